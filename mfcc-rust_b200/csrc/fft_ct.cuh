// Compile-time FFT building blocks for register-resident transforms.
//
// Everything here is resolved at compile time: loop indices (sfor), twiddle factors (constexpr
// sin/cos evaluated in double precision and emitted as FFMA immediates) and the data movement of
// the Stockham stages (pure register renaming after full unrolling).
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

namespace ssfft {

template <int I, int E, class F>
__device__ __forceinline__ void sfor(F&& f) {
  if constexpr (I < E) {
    f(std::integral_constant<int, I>{});
    sfor<I + 1, E>(static_cast<F&&>(f));
  }
}

constexpr double kPi = 3.14159265358979323846264338327950288;

// sin / cos on [-pi/4, pi/4] by Taylor series (truncation error < 1e-18)
constexpr double sin_core(double x) {
  const double x2 = x * x;
  double term = x, sum = x;
  for (int n = 1; n <= 11; ++n) {
    term *= -x2 / ((2.0 * n) * (2.0 * n + 1.0));
    sum += term;
  }
  return sum;
}
constexpr double cos_core(double x) {
  const double x2 = x * x;
  double term = 1.0, sum = 1.0;
  for (int n = 1; n <= 11; ++n) {
    term *= -x2 / ((2.0 * n - 1.0) * (2.0 * n));
    sum += term;
  }
  return sum;
}
// cos / sin of 2 pi m / n.  The octant is found on the integers, so the exact cases
// (multiples of pi/4) come out exact and everything else is accurate to ~1e-16.
constexpr double cos2pi(long long m, long long n);
constexpr double sin2pi(long long m, long long n) { return cos2pi(4 * m - n, 4 * n); }  // sin x = cos(x - pi/2)
constexpr double cos2pi(long long m, long long n) {
  m %= n;
  if (m < 0) m += n;                       // 0 <= m < n
  if (2 * m > n) m = n - m;                // cos(2pi - x) = cos x        -> 0 <= x <= pi
  bool neg = false;
  if (4 * m > n) { m = n - 2 * m; n *= 2; neg = true; if (m < 0) m = -m; }  // cos x = -cos(pi - x), x' = 2pi (n-2m)/(2n)
  // now 0 <= 2 pi m / n <= pi/2
  double v = 0.0;
  if (8 * m <= n) v = cos_core(2.0 * kPi * (double)m / (double)n);
  else v = sin_core(2.0 * kPi * (double)(n - 4 * m) / (double)(4 * n));   // cos x = sin(pi/2 - x)
  if (m == 0) v = 1.0;
  if (4 * m == n) v = 0.0;
  return neg ? -v : v;
}

// ---- packed complex arithmetic --------------------------------------------------------------
// sm_100 executes two FP32 operations per lane in one instruction on a 64-bit register pair
// (FADD2 / FMUL2 / FFMA2; PTX add/mul/fma.rn.f32x2), with per-operand swap (.LO_HI), broadcast and
// sign-pattern modifiers that ptxas folds in.  A complex number is exactly such a pair, so every
// butterfly below is written on whole float2 values: a complex add is one instruction, a twiddled
// radix-2 butterfly is three.  Each half is an ordinary IEEE fused multiply-add / add, so results
// are bit-identical to the scalar formulation.
__device__ __forceinline__ float2 cneg(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, cneg(b)); }
// a - i b and a + i b
__device__ __forceinline__ float2 cadd_mi(float2 a, float2 b) { return __fadd2_rn(a, make_float2(b.y, -b.x)); }
__device__ __forceinline__ float2 cadd_pi(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.y, b.x)); }
// a + (c - i s) b
__device__ __forceinline__ float2 cfma_tw(float c, float s, float2 b, float2 a) {
  // the sign pattern sits on the data operand so that c and s both stay broadcast immediates
  const float2 t = __ffma2_rn(make_float2(c, c), b, a);
  return __ffma2_rn(make_float2(s, s), make_float2(b.y, -b.x), t);
}
// (c - i s) v
__device__ __forceinline__ float2 cmul_tw(float c, float s, float2 v) {
  const float2 r = __fmul2_rn(make_float2(s, s), make_float2(v.y, -v.x));
  return __ffma2_rn(make_float2(c, c), v, r);
}
// a b for two run-time complex numbers
__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
  // b.x a + b.y (i a): the broadcast operands carry no sign, the swap / sign pattern sits on the pair a
  const float2 r = __fmul2_rn(make_float2(b.x, b.x), a);
  return __ffma2_rn(make_float2(b.y, b.y), make_float2(-a.y, a.x), r);
}

// radix-2 DIT butterfly with w = e^{-2 pi i M / N}: a' = a + w b, b' = a - w b = 2a - a'
template <int N, int M>
__device__ __forceinline__ void bfly2(float2& a, float2& b) {
  constexpr int m = ((M % N) + N) % N;
  if constexpr (m == 0) {
    const float2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
  } else if constexpr (4 * m == N) {  // w = -i : w b = (b.y, -b.x)
    const float2 t = a, s = b;
    a = cadd_mi(t, s);
    b = cadd_pi(t, s);
  } else if constexpr (2 * m == N) {  // w = -1
    const float2 t = a;
    a = csub(t, b);
    b = cadd(t, b);
  } else if constexpr (4 * m == 3 * N) {  // w = +i : w b = (-b.y, b.x)
    const float2 t = a, s = b;
    a = cadd_pi(t, s);
    b = cadd_mi(t, s);
  } else {
    constexpr float c = (float)cos2pi(m, N), s = (float)sin2pi(m, N);  // w = c - i s
    const float2 t = cfma_tw(c, s, b, a);
    b = __ffma2_rn(make_float2(2.0f, 2.0f), a, cneg(t));
    a = t;
  }
}

// multiply by e^{-2 pi i M / N}, compile-time
template <int N, int M>
__device__ __forceinline__ float2 twmul(float2 v) {
  constexpr int m = ((M % N) + N) % N;
  if constexpr (m == 0) return v;
  else if constexpr (4 * m == N) return make_float2(v.y, -v.x);
  else if constexpr (2 * m == N) return make_float2(-v.x, -v.y);
  else if constexpr (4 * m == 3 * N) return make_float2(-v.y, v.x);
  else {
    constexpr float c = (float)cos2pi(m, N), s = (float)sin2pi(m, N);
    return cmul_tw(c, s, v);
  }
}

// 5-point DFT, y_k = sum_n x_n e^{-2 pi i nk/5} (18 packed FP32 instructions)
__device__ __forceinline__ void dft5(float2& x0, float2& x1, float2& x2, float2& x3, float2& x4) {
  constexpr float c1 = (float)cos2pi(1, 5), c2 = (float)cos2pi(2, 5);
  constexpr float s1 = (float)sin2pi(1, 5), s2 = (float)sin2pi(2, 5);
  const float2 t1 = cadd(x1, x4), t2 = cadd(x2, x3);
  const float2 t3 = csub(x1, x4), t4 = csub(x2, x3);
  const float2 m1 = __ffma2_rn(make_float2(c2, c2), t2, __ffma2_rn(make_float2(c1, c1), t1, x0));
  const float2 m2 = __ffma2_rn(make_float2(c1, c1), t2, __ffma2_rn(make_float2(c2, c2), t1, x0));
  const float2 u1 = __ffma2_rn(make_float2(s2, s2), t4, __fmul2_rn(make_float2(s1, s1), t3));
  const float2 u2 = __ffma2_rn(make_float2(-s1, -s1), t4, __fmul2_rn(make_float2(s2, s2), t3));
  x0 = cadd(x0, cadd(t1, t2));
  x1 = cadd_mi(m1, u1);  // m1 - i u1
  x4 = cadd_pi(m1, u1);
  x2 = cadd_mi(m2, u2);
  x3 = cadd_pi(m2, u2);
}

// One Stockham stage on a register array: N points, radix R (2 or 5), sub-length NS.
//   out[(j / NS) * NS * R + k + r * NS] = DFT_R( in[j + r * N/R] * W_{NS*R}^{k r} ),  k = j % NS
// JNZ: in[j] is a structural zero for j >= JNZ (zero padding); only used by the first stage.
template <int N, int R, int NS, int JNZ = N>
__device__ __forceinline__ void stockham_stage(const float2 (&in)[N], float2 (&out)[N]) {
  static_assert(R == 2 || R == 5, "radix 2 or 5");
  constexpr int NB = N / R;
  sfor<0, NB>([&](auto jj) {
    constexpr int j = decltype(jj)::value;
    constexpr int k = j % NS;
    constexpr int o = (j / NS) * NS * R + k;
    if constexpr (R == 2 && NS == 1 && j + NB >= JNZ) {
      out[o] = in[j];  // a + 0, a - 0
      out[o + NS] = in[j];
    } else if constexpr (R == 2) {
      float2 a = in[j], b = in[j + NB];
      bfly2<NS * 2, k>(a, b);
      out[o] = a;
      out[o + NS] = b;
    } else {
      float2 v0 = in[j];
      float2 v1 = twmul<NS * 5, k>(in[j + NB]);
      float2 v2 = twmul<NS * 5, 2 * k>(in[j + 2 * NB]);
      float2 v3 = twmul<NS * 5, 3 * k>(in[j + 3 * NB]);
      float2 v4 = twmul<NS * 5, 4 * k>(in[j + 4 * NB]);
      dft5(v0, v1, v2, v3, v4);
      out[o] = v0;
      out[o + NS] = v1;
      out[o + 2 * NS] = v2;
      out[o + 3 * NS] = v3;
      out[o + 4 * NS] = v4;
    }
  });
}

// y[k] = sum_j x[j] e^{-2 pi i jk/N}, natural order in and out, for N in {5, 10, 20, 40}: a radix-5
// stage without twiddles followed by radix-2 stages with 6-FFMA butterflies.
template <int N>
__device__ __forceinline__ void fft_5x2(float2 (&x)[N]) {
  static_assert(N == 5 || N == 10 || N == 20 || N == 40, "5 * 2^a");
  float2 t[N];
  stockham_stage<N, 5, 1>(x, t);
  if constexpr (N == 5) {
    sfor<0, N>([&](auto i) { x[decltype(i)::value] = t[decltype(i)::value]; });
  } else {
    stockham_stage<N, 2, 5>(t, x);
    if constexpr (N >= 20) {
      stockham_stage<N, 2, 10>(x, t);
      if constexpr (N == 40) stockham_stage<N, 2, 20>(t, x);
      else sfor<0, N>([&](auto i) { x[decltype(i)::value] = t[decltype(i)::value]; });
    }
  }
}

// y[k] = sum_j x[j] e^{-2 pi i jk/N} for N = 8, 16 or 32: radix-2 Stockham stages in registers
template <int N, int JNZ = N>
__device__ __forceinline__ void fft_pow2(float2 (&x)[N]) {
  static_assert(N == 8 || N == 16 || N == 32, "8, 16 or 32");
  float2 t[N];
  stockham_stage<N, 2, 1, JNZ>(x, t);
  stockham_stage<N, 2, 2>(t, x);
  stockham_stage<N, 2, 4>(x, t);
  if constexpr (N == 8) {
    sfor<0, N>([&](auto i) { x[decltype(i)::value] = t[decltype(i)::value]; });
  } else {
    stockham_stage<N, 2, 8>(t, x);
    if constexpr (N == 32) {
      stockham_stage<N, 2, 16>(x, t);
      sfor<0, N>([&](auto i) { x[decltype(i)::value] = t[decltype(i)::value]; });
    }
  }
}

}  // namespace ssfft
