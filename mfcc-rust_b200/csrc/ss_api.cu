// C-ABI host layer: plan management, argument validation (the reference's panics become
// error codes), sharding of clips over the plan's devices, and the chunked
// H2D -> kernels -> D2H pipeline for host buffers.  See include/speechsauce_cuda.h.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>

#include "ss_internal.h"

namespace ss {

static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
static std::atomic<const char*> g_last_kernel{""};  // process-wide: multi-device plans launch from worker threads
void note_kernel(const char* name) { g_last_kernel.store(name, std::memory_order_relaxed); }

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
static int cuda_fail(cudaError_t e, const char* what) {
  return fail(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? SS_ERR_NO_DEVICE : SS_ERR_CUDA,
              "%s: %s", what, cudaGetErrorString(e));
}
#define SS_CUDA(call)                                        \
  do {                                                       \
    cudaError_t e__ = (call);                                \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call);    \
  } while (0)

constexpr int kLanes = 3;  // H2D / compute / D2H of consecutive chunks overlap on 3 streams

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return SS_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      e = cudaMalloc(&p, bytes);
      want = bytes;
    }
    if (e != cudaSuccess) return fail(SS_ERR_ALLOC, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    cap = want;
    return SS_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct Lane {
  cudaStream_t stream = nullptr;
  DevBuf in, out, scratch, mem;
};

struct DeviceCtx {
  int device = 0;
  int sm_count = 148;
  DevBuf tables;  // one allocation holding every device table
  DevTables tab;
  Fast512bTables* fast512b = nullptr;
  Fast512cTables* fast512c = nullptr;
  Stft400Tables* stft400 = nullptr;
  Stft400mTables* stft400m = nullptr;
  Stft2048Tables* stft2048 = nullptr;
  Stft2048mTables* stft2048m = nullptr;
  Lane lanes[kLanes];
  DevBuf dev_scratch;  // scratch for the *_dev entry points (cmvn partials)
  std::string err;     // worker-thread error message
};

}  // namespace ss

struct ss_plan {
  ss::Derived d;
  ss::MelBank mb;
  ss::FftPlanHost fft;
  bool fft_ok = false;
  std::vector<float> window, dct;
  std::vector<float> analysis_mem;  // [W-H], config.rs:126,162
  std::vector<ss::DeviceCtx> devs;
  bool fast512b = false;
  bool fast512c = false;
  bool stft400 = false;
  bool stft400m = false;
  bool stft2048 = false;
  bool stft2048m = false;
};

namespace ss {

static int upload_tables(ss_plan* pl, DeviceCtx* dc) {
  SS_CUDA(cudaSetDevice(dc->device));
  cudaDeviceProp prop;
  SS_CUDA(cudaGetDeviceProperties(&prop, dc->device));
  dc->sm_count = prop.multiProcessorCount;
  const Derived& d = pl->d;
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t o_tw = 0;
  const size_t o_win = align(o_tw + (pl->fft_ok ? pl->fft.tw.size() : 2) * sizeof(float));
  const size_t o_fs = align(o_win + pl->window.size() * sizeof(float));
  const size_t o_fl = align(o_fs + d.nf * sizeof(int));
  const size_t o_fo = align(o_fl + d.nf * sizeof(int));
  const size_t o_fw = align(o_fo + d.nf * sizeof(int));
  const size_t o_dct = align(o_fw + pl->mb.w.size() * sizeof(float));
  const size_t total = align(o_dct + std::max<size_t>(1, pl->dct.size()) * sizeof(float));
  int rc = dc->tables.reserve(total);
  if (rc) return rc;
  char* base = static_cast<char*>(dc->tables.p);
  if (pl->fft_ok)
    SS_CUDA(cudaMemcpy(base + o_tw, pl->fft.tw.data(), pl->fft.tw.size() * sizeof(float), cudaMemcpyHostToDevice));
  SS_CUDA(cudaMemcpy(base + o_win, pl->window.data(), pl->window.size() * sizeof(float), cudaMemcpyHostToDevice));
  SS_CUDA(cudaMemcpy(base + o_fs, pl->mb.start.data(), d.nf * sizeof(int), cudaMemcpyHostToDevice));
  SS_CUDA(cudaMemcpy(base + o_fl, pl->mb.len.data(), d.nf * sizeof(int), cudaMemcpyHostToDevice));
  SS_CUDA(cudaMemcpy(base + o_fo, pl->mb.wofs.data(), d.nf * sizeof(int), cudaMemcpyHostToDevice));
  SS_CUDA(cudaMemcpy(base + o_fw, pl->mb.w.data(), pl->mb.w.size() * sizeof(float), cudaMemcpyHostToDevice));
  if (!pl->dct.empty())
    SS_CUDA(cudaMemcpy(base + o_dct, pl->dct.data(), pl->dct.size() * sizeof(float), cudaMemcpyHostToDevice));
  dc->tab.tw = reinterpret_cast<const float2*>(base + o_tw);
  dc->tab.window = reinterpret_cast<const float*>(base + o_win);
  dc->tab.fstart = reinterpret_cast<const int*>(base + o_fs);
  dc->tab.flen = reinterpret_cast<const int*>(base + o_fl);
  dc->tab.fwofs = reinterpret_cast<const int*>(base + o_fo);
  dc->tab.fw = reinterpret_cast<const float*>(base + o_fw);
  dc->tab.dct = reinterpret_cast<const float*>(base + o_dct);
  dc->tab.n_stages = pl->fft.n_stages;
  for (int i = 0; i < kMaxStages; ++i) dc->tab.radix[i] = pl->fft.radix[i];
  for (auto& ln : dc->lanes) SS_CUDA(cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking));
  if (pl->fast512b) {
    int e = fast512b_build_tables(pl->d, pl->mb, &dc->fast512b);
    if (e) return cuda_fail((cudaError_t)e, "fast512b_build_tables");
  }
  if (pl->fast512c) {
    int e = fast512c_build_tables(pl->d, pl->mb, &dc->fast512c);
    if (e) return cuda_fail((cudaError_t)e, "fast512c_build_tables");
  }
  if (pl->stft400) {
    int e = stft400_build_tables(pl->d, pl->mb, pl->window, &dc->stft400);
    if (e) return cuda_fail((cudaError_t)e, "stft400_build_tables");
  }
  if (pl->stft400m) {
    int e = stft400m_build_tables(pl->d, pl->mb, pl->window, &dc->stft400m);
    if (e) return cuda_fail((cudaError_t)e, "stft400m_build_tables");
  }
  if (pl->stft2048) {
    int e = stft2048_build_tables(pl->d, pl->mb, pl->window, &dc->stft2048);
    if (e) return cuda_fail((cudaError_t)e, "stft2048_build_tables");
  }
  if (pl->stft2048m) {
    int e = stft2048m_build_tables(pl->d, pl->mb, pl->window, &dc->stft2048m);
    if (e) return cuda_fail((cudaError_t)e, "stft2048m_build_tables");
  }
  return SS_OK;
}

// ---- speechpy-framing features on one device, device pointers ---------------------------
static int features_dev(ss_plan* pl, DeviceCtx* dc, cudaStream_t stream, const float* d_pcm, int64_t n_clips,
                        int64_t clip_len, int64_t row_stride, const ss_feature_outputs& o, DevBuf* scratch) {
  const Derived& d = pl->d;
  const int64_t T = num_frames(clip_len, d.L, d.S, 0);
  FrameParams p;
  p.pcm = d_pcm;
  p.n_rows = n_clips;
  p.row_len = clip_len;
  p.row_stride = row_stride;
  p.T = T;
  p.T_out = T;
  p.N = (int)d.W;
  p.F = (int)d.F;
  p.L = (int)d.L;
  p.S = (int)d.S;
  p.nf = d.nf;
  p.nc = d.nc;
  p.mode = kModeSpeechpy;
  p.flags = d.cfg.flags;
  p.mag_scale = 1.0f / (float)d.W;  // processing.rs:180
  const float n = (float)(T * (int64_t)d.nf);  // feature.rs:126
  p.scale_00 = 1.0f / std::sqrt(4.0f * n);
  p.scale_rest = 1.0f / std::sqrt(2.0f * n);
  p.dc_elim = d.cfg.dc_elimination ? 1 : 0;
  p.ssc_r0 = 1.0f;
  p.ssc_step = d.F > 1 ? ((float)d.cfg.sample_rate / 2.0f - 1.0f) / (float)(d.F - 1) : 0.f;
  p.mfe = o.mfe;
  p.energy = o.energy;
  p.lmfe = o.lmfe;
  p.ssc = o.ssc;
  p.mfcc = o.mfcc;
  p.cmvn_mode = o.mfcc ? o.mfcc_cmvn_mode : 0;
  p.tab = dc->tab;
  bool cmvn_done = false;
  int e;
  const bool use_fast = (pl->fast512b || pl->fast512c) && !(d.cfg.flags & (SS_FLAG_FORCE_GENERIC | SS_FLAG_FRAMING_LITERAL));
  // SS_FAST512_GEN=2 selects the previous generation of the n_fft = 512 kernel for A/B measurements
  // (2: CUDA-core mel program + DCT, CTA per clip; default 3: tensor-core mel + DCT, warp per item)
  static const int gen = getenv("SS_FAST512_GEN") ? atoi(getenv("SS_FAST512_GEN")) : 3;
  if (use_fast) {
    if (pl->fast512c && (gen >= 3 || !pl->fast512b)) e = launch_fast512c(p, dc->fast512c, dc->sm_count, stream);
    else e = launch_fast512b(p, dc->fast512b, dc->sm_count, stream);
    cmvn_done = true;  // the fast path normalises in-kernel or enqueues its own cmvn pass
    if (e == -1000) {  // the fast path declined this shape; fall through to the generic kernel
      cmvn_done = false;
      e = launch_frames_generic(p, stream);
    } else if (e == -1001) {  // features written, cmvn left to us
      cmvn_done = false;
      e = 0;
    }
  } else {
    e = launch_frames_generic(p, stream);
  }
  if (e) return cuda_fail((cudaError_t)e, "feature kernel launch");
  if (p.cmvn_mode && !cmvn_done) {
    int rc = scratch->reserve(cmvn_scratch_bytes(n_clips, T, d.nc));
    if (rc) return rc;
    e = launch_cmvn(o.mfcc, n_clips, T, d.nc, p.cmvn_mode == SS_CMVN_MEAN_VAR, o.mfcc,
                    static_cast<float*>(scratch->p), stream);
    if (e) return cuda_fail((cudaError_t)e, "cmvn kernel launch");
  }
  return SS_OK;
}

static void stft_params(ss_plan* pl, DeviceCtx* dc, FrameParams* p, const float* d_pcm, int64_t ch, int64_t ttd,
                        int64_t row_stride) {
  const Derived& d = pl->d;
  p->pcm = d_pcm;
  p->n_rows = ch;
  p->row_len = ttd;
  p->row_stride = row_stride;
  p->Tp = stft_num_frames(ttd, d.H);
  p->n_pad = (int)d.n_pad;
  p->T = std::max<int64_t>(0, p->Tp - d.n_pad);
  p->T_out = p->Tp;
  p->N = (int)d.W;
  p->F = (int)d.F;
  p->L = (int)d.W;
  p->S = (int)d.H;
  p->nf = d.nf;
  p->nc = d.nc;
  p->mode = kModeStft;
  p->wnorm = d.wnorm;
  p->tab = dc->tab;
}

// mel_spectrogram / stft launch: a size-specialised kernel when one exists, else the generic one
static int launch_stft_any(ss_plan* pl, DeviceCtx* dc, const FrameParams& p, void* stream) {
  // SS_STFT400_GEN=1 selects the first-generation fft_points = 400 kernel for A/B measurements
  static const int gen400 = getenv("SS_STFT400_GEN") ? atoi(getenv("SS_STFT400_GEN")) : 2;
  if (pl->stft400m && dc->stft400m && gen400 >= 2 && (p.mel_out || p.stft_out) && !(pl->d.cfg.flags & SS_FLAG_FORCE_GENERIC)) {
    const int e = launch_stft400m(p, dc->stft400m, dc->sm_count, stream);
    if (e != -1000) return e;
  }
  if (pl->stft400 && dc->stft400 && p.mel_out && !p.stft_out && !(pl->d.cfg.flags & SS_FLAG_FORCE_GENERIC)) {
    const int e = launch_stft400(p, dc->stft400, dc->sm_count, stream);
    if (e != -1000) return e;
  }
  // SS_STFT2048_GEN=1 selects the first-generation fft_points = 2048 mel kernel for A/B measurements
  static const int gen2048 = getenv("SS_STFT2048_GEN") ? atoi(getenv("SS_STFT2048_GEN")) : 2;
  if (pl->stft2048m && dc->stft2048m && gen2048 >= 2 && p.mel_out && !p.stft_out && !(pl->d.cfg.flags & SS_FLAG_FORCE_GENERIC)) {
    const int e = launch_stft2048m(p, dc->stft2048m, dc->sm_count, stream);
    if (e != -1000) return e;
  }
  if (pl->stft2048 && dc->stft2048 && (p.mel_out || p.stft_out) && !(pl->d.cfg.flags & SS_FLAG_FORCE_GENERIC)) {
    const int e = launch_stft2048(p, dc->stft2048, dc->sm_count, stream);
    if (e != -1000) return e;
  }
  return launch_frames_generic(p, stream);
}

static int check_speechpy(ss_plan* pl, int64_t clip_len, int64_t row_stride, int64_t* T) {
  if (!pl) return fail(SS_ERR_INVALID_ARG, "null plan");
  const Derived& d = pl->d;
  if (d.librosa) return fail(SS_ERR_INVALID_ARG, "this plan was created with SS_FLAG_LIBROSA: use ss_librosa_melspectrogram");
  if (!d.speechpy_ok)
    return fail(SS_ERR_INVALID_ARG,
                "framing undefined for this config: need 0 < round(sr*frame_length)=%lld <= fft_points=%lld and "
                "round(sr*frame_stride)=%lld > 0 (ndrustfft panics when the frame is longer than the FFT)",
                (long long)d.L, (long long)d.W, (long long)d.S);
  if (!pl->fft_ok) return fail(SS_ERR_UNSUPPORTED, "fft_points=%lld has a prime factor > 61", (long long)d.W);
  if (d.W > 8192) return fail(SS_ERR_UNSUPPORTED, "fft_points=%lld > 8192", (long long)d.W);
  if (row_stride < clip_len) return fail(SS_ERR_INVALID_ARG, "row_stride < clip_len");
  *T = num_frames(clip_len, d.L, d.S, 0);
  if (*T < 1)
    return fail(SS_ERR_INVALID_ARG,
                "signal of %lld samples yields no frame (L=%lld, S=%lld): usize underflow in the reference "
                "(processing.rs:101-105)",
                (long long)clip_len, (long long)d.L, (long long)d.S);
  return SS_OK;
}

static int check_stft(ss_plan* pl, int64_t ch, int64_t ttd, int64_t row_stride) {
  if (!pl) return fail(SS_ERR_INVALID_ARG, "null plan");
  const Derived& d = pl->d;
  if (d.librosa) return fail(SS_ERR_INVALID_ARG, "this plan was created with SS_FLAG_LIBROSA: use ss_librosa_melspectrogram");
  if (!d.stft_ok)
    return fail(SS_ERR_INVALID_ARG,
                "stft undefined for this config: need 0 < frame_size=%lld <= fft_points/2=%lld (usize underflow "
                "at functions.rs:136 in the reference)",
                (long long)d.H, (long long)(d.W / 2));
  if (!pl->fft_ok) return fail(SS_ERR_UNSUPPORTED, "fft_points=%lld has a prime factor > 61", (long long)d.W);
  if (d.W > 8192) return fail(SS_ERR_UNSUPPORTED, "fft_points=%lld > 8192", (long long)d.W);
  if (ch < 0 || ttd < 0 || row_stride < ttd) return fail(SS_ERR_INVALID_ARG, "bad stft shape");
  return SS_OK;
}

// host-side restatement of the overlap memory in front of row `r` (functions.rs:153-160):
// the last W-H samples of  mem0 ++ pad(row 0) ++ ... ++ pad(row r-1)
static void host_mem_before_row(const ss_plan* pl, const float* pcm, int64_t ttd, int64_t row_stride, int64_t r,
                                const float* mem0, float* out) {
  const Derived& d = pl->d;
  const int64_t memlen = d.W - d.H;
  const int64_t span = stft_num_frames(ttd, d.H) * d.H;
  const int64_t total = memlen + r * span;
  for (int64_t i = 0; i < memlen; ++i) {
    int64_t q = total - memlen + i;
    float v;
    if (q < memlen) v = mem0 ? mem0[q] : 0.f;
    else {
      q -= memlen;
      const int64_t rr = span > 0 ? q / span : 0, o = span > 0 ? q % span : 0;
      v = o < ttd ? pcm[rr * row_stride + o] : 0.f;
    }
    out[i] = v;
  }
}

template <class Fn>
static int run_on_devices(ss_plan* pl, int64_t n_items, Fn fn) {
  const int nd = (int)pl->devs.size();
  if (nd == 0) return fail(SS_ERR_NO_DEVICE, "plan has no CUDA device (host-only plan); there is no CPU fallback");
  std::vector<int> rc(nd, 0);
  if (nd == 1) {
    rc[0] = fn(&pl->devs[0], (int64_t)0, n_items);
    if (rc[0]) pl->devs[0].err = g_err;
  } else {
    std::vector<std::thread> th;
    for (int i = 0; i < nd; ++i) {
      const int64_t a = n_items * i / nd, b = n_items * (i + 1) / nd;
      th.emplace_back([&, i, a, b] {
        rc[i] = (b > a) ? fn(&pl->devs[i], a, b) : 0;
        if (rc[i]) pl->devs[i].err = g_err;
      });
    }
    for (auto& t : th) t.join();
  }
  for (int i = 0; i < nd; ++i)
    if (rc[i]) {
      g_err = "device " + std::to_string(pl->devs[i].device) + ": " + pl->devs[i].err;
      return rc[i];
    }
  return SS_OK;
}

static int copy_rows_h2d(float* dst, const float* src, int64_t rows, int64_t len, int64_t row_stride,
                         cudaStream_t s) {
  if (rows == 0 || len == 0) return SS_OK;
  if (row_stride == len) {
    SS_CUDA(cudaMemcpyAsync(dst, src, (size_t)rows * len * sizeof(float), cudaMemcpyHostToDevice, s));
  } else {
    SS_CUDA(cudaMemcpy2DAsync(dst, len * sizeof(float), src, row_stride * sizeof(float), len * sizeof(float),
                              rows, cudaMemcpyHostToDevice, s));
  }
  return SS_OK;
}

static int sync_lanes(DeviceCtx* dc) {
  for (auto& ln : dc->lanes) SS_CUDA(cudaStreamSynchronize(ln.stream));
  return SS_OK;
}

// ---- default context for the config-less processing functions ---------------------------
struct DefaultCtx {
  cudaStream_t stream = nullptr;
  DevBuf a, b, c;
  bool ready = false;
};
static thread_local DefaultCtx g_ctx;
static int default_ctx(DefaultCtx** out) {
  if (!g_ctx.ready) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
      return fail(SS_ERR_NO_DEVICE, "no CUDA device visible (%s); there is no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    SS_CUDA(cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking));
    g_ctx.ready = true;
  }
  *out = &g_ctx;
  return SS_OK;
}

// host buffers -> device -> fn -> host buffers on the default context
template <class Fn>
static int host_unary(const float* in, size_t in_floats, float* out, size_t out_floats, Fn fn) {
  if (!in || !out) return fail(SS_ERR_INVALID_ARG, "null argument");
  DefaultCtx* c;
  int rc = default_ctx(&c);
  if (rc) return rc;
  if ((rc = c->a.reserve(std::max<size_t>(16, in_floats * 4)))) return rc;
  if ((rc = c->b.reserve(std::max<size_t>(16, out_floats * 4)))) return rc;
  if (in_floats) SS_CUDA(cudaMemcpyAsync(c->a.p, in, in_floats * 4, cudaMemcpyHostToDevice, c->stream));
  if ((rc = fn(c->stream, static_cast<float*>(c->a.p), static_cast<float*>(c->b.p)))) return rc;
  if (out_floats) SS_CUDA(cudaMemcpyAsync(out, c->b.p, out_floats * 4, cudaMemcpyDeviceToHost, c->stream));
  SS_CUDA(cudaStreamSynchronize(c->stream));
  return SS_OK;
}
}  // namespace ss

using namespace ss;

// ======================================================================= exported symbols
#pragma GCC visibility push(default)
extern "C" {

const char* ss_last_error(void) { return g_err.c_str(); }
const char* ss_version(void) { return "speechsauce-cuda 0.1.0 (sm_100a)"; }
int64_t ss_kernel_launch_count(void) { return g_launches.load(); }
const char* ss_last_kernel_name(void) { return g_last_kernel.load(std::memory_order_relaxed); }

int ss_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int ss_host_alloc(void** ptr, size_t bytes) {
  if (!ptr) return fail(SS_ERR_INVALID_ARG, "null ptr");
  cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocPortable);
  if (e != cudaSuccess) return cuda_fail(e, "cudaHostAlloc");
  return SS_OK;
}
int ss_host_free(void* ptr) {
  if (!ptr) return SS_OK;
  SS_CUDA(cudaFreeHost(ptr));
  return SS_OK;
}

int64_t ss_num_frames(int64_t n, int64_t frame_len, int64_t frame_step, int zero_padding) {
  return num_frames(n, frame_len, frame_step, zero_padding);
}
int64_t ss_stft_num_frames(int64_t ttd, int64_t hop) { return stft_num_frames(ttd, hop); }

int ss_plan_create(const ss_config* cfg, const int* devices, int n_devices, ss_plan** out) {
  if (!cfg || !out) return fail(SS_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  std::unique_ptr<ss_plan> pl(new ss_plan);
  std::string err = derive(*cfg, &pl->d);
  if (!err.empty()) return fail(SS_ERR_INVALID_ARG, "%s", err.c_str());
  build_melbank(pl->d, &pl->mb);
  build_window(pl->d, &pl->window);
  build_dct(pl->d, &pl->dct);
  pl->fft_ok = pl->d.W <= (1 << 20) && build_fft_plan((int)pl->d.W, &pl->fft);
  if (pl->d.stft_ok) pl->analysis_mem.assign((size_t)(pl->d.W - pl->d.H), 0.f);  // config.rs:162
  pl->fast512b = pl->fft_ok && fast512b_supported(pl->d, pl->mb);
  pl->fast512c = pl->fft_ok && fast512c_supported(pl->d, pl->mb);
  pl->stft400 = pl->fft_ok && stft400_supported(pl->d, pl->mb);
  pl->stft400m = pl->fft_ok && stft400m_supported(pl->d, pl->mb);
  pl->stft2048 = pl->fft_ok && stft2048_supported(pl->d, pl->mb);
  pl->stft2048m = pl->fft_ok && stft2048m_supported(pl->d, pl->mb);
  std::vector<int> devs;
  if (n_devices < 0) {
    int n = ss_device_count();
    if (n == 0) return fail(SS_ERR_NO_DEVICE, "no CUDA device visible; there is no CPU fallback");
    for (int i = 0; i < n; ++i) devs.push_back(i);
  } else {
    for (int i = 0; i < n_devices; ++i) devs.push_back(devices ? devices[i] : i);
  }
  if (!devs.empty()) {
    int n = ss_device_count();
    if (n == 0) return fail(SS_ERR_NO_DEVICE, "no CUDA device visible; there is no CPU fallback");
    for (int dv : devs)
      if (dv < 0 || dv >= n) return fail(SS_ERR_INVALID_ARG, "device %d out of range (0..%d)", dv, n - 1);
  }
  pl->devs.resize(devs.size());
  int prev = -1;
  if (!devs.empty()) cudaGetDevice(&prev);
  for (size_t i = 0; i < devs.size(); ++i) {
    pl->devs[i].device = devs[i];
    int rc = upload_tables(pl.get(), &pl->devs[i]);
    if (rc) {
      ss_plan_destroy(pl.release());
      return rc;
    }
  }
  if (prev >= 0) cudaSetDevice(prev);
  *out = pl.release();
  return SS_OK;
}

void ss_plan_destroy(ss_plan* pl) {
  if (!pl) return;
  int prev = -1;
  if (!pl->devs.empty()) cudaGetDevice(&prev);
  for (auto& dc : pl->devs) {
    cudaSetDevice(dc.device);
    for (auto& ln : dc.lanes) {
      if (ln.stream) {
        cudaStreamSynchronize(ln.stream);
        cudaStreamDestroy(ln.stream);
      }
      ln.in.release();
      ln.out.release();
      ln.scratch.release();
      ln.mem.release();
    }
    dc.tables.release();
    dc.dev_scratch.release();
    if (dc.fast512b) fast512b_free_tables(dc.fast512b);
    if (dc.fast512c) fast512c_free_tables(dc.fast512c);
    if (dc.stft400) stft400_free_tables(dc.stft400);
    if (dc.stft400m) stft400m_free_tables(dc.stft400m);
    if (dc.stft2048) stft2048_free_tables(dc.stft2048);
    if (dc.stft2048m) stft2048m_free_tables(dc.stft2048m);
  }
  if (prev >= 0) cudaSetDevice(prev);
  delete pl;
}

int ss_plan_get_info(const ss_plan* pl, ss_plan_info* info) {
  if (!pl || !info) return fail(SS_ERR_INVALID_ARG, "null argument");
  const Derived& d = pl->d;
  info->frame_len = d.L;
  info->frame_step = d.S;
  info->hop = d.H;
  info->n_fft = d.W;
  info->freq_size = d.F;
  info->n_pad = d.n_pad;
  info->num_filters = d.nf;
  info->num_cepstral = d.nc;
  info->wnorm = d.wnorm;
  info->n_devices = (int32_t)pl->devs.size();
  info->fast_path = (pl->fast512b || pl->fast512c) && !(d.cfg.flags & (SS_FLAG_FORCE_GENERIC | SS_FLAG_FRAMING_LITERAL));
  info->fast_stft = (pl->stft400 || pl->stft400m || pl->stft2048 || pl->stft2048m) && !(d.cfg.flags & SS_FLAG_FORCE_GENERIC);
  return SS_OK;
}

int ss_plan_get_analysis_mem(const ss_plan* pl, float* out, int64_t n) {
  if (!pl || !out) return fail(SS_ERR_INVALID_ARG, "null argument");
  if (n != (int64_t)pl->analysis_mem.size()) return fail(SS_ERR_INVALID_ARG, "analysis_mem has %zu floats", pl->analysis_mem.size());
  std::memcpy(out, pl->analysis_mem.data(), n * sizeof(float));
  return SS_OK;
}
int ss_plan_set_analysis_mem(ss_plan* pl, const float* in, int64_t n) {
  if (!pl) return fail(SS_ERR_INVALID_ARG, "null argument");
  if (!in) {
    std::fill(pl->analysis_mem.begin(), pl->analysis_mem.end(), 0.f);
    return SS_OK;
  }
  if (n != (int64_t)pl->analysis_mem.size()) return fail(SS_ERR_INVALID_ARG, "analysis_mem has %zu floats", pl->analysis_mem.size());
  std::memcpy(pl->analysis_mem.data(), in, n * sizeof(float));
  return SS_OK;
}

int ss_debug_filterbank(const ss_plan* pl, float* out) {
  if (!pl || !out) return fail(SS_ERR_INVALID_ARG, "null argument");
  const Derived& d = pl->d;
  std::memset(out, 0, sizeof(float) * d.nf * d.F);
  for (int i = 0; i < d.nf; ++i)
    for (int k = 0; k < pl->mb.len[i]; ++k) out[(size_t)i * d.F + pl->mb.start[i] + k] = pl->mb.w[pl->mb.wofs[i] + k];
  return SS_OK;
}
int ss_debug_window(const ss_plan* pl, float* out) {
  if (!pl || !out) return fail(SS_ERR_INVALID_ARG, "null argument");
  std::memcpy(out, pl->window.data(), pl->window.size() * sizeof(float));
  return SS_OK;
}

// -------------------------------------------------------------------------- features (host)
int ss_features(ss_plan* pl, const float* pcm, int64_t n_clips, int64_t clip_len, int64_t row_stride,
                const ss_feature_outputs* outs) {
  if (!pl || !pcm || !outs) return fail(SS_ERR_INVALID_ARG, "null argument");
  if (n_clips < 0) return fail(SS_ERR_INVALID_ARG, "n_clips < 0");
  int64_t T = 0;
  int rc = check_speechpy(pl, clip_len, row_stride, &T);
  if (rc) return rc;
  if (n_clips == 0) return SS_OK;
  const Derived& d = pl->d;
  const ss_feature_outputs o = *outs;
  const size_t per_clip_out =
      (size_t)T * ((o.mfe ? d.nf : 0) + (o.energy ? 1 : 0) + (o.lmfe ? d.nf : 0) + (o.ssc ? d.nf : 0) + (o.mfcc ? d.nc : 0)) *
      sizeof(float);
  const size_t per_clip_in = (size_t)clip_len * sizeof(float);
  return run_on_devices(pl, n_clips, [&](DeviceCtx* dc, int64_t c0, int64_t c1) -> int {
    SS_CUDA(cudaSetDevice(dc->device));
    // chunk so that H2D, kernels and D2H of neighbouring chunks overlap; whole clips only
    int64_t chunk = std::max<int64_t>(1, (int64_t)((64u << 20) / std::max<size_t>(1, per_clip_in + per_clip_out)));
    chunk = std::min(chunk, c1 - c0);
    int li = 0;
    for (int64_t a = c0; a < c1; a += chunk, li = (li + 1) % kLanes) {
      const int64_t n = std::min(chunk, c1 - a);
      Lane& ln = dc->lanes[li];
      int r = ln.in.reserve(n * per_clip_in);
      if (r) return r;
      r = ln.out.reserve(std::max<size_t>(16, n * per_clip_out + 5 * 256));
      if (r) return r;
      r = copy_rows_h2d(static_cast<float*>(ln.in.p), pcm + a * row_stride, n, clip_len, row_stride, ln.stream);
      if (r) return r;
      ss_feature_outputs dv = o;
      char* ob = static_cast<char*>(ln.out.p);
      auto carve = [&](float* host, size_t elems) -> float* {
        if (!host) return nullptr;
        float* p = reinterpret_cast<float*>(ob);
        ob += (elems * sizeof(float) + 255) & ~(size_t)255;
        return p;
      };
      dv.mfe = carve(o.mfe, (size_t)n * T * d.nf);
      dv.energy = carve(o.energy, (size_t)n * T);
      dv.lmfe = carve(o.lmfe, (size_t)n * T * d.nf);
      dv.ssc = carve(o.ssc, (size_t)n * T * d.nf);
      dv.mfcc = carve(o.mfcc, (size_t)n * T * d.nc);
      r = features_dev(pl, dc, ln.stream, static_cast<float*>(ln.in.p), n, clip_len, clip_len, dv, &ln.scratch);
      if (r) return r;
      auto back = [&](float* host, float* dev, size_t per_clip) -> int {
        if (!host) return SS_OK;
        SS_CUDA(cudaMemcpyAsync(host + (size_t)a * per_clip, dev, (size_t)n * per_clip * sizeof(float),
                                cudaMemcpyDeviceToHost, ln.stream));
        return SS_OK;
      };
      if ((r = back(o.mfe, dv.mfe, (size_t)T * d.nf))) return r;
      if ((r = back(o.energy, dv.energy, (size_t)T))) return r;
      if ((r = back(o.lmfe, dv.lmfe, (size_t)T * d.nf))) return r;
      if ((r = back(o.ssc, dv.ssc, (size_t)T * d.nf))) return r;
      if ((r = back(o.mfcc, dv.mfcc, (size_t)T * d.nc))) return r;
    }
    return sync_lanes(dc);
  });
}

int ss_mfe(ss_plan* pl, const float* pcm, int64_t n, int64_t len, int64_t stride, float* feat, float* energy) {
  ss_feature_outputs o = {};
  o.mfe = feat;
  o.energy = energy;
  return ss_features(pl, pcm, n, len, stride, &o);
}
int ss_lmfe(ss_plan* pl, const float* pcm, int64_t n, int64_t len, int64_t stride, float* out) {
  ss_feature_outputs o = {};
  o.lmfe = out;
  return ss_features(pl, pcm, n, len, stride, &o);
}
int ss_mfcc(ss_plan* pl, const float* pcm, int64_t n, int64_t len, int64_t stride, float* out, int cmvn_mode) {
  if (cmvn_mode < 0 || cmvn_mode > 2) return fail(SS_ERR_INVALID_ARG, "cmvn_mode must be 0, 1 or 2");
  ss_feature_outputs o = {};
  o.mfcc = out;
  o.mfcc_cmvn_mode = cmvn_mode;
  return ss_features(pl, pcm, n, len, stride, &o);
}
int ss_ssc(ss_plan* pl, const float* pcm, int64_t n, int64_t len, int64_t stride, float* out) {
  ss_feature_outputs o = {};
  o.ssc = out;
  return ss_features(pl, pcm, n, len, stride, &o);
}

int ss_features_dev(ss_plan* pl, int dev_index, void* stream, const float* d_pcm, int64_t n_clips,
                    int64_t clip_len, int64_t row_stride, const ss_feature_outputs* d_outs) {
  if (!pl || !d_pcm || !d_outs) return fail(SS_ERR_INVALID_ARG, "null argument");
  if (dev_index < 0 || dev_index >= (int)pl->devs.size())
    return fail(pl->devs.empty() ? SS_ERR_NO_DEVICE : SS_ERR_INVALID_ARG, "dev_index %d out of range", dev_index);
  int64_t T = 0;
  int rc = check_speechpy(pl, clip_len, row_stride, &T);
  if (rc) return rc;
  if (n_clips <= 0) return n_clips == 0 ? SS_OK : fail(SS_ERR_INVALID_ARG, "n_clips < 0");
  DeviceCtx* dc = &pl->devs[dev_index];
  return features_dev(pl, dc, (cudaStream_t)stream, d_pcm, n_clips, clip_len, row_stride, *d_outs, &dc->dev_scratch);
}

// ------------------------------------------------------------------------------ STFT (host)
static int stft_host(ss_plan* pl, const float* pcm, int64_t ch, int64_t ttd, int64_t row_stride, float* out_ri,
                     float* out_mel, float* mem_inout, int log_mode) {
  int rc = check_stft(pl, ch, ttd, row_stride);
  if (rc) return rc;
  if (!pcm && ch * ttd > 0) return fail(SS_ERR_INVALID_ARG, "null pcm");
  const Derived& d = pl->d;
  const int64_t memlen = d.W - d.H;
  const int64_t Tp = stft_num_frames(ttd, d.H);
  std::vector<float> mem0(memlen);
  std::memcpy(mem0.data(), mem_inout ? mem_inout : pl->analysis_mem.data(), memlen * sizeof(float));
  if (ch > 0 && Tp > 0) {
    const size_t per_row_in = (size_t)ttd * sizeof(float);
    const size_t per_row_out = (size_t)Tp * (out_ri ? (size_t)d.F * 2 : (size_t)d.nf) * sizeof(float);
    rc = run_on_devices(pl, ch, [&](DeviceCtx* dc, int64_t c0, int64_t c1) -> int {
      SS_CUDA(cudaSetDevice(dc->device));
      int64_t chunk = std::max<int64_t>(1, (int64_t)((64u << 20) / std::max<size_t>(1, per_row_in + per_row_out)));
      chunk = std::min(chunk, c1 - c0);
      int li = 0;
      std::vector<float> hmem(memlen * kLanes);
      for (int64_t a = c0; a < c1; a += chunk, li = (li + 1) % kLanes) {
        const int64_t n = std::min(chunk, c1 - a);
        Lane& ln = dc->lanes[li];
        int r = ln.in.reserve(std::max<size_t>(16, n * per_row_in));
        if (r) return r;
        if ((r = ln.out.reserve(n * per_row_out))) return r;
        if ((r = ln.mem.reserve(std::max<size_t>(16, memlen * sizeof(float))))) return r;
        // the chunk's first row continues the stream of the row before it (Q13): restate the
        // overlap memory on the host from the input itself -- no device collective needed
        SS_CUDA(cudaStreamSynchronize(ln.stream));  // hmem slot reuse
        float* hm = hmem.data() + (size_t)li * memlen;
        host_mem_before_row(pl, pcm, ttd, row_stride, a, mem0.data(), hm);
        SS_CUDA(cudaMemcpyAsync(ln.mem.p, hm, memlen * sizeof(float), cudaMemcpyHostToDevice, ln.stream));
        if ((r = copy_rows_h2d(static_cast<float*>(ln.in.p), pcm + a * row_stride, n, ttd, row_stride, ln.stream)))
          return r;
        FrameParams p;
        stft_params(pl, dc, &p, static_cast<float*>(ln.in.p), n, ttd, ttd);
        p.mem_in = static_cast<float*>(ln.mem.p);
        p.log_mode = log_mode;
        if (out_ri) p.stft_out = static_cast<float2*>(ln.out.p);
        else p.mel_out = static_cast<float*>(ln.out.p);
        int e = launch_stft_any(pl, dc, p, ln.stream);
        if (e) return cuda_fail((cudaError_t)e, "stft kernel launch");
        float* host_out = out_ri ? out_ri + (size_t)a * Tp * d.F * 2 : out_mel + (size_t)a * d.nf * Tp;
        SS_CUDA(cudaMemcpyAsync(host_out, ln.out.p, n * per_row_out, cudaMemcpyDeviceToHost, ln.stream));
      }
      return sync_lanes(dc);
    });
    if (rc) return rc;
  }
  // memory after the last row (functions.rs:153-160), kept in the plan like the reference
  // keeps it in the config, or handed back to the caller
  std::vector<float> mem1(memlen);
  host_mem_before_row(pl, pcm, ttd, row_stride, ch, mem0.data(), mem1.data());
  std::memcpy(mem_inout ? mem_inout : pl->analysis_mem.data(), mem1.data(), memlen * sizeof(float));
  return SS_OK;
}

int ss_stft(ss_plan* pl, const float* pcm, int64_t ch, int64_t ttd, int64_t row_stride, float* out_ri,
            float* mem_inout) {
  if (!pl || !out_ri) return fail(SS_ERR_INVALID_ARG, "null argument");
  return stft_host(pl, pcm, ch, ttd, row_stride, out_ri, nullptr, mem_inout, 0);
}
int ss_mel_spectrogram(ss_plan* pl, const float* pcm, int64_t ch, int64_t ttd, int64_t row_stride, float* out,
                       float* mem_inout, int log_mode) {
  if (!pl || !out) return fail(SS_ERR_INVALID_ARG, "null argument");
  if (log_mode < 0 || log_mode > 1) return fail(SS_ERR_INVALID_ARG, "log_mode must be 0 or 1");
  return stft_host(pl, pcm, ch, ttd, row_stride, nullptr, out, mem_inout, log_mode);
}

static int stft_dev(ss_plan* pl, int dev_index, void* stream, const float* d_pcm, int64_t ch, int64_t ttd,
                    int64_t row_stride, float* d_out_ri, float* d_out_mel, const float* d_mem_in, float* d_mem_out,
                    int log_mode) {
  if (!pl) return fail(SS_ERR_INVALID_ARG, "null plan");
  if (dev_index < 0 || dev_index >= (int)pl->devs.size())
    return fail(pl->devs.empty() ? SS_ERR_NO_DEVICE : SS_ERR_INVALID_ARG, "dev_index %d out of range", dev_index);
  int rc = check_stft(pl, ch, ttd, row_stride);
  if (rc) return rc;
  if (d_mem_in && d_mem_in == d_mem_out) return fail(SS_ERR_INVALID_ARG, "d_mem_in and d_mem_out must not alias");
  FrameParams p;
  stft_params(pl, &pl->devs[dev_index], &p, d_pcm, ch, ttd, row_stride);
  p.mem_in = d_mem_in;
  p.mem_out = d_mem_out;
  p.log_mode = log_mode;
  p.stft_out = reinterpret_cast<float2*>(d_out_ri);
  p.mel_out = d_out_mel;
  if (ch > 0 && p.Tp > 0) {
    int e = launch_stft_any(pl, &pl->devs[dev_index], p, stream);
    if (e) return cuda_fail((cudaError_t)e, "stft kernel launch");
  }
  int e = launch_stft_mem_update(p, stream);
  if (e) return cuda_fail((cudaError_t)e, "stft mem kernel launch");
  return SS_OK;
}
int ss_stft_dev(ss_plan* pl, int dev_index, void* stream, const float* d_pcm, int64_t ch, int64_t ttd,
                int64_t row_stride, float* d_out_ri, const float* d_mem_in, float* d_mem_out) {
  if (!d_out_ri) return fail(SS_ERR_INVALID_ARG, "null output");
  return stft_dev(pl, dev_index, stream, d_pcm, ch, ttd, row_stride, d_out_ri, nullptr, d_mem_in, d_mem_out, 0);
}
int ss_mel_spectrogram_dev(ss_plan* pl, int dev_index, void* stream, const float* d_pcm, int64_t ch, int64_t ttd,
                           int64_t row_stride, float* d_out, const float* d_mem_in, float* d_mem_out, int log_mode) {
  if (!d_out) return fail(SS_ERR_INVALID_ARG, "null output");
  return stft_dev(pl, dev_index, stream, d_pcm, ch, ttd, row_stride, nullptr, d_out, d_mem_in, d_mem_out, log_mode);
}

// ------------------------------------------------------------- librosa-compatible mel spectrogram
// (SS_FLAG_LIBROSA plans).  No overlap memory, no dropped or zero frames: every one
// of the 1 + ttd / hop centred frames is computed; rows are independent.
static int check_librosa(ss_plan* pl, int64_t ch, int64_t ttd, int64_t row_stride, int db_mode) {
  if (!pl) return fail(SS_ERR_INVALID_ARG, "null plan");
  const Derived& d = pl->d;
  if (!d.librosa) return fail(SS_ERR_INVALID_ARG, "ss_librosa_melspectrogram needs a plan created with SS_FLAG_LIBROSA");
  if (!d.stft_ok) return fail(SS_ERR_INVALID_ARG, "need 0 < hop=%lld <= fft_points=%lld", (long long)d.H, (long long)d.W);
  if (!pl->fft_ok) return fail(SS_ERR_UNSUPPORTED, "fft_points=%lld has a prime factor > 61", (long long)d.W);
  if (d.W > 8192) return fail(SS_ERR_UNSUPPORTED, "fft_points=%lld > 8192", (long long)d.W);
  if (ch < 0 || row_stride < ttd) return fail(SS_ERR_INVALID_ARG, "bad shape");
  if (ttd <= d.W / 2) return fail(SS_ERR_INVALID_ARG, "reflect padding needs more than n_fft/2 = %lld samples per row", (long long)(d.W / 2));
  if (db_mode < 0 || db_mode > 1) return fail(SS_ERR_INVALID_ARG, "db_mode must be 0 or 1");
  return SS_OK;
}

static void librosa_params(ss_plan* pl, DeviceCtx* dc, FrameParams* p, const float* d_pcm, int64_t ch, int64_t ttd,
                           int64_t row_stride, float* d_out, int db_mode) {
  const Derived& d = pl->d;
  p->pcm = d_pcm;
  p->n_rows = ch;
  p->row_len = ttd;
  p->row_stride = row_stride;
  p->Tp = 1 + ttd / d.H;
  p->n_pad = 0;
  p->T = p->Tp;
  p->T_out = p->Tp;
  p->N = (int)d.W;
  p->F = (int)d.F;
  p->L = (int)d.W;
  p->S = (int)d.H;
  p->nf = d.nf;
  p->nc = d.nc;
  p->mode = kModeLibrosa;
  p->wnorm = 1.0f;
  p->log_mode = db_mode ? 2 : 0;
  p->mel_out = d_out;
  p->tab = dc->tab;
}

// the register-FFT kernel when n_fft = 400 (Whisper's front end), else the generic one
static int launch_librosa_any(ss_plan* pl, DeviceCtx* dc, const FrameParams& p, void* stream) {
  if (pl->stft400m && dc->stft400m && !(pl->d.cfg.flags & SS_FLAG_FORCE_GENERIC)) {
    const int e = launch_stft400m(p, dc->stft400m, dc->sm_count, stream);
    if (e != -1000) return e;
  }
  return launch_frames_generic(p, stream);
}

int64_t ss_librosa_num_frames(int64_t ttd, int64_t hop) { return (hop > 0 && ttd >= 0) ? 1 + ttd / hop : -1; }

int ss_librosa_melspectrogram_dev(ss_plan* pl, int dev_index, void* stream, const float* d_pcm, int64_t ch, int64_t ttd,
                                  int64_t row_stride, float* d_out, int db_mode) {
  if (!d_out) return fail(SS_ERR_INVALID_ARG, "null output");
  int rc = check_librosa(pl, ch, ttd, row_stride, db_mode);
  if (rc) return rc;
  if (dev_index < 0 || dev_index >= (int)pl->devs.size())
    return fail(pl->devs.empty() ? SS_ERR_NO_DEVICE : SS_ERR_INVALID_ARG, "dev_index %d out of range", dev_index);
  if (ch == 0) return SS_OK;
  FrameParams p;
  librosa_params(pl, &pl->devs[dev_index], &p, d_pcm, ch, ttd, row_stride, d_out, db_mode);
  int e = launch_librosa_any(pl, &pl->devs[dev_index], p, stream);
  if (e) return cuda_fail((cudaError_t)e, "librosa mel kernel launch");
  return SS_OK;
}

int ss_librosa_melspectrogram(ss_plan* pl, const float* pcm, int64_t ch, int64_t ttd, int64_t row_stride, float* out,
                              int db_mode) {
  if (!out) return fail(SS_ERR_INVALID_ARG, "null output");
  int rc = check_librosa(pl, ch, ttd, row_stride, db_mode);
  if (rc) return rc;
  if (!pcm && ch > 0) return fail(SS_ERR_INVALID_ARG, "null pcm");
  if (ch == 0) return SS_OK;
  const Derived& d = pl->d;
  const int64_t Tp = 1 + ttd / d.H;
  const size_t per_row_in = (size_t)ttd * sizeof(float), per_row_out = (size_t)Tp * d.nf * sizeof(float);
  return run_on_devices(pl, ch, [&](DeviceCtx* dc, int64_t c0, int64_t c1) -> int {
    SS_CUDA(cudaSetDevice(dc->device));
    int64_t chunk = std::max<int64_t>(1, (int64_t)((64u << 20) / std::max<size_t>(1, per_row_in + per_row_out)));
    chunk = std::min(chunk, c1 - c0);
    int li = 0;
    for (int64_t a = c0; a < c1; a += chunk, li = (li + 1) % kLanes) {
      const int64_t n = std::min(chunk, c1 - a);
      Lane& ln = dc->lanes[li];
      int r = ln.in.reserve(std::max<size_t>(16, n * per_row_in));
      if (r) return r;
      if ((r = ln.out.reserve(n * per_row_out))) return r;
      if ((r = copy_rows_h2d(static_cast<float*>(ln.in.p), pcm + a * row_stride, n, ttd, row_stride, ln.stream))) return r;
      FrameParams p;
      librosa_params(pl, dc, &p, static_cast<float*>(ln.in.p), n, ttd, ttd, static_cast<float*>(ln.out.p), db_mode);
      int e = launch_librosa_any(pl, dc, p, ln.stream);
      if (e) return cuda_fail((cudaError_t)e, "librosa mel kernel launch");
      SS_CUDA(cudaMemcpyAsync(out + (size_t)a * d.nf * Tp, ln.out.p, n * per_row_out, cudaMemcpyDeviceToHost, ln.stream));
    }
    return sync_lanes(dc);
  });
}

// ------------------------------------------------------------- standalone processing functions
int ss_preemphasis_dev(void* stream, const float* d_in, float* d_out, int64_t n_rows, int64_t n,
                       int64_t row_stride, int64_t shift, float cof) {
  if (!d_in || !d_out) return fail(SS_ERR_INVALID_ARG, "null argument");
  if (d_in == d_out) return fail(SS_ERR_INVALID_ARG, "preemphasis cannot run in place (circular read)");
  if (n_rows < 0 || n <= 0 || row_stride < n) return fail(SS_ERR_INVALID_ARG, "bad shape");
  // s![..shift] vs s![-shift..] have different lengths for shift == 0 or |shift| > n: the reference panics
  if (shift == 0 || shift > n || -shift > n)
    return fail(SS_ERR_INVALID_ARG, "preemphasis: need 0 < |shift| <= n (the reference panics, processing.rs:43-50)");
  int e = launch_preemphasis(d_in, d_out, n_rows, n, row_stride, shift, cof, stream);
  if (e) return cuda_fail((cudaError_t)e, "preemphasis launch");
  return SS_OK;
}

int ss_preemphasis_batch(const float* in, float* out, int64_t n_rows, int64_t n, int64_t row_stride, int64_t shift,
                         float cof) {
  if (!in || !out) return fail(SS_ERR_INVALID_ARG, "null argument");
  if (n_rows < 0 || n <= 0 || row_stride < n) return fail(SS_ERR_INVALID_ARG, "bad shape");
  if (shift == 0 || shift > n || -shift > n)
    return fail(SS_ERR_INVALID_ARG, "preemphasis: need 0 < |shift| <= n (the reference panics, processing.rs:43-50)");
  DefaultCtx* c;
  int rc = default_ctx(&c);
  if (rc) return rc;
  // rows are independent: process in chunks of rows, contiguous on the device
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(n_rows, (256ll << 20) / (n * 4)));
  for (int64_t a = 0; a < n_rows; a += chunk) {
    const int64_t m = std::min(chunk, n_rows - a);
    if ((rc = c->a.reserve((size_t)m * n * 4))) return rc;
    if ((rc = c->b.reserve((size_t)m * n * 4))) return rc;
    if ((rc = copy_rows_h2d(static_cast<float*>(c->a.p), in + a * row_stride, m, n, row_stride, c->stream))) return rc;
    if ((rc = ss_preemphasis_dev(c->stream, static_cast<float*>(c->a.p), static_cast<float*>(c->b.p), m, n, n, shift, cof)))
      return rc;
    if (row_stride == n) {
      SS_CUDA(cudaMemcpyAsync(out + a * row_stride, c->b.p, (size_t)m * n * 4, cudaMemcpyDeviceToHost, c->stream));
    } else {
      SS_CUDA(cudaMemcpy2DAsync(out + a * row_stride, row_stride * 4, c->b.p, n * 4, n * 4, m, cudaMemcpyDeviceToHost,
                                c->stream));
    }
    SS_CUDA(cudaStreamSynchronize(c->stream));
  }
  return SS_OK;
}
int ss_preemphasis(const float* in, float* out, int64_t n, int64_t shift, float cof) {
  return ss_preemphasis_batch(in, out, 1, n, n, shift, cof);
}

int ss_stack_frames_dev(void* stream, const float* d_in, int64_t n, int64_t L, int64_t S, int zero_padding,
                        uint32_t flags, float* d_out, int64_t numframes) {
  if (!d_in || !d_out) return fail(SS_ERR_INVALID_ARG, "null argument");
  const int64_t T = num_frames(n, L, S, zero_padding);
  if (T < 0 || T != numframes) return fail(SS_ERR_INVALID_ARG, "numframes does not match ss_num_frames");
  int e = launch_stack_frames(d_in, n, L, S, T, flags, d_out, stream);
  if (e) return cuda_fail((cudaError_t)e, "stack_frames launch");
  return SS_OK;
}

int ss_stack_frames(const float* in, int64_t n, int64_t L, int64_t S, int zero_padding, uint32_t flags, float* out,
                    int64_t* numframes) {
  const int64_t T = num_frames(n, L, S, zero_padding);
  if (T < 0)
    return fail(SS_ERR_INVALID_ARG,
                "stack_frames: n=%lld, L=%lld, S=%lld yields no frame (usize underflow in the reference, "
                "processing.rs:91,101,105)",
                (long long)n, (long long)L, (long long)S);
  if (numframes) *numframes = T;
  if (!out) return SS_OK;
  if (!in) return fail(SS_ERR_INVALID_ARG, "null input");
  if (T == 0) return SS_OK;
  DefaultCtx* c;
  int rc = default_ctx(&c);
  if (rc) return rc;
  if ((rc = c->a.reserve((size_t)n * 4))) return rc;
  if ((rc = c->b.reserve((size_t)T * L * 4))) return rc;
  SS_CUDA(cudaMemcpyAsync(c->a.p, in, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  if ((rc = ss_stack_frames_dev(c->stream, static_cast<float*>(c->a.p), n, L, S, zero_padding, flags,
                                static_cast<float*>(c->b.p), T)))
    return rc;
  SS_CUDA(cudaMemcpyAsync(out, c->b.p, (size_t)T * L * 4, cudaMemcpyDeviceToHost, c->stream));
  SS_CUDA(cudaStreamSynchronize(c->stream));
  return SS_OK;
}

// power_spectrum has no config either: FFT tables are cached per (thread, n_fft)
struct PsCache {
  int n = 0;
  FftPlanHost plan;
  DevBuf tw;
};
static thread_local PsCache g_ps;
static int ps_tables(int64_t n_fft, DevTables* tab) {
  if (n_fft < 2 || n_fft > 8192) return fail(SS_ERR_UNSUPPORTED, "n_fft=%lld outside [2, 8192]", (long long)n_fft);
  if (g_ps.n != n_fft) {
    if (!build_fft_plan((int)n_fft, &g_ps.plan)) return fail(SS_ERR_UNSUPPORTED, "n_fft=%lld has a prime factor > 61", (long long)n_fft);
    int rc = g_ps.tw.reserve(g_ps.plan.tw.size() * sizeof(float));
    if (rc) return rc;
    SS_CUDA(cudaMemcpy(g_ps.tw.p, g_ps.plan.tw.data(), g_ps.plan.tw.size() * sizeof(float), cudaMemcpyHostToDevice));
    g_ps.n = (int)n_fft;
  }
  *tab = DevTables();
  tab->tw = static_cast<const float2*>(g_ps.tw.p);
  tab->n_stages = g_ps.plan.n_stages;
  for (int i = 0; i < kMaxStages; ++i) tab->radix[i] = g_ps.plan.radix[i];
  return SS_OK;
}

int ss_power_spectrum_dev(void* stream, const float* d_frames, int64_t rows, int64_t L, int64_t n_fft, float* d_out) {
  if (!d_frames || !d_out) return fail(SS_ERR_INVALID_ARG, "null argument");
  if (rows < 0 || L <= 0) return fail(SS_ERR_INVALID_ARG, "bad shape");
  if (L > n_fft) return fail(SS_ERR_INVALID_ARG, "frame_len %lld > n_fft %lld (ndrustfft panics, processing.rs:159)", (long long)L, (long long)n_fft);
  if (rows == 0) return SS_OK;
  FrameParams p;
  int rc = ps_tables(n_fft, &p.tab);
  if (rc) return rc;
  p.pcm = d_frames;
  p.n_rows = 1;
  p.row_len = rows * L;
  p.row_stride = rows * L;
  p.T = rows;
  p.T_out = rows;
  p.N = (int)n_fft;
  p.F = (int)(n_fft / 2 + 1);
  p.L = (int)L;
  p.S = (int)L;
  p.nf = 0;
  p.nc = 0;
  p.mode = kModeSpeechpy;
  p.mag_scale = 1.0f / (float)n_fft;
  p.pspec_out = d_out;
  int e = launch_frames_generic(p, stream);
  if (e) return cuda_fail((cudaError_t)e, "power_spectrum launch");
  return SS_OK;
}

int ss_power_spectrum(const float* frames, int64_t rows, int64_t L, int64_t n_fft, float* out) {
  if (!frames || !out) return fail(SS_ERR_INVALID_ARG, "null argument");
  if (rows < 0 || L <= 0) return fail(SS_ERR_INVALID_ARG, "bad shape");
  if (L > n_fft) return fail(SS_ERR_INVALID_ARG, "frame_len %lld > n_fft %lld (ndrustfft panics, processing.rs:159)", (long long)L, (long long)n_fft);
  if (rows == 0) return SS_OK;
  DefaultCtx* c;
  int rc = default_ctx(&c);
  if (rc) return rc;
  const int64_t F = n_fft / 2 + 1;
  if ((rc = c->a.reserve((size_t)rows * L * 4))) return rc;
  if ((rc = c->b.reserve((size_t)rows * F * 4))) return rc;
  SS_CUDA(cudaMemcpyAsync(c->a.p, frames, (size_t)rows * L * 4, cudaMemcpyHostToDevice, c->stream));
  if ((rc = ss_power_spectrum_dev(c->stream, static_cast<float*>(c->a.p), rows, L, n_fft, static_cast<float*>(c->b.p))))
    return rc;
  SS_CUDA(cudaMemcpyAsync(out, c->b.p, (size_t)rows * F * 4, cudaMemcpyDeviceToHost, c->stream));
  SS_CUDA(cudaStreamSynchronize(c->stream));
  return SS_OK;
}

// processing::log_power_spectrum (processing.rs:194-214): private and unused in the reference
static thread_local DevBuf g_logps_scratch;
int ss_log_power_spectrum_dev(void* stream, const float* d_frames, int64_t rows, int64_t L, int64_t n_fft, int normalize,
                              float* d_out) {
  int rc = ss_power_spectrum_dev(stream, d_frames, rows, L, n_fft, d_out);
  if (rc || rows == 0) return rc;
  if ((rc = g_logps_scratch.reserve(16))) return rc;
  int e = launch_log_power(d_out, rows * (n_fft / 2 + 1), normalize, g_logps_scratch.p, stream);
  if (e) return cuda_fail((cudaError_t)e, "log_power_spectrum launch");
  return SS_OK;
}

int ss_log_power_spectrum(const float* frames, int64_t rows, int64_t L, int64_t n_fft, int normalize, float* out) {
  if (!frames || !out) return fail(SS_ERR_INVALID_ARG, "null argument");
  if (rows < 0 || L <= 0) return fail(SS_ERR_INVALID_ARG, "bad shape");
  if (L > n_fft) return fail(SS_ERR_INVALID_ARG, "frame_len %lld > n_fft %lld (ndrustfft panics, processing.rs:159)", (long long)L, (long long)n_fft);
  if (rows == 0) return SS_OK;
  DefaultCtx* c;
  int rc = default_ctx(&c);
  if (rc) return rc;
  const int64_t F = n_fft / 2 + 1;
  if ((rc = c->a.reserve((size_t)rows * L * 4))) return rc;
  if ((rc = c->b.reserve((size_t)rows * F * 4))) return rc;
  SS_CUDA(cudaMemcpyAsync(c->a.p, frames, (size_t)rows * L * 4, cudaMemcpyHostToDevice, c->stream));
  if ((rc = ss_log_power_spectrum_dev(c->stream, static_cast<float*>(c->a.p), rows, L, n_fft, normalize, static_cast<float*>(c->b.p))))
    return rc;
  SS_CUDA(cudaMemcpyAsync(out, c->b.p, (size_t)rows * F * 4, cudaMemcpyDeviceToHost, c->stream));
  SS_CUDA(cudaStreamSynchronize(c->stream));
  return SS_OK;
}

static thread_local DevBuf g_cmvn_scratch;
int ss_cmvn_dev(void* stream, const float* d_in, int64_t n_mats, int64_t rows, int64_t cols, int varnorm,
                float* d_out) {
  if (!d_in || !d_out) return fail(SS_ERR_INVALID_ARG, "null argument");
  if (n_mats < 0 || cols < 0) return fail(SS_ERR_INVALID_ARG, "bad shape");
  if (rows <= 0) return fail(SS_ERR_INVALID_ARG, "cmvn: zero rows (mean_axis(..).unwrap() panics, processing.rs:270)");
  if (cols > 256) return fail(SS_ERR_UNSUPPORTED, "cmvn: more than 256 columns");
  if (n_mats == 0 || cols == 0) return SS_OK;
  int rc = g_cmvn_scratch.reserve(cmvn_scratch_bytes(n_mats, rows, cols));
  if (rc) return rc;
  int e = launch_cmvn(d_in, n_mats, rows, cols, varnorm, d_out, static_cast<float*>(g_cmvn_scratch.p), stream);
  if (e) return cuda_fail((cudaError_t)e, "cmvn launch");
  return SS_OK;
}

int ss_cmvn_batch(const float* in, int64_t n_mats, int64_t rows, int64_t cols, int varnorm, float* out) {
  if (!in || !out) return fail(SS_ERR_INVALID_ARG, "null argument");
  if (n_mats < 0 || cols < 0) return fail(SS_ERR_INVALID_ARG, "bad shape");
  if (rows <= 0) return fail(SS_ERR_INVALID_ARG, "cmvn: zero rows (mean_axis(..).unwrap() panics, processing.rs:270)");
  if (n_mats == 0 || cols == 0) return SS_OK;
  DefaultCtx* c;
  int rc = default_ctx(&c);
  if (rc) return rc;
  const size_t bytes = (size_t)n_mats * rows * cols * 4;
  if ((rc = c->a.reserve(bytes))) return rc;
  if ((rc = c->b.reserve(bytes))) return rc;
  SS_CUDA(cudaMemcpyAsync(c->a.p, in, bytes, cudaMemcpyHostToDevice, c->stream));
  if ((rc = ss_cmvn_dev(c->stream, static_cast<float*>(c->a.p), n_mats, rows, cols, varnorm, static_cast<float*>(c->b.p))))
    return rc;
  SS_CUDA(cudaMemcpyAsync(out, c->b.p, bytes, cudaMemcpyDeviceToHost, c->stream));
  SS_CUDA(cudaStreamSynchronize(c->stream));
  return SS_OK;
}
int ss_cmvn(const float* in, int64_t rows, int64_t cols, int varnorm, float* out) {
  return ss_cmvn_batch(in, 1, rows, cols, varnorm, out);
}

// ------------------------------------------------------------ "next" rows: deltas, cmvnw
int ss_derivative_extraction_dev(void* stream, const float* d_in, int64_t rows, int64_t cols, int dw, float* d_out) {
  if (!d_in || !d_out) return fail(SS_ERR_INVALID_ARG, "null argument");
  if (rows < 0 || cols < 0) return fail(SS_ERR_INVALID_ARG, "bad shape");
  if (dw < 1) return fail(SS_ERR_INVALID_ARG, "delta_windows must be >= 1 (zero scale in the reference, processing.rs:253)");
  if (d_in == d_out) return fail(SS_ERR_INVALID_ARG, "derivative_extraction cannot run in place");
  int e = launch_derivative(d_in, rows, cols, dw, d_out, stream);
  if (e) return cuda_fail((cudaError_t)e, "derivative launch");
  return SS_OK;
}
int ss_extract_derivative_feature_dev(void* stream, const float* d_in, int64_t rows, int64_t cols, float* d_out) {
  if (!d_in || !d_out) return fail(SS_ERR_INVALID_ARG, "null argument");
  if (rows < 0 || cols < 0) return fail(SS_ERR_INVALID_ARG, "bad shape");
  int e = launch_derivative_cube(d_in, rows, cols, d_out, stream);
  if (e) return cuda_fail((cudaError_t)e, "derivative cube launch");
  return SS_OK;
}
static thread_local DevBuf g_cmvnw_scratch;
int ss_cmvnw_dev(void* stream, const float* d_in, int64_t n_mats, int64_t rows, int64_t cols, int64_t win, int varnorm,
                 float* d_out) {
  if (!d_in || !d_out) return fail(SS_ERR_INVALID_ARG, "null argument");
  if (n_mats < 0 || rows < 0 || cols < 0) return fail(SS_ERR_INVALID_ARG, "bad shape");
  if (win < 1 || win % 2 != 1) return fail(SS_ERR_INVALID_ARG, "Windows size must be odd! (processing.rs:327)");
  if (cols > 256 || n_mats > 65535 || win > (1 << 24)) return fail(SS_ERR_UNSUPPORTED, "cmvnw: shape outside this build's limits");
  if (n_mats == 0 || rows == 0 || cols == 0) return SS_OK;
  float* scratch = nullptr;
  if (varnorm) {
    int rc = g_cmvnw_scratch.reserve((size_t)n_mats * rows * cols * sizeof(float));
    if (rc) return rc;
    scratch = static_cast<float*>(g_cmvnw_scratch.p);
  }
  int e = launch_cmvnw(d_in, n_mats, rows, cols, (int)win, varnorm, d_out, scratch, stream);
  if (e) return cuda_fail((cudaError_t)e, "cmvnw launch");
  return SS_OK;
}

int ss_derivative_extraction(const float* in, int64_t rows, int64_t cols, int dw, float* out) {
  if (rows < 0 || cols < 0) return fail(SS_ERR_INVALID_ARG, "bad shape");
  if (dw < 1) return fail(SS_ERR_INVALID_ARG, "delta_windows must be >= 1 (zero scale in the reference, processing.rs:253)");
  return host_unary(in, (size_t)rows * cols, out, (size_t)rows * cols,
                    [&](cudaStream_t s, float* a, float* b) { return ss_derivative_extraction_dev(s, a, rows, cols, dw, b); });
}
int ss_extract_derivative_feature(const float* in, int64_t rows, int64_t cols, float* out) {
  if (rows < 0 || cols < 0) return fail(SS_ERR_INVALID_ARG, "bad shape");
  return host_unary(in, (size_t)rows * cols, out, (size_t)rows * cols * 3,
                    [&](cudaStream_t s, float* a, float* b) { return ss_extract_derivative_feature_dev(s, a, rows, cols, b); });
}
int ss_cmvnw(const float* in, int64_t rows, int64_t cols, int64_t win, int varnorm, float* out) {
  if (rows < 0 || cols < 0) return fail(SS_ERR_INVALID_ARG, "bad shape");
  if (win < 1 || win % 2 != 1) return fail(SS_ERR_INVALID_ARG, "Windows size must be odd! (processing.rs:327)");
  return host_unary(in, (size_t)rows * cols, out, (size_t)rows * cols,
                    [&](cudaStream_t s, float* a, float* b) { return ss_cmvnw_dev(s, a, 1, rows, cols, win, varnorm, b); });
}

}  // extern "C"
#pragma GCC visibility pop
