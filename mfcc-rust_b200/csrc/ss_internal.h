// Internal declarations shared by the host layer (ss_api.cu, ss_tables.cpp) and the kernels.
#pragma once
#include <cstdint>
#include <vector_types.h>
#include <string>
#include <vector>

#include "speechsauce_cuda.h"

namespace ss {

constexpr float kEps = 1.1920929e-07f;          // f32::EPSILON, functions.rs:70
constexpr float kCmvnEps = 9.31322574615478515625e-10f;  // 2^-30, processing.rs:266
constexpr int kMaxStages = 16;

// ---- host-side derived configuration (pure C++, no CUDA; unit-tested without a GPU) ------
struct Derived {
  ss_config cfg;
  int64_t L = 0, S = 0, H = 0, W = 0, F = 0, n_pad = -1;
  int nf = 0, nc = 0;
  float wnorm = 0.f;
  float dct2_scale = 2.f;
  bool stft_ok = false;      // 0 < H <= W/2
  bool speechpy_ok = false;  // 0 < L <= W and S > 0
  bool librosa = false;      // SS_FLAG_LIBROSA: Hann window, centred reflect-padded frames, Slaney mel bank
};

// Mel filterbank in banded (CSR-like) form: filter i covers bins [start[i], start[i]+len[i])
// with weights w[wofs[i] ...].  Built exactly as feature::filterbanks (feature.rs:36-90).
struct MelBank {
  std::vector<int> start, len, wofs;
  std::vector<float> w;
  std::vector<int64_t> idx;  // the nf+2 bin edges
  int max_len = 0;
  int max_bin = -1;  // highest bin with a non-zero weight
};

struct FftPlanHost {
  int n = 0;
  int n_stages = 0;
  int radix[kMaxStages] = {0};
  std::vector<float> tw;  // interleaved cos/sin of e^{-2 pi i q / n}, q = 0..n-1
};

// Validates the nine SpeechConfig::new arguments and derives L/S/H/F/n_pad/wnorm in f32
// exactly as the reference does.  Returns an error string ("" on success).
std::string derive(const ss_config& cfg, Derived* out);
int64_t rust_round_f32(float x);
int64_t num_frames(int64_t n, int64_t L, int64_t S, int zero_padding);
int64_t stft_num_frames(int64_t ttd, int64_t H);
void build_melbank(const Derived& d, MelBank* out);
void build_window(const Derived& d, std::vector<float>* out);
// [nc][nf] DCT-II matrix including dct2_scale (no length-dependent scale)
void build_dct(const Derived& d, std::vector<float>* out);
bool build_fft_plan(int n, FftPlanHost* out);
// balanced assignment of filters to `lanes` lanes (longest-processing-time first)
void balance_filters(const MelBank& mb, int lanes, std::vector<std::vector<int>>* per_lane);

// ---- device-side parameter blocks --------------------------------------------------------
struct DevTables {
  const float2* tw = nullptr;     // [W] twiddles e^{-2 pi i q / W}
  const float* window = nullptr;  // [W] Vorbis window
  const int* fstart = nullptr;    // [nf]
  const int* flen = nullptr;      // [nf]
  const int* fwofs = nullptr;     // [nf]
  const float* fw = nullptr;      // banded weights
  const float* dct = nullptr;     // [nc][nf]
  int n_stages = 0;
  int radix[kMaxStages] = {0};
};

enum Mode : int { kModeSpeechpy = 0, kModeStft = 1, kModePowerSpectrum = 2, kModeLibrosa = 3 };

// One launch of the frame pipeline: `n_rows` rows (clips / channels), `T` frames each.
struct FrameParams {
  const float* pcm = nullptr;
  int64_t n_rows = 0, row_len = 0, row_stride = 0;
  int64_t T = 0;       // frames computed per row
  int64_t T_out = 0;   // frames stored per row (STFT: T' including the n_pad zero rows)
  int N = 0, F = 0;    // FFT size, bins
  int L = 0, S = 0;    // speechpy: frame length / step; STFT: W / H
  int nf = 0, nc = 0;
  int mode = kModeSpeechpy;
  uint32_t flags = 0;
  // STFT stream model (functions.rs:86-170): rows are zero-padded to Tp*H and concatenated
  // behind `mem_in` ([W-H], may be null = zeros); frame j of row r starts at virtual stream
  // position r*Tp*H + j*H.
  int64_t Tp = 0;
  int n_pad = 0;
  const float* mem_in = nullptr;
  float* mem_out = nullptr;
  float wnorm = 0.f;
  float mag_scale = 0.f;  // speechpy: 1/n_fft (processing.rs:180)
  // mfcc epilogue (feature.rs:126-146)
  float scale_rest = 0.f, scale_00 = 0.f;
  int dc_elim = 1;
  int cmvn_mode = 0;
  float ssc_r0 = 1.f, ssc_step = 0.f;
  int log_mode = 0;
  // outputs (device pointers, may be null)
  float* mfe = nullptr;
  float* energy = nullptr;
  float* lmfe = nullptr;
  float* ssc = nullptr;
  float* mfcc = nullptr;
  float2* stft_out = nullptr;  // [row][T_out][F]
  float* mel_out = nullptr;    // [row][nf][T_out]
  float* pspec_out = nullptr;  // [row][F] (power_spectrum of pre-framed rows)
  DevTables tab;
};

// ---- kernel launchers (each returns cudaError_t as int; all enqueue on `stream`) ---------
int launch_frames_generic(const FrameParams& p, void* stream);
int launch_stft_mem_update(const FrameParams& p, void* stream);
// second-generation n_fft == 512 kernel: 16 lanes x 16 points per frame (ss_kernels_fast512b.cu)
bool fast512b_supported(const Derived& d, const MelBank& mb);
struct Fast512bTables;
int fast512b_build_tables(const Derived& d, const MelBank& mb, Fast512bTables** out);
void fast512b_free_tables(Fast512bTables* t);
int launch_fast512b(const FrameParams& p, const Fast512bTables* t, int sm_count, void* stream);

// third-generation n_fft == 512 kernel: FFT as in the second, mel + DCT on mma.sync TF32x3 (ss_kernels_fast512c.cu)
bool fast512c_supported(const Derived& d, const MelBank& mb);
struct Fast512cTables;
int fast512c_build_tables(const Derived& d, const MelBank& mb, Fast512cTables** out);
void fast512c_free_tables(Fast512cTables* t);
int launch_fast512c(const FrameParams& p, const Fast512cTables* t, int sm_count, void* stream);

// fft_points == 400 register-FFT path for mel_spectrogram (ss_kernels_stft400.cu)
bool stft400_supported(const Derived& d, const MelBank& mb);
struct Stft400Tables;
int stft400_build_tables(const Derived& d, const MelBank& mb, const std::vector<float>& window, Stft400Tables** out);
void stft400_free_tables(Stft400Tables* t);
int launch_stft400(const FrameParams& p, const Stft400Tables* t, int sm_count, void* stream);

// second-generation fft_points == 400 kernel: warp-independent tiles, mel on mma.sync TF32x3 (ss_kernels_stft400m.cu)
bool stft400m_supported(const Derived& d, const MelBank& mb);
struct Stft400mTables;
int stft400m_build_tables(const Derived& d, const MelBank& mb, const std::vector<float>& window, Stft400mTables** out);
void stft400m_free_tables(Stft400mTables* t);
int launch_stft400m(const FrameParams& p, const Stft400mTables* t, int sm_count, void* stream);

// fft_points == 2048 register-FFT path for mel_spectrogram (ss_kernels_stft2048.cu)
bool stft2048_supported(const Derived& d, const MelBank& mb);
struct Stft2048Tables;
int stft2048_build_tables(const Derived& d, const MelBank& mb, const std::vector<float>& window, Stft2048Tables** out);
void stft2048_free_tables(Stft2048Tables* t);
int launch_stft2048(const FrameParams& p, const Stft2048Tables* t, int sm_count, void* stream);
// librosa mode for n_fft == 2048: |X|^2 rows (all 1025 bins) + the banded mel kernel
size_t stft2048_power_scratch_bytes(const FrameParams& p, int64_t* rows_per_launch);
int launch_stft2048_power(const FrameParams& p, const Stft2048Tables* t, float* scratch, int sm_count, void* stream);

// second-generation fft_points == 2048 mel kernel: CTA-shared |X|^2 tile, mel on mma.sync TF32x3 (ss_kernels_stft2048m.cu)
bool stft2048m_supported(const Derived& d, const MelBank& mb);
struct Stft2048mTables;
int stft2048m_build_tables(const Derived& d, const MelBank& mb, const std::vector<float>& window, Stft2048mTables** out);
void stft2048m_free_tables(Stft2048mTables* t);
int launch_stft2048m(const FrameParams& p, const Stft2048mTables* t, int sm_count, void* stream);

// processing::power_spectrum for n_fft == 512: the register FFT of the fused path without its epilogue (ss_kernels_pspec512.cu)
bool pspec512_supported(int64_t L, int64_t n_fft);
int launch_pspec512(const float* frames, int64_t rows, int64_t L, float* out, void* stream);

// fft_points == 512 on the STFT / librosa paths: register FFT + a banded mel kernel over a |X|^2 scratch (ss_kernels_stft512.cu)
bool stft512_supported(const Derived& d);
size_t stft512_scratch_bytes(const FrameParams& p, int64_t* rows_per_launch);
int launch_stft512(const FrameParams& p, float* scratch, int sm_count, void* stream);
int launch_mel_from_power(const float* power, const FrameParams& p, int64_t row0, int64_t n_rows, int F, void* stream);
// bytes of |X|^2 scratch per launch of the two-kernel paths (256 MB; SS_POWER_SLAB_KB shrinks it so that tests cross slabs)
int64_t power_slab_bytes();

int launch_preemphasis(const float* in, float* out, int64_t n_rows, int64_t n, int64_t row_stride,
                       int64_t shift, float cof, void* stream);
int launch_stack_frames(const float* in, int64_t n, int64_t L, int64_t S, int64_t T, uint32_t flags,
                        float* out, void* stream);
int launch_cmvn(const float* in, int64_t n_mats, int64_t rows, int64_t cols, int varnorm, float* out,
                float* scratch, void* stream);
size_t cmvn_scratch_bytes(int64_t n_mats, int64_t rows, int64_t cols);
int launch_log_power(float* x, int64_t n, int normalize, void* scratch16, void* stream);
int launch_derivative(const float* in, int64_t rows, int64_t cols, int dw, float* out, void* stream);
int launch_derivative_cube(const float* in, int64_t rows, int64_t cols, float* out, void* stream);
int launch_cmvnw(const float* in, int64_t n_mats, int64_t rows, int64_t cols, int win, int varnorm, float* out,
                 float* scratch, void* stream);

void count_launch(int n = 1);
void note_kernel(const char* name);  // name of the fused frame kernel just launched (ss_last_kernel_name)

}  // namespace ss
