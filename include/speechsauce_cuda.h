/* speechsauce_cuda.h — C ABI of the B200 (sm_100a) MFCC / mel-spectrogram hot path.
 *
 * This is the drop-in boundary for speechsauce (secretsauceai/mfcc-rust).  Every entry
 * point names the reference interface it replaces (file:line relative to the reference
 * repository root).  The `speechsauce-cuda` Rust crate (mfcc-rust_b200/speechsauce-cuda),
 * the C++ mirror (mfcc-rust_b200/cpp/speechsauce.hpp) and the Python ctypes mirror
 * (mfcc-rust_b200/speechsauce_b200) bind exactly these symbols.
 *
 * Conventions
 *   - plain pointers and sizes only; all arrays are contiguous row-major f32 unless a
 *     row stride (in elements) is given;
 *   - the caller owns every host buffer and sizes outputs with ss_num_frames /
 *     ss_stft_num_frames; the library owns the plan and all device memory;
 *   - every function returns SS_OK (0) or a negative ss_status; ss_last_error() gives a
 *     thread-local message.  The library never aborts and never throws across the ABI.
 *     (The reference panics on the same conditions; the Rust wrapper turns a non-zero
 *     status back into panic!.)
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with
 *     SS_ERR_NO_DEVICE;
 *   - a plan is not thread-safe (mirrors `SpeechConfig: Send + !Sync`, config.rs:126,130).
 */
#ifndef SPEECHSAUCE_CUDA_H
#define SPEECHSAUCE_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  SS_OK = 0,
  SS_ERR_INVALID_ARG = -1, /* a condition on which the reference panics (N < L, H > W/2, hi > sr/2 ...) */
  SS_ERR_CUDA = -2,        /* a CUDA runtime error; message holds cudaGetErrorString */
  SS_ERR_UNSUPPORTED = -3, /* valid for the reference but outside this build's limits */
  SS_ERR_NO_DEVICE = -4,   /* no CUDA device visible */
  SS_ERR_ALLOC = -5
} ss_status;

/* ss_config.flags */
#define SS_FLAG_FRAMING_LITERAL 0x1u /* reproduce processing.rs:110-120 as written (all-zero frames for T > 2) */
#define SS_FLAG_FORCE_GENERIC   0x2u /* never pick a size-specialised kernel (used by the parity tests) */
#define SS_FLAG_LIBROSA         0x4u /* the plan serves ss_librosa_melspectrogram only: periodic Hann window of
                                        fft_points, hop = trunc(frame_length*sample_rate), centred frames with
                                        reflect padding, Slaney mel scale with Slaney (area) normalisation.  The
                                        reference's README / Contributing.md name librosa compatibility as the
                                        project's direction; its code has no such path. */

/* cmvn_mode */
#define SS_CMVN_NONE 0
#define SS_CMVN_MEAN 1     /* processing::cmvn(x, false) */
#define SS_CMVN_MEAN_VAR 2 /* processing::cmvn(x, true)  */

/* The nine arguments of SpeechConfig::new (config.rs:140-150), in that order, plus flags.
 * Derived sizes are computed by the library exactly as the reference does (f32
 * arithmetic): L = round(sr*frame_length), S = round(sr*frame_stride) (processing.rs:77-78),
 * H = trunc(frame_length*sr) (config.rs:154), F = fft_points/2+1 (config.rs:179). */
typedef struct {
  int64_t sample_rate;
  int64_t fft_points;
  float frame_length;
  float frame_stride;
  int64_t num_cepstral;
  int64_t num_filters;
  float low_frequency;
  float high_frequency; /* <= 0 means sample_rate/2 (SpeechConfigBuilder::new, config.rs:44) */
  int32_t dc_elimination;
  uint32_t flags;
  float dct2_scale; /* 0 -> 2.0, ndrustfft's default nddct2 normalisation (feature.rs:123) */
} ss_config;

typedef struct {
  int64_t frame_len;   /* L */
  int64_t frame_step;  /* S */
  int64_t hop;         /* H = frame_size */
  int64_t n_fft;       /* W = window_size */
  int64_t freq_size;   /* F */
  int64_t n_pad;       /* W/H - 1, or -1 when the STFT is undefined for this config (H == 0 or H > W/2) */
  int64_t num_filters;
  int64_t num_cepstral;
  float wnorm;         /* config.rs:178 */
  int32_t n_devices;
  int32_t fast_path;   /* 1 when a size-specialised kernel serves the speechpy-framing path */
  int32_t fast_stft;   /* 1 when a size-specialised kernel serves the STFT path */
} ss_plan_info;

typedef struct ss_plan ss_plan;

/* Output selection for ss_features*: any pointer may be NULL (not computed / not stored). */
typedef struct {
  float* mfe;      /* [n][T][nf]  feature::mfe .0      feature.rs:200-233 */
  float* energy;   /* [n][T]      feature::mfe .1      feature.rs:216-219 */
  float* lmfe;     /* [n][T][nf]  feature::lmfe        feature.rs:242-245 */
  float* ssc;      /* [n][T][nf]  feature::ssc         README.md:15 (defined in DESIGN.md) */
  float* mfcc;     /* [n][T][nc]  feature::mfcc        feature.rs:99-148 */
  int32_t mfcc_cmvn_mode; /* SS_CMVN_*: processing::cmvn applied per clip to the mfcc block */
  int32_t reserved;
} ss_feature_outputs;

/* ---- library / device ---------------------------------------------------------------- */
int ss_device_count(void);
const char* ss_last_error(void);
const char* ss_version(void);
/* Host memory.  Every host entry point accepts pageable memory (what `ArrayView::as_ptr()` of an
 * ordinary ndarray / numpy array is, py-speechsauce/src/lib.rs:167-204): chunks are staged through
 * pinned rings by a small pool of copy threads (ss_copy_threads(), env SS_COPY_THREADS) while the
 * neighbouring chunks are on the wire.  Pinned memory (ss_host_alloc, or a caller's own buffer
 * registered once with ss_host_register) is DMA'd in place and skips that copy. */
int ss_host_alloc(void** ptr, size_t bytes);
int ss_host_free(void* ptr);
int ss_host_register(void* ptr, size_t bytes); /* cudaHostRegister(portable): pin a caller-owned range */
int ss_host_unregister(void* ptr);
int ss_host_is_pinned(const void* ptr);        /* 1 when the library will DMA `ptr` in place */
int ss_copy_threads(void);                     /* size of the staging copy pool */
/* the staging pool's 2-D copy on its own (host memory only, usable without a device): `rows` rows of `row_bytes`,
 * source rows `src_pitch` bytes apart (they may overlap: the frame tiles of a long clip), destination rows
 * `dst_pitch` >= row_bytes apart; blocks until done */
int ss_debug_copy2d(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t row_bytes, size_t rows);
/* frees the per-(device, stream) scratch the config-less *_dev entry points keep between calls */
int ss_release_scratch(void);

/* ---- plan (replaces SpeechConfig's FFT plan / window / overlap memory, config.rs:99-185) */
int ss_plan_create(const ss_config* cfg, const int* devices, int n_devices, ss_plan** out);
void ss_plan_destroy(ss_plan* plan);
int ss_plan_get_info(const ss_plan* plan, ss_plan_info* info);
/* Host-pipeline tuning (results never depend on it).  tile_frames: clips with more frames than this are
 * processed in frame tiles with an L - S sample halo so that one long stream still overlaps its H2D,
 * kernels and D2H (0 = automatic: only clips larger than 1.5 chunks).  chunk_bytes: input + output
 * bytes per pipeline chunk (0 = keep; default 64 MiB). */
int ss_plan_set_tuning(ss_plan* plan, int64_t tile_frames, int64_t chunk_bytes);
/* STFT overlap memory held by the plan (config.rs:126,162): [W-H] floats.  The reference
 * never resets it (functions.rs:89-93); these two calls make the hidden state explicit. */
int ss_plan_get_analysis_mem(const ss_plan* plan, float* out, int64_t n);
int ss_plan_set_analysis_mem(ss_plan* plan, const float* in, int64_t n); /* in == NULL -> zeros */

/* ---- frame counts (host arithmetic, usable without a device) ---------------------------- */
/* processing.rs:91-92,101-102: f32 division, ceil with zero padding, floor without; no "+1".
 * Returns -1 where the reference underflows (n < L, or no padding and fewer than one frame). */
int64_t ss_num_frames(int64_t n, int64_t frame_len, int64_t frame_step, int zero_padding);
/* functions.rs:97 (without n_pad): ceil(ttd as f32 / hop as f32) */
int64_t ss_stft_num_frames(int64_t ttd, int64_t hop);

/* ---- speechpy-framing features (host buffers; clips are sharded over the plan's devices) - */
/* pcm: [n_clips][clip_len] with row stride `row_stride` elements. T = ss_num_frames(clip_len, L, S, 0). */
int ss_features(ss_plan* plan, const float* pcm, int64_t n_clips, int64_t clip_len,
                int64_t row_stride, const ss_feature_outputs* outs);
int ss_mfe(ss_plan* plan, const float* pcm, int64_t n_clips, int64_t clip_len, int64_t row_stride,
           float* feat, float* energy);                       /* feature.rs:200 */
int ss_lmfe(ss_plan* plan, const float* pcm, int64_t n_clips, int64_t clip_len, int64_t row_stride,
            float* out);                                      /* feature.rs:242 */
int ss_mfcc(ss_plan* plan, const float* pcm, int64_t n_clips, int64_t clip_len, int64_t row_stride,
            float* out, int cmvn_mode);                       /* feature.rs:99 */
int ss_ssc(ss_plan* plan, const float* pcm, int64_t n_clips, int64_t clip_len, int64_t row_stride,
           float* out);                                       /* README.md:15 */

/* ---- STFT family ---------------------------------------------------------------------- */
/* functions::stft2 (functions.rs:86-123); stft1 (functions.rs:199-233) is ch == 1.
 * out_ri: [ch][T'][F][2] interleaved re/im, T' = ss_stft_num_frames(ttd, H); the last n_pad
 * time rows are zero.  analysis_mem_inout: [W-H] overlap memory read before row 0 and
 * overwritten with the memory after the last row; NULL -> the plan's own memory is used and
 * updated (the reference's behaviour: state lives in the config and is never reset). */
int ss_stft(ss_plan* plan, const float* pcm, int64_t ch, int64_t ttd, int64_t row_stride,
            float* out_ri, float* analysis_mem_inout);
/* feature::mel_spectrogram2 (feature.rs:163-174); mel_spectrogram1 (feature.rs:151-162) is
 * ch == 1.  out: [ch][nf][T'] (time-minor).  log_mode 0: as the reference; 1: ln(zero_handling(.)). */
int ss_mel_spectrogram(ss_plan* plan, const float* pcm, int64_t ch, int64_t ttd, int64_t row_stride,
                       float* out, float* analysis_mem_inout, int log_mode);

/* ---- standalone processing functions (no config in the reference; run on device 0 of the
 *      calling thread's default context) ------------------------------------------------- */
/* processing::preemphasis (processing.rs:31-53): out[i] = in[i] - cof*in[(i-shift) mod n], per row */
int ss_preemphasis(const float* in, float* out, int64_t n, int64_t shift, float cof);
int ss_preemphasis_batch(const float* in, float* out, int64_t n_rows, int64_t n, int64_t row_stride,
                         int64_t shift, float cof);
/* processing::stack_frames (processing.rs:65-129), filter = None.  out: [numframes][L];
 * *numframes is written even when out == NULL (size query). */
int ss_stack_frames(const float* in, int64_t n, int64_t frame_len, int64_t frame_step,
                    int zero_padding, uint32_t flags, float* out, int64_t* numframes);
/* processing::power_spectrum (processing.rs:179-181): (1/n_fft)*|rfft(zero-padded row)|, out [rows][n_fft/2+1] */
int ss_power_spectrum(const float* frames, int64_t rows, int64_t frame_len, int64_t n_fft, float* out);
/* librosa.feature.melspectrogram(y, sr, n_fft, hop_length, window="hann", center=True, pad_mode="reflect",
 * power=2.0, n_mels, fmin, fmax, htk=False, norm="slaney") for a plan created with SS_FLAG_LIBROSA.
 * out [ch][n_mels][1 + ttd / hop] (time-minor, like librosa's [..., n_mels, t]).  db_mode 0: power;
 * 1: librosa.power_to_db(S, ref=1.0, amin=1e-10, top_db=None) = 10 log10(max(1e-10, S)).  Needs ttd > n_fft/2. */
int ss_librosa_melspectrogram(ss_plan* plan, const float* pcm, int64_t ch, int64_t ttd, int64_t row_stride,
                              float* out, int db_mode);
int64_t ss_librosa_num_frames(int64_t ttd, int64_t hop); /* 1 + ttd / hop */
/* processing::log_power_spectrum (processing.rs:194-214; private and unused in the reference):
 * 10 log10 of the power spectrum, -200 below 1e-20; with normalize != 0 the LAST value above the
 * threshold in row-major order (not the maximum: the reference's `mx` is overwritten by every
 * element) is subtracted from every element.  out [rows][n_fft/2+1] */
int ss_log_power_spectrum(const float* frames, int64_t rows, int64_t frame_len, int64_t n_fft, int normalize, float* out);
/* processing::cmvn (processing.rs:265-300) on one [rows][cols] matrix / a batch of them */
int ss_cmvn(const float* in, int64_t rows, int64_t cols, int variance_normalization, float* out);
int ss_cmvn_batch(const float* in, int64_t n_mats, int64_t rows, int64_t cols,
                  int variance_normalization, float* out);

/* ---- "next" rows: delta features and sliding-window CMVN.  The reference's util::pad
 *      symmetric / edge modes are degenerate as written, so these follow the speechpy behaviour the
 *      Rust code transliterates (numpy 'edge' / 'symmetric' padding); see DESIGN.md. --------------- */
/* processing::derivative_extraction (processing.rs:222-254): in/out [rows][cols], edge-padded along features */
int ss_derivative_extraction(const float* in, int64_t rows, int64_t cols, int delta_windows, float* out);
/* feature::extract_derivative_feature (feature.rs:253-269): out [rows][cols][3] = static, delta, delta-delta */
int ss_extract_derivative_feature(const float* in, int64_t rows, int64_t cols, float* out_cube);
/* processing::cmvnw (processing.rs:315-371): sliding window (odd win_size), symmetric padding along time */
int ss_cmvnw(const float* in, int64_t rows, int64_t cols, int64_t win_size, int variance_normalization, float* out);
int ss_derivative_extraction_dev(void* cuda_stream, const float* d_in, int64_t rows, int64_t cols, int delta_windows,
                                 float* d_out);
int ss_extract_derivative_feature_dev(void* cuda_stream, const float* d_in, int64_t rows, int64_t cols, float* d_out_cube);
int ss_cmvnw_dev(void* cuda_stream, const float* d_in, int64_t n_mats, int64_t rows, int64_t cols, int64_t win_size,
                 int variance_normalization, float* d_out);

/* ---- device-resident variants: device pointers on plan device `dev_index`, enqueued on
 *      `cuda_stream` (a cudaStream_t; NULL = legacy default stream), no synchronisation.
 *      Used by the throughput harness and by callers that already hold data in HBM. ------- */
int ss_features_dev(ss_plan* plan, int dev_index, void* cuda_stream, const float* d_pcm,
                    int64_t n_clips, int64_t clip_len, int64_t row_stride,
                    const ss_feature_outputs* d_outs);
int ss_stft_dev(ss_plan* plan, int dev_index, void* cuda_stream, const float* d_pcm, int64_t ch,
                int64_t ttd, int64_t row_stride, float* d_out_ri, const float* d_mem_in,
                float* d_mem_out);
int ss_mel_spectrogram_dev(ss_plan* plan, int dev_index, void* cuda_stream, const float* d_pcm,
                           int64_t ch, int64_t ttd, int64_t row_stride, float* d_out,
                           const float* d_mem_in, float* d_mem_out, int log_mode);
int ss_preemphasis_dev(void* cuda_stream, const float* d_in, float* d_out, int64_t n_rows, int64_t n,
                       int64_t row_stride, int64_t shift, float cof);
int ss_stack_frames_dev(void* cuda_stream, const float* d_in, int64_t n, int64_t frame_len,
                        int64_t frame_step, int zero_padding, uint32_t flags, float* d_out,
                        int64_t numframes);
int ss_power_spectrum_dev(void* cuda_stream, const float* d_frames, int64_t rows, int64_t frame_len,
                          int64_t n_fft, float* d_out);
int ss_librosa_melspectrogram_dev(ss_plan* plan, int dev_index, void* cuda_stream, const float* d_pcm, int64_t ch,
                                  int64_t ttd, int64_t row_stride, float* d_out, int db_mode);
int ss_log_power_spectrum_dev(void* cuda_stream, const float* d_frames, int64_t rows, int64_t frame_len,
                              int64_t n_fft, int normalize, float* d_out);
int ss_cmvn_dev(void* cuda_stream, const float* d_in, int64_t n_mats, int64_t rows, int64_t cols,
                int variance_normalization, float* d_out);

/* ---- introspection used by the tests and the bench (host tables, launch accounting) ----- */
/* dense [nf][F] filterbank exactly as feature::filterbanks builds it (feature.rs:36-90) */
int ss_debug_filterbank(const ss_plan* plan, float* out_nf_by_F);
/* Vorbis window [W] (config.rs:151-160) */
int ss_debug_window(const ss_plan* plan, float* out_W);
/* number of kernels this library has launched from the calling process since load */
int64_t ss_kernel_launch_count(void);
/* name of the fused frame kernel this process launched last ("k_fast512c", "k_stft400m", ...; "" before the
   first launch); a pointer to a string literal */
const char* ss_last_kernel_name(void);

#ifdef __cplusplus
}
#endif
#endif /* SPEECHSAUCE_CUDA_H */
