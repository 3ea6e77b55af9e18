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

#include "ss_host.h"
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

static int reserve(DevBuf& b, size_t bytes) {
  cudaError_t e = b.reserve_raw(bytes);
  if (e != cudaSuccess) return fail(SS_ERR_ALLOC, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
  return SS_OK;
}
static int reserve(PinBuf& b, size_t bytes) {
  cudaError_t e = b.reserve_raw(bytes);
  if (e != cudaSuccess) return fail(SS_ERR_ALLOC, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
  return SS_OK;
}

constexpr int kMaxOuts = 5;
struct HostOut {
  float* host = nullptr;
  size_t per_row = 0;  // floats per row; rows are dense in the caller's buffer
};

// what a lane still owes the caller: the rows [a, a + n) of every output sit in its pinned ring
struct Pending {
  bool valid = false;
  int64_t a = 0, n = 0;
  int n_outs = 0;
  HostOut outs[kMaxOuts];
  size_t off[kMaxOuts] = {0};
};

struct Lane {
  cudaStream_t stream = nullptr;
  DevBuf in, out, scratch, mem;
  PinBuf hin, hout, hmem;  // pinned rings for pageable caller memory (and the STFT halo)
  Pending pending;
  std::vector<CopyPool::TicketPtr> out_tickets;
};

struct DeviceCtx {
  int device = 0;
  int sm_count = 148;
  bool ready = false;  // tables + streams are created on first use (a plan over 8 GPUs that only ever
  std::mutex init_mu;  // sees single clips touches device 0 alone)
  DevBuf tables;       // one allocation holding every device table
  DevTables tab;
  Fast512cTables* fast512c = nullptr;
  Stft400mTables* stft400m = nullptr;
  Stft2048Tables* stft2048 = nullptr;
  Stft2048mTables* stft2048m = nullptr;
  Lane lanes[kLanes];
  int next_lane = 0;
  DevBuf clip_mfcc;                       // un-normalised cepstra of a clip that is processed in frame tiles
  cudaEvent_t lane_evt[kLanes] = {nullptr, nullptr, nullptr};
  std::mutex scratch_mu;
  std::map<void*, DevBuf> dev_scratch;    // cmvn partials of ss_features_dev, per caller stream
  std::string err;                        // worker-thread error message
  DeviceCtx() = default;
  DeviceCtx(const DeviceCtx&) = delete;
};

}  // namespace ss

struct ss_plan {
  ss::Derived d;
  ss::MelBank mb;
  ss::FftPlanHost fft;
  bool fft_ok = false;
  std::vector<float> window, dct;
  std::vector<float> analysis_mem;  // [W-H], config.rs:126,162
  std::vector<std::unique_ptr<ss::DeviceCtx>> devs;
  std::vector<std::unique_ptr<ss::DeviceWorker>> workers;  // one persistent thread per device (multi-device plans)
  int64_t tile_frames = 0;        // 0: automatic (ss_plan_set_tuning)
  size_t chunk_bytes = 64u << 20;  // H2D + D2H bytes per pipeline chunk
  bool fast512c = false;
  bool stft400m = false;
  bool stft2048 = false;
  bool stft2048m = false;
};

namespace ss {

static int upload_tables(ss_plan* pl, DeviceCtx* dc) {
  int smc = 0;
  SS_CUDA(cudaDeviceGetAttribute(&smc, cudaDevAttrMultiProcessorCount, dc->device));
  dc->sm_count = smc;
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
  int rc = reserve(dc->tables, total);
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
  for (auto& ev : dc->lane_evt) SS_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  if (pl->fast512c) {
    int e = fast512c_build_tables(pl->d, pl->mb, &dc->fast512c);
    if (e) return cuda_fail((cudaError_t)e, "fast512c_build_tables");
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

// Tables, streams and events of a device are created by the first call that runs on it.  The caller's
// current device is `dc->device` (DeviceGuard or the device's worker thread).
static int ensure_device(ss_plan* pl, DeviceCtx* dc) {
  std::lock_guard<std::mutex> lk(dc->init_mu);
  if (dc->ready) return SS_OK;
  int rc = upload_tables(pl, dc);
  if (rc) return rc;
  dc->ready = true;
  return SS_OK;
}

// ---- speechpy-framing features on one device, device pointers ---------------------------
// `T_clip` / `first_tile`: when a long clip is processed in frame tiles the rows handed in are tiles of
// one clip; the length-dependent DCT scale (feature.rs:126) uses the clip's frame count and only the
// clip's very first frame is element [0,0] (feature.rs:128).  T_clip == 0: rows are whole clips.
// `T_rows`: frames per row when it is not the reference's count for `clip_len` samples (tiles).
static int features_dev(ss_plan* pl, DeviceCtx* dc, cudaStream_t stream, const float* d_pcm, int64_t n_clips,
                        int64_t clip_len, int64_t row_stride, const ss_feature_outputs& o, DevBuf* scratch,
                        int64_t T_clip = 0, bool first_tile = true, int64_t T_rows = 0) {
  const Derived& d = pl->d;
  const int64_t T = T_rows > 0 ? T_rows : num_frames(clip_len, d.L, d.S, 0);
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
  const float n = (float)((T_clip > 0 ? T_clip : T) * (int64_t)d.nf);  // feature.rs:126
  p.scale_00 = first_tile ? 1.0f / std::sqrt(4.0f * n) : 1.0f;           // D[1:,0] stays unscaled (feature.rs:128)
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
  const bool use_fast = pl->fast512c && !(d.cfg.flags & (SS_FLAG_FORCE_GENERIC | SS_FLAG_FRAMING_LITERAL));
  if (use_fast) {
    e = launch_fast512c(p, dc->fast512c, dc->sm_count, stream);
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
    int rc = reserve(*scratch, cmvn_scratch_bytes(n_clips, T, d.nc));
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

static DevBuf* plan_stream_scratch(DeviceCtx* dc, void* stream);

// fft_points = 512 / 1024: the register-FFT kernel + the banded mel kernel; the |X|^2 scratch lives per caller stream
static int launch_stft512_any(ss_plan* pl, DeviceCtx* dc, const FrameParams& p, void* stream) {
  if (!stft512_supported(pl->d) || (pl->d.cfg.flags & SS_FLAG_FORCE_GENERIC) || !(p.mel_out || p.stft_out)) return -1000;
  float* scratch = nullptr;
  if (p.mel_out) {
    DevBuf* sb = plan_stream_scratch(dc, stream);
    if (reserve(*sb, stft512_scratch_bytes(p, nullptr))) return (int)cudaErrorMemoryAllocation;
    scratch = static_cast<float*>(sb->p);
  }
  return launch_stft512(p, scratch, dc->sm_count, stream);
}

// mel_spectrogram / stft launch: a size-specialised kernel when one exists, else the generic one
static int launch_stft_any(ss_plan* pl, DeviceCtx* dc, const FrameParams& p, void* stream) {
  {
    const int e = launch_stft512_any(pl, dc, p, stream);
    if (e != -1000) return e;
  }
  if (pl->stft400m && dc->stft400m && (p.mel_out || p.stft_out) && !(pl->d.cfg.flags & SS_FLAG_FORCE_GENERIC)) {
    const int e = launch_stft400m(p, dc->stft400m, dc->sm_count, stream);
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

static int copy_rows_h2d(float* dst, const float* src, int64_t rows, int64_t len, int64_t row_stride,
                         cudaStream_t s);

// Retire what a lane still owes the caller: wait for its stream, then move the rows parked in the pinned
// ring to the caller's (pageable) arrays on the copy pool.  `async`: leave the copies in flight.
static int lane_retire(Lane& ln, bool async) {
  SS_CUDA(cudaStreamSynchronize(ln.stream));
  if (ln.pending.valid) {
    const Pending& pd = ln.pending;
    for (int i = 0; i < pd.n_outs; ++i) {
      if (!pd.outs[i].host) continue;
      const size_t bytes = (size_t)pd.n * pd.outs[i].per_row * sizeof(float);
      ln.out_tickets.push_back(CopyPool::instance().copy2d_async(pd.outs[i].host + (size_t)pd.a * pd.outs[i].per_row, bytes,
                                                                static_cast<char*>(ln.hout.p) + pd.off[i], bytes, bytes, 1));
    }
    ln.pending.valid = false;
  }
  if (!async) {
    for (auto& t : ln.out_tickets) CopyPool::instance().wait(t);
    ln.out_tickets.clear();
  }
  return SS_OK;
}
static int drain_lanes(DeviceCtx* dc) {
  int rc = SS_OK;
  for (auto& ln : dc->lanes) {
    int r = lane_retire(ln, false);
    if (r && !rc) rc = r;
  }
  return rc;
}
// after an error: nothing may stay in flight towards caller memory
static void abort_lanes(DeviceCtx* dc) {
  for (auto& ln : dc->lanes) {
    if (ln.stream) cudaStreamSynchronize(ln.stream);
    for (auto& t : ln.out_tickets) CopyPool::instance().wait(t);
    ln.out_tickets.clear();
    ln.pending.valid = false;
  }
  cudaGetLastError();
}

// A batch of rows on the host: dense outputs, input rows `row_stride` floats apart (rows may overlap:
// the frame tiles of one long clip are rows TtS apart and (Tt-1)S+L long).
struct RowBatch {
  const float* in = nullptr;
  int64_t row_len = 0, row_stride = 0;
  int n_outs = 0;
  HostOut outs[kMaxOuts];
  size_t chunk_bytes = 64u << 20;
};
// enqueue the kernels for rows [a, a + n) on ln.stream; d_in is [n][row_len] dense, d_outs[i] dense
using LaunchFn = std::function<int(Lane& ln, int64_t a, int64_t n, const float* d_in, float* const* d_outs)>;

// The chunked H2D -> kernels -> D2H pipeline over the device's three lanes.  Pinned caller memory is
// DMA'd in place; pageable memory goes through the lanes' pinned rings, filled and emptied by the copy
// pool while the neighbouring chunks are on the wire or in the kernels.
static int pipeline_rows(DeviceCtx* dc, const RowBatch& b, int64_t c0, int64_t c1, const LaunchFn& fn, bool drain) {
  if (c1 <= c0) return drain ? drain_lanes(dc) : SS_OK;
  static const bool force_stage = getenv("SS_FORCE_STAGING") && atoi(getenv("SS_FORCE_STAGING"));
  const bool in_pinned = !force_stage && host_ptr_is_pinned(b.in + c0 * b.row_stride);
  bool out_pinned = !force_stage;
  size_t per_out = 0;
  for (int i = 0; i < b.n_outs; ++i) {
    per_out += b.outs[i].per_row * sizeof(float);
    if (b.outs[i].host && b.outs[i].per_row && !host_ptr_is_pinned(b.outs[i].host)) out_pinned = false;
  }
  const size_t per_in = (size_t)b.row_len * sizeof(float);
  int64_t chunk = std::max<int64_t>(1, (int64_t)(b.chunk_bytes / std::max<size_t>(1, per_in + per_out)));
  chunk = std::min(chunk, c1 - c0);
  CopyPool& pool = CopyPool::instance();
  for (int64_t a = c0; a < c1; a += chunk) {
    const int64_t n = std::min(chunk, c1 - a);
    Lane& ln = dc->lanes[dc->next_lane];
    dc->next_lane = (dc->next_lane + 1) % kLanes;
    int r = lane_retire(ln, true);  // the lane's previous chunk: off the device, on its way to the caller
    if (r) return r;
    if ((r = reserve(ln.in, std::max<size_t>(16, n * per_in)))) return r;
    size_t off[kMaxOuts], total_out = 0;
    for (int i = 0; i < b.n_outs; ++i) {
      off[i] = total_out;
      if (b.outs[i].host) total_out += ((size_t)n * b.outs[i].per_row * sizeof(float) + 255) & ~(size_t)255;
    }
    if ((r = reserve(ln.out, std::max<size_t>(16, total_out + 5 * 256)))) return r;
    const float* src = b.in + a * b.row_stride;
    if (in_pinned) {
      if ((r = copy_rows_h2d(static_cast<float*>(ln.in.p), src, n, b.row_len, b.row_stride, ln.stream))) return r;
    } else {
      if ((r = reserve(ln.hin, std::max<size_t>(16, n * per_in)))) return r;
      pool.wait(pool.copy2d_async(ln.hin.p, per_in, src, (size_t)b.row_stride * sizeof(float), per_in, (size_t)n));
      SS_CUDA(cudaMemcpyAsync(ln.in.p, ln.hin.p, n * per_in, cudaMemcpyHostToDevice, ln.stream));
    }
    float* d_outs[kMaxOuts];
    for (int i = 0; i < b.n_outs; ++i)
      d_outs[i] = b.outs[i].host ? reinterpret_cast<float*>(static_cast<char*>(ln.out.p) + off[i]) : nullptr;
    if ((r = fn(ln, a, n, static_cast<const float*>(ln.in.p), d_outs))) return r;
    for (auto& t : ln.out_tickets) pool.wait(t);  // the ring is about to be overwritten
    ln.out_tickets.clear();
    if (out_pinned) {
      for (int i = 0; i < b.n_outs; ++i)
        if (b.outs[i].host && b.outs[i].per_row)
          SS_CUDA(cudaMemcpyAsync(b.outs[i].host + (size_t)a * b.outs[i].per_row, d_outs[i],
                                  (size_t)n * b.outs[i].per_row * sizeof(float), cudaMemcpyDeviceToHost, ln.stream));
    } else if (total_out) {
      if ((r = reserve(ln.hout, total_out))) return r;
      SS_CUDA(cudaMemcpyAsync(ln.hout.p, ln.out.p, total_out, cudaMemcpyDeviceToHost, ln.stream));
      Pending& pd = ln.pending;
      pd.valid = true;
      pd.a = a;
      pd.n = n;
      pd.n_outs = b.n_outs;
      for (int i = 0; i < b.n_outs; ++i) {
        pd.outs[i] = b.outs[i];
        pd.off[i] = off[i];
      }
    }
  }
  return drain ? drain_lanes(dc) : SS_OK;
}

// Shards [0, n_items) in contiguous blocks over the plan's devices and runs fn(dc, first, last) for each
// block: on the calling thread (device guarded) when one device is enough, else on the devices'
// persistent worker threads.  `bytes_hint` keeps small calls on the first device alone.
template <class Fn>
static int run_on_devices(ss_plan* pl, int64_t n_items, size_t bytes_hint, Fn fn) {
  const int nd_all = (int)pl->devs.size();
  if (nd_all == 0) return fail(SS_ERR_NO_DEVICE, "plan has no CUDA device (host-only plan); there is no CPU fallback");
  const int nd = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)nd_all, n_items, (int64_t)(bytes_hint / (4u << 20)) + 1}));
  auto body = [pl, &fn](DeviceCtx* dc, int64_t a, int64_t b) -> int {
    int rc = ensure_device(pl, dc);
    if (!rc) rc = fn(dc, a, b);
    if (rc) {
      dc->err = g_err;
      abort_lanes(dc);
    }
    return rc;
  };
  std::vector<int> rc(nd, 0);
  if (nd == 1) {
    DeviceCtx* dc = pl->devs[0].get();
    DeviceGuard guard(dc->device);
    rc[0] = body(dc, 0, n_items);
  } else {
    if (pl->workers.empty())
      for (auto& dc : pl->devs) pl->workers.emplace_back(new DeviceWorker(dc->device));
    for (int i = 0; i < nd; ++i) {
      const int64_t a = n_items * i / nd, b = n_items * (i + 1) / nd;
      DeviceCtx* dc = pl->devs[i].get();
      pl->workers[i]->post([&body, dc, a, b] { return b > a ? body(dc, a, b) : 0; });
    }
    for (int i = 0; i < nd; ++i) rc[i] = pl->workers[i]->wait();
  }
  for (int i = 0; i < nd; ++i)
    if (rc[i]) {
      g_err = "device " + std::to_string(pl->devs[i]->device) + ": " + pl->devs[i]->err;
      return rc[i];
    }
  return SS_OK;
}

// scratch of ss_features_dev's standalone cmvn, one buffer per caller stream
static DevBuf* plan_stream_scratch(DeviceCtx* dc, void* stream) {
  std::lock_guard<std::mutex> lk(dc->scratch_mu);
  return &dc->dev_scratch[stream];
}

// ---- default context for the config-less processing functions: per (thread, device) -------------
struct DefaultCtx {
  cudaStream_t stream = nullptr;
  DevBuf a, b, c;
  int device = -1;
};
struct DefaultCtxSet {
  std::map<int, DefaultCtx> by_dev;
  ~DefaultCtxSet() {  // thread exit: give the device memory back (the runtime may already be gone at process exit)
    for (auto& kv : by_dev) {
      DeviceGuard g(kv.first);
      if (kv.second.stream && cudaStreamQuery(kv.second.stream) != cudaErrorCudartUnloading) {
        cudaStreamSynchronize(kv.second.stream);
        kv.second.a.release();
        kv.second.b.release();
        kv.second.c.release();
        cudaStreamDestroy(kv.second.stream);
      }
    }
    cudaGetLastError();
  }
};
static thread_local DefaultCtxSet g_ctx;
static int default_ctx(DefaultCtx** out) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(SS_ERR_NO_DEVICE, "no CUDA device visible (%s); there is no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  int dev = 0;
  SS_CUDA(cudaGetDevice(&dev));
  DefaultCtx& c = g_ctx.by_dev[dev];
  if (!c.stream) {
    SS_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    c.device = dev;
  }
  *out = &c;
  return SS_OK;
}

static int copy_rows_h2d(float* dst, const float* src, int64_t rows, int64_t len, int64_t row_stride,
                         cudaStream_t s) {
  if (rows == 0 || len == 0) return SS_OK;
  if (row_stride == len) {
    SS_CUDA(cudaMemcpyAsync(dst, src, (size_t)rows * len * sizeof(float), cudaMemcpyHostToDevice, s));
  } else if (row_stride < len || rows == 1) {  // overlapping rows (the frame tiles of one long clip): a 2-D copy needs pitch >= width
    for (int64_t r = 0; r < rows; ++r)
      SS_CUDA(cudaMemcpyAsync(dst + r * len, src + r * row_stride, (size_t)len * sizeof(float), cudaMemcpyHostToDevice, s));
  } else {
    SS_CUDA(cudaMemcpy2DAsync(dst, len * sizeof(float), src, row_stride * sizeof(float), len * sizeof(float),
                              rows, cudaMemcpyHostToDevice, s));
  }
  return SS_OK;
}

// host buffers -> device -> fn -> host buffers on the default context
template <class Fn>
static int host_unary(const float* in, size_t in_floats, float* out, size_t out_floats, Fn fn) {
  if (!in || !out) return fail(SS_ERR_INVALID_ARG, "null argument");
  DefaultCtx* c;
  int rc = default_ctx(&c);
  if (rc) return rc;
  if ((rc = reserve(c->a, std::max<size_t>(16, in_floats * 4)))) return rc;
  if ((rc = reserve(c->b, std::max<size_t>(16, out_floats * 4)))) return rc;
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
int ss_host_register(void* ptr, size_t bytes) {
  if (!ptr || !bytes) return fail(SS_ERR_INVALID_ARG, "null / empty range");
  cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) return cuda_fail(e, "cudaHostRegister");
  return SS_OK;
}
int ss_host_unregister(void* ptr) {
  if (!ptr) return SS_OK;
  SS_CUDA(cudaHostUnregister(ptr));
  return SS_OK;
}
int ss_host_is_pinned(const void* ptr) { return ptr && host_ptr_is_pinned(ptr) ? 1 : 0; }
int ss_copy_threads(void) { return CopyPool::instance().threads(); }
int ss_debug_copy2d(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t row_bytes, size_t rows) {
  if ((!dst || !src) && rows > 0 && row_bytes > 0) return fail(SS_ERR_INVALID_ARG, "null argument");
  if (rows > 1 && (dst_pitch < row_bytes)) return fail(SS_ERR_INVALID_ARG, "destination rows overlap");
  CopyPool& pool = CopyPool::instance();
  pool.wait(pool.copy2d_async(dst, dst_pitch, src, src_pitch, row_bytes, rows));
  return SS_OK;
}
int ss_release_scratch(void) {
  ScratchRegistry::instance().clear();
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
  pl->fast512c = pl->fft_ok && fast512c_supported(pl->d, pl->mb);
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
  // device state (tables, streams, pinned rings, worker threads) is created by the first call that
  // runs on a device: creating a plan over all GPUs costs nothing on the ones a caller never reaches
  for (int dv : devs) {
    pl->devs.emplace_back(new DeviceCtx());
    pl->devs.back()->device = dv;
  }
  *out = pl.release();
  return SS_OK;
}

void ss_plan_destroy(ss_plan* pl) {
  if (!pl) return;
  pl->workers.clear();  // joins the worker threads
  for (auto& dcp : pl->devs) {
    DeviceCtx& dc = *dcp;
    if (!dc.ready) continue;
    DeviceGuard guard(dc.device);
    for (auto& ln : dc.lanes) {
      if (ln.stream) {
        cudaStreamSynchronize(ln.stream);
        cudaStreamDestroy(ln.stream);
      }
      for (auto& t : ln.out_tickets) CopyPool::instance().wait(t);
      ln.in.release();
      ln.out.release();
      ln.scratch.release();
      ln.mem.release();
      ln.hin.release();
      ln.hout.release();
      ln.hmem.release();
    }
    for (auto& ev : dc.lane_evt)
      if (ev) cudaEventDestroy(ev);
    dc.tables.release();
    dc.clip_mfcc.release();
    for (auto& kv : dc.dev_scratch) kv.second.release();
    if (dc.fast512c) fast512c_free_tables(dc.fast512c);
    if (dc.stft400m) stft400m_free_tables(dc.stft400m);
    if (dc.stft2048) stft2048_free_tables(dc.stft2048);
    if (dc.stft2048m) stft2048m_free_tables(dc.stft2048m);
  }
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
  info->fast_path = pl->fast512c && !(d.cfg.flags & (SS_FLAG_FORCE_GENERIC | SS_FLAG_FRAMING_LITERAL));
  info->fast_stft = (pl->stft400m || pl->stft2048 || pl->stft2048m || stft512_supported(d)) && !(d.cfg.flags & SS_FLAG_FORCE_GENERIC);
  return SS_OK;
}

int ss_plan_set_tuning(ss_plan* pl, int64_t tile_frames, int64_t chunk_bytes) {
  if (!pl) return fail(SS_ERR_INVALID_ARG, "null plan");
  if (tile_frames < 0 || chunk_bytes < 0) return fail(SS_ERR_INVALID_ARG, "negative tuning value");
  pl->tile_frames = tile_frames;
  if (chunk_bytes > 0) pl->chunk_bytes = (size_t)chunk_bytes;
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
  if (o.mfcc && (o.mfcc_cmvn_mode < 0 || o.mfcc_cmvn_mode > 2)) return fail(SS_ERR_INVALID_ARG, "cmvn_mode must be 0, 1 or 2");
  const size_t per_frame_out =
      ((o.mfe ? d.nf : 0) + (o.energy ? 1 : 0) + (o.lmfe ? d.nf : 0) + (o.ssc ? d.nf : 0) + (o.mfcc ? d.nc : 0)) * sizeof(float);
  const size_t per_clip_out = (size_t)T * per_frame_out;
  const size_t per_clip_in = (size_t)clip_len * sizeof(float);
  const size_t kChunk = pl->chunk_bytes;
  // a clip larger than a chunk (long-form streams: 230 MB in, 135 MB out per hour) is cut into frame
  // tiles with an L - S halo so that its own H2D, kernels and D2H overlap (long-context row)
  const int64_t tile_env = pl->tile_frames;
  const bool tiled = tile_env > 0 ? T > tile_env : per_clip_in + per_clip_out > kChunk + kChunk / 2;
  return run_on_devices(pl, n_clips, (size_t)n_clips * (per_clip_in + per_clip_out), [&](DeviceCtx* dc, int64_t c0, int64_t c1) -> int {
    if (!tiled) {
      RowBatch b;
      b.in = pcm;
      b.row_len = clip_len;
      b.row_stride = row_stride;
      b.chunk_bytes = kChunk;
      b.n_outs = 5;
      b.outs[0] = {o.mfe, (size_t)T * d.nf};
      b.outs[1] = {o.energy, (size_t)T};
      b.outs[2] = {o.lmfe, (size_t)T * d.nf};
      b.outs[3] = {o.ssc, (size_t)T * d.nf};
      b.outs[4] = {o.mfcc, (size_t)T * d.nc};
      return pipeline_rows(dc, b, c0, c1, [&](Lane& ln, int64_t, int64_t n, const float* d_in, float* const* dv) -> int {
        ss_feature_outputs dvo = o;
        dvo.mfe = dv[0];
        dvo.energy = dv[1];
        dvo.lmfe = dv[2];
        dvo.ssc = dv[3];
        dvo.mfcc = dv[4];
        return features_dev(pl, dc, ln.stream, d_in, n, clip_len, clip_len, dvo, &ln.scratch);
      }, true);
    }
    // ---- frame tiles: tile j of a clip = frames [j Tt, (j+1) Tt) = samples [j Tt S, j Tt S + (Tt-1) S + L).
    // The frame count of a tile is handed to the kernels explicitly: the reference's own count has no "+1"
    // (floor((N - L) / S), processing.rs:101) and would give Tt - 1 for exactly these samples.
    const int64_t Tt = tile_env > 0 ? tile_env : std::max<int64_t>(64, (int64_t)(kChunk / (d.S * sizeof(float) + per_frame_out)));
    const bool cmvn = o.mfcc && o.mfcc_cmvn_mode != SS_CMVN_NONE;
    for (int64_t c = c0; c < c1; ++c) {
      const float* clip = pcm + c * row_stride;
      if (cmvn) {
        // the statistics span the whole clip: tiles park their cepstra on the device, one cmvn pass at the end
        int r = reserve(dc->clip_mfcc, (size_t)T * d.nc * sizeof(float));
        if (r) return r;
        for (auto& ln : dc->lanes) SS_CUDA(cudaStreamSynchronize(ln.stream));  // previous clip's D2H of clip_mfcc
      }
      const int64_t n_full = T / Tt, rem = T - n_full * Tt;
      for (int pass = 0; pass < 2; ++pass) {  // the full tiles, then the shorter last one
        const int64_t frames = pass == 0 ? Tt : rem, j_base = pass == 0 ? 0 : n_full, n_tiles = pass == 0 ? n_full : 1;
        if (frames == 0 || n_tiles == 0) continue;
        const size_t f0 = (size_t)c * T + (size_t)j_base * Tt;  // first frame of the batch in the caller's arrays
        RowBatch b;
        b.in = clip + j_base * Tt * d.S;
        b.row_len = (frames - 1) * d.S + d.L;
        b.row_stride = Tt * d.S;
        b.chunk_bytes = 1;  // one tile per chunk
        b.n_outs = 5;
        auto at = [&](float* host, size_t per_frame) -> HostOut {
          HostOut h;
          if (host) {
            h.host = host + f0 * per_frame;
            h.per_row = (size_t)frames * per_frame;
          }
          return h;
        };
        b.outs[0] = at(o.mfe, d.nf);
        b.outs[1] = at(o.energy, 1);
        b.outs[2] = at(o.lmfe, d.nf);
        b.outs[3] = at(o.ssc, d.nf);
        b.outs[4] = cmvn ? HostOut() : at(o.mfcc, d.nc);
        const int64_t tile_len = b.row_len;
        int r = pipeline_rows(dc, b, 0, n_tiles, [&](Lane& ln, int64_t j, int64_t, const float* d_in, float* const* dv) -> int {
          ss_feature_outputs dvo = o;
          dvo.mfe = dv[0];
          dvo.energy = dv[1];
          dvo.lmfe = dv[2];
          dvo.ssc = dv[3];
          dvo.mfcc = cmvn ? static_cast<float*>(dc->clip_mfcc.p) + (size_t)(j_base + j) * Tt * d.nc : dv[4];
          dvo.mfcc_cmvn_mode = SS_CMVN_NONE;
          return features_dev(pl, dc, ln.stream, d_in, 1, tile_len, tile_len, dvo, &ln.scratch, T, j_base + j == 0, frames);
        }, false);
        if (r) return r;
      }
      if (cmvn) {
        // every lane holds tiles of this clip: join them on lane 0, normalise, hand the block back
        Lane& l0 = dc->lanes[0];
        for (int i = 1; i < kLanes; ++i) {
          SS_CUDA(cudaEventRecord(dc->lane_evt[i], dc->lanes[i].stream));
          SS_CUDA(cudaStreamWaitEvent(l0.stream, dc->lane_evt[i], 0));
        }
        int r = reserve(l0.scratch, cmvn_scratch_bytes(1, T, d.nc));
        if (r) return r;
        float* m = static_cast<float*>(dc->clip_mfcc.p);
        int e = launch_cmvn(m, 1, T, d.nc, o.mfcc_cmvn_mode == SS_CMVN_MEAN_VAR, m, static_cast<float*>(l0.scratch.p), l0.stream);
        if (e) return cuda_fail((cudaError_t)e, "cmvn kernel launch");
        // cudaMemcpyAsync stages pageable destinations itself (18.7 MB per hour of audio: not worth a ring)
        SS_CUDA(cudaMemcpyAsync(o.mfcc + (size_t)c * T * d.nc, m, (size_t)T * d.nc * sizeof(float), cudaMemcpyDeviceToHost, l0.stream));
      }
    }
    return drain_lanes(dc);
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
  DeviceCtx* dc = pl->devs[dev_index].get();
  DeviceGuard guard(dc->device);
  if ((rc = ensure_device(pl, dc))) return rc;
  return features_dev(pl, dc, (cudaStream_t)stream, d_pcm, n_clips, clip_len, row_stride, *d_outs,
                      plan_stream_scratch(dc, stream));
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
    rc = run_on_devices(pl, ch, (size_t)ch * (per_row_in + per_row_out), [&](DeviceCtx* dc, int64_t c0, int64_t c1) -> int {
      RowBatch b;
      b.in = pcm;
      b.row_len = ttd;
      b.row_stride = row_stride;
      b.chunk_bytes = pl->chunk_bytes;
      b.n_outs = 1;
      b.outs[0] = {out_ri ? out_ri : out_mel, per_row_out / sizeof(float)};
      return pipeline_rows(dc, b, c0, c1, [&](Lane& ln, int64_t a, int64_t n, const float* d_in, float* const* dv) -> int {
        int r = reserve(ln.mem, std::max<size_t>(16, memlen * sizeof(float)));
        if (r) return r;
        if ((r = reserve(ln.hmem, std::max<size_t>(16, memlen * sizeof(float))))) return r;
        // the chunk's first row continues the stream of the row before it (Q13): restate the
        // overlap memory on the host from the input itself -- no device collective needed.
        // (the lane was retired before this chunk started, so its halo slot is free)
        float* hm = static_cast<float*>(ln.hmem.p);
        host_mem_before_row(pl, pcm, ttd, row_stride, a, mem0.data(), hm);
        SS_CUDA(cudaMemcpyAsync(ln.mem.p, hm, memlen * sizeof(float), cudaMemcpyHostToDevice, ln.stream));
        FrameParams p;
        stft_params(pl, dc, &p, d_in, n, ttd, ttd);
        p.mem_in = static_cast<float*>(ln.mem.p);
        p.log_mode = log_mode;
        if (out_ri) p.stft_out = reinterpret_cast<float2*>(dv[0]);
        else p.mel_out = dv[0];
        int e = launch_stft_any(pl, dc, p, ln.stream);
        if (e) return cuda_fail((cudaError_t)e, "stft kernel launch");
        return SS_OK;
      }, true);
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
  if (!d_pcm && ch * ttd > 0) return fail(SS_ERR_INVALID_ARG, "null pcm");
  DeviceCtx* dc = pl->devs[dev_index].get();
  DeviceGuard guard(dc->device);
  if ((rc = ensure_device(pl, dc))) return rc;
  FrameParams p;
  stft_params(pl, dc, &p, d_pcm, ch, ttd, row_stride);
  p.mem_in = d_mem_in;
  p.mem_out = d_mem_out;
  p.log_mode = log_mode;
  p.stft_out = reinterpret_cast<float2*>(d_out_ri);
  p.mel_out = d_out_mel;
  if (ch > 0 && p.Tp > 0) {
    int e = launch_stft_any(pl, dc, p, stream);
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
  {
    const int e = launch_stft512_any(pl, dc, p, stream);
    if (e != -1000) return e;
  }
  if (pl->stft2048 && dc->stft2048 && pl->d.librosa && p.mel_out && !(pl->d.cfg.flags & SS_FLAG_FORCE_GENERIC)) {
    DevBuf* sb = plan_stream_scratch(dc, stream);
    if (reserve(*sb, stft2048_power_scratch_bytes(p, nullptr))) return (int)cudaErrorMemoryAllocation;
    const int e = launch_stft2048_power(p, dc->stft2048, static_cast<float*>(sb->p), dc->sm_count, stream);
    if (e != -1000) return e;
  }
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
  if (!d_pcm) return fail(SS_ERR_INVALID_ARG, "null pcm");
  DeviceCtx* dc = pl->devs[dev_index].get();
  DeviceGuard guard(dc->device);
  if ((rc = ensure_device(pl, dc))) return rc;
  FrameParams p;
  librosa_params(pl, dc, &p, d_pcm, ch, ttd, row_stride, d_out, db_mode);
  int e = launch_librosa_any(pl, dc, p, stream);
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
  return run_on_devices(pl, ch, (size_t)ch * (per_row_in + per_row_out), [&](DeviceCtx* dc, int64_t c0, int64_t c1) -> int {
    RowBatch b;
    b.in = pcm;
    b.row_len = ttd;
    b.row_stride = row_stride;
    b.chunk_bytes = pl->chunk_bytes;
    b.n_outs = 1;
    b.outs[0] = {out, per_row_out / sizeof(float)};
    return pipeline_rows(dc, b, c0, c1, [&](Lane& ln, int64_t, int64_t n, const float* d_in, float* const* dv) -> int {
      FrameParams p;
      librosa_params(pl, dc, &p, d_in, n, ttd, ttd, dv[0], db_mode);
      int e = launch_librosa_any(pl, dc, p, ln.stream);
      if (e) return cuda_fail((cudaError_t)e, "librosa mel kernel launch");
      return SS_OK;
    }, true);
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
    if ((rc = reserve(c->a, (size_t)m * n * 4))) return rc;
    if ((rc = reserve(c->b, (size_t)m * n * 4))) return rc;
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
  if ((rc = reserve(c->a, (size_t)n * 4))) return rc;
  if ((rc = reserve(c->b, (size_t)T * L * 4))) return rc;
  SS_CUDA(cudaMemcpyAsync(c->a.p, in, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  if ((rc = ss_stack_frames_dev(c->stream, static_cast<float*>(c->a.p), n, L, S, zero_padding, flags,
                                static_cast<float*>(c->b.p), T)))
    return rc;
  SS_CUDA(cudaMemcpyAsync(out, c->b.p, (size_t)T * L * 4, cudaMemcpyDeviceToHost, c->stream));
  SS_CUDA(cudaStreamSynchronize(c->stream));
  return SS_OK;
}

// power_spectrum has no config either: FFT tables are cached per (device, n_fft), process-wide.  The
// table is uploaded on the stream of the call that first needs it; an event orders every later stream
// behind that upload.
struct PsEntry {
  FftPlanHost plan;
  DevBuf tw;
  cudaEvent_t ready = nullptr;
  void* upload_stream = nullptr;
};
static std::mutex g_ps_mu;
static std::map<std::pair<int, int>, std::unique_ptr<PsEntry>> g_ps;
static int ps_tables(int64_t n_fft, void* stream, DevTables* tab) {
  if (n_fft < 2 || n_fft > 8192) return fail(SS_ERR_UNSUPPORTED, "n_fft=%lld outside [2, 8192]", (long long)n_fft);
  int dev = 0;
  SS_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_ps_mu);
  auto& slot = g_ps[{dev, (int)n_fft}];
  if (!slot) {
    std::unique_ptr<PsEntry> e(new PsEntry());
    if (!build_fft_plan((int)n_fft, &e->plan)) return fail(SS_ERR_UNSUPPORTED, "n_fft=%lld has a prime factor > 61", (long long)n_fft);
    int rc = reserve(e->tw, e->plan.tw.size() * sizeof(float));
    if (rc) return rc;
    SS_CUDA(cudaEventCreateWithFlags(&e->ready, cudaEventDisableTiming));
    SS_CUDA(cudaMemcpyAsync(e->tw.p, e->plan.tw.data(), e->plan.tw.size() * sizeof(float), cudaMemcpyHostToDevice,
                            (cudaStream_t)stream));
    SS_CUDA(cudaEventRecord(e->ready, (cudaStream_t)stream));
    e->upload_stream = stream;
    slot = std::move(e);
  } else if (slot->upload_stream != stream) {
    SS_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, slot->ready, 0));
  }
  *tab = DevTables();
  tab->tw = static_cast<const float2*>(slot->tw.p);
  tab->n_stages = slot->plan.n_stages;
  for (int i = 0; i < kMaxStages; ++i) tab->radix[i] = slot->plan.radix[i];
  return SS_OK;
}

static StreamScratch* stream_scratch(void* stream) {
  int dev = 0;
  cudaGetDevice(&dev);
  return ScratchRegistry::instance().get(dev, stream);
}

int ss_power_spectrum_dev(void* stream, const float* d_frames, int64_t rows, int64_t L, int64_t n_fft, float* d_out) {
  if (!d_frames || !d_out) return fail(SS_ERR_INVALID_ARG, "null argument");
  if (rows < 0 || L <= 0) return fail(SS_ERR_INVALID_ARG, "bad shape");
  if (L > n_fft) return fail(SS_ERR_INVALID_ARG, "frame_len %lld > n_fft %lld (ndrustfft panics, processing.rs:159)", (long long)L, (long long)n_fft);
  if (rows == 0) return SS_OK;
  static const bool no_fast = getenv("SS_PSPEC_GENERIC") && atoi(getenv("SS_PSPEC_GENERIC"));  // A/B switch
  if (!no_fast && pspec512_supported(L, n_fft)) {
    const int e = launch_pspec512(d_frames, rows, L, d_out, stream);
    if (e) return cuda_fail((cudaError_t)e, "power_spectrum launch");
    return SS_OK;
  }
  FrameParams p;
  int rc = ps_tables(n_fft, stream, &p.tab);
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
  if ((rc = reserve(c->a, (size_t)rows * L * 4))) return rc;
  if ((rc = reserve(c->b, (size_t)rows * F * 4))) return rc;
  SS_CUDA(cudaMemcpyAsync(c->a.p, frames, (size_t)rows * L * 4, cudaMemcpyHostToDevice, c->stream));
  if ((rc = ss_power_spectrum_dev(c->stream, static_cast<float*>(c->a.p), rows, L, n_fft, static_cast<float*>(c->b.p))))
    return rc;
  SS_CUDA(cudaMemcpyAsync(out, c->b.p, (size_t)rows * F * 4, cudaMemcpyDeviceToHost, c->stream));
  SS_CUDA(cudaStreamSynchronize(c->stream));
  return SS_OK;
}

// processing::log_power_spectrum (processing.rs:194-214): private and unused in the reference
int ss_log_power_spectrum_dev(void* stream, const float* d_frames, int64_t rows, int64_t L, int64_t n_fft, int normalize,
                              float* d_out) {
  int rc = ss_power_spectrum_dev(stream, d_frames, rows, L, n_fft, d_out);
  if (rc || rows == 0) return rc;
  DevBuf& sc = stream_scratch(stream)->logps;
  if ((rc = reserve(sc, 16))) return rc;
  int e = launch_log_power(d_out, rows * (n_fft / 2 + 1), normalize, sc.p, stream);
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
  if ((rc = reserve(c->a, (size_t)rows * L * 4))) return rc;
  if ((rc = reserve(c->b, (size_t)rows * F * 4))) return rc;
  SS_CUDA(cudaMemcpyAsync(c->a.p, frames, (size_t)rows * L * 4, cudaMemcpyHostToDevice, c->stream));
  if ((rc = ss_log_power_spectrum_dev(c->stream, static_cast<float*>(c->a.p), rows, L, n_fft, normalize, static_cast<float*>(c->b.p))))
    return rc;
  SS_CUDA(cudaMemcpyAsync(out, c->b.p, (size_t)rows * F * 4, cudaMemcpyDeviceToHost, c->stream));
  SS_CUDA(cudaStreamSynchronize(c->stream));
  return SS_OK;
}

int ss_cmvn_dev(void* stream, const float* d_in, int64_t n_mats, int64_t rows, int64_t cols, int varnorm,
                float* d_out) {
  if (!d_in || !d_out) return fail(SS_ERR_INVALID_ARG, "null argument");
  if (n_mats < 0 || cols < 0) return fail(SS_ERR_INVALID_ARG, "bad shape");
  if (rows <= 0) return fail(SS_ERR_INVALID_ARG, "cmvn: zero rows (mean_axis(..).unwrap() panics, processing.rs:270)");
  if (cols > 256) return fail(SS_ERR_UNSUPPORTED, "cmvn: more than 256 columns");
  if (n_mats == 0 || cols == 0) return SS_OK;
  DevBuf& sc = stream_scratch(stream)->cmvn;
  int rc = reserve(sc, cmvn_scratch_bytes(n_mats, rows, cols));
  if (rc) return rc;
  int e = launch_cmvn(d_in, n_mats, rows, cols, varnorm, d_out, static_cast<float*>(sc.p), stream);
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
  if ((rc = reserve(c->a, bytes))) return rc;
  if ((rc = reserve(c->b, bytes))) return rc;
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
int ss_cmvnw_dev(void* stream, const float* d_in, int64_t n_mats, int64_t rows, int64_t cols, int64_t win, int varnorm,
                 float* d_out) {
  if (!d_in || !d_out) return fail(SS_ERR_INVALID_ARG, "null argument");
  if (n_mats < 0 || rows < 0 || cols < 0) return fail(SS_ERR_INVALID_ARG, "bad shape");
  if (win < 1 || win % 2 != 1) return fail(SS_ERR_INVALID_ARG, "Windows size must be odd! (processing.rs:327)");
  if (cols > 256 || n_mats > 65535 || win > (1 << 24)) return fail(SS_ERR_UNSUPPORTED, "cmvnw: shape outside this build's limits");
  if (n_mats == 0 || rows == 0 || cols == 0) return SS_OK;
  float* scratch = nullptr;
  if (varnorm) {
    DevBuf& sc = stream_scratch(stream)->cmvnw;
    int rc = reserve(sc, (size_t)n_mats * rows * cols * sizeof(float));
    if (rc) return rc;
    scratch = static_cast<float*>(sc.p);
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
