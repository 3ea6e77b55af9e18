// Host-side runtime of the C-ABI: device / pinned buffers, the RAII device guard, the persistent
// per-device worker threads of a multi-device plan, the staging copy pool that feeds pageable caller
// memory through pinned rings, and the (device, stream)-keyed scratch registry of the *_dev entry points.
#pragma once
#include <cuda_runtime.h>
#include <emmintrin.h>

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace ss {

// Restores the caller's current device on scope exit (every host entry point that switches device).
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// Grow-only device allocation; remembers the device it lives on.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int device = -1;
  cudaError_t reserve_raw(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    release();
    cudaGetDevice(&device);
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      want = bytes;
      e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) {
      p = nullptr;
      return e;
    }
    cap = want;
    return cudaSuccess;
  }
  void release() {
    if (p) {
      if (device >= 0) {
        DeviceGuard g(device);
        cudaFree(p);
      } else {
        cudaFree(p);
      }
    }
    p = nullptr;
    cap = 0;
  }
};

// Grow-only pinned host allocation (portable: usable from every device of the plan).
struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve_raw(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    release();
    cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) {
      p = nullptr;
      return e;
    }
    cap = bytes;
    return cudaSuccess;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

// true when the driver can DMA straight from / to `ptr` (cudaHostAlloc / cudaHostRegister memory)
inline bool host_ptr_is_pinned(const void* ptr) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// memcpy with non-temporal stores: a staging copy never re-reads its destination, so the lines should
// not be read for ownership nor evict the rest of the cache (3 -> 2 memory passes per byte moved)
inline void stream_copy(char* dst, const char* src, size_t n) {
  if (n < (64u << 10)) {
    std::memcpy(dst, src, n);
    return;
  }
  const size_t head = (64 - (reinterpret_cast<uintptr_t>(dst) & 63)) & 63;
  if (head) {
    std::memcpy(dst, src, head);
    dst += head;
    src += head;
    n -= head;
  }
  const size_t blocks = n / 64;
  static const bool pf = !(getenv("SS_COPY_PREFETCH") && atoi(getenv("SS_COPY_PREFETCH")) == 0);
  for (size_t i = 0; i < blocks; ++i) {
    if (pf) _mm_prefetch(src + 1024, _MM_HINT_NTA);
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src) + 0);
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src) + 1);
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src) + 2);
    const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src) + 3);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst) + 0, a);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst) + 1, b);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst) + 2, c);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst) + 3, d);
    src += 64;
    dst += 64;
  }
  _mm_sfence();
  if (n & 63) std::memcpy(dst, src, n & 63);
}

// ---- staging copy pool ------------------------------------------------------------------------
// A caller's pageable ndarray cannot be DMA'd; cudaMemcpyAsync on it degrades to a synchronous,
// single-threaded bounce through the driver.  Instead the chunks are memcpy'd into / out of the lanes'
// pinned rings by a few persistent threads (the copies of neighbouring chunks overlap the DMA and the
// kernels), which keeps a pageable call close to the pinned rate.
class CopyPool {
 public:
  struct Ticket {
    std::atomic<int> left{0};
    std::mutex m;
    std::condition_variable cv;
  };
  using TicketPtr = std::shared_ptr<Ticket>;

  static CopyPool& instance() {
    static CopyPool* pool = new CopyPool();  // leaked on purpose: threads must outlive static destructors
    return *pool;
  }
  int threads() const { return (int)th_.size(); }

  // rows x row_bytes from src (stride src_pitch) to dst (stride dst_pitch); returns immediately
  TicketPtr copy2d_async(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t row_bytes, size_t rows) {
    auto t = std::make_shared<Ticket>();
    if (rows == 0 || row_bytes == 0) return t;
    if (rows * row_bytes <= (256u << 10)) {  // single clips: waking the pool costs more than the copy
      for (size_t r = 0; r < rows; ++r) std::memcpy((char*)dst + r * dst_pitch, (const char*)src + r * src_pitch, row_bytes);
      return t;
    }
    std::vector<Task> tasks;
    const size_t piece = 2u << 20;
    if ((dst_pitch == row_bytes && src_pitch == row_bytes) || rows == 1) {
      const size_t total = rows * row_bytes;
      for (size_t o = 0; o < total; o += piece)
        tasks.push_back({(char*)dst + o, (const char*)src + o, std::min(piece, total - o), 1, 0, 0, t});
    } else if (row_bytes >= piece) {  // long rows (the frame tiles of one stream): every row in pieces
      for (size_t r = 0; r < rows; ++r)
        for (size_t o = 0; o < row_bytes; o += piece)
          tasks.push_back({(char*)dst + r * dst_pitch + o, (const char*)src + r * src_pitch + o, std::min(piece, row_bytes - o), 1, 0, 0, t});
    } else {
      const size_t rows_per = std::max<size_t>(1, piece / row_bytes);
      for (size_t r = 0; r < rows; r += rows_per)
        tasks.push_back({(char*)dst + r * dst_pitch, (const char*)src + r * src_pitch, row_bytes,
                         std::min(rows_per, rows - r), dst_pitch, src_pitch, t});
    }
    t->left.store((int)tasks.size());
    {
      std::lock_guard<std::mutex> lk(m_);
      for (auto& k : tasks) q_.push_back(std::move(k));
    }
    cv_.notify_all();
    return t;
  }
  // the waiting thread helps with whatever is queued (its own pieces or a neighbour's)
  void wait(const TicketPtr& t) {
    if (!t) return;
    while (t->left.load(std::memory_order_acquire) > 0) {
      Task k;
      if (pop(&k)) run(k);
      else {
        std::unique_lock<std::mutex> lk(t->m);
        t->cv.wait_for(lk, std::chrono::microseconds(200), [&] { return t->left.load(std::memory_order_acquire) <= 0; });
      }
    }
  }

 private:
  struct Task {
    char* dst;
    const char* src;
    size_t bytes, rows, dpitch, spitch;
    TicketPtr t;
  };
  CopyPool() {
    unsigned hw = std::thread::hardware_concurrency();
    int n = (int)std::min(16u, std::max(2u, hw - 1));
    if (const char* e = getenv("SS_COPY_THREADS")) n = std::max(0, atoi(e));
    for (int i = 0; i < n; ++i) th_.emplace_back([this] { loop(); });
    for (auto& t : th_) t.detach();
  }
  bool pop(Task* k) {
    std::lock_guard<std::mutex> lk(m_);
    if (q_.empty()) return false;
    *k = std::move(q_.front());
    q_.pop_front();
    return true;
  }
  static void run(const Task& k) {
    static const bool nt = !(getenv("SS_COPY_NT") && atoi(getenv("SS_COPY_NT")) == 0);
    for (size_t r = 0; r < k.rows; ++r) {
      if (nt) stream_copy(k.dst + r * k.dpitch, k.src + r * k.spitch, k.bytes);
      else std::memcpy(k.dst + r * k.dpitch, k.src + r * k.spitch, k.bytes);
    }
    if (k.t->left.fetch_sub(1, std::memory_order_acq_rel) == 1) {
      std::lock_guard<std::mutex> lk(k.t->m);
      k.t->cv.notify_all();
    }
  }
  void loop() {
    for (;;) {
      Task k;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return !q_.empty(); });
        k = std::move(q_.front());
        q_.pop_front();
      }
      run(k);
    }
  }
  std::mutex m_;
  std::condition_variable cv_;
  std::deque<Task> q_;
  std::vector<std::thread> th_;
};

// ---- one persistent worker thread per device of a multi-device plan -----------------------------
// (the first version spawned a std::thread per device on every call)
class DeviceWorker {
 public:
  explicit DeviceWorker(int device) : device_(device), th_([this] { loop(); }) {}
  ~DeviceWorker() {
    {
      std::lock_guard<std::mutex> lk(m_);
      quit_ = true;
    }
    cv_.notify_all();
    if (th_.joinable()) th_.join();
  }
  void post(std::function<int()> fn) {
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = std::move(fn);
      has_job_ = true;
      done_ = false;
    }
    cv_.notify_all();
  }
  int wait() {
    std::unique_lock<std::mutex> lk(m_);
    cv_.wait(lk, [&] { return done_; });
    return rc_;
  }

 private:
  void loop() {
    cudaSetDevice(device_);
    for (;;) {
      std::function<int()> fn;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return has_job_ || quit_; });
        if (quit_) return;
        fn = std::move(job_);
        has_job_ = false;
      }
      int rc = fn();
      {
        std::lock_guard<std::mutex> lk(m_);
        rc_ = rc;
        done_ = true;
      }
      cv_.notify_all();
    }
  }
  int device_;
  std::mutex m_;
  std::condition_variable cv_;
  std::function<int()> job_;
  bool has_job_ = false, done_ = true, quit_ = false;
  int rc_ = 0;
  std::thread th_;
};

// ---- scratch of the config-less *_dev entry points, keyed by (device, stream) --------------------
// Two calls on different streams (or devices) never share a buffer; calls on one stream are ordered by
// the stream itself.  Entries live until ss_release_scratch() or process exit.
struct StreamScratch {
  DevBuf cmvn, cmvnw, logps;
};
class ScratchRegistry {
 public:
  static ScratchRegistry& instance() {
    static ScratchRegistry* r = new ScratchRegistry();
    return *r;
  }
  StreamScratch* get(int device, void* stream) {
    std::lock_guard<std::mutex> lk(m_);
    auto& e = map_[{device, stream}];
    if (!e) e.reset(new StreamScratch());
    return e.get();
  }
  void clear() {
    std::lock_guard<std::mutex> lk(m_);
    for (auto& kv : map_) {
      kv.second->cmvn.release();
      kv.second->cmvnw.release();
      kv.second->logps.release();
    }
    map_.clear();
  }

 private:
  std::mutex m_;
  std::map<std::pair<int, void*>, std::unique_ptr<StreamScratch>> map_;
};

}  // namespace ss
