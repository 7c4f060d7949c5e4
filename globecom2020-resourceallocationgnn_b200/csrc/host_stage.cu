// Host-side staging for the reference-facing entry points (BS.predict / BS.train_dnn with numpy inputs).
//
// The reference hands Keras a dict of ~3N+1 host arrays per call (BS_brain.py:495-504, :724-728): fp64, one array per
// node slot, and the adjacency as the dense Kronecker matrix kron(Adj, I_F) (:492-493).  Once the device step takes
// ~80 us, repacking those arrays on ONE host thread (numpy slice assignments into a pinned tensor) is the wall-clock
// sink of the call.  Here every input tensor is described by a few strided "views" of the caller's memory; a small
// persistent worker pool gathers + converts (fp64 -> fp32) + validates (adjacency in {0,1}) them straight into pinned
// staging, chunk by chunk, while the calling thread watches per-tensor completion and enqueues the host-to-device
// copy of a tensor the moment its last chunk lands -- so the DMA of one tensor overlaps the packing of the next.
// No device synchronisation happens until the single one at the end of the call.
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "host_stage.cuh"

namespace v2v {

namespace {

struct Chunk {
  const v2v_host_view* view;
  float* dst;            // base of the destination tensor (pinned)
  long r0, r1;           // row range of the view
  int tensor;            // completion counter index
  int check;             // kCheckBinary / kCheckNonzero / 0
  int pack_N;            // > 0: rows r0..r1 are whole graphs of pack_N nodes; pack their adjacency bit masks too
  uint32_t* in_mask;     // pinned [B][N] words (pack_N <= 32)
  uint32_t* out_mask;
};

}  // namespace

// 32 x 32 bit-matrix transpose, LSB-first: on return bit n of a[m] is the former bit m of a[n]
static inline void transpose32(uint32_t a[32]) {
  uint32_t m = 0x0000ffffu;
  for (int j = 16; j != 0; j >>= 1, m ^= (m << j)) {
    for (int k = 0; k < 32; k = (k + j + 1) & ~j) {
      const uint32_t t = ((a[k] >> j) ^ a[k + j]) & m;
      a[k] ^= t << j;
      a[k + j] ^= t;
    }
  }
}

// Copy into pinned staging with non-temporal stores: the destination is read next by the copy engine, not by a core, so
// there is no point in pulling its cache lines in (write-allocate) or in keeping them.  dst 16-byte aligned, n floats.
static inline void stream_copy_f32(float* dst, const float* src, size_t n) {
#if defined(__SSE2__)
  static const bool on = [] { const char* e = getenv("V2V_HOST_NT"); return !(e && e[0] == '0'); }();
  if (on && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0 && n >= 1024) {
    size_t i = 0;
    for (; i + 16 <= n; i += 16) {
      const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
      const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 4));
      const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 8));
      const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 12));
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 4), b);
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 8), c);
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 12), d);
    }
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
    return;
  }
#endif
  memcpy(dst, src, n * sizeof(float));
}

namespace {

struct Job {                                        // one host_stage_run call; kept alive by whoever still looks at it
  std::vector<Chunk> chunks;
  std::atomic<int> next{0}, active{0};
  std::atomic<int> remaining[kHostStageMaxTensors];
  std::atomic<int> flags[kHostStageMaxTensors];      // kFlagNonbinary | kFlagNonzero found in the tensor
};

class StagePool {
 public:
  static StagePool& get() {
    static StagePool* p = new StagePool();      // intentionally leaked: workers outlive static destructors
    return *p;
  }
  int threads() const { return n_threads_; }

  void submit(const std::shared_ptr<Job>& job) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      job_ = job;
      gen_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
  }
  // runs one work item if any is left; false when the queue is empty
  static bool run_one(Job& j) {
    const int n = (int)j.chunks.size();
    if (j.next.load(std::memory_order_relaxed) >= n) return false;
    const int i = j.next.fetch_add(1, std::memory_order_acq_rel);
    if (i >= n) return false;
    const Chunk& c = j.chunks[i];
    if (const int f = run_chunk(c)) j.flags[c.tensor].fetch_or(f, std::memory_order_relaxed);
    j.remaining[c.tensor].fetch_sub(1, std::memory_order_acq_rel);
    j.active.fetch_sub(1, std::memory_order_acq_rel);
    return true;
  }
  static void drain(Job& j) {
    while (run_one(j)) {}
  }
  static void cpu_relax() {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
  }

 private:
  StagePool() {
    int hw = (int)std::thread::hardware_concurrency();
    if (hw <= 0) hw = 4;
    int local_world = 1;
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) local_world = std::max(1, atoi(e));
    // the caller works as well: with several ranks per host every rank gets its share of the cores minus its own thread
    int n = local_world > 1 ? std::min(8, std::max(1, hw / local_world)) - 1 : std::min(8, std::max(1, hw / 2));
    if (const char* e = getenv("V2V_HOST_THREADS")) n = std::max(0, std::min(64, atoi(e)));
    n_threads_ = n;
    for (int i = 0; i < n; ++i) std::thread([this] { loop(); }).detach();
  }
  void loop() {
    unsigned long seen = 0;
    for (;;) {
      // spin briefly for the next call (they arrive back to back in a training loop), then sleep
      bool got = false;
      for (int spin = 0; spin < 20000; ++spin) {
        if (gen_.load(std::memory_order_acquire) != seen) { got = true; break; }
        cpu_relax();
      }
      std::shared_ptr<Job> job;
      {
        std::unique_lock<std::mutex> lk(mu_);
        if (!got) cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen; });
        seen = gen_.load(std::memory_order_acquire);
        job = job_;
      }
      if (job) drain(*job);
    }
  }

  // gather + convert rows [r0, r1) of a view; returns kFlag* bits according to c.check
  template <typename S>
  static int copy_rows(const Chunk& c) {
    const v2v_host_view& v = *c.view;
    const S* src = static_cast<const S*>(v.ptr);
    const bool plain_f32 = std::is_same<S, float>::value && v.col_stride == 1 && c.check == 0;
    if (plain_f32 && v.row_stride == v.cols && v.dst_row_stride == v.cols) {
      stream_copy_f32(c.dst + v.dst_off + c.r0 * v.cols, reinterpret_cast<const float*>(src) + c.r0 * v.cols,
                      (size_t)(c.r1 - c.r0) * v.cols);
      return 0;
    }
    int nonbinary = 0, nonzero = 0;
    const long cs = v.col_stride;
    for (long r = c.r0; r < c.r1; ++r) {
      const S* s = src + r * v.row_stride;
      float* d = c.dst + v.dst_off + r * v.dst_row_stride;
      if (plain_f32) {
        memcpy(d, s, (size_t)v.cols * sizeof(float));
      } else if (c.check == kCheckNonzero) {
        int nz = 0;
        for (long k = 0; k < v.cols; ++k) { const float x = (float)s[k * cs]; d[k] = x; nz |= (x != 0.f); }
        nonzero |= nz;
      } else if (c.check == kCheckBinary) {
        int nb = 0;
        for (long k = 0; k < v.cols; ++k) { const float x = (float)s[k * cs]; d[k] = x; nb |= (x != 0.f) & (x != 1.f); }
        nonbinary |= nb;
      } else if (cs == 1) {
        for (long k = 0; k < v.cols; ++k) d[k] = (float)s[k];
      } else {
        for (long k = 0; k < v.cols; ++k) d[k] = (float)s[k * cs];
      }
    }
    if (c.check == kCheckBinary) return nonbinary ? kFlagNonbinary : 0;
    if (c.check == kCheckNonzero) return nonzero ? kFlagNonzero : 0;
    return 0;
  }

  // adjacency work item: build the bit masks of graphs [r0/N, r1/N) straight from the caller's memory (no dense copy);
  // returns kFlagNonbinary if a value outside {0,1} was seen (the caller then stages the dense matrix as well)
  template <typename S>
  static int pack_rows(const Chunk& c) {
    const v2v_host_view& v = *c.view;
    const S* src = static_cast<const S*>(v.ptr);
    const int N = c.pack_N;
    const long cs = v.col_stride;
    int nonbinary = 0;
    for (long g = c.r0 / N; g < c.r1 / N; ++g) {
      uint32_t rows[32];
      uint32_t* om = c.out_mask + (size_t)g * N;
      for (int n = 0; n < N; ++n) {
        const S* a = src + (g * N + n) * v.row_stride;
        uint32_t w = 0;
        int m = 0;
#if defined(__SSE2__)
        if (std::is_same<S, float>::value && cs == 1) {
          const float* af = reinterpret_cast<const float*>(a);
          const __m128 zero = _mm_setzero_ps(), one = _mm_set1_ps(1.f);
          __m128 bad = _mm_setzero_ps();
          for (; m + 4 <= N; m += 4) {
            const __m128 x = _mm_loadu_ps(af + m);
            const __m128 nz = _mm_cmpneq_ps(x, zero);
            bad = _mm_or_ps(bad, _mm_and_ps(nz, _mm_cmpneq_ps(x, one)));
            w |= (uint32_t)_mm_movemask_ps(nz) << m;
          }
          nonbinary |= _mm_movemask_ps(bad);
        } else if (std::is_same<S, double>::value && cs == 1) {
          const double* ad = reinterpret_cast<const double*>(a);
          const __m128d zero = _mm_setzero_pd(), one = _mm_set1_pd(1.0);
          __m128d bad = _mm_setzero_pd();
          for (; m + 2 <= N; m += 2) {
            const __m128d x = _mm_loadu_pd(ad + m);
            const __m128d nz = _mm_cmpneq_pd(x, zero);
            bad = _mm_or_pd(bad, _mm_and_pd(nz, _mm_cmpneq_pd(x, one)));
            w |= (uint32_t)_mm_movemask_pd(nz) << m;
          }
          nonbinary |= _mm_movemask_pd(bad);
        }
#endif
        for (; m < N; ++m) {
          const float x = (float)a[m * cs];
          w |= (uint32_t)(x != 0.f) << m;
          nonbinary |= (x != 0.f) & (x != 1.f);
        }
        rows[n] = w;
        om[n] = w;
      }
      for (int n = N; n < 32; ++n) rows[n] = 0;
      transpose32(rows);
      uint32_t* im = c.in_mask + (size_t)g * N;
      for (int m = 0; m < N; ++m) im[m] = rows[m];
    }
    return nonbinary ? kFlagNonbinary : 0;
  }
  static int run_chunk(const Chunk& c) {
    if (c.pack_N > 0) return c.view->dtype == V2V_F64 ? pack_rows<double>(c) : pack_rows<float>(c);
    return c.view->dtype == V2V_F64 ? copy_rows<double>(c) : copy_rows<float>(c);
  }

  int n_threads_ = 0;
  std::mutex mu_;
  std::condition_variable cv_;
  std::atomic<unsigned long> gen_{0};
  std::shared_ptr<Job> job_;
};

}  // namespace

int host_stage_threads() { return StagePool::get().threads(); }

int host_stage_validate(const v2v_host_view* views, int n, long dst_elems, const char* what) {
  V2V_REQUIRE(n >= 0 && (n == 0 || views), "%s: null views", what);
  for (int i = 0; i < n; ++i) {
    const v2v_host_view& v = views[i];
    V2V_REQUIRE(v.ptr || v.rows * v.cols == 0, "%s: view %d has a null pointer", what, i);
    V2V_REQUIRE(v.dtype == V2V_F32 || v.dtype == V2V_F64, "%s: view %d dtype %d (host views are V2V_F32 or V2V_F64)", what, i, v.dtype);
    V2V_REQUIRE(v.rows >= 0 && v.cols >= 0 && v.dst_off >= 0 && v.dst_row_stride >= v.cols, "%s: view %d bad shape", what, i);
    if (v.rows > 0 && v.cols > 0)
      V2V_REQUIRE(v.dst_off + (v.rows - 1) * v.dst_row_stride + v.cols <= dst_elems,
                  "%s: view %d writes past the staging tensor (%ld floats)", what, i, dst_elems);
  }
  return 0;
}

int host_stage_run(const HostStageTensor* tensors, int n_tensors, cudaStream_t st, int* flags_out) {
  StagePool& pool = StagePool::get();
  V2V_REQUIRE(n_tensors <= kHostStageMaxTensors, "host_stage_run: too many tensors");
  auto job = std::make_shared<Job>();
  constexpr long kChunkElems = 16384;                 // 64 KB of fp32 output per work item
  for (int t = 0; t < n_tensors; ++t) {
    const HostStageTensor& T = tensors[t];
    int cnt = 0;
    for (int i = 0; i < T.n_views; ++i) {
      const v2v_host_view& v = T.views[i];
      if (v.rows <= 0 || v.cols <= 0) continue;
      long step = std::max<long>(1, kChunkElems / v.cols);
      if (T.pack_N > 0) step = std::max<long>(1, step / T.pack_N) * T.pack_N;      // whole graphs per work item
      for (long r = 0; r < v.rows; r += step) {
        job->chunks.push_back(Chunk{&v, T.pinned, r, std::min(v.rows, r + step), t, T.check, T.pack_N, T.pin_in_mask,
                                    T.pin_out_mask});
        ++cnt;
      }
    }
    job->flags[t].store(0, std::memory_order_relaxed);
    job->remaining[t].store(cnt, std::memory_order_relaxed);
  }
  job->active.store((int)job->chunks.size(), std::memory_order_release);
  if (pool.threads() > 0) pool.submit(job);
  // enqueue each tensor's H2D as soon as it is complete (tensors are listed in the order they should travel)
  int rc = 0;
  auto h2d = [&](void* dst, const void* src, size_t bytes) {
    if (!rc && dst && bytes > 0 && cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess)
      rc = fail("host staging: cudaMemcpyAsync failed: %s", cudaGetErrorString(cudaGetLastError()));
  };
  for (int t = 0; t < n_tensors; ++t) {
    const HostStageTensor& T = tensors[t];
    // the calling thread works too (one item at a time, so a finished tensor is enqueued at most one item late)
    while (job->remaining[t].load(std::memory_order_acquire) > 0)
      if (!StagePool::run_one(*job)) StagePool::cpu_relax();
    const int f = job->flags[t].load(std::memory_order_acquire);
    if (flags_out) flags_out[t] = f;
    if (T.pack_N > 0) {
      h2d(T.dev_in_mask, T.pin_in_mask, T.mask_bytes);
      h2d(T.dev_out_mask, T.pin_out_mask, T.mask_bytes);
      continue;                        // the dense matrix was not staged: see below
    }
    if (T.copy_if == 0 || (f & T.copy_if)) h2d(T.device, T.h2d_from ? T.h2d_from : T.pinned, T.bytes);
  }
  while (job->active.load(std::memory_order_acquire) > 0) StagePool::cpu_relax();
  // weighted adjacency (values outside {0,1}; never produced by the reference, :441-445): the kernels need the dense
  // matrix after all -- stage it now, as a plain tensor
  for (int t = 0; t < n_tensors && !rc; ++t) {
    const HostStageTensor& T = tensors[t];
    if (T.pack_N > 0 && T.pinned && (job->flags[t].load() & kFlagNonbinary)) {
      HostStageTensor D = T;
      D.pack_N = 0; D.check = 0; D.copy_if = 0;
      rc = host_stage_run(&D, 1, st, nullptr);
    }
  }
  return rc;
}

}  // namespace v2v
