// Result transport: the device-resident CSC -> the caller's host arrays (fegpu_makematrix_copy).
//
// The CSC of BASELINE config 2 is 8.27 GB (rowval 4.1 GB + nzval 4.1 GB + colptr), i.e. ~180 ms of PCIe Gen5 time against an
// 11 ms assembly: the device->host link is the end-to-end bottleneck, so the bytes that cross it are minimised.
//   rowval : row indices fit 32 bits (the dof map is int32 on the device), so they cross the link as int32 chunks
//            (k_narrow -> pinned staging ring) and are widened to the caller's Int64 array by a small pool of host threads
//            (AVX2 sign-extension + non-temporal stores) while the next chunks and nzval are in flight.  Measured on the
//            B200 hosts (profiles/r01_xfer_sweep.jsonl): 8.27 GB plain DMA 146-155 ms; this transport 134 ms with 4 threads
//            (more threads compete with the DMA for host memory bandwidth: 16 threads 142-149 ms).  The destination
//            may be pageable memory (a Julia Vector{Int}): the threads write it directly, no driver bounce buffer.
//   nzval  : DMA straight into the destination when it is page-locked; otherwise through the same staging ring with the
//            threads doing the memcpy (faster than the driver's single-threaded pageable path).
//   colptr : one plain copy.
// rowval and nzval chunks are interleaved on the copy stream so widening chunk c overlaps the transfer of nzval chunk c and
// rowval chunk c+1.  If the threads fall behind the link (the next chunk has already landed when they finish one) and the
// destination is page-locked, the next rowval chunk bypasses them as plain int64 DMA: the split balances itself.  This is a transport codec only: every value is produced on the device.
#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>

#include "fegpu_internal.h"

namespace {

constexpr int XF_NBUF = 4;                        // staging ring depth
constexpr size_t XF_CHUNK_MAX = (size_t)32 << 20;  // bytes per staging buffer (device + pinned host)

__global__ void k_narrow(const int64_t *__restrict__ in, int32_t *__restrict__ out, int64_t n) {
  const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < n) {
    const longlong2 a = *reinterpret_cast<const longlong2 *>(in + i4), b = *reinterpret_cast<const longlong2 *>(in + i4 + 2);
    *reinterpret_cast<int4 *>(out + i4) = make_int4((int)a.x, (int)a.y, (int)b.x, (int)b.y);
  } else {
    for (int64_t i = i4; i < n; i++) out[i] = (int32_t)in[i];
  }
}

class HostPool {
 public:
  explicit HostPool(int n) : n_(n) {
    for (int t = 0; t < n_; t++) th_.emplace_back([this, t] { loop(t); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      gen_++;
    }
    cv_.notify_all();
    for (auto &t : th_) t.join();
  }
  int size() const { return n_; }
  // runs fn(tid, nthreads) on every worker and returns when all are done
  void run(const std::function<void(int, int)> &fn) {
    start(fn);
    wait();
  }
  // asynchronous form: fn must stay alive until wait() returns
  void start(const std::function<void(int, int)> &fn) {
    std::lock_guard<std::mutex> lk(mu_);
    fn_ = &fn;
    pending_ = n_;
    gen_++;
    cv_.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void loop(int tid) {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(int, int)> *fn;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        fn = fn_;
      }
      (*fn)(tid, n_);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  int n_;
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int, int)> *fn_ = nullptr;
  uint64_t gen_ = 0;
  int pending_ = 0;
  bool stop_ = false;
};

__attribute__((target("avx2"))) void widen_avx2(const int32_t *src, int64_t *dst, size_t n) {
  size_t i = 0;
  while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31)) { dst[i] = src[i]; i++; }
  for (; i + 8 <= n; i += 8) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 4));
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), _mm256_cvtepi32_epi64(a));
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 4), _mm256_cvtepi32_epi64(b));
  }
  for (; i < n; i++) dst[i] = src[i];
  _mm_sfence();
}

__attribute__((target("avx512f"))) void widen_avx512(const int32_t *src, int64_t *dst, size_t n) {
  size_t i = 0;
  while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 63)) { dst[i] = src[i]; i++; }
  for (; i + 16 <= n; i += 16) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 8));
    _mm512_stream_si512(reinterpret_cast<__m512i *>(dst + i), _mm512_cvtepi32_epi64(a));
    _mm512_stream_si512(reinterpret_cast<__m512i *>(dst + i + 8), _mm512_cvtepi32_epi64(b));
  }
  for (; i < n; i++) dst[i] = src[i];
  _mm_sfence();
}

void widen_scalar(const int32_t *src, int64_t *dst, size_t n) {
  for (size_t i = 0; i < n; i++) dst[i] = src[i];
}

// dst <- src (n int64) with non-temporal stores (full 32-byte stores on the aligned body; callers pass long runs, a short
// run would leave partially filled write-combining buffers, which is far slower than cached stores)
__attribute__((target("avx2"))) void stream_copy_avx2(int64_t *dst, const int64_t *src, size_t n) {
  size_t i = 0;
  while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 63)) { dst[i] = src[i]; i++; }
  for (; i + 8 <= n; i += 8) {
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i)));
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 4), _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 4)));
  }
  for (; i < n; i++) dst[i] = src[i];
}

// Collects the row lists a thread produces while their destinations are consecutive in rowval (they are whenever the
// columns of consecutive nodes are consecutive, i.e. for node-major numberings) and writes them out as long streams.
struct RowSink {
  static constexpr size_t CAP = 16384;  // int64 entries (128 KB, L2 resident)
  std::vector<int64_t> buf;
  size_t n = 0;
  int64_t *start = nullptr;
  bool avx2;
  explicit RowSink(bool a) : buf(CAP), avx2(a) {}
  void flush() {
    if (!n) return;
    if (avx2 && n >= 64) stream_copy_avx2(start, buf.data(), n);
    else std::memcpy(start, buf.data(), n * sizeof(int64_t));
    n = 0;
  }
  void put(int64_t *dst, const int64_t *src, size_t len) {
    if (len > CAP) { flush(); std::memcpy(dst, src, len * sizeof(int64_t)); return; }
    if (n && (dst != start + n || n + len > CAP)) flush();
    if (!n) start = dst;
    std::memcpy(buf.data() + n, src, len * sizeof(int64_t));
    n += len;
  }
};

// Row indices of the columns of nodes [n_lo, n_hi) from the compressed pattern: the rows of every column of node n are the
// dofs of its neighbour nodes in (neighbour ascending, component ascending) order -- the same list for the node's ndn
// columns.  A decoder of what the device's symbolic phase produced (k_nbr / k_rows_sorted), not a pattern computation: no
// connectivity is touched on the host.
void expand_rows(const int32_t *nbr, const int64_t *nbrptr, const int32_t *dof, int ndn, int64_t nnodes, const int64_t *colptr, int64_t *rowval,
                 int64_t n_lo, int64_t n_hi, bool avx2) {
  std::vector<int64_t> rows;
  RowSink sink(avx2);
  for (int64_t n = n_lo; n < n_hi; n++) {
    const int64_t b = nbrptr[n], nu = nbrptr[n + 1] - b;
    if (nu == 0) continue;
    const size_t nr = (size_t)nu * ndn;
    if (rows.size() < nr) rows.resize(nr);
    int64_t *r = rows.data();
    if (ndn == 3) {
      const int32_t *d0 = dof, *d1 = dof + nnodes, *d2 = dof + 2 * nnodes;
      for (int64_t s = 0; s < nu; s++) {
        const int32_t m = nbr[b + s];
        r[3 * s] = (int64_t)d0[m] + 1;
        r[3 * s + 1] = (int64_t)d1[m] + 1;
        r[3 * s + 2] = (int64_t)d2[m] + 1;
      }
    } else {
      for (int64_t s = 0; s < nu; s++) {
        const int32_t m = nbr[b + s];
        for (int p = 0; p < ndn; p++) r[s * ndn + p] = (int64_t)dof[(int64_t)p * nnodes + m] + 1;
      }
    }
    for (int q = 0; q < ndn; q++) sink.put(rowval + (colptr[dof[(int64_t)q * nnodes + n]] - 1), r, nr);
  }
  sink.flush();
  _mm_sfence();
}

// Row indices of columns [c_lo, c_hi) from the column-stencil codec (fe_col_stencils, fegpu_csc_ops.cu): the rows of column c are
// c + 1 + offsets of the dictionary entry its id names.  A decoder of what the device produced and verified column by column.
// ids and colptr are the slices of the matrix's non-empty column window, which starts at column c0 (0-based).
void expand_stencils(const uint32_t *ids, const int32_t *lut, const int32_t *dict, int stride, const int64_t *colptr, int64_t *rowval, int64_t c_lo,
                     int64_t c_hi, int64_t c0, bool avx2) {
  int64_t rows[128];
  RowSink sink(avx2);
  for (int64_t c = c_lo; c < c_hi; c++) {
    const int32_t *d = dict + (size_t)lut[ids[c]] * stride;
    const int len = d[1];
    if (len == 0) continue;
    const int64_t c1 = c0 + c + 1;
    for (int k = 0; k < len; k++) rows[k] = c1 + d[2 + k];
    sink.put(rowval + (colptr[c] - 1), rows, (size_t)len);
  }
  sink.flush();
  _mm_sfence();
}

}  // namespace

struct Transfer {
  cudaStream_t stream = nullptr;
  cudaEvent_t ready = nullptr;             // results of the compute stream are complete
  cudaEvent_t done[XF_NBUF] = {};          // staging buffer b has landed in pinned memory
  void *d_stage[XF_NBUF] = {};
  void *h_stage[XF_NBUF] = {};
  HostPool *pool = nullptr;
  int simd = 0;                            // 0 scalar, 1 AVX2, 2 AVX-512
  size_t chunk_bytes = XF_CHUNK_MAX;       // FEGPU_XFER_CHUNK_MB (tuning knob, <= 32)
  bool narrow = true;                      // FEGPU_XFER_NARROW=0: plain int64 DMA (A/B measurements)
  int widen_threads = 4;                   // threads of the pool that widen / copy staged chunks
  bool compress = true;                    // FEGPU_XFER_COMPRESS=0: never send neighbour lists instead of row indices
  int64_t staged = 0, bypassed = 0;        // chunk counters (diagnostics)
  int64_t compressed = 0;                  // results whose row indices were rebuilt from neighbour lists
  bool stencil = true;                     // FEGPU_XFER_STENCIL=0: never send column-stencil ids instead of row indices
  int64_t stenciled = 0;                   // results whose row indices were rebuilt from column stencils
  void *h_meta = nullptr;                  // pinned: neighbour lists, their offsets, the dof map, colptr
  size_t meta_cap = 0;
  cudaEvent_t meta_done = nullptr;
  ~Transfer() {
    delete pool;
    if (h_meta) cudaFreeHost(h_meta);
    if (meta_done) cudaEventDestroy(meta_done);
    for (int b = 0; b < XF_NBUF; b++) {
      if (d_stage[b]) cudaFree(d_stage[b]);
      if (h_stage[b]) cudaFreeHost(h_stage[b]);
      if (done[b]) cudaEventDestroy(done[b]);
    }
    if (ready) cudaEventDestroy(ready);
    if (stream) cudaStreamDestroy(stream);
  }
};

void fe_transfer_free(Transfer *t) { delete t; }

static int32_t transfer_get(fegpu_ctx *ctx, Transfer **out) {
  if (ctx->xfer) { *out = ctx->xfer; return FEGPU_OK; }
  Transfer *t = new Transfer();
  ctx->xfer = t;  // owned by the context from here on (also on error paths)
  CUDA_TRY(ctx, cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking));
  CUDA_TRY(ctx, cudaEventCreateWithFlags(&t->ready, cudaEventDisableTiming));
  CUDA_TRY(ctx, cudaEventCreateWithFlags(&t->meta_done, cudaEventDisableTiming));
  for (int b = 0; b < XF_NBUF; b++) {
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&t->done[b], cudaEventDisableTiming));
    CUDA_TRY(ctx, cudaMalloc(&t->d_stage[b], XF_CHUNK_MAX));
    CUDA_TRY(ctx, cudaHostAlloc(&t->h_stage[b], XF_CHUNK_MAX, cudaHostAllocDefault));
  }
  int nt = (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  // measured on the B200 hosts (profiles/r01_xfer_sweep.jsonl, r01_xfer_compress.jsonl): widening int32 chunks is fastest with 4
  // threads (more only steal memory bandwidth from the DMA); decoding row indices from neighbour lists (non-temporal streams) is at the nzval DMA time with 4-8
  nt = std::min(nt, 8);
  // one process per GPU on a shared host (torchrun exports LOCAL_WORLD_SIZE): the ranks' pools together must not oversubscribe
  // the cores -- 8 ranks x 8 threads on a 32-core box slowed every rank's decode down (round-1 SCALE run)
  if (const char *e = std::getenv("LOCAL_WORLD_SIZE")) {
    const int lw = std::max(1, std::atoi(e));
    nt = std::max(2, std::min(nt, (int)std::thread::hardware_concurrency() / lw));
  }
  if (const char *e = std::getenv("FEGPU_HOST_THREADS")) nt = std::max(1, std::atoi(e));
  t->widen_threads = std::min(nt, 4);
  if (const char *e = std::getenv("FEGPU_WIDEN_THREADS")) t->widen_threads = std::max(1, std::min(nt, std::atoi(e)));
  t->pool = new HostPool(nt);
  __builtin_cpu_init();
  t->simd = __builtin_cpu_supports("avx2") ? 1 : 0;  // AVX-512 (FEGPU_XFER_SIMD=2) measured no faster: the loop is memory-bound
  if (std::getenv("FEGPU_XFER_SIMD") && std::atoi(std::getenv("FEGPU_XFER_SIMD")) >= 2 && __builtin_cpu_supports("avx512f")) t->simd = 2;
  else if (const char *e = std::getenv("FEGPU_XFER_SIMD")) t->simd = std::min(t->simd, std::max(0, std::atoi(e)));
  if (const char *e = std::getenv("FEGPU_XFER_CHUNK_MB")) t->chunk_bytes = std::min(XF_CHUNK_MAX, (size_t)std::max(1, std::atoi(e)) << 20);
  if (const char *e = std::getenv("FEGPU_XFER_NARROW")) t->narrow = std::atoi(e) != 0;
  if (const char *e = std::getenv("FEGPU_XFER_COMPRESS")) t->compress = std::atoi(e) != 0;
  if (const char *e = std::getenv("FEGPU_XFER_STENCIL")) t->stencil = std::atoi(e) != 0;
  *out = t;
  return FEGPU_OK;
}

static bool is_pinned(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

// One stream of `n` items: device source -> (optional narrowing) -> pinned ring -> host threads -> destination.
struct StagedJob {
  const void *d_src = nullptr;  // int64 (narrow) or raw bytes
  void *h_dst = nullptr;
  int64_t n = 0;                // items
  bool narrow = false;          // int64 -> int32 on the device, widened back on the host
  bool dst_pinned = false;      // a lagging host may be bypassed by plain DMA of the 8-byte items
  size_t item_dev = 0;          // bytes per item in the staging buffers
  int64_t per_chunk = 0;
  int64_t next = 0;             // first item not yet issued
};

int32_t fe_copy_result(fegpu_asm *as, int64_t *colptr, int64_t *rowval, double *nzval) {
  fegpu_ctx *ctx = as->ctx;
  Transfer *T = nullptr;
  FE_TRY(transfer_get(ctx, &T));
  cudaStream_t cs = T->stream;
  CUDA_TRY(ctx, cudaEventRecord(T->ready, ctx->stream));
  const int64_t nnz = as->r_nnz();
  bool joined = false;  // has the copy stream been ordered after everything queued on the caller's stream?
  auto join = [&]() -> int32_t {
    if (!joined) CUDA_TRY(ctx, cudaStreamWaitEvent(cs, T->ready, 0));
    joined = true;
    return FEGPU_OK;
  };

  StagedJob jobs[2];
  int njobs = 0;
  const double *direct_nz = nullptr;
  const int64_t *direct_rv = nullptr;

  // Row indices of a vector-field pattern: ndn^2 entries of rowval per (column node, neighbour node) pair, so the pair list
  // (int32 per pair) + the dof map cross the link instead -- config 2: 0.27 GB instead of 2.05 GB -- and the host threads
  // rebuild rowval while nzval is in flight.
  const int32_t *c_nbr = nullptr, *c_dof = nullptr;
  const int64_t *c_nbrptr = nullptr;
  int64_t c_total = 0, c_nnodes = 0;
  int c_ndn = 0;
  bool compressed = rowval && nnz && T->compress && !as->view.active && as->pat_src &&
                    fe_pattern_compressed(as->pat_src, &c_nbr, &c_nbrptr, &c_total, &c_dof, &c_ndn, &c_nnodes);
  const int32_t *h_nbr = nullptr, *h_dof = nullptr;
  const int64_t *h_nbrptr = nullptr, *h_colptr = nullptr;
  if (compressed) {
    const size_t b_nbr = ((size_t)c_total * 4 + 63) & ~(size_t)63, b_ptr = ((size_t)(c_nnodes + 1) * 8 + 63) & ~(size_t)63;
    const size_t b_dof = ((size_t)c_nnodes * c_ndn * 4 + 63) & ~(size_t)63, b_col = ((size_t)(as->ncols + 1) * 8 + 63) & ~(size_t)63;
    const size_t need = b_nbr + b_ptr + b_dof + b_col;
    if (T->meta_cap < need) {
      if (T->h_meta) cudaFreeHost(T->h_meta);
      T->h_meta = nullptr; T->meta_cap = 0;
      CUDA_TRY(ctx, cudaHostAlloc(&T->h_meta, need, cudaHostAllocDefault));
      T->meta_cap = need;
    }
    char *hm = static_cast<char *>(T->h_meta);
    // The pattern's arrays are final when its build is (the event of the pattern): when the form call was asynchronous
    // (fegpu_set_async) they cross the link while the integration and the numeric phase are still running, and the host
    // threads can start on rowval before nzval exists.
    static const bool early_off = std::getenv("FEGPU_EARLY_META") && std::atoi(std::getenv("FEGPU_EARLY_META")) == 0;  // A/B knob
    cudaEvent_t pr = early_off ? nullptr : fe_pattern_ready_event(as->pat_src);
    if (pr) CUDA_TRY(ctx, cudaStreamWaitEvent(cs, pr, 0));
    else FE_TRY(join());
    // small pieces first so the threads can start on the first nodes as early as possible
    CUDA_TRY(ctx, cudaMemcpyAsync(hm, as->d_colptr, (size_t)(as->ncols + 1) * 8, cudaMemcpyDeviceToHost, cs));
    CUDA_TRY(ctx, cudaMemcpyAsync(hm + b_col, c_nbrptr, (size_t)(c_nnodes + 1) * 8, cudaMemcpyDeviceToHost, cs));
    CUDA_TRY(ctx, cudaMemcpyAsync(hm + b_col + b_ptr, c_dof, (size_t)c_nnodes * c_ndn * 4, cudaMemcpyDeviceToHost, cs));
    CUDA_TRY(ctx, cudaMemcpyAsync(hm + b_col + b_ptr + b_dof, c_nbr, (size_t)c_total * 4, cudaMemcpyDeviceToHost, cs));
    CUDA_TRY(ctx, cudaEventRecord(T->meta_done, cs));
    FE_TRY(join());
    h_colptr = reinterpret_cast<const int64_t *>(hm);
    h_nbrptr = reinterpret_cast<const int64_t *>(hm + b_col);
    h_dof = reinterpret_cast<const int32_t *>(hm + b_col + b_ptr);
    h_nbr = reinterpret_cast<const int32_t *>(hm + b_col + b_ptr + b_dof);
    T->compressed++;
  }
  // Scalar fields and every other result whose row indices would cross as int32: one id per column + a dictionary of row-offset
  // lists (column stencils) instead -- config 4: 68 MB + a few KB instead of 1.82 GB.  The device builds and verifies the codec on
  // the copy stream (one round trip); a matrix without a small dictionary keeps the int32 path.
  const uint32_t *h_ids = nullptr;
  const int32_t *h_dict = nullptr;
  std::vector<int32_t> st_lut;
  int st_stride = 0, st_nd = 0;
  int64_t st_c0 = 0, st_nc = 0;  // the matrix's non-empty column window [st_c0, st_c0 + st_nc): only its colptr and ids cross the link
  bool stencil = false;
  if (!compressed && rowval && nnz >= ((int64_t)1 << 20) && T->stencil && !as->view.active) {
    cudaEvent_t pr = as->pat_src ? fe_pattern_ready_event(as->pat_src) : nullptr;
    if (pr) CUDA_TRY(ctx, cudaStreamWaitEvent(cs, pr, 0));
    else FE_TRY(join());
    uint32_t *d_ids = nullptr;
    int32_t *d_dict = nullptr;
    int nd = 0, maxlen = 0, cap = 0;
    int64_t cfirst = 0, clast = -1;
    FE_TRY(fe_col_stencils(ctx, as->ncols, as->r_colptr(), as->r_rowval(), cs, &d_ids, &d_dict, &nd, &maxlen, &cap, &cfirst, &clast, &stencil));
    if (stencil) {
      st_stride = 2 + maxlen;
      st_nd = nd;
      st_c0 = cfirst;
      st_nc = clast - cfirst + 1;
      const size_t b_col = ((size_t)(st_nc + 1) * 8 + 63) & ~(size_t)63, b_ids = ((size_t)st_nc * 4 + 63) & ~(size_t)63;
      const size_t b_dict = ((size_t)nd * st_stride * 4 + 63) & ~(size_t)63;
      const size_t need = b_col + b_ids + b_dict;
      if (T->meta_cap < need) {
        if (T->h_meta) cudaFreeHost(T->h_meta);
        T->h_meta = nullptr; T->meta_cap = 0;
        CUDA_TRY(ctx, cudaHostAlloc(&T->h_meta, need, cudaHostAllocDefault));
        T->meta_cap = need;
      }
      char *hm = static_cast<char *>(T->h_meta);
      CUDA_TRY(ctx, cudaMemcpyAsync(hm + b_col + b_ids, d_dict, (size_t)nd * st_stride * 4, cudaMemcpyDeviceToHost, cs));
      CUDA_TRY(ctx, cudaMemcpyAsync(hm, as->r_colptr() + st_c0, (size_t)(st_nc + 1) * 8, cudaMemcpyDeviceToHost, cs));
      CUDA_TRY(ctx, cudaMemcpyAsync(hm + b_col, d_ids + st_c0, (size_t)st_nc * 4, cudaMemcpyDeviceToHost, cs));
      CUDA_TRY(ctx, cudaEventRecord(T->meta_done, cs));
      fe_dev_free(ctx, d_ids, cs);
      fe_dev_free(ctx, d_dict, cs);
      h_colptr = reinterpret_cast<const int64_t *>(hm);
      h_ids = reinterpret_cast<const uint32_t *>(hm + b_col);
      h_dict = reinterpret_cast<const int32_t *>(hm + b_col + b_ids);
      st_lut.assign((size_t)cap, 0);
      c_nnodes = st_nc;  // the slices below run over the columns of the window
      c_total = nnz;
      T->stenciled++;
    }
  }
  const bool rebuilt = compressed || stencil;  // rowval is produced by the host threads, not shipped
  FE_TRY(join());
  // node (or column) range [lo, hi) of slice k of K, balanced by the number of row indices
  auto node_slice = [&](int64_t k, int64_t K, int64_t *lo, int64_t *hi) {
    auto cut = [&](int64_t j) -> int64_t {
      if (j <= 0) return 0;
      if (j >= K) return c_nnodes;
      const int64_t target = (int64_t)((__int128)c_total * j / K);
      if (stencil) return std::upper_bound(h_colptr, h_colptr + c_nnodes + 1, target + 1) - h_colptr - 1;
      return std::upper_bound(h_nbrptr, h_nbrptr + c_nnodes + 1, target) - h_nbrptr - 1;
    };
    *lo = cut(k); *hi = cut(k + 1);
  };
  auto expand_slice = [&](int64_t lo, int64_t hi) {
    if (stencil) expand_stencils(h_ids, st_lut.data(), h_dict, st_stride, h_colptr, rowval, lo, hi, st_c0, T->simd >= 1);
    else expand_rows(h_nbr, h_nbrptr, h_dof, c_ndn, c_nnodes, h_colptr, rowval, lo, hi, T->simd >= 1);
  };
  // share `tid` of `nth` of the caller's colptr.  Column stencils: only the window's colptr crossed the link; ahead of it every
  // column starts at 1, behind it at nnz + 1
  auto fill_colptr = [&](int tid, int nth) {
    const int64_t n = as->ncols + 1, per = (n + nth - 1) / nth, a = std::min(n, per * tid), b = std::min(n, a + per);
    if (b <= a) return;
    if (!stencil) { std::memcpy(colptr + a, h_colptr + a, (size_t)(b - a) * 8); return; }
    for (int64_t c = a; c < std::min(b, st_c0); c++) colptr[c] = 1;
    const int64_t wa = std::max(a, st_c0), wb = std::min(b, st_c0 + st_nc + 1);
    if (wb > wa) std::memcpy(colptr + wa, h_colptr + (wa - st_c0), (size_t)(wb - wa) * 8);
    for (int64_t c = std::max(a, st_c0 + st_nc + 1); c < b; c++) colptr[c] = nnz + 1;
  };
  auto meta_arrived = [&]() -> int32_t {  // the host may read the metadata from here on
    CUDA_TRY(ctx, cudaEventSynchronize(T->meta_done));
    if (stencil)
      for (int k = 0; k < st_nd; k++) st_lut[(size_t)h_dict[(size_t)k * st_stride]] = k;  // slot number -> dictionary entry
    return FEGPU_OK;
  };
  int64_t exp_done = 0, exp_total = 0;  // node slices of the expansion handed to the threads so far / in all

  if (rowval && nnz && !rebuilt) {
    if (!T->narrow && is_pinned(rowval)) {
      direct_rv = as->r_rowval();
    } else {
      StagedJob &j = jobs[njobs++];
      j.d_src = as->r_rowval(); j.h_dst = rowval; j.n = nnz; j.narrow = true; j.item_dev = 4; j.dst_pinned = is_pinned(rowval);
    }
  }
  if (nzval && nnz) {
    if (is_pinned(nzval)) {
      direct_nz = as->r_nzval();
    } else {
      StagedJob &j = jobs[njobs++];
      j.d_src = as->r_nzval(); j.h_dst = nzval; j.n = nnz; j.narrow = false; j.item_dev = 8;
    }
  }
  int64_t total_chunks = 0;
  for (int k = 0; k < njobs; k++) {
    jobs[k].per_chunk = (int64_t)(T->chunk_bytes / jobs[k].item_dev);
    total_chunks += (jobs[k].n + jobs[k].per_chunk - 1) / jobs[k].per_chunk;
  }
  if (colptr && !rebuilt) CUDA_TRY(ctx, cudaMemcpyAsync(colptr, as->r_colptr(), sizeof(int64_t) * (as->r_ncols() + 1), cudaMemcpyDeviceToHost, cs));
  if (direct_rv) CUDA_TRY(ctx, cudaMemcpyAsync(rowval, direct_rv, sizeof(int64_t) * nnz, cudaMemcpyDeviceToHost, cs));

  // page-locked nzval goes by plain DMA, sliced in between the staged chunks so the link never idles while the threads work
  int64_t nz_issued = 0;
  const int64_t nz_slice = direct_nz ? (total_chunks ? (nnz + total_chunks - 1) / total_chunks : nnz) : 0;
  auto issue_direct_nz = [&]() -> int32_t {
    if (!direct_nz || nz_issued >= nnz) return FEGPU_OK;
    const int64_t len = std::min(nz_slice, nnz - nz_issued);
    CUDA_TRY(ctx, cudaMemcpyAsync(nzval + nz_issued, direct_nz + nz_issued, sizeof(double) * len, cudaMemcpyDeviceToHost, cs));
    nz_issued += len;
    return FEGPU_OK;
  };

  std::function<void(int, int)> expand_all;
  if (rebuilt) {
    exp_total = std::max<int64_t>(total_chunks, 1);
    if (total_chunks == 0) {
      // nothing is staged (nzval page-locked or not requested): the link carries nzval by plain DMA while the threads expand
      while (direct_nz && nz_issued < nnz) FE_TRY(issue_direct_nz());
      FE_TRY(meta_arrived());
      expand_all = [&](int tid, int nth) {
        int64_t lo, hi;
        node_slice(tid, nth, &lo, &hi);
        expand_slice(lo, hi);
        if (colptr) fill_colptr(tid, nth);  // the caller's colptr: each thread writes its share
      };
      T->pool->run(expand_all);
      exp_done = exp_total;
    } else {
      FE_TRY(meta_arrived());
      if (colptr) fill_colptr(0, 1);
    }
  }

  struct Pending { int buf, job; int64_t off, len; };
  Pending ring[XF_NBUF];
  int head = 0, npend = 0;          // FIFO of staged chunks in flight
  int free_buf[XF_NBUF], nfree = XF_NBUF;
  for (int b = 0; b < XF_NBUF; b++) free_buf[b] = b;
  int rr = 0;                       // round-robin over the jobs
  bool lagging = false;             // the next staged chunk had already landed when the threads finished the previous one
  auto remaining = [&]() { for (int k = 0; k < njobs; k++) if (jobs[k].next < jobs[k].n) return true; return false; };

  while (remaining() || npend > 0) {
    int bypass = lagging ? 1 : 0;  // at most one bypassed chunk per consumed chunk
    while (remaining() && (npend < XF_NBUF - 1 || bypass > 0)) {
      while (jobs[rr].next >= jobs[rr].n) rr = (rr + 1) % njobs;
      StagedJob &j = jobs[rr];
      const int k = rr;
      rr = (rr + 1) % njobs;
      const int64_t off = j.next, len = std::min(j.per_chunk, j.n - off);
      if (bypass > 0 && j.narrow && j.dst_pinned) {
        // the host threads are the bottleneck right now: this chunk crosses the link as int64, straight to its destination
        CUDA_TRY(ctx, cudaMemcpyAsync(static_cast<int64_t *>(j.h_dst) + off, static_cast<const int64_t *>(j.d_src) + off, sizeof(int64_t) * len,
                                      cudaMemcpyDeviceToHost, cs));
        bypass--;
        T->bypassed++;
      } else {
        if (npend >= XF_NBUF - 1) break;
        const int b = free_buf[--nfree];
        if (j.narrow) {
          k_narrow<<<grid_for((len + 3) / 4, 256), 256, 0, cs>>>(static_cast<const int64_t *>(j.d_src) + off, static_cast<int32_t *>(T->d_stage[b]), len);
          ctx->launches++;
          CUDA_TRY(ctx, cudaMemcpyAsync(T->h_stage[b], T->d_stage[b], (size_t)len * 4, cudaMemcpyDeviceToHost, cs));
        } else {
          CUDA_TRY(ctx, cudaMemcpyAsync(T->h_stage[b], static_cast<const char *>(j.d_src) + (size_t)off * j.item_dev, (size_t)len * j.item_dev,
                                        cudaMemcpyDeviceToHost, cs));
        }
        CUDA_TRY(ctx, cudaEventRecord(T->done[b], cs));
        ring[(head + npend) % XF_NBUF] = Pending{b, k, off, len};
        npend++;
        T->staged++;
      }
      j.next = off + len;
      FE_TRY(issue_direct_nz());
    }
    if (npend == 0) continue;
    const Pending p = ring[head];
    head = (head + 1) % XF_NBUF;
    npend--;
    CUDA_TRY(ctx, cudaEventSynchronize(T->done[p.buf]));
    const StagedJob &j = jobs[p.job];
    const void *src = T->h_stage[p.buf];
    const int simd = T->simd;
    const int64_t exp_k = (rebuilt && exp_done < exp_total) ? exp_done++ : -1;
    const int wth = T->widen_threads;
    std::function<void(int, int)> fn = [&, src, p, simd, exp_k, wth](int tid, int nth) {
      if (exp_k >= 0) {  // this chunk's share of the row-index expansion
        int64_t lo, hi;
        node_slice(exp_k * nth + tid, exp_total * nth, &lo, &hi);
        expand_slice(lo, hi);
      }
      // slices are multiples of 16 items so the vector loops stay aligned
      if (tid >= wth) return;
      nth = wth;
      const int64_t per = (((p.len + nth - 1) / nth) + 15) & ~(int64_t)15;
      const int64_t lo = std::min<int64_t>(p.len, per * tid), hi = std::min<int64_t>(p.len, lo + per);
      if (hi <= lo) return;
      if (j.narrow) {
        const int32_t *s = static_cast<const int32_t *>(src) + lo;
        int64_t *d = static_cast<int64_t *>(j.h_dst) + p.off + lo;
        if (simd == 2) widen_avx512(s, d, (size_t)(hi - lo));
        else if (simd == 1) widen_avx2(s, d, (size_t)(hi - lo));
        else widen_scalar(s, d, (size_t)(hi - lo));
      } else {
        std::memcpy(static_cast<char *>(j.h_dst) + (size_t)(p.off + lo) * 8, static_cast<const char *>(src) + (size_t)lo * 8, (size_t)(hi - lo) * 8);
      }
    };
    T->pool->run(fn);
    free_buf[nfree++] = p.buf;
    lagging = npend > 0 && cudaEventQuery(T->done[ring[head].buf]) == cudaSuccess;
  }
  while (direct_nz && nz_issued < nnz) FE_TRY(issue_direct_nz());
  CUDA_TRY(ctx, cudaStreamSynchronize(cs));
  return FEGPU_OK;
}

extern "C" int32_t fegpu_transfer_stats(fegpu_ctx *ctx, int64_t *staged, int64_t *bypassed) {
  if (!ctx) return FEGPU_ERR_ARG;
  if (staged) *staged = ctx->xfer ? ctx->xfer->staged : 0;
  if (bypassed) *bypassed = ctx->xfer ? ctx->xfer->bypassed : 0;
  return FEGPU_OK;
}

extern "C" int32_t fegpu_transfer_compressed(fegpu_ctx *ctx, int64_t *results) {
  if (!ctx || !results) return FEGPU_ERR_ARG;
  *results = ctx->xfer ? ctx->xfer->compressed : 0;
  return FEGPU_OK;
}

extern "C" int32_t fegpu_transfer_stenciled(fegpu_ctx *ctx, int64_t *results) {
  if (!ctx || !results) return FEGPU_ERR_ARG;
  *results = ctx->xfer ? ctx->xfer->stenciled : 0;
  return FEGPU_OK;
}
