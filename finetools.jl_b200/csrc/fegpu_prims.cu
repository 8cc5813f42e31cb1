// Device-wide primitives: exclusive scan (int64), max reduction.  Three-kernel scan: per-tile scan + tile totals,
// recursive scan of the totals, offset add.  These run in the symbolic phase and the generic sort path.
#include "fegpu_internal.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>

bool fe_trace_on() {
  static const bool on = std::getenv("FEGPU_TRACE") && std::atoi(std::getenv("FEGPU_TRACE")) != 0;
  return on;
}

void fe_trace(const char *label) {
  if (!fe_trace_on()) return;
  static auto last = std::chrono::steady_clock::now();
  const auto now = std::chrono::steady_clock::now();
  std::fprintf(stderr, "[fegpu trace] %10.1f us  %s\n", std::chrono::duration<double, std::micro>(now - last).count(), label);
  last = std::chrono::steady_clock::now();
}

void fe_mark(fegpu_ctx *ctx, const char *name) {
  if (!ctx->marks_on) return;
  if (ctx->nmarks >= (int)ctx->marks.size()) {
    fegpu_ctx::Mark m{name, nullptr};
    if (cudaEventCreate(&m.ev) != cudaSuccess) return;
    ctx->marks.push_back(m);
  }
  ctx->marks[ctx->nmarks].name = name;
  if (cudaEventRecord(ctx->marks[ctx->nmarks].ev, ctx->stream) == cudaSuccess) ctx->nmarks++;
}

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int64_t warp_incl_scan(int64_t v) {
  const unsigned lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int64_t t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= (unsigned)d) v += t;
  }
  return v;
}

// Each block scans one tile: out[i] = exclusive prefix inside the tile; tile_sum[block] = total of the tile.
template <typename TIN>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const TIN *__restrict__ in, int64_t *__restrict__ out, int64_t n,
                                                             int64_t *__restrict__ tile_sum) {
  __shared__ int64_t warp_tot[SCAN_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int64_t v[SCAN_ITEMS];
  int64_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    int64_t idx = base + i;
    v[i] = (idx < n) ? (int64_t)in[idx] : 0;
    s += v[i];
  }
  int64_t incl = warp_incl_scan(s);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 31) warp_tot[w] = incl;
  __syncthreads();
  int64_t woff = 0;
#pragma unroll
  for (int k = 0; k < SCAN_THREADS / 32; k++)
    if (k < w) woff += warp_tot[k];
  int64_t run = woff + incl - s;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    int64_t idx = base + i;
    if (idx < n) out[idx] = run;
    run += v[i];
  }
  if (threadIdx.x == SCAN_THREADS - 1) tile_sum[blockIdx.x] = run;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(int64_t *__restrict__ out, int64_t n, const int64_t *__restrict__ tile_off,
                                                           int64_t base, int64_t *total_slot) {
  const int64_t off = tile_off[blockIdx.x] + base;
  const int64_t b = (int64_t)blockIdx.x * SCAN_TILE;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    int64_t idx = b + (int64_t)i * SCAN_THREADS + threadIdx.x;
    if (idx < n) out[idx] += off;
  }
  if (total_slot && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
    // total = offset of the (virtual) tile after the last one, stored by the caller in tile_off[gridDim.x]
    *total_slot = tile_off[gridDim.x] + base;
  }
}

// single block: exclusive scan of up to any length (loops), writes n+1 entries (out[n] = total)
__global__ void __launch_bounds__(1024) k_scan_small(const int64_t *__restrict__ in, int64_t *__restrict__ out, int64_t n) {
  __shared__ int64_t warp_tot[32];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t start = 0; start < n; start += 1024) {
    int64_t idx = start + threadIdx.x;
    int64_t v = (idx < n) ? in[idx] : 0;
    int64_t incl = warp_incl_scan(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    int64_t woff = 0;
    for (int k = 0; k < w; k++) woff += warp_tot[k];
    int64_t c = carry;
    if (idx < n) out[idx] = c + woff + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + woff + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = carry;
}

template <typename TIN>
int32_t scan_impl(fegpu_ctx *ctx, const TIN *d_in, int64_t *d_out, int64_t n, int64_t base, bool write_total, int64_t *total_host) {
  cudaStream_t st = ctx->stream;
  if (n <= 0) {
    if (write_total) CUDA_TRY(ctx, cudaMemcpyAsync(d_out, &base, sizeof(int64_t), cudaMemcpyHostToDevice, st));
    if (total_host) *total_host = base;
    if (write_total) CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return FEGPU_OK;
  }
  const int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  int64_t *d_tiles = nullptr;  // ntiles sums, then scanned in place into ntiles+1 offsets
  FE_TRY(fe_dev_alloc(ctx, (void **)&d_tiles, sizeof(int64_t) * (size_t)(2 * ntiles + 2), st));
  int64_t *d_sums = d_tiles, *d_offs = d_tiles + ntiles;  // offs has ntiles+1 entries
  k_scan_tiles<TIN><<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(d_in, d_out, n, d_sums);
  ctx->launches++;
  if (ntiles <= 65536) {
    k_scan_small<<<1, 1024, 0, st>>>(d_sums, d_offs, ntiles);
    ctx->launches++;
  } else {
    FE_TRY(scan_impl<int64_t>(ctx, d_sums, d_offs, ntiles, 0, true, nullptr));
  }
  int64_t *total_slot = write_total ? (d_out + n) : nullptr;
  k_scan_add<<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(d_out, n, d_offs, base, total_slot);
  ctx->launches++;
  if (total_host) {
    int64_t t = 0;
    CUDA_TRY(ctx, cudaMemcpyAsync(&t, d_offs + ntiles, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    *total_host = t + base;
  }
  fe_dev_free(ctx, d_tiles, st);
  CUDA_TRY(ctx, cudaGetLastError());
  return FEGPU_OK;
}

__global__ void k_max_i32(const int32_t *__restrict__ in, int64_t n, int32_t *out) {
  int32_t m = INT32_MIN;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = max(m, in[i]);
  for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

}  // namespace

int32_t fe_exclusive_scan_i64(fegpu_ctx *ctx, const int64_t *d_in, int64_t *d_out, int64_t n, int64_t base, bool write_total,
                              int64_t *total_host) {
  return scan_impl<int64_t>(ctx, d_in, d_out, n, base, write_total, total_host);
}
int32_t fe_exclusive_scan_i32_to_i64(fegpu_ctx *ctx, const int32_t *d_in, int64_t *d_out, int64_t n, int64_t base, bool write_total,
                                     int64_t *total_host) {
  return scan_impl<int32_t>(ctx, d_in, d_out, n, base, write_total, total_host);
}

// max(*d_out, in[0..n)) left on the device: no synchronisation (the caller initialises the slot and reads it back with its
// other scalars)
int32_t fe_max_i32_dev(fegpu_ctx *ctx, const int32_t *d_in, int64_t n, int32_t *d_out) {
  if (n > 0) {
    unsigned g = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8);
    k_max_i32<<<g, 256, 0, ctx->stream>>>(d_in, n, d_out);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
  }
  return FEGPU_OK;
}

int32_t fe_max_i32(fegpu_ctx *ctx, const int32_t *d_in, int64_t n, int32_t *max_host) {
  int32_t *d_m = nullptr;
  int32_t init = INT32_MIN;
  FE_TRY(fe_dev_alloc(ctx, (void **)&d_m, sizeof(int32_t), ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(d_m, &init, sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  if (n > 0) {
    unsigned g = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8);
    k_max_i32<<<g, 256, 0, ctx->stream>>>(d_in, n, d_m);
    ctx->launches++;
  }
  CUDA_TRY(ctx, cudaMemcpyAsync(max_host, d_m, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  fe_dev_free(ctx, d_m, ctx->stream);
  return FEGPU_OK;
}
