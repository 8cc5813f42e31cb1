// Device side of the OPTIONAL gather of row-block CSCs into one matrix (the reference's makematrix! returns one
// SparseMatrixCSC, AssemblyModule.jl:319-325).  Assembly itself needs no collective: every rank owns the rows of its nodes.
// When one matrix is wanted on a device, the exchange is
//
//   1. fegpu_block_counts     entries per column of this rank's block (diff of its colptr)                     [ncols] int64
//   2. all-gather of the counts over NCCL (host side: torch.distributed on the device buffers)           8 B x ncols x P
//   3. fegpu_gather_plan      global colptr = 1 + prefix sum of the column totals; nnz of every rank's block
//   4. contiguous slabs rowval / nzval of every block to the destination (NCCL send / recv)              16 B x nnz, once
//   5. fegpu_gather_place     interleave: column j of the block of rank r lands at colptr[j] + (entries of ranks < r in column j)
//   6. fegpu_gather_sort_columns  only if the owned dof ranges are not ordered by rank (free-first numberings): columns whose
//                              concatenation is not ascending are sorted by row
//
// No destination-index array crosses the wire: the positions are functions of the all-gathered counts.  The kernels are plain
// streaming copies (HBM-bound, 32 B / nnz read + written on the destination).
#include "fegpu_internal.h"

namespace {

__global__ void k_block_counts(const int64_t *__restrict__ colptr, int64_t ncols, int64_t *__restrict__ counts) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < ncols) counts[j] = colptr[j + 1] - colptr[j];
}

// column totals over the ranks (into tot[0..ncols)) and, per block of 256 columns, nothing else: the scan follows
__global__ void k_column_totals(const int64_t *__restrict__ allcounts, int world, int64_t ncols, int64_t *__restrict__ tot) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ncols) return;
  int64_t s = 0;
  for (int r = 0; r < world; r++) s += allcounts[(int64_t)r * ncols + j];
  tot[j] = s;
}

// nnz of every rank's block: one block of threads per rank, grid-stride partial sums + shuffle tree (exact integer sums)
__global__ void __launch_bounds__(256) k_rank_totals(const int64_t *__restrict__ allcounts, int64_t ncols, unsigned long long *__restrict__ nnz_rank) {
  const int r = blockIdx.y;
  int64_t s = 0;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < ncols; j += (int64_t)gridDim.x * blockDim.x) s += allcounts[(int64_t)r * ncols + j];
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(&nnz_rank[r], (unsigned long long)s);
}

// Place the block of rank `src`: 8 lanes per column copy its segment to the global position.
__global__ void __launch_bounds__(256) k_place(const int64_t *__restrict__ allcounts, int world, int src, int64_t ncols,
                                               const int64_t *__restrict__ gcolptr, const int64_t *__restrict__ src_start,
                                               const int64_t *__restrict__ src_rowval, const double *__restrict__ src_nzval,
                                               int64_t *__restrict__ rowval, double *__restrict__ nzval) {
  const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int gl = threadIdx.x & 7;
  if (j >= ncols) return;
  const int64_t len = allcounts[(int64_t)src * ncols + j];
  if (len == 0) return;
  int64_t before = 0;
  for (int r = 0; r < src; r++) before += allcounts[(int64_t)r * ncols + j];
  const int64_t d0 = gcolptr[j] - 1 + before, s0 = src_start[j];
  for (int64_t k = gl; k < len; k += 8) {
    rowval[d0 + k] = src_rowval[s0 + k];
    nzval[d0 + k] = src_nzval[s0 + k];
  }
}

// columns whose rows are not ascending (blocks whose dof ranges interleave): insertion sort, one thread per column.  Rare path.
__global__ void k_sort_unsorted_columns(const int64_t *__restrict__ gcolptr, int64_t ncols, int64_t *__restrict__ rowval, double *__restrict__ nzval,
                                        int *__restrict__ nfixed) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ncols) return;
  const int64_t b = gcolptr[j] - 1, e = gcolptr[j + 1] - 1;
  bool sorted = true;
  for (int64_t k = b + 1; k < e; k++) sorted = sorted && rowval[k - 1] < rowval[k];
  if (sorted) return;
  for (int64_t k = b + 1; k < e; k++) {
    const int64_t r = rowval[k];
    const double v = nzval[k];
    int64_t q = k - 1;
    while (q >= b && rowval[q] > r) {
      rowval[q + 1] = rowval[q];
      nzval[q + 1] = nzval[q];
      q--;
    }
    rowval[q + 1] = r;
    nzval[q + 1] = v;
  }
  atomicAdd(nfixed, 1);
}

struct Guard {
  int prev = -1;
  explicit Guard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~Guard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

}  // namespace

extern "C" {

int32_t fegpu_block_counts(fegpu_asm *as, int64_t *d_counts) {
  if (!as || !d_counts) return fegpu_fail(as ? as->ctx : nullptr, FEGPU_ERR_ARG, "NULL argument");
  fegpu_ctx *ctx = as->ctx;
  if (!as->have_result) return fegpu_fail(ctx, FEGPU_ERR_STATE, "no assembled matrix");
  Guard g(ctx->device);
  const int64_t n = as->r_ncols();
  if (n > 0) {
    k_block_counts<<<grid_for(n, 256), 256, 0, ctx->stream>>>(as->r_colptr(), n, d_counts);
    ctx->launches++;
  }
  CUDA_TRY(ctx, cudaGetLastError());
  return FEGPU_OK;
}

int32_t fegpu_gather_plan(fegpu_ctx *ctx, const int64_t *d_allcounts, int32_t world, int64_t ncols, int64_t *d_colptr, int64_t *d_nnz_rank) {
  if (!ctx || !d_allcounts || !d_colptr || !d_nnz_rank || world < 1 || ncols < 0) return fegpu_fail(ctx, FEGPU_ERR_ARG, "bad argument");
  Guard g(ctx->device);
  cudaStream_t st = ctx->stream;
  CUDA_TRY(ctx, cudaMemsetAsync(d_nnz_rank, 0, sizeof(int64_t) * world, st));
  if (ncols > 0) {
    k_column_totals<<<grid_for(ncols, 256), 256, 0, st>>>(d_allcounts, world, ncols, d_colptr);
    const dim3 grid((unsigned)std::min<int64_t>(grid_for(ncols, 256), (int64_t)ctx->sm_count * 4), (unsigned)world);
    k_rank_totals<<<grid, 256, 0, st>>>(d_allcounts, ncols, reinterpret_cast<unsigned long long *>(d_nnz_rank));
    ctx->launches += 2;
  }
  FE_TRY(fe_exclusive_scan_i64(ctx, d_colptr, d_colptr, ncols, 1, true, nullptr));  // in place: totals -> 1-based colptr [ncols+1]
  CUDA_TRY(ctx, cudaGetLastError());
  return FEGPU_OK;
}

int32_t fegpu_gather_place(fegpu_ctx *ctx, const int64_t *d_allcounts, int32_t world, int32_t src_rank, int64_t ncols, const int64_t *d_colptr,
                           const int64_t *d_src_rowval, const double *d_src_nzval, int64_t *d_rowval, double *d_nzval) {
  if (!ctx || !d_allcounts || !d_colptr || world < 1 || src_rank < 0 || src_rank >= world) return fegpu_fail(ctx, FEGPU_ERR_ARG, "bad argument");
  if (ncols == 0) return FEGPU_OK;
  if (!d_src_rowval || !d_src_nzval || !d_rowval || !d_nzval) return fegpu_fail(ctx, FEGPU_ERR_ARG, "NULL block arrays");
  Guard g(ctx->device);
  cudaStream_t st = ctx->stream;
  int64_t *d_start = nullptr;  // first entry of every column inside the source block: exclusive scan of its counts
  FE_TRY(fe_dev_alloc(ctx, (void **)&d_start, sizeof(int64_t) * (size_t)(ncols + 1), st));
  int32_t s = fe_exclusive_scan_i64(ctx, d_allcounts + (int64_t)src_rank * ncols, d_start, ncols, 0, true, nullptr);
  if (s == FEGPU_OK) {
    k_place<<<grid_for(ncols * 8, 256), 256, 0, st>>>(d_allcounts, world, src_rank, ncols, d_colptr, d_start, d_src_rowval, d_src_nzval, d_rowval, d_nzval);
    ctx->launches++;
    if (cudaGetLastError() != cudaSuccess) s = fegpu_fail(ctx, FEGPU_ERR_CUDA, "k_place launch failed");
  }
  fe_dev_free(ctx, d_start, st);
  return s;
}

int32_t fegpu_gather_sort_columns(fegpu_ctx *ctx, int64_t ncols, const int64_t *d_colptr, int64_t *d_rowval, double *d_nzval, int64_t *columns_sorted) {
  if (!ctx || !d_colptr) return fegpu_fail(ctx, FEGPU_ERR_ARG, "bad argument");
  Guard g(ctx->device);
  cudaStream_t st = ctx->stream;
  int *d_n = nullptr;
  FE_TRY(fe_dev_alloc(ctx, (void **)&d_n, sizeof(int), st));
  CUDA_TRY(ctx, cudaMemsetAsync(d_n, 0, sizeof(int), st));
  if (ncols > 0) {
    k_sort_unsorted_columns<<<grid_for(ncols, 128), 128, 0, st>>>(d_colptr, ncols, d_rowval, d_nzval, d_n);
    ctx->launches++;
  }
  int h = 0;
  CUDA_TRY(ctx, cudaMemcpyAsync(&h, d_n, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(ctx, cudaStreamSynchronize(st));
  fe_dev_free(ctx, d_n, st);
  if (columns_sorted) *columns_sorted = h;
  return FEGPU_OK;
}

}  // extern "C"
