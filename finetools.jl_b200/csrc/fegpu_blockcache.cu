// Device-memory block cache of a context: the symbolic phase allocates and frees ~25 arrays per pattern build (up to
// several GB each), always the same sizes for the same mesh.  The driver's stream-ordered pool (cudaMallocAsync) serves
// them from its free blocks, but its slow path -- re-mapping physical memory into a fresh virtual range when the free
// blocks do not line up -- showed up as host stalls of 10 ms to 200 ms with the GPU idle, at random, on back-to-back fresh
// assemblies of the 256^3 block (FEGPU_TRACE timeline, profiles/).  This cache keeps every freed block and hands it out
// again on an exact or near fit, so after the first rebuild the hot path makes no driver allocation call at all.
//
// Stream order: a block freed on stream S may be reused by work queued later on S without further ado.  Reuse on another
// stream T first records an event at the current tail of S (which is behind the free) and makes T wait for it.
#include <unordered_map>

#include "fegpu_internal.h"

struct BlockCache {
  struct Block {
    void *p;
    size_t cap;
    cudaStream_t stream;  // stream the block was freed on
  };
  std::vector<Block> free_list;
  std::unordered_map<void *, size_t> live;  // blocks handed out -> capacity
  size_t free_bytes = 0;
  size_t limit = 0;  // cached (free) bytes above which a miss first returns everything to the driver
  cudaEvent_t ev = nullptr;
  int64_t hits = 0, misses = 0;
};

static BlockCache *cache_of(fegpu_ctx *ctx) {
  if (!ctx->blocks) {
    ctx->blocks = new BlockCache();
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) tot = (size_t)64 << 30;
    ctx->blocks->limit = tot / 5 * 2;
    cudaEventCreateWithFlags(&ctx->blocks->ev, cudaEventDisableTiming);
  }
  return ctx->blocks;
}

static void release_free_blocks(BlockCache *c) {
  if (c->free_list.empty()) return;
  cudaDeviceSynchronize();  // pending work may still use a block that was freed in stream order
  for (auto &b : c->free_list) cudaFree(b.p);
  c->free_list.clear();
  c->free_bytes = 0;
}

int32_t fe_dev_alloc(fegpu_ctx *ctx, void **p, size_t bytes, cudaStream_t stream) {
  *p = nullptr;
  BlockCache *c = cache_of(ctx);
  bytes = (std::max<size_t>(bytes, 1) + 511) & ~(size_t)511;
  // smallest cached block that fits without wasting more than the request again (1 MB of slack for small ones); among
  // equals the one freed on this stream (no cross-stream wait)
  int best = -1;
  const size_t cap_max = bytes + std::max<size_t>(bytes, (size_t)1 << 20);
  for (int i = 0; i < (int)c->free_list.size(); i++) {
    const auto &b = c->free_list[i];
    if (b.cap < bytes || b.cap > cap_max) continue;
    if (best < 0 || b.cap < c->free_list[best].cap || (b.cap == c->free_list[best].cap && b.stream == stream && c->free_list[best].stream != stream))
      best = i;
  }
  if (best >= 0) {
    BlockCache::Block b = c->free_list[best];
    c->free_list[best] = c->free_list.back();
    c->free_list.pop_back();
    c->free_bytes -= b.cap;
    if (b.stream != stream) {
      // everything queued on the freeing stream so far (the free point included) must finish before the new owner writes
      if (!c->ev || cudaEventRecord(c->ev, b.stream) != cudaSuccess || cudaStreamWaitEvent(stream, c->ev, 0) != cudaSuccess) {
        cudaGetLastError();
        cudaDeviceSynchronize();  // e.g. the caller destroyed the stream it had lent to the context
      }
    }
    c->live[b.p] = b.cap;
    c->hits++;
    *p = b.p;
    return FEGPU_OK;
  }
  c->misses++;
  // many meshes of different sizes through one context: do not let the list (linear scan) or the footprint grow without bound
  if (c->free_bytes + bytes > c->limit || c->free_list.size() > 512) release_free_blocks(c);
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    release_free_blocks(c);
    e = cudaMalloc(p, bytes);
  }
  if (e != cudaSuccess) {
    *p = nullptr;
    return fegpu_fail(ctx, FEGPU_ERR_CUDA, std::string("cudaMalloc of ") + std::to_string(bytes) + " bytes: " + cudaGetErrorString(e));
  }
  c->live[*p] = bytes;
  return FEGPU_OK;
}

void fe_dev_free(fegpu_ctx *ctx, void *p, cudaStream_t stream) {
  if (!p) return;
  BlockCache *c = cache_of(ctx);
  auto it = c->live.find(p);
  if (it == c->live.end()) {  // not ours (should not happen): fall back to the driver
    cudaFree(p);
    return;
  }
  c->free_list.push_back(BlockCache::Block{p, it->second, stream});
  c->free_bytes += it->second;
  c->live.erase(it);
}

void fe_dev_cache_stats(fegpu_ctx *ctx, int64_t *hits, int64_t *misses, size_t *free_bytes) {
  BlockCache *c = cache_of(ctx);
  if (hits) *hits = c->hits;
  if (misses) *misses = c->misses;
  if (free_bytes) *free_bytes = c->free_bytes;
}

void fe_dev_cache_trim(fegpu_ctx *ctx) {
  if (ctx->blocks) release_free_blocks(ctx->blocks);
}

void fe_dev_cache_destroy(fegpu_ctx *ctx) {
  BlockCache *c = ctx->blocks;
  if (!c) return;
  release_free_blocks(c);
  // blocks still handed out belong to handles the caller has not destroyed yet (contract: handles go before their context)
  if (c->ev) cudaEventDestroy(c->ev);
  delete c;
  ctx->blocks = nullptr;
}
