// Mesh-structured CSC construction: replaces SparseArrays.sparse(I,J,V,m,n) (AssemblyModule.jl:319-325) for
// assemblies whose triplets come from element matrices scattered through one dof map.
//
// Symbolic phase (once per mesh + dof map + partition; cached in the dofmap):
//   node -> element adjacency (CSR, ascending element order)             k_count_adj / k_fill_adj / k_sort_adj
//   k_nbr (one warp per node): candidate neighbour nodes of all adjacent elements -> 32-bit bitonic sort (in registers,
//     shuffle network) -> unique list U; every candidate (adjacent element a, local row node li) finds its neighbour
//     slot by binary search: cslot[a*nne + li].
//   colptr from per-column counts (the ndn columns of a node share one row set)
//   k_rows: rowval = dofs of (neighbour, component) in ascending order; when node-major order is not already
//     ascending (free-first/fixed-last numbering, FieldModule.jl:360-377) a per-node sort produces a rank table.
// Numeric phase (every assembly): k_gather -- a group of 8/16/32 lanes per column node walks the node's adjacent elements
//   in ascending order; for each it reads the ndn contiguous element-matrix columns of that node (coalesced) and adds
//   every value into the node's accumulator slot (shared memory) given by cslot.  Ascending element order = the
//   reference's left-to-right duplicate sum; no atomics => bit-reproducible (test/test_basics.jl:3039-3045).
//
// The pattern equals sparse()'s: one entry per (row dof, col dof) pair sharing an element, explicit zeros kept, rows
// strictly increasing in a column, 1-based int64 colptr/rowval.
#include <cstdlib>

#include <cstdio>

#include "fegpu_internal.h"
#include "fegpu_pattern.h"


namespace {

constexpr int WPB = 4;  // warps per block in the per-node symbolic kernels
constexpr int GWPB = 8; // warps per block in the gather
#ifndef GATHER_MIN_CTAS
#define GATHER_MIN_CTAS 5  // 48 registers; 6 (40 registers, small spills) was measured: config 2 gather 3.32 -> 4.67 ms, config 4 unchanged
#endif
// GBATCH (template): adjacent elements whose loads are in flight together (k_gather, one-value-per-lane path)

struct SymParams {
  const int32_t *conn;
  const int32_t *elem_list;
  int64_t nactive;
  int nne;
  int64_t nnodes;
  const uint8_t *rowowned;
  const int32_t *dof;  // [ndn][nnodes]
  int ndn;
  // nodes that have at least one active element (ascending), nullptr = all nodes; the per-node kernels visit only these, so
  // a rank's symbolic cost follows its own share of the mesh (row-block partitions, boundary skins)
  const int32_t *anodes;
  int64_t na;
  // when set, the number of entries of `anodes` lives on the device (class lists built by k_classify_nodes) and `na` is only
  // an upper bound used to size the grid
  const int64_t *na_dev;
};
__device__ __forceinline__ int64_t active_node(const SymParams &S, int64_t idx) { return S.anodes ? (int64_t)S.anodes[idx] : idx; }

// The per-node passes of the symbolic phase run over the mesh's node window [lo, lo + nw) (fegpu_mesh::win_lo/hi: every node
// of an active element lies inside; the whole mesh when it is not partitioned), so that a rank's fixed costs follow its own
// share of the mesh.  flag / pos are window-relative, the list holds global node ids.
__global__ void k_flag_active(const int32_t *__restrict__ deg, int64_t lo, int64_t nw, int32_t *__restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nw) flag[i] = deg[lo + i] > 0 ? 1 : 0;
}
__global__ void k_compact_active(const int32_t *__restrict__ flag, const int64_t *__restrict__ pos, int64_t lo, int64_t nw,
                                 int32_t *__restrict__ list) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nw && flag[i]) list[pos[i]] = (int32_t)(lo + i);
}
// Entries outside [lo, hi] of a windowed prefix array (adjptr, nbrptr, colptr): `before` ahead of the window, the window's
// total (a[hi]) behind it, so every consumer still sees a complete, non-decreasing array.  Two arrays per launch.
__global__ void k_fill_outside(int64_t *__restrict__ a0, int64_t *__restrict__ a1, int64_t len, int64_t lo, int64_t hi, int64_t before) {
  const int64_t nout = len - (hi - lo + 1);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nout) return;
  const int64_t idx = (i < lo) ? i : i + (hi - lo + 1);
  if (a0) a0[idx] = (i < lo) ? before : a0[hi];
  if (a1) a1[idx] = (i < lo) ? before : a1[hi];
}

// The count pass keeps what its atomic returns -- the position of this (element, local node) inside the node's list -- so that
// the fill pass is a plain scatter without a second round of atomics.
__global__ void k_count_adj(SymParams S, int32_t *deg, uint16_t *__restrict__ arank, int *degenerate) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.nactive * S.nne) return;
  int64_t slot = i / S.nne;
  int lc = (int)(i % S.nne);
  int64_t e = S.elem_list ? S.elem_list[slot] : slot;
  const int32_t *c = S.conn + e * S.nne;
  int n = c[lc];
  arank[i] = (uint16_t)atomicAdd(&deg[n], 1);  // degrees >= 65535 / nne leave the structured path before the fill runs
  for (int k = 0; k < lc; k++)
    if (c[k] == n) *degenerate = 1;
}

__global__ void k_fill_adj(SymParams S, const int64_t *__restrict__ adjptr, const uint16_t *__restrict__ arank, int32_t *__restrict__ adj_slot,
                           uint8_t *__restrict__ adj_lc) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.nactive * S.nne) return;
  int64_t slot = i / S.nne;
  int lc = (int)(i % S.nne);
  int64_t e = S.elem_list ? S.elem_list[slot] : slot;
  int n = S.conn[e * S.nne + lc];
  int64_t pos = adjptr[n] + arank[i];
  adj_slot[pos] = (int32_t)slot;
  adj_lc[pos] = (uint8_t)lc;
}

// ascending slot order inside every node's list (the atomics above deliver an arbitrary order)
__global__ void k_sort_adj(SymParams S, const int64_t *adjptr, int32_t *adj_slot, uint8_t *adj_lc) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= S.na) return;
  const int64_t n = active_node(S, idx);
  int64_t b = adjptr[n], e = adjptr[n + 1];
  for (int64_t i = b + 1; i < e; i++) {
    int32_t s = adj_slot[i];
    uint8_t l = adj_lc[i];
    int64_t j = i - 1;
    while (j >= b && adj_slot[j] > s) {
      adj_slot[j + 1] = adj_slot[j];
      adj_lc[j + 1] = adj_lc[j];
      j--;
    }
    adj_slot[j + 1] = s;
    adj_lc[j + 1] = l;
  }
}

// in-warp bitonic sort of n (power of two) keys in shared memory, ascending
template <typename K>
__device__ void warp_bitonic(K *a, int n, int lane) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < n; i += 32) {
        int ixj = i ^ j;
        if (ixj > i) {
          K x = a[i], y = a[ixj];
          bool up = ((i & k) == 0);
          if ((x > y) == up) {
            a[i] = y;
            a[ixj] = x;
          }
        }
      }
      __syncwarp();
    }
  }
}

__device__ __forceinline__ int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// Bitonic sort of 32*KPL uint32 keys held KPL per lane (element index i = r*32 + lane), ascending, shuffle network.
template <int KPL>
__device__ __forceinline__ void reg_bitonic(uint32_t (&v)[KPL], int lane) {
#pragma unroll
  for (int k = 2; k <= 32 * KPL; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int r = 0; r < KPL; r++) {
        const int i = r * 32 + lane;
        const bool up = ((i & k) == 0);
        if (j >= 32) {
          const int rp = r ^ (j >> 5);
          if (rp > r) {  // both live in this lane
            uint32_t lo = min(v[r], v[rp]), hi = max(v[r], v[rp]);
            v[r] = up ? lo : hi;
            v[rp] = up ? hi : lo;
          }
        } else {
          const uint32_t o = __shfl_xor_sync(0xffffffffu, v[r], j);
          const bool lower = ((lane & j) == 0);
          v[r] = (lower == up) ? min(v[r], o) : max(v[r], o);
        }
      }
    }
  }
}

template <int KPL>
__device__ __forceinline__ void sort_via_regs(uint32_t *work, int lane) {
  uint32_t v[KPL];
#pragma unroll
  for (int r = 0; r < KPL; r++) v[r] = work[r * 32 + lane];
  reg_bitonic<KPL>(v, lane);
#pragma unroll
  for (int r = 0; r < KPL; r++) work[r * 32 + lane] = v[r];
  __syncwarp();
}

// One warp per node.  Shared per warp (uint32 words): el[maxdeg] | cand[capc] | work[capc] | uq[capc]
// NNE_T: nodes per element as a compile-time constant (0 = runtime) so the candidate decode needs no integer division.
template <int NNE_T>
__global__ void __launch_bounds__(WPB * 32) k_nbr(SymParams S, const int64_t *__restrict__ adjptr, const int32_t *__restrict__ adj_slot,
                                                  int maxdeg, int capc, int32_t *__restrict__ nnbr_out, int32_t *__restrict__ U,
                                                  uint16_t *__restrict__ cslot, uint8_t *__restrict__ sorted_flag, int *any_unsorted) {
  extern __shared__ uint32_t su[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int per_warp = maxdeg + 3 * capc;
  uint32_t *el = su + (size_t)w * per_warp;
  uint32_t *cand = el + maxdeg;
  uint32_t *work = cand + capc;
  uint32_t *uq = work + capc;
  const int nne = (NNE_T > 0) ? NNE_T : S.nne, ndn = S.ndn;
  for (int64_t idx = (int64_t)blockIdx.x * WPB + w; idx < S.na; idx += (int64_t)gridDim.x * WPB) {
    const int64_t n = active_node(S, idx);
    const int64_t ab = adjptr[n];
    const int deg = (int)(adjptr[n + 1] - ab);
    if (deg == 0) {
      if (lane == 0) {
        nnbr_out[n] = 0;
        sorted_flag[n] = 1;
      }
      continue;
    }
    const int ncand = deg * nne;
    int p2 = 32;
    while (p2 < ncand) p2 <<= 1;
    for (int a = lane; a < deg; a += 32) {
      int64_t slot = adj_slot[ab + a];
      el[a] = (uint32_t)(S.elem_list ? S.elem_list[slot] : slot);
    }
    __syncwarp();
    // candidates in triplet order k = a*nne + li; rows not owned by this rank are dropped (all ones)
    for (int k = lane; k < p2; k += 32) {
      uint32_t m = 0xffffffffu;
      if (k < ncand) {
        int a = k / nne, li = k - a * nne;
        uint32_t mm = (uint32_t)S.conn[(int64_t)el[a] * nne + li];
        if (!S.rowowned || S.rowowned[mm]) m = mm;
      }
      cand[k] = m;
      work[k] = m;
    }
    __syncwarp();
    switch (p2) {
      case 32: sort_via_regs<1>(work, lane); break;
      case 64: sort_via_regs<2>(work, lane); break;
      case 128: sort_via_regs<4>(work, lane); break;
      case 256: sort_via_regs<8>(work, lane); break;
      case 512: sort_via_regs<16>(work, lane); break;
      default: warp_bitonic(work, p2, lane); break;
    }
    int nu = 0;
    for (int base = 0; base < p2; base += 32) {
      int k = base + lane;
      uint32_t v = work[k];
      bool head = (v != 0xffffffffu) && (k == 0 || work[k - 1] != v);
      unsigned bal = __ballot_sync(0xffffffffu, head);
      if (head) uq[nu + __popc(bal & ((1u << lane) - 1))] = v;
      nu += __popc(bal);
    }
    __syncwarp();
    // neighbour slot of every candidate: binary search in the unique list
    uint16_t *cs = cslot + ab * nne;
    for (int k = lane; k < ncand; k += 32) {
      uint32_t m = cand[k];
      uint16_t s = 0xffffu;
      if (m != 0xffffffffu) {
        int lo = 0, hi = nu - 1;
        while (lo < hi) {
          int mid = (lo + hi) >> 1;
          if (uq[mid] < m) lo = mid + 1;
          else hi = mid;
        }
        s = (uint16_t)lo;
      }
      cs[k] = s;
    }
    // unique neighbour list to global; is the node-major dof order already ascending?
    int32_t *Un = U + ab * nne;
    for (int s = lane; s < nu; s += 32) Un[s] = (int32_t)uq[s];
    // node-major dof order ascending?  per neighbour: its own dofs ascending and its last dof below the next neighbour's first
    bool ok = true;
    for (int s = lane; s < nu; s += 32) {
      const int64_t u = uq[s];
      int prev = S.dof[u];
      for (int p = 1; p < ndn; p++) {
        const int d = S.dof[(int64_t)p * S.nnodes + u];
        if (d <= prev) ok = false;
        prev = d;
      }
      if (s + 1 < nu && S.dof[uq[s + 1]] <= prev) ok = false;
    }
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) {
      nnbr_out[n] = nu;
      sorted_flag[n] = ok ? 1 : 0;
      if (!ok) *any_unsorted = 1;
    }
    __syncwarp();
  }
}

// Bitonic sort of 32*KPL uint32 keys held KPL per lane in LANE-MAJOR order (position i = lane*KPL + r), ascending.  The
// stages with j < KPL compare registers of one lane (no shuffle); the others exchange register r with lane ^ (j/KPL).
template <int KPL>
__device__ __forceinline__ void reg_bitonic_lm(uint32_t (&v)[KPL], int lane) {
#pragma unroll
  for (int k = 2; k <= 32 * KPL; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j < KPL) {
#pragma unroll
        for (int r = 0; r < KPL; r++) {
          const int rp = r ^ j;
          if (rp > r) {
            const bool up = (((lane * KPL + r) & k) == 0);
            const uint32_t lo = min(v[r], v[rp]), hi = max(v[r], v[rp]);
            v[r] = up ? lo : hi;
            v[rp] = up ? hi : lo;
          }
        }
      } else {
        const int lj = j / KPL;
        const bool take_min = (((lane & lj) == 0) == (((lane * KPL) & k) == 0));
#pragma unroll
        for (int r = 0; r < KPL; r++) {
          const uint32_t o = __shfl_xor_sync(0xffffffffu, v[r], lj);
          v[r] = take_min ? min(v[r], o) : max(v[r], o);
        }
      }
    }
  }
}

// Neighbour lists of one node from packed keys (node << KB | candidate index), everything in registers.
//   keys: candidate k = a*nne + li of the node's adjacent element a (ascending element order) -> (conn[el_a][li] << KB) | k;
//         rows not owned by this rank keep k but carry the all-ones node field; padding is all ones
//   sort (lane-major bitonic network), head flags by comparing node fields of consecutive positions, slot = number of heads
//   before the position; every sorted key scatters its slot to cslot[k] (no search), heads write the unique list U.
template <int KPL, bool CHECK>
__device__ __forceinline__ void nbr_node(const SymParams &S, const int64_t n, const int64_t ab, const int deg, const int lane, const int KB,
                                         int32_t *__restrict__ adj_slot, uint8_t *__restrict__ adj_lc, int32_t *__restrict__ nnbr_out,
                                         int32_t *__restrict__ U,
                                         uint16_t *__restrict__ cslot, uint8_t *__restrict__ sorted_flag, int *any_unsorted, int32_t *smF,
                                         int32_t *smL) {
  const int nne = S.nne;
  const int ncand = deg * nne;
  const uint32_t nne_magic = 65536u / (uint32_t)nne + 1u;  // k / nne == (k * magic) >> 16 for k < 2048 (nne <= 32)
  const uint32_t kmask = (1u << KB) - 1u, dropped = 0xffffffffu >> KB;
  // the adjacency list arrives in the arbitrary order of k_fill_adj's atomics: sort it here (slot << 5 | local index, one key
  // per lane) and write it back -- ascending element order is what makes the numeric phase reproduce the reference's sums
  uint32_t el = 0;
  {
    uint32_t key[1] = {0xffffffffu};
    if (lane < deg) key[0] = ((uint32_t)adj_slot[ab + lane] << 5) | (uint32_t)adj_lc[ab + lane];
    reg_bitonic_lm<1>(key, lane);
    if (lane < deg) {
      const int64_t slot = key[0] >> 5;
      adj_slot[ab + lane] = (int32_t)slot;
      adj_lc[ab + lane] = (uint8_t)(key[0] & 31u);
      el = (uint32_t)(S.elem_list ? S.elem_list[slot] : slot);
    }
  }
  uint32_t v[KPL];
#pragma unroll
  for (int r = 0; r < KPL; r++) {
    const int k = r * 32 + lane;  // coalesced reads of the connectivity rows
    const int a = (k < ncand) ? (int)(((uint32_t)k * nne_magic) >> 16) : 0;
    const uint32_t e = __shfl_sync(0xffffffffu, el, a);
    uint32_t key = 0xffffffffu;
    if (k < ncand) {
      const int li = k - a * nne;
      uint32_t m = (uint32_t)S.conn[(int64_t)e * nne + li];
      if (S.rowowned && !S.rowowned[m]) m = dropped;
      key = (m << KB) | (uint32_t)k;
    }
    v[r] = key;
  }
  reg_bitonic_lm<KPL>(v, lane);
  // heads: position i = lane*KPL + r; the predecessor of r = 0 is the last register of the previous lane
  const uint32_t prev_lane_last = __shfl_up_sync(0xffffffffu, v[KPL - 1], 1);
  bool head[KPL];
  unsigned bal[KPL];
  int nu = 0;
#pragma unroll
  for (int r = 0; r < KPL; r++) {
    const uint32_t node = v[r] >> KB;
    const uint32_t pnode = (r == 0) ? (prev_lane_last >> KB) : (v[r - 1] >> KB);
    head[r] = (node != dropped) && ((r == 0 && lane == 0) || node != pnode);
    bal[r] = __ballot_sync(0xffffffffu, head[r]);
    nu += __popc(bal[r]);
  }
  // heads at positions before (lane, r): all registers of the lower lanes + registers r' <= r of this lane
  const unsigned lt = (1u << lane) - 1u;
  int before = 0;
#pragma unroll
  for (int r = 0; r < KPL; r++) before += __popc(bal[r] & lt);
  uint16_t *cs = cslot + ab * nne;
  int32_t *Un = U + ab * nne;
  bool ok = true;
#pragma unroll
  for (int r = 0; r < KPL; r++) {
    before += head[r] ? 1 : 0;
    const uint32_t node = v[r] >> KB, k = v[r] & kmask;
    if ((int)k < ncand) {  // padding carries k = 2^KB - 1 >= ncand (a node with ncand == 2^KB has no padding)
      if (node == dropped) cs[k] = 0xffffu;
      else cs[k] = (uint16_t)(before - 1);
    }
    if (head[r]) {
      Un[before - 1] = (int32_t)node;
      if (CHECK) {
        int prev = S.dof[node];
        smF[before - 1] = prev;
        for (int p = 1; p < S.ndn; p++) {
          const int d = S.dof[(int64_t)p * S.nnodes + node];
          if (d <= prev) ok = false;
          prev = d;
        }
        smL[before - 1] = prev;
      }
    }
  }
  if (CHECK) {
    __syncwarp();
    for (int s = lane; s + 1 < nu; s += 32)
      if (smF[s + 1] <= smL[s]) ok = false;
    ok = __all_sync(0xffffffffu, ok);
    __syncwarp();
  }
  if (lane == 0) {
    nnbr_out[n] = nu;
    if (CHECK) {
      sorted_flag[n] = ok ? 1 : 0;
      if (!ok) *any_unsorted = 1;
    }
  }
}

// One warp per node, register-only version of k_nbr.  Preconditions (checked by the host): maxdeg <= 32, node ids and candidate
// indices fit one 32-bit key (nnodes <= 2^(32-KB) - 2 with 2^KB >= padded candidates).  CHECK = false when the dof map is
// node-major ascending everywhere (k_dof_monotone): no per-node order test, sorted_flag is not written.
template <int MAXKPL, bool CHECK>
__global__ void __launch_bounds__(WPB * 32) k_nbr_fast(SymParams S, const int64_t *__restrict__ adjptr, int32_t *__restrict__ adj_slot,
                                                       uint8_t *__restrict__ adj_lc, int capc, int KB, int32_t *__restrict__ nnbr_out,
                                                       int32_t *__restrict__ U,
                                                       uint16_t *__restrict__ cslot, uint8_t *__restrict__ sorted_flag, int *any_unsorted) {
  extern __shared__ int32_t sfl[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int32_t *smF = CHECK ? sfl + (size_t)w * 2 * capc : nullptr;
  int32_t *smL = CHECK ? smF + capc : nullptr;
  const int nne = S.nne;
  const int64_t na = S.na_dev ? *S.na_dev : S.na;
  for (int64_t idx = (int64_t)blockIdx.x * WPB + w; idx < na; idx += (int64_t)gridDim.x * WPB) {
    const int64_t n = active_node(S, idx);
    const int64_t ab = adjptr[n];
    const int deg = (int)(adjptr[n + 1] - ab);
    if (deg == 0) {
      if (lane == 0) {
        nnbr_out[n] = 0;
        if (CHECK) sorted_flag[n] = 1;
      }
      continue;
    }
    const int ncand = deg * nne;
#define NBR_ARGS S, n, ab, deg, lane, KB, adj_slot, adj_lc, nnbr_out, U, cslot, sorted_flag, any_unsorted, smF, smL
    // MAXKPL (from the mesh's largest candidate count) bounds the variants compiled in, and with them the register count
    if (ncand <= 32) nbr_node<1, CHECK>(NBR_ARGS);
    else if (MAXKPL <= 2 || ncand <= 64) nbr_node<2, CHECK>(NBR_ARGS);
    else if (MAXKPL <= 4 || ncand <= 128) nbr_node<4, CHECK>(NBR_ARGS);
    else if (ncand <= 256) nbr_node<8, CHECK>(NBR_ARGS);
    else nbr_node<16, CHECK>(NBR_ARGS);
#undef NBR_ARGS
  }
}

// Same algorithm with LPN lanes per node (32/LPN nodes per warp) and KPL keys per lane: for small elements (H8: 64
// candidates = 16 lanes x 4 keys, two nodes per warp; Q4/T3: 8 lanes x 4 keys, four nodes per warp) every instruction of the
// network serves several nodes, and more stages stay inside a lane.  Preconditions: maxdeg <= LPN, candidates <= LPN*KPL.
template <int LPN, int KPL>
__device__ __forceinline__ void group_bitonic(uint32_t (&v)[KPL], int gl) {
  // Bitonic sorter written with ascending comparators only: the first step of every merge pairs position i with
  // i ^ (k-1) (the mirrored position), the following steps with i ^ j.  The lower position always keeps the minimum, so
  // in-lane steps are a plain (min, max) pair and cross-lane steps one predicated min/max per key.
#pragma unroll
  for (int k = 2; k <= LPN * KPL; k <<= 1) {
    if (k <= KPL) {
#pragma unroll
      for (int r = 0; r < KPL; r++) {
        const int rp = r ^ (k - 1);
        if (rp > r) {
          const uint32_t lo = min(v[r], v[rp]), hi = max(v[r], v[rp]);
          v[r] = lo;
          v[rp] = hi;
        }
      }
    } else {
      const int lm = k / KPL - 1;                          // lane part of the mirror mask; the register part is r ^ (KPL-1)
      const bool lower = (gl & (k / (2 * KPL))) == 0;      // the highest flipped bit decides which side this lane is on
      uint32_t o[KPL];
#pragma unroll
      for (int r = 0; r < KPL; r++) o[r] = __shfl_xor_sync(0xffffffffu, v[KPL - 1 - r], lm);
#pragma unroll
      for (int r = 0; r < KPL; r++) v[r] = lower ? min(v[r], o[r]) : max(v[r], o[r]);
    }
#pragma unroll
    for (int j = k >> 2; j > 0; j >>= 1) {
      if (j < KPL) {
#pragma unroll
        for (int r = 0; r < KPL; r++) {
          const int rp = r ^ j;
          if (rp > r) {
            const uint32_t lo = min(v[r], v[rp]), hi = max(v[r], v[rp]);
            v[r] = lo;
            v[rp] = hi;
          }
        }
      } else {
        const int lj = j / KPL;
        const bool lower = (gl & lj) == 0;
#pragma unroll
        for (int r = 0; r < KPL; r++) {
          const uint32_t o = __shfl_xor_sync(0xffffffffu, v[r], lj);
          v[r] = lower ? min(v[r], o) : max(v[r], o);
        }
      }
    }
  }
}

template <int LPN, int KPL, bool CHECK>
__global__ void __launch_bounds__(WPB * 32) k_nbr_group(SymParams S, const int64_t *__restrict__ adjptr, int32_t *__restrict__ adj_slot,
                                                        uint8_t *__restrict__ adj_lc, int capc, int KB, int32_t *__restrict__ nnbr_out,
                                                        int32_t *__restrict__ U,
                                                        uint16_t *__restrict__ cslot, uint8_t *__restrict__ sorted_flag, int *any_unsorted) {
  extern __shared__ int32_t sfl[];
  constexpr int NPW = 32 / LPN;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int g = lane / LPN, gl = lane % LPN;
  int32_t *smF = CHECK ? sfl + (size_t)(w * NPW + g) * 2 * capc : nullptr;
  int32_t *smL = CHECK ? smF + capc : nullptr;
  const int nne = S.nne;
  const uint32_t nne_magic = 65536u / (uint32_t)nne + 1u;
  const uint32_t kmask = (1u << KB) - 1u, dropped = 0xffffffffu >> KB;
  const unsigned gmask = (LPN == 32) ? 0xffffffffu : (((1u << LPN) - 1u) << (g * LPN));
  const unsigned lt = ((1u << lane) - 1u) & gmask;  // lower lanes of this node's group
  const int64_t stride = (int64_t)gridDim.x * WPB * NPW;
  const int64_t first = ((int64_t)blockIdx.x * WPB + w) * NPW;  // warp-uniform; the groups of a warp take consecutive nodes
  // software pipeline: the adjacency range and element ids of the next node are fetched while this one is sorted
  int64_t ab = 0;
  int deg = 0;
  uint32_t el = 0;
  int64_t n = -1;
  const int64_t na = S.na_dev ? *S.na_dev : S.na;
  auto fetch = [&](int64_t idx) {  // called by all lanes of the warp together
    n = -1; ab = 0; deg = 0; el = 0;
    uint32_t key[1] = {0xffffffffu};
    if (idx < na) {
      n = active_node(S, idx);
      ab = adjptr[n];
      deg = (int)(adjptr[n + 1] - ab);
      if (gl < deg) key[0] = ((uint32_t)adj_slot[ab + gl] << 5) | (uint32_t)adj_lc[ab + gl];
    }
    // sort the node's adjacency (k_fill_adj's atomics deliver an arbitrary order) and write it back, see nbr_node
    group_bitonic<LPN, 1>(key, gl);
    if (gl < deg) {
      const int64_t slot = key[0] >> 5;
      adj_slot[ab + gl] = (int32_t)slot;
      adj_lc[ab + gl] = (uint8_t)(key[0] & 31u);
      el = (uint32_t)(S.elem_list ? S.elem_list[slot] : slot);
    }
  };
  fetch(first + g);
  for (int64_t base = first; base < na; base += stride) {  // trip count is warp-uniform
    const int64_t n_cur = n, ab_cur = ab;
    const int deg_cur = deg;
    const uint32_t el_cur = el;
    const int ncand = deg_cur * nne;
    uint32_t v[KPL];
#pragma unroll
    for (int r = 0; r < KPL; r++) {
      const int k = r * LPN + gl;
      const int a = (k < ncand) ? (int)(((uint32_t)k * nne_magic) >> 16) : 0;
      const uint32_t e = __shfl_sync(0xffffffffu, el_cur, g * LPN + a);
      uint32_t key = 0xffffffffu;
      if (k < ncand) {
        const int li = k - a * nne;
        uint32_t m = (uint32_t)S.conn[(int64_t)e * nne + li];
        if (S.rowowned && !S.rowowned[m]) m = dropped;
        key = (m << KB) | (uint32_t)k;
      }
      v[r] = key;
    }
    fetch(base + stride + g);
    group_bitonic<LPN, KPL>(v, gl);
    const uint32_t prev_lane_last = __shfl_up_sync(0xffffffffu, v[KPL - 1], 1);
    bool head[KPL];
    int nu = 0, before = 0;
#pragma unroll
    for (int r = 0; r < KPL; r++) {
      const uint32_t node = v[r] >> KB;
      const uint32_t pnode = (r == 0) ? (prev_lane_last >> KB) : (v[r - 1] >> KB);
      head[r] = (node != dropped) && ((r == 0 && gl == 0) || node != pnode);
      const unsigned bal = __ballot_sync(0xffffffffu, head[r]);
      nu += __popc(bal & gmask);
      before += __popc(bal & lt);
    }
    if (deg_cur > 0) {
      uint16_t *cs = cslot + ab_cur * nne;
      int32_t *Un = U + ab_cur * nne;
      bool ok = true;
#pragma unroll
      for (int r = 0; r < KPL; r++) {
        before += head[r] ? 1 : 0;
        const uint32_t node = v[r] >> KB, k = v[r] & kmask;
        if ((int)k < ncand) cs[k] = (node == dropped) ? (uint16_t)0xffffu : (uint16_t)(before - 1);
        if (head[r]) {
          Un[before - 1] = (int32_t)node;
          if (CHECK) {
            int prev = S.dof[node];
            smF[before - 1] = prev;
            for (int p = 1; p < S.ndn; p++) {
              const int d = S.dof[(int64_t)p * S.nnodes + node];
              if (d <= prev) ok = false;
              prev = d;
            }
            smL[before - 1] = prev;
          }
        }
      }
      if (CHECK) {
        __syncwarp(gmask);
        for (int s2 = gl; s2 + 1 < nu; s2 += LPN)
          if (smF[s2 + 1] <= smL[s2]) ok = false;
        ok = (__ballot_sync(gmask, !ok) & gmask) == 0;
        __syncwarp(gmask);
      }
      if (gl == 0) {
        nnbr_out[n_cur] = nu;
        if (CHECK) {
          sorted_flag[n_cur] = ok ? 1 : 0;
          if (!ok) *any_unsorted = 1;
        }
      }
    } else if (n_cur >= 0 && gl == 0) {
      nnbr_out[n_cur] = 0;
      if (CHECK) sorted_flag[n_cur] = 1;
    }
  }
}

// Meshes whose nodes differ a lot in valence (T10: vertex nodes see ~24 elements, edge nodes ~5; H20: corners 8, edges 4):
// the nodes are binned by candidate count so that the many small nodes run the small-register instantiation of k_nbr_fast.
// Lists are filled with warp-aggregated atomics (their internal order is irrelevant: every node writes its own outputs).
__global__ void k_classify_nodes(SymParams S, const int32_t *__restrict__ deg, int32_t *__restrict__ l0, int32_t *__restrict__ l1,
                                 int32_t *__restrict__ l2, unsigned long long *__restrict__ counts) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int cls = -1;
  int32_t n = 0;
  if (idx < S.na) {
    n = (int32_t)active_node(S, idx);
    const int ncand = deg[n] * S.nne;
    cls = (ncand == 0) ? -1 : (ncand <= 64 ? 0 : (ncand <= 128 ? 1 : 2));
  }
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const unsigned m = __ballot_sync(0xffffffffu, cls == c);
    if (m == 0) continue;
    unsigned long long base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(&counts[c], (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (cls == c) (c == 0 ? l0 : (c == 1 ? l1 : l2))[base + __popc(m & ((1u << lane) - 1u))] = n;
  }
}

// is the dof map node-major ascending (dofs of a node ascending by component, below every dof of the next node)?  Then the
// rows of every column are ascending in (neighbour, component) order and no node needs an order test or a rank table.
// Over the node window only (every row and column node of the pattern is in it).  Also the dof range of the window's nodes,
// the part of colptr that needs a scan: range[0] = max(INT32_MAX - dof), range[1] = max(dof + 1) (zero-initialised slots).
__global__ void __launch_bounds__(256) k_dof_monotone(const int32_t *__restrict__ dof, int64_t nnodes, int ndn, int64_t lo, int64_t nw,
                                                      int *violated, int *__restrict__ range) {
  __shared__ int s_min[8], s_max[8];
  int dmin = INT32_MAX, dmax = -1;
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = lo + i;
    int prev = dof[n];
    dmin = min(dmin, prev);
    dmax = max(dmax, prev);
    for (int p = 1; p < ndn; p++) {
      const int d = dof[(int64_t)p * nnodes + n];
      bad = bad || d <= prev;
      prev = d;
      dmin = min(dmin, d);
      dmax = max(dmax, d);
    }
    if (i + 1 < nw && dof[n + 1] <= prev) bad = true;
  }
  for (int d = 16; d > 0; d >>= 1) {
    dmin = min(dmin, __shfl_xor_sync(0xffffffffu, dmin, d));
    dmax = max(dmax, __shfl_xor_sync(0xffffffffu, dmax, d));
  }
  if ((threadIdx.x & 31) == 0) {
    s_min[threadIdx.x >> 5] = dmin;
    s_max[threadIdx.x >> 5] = dmax;
  }
  if (bad) *violated = 1;
  __syncthreads();
  if (threadIdx.x == 0) {  // one pair of atomics per block: they all land on the same two words
    for (int k = 1; k < 8; k++) {
      dmin = min(dmin, s_min[k]);
      dmax = max(dmax, s_max[k]);
    }
    if (dmax >= 0) {
      atomicMax(&range[0], INT32_MAX - dmin);
      atomicMax(&range[1], dmax + 1);
    }
  }
}

__global__ void k_col_counts(SymParams S, const int32_t *nnbr, int64_t lo, int64_t nw, int64_t *colcount) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nw * S.ndn) return;
  int64_t n = lo + i % nw;
  int q = (int)(i / nw);
  if (nnbr[n] > 0) colcount[S.dof[(int64_t)q * S.nnodes + n]] = (int64_t)nnbr[n] * S.ndn;
}

// rowval for the nodes whose node-major dof order is already ascending (the common case: every node unless the numbering
// puts fixed dofs last).  LPN lanes per node so a warp keeps 32/LPN dependent-load chains (node -> neighbour -> dof) in
// flight; NDN compile-time (0 = runtime) so the (slot, component) decode needs no division.
template <int LPN, int NDN>
// vector fields: 32 registers and all 64 warps of an SM resident (the kernel waits on its node -> neighbour -> dof load chain:
// 1.39 -> 1.22 ms on config 2); the scalar instantiations spill at 32 registers and were slower, they keep 40
__global__ void __launch_bounds__(256, (NDN == 2 || NDN == 3) ? 8 : 6) k_rows_sorted(SymParams S, const int64_t *__restrict__ adjptr, const int32_t *__restrict__ nnbr,
                                                     const int64_t *__restrict__ nbrptr, const int32_t *__restrict__ U,
                                                     const uint8_t *__restrict__ sorted_flag, const int64_t *__restrict__ colptr,
                                                     int64_t *__restrict__ rowval, uint16_t *__restrict__ rank, int32_t *__restrict__ nbr_compact) {
  constexpr int QMAX = (NDN > 0) ? NDN : 6;
  constexpr int UNR = 4;  // row chunks whose neighbour -> dof loads are issued together
  const int ndn = (NDN > 0) ? NDN : S.ndn;
  const int gl = threadIdx.x % LPN;
  const int64_t idx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPN;
  if (idx >= S.na) return;
  const int64_t n = active_node(S, idx);
  const int nu = nnbr[n];
  const uint8_t sf = sorted_flag[n];
  const int64_t ab = adjptr[n];
  int64_t cbase[QMAX];
#pragma unroll
  for (int q = 0; q < QMAX; q++) cbase[q] = (q < ndn) ? colptr[S.dof[(int64_t)q * S.nnodes + n]] - 1 : 0;
  if (nu == 0 || !sf) return;
  const int32_t *Un = U + ab * S.nne;
  const int nr = nu * ndn;
  const int64_t nb = (rank || nbr_compact) ? nbrptr[n] : 0;
  if (nbr_compact)
    for (int s = gl; s < nu; s += LPN) nbr_compact[nb + s] = Un[s];
  for (int i0 = 0; i0 < nr; i0 += UNR * LPN) {
    int64_t rdof[UNR];
#pragma unroll
    for (int u = 0; u < UNR; u++) {
      const int i = i0 + u * LPN + gl;
      if (i < nr) {
        const int s = i / ndn, p = i - s * ndn;
        rdof[u] = (int64_t)S.dof[(int64_t)p * S.nnodes + Un[s]] + 1;
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; u++) {
      const int i = i0 + u * LPN + gl;
      if (i < nr) {
#pragma unroll
        for (int q = 0; q < QMAX; q++)
          if (q < ndn) rowval[cbase[q] + i] = rdof[u];
        if (rank) rank[nb * ndn + i] = (uint16_t)i;
      }
    }
  }
}

// rowval (+ rank).  One warp per node; shared per warp: cap2 uint64 keys, used only by nodes whose dof order needs a sort.
__global__ void __launch_bounds__(WPB * 32) k_rows(SymParams S, const int64_t *__restrict__ adjptr, const int32_t *__restrict__ nnbr,
                                                   const int64_t *__restrict__ nbrptr, const int32_t *__restrict__ U,
                                                   const uint8_t *__restrict__ sorted_flag, const int64_t *__restrict__ colptr, int cap2,
                                                   int64_t *__restrict__ rowval, uint16_t *__restrict__ rank) {
  extern __shared__ unsigned long long sk[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned long long *keys = sk + (size_t)w * cap2;
  const int ndn = S.ndn;
  for (int64_t idx = (int64_t)blockIdx.x * WPB + w; idx < S.na; idx += (int64_t)gridDim.x * WPB) {
    const int64_t n = active_node(S, idx);
    const int nu = nnbr[n];
    if (nu == 0 || sorted_flag[n]) continue;
    const int32_t *Un = U + adjptr[n] * S.nne;
    const int64_t nb = nbrptr[n];
    const int nr = nu * ndn;
    int64_t cbase[6];
    for (int q = 0; q < ndn; q++) cbase[q] = colptr[S.dof[(int64_t)q * S.nnodes + n]] - 1;
    if (sorted_flag[n]) {
      continue;  // written by k_rows_sorted
    } else {
      const int q2 = next_pow2(nr);
      for (int i = lane; i < q2; i += 32) {
        unsigned long long key = ~0ull;
        if (i < nr) {
          int s = i / ndn, p = i - s * ndn;
          key = ((unsigned long long)(unsigned)S.dof[(int64_t)p * S.nnodes + Un[s]] << 16) | (unsigned)i;
        }
        keys[i] = key;
      }
      __syncwarp();
      warp_bitonic(keys, q2, lane);
      for (int pos = lane; pos < nr; pos += 32) {
        unsigned i = (unsigned)(keys[pos] & 0xffffu);
        int64_t rdof = (int64_t)(keys[pos] >> 16) + 1;
        rank[nb * ndn + i] = (uint16_t)pos;
        for (int q = 0; q < ndn; q++) rowval[cbase[q] + pos] = rdof;
      }
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------------- numeric gather
struct GatherParams {
  int64_t nnodes;
  int nne, ndn;
  const int64_t *adjptr;
  const int32_t *adj_slot;
  const uint8_t *adj_lc;
  const int32_t *nnbr;
  const int64_t *nbrptr;
  const uint16_t *cslot;
  const uint16_t *rank;
  const int32_t *dof;
  const int64_t *colptr;
  const double *V;
  double *nzval;
  int maxnbr, maxdeg, maxcand;
  const int32_t *order;  // visiting order (active nodes, Morton-sorted) or nullptr = all nodes in natural order
  int64_t npos;          // visiting positions
};

// Offset (inside one element's values) of entry (row node li, row comp p; column node lc, column comp q).
// Full layout: emission order, column (lc, q) then row (li, p).  Compact layout (symmetric forms, fegpu_internal.h): the
// upper block (min, max), read transposed when the row node comes after the column node.
template <bool COMPACT>
__device__ __forceinline__ int elem_offset(int li, int p, int lc, int q, int ndn, int EM) {
  if (!COMPACT) return (lc * ndn + q) * EM + li * ndn + p;
  const int nd2 = ndn * ndn;
  return (li <= lc) ? nd2 * (lc * (lc + 1) / 2 + li) + q * ndn + p : nd2 * (li * (li + 1) / 2 + lc) + p * ndn + q;
}

// Numeric CSC phase.  LPN lanes per column node (32/LPN nodes per warp), NDN dofs per node (0 = runtime), RPL element-
// matrix rows per lane (>= ceil(EM/LPN); 0 = runtime loop), GBATCH adjacent elements whose loads are in flight together.
// The kernel is bound by dependent-load latency (node -> adjacency -> element values -> column pointer), so the design
// goal is many nodes in flight per SM (small LPN), all loads of a batch issued back to back, and the scalars of the
// next node / the column bases fetched early.
// Shared per node group: acc[maxnbr*ndn*ndn] doubles | base[maxdeg] int64 | cs[maxcand] uint16 (padded to 8 B)
template <int LPN, int NDN, bool COMPACT, int GBATCH, int RPL>
__global__ void __launch_bounds__(GWPB * 32, GATHER_MIN_CTAS) k_gather(const GatherParams G) {
  extern __shared__ double sacc[];
  constexpr int NPW = 32 / LPN;
  constexpr int QMAX = (NDN > 0) ? NDN : 6;
  constexpr int RMAX = (RPL > 0) ? RPL : 1;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int g = lane / LPN, gl = lane % LPN;
  const int ndn = (NDN > 0) ? NDN : G.ndn;
  const int nne = G.nne;
  const int EM = nne * ndn;
  const int64_t EM2 = COMPACT ? (int64_t)(nne * (nne + 1) / 2) * ndn * ndn : (int64_t)EM * EM;  // values per element
  const int acc_stride = G.maxnbr * ndn * ndn;
  const int cs_words = (G.maxcand + 3) / 4;
  const int grp_words = acc_stride + G.maxdeg + cs_words;
  double *acc = sacc + (size_t)(w * NPW + g) * grp_words;
  long long *base = reinterpret_cast<long long *>(acc + acc_stride);
  uint16_t *cs = reinterpret_cast<uint16_t *>(base + G.maxdeg);
  // node counts fit 32 bits (dof numbers are int32 on the device): 32-bit loop state keeps the kernel at 48 registers
  const int nnodes = (int)G.nnodes, npos = (int)G.npos;
  const int groups_total = (int)gridDim.x * GWPB * NPW;
  const int niter = (npos + groups_total - 1) / groups_total;
  // rows of the element matrix this lane adds: r = gl + j*LPN
  int rli[RMAX], rp[RMAX];
  bool rok[RMAX];
#pragma unroll
  for (int j = 0; j < RMAX; j++) {
    const int r = gl + j * LPN;
    rok[j] = r < EM;
    rli[j] = rok[j] ? r / ndn : 0;
    rp[j] = rok[j] ? r - rli[j] * ndn : 0;
  }
  // visiting position -> node (Morton order when G.order); node ids are prefetched two iterations ahead, scalars one
  int pos = (int)blockIdx.x * (GWPB * NPW) + w * NPW + g;
  auto node_at = [&](int p) -> int { return (p < npos && p >= 0) ? (G.order ? G.order[p] : p) : nnodes; };
  int n = node_at(pos);
  int n_next = node_at(pos + groups_total);
  int nn_pre = (n < nnodes) ? G.nnbr[n] : 0;
  int64_t ab_pre = (n < nnodes) ? G.adjptr[n] : 0;
  int deg_pre = (n < nnodes) ? (int)(G.adjptr[n + 1] - ab_pre) : 0;
  for (int it = 0; it < niter; it++) {
    const bool live = n < nnodes;
    const int nn = nn_pre;
    const int64_t ab = ab_pre;
    const int deg = (live && nn > 0) ? deg_pre : 0;
    const int per_col = nn * ndn;
    const int total = per_col * ndn;
    // stage the node's metadata (one round trip to memory), clear the accumulators
    // element base (in values) with the column node's local index packed into the low 6 bits
    for (int a = gl; a < deg; a += LPN) base[a] = (((long long)G.adj_slot[ab + a] * EM2) << 6) | (long long)G.adj_lc[ab + a];
    {
      const uint16_t *csg = G.cslot + ab * nne;
      for (int k = gl; k < deg * nne; k += LPN) cs[k] = csg[k];
    }
    // issued early, consumed late: column bases of this node, scalars of the next node
    long long cb[QMAX];
#pragma unroll
    for (int q = 0; q < QMAX; q++) cb[q] = (q < ndn && deg > 0) ? G.colptr[G.dof[(int64_t)q * G.nnodes + n]] - 1 : 0;
    const int n_next2 = node_at(pos + 2 * groups_total);
    if (n_next < nnodes) {
      nn_pre = G.nnbr[n_next];
      ab_pre = G.adjptr[n_next];
      deg_pre = (int)(G.adjptr[n_next + 1] - ab_pre);
    } else {
      nn_pre = 0; ab_pre = 0; deg_pre = 0;
    }
    for (int i = gl; i < total; i += LPN) acc[i] = 0.0;
    int maxdeg = deg;
#pragma unroll
    for (int d = LPN; d < 32; d <<= 1) maxdeg = max(maxdeg, __shfl_xor_sync(0xffffffffu, maxdeg, d));
    __syncwarp();
    if (RPL > 0) {
      for (int a0 = 0; a0 < maxdeg; a0 += GBATCH) {
        // Compact layout: the block of (row node li, column node lc) is stored once, as block (min, max) with the entry
        // (component i of min, component k of max) at k*ndn + i.  When li > lc the block is read in the SAME address order
        // as an upper block (lane component fastest: neighbouring lanes read neighbouring words) and what the lane holds is
        // then the transposed entry -- row component q, column component rp -- which only changes the accumulator index.
        double v[GBATCH][RMAX][QMAX];
        int lcs[GBATCH];
#pragma unroll
        for (int u = 0; u < GBATCH; u++) {
          lcs[u] = 0;
          if (a0 + u < deg) {
            const long long bu = base[a0 + u];
            const double *Vb = G.V + (bu >> 6);
            const int lc = (int)(bu & 63);
            lcs[u] = lc;
#pragma unroll
            for (int j = 0; j < RMAX; j++)
              if (rok[j]) {
                if (COMPACT) {
                  const int li = rli[j];
                  const int blk = (li <= lc) ? lc * (lc + 1) / 2 + li : li * (li + 1) / 2 + lc;
                  const double *Bp = Vb + ndn * ndn * blk + rp[j];
#pragma unroll
                  for (int q = 0; q < QMAX; q++)
                    if (q < ndn) v[u][j][q] = Bp[q * ndn];
                } else {
#pragma unroll
                  for (int q = 0; q < QMAX; q++)
                    if (q < ndn) v[u][j][q] = Vb[elem_offset<false>(rli[j], rp[j], lc, q, ndn, EM)];
                }
              }
          }
        }
#pragma unroll
        for (int u = 0; u < GBATCH; u++) {
          if (a0 + u < maxdeg) {  // warp-uniform
            if (a0 + u < deg) {
#pragma unroll
              for (int j = 0; j < RMAX; j++)
                if (rok[j]) {
                  const unsigned s = cs[(a0 + u) * nne + rli[j]];
                  if (s != 0xffffu) {
                    // transposed block: v[q] is (row component q, column component rp); selects, not branches (the lanes
                    // of a warp differ)
                    const bool tr = COMPACT && rli[j] > lcs[u];
                    double *dst = acc + (tr ? rp[j] * per_col + (int)s * ndn : (int)s * ndn + rp[j]);
                    const int stride = tr ? 1 : per_col;
#pragma unroll
                    for (int q = 0; q < QMAX; q++)
                      if (q < ndn) dst[q * stride] += v[u][j][q];
                  }
                }
            }
            __syncwarp();  // the next element may add into the same accumulator from another lane
          }
        }
      }
    } else {
      for (int a = 0; a < maxdeg; a++) {
        if (a < deg) {
          const long long ba = base[a];
          const double *Vb = G.V + (ba >> 6);
          const int lc = (int)(ba & 63);
          for (int r = gl; r < EM; r += LPN) {
            const int li = r / ndn, p = r - li * ndn;
            const unsigned s = cs[a * nne + li];
            if (s != 0xffffu) {
              double *dst = acc + s * ndn + p;
#pragma unroll
              for (int q = 0; q < QMAX; q++)
                if (q < ndn) dst[q * per_col] += Vb[elem_offset<COMPACT>(li, p, lc, q, ndn, EM)];
            }
          }
        }
        __syncwarp();
      }
    }
    if (live && nn > 0) {
      const int64_t nb = G.rank ? G.nbrptr[n] : 0;
#pragma unroll
      for (int q = 0; q < QMAX; q++) {
        if (q < ndn) {
          for (int rem = gl; rem < per_col; rem += LPN) {
            const int pos = G.rank ? G.rank[nb * ndn + rem] : rem;
            G.nzval[cb[q] + pos] = acc[q * per_col + rem];
          }
        }
      }
    }
    __syncwarp();
    n = n_next;
    n_next = n_next2;
    pos += groups_total;
  }
}

// Vector assembly (SysvecAssembler, AssemblyModule.jl:853-917) on the node -> element adjacency of the pattern: thread per
// (node, component) adds the node's entry of every adjacent element vector in ascending element order.
__global__ void k_vec_gather(int64_t nnodes, int nne, int ndn, const int64_t *__restrict__ adjptr, const int32_t *__restrict__ adj_slot,
                             const uint8_t *__restrict__ adj_lc, const int32_t *__restrict__ dof, const uint8_t *__restrict__ rowowned,
                             const double *__restrict__ elvec, double *__restrict__ F) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnodes * ndn) return;
  const int64_t n = i / ndn;
  const int p = (int)(i - n * ndn);
  if (rowowned && !rowowned[n]) return;
  const int64_t b = adjptr[n], e = adjptr[n + 1];
  if (b == e) return;
  const int EM = nne * ndn;
  double acc = 0.0;
  for (int64_t a = b; a < e; a++) acc += elvec[(int64_t)adj_slot[a] * EM + adj_lc[a] * ndn + p];
  F[dof[(int64_t)p * nnodes + n]] = acc;
}

// every array of the symbolic phase comes from the context's block cache (fegpu_blockcache.cu)
template <typename T>
int32_t dalloc(fegpu_ctx *ctx, T **p, size_t n) {
  *p = nullptr;
  return fe_dev_alloc(ctx, (void **)p, sizeof(T) * std::max<size_t>(n, 1), ctx->stream);
}

}  // namespace

void fe_pattern_retain(Pattern *p) {
  if (p) p->refs++;
}

void fe_pattern_free(Pattern *p) {
  if (!p) return;
  if (--p->refs > 0) return;  // an assembler result (or the dof map) still reads the arrays
  FE_TRACE("pattern_free: enter");
  // Stream-ordered frees into the context's block cache, no device synchronisation.  They are ordered on the stream the
  // blocks were ALLOCATED on (the build stream), behind the last use on the consumer stream by an event: a rebuild allocates
  // on that same stream again and gets the blocks back by plain stream order.
  cudaStream_t st = p->alloc_stream;
  if (p->stream != p->alloc_stream) {
    if (!p->ready) cudaEventCreateWithFlags(&p->ready, cudaEventDisableTiming);
    if (!p->ready || cudaEventRecord(p->ready, p->stream) != cudaSuccess || cudaStreamWaitEvent(st, p->ready, 0) != cudaSuccess) {
      cudaGetLastError();
      cudaDeviceSynchronize();
    }
  }
  void *ptrs[] = {p->d_colptr, p->d_rowval, p->d_adjptr, p->d_adj_slot, p->d_adj_lc, p->d_nnbr, p->d_nbrptr, p->d_cslot, p->d_rank, p->d_order, p->d_nbr,
                  p->t_deg, p->t_adj, p->t_cs};
  for (void *q : ptrs)
    if (q) fe_dev_free(p->ctx, q, st);
  if (p->ready) cudaEventDestroy(p->ready);
  delete p;
  FE_TRACE("pattern_free: done");
}
void fe_pattern_set_stream(Pattern *p, cudaStream_t s) { p->stream = s; }
cudaEvent_t fe_pattern_ready_event(const Pattern *p) { return p ? p->ready : nullptr; }
int64_t fe_pattern_nnz(const Pattern *p) { return p->nnz; }
bool fe_pattern_is_tile(const Pattern *p) { return p && p->tile; }
const int64_t *fe_pattern_colptr(const Pattern *p) { return p->d_colptr; }
const int64_t *fe_pattern_rowval(const Pattern *p) { return p->d_rowval; }
bool fe_pattern_compressed(const Pattern *p, const int32_t **nbr, const int64_t **nbrptr, int64_t *total_nbr, const int32_t **dof, int *ndn,
                           int64_t *nnodes) {
  if (!p || !p->d_nbr) return false;
  *nbr = p->d_nbr; *nbrptr = p->d_nbrptr; *total_nbr = p->total_nbr; *dof = p->d_dof; *ndn = p->ndn; *nnodes = p->nnodes;
  return true;
}

bool fe_pattern_usable(const fegpu_dofmap *dm) {
  return dm->injective && !dm->mesh->degenerate && dm->row_nall == dm->col_nall && dm->mesh->nne <= 32;
}

int32_t fe_pattern_build(fegpu_dofmap *dm, const std::function<int32_t(bool)> *fork) {
  fegpu_ctx *ctx = dm->ctx;
  fegpu_mesh *mesh = dm->mesh;
  cudaStream_t st = ctx->stream;
  {  // small stencils with a node-major affine dof map: thread-per-node kernels (fegpu_tile.cu)
    bool taken = false;
    FE_TRY(fe_tile_build(dm, fork, &taken));
    if (taken) return FEGPU_OK;
  }
  if (dm->pat) { fe_pattern_free(dm->pat); dm->pat = nullptr; }
  Pattern *P = new Pattern();
  dm->pat = P;  // owned by the dofmap from here on (freed with it, also on error paths)
  P->ctx = ctx;
  P->stream = st;
  P->alloc_stream = st;
  P->ncols = dm->col_nall;
  P->nrows = dm->row_nall;
  const int64_t nn = mesh->nnodes;
  const int nne = mesh->nne, ndn = dm->ndn;
  SymParams S{mesh->conn_act(), mesh->d_elem_list, mesh->nactive, nne, nn, mesh->d_rowowned, dm->d_dof, ndn, nullptr, nn, nullptr};
  const int64_t nadj = mesh->nactive * nne;

  uint16_t *d_arank = nullptr;
  int32_t *d_deg = nullptr, *d_U = nullptr, *d_aflag = nullptr, *d_anodes = nullptr, *d_cls = nullptr;
  unsigned long long *d_clscnt = nullptr;
  int64_t *d_apos = nullptr;
  uint8_t *d_sorted = nullptr;
  // [0] degenerate, [1] some node needs a dof sort, [2] the dof map is not node-major ascending, [3] INT32_MAX - smallest dof and
  // [4] 1 + largest dof of the window's nodes, [5] largest node degree, [6] largest neighbour count
  int *d_flags = nullptr;
  constexpr int NFLAGS = 8;
  auto cleanup = [&]() {
    void *ptrs[] = {d_deg, d_arank, d_U, d_sorted, d_flags, d_aflag, d_anodes, d_apos, d_cls, d_clscnt};
    for (void *q : ptrs)
      if (q) fe_dev_free(ctx, q, st);
  };
  auto bail = [&]() {  // the mesh cannot use the structured path: free everything, the caller takes the sort path
    mesh->degenerate = true;
    cleanup();
    fe_pattern_free(P);
    dm->pat = nullptr;
    return FEGPU_OK;
  };
#define PT(expr) do { int32_t _s = (expr); if (_s != FEGPU_OK) { cleanup(); return _s; } } while (0)
#define PC(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return fegpu_fail(ctx, FEGPU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } } while (0)
  FE_TRACE("build: enter");
  if (fe_trace_on()) {
    int64_t hits = 0, misses = 0;
    size_t fb = 0;
    fe_dev_cache_stats(ctx, &hits, &misses, &fb);
    std::fprintf(stderr, "[fegpu trace] block cache: %lld hits, %lld misses (driver allocations), %.2f GB cached\n", (long long)hits, (long long)misses, fb / 1e9);
  }
  // node window of the active elements: every per-node pass below runs over [lo, hi) only
  const int64_t lo = mesh->win_lo, hi = mesh->win_hi, nw = hi - lo;
  PT(dalloc(ctx, &d_deg, nn));
  PT(dalloc(ctx, &d_arank, nadj));
  PT(dalloc(ctx, &d_flags, NFLAGS));
  if (nw > 0) {
    PC(cudaMemsetAsync(d_deg + lo, 0, sizeof(int32_t) * nw, st));
  }
  PC(cudaMemsetAsync(d_flags, 0, sizeof(int) * NFLAGS, st));
  if (nw > 0) {
    k_dof_monotone<<<(unsigned)std::min<int64_t>(grid_for(nw, 256), (int64_t)ctx->sm_count * 8), 256, 0, st>>>(dm->d_dof, nn, ndn, lo, nw, d_flags + 2, d_flags + 3);
    ctx->launches++;
  }
  if (nadj > 0) {
    k_count_adj<<<grid_for(nadj, 256), 256, 0, st>>>(S, d_deg, d_arank, d_flags);
    ctx->launches++;
  }
  fe_mark(ctx, "sym:k_count_adj");
  PT(dalloc(ctx, &P->d_adjptr, nn + 1));
  PT(fe_exclusive_scan_i32_to_i64(ctx, d_deg + lo, P->d_adjptr + lo, nw, 0, true, nullptr));
  PT(fe_max_i32_dev(ctx, d_deg + lo, nw, d_flags + 5));
  // active nodes (at least one active element): compacted list, dropped again when it is every node
  int64_t na = nn;
  PT(dalloc(ctx, &d_aflag, nw));
  PT(dalloc(ctx, &d_apos, nw + 1));
  PT(dalloc(ctx, &d_anodes, nw));
  if (nw > 0) {
    k_flag_active<<<grid_for(nw, 256), 256, 0, st>>>(d_deg, lo, nw, d_aflag);
    ctx->launches++;
  }
  PT(fe_exclusive_scan_i32_to_i64(ctx, d_aflag, d_apos, nw, 0, true, nullptr));
  if (nw > 0) {
    k_compact_active<<<grid_for(nw, 256), 256, 0, st>>>(d_aflag, d_apos, lo, nw, d_anodes);
    ctx->launches++;
  }
  // one round trip for every scalar the host needs here
  int h_flags[NFLAGS] = {0};
  PC(cudaMemcpyAsync(h_flags, d_flags, sizeof(int) * NFLAGS, cudaMemcpyDeviceToHost, st));
  FE_TRACE("build: first passes queued");
  PC(cudaMemcpyAsync(&na, d_apos + nw, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  PC(cudaStreamSynchronize(st));
  FE_TRACE("build: sync A done");
  fe_mark(ctx, "sym:scans_a");
  if (h_flags[0]) return bail();
  if (na < nn) {
    S.anodes = d_anodes;
    S.na = na;
  }
  int32_t maxdeg = h_flags[5];
  // dof range of the window's nodes: the only part of colptr that is counted and scanned
  const int64_t dlo = h_flags[4] > 0 ? (int64_t)(INT32_MAX - h_flags[3]) : 0;
  const int64_t dhi = h_flags[4] > 0 ? (int64_t)h_flags[4] : 0;  // exclusive
  const bool monotone = h_flags[2] == 0;
  if (maxdeg < 1) maxdeg = 1;
  P->maxdeg = maxdeg;
  P->maxcand = maxdeg * nne;
  int capc = 32;
  while (capc < P->maxcand) capc <<= 1;
  // limits of the packed encodings: neighbour slots and rows per column < 65535
  const size_t smem1 = (size_t)WPB * (maxdeg + 3 * (size_t)capc) * sizeof(uint32_t);
  if (P->maxcand >= 65535 || (int64_t)P->maxcand * ndn >= 65535 || smem1 > 200 * 1024) return bail();
  if (fork) PT((*fork)(false));  // the structured path will be taken: independent work may start on another stream now
  FE_TRACE("build: fork (integration queued)");

  PT(dalloc(ctx, &P->d_adj_slot, nadj));
  PT(dalloc(ctx, &P->d_adj_lc, nadj));
  int KB = 5;
  while ((1 << KB) < capc) KB++;
  static const bool nbr_fast_off = std::getenv("FEGPU_NBR_FAST") && std::atoi(std::getenv("FEGPU_NBR_FAST")) == 0;  // A/B knob
  // register-only neighbour kernels: adjacency in one register per lane, (node, candidate) and (slot, local index) keys in 32 bits
  const bool nbr_fast = !nbr_fast_off && maxdeg <= 32 && capc <= 512 && (uint64_t)nn <= (uint64_t)(0xffffffffu >> KB) &&
                        mesh->nactive < ((int64_t)1 << 27);
  if (nadj > 0) {
    k_fill_adj<<<grid_for(nadj, 256), 256, 0, st>>>(S, P->d_adjptr, d_arank, P->d_adj_slot, P->d_adj_lc);
    ctx->launches++;
    if (!nbr_fast) {  // the register-only kernels sort every node's list themselves
      k_sort_adj<<<grid_for(S.na, 128), 128, 0, st>>>(S, P->d_adjptr, P->d_adj_slot, P->d_adj_lc);
      ctx->launches++;
    }
  }
  fe_mark(ctx, "sym:k_fill_adj");
  PT(dalloc(ctx, &P->d_nnbr, nn));
  PT(dalloc(ctx, &d_U, (size_t)nadj * nne));
  PT(dalloc(ctx, &d_sorted, nn));
  PT(dalloc(ctx, &P->d_cslot, (size_t)nadj * nne));
  if (S.anodes && nw > 0) {  // the per-node kernels skip inactive nodes: their outputs must still be defined (inside the window;
                             // nothing reads these two arrays outside it)
    PC(cudaMemsetAsync(P->d_nnbr + lo, 0, sizeof(int32_t) * nw, st));
    PC(cudaMemsetAsync(d_sorted + lo, 1, nw, st));
  }
  unsigned gridn = (unsigned)std::min<int64_t>((S.na + WPB - 1) / WPB, (int64_t)ctx->sm_count * 64);
  if (gridn == 0) gridn = 1;
  if (nbr_fast) {
    // register-only kernel; the per-node order test (and its shared memory) only when the dof map is not node-major ascending
    const size_t smf = monotone ? 0 : (size_t)WPB * 2 * capc * sizeof(int32_t);
#define LAUNCH_FAST(M)                                                                                                              \
  do {                                                                                                                              \
    if (monotone) k_nbr_fast<M, false><<<gridn, WPB * 32, 0, st>>>(S, P->d_adjptr, P->d_adj_slot, P->d_adj_lc, capc, KB, P->d_nnbr, d_U, P->d_cslot, d_sorted, d_flags + 1); \
    else k_nbr_fast<M, true><<<gridn, WPB * 32, smf, st>>>(S, P->d_adjptr, P->d_adj_slot, P->d_adj_lc, capc, KB, P->d_nnbr, d_U, P->d_cslot, d_sorted, d_flags + 1);       \
  } while (0)
#define LAUNCH_GROUP(L, K)                                                                                                          \
  do {                                                                                                                              \
    const unsigned gg = (unsigned)std::max<int64_t>(1, std::min<int64_t>((S.na + WPB * (32 / L) - 1) / (WPB * (32 / L)), (int64_t)ctx->sm_count * 64)); \
    if (monotone) k_nbr_group<L, K, false><<<gg, WPB * 32, 0, st>>>(S, P->d_adjptr, P->d_adj_slot, P->d_adj_lc, capc, KB, P->d_nnbr, d_U, P->d_cslot, d_sorted, d_flags + 1); \
    else k_nbr_group<L, K, true><<<gg, WPB * 32, smf * (32 / L), st>>>(S, P->d_adjptr, P->d_adj_slot, P->d_adj_lc, capc, KB, P->d_nnbr, d_U, P->d_cslot, d_sorted, d_flags + 1); \
  } while (0)
    static const bool group_off = std::getenv("FEGPU_NBR_GROUP") && std::atoi(std::getenv("FEGPU_NBR_GROUP")) == 0;  // A/B knob
    if (!group_off && capc <= 32 && maxdeg <= 8) LAUNCH_GROUP(8, 4);          // Q4 / T3 skins, T3/Q4 planar: four nodes per warp
    else if (!group_off && capc <= 64 && maxdeg <= 16) LAUNCH_GROUP(16, 4);   // H8: two nodes per warp
    else if (capc <= 64) LAUNCH_FAST(2);
    else {
      // mixed valences: one launch per candidate-count class, each over its own node list
      static const bool classes_off = std::getenv("FEGPU_NBR_CLASSES") && std::atoi(std::getenv("FEGPU_NBR_CLASSES")) == 0;  // A/B knob
      if (classes_off) {
        if (capc <= 128) LAUNCH_FAST(4);
        else LAUNCH_FAST(16);
      } else {
        PT(dalloc(ctx, &d_cls, (size_t)S.na * 3));
        PT(dalloc(ctx, &d_clscnt, 3));
        PC(cudaMemsetAsync(d_clscnt, 0, sizeof(unsigned long long) * 3, st));
        if (nw > 0) {
          PC(cudaMemsetAsync(P->d_nnbr + lo, 0, sizeof(int32_t) * nw, st));  // nodes without elements are in no list
          if (!monotone) PC(cudaMemsetAsync(d_sorted + lo, 1, nw, st));
        }
        k_classify_nodes<<<grid_for(S.na, 256), 256, 0, st>>>(S, d_deg, d_cls, d_cls + S.na, d_cls + 2 * S.na, d_clscnt);
        ctx->launches++;
        const SymParams S_all = S;
        for (int c = 0; c < 3; c++) {
          S.anodes = d_cls + (size_t)c * S_all.na;
          S.na_dev = reinterpret_cast<const int64_t *>(d_clscnt + c);
          if (c == 0) {  // <= 64 candidates: deg <= 16 for nne >= 4, so two nodes share a warp
            if (!group_off && nne >= 4) LAUNCH_GROUP(16, 4);
            else LAUNCH_FAST(2);
          } else if (c == 1) LAUNCH_FAST(4);
          else if (capc > 128) LAUNCH_FAST(16);
          ctx->launches++;
        }
        S = S_all;
        ctx->launches--;  // the common increment below counts one of them
      }
    }
#undef LAUNCH_GROUP
#undef LAUNCH_FAST
    if (monotone && nw > 0) PC(cudaMemsetAsync(d_sorted + lo, 1, nw, st));  // every node is in order; the kernel did not write the flags
  } else {
#define LAUNCH_NBR(N)                                                                                                  \
  do {                                                                                                                 \
    PC(cudaFuncSetAttribute(k_nbr<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));                      \
    k_nbr<N><<<gridn, WPB * 32, smem1, st>>>(S, P->d_adjptr, P->d_adj_slot, maxdeg, capc, P->d_nnbr, d_U, P->d_cslot, d_sorted, d_flags + 1); \
  } while (0)
    switch (nne) {
      case 3: LAUNCH_NBR(3); break;
      case 4: LAUNCH_NBR(4); break;
      case 8: LAUNCH_NBR(8); break;
      case 10: LAUNCH_NBR(10); break;
      case 20: LAUNCH_NBR(20); break;
      case 27: LAUNCH_NBR(27); break;
      default: LAUNCH_NBR(0); break;
    }
#undef LAUNCH_NBR
  }
  ctx->launches++;
  fe_mark(ctx, "sym:k_nbr");
  FE_TRACE("build: nbr kernel queued");
  PT(dalloc(ctx, &P->d_nbrptr, nn + 1));
  int64_t total_nbr = 0;
  PT(fe_exclusive_scan_i32_to_i64(ctx, P->d_nnbr + lo, P->d_nbrptr + lo, nw, 0, true, nullptr));
  PT(fe_max_i32_dev(ctx, P->d_nnbr + lo, nw, d_flags + 6));
  if (nw < nn) {  // adjptr / nbrptr stay complete prefix arrays for the consumers that walk every node (vector gather, transport)
    k_fill_outside<<<grid_for(nn - nw, 256), 256, 0, st>>>(P->d_adjptr, P->d_nbrptr, nn + 1, lo, hi, 0);
    ctx->launches++;
  }
  // column pointers: counts and scan over the dof range of the window, constants on both sides of it
  const int64_t ndw = dhi - dlo;
  PT(dalloc(ctx, &P->d_colptr, P->ncols + 1));
  if (ndw > 0) PC(cudaMemsetAsync(P->d_colptr + dlo, 0, sizeof(int64_t) * ndw, st));
  if (nw * ndn > 0) {
    k_col_counts<<<grid_for(nw * ndn, 256), 256, 0, st>>>(S, P->d_nnbr, lo, nw, P->d_colptr);
    ctx->launches++;
  }
  int64_t tot = 0;
  PT(fe_exclusive_scan_i64(ctx, P->d_colptr + dlo, P->d_colptr + dlo, ndw, 1, true, nullptr));
  if (ndw < P->ncols) {
    k_fill_outside<<<grid_for(P->ncols - ndw, 256), 256, 0, st>>>(P->d_colptr, nullptr, P->ncols + 1, dlo, dhi, 1);
    ctx->launches++;
  }
  PC(cudaMemcpyAsync(h_flags, d_flags, sizeof(int) * NFLAGS, cudaMemcpyDeviceToHost, st));
  PC(cudaMemcpyAsync(&total_nbr, P->d_nbrptr + hi, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  FE_TRACE("build: nbr + scans queued");
  PC(cudaMemcpyAsync(&tot, P->d_colptr + dhi, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  PC(cudaStreamSynchronize(st));
  FE_TRACE("build: sync B done");
  fe_mark(ctx, "sym:scans_b");
  P->maxnbr = std::max(h_flags[6], 1);
  P->nnz = tot - 1;
  if (P->nnz != total_nbr * ndn * ndn) {
    cleanup();
    return fegpu_fail(ctx, FEGPU_ERR_STATE, "internal: pattern size mismatch");
  }
  const bool need_rank = h_flags[1] != 0;
  PT(dalloc(ctx, &P->d_rowval, (size_t)P->nnz));
  if (need_rank) PT(dalloc(ctx, &P->d_rank, (size_t)(total_nbr * ndn)));
  P->total_nbr = total_nbr; P->d_dof = dm->d_dof; P->ndn = ndn; P->nnodes = nn;
  // neighbour lists kept for the transport: only where they are smaller than int32 row indices (ndn^2 columns-rows per pair)
  if (!need_rank && ndn >= 2 && total_nbr > 0) PT(dalloc(ctx, &P->d_nbr, (size_t)total_nbr));
  int cap2 = 32;
  while (cap2 < P->maxnbr * ndn) cap2 <<= 1;
  const size_t smem2 = need_rank ? (size_t)WPB * cap2 * sizeof(unsigned long long) : 0;
  if (smem2 > 200 * 1024) return bail();
  if (smem2 > 48 * 1024) PC(cudaFuncSetAttribute(k_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  {
    // lanes per node: a warp for vector fields (rows per column >= 81 for H8), 8 lanes for scalar fields (27 rows: four
    // nodes per warp keep four dependent-load chains in flight).  FEGPU_ROWS_LPN overrides (tuning knob).
    static const int lpn_env = std::getenv("FEGPU_ROWS_LPN") ? std::atoi(std::getenv("FEGPU_ROWS_LPN")) : 0;
    int rl = (ndn >= 3) ? 32 : (ndn == 2 ? 16 : 8);
    if (lpn_env == 8 || lpn_env == 16 || lpn_env == 32) rl = lpn_env;
    const unsigned gr = std::max(1u, grid_for(S.na * rl, 256));  // S.na == 0: an empty FESet, or a rank without nodes
#define ROWS_ARGS S, P->d_adjptr, P->d_nnbr, P->d_nbrptr, d_U, d_sorted, P->d_colptr, P->d_rowval, P->d_rank, P->d_nbr
#define ROWS_L(N)                                                              \
  do {                                                                         \
    if (rl == 8) k_rows_sorted<8, N><<<gr, 256, 0, st>>>(ROWS_ARGS);           \
    else if (rl == 16) k_rows_sorted<16, N><<<gr, 256, 0, st>>>(ROWS_ARGS);    \
    else k_rows_sorted<32, N><<<gr, 256, 0, st>>>(ROWS_ARGS);                  \
  } while (0)
    switch (ndn) {
      case 1: ROWS_L(1); break;
      case 2: ROWS_L(2); break;
      case 3: ROWS_L(3); break;
      default: ROWS_L(0); break;
    }
#undef ROWS_L
#undef ROWS_ARGS
    ctx->launches++;
  }
  if (need_rank) {  // only the nodes whose dof order needs a sort
    k_rows<<<gridn, WPB * 32, smem2, st>>>(S, P->d_adjptr, P->d_nnbr, P->d_nbrptr, d_U, d_sorted, P->d_colptr, cap2, P->d_rowval, P->d_rank);
    ctx->launches++;
  }
  PC(cudaGetLastError());
  fe_mark(ctx, "sym:k_rows");
  {  // node visiting order of the gather.  Morton order (FEGPU_GATHER_ORDER=1) was measured on configs 2-4: it costs a
     // radix sort of the nodes per pattern build (+0.2 ms at 2.1 M nodes, +3 ms at 17 M) and saves < 0.1 ms of gather on
     // meshes whose numbering is already local, so the default is the natural order of the active nodes.
    static const bool order_off = !(std::getenv("FEGPU_GATHER_ORDER") && std::atoi(std::getenv("FEGPU_GATHER_ORDER")) == 1);
    if (!order_off && S.na > 0) {
      PT(dalloc(ctx, &P->d_order, S.na));
      PT(fe_morton_order(mesh, S.anodes, S.na, P->d_order));
      P->norder = S.na;
    } else if (S.anodes && S.na > 0) {  // natural order, active nodes only
      PT(dalloc(ctx, &P->d_order, S.na));
      PC(cudaMemcpyAsync(P->d_order, S.anodes, sizeof(int32_t) * S.na, cudaMemcpyDeviceToDevice, st));
      P->norder = S.na;
    }
  }
  PC(cudaEventCreateWithFlags(&P->ready, cudaEventDisableTiming));
  PC(cudaEventRecord(P->ready, st));
  fe_mark(ctx, "sym:finish");
  FE_TRACE("build: rows queued");
  cleanup();
  FE_TRACE("build: temporaries freed");
#undef PT
#undef PC
  dm->pat_topo_version = mesh->topo_version;
  return FEGPU_OK;
}

int32_t fe_gather(fegpu_dofmap *dm, const double *d_V, bool compact, double *d_nzval, bool planes, int64_t vstride) {
  fegpu_ctx *ctx = dm->ctx;
  Pattern *P = dm->pat;
  fegpu_mesh *mesh = dm->mesh;
  if (!P) return fegpu_fail(ctx, FEGPU_ERR_STATE, "no pattern");
  if (P->nnz == 0) return FEGPU_OK;
  if (P->tile) return fe_tile_gather(dm, d_V, compact, planes, vstride, d_nzval);
  if (planes) return fegpu_fail(ctx, FEGPU_ERR_STATE, "internal: plane layout without a thread-per-node pattern");
  GatherParams G{mesh->nnodes, mesh->nne, dm->ndn, P->d_adjptr, P->d_adj_slot, P->d_adj_lc, P->d_nnbr, P->d_nbrptr, P->d_cslot,
                 P->d_rank, dm->d_dof, P->d_colptr, d_V, d_nzval, P->maxnbr, P->maxdeg, P->maxcand, P->d_order, P->d_order ? P->norder : mesh->nnodes};
  const int EM = mesh->nne * dm->ndn;
  // tuning knobs (measured defaults below): FEGPU_GATHER_LPN lanes per node, FEGPU_GATHER_BATCH elements in flight
  static const int batch_env = std::getenv("FEGPU_GATHER_BATCH") ? std::atoi(std::getenv("FEGPU_GATHER_BATCH")) : 0;
  static const int lpn_env = std::getenv("FEGPU_GATHER_LPN") ? std::atoi(std::getenv("FEGPU_GATHER_LPN")) : 0;
  // measured on config 2 (profiles/r01_gather_knobs.txt): two rows per lane with one element in flight beats one row per lane
  // for the compact layout; one row per lane keeps four elements in flight
  int lpn = (EM <= 16) ? 8 : (EM <= 32 ? 16 : 32);
  if (lpn_env == 8 || lpn_env == 16 || lpn_env == 32) lpn = lpn_env;
  if (dm->ndn > 3) lpn = 32;  // runtime-ndn instantiation exists for 32 lanes only
  int rpl = (EM + lpn - 1) / lpn;
  if (rpl > 4) rpl = 0;  // runtime row loop
  const int batch = (rpl == 1) ? ((batch_env == 1 || batch_env == 2) ? 1 : 4) : 1;
  const int npw = 32 / lpn;
  const size_t smem = (size_t)GWPB * npw * ((size_t)P->maxnbr * dm->ndn * dm->ndn + P->maxdeg + (P->maxcand + 3) / 4) * sizeof(double);
  if (smem > 200 * 1024) return fegpu_fail(ctx, FEGPU_ERR_STATE, "internal: gather accumulators exceed shared memory");
  const int64_t per_block = (int64_t)GWPB * npw;
  // persistent launch: exactly the CTAs that are resident at once (occupancy x SMs), so there is no partial last wave
  static const int waves_env = std::getenv("FEGPU_GATHER_WAVES") ? std::atoi(std::getenv("FEGPU_GATHER_WAVES")) : 0;
  static const int gen_carveout = std::getenv("FEGPU_GATHER_GEN_CARVEOUT") ? std::atoi(std::getenv("FEGPU_GATHER_GEN_CARVEOUT")) : 85;  // measured: H20 96^3 gather 12.7 -> 10.9 ms at 85 % (and at 70 %), T10 unchanged; -1 = the driver's choice
  const int64_t need_blocks = std::max<int64_t>(1, (G.npos + per_block - 1) / per_block);
  unsigned grid = 1;
#define LG5(L, N, C, B, R)                                                                                                          \
  do {                                                                                                                              \
    if (smem > 48 * 1024) CUDA_TRY(ctx, cudaFuncSetAttribute(k_gather<L, N, C, B, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    int occ = 0;                                                                                                                    \
    CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_gather<L, N, C, B, R>, GWPB * 32, smem));                   \
    if (gen_carveout > 0 && smem + 1024 <= (size_t)gen_carveout * 228 * 1024 / 100) {                                               \
      /* leave L1 for the lines in flight (see fe_tile_gather): fewer resident CTAs, sized from the carve-out */                    \
      CUDA_TRY(ctx, cudaFuncSetAttribute(k_gather<L, N, C, B, R>, cudaFuncAttributePreferredSharedMemoryCarveout, gen_carveout));   \
      occ = std::min<int>(occ, (int)(((size_t)gen_carveout * 228 * 1024 / 100) / (smem + 1024)));                                   \
    }                                                                                                                               \
    grid = (unsigned)std::min<int64_t>(need_blocks, (int64_t)ctx->sm_count * std::max(occ, 1) * (waves_env > 0 ? waves_env : 1)); \
    k_gather<L, N, C, B, R><<<grid, GWPB * 32, smem, ctx->stream>>>(G);                                                             \
  } while (0)
#define LG_R(L, N, C)                                                            \
  do {                                                                           \
    if (rpl == 1) { if (batch == 4) LG5(L, N, C, 4, 1); else LG5(L, N, C, 1, 1); } \
    else if (rpl == 2) LG5(L, N, C, 1, 2);                                       \
    else if (rpl == 3) LG5(L, N, C, 1, 3);                                       \
    else if (rpl == 4) LG5(L, N, C, 1, 4);                                       \
    else LG5(L, N, C, 1, 0);                                                     \
  } while (0)
#define LG_C(L, N)                     \
  do {                                 \
    if (compact) LG_R(L, N, true);     \
    else LG_R(L, N, false);            \
  } while (0)
#define LG_L(N)                        \
  do {                                 \
    if (lpn == 8) LG_C(8, N);          \
    else if (lpn == 16) LG_C(16, N);   \
    else LG_C(32, N);                  \
  } while (0)
  switch (dm->ndn) {
    case 1: LG_L(1); break;
    case 2: LG_L(2); break;
    case 3: LG_L(3); break;
    default: LG_C(32, 0); break;
  }
#undef LG5
#undef LG_R
#undef LG_C
#undef LG_L
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return FEGPU_OK;
}

int32_t fe_vec_gather(fegpu_dofmap *dm, const double *d_elvec, double *d_F) {
  fegpu_ctx *ctx = dm->ctx;
  Pattern *P = dm->pat;
  fegpu_mesh *mesh = dm->mesh;
  if (!P) return fegpu_fail(ctx, FEGPU_ERR_STATE, "no pattern");
  CUDA_TRY(ctx, cudaMemsetAsync(d_F, 0, sizeof(double) * (size_t)std::max<int64_t>(dm->row_nall, 1), ctx->stream));
  if (P->tile) return fe_tile_vec_gather(dm, d_elvec, d_F);
  const int64_t n = mesh->nnodes * dm->ndn;
  if (n == 0) return FEGPU_OK;
  k_vec_gather<<<grid_for(n, 256), 256, 0, ctx->stream>>>(mesh->nnodes, mesh->nne, dm->ndn, P->d_adjptr, P->d_adj_slot, P->d_adj_lc, dm->d_dof,
                                                         mesh->d_rowowned, d_elvec, d_F);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return FEGPU_OK;
}
