// Internal declarations shared by the translation units of libfinegpu.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>
#include <functional>
#include <string>
#include <vector>

#include "../../include/fegpu.h"

#define FEGPU_MAX_NNE 27
#define FEGPU_MAX_NPTS 64
#define FEGPU_TAB_DOUBLES 6144  // N + dN tables: npts*nne*(1+mdim) <= this

struct fegpu_ctx {
  int device = 0;
  cudaStream_t stream = 0;
  // fresh assemblies run the element integration on a second stream while the symbolic phase (pattern build) runs on
  // `stream`: the two are independent until the numeric phase (FP64-bound vs. integer/latency-bound kernels)
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_fork = nullptr;
  bool overlap = true;  // fegpu_set_overlap / FEGPU_OVERLAP=0: strictly serial phases (per-kernel timing)
  bool async = false;
  int64_t launches = 0;
  int sm_count = 148;
  std::string err;
  struct Transfer *xfer = nullptr;  // staging ring + host threads of the result transport (fegpu_transfer.cu), lazily built
  struct BlockCache *blocks = nullptr;  // device-memory block cache of the symbolic phase (fegpu_blockcache.cu), lazily built
  // named timing marks on the context's stream (fegpu_marks_begin / fegpu_marks_read): per-kernel times of one serial step
  bool marks_on = false;
  struct Mark {
    const char *name;
    cudaEvent_t ev;
  };
  std::vector<Mark> marks;
  int nmarks = 0;
  // quadrature tables / coefficients of the specialised H8 kernels live in __constant__ memory, which every context of a
  // device shares: the owner of the current contents (fegpu_h8.cu re-uploads when another context or other data come along)
  uint64_t const_epoch = 0;
};
void fe_mark(fegpu_ctx *ctx, const char *name);  // records an event named `name` on ctx->stream when marks are on

struct Pattern;  // fegpu_pattern.cu

struct fegpu_mesh {
  fegpu_ctx *ctx = nullptr;
  int etype = 0, nne = 0, mdim = 0, sdim = 0;
  int64_t nelem = 0, nnodes = 0;
  int32_t *d_conn = nullptr;  // [nelem][nne] 0-based, in INTERNAL element order (ascending smallest node id, fe_order_elements)
  int32_t *d_orig = nullptr;  // [nelem] the caller's element id of internal element i (nullptr = identity)
  double *d_xyz = nullptr;    // [sdim][nnodes]
  double rm[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // constant material coordinate system matrix, sdim x mdim col-major (fegpu_csys_set)
  bool use_rm = false;        // false = identity (the FEMMBase default, FEMMBaseModule.jl:82-84)
  double otherdim = 1.0;      // constant other-dimension of the IntegDomain (thickness of a 2-manifold, fegpu_otherdimension_set)
  double *d_uvel = nullptr;   // [sdim][nnodes] nodal field of bilform_convection (fegpu_bilform_convection uploads it)
  // quadrature tables (device copy): N [npts][nne], dN [npts][mdim][nne], w [npts]
  int npts = 0;
  double *d_tab = nullptr;  // N then dN
  double *d_w = nullptr;
  std::vector<double> h_tab, h_w;
  // partition (multi-GPU row blocks)
  bool partitioned = false;
  int64_t nactive = 0;              // == nelem when not partitioned
  int32_t *d_elem_list = nullptr;   // active (internal) element ids ascending; nullptr = the contiguous range elem_base .. elem_base + nactive
  int64_t elem_base = 0;            // (slab partitions of a mesh in internal order activate a contiguous range: no indirection then)
  const int32_t *conn_act() const { return d_conn + elem_base * nne; }  // connectivity row of active slot 0 when d_elem_list is null
  uint8_t *d_rowowned = nullptr;    // per node, nullptr = all owned
  bool own_contig = false;          // the owned nodes are exactly the range [own_lo, own_hi) (slab / reordered partitions)
  int64_t own_lo = 0, own_hi = 0;
  int64_t win_lo = 0, win_hi = 0;   // node window [lo, hi) that contains every node of an active element ([0, nnodes) when not partitioned)
  uint64_t topo_version = 1;        // bumped when the active set / ownership changes
  uint64_t adj_collide_version = 0; // topo_version for which the atomic-free adjacency placement lost an entry (fegpu_tile.cu)
  bool degenerate = false;          // some element lists a node twice -> generic sort path
  double bbox_lo[3] = {0, 0, 0}, bbox_hi[3] = {0, 0, 0};  // of the coordinates at upload (node visiting order of the gather)
};

struct fegpu_dofmap {
  fegpu_ctx *ctx = nullptr;
  fegpu_mesh *mesh = nullptr;
  int ndn = 0;
  int64_t row_nall = 0, col_nall = 0;
  int32_t *d_dof = nullptr;  // [ndn][nnodes] 0-based
  bool injective = true;
  Pattern *pat = nullptr;
  uint64_t pat_topo_version = 0;
  uint64_t tile_failed_version = 0;  // mesh->topo_version for which the thread-per-node path's preconditions failed (fegpu_tile.cu)
};

struct fegpu_asm {
  fegpu_ctx *ctx = nullptr;
  // element-matrix values of the last bilform call: [nactive][EM*EM], reference emission order
  double *d_V = nullptr;
  size_t V_cap = 0;  // doubles
  int64_t V_n = 0;
  int last_EM = 0;
  bool V_compact = false;  // d_V holds the compact symmetric layout (fe_compact_size per element)
  bool V_planes = false;   // d_V is in plane form: value k of slot s at d_V[k * V_stride + s]
  int64_t V_stride = 0;
  // result
  int64_t nrows = 0, ncols = 0, nnz = 0;
  const int64_t *d_colptr = nullptr;  // borrowed from a Pattern or == own_colptr
  const int64_t *d_rowval = nullptr;
  Pattern *pat_src = nullptr;         // the pattern d_colptr/d_rowval are borrowed from, retained (nullptr: the assembler's own arrays)
  double *d_nzval = nullptr;
  size_t nz_cap = 0;
  int64_t *own_colptr = nullptr, *own_rowval = nullptr;
  size_t own_colptr_cap = 0, own_rowval_cap = 0;
  bool have_result = false;
  bool pattern_cached = false;
  // assembled vector (SysvecAssembler semantics: linform_dot / distribloads and the generic vector protocol)
  double *d_F = nullptr;
  size_t F_cap = 0;
  int64_t F_n = 0;
  bool have_vector = false, vec_started = false;
  int64_t v_row_nall = 0;
  std::vector<int64_t> hvI;
  std::vector<double> hvV;
  // optional view of the result (sub-block and/or exact zeros dropped, fegpu_csc_ops.cu); the accessors below pick it
  struct View {
    bool active = false;
    int64_t nrows = 0, ncols = 0, nnz = 0;
    int64_t *own_colptr = nullptr, *own_rowval = nullptr;
    double *own_nzval = nullptr;
    size_t colptr_cap = 0, rowval_cap = 0, nzval_cap = 0;
  } view;
  int64_t r_nrows() const { return view.active ? view.nrows : nrows; }
  int64_t r_ncols() const { return view.active ? view.ncols : ncols; }
  int64_t r_nnz() const { return view.active ? view.nnz : nnz; }
  const int64_t *r_colptr() const { return view.active ? view.own_colptr : d_colptr; }
  const int64_t *r_rowval() const { return view.active ? view.own_rowval : d_rowval; }
  const double *r_nzval() const { return view.active ? view.own_nzval : d_nzval; }
  // generic protocol staging (host side, flushed to the device at makematrix)
  bool started = false;
  bool symmetric = false;  // SysmatAssemblerSparseSymm semantics (fegpu_asm_set_symmetric)
  int lump = 0;            // 0 none, 1 SysmatAssemblerSparseDiag, 2 SysmatAssemblerSparseHRZLumpingSymm (fegpu_asm_set_lumping)
  std::vector<int32_t> hN; // generic protocol with lumping: size of every staged square element matrix
  int64_t g_row_nall = 0, g_col_nall = 0;
  std::vector<int64_t> hI, hJ;
  std::vector<double> hV;
  // timings
  // [0] start, [1] symbolic done, [2] integration done, [3] end, [4] integration start, [5] numeric start
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool ev_valid = false;
};

// ---- error helpers -----------------------------------------------------------------------------------
int32_t fegpu_fail(fegpu_ctx *ctx, int32_t code, const std::string &msg);
#define CUDA_TRY(ctx, expr)                                                                              \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess)                                                                               \
      return fegpu_fail((ctx), FEGPU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));      \
  } while (0)
#define FE_TRY(expr)                \
  do {                              \
    int32_t _s = (expr);            \
    if (_s != FEGPU_OK) return _s;  \
  } while (0)

// FEGPU_TRACE=1: host wall-clock marks (microseconds since the previous mark) on stderr, to find host-side stalls
void fe_trace(const char *label);
bool fe_trace_on();
#define FE_TRACE(label) fe_trace(label)

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

// ---- device-memory block cache (fegpu_blockcache.cu): stream-ordered alloc / free without driver calls on the hot path
int32_t fe_dev_alloc(fegpu_ctx *ctx, void **p, size_t bytes, cudaStream_t stream);
void fe_dev_free(fegpu_ctx *ctx, void *p, cudaStream_t stream);  // p may be reused by work queued on `stream` from now on
void fe_dev_cache_stats(fegpu_ctx *ctx, int64_t *hits, int64_t *misses, size_t *free_bytes);
void fe_dev_cache_trim(fegpu_ctx *ctx);     // cudaFree every cached (free) block
void fe_dev_cache_destroy(fegpu_ctx *ctx);

// ---- primitives (fegpu_prims.cu) -----------------------------------------------------------------------
// out[i] = sum_{k<i} in[k] (+ base); out may alias in for the int64 version.  n+1 entries are written when
// write_total is true (out[n] = total).  Returns the total through *total_host if non-null (synchronises).
int32_t fe_exclusive_scan_i64(fegpu_ctx *ctx, const int64_t *d_in, int64_t *d_out, int64_t n, int64_t base, bool write_total,
                              int64_t *total_host);
int32_t fe_exclusive_scan_i32_to_i64(fegpu_ctx *ctx, const int32_t *d_in, int64_t *d_out, int64_t n, int64_t base, bool write_total,
                                     int64_t *total_host);
int32_t fe_max_i32(fegpu_ctx *ctx, const int32_t *d_in, int64_t n, int32_t *max_host);
int32_t fe_max_i32_dev(fegpu_ctx *ctx, const int32_t *d_in, int64_t n, int32_t *d_out);

// ---- integration (fegpu_integrate.cu) ------------------------------------------------------------------
enum { FORM_DIFF_ISO = 0, FORM_DIFF_GEN = 1, FORM_ELASTIC = 2, FORM_DOT = 3, FORM_CONVECTION = 4, FORM_DIV_GRAD = 5,
       FORM_LINDOT = 6 /* linform_dot: element VECTORS, [nactive][EM] */,
       FORM_MASSLIKE = 7 /* bilform_masslike: rectangular ndn x EM element matrices, [nactive][ndn*EM] column-major */ };
struct FormArgs {
  int form;
  int ndn;
  double coef[36];  // kappa (mdim x mdim col-major) | C (6x6 col-major) | c (ndn x ndn col-major); coef[0] = scalar kappa
  int m;            // bilform_dot manifold dimension
  double otherdim;
  double rm[9];     // constant CSys matrix (sdim x mdim col-major) for the general diffusion and the elasticity forms
  bool use_rm;      // false = identity
  const double *d_uvel = nullptr;  // bilform_convection: nodal convective velocity on the device, [sdim][nnodes]
  bool compact;     // symmetric forms only: write the compact upper-block layout (fe_compact_size) instead of full matrices
  // struct-of-arrays output for the thread-per-node numeric kernel: value k of the element in slot s at V[k * vstride + s]
  // (k = the position inside the compact / full element record).  Only the kernels of fe_integrate_supports_planes.
  bool planes = false;
  int64_t vstride = 0;
};
// Does the 6 x 6 material matrix (column-major, strain order xx,yy,zz,xy,xz,yz) have the cubic-symmetry form
//   [D00 on the normal diagonal, lam off it] (+) mu I_3, every other entry exactly zero
// (isotropic materials: MatDeforElastIso builds exactly this)?  Then out = {D00, lam, mu} and the elasticity kernels may use their
// outer-product formulation (fegpu_h8.cu: outer_acc / iso_block).  FEGPU_ELASTIC_ISO=0 switches the shortcut off (A/B knob).
static inline bool fe_elastic_cubic(const double *D, double out[3]) {
  static const bool off = std::getenv("FEGPU_ELASTIC_ISO") && std::atoi(std::getenv("FEGPU_ELASTIC_ISO")) == 0;
  if (off) return false;
  const double d00 = D[0], lam = D[1], mu = D[3 + 6 * 3];
  for (int j = 0; j < 6; j++)
    for (int i = 0; i < 6; i++) {
      const double v = D[i + 6 * j];
      const double want = (i == j) ? (i < 3 ? d00 : mu) : ((i < 3 && j < 3) ? lam : 0.0);
      if (!(v == want)) return false;
    }
  out[0] = d00; out[1] = lam; out[2] = mu;
  return true;
}

// Compact layout of a symmetric element matrix (nne nodes x ndn dofs): the upper block triangle, block (a <= b) of
// ndn x ndn values (column-major: row comp i, col comp j at j*ndn + i) at ndn*ndn*(b(b+1)/2 + a); diagonal blocks are stored
// in full (mirrored).  This is what the symmetric forms write on the mesh-structured path: (1 + 1/nne)/2 of the bytes.
static inline int fe_compact_size(int nne, int ndn) { return nne * (nne + 1) / 2 * ndn * ndn; }
// forms whose kernels compute the upper triangle once and mirror it (they may write the compact layout); bilform_div_grad's
// element matrices are symmetric too but the reference forms them in full (FEMMBaseModule.jl:1694-1707), and so do we
static inline bool fe_form_symmetric(int form) { return form <= 2; }
static inline bool fe_form_values_symmetric(int form) { return form <= 2 || form == 5; }
int32_t fe_integrate(fegpu_mesh *mesh, const FormArgs &fa, double *d_V);
// does the integration kernel that will run write the compact symmetric layout when fa.compact is set?
bool fe_integrate_supports_compact(const fegpu_mesh *mesh, const FormArgs &fa);
bool fe_integrate_supports_planes(const fegpu_mesh *mesh, const FormArgs &fa);  // ... the plane layout when fa.planes is set?

// ---- pattern + gather (fegpu_pattern.cu) ---------------------------------------------------------------
// `fork` (optional) is invoked once, on the calling thread, as soon as the build knows it will not fall back to the sort path
// for an early reason (degenerate elements, encoding limits): the caller launches independent work on another stream there
// fork(tile): tile = the thread-per-node path is being taken (the integration may write the plane layout); it may be invoked a
// second time with tile = false when that path's preconditions turn out violated
int32_t fe_pattern_build(fegpu_dofmap *dm, const std::function<int32_t(bool)> *fork = nullptr);
// Patterns are shared: the dof map that built one and every assembler whose result borrows its colptr / rowval hold a
// reference (an invalidated or rebuilt pattern must not pull the arrays from under a result that is still being read).
void fe_pattern_retain(Pattern *p);
void fe_pattern_free(Pattern *p);  // drops one reference; the arrays go back to the block cache with the last one
// thread-per-node kernels for small stencils (fegpu_tile.cu); *taken = false: preconditions not met, run the general path
int32_t fe_tile_build(fegpu_dofmap *dm, const std::function<int32_t(bool)> *fork, bool *taken);
bool fe_tile_candidate(const fegpu_dofmap *dm);  // would fe_tile_build try the thread-per-node path for this dof map?
bool fe_pattern_is_tile(const Pattern *p);
int32_t fe_tile_gather(fegpu_dofmap *dm, const double *d_V, bool compact, bool planes, int64_t vstride, double *d_nzval);
int32_t fe_tile_vec_gather(fegpu_dofmap *dm, const double *d_elvec, double *d_F);
void fe_pattern_set_stream(Pattern *p, cudaStream_t s);  // stream its stream-ordered frees are queued on
cudaEvent_t fe_pattern_ready_event(const Pattern *p);    // completes when every array of the pattern is final
int64_t fe_pattern_nnz(const Pattern *p);
const int64_t *fe_pattern_colptr(const Pattern *p);
const int64_t *fe_pattern_rowval(const Pattern *p);
// compressed row structure for the transport (false when the pattern keeps none): per-node ascending neighbour lists + dof map
// column-stencil codec of the result transport (fegpu_csc_ops.cu): one id per column + a dictionary of row-offset lists
int32_t fe_col_stencils(fegpu_ctx *ctx, int64_t ncols, const int64_t *d_colptr, const int64_t *d_rowval, cudaStream_t stream, uint32_t **d_ids,
                        int32_t **d_dict, int *ndict, int *maxlen, int *cap, int64_t *col_first, int64_t *col_last, bool *ok);
bool fe_pattern_compressed(const Pattern *p, const int32_t **nbr, const int64_t **nbrptr, int64_t *total_nbr, const int32_t **dof, int *ndn,
                           int64_t *nnodes);
bool fe_pattern_usable(const fegpu_dofmap *dm);  // mesh-structured fast path applicable?
int32_t fe_gather(fegpu_dofmap *dm, const double *d_V, bool compact, double *d_nzval, bool planes = false, int64_t vstride = 0);
// element vectors [nactive][nne*ndn] -> dense vector F[row_nall] (zeroed here): every node sums its adjacent elements'
// entries in ascending element order (= the reference's element loop); rows of nodes this rank does not own stay zero
int32_t fe_vec_gather(fegpu_dofmap *dm, const double *d_elvec, double *d_F);
// compact symmetric layout -> full element matrices in emission order (raw-COO export only)
int32_t fe_expand_compact(fegpu_ctx *ctx, const double *d_Vc, double *d_Vfull, int64_t nelem, int nne, int ndn, const int32_t *d_perm = nullptr,
                          int64_t vstride = 0 /* > 0: d_Vc is in plane form */);

// ---- generic COO -> CSC by sort (fegpu_sort.cu) --------------------------------------------------------
// d_I, d_J 1-based int64, n triplets in emission order.  Fills the assembler's own colptr/rowval/nzval.
int32_t fe_coo_to_csc(fegpu_asm *as, int64_t n, const int64_t *d_I, const int64_t *d_J, const double *d_V, int64_t nrows,
                      int64_t ncols);
// emit the reference-order (I, J) of a bilform assembly (AssemblyModule.jl:266-279) from conn + dof map
int32_t fe_emit_ij(fegpu_dofmap *dm, int64_t *d_I, int64_t *d_J, const int32_t *d_perm = nullptr);
// internal element order (ascending smallest node id) at upload; emission-order permutation for the raw-COO export
int32_t fe_order_elements(fegpu_mesh *mesh);
int32_t fe_emission_order(fegpu_mesh *mesh, int32_t *d_perm /* [nactive]: rank in the caller's order -> slot */);
// Morton order of (a subset of) the nodes (spatial locality for the gather's L2 reuse): d_order [n]
int32_t fe_morton_order(fegpu_mesh *mesh, const int32_t *d_nodes /* subset or nullptr = all */, int64_t n, int32_t *d_order);

// ---- result transport (fegpu_transfer.cu) --------------------------------------------------------------
// device CSC -> caller's host arrays; rowval crosses the link as int32 and is widened by host threads
int32_t fe_copy_result(fegpu_asm *as, int64_t *colptr, int64_t *rowval, double *nzval);
void fe_transfer_free(struct Transfer *t);
// ---- views of the assembled CSC (fegpu_csc_ops.cu): A[r0:r1, c0:c1] (1-based, inclusive), optionally without exact zeros
int32_t fe_csc_view(fegpu_asm *as, int64_t r0, int64_t r1, int64_t c0, int64_t c1, bool drop_zeros);

int32_t fe_asm_reserve(fegpu_asm *as, double **buf, size_t *cap, size_t need_doubles);
int32_t fe_reserve_bytes(fegpu_ctx *ctx, void **buf, size_t *cap, size_t need_bytes);
