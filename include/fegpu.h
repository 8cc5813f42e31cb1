/*
 * fegpu.h -- C ABI of libfinegpu.so: B200 (sm_100a) element integration + sparse assembly behind
 * FinEtools.jl's assembler protocol.
 *
 * FinEtools.jl is pure Julia and has no FFI on this path; the entry points below are what a
 * `SysmatAssemblerSparseGPU <: AbstractSysmatAssembler` shim binds with `ccall` (see INTEGRATION.md and
 * finetools.jl_b200/julia/FinEtoolsGPU.jl).  Each one names the reference interface (file:line under
 * /root/reference/src) it replaces.
 *
 * Conventions
 *  - every function returns int32 status: 0 = OK, <0 = error; fegpu_last_error() gives the message.  The
 *    messages for dof-range violations are the reference's own strings (AssemblyModule.jl:265-273).
 *  - all index arrays crossing the boundary are 1-based int64 (Julia Int), all reals are IEEE binary64,
 *    matrices are column-major (Julia layout).
 *  - the caller owns every host pointer before and after each call; the library copies in / out and keeps
 *    no host pointers.  Device memory belongs to the handles.
 *  - calls are blocking (stream-synchronised on return) unless fegpu_set_async(ctx, 1) was called.
 *  - there is NO CPU fallback: without a CUDA device fegpu_create fails.
 */
#ifndef FEGPU_H
#define FEGPU_H
#include <stdint.h>

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif
#ifdef __cplusplus
extern "C" {
#endif

typedef struct fegpu_ctx fegpu_ctx;       /* one GPU, one stream                                            */
typedef struct fegpu_mesh fegpu_mesh;     /* FESet connectivity + geometry NodalField + quadrature tables   */
typedef struct fegpu_dofmap fegpu_dofmap; /* NodalField u.dofnums + the cached sparsity pattern             */
typedef struct fegpu_asm fegpu_asm;       /* SysmatAssemblerSparse state: triplet values, CSC result        */

/* element type codes (FESetModule.jl: T3 :661, Q4 :709, T4 :1342, T10 :1393, H8 :952, H20 :1027, H27 :1215) */
enum { FEGPU_T3 = 1, FEGPU_Q4 = 2, FEGPU_T4 = 3, FEGPU_T10 = 4, FEGPU_H8 = 5, FEGPU_H20 = 6, FEGPU_H27 = 7 };

enum {
  FEGPU_OK = 0,
  FEGPU_ERR_CUDA = -1,         /* CUDA runtime error / no device                                              */
  FEGPU_ERR_ARG = -2,          /* bad argument                                                                */
  FEGPU_ERR_COL_LT1 = -11,     /* "Column degree of freedom < 1"      AssemblyModule.jl:268                   */
  FEGPU_ERR_COL_GT = -12,      /* "Column degree of freedom > size"   AssemblyModule.jl:269                   */
  FEGPU_ERR_ROW_LT1 = -13,     /* "Row degree of freedom < 1"         AssemblyModule.jl:272                   */
  FEGPU_ERR_ROW_GT = -14,      /* "Row degree of freedom > size"      AssemblyModule.jl:273                   */
  FEGPU_ERR_MATSIZE = -15,     /* "Wrong size of matrix"              AssemblyModule.jl:265                   */
  FEGPU_ERR_MANIFOLD = -16,    /* "That is the only acceptable option here." IntegDomainModule.jl:543,600     */
  FEGPU_ERR_STATE = -17        /* call out of order (e.g. makematrix before any assembly)                     */
};

/* -- context ------------------------------------------------------------------------------------------ */
int32_t fegpu_create(fegpu_ctx **ctx, int32_t device);
int32_t fegpu_destroy(fegpu_ctx *ctx);
const char *fegpu_last_error(fegpu_ctx *ctx); /* ctx may be NULL: last error of the calling thread        */
/* run all kernels on this cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); default = stream 0 */
int32_t fegpu_set_stream(fegpu_ctx *ctx, void *cuda_stream);
/* async_on = 1: the form calls (fegpu_bilform_*, fegpu_linform_dot) return once their device work is queued; every call that
 * hands results to the host (fegpu_makematrix_copy*, fegpu_coo_copy, fegpu_last_timings, ...) synchronises.  A shim that
 * fetches the matrix right after the form switches it on around the pair: fegpu_makematrix_copy then ships the pattern's
 * arrays, and its host threads rebuild rowval, while the integration and the numeric phase are still running.  Default: off. */
int32_t fegpu_set_async(fegpu_ctx *ctx, int32_t async_on);
/* Fresh assemblies (no cached pattern) run the element integration on a second stream, concurrently with the symbolic
 * phase; both are joined before the numeric phase.  overlap_on = 0 (or FEGPU_OVERLAP=0) makes the phases strictly serial,
 * which is what per-kernel timings (fegpu_last_timings) should be taken with.  Default: on.                  */
int32_t fegpu_set_overlap(fegpu_ctx *ctx, int32_t overlap_on);
int32_t fegpu_synchronize(fegpu_ctx *ctx);
/* The symbolic phase keeps every device block it frees (pattern arrays, temporaries) in a per-context cache so that rebuilding
 * a pattern makes no driver allocation call.  This hands the cached blocks back to the driver (after a device
 * synchronisation), e.g. before another library needs the memory; handles and their results stay valid.               */
int32_t fegpu_cache_release(fegpu_ctx *ctx);
/* number of CUDA kernels this context has launched so far (bench.py's gpu_launches)                     */
int64_t fegpu_launch_count(fegpu_ctx *ctx);
/* roofline denominators measured by the library itself: FP64 FMA peak (TFLOP/s) and copy bandwidth (GB/s) */
int32_t fegpu_measure_peaks(fegpu_ctx *ctx, double *dfma_tflops, double *copy_gbs);
/* Page-locked host memory for result arrays that are re-used across assemblies (colptr / rowval / nzval of a shim that keeps
 * its buffers): DMA into pinned memory runs at link speed, into pageable memory at about half of it.  Plain cudaHostAlloc /
 * cudaFreeHost; the caller owns the block. */
int32_t fegpu_host_alloc(void **p, int64_t bytes);
int32_t fegpu_host_free(void *p);
/* Per-kernel device times of ONE step: fegpu_marks_begin starts recording named CUDA events on the context's stream at the
 * boundaries of the library's kernels (run the step with fegpu_set_overlap(ctx, 0) so that it stays on one stream);
 * fegpu_marks_read synchronises, stops recording and writes "name=ms;name=ms;..." -- the time since the previous mark --
 * into buf (NUL-terminated, truncated at cap).  Measurement aid of bench.py (roofline.achieved); nothing of the reference. */
int32_t fegpu_marks_begin(fegpu_ctx *ctx);
int32_t fegpu_marks_read(fegpu_ctx *ctx, char *buf, int64_t cap);

/* -- mesh: replaces the per-element gathers of fes.conn (FESetModule.jl:61) and geom.values
 *    (gathervalues_asmat!, FieldModule.jl:263-275): uploaded once ------------------------------------- */
int32_t fegpu_mesh_upload(fegpu_ctx *ctx, int32_t etype, int64_t nelem, const int64_t *conn /* [nelem][nne] 1-based */,
                          int64_t nnodes, int32_t sdim, const double *xyz /* nnodes x sdim col-major */, fegpu_mesh **mesh);
int32_t fegpu_mesh_destroy(fegpu_mesh *mesh);
/* new coordinates for the same connectivity (pattern cache stays valid) */
int32_t fegpu_geom_update(fegpu_mesh *mesh, const double *xyz);
/* The same for a partitioned mesh (fegpu_partition_set): only the coordinates of the node window [lo, hi) that holds every
 * node of this rank's active elements cross the link (xyz is still the whole nnodes x sdim array).  fegpu_mesh_window
 * reports that window (0-based, hi exclusive; the whole mesh when it is not partitioned) and the active element count. */
int32_t fegpu_geom_update_window(fegpu_mesh *mesh, const double *xyz);
int32_t fegpu_mesh_window(fegpu_mesh *mesh, int64_t *lo, int64_t *hi, int64_t *nactive);
/* quadrature tables exactly as the caller's integrationdata() produced them (IntegDomainModule.jl:631-648):
 * Ns [npts][nne], gradNpar [npts][mdim][nne] (i.e. each point's nne x mdim matrix, column-major), w [npts] */
int32_t fegpu_rule_set(fegpu_mesh *mesh, int32_t npts, const double *Ns, const double *gradNpar, const double *w);
/* constant material coordinate system of the FEMM, CSys(csmat) (CSysModule.jl:133-144): csmat sdim x mdim column-major, NULL =
 * identity (CSys(dim), the FEMMBase default FEMMBaseModule.jl:82-84).  Used by the general bilform_diffusion (RmTJ = Rm' J,
 * :1495-1497) and by bilform_lin_elastic (:1801-1803, blmat! with Rm).  Position-dependent systems are not eligible. */
int32_t fegpu_csys_set(fegpu_mesh *mesh, const double *csmat);
/* constant other-dimension of the IntegDomain (IntegDomainModule.jl:73-82, 150-152): the thickness that multiplies the
 * surface Jacobian of a 2-manifold in Jacobianvolume (:504-517); used by the gradient forms (diffusion, convection,
 * div_grad) on planar meshes.  Default 1.0 (otherdimensionunity).  bilform_dot / masslike / linform_dot take it per call. */
int32_t fegpu_otherdimension_set(fegpu_mesh *mesh, double otherdimension);
/* multi-GPU row-block ownership (pointpartitioning semantics, MeshModificationModule.jl:1029): this context
 * integrates every element touching a node with node_owner[n] == my_rank and keeps the matrix rows of those
 * nodes' dofs.  node_owner == NULL restores the single-GPU behaviour. */
int32_t fegpu_partition_set(fegpu_mesh *mesh, const int32_t *node_owner /* [nnodes] */, int32_t my_rank);

/* -- dof map: replaces gatherdofnums! (FieldModule.jl:304-314); validates the range checks of assemble!
 *    (AssemblyModule.jl:268-273) once, at upload --------------------------------------------------------- */
int32_t fegpu_dofmap_upload(fegpu_ctx *ctx, fegpu_mesh *mesh, int32_t ndn, const int64_t *dofnums /* nnodes x ndn col-major */,
                            int64_t row_nalldofs, int64_t col_nalldofs, fegpu_dofmap **dofmap);
int32_t fegpu_dofmap_destroy(fegpu_dofmap *dofmap);

/* -- assembler ------------------------------------------------------------------------------------------ */
int32_t fegpu_asm_create(fegpu_ctx *ctx, fegpu_asm **as);
int32_t fegpu_asm_destroy(fegpu_asm *as);
/* on != 0: SysmatAssemblerSparseSymm semantics (AssemblyModule.jl:342-583).  fegpu_assemble keeps the lower triangle of each
 * (square) element matrix (:517-530, "Size mismatch" :510); fegpu_makematrix and the bilform calls deliver
 * S + transpose(S) with the diagonal halved (:576-579), i.e. the symmetric matrix WITHOUT entries that sum to exactly 0.0. */
int32_t fegpu_asm_set_symmetric(fegpu_asm *as, int32_t on);
/* SysmatAssemblerSparseDiag (AssemblyModule.jl:599-794; mode 1) and SysmatAssemblerSparseHRZLumpingSymm (:943-1141; mode 2):
 * every square element matrix contributes only its diagonal, for HRZ scaled by sum(mat) / trace(mat) (:1085-1090); the result
 * is sparse(I = J = dof, V): a diagonal matrix with a stored entry for every dof that appears in an element.  Applies to the
 * bilinear forms and to the generic startassembly!/assemble!/makematrix! protocol ("Size mismatch", "Row and column info do
 * not agree", "Diagonal sparse matrix is assumed to be assembled from square matrices" as in the reference).  Mode 0 = off. */
int32_t fegpu_asm_set_lumping(fegpu_asm *as, int32_t mode);

/* The three bilinear forms.  Each call = startassembly! + the whole element loop + makematrix!
 * (the CSC stays on the device until fegpu_makematrix_copy). */
/* bilform_diffusion, FEMMBaseModule.jl:1462-1535.  kappa_kind 0: scalar (kappa[1]) -> _iso path :1508;
 * 1: mdim x mdim col-major -> _general path :1476 */
int32_t fegpu_bilform_diffusion(fegpu_mesh *mesh, fegpu_dofmap *dofmap, int32_t kappa_kind, const double *kappa, fegpu_asm *as);
/* bilform_lin_elastic with DeforModelRed3D, FEMMBaseModule.jl:1774-1813.  C: 6 x 6 col-major */
int32_t fegpu_bilform_lin_elastic(fegpu_mesh *mesh, fegpu_dofmap *dofmap, const double *C, fegpu_asm *as);
/* bilform_dot, FEMMBaseModule.jl:1335-1366.  c: ndn x ndn col-major; m: manifold dimension kwarg;
 * otherdim: the constant other-dimension of the IntegDomain (IntegDomainModule.jl:150-152: 1.0) */
int32_t fegpu_bilform_dot(fegpu_mesh *mesh, fegpu_dofmap *dofmap, const double *c, int32_t m, double otherdim, fegpu_asm *as);
/* bilform_convection, FEMMBaseModule.jl:1583-1625.  dofmap: the scalar field Q (1 dof per node);
 * uvel: nodal values of the convective velocity field u, nnodes x sdim col-major (NodalField.values), copied to the device
 * by the call; rho: the constant of rhof (evaluated by the reference but absent from its integrand).  Non-symmetric. */
int32_t fegpu_bilform_convection(fegpu_mesh *mesh, fegpu_dofmap *dofmap, const double *uvel, double rho, fegpu_asm *as);
/* bilform_div_grad, FEMMBaseModule.jl:1672-1713.  dofmap: vector field with sdim dofs per node; mu: constant viscosity */
int32_t fegpu_bilform_div_grad(fegpu_mesh *mesh, fegpu_dofmap *dofmap, double mu, fegpu_asm *as);
/* bilform_masslike, FEMMBaseModule.jl:1865-1912: the test function is the indicator of each element, so the matrix is
 * (nelem * ndn) x nalldofs with element e owning rows (e-1)*ndn + 1 .. e*ndn; c: ndn x ndn col-major.  Built through the
 * generic sort path (plain sparse assembler only, no row-block partitions). */
int32_t fegpu_bilform_masslike(fegpu_mesh *mesh, fegpu_dofmap *dofmap, const double *c, int32_t m, double otherdim, fegpu_asm *as);

/* -- vectors: linform_dot / distribloads and SysvecAssembler ---------------------- */
/* linform_dot, FEMMBaseModule.jl:1207-1244 (distribloads :1277-1297 forwards a ForceIntensity's cache to it): F_i = int N_i f
 * over the m-dimensional manifold Jacobian; force: ndn constant components.  The element vectors are summed per node in
 * ascending element order on the mesh's node -> element adjacency (no atomics).  Result: fegpu_makevector_copy.        */
int32_t fegpu_linform_dot(fegpu_mesh *mesh, fegpu_dofmap *dofmap, const double *force, int32_t m, double otherdim, fegpu_asm *as);
/* SysvecAssembler protocol, AssemblyModule.jl:853-917: startassembly!(a, row_nalldofs), assemble!(a, vec, dofnums) with the
 * reference's "Row degree of freedom < 1" / "> size" errors, makevector!(a).  Entries are staged on the host and summed on
 * the device (duplicates left to right). */
int32_t fegpu_vec_startassembly(fegpu_asm *as, int64_t row_nalldofs);
int32_t fegpu_vec_assemble(fegpu_asm *as, const double *vec, const int64_t *dofnums, int64_t n);
int32_t fegpu_makevector(fegpu_asm *as);
int32_t fegpu_makevector_size(fegpu_asm *as, int64_t *n);
int32_t fegpu_makevector_copy(fegpu_asm *as, double *F /* [row_nalldofs] */);

/* Generic assembler protocol for any other caller (AssemblyModule.jl:209-282): host triplets are staged to
 * the device and the CSC is built there by a 64-bit key sort + segmented sum. */
int32_t fegpu_startassembly(fegpu_asm *as, int64_t elem_mat_nrows, int64_t elem_mat_ncols, int64_t n_elem_mats,
                            int64_t row_nalldofs, int64_t col_nalldofs);
/* assemble!: mat is nrows x ncols col-major */
int32_t fegpu_assemble(fegpu_asm *as, const double *mat, const int64_t *dofnums_row, int64_t nrows, const int64_t *dofnums_col,
                       int64_t ncols);
/* bulk form of the same: n ready triplets in emission order */
int32_t fegpu_triplets_append(fegpu_asm *as, int64_t n, const int64_t *I, const int64_t *J, const double *V);
/* makematrix!, AssemblyModule.jl:304-329 (= SparseArrays.sparse(I,J,V,m,n): duplicates summed left to right,
 * explicit zeros kept, rows strictly increasing per column).  Builds the CSC on the device. */
int32_t fegpu_makematrix(fegpu_asm *as);

/* Result access.  sizes: after a bilform call or fegpu_makematrix. copy: straight into the arrays handed to
 * SparseMatrixCSC(m, n, colptr, rowval, nzval): 1-based int64. */
int32_t fegpu_makematrix_sizes(fegpu_asm *as, int64_t *nrows, int64_t *ncols, int64_t *nnz);
int32_t fegpu_makematrix_copy(fegpu_asm *as, int64_t *colptr /* ncols+1 */, int64_t *rowval /* nnz */, double *nzval /* nnz */);
/* View of the assembled matrix (built on the device, the full result stays available for further views):
 *   A[row_first:row_last, col_first:col_last] (1-based, inclusive; stored zeros kept, rows rebased) -- replaces
 *   matrix_blocked_ff/fd/df/dd (MatrixUtilityModule.jl:675-793) and SysmatAssemblerFFBlock's makematrix!
 *   (AssemblyModule.jl:1149-1231);
 *   drop_exact_zeros != 0 removes entries equal to 0.0 -- what SysmatAssemblerSparseSymm's `S + transpose(S)` leaves
 *   (AssemblyModule.jl:551-583; SparseArrays' zero-preserving map stores only non-zero sums).
 * Afterwards sizes / copy / device refer to the view; (1, nrows, 1, ncols, 0) restores the full matrix. */
int32_t fegpu_makematrix_view(fegpu_asm *as, int64_t row_first, int64_t row_last, int64_t col_first, int64_t col_last,
                              int32_t drop_exact_zeros);
/* Transport notes (fegpu_transfer.cu): for results above 1 M nonzeros rowval crosses the PCIe link as int32 and is widened
 * into `rowval` by host threads (FEGPU_HOST_THREADS, default min(cores, 4)); destinations may be pageable or page-locked.
 * fegpu_transfer_stats: staged (int32 + widen) and bypassed (plain int64 DMA) chunk counts so far. */
int32_t fegpu_transfer_stats(fegpu_ctx *ctx, int64_t *staged_chunks, int64_t *bypassed_chunks);
/* Vector fields (ndn >= 2) whose node-major dof order is ascending at every node: rowval does not cross the link at all; the
 * per-node neighbour lists of the device's symbolic phase (int32 per node pair, 1/ndn^2 of the row indices) and the dof map
 * do, and the host threads decode them into `rowval` while nzval is in flight (FEGPU_XFER_COMPRESS=0 turns it off).
 * fegpu_transfer_compressed: number of results delivered that way so far. */
int32_t fegpu_transfer_compressed(fegpu_ctx *ctx, int64_t *results);
/* Every other result of at least 2^20 non-zeros (scalar fields, numberings that are not node-major): the row indices of column j
 * are j + a list of offsets, and meshes repeat a few such lists, so one id per column (4 B) + a dictionary of offset lists cross
 * the link instead of 4 B per non-zero, and the host threads rebuild `rowval` (the codec is built and verified column by column
 * on the device; a matrix with more than 4096 distinct column shapes keeps the int32 transport; FEGPU_XFER_STENCIL=0 turns it
 * off).  fegpu_transfer_stenciled: number of results delivered that way so far.  Nothing of the reference: a transport codec. */
int32_t fegpu_transfer_stenciled(fegpu_ctx *ctx, int64_t *results);
/* nzval only (re-assembly on a cached pattern: colptr/rowval did not change) */
int32_t fegpu_makematrix_copy_values(fegpu_asm *as, double *nzval);
/* device pointers of the current result (valid until the next assembly on this assembler) */
int32_t fegpu_makematrix_device(fegpu_asm *as, const int64_t **d_colptr, const int64_t **d_rowval, const double **d_nzval);
/* raw COO of the last bilform assembly in the reference's emission order (setnomatrixresult(true) flows,
 * AssemblyModule.jl:309-317): n = nelem*elmdim^2 triplets */
int32_t fegpu_coo_copy(fegpu_asm *as, fegpu_mesh *mesh, fegpu_dofmap *dofmap, int64_t *I, int64_t *J, double *V);

/* device milliseconds (CUDA events) of the phases of the last bilform / makematrix call:
 * [0] element integration, [1] symbolic pattern build (0 when cached), [2] numeric CSC gather-sum or sort path,
 * [3] whole call */
/* Optional gather of the ranks' row-block CSCs into ONE matrix on a device (multi-GPU; the reference's makematrix! returns one
 * SparseMatrixCSC, AssemblyModule.jl:319-325).  All pointers are DEVICE pointers (e.g. torch tensors' data_ptr()); the collective
 * steps in between -- an all-gather of the counts, contiguous rowval / nzval slabs to the destination -- are the host side's NCCL
 * calls (finetools.jl_b200/parallel.py: torch.distributed).  No index array crosses the wire: positions follow from the counts.
 *   fegpu_block_counts        d_counts[j] = entries of column j in this assembler's (block) result               [ncols]
 *   fegpu_gather_plan         from the all-gathered counts [world][ncols]: global 1-based colptr [ncols+1], nnz per rank [world]
 *   fegpu_gather_place        copy the block of rank src_rank (its rowval / nzval slabs as received) into the global arrays:
 *                             column j lands at colptr[j] + (entries of the ranks below src_rank in column j)
 *   fegpu_gather_sort_columns only when the ranks' dof ranges interleave (free-first numberings): sorts the columns whose
 *                             concatenation is not ascending; *columns_sorted = how many needed it                         */
int32_t fegpu_block_counts(fegpu_asm *as, int64_t *d_counts);
int32_t fegpu_gather_plan(fegpu_ctx *ctx, const int64_t *d_allcounts, int32_t world, int64_t ncols, int64_t *d_colptr, int64_t *d_nnz_rank);
int32_t fegpu_gather_place(fegpu_ctx *ctx, const int64_t *d_allcounts, int32_t world, int32_t src_rank, int64_t ncols, const int64_t *d_colptr,
                           const int64_t *d_src_rowval, const double *d_src_nzval, int64_t *d_rowval, double *d_nzval);
int32_t fegpu_gather_sort_columns(fegpu_ctx *ctx, int64_t ncols, const int64_t *d_colptr, int64_t *d_rowval, double *d_nzval, int64_t *columns_sorted);
int32_t fegpu_last_timings(fegpu_asm *as, double ms[4]);
/* was the sparsity pattern of the last bilform call served from the cache (1) or built (0)? */
int32_t fegpu_pattern_was_cached(fegpu_asm *as);
/* forget a dofmap's cached pattern (forces the next assembly to rebuild it) */
int32_t fegpu_pattern_invalidate(fegpu_dofmap *dofmap);
/* Which symbolic path built the dof map's cached pattern: 0 = none cached, 1 = group / warp kernels (fegpu_pattern.cu: every
 * element type and numbering), 2 = thread-per-node kernels (fegpu_tile.cu: small stencils, node-major affine dof map). */
int32_t fegpu_pattern_path(fegpu_dofmap *dofmap);

#ifdef __cplusplus
}
#endif
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#endif
