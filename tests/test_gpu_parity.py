"""GPU parity: the CUDA path through the C ABI versus the CPU oracle on the same inputs.
colptr/rowval must be bit-exact, nzval within 1e-12 of the matrix max-abs (BASELINE.json north_star)."""
import numpy as np
import pytest

from conftest import KAPPA3, isotropic_C
from helpers import assert_parity, gpu_csc, make_field, oracle_csc

pytestmark = pytest.mark.gpu


def _mesh(fe, et, n=3):
    dims = (1.3, 3.1, 2.7)
    if et == "H8":
        return fe.H8block(*dims, n, n + 1, n + 2)
    if et == "H20":
        return fe.H20block(*dims, n, n, n + 1)
    if et == "H27":
        return fe.H27block(*dims, n, n, n)
    if et == "T4":
        return fe.T4block(*dims, n, n + 1, n)
    if et == "T10":
        return fe.T10block(*dims, n, n, n + 1)
    raise ValueError(et)


def _distort(fens):
    x = fens.xyz
    h = 0.05
    x[:, 0] += h * np.sin(3 * x[:, 1]) * np.cos(2 * x[:, 2])
    x[:, 1] += h * np.sin(3 * x[:, 2]) * np.cos(2 * x[:, 0])
    x[:, 2] += h * np.sin(3 * x[:, 0]) * np.cos(2 * x[:, 1])
    return fens


VOL_RULES = {"H8": ("gauss", 2), "H20": ("gauss", 3), "H27": ("gauss", 3), "T4": ("tet", 1), "T10": ("tet", 4)}


def _rule(fe, et):
    kind, k = VOL_RULES[et]
    return fe.GaussRule(3, k) if kind == "gauss" else fe.TetRule(k)


@pytest.mark.parametrize("et", ["H8", "H20", "H27", "T4", "T10"])
@pytest.mark.parametrize("kappa", ["matrix", "scalar"])
def test_diffusion_parity(fe, orc, gpu_ctx, et, kappa):
    fens, fes = _mesh(fe, et)
    _distort(fens)
    u = make_field(fe, fens, 1)
    rule = _rule(fe, et)
    coef = KAPPA3 if kappa == "matrix" else 1.7
    ref, _ = oracle_csc(orc, "diffusion", et, fes, fens, u, rule, coef)
    got, _ = gpu_csc(fe, "diffusion", fes, fens, u, rule, coef)
    assert_parity(ref, got)


@pytest.mark.parametrize("et", ["H8", "H20", "H27", "T4", "T10"])
def test_elastic_parity(fe, orc, gpu_ctx, et):
    fens, fes = _mesh(fe, et, 2)
    _distort(fens)
    u = make_field(fe, fens, 3)
    rule = _rule(fe, et)
    C = isotropic_C()
    C[0, 3] = C[3, 0] = 0.05  # a little anisotropy so every D entry matters
    ref, _ = oracle_csc(orc, "elastic", et, fes, fens, u, rule, C)
    got, _ = gpu_csc(fe, "elastic", fes, fens, u, rule, C)
    assert_parity(ref, got)


@pytest.mark.parametrize("et", ["H8", "H20", "H27", "T4", "T10"])
@pytest.mark.parametrize("material", ["isotropic", "cubic"])
def test_elastic_parity_cubic_symmetry_shortcut(fe, orc, gpu_ctx, et, material):
    """A material matrix of the cubic-symmetry form (isotropic: what MatDeforElastIso produces; cubic: D00 - lam != 2 mu) takes the
    outer-product formulation of the elasticity kernels (fe_elastic_cubic, csrc/fegpu_internal.h): same matrix as the reference's
    B' D B loop to rounding, pattern bit-exact, K - K' == 0 exactly (test/test_forms.jl:441-442)."""
    fens, fes = _mesh(fe, et, 2)
    _distort(fens)
    u = make_field(fe, fens, 3)
    rule = _rule(fe, et)
    C = isotropic_C(E=3.1, nu=0.27)
    if material == "cubic":
        C[np.arange(3), np.arange(3)] *= 1.3
        C[3:, 3:] *= 0.7
    ref, _ = oracle_csc(orc, "elastic", et, fes, fens, u, rule, C)
    got, _ = gpu_csc(fe, "elastic", fes, fens, u, rule, C)
    assert_parity(ref, got)
    import scipy.sparse as sp
    K = sp.csc_matrix((got[2], got[1] - 1, got[0] - 1), shape=(u.nalldofs(), u.nalldofs()))
    assert abs(K - K.T).max() == 0.0


@pytest.mark.parametrize("et", ["H8", "H20", "H27", "T4", "T10"])
@pytest.mark.parametrize("ndn", [1, 3])
def test_dot_parity(fe, orc, gpu_ctx, et, ndn):
    fens, fes = _mesh(fe, et, 2)
    _distort(fens)
    u = make_field(fe, fens, ndn)
    rule = _rule(fe, et)
    c = np.array([[1.0]]) if ndn == 1 else np.array([[2.0, 0.1, 0.0], [0.3, 1.0, 0.2], [0.0, 0.4, 3.0]])
    ref, _ = oracle_csc(orc, "dot", et, fes, fens, u, rule, c)
    got, _ = gpu_csc(fe, "dot", fes, fens, u, rule, c)
    assert_parity(ref, got)


@pytest.mark.parametrize("et,m", [("Q4", 2), ("T3", 2), ("Q4", 3)])
def test_surface_dot_parity(fe, orc, gpu_ctx, et, m):
    """Boundary mass on the Q4 / T3 skins of a volume block (config 5): sdim = 3, manifold dim = 2."""
    if et == "Q4":
        fens, vol = fe.H8block(1.3, 3.1, 2.7, 3, 4, 2)
        rule = fe.GaussRule(2, 2)
    else:
        fens, vol = fe.T4block(1.3, 3.1, 2.7, 3, 4, 2)
        rule = fe.TriRule(3)
    _distort(fens)
    bfes = fe.meshboundary(vol)
    u = make_field(fe, fens, 1)
    c = np.array([[1.0]])
    ref, _ = oracle_csc(orc, "dot", et, bfes, fens, u, rule, c, m=m, otherdim=1.0)
    got, _ = gpu_csc(fe, "dot", bfes, fens, u, rule, c, m=m)
    assert_parity(ref, got)
    # interior nodes have empty columns: colptr must still have ncols + 1 entries
    assert got[0].size == u.nalldofs() + 1


@pytest.mark.parametrize("et", ["Q4", "T3"])
def test_planar_diffusion_parity(fe, orc, gpu_ctx, et):
    fens, fes = (fe.Q4block(2.0, 1.0, 5, 4) if et == "Q4" else fe.T3block(2.0, 1.0, 5, 4))
    u = make_field(fe, fens, 1)
    rule = fe.GaussRule(2, 2) if et == "Q4" else fe.TriRule(3)
    kap = np.array([[1.5, 0.2], [0.2, 2.5]])
    ref, _ = oracle_csc(orc, "diffusion", et, fes, fens, u, rule, kap)
    got, _ = gpu_csc(fe, "diffusion", fes, fens, u, rule, kap)
    assert_parity(ref, got)


def test_config1_h8_20cube(fe, orc, gpu_ctx):
    """BASELINE config 1 + the fingerprints of SURVEY.md appendix B."""
    fens, fes = fe.H8block(12.0, 1.1, 0.32, 20, 20, 20)
    u = make_field(fe, fens, 1)
    rule = fe.GaussRule(3, 2)
    ref, _ = oracle_csc(orc, "diffusion", "H8", fes, fens, u, rule, KAPPA3)
    got, a = gpu_csc(fe, "diffusion", fes, fens, u, rule, KAPPA3)
    assert_parity(ref, got)
    colptr, rowval, nzval = got[0], got[1], got[2]
    assert nzval.size == 61 ** 3
    np.testing.assert_array_equal(colptr[:5], [1, 9, 21, 33, 45])
    np.testing.assert_array_equal(rowval[:8], [1, 2, 22, 23, 442, 443, 463, 464])
    assert abs(nzval[0] - 0.88226262626262) < 1e-12


def test_ebc_permuted_dofnums(fe, orc, gpu_ctx):
    """Free-first / fixed-last numbering (FieldModule.jl:360-377) makes dofnums a non-trivial permutation."""
    fens, fes = fe.H8block(1.3, 3.1, 2.7, 3, 3, 3)
    u = make_field(fe, fens, 3, fixed_nodes=[1, 2, 3, 17, 40], fixed_comp=None)
    fe.setebc(u, [5, 9], True, 2, 0.0)
    fe.numberdofs(u)
    rule = fe.GaussRule(3, 2)
    C = isotropic_C()
    ref, _ = oracle_csc(orc, "elastic", "H8", fes, fens, u, rule, C)
    got, _ = gpu_csc(fe, "elastic", fes, fens, u, rule, C)
    assert_parity(ref, got)


def test_reassembly_is_bit_identical_and_cached(fe, orc, gpu_ctx):
    """test/test_basics.jl:3039-3045: repeated assemblies with one assembler are bit-identical; the second one is
    served from the cached pattern."""
    fens, fes = fe.T10block(1.0, 1.0, 1.0, 3, 3, 3)
    u = make_field(fe, fens, 1)
    rule = fe.TetRule(4)
    c = np.array([[1.0]])
    a = fe.SysmatAssemblerSparseGPU(0.0)
    femm = fe.FEMMBase(fe.IntegDomain(fes, rule))
    geom = fe.NodalField(fens.xyz)
    r1 = fe.bilform_dot(femm, a, geom, u, fe.DataCache(c), raw=True)
    assert not a.pattern_was_cached()
    r2 = fe.bilform_dot(femm, a, geom, u, fe.DataCache(c), raw=True)
    assert a.pattern_was_cached()
    r3 = fe.bilform_dot(femm, a, geom, u, fe.DataCache(c), raw=True)
    for k in range(3):
        np.testing.assert_array_equal(r1[k], r2[k])
        np.testing.assert_array_equal(r1[k], r3[k])
    # moved geometry, same pattern: values change, pattern arrays do not
    geom.values[:, 0] *= 1.5
    r4 = fe.bilform_dot(femm, a, geom, u, fe.DataCache(c), raw=True)
    assert a.pattern_was_cached()
    np.testing.assert_array_equal(r1[1], r4[1])
    fens2 = fe.FENodeSet(geom.values)
    ref, _ = oracle_csc(orc, "dot", "T10", fes, fens2, u, rule, c)
    assert_parity(ref, r4)


def test_raw_coo_matches_reference_emission_order(fe, orc, gpu_ctx):
    fens, fes = fe.H8block(1.3, 3.1, 2.7, 2, 3, 2)
    u = make_field(fe, fens, 1)
    rule = fe.GaussRule(3, 2)
    ref, (I, J, V) = oracle_csc(orc, "diffusion", "H8", fes, fens, u, rule, KAPPA3)
    got, a = gpu_csc(fe, "diffusion", fes, fens, u, rule, KAPPA3)
    gI, gJ, gV = a.coo()
    np.testing.assert_array_equal(gI, I)
    np.testing.assert_array_equal(gJ, J)
    assert np.abs(gV - V).max() <= 1e-12 * np.abs(V).max()


def test_generic_protocol_testA(fe, gpu_ctx):
    """test/test_basics.jl:72-129: two dense blocks into a 7x7 matrix (known answer testA, tol 1e-5)."""
    m1 = np.array([[0.24406, 0.599773, 0.833404, 0.0420141],
                   [0.786024, 0.00206713, 0.995379, 0.780298],
                   [0.845816, 0.198459, 0.355149, 0.224996]])
    m1 = m1.T @ m1
    i1 = [5, 2, 1, 4]
    m2 = np.array([[0.146618, 0.53471, 0.614342, 0.737833],
                   [0.479719, 0.41354, 0.00760941, 0.836455],
                   [0.254868, 0.476189, 0.460794, 0.00919633],
                   [0.159064, 0.261821, 0.317078, 0.77646],
                   [0.643538, 0.429817, 0.59788, 0.958909]])
    m2 = m2.T @ m2
    i2 = [2, 3, 1, 5]
    testA = np.array([[2.85928, 1.21875, 0.891063, 0.891614, 2.56958, 0.0, 0.0],
                      [1.21875, 1.15515, 0.716396, 0.0714644, 1.56825, 0.0, 0.0],
                      [0.891063, 0.716396, 0.936979, 0.0, 1.36026, 0.0, 0.0],
                      [0.891614, 0.0714644, 0.0, 0.661253, 0.813892, 0.0, 0.0],
                      [2.56958, 1.56825, 1.36026, 0.813892, 4.15934, 0.0, 0.0],
                      [0.0] * 7, [0.0] * 7])
    a = fe.SysmatAssemblerSparseGPU(0.0)
    fe.startassembly(a, 5, 5, 3, 7, 7)
    fe.assemble(a, m1, i1, i1)
    fe.assemble(a, m2, i2, i2)
    A = fe.makematrix(a)
    assert np.abs(testA - A.toarray()).max() < 1.0e-5
    # the assembler is reusable right away (AssemblyModule.jl:327)
    fe.startassembly(a, 5, 5, 3, 7, 7)
    fe.assemble(a, m1, i1, i1)
    B = fe.makematrix(a)
    M = np.zeros((7, 7))
    M[np.ix_(np.array(i1) - 1, np.array(i1) - 1)] += m1
    assert np.abs(M - B.toarray()).max() < 1e-14


def test_generic_protocol_random_blocks_vs_sparse_oracle(fe, orc, gpu_ctx):
    """test/test_basics.jl:2087-2172 pattern: many small blocks with repeated dofs, rectangular target; compared with the
    oracle's sparse() bit for bit in colptr/rowval and to rounding in nzval (same left-to-right sum order => exact)."""
    rng = np.random.default_rng(1234)
    nr, nc = 137, 91
    a = fe.SysmatAssemblerSparseGPU(0.0)
    fe.startassembly(a, 4, 3, 10, nr, nc)  # deliberately undersized, like the reference's buffer-growth test
    Is, Js, Vs = [], [], []
    for _ in range(1500):
        dr = rng.integers(1, nr + 1, size=4)
        dc = rng.integers(1, nc + 1, size=3)
        m = rng.standard_normal((4, 3))
        fe.assemble(a, m, dr, dc)
        for j in range(3):
            for i in range(4):
                Is.append(dr[i]); Js.append(dc[j]); Vs.append(m[i, j])
    colptr, rowval, nzval, mm, nn = fe.makematrix(a, raw=True)
    cp, rv, nz = orc.sparse(np.array(Is), np.array(Js), np.array(Vs), nr, nc)
    assert (mm, nn) == (nr, nc)
    np.testing.assert_array_equal(colptr, cp)
    np.testing.assert_array_equal(rowval, rv)
    np.testing.assert_array_equal(nzval, nz)  # same summation order: bit-exact


def test_dof_range_errors_use_reference_strings(fe, gpu_ctx):
    a = fe.SysmatAssemblerSparseGPU(0.0)
    fe.startassembly(a, 2, 2, 1, 5, 5)
    with pytest.raises(fe.FEGPUError, match="Row degree of freedom > size"):
        fe.assemble(a, np.eye(2), [1, 6], [1, 2])
    with pytest.raises(fe.FEGPUError, match="Column degree of freedom < 1"):
        fe.assemble(a, np.eye(2), [1, 2], [0, 2])
    with pytest.raises(fe.FEGPUError, match="Wrong size of matrix"):
        fe.assemble(a, np.eye(3), [1, 2], [1, 2])
    # through the form path: a dof number beyond nalldofs is caught at upload
    fens, fes = fe.H8block(1, 1, 1, 2, 2, 2)
    u = make_field(fe, fens, 1)
    u.dofnums = u.dofnums.copy()  # numberdofs hands out a read-only array; a hand-edited numbering is a new array
    u.dofnums[3, 0] = u.nalldofs() + 5
    with pytest.raises(fe.FEGPUError, match="degree of freedom > size"):
        gpu_csc(fe, "diffusion", fes, fens, u, fe.GaussRule(3, 2), KAPPA3)


def test_non_injective_dofmap_takes_sort_path(fe, orc, gpu_ctx):
    """Two nodes tied to one dof (periodic-style numbering): the mesh-structured pattern does not apply; the generic
    sort path must give the reference's sparse() result."""
    fens, fes = fe.H8block(1.0, 1.0, 1.0, 3, 2, 2)
    u = make_field(fe, fens, 1)
    u.dofnums = u.dofnums.copy()
    u.dofnums[u.dofnums == u.nalldofs()] = 1  # last node shares dof 1
    rule = fe.GaussRule(3, 2)
    ref, _ = oracle_csc(orc, "diffusion", "H8", fes, fens, u, rule, KAPPA3)
    got, _ = gpu_csc(fe, "diffusion", fes, fens, u, rule, KAPPA3)
    assert_parity(ref, got)


def test_degenerate_element_takes_sort_path(fe, orc, gpu_ctx):
    """A collapsed hexahedron (one node listed twice) is legal input for assemble!; duplicates inside one element
    matrix are summed by sparse()."""
    fens, fes = fe.H8block(1.0, 1.0, 1.0, 2, 2, 2)
    fes.conn[0, 1] = fes.conn[0, 0]
    u = make_field(fe, fens, 1)
    rule = fe.GaussRule(3, 2)
    c = np.array([[1.0]])
    ref, _ = oracle_csc(orc, "dot", "H8", fes, fens, u, rule, c)
    got, _ = gpu_csc(fe, "dot", fes, fens, u, rule, c)
    assert_parity(ref, got)


@pytest.mark.parametrize("nparts", [2, 4])
def test_row_block_partition_reassembles_the_matrix(fe, orc, gpu_ctx, nparts):
    """Multi-GPU semantics on one device: rank p keeps the rows of its nodes; the blocks are disjoint and their union is
    the single-GPU matrix (same pattern, values to rounding)."""
    fens, fes = fe.H8block(1.0, 1.0, 1.0, 4, 3, 6)
    u = make_field(fe, fens, 3)
    rule = fe.GaussRule(3, 2)
    C = isotropic_C()
    ref, _ = oracle_csc(orc, "elastic", "H8", fes, fens, u, rule, C)
    n = u.nalldofs()
    import scipy.sparse as sp
    full = sp.csc_matrix((ref[2], ref[1] - 1, ref[0] - 1), shape=(n, n))
    owner = fe.slab_owner(fens.count(), nparts)
    total = sp.csc_matrix((n, n))
    nnz_sum = 0
    for p in range(nparts):
        got, _ = gpu_csc(fe, "elastic", fes, fens, u, rule, C, node_owner=owner, my_rank=p)
        colptr, rowval, nzval, mm, nn = got
        blk = sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(n, n))
        rows_owned = np.zeros(n, bool)
        rows_owned[(u.dofnums[owner == p] - 1).reshape(-1)] = True
        assert rows_owned[rowval - 1].all()
        # rows strictly increasing inside every column
        for j in range(0, n, 7):
            seg = rowval[colptr[j] - 1: colptr[j + 1] - 1]
            assert np.all(np.diff(seg) > 0)
        nnz_sum += nzval.size
        total = total + blk
    assert nnz_sum == ref[2].size
    diff = (total - full)
    assert abs(diff).max() <= 1e-12 * np.abs(ref[2]).max()


def _row_block(ref, keep_rows, n):
    """The rows `keep_rows` (bool per dof) of the CSC triple `ref`, every column kept."""
    mask = keep_rows[ref[1] - 1]
    col_of = np.repeat(np.arange(n), np.diff(ref[0]))
    cnt = np.bincount(col_of[mask], minlength=n)
    return np.concatenate(([1], 1 + np.cumsum(cnt))).astype(np.int64), ref[1][mask], ref[2][mask]


@pytest.mark.parametrize("variant", ["slab_elastic", "slab_ebc", "inertial_scalar", "t10_three_slabs", "tail_nodes_unused"])
def test_partition_node_window_variants(fe, orc, gpu_ctx, variant):
    """The symbolic phase of a partitioned mesh runs over the node window of the rank's active elements (and colptr over
    the dof range of that window).  Every rank's block must still be the oracle's rows-of-owned-nodes block, bit-exact in
    colptr / rowval (complete colptr: constant before and after the window), for contiguous slabs, a numbering with the
    fixed dofs last (non-monotone dof map, dof range != node window), a non-contiguous owner map (recursive inertial
    bisection), mixed-valence T10 nodes, and a mesh whose last nodes belong to no element."""
    rule = fe.GaussRule(3, 2)
    if variant == "slab_elastic":
        fens, fes = fe.H8block(1.0, 1.0, 1.0, 3, 4, 9)
        u = make_field(fe, fens, 3)
        form, et, coef, nparts = "elastic", "H8", isotropic_C(), 4
        owner = fe.slab_owner(fens.count(), nparts)
    elif variant == "slab_ebc":
        fens, fes = fe.H8block(1.0, 1.0, 1.0, 3, 3, 8)
        u = make_field(fe, fens, 3, fixed_nodes=np.arange(5, fens.count(), 7) + 1, fixed_comp=[1, 3])
        form, et, coef, nparts = "elastic", "H8", isotropic_C(), 3
        owner = fe.slab_owner(fens.count(), nparts)
    elif variant == "inertial_scalar":
        fens, fes = fe.H8block(2.0, 1.0, 1.5, 6, 4, 5)
        u = make_field(fe, fens, 1)
        form, et, coef, nparts = "diffusion", "H8", KAPPA3, 4
        owner = fe.pointpartitioning(fens.xyz, nparts) - 1
    elif variant == "t10_three_slabs":
        fens, fes = fe.T10block(1.0, 1.0, 1.0, 2, 2, 5)
        u = make_field(fe, fens, 1)
        rule = fe.TetRule(4)
        form, et, coef, nparts = "dot", "T10", np.array([[1.0]]), 3
        owner = fe.slab_owner(fens.count(), nparts)
    else:
        fens, fes = fe.H8block(1.0, 1.0, 1.0, 3, 3, 6)
        fes = fes.subset(np.arange(27))  # the three lowest element layers: the upper nodes belong to no element
        u = make_field(fe, fens, 1)
        form, et, coef, nparts = "diffusion", "H8", 1.7, 2
        owner = fe.slab_owner(fens.count(), nparts)
    owner = np.asarray(owner, dtype=np.int32)
    ref, _ = oracle_csc(orc, form, et, fes, fens, u, rule, coef)
    n = u.nalldofs()
    nnz_sum = 0
    for p in range(nparts):
        keep = np.zeros(n, bool)
        keep[(u.dofnums[owner == p] - 1).reshape(-1)] = True
        blk = _row_block(ref, keep, n)
        got, a = gpu_csc(fe, form, fes, fens, u, rule, coef, node_owner=owner, my_rank=p)
        scale = np.abs(ref[2]).max()
        np.testing.assert_array_equal(got[0], blk[0])
        np.testing.assert_array_equal(got[1], blk[1])
        assert got[2].size == blk[2].size
        if blk[2].size:
            assert np.abs(got[2] - blk[2]).max() <= 1e-12 * scale
        # a second, cached assembly on the same partition is bit-identical
        got2, _ = gpu_csc(fe, form, fes, fens, u, rule, coef, assembler=a, node_owner=owner, my_rank=p)
        assert a.pattern_was_cached()
        np.testing.assert_array_equal(got2[2], got[2])
        nnz_sum += got[2].size
    assert nnz_sum == ref[2].size


@pytest.mark.parametrize("ndn", [1, 3])
def test_empty_feset_and_rank_without_nodes(fe, orc, gpu_ctx, ndn):
    """Edge cases of the mesh-structured path: (a) an FESet without elements -- the reference's startassembly!(…, 0, …) +
    makematrix! gives sparse(Int[], Int[], Float64[], n, n): colptr all ones, no stored entry; (b) a rank that owns no node
    (empty node window): the same empty block, and the other rank's block is then the whole matrix; (c) one single element."""
    rule = fe.GaussRule(3, 2)
    fens, fes = fe.H8block(1.0, 2.0, 3.0, 2, 2, 2)
    u = make_field(fe, fens, ndn)
    n = u.nalldofs()
    form, coef = ("elastic", isotropic_C()) if ndn == 3 else ("diffusion", KAPPA3)
    # (a)
    got, a = gpu_csc(fe, form, fes.subset(np.arange(0)), fens, u, rule, coef)
    np.testing.assert_array_equal(got[0], np.ones(n + 1, np.int64))
    assert got[1].size == 0 and got[2].size == 0 and (got[3], got[4]) == (n, n)
    cp, rv, nz = orc.sparse(np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0), n, n)
    np.testing.assert_array_equal(got[0], cp)
    # (b)
    ref, _ = oracle_csc(orc, form, "H8", fes, fens, u, rule, coef)
    owner = np.zeros(fens.count(), np.int32)
    got1, _ = gpu_csc(fe, form, fes, fens, u, rule, coef, node_owner=owner, my_rank=1)
    np.testing.assert_array_equal(got1[0], np.ones(n + 1, np.int64))
    assert got1[1].size == 0 and got1[2].size == 0
    got0, _ = gpu_csc(fe, form, fes, fens, u, rule, coef, node_owner=owner, my_rank=0)
    assert_parity(ref, got0)
    # (c)
    one = fes.subset(np.arange(3, 4))
    ref1, _ = oracle_csc(orc, form, "H8", one, fens, u, rule, coef)
    gotc, _ = gpu_csc(fe, form, one, fens, u, rule, coef)
    assert_parity(ref1, gotc)


def test_block_cache_release_keeps_results_valid(fe, orc, gpu_ctx):
    """fegpu_cache_release hands the symbolic phase's cached device blocks back to the driver: the resident result, the cached
    pattern and later fresh assemblies (which allocate anew) are unaffected, bit for bit."""
    fens, fes = fe.H8block(1.0, 1.0, 1.0, 5, 4, 3)
    u = make_field(fe, fens, 3)
    rule = fe.GaussRule(3, 2)
    ref, _ = oracle_csc(orc, "elastic", "H8", fes, fens, u, rule, isotropic_C())
    got, a = gpu_csc(fe, "elastic", fes, fens, u, rule, isotropic_C())
    assert_parity(ref, got)
    a.ctx.release_cache()
    again = a._fetch(True)                       # the resident CSC is still there
    np.testing.assert_array_equal(again[1], got[1])
    np.testing.assert_array_equal(again[2], got[2])
    got2, _ = gpu_csc(fe, "elastic", fes, fens, u, rule, isotropic_C(), assembler=a)   # cached pattern
    assert a.pattern_was_cached()
    np.testing.assert_array_equal(got2[2], got[2])
    a.invalidate_patterns()
    a.ctx.release_cache()
    got3, _ = gpu_csc(fe, "elastic", fes, fens, u, rule, isotropic_C(), assembler=a)   # fresh build from an empty cache
    assert not a.pattern_was_cached()
    assert_parity(ref, got3)
    np.testing.assert_array_equal(got3[2], got[2])


@pytest.mark.parametrize("et,ndn,form", [("H8", 3, "elastic"), ("T10", 1, "dot"), ("H20", 1, "diffusion"), ("T4", 3, "elastic")])
def test_scrambled_mesh_numbering_parity(fe, orc, gpu_ctx, et, ndn, form):
    """Nothing in the path may lean on the block generators' numbering: the nodes are renumbered by a random permutation, the
    elements shuffled, and the dof numbers are an arbitrary permutation of 1..nalldofs (a3 of SURVEY 8: "arbitrary
    permutation") -- no locality, no node-major order, every column needs the sorted-rows path.  Whole matrix and a 3-way
    partition by recursive inertial bisection (owner map without locality in the node ids: the node window is everything)."""
    rng = np.random.default_rng(7)
    fens, fes = _mesh(fe, et, 3 if et in ("H8", "T4") else 2)
    _distort(fens)
    nn = fens.count()
    perm = rng.permutation(nn)                      # old node i becomes node perm[i]
    xyz = np.empty_like(fens.xyz)
    xyz[perm] = fens.xyz
    conn = perm[fes.conn - 1] + 1
    conn = np.ascontiguousarray(conn[rng.permutation(conn.shape[0])])
    fens2, fes2 = fe.FENodeSet(xyz), type(fes)(conn)
    u = fe.NodalField(np.zeros((nn, ndn)))
    u.dofnums[:] = (rng.permutation(nn * ndn) + 1).reshape(nn, ndn)
    u._dofver += 1
    rule = _rule(fe, et)
    coef = {"elastic": isotropic_C(), "dot": np.array([[1.0]]), "diffusion": KAPPA3}[form]
    ref, _ = oracle_csc(orc, form, et, fes2, fens2, u, rule, coef)
    got, _ = gpu_csc(fe, form, fes2, fens2, u, rule, coef)
    assert_parity(ref, got)
    n = u.nalldofs()
    owner = np.asarray(fe.pointpartitioning(fens2.xyz, 3) - 1, dtype=np.int32)
    nnz_sum = 0
    for p in range(int(owner.max()) + 1):
        keep = np.zeros(n, bool)
        keep[(u.dofnums[owner == p] - 1).reshape(-1)] = True
        blk = _row_block(ref, keep, n)
        gotp, _ = gpu_csc(fe, form, fes2, fens2, u, rule, coef, node_owner=owner, my_rank=p)
        np.testing.assert_array_equal(gotp[0], blk[0])
        np.testing.assert_array_equal(gotp[1], blk[1])
        if blk[2].size:
            assert np.abs(gotp[2] - blk[2]).max() <= 1e-12 * np.abs(ref[2]).max()
        nnz_sum += gotp[2].size
    assert nnz_sum == ref[2].size


def _sampled_symmetry(colptr, rowval, nzval, cols):
    """K[i,j] == K[j,i] bit for bit on the entries of the sampled columns (binary search in the partner column)."""
    for j in cols:
        lo, hi = colptr[j] - 1, colptr[j + 1] - 1
        for k in range(lo, hi, max(1, (hi - lo) // 9)):
            i = rowval[k] - 1
            seg = rowval[colptr[i] - 1: colptr[i + 1] - 1]
            pos = np.searchsorted(seg, j + 1)
            assert pos < seg.size and seg[pos] == j + 1
            assert nzval[colptr[i] - 1 + pos] == nzval[k]


def test_full_size_config2_properties(fe, gpu_ctx):
    """BASELINE config 2 at full size (128^3 H8 elasticity, 2.1 M elements, 1.2 G triplets): size-independent properties.
    nnz = 9*385^3; rows ascending; K exactly symmetric (the element triangle is mirrored, same summation order on both
    sides); rigid translations in the null space; a second assembly is bit-identical and served from the cached pattern."""
    import scipy.sparse as sp
    n = 128
    fens, fes = fe.H8block(1.0, 1.0, 1.0, n, n, n)
    u = make_field(fe, fens, 3)
    rule = fe.GaussRule(3, 2)
    a = fe.SysmatAssemblerSparseGPU(0.0)
    femm = fe.FEMMBase(fe.IntegDomain(fes, rule))
    geom = fe.NodalField(fens.xyz)
    colptr, rowval, nzval, m_, n_ = fe.bilform_lin_elastic(femm, a, geom, u, fe.DeforModelRed3D, fe.DataCache(isotropic_C()), raw=True)
    assert nzval.size == 9 * 385 ** 3 and m_ == n_ == 3 * 129 ** 3
    assert colptr[0] == 1 and colptr[-1] == nzval.size + 1 and np.all(np.diff(colptr) > 0)
    d = np.diff(rowval)
    d[colptr[1:-1] - 2] = 1
    assert np.all(d > 0)  # rows strictly increasing inside every column
    del d
    K = sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(m_, n_))
    scale = np.abs(nzval).max()
    for comp in range(3):
        v = np.zeros(m_)
        v[u.dofnums[:, comp] - 1] = 1.0
        assert np.abs(K @ v).max() <= 1e-10 * scale
    del K
    _sampled_symmetry(colptr, rowval, nzval, np.arange(0, n_, 150001))
    nz2 = np.empty_like(nzval)
    fe.bilform_lin_elastic(femm, a, geom, u, fe.DeforModelRed3D, fe.DataCache(isotropic_C()), raw=True, out=(colptr, rowval, nz2))
    assert a.pattern_was_cached()
    np.testing.assert_array_equal(nzval, nz2)


def test_full_size_config4_properties(fe, gpu_ctx):
    """BASELINE config 4 on one GPU (256^3 H8 diffusion, 16.8 M elements): nnz = 769^3, constants in the null space,
    symmetry, the same matrix from two row-block halves (multi-GPU semantics on one device)."""
    import scipy.sparse as sp
    n = 256
    fens, fes = fe.H8block(1.0, 1.0, 1.0, n, n, n)
    u = make_field(fe, fens, 1)
    rule = fe.GaussRule(3, 2)
    a = fe.SysmatAssemblerSparseGPU(0.0)
    femm = fe.FEMMBase(fe.IntegDomain(fes, rule))
    geom = fe.NodalField(fens.xyz)
    colptr, rowval, nzval, m_, n_ = fe.bilform_diffusion(femm, a, geom, u, fe.DataCache(KAPPA3), raw=True)
    assert nzval.size == 769 ** 3 and m_ == n_ == 257 ** 3
    K = sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(m_, n_))
    assert np.abs(K @ np.ones(m_)).max() <= 1e-10 * np.abs(nzval).max()
    del K
    _sampled_symmetry(colptr, rowval, nzval, np.arange(0, n_, 400009))
    owner = fe.slab_owner(fens.count(), 2)
    nnz_blocks = 0
    for p in range(2):
        cp, rv, nz, _, _ = fe.bilform_diffusion(femm, a, geom, u, fe.DataCache(KAPPA3), raw=True, node_owner=owner, my_rank=p)
        nnz_blocks += nz.size
        # a column in the middle of rank p's slab lies entirely inside the block: identical to the full matrix's column
        j = int(np.nonzero(owner == p)[0][owner[owner == p].size // 2])
        full = slice(colptr[j] - 1, colptr[j + 1] - 1)
        blk = slice(cp[j] - 1, cp[j + 1] - 1)
        np.testing.assert_array_equal(rv[blk], rowval[full])
        np.testing.assert_array_equal(nz[blk], nzval[full])
    assert nnz_blocks == nzval.size


def test_config5_mixed_surface_volume_properties(fe, gpu_ctx):
    """BASELINE config 5 (mixed surface / volume run), one rank's share of it: H20 elasticity (GaussRule(3,3), the register-tiled
    lane-per-tile kernel) on a 40^3 block, and the boundary mass of the FULL 96^3 skins (Q4 GaussRule(2,2), T3 TriRule(3),
    m = 2).  Size-independent properties: nnz of the serendipity pattern 9 (234 n^3 + 141 n^2 + 24 n + 1),
    rows ascending, exact symmetry, rigid translations in the null space, a 4-way slab partition's blocks adding up to nnz;
    the skin mass matrices sum to the surface area 6."""
    import scipy.sparse as sp
    n = 40
    fens, fes = fe.H20block(1.0, 1.0, 1.0, n, n, n)
    u = make_field(fe, fens, 3)
    rule = fe.GaussRule(3, 3)
    a = fe.SysmatAssemblerSparseGPU(0.0)
    femm = fe.FEMMBase(fe.IntegDomain(fes, rule))
    geom = fe.NodalField(fens.xyz)
    colptr, rowval, nzval, m_, n_ = fe.bilform_lin_elastic(femm, a, geom, u, fe.DeforModelRed3D, fe.DataCache(isotropic_C()), raw=True)
    assert nzval.size == 9 * (234 * n ** 3 + 141 * n ** 2 + 24 * n + 1) and m_ == n_ == 3 * fens.count()
    d = np.diff(rowval)
    d[colptr[1:-1] - 2] = 1
    assert np.all(d > 0)
    del d
    K = sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(m_, n_))
    scale = np.abs(nzval).max()
    for comp in range(3):
        v = np.zeros(m_)
        v[u.dofnums[:, comp] - 1] = 1.0
        assert np.abs(K @ v).max() <= 1e-10 * scale
    del K
    _sampled_symmetry(colptr, rowval, nzval, np.arange(0, n_, 50021))
    owner = fe.slab_owner(fens.count(), 4)
    a.setnomatrixresult(True)
    nnz_blocks = 0
    for p in range(4):
        fe.bilform_lin_elastic(femm, a, geom, u, fe.DeforModelRed3D, fe.DataCache(isotropic_C()), raw=True, node_owner=owner, my_rank=p)
        nnz_blocks += a.sizes()[2]
    assert nnz_blocks == nzval.size
    del colptr, rowval, nzval
    # the skins at BASELINE size
    nb = 96
    for mesher, srule in ((fe.H8block, fe.GaussRule(2, 2)), (fe.T4block, fe.TriRule(3))):
        vf, vol = mesher(1.0, 1.0, 1.0, nb, nb, nb)
        skin = fe.meshboundary(vol)
        psi = make_field(fe, vf, 1)
        sa = fe.SysmatAssemblerSparseGPU(0.0)
        sfemm = fe.FEMMBase(fe.IntegDomain(skin, srule))
        cp, rv, nz, sm, sn = fe.bilform_dot(sfemm, sa, fe.NodalField(vf.xyz), psi, fe.DataCache(np.array([[1.0]])), m=2, raw=True)
        assert sm == sn == vf.count() and cp[-1] == nz.size + 1
        assert abs(nz.sum() - 6.0) <= 1e-9
        on_skin = np.unique(skin.conn) - 1
        cols = np.diff(cp)
        assert np.count_nonzero(cols) == on_skin.size and np.all(cols[psi.dofnums[on_skin, 0] - 1] > 0)


def test_full_size_config3_t10_mass_cached_reassembly(fe, gpu_ctx):
    """BASELINE config 3: consistent mass on a distorted T10 block (100^3 cells, 6 M quadratic tets), TetRule(4), then
    re-assembly on the cached pattern after the geometry moved.  1'M1 = sum of the tet volumes (straight-edged T10)."""
    n = 100
    f4, s4 = fe.T4block(1.0, 1.0, 1.0, n, n, n)
    h = 1.0 / n
    x = f4.xyz
    x0 = x.copy()
    x[:, 0] += 0.2 * h * np.sin(3 * np.pi * x0[:, 1]) * np.cos(2 * np.pi * x0[:, 2])
    x[:, 1] += 0.2 * h * np.sin(3 * np.pi * x0[:, 2]) * np.cos(2 * np.pi * x0[:, 0])
    x[:, 2] += 0.2 * h * np.sin(3 * np.pi * x0[:, 0]) * np.cos(2 * np.pi * x0[:, 1])
    fens, fes = fe.T4toT10(f4, s4)
    assert fes.count() == 6 * n ** 3 and fens.count() == (2 * n + 1) ** 3

    def tetvol(xyz):
        c = fes.conn[:, :4] - 1
        e1, e2, e3 = xyz[c[:, 1]] - xyz[c[:, 0]], xyz[c[:, 2]] - xyz[c[:, 0]], xyz[c[:, 3]] - xyz[c[:, 0]]
        return (np.einsum("ij,ij->i", np.cross(e1, e2), e3) / 6.0).sum()

    u = make_field(fe, fens, 1)
    a = fe.SysmatAssemblerSparseGPU(0.0)
    femm = fe.FEMMBase(fe.IntegDomain(fes, fe.TetRule(4)))
    geom = fe.NodalField(fens.xyz)
    colptr, rowval, nzval, m_, n_ = fe.bilform_dot(femm, a, geom, u, fe.DataCache(np.eye(1)), raw=True)
    assert nzval.size == 230 * n ** 3 + 138 * n ** 2 + 24 * n + 1
    vol = tetvol(fens.xyz)
    assert abs(nzval.sum() - vol) <= 1e-9 * vol
    # move the geometry (affine stretch keeps mid-edge nodes at the midpoints), re-assemble on the cached pattern
    geom.values[:, 2] *= 1.25
    nz2 = np.empty_like(nzval)
    fe.bilform_dot(femm, a, geom, u, fe.DataCache(np.eye(1)), raw=True, out=(colptr, rowval, nz2))
    assert a.pattern_was_cached()
    assert abs(nz2.sum() - 1.25 * vol) <= 1e-9 * vol
    assert np.abs(nz2 - 1.25 * nzval).max() <= 1e-12 * np.abs(nz2).max()


def test_golden_fixtures_on_gpu(fe, gpu_ctx):
    """The committed golden CSC fixtures (tests/golden/, generated by make_golden.py with the oracle) against the CUDA path."""
    import glob
    import os
    from golden.make_golden import CASES, build_case
    root = os.path.dirname(os.path.abspath(__file__))
    files = sorted(glob.glob(os.path.join(root, "golden", "*.npz")))
    assert len(files) == len(CASES)
    for f in files:
        g = np.load(f)
        name = os.path.basename(f)[:-4]
        fens, fes, u, rule, coef, form, et, kw = build_case(fe, CASES[name])
        got, _ = gpu_csc(fe, form, fes, fens, u, rule, coef, **kw)
        assert_parity((g["colptr"], g["rowval"], g["nzval"]), got)


@pytest.mark.parametrize("pinned", [False, True])
def test_result_transport_large(fe, orc, gpu_ctx, pinned):
    """Results above 1 M nonzeros cross the link through the staged transport (int32 row indices widened by host threads,
    fegpu_transfer.cu); the arrays that arrive must be the oracle's, for pageable and for page-locked destinations."""
    fens, fes = fe.H8block(1.0, 2.0, 3.0, 18, 18, 18)
    _distort(fens)
    u = make_field(fe, fens, 3)
    rule = fe.GaussRule(3, 2)
    C = isotropic_C()
    ref, _ = oracle_csc(orc, "elastic", "H8", fes, fens, u, rule, C)
    nnz, n = ref[2].size, u.nalldofs()
    assert nnz >= (1 << 20)
    if pinned:
        import torch
        out = (torch.empty(n + 1, dtype=torch.int64, pin_memory=True).numpy(), torch.empty(nnz + 3, dtype=torch.int64, pin_memory=True).numpy()[3:],
               torch.empty(nnz, dtype=torch.float64, pin_memory=True).numpy())
    else:
        out = (np.empty(n + 1, np.int64), np.empty(nnz + 1, np.int64)[1:], np.empty(nnz + 1, np.float64)[1:])  # odd alignment on purpose
    got, a = gpu_csc(fe, "elastic", fes, fens, u, rule, C, out=out)
    assert got[1] is out[1] and got[2] is out[2]
    assert_parity(ref, got)
    vals = np.full(nnz, np.nan)
    a.fetch_values(vals)
    np.testing.assert_array_equal(vals, got[2])


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("variant", ["block", "partition", "skin", "random"])
def test_result_transport_column_stencils(fe, orc, gpu_ctx, pinned, variant):
    """Scalar fields (and everything else that would ship int32 row indices): one id per column + a dictionary of row-offset lists
    cross the link and the host threads rebuild rowval (fe_col_stencils, csrc/fegpu_csc_ops.cu; decoder in fegpu_transfer.cu).  The
    arrays that arrive must be the oracle's bit for bit.  'random': a matrix without a small dictionary (generic assemble! protocol,
    random dofs) must fall back to the int32 transport and still arrive intact."""
    rng = np.random.default_rng(11)
    kw = {}
    if variant == "random":
        n, ne, em = 60000, 20000, 8
        a = fe.SysmatAssemblerSparseGPU(0.0)
        fe.startassembly(a, em, em, ne, n, n)
        dofs = rng.integers(1, n + 1, size=(ne, em))
        mats = rng.standard_normal((ne, em, em))
        for e in range(ne):
            fe.assemble(a, mats[e], dofs[e], dofs[e])
        I = np.repeat(dofs[:, None, :], em, axis=1)   # entry (e, c, r): row dof = dofs[e, r], column-major emission
        J = np.repeat(dofs[:, :, None], em, axis=2)   #                   column dof = dofs[e, c]
        V = np.transpose(mats, (0, 2, 1))
        ref = orc.sparse(I.reshape(-1).astype(np.int64), J.reshape(-1).astype(np.int64), np.ascontiguousarray(V).reshape(-1), n, n)
        nnz = ref[2].size
        assert nnz >= (1 << 20)
        before = gpu_ctx.transfer_stats()["stenciled_results"]
        got = fe.makematrix(a, raw=True)
        assert_parity(ref, got, tol=1e-15)
        assert gpu_ctx.transfer_stats()["stenciled_results"] == before
        return
    if variant == "skin":
        fens, vol = fe.H8block(1.0, 2.0, 3.0, 300, 300, 3)
        fes = fe.meshboundary(vol)
        rule, form, coef, et, kw2 = fe.GaussRule(2, 2), "dot", np.array([[1.3]]), "Q4", dict(m=2)
        okw = dict(m=2, otherdim=1.0)
    else:
        fens, fes = fe.H8block(1.0, 2.0, 3.0, 36, 36, 36)
        _distort(fens)
        rule, form, coef, et, kw2, okw = fe.GaussRule(3, 2), "diffusion", KAPPA3, "H8", {}, {}
    u = make_field(fe, fens, 1)
    ref, _ = oracle_csc(orc, form, et, fes, fens, u, rule, coef, **okw)
    n = u.nalldofs()
    if variant == "partition":
        owner = fe.slab_owner(fens.count(), 3)
        kw = dict(node_owner=owner, my_rank=1)
        keep = np.zeros(n, bool)
        keep[(u.dofnums[owner == 1] - 1).reshape(-1)] = True
        mask = keep[ref[1] - 1]
        col_of = np.repeat(np.arange(n), np.diff(ref[0]))
        cnt = np.bincount(col_of[mask], minlength=n)
        ref = (np.concatenate(([1], 1 + np.cumsum(cnt))).astype(np.int64), ref[1][mask], ref[2][mask])
    nnz = ref[2].size
    big = nnz >= (1 << 20)
    before = gpu_ctx.transfer_stats()["stenciled_results"]
    if pinned:
        import torch
        out = (torch.empty(n + 1, dtype=torch.int64, pin_memory=True).numpy(), torch.empty(nnz + 3, dtype=torch.int64, pin_memory=True).numpy()[3:],
               torch.empty(nnz, dtype=torch.float64, pin_memory=True).numpy())
    else:
        out = (np.empty(n + 1, np.int64), np.empty(nnz + 1, np.int64)[1:], np.empty(nnz + 1, np.float64)[1:])
    for o in out:
        o[...] = -7
    got, a = gpu_csc(fe, form, fes, fens, u, rule, coef, out=out, **kw2, **kw)
    assert_parity(ref, got)
    assert gpu_ctx.transfer_stats()["stenciled_results"] - before == (1 if big else 0)
    assert variant == "partition" or big


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("variant", ["natural", "ebc", "partition", "dot2"])
def test_result_transport_compressed_rows(fe, orc, gpu_ctx, pinned, variant):
    """Vector fields: rowval is rebuilt on the host from the device's neighbour lists + dof map (fegpu_transfer.cu) instead
    of crossing the link; the arrays that arrive must be the oracle's.  'ebc' (free-first numbering: node-major dof order not
    ascending everywhere) must fall back to the int32 transport; a row-block partition and a 2-dof field must not."""
    fens, fes = fe.H8block(1.0, 2.0, 3.0, 22, 22, 22)
    _distort(fens)
    rule = fe.GaussRule(3, 2)
    kw = {}
    if variant == "dot2":
        u = make_field(fe, fens, 2)
        form, coef, et = "dot", np.array([[2.0, 0.5], [0.25, 3.0]]), "H8"
    else:
        fixed = list(range(5, fens.count(), 37)) if variant == "ebc" else None
        u = make_field(fe, fens, 3, fixed_nodes=fixed, fixed_comp=None)
        form, coef, et = "elastic", isotropic_C(), "H8"
    ref, _ = oracle_csc(orc, form, et, fes, fens, u, rule, coef)
    n = u.nalldofs()
    if variant == "partition":
        owner = fe.slab_owner(fens.count(), 2)
        kw = dict(node_owner=owner, my_rank=1)
        keep = np.zeros(n, bool)
        keep[(u.dofnums[owner == 1] - 1).reshape(-1)] = True
        mask = keep[ref[1] - 1]  # the rank's block: the rows of its nodes, every column
        col_of = np.repeat(np.arange(n), np.diff(ref[0]))
        cnt = np.bincount(col_of[mask], minlength=n)
        ref = (np.concatenate(([1], 1 + np.cumsum(cnt))).astype(np.int64), ref[1][mask], ref[2][mask])
    nnz = ref[2].size
    assert nnz >= (1 << 20)
    before = gpu_ctx.transfer_stats()["compressed_results"]
    if pinned:
        import torch
        out = (torch.empty(n + 1, dtype=torch.int64, pin_memory=True).numpy(), torch.empty(nnz + 3, dtype=torch.int64, pin_memory=True).numpy()[3:],
               torch.empty(nnz, dtype=torch.float64, pin_memory=True).numpy())
    else:
        out = (np.empty(n + 1, np.int64), np.empty(nnz + 1, np.int64)[1:], np.empty(nnz + 1, np.float64)[1:])
    for o in out:
        o[...] = -7
    got, a = gpu_csc(fe, form, fes, fens, u, rule, coef, out=out, **kw)
    assert_parity(ref, got)
    used = gpu_ctx.transfer_stats()["compressed_results"] - before
    assert used == (0 if variant == "ebc" else 1)


def test_symm_assembler_testA_on_gpu(fe, orc, gpu_ctx):
    """test/test_basics.jl:119-129 through the generic protocol of the symmetric assembler."""
    from test_oracle_pins import _testA_blocks
    a = fe.SysmatAssemblerSparseSymmGPU(0.0)
    assert a.expectedntriples(5, 5, 3) == 45
    fe.startassembly(a, 5, 5, 3, 7, 7)
    I, J, V = [], [], []
    for m, ii in _testA_blocks():
        fe.assemble(a, m, ii, ii)
        for j in range(4):
            for i in range(j, 4):
                I.append(ii[i]); J.append(ii[j]); V.append(m[i, j])
    got = fe.makematrix(a, raw=True)
    ref = orc.sparse_symm(np.array(I), np.array(J), np.array(V), 7)
    assert_parity(ref, got)
    with pytest.raises(fe.FEGPUError, match="Size mismatch"):
        fe.startassembly(a, 2, 3, 1, 7, 7)
        fe.assemble(a, np.zeros((2, 3)), [1, 2], [1, 2, 3])


@pytest.mark.parametrize("form,et,ndn", [("diffusion", "H8", 1), ("elastic", "H8", 3), ("dot", "T10", 1), ("elastic", "T4", 3)])
def test_symm_assembler_forms(fe, orc, gpu_ctx, form, et, ndn):
    """The reference's default assembler (SysmatAssemblerSparseSymm, FEMMBaseModule.jl:1374,1408,1543,1822) on a distorted
    mesh (no exact cancellations): same pattern as the oracle's S + transpose(S), values within tolerance; on an axis-aligned
    block the exact zeros must disappear from the pattern."""
    fens, fes = _mesh(fe, et, 3)
    _distort(fens)
    u = make_field(fe, fens, ndn)
    rule = _rule(fe, et)
    coef = {"diffusion": KAPPA3, "elastic": isotropic_C(), "dot": np.array([[1.3]])}[form]
    _, (I, J, V) = oracle_csc(orc, form, et, fes, fens, u, rule, coef)
    low = orc.lower_triangle_mask(fes.nne * ndn, fes.count())
    ref = orc.sparse_symm(I[low], J[low], V[low], u.nalldofs())
    a = fe.SysmatAssemblerSparseSymmGPU(0.0)
    got, _ = gpu_csc(fe, form, fes, fens, u, rule, coef, assembler=a)
    assert_parity(ref, got)
    A = orc.to_scipy(got[0], got[1], got[2], got[3], got[4])
    assert abs(A - A.T).max() == 0.0


def test_symm_assembler_drops_exact_zeros(fe, orc, gpu_ctx):
    """Entries that sum to exactly 0.0 are stored by SysmatAssemblerSparse but not by the symmetric assembler.  A conductivity
    with a zero row/column (kappa = diag(1, 0, 0)) on axis-aligned bricks decouples nodes that differ only in y or z: exact zeros."""
    fens, fes = fe.H8block(2.0, 2.0, 2.0, 3, 3, 3)
    u = make_field(fe, fens, 1)
    rule = fe.GaussRule(3, 2)
    for kap in (np.zeros((3, 3)), np.diag([1.0, 0.0, 0.0])):
        full, _ = gpu_csc(fe, "diffusion", fes, fens, u, rule, kap)
        symm, _ = gpu_csc(fe, "diffusion", fes, fens, u, rule, kap, assembler=fe.SysmatAssemblerSparseSymmGPU(0.0))
        nzero = int((full[2] == 0.0).sum())
        assert symm[2].size == full[2].size - nzero and (symm[2] != 0.0).all()
        A, B = orc.to_scipy(*full[:3], full[3], full[4]), orc.to_scipy(*symm[:3], symm[3], symm[4])
        assert abs(A - B).max() == 0.0
        if not kap.any():
            assert nzero == full[2].size and symm[2].size == 0 and (symm[0] == 1).all()


def test_ffblock_assembler_and_matrix_blocked(fe, orc, gpu_ctx):
    """test/test_forms.jl:542-566: SysmatAssemblerFFBlock(nfreedofs) == matrix_blocked_ff(full assembly); all four blocks
    against the oracle's Julia-range-indexing restatement."""
    fens, fes = fe.H8block(1.0, 2.0, 3.0, 5, 4, 3)
    _distort(fens)
    u = make_field(fe, fens, 3, fixed_nodes=np.array([1, 2, 3, 17, 40]), fixed_comp=None)
    nf, n = u.nfreedofs(), u.nalldofs()
    assert 0 < nf < n
    rule = fe.GaussRule(3, 2)
    C = isotropic_C()
    ref, _ = oracle_csc(orc, "elastic", "H8", fes, fens, u, rule, C)
    a = fe.SysmatAssemblerSparseGPU(0.0)
    full, _ = gpu_csc(fe, "elastic", fes, fens, u, rule, C, assembler=a)
    assert_parity(ref, full)
    blocks = {"ff": (1, nf, 1, nf), "fd": (1, nf, nf + 1, n), "df": (nf + 1, n, 1, nf), "dd": (nf + 1, n, nf + 1, n)}
    for name, rng in blocks.items():
        got = getattr(fe, "matrix_blocked_" + name)(a, nf, nf, raw=True)
        assert (got[3], got[4]) == (rng[1] - rng[0] + 1, rng[3] - rng[2] + 1)
        assert_parity(orc.matrix_block(ref, *rng), got)
    again = a.makematrix(raw=True)                      # the full matrix is still there after cutting blocks
    assert_parity(ref, again)
    ff = fe.SysmatAssemblerFFBlock(nf)
    got, _ = gpu_csc(fe, "elastic", fes, fens, u, rule, C, assembler=ff)
    assert_parity(orc.matrix_block(ref, 1, nf, 1, nf), got)
    # generic protocol through the wrapper (AssemblyModule.jl:1169-1231)
    ff2 = fe.SysmatAssemblerFFBlock(3, 2)
    fe.startassembly(ff2, 2, 2, 2, 4, 4)
    fe.assemble(ff2, np.array([[1.0, 2.0], [3.0, 4.0]]), [1, 4], [1, 2])
    fe.assemble(ff2, np.array([[5.0, 0.0], [7.0, 8.0]]), [2, 3], [2, 3])
    got = fe.makematrix(ff2, raw=True)
    I, J, V = np.array([1, 4, 1, 4, 2, 3, 2, 3]), np.array([1, 1, 2, 2, 2, 2, 3, 3]), np.array([1.0, 3.0, 2.0, 4.0, 5.0, 7.0, 0.0, 8.0])
    assert_parity(orc.matrix_block(orc.sparse(I, J, V, 4, 4), 1, 3, 1, 2), got)
    with pytest.raises(fe.FEGPUError, match="too many rows"):
        fe.matrix_blocked_ff(a, n + 1)


@pytest.mark.parametrize("et,rule_order", [("H8", 3), ("T10", None), ("H20", None), ("T4", None)])
def test_elastic_sort_path_full_layout(fe, orc, gpu_ctx, et, rule_order):
    """Elasticity through the generic sort path (a dof shared by two nodes rules out the mesh-structured pattern): the
    integration kernels then write full element matrices in emission order, including H8 with a rule the dedicated H8
    kernel does not take (3x3x3)."""
    fens, fes = _mesh(fe, et, 2)
    _distort(fens)
    u = make_field(fe, fens, 3)
    u.dofnums = u.dofnums.copy()
    u.dofnums[u.dofnums == u.nalldofs()] = 2
    rule = fe.GaussRule(3, rule_order) if rule_order else _rule(fe, et)
    C = isotropic_C()
    C[1, 4] = C[4, 1] = -0.07
    ref, _ = oracle_csc(orc, "elastic", et, fes, fens, u, rule, C)
    got, _ = gpu_csc(fe, "elastic", fes, fens, u, rule, C)
    assert_parity(ref, got)


def test_elastic_h8_gauss3_structured_path(fe, orc, gpu_ctx):
    """H8 with GaussRule(3,3): the register-tiled elasticity kernel writing the compact layout."""
    fens, fes = _mesh(fe, "H8", 3)
    _distort(fens)
    u = make_field(fe, fens, 3)
    rule = fe.GaussRule(3, 3)
    ref, (I, J, V) = oracle_csc(orc, "elastic", "H8", fes, fens, u, rule, isotropic_C())
    got, a = gpu_csc(fe, "elastic", fes, fens, u, rule, isotropic_C())
    assert_parity(ref, got)
    gi, gj, gv = a.coo()                       # raw-COO export expands the compact layout back to emission order
    assert np.array_equal(gi, I) and np.array_equal(gj, J)
    assert np.abs(gv - V).max() <= 1e-12 * np.abs(V).max()


@pytest.mark.parametrize("et", ["Q4", "T3"])
def test_planar_dot_parity(fe, orc, gpu_ctx, et):
    """Mass matrix of a planar (sdim = 2) mesh, m = 2 and m = 3 with a non-unit other dimension (thickness)."""
    fens, fes = (fe.Q4block(2.0, 1.0, 5, 4) if et == "Q4" else fe.T3block(2.0, 1.0, 5, 4))
    fens.xyz[:, 0] += 0.03 * np.sin(3 * fens.xyz[:, 1])
    u = make_field(fe, fens, 1)
    rule = fe.GaussRule(2, 2) if et == "Q4" else fe.TriRule(3)
    c = np.array([[2.5]])
    for m in (2, 3):
        ref, _ = oracle_csc(orc, "dot", et, fes, fens, u, rule, c, m=m, otherdim=1.0)
        got, _ = gpu_csc(fe, "dot", fes, fens, u, rule, c, m=m)
        assert_parity(ref, got)


def _planar_mesh(fe, et):
    fens, fes = (fe.Q4block(2.0, 3.0, 5, 4) if et == "Q4" else fe.T3block(2.0, 3.0, 5, 4))
    x = fens.xyz
    x[:, 0] += 0.04 * np.sin(3 * x[:, 1])
    x[:, 1] += 0.04 * np.sin(2 * x[:, 0])
    return fens, fes


@pytest.mark.parametrize("et", ["H8", "H20", "T4", "T10", "Q4", "T3"])
def test_convection_parity(fe, orc, gpu_ctx, et):
    """bilform_convection (FEMMBaseModule.jl:1583-1625): non-symmetric scalar form with a nodal velocity
    field; pattern bit-exact, values within 1e-12; plus the reference's own identity Psi' K Q = (u . grad q) V for linear q
    (test/test_forms.jl:196-231)."""
    planar = et in ("Q4", "T3")
    if planar:
        fens, fes = _planar_mesh(fe, et)
        rule = fe.GaussRule(2, 2) if et == "Q4" else fe.TriRule(3)
    else:
        fens, fes = _mesh(fe, et)
        _distort(fens)
        rule = _rule(fe, et)
    sdim = fens.xyz.shape[1]
    q = make_field(fe, fens, 1)
    x = fens.xyz
    uvals = np.stack([3.1 + 0.2 * x[:, 1], -2.7 + 0.1 * x[:, 0], 0.4 - 0.3 * x[:, 0]][:sdim], axis=1)
    ref, _ = oracle_csc(orc, "convection", et, fes, fens, q, rule, uvals)
    got, _ = gpu_csc(fe, "convection", fes, fens, q, rule, uvals)
    assert_parity(ref, got)
    # identity with a constant velocity and linear q, on the undistorted block (known volume)
    uc = np.tile([3.1, -2.7, 0.4][:sdim], (fens.count(), 1))
    if not planar:
        import scipy.sparse as sp
        fens2, fes2 = _mesh(fe, et)
        q2 = make_field(fe, fens2, 1)
        got, _ = gpu_csc(fe, "convection", fes2, fens2, q2, rule, uc)
        n = q2.nalldofs()
        K = sp.csc_matrix((got[2], got[1] - 1, got[0] - 1), shape=(n, n))
        grad = np.array([0.3, 0.4, 0.5])
        Q = np.zeros(n)
        Q[q2.dofnums[:, 0] - 1] = -0.1 + fens2.xyz @ grad
        vol = 1.3 * 3.1 * 2.7
        assert abs(np.ones(n) @ (K @ Q) - (grad @ uc[0]) * vol) <= 1e-9 * vol
    with pytest.raises(fe.FEGPUError):  # "must be able to assemble unsymmetric matrices", FEMMBaseModule.jl:1634
        gpu_csc(fe, "convection", fes, fens, q, rule, uc, assembler=fe.SysmatAssemblerSparseSymmGPU(0.0))


@pytest.mark.parametrize("et", ["H8", "H20", "H27", "T4", "T10", "Q4", "T3"])
def test_div_grad_parity(fe, orc, gpu_ctx, et):
    """bilform_div_grad (FEMMBaseModule.jl:1672-1713): parity with the oracle; the reference's identities v' G v = 0 for a
    constant field and = 2 mu V |sym grad u|^2 for a linear one (test/test_forms.jl:234-300) on the undistorted hexahedra."""
    planar = et in ("Q4", "T3")
    if planar:
        fens, fes = _planar_mesh(fe, et)
        rule = fe.GaussRule(2, 2) if et == "Q4" else fe.TriRule(3)
    else:
        fens, fes = _mesh(fe, et, 2)
        _distort(fens)
        rule = _rule(fe, et)
    sdim = fens.xyz.shape[1]
    u = make_field(fe, fens, sdim)
    mu = 0.13377
    ref, _ = oracle_csc(orc, "div_grad", et, fes, fens, u, rule, mu)
    got, _ = gpu_csc(fe, "div_grad", fes, fens, u, rule, mu)
    assert_parity(ref, got)
    import scipy.sparse as sp
    n = u.nalldofs()
    G = sp.csc_matrix((got[2], got[1] - 1, got[0] - 1), shape=(n, n))
    vc = np.zeros(n)
    vc[u.dofnums - 1] = np.array([3.1, -2.7, -0.77][:sdim])
    assert abs(vc @ (G @ vc)) <= 1e-9 * np.abs(got[2]).max() * n
    assert abs(G - G.T).max() <= 1e-14 * np.abs(got[2]).max()  # symmetric to rounding ((factor*g_a)*g_b vs (factor*g_b)*g_a), as in the reference
    if et == "H8":
        W, L, t = 11.1, 12.0, 7.32
        fens, fes = fe.H8block(L, W, t, 2, 4, 3)
        u = make_field(fe, fens, 3)
        a, b, c, d = (-0.33, 2 / 3, -1.67, 2 / 7)
        x = fens.xyz
        vals = np.stack([a + b * x[:, 0] + c * x[:, 1] + d * x[:, 2], b + c * x[:, 0] + d * x[:, 1] + a * x[:, 2],
                         c + d * x[:, 0] + a * x[:, 1] + b * x[:, 2]], 1)
        got, _ = gpu_csc(fe, "div_grad", fes, fens, u, fe.GaussRule(3, 2), mu)
        n = u.nalldofs()
        G = sp.csc_matrix((got[2], got[1] - 1, got[0] - 1), shape=(n, n))
        v = np.zeros(n)
        v[u.dofnums - 1] = vals
        gradu = np.array([[b, c, d], [c, d, a], [d, a, b]])
        gs = (gradu + gradu.T) / 2
        true = 2 * mu * W * L * t * (gs ** 2).sum()
        assert abs(v @ (G @ v) - true) <= 1e-5 * true


@pytest.mark.parametrize("et,ndn", [("H8", 1), ("H8", 3), ("H20", 3), ("H27", 1), ("T4", 3), ("T10", 1), ("T10", 3)])
def test_linform_dot_volume_parity(fe, orc, gpu_ctx, et, ndn):
    """linform_dot / distribloads (FEMMBaseModule.jl:1207-1297) into a SysvecAssembler: every entry of the assembled vector
    against the oracle (same element order of the sums: tolerance 1e-12 of the largest entry is rounding from FMA only)."""
    fens, fes = _mesh(fe, et)
    _distort(fens)
    rule = _rule(fe, et)
    P = make_field(fe, fens, ndn, fixed_nodes=[2, 5] if ndn == 3 else None, fixed_comp=None)
    force = np.array([11.0, -3.5, 0.25][:ndn])
    ref = orc.linform_dot(et, fes.conn, fens.xyz, P.dofnums, P.nalldofs(), rule.param_coords, rule.weights, force)
    femm = fe.FEMMBase(fe.IntegDomain(fes, rule))
    a = fe.SysvecAssemblerGPU(0.0)
    F = fe.linform_dot(femm, a, fe.NodalField(fens.xyz), P, fe.DataCache(force), 3)
    assert F.shape == ref.shape
    assert np.abs(F - ref).max() <= 1e-12 * np.abs(ref).max()
    F2 = fe.distribloads(femm, a, fe.NodalField(fens.xyz), P, fe.ForceIntensity(force), 3)
    np.testing.assert_array_equal(F, F2)  # repeated assembly is bit-identical (no atomics)


@pytest.mark.parametrize("et", ["Q4", "T3"])
def test_linform_dot_surface_traction(fe, orc, gpu_ctx, et):
    """Surface tractions: the boundary of a block (Q4 / T3 in 3-D), m = 2; sum(F) = traction x area."""
    if et == "Q4":
        fens, vol = fe.H8block(1.0, 2.0, 3.0, 3, 4, 5)
        rule = fe.GaussRule(2, 2)
    else:
        fens, vol = fe.T4block(1.0, 2.0, 3.0, 3, 4, 5)
        rule = fe.TriRule(3)
    bfes = fe.meshboundary(vol)
    P = make_field(fe, fens, 3)
    force = np.array([2.0, -1.0, 0.5])
    ref = orc.linform_dot(et, bfes.conn, fens.xyz, P.dofnums, P.nalldofs(), rule.param_coords, rule.weights, force, m=2)
    femm = fe.FEMMBase(fe.IntegDomain(bfes, rule))
    F = fe.linform_dot(femm, fe.SysvecAssemblerGPU(0.0), fe.NodalField(fens.xyz), P, fe.DataCache(force), 2)
    assert np.abs(F - ref).max() <= 1e-12 * np.abs(ref).max()
    area = 2 * (1 * 2 + 2 * 3 + 1 * 3)
    comp = F[P.dofnums - 1].sum(axis=0)
    assert np.abs(comp - force * area).max() <= 1e-10 * area
    with pytest.raises(fe.FEGPUError):
        fe.linform_dot(femm, fe.SysvecAssemblerGPU(0.0), fe.NodalField(fens.xyz), P, fe.DataCache(force), 1)


def test_distribloads_reference_identity_and_partition(fe, orc, gpu_ctx):
    """test/test_forms.jl:158-193: sum(F) = L W t f; and the row blocks of a 2-way partition add up to the vector."""
    W, L, t = 1.1, 12.0, 4.32
    fens, fes = fe.H8block(L, W, t, 2, 4, 3)
    psi = make_field(fe, fens, 1)
    femm = fe.FEMMBase(fe.IntegDomain(fes, fe.GaussRule(3, 2)))
    geom = fe.NodalField(fens.xyz)
    F = fe.distribloads(femm, fe.SysvecAssemblerGPU(0.0), geom, psi, fe.ForceIntensity([11.0]), 3)
    assert abs(F.sum() - L * W * t * 11.0) / 667 <= 1.0e-5
    F2 = fe.distribloads(femm, fe.SysvecAssemblerGPU(0.0), geom, psi, fe.ForceIntensity(11.0), 3)
    np.testing.assert_array_equal(F, F2)
    owner = fe.slab_owner(fens.count(), 2)
    parts = [fe.linform_dot(femm, fe.SysvecAssemblerGPU(0.0), geom, psi, fe.DataCache(np.array([11.0])), 3, node_owner=owner, my_rank=p)
             for p in range(2)]
    for p in range(2):
        notmine = np.ones(psi.nalldofs(), bool)
        notmine[psi.dofnums[owner == p, 0] - 1] = False
        assert not parts[p][notmine].any()
    np.testing.assert_array_equal(parts[0] + parts[1], F)


def test_sysvec_assembler_protocol(fe, gpu_ctx):
    """startassembly!/assemble!/makevector! of SysvecAssembler (AssemblyModule.jl:884-917) with the reference's range errors."""
    a = fe.SysvecAssemblerGPU(0.0)
    a.startassembly(7)
    a.assemble(np.array([1.0, 2.0, 3.0]), [5, 2, 1])
    a.assemble(np.array([10.0, 20.0]), [2, 7])
    a.assemble(np.array([0.5]), [2])
    F = fe.makevector(a)
    np.testing.assert_array_equal(F, [3.0, 12.5, 0.0, 0.0, 1.0, 0.0, 20.0])
    a.startassembly(3)
    with pytest.raises(fe.FEGPUError, match="Row degree of freedom < 1"):
        a.assemble(np.array([1.0]), [0])
    with pytest.raises(fe.FEGPUError, match="Row degree of freedom > size"):
        a.assemble(np.array([1.0]), [4])
    a.startassembly(3)
    np.testing.assert_array_equal(fe.makevector(a), [0.0, 0.0, 0.0])


@pytest.mark.parametrize("kind", ["diag", "hrz"])
@pytest.mark.parametrize("et,ndn", [("H8", 3), ("T10", 1), ("H20", 3), ("T4", 3)])
def test_lumped_mass_assemblers(fe, orc, gpu_ctx, kind, et, ndn):
    """SysmatAssemblerSparseDiag / SysmatAssemblerSparseHRZLumpingSymm (AssemblyModule.jl:599-794, 943-1141) fed by bilform_dot:
    pattern (the diagonal entries of the dofs that appear in an element) bit-exact, values within 1e-12; HRZ preserves the mass:
    sum(M_lumped) = ndn * rho * V."""
    fens, fes = _mesh(fe, et)
    _distort(fens)
    rule = _rule(fe, et)
    u = make_field(fe, fens, ndn, fixed_nodes=[3, 4] if ndn == 3 else None, fixed_comp=None)
    c = 2.5 * np.eye(ndn)
    n = u.nalldofs()
    I, J, V = orc.bilform_dot_coo(et, fes.conn, fens.xyz, u.dofnums, n, rule.param_coords, rule.weights, c)
    Id, Vd = orc.lumped_coo(I, J, V, fes.nne * ndn, 1 if kind == "diag" else 2)
    ref = orc.sparse(Id, Id, Vd, n, n)
    a = fe.SysmatAssemblerSparseDiagGPU(0.0) if kind == "diag" else fe.SysmatAssemblerSparseHRZLumpingSymmGPU(0.0)
    assert a.expectedntriples(24, 24, 10) == 240
    got, _ = gpu_csc(fe, "dot", fes, fens, u, rule, c, assembler=a)
    assert_parity(ref, got)
    assert np.array_equal(np.repeat(np.arange(1, n + 1), np.diff(got[0])), got[1])  # diagonal
    if kind == "hrz":
        full = orc.sparse(I, J, V, n, n)
        assert abs(got[2].sum() - full[2].sum()) <= 1e-10 * abs(full[2].sum())


def test_lumped_assemblers_generic_protocol(fe, orc, gpu_ctx):
    """startassembly!/assemble!/makematrix! of the diagonal and HRZ assemblers with host element matrices, and their errors."""
    rng = np.random.default_rng(5)
    mats = [rng.random((k, k)) + k * np.eye(k) for k in (4, 4, 3)]
    dofs = [[1, 4, 6, 2], [2, 3, 7, 1], [7, 5, 1]]
    for cls, mode in ((fe.SysmatAssemblerSparseDiagGPU, 1), (fe.SysmatAssemblerSparseHRZLumpingSymmGPU, 2)):
        a = cls(0.0)
        a.startassembly(4, 4, 3, 7, 7)
        for m, d in zip(mats, dofs):
            a.assemble(m, d, d)
        colptr, rowval, nzval, mm, nn = a.makematrix(raw=True)
        I = np.concatenate([np.asarray(d, np.int64) for d in dofs])
        V = np.concatenate([np.diag(m) * ((m.sum() / np.trace(m)) if mode == 2 else 1.0) for m in mats])
        ref = orc.sparse(I, I, V, 7, 7)
        assert (mm, nn) == (7, 7)
        np.testing.assert_array_equal(colptr, ref[0])
        np.testing.assert_array_equal(rowval, ref[1])
        assert np.abs(nzval - ref[2]).max() <= 1e-13 * np.abs(ref[2]).max()
        with pytest.raises(fe.FEGPUError, match="Size mismatch"):
            a.startassembly(4, 4, 1, 7, 7)
            a.assemble(np.ones((4, 3)), [1, 2, 3, 4], [1, 2, 3])
        b = cls(0.0)
        with pytest.raises(fe.FEGPUError, match="square matrices"):
            b.startassembly(4, 3, 1, 7, 7)
        with pytest.raises(fe.FEGPUError, match="Row and column info do not agree"):
            b.startassembly(4, 4, 1, 7, 8)


@pytest.mark.parametrize("et,ndn,m", [("H8", 1, 3), ("T10", 3, 3), ("H20", 1, 3), ("Q4", 1, 2), ("T3", 2, 2)])
def test_masslike_parity(fe, orc, gpu_ctx, et, ndn, m):
    """bilform_masslike (FEMMBaseModule.jl:1865-1912): rectangular matrix with element-numbered rows; pattern bit-exact, values
    within 1e-12; row sums of a scalar field are c times the element measures."""
    if et in ("Q4", "T3"):
        fens, vol = (fe.H8block if et == "Q4" else fe.T4block)(1.0, 2.0, 3.0, 3, 2, 2)
        fes = fe.meshboundary(vol)
        rule = fe.GaussRule(2, 2) if et == "Q4" else fe.TriRule(3)
    else:
        fens, fes = _mesh(fe, et)
        _distort(fens)
        rule = _rule(fe, et)
    phi = make_field(fe, fens, ndn)
    c = np.array([[2.0, 0.3, -0.1], [0.0, 1.5, 0.2], [0.4, 0.0, 3.0]])[:ndn, :ndn]
    n = phi.nalldofs()
    I, J, V = orc.bilform_masslike_coo(et, fes.conn, fens.xyz, phi.dofnums, n, rule.param_coords, rule.weights, c, m=m)
    ref = orc.sparse(I, J, V, fes.count() * ndn, n)
    femm = fe.FEMMBase(fe.IntegDomain(fes, rule))
    got = fe.bilform_masslike(femm, fe.SysmatAssemblerSparseGPU(0.0), fe.NodalField(fens.xyz), phi, fe.DataCache(c), m=m, raw=True)
    assert (got[3], got[4]) == (fes.count() * ndn, n)
    assert_parity(ref, got)
    if ndn == 1:
        import scipy.sparse as sp
        M = sp.csc_matrix((got[2], got[1] - 1, got[0] - 1), shape=(got[3], got[4]))
        measure = np.asarray(M.sum(axis=1)).reshape(-1) / c[0, 0]
        total = {"H8": 1.3 * 3.1 * 2.7, "H20": 1.3 * 3.1 * 2.7, "Q4": 2 * (2 + 6 + 3)}.get(et)
        assert (measure > 0).all()
        if total is not None and et == "Q4":
            assert abs(measure.sum() - total) <= 1e-10 * total


@pytest.mark.parametrize("et", ["Q4", "T3"])
def test_planar_forms_with_thickness(fe, orc, gpu_ctx, et):
    """A constant other-dimension (IntegDomain(fes, rule, t), IntegDomainModule.jl:73-82): Jacobianvolume of a 2-manifold is the
    surface Jacobian times the thickness (:504-517).  Diffusion, convection and div_grad against the oracle, and the reference's
    convection identity Psi' K Q = (b u_x + c u_y) W L t (test/test_forms.jl:196-231, there on Q8)."""
    import scipy.sparse as sp
    W, L, t = 6.1, 12.0, 2.32
    fens, fes = (fe.Q4block(L, W, 3, 4) if et == "Q4" else fe.T3block(L, W, 3, 4))
    rule = fe.GaussRule(2, 2) if et == "Q4" else fe.TriRule(3)
    geom = fe.NodalField(fens.xyz)
    femm = fe.FEMMBase(fe.IntegDomain(fes, rule, t))
    q = make_field(fe, fens, 1)
    n = q.nalldofs()
    x = fens.xyz
    # convection
    uv = np.tile([3.1, -2.7], (fens.count(), 1))
    I, J, V = orc.bilform_convection_coo(et, fes.conn, x, uv, q.dofnums, n, rule.param_coords, rule.weights, 1.0, otherdim=t)
    ref = orc.sparse(I, J, V, n, n)
    got = fe.bilform_convection(femm, fe.SysmatAssemblerSparseGPU(0.0), geom, fe.NodalField(uv), q, fe.DataCache(1.0), raw=True)
    assert_parity(ref, got)
    K = sp.csc_matrix((got[2], got[1] - 1, got[0] - 1), shape=(n, n))
    a_, b_, c_ = (-0.1, +0.3, +0.4)
    Q = np.zeros(n)
    Q[q.dofnums[:, 0] - 1] = a_ + b_ * x[:, 0] + c_ * x[:, 1]
    assert abs(np.ones(n) @ (K @ Q) - (b_ * 3.1 + c_ * -2.7) * (W * L * t)) / (W * L * t) <= 1.0e-5
    # diffusion
    kap = np.array([[1.5, 0.2], [0.2, 2.5]])
    I, J, V = orc.bilform_diffusion_coo(et, fes.conn, x, q.dofnums, n, rule.param_coords, rule.weights, kap, otherdim=t)
    got = fe.bilform_diffusion(femm, fe.SysmatAssemblerSparseGPU(0.0), geom, q, fe.DataCache(kap), raw=True)
    assert_parity(orc.sparse(I, J, V, n, n), got)
    I1, J1, V1 = orc.bilform_diffusion_coo(et, fes.conn, x, q.dofnums, n, rule.param_coords, rule.weights, kap)
    assert np.abs(V - t * V1).max() <= 1e-13 * np.abs(V).max()
    # div_grad
    u = make_field(fe, fens, 2)
    I, J, V = orc.bilform_div_grad_coo(et, fes.conn, x, u.dofnums, u.nalldofs(), rule.param_coords, rule.weights, 0.13, otherdim=t)
    got = fe.bilform_div_grad(femm, fe.SysmatAssemblerSparseGPU(0.0), geom, u, fe.DataCache(0.13), raw=True)
    assert_parity(orc.sparse(I, J, V, u.nalldofs(), u.nalldofs()), got)


def _rotation():
    a, b = 0.7, -0.4
    Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    Rx = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
    return Rz @ Rx


@pytest.mark.parametrize("et", ["H8", "T10", "H20", "T4"])
def test_constant_material_csys(fe, orc, gpu_ctx, et):
    """FEMMBase(integdomain, CSys(csmat)) with a constant rotation (CSysModule.jl:133-144): orthotropic elasticity and anisotropic
    diffusion in rotated material axes against the oracle; an isotropic material must not see the rotation; diffusion with
    (Rm, kappa) equals diffusion with (I, Rm kappa Rm')."""
    fens, fes = _mesh(fe, et, 2)
    _distort(fens)
    rule = _rule(fe, et)
    Rm = _rotation()
    geom = fe.NodalField(fens.xyz)
    u = make_field(fe, fens, 3)
    n = u.nalldofs()
    Corth = isotropic_C() + np.diag([3.0, 0.5, 1.0, 0.2, 0.7, 0.1])
    Corth[0, 1] = Corth[1, 0] = 0.9
    femm_r = fe.FEMMBase(fe.IntegDomain(fes, rule), fe.CSys(Rm))
    femm_i = fe.FEMMBase(fe.IntegDomain(fes, rule))
    I, J, V = orc.bilform_lin_elastic_coo(et, fes.conn, fens.xyz, u.dofnums, n, rule.param_coords, rule.weights, Corth, Rm=Rm)
    ref = orc.sparse(I, J, V, n, n)
    got = fe.bilform_lin_elastic(femm_r, fe.SysmatAssemblerSparseGPU(0.0), geom, u, fe.DeforModelRed3D, fe.DataCache(Corth), raw=True)
    assert_parity(ref, got)
    got_id = fe.bilform_lin_elastic(femm_i, fe.SysmatAssemblerSparseGPU(0.0), geom, u, fe.DeforModelRed3D, fe.DataCache(Corth), raw=True)
    assert np.abs(got[2] - got_id[2]).max() > 1e-3 * np.abs(got[2]).max()  # the rotation matters for an orthotropic material
    iso_r = fe.bilform_lin_elastic(femm_r, fe.SysmatAssemblerSparseGPU(0.0), geom, u, fe.DeforModelRed3D, fe.DataCache(isotropic_C()), raw=True)
    iso_i = fe.bilform_lin_elastic(femm_i, fe.SysmatAssemblerSparseGPU(0.0), geom, u, fe.DeforModelRed3D, fe.DataCache(isotropic_C()), raw=True)
    np.testing.assert_array_equal(iso_r[1], iso_i[1])
    assert np.abs(iso_r[2] - iso_i[2]).max() <= 1e-12 * np.abs(iso_i[2]).max()
    q = make_field(fe, fens, 1)
    nq = q.nalldofs()
    I, J, V = orc.bilform_diffusion_coo(et, fes.conn, fens.xyz, q.dofnums, nq, rule.param_coords, rule.weights, KAPPA3, Rm=Rm)
    got = fe.bilform_diffusion(femm_r, fe.SysmatAssemblerSparseGPU(0.0), geom, q, fe.DataCache(KAPPA3), raw=True)
    assert_parity(orc.sparse(I, J, V, nq, nq), got)
    rot = fe.bilform_diffusion(femm_i, fe.SysmatAssemblerSparseGPU(0.0), geom, q, fe.DataCache(Rm @ KAPPA3 @ Rm.T), raw=True)
    assert np.abs(got[2] - rot[2]).max() <= 1e-12 * np.abs(rot[2]).max()
    # the scalar (iso) diffusion path never looks at the coordinate system (FEMMBaseModule.jl:1508-1535)
    s_r = fe.bilform_diffusion(femm_r, fe.SysmatAssemblerSparseGPU(0.0), geom, q, fe.DataCache(1.7), raw=True)
    s_i = fe.bilform_diffusion(femm_i, fe.SysmatAssemblerSparseGPU(0.0), geom, q, fe.DataCache(1.7), raw=True)
    np.testing.assert_array_equal(s_r[2], s_i[2])
