"""The thread-per-node CSC kernels (csrc/fegpu_tile.cu) against the oracle, and the ownership / caching rules around them:
which symbolic path a mesh takes, partitions with contiguous and scattered ownership, many tiles (decoupled look-back),
contexts sharing a device, fresh assemblers re-using the context's cached pattern, results outliving an invalidated pattern."""
import numpy as np
import pytest

from conftest import KAPPA3, isotropic_C
from helpers import assert_parity, gpu_csc, make_field, oracle_csc

pytestmark = pytest.mark.gpu


def _path(fe, ctx, fes, u):
    from finetools_jl_b200 import _lib
    return _lib.lib().fegpu_pattern_path(ctx.device_mesh(fes).dofmap(u))


def _distort(fens, amp=0.07):
    x = fens.xyz.copy()
    z = x[:, 2] if x.shape[1] > 2 else 0.37 * x[:, 0]
    fens.xyz[:, 0] += amp * np.sin(2.1 * x[:, 1] + 0.3) * np.cos(1.7 * z)
    fens.xyz[:, 1] += amp * np.sin(1.3 * z + 0.1) * np.cos(2.3 * x[:, 0])
    if x.shape[1] > 2:
        fens.xyz[:, 2] += amp * np.sin(1.9 * x[:, 0] + 0.2) * np.cos(1.1 * x[:, 1])


CASES = [
    # (element type, mesh builder args, ndn, form, coefficient, rule)
    ("H8", (1.0, 2.0, 3.0, 9, 7, 11), 1, "diffusion", KAPPA3, ("gauss", 3, 2)),
    ("H8", (1.0, 2.0, 3.0, 9, 7, 11), 1, "diffusion", 2.5, ("gauss", 3, 2)),
    ("H8", (1.0, 2.0, 3.0, 6, 7, 5), 3, "elastic", None, ("gauss", 3, 2)),
    ("H8", (1.0, 2.0, 3.0, 6, 7, 5), 2, "dot", np.array([[2.0, 0.5], [0.25, 3.0]]), ("gauss", 3, 2)),
    ("H8", (1.0, 2.0, 3.0, 5, 4, 6), 3, "dot", np.array([[2.0, 0.5, 0.1], [0.25, 3.0, 0.2], [0.3, 0.4, 4.0]]), ("gauss", 3, 2)),
    ("H8", (1.0, 2.0, 3.0, 5, 4, 6), 1, "diffusion", KAPPA3, ("gauss", 3, 3)),   # generic integration kernel, full element matrices
    ("Q4", (2.0, 3.0, 17, 13), 1, "diffusion", np.array([[1.5, 0.2], [0.2, 2.5]]), ("gauss", 2, 2)),
    ("Q4", (2.0, 3.0, 17, 13), 2, "dot", np.array([[2.0, 0.5], [0.25, 3.0]]), ("gauss", 2, 2)),
    ("T3", (2.0, 3.0, 15, 12), 1, "diffusion", np.array([[1.5, 0.2], [0.2, 2.5]]), ("tri", 3)),
    ("T3", (2.0, 3.0, 15, 12), 1, "dot", np.array([[1.0]]), ("tri", 3)),
]


def _build(fe, et, margs):
    mesher = {"H8": fe.H8block, "Q4": fe.Q4block, "T3": fe.T3block}[et]
    fens, fes = mesher(*margs)
    _distort(fens)
    return fens, fes


def _rule(fe, spec):
    return fe.GaussRule(spec[1], spec[2]) if spec[0] == "gauss" else fe.TriRule(spec[1])


@pytest.mark.parametrize("case", range(len(CASES)))
def test_tile_path_parity(fe, orc, gpu_ctx, case):
    et, margs, ndn, form, coef, rspec = CASES[case]
    if coef is None:
        coef = isotropic_C()
        coef[1, 4] = coef[4, 1] = -0.07
    fens, fes = _build(fe, et, margs)
    u = make_field(fe, fens, ndn)
    rule = _rule(fe, rspec)
    kw = {"m": 2} if (form == "dot" and et in ("Q4", "T3")) else {}  # planar meshes: surface Jacobian
    ref, _ = oracle_csc(orc, form, et, fes, fens, u, rule, coef, **kw)
    got, a = gpu_csc(fe, form, fes, fens, u, rule, coef, **kw)
    assert _path(fe, gpu_ctx, fes, u) == 2
    assert_parity(ref, got)
    # cached re-assembly: same pattern arrays, bit-identical values (no atomics anywhere)
    got2, _ = gpu_csc(fe, form, fes, fens, u, rule, coef, assembler=a, **kw)
    assert a.pattern_was_cached()
    np.testing.assert_array_equal(got2[2], got[2])


@pytest.mark.parametrize("ndn,form", [(1, "diffusion"), (3, "elastic")])
@pytest.mark.parametrize("ownership", ["slab", "scattered", "bisection"])
def test_tile_path_partitions(fe, orc, gpu_ctx, ndn, form, ownership):
    """Row-block partitions through the thread-per-node kernels: contiguous ownership (two comparisons per candidate), an owner
    map without any locality (byte map; the node window is the whole mesh) and recursive inertial bisection.  Every rank's block
    is the oracle's rows-of-owned-nodes block; the blocks add up to the matrix."""
    fens, fes = fe.H8block(1.0, 2.0, 3.0, 7, 6, 9)
    _distort(fens)
    u = make_field(fe, fens, ndn)
    rule = fe.GaussRule(3, 2)
    coef = KAPPA3 if form == "diffusion" else isotropic_C()
    ref, (I, J, V) = oracle_csc(orc, form, "H8", fes, fens, u, rule, coef)
    n = u.nalldofs()
    nn = fens.count()
    P = 3 if ownership != "bisection" else 4
    if ownership == "slab":
        owner = fe.slab_owner(nn, P)
    elif ownership == "scattered":
        owner = np.random.default_rng(1).integers(0, P, size=nn).astype(np.int32)
    else:
        owner = (fe.pointpartitioning(fens.xyz, P) - 1).astype(np.int32)
    nnz_sum = 0
    a = fe.SysmatAssemblerSparseGPU(0.0)
    for p in range(P):
        owned = np.zeros(n + 1, bool)
        owned[u.dofnums[owner == p].reshape(-1)] = True
        keep = owned[I]
        blk = orc.sparse(I[keep], J[keep], V[keep], n, n)
        got, _ = gpu_csc(fe, form, fes, fens, u, rule, coef, assembler=a, node_owner=owner, my_rank=p)
        assert _path(fe, gpu_ctx, fes, u) == 2
        np.testing.assert_array_equal(got[0], blk[0])
        np.testing.assert_array_equal(got[1], blk[1])
        if blk[2].size:
            assert np.abs(got[2] - blk[2]).max() <= 1e-12 * np.abs(ref[2]).max()
        nnz_sum += got[2].size
    assert nnz_sum == ref[2].size


def test_tile_path_many_tiles_and_unused_nodes(fe, orc, gpu_ctx):
    """~1 400 tiles of 128 nodes (the decoupled look-back runs over many CTAs) on a mesh whose node set has unused nodes in the
    middle and at the end (columns without entries inside the window), 2 dofs per node."""
    fens, fes = fe.H8block(1.0, 1.0, 1.0, 55, 55, 55)
    _distort(fens, 0.002)
    nn = fens.count()
    # drop a slab of elements in the middle: its interior nodes stay in the node set but belong to no element
    cz = fens.xyz[fes.conn[:, 0] - 1, 2]
    keep = (cz < 0.45) | (cz > 0.6)
    fes = type(fes)(np.ascontiguousarray(fes.conn[keep]))
    xyz = np.vstack([fens.xyz, np.full((37, 3), 9.0)])  # and 37 unused trailing nodes
    fens2 = fe.FENodeSet(xyz)
    u = make_field(fe, fens2, 2)
    rule = fe.GaussRule(3, 2)
    c = np.array([[2.0, 0.5], [0.25, 3.0]])
    ref, _ = oracle_csc(orc, "dot", "H8", fes, fens2, u, rule, c)
    got, a = gpu_csc(fe, "dot", fes, fens2, u, rule, c)
    assert _path(fe, gpu_ctx, fes, u) == 2
    assert_parity(ref, got)
    assert np.count_nonzero(np.diff(got[0]) == 0) >= 2 * 37


def test_internal_element_order_is_invisible(fe, orc, gpu_ctx):
    """The library stores elements in its own order (ascending smallest node id); whatever order the FESet lists them in, the raw
    COO export comes back in the CALLER's emission order (AssemblyModule.jl:261-280) and bilform_masslike numbers its rows by the
    caller's element ids (FEMMBaseModule.jl:1907-1908) -- single GPU and as a rank of a partition."""
    rng = np.random.default_rng(7)
    fens, fes = fe.H8block(1.0, 2.0, 3.0, 5, 4, 6)
    _distort(fens)
    fes = type(fes)(np.ascontiguousarray(fes.conn[rng.permutation(fes.count())]))  # element order unrelated to the node numbering
    rule = fe.GaussRule(3, 2)
    for ndn, form, coef in ((1, "diffusion", KAPPA3), (3, "elastic", isotropic_C())):
        u = make_field(fe, fens, ndn)
        ref, (I, J, V) = oracle_csc(orc, form, "H8", fes, fens, u, rule, coef)
        got, a = gpu_csc(fe, form, fes, fens, u, rule, coef)
        assert_parity(ref, got)
        gI, gJ, gV = a.coo()
        np.testing.assert_array_equal(gI, I)
        np.testing.assert_array_equal(gJ, J)
        assert np.abs(gV - V).max() <= 1e-12 * np.abs(V).max()
    # masslike: element-numbered rows
    phi = make_field(fe, fens, 1)
    c = np.array([[1.7]])
    I, J, V = orc.bilform_masslike_coo("H8", fes.conn, fens.xyz, phi.dofnums, phi.nalldofs(), rule.param_coords, rule.weights, c)
    refm = orc.sparse(I, J, V, fes.count(), phi.nalldofs())
    am = fe.SysmatAssemblerSparseGPU(0.0)
    gotm = fe.bilform_masslike(fe.FEMMBase(fe.IntegDomain(fes, rule)), am, fe.NodalField(fens.xyz), phi, fe.DataCache(c), raw=True)
    assert_parity(refm, gotm)


@pytest.mark.parametrize("et,ndn", [("H8", 4), ("H8", 6), ("T10", 5), ("Q4", 4)])
def test_dot_more_than_three_dofs(fe, orc, gpu_ctx, et, ndn):
    """bilform_dot has no cap on the dofs per node in the reference (FEMMBaseModule.jl:1355-1360): 4..6 (shell-like fields) go through
    the Kronecker path of the integration and the runtime-ndn symbolic / numeric kernels."""
    rng = np.random.default_rng(ndn)
    if et == "H8":
        fens, fes = fe.H8block(1.0, 2.0, 3.0, 4, 3, 5)
        rule, kw = fe.GaussRule(3, 2), {}
    elif et == "T10":
        fens, fes = fe.T10block(1.0, 2.0, 3.0, 2, 3, 2)
        rule, kw = fe.TetRule(4), {}
    else:
        fens, fes = fe.Q4block(2.0, 3.0, 7, 5)
        rule, kw = fe.GaussRule(2, 2), {"m": 2}
    _distort(fens, 0.03)
    u = make_field(fe, fens, ndn)
    c = rng.standard_normal((ndn, ndn))
    ref, _ = oracle_csc(orc, "dot", et, fes, fens, u, rule, c, **kw)
    got, _ = gpu_csc(fe, "dot", fes, fens, u, rule, c, **kw)
    assert_parity(ref, got)


def test_adjacency_placement_collisions_are_caught(fe, orc, gpu_ctx):
    """The adjacency table is first filled without atomics, an element's entry going to the plane of the node's LOCAL index; that is
    collision free only when the elements around a node see it at different local indices.  Rotating the connectivity of random
    elements about their axis (a valid H8 renumbering: same geometry, positive Jacobian) makes neighbours claim the same plane: the
    build must notice (entry count) and redo the table with atomics.  Fresh, cached, and as a rank of a partition."""
    rng = np.random.default_rng(11)
    fens, fes = fe.H8block(1.0, 2.0, 3.0, 9, 8, 7)
    _distort(fens, 0.04)
    conn = fes.conn.copy()
    rot = np.array([1, 2, 3, 0, 5, 6, 7, 4])
    for e in rng.choice(conn.shape[0], size=conn.shape[0] // 3, replace=False):
        for _ in range(int(rng.integers(1, 4))):
            conn[e] = conn[e][rot]
    fes2 = type(fes)(np.ascontiguousarray(conn))
    rule = fe.GaussRule(3, 2)
    for ndn, form, coef in ((1, "diffusion", KAPPA3), (3, "elastic", isotropic_C())):
        u = make_field(fe, fens, ndn)
        ref, (I, J, V) = oracle_csc(orc, form, "H8", fes2, fens, u, rule, coef)
        got, a = gpu_csc(fe, form, fes2, fens, u, rule, coef)
        assert _path(fe, gpu_ctx, fes2, u) == 2
        assert_parity(ref, got)
        a.invalidate_patterns()
        got, _ = gpu_csc(fe, form, fes2, fens, u, rule, coef, assembler=a)   # the collision is remembered: straight to the atomics
        assert_parity(ref, got)
        owner = fe.slab_owner(fens.count(), 2)
        n = u.nalldofs()
        owned = np.zeros(n + 1, bool)
        owned[u.dofnums[owner == 1].reshape(-1)] = True
        blk = orc.sparse(I[owned[I]], J[owned[I]], V[owned[I]], n, n)
        g, _ = gpu_csc(fe, form, fes2, fens, u, rule, coef, assembler=a, node_owner=owner, my_rank=1)
        np.testing.assert_array_equal(g[0], blk[0])
        np.testing.assert_array_equal(g[1], blk[1])
        assert np.abs(g[2] - blk[2]).max() <= 1e-12 * np.abs(ref[2]).max()


def test_general_path_still_taken_when_preconditions_fail(fe, orc, gpu_ctx):
    """Free-first / fixed-last numbering (not affine) and a permuted dof map fall back to the group kernels -- same arrays."""
    fens, fes = fe.H8block(1.0, 2.0, 3.0, 6, 5, 7)
    _distort(fens)
    rule = fe.GaussRule(3, 2)
    u = make_field(fe, fens, 1, fixed_nodes=[1, 2, 3, 50, 51], fixed_comp=None)
    ref, _ = oracle_csc(orc, "diffusion", "H8", fes, fens, u, rule, KAPPA3)
    got, _ = gpu_csc(fe, "diffusion", fes, fens, u, rule, KAPPA3)
    assert _path(fe, gpu_ctx, fes, u) == 1
    assert_parity(ref, got)
    u2 = fe.NodalField(np.zeros((fens.count(), 1)))
    u2.dofnums = (np.random.default_rng(0).permutation(fens.count()) + 1).reshape(-1, 1).astype(np.int64)
    ref, _ = oracle_csc(orc, "diffusion", "H8", fes, fens, u2, rule, KAPPA3)
    got, _ = gpu_csc(fe, "diffusion", fes, fens, u2, rule, KAPPA3)
    assert _path(fe, gpu_ctx, fes, u2) == 1
    assert_parity(ref, got)


def test_high_valence_mesh_falls_back(fe, orc, gpu_ctx):
    """T4 blocks have up to 24+ elements at a node: above the thread-per-node capacity (16) -> general path, remembered."""
    fens, fes = fe.T4block(1.0, 1.0, 1.0, 4, 4, 4)
    u = make_field(fe, fens, 1)
    rule = fe.TetRule(4)
    ref, _ = oracle_csc(orc, "diffusion", "T4", fes, fens, u, rule, KAPPA3)
    got, a = gpu_csc(fe, "diffusion", fes, fens, u, rule, KAPPA3)
    assert _path(fe, gpu_ctx, fes, u) == 1
    assert_parity(ref, got)
    a.invalidate_patterns()
    got, _ = gpu_csc(fe, "diffusion", fes, fens, u, rule, KAPPA3, assembler=a)
    assert _path(fe, gpu_ctx, fes, u) == 1
    assert_parity(ref, got)


def test_fresh_assembler_per_call_reuses_the_context_cache(fe, orc, gpu_ctx):
    """The reference builds an assembler per call (FEMMBaseModule.jl:1374, 1408): the device mesh and the cached pattern belong
    to the context, so the second call -- with a NEW assembler -- is a cached re-assembly, and nothing is uploaded twice."""
    fens, fes = fe.H8block(1.0, 2.0, 3.0, 8, 8, 8)
    u = make_field(fe, fens, 3)
    rule = fe.GaussRule(3, 2)
    ref, _ = oracle_csc(orc, "elastic", "H8", fes, fens, u, rule, isotropic_C())
    got1, a1 = gpu_csc(fe, "elastic", fes, fens, u, rule, isotropic_C())
    assert not a1.pattern_was_cached()
    dm = gpu_ctx.device_mesh(fes)
    got2, a2 = gpu_csc(fe, "elastic", fes, fens, u, rule, isotropic_C())
    assert a2 is not a1 and a2.pattern_was_cached()
    assert gpu_ctx.device_mesh(fes) is dm and len(dm.dofmaps) == 1
    assert_parity(ref, got2)
    np.testing.assert_array_equal(got1[2], got2[2])
    # an edited numbering (new array) is NOT served from the cache
    u.dofnums = u.dofnums.copy()
    u.dofnums[[0, 1], :] = u.dofnums[[1, 0], :]
    ref3, _ = oracle_csc(orc, "elastic", "H8", fes, fens, u, rule, isotropic_C())
    got3, a3 = gpu_csc(fe, "elastic", fes, fens, u, rule, isotropic_C())
    assert not a3.pattern_was_cached()
    assert_parity(ref3, got3)
    # a different rule on the same FESet re-uploads the tables (stiffness with 2x2x2, then 3x3x3)
    rule3 = fe.GaussRule(3, 3)
    ref4, _ = oracle_csc(orc, "elastic", "H8", fes, fens, u, rule3, isotropic_C())
    got4, _ = gpu_csc(fe, "elastic", fes, fens, u, rule3, isotropic_C())
    assert_parity(ref4, got4)
    # an edited owner map (same object, same rank, edited in place) is noticed
    owner = np.zeros(fens.count(), np.int32)
    owner[fens.count() // 2:] = 1
    gpu_csc(fe, "elastic", fes, fens, u, rule, isotropic_C(), node_owner=owner, my_rank=1)
    owner[: fens.count() // 4] = 1
    g, ax = gpu_csc(fe, "elastic", fes, fens, u, rule, isotropic_C(), node_owner=owner, my_rank=1)
    n = u.nalldofs()
    _, (I, J, V) = oracle_csc(orc, "elastic", "H8", fes, fens, u, rule, isotropic_C())
    owned = np.zeros(n + 1, bool)
    owned[u.dofnums[owner == 1].reshape(-1)] = True
    blk = orc.sparse(I[owned[I]], J[owned[I]], V[owned[I]], n, n)
    np.testing.assert_array_equal(g[0], blk[0])
    np.testing.assert_array_equal(g[1], blk[1])


def test_result_survives_pattern_invalidation_and_rebuild(fe, orc, gpu_ctx):
    """An assembler's result borrows colptr / rowval from the pattern: invalidating the pattern, or rebuilding it for another
    partition, must not pull the arrays from under a result that has not been fetched yet (the pattern is reference-counted)."""
    fens, fes = fe.H8block(1.0, 2.0, 3.0, 10, 9, 8)
    u = make_field(fe, fens, 1)
    rule = fe.GaussRule(3, 2)
    ref, _ = oracle_csc(orc, "diffusion", "H8", fes, fens, u, rule, KAPPA3)
    femm = fe.FEMMBase(fe.IntegDomain(fes, rule))
    geom = fe.NodalField(fens.xyz)
    a = fe.SysmatAssemblerSparseGPU(0.0)
    a.setnomatrixresult(True)  # assemble, leave the CSC on the device
    fe.bilform_diffusion(femm, a, geom, u, fe.DataCache(KAPPA3), raw=True)
    a.setnomatrixresult(False)
    a.invalidate_patterns()
    # another assembler rebuilds the pattern for a partition and churns the block cache
    b = fe.SysmatAssemblerSparseGPU(0.0)
    owner = fe.slab_owner(fens.count(), 2)
    for p in (0, 1, 0):
        fe.bilform_diffusion(femm, b, geom, u, fe.DataCache(KAPPA3), raw=True, node_owner=owner, my_rank=p)
    got = a._fetch(True)
    assert_parity(ref, got)


def test_out_arrays_are_validated(fe, gpu_ctx):
    fens, fes = fe.H8block(1.0, 1.0, 1.0, 3, 3, 3)
    u = make_field(fe, fens, 1)
    rule = fe.GaussRule(3, 2)
    got, a = gpu_csc(fe, "diffusion", fes, fens, u, rule, KAPPA3)
    n, nnz = got[0].size - 1, got[2].size
    good = (np.empty(n + 1, np.int64), np.empty(nnz, np.int64), np.empty(nnz))
    for bad in ((np.empty(n, np.int64), good[1], good[2]), (good[0], np.empty(nnz, np.int32), good[2]),
                (good[0], good[1], np.empty(nnz - 1)), (good[0], good[1], np.empty(2 * nnz)[::2])):
        with pytest.raises(fe.FEGPUError, match="out\\["):
            gpu_csc(fe, "diffusion", fes, fens, u, rule, KAPPA3, assembler=a, out=bad)


def test_two_contexts_share_a_device(fe, orc, gpu_ctx):
    """Two contexts (own streams) on one device, forms queued asynchronously on both with DIFFERENT coefficients and rules: the
    quadrature tables and coefficients travel with each launch (kernel parameters), so neither can see the other's."""
    import torch
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    c1 = fe.GPUContext(0, stream=s1.cuda_stream)
    c2 = fe.GPUContext(0, stream=s2.cuda_stream)
    fens, fes = fe.H8block(1.0, 2.0, 3.0, 24, 24, 24)
    _distort(fens)
    rule = fe.GaussRule(3, 2)
    u3, u1 = make_field(fe, fens, 3), make_field(fe, fens, 1)
    C = isotropic_C()
    C2 = isotropic_C(E=7.0, nu=0.2)
    refs = [oracle_csc(orc, "elastic", "H8", fes, fens, u3, rule, C)[0], oracle_csc(orc, "diffusion", "H8", fes, fens, u1, rule, KAPPA3)[0],
            oracle_csc(orc, "elastic", "H8", fes, fens, u3, rule, C2)[0]]
    femm = fe.FEMMBase(fe.IntegDomain(fes, rule))
    geom = fe.NodalField(fens.xyz)
    a1, a2, a3 = fe.SysmatAssemblerSparseGPU(0.0, ctx=c1), fe.SysmatAssemblerSparseGPU(0.0, ctx=c2), fe.SysmatAssemblerSparseGPU(0.0, ctx=c2)
    for a in (a1, a2, a3):
        a.setnomatrixresult(True)
    c1.set_async(True)
    c2.set_async(True)
    for _ in range(3):  # interleave launches of the two contexts
        fe.bilform_lin_elastic(femm, a1, geom, u3, fe.DeforModelRed3D, fe.DataCache(C), raw=True)
        fe.bilform_diffusion(femm, a2, geom, u1, fe.DataCache(KAPPA3), raw=True)
        fe.bilform_lin_elastic(femm, a3, geom, u3, fe.DeforModelRed3D, fe.DataCache(C2), raw=True)
    c1.set_async(False)
    c2.set_async(False)
    for a, ref in zip((a1, a2, a3), refs):
        a.setnomatrixresult(False)
        assert_parity(ref, a._fetch(True))
    if torch.cuda.device_count() >= 2:  # the per-device kernel attribute (dynamic shared memory of k_h8_elastic) on a second device
        c3 = fe.GPUContext(1)
        a4 = fe.SysmatAssemblerSparseGPU(0.0, ctx=c3)
        got = fe.bilform_lin_elastic(femm, a4, geom, u3, fe.DeforModelRed3D, fe.DataCache(C), raw=True)
        assert_parity(refs[0], got)
        c3.release_meshes()
    for c in (c1, c2):
        c.release_meshes()
        c.release_cache()
