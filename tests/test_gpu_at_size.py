"""Entry-wise parity AT THE BENCHMARKED SIZES (BASELINE configs 2-5), single GPU and as ranks of an 8-way row-block partition.

The oracle cannot assemble 16.8 M elements in seconds, but a COLUMN of the matrix only receives triplets from the elements that
contain its node.  So for K random nodes the oracle assembles the patch of elements touching them (same element order, same
dof map, same arithmetic as a full run) and its columns of those nodes must be the GPU's: rowval bit-exact, nzval within
1e-12 of the matrix's max-abs entry (north_star tolerance).  For a rank of a partition the oracle's triplets are cut down to
the rows of the rank's own nodes first (the halo rule: a triplet belongs to the owner of its row node)."""
import ctypes as C

import numpy as np
import pytest

from conftest import KAPPA3, isotropic_C
from helpers import make_field

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _oracle_patch(orc, form, etname, conn, xyz, u, rule, coef, nodes0):
    """(colptr, rowval, nzval) of the elements that touch one of `nodes0` (0-based), assembled by the oracle in mesh order."""
    mask = np.isin(conn, np.asarray(nodes0) + 1).any(axis=1)
    sub = np.ascontiguousarray(conn[mask])
    n = u.nalldofs()
    if form == "diffusion":
        I, J, V = orc.bilform_diffusion_coo(etname, sub, xyz, u.dofnums, n, rule.param_coords, rule.weights, coef)
    elif form == "elastic":
        I, J, V = orc.bilform_lin_elastic_coo(etname, sub, xyz, u.dofnums, n, rule.param_coords, rule.weights, coef)
    else:
        I, J, V = orc.bilform_dot_coo(etname, sub, xyz, u.dofnums, n, rule.param_coords, rule.weights, coef)
    return I, J, V, n


def _check_columns(orc, trip, u, nodes0, got, scale, owner=None, rank=None, col_offset=0):
    I, J, V, n = trip
    if owner is not None:
        owned = np.zeros(n + 1, bool)
        owned[u.dofnums[owner == rank].reshape(-1)] = True
        keep = owned[I]
        I, J, V = I[keep], J[keep], V[keep]
    cp, rv, nz = orc.sparse(I, J, V, n, n)
    colptr, rowval, nzval = got
    checked = 0
    worst = 0.0
    for node in nodes0:
        for q in range(u.ndofs()):
            j = int(u.dofnums[node, q]) - 1
            ref_r, ref_v = rv[cp[j] - 1: cp[j + 1] - 1], nz[cp[j] - 1: cp[j + 1] - 1]
            jj = j - col_offset
            g_r, g_v = rowval[colptr[jj] - 1: colptr[jj + 1] - 1], nzval[colptr[jj] - 1: colptr[jj + 1] - 1]
            np.testing.assert_array_equal(g_r, ref_r)
            if ref_v.size:
                worst = max(worst, float(np.abs(g_v - ref_v).max()))
            checked += ref_v.size
    assert worst <= TOL * scale, "nzval error %.3e > %.0e * %.3e" % (worst, TOL, scale)
    return checked


def _pattern_path(fe, ctx, fes, u):
    from finetools_jl_b200 import _lib
    return _lib.lib().fegpu_pattern_path(ctx.device_mesh(fes).dofmap(u))


def _sample_nodes(rng, nn, k, extra=()):
    s = set(int(x) for x in rng.integers(0, nn, size=k))
    s.update(int(e) for e in extra)
    return np.array(sorted(s), dtype=np.int64)


def _release(ctx):
    ctx.release_meshes()
    ctx.release_cache()


@pytest.mark.parametrize("cfg", ["c2", "c4"])
def test_sampled_column_parity_h8_full_size(fe, orc, gpu_ctx, cfg):
    """BASELINE config 2 (128^3 H8 elasticity) and config 4 (256^3 H8 diffusion) at full size: ~200 random nodes plus mesh corners,
    an edge and a face node; then ranks 0, 3 and 7 of the 8-way z-slab partition the scaling bench uses."""
    rng = np.random.default_rng(20261017)
    n = 128 if cfg == "c2" else 256
    ndn, form, coef = (3, "elastic", isotropic_C()) if cfg == "c2" else (1, "diffusion", KAPPA3)
    fens, fes = fe.H8block(1.0, 1.0, 1.0, n, n, n)
    u = make_field(fe, fens, ndn)
    rule = fe.GaussRule(3, 2)
    femm = fe.FEMMBase(fe.IntegDomain(fes, rule))
    geom = fe.NodalField(fens.xyz)
    nn = fens.count()
    a = fe.SysmatAssemblerSparseGPU(0.0)

    def run(**kw):
        if form == "elastic":
            return fe.bilform_lin_elastic(femm, a, geom, u, fe.DeforModelRed3D, fe.DataCache(coef), raw=True, **kw)
        return fe.bilform_diffusion(femm, a, geom, u, fe.DataCache(coef), raw=True, **kw)

    colptr, rowval, nzval, m_, n_ = run()
    assert _pattern_path(fe, gpu_ctx, fes, u) == 2  # the thread-per-node kernels built it
    assert nzval.size == (9 * 385 ** 3 if cfg == "c2" else 769 ** 3)
    scale = float(np.abs(nzval).max())
    e = n + 1
    special = [0, n, nn - 1, e * e * (n // 2), e * (n // 2) + 3, e * e * 5 + e * 7 + 11]
    nodes = _sample_nodes(rng, nn, 200, special)
    trip = _oracle_patch(orc, form, "H8", fes.conn, fens.xyz, u, rule, coef, nodes)
    assert _check_columns(orc, trip, u, nodes, (colptr, rowval, nzval), scale) > 5000
    del colptr, rowval, nzval
    owner = fe.slab_owner(nn, 8)
    for rank in (0, 3, 7):
        cp, rv, nz, _, _ = run(node_owner=owner, my_rank=rank)
        assert _pattern_path(fe, gpu_ctx, fes, u) == 2
        lo, hi = np.nonzero(owner == rank)[0][[0, -1]]
        # nodes inside the slab, on both of its interfaces, and just outside (columns that hold only a few owned rows)
        near = [lo, lo + 1, hi, hi - 1, max(lo - e * e, 0), min(hi + e * e, nn - 1), max(lo - 1, 0), min(hi + 1, nn - 1)]
        rnodes = _sample_nodes(rng, hi - lo + 1, 60) + lo
        rnodes = np.array(sorted(set(rnodes.tolist()) | set(int(x) for x in near)), dtype=np.int64)
        trip = _oracle_patch(orc, form, "H8", fes.conn, fens.xyz, u, rule, coef, rnodes)
        assert _check_columns(orc, trip, u, rnodes, (cp, rv, nz), scale, owner=owner, rank=rank) > 1000
    a = None
    _release(gpu_ctx)


def test_sampled_column_parity_t10_full_size(fe, orc, gpu_ctx):
    """BASELINE config 3: consistent mass on the distorted 6 M-element T10 block (general symbolic path: mixed valences)."""
    rng = np.random.default_rng(3)
    n = 100
    f4, s4 = fe.T4block(1.0, 1.0, 1.0, n, n, n)
    h = 1.0 / n
    x0 = f4.xyz.copy()
    f4.xyz[:, 0] += 0.2 * h * np.sin(3 * np.pi * x0[:, 1]) * np.cos(2 * np.pi * x0[:, 2])
    f4.xyz[:, 1] += 0.2 * h * np.sin(3 * np.pi * x0[:, 2]) * np.cos(2 * np.pi * x0[:, 0])
    f4.xyz[:, 2] += 0.2 * h * np.sin(3 * np.pi * x0[:, 0]) * np.cos(2 * np.pi * x0[:, 1])
    fens, fes = fe.T4toT10(f4, s4)
    u = make_field(fe, fens, 1)
    rule = fe.TetRule(4)
    femm = fe.FEMMBase(fe.IntegDomain(fes, rule))
    geom = fe.NodalField(fens.xyz)
    a = fe.SysmatAssemblerSparseGPU(0.0)
    c = np.array([[1.0]])
    colptr, rowval, nzval, _, _ = fe.bilform_dot(femm, a, geom, u, fe.DataCache(c), raw=True)
    assert _pattern_path(fe, gpu_ctx, fes, u) == 1
    scale = float(np.abs(nzval).max())
    nn = fens.count()
    nodes = _sample_nodes(rng, nn, 150, [0, f4.count() - 1, f4.count(), nn - 1])  # vertex nodes and mid-edge nodes
    trip = _oracle_patch(orc, "dot", "T10", fes.conn, fens.xyz, u, rule, c, nodes)
    assert _check_columns(orc, trip, u, nodes, (colptr, rowval, nzval), scale) > 3000
    del colptr, rowval, nzval
    owner = fe.slab_owner(nn, 8)
    for rank in (2, 7):
        cp, rv, nz, _, _ = fe.bilform_dot(femm, a, geom, u, fe.DataCache(c), raw=True, node_owner=owner, my_rank=rank)
        lo, hi = np.nonzero(owner == rank)[0][[0, -1]]
        rnodes = np.array(sorted(set((_sample_nodes(rng, hi - lo + 1, 50) + lo).tolist()) | {int(lo), int(hi)}), dtype=np.int64)
        trip = _oracle_patch(orc, "dot", "T10", fes.conn, fens.xyz, u, rule, c, rnodes)
        assert _check_columns(orc, trip, u, rnodes, (cp, rv, nz), scale, owner=owner, rank=rank) > 500
    a = None
    _release(gpu_ctx)


def test_sampled_column_parity_h20_96cube(fe, orc, gpu_ctx):
    """BASELINE config 5 volume part at its stated size: H20 96^3 elasticity, GaussRule(3,3) (884 736 elements, 1.87 G nnz).  The
    30 GB result stays on the device; clusters of columns come back through the library's own block view
    (fegpu_makematrix_view, the FF-block machinery) and are compared with the oracle's patch columns.  Then rank 5 of 8."""
    rng = np.random.default_rng(5)
    n = 96
    fens, fes = fe.H20block(1.0, 1.0, 1.0, n, n, n)
    u = make_field(fe, fens, 3)
    rule = fe.GaussRule(3, 3)
    femm = fe.FEMMBase(fe.IntegDomain(fes, rule))
    geom = fe.NodalField(fens.xyz)
    a = fe.SysmatAssemblerSparseGPU(0.0)
    C6 = isotropic_C()
    a.setnomatrixresult(True)
    fe.bilform_lin_elastic(femm, a, geom, u, fe.DeforModelRed3D, fe.DataCache(C6), raw=True)
    a.setnomatrixresult(False)
    m_, n_, nnz = a.sizes()
    assert nnz == 9 * (234 * n ** 3 + 141 * n ** 2 + 24 * n + 1) and m_ == n_ == 3 * fens.count()
    nn = fens.count()
    nvert = (n + 1) ** 3
    starts = [0, nvert - 13, nvert + 5, nn - 12] + [int(x) for x in rng.integers(0, nn - 12, size=8)]
    total = 0
    scale = None
    for s0 in starts:  # 12 consecutive nodes = 36 consecutive columns (default numbering)
        nodes = np.arange(s0, s0 + 12)
        c0, c1 = int(u.dofnums[s0, 0]), int(u.dofnums[s0 + 11, 2])
        a.view(1, m_, c0, c1)
        cp, rv, nz, _, _ = a._fetch(True)
        a.view_reset()
        if scale is None:
            scale = 0.0
        scale = max(scale, float(np.abs(nz).max()))
        trip = _oracle_patch(orc, "elastic", "H20", fes.conn, fens.xyz, u, rule, C6, nodes)
        total += _check_columns(orc, trip, u, nodes, (cp, rv, nz), scale, col_offset=c0 - 1)
    assert total > 20000
    owner = fe.slab_owner(nn, 8)
    rank = 5
    a.setnomatrixresult(True)
    fe.bilform_lin_elastic(femm, a, geom, u, fe.DeforModelRed3D, fe.DataCache(C6), raw=True, node_owner=owner, my_rank=rank)
    a.setnomatrixresult(False)
    lo, hi = np.nonzero(owner == rank)[0][[0, -1]]
    for s0 in (int(lo) - 6, int(hi) - 6, int((lo + hi) // 2)):
        nodes = np.arange(s0, s0 + 12)
        c0, c1 = int(u.dofnums[s0, 0]), int(u.dofnums[s0 + 11, 2])
        a.view(1, m_, c0, c1)
        cp, rv, nz, _, _ = a._fetch(True)
        a.view_reset()
        trip = _oracle_patch(orc, "elastic", "H20", fes.conn, fens.xyz, u, rule, C6, nodes)
        _check_columns(orc, trip, u, nodes, (cp, rv, nz), scale, owner=owner, rank=rank, col_offset=c0 - 1)
    a = None
    _release(gpu_ctx)


def test_skins_96cube_parity_as_partition_ranks(fe, orc, gpu_ctx):
    """BASELINE config 5 surface part: Q4 / T3 boundary mass of the 96^3 blocks, full entry-wise parity (the skins are small enough
    for the oracle), single GPU and every rank of an 8-way partition of the VOLUME's nodes (most ranks own a ring of the skin)."""
    nb = 96
    c = np.array([[1.0]])
    for mesher, srule, et in ((fe.H8block, fe.GaussRule(2, 2), "Q4"), (fe.T4block, fe.TriRule(3), "T3")):
        vf, vol = mesher(1.0, 1.0, 1.0, nb, nb, nb)
        skin = fe.meshboundary(vol)
        psi = make_field(fe, vf, 1)
        n = psi.nalldofs()
        I, J, V = orc.bilform_dot_coo(et, skin.conn, vf.xyz, psi.dofnums, n, srule.param_coords, srule.weights, c, m=2)
        ref = orc.sparse(I, J, V, n, n)
        sa = fe.SysmatAssemblerSparseGPU(0.0)
        sfemm = fe.FEMMBase(fe.IntegDomain(skin, srule))
        geom = fe.NodalField(vf.xyz)
        got = fe.bilform_dot(sfemm, sa, geom, psi, fe.DataCache(c), m=2, raw=True)
        assert _pattern_path(fe, gpu_ctx, skin, psi) == 2
        np.testing.assert_array_equal(got[0], ref[0])
        np.testing.assert_array_equal(got[1], ref[1])
        scale = np.abs(ref[2]).max()
        assert np.abs(got[2] - ref[2]).max() <= TOL * scale
        owner = fe.slab_owner(vf.count(), 8)
        nnz_sum = 0
        for rank in range(8):
            owned = np.zeros(n + 1, bool)
            owned[psi.dofnums[owner == rank].reshape(-1)] = True
            keep = owned[I]
            blk = orc.sparse(I[keep], J[keep], V[keep], n, n)
            g = fe.bilform_dot(sfemm, sa, geom, psi, fe.DataCache(c), m=2, raw=True, node_owner=owner, my_rank=rank)
            np.testing.assert_array_equal(g[0], blk[0])
            np.testing.assert_array_equal(g[1], blk[1])
            if blk[2].size:
                assert np.abs(g[2] - blk[2]).max() <= TOL * scale
            nnz_sum += g[2].size
        assert nnz_sum == ref[2].size
    _release(gpu_ctx)
