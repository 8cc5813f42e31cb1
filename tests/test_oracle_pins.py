"""Pins the CPU oracle to the reference's own known-answer tests and identities:
  gradN golden vectors      test/test_basics.jl:1156-1169 (Q4), :1188-1207 (H8)
  7x7 assembler matrix      test/test_basics.jl:78-129
  dense-kernel identities   test/test_basics.jl:519-576
  form identities           test/test_forms.jl:5-24, 33-49, 58-76, 84-102, 136-153, 305-327, 335-365, 415-447
Runs without a GPU."""
import numpy as np
import pytest

from conftest import KAPPA3, isotropic_C


def _gradN_identity(dN):
    return dN  # redJ = identity: gradN == gradNparams


def test_gradN_golden_h8(orc):
    got = _gradN_identity(orc.bfundpar("H8", [0.57, 0.57, -0.57]))
    gold = np.array([[-0.0843875, -0.0843875, -0.023112500000000005], [0.0843875, -0.30811249999999996, -0.0843875],
                     [0.30811249999999996, 0.30811249999999996, -0.30811249999999996], [-0.30811249999999996, 0.0843875, -0.0843875],
                     [-0.023112500000000005, -0.023112500000000005, 0.023112500000000005], [0.023112500000000005, -0.0843875, 0.0843875],
                     [0.0843875, 0.0843875, 0.30811249999999996], [-0.0843875, 0.023112500000000005, 0.0843875]])
    assert np.linalg.norm(got - gold) <= 1.0e-6
    assert np.abs(got - gold).max() <= 1e-16  # in fact the digits are identical


def test_gradN_golden_q4(orc):
    got = orc.bfundpar("Q4", [0.57, 0.57])
    gold = np.array([[-0.10750000000000001, -0.10750000000000001], [0.10750000000000001, -0.39249999999999996],
                     [0.39249999999999996, 0.39249999999999996], [-0.39249999999999996, 0.10750000000000001]])
    assert np.linalg.norm(got - gold) <= 1.0e-6
    assert np.abs(got - gold).max() <= 1e-16


@pytest.mark.parametrize("et,nne,mdim", [("T3", 3, 2), ("Q4", 4, 2), ("T4", 4, 3), ("T10", 10, 3), ("H8", 8, 3), ("H20", 20, 3), ("H27", 27, 3)])
def test_host_basis_matches_oracle_transcription(orc, fe, et, nne, mdim):
    """The host package's basis functions (written from the element definitions) against the oracle's literal
    transcription of the reference expressions; plus partition of unity and a finite-difference derivative check."""
    fes = fe.FESET_BY_NAME[et](np.arange(1, nne + 1).reshape(1, nne))
    rng = np.random.default_rng(7)
    for _ in range(5):
        pc = rng.uniform(0.05, 0.3, size=mdim) if et.startswith("T") else rng.uniform(-0.9, 0.9, size=mdim)
        N_h, dN_h = fes.bfun(pc).reshape(-1), fes.bfundpar(pc)
        N_o, dN_o = orc.bfun(et, pc), orc.bfundpar(et, pc)
        assert np.abs(N_h - N_o).max() <= 4e-16
        assert np.abs(dN_h - dN_o).max() <= 2e-15
        assert abs(N_o.sum() - 1.0) <= 1e-14
        assert np.abs(dN_o.sum(axis=0)).max() <= 1e-13
        h = 1e-6
        for d in range(mdim):
            e = np.zeros(mdim)
            e[d] = h
            fd = (orc.bfun(et, pc + e) - orc.bfun(et, pc - e)) / (2 * h)
            assert np.abs(fd - dN_o[:, d]).max() <= 1e-8


def test_rules_match_and_integrate(orc, fe):
    for dim, order in [(1, 2), (2, 2), (3, 2), (3, 3), (2, 3), (3, 4)]:
        r = fe.GaussRule(dim, order)
        pc, w = orc.gauss_rule(dim, order)
        np.testing.assert_array_equal(r.param_coords, pc)
        np.testing.assert_array_equal(r.weights.reshape(-1), w)
        assert abs(w.sum() - 2.0 ** dim) < 1e-13
    # tensor order: first coordinate slowest (IntegRuleModule.jl:374-390)
    r = fe.GaussRule(3, 2)
    assert r.param_coords[1, 2] > 0 and r.param_coords[1, 0] < 0 and r.param_coords[4, 0] > 0
    for n in (1, 4, 5):
        r = fe.TetRule(n)
        pc, w = orc.tet_rule(n)
        np.testing.assert_array_equal(r.param_coords, pc)
        np.testing.assert_array_equal(r.weights.reshape(-1), w)
        assert abs(w.sum() - 1.0 / 6) < 1e-15
    assert fe.TetRule(4).param_coords[1, 0] == 0.58541020  # the reference's 8-digit constant
    for n in (1, 3):
        r = fe.TriRule(n)
        pc, w = orc.tri_rule(n)
        np.testing.assert_array_equal(r.param_coords, pc)
        np.testing.assert_array_equal(r.weights.reshape(-1), w)
        assert abs(w.sum() - 0.5) < 1e-15


def test_dense_kernels_vs_matrix_expressions(orc):
    """test/test_basics.jl:519-576 (rel. tol 1e-9 there)."""
    L = orc.lib()
    rng = np.random.default_rng(3)
    N = 8
    g2 = rng.random((N, 2))
    Kv = np.zeros(N * N)
    L.orc_add_mggt_ut_only(Kv, np.ascontiguousarray(g2.T).reshape(-1), 3.0, N, 2)
    L.orc_complete_lt(Kv, N)
    K = Kv.reshape(N, N).T
    ref = 3.0 * (g2 @ g2.T)
    assert np.linalg.norm(K - ref) / np.linalg.norm(ref) <= 1e-9

    g3 = rng.random((N, 3))
    kap = rng.random((3, 3))
    kap = kap + kap.T
    Kv = np.zeros(N * N)
    scratch = np.zeros(3 * N)
    L.orc_add_gkgt_ut_only(Kv, np.ascontiguousarray(g3.T).reshape(-1), 0.33, np.ascontiguousarray(kap.T).reshape(-1), scratch, N, 3)
    L.orc_complete_lt(Kv, N)
    ref = 0.33 * (g3 @ kap @ g3.T)
    assert np.linalg.norm(Kv.reshape(N, N).T - ref) / np.linalg.norm(ref) <= 1e-9

    N = 12
    B = rng.random((3, N))
    D = rng.random((3, 3))
    D = D + D.T
    Kv = np.zeros(N * N)
    DB = np.zeros(3 * N)
    L.orc_add_btdb_ut_only(Kv, np.ascontiguousarray(B.T).reshape(-1), 0.33, np.ascontiguousarray(D.T).reshape(-1), DB, 3, N)
    L.orc_complete_lt(Kv, N)
    ref = 0.33 * (B.T @ D @ B)
    assert np.linalg.norm(Kv.reshape(N, N).T - ref) / np.linalg.norm(ref) <= 1e-9


def test_assembler_testA(orc):
    """test/test_basics.jl:78-129: two dense blocks into a 7x7 matrix."""
    m1 = np.array([[0.24406, 0.599773, 0.833404, 0.0420141], [0.786024, 0.00206713, 0.995379, 0.780298],
                   [0.845816, 0.198459, 0.355149, 0.224996]])
    m1 = m1.T @ m1
    i1 = np.array([5, 2, 1, 4], dtype=np.int64)
    m2 = np.array([[0.146618, 0.53471, 0.614342, 0.737833], [0.479719, 0.41354, 0.00760941, 0.836455],
                   [0.254868, 0.476189, 0.460794, 0.00919633], [0.159064, 0.261821, 0.317078, 0.77646],
                   [0.643538, 0.429817, 0.59788, 0.958909]])
    m2 = m2.T @ m2
    i2 = np.array([2, 3, 1, 5], dtype=np.int64)
    testA = np.array([[2.85928, 1.21875, 0.891063, 0.891614, 2.56958, 0.0, 0.0], [1.21875, 1.15515, 0.716396, 0.0714644, 1.56825, 0.0, 0.0],
                      [0.891063, 0.716396, 0.936979, 0.0, 1.36026, 0.0, 0.0], [0.891614, 0.0714644, 0.0, 0.661253, 0.813892, 0.0, 0.0],
                      [2.56958, 1.56825, 1.36026, 0.813892, 4.15934, 0.0, 0.0], [0.0] * 7, [0.0] * 7])
    L = orc.lib()
    I, J, V = np.zeros(32, np.int64), np.zeros(32, np.int64), np.zeros(32)
    p = np.zeros(1, np.int64)
    for m, idx in ((m1, i1), (m2, i2)):
        rc = L.orc_assemble(I, J, V, p, np.ascontiguousarray(m.T).reshape(-1), idx, 4, idx, 4, 7, 7)
        assert rc == 0
    assert p[0] == 32
    # emission order: column-major walk (AssemblyModule.jl:266-279)
    assert (I[:4] == i1).all() and (J[:4] == i1[0]).all()
    cp, rv, nz = orc.sparse(I, J, V, 7, 7)
    A = orc.to_scipy(cp, rv, nz, 7, 7).toarray()
    assert np.abs(testA - A).max() < 1.0e-5
    # range checks with the reference's precedence
    assert L.orc_assemble(I, J, V, np.zeros(1, np.int64), np.zeros(4), np.array([1, 9], np.int64), 2, np.array([1, 2], np.int64), 2, 7, 7) == 4
    assert L.orc_assemble(I, J, V, np.zeros(1, np.int64), np.zeros(4), np.array([1, 2], np.int64), 2, np.array([0, 2], np.int64), 2, 7, 7) == 1


def test_sparse_semantics(orc):
    """sparse(): duplicates summed left to right, explicit zeros kept, rows ascending, rectangular sizes."""
    I = np.array([3, 1, 3, 2, 3, 1], np.int64)
    J = np.array([2, 1, 2, 4, 2, 1], np.int64)
    V = np.array([1e16, 1.0, 1.0, 0.0, -1e16, -1.0])
    cp, rv, nz = orc.sparse(I, J, V, 3, 5)
    np.testing.assert_array_equal(cp, [1, 2, 3, 3, 4, 4])
    np.testing.assert_array_equal(rv, [1, 3, 2])
    assert nz[0] == 0.0                      # 1 + (-1): a stored zero stays
    assert nz[1] == (1e16 + 1.0) - 1e16      # left-to-right: (1e16 + 1) - 1e16 == 0.0 in binary64, not 1.0
    assert nz[2] == 0.0                      # explicit zero kept
    import scipy.sparse as sp
    rng = np.random.default_rng(0)
    n = 5000
    I = rng.integers(1, 60, n); J = rng.integers(1, 45, n); V = rng.standard_normal(n)
    cp, rv, nz = orc.sparse(I, J, V, 59, 44)
    ref = sp.coo_matrix((V, (I - 1, J - 1)), shape=(59, 44)).tocsc()
    ref.sort_indices()
    np.testing.assert_array_equal(cp - 1, ref.indptr)
    np.testing.assert_array_equal(rv - 1, ref.indices)
    assert np.abs(nz - ref.data).max() < 1e-12
    with pytest.raises(ValueError):
        orc.sparse(np.array([4], np.int64), np.array([1], np.int64), np.array([1.0]), 3, 3)


def _quad_form(orc, form, et, fens, fes, field_vals, rule, coef, **kw):
    import finetools_jl_b200 as fe
    u = fe.NodalField(field_vals)
    fe.numberdofs(u)
    n = u.nalldofs()
    if form == "diffusion":
        I, J, V = orc.bilform_diffusion_coo(et, fes.conn, fens.xyz, u.dofnums, n, rule.param_coords, rule.weights, coef)
    elif form == "elastic":
        I, J, V = orc.bilform_lin_elastic_coo(et, fes.conn, fens.xyz, u.dofnums, n, rule.param_coords, rule.weights, coef)
    else:
        I, J, V = orc.bilform_dot_coo(et, fes.conn, fens.xyz, u.dofnums, n, rule.param_coords, rule.weights, coef, **kw)
    cp, rv, nz = orc.sparse(I, J, V, n, n)
    K = orc.to_scipy(cp, rv, nz, n, n)
    v = fe.gathersysvec(u)
    return float(v @ (K @ v)), K


def test_form_identities_of_the_reference(orc, fe):
    W, L, t = 1.1, 12.0, 0.32
    fens, fes = fe.H8block(L, W, t, 2, 4, 3)
    x = fens.xyz
    ones = np.ones((fens.count(), 1))
    g2 = fe.GaussRule(3, 2)
    # bilform_dot: v'Gv = volume (test_forms.jl:5-24), 3 dofs: 3 * volume (test_basics.jl:1001-1013)
    q, _ = _quad_form(orc, "dot", "H8", fens, fes, ones, g2, np.eye(1))
    assert abs(q - W * L * t) / (W * L * t) <= 1e-5
    q, _ = _quad_form(orc, "dot", "H8", fens, fes, np.ones((fens.count(), 3)), g2, np.eye(3))
    assert abs(q - 3 * W * L * t) / (W * L * t) <= 1e-5
    # bilform_diffusion: constant field -> 0 (:33-49); linear field -> (b^2+c^2+d^2) V (:58-76)
    q, _ = _quad_form(orc, "diffusion", "H8", fens, fes, 0.3 * ones, g2, np.eye(3))
    assert abs(q) / (W * L * t) <= 1e-5
    a, b, c, d = -0.1, 0.3, 0.4, -0.5
    lin = (a + b * x[:, 0] + c * x[:, 1] + d * x[:, 2]).reshape(-1, 1)
    q, _ = _quad_form(orc, "diffusion", "H8", fens, fes, lin, g2, np.eye(3))
    assert abs(q - (b * b + c * c + d * d) * W * L * t) / (W * L * t) <= 1e-5
    # scalar-kappa iso path (:136-153)
    q, _ = _quad_form(orc, "diffusion", "H8", fens, fes, lin, g2, 1.0)
    assert abs(q - (b * b + c * c + d * d) * W * L * t) / (W * L * t) <= 1e-5
    # H20 with GaussRule(3,3) (:84-102)
    f20, s20 = fe.H20block(L, W, t, 2, 4, 3)
    x20 = f20.xyz
    lin20 = (a + b * x20[:, 0] + c * x20[:, 1] + d * x20[:, 2]).reshape(-1, 1)
    q, _ = _quad_form(orc, "diffusion", "H20", f20, s20, lin20, fe.GaussRule(3, 3), np.eye(3))
    assert abs(q - (b * b + c * c + d * d) * W * L * t) / (W * L * t) <= 1e-5
    # T10 and H27 see the same linear field exactly too
    f10, s10 = fe.T10block(L, W, t, 2, 4, 3)
    x10 = f10.xyz
    lin10 = (a + b * x10[:, 0] + c * x10[:, 1] + d * x10[:, 2]).reshape(-1, 1)
    q, _ = _quad_form(orc, "diffusion", "T10", f10, s10, lin10, fe.TetRule(4), np.eye(3))
    assert abs(q - (b * b + c * c + d * d) * W * L * t) / (W * L * t) <= 1e-5
    f27, s27 = fe.H27block(L, W, t, 2, 2, 2)
    x27 = f27.xyz
    lin27 = (a + b * x27[:, 0] + c * x27[:, 1] + d * x27[:, 2]).reshape(-1, 1)
    q, _ = _quad_form(orc, "diffusion", "H27", f27, s27, lin27, fe.GaussRule(3, 3), np.eye(3))
    assert abs(q - (b * b + c * c + d * d) * W * L * t) / (W * L * t) <= 1e-5


def test_elastic_identities_of_the_reference(orc, fe):
    W, L, t = 11.1, 12.0, 7.32
    fens, fes = fe.H8block(L, W, t, 2, 4, 3)
    x = fens.xyz
    g2 = fe.GaussRule(3, 2)
    mu = 0.00133
    C = np.diag([2 * mu, 2 * mu, 2 * mu, mu, mu, mu])
    # rigid translation -> 0 (test_forms.jl:305-327)
    rigid = np.column_stack([np.full(fens.count(), 3.1), np.full(fens.count(), -2.7), np.full(fens.count(), -0.77)])
    q, K = _quad_form(orc, "elastic", "H8", fens, fes, rigid, g2, C)
    assert abs(q) / (W * L * t) <= 1e-5
    # exact symmetry G - G' == 0 (:441-442): the triangle is mirrored, summation order is the same on both sides
    assert abs(K - K.T).max() == 0.0
    # linear displacement -> 2 mu V |sym grad u|^2 (:335-365)
    a, b, c, d = -0.33, 2 / 3, -1.67, 2 / 7
    mu = 0.13377
    C = np.diag([2 * mu, 2 * mu, 2 * mu, mu, mu, mu])
    uu = np.column_stack([a + b * x[:, 0] + c * x[:, 1] + d * x[:, 2], b + c * x[:, 0] + d * x[:, 1] + a * x[:, 2],
                          c + d * x[:, 0] + a * x[:, 1] + b * x[:, 2]])
    gradu = np.array([[b, c, d], [c, d, a], [d, a, b]])
    gs = (gradu + gradu.T) / 2
    int_true = 2 * mu * (W * L * t) * (gs ** 2).sum()
    q, _ = _quad_form(orc, "elastic", "H8", fens, fes, uu, g2, C)
    assert abs(q - int_true) / int_true <= 1e-5


def test_surface_mass_is_the_area(orc, fe):
    """Jacobiansurface on the boundary skin (sdim 3, manifold 2): 1' M 1 = surface area."""
    Lx, Ly, Lz = 1.3, 3.1, 2.7
    area = 2 * (Lx * Ly + Ly * Lz + Lx * Lz)
    fens, vol = fe.H8block(Lx, Ly, Lz, 3, 2, 4)
    q4 = fe.meshboundary(vol)
    q, _ = _quad_form(orc, "dot", "Q4", fens, q4, np.ones((fens.count(), 1)), fe.GaussRule(2, 2), np.eye(1), m=2)
    assert abs(q - area) / area < 1e-12
    fens, vol = fe.T4block(Lx, Ly, Lz, 3, 2, 4)
    t3 = fe.meshboundary(vol)
    q, _ = _quad_form(orc, "dot", "T3", fens, t3, np.ones((fens.count(), 1)), fe.TriRule(3), np.eye(1), m=2)
    assert abs(q - area) / area < 1e-12


def test_survey_fingerprints(orc, fe):
    """SURVEY.md appendix B (independent NumPy restatement written at survey time)."""
    fens, fes = fe.H8block(12.0, 1.1, 0.32, 20, 20, 20)
    u = fe.NodalField(np.zeros((fens.count(), 1)))
    fe.numberdofs(u)
    r = fe.GaussRule(3, 2)
    I, J, V = orc.bilform_diffusion_coo("H8", fes.conn, fens.xyz, u.dofnums, u.nalldofs(), r.param_coords, r.weights, KAPPA3)
    cp, rv, nz = orc.sparse(I, J, V, u.nalldofs(), u.nalldofs())
    assert nz.size == 226981
    np.testing.assert_array_equal(cp[:5], [1, 9, 21, 33, 45])
    np.testing.assert_array_equal(rv[:8], [1, 2, 22, 23, 442, 443, 463, 464])
    K = orc.to_scipy(cp, rv, nz, u.nalldofs(), u.nalldofs())
    assert abs(K[0, 0] - 0.88226262626262) < 1e-13 and abs(K[1, 0] - 0.44003964646464) < 1e-13
    assert abs(K[499, 499] - 6.80650101010100) < 1e-12
    assert abs(K.diagonal().sum() - 54452.00808080) < 1e-7
    fens, fes = fe.H8block(1, 1, 1, 4, 4, 4)
    u = fe.NodalField(np.zeros((fens.count(), 3)))
    fe.numberdofs(u)
    I, J, V = orc.bilform_lin_elastic_coo("H8", fes.conn, fens.xyz, u.dofnums, u.nalldofs(), r.param_coords, r.weights, isotropic_C())
    cp, rv, nz = orc.sparse(I, J, V, u.nalldofs(), u.nalldofs())
    assert nz.size == 19773
    np.testing.assert_array_equal(cp[:4], [1, 25, 49, 73])
    K = orc.to_scipy(cp, rv, nz, u.nalldofs(), u.nalldofs())
    assert abs(K[0, 0] - 0.058760683760683) < 1e-14 and abs(K.diagonal().sum() - 90.2564102564102) < 1e-11
    fens, fes = fe.T10block(1.3, 3.1, 2.7, 3, 2, 4)
    assert fens.count() == 315 and fes.count() == 144
    u = fe.NodalField(np.zeros((fens.count(), 1)))
    fe.numberdofs(u)
    r4 = fe.TetRule(4)
    I, J, V = orc.bilform_dot_coo("T10", fes.conn, fens.xyz, u.dofnums, u.nalldofs(), r4.param_coords, r4.weights, np.eye(1))
    cp, rv, nz = orc.sparse(I, J, V, u.nalldofs(), u.nalldofs())
    assert nz.size == 6789
    assert abs(nz.sum() - 10.881) < 1e-12


def _testA_blocks():
    m1 = np.array([[0.24406, 0.599773, 0.833404, 0.0420141], [0.786024, 0.00206713, 0.995379, 0.780298],
                   [0.845816, 0.198459, 0.355149, 0.224996]])
    m2 = np.array([[0.146618, 0.53471, 0.614342, 0.737833], [0.479719, 0.41354, 0.00760941, 0.836455],
                   [0.254868, 0.476189, 0.460794, 0.00919633], [0.159064, 0.261821, 0.317078, 0.77646],
                   [0.643538, 0.429817, 0.59788, 0.958909]])
    return (m1.T @ m1, np.array([5, 2, 1, 4], dtype=np.int64)), (m2.T @ m2, np.array([2, 3, 1, 5], dtype=np.int64))


def test_symm_assembler_testA(orc):
    """test/test_basics.jl:119-129: SysmatAssemblerSparseSymm on the same two blocks equals the general assembler's
    matrix and is symmetric; the empty rows/columns 6, 7 stay empty."""
    I, J, V = [], [], []
    M = np.zeros((7, 7))
    for m, ii in _testA_blocks():
        M[np.ix_(ii - 1, ii - 1)] += m
        for j in range(4):            # lower triangle, column outer (AssemblyModule.jl:517-530)
            for i in range(j, 4):
                I.append(ii[i]); J.append(ii[j]); V.append(m[i, j])
    cp, rv, nz = orc.sparse_symm(np.array(I), np.array(J), np.array(V), 7)
    A = orc.to_scipy(cp, rv, nz, 7, 7).toarray()
    assert np.abs(A - M).max() < 1e-5 and np.abs(A - A.T).max() == 0.0
    assert cp[-1] == cp[5] and nz.size == 23        # 25 pattern entries minus the (3,4)/(4,3) pair that never meets
    # exact cancellation is dropped by the sparse `+` (zero-preserving map), a stored zero too
    cp, rv, nz = orc.sparse_symm(np.array([2, 2, 3], np.int64), np.array([1, 1, 3], np.int64), np.array([1.5, -1.5, 0.0]), 3)
    assert nz.size == 0 and (cp == 1).all()


def test_matrix_block_semantics(orc):
    """matrix_blocked_ff/fd/df/dd = Julia range indexing (MatrixUtilityModule.jl:675-793): stored zeros survive, rows rebased."""
    I = np.array([1, 3, 4, 2, 4, 1, 3], np.int64)
    J = np.array([1, 1, 1, 2, 3, 4, 4], np.int64)
    V = np.array([1.0, 0.0, 3.0, 4.0, 5.0, 6.0, 7.0])
    csc = orc.sparse(I, J, V, 4, 4)
    full = orc.to_scipy(*csc, 4, 4).toarray()
    nf = 2
    for (r0, r1, c0, c1) in ((1, nf, 1, nf), (1, nf, nf + 1, 4), (nf + 1, 4, 1, nf), (nf + 1, 4, nf + 1, 4)):
        cp, rv, nz = orc.matrix_block(csc, r0, r1, c0, c1)
        B = orc.to_scipy(cp, rv, nz, r1 - r0 + 1, c1 - c0 + 1).toarray()
        assert np.array_equal(B, full[r0 - 1:r1, c0 - 1:c1])
    cp, rv, nz = orc.matrix_block(csc, 3, 4, 1, 2)
    assert list(cp) == [1, 3, 3] and list(rv) == [1, 2] and list(nz) == [0.0, 3.0]   # the stored zero of (3,1) is kept


def test_sibling_form_identities_of_the_reference(orc, fe):
    """SURVEY.md 8(f) rank 3 restatements pinned to the reference's own tests: bilform_div_grad v'Gv = 0 for a constant field
    and = 2 mu V |sym grad u|^2 for a linear one (test/test_forms.jl:234-300), bilform_convection Psi' K Q = (u . grad q) V
    (:196-231, here on H8), distribloads / linform_dot sum(F) = f V (:158-193)."""
    import scipy.sparse as sp
    W, L, t = 11.1, 12.0, 7.32
    fens, fes = fe.H8block(L, W, t, 2, 4, 3)
    x = fens.xyz
    rule = fe.GaussRule(3, 2)
    u = fe.NodalField(np.zeros((fens.count(), 3)))
    fe.numberdofs(u)
    n = u.nalldofs()
    a, b, c, d = (-0.33, 2 / 3, -1.67, 2 / 7)
    mu = 0.13377
    I, J, V = orc.bilform_div_grad_coo("H8", fes.conn, x, u.dofnums, n, rule.param_coords, rule.weights, mu)
    cp, rv, nz = orc.sparse(I, J, V, n, n)
    G = sp.csc_matrix((nz, rv - 1, cp - 1), shape=(n, n))
    v = np.zeros(n)
    v[u.dofnums - 1] = np.stack([a + b * x[:, 0] + c * x[:, 1] + d * x[:, 2], b + c * x[:, 0] + d * x[:, 1] + a * x[:, 2],
                                 c + d * x[:, 0] + a * x[:, 1] + b * x[:, 2]], 1)
    gradu = np.array([[b, c, d], [c, d, a], [d, a, b]])
    true = 2 * mu * W * L * t * (((gradu + gradu.T) / 2) ** 2).sum()
    assert abs(v @ (G @ v) - true) / true <= 1.0e-5
    vc = np.zeros(n)
    vc[u.dofnums - 1] = np.array([3.1, -2.7, -0.77])
    assert abs(vc @ (G @ vc)) / (W * L * t) <= 1.0e-5
    q = fe.NodalField(np.zeros((fens.count(), 1)))
    fe.numberdofs(q)
    nq = q.nalldofs()
    uv = np.tile([3.1, -2.7, 0.4], (fens.count(), 1))
    I, J, V = orc.bilform_convection_coo("H8", fes.conn, x, uv, q.dofnums, nq, rule.param_coords, rule.weights, 1.0)
    cp, rv, nz = orc.sparse(I, J, V, nq, nq)
    K = sp.csc_matrix((nz, rv - 1, cp - 1), shape=(nq, nq))
    Q = np.zeros(nq)
    Q[q.dofnums[:, 0] - 1] = -0.1 + 0.3 * x[:, 0] + 0.4 * x[:, 1] + 0.5 * x[:, 2]
    assert abs(np.ones(nq) @ (K @ Q) - (0.3 * 3.1 + 0.4 * -2.7 + 0.5 * 0.4) * W * L * t) / (W * L * t) <= 1.0e-5
    F = orc.linform_dot("H8", fes.conn, x, q.dofnums, nq, rule.param_coords, rule.weights, [11.0])
    assert abs(F.sum() - L * W * t * 11.0) / 667 <= 1.0e-5


def _assemble_blocks(orc, blocks, nrows, ncols, cap=64):
    L = orc.lib()
    I, J, V = np.zeros(cap, np.int64), np.zeros(cap, np.int64), np.zeros(cap)
    p = np.zeros(1, np.int64)
    for m, dr, dc in blocks:
        m = np.asarray(m, dtype=np.float64)
        rc = L.orc_assemble(I, J, V, p, np.ascontiguousarray(m.T).reshape(-1), np.asarray(dr, np.int64), len(dr),
                            np.asarray(dc, np.int64), len(dc), nrows, ncols)
        assert rc == 0
    n = int(p[0])
    return I[:n], J[:n], V[:n]


RECT_M1 = [[0.24406, 0.599773, 0.833404, 0.0420141], [0.786024, 0.00206713, 0.995379, 0.780298], [0.845816, 0.198459, 0.355149, 0.224996]]
RECT_M2 = [[0.146618, 0.53471, 0.614342, 0.737833], [0.479719, 0.41354, 0.00760941, 0.836455], [0.254868, 0.476189, 0.460794, 0.00919633],
           [0.159064, 0.261821, 0.317078, 0.77646], [0.643538, 0.429817, 0.59788, 0.958909]]


def test_assembler_rectangular_blocks_kat(orc):
    """test/test_basics.jl:1465-1496: a 3x4 and a 5x4 block with different row and column dof lists into a 7x7 matrix; the result
    equals the dense scatter-add refa[rows, cols] += m (tolerance of the reference's test: 1e-5; here exact up to one rounding)."""
    blocks = [(RECT_M1, [1, 7, 5], [5, 2, 1, 4]), (RECT_M2, [2, 3, 1, 4, 5], [6, 7, 3, 4])]
    refa = np.zeros((7, 7))
    for m, dr, dc in blocks:
        refa[np.ix_(np.array(dr) - 1, np.array(dc) - 1)] += np.array(m)
    I, J, V = _assemble_blocks(orc, blocks, 7, 7)
    assert I.size == 12 + 20
    # emission order of assemble! (AssemblyModule.jl:261-280): column by column, rows inner
    assert list(I[:3]) == [1, 7, 5] and list(J[:3]) == [5, 5, 5] and V[1] == 0.786024
    cp, rv, nz = orc.sparse(I, J, V, 7, 7)
    A = orc.to_scipy(cp, rv, nz, 7, 7).toarray()
    assert np.abs(refa - A).max() < 1e-15
    assert nz.size == np.count_nonzero(refa) == 12 + 20 - 2   # (1,4) and (5,4) are hit by both blocks


def test_assembler_nomatrixresult_flow_kat(orc):
    """test/test_basics.jl:1580-1608: the second block goes to rows [2 3 1 7 5]; the reference pins five entries of the matrix built
    after `setnomatrixresult(a, false)`: A[1,1] = 0.833404, A[5,1] = 0.355149, A[7,6] = 0.159064, A[3,7] = 0.41354, A[7,7] = 0.261821."""
    blocks = [(RECT_M1, [1, 7, 5], [5, 2, 1, 4]), (RECT_M2, [2, 3, 1, 7, 5], [6, 7, 3, 4])]
    I, J, V = _assemble_blocks(orc, blocks, 7, 7)
    cp, rv, nz = orc.sparse(I, J, V, 7, 7)
    A = orc.to_scipy(cp, rv, nz, 7, 7).toarray()
    for (i, j), v in {(1, 1): 0.833404, (5, 1): 0.355149, (7, 6): 0.159064, (3, 7): 0.41354, (7, 7): 0.261821}.items():
        assert abs(A[i - 1, j - 1] - v) <= 1e-12, (i, j)


def test_assembler_rectangular_blocks_golden_matrix(orc):
    """test/test_miscellaneous.jl:1944-1976: the same two rectangular blocks (rows [1 7 5] / [2 3 1 7 5]) against the full 7x7
    golden matrix of the reference's test (5 significant digits, tolerance 1e-5 as there; note that the reference's check is
    one-sided, maximum(G - A) < 1e-5 -- here both signs)."""
    G = np.array([[0.833404, 0.599773, 0.460794, 0.0512104, 0.24406, 0.254868, 0.476189],
                  [0.0, 0.0, 0.614342, 0.737833, 0.0, 0.146618, 0.53471],
                  [0.0, 0.0, 0.00760941, 0.836455, 0.0, 0.479719, 0.41354],
                  [0.0] * 7,
                  [0.355149, 0.198459, 0.59788, 1.1839, 0.845816, 0.643538, 0.429817],
                  [0.0] * 7,
                  [0.995379, 0.00206713, 0.317078, 1.55676, 0.786024, 0.159064, 0.261821]])
    blocks = [(RECT_M1, [1, 7, 5], [5, 2, 1, 4]), (RECT_M2, [2, 3, 1, 7, 5], [6, 7, 3, 4])]
    I, J, V = _assemble_blocks(orc, blocks, 7, 7)
    cp, rv, nz = orc.sparse(I, J, V, 7, 7)
    A = orc.to_scipy(cp, rv, nz, 7, 7).toarray()
    assert np.abs(G - A).max() < 1.0e-5
    # the three entries hit by both blocks are sums in assembly order: (1,4), (5,4), (7,4)
    assert A[0, 3] == 0.0420141 + 0.00919633 and A[4, 3] == 0.224996 + 0.958909 and A[6, 3] == 0.780298 + 0.77646


def test_cubic_symmetry_identity_of_the_elasticity_kernels(orc, fe):
    """The GPU elasticity kernels integrate a material matrix of the cubic-symmetry form (isotropic: MatDeforElastIso) through
    B_a(:,i)' D B_b(:,j) = lam g_a,i g_b,j + mu g_a,j g_b,i (i != j),  D00 g_a,i g_b,i + mu sum_{k != i} g_a,k g_b,k (i == j)
    (csrc/fegpu_h8.cu: outer_acc / iso_block).  Checked here on the CPU against the oracle's literal B' D B loop
    (FEMMBaseModule.jl:1774-1813, DeforModelRedModule.jl:463-468) for one distorted H8 element, isotropic and cubic D."""
    rng = np.random.default_rng(3)
    fens, fes = fe.H8block(1.0, 1.3, 0.7, 1, 1, 1)
    xyz = fens.xyz + 0.08 * rng.standard_normal(fens.xyz.shape)
    rule = fe.GaussRule(3, 2)
    u = fe.NodalField(np.zeros((8, 3)))
    fe.numberdofs(u)
    for cubic in (False, True):
        lam, mu, d00 = 0.9, 0.55, 0.9 + 2 * 0.55
        if cubic:
            d00, mu = 2.7, 0.31
        D = np.zeros((6, 6))
        D[:3, :3] = lam
        D[np.arange(3), np.arange(3)] = d00
        D[3:, 3:] = mu * np.eye(3)
        I, J, V = orc.bilform_lin_elastic_coo("H8", fes.conn, xyz, u.dofnums, 24, rule.param_coords, rule.weights, D)
        K_ref = np.zeros((24, 24))
        np.add.at(K_ref, (I - 1, J - 1), V)
        K = np.zeros((24, 24))
        conn = fes.conn[0] - 1
        for pc, w in zip(rule.param_coords, rule.weights):
            dNpar = np.asarray(orc.bfundpar("H8", pc)).reshape(8, 3)
            Jm = xyz[conn].T @ dNpar                      # J[s, d] = sum_a x[a, s] dN[a, d]
            g = dNpar @ np.linalg.inv(Jm)                 # gradN! (FESetModule.jl:507-544)
            Jw = np.linalg.det(Jm) * w
            for a in range(8):
                for b in range(8):
                    P = Jw * np.outer(g[a], g[b])         # P[i, j] = Jw g_a,i g_b,j
                    blk = lam * P + mu * P.T
                    tr = np.trace(P)
                    for i in range(3):
                        blk[i, i] = d00 * P[i, i] + mu * (tr - P[i, i])
                    K[np.ix_(u.dofnums[conn[a]] - 1, u.dofnums[conn[b]] - 1)] += blk
        assert np.abs(K - K_ref).max() <= 1e-13 * np.abs(K_ref).max()
