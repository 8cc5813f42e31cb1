"""Shared helpers for the parity tests: run the same problem through the oracle and through the C ABI."""
import numpy as np


def make_field(fe, fens, ndn, fixed_nodes=None, fixed_comp=None):
    u = fe.NodalField(np.zeros((fens.count(), ndn)))
    if fixed_nodes is not None:
        fe.setebc(u, fixed_nodes, True, fixed_comp, 0.0)
    fe.numberdofs(u)
    return u


def oracle_csc(orc, form, etname, fes, fens, u, rule, coef, **kw):
    n = u.nalldofs()
    if form == "diffusion":
        I, J, V = orc.bilform_diffusion_coo(etname, fes.conn, fens.xyz, u.dofnums, n, rule.param_coords, rule.weights, coef)
    elif form == "elastic":
        I, J, V = orc.bilform_lin_elastic_coo(etname, fes.conn, fens.xyz, u.dofnums, n, rule.param_coords, rule.weights, coef)
    elif form == "convection":  # coef = nodal velocity values (nnodes x sdim)
        I, J, V = orc.bilform_convection_coo(etname, fes.conn, fens.xyz, coef, u.dofnums, n, rule.param_coords, rule.weights, 1.0)
    elif form == "div_grad":
        I, J, V = orc.bilform_div_grad_coo(etname, fes.conn, fens.xyz, u.dofnums, n, rule.param_coords, rule.weights, coef)
    else:
        I, J, V = orc.bilform_dot_coo(etname, fes.conn, fens.xyz, u.dofnums, n, rule.param_coords, rule.weights, coef, **kw)
    return orc.sparse(I, J, V, n, n), (I, J, V)


def gpu_csc(fe, form, fes, fens, u, rule, coef, assembler=None, m=3, **kw):
    a = assembler if assembler is not None else fe.SysmatAssemblerSparseGPU(0.0)
    femm = fe.FEMMBase(fe.IntegDomain(fes, rule))
    geom = fe.NodalField(fens.xyz)
    if form == "diffusion":
        out = fe.bilform_diffusion(femm, a, geom, u, fe.DataCache(coef), raw=True, **kw)
    elif form == "elastic":
        out = fe.bilform_lin_elastic(femm, a, geom, u, fe.DeforModelRed3D, fe.DataCache(coef), raw=True, **kw)
    elif form == "convection":
        out = fe.bilform_convection(femm, a, geom, fe.NodalField(coef), u, fe.DataCache(1.0), raw=True, **kw)
    elif form == "div_grad":
        out = fe.bilform_div_grad(femm, a, geom, u, fe.DataCache(coef), raw=True, **kw)
    else:
        out = fe.bilform_dot(femm, a, geom, u, fe.DataCache(coef), m=m, raw=True, **kw)
    return out, a


def assert_parity(ref, got, tol=1e-12):
    """colptr / rowval bit-exact; nzval within tol * max|ref| (north_star tolerance: 1e-12)."""
    (cp, rv, nz) = ref
    colptr, rowval, nzval = got[0], got[1], got[2]
    assert colptr.dtype == np.int64 and rowval.dtype == np.int64 and nzval.dtype == np.float64
    np.testing.assert_array_equal(colptr, cp)
    np.testing.assert_array_equal(rowval, rv)
    scale = np.abs(nz).max() if nz.size else 1.0
    err = np.abs(nzval - nz).max() if nz.size else 0.0
    assert err <= tol * scale, "nzval error %.3e > %.1e * %.3e" % (err, tol, scale)
    return err / scale if scale else 0.0
