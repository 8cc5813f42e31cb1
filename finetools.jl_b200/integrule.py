"""Quadrature rules on the hot path (reference: src/IntegRuleModule.jl:41-47 TriRule, :206-398 GaussRule,
:483-516 TetRule).  The constants are the reference's truncated decimals, NOT the exact abscissae: parity with
the reference requires these digits."""
import numpy as np

_G1 = {
    1: ([0.0], [2.0]),
    2: ([-0.577350269189626, 0.577350269189626], [1.0, 1.0]),
    3: ([-0.774596669241483, 0.0, 0.774596669241483], [0.5555555555555556, 0.8888888888888889, 0.5555555555555556]),
    4: ([-0.86113631159405, -0.33998104358486, 0.33998104358486, 0.86113631159405],
        [0.34785484513745, 0.65214515486255, 0.65214515486255, 0.34785484513745]),
}


class AbstractIntegRule:
    npts = 0
    param_coords = None  # (npts, dim)
    weights = None       # (npts, 1)


class GaussRule(AbstractIntegRule):
    """GaussRule(dim, order): tensor-product rule on [-1,1]^dim; first coordinate varies slowest
    (IntegRuleModule.jl:360-390)."""

    def __init__(self, dim=1, order=1):
        if not 1 <= dim <= 3:
            raise ValueError("Gauss rule of dimension %d not available" % dim)
        if order not in _G1:
            raise ValueError("Gauss rule of order %d not available" % order)
        x, w = (np.array(v) for v in _G1[order])
        self.dim, self.order = dim, order
        self.npts = order ** dim
        grids = np.meshgrid(*([np.arange(order)] * dim), indexing="ij")
        idx = np.stack([g.reshape(-1) for g in grids], axis=1)
        self.param_coords = x[idx]
        ww = w[idx[:, 0]]
        for d in range(1, dim):
            ww = ww * w[idx[:, d]]
        self.weights = ww.reshape(-1, 1)


class TetRule(AbstractIntegRule):
    def __init__(self, npts=1):
        if npts == 1:
            pc, w = [[0.25, 0.25, 0.25]], np.array([1.0]) / 6.0
        elif npts == 4:
            a, b = 0.13819660, 0.58541020
            pc, w = [[a, a, a], [b, a, a], [a, b, a], [a, a, b]], np.full(4, 0.041666666666666666667)
        elif npts == 5:
            a, b, c, d, e = 1.0 / 6.0, 0.25, 0.5, -0.8, 0.45
            pc, w = [[b, b, b], [c, a, a], [a, c, a], [a, a, c], [a, a, a]], np.array([d, e, e, e, e]) / 6
        else:
            raise ValueError("Unknown number of integration points")
        self.npts = npts
        self.param_coords = np.array(pc, dtype=np.float64)
        self.weights = np.asarray(w, dtype=np.float64).reshape(-1, 1)


class TriRule(AbstractIntegRule):
    def __init__(self, npts=1):
        if npts == 1:
            pc, w = [[1.0 / 3.0, 1.0 / 3.0]], np.array([1.0]) / 2.0
        elif npts == 3:
            pc, w = [[2.0 / 3, 1.0 / 6], [1.0 / 6, 2.0 / 3], [1.0 / 6, 1.0 / 6]], np.array([1.0 / 3, 1.0 / 3, 1.0 / 3]) / 2
        else:
            raise ValueError("TriRule(%d) is outside the hot-path scope (1 and 3 points are provided)" % npts)
        self.npts = npts
        self.param_coords = np.array(pc, dtype=np.float64)
        self.weights = np.asarray(w, dtype=np.float64).reshape(-1, 1)
