"""Finite element sets: connectivity holder + basis functions, mirroring FinEtools' FESetModule for the
element types on the assembly hot path (T3, Q4, T4, T10, H8, H20, H27).

Reference: src/FESetModule.jl:57-93 (the `conn` holder), basis functions :665/:678 (T3), :713/:724 (Q4),
:1344/:1355 (T4), :1395/:1415 (T10), :955/:977 (H8), :1030/:1095 (H20), :1217/:1260 (H27); boundary
connectivity tables :994-1004 (H8), :1366-1369 (T4).

The basis functions here are written from the element definitions (signed-corner tables for the hexahedra,
1-D Lagrange factors for H27) rather than as expression lists; they agree with the reference's expressions to
rounding (checked against the oracle's literal transcription in tests/test_oracle_pins.py::test_host_basis_matches_oracle_transcription).
"""
import numpy as np

# C-ABI element type codes (include/fegpu.h)
ETYPE = {"T3": 1, "Q4": 2, "T4": 3, "T10": 4, "H8": 5, "H20": 6, "H27": 7}

# corner signs of the 8 hexahedron vertices in FinEtools node order
_HS = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1],
                [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=np.float64)
# H20 mid-edge nodes 9..20: (xi, eta, zeta) position with 0 on the edge's running coordinate
_H20E = np.array([[0, -1, -1], [1, 0, -1], [0, 1, -1], [-1, 0, -1],
                  [0, -1, 1], [1, 0, 1], [0, 1, 1], [-1, 0, 1],
                  [-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], dtype=np.float64)
# H27: position index (-1, 0, +1) of every node along xi, eta, zeta
_H27P = np.vstack([_HS, _H20E,
                   np.array([[0, 0, -1], [0, -1, 0], [1, 0, 0], [0, 1, 0], [-1, 0, 0], [0, 0, 1], [0, 0, 0]], dtype=np.float64)])


def _lag3(p, x):
    """Quadratic 1-D Lagrange function attached to position p in {-1,0,1} and its derivative."""
    if p < 0:
        return 0.5 * x * (x - 1.0), x - 0.5
    if p > 0:
        return 0.5 * x * (x + 1.0), x + 0.5
    return (1.0 - x) * (1.0 + x), -2.0 * x


class AbstractFESet:
    """Connectivity is an (nelem, nne) int64 array, 1-based node numbers (FESetModule.jl:61)."""
    name = None
    nne = 0
    mdim = 0

    def __init__(self, conn):
        conn = np.ascontiguousarray(conn, dtype=np.int64)
        if conn.ndim != 2 or conn.shape[1] != self.nne:
            raise ValueError("Connectivity of %s needs %d columns" % (self.name, self.nne))
        self.conn = conn
        self.label = np.zeros(conn.shape[0], dtype=np.int64)

    def count(self):
        return self.conn.shape[0]

    def __len__(self):
        return self.conn.shape[0]

    @property
    def etype(self):
        return ETYPE[self.name]

    def nodesperelem(self):
        return self.nne

    def manifdim(self):
        return self.mdim

    # bfun returns an (nne, 1) matrix, bfundpar an (nne, mdim) matrix, like the reference
    def bfun(self, pc):
        return self._bfun(np.asarray(pc, dtype=np.float64)).reshape(self.nne, 1)

    def bfundpar(self, pc):
        return self._bfundpar(np.asarray(pc, dtype=np.float64)).reshape(self.nne, self.mdim)

    def subset(self, idx):
        out = type(self)(self.conn[np.asarray(idx)])
        out.label = self.label[np.asarray(idx)]
        return out


class FESetT3(AbstractFESet):
    name, nne, mdim = "T3", 3, 2

    def _bfun(self, p):
        return np.array([1 - p[0] - p[1], p[0], p[1]])

    def _bfundpar(self, p):
        return np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])


class FESetQ4(AbstractFESet):
    name, nne, mdim = "Q4", 4, 2
    _S = _HS[:4, :2]

    def _bfun(self, p):
        s = self._S
        return 0.25 * (1.0 + s[:, 0] * p[0]) * (1.0 + s[:, 1] * p[1])

    def _bfundpar(self, p):
        s = self._S
        return np.column_stack([s[:, 0] * (1.0 + s[:, 1] * p[1]) * 0.25, s[:, 1] * (1.0 + s[:, 0] * p[0]) * 0.25])


class FESetT4(AbstractFESet):
    name, nne, mdim = "T4", 4, 3

    def _bfun(self, p):
        return np.array([1 - p[0] - p[1] - p[2], p[0], p[1], p[2]])

    def _bfundpar(self, p):
        return np.array([[-1.0, -1.0, -1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])

    def boundaryconn(self):
        c = self.conn
        return np.vstack([c[:, [0, 2, 1]], c[:, [0, 1, 3]], c[:, [1, 2, 3]], c[:, [0, 3, 2]]])

    boundaryfe = FESetT3


class FESetT10(AbstractFESet):
    name, nne, mdim = "T10", 10, 3
    # mid-edge node k (5..10) sits between vertices _E[k] (T4toT10 edge table, MeshTetrahedronModule.jl:160)
    _E = [(0, 1), (1, 2), (2, 0), (3, 0), (3, 1), (3, 2)]

    def _bary(self, p):
        L = np.array([1 - p[0] - p[1] - p[2], p[0], p[1], p[2]])
        dL = np.array([[-1.0, -1.0, -1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
        return L, dL

    def _bfun(self, p):
        L, _ = self._bary(p)
        return np.array([L[i] * (2 * L[i] - 1) for i in range(4)] + [4 * L[a] * L[b] for a, b in self._E])

    def _bfundpar(self, p):
        L, dL = self._bary(p)
        rows = [(4 * L[i] - 1) * dL[i] for i in range(4)] + [4 * (L[a] * dL[b] + L[b] * dL[a]) for a, b in self._E]
        return np.array(rows)


class FESetH8(AbstractFESet):
    name, nne, mdim = "H8", 8, 3

    def _bfun(self, p):
        s = _HS
        return (1.0 + s[:, 0] * p[0]) * (1.0 + s[:, 1] * p[1]) * (1.0 + s[:, 2] * p[2]) / 8.0

    def _bfundpar(self, p):
        s = _HS
        f = [1.0 + s[:, d] * p[d] for d in range(3)]
        return np.column_stack([s[:, 0] * f[1] * f[2], s[:, 1] * f[0] * f[2], s[:, 2] * f[0] * f[1]]) / 8.0

    def boundaryconn(self):
        c = self.conn
        return np.vstack([c[:, [0, 3, 2, 1]], c[:, [0, 1, 5, 4]], c[:, [1, 2, 6, 5]],
                          c[:, [2, 3, 7, 6]], c[:, [3, 0, 4, 7]], c[:, [5, 6, 7, 4]]])

    boundaryfe = FESetQ4


class FESetH20(AbstractFESet):
    name, nne, mdim = "H20", 20, 3

    def _eval(self, p):
        N = np.zeros(20)
        dN = np.zeros((20, 3))
        x = np.asarray(p, dtype=np.float64)
        for i in range(8):  # serendipity corner: (1+sx x)(1+sy y)(1+sz z)(sx x + sy y + sz z - 2)/8
            s = _HS[i]
            f = 1.0 + s * x
            g = s[0] * x[0] + s[1] * x[1] + s[2] * x[2] - 2.0
            N[i] = f[0] * f[1] * f[2] * g / 8.0
            for d in range(3):
                o = [k for k in range(3) if k != d]
                dN[i, d] = s[d] * f[o[0]] * f[o[1]] * (g + f[d]) / 8.0
        for k in range(12):  # mid-edge: (1 - t^2)(1+sa a)(1+sb b)/4 with t the running coordinate
            e = _H20E[k]
            t = int(np.where(e == 0)[0][0])
            o = [d for d in range(3) if d != t]
            fa, fb = 1.0 + e[o[0]] * x[o[0]], 1.0 + e[o[1]] * x[o[1]]
            q = (1.0 - x[t]) * (1.0 + x[t])
            N[8 + k] = q * fa * fb / 4.0
            dN[8 + k, t] = -2.0 * x[t] * fa * fb / 4.0
            dN[8 + k, o[0]] = q * e[o[0]] * fb / 4.0
            dN[8 + k, o[1]] = q * fa * e[o[1]] / 4.0
        return N, dN

    def _bfun(self, p):
        return self._eval(p)[0]

    def _bfundpar(self, p):
        return self._eval(p)[1]


class FESetH27(AbstractFESet):
    name, nne, mdim = "H27", 27, 3

    def _eval(self, p):
        N = np.zeros(27)
        dN = np.zeros((27, 3))
        for i in range(27):
            l = [_lag3(_H27P[i, d], float(p[d])) for d in range(3)]
            N[i] = l[0][0] * l[1][0] * l[2][0]
            dN[i, 0] = l[0][1] * l[1][0] * l[2][0]
            dN[i, 1] = l[0][0] * l[1][1] * l[2][0]
            dN[i, 2] = l[0][0] * l[1][0] * l[2][1]
        return N, dN

    def _bfun(self, p):
        return self._eval(p)[0]

    def _bfundpar(self, p):
        return self._eval(p)[1]


FESET_BY_NAME = {c.name: c for c in (FESetT3, FESetQ4, FESetT4, FESetT10, FESetH8, FESetH20, FESetH27)}
