"""Generates the golden CSC fixtures of this directory with the CPU oracle (run once in the build container:
`python tests/golden/make_golden.py`).  The reference is Julia-only and cannot run here, so these vectors pin the ORACLE
(itself pinned to the reference's known-answer tests in tests/test_oracle_pins.py), and through it the GPU path: the
`-m gpu` tests compare the CUDA result with the same files."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

KAPPA3 = np.array([[1.5, 0.2, 0.1], [0.2, 2.5, 0.3], [0.1, 0.3, 3.5]])


def _iso(E=1.0, nu=0.3):
    lam = E * nu / ((1 + nu) * (1 - 2 * nu))
    mu = E / (2 * (1 + nu))
    C = np.zeros((6, 6))
    C[:3, :3] = lam
    C[np.arange(3), np.arange(3)] += 2 * mu
    C[3:, 3:] = mu * np.eye(3)
    return C


# name -> (mesh generator name, args, ndn, rule, form, coefficient, fixed nodes)
CASES = {
    "h8_diffusion_6x5x4": ("H8block", (12.0, 1.1, 0.32, 6, 5, 4), 1, ("gauss", 3, 2), "diffusion", KAPPA3, None),
    "h8_elastic_4cube": ("H8block", (1.0, 1.0, 1.0, 4, 4, 4), 3, ("gauss", 3, 2), "elastic", _iso(), None),
    "h8_elastic_ebc": ("H8block", (1.0, 2.0, 3.0, 3, 3, 3), 3, ("gauss", 3, 2), "elastic", _iso(), [1, 2, 3, 30]),
    "t10_mass_3x2x4": ("T10block", (1.3, 3.1, 2.7, 3, 2, 4), 1, ("tet", 4), "dot", np.eye(1), None),
    "h20_elastic_2cube": ("H20block", (1.0, 1.0, 1.0, 2, 2, 2), 3, ("gauss", 3, 3), "elastic", _iso(), None),
    "q4_skin_mass": ("skinQ4", (1.3, 3.1, 2.7, 3, 2, 2), 1, ("gauss", 2, 2), "dot", np.eye(1), None),
    "t3_skin_mass": ("skinT3", (1.3, 3.1, 2.7, 2, 2, 2), 1, ("tri", 3), "dot", np.eye(1), None),
}


def build_case(fe, case):
    gen, args, ndn, rule, form, coef, fixed = case
    kw = {}
    if gen == "skinQ4":
        fens, vol = fe.H8block(*args)
        fes, et = fe.meshboundary(vol), "Q4"
        kw = {"m": 2}
    elif gen == "skinT3":
        fens, vol = fe.T4block(*args)
        fes, et = fe.meshboundary(vol), "T3"
        kw = {"m": 2}
    else:
        fens, fes = getattr(fe, gen)(*args)
        et = fes.name
    u = fe.NodalField(np.zeros((fens.count(), ndn)))
    if fixed is not None:
        fe.setebc(u, fixed, True, None, 0.0)
    fe.numberdofs(u)
    r = {"gauss": lambda: fe.GaussRule(rule[1], rule[2]), "tet": lambda: fe.TetRule(rule[1]), "tri": lambda: fe.TriRule(rule[1])}[rule[0]]()
    return fens, fes, u, r, coef, form, et, kw


def main():
    import finetools_jl_b200 as fe
    from oracle import oracle as orc
    from helpers import oracle_csc
    for name, case in CASES.items():
        fens, fes, u, rule, coef, form, et, kw = build_case(fe, case)
        (cp, rv, nz), _ = oracle_csc(orc, form, et, fes, fens, u, rule, coef, **kw)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), colptr=cp, rowval=rv, nzval=nz)
        print(name, "nnz", nz.size)


if __name__ == "__main__":
    main()
