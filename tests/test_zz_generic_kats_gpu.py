"""Known-answer tests of the reference's assembler protocol with RECTANGULAR element blocks, through the C ABI on the GPU
(generic startassembly! / assemble! / makematrix! path = 64-bit key radix sort + segmented sum).  The same vectors pin the CPU
oracle in tests/test_oracle_pins.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

M1 = np.array([[0.24406, 0.599773, 0.833404, 0.0420141], [0.786024, 0.00206713, 0.995379, 0.780298], [0.845816, 0.198459, 0.355149, 0.224996]])
M2 = np.array([[0.146618, 0.53471, 0.614342, 0.737833], [0.479719, 0.41354, 0.00760941, 0.836455], [0.254868, 0.476189, 0.460794, 0.00919633],
               [0.159064, 0.261821, 0.317078, 0.77646], [0.643538, 0.429817, 0.59788, 0.958909]])


def test_rectangular_blocks_dense_scatter_add(fe, gpu_ctx):
    """test/test_basics.jl:1465-1496: refa[rows, cols] += m for a 3x4 and a 5x4 block, 7x7 target."""
    blocks = [(M1, [1, 7, 5], [5, 2, 1, 4]), (M2, [2, 3, 1, 4, 5], [6, 7, 3, 4])]
    refa = np.zeros((7, 7))
    a = fe.SysmatAssemblerSparseGPU(0.0)
    fe.startassembly(a, 5, 5, 3, 7, 7)
    for m, dr, dc in blocks:
        refa[np.ix_(np.array(dr) - 1, np.array(dc) - 1)] += m
        fe.assemble(a, m, dr, dc)
    A = fe.makematrix(a)
    assert A.shape == (7, 7)
    assert np.abs(refa - A.toarray()).max() < 1e-15


def test_rectangular_blocks_golden_matrix(fe, gpu_ctx):
    """test/test_miscellaneous.jl:1944-1976: rows [1 7 5] / [2 3 1 7 5] against the reference's golden 7x7 matrix (tol 1e-5), and
    the entries pinned by test/test_basics.jl:1602-1606."""
    G = np.array([[0.833404, 0.599773, 0.460794, 0.0512104, 0.24406, 0.254868, 0.476189],
                  [0.0, 0.0, 0.614342, 0.737833, 0.0, 0.146618, 0.53471],
                  [0.0, 0.0, 0.00760941, 0.836455, 0.0, 0.479719, 0.41354],
                  [0.0] * 7,
                  [0.355149, 0.198459, 0.59788, 1.1839, 0.845816, 0.643538, 0.429817],
                  [0.0] * 7,
                  [0.995379, 0.00206713, 0.317078, 1.55676, 0.786024, 0.159064, 0.261821]])
    a = fe.SysmatAssemblerSparseGPU(0.0)
    fe.startassembly(a, 5, 5, 3, 7, 7)
    fe.assemble(a, M1, [1, 7, 5], [5, 2, 1, 4])
    fe.assemble(a, M2, [2, 3, 1, 7, 5], [6, 7, 3, 4])
    A = fe.makematrix(a).toarray()
    assert np.abs(G - A).max() < 1.0e-5
    for (i, j), v in {(1, 1): 0.833404, (5, 1): 0.355149, (7, 6): 0.159064, (3, 7): 0.41354, (7, 7): 0.261821}.items():
        assert abs(A[i - 1, j - 1] - v) <= 1e-12
    # duplicates are summed left to right, in assembly order
    assert A[0, 3] == 0.0420141 + 0.00919633 and A[4, 3] == 0.224996 + 0.958909 and A[6, 3] == 0.780298 + 0.77646
