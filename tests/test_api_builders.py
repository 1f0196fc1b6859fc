"""CPU: the DLT design-matrix builders of the drop-in module (pure layout code, no kernel involved) against the matrices
the LIVE reference builds (tests/golden/api.npz, make_golden_full.py): calc_corresp / calc_correspLinear /
calc_correspCollective / calc_correspLinearCollective (reference homography.py:4-69)."""
import numpy as np


def test_design_matrices_equal_reference(golden):
    from ransac_with_homography_b200 import homography as hg
    z = golden.file("api")
    A, B = golden.points()
    inl = z["inl"]
    m = hg.calc_corresp(z["u4"], z["v4"])
    assert m.dtype == np.float32 and np.array_equal(m, z["corresp"])
    a, b = hg.calc_correspLinear(z["u4"], z["v4"])
    assert a.dtype == np.float32 and b.dtype == np.float32
    assert np.array_equal(a, z["corresp_linear_A"]) and np.array_equal(b, z["corresp_linear_b"])
    m = hg.calc_correspCollective(A[inl], B[inl])
    assert m.shape == (2 * len(inl), 9) and np.array_equal(m, z["corresp_collective"])
    a, b = hg.calc_correspLinearCollective(A[inl], B[inl])
    assert np.array_equal(a, z["corresp_linear_collective_A"]) and np.array_equal(b, z["corresp_linear_collective_b"])
