"""CPU: the oracle against the LIVE reference at full size (BASELINE configs 1 and 2 on the reference's own images),
pinned by the digests of tests/golden/full_size.json, and the small API vectors of tests/golden/api.npz."""
import numpy as np
import pytest

import fullsize
import oracle
from oracle import ref_numpy as port

CASES = sorted(fullsize.meta()["cases"].keys())


def test_fixture_images_decode_identically():
    pytest.importorskip("cv2")
    assert fullsize.decoder_matches()


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_at_full_size(name):
    pytest.importorskip("cv2")
    if not fullsize.decoder_matches():
        pytest.skip("this host's JPEG decoder differs from the one the digests were made with")
    rec = fullsize.meta()["cases"][name]
    out, off = fullsize.oracle_case(name)
    assert list(out.shape) == rec["shape"]
    assert out.reshape(-1)[::fullsize.STEP].tolist() == rec["sample"]
    assert fullsize.same(out, rec)
    if "off" in rec:
        assert off == rec["off"]


def test_port_design_matrices_equal_reference(golden):
    """calc_corresp / calc_correspLinear / ..Collective (homography.py:4-69) restated in oracle/ref_numpy.py."""
    z = golden.file("api")
    assert np.array_equal(port.dlt_rows_null(z["u4"], z["v4"]), z["corresp"])
    A, b = port.dlt_rows_inhom(z["u4"], z["v4"])
    assert np.array_equal(A, z["corresp_linear_A"]) and np.array_equal(b, z["corresp_linear_b"])
