"""Full-size parity inputs shared by the CPU and GPU tests: the reference's own fixture images (tests/golden/*.jpg) and
the digests of the reference's outputs on them (tests/golden/full_size.json, written by make_golden_full.py from the
live reference).  `oracle_case(name)` evaluates one case with the CPU oracle (test infrastructure)."""
import hashlib
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STEP = 9973
_cache = {}


def meta():
    if "meta" not in _cache:
        with open(os.path.join(GOLDEN, "full_size.json")) as f:
            _cache["meta"] = json.load(f)
    return _cache["meta"]


def image(name):
    """RGB uint8 image as the reference loads it (cv2.imread(...)[:, :, ::-1], ransac.py:291-294)."""
    if name not in _cache:
        import cv2
        _cache[name] = np.ascontiguousarray(cv2.imread(os.path.join(GOLDEN, name))[:, :, ::-1])
    return _cache[name].copy()


def record(a):
    a = np.ascontiguousarray(a)
    return {"shape": list(a.shape), "dtype": str(a.dtype), "sha256": hashlib.sha256(a.tobytes()).hexdigest(),
            "sample": a.reshape(-1).view(np.uint8)[::STEP].tolist()}


def same(a, rec):
    r = record(a)
    return all(r[k] == rec[k] for k in ("shape", "dtype", "sha256"))


def decoder_matches():
    """The JPEG decoder of this host gives the pixels the digests were made from."""
    return all(same(image(f), rec) for f, rec in meta()["images"].items())


SCAN_TAGS = {"a4_300dpi": (3508, 2480), "a4_app": (1188, 840)}


def scan_uv(tag):
    hh, ww = SCAN_TAGS[tag]
    u = np.array(meta()["scan_u"], np.float64)
    v = np.array([[0.0, ww, ww, 0], [0, 0, hh, hh], [1, 1, 1, 1]])
    return u, v


def crop_transform_image(imgn, v, mx, my):
    """The crop of transformImage (homography.py:226-228)."""
    sx = int(v[0, 0] - mx); sy = int(v[1, 0] - my)
    ex = int(v[0, 2] - mx); ey = int(v[1, 2] - my)
    return imgn[sy:ey + 1, sx:ex + 1, :]


def oracle_case(name):
    """(uint8 array, extras) of one full-size case by the CPU oracle, with the H the reference used."""
    import oracle
    from oracle import ref_numpy as port
    m = meta()
    c = m["cases"][name]
    Hs = np.array(m["H_stitch"])
    if name == "c1_tih_bilinear":
        o, mx, my = oracle.wrapPerspective(image("foto1A.jpg"), Hs, "bilinear")
        return oracle.to_u8(o), [mx, my]
    if name == "c1_bbox_nn":
        o, mx, my = oracle.wrapPerspective(image("foto1A.jpg"), Hs, "nn")
        return oracle.to_u8(o), [mx, my]
    if name == "c1_pano_paste":
        return port.stitch_panorama(image("foto1B.jpg"), image("foto1A.jpg"), Hs), None
    if name == "c1_pano_blend":
        return port.stitch_panorama(image("foto1B.jpg"), image("foto1A.jpg"), Hs, "Rate", 0.2), None
    if name == "c1_pano_paste_swapped":
        return port.stitch_panorama(image("foto1A.jpg"), image("foto1B.jpg"), np.array(m["H_stitch_inv"])), None
    if name.startswith("c2_a4"):
        tag, conv = name[3:].rsplit("_", 1)
        u, v = scan_uv(tag)
        o, _, _ = oracle.wrapPerspectiveScan(image("cv.jpg"), np.array(c["H"]), list(SCAN_TAGS[tag]), conv)
        return crop_transform_image(oracle.to_u8(o), v, 0, 0), None
    if name == "c2_nobox_bilinear":
        v = np.array([[0.0, 840, 840, 0], [0, 0, 1188, 1188], [1, 1, 1, 1]])
        o, mx, my = oracle.wrapPerspective(image("cv.jpg"), np.array(c["H"]), "bilinear")
        return crop_transform_image(oracle.to_u8(o), v, mx, my), [mx, my]
    if name == "notebook_bilinear":
        o, mx, my = oracle.wrapPerspective(image("notebook.jpg"), np.array(m["H_notebook"]), "bilinear")
        return oracle.to_u8(o), [mx, my]
    raise KeyError(name)
