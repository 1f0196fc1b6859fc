import os, sys, struct
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import orb_numpy as O
os.environ["B200GEO_ORB_DUMP"] = "/tmp/orb_dump.bin"
from ransac_with_homography_b200.ransac import orb_detect_and_compute
z = np.load("tests/golden/orb.npz")
rgb = z["small_rgb"]
kp, ds = orb_detect_and_compute(rgb, 500, as_arrays=True)
raw = open("/tmp/orb_dump.bin", "rb").read()
nbytes = struct.unpack_from("<q", raw, 0)[0]
lvsz = 8 * 40 + 8
L = raw[8:8 + lvsz]
print("sizeof L guess", lvsz, "file", len(raw), "expected", 8 + lvsz + 2 * nbytes)
pyr = np.frombuffer(raw, np.uint8, nbytes, 8 + lvsz)
sc = np.frombuffer(raw, np.uint8, nbytes, 8 + lvsz + nbytes)
gray = O.rgb2gray(rgb)
opyr = O.pyramid(gray)
for lv in range(8):
    w, h, pitch, pw, ph = struct.unpack_from("<5i", L, lv * 40)
    ofs, = struct.unpack_from("<q", L, lv * 40 + 24)
    co, cc = struct.unpack_from("<2I", L, lv * 40 + 32)
    ext = pyr[ofs:ofs + pitch * (h + 64)].reshape(h + 64, pitch)[:, :w + 64]
    s = sc[ofs:ofs + pitch * (h + 64)].reshape(h + 64, pitch)[32:32 + h, 32:32 + w]
    oimg, oext = opyr[lv]
    osc = O.fast_scores(oimg)
    print(lv, (w, h, pitch, pw, ph, ofs, co, cc), "oracle size", oimg.shape, "ext mismatch", int((ext != oext).sum()) if ext.shape == oext.shape else "shape", "score mismatch", int((s != osc).sum()) if s.shape == osc.shape else "shape", "nonzero gpu", int((s > 0).sum()), "oracle", int((osc > 0).sum()))
    if lv == 0 and s.shape == osc.shape:
        bad = np.argwhere(s != osc)[:5]
        for y, x in bad: print("   at", y, x, "gpu", s[y, x], "oracle", osc[y, x])
img = opyr[0][0].astype(int)
s0 = sc[0:240 * (121 + 64)].reshape(121 + 64, 240)[32:32 + 121, 32:32 + 163]
for (y, x) in [(3, 5), (6, 3), (6, 4), (10, 10)]:
    d = [img[y + dy, x + dx] - img[y, x] for dx, dy in O.RING]
    print((y, x), "centre", img[y, x], "d", d, "gpu", s0[y, x], "oracle", O.fast_scores(opyr[0][0])[y, x])
