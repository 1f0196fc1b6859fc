"""Latency of the reference app's two single-frame configurations (BASELINE.json configs 1 and 2)
through the drop-in Python API (host numpy arrays in and out), beside the numpy port on the host.

    python tools/bench_app.py            # prints one JSON object

C1: RANSAC on the reference's fixture (185 matches; settings A: th=5,d=70,k=1000 of ransac.example0
    and B: th=4,d=95,k=1500 of app.py:364-367) + stitchPanorama warp of a 683x1024 RGB image by the
    fitted H (synthetic pixels of the reference's shapes: the JPEGs do not travel to the GPU box).
C2: Scanner: 4-point rectification of a 755x1226 image to A4 at 300 dpi (2480x3508), bilinear."""
import io, json, os, sys, time
from contextlib import redirect_stdout
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import ref_numpy as port
from ransac_with_homography_b200 import homography as hg
from ransac_with_homography_b200.ransac import HomoModel, RANSAC

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def best(fn, n=5):
    ts = []
    for _ in range(n):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return min(ts) * 1e3


z = np.load(os.path.join(ROOT, "tests", "golden", "fixture_points.npz"))
A, B = z["ptsA"], z["ptsB"]
rs = np.random.RandomState(0)
imgT = rs.randint(0, 256, (683, 1024, 3)).astype(np.uint8)
imgQ = rs.randint(0, 256, (683, 1024, 3)).astype(np.uint8)
out = {}
for tag, (th, d, k) in {"A": (5, 70, 1000), "B": (4, 95, 1500)}.items():
    idx = np.random.RandomState(1).randint(0, A.shape[0], (k, 4))
    res = {}

    def ours():
        with redirect_stdout(io.StringIO()):
            res["H"] = RANSAC(HomoModel(th=th, d=d, n=4), k=k).run([A.T, B.T], method="fwd", idx=idx)[0]

    def ref():
        res["Hr"] = port.ransac_run(A.T, B.T, th=th, d=d, n=4, k=k, method="fwd", idx=idx)[0]
    ours()
    out[f"C1_ransac_{tag}_ms"] = {"ours": best(ours), "numpy_port": best(ref, 3)}
H = res["H"]
hg.stitchPanorama(imgQ.copy(), imgT.copy(), H)
out["C1_stitchPanorama_ms"] = {"ours": best(lambda: hg.stitchPanorama(imgQ.copy(), imgT.copy(), H)),
                               "numpy_port_warp_only": best(lambda: port.warp_bbox(imgT.copy(), H, True)[0].astype(np.uint8), 2)}
with redirect_stdout(io.StringIO()):
    hg.stitchPanorama(imgQ.copy(), imgT.copy(), H, blending="Rate", blendrate=0.2)

    def ours_blend():
        hg.stitchPanorama(imgQ.copy(), imgT.copy(), H, blending="Rate", blendrate=0.2)

    def ref_blend():
        port.stitch_panorama(imgQ.copy(), imgT.copy(), H, "Rate", 0.2)
    out["C1_stitchPanorama_blend_ms"] = {"ours": best(ours_blend), "numpy_port": best(ref_blend, 1)}
# the matcher between ORB and RANSAC (ransac.py:258-259): 500 x 500 ORB-sized descriptors, cross-check
from ransac_with_homography_b200.ransac import match_hamming
dA, dB = rs.randint(0, 256, (500, 32)).astype(np.uint8), rs.randint(0, 256, (500, 32)).astype(np.uint8)
match_hamming(dA, dB)
out["C1_match_500x500_ms"] = {"ours": best(lambda: match_hamming(dA, dB))}
try:
    import cv2
    bf = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True)
    out["C1_match_500x500_ms"]["opencv_cpu"] = best(lambda: bf.match(dA, dB))
except ImportError:
    pass
img = rs.randint(0, 256, (755, 1226, 3)).astype(np.uint8)
u = np.array([[100, 560, 740, 80], [160, 125, 490, 555], [1, 1, 1, 1]], np.float64)
v = np.array([[0, 2480, 2480, 0], [0, 0, 3508, 3508], [1, 1, 1, 1]], np.float64)
hg.transformImage(img.copy(), u, v, box=[3508, 2480])
Hs = port.homography_normal_eq(u.T[:, :2], v.T[:, :2])
out["C2_scanner_2480x3508_ms"] = {"ours": best(lambda: hg.transformImage(img.copy(), u, v, box=[3508, 2480])),
                                  "numpy_port": best(lambda: port.warp_scan(img.copy(), Hs, [3508, 2480], True)[0].astype(np.uint8), 1)}
v2 = np.array([[0, 840, 840, 0], [0, 0, 1188, 1188], [1, 1, 1, 1]], np.float64)
hg.transformImage(img.copy(), u, v2, box=[1188, 840])
out["C2_scanner_840x1188_ms"] = {"ours": best(lambda: hg.transformImage(img.copy(), u, v2, box=[1188, 840]))}
# the whole Stitch button (ransac.example1 / app.py:364): stitching(foto1A, foto1B, "fwd", blending="Rate") on the reference's own pair,
# ours (no OpenCV on the path) against the unmodified reference (oracle/_ref: cv2 ORB + BFMatcher, numpy RANSAC, numpy warp + blend)
try:
    import cv2
    from ransac_with_homography_b200.ransac import stitching
    imA = np.ascontiguousarray(cv2.imread(os.path.join(ROOT, "tests", "golden", "foto1A.jpg"))[:, :, ::-1])
    imB = np.ascontiguousarray(cv2.imread(os.path.join(ROOT, "tests", "golden", "foto1B.jpg"))[:, :, ::-1])

    def ours_stitch():
        np.random.seed(0)
        with redirect_stdout(io.StringIO()):
            return stitching(imA.copy(), imB.copy(), "fwd", blending="Rate", blendrate=0.2)
    ours_stitch()
    rec = {"ours": best(ours_stitch)}
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if os.path.exists(os.path.join(ref_dir, "ransac.py")):
        sys.path.insert(0, ref_dir)
        cwd = os.getcwd()
        os.chdir(ref_dir)
        try:
            import ransac as ref_ransac              # the unmodified reference

            def ref_stitch():
                np.random.seed(0)
                with redirect_stdout(io.StringIO()):
                    return ref_ransac.stitching(imA.copy(), imB.copy(), "fwd", blending="Rate", blendrate=0.2)
            rec["reference"] = best(ref_stitch, 2)
        finally:
            os.chdir(cwd)
    out["C1_stitching_foto1A_foto1B_ms"] = rec
except ImportError:
    pass
out["host_cores"] = os.cpu_count()
print(json.dumps(out))
