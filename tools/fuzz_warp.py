"""Randomised parity check of the fast uint8 warp (K4) against the float64 kernel (itself bit-identical to the
oracle): random frame sizes, batch sizes, homography families (corner jitter, rotations, zooms, strong perspective,
large translations), samplers, output grids (Scan / bbox) and source layout hints (none, 2-D, 3-D).
nn must be identical, bilinear within +-1.      python tools/fuzz_warp.py [cases] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ransac_with_homography_b200 import _cabi, device


def rand_h(rs, h, w):
    kind = rs.randint(0, 6)
    if kind == 0:      # corner jitter
        src = np.array([[0, 0], [w, 0], [w, h], [0, h]], np.float64)
        dst = src + rs.uniform(-0.15, 0.15, (4, 2)) * [w, h]
        A = np.zeros((8, 8)); b = np.zeros(8)
        for j, ((x, y), (X, Y)) in enumerate(zip(src, dst)):
            A[2 * j] = [x, y, 1, 0, 0, 0, -x * X, -y * X]; A[2 * j + 1] = [0, 0, 0, x, y, 1, -x * Y, -y * Y]
            b[2 * j], b[2 * j + 1] = X, Y
        return np.append(np.linalg.solve(A, b), 1).reshape(3, 3)
    if kind == 1:      # rotation about the centre
        a = rs.uniform(0, 2 * np.pi)
        T = np.array([[1, 0, w / 2], [0, 1, h / 2], [0, 0, 1.0]])
        return T @ np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]) @ np.linalg.inv(T)
    if kind == 2:      # zoom
        z = rs.uniform(0.25, 4.0)
        return np.array([[z, 0, rs.uniform(-40, 40)], [0, z * rs.uniform(0.7, 1.4), rs.uniform(-40, 40)], [0, 0, 1.0]])
    if kind == 3:      # strong perspective
        return np.array([[1.0, 0.1, 5], [0.05, 1.0, -7], [rs.uniform(-5e-3, 5e-3), rs.uniform(-5e-3, 5e-3), 1.0]])
    if kind == 4:      # integer / half-integer translation (exactness cases), incl. out of the frame
        return np.array([[1.0, 0, rs.randint(-w, w) / 2.0], [0, 1.0, rs.randint(-h, h) / 2.0], [0, 0, 1.0]])
    return np.array([[rs.uniform(0.8, 1.2), rs.uniform(-0.3, 0.3), rs.uniform(-w / 2, w / 2)],
                     [rs.uniform(-0.3, 0.3), rs.uniform(0.8, 1.2), rs.uniform(-h / 2, h / 2)], [rs.uniform(-3e-4, 3e-4), rs.uniform(-3e-4, 3e-4), 1.0]])


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rs = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    ctx = _cabi.context(0)
    bad = 0
    worst = 0.0
    for c in range(cases):
        h, w = int(rs.randint(48, 700)), int(rs.randint(86, 900))
        B = int(rs.randint(1, 6))
        sampler = "nn" if rs.rand() < 0.4 else "bilinear"
        frames = rs.randint(0, 256, (B, h, w, 3)).astype(np.uint8)
        if rs.rand() < 0.3:
            frames[0, :, : w // 2] = 77
        Hs = np.stack([rand_h(rs, h, w) for _ in range(B)])
        form = "scan" if rs.rand() < 0.5 else "bbox"
        src_t = torch.from_numpy(frames).cuda()
        if form == "scan":
            oh, ow = (h, w) if rs.rand() < 0.5 else (int(rs.randint(40, 600)), int(rs.randint(40, 800)))
            jobs = device.scan_jobs(Hs, (h, w), (oh, ow))
        else:
            jobs, _ = device.bbox_jobs(Hs, (h, w))
            if int((jobs["dst_pitch"] * jobs["oh"]).sum()) > 300e6 or (jobs["ow"] <= 0).any() or (jobs["oh"] <= 0).any():
                continue
        total = int((jobs["dst_offset"] + jobs["dst_pitch"] * jobs["oh"]).max())
        jd = device.upload_jobs(jobs, "cuda")
        mo, mh = int(jobs["ow"].max()), int(jobs["oh"].max())
        layout = [None, (w * 3, B * h), (w * 3, B * h, h, w * 3)][rs.randint(0, 3)]
        if os.environ.get("FUZZ_ONLY") and c != int(os.environ["FUZZ_ONLY"]):
            continue
        if os.environ.get("FUZZ_FRAME"):
            f = int(os.environ["FUZZ_FRAME"])
            frames, Hs, B = frames[f:f + 1].copy(), Hs[f:f + 1], 1
            src_t = torch.from_numpy(frames).cuda()
            jobs = device.scan_jobs(Hs, (h, w), (jobs["oh"][0], jobs["ow"][0])) if form == "scan" else device.bbox_jobs(Hs, (h, w))[0]
            total = int((jobs["dst_offset"] + jobs["dst_pitch"] * jobs["oh"]).max())
            jd = device.upload_jobs(jobs, "cuda")
            mo, mh = int(jobs["ow"].max()), int(jobs["oh"].max())
            layout = None
            print("H =", repr(Hs[0]), flush=True)
        if os.environ.get("FUZZ_LAYOUT"):
            layout = [None, (w * 3, B * h), (w * 3, B * h, h, w * 3)][int(os.environ["FUZZ_LAYOUT"])]
        out = {}
        if os.environ.get("FUZZ_VERBOSE"):
            print(f"case {c}: {sampler} {form} {h}x{w} B={B} out {mh}x{mo} layout={layout}", flush=True)
        host_pipe = rs.rand() < 0.2          # the host-buffer pipeline (three streams, chunked) instead of the device-resident call
        chunk = int(rs.randint(1, 5))
        for m in (("exact", "fast") if os.environ.get("FUZZ_EXACT_FIRST") else ("fast", "exact")):
            ctx.warp_exact(m == "exact")
            ctx.warp_variant(int(os.environ.get("WARP_VARIANT", "0")))
            if host_pipe and m == "fast":
                out_h = np.full((total,), 7, np.uint8)
                device.warp_host(frames, out_h, jobs, 3, sampler, chunk_jobs=chunk)
                out[m] = out_h.astype(np.int16)
                continue
            dst_t = torch.full((total,), 7, dtype=torch.uint8, device="cuda")
            device.warp_resident(src_t, dst_t, jd, B, mo, mh, 3, sampler, layout=layout if m == "fast" else None)
            torch.cuda.synchronize()
            out[m] = dst_t.cpu().numpy().astype(np.int16)
        ctx.warp_exact(False)
        d = np.abs(out["fast"] - out["exact"])
        tol = 0 if sampler == "nn" else 1
        share = float((d > 0).mean())
        worst = max(worst, share if sampler == "bilinear" else 0.0)
        if d.max() > tol or (sampler == "bilinear" and share > 0.03):
            bad += 1
            print(f"FAIL case {c}: {sampler} {form} {h}x{w} B={B} layout={layout} host_pipe={host_pipe}/{chunk} max {int(d.max())} share {share:.4f}")
    print(f"{cases} cases, {bad} failures, worst bilinear off-by-one share {worst:.4f}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
