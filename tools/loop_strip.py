"""Per-call durations of back-to-back fast-warp launches (C4 shape).   python tools/loop_strip.py [frames] [gap_ms]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import frame_homographies
from ransac_with_homography_b200 import _cabi, device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
gap = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
h, w = 1080, 1920
ctx = _cabi.context(0)
ctx.warp_variant(int(os.environ.get("WARP_VARIANT", "0")))
src = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda")
dst = torch.empty_like(src)
jobs = device.upload_jobs(device.scan_jobs(frame_homographies(n, h, w, 0.05), (h, w), (h, w)), "cuda")
lay = (w * 3, n * h, h, w * 3)
for _ in range(3):
    device.warp_resident(src, dst, jobs, n, w, h, 3, "bilinear", layout=lay)
torch.cuda.synchronize()
R = 12
ev = [torch.cuda.Event(enable_timing=True) for _ in range(R + 1)]
ev[0].record()
for i in range(R):
    device.warp_resident(src, dst, jobs, n, w, h, 3, "bilinear", layout=lay)
    ev[i + 1].record()
    if gap:
        torch.cuda.synchronize(); time.sleep(gap / 1e3); ev[i + 1].record()
torch.cuda.synchronize()
print("frames", n, "gap", gap, "per-call ms:", " ".join(f"{ev[i].elapsed_time(ev[i+1]):.2f}" for i in range(R)))
if os.environ.get("NVML"):
    import threading, pynvml
    pynvml.nvmlInit(); hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
    samples = []; stop = threading.Event()
    def loop():
        while not stop.is_set():
            try:
                samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_MEM),
                                pynvml.nvmlDeviceGetPowerUsage(hnd) / 1000.0, pynvml.nvmlDeviceGetCurrentClocksEventReasons(hnd),
                                pynvml.nvmlDeviceGetTemperature(hnd, 0)))
            except Exception as e:
                samples.append((time.perf_counter(), str(e)))
            stop.wait(0.01)
    th = threading.Thread(target=loop, daemon=True); th.start()
    t0 = time.perf_counter()
    for i in range(60):
        device.warp_resident(src, dst, jobs, n, w, h, 3, "bilinear", layout=lay)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    stop.set(); th.join()
    print("60 calls: %.2f ms per call" % ((t1 - t0) * 1e3 / 60))
    for s in samples[::6]:
        print(" t=%.3f" % (s[0] - t0), *s[1:])
