"""CPU check of the fast warp kernel's coordinate error bound (csrc/warp_fast.cuh, wf_setup / wf_rel16).

The kernel evaluates the source position of every pixel of a 64x40 tile in float32 from nine
tile-local coefficients and decides masks / nn rounding on those coordinates only when they are
further than `band` from the decision boundary.  This test mirrors the set-up arithmetic in numpy
(float64 set-up, coefficients rounded to float32, float32 evaluation with fused multiply-adds and a
reciprocal that is allowed to be 1 ulp off), evaluates whole tiles and requires the float32
coordinates to stay within HALF the band the kernel uses, for every tile the set-up does not send to
the float64 path.  No GPU needed: float32 fma / mul are IEEE and reproduced exactly here.
"""
import numpy as np
import pytest

TW, TH = 64, 40
F = np.float32


def fma32(a, b, c):
    # float32 x float32 is exact in float64; the sum is rounded to float64 and then to float32
    # (differs from a true fma only by a double rounding far below the tested margin)
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def setup(invH, x0, dx, y0, dy, c0, r0):
    """wf_setup, float64 part: returns the float32 coefficients, the anchor, band (px) or None if SLOW."""
    cm, rm = c0 + TW // 2, r0 + TH // 2
    Xm, Ym = cm * dx + x0, rm * dy + y0
    h = invH.ravel()
    zx, zy, zw = h[0] * Xm + h[1] * Ym + h[2], h[3] * Xm + h[4] * Ym + h[5], h[6] * Xm + h[7] * Ym + h[8]
    if not (1e-200 < abs(zw) < 1e200):
        return None
    s = 1.0 / zw
    xs, ys = zx * s, zy * s
    if not (abs(xs) < 4.0e6 and abs(ys) < 4.0e6):
        return None
    xc, yc = np.rint(xs), np.rint(ys)
    ax, bx = (h[0] - xc * h[6]) * s * dx, (h[1] - xc * h[7]) * s * dy
    ay, by = (h[3] - yc * h[6]) * s * dx, (h[4] - yc * h[7]) * s * dy
    gw, hw = h[6] * s * dx, h[7] * s * dy
    q = abs(gw) * (TW / 2) + abs(hw) * (TH / 2)
    if not q <= 0.25:
        return None
    Sx = abs(xs - xc) + abs(ax) * (TW / 2) + abs(bx) * (TH / 2)
    Sy = abs(ys - yc) + abs(ay) * (TW / 2) + abs(by) * (TH / 2)
    sc = 1.0 / 65536.0
    co = dict(nx0=F(xs - xc), ax=F(ax), bx=F(bx), ny0=F(ys - yc), ay=F(ay), by=F(by), w0=F(sc), gw=F(gw * sc), hw=F(hw * sc))
    return co, xc, yc, max(Sx, Sy)


def rel32(co, di, dj, rcp_ulp):
    """wf_compute_full's evaluation order: base at the thread's first pixel, then row and column steps."""
    di, dj = di.astype(F), dj.astype(F)
    nx = fma32(co["ax"], di, fma32(co["bx"], dj, np.full_like(di, co["nx0"])))
    ny = fma32(co["ay"], di, fma32(co["by"], dj, np.full_like(di, co["ny0"])))
    w = fma32(co["gw"], di, fma32(co["hw"], dj, np.full_like(di, co["w0"])))
    rw = (F(1) / w).astype(F)
    rw = np.nextafter(rw, F(np.inf) if rcp_ulp > 0 else F(-np.inf)) if rcp_ulp else rw
    return (nx * rw).astype(F).astype(np.float64) / 65536.0, (ny * rw).astype(F).astype(np.float64) / 65536.0


def homographies(rs, h, w):
    out = []
    for _ in range(40):   # corner jitter (the benchmark's family), up to +-10 %
        src = np.array([[0, 0], [w, 0], [w, h], [0, h]], np.float64)
        dst = src + rs.uniform(-0.1, 0.1, (4, 2)) * [w, h]
        A = np.zeros((8, 8)); b = np.zeros(8)
        for j, ((x, y), (X, Y)) in enumerate(zip(src, dst)):
            A[2 * j] = [x, y, 1, 0, 0, 0, -x * X, -y * X]; A[2 * j + 1] = [0, 0, 0, x, y, 1, -x * Y, -y * Y]
            b[2 * j], b[2 * j + 1] = X, Y
        out.append(np.append(np.linalg.solve(A, b), 1).reshape(3, 3))
    for _ in range(20):   # rotations about the centre, zooms, strong perspective
        a = rs.uniform(0, 2 * np.pi)
        T = np.array([[1, 0, w / 2], [0, 1, h / 2], [0, 0, 1.0]])
        out.append(T @ np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]) @ np.linalg.inv(T))
        z = rs.uniform(0.4, 3.0)
        out.append(np.array([[z, 0, rs.uniform(-30, 30)], [0, z * rs.uniform(0.8, 1.2), rs.uniform(-30, 30)], [0, 0, 1.0]]))
        out.append(np.array([[1.0, 0.1, 5], [0.05, 1.0, -7], [rs.uniform(-2e-3, 2e-3), rs.uniform(-2e-3, 2e-3), 1.0]]))
    return out


@pytest.mark.parametrize("grid", ["scan", "bbox"])
def test_float32_tile_coordinates_stay_inside_half_the_band(grid):
    rs = np.random.RandomState(3)
    h, w = 1080, 1920
    jj, ii = np.meshgrid(np.arange(TH), np.arange(TW), indexing="ij")
    worst, n_tiles, n_slow = 0.0, 0, 0
    for H in homographies(rs, h, w):
        invH = np.linalg.inv(H)
        if grid == "scan":
            x0, dx, y0, dy = 0.0, w / (w - 1), 0.0, h / (h - 1)
        else:
            x0, dx, y0, dy = float(rs.randint(-300, 300)), 1.0, float(rs.randint(-300, 300)), 1.0
        for _ in range(12):
            c0, r0 = int(rs.randint(0, w // TW)) * TW, int(rs.randint(0, h // TH)) * TH
            st = setup(invH, x0, dx, y0, dy, c0, r0)
            if st is None:
                n_slow += 1
                continue
            co, xc, yc, S = st
            # float64 truth, exactly the reference's formulation (homography.py:166-175)
            X, Y = (c0 + ii) * dx + x0, (r0 + jj) * dy + y0
            z = invH @ np.vstack([X.ravel(), Y.ravel(), np.ones(X.size)])
            tx, ty = z[0] / z[2] - xc, z[1] / z[2] - yc
            R = max(np.abs(tx).max(), np.abs(ty).max())
            if R >= 1024.0:
                n_slow += 1
                continue
            band = 2.4e-7 * (6.0 * S + 10.0 * R) + 4.0e-6          # wf_setup
            if band > 1.0 / 64:
                n_slow += 1
                continue
            for ulp in (-1, 0, 1):
                rx, ry = rel32(co, (ii - TW // 2).ravel(), (jj - TH // 2).ravel(), ulp)
                err = max(np.abs(rx - tx).max(), np.abs(ry - ty).max())
                worst = max(worst, err / band)
                assert err <= 0.5 * band, (grid, c0, r0, err, band)
            n_tiles += 1
    assert n_tiles > 500 and n_slow < n_tiles          # the test must actually exercise the fast path
