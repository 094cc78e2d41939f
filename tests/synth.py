"""Deterministic synthetic inputs shared by tests and bench.py: band-limited noise
texture and the textured-plane motion model built on the reference's own flow equations
(numerical_simulation/simulation.py:7-12)."""
import numpy as np


def _cubic_up(grid, h, w):
    """separable Catmull-Rom upsampling of a coarse grid to (h, w) (no cv2 dependency)."""
    gh, gw = grid.shape

    def interp_axis(a, n_out, axis):
        n_in = a.shape[axis]
        pos = (np.arange(n_out) + 0.5) * (n_in - 3) / n_out + 1.0 - 0.5
        i0 = np.floor(pos).astype(int)
        t = pos - i0
        idx = [np.clip(i0 + k, 0, n_in - 1) for k in (-1, 0, 1, 2)]
        wts = [((-t + 2) * t - 1) * t / 2, (((3 * t - 5) * t) * t + 2) / 2, ((-3 * t + 4) * t + 1) * t / 2, ((t - 1) * t * t) / 2]
        out = 0
        for ii, ww in zip(idx, wts):
            taken = np.take(a, ii, axis=axis)
            shape = [1] * a.ndim
            shape[axis] = n_out
            out = out + taken * ww.reshape(shape)
        return out
    return interp_axis(interp_axis(grid, h, 0), w, 1)


def _blur(a, sigma):
    r = int(np.ceil(3 * sigma))
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
    k /= k.sum()
    p = np.pad(a, ((r, r), (0, 0)), mode="reflect")
    a = sum(p[i:i + a.shape[0]] * k[i] for i in range(2 * r + 1))
    p = np.pad(a, ((0, 0), (r, r)), mode="reflect")
    return sum(p[:, i:i + a.shape[1]] * k[i] for i in range(2 * r + 1))


def texture(h, w, seed):
    rng = np.random.default_rng(1000 + seed)
    coarse = rng.random((h // 8 + 3, w // 8 + 3))
    fine = rng.random((h, w))
    a = _cubic_up(coarse, h, w) + 0.3 * _blur(fine, 1.5)
    a = (a - a.min()) / (a.max() - a.min())
    return np.round(a * 255).astype(np.uint8)


def bilinear_sample(img, xs, ys):
    """img float (H,W); xs, ys float arrays; reflect-101 borders."""
    h, w = img.shape

    def refl(i, n):
        i = np.abs(i)
        i = np.where(i >= n, 2 * n - 2 - i, i)
        return np.clip(i, 0, n - 1)
    x0 = np.floor(xs).astype(int)
    y0 = np.floor(ys).astype(int)
    fx = xs - x0
    fy = ys - y0
    x0r, x1r, y0r, y1r = refl(x0, w), refl(x0 + 1, w), refl(y0, h), refl(y0 + 1, h)
    return ((1 - fy) * ((1 - fx) * img[y0r, x0r] + fx * img[y0r, x1r])
            + fy * ((1 - fx) * img[y1r, x0r] + fx * img[y1r, x1r]))


def flow_model(x, y, v, w, d, n):
    """u(X) of simulation.py:7-12 with t=0 at normalised image coordinates (vectorised)."""
    nx = n[0] * x + n[1] * y + n[2]
    cx = w[1] - w[2] * y
    cy = w[2] * x - w[0]
    cz = w[0] * y - w[1] * x
    s = nx / d
    return s * (v[0] - v[2] * x) + (cx - cz * x), s * (v[1] - v[2] * y) + (cy - cz * y)


def motion(pair_id, w, h, max_disp=20.0):
    """Random (v, omega, d, n) of SURVEY 8d, v rescaled so the largest displacement is <= max_disp px."""
    rng = np.random.default_rng(2000 + pair_id)
    v = rng.uniform(-1, 1, 3)
    om = rng.normal(0, 0.05, 3)
    d = rng.uniform(0.5, 5.0)
    n = np.array([rng.normal(0, 0.05), rng.normal(0, 0.05), 1.0])
    n /= np.linalg.norm(n)
    f = 0.8 * w
    dt = 1.0 / 30.0
    cx, cy = (w / 2 if w % 2 == 0 else (w + 1) / 2), (h / 2 if h % 2 == 0 else (h + 1) / 2)
    xs = (np.array([0, w - 1, 0, w - 1, w / 2]) - cx) / f
    ys = (np.array([0, 0, h - 1, h - 1, h / 2]) - cy) / f
    for _ in range(8):
        ux, uy = flow_model(xs, ys, v, om, d, n)
        m = np.max(np.hypot(ux, uy)) * f * dt
        if m <= max_disp:
            break
        v *= 0.7 * max_disp / m
        om *= 0.7 * max_disp / m
    return dict(v=v, w=om, d=d, n=n, f=f, dt=dt, cx=cx, cy=cy)


def warp_pair(img, mo):
    """Second frame = first frame moved by the displacement field f*dt*u(X) (backward warp, three
    fixed-point iterations so that a point at p in frame 1 lands at p + disp(p) in frame 2)."""
    h, w = img.shape
    f, dt, cx, cy = mo["f"], mo["dt"], mo["cx"], mo["cy"]
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    sx, sy = xx.copy(), yy.copy()
    for _ in range(3):
        ux, uy = flow_model((sx - cx) / f, (sy - cy) / f, mo["v"], mo["w"], mo["d"], mo["n"])
        sx = xx - ux * f * dt
        sy = yy - uy * f * dt
    out = bilinear_sample(img.astype(np.float64), sx, sy)
    return np.clip(np.round(out), 0, 255).astype(np.uint8)


def make_pair(h, w, stream_id=0, pair_id=0, max_disp=20.0):
    img = texture(h, w, stream_id)
    mo = motion(pair_id, w, h, max_disp)
    return img, warp_pair(img, mo), mo


def affine_pair(h, w, seed=0, shift=(2.3, -1.7), rot=0.01, scale=1.005):
    """Cheap pair for parity tests: frame 2 = affine warp of a texture."""
    img = texture(h, w, seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    c, s = np.cos(rot) * scale, np.sin(rot) * scale
    sx = c * (xx - w / 2) + s * (yy - h / 2) + w / 2 - shift[0]
    sy = -s * (xx - w / 2) + c * (yy - h / 2) + h / 2 - shift[1]
    out = bilinear_sample(img.astype(np.float64), sx, sy)
    return img, np.clip(np.round(out), 0, 255).astype(np.uint8)
