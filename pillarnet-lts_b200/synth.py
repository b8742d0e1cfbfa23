"""Seeded synthetic LiDAR frames shaped like the two benchmark datasets.

No dataset or network is available, so the benchmark and the parity tests run on generated point
clouds: a spinning multi-beam sensor over a ground plane with smoothed random "facades" per azimuth,
multi-sweep accumulation with ego motion (nuScenes: 10 sweeps, 32 beams; Waymo: 1 sweep, 64 beams),
close-point removal (det3d/datasets/pipelines/loading.py:37-46) and a per-sweep time channel
(loading.py:135-140).  Output rows are [x, y, z, intensity, dt|elongation] float32 — the `points`
entries of det3d's example dict.
"""
import numpy as np

PRESETS = {
    # beams, elevation range (deg), azimuth steps, sweeps, sensor height, max range
    "nuscenes": dict(beams=32, elev=(-30.0, 10.0), az=1090, sweeps=10, height=1.84, max_range=70.0,
                     ego_step=0.5, dropout=0.18),
    "waymo": dict(beams=64, elev=(-17.6, 2.4), az=2650, sweeps=1, height=2.1, max_range=76.0,
                  ego_step=0.0),
}


def _smooth_periodic(rng, n, k):
    x = rng.normal(size=n)
    f = np.fft.rfft(x)
    f[k:] = 0
    y = np.fft.irfft(f, n)
    return (y - y.min()) / max(y.max() - y.min(), 1e-9)


def make_frame(kind="nuscenes", seed=0, n_points=None):
    """Returns (N,5) float32.  `n_points` resamples the frame to an exact size (config-5 sweeps)."""
    p = PRESETS[kind]
    rng = np.random.default_rng(seed)
    elev = np.deg2rad(np.linspace(p["elev"][0], p["elev"][1], p["beams"]))
    az = np.linspace(-np.pi, np.pi, p["az"], endpoint=False)
    facade_r = 6.0 + 44.0 * _smooth_periodic(rng, p["az"], 24) ** 1.5
    facade_h = 2.0 + 8.0 * _smooth_periodic(rng, p["az"], 12)
    out = []
    for s in range(p["sweeps"]):
        shift = np.array([p["ego_step"] * s, 0.02 * s])
        A, E = np.meshgrid(az + rng.uniform(0, 2 * np.pi / p["az"]), elev, indexing="ij")
        fr = np.roll(facade_r, rng.integers(0, 3))[:, None] * np.ones_like(E)
        fh = facade_h[:, None] * np.ones_like(E)
        with np.errstate(divide="ignore", invalid="ignore"):
            r_ground = np.where(E < -1e-3, p["height"] / np.tan(-E), np.inf)
        z_at_facade = p["height"] + fr * np.tan(E)
        hit_facade = (z_at_facade > 0) & (z_at_facade < fh)
        r = np.where(hit_facade & (fr < r_ground), fr, r_ground)
        ok = np.isfinite(r) & (r < p["max_range"])
        ok &= rng.random(r.shape) > p.get("dropout", 0.08)  # dropout
        r = r * (1.0 + rng.normal(0, 0.002, r.shape))
        x = r * np.cos(A) - shift[0]
        y = r * np.sin(A) - shift[1]
        z = r * np.tan(E)  # relative to the sensor; ground is at -height
        ok &= ~((np.abs(x) < 1.0) & (np.abs(y) < 1.0))  # remove_close
        n = int(ok.sum())
        inten = rng.random(n)
        if kind == "nuscenes":
            last = np.full(n, 0.05 * s)
        else:
            inten = np.tanh(inten * 2.0)
            last = rng.random(n) * 0.5  # elongation
        out.append(np.stack([x[ok], y[ok], z[ok], inten, last], axis=1))
    pts = np.concatenate(out).astype(np.float32)
    if n_points is not None:
        idx = rng.integers(0, len(pts), n_points) if n_points > len(pts) else rng.permutation(len(pts))[:n_points]
        pts = pts[np.sort(idx)]
        if n_points > len(out) and n_points > 0:
            pts = pts + rng.normal(0, 0.03, pts.shape).astype(np.float32) * np.array([1, 1, 1, 0, 0], np.float32)
    return pts


def make_batch(kind, n_frames, seed0=0, n_points=None):
    """list of frames with seeds seed0 .. seed0+n_frames-1"""
    return [make_frame(kind, seed0 + i, n_points) for i in range(n_frames)]


def synthetic_head_maps(rng, B, H, W, channels, hm_slice, n_peaks=1500, peak_logit=(-1.0, 3.0)):
    """Head maps with a realistic number of above-threshold heat-map cells (random-init heads have an
    hm bias of -2.19 and would yield no candidates)."""
    m = rng.normal(0, 0.5, (B, H, W, channels)).astype(np.float32)
    m[..., hm_slice] = rng.normal(-4.0, 0.7, m[..., hm_slice].shape).astype(np.float32)
    for b in range(B):
        ii = rng.integers(0, H, n_peaks)
        jj = rng.integers(0, W, n_peaks)
        kk = rng.integers(hm_slice.start, hm_slice.stop, n_peaks)
        m[b, ii, jj, kk] = rng.uniform(peak_logit[0], peak_logit[1], n_peaks).astype(np.float32)
    return m
