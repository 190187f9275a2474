"""Seeded synthetic inputs shared by the CPU and GPU tests (numpy only)."""
import numpy as np


def uniform_spheres(n, seed=0, rmax=0.1):
    """tests/hitcounts/hitcounts.cu:52-54: centres U[0,1)^3, radii U[0,rmax)."""
    rng = np.random.default_rng(seed)
    s = rng.random((n, 4), dtype=np.float32)
    s[:, 3] *= np.float32(rmax)
    return s


def clustered_spheres(n, seed=0, n_halos=16):
    """Small Gadget-like set: 30 % uniform + 70 % in Plummer-ish halos, h from local density."""
    rng = np.random.default_rng(seed)
    n_bg = int(0.3 * n)
    pos = np.empty((n, 3), np.float64)
    pos[:n_bg] = rng.random((n_bg, 3))
    centres = rng.random((n_halos, 3))
    scale = 10 ** rng.uniform(np.log10(0.004), np.log10(0.03), n_halos)
    which = rng.integers(0, n_halos, n - n_bg)
    u = rng.random(n - n_bg) * 0.95 + 1e-4
    r = scale[which] / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
    v = rng.normal(size=(n - n_bg, 3))
    v /= np.linalg.norm(v, axis=1)[:, None]
    pos[n_bg:] = centres[which] + v * r[:, None]
    pos = np.mod(pos, 1.0)
    # smoothing length from a crude local density (uniform part + halo profile)
    dens = np.full(n, 0.3 * n)
    for h in range(n_halos):
        d2 = ((pos - centres[h]) ** 2).sum(1)
        m = 0.7 * n * (which == h).sum() / max(1, n - n_bg)
        a = scale[h]
        dens += 3 * m / (4 * np.pi * a ** 3) * (1 + d2 / a ** 2) ** -2.5
    hsml = np.clip((3 * 32 / (4 * np.pi * dens)) ** (1 / 3), 1e-5, 0.1)
    s = np.empty((n, 4), np.float32)
    s[:, :3] = pos
    s[:, 3] = hsml
    s[:, :3] = np.minimum(s[:, :3], np.float32(0.99999994))
    return s


def isotropic_rays(n, origin=(0.5, 0.5, 0.5), length=2.0, seed=1, sort=True):
    """Isotropic rays from one origin, ordered by the direction Morton key like
    uniform_random_rays (cuda/kernels/gen_rays.cuh:126-161,483)."""
    rng = np.random.default_rng(seed)
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1)[:, None]
    rays = np.zeros((n, 7), np.float32)
    rays[:, :3] = v
    rays[:, 3:6] = np.asarray(origin, np.float32)
    rays[:, 6] = length
    if sort:
        q = np.clip(((rays[:, :3] + 1) / 2 * 1023).astype(np.int64), 0, 1023)
        key = np.zeros(n, np.int64)
        for b in range(10):
            for k in range(3):
                key |= ((q[:, k] >> b) & 1) << (3 * b + k)
        rays = rays[np.argsort(key, kind="stable")]
    return np.ascontiguousarray(rays)


def ortho_rays_z(n_side, lo=0.0, hi=1.0):
    """Pixel-centre rays in -z over [lo,hi]^2 (tests/helper/rays.cuh:55-79 geometry)."""
    c = (np.arange(n_side, dtype=np.float64) + 0.5) / n_side * (hi - lo) + lo
    x, y = np.meshgrid(c, c[::-1])
    rays = np.zeros((n_side * n_side, 7), np.float32)
    rays[:, 2] = -1.0
    rays[:, 3] = x.ravel()
    rays[:, 4] = y.ravel()
    rays[:, 5] = hi + (hi - lo) * 0.5
    rays[:, 6] = 2.0 * (hi - lo)
    return rays
