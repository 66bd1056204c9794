"""Seeded synthetic trajectories for tests and bench.

"HP35-like" = mixture of K isotropic Gaussians, centres U[-2.5,2.5]^D, sigma_k U[0.15,0.6],
weights proportional to U[0.1,1.1]; values stored as float32, row-major [N][D] like the
reference's Tools::read_coords result (tools.hxx:39-111).
"""
import numpy as np

# (n_rows, n_cols, radii, K, seed) of the BASELINE.json configs; fe_radius_index: the radius whose populations give the
# free energies the neighbour search runs on (C3: r = 1.0, median population ~700 of 1M -- the smallest radii leave
# almost every frame alone, the largest make whole clusters tie)
CONFIGS = {
    "C1": dict(n=100_000, d=5, radii=[0.1, 0.2, 0.3, 0.4, 0.5], k=12, seed=1, fe_radius_index=2),
    "C2": dict(n=1_000_000, d=5, radii=[0.3], k=12, seed=2, fe_radius_index=0),
    "C3": dict(n=1_000_000, d=10, radii=[round(0.1 * i, 1) for i in range(1, 21)], k=12, seed=3, fe_radius_index=9),
    "C4": dict(n=5_000_000, d=3, radii=[0.15], k=8, seed=4, fe_radius_index=0),
    "C5": dict(n=500_000, d=128, radii=[1.0], k=16, seed=5, fe_radius_index=0),
}


def gaussian_mixture(n, d, k=12, seed=1, lo=-2.5, hi=2.5, smin=0.15, smax=0.6):
    rng = np.random.Generator(np.random.PCG64(seed))
    centres = rng.uniform(lo, hi, size=(k, d))
    sig = rng.uniform(smin, smax, size=k)
    w = rng.uniform(0.1, 1.1, size=k)
    w /= w.sum()
    comp = rng.choice(k, size=n, p=w)
    x = centres[comp] + rng.standard_normal(size=(n, d)) * sig[comp, None]
    return np.ascontiguousarray(x.astype(np.float32))


def contact_like(n, d=128, k=16, seed=5, sigma=0.05):
    """C5: contact-feature-like data, centres U[0,1]^d, clipped to [0,1]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    centres = rng.uniform(0.0, 1.0, size=(k, d))
    comp = rng.integers(0, k, size=n)
    x = centres[comp] + rng.standard_normal(size=(n, d)) * sigma
    return np.ascontiguousarray(np.clip(x, 0.0, 1.0).astype(np.float32))


def config_data(name, n=None):
    c = CONFIGS[name]
    n = c["n"] if n is None else n
    if name == "C5":
        return contact_like(n, c["d"], c["k"], c["seed"])
    return gaussian_mixture(n, c["d"], c["k"], c["seed"])
