"""GPU parity tests in the shape of the BASELINE.json configs and of the paths the benchmark times (pytest -m gpu):

  * C1 (100k x 5, radii 0.1..0.5) in full against the oracle: every population, free energy, neighbour index and d2 bit;
  * C3 shape (100k x 10, the 20 radii) in full against the oracle -- the table-driven multi-radius kernel (pops_bin_kernel)
    over several row blocks, pruned tiles, dense and sparse steps -- and C3 at its full 1M frames through sampled rows;
  * the multi-radius kernel on every specialised dimension, with radius lists that stress its cell table;
  * sharded scans: position ranges (contiguous, uneven, tile-aligned or not) and the block-cyclic shards the
    one-process-per-GPU driver uses must concatenate to the full scan, for D <= 16.
"""
import os

import numpy as np
import pytest

from clustering_b200 import density
from clustering_b200.synth import CONFIGS, config_data, gaussian_mixture

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same_neighbours(a, b):
    return (np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2]) and np.array_equal(bits(a[1]), bits(b[1]))
            and np.array_equal(bits(a[3]), bits(b[3])))


# ---------------------------------------------------------------- BASELINE configs in full --------
def test_c1_full_differential(oracle):
    """BASELINE configs[0]: 100k frames x 5 dims, radii 0.1 .. 0.5: populations + free energies for every radius,
    plus the neighbour search on the middle radius, all compared value by value with the oracle."""
    cfg = CONFIGS["C1"]
    x = config_data("C1")
    radii = np.asarray(cfg["radii"], np.float32)
    po = oracle.populations(x, radii)
    pg = density.calculate_populations(x, radii)
    assert np.array_equal(po, pg)
    for r in range(len(radii)):
        assert np.array_equal(bits(oracle.free_energies(po[r])), bits(density.calculate_free_energies(pg[r])))
    fe = oracle.free_energies(po[2])
    assert same_neighbours(oracle.nearest_neighbors(x, fe), density.nearest_neighbors(x, fe))


def test_c3_shape_full_differential(oracle):
    """C3's shape at 100k frames: 10 dims, the 20 radii 0.1 .. 2.0 (98 row blocks, tiles pruned by box and sphere, dense
    and sparse steps of the table-driven kernel), every count against the oracle; neighbours on r = 1.0."""
    cfg = CONFIGS["C3"]
    x = config_data("C3", 100_000)
    radii = np.asarray(cfg["radii"], np.float32)
    po = oracle.populations(x, radii)
    pg = density.calculate_populations(x, radii)
    assert np.array_equal(po, pg)
    r_fe = cfg["fe_radius_index"]
    fe = oracle.free_energies(po[r_fe])
    assert np.array_equal(bits(fe), bits(density.calculate_free_energies(pg[r_fe])))
    assert same_neighbours(oracle.nearest_neighbors(x, fe), density.nearest_neighbors(x, fe))


def test_c3_full_size_sampled_rows(oracle):
    """C3 at its full size (1M x 10, 20 radii): sampled rows against the oracle's scalar distance over ALL frames, the
    symmetry of the counts, and the neighbours of the sampled rows."""
    cfg = CONFIGS["C3"]
    x = config_data("C3")
    n, d = x.shape
    radii = np.asarray(cfg["radii"], np.float32)
    pops = density.calculate_populations(x, radii)
    r2 = (radii * radii).astype(np.float32)
    assert np.all(np.diff(pops.astype(np.int64), axis=0) >= 0)              # radii ascend: counts are nested
    for r in range(len(radii)):
        assert int(pops[r].astype(np.int64).sum() - n) % 2 == 0             # every pair is counted from both ends
    fe = density.calculate_free_energies(pops[cfg["fe_radius_index"]])
    ni, nd, hi, hd = density.nearest_neighbors(x, fe)
    rng = np.random.default_rng(11)
    for i in rng.choice(n, 10, replace=False):
        diff = x - x[i]
        approx = np.einsum("ij,ij->i", diff, diff)
        cand = np.nonzero(approx < r2.max() * np.float32(1.001) + np.float32(1e-6))[0]
        exact = np.array([oracle.dist2(x[i], x[j]) for j in cand], np.float32)
        for r in range(len(radii)):
            assert pops[r, i] == 1 + int(np.count_nonzero((exact < r2[r]) & (cand != i))), (i, r)
        approx[i] = np.inf
        c2 = np.nonzero(approx <= approx.min() * np.float32(1.001) + np.float32(1e-7))[0]
        e2 = np.array([oracle.dist2(x[i], x[j]) for j in c2], np.float32)
        assert ni[i] == c2[np.flatnonzero(e2 == e2.min())[0]] and bits(nd[i]) == bits(e2.min()), i
        lower = np.nonzero(fe < fe[i])[0]
        if lower.size:
            a2 = approx[lower]
            c3 = lower[np.nonzero(a2 <= a2.min() * np.float32(1.001) + np.float32(1e-7))[0]]
            e3 = np.array([oracle.dist2(x[i], x[j]) for j in c3], np.float32)
            assert hi[i] == c3[np.flatnonzero(e3 == e3.min())[0]] and bits(hd[i]) == bits(e3.min()), i
        else:
            assert hi[i] == n + 1
    has = hi <= n
    assert np.all(fe[hi[has]] < fe[has]) and np.all(hd >= nd)


# ---------------------------------------------------------------- the multi-radius kernel ---------
@pytest.mark.parametrize("d", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16])
def test_bin_mode_all_dims(oracle, d):
    n = 4000
    x = gaussian_mixture(n, d, seed=5100 + d)
    x[33] = x[4]
    x[n - 1] = x[4]                                    # exact duplicates
    s = np.float32(np.sqrt(d))
    lists = [
        np.linspace(0.05, 1.0, 20, dtype=np.float32) * s,                      # the C3 pattern
        np.array([0.9, 0.1, 0.5, 0.5, 0.3, 0.7, 0.2, 0.0], np.float32) * s,    # unsorted, a duplicate, a zero radius
        np.linspace(0.02, 1.3, 31, dtype=np.float32) * s,                      # a full pass
    ]
    for radii in lists:
        assert np.array_equal(oracle.populations(x, radii), density.calculate_populations(x, radii)), (d, radii)


def test_bin_mode_table_stress(oracle):
    """Radius lists at the limits of the cell table: more than one pass (> 31 radii), radii so close that the table needs
    its finest resolution, radii too close for any table (the histogram kernel takes over), tiny radii next to a huge
    one, and data far from the origin (wide error bands)."""
    x = gaussian_mixture(6000, 5, seed=88)
    base = np.float32(0.6)
    cases = [
        np.linspace(0.03, 1.6, 45, dtype=np.float32),                                      # two passes
        np.array([0.6, 0.6003, 0.61, 0.3, 0.9, 1.2], np.float32),                          # fine table
        np.array([base, np.nextafter(base, np.float32(1)), 0.2, 0.4, 0.8], np.float32),    # no table: 1 ulp apart
        np.array([0.001, 0.002, 0.004, 3.0, 0.5], np.float32),                             # six octaves apart
    ]
    for radii in cases:
        assert np.array_equal(oracle.populations(x, radii), density.calculate_populations(x, radii)), radii
    y = gaussian_mixture(3000, 4, seed=89) + np.float32(1000.0)
    radii = np.linspace(0.05, 1.2, 12, dtype=np.float32)
    assert np.array_equal(oracle.populations(y, radii), density.calculate_populations(y, radii))


def test_bin_mode_lattice_on_the_radii(oracle):
    """Integer lattice: thousands of pairs sit exactly ON several of the radii (strict '<' through the table's band)."""
    g = np.stack(np.meshgrid(np.arange(13), np.arange(13), np.arange(13), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    radii = np.array([1.0, np.sqrt(2.0), np.sqrt(3.0), 2.0, np.sqrt(5.0), np.sqrt(6.0), 3.0, 2.5], np.float32)
    assert np.array_equal(oracle.populations(g, radii), density.calculate_populations(g, radii))


@pytest.mark.parametrize("mode", ["count", "bin", "hist"])
def test_population_kernels_agree(oracle, mode):
    """The three population kernels are interchangeable: each one forced in turn on the same input."""
    x = gaussian_mixture(20011, 6, seed=61)
    radii = np.array([0.2, 0.35, 0.5, 0.65, 0.8], np.float32)
    want = oracle.populations(x, radii)
    old = os.environ.get("DCB200_POPS_MODE")
    os.environ["DCB200_POPS_MODE"] = mode
    try:
        got = density.calculate_populations(x, radii)
    finally:
        if old is None:
            del os.environ["DCB200_POPS_MODE"]
        else:
            os.environ["DCB200_POPS_MODE"] = old
    assert np.array_equal(want, got)
