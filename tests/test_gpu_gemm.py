"""GPU parity tests of the GEMM-form (tcgen05 tensor core) path that serves 17 <= n_cols <= 256
(clustering_b200/csrc/gemm_kernels.cuh).  Everything goes through the C ABI and is compared bit for bit with the
CPU oracle; the tensor-core value only filters, so populations, neighbour indices and squared distances must be
identical to the reference's arithmetic."""
import os

import numpy as np
import pytest
import torch

from clustering_b200 import density
from clustering_b200.session import Session
from clustering_b200.synth import gaussian_mixture, contact_like

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same_nn(a, b):
    return (np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2]) and np.array_equal(bits(a[1]), bits(b[1]))
            and np.array_equal(bits(a[3]), bits(b[3])))


def test_gemm_path_is_active_and_bounded():
    """17 <= d <= 256 in spatial order runs on the tensor cores, other inputs do not; with DCB200_GEMM_CHECK=1 every pair
    is also evaluated exactly and the observed |fast - exact| must stay inside the proven band (ratio < 1)."""
    s = Session(0)
    for d, want in ((16, False), (17, True), (32, True), (128, True), (256, True), (257, False)):
        s.set_coords(gaussian_mixture(700, d, seed=d))
        assert s.gemm_info()[0] == want, d
    s.set_coords(gaussian_mixture(700, 64, seed=1), keep_order=True)
    assert not s.gemm_info()[0]
    os.environ["DCB200_GEMM_CHECK"] = "1"
    try:
        worst = 0.0
        for d, data in ((17, gaussian_mixture(3000, 17, k=4, seed=3) * np.float32(7.0)), (24, contact_like(3000, 24, k=4, seed=4) - np.float32(9.0)),
                        (32, gaussian_mixture(3000, 32, k=4, seed=5)), (128, contact_like(3000, 128, k=4, seed=6)),
                        (200, gaussian_mixture(2000, 200, k=3, seed=7) + np.float32(40.0)),
                        (256, contact_like(2000, 256, k=3, seed=8) * np.float32(30.0))):
            s.set_coords(data)
            r = float(np.sqrt(np.median(((data[:200, None, :] - data[None, :200, :]) ** 2).sum(-1))))
            s.populations([r])
            s.sync()
            active, ratio = s.gemm_info()
            assert active and 0.0 < ratio < 1.0, (d, ratio)
            worst = max(worst, ratio)
        print(f"max observed |fast - exact| / band = {worst:.3f}")
    finally:
        del os.environ["DCB200_GEMM_CHECK"]
    s.close()


@pytest.mark.parametrize("d", [17, 20, 24, 31, 32, 33, 40, 64, 100, 128, 129, 200, 256])
def test_gemm_vs_oracle_dims(oracle, d):
    n = 1300                                            # not a multiple of the tile: padded rows and columns
    x = contact_like(n, d, k=4, seed=700 + d) if d % 2 == 0 else gaussian_mixture(n, d, k=4, seed=700 + d)
    x[17] = x[3]
    x[n - 1] = x[3]                                     # exact duplicates: d2 = 0 neighbours are legal
    dd = ((x[:300, None, :] - x[None, :300, :]) ** 2).sum(-1)
    r_in = float(np.sqrt(np.percentile(dd[dd > 0], 20)))
    for radii in ([r_in], [r_in, 0.7 * r_in, r_in], np.linspace(0.3 * r_in, 3.0 * r_in, 9), [0.0, r_in]):
        radii = np.array(radii, np.float32)
        po, pg = oracle.populations(x, radii), density.calculate_populations(x, radii)
        assert np.array_equal(po, pg), (d, radii)
    po = oracle.populations(x, np.array([r_in], np.float32))
    fe = oracle.free_energies(po[0])
    assert same_nn(oracle.nearest_neighbors(x, fe), density.nearest_neighbors(x, fe))


def test_gemm_binary_lattice_boundary_pairs(oracle):
    # binary vectors: squared distances are integers (Hamming distances), thousands of pairs sit exactly ON the
    # radius; strict '<' must hold although the tensor-core value is only TF32-accurate.  Massive neighbour ties.
    rng = np.random.default_rng(3)
    base = rng.integers(0, 2, size=(6, 48))
    x = base[rng.integers(0, 6, size=2000)].copy()
    flip = rng.random(x.shape) < 0.06
    x = np.ascontiguousarray(np.where(flip, 1 - x, x).astype(np.float32))
    radii = np.array([2.0, 3.0, np.sqrt(5.0), 4.0, 1.0], np.float32)
    po, pg = oracle.populations(x, radii), density.calculate_populations(x, radii)
    assert np.array_equal(po, pg)
    fe = oracle.free_energies(po[1])
    assert same_nn(oracle.nearest_neighbors(x, fe), density.nearest_neighbors(x, fe))
    y = x + np.float32(1000.0)                          # far from the origin: centring must not change a single count
    assert np.array_equal(oracle.populations(y, radii), density.calculate_populations(y, radii))


def test_gemm_medium_pruned_vs_oracle(oracle):
    # many row tiles, several work items per row tile, well separated clusters (tile pairs are pruned), n % 128 != 0
    x = contact_like(12011, 64, k=10, seed=91)
    radii = np.array([0.35, 0.45], np.float32)
    po, pg = oracle.populations(x, radii), density.calculate_populations(x, radii)
    assert np.array_equal(po, pg)
    assert po[1].max() > 100
    fe = oracle.free_energies(po[0])
    assert same_nn(oracle.nearest_neighbors(x, fe), density.nearest_neighbors(x, fe))


def test_gemm_matches_ffma_path_and_shards():
    """Larger input than the oracle handles quickly: tensor-core path against the run-time-D FFMA kernels of the same library
    (DCB200_GEMM=0), and row shards (multiples of 128: tensor cores; anything else: FFMA kernels) against the full scan."""
    x = contact_like(40000, 128, k=12, seed=17)
    radii = np.array([0.9, 1.0, 0.7], np.float32)
    pg = density.calculate_populations(x, radii)
    fe = density.calculate_free_energies(pg[1])
    ng = density.nearest_neighbors(x, fe)
    os.environ["DCB200_GEMM"] = "0"
    try:
        pf = density.calculate_populations(x, radii)
        nf = density.nearest_neighbors(x, fe)
    finally:
        del os.environ["DCB200_GEMM"]
    assert np.array_equal(pg, pf)
    assert same_nn(ng, nf)
    s = Session(0)
    s.set_coords(x)
    assert s.gemm_info()[0]
    full = s.populations(radii)
    # 0: two row tiles per item; 12928 = 101 x 128: one row tile per item (odd tile boundary); 25000: not a tile boundary -> FFMA kernels
    parts = [s.populations(radii, 0, 12928), s.populations(radii, 12928, 25000), s.populations(radii, 25000, 40000)]
    assert torch.equal(full, torch.cat(parts, dim=1))
    fe_dev = torch.from_numpy(fe).to(full.device)
    s.nn_prepare(fe_dev)
    kfull = s.nn_scan()
    kparts = torch.cat([s.nn_scan(0, 20480), s.nn_scan(20480, 30848), s.nn_scan(30848, 40000)], dim=1)      # 30848 = 241 x 128
    assert torch.equal(kfull, kparts)
    s.close()


def test_gemm_edge_cases(oracle):
    """Partial single tile, exactly one tile, a single frame, all frames identical, a radius far below the TF32 resolution
    of the data (every pair goes through the exact path), a radius that contains everything, NaN input."""
    from clustering_b200 import lib
    for n in (1, 2, 97, 128, 129, 256, 257):
        x = contact_like(n, 48, k=2, seed=n)
        radii = np.array([0.45, 1e-4, 50.0], np.float32)
        po, pg = oracle.populations(x, radii), density.calculate_populations(x, radii)
        assert np.array_equal(po, pg), n
        fe = oracle.free_energies(po[0])
        assert same_nn(oracle.nearest_neighbors(x, fe), density.nearest_neighbors(x, fe)), n
    x = np.tile(contact_like(1, 64, seed=3), (700, 1))                     # 700 identical frames: every distance is 0
    po, pg = oracle.populations(x, [0.1, 0.0]), density.calculate_populations(x, [0.1, 0.0])
    assert np.array_equal(po, pg) and pg[0].min() == 700 and pg[1].max() == 1
    fe = np.zeros(700, np.float32)
    fe[::7] = 1.0
    assert same_nn(oracle.nearest_neighbors(x, fe), density.nearest_neighbors(x, fe))
    x = gaussian_mixture(900, 40, k=3, seed=12) * np.float32(1e3) + np.float32(3e4)   # large scale: wide bands, many rechecks
    dd = ((x[:200, None, :] - x[None, :200, :]) ** 2).sum(-1)
    r = float(np.sqrt(np.percentile(dd[dd > 0], 10)))
    assert np.array_equal(oracle.populations(x, [r, 0.5 * r]), density.calculate_populations(x, [r, 0.5 * r]))
    bad = contact_like(300, 64, seed=1)
    bad[17, 5] = np.nan
    with pytest.raises(lib.Dcb200Error):
        density.calculate_populations(bad, [0.5])
