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


# ---------------------------------------------------------------- sharded scans (D <= 16) ----------
def _full_scan(sess, x, radii, fe_index):
    import torch
    n = len(x)
    sess.set_coords(x)
    pops_pos = sess.populations(radii, 0, n)
    pops = sess.to_frame_order(pops_pos)
    fe = sess.free_energies(pops[fe_index].contiguous())
    sess.nn_prepare(fe)
    keys = sess.nn_scan(0, n)
    nn = sess.nn_finish(keys)
    sess.sync()
    return pops_pos, pops, fe, keys, nn


@pytest.mark.parametrize("d,radii", [(5, [0.3]), (5, [0.1, 0.2, 0.3, 0.4, 0.5]), (10, [0.4, 0.7, 1.0, 1.3, 1.6, 2.0]), (16, [1.2, 2.0])])
def test_position_range_shards_concatenate_to_the_full_scan(oracle, d, radii):
    """dcb200_ctx_populations / dcb200_ctx_nn_scan over 2, 3 and 8 uneven position ranges -- starting on a tile, on a row
    block, or anywhere (the count / table kernels need tile-aligned groups, other ranges fall back to the histogram
    kernel) -- concatenated == the full scan == the oracle."""
    import torch
    from clustering_b200.session import Session
    n = 21_500
    x = gaussian_mixture(n, d, seed=7000 + d)
    radii = np.asarray(radii, np.float32)
    sess = Session(0)
    pops_pos, pops, fe, keys, nn = _full_scan(sess, x, radii, 0)
    want = oracle.populations(x, radii)
    assert np.array_equal(pops.cpu().numpy().view(np.uint32), want)
    fe_o = oracle.free_energies(want[0])
    assert np.array_equal(bits(fe.cpu().numpy()), bits(fe_o))
    nn_o = oracle.nearest_neighbors(x, fe_o)
    assert same_neighbours(nn_o, tuple(t.cpu().numpy().view(np.uint32) if t.dtype == torch.int32 else t.cpu().numpy() for t in nn))
    for cuts in ([0, 10_240, n], [0, 7_168, 14_000, n], [0, 1_024, 1_152, 5_000, 9_999, 10_000, 16_384, 21_499, n], [0, 333, n]):
        parts_p = [sess.populations(radii, b, e) for b, e in zip(cuts[:-1], cuts[1:])]
        parts_k = [sess.nn_scan(b, e) for b, e in zip(cuts[:-1], cuts[1:])]
        sess.sync()
        assert torch.equal(torch.cat(parts_p, dim=1), pops_pos), cuts
        assert torch.equal(torch.cat(parts_k, dim=1), keys), cuts
    sess.close()


@pytest.mark.parametrize("d,radii,n", [(5, [0.3], 40_000), (10, [0.1 * i for i in range(1, 21)], 30_011), (3, [0.15, 0.3], 9_000),
                                       (20, [2.0, 2.6, 3.2], 5_000)])        # d = 20: GEMM-form path, contiguous shards
def test_block_cyclic_shards_assemble_to_the_full_scan(oracle, d, radii, n):
    """The shard entry points the one-process-per-GPU driver uses (dcb200_ctx_*_shard, blocks of 1024 positions dealt
    round-robin) for 2, 3 and 8 emulated ranks on one device: gathered and assembled by the library's own kernels ==
    the full scan == the oracle, in frame order."""
    import torch
    from clustering_b200.session import Session
    x = gaussian_mixture(n, d, seed=7100 + d)
    radii = np.asarray(radii, np.float32)
    sess = Session(0)
    _, pops, fe, _, nn = _full_scan(sess, x, radii, 0)
    assert np.array_equal(pops.cpu().numpy().view(np.uint32), oracle.populations(x, radii))
    for w in (2, 3, 8):
        cap = sess.shard_capacity(w)
        assert sum(sess.shard_rows(r, w) for r in range(w)) == n
        # the library works on its own (non-blocking) stream: buffers are allocated up front (torch.empty launches nothing)
        # and torch only looks at the results after sess.sync()
        gp = torch.empty((w, len(radii), cap), dtype=torch.int32, device=sess.dev)            # what the all-gather delivers
        gk = torch.empty((w, 2, cap), dtype=torch.int64, device=sess.dev)
        for r in range(w):
            sess.populations_shard(radii, r, w, out=gp[r])
            sess.nn_scan_shard(r, w, out=gk[r])
        got = sess.shards_to_frame_order(gp, len(radii), w)
        got_nn = sess.nn_finish_shards(gk, w)
        sess.sync()
        assert torch.equal(got, pops), w
        for a, b in zip(got_nn, nn):
            assert torch.equal(a.view(torch.int32), b.view(torch.int32)), w
    sess.close()


def test_density_pass_from_the_default_stream(oracle):
    """clustering_b200.dist.DensityPass and ScreeningPass driven from torch's default stream (the session stream is
    non-blocking: the passes order themselves after the caller's stream and hand their results back to it)."""
    import torch
    from clustering_b200.dist import DensityPass, ScreeningPass
    from clustering_b200.session import Session
    n, d = 12_000, 4
    x = gaussian_mixture(n, d, k=5, seed=7300)
    radii = np.array([0.25, 0.4], np.float32)
    sess = Session(0)
    dp = DensityPass(sess, n, radii)
    xd = torch.from_numpy(x).cuda() * 1.0                    # produced on the default stream, consumed by the session stream
    pops, fe, nn = dp.run(xd, 1)
    pops_h = pops.cpu().numpy().view(np.uint32)              # default-stream copies: must see the finished results
    want = oracle.populations(x, radii)
    assert np.array_equal(pops_h, want)
    fe_o = oracle.free_energies(want[1])
    assert np.array_equal(bits(fe.cpu().numpy()), bits(fe_o))
    nn_o = oracle.nearest_neighbors(x, fe_o)
    assert np.array_equal(nn[0].cpu().numpy().view(np.uint32), nn_o[0]) and np.array_equal(bits(nn[3].cpu().numpy()), bits(nn_o[3]))
    # screening pass on the same data, two thresholds, comp initialised on the default stream
    order = density.sorted_free_energies(fe_o)
    xs = np.ascontiguousarray(x[order])
    cut = np.float32(4.0 * density.compute_sigma2(nn_o[1]))
    sp = ScreeningPass(sess, torch.from_numpy(xs).cuda())
    comp = torch.arange(n, dtype=torch.int32, device="cuda")
    prev_o, m_prev = None, 0
    for t in (np.float32(0.8), np.float32(2.5)):
        m_new = int(np.searchsorted(fe_o[order], t, side="right"))
        sp.step(m_prev, m_new, float(cut), comp)
        rep = comp[:m_new].cpu().numpy()
        _, lab = np.unique(rep, return_inverse=True)
        labels = np.zeros(n, np.uint32)
        labels[order[:m_new]] = lab + 1
        prev_o = oracle.screening(fe_o, nn_o[1], t, x, prev_o)
        assert np.array_equal(labels, prev_o.astype(np.uint32)), float(t)
        m_prev = m_new
    sess.close()


# ---------------------------------------------------------------- fused run, screening runs, arbitrary initial clusters ----
def test_density_run_equals_the_separate_entry_points(oracle):
    """dcb200_density_run (one upload, one layout build) == dcb200_populations + dcb200_free_energies +
    dcb200_nearest_neighbors == the oracle, including the free energies of every radius."""
    n, d = 30_000, 5
    x = gaussian_mixture(n, d, seed=8100)
    radii = np.array([0.2, 0.3, 0.45], np.float32)
    r = density.density_run(x, radii, fe_radius_index=1, neighbors=True, all_free_energies=True)
    want = oracle.populations(x, radii)
    assert np.array_equal(r["pops"], want)
    for k in range(len(radii)):
        assert np.array_equal(bits(r["fe_all"][k]), bits(oracle.free_energies(want[k])))
    fe = oracle.free_energies(want[1])
    assert np.array_equal(bits(r["fe"]), bits(fe))
    assert same_neighbours(oracle.nearest_neighbors(x, fe), r["nn"])
    assert same_neighbours(density.nearest_neighbors(x, fe), r["nn"])
    only = density.density_run(x, radii[:1], neighbors=False)
    assert only["nn"] is None and np.array_equal(only["pops"][0], want[0])


def test_screening_run_equals_call_per_threshold(oracle):
    """All thresholds through ONE screening run (sorted once, coordinates resident) == one dcb200_screening call per
    threshold with the previous labels == the oracle, at every threshold."""
    x = gaussian_mixture(7000, 3, k=6, seed=8200)
    x[15] = x[3]
    fe = oracle.free_energies(oracle.populations(x, [0.3])[0])
    _, nd, _, _ = oracle.nearest_neighbors(x, fe)
    prev_o = prev_g = None
    with density.ScreeningRun(fe, nd, x) as run:
        t = np.float32(0.1)
        while t < fe.max() + 0.2:
            lo = oracle.screening(fe, nd, t, x, prev_o)
            lr = run.next(t)
            lg = density.screening(fe, nd, t, x, prev_g)
            assert np.array_equal(lo.astype(np.uint32), lr), float(t)
            assert np.array_equal(lr, lg), float(t)
            prev_o, prev_g = lo, lg
            t = np.float32(t + np.float32(0.35))
        with pytest.raises(Exception):
            run.next(np.float32(0.05))                     # thresholds must not decrease within a run


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_screening_accepts_arbitrary_initial_clusters(oracle, seed):
    """initial_clusters that are NOT the labels of a lower threshold (density_clustering.cpp:394-427 takes any labelling):
    random names with gaps on a random subset of the frames, some of them above the threshold, some clusters far apart
    sharing a name -- labels must equal the oracle's literal restatement of the reference."""
    rng = np.random.default_rng(seed)
    n = 3000
    x = gaussian_mixture(n, 3, k=5, seed=8300 + seed)
    fe = oracle.free_energies(oracle.populations(x, [0.3])[0])
    _, nd, _, _ = oracle.nearest_neighbors(x, fe)
    init = np.zeros(n, np.uint32)
    chosen = rng.choice(n, n // 3, replace=False)
    init[chosen] = rng.choice(np.array([2, 3, 7, 11, 40], np.uint32), chosen.size)
    for t in (np.float32(1.0), np.float32(2.5), np.float32(fe.max() + 1)):
        want = oracle.screening(fe, nd, t, x, init.astype(np.uint64))
        got = density.screening(fe, nd, t, x, init)
        assert np.array_equal(got, want.astype(np.uint32)), float(t)
    # a previous result with holes punched into it
    base = density.screening(fe, nd, np.float32(1.5), x, None)
    holes = base.copy()
    holes[rng.choice(n, n // 5, replace=False)] = 0
    want = oracle.screening(fe, nd, np.float32(2.0), x, holes.astype(np.uint64))
    assert np.array_equal(density.screening(fe, nd, np.float32(2.0), x, holes), want.astype(np.uint32))


# ---------------------------------------------------------------- scheduling knobs ----------------
@pytest.mark.parametrize("knobs", [
    {"DCB200_ITEMS_PER_CTA": "1"},          # 16 column ranges per row block (round 1)
    {"DCB200_ITEMS_PER_CTA": "4096"},       # the smallest ranges the kernels allow
    {"DCB200_BIN_STEAL": "0"},              # warps only work on their own row group's units
    {"DCB200_BIN_DENSE_LANES": "1"},        # every active step through the branch-free table path
    {"DCB200_BIN_DENSE_LANES": "33"},       # every active step candidate by candidate
    {"DCB200_SUPER_PRUNE": "0"},            # no coarse level in the producers' pruning
    {"DCB200_AXIS_PRUNE": "2", "DCB200_BIN_PROJ": "2"},
    {"DCB200_NN_WINDOW": "1", "DCB200_NN_SEED_W": "1"},      # hardly any warm start of the neighbour thresholds
    {"DCB200_NN_WINDOW": "64", "DCB200_NN_SEED_W": "16"},
])
def test_scheduling_knobs_never_change_results(oracle, knobs):
    """How the pair matrix is cut into work items, which warp takes which unit, what is pruned at which level and which
    path bins a step are scheduling decisions: populations, neighbours and screening labels must not depend on them."""
    n, d = 9000, 10
    x = gaussian_mixture(n, d, seed=8400)
    radii = np.linspace(0.1, 2.0, 20, dtype=np.float32)
    want = oracle.populations(x, radii)
    fe = oracle.free_energies(want[9])
    nn_want = oracle.nearest_neighbors(x, fe)
    old = {k: os.environ.get(k) for k in knobs}
    try:
        os.environ.update(knobs)
        r = density.density_run(x, radii, 9)
        assert np.array_equal(r["pops"], want)
        assert same_neighbours(nn_want, r["nn"])
        assert np.array_equal(density.calculate_populations(x, radii[[3]]), want[[3]])        # count mode
        lab = density.screening(fe, nn_want[1], np.float32(2.0), x, None)
        assert np.array_equal(lab, oracle.screening(fe, nn_want[1], np.float32(2.0), x, None).astype(np.uint32))
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
