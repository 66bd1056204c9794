"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every result of libdcb200.so (through the C ABI)
is compared bit-for-bit with the golden fixtures produced by the compiled reference and with the CPU
oracle on fresh seeded inputs.  No reference sources are needed at run time."""
import numpy as np
import pytest

from clustering_b200 import density
from clustering_b200.synth import gaussian_mixture, contact_like
from conftest import golden_cases

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


# ---------------------------------------------------------------- golden fixtures ----------------
@pytest.mark.parametrize("name", golden_cases())
def test_populations_free_energies_match_golden(golden, name):
    g = golden(name)
    pops = density.calculate_populations(g["coords"], g["radii"])
    assert np.array_equal(pops, g["pops"])
    for r in range(len(g["radii"])):
        assert np.array_equal(bits(density.calculate_free_energies(pops[r])), bits(g["fe"][r]))


@pytest.mark.parametrize("name", golden_cases())
def test_nearest_neighbors_match_golden(golden, name):
    g = golden(name)
    fe = g["fe"][int(g["r_scr"])]
    ni, nd, hi, hd = density.nearest_neighbors(g["coords"], fe)
    assert np.array_equal(ni, g["nn_idx"]) and np.array_equal(hi, g["hd_idx"])
    assert np.array_equal(bits(nd), bits(g["nn_d2"])) and np.array_equal(bits(hd), bits(g["hd_d2"]))
    assert density.compute_sigma2(nd) == float(g["sigma2"])
    assert np.array_equal(density.sorted_free_energies(fe), g["order"])


@pytest.mark.parametrize("name", golden_cases())
def test_screening_matches_golden(golden, name):
    g = golden(name)
    fe = g["fe"][int(g["r_scr"])]
    prev = None
    for k, t in enumerate(g["thresholds"]):          # incremental, exactly like Density::main (:806-816)
        lab = density.screening(fe, g["nn_d2"], t, g["coords"], prev)
        assert np.array_equal(lab, g["labels"][k]), (name, k)
        prev = lab
    # from scratch at a late threshold: same labels (numbering by first sorted member)
    k = len(g["thresholds"]) - 1
    lab = density.screening(fe, g["nn_d2"], g["thresholds"][k], g["coords"], None)
    assert np.array_equal(lab, g["labels"][k])


# ---------------------------------------------------------------- oracle, fresh inputs -----------
@pytest.mark.parametrize("d", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 33])
def test_vs_oracle_all_dims(oracle, d):
    n = 3000 if d <= 16 else 1200
    x = gaussian_mixture(n, d, seed=2000 + d)
    x[17] = x[3]
    x[n - 1] = x[3]                                 # exact duplicates: d2 = 0 neighbours are legal
    radii = np.array([0.3, 0.55, 0.15, 0.3], np.float32) * np.float32(np.sqrt(d))   # unsorted, one duplicate
    po = oracle.populations(x, radii)
    pg = density.calculate_populations(x, radii)
    assert np.array_equal(po, pg)
    fe = oracle.free_energies(po[0])
    assert np.array_equal(bits(fe), bits(density.calculate_free_energies(pg[0])))
    a, b = oracle.nearest_neighbors(x, fe), density.nearest_neighbors(x, fe)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
    assert np.array_equal(bits(a[1]), bits(b[1])) and np.array_equal(bits(a[3]), bits(b[3]))


@pytest.mark.parametrize("d", [1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 16])
def test_count_mode_one_and_two_radii(oracle, d):
    # <= 2 distinct radii run the branch-free count kernel (pops_count_kernel): decisions by sign bit + exact band recheck
    n = 5000
    x = gaussian_mixture(n, d, seed=3100 + d)
    x[100] = x[7]
    x[n - 2] = x[7]
    s = np.float32(np.sqrt(d))
    for radii in ([0.3 * s], [0.25 * s, 0.5 * s], [0.4 * s, 0.4 * s, 0.2 * s], [0.0], [0.0, 0.3 * s]):
        radii = np.array(radii, np.float32)
        assert np.array_equal(oracle.populations(x, radii), density.calculate_populations(x, radii)), (d, radii)


def test_count_mode_boundary_lattice_and_offset(oracle):
    g = np.stack(np.meshgrid(np.arange(14), np.arange(14), np.arange(14), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    for radii in ([1.0], [np.sqrt(2.0)], [2.0, 3.0], [np.sqrt(5.0), 1.0]):
        radii = np.array(radii, np.float32)
        po, pg = oracle.populations(g, radii), density.calculate_populations(g, radii)
        assert np.array_equal(po, pg), radii
    assert density.calculate_populations(g, [1.0]).max() == 1
    x = gaussian_mixture(4000, 3, seed=99) + np.float32(500.0)          # far from the origin: wide error band, many rechecks
    for radii in ([0.2], [0.1, 0.3]):
        assert np.array_equal(oracle.populations(x, np.array(radii, np.float32)), density.calculate_populations(x, radii))


def test_count_mode_medium_vs_oracle(oracle):
    # several row blocks / work items / pruned tiles; n not a multiple of the padding
    x = gaussian_mixture(30011, 5, seed=41)
    for radii in ([0.3], [0.2, 0.45]):
        radii = np.array(radii, np.float32)
        assert np.array_equal(oracle.populations(x, radii), density.calculate_populations(x, radii))
    x = gaussian_mixture(4096, 3, k=3, seed=42)                        # n == padded size: clamped rows are real rows
    assert np.array_equal(oracle.populations(x, np.array([0.15], np.float32)), density.calculate_populations(x, [0.15]))


def test_offset_data_and_many_radii(oracle):
    # far from the origin (the centring of the fast path must not change a single count) and > 31 radii (two passes)
    x = gaussian_mixture(2500, 4, seed=77) + np.float32(1000.0)
    radii = np.linspace(0.05, 1.5, 40).astype(np.float32)
    assert np.array_equal(oracle.populations(x, radii), density.calculate_populations(x, radii))


def test_boundary_pairs_lattice(oracle):
    # integer lattice: many pairs sit exactly ON the radius (d2 == r2 must not count: strict '<')
    g = np.stack(np.meshgrid(np.arange(12), np.arange(12), np.arange(12), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    radii = np.array([1.0, np.sqrt(2.0), 2.0, 3.0, 2.5], np.float32)
    po, pg = oracle.populations(g, radii), density.calculate_populations(g, radii)
    assert np.array_equal(po, pg)
    assert pg[0].max() == 1                          # d == 1.0 exactly is not inside
    fe = oracle.free_energies(po[3])
    a, b = oracle.nearest_neighbors(g, fe), density.nearest_neighbors(g, fe)      # massive ties: smallest index wins
    for u, v in zip(a, b):
        assert np.array_equal(bits(u) if u.dtype == np.float32 else u, bits(v) if v.dtype == np.float32 else v)


def test_kat_small_cases():
    x = np.arange(10, dtype=np.float32).reshape(-1, 1)
    p = density.calculate_populations(x, [1.5, 1.0, 2.0])
    assert p[0].tolist() == [2, 3, 3, 3, 3, 3, 3, 3, 3, 2]
    assert p[1].tolist() == [1] * 10
    assert p[2].tolist() == [2, 3, 3, 3, 3, 3, 3, 3, 3, 2]
    x = np.array([[0, 0], [1, 0], [-1, 0], [0, 0], [5, 5]], np.float32)
    fe = np.array([0.5, 0.1, 0.1, 0.5, 0.0], np.float32)
    ni, nd, hi, hd = density.nearest_neighbors(x, fe)
    assert ni.tolist() == [3, 0, 0, 0, 1] and nd.tolist() == [0, 1, 1, 0, 41]
    assert hi.tolist() == [1, 4, 4, 1, 6]
    assert hd[4] == np.finfo(np.float32).max and hd[1] == 41
    one = np.array([[1.0, 2.0, 3.0]], np.float32)
    assert density.calculate_populations(one, [0.5]).tolist() == [[1]]
    ni, nd, hi, hd = density.nearest_neighbors(one, np.zeros(1, np.float32))
    assert ni[0] == 2 and hi[0] == 2 and nd[0] == np.finfo(np.float32).max
    fe = density.calculate_free_energies(np.array([5, 7, 7, 1], np.uint32))
    assert bits(fe)[1] == 0x80000000


def test_free_energy_formula_all_pops(oracle):
    for mx in (1000, 16390, 123457, 1000000, 4999999):
        p = np.arange(1, mx + 1, dtype=np.uint32)
        assert np.array_equal(bits(density.calculate_free_energies(p)), bits(oracle.free_energies(p))), mx


def test_screening_vs_oracle_fresh(oracle):
    x = gaussian_mixture(1500, 4, k=6, seed=78)
    x[10] = x[2]
    fe = oracle.free_energies(oracle.populations(x, [0.35])[0])
    _, nd, _, _ = oracle.nearest_neighbors(x, fe)
    prev_o = prev_g = None
    t = np.float32(0.1)
    while t < fe.max() + 0.1:
        lo = oracle.screening(fe, nd, t, x, prev_o)
        lg = density.screening(fe, nd, t, x, prev_g)
        assert np.array_equal(lo.astype(np.uint32), lg), float(t)
        prev_o, prev_g = lo, lg
        t = np.float32(t + np.float32(0.3))


def test_high_dim_contact_like(oracle):
    x = contact_like(900, 128, k=6, seed=21)
    radii = np.array([1.0, 0.8], np.float32)
    po = oracle.populations(x, radii)
    assert np.array_equal(po, density.calculate_populations(x, radii))
    fe = oracle.free_energies(po[0])
    a, b = oracle.nearest_neighbors(x, fe), density.nearest_neighbors(x, fe)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
    assert np.array_equal(bits(a[1]), bits(b[1])) and np.array_equal(bits(a[3]), bits(b[3]))


# ---------------------------------------------------------------- size-independent properties ----
def test_medium_size_properties(oracle):
    """60k x 5 (several row blocks, column items and tiles): sampled rows against the oracle's scalar
    distance, symmetry of the counts, and NN consistency."""
    n, d = 60000, 5
    x = gaussian_mixture(n, d, seed=5)
    radii = np.array([0.1, 0.2, 0.3, 0.4, 0.5], np.float32)
    pops = density.calculate_populations(x, radii)
    rng = np.random.default_rng(0)
    rows = rng.choice(n, 24, replace=False)
    r2 = (radii * radii).astype(np.float32)
    for i in rows:
        diff = (x - x[i]).astype(np.float32)
        approx = (diff * diff).sum(1)
        cand = np.nonzero(approx < r2.max() * 1.001 + 1e-6)[0]
        exact = np.array([oracle.dist2(x[i], x[j]) for j in cand], np.float32)
        for r in range(len(radii)):
            assert pops[r, i] == 1 + int(np.count_nonzero((exact < r2[r]) & (cand != i))), (i, r)
    assert int(pops[:, :].astype(np.int64).sum() - pops.size) % 2 == 0     # every pair is counted from both ends
    fe = density.calculate_free_energies(pops[2])
    ni, nd, hi, hd = density.nearest_neighbors(x, fe)
    for i in rows:
        diff = (x - x[i]).astype(np.float32)
        approx = (diff * diff).sum(1)
        approx[i] = np.inf
        cand = np.nonzero(approx <= approx.min() * 1.001 + 1e-7)[0]
        exact = np.array([oracle.dist2(x[i], x[j]) for j in cand], np.float32)
        j = cand[np.flatnonzero(exact == exact.min())[0]]
        assert ni[i] == j and bits(nd[i]) == bits(exact.min())
        lower = np.nonzero(fe < fe[i])[0]
        if lower.size:
            a2 = approx[lower]
            c2 = lower[np.nonzero(a2 <= a2.min() * 1.001 + 1e-7)[0]]
            e2 = np.array([oracle.dist2(x[i], x[j]) for j in c2], np.float32)
            assert hi[i] == c2[np.flatnonzero(e2 == e2.min())[0]] and bits(hd[i]) == bits(e2.min())
        else:
            assert hi[i] == n + 1
    assert np.all(hd >= nd)


def test_sharded_screening_forests_merge_to_the_oracle_labels(oracle):
    """The session-level screening the one-process-per-GPU driver uses (clustering_b200/dist.py: ScreeningPass): three
    'ranks' scan their share of the new rows into their own forests, the forests are unioned on the device, and the
    result must give the oracle's labels at every threshold."""
    import torch
    from clustering_b200.dist import screen_cuts
    from clustering_b200.session import Session
    x = gaussian_mixture(5000, 3, k=6, seed=404)
    fe = oracle.free_energies(oracle.populations(x, [0.3])[0])
    _, nd, _, _ = oracle.nearest_neighbors(x, fe)
    order = density.sorted_free_energies(fe)
    xs = np.ascontiguousarray(x[order])
    fes = fe[order]
    cut = np.float32(4.0 * density.compute_sigma2(nd))
    G = 3
    sess = [Session(0) for _ in range(G)]
    for s in sess:
        s.set_coords(xs, keep_order=True)
    dev = sess[0].dev
    comp = torch.arange(len(x), dtype=torch.int32, device=dev)
    prev_o, m_prev = None, 0
    t = np.float32(0.2)
    while t < fe.max() + 0.3:
        m_new = int(np.searchsorted(fes, t, side="right"))
        cuts = screen_cuts(m_prev, m_new, G)
        forests = []
        for g in range(G):
            c = comp.clone()
            c[m_prev:m_new] = torch.arange(m_prev, m_new, dtype=torch.int32, device=dev)
            sess[g].screening_scan(m_prev, m_new, float(cut), c, cuts[g], cuts[g + 1])
            sess[g].screening_flatten(m_new, c)
            sess[g].sync()
            forests.append(c)
        comp = forests[0]
        for g in range(1, G):
            sess[0].screening_merge(m_new, comp, forests[g])
        sess[0].screening_flatten(m_new, comp)
        sess[0].sync()
        rep = comp[:m_new].cpu().numpy()
        roots, lab = np.unique(rep, return_inverse=True)        # ascending representative = the reference's numbering
        labels = np.zeros(len(x), np.uint32)
        labels[order[:m_new]] = lab + 1
        prev_o = oracle.screening(fe, nd, t, x, prev_o)
        assert np.array_equal(labels, prev_o.astype(np.uint32)), float(t)
        m_prev = m_new
        t = np.float32(t + np.float32(0.4))
    for s in sess:
        s.close()


@pytest.mark.parametrize("name", ["C2", "C5"])
def test_full_size_properties(oracle, name):
    """BASELINE.json's full sizes (C2: 1M x 5 on the FFMA kernels, C5: 500k x 128 on the tensor cores): size-independent
    properties -- every pair is counted from both ends, sampled rows against the oracle's scalar distance over ALL frames,
    nearest neighbours of sampled rows, and the lower-free-energy neighbour really has a lower free energy."""
    from clustering_b200.synth import CONFIGS, config_data
    cfg = CONFIGS[name]
    x = config_data(name)
    n, d = x.shape
    radii = np.asarray(cfg["radii"][:1], np.float32)
    pops = density.calculate_populations(x, radii)[0]
    assert int(pops.astype(np.int64).sum() - n) % 2 == 0
    fe = density.calculate_free_energies(pops)
    ni, nd, hi, hd = density.nearest_neighbors(x, fe)
    rng = np.random.default_rng(7)
    r2 = np.float32(radii[0] * radii[0])
    for i in rng.choice(n, 12, replace=False):
        diff = x - x[i]
        approx = np.einsum("ij,ij->i", diff, diff)
        cand = np.nonzero(approx < r2 * np.float32(1.001) + np.float32(1e-6))[0]
        exact = np.array([oracle.dist2(x[i], x[j]) for j in cand], np.float32)
        assert pops[i] == 1 + int(np.count_nonzero((exact < r2) & (cand != i))), (name, i)
        approx[i] = np.inf
        c2 = np.nonzero(approx <= approx.min() * np.float32(1.001) + np.float32(1e-7))[0]
        e2 = np.array([oracle.dist2(x[i], x[j]) for j in c2], np.float32)
        assert ni[i] == c2[np.flatnonzero(e2 == e2.min())[0]] and bits(nd[i]) == bits(e2.min()), (name, i)
        lower = np.nonzero(fe < fe[i])[0]
        if lower.size:
            a2 = approx[lower]
            c3 = lower[np.nonzero(a2 <= a2.min() * np.float32(1.001) + np.float32(1e-7))[0]]
            e3 = np.array([oracle.dist2(x[i], x[j]) for j in c3], np.float32)
            assert hi[i] == c3[np.flatnonzero(e3 == e3.min())[0]] and bits(hd[i]) == bits(e3.min()), (name, i)
        else:
            assert hi[i] == n + 1
    has = hi <= n
    assert np.all(fe[hi[has]] < fe[has]) and np.all(hd >= nd)


def test_full_size_screening_properties():
    """C4 (5M x 3): the incremental screening over rising thresholds gives the same labels as one from-scratch call at the
    last threshold; labels are 1..K without gaps below the threshold and 0 above it; clusters only ever merge or grow."""
    from clustering_b200.synth import CONFIGS, config_data
    cfg = CONFIGS["C4"]
    x = config_data("C4")
    pops = density.calculate_populations(x, np.asarray(cfg["radii"][:1], np.float32))[0]
    fe = density.calculate_free_energies(pops)
    _, nd, _, _ = density.nearest_neighbors(x, fe)
    prev = None
    for t in (np.float32(0.4), np.float32(1.2), np.float32(2.0)):
        lab = density.screening(fe, nd, t, x, prev)
        below = fe <= t
        assert np.all(lab[~below] == 0) and np.all(lab[below] > 0)
        k = int(lab.max())
        assert np.array_equal(np.unique(lab[below]), np.arange(1, k + 1, dtype=lab.dtype))
        if prev is not None:
            old = prev > 0                              # frames clustered before stay together
            pairs = np.unique(np.stack([prev[old], lab[old]], 1), axis=0)
            assert len(np.unique(pairs[:, 0])) == len(pairs)      # an old cluster maps to exactly one new cluster
        prev = lab
    scratch = density.screening(fe, nd, np.float32(2.0), x, None)
    assert np.array_equal(scratch, prev)
