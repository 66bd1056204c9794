"""CPU tests that pin the oracle (oracle/dc_oracle.c):
  * against the committed golden fixtures produced by the compiled reference (tests/golden/make_golden.py)
  * against the compiled reference itself on fresh seeded inputs, where oracle/_ref exists
  * hand-computable known-answer cases for the semantics chosen as the parity contract
"""
import numpy as np
import pytest

from clustering_b200.synth import gaussian_mixture
from conftest import golden_cases


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


# ---------------------------------------------------------------- golden fixtures ----------------
@pytest.mark.parametrize("name", golden_cases())
def test_oracle_matches_golden(oracle, golden, name):
    g = golden(name)
    x = g["coords"]
    pops = oracle.populations(x, g["radii"])
    assert np.array_equal(pops, g["pops"])
    for r in range(len(g["radii"])):
        assert np.array_equal(bits(oracle.free_energies(pops[r])), bits(g["fe"][r]))
    fe = g["fe"][int(g["r_scr"])]
    ni, nd, hi, hd = oracle.nearest_neighbors(x, fe)
    assert np.array_equal(ni, g["nn_idx"]) and np.array_equal(hi, g["hd_idx"])
    assert np.array_equal(bits(nd), bits(g["nn_d2"])) and np.array_equal(bits(hd), bits(g["hd_d2"]))
    assert np.array_equal(oracle.sorted_free_energies(fe), g["order"].astype(np.uint64))
    assert oracle.sigma2(nd) == float(g["sigma2"])


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_screening_matches_golden(oracle, golden, name):
    g = golden(name)
    fe = g["fe"][int(g["r_scr"])]
    prev = None
    # every 4th threshold incrementally (keeps the O(M^2) restatement within seconds) ...
    for k in range(0, len(g["thresholds"]), 4):
        prev = g["labels"][k - 1].astype(np.uint64) if k > 0 else None
        lab = oracle.screening(fe, g["nn_d2"], g["thresholds"][k], g["coords"], prev)
        assert np.array_equal(lab, g["labels"][k].astype(np.uint64)), (name, k)
    # ... and one from scratch (no initial clusters): same partition, numbering by first sorted member
    k = len(g["thresholds"]) // 3
    lab = oracle.screening(fe, g["nn_d2"], g["thresholds"][k], g["coords"], None)
    ref_lab = g["labels"][k]
    assert np.array_equal(lab == 0, ref_lab == 0)
    pairs = set(zip(lab.tolist(), ref_lab.tolist()))
    assert len(pairs) == len(set(lab.tolist())) == len(set(ref_lab.tolist()))  # same partition


def test_oracle_assign_low_density_matches_golden(oracle, golden):
    for name in golden_cases():
        g = golden(name)
        fe = g["fe"][int(g["r_scr"])]
        mid = g["labels"][len(g["labels"]) // 2].astype(np.uint64)
        assigned = oracle.assign_low_density_frames(mid, g["hd_idx"], fe)
        # golden stores sorted_cluster_names(assign(...)): same partition, sizes non-increasing with rank
        final = g["microstates_from_mid"]
        pairs = set(zip(assigned.tolist(), final.tolist()))
        assert len(pairs) == len(set(assigned.tolist())) == len(set(final.tolist()))


# ---------------------------------------------------------------- live reference -----------------
@pytest.mark.ref
@pytest.mark.parametrize("d", [1, 2, 3, 4, 5, 6, 7, 8, 10, 11, 33])
def test_oracle_vs_compiled_reference(oracle, ref, d):
    n = 1500 if d <= 16 else 700
    x = gaussian_mixture(n, d, seed=1000 + d)
    x[17] = x[3]
    x[n - 1] = x[3]
    radii = np.array([0.3, 0.55, 0.15], np.float32) * np.float32(np.sqrt(d))
    po, pr = oracle.populations(x, radii), ref.populations(x, radii)
    assert np.array_equal(po.astype(np.uint64), pr)
    fe_o, fe_r = oracle.free_energies(po[0]), ref.free_energies(pr[0])
    assert np.array_equal(bits(fe_o), bits(fe_r))
    a, b = oracle.nearest_neighbors(x, fe_o), ref.nearest_neighbors(x, fe_r)
    assert np.array_equal(a[0].astype(np.uint64), b[0]) and np.array_equal(a[2].astype(np.uint64), b[2])
    assert np.array_equal(bits(a[1]), bits(b[1])) and np.array_equal(bits(a[3]), bits(b[3]))
    assert np.array_equal(oracle.sorted_free_energies(fe_o), ref.sorted_free_energies(fe_r))
    assert oracle.sigma2(a[1]) == ref.sigma2(b[0], b[1])


@pytest.mark.ref
def test_oracle_free_energy_formula_all_pops(oracle, ref):
    # every population value 1..max for several maxima (reciprocal form is the compiled one)
    for mx in (1000, 16390, 123457):
        p = np.arange(1, mx + 1, dtype=np.uint32)
        assert np.array_equal(bits(oracle.free_energies(p)), bits(ref.free_energies(p.astype(np.uint64))))
    fe = oracle.free_energies(np.array([5, 7, 7, 1], np.uint32))
    assert bits(fe)[1] == 0x80000000  # -0.0f for the maximal population


@pytest.mark.ref
def test_oracle_duplicate_radius_semantics(oracle, ref):
    # a radius listed twice is one map entry incremented twice per hit (density_clustering.cpp:131-134,180)
    x = gaussian_mixture(400, 3, seed=5)
    radii = np.array([0.3, 0.5, 0.3], np.float32)
    assert np.array_equal(oracle.populations(x, radii).astype(np.uint64), ref.populations(x, radii))


@pytest.mark.ref
def test_oracle_screening_vs_compiled_reference(oracle, ref):
    x = gaussian_mixture(1200, 4, k=6, seed=77)
    x[10] = x[2]
    fe = oracle.free_energies(oracle.populations(x, [0.35])[0])
    ni, nd, _, _ = oracle.nearest_neighbors(x, fe)
    prev_o = prev_r = None
    t = np.float32(0.1)
    while t < fe.max() + 0.1:
        lo = oracle.screening(fe, nd, t, x, prev_o)
        lr = ref.screening(fe, ni.astype(np.uint64), nd, t, x, prev_r)
        assert np.array_equal(lo, lr), float(t)
        prev_o, prev_r = lo, lr
        t = np.float32(t + np.float32(0.3))


# ---------------------------------------------------------------- known-answer cases -------------
def test_kat_dist2_order():
    from _oracle import Oracle
    o = Oracle()
    # D=1..3: plain left-to-right; exact small integers
    assert o.dist2([0], [3]) == 9
    assert o.dist2([0, 0], [3, 4]) == 25
    assert o.dist2([1, 2, 3], [4, 6, 3]) == 25
    # D=6 rounding order: ((c0^2 + c4^2) + (c1^2 + c5^2) ... ) restated by hand in float32
    rng = np.random.default_rng(3)
    for d in (4, 5, 6, 7, 8, 9, 10, 11, 12, 128):
        x = rng.standard_normal(d).astype(np.float32)
        y = rng.standard_normal(d).astype(np.float32)
        c = (x - y).astype(np.float32)
        m = (c * c).astype(np.float32)
        a = np.zeros(4, np.float32)
        k4 = 4 * (d // 4)
        for k in range(0, k4, 4):
            a = (a + m[k:k + 4]).astype(np.float32)
        l0, l1 = np.float32(a[0] + a[2]), np.float32(a[1] + a[3])
        k = k4
        if d - k >= 2:
            l0 = np.float32(l0 + m[k]); l1 = np.float32(l1 + m[k + 1]); s = np.float32(l1 + l0); k += 2
        else:
            s = np.float32(l0 + l1)
        if k < d:
            s = np.float32(s + m[k])
        assert bits(o.dist2(x, y)) == bits(s)


def test_kat_lattice_populations(oracle):
    # points 0..9 on a line with spacing 1: radius 1.5 -> 2 neighbours inside, ends 1; strict '<' at r = 1.0 -> none
    x = np.arange(10, dtype=np.float32).reshape(-1, 1)
    p = oracle.populations(x, [1.5, 1.0, 2.0])
    assert p[0].tolist() == [2, 3, 3, 3, 3, 3, 3, 3, 3, 2]
    assert p[1].tolist() == [1] * 10          # d2 == r2 is NOT inside (strict '<')
    assert p[2].tolist() == [2, 3, 3, 3, 3, 3, 3, 3, 3, 2]   # d = 2 is on the boundary -> excluded


def test_kat_nn_ties_sentinel_duplicates(oracle):
    x = np.array([[0, 0], [1, 0], [-1, 0], [0, 0], [5, 5]], np.float32)
    fe = np.array([0.5, 0.1, 0.1, 0.5, 0.0], np.float32)
    ni, nd, hi, hd = oracle.nearest_neighbors(x, fe)
    assert ni.tolist() == [3, 0, 0, 0, 1]          # duplicate is a legal neighbour (d2 = 0); ties -> smallest j
    assert nd.tolist() == [0, 1, 1, 0, 41]
    assert hi.tolist() == [1, 4, 4, 1, 6]          # strict fe[j] < fe[i]; none -> (N+1, FLT_MAX)
    assert hd[4] == np.finfo(np.float32).max and hd[1] == 41


def test_kat_single_frame(oracle):
    x = np.array([[1.0, 2.0, 3.0]], np.float32)
    assert oracle.populations(x, [0.5]).tolist() == [[1]]
    ni, nd, hi, hd = oracle.nearest_neighbors(x, np.zeros(1, np.float32))
    assert ni[0] == 2 and hi[0] == 2 and nd[0] == np.finfo(np.float32).max


def test_kat_screening_two_chains(oracle):
    # two chains of points with spacing 0.1 separated by a gap of 1.0: sigma2 = 0.01, cut = 4*sigma2 = 0.04
    # -> consecutive points connect (d2 = 0.01), the chains do not; numbering follows the lowest free energy.
    a = np.stack([np.arange(30) * 0.1, np.zeros(30)], 1)
    b = np.stack([np.arange(20) * 0.1 + 3.9, np.zeros(20)], 1)
    x = np.vstack([b, a]).astype(np.float32)                       # chain b comes first in frame order
    fe = np.concatenate([0.5 + 0.001 * np.arange(20), 0.001 * np.arange(30)]).astype(np.float32)
    _, nd, _, _ = oracle.nearest_neighbors(x, fe)
    lab = oracle.screening(fe, nd, np.float32(10.0), x, None)
    assert set(lab[20:]) == {1} and set(lab[:20]) == {2}
    # threshold below chain b's free energies: chain b stays unassigned (0)
    lab = oracle.screening(fe, nd, np.float32(0.2), x, None)
    assert set(lab[20:]) == {1} and set(lab[:20]) == {0}
