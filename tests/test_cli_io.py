"""File formats and command-line front end of `clustering density` (CPU-only tests: no compute calls).

The writers are compared byte for byte with files produced by the reference's own writers (tools.cpp, compiled
unmodified into oracle/_ref/libdcref.so): live where that library exists, and always against the copies committed
under tests/golden/io/ (made by tests/golden/make_golden_io.py)."""
import os
import subprocess

import numpy as np
import pytest

from clustering_b200 import io as dio
from clustering_b200 import density

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "io")
CLI = os.path.join(ROOT, "clustering_b200", "clustering")

HEADER = "# clustering v1.3.2 - density\n#\n# Created Thu Jan  1 00:00:00 1970\n# by following command:\n#\n# clustering density -f x \n"
COMMENTS = {"clustering_radius": 0.3, "lumping_radius": 0.2165063, "screening_from": 0.0, "cmin": 0.0, "limits": 0.0}


def sample_arrays():
    pops = np.array([1, 16390, 4999999, 7, 123456], np.uint32)
    fe = np.array([-0.0, 5.960464e-08, 1.2345678, 15.424948, 0.0, 1e-30, 3.4e38], np.float32)
    ni = np.array([3, 0, 999999, 12, 5], np.uint32)
    nd = np.array([0.0123457, 0.0, 123457.0, 1e-07, 2.5], np.float32)
    hi = np.array([1000001, 2, 7, 13, 4294967295], np.uint32)
    hd = np.array([np.finfo(np.float32).max, 1.5, 1e-07, 0.333333343, 1e10], np.float32)
    states = np.array([0, 1, 2, 117, 3, 3], np.uint32)
    return pops, fe, (ni, nd, hi, hd), states


def write_all(w, d):
    pops, fe, nb, states = sample_arrays()
    w.write_pops(os.path.join(d, "pops"), pops, HEADER, COMMENTS)
    w.write_fes(os.path.join(d, "fe"), fe, HEADER, COMMENTS)
    w.write_neighborhood(os.path.join(d, "nn"), *nb, HEADER, COMMENTS)
    w.write_states(os.path.join(d, "states"), states, HEADER, COMMENTS)


@pytest.mark.parametrize("name", ["pops", "fe", "nn", "states"])
def test_writers_match_reference_golden_bytes(tmp_path, name):
    write_all(dio, str(tmp_path))
    with open(os.path.join(GOLD, name), "rb") as f:
        want = f.read()
    with open(os.path.join(str(tmp_path), name), "rb") as f:
        assert f.read() == want


@pytest.mark.ref
def test_writers_match_reference_live(tmp_path, ref):
    a, b = tmp_path / "ours", tmp_path / "ref"
    a.mkdir(); b.mkdir()
    write_all(dio, str(a))
    write_all(ref, str(b))
    rng = np.random.default_rng(3)
    fe = np.concatenate([rng.gamma(2.0, 2.0, 5000), 10.0 ** rng.uniform(-12, 12, 5000)]).astype(np.float32)
    dio.write_fes(str(a / "fe_big"), fe, HEADER, {})
    ref.write_fes(str(b / "fe_big"), fe, HEADER, {})
    nd = (10.0 ** rng.uniform(-9, 9, 5000)).astype(np.float32)
    idx = rng.integers(0, 5_000_000, 5000).astype(np.uint32)
    dio.write_neighborhood(str(a / "nn_big"), idx, nd, idx[::-1].copy(), nd[::-1].copy(), HEADER, {})
    ref.write_neighborhood(str(b / "nn_big"), idx, nd, idx[::-1].copy(), nd[::-1].copy(), HEADER, {})
    for name in ("pops", "fe", "nn", "states", "fe_big", "nn_big"):
        assert (a / name).read_bytes() == (b / name).read_bytes(), name


def test_readers_round_trip_and_comments(tmp_path):
    pops, fe, nb, states = sample_arrays()
    d = str(tmp_path)
    write_all(dio, d)
    assert np.array_equal(dio.read_column(os.path.join(d, "pops"), int), pops)
    assert np.array_equal(dio.read_column(os.path.join(d, "states"), int), states)
    back = dio.read_column(os.path.join(d, "fe"), float)
    assert np.allclose(back, fe, rtol=1e-6, atol=0) and back.size == fe.size
    ni, nd, hi, hd = dio.read_neighborhood(os.path.join(d, "nn"))
    assert np.array_equal(ni, nb[0]) and np.array_equal(hi, nb[2])
    assert np.allclose(nd, nb[1], rtol=1e-5) and np.allclose(hd, nb[3], rtol=1e-5)
    assert abs(dio.read_comment(os.path.join(d, "fe"), "clustering_radius") - 0.3) < 1e-6
    assert abs(dio.read_comment(os.path.join(d, "nn"), "lumping_radius") - 0.21651) < 1e-6
    assert dio.read_comment(os.path.join(d, "nn"), "screening_from", 0.25) == 0.25       # zero values are not written


def test_read_coords_semantics(tmp_path):
    p = str(tmp_path / "coords")
    with open(p, "w") as f:
        f.write("\n1.5 2 -3e-1\n4 5 6\n\n7   8\t9\n 10 11 12")        # leading blank line, tabs, no trailing newline
    x = dio.read_coords(p)
    assert x.shape == (4, 3)
    assert np.array_equal(x, np.array([[1.5, 2, -0.3], [4, 5, 6], [7, 8, 9], [10, 11, 12]], np.float32))


@pytest.mark.ref
def test_read_coords_matches_reference(tmp_path, ref):
    rng = np.random.default_rng(5)
    x = rng.normal(size=(257, 5)).astype(np.float32)
    p = str(tmp_path / "coords")
    np.savetxt(p, x, fmt="%.6f")
    assert np.array_equal(dio.read_coords(p), ref.read_coords(p))


@pytest.mark.ref
def test_microstate_assignment_matches_reference(ref, oracle):
    from clustering_b200.synth import gaussian_mixture
    x = gaussian_mixture(3000, 3, k=5, seed=9)
    pops = ref.populations(x, np.array([0.3], np.float32))[0]
    fe = ref.free_energies(pops)
    ni, nd, hi, hd = ref.nearest_neighbors(x, fe)
    lab = ref.screening(fe, ni, nd, np.float32(1.0), x, None)
    want = ref.assign_low_density_frames(lab, hi, hd, fe)
    got = density.assign_low_density_frames(lab, hi, fe)
    assert np.array_equal(got, want.astype(np.uint32))
    assert np.array_equal(density.sorted_cluster_names(got), ref.sorted_cluster_names(want).astype(np.uint32))
    # many equal populations: the unstable sort's tie order must match
    st = np.repeat(np.arange(1, 41), 5).astype(np.uint32)
    np.random.default_rng(1).shuffle(st)
    assert np.array_equal(density.sorted_cluster_names(st), ref.sorted_cluster_names(st).astype(np.uint32))


def test_microstate_golden():
    g = np.load(os.path.join(ROOT, "tests", "golden", "io", "microstates.npz"))
    got = density.assign_low_density_frames(g["initial"], g["hd_idx"], g["fe"])
    assert np.array_equal(got, g["assigned"])
    assert np.array_equal(density.sorted_cluster_names(got), g["named"])
    assert np.array_equal(density.sorted_cluster_names(g["ties"]), g["ties_named"])


# ---- command line (argument handling happens before any device is touched) -------------------------------------
def run_cli(*args):
    return subprocess.run([CLI, *args], capture_output=True, text=True)


def test_cli_help_and_usage():
    r = run_cli("density", "-h")
    assert r.returncode == 0
    for flag in ("--file", "--radius", "--threshold-screening", "--output", "--input", "--radii", "--population", "--free-energy",
                 "--free-energy-input", "--nearest-neighbors", "--nearest-neighbors-input", "--nthreads", "--verbose"):
        assert flag in r.stdout
    for short in "f r T o i R p d D b B n v".split():
        assert f"-{short} [" in r.stdout
    r = run_cli()
    assert r.returncode != 0 and "usage:" in r.stderr
    r = run_cli("mpp", "-h")
    assert r.returncode != 0


def test_cli_argument_errors():
    r = run_cli("density", "-r", "0.3")
    assert r.returncode != 0 and "error parsing arguments" in r.stderr and "--file" in r.stderr
    r = run_cli("density", "-f", "x", "--bogus", "1")
    assert r.returncode != 0 and "unrecognised option" in r.stderr
    r = run_cli("density", "-f", "x", "-r", "abc")
    assert r.returncode != 0 and "invalid" in r.stderr
    r = run_cli("density", "-f", "x", "-r")
    assert r.returncode != 0 and "missing" in r.stderr
    r = run_cli("density", "-f", "x", "-f", "y")
    assert r.returncode != 0 and "more than once" in r.stderr
    r = run_cli("density", "-f", "x", "--rad", "1")            # ambiguous abbreviation (radius / radii)
    assert r.returncode != 0 and "ambiguous" in r.stderr


def test_label_side_channel_roundtrip(tmp_path):
    """the binary container of the per-threshold label files.  A reader gets the labels from the container
    only while the ASCII file it stands for is unchanged on disk; otherwise (no record, other size, other name pattern) it
    parses the ASCII file -- the format of record -- like read_clustered_trajectory (network_builder.cpp:411-437)."""
    from clustering_b200 import io
    rng = np.random.default_rng(5)
    base = str(tmp_path / "clust")
    labels = {}
    for k, t in enumerate((0.1, 0.2, 1.25)):
        lab = rng.integers(0, 50, size=4000).astype(np.uint32)
        fname = base + ".%0.2f" % t
        io.write_states(fname, lab, "# header\n", {"screening_from": 0.1})
        io.write_states_record(fname, lab, truncate=(k == 0))
        labels[fname] = lab
    assert os.path.exists(base + ".dcb200labels")
    for fname, lab in labels.items():
        got, from_bin = io.read_states(fname)
        assert from_bin and np.array_equal(got, lab)
        assert np.array_equal(io.read_column(fname, int), lab)            # the ASCII file says the same
    # the ASCII file was rewritten by somebody else: the stale record must not be used
    fname = base + ".0.20"
    other = np.arange(3000, dtype=np.uint32)
    io.write_states(fname, other, "# header\n", {})
    got, from_bin = io.read_states(fname)
    assert not from_bin and np.array_equal(got, other)
    # a re-run appends a fresh record: the last one wins
    io.write_states_record(fname, other)
    got, from_bin = io.read_states(fname)
    assert from_bin and np.array_equal(got, other)
    # files without a threshold suffix have no container
    plain = str(tmp_path / "states.dat")
    io.write_states(plain, other, "", {})
    got, from_bin = io.read_states(plain)
    assert not from_bin and np.array_equal(got, other)
    with pytest.raises(Exception):
        io.write_states_record(plain, other)
    # errors are statuses, not exits
    with pytest.raises(Exception):
        io.read_states(str(tmp_path / "missing.0.10"))
