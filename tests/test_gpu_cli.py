"""End-to-end runs of the `clustering density` command-line driver on the GPU box, checked against the CPU oracle:
same flags, file names and file contents as the reference's driver (density_clustering.cpp:559-825)."""
import os
import subprocess

import numpy as np
import pytest

from clustering_b200 import io as dio
from clustering_b200.synth import gaussian_mixture

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "clustering_b200", "clustering")


def data_lines(path):
    with open(path) as f:
        return [ln.rstrip("\n") for ln in f if not ln.startswith("#")]


def comment_lines(path):
    with open(path) as f:
        return [ln.rstrip("\n") for ln in f if ln.startswith("#@")]


def run(*args, cwd):
    r = subprocess.run([CLI, "density", *args], cwd=cwd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return r


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    x = gaussian_mixture(2500, 3, k=5, seed=31)
    x[40] = x[4]
    np.savetxt(d / "coords", x, fmt="%.6f")
    return d, dio.read_coords(str(d / "coords"))          # what every reader of that file sees


def thresholds(t_from, t_step, t_to):
    out = []
    t_from, t_step, t_to = np.float32(t_from), np.float32(t_step), np.float32(t_to)
    lo = np.float32(np.float32(t_to - np.float32(t_step / np.float32(10.0))) + t_step)
    hi = np.float32(np.float32(t_to + np.float32(t_step / np.float32(10.0))) + t_step)
    t = t_from
    while t < lo and not (hi < t):
        out.append(t)
        t = np.float32(t + t_step)
    return out


def test_single_radius_pops_fe_nn(workdir, oracle):
    d, x = workdir
    run("-f", "coords", "-r", "0.3", "-p", "pop", "-d", "fe", "-b", "nn", cwd=d)
    pops = oracle.populations(x, np.array([0.3], np.float32))[0]
    fe = oracle.free_energies(pops)
    ni, nd, hi, hd = oracle.nearest_neighbors(x, fe)
    assert data_lines(d / "pop") == [str(int(p)) for p in pops]
    assert data_lines(d / "fe") == ["%e" % float(v) for v in fe]
    assert data_lines(d / "nn") == ["%d %g %d %g" % (a, float(b), c, float(e)) for a, b, c, e in zip(ni, nd, hi, hd)]
    lump = np.float32(np.sqrt(4 * oracle.sigma2(nd)))
    assert comment_lines(d / "nn") == ["#@   clustering_radius = 0.30000", "#@   lumping_radius = %.5f" % float(lump)]
    assert comment_lines(d / "pop") == ["#@   clustering_radius = 0.30000"]
    with open(d / "pop") as f:
        head = f.read().split("\n")
    assert head[0] == "# clustering v1.3.2 - density" and "# point density of each frame" in head


def test_multi_radius_files(workdir, oracle):
    d, x = workdir
    run("-f", "coords", "-R", "0.2", "0.1", "0.4", "0.2", "-p", "mpop", "-d", "mfe", "-v", cwd=d)
    radii = np.array([0.2, 0.1, 0.4, 0.2], np.float32)
    pops = oracle.populations(x, radii)
    for r, name in ((1, "0.100000"), (0, "0.200000"), (2, "0.400000")):
        assert data_lines(d / f"mpop_{name}") == [str(int(p)) for p in pops[r]]
        assert data_lines(d / f"mfe_{name}") == ["%e" % float(v) for v in oracle.free_energies(pops[r])]
    assert sorted(p for p in os.listdir(d) if p.startswith("mpop_")) == ["mpop_0.100000", "mpop_0.200000", "mpop_0.400000"]


def test_screening_in_one_run_and_microstates(workdir, oracle):
    d, x = workdir
    run("-f", "coords", "-r", "0.3", "-d", "fe2", "-b", "nn2", "-T", "0.1", "0.4", "2.0", "-o", "clust", cwd=d)
    pops = oracle.populations(x, np.array([0.3], np.float32))[0]
    fe = oracle.free_energies(pops)
    ni, nd, hi, hd = oracle.nearest_neighbors(x, fe)
    prev = None
    ts = thresholds(0.1, 0.4, 2.0)
    assert len(ts) >= 5
    for t in ts:
        lab = oracle.screening(fe, nd, t, x, prev)
        assert data_lines(d / ("clust.%0.2f" % float(t))) == [str(int(v)) for v in lab], float(t)
        prev = lab
    assert "#@   screening_step = 0.40000" in comment_lines(d / ("clust.%0.2f" % float(ts[0])))
    # second invocation, like the reference workflow: free energies and neighbours re-used from the files
    last = "clust.%0.2f" % float(ts[2])
    run("-f", "coords", "-r", "0.3", "-D", "fe2", "-B", "nn2", "-i", last, "-o", "micro", cwd=d)
    fe_file = dio.read_column(str(d / "fe2"), float)
    ni_f, nd_f, hi_f, hd_f = dio.read_neighborhood(str(d / "nn2"))
    init = dio.read_column(str(d / last), int)
    order = oracle.sorted_free_energies(fe_file)
    want = oracle.assign_low_density_frames(init, hi_f, fe_file)
    got = np.array([int(v) for v in data_lines(d / "micro")], np.uint32)
    assert got.size == x.shape[0] and got.min() >= 1
    # every frame ends in the state its chain of lower-free-energy neighbours leads to
    st = init.astype(np.int64).copy()
    for k in order:
        if st[k] == 0 and hi_f[k] < st.size:
            st[k] = st[hi_f[k]]
    assert np.array_equal(st, want.astype(np.int64))
    names, counts = np.unique(st, return_counts=True)
    assert np.array_equal(np.sort(np.unique(got, return_counts=True)[1]), np.sort(counts))       # same partition sizes
    # the partition itself is the same (names differ: renamed by decreasing population)
    pairs = set(zip(st.tolist(), got.tolist()))
    assert len(pairs) == names.size
    # decreasing population: state 1 is the largest
    assert np.all(np.diff(np.unique(got, return_counts=True)[1]) <= 0)


def test_screening_defaults_and_lumping_radius(workdir, oracle):
    d, x = workdir
    # no -r: the clustering radius is the lumping radius of a first pass with radius 1; -T -1: defaults 0.1 / 0.1 / max FE
    run("-f", "coords", "-p", "pop3", "-T", "-1", "-o", "c3", cwd=d)
    p1 = oracle.populations(x, np.array([1.0], np.float32))[0]
    nd1 = oracle.nearest_neighbors(x, oracle.free_energies(p1))[1]
    r_lump = np.float32(np.sqrt(4 * oracle.sigma2(nd1)))
    pops = oracle.populations(x, np.array([r_lump], np.float32))[0]
    assert data_lines(d / "pop3") == [str(int(p)) for p in pops]
    fe = oracle.free_energies(pops)
    nd = oracle.nearest_neighbors(x, fe)[1]
    ts = thresholds(0.1, 0.1, fe.max())
    files = sorted(p for p in os.listdir(d) if p.startswith("c3."))
    assert files == sorted("c3.%0.2f" % float(t) for t in ts)
    prev = None
    for t in ts:
        lab = oracle.screening(fe, nd, t, x, prev)
        prev = lab
    assert data_lines(d / ("c3.%0.2f" % float(ts[-1]))) == [str(int(v)) for v in prev]


def test_cli_semantic_errors(workdir):
    d, _ = workdir
    for args, msg in ((("-f", "coords", "-R", "0.1", "0.2", "-o", "x", "-p", "p"), "several radii"),
                      (("-f", "coords", "-R", "0.1", "0.2", "-b", "n", "-p", "p"), "several radii"),
                      (("-f", "coords", "-r", "0.3", "-o", "x"), "one of -T/-i is needed"),
                      (("-f", "coords", "-r", "0.3", "-T", "0.123", "-o", "x"), "two digits"),
                      (("-f", "coords", "-i", "a", "-d", "b", "-o", "x"), "-D/-B should be used"),
                      (("-f", "nonexistent", "-r", "0.3", "-p", "p"), "cannot open file")):
        r = subprocess.run([CLI, "density", *args], cwd=d, capture_output=True, text=True)
        assert r.returncode != 0 and msg in r.stderr, (args, r.stderr)
