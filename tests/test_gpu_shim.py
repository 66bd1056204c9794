"""The C++ boundary (clustering_b200/include/dcb200/density_cuda.hpp: the reference's own Clustering::Density::CUDA
signatures over libdcb200.so), driven through the compiled test driver and compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

from clustering_b200.synth import gaussian_mixture

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "clustering_b200", "shim_check")


def test_cpp_shim_matches_oracle(oracle, tmp_path):
    assert os.path.exists(DRIVER), "build clustering_b200/shim_check with make -C clustering_b200/csrc"
    n, d = 2500, 3
    x = gaussian_mixture(n, d, k=5, seed=321)
    x[40] = x[4]
    radii = [0.25, 0.4]
    x.tofile(tmp_path / "c.f32")
    prefix = str(tmp_path / "out")
    out = subprocess.run([DRIVER, str(tmp_path / "c.f32"), str(n), str(d), prefix, "0.4"] + [repr(r) for r in radii],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    pops = np.fromfile(prefix + ".pops.u32", np.uint32).reshape(len(radii), n)
    assert np.array_equal(pops, oracle.populations(x, radii))
    fe = np.fromfile(prefix + ".fe.f32", np.float32)
    assert np.array_equal(fe.view(np.uint32), oracle.free_energies(pops[0]).view(np.uint32))
    idx = np.fromfile(prefix + ".nn.u32", np.uint32).reshape(2, n)
    d2 = np.fromfile(prefix + ".nnd.f32", np.float32).reshape(2, n)
    ni, nd, hi, hd = oracle.nearest_neighbors(x, fe)
    assert np.array_equal(idx[0], ni) and np.array_equal(idx[1], hi)
    assert np.array_equal(d2[0].view(np.uint32), nd.view(np.uint32)) and np.array_equal(d2[1].view(np.uint32), hd.view(np.uint32))
    thr = np.fromfile(prefix + ".thr.f32", np.float32)
    lab = np.fromfile(prefix + ".lab.u32", np.uint32).reshape(len(thr), n)
    prev = None
    for k, t in enumerate(thr):
        prev = oracle.screening(fe, nd, t, x, prev)
        assert np.array_equal(lab[k], prev.astype(np.uint32)), k
    # arbitrary initial clusters (not a prefix of the free-energy order, names with gaps)
    init = np.fromfile(prefix + ".arbinit.u32", np.uint32)
    arb = np.fromfile(prefix + ".arb.u32", np.uint32)
    assert np.count_nonzero(init) > 0 and np.count_nonzero(init == 0) > 0
    want = oracle.screening(fe, nd, thr[-1], x, init.astype(np.uint64))
    assert np.array_equal(arb, want.astype(np.uint32))


def test_cpp_shim_exits_like_the_reference_on_bad_input(tmp_path):
    # error convention of the reference's CUDA path: message on stderr + exit(EXIT_FAILURE), no exceptions
    out = subprocess.run([DRIVER, str(tmp_path / "missing.f32"), "10", "2", str(tmp_path / "o"), "0.5", "0.3"],
                         capture_output=True, text=True)
    assert out.returncode != 0 and "cannot read" in out.stderr
