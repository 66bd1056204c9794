"""Multi-GPU paths (pytest -m gpu; skipped on a box with one GPU):

  * the in-process path of the host-pointer C ABI (what the `clustering` binary runs): one process, worker thread per GPU,
    one upload + NCCL broadcast, block-cyclic shards, ncclAllGather inside libdcb200.so -- against the oracle and against
    the same calls restricted to one GPU, with and without NCCL (DCB200_NO_NCCL=1: host assembly);
  * the one-process-per-GPU driver (torchrun + clustering_b200.dist.DensityPass): bench.py's own sharded-vs-unsharded
    assertion on a reduced workload.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from clustering_b200 import density, lib
from clustering_b200.synth import gaussian_mixture

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        return lib.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_gpus() < 2, reason="needs at least two GPUs")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@needs2
@pytest.mark.parametrize("d,radii", [(5, [0.3]), (10, [0.2 * i for i in range(1, 11)]), (20, [1.5, 2.0])])
def test_inprocess_multi_gpu_equals_one_gpu_and_oracle(oracle, d, radii):
    n = 23_000 if d <= 16 else 6_000
    x = gaussian_mixture(n, d, seed=9000 + d)
    radii = np.asarray(radii, np.float32)
    want = oracle.populations(x, radii)
    fe = oracle.free_energies(want[0])
    nn_o = oracle.nearest_neighbors(x, fe)
    results = []
    try:
        for g in (1, 2, min(_gpus(), 8)):
            lib.set_gpus(g)
            r = density.density_run(x, radii, 0, neighbors=True, all_free_energies=True)
            assert np.array_equal(r["pops"], want), g
            assert np.array_equal(bits(r["fe"]), bits(fe)), g
            for a, b in zip(nn_o, r["nn"]):
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), g
            p = density.calculate_populations(x, radii)                 # the separate entry points on the same gang
            assert np.array_equal(p, want), g
            for a, b in zip(nn_o, density.nearest_neighbors(x, fe)):
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), g
            results.append(r)
    finally:
        lib.set_gpus(0)


@needs2
def test_inprocess_multi_gpu_screening_run(oracle):
    x = gaussian_mixture(9_000, 3, k=6, seed=9100)
    fe = oracle.free_energies(oracle.populations(x, [0.3])[0])
    _, nd, _, _ = oracle.nearest_neighbors(x, fe)
    try:
        lib.set_gpus(min(_gpus(), 4))
        prev = None
        with density.ScreeningRun(fe, nd, x) as run:
            t = np.float32(0.2)
            while t < fe.max() + 0.3:
                prev = oracle.screening(fe, nd, t, x, prev)
                assert np.array_equal(run.next(t), prev.astype(np.uint32)), float(t)
                t = np.float32(t + np.float32(0.5))
    finally:
        lib.set_gpus(0)


@needs2
def test_host_assembly_without_nccl(oracle, tmp_path):
    """DCB200_NO_NCCL=1: the same entry points with the shards assembled through the host (fresh process: the gang is cached)."""
    code = """
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from clustering_b200 import density, lib
from clustering_b200.synth import gaussian_mixture
from _oracle import Oracle
o = Oracle()
x = gaussian_mixture(15000, 4, seed=9200)
radii = np.array([0.25, 0.4], np.float32)
lib.set_gpus(2)
r = density.density_run(x, radii, 1)
want = o.populations(x, radii)
assert np.array_equal(r["pops"], want)
fe = o.free_energies(want[1])
for a, b in zip(o.nearest_neighbors(x, fe), r["nn"]):
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
print("host assembly ok")
""" % (ROOT, os.path.join(ROOT, "tests"))
    env = dict(os.environ, DCB200_NO_NCCL="1")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and "host assembly ok" in out.stdout, out.stderr[-2000:]


@needs2
def test_torchrun_sharded_density_pass_equals_unsharded():
    """bench.py under torchrun on two GPUs with a reduced workload: it asserts itself that the gathered block-cyclic result
    equals rank 0's unsharded scan, and reports the in-process C++ path next to it."""
    port = str(29700 + os.getpid() % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", port, os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "2", "--warmup", "1", "--frames", "150000",
           "--no-cpu-baseline"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["parity"]["sharded_equals_unsharded"] is True
    assert line["parity"]["pops_checksum"] == line["parity"]["unsharded_pops_checksum"]
    assert line["cxx_inprocess"].get("equals_torchrun_result") is True, line["cxx_inprocess"]
