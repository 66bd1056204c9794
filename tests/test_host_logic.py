"""CPU tests of the host side: the C ABI library loads and exports exactly what include/dcb200.h declares,
the host-only entry points agree with the oracle, compute entry points fail loudly without a GPU, and the
multi-rank sharding/gather logic works under gloo with world_size 2."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "dcb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dcb200_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from clustering_b200 import lib
    L = lib.load()
    declared = header_symbols()
    assert declared, "no declarations found in include/dcb200.h"
    assert sorted(lib.SYMBOLS) == declared
    for s in declared:
        assert hasattr(L, s), s
    out = subprocess.run(["nm", "-D", "--defined-only", lib.SO_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (dcb200_[a-z0-9_]+)", out))
    assert set(declared) <= exported
    assert L.dcb200_version() == 200


def test_no_oracle_in_product():
    # the product never links, loads or imports the checker
    out = subprocess.run(["ldd", os.path.join(ROOT, "clustering_b200", "libdcb200.so")], capture_output=True, text=True).stdout
    assert "dcoracle" not in out and "dcref" not in out
    for dirpath, _, files in os.walk(os.path.join(ROOT, "clustering_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "_oracle" not in src and "libdcoracle" not in src and "libdcref" not in src, f


def test_sorted_free_energies_and_sigma2_match_oracle(oracle):
    from clustering_b200 import density
    rng = np.random.default_rng(1)
    pops = rng.integers(1, 60, size=20000).astype(np.uint32)      # many ties, like real populations
    fe = oracle.free_energies(pops)
    assert np.array_equal(density.sorted_free_energies(fe), oracle.sorted_free_energies(fe).astype(np.uint32))
    d2 = rng.random(20000).astype(np.float32)
    assert density.compute_sigma2(d2) == oracle.sigma2(d2)


def test_compute_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from clustering_b200 import density, lib
    with pytest.raises(lib.Dcb200Error):
        density.calculate_populations(np.zeros((8, 3), np.float32), [0.5])
    with pytest.raises(lib.Dcb200Error):
        density.nearest_neighbors(np.zeros((8, 3), np.float32), np.zeros(8, np.float32))
    with pytest.raises(lib.Dcb200Error):
        density.screening(np.arange(8, dtype=np.float32), np.ones(8, np.float32), 10.0, np.random.rand(8, 3).astype(np.float32))


def test_shard_bounds_cover_everything():
    from clustering_b200.dist import shard_bounds, shard_positions, shard_size
    for n in (1, 1023, 1024, 1025, 100_000, 1_000_000, 5_000_001):
        for w in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(w):
                b, e = shard_bounds(n, w, r)
                assert b == min(prev, n) and b <= e <= n
                assert b % 1024 == 0 or b == n
                prev = e
            assert prev == n
            assert shard_size(n, w) * w >= n
    # block-cyclic shards: a partition of the positions, every shard within the common capacity, and the library's
    # own capacity formula agrees with the host-side one
    from clustering_b200 import lib
    L = lib.load()
    for n in (1, 1023, 1024, 1025, 5000, 100_000):
        for w in (1, 2, 3, 8):
            seen = []
            for r in range(w):
                p = shard_positions(n, w, r)
                assert len(p) <= shard_size(n, w)
                seen += p
            assert sorted(seen) == list(range(n))
            assert L.dcb200_shard_capacity(n, w) == shard_size(n, w)


_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from clustering_b200.dist import gather_shards, shard_bounds, shard_positions, shard_size, shards_to_position_order
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=int(sys.argv[4]))
w, r = dist.get_world_size(), dist.get_rank()
for n in (5000, 1024, 3000, 9001):
    cap = shard_size(n, w)
    want = torch.stack([torch.arange(n) * 3 + 1, -torch.arange(n)])
    # block-cyclic shards (n_cols <= 16): what a rank computes for ITS positions, padded to the common capacity
    pos = torch.tensor(shard_positions(n, w, r), dtype=torch.int64)
    local = torch.full((2, cap), -7, dtype=torch.int64)
    local[0, :len(pos)] = pos * 3 + 1
    local[1, :len(pos)] = -pos
    full = shards_to_position_order(gather_shards(local, w), n, cyclic=True)
    assert full.shape == (2, n) and torch.equal(full, want), (n, r)
    # contiguous shards (GEMM-form path)
    b, e = shard_bounds(n, w, r)
    pos = torch.arange(b, e, dtype=torch.int64)
    local = torch.full((2, cap), -7, dtype=torch.int64)
    local[0, :e - b] = pos * 3 + 1
    local[1, :e - b] = -pos
    out = torch.empty((w, 2, cap), dtype=torch.int64)
    full = shards_to_position_order(gather_shards(local, w, out=out), n, cyclic=False)
    assert torch.equal(full, want), (n, r)
dist.barrier()
dist.destroy_process_group()
print("rank", r, "ok")
"""


@pytest.mark.parametrize("world_size", [2, 3])
def test_gather_of_position_shards_gloo(tmp_path, world_size):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29600 + os.getpid() % 300 + world_size)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r), str(world_size)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(world_size)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o


def test_screening_row_cuts_balance_pairs():
    from clustering_b200.dist import screen_cuts
    for m_prev, m_new, w in ((0, 1000, 4), (900, 1000, 3), (0, 5, 8), (10, 10, 2), (123456, 5000000, 8)):
        cuts = screen_cuts(m_prev, m_new, w)
        assert cuts[0] == m_prev and cuts[-1] == m_new and len(cuts) == w + 1
        assert all(cuts[g] <= cuts[g + 1] for g in range(w))
        if m_new - m_prev > 100 * w:
            pairs = [sum(range(cuts[g], cuts[g + 1])) if m_new < 10000 else (cuts[g + 1] ** 2 - cuts[g] ** 2) / 2 for g in range(w)]
            assert max(pairs) < 1.2 * (sum(pairs) / w) + m_new


def test_bench_reference_arm_under_torchrun():
    """The driver launches both arms as `python -m torch.distributed.run ... bench.py --gpus N --steps K --warmup W [--impl reference]`:
    torchrun's own parser must let every bench.py option through (it rejected `--n` as an ambiguous abbreviation of its own
    options once), rank 0 alone prints the reference arm's line and the other rank exits 0 without work."""
    import json
    import subprocess
    port = str(29800 + os.getpid() % 150)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", port, os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "1", "--warmup", "0", "--impl", "reference",
           "--workload", "C1", "--frames", "3000"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "density_run_throughput" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["e2e"]["h2d_bytes_per_step"] == 0
