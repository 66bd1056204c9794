#!/usr/bin/env python
"""Benchmark of the `clustering density` hot path (BASELINE.json metric: density run wall-time & Gpair.dim/s, 1M frames,
1/2/4/8 B200 vs OpenMP host).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C3|C1|C2|C4|C5] [--impl ours|reference]

One step = one full density run over the workload: coordinate layout build, populations for ALL radii of the workload,
free energies, nearest neighbours (+ nearest neighbour with lower free energy) on the free energies of the workload's
`fe_radius_index`.  Default workload = C3, the configuration BASELINE.json's north_star targets: 1M frames x 10 dims,
the 20 radii 0.1 .. 2.0 (synthetic "HP35-like" Gaussian mixture, seed 3).  C4 adds the screening loop (all thresholds).

  value  whole-job Gpair.dim/s with the coordinates resident in HBM when the timed region starts
         (pair.dim = one (x_ik - x_jk)^2 term of the ordered N x N pair matrix; two pair scans per step)
  e2e    the same metric through the host-pointer C ABI (dcb200_density_run) with PINNED HOST buffers: the upload of
         the coordinates and the download of every result are inside the timed region
  N > 1  one process per GPU (torchrun), block-cyclic row shards, per-shard results assembled with ONE NCCL all-gather
         per stage on the device; fixed total work => "strong" scaling.  Before the line is printed the sharded result
         is compared with an unsharded run of rank 0 (checksum in the line).
  --impl reference   the reference's own OpenMP CPU implementation (oracle/_ref/libdcref.so, compiled from the
         unmodified reference sources; else the C port in oracle/) on a bounded sample of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "density_run_throughput"
UNIT = "Gpair.dim/s"
FLOP_PER_PAIR_DIM = 3.0          # one (x-y)^2 accumulate = FSUB + FFMA = 3 flop
FLOP_EXECUTED_PER_PAIR_DIM = 2.0  # what the kernels issue: one FFMA per pair.dim (|y|^2 - 2x.y form)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", dest="n", type=int, default=None,
                    help="override the frame count (tests only; invalid as a bench line).  Not `--n`: torchrun's own parser rejects it as ambiguous")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload(name, n=None):
    from clustering_b200.synth import CONFIGS, config_data
    c = dict(CONFIGS[name])
    if n is not None:
        c["n"] = n
    x = config_data(name, c["n"])
    return c, x


def ncu_traffic(workload_name, scan):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the workload's population ("pops") or neighbour ("nn")
    scan kernel, from profiles/ncu_traffic.json (written by scripts/ncu_summary.py --traffic from an `ncu --set full`
    capture of that workload at its full size); None if not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(workload_name, {}).get(scan)
    except Exception:
        return None


def workload_label(name, cfg):
    """the same `config.workload` string in both arms"""
    r = cfg["radii"]
    radii = f"{len(r)} radii {r[0]:g} .. {r[-1]:g}" if len(r) > 3 else "radius " + ", ".join(f"{v:g}" for v in r)
    tail = " + screening over all thresholds" if name == "C4" else ""
    return (f"{name}: clustering density, {cfg['n']} frames x {cfg['d']} dims, {radii}: populations + free energies + "
            f"nearest neighbours (-b) on the free energies of r = {r[cfg['fe_radius_index']]:g}{tail}")


def pair_dims_per_step(n, d):
    return 2.0 * float(n) * float(n) * float(d)      # populations scan + neighbour scan, ordered N x N pairs each


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        hi = [s for s in sm if s >= 0.5 * max(sm)] or sm          # samples under load
        return {"sm_mhz": float(np.median(hi)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own OpenMP implementation (oracle/_ref) or the C port (oracle/)
# ------------------------------------------------------------------------------------------------
def cpu_impl():
    from _oracle import Oracle, REF_SO
    if os.path.exists(REF_SO):
        from _oracle import Ref
        r = Ref()
        r.set_threads(os.cpu_count() or 1)              # torchrun pins OMP_NUM_THREADS=1; the baseline uses every core
        return "reference", r, r.max_threads()
    o = Oracle()
    return "port", o, os.cpu_count()


def cpu_step(kind, impl, x, radii, fe_index):
    """populations for all radii + free energies + neighbours on the CPU; returns seconds."""
    t0 = time.perf_counter()
    pops = impl.populations(x, np.asarray(radii, np.float32))
    fe = impl.free_energies(pops[fe_index])
    impl.nearest_neighbors(x, fe)
    return time.perf_counter() - t0


def cpu_sample_size(kind, impl, x_full, radii, fe_index, target_s):
    """frames of the workload whose CPU step takes about target_s (brute-force NN is exactly quadratic)."""
    n0 = min(6000, len(x_full))
    t = cpu_step(kind, impl, x_full[:n0], radii, fe_index)
    n = int(n0 * (target_s / max(t, 1e-3)) ** 0.5)
    return max(2000, min(len(x_full), n))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, x = workload(args.workload, args.n)
    d = cfg["d"]
    kind, impl, cores = cpu_impl()
    total = args.steps + args.warmup
    target = min(4.0, 150.0 / max(total, 1))
    ns = cpu_sample_size(kind, impl, x, cfg["radii"], cfg["fe_radius_index"], target)
    xs = np.ascontiguousarray(x[:ns])
    for _ in range(args.warmup):
        cpu_step(kind, impl, xs, cfg["radii"], cfg["fe_radius_index"])
    t = [cpu_step(kind, impl, xs, cfg["radii"], cfg["fe_radius_index"]) for _ in range(args.steps)]
    sec = float(np.mean(t))
    val = pair_dims_per_step(ns, d) / sec / 1e9
    sample = (f"first {ns} frames of {args.workload} ({cfg['n']}x{d}), populations ({len(cfg['radii'])} radii) + free energies + "
              "neighbours per step; the CPU populations are box-pruned, pair.dims counted as the full N x N matrix")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_label(args.workload, cfg), "sample_frames": ns},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def reference_cuda_arm(x, radii, fe_index, d, density):
    """The reference's own CUDA path (unmodified .cu files built for sm_100a, oracle/_ref/libdcrefcuda.so) on a bounded sample of
    the workload, next to this library on the same sample: the "kernel to beat" of BASELINE.json's north_star.  Reported, never
    used as a parity oracle (its semantics differ from the CPU path: d2 <= r2, duplicates excluded)."""
    try:
        from _oracle import RefCuda, REFCUDA_SO
        if not os.path.exists(REFCUDA_SO):
            return {"unavailable": "oracle/_ref/libdcrefcuda.so not built (make -C oracle refcuda needs /root/reference)"}
        if d > 12:
            return {"unavailable": f"the reference's population kernel needs 2*512*n_cols*4 B <= 48 KB of shared memory: n_cols <= 12, workload has {d}"}
        rc = RefCuda()
        ns = min(len(x), 200_000)
        xs = np.ascontiguousarray(x[:ns])
        rc.populations(xs[:20000], radii)                    # warm-up (context, module load)
        t0 = time.perf_counter(); pr = rc.populations(xs, radii); t_p = time.perf_counter() - t0
        fe = density.calculate_free_energies(pr[fe_index].astype(np.uint32))
        t0 = time.perf_counter(); rc.nearest_neighbors(xs, fe); t_n = time.perf_counter() - t0
        density.density_run(xs[:20000], radii, fe_index)
        t0 = time.perf_counter(); po = density.density_run(xs, radii, fe_index)["pops"]; t_o = time.perf_counter() - t0
        return {"sample": f"first {ns} frames of the workload, all {len(radii)} radii; host-pointer calls (H2D/D2H inside), wall clock",
                "populations_ms": t_p * 1e3, "nearest_neighbors_ms": t_n * 1e3,
                "value": pair_dims_per_step(ns, d) / (t_p + t_n) / 1e9, "unit": UNIT,
                "ours_same_sample": {"density_run_ms": t_o * 1e3, "value": pair_dims_per_step(ns, d) / t_o / 1e9},
                "population_counts_differing": int(np.count_nonzero(pr.astype(np.uint32) != po)),
                "note": "differences in counts come from the reference CUDA path's own rounding (sequential FMA) and '<=' test"}
    except Exception as ex:
        return {"unavailable": f"failed: {ex}"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def checksum(pops, nn):
    """order-sensitive 64-bit checksums of the populations and of the four neighbour arrays (device tensors)."""
    import torch
    n = pops.shape[-1]
    w = (torch.arange(n, device=pops.device, dtype=torch.int64) % 1000003) + 1
    cp = int((pops.to(torch.int64) * w).sum().item())
    cn = 0
    for t in nn:
        cn = (cn * 31 + int((t.view(torch.int32).to(torch.int64) * w).sum().item())) % (1 << 62)
    return cp, cn


def run_ours(args):
    import torch
    import torch.distributed as dist
    from clustering_b200 import density, lib
    from clustering_b200.session import Session

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")       # host-side barriers (an NCCL barrier spins ON the GPUs while it waits)
    cfg, x = workload(args.workload, args.n)
    n, d = x.shape
    radii = np.asarray(cfg["radii"], np.float32)
    r_fe = cfg["fe_radius_index"]
    screening_workload = args.workload == "C4"
    sess = Session(local)
    stream = sess.torch_stream()

    from clustering_b200.dist import DensityPass
    dpass = DensityPass(sess, n, radii)

    x_dev = torch.from_numpy(x).to(dev)                # resident in HBM before the timed region
    x_pin = torch.from_numpy(x).pin_memory()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > L2 (126 MB)
    out_host = [torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory(),
                torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory()]
    pops_host = torch.empty((radii.size, n), dtype=torch.int32).pin_memory()
    fe_host = torch.empty(n, dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()

    def step(coords, timed=False):
        """one density pass; coords: device tensor (resident) or pinned host tensor (per-rank e2e at N > 1)."""
        pops, fe, nn = dpass.run(coords if coords.is_cuda else coords.numpy(), r_fe, timed=timed)
        if not coords.is_cuda and rank == 0:             # e2e: the results land in rank 0's host buffers (ONE download, like
            with torch.cuda.stream(stream):              # the in-process path and the reference's single process)
                pops_host.copy_(pops, non_blocking=True)
                fe_host.copy_(fe, non_blocking=True)
                for h, t in zip(out_host, nn):
                    h.copy_(t, non_blocking=True)
        return pops, fe, nn

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stage_acc = {}

    def timed(coords, steps, stages=False):
        """per-step CUDA events on the session stream, L2 flushed between steps (outside the events)."""
        tot = 0.0
        for _ in range(steps):
            with torch.cuda.stream(stream):
                flush_buf.zero_()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step(coords, timed=stages)
            e1.record(stream)
            e1.synchronize()
            tot += e0.elapsed_time(e1)
            if stages:
                for k, v in dpass.stage_ms().items():
                    stage_acc[k] = stage_acc.get(k, 0.0) + v / steps
        t = torch.tensor([tot], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps                   # ms per step, max over ranks

    # ---- warm-up, FFMA peak, timed region --------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                                 # nvidia-smi needs ~1 s to deliver its first sample
    for _ in range(max(args.warmup, 3)):
        last = step(x_dev)
    barrier()
    tensor_path = sess.gemm_info()[0]                   # 17 <= n_cols <= 256: the scans run on the tensor cores (tcgen05, TF32)
    peak_tflops = sess.tf32_peak(300.0) if tensor_path else sess.ffma_peak(300.0)
    s0 = sess.stats(reset=True)
    barrier()
    ms = timed(x_dev, args.steps, stages=True)
    barrier()
    s1 = sess.stats()
    # keep the load on until the sampler has seen it (short timed regions), then stop it
    t_end = time.time() + 1.5
    while time.time() < t_end:
        last = step(x_dev)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    counts = torch.tensor([s1["launches"] - s0["launches"], s1["pairs_evaluated"], s1["pairs_scheduled"]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(counts)
    launches = counts[0]
    evaluated_frac = float(counts[1].item()) / max(1.0, float(counts[2].item()))
    pd = pair_dims_per_step(n, d)
    value = pd / (ms * 1e-3) / 1e9

    # per-rank stage times (CUDA events inside the pass), min / max over ranks: names the limiter of the scaling curve
    names = ["layout", "pops", "gather+fe+ranks", "nn", "gather+finish"]
    mine = torch.tensor([stage_acc.get(k, 0.0) for k in names], dtype=torch.float64, device=dev)
    lo, hi = mine.clone(), mine.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    stages = {k: {"min_ms": float(lo[i]), "max_ms": float(hi[i])} for i, k in enumerate(names)}

    # ---- parity of the sharded result: the gathered populations / neighbours == an unsharded scan of rank 0 -------------
    parity = None
    if world > 1:
        cs = checksum(last[0], last[2])
        if rank == 0:
            from clustering_b200.dist import _OnSessionStream
            with _OnSessionStream(sess):
                sess.set_coords(x_dev)
                full_p = sess.to_frame_order(sess.populations(radii, 0, n))
                fe_full = sess.free_energies(full_p[r_fe].contiguous())
                full_nn = sess.nearest_neighbors(fe_full)
            torch.cuda.synchronize()
            cs1 = checksum(full_p, full_nn)
            same = bool(torch.equal(full_p, last[0])) and all(bool(torch.equal(a.view(torch.int32), b.view(torch.int32)))
                                                              for a, b in zip(full_nn, last[2]))
            parity = {"sharded_equals_unsharded": same, "pops_checksum": cs[0], "nn_checksum": cs[1],
                      "unsharded_pops_checksum": cs1[0], "unsharded_nn_checksum": cs1[1]}
        flag = torch.tensor([1 if (parity is None or parity["sharded_equals_unsharded"]) else 0], device=dev)
        dist.broadcast(flag, 0)
        if int(flag.item()) != 1:
            raise SystemExit(f"bench.py: the sharded result on {world} GPUs differs from the unsharded scan: {parity}")
    else:
        cs = checksum(last[0], last[2])
        parity = {"pops_checksum": cs[0], "nn_checksum": cs[1]}

    # ---- the two pair scans alone on this rank's shard, for the roofline (the slower one is the dominant kernel) ---------
    def time_scan(fn, reps=3):
        fn()
        sess.stats(reset=True)
        kt = []
        for _ in range(reps):
            with torch.cuda.stream(stream):
                flush_buf.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            e1.synchronize()
            kt.append(e0.elapsed_time(e1))
        return float(np.mean(kt)), sess.stats()["pairs_evaluated"] / float(reps)

    barrier()
    with torch.cuda.stream(stream):
        sess.set_coords(x_dev)
        sess.nn_prepare(last[1])
    if world == 1:
        keys_loc = torch.empty((2, n), dtype=torch.int64, device=dev)
        pops_loc = torch.empty((radii.size, n), dtype=torch.int32, device=dev)
        nn_ms, nn_pairs = time_scan(lambda: sess.nn_scan(0, n, out=keys_loc))
        pops_ms, pops_pairs = time_scan(lambda: sess.populations(radii, 0, n, out=pops_loc))
    else:
        nn_ms, nn_pairs = time_scan(lambda: sess.nn_scan_shard(rank, world, out=dpass.keys_loc))
        pops_ms, pops_pairs = time_scan(lambda: sess.populations_shard(radii, rank, world, out=dpass.pops_loc))
    scans = {"nn_kernel (neighbour scan)": (nn_ms, nn_pairs), "population kernel (multi-radius scan)": (pops_ms, pops_pairs)}
    kname = max(scans, key=lambda k: scans[k][0])
    kms, k_pairs = scans[kname]
    barrier()

    # ---- end to end: host buffers through the C ABI ----------------------------------------------
    screening_info = None
    if world == 1:
        lib.set_gpus(1)                                  # the in-process path must use ONE GPU at N = 1
        outs = dict(pops=pops_host.numpy().view(np.uint32), fe=fe_host.numpy(),
                    nn=(out_host[0].numpy().view(np.uint32), out_host[1].numpy(), out_host[2].numpy().view(np.uint32), out_host[3].numpy()))
        xp = x_pin.numpy()
        lab_host = torch.empty(n, dtype=torch.int32).pin_memory().numpy().view(np.uint32) if screening_workload else None

        def e2e_step():
            r = density.density_run(xp, radii, r_fe, neighbors=True, out=outs)
            if screening_workload:
                with density.ScreeningRun(r["fe"], r["nn"][1], xp) as run:
                    t, t_to, k, lab = np.float32(0.1), float(r["fe"].max()), 0, None
                    while (t < np.float32(t_to - 0.01 + 0.1)) and not (np.float32(t_to + 0.01 + 0.1) < t):
                        lab = run.next(t, out=lab_host)
                        t = np.float32(t + np.float32(0.1))
                        k += 1
                return k, int(lab.max()), int((lab.astype(np.int64) * (np.arange(n) % 1000003 + 1)).sum())
            return None
        e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            info = e2e_step()
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        if info:
            screening_info = {"thresholds": info[0], "clusters_at_last_threshold": info[1], "labels_checksum": info[2]}
        h2d = x.nbytes + (x.nbytes if screening_workload else 0)
        d2h = radii.size * 4 * n + 4 * n + 16 * n + (screening_info["thresholds"] * 4 * n if screening_info else 0)
        e2e_api = ("host-pointer C ABI, pinned host buffers: dcb200_density_run (upload, populations of all radii, free energies, "
                   "neighbours, downloads)" + (" + dcb200_screening_begin/next/end over all thresholds" if screening_workload else "") +
                   "; wall clock around the calls")
        cxx = None
    else:
        step(x_pin)
        e2e_ms = timed(x_pin, args.steps)
        h2d = x.nbytes
        d2h = radii.size * 4 * n + 4 * n + 16 * n
        e2e_api = ("session C ABI per rank with pinned host buffers: H2D of the coordinates on every rank (h2d_bytes_per_step is per rank), "
                   "block-cyclic sharded scans, NCCL all-gathers, D2H of all results on rank 0 (CUDA events, max over ranks)")
        # the C++ in-process path (what the `clustering` binary runs): one process, N GPUs, NCCL inside libdcb200.so.
        # The other ranks wait on the HOST meanwhile, so that their GPUs are free for rank 0's worker threads.
        cxx = None
        barrier()
        dist.barrier(group=cpu_group)
        if rank == 0:
            try:
                lib.set_gpus(world)
                outs = dict(pops=pops_host.numpy().view(np.uint32), fe=fe_host.numpy(),
                            nn=(out_host[0].numpy().view(np.uint32), out_host[1].numpy(), out_host[2].numpy().view(np.uint32), out_host[3].numpy()))
                xp = x_pin.numpy()
                density.density_run(xp, radii, r_fe, neighbors=True, out=outs)          # creates contexts + communicators
                density.density_run(xp, radii, r_fe, neighbors=True, out=outs)
                t0 = time.perf_counter()
                reps = max(3, args.steps // 2)
                for _ in range(reps):
                    density.density_run(xp, radii, r_fe, neighbors=True, out=outs)
                cms = (time.perf_counter() - t0) * 1e3 / reps
                same = bool(np.array_equal(outs["pops"], last[0].cpu().numpy().view(np.uint32)) and
                            np.array_equal(outs["nn"][0], last[2][0].cpu().numpy().view(np.uint32)) and
                            np.array_equal(outs["nn"][3].view(np.uint32), last[2][3].cpu().numpy().view(np.uint32)))
                cxx = {"ms_per_step": cms, "value": pd / (cms * 1e-3) / 1e9, "unit": UNIT, "equals_torchrun_result": same,
                       "api": f"one process, {world} GPUs: dcb200_density_run (one H2D + NCCL broadcast, block-cyclic shards, ncclAllGather, one D2H)"}
            except Exception as ex:
                cxx = {"failed": str(ex)}
        dist.barrier(group=cpu_group)
        barrier()
    e2e_value = pd / (e2e_ms * 1e-3) / 1e9

    # per-GPU roofline numbers of every rank (max kernel time decides the step)
    roof = torch.tensor([kms, k_pairs, pops_ms, pops_pairs, nn_ms, nn_pairs], dtype=torch.float64, device=dev)
    roofs = [torch.zeros_like(roof) for _ in range(world)] if world > 1 else [roof]
    if world > 1:
        dist.all_gather(roofs, roof)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "tf32 filter + f32 exact recheck" if tensor_path else "f32", "data": "synthetic",
            "config": {"workload": workload_label(args.workload, cfg),
                       "step": "layout build + populations (all radii) + free energies + nearest neighbours (+ lower-free-energy neighbour)",
                       "pair_dims_per_step": pd,
                       "parallelism": f"rows dealt block-cyclically (1024-row blocks) to {world} GPU(s), coords replicated, one NCCL all-gather "
                                      "of the populations and one of the neighbour keys",
                       "l2": "flushed between timed steps (256 MiB write)", "seed": cfg["seed"],
                       "pairs_evaluated_frac": evaluated_frac},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "api": e2e_api},
            "gpu_launches": int(launches.item()),
            "stages_ms": stages,
            "parity": parity,
        }
        if screening_info:
            line["screening"] = screening_info
        if cxx is not None:
            line["cxx_inprocess"] = cxx
        pairs = float(n) * float(n)
        rows_frac = 1.0 / world
        achieved = FLOP_EXECUTED_PER_PAIR_DIM * k_pairs * d / (kms * 1e-3) / 1e12
        line["roofline"] = {
            "bound": "tensor" if tensor_path else "fp32", "kernel": kname + (f" (rank 0's shard of {world})" if world > 1 else ""),
            "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
            "kernel_ms": kms,
            "pairs_evaluated_per_launch": k_pairs, "pairs_evaluated_frac": k_pairs / (pairs * rows_frac),
            "evaluated_gpair_dim_per_s": k_pairs * d / (kms * 1e-3) / 1e9,
            "effective_gpair_dim_per_s": pairs * rows_frac * d / (kms * 1e-3) / 1e9,
            "algorithmic_3flop_tflops_on_evaluated_pairs": FLOP_PER_PAIR_DIM * k_pairs * d / (kms * 1e-3) / 1e12,
            "peak_source": ("tcgen05.mma kind::tf32 128x128x8 back to back on resident operands, measured in this run "
                            "(dcb200_ctx_tf32_peak); MEASURED_PEAKS.json holds bf16 only (TF32 runs at half the bf16 rate)") if tensor_path else
                           "FFMA-only microbenchmark in this run (dcb200_ctx_ffma_peak); MEASURED_PEAKS.json has no FP32 entry",
            "traffic": ncu_traffic(args.workload, "nn" if kname.startswith("nn") else "pops") if args.n is None else None,
            "other_scans": {k: {"kernel_ms": v[0], "pairs_evaluated_frac": v[1] / (pairs * rows_frac),
                                "achieved_tflops": FLOP_EXECUTED_PER_PAIR_DIM * v[1] * d / (v[0] * 1e-3) / 1e12}
                            for k, v in scans.items() if k != kname},
            "note": ("GEMM-form path (n_cols >= 17): the roofline is the tensor pipe (TF32). achieved = 2 flop per pair.dim x the pairs of "
                     "the 128 x 128 tile pairs the kernel really multiplied; pruned tile pairs are not counted; the headline value "
                     "counts the full N x N matrix") if tensor_path else
                    "compute-bound path (SURVEY.md 8d): the roofline is the FP32 FFMA pipe, not HBM. achieved = 2 flop (one FFMA) per "
                    "pair.dim x the pairs the kernel really evaluated = (warp, tile) scans x 128 rows x 128 columns; tiles out of reach "
                    "of a row group are pruned and not counted; the headline value counts the full N x N matrix. "
                    "traffic = dram bytes per launch from the committed ncu capture (profiles/), null if none",
        }
        if world > 1:
            line["roofline"]["per_gpu"] = [
                {"rank": g, "pops_ms": float(r[2]), "pops_tflops": FLOP_EXECUTED_PER_PAIR_DIM * float(r[3]) * d / (float(r[2]) * 1e-3) / 1e12,
                 "nn_ms": float(r[4]), "nn_tflops": FLOP_EXECUTED_PER_PAIR_DIM * float(r[5]) * d / (float(r[4]) * 1e-3) / 1e12,
                 "frac_of_peak_dominant": FLOP_EXECUTED_PER_PAIR_DIM * float(r[1]) * d / (float(r[0]) * 1e-3) / 1e12 / peak_tflops}
                for g, r in enumerate(roofs)]
        if not args.no_cpu_baseline:
            try:
                kind, impl, cores = cpu_impl()
                ns = cpu_sample_size(kind, impl, x, radii, r_fe, 12.0)
                sec = cpu_step(kind, impl, np.ascontiguousarray(x[:ns]), radii, r_fe)
                line["cpu_baseline"] = {
                    "value": pair_dims_per_step(ns, d) / sec / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
                    "sample": f"first {ns} frames of the workload, one populations ({radii.size} radii) + free energies + neighbours step in {sec:.1f} s "
                              "(brute-force NN is quadratic; the CPU populations are box-pruned, counted as full N x N)"}
            except Exception as ex:                       # the baseline must never take the bench line down
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
        if not args.no_cpu_baseline and world == 1:
            line["reference_cuda"] = reference_cuda_arm(x, radii, r_fe, d, density)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
