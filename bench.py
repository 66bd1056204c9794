#!/usr/bin/env python
"""Benchmark of the `clustering density` hot path (BASELINE.json metric: density run wall-time & Gpair.dim/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2|C1|C3|C4|C5] [--impl ours|reference]

One step = one full density pass over the workload: coordinate layout build, populations, free energies,
nearest neighbours (+ nearest neighbour with lower free energy).  Default workload = BASELINE.json
configs[1] (C2: 1M frames x 5 dims, radius 0.3, seed 2; synthetic "HP35-like" Gaussian mixture).

  value  whole-job Gpair.dim/s with the coordinates resident in HBM when the timed region starts
         (pair.dim = one (x_ik - x_jk)^2 term of the ordered N x N pair matrix; two pair scans per step)
  e2e    same metric through the C ABI with HOST buffers (H2D/D2H inside the timed region)
  N > 1  one process per GPU (torchrun), rows sharded, per-shard results assembled with NCCL all-gather
         on the device; fixed total work => "strong" scaling
  --impl reference   the reference's own OpenMP CPU implementation (oracle/_ref/libdcref.so, compiled from
         the unmodified reference sources; else the C port in oracle/) on a bounded sample of the workload.
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=None, help="override the frame count (debugging only; invalid as a bench line)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload(name, n=None):
    from clustering_b200.synth import CONFIGS, config_data
    c = dict(CONFIGS[name])
    if n is not None:
        c["n"] = n
    x = config_data(name, c["n"])
    return c, x


def ncu_traffic(kernel_name):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the kernel, from profiles/ncu_traffic.json
    (written by scripts/ncu_summary.py from an `ncu --set full` capture of this workload); None if not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        for key, val in t.items():
            if key in kernel_name:
                return val
    except Exception:
        pass
    return None


def workload_label(name, cfg):
    """the same `config.workload` string in both arms"""
    radii = ", ".join(f"{r:g}" for r in cfg["radii"])
    return (f"{name}: clustering density, {cfg['n']} frames x {cfg['d']} dims, radius {radii}: "
            "populations + free energies + nearest neighbours (-b)")


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


def cpu_step(kind, impl, x, radius):
    """pops + FE + NN on the CPU; returns seconds."""
    t0 = time.perf_counter()
    pops = impl.populations(x, np.array([radius], np.float32))
    fe = impl.free_energies(pops[0])
    impl.nearest_neighbors(x, fe)
    return time.perf_counter() - t0


def cpu_sample_size(kind, impl, x_full, radius, target_s):
    """frames of the workload whose CPU step takes about target_s (brute-force NN is exactly quadratic)."""
    n0 = min(8000, len(x_full))
    t = cpu_step(kind, impl, x_full[:n0], radius)
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
    ns = cpu_sample_size(kind, impl, x, cfg["radii"][0], target)
    xs = np.ascontiguousarray(x[:ns])
    for _ in range(args.warmup):
        cpu_step(kind, impl, xs, cfg["radii"][0])
    t = [cpu_step(kind, impl, xs, cfg["radii"][0]) for _ in range(args.steps)]
    sec = float(np.mean(t))
    val = pair_dims_per_step(ns, d) / sec / 1e9
    sample = f"first {ns} frames of {args.workload} ({cfg['n']}x{d}), pops(r={cfg['radii'][0]})+FE+NN per step; pops is box-pruned on the CPU, pair.dims counted as the full N x N matrix"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_label(args.workload, cfg), "sample_frames": ns},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def reference_cuda_arm(x, radii, d, density):
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
        ns = min(len(x), 250_000)
        xs = np.ascontiguousarray(x[:ns])
        r1 = np.ascontiguousarray(radii[:1])
        rc.populations(xs[:20000], r1)                       # warm-up (context, module load)
        t0 = time.perf_counter(); pr = rc.populations(xs, r1); t_p = time.perf_counter() - t0
        fe = density.calculate_free_energies(pr[0].astype(np.uint32))
        t0 = time.perf_counter(); rc.nearest_neighbors(xs, fe); t_n = time.perf_counter() - t0
        density.calculate_populations(xs[:20000], r1)
        t0 = time.perf_counter(); po = density.calculate_populations(xs, r1); t_po = time.perf_counter() - t0
        t0 = time.perf_counter(); density.nearest_neighbors(xs, fe); t_no = time.perf_counter() - t0
        return {"sample": f"first {ns} frames of the workload, radius {float(r1[0]):g}; host-pointer calls (H2D/D2H inside), wall clock",
                "populations_ms": t_p * 1e3, "nearest_neighbors_ms": t_n * 1e3,
                "value": pair_dims_per_step(ns, d) / (t_p + t_n) / 1e9, "unit": UNIT,
                "ours_same_sample": {"populations_ms": t_po * 1e3, "nearest_neighbors_ms": t_no * 1e3,
                                     "value": pair_dims_per_step(ns, d) / (t_po + t_no) / 1e9},
                "population_counts_differing": int(np.count_nonzero(pr[0] != po[0])),
                "note": "differences in counts come from the reference CUDA path's own rounding (sequential FMA) and '<=' test"}
    except Exception as ex:
        return {"unavailable": f"failed: {ex}"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg, x = workload(args.workload, args.n)
    n, d = x.shape
    radii = np.asarray(cfg["radii"], np.float32)
    r_fe = 0                                            # free energies / neighbours from the first radius
    sess = Session(local)
    stream = sess.torch_stream()

    from clustering_b200.dist import DensityPass, shard_bounds
    b, e = shard_bounds(n, world, rank)
    dpass = DensityPass(sess, n, radii)

    x_dev = torch.from_numpy(x).to(dev)                # resident in HBM before the timed region
    x_pin = torch.from_numpy(x).pin_memory()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > L2 (126 MB)
    keys_loc = torch.zeros((2, n), dtype=torch.int64, device=dev) if world == 1 else None
    out_host = [torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory(),
                torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory()]
    pops_host = torch.empty((radii.size, n), dtype=torch.int32).pin_memory()
    fe_host = torch.empty(n, dtype=torch.float32).pin_memory()

    def step(coords):
        """one density pass; coords: device tensor (resident) or pinned host tensor (e2e)."""
        with torch.cuda.stream(stream):
            pops, fe, nn = dpass.run(coords if coords.is_cuda else coords.numpy(), r_fe)
            if not coords.is_cuda:                       # e2e: results back in host memory
                pops_host.copy_(pops, non_blocking=True)
                fe_host.copy_(fe, non_blocking=True)
                for h, t in zip(out_host, nn):
                    h.copy_(t, non_blocking=True)
        return pops, fe, nn

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(coords, steps):
        """per-step CUDA events on the session stream, L2 flushed between steps (outside the events)."""
        tot = 0.0
        for _ in range(steps):
            with torch.cuda.stream(stream):
                flush_buf.zero_()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step(coords)
            e1.record(stream)
            e1.synchronize()
            tot += e0.elapsed_time(e1)
        t = torch.tensor([tot], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps                   # ms per step, max over ranks

    # ---- warm-up, FFMA peak, timed region --------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                                 # nvidia-smi needs ~1 s to deliver its first sample
    for _ in range(max(args.warmup, 3)):
        step(x_dev)
    barrier()
    tensor_path = sess.gemm_info()[0]                   # 17 <= n_cols <= 256: the scans run on the tensor cores (tcgen05, TF32)
    peak_tflops = sess.tf32_peak(300.0) if tensor_path else sess.ffma_peak(300.0)
    s0 = sess.stats(reset=True)
    barrier()
    ms = timed(x_dev, args.steps)
    barrier()
    s1 = sess.stats()
    # keep the load on until the sampler has seen it (short timed regions), then stop it
    t_end = time.time() + 1.5
    while time.time() < t_end:
        step(x_dev)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    counts = torch.tensor([s1["launches"] - s0["launches"], s1["pairs_evaluated"], s1["pairs_scheduled"]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(counts)
    launches = counts[0]
    evaluated_frac = float(counts[1].item()) / max(1.0, float(counts[2].item()))
    pd = pair_dims_per_step(n, d)
    value = pd / (ms * 1e-3) / 1e9

    # ---- the two pair scans alone, for the roofline (the slower one is the dominant kernel) ---------
    kms = None
    if world == 1:
        def time_scan(fn, reps=5):
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

        pops_loc = torch.zeros((radii.size, n), dtype=torch.int32, device=dev)
        with torch.cuda.stream(stream):
            fe_dev = sess.free_energies(sess.to_frame_order(sess.populations(radii, 0, n))[r_fe].contiguous())
        sess.nn_prepare(fe_dev)
        # nn_scan = seed kernel + block bounds + neighbourhood pass + full pass of nn_kernel (the last one is > 90 % of it)
        nn_ms, nn_pairs = time_scan(lambda: sess.nn_scan(0, n, out=keys_loc))
        pops_ms, pops_pairs = time_scan(lambda: sess.populations(radii, 0, n, out=pops_loc))
        scans = {"nn_kernel (neighbour scan)": (nn_ms, nn_pairs), "pops kernel (population scan)": (pops_ms, pops_pairs)}
        kname = max(scans, key=lambda k: scans[k][0])
        kms, k_pairs = scans[kname]

    # ---- end to end: host buffers through the C ABI ----------------------------------------------
    if world == 1:
        def e2e_step():
            pops = density.calculate_populations(x, radii)
            fe = density.calculate_free_energies(pops[r_fe])
            density.nearest_neighbors(x, fe)
        e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        h2d = 2 * x.nbytes + 4 * n + 4 * n                # coords for each of the two scans, pops, fe
        d2h = radii.size * 4 * n + 4 * n + 2 * 8 * n + 4 * n   # pops, fe, neighbour keys, frame order
        e2e_api = "host-pointer C ABI: dcb200_populations + dcb200_free_energies + dcb200_nearest_neighbors (wall clock)"
    else:
        step(x_pin)
        e2e_ms = timed(x_pin, args.steps)
        h2d = x.nbytes
        d2h = radii.size * 4 * n + 4 * n + 16 * n
        e2e_api = "session C ABI per rank with pinned host buffers: H2D coords, sharded scans, NCCL all-gather, D2H results (CUDA events, max over ranks)"
    e2e_value = pd / (e2e_ms * 1e-3) / 1e9

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "tf32 filter + f32 exact recheck" if tensor_path else "f32", "data": "synthetic",
            "config": {"workload": workload_label(args.workload, cfg),
                       "step": "layout build + populations + free energies + nearest neighbours (+ lower-free-energy neighbour)",
                       "pair_dims_per_step": pd, "parallelism": f"rows sharded over {world} GPU(s), coords replicated, NCCL all-gather of populations and neighbour keys",
                       "l2": "flushed between timed steps (256 MiB write)", "seed": cfg["seed"],
                       "pairs_evaluated_frac": evaluated_frac},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "api": e2e_api},
            "gpu_launches": int(launches.item()),
        }
        if kms is not None:
            pairs = float(n) * float(n)
            achieved = FLOP_EXECUTED_PER_PAIR_DIM * k_pairs * d / (kms * 1e-3) / 1e12
            line["roofline"] = {
                "bound": "tensor" if tensor_path else "fp32", "kernel": kname,
                "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
                "kernel_ms": kms,
                "pairs_evaluated_per_launch": k_pairs, "pairs_evaluated_frac": k_pairs / pairs,
                "evaluated_gpair_dim_per_s": k_pairs * d / (kms * 1e-3) / 1e9,
                "effective_gpair_dim_per_s": pairs * d / (kms * 1e-3) / 1e9,
                "algorithmic_3flop_tflops_on_evaluated_pairs": FLOP_PER_PAIR_DIM * k_pairs * d / (kms * 1e-3) / 1e12,
                "peak_source": ("tcgen05.mma kind::tf32 128x128x8 back to back on resident operands, measured in this run "
                                "(dcb200_ctx_tf32_peak); MEASURED_PEAKS.json holds bf16 only (TF32 runs at half the bf16 rate)") if tensor_path else
                               "FFMA-only microbenchmark in this run (dcb200_ctx_ffma_peak); MEASURED_PEAKS.json has no FP32 entry",
                "traffic": ncu_traffic(("gscan_" + ("nn" if kname.startswith("nn") else "pops")) if tensor_path else kname),
                "other_scans": {k: {"kernel_ms": v[0], "pairs_evaluated_frac": v[1] / pairs,
                                    "achieved_tflops": FLOP_EXECUTED_PER_PAIR_DIM * v[1] * d / (v[0] * 1e-3) / 1e12}
                                for k, v in scans.items() if k != kname},
                "note": ("GEMM-form path (n_cols >= 17): the roofline is the tensor pipe (TF32). achieved = 2 flop per pair.dim x the pairs of "
                         "the 128 x 128 tile pairs the kernel really multiplied; pruned tile pairs are not counted; the headline value "
                         "counts the full N x N matrix") if tensor_path else
                        "compute-bound path (SURVEY.md 8d): the roofline is the FP32 FFMA pipe, not HBM. achieved = 2 flop (one FFMA) per "
                        "pair.dim x the pairs the kernel really evaluated = (warp, tile) scans x 128 rows x 128 columns; tiles out of reach "
                        "of a row block / of a warp's rows are pruned and not counted; the headline value counts the full N x N matrix. "
                        "traffic = dram bytes per launch from the committed ncu capture (profiles/), null if none",
            }
        if not args.no_cpu_baseline:
            try:
                kind, impl, cores = cpu_impl()
                ns = cpu_sample_size(kind, impl, x, float(radii[0]), 12.0)
                sec = cpu_step(kind, impl, np.ascontiguousarray(x[:ns]), float(radii[0]))
                line["cpu_baseline"] = {
                    "value": pair_dims_per_step(ns, d) / sec / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
                    "sample": f"first {ns} frames of the workload, one pops+FE+NN step in {sec:.1f} s (brute-force NN is quadratic; pops is box-pruned, counted as full N x N)"}
            except Exception as ex:                       # the baseline must never take the bench line down
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
        if not args.no_cpu_baseline and world == 1:
            line["reference_cuda"] = reference_cuda_arm(x, radii, d, density)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
