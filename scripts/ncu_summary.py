"""Summarises an `ncu --page raw --csv` dump: the metrics the roofline discussion needs, the warp stall samples, and
(with --traffic OUT.json) dram bytes per launch of each kernel for bench.py's `roofline.traffic`.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv
    python scripts/ncu_summary.py raw.csv [--traffic profiles/ncu_traffic.json]
"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__cycles_active.avg', 'sm__cycles_active.avg', 'gpc__cycles_elapsed.max']
for w in want:
    if w in col:
        print(f"{w:72s}", [units[col[w]]] + [r[col[w]][:44] for r in data])
print("--- warp stall samples (pcsamp) ---")
for i, h in enumerate(hdr):
    if 'pcsamp_warps_issue_stalled' in h and not h.endswith('_not_issued'):
        vals = [r[i] for r in data]
        try:
            if max(float(v) for v in vals) > 0:
                print(f"{h.replace('smsp__pcsamp_warps_issue_stalled_', ''):40s}", vals)
        except ValueError:
            pass

if '--traffic' in sys.argv:
    # python scripts/ncu_summary.py raw.csv --traffic profiles/ncu_traffic.json --workload C3 [--source "capture name"]
    out = sys.argv[sys.argv.index('--traffic') + 1]
    wl = sys.argv[sys.argv.index('--workload') + 1]
    src = sys.argv[sys.argv.index('--source') + 1] if '--source' in sys.argv else sys.argv[1]

    def to_bytes(v, unit):
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]
        return float(v.replace(',', '')) * scale

    traffic = {}
    for r in data:
        name = r[col['Kernel Name']]
        key = 'nn' if 'nn_kernel' in name else 'pops' if 'pops' in name else None
        if key is None:
            continue
        ms = float(r[col['gpu__time_duration.sum']].replace(',', ''))
        if 'dram__bytes_read.sum' in col:
            b = to_bytes(r[col['dram__bytes_read.sum']], units[col['dram__bytes_read.sum']]) + \
                to_bytes(r[col['dram__bytes_write.sum']], units[col['dram__bytes_write.sum']])
        else:       # section captures (no --set full): bytes per second over the launch duration
            c = col['dram__bytes.sum.per_second']
            b = to_bytes(r[c], units[c].split('/')[0]) * ms * 1e-3
        if key not in traffic or ms > traffic[key][1]:        # the longest launch of a kernel = its main pass
            traffic[key] = (b, ms)
    try:
        allw = json.load(open(out))
    except Exception:
        allw = {}
    allw[wl] = dict({k: v[0] for k, v in traffic.items()}, source=src)
    json.dump(allw, open(out, 'w'), indent=1)
    print("traffic (dram bytes per launch) ->", out, allw[wl])
