"""Prints the metrics of an `ncu --page raw --csv` dump that the roofline discussion needs."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__cycles_active.avg', 'sm__cycles_active.avg']
for w in want:
    idx = [i for i, h in enumerate(hdr) if h == w]
    if idx:
        print(f"{w:72s}", [rows[1][idx[0]]] + [r[idx[0]][:44] for r in rows[2:]])
print("--- warp stall samples (pcsamp) ---")
for i, h in enumerate(hdr):
    if 'pcsamp_warps_issue_stalled' in h and not h.endswith('_not_issued'):
        vals = [r[i] for r in rows[2:]]
        try:
            if max(float(v) for v in vals) > 0:
                print(f"{h.replace('smsp__pcsamp_warps_issue_stalled_', ''):40s}", vals)
        except ValueError:
            pass
