#!/usr/bin/env python
"""ncu .ncu-rep (--set full, one launch) -> small JSON summary for profiles/.  usage: ncu_kernel_summary.py rep out.json 'description'"""
import csv, json, subprocess, sys
rep, out, what = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__inst_executed.sum',
        'sm__inst_executed.sum.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'smsp__warps_eligible.avg.per_cycle_active']
m = {h: {"unit": units[i], "value": vals[i]} for i, h in enumerate(hdr) if h in keep}
stall = {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''): float(vals[i]) for i, h in enumerate(hdr)
         if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')}
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines())); sh = rows[1]; data = rows[2:]; ix = {h: i for i, h in enumerate(sh)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
stalls = [h for h in sh if h.startswith('stall_') and 'Not' not in h]
tot = sum(f(r, '# Samples') for r in data)
top = []
for r in sorted(data, key=lambda r: -f(r, '# Samples'))[:12]:
    st = max(((f(r, s), s) for s in stalls))
    top.append({"sass": r[ix['Source']].strip()[:70], "samples_pct": round(100 * f(r, '# Samples') / max(tot, 1), 2), "executed": f(r, 'Instructions Executed'),
                "main_stall": st[1]})
json.dump({"what": what, "metrics": m, "stall_cycles_per_issued_instruction": dict(sorted(stall.items(), key=lambda kv: -kv[1])[:8]),
           "hottest_instructions": top}, open(out, 'w'), indent=1)
print(open(out).read()[:3000])
