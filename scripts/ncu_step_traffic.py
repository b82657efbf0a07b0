#!/usr/bin/env python
"""Summarise an ncu launch list (csv with gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum) of ONE
Score() step into profiles/score_step_traffic.json.  usage: ncu_step_traffic.py launches.csv n_prompts out.json"""
import csv, json, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith('=='))]
hdr = rows[0]; ki = hdr.index('Kernel Name'); mi = hdr.index('Metric Name'); vi = hdr.index('Metric Value'); ii = hdr.index('ID'); ui = hdr.index('Metric Unit')
n = int(sys.argv[2])
launches = {}
for r in rows[1:]:
    d = launches.setdefault(r[ii], {"kernel": r[ki].split('(')[0][-40:]})
    v = float(r[vi].replace(',', ''))
    u = r[ui].lower()
    if 'byte' in u:
        v *= {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)
    if u in ('usecond', 'us'): v *= 1e3
    if u in ('msecond', 'ms'): v *= 1e6
    d[r[mi]] = v
per = {}
tot_t = tot_b = 0.0
for d in launches.values():
    k = d["kernel"]
    t = d.get('gpu__time_duration.sum', 0.0); b = d.get('dram__bytes_read.sum', 0.0) + d.get('dram__bytes_write.sum', 0.0)
    e = per.setdefault(k, {"launches": 0, "time_us": 0.0, "dram_bytes": 0.0})
    e["launches"] += 1; e["time_us"] += t / 1e3; e["dram_bytes"] += b
    tot_t += t / 1e3; tot_b += b
out = {"what": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over ONE Score() step "
               "(cold-cache, serialised launches: compare shares, not absolutes)", "n_prompts": n, "kernels": per,
       "total_time_us": tot_t, "dram_bytes_per_step": tot_b, "dram_bytes_per_prompt": tot_b / n}
for k, e in per.items():
    e["time_share"] = e["time_us"] / tot_t
json.dump(out, open(sys.argv[3], 'w'), indent=1)
print(json.dumps(out, indent=1))
