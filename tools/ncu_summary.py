"""Summarise an .ncu-rep (first kernel): key metrics, stall reasons, hottest SASS lines.
usage: python tools/ncu_summary.py report.ncu-rep [n_lines]"""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = dict(zip(hdr, zip(units, vals)))
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sectors.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for k in keys:
    if k in m: print(f"{k:90s} {m[k][1]} {m[k][0]}")
print("-- stall reasons (warps per issue-active) --")
st = [(float(v[1]), k.split("issue_stalled_")[1].split("_per")[0]) for k, v in m.items()
      if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and v[1]]
for v, k in sorted(st, reverse=True)[:9]: print(f"  {k:28s} {v:.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
body = [r for r in rows[2:] if len(r) > 5 and r[2].strip().isdigit() and r[1].strip()]
tot = sum(int(r[2]) for r in body)
c = Counter(); ex = Counter()
for r in body:
    op = [o for o in r[1].split() if not o.startswith('@')][0]
    c[op] += int(r[2]); ex[op] += int(r[5])
print("-- stall samples by opcode --")
for op, n in c.most_common(12): print(f"  {op:24s} {100*n/tot:5.1f}%  executed {ex[op]}")
print("-- hottest SASS lines --")
for r in sorted(body, key=lambda r: -int(r[2]))[:ntop]:
    print(f"  {r[0][-5:]} {100*int(r[2])/tot:5.1f}% exec {r[5]:>10s}  {r[1].strip()[:100]}")
