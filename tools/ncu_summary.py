#!/usr/bin/env python
"""Summarise one `ncu --set full` capture (raw csv + optional SASS source csv) of a kernel.
usage: tools/ncu_summary.py gpurun_out/ncu_<tag> [samples_per_launch]"""
import csv, sys
from collections import Counter
tag = sys.argv[1]
nsamp = float(sys.argv[2]) if len(sys.argv) > 2 else 2048 * 15388.0
rows = list(csv.reader(open(tag + "_raw.csv")))
hdr, vals = rows[0], rows[-1]
m = {}
for h, v in zip(hdr, vals):
    try:
        m[h] = float(v.replace(",", ""))
    except ValueError:
        m[h] = v
def g(k): return m.get(k, float("nan"))
cyc = g("sm__cycles_elapsed.max") if "sm__cycles_elapsed.max" in m else g("sm__cycles_elapsed.avg")
print("kernel:", m.get("Kernel Name"), "| duration us:", g("gpu__time_duration.sum") / 1e3 if g("gpu__time_duration.sum") > 1e4 else g("gpu__time_duration.sum"))
print(f"SM cycles {cyc:.0f} -> {cyc * 148 / nsamp:.2f} clk/sample/SM; regs {g('launch__registers_per_thread'):.0f}; warps/SM {g('sm__warps_active.avg.per_cycle_active'):.1f}")
print(f"issue slots busy {g('sm__inst_issued.avg.pct_of_peak_sustained_active'):.1f}%  fma pipe {g('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'):.1f}%  fmaheavy {g('sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f}%  alu {g('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'):.1f}%  lsu {g('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'):.1f}%")
print(f"smem wavefronts {g('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'):.1f}% of peak ({g('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum') / nsamp:.2f}/sample), bank conflicts {g('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum') / nsamp:.2f}/sample")
print(f"dram read {g('dram__bytes_read.sum')} write {g('dram__bytes_write.sum')} (units as exported); dram throughput {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f}%")
print(f"instructions {g('smsp__inst_executed.sum'):.0f} -> {g('smsp__inst_executed.sum') * 32 / nsamp:.0f} thread-instr/sample")
st = sorted(((v, h) for h, v in m.items() if "warps_issue_stalled" in h and h.endswith("per_issue_active.ratio") and isinstance(v, float)), reverse=True)
print("stalls (warp-cycles per issue):", ", ".join(f"{h.split('stalled_')[1].split('_per_')[0]} {v:.2f}" for v, h in st[:8]))
try:
    src = list(csv.reader(open(tag + "_src.csv")))
    hd, data = src[1], src[2:]
    iS, iE, iW = hd.index("Source"), hd.index("Instructions Executed"), hd.index("Warp Stall Sampling (All Samples)")
    iSh, iX = hd.index("L1 Wavefronts Shared"), hd.index("L1 Wavefronts Shared Excessive")
    tot = sum(int(r[iE]) for r in data); tots = sum(int(r[iW]) for r in data)
    c = Counter(); cs = Counter(); wf = Counter(); wx = Counter()
    for r in data:
        t = r[iS].split()
        op = t[1] if t[0].startswith("@") else t[0]
        c[op] += int(r[iE]); cs[op] += int(r[iW]); wf[op] += int(r[iSh] or 0); wx[op] += int(r[iX] or 0)
    print("top opcodes: " + "; ".join(f"{op} {n / tot * 100:.1f}%i/{cs[op] / tots * 100:.1f}%s" for op, n in c.most_common(14)))
    print("smem wavefronts by opcode: " + "; ".join(f"{op} {wf[op] / nsamp:.2f} (excess {wx[op] / nsamp:.2f})" for op in wf if wf[op]))
    bars = [i for i, r in enumerate(data) if "BAR.SYNC" in r[iS]]
    prev = 0
    for b in bars + [len(data)]:
        seg = data[prev:b]
        print(f"  SASS [{prev},{b}): {sum(int(r[iE]) for r in seg) / tot * 100:.1f}% instr, {sum(int(r[iW]) for r in seg) / tots * 100:.1f}% stall samples")
        prev = b
except FileNotFoundError:
    pass
