"""Compact summary of one ncu report (first kernel): duration, instructions per cell, issue/pipe utilisation, stalls per issue,
DRAM/L2 traffic per cell, opcode mix and a bucketed walk along the SASS.  usage: python tools/ncu_summary.py rep ncells [bucket]"""
import collections, csv, io, re, subprocess, sys
rep, ncells = sys.argv[1], float(sys.argv[2])
B = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, v = rows[0], rows[2]
m = {n: x for n, x in zip(h, v)}
def f(k):
    try: return float(m[k].replace(",", ""))
    except Exception: return float("nan")
dur = f("gpu__time_duration.sum")
print(f"kernel {m.get('Kernel Name','?')[:60]} grid {m.get('launch__grid_size')} block {m.get('launch__block_size')} regs {m.get('launch__registers_per_thread')}")
print(f"duration {dur:.3f} {rows[1][h.index('gpu__time_duration.sum')]}  -> {ncells / dur / 1e3:.2f} M cells/s (under ncu)")
print(f"inst/cell {f('smsp__inst_executed.sum') / ncells:.0f}  issue active {f('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f}%  warps active {f('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f}%")
print(f"dram read {f('dram__bytes_read.sum') * 1e9 / ncells:.0f} B/cell write {f('dram__bytes_write.sum') * 1e9 / ncells:.0f} B/cell (units {rows[1][h.index('dram__bytes_read.sum')]}); L2->L1 {f('l1tex__m_xbar2l1tex_read_bytes.sum')} {rows[1][h.index('l1tex__m_xbar2l1tex_read_bytes.sum')]}")
print(f"smem wavefronts/cell {f('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum') / ncells:.0f} ({f('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'):.1f}% of peak)")
for k in h:
    if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
        x = f(k)
        if x > 0.08: print(f"  stall {k.split('stalled_')[1].split('_per_issue')[0]:22s} {x:.2f}")
for k in ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"):
    if k in m: print(f"  {k.split('.')[0]:44s} {m[k]}%")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ci, cs, cz = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
data = []
for r in rows[hi + 1:]:
    try: data.append((r[cz], int(r[cs] or 0), int(r[ci] or 0)))
    except Exception: pass
op = collections.Counter(); tot = 0
for s, _, n in data:
    mm = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z0-9_]+(?:\.[0-9]+)?)", s)
    if mm: op[mm.group(1)] += n; tot += n
print(f"SASS: {len(data)} static instructions ({len(data) * 16 // 1024} KB); executed {tot / ncells:.0f}/cell")
print("  " + "  ".join(f"{k} {c / ncells:.0f}" for k, c in op.most_common(28)))
if B:
    ts = sum(d[1] for d in data)
    for b in range(0, len(data), B):
        ch = data[b:b + B]
        ops = collections.Counter()
        for s, _, n in ch:
            mm = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", s)
            if mm: ops[mm.group(1)] += 1
        key = [k for k in ("LDGSTS", "CREDUX", "DMMA", "LDG", "STG", "MUFU", "SHFL", "VOTE", "DFMA", "LDS", "STS") if ops[k]]
        print(f"{b:5d} samp {100 * sum(d[1] for d in ch) / ts:5.1f}% inst/cell {sum(d[2] for d in ch) / ncells:6.0f}  " + " ".join(f"{k}:{ops[k]}" for k in key))
