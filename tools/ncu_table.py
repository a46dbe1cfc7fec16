"""One markdown row per ncu report: duration, DRAM GB/s (read+write bytes / duration), FP64 TFLOP/s where given, issue-active,
tensor pipe, top stalls.  usage: python tools/ncu_table.py name=units:rep ...  (units = items per launch, for the rate)"""
import csv, io, subprocess, sys
print("| kernel (capture) | launch | time | rate | DRAM read+write | % of HBM peak (6454.6 GB/s) | issue active | tensor pipe | FP64 pipe | regs | top stalls (cycles per issue) |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for a in sys.argv[1:]:
    name, rest = a.rsplit("=", 1); units, rep = rest.split(":", 1); units = float(units)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u, v = rows[0], rows[1], rows[2]
    m = dict(zip(h, v)); un = dict(zip(h, u))
    def f(k):
        try: return float(m[k].replace(",", ""))
        except Exception: return float("nan")
    def to_s(k):
        x = f(k); return x * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(un[k], 1e-3)
    def to_b(k):
        x = f(k); return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(un[k], 1.0)
    t = to_s("gpu__time_duration.sum")
    by = to_b("dram__bytes_read.sum") + to_b("dram__bytes_write.sum")
    st = sorted(((f(k), k.split("stalled_")[1].split("_per_issue")[0]) for k in h
                 if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")
                 and "selected" not in k), reverse=True)[:3]
    kn = m.get("Kernel Name", "?").split("(")[0][-44:]
    print(f"| `{kn}` ({name}) | {m.get('launch__grid_size')} x {m.get('launch__block_size')} | {t * 1e3:.3f} ms | {units / t / 1e6:.1f} M/s | "
          f"{by / 1e9:.2f} GB = {by / t / 1e9:.0f} GB/s | {100 * by / t / 6454.6e9:.0f} % | {f('smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} % | "
          f"{f('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.0f} % | {f('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):.0f} % | "
          f"{m.get('launch__registers_per_thread')} | " + ", ".join(f"{n} {x:.2f}" for x, n in st) + " |")
