"""Static SASS instruction count per CUDA source line of one kernel (no GPU needed).
usage: python tools/sass_lines.py <object or .so> <kernel-name substring> [top]
Runs cuobjdump -xelf + nvdisasm -g and aggregates the `//## File ..., line N` markers."""
import os, re, subprocess, sys, tempfile, collections, glob
obj, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cnt = collections.Counter(); ops = collections.Counter(); total = 0
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
    on = False; line = None
    for l in dis:
        if l.startswith(".text."):
            on = pat in l
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if m:
            cnt[line] += 1; total += 1; ops[m.group(1)] += 1
print("total instructions:", total, "=", total * 16 // 1024, "KB")
print("by opcode:", ", ".join(f"{k} {v}" for k, v in ops.most_common(25)))
src = {}
for (f, n), c in cnt.most_common(top):
    if f not in src:
        cands = glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "**", f), recursive=True)
        src[f] = open(cands[0]).read().splitlines() if cands else []
    text = src[f][n - 1].strip()[:100] if n - 1 < len(src[f]) else ""
    print(f"{c:6d}  {f}:{n}  {text}")
