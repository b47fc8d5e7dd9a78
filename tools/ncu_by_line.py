"""Aggregate an ncu SASS-level source page by CUDA source line.
usage: ncu_by_line.py <report.ncu-rep> <cubin> <mangled-kernel-substring> [top]"""
import collections, csv, os, re, subprocess, sys
rep, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and kern in l)
instr, cur = [], ("?", 0)
for l in sass[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        if len(instr) > 100:
            break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
    elif re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l):
        instr.append((cur, l.strip()))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hdr = next(r for r in rows if r and r[0] == "Address")
data = [r for r in rows[rows.index(hdr) + 1:] if len(r) > 5]
ix = {h: k for k, h in enumerate(hdr)}
assert abs(len(data) - len(instr)) < 8, (len(data), len(instr))
agg = collections.defaultdict(lambda: [0, 0, 0])
byfile = collections.defaultdict(lambda: [0, 0, 0])
for k in range(min(len(data), len(instr))):
    (f, ln), _ = instr[k]
    r = data[k]
    v = (int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]]), int(r[ix["# Samples"]]))
    for j in range(3):
        agg[(f, ln)][j] += v[j]
        byfile[f][j] += v[j]
tot = [sum(v[j] for v in agg.values()) for j in range(3)]
print(f"warp instr {tot[0]:.4g}  thread instr {tot[1]:.4g}  avg lanes {tot[1]/tot[0]:.2f}  samples {tot[2]}")
for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:22s} inst {100*v[0]/tot[0]:5.1f}%  samples {100*v[2]/tot[2]:5.1f}%  lanes {v[1]/max(v[0],1):5.1f}")
src = {}
for f in byfile:
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "light_transport_b200", "csrc", f)
    if os.path.exists(p):
        src[f] = open(p).read().split("\n")
for (f, ln), (ie, te, smp) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    s = src[f][ln - 1].strip()[:100] if f in src and ln - 1 < len(src[f]) else ""
    print(f"{f:18s}:{ln:4d} inst {100*ie/tot[0]:5.1f}% samp {100*smp/tot[2]:5.1f}% lanes {te/max(ie,1):5.1f} | {s}")
