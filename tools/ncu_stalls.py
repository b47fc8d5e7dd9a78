"""Stall reasons of one kernel by CUDA source line, from an ncu report captured with --import-source on.
usage: ncu_stalls.py <report.ncu-rep> <cubin> <mangled-kernel-substring> [top]
(the cubin: cuobjdump -xelf all <object>; it must be the build the report was taken from)"""
import collections, csv, re, subprocess, sys
rep, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 20
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and kern in l)
instr, cur = [], ("?", 0)
for l in sass[start + 1:]:
    if (l.startswith(".text.") or l.startswith(".section")) and len(instr) > 100:
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
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]]) for r in data)
agg = collections.defaultdict(collections.Counter)
for k in range(min(len(data), len(instr))):
    r = data[k]
    for reason in reasons:
        agg[instr[k][0]][reason] += int(r[ix[reason]] or 0)
    agg[instr[k][0]]["all"] += int(r[ix["# Samples"]])
    agg[instr[k][0]]["inst"] += int(r[ix["Instructions Executed"]])
tr = collections.Counter()
for v in agg.values():
    tr.update(v)
print("samples", tot, "warp instr", tr["inst"])
print("  ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in tr.most_common() if k.startswith("stall_") and v > 0.005 * tot))
for reason in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_branch_resolving"):
    print(f"-- top lines by {reason}")
    for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][reason])[:top]:
        print(f"   {f}:{ln:<5d} {reason[6:]} {100 * v[reason] / tot:5.2f}%   all {100 * v['all'] / tot:5.2f}%   inst {100 * v['inst'] / tr['inst']:5.2f}%")
