"""Instruction mix of a kernel from an ncu report, bucketed by how often each SASS instruction executed:
python tools/ncu_buckets.py report.ncu-rep <kernel regex> [launch index]"""
import csv, io, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]
idx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{rx}",
                      "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
iS, iE, iW = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
iX = hdr.index("L1 Wavefronts Shared Excessive")
data = [(int(r[iE] or 0), int(r[iW] or 0), r[iS].strip(), int(r[iX] or 0)) for r in rows[2:] if len(r) > max(iW, iE, iX) and r[iE].isdigit()]
tot = sum(d[0] for d in data)
tots = sum(d[1] for d in data)
print("kernel", rows[0][1][:90], "| static instr", len(data), "| executed", tot, "| samples", tots)
# buckets: group consecutive instructions with similar exec counts
b = collections.OrderedDict()
for e, w, s, x in data:
    if e == 0:
        continue
    key = round(e, -len(str(e)) + 2)
    d = b.setdefault(key, [0, 0, 0, collections.Counter(), 0])
    d[0] += 1; d[1] += e; d[2] += w; d[3][s.split()[0] if not s.startswith("@") else s.split()[1]] += 1; d[4] += x
for key, (n, e, w, ops, x) in sorted(b.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"exec~{key:>10d}: {n:5d} static, {100.0 * e / tot:5.1f}% of executed, {100.0 * w / max(tots, 1):5.1f}% of samples, smem excess {x}; "
          + " ".join(f"{o}:{c}" for o, c in ops.most_common(12)))
