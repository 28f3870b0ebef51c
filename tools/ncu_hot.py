"""Top stall instructions of a kernel from an ncu report: python tools/ncu_hot.py report.ncu-rep <kernel regex> <launch index> [n]"""
import csv, subprocess, sys, io
rep, rx, idx = sys.argv[1], sys.argv[2], int(sys.argv[3])
n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--launch-skip", str(idx),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
i_src, i_s, i_ex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = []
for k, r in enumerate(rows[2:]):
    if len(r) <= i_ex:
        continue
    try:
        data.append((int(r[i_s] or 0), int(r[i_ex] or 0), k, r[i_src].strip()))
    except ValueError:
        pass
half = len(data) // 2 if len(data) > 2 and data[0][3] == data[len(data) // 2][3] else len(data)
data = data[:half]
tot = sum(d[0] for d in data)
print(rows[0][1][:100], "total samples", tot, "instructions", len(data))
for d in sorted(data, reverse=True)[:n]:
    print(f"{d[0]:7d} {100.0 * d[0] / max(tot, 1):5.1f}%  exec {d[1]:9d}  #{d[2]:5d}  {d[3][:110]}")
