"""Print the hot SASS lines (executed count, stall samples) of one kernel in an ncu report."""
import csv, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
iS, iE, iW = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
print(rows[0][:2])
mx = max(int(r[iE]) for r in data)
tot_s = sum(int(r[iW]) for r in data)
print("total samples", tot_s, "max exec", mx)
for r in data:
    if int(r[iE]) >= thr * mx:
        print("%9s %6s  %s" % (r[iE], r[iW], r[iS].strip()[:120]))
