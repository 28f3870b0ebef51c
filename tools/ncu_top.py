import csv, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
iS, iE, iW = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
print(rows[0][:2]); print("total samples", sum(int(r[iW]) for r in data))
idx = sorted(range(len(data)), key=lambda i: -int(data[i][iW]))[:25]
for i in sorted(idx):
    r = data[i]
    top = sorted(((int(r[c]), hdr[c]) for c in stall_cols if r[c].isdigit()), reverse=True)[:2]
    print("%4d %9s %6s  %-60s %s" % (i, r[iE], r[iW], r[iS].strip()[:60], top))
