"""usage: python scripts/ncu_hot.py <report.ncu-rep> [top]  -- top stall sites of the first kernel in an ncu report (source page)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]; body = rows[2:]
ia, isrc, isamp, iex = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[isamp] or 0) for r in body)
print("total samples", tot)
idx = sorted(range(len(body)), key=lambda k: -int(body[k][isamp] or 0))[:top]
for k in sorted(idx):
    r = body[k]
    reasons = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall), reverse=True)[:3]
    print(f"{k:5d} {int(r[isamp]):7d} {100*int(r[isamp])/tot:5.1f}% ex={r[iex]:>9s} {r[isrc].strip()[:70]:70s} {reasons}")
