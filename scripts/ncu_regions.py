"""Stall samples of an `ncu --page source --csv --print-source sass` export, summed per block of SASS instructions.
usage: python scripts/ncu_regions.py <src.csv> [block size, default 50]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
step = int(sys.argv[2]) if len(sys.argv) > 2 else 50
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
keys = ['stall_wait', 'stall_math', 'stall_not_selected', 'stall_selected', 'stall_short_sb', 'stall_long_sb', 'stall_dispatch',
        'stall_branch_resolving', 'stall_no_inst', 'stall_barrier', 'stall_mio']
N = sum(int(r[ix['# Samples']]) for r in data)
print("kernel:", rows[0][1][:100])
print("total samples", N)
tot = {k: sum(int(r[ix[k]] or 0) for r in data) for k in keys}
print("all        100.00%                      " + " ".join(f"{100 * tot[k] / N:6.2f}" for k in keys))
print("range      samples%  execM(max) fp64instr " + " ".join(k[6:12].rjust(6) for k in keys))
for a in range(0, len(data), step):
    seg = data[a:a + step]
    s = sum(int(r[ix['# Samples']]) for r in seg)
    if s < N * 0.002:
        continue
    ex = max(int(r[ix['Instructions Executed']]) for r in seg) / 1e6
    f = sum(1 for r in seg if any(o in r[ix['Source']] for o in ('DFMA', 'DADD', 'DMUL')))
    marks = [r[ix['Source']].split()[0] for r in seg if any(o in r[ix['Source']] for o in ('USETMAXREG', 'SHFL', 'STS.128', 'BAR.SYNC', 'F2I', 'UBLKCP', 'ATOMS'))]
    print(f"{a:4d}-{a + step:4d} {100 * s / N:6.2f}%  {ex:8.1f} {f:4d}      " + " ".join(f"{100 * sum(int(r[ix[k]] or 0) for r in seg) / N:6.2f}" for k in keys) + "  " + ",".join(sorted(set(marks))))
