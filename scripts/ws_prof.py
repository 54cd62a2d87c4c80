"""Phase timers of the warp-specialised blind rotation (debug build: RS_NVCC_EXTRA=-DRS_WS_PROF python -m redsec_b200.build --force).
usage: python scripts/ws_prof.py [count ...]"""
import ctypes as C, sys, numpy as np
sys.path.insert(0, '.')
import redsec_b200 as rs
from oracle import oracle as O
ks = O.keygen(0)
eng = rs.Engine(0); eng.load_eval_key(ks.bsk, ks.ksk)
lib = eng.lib
FRONT = ["wait accready", "rot_diff/src", "producer duty", "wait xempty", "digits+pass1+store", "-", "-", "loop"]
BACK = ["wait xfull", "lds+pass2+shfl+pass3", "wait slab", "MAC", "inverse+acc", "-", "-", "row head"]
for count in [int(a) for a in sys.argv[1:]] or [148, 592]:
    ct = O.encrypt(np.full(count, 0x20000000), 2.0**-25, ks.lwe_key, 3)
    dev = eng.upload(ct); out = eng.alloc(count)
    eng.pbs(dev, 0x20000000, out); eng.sync()
    buf = (C.c_longlong * 96)()
    lib.rs_debug_ws_prof(buf)
    a = np.array(buf[:]).reshape(12, 8)
    print(f"== count {count} (CTA 0): cycles per row (7000 rows)")
    for w in range(12):
        names = BACK if w < 8 else FRONT
        tot = a[w].sum()
        if tot == 0: continue
        print(f" warp {w:2d} {'back ' if w < 8 else 'front'} total {tot/7000:7.0f}: " + ", ".join(f"{n} {a[w][k]/7000:6.0f}" for k, n in enumerate(names) if n != "-"))
    rw = (C.c_longlong * 60)()
    if hasattr(lib, "rs_debug_ws_rowwait"):
        lib.rs_debug_ws_rowwait(rw)
        r = np.array(rw[:]).reshape(3, 20) / 350.0
        print("   back warp 0, cycles per step waiting for row r of the step :", " ".join(f"{x:5.0f}" for x in r[0]))
        print("   front warp 8, cycles per step waiting for a free slot, row r:", " ".join(f"{x:5.0f}" for x in r[1]))
        print("   front warp 8, cycles per step waiting for accumulator 0 / 1 :", " ".join(f"{x:5.0f}" for x in r[2][:2]))
