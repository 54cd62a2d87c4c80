import sys, ctypes, numpy as np
sys.path.insert(0, '.')
import redsec_b200 as rs
from redsec_b200 import _lib
from oracle import oracle as O
ks = O.keygen(0)
eng = rs.Engine(0)
eng.load_eval_key(ks.bsk, ks.ksk)
lib = _lib.load()
buf = (ctypes.c_ulonglong * 8)()
for variant in (0, 1):
    eng.set_tuning(variant)
    for count in (592, 2368):
        ct = O.encrypt(np.full(count, 0x20000000), 2.0**-25, ks.lwe_key, 3)
        dev = eng.upload(ct); out = eng.alloc(count)
        eng.pbs(dev, 0x20000000, out); eng.sync()
        lib.rs_debug_stats(buf, 1)
        eng.pbs(dev, 0x20000000, out); eng.sync()
        lib.rs_debug_stats(buf, 1)
        v = list(buf)
        print(f"variant={variant} count={count}: warp-cycles {v[0]:.3e} wait-cycles {v[1]:.3e} ({100*v[1]/v[0]:.1f}%) first-test failures {v[2]} of {v[3]} ({100*v[2]/max(v[3],1):.1f}%) per group {v[4:8]}; avg wait cyc/row {v[1]/max(v[3],1):.0f}")
