"""Tensor-core keyswitch (variant 2, tcgen05.mma kind::i8) against the default shared-memory kernel and the oracle; timing at 2^16."""
import os, sys, time
import numpy as np
sys.path.insert(0, '.')
import redsec_b200 as rs
from oracle import oracle as O

ks = O.keygen(0)
rng = np.random.default_rng(4)
for swap in (0,):
    eng = rs.Engine(0); eng.load_eval_key(ks.bsk, ks.ksk)
    ok_all = True
    for count in (5, 256, 300, 1000):
        ext = rng.integers(0, 2 ** 32, size=(count, 1025), dtype=np.uint64).astype(np.uint32)
        eng.set_ks_variant(0); ref = eng.keyswitch(ext)
        try:
            eng.set_ks_variant(2); got = eng.keyswitch(ext)
        except Exception as e:
            print("swap", swap, "count", count, "FAILED:", str(e)[:200]); ok_all = False; break
        same = np.array_equal(got, ref)
        ok_all &= same
        bad_rows = int((got != ref).any(axis=1).sum())
        print(f"swap {swap} count {count}: equal to the default kernel: {same} (rows differing {bad_rows}, words differing {int((got != ref).sum())})")
        if count == 5:
            print("   default == oracle:", np.array_equal(ref, O.keyswitch(ext, ks)))
    if ok_all:
        n = 1 << 16
        ct = O.encrypt(np.full(n, 1 << 29), 2.0 ** -25, ks.lwe_key, 3)
        d = eng.upload(ct); out = eng.alloc(n)
        for v in (0, 2):
            eng.set_ks_variant(v)
            eng.pbs(d, 1 << 29, out); eng.sync()
            eng.profile(True); eng.profile_reset()
            eng.pbs(d, 1 << 29, out); eng.pbs(d, 1 << 29, out); eng.sync()
            ks_ms, ks_n = eng.profile_get(1)
            eng.profile(False)
            res = eng.download(out)
            print(f"variant {v}: keyswitch {ks_ms / 2:.2f} ms per 2^16 ({ks_n} launches incl. prep/init)", "hash", int(res.astype(np.uint64).sum() % (1 << 61)))
    eng.close()
    if ok_all:
        break
