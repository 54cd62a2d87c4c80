"""compute-sanitizer workload for the end-of-round-2 kernels: the producer-warp blind rotation in all three row-split modes (counts
3, 150, 300, 600 -> SPLIT 4, 2, 2, 1), the tensor-core keyswitch (count 2100 through rs_keyswitch, ragged last tile), each compared
with the oracle on a sample.  usage: compute-sanitizer --tool memcheck python scripts/sanitize_run2.py"""
import sys, numpy as np
sys.path.insert(0, '.')
import redsec_b200 as rs
from redsec_b200 import client
from oracle import oracle as O
ks = client.keygen(0)
eng = rs.Engine(0); eng.load_eval_key(ks.bsk, ks.ksk)
oks = O.KeySet(ks.lwe_key, ks.tlwe_key, ks.bsk, ks.ksk)
rng = np.random.default_rng(1)
for n in (3, 150, 300, 600):
    x = client.encrypt(rng.integers(-900, 900, n) * client.UNIT, ks.lwe_key, client.SECALPHA, 3)
    got = eng.download(eng.pbs(eng.upload(x), client.UNIT))
    idx = np.unique(np.concatenate([rng.choice(n, min(n, 4), replace=False), [0, n - 1]]))
    assert np.array_equal(got[idx], O.pbs(x[idx], client.UNIT, oks)), n
ext = rng.integers(0, 2 ** 32, size=(2100, 1025), dtype=np.uint64).astype(np.uint32)
got = eng.keyswitch(ext)
idx = np.array([0, 1, 255, 256, 2047, 2048, 2099])
assert np.array_equal(got[idx], O.keyswitch(ext[idx], oks))
eng.close()
print("sanitize run 2 ok")
