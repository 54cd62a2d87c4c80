"""Small hot-path run for compute-sanitizer (memcheck / racecheck): 6 sign bootstraps, 6 test-vector bootstraps, 9 NAND gates,
checked against the oracle.  usage: compute-sanitizer --tool memcheck python scripts/sanitize_run.py"""
import sys, numpy as np
sys.path.insert(0, '.')
import redsec_b200 as rs
from redsec_b200 import client
from oracle import oracle as O
ks = client.keygen(0)
eng = rs.Engine(0); eng.load_eval_key(ks.bsk, ks.ksk)
oks = O.KeySet(ks.lwe_key, ks.tlwe_key, ks.bsk, ks.ksk)
rng = np.random.default_rng(1)
x = client.encrypt(rng.integers(-900, 900, 6) * client.UNIT, ks.lwe_key, client.SECALPHA, 3)
assert np.array_equal(eng.download(eng.pbs(eng.upload(x), client.UNIT)), O.pbs(x, client.UNIT, oks))
luts = (rng.integers(-1000, 1000, size=(4, 1024)) * client.UNIT & 0xFFFFFFFF).astype(np.uint32)
assert np.array_equal(eng.download(eng.pbs_lut(eng.upload(x), luts)), O.pbs_lut(x, luts, oks))
a = client.encrypt_bits(rng.integers(0, 2, 9), ks.lwe_key, seed=1); b = client.encrypt_bits(rng.integers(0, 2, 9), ks.lwe_key, seed=2)
assert np.array_equal(eng.download(eng.gate("NAND", eng.upload(a), eng.upload(b), client.EIGHTH)), O.gate("NAND", a, b, client.EIGHTH, oks))
eng.close()
print("sanitize run ok")
