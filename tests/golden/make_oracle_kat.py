"""Writes tests/golden/oracle_kat.json: SHA-256 known answers of the oracle on the shared keyset (seed 0).
The oracle is 'parity unpinned' against upstream TFHE (no fixture exists); these vectors pin it against itself so
that any change to the restatement is caught, after the restatement was validated by (a) exact-integer == FFT blind
rotation, (b) gate truth tables, (c) the GPU path reproducing it bit for bit."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

h = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
ks = O.keygen(0)
MU8 = 1 << 29
ct = O.encrypt(np.array([MU8, (-MU8) & 0xFFFFFFFF, MU8]), 2.0 ** -25, ks.lwe_key, 5)
out = {"lwe_key": h(ks.lwe_key), "tlwe_key": h(ks.tlwe_key), "bsk": h(ks.bsk), "ksk": h(ks.ksk), "ct": h(ct),
       "pbs_fft": h(O.pbs(ct, MU8, ks)), "pbs_exact_equal": bool(np.array_equal(O.pbs(ct[:1], MU8, ks, exact=True), O.pbs(ct[:1], MU8, ks)))}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "oracle_kat.json"), "w"), indent=1)
print(out)
