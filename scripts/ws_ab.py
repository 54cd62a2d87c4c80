"""A/B of blind-rotate build knobs on one box: each setting runs in its own process (the knobs are read at rs_create), times
rs_pbs_batch at a few sizes and prints a digest of the output ciphertexts -- every setting must print the same digest.
usage: python scripts/ws_ab.py "RS_WS_PRODUCER=0" "RS_WS_PRODUCER=1" ... [--counts 592,4096,444]"""
import hashlib, os, subprocess, sys, time

CHILD = r'''
import sys, time, hashlib, numpy as np
sys.path.insert(0, '.')
import redsec_b200 as rs
from redsec_b200 import client
counts = [int(c) for c in sys.argv[1].split(',')]
ks = client.keygen(0)
eng = rs.Engine(0); eng.load_eval_key(ks.bsk, ks.ksk)
rng = np.random.default_rng(0)
for n in counts:
    ct = client.encrypt(rng.integers(-500, 500, n) * client.UNIT, ks.lwe_key, client.SECALPHA, 3)
    d = eng.upload(ct); out = eng.alloc(n)
    eng.pbs(d, client.UNIT, out); eng.sync()
    eng.profile(True); eng.profile_reset()
    reps = 3 if n < 20000 else 2
    for _ in range(reps):
        eng.pbs(d, client.UNIT, out)
    eng.sync()
    br, _ = eng.profile_get(0)
    eng.profile(False)
    h = hashlib.sha256(eng.download(out).tobytes()).hexdigest()[:16]
    print(f"  count {n:6d}: blind-rotate {br/reps:9.3f} ms  -> {n/(br/reps)*1e3:9.0f} PBS/s   digest {h}", flush=True)
    d.free(); out.free()
'''

args = [a for a in sys.argv[1:] if not a.startswith('--counts')]
counts = '592,4096'
for a in sys.argv[1:]:
    if a.startswith('--counts'):
        counts = a.split('=', 1)[1]
for setting in args or ['']:
    env = dict(os.environ)
    for kv in setting.split():
        k, v = kv.split('=', 1)
        env[k] = v
    print(f"== {setting or '(default)'}", flush=True)
    r = subprocess.run([sys.executable, '-c', CHILD, counts], env=env, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout)
    if r.returncode != 0:
        sys.stdout.write(r.stderr[-2000:])
