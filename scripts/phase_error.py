"""Pre-rounding error distribution of the FFT-based blind rotation (north star: 'torus phase error ... reported as a distribution').
GPU part needs the debug build:  RS_NVCC_EXTRA=-DRS_ERR_STATS python -c "from redsec_b200 import build; build.build(force=True)"
The blind rotation rounds each inverse-transform output x to rint(x); as long as |x - rint(x)| < 0.5 the result is the exact
integer convolution, i.e. the ciphertext equals the exact-integer reference bit for bit (delta-phase = 0).  This script reports the
distribution of |x - rint(x)| for the GPU kernel (350 steps x 2048 outputs per bootstrap) and for the CPU oracle's FFT variant."""
import ctypes as C, json, sys, numpy as np
sys.path.insert(0, '.')
from oracle import oracle as O
ks = O.keygen(0)
count = int(sys.argv[1]) if len(sys.argv) > 1 else 592
rng = np.random.default_rng(3)
msg = np.where(rng.integers(0, 2, count) == 1, 0x20000000, 0xE0000000).astype(np.uint32)
ct = O.encrypt(msg, 2.0**-25, ks.lwe_key, 5)
out = {"bins": ["<1e-6", "<1e-5", "<1e-4", "<1e-3", "<1e-2", "<1e-1", "<=0.5"], "bound_for_exactness": 0.5}
st = np.zeros(3)
n_cpu = min(count, 16)
O.pbs(ct[:n_cpu], 0x20000000, ks, stats=st)
out["cpu_oracle_fft"] = {"bootstraps": n_cpu, "max_abs_error": float(st[0]), "mean_abs_error": float(st[1] / max(st[2], 1)), "samples": int(st[2])}
try:
    import redsec_b200 as rs
    eng = rs.Engine(0); eng.load_eval_key(ks.bsk, ks.ksk)
    got = eng.download(eng.pbs(eng.upload(ct), 0x20000000))
    hist = (C.c_ulonglong * 8)(); mx = C.c_double()
    if eng.lib.rs_debug_err_stats(hist, C.byref(mx)) == 0:
        h = [int(v) for v in hist[:7]]
        out["gpu_ws_kernel"] = {"bootstraps": count, "samples": int(sum(h)), "histogram": h, "max_abs_error": mx.value,
                                "ciphertexts_equal_exact_oracle": bool(np.array_equal(got[:4], O.pbs(ct[:4], 0x20000000, ks, exact=True)))}
except AttributeError:
    out["gpu_ws_kernel"] = "library not built with -DRS_ERR_STATS"
print(json.dumps(out, indent=1))
