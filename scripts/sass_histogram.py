"""SASS instruction histograms of the product kernels (cuobjdump -sass on the in-tree libredsec_b200.so, no GPU needed):
the evidence for the TMA / mbarrier / setmaxnreg / tensor-memory claims in DESIGN.md.  Writes profiles/r2_sass_histograms.txt."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "redsec_b200", "libredsec_b200.so")
KEYS = ["UBLKCP", "UTMALDG", "SYNCS", "USETMAXREG", "DFMA", "DADD", "DMUL", "DMMA", "F2I", "I2F", "LDS", "STS", "LDG", "STG", "ATOMG", "RED",
        "SHFL", "BAR", "MEMBAR", "UTCCP", "LDTM", "UTCBAR", "UTCHMMA", "UTCIMMA", "UTCQMMA", "UTCATOMSWS", "TCGEN"]
WANT = [("blind_rotate_ws_kernelILi5ELi3ELi1ELb0ELb1", "blind_rotate_ws_kernel<5,3,SPLIT=1,STRESS=0,PRODUCER=1> (default, full waves; 16 warps, BSK producer warp)"),
        ("blind_rotate_ws_kernelILi5ELi3ELi2ELb0ELb1", "blind_rotate_ws_kernel<5,3,2,0,1> (row-split, <= 2 ciphertexts per SM)"),
        ("blind_rotate_ws_kernelILi5ELi3ELi4ELb0ELb1", "blind_rotate_ws_kernel<5,3,4,0,1> (row-split, <= 1 ciphertext per SM)"),
        ("blind_rotate_ws_kernelILi5ELi3ELi1ELb0ELb0", "blind_rotate_ws_kernel<5,3,1,0,PRODUCER=0> (round-1 shape: 12 warps, front warps claim the slabs; RS_WS_PRODUCER=0)"),
        ("keyswitch_mma_kernel", "keyswitch_mma_kernel (default keyswitch for batches >= 2048: tcgen05.mma kind::i8 = UTCIMMA, accumulators in TMEM, LDTM epilogue)"),
        ("keyswitch_tiled_kernelILi64", "keyswitch_tiled_kernel<64> (shared-memory gather keyswitch, small batches)"),
        ("lwe_conv_kernelILb0", "lwe_conv_kernel<false>"),
        ("lwe_lincomb_kernel", "lwe_lincomb_kernel"),
        ("bsk_to_fourier_kernel", "bsk_to_fourier_kernel (key conversion)"),
        ("blind_rotate_tm_kernelILi5ELi3E", "blind_rotate_tm_kernel<5,3> (variant 3: BSK through tensor memory; not the default)")]
txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []
    elif cur and re.search(r"/\*[0-9a-f]{4}\*/", line):
        ins = line.split("*/", 1)[1].strip().split(";")[0]
        funcs[cur].append(ins)
out = ["SASS instruction histograms, cuobjdump -sass redsec_b200/libredsec_b200.so (sm_100a), produced by scripts/sass_histogram.py",
       "Counts are STATIC instruction counts of the whole kernel (prologue, key/twiddle setup and all roles), not per-row dynamic counts.",
       "UBLKCP = cp.async.bulk (1-D TMA bulk copy), SYNCS = mbarrier operations, USETMAXREG = setmaxnreg, UTCCP = tcgen05.cp, LDTM = tcgen05.ld.", ""]
for key, title in WANT:
    for name, body in funcs.items():
        if key in name:
            c = collections.Counter()
            for ins in body:
                op = ins.split()[0] if not ins.startswith("@") else ins.split()[1]
                for k in KEYS:
                    if op.startswith(k):
                        c[k] += 1
            reuse = sum(1 for ins in body if ins.split()[0].startswith("DFMA") and ".reuse" in ins)
            out.append(f"== {title}\n   mangled: {name}\n   {len(body)} instructions; " + ", ".join(f"{k} {c[k]}" for k in KEYS if c[k]) +
                       f"; DFMA with a .reuse operand {reuse}")
            samples = [ins for ins in body if any(ins.split()[0].startswith(k) for k in ("UBLKCP", "USETMAXREG", "UTCCP", "LDTM", "UTCIMMA", "UTCBAR"))][:5]
            for s_ in samples:
                out.append("      " + s_)
            out.append("")
            break
    else:
        out.append(f"== {title}: NOT FOUND in the library\n")
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
open(os.path.join(ROOT, "profiles", "r2_sass_histograms.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
