"""CPU: the claims DESIGN.md makes about the compiled kernels, checked on the in-tree library with cuobjdump (no GPU needed):
TMA bulk copies, mbarriers, setmaxnreg, tcgen05 in the tensor-core keyswitch, and L2-only loads of ciphertext data."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "redsec_b200", "libredsec_b200.so")


@pytest.fixture(scope="module")
def sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump is not installed")
    if not os.path.exists(SO):
        pytest.skip("libredsec_b200.so is not built")
    txt = subprocess.run([exe, "-sass", SO], capture_output=True, text=True, check=True).stdout
    funcs, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), collections.Counter())
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    return funcs


def _one(funcs, fragment):
    hits = [c for name, c in funcs.items() if fragment in name]
    assert len(hits) == 1, f"{fragment}: {len(hits)} kernels match"
    return hits[0]


def _count(c, prefix):
    return sum(v for k, v in c.items() if k.startswith(prefix))


def test_default_blind_rotation_is_tma_fed_and_warp_specialised(sass):
    c = _one(sass, "blind_rotate_ws_kernelILi5ELi3ELi1ELb0ELb1E")       # <5 stages, 3 slots, un-split, no stress, producer warp>
    assert _count(c, "UBLKCP") >= 1                      # cp.async.bulk: the BSK slabs
    assert _count(c, "USETMAXREG") == 3                  # producer 24 / front 120 / back 184
    assert _count(c, "SYNCS") >= 40                      # mbarrier rings
    assert _count(c, "DFMA") + _count(c, "DADD") + _count(c, "DMUL") > 1000
    assert _count(c, "LDG") > 0 and all("CONSTANT" not in k for k in c if k.startswith("LDG"))      # no non-coherent ciphertext loads


def test_tensor_core_keyswitch_uses_tcgen05(sass):
    c = _one(sass, "keyswitch_mma_kernel")
    assert _count(c, "UTCIMMA") >= 4                     # tcgen05.mma kind::i8: two K-steps x two M tiles per stage
    assert _count(c, "LDTM") >= 1                        # tcgen05.ld epilogue
    assert _count(c, "UTCBAR") >= 3                      # tcgen05.commit onto the rings' barriers
    assert _count(c, "UBLKCP") >= 1                      # the key tiles arrive by TMA bulk copy
    assert all("CONSTANT" not in k for k in c if k.startswith("LDG"))


def test_kernels_that_read_ciphertexts_do_it_through_l2(sass):
    """data written by an earlier kernel is never read through the non-coherent L1 path."""
    for fragment in ("gate_linear_kernel", "lwe_axpby_kernel", "lwe_pad_kernel", "lwe_unpad_kernel", "lwe_interleave_kernel",
                     "keyswitch_tiled_kernelILi64E", "keyswitch_init_kernel", "ks_mma_prep_kernel", "kskb_build_kernel", "lwe_add_const_kernel"):
        c = _one(sass, fragment)
        loads = {k: v for k, v in c.items() if k.startswith("LDG")}
        assert loads and all("STRONG.GPU" in k for k in loads), (fragment, loads)
    for fragment in ("lwe_conv_kernelILb0E", "lwe_lincomb_kernel"):           # weights / CSR tables may use the read-only path, rows may not
        c = _one(sass, fragment)
        assert any("STRONG.GPU" in k for k in c if k.startswith("LDG")), fragment
