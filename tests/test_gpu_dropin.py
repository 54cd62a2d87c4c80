"""Drop-in check (SURVEY 8b "B-outer", rows f3 / b-ops): the reference's own, UNMODIFIED GPU drivers (nets/*/main.cu + net.cu,
built by dropin/build.sh with the reference Makefile's command lines against this repo's REDcuFHE facade) run on the GPU box
and must write the same output ciphertexts as the engine's own EncryptedNet on the same keyset and image ciphertexts.

Trees: `tree` = Layer-level facade (this repo's IntLayer / BinLayer); `tree_func` = the REFERENCE's own
lib/GPU/{Bin,Int}Layer.cu, unmodified, sequencing this repo's BinFunc:: / IntFunc:: classes (a network composed from Func
objects); `*_g<N>` = the same built with NUM_GPUS = N: one process, one host thread and one engine context per GPU, layers
neuron-sharded with an NCCL all-gather between them -- the output must equal the single-GPU ciphertexts bit for bit."""
import os
import subprocess

import numpy as np
import pytest

from redsec_b200 import netspec

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "dropin", "_build")

CASES = [("tree", "mnist/sign1024x1", 1), ("tree", "mnist/sign1024x3", 1), ("tree", "cifar/binarynet_small", 1),
         ("tree_func", "mnist/sign1024x1", 1), ("tree_func", "cifar/binarynet_small", 1),
         ("tree_g2", "mnist/sign1024x1", 2), ("tree_g2", "cifar/binarynet_small", 2), ("tree_func_g2", "cifar/binarynet_small", 2),
         ("tree_g8", "cifar/binarynet", 8)]


def _gpu_count() -> int:
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("tree,name,ngpu", CASES, ids=[f"{t}-{n}" for t, n, _ in CASES])
def test_reference_gpu_driver_runs_on_the_engine(engine, keyset, tree, name, ngpu):
    from redsec_b200 import client, nets
    exe = os.path.join(BUILD, tree, "nets", name, "gpu-encrypt.out")
    if not os.path.exists(exe):
        pytest.skip("dropin/_build is not built (needs the reference tree at build time: dropin/build.sh)")
    if _gpu_count() < ngpu:
        pytest.skip(f"needs {ngpu} GPUs on the box, {_gpu_count()} present")
    cdir = os.path.join(BUILD, tree, "client")
    os.makedirs(cdir, exist_ok=True)
    ks = client.KeySet(keyset.lwe_key, keyset.tlwe_key, keyset.bsk, keyset.ksk)
    eval_key = os.path.join(cdir, "eval.key")
    if not os.path.exists(eval_key):
        client.write_keys(ks, os.path.join(cdir, "secret.key"), eval_key)
    spec = netspec.NETS[name]()
    label, px = netspec.load_image_csv(spec["image"])
    ct = client.encrypt_image(px, ks.lwe_key, seed=11)
    client.write_ctxt(os.path.join(cdir, "image.ctxt"), ct, variance=2.0 ** -30)
    out_path = os.path.join(cdir, "network_output.ctxt")
    if os.path.exists(out_path):
        os.remove(out_path)                      # the driver appends (main.cu:82)
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.pathsep.join([os.path.join(BUILD, "lib"), os.path.join(ROOT, "redsec_b200"),
                                              env.get("LD_LIBRARY_PATH", "")])
    r = subprocess.run([exe], cwd=os.path.dirname(exe), env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Inference Time" in r.stdout
    got = client.read_ctxt(out_path, 10)
    net = nets.EncryptedNet(engine, spec)
    want = engine.download(net.run(engine.upload(ct)))
    net.close()
    assert np.array_equal(got, want), "reference driver over the facade and EncryptedNet disagree"
    scores = client.decrypt(got, ks.lwe_key, 4096)
    print(tree, name, "scores", scores.tolist(), "label", label, r.stdout.strip().splitlines()[-1])
    if name.startswith("mnist"):
        assert int(np.argmax(scores)) == label


def test_ops_level_surface_matches_the_oracle(oracle, keyset):
    """Row b-ops: one call of each kind of the gates.cuh / BinOps:: / IntOps:: surface (count-1 calls on the batch API) through a
    C++ caller linked like a net driver; every result ciphertext must equal the oracle's."""
    from redsec_b200 import client
    exe = os.path.join(BUILD, "tree", "opscheck", "ops_check.out")
    if not os.path.exists(exe):
        pytest.skip("dropin/_build is not built")
    O = oracle
    cdir = os.path.join(BUILD, "tree", "client")
    os.makedirs(cdir, exist_ok=True)
    ks = client.KeySet(keyset.lwe_key, keyset.tlwe_key, keyset.bsk, keyset.ksk)
    if not os.path.exists(os.path.join(cdir, "eval.key")):
        client.write_keys(ks, os.path.join(cdir, "secret.key"), os.path.join(cdir, "eval.key"))
    E, U = 1 << 29, 1 << 20
    bits = [1, 0, 1]
    gate_ct = O.encrypt([E if v else -E for v in bits], 2.0 ** -25, keyset.lwe_key, 71)
    int_ct = O.encrypt([37 * U, -12 * U], 2.0 ** -25, keyset.lwe_key, 72)
    client.write_ctxt(os.path.join(cdir, "ops_in.ctxt"), np.concatenate([gate_ct, int_ct]), variance=2.0 ** -50)
    out_path = os.path.join(cdir, "ops_out.ctxt")
    if os.path.exists(out_path):
        os.remove(out_path)
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.pathsep.join([os.path.join(BUILD, "lib"), os.path.join(ROOT, "redsec_b200"), env.get("LD_LIBRARY_PATH", "")])
    r = subprocess.run([exe], cwd=os.path.dirname(exe), env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = client.read_ctxt(out_path, 18)
    a, b, c = gate_ct[0:1], gate_ct[1:2], gate_ct[2:3]
    x, y = int_ct[0:1], int_ct[1:2]
    g = lambda op, p, q: O.gate(op, p, q, E, keyset)
    triv = lambda v: np.concatenate([np.zeros(350, np.uint32), np.array([v & 0xFFFFFFFF], np.uint32)])[None, :]
    t0 = g("XOR", a, b)
    s = g("XOR", t0, c)
    carry = g("OR", g("AND", c, t0), g("AND", a, b))
    want = [g("NAND", a, b), g("OR", a, b), g("AND", a, b), g("NOR", a, b), g("XOR", a, b), g("XNOR", a, b),
            x + y, x - y, x * np.uint32(3), np.uint32(0) - x,
            O.pbs(x + y, U, keyset), O.pbs(x, 2 * U, keyset), s, carry, g("OR", a, b), np.uint32(0) - a, x - y, x + triv(17 * U)]
    for i, w in enumerate(want):
        assert np.array_equal(got[i], np.asarray(w, dtype=np.uint32).reshape(-1)), f"ops_check result {i}"
    dec = (O.phase(got[:6], keyset.lwe_key).astype(np.int32) > 0).astype(int).tolist()
    assert dec == [0 if (bits[0] & bits[1]) else 1, bits[0] | bits[1], bits[0] & bits[1], 1 - (bits[0] | bits[1]), bits[0] ^ bits[1], 1 - (bits[0] ^ bits[1])]
