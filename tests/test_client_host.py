"""CPU: host-side product code -- client tools (keygen / encrypt / decrypt / files), weight-file handling and the
neuron-partition arithmetic -- against the oracle."""
import os

import numpy as np

from redsec_b200 import client, netspec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_keygen_equals_oracle_spec(oracle, keyset):
    ks = client.keygen(0)
    assert np.array_equal(ks.lwe_key, keyset.lwe_key) and np.array_equal(ks.tlwe_key, keyset.tlwe_key)
    assert np.array_equal(ks.bsk, keyset.bsk) and np.array_equal(ks.ksk, keyset.ksk)
    other = client.keygen(1)
    assert not np.array_equal(other.lwe_key, ks.lwe_key)


def test_encrypt_decrypt_match_oracle_and_round_trip(oracle, keyset):
    rng = np.random.default_rng(0)
    msgs = rng.integers(-2047, 2048, 200)
    ct = client.encrypt(msgs * client.UNIT, keyset.lwe_key, client.SECALPHA, seed=3)
    assert np.array_equal(ct, oracle.encrypt((msgs * client.UNIT) & 0xFFFFFFFF, client.SECALPHA, keyset.lwe_key, 3))
    assert np.array_equal(client.decrypt(ct, keyset.lwe_key, 4096), msgs)
    assert np.array_equal(client.decrypt(ct, keyset.lwe_key, 4096), oracle.decrypt(ct, keyset.lwe_key, 4096))
    # edge: message space boundary 2048 decrypts to +2048 (centred to (-2048, 2048], client/decrypt_image.cpp:53-58)
    edge = client.encrypt(np.array([2048, -2048]) * client.UNIT, keyset.lwe_key, 0.0, seed=4)
    assert list(client.decrypt(edge, keyset.lwe_key, 4096)) == [2048, 2048]


def test_image_encoding(oracle, keyset):
    label, px = netspec.load_image_csv(netspec.NETS["mnist/sign1024x1"]()["image"])
    assert label == 0 and len(px) == 784
    ct = client.encrypt_image(px, keyset.lwe_key, seed=5)
    assert ct.shape == (784, 351)                     # every pixel encrypted (reference defect R1 not reproduced)
    assert np.array_equal(client.decrypt(ct, keyset.lwe_key, 4096), 2 * np.asarray(px) - 255)


def test_key_and_ciphertext_files_round_trip(tmp_path, keyset):
    ks = client.KeySet(keyset.lwe_key, keyset.tlwe_key, keyset.bsk, keyset.ksk)
    sk, ek = str(tmp_path / "secret.key"), str(tmp_path / "eval.key")
    client.write_keys(ks, sk, ek)
    back = client.read_keys(sk, ek)
    assert all(np.array_equal(getattr(back, f), getattr(ks, f)) for f in ("lwe_key", "tlwe_key", "bsk", "ksk"))
    ct = client.encrypt(np.arange(10) * client.UNIT, ks.lwe_key, client.SECALPHA, seed=6)
    p = str(tmp_path / "image.ctxt")
    client.write_ctxt(p, ct[:4]); client.write_ctxt(p, ct[4:], append=True)      # WriteCtxtToFileRed appends (main.cu:82)
    assert os.path.getsize(p) == 10 * (351 * 4 + 8)
    assert np.array_equal(client.read_ctxt(p, 10), ct)
    # a truncated or foreign key file is rejected
    open(ek, "r+b").write(b"XXXX")
    try:
        client.read_keys(None, ek)
        assert False, "corrupt eval.key accepted"
    except Exception:
        pass


def test_weight_writer_reader_round_trip(tmp_path):
    from oracle import layers_oracle as LO
    spec = netspec.tiny_cifar_like()
    spec["weights"] = netspec.write_random_weights(spec, str(tmp_path / "w.dat"), seed=1, p_zero=0.3)
    layers = LO.prepare(spec, spec["weights"])       # asserts the file is consumed exactly
    assert [l.out_dims for l in layers] == [(8, 8, 3), (8, 8, 16), (4, 4, 16), (1, 1, 32), (1, 1, 10)]
    w = layers[1].weights
    assert set(np.unique(w)) <= {-1, 0, 1} and 0.15 < np.mean(w == 0) < 0.45
    assert LO.count_bootstraps(layers) == 192 + 1024 + 1024 + 3 * 256 + 32


def test_bootstrap_counts_per_net():
    from oracle import layers_oracle as LO
    want = {"mnist/sign1024x1": 1220, "mnist/sign1024x2": 2244, "mnist/sign1024x3": 3268, "mnist/cnn_builder": 4832,
            "cifar/binarynet": 463872 + 172032, "cifar/binarynet_small": 232960 + 86528}
    for name, n in want.items():
        spec = netspec.NETS[name]()
        assert LO.count_bootstraps(LO.prepare(spec, spec["weights"])) == n, name


def test_shard_ranges_partition_channels():
    from redsec_b200 import nets
    for channels in (128, 1024, 16, 10, 3):
        for world in (1, 2, 4, 8):
            got = [nets.shard_range(channels, True, r, world) for r in range(world)]
            if channels % world == 0:
                assert got[0][0] == 0 and got[-1][1] == channels
                assert all(got[i][1] == got[i + 1][0] for i in range(world - 1))
            else:
                assert all(g == (0, channels) for g in got)      # not shardable -> replicated
    assert nets.shard_range(3, False, 1, 2) == (0, 3)             # no conv stage -> replicated


def test_chacha20_known_answer():
    """The CSPRNG behind rs_keygen_secure / rs_lwe_encrypt_secure is ChaCha20 (64-bit counter, 64-bit nonce): all-zero key and
    nonce give the published keystream 76 b8 e0 ad a0 f1 3d 90 40 5d 6a e5 53 86 bd 28 ... (Bernstein's test vector, also
    draft-agl-tls-chacha20poly1305 TC1); the second block continues it, and another stream index gives different words."""
    import ctypes as C
    from redsec_b200 import _lib
    lib = _lib.load()
    key = np.zeros(8, np.uint32)
    out = np.zeros(32, np.uint32)
    assert lib.rs_selftest_chacha20(key.ctypes.data, 0, 0, out.ctypes.data, 32) == 0
    want = bytes.fromhex("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7"
                         "da41597c5157488d7724e03fb8d84a376a43b8f41518a11cc387b669b2ee6586"
                         "9f07e7be5551387a98ba977c732d080dcb0f29a048e3656912c6533e32ee7aed"
                         "29b721769ce64e43d57133b074d839d531ed1f28510afb45ace10a1f4b794d6f")
    assert out.astype("<u4").tobytes() == want
    other = np.zeros(16, np.uint32)
    assert lib.rs_selftest_chacha20(key.ctypes.data, 5, 1, other.ctypes.data, 16) == 0
    assert not np.array_equal(other, out[:16])


def test_secure_keygen_and_encryption_are_fresh_every_call():
    """the default client paths draw from the OS (ChaCha20 keyed by getrandom): two key generations differ, two
    encryptions of the same message share neither masks nor noise, and both decrypt; explicit seeds stay reproducible."""
    from redsec_b200 import client
    k1, k2 = client.keygen(), client.keygen()
    assert not np.array_equal(k1.lwe_key, k2.lwe_key) and not np.array_equal(k1.bsk[:2048], k2.bsk[:2048])
    assert set(np.unique(k1.lwe_key)) <= {0, 1} and 100 < int(k1.lwe_key.sum()) < 250
    msg = np.arange(-8, 8) * client.UNIT
    c1, c2 = client.encrypt(msg, k1.lwe_key, client.SECALPHA), client.encrypt(msg, k1.lwe_key, client.SECALPHA)
    assert not np.array_equal(c1[:, :350], c2[:, :350])
    assert np.array_equal(client.decrypt(c1, k1.lwe_key), np.arange(-8, 8)) and np.array_equal(client.decrypt(c2, k1.lwe_key), np.arange(-8, 8))
    # the key material is usable: a key-switching key row decrypts (under the LWE key) to h * s'_i / base^(j+1)
    ks_row = k1.ksk.reshape(1024, 9, 8, 351)[3, 0, 5][None, :]
    ph = int(client.phase(ks_row, k1.lwe_key)[0])
    expect = (5 * int(k1.tlwe_key[3])) << 29
    assert abs(((ph - expect + 2 ** 31) % 2 ** 32) - 2 ** 31) < 2 ** 12
    s1, s2 = client.encrypt(msg, k1.lwe_key, client.SECALPHA, seed=9), client.encrypt(msg, k1.lwe_key, client.SECALPHA, seed=9)
    assert np.array_equal(s1, s2)


def test_client_executables_round_trip_through_files(tmp_path):
    """client/Makefile workflow (reference client/Makefile:1-25): keygen -> image_converter.py -> encrypt-image -> decrypt-image,
    files only.  The tools' files are the library's formats: image.ctxt written by encrypt.out equals client.encrypt_image
    with the same test seed and decrypts to 2p-255 for ALL 784 pixels (defect R1 not reproduced); decrypt.out recovers the
    argmax of the scores in network_output.ctxt."""
    import shutil
    import subprocess
    from redsec_b200 import client, netspec
    cdir = os.path.join(ROOT, "client")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call(["make", "-C", cdir, "all", f"CXX={gxx}"], stdout=subprocess.DEVNULL)
    run = lambda *a: subprocess.run(list(a), cwd=tmp_path, check=True, capture_output=True, text=True).stdout
    assert "deterministic TEST keyset" in run(os.path.join(cdir, "keygen.out"), "-seclevel", "128", "--seed", "0")
    ks = client.read_keys(str(tmp_path / "secret.key"), str(tmp_path / "eval.key"))
    ref = client.keygen(0)
    assert np.array_equal(ks.lwe_key, ref.lwe_key) and np.array_equal(ks.bsk, ref.bsk) and np.array_equal(ks.ksk, ref.ksk)
    run("python3", os.path.join(cdir, "image_converter.py"), "--format", "MNIST", "--image", os.path.join(ROOT, "data", "client", "mnist_test.csv"))
    assert (tmp_path / "image.ptxt").read_text().startswith("0,28,28,1,")
    run(os.path.join(cdir, "encrypt.out"), "image.ptxt", "--seed", "11")
    label, px = netspec.load_image_csv(os.path.join(ROOT, "data", "client", "mnist_test.csv"))
    ct = client.read_ctxt(str(tmp_path / "image.ctxt"), 784)
    assert np.array_equal(ct, client.encrypt_image(px, ks.lwe_key, seed=11))
    assert np.array_equal(client.decrypt(ct, ks.lwe_key), 2 * np.asarray(px) - 255)
    run(os.path.join(cdir, "encrypt.out"), "image.ptxt")                                   # OS entropy: a different ciphertext, same pixels
    ct2 = client.read_ctxt(str(tmp_path / "image.ctxt"), 784)
    assert not np.array_equal(ct2, ct) and np.array_equal(client.decrypt(ct2, ks.lwe_key), 2 * np.asarray(px) - 255)
    scores = np.array([-143, 12, 300, -7, 299, 0, -2048 + 1, 2047, 5, -1])
    client.write_ctxt(str(tmp_path / "network_output.ctxt"), client.encrypt(scores * client.UNIT, ks.lwe_key, 2.0 ** -25, seed=3))
    out = run(os.path.join(cdir, "decrypt.out"), "MNIST")
    assert "Classification Result: 7" in out and "Scores: -143 12 300 -7 299 0 -2047 2047 5 -1" in out
    for f in ("keygen.out", "encrypt.out", "decrypt.out"):
        os.remove(os.path.join(cdir, f))
