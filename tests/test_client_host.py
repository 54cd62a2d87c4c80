"""CPU: host-side product code -- client tools (keygen / encrypt / decrypt / files), weight-file handling and the
neuron-partition arithmetic -- against the oracle."""
import os

import numpy as np

from redsec_b200 import client, netspec


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
