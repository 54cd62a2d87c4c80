"""GPU parity of the bootstrap hot path against the CPU oracle, through the C-ABI.
Bar (north star): bit-exact ciphertexts for keyswitch / sample extract / full PBS on identical inputs."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MU8 = 0x20000000       # 1/8
MU4096 = 0x00100000    # 1/4096


def _rand_bits_ct(O, ks, count, mu, alpha, seed):
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 2, count)
    msg = np.where(bits == 1, mu, (-mu) & 0xFFFFFFFF)
    return bits, O.encrypt(msg, alpha, ks.lwe_key, seed)


def test_blind_rotate_extract_bit_exact(oracle, keyset, engine):
    O = oracle
    bits, ct = _rand_bits_ct(O, keyset, 13, MU8, 2.0 ** -25, 11)
    dev = engine.upload(ct)
    ext_gpu = engine.blind_rotate(dev, MU8)
    for c in range(ct.shape[0]):
        acc = O.blind_rotate(ct[c], MU8, keyset, exact=(c < 2))
        assert np.array_equal(ext_gpu[c], O.sample_extract(acc)), f"ciphertext {c}"


@pytest.mark.parametrize("variant", [3, 2], ids=["smem-gather", "tensor-cores"])
@pytest.mark.parametrize("count", [9, 257, 1500])
def test_keyswitch_bit_exact(oracle, keyset, engine, count, variant):
    """Both keyswitch kernels against the oracle on random extracted samples: the shared-memory gather kernel (TILE 16 / 32 paths,
    split sum with atomics) and the exact int8 GEMM on the tensor cores (tcgen05.mma kind::i8; ragged last M tile)."""
    rng = np.random.default_rng(5 + count)
    ext = rng.integers(0, 2 ** 32, size=(count, 1025), dtype=np.uint64).astype(np.uint32)
    ext[0, :1024] = 0                       # all digits 0: only (0, b) survives
    ext[1, :1024] = 0xFFFFFFFF              # carries through every digit
    engine.set_ks_variant(variant)
    try:
        got = engine.keyswitch(ext)
    finally:
        engine.set_ks_variant(0)
    want = oracle.keyswitch(ext, keyset)
    assert np.array_equal(got, want)


def test_keyswitch_variants_agree_at_scale(engine, keyset):
    """20 001 ciphertexts (ragged last tile in every kernel): the default (auto: tensor cores at this size), the shared-memory
    gather kernel (TILE 64 path) and the un-tiled gather kernel give the same words (all three are bit-exact vs the oracle on
    the small cases above; integer sums, so any difference is a bug)."""
    rng = np.random.default_rng(9)
    ext = rng.integers(0, 2 ** 32, size=(20001, 1025), dtype=np.uint64).astype(np.uint32)
    got = engine.keyswitch(ext)
    try:
        for v in (1, 3):
            engine.set_ks_variant(v)
            assert np.array_equal(engine.keyswitch(ext), got), f"keyswitch variant {v}"
    finally:
        engine.set_ks_variant(0)


@pytest.mark.parametrize("count", [1, 5, 6, 7, 200, 300])
def test_pbs_bit_exact_and_decrypts(oracle, keyset, engine, count):
    O = oracle
    # inputs +-200/4096: far enough from 0 that the modswitch noise (sigma ~ 7.7/4096) cannot flip the sign
    bits, ct = _rand_bits_ct(O, keyset, count, 200 * MU4096, 2.0 ** -15, 100 + count)
    got = engine.download(engine.pbs(engine.upload(ct), MU4096))
    want = O.pbs(ct, MU4096, keyset)
    assert np.array_equal(got, want)
    dec = O.decrypt(got, keyset.lwe_key, 4096)
    assert np.array_equal(dec, np.where(bits == 1, 1, -1))


def test_pbs_random_and_degenerate_ciphertexts_bit_exact(oracle, keyset, engine):
    """3 000 arbitrary LWE rows (uniform random words, not encryptions: every mod-switched value occurs, including ~500
    coefficients that round to 0 and skip their blind-rotate step) plus degenerate rows: trivial samples (a = 0) at phases 0,
    1/2, +-3/4096 and -1/4096 (exactly half a slot below 0: rounds to slot 0, so +mu), an all-ones mask, and a mask whose every
    coefficient rounds to 2N-1.  Bit-exact vs the oracle."""
    rng = np.random.default_rng(77)
    ct = rng.integers(0, 2 ** 32, size=(3000, 351), dtype=np.uint64).astype(np.uint32)
    special = np.zeros((7, 351), np.uint32)
    special[0, 350] = 0                       # trivial, phase exactly 0 -> +mu
    special[1, 350] = 0x80000000              # trivial, phase 1/2
    special[2, 350] = 3 * MU4096
    special[3, 350] = (-3 * MU4096) & 0xFFFFFFFF
    special[4, 350] = (-MU4096) & 0xFFFFFFFF  # -1/4096 = -half a slot: modSwitch rounds it to slot 0
    special[5, :] = 0xFFFFFFFF                # every coefficient rounds to 0 mod 2N (wraps)
    special[6, :350] = 0xFFE00000 - (1 << 20)  # rounds to 2N-1
    ct = np.concatenate([special, ct])
    got = engine.download(engine.pbs(engine.upload(ct), MU4096))
    want = oracle.pbs(ct, MU4096, keyset)
    assert np.array_equal(got, want)
    dec = oracle.decrypt(got[:5], keyset.lwe_key, 4096)
    assert list(dec) == [1, -1, 1, -1, 1]


def test_unbinarize_mu_and_empty_batch(oracle, keyset, engine):
    """BinOps::unbinarize_int is the same bootstrap with mu = 1/2048 (MULTIBIT_SPACE, lib/BinOps_enc.cpp:188-192, lib/Layer.h:35);
    an empty batch is a no-op; in == out aliasing is allowed."""
    mu = oracle.to_torus(1, 2048)
    bits, ct = _rand_bits_ct(oracle, keyset, 11, 300 * MU4096, 2.0 ** -15, 55)
    dev = engine.upload(ct)
    got = engine.download(engine.pbs(dev, mu, out=dev))            # in place
    assert np.array_equal(got, oracle.pbs(ct, mu, keyset))
    assert np.array_equal(oracle.decrypt(got, keyset.lwe_key, 2048), np.where(bits == 1, 1, -1))
    assert engine.lib.rs_pbs_batch(engine.ctx, dev.ptr, dev.ptr, 0, mu) == 0
    assert engine.lib.rs_gate_batch(engine.ctx, 0, dev.ptr, dev.ptr, dev.ptr, 0, mu) == 0
    assert np.array_equal(engine.download(dev), got)                # untouched by the empty calls


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_test_vector_bootstrap_bit_exact(oracle, keyset, engine, variant):
    """rs_pbs_lut_batch (row f4): per-ciphertext test vectors, ciphertext c uses table c % m.  Bit-exact vs the oracle and the
    output decrypts to the table entry of the slot the rotation ends on, on both halves of the torus."""
    from oracle import layers_oracle as LO
    rng = np.random.default_rng(31)
    luts = (rng.integers(-2000, 2000, size=(5, 1024)) * LO.UNIT & 0xFFFFFFFF).astype(np.uint32)
    msgs = (rng.integers(-2048, 2048, size=37) * LO.UNIT & 0xFFFFFFFF).astype(np.uint32)
    ct = oracle.encrypt(msgs, 2.0 ** -15, keyset.lwe_key, 6)
    engine.set_tuning(variant)
    try:
        got = engine.download(engine.pbs_lut(engine.upload(ct), luts))
    finally:
        engine.set_tuning(0)
    assert np.array_equal(got, oracle.pbs_lut(ct, luts, keyset))
    want, slot = LO.predicted_lut_message(ct, luts, keyset.lwe_key)
    err = (oracle.phase(got, keyset.lwe_key).astype(np.int64) - want.astype(np.int64) + 2 ** 31) % 2 ** 32 - 2 ** 31
    assert np.max(np.abs(err)) < LO.UNIT // 4 and len(set(slot // 1024)) == 2


def test_test_vector_bootstrap_rejected_by_tensor_memory_variant(engine, oracle, keyset):
    import redsec_b200 as rs
    ct = oracle.encrypt(np.zeros(4, np.uint32), 2.0 ** -15, keyset.lwe_key, 1)
    engine.set_tuning(3)
    try:
        with pytest.raises(rs.RsError):
            engine.pbs_lut(engine.upload(ct), np.zeros((1, 1024), np.uint32))
    finally:
        engine.set_tuning(0)


@pytest.mark.parametrize("groups", [0, 1, 2, 3, 4])
def test_variants_agree(oracle, keyset, engine, groups):
    bits, ct = _rand_bits_ct(oracle, keyset, 29, MU8, 2.0 ** -25, 77)
    engine.set_tuning(groups)
    try:
        got = engine.download(engine.pbs(engine.upload(ct), MU8))
    finally:
        engine.set_tuning(0)
    assert np.array_equal(got, oracle.pbs(ct, MU8, keyset))


@pytest.mark.parametrize("op,fn", [("NAND", lambda a, b: 1 - (a & b)), ("OR", lambda a, b: a | b),
                                    ("AND", lambda a, b: a & b), ("NOR", lambda a, b: 1 - (a | b)),
                                    ("XOR", lambda a, b: a ^ b), ("XNOR", lambda a, b: 1 - (a ^ b))])
def test_gates_bit_exact(oracle, keyset, engine, op, fn):
    O = oracle
    a_bits, a = _rand_bits_ct(O, keyset, 40, MU8, 2.0 ** -25, 21)
    b_bits, b = _rand_bits_ct(O, keyset, 40, MU8, 2.0 ** -25, 22)
    got = engine.download(engine.gate(op, engine.upload(a), engine.upload(b), MU8))
    want = O.gate(op, a, b, MU8, keyset)
    assert np.array_equal(got, want)
    dec = (O.decrypt(got, keyset.lwe_key, 8) > 0).astype(int)
    assert np.array_equal(dec, fn(a_bits, b_bits))


def test_host_entry_points_match_device(oracle, keyset, engine):
    bits, ct = _rand_bits_ct(oracle, keyset, 17, MU8, 2.0 ** -25, 31)
    got = engine.pbs_host(ct, MU8)
    assert np.array_equal(got, oracle.pbs(ct, MU8, keyset))
    got2 = engine.gate_host("NAND", ct, ct[::-1].copy(), MU8)
    assert np.array_equal(got2, oracle.gate("NAND", ct, ct[::-1].copy(), MU8, keyset))


def test_2pow16_gates_truth_table_and_oracle_sample(oracle, keyset, engine):
    """BASELINE config 2 at full size: 2^16 independent NAND and XNOR gates in one launch each; EVERY gate decrypts to the
    truth table, and a 4 096-gate sample (a contiguous block, a strided set and the batch edges) is ciphertext-equal to
    the CPU oracle."""
    n = 1 << 16
    rng = np.random.default_rng(1)
    a_bits, b_bits = rng.integers(0, 2, n), rng.integers(0, 2, n)
    mu8 = 1 << 29
    a = oracle.encrypt(np.where(a_bits == 1, mu8, -mu8), 2.0 ** -25, keyset.lwe_key, 101)
    b = oracle.encrypt(np.where(b_bits == 1, mu8, -mu8), 2.0 ** -25, keyset.lwe_key, 102)
    d_a, d_b = engine.upload(a), engine.upload(b)
    sample = np.unique(np.concatenate([np.arange(0, 2048), np.arange(2048, n, 31)[:2040], [n - 1, n - 2, 591, 592, 593]]))
    for op, truth in (("NAND", 1 - (a_bits & b_bits)), ("XNOR", 1 - (a_bits ^ b_bits))):
        got = engine.download(engine.gate(op, d_a, d_b, mu8))
        dec = (oracle.phase(got, keyset.lwe_key).astype(np.int32) > 0).astype(np.int64)
        assert np.array_equal(dec, truth), f"{op}: {int((dec != truth).sum())} of 2^16 gates decrypt wrong"
        want = oracle.gate(op, a[sample], b[sample], mu8, keyset)
        assert np.array_equal(got[sample], want), f"{op}: sampled gates differ from the oracle"


def test_row_split_under_a_lagging_back_warp_pair(oracle, keyset):
    """ADVICE r1 (BSK ring parity aliasing in row-split mode): the stress instantiation delays one back-warp pair of every CTA by
    about a row per row, so front warps of the other slots run their slab claims as far ahead as the rings allow.  With the
    in-order slab issue the ciphertexts still equal the oracle's for batch sizes that use the split kernel (<= 2 per SM)."""
    import redsec_b200 as rs
    os.environ["RS_WS_STRESS"] = "1"
    try:
        eng = rs.Engine(0)
    finally:
        del os.environ["RS_WS_STRESS"]
    eng.load_eval_key(keyset.bsk, keyset.ksk)
    rng = np.random.default_rng(12)
    for count in (3, 150, 296):
        mu = rng.integers(-1500, 1500, count) * (1 << 20)
        ct = oracle.encrypt(mu, 2.0 ** -15, keyset.lwe_key, 300 + count)
        got = eng.download(eng.pbs(eng.upload(ct), 1 << 20))
        k = min(count, 24)
        idx = np.unique(np.concatenate([rng.choice(count, k, replace=False), [0, count - 1]]))
        assert np.array_equal(got[idx], oracle.pbs(ct[idx], 1 << 20, keyset)), f"count {count}"
        # the un-stressed engine gives the same batch
    eng.close()


def test_claim_protocol_build_agrees_with_the_producer_warp_build(oracle, keyset, engine):
    """The default blind rotation has a dedicated BSK producer warp (16-warp build); RS_WS_PRODUCER=0 selects the 12-warp build
    whose front warps claim the slabs.  Same ciphertexts from both, in every row-split mode, and both equal the oracle."""
    import redsec_b200 as rs
    os.environ["RS_WS_PRODUCER"] = "0"
    try:
        legacy = rs.Engine(0)
    finally:
        del os.environ["RS_WS_PRODUCER"]
    legacy.load_eval_key(keyset.bsk, keyset.ksk)
    rng = np.random.default_rng(77)
    for count in (2, 149, 300, 601):          # split 4, split 2, un-split partial wave, un-split with a second wave
        mu = rng.integers(-1500, 1500, count) * (1 << 20)
        ct = oracle.encrypt(mu, 2.0 ** -15, keyset.lwe_key, 500 + count)
        a = engine.download(engine.pbs(engine.upload(ct), 1 << 20))
        b = legacy.download(legacy.pbs(legacy.upload(ct), 1 << 20))
        assert np.array_equal(a, b), f"count {count}"
        idx = np.unique(np.concatenate([rng.choice(count, min(count, 12), replace=False), [0, count - 1]]))
        assert np.array_equal(a[idx], oracle.pbs(ct[idx], 1 << 20, keyset)), f"count {count}"
    legacy.close()


def test_producer_knobs_do_not_change_ciphertexts(oracle, keyset, engine):
    """RS_WS_GATE (every n-th wave of a long launch waits for the earlier waves: bounded spin on a counter the front warps bump) and
    RS_WS_LOOKAHEAD (slabs the producer warp requests ahead of the slowest consumer) only move work in time."""
    import redsec_b200 as rs
    rng = np.random.default_rng(78)
    count = 1300                                   # 325 CTAs = 2.2 waves: the second and third wave are gated
    mu = rng.integers(-1500, 1500, count) * (1 << 20)
    ct = oracle.encrypt(mu, 2.0 ** -15, keyset.lwe_key, 900)
    want = engine.download(engine.pbs(engine.upload(ct), 1 << 20))
    idx = np.unique(np.concatenate([rng.choice(count, 16, replace=False), [0, 591, 592, count - 1]]))
    assert np.array_equal(want[idx], oracle.pbs(ct[idx], 1 << 20, keyset))
    for env in ({"RS_WS_GATE": "1"}, {"RS_WS_GATE": "2", "RS_WS_LOOKAHEAD": "2"}, {"RS_WS_LOOKAHEAD": "1"}):
        os.environ.update(env)
        try:
            eng = rs.Engine(0)
        finally:
            for k in env:
                del os.environ[k]
        eng.load_eval_key(keyset.bsk, keyset.ksk)
        got = eng.download(eng.pbs(eng.upload(ct), 1 << 20))
        eng.close()
        assert np.array_equal(got, want), env


def test_small_last_wave_runs_row_split(oracle, keyset, engine):
    """A multi-wave batch whose last wave holds <= 2 ciphertexts per SM is cut into an un-split launch and a row-split one
    (api.cu launch_blind_rotate); test-vector batches only when the cut falls on a multiple of the table count.  Ciphertexts
    equal the oracle's on both sides of the cut, and equal the single-launch result (RS_WS_TAIL_SPLIT=0)."""
    import redsec_b200 as rs
    os.environ["RS_WS_TAIL_SPLIT"] = "0"
    try:
        whole = rs.Engine(0)
    finally:
        del os.environ["RS_WS_TAIL_SPLIT"]
    whole.load_eval_key(keyset.bsk, keyset.ksk)
    rng = np.random.default_rng(79)
    for count in (592 + 100, 2 * 592 + 280):
        mu = rng.integers(-1500, 1500, count) * (1 << 20)
        ct = oracle.encrypt(mu, 2.0 ** -15, keyset.lwe_key, 700 + count)
        got = engine.download(engine.pbs(engine.upload(ct), 1 << 20))
        assert np.array_equal(got, whole.download(whole.pbs(whole.upload(ct), 1 << 20))), count
        head = count - count % 592
        idx = np.unique(np.concatenate([rng.choice(count, 10, replace=False), [0, head - 1, head, head + 1, count - 1]]))
        assert np.array_equal(got[idx], oracle.pbs(ct[idx], 1 << 20, keyset)), count
    count = 592 + 60
    ct = oracle.encrypt(rng.integers(-1500, 1500, count) * (1 << 20), 2.0 ** -15, keyset.lwe_key, 990)
    for tables in (4, 3):                     # 592 % 4 == 0: cut; 592 % 3 != 0: one launch
        luts = (rng.integers(-1000, 1000, size=(tables, 1024)) * (1 << 20) & 0xFFFFFFFF).astype(np.uint32)
        got = engine.download(engine.pbs_lut(engine.upload(ct), luts))
        idx = np.array([0, 1, 2, 3, 590, 591, 592, 593, 594, 595, count - 1])
        want = np.concatenate([oracle.pbs_lut(ct[i:i + 1], luts[(i % tables):(i % tables) + 1], keyset) for i in idx])
        assert np.array_equal(got[idx], want), tables
    whole.close()


def test_second_key_load_rebuilds_every_keyswitch_table(oracle, keyset):
    """rs_load_eval_key twice on one context: the keyswitch tables derived from the key -- the tiled table of the gather kernel and
    the byte-limb table of the tensor-core kernel, which is built lazily on first use -- must follow the second key."""
    import redsec_b200 as rs
    rng = np.random.default_rng(80)
    ksk2 = rng.integers(0, 2 ** 32, size=keyset.ksk.size, dtype=np.uint64).astype(np.uint32)     # any table is a valid keyswitch key
    second = oracle.KeySet(keyset.lwe_key, keyset.tlwe_key, keyset.bsk, ksk2)
    ext = rng.integers(0, 2 ** 32, size=(600, 1025), dtype=np.uint64).astype(np.uint32)
    eng = rs.Engine(0)
    try:
        for ks in (keyset, second):
            eng.load_eval_key(ks.bsk, ks.ksk)
            want = oracle.keyswitch(ext, ks)
            for variant in (2, 3):                 # tensor cores, shared-memory gather
                eng.set_ks_variant(variant)
                assert np.array_equal(eng.keyswitch(ext), want), f"variant {variant}"
    finally:
        eng.close()
