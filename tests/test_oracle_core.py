"""CPU: pins the oracle (tests/golden/oracle_kat.json, made by tests/golden/make_oracle_kat.py) and checks its
internal consistency: exact-integer and FFT blind rotations agree bit for bit, gates obey their truth tables,
keyswitch/sample-extract are consistent with decryption under the corresponding keys."""
import hashlib
import json
import os

import numpy as np
import pytest

MU8 = 1 << 29
KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_kat.json")))


def _h(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_modswitch_known_answers(oracle):
    O = oracle
    assert O.to_torus(1, 8) == 0x20000000 and O.to_torus(-1, 8) == 0xE0000000
    assert O.to_torus(1, 4096) == 1 << 20 and O.to_torus(-5, 4096) == (-5 << 20) & 0xFFFFFFFF
    # modSwitchFromTorus32(x, 2N) == ((x<<32)+2^52)>>53 (lib/GPU/gates.cu:39-42)
    for x in (0, 1, (1 << 20) - 1, 1 << 20, 1 << 21, 0x7FFFFFFF, 0x80000000, 0xFFEFFFFF, 0xFFF00000, 0xFFFFFFFF):
        assert O.from_torus(x, 2048) == ((((x << 32) + (1 << 52)) & ((1 << 64) - 1)) >> 53)


def test_keyset_and_pbs_known_answers(oracle, keyset):
    O = oracle
    assert _h(keyset.lwe_key) == KAT["lwe_key"] and _h(keyset.tlwe_key) == KAT["tlwe_key"]
    assert _h(keyset.bsk) == KAT["bsk"] and _h(keyset.ksk) == KAT["ksk"]
    ct = O.encrypt(np.array([MU8, (-MU8) & 0xFFFFFFFF, MU8]), 2.0 ** -25, keyset.lwe_key, 5)
    assert _h(ct) == KAT["ct"]
    assert _h(O.pbs(ct, MU8, keyset)) == KAT["pbs_fft"]


def test_exact_and_fft_blind_rotation_agree(oracle, keyset):
    O = oracle
    ct = O.encrypt(np.array([MU8, (-MU8) & 0xFFFFFFFF]), 2.0 ** -25, keyset.lwe_key, 6)
    stats = np.zeros(3)
    for c in range(2):
        exact = O.blind_rotate(ct[c], MU8, keyset, exact=True)
        fft = O.blind_rotate(ct[c], MU8, keyset, exact=False, stats=stats)
        assert np.array_equal(exact, fft)
    # phase-error bound of the FFT path: pre-rounding |x - rint(x)| (must stay < 0.5, observed ~1e-3)
    assert stats[0] < 0.05, stats


def test_sample_extract_and_keyswitch_semantics(oracle, keyset):
    O = oracle
    ct = O.encrypt(np.array([MU8]), 2.0 ** -25, keyset.lwe_key, 7)
    acc = O.blind_rotate(ct[0], MU8, keyset, exact=False)
    ext = O.sample_extract(acc)
    # phase of the extracted sample under the TLWE key equals coefficient 0 of the accumulator's phase: ~ +1/8
    s = keyset.tlwe_key.astype(np.uint64)
    phase = (int(ext[1024]) - int(np.sum(ext[:1024].astype(np.uint64) * s) % (1 << 32))) % (1 << 32)
    assert abs(phase - MU8) < (1 << 24)
    out = O.keyswitch(ext, keyset)
    ph = int(O.phase(out, keyset.lwe_key)[0])
    assert abs(ph - MU8) < (1 << 25)


@pytest.mark.parametrize("op,fn", [("NAND", lambda a, b: 1 - (a & b)), ("OR", lambda a, b: a | b), ("AND", lambda a, b: a & b),
                                    ("NOR", lambda a, b: 1 - (a | b)), ("XOR", lambda a, b: a ^ b), ("XNOR", lambda a, b: 1 - (a ^ b))])
def test_gate_truth_tables(oracle, keyset, op, fn):
    O = oracle
    a_bits = np.array([0, 0, 1, 1, 0, 1, 1, 0])
    b_bits = np.array([0, 1, 0, 1, 1, 1, 0, 0])
    enc = lambda bits, seed: O.encrypt(np.where(bits == 1, MU8, (-MU8) & 0xFFFFFFFF), 2.0 ** -25, keyset.lwe_key, seed)
    out = O.gate(op, enc(a_bits, 1), enc(b_bits, 2), MU8, keyset)
    assert np.array_equal((O.decrypt(out, keyset.lwe_key, 8) > 0).astype(int), fn(a_bits, b_bits))


def test_sign_bootstrap_maps_zero_phase_to_plus_mu(oracle, keyset):
    # tfhe_bootstrap_FFT maps phase exactly 0 to +mu: trivial sample (0, 0)
    O = oracle
    ct = np.zeros((1, 351), np.uint32)
    out = O.pbs(ct, 1 << 20, keyset)
    assert O.decrypt(out, keyset.lwe_key, 4096)[0] == 1


def test_test_vector_bootstrap_exact_fft_and_slot_semantics(oracle, keyset):
    """Row f4: programmable bootstrap with caller-supplied test vectors.  Exact-integer and FFT variants agree bit for bit;
    the output decrypts to the table entry of the slot the blind rotation ends on (negated on the upper half of the torus);
    a constant table reproduces the sign bootstrap."""
    from oracle import layers_oracle as LO
    O = oracle
    rng = np.random.default_rng(21)
    luts = (rng.integers(-2000, 2000, size=(3, 1024)) * LO.UNIT & 0xFFFFFFFF).astype(np.uint32)
    msgs = (rng.integers(-2048, 2048, size=6) * LO.UNIT & 0xFFFFFFFF).astype(np.uint32)
    ct = O.encrypt(msgs, 2.0 ** -15, keyset.lwe_key, 5)
    fft = O.pbs_lut(ct, luts, keyset)
    assert np.array_equal(fft[:2], O.pbs_lut(ct[:2], luts, keyset, exact=True))
    want, slot = LO.predicted_lut_message(ct, luts, keyset.lwe_key)
    got = O.phase(fft, keyset.lwe_key)
    err = (got.astype(np.int64) - want.astype(np.int64) + 2 ** 31) % 2 ** 32 - 2 ** 31
    assert np.max(np.abs(err)) < LO.UNIT // 4           # output noise only
    assert len(set(slot // 1024)) == 2                   # both halves of the torus were exercised
    const = np.full((1, 1024), LO.UNIT, dtype=np.uint32)
    assert np.array_equal(O.pbs_lut(ct[:3], const, keyset), O.pbs(ct[:3], LO.UNIT, keyset))


def test_one_blind_rotate_step_against_numpy(oracle, keyset):
    """Independent restatement of ONE blind-rotate step in numpy int64: an LWE sample whose mask has a single
    non-zero coefficient runs exactly one external product.  Pins the rotation direction (acc = X^{-barb} tv, then
    acc += BK_i (.) ((X^{a_i} - 1) acc)), the gadget decomposition (offset, digits in [-4, 3], level p <-> 2^{32-3(p+1)}), the row
    order of the bootstrapping key (row = c*l + p, two output polynomials) and the negacyclic product, for both oracle variants."""
    O = oracle
    N, l = 1024, 10
    rng = np.random.default_rng(3)
    mu = 0x00100000
    for i, a_i, b in [(7, 0x12345678, 0x9ABCDEF0), (349, 0xFFD00000, 0x00000000), (0, 0x80000000, 0x7FFFFFFF)]:
        lwe = np.zeros(351, np.uint32)
        lwe[i], lwe[350] = a_i, b
        ms = lambda x: ((int(x) + (1 << 20)) >> 21) % (2 * N)            # modSwitchFromTorus32(x, 2N)
        bara, barb = ms(a_i), ms(b)
        assert bara != 0

        def mul_x(poly, k):                                              # X^k * poly mod X^N + 1, k in [0, 2N)
            out = np.zeros(N, np.int64)
            for j in range(N):
                t = j + k
                s = -1 if (t // N) % 2 else 1
                out[t % N] = s * poly[j]
            return out
        tv = np.full(N, mu, np.int64)
        acc = [np.zeros(N, np.int64), mul_x(tv, (2 * N - barb) % (2 * N))]
        off = sum(4 << (32 - 3 * p) for p in range(1, l + 1)) & 0xFFFFFFFF
        res = [np.zeros(N, np.int64), np.zeros(N, np.int64)]
        bsk = np.asarray(keyset.bsk, dtype=np.uint32).reshape(350, 2 * l, 2, N)
        for c in range(2):
            tmp = ((mul_x(acc[c], bara) - acc[c]) + off) & 0xFFFFFFFF
            for p in range(l):
                d = ((tmp >> (32 - 3 * (p + 1))) & 7) - 4
                for co in range(2):
                    B = bsk[i, c * l + p, co].astype(np.int32).astype(np.int64)     # signed representative: same product mod 2^32
                    full = np.convolve(d, B)                                      # exact in int64 (|d| <= 4, |B| < 2^31, 1024 terms)
                    neg = full[:N].copy()
                    neg[:N - 1] -= full[N:]
                    res[co] += neg
        want = np.concatenate([(acc[0] + res[0]) & 0xFFFFFFFF, (acc[1] + res[1]) & 0xFFFFFFFF]).astype(np.uint32)
        assert np.array_equal(O.blind_rotate(lwe, mu, keyset, exact=True), want), (i, "exact")
        assert np.array_equal(O.blind_rotate(lwe, mu, keyset, exact=False), want), (i, "fft")
