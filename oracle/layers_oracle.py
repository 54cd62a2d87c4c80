"""CPU restatement of REDsec's layer forward (plaintext twin and encrypted CPU path) -- TEST INFRASTRUCTURE ONLY.

Follows the reference loops, not the product's kernels:
  * weight file blocks      lib/BinOps_enc.cpp:247-297 (ternary: tag, MSB-first bits, sign then is_zero; ints: tag + int32)
  * convolution             lib/BinFunc.cpp:142-330 / lib/IntFunc.cpp:152-319 (retrieve_dims :348-366, indices :388-404)
  * sum pooling             lib/IntFunc.cpp:643-700, lib/BinFunc.cpp:677-732
  * sign activation + bias  lib/BinFunc.cpp:1044-1075, lib/IntFunc.cpp:860-889; plaintext binarize lib/BinOps.cpp:207-217
  * max pooling             lib/BinFunc.cpp:880-925 (as an OR tree at +-1/8, SURVEY.md H2 / defect R3)
  * layer sequencing        lib/BinLayer.cpp:150-241, lib/IntLayer.cpp:153-235
  * DoReFa ReLU (row f4)    lib/IntFunc.cpp:934-973 plaintext branch: (slope*x + bias) >> slope_bits, clamp to [0, 2^shift_bits);
                            slope_bits from lib/IntFunc.cpp:812-815; plaintext IntFunc conv: weight -1 contributes ~x = -x-1
                            (IntOps::invert, lib/IntOps.cpp:72-82), zero weight contributes 0 (lib/IntFunc.cpp:270).
                            Encrypted form = ONE test-vector bootstrap per neuron (relu_test_vectors below).
The plaintext path is pinned against the reference's own plaintext build (oracle/_ref, golden scores in
tests/golden/ptxt_scores.json).  The encrypted path is "parity unpinned" like the rest of the oracle (no TFHE here).
A net spec is plain data (see redsec_b200/netspec.py); it is passed in by the caller, this module imports nothing
from the product.
"""
from __future__ import annotations

import numpy as np

from . import oracle as O

UNIT = 1 << 20      # 1/4096
EIGHTH = 1 << 29    # 1/8


# ----------------------------------------------------------------------------------------------- weight file
def _read_ternary(buf: memoryview, pos: int, length: int):
    tag = buf[pos]
    assert tag in (1, 2), f"bad ternary tag {tag}"
    nbits = 1 if tag == 1 else 2
    nbytes = (length * nbits + 7) // 8
    bits = np.unpackbits(np.frombuffer(buf[pos + 1: pos + 1 + nbytes], dtype=np.uint8))  # MSB first
    if nbits == 1:
        w = np.where(bits[:length] == 1, 1, -1)
    else:
        sign, zero = bits[0:2 * length:2], bits[1:2 * length:2]
        w = np.where(zero == 1, 0, np.where(sign == 1, 1, -1))
    return w.astype(np.int8), pos + 1 + nbytes


def _read_ints(buf: memoryview, pos: int, length: int):
    tag = buf[pos]
    assert tag in (3, 4), f"bad int tag {tag}"
    v = np.frombuffer(buf[pos + 1: pos + 1 + 4 * length], dtype="<i4").copy()
    return v, pos + 1 + 4 * length


def _same_out(n, stride):
    return (n - 1) // stride + 1


class PreparedLayer:
    pass


def prepare(spec: dict, weights_path: str):
    """Dimension bookkeeping + weight loading (the E_PREP pass of {Bin,Int}Layer::run)."""
    buf = memoryview(open(weights_path, "rb").read())
    pos = 0
    h, w, dep = spec["input"]
    scale = float(spec.get("input_scale", 255))           # tDimensions.scale (nets/*/net.cpp "lay_dim.scale")
    twin_conv = any(ls["act"] == "relu" for ls in spec["layers"])   # ReLU nets follow the plaintext twin's IntFunc conv
    layers = []
    for ls in spec["layers"]:
        L = PreparedLayer()
        L.spec = ls
        conv = ls["conv"]
        if conv in ("fc", "fc_final"):
            dep, h, w = dep * h * w, 1, 1
        L.has_conv = conv != "none"
        if L.has_conv:
            wh, ww = (1, 1) if conv.startswith("fc") else ls["conv_win"]
            sh, sw = (1, 1) if conv.startswith("fc") else ls["conv_stride"]
            same = True if conv.startswith("fc") else ls["conv_same_pad"]
            L.cin = (h, w, dep)
            if same:
                oh, ow = _same_out(h, sh), _same_out(w, sw)
                ofh = (wh - 1) // 2 if sh == 1 else (oh * sh - h) // 2
                ofw = (ww - 1) // 2 if sw == 1 else (ow * sw - w) // 2
            else:
                ofh = ofw = 0
                oh, ow = (h - 2 * ((wh - 1) // 2)) // sh, (w - 2 * ((ww - 1) // 2)) // sw
            L.conv_geom = (wh, ww, sh, sw, ofh, ofw, oh, ow)
            od = ls["depth"]
            flat, pos = _read_ternary(buf, pos, wh * ww * dep * od)
            L.weights = flat.reshape(wh, ww, dep, od)        # index ((fh*W+fw)*in_dep+di)*OutDepth+od
            h, w, dep = oh, ow, od
        L.has_sumpool = ls["pool"] == "sum"
        if L.has_sumpool:
            ph_, pw_ = ls["pool_win"]
            sh, sw = ls["pool_stride"]
            L.sp_in = (h, w)
            if ls["pool_same_pad"]:
                oh, ow = _same_out(h, sh), _same_out(w, sw)
                ofh = (ph_ - 1) // 2 if sh == 1 else (oh * sh - h) // 2
                ofw = (pw_ - 1) // 2 if sw == 1 else (ow * sw - w) // 2
            else:
                ofh = ofw = 0
                oh, ow = (h - ph_ // 2 - 1) // sh + 1, (w - pw_ // 2 - 1) // sw + 1
            L.sp_geom = (ph_, pw_, sh, sw, ofh, ofw, oh, ow)
            h, w = oh, ow
            scale *= ph_ * pw_                              # lib/IntFunc.cpp:629
        L.q_dims = (h, w, dep)
        L.bias, pos = _read_ints(buf, pos, dep)
        L.twin_conv = twin_conv and ls["kind"] == "int" and L.has_conv
        if L.twin_conv:
            L.neg_count = (L.weights == -1).sum(axis=(0, 1, 2)).astype(np.int64)   # per output channel
        L.slope = None
        if ls["act"] == "relu":
            L.shift_bits = ls["shift_bits"]
            if ls.get("e_bias") == 2 and L.shift_bits > 1:
                L.slope, pos = _read_ints(buf, pos, dep)
            sc_b = 0
            while (1 << sc_b) < scale:
                sc_b += 1
            L.slope_bits = 8 + sc_b - L.shift_bits           # SLOPE_BITS=8 (lib/IntFunc.cpp:45,815)
            scale = float((1 << L.shift_bits) - 1)           # lib/IntFunc.cpp:836-837
        elif ls["act"] == "sign":
            scale = 1.0
        L.has_maxpool = ls["pool"] == "max" and ls["act"] == "sign" and conv != "fc_final"
        if L.has_maxpool:
            ph_, pw_ = ls["pool_win"]
            sh, sw = ls["pool_stride"]
            if ls["pool_same_pad"]:
                oh, ow = _same_out(h, sh), _same_out(w, sw)
            else:
                oh, ow = h // ph_, w // pw_
            L.mp_geom = (ph_, pw_, sh, sw, oh, ow)
            h, w = oh, ow
        L.out_dims = (h, w, dep)
        layers.append(L)
    assert pos == len(buf), f"weight file not consumed exactly: {pos} of {len(buf)}"
    return layers


# ----------------------------------------------------------------------------------------------- generic linear stage
def _conv(L, x, zero_term, pad_term):
    """x: [h][w][dep][...] array (ints or uint32 LWE rows).  Returns [oh][ow][od][...].
    zero_term / pad_term: what a zero weight / padded position contributes (an array broadcastable to x[0,0,0])."""
    wh, ww, sh, sw, ofh, ofw, oh, ow = L.conv_geom
    h, w, dep = L.cin
    od = L.weights.shape[3]
    tail = x.shape[3:]
    out = np.zeros((oh, ow, od) + tail, dtype=x.dtype)
    wpos = (L.weights == 1)
    wneg = (L.weights == -1)
    wzero = (L.weights == 0)
    for ph in range(oh):
        for pw in range(ow):
            acc = np.zeros((od,) + tail, dtype=x.dtype)
            for fh in range(wh):
                iy = fh + ph * sh - ofh
                for fw in range(ww):
                    ix = fw + pw * sw - ofw
                    if not (0 <= iy < h and 0 <= ix < w):       # oob (same padding)
                        if pad_term is not None:
                            acc += (dep * pad_term).astype(x.dtype)
                        continue
                    v = x[iy, ix]                                 # [dep][...]
                    P = wpos[fh, fw].astype(x.dtype)              # [dep][od]
                    M = wneg[fh, fw].astype(x.dtype)
                    if v.ndim == 1:
                        acc += P.T @ v - M.T @ v
                    else:
                        acc += P.T @ v - M.T @ v                  # uint32 wraps mod 2^32
                    if zero_term is not None:
                        nz = wzero[fh, fw].sum(axis=0).astype(x.dtype)   # zero weights per od
                        acc += nz.reshape((od,) + (1,) * len(tail)) * zero_term
            out[ph, pw] = acc
    return out


def _sumpool(L, x):
    ph_, pw_, sh, sw, ofh, ofw, oh, ow = L.sp_geom
    h, w = L.sp_in
    out = np.zeros((oh, ow) + x.shape[2:], dtype=x.dtype)
    for a in range(oh):
        for b in range(ow):
            for fh in range(ph_):
                iy = a * sh - ofh + fh
                if not 0 <= iy < h:
                    continue
                for fw in range(pw_):
                    ix = b * sw - ofw + fw
                    if 0 <= ix < w:
                        out[a, b] += x[iy, ix]
    return out


# ----------------------------------------------------------------------------------------------- plaintext twin
def plain_forward(layers, pixels, enc_conv_semantics=False, collect=None):
    """Plaintext twin (lib/unenc).  pixels: flat (h,w,c) ints already mapped x=2p-255.  Returns class scores.
    enc_conv_semantics: apply the encrypted IntFunc convention (zero weight / padding contribute -1,
    lib/IntFunc.cpp:268,277) instead of the plaintext one (0), for builder-defined integer conv layers."""
    x = np.asarray(pixels, dtype=np.int64)
    for L in layers:
        ls = L.spec
        if ls["kind"] == "bin":
            x = 2 * x - 1                                 # bits 0/1 -> -1/+1 (lib/BinFunc.cpp:248,261)
        if not L.has_conv:
            x = x.reshape((L.sp_in if L.has_sumpool else L.q_dims[:2]) + (L.q_dims[2],))
        if L.has_conv:
            x = x.reshape(L.cin)
            int_mode = ls["kind"] == "int" and enc_conv_semantics and not L.twin_conv
            term = np.int64(-1) if int_mode else None
            x = _conv(L, x, term, term)
            if L.twin_conv:
                x = x - L.neg_count                       # weight -1 contributes ~x = -x - 1
        if L.has_sumpool:
            x = _sumpool(L, x)
        if ls["act"] == "relu":
            x = relu_shift(L, x)
            if collect is not None:
                collect.append(x.copy())
            x = x.reshape(-1)
            continue
        x = x + L.bias.astype(np.int64)
        if ls["act"] == "none":
            if L is layers[-1]:
                return x.reshape(-1)
            if collect is not None:
                collect.append(x.copy())
            x = x.reshape(-1)
            continue
        x = (x >= 0).astype(np.int64)                    # binarize: val<0 -> 0 else 1
        if L.has_maxpool:
            ph_, pw_, sh, sw, oh, ow = L.mp_geom
            y = np.zeros((oh, ow, x.shape[2]), dtype=np.int64)
            for a in range(oh):
                for b in range(ow):
                    y[a, b] = x[a * sh: a * sh + ph_, b * sw: b * sw + pw_].reshape(-1, x.shape[2]).max(axis=0)
            x = y
        if collect is not None:
            collect.append(x.copy())
        x = x.reshape(-1)
    return x


def relu_shift(L, x):
    """Plaintext DoReFa ReLU of lib/IntFunc.cpp:964-967 on integers x[..., channel]."""
    x = np.asarray(x, dtype=np.int64)
    slope = L.slope.astype(np.int64) if L.slope is not None else np.ones(x.shape[-1], dtype=np.int64)
    v = (x * slope + L.bias.astype(np.int64)) >> L.slope_bits          # IntOps::shift: arithmetic >>
    return np.clip(v, 0, (1 << L.shift_bits) - 1)                       # IntOps::relu


def relu_test_vectors(L):
    """Per-channel test vectors [dep][1024] (torus32) of the encrypted ReLU: the neuron value x sits on the torus in units
    of 1/4096, a bootstrap resolves 2N = 2048 phase slots, so slot j <-> x = 2j.  The staircase f saturates on both sides, so
    g = f - (2^shift_bits - 1)/2 is made negacyclic: slots [0,512) hold g(2j) (small positive x; large negative x sees
    -g = saturated value), slots [512,1024) hold -g(2j - 2048) (small negative x).  The caller adds the constant
    (2^shift_bits - 1)/2 back after the bootstrap."""
    dep = L.q_dims[2]
    half = ((1 << L.shift_bits) - 1) * (UNIT // 2)
    j = np.arange(1024, dtype=np.int64)
    xs = np.where(j < 512, 2 * j, 2 * j - 2048)                         # [1024]
    f = relu_shift(L, np.broadcast_to(xs[:, None], (1024, dep)))        # [1024][dep]
    g = f * UNIT - half
    tv = np.where((j < 512)[:, None], g, -g)
    return np.ascontiguousarray((tv.T & 0xFFFFFFFF).astype(np.uint32)), half


def predicted_lut_message(ct, luts, lwe_key):
    """What a test-vector bootstrap of ct[c] with luts[c % m] must decrypt to (torus32, before output noise): the blind rotation
    ends at slot (barb - sum bara_i s_i) mod 2N, slots [N,2N) read the negated table."""
    ct = np.ascontiguousarray(ct, dtype=np.uint32).reshape(-1, O.LWE_WORDS)
    luts = np.ascontiguousarray(luts, dtype=np.uint32).reshape(-1, O.N)
    bar = (((ct.astype(np.uint64) << np.uint64(32)) + np.uint64(1 << 52)) >> np.uint64(53)).astype(np.int64) % (2 * O.N)  # round to 2N
    slot = (bar[:, O.n] - bar[:, :O.n] @ np.asarray(lwe_key, dtype=np.int64)) % (2 * O.N)
    rows = np.arange(ct.shape[0]) % luts.shape[0]
    v = luts[rows, slot % O.N].astype(np.int64)
    return np.where(slot < O.N, v, -v) & 0xFFFFFFFF, slot


def plain_layer_preact(L, x_in):
    """Pre-activation integers [h][w][c] of ONE layer of the plaintext twin for the inputs that layer receives
    (bits 0/1 for a BinLayer, integers for an IntLayer); `encrypted` conventions for IntLayer convs.  Used by the
    full-size teacher-forced check: feed the DECRYPTED inputs of an encrypted layer, compare signs away from 0."""
    ls = L.spec
    x = np.asarray(x_in, dtype=np.int64)
    if ls["kind"] == "bin":
        x = 2 * x - 1
    if L.has_conv:
        x = x.reshape(L.cin)
        term = np.int64(-1) if ls["kind"] == "int" else None
        x = _conv(L, x, term, term)
    else:
        x = x.reshape((L.sp_in if L.has_sumpool else L.q_dims[:2]) + (L.q_dims[2],))
    if L.has_sumpool:
        x = _sumpool(L, x)
    return x.reshape(L.q_dims) + L.bias.astype(np.int64)


def plain_maxpool(L, bits):
    ph_, pw_, sh, sw, oh, ow = L.mp_geom
    y = np.zeros((oh, ow, bits.shape[2]), dtype=bits.dtype)
    for a in range(oh):
        for b in range(ow):
            y[a, b] = bits[a * sh: a * sh + ph_, b * sw: b * sw + pw_].reshape(-1, bits.shape[2]).max(axis=0)
    return y


# ----------------------------------------------------------------------------------------------- encrypted CPU path
def enc_linear(L, ct):
    """Bootstrap-free part of one layer on LWE rows [count][351] -> [(h,w,c) flat][351] (uint32 wrap-around)."""
    ls = L.spec
    x = np.ascontiguousarray(ct, dtype=np.uint32)
    if L.has_conv:
        x = x.reshape(L.cin + (O.LWE_WORDS,))
        term = None
        if ls["kind"] == "int" and not L.twin_conv:     # trivial sample (0, -1/4096)
            term = np.zeros(O.LWE_WORDS, dtype=np.uint32)
            term[O.n] = (-UNIT) & 0xFFFFFFFF
        x = _conv(L, x, term, term)
        if L.twin_conv:                                 # plaintext-twin IntFunc conv: weight -1 contributes -x - 1
            x[..., O.n] -= (L.neg_count * UNIT & 0xFFFFFFFF).astype(np.uint32)[None, None, :]
    if L.has_sumpool:
        if not L.has_conv:
            x = x.reshape(L.sp_in + (L.q_dims[2], O.LWE_WORDS))
        x = _sumpool(L, x)
    x = x.reshape(L.q_dims + (O.LWE_WORDS,)).copy()
    if ls["act"] != "relu":                             # the ReLU bias lives inside the test vector (slope*x + bias)
        x[..., O.n] += (L.bias.astype(np.int64) * UNIT & 0xFFFFFFFF).astype(np.uint32)[None, None, :]
    return x.reshape(-1, O.LWE_WORDS)


def enc_linear_rows(L, ct, idx):
    """Rows `idx` (flat indices into the (h,w,c) pre-activation array) of enc_linear(L, ct), computed one output at a time
    from the reference's index formulas (lib/BinFunc.cpp:228-310 window gather, lib/IntFunc.cpp:665-697 pooling) without
    forming the whole layer: what the sampled full-size parity tests use (CIFAR conv2 is 53 G word-adds as a whole)."""
    ls = L.spec
    x = np.ascontiguousarray(ct, dtype=np.uint32)
    W = O.LWE_WORDS
    qh_, qw_, qd_ = L.q_dims
    unit_b = np.zeros(W, dtype=np.uint32)
    unit_b[O.n] = UNIT
    if L.has_conv:
        x = x.reshape(L.cin + (W,))
        wh, ww, sh, sw, ofh, ofw, oh, ow = L.conv_geom
        h, w, dep = L.cin
        int_mode = ls["kind"] == "int" and not L.twin_conv
    elif L.has_sumpool:
        x = x.reshape(L.sp_in + (qd_, W))
    else:
        x = x.reshape(L.q_dims + (W,))

    def conv_at(ph, pw, od):
        if not L.has_conv:
            return x[ph, pw, od]
        acc = np.zeros(W, dtype=np.uint32)
        extra = 0                                       # count of -1/4096 contributions (IntFunc zero weight / padding)
        for fh in range(wh):
            iy = fh + ph * sh - ofh
            for fw in range(ww):
                ix = fw + pw * sw - ofw
                wv = L.weights[fh, fw, :, od].astype(np.int64)
                if not (0 <= iy < h and 0 <= ix < w):
                    if int_mode:
                        extra += dep
                    continue
                v = x[iy, ix]                           # [dep][W]
                pos, neg = wv == 1, wv == -1
                acc += v[pos].sum(axis=0, dtype=np.uint32)
                acc -= v[neg].sum(axis=0, dtype=np.uint32)
                if int_mode:
                    extra += int((wv == 0).sum())
        with np.errstate(over="ignore"):
            acc[O.n] -= np.uint32((extra * UNIT) & 0xFFFFFFFF)
            if L.twin_conv:
                acc[O.n] -= np.uint32((int(L.neg_count[od]) * UNIT) & 0xFFFFFFFF)
        return acc

    out = np.zeros((len(idx), W), dtype=np.uint32)
    with np.errstate(over="ignore"):
        _rows_loop(L, idx, out, conv_at, unit_b)
    return out


def _rows_loop(L, idx, out, conv_at, unit_b):
    ls = L.spec
    W = O.LWE_WORDS
    qh_, qw_, qd_ = L.q_dims
    for k, flat in enumerate(np.asarray(idx, dtype=np.int64)):
        c = int(flat % qd_); qw = int((flat // qd_) % qw_); qh = int(flat // (qd_ * qw_))
        if L.has_sumpool:
            ph_, pw_, sh2, sw2, ofh2, ofw2, oh2, ow2 = L.sp_geom
            ih, iw = L.sp_in
            acc = np.zeros(W, dtype=np.uint32)
            for fh in range(ph_):
                iy = qh * sh2 - ofh2 + fh
                if not 0 <= iy < ih:
                    continue
                for fw in range(pw_):
                    ix = qw * sw2 - ofw2 + fw
                    if 0 <= ix < iw:
                        acc += conv_at(iy, ix, c)
        else:
            acc = conv_at(qh, qw, c).copy()
        if ls["act"] != "relu":
            acc[O.n] += np.uint32((int(L.bias[c]) * UNIT) & 0xFFFFFFFF)
        out[k] = acc


def enc_layer_rows(L, ct, idx, ks, threads=0):
    """Rows `idx` of enc_layer_forward(L, ct, ks): flat indices into the layer OUTPUT ((oh,ow,c) after a max-pool).  A pooled
    output costs 4 sign bootstraps + 3 ORs, any other one bootstrap (none for an activation-free layer)."""
    idx = np.asarray(idx, dtype=np.int64)
    if L.spec["act"] == "none":
        return enc_linear_rows(L, ct, idx)
    if L.spec["act"] == "relu":
        lin = enc_linear_rows(L, ct, idx)
        tv, half = relu_test_vectors(L)
        # pbs_lut uses table c % len(luts) for ciphertext c: hand each sampled row its own channel's table
        out = O.pbs_lut(lin, tv[idx % L.q_dims[2]], ks, threads=threads)
        out[:, O.n] += np.uint32(half)
        return out
    if not L.has_maxpool:
        return O.pbs(enc_linear_rows(L, ct, idx), UNIT, ks, threads=threads)
    ph_, pw_, sh, sw, oh, ow = L.mp_geom
    assert (ph_, pw_) == (2, 2), "oracle OR tree restated for the 2x2 windows the shipped nets use"
    qh_, qw_, qd_ = L.q_dims
    c = idx % qd_; b = (idx // qd_) % ow; a = idx // (qd_ * ow)
    def q_index(fh, fw):
        return ((a * sh + fh) * qw_ + (b * sw + fw)) * qd_ + c
    bits = [O.pbs(enc_linear_rows(L, ct, q_index(fh, fw)), EIGHTH, ks, threads=threads) for fh in range(2) for fw in range(2)]
    top = O.gate("OR", bits[0], bits[1], EIGHTH, ks, threads=threads)
    bot = O.gate("OR", bits[2], bits[3], EIGHTH, ks, threads=threads)
    return O.gate("OR", top, bot, UNIT, ks, threads=threads)


def enc_layer_forward(L, ct, ks, threads=0):
    """One encrypted layer: linear part, one sign bootstrap per neuron, max-pool OR tree."""
    lin = enc_linear(L, ct)
    if L.spec["act"] == "none":
        return lin
    if L.spec["act"] == "relu":                         # ONE test-vector bootstrap per neuron, then + (2^shift_bits-1)/2
        tv, half = relu_test_vectors(L)
        out = O.pbs_lut(lin, tv, ks, threads=threads)
        out[:, O.n] += np.uint32(half)
        return out
    if not L.has_maxpool:
        return O.pbs(lin, UNIT, ks, threads=threads)
    bits = O.pbs(lin, EIGHTH, ks, threads=threads).reshape(L.q_dims + (O.LWE_WORDS,))
    ph_, pw_, sh, sw, oh, ow = L.mp_geom
    assert (ph_, pw_) == (2, 2), "oracle OR tree restated for the 2x2 windows the shipped nets use"
    a = bits[0:oh * sh:sh, 0:ow * sw:sw].reshape(-1, O.LWE_WORDS)
    b = bits[0:oh * sh:sh, 1:ow * sw:sw].reshape(-1, O.LWE_WORDS)
    c = bits[1:oh * sh:sh, 0:ow * sw:sw].reshape(-1, O.LWE_WORDS)
    d = bits[1:oh * sh:sh, 1:ow * sw:sw].reshape(-1, O.LWE_WORDS)
    # window element order (fh,fw) = (0,0),(0,1),(1,0),(1,1): level 1 pairs (0,1) and (2,3); level 2 emits +-1/4096
    top = O.gate("OR", a, b, EIGHTH, ks, threads=threads)
    bot = O.gate("OR", c, d, EIGHTH, ks, threads=threads)
    return O.gate("OR", top, bot, UNIT, ks, threads=threads)


def enc_forward(layers, ct, ks, threads=0, collect=None):
    x = ct
    for L in layers:
        x = enc_layer_forward(L, x, ks, threads)
        if collect is not None:
            collect.append(x.copy())
    return x


def encode_pixels(pixels):
    """client/encrypt_image.cpp:76: ptxt = 2*pixel - 255, mu = modSwitchToTorus32(ptxt, 4096)."""
    v = 2 * np.asarray(pixels, dtype=np.int64) - 255
    return (v * UNIT) & 0xFFFFFFFF


def count_bootstraps(layers):
    n = 0
    for L in layers:
        if L.spec["act"] not in ("sign", "relu"):
            continue
        h, w, d = L.q_dims
        n += h * w * d
        if L.has_maxpool:
            oh, ow = L.mp_geom[4], L.mp_geom[5]
            n += 3 * oh * ow * d
    return n
