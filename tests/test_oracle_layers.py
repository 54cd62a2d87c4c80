"""CPU: the oracle's plaintext layer restatement against golden scores produced by the reference's own plaintext
build (tests/golden/ptxt_scores.json, made by tests/golden/make_golden.py from oracle/_ref)."""
import json
import os

import numpy as np
import pytest

from redsec_b200 import netspec

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ptxt_scores.json")))


def _plain_scores(spec, row=0, image=None):
    from oracle import layers_oracle as LO
    layers = LO.prepare(spec, spec["weights"])
    label, px = netspec.load_image_csv(image or spec["image"], row)
    return label, LO.plain_forward(layers, netspec.map_pixels(spec, px))


@pytest.mark.parametrize("net,csv", [("mnist/sign1024x1", "client/mnist_test.csv"), ("mnist/sign1024x2", "client/mnist_test.csv"),
                                      ("mnist/sign1024x3", "client/mnist_test.csv")])
def test_mnist_plain_scores_match_reference(net, csv):
    spec = netspec.NETS[net]()
    label, scores = _plain_scores(spec)
    gold = GOLD[f"{net}|{csv}|1"][0]
    assert label == gold["label"]
    assert list(scores) == gold["scores"]


def test_mnist_20_rows_match_reference():
    spec = netspec.NETS["mnist/sign1024x1"]()
    gold = GOLD["mnist/sign1024x1|nets/mnist/mnist_data.csv|20"]
    img = os.path.join(netspec.DATA, "nets", "mnist", "mnist_data_20.csv")
    for row, g in enumerate(gold):
        label, scores = _plain_scores(spec, row, img)
        assert label == g["label"] and list(scores) == g["scores"], row


@pytest.mark.parametrize("net", ["mnist/relu1024x1", "mnist/relu1024x2", "mnist/relu1024x3"])
def test_relu_nets_plain_scores_match_reference(net):
    """Row f4: the DoReFa-ReLU restatement (slope multiply, bias, arithmetic shift, clamp; ~x for weight -1) reproduces the
    reference's plaintext build on the client image and on 20 MNIST rows."""
    spec = netspec.NETS[net]()
    label, scores = _plain_scores(spec)
    gold = GOLD[f"{net}|client/mnist_test.csv|1"][0]
    assert label == gold["label"] and list(scores) == gold["scores"]
    img = os.path.join(netspec.DATA, "nets", "mnist", "mnist_data_20.csv")
    for row, g in enumerate(GOLD[f"{net}|nets/mnist/mnist_data.csv|20"]):
        label, scores = _plain_scores(spec, row, img)
        assert label == g["label"] and list(scores) == g["scores"], row


def test_cifar_small_plain_scores_match_reference():
    spec = netspec.NETS["cifar/binarynet_small"]()
    label, scores = _plain_scores(spec)
    gold = GOLD["cifar/binarynet_small|client/cifar_test.csv|1"][0]
    assert label == gold["label"] and list(scores) == gold["scores"]


def test_weight_file_size_identity():
    # sign1024x1 = 5 + (1+50176) + (1+4096) + (1+2560) + (1+40) bytes
    assert os.path.getsize(netspec.NETS["mnist/sign1024x1"]()["weights"]) == 56881


def test_relu_test_vectors_reproduce_the_staircase():
    """The encrypted ReLU's per-channel tables: reading slot x/2 (with the negacyclic sign on the upper half) and adding the
    constant back gives the plaintext staircase at every even x in [-2048, 2048)."""
    from oracle import layers_oracle as LO
    spec = netspec.NETS["mnist/relu1024x2"]()
    layers = LO.prepare(spec, spec["weights"])
    for L in layers[1:3]:
        assert (L.shift_bits, L.slope_bits) == ((4, 6) if L is layers[1] else (4, 8))
        tv, half = LO.relu_test_vectors(L)
        xs = np.arange(-2048, 2048, 2)
        slot = (xs // 2) % 2048
        v = tv[:, slot % 1024].astype(np.int64)
        got = ((np.where(slot < 1024, v, -v) + half) & 0xFFFFFFFF) // LO.UNIT
        want = LO.relu_shift(L, np.broadcast_to(xs[:, None], (xs.size, tv.shape[0]))).T
        inner = np.abs(xs) < 1024                        # small |x|: exact by construction
        assert np.array_equal(got[:, inner], want[:, inner])
        # beyond +-1024 the table is the negacyclic image g(x +- 2048) = -g(x) of the inner part: exact wherever x and
        # x -+ 2048 are saturated at opposite ends, i.e. up to |x| < 2048 - |transition position|
        mid = np.abs(xs) < 512
        lo, hi = want[:, mid][:, 0], want[:, mid][:, -1]
        sat = ((lo == 0) & (hi == 15)) | ((lo == 15) & (hi == 0))      # transition inside (-512, 512)
        assert sat.sum() > 100
        wide = np.abs(xs) < 1536
        assert np.array_equal(got[sat][:, wide], want[sat][:, wide])


def test_relu_weight_file_size_identity():
    """SURVEY.md 5.4 with the slope block of a ReLU+BN layer: relu1024x1 = bias(1+4) + [tern(1+196*1024*2/8) + bias(1+4096) +
    slope(1+4096)] + [tern(1+1024*10*2/8) + bias(1+40)] bytes, and prepare() consumes the file exactly."""
    from oracle import layers_oracle as LO
    spec = netspec.NETS["mnist/relu1024x1"]()
    assert os.path.getsize(spec["weights"]) == 5 + (1 + 50176) + 2 * (1 + 4096) + (1 + 2560) + (1 + 40) == 60978
    layers = LO.prepare(spec, spec["weights"])
    assert [L.q_dims for L in layers] == [(14, 14, 1), (1, 1, 1024), (1, 1, 10)]
    assert layers[1].slope is not None and layers[1].slope.min() > 0 and layers[1].twin_conv and not layers[0].twin_conv
    assert netspec.map_pixels(spec, [0, 99, 100, 199, 200, 255]).tolist() == [-1, -1, 0, 0, 1, 1]


@pytest.mark.parametrize("name", ["mnist/cnn_builder", "mnist/relu1024x1", "cifar/binarynet_small"])
def test_sampled_linear_rows_equal_the_whole_layer(name):
    """enc_linear_rows (what the full-size GPU parity tests sample with) against enc_linear on whole layers: integer conv with
    the -1/4096 zero/padding convention + sum-pool (builder CNN), plaintext-twin conv (ReLU nets), binary conv with same
    padding (CIFAR), FC after flatten, conv-less input layers.  Random uint32 rows, no bootstraps: exact mod 2^32."""
    from oracle import layers_oracle as LO
    from oracle import oracle as O
    spec = netspec.NETS[name]()
    layers = LO.prepare(spec, spec["weights"])
    rng = np.random.default_rng(5)
    h, w, c = spec["input"]
    count = h * w * c
    for li, L in enumerate(layers[:3]):
        ct = rng.integers(0, 2 ** 32, size=(count, O.LWE_WORDS), dtype=np.uint64).astype(np.uint32)
        n = int(np.prod(L.q_dims))
        idx = np.unique(np.concatenate([rng.integers(0, n, 24), [0, n - 1]]))
        if name.startswith("cifar") and li == 2:
            break                                        # 128 x 128 x 9 x 1024 pixels as a whole layer: covered by layer 1
        want = LO.enc_linear(L, ct)
        assert np.array_equal(LO.enc_linear_rows(L, ct, idx), want[idx]), (name, li)
        count = int(np.prod(L.out_dims))
