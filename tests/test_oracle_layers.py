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
    return label, LO.plain_forward(layers, 2 * np.asarray(px) - 255)


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


def test_cifar_small_plain_scores_match_reference():
    spec = netspec.NETS["cifar/binarynet_small"]()
    label, scores = _plain_scores(spec)
    gold = GOLD["cifar/binarynet_small|client/cifar_test.csv|1"][0]
    assert label == gold["label"] and list(scores) == gold["scores"]


def test_weight_file_size_identity():
    # sign1024x1 = 5 + (1+50176) + (1+4096) + (1+2560) + (1+40) bytes
    assert os.path.getsize(netspec.NETS["mnist/sign1024x1"]()["weights"]) == 56881
