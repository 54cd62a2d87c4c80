"""Network architectures as plain data (what compiler/compiler.py bakes into nets/*/net.cpp).

References: nets/mnist/sign1024x{1,2,3}/net.cpp:60-112, nets/cifar/binarynet/net.cpp:92-215,
nets/cifar/binarynet_small/net.cpp.  `mnist_cnn` is builder-defined (no CNN ships in nets/mnist).
Enum spellings map to lib/Layer.h:58-101.
"""
from __future__ import annotations

import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "data")

CONV = {"none": 0, "conv": 1, "fc": 2, "fc_final": 3}
POOL = {"none": 0, "max": 1, "sum": 2}
ACT = {"none": 0, "sign": 1, "relu": 2}


def _layer(kind, conv, depth, pool, act, conv_win=(1, 1), conv_stride=(1, 1), conv_same_pad=False,
           pool_win=(2, 2), pool_stride=(2, 2), pool_same_pad=False, e_bias=0, version=2, shift_bits=0):
    return dict(kind=kind, conv=conv, depth=depth, pool=pool, act=act, conv_win=conv_win, conv_stride=conv_stride,
                conv_same_pad=conv_same_pad, pool_win=pool_win, pool_stride=pool_stride, pool_same_pad=pool_same_pad,
                e_bias=e_bias, version=version, shift_bits=shift_bits)


def mnist_sign(n_hidden: int) -> dict:
    layers = [_layer("int", "none", 1, "sum", "sign")]
    layers += [_layer("bin", "fc", 1024, "none", "sign", e_bias=2) for _ in range(n_hidden)]
    layers += [_layer("bin", "fc_final", 10, "none", "none")]
    name = f"mnist/sign1024x{n_hidden}"
    return dict(name=name, input=(28, 28, 1), layers=layers, weights=os.path.join(DATA, "nets", name, "var_prep.dat"),
                image=os.path.join(DATA, "client", "mnist_test.csv"))


def mnist_relu(n_hidden: int) -> dict:
    """nets/mnist/relu1024x{1,2,3}/net.cpp:96-170: all-IntLayer MLP with 4-bit DoReFa ReLU.
    Inputs are ternarised pixels x = pixel/100 - 1 (nets/mnist/relu1024x1/main.cpp:203), lay_dim.scale = 1."""
    layers = [_layer("int", "none", 1, "sum", "none")]
    layers += [_layer("int", "fc", 1024, "none", "relu", e_bias=2, shift_bits=4) for _ in range(n_hidden)]
    layers += [_layer("int", "fc", 10, "none", "none")]
    name = f"mnist/relu1024x{n_hidden}"
    return dict(name=name, input=(28, 28, 1), layers=layers, weights=os.path.join(DATA, "nets", name, "var_prep.dat"),
                image=os.path.join(DATA, "client", "mnist_test.csv"), input_scale=1, input_map="relu")


def cifar_binarynet(small: bool = False) -> dict:
    widths = (64, 64, 128, 128, 256, 256, 512, 512) if small else (128, 128, 256, 256, 512, 512, 1024, 1024)
    conv = dict(conv_win=(3, 3), conv_stride=(1, 1), conv_same_pad=True, e_bias=2)
    layers = [_layer("int", "none", 1, "none", "sign")]
    for i in range(6):
        layers.append(_layer("bin", "conv", widths[i], "max" if i % 2 == 1 else "none", "sign", **conv))
    layers.append(_layer("bin", "fc", widths[6], "none", "sign", e_bias=2))
    layers.append(_layer("bin", "fc", widths[7], "none", "sign", e_bias=2))
    layers.append(_layer("bin", "fc_final", 10, "none", "none"))
    name = "cifar/binarynet_small" if small else "cifar/binarynet"
    return dict(name=name, input=(32, 32, 3), layers=layers, weights=os.path.join(DATA, "nets", name, "var_prep.dat"),
                image=os.path.join(DATA, "client", "cifar_test.csv"))


def mnist_cnn() -> dict:
    """Builder-defined discretized MNIST CNN: 5-bit inputs, random ternary weights."""
    conv = dict(conv_win=(3, 3), conv_stride=(1, 1), conv_same_pad=True)
    layers = [
        _layer("int", "none", 1, "none", "none"),
        _layer("int", "conv", 16, "sum", "sign", **conv),
        _layer("bin", "conv", 32, "sum", "sign", **conv),
        _layer("bin", "fc", 128, "none", "sign"),
        _layer("bin", "fc_final", 10, "none", "none"),
    ]
    return dict(name="mnist/cnn_builder", input=(28, 28, 1), layers=layers,
                weights=os.path.join(DATA, "nets", "mnist", "cnn_builder", "var_prep.dat"),
                image=os.path.join(DATA, "client", "mnist_test.csv"), five_bit_inputs=True)


NETS = {
    "mnist/sign1024x1": lambda: mnist_sign(1),
    "mnist/sign1024x2": lambda: mnist_sign(2),
    "mnist/sign1024x3": lambda: mnist_sign(3),
    "mnist/relu1024x1": lambda: mnist_relu(1),
    "mnist/relu1024x2": lambda: mnist_relu(2),
    "mnist/relu1024x3": lambda: mnist_relu(3),
    "cifar/binarynet": lambda: cifar_binarynet(False),
    "cifar/binarynet_small": lambda: cifar_binarynet(True),
    "mnist/cnn_builder": mnist_cnn,
}


def load_image_csv(path: str, row: int = 0):
    """label, pixels (flat (h,w,c) as in the CSV) -- client/image_converter.py:27-37 keeps this order."""
    with open(path) as f:
        lines = [l for l in f.read().splitlines() if l and l[0].isdigit()]
    vals = [int(v) for v in lines[row].split(",") if v != ""]
    return vals[0], vals[1:]


def map_pixels(spec: dict, pixels):
    """Pixel -> plaintext integer the client encrypts: 2p-255 (client/encrypt_image.cpp:76), 5-bit 2(p>>3)-31 for the
    builder CNN, or the ternarised p/100-1 of the ReLU nets (nets/mnist/relu1024x1/main.cpp:203)."""
    import numpy as np
    px = np.asarray(pixels, dtype=np.int64)
    if spec.get("input_map") == "relu":
        return px // 100 - 1
    if spec.get("five_bit_inputs"):
        return 2 * (px >> 3) - 31
    return 2 * px - 255


# ---------------------------------------------------------------------------------------------- synthetic weights
def layer_shapes(spec: dict):
    """Per layer: (ternary weight count or 0, bias length) following the E_PREP dimension pass."""
    h, w, dep = spec["input"]
    shapes = []
    for ls in spec["layers"]:
        nw = 0
        if ls["conv"] in ("fc", "fc_final"):
            dep, h, w = dep * h * w, 1, 1
        if ls["conv"] != "none":
            wh, ww = (1, 1) if ls["conv"].startswith("fc") else ls["conv_win"]
            sh, sw = (1, 1) if ls["conv"].startswith("fc") else ls["conv_stride"]
            same = True if ls["conv"].startswith("fc") else ls["conv_same_pad"]
            nw = wh * ww * dep * ls["depth"]
            if same:
                h, w = (h - 1) // sh + 1, (w - 1) // sw + 1
            else:
                h, w = (h - 2 * ((wh - 1) // 2)) // sh, (w - 2 * ((ww - 1) // 2)) // sw
            dep = ls["depth"]
        if ls["pool"] == "sum":
            (ph, pw), (sh, sw) = ls["pool_win"], ls["pool_stride"]
            if ls["pool_same_pad"]:
                h, w = (h - 1) // sh + 1, (w - 1) // sw + 1
            else:
                h, w = (h - ph // 2 - 1) // sh + 1, (w - pw // 2 - 1) // sw + 1
        shapes.append((nw, dep))
        if ls["pool"] == "max" and ls["act"] == "sign":
            (ph, pw), (sh, sw) = ls["pool_win"], ls["pool_stride"]
            h, w = ((h - 1) // sh + 1, (w - 1) // sw + 1) if ls["pool_same_pad"] else (h // ph, w // pw)
    return shapes


def write_random_weights(spec: dict, path: str, seed: int, p_zero: float = 0.1, bias_range: int = 8):
    """var_prep.dat (SURVEY.md 5.4 layout; writer twin of lib/BinOps.cpp:393-467) with i.i.d. ternary weights
    P(-1,0,+1) = ((1-p_zero)/2, p_zero, (1-p_zero)/2) and uniform integer biases in [-bias_range, bias_range]."""
    import numpy as np
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        for nw, nb in layer_shapes(spec):
            if nw:
                u = rng.random(nw)
                zero = u < p_zero
                sign = rng.integers(0, 2, nw).astype(np.uint8)
                bits = np.empty(2 * nw, np.uint8)
                bits[0::2] = sign
                bits[1::2] = zero
                f.write(bytes([2]))                       # TERN_FMT
                f.write(np.packbits(bits).tobytes())      # MSB first
            f.write(bytes([4]))                           # INT32_FMT
            f.write(rng.integers(-bias_range, bias_range + 1, nb).astype("<i4").tobytes())
    return path


def tiny_cifar_like() -> dict:
    """Small conv + max-pool net for parity tests (same layer kinds as nets/cifar/binarynet, 8x8x3 input)."""
    conv = dict(conv_win=(3, 3), conv_stride=(1, 1), conv_same_pad=True, e_bias=2)
    layers = [_layer("int", "none", 1, "none", "sign"),
              _layer("bin", "conv", 16, "none", "sign", **conv),
              _layer("bin", "conv", 16, "max", "sign", **conv),
              _layer("bin", "fc", 32, "none", "sign", e_bias=2),
              _layer("bin", "fc_final", 10, "none", "none")]
    return dict(name="test/tiny_cifar", input=(8, 8, 3), layers=layers, weights=None, image=None)
