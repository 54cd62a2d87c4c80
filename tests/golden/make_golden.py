"""Regenerates tests/golden/ptxt_scores.json from the reference's own plaintext build (oracle/_ref, built by
oracle/build_ref.sh from /root/reference).  Run in the build container only; the JSON is committed."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("REF", "/root/reference")
subprocess.check_call([os.path.join(ROOT, "oracle", "build_ref.sh")])
out = {}
cases = [("mnist/sign1024x1", "client/mnist_test.csv", 1), ("mnist/sign1024x2", "client/mnist_test.csv", 1),
         ("mnist/sign1024x3", "client/mnist_test.csv", 1), ("cifar/binarynet", "client/cifar_test.csv", 1),
         ("cifar/binarynet_small", "client/cifar_test.csv", 1), ("mnist/sign1024x1", "nets/mnist/mnist_data.csv", 20),
         # DoReFa-ReLU nets (row f4): inputs x = pixel/100 - 1 (nets/mnist/relu1024x1/main.cpp:203)
         ("mnist/relu1024x1", "client/mnist_test.csv", 1), ("mnist/relu1024x2", "client/mnist_test.csv", 1),
         ("mnist/relu1024x3", "client/mnist_test.csv", 1), ("mnist/relu1024x1", "nets/mnist/mnist_data.csv", 20),
         ("mnist/relu1024x2", "nets/mnist/mnist_data.csv", 20), ("mnist/relu1024x3", "nets/mnist/mnist_data.csv", 20)]
for net, csv, rows in cases:
    exe = os.path.join(ROOT, "oracle", "_ref", "ptxt_" + net.replace("/", "_"))
    extra = ["relu"] if "relu" in net else []
    txt = subprocess.check_output([exe, os.path.join(REF, csv), str(rows)] + extra, cwd=os.path.join(REF, "nets", net), text=True)
    res = []
    for line in txt.splitlines():
        if line.startswith("label"):
            parts = line.split()
            res.append({"label": int(parts[1]), "scores": [int(v) for v in parts[3:]]})
    out[f"{net}|{csv}|{rows}"] = res
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "ptxt_scores.json"), "w"), indent=1)
print("wrote", len(out), "cases")
