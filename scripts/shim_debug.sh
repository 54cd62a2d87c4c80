# runs the Func-level drop-in of cifar/binarynet_small on 1 and on 2 GPUs with per-layer checksums
python - <<'PY'
import os, sys
sys.path.insert(0, '.')
from redsec_b200 import client, netspec
ks = client.keygen(0)
spec = netspec.NETS["cifar/binarynet_small"]()
label, px = netspec.load_image_csv(spec["image"])
ct = client.encrypt_image(px, ks.lwe_key, seed=11)
for tree in ("tree_func", "tree_func_g2"):
    cdir = f"dropin/_build/{tree}/client"
    os.makedirs(cdir, exist_ok=True)
    client.write_keys(ks, cdir + "/secret.key", cdir + "/eval.key")
    client.write_ctxt(cdir + "/image.ctxt", ct, variance=2.0 ** -30)
PY
export LD_LIBRARY_PATH=$PWD/dropin/_build/lib:$PWD/redsec_b200:$LD_LIBRARY_PATH
for tree in tree_func tree_func_g2; do
  echo "== $tree"
  ( cd dropin/_build/$tree/nets/cifar/binarynet_small && RS_SHIM_DEBUG=1 ./gpu-encrypt.out 2>&1 | grep -E "shim:|Inference" )
done
