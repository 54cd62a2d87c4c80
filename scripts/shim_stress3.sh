python - <<'PY'
import os, sys
sys.path.insert(0, '.')
from redsec_b200 import client, netspec
ks = client.keygen(0)
spec = netspec.NETS["cifar/binarynet_small"]()
label, px = netspec.load_image_csv(spec["image"])
ct = client.encrypt_image(px, ks.lwe_key, seed=11)
for tree in ("tree", "tree_func_g2", "tree_g2"):
    cdir = f"dropin/_build/{tree}/client"
    os.makedirs(cdir, exist_ok=True)
    client.write_keys(ks, cdir + "/secret.key", cdir + "/eval.key")
    client.write_ctxt(cdir + "/image.ctxt", ct, variance=2.0 ** -30)
PY
export LD_LIBRARY_PATH=$PWD/dropin/_build/lib:$PWD/redsec_b200:$LD_LIBRARY_PATH
run() { ( cd dropin/_build/$1/nets/cifar/binarynet_small && rm -f ../../../client/network_output.ctxt && ./gpu-encrypt.out > /dev/null 2>&1; md5sum ../../../client/network_output.ctxt | cut -c1-8 ); }
echo "2 GPUs, no other process:   $(for i in 1 2 3 4 5 6 7 8 9 10; do run tree_func_g2; done | tr '\n' ' ')"
python -c "
import sys, time; sys.path.insert(0, '.')
import redsec_b200 as rs
from redsec_b200 import client
ks = client.keygen(0); e = rs.Engine(0); e.load_eval_key(ks.bsk, ks.ksk); print('holder ready', flush=True); time.sleep(900)" &
HOLD=$!
sleep 25
echo "1 GPU, idle holder process: $(for i in 1 2 3 4 5 6 7 8 9 10; do run tree; done | tr '\n' ' ')"
kill $HOLD
