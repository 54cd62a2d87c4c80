python - <<'PY'
import os, sys
sys.path.insert(0, '.')
from redsec_b200 import client, netspec
ks = client.keygen(0)
spec = netspec.NETS["cifar/binarynet_small"]()
label, px = netspec.load_image_csv(spec["image"])
ct = client.encrypt_image(px, ks.lwe_key, seed=11)
for tree in ("tree_func_g2",):
    cdir = f"dropin/_build/{tree}/client"
    os.makedirs(cdir, exist_ok=True)
    client.write_keys(ks, cdir + "/secret.key", cdir + "/eval.key")
    client.write_ctxt(cdir + "/image.ctxt", ct, variance=2.0 ** -30)
PY
export LD_LIBRARY_PATH=$PWD/dropin/_build/lib:$PWD/redsec_b200:$LD_LIBRARY_PATH
cd dropin/_build/tree_func_g2/nets/cifar/binarynet_small
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  RS_SHIM_DEBUG=1 ./gpu-encrypt.out 2>&1 | grep "shim:" | sed 's/shim: //' > /tmp/run_$i.txt
  if [ $i -gt 1 ] && ! cmp -s /tmp/run_1.txt /tmp/run_$i.txt; then echo "run $i differs from run 1:"; diff /tmp/run_1.txt /tmp/run_$i.txt | head -8; fi
done
echo "run 1 final:"; tail -2 /tmp/run_1.txt
