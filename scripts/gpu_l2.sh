# L2 residency hint sweep: DRAM bytes and time of one 2^16-ciphertext blind-rotate launch per RS_L2_KEEP value
mkdir -p gpurun_out
for K in 0.25 0.35 0.45; do
  RS_L2_KEEP=$K bash scripts/gpu_traffic.sh l2_$K > /dev/null 2>&1
  echo "== RS_L2_KEEP=$K"; grep -E "blind_rotate" gpurun_out/traffic_l2_$K.csv | awk -F'","' '{print $13, $15}' | tr -d '"'
done
