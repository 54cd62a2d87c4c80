export RS_TRAFFIC_COUNT=65536
bash scripts/gpu_traffic_ab.sh "RS_WS_GATE=0" "RS_WS_GATE=1" "RS_WS_GATE=4" "RS_WS_GATE=16" 2>&1 | grep "==\|dram__bytes_read\|gpu__time\|lts__t_sector_hit" | tee gpurun_out/traffic_gate.log
timeout 200 python scripts/ws_ab.py "RS_WS_GATE=0" "RS_WS_GATE=1" "RS_WS_GATE=4" "RS_WS_GATE=16" --counts=592,4096,16384,65536 2>&1 | tee gpurun_out/ws_ab_gate.log
