export RS_TRAFFIC_COUNT=65536
bash scripts/gpu_traffic_ab.sh "RS_WS_LOOKAHEAD=5" "RS_WS_LOOKAHEAD=3" "RS_WS_LOOKAHEAD=2" "RS_WS_LOOKAHEAD=1" 2>&1 | grep "==\|dram__bytes_read\|gpu__time\|lts__t_sector_hit" | tee gpurun_out/traffic_lookahead.log
timeout 200 python scripts/ws_ab.py "RS_WS_LOOKAHEAD=5" "RS_WS_LOOKAHEAD=3" "RS_WS_LOOKAHEAD=2" "RS_WS_LOOKAHEAD=1" --counts=592,16384,65536 2>&1 | tee gpurun_out/ws_ab_lookahead.log
