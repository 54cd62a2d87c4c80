for cfg in "RS_LANES_MAX=4" "RS_LANES_MAX=1" "RS_LANES_MAX=2" "RS_LANE_SERIAL=1"; do
  echo "== $cfg: $(env $cfg python scripts/lanes_stress.py 32 32 8 32 100 2>&1 | grep -E "^run|total" | awk '/^run/ {n++} /total/ {print n+0, "bad runs;", $0}')"
done
