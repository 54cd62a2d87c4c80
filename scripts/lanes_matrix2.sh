echo "== 2 lanes, 32x32x8->32, 300 runs: $(python scripts/lanes_stress.py 32 32 8 32 300 2>&1 | grep -E "^run|total" | awk '/^run/ {n++} /total/ {print n+0, "bad runs;", $0}')"
echo "== 2 lanes, 16x16x8->64, 200 runs: $(python scripts/lanes_stress.py 16 16 8 64 200 2>&1 | grep -E "^run|total" | awk '/^run/ {n++} /total/ {print n+0, "bad runs;", $0}')"
