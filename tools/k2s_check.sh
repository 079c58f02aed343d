# parity + the three workloads whose time is the thread-per-unit ANS kernel
echo "all: $(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2)"
echo "2-state: $(python bench.py --quick --no-extra --nstates 2 --steps 5 --warmup 3 2>/dev/null | python -c 'import json,sys; j=json.load(sys.stdin); print(j["ms_per_step"], j["roofline"]["stages_ms"])')"
echo "4-state: $(python bench.py --quick --no-extra --nstates 4 --steps 5 --warmup 3 2>/dev/null | python -c 'import json,sys; j=json.load(sys.stdin); print(j["ms_per_step"], j["roofline"]["stages_ms"])')"
echo "mic3: $(python tools/mic3_bench.py --side 16384 --steps 3 --warmup 1 --no-e2e --no-cpu 2>/dev/null | python -c 'import json,sys; j=json.load(sys.stdin); print(j["ms_per_step"], j["stages_ms"])')"
FRAMES=12 python tools/mic2_profile.py 2>&1 | grep -E "^independent|^temporal|exact"
