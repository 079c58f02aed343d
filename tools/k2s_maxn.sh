# A/B: which state counts take the thread-per-unit ANS kernel (MICGPU_K2_SERIAL_MAXN; parity first with every N on it)
echo "parity MAXN=8: $(MICGPU_K2_SERIAL_MAXN=8 python -m pytest tests -m gpu -x -q 2>&1 | tail -1)"
for ns in 8 4; do for m in 2 $ns; do
echo "nstates=$ns maxn=$m: $(MICGPU_K2_SERIAL_MAXN=$m python bench.py --quick --no-extra --nstates $ns --steps 5 --warmup 3 2>/dev/null | python -c 'import json,sys; j=json.load(sys.stdin); print(j["ms_per_step"], j["roofline"]["stages_ms"])')"
done; done
