# A/B: units per warp of the packed ANS kernel (MICGPU_K2_UPW; default 4 for 8-state)
for u in 4 2 1; do
echo "upw=$u: $(MICGPU_K2_UPW=$u python bench.py --quick --no-extra --steps 5 --warmup 3 2>/dev/null | python -c 'import json,sys; j=json.load(sys.stdin); print(j["ms_per_step"], j["roofline"]["stages_ms"])')"
done
