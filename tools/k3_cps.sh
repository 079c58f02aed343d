# A/B: K3 kernels of the unit ranges chained (MICGPU_K3_CHAIN=1) with fewer K3 CTAs per SM, so K4 of range p co-resides with K3 of range p+1
for cps in 8 5 4 3 2; do for parts in 4 8; do
echo "chain=1 cps=$cps parts=$parts: $(MICGPU_K3_CHAIN=1 MICGPU_K3_CPS=$cps MICGPU_PARTS=$parts python bench.py --quick --no-extra --steps 8 --warmup 3 2>/dev/null | python -c 'import json,sys; j=json.load(sys.stdin); print(j["ms_per_step"], j["value"])')"
done; done
