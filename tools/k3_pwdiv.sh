# A/B: header density from which a window takes the parallel walk (MICGPU_PW_DIV: headers per consumed stretch >= CHUNK / div)
for d in 32 64 128 512; do
echo "div=$d wavelet x16: $(MICGPU_PW_DIV=$d N=16 python tools/wavelet_batch.py 2>&1 | tail -2 | head -1)"
echo "div=$d mic3 shape 6: $(MICGPU_PW_DIV=$d MICGPU_K3_SHAPE=6 python tools/mic3_bench.py --side 16384 --steps 3 --warmup 1 --no-e2e --no-cpu 2>/dev/null | python -c 'import json,sys; j=json.load(sys.stdin); print(j["ms_per_step"], j["stages_ms"]["k_rle_expand"])')"
done
