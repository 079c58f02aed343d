python -m pytest tests/test_gpu_wsi.py tests/test_gpu_multi.py -x -q 2>&1 | tail -1
for s in 1 2 4 8; do echo "split=$s: $(MICGPU_WSI_SPLIT=$s python tools/mic3_bench.py --side 32768 --steps 3 --warmup 1 --no-cpu 2>/dev/null | python -c 'import json,sys; j=json.load(sys.stdin); print(j["ms_per_step"], j["e2e"])')"; done
