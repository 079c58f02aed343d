# parity of the parallel header walk (K3 shapes 5 / 6), then its effect on the run-heavy workloads
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for s in 5 6; do echo "forced shape $s: $(MICGPU_K3_SHAPE=$s python -m pytest tests -m gpu -x -q 2>&1 | tail -1)"; done
for s in 2 6; do echo "mic3 shape $s: $(MICGPU_K3_SHAPE=$s python tools/mic3_bench.py --side 16384 --steps 3 --warmup 1 --no-e2e --no-cpu 2>/dev/null | python -c 'import json,sys; j=json.load(sys.stdin); print(j["ms_per_step"], j["stages_ms"])')"; done
for s in 0 5; do echo "mic2 shape $s:"; MICGPU_K3_SHAPE=$s FRAMES=12 python tools/mic2_profile.py 2>&1 | grep -E "^independent|^temporal|exact"; done
for s in 0 5; do echo "wavelet x16 shape $s: $(MICGPU_K3_SHAPE=$s N=16 python tools/wavelet_batch.py 2>&1 | tail -2 | head -1)"; done
