#!/bin/bash
# A/B of the K3 launch shapes (MICGPU_K3_SHAPE: 0 = <256,4096>, 1 = <128,4096>, 2 = <128,2048>, 3 = <64,2048>, 4 = <64,1024>):
# parity subset first, then the per-kernel times of the PICS-8 batch and of the MIC3 tile workload.
O=gpurun_out/k3shapes
mkdir -p $O
for s in 0 1 2 3 4; do
  export MICGPU_K3_SHAPE=$s
  timeout 600 python -m pytest tests/test_gpu_decode.py tests/test_gpu_wsi.py tests/test_gpu_robustness.py tests/test_gpu_wavelet.py -x -q 2>&1 | tail -1 > $O/pytest_$s.txt
  python bench.py --quick --no-extra --steps 5 --warmup 3 > $O/pics_$s.json 2> $O/pics_$s.err
  python bench.py --quick --no-extra --nstates 2 --steps 3 --warmup 2 > $O/pics2_$s.json 2> $O/pics2_$s.err
  python tools/mic3_bench.py --side 16384 --steps 3 --warmup 1 --no-e2e --no-cpu > $O/mic3_$s.json 2> $O/mic3_$s.err
  echo "shape $s: $(cat $O/pytest_$s.txt)"
  python - <<PY
import json
for f in ("pics","pics2","mic3"):
    try:
        j=json.load(open("$O/%s_$s.json"%f))
        st=j.get("roofline",{}).get("stages_ms") or j.get("stages_ms")
        print("  ",f,j.get("ms_per_step"),st)
    except Exception as e: print("  ",f,"ERR",e)
PY
done
