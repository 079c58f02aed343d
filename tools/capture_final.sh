#!/bin/bash
# Reduced end-of-round capture (the kernels of the 8-state step are unchanged since tools/capture_round.sh ran; what moved
# is the thread-per-unit ANS kernel and the Huffman back end).  Writes gpurun_out/final3/; the .ncu-rep files are
# exported to CSV off the GPU box (tools/ncu_summary.py reads either).
set -u
O=gpurun_out/final3
mkdir -p $O
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm,clocks.max.mem,pcie.link.gen.max,pcie.link.width.max --format=csv > $O/box.txt 2>&1
nproc >> $O/box.txt
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1
timeout 400 python bench.py > $O/bench.json 2> $O/bench.err
timeout 200 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err
timeout 120 python bench.py --quick --steps 2 --warmup 3 > $O/quick.json 2>/dev/null && \
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv \
      python bench.py --quick --steps 2 --warmup 3 > $O/ncu_launches.log 2>&1
MICGPU_PARTS=1 timeout 120 python bench.py --quick --nstates 2 --steps 1 --warmup 3 > /dev/null 2>&1 && \
  MICGPU_PARTS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_ans_decode_serial" -s 4 -c 1 \
      -o $O/prof_pics8_2state python bench.py --quick --nstates 2 --steps 1 --warmup 3 > $O/ncu_pics8_2state.log 2>&1
timeout 120 python tools/huff_bench.py > $O/huff_bench.txt 2>&1 && \
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_huff_decode" -s 2 -c 1 \
      -o $O/prof_huff python tools/huff_bench.py > $O/ncu_huff.log 2>&1
N=64 timeout 200 python tools/wavelet_batch.py > $O/wavelet_x64.txt 2>&1
timeout 200 python tools/mic2_multi.py > $O/mic2_96.json 2> $O/mic2_96.err
timeout 150 python bench.py --nstates 4 --no-extra > $O/bench_4state.json 2> $O/bench_4state.err
du -sh $O; tail -2 $O/pytest_gpu.txt; tail -1 $O/smoke.txt; cut -c1-400 $O/bench.json
