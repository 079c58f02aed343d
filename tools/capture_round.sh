#!/bin/bash
# Collects the evidence committed under profiles/ (run on the GPU box through gpurun; writes gpurun_out/final/).
# One profiler per command, each only after the same command exited 0 without it.
set -u
O=gpurun_out/final
mkdir -p $O
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm,clocks.max.mem,pcie.link.gen.max,pcie.link.width.max --format=csv > $O/box.txt 2>&1
nproc >> $O/box.txt
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1
python bench.py > $O/bench_8state.json 2> $O/bench_8state.err
python bench.py --nstates 4 > $O/bench_4state.json 2> $O/bench_4state.err
python bench.py --nstates 2 > $O/bench_2state.json 2> $O/bench_2state.err
python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err
python bench.py --quick --steps 2 --warmup 3 > $O/quick.json 2>/dev/null && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
      python bench.py --quick --steps 2 --warmup 3 > $O/ncu_launches.log 2>&1
python bench.py --quick --steps 1 --warmup 1 > /dev/null 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"k_build_tables|k_ans_decode|k_rle_expand|k_delta_wavefront" -c 4 \
      -o $O/prof_decode python bench.py --quick --steps 1 --warmup 1 > $O/ncu_full.log 2>&1
python tools/bench_configs.py > $O/configs.json 2> $O/configs.err
tail -2 $O/pytest_gpu.txt; cat $O/smoke.txt | tail -1; cat $O/bench_8state.json | cut -c1-400
