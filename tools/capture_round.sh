#!/bin/bash
# Collects the evidence committed under profiles/ (run on the GPU box through gpurun; writes gpurun_out/final/).
# One profiler per command, each only after the same command exited 0 without it.
set -u
O=gpurun_out/final
mkdir -p $O
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm,clocks.max.mem,pcie.link.gen.max,pcie.link.width.max --format=csv > $O/box.txt 2>&1
nproc >> $O/box.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1
python bench.py > $O/bench.json 2> $O/bench.err
python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err
python bench.py --nstates 4 --no-extra > $O/bench_4state.json 2> $O/bench_4state.err
# launch list of two timed steps (cold-cache, serialised: compare shares, not absolutes)
python bench.py --quick --steps 2 --warmup 3 > $O/quick.json 2>/dev/null && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv \
      python bench.py --quick --steps 2 --warmup 3 > $O/ncu_launches.log 2>&1
# every decode kernel of one step, one launch each (MICGPU_PARTS=1), full set with source
MICGPU_PARTS=1 python bench.py --quick --steps 1 --warmup 3 > /dev/null 2>&1 && \
  MICGPU_PARTS=1 ncu --set full --clock-control none --import-source on \
      -k regex:"k_parse_ncount|k_build_dtable|k_ans_decode|k_rle_expand|k_delta_rowscan" -s 20 -c 5 \
      -o $O/prof_pics8 python bench.py --quick --steps 1 --warmup 3 > $O/ncu_pics8.log 2>&1
MICGPU_PARTS=1 python bench.py --quick --nstates 2 --steps 1 --warmup 3 > /dev/null 2>&1 && \
  MICGPU_PARTS=1 ncu --set full --clock-control none --import-source on -k regex:"k_ans_decode_serial" -s 4 -c 1 \
      -o $O/prof_pics8_2state python bench.py --quick --nstates 2 --steps 1 --warmup 3 > $O/ncu_pics8_2state.log 2>&1
# MIC3 tile workload (8192^2 window = 1024 tiles keeps the replays short)
MICGPU_PARTS=1 python tools/mic3_bench.py --side 8192 --steps 1 --warmup 1 --no-e2e --no-cpu > $O/mic3_8k.json 2>/dev/null && \
  MICGPU_PARTS=1 ncu --set full --clock-control none --import-source on \
      -k regex:"k_parse_ncount|k_build_dtable|k_ans_decode|k_rle_expand|k_delta_rowscan|k_tile_blit" -s 6 -c 6 \
      -o $O/prof_mic3 python tools/mic3_bench.py --side 8192 --steps 1 --warmup 1 --no-e2e --no-cpu > $O/ncu_mic3.log 2>&1
# front-end kernels (wavelet lifting, temporal, tile planes, pyramid): one launch each
python tools/frontend_kernels.py > $O/frontends.json 2> $O/frontends.err && \
  ncu --set full --clock-control none -k regex:"k_wt53|k_wavelet|k_temporal|k_tile_planes|k_tile_blit|k_downsample|k_plane|k_grad|k_row_costs" -c 90 \
      -o $O/prof_frontends python tools/frontend_kernels.py > $O/ncu_frontends.log 2>&1
# gpurun merges at most 64 MiB back: keep the raw-metric CSV of every capture, the reports only of the two main workloads
for r in prof_pics8 prof_pics8_2state prof_mic3 prof_frontends; do
  [ -f $O/$r.ncu-rep ] && ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null
done
rm -f $O/prof_frontends.ncu-rep $O/prof_pics8_2state.ncu-rep
python tools/bench_configs.py > $O/configs.json 2> $O/configs.err
N=64 python tools/wavelet_batch.py > $O/wavelet_x64.txt 2>&1
python tools/mic2_multi.py > $O/mic2_96.json 2> $O/mic2_96.err
du -sh $O; tail -2 $O/pytest_gpu.txt; cat $O/smoke.txt | tail -1; cat $O/bench.json | cut -c1-600
