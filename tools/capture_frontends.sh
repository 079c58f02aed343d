O=gpurun_out/final
mkdir -p $O
python -m pytest tests/test_gpu_wavelet.py tests/test_gpu_encode.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -1
python tools/frontend_kernels.py > $O/frontends.json 2> $O/frontends.err && \
  ncu --set full --clock-control none -k regex:"k_wt53|k_wavelet|k_temporal|k_tile_planes|k_tile_blit|k_downsample|k_plane|k_grad|k_row_costs" -c 90 \
      -o $O/prof_frontends python tools/frontend_kernels.py > $O/ncu_frontends.log 2>&1
ncu -i $O/prof_frontends.ncu-rep --page raw --csv > $O/prof_frontends.raw.csv 2>/dev/null
rm -f $O/prof_frontends.ncu-rep
N=64 python tools/wavelet_batch.py 2>&1 | tail -3 | tee $O/wavelet_x64.txt
