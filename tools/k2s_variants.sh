# A/B of K2s builds (tools/probe/libmicgpu_<variant>.so, built with MICGPU_NVCC_EXTRA=-D... from a tree that carries the
# variants as macros): parity of each, then the kernel time on 2-state strips, 4-state strips, MIC3 planes (4096 tiles)
# and MIC2 frames (12 frames: independent, temporal).  Results of the session that added the packed stores (ms):
#   variant                                   2-state  4-state  MIC3   MIC2 ind / temporal
#   one store per round (before)               11.34    10.43   2.346   99.1 / 123.3
#   16 symbols as two 16 B stores (kept)       11.37    10.33   2.351   84.7 / 104.8
#   + byte-offset addressing of the ring word  11.59    10.63   2.391   81.4 / 100.4   (one instruction less per pair, slower in full warps)
#   byte-offset addressing alone               11.90    10.50   2.458  110.3 / 136.3
#   + 32-bit window rebuilt per pair instead   11.73    10.49   2.415   83.6 / 103.7   (no low half to shift: two shifts less per pair)
#     of a 64-bit buffer
# i.e. fewer instructions did not buy time where a warp carries several units; only the packed stores were kept.
for v in ${VARIANTS:-base win}; do
cp tools/probe/libmicgpu_$v.so medical-image-codec_b200/libmicgpu.so
echo "$v parity: $(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1)"
echo "$v 2-state: $(python bench.py --quick --no-extra --nstates 2 --steps 5 --warmup 3 2>/dev/null | python -c 'import json,sys; j=json.load(sys.stdin); print(j["roofline"]["stages_ms"]["k_ans_decode_serial<2>"])')"
echo "$v 4-state: $(python bench.py --quick --no-extra --nstates 4 --steps 5 --warmup 3 2>/dev/null | python -c 'import json,sys; j=json.load(sys.stdin); print(j["roofline"]["stages_ms"]["k_ans_decode_serial<4>"])')"
echo "$v mic3: $(python tools/mic3_bench.py --side 16384 --steps 3 --warmup 1 --no-e2e --no-cpu 2>/dev/null | python -c 'import json,sys; j=json.load(sys.stdin); print(j["stages_ms"]["k_ans_decode_serial<2>"])')"
echo "$v mic2: $(FRAMES=12 python tools/mic2_profile.py 2>&1 | grep -E "^independent|^temporal" | grep -o "k_ans_decode_serial<2>', [0-9.]*" | tr '\n' ' ')"
done
