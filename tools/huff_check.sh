# parity of the Huffman back end (self-synchronising and one-thread decode), the whole GPU suite, 4-state wavelet streams on either ANS kernel
echo "huff: $(timeout 300 python -m pytest tests/test_gpu_huff.py -x -q 2>&1 | tail -15)"
echo "huff serial: $(MICGPU_HUFF_SERIAL=1 timeout 300 python -m pytest tests/test_gpu_huff.py -x -q 2>&1 | tail -3)"
echo "all: $(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)"
for m in 2 4; do echo "wavelet x16 maxn=$m: $(MICGPU_K2_SERIAL_MAXN=$m N=16 python tools/wavelet_batch.py 2>&1 | tail -2 | head -1)"; done
