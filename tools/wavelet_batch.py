"""BASELINE configs[2] at its full batch: 64 synthetic 4096x3328 mammograms through WaveletV2 (5 levels), host calls."""
import importlib, os, sys, time
import numpy as np
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mic = importlib.import_module("medical-image-codec_b200")
synth = importlib.import_module("medical-image-codec_b200.synth")
n = int(os.environ.get("N", "64"))
with ThreadPoolExecutor(16) as ex:
    base = list(ex.map(lambda i: synth.mammo_image(1 + i).ravel(), range(min(n, 16))))
imgs = [base[i % len(base)] for i in range(n)]
raw = sum(i.nbytes for i in imgs)
for it in range(2):
    t = time.perf_counter(); blobs = mic.WaveletV2CompressBatch(imgs, 4096, 3328, [int(i.max()) for i in imgs], 5); te = time.perf_counter() - t
    t = time.perf_counter(); outs = mic.WaveletV2DecompressBatch(blobs); td = time.perf_counter() - t
    print("n=%d encode %.0f ms (%.2f GB/s)  decode %.0f ms (%.2f GB/s)  ratio %.2f" % (n, te * 1e3, raw / te / 1e9, td * 1e3, raw / td / 1e9, raw / sum(len(b) for b in blobs)))
ok = all(np.array_equal(np.asarray(o[0]).ravel(), i) for o, i in zip(outs[:4], imgs[:4]))
print("first images exact:", ok)
# the same decode into pinned host buffers that are reused across calls (what bench.py's e2e legs and tools/mic2_multi.py
# do): fresh numpy arrays above make the D2H copies fault 1.7 GB of pages in, which is the caller's allocator, not the codec
import ctypes as C
lib = mic.lib
views = [np.frombuffer(b, np.uint8) for b in blobs]
h_in = [lib.micgpu_host_alloc(v.size + 256) for v in views[: min(n, 16)]]
for p, v in zip(h_in, views):
    C.memmove(p, v.ctypes.data, v.size)
h_out = [lib.micgpu_host_alloc(4096 * 3328 * 2) for _ in range(n)]
bp = (C.c_void_p * n)(*[h_in[i % len(h_in)] for i in range(n)])
ln = (C.c_size_t * n)(*[views[i % len(h_in)].size for i in range(n)])
op = (C.c_void_p * n)(*h_out)
cp = (C.c_size_t * n)(*[4096 * 3328] * n)
rs, cs, st = (C.c_int * n)(), (C.c_int * n)(), (C.c_int * n)()
for it in range(3):
    t = time.perf_counter(); rc = lib.micgpu_wavelet_v2_decompress_batch(n, bp, ln, op, cp, rs, cs, st); td = time.perf_counter() - t
    assert rc == 0, rc
    print("n=%d decode into pinned, reused buffers %.0f ms (%.2f GB/s)" % (n, td * 1e3, raw / td / 1e9))
o0 = np.ctypeslib.as_array(C.cast(h_out[n - 1], C.POINTER(C.c_uint16)), shape=(4096 * 3328,))
print("last image exact:", np.array_equal(o0, imgs[n - 1]))
