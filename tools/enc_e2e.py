"""End-to-end batched PICS-8 encode through the C ABI with pinned host buffers (the encode twin of bench.py's e2e leg)."""
import ctypes as C, importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
api = importlib.import_module("medical-image-codec_b200.api")
synth = importlib.import_module("medical-image-codec_b200.synth")
from concurrent.futures import ThreadPoolExecutor
W, H, n, distinct = 2577, 2048, int(os.environ.get("N", "256")), 32
npx = W * H
with ThreadPoolExecutor(16) as ex:
    imgs = list(ex.map(lambda i: synth.xr_image(1 + i, W, H).ravel(), range(distinct)))
hin = api.lib.micgpu_host_alloc(n * npx * 2)
cap = npx * 2 + 65536
hout = api.lib.micgpu_host_alloc(n * cap)
a_in = np.ctypeslib.as_array(C.cast(hin, C.POINTER(C.c_uint16)), shape=(n * npx,))
for i in range(n):
    a_in[i * npx:(i + 1) * npx] = imgs[i % distinct]
px = (C.c_void_p * n)(*[hin + i * npx * 2 for i in range(n)])
mv = (C.c_uint16 * n)(*[int(imgs[i % distinct].max()) for i in range(n)])
outs = (C.c_void_p * n)(*[hout + i * cap for i in range(n)])
caps = (C.c_size_t * n)(*([cap] * n))
lens = (C.c_size_t * n)()
st = (C.c_int * n)()
for it in range(4):
    t = time.perf_counter()
    rc = api.lib.micgpu_pics_compress_batch(n, px, W, H, mv, 8, 8, outs, caps, lens, st)
    dt = time.perf_counter() - t
    tot = sum(lens)
    print("encode e2e call %d rc=%d: %.1f ms, %.2f GB/s of raw pixels, ratio %.3f" % (it, rc, dt * 1e3, n * npx * 2 / dt / 1e9, n * npx * 2 / max(tot, 1)))
# byte identity of image 0 against the single-context path
one = api.CompressParallelStrips(imgs[0], W, H, int(imgs[0].max()), 8, 8)
got = bytes(np.ctypeslib.as_array(C.cast(hout, C.POINTER(C.c_uint8)), shape=(lens[0],)))
print("image 0 identical to the one-shot call:", got == one)
