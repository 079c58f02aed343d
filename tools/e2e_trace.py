"""Timeline of one micgpu_pics_decompress_batch call (MICGPU_TRACE=1): where the end-to-end time goes."""
import ctypes as C, importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MICGPU_TRACE"] = "1"
sys.argv = [sys.argv[0]]
import bench
api = importlib.import_module("medical-image-codec_b200.api")
n = int(os.environ.get("N", "256"))
blobs, raw_per, _ = bench.make_inputs(n, 32, 8, 1, use_gpu=True)
offs, tot = [], 0
for b in blobs:
    offs.append(tot); tot += (len(b) + 63) & ~63
hc = api.lib.micgpu_host_alloc(tot + 256); ho = api.lib.micgpu_host_alloc(n * raw_per)
h_comp = np.ctypeslib.as_array(C.cast(hc, C.POINTER(C.c_uint8)), shape=(tot + 256,))
for b, o in zip(blobs, offs): h_comp[o:o + len(b)] = np.frombuffer(b, np.uint8)
bp = (C.c_void_p * n)(*[hc + o for o in offs]); ln = (C.c_size_t * n)(*[len(b) for b in blobs])
op = (C.c_void_p * n)(*[ho + i * raw_per for i in range(n)]); cp = (C.c_size_t * n)(*([raw_per // 2] * n))
f = api.lib.micgpu_pics_decompress_batch
for it in range(3):
    t = time.perf_counter()
    rc = f(n, bp, ln, op, cp, None)
    print("call %d rc=%d %.2f ms" % (it, rc, (time.perf_counter() - t) * 1e3), file=sys.stderr)
