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
