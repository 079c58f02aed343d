"""Wall time of one batched PICS-8 encode call (pageable numpy inputs) for the kernel-vs-host split; run under ncu for kernel times."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mic = importlib.import_module("medical-image-codec_b200")
synth = importlib.import_module("medical-image-codec_b200.synth")
W, H, n = 2577, 2048, int(os.environ.get("N", "16"))
imgs = [synth.xr_image(1 + i, W, H).ravel() for i in range(n)]
mx = [int(p.max()) for p in imgs]
for it in range(3):
    t = time.perf_counter()
    out = mic.CompressParallelStripsBatch(imgs, W, H, mx, 8, 8)
    dt = time.perf_counter() - t
    print("encode call %d: %.1f ms, %.2f GB/s raw, ratio %.3f" % (it, dt * 1e3, n * W * H * 2 / dt / 1e9, n * W * H * 2 / sum(len(o) for o in out)))
