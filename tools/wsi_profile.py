"""Wall time of a level-0 WSI tile-batch decode vs. the kernels in it (run under ncu for the kernel list)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mic = importlib.import_module("medical-image-codec_b200")
synth = importlib.import_module("medical-image-codec_b200.synth")
W, H = 8192, 6144
rgb = synth.wsi_region(11, 14000, 20000, W, H, 100000, 80000)
blob = mic.CompressWSI(rgb, W, H)
hdr = mic.ReadWSIHeader(blob)
tiles = [(0, tx, ty) for ty in range(hdr["Levels"][0][3]) for tx in range(hdr["Levels"][0][2])]
for it in range(3):
    t = time.perf_counter(); out = mic.DecompressWSITiles(blob, tiles); dt = time.perf_counter() - t
    print("decode %d tiles: %.1f ms, %.2f GB/s" % (len(tiles), dt * 1e3, rgb.nbytes / dt / 1e9))
