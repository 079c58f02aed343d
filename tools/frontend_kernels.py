"""Runs every front-end kernel of the path once on BASELINE-shaped inputs (for the ncu capture in tools/capture_round.sh)
and prints CUDA-event-free wall numbers for context: WaveletV2 decode + encode (4096x3328), MIC2 temporal decode + encode
(2457x1996 frames), MIC3 encode of a 4096x4096 window (tile planes, plane statistics, 2x2 pyramid) and its tile decode (blit)."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

g.build()
mic = importlib.import_module("medical-image-codec_b200")
synth = importlib.import_module("medical-image-codec_b200.synth")
out = {}
imgs = [synth.mammo_image(1 + i).ravel() for i in range(2)]
t0 = time.perf_counter()
blobs = mic.WaveletV2CompressBatch(imgs, 4096, 3328, [int(i.max()) for i in imgs], 5)
out["wavelet_encode_s"] = round(time.perf_counter() - t0, 3)
t0 = time.perf_counter()
res = mic.WaveletV2DecompressBatch(blobs)
out["wavelet_decode_s"] = round(time.perf_counter() - t0, 3)
assert np.array_equal(res[0][0], imgs[0])
st = synth.tomo_stack(7, 6)
blob = mic.CompressMultiFrame(st, 1996, 2457, 1023, True)
frames, _ = mic.DecompressMultiFrame(blob)
assert np.array_equal(np.asarray(frames).reshape(st.shape), st)
rgb = synth.wsi_region(11, 20000, 20000, 4096, 4096, 100000, 80000)
wsi = mic.CompressWSI(rgb, 4096, 4096)
hdr = mic.ReadWSIHeader(wsi)
tiles = mic.DecompressWSITileRange(wsi, 0, hdr["TotalTiles"])
assert np.array_equal(tiles[0].reshape(256, 256, 3), rgb[:256, :256])
out["ok"] = True
print(json.dumps(out))
