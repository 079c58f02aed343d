#!/usr/bin/env python
"""Host-call throughput of the other BASELINE.json configs (configs[0], [2], [3], [4]) through the Go-named API.
These are parity-test shapes, not the bench line (bench.py measures configs[1]); numbers go to DESIGN.md for context.
GB/s = raw pixel bytes / wall time of the call (pageable host buffers, copies included)."""
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


def timed(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    return (time.perf_counter() - t0) / reps, r


out = {}
# configs[0]: CT 512x512, 1-state
ct = np.fromfile(os.path.join(ROOT, "tests", "golden", "CT_512_512_image.bin"), np.uint16)
te, blob = timed(lambda: mic.CompressSingleFrame(ct, 512, 512, int(ct.max()), 1), 10)
td, _ = timed(lambda: mic.DecompressSingleFrame(blob, 512, 512), 10)
out["c1_ct512_1state"] = {"encode_GBps": ct.nbytes / te / 1e9, "decode_GBps": ct.nbytes / td / 1e9, "ratio": ct.nbytes / len(blob), "note": "single 512x512 unit, tableLog 16: latency bound"}
# configs[2]: mammo wavelet, 8 images
imgs = [synth.mammo_image(1 + i).ravel() for i in range(8)]
raw = sum(i.nbytes for i in imgs)
te, blobs = timed(lambda: mic.WaveletV2CompressBatch(imgs, 4096, 3328, [int(i.max()) for i in imgs], 5), 2)
td, _ = timed(lambda: mic.WaveletV2DecompressBatch(blobs), 2)
out["c3_mammo_wavelet_x8"] = {"encode_GBps": raw / te / 1e9, "decode_GBps": raw / td / 1e9, "ratio": raw / sum(len(b) for b in blobs)}
# configs[3]: tomo MIC2, 24 frames
st = synth.tomo_stack(7, 24)
for temporal in (False, True):
    te, blob = timed(lambda: mic.CompressMultiFrame(st, 1996, 2457, 1023, temporal), 2)
    td, _ = timed(lambda: mic.DecompressMultiFrame(blob), 2)
    out[f"c4_tomo_x24_temporal{int(temporal)}"] = {"encode_GBps": st.nbytes / te / 1e9, "decode_GBps": st.nbytes / td / 1e9, "ratio": st.nbytes / len(blob)}
# configs[4]: WSI window 8192x6144
W, H = 8192, 6144
rgb = synth.wsi_region(11, 14000, 20000, W, H, 100000, 80000)
te, blob = timed(lambda: mic.CompressWSI(rgb, W, H), 1)
hdr = mic.ReadWSIHeader(blob)
tiles = [(0, tx, ty) for ty in range(hdr["Levels"][0][3]) for tx in range(hdr["Levels"][0][2])]
td, _ = timed(lambda: mic.DecompressWSITiles(blob, tiles), 2)
first0, n0 = hdr["Levels"][0][4], hdr["Levels"][0][2] * hdr["Levels"][0][3]
td2, _ = timed(lambda: mic.DecompressWSITileRange(blob, first0, n0), 2)
td3, _ = timed(lambda: mic.DecompressWSITileRange(blob, 0, hdr["TotalTiles"]), 2)
out["c5_wsi_8192x6144"] = {"encode_GBps": rgb.nbytes / te / 1e9, "decode_level0_GBps": rgb.nbytes / td / 1e9,
                           "decode_level0_tile_range_GBps": n0 * 196608 / td2 / 1e9, "decode_all_levels_tile_range_GBps": hdr["TotalTiles"] * 196608 / td3 / 1e9,
                           "ratio": rgb.nbytes / len(blob), "tiles": len(tiles), "total_tiles": hdr["TotalTiles"],
                           "note": "tile_range = micgpu_wsi_decompress_tile_range (one H2D, one D2H, full tiles); pageable buffers"}
print(json.dumps({k: {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()} for k, v in out.items()}, indent=1))
