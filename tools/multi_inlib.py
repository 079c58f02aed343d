"""In-library multi-GPU: ONE process, ONE call per container, micgpu_init(device list).

    python tools/multi_inlib.py [--images 256] [--side 32768] [--frames 96] [--skip-mic2]

For 1 device and for every visible device: host-call throughput (pinned host buffers, H2D / D2H inside the call) of
  * micgpu_pics_decompress_batch      BASELINE configs[1]: PICS-8 8-state, 2577x2048, `images` radiographs
  * micgpu_wsi_decompress_tile_range  BASELINE configs[4] shape: level 0 of a side x side window, 256x256 RGB8 tiles
  * micgpu_mic2_decompress            BASELINE configs[3]: 96 frames of 2457x1996, independent and temporal
with the outputs of the N-device call compared byte for byte with the 1-device call and spot-checked against the source
pixels, and the busy fraction of every GPU sampled by nvidia-smi while the N-device calls run.  Prints one JSON object."""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


class Busy:
    """utilization.gpu of every device, sampled every 100 ms"""

    def __init__(self):
        self.rows = []
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=index,utilization.gpu", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.rows.append(l.strip()) for l in self.p.stdout], daemon=True).start()
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return None
        time.sleep(0.2)
        self.p.terminate()
        mx = {}
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) == 2 and f[0].isdigit() and f[1].isdigit():
                mx[int(f[0])] = max(mx.get(int(f[0]), 0), int(f[1]))
        return {"max_utilization_pct": [mx[k] for k in sorted(mx)], "gpus_active": sum(1 for v in mx.values() if v > 0)}


def timed(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=256)
    ap.add_argument("--distinct", type=int, default=16)
    ap.add_argument("--side", type=int, default=32768)
    ap.add_argument("--frames", type=int, default=96)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--skip-mic2", action="store_true")
    a = ap.parse_args()
    import __graft_entry__ as g

    g.build()
    mic = importlib.import_module("medical-image-codec_b200")
    synth = importlib.import_module("medical-image-codec_b200.synth")
    api = mic.api
    lib = api.lib
    ndev = lib.micgpu_device_count()
    out = {"devices": ndev, "host_threads": len(os.sched_getaffinity(0))}
    dev_sets = [[0]] + ([list(range(ndev))] if ndev > 1 else [])

    def pinned(nbytes, dtype):
        p = lib.micgpu_host_alloc(nbytes)
        if not p:
            raise SystemExit("pinned allocation failed: " + api.last_error())
        return p, np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,)).view(dtype)

    # ---- PICS batch ----------------------------------------------------------------------------------------------------
    W, H = 2577, 2048
    imgs = [synth.xr_image(1 + i, W, H).ravel() for i in range(a.distinct)]
    uniq = mic.CompressParallelStripsBatch(imgs, W, H, [int(p.max()) for p in imgs], 8, 8)
    blobs = [uniq[i % a.distinct] for i in range(a.images)]
    offs, tot = [], 0
    for b in blobs:
        offs.append(tot)
        tot += (len(b) + 63) & ~63
    pin_c, h_comp = pinned(tot + 256, np.uint8)
    raw = a.images * W * H * 2
    pin_o, h_out = pinned(raw, np.uint16)
    for b, o in zip(blobs, offs):
        h_comp[o:o + len(b)] = np.frombuffer(b, np.uint8)
    n = a.images
    bp = (C.c_void_p * n)(*[pin_c + o for o in offs])
    ln = (C.c_size_t * n)(*[len(b) for b in blobs])
    op = (C.c_void_p * n)(*[pin_o + i * W * H * 2 for i in range(n)])
    cp = (C.c_size_t * n)(*[W * H] * n)
    st = (C.c_int * n)()

    def pics_call():
        rc = lib.micgpu_pics_decompress_batch(n, bp, ln, op, cp, st)
        if rc:
            raise RuntimeError(f"micgpu_pics_decompress_batch rc={rc}: {api.last_error()}")

    res, crc0 = {}, None
    for devs in dev_sets:
        mic.Init(devs)
        h_out[:] = 0
        busy = Busy()
        if len(devs) > 1:
            busy.start()
        dt = timed(pics_call, a.reps)
        b = busy.stop() if len(devs) > 1 else None
        crc = zlib.crc32(h_out.view(np.uint8))
        crc0 = crc if crc0 is None else crc0
        exact = all(np.array_equal(h_out[i * W * H:(i + 1) * W * H], imgs[i % a.distinct]) for i in (0, n // 2, n - 1))
        res[f"{len(devs)}gpu"] = {"GBps": round(raw / dt / 1e9, 2), "ms": round(dt * 1e3, 1), "source_pixels_exact": bool(exact),
                                  "same_bytes_as_1gpu": crc == crc0, **({"busy": b} if b else {})}
        lib.micgpu_shutdown()
    out["pics8_batch"] = {"workload": f"PICS-8 8-state, {n} x {W}x{H}", "compressed_MB": round(tot / 1e6, 1), **res}
    lib.micgpu_host_free(pin_c)
    lib.micgpu_host_free(pin_o)
    print(json.dumps(out), file=sys.stderr, flush=True)

    # ---- MIC3 tile range -----------------------------------------------------------------------------------------------
    import mic3_bench as m3

    src = m3.make_source(11)
    src_blob = np.frombuffer(mic.CompressWSI(src.ravel(), m3.SRC_SIDE, m3.SRC_SIDE, 3, 8, m3.TILE, m3.TILE, 1), np.uint8)
    blob = m3.assemble(src_blob, m3.SRC_SIDE, a.side)
    n_side = a.side // m3.TILE
    n_tiles = n_side * n_side
    tb = m3.TILE * m3.TILE * 3
    pin_c, h_comp = pinned(blob.size + 256, np.uint8)
    h_comp[:blob.size] = blob
    pin_o, h_out = pinned(n_tiles * tb, np.uint8)
    tst = (C.c_int * n_tiles)()

    def wsi_call():
        rc = lib.micgpu_wsi_decompress_tile_range(pin_c, blob.size, C.c_uint64(0), C.c_uint64(n_tiles), pin_o, n_tiles * tb, tst)
        if rc:
            raise RuntimeError(f"micgpu_wsi_decompress_tile_range rc={rc}: {api.last_error()}")

    res, crc0 = {}, None
    s = m3.SRC_SIDE // m3.TILE
    for devs in dev_sets:
        mic.Init(devs)
        h_out[:] = 0
        busy = Busy()
        if len(devs) > 1:
            busy.start()
        dt = timed(wsi_call, a.reps)
        b = busy.stop() if len(devs) > 1 else None
        crc = zlib.crc32(h_out)
        crc0 = crc if crc0 is None else crc0
        exact = True
        for (tx, ty) in ((0, 0), (n_side // 2, n_side // 3), (n_side - 1, n_side - 1)):
            idx = ty * n_side + tx
            sx, sy = (tx % s) * m3.TILE, (ty % s) * m3.TILE
            exact &= bool(np.array_equal(h_out[idx * tb:(idx + 1) * tb].reshape(m3.TILE, m3.TILE, 3), src[sy:sy + m3.TILE, sx:sx + m3.TILE]))
        res[f"{len(devs)}gpu"] = {"GBps": round(n_tiles * tb / dt / 1e9, 2), "ms": round(dt * 1e3, 1), "source_pixels_exact": exact,
                                  "same_bytes_as_1gpu": crc == crc0, **({"busy": b} if b else {})}
        lib.micgpu_shutdown()
    out["mic3_tile_range"] = {"workload": f"MIC3 level 0, {a.side}x{a.side} RGB8, {n_tiles} tiles of 256x256", "compressed_MB": round(blob.size / 1e6, 1), **res}
    lib.micgpu_host_free(pin_c)
    lib.micgpu_host_free(pin_o)
    print(json.dumps(out), file=sys.stderr, flush=True)

    # ---- MIC2 ------------------------------------------------------------------------------------------------------------
    if not a.skip_mic2:
        rows, cols = 2457, 1996
        stk = synth.tomo_stack(7, a.frames, rows, cols)
        fpx = rows * cols
        raw = a.frames * fpx * 2
        pin_o, h_out = pinned(raw, np.uint16)
        for temporal in (False, True):
            blob = mic.CompressMultiFrame(stk.ravel(), cols, rows, 1023, temporal)
            pin_c, h_comp = pinned(len(blob) + 256, np.uint8)
            h_comp[:len(blob)] = np.frombuffer(blob, np.uint8)
            w, h, fr, tp = C.c_int(), C.c_int(), C.c_int(), C.c_int()

            def mic2_call():
                rc = lib.micgpu_mic2_decompress(pin_c, len(blob), pin_o, a.frames * fpx, C.byref(w), C.byref(h), C.byref(fr), C.byref(tp))
                if rc:
                    raise RuntimeError(f"micgpu_mic2_decompress rc={rc}: {api.last_error()}")

            res = {}
            for devs in dev_sets:
                mic.Init(devs)
                h_out[:] = 0
                busy = Busy()
                if len(devs) > 1:
                    busy.start()
                dt = timed(mic2_call, max(1, a.reps - 1))
                b = busy.stop() if len(devs) > 1 else None
                exact = bool(np.array_equal(h_out.reshape(stk.shape), stk))
                res[f"{len(devs)}gpu"] = {"GBps": round(raw / dt / 1e9, 2), "ms": round(dt * 1e3, 1), "bit_exact": exact, **({"busy": b} if b else {})}
                lib.micgpu_shutdown()
            out["mic2_temporal" if temporal else "mic2_independent"] = {
                "workload": f"MIC2 {a.frames} frames of {cols}x{rows}", "compressed_MB": round(len(blob) / 1e6, 1), **res}
            lib.micgpu_host_free(pin_c)
        lib.micgpu_host_free(pin_o)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
