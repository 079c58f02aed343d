"""BASELINE configs[3]: MIC2 multi-frame decode (independent and temporal ZigZag modes) of a synthetic 2457x1996x96
16-bit tomosynthesis stack in ONE library call spread over the visible GPUs (micgpu_init + micgpu_mic2_decompress).

    python tools/mic2_multi.py [--frames 96] [--gpus N]      (N = 0: every visible device)

Prints one JSON object: per mode the host-call GB/s (pinned buffers, H2D/D2H inside), bit-exactness against the source
stack, and the compressed size.  The stack is encoded by the product's CUDA encoder (CompressMultiFrame)."""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=96)
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import __graft_entry__ as g

    g.build()
    mic = importlib.import_module("medical-image-codec_b200")
    synth = importlib.import_module("medical-image-codec_b200.synth")
    api = mic.api
    rows, cols = 2457, 1996
    t0 = time.time()
    st = synth.tomo_stack(7, a.frames, rows, cols)
    t_gen = time.time() - t0
    have = api.lib.micgpu_device_count()
    n = have if a.gpus <= 0 else min(a.gpus, have)
    out = {"workload": f"MIC2 {a.frames} frames of {cols}x{rows} 10-bit, synthetic tomosynthesis stack", "gpus": n, "gen_s": round(t_gen, 1)}
    fpx = rows * cols
    raw = a.frames * fpx * 2
    h_out = api.lib.micgpu_host_alloc(raw)
    hout = np.ctypeslib.as_array(C.cast(h_out, C.POINTER(C.c_uint16)), shape=(a.frames * fpx,))
    for temporal in (False, True):
        t0 = time.time()
        blob = mic.CompressMultiFrame(st.ravel(), cols, rows, 1023, temporal)
        t_enc = time.time() - t0
        h_in = api.lib.micgpu_host_alloc(len(blob) + 256)
        hin = np.ctypeslib.as_array(C.cast(h_in, C.POINTER(C.c_uint8)), shape=(len(blob),))
        hin[:] = np.frombuffer(blob, np.uint8)
        res = {}
        for devs in ([0], list(range(n))) if n > 1 else ([0],):
            mic.Init(devs)
            w, h, fr, tp = C.c_int(), C.c_int(), C.c_int(), C.c_int()

            def call():
                rc = api.lib.micgpu_mic2_decompress(h_in, len(blob), h_out, a.frames * fpx, C.byref(w), C.byref(h), C.byref(fr), C.byref(tp))
                if rc != 0:
                    raise RuntimeError(f"micgpu_mic2_decompress rc={rc}: {api.last_error()}")

            hout[:] = 0
            call()
            exact = bool(np.array_equal(hout.reshape(st.shape), st))
            t0 = time.perf_counter()
            for _ in range(a.reps):
                call()
            dt = (time.perf_counter() - t0) / a.reps
            res[f"{len(devs)}gpu"] = {"GBps": round(raw / dt / 1e9, 3), "ms": round(dt * 1e3, 1), "bit_exact": exact}
            api.lib.micgpu_shutdown()
        out["temporal" if temporal else "independent"] = {"compressed_MB": round(len(blob) / 1e6, 1), "ratio": round(raw / len(blob), 3),
                                                          "encode_host_call_s": round(t_enc, 2), **res}
        api.lib.micgpu_host_free(h_in)
    api.lib.micgpu_host_free(h_out)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
