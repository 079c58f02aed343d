"""MIC3 (BASELINE configs[4]) measurement leg: level-0 tile decode of a large procedural slide window.

Used by bench.py (extra key `mic3` of the JSON line) and runnable on its own:
    python tools/mic3_bench.py [--side 32768] [--steps 5]

Workload: a `side` x `side` RGB8 window made of 256x256 tiles (side 32768 -> 16384 tiles, 3.2 GB of pixels).  The
window is the procedural slide of SURVEY 8(d) (white background, elliptical H&E tissue ~35 % of the area, nuclei)
generated at 8192 x 8192 and repeated 4 x 4 -- the same "distinct inputs cycled, separate buffers" device as the PICS
leg -- encoded by the product's own CUDA encoder (CompressWSI, one pyramid level) and assembled into ONE MIC3
container whose tile table holds every tile separately (wsiformat.go:98-164).
"""
from __future__ import annotations

import ctypes as C
import importlib
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "medical-image-codec_b200"
TILE = 256
SRC_SIDE = 8192


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def make_source(seed: int, side: int = SRC_SIDE) -> np.ndarray:
    synth = importlib.import_module(PKG + ".synth")
    bands = 16
    rows = side // bands
    with ThreadPoolExecutor(max_workers=min(bands, host_threads())) as ex:
        parts = list(ex.map(lambda b: synth.wsi_region(seed, 0, b * rows, side, rows, side, side), range(bands)))
    return np.concatenate(parts, axis=0)


def parse_tiles(blob: np.ndarray):
    """-> (list of (offset, length) absolute in blob, n_levels)"""
    nlv = int(blob[28]) | (int(blob[29]) << 8)
    total = int.from_bytes(blob[32:40].tobytes(), "little")
    table_off = 48 + 20 * nlv
    data_off = table_off + 16 * total
    tab = np.frombuffer(blob[table_off:data_off].tobytes(), dtype="<u8").reshape(total, 2)
    return [(data_off + int(o), int(l)) for o, l in tab], nlv


def assemble(src_blob: np.ndarray, src_side: int, side: int) -> np.ndarray:
    """MIC3 container of a side x side window whose tile (tx, ty) is tile (tx % s, ty % s) of the source container."""
    tiles, nlv = parse_tiles(src_blob)
    s = src_side // TILE
    n = side // TILE
    assert nlv == 1 and len(tiles) == s * s
    order = [(ty % s) * s + (tx % s) for ty in range(n) for tx in range(n)]
    lens = np.array([tiles[i][1] for i in order], dtype=np.uint64)
    offs = np.zeros(len(order), dtype=np.uint64)
    offs[1:] = np.cumsum(lens)[:-1]
    hdr = bytearray(48)
    hdr[0:4] = b"MIC3"
    hdr[4:8] = (1).to_bytes(4, "little")
    hdr[8:12] = side.to_bytes(4, "little")
    hdr[12:16] = side.to_bytes(4, "little")
    hdr[16:20] = TILE.to_bytes(4, "little")
    hdr[20:24] = TILE.to_bytes(4, "little")
    hdr[24:26] = (3).to_bytes(2, "little")
    hdr[26] = 8
    hdr[27] = src_blob[27]
    hdr[28:30] = (1).to_bytes(2, "little")
    hdr[32:40] = len(order).to_bytes(8, "little")
    lvl = b"".join(int(v).to_bytes(4, "little") for v in (side, side, n, n, 0))
    table = np.empty((len(order), 2), dtype="<u8")
    table[:, 0] = offs
    table[:, 1] = lens
    total = 48 + 20 + table.nbytes + int(lens.sum())
    out = np.empty(total, np.uint8)
    p = 0
    for chunk in (np.frombuffer(bytes(hdr), np.uint8), np.frombuffer(lvl, np.uint8), table.view(np.uint8).ravel()):
        out[p:p + chunk.size] = chunk
        p += chunk.size
    for i in order:
        o, l = tiles[i]
        out[p:p + l] = src_blob[o:o + l]
        p += l
    assert p == total
    return out


def run(mic, torch, steps: int, warmup: int, side: int = 32768, seed: int = 11, e2e: bool = True, cpu: bool = True, log=print):
    api = mic.api
    t0 = time.time()
    src = make_source(seed)
    t_gen = time.time() - t0
    t0 = time.time()
    src_blob = np.frombuffer(mic.CompressWSI(src.ravel(), SRC_SIDE, SRC_SIDE, 3, 8, TILE, TILE, 1), np.uint8)
    t_enc = time.time() - t0
    blob = assemble(src_blob, SRC_SIDE, side)
    n_side = side // TILE
    n_tiles = n_side * n_side
    log(f"[mic3] source {SRC_SIDE}^2 generated in {t_gen:.1f}s, encoded in {t_enc:.2f}s ({src_blob.size / 1e6:.1f} MB, ratio "
        f"{src.size / src_blob.size:.2f}); container {side}^2: {n_tiles} tiles, {blob.size / 1e6:.1f} MB")

    # ---- device-resident arm -------------------------------------------------------------------------------
    t0 = time.time()
    plan = api.WsiPlan(blob, 0, n_tiles, torch.cuda.current_device())
    t_plan = time.time() - t0
    d_span = torch.empty(plan.span_len + 256, dtype=torch.uint8, device="cuda")
    d_span[: plan.span_len].copy_(torch.from_numpy(blob[plan.span_off: plan.span_off + plan.span_len]))
    d_out = torch.empty(plan.out_bytes, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        plan.run_device(d_span.data_ptr(), d_out.data_ptr(), stream)

    for _ in range(max(warmup, 1)):
        step()
    assert not any(plan.status(stream)), "tile decode failed"
    # spot checks: decoded tiles equal the generator's pixels and the CPU oracle's decode of the same container
    from oracle.oracle import Oracle

    orc = Oracle()
    s = SRC_SIDE // TILE
    checks = [(0, 0), (n_side // 2, n_side // 2), (n_side - 1, n_side - 1), ((s + 13) % n_side, (2 * s + 17) % n_side), (n_side // 2 + 3, 5), (17, n_side - 9)]
    blob_b = blob.tobytes() if blob.size < (1 << 31) else None
    for tx, ty in checks:
        idx = ty * n_side + tx
        got = d_out[idx * plan.tile_bytes:(idx + 1) * plan.tile_bytes].cpu().numpy().reshape(TILE, TILE, 3)
        sx, sy = (tx % s) * TILE, (ty % s) * TILE
        assert np.array_equal(got, src[sy:sy + TILE, sx:sx + TILE]), f"tile ({tx},{ty}) differs from the source pixels"
        if blob_b is not None:
            ref, rw, rh = orc.wsi_decompress_tile(blob_b, 0, tx, ty)
            assert (rw, rh) == (TILE, TILE) and np.array_equal(got.ravel(), np.asarray(ref, np.uint8)), f"tile ({tx},{ty}) differs from the oracle"
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms_dev = e0.elapsed_time(e1) / steps
    launches = plan.last_launches
    plan.set_profiling(True)
    acc = {}
    for _ in range(2):
        step()
        for name, ms in plan.kernel_times():
            acc[name] = acc.get(name, 0.0) + ms / 2
    plan.set_profiling(False)
    out_bytes, span_len, n_units = plan.out_bytes, plan.span_len, plan.n_units
    res = {
        "workload": f"MIC3 level-0 tile decode, {side}x{side} RGB8 procedural slide window ({n_tiles} tiles of 256x256, YCoCg-R planes, "
                    f"2-state FSE; one {SRC_SIDE}^2 source window repeated, every tile a separate blob)",
        "value": round(out_bytes / (ms_dev * 1e-3) / 1e9, 3), "unit": "GB/s (RGB bytes)", "ms_per_step": round(ms_dev, 4),
        "n_tiles": n_tiles, "n_units": n_units, "ratio": round(out_bytes / span_len, 3), "gpu_launches_per_step": launches,
        "plan_ms_once": round(t_plan * 1e3, 1), "stages_ms": {k: round(v, 4) for k, v in acc.items()},
        "parity": f"{len(checks)} tiles bit-exact vs generator pixels and vs the CPU oracle (restatement; MIC3 bytes are parity-unpinned vs Go)",
    }
    if acc:
        dom = max(acc.items(), key=lambda kv: kv[1])
        alg = span_len + out_bytes
        res["roofline"] = {"bound": "hbm", "kernel": dom[0], "kernel_ms": round(dom[1], 4), "algorithmic_bytes_per_launch": alg,
                           "achieved": round(alg / (dom[1] * 1e-3) / 1e9, 2), "pipeline_achieved": round(alg / (ms_dev * 1e-3) / 1e9, 2)}
    plan.close()
    del d_out, d_span
    torch.cuda.empty_cache()

    # ---- end to end: host buffers in and out through micgpu_wsi_decompress_tile_range -----------------------------
    if e2e:
        h_in = api.lib.micgpu_host_alloc(blob.size + 256)
        h_out = api.lib.micgpu_host_alloc(out_bytes)
        if h_in and h_out:
            hin = np.ctypeslib.as_array(C.cast(h_in, C.POINTER(C.c_uint8)), shape=(blob.size + 256,))
            hin[: blob.size] = blob
            st = (C.c_int * n_tiles)()

            def step_e2e():
                rc = api.lib.micgpu_wsi_decompress_tile_range(h_in, blob.size, C.c_uint64(0), C.c_uint64(n_tiles), h_out, out_bytes, st)
                if rc != 0:
                    raise RuntimeError("micgpu_wsi_decompress_tile_range rc=%d: %s" % (rc, api.last_error()))

            step_e2e()
            hout = np.ctypeslib.as_array(C.cast(h_out, C.POINTER(C.c_uint8)), shape=(out_bytes,))
            tx, ty = checks[1]
            idx = ty * n_side + tx
            sx, sy = (tx % s) * TILE, (ty % s) * TILE
            assert np.array_equal(hout[idx * TILE * TILE * 3:(idx + 1) * TILE * TILE * 3].reshape(TILE, TILE, 3), src[sy:sy + TILE, sx:sx + TILE])
            reps = max(1, min(steps, 3))
            t0 = time.perf_counter()
            for _ in range(reps):
                step_e2e()
            ms = (time.perf_counter() - t0) * 1e3 / reps
            res["e2e"] = {"value": round(out_bytes / (ms * 1e-3) / 1e9, 3), "unit": "GB/s (RGB bytes)", "ms_per_step": round(ms, 2),
                          "h2d_bytes_per_step": int(span_len), "d2h_bytes_per_step": int(out_bytes), "steps": reps,
                          "api": "micgpu_wsi_decompress_tile_range (pinned host buffers; plans the 16384 tiles inside the call)"}
        if h_in:
            api.lib.micgpu_host_free(h_in)
        if h_out:
            api.lib.micgpu_host_free(h_out)

    # ---- CPU baseline: the oracle's tile decoder on the host cores, bounded sample ---------------------------------
    if cpu:
        nthreads = host_threads()
        src_b = src_blob.tobytes()
        sample = [(i % s, (i * 7) % s) for i in range(4 * nthreads)]

        def work(t):
            o2 = Oracle()
            return o2.wsi_decompress_tile(src_b, 0, t[0], t[1])[1]

        with ThreadPoolExecutor(max_workers=nthreads) as ex:
            list(ex.map(work, sample[:nthreads]))
            reps, t0 = 0, time.perf_counter()
            while reps < 2 or (time.perf_counter() - t0 < 4.0 and reps < 40):
                list(ex.map(work, sample))
                reps += 1
            dt = (time.perf_counter() - t0) / reps
        res["cpu_baseline"] = {"value": round(len(sample) * TILE * TILE * 3 / dt / 1e9, 4), "unit": "GB/s (RGB bytes)", "cores": nthreads, "kind": "port",
                               "sample": f"{len(sample)} tiles x {reps} reps of the source window through the CPU oracle (C restatement of "
                                         f"DecompressWSITile, wsicompress.go:175-217; the reference's C twin has no MIC3 path), {nthreads} host threads"}
    return res


if __name__ == "__main__":
    import argparse
    import json

    import torch

    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=32768)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    import __graft_entry__ as g

    g.build()
    mic = importlib.import_module(PKG)
    print(json.dumps(run(mic, torch, a.steps, a.warmup, a.side, e2e=not a.no_e2e, cpu=not a.no_cpu, log=lambda *x: print(*x, file=sys.stderr))))
