"""BASELINE configs[4] at its named scale: a 100000 x 80000 RGB8 slide, 256x256 tiles, 6 pyramid levels = 163 487 tiles,
decoded through micgpu_wsi_decompress_tile_range over every visible GPU (micgpu_init), one call per 16 384-tile range.

    python tools/mic3_fullscale.py [--width 100000] [--height 80000] [--levels 6] [--chunk 16384]

The slide cannot be materialised (24 GB of pixels at level 0): the procedural slide of SURVEY 8(d) is generated at
8192 x 8192 with 6 pyramid levels, encoded by the product's CUDA encoder (CompressWSI), and a container of the full
geometry is assembled whose tile (tx, ty) of level l is tile (tx mod s_l, ty mod s_l) of the source (s_l = 32 >> l; the
window period 8192 is a whole number of tiles on each of the 6 levels).  The container is a valid MIC3 file of the
full size: 48-byte header, 6 level descriptors, 163 487 u64 table entries, > 4 GB of tile data.  Every range is checked
on a few tiles against the pixels the single-tile call returns for the source container, level 0 also against the
generator's pixels.  Prints one JSON object."""
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
sys.path.insert(0, os.path.join(ROOT, "tools"))
TILE = 256


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=100000)
    ap.add_argument("--height", type=int, default=80000)
    ap.add_argument("--levels", type=int, default=6)
    ap.add_argument("--chunk", type=int, default=16384)
    a = ap.parse_args()
    import __graft_entry__ as g
    import mic3_bench as m3

    g.build()
    mic = importlib.import_module("medical-image-codec_b200")
    api, lib = mic.api, mic.api.lib
    ndev = lib.micgpu_device_count()
    t0 = time.time()
    src = m3.make_source(11)
    SS = m3.SRC_SIDE
    src_blob = np.frombuffer(mic.CompressWSI(src.ravel(), SS, SS, 3, 8, TILE, TILE, a.levels), np.uint8)
    shdr = mic.ReadWSIHeader(src_blob)
    assert len(shdr["Levels"]) == a.levels
    stiles, _ = m3.parse_tiles(src_blob)
    t_src = time.time() - t0
    # ---- geometry of the full slide (computeLevels, wsiformat.go:245-285: floor halves, min 1) -------------------------
    levels, first = [], 0
    w, h = a.width, a.height
    for l in range(a.levels):
        tx, ty = (w + TILE - 1) // TILE, (h + TILE - 1) // TILE
        levels.append((w, h, tx, ty, first))
        first += tx * ty
        w, h = max(1, w // 2), max(1, h // 2)
    total = first
    order = np.empty(total, np.int64)
    for l, (lw, lh, tx, ty, f0) in enumerate(levels):
        _, _, stx, sty, sf = shdr["Levels"][l]
        yy, xx = np.divmod(np.arange(tx * ty), tx)
        order[f0:f0 + tx * ty] = sf + (yy % sty) * stx + (xx % stx)
    slen = np.array([t[1] for t in stiles], np.uint64)
    soff = np.array([t[0] for t in stiles], np.int64)
    lens = slen[order]
    offs = np.zeros(total, np.uint64)
    offs[1:] = np.cumsum(lens)[:-1]
    hdr = bytearray(48)
    hdr[0:4] = b"MIC3"
    hdr[4:8] = (1).to_bytes(4, "little")
    hdr[8:12] = a.width.to_bytes(4, "little")
    hdr[12:16] = a.height.to_bytes(4, "little")
    hdr[16:20] = TILE.to_bytes(4, "little")
    hdr[20:24] = TILE.to_bytes(4, "little")
    hdr[24:26] = (3).to_bytes(2, "little")
    hdr[26] = 8
    hdr[27] = src_blob[27]
    hdr[28:30] = a.levels.to_bytes(2, "little")
    hdr[32:40] = total.to_bytes(8, "little")
    lv = b"".join(int(v).to_bytes(4, "little") for L in levels for v in L)
    table = np.empty((total, 2), "<u8")
    table[:, 0] = offs
    table[:, 1] = lens
    data_off = 48 + len(lv) + table.nbytes
    size = data_off + int(lens.sum())
    pin = lib.micgpu_host_alloc(size + 256)
    if not pin:
        raise SystemExit("pinned allocation of the container failed: " + api.last_error())
    blob = np.ctypeslib.as_array(C.cast(pin, C.POINTER(C.c_uint8)), shape=(size + 256,))
    p = 0
    for chunk in (np.frombuffer(bytes(hdr), np.uint8), np.frombuffer(lv, np.uint8), table.view(np.uint8).ravel()):
        blob[p:p + chunk.size] = chunk
        p += chunk.size
    t0 = time.time()
    for i in range(total):
        o, n = int(soff[order[i]]), int(lens[i])
        blob[p:p + n] = src_blob[o:o + n]
        p += n
    assert p == size
    t_asm = time.time() - t0
    info = mic.ReadWSIHeader(blob[:size])
    assert info["TotalTiles"] == total and info["Levels"] == [tuple(L) for L in levels]
    tb = TILE * TILE * 3
    pout = lib.micgpu_host_alloc(a.chunk * tb)
    hout = np.ctypeslib.as_array(C.cast(pout, C.POINTER(C.c_uint8)), shape=(a.chunk * tb,))
    st = (C.c_int * a.chunk)()
    out = {"workload": f"MIC3 {a.width}x{a.height} RGB8, {a.levels} levels, {total} tiles of 256x256 (procedural slide, 8192^2 period)",
           "container_GB": round(size / 1e9, 3), "tiles_per_level": [L[2] * L[3] for L in levels], "source_s": round(t_src, 1), "assemble_s": round(t_asm, 1),
           "host_threads": len(os.sched_getaffinity(0))}
    # tiles to check per range: decoded through the single-tile call on the SOURCE container
    rng = np.random.default_rng(5)
    for devs in ([0], list(range(ndev))) if ndev > 1 else ([0],):
        mic.Init(devs)
        t_all, checked, ok, per_call = 0.0, 0, True, []
        # untimed first pass over every range: the contexts' grow-only scratch reaches its final size (a range of an upper
        # level has more coded planes per tile than one of level 0, and a reallocation is a cudaFree + cudaMalloc of GBs)
        for f in range(0, total, a.chunk):
            n = min(a.chunk, total - f)
            lib.micgpu_wsi_decompress_tile_range(pin, size, C.c_uint64(f), C.c_uint64(n), pout, n * tb, st)
        for f in range(0, total, a.chunk):
            n = min(a.chunk, total - f)
            t0 = time.perf_counter()
            rc = lib.micgpu_wsi_decompress_tile_range(pin, size, C.c_uint64(f), C.c_uint64(n), pout, n * tb, st)
            per_call.append(round((time.perf_counter() - t0) * 1e3, 1))
            t_all += per_call[-1] * 1e-3
            if rc:
                raise RuntimeError(f"micgpu_wsi_decompress_tile_range rc={rc}: {api.last_error()}")
            for t in rng.integers(0, n, 3):
                gi = f + int(t)
                l = max(k for k, L in enumerate(levels) if L[4] <= gi)
                _, _, stx, sty, sf = shdr["Levels"][l]
                si = int(order[gi]) - sf
                ref, rw, rh = mic.DecompressWSITile(src_blob, l, si % stx, si // stx)
                got = hout[int(t) * tb:(int(t) + 1) * tb].reshape(TILE, TILE, 3)
                ok &= bool(np.array_equal(got[:rh, :rw], ref.reshape(rh, rw, 3)))
                if l == 0:
                    sx, sy = (si % stx) * TILE, (si // stx) * TILE
                    ok &= bool(np.array_equal(got, src[sy:sy + TILE, sx:sx + TILE]))
                checked += 1
        out[f"{len(devs)}gpu"] = {"GBps_rgb": round(total * tb / t_all / 1e9, 2), "seconds": round(t_all, 3), "calls": (total + a.chunk - 1) // a.chunk,
                                  "tiles_checked": checked, "bit_exact": ok, "ms_per_call": per_call}
        lib.micgpu_shutdown()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
