"""GPU robustness: damaged containers and streams must come back as negative statuses -- never a hang, a fault or
stale pixels -- and must not disturb the healthy units that share their launch."""
import ctypes as C

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def _batch(mic, blobs, dims):
    n = len(blobs)
    views = [np.frombuffer(bytes(b), np.uint8) for b in blobs]
    outs = [np.full(w * h, 0xABCD, np.uint16) for (w, h) in dims]
    bp = (C.c_void_p * n)(*[v.ctypes.data for v in views]); ln = (C.c_size_t * n)(*[v.size for v in views])
    op = (C.c_void_p * n)(*[o.ctypes.data for o in outs]); cp = (C.c_size_t * n)(*[o.size for o in outs])
    st = (C.c_int * n)()
    rc = mic.lib.micgpu_pics_decompress_batch(n, bp, ln, op, cp, st)
    return rc, list(st), outs


def test_fuzz_corrupt_units_in_mixed_batches(mic, oracle, synth):
    """>= 200 seeded damages (bit flips, truncation, inflated symbol count, broken ncount header) on mixed 2/4/8-state
    batches: the call returns, the damaged image never passes silently when the damage is detectable by construction,
    and every neighbour decodes exactly."""
    shapes = [(97, 33), (300, 60), (64, 64), (211, 47), (500, 17), (128, 40), (257, 29), (40, 150), (333, 31), (256, 12)]
    imgs, blobs, dims = [], [], []
    for i in range(20):
        w, h = shapes[i % len(shapes)]
        im = synth.xr_image(900 + i, w, h).ravel()
        imgs.append(im); dims.append((w, h))
        # strips of a few rows are too short for the 4-/8-state coders at 12 bits: fewer strips for the flat shapes
        strips = (1, 2, 4, 3)[i % 4] if h >= 40 else 1
        blobs.append(oracle.pics_compress(im, w, h, int(im.max()), strips, (2, 4, 8)[i % 3]))
    rc, st, outs = _batch(mic, blobs, dims)
    assert rc == 0 and all(np.array_equal(o, im) for o, im in zip(outs, imgs))
    kinds = ("flip", "truncate", "count", "ncount")
    detected = {k: 0 for k in kinds}
    for seed in range(208):
        rng = np.random.default_rng(seed)
        victim = int(rng.integers(0, len(blobs)))
        kind = kinds[seed % len(kinds)]
        b = bytearray(blobs[victim])
        ns = int.from_bytes(b[12:16], "little")
        hdr = 20 + 8 * ns
        s = int(rng.integers(0, ns))
        off, ln = int.from_bytes(b[20 + 8 * s:24 + 8 * s], "little"), int.from_bytes(b[24 + 8 * s:28 + 8 * s], "little")
        f0 = hdr + off
        if kind == "flip":          # random bit flips anywhere in the strip (header, ncount, bitstream)
            for _ in range(int(rng.integers(1, 6))):
                b[f0 + int(rng.integers(0, ln))] ^= 1 << int(rng.integers(0, 8))
        elif kind == "truncate":    # the strip loses its tail; last byte kept non-zero so that the bit reader starts
            keep = int(rng.integers(8, max(9, ln - 4)))
            b[24 + 8 * s:28 + 8 * s] = keep.to_bytes(4, "little")
            b[f0 + keep - 1] |= 1
        elif kind == "count":       # the symbol count of the frame prefix is raised
            c = int.from_bytes(b[f0 + 2:f0 + 6], "little")
            b[f0 + 2:f0 + 6] = (c + int(rng.integers(64, 1 << 16))).to_bytes(4, "little")
        else:                       # bytes of the ncount header are overwritten
            for k in range(6, 6 + int(rng.integers(2, 12))):
                b[f0 + k] = int(rng.integers(0, 256))
        trial = list(blobs)
        trial[victim] = bytes(b)
        rc, st, outs = _batch(mic, trial, dims)
        for i in range(len(blobs)):
            if i == victim:
                continue
            assert st[i] == 0 and np.array_equal(outs[i], imgs[i]), f"seed {seed} ({kind} on image {victim}): neighbour {i} disturbed"
        if st[victim] != 0:
            detected[kind] += 1
            assert rc != 0
        elif kind in ("truncate", "count"):
            pytest.fail(f"seed {seed}: {kind} on image {victim} strip {s} was accepted")
    # flips and header noise are caught whenever they break the framing; most of them do
    assert detected["truncate"] == 52 and detected["count"] == 52
    assert detected["ncount"] >= 40, detected


def test_pics_short_strip_table_leaves_zero_rows(mic, oracle, synth):
    """A header whose strips stop short of the image height: the reference returns the rows nobody wrote as make() left
    them, zero (parallelstrips.go:288).  The device output is recycled scratch, so a previous decode must not show."""
    w, h = 160, 90
    img = synth.xr_image(31, w, h).ravel()
    full = oracle.pics_compress(img, w, h, int(img.max()), 3, 8)
    mic.DecompressParallelStrips(full)        # leaves non-zero pixels in the scratch output
    b = bytearray(full)
    b[8:12] = (h + 50).to_bytes(4, "little")  # 3 strips x 30 rows, image now 140 rows
    got, ow, oh = mic.DecompressParallelStrips(bytes(b))
    assert (ow, oh) == (w, h + 50)
    assert np.array_equal(got[: w * h], img) and not got[w * h:].any()
    ref, _, _ = oracle.pics_decompress(bytes(b))
    assert np.array_equal(got, ref)


def test_mic2_residual_of_wrong_length_is_rejected(mic, oracle, synth):
    """A temporal residual frame must expand to exactly width*height words (multiframecompress.go:165-175 + TemporalDeltaDecode);
    a shorter one used to succeed and leave stale scratch in the running sum."""
    w, h, nf = 64, 48, 4
    st = synth.tomo_stack(9, nf, h, w)     # (frames, rows, cols)
    blob = bytearray(oracle.mic2_compress(st.ravel(), w, h, 1023, True))
    res = oracle.temporal_encode(st[2].ravel(), st[1].ravel())
    short = oracle.compress_residual_frame(res[: w * h - 100], int(res.max()))
    # rebuild the container with frame 2 replaced
    n = int.from_bytes(blob[12:16], "little")
    data_off = 20 + 8 * n
    frames = []
    for i in range(n):
        o, l = int.from_bytes(blob[20 + 8 * i:24 + 8 * i], "little"), int.from_bytes(blob[24 + 8 * i:28 + 8 * i], "little")
        frames.append(bytes(blob[data_off + o:data_off + o + l]))
    frames[2] = short
    out = bytearray(blob[:20])
    off = 0
    for f in frames:
        out += off.to_bytes(4, "little") + len(f).to_bytes(4, "little")
        off += len(f)
    for f in frames:
        out += f
    with pytest.raises(mic.MicGpuError):
        mic.DecompressMultiFrame(bytes(out))
    frames_ok, _ = mic.DecompressMultiFrame(bytes(blob))
    assert np.array_equal(frames_ok.reshape(st.shape), st)


def _wsi_blob(oracle, synth):
    rgb = synth.wsi_region(11, 0, 0, 600, 400, 600, 400)
    return rgb, oracle.wsi_compress(rgb, 600, 400, 3, 8, 256, 256, 0)


def test_mic3_crafted_headers_are_rejected(mic, oracle, synth):
    """Every MIC3 field that feeds an index computation is file data: wrapped tile offsets, level
    descriptors that disagree with the tile grid, oversized level counts.  Each must fail cleanly; the context must
    stay usable (a GPU fault would be sticky)."""
    rgb, blob = _wsi_blob(oracle, synth)
    hdr = mic.ReadWSIHeader(blob)
    nlv = len(hdr["Levels"])
    table_off = 48 + 20 * nlv
    good = mic.DecompressWSITile(blob, 0, 1, 1)

    def expect_fail(b, what):
        with pytest.raises(mic.MicGpuError):
            mic.DecompressWSITile(bytes(b), 0, 1, 1)
        with pytest.raises(mic.MicGpuError):
            mic.DecompressWSIRegion(bytes(b), 0, 200, 100, 300, 250)

    tiles_x = (600 + 255) // 256
    idx = 1 * tiles_x + 1
    b = bytearray(blob)                       # (a) offset + length wraps around 2^64
    b[table_off + 16 * idx:table_off + 16 * idx + 8] = ((1 << 64) - 5000).to_bytes(8, "little")
    b[table_off + 16 * idx + 8:table_off + 16 * idx + 16] = (4500).to_bytes(8, "little")
    expect_fail(b, "wrapped offset")
    b = bytearray(blob)                       # (b) level 0 claims to be 10 pixels wide while keeping its 3x2 tile grid
    b[48:52] = (10).to_bytes(4, "little")
    expect_fail(b, "level width")
    b = bytearray(blob)                       # (c) tile grid larger than the level
    b[48 + 8:48 + 12] = (40).to_bytes(4, "little")
    expect_fail(b, "tile grid")
    b = bytearray(blob)                       # (d) first tile index beyond the table
    b[48 + 16:48 + 20] = (1 << 30).to_bytes(4, "little")
    expect_fail(b, "first tile")
    b = bytearray(blob)                       # (e) 65535 levels
    b[28:30] = (65535).to_bytes(2, "little")
    with pytest.raises(mic.MicGpuError):
        mic.ReadWSIHeader(bytes(b))
    b = bytearray(blob)                       # (f) plane lengths of a tile blob exceed the blob
    o = int.from_bytes(blob[table_off + 16 * idx:table_off + 16 * idx + 8], "little")
    data_off = table_off + 16 * hdr["TotalTiles"]
    b[data_off + o:data_off + o + 4] = (1 << 31).to_bytes(4, "little")
    with pytest.raises(mic.MicGpuError):
        mic.DecompressWSITile(bytes(b), 0, 1, 1)
    # the context survived all of it
    again = mic.DecompressWSITile(blob, 0, 1, 1)
    assert np.array_equal(again[0], good[0])
