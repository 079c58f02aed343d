"""GPU parity for MIC3 tiles / regions and MICR RGB payloads vs the CPU oracle (bit-exact bytes)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def slide(oracle, synth):
    W, H = 700, 533
    rgb = synth.wsi_region(11, 900, 700, W, H, 2500, 2000)     # straddles the tissue edge: white + H&E + nuclei
    rgb[400:, :300] = 255                                       # constant planes (modes 0 / 1)
    rgb[:40, 600:] = 0
    return rgb, oracle.wsi_compress(rgb.ravel(), W, H, 3, 8, 256, 256, 0)


def test_header(mic, oracle, slide):
    rgb, blob = slide
    h, o = mic.ReadWSIHeader(blob), oracle.wsi_header(blob)
    assert (h["Width"], h["Height"], h["TileWidth"], h["TileHeight"], h["Channels"], h["BitsPerSample"]) == (700, 533, 256, 256, 3, 8)
    assert h["ColorTransform"] and h["Levels"] == o["levels"] and h["TotalTiles"] == o["total_tiles"]


def test_every_tile_of_every_level(mic, oracle, slide):
    # TestWSICompressMedium / TestWSIPyramidLevels / TestWSICompressOddDimensions wsi_test.go:493-780
    rgb, blob = slide
    hdr = mic.ReadWSIHeader(blob)
    req = [(l, tx, ty) for l, (w, h, ntx, nty, _) in enumerate(hdr["Levels"]) for ty in range(nty) for tx in range(ntx)]
    got = mic.DecompressWSITiles(blob, req)
    for (l, tx, ty), (px, w, h) in zip(req, got):
        ref, rw, rh = oracle.wsi_decompress_tile(blob, l, tx, ty)
        assert (w, h) == (rw, rh) and np.array_equal(px, ref), (l, tx, ty)
    # level 0 equals the source
    px, w, h = mic.DecompressWSITile(blob, 0, 2, 2)
    assert (w, h) == (700 - 512, 533 - 512)
    assert np.array_equal(px.reshape(h, w, 3), rgb[512:, 512:])


def test_region_cross_tile(mic, oracle, slide):
    # TestWSIRegionCrossTile wsi_test.go:700+
    rgb, blob = slide
    for (l, x, y, w, h) in [(0, 200, 180, 150, 120), (0, 0, 0, 700, 533), (0, 650, 500, 200, 200), (1, 100, 90, 200, 150), (2, 0, 0, 175, 133)]:
        got, gw, gh = mic.DecompressWSIRegion(blob, l, x, y, w, h)
        ref, rw, rh = oracle.wsi_decompress_region(blob, l, x, y, w, h)
        assert (gw, gh) == (rw, rh) and np.array_equal(got, ref), (l, x, y, w, h)
    got, gw, gh = mic.DecompressWSIRegion(blob, 0, 200, 180, 150, 120)
    assert np.array_equal(got.reshape(gh, gw, 3), rgb[180:300, 200:350])
    with pytest.raises(mic.MicGpuError):
        mic.DecompressWSIRegion(blob, 0, 700, 0, 10, 10)         # empty region
    with pytest.raises(mic.MicGpuError):
        mic.DecompressWSITile(blob, 9, 0, 0)


def test_rgb_payload(mic, oracle, synth):
    # CompressRGB/DecompressRGB rgbcompress.go:25-33 (MICR payload); odd size, no tile padding
    rgb = synth.wsi_region(5, 1000, 900, 301, 203, 2500, 2000)
    blob = oracle.rgb_compress(rgb.ravel(), 301, 203, True)
    got = mic.DecompressRGB(blob, 301, 203)
    assert np.array_equal(got, oracle.rgb_decompress(blob, 301, 203, True))
    assert np.array_equal(got, rgb.ravel())


@pytest.mark.parametrize("bps", [8, 16])
def test_greyscale_slide(mic, oracle, synth, bps):
    # compressGreyTileBlob / decompressGreyTileBlob wsicompress.go:366-369,476-484
    W, H = 300, 270
    g = synth.xr_image(4, W, H)
    px = (g >> 4).astype(np.uint8).tobytes() if bps == 8 else (g >> 3).astype("<u2").tobytes()
    # (one pyramid level for 16-bit: the tiny top levels have more distinct symbols than FSE table cells and the
    # reference coder rejects them)
    blob = oracle.wsi_compress(np.frombuffer(px, np.uint8), W, H, 1, bps, 128, 128, 0 if bps == 8 else 1)
    hdr = mic.ReadWSIHeader(blob)
    for l, (w, h, ntx, nty, _) in enumerate(hdr["Levels"]):
        for ty in range(nty):
            for tx in range(ntx):
                got, gw, gh = mic.DecompressWSITile(blob, l, tx, ty)
                ref, rw, rh = oracle.wsi_decompress_tile(blob, l, tx, ty)
                assert (gw, gh) == (rw, rh) and np.array_equal(got, ref)
    full, fw, fh = mic.DecompressWSIRegion(blob, 0, 0, 0, W, H)
    assert np.array_equal(full.tobytes(), px)


def _full_tiles_from_oracle(oracle, blob, hdr):
    """Every tile of the table as a full (zero-padded) tile, from the oracle's cropped tiles."""
    tw, th = hdr["TileWidth"], hdr["TileHeight"]
    bpp = hdr["Channels"] * (2 if hdr["BitsPerSample"] == 16 else 1)
    out = []
    for l, (w, h, ntx, nty, first) in enumerate(hdr["Levels"]):
        for ty in range(nty):
            for tx in range(ntx):
                ref, rw, rh = oracle.wsi_decompress_tile(blob, l, tx, ty)
                full = np.zeros((th, tw * bpp), np.uint8)
                full[:rh, : rw * bpp] = np.asarray(ref, np.uint8).reshape(rh, rw * bpp)
                out.append(full.ravel())
    return out


def test_tile_range_batch(mic, oracle, slide):
    """The batch / serving path: tile-index ranges of the table decode to full tiles, in one call and through a plan
    that runs from device-resident bytes (micgpu_wsi_plan_*).  Edge tiles carry the encoder's zero padding."""
    import torch

    rgb, blob = slide
    hdr = mic.ReadWSIHeader(blob)
    ref = _full_tiles_from_oracle(oracle, blob, hdr)
    n = hdr["TotalTiles"]
    assert n == len(ref)
    got = mic.DecompressWSITileRange(blob, 0, n)
    for i in range(n):
        assert np.array_equal(got[i], ref[i]), i
    # sub-ranges, including one that starts inside level 0 and ends in level 1, and an empty one
    for first, cnt in [(2, 5), (7, n - 7), (n - 1, 1), (3, 0)]:
        part = mic.DecompressWSITileRange(blob, first, cnt)
        assert part.shape[0] == cnt
        for i in range(cnt):
            assert np.array_equal(part[i], ref[first + i]), (first, i)
    with pytest.raises(mic.MicGpuError):
        mic.DecompressWSITileRange(blob, n - 1, 2)
    # plan once, run twice from device memory
    a = np.frombuffer(blob, np.uint8)
    plan = mic.WsiPlan(blob, 1, n - 2)
    assert plan.tile_bytes == 256 * 256 * 3 and plan.out_bytes == (n - 2) * plan.tile_bytes and plan.n_units > 0
    d_span = torch.zeros(plan.span_len + 256, dtype=torch.uint8, device="cuda")
    d_span[: plan.span_len] = torch.from_numpy(a[plan.span_off: plan.span_off + plan.span_len].copy()).cuda()
    d_out = torch.zeros(plan.out_bytes, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        d_out.fill_(0x5A)
        plan.run_device(d_span.data_ptr(), d_out.data_ptr(), stream)
        assert not any(plan.status(stream))
        out = d_out.cpu().numpy().reshape(n - 2, plan.tile_bytes)
        for i in range(n - 2):
            assert np.array_equal(out[i], ref[1 + i]), i
    assert plan.last_launches >= 4
    plan.close()


def test_tile_range_reports_the_damaged_tile(mic, oracle, slide):
    rgb, blob = slide
    hdr = mic.ReadWSIHeader(blob)
    n = hdr["TotalTiles"]
    nlv = len(hdr["Levels"])
    table_off = 48 + 20 * nlv
    data_off = table_off + 16 * n
    b = bytearray(blob)
    o = int.from_bytes(blob[table_off + 16 * 4: table_off + 16 * 4 + 8], "little")
    b[data_off + o + 1] ^= 0xFF                 # plane length of tile 4 no longer fits its blob
    b[data_off + o + 3] |= 0x40
    a = np.frombuffer(bytes(b), np.uint8)
    tile_bytes = 256 * 256 * 3
    out = np.zeros(n * tile_bytes, np.uint8)
    import ctypes as C

    st = (C.c_int * n)()
    rc = mic.lib.micgpu_wsi_decompress_tile_range(a.ctypes.data, a.size, C.c_uint64(0), C.c_uint64(n), out.ctypes.data, out.size, st)
    assert rc != 0 and st[4] != 0
    good = mic.DecompressWSITileRange(blob, 0, n)
    for i in range(n):
        if i != 4:
            assert st[i] == 0 and np.array_equal(out[i * tile_bytes:(i + 1) * tile_bytes], good[i]), i


# ---- region / viewport serving from a resident slide ------------------------------------------------------
def test_slide_viewport_batches(mic, oracle, slide):
    """One handle, headers parsed once, the caller's buffer released; batches of rectangles across pyramid levels equal
    DecompressWSIRegion of the oracle (wsicompress.go:220-296) rectangle by rectangle, clamping included."""
    rgb, blob = slide
    buf = bytearray(blob)
    s = mic.WsiSlide(buf)
    for i in range(len(buf)):           # the handle keeps what it needs: scribble over the caller's copy
        buf[i] = 0xAA
    hdr = s.header
    rects = [(0, 0, 0, 700, 533),                    # the whole level
             (0, 250, 250, 20, 20),                  # corner shared by four tiles
             (0, 255, 100, 2, 300),                  # a two-pixel column across a tile seam
             (0, 600, 500, 400, 400),                # clamped on both axes
             (0, 699, 532, 1, 1)]                    # the last pixel
    for l, (w, h, _, _, _) in enumerate(hdr["Levels"][1:], start=1):
        rects += [(l, 0, 0, w, h), (l, w // 3, h // 4, max(1, w // 2), max(1, h // 2))]
    before = s.stats()
    got = s.regions(rects)
    after = s.stats()
    assert after["requests"] - before["requests"] == len(rects)
    # the five level-0 rectangles share the level's nine tiles: every tile is decoded once per batch
    ntiles_all = sum(tx * ty for (_, _, tx, ty, _) in hdr["Levels"])
    assert after["tiles_decoded"] - before["tiles_decoded"] <= ntiles_all
    for r, (px, w, h, st) in zip(rects, got):
        ref, rw, rh = oracle.wsi_decompress_region(blob, *r)
        assert st == 0 and (w, h) == (rw, rh) and np.array_equal(px, ref), r
    full = got[0][0].reshape(533, 700, 3)
    assert np.array_equal(full, rgb)
    # single-rectangle call, repeated: same bytes every time (scratch is recycled between requests)
    for _ in range(3):
        px, w, h = s.region(0, 100, 90, 333, 222)
        assert (w, h) == (333, 222) and np.array_equal(px.reshape(h, w, 3), rgb[90:312, 100:433])
    s.close()


def test_slide_bad_requests_do_not_poison_the_batch(mic, oracle, slide):
    rgb, blob = slide
    s = mic.WsiSlide(blob)
    rects = [(0, 10, 10, 50, 40), (7, 0, 0, 5, 5), (0, -1, 0, 5, 5), (0, 700, 0, 5, 5), (0, 0, 0, 0, 9), (1, 3, 2, 40, 30)]
    res = s.regions(rects, raise_on_error=False)
    assert [r[3] == 0 for r in res] == [True, False, False, False, False, True]
    for r, (px, w, h, st) in zip(rects, res):
        if st == 0:
            ref, rw, rh = oracle.wsi_decompress_region(blob, *r)
            assert (w, h) == (rw, rh) and np.array_equal(px, ref)
        else:
            assert (w, h) == (0, 0)
    with pytest.raises(mic.MicGpuError):
        s.regions(rects)
    s.close()
    # a damaged tile fails the rectangles that touch it and only those
    hdr = mic.ReadWSIHeader(blob)
    n_tiles = hdr["TotalTiles"]
    table = 48 + 20 * len(hdr["Levels"])
    data = table + 16 * n_tiles
    off4 = int.from_bytes(blob[table + 16 * 4:table + 16 * 4 + 8], "little")      # tile (1,1) of level 0
    dmg = bytearray(blob)
    for k in range(40, 60):
        dmg[data + off4 + k] ^= 0x5A
    s = mic.WsiSlide(bytes(dmg))
    res = s.regions([(0, 0, 0, 200, 200), (0, 300, 300, 100, 100), (0, 520, 10, 100, 100)], raise_on_error=False)
    assert res[0][3] == 0 and res[2][3] == 0
    assert np.array_equal(res[0][0].reshape(200, 200, 3), rgb[:200, :200])
    if res[1][3] == 0:      # the damage may decode to wrong pixels without tripping a check (Go behaves the same way)
        assert not np.array_equal(res[1][0].reshape(100, 100, 3), rgb[300:400, 300:400])
    s.close()
    with pytest.raises(mic.MicGpuError):
        mic.WsiSlide(b"MIC3" + bytes(blob[4:40]))
