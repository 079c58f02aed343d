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
