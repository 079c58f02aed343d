"""GPU parity for WaveletV2 (5/3 integer lifting + RLE + 4-state FSE) decode vs the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _field(rows, cols, seed):
    i = np.arange(rows * cols)
    y, x = i // cols, i % cols
    return ((y * 5 + x * 3) % 3000 + (i * 97 + 13 + seed) % 7).astype(np.uint16)


@pytest.mark.parametrize("rows,cols,levels", [(64, 96, 3), (96, 64, 5), (128, 128, 5), (127, 129, 4), (256, 512, 8), (3, 2000, 5), (300, 200, 1),
                                              (255, 257, 2), (513, 130, 6)])
def test_wavelet_v2_decode(mic, oracle, rows, cols, levels):
    # dimension / level coverage of TestWaveletV2MultiLevelRoundTrip, TestWavelet2DSeparatedRoundTrip (odd sizes,
    # early level stop) through the whole pipeline
    src = _field(rows, cols, 0)
    blob = oracle.wavelet_v2_compress(src, rows, cols, 4095, levels)
    got, r, c = mic.WaveletV2RLEFSEDecompressU16(blob)
    ref, rr, cc = oracle.wavelet_v2_decompress(blob)
    assert (r, c) == (rr, cc) == (rows, cols)
    assert np.array_equal(got, ref) and np.array_equal(got, src)


@pytest.mark.parametrize("rows,cols,levels", [(127, 136, 4), (255, 264, 3), (65, 72, 6), (33, 4096, 4), (129, 520, 5), (64, 128, 1),
                                              (1000, 16, 3), (97, 2056, 2)])
def test_wavelet_v2_decode_tma_tiles(mic, oracle, rows, cols, levels):
    """Shapes the fused TMA level kernel takes (cols % 8 == 0): odd row counts, tiles cut by the right / bottom border,
    levels whose low band is odd x odd, one-tile and many-tile levels, boxes whose halo leaves the plane."""
    rng = np.random.default_rng(rows * 7 + cols)
    src = (_field(rows, cols, 3).astype(np.int64) + rng.integers(0, 500, rows * cols)).astype(np.uint16)
    blob = oracle.wavelet_v2_compress(src, rows, cols, 4095, levels)
    got, r, c = mic.WaveletV2RLEFSEDecompressU16(blob)
    assert (r, c) == (rows, cols) and np.array_equal(got, src)
    # two images of one geometry share the launch sequence (grid.z): the second must not see the first's planes
    src2 = np.roll(src, 17)
    blob2 = oracle.wavelet_v2_compress(src2, rows, cols, 4095, levels)
    res = mic.WaveletV2DecompressBatch([blob, blob2])
    assert np.array_equal(res[0][0], src) and np.array_equal(res[1][0], src2)


def test_wavelet_mammo_like(mic, oracle, synth):
    # BASELINE config 3 content at reduced size: 14-bit, ~45 % zero background, 5 levels
    img = synth.mammo_image(1, 1024, 832).ravel()
    blob = oracle.wavelet_v2_compress(img, 1024, 832, int(img.max()), 5)
    got, r, c = mic.WaveletV2SIMDRLEFSEDecompressU16(blob)
    assert (r, c) == (1024, 832) and np.array_equal(got, img)


def test_wavelet_escape_triples(mic, oracle):
    # coefficients beyond +-32767 travel as 65535,hi,lo (waveletfsecompressu16.go:31-37)
    rng = np.random.default_rng(3)
    src = (rng.integers(0, 2, 96 * 96) * 65535).astype(np.uint16)
    blob = oracle.wavelet_v2_compress(src, 96, 96, 65535, 3)
    got, _, _ = mic.WaveletV2RLEFSEDecompressU16(blob)
    assert np.array_equal(got, src)


def test_wavelet_batch_mixed_geometry(mic, oracle):
    cases = [(128, 128, 5, 1), (128, 128, 5, 2), (96, 64, 3, 3), (128, 128, 5, 4), (127, 129, 4, 5)]
    blobs = [oracle.wavelet_v2_compress(_field(r, c, s), r, c, 4095, l) for r, c, l, s in cases]
    res = mic.WaveletV2DecompressBatch(blobs)
    for (px, r, c), (rr, cc, _, s) in zip(res, cases):
        assert (r, c) == (rr, cc) and np.array_equal(px, _field(rr, cc, s))


def test_wavelet_bad_magic(mic, oracle):
    blob = bytearray(oracle.wavelet_v2_compress(_field(64, 96, 0), 64, 96, 4095, 3))
    blob[12] = 0x02
    with pytest.raises(mic.MicGpuError):
        mic.WaveletV2RLEFSEDecompressU16(bytes(blob))


# ---- SURVEY 8(f).4 (wavelet half): the V1 layouts ---------------------------------------------------------------------------
@pytest.mark.parametrize("with_rle", [0, 1])
@pytest.mark.parametrize("rows,cols,levels", [(64, 40, 1), (65, 33, 2), (128, 96, 3), (257, 129, 4), (300, 517, 4), (40, 1000, 9), (9, 4000, 3), (4000, 9, 4)])
def test_wavelet_v1_layouts_decode(mic, oracle, synth, with_rle, rows, cols, levels):
    """WaveletFSEDecompressU16 / WaveletRLEFSEDecompressU16 (waveletfsecompressu16.go:124-163, 624-669): raster-order
    coefficients, interleaved in-place lifting on the top-left corner per level (waveletInverse2DRegion :180-189), levels
    clamped to 4 by the encoder, early stop below 2 rows / columns."""
    px, mx = _field(rows, cols, 5), 4095
    blob = oracle.wavelet_v1_compress(px, rows, cols, mx, levels, with_rle)
    ref, r, c = oracle.wavelet_v1_decompress(blob, with_rle)
    assert (r, c) == (rows, cols) and np.array_equal(ref, px)
    got, gr, gc = (mic.WaveletRLEFSEDecompressU16 if with_rle else mic.WaveletFSEDecompressU16)(blob)
    assert (gr, gc) == (rows, cols) and np.array_equal(got, px)


def test_wavelet_v1_reference_images_and_escapes(mic, oracle):
    import os
    from conftest import GOLDEN

    for name, w, h in (("MR_256_256", 256, 256), ("CT_512_512", 512, 512)):   # CT: full 16-bit range -> escape triples
        px = np.fromfile(os.path.join(GOLDEN, f"{name}_image.bin"), dtype="<u2")
        for with_rle in (0, 1):
            blob = oracle.wavelet_v1_compress(px, h, w, int(px.max()), 3, with_rle)
            got, r, c = (mic.WaveletRLEFSEDecompressU16 if with_rle else mic.WaveletFSEDecompressU16)(blob)
            assert (r, c) == (h, w) and np.array_equal(got, px), (name, with_rle)
    # damaged streams: no hang, an error or the oracle's verdict
    blob = bytearray(oracle.wavelet_v1_compress(px, 512, 512, int(px.max()), 2, 1))
    for bad in (bytes(blob[:14]), bytes(blob[:15]) + b"\x00" * 8, bytes(blob[:40])):
        with pytest.raises(mic.MicGpuError):
            mic.WaveletRLEFSEDecompressU16(bad)


@pytest.mark.parametrize("with_rle", [0, 1])
@pytest.mark.parametrize("rows,cols,levels", [(64, 40, 1), (65, 33, 2), (128, 96, 3), (257, 129, 4), (300, 517, 4), (40, 1000, 9), (9, 4000, 3), (4000, 9, 4),
                                               (2, 2, 3), (3, 1, 2)])
def test_wavelet_v1_layouts_encode_bytes(mic, oracle, with_rle, rows, cols, levels):
    """WaveletFSECompressU16 / WaveletRLEFSECompressU16 (waveletfsecompressu16.go:71-123, 551-623): the device encoder
    (interleaved forward lifting, raster-order pack, FSE-4 directly or behind RLE) emits the oracle's bytes, and they decode."""
    px, mx = _field(rows, cols, 7), 4095
    enc = mic.WaveletRLEFSECompressU16 if with_rle else mic.WaveletFSECompressU16
    dec = mic.WaveletRLEFSEDecompressU16 if with_rle else mic.WaveletFSEDecompressU16
    try:
        want = oracle.wavelet_v1_compress(px, rows, cols, mx, levels, with_rle)
    except Exception:
        with pytest.raises(mic.MicGpuError):          # e.g. a 3 x 1 image: the FSE stage rejects the tiny stream in both
            enc(px, rows, cols, mx, levels)
        return
    got = enc(px, rows, cols, mx, levels)
    assert got == want
    out, r, c = dec(got)
    assert (r, c) == (rows, cols) and np.array_equal(out, px)


def test_wavelet_v1_encode_reference_images_and_escapes(mic, oracle):
    import os
    from conftest import GOLDEN

    for name, w, h in (("MR_256_256", 256, 256), ("CT_512_512", 512, 512)):   # CT: full 16-bit range -> escape triples
        px = np.fromfile(os.path.join(GOLDEN, f"{name}_image.bin"), dtype="<u2")
        for with_rle in (0, 1):
            enc = mic.WaveletRLEFSECompressU16 if with_rle else mic.WaveletFSECompressU16
            assert enc(px, h, w, int(px.max()), 3) == oracle.wavelet_v1_compress(px, h, w, int(px.max()), 3, with_rle), (name, with_rle)
