"""Size-independent properties at the BASELINE.json shapes (the oracle is too slow to run whole batches, so full-size
cases use round trips, a few oracle spot checks and the reference decoder)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_config1_ct_single_state(mic, oracle):
    # configs[0]: Delta+RLE+FSE round trip of CT_512_512_image.bin with the 1-state coder (DeltaRleFSETest fseu16_test.go:679)
    import os
    from conftest import GOLDEN

    img = np.fromfile(os.path.join(GOLDEN, "CT_512_512_image.bin"), np.uint16)
    blob = mic.CompressSingleFrame(img, 512, 512, int(img.max()), 1)
    assert blob == oracle.compress_single_frame(img, 512, 512, int(img.max()), 1)
    assert np.array_equal(mic.DecompressSingleFrame(blob, 512, 512), img)
    assert round(img.nbytes / len(blob), 2) == 2.24           # README.md:270


def test_config2_batch_roundtrip(mic, oracle, reftwin, synth):
    # configs[1]: 2577x2048 PICS-8, a slice of the batch: encode -> decode round trip, first blob also vs the oracle,
    # 2-state container also through the reference's pthread decoder
    w, h = 2577, 2048
    imgs = [synth.xr_image(300 + i, w, h).ravel() for i in range(6)]
    mx = [int(i.max()) for i in imgs]
    blobs = mic.CompressParallelStripsBatch(imgs, w, h, mx, 8, 8)
    assert blobs[0] == oracle.pics_compress(imgs[0], w, h, mx[0], 8, 8)
    for (px, ow, oh), im in zip(mic.DecompressParallelStripsBatch(blobs), imgs):
        assert (ow, oh) == (w, h) and np.array_equal(px, im)
    two = mic.CompressParallelStrips(imgs[1], w, h, mx[1], 8, 2)
    assert np.array_equal(reftwin.decompress_parallel(two, w, h, 8), imgs[1])


def test_config3_mammo_wavelet(mic, oracle, synth):
    # configs[2]: 4096x3328 14-bit, WaveletV2 5 levels (stored levels = 5: 4096x3328 -> ... -> 256x208, LL 128x104)
    rows, cols = 4096, 3328
    imgs = [synth.mammo_image(1 + i, rows, cols).ravel() for i in range(2)]
    blobs = mic.WaveletV2CompressBatch(imgs, rows, cols, [int(i.max()) for i in imgs], 5)
    assert blobs[0][10] == 5
    assert blobs[0] == oracle.wavelet_v2_compress(imgs[0], rows, cols, int(imgs[0].max()), 5)
    for (px, r, c), im in zip(mic.WaveletV2DecompressBatch(blobs), imgs):
        assert (r, c) == (rows, cols) and np.array_equal(px, im)


@pytest.mark.parametrize("temporal", [False, True])
def test_config4_tomo_mic2(mic, oracle, synth, temporal):
    # configs[3]: 2457x1996 10-bit frames (a 6-frame slice of the 96-frame stack)
    st = synth.tomo_stack(7, 6, 2457, 1996)
    blob = mic.CompressMultiFrame(st, 1996, 2457, 1023, temporal)
    frames, hdr = mic.DecompressMultiFrame(blob)
    assert hdr == {"Width": 1996, "Height": 2457, "FrameCount": 6, "Temporal": temporal}
    assert np.array_equal(frames, st)
    assert np.array_equal(mic.DecompressFrame(blob, 3), st[3])
    # the first frame (spatial) and one later frame vs the oracle
    ref = oracle.mic2_compress(st[:2].ravel(), 1996, 2457, 1023, temporal)
    mine = mic.CompressMultiFrame(st[:2], 1996, 2457, 1023, temporal)
    assert mine == ref


def test_config5_wsi_window(mic, oracle, synth):
    # configs[4]: a 2300x1700 window of the 100000x80000 procedural slide that straddles the tissue edge; 256x256 tiles,
    # auto pyramid (5 levels), YCoCg-R
    W, H = 2300, 1700
    rgb = synth.wsi_region(11, 18000, 24000, W, H, 100000, 80000)
    blob = mic.CompressWSI(rgb, W, H)
    hdr = mic.ReadWSIHeader(blob)
    assert len(hdr["Levels"]) == 5 and hdr["Levels"][0][:4] == (W, H, 9, 7)
    full, fw, fh = mic.DecompressWSIRegion(blob, 0, 0, 0, W, H)
    assert (fw, fh) == (W, H) and np.array_equal(full.reshape(H, W, 3), rgb)
    # level 2 equals two box-filter steps (Downsample2xRGB wsipyramid.go:10-33)
    d1, w1, h1 = oracle.downsample2x_rgb(rgb.ravel(), W, H)
    d2, w2, h2 = oracle.downsample2x_rgb(d1, w1, h1)
    l2, lw, lh = mic.DecompressWSIRegion(blob, 2, 0, 0, w2, h2)
    assert (lw, lh) == (w2, h2) and np.array_equal(l2, d2)
    # one interior tile blob byte-for-byte vs the oracle's tile coder (CompressRGB on the padded tile)
    tile = np.ascontiguousarray(rgb[256:512, 512:768])
    assert mic.CompressRGB(tile, 256, 256) == oracle.rgb_compress(tile.ravel(), 256, 256, True)
