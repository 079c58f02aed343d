"""GPU parity for the encode direction: every stage and the finished streams must equal the CPU oracle byte for byte
(and therefore the reference C encoder for 2/4/8-state frames, see tests/golden/streams.json)."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _img(seed, w, h, noise=30):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    return (1500 + 3 * x + 5 * y + rng.integers(0, noise, (h, w))).astype(np.uint16).ravel()


# ---- stage 1: RLE state machine in closed form ---------------------------------------------------
@pytest.mark.parametrize("maxv", [255, 4095])
def test_rle_compress_run_structures(mic, oracle, maxv):
    mid = (1 << (int(maxv).bit_length() - 1)) - 1
    P = mid - 3
    rng = np.random.default_rng(maxv)
    cases = []
    # constant runs at every residue around the forced-flush boundaries (rlecompressu16.go:58)
    for L in [1, 2, 3, 4, mid - 2, mid - 1, mid, mid + 1, mid + P - 1, mid + P, mid + P + 1, mid + 2 * P, 3 * mid + 7]:
        cases.append(np.full(L, 9, np.uint16))
    # literal stretches around the same boundaries, followed by a run (m+2 rule) or by the end of the stream
    for m in [1, 2, mid - 3, mid - 2, mid - 1, mid, mid + 1, mid + P - 2, mid + P - 1, mid + P, 2 * mid + 5]:
        lit = (np.arange(m) % 7 + (np.arange(m) // 7) % 3 * 10).astype(np.uint16)   # no 3 equal in a row
        lit[1::2] += 20
        cases.append(lit)
        cases.append(np.concatenate([lit, np.full(5, 3, np.uint16)]))
        cases.append(np.concatenate([np.full(4, 3, np.uint16), lit, np.full(3, 5, np.uint16), lit[: max(1, m // 2)]]))
    # pairs stay literal, adjacent runs of different values, run at the very start / end
    cases.append(np.array([1, 1, 2, 2, 3, 3, 3, 4, 4, 4, 4, 5, 6, 6, 7], np.uint16))
    cases.append(np.array([5, 5, 5, 6, 6, 6, 7, 7, 7], np.uint16))
    cases.append(rng.integers(0, 3, 5000).astype(np.uint16))            # many short runs
    cases.append(rng.integers(0, 2, 20000).astype(np.uint16))
    cases.append(np.repeat(rng.integers(0, 200, 300), rng.integers(1, 400, 300)).astype(np.uint16))
    for k, v in enumerate(cases):
        if v.max() > maxv:
            v = (v % (maxv + 1)).astype(np.uint16)
        got = mic.RleCompress(v, maxv)
        ref = oracle.rle_compress(v, maxv)
        assert np.array_equal(got, ref), (maxv, k, v.size)


@pytest.mark.parametrize("w,h", [(64, 64), (257, 129), (611, 403), (1000, 37), (33, 31)])
def test_delta_rle_compress(mic, oracle, synth, w, h):
    img = synth.xr_image(w * 7 + h, w, h).ravel()          # zero borders -> long same-runs, ramp + noise -> literals
    mx = int(img.max())
    assert np.array_equal(mic.DeltaRleCompress(img, w, h, mx), oracle.delta_rle_compress(img, w, h, mx))
    # caller-supplied maxValue larger than the data (Go semantics)
    assert np.array_equal(mic.DeltaRleCompress(img, w, h, 65535), oracle.delta_rle_compress(img, w, h, 65535))


def test_delta_rle_escapes(mic, oracle):
    w, h = 500, 120
    img = np.zeros((h, w), np.uint16)
    img[:, 100:200] = 4095
    img[10:50, 250:400] = 2000
    img[60:, ::2] = 4095
    img[100:, :] = 777
    img = img.ravel()
    assert np.array_equal(mic.DeltaRleCompress(img, w, h, 4095), oracle.delta_rle_compress(img, w, h, 4095))


# ---- finished frames -------------------------------------------------------------------------------
@pytest.mark.parametrize("nstates", [1, 2, 4, 8])
@pytest.mark.parametrize("w,h", [(256, 256), (611, 403), (1000, 37), (64, 64)])
def test_single_frame_bytes(mic, oracle, synth, nstates, w, h):
    img = synth.xr_image(w + h, w, h).ravel()
    mx = int(img.max())
    got = mic.CompressSingleFrame(img, w, h, mx, nstates)
    ref = oracle.compress_single_frame(img, w, h, mx, nstates)
    assert got == ref
    assert np.array_equal(mic.DecompressSingleFrame(got, w, h), img)


def test_golden_reference_encoder_bytes(mic):
    """CT 512x512 and MR 256x256: the CUDA encoder reproduces the REFERENCE C encoder's bytes (hashes in
    tests/golden/streams.json were taken from oracle/_ref, the compiled reference), incl. tableLog 16 for CT."""
    gold = json.load(open(os.path.join(GOLDEN, "streams.json")))
    for name, w, h in (("CT_512_512", 512, 512), ("MR_256_256", 256, 256)):
        img = np.fromfile(os.path.join(GOLDEN, f"{name}_image.bin"), np.uint16)
        for ns in (1, 2, 4, 8):
            blob = mic.CompressSingleFrame(img, w, h, int(img.max()), ns)
            g = gold[name][f"frame_{ns}state"]
            assert (len(blob), hashlib.sha256(blob).hexdigest()) == (g["len"], g["sha256"]), (name, ns)


def test_16bit_tablelog16(mic, oracle):
    rng = np.random.default_rng(7)
    w, h = 530, 520
    img = (np.cumsum(rng.integers(-300, 301, w * h)) % 65536).astype(np.uint16)
    img[::97] = 65535
    for ns in (2, 8):
        assert mic.CompressSingleFrame(img, w, h, 65535, ns) == oracle.compress_single_frame(img, w, h, 65535, ns)


def test_rejects_match_reference(mic, oracle):
    from oracle.oracle import OracleError

    # constant image: the RLE stream still has several distinct symbols, so it codes fine; a 1x1 image is rejected
    img = np.full(64 * 64, 300, np.uint16)
    assert mic.CompressSingleFrame(img, 64, 64, 300, 2) == oracle.compress_single_frame(img, 64, 64, 300, 2)
    # tiny / degenerate inputs: same outcome as the reference semantics (same bytes, or both reject)
    # (maxValue below 16 is documented as unsupported: midCount <= 3 degenerates the reference RLE)
    for px, w, h, mx in [([500], 1, 1, 500), ([5, 9], 2, 1, 255), ([1, 2, 3, 4, 5, 6, 7, 8, 9], 3, 3, 300), (list(range(40)), 40, 1, 1000),
                         ([7] * 12, 4, 3, 255), ([0] * 50, 10, 5, 255)]:
        a = np.array(px, np.uint16)
        try:
            ref = oracle.compress_single_frame(a, w, h, mx, 8)
        except OracleError:
            ref = None
        try:
            got = mic.CompressSingleFrame(a, w, h, mx, 8)
        except mic.MicGpuError:
            got = None
        assert got == ref, (px, got, ref)


@pytest.mark.parametrize("nstates", [2, 4, 8])
@pytest.mark.parametrize("strips", [1, 3, 8])
def test_pics_bytes(mic, oracle, synth, nstates, strips):
    w, h = 611, 403
    img = synth.xr_image(11, w, h).ravel()
    got = mic.CompressParallelStrips(img, w, h, int(img.max()), strips, nstates)
    assert got == oracle.pics_compress(img, w, h, int(img.max()), strips, nstates)
    px, ow, oh = mic.DecompressParallelStrips(got)
    assert (ow, oh) == (w, h) and np.array_equal(px, img)


def test_pics_batch_full_geometry(mic, oracle, synth):
    # BASELINE config 2 geometry, two images: bytes equal the oracle's, and the reference decodes them
    w, h = 2577, 2048
    imgs = [synth.xr_image(s, w, h).ravel() for s in (1, 2)]
    blobs = mic.CompressParallelStripsBatch(imgs, w, h, [int(i.max()) for i in imgs], 8, 8)
    for im, b in zip(imgs, blobs):
        assert b == oracle.pics_compress(im, w, h, int(im.max()), 8, 8)


def test_twin_compress_symbols(mic, oracle, reftwin, synth):
    import ctypes as C

    w, h = 320, 240
    img = synth.xr_image(4, w, h).ravel()
    for ns, name in ((2, "two"), (4, "four"), (8, "eight")):
        out = np.zeros(2 * w * h + 4096, np.uint8)
        n = C.c_size_t()
        rc = getattr(mic.lib, f"mic_compress_{name}_state")(img.ctypes.data, w, h, out.ctypes.data, out.size, C.byref(n))
        assert rc == 0
        assert out[: n.value].tobytes() == reftwin.compress(img, w, h, ns)     # byte-identical to the reference binary


# ---- containers and front ends ---------------------------------------------------------------------
@pytest.mark.parametrize("temporal", [False, True])
def test_mic2_bytes(mic, oracle, synth, temporal):
    st = synth.tomo_stack(7, 5, 128, 128)
    got = mic.CompressMultiFrame(st, 128, 128, 1023, temporal)
    assert got == oracle.mic2_compress(st.ravel(), 128, 128, 1023, temporal)
    frames, hdr = mic.DecompressMultiFrame(got)
    assert hdr["Temporal"] == temporal and np.array_equal(frames, st)


def test_mic2_larger_temporal(mic, oracle, synth):
    st = synth.tomo_stack(3, 4, 301, 257)
    got = mic.CompressMultiFrame(st, 257, 301, 1023, True)
    assert got == oracle.mic2_compress(st.ravel(), 257, 301, 1023, True)


def _field(rows, cols, seed):
    i = np.arange(rows * cols)
    y, x = i // cols, i % cols
    return ((y * 5 + x * 3) % 3000 + (i * 97 + 13 + seed) % 7).astype(np.uint16)


@pytest.mark.parametrize("rows,cols,levels", [(64, 96, 3), (96, 64, 5), (128, 128, 5), (127, 129, 4), (256, 512, 8), (3, 2000, 5), (300, 200, 1), (513, 130, 6)])
def test_wavelet_bytes(mic, oracle, rows, cols, levels):
    src = _field(rows, cols, 0)
    got = mic.WaveletV2RLEFSECompressU16(src, rows, cols, 4095, levels)
    assert got == oracle.wavelet_v2_compress(src, rows, cols, 4095, levels)
    px, r, c = mic.WaveletV2RLEFSEDecompressU16(got)
    assert (r, c) == (rows, cols) and np.array_equal(px, src)


def test_wavelet_bytes_mammo_and_escapes(mic, oracle, synth):
    img = synth.mammo_image(1, 1024, 832).ravel()
    assert mic.WaveletV2SIMDRLEFSECompressU16(img, 1024, 832, int(img.max()), 5) == oracle.wavelet_v2_compress(img, 1024, 832, int(img.max()), 5)
    rng = np.random.default_rng(3)
    src = (rng.integers(0, 2, 96 * 96) * 65535).astype(np.uint16)        # coefficients beyond +-32767 -> escape triples
    assert mic.WaveletV2RLEFSECompressU16(src, 96, 96, 65535, 3) == oracle.wavelet_v2_compress(src, 96, 96, 65535, 3)
    blobs = mic.WaveletV2CompressBatch([_field(128, 128, s) for s in range(3)], 128, 128, [4095] * 3, 5)
    for s, b in enumerate(blobs):
        assert b == oracle.wavelet_v2_compress(_field(128, 128, s), 128, 128, 4095, 5)


def test_rgb_bytes(mic, oracle, synth):
    rgb = synth.wsi_region(5, 1000, 900, 301, 203, 2500, 2000)
    got = mic.CompressRGB(rgb, 301, 203)
    assert got == oracle.rgb_compress(rgb.ravel(), 301, 203, True)
    assert np.array_equal(mic.DecompressRGB(got, 301, 203), rgb.ravel())


def test_wsi_bytes(mic, oracle, synth):
    # tissue edge + white background + constant regions: plane modes 0/1/2, edge-tile zero padding, 3 pyramid levels
    W, H = 700, 533
    rgb = synth.wsi_region(11, 900, 700, W, H, 2500, 2000)
    rgb[400:, :300] = 255
    rgb[:40, 600:] = 0
    got = mic.CompressWSI(rgb, W, H)
    ref = oracle.wsi_compress(rgb.ravel(), W, H, 3, 8, 256, 256, 0)
    assert got == ref
    got2 = mic.CompressWSI(rgb, W, H, tile_width=128, tile_height=64, pyramid_levels=2)
    assert got2 == oracle.wsi_compress(rgb.ravel(), W, H, 3, 8, 128, 64, 2)


@pytest.mark.parametrize("bps", [8, 16])
def test_wsi_grey_bytes(mic, oracle, synth, bps):
    W, H = 300, 270
    g = synth.xr_image(4, W, H)
    px = (g >> 4).astype(np.uint8).tobytes() if bps == 8 else (g >> 3).astype("<u2").tobytes()
    lv = 0 if bps == 8 else 1
    got = mic.CompressWSI(np.frombuffer(px, np.uint8), W, H, channels=1, bits_per_sample=bps, tile_width=128, tile_height=128, pyramid_levels=lv)
    assert got == oracle.wsi_compress(np.frombuffer(px, np.uint8), W, H, 1, bps, 128, 128, lv)


def test_uncodable_stream_fails_like_the_reference(mic, oracle):
    """A temporal residual with more distinct symbols than FSE table cells cannot be coded by the reference either
    (normalizeCount2 underflows, fsecompressu16.go:582-667; the oracle reports it as an internal error).  The GPU
    encoder must report an error on every tier of the 8->4->2->1 ladder instead of retrying on tables it never built
    (that retry used to read outside the state table)."""
    rng = np.random.default_rng(5)
    w, h = 24, 25
    n = w * h
    k = int(rng.integers(150, 400))
    vals = rng.integers(0, 1000, k)
    d = vals[rng.integers(0, k, n)]
    f0 = rng.integers(2000, 3000, n).astype(np.uint16)
    f1 = (f0.astype(np.int64) + d - 500).astype(np.uint16)
    st = np.stack([f0, f1])
    with pytest.raises(Exception):
        oracle.mic2_compress(st.ravel(), w, h, 4095, True)
    for _ in range(3):      # repeated: the encoder context must stay usable after the failure
        with pytest.raises(mic.MicGpuError):
            mic.CompressMultiFrame(st, w, h, 4095, True)
    y, x = np.mgrid[0:64, 0:64]
    good = np.stack([(1000 + 3 * x + 5 * y + rng.integers(0, 8, (64, 64))).astype(np.uint16) for _ in range(2)])
    assert mic.CompressMultiFrame(good, 64, 64, 4095, True) == oracle.mic2_compress(good.ravel(), 64, 64, 4095, True)


def test_pics_batch_encode_threaded_contexts(mic, oracle, synth):
    """A batch of >= 16 images and >= 64 MB goes through several encoder contexts on worker threads; every image must
    still be byte-identical to the reference encoder, in order."""
    w, h = 2048, 1024
    base = [synth.xr_image(70 + i, w, h).ravel() for i in range(2)]
    imgs = [base[i % 2] for i in range(17)]
    blobs = mic.CompressParallelStripsBatch(imgs, w, h, [int(i.max()) for i in imgs], 8, 8)
    want = [oracle.pics_compress(b, w, h, int(b.max()), 8, 8) for b in base]
    assert len(blobs) == 17
    for i, b in enumerate(blobs):
        assert b == want[i % 2], f"image {i}"


@pytest.mark.parametrize("shape", [(256, 256), (300, 201), (97, 33), (130, 57), (211, 160), (128, 40)])
def test_rans8_frame_bytes(mic, oracle, synth, shape):
    """RANSCompressU16EightState (rans8state.go:31-220, ransu16.go:139-213) over the Delta+RLE symbols: the CUDA encoder's
    frame equals the oracle's byte for byte (every tail residue of len mod 8 through the shapes) and decodes back."""
    w, h = shape
    img = synth.xr_image(40 + w, w, h).ravel()
    mx = int(img.max())
    sym = oracle.delta_rle_compress(img, w, h, mx)
    want = oracle.fse_compress(sym, 108)
    got = mic.CompressSingleFrame(img, w, h, mx, 108)
    assert got[:2] == b"\xff\x08"
    assert got == want
    assert np.array_equal(mic.DecompressSingleFrame(got, w, h), img)


def test_rans8_rejects_like_the_reference(mic, oracle):
    # constant image: the symbol stream is one RLE run -> ErrIncompressible / ErrUseRLE, and no ladder behind rANS
    img = np.full(64 * 64, 700, np.uint16)
    sym = oracle.delta_rle_compress(img, 64, 64, 700)
    with pytest.raises(Exception):
        oracle.fse_compress(sym, 108)
    with pytest.raises(mic.MicGpuError):
        mic.CompressSingleFrame(img, 64, 64, 700, 108)
