"""the gradient-adaptive predictor (deltagradrlecompressu16.go, CompressSingleFrameGrad /
DecompressSingleFrameGrad, multiframecompress.go:111-142) and the PICA container (parallelstripsadaptive.go:54-289)
through the C ABI: decoded pixels and compressed bytes against the CPU oracle (a restatement of the Go source whose
ratios and per-strip predictor choices match docs/adaptive-compression.md; byte parity versus the Go encoder itself is
unpinned, as for every format the reference's C twin does not cover)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _flags(blob):
    n = int.from_bytes(blob[12:16], "little")
    return [int.from_bytes(blob[28 + 16 * i:32 + 16 * i], "little") for i in range(n)]


def _golden(name, w, h):
    return np.fromfile(os.path.join(GOLDEN, f"{name}_image.bin"), dtype="<u2"), w, h


@pytest.mark.parametrize("w,h", [(1, 1), (1, 37), (37, 1), (2, 2), (3, 3), (8, 5), (33, 70), (64, 64), (65, 33), (517, 263), (2577, 96), (300, 1031)])
def test_grad_frame_decode(mic, oracle, synth, w, h):
    # every geometry class of the wavefront: one row / one column (no gradient context at all), fewer rows than a band,
    # several bands per warp (h > 256), widths around the 32-step block size
    if w * h < 64:
        # tiny frames: 5-bit content, the short ncount header is what lets the FSE tiers accept them (test_ragged_sizes)
        i = np.arange(w * h)
        px, mx = (20 + (i % 3) + (i // max(w, 1)) % 2).astype(np.uint16), 31
    else:
        img = synth.smooth_image(w + h, w, h)
        px, mx = img.ravel(), int(img.max())
    frame = oracle.compress_single_frame_grad(px, w, h, mx)
    want = oracle.decompress_single_frame_grad(frame, w, h)
    assert np.array_equal(want, px)
    assert np.array_equal(mic.DecompressSingleFrameGrad(frame, w, h), px)


def test_grad_frame_reference_images(mic, oracle):
    for name, w, h in (("MR_256_256", 256, 256), ("CT_512_512", 512, 512)):
        px, _, _ = _golden(name, w, h)
        mx = int(px.max())
        frame = oracle.compress_single_frame_grad(px, w, h, mx)
        assert mic.CompressSingleFrameGrad(px, w, h, mx) == frame          # bytes
        assert np.array_equal(mic.DecompressSingleFrameGrad(frame, w, h), px)   # pixels (CT: tableLog 16, escapes)


def test_grad_escapes_and_wrap(mic, oracle, synth):
    # escapes (|diff| >= threshold -> delimiter + raw pixel), literals equal to the delimiter, and a stream whose
    # reconstruction wraps uint16 (deltagradrlecompressu16.go:120 `uint16(predicted + diff)`)
    w, h = 203, 77
    img = synth.smooth_image(5, w, h).astype(np.int64)
    rng = np.random.default_rng(11)
    ys, xs = rng.integers(0, h, 300), rng.integers(0, w, 300)
    img[ys, xs] = rng.choice([0, 4095, 4094, 1, 2047], 300)
    px = img.astype(np.uint16).ravel()
    frame = oracle.compress_single_frame_grad(px, w, h, 4095)
    assert mic.CompressSingleFrameGrad(px, w, h, 4095) == frame
    assert np.array_equal(mic.DecompressSingleFrameGrad(frame, w, h), px)
    # wrap: decode an avg-predictor stream with the gradient decoder and vice versa -- pixels are garbage, but the SAME
    # garbage as the reference arithmetic produces (including values that leave [0, 65535] and wrap)
    avg = oracle.compress_single_frame(px, w, h, 4095, 2)
    assert np.array_equal(mic.DecompressSingleFrameGrad(avg, w, h), oracle.decompress_single_frame_grad(avg, w, h))
    assert np.array_equal(mic.DecompressSingleFrame(frame, w, h), oracle.decompress_single_frame(frame, w, h))


@pytest.mark.parametrize("w,h,noisy_from,strips", [(320, 200, 120, 1), (320, 200, 120, 4), (517, 263, 100, 8), (2577, 512, 256, 8), (64, 40, None, 100)])
def test_pica_bytes_and_pixels(mic, oracle, synth, w, h, noisy_from, strips):
    img = synth.smooth_image(9, w, h, noisy_from=noisy_from)
    px, mx = img.ravel(), int(img.max())
    assert mic.AdaptiveStripBoundaries(px, w, h, strips) == oracle.pica_boundaries(px, w, h, strips)
    if strips >= h:
        # one strip per row: single-row frames of this content are rejected by FSE in the reference; both sides must fail
        with pytest.raises(Exception):
            oracle.pica_compress(px, w, h, mx, strips)
        with pytest.raises(mic.MicGpuError):
            mic.CompressParallelStripsAdaptive(px, w, h, mx, strips)
        return
    want = oracle.pica_compress(px, w, h, mx, strips)
    got = mic.CompressParallelStripsAdaptive(px, w, h, mx, strips)
    assert got == want
    if noisy_from is not None and strips >= 4:
        assert set(_flags(want)) == {0, 1}          # both predictors in one container
    out, ow, oh = mic.DecompressParallelStripsAdaptive(want)
    assert (ow, oh) == (w, h) and np.array_equal(out, px)


def test_pica_reference_images(mic, oracle):
    # docs/adaptive-compression.md:74-75: MR keeps the gradient predictor on 3 of 4 strips, CT on none
    for name, w, h, ngrad in (("MR_256_256", 256, 256, 3), ("CT_512_512", 512, 512, 0)):
        px, _, _ = _golden(name, w, h)
        mx = int(px.max())
        want = oracle.pica_compress(px, w, h, mx, 4)
        assert sum(_flags(want)) == ngrad
        assert mic.CompressParallelStripsAdaptive(px, w, h, mx, 4) == want
        out, ow, oh = mic.DecompressParallelStripsAdaptive(want)
        assert (ow, oh) == (w, h) and np.array_equal(out, px)


def test_pica_in_a_mixed_plan(mic, oracle, synth):
    # PICA and PICS containers in one decoder plan: avg and gradient units share K1-K3 and split at the predictor kernel
    w, h = 333, 150
    a = synth.smooth_image(3, w, h, noisy_from=70).ravel()
    b = synth.xr_image(4, w, h).ravel()
    pica = np.frombuffer(oracle.pica_compress(a, w, h, int(a.max()), 5), np.uint8)
    pics = np.frombuffer(oracle.pics_compress(b, w, h, int(b.max()), 3, 8), np.uint8)
    assert set(_flags(bytes(pica))) == {0, 1}
    off2 = (pica.size + 63) & ~63
    comp = np.zeros(off2 + pics.size + 256, np.uint8)
    comp[:pica.size] = pica
    comp[off2:off2 + pics.size] = pics
    import ctypes as C

    d = mic.lib.micgpu_decoder_create(0)
    assert d
    try:
        assert mic.lib.micgpu_decoder_begin(d) == 0
        ww, hh = C.c_int(), C.c_int()
        assert mic.lib.micgpu_decoder_add_pica(d, pica.ctypes.data, pica.size, 0, 0, C.byref(ww), C.byref(hh)) == 0
        assert (ww.value, hh.value) == (w, h)
        assert mic.lib.micgpu_decoder_add_pics(d, pics.ctypes.data, pics.size, off2, w * h, C.byref(ww), C.byref(hh)) == 0
        assert mic.lib.micgpu_decoder_commit(d) == 0
        out = np.zeros(2 * w * h, np.uint16)
        assert mic.lib.micgpu_decoder_run_host(d, comp.ctypes.data, comp.size, out.ctypes.data, out.size) == 0
    finally:
        mic.lib.micgpu_decoder_destroy(d)
    assert np.array_equal(out[:w * h], a) and np.array_equal(out[w * h:], b)


def test_pica_bad_containers(mic, oracle, synth):
    w, h = 200, 120
    px = synth.smooth_image(2, w, h, noisy_from=60).ravel()
    good = oracle.pica_compress(px, w, h, int(px.max()), 3)

    def patch(off, val):
        return good[:off] + int(val).to_bytes(4, "little") + good[off + 4:]

    bads = [b"PICS" + good[4:], good[:15], good[:40],          # magic, short header, truncated strip table
            patch(4, 0), patch(12, 0), patch(12, 1 << 30),       # zero width, zero / absurd strip count
            patch(16 + 16, 0),                                   # second strip starts at row 0 (first strip gets no rows)
            patch(16 + 32, h),                                   # last strip starts past the image
            patch(16 + 16 + 8, len(good)),                       # strip length beyond the blob
            patch(16 + 4, 1 << 31)]                              # strip offset beyond the blob
    for bad in bads:
        with pytest.raises(mic.MicGpuError):
            mic.DecompressParallelStripsAdaptive(bad)
    # a damaged frame inside an intact container: negative status, no hang
    dmg = bytearray(good)
    dmg[16 + 48 + 9] ^= 0xFF
    try:
        out, _, _ = mic.DecompressParallelStripsAdaptive(bytes(dmg))
    except mic.MicGpuError:
        pass
    # flags with unknown high bits: only bit 0 selects the predictor (parallelstripsadaptive.go:186)
    out, _, _ = mic.DecompressParallelStripsAdaptive(patch(16 + 12, int.from_bytes(good[28:32], "little") | 0xF0))
    assert np.array_equal(out, px)
