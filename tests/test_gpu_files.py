"""the .mic file wrappers of cmd/mic-compress (MIC1 = 20-byte header + one frame, MICR = 12-byte header +
CompressRGB blob) and the magic dispatch of cmd/mic-wasm's decodeMicFile, through the C ABI."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _mic1_by_hand(oracle, img, w, h, mx, nstates):
    frame = oracle.compress_single_frame(img, w, h, mx, nstates)
    return b"MIC1" + w.to_bytes(4, "little") + h.to_bytes(4, "little") + (1).to_bytes(4, "little") + len(frame).to_bytes(4, "little") + frame


@pytest.mark.parametrize("nstates", [2, 4, 8])
def test_mic1_file_bytes_and_round_trip(mic, oracle, nstates):
    # the reference's own MR image: the file the CLI would write (writeMicFile over compressImage[4State], main.go:26-105)
    img = np.fromfile(os.path.join(GOLDEN, "MR_256_256_image.bin"), dtype="<u2")
    mx = int(img.max())
    want = _mic1_by_hand(oracle, img, 256, 256, mx, nstates)
    got = mic.WriteMIC1(img, 256, 256, mx, nstates)
    assert got == want
    kind, px, w, h = mic.DecodeMicFile(got)
    assert (kind, w, h) == ("MIC1", 256, 256) and np.array_equal(px, img)


def test_micr_file(mic, oracle, synth):
    rgb = synth.wsi_region(5, 1000, 900, 301, 203, 2500, 2000)
    blob = oracle.rgb_compress(rgb.ravel(), 301, 203, True)
    want = b"MICR" + (301).to_bytes(4, "little") + (203).to_bytes(4, "little") + blob
    got = mic.WriteMICR(rgb, 301, 203)
    assert got == want
    kind, px, w, h = mic.DecodeMicFile(got)
    assert (kind, w, h) == ("MICR", 301, 203) and np.array_equal(px, rgb.ravel())


def test_magic_dispatch_and_bad_files(mic, oracle, synth):
    img = synth.xr_image(3, 200, 120).ravel()
    pics = oracle.pics_compress(img, 200, 120, int(img.max()), 4, 2)
    kind, px, w, h = mic.DecodeMicFile(pics)
    assert (kind, w, h) == ("PICS", 200, 120) and np.array_equal(px, img)
    st = synth.tomo_stack(7, 3, 64, 64)
    kind, frames, w, h = mic.DecodeMicFile(oracle.mic2_compress(st.ravel(), 64, 64, 1023, True))
    assert kind == "MIC2" and np.array_equal(np.asarray(frames).reshape(st.shape), st)
    good = mic.WriteMIC1(img, 200, 120, int(img.max()), 2)
    for bad in (b"MICX" + good[4:], good[:19], good[:12] + (2).to_bytes(4, "little") + good[16:],      # magic, size, pipeline
                good[:16] + (len(good)).to_bytes(4, "little") + good[20:], good[:4] + bytes(8) + good[12:]):   # length, dimensions
        with pytest.raises(mic.MicGpuError):
            mic.DecodeMicFile(bad)
    assert mic.lib.micgpu_file_kind(None, 0) == 0
