"""SURVEY 8(f).4, Huffman half: canonical-Huffman streams (canhuffmandecompressu16.go) and the Delta+RLE+Huffman
composition (deltarlehuffdecompressu16.go) through the C ABI, against the oracle (oracle/mic_oracle_huff.c).  The streams
come from the oracle's encoder; the decoder is determined by the stream (the header carries symbol order and code lengths)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

REF_INPUT = np.array([256, 256, 256, 1025, 457, 457, 457, 8000, 1, 65534], np.uint16)   # canhuffmancompressu16_test.go:15


def _streams():
    rng = np.random.default_rng(11)
    yield "reference test vector", REF_INPUT
    yield "one symbol", np.array([5], np.uint16)
    yield "two symbols", np.array([5, 6], np.uint16)
    yield "constant", np.full(100, 9, np.uint16)
    yield "all zero (0-bit codes)", np.zeros(1000, np.uint16)
    yield "two-letter alphabet", rng.integers(0, 2, 5000).astype(np.uint16)
    yield "8-bit uniform", rng.integers(0, 256, 20000).astype(np.uint16)
    yield "12-bit geometric", np.minimum(rng.geometric(0.02, 300000), 4095).astype(np.uint16)
    yield "16-bit uniform (list cut, escapes)", rng.integers(0, 65536, 60000).astype(np.uint16)
    yield "delimiter value in the data", np.array([4095] * 50 + [7] * 20 + [4095] + [3] * 600, np.uint16)
    yield "skewed (1-bit code, long tail)", np.where(rng.random(400000) < 0.9, 0, rng.integers(1, 3000, 400000)).astype(np.uint16)
    yield "shorter than one subsequence", rng.integers(0, 16, 40).astype(np.uint16)


@pytest.mark.parametrize("name,sym", list(_streams()), ids=[n for n, _ in _streams()])
def test_huffman_symbols(mic, oracle, name, sym, monkeypatch):
    blob = oracle.huff_compress(sym)
    want = oracle.huff_decompress(blob)
    assert np.array_equal(want, sym)
    got = mic.CanHuffmanDecompressU16(blob)
    assert got.size == sym.size and np.array_equal(got, sym)


@pytest.mark.parametrize("name,sym", list(_streams()), ids=[n for n, _ in _streams()])
def test_huffman_encoder_bytes(mic, oracle, name, sym):
    # the device encoder (histogram + bit emission on the GPU, code construction on the host) against the restatement:
    # same stream byte for byte, and the stream decodes on the device
    want = oracle.huff_compress(sym)
    got = mic.CanHuffmanCompressU16(sym)
    assert got == want
    assert np.array_equal(mic.CanHuffmanDecompressU16(got), sym)


def test_delta_rle_huff_encoder_bytes(mic, oracle, synth):
    for (w, h, img) in [(5, 2, REF_INPUT), (333, 217, synth.xr_image(9, 333, 217).ravel()),
                        (256, 256, np.fromfile(os.path.join(GOLDEN, "MR_256_256_image.bin"), dtype="<u2"))]:
        mx = int(img.max())
        got = mic.DeltaRleHuffCompressU16(img, w, h, mx)
        assert got == oracle.delta_rle_huff_compress(img, w, h, mx)
        assert np.array_equal(mic.DeltaRleHuffDecompressU16(got, w, h), img)


def test_huffman_encoder_large_and_small_buffers(mic, oracle):
    rng = np.random.default_rng(17)
    sym = np.minimum(rng.geometric(0.05, 3_000_000), 1023).astype(np.uint16)      # 733 chunks: the chunk scan and word seams
    blob = mic.CanHuffmanCompressU16(sym)
    assert blob == oracle.huff_compress(sym)
    out = np.empty(16, np.uint8)
    import ctypes as C
    got = C.c_size_t()
    rc = mic.lib.micgpu_huff_compress(sym.ctypes.data, sym.size, out.ctypes.data, out.size, C.byref(got))
    assert rc == mic.api.E_SIZE and got.value == len(blob)


def test_huffman_multi_tile_stream(mic, oracle):
    # 4 M symbols: ~250 tiles of 256 subsequences; the carried start of every tile and the running output offset are exercised
    rng = np.random.default_rng(3)
    sym = np.minimum(rng.geometric(0.05, 4_000_000), 1023).astype(np.uint16)
    blob = oracle.huff_compress(sym)
    assert np.array_equal(mic.CanHuffmanDecompressU16(blob), sym)


@pytest.mark.parametrize("name,w,h", [("MR_256_256_image.bin", 256, 256), ("CT_512_512_image.bin", 512, 512)])
def test_delta_rle_huff_reference_images(mic, oracle, name, w, h):
    img = np.fromfile(os.path.join(GOLDEN, name), dtype="<u2")
    blob = oracle.delta_rle_huff_compress(img, w, h, int(img.max()))
    assert np.array_equal(oracle.delta_rle_huff_decompress(blob, w, h).ravel(), img)
    assert np.array_equal(mic.DeltaRleHuffDecompressU16(blob, w, h), img)


@pytest.mark.parametrize("w,h", [(5, 2), (1, 1), (1, 40), (40, 1), (7, 5), (333, 217), (1000, 64)])
def test_delta_rle_huff_geometries(mic, oracle, synth, w, h):
    img = REF_INPUT if (w, h) == (5, 2) else (synth.xr_image(w * 31 + h, w, h).ravel() if w * h > 64 else
                                                 (np.arange(w * h, dtype=np.uint16) % 9 + 100))
    blob = oracle.delta_rle_huff_compress(img, w, h, int(img.max()))
    assert np.array_equal(mic.DeltaRleHuffDecompressU16(blob, w, h), img)


def test_huffman_units_in_a_plan_beside_fse_units(mic, oracle, synth):
    # one plan, three units: an FSE frame, a Huffman frame of another image, a Huffman-coded RLE stream
    a, b = synth.xr_image(1, 200, 120).ravel(), synth.xr_image(2, 96, 80).ravel()
    fa = oracle.compress_single_frame(a, 200, 120, int(a.max()), 2)
    hb = oracle.delta_rle_huff_compress(b, 96, 80, int(b.max()))
    res = (np.arange(5000) % 7 == 0).astype(np.uint16) * 300
    hr = oracle.huff_compress(oracle.rle_compress(res, 300))
    assert np.array_equal(oracle.rle_decompress(oracle.huff_decompress(hr)), res)
    comp = np.frombuffer(fa + bytes(-len(fa) % 64) + hb + bytes(-len(hb) % 64) + hr + bytes(256), np.uint8)
    o1 = len(fa) + (-len(fa) % 64)
    o2 = o1 + len(hb) + (-len(hb) % 64)
    d = mic.Decoder()
    d.begin()
    d.add_unit(fa, 0, 0, 200, 120, 0)
    d.add_huff_unit(hb, o1, 0, 96, 80, 24000)
    d.add_huff_unit(hr, o2, 1, 5000, 1, 24000 + 96 * 80)
    d.commit()
    out = np.zeros(24000 + 96 * 80 + 5000, np.uint16)
    d.run_host(comp, out)
    assert np.array_equal(out[:24000], a) and np.array_equal(out[24000:24000 + 7680], b) and np.array_equal(out[31680:], res)
    d.close()


def test_huffman_corrupt_streams(mic, oracle):
    rng = np.random.default_rng(5)
    sym = np.minimum(rng.geometric(0.05, 50000), 1023).astype(np.uint16)
    blob = oracle.huff_compress(sym)
    with pytest.raises(mic.MicGpuError):
        mic.CanHuffmanDecompressU16(blob[:8])                                   # shorter than the fixed header
    with pytest.raises(mic.MicGpuError):
        mic.CanHuffmanDecompressU16(blob[: len(blob) // 2])                     # the bits run out
    with pytest.raises(mic.MicGpuError):
        mic.CanHuffmanDecompressU16(blob[:6] + bytes([40]) + blob[7:])          # maxCodeLength 40
    bad = bytearray(blob)
    bad[7], bad[8] = 0xFF, 0xFF                                                 # a symbol list longer than the stream
    with pytest.raises(mic.MicGpuError):
        mic.CanHuffmanDecompressU16(bytes(bad))
    # seeded bit flips: the call returns (symbols or an error), and an undamaged stream decoded afterwards is exact
    for seed in range(24):
        r = np.random.default_rng(seed)
        dmg = bytearray(blob)
        for _ in range(3):
            i = int(r.integers(4, len(dmg)))                                    # keep the symbol count
            dmg[i] ^= 1 << int(r.integers(0, 8))
        try:
            mic.CanHuffmanDecompressU16(bytes(dmg))
        except mic.MicGpuError:
            pass
    assert np.array_equal(mic.CanHuffmanDecompressU16(blob), sym)
