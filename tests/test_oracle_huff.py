"""Canonical-Huffman back end of the oracle (oracle/mic_oracle_huff.c): the reference's own test inputs
(canhuffmancompressu16_test.go:11-108), the stream layout of WriteTable (canhuffmancompressu16.go:119-137) checked field by
field, and the properties a canonical prefix code must have.  Encoder bytes are unpinned against Go (tie order of
sort.Slice); the decoder is determined by the stream."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

REF_INPUT = np.array([256, 256, 256, 1025, 457, 457, 457, 8000, 1, 65534], np.uint16)   # canhuffmancompressu16_test.go:15,53


def _bits(blob, pos, n):
    v = 0
    for i in range(n):
        v = (v << 1) | ((blob[(pos + i) >> 3] >> (7 - ((pos + i) & 7))) & 1)
    return v


def _header(blob):
    n, mx, ml, nl = _bits(blob, 0, 32), _bits(blob, 32, 16), _bits(blob, 48, 8), _bits(blob, 56, 16)
    depth, lb = int(mx).bit_length(), int(ml).bit_length()
    syms = [_bits(blob, 72 + j * depth, depth) for j in range(nl)]
    lens = [_bits(blob, 72 + nl * depth + j * lb, lb) for j in range(nl)]
    return n, mx, ml, syms, lens, 72 + nl * (depth + lb)


def test_reference_test_vector_round_trip_and_layout(oracle):
    blob = oracle.huff_compress(REF_INPUT)
    assert np.array_equal(oracle.huff_decompress(blob), REF_INPUT)
    n, mx, ml, syms, lens, data_at = _header(blob)
    assert (n, mx) == (10, 65534)
    # every distinct symbol is in the list (6 symbols cost at most 14 bits) plus the delimiter 2^16 - 1, which no input
    # symbol equals here: its frequency is 0 and it still gets a code (AddDelimiterToSymbolList, :190-206)
    assert sorted(syms) == [1, 256, 457, 1025, 8000, 65534, 65535]
    assert ml == max(lens) and sum(2.0 ** -l for l in lens) == 1.0      # a complete prefix code
    by = dict(zip(syms, lens))
    assert by[256] <= by[1025] and by[457] <= by[8000]                   # more frequent symbols never get longer codes
    assert lens == sorted(lens, reverse=True)                            # the list is written rarest first (ascending frequency)
    # the body: every symbol's code, then maxCodeLength + pixelDepth zero bits, then zero padding to a byte (:65-80)
    body = sum(by[int(s)] for s in REF_INPUT)
    assert len(blob) == (data_at + body + ml + 16 + 7) // 8


def test_delta_rle_huff_reference_vector(oracle):
    # TestDeltaRLEHuffmanCompression (canhuffmancompressu16_test.go:51-108): 5 x 2 image, maxValue 65534
    sym = oracle.delta_rle_compress(REF_INPUT, 5, 2, 65534)
    blob = oracle.huff_compress(sym)
    assert np.array_equal(oracle.huff_decompress(blob), sym)
    assert blob == oracle.delta_rle_huff_compress(REF_INPUT, 5, 2, 65534)
    assert np.array_equal(oracle.delta_rle_huff_decompress(blob, 5, 2).ravel(), REF_INPUT)


@pytest.mark.parametrize("name,w,h", [("MR_256_256_image.bin", 256, 256), ("CT_512_512_image.bin", 512, 512)])
def test_reference_images(oracle, name, w, h):
    img = np.fromfile(os.path.join(GOLDEN, name), dtype="<u2")
    blob = oracle.delta_rle_huff_compress(img, w, h, int(img.max()))
    assert np.array_equal(oracle.delta_rle_huff_decompress(blob, w, h).ravel(), img)
    n, mx, ml, syms, lens, _ = _header(blob)
    assert ml <= 16 and len(set(syms)) == len(syms)
    # Huffman over the Delta+RLE symbols lands near the FSE result on the same symbols (README.md: both ~2.2-2.4x on MR)
    fse = oracle.compress_single_frame(img, w, h, int(img.max()), 2)
    assert 0.85 < len(blob) / len(fse) < 1.25


def test_escape_and_degenerate_streams(oracle):
    rng = np.random.default_rng(7)
    # > 2^14 distinct symbols: OptimizeSymbolCount (:168-186) must cut the list, the rest escapes through the delimiter
    x = rng.integers(0, 65536, 60000).astype(np.uint16)
    blob = oracle.huff_compress(x)
    n, mx, ml, syms, lens, _ = _header(blob)
    assert len(syms) < 40000 and 65535 in syms and np.array_equal(oracle.huff_decompress(blob), x)
    # symbols equal to the delimiter value itself are escaped like any other unlisted symbol
    y = np.array([4095] * 50 + [7] * 20 + [4095], np.uint16)
    assert np.array_equal(oracle.huff_decompress(oracle.huff_compress(y)), y)
    # one distinct symbol; all zeros (pixelDepth 0: every symbol costs no bits at all); a single symbol
    for z in (np.full(100, 9, np.uint16), np.zeros(1000, np.uint16), np.array([5], np.uint16)):
        b = oracle.huff_compress(z)
        assert np.array_equal(oracle.huff_decompress(b), z)
    assert len(oracle.huff_compress(np.zeros(1000, np.uint16))) == 9


def test_corrupt_streams_are_rejected(oracle):
    from oracle.oracle import OracleError

    blob = oracle.huff_compress(np.arange(300, dtype=np.uint16) % 37)
    for bad in (blob[:8], blob[: len(blob) - 3], blob[:6] + bytes([40]) + blob[7:]):
        with pytest.raises(OracleError):
            oracle.huff_decompress(bad)
