"""CPU oracle vs every known-answer value and deterministic round-trip input the reference's own tests
hold for the hot path.  Reference test cited next to each case."""
import hashlib
import json
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


# ---- value-level pins ---------------------------------------------------------------------
def test_ycocg_known_values(oracle):
    # TestYCoCgRKnownValues wsi_test.go:197-215: RGB(200,100,50) -> Y=112, Co=ZigZag(150)=300, Cg=ZigZag(-25)=49
    y, co, cg = oracle.ycocg_forward(np.array([200, 100, 50], np.uint8))
    assert (int(y[0]), int(co[0]), int(cg[0])) == (112, 300, 49)
    assert oracle.ycocg_inverse(y, co, cg).tolist() == [200, 100, 50]


def test_ycocg_exhaustive_8bit(oracle):
    # TestYCoCgRExhaustive8Bit wsi_test.go:170: all 2^24 colours round-trip; Y in [0,255], ZigZag(Co/Cg) in [0,510]
    v = np.arange(1 << 24, dtype=np.uint32)
    rgb = np.stack([(v >> 16) & 255, (v >> 8) & 255, v & 255], axis=1).astype(np.uint8).ravel()
    y, co, cg = oracle.ycocg_forward(rgb)
    assert y.max() <= 255 and co.max() <= 510 and cg.max() <= 510
    assert np.array_equal(oracle.ycocg_inverse(y, co, cg), rgb)


def test_downsample_rgb_known(oracle):
    # TestDownsample2xRGB wsi_test.go:229: 2x2 block (10,20,100),(20,30,100),(30,40,100),(40,30,100) -> (25,30,100)
    src = np.array([10, 20, 100, 20, 30, 100, 30, 40, 100, 40, 30, 100], np.uint8)
    dst, w, h = oracle.downsample2x_rgb(src, 2, 2)
    assert (w, h) == (1, 1) and dst.tolist() == [25, 30, 100]


def test_downsample_odd_dimensions(oracle):
    # TestDownsample2xOddDimensions wsi_test.go:256: 5x3 -> 2x1 (odd trailing row/col dropped)
    src = np.arange(5 * 3 * 3, dtype=np.uint8)
    dst, w, h = oracle.downsample2x_rgb(src, 5, 3)
    assert (w, h) == (2, 1) and dst.size == 6
    g, w, h = oracle.downsample2x_grey(np.array([1, 2, 3, 4, 5, 6, 7, 8, 9], np.uint16), 3, 3)
    assert (w, h) == (1, 1) and g.tolist() == [(1 + 2 + 4 + 5 + 2) // 4]
    _, w, h = oracle.downsample2x_grey(np.array([1, 2], np.uint16), 2, 1)
    assert (w, h) == (0, 0)


def test_temporal_delta_edge_cases(oracle):
    # TestTemporalDeltaEdgeCases multiframe_test.go:48-60: {0,65535,32768,100} vs {1,65535,65535,1}... exact by wrap
    prev = np.array([1, 65535, 65535, 1], np.uint16)
    cur = np.array([0, 65535, 32768, 100], np.uint16)
    res = oracle.temporal_encode(cur, prev)
    assert np.array_equal(oracle.temporal_decode(res, prev), cur)
    # every pair round-trips because both directions wrap mod 2^16
    rng = np.random.default_rng(0)
    a, b = rng.integers(0, 65536, 4096).astype(np.uint16), rng.integers(0, 65536, 4096).astype(np.uint16)
    assert np.array_equal(oracle.temporal_decode(oracle.temporal_encode(a, b), b), a)


def test_zigzag_roundtrip(oracle):
    # TestZigZagRoundTrip waveletu16_test.go:57-67 and deltazigzagcompressu16.go:108-116
    for x in (0, 1, -1, 2, -2, 127, -128, 32767, -32768, 12345, -12345):
        z = oracle.zigzag(x)
        assert oracle.unzigzag(z) == x
    assert [oracle.zigzag(v) for v in (0, -1, 1, -2, 2)] == [0, 1, 2, 3, 4]


# ---- literal vectors of the stage tests ----------------------------------------------------
def test_rle_literal_vector(oracle):
    # TestRleCompressionU16 rlecompressu16_test.go:11: {256,256,256,1025,457,457,457,8000,1}, Init(3,3,8000)
    src = np.array([256, 256, 256, 1025, 457, 457, 457, 8000, 1], np.uint16)
    enc = oracle.rle_compress(src, 8000)
    assert enc[0] == 8000 and (int(enc[1]) << 16 | int(enc[2])) == 9
    assert np.array_equal(oracle.rle_decompress(enc), src)


def test_delta_rle_literal_vector(oracle):
    # TestDeltaRleCompression deltacompressu16_test.go:35: 3x3, max 8000
    src = np.array([256, 300, 468, 1025, 457, 399, 4096, 8000, 1], np.uint16)
    sym = oracle.delta_rle_compress(src, 3, 3, 8000)
    assert sym[0] == 8191                       # delimiter = (1<<13)-1 is word 0 (deltarlecompressu16.go:28)
    assert np.array_equal(oracle.delta_rle_decompress(sym, 3, 3), src)


def test_rle_long_constant_runs_split(oracle):
    # rlecompressu16.go:58: forced flush at len(buf) >= midCount-1 splits constant runs into midCount-3 pieces
    src = np.full(10000, 77, np.uint16)
    enc = oracle.rle_compress(src, 255)          # depth 8 -> midCount 127
    body = enc[3:]
    assert body[0] == 127 - 3 and body[1] == 77
    assert np.array_equal(oracle.rle_decompress(enc), src)


# ---- FSE tail-residue inputs and magic bytes -------------------------------------------------
@pytest.mark.parametrize("n", [999, 101, 102, 103, 1001, 1002, 1003])
@pytest.mark.parametrize("mod", [17, 8])
def test_fse2_edge_lengths(oracle, n, mod):
    # TestFSE2StateEdgeCases fse2state_test.go:178
    src = (np.arange(n) % mod).astype(np.uint16)
    for coder in (1, 2, 4):
        blob = oracle.fse_compress(src, coder)
        assert np.array_equal(oracle.fse_decompress(blob), src)


@pytest.mark.parametrize("n", [9, 10, 11, 12, 13, 14, 15, 100, 101, 102, 103, 104, 105, 106, 107, 1001, 1007])
def test_fse8_edge_lengths(oracle, n):
    # TestFSE8StateEdgeCases fse8state_test.go:106 (i%4) and TestRANS8StateRoundtrip
    from oracle.oracle import OracleError

    src = (np.arange(n) % 4).astype(np.uint16)
    for coder in (8, 108):
        try:
            blob = oracle.fse_compress(src, coder)
        except OracleError as e:
            assert e.code == -1        # ErrIncompressible for the shortest inputs, as in Go
            continue
        assert np.array_equal(oracle.fse_decompress(blob), src)


def test_fse_rejects(oracle):
    from oracle.oracle import ERR_INCOMPRESSIBLE, ERR_USE_RLE, OracleError

    with pytest.raises(OracleError) as e:
        oracle.fse_compress(np.full(1000, 5, np.uint16), 2)     # constant -> ErrUseRLE (fse2state_test.go:184)
    assert e.value.code == ERR_USE_RLE
    with pytest.raises(OracleError) as e:
        oracle.fse_compress(np.array([1, 2], np.uint16), 2)     # 2 elements -> error (fse2state_test.go:192)
    assert e.value.code == ERR_INCOMPRESSIBLE
    with pytest.raises(OracleError) as e:
        oracle.fse_compress(np.arange(7, dtype=np.uint16), 8)   # len <= 7 (fse8state.go:33)
    assert e.value.code == ERR_INCOMPRESSIBLE


def test_magic_bytes_and_autodetect(oracle):
    # Test*MagicBytes / Test*AutoDetect fse8state_test.go:46,78: [0xFF,0x02/0x04/0x84/0x08] + count u32 LE
    src = (np.arange(5000) % 23).astype(np.uint16)
    for coder, magic in ((2, 0x02), (4, 0x04), (8, 0x84), (108, 0x08)):
        blob = oracle.fse_compress(src, coder)
        assert blob[0] == 0xFF and blob[1] == magic and struct.unpack_from("<I", blob, 2)[0] == 5000
        assert np.array_equal(oracle.fse_decompress(blob), src)
    one = oracle.fse_compress(src, 1)
    assert np.array_equal(oracle.fse_decompress(one), src)
    bad = bytearray(oracle.fse_compress(src, 8))
    bad[1] = 0x55                                              # corrupt magic -> falls to 1-state parse -> error or garbage
    try:
        out = oracle.fse_decompress(bytes(bad))
        assert not np.array_equal(out, src)
    except Exception:
        pass


# ---- wavelet: deterministic inputs and dimension lists ----------------------------------------
@pytest.mark.parametrize("n", [2, 3, 4, 5, 8, 15, 16, 17, 31, 64, 100, 255, 256])
def test_wavelet_1d_roundtrip(oracle, n):
    # TestWavelet1DRoundTrip waveletu16_test.go:14: input i*37+100
    src = (np.arange(n) * 37 + 100).astype(np.int32)
    fwd = oracle.wt53_forward_1d(src)
    assert np.array_equal(oracle.wt53_inverse_1d(fwd), src)


@pytest.mark.parametrize("rows,cols", [(4, 4), (5, 7), (8, 8), (16, 9), (63, 65), (64, 32), (256, 256)])
def test_wavelet_2d_separated_roundtrip(oracle, rows, cols):
    # TestWavelet2DSeparatedRoundTrip waveletu16_test.go:190: input (i*131+7)%65536
    src = ((np.arange(rows * cols) * 131 + 7) % 65536).astype(np.int32)
    fwd = oracle.wt53_forward_2d(src, rows, cols)
    assert np.array_equal(oracle.wt53_inverse_2d(fwd, rows, cols), src)


def test_wavelet_lifting_known_small(oracle):
    # 5/3 lifting by hand (waveletu16.go:26-71): [10,20,30,40] -> d0 = 20-((10+30)>>1)=0, d1 = 40-((30+30)>>1)=10,
    # s0 = 10+((0+0+2)>>2)=10, s1 = 30+((0+10+2)>>2)=33
    assert oracle.wt53_forward_1d(np.array([10, 20, 30, 40], np.int32)).tolist() == [10, 0, 33, 10]


@pytest.mark.parametrize("levels", [1, 2, 3, 5])
@pytest.mark.parametrize("dim", [8, 32, 64, 128])
def test_wavelet_multilevel_transform_roundtrip(oracle, levels, dim):
    # TestWaveletV2MultiLevelRoundTrip waveletu16_test.go:212 (transform only, top-left LL recursion, early stop at r<2||c<2)
    src = ((np.arange(dim * dim) * 131 + 7) % 65536).astype(np.int32)
    data, dims, r, c = src.copy(), [], dim, dim
    for _ in range(levels):
        if r < 2 or c < 2:
            break
        dims.append((r, c))
        data = oracle.wt53_forward_2d(data, r, c, dim)
        r, c = (r + 1) // 2, (c + 1) // 2
    for r, c in reversed(dims):
        data = oracle.wt53_inverse_2d(data, r, c, dim)
    assert np.array_equal(data, src)


@pytest.mark.parametrize("rows,cols,levels", [(64, 96, 3), (96, 64, 5), (128, 128, 5), (127, 129, 4), (256, 512, 8), (3, 2000, 5), (300, 200, 1)])
def test_wavelet_v2_multilevel_roundtrip(oracle, rows, cols, levels):
    # full WaveletV2 pipeline on a smooth field + small deterministic texture (the (i*97+13)%4096 ramp of
    # waveletu16_test.go:352 has more distinct coefficients than FSE table cells at these sizes and is rejected
    # by FSECompressU16FourState, exactly as in Go: the V2 pipeline has no fallback)
    i = np.arange(rows * cols)
    y, x = i // cols, i % cols
    src = ((y * 5 + x * 3) % 3000 + (i * 97 + 13) % 7).astype(np.uint16)
    blob = oracle.wavelet_v2_compress(src, rows, cols, 4095, levels)
    r, c, mx, lv = struct.unpack_from("<IIHB", blob, 0)
    assert (r, c, mx) == (rows, cols, 4095) and 1 <= lv <= levels
    assert blob[11] == 0xFF and blob[12] == 0x04           # 4-state FSE, no fallback (waveletfsecompressu16.go:353)
    px, rr, cc = oracle.wavelet_v2_decompress(blob)
    assert (rr, cc) == (rows, cols) and np.array_equal(px, src)


def test_wavelet_escape_coefficients(oracle):
    # coefficients beyond +-32767 use the 65535,hi,lo escape triple (waveletfsecompressu16.go:31-37)
    rng = np.random.default_rng(3)
    src = rng.integers(0, 2, 64 * 64).astype(np.uint16) * 65535
    blob = oracle.wavelet_v2_compress(src, 64, 64, 65535, 3)
    px, _, _ = oracle.wavelet_v2_decompress(blob)
    assert np.array_equal(px, src)


# ---- containers --------------------------------------------------------------------------------
def test_mic2_header_layout(oracle, synth):
    # TestMIC2HeaderRoundtrip multiframe_test.go:62: "MIC2" w h n u32, flags (bit0 spatial, bit1 temporal), 3 reserved
    st = synth.tomo_stack(7, 5, 128, 128)
    for temporal in (False, True):
        blob = oracle.mic2_compress(st.ravel(), 128, 128, 1023, temporal)
        assert blob[:4] == b"MIC2" and struct.unpack_from("<3I", blob, 4) == (128, 128, 5)
        assert blob[16] == (0x03 if temporal else 0x01) and blob[17:20] == b"\0\0\0"
        offs = [struct.unpack_from("<2I", blob, 20 + 8 * i) for i in range(5)]
        assert offs[0][0] == 0 and all(offs[i + 1][0] == offs[i][0] + offs[i][1] for i in range(4))
        assert 20 + 40 + offs[-1][0] + offs[-1][1] == len(blob)
        frames, t = oracle.mic2_decompress(blob)
        assert t == temporal and np.array_equal(frames, st)
        for idx in range(5):
            assert np.array_equal(oracle.mic2_decompress_frame(blob, idx), st[idx])


def test_pics_format_and_clamps(oracle, synth):
    # TestParallelStripsFormatValidation parallelstrips_test.go:82, TestParallelStripsSingleRowImage :121
    img = synth.xr_image(2, 297, 203).ravel()
    blob = oracle.pics_compress(img, 297, 203, int(img.max()), 8, 2)
    w, h, n, sh = struct.unpack_from("<4I", blob, 4)
    assert blob[:4] == b"PICS" and (w, h) == (297, 203) and sh == (203 + 7) // 8 and n == (203 + sh - 1) // sh
    px, ow, oh = oracle.pics_decompress(blob)
    assert (ow, oh) == (297, 203) and np.array_equal(px, img)
    from oracle.oracle import OracleError

    with pytest.raises(OracleError):
        oracle.pics_decompress(b"PICX" + blob[4:])
    with pytest.raises(OracleError):
        oracle.pics_decompress(blob[:24])
    # numStrips > height clamps to height (parallelstrips.go:62-64)
    row = synth.xr_image(5, 4000, 1).ravel()
    one = oracle.pics_compress(row, 4000, 1, int(row.max()), 16, 2)
    assert struct.unpack_from("<4I", one, 4) == (4000, 1, 1, 1)
    assert np.array_equal(oracle.pics_decompress(one)[0], row)


def test_mic3_header_and_tiles(oracle, synth):
    # TestMIC3HeaderRoundtrip wsi_test.go:289, TestWSICompressOddDimensions, TestWSIPyramidLevels, TestWSIRegionCrossTile :493-780
    W, H = 600, 391
    rgb = synth.wsi_region(11, 100, 50, W, H, 2048, 1024)
    blob = oracle.wsi_compress(rgb.ravel(), W, H, 3, 8, 256, 256, 0)
    hdr = oracle.wsi_header(blob)
    assert blob[:4] == b"MIC3" and struct.unpack_from("<I", blob, 4)[0] == 1
    assert (hdr["width"], hdr["height"], hdr["tile_w"], hdr["tile_h"], hdr["channels"], hdr["bps"]) == (W, H, 256, 256, 3, 8)
    assert hdr["color_transform"] == 1
    # autoLevelCount: 600x391 -> 300x195 -> 150x97 (fits one tile) = 3 levels; floor halving (wsiformat.go:260-285)
    assert hdr["nlevels"] == 3
    assert [lv[:4] for lv in hdr["levels"]] == [(600, 391, 3, 2), (300, 195, 2, 1), (150, 97, 1, 1)]
    assert [lv[4] for lv in hdr["levels"]] == [0, 6, 8] and hdr["total_tiles"] == 9
    # level-0 tiles reproduce the source (edge tiles cropped, wsicompress.go:200-216)
    for ty in range(2):
        for tx in range(3):
            t, tw, th = oracle.wsi_decompress_tile(blob, 0, tx, ty)
            exp = rgb[ty * 256:ty * 256 + th, tx * 256:tx * 256 + tw]
            assert (tw, th) == (exp.shape[1], exp.shape[0]) and np.array_equal(t.reshape(th, tw, 3), exp)
    # region crossing four tiles
    reg, rw, rh = oracle.wsi_decompress_region(blob, 0, 200, 180, 150, 120)
    assert (rw, rh) == (150, 120) and np.array_equal(reg.reshape(120, 150, 3), rgb[180:300, 200:350])
    # level 1 equals the 2x2 box filter of level 0
    ds, dw, dh = oracle.downsample2x_rgb(rgb.ravel(), W, H)
    reg1, _, _ = oracle.wsi_decompress_region(blob, 1, 0, 0, dw, dh)
    assert np.array_equal(reg1, ds)


def test_wsi_plane_modes(oracle):
    # plane modes 0/1/2/3 (wsicompress.go:17-22,373-421)
    n = 64 * 64
    assert oracle.wsi_plane_compress(np.zeros(n, np.uint16), 64, 64) == b"\x00"
    assert oracle.wsi_plane_compress(np.full(n, 300, np.uint16), 64, 64) == b"\x01" + struct.pack("<H", 300)
    rng = np.random.default_rng(1)
    smooth = (np.arange(n) % 64 + rng.integers(0, 3, n)).astype(np.uint16)
    b2 = oracle.wsi_plane_compress(smooth, 64, 64)
    assert b2[0] == 2 and np.array_equal(oracle.wsi_plane_decompress(b2, 64, 64), smooth)
    # raw fallback needs ErrIncompressible from every FSE tier (wsicompress.go:403-414): a 2x2 plane whose
    # symbol stream has no repeated symbol (maxCount == 1, fsecompressu16.go:43)
    tiny = np.array([300, 10, 20, 35], np.uint16)
    b3 = oracle.wsi_plane_compress(tiny, 2, 2)
    assert b3[0] == 3 and len(b3) == 1 + 2 * 4 and np.array_equal(oracle.wsi_plane_decompress(b3, 2, 2), tiny)


def test_rgb_compress_roundtrip(oracle, synth):
    # CompressRGB / DecompressRGB rgbcompress.go:25-33
    rgb = synth.wsi_region(3, 0, 0, 96, 80, 96, 80)
    blob = oracle.rgb_compress(rgb.ravel(), 96, 80, True)
    ly, lco, lcg = struct.unpack_from("<3I", blob, 0)
    assert 12 + ly + lco + lcg == len(blob)
    assert np.array_equal(oracle.rgb_decompress(blob, 96, 80, True), rgb.ravel())


# ---- golden streams -----------------------------------------------------------------------------
def test_golden_streams(oracle, synth):
    """Lengths + sha256 of compressed streams; the 2/4/8-state entries were produced by the reference's own
    C encoder (tests/golden/make_golden.py asserts oracle == reference twin byte for byte)."""
    gold = json.load(open(os.path.join(GOLDEN, "streams.json")))
    inputs = {
        "CT_512_512": (np.fromfile(os.path.join(GOLDEN, "CT_512_512_image.bin"), np.uint16), 512, 512),
        "MR_256_256": (np.fromfile(os.path.join(GOLDEN, "MR_256_256_image.bin"), np.uint16), 256, 256),
        "xr_seed1_611x403": (synth.xr_image(1, 611, 403).ravel(), 611, 403),
    }
    for name, (img, w, h) in inputs.items():
        g = gold[name]
        assert sha(img.tobytes()) == g["pixels_sha256"]
        mx = int(img.max())
        for ns in (1, 2, 4, 8):
            a = oracle.compress_single_frame(img, w, h, mx, ns)
            assert (len(a), sha(a)) == (g[f"frame_{ns}state"]["len"], g[f"frame_{ns}state"]["sha256"]), (name, ns)
            assert np.array_equal(oracle.decompress_single_frame(a, w, h), img)
        sym = oracle.delta_rle_compress(img, w, h, mx)
        assert (int(sym.size), sha(sym.tobytes())) == (g["delta_rle_symbols"]["len"], g["delta_rle_symbols"]["sha256"])
        a = oracle.fse_compress(sym, 108)
        assert sha(a) == g["frame_rans8"]["sha256"]
        a = oracle.pics_compress(img, w, h, mx, 8, 2)
        assert sha(a) == g["pics8_2state"]["sha256"]
        a = oracle.wavelet_v2_compress(img, h, w, mx, 5)
        assert sha(a) == g["wavelet_v2_5lv"]["sha256"]
        assert np.array_equal(oracle.wavelet_v2_decompress(a)[0], img)
    # published ratios: CT 2.24x, MR 2.35x (README.md:269-270)
    assert round(524288 / gold["CT_512_512"]["frame_2state"]["len"], 2) == 2.24
    assert round(131072 / gold["MR_256_256"]["frame_2state"]["len"], 2) == 2.35
    st = synth.tomo_stack(7, 5, 128, 128)
    for temporal in (False, True):
        a = oracle.mic2_compress(st.ravel(), 128, 128, 1023, temporal)
        assert sha(a) == gold[f"mic2_tomo_seed7_5x128x128_temporal{int(temporal)}"]["sha256"]
    rgb = synth.wsi_region(11, 300, 200, 512, 384, 1024, 768)
    a = oracle.wsi_compress(rgb.ravel(), 512, 384, 3, 8, 256, 256, 0)
    assert sha(a) == gold["mic3_wsi_seed11_512x384"]["sha256"]


def test_oracle_matches_reference_twin_bytes(oracle, reftwin, synth):
    """Live cross-check against the compiled reference C encoder/decoder (oracle/_ref)."""
    cases = [(np.fromfile(os.path.join(GOLDEN, "MR_256_256_image.bin"), np.uint16), 256, 256),
             (synth.xr_image(9, 333, 257).ravel(), 333, 257)]
    for img, w, h in cases:
        for ns in (2, 4, 8):
            a = oracle.compress_single_frame(img, w, h, int(img.max()), ns)
            assert a == reftwin.compress(img, w, h, ns)
            assert np.array_equal(reftwin.decompress(a, w, h, ns, simd=True), img)
            assert np.array_equal(reftwin.decompress(a, w, h, ns, simd=False), img)


# ---- gradient-adaptive predictor and PICA ----------------------------------------------------------------
def test_grad_predictor_reference_vector(oracle):
    # TestGradDeltaCompress (deltagradcompressu16_test.go:12-30): the 3x3 ramp round-trips; the symbols are checked by hand:
    # row 0 predicts from the left, column 0 from the top, the interior from gradPredict (deltagradcompressu16.go:149-166)
    px = np.array([100, 110, 120, 105, 115, 125, 110, 120, 130], np.uint16)
    sym = oracle.grad_delta_rle_compress(px, 3, 3, 8000)
    thr = (1 << 12) - 1                       # bits.Len16(8000) = 13
    assert sym[0] == (1 << 13) - 1            # RleCompressU16.Init word: the delimiter (deltagradrlecompressu16.go:30)
    assert np.array_equal(oracle.grad_delta_rle_decompress(sym, 3, 3), px)
    # gradPredict(w=105, n=110, nw=100, ne=120): avg 107, g = 5 + 10 = 15, corr = 20 >> 3 = 2 (limit 7) -> 109; 115 - 109 = 6
    # (1,2): w=115 n=120 nw=110 ne=nw (last column) -> avg 117, corr 0 -> 117; 125 - 117 = 8
    d = [int(v) - thr for v in _expand_rle(sym)[1:]]
    assert d == [100, 10, 10, 5, 6, 8, 5, 6, 8]


def _expand_rle(sym):
    """RleDecompressU16.DecodeNext2 over a spatial symbol stream (rledecompressu16.go:59-97), in Python for tiny inputs."""
    sym = [int(v) for v in sym]
    depth = sym[0].bit_length()
    mid = (1 << (depth - 1)) - 1
    out, i = [], 1
    while i < len(sym):
        c = sym[i]
        if c <= mid:
            out += [sym[i + 1]] * c
            i += 2
        else:
            n = c - mid
            out += sym[i + 1:i + 1 + n]
            i += 1 + n
    return out


@pytest.mark.parametrize("name,w,h,avg,grad,pics4,pica4,ngrad", [
    # docs/adaptive-compression.md:32-33,74-75 (published ratios; the MR rows of that table are 0.2 % below what the
    # reference encoder's byte counts give here for BOTH predictors, the CT rows match to the last digit)
    ("MR_256_256", 256, 256, 2.348, 2.374, 2.284, 2.309, 3),
    ("CT_512_512", 512, 512, 2.237, 2.182, 2.145, 2.112, 0),
])
def test_grad_and_pica_published_ratios(oracle, name, w, h, avg, grad, pics4, pica4, ngrad):
    px = np.fromfile(os.path.join(GOLDEN, f"{name}_image.bin"), dtype="<u2")
    mx, raw = int(px.max()), w * h * 2
    g = oracle.compress_single_frame_grad(px, w, h, mx)
    a = oracle.compress_single_frame(px, w, h, mx, 2)
    assert np.array_equal(oracle.decompress_single_frame_grad(g, w, h), px)
    assert abs(raw / len(a) - avg) < 0.006 and abs(raw / len(g) - grad) < 0.006
    pica = oracle.pica_compress(px, w, h, mx, 4)
    pics = oracle.pics_compress(px, w, h, mx, 4, 2)
    assert abs(raw / len(pics) - pics4) < 0.0006 and abs(raw / len(pica) - pica4) < 0.0006
    n = int.from_bytes(pica[12:16], "little")
    flags = [int.from_bytes(pica[28 + 16 * i:32 + 16 * i], "little") for i in range(n)]
    assert n == 4 and sum(flags) == ngrad          # "Grad strips" column of the same table
    got, gw, gh = oracle.pica_decompress(pica)
    assert (gw, gh) == (w, h) and np.array_equal(got, px)


def test_pica_layout_and_boundaries(oracle, synth):
    img = synth.smooth_image(9, 320, 200, noisy_from=120)
    px, mx = img.ravel(), int(img.max())
    blob = oracle.pica_compress(px, 320, 200, mx, 3)
    assert blob[:4] == b"PICA" and [int.from_bytes(blob[4 + 4 * i:8 + 4 * i], "little") for i in range(3)] == [320, 200, 3]
    starts = oracle.pica_boundaries(px, 320, 200, 3)
    ent = [[int.from_bytes(blob[16 + 16 * s + 4 * k:20 + 16 * s + 4 * k], "little") for k in range(4)] for s in range(3)]
    assert [e[0] for e in ent] == starts and starts[0] == 0 and starts == sorted(starts)
    off = 0
    for s, (y0, o, ln, fl) in enumerate(ent):     # offsets are running sums, every strip decodes on its own
        assert o == off and fl in (0, 1)
        off += ln
        y1 = ent[s + 1][0] if s + 1 < 3 else 200
        frame = blob[16 + 48 + o:16 + 48 + o + ln]
        dec = oracle.decompress_single_frame_grad(frame, 320, y1 - y0) if fl else oracle.decompress_single_frame(frame, 320, y1 - y0)
        assert np.array_equal(dec, px[y0 * 320:y1 * 320])
        # the kept frame is the smaller of the two (the gradient one on a tie), parallelstripsadaptive.go:98-107
        a = oracle.compress_single_frame(px[y0 * 320:y1 * 320], 320, y1 - y0, mx, 2)
        g = oracle.compress_single_frame_grad(px[y0 * 320:y1 * 320], 320, y1 - y0, mx)
        assert frame == (g if len(g) <= len(a) else a) and fl == int(len(g) <= len(a))
    assert len(blob) == 16 + 48 + off
    # clamps and degenerate partitions (parallelstripsadaptive.go:62-70,215-224,243-251)
    assert oracle.pica_boundaries(px, 320, 200, 1) == [0]
    assert oracle.pica_boundaries(px, 320, 200, 500) == list(range(200))
    flat = np.full(64 * 40, 777, np.uint16)
    assert oracle.pica_boundaries(flat, 64, 40, 4) == [0, 10, 20, 30]   # total cost 0 -> equal heights
    got, w, h = oracle.pica_decompress(blob)
    assert (w, h) == (320, 200) and np.array_equal(got, px)
    for bad in (b"PICS" + blob[4:], blob[:15], blob[:12] + (0).to_bytes(4, "little") + blob[16:], blob[:40]):
        with pytest.raises(Exception):
            oracle.pica_decompress(bad)


@pytest.mark.parametrize("name,w,h,wav5,pics", [
    # README.md:268-269 ("Wavelet" column, WaveletV2SIMD at 5 levels as in waveletu16_test.go:391) and
    # docs/parallel-strips.md:98-99 (1 / 2 / 4 / 8 strips, two-state strips)
    ("MR_256_256", 256, 256, 2.38, (2.35, 2.33, 2.28, 2.21)),
    ("CT_512_512", 512, 512, 1.67, (2.24, 2.21, 2.15, 1.96)),
])
def test_published_ratios_wavelet_and_strip_counts(oracle, name, w, h, wav5, pics):
    """The WaveletV2 stream has no golden vector and the reference's C twin does not cover it: its size on the two images
    the repository ships is pinned to the three digits the reference publishes (a coefficient-order, escape or RLE
    mistake moves the ratio in the first or second digit).  The same for PICS at every published strip count."""
    px = np.fromfile(os.path.join(GOLDEN, f"{name}_image.bin"), dtype="<u2")
    mx, raw = int(px.max()), w * h * 2
    blob = oracle.wavelet_v2_compress(px, h, w, mx, 5)
    assert abs(raw / len(blob) - wav5) < 0.005
    got, rows, cols = oracle.wavelet_v2_decompress(blob)
    assert (rows, cols) == (h, w) and np.array_equal(got, px)
    for n, want in zip((1, 2, 4, 8), pics):
        assert abs(raw / len(oracle.pics_compress(px, w, h, mx, n, 2)) - want) < 0.006   # published to two decimals, some rounded twice (2.1446 -> 2.145 -> 2.15)


@pytest.mark.parametrize("with_rle", [0, 1])
def test_wavelet_v1_layouts_roundtrip(oracle, with_rle):
    """WaveletFSECompressU16 / WaveletRLEFSECompressU16 (waveletfsecompressu16.go:71-189, 551-669): header layout (11 / 15
    bytes, the RLE variant stores the coefficient-stream length), level clamp to 4 and pixel-exact round trips on the
    reference's own images (TestWaveletFSECompress / TestWaveletRLEFSECompress shapes, waveletu16_test.go)."""
    for name, w, h in (("MR_256_256", 256, 256), ("CT_512_512", 512, 512)):
        px = np.fromfile(os.path.join(GOLDEN, f"{name}_image.bin"), dtype="<u2")
        mx = int(px.max())
        for levels in (1, 3, 9):
            blob = oracle.wavelet_v1_compress(px, h, w, mx, levels, with_rle)
            assert int.from_bytes(blob[0:4], "little") == h and int.from_bytes(blob[4:8], "little") == w
            assert int.from_bytes(blob[8:10], "little") == mx and blob[10] == min(levels, 4)
            hdr = 15 if with_rle else 11
            assert blob[hdr] == 0xFF and blob[hdr + 1] == 0x04          # FSECompressU16FourState magic
            if with_rle:
                assert int.from_bytes(blob[11:15], "little") >= w * h   # one word per coefficient, three per escape
            got, r, c = oracle.wavelet_v1_decompress(blob, with_rle)
            assert (r, c) == (h, w) and np.array_equal(got, px)


@pytest.mark.parametrize("name,w,h,frame,pics,wav_ratio,wav_mib", [
    # results/20260518-112009/01-all-codecs-decompress.txt:11-13,22-26 (MR) and :28-30,39-43 (CT): `ratio` of MIC-Go (the 2-state
    # CompressSingleFrame), MIC-4state, MIC-8state and PICS-2 / -4 / -8, all measured with the GO encoder, four significant
    # digits; 06-wavelet-simd.txt:5-6: WaveletV2SIMD `ratio` and `comp` (MiB) of the same two images
    ("MR_256_256", 256, 256, (2.353, 2.353, 2.352), (2.325, 2.284, 2.214), 2.381, 0.05250),
    ("CT_512_512", 512, 512, (2.237, 2.237, 2.237), (2.214, 2.145, 1.962), 1.669, 0.2995),
])
def test_published_four_digit_ratios_of_the_go_encoder(oracle, name, w, h, frame, pics, wav_ratio, wav_mib):
    """The reference's committed benchmark logs print the compressed size of its two shipped images to four significant
    digits for streams its GO encoder produced (not the C twin): a pin of +-0.02 % (about 12 bytes on the MR image) on the
    oracle's 2/4/8-state frames, its PICS containers at 2 / 4 / 8 strips and its WaveletV2 stream.  Not byte identity -- the
    parity status stays "unpinned" for the formats the C twin does not cover -- but a count that is off by one run, one
    escape or one table entry on a 55 KB stream moves the fourth digit."""
    px = np.fromfile(os.path.join(GOLDEN, f"{name}_image.bin"), dtype="<u2")
    mx, raw = int(px.max()), w * h * 2
    for n, want in zip((2, 4, 8), frame):
        assert abs(raw / len(oracle.compress_single_frame(px, w, h, mx, n)) - want) < 0.00055, ("frame", n)
    for n, want in zip((2, 4, 8), pics):
        assert abs(raw / len(oracle.pics_compress(px, w, h, mx, n, 2)) - want) < 0.00055, ("pics", n)
    blob = oracle.wavelet_v2_compress(px, h, w, mx, 5)
    assert abs(raw / len(blob) - wav_ratio) < 0.00055
    assert abs(len(blob) / 2 ** 20 - wav_mib) < 0.55 * 10 ** (np.floor(np.log10(wav_mib)) - 3)   # half a unit of the fourth digit
