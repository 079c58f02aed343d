"""GPU parity: CUDA decode path (through the C ABI) vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _frame(oracle, img, w, h, coder):
    """One FSE frame of a spatial unit with the requested entropy coder."""
    mx = int(img.max())
    if coder in (1, 2, 4, 8):
        return oracle.compress_single_frame(img, w, h, mx, coder)
    sym = oracle.delta_rle_compress(img, w, h, mx)
    return oracle.fse_compress(sym, coder)


@pytest.mark.parametrize("coder", [1, 2, 4, 8, 108])
def test_single_frame_mr(mic, oracle, coder):
    # 256x256 12-bit MR: tableLog 13 (the reference's own round-trip image, fseu16_test.go:28-53)
    rng = np.random.default_rng(5)
    w, h = 256, 256
    y, x = np.mgrid[0:h, 0:w]
    img = ((np.sin(x / 17.0) + np.cos(y / 23.0) + 2) * 700 + rng.integers(0, 40, (h, w))).astype(np.uint16).ravel()
    blob = _frame(oracle, img, w, h, coder)
    got = mic.DecompressSingleFrame(blob, w, h)
    assert np.array_equal(got, oracle.decompress_single_frame(blob, w, h))
    assert np.array_equal(got, img)


@pytest.mark.parametrize("coder", [1, 2, 8])
def test_single_frame_16bit_tablelog16(mic, oracle, coder):
    # full 16-bit range forces tableLog 16: the decode table stays in L2 (mode 2)
    rng = np.random.default_rng(7)
    # (needs >= 2^18 symbols: optimalTableLog caps tableLog at highBits(n-1)-2, fsecompressu16.go:483-486)
    w, h = 530, 520
    img = (np.cumsum(rng.integers(-300, 301, w * h)) % 65536).astype(np.uint16)
    img[::97] = 65535
    blob = _frame(oracle, img, w, h, coder)
    sym = oracle.delta_rle_compress(img, w, h, 65535)
    assert oracle.fse_table_info(sym)[0] == 16
    got = mic.DecompressSingleFrame(blob, w, h)
    assert np.array_equal(got, img)


@pytest.mark.parametrize("w,h", [(1, 1), (2, 2), (1, 40), (40, 1), (3, 3), (7, 5), (31, 33), (33, 31), (64, 64), (257, 129), (1000, 37)])
@pytest.mark.parametrize("coder", [1, 2, 4, 8])
def test_ragged_sizes(mic, oracle, w, h, coder):
    # Single-row, single-column and tiny units (the reference pins one-row strips, parallelstrips_test.go:121).  The
    # encoder only accepts such short streams when the ncount header is short, i.e. for a small alphabet: 5-bit content.
    # Below ~8 symbols the ladder of multiframecompress.go:15-93 ends on a lower tier than asked; the decoder must
    # follow whatever magic the frame carries.
    i = np.arange(w * h)
    img = (20 + (i % 3) + (i // max(w, 1)) % 2).astype(np.uint16)
    blob = oracle.compress_single_frame(img, w, h, 31, coder)
    assert np.array_equal(oracle.decompress_single_frame(blob, w, h), img)
    got = mic.DecompressSingleFrame(blob, w, h)
    assert np.array_equal(got, img)


def test_pics_more_strips_than_rows(mic, oracle, synth):
    # TestParallelStripsSingleRowImage (parallelstrips_test.go:118-144): numStrips > height clamps to one-row strips
    import os
    from conftest import GOLDEN

    w, rows = 256, 2    # the first two rows of the reference's MR image, as in the Go test
    img = np.fromfile(os.path.join(GOLDEN, "MR_256_256_image.bin"), dtype="<u2")[: w * rows].copy()
    for nstates in (2, 8):
        blob = oracle.pics_compress(img, w, rows, int(img.max()), 256, nstates)
        assert int.from_bytes(blob[12:16], "little") == 2 and int.from_bytes(blob[16:20], "little") == 1
        got, ow, oh = mic.DecompressParallelStrips(blob)
        assert (ow, oh) == (w, rows) and np.array_equal(got, img)


@pytest.mark.parametrize("w", [296, 297, 298, 299, 300, 301, 302, 303])
@pytest.mark.parametrize("shift", [0, 3, 5])
def test_row_phase_all_residues(mic, oracle, w, shift):
    # every W mod 8 (row-to-row phase step of the wavefront) x several output alignments, no zero borders
    h = 70
    rng = np.random.default_rng(w * 10 + shift)
    y, x = np.mgrid[0:h, 0:w]
    img = (1500 + 3 * x + 5 * y + rng.integers(0, 30, (h, w))).astype(np.uint16)
    blob = np.frombuffer(oracle.compress_single_frame(img.ravel(), w, h, int(img.max()), 8), np.uint8)
    dec = mic.Decoder(0)
    dec.begin()
    dec.add_unit(blob, 0, 0, w, h, shift)
    dec.commit()
    comp = np.zeros(blob.size + 256, np.uint8)
    comp[: blob.size] = blob
    out = np.zeros(w * h + shift, np.uint16)
    dec.run_host(comp, out)
    assert np.array_equal(out[shift:], img.ravel())
    dec.close()


def test_escapes_and_runs(mic, oracle):
    # sharp edges (escape + literal), literals equal to the delimiter, long constant runs
    w, h = 500, 120
    img = np.zeros((h, w), np.uint16)
    img[:, 100:200] = 4095          # literal == delimiter (depth 12)
    img[10:50, 250:400] = 2000
    img[60:, ::2] = 4095
    img[60:, 1::2] = 0
    img[100:, :] = 777
    img = img.ravel()
    for coder in (2, 4, 8):
        blob = _frame(oracle, img, w, h, coder)
        got = mic.DecompressSingleFrame(blob, w, h)
        assert np.array_equal(got, img), coder


@pytest.mark.parametrize("nstates", [2, 4, 8])
@pytest.mark.parametrize("strips", [1, 3, 8])
def test_pics(mic, oracle, synth, nstates, strips):
    w, h = 611, 403
    img = synth.xr_image(11, w, h).ravel()
    blob = oracle.pics_compress(img, w, h, int(img.max()), strips, nstates)
    got, ow, oh = mic.DecompressParallelStrips(blob)
    assert (ow, oh) == (w, h)
    assert np.array_equal(got, img)


def test_pics_full_size_strip_table(mic, oracle, synth, reftwin):
    # BASELINE config 2 geometry for one image: 2577x2048, 8 strips of 256 rows, 8-state and 2-state
    w, h = 2577, 2048
    img = synth.xr_image(1, w, h).ravel()
    for nstates in (8, 2):
        blob = oracle.pics_compress(img, w, h, int(img.max()), 8, nstates)
        got, _, _ = mic.DecompressParallelStrips(blob)
        assert np.array_equal(got, img)
    # the reference's own pthread decoder agrees on the 2-state container
    assert np.array_equal(reftwin.decompress_parallel(blob, w, h, 8), got)


def test_pics_batch(mic, oracle, synth):
    w, h = 300, 200
    imgs = [synth.xr_image(20 + i, w, h).ravel() for i in range(5)]
    blobs = [oracle.pics_compress(im, w, h, int(im.max()), 4, 8 if i % 2 else 2) for i, im in enumerate(imgs)]
    res = mic.DecompressParallelStripsBatch(blobs)
    for (px, ow, oh), im in zip(res, imgs):
        assert (ow, oh) == (w, h)
        assert np.array_equal(px, im)


def test_pics_batch_pipelined(mic, oracle, synth):
    # >= 16 images take the chunked path (H2D / kernels / D2H of consecutive chunks overlap on several decoder contexts);
    # one corrupt image in the middle must be reported without disturbing the others
    w, h = 211, 160
    imgs = [synth.xr_image(100 + i, w, h).ravel() for i in range(37)]
    blobs = [oracle.pics_compress(im, w, h, int(im.max()), 4, (2, 4, 8)[i % 3]) for i, im in enumerate(imgs)]
    res = mic.DecompressParallelStripsBatch(blobs)
    for (px, ow, oh), im in zip(res, imgs):
        assert (ow, oh) == (w, h) and np.array_equal(px, im)
    bad = list(blobs)
    bad[20] = bad[20][:200] + bytes(len(bad[20]) - 200)
    with pytest.raises(mic.MicGpuError):
        mic.DecompressParallelStripsBatch(bad)


@pytest.mark.parametrize("temporal", [False, True])
def test_mic2(mic, oracle, synth, temporal):
    # 128x128x5 like TestMultiFrame{Independent,Temporal}Roundtrip (multiframe_test.go:149,192)
    st = synth.tomo_stack(7, 5, 128, 128)
    blob = oracle.mic2_compress(st.ravel(), 128, 128, 1023, temporal)
    frames, hdr = mic.DecompressMultiFrame(blob)
    assert hdr["Temporal"] == temporal and hdr["FrameCount"] == 5
    assert np.array_equal(frames, st)
    for idx in (0, 2, 4):
        assert np.array_equal(mic.DecompressFrame(blob, idx), st[idx])


def test_twin_symbols(mic, oracle, synth):
    import ctypes as C

    w, h = 320, 240
    img = synth.xr_image(4, w, h).ravel()
    for nstates, name in ((2, "two"), (4, "four"), (8, "eight")):
        blob = np.frombuffer(oracle.compress_single_frame(img, w, h, int(img.max()), nstates), np.uint8)
        for suffix in ("", "_simd"):
            out = np.zeros(w * h, np.uint16)
            rc = getattr(mic.lib, f"mic_decompress_{name}_state{suffix}")(blob.ctypes.data, blob.size, out.ctypes.data, w, h)
            assert rc == 0 and np.array_equal(out, img)
    pics = np.frombuffer(oracle.pics_compress(img, w, h, int(img.max()), 8, 4), np.uint8)
    out = np.zeros(w * h, np.uint16)
    assert mic.lib.mic_decompress_parallel(pics.ctypes.data, pics.size, out.ctypes.data, w, h, 8) == 0
    assert np.array_equal(out, img)
    assert mic.lib.mic_decompress_parallel(pics.ctypes.data, pics.size, out.ctypes.data, w + 1, h, 8) == -1


def test_corrupt_streams_report_errors(mic, oracle, synth):
    w, h = 200, 100
    img = synth.xr_image(9, w, h).ravel()
    blob = bytearray(oracle.compress_single_frame(img, w, h, int(img.max()), 8))
    trunc = bytes(blob[: len(blob) // 2])
    with pytest.raises(mic.MicGpuError):
        mic.DecompressSingleFrame(trunc + b"\x00", w, h)       # zero last byte: bitreader.go:33-36
    with pytest.raises(mic.MicGpuError):
        mic.DecompressParallelStrips(b"PICX" + bytes(40))


def test_packed_warp_mixed_units_and_statuses(mic, oracle, synth):
    """Units of different length, state count and health share the lanes of one ANS warp (32/N units per warp):
    healthy images must decode exactly and every damaged one must report its own status (parallelstrips.go:324-328)."""
    import ctypes as C

    shapes = [(97, 33), (300, 200), (64, 64), (211, 160), (1000, 37), (128, 128), (257, 129), (40, 300)]
    imgs, blobs, dims = [], [], []
    for i in range(24):
        w, h = shapes[i % len(shapes)]
        im = synth.xr_image(500 + i, w, h).ravel()
        imgs.append(im); dims.append((w, h))
        blobs.append(bytearray(oracle.pics_compress(im, w, h, int(im.max()), (1, 2, 4, 3)[i % 4], 8 if i % 3 else 4)))
    damaged = {5: "truncate", 11: "flip", 18: "count"}
    for i, how in damaged.items():
        b = blobs[i]
        ns = int.from_bytes(b[12:16], "little")
        hdr = 20 + 8 * ns
        off0, len0 = int.from_bytes(b[20:24], "little"), int.from_bytes(b[24:28], "little")
        f0 = hdr + off0
        if how == "truncate":       # strip 0 loses the tail of its bitstream (length field kept consistent, last byte non-zero)
            b[24:28] = (len0 // 2).to_bytes(4, "little")
            b[f0 + len0 // 2 - 1] |= 1
        elif how == "flip":         # bits flipped inside the ncount header / first payload bytes
            for k in range(8, 24):
                b[f0 + k] ^= 0x5A
        else:                       # symbol count of the frame prefix raised: the decoder must stop at the end of the stream
            c = int.from_bytes(b[f0 + 2:f0 + 6], "little")
            b[f0 + 2:f0 + 6] = (c + 4096).to_bytes(4, "little")
    n = len(blobs)
    views = [np.frombuffer(bytes(b), np.uint8) for b in blobs]
    outs = [np.zeros(w * h, np.uint16) for (w, h) in dims]
    bp = (C.c_void_p * n)(*[v.ctypes.data for v in views]); ln = (C.c_size_t * n)(*[v.size for v in views])
    op = (C.c_void_p * n)(*[o.ctypes.data for o in outs]); cp = (C.c_size_t * n)(*[o.size for o in outs])
    st = (C.c_int * n)()
    rc = mic.lib.micgpu_pics_decompress_batch(n, bp, ln, op, cp, st)
    assert rc != 0                                   # first failing image is reported
    for i in range(n):
        if i in damaged:
            assert st[i] != 0, f"image {i} ({damaged[i]}) was accepted"
        else:
            assert st[i] == 0 and np.array_equal(outs[i], imgs[i]), f"healthy image {i} disturbed"


@pytest.mark.parametrize("cuts", [(0, 3, 7), (0, 1, 4, 7), (0, 7, 7)])
def test_mic2_temporal_sharded_ranges(mic, oracle, synth, cuts):
    """a temporal stack cut into frame ranges (one per rank) decodes range by range -- relative sums for a
    range that starts inside the stack -- and the carry exchange (exclusive scan of each range's last frame, then
    micgpu_temporal_add_carry) reproduces the serial decode bit for bit.  The ranges run one after the other on this GPU."""
    import importlib

    import torch

    shard = importlib.import_module("medical-image-codec_b200.shard")
    w, h, nf = 96, 80, 7
    st = synth.tomo_stack(3, nf, w, h)
    blob = np.frombuffer(oracle.mic2_compress(st.ravel(), w, h, 1023, True), np.uint8)
    fpx = w * h
    d_comp = torch.zeros(blob.size + 256, dtype=torch.uint8, device="cuda")
    d_comp[: blob.size] = torch.from_numpy(blob.copy()).cuda()
    stream = torch.cuda.current_stream().cuda_stream
    outs, lasts = [], []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        d_out = torch.zeros(max(hi - lo, 1) * fpx, dtype=torch.int16, device="cuda")
        if hi > lo:
            dec = mic.Decoder(0)
            dec.begin()
            assert dec.add_mic2_range(blob, 0, 0, lo, hi - lo) == (w, h, nf, True)
            dec.commit()
            dec.run_device(d_comp.data_ptr(), blob.size, d_out.data_ptr(), (hi - lo) * fpx, stream)
            assert not any(dec.unit_status(stream))
            dec.close()
            lasts.append(d_out[(hi - lo - 1) * fpx:(hi - lo) * fpx].clone())
        else:
            lasts.append(torch.zeros(fpx, dtype=torch.int16, device="cuda"))   # an empty range carries nothing
        outs.append((lo, hi, d_out))
    stacked = torch.stack(lasts)
    for r, (lo, hi, d_out) in enumerate(outs):
        carry = shard.mic2_temporal_carry(stacked, r)
        if carry is not None and hi > lo:
            mic.temporal_add_carry(d_out.data_ptr(), carry.contiguous().data_ptr(), fpx, hi - lo, stream)
        torch.cuda.synchronize()
        got = d_out.cpu().numpy().view(np.uint16)[: (hi - lo) * fpx]
        assert np.array_equal(got, st[lo:hi].ravel()), f"range {lo}:{hi}"


def test_temporal_carry_fused_over_peer_pointers(mic, oracle, synth):
    """The fused form of the exchange: one kernel reads the last frames of the earlier ranges (here: other buffers on this
    GPU, on a multi-GPU node: peer memory opened through CUDA IPC) and finishes the local frames.  Also checks that a
    library-allocated buffer exports an IPC handle."""
    import torch

    w, h, nf = 96, 80, 7        # frame_px % 8 == 0: vector path; the second stack exercises the scalar path
    for (ww, hh) in ((w, h), (97, 41)):
        st = synth.tomo_stack(5, nf, ww, hh)
        fpx = ww * hh
        full = st.reshape(nf, fpx).astype(np.uint16)
        cuts = (0, 2, 3, 7)
        # relative sums of every range, as micgpu_decoder_add_mic2_range + run_device leave them
        rel = [full[lo:hi].copy() if lo == 0 else (full[lo:hi] - full[lo - 1]).astype(np.uint16) for lo, hi in zip(cuts[:-1], cuts[1:])]
        lasts = []
        for r in rel:
            p = mic.device_alloc(fpx * 2)
            t = torch.from_numpy(r[-1].astype(np.int16)).cuda()
            lasts.append((p, t))      # p: a library allocation (IPC-exportable); t: the frame itself
        stream = torch.cuda.current_stream().cuda_stream
        for k in range(1, len(rel)):
            d = torch.from_numpy(rel[k].astype(np.int16)).cuda().contiguous()
            mic.temporal_add_carry_peers(d.data_ptr(), [lasts[q][1].data_ptr() for q in range(k)], fpx, rel[k].shape[0], stream)
            torch.cuda.synchronize()
            lo, hi = cuts[k], cuts[k + 1]
            assert np.array_equal(d.cpu().numpy().view(np.uint16), full[lo:hi]), (ww, hh, k)
        handle = mic.ipc_export(lasts[0][0])
        assert len(handle) == 64 and any(handle)
        for p, _ in lasts:
            mic.device_free(p)


@pytest.mark.parametrize("w,h", [(300, 40), (256, 64), (2577, 9)])
def test_predictor_uint16_wrap_follows_the_reference(mic, oracle, w, h):
    """uint16(pred + diff) wraps in the reference (deltarlecompressu16.go:123).  No encoder emits such a stream, but a
    decoder has to follow it: symbols are tampered after Delta+RLE so that pixels overflow / underflow, then FSE-coded.
    The row-scan predictor kernel detects the wrap and hands the unit to the pixel-by-pixel wavefront kernel."""
    rng = np.random.default_rng(w + h)
    img = (65000 + rng.integers(0, 400, w * h)).astype(np.uint16)
    img[::7] = rng.integers(0, 300, img[::7].size).astype(np.uint16)      # near both ends of the range
    sym = oracle.delta_rle_compress(img, w, h, 65535).copy()
    depth = 16
    thr = (1 << (depth - 1)) - 1
    # push plain diff symbols (not headers / delimiters / literals) towards +thr or 0 so that pred + diff leaves [0, 65535]
    ref_px = oracle.delta_rle_decompress(sym, w, h)
    assert np.array_equal(ref_px, img)
    cand = np.flatnonzero((sym > thr - 2000) & (sym < thr + 2000))
    cand = cand[(cand > 8)]
    pick = cand[:: max(1, cand.size // 40)]
    sym[pick[0::2]] = 2 * thr - 3
    sym[pick[1::2]] = 5
    try:
        want = oracle.delta_rle_decompress(sym, w, h)
    except Exception:
        pytest.skip("tampered symbol stream no longer parses as RLE")
    for coder in (2, 8):
        blob = oracle.fse_compress(sym, coder)
        got = mic.DecompressSingleFrame(blob, w, h)
        assert np.array_equal(got, want), coder
    assert not np.array_equal(want, img)


@pytest.mark.parametrize("coder", [1, 2, 4, 8])
def test_dense_streams_on_wide_tables(mic, oracle, coder):
    # near-incompressible 15-bit data at tableLog 16: every symbol costs 14-15 bits, so the thread-per-unit kernel consumes
    # its bitstream ring at the highest rate the format allows (three ring checks in a row take > 673 bits: the case that
    # needs the eight-quarter ring of k_ans_serial.cu); RLE-kind unit = the layout of a temporal residual frame
    rng = np.random.default_rng(coder)
    res = rng.integers(0, 24000, 300_000).astype(np.uint16)
    sym = oracle.rle_compress(res, 23999)
    blob = np.frombuffer(oracle.fse_compress(sym, coder), np.uint8)
    hdr = 0 if coder == 1 else 6
    assert (int(blob[hdr]) & 0xF) + 5 == 16 and blob.size * 8 / sym.size > 14.0      # tableLog 16, > 14 bits per symbol
    dec = mic.Decoder(0)
    dec.begin()
    dec.add_unit(blob, 0, 1, res.size, 1, 0)
    dec.commit()
    comp = np.zeros(blob.size + 256, np.uint8)
    comp[: blob.size] = blob
    out = np.zeros(res.size, np.uint16)
    dec.run_host(comp, out)
    assert np.array_equal(out, res)
    dec.close()
