"""In-library multi-device paths (micgpu_init): the one-call batch entry points split their unit
lists by compressed bytes over the device list, one host thread + context per entry.  On a one-GPU box the list names
device 0 several times, which runs the same threading, partition and carry-exchange code (the "peers" are contexts on
the same device); with more GPUs visible the same tests spread over them (gpurun --gpus N)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


@pytest.fixture()
def devlist(mic):
    n = mic.lib.micgpu_device_count()
    devs = list(range(n)) if n >= 2 else [0, 0, 0]
    yield devs
    mic.lib.micgpu_shutdown()       # forgets the device list: later tests run single-device again


def test_init_rejects_bad_lists(mic):
    bad = (C.c_int * 1)(99)
    assert mic.lib.micgpu_init(bad, 1) < 0
    assert mic.Init([0]) == 1
    mic.lib.micgpu_shutdown()


def test_pics_batch_over_devices(mic, oracle, synth, devlist):
    assert mic.Init(devlist) == len(devlist)
    w, h = 211, 160
    imgs = [synth.xr_image(300 + i, w, h).ravel() for i in range(23)]
    blobs = [oracle.pics_compress(im, w, h, int(im.max()), 4, (2, 4, 8)[i % 3]) for i, im in enumerate(imgs)]
    res = mic.DecompressParallelStripsBatch(blobs)
    for (px, ow, oh), im in zip(res, imgs):
        assert (ow, oh) == (w, h) and np.array_equal(px, im)
    # a damaged image in the middle is reported, its neighbours on every device are exact
    bad = list(blobs)
    bad[11] = bad[11][:300] + bytes(len(bad[11]) - 300)
    n = len(bad)
    views = [np.frombuffer(b, np.uint8) for b in bad]
    outs = [np.zeros(w * h, np.uint16) for _ in range(n)]
    bp = (C.c_void_p * n)(*[v.ctypes.data for v in views]); ln = (C.c_size_t * n)(*[v.size for v in views])
    op = (C.c_void_p * n)(*[o.ctypes.data for o in outs]); cp = (C.c_size_t * n)(*[o.size for o in outs])
    st = (C.c_int * n)()
    assert mic.lib.micgpu_pics_decompress_batch(n, bp, ln, op, cp, st) != 0
    for i in range(n):
        if i == 11:
            assert st[i] != 0
        else:
            assert st[i] == 0 and np.array_equal(outs[i], imgs[i]), i


@pytest.mark.parametrize("temporal", [False, True])
def test_mic2_over_devices(mic, oracle, synth, devlist, temporal):
    """DecompressMultiFrame in ONE call over the device list; temporal mode exercises the fused carry kernel that reads
    the parked last frames of the earlier ranges (peer memory when the entries are different GPUs)."""
    assert mic.Init(devlist) == len(devlist)
    nf, rows, cols = 13, 72, 88
    st = synth.tomo_stack(21, nf, rows, cols)
    blob = oracle.mic2_compress(st.ravel(), cols, rows, 1023, temporal)
    frames, hdr = mic.DecompressMultiFrame(blob)
    assert hdr["Temporal"] == temporal and hdr["FrameCount"] == nf
    assert np.array_equal(np.asarray(frames).reshape(st.shape), st)
    # odd frame size: the carry kernel's scalar path
    st2 = synth.tomo_stack(22, 9, 37, 41)
    blob2 = oracle.mic2_compress(st2.ravel(), 41, 37, 1023, temporal)
    frames2, _ = mic.DecompressMultiFrame(blob2)
    assert np.array_equal(np.asarray(frames2).reshape(st2.shape), st2)


def test_wsi_tile_range_over_devices(mic, oracle, synth, devlist):
    assert mic.Init(devlist) == len(devlist)
    W, H = 1100, 790
    rgb = synth.wsi_region(11, 700, 500, W, H, 2500, 2000)
    blob = oracle.wsi_compress(rgb.ravel(), W, H, 3, 8, 256, 256, 0)
    hdr = mic.ReadWSIHeader(blob)
    n = hdr["TotalTiles"]
    got = mic.DecompressWSITileRange(blob, 0, n)
    mic.lib.micgpu_shutdown()
    single = mic.DecompressWSITileRange(blob, 0, n)
    assert np.array_equal(got, single)
    # level 0, tile (1, 1) against the source pixels
    tx_n = hdr["Levels"][0][2]
    t = got[1 * tx_n + 1].reshape(256, 256, 3)
    assert np.array_equal(t, rgb[256:512, 256:512])
