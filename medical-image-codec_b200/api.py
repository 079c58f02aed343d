"""ctypes binding of libmicgpu.so, named after the reference's Go API.

Go API mirrored here (reference file:line):
  DecompressParallelStrips   parallelstrips.go:270
  DecompressSingleFrame      multiframecompress.go:97
  DecompressMultiFrame       multiframecompress.go:227
  DecompressFrame            multiframecompress.go:266
Errors surface as MicGpuError carrying the C ABI's negative error class, the
way ojph/mic_c.go:36-38 turns rc != 0 into a Go error.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmicgpu.so")

E_HEADER, E_NCOUNT, E_ALLOC, E_DTABLE, E_BITSTREAM, E_RLE, E_SIZE, E_UNSUPPORTED, E_CUDA = -1, -2, -3, -4, -6, -8, -9, -10, -20
E_INCOMPRESSIBLE, E_USE_RLE, E_INTERNAL = -11, -12, -13
KIND_SPATIAL, KIND_RLE = 0, 1


class MicGpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"micgpu rc={code}: {msg}")
        self.code = code


def _load() -> C.CDLL:
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(micgpu has no CPU fallback)"
        )
    return C.CDLL(_LIB_PATH)


lib = _load()

_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_ip = C.POINTER(C.c_int)

lib.micgpu_last_error.restype = C.c_char_p
lib.micgpu_device_count.restype = C.c_int
lib.micgpu_host_alloc.restype = C.c_void_p
lib.micgpu_host_alloc.argtypes = [C.c_size_t]
lib.micgpu_host_free.argtypes = [C.c_void_p]
lib.micgpu_decoder_create.restype = C.c_void_p
lib.micgpu_decoder_create.argtypes = [C.c_int]
lib.micgpu_decoder_destroy.argtypes = [C.c_void_p]
lib.micgpu_decoder_begin.argtypes = [C.c_void_p]
lib.micgpu_decoder_add_unit.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_int, C.c_uint32, C.c_uint32, C.c_uint64]
lib.micgpu_decoder_add_pics.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64, _ip, _ip]
lib.micgpu_decoder_add_mic2.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64, _ip, _ip, _ip, _ip]
lib.micgpu_decoder_add_mic2_range.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_int, C.c_int, _ip, _ip, _ip, _ip]
lib.micgpu_temporal_add_carry.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
lib.micgpu_temporal_add_carry_peers.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_uint64, C.c_int, C.c_void_p]
lib.micgpu_device_alloc.argtypes = [C.c_size_t]
lib.micgpu_device_alloc.restype = C.c_void_p
lib.micgpu_device_free.argtypes = [C.c_void_p]
lib.micgpu_device_free.restype = None
lib.micgpu_ipc_export.argtypes = [C.c_void_p, C.c_void_p]
lib.micgpu_ipc_open.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
lib.micgpu_ipc_close.argtypes = [C.c_void_p]
lib.micgpu_decoder_commit.argtypes = [C.c_void_p]
lib.micgpu_decoder_unit_count.argtypes = [C.c_void_p]
lib.micgpu_decoder_run_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
lib.micgpu_decoder_unit_status.argtypes = [C.c_void_p, _ip, C.c_int, C.c_void_p]
lib.micgpu_decoder_last_launches.argtypes = [C.c_void_p]
lib.micgpu_decoder_set_profiling.argtypes = [C.c_void_p, C.c_int]
lib.micgpu_decoder_kernel_times.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_float), C.c_int]
lib.micgpu_decoder_run_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
lib.micgpu_pics_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, _ip, _ip]
lib.micgpu_pics_decompress_batch.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), _ip]
lib.micgpu_decompress_single_frame.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int]
lib.micgpu_wsi_open.restype = C.c_void_p
lib.micgpu_wsi_open.argtypes = [C.c_int, C.c_void_p, C.c_size_t]
lib.micgpu_wsi_close.argtypes = [C.c_void_p]
lib.micgpu_wsi_slide_regions.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _ip, _ip, _ip, C.c_void_p, C.c_void_p, _ip, _ip, _ip]
lib.micgpu_wsi_slide_regions_device.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _ip, _ip, _ip, C.c_void_p, C.c_void_p, _ip, _ip, _ip]
lib.micgpu_wsi_slide_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, _ip]
lib.micgpu_decompress_single_frame_grad.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int]
lib.micgpu_pica_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, _ip, _ip]
lib.micgpu_decoder_add_pica.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64, _ip, _ip]
lib.micgpu_mic2_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, _ip, _ip, _ip, _ip]
lib.micgpu_mic2_decompress_frame.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, _ip, _ip]
lib.micgpu_file_kind.argtypes = [C.c_void_p, C.c_size_t]
lib.micgpu_mic1_compress.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint16, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
lib.micgpu_mic1_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, _ip, _ip]
lib.micgpu_micr_compress.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
lib.micgpu_micr_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, _ip, _ip]
lib.micgpu_init.argtypes = [_ip, C.c_int]
lib.micgpu_wsi_plan_tiles.restype = C.c_void_p
lib.micgpu_wsi_plan_tiles.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64]
lib.micgpu_wsi_plan_destroy.argtypes = [C.c_void_p]
lib.micgpu_wsi_plan_destroy.restype = None
_u64p = C.POINTER(C.c_uint64)
lib.micgpu_wsi_plan_info.argtypes = [C.c_void_p, _u64p, _u64p, _u64p, _u64p, _ip]
lib.micgpu_wsi_plan_run_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
lib.micgpu_wsi_plan_status.argtypes = [C.c_void_p, _ip, C.c_int, C.c_void_p]
lib.micgpu_wsi_plan_launches.argtypes = [C.c_void_p]
lib.micgpu_wsi_plan_kernel_times.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_float), C.c_int]
lib.micgpu_wsi_decompress_tile_range.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_void_p, C.c_size_t, _ip]


class WsiInfo(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("tile_w", C.c_int), ("tile_h", C.c_int), ("channels", C.c_int),
                ("bits_per_sample", C.c_int), ("color_transform", C.c_int), ("n_levels", C.c_int), ("total_tiles", C.c_uint64),
                ("level_w", C.c_int * 32), ("level_h", C.c_int * 32), ("tiles_x", C.c_int * 32), ("tiles_y", C.c_int * 32),
                ("first_tile", C.c_int * 32)]


lib.micgpu_wsi_read_header.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(WsiInfo)]
lib.micgpu_wsi_decompress_tile.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, _ip, _ip]
lib.micgpu_wsi_decompress_tiles.argtypes = [C.c_void_p, C.c_size_t, C.c_int, _ip, _ip, _ip, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), _ip, _ip, _ip]
lib.micgpu_wsi_decompress_region.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, _ip, _ip]
lib.micgpu_rgb_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
lib.micgpu_wavelet_v2_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, _ip, _ip]
lib.micgpu_huff_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
lib.micgpu_delta_rle_huff_compress.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint16, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
lib.micgpu_huff_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
lib.micgpu_delta_rle_huff_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int]
lib.micgpu_decoder_add_huff_unit.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_int, C.c_uint32, C.c_uint32, C.c_uint64]
lib.micgpu_wavelet_v1_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, _ip, _ip]
lib.micgpu_wavelet_v2_decompress_batch.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), _ip, _ip, _ip]
_szp = C.POINTER(C.c_size_t)
lib.micgpu_delta_rle_compress.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint16, C.c_void_p, C.c_size_t, _szp]
lib.micgpu_rle_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_uint16, C.c_void_p, C.c_size_t, _szp]
lib.micgpu_compress_single_frame.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint16, C.c_int, C.c_void_p, C.c_size_t, _szp]
lib.micgpu_pics_compress.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint16, C.c_int, C.c_int, C.c_void_p, C.c_size_t, _szp]
lib.micgpu_compress_single_frame_grad.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint16, C.c_void_p, C.c_size_t, _szp]
lib.micgpu_pica_compress.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint16, C.c_int, C.c_void_p, C.c_size_t, _szp]
lib.micgpu_pica_boundaries.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _ip, _ip]
lib.micgpu_pics_compress_batch.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.POINTER(C.c_uint16), C.c_int, C.c_int,
                                           C.POINTER(C.c_void_p), _szp, _szp, _ip]
lib.micgpu_mic2_compress.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint16, C.c_int, C.c_void_p, C.c_size_t, _szp]
lib.micgpu_wavelet_v1_compress.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint16, C.c_int, C.c_int, C.c_void_p, C.c_size_t, _szp]
lib.micgpu_wavelet_v2_compress.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint16, C.c_int, C.c_void_p, C.c_size_t, _szp]
lib.micgpu_wavelet_v2_compress_batch.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.POINTER(C.c_uint16), C.c_int,
                                                 C.POINTER(C.c_void_p), _szp, _szp, _ip]
lib.micgpu_rgb_compress.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, _szp]
lib.micgpu_wsi_compress.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, _szp]
for _n in ("two", "four", "eight"):
    getattr(lib, f"mic_compress_{_n}_state").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, _szp]
for _n in ("two", "four", "eight"):
    for _s in ("", "_simd"):
        getattr(lib, f"mic_decompress_{_n}_state{_s}").argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int]
lib.mic_decompress_parallel.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_int]
lib.mic_decompress_parallel_scalar.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_int]


def temporal_add_carry(d_frames_ptr: int, d_carry_ptr: int, frame_px: int, nframes: int, stream_ptr: int = 0):
    """frames[f] += carry (mod 2^16) on device pointers: the exchange step of a sharded temporal MIC2 stack."""
    _check(lib.micgpu_temporal_add_carry(d_frames_ptr, d_carry_ptr, frame_px, nframes, stream_ptr))


def temporal_add_carry_peers(d_frames_ptr: int, peer_last_ptrs, frame_px: int, nframes: int, stream_ptr: int = 0):
    """Fused exchange: frames[f] += sum of the peers' last frames (device pointers, possibly on peer GPUs)."""
    n = len(peer_last_ptrs)
    arr = (C.c_void_p * max(n, 1))(*peer_last_ptrs)
    _check(lib.micgpu_temporal_add_carry_peers(d_frames_ptr, arr, n, frame_px, nframes, stream_ptr))


def device_alloc(nbytes: int) -> int:
    p = lib.micgpu_device_alloc(nbytes)
    if not p:
        raise MicGpuError(E_ALLOC, last_error())
    return p


def device_free(ptr: int):
    lib.micgpu_device_free(ptr)


def ipc_export(d_ptr: int) -> bytes:
    buf = C.create_string_buffer(64)
    _check(lib.micgpu_ipc_export(d_ptr, buf))
    return buf.raw


def ipc_open(handle: bytes) -> int:
    p = C.c_void_p()
    _check(lib.micgpu_ipc_open(C.create_string_buffer(handle, 64), C.byref(p)))
    return p.value


def ipc_close(d_ptr: int):
    _check(lib.micgpu_ipc_close(d_ptr))


def last_error() -> str:
    return (lib.micgpu_last_error() or b"").decode("utf-8", "replace")


def _check(rc: int) -> None:
    if rc != 0:
        raise MicGpuError(rc, last_error())


def _bytes_view(b) -> np.ndarray:
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b, dtype=np.uint8)
    return np.frombuffer(b, dtype=np.uint8)


def _rd32(a: np.ndarray, off: int) -> int:
    return int(a[off]) | int(a[off + 1]) << 8 | int(a[off + 2]) << 16 | int(a[off + 3]) << 24


# ---- Go-named one-shot calls (host buffers) --------------------------------------
def DecompressParallelStrips(compressed):
    """parallelstrips.go:270 -> (pixels[h*w] uint16, width, height)."""
    a = _bytes_view(compressed)
    if a.size < 20 or bytes(a[:4]) != b"PICS":
        raise MicGpuError(E_HEADER, "parallelstrips: invalid magic")
    w, h = _rd32(a, 4), _rd32(a, 8)
    if w <= 0 or h <= 0 or w * h > (1 << 34):
        raise MicGpuError(E_HEADER, "parallelstrips: invalid dimensions")
    out = np.empty(w * h, np.uint16)
    ow, oh = C.c_int(), C.c_int()
    _check(lib.micgpu_pics_decompress(a.ctypes.data, a.size, out.ctypes.data, out.size, C.byref(ow), C.byref(oh)))
    return out, ow.value, oh.value


def DecompressParallelStripsBatch(blobs):
    """Batch of independent PICS images in one launch sequence -> list of (pixels, w, h)."""
    views = [_bytes_view(b) for b in blobs]
    n = len(views)
    outs, dims = [], []
    for a in views:
        if a.size < 20 or bytes(a[:4]) != b"PICS":
            raise MicGpuError(E_HEADER, "parallelstrips: invalid magic")
        w, h = _rd32(a, 4), _rd32(a, 8)
        dims.append((w, h))
        outs.append(np.empty(w * h, np.uint16))
    bp = (C.c_void_p * n)(*[a.ctypes.data for a in views])
    ln = (C.c_size_t * n)(*[a.size for a in views])
    op = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
    cp = (C.c_size_t * n)(*[o.size for o in outs])
    st = (C.c_int * n)()
    _check(lib.micgpu_pics_decompress_batch(n, bp, ln, op, cp, st))
    return [(o, w, h) for o, (w, h) in zip(outs, dims)]


def DecompressSingleFrame(compressed, width: int, height: int) -> np.ndarray:
    """multiframecompress.go:97."""
    a = _bytes_view(compressed)
    out = np.empty(width * height, np.uint16)
    _check(lib.micgpu_decompress_single_frame(a.ctypes.data, a.size, out.ctypes.data, width, height))
    return out


def DecompressSingleFrameGrad(compressed, width: int, height: int) -> np.ndarray:
    """multiframecompress.go:129 (gradient-adaptive predictor)."""
    a = _bytes_view(compressed)
    out = np.empty(width * height, np.uint16)
    _check(lib.micgpu_decompress_single_frame_grad(a.ctypes.data, a.size, out.ctypes.data, width, height))
    return out


def DecompressParallelStripsAdaptive(compressed):
    """parallelstripsadaptive.go:143 -> (pixels[h*w] uint16, width, height)."""
    a = _bytes_view(compressed)
    if a.size < 16 or bytes(a[:4]) != b"PICA":
        raise MicGpuError(E_HEADER, "pica: invalid magic")
    w, h = _rd32(a, 4), _rd32(a, 8)
    if w <= 0 or h <= 0 or w * h > (1 << 34):
        raise MicGpuError(E_HEADER, "pica: invalid dimensions")
    out = np.empty(w * h, np.uint16)
    ow, oh = C.c_int(), C.c_int()
    _check(lib.micgpu_pica_decompress(a.ctypes.data, a.size, out.ctypes.data, out.size, C.byref(ow), C.byref(oh)))
    return out, ow.value, oh.value


def DecompressMultiFrame(data):
    """multiframecompress.go:227 -> (frames[n,h,w] uint16, header dict)."""
    a = _bytes_view(data)
    if a.size < 20 or bytes(a[:4]) != b"MIC2":
        raise MicGpuError(E_HEADER, "MIC2: invalid magic")
    w, h, n = _rd32(a, 4), _rd32(a, 8), _rd32(a, 12)
    out = np.empty(max(n * w * h, 1), np.uint16)
    ow, oh, on, ot = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    _check(lib.micgpu_mic2_decompress(a.ctypes.data, a.size, out.ctypes.data, out.size, C.byref(ow), C.byref(oh), C.byref(on), C.byref(ot)))
    hdr = {"Width": ow.value, "Height": oh.value, "FrameCount": on.value, "Temporal": bool(ot.value)}
    return out[: n * w * h].reshape(n, h, w), hdr


def DecompressFrame(data, frame_idx: int):
    """multiframecompress.go:266 -> pixels[h,w]."""
    a = _bytes_view(data)
    if a.size < 20 or bytes(a[:4]) != b"MIC2":
        raise MicGpuError(E_HEADER, "MIC2: invalid magic")
    w, h = _rd32(a, 4), _rd32(a, 8)
    out = np.empty(max(w * h, 1), np.uint16)
    ow, oh = C.c_int(), C.c_int()
    _check(lib.micgpu_mic2_decompress_frame(a.ctypes.data, a.size, frame_idx, out.ctypes.data, out.size, C.byref(ow), C.byref(oh)))
    return out[: w * h].reshape(h, w)


def ReadWSIHeader(data) -> dict:
    """wsicompress.go:299 -> WSIHeader as a dict (Levels: list of (w, h, tilesX, tilesY, firstTileIdx))."""
    a = _bytes_view(data)
    info = WsiInfo()
    _check(lib.micgpu_wsi_read_header(a.ctypes.data, a.size, C.byref(info)))
    n = info.n_levels
    return {"Width": info.width, "Height": info.height, "TileWidth": info.tile_w, "TileHeight": info.tile_h,
            "Channels": info.channels, "BitsPerSample": info.bits_per_sample, "ColorTransform": bool(info.color_transform),
            "TotalTiles": info.total_tiles,
            "Levels": [(info.level_w[i], info.level_h[i], info.tiles_x[i], info.tiles_y[i], info.first_tile[i]) for i in range(min(n, 32))]}


def _wsi_bpp(hdr) -> int:
    return hdr["Channels"] * (2 if hdr["BitsPerSample"] == 16 else 1)


def DecompressWSITile(data, level: int, tile_x: int, tile_y: int):
    """wsicompress.go:175 -> (bytes as uint8 array, width, height); edge tiles cropped."""
    a = _bytes_view(data)
    hdr = ReadWSIHeader(a)
    out = np.empty(hdr["TileWidth"] * hdr["TileHeight"] * _wsi_bpp(hdr), np.uint8)
    w, h = C.c_int(), C.c_int()
    _check(lib.micgpu_wsi_decompress_tile(a.ctypes.data, a.size, level, tile_x, tile_y, out.ctypes.data, out.size, C.byref(w), C.byref(h)))
    return out[: w.value * h.value * _wsi_bpp(hdr)], w.value, h.value


def DecompressWSITiles(data, tiles):
    """Batch of (level, tx, ty) tiles of one container in one launch sequence -> list of (bytes, w, h)."""
    a = _bytes_view(data)
    hdr = ReadWSIHeader(a)
    n = len(tiles)
    cap = hdr["TileWidth"] * hdr["TileHeight"] * _wsi_bpp(hdr)
    outs = [np.empty(cap, np.uint8) for _ in range(n)]
    lv = (C.c_int * n)(*[t[0] for t in tiles])
    tx = (C.c_int * n)(*[t[1] for t in tiles])
    ty = (C.c_int * n)(*[t[2] for t in tiles])
    op = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
    cp = (C.c_size_t * n)(*[cap] * n)
    ws, hs, st = (C.c_int * n)(), (C.c_int * n)(), (C.c_int * n)()
    _check(lib.micgpu_wsi_decompress_tiles(a.ctypes.data, a.size, n, lv, tx, ty, op, cp, ws, hs, st))
    bpp = _wsi_bpp(hdr)
    return [(outs[i][: ws[i] * hs[i] * bpp], ws[i], hs[i]) for i in range(n)]


def DecompressWSIRegion(data, level: int, x: int, y: int, w: int, h: int):
    """wsicompress.go:220 -> (bytes as uint8 array, width, height) clamped to the level extent."""
    a = _bytes_view(data)
    hdr = ReadWSIHeader(a)
    out = np.empty(max(w, 1) * max(h, 1) * _wsi_bpp(hdr), np.uint8)
    ow, oh = C.c_int(), C.c_int()
    _check(lib.micgpu_wsi_decompress_region(a.ctypes.data, a.size, level, x, y, w, h, out.ctypes.data, out.size, C.byref(ow), C.byref(oh)))
    return out[: ow.value * oh.value * _wsi_bpp(hdr)], ow.value, oh.value


def Init(devices=None) -> int:
    """micgpu_init: spread the one-call batch entry points over these devices (None = all visible)."""
    if devices is None:
        rc = lib.micgpu_init(None, 0)
    else:
        arr = (C.c_int * len(devices))(*devices)
        rc = lib.micgpu_init(arr, len(devices))
    if rc <= 0:
        raise MicGpuError(rc, last_error())
    return rc


def DecompressWSITileRange(data, first_tile: int, n_tiles: int):
    """Tiles [first, first+n) of the tile table as full (uncropped) tiles -> uint8 array (n_tiles, tile_bytes)."""
    a = _bytes_view(data)
    hdr = ReadWSIHeader(a)
    tile_bytes = hdr["TileWidth"] * hdr["TileHeight"] * _wsi_bpp(hdr)
    out = np.empty(max(n_tiles, 1) * tile_bytes, np.uint8)
    st = (C.c_int * max(n_tiles, 1))()
    _check(lib.micgpu_wsi_decompress_tile_range(a.ctypes.data, a.size, C.c_uint64(first_tile), C.c_uint64(n_tiles), out.ctypes.data, out.size, st))
    return out[: n_tiles * tile_bytes].reshape(n_tiles, tile_bytes)


class WsiSlide:
    """micgpu_wsi_open / micgpu_wsi_slide_regions: a slide resident in device memory, headers parsed once; serves batches
    of rectangles from any pyramid levels (the DecompressWSIRegion call of a viewer, wsicompress.go:220)."""

    def __init__(self, data, device: int = 0):
        a = _bytes_view(data)
        self.p = lib.micgpu_wsi_open(device, a.ctypes.data, a.size)
        if not self.p:
            raise MicGpuError(-1, last_error())
        self.header = ReadWSIHeader(a)
        self.bpp = _wsi_bpp(self.header)

    def close(self):
        if self.p:
            lib.micgpu_wsi_close(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def regions(self, rects, raise_on_error: bool = True):
        """rects: list of (level, x, y, w, h) -> list of (uint8 array, width, height, status)."""
        n = len(rects)
        outs = [np.empty(max(r[3], 1) * max(r[4], 1) * self.bpp, np.uint8) for r in rects]
        col = lambda k: (C.c_int * n)(*[int(r[k]) for r in rects])
        op = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        cp = (C.c_size_t * n)(*[o.size for o in outs])
        ow, oh, st = (C.c_int * n)(), (C.c_int * n)(), (C.c_int * n)()
        rc = lib.micgpu_wsi_slide_regions(self.p, n, col(0), col(1), col(2), col(3), col(4), op, cp, ow, oh, st)
        if rc and raise_on_error:
            raise MicGpuError(rc, last_error())
        return [(outs[i][: ow[i] * oh[i] * self.bpp], ow[i], oh[i], st[i]) for i in range(n)]

    def region(self, level: int, x: int, y: int, w: int, h: int):
        px, ow, oh, _ = self.regions([(level, x, y, w, h)])[0]
        return px, ow, oh

    def stats(self):
        rq, td, ll = C.c_uint64(), C.c_uint64(), C.c_int()
        _check(lib.micgpu_wsi_slide_stats(self.p, C.byref(rq), C.byref(td), C.byref(ll)))
        return {"requests": rq.value, "tiles_decoded": td.value, "last_launches": ll.value}


class WsiPlan:
    """micgpu_wsi_plan_*: plan a tile range once, run it from device-resident bytes (pointers are raw device addresses)."""

    def __init__(self, data, first_tile: int, n_tiles: int, device: int = 0):
        a = _bytes_view(data)
        self.p = lib.micgpu_wsi_plan_tiles(device, a.ctypes.data, a.size, C.c_uint64(first_tile), C.c_uint64(n_tiles))
        if not self.p:
            raise MicGpuError(-1, last_error())
        so, sl, tb, ob, nu = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_int()
        _check(lib.micgpu_wsi_plan_info(self.p, C.byref(so), C.byref(sl), C.byref(tb), C.byref(ob), C.byref(nu)))
        self.span_off, self.span_len, self.tile_bytes, self.out_bytes, self.n_units = so.value, sl.value, tb.value, ob.value, nu.value
        self.n_tiles = n_tiles

    def close(self):
        if self.p:
            lib.micgpu_wsi_plan_destroy(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_device(self, d_span_ptr: int, d_out_ptr: int, stream_ptr: int = 0):
        _check(lib.micgpu_wsi_plan_run_device(self.p, C.c_void_p(d_span_ptr), C.c_void_p(d_out_ptr), C.c_void_p(stream_ptr)))

    def status(self, stream_ptr: int = 0):
        st = (C.c_int * max(self.n_tiles, 1))()
        lib.micgpu_wsi_plan_status(self.p, st, self.n_tiles, C.c_void_p(stream_ptr))
        return list(st)[: self.n_tiles]

    @property
    def last_launches(self) -> int:
        return lib.micgpu_wsi_plan_launches(self.p)

    def set_profiling(self, on: bool):
        _check(lib.micgpu_wsi_plan_kernel_times(self.p, 1 if on else 0, None, 0, None, 0))

    def kernel_times(self):
        names = C.create_string_buffer(4096)
        ms = (C.c_float * 64)()
        n = lib.micgpu_wsi_plan_kernel_times(self.p, -1, names, 4096, ms, 64)
        if n < 0:
            raise MicGpuError(n, last_error())
        nm = names.value.decode().split(";") if n else []
        return [(nm[i], ms[i]) for i in range(n)]


def DecompressRGB(data, width: int, height: int) -> np.ndarray:
    """rgbcompress.go:31 -> interleaved RGB8."""
    a = _bytes_view(data)
    out = np.empty(width * height * 3, np.uint8)
    _check(lib.micgpu_rgb_decompress(a.ctypes.data, a.size, width, height, out.ctypes.data))
    return out


def WaveletV2RLEFSEDecompressU16(compressed):
    """waveletfsecompressu16.go:374 (and the SIMD variant :493) -> (pixels[rows*cols] uint16, rows, cols)."""
    a = _bytes_view(compressed)
    if a.size < 11:
        raise MicGpuError(E_HEADER, "compressed data too short")
    rows, cols = _rd32(a, 0), _rd32(a, 4)
    out = np.empty(max(rows * cols, 1), np.uint16)
    r, c = C.c_int(), C.c_int()
    _check(lib.micgpu_wavelet_v2_decompress(a.ctypes.data, a.size, out.ctypes.data, out.size, C.byref(r), C.byref(c)))
    return out[: rows * cols], r.value, c.value


WaveletV2SIMDRLEFSEDecompressU16 = WaveletV2RLEFSEDecompressU16


def CanHuffmanCompressU16(symbols) -> bytes:
    """canhuffmancompressu16.go:46-81 (Init + Compress) -> one canonical-Huffman stream."""
    a = np.ascontiguousarray(symbols, dtype=np.uint16).ravel()
    cap = 9 + 5 * 65536 + 4 * a.size + 16      # header + symbol list + every symbol escaped (<= 32 bits each)
    out = np.empty(cap, np.uint8)
    got = C.c_size_t()
    _check(lib.micgpu_huff_compress(a.ctypes.data, a.size, out.ctypes.data, out.size, C.byref(got)))
    return out[: got.value].tobytes()


def DeltaRleHuffCompressU16(pixels, width: int, height: int, max_value: int) -> bytes:
    """DeltaRleCompressU16.Compress -> CanHuffmanCompressU16 (fseu16_test.go:881-889)."""
    a = np.ascontiguousarray(pixels, dtype=np.uint16).ravel()
    cap = 9 + 5 * 65536 + 4 * (3 * a.size + 4096) + 16
    out = np.empty(cap, np.uint8)
    got = C.c_size_t()
    _check(lib.micgpu_delta_rle_huff_compress(a.ctypes.data, width, height, max_value, out.ctypes.data, out.size, C.byref(got)))
    return out[: got.value].tobytes()


def CanHuffmanDecompressU16(compressed) -> np.ndarray:
    """canhuffmandecompressu16.go:31-108 (Init + ReadTable + Decompress) -> the symbols of one canonical-Huffman stream."""
    a = _bytes_view(compressed)
    if a.size < 9:
        raise MicGpuError(E_HEADER, "compressed data too short")
    n = (int(a[0]) << 24) | (int(a[1]) << 16) | (int(a[2]) << 8) | int(a[3])
    out = np.empty(max(n, 1), np.uint16)
    got = C.c_size_t()
    _check(lib.micgpu_huff_decompress(a.ctypes.data, a.size, out.ctypes.data, n, C.byref(got)))
    return out[: got.value]


def DeltaRleHuffDecompressU16(compressed, width: int, height: int) -> np.ndarray:
    """deltarlehuffdecompressu16.go:19 (Huffman -> RLE -> inverse avg(top,left) predictor) -> pixels."""
    a = _bytes_view(compressed)
    out = np.empty(width * height, np.uint16)
    _check(lib.micgpu_delta_rle_huff_decompress(a.ctypes.data, a.size, out.ctypes.data, width, height))
    return out


def _wavelet_v1(compressed, with_rle: int):
    a = _bytes_view(compressed)
    if a.size < (15 if with_rle else 11):
        raise MicGpuError(E_HEADER, "compressed data too short")
    rows, cols = _rd32(a, 0), _rd32(a, 4)
    if rows <= 0 or cols <= 0 or rows * cols > (1 << 31):
        raise MicGpuError(E_HEADER, "bad wavelet header")
    out = np.empty(rows * cols, np.uint16)
    r, c = C.c_int(), C.c_int()
    _check(lib.micgpu_wavelet_v1_decompress(a.ctypes.data, a.size, with_rle, out.ctypes.data, out.size, C.byref(r), C.byref(c)))
    return out, r.value, c.value


def WaveletFSEDecompressU16(compressed):
    """waveletfsecompressu16.go:124 (V1 layout: FSE-4 only, raster order) -> (pixels, rows, cols)."""
    return _wavelet_v1(compressed, 0)


def WaveletRLEFSEDecompressU16(compressed):
    """waveletfsecompressu16.go:624 (V1 layout with RLE) -> (pixels, rows, cols)."""
    return _wavelet_v1(compressed, 1)


def WaveletV2DecompressBatch(blobs):
    """n WaveletV2 streams in one launch sequence -> list of (pixels, rows, cols)."""
    views = [_bytes_view(b) for b in blobs]
    n = len(views)
    dims = [(_rd32(a, 0), _rd32(a, 4)) for a in views]
    outs = [np.empty(max(r * c, 1), np.uint16) for r, c in dims]
    bp = (C.c_void_p * n)(*[a.ctypes.data for a in views])
    ln = (C.c_size_t * n)(*[a.size for a in views])
    op = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
    cp = (C.c_size_t * n)(*[o.size for o in outs])
    rs, cs, st = (C.c_int * n)(), (C.c_int * n)(), (C.c_int * n)()
    _check(lib.micgpu_wavelet_v2_decompress_batch(n, bp, ln, op, cp, rs, cs, st))
    return [(o[: r * c], r, c) for o, (r, c) in zip(outs, dims)]


# ---- encode direction ---------------------------------------------------------------------------
def _u16(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint16).ravel()


def DeltaRleCompress(pixels, width: int, height: int, max_value: int) -> np.ndarray:
    """DeltaRleCompressU16.Compress (deltarlecompressu16.go:24) -> RLE symbol stream."""
    a = _u16(pixels)
    out = np.empty(2 * a.size + a.size // 2 + 64, np.uint16)
    n = C.c_size_t()
    _check(lib.micgpu_delta_rle_compress(a.ctypes.data, width, height, max_value, out.ctypes.data, out.size, C.byref(n)))
    return out[: n.value].copy()


def RleCompress(symbols, max_value: int) -> np.ndarray:
    """RleCompressU16.Init(len,1,maxValue) + Compress (rlecompressu16.go:15-93)."""
    a = _u16(symbols)
    out = np.empty(a.size + a.size // 4 + 64, np.uint16)
    n = C.c_size_t()
    _check(lib.micgpu_rle_compress(a.ctypes.data, a.size, max_value, out.ctypes.data, out.size, C.byref(n)))
    return out[: n.value].copy()


def CompressSingleFrame(pixels, width: int, height: int, max_value: int, nstates: int = 2) -> bytes:
    """CompressSingleFrame / 4State / 8State (multiframecompress.go:15,38,67); nstates=1: FSECompressU16 tier only."""
    a = _u16(pixels)
    out = np.empty(4 * a.size + 8192, np.uint8)
    n = C.c_size_t()
    _check(lib.micgpu_compress_single_frame(a.ctypes.data, width, height, max_value, nstates, out.ctypes.data, out.size, C.byref(n)))
    return out[: n.value].tobytes()


def CompressSingleFrame4State(pixels, width, height, max_value) -> bytes:
    return CompressSingleFrame(pixels, width, height, max_value, 4)


def CompressSingleFrame8State(pixels, width, height, max_value) -> bytes:
    return CompressSingleFrame(pixels, width, height, max_value, 8)


def CompressParallelStrips(pixels, width: int, height: int, max_value: int, num_strips: int, nstates: int = 2) -> bytes:
    """CompressParallelStrips / 4State / 8State (parallelstrips.go:55,128,199)."""
    a = _u16(pixels)
    if a.size != width * height:
        raise MicGpuError(E_HEADER, f"parallelstrips: pixel count {a.size} != width*height {width * height}")
    out = np.empty(4 * a.size + 8192 * max(num_strips, 1), np.uint8)
    n = C.c_size_t()
    _check(lib.micgpu_pics_compress(a.ctypes.data, width, height, max_value, num_strips, nstates, out.ctypes.data, out.size, C.byref(n)))
    return out[: n.value].tobytes()


def CompressSingleFrameGrad(pixels, width: int, height: int, max_value: int) -> bytes:
    """CompressSingleFrameGrad (multiframecompress.go:111)."""
    a = _u16(pixels)
    out = np.empty(4 * a.size + 8192, np.uint8)
    n = C.c_size_t()
    _check(lib.micgpu_compress_single_frame_grad(a.ctypes.data, width, height, max_value, out.ctypes.data, out.size, C.byref(n)))
    return out[: n.value].tobytes()


def CompressParallelStripsAdaptive(pixels, width: int, height: int, max_value: int, num_strips: int) -> bytes:
    """CompressParallelStripsAdaptive (parallelstripsadaptive.go:54)."""
    a = _u16(pixels)
    if a.size != width * height:
        raise MicGpuError(E_HEADER, f"pica: pixel count {a.size} != width*height {width * height}")
    out = np.empty(4 * a.size + 8192 * max(min(num_strips, height), 1), np.uint8)
    n = C.c_size_t()
    _check(lib.micgpu_pica_compress(a.ctypes.data, width, height, max_value, num_strips, out.ctypes.data, out.size, C.byref(n)))
    return out[: n.value].tobytes()


def AdaptiveStripBoundaries(pixels, width: int, height: int, num_strips: int):
    """adaptiveStripBoundaries (parallelstripsadaptive.go:214)."""
    a = _u16(pixels)
    starts = (C.c_int * max(1, min(num_strips, height)))()
    n = C.c_int()
    _check(lib.micgpu_pica_boundaries(a.ctypes.data, width, height, num_strips, starts, C.byref(n)))
    return list(starts[: n.value])


def CompressParallelStripsBatch(images, width: int, height: int, max_values, num_strips: int, nstates: int = 2):
    """n images of one geometry in one launch sequence -> list of PICS blobs."""
    arrs = [_u16(im) for im in images]
    n = len(arrs)
    outs = [np.empty(4 * width * height + 8192 * max(num_strips, 1), np.uint8) for _ in range(n)]
    pp = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    mv = (C.c_uint16 * n)(*[int(m) for m in max_values])
    op = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
    cp = (C.c_size_t * n)(*[o.size for o in outs])
    ol = (C.c_size_t * n)()
    st = (C.c_int * n)()
    _check(lib.micgpu_pics_compress_batch(n, pp, width, height, mv, num_strips, nstates, op, cp, ol, st))
    return [outs[i][: ol[i]].tobytes() for i in range(n)]


def CompressMultiFrame(frames, width: int, height: int, max_value: int, temporal: bool) -> bytes:
    """multiframecompress.go:179 (frames: array [n, h, w] or list of flat frames)."""
    a = np.ascontiguousarray(np.asarray(frames), dtype=np.uint16).ravel()
    n = a.size // (width * height)
    out = np.empty(4 * a.size + 8192 * max(n, 1), np.uint8)
    ol = C.c_size_t()
    _check(lib.micgpu_mic2_compress(a.ctypes.data, width, height, n, max_value, int(bool(temporal)), out.ctypes.data, out.size, C.byref(ol)))
    return out[: ol.value].tobytes()


def WaveletV2RLEFSECompressU16(pixels, rows: int, cols: int, max_value: int, levels: int) -> bytes:
    """waveletfsecompressu16.go:303 (and the SIMD variant :427)."""
    a = _u16(pixels)
    out = np.empty(8 * a.size + 8192, np.uint8)
    ol = C.c_size_t()
    _check(lib.micgpu_wavelet_v2_compress(a.ctypes.data, rows, cols, max_value, levels, out.ctypes.data, out.size, C.byref(ol)))
    return out[: ol.value].tobytes()


WaveletV2SIMDRLEFSECompressU16 = WaveletV2RLEFSECompressU16


def _wavelet_v1_compress(pixels, rows, cols, max_value, levels, with_rle) -> bytes:
    a = _u16(pixels)
    out = np.empty(8 * a.size + 8192, np.uint8)
    ol = C.c_size_t()
    _check(lib.micgpu_wavelet_v1_compress(a.ctypes.data, rows, cols, max_value, levels, with_rle, out.ctypes.data, out.size, C.byref(ol)))
    return out[: ol.value].tobytes()


def WaveletFSECompressU16(pixels, rows: int, cols: int, max_value: int, levels: int) -> bytes:
    """waveletfsecompressu16.go:71 (V1 layout: interleaved lifting, raster order, 4-state FSE)."""
    return _wavelet_v1_compress(pixels, rows, cols, max_value, levels, 0)


def WaveletRLEFSECompressU16(pixels, rows: int, cols: int, max_value: int, levels: int) -> bytes:
    """waveletfsecompressu16.go:551 (V1 layout behind the RLE layer)."""
    return _wavelet_v1_compress(pixels, rows, cols, max_value, levels, 1)


def WaveletV2CompressBatch(images, rows: int, cols: int, max_values, levels: int):
    arrs = [_u16(im) for im in images]
    n = len(arrs)
    outs = [np.empty(8 * rows * cols + 8192, np.uint8) for _ in range(n)]
    pp = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    mv = (C.c_uint16 * n)(*[int(m) for m in max_values])
    op = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
    cp = (C.c_size_t * n)(*[o.size for o in outs])
    ol = (C.c_size_t * n)()
    st = (C.c_int * n)()
    _check(lib.micgpu_wavelet_v2_compress_batch(n, pp, rows, cols, mv, levels, op, cp, ol, st))
    return [outs[i][: ol[i]].tobytes() for i in range(n)]


def CompressRGB(rgb, width: int, height: int) -> bytes:
    """rgbcompress.go:25."""
    a = np.ascontiguousarray(rgb, dtype=np.uint8).ravel()
    out = np.empty(8 * a.size + 8192, np.uint8)
    ol = C.c_size_t()
    _check(lib.micgpu_rgb_compress(a.ctypes.data, width, height, out.ctypes.data, out.size, C.byref(ol)))
    return out[: ol.value].tobytes()


def CompressWSI(pixels, width: int, height: int, channels: int = 3, bits_per_sample: int = 8, tile_width: int = 256, tile_height: int = 256,
                pyramid_levels: int = 0) -> bytes:
    """wsicompress.go:27 with WSIOptions{TileWidth, TileHeight, PyramidLevels}; ColorTransform defaults on for RGB."""
    a = np.ascontiguousarray(pixels, dtype=np.uint8).ravel()
    out = np.empty(8 * a.size + (1 << 20), np.uint8)
    ol = C.c_size_t()
    _check(lib.micgpu_wsi_compress(a.ctypes.data, width, height, channels, bits_per_sample, tile_width, tile_height, pyramid_levels,
                                   out.ctypes.data, out.size, C.byref(ol)))
    return out[: ol.value].tobytes()


# ---- .mic files (cmd/mic-compress / cmd/mic-wasm) ----------------------------------------------
def WriteMIC1(pixels, width: int, height: int, max_value: int, nstates: int = 2) -> bytes:
    """writeMicFile(compressImage[4State](...)) (cmd/mic-compress/main.go:26-105) -> the bytes of the .mic file."""
    a = _u16(pixels)
    out = np.empty(4 * a.size + 8192, np.uint8)
    n = C.c_size_t()
    _check(lib.micgpu_mic1_compress(a.ctypes.data, width, height, max_value, nstates, out.ctypes.data, out.size, C.byref(n)))
    return out[: n.value].tobytes()


def WriteMICR(rgb, width: int, height: int) -> bytes:
    a = np.ascontiguousarray(rgb, dtype=np.uint8).ravel()
    out = np.empty(8 * a.size + 8192, np.uint8)
    n = C.c_size_t()
    _check(lib.micgpu_micr_compress(a.ctypes.data, width, height, out.ctypes.data, out.size, C.byref(n)))
    return out[: n.value].tobytes()


def DecodeMicFile(data):
    """decodeMicFile (cmd/mic-wasm/main.go:52-131): dispatch on the magic.  -> (kind, pixels, width, height) with kind in
    MIC1 / MICR / PICS / MIC2 (all frames) / MIC3 (header dict instead of pixels)."""
    a = _bytes_view(data)
    kind = lib.micgpu_file_kind(a.ctypes.data, a.size)
    if kind == 1:
        if a.size < 20:
            raise MicGpuError(-1, "MIC1 file too small")
        w, h = _rd32(a, 4), _rd32(a, 8)
        out = np.empty(max(w * h, 1), np.uint16)
        ww, hh = C.c_int(), C.c_int()
        _check(lib.micgpu_mic1_decompress(a.ctypes.data, a.size, out.ctypes.data, out.size, C.byref(ww), C.byref(hh)))
        return "MIC1", out[: ww.value * hh.value], ww.value, hh.value
    if kind == 4:
        if a.size < 12:
            raise MicGpuError(-1, "MICR file too small")
        w, h = _rd32(a, 4), _rd32(a, 8)
        out = np.empty(max(w * h * 3, 1), np.uint8)
        ww, hh = C.c_int(), C.c_int()
        _check(lib.micgpu_micr_decompress(a.ctypes.data, a.size, out.ctypes.data, out.size, C.byref(ww), C.byref(hh)))
        return "MICR", out[: ww.value * hh.value * 3], ww.value, hh.value
    if kind == 5:
        px, w, h = DecompressParallelStrips(a)
        return "PICS", px, w, h
    if kind == 2:
        frames, hdr = DecompressMultiFrame(a)
        return "MIC2", frames, hdr["Width"], hdr["Height"]
    if kind == 3:
        hdr = ReadWSIHeader(a)
        return "MIC3", hdr, hdr["Width"], hdr["Height"]
    raise MicGpuError(-1, "invalid .mic magic")


# ---- batch decoder over device-resident buffers -------------------------------------
class Decoder:
    """Plan a batch from host copies of the streams, then run it on device buffers."""

    def __init__(self, device: int = 0):
        self._h = lib.micgpu_decoder_create(device)
        if not self._h:
            raise MicGpuError(E_CUDA, last_error())
        self._keep = []

    def close(self):
        if self._h:
            lib.micgpu_decoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def begin(self):
        self._keep = []
        _check(lib.micgpu_decoder_begin(self._h))

    def add_unit(self, frame, comp_off: int, kind: int, width: int, height: int, out_off: int) -> int:
        a = _bytes_view(frame)
        self._keep.append(a)
        rc = lib.micgpu_decoder_add_unit(self._h, a.ctypes.data, a.size, comp_off, kind, width, height, out_off)
        if rc < 0:
            _check(rc)
        return rc

    def add_huff_unit(self, stream, comp_off: int, kind: int, width: int, height: int, out_off: int) -> int:
        a = _bytes_view(stream)
        self._keep.append(a)
        rc = lib.micgpu_decoder_add_huff_unit(self._h, a.ctypes.data, a.size, comp_off, kind, width, height, out_off)
        if rc < 0:
            _check(rc)
        return rc

    def add_pics(self, blob, comp_off: int, out_off: int):
        a = _bytes_view(blob)
        w, h = C.c_int(), C.c_int()
        _check(lib.micgpu_decoder_add_pics(self._h, a.ctypes.data, a.size, comp_off, out_off, C.byref(w), C.byref(h)))
        return w.value, h.value

    def add_mic2(self, blob, comp_off: int, out_off: int):
        a = _bytes_view(blob)
        w, h, n, t = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _check(lib.micgpu_decoder_add_mic2(self._h, a.ctypes.data, a.size, comp_off, out_off, C.byref(w), C.byref(h), C.byref(n), C.byref(t)))
        return w.value, h.value, n.value, bool(t.value)

    def add_mic2_range(self, blob, comp_off: int, out_off: int, first_frame: int, frame_count: int):
        """Frames [first_frame, first_frame + frame_count) of a MIC2 stack (one rank's shard); see include/micgpu.h."""
        a = _bytes_view(blob)
        w, h, n, t = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _check(lib.micgpu_decoder_add_mic2_range(self._h, a.ctypes.data, a.size, comp_off, out_off, first_frame, frame_count,
                                                 C.byref(w), C.byref(h), C.byref(n), C.byref(t)))
        return w.value, h.value, n.value, bool(t.value)

    def commit(self):
        _check(lib.micgpu_decoder_commit(self._h))

    @property
    def unit_count(self) -> int:
        return lib.micgpu_decoder_unit_count(self._h)

    @property
    def last_launches(self) -> int:
        return lib.micgpu_decoder_last_launches(self._h)

    def run_device(self, d_comp_ptr: int, comp_bytes: int, d_out_ptr: int, out_elems: int, stream_ptr: int = 0):
        _check(lib.micgpu_decoder_run_device(self._h, d_comp_ptr, comp_bytes, d_out_ptr, out_elems, stream_ptr))

    def unit_status(self, stream_ptr: int = 0, raise_on_error: bool = True):
        n = self.unit_count
        st = (C.c_int * max(n, 1))()
        rc = lib.micgpu_decoder_unit_status(self._h, st, n, stream_ptr)
        if rc and raise_on_error:
            _check(rc)
        return list(st)[:n]

    def set_profiling(self, on: bool):
        _check(lib.micgpu_decoder_set_profiling(self._h, int(on)))

    def kernel_times(self):
        """[(kernel name, ms)] of the last run (requires set_profiling(True)); waits for the run."""
        names = C.create_string_buffer(4096)
        ms = (C.c_float * 64)()
        n = lib.micgpu_decoder_kernel_times(self._h, names, 4096, ms, 64)
        if n < 0:
            _check(n)
        nm = names.value.decode().split(";") if n else []
        return [(nm[i], float(ms[i])) for i in range(n)]

    def run_host(self, comp: np.ndarray, out: np.ndarray):
        _check(lib.micgpu_decoder_run_host(self._h, comp.ctypes.data, comp.size, out.ctypes.data, out.size))
