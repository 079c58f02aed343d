"""ctypes front end for the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  The product package never does.

`Oracle`  wraps oracle/libmicoracle.so (C restatement of the Go semantics).
`RefTwin` wraps oracle/_ref/libmicref.so (the reference's own C twin compiled
from /root/reference/ojph/*.c by oracle/Makefile; covers 2/4/8-state
Delta+RLE+FSE encode/decode and PICS decode only).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

ERR_INCOMPRESSIBLE = -1
ERR_USE_RLE = -2
ERR_CORRUPT = -3
ERR_ARG = -4

FSE1, FSE2, FSE4, FSE8, RANS8 = 1, 2, 4, 8, 108


class OracleError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"{what}: oracle rc={code}")
        self.code = code


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when the reference tree is present)."""
    so = os.path.join(_HERE, "libmicoracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("mic_oracle.c", "mic_oracle_huff.c", "mic_oracle.h")]
    need = force or not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(f) for f in srcs)
    need_ref = os.path.isdir("/root/reference/ojph") and not os.path.exists(os.path.join(_HERE, "_ref", "libmicref.so"))
    if need or need_ref:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_i32p = C.POINTER(C.c_int32)


def _ptr(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def _as_u16(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint16)


def _as_u8(a) -> np.ndarray:
    if isinstance(a, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(a), dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


class Oracle:
    def __init__(self):
        build()
        self.lib = C.CDLL(os.path.join(_HERE, "libmicoracle.so"))
        self.lib.orc_free.argtypes = [C.c_void_p]
        self.lib.orc_zigzag.restype = C.c_uint16
        self.lib.orc_zigzag.argtypes = [C.c_int16]
        self.lib.orc_unzigzag.restype = C.c_int16
        self.lib.orc_unzigzag.argtypes = [C.c_uint16]

    # -- helpers ---------------------------------------------------------
    def _take_u8(self, p, n) -> bytes:
        out = C.string_at(p, n.value) if n.value else b""
        self.lib.orc_free(p)
        return out

    def _take_u16(self, p, n) -> np.ndarray:
        out = np.ctypeslib.as_array(p, shape=(max(n.value, 1),))[: n.value].copy() if n.value else np.zeros(0, np.uint16)
        self.lib.orc_free(p)
        return out

    @staticmethod
    def _chk(rc, what):
        if rc != 0:
            raise OracleError(rc, what)

    # -- L1 ----------------------------------------------------------------
    def fse_compress(self, sym, coder: int) -> bytes:
        a = _as_u16(sym)
        p, n = _u8p(), C.c_size_t()
        self._chk(self.lib.orc_fse_compress(_ptr(a, _u16p), C.c_size_t(a.size), coder, C.byref(p), C.byref(n)), "fse_compress")
        return self._take_u8(p, n)

    def fse_decompress(self, blob) -> np.ndarray:
        a = _as_u8(blob)
        p, n = _u16p(), C.c_size_t()
        self._chk(self.lib.orc_fse_decompress_auto(_ptr(a, _u8p), C.c_size_t(a.size), C.byref(p), C.byref(n)), "fse_decompress")
        return self._take_u16(p, n)

    def huff_compress(self, sym) -> bytes:
        a = _as_u16(sym)
        p, n = _u8p(), C.c_size_t()
        self._chk(self.lib.orc_huff_compress(_ptr(a, _u16p), C.c_size_t(a.size), C.byref(p), C.byref(n)), "huff_compress")
        return self._take_u8(p, n)

    def huff_decompress(self, blob) -> np.ndarray:
        a = _as_u8(blob)
        p, n = _u16p(), C.c_size_t()
        self._chk(self.lib.orc_huff_decompress(_ptr(a, _u8p), C.c_size_t(a.size), C.byref(p), C.byref(n)), "huff_decompress")
        return self._take_u16(p, n)

    def delta_rle_huff_compress(self, px, width, height, max_value) -> bytes:
        a = _as_u16(px)
        p, n = _u8p(), C.c_size_t()
        self._chk(self.lib.orc_delta_rle_huff_compress(_ptr(a, _u16p), width, height, C.c_uint16(max_value), C.byref(p), C.byref(n)),
                  "delta_rle_huff_compress")
        return self._take_u8(p, n)

    def delta_rle_huff_decompress(self, blob, width, height) -> np.ndarray:
        a = _as_u8(blob)
        o = np.empty(width * height, dtype=np.uint16)
        self._chk(self.lib.orc_delta_rle_huff_decompress(_ptr(a, _u8p), C.c_size_t(a.size), width, height, _ptr(o, _u16p)),
                  "delta_rle_huff_decompress")
        return o.reshape(height, width)

    def fse_table_info(self, sym):
        a = _as_u16(sym)
        tl, sl = C.c_int(), C.c_int()
        norm = np.zeros(65536, np.int32)
        self._chk(self.lib.orc_fse_table_info(_ptr(a, _u16p), C.c_size_t(a.size), C.byref(tl), C.byref(sl), _ptr(norm, _i32p)), "fse_table_info")
        return tl.value, sl.value, norm[: sl.value].copy()

    # -- L2 ----------------------------------------------------------------
    def rle_compress(self, sym, max_value: int) -> np.ndarray:
        a = _as_u16(sym)
        p, n = _u16p(), C.c_size_t()
        self._chk(self.lib.orc_rle_compress(_ptr(a, _u16p), C.c_size_t(a.size), C.c_uint16(max_value), C.byref(p), C.byref(n)), "rle_compress")
        return self._take_u16(p, n)

    def rle_decompress(self, sym) -> np.ndarray:
        a = _as_u16(sym)
        p, n = _u16p(), C.c_size_t()
        self._chk(self.lib.orc_rle_decompress(_ptr(a, _u16p), C.c_size_t(a.size), C.byref(p), C.byref(n)), "rle_decompress")
        return self._take_u16(p, n)

    def delta_rle_compress(self, px, width, height, max_value) -> np.ndarray:
        a = _as_u16(px)
        assert a.size == width * height
        p, n = _u16p(), C.c_size_t()
        self._chk(self.lib.orc_delta_rle_compress(_ptr(a, _u16p), width, height, C.c_uint16(max_value), C.byref(p), C.byref(n)), "delta_rle_compress")
        return self._take_u16(p, n)

    def delta_rle_decompress(self, sym, width, height) -> np.ndarray:
        a = _as_u16(sym)
        out = np.zeros(width * height, np.uint16)
        self._chk(self.lib.orc_delta_rle_decompress(_ptr(a, _u16p), C.c_size_t(a.size), width, height, _ptr(out, _u16p)), "delta_rle_decompress")
        return out

    def zigzag(self, x: int) -> int:
        return self.lib.orc_zigzag(x)

    def unzigzag(self, x: int) -> int:
        return self.lib.orc_unzigzag(x)

    def temporal_encode(self, cur, prev) -> np.ndarray:
        c, p = _as_u16(cur), _as_u16(prev)
        out = np.zeros(c.size, np.uint16)
        self.lib.orc_temporal_encode(_ptr(c, _u16p), _ptr(p, _u16p), C.c_size_t(c.size), _ptr(out, _u16p))
        return out

    def temporal_decode(self, res, prev) -> np.ndarray:
        r, p = _as_u16(res), _as_u16(prev)
        out = np.zeros(r.size, np.uint16)
        self.lib.orc_temporal_decode(_ptr(r, _u16p), _ptr(p, _u16p), C.c_size_t(r.size), _ptr(out, _u16p))
        return out

    def ycocg_forward(self, rgb):
        a = _as_u8(rgb)
        n = a.size // 3
        y, co, cg = (np.zeros(n, np.uint16) for _ in range(3))
        self.lib.orc_ycocg_forward(_ptr(a, _u8p), C.c_size_t(n), _ptr(y, _u16p), _ptr(co, _u16p), _ptr(cg, _u16p))
        return y, co, cg

    def ycocg_inverse(self, y, co, cg) -> np.ndarray:
        y, co, cg = _as_u16(y), _as_u16(co), _as_u16(cg)
        out = np.zeros(y.size * 3, np.uint8)
        self.lib.orc_ycocg_inverse(_ptr(y, _u16p), _ptr(co, _u16p), _ptr(cg, _u16p), C.c_size_t(y.size), _ptr(out, _u8p))
        return out

    def downsample2x_rgb(self, src, w, h):
        a = _as_u8(src)
        out = np.zeros(max((w // 2) * (h // 2) * 3, 1), np.uint8)
        nw, nh = C.c_int(), C.c_int()
        self.lib.orc_downsample2x_rgb(_ptr(a, _u8p), w, h, _ptr(out, _u8p), C.byref(nw), C.byref(nh))
        return out[: nw.value * nh.value * 3], nw.value, nh.value

    def downsample2x_grey(self, src, w, h):
        a = _as_u16(src)
        out = np.zeros(max((w // 2) * (h // 2), 1), np.uint16)
        nw, nh = C.c_int(), C.c_int()
        self.lib.orc_downsample2x_grey(_ptr(a, _u16p), w, h, _ptr(out, _u16p), C.byref(nw), C.byref(nh))
        return out[: nw.value * nh.value], nw.value, nh.value

    def wt53_forward_1d(self, data, n=None, offset=0, stride=1) -> np.ndarray:
        a = np.ascontiguousarray(data, dtype=np.int32).copy()
        self.lib.orc_wt53_forward_1d(_ptr(a, _i32p), offset, a.size if n is None else n, stride)
        return a

    def wt53_inverse_1d(self, data, n=None, offset=0, stride=1) -> np.ndarray:
        a = np.ascontiguousarray(data, dtype=np.int32).copy()
        self.lib.orc_wt53_inverse_1d(_ptr(a, _i32p), offset, a.size if n is None else n, stride)
        return a

    def wt53_forward_2d(self, data, rows, cols, full_cols=None) -> np.ndarray:
        a = np.ascontiguousarray(data, dtype=np.int32).copy()
        self.lib.orc_wt53_forward_2d(_ptr(a, _i32p), rows, cols, cols if full_cols is None else full_cols)
        return a

    def wt53_inverse_2d(self, data, rows, cols, full_cols=None) -> np.ndarray:
        a = np.ascontiguousarray(data, dtype=np.int32).copy()
        self.lib.orc_wt53_inverse_2d(_ptr(a, _i32p), rows, cols, cols if full_cols is None else full_cols)
        return a

    # -- L3 ----------------------------------------------------------------
    def compress_single_frame(self, px, width, height, max_value, nstates=2) -> bytes:
        a = _as_u16(px)
        assert a.size == width * height
        p, n = _u8p(), C.c_size_t()
        self._chk(self.lib.orc_compress_single_frame(_ptr(a, _u16p), width, height, C.c_uint16(max_value), nstates, C.byref(p), C.byref(n)), "compress_single_frame")
        return self._take_u8(p, n)

    def decompress_single_frame(self, blob, width, height) -> np.ndarray:
        a = _as_u8(blob)
        out = np.zeros(width * height, np.uint16)
        self._chk(self.lib.orc_decompress_single_frame(_ptr(a, _u8p), C.c_size_t(a.size), width, height, _ptr(out, _u16p)), "decompress_single_frame")
        return out

    def compress_residual_frame(self, res, max_value) -> bytes:
        a = _as_u16(res)
        p, n = _u8p(), C.c_size_t()
        self._chk(self.lib.orc_compress_residual_frame(_ptr(a, _u16p), C.c_size_t(a.size), C.c_uint16(max_value), C.byref(p), C.byref(n)), "compress_residual_frame")
        return self._take_u8(p, n)

    def decompress_residual_frame(self, blob) -> np.ndarray:
        a = _as_u8(blob)
        p, n = _u16p(), C.c_size_t()
        self._chk(self.lib.orc_decompress_residual_frame(_ptr(a, _u8p), C.c_size_t(a.size), C.byref(p), C.byref(n)), "decompress_residual_frame")
        return self._take_u16(p, n)

    def wavelet_v2_compress(self, px, rows, cols, max_value, levels) -> bytes:
        a = _as_u16(px)
        assert a.size == rows * cols
        p, n = _u8p(), C.c_size_t()
        self._chk(self.lib.orc_wavelet_v2_compress(_ptr(a, _u16p), rows, cols, C.c_uint16(max_value), levels, C.byref(p), C.byref(n)), "wavelet_v2_compress")
        return self._take_u8(p, n)

    def wavelet_v2_decompress(self, blob):
        a = _as_u8(blob)
        p = _u16p()
        r, c = C.c_int(), C.c_int()
        self._chk(self.lib.orc_wavelet_v2_decompress(_ptr(a, _u8p), C.c_size_t(a.size), C.byref(p), C.byref(r), C.byref(c)), "wavelet_v2_decompress")
        n = C.c_size_t(r.value * c.value)
        return self._take_u16(p, n), r.value, c.value

    def wavelet_v1_compress(self, px, rows, cols, max_value, levels, with_rle) -> bytes:
        a = _as_u16(px)
        p, n = _u8p(), C.c_size_t()
        self._chk(self.lib.orc_wavelet_v1_compress(_ptr(a, _u16p), rows, cols, C.c_uint16(max_value), levels, int(bool(with_rle)), C.byref(p), C.byref(n)), "wavelet_v1_compress")
        return self._take_u8(p, n)

    def wavelet_v1_decompress(self, blob, with_rle):
        a = _as_u8(blob)
        p = _u16p()
        r, c = C.c_int(), C.c_int()
        self._chk(self.lib.orc_wavelet_v1_decompress(_ptr(a, _u8p), C.c_size_t(a.size), int(bool(with_rle)), C.byref(p), C.byref(r), C.byref(c)), "wavelet_v1_decompress")
        return self._take_u16(p, C.c_size_t(r.value * c.value)), r.value, c.value

    # -- L4 ----------------------------------------------------------------
    def pics_compress(self, px, width, height, max_value, num_strips, nstates=2) -> bytes:
        a = _as_u16(px)
        assert a.size == width * height
        p, n = _u8p(), C.c_size_t()
        self._chk(self.lib.orc_pics_compress(_ptr(a, _u16p), width, height, C.c_uint16(max_value), num_strips, nstates, C.byref(p), C.byref(n)), "pics_compress")
        return self._take_u8(p, n)

    def pics_decompress(self, blob):
        a = _as_u8(blob)
        p = _u16p()
        w, h = C.c_int(), C.c_int()
        self._chk(self.lib.orc_pics_decompress(_ptr(a, _u8p), C.c_size_t(a.size), C.byref(p), C.byref(w), C.byref(h)), "pics_decompress")
        return self._take_u16(p, C.c_size_t(w.value * h.value)), w.value, h.value

    def pica_boundaries(self, px, width, height, num_strips):
        a = _as_u16(px)
        starts = (C.c_int * max(1, min(num_strips, height)))()
        n = self.lib.orc_pica_boundaries(_ptr(a, _u16p), width, height, num_strips, starts)
        return list(starts[:n])

    def pica_compress(self, px, width, height, max_value, num_strips) -> bytes:
        a = _as_u16(px)
        assert a.size == width * height
        p, n = _u8p(), C.c_size_t()
        self._chk(self.lib.orc_pica_compress(_ptr(a, _u16p), width, height, C.c_uint16(max_value), num_strips, C.byref(p), C.byref(n)), "pica_compress")
        return self._take_u8(p, n)

    def pica_decompress(self, blob):
        a = _as_u8(blob)
        p = _u16p()
        w, h = C.c_int(), C.c_int()
        self._chk(self.lib.orc_pica_decompress(_ptr(a, _u8p), C.c_size_t(a.size), C.byref(p), C.byref(w), C.byref(h)), "pica_decompress")
        return self._take_u16(p, C.c_size_t(w.value * h.value)), w.value, h.value

    def compress_single_frame_grad(self, px, width, height, max_value) -> bytes:
        a = _as_u16(px)
        p, n = _u8p(), C.c_size_t()
        self._chk(self.lib.orc_compress_single_frame_grad(_ptr(a, _u16p), width, height, C.c_uint16(max_value), C.byref(p), C.byref(n)), "compress_single_frame_grad")
        return self._take_u8(p, n)

    def decompress_single_frame_grad(self, blob, width, height):
        a = _as_u8(blob)
        out = np.empty(width * height, np.uint16)
        self._chk(self.lib.orc_decompress_single_frame_grad(_ptr(a, _u8p), C.c_size_t(a.size), width, height, _ptr(out, _u16p)), "decompress_single_frame_grad")
        return out

    def grad_delta_rle_compress(self, px, width, height, max_value):
        a = _as_u16(px)
        p, n = _u16p(), C.c_size_t()
        self._chk(self.lib.orc_grad_delta_rle_compress(_ptr(a, _u16p), width, height, C.c_uint16(max_value), C.byref(p), C.byref(n)), "grad_delta_rle_compress")
        return self._take_u16(p, n)

    def grad_delta_rle_decompress(self, sym, width, height):
        a = _as_u16(sym)
        out = np.empty(width * height, np.uint16)
        self._chk(self.lib.orc_grad_delta_rle_decompress(_ptr(a, _u16p), C.c_size_t(a.size), width, height, _ptr(out, _u16p)), "grad_delta_rle_decompress")
        return out

    def mic2_compress(self, frames, width, height, max_value, temporal) -> bytes:
        a = _as_u16(frames)
        nframes = a.size // (width * height)
        p, n = _u8p(), C.c_size_t()
        self._chk(self.lib.orc_mic2_compress(_ptr(a, _u16p), width, height, nframes, C.c_uint16(max_value), int(bool(temporal)), C.byref(p), C.byref(n)), "mic2_compress")
        return self._take_u8(p, n)

    def mic2_decompress(self, blob):
        a = _as_u8(blob)
        p = _u16p()
        w, h, nf, t = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._chk(self.lib.orc_mic2_decompress(_ptr(a, _u8p), C.c_size_t(a.size), C.byref(p), C.byref(w), C.byref(h), C.byref(nf), C.byref(t)), "mic2_decompress")
        fr = self._take_u16(p, C.c_size_t(w.value * h.value * nf.value))
        return fr.reshape(nf.value, h.value, w.value), bool(t.value)

    def mic2_decompress_frame(self, blob, idx):
        a = _as_u8(blob)
        p = _u16p()
        w, h = C.c_int(), C.c_int()
        self._chk(self.lib.orc_mic2_decompress_frame(_ptr(a, _u8p), C.c_size_t(a.size), idx, C.byref(p), C.byref(w), C.byref(h)), "mic2_decompress_frame")
        return self._take_u16(p, C.c_size_t(w.value * h.value)).reshape(h.value, w.value)

    def rgb_compress(self, rgb, width, height, color_transform=True) -> bytes:
        a = _as_u8(rgb)
        assert a.size == width * height * 3
        p, n = _u8p(), C.c_size_t()
        self._chk(self.lib.orc_rgb_compress(_ptr(a, _u8p), width, height, int(color_transform), C.byref(p), C.byref(n)), "rgb_compress")
        return self._take_u8(p, n)

    def rgb_decompress(self, blob, width, height, color_transform=True) -> np.ndarray:
        a = _as_u8(blob)
        out = np.zeros(width * height * 3, np.uint8)
        self._chk(self.lib.orc_rgb_decompress(_ptr(a, _u8p), C.c_size_t(a.size), width, height, int(color_transform), _ptr(out, _u8p)), "rgb_decompress")
        return out

    def wsi_plane_compress(self, plane, width, height) -> bytes:
        a = _as_u16(plane)
        p, n = _u8p(), C.c_size_t()
        self._chk(self.lib.orc_wsi_plane_compress(_ptr(a, _u16p), width, height, C.byref(p), C.byref(n)), "wsi_plane_compress")
        return self._take_u8(p, n)

    def wsi_plane_decompress(self, blob, width, height) -> np.ndarray:
        a = _as_u8(blob)
        out = np.zeros(width * height, np.uint16)
        self._chk(self.lib.orc_wsi_plane_decompress(_ptr(a, _u8p), C.c_size_t(a.size), width, height, _ptr(out, _u16p)), "wsi_plane_decompress")
        return out

    def wsi_compress(self, pixels, width, height, channels=3, bps=8, tile_w=256, tile_h=256, pyramid_levels=0) -> bytes:
        a = _as_u8(pixels)
        p, n = _u8p(), C.c_size_t()
        self._chk(self.lib.orc_wsi_compress(_ptr(a, _u8p), width, height, channels, bps, tile_w, tile_h, pyramid_levels, C.byref(p), C.byref(n)), "wsi_compress")
        return self._take_u8(p, n)

    def wsi_decompress_tile(self, blob, level, tx, ty):
        a = _as_u8(blob)
        p, n = _u8p(), C.c_size_t()
        tw, th = C.c_int(), C.c_int()
        self._chk(self.lib.orc_wsi_decompress_tile(_ptr(a, _u8p), C.c_size_t(a.size), level, tx, ty, C.byref(p), C.byref(n), C.byref(tw), C.byref(th)), "wsi_decompress_tile")
        return np.frombuffer(self._take_u8(p, n), np.uint8).copy(), tw.value, th.value

    def wsi_decompress_region(self, blob, level, x, y, w, h):
        a = _as_u8(blob)
        p, n = _u8p(), C.c_size_t()
        ow, oh = C.c_int(), C.c_int()
        self._chk(self.lib.orc_wsi_decompress_region(_ptr(a, _u8p), C.c_size_t(a.size), level, x, y, w, h, C.byref(p), C.byref(n), C.byref(ow), C.byref(oh)), "wsi_decompress_region")
        return np.frombuffer(self._take_u8(p, n), np.uint8).copy(), ow.value, oh.value

    def wsi_header(self, blob):
        a = _as_u8(blob)
        v = [C.c_int() for _ in range(8)]
        info = (C.c_int * (16 * 5))()
        total = C.c_uint64()
        self._chk(self.lib.orc_wsi_header(_ptr(a, _u8p), C.c_size_t(a.size), *[C.byref(x) for x in v], info, 16, C.byref(total)), "wsi_header")
        names = ["width", "height", "tile_w", "tile_h", "channels", "bps", "color_transform", "nlevels"]
        d = {k: x.value for k, x in zip(names, v)}
        d["total_tiles"] = total.value
        d["levels"] = [tuple(info[l * 5 + k] for k in range(5)) for l in range(min(d["nlevels"], 16))]
        return d


class RefTwin:
    """The reference's own C twin (ojph/mic_{compress,decompress}_c.c, mic_parallel.c)."""

    def __init__(self):
        build()
        path = os.path.join(_HERE, "_ref", "libmicref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)

    def compress(self, px, width, height, nstates) -> bytes:
        a = _as_u16(px)
        fn = {2: self.lib.mic_compress_two_state, 4: self.lib.mic_compress_four_state, 8: self.lib.mic_compress_eight_state}[nstates]
        cap = 2 * width * height + 4096
        out = np.zeros(cap, np.uint8)
        n = C.c_size_t()
        rc = fn(_ptr(a, _u16p), width, height, _ptr(out, _u8p), C.c_size_t(cap), C.byref(n))
        if rc != 0:
            raise OracleError(rc, "ref mic_compress")
        return out[: n.value].tobytes()

    def decompress(self, blob, width, height, nstates, simd=True) -> np.ndarray:
        a = _as_u8(blob)
        name = {2: "two", 4: "four", 8: "eight"}[nstates]
        fn = getattr(self.lib, f"mic_decompress_{name}_state" + ("_simd" if simd else ""))
        out = np.zeros(width * height, np.uint16)
        rc = fn(_ptr(a, _u8p), C.c_size_t(a.size), _ptr(out, _u16p), width, height)
        if rc != 0:
            raise OracleError(rc, "ref mic_decompress")
        return out

    def decompress_parallel(self, blob, width, height, max_threads=0, out=None) -> np.ndarray:
        a = _as_u8(blob)
        if out is None:
            out = np.zeros(width * height, np.uint16)
        rc = self.lib.mic_decompress_parallel(_ptr(a, _u8p), C.c_size_t(a.size), _ptr(out, _u16p), width, height, max_threads)
        if rc != 0:
            raise OracleError(rc, "ref mic_decompress_parallel")
        return out
