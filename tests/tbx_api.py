"""ctypes wrapper of the TEST-ONLY sequential drivers of csrc/xray_pyramid.h (tests/cpu_backend/xray_pyramid_cpu.cpp): the
PCV_HD functions the X-ray pyramid kernels call, run with plain loops so that they can be checked against the oracle
without a GPU."""
import ctypes as C
import os

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpu_backend", "_build", "libtbx.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_SO)
        L.tbx_build_parent.argtypes = [C.POINTER(C.c_void_p), C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
        L.tbx_background.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.tbx_binned.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p,
                                 C.c_void_p]
        L.tbx_bin_of.restype = C.c_int64
        L.tbx_bin_of.argtypes = [C.c_float, C.c_double]
        L.tbx_quad_rect_of.argtypes = [C.c_uint8, C.c_uint64, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.tbx_rect_and_levels.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_uint32, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.tbx_encode_png.restype = C.c_int64
        L.tbx_encode_png.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64]
        L.tbx_encode_xray_meta.restype = C.c_int64
        L.tbx_encode_xray_meta.argtypes = [C.c_double, C.c_double, C.c_double, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        L.tbx_node_name.argtypes = [C.c_uint8, C.c_uint64, C.c_char_p, C.c_int]
        _lib = L
    return _lib


def encode_png(rgba):
    img = np.ascontiguousarray(rgba, np.uint8)
    out = np.zeros(img.size + img.shape[0] + (1 << 16), np.uint8)
    n = lib().tbx_encode_png(img.ctypes.data, img.shape[1], img.shape[0], out.ctypes.data, len(out))
    assert n > 0
    return out[:n].tobytes()


def encode_xray_meta(min_x, min_y, edge, deepest, tile, nodes):
    lv = np.array([l for l, _ in nodes], np.uint32)
    ix = np.array([i for _, i in nodes], np.uint64)
    out = np.zeros(1 << 16, np.uint8)
    n = lib().tbx_encode_xray_meta(min_x, min_y, edge, deepest, tile, lv.ctypes.data, ix.ctypes.data, len(nodes), out.ctypes.data, len(out))
    assert n > 0
    return out[:n].tobytes()


def node_name(level, index):
    buf = C.create_string_buffer(80)
    lib().tbx_node_name(level, index, buf, 80)
    return buf.value.decode()


def build_parent(children, background, tile_px):
    child_px = next(c.shape[0] for c in children if c is not None)
    keep = [np.ascontiguousarray(c, np.uint8) if c is not None else None for c in children]
    ptrs = (C.c_void_p * 4)(*[c.ctypes.data if c is not None else None for c in keep])
    bg = np.asarray(background, np.uint8)
    out = np.zeros((tile_px, tile_px, 4), np.uint8)
    assert lib().tbx_build_parent(ptrs, child_px, bg.ctypes.data, tile_px, out.ctypes.data) == 0
    return out


def background(rgba, bg):
    out = np.ascontiguousarray(rgba, np.uint8).copy()
    b = np.asarray(bg, np.uint8)
    lib().tbx_background(out.ctypes.data, out.size // 4, b.ctypes.data)
    return out


def binned(pixel, attr, value, bin_size, npix, stride, bin_cap=1 << 10, col_cap=None):
    pixel = np.ascontiguousarray(pixel, np.uint32)
    attr = np.ascontiguousarray(attr, np.float32)
    value = np.ascontiguousarray(value, np.float32).reshape(len(pixel), -1)
    ncomp = value.shape[1]
    col_cap = col_cap or 2 * len(pixel) + 1024
    pix_sum = np.zeros(npix * stride, np.float32)
    pix_bins = np.zeros(npix, np.uint32)
    err = lib().tbx_binned(len(pixel), pixel.ctypes.data, attr.ctypes.data, value.ctypes.data, ncomp, float(bin_size), bin_cap, col_cap, npix, stride,
                           pix_sum.ctypes.data, pix_bins.ctypes.data)
    return err, pix_sum.reshape(npix, stride), pix_bins


def bin_of(attr, bin_size):
    return int(lib().tbx_bin_of(float(attr), float(bin_size)))


def quad_rect_of(level, index, root):
    out = (C.c_double * 3)()
    lib().tbx_quad_rect_of(level, index, (C.c_double * 3)(*root), out)
    return tuple(out)


def rect_and_levels(bmin, bmax, tile_px, pixel_size_m):
    rect = (C.c_double * 3)()
    lv = C.c_int()
    rc = lib().tbx_rect_and_levels((C.c_double * 3)(*bmin), (C.c_double * 3)(*bmax), tile_px, pixel_size_m, rect, C.byref(lv))
    return None if rc else (tuple(rect), lv.value)
