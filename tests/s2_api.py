"""ctypes access to the oracle's S2 restatement (oracle/oracle_s2.hpp) and to the TEST-ONLY sequential drivers of the product's
csrc/s2.h (tests/cpu_backend/s2_cpu.cpp)."""
import ctypes as C
import os

import numpy as np

import oracle_api as O

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpu_backend", "_build", "libtbs.so")
_tb = None
_orc = None
u64p = C.POINTER(C.c_uint64)


def orc():
    global _orc
    if _orc is None:
        L = O.lib()
        L.orc_s2_cell_ids.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
        L.orc_s2_face_ij.argtypes = [C.c_uint64] + [C.POINTER(C.c_int)] * 3
        L.orc_s2_centre.argtypes = [C.c_uint64, C.POINTER(C.c_double)]
        L.orc_s2_from_face_ij.restype = C.c_uint64
        L.orc_s2_from_face_ij.argtypes = [C.c_int] * 3
        L.orc_s2_parent.restype = C.c_uint64
        L.orc_s2_parent.argtypes = [C.c_uint64, C.c_int]
        L.orc_s2_next.restype = C.c_uint64
        L.orc_s2_next.argtypes = [C.c_uint64]
        L.orc_s2_level.argtypes = [C.c_uint64]
        L.orc_s2_token.argtypes = [C.c_uint64, C.c_char_p, C.c_int]
        L.orc_s2_normalize.restype = C.c_uint64
        L.orc_s2_normalize.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_s2_union_test.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.orc_s2_split.restype = C.c_void_p
        L.orc_s2_split.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
        L.orc_s2_split_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), u64p, C.POINTER(C.c_double), u64p]
        L.orc_s2_split_cells.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_s2_split_free.argtypes = [C.c_void_p]
        _orc = L
    return _orc


def tb():
    global _tb
    if _tb is None:
        L = C.CDLL(_SO)
        L.tbs_cell_ids.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        L.tbs_from_face_ij.restype = C.c_uint64
        L.tbs_from_face_ij.argtypes = [C.c_int] * 3
        L.tbs_normalize.restype = C.c_uint64
        L.tbs_normalize.argtypes = [C.c_void_p, C.c_uint64]
        L.tbs_union_test.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.tbs_token.argtypes = [C.c_uint64, C.c_char_p, C.c_int]
        L.tbs_from_token.argtypes = [C.c_char_p, u64p]
        L.tbs_level.argtypes = [C.c_uint64]
        L.tbs_is_valid.argtypes = [C.c_uint64]
        _tb = L
    return _tb


def _xyz(P):
    P = np.ascontiguousarray(P, np.float64)
    return [np.ascontiguousarray(P[:, k]) for k in range(3)]


def oracle_cell_ids(P, level):
    x, y, z = _xyz(P)
    out = np.zeros(len(x), np.uint64)
    orc().orc_s2_cell_ids(len(x), x.ctypes.data, y.ctypes.data, z.ctypes.data, 1, level, out.ctypes.data)
    return out


def product_cell_ids(P, level):
    x, y, z = _xyz(P)
    out = np.zeros(len(x), np.uint64)
    valid = np.zeros(len(x), np.uint8)
    tb().tbs_cell_ids(len(x), x.ctypes.data, y.ctypes.data, z.ctypes.data, 1, level, out.ctypes.data, valid.ctypes.data)
    return out, valid.astype(bool)


def face_ij(cid):
    f, i, j = C.c_int(), C.c_int(), C.c_int()
    orc().orc_s2_face_ij(int(cid), C.byref(f), C.byref(i), C.byref(j))
    return f.value, i.value, j.value


def centre(cid):
    out = (C.c_double * 3)()
    orc().orc_s2_centre(int(cid), out)
    return np.array(out)


def token(cid, product=False):
    buf = C.create_string_buffer(32)
    (tb().tbs_token if product else orc().orc_s2_token)(int(cid), buf, 32)
    return buf.value.decode()


def normalize(ids, product=False):
    a = np.ascontiguousarray(ids, np.uint64).copy()
    n = (tb().tbs_normalize if product else orc().orc_s2_normalize)(a.ctypes.data, len(a))
    return a[:n]


def union_test(cu, ids, product=False):
    cu = np.ascontiguousarray(cu, np.uint64)
    ids = np.ascontiguousarray(ids, np.uint64)
    c = np.zeros(len(ids), np.uint8)
    i = np.zeros(len(ids), np.uint8)
    (tb().tbs_union_test if product else orc().orc_s2_union_test)(cu.ctypes.data, len(cu), ids.ctypes.data, len(ids), c.ctypes.data, i.ctypes.data)
    return c.astype(bool), i.astype(bool)


def split(P, level):
    """S2Splitter::write + get_meta restated: dict(ok, bad_index, bmin, bmax, ids, counts, order)."""
    x, y, z = _xyz(P)
    h = orc().orc_s2_split(len(x), x.ctypes.data, y.ctypes.data, z.ctypes.data, 1, level)
    try:
        ok, bad, nc = C.c_int(), C.c_uint64(), C.c_uint64()
        bb = (C.c_double * 6)()
        orc().orc_s2_split_info(h, C.byref(ok), C.byref(bad), bb, C.byref(nc))
        ids = np.zeros(nc.value, np.uint64)
        counts = np.zeros(nc.value, np.uint64)
        order = np.zeros(len(x) if ok.value else 0, np.uint64)
        if ok.value:
            orc().orc_s2_split_cells(h, ids.ctypes.data, counts.ctypes.data, order.ctypes.data)
        return dict(ok=bool(ok.value), bad_index=bad.value, bmin=np.array(bb[:3]), bmax=np.array(bb[3:]), ids=ids, counts=counts, order=order)
    finally:
        orc().orc_s2_split_free(h)
