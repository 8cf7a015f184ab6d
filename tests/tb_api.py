"""ctypes wrapper of the TEST-ONLY CPU backend (tests/cpu_backend): runs csrc/build_host.hpp with sequential
stand-ins for the CUDA kernels so the host orchestration can be checked against the oracle without a GPU."""
import ctypes as C
import os

import numpy as np

from point_cloud_viewer_b200._native import NodeMeta
from point_cloud_viewer_b200 import node_name, ENC_BYTES

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpu_backend", "_build", "libtb.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_SO)
        L.tb_build.restype = C.c_void_p
        L.tb_build.argtypes = [C.c_uint64] + [C.c_void_p] * 3 + [C.c_uint64, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_char_p, C.c_int]
        L.tb_num_nodes.restype = C.c_uint64
        L.tb_num_nodes.argtypes = [C.c_void_p]
        L.tb_nodes.argtypes = [C.c_void_p, C.c_void_p]
        L.tb_xyz_bytes.restype = C.c_uint64
        L.tb_xyz_bytes.argtypes = [C.c_void_p]
        L.tb_passes.restype = C.c_uint32
        L.tb_passes.argtypes = [C.c_void_p]
        L.tb_download.argtypes = [C.c_void_p] * 5
        L.tb_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


class TbTree:
    def __init__(self, x, y, z, rgb, resolution, bmin, bmax, max_points, G, intensity=None, stride=1):
        L = lib()
        n = len(rgb) // 3
        err = C.create_string_buffer(512)
        h = L.tb_build(n, x.ctypes.data, y.ctypes.data, z.ctypes.data, stride, rgb.ctypes.data, intensity.ctypes.data if intensity is not None else None,
                       float(resolution), (C.c_double * 3)(*[float(v) for v in bmin]), (C.c_double * 3)(*[float(v) for v in bmax]), max_points, G, err, 512)
        if not h:
            raise RuntimeError(err.value.decode())
        nn = L.tb_num_nodes(h)
        arr = (NodeMeta * max(nn, 1))()
        L.tb_nodes(h, arr)
        self.passes = L.tb_passes(h)
        self.has_intensity = intensity is not None
        xb = L.tb_xyz_bytes(h)
        self.xyz, self.rgb, self.src = np.zeros(max(xb, 1), np.uint8), np.zeros(max(3 * n, 1), np.uint8), np.zeros(max(n, 1), np.uint32)
        self.inten = np.zeros(max(n, 1), np.float32) if intensity is not None else None
        L.tb_download(h, self.xyz.ctypes.data, self.rgb.ctypes.data, self.inten.ctypes.data if self.inten is not None else None, self.src.ctypes.data)
        L.tb_free(h)
        self.nodes = {}
        for i in range(nn):
            m = arr[i]
            self.nodes[node_name(m.id_high, m.id_low)] = dict(num_points=m.num_points, enc=m.position_encoding, cube=(m.cube_min[0], m.cube_min[1], m.cube_min[2], m.cube_edge),
                                                              point_offset=m.point_offset, xyz_byte_offset=m.xyz_byte_offset)

    def node_data(self, name):
        m = self.nodes[name]
        n, po, bo, bpc = m["num_points"], m["point_offset"], m["xyz_byte_offset"], ENC_BYTES[m["enc"]]
        return (self.xyz[bo:bo + n * 3 * bpc], self.rgb[3 * po:3 * (po + n)], self.inten[po:po + n] if self.inten is not None else None, self.src[po:po + n].astype(np.uint64))
