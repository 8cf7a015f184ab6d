"""The exact fast paths of csrc/chain.h (integer-built unit fractions, Markstein division with a known reciprocal,
clamp-free integer encode) return bit-identical results to the plain IEEE operators.  Host execution; the GPU parity
tests exercise the same functions on the device."""
import ctypes as C

import tb_api


def _lib():
    L = tb_api.lib()
    for f in ("tb_check_unit_frac", "tb_check_div", "tb_check_codec"):
        getattr(L, f).restype = C.c_uint64
    L.tb_check_div.argtypes = [C.c_uint64, C.c_uint64]
    L.tb_check_codec.argtypes = [C.c_uint64, C.c_uint64]
    return L


def test_unit_fractions_exhaustive():
    assert _lib().tb_check_unit_frac() == 0


def test_division_with_known_reciprocal():
    assert _lib().tb_check_div(20_000_000, 12345) == 0


def test_codec_fast_equals_plain():
    assert _lib().tb_check_codec(3_000_000, 99) == 0
