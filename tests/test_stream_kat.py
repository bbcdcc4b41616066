"""Known-answer tests of the sample stream (reference utility.h:90-103 = std::mt19937() + generate_canonical)."""
import ctypes as C

import numpy as np

import parity as P


def lib():
    L = C.CDLL(P.ORACLE_LIB, mode=C.RTLD_LOCAL)
    L.orc_mt19937_nth.restype = C.c_uint32
    L.orc_mt19937_nth.argtypes = [C.c_uint32, C.c_uint64]
    L.orc_random01_from_u32.restype = C.c_float
    L.orc_random01_from_u32.argtypes = [C.c_uint32]
    return L


def test_mt19937_known_answers():
    L = lib()
    # first outputs of std::mt19937(5489) and the 10000th (ISO C++ [rand.predef])
    assert [L.orc_mt19937_nth(5489, i) for i in range(3)] == [3499211612, 581869302, 3890346734]
    assert L.orc_mt19937_nth(5489, 9999) == 4123659995


def test_mt19937_matches_numpy():
    L = lib()
    rs = np.random.RandomState()
    bg = np.random.MT19937()
    bg._legacy_seeding(5489)
    raw = np.random.Generator(bg).bit_generator.random_raw(2000)
    for i in (0, 1, 623, 624, 625, 1247, 1999):
        assert L.orc_mt19937_nth(5489, i) == int(raw[i])


def test_random01_conversion():
    L = lib()
    assert L.orc_random01_from_u32(0) == 0.0
    assert L.orc_random01_from_u32(1 << 31) == 0.5
    # float(u) rounds to 2^32 for the top 128 values: generate_canonical clamps below 1
    assert L.orc_random01_from_u32(0xFFFFFFFF) == np.nextafter(np.float32(1), np.float32(0))
    assert L.orc_random01_from_u32(0xFFFFFF7F) < 1.0
    u = 3499211612
    assert L.orc_random01_from_u32(u) == np.float32(np.float32(u) / np.float32(4294967296.0))
