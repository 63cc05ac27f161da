"""The GPU encoder's per-thread code (zb_encode.cuh) replayed on the CPU: every frame it writes must be accepted
by the oracle (= the reference decoder's rules) and by libzstd, round-trip bit-exact, and stay within the ratio
band of libzstd at the same level (the reference has no compressor to compare with)."""
import ctypes
import random

import pytest

from tests import helpers
from tools import corpus, zstd_ref


@pytest.fixture(scope="module")
def enc(hostsim, oracle):
    lib = hostsim.lib
    lib.hostsim_compress.restype = ctypes.c_uint32
    lib.hostsim_compress.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int]

    def compress(data, level, checksum=True, cap=None):
        data = bytes(data)
        cap = len(data) + len(data) // 128 + 128 if cap is None else cap
        buf = ctypes.create_string_buffer(max(cap, 1))
        r = lib.hostsim_compress(buf, cap, data, len(data), level, 1 if checksum else 0)
        if helpers.is_err(r):
            return r, None
        f = buf.raw[:r]
        if checksum:   # on the GPU the checksum kernel appends it; emulate with the oracle's XXH64
            f += (oracle.xxh64(data) & 0xFFFFFFFF).to_bytes(4, "little")
        return len(f), f
    return compress


SIZES = [0, 1, 2, 5, 17, 63, 64, 100, 255, 256, 1000, 1024, 4096, 5000, 16384, 65536, 70000, 131072, 131073, 200000, 300000]


def test_frames_are_accepted_and_round_trip(enc, oracle):
    rng = random.Random(5)
    for t in range(150):
        n = rng.choice(SIZES)
        data = helpers.sample_payload(rng, t % 6, n)
        for level in (1, 2, 3):
            for checksum in (True, False):
                r, f = enc(data, level, checksum)
                assert f is not None
                ro, oo, over = oracle.decompress(f, n)
                assert ro == n and oo == data, (t, n, level, hex(ro))
                assert zstd_ref.decompress(f, n) == data


def test_ratio_band_against_libzstd(enc):
    for kind in ("log", "tick"):
        raw = corpus.make(kind, 2 << 20).tobytes()
        for chunk in (65536, 131072):
            for level in (1, 2, 3):
                ours = sum(enc(raw[i:i + chunk], level)[0] for i in range(0, len(raw), chunk))
                ref = sum(len(zstd_ref.compress(raw[i:i + chunk], level)) for i in range(0, len(raw), chunk))
                assert ours <= ref * 1.03, (kind, chunk, level, ours, ref)


def test_warp_matcher_emulation_round_trips_and_ratio(hostsim, oracle):
    """The lock-step CPU emulation of the GPU's warp-parallel match finder (k_enc_match): valid frames, ratio band."""
    lib = hostsim.lib
    lib.hostsim_compress_warp.restype = ctypes.c_uint32
    lib.hostsim_compress_warp.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int]

    def comp(data, level):
        cap = len(data) + len(data) // 128 + 128
        buf = ctypes.create_string_buffer(max(cap, 1))
        r = lib.hostsim_compress_warp(buf, cap, bytes(data), len(data), level, 0)
        assert not helpers.is_err(r)
        return buf.raw[:r]
    rng = random.Random(9)
    for t in range(90):
        n = rng.choice(SIZES)
        data = helpers.sample_payload(rng, t % 6, n)
        for level in (1, 2, 3):
            f = comp(data, level)
            ro, oo, _ = oracle.decompress(f, n)
            assert ro == n and oo == data, (t, n, level, hex(ro))
    for kind in ("log", "tick"):
        raw = corpus.make(kind, 1 << 20).tobytes()
        for chunk in (65536, 131072):
            for level in (1, 2, 3):
                ours = sum(len(comp(raw[i:i + chunk], level)) for i in range(0, len(raw), chunk))
                ref = sum(len(zstd_ref.compress(raw[i:i + chunk], level, checksum=False)) for i in range(0, len(raw), chunk))
                assert ours <= ref * 1.03, (kind, chunk, level, ours, ref)


def test_incompressible_and_constant_inputs(enc, oracle):
    rnd = corpus.random_(200000).tobytes()
    for level in (1, 2, 3):
        r, f = enc(rnd, level)
        assert r <= len(rnd) + 3 * 2 + 9 + 4 + 8          # raw blocks: bounded expansion
        assert oracle.decompress(f, len(rnd))[1] == rnd
        r, f = enc(b"\x07" * 300000, level)
        assert r < 64                                      # RLE blocks
        assert oracle.decompress(f, 300000)[1] == b"\x07" * 300000


def test_destination_too_small(enc):
    data = corpus.log(5000).tobytes()
    r, f = enc(data, 3, True, cap=100)
    assert r == helpers.err(70) and f is None
