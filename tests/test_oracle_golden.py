"""Pins the CPU oracle: the reference's own golden vectors and KATs, then agreement with libzstd 1.5.5."""
import ctypes
import json
import os
import random

from tests import helpers
from tools import zstd_ref


def test_reference_golden_vectors(oracle):
    # csharp/test/TestDecompress.cs:53-99 and java/src/test/java/com/epam/deltix/zstd/TestDecompress.java:6-20
    for name, frame, raw in helpers.golden_vectors():
        assert oracle.get_decompressed_size(frame) == len(raw), name
        r, out, _ = oracle.decompress(frame, len(raw))
        assert r == len(raw) and out == raw, name


def test_golden_checksums(oracle):
    # stored frame checksums are the low 32 bits of XXH64(seed 0) of the content
    for name, frame, raw in helpers.golden_vectors():
        assert oracle.xxh64(raw) & 0xFFFFFFFF == int.from_bytes(frame[-4:], "little"), name


def test_xxh64_known_answers(oracle):
    assert oracle.xxh64(b"") == 0xEF46DB3751D8E999
    assert oracle.xxh64(b"a") == 0xD24EC4F1A98C6E5B
    assert oracle.xxh64(b"abc") == 0x44BC2CF5AD770999
    assert oracle.xxh64(b"Nobody inspects the spammish repetition") == 0xFBCEA83C8A378BF1


def test_predefined_tables_match_reference(oracle):
    # csharp/src/ZStdDecompress.cs:833-934 (LL_/OF_/ML_defaultDTable) = BuildFSETable(default norms)
    tables = json.load(open(os.path.join(helpers.GOLDEN, "default_tables.json")))
    for i, name in enumerate(("LL", "OF", "ML")):
        arr = (ctypes.c_uint32 * (4 * 64))()
        lg = oracle.lib.oracle_default_table(i, arr)
        got = [[arr[4 * k], arr[4 * k + 1], arr[4 * k + 2], arr[4 * k + 3]] for k in range(1 << lg)]
        assert got == tables[name], name


def test_agrees_with_libzstd_on_valid_frames(oracle):
    for frame, data in helpers.make_frames(101, 240):
        r, out, _ = oracle.decompress(frame, len(data))
        assert r == len(data) and out == data
        assert zstd_ref.decompress(frame, len(data)) == data


def test_error_codes_on_simple_corruptions(oracle):
    frame, data = helpers.make_frames(5, 8, sizes=[5000])[1]
    n = len(data)
    assert oracle.decompress(frame, n - 1)[0] == helpers.err(70)                 # dstSize_tooSmall
    assert oracle.decompress(b"\x00" * 8 + frame, n)[0] == helpers.err(10)        # prefix_unknown
    assert oracle.decompress(frame[:3], n)[0] == helpers.err(72)                 # srcSize_wrong
    assert oracle.decompress(b"", n)[0] == 0                                      # empty input decodes to nothing
    f2 = bytearray(frame); f2[4] |= 0x08
    assert oracle.decompress(bytes(f2), n)[0] == helpers.err(14)                  # reserved bit
    f3 = bytearray(zstd_ref.compress(data, 3, checksum=True)); f3[-1] ^= 0xFF
    assert oracle.decompress(bytes(f3), n)[0] == helpers.err(22)                  # checksum_wrong


def test_multiframe_and_skippable(oracle):
    (f1, d1), (f2, d2) = helpers.make_frames(9, 2, sizes=[3000, 70000])
    item = helpers.skippable(b"hello") + f1 + helpers.skippable(b"", 3) + f2 + helpers.skippable(b"x" * 20)
    r, out, _ = oracle.decompress(item, len(d1) + len(d2))
    assert r == len(d1) + len(d2) and out == d1 + d2
    assert oracle.get_decompressed_size(helpers.skippable(b"abc")) == 0
