"""The C-ABI library builds, loads and exports every symbol include/zstdb200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import zstandard_b200 as zb
from zstandard_b200 import build as zbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "zstdb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zstdb200_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    lib_path = zbuild.build()
    lib = ctypes.CDLL(lib_path)
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(zb.ABI.keys())


def test_host_only_entry_points():
    lib = zb.load_library()
    assert lib.zstdb200_version().startswith(b"zstdb200")
    assert lib.zstdb200_is_error((-20) & 0xFFFFFFFF) == 1 and lib.zstdb200_is_error(12345) == 0
    assert lib.zstdb200_compress_bound(65536) >= 65536 + 16
    # GetDecompressedSize is a pure header parse (ZStdDecompress.cs:590-622): check it against the golden vectors
    from tests import helpers
    for name, frame, raw in helpers.golden_vectors():
        assert zb.ZStdDecompress.GetDecompressedSize(frame) == len(raw), name
    assert zb.ZStdDecompress.GetDecompressedSize(helpers.skippable(b"abc")) == 0
    assert zb.ZStdDecompress.GetDecompressedSize(b"\x28\xb5\x2f") == 0


def test_get_decompressed_size_matches_oracle(oracle):
    from tests import helpers
    from tools import zstd_ref
    import random
    rng = random.Random(3)
    for frame, data in helpers.make_frames(31, 60):
        for f in (frame, frame[:rng.randrange(len(frame) + 1)], helpers.mutate(rng, frame) if len(frame) > 1 else frame):
            assert zb.ZStdDecompress.GetDecompressedSize(f) == oracle.get_decompressed_size(f)
