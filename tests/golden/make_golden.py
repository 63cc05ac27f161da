#!/usr/bin/env python3
"""Extracts the reference's own golden vectors into small binary fixtures.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py

Sources
  * csharp/test/TestDecompress.cs:58-89  -> csharp_alphabet.zst   (484-byte frame)
    expected plaintext: csharp/test/TestDecompress.cs:28-43 driven by the xorshift128+ of
    csharp/test/XorShift128Plus.cs:45-53, seed (42, 24) -> csharp_alphabet.raw (3409 bytes)
  * java/src/test/java/com/epam/deltix/zstd/TestDecompress.java:8-10 -> java_abc.zst (51 bytes)
    (the Java test asserts nothing about content; the frame carries an XXH64 content checksum and a
    100000-byte content size, and its only literal run is "abc..za", so the plaintext is the alphabet
    cycle; java_abc.raw is produced by the system libzstd decoding that frame)
  * csharp/src/ZStdDecompress.cs:833-934 -> default_tables.json (predefined LL/OF/ML decode tables,
    the KAT for the FSE table builder applied to ZStdInternal.cs:164-196)
"""
import ctypes, json, os, re, sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
M64 = (1 << 64) - 1


def xorshift_text():
    s = [42, 24]

    def nxt():
        x, y = s
        s[0] = y
        x ^= (x << 23) & M64
        s[1] = x ^ y ^ (x >> 17) ^ (y >> 26)
        return (s[1] + y) & M64

    alpha = "abcdefghijklmnopqrstuvwxyz"
    n = len(alpha)
    alpha2 = alpha + alpha
    out = []
    for _ in range(256):
        i = nxt() % n
        l = nxt() % n
        out.append(alpha2[i:i + l])
    return "".join(out).encode("ascii")


def main():
    cs = open(f"{REF}/csharp/test/TestDecompress.cs").read()
    body = cs[cs.index("byte[] compressedData = {"):]
    body = body[:body.index("};")]
    frame = bytes(int(h, 16) for h in re.findall(r"0x([0-9A-Fa-f]{2})", body))
    assert len(frame) == 484, len(frame)
    open(f"{HERE}/csharp_alphabet.zst", "wb").write(frame)
    raw = xorshift_text()
    assert len(raw) == 3409, len(raw)
    open(f"{HERE}/csharp_alphabet.raw", "wb").write(raw)

    jv = open(f"{REF}/java/src/test/java/com/epam/deltix/zstd/TestDecompress.java").read()
    body = jv[jv.index("compressedData = {"):]
    body = body[:body.index("};")]
    jframe = bytes(int(v) & 0xFF for v in re.findall(r"-?\d+", body[body.index("{"):]))
    assert len(jframe) == 51, len(jframe)
    open(f"{HERE}/java_abc.zst", "wb").write(jframe)
    z = ctypes.CDLL("libzstd.so.1")
    z.ZSTD_decompress.restype = ctypes.c_size_t
    z.ZSTD_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
    buf = ctypes.create_string_buffer(100000)
    n = z.ZSTD_decompress(buf, 100000, jframe, len(jframe))
    assert n == 100000, n
    jraw = buf.raw[:n]
    assert jraw == (b"abcdefghijklmnopqrstuvwxyz" * 4000)[:100000]
    # stored as a generator description, not 100 kB of bytes
    json.dump({"pattern": "abcdefghijklmnopqrstuvwxyz", "length": 100000},
              open(f"{HERE}/java_abc.raw.json", "w"))

    # predefined tables: rows "new SeqSymbol(a, b, c, d)" in declaration order; first row per table is the header
    src = open(f"{REF}/csharp/src/ZStdDecompress.cs").read()
    tables = {}
    for name in ("LL", "OF", "ML"):
        seg = src[src.index(f"{name}_defaultDTableArray = new SeqSymbol"):]
        seg = seg[:seg.index("};")]
        rows = re.findall(r"new SeqSymbol\(\s*(\d+),\s*(\d+),\s*(\d+),\s*([A-Za-z_0-9]+)\)", seg)
        cells = [[int(a), int(b), int(c), int(d)] for a, b, c, d in rows[1:]]
        tables[name] = cells
    assert len(tables["LL"]) == 64 and len(tables["OF"]) == 32 and len(tables["ML"]) == 64
    json.dump(tables, open(f"{HERE}/default_tables.json", "w"))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    sys.exit(main())
