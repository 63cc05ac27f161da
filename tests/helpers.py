"""Shared test plumbing: oracle / host-simulator loaders and frame generators.

Only tests may touch oracle/ (it is the checker, never the product)."""
import ctypes
import json
import os
import random
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ERR_FLOOR = (-120) & 0xFFFFFFFF


def is_err(r):
    return int(r) > ERR_FLOOR


def err(code):
    return (-code) & 0xFFFFFFFF


class Oracle:
    def __init__(self):
        path = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
        src = os.path.join(ROOT, "oracle", "zstd_oracle.cpp")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
        o = ctypes.CDLL(path)
        o.oracle_decompress.restype = ctypes.c_uint32
        o.oracle_decompress.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint32]
        o.oracle_decompress_diag.restype = ctypes.c_uint32
        o.oracle_decompress_diag.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint32, ctypes.POINTER(ctypes.c_int)]
        o.oracle_get_decompressed_size.restype = ctypes.c_uint64
        o.oracle_get_decompressed_size.argtypes = [ctypes.c_char_p, ctypes.c_uint32]
        o.oracle_xxh64.restype = ctypes.c_uint64
        o.oracle_xxh64.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint64]
        o.oracle_default_table.restype = ctypes.c_uint32
        o.oracle_default_table.argtypes = [ctypes.c_int, ctypes.c_void_p]
        self.lib = o

    def decompress(self, frame, cap):
        """-> (result code, bytes or None, overread flag)"""
        frame = bytes(frame)
        buf = ctypes.create_string_buffer(max(cap, 1))
        ov = ctypes.c_int(0)
        r = self.lib.oracle_decompress_diag(buf, cap, frame, len(frame), ctypes.byref(ov))
        return r, (buf.raw[:r] if not is_err(r) else None), ov.value

    def get_decompressed_size(self, frame):
        frame = bytes(frame)
        return self.lib.oracle_get_decompressed_size(frame, len(frame))

    def xxh64(self, data, seed=0):
        data = bytes(data)
        return self.lib.oracle_xxh64(data, len(data), seed)


class HostSim:
    """g++ build of the GPU decoder's __host__ __device__ entropy stages (tests/hostsim/hostsim.cpp)."""

    def __init__(self):
        d = os.path.join(ROOT, "tests", "hostsim")
        path = os.path.join(d, "_build", "libhostsim.so")
        deps = [os.path.join(d, "hostsim.cpp")] + [os.path.join(ROOT, "zstandard_b200", "csrc", f)
                                                  for f in ("zb_common.cuh", "zb_format.cuh", "zb_decode.cuh", "zb_blocks.cuh", "zb_encode.cuh")] + [os.path.join(d, "serial_encoder.h")]
        if not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(p) for p in deps):
            os.makedirs(os.path.dirname(path), exist_ok=True)
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-static-libstdc++", "-static-libgcc",
                            "-o", path, deps[0]], check=True)
        h = ctypes.CDLL(path)
        h.hostsim_decompress.restype = ctypes.c_uint32
        h.hostsim_decompress.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint32,
                                         ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_int)]
        h.hostsim_decompress2.restype = ctypes.c_uint32
        h.hostsim_decompress2.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint32,
                                          ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_int),
                                          ctypes.POINTER(ctypes.c_uint32), ctypes.c_void_p]
        h.hostsim_set_par.argtypes = [ctypes.c_int]
        h.hostsim_par_frames.restype = ctypes.c_int
        h.hostsim_set_huf_root.argtypes = [ctypes.c_int]
        h.hostsim_set_seq_cap.argtypes = [ctypes.c_int]
        h.hostsim_deferred.restype = ctypes.c_int
        h.hostsim_class_frames.argtypes = [ctypes.c_int]
        h.hostsim_class_frames.restype = ctypes.c_int
        self.lib = h

    def set_par(self, on):
        """0: every frame on the frame-serial replay; 1 (default): sound multi-block frames take the block-parallel one."""
        self.lib.hostsim_set_par(int(on))

    def par_frames(self):
        return self.lib.hostsim_par_frames()

    def set_huf_root(self, root_log):
        """Root-table log of the replayed Huffman stage: 9 (k_huf<9>, frames of few literals; the default here because it
        exercises the long-code path on every larger table) or 11 (k_huf<11> / k_huf_blk)."""
        self.lib.hostsim_set_huf_root(int(root_log))

    def decompress(self, frame, cap, oracle):
        """One item through the replayed stages, one pass per data frame; the oracle's XXH64 stands in for k_xxh."""
        frame = bytes(frame)
        buf = ctypes.create_string_buffer(max(cap, 1))
        tr, nx, lb = ctypes.c_uint32(0), ctypes.c_int(0), ctypes.c_uint32(0)
        xxh = ctypes.cast(oracle.lib.oracle_xxh64, ctypes.c_void_p)
        r = self.lib.hostsim_decompress2(buf, cap, frame, len(frame), ctypes.byref(tr), ctypes.byref(nx), ctypes.byref(lb), xxh)
        out = buf.raw[:r] if not is_err(r) else None
        return r, out


def golden_vectors():
    """[(name, frame bytes, expected plaintext)] from the reference's own tests (tests/golden/make_golden.py)."""
    cs = open(os.path.join(GOLDEN, "csharp_alphabet.zst"), "rb").read()
    cs_raw = open(os.path.join(GOLDEN, "csharp_alphabet.raw"), "rb").read()
    jv = open(os.path.join(GOLDEN, "java_abc.zst"), "rb").read()
    spec = json.load(open(os.path.join(GOLDEN, "java_abc.raw.json")))
    pat = spec["pattern"].encode()
    jv_raw = (pat * (spec["length"] // len(pat) + 1))[:spec["length"]]
    return [("csharp_alphabet", cs, cs_raw), ("java_abc", jv, jv_raw)]


def sample_payload(rng, kind, n):
    from tools import corpus
    if kind == 0:
        return corpus.log(n, (rng.randrange(1 << 30), 24)).tobytes()
    if kind == 1:
        return corpus.tick(n, (rng.randrange(1 << 30), 25)).tobytes()
    if kind == 2:
        return corpus.random_(n, (rng.randrange(1 << 30), 26)).tobytes()
    if kind == 3:
        return bytes([rng.choice(b"ab")]) * n
    if kind == 4:
        return bytes(rng.choice(b"abcdefgh") for _ in range(n))
    words = [bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz ") for _ in range(rng.randint(2, 9))) for _ in range(50)]
    return b" ".join(rng.choice(words) for _ in range(n // 5 + 1))[:n]


SIZES = [0, 1, 2, 5, 17, 100, 1000, 4096, 5000, 65536, 70000, 131072, 200000, 300000]
LEVELS = [1, 2, 3, 3, 3, 5, 9, 19, -1, -5]


def make_frames(seed, count, sizes=SIZES, levels=LEVELS):
    """Deterministic mix of libzstd frames: all literal modes, table modes, raw/RLE blocks, multi-block, +/- checksum."""
    from tools import zstd_ref
    rng = random.Random(seed)
    out = []
    for t in range(count):
        n = rng.choice(sizes)
        data = sample_payload(rng, t % 6, n)
        frame = zstd_ref.compress(data, rng.choice(levels), checksum=rng.random() < 0.7, content_size=rng.random() < 0.8)
        out.append((frame, data))
    return out


def mutate(rng, frame):
    b = bytearray(frame)
    mode = rng.randrange(4)
    if mode == 0:
        i = rng.randrange(len(b)); b[i] ^= 1 << rng.randrange(8)
    elif mode == 1:
        b = b[:rng.randrange(len(b))]
    elif mode == 2:
        for _ in range(3):
            i = rng.randrange(len(b)); b[i] = rng.randrange(256)
    else:
        i = rng.randrange(min(len(b), 40)); b[i] ^= 1 << rng.randrange(8)
    return bytes(b)


def large_window_frames(seed=5, count=60):
    """libzstd frames whose header declares a window above 16 MiB (descriptors 0x71: 18 MiB, 0x78: 32 MiB, 0x80: 64 MiB,
    0x8b: 88 MiB) over ordinary small contents — a declared window may always be larger than needed.  Above 16 MiB the
    reference switches blocks whose offset table has enough long-offset cells (the predefined table does) to the
    sequence loop that runs four sequences ahead (ZStdDecompress.cs:1898-1905); above 32 MiB offsets of 25+ bits
    are read in two parts (:1494-1501 / :1642-1647)."""
    from tools import zstd_ref
    rng = random.Random(seed)
    out = []
    for t in range(count):
        n = rng.choice([60, 200, 700, 1500, 4000, 20000, 70000])
        data = sample_payload(rng, t % 6, n)
        frame = bytearray(zstd_ref.compress(data, rng.choice(LEVELS), checksum=rng.random() < 0.5, content_size=False))
        assert not (frame[4] & 0x20)                  # not single-segment: byte 5 is the window descriptor
        frame[5] = rng.choice([0x71, 0x78, 0x80, 0x8b])
        out.append((bytes(frame), data))
    return out


def skippable(payload, nibble=0):
    return (0x184D2A50 + nibble).to_bytes(4, "little") + len(payload).to_bytes(4, "little") + payload


def multi_frame_items(seed=11, count=40):
    """[(item bytes, capacity)] of items holding several data frames (DecompressMultiFrame, ZStdDecompress.cs:2096-2160):
    2..6 frames with and without checksum, skippable frames between them, one corrupted / truncated / short-capacity
    variant each."""
    rng = random.Random(seed)
    pool = make_frames(seed, 60, sizes=[0, 1, 100, 1000, 5000, 70000, 140000])
    items = []
    for t in range(count):
        k = rng.randint(2, 6)
        parts = [pool[rng.randrange(len(pool))] for _ in range(k)]
        blob, total = b"", 0
        for frame, data in parts:
            if rng.random() < 0.3:
                blob += skippable(bytes(rng.randrange(256) for _ in range(rng.randint(0, 20))))
            blob += frame
            total += len(data)
        if rng.random() < 0.3:
            blob += skippable(b"end")
        items.append((blob, total))
        items.append((blob, total + 7))
        if total:
            items.append((blob, max(0, total - 1 - rng.randrange(min(total, 3000)))))   # some frame no longer fits
        items.append((blob[:len(blob) - 1 - rng.randrange(min(len(blob) - 1, 200))], total))   # truncated in the tail
        last = parts[-1][0]
        if len(last) > 12:                                                  # corruption inside the last data frame
            pos = len(blob) - rng.randrange(4, min(len(last), 60))
            mutated = bytearray(blob); mutated[pos] ^= 1 << rng.randrange(8)
            items.append((bytes(mutated), total))
        first = parts[0][0]
        if len(first) > 12 and len(parts[0][1]) > 0:                        # corruption inside the first data frame
            mutated = bytearray(blob); mutated[blob.index(first) + len(first) - 2] ^= 0x10
            items.append((bytes(mutated), total))
    return items


def huf12_frame(n=700, four_streams=True, seed=5, raw_prefix=0):
    """Hand-made frame whose Huffman table has tableLog 12 — legal for the reference (HufDecompress.cs:128, ReadStats
    EntropyCommon.cs:198-269) although no encoder emits it.  Direct 4-bit weights 11,11,11,10,9,...,2,1 for symbols
    0..12, the 14th symbol's weight (1) is implied; literals only, no sequences.  Returns (frame, plaintext)."""
    rng = random.Random(seed)
    weights = [11, 11, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 1]
    # canonical positions as the decoder lays them out: ascending weight, ascending symbol within a weight
    start, pos = {}, 0
    for w in range(1, 12):
        for sym, ws in enumerate(weights):
            if ws == w:
                start[sym] = pos
                pos += 1 << (w - 1)
    assert pos == 4096
    code = {sym: (start[sym] >> (w - 1), 13 - w) for sym, w in enumerate(weights)}      # (value, nbBits)
    data = bytes(rng.choice([0, 0, 0, 1, 1, 2, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13]) for _ in range(n))

    def stream(symbols):
        acc, nbits = 0, 0
        for sym in reversed(symbols):                 # the decoder reads backwards: last written is first read
            v, nb = code[sym]
            acc |= v << nbits
            nbits += nb
        acc |= 1 << nbits                             # end mark
        return acc.to_bytes(nbits // 8 + 1, "little")

    if four_streams:
        seg = (n + 3) // 4
        parts = [stream(data[0:seg]), stream(data[seg:2 * seg]), stream(data[2 * seg:3 * seg]), stream(data[3 * seg:])]
        body = b"".join(len(p).to_bytes(2, "little") for p in parts[:3]) + b"".join(parts)
    else:
        body = stream(data)
    nib = weights[:13] + [0]
    hdr = bytes([127 + 13]) + bytes((nib[i] << 4) | nib[i + 1] for i in range(0, 14, 2))
    comp = hdr + body
    assert n < 1024 and len(comp) < 1024
    lhc = 2 | ((1 if four_streams else 0) << 2) | (n << 4) | (len(comp) << 14)
    block = lhc.to_bytes(3, "little") + comp + b"\x00"                                  # + "0 sequences"
    blocks = b""
    plain = b""
    if raw_prefix:                                                                      # a raw block first: table built in block 1
        pre = bytes(rng.randrange(256) for _ in range(raw_prefix))
        blocks += ((raw_prefix << 3) | 0).to_bytes(3, "little") + pre
        plain += pre
    blocks += ((len(block) << 3) | (2 << 1) | 1).to_bytes(3, "little") + block
    plain += data
    total = len(plain)
    frame = b"\x28\xb5\x2f\xfd" + bytes([0x60]) + (total - 256).to_bytes(2, "little") + blocks    # single segment, 2-byte FCS
    return frame, plain


def header_variant_frames(seed=3):
    """[(frame, capacity)]: every frame-header shape the reference parses (ZStdDecompress.cs:389-499) around one raw
    block — FCS field of 0/1/2/4/8 bytes, single-segment or window descriptor (incl. too large), dictionary-id field of
    0/1/2/4 bytes (zero and non-zero), checksum flag with a right and a wrong checksum, the reserved bit, FCS that
    disagrees with the content, truncations inside the header."""
    rng = random.Random(seed)
    import ctypes
    out = []
    xxh = Oracle().xxh64
    for n in (0, 1, 200, 255, 256, 300, 65535 + 256, 70000):
        payload = bytes(rng.randrange(256) for _ in range(n))
        block = ((n << 3) | 1).to_bytes(3, "little") + payload                      # one raw block, last
        for single in (0, 1):
            for fcs_id in (0, 1, 2, 3):
                for did in (0, 1, 2, 3):
                    for checksum in (0, 1):
                        fhd = (fcs_id << 6) | (single << 5) | (checksum << 2) | did
                        hdr = b"\x28\xb5\x2f\xfd" + bytes([fhd])
                        if not single:
                            hdr += bytes([rng.choice([0x00, 0x38, 0x50, 0xA0, 0xA8, 0xF8])])     # window descriptor (some > 2^30)
                        dict_id = rng.choice([0, 0, 0, 7])
                        hdr += dict_id.to_bytes([0, 1, 2, 4][did], "little") if did else b""
                        if fcs_id == 0:
                            hdr += bytes([n & 0xFF]) if single else b""
                        elif fcs_id == 1:
                            hdr += ((n - 256) & 0xFFFF).to_bytes(2, "little")
                        elif fcs_id == 2:
                            hdr += n.to_bytes(4, "little")
                        else:
                            hdr += n.to_bytes(8, "little")
                        frame = hdr + block
                        if checksum:
                            frame += (xxh(payload) & 0xFFFFFFFF).to_bytes(4, "little")
                        out.append((frame, n))
                        if rng.random() < 0.15:
                            out.append((frame, max(0, n - 1)))                                   # capacity one short
                        if rng.random() < 0.15 and checksum:
                            bad = bytearray(frame); bad[-1] ^= 0x40
                            out.append((bytes(bad), n))                                          # wrong checksum
                        if rng.random() < 0.1:
                            out.append((frame[:rng.randrange(1, len(hdr) + 3)], n))               # cut inside the header / block header
                        if rng.random() < 0.05:
                            bad = bytearray(frame); bad[4] |= 0x08
                            out.append((bytes(bad), n))                                          # reserved bit
    return out


def sequence_count_frames():
    """[(frame, plaintext, nbSeq of its first block)] covering the three encodings of the sequence count
    (DecodeSeqHeaders :1121-1136): one byte (< 128), two bytes (< 0x7F00) and 255 + u16 (>= 0x7F00)."""
    from tools import zstd_ref
    rng = random.Random(1)
    toks = [bytes(rng.randrange(256) for _ in range(3)) for _ in range(64)]
    dense = b"".join(bytes([rng.randrange(256)]) + rng.choice(toks) for _ in range(40000))[:131072]
    out = []
    o = Oracle()
    o.lib.oracle_decompress_trace.restype = ctypes.c_uint32
    o.lib.oracle_trace_nseq.restype = ctypes.c_uint64
    for data, level in ((dense[:400], 19), (dense[:20000], 19), (dense, 19)):
        f = zstd_ref.compress(data, level, checksum=True)
        dst = ctypes.create_string_buffer(len(data))
        assert o.lib.oracle_decompress_trace(dst, len(data), f, len(f)) == len(data)
        out.append((f, data, int(o.lib.oracle_trace_nseq())))
    return out


def rle_modes_frame(nseq=100, tail=3, seed=9, lit_byte=0x78):
    """Hand-made frame for the modes no encoder at hand emits: an RLE literals section (DecodeLiteralsBlock :755-772,
    type 1) and all three sequence tables in RLE mode (BuildSeqTable :1046-1050, mode 1: one symbol, zero state bits).
    Every sequence is litLength 1 (code 1), matchLength 8 (code 5), offset code 2 (offset 1..4 from 2 extra bits), so
    the bitstream is just 2 bits per sequence.  Returns (frame, plaintext)."""
    rng = random.Random(seed)
    extras = [0] + [rng.randrange(4) for _ in range(nseq - 1)]      # the first match can only reach back 1 byte
    out = bytearray()
    for e in extras:
        out.append(lit_byte)
        off = 1 + e
        for _ in range(8):
            out.append(out[-off])
    out += bytes([lit_byte]) * tail
    nlit = nseq + tail
    if nlit < 32:
        lit = bytes([(nlit << 3) | 1])
    elif nlit < 4096:
        lit = (1 | (1 << 2) | (nlit << 4)).to_bytes(2, "little")
    else:
        lit = (1 | (3 << 2) | (nlit << 4)).to_bytes(3, "little")
    lit += bytes([lit_byte])
    if nseq < 128:
        cnt = bytes([nseq])
    elif nseq < 0x7F00:
        cnt = bytes([(nseq >> 8) + 128, nseq & 0xFF])
    else:
        cnt = b"\xff" + (nseq - 0x7F00).to_bytes(2, "little")
    acc, nbits = 0, 0
    for e in reversed(extras):                                      # the decoder reads sequence 0 first, from the top
        acc |= e << nbits
        nbits += 2
    acc |= 1 << nbits
    stream = acc.to_bytes(nbits // 8 + 1, "little")
    seq = cnt + bytes([(1 << 6) | (1 << 4) | (1 << 2), 1, 2, 5]) + stream   # modes RLE/RLE/RLE; symbols LL 1, OF 2, ML 5
    block = lit + seq
    assert len(block) < (1 << 17)
    total = len(out)
    hdr = b"\x28\xb5\x2f\xfd" + bytes([0xA0]) + total.to_bytes(4, "little")             # single segment, 4-byte FCS
    frame = hdr + ((len(block) << 3) | (2 << 1) | 1).to_bytes(3, "little") + block
    return frame, bytes(out)


def frame_modes(frame):
    """Counter of the format modes one data frame uses: ("block", type), ("lit", type, sizeFormat), ("nseq", bytes),
    ("LL" | "OF" | "ML", mode) — a coverage probe for the test corpora (walks headers only)."""
    import collections
    c = collections.Counter()
    if frame[:4] != b"\x28\xb5\x2f\xfd":
        return c
    fhd = frame[4]; did = fhd & 3; single = (fhd >> 5) & 1; fcs = fhd >> 6
    p = 5 + (0 if single else 1) + [0, 1, 2, 4][did] + [1 if single else 0, 2, 4, 8][fcs]
    while p + 3 <= len(frame):
        h = int.from_bytes(frame[p:p + 3], "little"); p += 3
        last, t, sz = h & 1, (h >> 1) & 3, h >> 3
        c[("block", t)] += 1
        if t == 2:
            b = frame[p:p + sz]
            lt, sf = b[0] & 3, (b[0] >> 2) & 3
            c[("lit", lt, sf)] += 1
            if lt < 2:
                lh = [1, 2, 1, 3][sf]
                n = (b[0] >> 3) if lh == 1 else (int.from_bytes(b[:lh], "little") >> 4)
                q = lh + (n if lt == 0 else 1)
            else:
                lh = [3, 3, 4, 5][sf]; v = int.from_bytes(b[:5], "little")
                cs = ((v >> 14) & 0x3FF) if lh == 3 else (((v >> 18) & 0x3FFF) if lh == 4 else ((v >> 22) & 0x3FFFF))
                q = lh + cs
            if q < len(b):
                ns = b[q]
                if ns == 0:
                    c[("nseq", 0)] += 1
                else:
                    w = 1 if ns < 128 else (2 if ns < 255 else 3)
                    c[("nseq", w)] += 1
                    m = b[q + w]
                    c[("LL", (m >> 6) & 3)] += 1; c[("OF", (m >> 4) & 3)] += 1; c[("ML", (m >> 2) & 3)] += 1
            p += sz
        elif t == 0:
            p += sz
        elif t == 1:
            p += 1
        else:
            break
        if last:
            break
    return c


def repeat_mode_frames():
    """[(frame, plaintext)] whose blocks use repeat-mode (mode 3) LL, OF and ML tables and treeless literals
    (libzstd level 19 on stationary multi-block inputs)."""
    from tools import zstd_ref
    rng = random.Random(7)
    B = bytes(rng.randrange(256) for _ in range(8))
    a = b"".join(bytes([rng.randrange(256)]) + B for _ in range(40000))
    toks = [bytes(rng.randrange(256) for _ in range(6)) for _ in range(200)]
    b = b"".join(rng.choice(toks) for _ in range(131072 // 6 + 1))[:131072] + b"".join(b"x" + rng.choice(toks) for _ in range(8000))
    return [(zstd_ref.compress(d, 19, checksum=True), d) for d in (a, b)]
