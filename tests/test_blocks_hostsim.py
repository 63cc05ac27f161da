"""Block-parallel decoding of multi-block frames (zstandard_b200/csrc/zb_blocks.cuh), replayed on the CPU.

Every compressed block of a structurally sound multi-block frame is decoded as a unit of its own: tables it inherits
("repeat" sequence tables, "treeless" literals — ZStdDecompress.cs:1062-1064, 746-747) are rebuilt from the header of the
block that defined them, the repeat offsets (:1576, :1596) are carried symbolically and resolved in block order.  These
tests drive that replay (k_parse's walk, k_huf_blk, k_seq_blk, a serial stand-in for k_exec_big) against the oracle and
check that it was the block-parallel path that ran."""
import ctypes
import random

from tests import helpers
from tests.test_dictionary import _dict_oracle


def _big_frames(seed, count):
    from tools import zstd_ref
    rng = random.Random(seed)
    out = []
    for t in range(count):
        n = rng.choice([140000, 200000, 262144, 300000, 524288, 1 << 20, 1500000])
        data = helpers.sample_payload(rng, t % 6, n)
        if t % 7 == 3:      # a raw stretch in the middle: raw blocks between compressed ones
            data = data[:n // 3] + helpers.sample_payload(rng, 2, 150000) + data[n // 3:]
        frame = zstd_ref.compress(data, rng.choice([1, 2, 3, 3, 3, 5, 9, 19, -1]), checksum=rng.random() < 0.7, content_size=rng.random() < 0.8)
        out.append((frame, data))
    return out


def _inheritance_census(frames):
    """(blocks, treeless literal blocks, blocks with a repeat-mode sequence table) over the frames' block headers."""
    blocks = treeless = repeat = 0
    for frame, _ in frames:
        fhd = frame[4]
        pos = 5 + (0 if fhd & 0x20 else 1) + [0, 1, 2, 4][fhd & 3] + ([0, 2, 4, 8][fhd >> 6] or (1 if fhd & 0x20 else 0))
        while True:
            h = int.from_bytes(frame[pos:pos + 3], "little"); pos += 3
            last, typ, size = h & 1, (h >> 1) & 3, h >> 3
            if typ == 2:
                blocks += 1
                b = frame[pos:pos + size]
                lt, lhl = b[0] & 3, (b[0] >> 2) & 3
                if lt >= 2:
                    treeless += lt == 3
                    lhs = [3, 3, 4, 5][lhl]
                    w = int.from_bytes(b[:5], "little")
                    csz = [(w >> 14) & 0x3FF, (w >> 14) & 0x3FF, (w >> 18) & 0x3FFF, (w >> 22) & 0x3FFFF][lhl]
                    sp = lhs + csz
                else:
                    lhs = [1, 2, 1, 3][lhl]
                    lsz = [b[0] >> 3, int.from_bytes(b[:2], "little") >> 4, b[0] >> 3, int.from_bytes(b[:3], "little") >> 4][lhl]
                    sp = lhs + (lsz if lt == 0 else 1)
                n = b[sp]
                if n:
                    sp += 1 if n < 128 else (3 if n == 255 else 2)
                    m = b[sp]
                    repeat += any(((m >> s) & 3) == 3 for s in (6, 4, 2))
            pos += 1 if typ == 1 else size
            if last:
                break
    return blocks, treeless, repeat


def test_multi_block_frames_take_the_block_parallel_replay_and_match_the_oracle(hostsim, oracle):
    frames = _big_frames(301, 36)
    blocks, treeless, repeat = _inheritance_census(frames)
    assert blocks > 100 and treeless > 5 and repeat > 5, (blocks, treeless, repeat)   # the inheritance cases are all present
    before = hostsim.par_frames()
    for frame, data in frames:
        n = len(data)
        for cap in (n, n + 13, n - 1, n // 2):
            ro, oo, _ = oracle.decompress(frame, cap)
            rh, oh = hostsim.decompress(frame, cap, oracle)
            assert ro == rh and oo == oh, (n, cap, hex(ro), hex(rh))
    assert hostsim.par_frames() - before >= 2 * len(frames)      # (RLE / raw-only payloads and short capacities stay frame-serial)
    # the same frames on the frame-serial replay: identical verdicts
    hostsim.set_par(0)
    try:
        base = hostsim.par_frames()
        for frame, data in frames[:8]:
            assert hostsim.decompress(frame, len(data), oracle)[0] == len(data)
        assert hostsim.par_frames() == base
    finally:
        hostsim.set_par(1)


def test_fuzzed_multi_block_frames_same_result_code_and_bytes(hostsim, oracle):
    rng = random.Random(78)
    frames = _big_frames(302, 24)
    before = hostsim.par_frames()
    checked = errors = 0
    for frame, data in frames:
        for _ in range(14):
            b = helpers.mutate(rng, frame)
            if rng.random() < 0.5:          # damage inside a later block rather than near the frame header
                i = rng.randrange(len(frame) // 3, len(frame)); bb = bytearray(frame); bb[i] ^= 1 << rng.randrange(8); b = bytes(bb)
            cap = max(0, len(data) + rng.choice([0, 0, 0, 5, -1, -100, 1000]))
            ro, oo, _ = oracle.decompress(b, cap)
            rh, oh = hostsim.decompress(b, cap, oracle)
            assert ro == rh and oo == oh, (len(b), cap, hex(ro), hex(rh))
            checked += 1; errors += helpers.is_err(ro)
    assert errors > checked // 3 and hostsim.par_frames() - before > checked // 3


def test_repeat_offsets_across_blocks_are_resolved_symbolically(hostsim, oracle):
    """Period-p data makes nearly every sequence a repeat-offset match, across block boundaries too: the first offsets of
    every block then come from the previous block's history (the symbolic records), including the 'rep0 - 1' code."""
    from tools import zstd_ref
    rng = random.Random(9)
    before = hostsim.par_frames()
    for t in range(12):
        period = rng.choice([5, 24, 24, 100, 333])
        unit = bytes(rng.randrange(256) for _ in range(period))
        body = bytearray(unit * (400000 // period + 1))[:400000]
        for _ in range(rng.choice([0, 40, 400])):     # sparse edits: literals between repeat matches
            body[rng.randrange(len(body))] = rng.randrange(256)
        data = bytes(body)
        frame = zstd_ref.compress(data, rng.choice([1, 3, 5, 19]), checksum=True)
        ro, oo, _ = oracle.decompress(frame, len(data))
        rh, oh = hostsim.decompress(frame, len(data), oracle)
        assert ro == rh == len(data) and oo == oh == data
    assert hostsim.par_frames() - before == 12


def test_multi_block_frames_with_dictionaries(hostsim, oracle):
    from tools import corpus, zstd_ref
    lib = hostsim.lib
    lib.hostsim_set_dict.argtypes = [ctypes.c_char_p, ctypes.c_uint32]
    run = _dict_oracle(oracle)
    rng = random.Random(23)
    log = corpus.log(5 << 20).tobytes()
    trained = zstd_ref.train_dict([log[i:i + 4096] for i in range(0, 2 << 20, 4096)], 32768)
    raw = log[:30000]
    before = hostsim.par_frames()
    try:
        for d in (trained, raw):
            lib.hostsim_set_dict(d, len(d))
            for k in range(10):
                n = rng.choice([150000, 300000, 600000])
                i = rng.randrange(2 << 20, (5 << 20) - n)
                data = log[i:i + n]
                frame = zstd_ref.compress_with_dict(data, d, rng.choice([1, 3, 5]), True, True)
                if k % 3 == 2:
                    frame = helpers.mutate(rng, frame)
                cap = len(data) + rng.choice([0, 9, -1])
                want, out = run(frame, cap, d)
                r, got = hostsim.decompress(frame, cap, oracle)
                assert r == want, (k, hex(r), hex(want))
                if out is not None:
                    assert got == out
    finally:
        lib.hostsim_set_dict(None, 0)
    assert hostsim.par_frames() - before >= 10
