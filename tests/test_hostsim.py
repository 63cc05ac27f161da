"""CPU replay of the GPU decoder's entropy stages (same source, compiled by g++) against the oracle.

This is how the kernels' parsers, table builders and bit readers are validated in the GPU-less build container;
the warp-cooperative execute stage and the checksum kernel are covered by the -m gpu tests."""
import random

from tests import helpers


def test_golden_vectors(hostsim, oracle):
    for name, frame, raw in helpers.golden_vectors():
        r, out = hostsim.decompress(frame, len(raw), oracle)
        assert r == len(raw) and out == raw, name


def test_valid_frames_match_oracle(hostsim, oracle):
    for frame, data in helpers.make_frames(201, 200):
        n = len(data)
        for cap in (n, n + 13, max(0, n - 1)):
            ro, oo, _ = oracle.decompress(frame, cap)
            rh, oh = hostsim.decompress(frame, cap, oracle)
            assert ro == rh and oo == oh, (n, cap, hex(ro), hex(rh))


def test_fuzzed_frames_same_result_code(hostsim, oracle):
    rng = random.Random(77)
    frames = helpers.make_frames(202, 120)
    for frame, data in frames:
        if len(frame) < 12:
            continue
        for _ in range(10):
            b = helpers.mutate(rng, frame)
            cap = max(0, len(data) + rng.choice([0, 0, 0, 5, -1, -100, 1000]))
            ro, oo, _ = oracle.decompress(b, cap)
            rh, oh = hostsim.decompress(b, cap, oracle)
            # same result code (error codes included) and bytes: the stages reproduce the reference's order of checks
            # and what its 4-byte bit container yields when a sequence reads below the stream start
            assert ro == rh and oo == oh, (b.hex()[:80], cap, hex(ro), hex(rh))


def test_items_with_several_data_frames(hostsim, oracle):
    """One pass per data frame (the loop api.cu runs around the kernels) against the reference's multi-frame loop."""
    for blob, cap in helpers.multi_frame_items():
        ro, oo, _ = oracle.decompress(blob, cap)
        rh, oh = hostsim.decompress(blob, cap, oracle)
        assert ro == rh and oo == oh, (len(blob), cap, hex(ro), hex(rh))


def test_huffman_table_log_12_is_folded_correctly(hostsim, oracle):
    """The kernels keep 2^11 cells per Huffman table and fold a log-12 table into it (zb_format.cuh huf_fill_table)."""
    for four in (True, False):
        for n in (300, 700, 1001):
            for prefix in (0, 50):
                frame, plain = helpers.huf12_frame(n, four, seed=n + prefix, raw_prefix=prefix)
                ro, oo, _ = oracle.decompress(frame, len(plain))
                assert ro == len(plain) and oo == plain, "crafted frame is not valid for the oracle"
                rh, oh = hostsim.decompress(frame, len(plain), oracle)
                assert rh == ro and oh == plain
                for cut in (1, 3):                    # truncations: same verdicts
                    bad = frame[:-cut]
                    assert hostsim.decompress(bad, len(plain), oracle)[0] == oracle.decompress(bad, len(plain))[0]
                assert hostsim.decompress(frame, 10, oracle)[0] == oracle.decompress(frame, 10)[0]    # dry validation path


def test_every_frame_header_shape(hostsim, oracle):
    from tools import zstd_ref
    items = helpers.header_variant_frames()
    assert len(items) > 600
    n_ok = 0
    for frame, cap in items:
        ro, oo, _ = oracle.decompress(frame, cap)
        rh, oh = hostsim.decompress(frame, cap, oracle)
        assert ro == rh and oo == oh, (frame[:16].hex(), cap, hex(ro), hex(rh))
        if not helpers.is_err(ro):
            n_ok += 1
            z = zstd_ref.decompress(frame, cap)        # libzstd refuses some legal shapes (windows > 2^27): only compare when it answers
            assert z is None or z == oo
    assert n_ok > 100


def test_sequence_count_encodings(hostsim, oracle):
    frames = helpers.sequence_count_frames()
    counts = [n for _, _, n in frames]
    assert counts[0] < 128 <= counts[1] < 0x7F00 <= counts[2], counts
    for f, data, _ in frames:
        rh, oh = hostsim.decompress(f, len(data), oracle)
        assert rh == len(data) and oh == data


def test_rle_literals_and_rle_mode_tables(hostsim, oracle):
    """Modes no encoder at hand emits (hand-made frames, accepted by libzstd): RLE literals with 1/2/3-byte headers,
    LL/OF/ML tables in RLE mode; plus mutations of those frames (same result codes)."""
    from tools import zstd_ref
    rng = random.Random(12)
    for nseq, tail in ((1, 0), (5, 3), (28, 3), (100, 3), (127, 0), (128, 5), (4000, 200), (5000, 0), (14000, 7)):
        f, p = helpers.rle_modes_frame(nseq, tail, seed=nseq)
        assert zstd_ref.decompress(f, len(p)) == p
        for frame, cap in [(f, len(p)), (f, len(p) - 1), (f[:-1], len(p))] + [(helpers.mutate(rng, f), len(p)) for _ in range(12)]:
            ro, oo, _ = oracle.decompress(frame, cap)
            rh, oh = hostsim.decompress(frame, cap, oracle)
            assert ro == rh and oo == oh, (nseq, tail, frame.hex()[:60], hex(ro), hex(rh))


def test_repeat_mode_tables(hostsim, oracle):
    import collections
    frames = helpers.repeat_mode_frames()
    modes = collections.Counter()
    for f, _ in frames:
        modes.update(helpers.frame_modes(f))
    assert modes[("LL", 3)] and modes[("OF", 3)] and modes[("ML", 3)], modes       # the probe found repeat mode for every kind
    for f, data in frames:
        for cap in (len(data), len(data) - 1):
            ro, oo, _ = oracle.decompress(f, cap)
            rh, oh = hostsim.decompress(f, cap, oracle)
            assert ro == rh and oo == oh
        assert oracle.decompress(f, len(data))[1] == data


def test_maximum_match_length_records(hostsim, oracle):
    """A block that is one repeat-offset match of 131 072 bytes (matchLength needs 18 bits in the sequence record), and
    frames of long runs in general."""
    from tools import zstd_ref
    for data in (bytes(1 << 20), b"ab" * (300 << 10), bytes(131072 + 7)):
        frame = zstd_ref.compress(data, 3, checksum=True)
        want, out, _ = oracle.decompress(frame, len(data))
        assert want == len(data) and out == data
        r, got = hostsim.decompress(frame, len(data), oracle)
        assert r == want and got == data


def test_large_window_frames_use_the_look_ahead_sequence_loop(hostsim, oracle):
    """Windows above 16 MiB: same bytes, and on damaged frames the result code of the loop that executes four
    sequences behind the decoder (DecompressSequencesLong, ZStdDecompress.cs:1708-1787)."""
    rng = random.Random(31)
    n_long_fail = 0
    for frame, data in helpers.large_window_frames():
        ro, oo, _ = oracle.decompress(frame, len(data))
        assert ro == len(data) and oo == data
        rh, oh = hostsim.decompress(frame, len(data), oracle)
        assert rh == ro and oh == oo
        for _ in range(30):
            b = bytearray(helpers.mutate(rng, frame))
            if len(b) > 5:
                b[:6] = frame[:6]                     # keep the header: the mutations are for the blocks
            cap = max(0, len(data) + rng.choice([0, 0, 5, -1, -7, -40, -len(data) // 3, 1000]))
            ro, oo, _ = oracle.decompress(bytes(b), cap)
            rh, oh = hostsim.decompress(bytes(b), cap, oracle)
            assert ro == rh and oo == oh, (bytes(b).hex()[:80], cap, hex(ro), hex(rh))
            n_long_fail += helpers.is_err(ro)
    assert n_long_fail > 300


def test_both_root_table_sizes_of_the_huffman_stage(hostsim, oracle):
    """The Huffman kernels keep a root table of 2^9 (frames of few literals) or 2^11 cells in shared memory and look longer
    codes up in the full table (zb_format.cuh huf_fill_root, zb_decode.cuh huf_cell).  The replay defaults to 2^9, which sends
    every code of 10+ bits through the long path; this runs valid, damaged and log-12 frames with 2^11 as well."""
    rng = random.Random(5)
    frames = helpers.make_frames(203, 80)
    try:
        for root_log in (11, 9):
            hostsim.set_huf_root(root_log)
            for frame, data in frames:
                ro, oo, _ = oracle.decompress(frame, len(data))
                rh, oh = hostsim.decompress(frame, len(data), oracle)
                assert ro == rh and oo == oh, (root_log, len(data))
                if len(frame) > 12:
                    b = helpers.mutate(rng, frame)
                    ro, oo, _ = oracle.decompress(b, len(data))
                    rh, oh = hostsim.decompress(b, len(data), oracle)
                    assert ro == rh and oo == oh, (root_log, "mutated", len(data))
            for four in (True, False):
                frame, plain = helpers.huf12_frame(700, four, seed=9, raw_prefix=0)
                assert hostsim.decompress(frame, len(plain), oracle) == (len(plain), plain)
    finally:
        hostsim.set_huf_root(9)


def test_small_sequence_tables_hand_larger_frames_over(hostsim, oracle):
    """k_seq runs as three instantiations with room for table logs 6 / 8 / the format's maximum; a frame whose tables do not
    fit the one k_parse chose is handed to the full-size one (zb_format.cuh read_seq_table capLog, SeqEmitter::defer).  The
    replay makes every frame start in the small instantiation: same results, and the hand-over did happen."""
    rng = random.Random(6)
    frames = helpers.make_frames(204, 90)
    before = hostsim.lib.hostsim_deferred()
    try:
        for cap_log in (6, 8):
            hostsim.lib.hostsim_set_seq_cap(cap_log)
            for frame, data in frames:
                ro, oo, _ = oracle.decompress(frame, len(data))
                rh, oh = hostsim.decompress(frame, len(data), oracle)
                assert ro == rh and oo == oh, (cap_log, len(data))
                if len(frame) > 12:
                    b = helpers.mutate(rng, frame)
                    ro, oo, _ = oracle.decompress(b, len(data))
                    rh, oh = hostsim.decompress(b, len(data), oracle)
                    assert ro == rh and oo == oh, (cap_log, "mutated", len(data))
    finally:
        hostsim.lib.hostsim_set_seq_cap(9)
    assert hostsim.lib.hostsim_deferred() - before > 20


def test_sequence_class_guesses_hold_for_libzstd_frames(hostsim, oracle):
    """k_parse's rule (zb_format.cuh first_block_classes): at most 512 sequences in the first block -> tables of at most
    2^6 / 2^6 / 2^7 cells, at most 2 048 -> 2^8.  The rule is a performance guess, not a correctness condition; this test pins
    that it IS right for libzstd frames of the benchmark's shapes (log text, tick records, mixed; 1-64 KiB; levels 1-19) -
    no frame is handed over - that the small classes really get used, and that results equal the oracle's either way."""
    from tools import corpus, zstd_ref
    lib = hostsim.lib
    base = [lib.hostsim_class_frames(c) for c in range(3)]
    deferred = lib.hostsim_deferred()
    lib.hostsim_set_seq_cap(0)
    try:
        for kind, total in (("log", 1 << 20), ("tick", 1 << 20), ("mixed", 1 << 19)):
            raw = corpus.make(kind, total).tobytes()
            for chunk, level in ((1024, 3), (4096, 1), (4096, 3), (4096, 19), (16384, 3), (16384, 9), (65536, 3), (65536, 19)):
                for pos in range(0, min(total, 6 * chunk), chunk):
                    data = raw[pos:pos + chunk]
                    frame = zstd_ref.compress(data, level, checksum=True)
                    ro, oo, _ = oracle.decompress(frame, len(data))
                    assert (ro, oo) == (len(data), data)
                    assert hostsim.decompress(frame, len(data), oracle) == (len(data), data), (kind, chunk, level, pos)
    finally:
        lib.hostsim_set_seq_cap(9)
    used = [lib.hostsim_class_frames(c) - base[c] for c in range(3)]
    assert used[1] > 20 and used[2] > 20 and used[0] > 20, used       # A, B and the full-size class all occur
    assert lib.hostsim_deferred() == deferred, "a libzstd frame needed larger tables than its class provides"
