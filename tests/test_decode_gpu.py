"""Parity tests proper: the CUDA decode path, called through the C ABI, against the oracle.

Bit-exact bar: decoded bytes and result codes (error codes included) equal the oracle's on the same inputs
(integer/byte work, no tolerance) — also for corrupted frames, where the reference's 4-byte bit container reads
garbage below the stream start before it notices (zb_decode.cuh "over-read emulation")."""
import os
import random
import sys

import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu


def _gpu_decode(ctx, frames_and_caps):
    srcs = [f for f, _ in frames_and_caps]
    dsts = [np.zeros(max(c, 1), dtype=np.uint8)[:c] for _, c in frames_and_caps]
    res = ctx.decompress_batch(srcs, dsts)
    return res, dsts


def test_reference_golden_vectors_through_public_api(gpu_ctx):
    # same assertions as csharp/test/TestDecompress.cs:91-98, through the mirror of ZStdDecompress
    import zstandard_b200 as zb
    for name, frame, raw in helpers.golden_vectors():
        n = zb.ZStdDecompress.GetDecompressedSize(frame)
        assert n == len(raw), name
        out = np.zeros(n, dtype=np.uint8)
        r = zb.ZStdDecompress.Decompress(out, frame)
        assert r == n and out.tobytes() == raw, name
        out2 = np.zeros(n + 100, dtype=np.uint8)
        assert zb.ZStdDecompress.Decompress(out2, n + 100, frame, len(frame)) == n and out2[:n].tobytes() == raw


def test_batch_matches_oracle_on_valid_frames(gpu_ctx, oracle):
    frames = helpers.make_frames(301, 300)
    items = []
    for frame, data in frames:
        n = len(data)
        for cap in (n, n + 13, max(0, n - 1)):
            items.append((frame, cap))
    res, dsts = _gpu_decode(gpu_ctx, items)
    for (frame, cap), r, d in zip(items, res, dsts):
        ro, oo, _ = oracle.decompress(frame, cap)
        assert int(r) == ro, (len(frame), cap, hex(int(r)), hex(ro))
        if not helpers.is_err(ro):
            assert d[:ro].tobytes() == oo


def test_edge_cases(gpu_ctx, oracle):
    (f1, d1), (f2, d2) = helpers.make_frames(9, 2, sizes=[3000, 70000])
    bad_sum = bytearray(helpers.make_frames(5, 8, sizes=[5000], levels=[3])[1][0])
    cases = [
        (b"", 10),                                                    # empty input -> 0
        (b"\x28\xb5", 10),                                            # shorter than a frame prefix
        (helpers.skippable(b"hello"), 10),                            # only a skippable frame -> 0
        (helpers.skippable(b"hello") + f1, len(d1)),                  # skippable then frame
        (f1 + helpers.skippable(b"tail" * 5), len(d1)),               # frame then skippable
        (f1 + b"\x00", len(d1)),                                      # 1 trailing byte -> srcSize_wrong
        (f1 + b"\x00" * 8, len(d1)),                                  # trailing garbage -> prefix_unknown
        (b"\x00" * 8 + f1, len(d1)),                                  # leading garbage -> prefix_unknown
        (f1[:len(f1) // 2], len(d1)),                                 # truncated
        (f2, len(d2) - 1), (f2, 0),                                   # dst too small
    ]
    for frame, data in helpers.make_frames(5, 8, sizes=[5000]):
        f = bytearray(frame); f[4] |= 0x08
        cases.append((bytes(f), len(data)))                           # reserved bit
    res, dsts = _gpu_decode(gpu_ctx, cases)
    for (frame, cap), r, d in zip(cases, res, dsts):
        ro, oo, _ = oracle.decompress(frame, cap)
        assert int(r) == ro, (frame[:12].hex(), cap, hex(int(r)), hex(ro))
        if not helpers.is_err(ro):
            assert d[:ro].tobytes() == oo
    del bad_sum


def test_huffman_table_log_12(gpu_ctx, oracle):
    """Log-12 Huffman tables (legal, never emitted by encoders) are folded into the kernels' 2^11-cell tables."""
    items = []
    for four in (True, False):
        for n in (300, 700, 1001):
            for prefix in (0, 50):
                frame, plain = helpers.huf12_frame(n, four, seed=n + prefix, raw_prefix=prefix)
                items += [(frame, len(plain)), (frame[:-1], len(plain)), (frame, 10)]
    ordinary = [(f, len(d)) for f, d in helpers.make_frames(306, 20, sizes=[5000, 70000])]
    mixed = []
    for k, it in enumerate(items):                     # folded and ordinary tables side by side in one CTA
        mixed += [it, ordinary[k % len(ordinary)]]
    res, dsts = _gpu_decode(gpu_ctx, mixed)
    for (frame, cap), r, d in zip(mixed, res, dsts):
        ro, oo, _ = oracle.decompress(frame, cap)
        assert int(r) == ro, (len(frame), cap, hex(ro), hex(int(r)))
        if not helpers.is_err(ro):
            assert d[:ro].tobytes() == oo


def test_checksum_mismatch_is_reported(gpu_ctx, oracle):
    from tools import zstd_ref
    rng = random.Random(4)
    items = []
    for n in (1, 31, 32, 33, 1000, 65536, 70001):
        data = helpers.sample_payload(rng, 0, n)
        f = bytearray(zstd_ref.compress(data, 3, checksum=True))
        items.append((bytes(f), n))
        f[-1] ^= 0x5A
        items.append((bytes(f), n))
    res, _ = _gpu_decode(gpu_ctx, items)
    for k, ((frame, cap), r) in enumerate(zip(items, res)):
        ro, _, _ = oracle.decompress(frame, cap)
        assert int(r) == ro
        assert helpers.is_err(ro) == bool(k & 1)


def test_fuzzed_frames_same_result_code_and_bytes(gpu_ctx, oracle):
    rng = random.Random(78)
    frames = helpers.make_frames(302, 150)
    items = []
    for frame, data in frames:
        if len(frame) < 12:
            continue
        for _ in range(16):
            b = helpers.mutate(rng, frame)
            if rng.random() < 0.3 and len(b) > 4:
                b = helpers.mutate(rng, b)
            items.append((b, max(0, len(data) + rng.choice([0, 0, 0, 5, -1, -100, 1000]))))
    res, dsts = _gpu_decode(gpu_ctx, items)
    n_ok = 0
    for (frame, cap), r, d in zip(items, res, dsts):
        ro, oo, _ = oracle.decompress(frame, cap)
        assert int(r) == ro, (frame.hex()[:80], cap, hex(ro), hex(int(r)))
        if not helpers.is_err(ro):
            assert d[:ro].tobytes() == oo
            n_ok += 1
    assert n_ok > 20     # the mutations leave some frames decodable (stored blocks, skipped fields)


def test_items_with_several_data_frames(gpu_ctx, oracle):
    """DecompressMultiFrame (ZStdDecompress.cs:2096-2160): data frames and skippable frames concatenated in one item
    decode back to back into the item's destination; one pass of the pipeline per data frame."""
    items = helpers.multi_frame_items()
    singles = [(f, len(d)) for f, d in helpers.make_frames(305, 50)]
    mixed = []
    for k, it in enumerate(items):          # multi-frame items scattered among ordinary ones
        mixed.append(it)
        mixed.append(singles[k % len(singles)])
    res, dsts = _gpu_decode(gpu_ctx, mixed)
    n_multi_ok = 0
    for k, ((blob, cap), r, d) in enumerate(zip(mixed, res, dsts)):
        ro, oo, _ = oracle.decompress(blob, cap)
        assert int(r) == ro, (k, len(blob), cap, hex(ro), hex(int(r)))
        if not helpers.is_err(ro):
            assert d[:ro].tobytes() == oo
            n_multi_ok += (k % 2 == 0)
    assert n_multi_ok >= 40
    # many frames in one item (one pass each)
    from tools import zstd_ref
    parts = [helpers.sample_payload(random.Random(k), k % 6, 300 + 37 * k) for k in range(40)]
    blob = b"".join(zstd_ref.compress(p, 3, checksum=bool(k & 1)) for k, p in enumerate(parts))
    raw = b"".join(parts)
    res, dsts = _gpu_decode(gpu_ctx, [(blob, len(raw)), (blob, len(raw) - 1)])
    assert int(res[0]) == len(raw) and dsts[0].tobytes() == raw
    assert int(res[1]) == oracle.decompress(blob, len(raw) - 1)[0]


def test_one_bad_item_does_not_poison_the_batch(gpu_ctx, oracle):
    frames = helpers.make_frames(303, 40, sizes=[4096, 65536])
    items = [(f, len(d)) for f, d in frames]
    items[7] = (b"\xff" * 100, 100)
    items[19] = (items[19][0][:20], items[19][1])
    res, dsts = _gpu_decode(gpu_ctx, items)
    for k, ((frame, cap), r, d) in enumerate(zip(items, res, dsts)):
        ro, oo, _ = oracle.decompress(frame, cap)
        assert int(r) == ro
        if k not in (7, 19):
            assert not helpers.is_err(ro) and d[:ro].tobytes() == oo


def test_device_pointer_api(gpu_ctx, oracle):
    import torch
    frames = helpers.make_frames(304, 64, sizes=[100, 4096, 65536, 131072])
    blob = b"".join(f for f, _ in frames)
    soff = np.cumsum([0] + [len(f) for f, _ in frames])[:-1].astype(np.uint64)
    ssz = np.array([len(f) for f, _ in frames], dtype=np.uint32)
    cap = np.array([len(d) for _, d in frames], dtype=np.uint32)
    doff = np.cumsum([0] + [(int(c) + 15) // 16 * 16 for c in cap])[:-1].astype(np.uint64)
    dev = torch.device("cuda:0")
    t_src = torch.frombuffer(bytearray(blob) + bytearray(16), dtype=torch.uint8).to(dev)
    t_dst = torch.zeros(int(doff[-1]) + int(cap[-1]) + 64, dtype=torch.uint8, device=dev)
    t_soff = torch.from_numpy(soff.view(np.int64)).to(dev)
    t_doff = torch.from_numpy(doff.view(np.int64)).to(dev)
    t_ssz = torch.from_numpy(ssz.view(np.int32)).to(dev)
    t_cap = torch.from_numpy(cap.view(np.int32)).to(dev)
    t_res = torch.zeros(len(frames), dtype=torch.int32, device=dev)
    ts = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    stream = ts.cuda_stream
    gpu_ctx.decompress_batch_device(t_src.data_ptr(), t_soff.data_ptr(), t_ssz.data_ptr(), t_dst.data_ptr(), t_doff.data_ptr(),
                                    t_cap.data_ptr(), t_res.data_ptr(), len(frames), stream=stream)
    torch.cuda.synchronize()
    res = t_res.cpu().numpy().view(np.uint32)
    out = t_dst.cpu().numpy()
    for k, (frame, data) in enumerate(frames):
        assert int(res[k]) == len(data)
        assert out[int(doff[k]):int(doff[k]) + len(data)].tobytes() == data
    # the same frames as ONE item (64 data frames back to back): 64 passes on the device-pointer path
    total = sum(len(d) for _, d in frames)
    one_soff = torch.zeros(1, dtype=torch.int64, device=dev); one_doff = torch.zeros(1, dtype=torch.int64, device=dev)
    one_ssz = torch.tensor([len(blob)], dtype=torch.int32, device=dev); one_cap = torch.tensor([total], dtype=torch.int32, device=dev)
    one_res = torch.zeros(1, dtype=torch.int32, device=dev)
    t_dst.zero_()
    gpu_ctx.decompress_batch_device(t_src.data_ptr(), one_soff.data_ptr(), one_ssz.data_ptr(), t_dst.data_ptr(), one_doff.data_ptr(),
                                    one_cap.data_ptr(), one_res.data_ptr(), 1, stream=stream)
    torch.cuda.synchronize()
    assert int(one_res.cpu().numpy().view(np.uint32)[0]) == total
    assert t_dst[:total].cpu().numpy().tobytes() == b"".join(d for _, d in frames)
    ms = gpu_ctx.decompress_batch_device_timed(t_src.data_ptr(), t_soff.data_ptr(), t_ssz.data_ptr(), t_dst.data_ptr(), t_doff.data_ptr(),
                                               t_cap.data_ptr(), t_res.data_ptr(), len(frames), stream=stream)
    assert set(ms) == {"k_parse", "k_huf", "k_seq", "k_exec", "k_xxh"} and all(v >= 0 for v in ms.values())


@pytest.mark.parametrize("kind,chunk", [("log", 65536), ("tick", 65536), ("mixed", 65536), ("tick", 4096), ("log", 1 << 20)])
def test_full_size_round_trip_properties(gpu_ctx, kind, chunk):
    """Size-independent properties at bench-like sizes: every frame decodes to exactly its chunk of the corpus
    (libzstd frames carry XXH64 checksums, which the GPU path verifies itself) — 32 MiB per case."""
    from tools import corpus, zstd_ref
    total = 32 << 20
    raw = corpus.make(kind, total)
    blob, off = zstd_ref.compress_chunks(raw, chunk, level=3, checksum=True)
    n = len(off) - 1
    srcs = [blob[int(off[i]):int(off[i + 1])] for i in range(n)]
    out = np.zeros(total, dtype=np.uint8)
    dsts = [out[i * chunk:min(total, (i + 1) * chunk)] for i in range(n)]
    res = gpu_ctx.decompress_batch(srcs, dsts)
    want = np.array([d.size for d in dsts], dtype=np.uint32)
    assert (res == want).all(), [hex(int(r)) for r in res[res != want][:5]]
    assert (out == raw).all()


def test_one_context_over_two_devices(oracle):
    """one context fans a batch out over its devices by bytes, no collective; results are those of a
    single device.  Skipped on single-GPU boxes."""
    import torch
    import zstandard_b200 as zb
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx = zb.Context(devices=[0, 1], max_batch_bytes=64 << 20)
    try:
        frames = helpers.make_frames(307, 150)
        items = [(f, len(d)) for f, d in frames] + [(b"\xff" * 50, 10)]
        res, dsts = _gpu_decode(ctx, items)
        for (frame, cap), r, d in zip(items, res, dsts):
            ro, oo, _ = oracle.decompress(frame, cap)
            assert int(r) == ro
            if not helpers.is_err(ro):
                assert d[:ro].tobytes() == oo
        payloads = [d for _, d in frames if len(d)]
        outs = [np.zeros(zb.ZStdCompress.CompressBound(len(p)), dtype=np.uint8) for p in payloads]
        cres = ctx.compress_batch(payloads, outs, level=3, checksum=True)
        for p, r, o in zip(payloads, cres, outs):
            assert not helpers.is_err(int(r))
            ro, oo, _ = oracle.decompress(o[:int(r)].tobytes(), len(p))
            assert ro == len(p) and oo == p
        assert ctx.device_count == 2
    finally:
        ctx.close()


def test_batches_larger_than_the_context_arena_are_cut_into_sub_batches():
    """A small context (4 MiB of arena) takes a 24 MiB batch in several sub-batches; an item whose content cannot fit at
    all is a batch-level error (`zstdb200_last_error`), not a crash."""
    import zstandard_b200 as zb
    from tools import corpus, zstd_ref
    ctx = zb.Context(max_batch_bytes=4 << 20)
    try:
        total, chunk = 24 << 20, 65536
        raw = corpus.make("mixed", total)
        blob, off = zstd_ref.compress_chunks(raw, chunk, level=3, checksum=True)
        n = len(off) - 1
        srcs = [blob[int(off[i]):int(off[i + 1])] for i in range(n)]
        out = np.zeros(total, dtype=np.uint8)
        dsts = [out[i * chunk:(i + 1) * chunk] for i in range(n)]
        res = ctx.decompress_batch(srcs, dsts)
        assert (res == chunk).all() and (out == raw).all()
        chunks = [raw[i * chunk:(i + 1) * chunk] for i in range(n)]
        outs = [np.zeros(zb.ZStdCompress.CompressBound(chunk), dtype=np.uint8) for _ in range(n)]
        cres = ctx.compress_batch(chunks, outs, level=1, checksum=True)
        for k in range(0, n, 11):
            assert zstd_ref.decompress(outs[k][:int(cres[k])].tobytes(), chunk) == chunks[k].tobytes()
        # a destination larger than the arena is fine as long as the content fits (capacity clamp, api.cu) ...
        big = np.zeros(8 << 20, dtype=np.uint8)
        assert int(ctx.decompress_batch([srcs[0]], [big])[0]) == chunk and (big[:chunk] == raw[:chunk]).all()
        # ... an item whose content cannot fit at all is the batch-level error
        huge = zstd_ref.compress(raw[:8 << 20].tobytes(), 1, checksum=False)
        with pytest.raises(RuntimeError):
            ctx.decompress_batch([huge], [big])
    finally:
        ctx.close()


def test_multi_megabyte_frames_with_large_windows(gpu_ctx, oracle):
    """multi-block frames with windows far beyond one block (libzstd level 19 on 8 MiB: 8 MiB window,
    repeat-mode tables, cross-block repeat offsets) decode bit-exact; the encoder's multi-block frames are accepted
    by libzstd and by the oracle."""
    import zstandard_b200 as zb
    from tools import corpus, zstd_ref
    items = []
    for kind, n, lvl in (("log", 8 << 20, 19), ("tick", 6 << 20, 12), ("mixed", 5 << 20, 3), ("log", 3 << 20, 1)):
        raw = corpus.make(kind, n).tobytes()
        items.append((zstd_ref.compress(raw, lvl, checksum=True), raw))
    dsts = [np.zeros(len(r), dtype=np.uint8) for _, r in items]
    res = gpu_ctx.decompress_batch([f for f, _ in items], dsts)
    for (f, raw), r, d in zip(items, res, dsts):
        assert int(r) == len(raw) and d.tobytes() == raw
    ro, oo, _ = oracle.decompress(items[0][0], len(items[0][1]))
    assert ro == len(items[0][1]) and oo == items[0][1]
    outs = [np.zeros(zb.ZStdCompress.CompressBound(len(r)), dtype=np.uint8) for _, r in items]
    cres = gpu_ctx.compress_batch([r for _, r in items], outs, level=3)
    for (f, raw), r, o in zip(items, cres, outs):
        frame = o[:int(r)].tobytes()
        assert zstd_ref.decompress(frame, len(raw)) == raw
    ro, oo, _ = oracle.decompress(outs[3][:int(cres[3])].tobytes(), len(items[3][1]))
    assert ro == len(items[3][1]) and oo == items[3][1]


def test_every_frame_header_shape(gpu_ctx, oracle):
    """ZSTD_getFrameHeader_advanced / frameHeaderSize (ZStdDecompress.cs:389-499): all field-size combinations, and
    zstdb200_get_decompressed_size against the oracle's GetDecompressedSize for each."""
    import zstandard_b200 as zb
    items = helpers.header_variant_frames()
    res, dsts = _gpu_decode(gpu_ctx, items)
    for (frame, cap), r, d in zip(items, res, dsts):
        ro, oo, _ = oracle.decompress(frame, cap)
        assert int(r) == ro, (frame[:16].hex(), cap, hex(ro), hex(int(r)))
        if not helpers.is_err(ro):
            assert d[:ro].tobytes() == oo
        assert zb.ZStdDecompress.GetDecompressedSize(frame) == oracle.get_decompressed_size(frame)


def test_sequence_count_encodings(gpu_ctx):
    """nbSeq in one byte, two bytes and the 255 + u16 form (>= 0x7F00 sequences in one block), DecodeSeqHeaders :1121-1136."""
    frames = helpers.sequence_count_frames()
    counts = [n for _, _, n in frames]
    assert counts[0] < 128 <= counts[1] < 0x7F00 <= counts[2], counts
    res, dsts = _gpu_decode(gpu_ctx, [(f, len(d)) for f, d, _ in frames])
    for (f, data, _), r, d in zip(frames, res, dsts):
        assert int(r) == len(data) and d.tobytes() == data


def test_rle_literals_and_rle_mode_tables(gpu_ctx, oracle):
    """RLE literals sections (1/2/3-byte headers) and LL/OF/ML tables in RLE mode: hand-made frames and their mutations."""
    rng = random.Random(12)
    items = []
    for nseq, tail in ((1, 0), (5, 3), (28, 3), (100, 3), (127, 0), (128, 5), (4000, 200), (5000, 0), (14000, 7)):
        f, p = helpers.rle_modes_frame(nseq, tail, seed=nseq)
        items += [(f, len(p)), (f, len(p) - 1), (f[:-1], len(p))] + [(helpers.mutate(rng, f), len(p)) for _ in range(12)]
    res, dsts = _gpu_decode(gpu_ctx, items)
    for (frame, cap), r, d in zip(items, res, dsts):
        ro, oo, _ = oracle.decompress(frame, cap)
        assert int(r) == ro, (frame.hex()[:60], cap, hex(ro), hex(int(r)))
        if not helpers.is_err(ro):
            assert d[:ro].tobytes() == oo


def test_repeat_mode_tables(gpu_ctx, oracle):
    frames = helpers.repeat_mode_frames()
    res, dsts = _gpu_decode(gpu_ctx, [(f, len(d)) for f, d in frames] + [(f, len(d) - 1) for f, d in frames])
    for k, (f, data) in enumerate(frames):
        assert int(res[k]) == len(data) and dsts[k].tobytes() == data
        assert int(res[k + len(frames)]) == oracle.decompress(f, len(data) - 1)[0]


def test_long_offset_regime_window_above_32_mib(gpu_ctx, oracle):
    """Offset codes >= 25 in a frame whose window exceeds 2^25 (ZStdDecompress.cs:1494-1501 split read, :1898-1905 dispatch):
    2 MiB of random bytes, 34 MiB of zeros, the same 2 MiB again — libzstd's long-distance matcher codes the second copy
    as matches 36 MiB back."""
    from tools import zstd_ref
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, 2 << 20, dtype=np.uint8).tobytes()
    data = a + bytes(34 << 20) + a
    c = zstd_ref._Z.ZSTD_createCCtx()
    try:
        for param, value in ((zstd_ref.ZSTD_c_compressionLevel, 3), (zstd_ref.ZSTD_c_checksumFlag, 1), (zstd_ref.ZSTD_c_windowLog, 27), (160, 1)):
            zstd_ref._Z.ZSTD_CCtx_setParameter(c, param, value)          # 160 = ZSTD_c_enableLongDistanceMatching
        import ctypes
        cap = zstd_ref.compress_bound(len(data))
        buf = ctypes.create_string_buffer(cap)
        n = zstd_ref._Z.ZSTD_compress2(c, buf, cap, data, len(data))
        assert not zstd_ref._Z.ZSTD_isError(n)
        frame = buf.raw[:n]
    finally:
        zstd_ref._Z.ZSTD_freeCCtx(c)
    assert len(frame) < (3 << 20)                  # the second copy was found: the frame holds the random block once
    assert (frame[4] >> 5) & 1 and len(data) > (1 << 25)   # single-segment frame: window = content size, above 2^25
    ro, oo, _ = oracle.decompress(frame, len(data))
    assert ro == len(data) and oo == data
    items = [(frame, len(data)), (frame, len(data) - 1), (frame[:-5], len(data))]
    res, dsts = _gpu_decode(gpu_ctx, items)
    for (f, cap), r, d in zip(items, res, dsts):
        want, out, _ = oracle.decompress(f, cap)
        assert int(r) == want, (hex(int(r)), hex(want))
        if not helpers.is_err(want):
            assert d[:want].tobytes() == out


def test_large_window_frames_use_the_look_ahead_sequence_loop(gpu_ctx, oracle):
    """Declared windows above 16 MiB: the reference runs blocks with enough long-offset cells through the loop that
    executes four sequences behind the decoder (ZStdDecompress.cs:1898-1905, :1708-1787) — same bytes, but a damaged
    block reports what THAT loop would have reached."""
    rng = random.Random(32)
    items = []
    for frame, data in helpers.large_window_frames(seed=6, count=80):
        items.append((frame, len(data)))
        for _ in range(30):
            b = bytearray(helpers.mutate(rng, frame))
            if len(b) > 5:
                b[:6] = frame[:6]
            items.append((bytes(b), max(0, len(data) + rng.choice([0, 0, 5, -1, -7, -40, -len(data) // 3, 1000]))))
    res, dsts = _gpu_decode(gpu_ctx, items)
    n_err = 0
    for (frame, cap), r, d in zip(items, res, dsts):
        ro, oo, _ = oracle.decompress(frame, cap)
        assert int(r) == ro, (frame.hex()[:80], cap, hex(ro), hex(int(r)))
        if helpers.is_err(ro):
            n_err += 1
        else:
            assert d[:ro].tobytes() == oo
    assert n_err > 400


def test_capacity_larger_than_the_context_arena():
    """The reference accepts any dstCapacity (ZStdDecompress.cs:2182-2191): a small frame decoded into a large reusable
    scratch buffer must not be refused because the buffer is larger than the context's max_batch_bytes; a frame that
    overshoots its declared content size still gets the reference's verdict, not the clamped capacity's."""
    import zstandard_b200 as zb
    from tools import zstd_ref
    orc = helpers.Oracle()
    rng = random.Random(12)
    ctx = zb.Context(max_batch_bytes=1 << 20)
    try:
        data = helpers.sample_payload(rng, 0, 50000)
        frame = zstd_ref.compress(data, 3, checksum=True)
        lying = bytearray(frame)                   # FCS field (2 bytes at offset 5, value - 256) declares 100 bytes less
        assert lying[4] >> 6 == 1 and (lying[4] >> 5) & 1
        fcs = int.from_bytes(lying[5:7], "little") - 100
        lying[5:7] = fcs.to_bytes(2, "little")
        items = [(frame, 4 << 20), (bytes(lying), 4 << 20), (frame, 50000), (bytes(lying), 50000 - 100), (frame, 49999)]
        dsts = [np.zeros(c, dtype=np.uint8) for _, c in items]
        res = ctx.decompress_batch([f for f, _ in items], dsts)
        for (f, cap), r, d in zip(items, res, dsts):
            want, out, _ = orc.decompress(f, cap)
            assert int(r) == want, (cap, hex(int(r)), hex(want))
            if not helpers.is_err(want):
                assert d[:want].tobytes() == out
        # compress: a destination far larger than the bound is fine as well
        out = np.zeros(8 << 20, dtype=np.uint8)
        r = ctx.compress_batch([np.frombuffer(data, dtype=np.uint8)], [out], level=1, checksum=True)[0]
        assert not helpers.is_err(int(r))
        ro, oo, _ = orc.decompress(out[:int(r)].tobytes(), len(data))
        assert ro == len(data) and oo == data
    finally:
        ctx.close()


def test_pinned_contiguous_buffers_take_the_direct_path_and_pageable_ones_the_staged_path(gpu_ctx, oracle):
    """The same batch through pinned memory laid out back to back (direct DMA), through pageable numpy arrays and through
    a non-contiguous destination (refused)."""
    import ctypes
    import zstandard_b200 as zb
    lib = zb.load_library()
    frames = helpers.make_frames(77, 40, sizes=[3000, 20000, 65536])
    total_src = sum(len(f) for f, _ in frames); total_dst = sum(len(d) for _, d in frames)
    lib.zstdb200_host_alloc.restype = ctypes.c_void_p
    hs = lib.zstdb200_host_alloc(total_src + 64); hd = lib.zstdb200_host_alloc(total_dst + 64)
    try:
        src = np.ctypeslib.as_array((ctypes.c_uint8 * total_src).from_address(hs))
        dst = np.ctypeslib.as_array((ctypes.c_uint8 * total_dst).from_address(hd))
        srcs, dsts, so, do = [], [], 0, 0
        for f, d in frames:
            src[so:so + len(f)] = np.frombuffer(f, dtype=np.uint8)
            srcs.append(src[so:so + len(f)]); dsts.append(dst[do:do + len(d)])
            so += len(f); do += len(d)
        res = gpu_ctx.decompress_batch(srcs, dsts)
        pageable = [np.zeros(len(d), dtype=np.uint8) for _, d in frames]
        res2 = gpu_ctx.decompress_batch([f for f, _ in frames], pageable)
        for (f, d), r, r2, o, o2 in zip(frames, res, res2, dsts, pageable):
            assert int(r) == int(r2) == len(d)
            assert o.tobytes() == d and o2.tobytes() == d
    finally:
        lib.zstdb200_host_free(ctypes.c_void_p(hs)); lib.zstdb200_host_free(ctypes.c_void_p(hd))
    strided = np.zeros(2 * len(frames[0][1]), dtype=np.uint8)[::2]
    with pytest.raises(ValueError):
        gpu_ctx.decompress_batch([frames[0][0]], [strided])


@pytest.mark.skipif(os.environ.get("ZSTDB200_TEST_INNER") == "1", reason="already inside the re-run")
@pytest.mark.parametrize("a_max,b_max", [("1000000", "1000000"), ("0", "1000000")])
def test_parity_suite_with_wrong_sequence_class_guesses(a_max, b_max):
    """k_parse sends frames of few sequences to sequence-kernel instantiations with small tables; a frame whose tables do
    not fit is handed to the full-size instantiation (k_seq_t, SeqEmitter::defer).  With real encoders that almost never
    happens, so the oracle-parity tests of this file and of the dictionary suite run again with every frame starting in
    the smallest (then in the middle) class: most 64 KiB frames are handed over, results must not change."""
    import subprocess
    env = dict(os.environ, ZSTDB200_SEQ_A_MAX=a_max, ZSTDB200_SEQ_B_MAX=b_max, ZSTDB200_TEST_INNER="1")
    keep = ("valid_frames or edge_cases or fuzzed or several_data_frames or bad_item or golden or sequence_count or rle_literals "
            "or repeat_mode or look_ahead or round_trip_properties or dictionary")
    p = subprocess.run([sys.executable, "-m", "pytest", "tests/test_decode_gpu.py", "tests/test_dictionary.py", "-m", "gpu", "-x", "-q", "-k", keep],
                       capture_output=True, text=True, env=env, cwd=helpers.ROOT, timeout=1500)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    assert " passed" in p.stdout and "no tests ran" not in p.stdout
