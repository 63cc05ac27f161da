"""Dictionary decode: ZSTD_decompress_usingDict, csharp/src/ZStdDecompress.cs:2162-2167 with the
dictionary loading of :2366-2475 (present in the reference, unreachable from its public API, which passes null, :2171).

Pins: the reference ships no dictionary test, so this row is pinned by oracle <-> libzstd 1.5.5 agreement on dictionary
frames (trained dictionaries from ZDICT_trainFromBuffer and raw-content dictionaries) — parity unpinned by the reference
itself.  CPU tests: oracle vs libzstd, and the g++ replay of the kernels' stage code vs the oracle.  GPU tests: the CUDA
path through zstdb200_load_dictionary + zstdb200_decompress_batch vs the oracle (result codes and bytes)."""
import ctypes
import random

import numpy as np
import pytest

from tests import helpers


def _dict_oracle(oracle):
    lib = oracle.lib
    lib.oracle_decompress_using_dict.restype = ctypes.c_uint32
    lib.oracle_decompress_using_dict.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint32]

    def run(frame, cap, dictionary):
        buf = ctypes.create_string_buffer(max(cap, 1) + 8)
        r = lib.oracle_decompress_using_dict(buf, cap, frame, len(frame), dictionary if dictionary else None, len(dictionary) if dictionary else 0)
        return r, (buf.raw[:r] if not helpers.is_err(r) else None)
    return run


def _material(seed, count):
    """(trained dictionary, raw-content dictionary, [(frame, plaintext, which dictionary)]): 4 KiB-ish log / tick messages, the
    shape dictionaries exist for, plus larger and multi-block payloads, levels 1-19, with and without checksum / dict id."""
    from tools import corpus, zstd_ref
    rng = random.Random(seed)
    log, tick = corpus.log(6 << 20).tobytes(), corpus.tick(2 << 20).tobytes()
    trained = zstd_ref.train_dict([log[i:i + 4096] for i in range(0, 3 << 20, 4096)], 32768)
    raw = log[:30000]
    items = []
    for k in range(count):
        n = rng.choice([0, 1, 50, 500, 4096, 4096, 4096, 20000, 70000, 140000])
        i = rng.randrange(3 << 20, (6 << 20) - n)
        data = log[i:i + n] if k % 4 else tick[:n]
        which = k % 2
        frame = zstd_ref.compress_with_dict(data, trained if which == 0 else raw, rng.choice([1, 3, 5, 19]), rng.random() < 0.7, rng.random() < 0.8)
        items.append((frame, data, which))
    return trained, raw, items


def test_oracle_agrees_with_libzstd_on_dictionary_frames(oracle):
    from tools import zstd_ref
    run = _dict_oracle(oracle)
    trained, raw, items = _material(21, 120)
    assert int.from_bytes(trained[:4], "little") == 0xEC30A437
    used_dict = 0
    for frame, data, which in items:
        d = trained if which == 0 else raw
        r, out = run(frame, len(data), d)
        assert r == len(data) and out == data
        assert zstd_ref.decompress_with_dict(frame, len(data), d) == data
        # without the dictionary: dictionary_wrong when the frame names one, otherwise whatever the reference decodes
        r0, _ = run(frame, len(data), b"")
        assert r0 == oracle.decompress(frame, len(data))[0]
        used_dict += r0 != len(data)
    assert used_dict > 40          # most frames really depend on their dictionary
    # a frame that names another dictionary id is refused (:633)
    frame, data, _ = next(it for it in items if it[2] == 0 and (it[0][4] & 3))
    other = bytearray(trained); other[4] ^= 0x55
    assert run(frame, len(data), bytes(other))[0] == helpers.err(32)
    # a malformed entropy section fails every data frame with dictionary_corrupted (:2501-2507)
    broken = bytearray(trained); broken[8:40] = bytes(32)
    assert run(frame, len(data), bytes(broken))[0] == helpers.err(30)


def test_replayed_stages_match_the_oracle_with_dictionaries(hostsim, oracle):
    lib = hostsim.lib
    lib.hostsim_set_dict.argtypes = [ctypes.c_char_p, ctypes.c_uint32]
    run = _dict_oracle(oracle)
    trained, raw, items = _material(22, 160)
    rng = random.Random(5)
    try:
        for which, d in ((0, trained), (1, raw)):
            lib.hostsim_set_dict(d, len(d))
            for k, (frame, data, w) in enumerate(items):
                if w != which:
                    continue
                if k % 5 == 0 and len(frame) > 10:
                    frame = helpers.mutate(rng, frame)
                cap = max(0, len(data) + rng.choice([0, 0, 7, -1]))
                want, out = run(frame, cap, d)
                r, got = hostsim.decompress(frame, cap, oracle)
                assert r == want, (k, hex(r), hex(want))
                if out is not None:
                    assert got == out
        broken = bytearray(trained); broken[8:40] = bytes(32)
        lib.hostsim_set_dict(bytes(broken), len(broken))
        frame, data, _ = items[0]
        assert hostsim.decompress(frame, len(data), oracle)[0] == run(frame, len(data), bytes(broken))[0] == helpers.err(30)
    finally:
        lib.hostsim_set_dict(None, 0)


@pytest.mark.gpu
def test_gpu_dictionary_decode_matches_the_oracle(oracle):
    import zstandard_b200 as zb
    run = _dict_oracle(oracle)
    trained, raw, items = _material(23, 400)
    rng = random.Random(9)
    ctx = zb.Context(max_batch_bytes=64 << 20)
    try:
        for which, d in ((0, trained), (1, raw)):
            ctx.load_dictionary(d)
            batch = []
            for k, (frame, data, w) in enumerate(items):
                if w != which:
                    continue
                if k % 5 == 0 and len(frame) > 10:
                    frame = helpers.mutate(rng, frame)
                batch.append((frame, max(0, len(data) + rng.choice([0, 0, 7, -1]))))
            dsts = [np.zeros(max(c, 1), dtype=np.uint8)[:c] for _, c in batch]
            res = ctx.decompress_batch([f for f, _ in batch], dsts)
            n_ok = 0
            for (frame, cap), r, o in zip(batch, res, dsts):
                want, out = run(frame, cap, d)
                assert int(r) == want, (hex(int(r)), hex(want))
                if out is not None:
                    assert o[:want].tobytes() == out
                    n_ok += 1
            assert n_ok > 100
        # wrong id, malformed dictionary, and back to no dictionary
        frame, data, _ = next(it for it in items if it[2] == 0 and (it[0][4] & 3))
        out = np.zeros(len(data), dtype=np.uint8)
        other = bytearray(trained); other[4] ^= 0x55
        ctx.load_dictionary(bytes(other))
        assert int(ctx.decompress_batch([frame], [out])[0]) == helpers.err(32)
        broken = bytearray(trained); broken[8:40] = bytes(32)
        ctx.load_dictionary(bytes(broken))
        assert int(ctx.decompress_batch([frame], [out])[0]) == run(frame, len(data), bytes(broken))[0] == helpers.err(30)
        ctx.load_dictionary(None)
        assert int(ctx.decompress_batch([frame], [out])[0]) == oracle.decompress(frame, len(data))[0]
        plain = helpers.make_frames(3, 10)
        outs = [np.zeros(len(p), dtype=np.uint8) for _, p in plain]
        res = ctx.decompress_batch([f for f, _ in plain], outs)
        assert all(int(r) == len(p) and o.tobytes() == p for (f, p), r, o in zip(plain, res, outs))
    finally:
        ctx.close()
