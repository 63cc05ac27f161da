"""Block-parallel decoding of multi-block frames on the GPU (k_parse's walk, k_huf_blk, k_seq_blk, k_exec_big —
zstandard_b200/csrc/zb_blocks.cuh, decode_kernels.cu), through the C ABI, against the oracle: bit-exact bytes and
result codes on valid and damaged frames, mixed with single-block frames in the same batch."""
import ctypes
import os
import random

import numpy as np
import pytest

from tests import helpers
from tests.test_blocks_hostsim import _big_frames
from tests.test_dictionary import _dict_oracle

pytestmark = pytest.mark.gpu


def _gpu_decode(ctx, frames_and_caps):
    srcs = [f for f, _ in frames_and_caps]
    dsts = [np.zeros(max(c, 1), dtype=np.uint8)[:c] for _, c in frames_and_caps]
    res = ctx.decompress_batch(srcs, dsts)
    return res, dsts


def _check(items, res, dsts, oracle):
    for (frame, cap), r, d in zip(items, res, dsts):
        ro, oo, _ = oracle.decompress(frame, cap)
        assert int(r) == ro, (len(frame), cap, hex(int(r)), hex(ro))
        if not helpers.is_err(ro):
            assert d[:ro].tobytes() == oo


def test_multi_block_frames_match_the_oracle(gpu_ctx, oracle):
    frames = _big_frames(301, 36)
    small = helpers.make_frames(303, 40)             # single-block frames interleaved: both paths in one launch
    items = []
    for k, (frame, data) in enumerate(frames):
        n = len(data)
        for cap in (n, n + 13, n - 1, n // 2):
            items.append((frame, cap))
        f2, d2 = small[k]
        items.append((f2, len(d2)))
    res, dsts = _gpu_decode(gpu_ctx, items)
    _check(items, res, dsts, oracle)


def test_fuzzed_multi_block_frames_same_result_code_and_bytes(gpu_ctx, oracle):
    rng = random.Random(78)
    frames = _big_frames(302, 24)
    items = []
    for frame, data in frames:
        for _ in range(14):
            b = helpers.mutate(rng, frame)
            if rng.random() < 0.5:
                i = rng.randrange(len(frame) // 3, len(frame)); bb = bytearray(frame); bb[i] ^= 1 << rng.randrange(8); b = bytes(bb)
            items.append((b, max(0, len(data) + rng.choice([0, 0, 0, 5, -1, -100, 1000]))))
    res, dsts = _gpu_decode(gpu_ctx, items)
    _check(items, res, dsts, oracle)


def test_repeat_offsets_across_blocks(gpu_ctx, oracle):
    from tools import zstd_ref
    rng = random.Random(9)
    items, raws = [], []
    for t in range(12):
        period = rng.choice([5, 24, 24, 100, 333])
        unit = bytes(rng.randrange(256) for _ in range(period))
        body = bytearray(unit * (400000 // period + 1))[:400000]
        for _ in range(rng.choice([0, 40, 400])):
            body[rng.randrange(len(body))] = rng.randrange(256)
        data = bytes(body)
        items.append((zstd_ref.compress(data, rng.choice([1, 3, 5, 19]), checksum=True), len(data))); raws.append(data)
    res, dsts = _gpu_decode(gpu_ctx, items)
    for r, d, raw in zip(res, dsts, raws):
        assert int(r) == len(raw) and d.tobytes() == raw


def test_multi_block_frames_with_dictionaries(oracle):
    import zstandard_b200 as zb
    from tools import corpus, zstd_ref
    run = _dict_oracle(oracle)
    rng = random.Random(23)
    log = corpus.log(5 << 20).tobytes()
    trained = zstd_ref.train_dict([log[i:i + 4096] for i in range(0, 2 << 20, 4096)], 32768)
    raw = log[:30000]
    ctx = zb.Context(max_batch_bytes=32 << 20)
    try:
        for d in (trained, raw):
            ctx.load_dictionary(d)
            items = []
            for k in range(10):
                n = rng.choice([150000, 300000, 600000])
                i = rng.randrange(2 << 20, (5 << 20) - n)
                data = log[i:i + n]
                frame = zstd_ref.compress_with_dict(data, d, rng.choice([1, 3, 5]), True, True)
                if k % 3 == 2:
                    frame = helpers.mutate(rng, frame)
                items.append((frame, len(data) + rng.choice([0, 9, -1])))
            res, dsts = _gpu_decode(ctx, items)
            for (frame, cap), r, o in zip(items, res, dsts):
                want, out = run(frame, cap, d)
                assert int(r) == want, (hex(int(r)), hex(want))
                if out is not None:
                    assert o[:want].tobytes() == out
    finally:
        ctx.close()


def test_block_parallel_and_frame_serial_kernels_agree_at_size(oracle):
    """64 MiB of 1 MiB tick frames (8 blocks each, repeat tables and treeless literals among them): the default
    (block-parallel) context and a ZSTDB200_PAR=0 one produce the same bytes; a checksum of checksums stands in for the
    oracle at this size, the oracle checks a sample."""
    import subprocess, sys, json
    code = r'''
import json, sys, hashlib
import numpy as np
import zstandard_b200 as zb
from tools import corpus, zstd_ref
raw = corpus.tick(64 << 20)
blob, off = zstd_ref.compress_chunks(raw, 1 << 20, level=3, checksum=True)
frames = [blob[int(off[i]):int(off[i + 1])].tobytes() for i in range(len(off) - 1)]
ctx = zb.Context(max_batch_bytes=80 << 20)
dsts = [np.zeros(1 << 20, dtype=np.uint8) for _ in frames]
res = ctx.decompress_batch(frames, dsts)
ok = all(int(r) == (1 << 20) for r in res) and b"".join(d.tobytes() for d in dsts) == raw.tobytes()
print(json.dumps({"ok": bool(ok), "launches": int(ctx.kernel_launches)}))
'''
    outs = []
    for par in ("1", "0"):
        env = dict(os.environ, ZSTDB200_PAR=par)
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=helpers.ROOT, timeout=900)
        assert p.returncode == 0, p.stderr[-2000:]
        outs.append(json.loads(p.stdout.strip().split("\n")[-1]))
    assert outs[0]["ok"] and outs[1]["ok"]
    assert outs[0]["launches"] > outs[1]["launches"]          # the block-parallel kernels did launch
