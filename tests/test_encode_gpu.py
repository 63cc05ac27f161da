"""GPU encoder through the C ABI: frames must be accepted by the oracle (the reference decoder's rules) and libzstd,
round-trip bit-exact, and stay in the ratio band.  (The entropy stage is the code tests/test_encode_hostsim.py replays
on the CPU; the match stage is warp-parallel on the GPU and may parse differently from the serial replay.)"""
import random

import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu

SIZES = [0, 1, 2, 5, 17, 63, 64, 100, 255, 256, 1000, 1024, 4096, 5000, 16384, 65536, 70000, 131072, 131073, 200000, 300000]


def _compress(ctx, payloads, level, checksum=True, caps=None):
    import zstandard_b200 as zb
    caps = caps or [zb.ZStdCompress.CompressBound(len(p)) for p in payloads]
    dsts = [np.zeros(max(c, 1), dtype=np.uint8)[:c] for c in caps]
    res = ctx.compress_batch(payloads, dsts, level=level, checksum=checksum)
    return res, dsts


def test_frames_round_trip_through_reference_rules(gpu_ctx, oracle):
    from tools import zstd_ref
    rng = random.Random(6)
    payloads = [helpers.sample_payload(rng, t % 6, rng.choice(SIZES)) for t in range(120)]
    for level in (1, 2, 3):
        for checksum in (True, False):
            res, dsts = _compress(gpu_ctx, payloads, level, checksum)
            for p, r, d in zip(payloads, res, dsts):
                assert not helpers.is_err(int(r)), hex(int(r))
                f = d[:int(r)].tobytes()
                ro, oo, _ = oracle.decompress(f, len(p))
                assert ro == len(p) and oo == p, (len(p), level, hex(ro))
                assert zstd_ref.decompress(f, len(p)) == p


def test_gpu_matches_lockstep_cpu_emulation(gpu_ctx, hostsim):
    """tests/hostsim emulates the warp-parallel match finder lane by lane: the GPU must produce the same bytes."""
    import ctypes
    lib = hostsim.lib
    lib.hostsim_compress_warp2.restype = ctypes.c_uint32
    lib.hostsim_compress_warp2.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_uint32]
    rng = random.Random(8)
    payloads = [helpers.sample_payload(rng, t % 6, rng.choice(SIZES)) for t in range(48)]
    small = [p for p in payloads if len(p) <= 131072]
    for batch in (payloads, small):                 # the table sizes follow the largest chunk of the call (zb_encode.cuh enc_hlog_*)
        biggest = max(len(p) for p in batch)
        for level in (1, 2, 3):
            res, dsts = _compress(gpu_ctx, batch, level, False)
            for p, r, d in zip(batch, res, dsts):
                cap = len(p) + len(p) // 128 + 128
                buf = ctypes.create_string_buffer(cap)
                n = lib.hostsim_compress_warp2(buf, cap, p, len(p), level, 0, biggest)
                assert int(r) == n and d[:n].tobytes() == buf.raw[:n], (len(p), level, int(r), n)


def test_gpu_frames_decode_on_gpu(gpu_ctx):
    rng = random.Random(7)
    payloads = [helpers.sample_payload(rng, t % 6, rng.choice(SIZES)) for t in range(60)]
    res, dsts = _compress(gpu_ctx, payloads, 3, True)
    frames = [d[:int(r)] for r, d in zip(res, dsts)]
    outs = [np.zeros(max(len(p), 1), dtype=np.uint8)[:len(p)] for p in payloads]
    dres = gpu_ctx.decompress_batch(frames, outs)
    for p, r, o in zip(payloads, dres, outs):
        assert int(r) == len(p) and o.tobytes() == p


@pytest.mark.parametrize("kind", ["log", "tick", "mixed"])
def test_ratio_band_and_properties_at_size(gpu_ctx, kind):
    """16 MiB in 128 KiB chunks (BASELINE.json config 3 shape): ratio within 3 % of libzstd at the same level,
    every frame decodes back (through libzstd, an independent decoder) to its chunk."""
    from tools import corpus, zstd_ref
    total, chunk = 16 << 20, 131072
    raw = corpus.make(kind, total)
    chunks = [raw[i:i + chunk] for i in range(0, total, chunk)]
    for level in (1, 2, 3):
        res, dsts = _compress(gpu_ctx, chunks, level, True)
        assert not any(helpers.is_err(int(r)) for r in res)
        ours = int(res.astype(np.int64).sum())
        _, off = zstd_ref.compress_chunks(raw, chunk, level=level, checksum=True)
        ref = int(off[-1])
        assert ours <= ref * 1.03, (kind, level, ours, ref)
        for k in range(0, len(chunks), 7):
            assert zstd_ref.decompress(dsts[k][:int(res[k])].tobytes(), chunk) == chunks[k].tobytes()


@pytest.mark.parametrize("chunk", [4096, 16384, 65536, 262144, 1048576])
def test_ratio_band_over_the_chunk_size_sweep(gpu_ctx, chunk):
    """BASELINE.json configs[4]: tick records in chunks of 4 KiB .. 1 MiB.  Ratio within 3 % of libzstd 1.5.5 at the same
    level (the reference has no compressor: this band is unpinned by the reference), frames decode back through libzstd."""
    from tools import corpus, zstd_ref
    total = 8 << 20
    raw = corpus.make("tick", total)
    chunks = [raw[i:i + chunk] for i in range(0, total, chunk)]
    for level in (1, 3):
        res, dsts = _compress(gpu_ctx, chunks, level, True)
        assert not any(helpers.is_err(int(r)) for r in res)
        ours = int(res.astype(np.int64).sum())
        _, off = zstd_ref.compress_chunks(raw, chunk, level=level, checksum=True)
        ref = int(off[-1])
        assert ours <= ref * 1.03, (chunk, level, total / ours, total / ref)
        for k in range(0, len(chunks), max(1, len(chunks) // 16)):
            assert zstd_ref.decompress(dsts[k][:int(res[k])].tobytes(), chunk) == chunks[k].tobytes()


def test_destination_too_small_is_per_item(gpu_ctx):
    from tools import corpus
    a, b = corpus.log(5000).tobytes(), corpus.tick(5000).tobytes()
    res, dsts = _compress(gpu_ctx, [a, b, a], 3, True, caps=[6000, 50, 6000])
    assert not helpers.is_err(int(res[0])) and not helpers.is_err(int(res[2]))
    assert int(res[1]) == helpers.err(70)


def test_device_pointer_api_and_kernel_times(gpu_ctx):
    """zstdb200_compress_batch_device(_timed): device-resident chunks in, frames out; every frame decodes back."""
    import torch
    from tools import corpus, zstd_ref
    import zstandard_b200 as zb
    chunk, n = 32768, 96
    raw = corpus.make("mixed", chunk * n)
    bound = (zb.ZStdCompress.CompressBound(chunk) + 15) // 16 * 16
    dev = torch.device("cuda:0")
    t_src = torch.zeros(chunk * n + 64, dtype=torch.uint8, device=dev)
    t_src[:chunk * n] = torch.from_numpy(raw).to(dev)
    t_dst = torch.zeros(n * bound + 64, dtype=torch.uint8, device=dev)
    t_soff = torch.arange(n, dtype=torch.int64, device=dev) * chunk
    t_doff = torch.arange(n, dtype=torch.int64, device=dev) * bound
    t_ssz = torch.full((n,), chunk, dtype=torch.int32, device=dev)
    t_cap = torch.full((n,), bound, dtype=torch.int32, device=dev)
    t_res = torch.zeros(n, dtype=torch.int32, device=dev)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    for level in (1, 2, 3):
        ms = gpu_ctx.compress_batch_device_timed(level, True, t_src.data_ptr(), t_soff.data_ptr(), t_ssz.data_ptr(), t_dst.data_ptr(),
                                                 t_doff.data_ptr(), t_cap.data_ptr(), t_res.data_ptr(), n, stream=st.cuda_stream)
        assert set(ms) == {"k_enc_match", "k_enc_entropy", "k_enc_xxh"} and all(v >= 0 for v in ms.values())
        res = t_res.cpu().numpy().view(np.uint32)
        out = t_dst.cpu().numpy()
        for k in range(n):
            assert not helpers.is_err(int(res[k]))
            assert zstd_ref.decompress(out[k * bound:k * bound + int(res[k])].tobytes(), chunk) == raw[k * chunk:(k + 1) * chunk].tobytes()


def test_bad_arguments_are_batch_level_errors(gpu_ctx):
    import zstandard_b200 as zb
    src = np.frombuffer(b"x" * 100, dtype=np.uint8)
    dst = np.zeros(200, dtype=np.uint8)
    with pytest.raises(RuntimeError):
        gpu_ctx.compress_batch([src], [dst], level=7)          # only levels 1..3 exist
    assert len(gpu_ctx.compress_batch([], [])) == 0           # an empty batch is not an error
    assert len(gpu_ctx.decompress_batch([], [])) == 0
    assert zb.ZStdCompress.CompressBound(0) > 0


def test_generous_destination_capacities_fit_the_staging():
    """Destination capacities well above the compress bound: the sum over a sub-batch may exceed the batch size; the
    device staging has to be sized for it (it once was not)."""
    import zstandard_b200 as zb
    from tools import corpus, zstd_ref
    ctx = zb.Context(max_batch_bytes=4 << 20)
    try:
        chunk, n = 131072, 30
        raw = corpus.make("log", chunk * n)
        chunks = [raw[i * chunk:(i + 1) * chunk] for i in range(n)]
        outs = [np.zeros(300 * 1024, dtype=np.uint8) for _ in range(n)]
        res = ctx.compress_batch(chunks, outs, level=3, checksum=True)
        for c, r, o in zip(chunks, res, outs):
            assert not helpers.is_err(int(r))
            assert zstd_ref.decompress(o[:int(r)].tobytes(), chunk) == c.tobytes()
        big = np.zeros(n * 300 * 1024, dtype=np.uint8)                      # the same, destinations back to back (direct DMA)
        outs2 = [big[i * 300 * 1024:(i + 1) * 300 * 1024] for i in range(n)]
        res2 = ctx.compress_batch(chunks, outs2, level=1, checksum=False)
        for c, r, o in zip(chunks, res2, outs2):
            assert zstd_ref.decompress(o[:int(r)].tobytes(), chunk) == c.tobytes()
    finally:
        ctx.close()
