#!/usr/bin/env python3
"""Minimal device-resident compress loop for profilers / quick timing.

    python tools/profile_encode.py [--bytes N] [--corpus log] [--chunk 131072] [--level 3] [--iters 2]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bytes", type=int, default=256 << 20)
    ap.add_argument("--corpus", default="log")
    ap.add_argument("--chunk", type=int, default=131072)
    ap.add_argument("--level", type=int, default=3)
    ap.add_argument("--iters", type=int, default=2)
    args = ap.parse_args()
    import torch
    import zstandard_b200 as zb
    from tools import corpus, zstd_ref
    raw = corpus.make(args.corpus, args.bytes)
    total, chunk = len(raw), args.chunk
    n = (total + chunk - 1) // chunk
    bound = zb.ZStdCompress.CompressBound(chunk)
    bound = (bound + 15) // 16 * 16
    dev = torch.device("cuda:0")
    ctx = zb.Context(devices=[0], max_batch_bytes=max(total, 1 << 20))
    t_src = torch.empty(total + 64, dtype=torch.uint8, device=dev)
    t_src[:total] = torch.from_numpy(raw).to(dev)
    t_dst = torch.zeros(n * bound + 64, dtype=torch.uint8, device=dev)
    t_soff = torch.from_numpy(np.arange(n, dtype=np.int64) * chunk).to(dev)
    ssz = np.array([min(chunk, total - i * chunk) for i in range(n)], dtype=np.int32)
    t_ssz = torch.from_numpy(ssz).to(dev)
    t_doff = torch.from_numpy(np.arange(n, dtype=np.int64) * bound).to(dev)
    t_dcap = torch.full((n,), bound, dtype=torch.int32, device=dev)
    t_res = torch.zeros(n, dtype=torch.int32, device=dev)
    s = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    times = []
    for _ in range(args.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        ctx.compress_batch_device(args.level, True, t_src.data_ptr(), t_soff.data_ptr(), t_ssz.data_ptr(), t_dst.data_ptr(), t_doff.data_ptr(),
                                  t_dcap.data_ptr(), t_res.data_ptr(), n, stream=s.cuda_stream)
        e1.record(s)
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    kms = ctx.compress_batch_device_timed(args.level, True, t_src.data_ptr(), t_soff.data_ptr(), t_ssz.data_ptr(), t_dst.data_ptr(),
                                          t_doff.data_ptr(), t_dcap.data_ptr(), t_res.data_ptr(), n, stream=s.cuda_stream)
    print("kernel_ms", {k: round(v, 3) for k, v in kms.items()})
    res = t_res.cpu().numpy().view(np.uint32)
    assert (res < 0xFFFFFF88).all()
    comp = int(res.astype(np.int64).sum())
    out = t_dst.cpu().numpy()
    for k in range(0, n, max(1, n // 16)):
        f = out[k * bound:k * bound + int(res[k])].tobytes()
        assert zstd_ref.decompress(f, int(ssz[k])) == raw[k * chunk:k * chunk + int(ssz[k])].tobytes()
    t0 = time.perf_counter()
    _, off = zstd_ref.compress_chunks(raw, chunk, level=args.level, checksum=True, threads=os.cpu_count() or 1)
    cpu_s = time.perf_counter() - t0
    print("frames", n, "raw", total, "ours", comp, "ratio", round(total / comp, 4), "libzstd ratio", round(total / int(off[-1]), 4),
          "ms", [round(t, 3) for t in times], "GB/s", round(total / min(times) / 1e6, 2), "libzstd all-core GB/s", round(total / cpu_s / 1e9, 2))
    ctx.close()


if __name__ == "__main__":
    main()
