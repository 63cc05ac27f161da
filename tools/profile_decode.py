#!/usr/bin/env python3
"""Minimal device-resident decode loop for profilers (ncu): same workload builder as bench.py, nothing else.

    python tools/profile_decode.py [--bytes N] [--corpus log] [--chunk 65536] [--iters 2]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bytes", type=int, default=256 << 20)
    ap.add_argument("--corpus", default="log")
    ap.add_argument("--chunk", type=int, default=65536)
    ap.add_argument("--level", type=int, default=3)
    ap.add_argument("--iters", type=int, default=2)
    args = ap.parse_args()
    import torch
    import zstandard_b200 as zb
    from tools import corpus, zstd_ref
    raw = corpus.make(args.corpus, args.bytes)
    blob, off = zstd_ref.compress_chunks(raw, args.chunk, level=args.level, checksum=True, threads=os.cpu_count() or 1)
    n, total, comp = len(off) - 1, len(raw), int(off[-1])
    dev = torch.device("cuda:0")
    ctx = zb.Context(devices=[0], max_batch_bytes=total)
    t_src = torch.empty(comp + 64, dtype=torch.uint8, device=dev)
    t_src[:comp] = torch.from_numpy(blob).to(dev)
    t_dst = torch.zeros(total + 64, dtype=torch.uint8, device=dev)
    t_soff = torch.from_numpy(off[:-1].astype(np.int64)).to(dev)
    t_ssz = torch.from_numpy(np.diff(off).astype(np.int32)).to(dev)
    t_doff = torch.from_numpy(np.arange(n, dtype=np.int64) * args.chunk).to(dev)
    dcap = np.array([min(args.chunk, total - i * args.chunk) for i in range(n)], dtype=np.int32)
    t_dcap = torch.from_numpy(dcap).to(dev)
    t_res = torch.zeros(n, dtype=torch.int32, device=dev)
    s = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    for _ in range(args.iters):
        ctx.decompress_batch_device(t_src.data_ptr(), t_soff.data_ptr(), t_ssz.data_ptr(), t_dst.data_ptr(), t_doff.data_ptr(),
                                    t_dcap.data_ptr(), t_res.data_ptr(), n, stream=s.cuda_stream)
        torch.cuda.synchronize()
    res = t_res.cpu().numpy().view(np.uint32)
    assert (res == dcap.view(np.uint32)).all()
    assert torch.equal(t_dst[:total].cpu(), torch.from_numpy(raw))
    ms = ctx.decompress_batch_device_timed(t_src.data_ptr(), t_soff.data_ptr(), t_ssz.data_ptr(), t_dst.data_ptr(), t_doff.data_ptr(),
                                           t_dcap.data_ptr(), t_res.data_ptr(), n, stream=s.cuda_stream)
    print("frames", n, "raw", total, "compressed", comp, "kernel_ms", {k: round(v, 4) for k, v in ms.items()})
    ctx.close()


if __name__ == "__main__":
    main()
