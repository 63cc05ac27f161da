#!/usr/bin/env python3
"""PCIe floor of the host-buffer decode step on this box: every rank copies 1 GiB device->pinned host and 220 MB
pinned host->device at the same time (the bytes bench.py's e2e arm moves per step), barrier-timed, max over ranks.

    python -m torch.distributed.run --nproc-per-node N tools/pcie_floor.py     (or plain python for N = 1)
"""
import os
import time

import torch


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    out_b, in_b = 1 << 30, 220 << 20
    h_out = torch.empty(out_b, dtype=torch.uint8).pin_memory(); d_out = torch.empty(out_b, dtype=torch.uint8, device=dev)
    h_in = torch.empty(in_b, dtype=torch.uint8).pin_memory(); d_in = torch.empty(in_b, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def step():
        with torch.cuda.stream(s1):
            h_out.copy_(d_out, non_blocking=True)
        with torch.cuda.stream(s2):
            d_in.copy_(h_in, non_blocking=True)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print({"n_gpus": world, "ms_per_step": round(float(t.item()) * 1e3, 2), "decode_e2e_ceiling_GBps": round(world * out_b / float(t.item()) / 1e9, 1)})
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
