#!/usr/bin/env python3
"""bench.py — headline benchmark of libzstdb200 (BASELINE.json: GB/s uncompressed, 64 KiB frames).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload decode64k|compress128k]

One process per GPU (under torchrun for N > 1: RANK/LOCAL_RANK/WORLD_SIZE from the env).  Frames are independent,
so ranks shard the work with no data-path collective ("scaling": "weak": every rank processes --bytes of its own
synthetic data); torch.distributed is used only for the barrier and the max-over-ranks time.

A step = one pass of the hot path over one batch: decode64k (BASELINE.json configs[1]) decodes --bytes (1 GiB)
of libzstd-level-3 compressed log text held as 64 KiB frames.
  value     device-resident pipeline: frames and outputs already in HBM, CUDA events on the launching stream
  e2e       the same work through the host-buffer C-ABI call (pinned host memory -> H2D -> kernels -> D2H)
  roofline  algorithmic bytes (sum of frame bytes + content bytes) / duration of the dominant kernel
  cpu_baseline  the oracle (C++ port of the reference decoder; the C# reference cannot run here) on all host cores
`--impl reference` times that oracle port alone, rank 0 only, as the reference arm.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHUNK = {"decode64k": 65536, "compress128k": 131072}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="decode64k", choices=list(CHUNK))
    ap.add_argument("--bytes", type=int, default=1 << 30, help="uncompressed bytes per GPU per step")
    ap.add_argument("--corpus", default="log", choices=["log", "tick", "random", "mixed"])
    ap.add_argument("--level", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (B200_PROFILING.md clocks line, via NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def prepare_decode(args, rank):
    """-> dict with raw corpus, frame blob, offsets (uint64 n+1) for this rank's shard."""
    from tools import corpus, zstd_ref
    chunk = CHUNK[args.workload]
    raw = corpus.make(args.corpus, args.bytes, shard=rank)
    blob, off = zstd_ref.compress_chunks(raw, chunk, level=args.level, checksum=True, threads=max(1, os.cpu_count() or 1))
    return {"raw": raw, "blob": blob, "off": off, "chunk": chunk, "n": len(off) - 1}


class OracleBatch:
    """The oracle's threaded batch decode over host pointer arrays (cpu_baseline / reference arm)."""

    def __init__(self, w):
        from tests import helpers
        self.lib = helpers.Oracle().lib
        n, chunk, total = w["n"], w["chunk"], len(w["raw"])
        self.out = np.zeros(total, dtype=np.uint8)
        off = w["off"]
        base_s, base_d = w["blob"].ctypes.data, self.out.ctypes.data
        self.sp = (ctypes.c_void_p * n)(*[base_s + int(off[i]) for i in range(n)])
        self.dp = (ctypes.c_void_p * n)(*[base_d + i * chunk for i in range(n)])
        self.ss = np.diff(off).astype(np.uint32)
        self.dc = np.array([min(chunk, total - i * chunk) for i in range(n)], dtype=np.uint32)
        self.res = np.zeros(n, dtype=np.uint32)
        self.n, self.total = n, total
        self.lib.oracle_decompress_batch.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_uint64, ctypes.c_int]
        self.lib.oracle_decompress_batch.restype = None

    def run(self, threads, count=None):
        n = self.n if count is None else count
        t = time.perf_counter()
        self.lib.oracle_decompress_batch(self.sp, self.ss.ctypes.data, self.dp, self.dc.ctypes.data, self.res.ctypes.data, n, threads)
        dt = time.perf_counter() - t
        assert (self.res[:n] == self.dc[:n]).all(), "oracle failed to decode the workload"
        return dt, int(self.dc[:n].sum())


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU algorithm for this path (oracle port; C# cannot run in this image)."""
    if rank != 0:
        return
    w = prepare_decode(args, 0)
    ob = OracleBatch(w)
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        ob.run(cores)
    t = 0.0
    for _ in range(args.steps):
        dt, nbytes = ob.run(cores)
        t += dt
    gbs = nbytes * args.steps / t / 1e9
    line = {"impl": "reference", "metric": "decompress_GBps_uncompressed", "value": round(gbs, 3), "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * t / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, w),
            "cpu_baseline": {"value": round(gbs, 3), "unit": "GB/s", "cores": cores, "kind": "port",
                             "sample": f"full step: {w['n']} frames, {nbytes} bytes, {args.steps} steps"},
            "e2e": {"value": round(gbs, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, w):
    return {"workload": f"{args.workload}: batched decompression of {args.bytes} B/GPU of libzstd-1.5.5 level-{args.level} "
                        f"{args.corpus} text in {w['chunk'] // 1024} KiB independent frames with XXH64 checksums",
            "frames_per_gpu": w["n"], "compressed_bytes_per_gpu": int(w["off"][-1]),
            "ratio": round(len(w["raw"]) / int(w["off"][-1]), 3),
            "l2": "inputs+outputs per step exceed the 126 MB L2 (no flush needed)", "sharding": "by frame, no collective"}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload != "decode64k":
        raise SystemExit("only decode64k is wired into bench.py so far")

    import torch
    import zstandard_b200 as zb
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    w = prepare_decode(args, rank)
    n, chunk, total = w["n"], w["chunk"], len(w["raw"])
    comp = int(w["off"][-1])
    ctx = zb.Context(devices=[local], max_batch_bytes=total)
    lib = zb.load_library()

    # ---------------- device-resident arm ----------------
    t_src = torch.empty(comp + 64, dtype=torch.uint8, device=dev)
    t_src[:comp] = torch.from_numpy(w["blob"]).to(dev)
    t_dst = torch.zeros(total + 64, dtype=torch.uint8, device=dev)
    soff = w["off"][:-1].astype(np.int64)
    ssz = np.diff(w["off"]).astype(np.int32)
    doff = (np.arange(n, dtype=np.int64) * chunk)
    dcap = np.array([min(chunk, total - i * chunk) for i in range(n)], dtype=np.int32)
    t_soff, t_ssz = torch.from_numpy(soff).to(dev), torch.from_numpy(ssz).to(dev)
    t_doff, t_dcap = torch.from_numpy(doff).to(dev), torch.from_numpy(dcap).to(dev)
    t_res = torch.zeros(n, dtype=torch.int32, device=dev)
    # a non-default stream: handle 0 would mean "the context's own stream" to the C ABI
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    dargs = (t_src.data_ptr(), t_soff.data_ptr(), t_ssz.data_ptr(), t_dst.data_ptr(), t_doff.data_ptr(), t_dcap.data_ptr(),
             t_res.data_ptr(), n)

    def step_device():
        ctx.decompress_batch_device(*dargs, stream=stream)

    for _ in range(max(3, args.warmup)):
        step_device()
    torch.cuda.synchronize()
    # correctness of what is being timed: every frame decoded to its chunk
    res = t_res.cpu().numpy().view(np.uint32)
    assert (res == dcap.view(np.uint32)).all(), "decode failed on the bench workload"
    assert torch.equal(t_dst[:total].cpu(), torch.from_numpy(w["raw"])), "decoded bytes differ from the corpus"

    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.kernel_launches - l0
    sampler.stop_flag = True
    sampler.join()
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())

    # per-kernel device times (CUDA events between the kernels, separate synchronised runs)
    kms = {}
    reps = max(3, args.steps)
    for _ in range(reps):
        for k, v in ctx.decompress_batch_device_timed(*dargs, stream=stream).items():
            kms[k] = kms.get(k, 0.0) + v / reps
    dom = max(kms, key=kms.get)
    algo_bytes = comp + total
    peak, peak_src = measured_peak()
    achieved = algo_bytes / (kms[dom] * 1e-3) / 1e9

    # ---------------- end-to-end arm: host buffers through the C ABI ----------------
    h_src = lib.zstdb200_host_alloc(comp + 64)
    h_dst = lib.zstdb200_host_alloc(total + 64)
    ctypes.memmove(h_src, w["blob"].ctypes.data, comp)
    sp = (ctypes.c_void_p * n)(*[h_src + int(w["off"][i]) for i in range(n)])
    dp = (ctypes.c_void_p * n)(*[h_dst + i * chunk for i in range(n)])
    ss_u, dc_u, res_u = ssz.view(np.uint32).copy(), dcap.view(np.uint32).copy(), np.zeros(n, dtype=np.uint32)
    u32p = ctypes.POINTER(ctypes.c_uint32)

    def step_e2e():
        rc = lib.zstdb200_decompress_batch(ctx.handle, sp, ss_u.ctypes.data_as(u32p), dp, dc_u.ctypes.data_as(u32p),
                                           res_u.ctypes.data_as(u32p), n)
        assert rc == 0, ctx.last_error()

    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    assert (res_u == dc_u).all()
    got = np.ctypeslib.as_array(ctypes.cast(h_dst, ctypes.POINTER(ctypes.c_uint8)), shape=(total,))
    assert (got == w["raw"]).all(), "e2e bytes differ"
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_s = float(t_e.item())
    lib.zstdb200_host_free(h_src)
    lib.zstdb200_host_free(h_dst)

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ob = OracleBatch(w)
        cores = os.cpu_count() or 1
        ob.run(cores, min(n, 1024))
        tt, nb, runs = 0.0, 0, 0
        while tt < 3.0 and runs < 50:
            dt, b = ob.run(cores)
            tt += dt; nb += b; runs += 1
        cpu = {"value": round(nb / tt / 1e9, 3), "unit": "GB/s", "cores": cores, "kind": "port",
               "sample": f"{runs} pass(es) over the full step ({n} frames, {total} B) = {tt * cores:.1f} core-seconds; "
                         "oracle = C++ restatement of the C# decoder, static frame partition over all cores"}

    if rank == 0:
        value = total * world * args.steps / (ms * 1e-3) / 1e9
        line = {
            "metric": "decompress_GBps_uncompressed", "value": round(value, 3), "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, w),
            "e2e": {"value": round(total * world * args.steps / e2e_s / 1e9, 3), "unit": "GB/s",
                    "h2d_bytes_per_step": int(comp + n * 24), "d2h_bytes_per_step": int(total + n * 4)},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 5),
                         "traffic": None, "kernel": dom, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(algo_bytes),
                         "pipeline_frac": round(algo_bytes * args.steps / (ms * 1e-3) / 1e9 / peak, 5)},
            "kernel_ms": {k: round(v, 4) for k, v in kms.items()},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
