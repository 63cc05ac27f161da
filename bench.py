#!/usr/bin/env python3
"""bench.py — headline benchmark of libzstdb200 (BASELINE.json: GB/s uncompressed, 64 KiB frames).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload decode64k|compress128k]
                    [--corpus log|tick|random|mixed] [--chunk BYTES] [--level 1..3] [--bytes N]

One process per GPU (under torchrun for N > 1: RANK/LOCAL_RANK/WORLD_SIZE from the env).  Frames are independent,
so ranks shard the work with no data-path collective ("scaling": "weak": every rank processes --bytes of its own
synthetic data); torch.distributed is used only for the barrier and the max-over-ranks time.

A step = one pass of the hot path over one batch:
  decode64k    (default, BASELINE.json configs[1]) decodes --bytes (1 GiB) of libzstd level-3 compressed log text held
               as 64 KiB frames with XXH64 checksums
  compress128k (configs[2]) compresses --bytes of log text in 128 KiB chunks at --level with XXH64 checksums
Reported:
  value         device-resident pipeline: inputs and outputs already in HBM, CUDA events on the launching stream
  e2e           the same work through the host-buffer C-ABI call (pinned host memory -> H2D -> kernels -> D2H)
  roofline      algorithmic bytes (bytes in + bytes out of the codec) / duration of the dominant kernel; `traffic` = that
                kernel's DRAM bytes per launch from the committed ncu capture of this command (profiles/traffic.json)
The other BASELINE.json configs are the same two workloads with other shapes: --corpus mixed (configs[3]), --corpus tick
--chunk 4096..1048576 (configs[4]); profiles/ holds the lines measured for them.
  cpu_baseline  decode: the oracle (C++ port of the reference decoder; the C# reference cannot run in this image), with
                libzstd 1.5.5's own decoder beside it as `cpu_baseline_libzstd`; compress: libzstd 1.5.5 (the reference
                has no compressor) — all host cores in every case
`--impl reference` times that CPU baseline alone, rank 0 only, as the reference arm.
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
METRIC = {"decode64k": "decompress_GBps_uncompressed", "compress128k": "compress_GBps_uncompressed"}


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
    ap.add_argument("--chunk", type=int, default=0, help="frame / chunk size in bytes (default: the workload's 64 KiB / 128 KiB)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def recorded_traffic(workload_key, kernel):
    """DRAM bytes per launch of `kernel` from the ncu --set full capture of this bench command (profiles/traffic.json,
    written by tools/ncu_traffic.py from the committed capture); None when no capture exists for this workload."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    rec = json.load(open(p)).get(workload_key, {}).get(kernel)
    return int(rec["dram_bytes"]) if rec else None


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
            time.sleep(0.01)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------------
# workload: host-side description of one step.  `src` is the codec's input, `dst` its output.
# ----------------------------------------------------------------------------------------------------------
def prepare(args, rank):
    from tools import corpus, zstd_ref
    import zstandard_b200 as zb
    chunk = args.chunk or CHUNK[args.workload]
    raw = corpus.make(args.corpus, args.bytes, shard=rank)
    total = len(raw)
    n = (total + chunk - 1) // chunk
    raw_off = np.arange(n + 1, dtype=np.uint64) * chunk
    raw_off[-1] = total
    w = {"raw": raw, "chunk": chunk, "n": n, "total": total, "kind": args.workload}
    blob, off = zstd_ref.compress_chunks(raw, chunk, level=args.level, checksum=True, threads=max(1, os.cpu_count() or 1))
    w["ref_compressed_bytes"] = int(off[-1])
    if args.workload == "decode64k":
        w.update(src=blob, src_off=off, dst_cap=np.diff(raw_off).astype(np.uint32), dst_stride=chunk)
    else:
        bound = (zb.ZStdCompress.CompressBound(chunk) + 15) // 16 * 16
        w.update(src=raw, src_off=raw_off, dst_cap=np.full(n, bound, dtype=np.uint32), dst_stride=bound)
    return w


def workload_config(args, w):
    if w["kind"] == "decode64k":
        what = (f"decode{w['chunk'] // 1024}k: batched decompression of {w['total']} B/GPU of libzstd-1.5.5 level-{args.level} {args.corpus} text held as "
                f"{w['chunk'] // 1024} KiB independent frames with XXH64 checksums")
    else:
        what = (f"compress{w['chunk'] // 1024}k: batched level-{args.level} compression of {w['total']} B/GPU of {args.corpus} text in "
                f"{w['chunk'] // 1024} KiB chunks, one frame each, with XXH64 checksums")
    return {"workload": what, "frames_per_gpu": w["n"], "libzstd_compressed_bytes_per_gpu": w["ref_compressed_bytes"],
            "libzstd_ratio": round(w["total"] / w["ref_compressed_bytes"], 4),
            "l2": "inputs+outputs per step exceed the 126 MB L2 (no flush needed)", "sharding": "by frame, no collective"}


class CpuBaseline:
    """All-core CPU arm: oracle batch decode (decode64k) or libzstd compress (compress128k)."""

    def __init__(self, args, w):
        self.w, self.args = w, args
        self.cores = os.cpu_count() or 1
        if w["kind"] == "decode64k":
            from tests import helpers
            self.lib = helpers.Oracle().lib
            n, total = w["n"], w["total"]
            self.out = np.zeros(total, dtype=np.uint8)
            off = w["src_off"]
            bs, bd = w["src"].ctypes.data, self.out.ctypes.data
            self.sp = (ctypes.c_void_p * n)(*[bs + int(off[i]) for i in range(n)])
            self.dp = (ctypes.c_void_p * n)(*[bd + i * w["chunk"] for i in range(n)])
            self.ss = np.diff(off).astype(np.uint32)
            self.dc = w["dst_cap"].copy()
            self.res = np.zeros(n, dtype=np.uint32)
            self.lib.oracle_decompress_batch.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_uint64, ctypes.c_int]
            self.lib.oracle_decompress_batch.restype = None
            self.kind, self.what = "port", "oracle = C++ restatement of the reference C# decoder, static frame partition over all cores"
            # second CPU baseline (SURVEY.md §8d ii): libzstd 1.5.5 ZSTD_decompress on all cores (tools/zstd_mt.c)
            self.zmt = None
            zp = os.path.join(ROOT, "tools", "_build", "libzstdmt.so")
            if os.path.exists(zp):
                self.zmt = ctypes.CDLL(zp)
                self.zmt.zmt_decompress.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64,
                                                    ctypes.c_uint64, ctypes.c_int]
        else:
            self.kind, self.what = "port", ("libzstd 1.5.5 ZSTD_compress2 (the reference ships no compressor; libzstd is the stand-in "
                                             "comparator), one context per core")

    def run(self):
        """-> (seconds, uncompressed bytes processed)"""
        w = self.w
        t = time.perf_counter()
        if w["kind"] == "decode64k":
            self.lib.oracle_decompress_batch(self.sp, self.ss.ctypes.data, self.dp, self.dc.ctypes.data, self.res.ctypes.data, w["n"], self.cores)
            dt = time.perf_counter() - t
            assert (self.res == self.dc).all(), "oracle failed to decode the workload"
        else:
            from tools import zstd_ref
            zstd_ref.compress_chunks(w["raw"], w["chunk"], level=self.args.level, checksum=True, threads=self.cores)
            dt = time.perf_counter() - t
        return dt, w["total"]


def libzstd_decode_baseline(cb):
    """-> {"value": GB/s, ...} of libzstd's own decoder on all cores over the same frames, or None."""
    if getattr(cb, "zmt", None) is None:
        return None
    w = cb.w
    off = np.ascontiguousarray(w["src_off"], dtype=np.uint64)
    call = lambda: cb.zmt.zmt_decompress(w["src"].ctypes.data, off.ctypes.data, w["n"], cb.out.ctypes.data, w["chunk"], w["total"], cb.cores)
    if call() != 0:
        return None
    tt, runs = 0.0, 0
    while tt < 1.5 and runs < 20:
        t = time.perf_counter()
        call()
        tt += time.perf_counter() - t
        runs += 1
    return {"value": round(w["total"] * runs / tt / 1e9, 3), "unit": "GB/s", "cores": cb.cores, "kind": "libzstd 1.5.5 ZSTD_decompress",
            "sample": f"{runs} pass(es) over the full step"}


def run_reference(args, rank):
    if rank != 0:
        return
    w = prepare(args, 0)
    cb = CpuBaseline(args, w)
    for _ in range(args.warmup):
        cb.run()
    t = 0.0
    for _ in range(args.steps):
        dt, nbytes = cb.run()
        t += dt
    gbs = nbytes * args.steps / t / 1e9
    line = {"impl": "reference", "metric": METRIC[args.workload], "value": round(gbs, 3), "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * t / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(args, w),
            "cpu_baseline": {"value": round(gbs, 3), "unit": "GB/s", "cores": cb.cores, "kind": cb.kind,
                             "sample": f"full step ({w['n']} frames, {nbytes} B) x {args.steps} steps; {cb.what}"},
            "e2e": {"value": round(gbs, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import zstandard_b200 as zb
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    w = prepare(args, rank)
    n, total, decode = w["n"], w["total"], w["kind"] == "decode64k"
    src_bytes = int(w["src_off"][-1])
    dst_span = n * w["dst_stride"]
    ctx = zb.Context(devices=[local], max_batch_bytes=max(total, 1 << 20))
    lib = zb.load_library()

    # ---------------- device-resident arm ----------------
    t_src = torch.empty(src_bytes + 64, dtype=torch.uint8, device=dev)
    t_src[:src_bytes] = torch.from_numpy(w["src"]).to(dev)
    t_dst = torch.zeros(dst_span + 64, dtype=torch.uint8, device=dev)
    soff = w["src_off"][:-1].astype(np.int64)
    ssz = np.diff(w["src_off"]).astype(np.int32)
    doff = np.arange(n, dtype=np.int64) * w["dst_stride"]
    dcap = w["dst_cap"].view(np.int32)
    t_soff, t_ssz = torch.from_numpy(soff).to(dev), torch.from_numpy(ssz).to(dev)
    t_doff, t_dcap = torch.from_numpy(doff).to(dev), torch.from_numpy(dcap.copy()).to(dev)
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
        if decode:
            ctx.decompress_batch_device(*dargs, stream=stream)
        else:
            ctx.compress_batch_device(args.level, True, *dargs, stream=stream)

    def verify(res_u32, out_bytes):
        """what is being timed must be right: decode -> the corpus; compress -> frames an independent decoder accepts"""
        if decode:
            assert (res_u32 == w["dst_cap"]).all(), "decode failed on the bench workload"
            assert (out_bytes[:total] == w["raw"]).all(), "decoded bytes differ from the corpus"
            return None
        from tools import zstd_ref
        assert (res_u32 < 0xFFFFFF88).all(), "compress failed on the bench workload"
        stride = w["dst_stride"]
        for k in range(0, n, max(1, n // 64)):
            f = out_bytes[k * stride:k * stride + int(res_u32[k])].tobytes()
            lo = k * w["chunk"]
            assert zstd_ref.decompress(f, int(ssz[k])) == w["raw"][lo:lo + int(ssz[k])].tobytes(), "frame does not round-trip"
        return int(res_u32.astype(np.int64).sum())

    for _ in range(max(3, args.warmup)):
        step_device()
    torch.cuda.synchronize()
    our_compressed = verify(t_res.cpu().numpy().view(np.uint32), t_dst.cpu().numpy())

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
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())

    # per-kernel device times (CUDA events between the kernels, separate synchronised runs)
    kms = {}
    reps = max(3, args.steps)
    if decode:
        for _ in range(reps):
            for k, v in ctx.decompress_batch_device_timed(*dargs, stream=stream).items():
                kms[k] = kms.get(k, 0.0) + v / reps
    else:
        for _ in range(reps):
            for k, v in ctx.compress_batch_device_timed(args.level, True, *dargs, stream=stream).items():
                kms[k] = kms.get(k, 0.0) + v / reps
    dom = max(kms, key=kms.get)
    out_bytes_algo = total if decode else our_compressed
    algo_bytes = src_bytes + out_bytes_algo
    peak, peak_src = measured_peak()
    achieved = algo_bytes / (kms[dom] * 1e-3) / 1e9

    # ---------------- end-to-end arm: host buffers through the C ABI ----------------
    h_src = lib.zstdb200_host_alloc(src_bytes + 64)
    h_dst = lib.zstdb200_host_alloc(dst_span + 64)
    ctypes.memmove(h_src, w["src"].ctypes.data, src_bytes)
    sp = (ctypes.c_void_p * n)(*[h_src + int(w["src_off"][i]) for i in range(n)])
    dp = (ctypes.c_void_p * n)(*[h_dst + i * w["dst_stride"] for i in range(n)])
    ss_u, dc_u, res_u = ssz.view(np.uint32).copy(), w["dst_cap"].copy(), np.zeros(n, dtype=np.uint32)
    u32p = ctypes.POINTER(ctypes.c_uint32)

    def step_e2e():
        if decode:
            rc = lib.zstdb200_decompress_batch(ctx.handle, sp, ss_u.ctypes.data_as(u32p), dp, dc_u.ctypes.data_as(u32p),
                                               res_u.ctypes.data_as(u32p), n)
        else:
            rc = lib.zstdb200_compress_batch(ctx.handle, args.level, 1, sp, ss_u.ctypes.data_as(u32p), dp, dc_u.ctypes.data_as(u32p),
                                             res_u.ctypes.data_as(u32p), n)
        assert rc == 0, ctx.last_error()

    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    got = np.ctypeslib.as_array(ctypes.cast(h_dst, ctypes.POINTER(ctypes.c_uint8)), shape=(dst_span,))
    verify(res_u, got)
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag = True      # clocks are sampled through both timed regions (device-resident and host-buffer)
    sampler.join()
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_s = float(t_e.item())
    lib.zstdb200_host_free(h_src)
    lib.zstdb200_host_free(h_dst)

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    cpu, cpu_libzstd = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = CpuBaseline(args, w)
        cb.run()
        tt, nb, runs = 0.0, 0, 0
        while tt < 3.0 and runs < 50:
            dt, b = cb.run()
            tt += dt; nb += b; runs += 1
        cpu = {"value": round(nb / tt / 1e9, 3), "unit": "GB/s", "cores": cb.cores, "kind": cb.kind,
               "sample": f"{runs} pass(es) over the full step ({n} frames, {total} B) = {tt * cb.cores:.1f} core-seconds; {cb.what}"}
        cpu_libzstd = libzstd_decode_baseline(cb) if decode else None

    if rank == 0:
        value = total * world * args.steps / (ms * 1e-3) / 1e9
        d2h = (total if decode else our_compressed) + n * 4
        line = {
            "metric": METRIC[args.workload], "value": round(value, 3), "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, w),
            "e2e": {"value": round(total * world * args.steps / e2e_s / 1e9, 3), "unit": "GB/s",
                    "h2d_bytes_per_step": int(src_bytes + n * 24), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 5),
                         "traffic": recorded_traffic(f"{args.workload}/{args.corpus}/{w['chunk']}/{total}/L{args.level}", dom), "kernel": dom, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(algo_bytes),
                         "pipeline_frac": round(algo_bytes * args.steps / (ms * 1e-3) / 1e9 / peak, 5)},
            "kernel_ms": {k: round(v, 4) for k, v in kms.items()},
            "cpu_baseline": cpu,
        }
        if cpu is not None and cpu_libzstd is not None:
            line["cpu_baseline_libzstd"] = cpu_libzstd
        if not decode:
            line["config"]["our_compressed_bytes_per_gpu"] = our_compressed
            line["config"]["our_ratio"] = round(total / our_compressed, 4)
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
