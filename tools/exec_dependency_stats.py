#!/usr/bin/env python3
"""Dependency structure of the execute stage's work, from the real sequence records (analysis aid; CPU only).

The execute stage (k_exec, ExecSequence ZStdDecompress.cs:1265-1352) copies, per sequence, a literal run and a match
whose source may be bytes that an earlier sequence of the same frame produced.  This script decodes frames with the CPU
replay of the kernels' own entropy code (tests/hostsim: the records k_seq writes), and reports per frame

  sequences, bytes per sequence, share of matches whose source overlaps their own destination (offset < matchLength),
  depth  = length of the longest chain "match reads bytes written by an earlier match" (the number of rounds a
           level-synchronous execution needs),
  rounds32 = rounds the shipped scheme needs: records in groups of 32 in stream order, a match is ready once every
           earlier match of its group that overlaps its source has been written (k_exec's dependency rounds),
  level width = sequences / depth.

    python tools/exec_dependency_stats.py [--corpus log|tick] [--chunk 65536] [--frames 16] [--json out.json]
DESIGN.md sections 7 and 10 quote these numbers."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def records_of(hs, frame, cap):
    u8p, u32p = ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_uint32)
    src = np.frombuffer(frame, dtype=np.uint8)
    lit = np.zeros(cap + 64, dtype=np.uint8)
    max_recs = 2 * (cap // 3) + 64
    rec = np.zeros(4 * max_recs, dtype=np.uint32)
    info = np.zeros(8, dtype=np.uint32)
    f = hs.lib.hostsim_stages
    f.restype = ctypes.c_uint32
    f.argtypes = [u8p, ctypes.c_uint32, ctypes.c_uint32, u8p, u32p, ctypes.c_uint32, u32p]
    r = f(src.ctypes.data_as(u8p), src.size, cap, lit.ctypes.data_as(u8p), rec.ctypes.data_as(u32p), max_recs, info.ctypes.data_as(u32p))
    assert r == 0, hex(r)
    return rec.reshape(-1, 4)


def frame_stats(rec):
    """rec: the frame's records (block header records included).  Single- and multi-block frames alike."""
    i, out_base = 0, 0
    seqs = []                                     # (dst of match, offset, matchLength)
    while i < len(rec):
        count, out_bytes = int(rec[i, 0]), int(rec[i, 1])
        if count == 0 and out_bytes == 0 and int(rec[i, 2]) == 0:
            break
        for k in range(i + 1, i + 1 + count):
            x, y, z, w = (int(v) for v in rec[k])
            ll = w & 0x1FFFF
            ml = (w >> 17) | (((y >> 18) & 7) << 15)
            seqs.append((out_base + x + ll, z, ml))
        out_base += out_bytes
        i += 1 + count
    n = len(seqs)
    if n == 0:
        return None
    total = seqs[-1][0] + seqs[-1][2]
    # level of every output byte = level of the match that wrote it (literals: 0)
    level_of = np.zeros(total + 1, dtype=np.int32)
    depth, overl = 0, 0
    lv = np.zeros(n, dtype=np.int32)
    for s, (dst, off, ml) in enumerate(seqs):
        a = dst - off
        if off < ml:
            overl += 1
        if a < 0:
            a = 0
        src_hi = min(dst, a + ml)
        l = 1 + (int(level_of[a:src_hi].max()) if src_hi > a else 0)
        level_of[dst:dst + ml] = l
        lv[s] = l
        depth = max(depth, l)
    # the shipped scheme: groups of 32 in stream order; inside a group a match waits for the earlier matches of the group
    # whose destination overlaps its source
    rounds = 0
    for g in range(0, n, 32):
        grp = seqs[g:g + 32]
        rl = [0] * len(grp)
        for j, (dst, off, ml) in enumerate(grp):
            a, b = dst - off, min(dst, dst - off + ml)
            r = 1
            for q in range(j):
                d2, _, m2 = grp[q]
                if d2 < b and d2 + m2 > a:
                    r = max(r, rl[q] + 1)
            rl[j] = r
        rounds += max(rl)
    return {"sequences": n, "bytes_per_sequence": round(total / n, 2), "self_overlapping_pct": round(100.0 * overl / n, 1),
            "depth": int(depth), "level_width": round(n / depth, 1), "rounds32": rounds, "rounds32_per_group": round(rounds / ((n + 31) // 32), 2)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--corpus", default="log")
    ap.add_argument("--chunk", type=int, default=65536)
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    from tests import helpers
    from tools import corpus, zstd_ref
    hs = helpers.HostSim()
    raw = corpus.make(args.corpus, args.chunk * args.frames).tobytes()
    rows = []
    for k in range(args.frames):
        data = raw[k * args.chunk:(k + 1) * args.chunk]
        st = frame_stats(records_of(hs, zstd_ref.compress(data, 3, checksum=True), len(data)))
        if st:
            rows.append(st)
    keys = rows[0].keys()
    mean = {k: round(float(np.mean([r[k] for r in rows])), 2) for k in keys}
    out = {"corpus": args.corpus, "chunk": args.chunk, "frames": len(rows), "mean": mean, "min_depth": min(r["depth"] for r in rows),
           "max_depth": max(r["depth"] for r in rows)}
    print(json.dumps(out))
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"summary": out, "frames": rows}, f, indent=1)


if __name__ == "__main__":
    main()
