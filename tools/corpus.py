"""Deterministic synthetic corpora: log text, tick records, random, mixed.

`log/tick/random_/mixed` call the C generator (tools/corpus.c, built to tools/_build/libcorpus.so);
`py_log/py_tick/py_random` are a byte-identical pure-Python mirror for small sizes (the cross-check).
PRNG: the reference test suite's xorshift128+ (csharp/test/XorShift128Plus.cs:45-53).
Test/bench infrastructure only.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
M64 = (1 << 64) - 1


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libcorpus.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        for f in ("corpus_log", "corpus_tick", "corpus_random"):
            getattr(_LIB, f).argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64]
            getattr(_LIB, f).restype = None
        _LIB.corpus_mixed.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64]
        _LIB.corpus_mixed.restype = None
    return _LIB


def _gen(name, n, s0, s1):
    out = np.empty(n, dtype=np.uint8)
    if n:
        getattr(_lib(), name)(out.ctypes.data, n, s0, s1)
    return out


def log(n, seed=(42, 24)):
    return _gen("corpus_log", n, *seed)


def tick(n, seed=(43, 25)):
    return _gen("corpus_tick", n, *seed)


def random_(n, seed=(44, 26)):
    return _gen("corpus_random", n, *seed)


def mixed(n, chunk=65536, seed_shift=0):
    out = np.empty(n, dtype=np.uint8)
    if n:
        _lib().corpus_mixed(out.ctypes.data, n, chunk, seed_shift)
    return out


def make(kind, n, shard=0):
    """kind in {log,tick,random,mixed}; `shard` perturbs the seed so ranks get different data."""
    if kind == "log":
        return log(n, (42 + 1000 * shard, 24))
    if kind == "tick":
        return tick(n, (43 + 1000 * shard, 25))
    if kind == "random":
        return random_(n, (44 + 1000 * shard, 26))
    if kind == "mixed":
        return mixed(n, 65536, 1000 * shard)
    raise ValueError(kind)


# ------------------------------ pure-Python mirror ------------------------------
class XorShift128Plus:
    def __init__(self, s0, s1):
        self.s = [s0 & M64, s1 & M64]

    def next(self):
        x, y = self.s
        self.s[0] = y
        x ^= (x << 23) & M64
        self.s[1] = x ^ y ^ (x >> 17) ^ (y >> 26)
        return (self.s[1] + y) & M64


_LEVELS = ["INFO"] * 8 + ["DEBUG"] * 4 + ["WARN"] * 2 + ["ERROR", "TRACE"]
_COMPONENTS = ["http.server", "auth.session", "cache.lru", "net.pool", "sched.worker", "rpc.client", "runtime.gc",
               "storage.wal"]
_MDAYS = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]


def py_log(n, seed=(42, 24)):
    r = XorShift128Plus(*seed)
    out = bytearray()
    t_ms = 0
    while len(out) < n:
        t_ms += 1 + r.next() % 2000
        ms, s = t_ms % 1000, t_ms // 1000
        sec, mi, h, d = s % 60, (s // 60) % 60, (s // 3600) % 24, s // 86400
        year = 2026 + d // 365
        d %= 365
        mo = 0
        while d >= _MDAYS[mo]:
            d -= _MDAYS[mo]
            mo += 1
        line = "%04d-%02d-%02dT%02d:%02d:%02d.%03dZ [" % (year, mo + 1, d + 1, h, mi, sec, ms)
        line += _LEVELS[r.next() % 16] + "] "
        line += _COMPONENTS[r.next() % 8] + " - "
        a, b, c = r.next(), r.next(), r.next()
        k = r.next() % 7
        if k == 0:
            line += "request %d completed in %d ms status=%d" % (a % 1000000, b % 900, 200 if c % 8 else 500 + c % 4)
        elif k == 1:
            line += "user %d logged in from 10.%d.%d.%d" % (a % 50000, b % 256, (b >> 8) % 256, c % 256)
        elif k == 2:
            line += "cache miss for key item:%d shard=%d" % (a % 100000, b % 64)
        elif k == 3:
            line += "connection %d closed after %d bytes" % (a % 65536, b % 10000000)
        elif k == 4:
            line += "scheduled job %d took %d us queue_depth=%d" % (a % 4096, b % 250000, c % 128)
        elif k == 5:
            line += "retry %d/5 for upstream svc-%d" % (1 + a % 5, b % 32)
        else:
            line += "gc pause %d ms heap=%d MB" % (a % 120, 512 + b % 7680)
        out += line.encode("ascii") + b"\n"
    return bytes(out[:n])


def py_tick(n, seed=(43, 25)):
    import struct
    r = XorShift128Plus(*seed)
    out = bytearray()
    ts = 1767225600000000
    price = [100000 + 1000 * i for i in range(64)]
    while len(out) < n:
        ts += 1 + r.next() % 5000
        sym = r.next() % 64
        price[sym] += r.next() % 11 - 5
        qty = 100 * (1 + r.next() % 499)
        v = r.next()
        out += struct.pack("<QIiIHBB", ts, sym, price[sym], qty, v % 16, (v >> 8) & 1, 0)
    return bytes(out[:n])


def py_random(n, seed=(44, 26)):
    r = XorShift128Plus(*seed)
    out = bytearray()
    while len(out) < n:
        out += r.next().to_bytes(8, "little")
    return bytes(out[:n])
