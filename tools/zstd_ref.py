"""ctypes view of the system libzstd 1.5.5 (facebook/zstd) — shared object only, prototypes declared by hand.

Role: the reference repo has no compressor, so libzstd is (i) the producer of
"reference-compressed frames" for decode tests/bench, (ii) the ratio comparator for the GPU encoder and
(iii) a second decoder to cross-check the oracle. Test/bench infrastructure only.
"""
import ctypes
import threading

import numpy as np

_Z = ctypes.CDLL("libzstd.so.1")
_Z.ZSTD_compressBound.restype = ctypes.c_size_t
_Z.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
_Z.ZSTD_isError.restype = ctypes.c_uint
_Z.ZSTD_isError.argtypes = [ctypes.c_size_t]
_Z.ZSTD_createCCtx.restype = ctypes.c_void_p
_Z.ZSTD_freeCCtx.argtypes = [ctypes.c_void_p]
_Z.ZSTD_CCtx_setParameter.restype = ctypes.c_size_t
_Z.ZSTD_CCtx_setParameter.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
_Z.ZSTD_compress2.restype = ctypes.c_size_t
_Z.ZSTD_compress2.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
_Z.ZSTD_decompress.restype = ctypes.c_size_t
_Z.ZSTD_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
_Z.ZSTD_getFrameContentSize.restype = ctypes.c_uint64
_Z.ZSTD_getFrameContentSize.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
_Z.ZSTD_versionNumber.restype = ctypes.c_uint

_Z.ZSTD_CCtx_loadDictionary.restype = ctypes.c_size_t
_Z.ZSTD_CCtx_loadDictionary.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
ZSTD_c_compressionLevel = 100
ZSTD_c_windowLog = 101
ZSTD_c_contentSizeFlag = 200
ZSTD_c_checksumFlag = 201


def version():
    return _Z.ZSTD_versionNumber()


def compress_bound(n):
    return _Z.ZSTD_compressBound(n)


def compress(data, level=3, checksum=True, content_size=True, window_log=None):
    data = bytes(data) if not isinstance(data, (bytes, bytearray)) else data
    c = _Z.ZSTD_createCCtx()
    try:
        _Z.ZSTD_CCtx_setParameter(c, ZSTD_c_compressionLevel, level)
        _Z.ZSTD_CCtx_setParameter(c, ZSTD_c_checksumFlag, 1 if checksum else 0)
        _Z.ZSTD_CCtx_setParameter(c, ZSTD_c_contentSizeFlag, 1 if content_size else 0)
        if window_log is not None:
            _Z.ZSTD_CCtx_setParameter(c, ZSTD_c_windowLog, window_log)
        cap = _Z.ZSTD_compressBound(len(data))
        buf = ctypes.create_string_buffer(cap)
        n = _Z.ZSTD_compress2(c, buf, cap, bytes(data), len(data))
        if _Z.ZSTD_isError(n):
            raise RuntimeError("ZSTD_compress2 failed")
        return buf.raw[:n]
    finally:
        _Z.ZSTD_freeCCtx(c)


def decompress(frame, cap):
    buf = ctypes.create_string_buffer(max(cap, 1))
    n = _Z.ZSTD_decompress(buf, cap, bytes(frame), len(frame))
    if _Z.ZSTD_isError(n):
        return None
    return buf.raw[:n]


def compress_chunks(raw, chunk, level=3, checksum=True, threads=8, dictionary=None):
    """Compress raw (np.uint8 array) as independent frames of `chunk` bytes (each with `dictionary` when given).

    Returns (blob np.uint8, offsets np.uint64[n+1]) with frame i = blob[offsets[i]:offsets[i+1]]."""
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    n = (len(raw) + chunk - 1) // chunk
    bound = _Z.ZSTD_compressBound(chunk)
    tmp = np.empty(n * bound, dtype=np.uint8)
    sizes = np.zeros(n, dtype=np.uint64)

    def work(t):
        c = _Z.ZSTD_createCCtx()
        _Z.ZSTD_CCtx_setParameter(c, ZSTD_c_compressionLevel, level)
        _Z.ZSTD_CCtx_setParameter(c, ZSTD_c_checksumFlag, 1 if checksum else 0)
        if dictionary:
            assert not _Z.ZSTD_isError(_Z.ZSTD_CCtx_loadDictionary(c, dictionary, len(dictionary)))   # sticks to the context
        for i in range(t, n, threads):
            lo = i * chunk
            ln = min(chunk, len(raw) - lo)
            r = _Z.ZSTD_compress2(c, tmp.ctypes.data + i * bound, bound, raw.ctypes.data + lo, ln)
            assert not _Z.ZSTD_isError(r)
            sizes[i] = r
        _Z.ZSTD_freeCCtx(c)

    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    offsets = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(sizes, out=offsets[1:])
    blob = np.empty(int(offsets[-1]), dtype=np.uint8)
    for i in range(n):
        blob[int(offsets[i]):int(offsets[i + 1])] = tmp[i * bound:i * bound + int(sizes[i])]
    return blob, offsets


# ---- dictionaries (libzstd + ZDICT): producers of dictionary frames for the 8f-3 tests ----
_Z.ZDICT_trainFromBuffer.restype = ctypes.c_size_t
_Z.ZDICT_trainFromBuffer.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint]
_Z.ZDICT_isError.restype = ctypes.c_uint
_Z.ZDICT_isError.argtypes = [ctypes.c_size_t]
_Z.ZSTD_CCtx_loadDictionary.restype = ctypes.c_size_t
_Z.ZSTD_CCtx_loadDictionary.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
_Z.ZSTD_createDCtx.restype = ctypes.c_void_p
_Z.ZSTD_freeDCtx.argtypes = [ctypes.c_void_p]
_Z.ZSTD_decompress_usingDict.restype = ctypes.c_size_t
_Z.ZSTD_decompress_usingDict.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
ZSTD_c_dictIDFlag = 202


def train_dict(samples, capacity=16384):
    """ZDICT_trainFromBuffer over a list of byte strings -> a dictionary with entropy tables (magic 0xEC30A437)."""
    blob = b"".join(samples)
    sizes = (ctypes.c_size_t * len(samples))(*[len(s) for s in samples])
    buf = ctypes.create_string_buffer(capacity)
    n = _Z.ZDICT_trainFromBuffer(buf, capacity, blob, sizes, len(samples))
    if _Z.ZDICT_isError(n):
        raise RuntimeError("ZDICT_trainFromBuffer failed")
    return buf.raw[:n]


def compress_with_dict(data, dictionary, level=3, checksum=True, dict_id_flag=True):
    data = bytes(data)
    c = _Z.ZSTD_createCCtx()
    try:
        _Z.ZSTD_CCtx_setParameter(c, ZSTD_c_compressionLevel, level)
        _Z.ZSTD_CCtx_setParameter(c, ZSTD_c_checksumFlag, 1 if checksum else 0)
        _Z.ZSTD_CCtx_setParameter(c, ZSTD_c_dictIDFlag, 1 if dict_id_flag else 0)
        r = _Z.ZSTD_CCtx_loadDictionary(c, dictionary, len(dictionary))
        if _Z.ZSTD_isError(r):
            raise RuntimeError("ZSTD_CCtx_loadDictionary failed")
        cap = _Z.ZSTD_compressBound(len(data))
        buf = ctypes.create_string_buffer(cap)
        n = _Z.ZSTD_compress2(c, buf, cap, data, len(data))
        if _Z.ZSTD_isError(n):
            raise RuntimeError("ZSTD_compress2 failed")
        return buf.raw[:n]
    finally:
        _Z.ZSTD_freeCCtx(c)


def decompress_with_dict(frame, cap, dictionary):
    d = _Z.ZSTD_createDCtx()
    try:
        buf = ctypes.create_string_buffer(max(cap, 1))
        n = _Z.ZSTD_decompress_usingDict(d, buf, cap, bytes(frame), len(frame), dictionary, len(dictionary))
        if _Z.ZSTD_isError(n):
            return None
        return buf.raw[:n]
    finally:
        _Z.ZSTD_freeDCtx(d)
