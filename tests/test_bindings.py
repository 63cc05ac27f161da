"""The Java-side surface of the reference: `ZstdDecompressor.decompress(byte[],int,int,byte[],int,int)`
and `getDecompressedSize` (java/src/main/java/com/epam/deltix/zstd/ZstdDecompressor.java:22-33).

No JDK exists in this image, so the JNI glue (bindings/java/zstdb200_jni.c) is compiled against a stand-in jni.h and
driven through a JNIEnv made of C arrays (tests/jni_stub/): that checks its argument handling, that no JNI call is made
inside a critical region, and — on the GPU box — that the reference's Java golden vector decodes through it."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from tests import helpers

ROOT = helpers.ROOT


class _Arr(ctypes.Structure):
    _fields_ = [("len", ctypes.c_int32), ("kind", ctypes.c_int), ("data", ctypes.c_void_p), ("pinned", ctypes.c_int)]


@pytest.fixture(scope="module")
def glue():
    import zstandard_b200 as zb
    zb.load_library()
    out = os.path.join(ROOT, "tests", "jni_stub", "_build", "libzstdb200_jni_test.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["gcc", "-std=c11", "-Wall", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "tests", "jni_stub"), "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "bindings", "java", "zstdb200_jni.c"), os.path.join(ROOT, "tests", "jni_stub", "fake_env.c"), "-o", out,
                    "-L" + os.path.join(ROOT, "zstandard_b200"), "-l:libzstdb200.so", "-Wl,--no-undefined",
                    "-Wl,-rpath," + os.path.join(ROOT, "zstandard_b200")], check=True)
    lib = ctypes.CDLL(out)
    lib.fake_env.restype = ctypes.c_void_p
    lib.fake_array.restype = ctypes.POINTER(_Arr); lib.fake_array.argtypes = [ctypes.c_int, ctypes.c_int32, ctypes.c_void_p]
    lib.fake_thrown.restype = ctypes.c_char_p
    lib.fake_string.restype = ctypes.c_char_p; lib.fake_string.argtypes = [ctypes.c_void_p]
    P = "Java_com_epam_deltix_zstd_ZstdDecompressor_"
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    for name, res, args in (("create0", i64, [vp, vp, i64]), ("destroy0", None, [vp, vp, i64]), ("isError0", ctypes.c_uint8, [vp, vp, i32]),
                            ("errorName0", vp, [vp, vp, i32]), ("lastError0", vp, [vp, vp, i64]),
                            ("decompress0", i32, [vp, vp, i64, vp, i32, i32, vp, i32, i32]),
                            ("decompressBatch0", i32, [vp, vp, i64, vp, vp, vp, vp, vp, vp, vp]),
                            ("getDecompressedSize0", i64, [vp, vp, vp, i32, i32])):
        f = getattr(lib, P + name); f.restype = res; f.argtypes = args
        setattr(lib, name, f)
    return lib


def _bytes(lib, data):
    buf = np.frombuffer(bytearray(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    return lib.fake_array(0, buf.size, buf.ctypes.data), buf


def test_glue_exports_and_header_queries(glue):
    env = glue.fake_env()
    for name, frame, raw in helpers.golden_vectors():
        a, keep = _bytes(glue, b"xx" + frame)
        assert glue.getDecompressedSize0(env, None, a, 2, len(frame)) == len(raw), name
    # no content-size field -> -1; bad magic / short input -> RuntimeException (ZstdFrameDecompressor.java:922-940)
    from tools import zstd_ref
    f = zstd_ref.compress(b"hello world" * 10, 3, checksum=False, content_size=False)
    a, keep = _bytes(glue, f)
    assert glue.getDecompressedSize0(env, None, a, 0, len(f)) == -1
    glue.fake_clear()
    a, keep = _bytes(glue, b"\x00" * 16)
    glue.getDecompressedSize0(env, None, a, 0, 16)
    assert glue.fake_thrown() == b"Invalid magic prefix"
    glue.fake_clear()
    glue.getDecompressedSize0(env, None, a, 0, 3)
    assert glue.fake_thrown() == b"Not enough input bytes"
    assert glue.isError0(env, None, -70) == 1 and glue.isError0(env, None, 100) == 0
    assert glue.fake_string(glue.errorName0(env, None, -20)) == b"corruption_detected"
    assert glue.fake_violations() == 0


def test_python_mirror_of_the_java_class_header_queries():
    import zstandard_b200 as zb
    for name, frame, raw in helpers.golden_vectors():
        assert zb.ZstdDecompressor.getDecompressedSize(b"abc" + frame, 3, len(frame)) == len(raw)
    with pytest.raises(RuntimeError, match="Invalid magic prefix"):
        zb.ZstdDecompressor.getDecompressedSize(bytes(16), 0, 16)
    with pytest.raises(RuntimeError, match="Not enough input bytes"):
        zb.ZstdDecompressor.getDecompressedSize(bytes(16), 0, 2)
    assert zb.ZstdDecompressor().decompress(b"", 0, 0, bytearray(4), 0, 0) == 0


@pytest.mark.gpu
def test_java_golden_vector_through_the_jni_glue_and_the_mirror(glue, oracle):
    import zstandard_b200 as zb
    env = glue.fake_env()
    ctx = glue.create0(env, None, 16 << 20)
    assert ctx
    try:
        items = [(f, r) for _, f, r in helpers.golden_vectors()] + helpers.make_frames(41, 30, sizes=[0, 1, 100, 5000, 70000, 200000])
        # one call per frame: java/src/test/java/com/epam/deltix/zstd/TestDecompress.java's shape, with offsets
        for frame, raw in items:
            a, ka = _bytes(glue, b"\x01\x02\x03" + frame)
            out = np.zeros(len(raw) + 9, dtype=np.uint8)
            o, ko = _bytes(glue, out)
            r = glue.decompress0(env, None, ctx, a, 3, len(frame), o, 5, len(raw))
            assert r == len(raw) and out[5:5 + len(raw)].tobytes() == raw and not out[:5].any() and not out[5 + len(raw):].any()
            assert a.contents.pinned == 0 and o.contents.pinned == 0
        # the batched overload, with one damaged frame in the middle
        n = len(items)
        frames = [bytearray(f) for f, _ in items]
        frames[3][len(frames[3]) // 2] ^= 0x40
        ins = [_bytes(glue, bytes(f)) for f in frames]
        outs = [_bytes(glue, np.zeros(max(len(r), 1), dtype=np.uint8)) for _, r in items]
        ptr_t = ctypes.c_void_p * n
        ia = ptr_t(*[ctypes.cast(a, ctypes.c_void_p).value for a, _ in ins]); oa = ptr_t(*[ctypes.cast(a, ctypes.c_void_p).value for a, _ in outs])
        zeros = np.zeros(n, dtype=np.int32); ilen = np.array([len(f) for f in frames], dtype=np.int32); mlen = np.array([len(r) for _, r in items], dtype=np.int32)
        res = np.zeros(n, dtype=np.int32)
        mk = lambda v: glue.fake_array(1, n, v.ctypes.data)
        rc = glue.decompressBatch0(env, None, ctx, glue.fake_array(2, n, ctypes.addressof(ia)), mk(zeros), mk(ilen),
                                   glue.fake_array(2, n, ctypes.addressof(oa)), mk(zeros), mk(mlen), mk(res))
        assert rc == 0
        for k, ((frame, raw), (_, ob)) in enumerate(zip(items, outs)):
            want, _, _ = oracle.decompress(bytes(frames[k]), len(raw))
            assert int(res[k]) & 0xFFFFFFFF == want, k
            if not helpers.is_err(want):
                assert ob[:want].tobytes() == raw
        assert glue.fake_violations() == 0           # no JNI call inside a critical region
    finally:
        glue.destroy0(env, None, ctx)
    # the Python mirror of the Java class: same contract
    dec = zb.ZstdDecompressor()
    for frame, raw in items[:6]:
        out = bytearray(len(raw) + 4)
        assert dec.decompress(b"zz" + frame, 2, len(frame), out, 4, len(raw)) == len(raw) and bytes(out[4:]) == raw
    with pytest.raises(RuntimeError, match="corruption_detected|checksum_wrong|srcSize_wrong|dstSize_tooSmall"):
        dec.decompress(bytes(frames[3]), 0, len(frames[3]), bytearray(len(items[3][1])), 0, len(items[3][1]))


# ---- C# (P/Invoke) surface: bindings/csharp/ZStdB200.cs -------------------------------------------------
# No .NET SDK exists in this image either.  The shim is thin (pin, call, return the code), so what can go wrong without a
# compiler is the marshalling: a wrong entry-point name, argument count, or argument width.  This checks every
# [DllImport] declaration against the prototype in include/zstdb200.h.
_C_WIDTH = {"int": "i32", "uint32_t": "u32", "uint64_t": "u64", "size_t": "ptr", "void": "void"}
_CS_WIDTH = {"int": "i32", "uint": "u32", "ulong": "u64", "UIntPtr": "ptr", "IntPtr": "ptr", "void": "void"}


def _c_kind(t):
    t = t.replace("const", " ").strip()
    if "*" in t:
        return "ptr"
    return _C_WIDTH[t.split()[0]]


def _cs_kind(t):
    t = t.replace("out ", " ").strip() if not t.startswith("out ") else "IntPtr*"
    if "*" in t:
        return "ptr"
    return _CS_WIDTH[t.split()[0]]


def _header_prototypes():
    import re
    text = open(os.path.join(ROOT, "include", "zstdb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z_0-9 \*]*?)\b(zstdb200_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        kinds = []
        for prm in params:
            ty = prm.rsplit(" ", 1)[0] if not prm.endswith("*") else prm        # drop the parameter name
            if "*" in prm:
                ty = prm
            kinds.append(_c_kind(ty))
        protos[name] = (_c_kind(ret), kinds)
    return protos


def test_csharp_pinvoke_declarations_match_the_header():
    import re
    src = open(os.path.join(ROOT, "bindings", "csharp", "ZStdB200.cs")).read()
    decls = re.findall(r"\[DllImport\(Lib\)\]\s*internal static extern\s+([A-Za-z\*]+)\s+(zstdb200_[a-z_0-9]+)\(([^)]*)\);", src)
    assert len(decls) >= 14
    protos = _header_prototypes()
    import zstandard_b200 as zb
    lib = zb.load_library()
    for ret, name, args in decls:
        assert name in protos and hasattr(lib, name), name
        params = [a.strip() for a in args.split(",")] if args.strip() else []
        kinds = []
        for prm in params:
            ty = prm.rsplit(" ", 1)[0]
            kinds.append("ptr" if (ty.startswith("out ") or "*" in ty) else _CS_WIDTH[ty])
        c_ret, c_kinds = protos[name]
        assert _cs_kind(ret) == c_ret, (name, ret, c_ret)
        assert kinds == c_kinds, (name, kinds, c_kinds)
    # the public surface keeps the reference's method names and adds the batched overloads (ZStdDecompress.cs:590-607, 2182-2191)
    for sig in ("public static ulong GetDecompressedSize(byte[] src)", "public static ulong GetDecompressedSize(byte[] src, uint srcSize)",
                "public static uint Decompress(byte[] dst, uint dstCapacity, byte[] src, uint srcSize)", "public static uint Decompress(byte[] dst, byte[] src)",
                "public static void Decompress(IReadOnlyList<ArraySegment<byte>> srcs, IReadOnlyList<ArraySegment<byte>> dsts, uint[] results)",
                "public static void Compress(IReadOnlyList<ArraySegment<byte>> srcs, IReadOnlyList<ArraySegment<byte>> dsts, uint[] results, int level = 3, bool checksum = true)"):
        assert sig in src, sig
    proj = open(os.path.join(ROOT, "bindings", "csharp", "Zstandard.B200.csproj")).read()
    assert "<TargetFramework>netstandard2.0</TargetFramework>" in proj and "NativeMemory." not in src and "Span<" not in src
