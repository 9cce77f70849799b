"""CPU: pins oracle/lz4_ref.c against the real upstream codec (system liblz4.so.1, what CodecLz4 wraps) and
the compression ratios printed in /root/reference/docs/src/index.md:52-63."""
import ctypes as C

import numpy as np
import pytest


def _sys(oracle):
    S = oracle.system_liblz4()
    if S is None:
        pytest.skip("liblz4.so.1 not present")
    return S


def _sys_compress(S, b, accel=2):
    cap = S.LZ4_compressBound(len(b))
    d = C.create_string_buffer(max(cap, 1))
    n = S.LZ4_compress_fast(b, d, len(b), cap, accel)
    return d.raw[:n]


def _sys_decompress(S, b, n):
    d = C.create_string_buffer(max(n, 1))
    return S.LZ4_decompress_safe(b, d, len(b), n), d.raw[:n]


def _bodies(oracle):
    rng = np.random.default_rng(11)
    N = 65536
    brands = ["apple", "samsung", "huawai", "microsoft", "dell", "xbox", "sony", "intel"]
    return {
        "id": np.arange(1, N + 1, dtype=np.int64).tobytes(),
        "rand1000": rng.integers(1, 1001, N).astype(np.int64).tobytes(),
        "brands": oracle.block_body("String", [brands[i] for i in rng.integers(0, 8, N)], 0, N),
        "price": (1 + 0.1 * rng.integers(0, 19991, N)).astype(np.float64).tobytes(),
        "f01": rng.random(N).tobytes(),
        "zeros": bytes(70000),
        "tiny": b"abc",
        "twelve": b"abcabcabcabc",
        "thirteen": b"aaaaaaaaaaaaa",
        "mixed": bytes(300) + rng.integers(0, 256, 999).astype(np.uint8).tobytes() + b"xyz" * 1000,
    }


def test_documented_compression_ratios(oracle):
    """docs/src/index.md:52-63: id 2.0, rand(1:1000) 2.55, 8 brands 2.85, rand(1.:0.1:2000.) 1.93 (accel 2, 65536-row blocks)."""
    b = _bodies(oracle)
    for name, want in [("id", 2.0), ("rand1000", 2.55), ("brands", 2.85), ("price", 1.93)]:
        for comp in (oracle.lz4_compress(b[name], 2), oracle.compress_block(b[name])):
            assert abs(len(b[name]) / len(comp) - want) < 0.02, (name, len(b[name]) / len(comp))


def test_round_trip_both_directions_with_upstream_liblz4(oracle):
    S = _sys(oracle)
    for name, body in _bodies(oracle).items():
        for accel in (1, 2, 8):
            own = oracle.lz4_compress(body, accel)
            up = _sys_compress(S, body, accel)
            assert oracle.lz4_decompress(own, len(body)) == body, name
            assert oracle.lz4_decompress(up, len(body)) == body, name           # restated decoder reads upstream streams
            n, out = _sys_decompress(S, own, len(body))
            assert n == len(body) and out == body, name                          # upstream decoder reads restated streams
            assert len(own) == len(up), name                                     # same greedy parse as LZ4_compress_fast


def test_corrupt_streams_are_rejected_like_upstream(oracle):
    S = _sys(oracle)
    rng = np.random.default_rng(5)
    body = rng.integers(1, 50, 5000).astype(np.int64).tobytes()
    good = oracle.compress_block(body)
    disagreements = 0
    for trial in range(300):
        bad = bytearray(good)
        k = int(rng.integers(0, len(bad)))
        bad[k] ^= int(rng.integers(1, 256))
        if trial % 3 == 0:
            bad = bad[: int(rng.integers(1, len(bad)))]
        n_up, out_up = _sys_decompress(S, bytes(bad), len(body))
        try:
            out = oracle.lz4_decompress(bytes(bad), len(body))
            ok = True
        except oracle.OracleError:
            ok = False
        if ok != (n_up == len(body)):
            disagreements += 1
        elif ok:
            assert out == out_up
    assert disagreements == 0


def test_block_stream_framing(oracle):
    """test/block_streams.jl:11-67: one compressed block of 64000 Int64 in 1..100000 reads back equal; a second
    block of another size follows a skipped first one."""
    import struct
    rng = np.random.default_rng(1)
    a = rng.integers(1, 100001, 64000).astype(np.int64)
    b = rng.integers(1, 100001, 74000).astype(np.int64)
    frames = []
    for arr in (a, b):
        body = arr.tobytes()
        comp = oracle.compress_block(body)
        frames.append(struct.pack("<iqq", len(arr), len(body), len(comp)) + comp)
    rows, body = oracle.decode_block(frames[0], a.nbytes)
    assert rows == 64000 and np.array_equal(np.frombuffer(body, np.int64), a)
    # skip_block: header + compressed bytes, then the next block decodes
    stream = frames[0] + frames[1]
    r0, _, c0 = struct.unpack("<iqq", stream[:20])
    assert r0 == 64000
    rows, body = oracle.decode_block(stream[20 + c0:], b.nbytes)
    assert rows == 74000 and np.array_equal(np.frombuffer(body, np.int64), b)
