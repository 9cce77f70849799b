"""The lane-per-block LZ4 decoder's state machine (csrc/lz4_lane_core.cuh) on the CPU.

tests/native/lane_sim.cpp runs ONE lane of lz4_decode_lane.cu with the kernel's own schedule (steps, piece queue, flush of
128-byte units, window refill that lands a round later) over the same header the kernel compiles, so the parse / emit logic,
the ring and window arithmetic and the acceptance rules of LZ4_decompress_safe (the call the reference makes in read_block,
/root/reference/src/io/BlockStreams.jl:110-112) are pinned against the oracle codec without a GPU.  The GPU tests
(test_gpu_parity.py, variant "lane") then cover the warp-cooperative parts."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "lane_sim.cpp")
CORE = os.path.join(HERE, "..", "dataframedbs.jl_b200", "csrc", "lz4_lane_core.cuh")
SO = os.path.join(HERE, "native", "liblane_sim.so")


@pytest.fixture(scope="module")
def sim():
    if not os.path.exists(SO) or max(os.path.getmtime(SRC), os.path.getmtime(CORE)) > os.path.getmtime(SO):
        subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-shared", "-fPIC", "-o", SO, SRC], check=True)
    L = C.CDLL(SO)
    L.lane_sim_decode2.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_long), C.c_int]

    def dec(comp, origin, hot_period=1):
        """hot_period 1: every step is a full step (general columns); 8: the hot-step schedule of word-regular columns"""
        out = C.create_string_buffer(max(origin, 1))
        st = (C.c_long * 8)()
        rc = L.lane_sim_decode2(comp, len(comp), out, origin, st, hot_period)
        return rc, out.raw[:origin], list(st)
    return dec


def _bodies(oracle):
    rng = np.random.default_rng(7)
    N = 65536
    brands = ["apple", "samsung", "huawai", "microsoft", "dell", "xbox", "sony", "intel"]
    return {
        "rand100": rng.integers(1, 101, N).astype(np.int64).tobytes(),
        "rand1000": rng.integers(1, 1001, N).astype(np.int64).tobytes(),
        "iseq": np.arange(1, N + 1, dtype=np.int64).tobytes(),
        "price": (1 + 0.1 * rng.integers(0, 19991, N)).astype(np.float64).tobytes(),
        "f01": rng.random(N).tobytes(),
        "zeros": bytes(100000), "short3": b"abc", "len13": b"0123456789abc", "len1": b"x",
        "rle_then_random": bytes(5000) + rng.integers(0, 256, 7001).astype(np.uint8).tobytes() + b"ab" * 3000,
        "brands": oracle.block_body("String", [brands[i] for i in rng.integers(0, 8, N)], 0, N),
        "missing_int": oracle.block_body("Missing(Int64)", (rng.integers(1, 101, N).astype(np.int64), rng.random(N) < 0.1), 0, N),
        "missing_float": oracle.block_body("Missing(Float64)", (rng.random(N), rng.random(N) < 0.1), 0, N),
        "decimals": oracle.block_body("String", [str(int(v)) for v in rng.integers(-2**31, 2**31, N // 4)], 0, N // 4),
        "repeats": b"".join(rng.integers(0, 256, 700).astype(np.uint8).tobytes() * 3 for _ in range(40)),
        "small_ints": rng.integers(0, 4, 30011).astype(np.uint8).tobytes(),
        "period3": (b"abc" * 20000)[:50001],
        "odd_sizes": rng.integers(0, 2, 777).astype(np.uint8).tobytes(),
    }


def test_every_body_kind_decodes_to_the_reference_bytes(sim, oracle):
    for name, body in _bodies(oracle).items():
        for comp in (oracle.compress_block(body), oracle.lz4_compress(body, 1)):
            rc, out, st = sim(comp, len(body))
            assert rc == 0 and out == body, (name, rc)
            # pieces of at most 8 output bytes: their number stays within 2.5 per output word
            assert st[1] <= 2.5 * (len(body) / 8) + 16, (name, st)
            rc, out, st = sim(comp, len(body), 8)          # the hot-step schedule must decode anything too, only slower
            assert rc == 0 and out == body, (name, "hot", rc)
    # word-regular bodies take the two-plain-tokens hot path for nearly every piece
    body = _bodies(oracle)["rand100"]
    rc, out, st = sim(oracle.compress_block(body), len(body), 8)
    assert rc == 0 and st[4] >= 0.98 * st[1], st


def test_empty_and_degenerate_blocks(sim, oracle):
    assert sim(b"\x00", 0)[0] == 0                     # what LZ4 emits for an empty body
    assert sim(b"\x00", 1)[0] != 0
    assert sim(b"\x10a", 0)[0] != 0
    assert sim(b"", 5)[0] != 0
    for body in (b"a", b"ab" * 3, bytes(12), bytes(13)):
        comp = oracle.compress_block(body)
        rc, out, _ = sim(comp, len(body))
        assert rc == 0 and out == body


def test_accepts_and_rejects_exactly_what_lz4_decompress_safe_does(sim, oracle):
    """@assert size == sizes.origin "decompression error" (BlockStreams.jl:112): damaged streams, truncated streams and wrong
    `origin` values are refused exactly when the CPU codec refuses them; accepted ones decode to the same bytes."""
    rng = np.random.default_rng(23)
    brands = ["apple", "samsung", "huawai", "microsoft", "dell", "xbox", "sony", "intel"]
    bodies = [rng.integers(1, 101, 4096).astype(np.int64).tobytes(),
              oracle.block_body("String", [brands[i] for i in rng.integers(0, 8, 4096)], 0, 4096),
              oracle.block_body("Missing(Float64)", (rng.random(4096), rng.random(4096) < 0.1), 0, 4096)]
    nacc = nrej = 0
    for body in bodies:
        good = oracle.compress_block(body)
        for _ in range(400):
            bad = bytearray(good)
            for _ in range(int(rng.integers(1, 4))):
                bad[int(rng.integers(0, len(bad)))] = int(rng.integers(0, 256))
            if rng.random() < 0.2:
                bad = bad[: int(rng.integers(1, len(bad)))]
            bad = bytes(bad)
            try:
                ref = oracle.lz4_decompress(bad, len(body))
            except oracle.OracleError:
                ref = None
            for hp in (1, 8):
                rc, out, _ = sim(bad, len(body), hp)
                if ref is None:
                    assert rc != 0
                else:
                    assert rc == 0 and out == ref
            nrej += ref is None
            nacc += ref is not None
        for d in (-8, -1, 1, 8, 100):
            try:
                oracle.lz4_decompress(good, len(body) + d)
                accepted = True
            except oracle.OracleError:
                accepted = False
            assert (sim(good, len(body) + d)[0] == 0) == accepted
    assert nacc > 100 and nrej > 100
