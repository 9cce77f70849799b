"""The batch rules of the two warp-per-block decoders of round 2, as Python models (tests/models/), against the oracle's codec.

lz4_decode_spec.cu and lz4_decode_bytes.cu replace read_block's LZ4_decompress_safe call (src/io/BlockStreams.jl:101-119) with
decoders that VERIFY token positions instead of walking them.  The kernels are checked against the oracle on the GPU
(tests/test_gpu_parity.py: every body kind, ragged blocks, corrupt and fuzzed streams); these tests pin the rules themselves on
the CPU: the models decode real liblz4-format streams to exactly the original bytes, reading every source word / byte from
where the kernel would (the models assert that the ring holds what global memory holds), and they exercise each batch shape."""
import numpy as np
import pytest

from models.bytes_model import decode_block_bytes
from models.spec_model import decode_block_spec

N = 16384


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    oracle.build()
    return oracle


def _bodies(O):
    rng = np.random.default_rng(11)
    brands = ["apple", "samsung", "huawai", "microsoft", "dell", "xbox", "sony", "intel"]
    return {
        "rand100": rng.integers(1, 101, N).astype(np.int64).tobytes(),
        "rand1000": rng.integers(1, 1001, N).astype(np.int64).tobytes(),
        "sorted": np.arange(1, N + 1, dtype=np.int64).tobytes(),
        "rand4": rng.integers(1, 5, N).astype(np.int64).tobytes(),
        "missing_int": O.block_body("Missing(Int64)", (rng.integers(1, 101, N).astype(np.int64), rng.random(N) < 0.1), 0, N),
        "brands": O.block_body("String", [brands[i] for i in rng.integers(0, 8, N)], 0, N),
        "missing_brands": O.block_body("Missing(String)", [None if m else brands[i] for i, m in zip(rng.integers(0, 8, N), rng.random(N) < 0.1)], 0, N),
        "decimals": O.block_body("String", [str(int(v)) for v in rng.integers(-2**31, 2**31, N // 4)], 0, N // 4),
    }


def test_spec_model_decodes_word_regular_bodies(O):
    shapes = {}
    for name, body in _bodies(O).items():
        if name in ("brands", "missing_brands", "decimals"):
            continue
        for comp in (O.compress_block(body), O.lz4_compress(body, 1)):
            got, st = decode_block_spec(comp, len(body))
            assert got == body, name
            for k, v in st.items():
                shapes[k] = shapes.get(k, 0) + v
    # every batch shape of the kernel occurred: full, full + one two-word sequence, full + one (1, 7) sequence, run + closing sequence, one sequence
    assert all(shapes.get(k, 0) > 0 for k in ("fb", "fb1", "fb1b", "nfb", "one")), shapes


def test_bytes_model_decodes_byte_streams(O):
    total = {}
    for name, body in _bodies(O).items():
        comp = O.compress_block(body)
        got, st = decode_block_bytes(comp, len(body))
        assert got == body, name
        if name == "brands":
            # the stream this decoder is chosen for: nearly everything goes through batches, closing matches stay inside them
            assert st["batches"] > 10 * st["one"] and st["closing"] > 0, st
        for k, v in st.items():
            total[k] = total.get(k, 0) + v
    assert total["batches"] > 0 and total["one"] > 0 and total["rounds"] > 0
