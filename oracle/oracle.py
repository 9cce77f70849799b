"""ctypes binding of the CPU oracle (oracle/liboracle.so) -- TEST INFRASTRUCTURE ONLY.

Also holds the fixture writer that produces tables in the reference's on-disk format
(/root/reference/src/io/table_io.jl:9-19, filesystem.jl:14-23, BlockStreams.jl:36-60,
blocks.jl:2-33) from numpy / Python data, the way `create_table(path; from=df, block_size=...)`
does in the reference's tests (test/view.jl:9-15).
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

K_NAMES = {1: "Int8", 2: "Int16", 3: "Int32", 4: "Int64", 5: "Int128", 6: "UInt8", 7: "UInt16", 8: "UInt32", 9: "UInt64",
           10: "UInt128", 11: "Float16", 12: "Float32", 13: "Float64", 14: "Bool", 15: "Char", 16: "String", 17: "Date",
           18: "DateTime", 19: "Time", 20: "Tuple"}
NP_DTYPES = {"Int8": np.int8, "Int16": np.int16, "Int32": np.int32, "Int64": np.int64, "UInt8": np.uint8, "UInt16": np.uint16,
             "UInt32": np.uint32, "UInt64": np.uint64, "Float16": np.float16, "Float32": np.float32, "Float64": np.float64,
             "Bool": np.bool_, "Char": np.uint32, "Date": np.int64, "DateTime": np.int64, "Time": np.int64}


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"oracle error {code}: {msg}")
        self.code = code
        self.msg = msg


class _Col(C.Structure):
    _fields_ = [("kind", C.c_int32), ("nullable", C.c_int32), ("elsize", C.c_int32), ("is_expr", C.c_int32),
                ("nrows", C.c_int64), ("values", C.c_void_p), ("missing", C.c_void_p), ("sizes", C.c_void_p),
                ("chars", C.c_void_p), ("nchars", C.c_int64), ("cap_rows", C.c_size_t), ("cap_chars", C.c_size_t)]


class Agg(C.Structure):
    _fields_ = [("count", C.c_int64), ("nmissing", C.c_int64), ("sum_i64", C.c_int64), ("sum_fold", C.c_double),
                ("sum_kahan", C.c_double), ("min_i64", C.c_int64), ("max_i64", C.c_int64), ("min_f64", C.c_double),
                ("max_f64", C.c_double), ("has_nan", C.c_int32), ("kind", C.c_int32)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("lz4_ref.c", "dfdb_oracle.c", "gen.c", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_last_error.restype = C.c_char_p
        L.orc_lz4_compress_bound.argtypes = [C.c_int]
        L.orc_lz4_compress.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_lz4_decompress_safe.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orc_compress_block.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_table_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.orc_table_close.argtypes = [C.c_void_p]
        L.orc_table_ncols.argtypes = [C.c_void_p]
        L.orc_table_ncols.restype = C.c_int64
        L.orc_table_block_size.argtypes = [C.c_void_p]
        L.orc_table_block_size.restype = C.c_int64
        L.orc_table_nrows.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.orc_table_col.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        L.orc_count.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.POINTER(C.c_int64)]
        L.orc_mask.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        L.orc_materialize.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.POINTER(C.c_void_p)]
        L.orc_mat_free.argtypes = [C.c_void_p]
        L.orc_mat_ncols.argtypes = [C.c_void_p]
        L.orc_mat_col.argtypes = [C.c_void_p, C.c_int]
        L.orc_mat_col.restype = C.POINTER(_Col)
        L.orc_aggregate.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_int32, C.POINTER(Agg)]
        L.orc_aggregate_blocks.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.POINTER(Agg)]
        L.orc_materialize_hash_blocks.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_uint64),
                                                  C.POINTER(C.c_uint64), C.POINTER(C.c_int64), C.c_int32]
        L.orc_hash_fixed.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_hash_fixed.restype = None
        L.orc_hash_strings.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_hash_strings.restype = None
        L.orc_decode_block.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int32)]
        L.orc_decode_block.restype = C.c_int64
        L.orc_parse_type.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_gen_table.argtypes = [C.c_char_p, C.c_char_p, C.c_int64, C.c_int64, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_int64)]
        L.orc_have_liblz4.restype = C.c_int
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise OracleError(rc, lib().orc_last_error().decode("utf-8", "replace"))


# ---- codec -----------------------------------------------------------------------------------

def lz4_compress(data: bytes, accel: int = 2) -> bytes:
    L = lib()
    cap = L.orc_lz4_compress_bound(len(data))
    dst = C.create_string_buffer(cap)
    n = L.orc_lz4_compress(data, dst, len(data), cap, accel)
    if n <= 0:
        raise OracleError(-1, "compress failed")
    return dst.raw[:n]


def lz4_decompress(data: bytes, origin: int) -> bytes:
    dst = C.create_string_buffer(max(origin, 1))
    n = lib().orc_lz4_decompress_safe(data, dst, len(data), origin)
    if n != origin:
        raise OracleError(n, "decompression error")
    return dst.raw[:origin]


def compress_block(body: bytes, prefer_system: bool = True) -> bytes:
    """The reference's codec call (LZ4_compress_fast, acceleration 2); system liblz4 when present."""
    L = lib()
    cap = L.orc_lz4_compress_bound(len(body))
    dst = C.create_string_buffer(cap)
    n = L.orc_compress_block(body, dst, len(body), cap, 1 if prefer_system else 0)
    if n <= 0:
        raise OracleError(-1, "compress failed")
    return dst.raw[:n]


def system_liblz4():
    """The real upstream codec (what CodecLz4 wraps), for pinning lz4_ref.c.  None when absent."""
    try:
        L = C.CDLL("liblz4.so.1")
    except OSError:
        return None
    L.LZ4_compress_fast.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
    L.LZ4_decompress_safe.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
    L.LZ4_compressBound.argtypes = [C.c_int]
    return L


# ---- fixture writer ----------------------------------------------------------------------------

def _jl_string(s: str) -> bytes:
    b = s.encode("utf-8")
    return struct.pack("<i", len(b)) + b


def block_body(typestring: str, data, lo: int, hi: int, garbage: int = 0x5A) -> bytes:
    """write_block_body blocks.jl:2-33 for rows [lo, hi)."""
    nullable = typestring.startswith("Missing(")
    base = typestring[8:-1] if nullable else typestring
    if base == "String":
        sizes, chars = [], []
        for s in data[lo:hi]:
            if s is None:
                sizes.append(-1)
            else:
                b = s.encode("utf-8") if isinstance(s, str) else bytes(s)
                sizes.append(len(b))
                chars.append(b)
        blob = b"".join(chars)
        return struct.pack("<i", len(blob)) + np.asarray(sizes, dtype="<i4").tobytes() + blob
    dt = np.dtype(NP_DTYPES[base]).newbyteorder("<")
    if nullable:
        values, missing = data
        v = np.array(values[lo:hi], dtype=dt, copy=True)
        m = np.asarray(missing[lo:hi], dtype=bool)
        rows = hi - lo
        bits = np.zeros((rows + 63) // 64 * 64, dtype=np.uint8)
        bits[:rows] = m
        words = np.packbits(bits.reshape(-1, 8), axis=1, bitorder="little").reshape(-1)   # BitArray chunks, LSB first
        # bytes under a missing bit are unspecified in the reference (src/common/missings.jl:1)
        vb = v.view(np.uint8).reshape(rows, dt.itemsize).copy()
        vb[m] = garbage
        return words.tobytes() + vb.tobytes()
    return np.ascontiguousarray(np.asarray(data[lo:hi], dtype=dt)).tobytes()


def write_table(path: str, columns, block_size: int = 65536, prefer_system_lz4: bool = True, ids=None) -> None:
    """columns: list of (name, typestring, data).  data: ndarray | (values, missing_mask) | list[str|None]."""
    os.makedirs(path, exist_ok=True)
    ids = ids or list(range(1, len(columns) + 1))
    with open(os.path.join(path, "meta.bin"), "wb") as f:
        f.write(struct.pack("<qqq", 1, block_size, len(columns)))
        for cid, (name, ts, _) in zip(ids, columns):
            f.write(struct.pack("<q", cid) + _jl_string(name) + _jl_string(ts))
    for cid, (name, ts, data) in zip(ids, columns):
        n = len(data[0]) if isinstance(data, tuple) else len(data)
        with open(os.path.join(path, f"{cid}.bin"), "wb") as f:
            f.write(struct.pack("<q", block_size) + _jl_string(ts))
            for lo in range(0, n, block_size):
                hi = min(n, lo + block_size)
                body = block_body(ts, data, lo, hi)
                if len(body) == 0:      # commit_block_write! skips empty bodies (BlockStreams.jl:38)
                    continue
                comp = compress_block(body, prefer_system_lz4)
                f.write(struct.pack("<iqq", hi - lo, len(body), len(comp)) + comp)


def gen_table(path: str, spec: str, nrows: int, block_size: int = 65536, seed: int = 0xDFDB0000, nthreads: int | None = None,
              prefer_system_lz4: bool = True):
    """Synthetic table by oracle/gen.c.  Returns (uncompressed_bytes, compressed_bytes)."""
    stats = (C.c_int64 * 2)()
    nthreads = nthreads or os.cpu_count() or 1
    rc = lib().orc_gen_table(path.encode(), spec.encode(), nrows, block_size, seed, nthreads, int(prefer_system_lz4), stats)
    if rc != 0:
        raise OracleError(rc, "gen_table failed")
    return stats[0], stats[1]


# ---- scans -------------------------------------------------------------------------------------

class FlatStrings:
    """FlatStringsVector layout: sizes (Int32, -1 = missing) + flat chars (FlatStringsVectors.jl:5-9)."""

    def __init__(self, sizes: np.ndarray, chars: bytes):
        self.sizes = sizes
        self.chars = chars

    def tolist(self):
        out, pos = [], 0
        for s in self.sizes.tolist():
            if s < 0:
                out.append(None)
            else:
                out.append(self.chars[pos:pos + s].decode("utf-8"))
                pos += s
        return out

    def __len__(self):
        return len(self.sizes)


class OracleTable:
    def __init__(self, path: str):
        self._h = C.c_void_p()
        _check(lib().orc_table_open(path.encode(), C.byref(self._h)))
        self.path = path

    def close(self):
        if self._h:
            lib().orc_table_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def block_size(self):
        return lib().orc_table_block_size(self._h)

    def nrows(self) -> int:
        n = C.c_int64()
        _check(lib().orc_table_nrows(self._h, C.byref(n)))
        return n.value

    def columns(self):
        out = []
        for i in range(lib().orc_table_ncols(self._h)):
            cid = C.c_int64()
            name = C.create_string_buffer(256)
            ts = C.create_string_buffer(256)
            _check(lib().orc_table_col(self._h, i, C.byref(cid), name, 256, ts, 256))
            out.append((cid.value, name.value.decode(), ts.value.decode()))
        return out

    def count(self, plan: bytes) -> int:
        n = C.c_int64()
        _check(lib().orc_count(self._h, plan, len(plan), C.byref(n)))
        return n.value

    def mask(self, plan: bytes) -> np.ndarray:
        n = self.nrows()
        buf = np.zeros(n + 1, dtype=np.uint8)
        seen = C.c_int64()
        _check(lib().orc_mask(self._h, plan, len(plan), buf.ctypes.data, n, C.byref(seen)))
        return buf[:n].astype(bool)

    def materialize(self, plan: bytes):
        """-> list of columns: ndarray | (ndarray values, ndarray missing) | FlatStrings."""
        h = C.c_void_p()
        _check(lib().orc_materialize(self._h, plan, len(plan), C.byref(h)))
        try:
            out = []
            for i in range(lib().orc_mat_ncols(h)):
                c = lib().orc_mat_col(h, i).contents
                kind = K_NAMES.get(c.kind, "Int64" if c.kind == 0 else None)
                n = c.nrows
                if kind == "String":
                    sizes = np.frombuffer(C.string_at(c.sizes, 4 * n), dtype="<i4").copy() if n else np.zeros(0, "<i4")
                    chars = C.string_at(c.chars, c.nchars) if c.nchars else b""
                    out.append(FlatStrings(sizes, chars))
                    continue
                if kind in NP_DTYPES:
                    dt = np.dtype(NP_DTYPES[kind])
                    vals = np.frombuffer(C.string_at(c.values, n * c.elsize), dtype=dt).copy() if n else np.zeros(0, dt)
                else:
                    vals = np.frombuffer(C.string_at(c.values, n * c.elsize), dtype=np.uint8).reshape(n, c.elsize).copy() if n else np.zeros((0, c.elsize), np.uint8)
                if c.nullable:
                    miss = np.frombuffer(C.string_at(c.missing, n), dtype=np.uint8).astype(bool) if n else np.zeros(0, bool)
                    out.append((vals, miss))
                else:
                    out.append(vals)
            return out
        finally:
            lib().orc_mat_free(h)

    def aggregate(self, plan: bytes, proj_idx: int = 0) -> Agg:
        a = Agg()
        _check(lib().orc_aggregate(self._h, plan, len(plan), proj_idx, C.byref(a)))
        return a

    def aggregate_blocks(self, plan: bytes, proj_idx: int, blk_lo: int, blk_hi: int) -> Agg:
        a = Agg()
        _check(lib().orc_aggregate_blocks(self._h, plan, len(plan), proj_idx, blk_lo, blk_hi, C.byref(a)))
        return a

    def materialize_hash_mt(self, plan: bytes, nblocks: int, nthreads: int, ncols: int):
        """Content hash per column of materialize(plan) over the whole table without holding the result: every thread
        materializes one block range, hashes it with range-local positions and frees it; the ranges are re-based in block
        order (A + 2 * base * B).  Returns ([(A, B)] per column, selected rows).  Predicate-only plans."""
        nthreads = max(1, min(nthreads, nblocks))
        bounds = [(nblocks * i) // nthreads for i in range(nthreads + 1)]

        def job(k):
            A, B, n = (C.c_uint64 * ncols)(), (C.c_uint64 * ncols)(), C.c_int64()
            _check(lib().orc_materialize_hash_blocks(self._h, plan, len(plan), bounds[k], bounds[k + 1], A, B, C.byref(n), ncols))
            return list(A), list(B), n.value
        with ThreadPoolExecutor(nthreads) as ex:
            parts = list(ex.map(job, range(nthreads)))
        out, base = [[0, 0] for _ in range(ncols)], 0
        for A, B, n in parts:
            for c in range(ncols):
                out[c][0] = (out[c][0] + A[c] + 2 * base * B[c]) & _M64
                out[c][1] = (out[c][1] + B[c]) & _M64
            base += n
        return [tuple(x) for x in out], base

    def aggregate_mt(self, plan: bytes, proj_idx: int, nblocks: int, nthreads: int, blk_lo: int = 0):
        """Thread-per-block-range driver over the single-threaded oracle (predicate-only plans).
        Returns the list of per-range partials in block order."""
        nthreads = max(1, min(nthreads, nblocks))
        bounds = [blk_lo + (nblocks * i) // nthreads for i in range(nthreads + 1)]
        with ThreadPoolExecutor(nthreads) as ex:
            futs = [ex.submit(self.aggregate_blocks, plan, proj_idx, bounds[i], bounds[i + 1]) for i in range(nthreads)]
            return [f.result() for f in futs]


_M64 = (1 << 64) - 1


def hash_columns(cols, nthreads: int = 1):
    """(A, B) content hash per column of a materialized result (see orc_hash_fixed in dfdb_oracle.c): `cols` is a list of
    numpy arrays (fixed width), (values, missing) pairs (nullable; missing = bool / uint8 per row) or (sizes, chars)
    pairs with sizes.dtype == int32 (strings).  Chunks are hashed on `nthreads` threads and re-based."""
    L = lib()
    out = []
    for col in cols:
        if isinstance(col, tuple) and len(col) == 3 and col[0] == "str":
            _, sizes, chars = col
            sizes = np.ascontiguousarray(sizes, dtype=np.int32)
            chars = np.ascontiguousarray(np.frombuffer(chars, dtype=np.uint8) if isinstance(chars, (bytes, bytearray, memoryview)) else chars, dtype=np.uint8)
            n = len(sizes)
            bounds = [n * i // nthreads for i in range(nthreads + 1)]
            coff = np.concatenate([[0], np.cumsum(np.maximum(sizes, 0), dtype=np.int64)]) if n else np.zeros(1, np.int64)

            def job(k, sizes=sizes, chars=chars, coff=coff, bounds=bounds):
                a, b = C.c_uint64(), C.c_uint64()
                lo, hi = bounds[k], bounds[k + 1]
                L.orc_hash_strings(sizes.ctypes.data + 4 * lo, chars.ctypes.data + int(coff[lo]), hi - lo, lo, C.byref(a), C.byref(b))
                return a.value, b.value
        else:
            if isinstance(col, tuple):
                vals, miss = col
                miss = np.ascontiguousarray(miss, dtype=np.uint8)
            else:
                vals, miss = col, None
            vals = np.ascontiguousarray(vals)
            n, es = len(vals), vals.dtype.itemsize
            bounds = [n * i // nthreads for i in range(nthreads + 1)]

            def job(k, vals=vals, miss=miss, es=es, bounds=bounds):
                a, b = C.c_uint64(), C.c_uint64()
                lo, hi = bounds[k], bounds[k + 1]
                L.orc_hash_fixed(vals.ctypes.data + es * lo, (miss.ctypes.data + lo) if miss is not None else None, es, hi - lo, lo, C.byref(a), C.byref(b))
                return a.value, b.value
        with ThreadPoolExecutor(max(1, nthreads)) as ex:
            parts = list(ex.map(job, range(nthreads)))
        out.append((sum(p[0] for p in parts) & _M64, sum(p[1] for p in parts) & _M64))
    return out


def decode_block(framed: bytes, cap: int):
    out = C.create_string_buffer(max(cap, 1))
    rows = C.c_int32()
    n = lib().orc_decode_block(framed, len(framed), out, cap, C.byref(rows))
    if n < 0:
        raise OracleError(int(n), "decompression error")
    return rows.value, out.raw[:n]
