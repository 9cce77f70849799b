"""CPU, world size 2, gloo: the N>1 host logic -- block-range shards from the C ABI, all-gather of the per-rank
dfdb_agg partials and the fixed rank-order fold.  The per-shard partials come from the oracle here (no GPU in this
container); on the GPU box tests/test_gpu_parity.py::test_sharded_scan_folds_to_the_unsharded_result covers the
same fold with partials produced by the CUDA kernels."""
import ctypes as C
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, path, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import dfdb_b200 as D
    from dfdb_b200 import _capi
    from dfdb_b200.dist import allgather_counts, allgather_fold, allreduce_count
    from oracle import oracle as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        t = D.open_table(path, rank=rank, world=world)          # host only: parses headers, sets the shard
        L = _capi.lib()
        lo, hi = C.c_int64(), C.c_int64()
        _capi.check(L.dfdb_table_shard_range(t._h, C.byref(lo), C.byref(hi), None, None))
        v = t[(t.a > 25) & (t.a <= 75), ["b"]]
        pb = D.plan_bytes(v.b)
        ot = O.OracleTable(path)
        ref = ot.aggregate_blocks(pb, 0, lo.value, hi.value)     # stands in for dfdb_scan_aggregate on this shard
        mine = _capi.Agg()
        mine.count, mine.nmissing, mine.sum_i64 = ref.count, ref.nmissing, ref.sum_i64
        mine.sum_f64, mine.sum_f64_lo = ref.sum_kahan, 0.0
        mine.min_f64, mine.max_f64, mine.min_i64, mine.max_i64 = ref.min_f64, ref.max_f64, ref.min_i64, ref.max_i64
        mine.value_class = 3 if ref.count else 0
        folded = allgather_fold(mine)
        total = allreduce_count(ref.count)
        # survivor counts of a sharded range stage travel the same way (dfdb_scan_exchange_count / _offset)
        counts = allgather_counts(ref.count)
        assert len(counts) == world and counts[rank] == ref.count and sum(counts) == total
        whole = ot.aggregate(pb, 0)
        q.put((rank, lo.value, hi.value, folded.count, total, folded.sum_f64 + folded.sum_f64_lo, folded.min_f64, folded.max_f64,
               whole.count, whole.sum_kahan, whole.min_f64, whole.max_f64))
    finally:
        dist.destroy_process_group()


def test_two_rank_shards_fold_to_the_whole_table(tmp_path, oracle):
    import torch.multiprocessing as mp
    path = str(tmp_path / "t")
    oracle.gen_table(path, "a:Int64:iuniform:1:100;b:Float64:funiform", 7 * 4096 + 123, 4096, 0xDFDB0005, 2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, path, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, *rest0), (r1, lo1, hi1, *rest1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 4, 4, 8)                 # contiguous block ranges, same for every column
    assert rest0 == rest1                                        # every rank holds the same folded result
    cnt, total, s, mn, mx, wcnt, wsum, wmn, wmx = rest0
    assert cnt == total == wcnt and mn == wmn and mx == wmx
    assert abs(s - wsum) <= 1e-12 * abs(wsum)
