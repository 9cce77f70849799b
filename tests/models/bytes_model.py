"""Lane-exact Python model of the batch rules of lz4_decode_bytes.cu (K1 for match-only byte streams).

Test infrastructure (see spec_model.py): which sequences a batch takes -- bare matches (token 0x0M, two offset bytes: 3 stream
bytes each) and at most one closing match with a single length-extension byte --, the prefix sum that places them, the
per-byte source index P[j] and its collapse by pointer jumping, and where a root byte is fetched from (the warp's ring of recent
output or global memory).  Checked on the CPU against the oracle's codec (tests/test_decoder_models.py)."""

BY_RING = 4096
BY_CLOSE_MAX = 82
BY_MAXT = 31 * 18 + BY_CLOSE_MAX
BY_NEAR = BY_RING - BY_MAXT - 64


def decode_block_bytes(src, origin):
    n_src = len(src)
    src = bytes(src) + bytes(4096)
    out = bytearray(origin)
    ring = [None] * BY_RING
    ip = op = 0
    ring_from = 0
    lim_b = origin - 12 if origin >= 12 else 0
    ip_lim = n_src - (3 * 32 + 12) if n_src >= 3 * 32 + 12 else -1
    stats = {"batches": 0, "closing": 0, "one": 0, "rounds": 0, "far": 0}
    done = False
    while not done:
        batch = False
        if ip <= ip_lim:
            x = [int.from_bytes(src[ip + 3 * l: ip + 3 * l + 4], "little") for l in range(32)]
            tok = [v & 0xff for v in x]
            off = [(v >> 8) & 0xffff for v in x]
            ext = [v >> 24 for v in x]
            M = [(t & 15) + 4 for t in tok]
            shape = [(tok[l] & 0xf0) == 0 and (tok[l] & 15) != 15 and off[l] != 0 for l in range(32)]
            n0 = shape.index(False) if not all(shape) else 32
            closing = n0 < 32 and (tok[n0] & 0xf0) == 0 and (tok[n0] & 15) == 15 and off[n0] != 0 and ext[n0] <= BY_CLOSE_MAX - 19
            Mk = [M[l] if l < n0 else (19 + ext[l] if (closing and l == n0) else 0) for l in range(32)]
            inc, acc = [], 0
            for l in range(32):
                acc += Mk[l]
                inc.append(acc)
            o = [inc[l] - Mk[l] for l in range(32)]
            ok = [l < n0 + (1 if closing else 0) and off[l] <= op + o[l] and op + inc[l] <= lim_b for l in range(32)]
            n = ok.index(False) if not all(ok) else 32
            if n > 0:
                T = inc[n - 1]
                assert T <= BY_MAXT
                adv = 3 * n + (1 if n > n0 else 0)
                P = [0] * T
                for l in range(n):
                    for i in range(Mk[l]):
                        P[o[l] + i] = o[l] + i - off[l]
                # pointer jumping: every byte ends up pointing before the batch (negative = distance back from the batch start)
                while any(p >= 0 for p in P):
                    stats["rounds"] += 1
                    snap = list(P)
                    for j in range(T):
                        if snap[j] >= 0:
                            P[j] = snap[snap[j]]
                vals = []
                for j in range(T):
                    back = -P[j]
                    a = op - back
                    assert 1 <= back <= 65535 and a >= 0
                    if back <= BY_NEAR and a >= ring_from:
                        b = ring[a & (BY_RING - 1)]
                        assert b is not None and b == out[a], (a, op)
                    else:
                        b = out[a]
                        stats["far"] += 1
                    vals.append(b)
                for j in range(T):
                    out[op + j] = vals[j]
                    ring[(op + j) & (BY_RING - 1)] = vals[j]
                ip += adv
                op += T
                batch = True
                stats["batches"] += 1
                stats["closing"] += 1 if n > n0 else 0
        if not batch:
            # anything else: one sequence (decode_one_sequence)
            stats["one"] += 1
            op_was = op
            t = src[ip]; ip += 1
            L = t >> 4
            if L == 15:
                while True:
                    e = src[ip]; ip += 1; L += e
                    if e != 255:
                        break
            out[op:op + L] = src[ip:ip + L]; ip += L; op += L
            if ip == n_src:
                done = True
            else:
                of = src[ip] | (src[ip + 1] << 8); ip += 2
                Mm = t & 15
                if Mm == 15:
                    while True:
                        e = src[ip]; ip += 1; Mm += e
                        if e != 255:
                            break
                Mm += 4
                for i in range(Mm):
                    out[op + i] = out[op - of + i]
                op += Mm
            if op - op_was <= 128:
                for a in range(op_was, op):
                    ring[a & (BY_RING - 1)] = out[a]
            else:
                ring_from = op
    assert op == origin and ip == n_src
    return bytes(out), stats
