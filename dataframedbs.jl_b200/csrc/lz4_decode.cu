// lz4_decode.cu -- K1: raw LZ4 block decode on sm_100a, one warp per compressed column block.
//
// Replaces read_block's LZ4_decompress_safe call (/root/reference/src/io/BlockStreams.jl:101-119,
// liblz4 through CodecLz4) for whole batches of independent column blocks.  Like the _safe decoder it
// never reads outside the compressed payload nor writes outside `origin` bytes, and it reports a
// per-block status instead of the reference's `@assert size == sizes.origin "decompression error"`.
//
// Algorithm (decode_batched): the serial part of LZ4 is only the *position* of each token.  A warp
//   A. walks 32 tokens of the smem-staged compressed stream with a 3-instruction dependent chain
//      (LDS token, shift, add) and records the 32 token positions;
//   B. then handles the 32 sequences lane-parallel: every lane reads its own token/offset, a warp
//      prefix sum of (literals + match) lengths gives every lane its output position, literals are
//      copied from the staged stream into an smem output staging area, matches are copied in
//      dependency waves (a lane whose source bytes are already final runs immediately; sources inside
//      the current batch wait for the frontier of completed lanes);
//   C. the staged output is flushed to HBM with coalesced 16-byte stores.
// Sequences with length extensions (literal or match nibble == 15) and the last sequence of a block
// take the warp-cooperative path decode_one_sequence, whose literal copy is a 16-byte vectorised,
// re-aligning warp memcpy -- that path is what incompressible Float64 blocks (one giant literal run)
// exercise, so it has to run at copy bandwidth.
//
// Algorithmic bytes per block (roofline): compressed bytes read + origin bytes written.
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace dfdb {

namespace {

constexpr int WARPS_PER_CTA = 8;
constexpr int WIN = 2048;              // bytes of compressed stream staged per warp
constexpr int WIN_PAD = 64;
constexpr int BATCH_IN_MAX = 32 * 17;  // 32 sequences * (token + 14 literals + offset)
constexpr int STG = 1280;              // output staging: 15 carried bytes + 32 * (14 + 18) bytes, rounded up

enum { E_OK = 0, E_TRUNCATED = 1, E_OFFSET = 2, E_OVERFLOW = 3, E_SIZE = 4 };

struct WarpSmem {
    __align__(16) uint8_t win[WIN + WIN_PAD];
    __align__(16) uint8_t stg[STG];
    uint16_t pos[32];
};

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// ---- warp memcpy global -> global, arbitrary alignment, vectorised on the destination -----------
__device__ __forceinline__ uint4 shift_combine(const uint4 lo, const uint4 hi, int q, int r8)
{
    uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    uint4 o;
    switch (q) {   // q is uniform over the whole copy
    case 0: o.x = __funnelshift_r(w[0], w[1], r8); o.y = __funnelshift_r(w[1], w[2], r8); o.z = __funnelshift_r(w[2], w[3], r8); o.w = __funnelshift_r(w[3], w[4], r8); break;
    case 1: o.x = __funnelshift_r(w[1], w[2], r8); o.y = __funnelshift_r(w[2], w[3], r8); o.z = __funnelshift_r(w[3], w[4], r8); o.w = __funnelshift_r(w[4], w[5], r8); break;
    case 2: o.x = __funnelshift_r(w[2], w[3], r8); o.y = __funnelshift_r(w[3], w[4], r8); o.z = __funnelshift_r(w[4], w[5], r8); o.w = __funnelshift_r(w[5], w[6], r8); break;
    default: o.x = __funnelshift_r(w[3], w[4], r8); o.y = __funnelshift_r(w[4], w[5], r8); o.z = __funnelshift_r(w[5], w[6], r8); o.w = __funnelshift_r(w[6], w[7], r8); break;
    }
    return o;
}

// `src_limit` = one past the last readable byte of the source buffer rounded up to 16 (payload slots are
// 16-byte padded), so the aligned over-read of the last chunk stays inside the slot.
__device__ void warp_copy(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, int64_t n)
{
    const uint32_t lane = lane_id();
    int64_t head = (int64_t)((16 - ((uintptr_t)dst & 15)) & 15);
    if (head > n) head = n;
    if ((int64_t)lane < head) dst[lane] = src[lane];
    dst += head; src += head; n -= head;
    const int64_t nchunks = n >> 4;
    if (nchunks > 0) {
        uint4 *d16 = reinterpret_cast<uint4 *>(dst);
        const int s = (int)((uintptr_t)src & 15);
        if (s == 0) {
            const uint4 *s16 = reinterpret_cast<const uint4 *>(src);
            int64_t c = lane;
            for (; c + 96 < nchunks; c += 128) {
                uint4 v0 = __ldg(s16 + c), v1 = __ldg(s16 + c + 32), v2 = __ldg(s16 + c + 64), v3 = __ldg(s16 + c + 96);
                d16[c] = v0; d16[c + 32] = v1; d16[c + 64] = v2; d16[c + 96] = v3;
            }
            for (; c < nchunks; c += 32) d16[c] = __ldg(s16 + c);
        } else {
            const uint4 *s16 = reinterpret_cast<const uint4 *>(src - s);   // aligned base; chunk c needs s16[c], s16[c+1]
            const int q = s >> 2, r8 = (s & 3) * 8;
            int64_t c = lane;
            for (; c + 96 < nchunks; c += 128) {
                uint4 a0 = __ldg(s16 + c), b0 = __ldg(s16 + c + 1);
                uint4 a1 = __ldg(s16 + c + 32), b1 = __ldg(s16 + c + 33);
                uint4 a2 = __ldg(s16 + c + 64), b2 = __ldg(s16 + c + 65);
                uint4 a3 = __ldg(s16 + c + 96), b3 = __ldg(s16 + c + 97);
                d16[c] = shift_combine(a0, b0, q, r8);
                d16[c + 32] = shift_combine(a1, b1, q, r8);
                d16[c + 64] = shift_combine(a2, b2, q, r8);
                d16[c + 96] = shift_combine(a3, b3, q, r8);
            }
            for (; c < nchunks; c += 32) d16[c] = shift_combine(__ldg(s16 + c), __ldg(s16 + c + 1), q, r8);
        }
    }
    const int64_t done = nchunks << 4;
    const int64_t tail = n - done;
    if ((int64_t)lane < tail) dst[done + lane] = src[done + lane];
}

// length extension bytes (LZ4: add bytes while they are 255); cooperative over the warp.
// returns false when the stream ends inside the extension.
__device__ __forceinline__ bool read_length_ext(const uint8_t *__restrict__ src, int64_t &ip, int64_t comp_len, int64_t &len)
{
    const uint32_t lane = lane_id();
    for (;;) {
        int64_t p = ip + lane;
        uint32_t b = p < comp_len ? src[p] : 0u;     // past the end reads as a terminator and is caught below
        uint32_t stop = __ballot_sync(0xffffffffu, b != 255u);
        if (stop == 0) { len += 255 * 32; ip += 32; if (ip >= comp_len) return false; continue; }
        int f = __ffs(stop) - 1;
        uint32_t last = __shfl_sync(0xffffffffu, b, f);
        if (ip + f >= comp_len) return false;
        len += 255 * f + last;
        ip += f + 1;
        return true;
    }
}

// One sequence, whole warp cooperating, direct global I/O.  All bytes < op are final in global memory
// on entry and on exit.  Returns E_* ; sets done when the block's last sequence was consumed.
__device__ int decode_one_sequence(const uint8_t *__restrict__ src, int64_t comp_len, uint8_t *dst, int64_t origin,
                                   int64_t &ip, int64_t &op, bool &done)
{
    const uint32_t lane = lane_id();
    if (ip >= comp_len) return E_TRUNCATED;
    const uint32_t token = src[ip];
    ip += 1;
    int64_t L = token >> 4;
    if (L == 15 && !read_length_ext(src, ip, comp_len, L)) return E_TRUNCATED;
    if (ip + L > comp_len) return E_TRUNCATED;
    if (op + L > origin) return E_OVERFLOW;
    if (L > 0) warp_copy(dst + op, src + ip, L);
    ip += L;
    op += L;
    if (ip == comp_len) { done = true; return E_OK; }   // last sequence: literals only
    if (ip + 2 > comp_len) return E_TRUNCATED;
    const uint32_t off = (uint32_t)src[ip] | ((uint32_t)src[ip + 1] << 8);
    ip += 2;
    int64_t M = token & 15;
    if (M == 15 && !read_length_ext(src, ip, comp_len, M)) return E_TRUNCATED;
    M += 4;
    if (off == 0 || (int64_t)off > op) return E_OFFSET;
    if (op + M > origin) return E_OVERFLOW;
    __syncwarp();                                       // literal bytes visible to the whole warp
    // every source byte is < op, i.e. already final: the copy is fully parallel even when it overlaps
    uint8_t *m_dst = dst + op;
    const uint8_t *m_src = dst + op - off;
    if ((int64_t)off >= M) {
        for (int64_t i = lane; i < M; i += 32) m_dst[i] = __ldcg(m_src + i);
    } else {
        for (int64_t i = lane; i < M; i += 32) m_dst[i] = __ldcg(m_src + ((uint32_t)i % off));
    }
    op += M;
    __syncwarp();
    return E_OK;
}

// ---- baseline decoder: one sequence at a time (kept for A/B checks, option "lz4_simple") ---------
__device__ int decode_simple(const uint8_t *__restrict__ src, int64_t comp_len, uint8_t *dst, int64_t origin)
{
    int64_t ip = 0, op = 0;
    bool done = false;
    while (!done) {
        int e = decode_one_sequence(src, comp_len, dst, origin, ip, op, done);
        if (e) return e;
    }
    return (op == origin && ip == comp_len) ? E_OK : E_SIZE;
}

__device__ __forceinline__ uint8_t stg_or_global(const WarpSmem &sm, const uint8_t *dst, int64_t stg_base, int64_t x)
{
    return x >= stg_base ? sm.stg[x - stg_base] : __ldcg(dst + x);
}

// ---- batched decoder ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int K>
__device__ __forceinline__ void chain_steps(uint32_t &b, uint32_t pos_sa)
{
    uint32_t t;
    asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(t) : "r"(b), "n"(3 * K));
    asm volatile("st.shared.u16 [%0+%1], %2;" ::"r"(pos_sa), "n"(2 * K), "h"((uint16_t)b));
    b += t >> 4;
    if constexpr (K + 1 < 32) chain_steps<K + 1>(b, pos_sa);
}

// Block sizes are < 2^31 (LZ4_MAX_INPUT_SIZE), so every stream / output offset is kept in 32 bits.
__device__ int decode_batched(WarpSmem &sm, const uint8_t *__restrict__ src, int64_t comp_len64, uint8_t *dst, int64_t origin64)
{
    const uint32_t lane = lane_id();
    const int32_t comp_len = (int32_t)comp_len64, origin = (int32_t)origin64;
    const int32_t comp_pad = (comp_len + 15) & ~15;
    const uint32_t win_sa = smem_addr(sm.win), pos_sa = smem_addr(sm.pos);
    int32_t ip = 0, op = 0;
    int32_t win_base = -2 * WIN;            // forces the first refill
    int32_t stg_base = 0;                   // 16-aligned; staging holds out bytes [stg_base, op)
    bool done = false;

    while (!done) {
        // ---- input window -------------------------------------------------------------------
        if (ip < win_base || ip + BATCH_IN_MAX > win_base + WIN) {
            win_base = ip & ~15;
            const uint4 *s16 = reinterpret_cast<const uint4 *>(src + win_base);
            uint4 *w16 = reinterpret_cast<uint4 *>(sm.win);
#pragma unroll
            for (int c = 0; c < (WIN + WIN_PAD) / 16 / 32 + 1; c++) {
                int ch = c * 32 + lane;
                if (ch < (WIN + WIN_PAD) / 16) {
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if (win_base + ch * 16 < comp_pad) v = __ldg(s16 + ch);
                    w16[ch] = v;
                }
            }
            __syncwarp();
        }
        // ---- A: token chain, 3 instructions per token: LDS.U8 [b + 3k], STS.U16 [pos + 2k], b += token >> 4
        // (b_k = window address of token k minus 3k, so the "+3" of every sequence rides in the immediates)
        {
            uint32_t b = win_sa + (uint32_t)(ip - win_base);
            chain_steps<0>(b, pos_sa);
        }
        __syncwarp();
        // ---- B: lane-parallel sequences -----------------------------------------------------
        const uint32_t mypos = (uint32_t)sm.pos[lane] + 3u * lane - (win_sa & 0xffffu);
        const uint32_t tok = sm.win[mypos];
        const uint32_t L = tok >> 4, Mn = tok & 15, M = Mn + 4;
        const int32_t seq_end_in = win_base + (int32_t)(mypos + 3 + L);
        const bool simple = (L < 15) && (Mn < 15) && (seq_end_in < comp_len);
        const uint32_t off = simple ? ((uint32_t)sm.win[mypos + 1 + L] | ((uint32_t)sm.win[mypos + 2 + L] << 8)) : 1u;
        // "regular" = one aligned 8-byte output word whose match source is an aligned word
        const bool regular = simple && (L + M == 8) && ((off & 7u) == 0) && (off != 0);
        const uint32_t bs = __ballot_sync(0xffffffffu, simple);
        const int nvalid = (bs == 0xffffffffu) ? 32 : (__ffs(~bs) - 1);
        if (nvalid == 0) {
            // global memory is complete below op: run one cooperative sequence, then re-seed the staging
            int64_t ip64 = ip, op64 = op;
            int e = decode_one_sequence(src, comp_len64, dst, origin64, ip64, op64, done);
            if (e) return e;
            ip = (int32_t)ip64;
            op = (int32_t)op64;
            stg_base = op & ~15;
            const int tail = op - stg_base;
            if ((int)lane < tail) sm.stg[lane] = __ldcg(dst + stg_base + lane);
            __syncwarp();
            continue;
        }
        // ---- one pass over the nvalid sequences of the batch ----------------------------------------------
        // Output positions from one warp prefix sum; "regular" lanes (one aligned 8-byte word whose source is
        // an aligned word: L + M == 8, off % 8 == 0, o % 8 == 0) are resolved by word forwarding, the few
        // remaining lanes by byte-granular dependency waves; the batch leaves through the smem staging area.
        const bool active = (int)lane < nvalid;
        const uint32_t len = active ? L + M : 0u;
        uint32_t incl = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if ((int)lane >= d) incl += v;
        }
        const int32_t total = (int32_t)__shfl_sync(0xffffffffu, incl, 31);
        const int32_t o = op + (int32_t)(incl - len);   // my literals start here
        const int32_t m_dst = o + (int32_t)L;           // my match starts here
        const int32_t m_src = m_dst - (int32_t)off;
        bool err = active && (off == 0 || m_src < 0);
        if (op + total > origin) return E_OVERFLOW;      // uniform
        const int32_t sw = o - (int32_t)off;             // source word of a regular lane
        bool reg = active && regular && ((o & 7) == 0) && !err;
        // in-batch source: the producer must be the regular lane that owns exactly that word
        int dep = -1;
        {
            const int guess = (int)lane - (int)(off >> 3);
            const int j = guess < 0 ? 0 : guess;
            const int32_t oj = __shfl_sync(0xffffffffu, o, j);
            const uint32_t regmask = __ballot_sync(0xffffffffu, reg);
            if (reg && sw + 8 > op) {                    // source word not entirely below the batch start
                if (guess >= 0 && oj == sw && ((regmask >> j) & 1u)) dep = j;
                else reg = false;                        // source lies inside a non-regular sequence: byte path
            }
        }
        // a regular lane whose producer was demoted must be demoted too (rare): iterate to a fixed point
        for (;;) {
            const uint32_t regmask = __ballot_sync(0xffffffffu, reg);
            const bool drop = reg && dep >= 0 && !((regmask >> dep) & 1u);
            if (!__any_sync(0xffffffffu, drop)) break;
            if (drop) { reg = false; dep = -1; }
        }
        const uint32_t regmask = __ballot_sync(0xffffffffu, reg);
        const uint32_t genmask = __ballot_sync(0xffffffffu, active && !reg && !err);

        // ---- regular lanes: word forwarding ------------------------------------------------------------
        if (regmask) {
            unsigned long long lit = 0;
            if (__any_sync(0xffffffffu, reg && L > 0)) {
                if (reg) {
#pragma unroll
                    for (uint32_t i = 0; i < 4; i++)
                        if (i < L) lit |= (unsigned long long)sm.win[mypos + 1 + i] << (8 * i);
                }
            }
            uint32_t Lc = reg ? L : 0;                        // low Lc bytes of my word are already known
            unsigned long long val = lit;
            bool fin = !reg || dep < 0;
            if (reg && dep < 0) {
                const unsigned long long W = __ldcg(reinterpret_cast<const unsigned long long *>(dst + sw));
                val = lit | (W & (~0ull << (8 * L)));
            }
            while (__any_sync(0xffffffffu, !fin)) {
                const int j = dep < 0 ? (int)lane : dep;
                const unsigned long long vj = __shfl_sync(0xffffffffu, val, j);
                const unsigned long long lj = __shfl_sync(0xffffffffu, lit, j);
                const uint32_t Lj = __shfl_sync(0xffffffffu, Lc, j);
                const int dj = __shfl_sync(0xffffffffu, dep, j);
                const uint32_t finmask = __ballot_sync(0xffffffffu, fin);
                if (!fin) {
                    const unsigned long long keep = ~0ull << (8 * Lc);
                    if ((finmask >> j) & 1u) { val = lit | (vj & keep); fin = true; dep = -1; }
                    else { lit |= lj & keep; Lc = Lc > Lj ? Lc : Lj; dep = dj; }
                }
            }
            if (reg) *reinterpret_cast<unsigned long long *>(sm.stg + (o - stg_base)) = val;
        }
        // ---- remaining lanes: literals + matches in dependency waves ------------------------------------------
        if (genmask) {
            const bool gen = (genmask >> lane) & 1u;
            if (gen) {
                for (uint32_t i = 0; i < L; i++) sm.stg[o - stg_base + i] = sm.win[mypos + 1 + i];
            }
            __syncwarp();
            // Frontier F: every output byte < F is final (regular lanes are already done).
            const int32_t src_end = (m_src + (int32_t)M < m_dst) ? m_src + (int32_t)M : m_dst;   // bytes needed from other producers end here
            uint32_t pending = genmask;
            int P = __ffs(pending) - 1;                                                     // first lane whose match is not done
            int32_t F = __shfl_sync(0xffffffffu, m_dst, P);
            while (pending) {
                const bool mine = (pending >> lane) & 1u;
                const bool ready = mine && (src_end <= F || (int)lane == P);
                if (ready) {
                    const int32_t sd = m_dst - stg_base;
                    uint32_t i = 0;
                    if (off >= 8 && ((m_dst | m_src) & 7) == 0) {
                        for (; i + 8 <= M; i += 8) {
                            unsigned long long v;
                            const int32_t x = m_src + (int32_t)i;
                            if (x >= stg_base) v = *reinterpret_cast<const unsigned long long *>(sm.stg + (x - stg_base));
                            else v = __ldcg(reinterpret_cast<const unsigned long long *>(dst + x));
                            *reinterpret_cast<unsigned long long *>(sm.stg + sd + i) = v;
                        }
                    }
                    // byte-serial per lane: correct for self-overlapping matches (off < M) as well
                    for (; i < M; i++) {
                        const int32_t x = m_src + (int32_t)i;
                        sm.stg[sd + i] = x >= stg_base ? sm.stg[x - stg_base] : __ldcg(dst + x);
                    }
                }
                __syncwarp();
                pending &= ~__ballot_sync(0xffffffffu, ready);
                if (pending) {
                    P = __ffs(pending) - 1;
                    F = __shfl_sync(0xffffffffu, m_dst, P);
                }
            }
        }
        if (__any_sync(0xffffffffu, err)) return E_OFFSET;
        __syncwarp();
        // ---- C: flush every 16-byte chunk the batch touched (the partial last chunk too, so global memory is
        //      always complete below op); the tail bytes stay in the staging area for the next batch ----------
        const int32_t new_op = op + total;
        const int nfull = (new_op >> 4) - (stg_base >> 4);
        const int nchunks = nfull + ((new_op & 15) ? 1 : 0);
        {
            uint4 *d16 = reinterpret_cast<uint4 *>(dst + stg_base);
            const uint4 *g16 = reinterpret_cast<const uint4 *>(sm.stg);
            for (int c = lane; c < nchunks; c += 32) d16[c] = g16[c];
        }
        const int tail = new_op & 15;
        uint8_t tb = 0;
        if (nfull > 0 && (int)lane < tail) tb = sm.stg[nfull * 16 + lane];
        __syncwarp();
        if (nfull > 0 && (int)lane < tail) sm.stg[lane] = tb;
        stg_base += nfull * 16;
        op = new_op;
        ip = __shfl_sync(0xffffffffu, seq_end_in, nvalid - 1);
        __syncwarp();
    }
    // block finished inside decode_one_sequence (which leaves everything < op in global memory)
    return (op == origin && ip == comp_len) ? E_OK : E_SIZE;
}

}  // namespace

__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 4) lz4_decode_kernel(DecodeArgs args, unsigned int *counter, int simple_mode)
{
    __shared__ WarpSmem smem[WARPS_PER_CTA];
    const int warp = threadIdx.x >> 5;
    const uint32_t lane = lane_id();
    const unsigned int njobs = (unsigned int)args.ncols * (unsigned int)args.nblocks;
    for (;;) {
        unsigned int job = 0;
        if (lane == 0) job = atomicAdd(counter, 1u);
        job = __shfl_sync(0xffffffffu, job, 0);
        if (job >= njobs) break;
        const int c = (int)(job % (unsigned int)args.ncols);
        const int b = args.blk0 + (int)(job / (unsigned int)args.ncols);
        const DecodeCol &col = args.col[c];
        const uint8_t *src = col.comp + col.comp_off[b];
        uint8_t *dst = col.out + col.dec_off[b];
        const int64_t comp_len = col.comp_len[b], origin = col.origin[b];
        int e = simple_mode ? decode_simple(src, comp_len, dst, origin) : decode_batched(smem[warp], src, comp_len, dst, origin);
        if (lane == 0) col.status[b] = e;
        __syncwarp();
    }
}

int launch_lz4_decode(const DecodeArgs &args, unsigned int *d_counter, int sm_count, int simple_mode, cudaStream_t stream)
{
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(lz4_decode_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        configured = true;
    }
    cudaMemsetAsync(d_counter, 0, sizeof(unsigned int), stream);
    const long long njobs = (long long)args.ncols * args.nblocks;
    long long ctas = (njobs + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    const long long max_ctas = (long long)sm_count * 4;     // 32 resident warps per SM, persistent over the job queue
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    lz4_decode_kernel<<<(unsigned int)ctas, WARPS_PER_CTA * 32, 0, stream>>>(args, d_counter, simple_mode);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace dfdb
