// lz4_decode_bytes.cu -- K1 for match-only byte streams (String columns of few distinct values): one warp per column block,
// token positions VERIFIED instead of walked, byte-granular copies resolved by pointer jumping.
//
// Replaces read_block's LZ4_decompress_safe call (/root/reference/src/io/BlockStreams.jl:101-119) for bodies such as the
// `brand` strings of the reference's docs: after the first few hundred bytes LZ4 emits no literals at all -- 99.9 % of the
// sequences are a bare match (token 0x0M, two offset bytes), i.e. exactly 3 stream bytes -- so, as in lz4_decode_spec.cu,
// token k of a run sits at p0 + 3k and every lane can check its own.  What differs from the word decoder: the matches are
// 4..18 bytes at any offset, mostly a few dozen bytes back, so output positions are a prefix sum over the match lengths and
// most sources lie INSIDE the batch.  A batch therefore lays out one int per output byte, P[j] = the byte it copies (an
// earlier byte of the batch, or -distance for a byte before the batch), collapses the chains by pointer jumping in shared
// memory (P[j] = P[P[j]] until every entry is negative: O(log) rounds, every lane on its own bytes), and then every byte
// fetches its root from the warp's ring of recent output (shared memory) or, older than that, from global memory (L2).
// Anything that is not a bare match ends the batch and is decoded by the whole warp (decode_one_sequence).
//
// Memory-safe on any input (every access bounds-checked, per-block status); the accept / reject verdict of damaged streams is
// the lane decoder's, taken at load (api.cu).  Algorithmic bytes per block (roofline): compressed bytes read + origin bytes written.
#include <cuda_runtime.h>

#include "kernels.cuh"
#include "lz4_common.cuh"

namespace dfdb {
namespace {

using namespace lz4;

constexpr int BY_WARPS = 8;
constexpr uint32_t BY_RING = 4096;                   // bytes of recent output per warp (a power of two)
constexpr uint32_t BY_CLOSE_MAX = 82;                // longest closing match taken inside a batch (19 + one extension byte <= 63)
constexpr uint32_t BY_MAXT = 31 * 18 + BY_CLOSE_MAX;  // output bytes of a batch: bare matches of at most 18 bytes and one closing match
constexpr uint32_t BY_NEAR = BY_RING - BY_MAXT - 64; // a source at most this far back is read from the ring (this batch's writes stay off it)
constexpr uint32_t BY_WARP_SMEM = BY_RING + 4 * BY_MAXT;     // ring + P
constexpr uint32_t BY_SMEM = BY_WARPS * BY_WARP_SMEM + BY_RING;   // + alignment slack

__device__ __forceinline__ uint32_t ldg_u32(const uint8_t *p) { return __ldg(reinterpret_cast<const unsigned int *>(p)); }
// the 4 stream bytes at tp (two aligned words; payload buffers carry slack behind the last block)
__device__ __forceinline__ uint32_t load_stream4(const uint8_t *__restrict__ src, uint32_t tp)
{
    const uint8_t *a = src + (tp & ~3u);
    return __funnelshift_r(ldg_u32(a), ldg_u32(a + 4), tp * 8u);
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ int lds_s32(uint32_t a) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_s32(uint32_t a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// one block, whole warp; returns E_*.  ring_s: this warp's ring (BY_RING aligned), P behind it.
__device__ int decode_block_bytes(const uint8_t *__restrict__ src, uint32_t comp_len, uint8_t *dst, uint32_t origin, uint32_t ring_s)
{
    uint32_t lane = lane_id();
    asm volatile("" : "+r"(lane));
    const uint32_t P_s = ring_s + BY_RING;
    uint32_t ip = 0, op = 0;
    bool done = false;
    uint32_t ring_from = 0;                  // output bytes >= this one (and within the ring's reach) are in the ring
    // matches of the fast path never write the block's last 12 bytes (the end-of-block rules stay with the one-sequence path),
    // and the fast path needs 3 * 32 + 8 stream bytes ahead
    const uint32_t lim_b = origin >= 12u ? origin - 12u : 0u;
    int32_t ip_lim = comp_len >= 3u * 32u + 12u ? (int32_t)(comp_len - (3u * 32u + 12u)) : -1;
    asm volatile("" : "+r"(ip_lim));
    if (lane < 8u) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + 128u * lane));
    uint32_t tp = 3u * lane;
    uint32_t x = load_stream4(src, tp);
    while (!done) {
        bool batch = false;
        if ((int32_t)ip <= ip_lim) {
            const uint32_t tok = x & 0xffu, off = (x >> 8) & 0xffffu;
            const uint32_t M = (tok & 15u) + 4u;
            const bool shape = (tok & 0xf0u) == 0 && (tok & 15u) != 15u && off != 0;
            const uint32_t bad0 = ~__ballot_sync(FULL, shape);
            const uint32_t n0 = bad0 ? (uint32_t)__ffs(bad0) - 1u : 32u;
            // the sequence that ends the run may still be a match without literals whose length takes ONE extension byte (19 .. 82
            // bytes: two brand names in a row): 4 stream bytes, and its lane has them.  It closes the batch.
            const uint32_t ext = x >> 24;
            const bool closing = lane == n0 && (tok & 0xf0u) == 0 && (tok & 15u) == 15u && off != 0 && ext <= BY_CLOSE_MAX - 19u;
            const uint32_t ncl = __ballot_sync(FULL, closing) ? 1u : 0u;
            const uint32_t Mk = lane < n0 ? M : (closing ? 19u + ext : 0u);
            uint32_t inc = Mk;                                          // inclusive prefix sum of the match lengths
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, inc, d);
                if ((int)lane >= d) inc += t;
            }
            const uint32_t o = inc - Mk;                                // where this lane's match starts in the batch
            const bool ok = lane < n0 + ncl && off <= op + o && op + inc <= lim_b;
            const uint32_t bad = ~__ballot_sync(FULL, ok);
            const uint32_t n = bad ? (uint32_t)__ffs(bad) - 1u : 32u;   // sequences of the batch (the closing one included)
            if (n > 0) {
                const uint32_t T = __shfl_sync(FULL, inc, n - 1u);      // output bytes of the batch
                const uint32_t adv = 3u * n + (n > n0 ? 1u : 0u);       // (the closing sequence is 4 stream bytes)
                const uint32_t nip = ip + adv;
                tp += adv;
                const uint32_t nx = load_stream4(src, tp);              // the next batch's bytes travel while this one is resolved
                if ((nip ^ ip) >> 7) { if (lane == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (((nip >> 7) + 8u) << 7))); }
                // ---- lay out the batch: P[j] = index of the batch byte that byte j copies, or -(distance back from the batch start) ----
                if (lane < n) {
                    const int rel = (int)o - (int)off;                  // source of the match's first byte, relative to the batch start
                    const uint32_t a = P_s + 4u * o;
                    for (uint32_t i = 0; i < Mk; i++) sts_s32(a + 4u * i, rel + (int)i);
                }
                __syncwarp();
                // ---- collapse the chains: every byte ends up pointing before the batch.  A lane owns bytes lane, lane + 32, ... and
                // keeps a bit per byte that still points into the batch (about half of them never do), so a round touches only those ----
                uint32_t um = 0;
                for (uint32_t j = lane, k = 0; j < T; j += 32u, k++)
                    if (lds_s32(P_s + 4u * j) >= 0) um |= 1u << k;
                while (__any_sync(FULL, um != 0)) {
                    for (uint32_t m = um; m; m &= m - 1u) {
                        const uint32_t k = (uint32_t)__ffs(m) - 1u;
                        const uint32_t pa = P_s + 4u * (lane + 32u * k);
                        const int q = lds_s32(P_s + 4u * (uint32_t)lds_s32(pa));
                        sts_s32(pa, q);
                        if (q < 0) um &= ~(1u << k);
                    }
                    __syncwarp();
                }
                // ---- fetch the roots, write the bytes (global memory and the ring): bytes up to the output's next 4-byte boundary
                // one per lane, then a word per lane ----
                auto root_byte = [&](uint32_t j) -> uint32_t {
                    const uint32_t back = (uint32_t)(-lds_s32(P_s + 4u * j));     // 1 .. 65535 bytes before the batch start
                    const uint32_t a = op - back;
                    if (back <= BY_NEAR && a >= ring_from) return lds_u8(ring_s + (a & (BY_RING - 1u)));
                    return __ldcg(dst + a);
                };
                const uint32_t head = min((0u - op) & 3u, T);
                if (lane < head) {
                    const uint32_t b = root_byte(lane);
                    dst[op + lane] = (uint8_t)b;
                    sts_u8(ring_s + ((op + lane) & (BY_RING - 1u)), b);
                }
                const uint32_t words = (T - head) >> 2, tail0 = head + 4u * words;
                for (uint32_t w = lane; w < words; w += 32u) {
                    const uint32_t j = head + 4u * w;
                    const uint32_t v = root_byte(j) | (root_byte(j + 1u) << 8) | (root_byte(j + 2u) << 16) | (root_byte(j + 3u) << 24);
                    *reinterpret_cast<uint32_t *>(dst + op + j) = v;
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(ring_s + ((op + j) & (BY_RING - 1u))), "r"(v) : "memory");
                }
                if (lane < T - tail0) {
                    const uint32_t b = root_byte(tail0 + lane);
                    dst[op + tail0 + lane] = (uint8_t)b;
                    sts_u8(ring_s + ((op + tail0 + lane) & (BY_RING - 1u)), b);
                }
                __syncwarp();
                ip = nip;
                op += T;
                x = nx;
                batch = true;
            }
        }
        if (!batch) {
            // ---- anything else: one sequence, whole warp ----
            int64_t ip64 = ip, op64 = op;
            const int e = decode_one_sequence(src, comp_len, dst, origin, ip64, op64, done);
            if (e) return e;
            const uint32_t op_was = op;
            ip = (uint32_t)ip64;
            op = (uint32_t)op64;
            // what this path wrote is in global memory only: a short sequence (the usual case: a length extension, a literal or two)
            // is copied into the ring at once -- one round trip to L2 here instead of one per source byte in the batches that
            // follow --, a long one restarts the ring
            if (op - op_was <= 128u) {
                for (uint32_t a = op_was + lane; a < op; a += 32u) sts_u8(ring_s + (a & (BY_RING - 1u)), __ldcg(dst + a));
                __syncwarp();
            } else {
                ring_from = op;
            }
            tp = ip + 3u * lane;
            x = load_stream4(src, tp);
        }
    }
    return (op == origin && ip == comp_len) ? E_OK : E_SIZE;
}

__global__ void __launch_bounds__(BY_WARPS * 32, 3) lz4_decode_bytes_kernel(const __grid_constant__ DecodeArgs args, unsigned int *counter)
{
    extern __shared__ unsigned char by_dyn[];
    const uint32_t ring_s = ((smem_addr(by_dyn) + BY_RING - 1u) & ~(BY_RING - 1u)) + (threadIdx.x >> 5) * BY_WARP_SMEM;
    const uint32_t lane = lane_id();
    const long long njobs = (long long)args.ncols * args.nblocks;
    for (;;) {
        unsigned int job = 0;
        if (lane == 0) job = atomicAdd(counter, 1u);
        job = __shfl_sync(FULL, job, 0);
        if ((long long)job >= njobs) return;
        const int c = (int)(job % args.ncols);
        const int b = args.blk0 + (int)(job / args.ncols);
        const DecodeCol &col = args.col[c];
        if (col.skip && col.skip[b]) continue;
        const uint8_t *src = col.comp + col.comp_off[b];
        uint8_t *dst = col.out + col.dec_off[b];
        const uint32_t comp_len = (uint32_t)col.comp_len[b], origin = (uint32_t)col.origin[b];
        int e;
        if (origin == 0) e = (comp_len == 1 && src[0] == 0) ? E_OK : E_SIZE;
        else if (comp_len == 0) e = E_TRUNCATED;
        else if (((uintptr_t)src & 3u) || ((uintptr_t)dst & 3u)) e = decode_simple(src, comp_len, dst, origin);   // (the batch path loads and stores words)
        else e = decode_block_bytes(src, comp_len, dst, origin, ring_s);
        if (lane == 0) col.status[b] = e;
        __syncwarp();
    }
}

}  // namespace

int launch_lz4_decode_bytes(const DecodeArgs &args, unsigned int *d_counter, int sm_count, cudaStream_t stream, int cta_limit)
{
    const long long njobs = (long long)args.ncols * args.nblocks;
    if (njobs <= 0) return 0;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(lz4_decode_bytes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BY_SMEM) != cudaSuccess) return 1;
        configured = true;
    }
    cudaMemsetAsync(d_counter, 0, sizeof(unsigned int), stream);
    long long ctas = (njobs + BY_WARPS - 1) / BY_WARPS;
    const long long max_ctas = cta_limit > 0 ? cta_limit : (long long)sm_count * 3;
    if (ctas > max_ctas) ctas = max_ctas;
    lz4_decode_bytes_kernel<<<(unsigned int)ctas, BY_WARPS * 32, BY_SMEM, stream>>>(args, d_counter);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace dfdb
