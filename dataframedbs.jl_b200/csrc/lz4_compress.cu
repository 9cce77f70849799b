// lz4_compress.cu -- raw LZ4 block compression on sm_100a, one warp per column block (write path).
//
// Stands in for the LZ4_compress_fast(src, dst, n, LZ4_compressBound(n), 2) call of commit_block_write!
// (/root/reference/src/io/BlockStreams.jl:36-60).  The output is a valid raw LZ4 block that LZ4_decompress_safe -- the
// reference's reader (BlockStreams.jl:110-112), system liblz4 and every decoder of this repository -- decodes to the same
// body; the compressed BYTES are not liblz4's (the reference never compares them), the ratio is.
//
// Algorithm: the greedy single-probe parse of LZ4's fast mode, with the probe done 32 positions at a time.  One warp
// owns one body and a private 4096-entry hash table of positions in shared memory.  Per round every lane hashes the 5
// bytes at its own position (ip + lane), looks up the candidate and verifies it (4 equal bytes, distance 1..65535); the
// lowest matching position wins, like the sequential scan that would have met it first, and the positions up to it are
// recorded in the table.  The winner's match is extended backwards over pending literals and forwards 64 bytes per step (8 bytes per
// lane), the sequence (literals + offset + lengths) is written by the whole warp, and the scan resumes behind the match.
// End-of-block rules of the format: no match starts within the last 12 bytes or ends within the last 5, the block ends
// with a literal-only sequence.
//
// One body is at most ~1 MB (a column block), so one warp per body keeps thousands of bodies in flight; the kernel is
// latency / LSU bound, not HBM bound: algorithmic bytes are n read + compressed written per body.
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace dfdb {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int CMP_WARPS = 8;
constexpr int CMP_THREADS = CMP_WARPS * 32;
constexpr int HASH_LOG = 12;
constexpr int HASH_SIZE = 1 << HASH_LOG;
constexpr uint32_t MFLIMIT = 12, LASTLITERALS = 5, MINMATCH = 4, MAX_DISTANCE = 65535;

__device__ __forceinline__ uint32_t load4(const uint8_t *p)
{
    const uintptr_t a = (uintptr_t)p;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    const uint32_t lo = w[0];
    if (sh == 0) return lo;
    return __funnelshift_r(lo, w[1], sh);     // (bodies are padded: the word behind the last byte is readable)
}
__device__ __forceinline__ uint64_t load8(const uint8_t *p)
{
    return (uint64_t)load4(p) | ((uint64_t)load4(p + 4) << 32);
}
// LZ4's own choice on 64-bit targets: the table is keyed on FIVE bytes although a match needs only four.  A four-byte key finds
// many bare four-byte matches (3 bytes of output for 4 of input) that keep the parse from reaching the longer match one or
// two bytes later: brand strings came out 45 % larger than liblz4's with it, and within 0.1 % with this one.
__device__ __forceinline__ uint32_t hash5(uint64_t seq) { return (uint32_t)(((seq << 24) * 889523592379ull) >> (64 - HASH_LOG)); }

// length extension bytes of the format: 255, 255, ..., rest
__device__ __forceinline__ uint32_t write_len_ext(uint8_t *dst, uint32_t op, uint32_t len, uint32_t lane)
{
    // len >= 15 already went into the token nibble
    const uint32_t r = len - 15, nfull = r / 255;
    for (uint32_t i = lane; i < nfull; i += 32) dst[op + i] = 255;
    if (lane == 0) dst[op + nfull] = (uint8_t)(r - nfull * 255);
    return op + nfull + 1;
}

__global__ void __launch_bounds__(CMP_THREADS) lz4_compress_kernel(const CompressArgs A, unsigned int *counter)
{
    extern __shared__ uint32_t tables[];                       // CMP_WARPS x HASH_SIZE positions
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *table = tables + warp * HASH_SIZE;
    for (;;) {
        unsigned int job = 0;
        if (lane == 0) job = atomicAdd(counter, 1u);
        job = __shfl_sync(FULL, job, 0);
        if (job >= (unsigned int)A.nblocks) return;
        const uint8_t *src = A.src + A.src_off[job];
        uint8_t *dst = A.dst + A.dst_off[job];
        const uint32_t n = (uint32_t)A.src_len[job];
        for (int i = lane; i < HASH_SIZE; i += 32) table[i] = 0;
        __syncwarp();
        uint32_t ip = 0, anchor = 0, op = 0;
        const uint32_t mflimit = n >= MFLIMIT + 1 ? n - MFLIMIT : 0;        // a match may start at positions <= mflimit (when n > 12)
        const uint32_t matchlimit = n >= LASTLITERALS ? n - LASTLITERALS : 0;
        if (n > MFLIMIT) {
            // position 0 can never match (nothing before it); it is only recorded
            while (ip <= mflimit) {
                const uint32_t p = ip + lane;
                const bool inside = p <= mflimit;
                uint32_t seq = 0, h = 0, cand = 0;
                if (inside) {
                    const uint64_t seq8 = load8(src + p);
                    seq = (uint32_t)seq8;
                    h = hash5(seq8);
                    cand = table[h];
                }
                __syncwarp();
                const bool hit = inside && cand < p && p - cand <= MAX_DISTANCE && load4(src + cand) == seq;
                const unsigned hits = __ballot_sync(FULL, hit);
                // Record the positions the sequential scan would have examined: everything up to the match (all 32 without one).
                // Recording the positions behind it too would make the next round find ITSELF instead of the older occurrence.
                // Lanes with the same hash: the highest position wins, as it would sequentially (and the output is reproducible).
                {
                    const int upto = hits ? __ffs(hits) - 1 : 31;
                    const bool rec = inside && (int)lane <= upto;
                    const unsigned peers = __match_any_sync(FULL, rec ? h : 0xffffffffu - lane);
                    if (rec && (peers >> lane) == 1u) table[h] = p;
                }
                __syncwarp();
                if (hits == 0) { ip += 32; continue; }
                const int f = __ffs(hits) - 1;
                uint32_t m = ip + (uint32_t)f, ref = __shfl_sync(FULL, cand, f);
                // backwards over the pending literals (LZ4 does the same: the match may start earlier than where it was found)
                while (m > anchor && ref > 0 && src[m - 1] == src[ref - 1]) { m--; ref--; }
                // forwards: 8 bytes per lane, 256 per step
                uint32_t mlen = MINMATCH;
                for (;;) {
                    const uint32_t q = m + mlen + lane * 8;
                    uint32_t same = 0;                         // equal bytes of this lane's 8 (0..8), 0 when outside
                    if (q < matchlimit) {
                        const uint32_t room = matchlimit - q;
                        const uint64_t x = load8(src + q) ^ load8(src + ref + mlen + lane * 8);
                        same = x ? (uint32_t)(__ffsll((long long)x) - 1) >> 3 : 8;
                        if (same > room) same = room;
                    }
                    const unsigned partial = __ballot_sync(FULL, same < 8);
                    if (partial == 0) { mlen += 256; continue; }
                    const int g = __ffs(partial) - 1;
                    mlen += 8 * (uint32_t)g + __shfl_sync(FULL, same, g);
                    break;
                }
                // ---- emit the sequence: token | literal length ext | literals | offset | match length ext ----
                const uint32_t lit = m - anchor, ml = mlen - MINMATCH;
                if (lane == 0) dst[op] = (uint8_t)(((lit < 15 ? lit : 15) << 4) | (ml < 15 ? ml : 15));
                op += 1;
                if (lit >= 15) op = write_len_ext(dst, op, lit, lane);
                for (uint32_t i = lane; i < lit; i += 32) dst[op + i] = src[anchor + i];
                op += lit;
                if (lane == 0) { dst[op] = (uint8_t)((m - ref) & 0xff); dst[op + 1] = (uint8_t)((m - ref) >> 8); }
                op += 2;
                if (ml >= 15) op = write_len_ext(dst, op, ml, lane);
                ip = m + mlen;
                anchor = ip;
                // (like LZ4_putPosition(ip - 2): positions inside the match are not indexed, its tail is)
                if (lane == 0 && ip >= 2 && ip - 2 <= mflimit) table[hash5(load8(src + ip - 2))] = ip - 2;
                __syncwarp();
            }
        }
        // ---- last sequence: literals only ----
        const uint32_t lit = n - anchor;
        if (lane == 0) dst[op] = (uint8_t)((lit < 15 ? lit : 15) << 4);
        op += 1;
        if (lit >= 15) op = write_len_ext(dst, op, lit, lane);
        for (uint32_t i = lane; i < lit; i += 32) dst[op + i] = src[anchor + i];
        op += lit;
        if (lane == 0) A.dst_len[job] = (int64_t)op;
        __syncwarp();
    }
}

}  // namespace

int launch_lz4_compress(const CompressArgs &a, unsigned int *d_counter, int sm_count, cudaStream_t stream)
{
    if (a.nblocks <= 0) return 0;
    constexpr int smem = CMP_WARPS * HASH_SIZE * 4;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(lz4_compress_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 1;
        configured = true;
    }
    cudaMemsetAsync(d_counter, 0, sizeof(unsigned int), stream);
    int ctas = (a.nblocks + CMP_WARPS - 1) / CMP_WARPS;
    if (ctas > sm_count) ctas = sm_count;                     // 128 KB of hash tables per CTA: one CTA per SM
    lz4_compress_kernel<<<ctas, CMP_THREADS, smem, stream>>>(a, d_counter);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace dfdb
