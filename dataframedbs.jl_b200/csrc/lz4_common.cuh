// lz4_common.cuh -- device helpers shared by the LZ4 block decoders (lz4_decode.cu, lz4_decode_v3.cu, lz4_decode_spec.cu).
//
// Everything here follows the raw LZ4 block format decoded by LZ4_decompress_safe, the call the reference
// makes in read_block (/root/reference/src/io/BlockStreams.jl:110-112): never read outside the compressed
// payload, never write outside `origin` bytes, report instead of assert.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace dfdb {
namespace lz4 {

enum { E_OK = 0, E_TRUNCATED = 1, E_OFFSET = 2, E_OVERFLOW = 3, E_SIZE = 4, E_INTERNAL = 5 };

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- warp memcpy global -> global, arbitrary alignment, vectorised on the destination -----------
__device__ __forceinline__ uint4 shift_combine(const uint4 lo, const uint4 hi, int q, int r8)
{
    uint4 o;
    switch (q) {   // q is uniform over the whole copy
    case 0: o.x = __funnelshift_r(lo.x, lo.y, r8); o.y = __funnelshift_r(lo.y, lo.z, r8); o.z = __funnelshift_r(lo.z, lo.w, r8); o.w = __funnelshift_r(lo.w, hi.x, r8); break;
    case 1: o.x = __funnelshift_r(lo.y, lo.z, r8); o.y = __funnelshift_r(lo.z, lo.w, r8); o.z = __funnelshift_r(lo.w, hi.x, r8); o.w = __funnelshift_r(hi.x, hi.y, r8); break;
    case 2: o.x = __funnelshift_r(lo.z, lo.w, r8); o.y = __funnelshift_r(lo.w, hi.x, r8); o.z = __funnelshift_r(hi.x, hi.y, r8); o.w = __funnelshift_r(hi.y, hi.z, r8); break;
    default: o.x = __funnelshift_r(lo.w, hi.x, r8); o.y = __funnelshift_r(hi.x, hi.y, r8); o.z = __funnelshift_r(hi.y, hi.z, r8); o.w = __funnelshift_r(hi.z, hi.w, r8); break;
    }
    return o;
}

// Payload slots are 16-byte aligned and 16-byte padded, so the aligned over-read of the last source chunk
// stays inside the slot.
__device__ __forceinline__ void warp_copy(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, int64_t n)
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
        uint32_t stop = __ballot_sync(FULL, b != 255u);
        if (stop == 0) { len += 255 * 32; ip += 32; if (ip >= comp_len) return false; continue; }
        int f = __ffs(stop) - 1;
        uint32_t last = __shfl_sync(FULL, b, f);
        if (ip + f >= comp_len) return false;
        len += 255 * f + last;
        ip += f + 1;
        return true;
    }
}

// One sequence, whole warp cooperating, direct global I/O.  All bytes < op are final in global memory
// on entry and on exit.  Returns E_* ; sets done when the block's last sequence was consumed.
static __device__ __noinline__ int decode_one_sequence(const uint8_t *__restrict__ src, int64_t comp_len, uint8_t *dst, int64_t origin,
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

// Baseline decoder: one sequence at a time (A/B option "lz4_simple"; also the path for blocks too large
// for the 24-bit stream / output positions of the batched decoders).
static __device__ __noinline__ int decode_simple(const uint8_t *__restrict__ src, int64_t comp_len, uint8_t *dst, int64_t origin)
{
    int64_t ip = 0, op = 0;
    bool done = false;
    while (!done) {
        int e = decode_one_sequence(src, comp_len, dst, origin, ip, op, done);
        if (e) return e;
    }
    return (op == origin && ip == comp_len) ? E_OK : E_SIZE;
}

}  // namespace lz4
}  // namespace dfdb
