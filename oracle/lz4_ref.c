/*
 * oracle/lz4_ref.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of the raw LZ4 *block* format used by the reference for every
 * column block (one LZ4 block per column block):
 *   reference call sites  /root/reference/src/io/BlockStreams.jl:39     LZ4_compressBound
 *                         /root/reference/src/io/BlockStreams.jl:42-48  LZ4_compress_fast(src,dst,n,cap,2)
 *                         /root/reference/src/io/BlockStreams.jl:110-112 LZ4_decompress_safe + size assert
 * The codec itself is a third-party dependency that is NOT under /root/reference:
 *   Julia package CodecLz4 (Project.toml:8, compat ">= 0.3.0", no Manifest => unpinned),
 *   which wraps upstream liblz4.  The block format is frozen and published
 *   (lz4_Block_format.md): sequences of
 *     token(1B: hi nibble = literal length, lo nibble = match length - 4)
 *     [literal length extension bytes: add each byte, stop after a byte != 255]
 *     literals
 *     offset (2B little endian, 1..65535, counted back from the current output position)
 *     [match length extension bytes]
 *   the last sequence stops after its literals.  End-of-block rules an encoder must follow:
 *   the last 5 bytes are literals and the last match starts >= 12 bytes before the end.
 *
 * Pinning: tests/test_oracle_lz4.py checks this file against the system liblz4.so.1
 * (same upstream library the reference binds) in both directions on random, RLE,
 * integer-column and string-column bodies, and reproduces the compression ratios printed in
 * /root/reference/docs/src/index.md:52-63 (2.0 / 2.55 / 2.85 / 1.93).
 */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

#define ORC_API __attribute__((visibility("default")))

#define LZ4_MINMATCH 4
#define LZ4_MFLIMIT 12
#define LZ4_LASTLITERALS 5
#define LZ4_MAX_INPUT 0x7E000000

static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

ORC_API int orc_lz4_compress_bound(int n)
{
    if (n < 0 || n > LZ4_MAX_INPUT) return 0;
    return n + n / 255 + 16;
}

/*
 * Safe decoder: never reads outside [src, src+srcSize) nor writes outside [dst, dst+dstCap).
 * Returns the number of bytes written, or a negative value on a malformed stream.  Mirrors the
 * acceptance rules of liblz4's LZ4_decompress_safe for full (non-partial) decoding:
 *   - a literal run that ends within MFLIMIT bytes of the end of the output buffer, or whose
 *     input ends within 2+1+LASTLITERALS bytes of the end of input, must be the last sequence;
 *   - a match may not end inside the last LASTLITERALS bytes of the output buffer;
 *   - offset 0 and offsets reaching before the start of the output are rejected.
 */
ORC_API int orc_lz4_decompress_safe(const uint8_t *src, uint8_t *dst, int srcSize, int dstCap)
{
    if (src == NULL || srcSize < 0 || dstCap < 0) return -1;
    if (dstCap == 0) return (srcSize == 1 && src[0] == 0) ? 0 : -1;
    if (srcSize == 0) return -1;

    const uint8_t *ip = src, *const iend = src + srcSize;
    uint8_t *op = dst, *const oend = dst + dstCap;

    for (;;) {
        if (ip >= iend) return -2;
        unsigned token = *ip++;
        size_t length = token >> 4;
        if (length == 15) {
            unsigned s;
            do {
                if (ip >= iend) return -3;
                s = *ip++;
                length += s;
            } while (s == 255);
            if (length > (size_t)LZ4_MAX_INPUT) return -4;
        }
        /* literals */
        if ((size_t)(oend - op) < length || (size_t)(iend - ip) < length) return -5;
        int near_out_end = ((size_t)(op - dst) + length + LZ4_MFLIMIT > (size_t)dstCap);
        int near_in_end = ((size_t)(ip - src) + length + (2 + 1 + LZ4_LASTLITERALS) > (size_t)srcSize);
        if (near_out_end || near_in_end) {
            if (ip + length != iend) return -6;       /* must be the last sequence */
            memcpy(op, ip, length);
            op += length;
            break;
        }
        memcpy(op, ip, length);
        op += length;
        ip += length;
        /* match */
        unsigned offset = (unsigned)ip[0] | ((unsigned)ip[1] << 8);
        ip += 2;
        if (offset == 0) return -7;
        if ((size_t)(op - dst) < offset) return -8;
        const uint8_t *match = op - offset;
        length = token & 15;
        if (length == 15) {
            unsigned s;
            do {
                if ((size_t)(ip - src) + LZ4_LASTLITERALS >= (size_t)srcSize) return -9;
                s = *ip++;
                length += s;
            } while (s == 255);
            if (length > (size_t)LZ4_MAX_INPUT) return -10;
        }
        length += LZ4_MINMATCH;
        if ((size_t)(oend - op) < length) return -11;
        if ((size_t)(op - dst) + length + LZ4_LASTLITERALS > (size_t)dstCap) return -12;   /* last 5 bytes must be literals */
        /* byte-serial copy: overlapping matches (offset < length) replicate a pattern */
        for (size_t i = 0; i < length; i++) op[i] = match[i];
        op += length;
    }
    return (int)(op - dst);
}

/* ------------------------------------------------------------------------------------------ */
/* Greedy single-probe compressor in the style of LZ4_compress_fast (hash of 5 bytes on 64-bit
 * little endian, 4096-entry position table, acceleration-controlled skip).  It emits valid LZ4
 * blocks; it is NOT required to be byte-identical with any liblz4 version (compressed bytes are
 * never compared, only decoded bytes -- SURVEY.md section 8c).                                  */

#define HASH_LOG 12
static inline uint32_t hash5(uint64_t seq)
{
    const uint64_t prime5 = 889523592379ULL;
    return (uint32_t)(((seq << 24) * prime5) >> (64 - HASH_LOG));
}

static inline unsigned match_count(const uint8_t *p, const uint8_t *m, const uint8_t *limit)
{
    const uint8_t *start = p;
    while (p + 8 <= limit) {
        uint64_t d = rd64(p) ^ rd64(m);
        if (d) return (unsigned)(p - start) + (unsigned)(__builtin_ctzll(d) >> 3);
        p += 8; m += 8;
    }
    while (p < limit && *p == *m) { p++; m++; }
    return (unsigned)(p - start);
}

ORC_API int orc_lz4_compress(const uint8_t *src, uint8_t *dst, int n, int cap, int accel)
{
    if (n < 0 || n > LZ4_MAX_INPUT) return 0;
    if (accel < 1) accel = 1;
    if (cap < orc_lz4_compress_bound(n)) return 0;   /* caller always provides the bound (BlockStreams.jl:39-40) */
    uint32_t table[1 << HASH_LOG];
    memset(table, 0, sizeof table);

    const uint8_t *ip = src, *anchor = src;
    const uint8_t *const iend = src + n;
    const uint8_t *const mflimit_p1 = iend - LZ4_MFLIMIT + 1;
    const uint8_t *const matchlimit = iend - LZ4_LASTLITERALS;
    uint8_t *op = dst;

    if (n >= LZ4_MFLIMIT + 1) {
        table[hash5(rd64(ip))] = 0;
        ip++;
        uint32_t forward_h = hash5(rd64(ip));
        for (;;) {
            const uint8_t *match;
            uint8_t *token;
            {
                const uint8_t *forward_ip = ip;
                unsigned step = 1, search_nb = (unsigned)accel << 6;
                for (;;) {
                    uint32_t h = forward_h;
                    uint32_t cur = (uint32_t)(forward_ip - src);
                    uint32_t mi = table[h];
                    ip = forward_ip;
                    forward_ip += step;
                    step = search_nb++ >> 6;
                    if (forward_ip > mflimit_p1) goto last_literals;
                    match = src + mi;
                    forward_h = hash5(rd64(forward_ip));
                    table[h] = cur;
                    if (mi + 65535u < cur) continue;
                    if (mi < cur && rd32(match) == rd32(ip)) break;
                }
            }
            while (ip > anchor && match > src && ip[-1] == match[-1]) { ip--; match--; }
            {
                unsigned lit = (unsigned)(ip - anchor);
                token = op++;
                if (lit >= 15) {
                    unsigned len = lit - 15;
                    *token = 15 << 4;
                    for (; len >= 255; len -= 255) *op++ = 255;
                    *op++ = (uint8_t)len;
                } else {
                    *token = (uint8_t)(lit << 4);
                }
                memcpy(op, anchor, lit);
                op += lit;
            }
        next_match:
            {
                unsigned off = (unsigned)(ip - match);
                *op++ = (uint8_t)off;
                *op++ = (uint8_t)(off >> 8);
                unsigned mc = match_count(ip + LZ4_MINMATCH, match + LZ4_MINMATCH, matchlimit);
                ip += mc + LZ4_MINMATCH;
                if (mc >= 15) {
                    *token += 15;
                    mc -= 15;
                    while (mc >= 255) { *op++ = 255; mc -= 255; }
                    *op++ = (uint8_t)mc;
                } else {
                    *token += (uint8_t)mc;
                }
            }
            anchor = ip;
            if (ip >= mflimit_p1) break;
            table[hash5(rd64(ip - 2))] = (uint32_t)(ip - 2 - src);
            {
                uint32_t h = hash5(rd64(ip));
                uint32_t cur = (uint32_t)(ip - src);
                uint32_t mi = table[h];
                match = src + mi;
                table[h] = cur;
                if (mi + 65535u >= cur && mi < cur && rd32(match) == rd32(ip)) {
                    token = op++;
                    *token = 0;
                    goto next_match;
                }
            }
            ip++;
            forward_h = hash5(rd64(ip));
        }
    }
last_literals:
    {
        unsigned lit = (unsigned)(iend - anchor);
        if (lit >= 15) {
            unsigned len = lit - 15;
            *op++ = 15 << 4;
            for (; len >= 255; len -= 255) *op++ = 255;
            *op++ = (uint8_t)len;
        } else {
            *op++ = (uint8_t)(lit << 4);
        }
        memcpy(op, anchor, lit);
        op += lit;
    }
    return (int)(op - dst);
}
