// lz4_lane_core.cuh -- the per-lane state machine of the lane-per-block LZ4 decoder (lz4_decode_lane.cu).
//
// One LANE decodes one compressed column block from start to end (read_block's LZ4_decompress_safe call,
// /root/reference/src/io/BlockStreams.jl:101-119), as a software pipeline of two independent halves:
//
//   parse  (Parser::step)   walks the token chain through a 16-byte register window over the lane's private
//                           shared-memory window of the compressed stream and describes ONE PIECE per step: at most
//                           8 output bytes that are either literal bytes (taken from the stream right away) or a copy
//                           from at most 65535 bytes back in the output.  A copy whose source will still be in the
//                           lane's shared-memory history ring when the piece is emitted is described by its position
//                           (NEAR); an older, 8-byte aligned source is already final in global memory and is fetched at
//                           parse time straight into the piece queue (an asynchronous copy that has the whole pipeline
//                           depth to land); an older unaligned source is loaded at emit time (FARSLOW, rare).
//                           The pieces go through a DEPTH-deep queue in shared memory (Mem::put_data / put_far / get_data),
//                           so the step is ONE copy of the code in a rolled loop: the kernel runs one warp per SM
//                           sub-partition and an unrolled step loop does not fit the instruction cache.
//   emit   (Emitter::step)  DEPTH steps later appends the piece to the output: merges it into the partial output word,
//                           keeps every byte below `op` in the history ring, and leaves complete 128-byte units for the
//                           warp-cooperative flush to global memory.
//
// The acceptance rules are liblz4's (LZ4_decompress_safe, full-block mode), restated from the published block
// format the same way oracle/lz4_ref.c does: a literal run that ends within 12 bytes of the end of the output or
// whose input ends within 8 bytes of the end of the input must be the last sequence and end exactly at the end
// of the input; a match may not end within the last 5 bytes of the output; offset 0 and offsets reaching before
// the start of the output are rejected; the decoded size must equal `origin` (BlockStreams.jl:112).
//
// Everything here is plain C++ over a `Mem` policy (window / ring / global accessors), so that the same code
// runs inside the CUDA kernel and inside the host harness tests/lane_sim.cpp, which checks it against the oracle
// without a GPU.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define LZL_HD __host__ __device__ __forceinline__
#else
#define LZL_HD inline
#endif

namespace dfdb {
namespace lane {

constexpr int RING_WORDS = 160;          // history ring per lane (8-byte words); a multiple of UNIT_WORDS
constexpr int RING_STRIDE = RING_WORDS + 1;   // odd word stride between lanes: lanes at the same ring slot hit different banks
constexpr int WIN_BYTES = 256;           // compressed-stream window per lane (circular by stream position)
constexpr int DEPTH = 8;                 // steps between parse and emit (the piece queue holds SUBS * DEPTH pieces); a power of two
constexpr int ROUND = 16;                // steps between two cooperative flush / refill rounds (a multiple of DEPTH)
constexpr int UNIT_BYTES = 128;          // flush unit
constexpr int UNIT_WORDS = UNIT_BYTES / 8;
constexpr int NEAR_WORDS = RING_WORDS - 2;    // a source at most this many words behind the piece is read from the ring
constexpr uint32_t STEP_LOOKAHEAD = 40;  // stream bytes past `ip` a step may touch (token + one length chunk + one literal piece + register window)
constexpr uint32_t MAX_POS = 1u << 25;   // piece descriptors carry 25-bit output positions
constexpr int SUBS = 2;                  // pieces per lane and step (queue sub-slots A, B)
constexpr int HOT_PERIOD = 8;            // word-regular columns: every HOT_PERIOD-th step is a full step (fast paths + the general state machine);
                                         // the others only run the two-plain-tokens hot path, and a lane that meets anything else waits for the
                                         // next full step.  It divides DEPTH, so the pieces of full steps are emitted in full steps.

static_assert(RING_WORDS % UNIT_WORDS == 0, "a flush unit never wraps inside the ring");
static_assert(ROUND % DEPTH == 0, "a round is a whole number of passes over the descriptor registers");
static_assert(DEPTH % HOT_PERIOD == 0 && ROUND % HOT_PERIOD == 0, "pieces parsed in a full step are emitted DEPTH steps later, which must be a full step again");
// (With SUBS pieces per step:) a piece parsed at output position opp is emitted at most SUBS * DEPTH pieces (8 bytes each) later; every flush round leaves less than
// UNIT_BYTES unflushed per unit round and the emitter advances at most SUBS * 8 * ROUND bytes between two rounds, so
// flushed > opp - UNIT_BYTES - 8 * ROUND - 8 * DEPTH whenever a piece is parsed: a source that is not final in global memory then
// (two aligned words) is still in the ring at emit time.
static_assert(NEAR_WORDS * 8 >= UNIT_BYTES + SUBS * (8 * ROUND + 8 * DEPTH) + 24, "a source is either in the ring at emit time or final in global memory at parse time");

enum : int { E_OK = 0, E_TRUNCATED = 1, E_OFFSET = 2, E_OVERFLOW = 3, E_SIZE = 4, E_INTERNAL = 5, E_ENDRULE = 6 };

// piece kinds
//   general pieces (any alignment, 1..8 bytes):  K_DATA (bytes in the queue's data slot), K_NEAR (source in the ring), K_FARSLOW
//   word pieces (the fast path: one whole 8-byte aligned output word whose match source is 8-byte aligned too):
//     K_WORD   the word `offw` words back in the ring, its first `nl` bytes replaced by literal bytes from the data slot
//     K_WORDQ  the 8 bytes in the data slot (a source older than the ring, fetched at parse time)
enum : uint32_t { K_NONE = 0, K_DATA = 1, K_NEAR = 2, K_FARSLOW = 3, K_WORD = 4, K_WORDQ = 5 };

// a piece descriptor: kind | n << 3 | source position << 7; the 8 data bytes of a K_DATA piece travel beside it in the queue slot
LZL_HD uint32_t piece_kind(uint32_t m) { return m & 7u; }
LZL_HD uint32_t piece_n(uint32_t m) { return (m >> 3) & 15u; }
LZL_HD uint32_t piece_src(uint32_t m) { return m >> 7; }
LZL_HD uint32_t word_nl(uint32_t m) { return (m >> 3) & 7u; }       // K_WORD: leading literal bytes
LZL_HD uint32_t word_offw(uint32_t m) { return m >> 6; }            // K_WORD: distance of the source word in words

// 8 bytes starting `sh` bits (a multiple of 8, < 64) into the 16 bytes (lo, hi)
LZL_HD uint64_t funnel64(uint64_t lo, uint64_t hi, uint32_t sh)
{
#if defined(__CUDA_ARCH__)
    uint32_t a0 = (uint32_t)lo, a1 = (uint32_t)(lo >> 32), a2 = (uint32_t)hi, a3 = (uint32_t)(hi >> 32);
    if (sh & 32u) { a0 = a1; a1 = a2; a2 = a3; }
    const uint32_t r0 = __funnelshift_r(a0, a1, sh), r1 = __funnelshift_r(a1, a2, sh);   // (shift taken modulo 32)
    return (uint64_t)r0 | ((uint64_t)r1 << 32);
#else
    return sh ? (lo >> sh) | (hi << (64 - sh)) : lo;
#endif
}

LZL_HD uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }

// number of leading 0xFF bytes of x (0..8)
LZL_HD uint32_t leading_ff(uint64_t x)
{
    const uint64_t y = ~x;
    if (y == 0) return 8;
#if defined(__CUDA_ARCH__)
    return (uint32_t)(__ffsll((long long)y) - 1) >> 3;
#else
    return (uint32_t)__builtin_ctzll(y) >> 3;
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// parse side
// ---------------------------------------------------------------------------------------------------------------
enum : uint32_t { PS_IDLE = 0, PS_START, PS_TOKEN, PS_LITEXT, PS_LITSTART, PS_LIT, PS_MHDR, PS_MATEXT, PS_MATSTART, PS_MATCH, PS_END, PS_ERR };

struct Parser {
    uint32_t ip, ip_end;       // stream position of the next unread byte / compressed size
    uint32_t opp, op_end;      // output position of the next piece / origin
    uint32_t lit_rem, mat_rem, off;
    uint32_t st, tok_m, last, err;
    uint32_t win_fill;         // the window holds valid stream bytes below this position
    uint64_t w0, w1;           // stream bytes [ip & ~7, (ip & ~7) + 16)

    LZL_HD void reset(uint32_t comp_len, uint32_t origin)
    {
        ip = 0; ip_end = comp_len; opp = 0; op_end = origin;
        lit_rem = mat_rem = off = 0;
        st = PS_START; tok_m = 0; last = 0; err = 0;
        win_fill = 0; w0 = w1 = 0;
    }
    LZL_HD bool finished() const { return st == PS_END || st == PS_ERR; }

    template <class Mem>
    LZL_HD void advance(Mem &mem, uint32_t k)   // k <= 8
    {
        const uint32_t old = ip;
        ip += k;
        if ((ip ^ old) & ~7u) {
            w0 = w1;
            w1 = mem.win_read((ip & ~7u) + 8);
        }
    }
    LZL_HD uint64_t peek() const { return funnel64(w0, w1, (ip & 7u) * 8u); }
    LZL_HD void fail(int e) { st = PS_ERR; err = (uint32_t)e; }

    // Fast path: ONE WORD PIECE -- a whole aligned output word from an aligned source word, with up to 5 leading literal
    // bytes -- straight-line and without touching any state unless it succeeds.  Covers what LZ4 emits for 8-byte
    // columns: (0 literals, match 8k), (1, 7), (2, 6) ... with offsets that are multiples of 8, and the following words of a
    // longer aligned match.  Everything else (and everything that has to be refused) is left to step().  Same checks as
    // step(): the fast path is only taken when every one of them passes.
    template <class Mem>
    LZL_HD bool fast(Mem &mem, uint32_t flushed, uint32_t slot, uint32_t &m)
    {
        // Written without early exits: one flag per condition, state updated through selects, memory operations predicated --
        // the warp runs one lane-decoder per lane and a single warp per scheduler, so straight-line code (instruction-level
        // parallelism, no branch bubbles) is what makes a step short.
        const bool idle = st == PS_IDLE || st == PS_END || st == PS_ERR;
        const bool isM = st == PS_MATCH, isT = st == PS_TOKEN;
        const bool okM = isM && mat_rem >= 8u && ((opp | off) & 7u) == 0;
        const uint64_t x = peek();
        const uint32_t t = (uint32_t)x & 0xffu, L = t >> 4, Mn = t & 15u;
        const uint32_t o = (uint32_t)(x >> ((8u + 8u * L) & 63u)) & 0xffffu;      // (garbage when L > 5: not used then)
        const uint32_t total = L + Mn + 4u;
        const bool win_ok = !(ip + STEP_LOOKAHEAD > win_fill && win_fill < ip_end);
        // not the last sequence: the literal run ends at least 12 bytes before the end of the output and 8 before the end of the input
        const bool okT = isT && (opp & 7u) == 0 && win_ok && L <= 5u && Mn != 15u && ip + 1u + L + 8u <= ip_end && opp + L + 12u <= op_end &&
                         (o & 7u) == 0 && o != 0 && o <= opp && (total & 7u) == 0 && opp + total + 5u <= op_end;
        const uint32_t offx = okM ? off : o, Lx = okM ? 0u : L;
        const uint32_t offw = offx >> 3;
        const bool near = offw <= (uint32_t)NEAR_WORDS;
        const bool ok = (okM || okT) && (near || (Lx == 0 && opp - offx + 8u <= flushed));
        if (ok && Lx) mem.put_data(slot, x >> 8);
        if (ok && !near) mem.put_far(slot, opp - offx);
        m = idle ? (uint32_t)K_NONE : near ? (K_WORD | (Lx << 3) | (offw << 6)) : (uint32_t)K_WORDQ;
        const bool tok = ok && okT;
        // advance(3 + L) for a parsed token
        {
            const uint32_t nip = ip + (tok ? 3u + L : 0u);
            const bool cross = ((nip ^ ip) & ~7u) != 0;
            ip = nip;
            if (cross) {
                w0 = w1;
                w1 = mem.win_read((ip & ~7u) + 8u);
            }
        }
        off = tok ? o : off;
        tok_m = tok ? Mn : tok_m;
        mat_rem = tok ? total - 8u : (ok ? mat_rem - 8u : mat_rem);
        opp += ok ? 8u : 0u;
        st = ok ? (mat_rem ? (uint32_t)PS_MATCH : (uint32_t)PS_TOKEN) : st;
        return ok || idle;
    }

    // Hot path: TWO plain tokens at once -- token byte 0x04 (no literals, match of 8) with offsets that are multiples of 8, at an
    // aligned output position: what LZ4 emits for ~98 % of the values of an 8-byte column.  Both tokens sit in the 8 bytes at ip.
    // All or nothing: either both word pieces are made (true) or the lane's state is untouched (false; mA = mB = K_NONE).
    template <class Mem>
    LZL_HD bool fast2(Mem &mem, uint32_t flushed, uint32_t slotA, uint32_t slotB, uint32_t &mA, uint32_t &mB)
    {
        const uint64_t x = peek();
        const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
        const uint32_t o1 = (lo >> 8) & 0xffffu, o2 = hi & 0xffffu;
        const bool win_ok = !(ip + STEP_LOOKAHEAD > win_fill && win_fill < ip_end);
        // (bitwise, not short-circuit: the flags are independent and cheap, branches are what costs here)
        const bool shape = (st == PS_TOKEN) & ((opp & 7u) == 0) & win_ok & ((lo & 0xff0000ffu) == 0x04000004u) & (((o1 | o2) & 7u) == 0) &
                           (o1 != 0) & (o2 != 0) & (o1 <= opp) & (o2 <= opp + 8u) & (opp + 28u <= op_end) & (ip + 12u <= ip_end);
        const uint32_t ow1 = o1 >> 3, ow2 = o2 >> 3;
        const bool n1 = ow1 <= (uint32_t)NEAR_WORDS, n2 = ow2 <= (uint32_t)NEAR_WORDS;
        const bool f1 = opp - o1 + 8u <= flushed, f2 = opp + 16u - o2 <= flushed;
        const bool ok = shape & (n1 | f1) & (n2 | f2);
        if (ok && !n1) mem.put_far(slotA, opp - o1);
        if (ok && !n2) mem.put_far(slotB, opp + 8u - o2);
        mA = !ok ? (uint32_t)K_NONE : n1 ? (K_WORD | (ow1 << 6)) : (uint32_t)K_WORDQ;
        mB = !ok ? (uint32_t)K_NONE : n2 ? (K_WORD | (ow2 << 6)) : (uint32_t)K_WORDQ;
        {
            const uint32_t nip = ip + (ok ? 6u : 0u);
            const bool cross = ((nip ^ ip) & ~7u) != 0;
            ip = nip;
            if (cross) {
                w0 = w1;
                w1 = mem.win_read((ip & ~7u) + 8u);
            }
        }
        off = ok ? o2 : off;
        tok_m = ok ? 4u : tok_m;
        opp += ok ? 16u : 0u;
        return ok;
    }

    // One step: at most one piece, described by the return value; its data (if any) goes to queue slot `slot`.
    // `flushed` = output bytes of this block that are final in global memory.
    template <class Mem>
    LZL_HD uint32_t step(Mem &mem, uint32_t flushed, uint32_t slot)
    {
        const uint32_t none = K_NONE;
        if (st == PS_IDLE || st == PS_END || st == PS_ERR) return none;
        // the step reads at most STEP_LOOKAHEAD bytes past ip: wait for the refill unless the stream ends before that
        if (ip + STEP_LOOKAHEAD > win_fill && win_fill < ip_end) return none;
        if (st == PS_START) {
            w0 = mem.win_read(0);
            w1 = mem.win_read(8);
            st = PS_TOKEN;
        }
        if (st == PS_TOKEN) {
            if (ip >= ip_end) { fail(E_TRUNCATED); return none; }
            const uint64_t x = peek();
            const uint32_t t = (uint32_t)x & 0xffu;
            tok_m = t & 15u;
            // the common case in one go: no literals, no length extension, not near the end of either buffer
            if ((t >> 4) == 0 && tok_m != 15u && ip + 1 + 8 <= ip_end && opp + 12 <= op_end) {
                off = (uint32_t)(x >> 8) & 0xffffu;
                advance(mem, 3);
                if (off == 0 || off > opp) { fail(E_OFFSET); return none; }
                mat_rem = tok_m + 4;
                if (opp + mat_rem + 5 > op_end) { fail(mat_rem > op_end - opp ? E_OVERFLOW : E_ENDRULE); return none; }
                st = PS_MATCH;
            } else {
                advance(mem, 1);
                lit_rem = t >> 4;
                st = lit_rem == 15u ? PS_LITEXT : PS_LITSTART;
            }
        }
        if (st == PS_LITEXT) {
            const uint64_t x = peek();
            const uint32_t c = leading_ff(x);
            if (c == 8) {
                if (ip + 8 > ip_end) { fail(E_TRUNCATED); return none; }
                lit_rem += 255u * 8u;
                if (lit_rem > 0x7E000000u) { fail(E_OVERFLOW); return none; }
                advance(mem, 8);
                return none;
            }
            if (ip + c + 1 > ip_end) { fail(E_TRUNCATED); return none; }
            lit_rem += 255u * c + ((uint32_t)(x >> (8 * c)) & 0xffu);
            advance(mem, c + 1);
            st = PS_LITSTART;
        }
        if (st == PS_LITSTART) {
            if (lit_rem > op_end - opp) { fail(E_OVERFLOW); return none; }
            if (lit_rem > ip_end - ip) { fail(E_TRUNCATED); return none; }
            // a run that ends near the end of the output or of the input must be the last sequence
            last = (opp + lit_rem + 12 > op_end || ip + lit_rem + 8 > ip_end) ? 1u : 0u;
            if (last && ip + lit_rem != ip_end) { fail(E_ENDRULE); return none; }
            st = PS_LIT;
        }
        if (st == PS_LIT) {
            if (lit_rem > 0) {
                const uint32_t n = umin(8u, lit_rem);
                mem.put_data(slot, peek());
                const uint32_t p = K_DATA | (n << 3);
                advance(mem, n);
                lit_rem -= n;
                opp += n;
                if (lit_rem == 0) st = last ? PS_END : PS_MHDR;
                if (st == PS_END && opp != op_end) fail(E_SIZE);
                return p;
            }
            st = last ? PS_END : PS_MHDR;
            if (st == PS_END) { if (opp != op_end) fail(E_SIZE); return none; }
        }
        if (st == PS_MHDR) {
            // (not the last sequence: at least 8 stream bytes follow the literals, so the offset is inside the stream)
            off = (uint32_t)peek() & 0xffffu;
            advance(mem, 2);
            if (off == 0 || off > opp) { fail(E_OFFSET); return none; }
            mat_rem = tok_m;
            st = tok_m == 15u ? PS_MATEXT : PS_MATSTART;
        }
        if (st == PS_MATEXT) {
            const uint64_t x = peek();
            const uint32_t c = leading_ff(x);
            // every extension byte must lie more than 5 bytes before the end of the input
            if (c == 8) {
                if (ip + 7 + 5 >= ip_end) { fail(E_TRUNCATED); return none; }
                mat_rem += 255u * 8u;
                if (mat_rem > 0x7E000000u) { fail(E_OVERFLOW); return none; }
                advance(mem, 8);
                return none;
            }
            if (ip + c + 5 >= ip_end) { fail(E_TRUNCATED); return none; }
            mat_rem += 255u * c + ((uint32_t)(x >> (8 * c)) & 0xffu);
            advance(mem, c + 1);
            st = PS_MATSTART;
        }
        if (st == PS_MATSTART) {
            mat_rem += 4;
            if (mat_rem > op_end - opp) { fail(E_OVERFLOW); return none; }
            if (opp + mat_rem + 5 > op_end) { fail(E_ENDRULE); return none; }   // the last 5 bytes are literals
            st = PS_MATCH;
        }
        if (st == PS_MATCH) {
            // an overlapping match (offset < 8) repeats a pattern: after `off` bytes the pattern is there twice, so the
            // distance doubles until a full word can be copied
            const uint32_t n = umin(umin(8u, mat_rem), off);
            const uint32_t src = opp - off;
            uint32_t p;
            if ((opp >> 3) - (src >> 3) <= (uint32_t)NEAR_WORDS) {
                p = K_NEAR | (n << 3) | (src << 7);
            } else if ((src & ~7u) + 16 <= flushed) {
                if ((src & 7u) == 0) {
                    mem.put_far(slot, src);
                    p = K_DATA | (n << 3);
                } else {
                    p = K_FARSLOW | (n << 3) | (src << 7);
                }
            } else {
                fail(E_INTERNAL);   // (ruled out by the static_assert above)
                return none;
            }
            opp += n;
            mat_rem -= n;
            if (off < 8u) off += off;
            if (mat_rem == 0) st = PS_TOKEN;
            return p;
        }
        return none;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// emit side
// ---------------------------------------------------------------------------------------------------------------
struct Emitter {
    uint32_t op;        // output bytes emitted; every byte below op is in the ring (or flushed)
    uint32_t flushed;   // output bytes final in global memory (a multiple of UNIT_BYTES, may exceed op once at the end)
    uint32_t rw;        // ring slot of the word that holds op
    uint64_t acc;       // the partial word at op

    LZL_HD void reset() { op = 0; flushed = 0; rw = 0; acc = 0; }

    // word pieces (op is 8-byte aligned whenever one arrives: the parser only makes them at aligned positions); no early exits
    template <class Mem>
    LZL_HD bool fast(Mem &mem, uint32_t p, uint32_t slot)
    {
        const uint32_t kind = piece_kind(p);
        const bool isW = kind == K_WORD, ok = kind >= K_WORD;
        int s = (int)rw - (int)word_offw(p);
        s += s < 0 ? RING_WORDS : 0;
        const uint64_t q = mem.get_data(slot);                      // literal bytes (K_WORD) or the whole word (K_WORDQ)
        const uint64_t r = mem.ring_load(isW ? (uint32_t)s : rw);   // (any valid slot when the piece is not a K_WORD)
        const uint32_t nl = isW ? word_nl(p) : 8u;
        const uint64_t keep = nl >= 8u ? (uint64_t)0 : ~(uint64_t)0 << (8u * nl);
        const uint64_t w = (r & keep) | (q & ~keep);
        if (ok) mem.ring_store(rw, w);
        const uint32_t nrw = rw + 1 == (uint32_t)RING_WORDS ? 0 : rw + 1;
        rw = ok ? nrw : rw;
        op += ok ? 8u : 0u;
        return ok || kind == K_NONE;
    }

    // Hot path: the two pieces of a hot step -- both whole words without literal bytes (K_WORD with nl == 0 / K_WORDQ), or both K_NONE.
    template <class Mem>
    LZL_HD void fast2(Mem &mem, uint32_t pA, uint32_t pB, uint32_t slotA, uint32_t slotB)
    {
        const bool ok = piece_kind(pA) >= K_WORD;                 // (a hot step makes both pieces or none)
        const uint32_t rw1 = rw + 1 == (uint32_t)RING_WORDS ? 0 : rw + 1;
        int s1 = (int)rw - (int)word_offw(pA), s2 = (int)rw1 - (int)word_offw(pB);
        s1 += s1 < 0 ? RING_WORDS : 0;
        s2 += s2 < 0 ? RING_WORDS : 0;
        const uint64_t qA = mem.get_data(slotA), qB = mem.get_data(slotB);
        const uint64_t rA = mem.ring_load(piece_kind(pA) == K_WORD ? (uint32_t)s1 : rw);
        const uint64_t wA = piece_kind(pA) == K_WORD ? rA : qA;
        if (ok) mem.ring_store(rw, wA);
        const uint64_t rB = mem.ring_load(piece_kind(pB) == K_WORD ? (uint32_t)s2 : rw);    // (after the store: the source may be the word just written)
        const uint64_t wB = piece_kind(pB) == K_WORD ? rB : qB;
        if (ok) mem.ring_store(rw1, wB);
        const uint32_t rw2 = rw1 + 1 == (uint32_t)RING_WORDS ? 0 : rw1 + 1;
        rw = ok ? rw2 : rw;
        op += ok ? 16u : 0u;
    }

    template <class Mem>
    LZL_HD void step(Mem &mem, uint32_t p, uint32_t slot)
    {
        const uint32_t kind = piece_kind(p);
        if (kind == K_NONE) return;
        const uint32_t n = piece_n(p);
        uint64_t d;
        if (kind == K_DATA) {
            d = mem.get_data(slot);
        } else {
            const uint32_t src = piece_src(p);
            uint64_t lo, hi = 0;
            if (kind == K_NEAR) {
                int s = (int)rw - (int)((op >> 3) - (src >> 3));
                if (s < 0) s += RING_WORDS;
                lo = mem.ring_load((uint32_t)s);
                if (src & 7u) hi = mem.ring_load((uint32_t)(s + 1 == RING_WORDS ? 0 : s + 1));
            } else {
                lo = mem.out_load(src & ~7u);
                hi = mem.out_load((src & ~7u) + 8);
            }
            d = funnel64(lo, hi, (src & 7u) * 8u);
        }
        if (n < 8) d &= ((uint64_t)1 << (8 * n)) - 1;
        const uint32_t fill = op & 7u;
        acc |= d << (8 * fill);
        mem.ring_store(rw, acc);
        if (fill + n >= 8) {
            rw = rw + 1 == (uint32_t)RING_WORDS ? 0 : rw + 1;
            acc = fill ? d >> (8 * (8 - fill)) : 0;
            if (fill + n > 8) mem.ring_store(rw, acc);
        }
        op += n;
    }
    // complete units waiting for the flush
    LZL_HD bool unit_ready() const { return op >= flushed + UNIT_BYTES; }
    LZL_HD uint32_t flush_slot() const { return (flushed >> 3) % (uint32_t)RING_WORDS; }
};

}  // namespace lane
}  // namespace dfdb
