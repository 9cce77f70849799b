// lz4_decode_v3.cu -- K1, walker / consumer flavour: raw LZ4 block decode on sm_100a for every column the warp-per-block decoder
// (lz4_decode_spec.cu) is not made for: strings, literal-heavy bodies such as Union{Float64,Missing}, shifted Float64 grids,
// nullable integer columns.
//
// Replaces read_block's LZ4_decompress_safe call (/root/reference/src/io/BlockStreams.jl:101-119, liblz4 via
// CodecLz4) for whole batches of independent column blocks, with the same safety contract: never reads
// outside the compressed payload, never writes outside `origin`, per-block status instead of the
// reference's `@assert size == sizes.origin "decompression error"`.
//
// Why this shape.  The only inherently serial part of an LZ4 block is finding where each token starts (token k+1 starts
// 3 + literal_length(k) bytes after token k).  The first-generation kernel (lz4_decode.cu) let a whole warp walk that chain
// (3 instructions, ~35 cycles of latency per token, 1 useful lane of 32).  Here one persistent CTA per SM keeps NSLOT column
// blocks in flight at once: NWALK walker warps whose LANES walk the token chains of NSLOT blocks through shared-memory
// windows of their compressed streams and emit one 4-byte ring entry (the token's stream position) per sequence -- 32 chains
// per warp advance in the latency of one --, and NCONS consumer warps with SPC block slots each that take 32 entries at a
// time and materialise them lane-parallel.  "Word-regular" runs (8-byte aligned output, offset and length multiples of 8)
// are expanded to one output word per lane (far sources: one 8-byte load; in-batch sources: warp shuffles); everything else
// goes through the generic batches below.  Ring validity comes from the walker's release store of its entry count; the
// window fill level and restart commands are published with release stores.  (Until round 2 a second copy of this kernel,
// lz4_decode_v2.cu, tuned for word-regular columns only, sat beside it; the warp-per-block decoder took over those columns
// and the copy was removed -- on the one kind it still served, Union{Int64,Missing}, this kernel is the faster one.)
// Beyond that organisation:
//   * the walker gets past tokens with length extensions on its own when the extension bytes are in the window
//     (walker_resolve); the entry stays in the ring, the consumer recognises it by its nibbles.  Only the last
//     sequence of a block and extensions that reach beyond the window still park the lane until the consumer posts a
//     restart command.  The hint word carries two flags: PARKED and STARVED (next token beyond the window: the
//     consumer then takes a partial batch, because the window only moves when entries are consumed);
//   * a sequence that is done one at a time (special_run) reads the stream through a 256-byte register chunk -- no
//     dependent global loads for token / extension bytes / literals / offset -- and keeps a short match pending in
//     registers across the parse of the next sequence; while the walker is parked the warp carries on over the
//     following sequences;
//   * generic batches (generic_batch) fetch a match source with aligned 8-byte loads issued together -- from global
//     memory or from the staging area -- and store from registers; byte-serial copies remain only for overlapping
//     matches.  The partial 16-byte output chunk a batch ends in stays in the staging area for the next batch;
//   * in-batch chains of word-regular batches (sorted / sequential columns: every word copies its predecessor) are
//     collapsed by pointer jumping (resolve_chains, burst<true>) instead of one dependency wave per word;
//   * every block start flushes the ring (command -> acknowledgement with the ring position the new block's entries
//     start at), so a block that ends in an error is dropped where it stands.
// Which flavour a column gets is decided at load from a token sample (api.cu: sample_flavour).
//
// Algorithmic bytes per block (roofline): compressed bytes read + origin bytes written.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "kernels.cuh"
#include "lz4_common.cuh"

namespace dfdb {

namespace {

using namespace lz4;

constexpr int NWALK = 2;                    // walker warps
constexpr int NCONS = 30;                   // consumer warps
constexpr int SPC = 2;                      // block slots per consumer warp
constexpr int NSLOT = NCONS * SPC;          // 60 blocks in flight per SM
constexpr int NSLOT_PAD = NWALK * 32;
constexpr int V2_THREADS = (NWALK + NCONS) * 32;
constexpr int W = 1024, WM = W - 1;         // compressed-stream window per slot (circular, by stream position)
constexpr int R = 128, RM = R - 1;          // ring entries per slot
constexpr int MAX_BURST = 16;
constexpr int READY_MIN = 96;               // ring entries that wake a consumer up (R - READY_MIN is the walker's slack)
constexpr unsigned IDLE_SLEEP_NS = 1000;     // consumer with nothing ready: about the time a walker needs for one batch               // batches a consumer takes from one slot before it looks at the other
constexpr int STG = 1088;                   // staging per consumer warp: 15 carried + 32 * (14 + 18) bytes, padded
constexpr uint32_t LIM_EXIT = 0xffffffffu;
// hint word (walker -> consumer): entries emitted so far | flags
constexpr uint32_t H_PARKED = 0x80000000u;  // the last entry is a special token the walker could not get past: it waits for a command
constexpr uint32_t H_STARVED = 0x40000000u; // the walker's next token lies beyond the window
constexpr uint32_t H_CNT = 0x3fffffffu;
constexpr int HYST = 8;                     // plain tokens in a row that end a consumer-side run of special sequences
constexpr uint32_t POS_CAP = 1u << 30;      // larger blocks take the one-sequence-at-a-time path
constexpr int REG_MIN = 16;                 // shortest leading word-regular run worth its own batch
static_assert(NSLOT <= NSLOT_PAD, "every slot needs a walker lane");
static_assert(V2_THREADS <= 1024, "one CTA");

enum { SLOT_EMPTY = 0, SLOT_ACTIVE = 1, SLOT_RETIRED = 2 };

// diagnostics (DFDB_LZ4_STATS=1): per-launch totals
enum { ST_W_ROUNDS = 0, ST_W_COMMITS, ST_W_RINGFULL, ST_W_WINEMPTY, ST_W_PARKED, ST_W_SLEEPS,
       ST_C_POLLS, ST_C_SLEEPS, ST_C_PS_CALLS, ST_C_PS_FALSE, ST_C_REG, ST_C_REGSEQ, ST_C_GEN, ST_C_GENSEQ, ST_C_SPECIAL,
       ST_C_PS_CYCLES, ST_C_LOOP_CYCLES, ST_C_REFILLS, ST_W_CYCLES, ST_C_START_CYCLES, ST_C_SLOW_CALLS, ST_C_SLOW_CYCLES,
       ST_C_SPECIAL_CYCLES, ST_C_SLEEP_CYCLES, ST_G_PRO, ST_G_HEAD, ST_G_BATCH, ST_G_FLUSH, ST_G_EPI, ST_G_WAVES, ST_COUNT };
#ifdef DFDB_LZ4_STATS
__device__ unsigned long long *g_stats;
#define STAT_ADD(i, v) do { if (g_stats) atomicAdd(&g_stats[i], (unsigned long long)(v)); } while (0)
#define STATS_ON (g_stats != nullptr)
// per-batch counters are kept per warp and added to the global totals once, when the warp retires: a global atomic per
// batch from every warp perturbs what it measures
__device__ unsigned long long g_warp_stats[148 * 32][8];
#define WSTAT_ADD(k, v) do { if (g_stats) g_warp_stats[blockIdx.x * 32 + (threadIdx.x >> 5)][k] += (unsigned long long)(v); } while (0)
#else
#define STAT_ADD(i, v) do { } while (0)
#define WSTAT_ADD(k, v) do { } while (0)
#define STATS_ON false
#endif

struct SlotJob {                // consumer-private state of one block slot
    const uint8_t *src;
    uint8_t *dst;
    int32_t *status;
    uint32_t comp_len, origin;
    uint32_t op;                // output bytes produced so far (global memory is complete below op)
    uint32_t ip;                // stream position of the first unconsumed token
    uint32_t tail;              // ring entries consumed
    uint32_t whi;               // window holds stream bytes [whi - W, whi)
    uint32_t state, err, seq;
    uint32_t pend;              // window fill level once the refill in flight lands (0 = none in flight)
    uint32_t rphase;            // parity of the slot's refill mbarrier
    uint32_t chainy, pad[2];    // burst flavour: batches are chains of in-batch copies
};

struct V2Smem {
    __align__(1024) uint8_t win[NSLOT][W];   // first, 1024-byte aligned: the walker forms addresses with one LOP3
    uint32_t tail[NSLOT_PAD], whi[NSLOT_PAD], cmd_seq[NSLOT_PAD], cmd_p[NSLOT_PAD], cmd_lim[NSLOT_PAD];
    uint32_t cmd_ack[NSLOT_PAD], cmd_head[NSLOT_PAD];   // walker -> consumer: last command taken, and the ring position its entries start at
    uint32_t dummy[NSLOT_PAD];               // sink for the ring stores of walker lanes that do not commit a step
    uint32_t hint[NSLOT_PAD];                // walker -> consumer: entries emitted so far | parked << 31 (release store once per round)
    SlotJob job[NSLOT];
    unsigned long long rbar[NSLOT];          // mbarrier per slot: completion of the asynchronous window refill (32 arrivals)
    uint32_t ring[NSLOT][R + 1];             // token positions; +1: consecutive slots start one bank apart
    uint32_t stg_slot[NCONS], stg_op[NCONS];  // whose output chunk sits in stg[c][0..16): slot and output position (generic batches)
    __align__(16) uint8_t stg[NCONS][STG];
};

// ---- shared-memory accessors with explicit semantics (cross-warp traffic must not live in registers) ----
__device__ __forceinline__ uint32_t ld_acq(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_addr(p)) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_rlx(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_addr(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_rel(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_addr(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void st_rlx(uint32_t *p, uint32_t v)
{
    asm volatile("st.relaxed.cta.shared.u32 [%0], %1;" ::"r"(smem_addr(p)), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t sa)
{
    uint32_t v;
    asm volatile("ld.volatile.shared.u8 %0, [%1];" : "=r"(v) : "r"(sa) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t sa, uint32_t x)
{
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(sa), "r"(x) : "memory");
}

// Out of line (the walker's loop should stay small): the lane stopped on the token of ring entry head - 1.  Reads its
// length extension bytes when they are in the window and returns true with p behind the sequence; false = stay parked.
__device__ __noinline__ bool walker_resolve(uint32_t win_sa, uint32_t ring_sa, uint32_t head, uint32_t whi_c, uint32_t lim, uint32_t &p)
{
    uint32_t ps;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(ps) : "r"(ring_sa + ((head - 1u) & RM) * 4) : "memory");
    const uint32_t t = lds_u8(win_sa | (ps & WM));
    uint32_t q = ps + 1, L = t >> 4;
    if (L == 15u) {
        for (;;) {
            if (q >= whi_c) return false;
            const uint32_t e = lds_u8(win_sa | (q & WM));
            q++;
            L += e;
            if (e != 255u) break;
            if (L > 65536u) return false;
        }
    }
    q += L + 2;                                   // behind the offset
    if (q > lim) return false;                    // last sequence (or a truncated stream): the consumer closes the block
    if ((t & 15u) == 15u) {
        for (;;) {
            if (q >= whi_c || q >= lim) return false;
            const uint32_t e = lds_u8(win_sa | (q & WM));
            q++;
            if (e != 255u) break;
        }
    }
    p = q;
    return true;
}

// =====================================================================================================
// walker: lane = block slot.  Emits {p | token << 24, o | gen << 24 | special << 31} per sequence.
// =====================================================================================================
__device__ void walker(V2Smem &S, int slot)
{
    constexpr int UNROLL = 8;
    const bool has = slot < NSLOT;
    const int sl = has ? slot : 0;
    const uint32_t win_sa = smem_addr(S.win[sl]);     // 1024-byte aligned: window address = win_sa | (p & WM)
    const uint32_t ring_sa = smem_addr(S.ring[sl]);
    const uint32_t dummy_sa = smem_addr(&S.dummy[threadIdx.x & (NSLOT_PAD - 1)]);   // where non-committing lanes store
    uint32_t p = 0, lim = 0, head = 0, seen = 0, whi_c = 0, tail_c = 0;
    bool running = false, finished = !has;
#ifdef DFDB_LZ4_STATS
    unsigned int st_rounds = 0, st_full = 0, st_empty = 0, st_parked = 0, st_sleeps = 0;
    const long long st_t0 = clock64();
#endif
    for (;;) {
        // flow control and commands, once per UNROLL steps; all lanes issue the same three loads (no divergence)
        // (the command sequence first: a new command's window level must not be older than the command)
        const uint32_t seq_n = ld_acq(&S.cmd_seq[sl]), whi_n = ld_acq(&S.whi[sl]), tail_n = ld_rlx(&S.tail[sl]);
        if (running) { whi_c = whi_n; tail_c = tail_n; }
        if (!finished && seq_n != seen) {
            seen = seq_n;
            lim = ld_rlx(&S.cmd_lim[sl]);
            if (lim == LIM_EXIT) { finished = true; running = false; }
            else { p = ld_rlx(&S.cmd_p[sl]); running = true; whi_c = whi_n; tail_c = tail_n; }
            // a command flushes the ring: whatever this lane emitted so far belongs to an abandoned block
            if (has) { st_rlx(&S.hint[sl], head); st_rlx(&S.cmd_head[sl], head); st_rel(&S.cmd_ack[sl], seq_n); }
        }
        const bool run0 = running;
        bool starved = false;
        const uint32_t room = running ? (uint32_t)R - (head - tail_c) : 0u;
        const uint32_t head0 = head;
#ifdef DFDB_LZ4_STATS
        if (STATS_ON && has && !finished) {
            st_rounds++;
            if (!running) st_parked++;
            else if (room == 0) st_full++;
            else if (p >= whi_c) st_empty++;
        }
#endif
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            // The walker does nothing but the chain: one ring entry = the stream position of a token.  Straight-line
            // and branch-free: the token load is speculative (any window address is readable), a lane that does not
            // commit stores to a dummy word and keeps its state.  Dependent chain of a step: LDS -> LEA.HI/VIADD
            // (next p) -> SEL -> LOP3 (next address).
            const uint32_t t = lds_u8(win_sa | (p & WM));
            const bool inwin = p < whi_c;
            if (u == UNROLL - 1) starved = running & !inwin;   // (as of the last step: a lane that ran dry later shows up a round late)
            const bool commit = running & inwin & (room > (uint32_t)u);
            const uint32_t pn = p + 3 + (t >> 4);
            const bool special = (t >= 0xf0u) | ((t & 15u) == 15u) | (pn > lim);   // length extensions / last sequence
            sts_u32(commit ? ring_sa + (head & RM) * 4 : dummy_sa, p);
            head += commit ? 1u : 0u;
            p = commit ? pn : p;                               // a parked lane gets a fresh position with its next command
            running = running & !(commit & special);           // parked until the consumer posts the position behind this sequence
        }
        const bool any_commit = head != head0;
        // Resolve step (rare, divergent): a lane that stopped on a token with length extensions in this round reads the
        // extension bytes itself when they are in the window, and walks on behind the sequence.  The entry stays in the
        // ring; the consumer recognises it by its nibbles and does that sequence on its own.  Not resolvable (extension
        // bytes beyond the window, last sequence): the lane stays parked and the consumer restarts it with a command.
        if (run0 & !running) running = walker_resolve(win_sa, ring_sa, head, whi_c, lim, p);
        if (has)   // entries below head are visible
            st_rel(&S.hint[sl], head | ((!running && !finished) ? H_PARKED : 0u) | (starved ? H_STARVED : 0u));
        if (__all_sync(FULL, finished)) break;
        if (!__any_sync(FULL, any_commit)) {
            __nanosleep(60);
#ifdef DFDB_LZ4_STATS
            st_sleeps++;
#endif
        }
    }
#ifdef DFDB_LZ4_STATS
    if (STATS_ON && has) {
        STAT_ADD(ST_W_ROUNDS, st_rounds); STAT_ADD(ST_W_COMMITS, head); STAT_ADD(ST_W_RINGFULL, st_full);
        STAT_ADD(ST_W_WINEMPTY, st_empty); STAT_ADD(ST_W_PARKED, st_parked);
        if ((threadIdx.x & 31) == 0) { STAT_ADD(ST_W_SLEEPS, st_sleeps); STAT_ADD(ST_W_CYCLES, clock64() - st_t0); }
    }
#endif
}

// =====================================================================================================
// consumer
// =====================================================================================================

// ---- asynchronous window refill: LDGSTS (cp.async) straight into the circular window, completion on the slot's
//      mbarrier, fill level published to the walker when the data has landed -----------------------------------
__device__ __forceinline__ bool refill_landed(V2Smem &S, int s, SlotJob &J, bool block)
{
    if (!J.pend) return false;
    const uint32_t bar = smem_addr(&S.rbar[s]);
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(J.rphase)
            : "memory");
        ok = __all_sync(FULL, ok != 0);
    } while (block && !ok);
    if (!ok) return false;
    if (lane_id() == 0) {
        J.whi = J.pend;
        st_rel(&S.whi[s], J.pend);
        J.pend = 0;
        J.rphase ^= 1u;
    }
    __syncwarp();
    return true;
}

// Start loading up to 512 more stream bytes (whole warp; at most one refill in flight per slot).  Returns bytes requested.
__device__ __forceinline__ uint32_t refill_issue(V2Smem &S, int s, SlotJob &J, uint32_t min_n)
{
    const uint32_t lane = lane_id();
    if (J.pend) return 0;
    const uint32_t whi = J.whi;
    const uint32_t whi_max = (J.comp_len + 16) & ~15u;          // strictly beyond the last stream byte
    if (whi >= whi_max) return 0;
    const uint32_t room = (J.ip & ~15u) + (uint32_t)W - whi;     // bytes below J.ip are consumed
    uint32_t n = room < 512u ? room : 512u;
    if (whi_max - whi < n) n = whi_max - whi;
    if (n == 0 || (n < min_n && whi + n < whi_max)) return 0;
    if (lane * 16 < n) {
        const uint32_t pos = whi + lane * 16;
        const bool in = pos < ((J.comp_len + 15) & ~15u);       // payload slots are padded to 16 bytes; beyond: zero fill
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr(&S.win[s][pos & WM])),
                     "l"(J.src + (in ? pos : 0u)), "r"(in ? 16 : 0)
                     : "memory");
    }
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_addr(&S.rbar[s])) : "memory");
    if (lane == 0) J.pend = whi + n;
    __syncwarp();
    return n;
}

// Synchronous flavour for the rare places that cannot proceed without the bytes.
__device__ __forceinline__ uint32_t refill(V2Smem &S, int s, SlotJob &J, uint32_t min_n)
{
    refill_landed(S, s, J, true);
    const uint32_t n = refill_issue(S, s, J, min_n);
    if (n) refill_landed(S, s, J, true);
    return n;
}

__device__ __forceinline__ void post_cmd(V2Smem &S, int s, SlotJob &J, uint32_t p, uint32_t lim)
{
    if (lane_id() == 0) {
        st_rlx(&S.cmd_p[s], p);
        st_rlx(&S.cmd_lim[s], lim);
        J.seq += 1;
        st_rel(&S.cmd_seq[s], J.seq);
    }
    __syncwarp();
}

__device__ __forceinline__ void finish_block(SlotJob &J, int e)
{
    if (lane_id() == 0) { *J.status = e; J.state = SLOT_EMPTY; }
    __syncwarp();
}

// Generic batches assemble their output in the warp's staging area; flush every 16-byte chunk of
// [stg_base, new_op) -- the partial last chunk too, so that global memory is always complete below op.
__device__ __forceinline__ void flush(const SlotJob &J, const uint8_t *stg, uint32_t stg_base, uint32_t new_op)
{
    const uint32_t lane = lane_id();
    const int nchunks = (int)((new_op + 15u) >> 4) - (int)(stg_base >> 4);
    uint4 *d16 = reinterpret_cast<uint4 *>(J.dst + stg_base);
    const uint4 *g16 = reinterpret_cast<const uint4 *>(stg);
    for (int c = lane; c < nchunks; c += 32) d16[c] = g16[c];
    __syncwarp();
}

// In-batch sources of a word-regular batch.  Word i is  lit_i | (word[dep_i] & keep_i)  (low half; the high half is a plain
// copy).  Dependency waves over warp shuffles settle the usual case (a producer a few words back that came from
// memory) in one or two rounds; when many words are still open after the first wave they form chains (sorted /
// sequential columns: every word copies its predecessor), and those are collapsed by pointer jumping -- the maps compose, (lit, keep) o (lit', keep') = (lit | lit' & keep, keep & keep')
// -- in at most five more rounds instead of one round per word.
__device__ __noinline__ void resolve_chains(bool &fin, int dep, uint32_t litw, uint32_t keep, uint32_t &vlo, uint32_t &vhi)
{
    const uint32_t lane = lane_id();
    while (__any_sync(FULL, !fin)) {
        const uint32_t finmask = __ballot_sync(FULL, fin);
        const int j = fin ? (int)lane : dep;
        const uint32_t a = __shfl_sync(FULL, vlo, j), b = __shfl_sync(FULL, vhi, j);
        const uint32_t pl = __shfl_sync(FULL, litw, j), pk = __shfl_sync(FULL, keep, j);
        const int pd = __shfl_sync(FULL, dep, j);
        if (!fin) {
            if ((finmask >> j) & 1u) { vlo = litw | (a & keep); vhi = b; fin = true; }
            else { litw |= pl & keep; keep &= pk; dep = pd; }
        }
    }
}
__device__ __forceinline__ void resolve_in_batch(bool &fin, int dep, uint32_t litw, uint32_t keep, uint32_t &vlo, uint32_t &vhi)
{
    const uint32_t lane = lane_id();
    bool first = true;
    for (;;) {
        const uint32_t finmask = __ballot_sync(FULL, fin);
        if (finmask == FULL) return;
        // many words still open after the first wave: chains, not the odd producer a few words back
        if (!first && __popc(~finmask) > 12) { resolve_chains(fin, dep, litw, keep, vlo, vhi); return; }
        first = false;
        const int j = fin ? (int)lane : dep;
        const uint32_t a = __shfl_sync(FULL, vlo, j), b = __shfl_sync(FULL, vhi, j);
        if (!fin && ((finmask >> j) & 1u)) { vlo = litw | (a & keep); vhi = b; fin = true; }
    }
}

// ---- word-regular batch: lanes [0, nreg) hold sequences with o % 8 == 0, off % 8 == 0, (L + M) % 8 == 0, L <= 2.
//      One output word per lane, written straight to global memory (a coalesced 256-byte store per batch). ----
__device__ __forceinline__ int regular_batch(const SlotJob &J, uint32_t o0, uint32_t o, uint32_t len, uint32_t L, uint32_t offv,
                                             uint32_t lit, int nreg, uint32_t *new_op)
{
    const uint32_t lane = lane_id();
    const uint32_t fw = (o - o0) >> 3;                        // first output word of my sequence
    const uint32_t tw = __shfl_sync(FULL, fw + (len >> 3), nreg - 1);   // words in the batch (<= 32)
    if (o0 + 8 * tw > J.origin) return E_OVERFLOW;
    const bool wl = lane < tw;                                // from here on: lane = output word
    uint32_t offw = offv, Lw = L, litw = lit;
    if ((int)tw != nreg) {                                    // some sequence spans two words: expand sequences to words
        const uint32_t startmask = __reduce_or_sync(FULL, (int)lane < nreg ? (1u << fw) : 0u);
        int sq = __popc(startmask & (0xffffffffu >> (31 - lane))) - 1;
        if (!wl) sq = 0;
        const uint32_t pk = __shfl_sync(FULL, offv | (L << 16) | (fw << 24), sq);
        const uint32_t lit_s = __shfl_sync(FULL, lit, sq);
        const bool k0 = lane == (pk >> 24);                   // first word of its sequence carries the literals
        offw = pk & 0xffffu;
        Lw = k0 ? ((pk >> 16) & 0xffu) : 0u;
        litw = k0 ? lit_s : 0u;
    }
    const uint32_t keep = 0xffffffffu << (8 * Lw);            // Lw <= 2
    const int dep = (int)lane - (int)(offw >> 3);             // producer word inside the batch, or < 0: already in memory
    const uint32_t ow = o0 + 8 * lane;
    bool fin = !wl;
    uint32_t vlo = 0, vhi = 0;
    if (wl && dep < 0) {
        const uint2 far = __ldcg(reinterpret_cast<const uint2 *>(J.dst + (ow - offw)));
        vlo = litw | (far.x & keep);
        vhi = far.y;
        fin = true;
    }
    resolve_in_batch(fin, dep, litw, keep, vlo, vhi);
    if (wl) *reinterpret_cast<uint2 *>(J.dst + ow) = make_uint2(vlo, vhi);
    __syncwarp();
    *new_op = o0 + 8 * tw;
    return E_OK;
}

// ---- generic batch: lanes [0, g) hold simple sequences of any alignment; byte-granular dependency waves ----
__device__ __forceinline__ int generic_batch(const uint8_t *win, uint8_t *stg, const SlotJob &J, uint32_t o0, uint32_t lit_pos, uint32_t off_pos,
                                             uint32_t o, uint32_t L, uint32_t M, int g, uint32_t *new_op)
{
    const uint32_t lane = lane_id();
    const bool active = (int)lane < g;
    const int32_t stg_base = (int32_t)(o0 & ~15u);
    const uint32_t off = active ? ((uint32_t)win[off_pos & WM] | ((uint32_t)win[(off_pos + 1) & WM] << 8)) : 1u;
    const uint32_t end = __shfl_sync(FULL, o + L + M, g - 1);
    if (end > J.origin) return E_OVERFLOW;
    const int32_t m_dst = (int32_t)(o + L), m_src = m_dst - (int32_t)off;
    const bool bad = active && (off == 0 || m_src < 0);
    if (__any_sync(FULL, bad)) return E_OFFSET;
    // Literals: short runs (no length extension) lane by lane, long ones by the whole warp, one sequence after the other.
    if (active && L < 15u)
        for (uint32_t i = 0; i < L; i++) stg[(int32_t)o - stg_base + (int32_t)i] = win[(lit_pos + i) & WM];
    uint32_t longl = __ballot_sync(FULL, active && L >= 15u);
    while (longl) {
        const int k = __ffs(longl) - 1;
        longl &= longl - 1;
        const uint32_t kL = __shfl_sync(FULL, L, k), kpos = __shfl_sync(FULL, lit_pos, k);
        const int32_t kd = __shfl_sync(FULL, (int32_t)o - stg_base, k);
        for (uint32_t i = lane; i < kL; i += 32) stg[kd + (int32_t)i] = win[(kpos + i) & WM];
    }
    __syncwarp();
    // Matches.  (1) Every sequence whose source lies wholly below the staging area copies it from global memory, all of
    // them at once: aligned 8-byte loads issued together (only the words that hold source bytes, so nothing beyond the
    // bytes already written is touched), bytes stored from registers.  (2) The others -- sources inside the batch or
    // straddling its start -- go in stream order, the whole warp on one sequence, lane = byte: no dependency analysis, no
    // waves in which a handful of lanes work while 32 pay, and a self-overlapping match is just a source index modulo
    // its offset.
    const bool far = active && M <= 18u && m_src + (int32_t)M <= stg_base;
    if (far) {
        const int32_t sd = m_dst - stg_base;
        const uint32_t sh = (uint32_t)m_src & 7u, span = sh + M;                          // M <= 18: span <= 25
        const unsigned long long *g8 = reinterpret_cast<const unsigned long long *>(J.dst + (m_src - (int32_t)sh));
        unsigned long long w0 = __ldcg(g8), w1 = 0, w2 = 0, w3 = 0;
        if (span > 8u) w1 = __ldcg(g8 + 1);
        if (span > 16u) w2 = __ldcg(g8 + 2);
        if (span > 24u) w3 = __ldcg(g8 + 3);
        const uint32_t s8b = sh * 8;
        if (s8b) {
            w0 = (w0 >> s8b) | (w1 << (64 - s8b));
            w1 = (w1 >> s8b) | (w2 << (64 - s8b));
            w2 = (w2 >> s8b) | (w3 << (64 - s8b));
        }
#pragma unroll
        for (int i = 0; i < 18; i++) {
            const unsigned long long w = i < 8 ? w0 : (i < 16 ? w1 : w2);
            if ((uint32_t)i < M) stg[sd + i] = (uint8_t)(w >> (8 * (i & 7)));
        }
    }
    __syncwarp();
    uint32_t rest = __ballot_sync(FULL, active && !far);
    const uint32_t pk = (uint32_t)(m_dst - stg_base) | (M << 16);                         // staging index < 2048, M < 512
    // Two short sequences go in one step (lanes 0-15 / 16-31) when the second does not read what the first writes: every
    // lane checks that against its predecessor among the sequences left.
    uint32_t pairable;
    {
        const uint32_t below = rest & ((1u << lane) - 1u);
        const int prev = below ? 31 - __clz(below) : 0;
        const uint32_t qp = __shfl_sync(FULL, pk, prev);
        const int32_t pdst = (int32_t)(qp & 0xffffu), pM = (int32_t)(qp >> 16);
        const int32_t s0 = m_src - stg_base, s1 = s0 + (int32_t)(off < M ? off : M);          // my source bytes (staging indices)
        pairable = __ballot_sync(FULL, below && M <= 16u && pM <= 16 && (s1 <= pdst || s0 >= pdst + pM));
    }
    while (rest) {
        const int k = __ffs(rest) - 1;
        rest &= rest - 1;
        const int k2 = rest ? __ffs(rest) - 1 : 0;
        const bool pair = rest && ((pairable >> k2) & 1u);
        if (pair) rest &= rest - 1;
        const int sel = pair && lane >= 16 ? k2 : k;
        const uint32_t q = __shfl_sync(FULL, pk, sel), koff = __shfl_sync(FULL, off, sel);
        const uint32_t kM = q >> 16;
        const int32_t kdst = (int32_t)(q & 0xffffu);
        // the bytes [m_src, m_src + min(off, M)) are final, and byte i of the match is byte i mod off of them
        for (uint32_t i = pair ? (lane & 15u) : lane; i < kM; i += pair ? 16u : 32u) {
            const uint32_t r = koff < kM ? i % koff : i;
            const int32_t x = kdst - (int32_t)koff + (int32_t)r;      // staging index of the source byte (negative: below the staging area)
            stg[kdst + (int32_t)i] = x >= 0 ? stg[x] : __ldcg(J.dst + (stg_base + x));
        }
        __syncwarp();
        if (STATS_ON && lane == 0) WSTAT_ADD(5, 1);
    }
    *new_op = end;
    return E_OK;
}

// ---- one sequence at a time, whole warp: the path of "special" tokens (length extensions, last sequence) ------------
// The stream is read through a 256-byte register chunk (one aligned 8-byte word per lane, bytes fetched with warp
// shuffles), so a sequence costs no dependent global loads for its token / extension bytes / literals / offset; the
// match bytes of a short match stay pending in registers across the parse of the next sequence (their load latency
// hides behind it).  Same checks, in the same order, as decode_one_sequence.
__device__ __forceinline__ uint2 chunk_load(const uint8_t *src, uint32_t lim16, uint32_t base)
{
    const uint32_t a = base + 8 * lane_id();                  // base is a multiple of 8, payload slots are 16-byte aligned and padded
    uint2 w = make_uint2(0u, 0u);
    if (a < lim16) w = __ldg(reinterpret_cast<const uint2 *>(src + a));
    const uint32_t nx = base + 256 + 128 * lane_id();         // the next chunk's two lines: into L2 while this one is parsed
    if (lane_id() < 2 && nx < lim16) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + nx));
    return w;
}
__device__ __forceinline__ uint32_t chunk_byte_u(const uint2 w, uint32_t k)     // k < 256, uniform over the warp
{
    return (__shfl_sync(FULL, (k & 4u) ? w.y : w.x, k >> 3) >> ((k & 3u) * 8)) & 0xffu;
}
__device__ __forceinline__ uint32_t chunk_byte_v(const uint2 w, uint32_t k)     // k < 256, any value per lane
{
    const uint32_t lo = __shfl_sync(FULL, w.x, k >> 3), hi = __shfl_sync(FULL, w.y, k >> 3);
    return (((k & 4u) ? hi : lo) >> ((k & 3u) * 8)) & 0xffu;
}

// Decodes the sequence at ip.  single: just that one (the walker got past it on its own and has gone on emitting
// entries).  Otherwise the walker is parked behind it, and the run goes on until HYST plain tokens in a row have been
// seen (a walker restart costs far more than a sequence done here).
// Returns E_*; on E_OK either done (last sequence consumed) or ip is the position of the next token.
__device__ __noinline__ int special_run(const SlotJob &J, uint32_t &ip_io, uint32_t &op_io, bool &done, bool single)
{
    const uint32_t lane = lane_id();
    const uint8_t *const src = J.src;
    uint8_t *const dst = J.dst;
    const uint32_t comp_len = J.comp_len, origin = J.origin, lim16 = (comp_len + 15u) & ~15u;
    uint32_t ip = ip_io, op = op_io;
    uint32_t base = ip & ~7u, k = ip - base;
    uint2 w = chunk_load(src, lim16, base);
    uint32_t pend_b = 0;
    uint8_t *pend_a = nullptr;                                // this lane's pending match byte (nullptr: none)
    int err = E_OK;
    bool first = true;
    int nplain = 0;
#define DFDB_ENSURE(n)                                                                                   \
    if (k + (n) > 256u) {                                                                                \
        const uint32_t pos_ = base + k;                                                                  \
        base = pos_ & ~7u; k = pos_ - base;                                                              \
        w = chunk_load(src, lim16, base);                                                                \
    }
    for (;;) {
        ip = base + k;
        if (ip >= comp_len) { err = E_TRUNCATED; break; }
        DFDB_ENSURE(1u)
        const uint32_t t = chunk_byte_u(w, k);
        uint32_t L = t >> 4;
        if (!first) {
            if (single) break;
            if (t < 0xf0u && (t & 15u) != 15u && ip + 3u + L <= comp_len) {
                if (++nplain > HYST) break;                                               // plain tokens: back to the walker
            } else nplain = 0;
        }
        first = false;
        k++;
        if (L == 15u) {
            for (;;) {
                if (base + k >= comp_len) { err = E_TRUNCATED; break; }
                DFDB_ENSURE(1u)
                const uint32_t e = chunk_byte_u(w, k);
                k++;
                L += e;
                if (e != 255u) break;
                if (L > origin) break;                                                    // caught as overflow below
            }
            if (err) break;
        }
        if (base + k + L > comp_len || base + k + L < L) { err = E_TRUNCATED; break; }
        if (op + L > origin) { err = E_OVERFLOW; break; }
        if (L <= 240u) {
            DFDB_ENSURE(L)
            for (uint32_t j0 = 0; j0 < L; j0 += 32) {
                const uint32_t j = j0 + lane, kk = k + j < 255u ? k + j : 255u;
                const uint32_t b = chunk_byte_v(w, kk);
                if (j < L) dst[op + j] = (uint8_t)b;
            }
        } else {
            warp_copy(dst + op, src + base + k, (int64_t)L);
        }
        k += L;
        op += L;
        ip = base + k;
        if (ip == comp_len) { done = true; break; }                                       // last sequence: literals only
        if (ip + 2u > comp_len) { err = E_TRUNCATED; break; }
        DFDB_ENSURE(2u)
        const uint32_t off = chunk_byte_u(w, k) | (chunk_byte_u(w, k + 1) << 8);
        k += 2;
        uint32_t M = t & 15u;
        if (M == 15u) {
            for (;;) {
                if (base + k >= comp_len) { err = E_TRUNCATED; break; }
                DFDB_ENSURE(1u)
                const uint32_t e = chunk_byte_u(w, k);
                k++;
                M += e;
                if (e != 255u) break;
                if (M > origin) break;
            }
            if (err) break;
        }
        M += 4;
        if (off == 0 || off > op) { err = E_OFFSET; break; }
        if (op + M > origin) { err = E_OVERFLOW; break; }
        if (pend_a) { *pend_a = (uint8_t)pend_b; pend_a = nullptr; }
        __syncwarp();                                          // literals and the previous match are visible to the whole warp
        // every source byte is < op, i.e. final: the copy is fully parallel even when it overlaps itself
        uint8_t *m_dst = dst + op;
        const uint8_t *m_src = m_dst - off;
        if (M <= 32u) {
            if (lane < M) { pend_b = __ldcg(m_src + (off >= M ? lane : lane % off)); pend_a = m_dst + lane; }
        } else if (off >= M) {
            for (uint32_t i = lane; i < M; i += 32) m_dst[i] = __ldcg(m_src + i);
        } else {
            for (uint32_t i = lane; i < M; i += 32) m_dst[i] = __ldcg(m_src + (i % off));
        }
        op += M;
    }
#undef DFDB_ENSURE
    if (pend_a) *pend_a = (uint8_t)pend_b;
    __syncwarp();
    ip_io = base + k;
    op_io = op;
    return err;
}

// A special entry (length extensions / last sequence) heads the ring: the whole warp does that sequence against global
// memory.  parked: the walker waits behind it, so the warp may carry on over the following sequences before it restarts
// the walker lane (or finishes the block).  Not parked: the walker has already gone on; just that one sequence.
__device__ __noinline__ void special_step(V2Smem &S, int s, SlotJob &J, uint32_t p0, bool parked)
{
    const uint32_t lane = lane_id();
    uint32_t ip = p0, op = J.op;
    bool done = false;
    const int e = special_run(J, ip, op, done, !parked);
    // a finished block -- good or bad -- leaves its ring as it is: the next block's first command flushes it
    if (e) { finish_block(J, e); return; }
    if (done) { finish_block(J, (op == J.origin && ip == J.comp_len) ? E_OK : E_SIZE); return; }
    if (parked || ip >= J.whi) refill_landed(S, s, J, true);
    if (lane == 0) {
        J.op = op;
        J.ip = ip;
        J.tail += 1;
        st_rlx(&S.tail[s], J.tail);
        // the next token lies beyond the window (long literal run): re-base the window.  The walker cannot be reading
        // there: parked, or waiting for exactly this position to come into the window.
        if (ip >= J.whi) { J.whi = ip & ~15u; st_rlx(&S.whi[s], J.whi); }
    }
    __syncwarp();
    if (parked) {
        while (refill(S, s, J, 16)) { }
        post_cmd(S, s, J, ip, J.comp_len);
    } else {
        refill_issue(S, s, J, 16);
    }
}

// Claim job `job` for slot s.  Trivial and oversized blocks are finished on the spot.
__device__ __noinline__ void start_job(V2Smem &S, int s, SlotJob &J, const DecodeArgs &args, unsigned int job)
{
    const uint32_t lane = lane_id();
    refill_landed(S, s, J, true);                              // a refill of the previous block may still be landing
    const int c = (int)(job % (unsigned int)args.ncols);
    const int b = args.blk0 + (int)(job / (unsigned int)args.ncols);
    const DecodeCol &col = args.col[c];
    if (col.skip && col.skip[b]) return;                       // body is referenced in place (stored block)
    const uint8_t *src = col.comp + col.comp_off[b];
    uint8_t *dst = col.out + col.dec_off[b];
    const int64_t comp_len = col.comp_len[b], origin = col.origin[b];
    if (comp_len <= 0) {
        if (lane == 0) col.status[b] = E_TRUNCATED;
        return;
    }
    if (comp_len >= (int64_t)POS_CAP || origin >= (int64_t)POS_CAP) {
        const int e = decode_simple(src, comp_len, dst, origin);
        if (lane == 0) col.status[b] = e;
        return;
    }
    if (lane == 0) {
        J.src = src; J.dst = dst; J.status = &col.status[b];
        J.comp_len = (uint32_t)comp_len; J.origin = (uint32_t)origin;
        J.op = 0; J.ip = 0; J.whi = 0; J.err = 0; J.pend = 0; J.chainy = 0;
        J.state = SLOT_ACTIVE;
        st_rlx(&S.whi[s], 0u);
    }
    __syncwarp();
    while (refill(S, s, J, 16)) { }
    post_cmd(S, s, J, 0u, (uint32_t)comp_len);
    // the command flushes the ring (a block that ended in an error is dropped where it stood, its entries with it):
    // the walker answers with the ring position the new block's entries start at
    while (ld_acq(&S.cmd_ack[s]) != J.seq) __nanosleep(100);
    const uint32_t hd = ld_rlx(&S.cmd_head[s]);
    if (lane == 0) { J.tail = hd; st_rlx(&S.tail[s], hd); }
    __syncwarp();
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v)
{
    const uint32_t lane = lane_id();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u = __shfl_up_sync(FULL, v, d);
        if ((int)lane >= d) v += u;
    }
    return v;
}

// One batch of an active slot whose ring looks ready -- any kind of entries.  Returns 0: nothing to do yet, 1: progress,
// 3: progress, and the batch was a generic one (more of the same is likely to follow).
__device__ __noinline__ int process_slot(V2Smem &S, int s, SlotJob &J, uint8_t *stg, bool force)
{
    const uint32_t lane = lane_id();
    const long long tp0 = STATS_ON ? clock64() : 0;
    const uint32_t h = ld_acq(&S.hint[s]);                     // ring entries below the count are visible
    const uint32_t avail = (h - J.tail) & H_CNT;
    const int navail = avail < 32u ? (int)avail : 32;
    if (navail == 0) return 0;
    const uint32_t p = (int)lane < navail ? S.ring[s][(J.tail + lane) & RM] : 0u;
    const uint32_t tok = (int)lane < navail ? S.win[s][p & WM] : 0u;   // the walker only emits tokens that are in the window
    // Every lane parses its sequence header.  A length extension of one byte (lengths up to 269 / 273) is read here --
    // the walker got past such a token because its extension bytes are in the window -- and the sequence stays in the
    // batch; longer extensions, and the last entry of a parked walker whatever its nibbles say, are "big": done one at
    // a time by special_step.
    const bool parked_last = (h & H_PARKED) && avail <= 32u;
    const bool ext = tok >= 0xf0u || (tok & 15u) == 15u;
    uint32_t L = tok >> 4, M = (tok & 15u) + 4, lit_pos = p + 1;
    bool big = parked_last && (int)lane == navail - 1;
    if ((int)lane < navail && L == 15u) {
        const uint32_t e = S.win[s][lit_pos & WM];
        lit_pos++;
        L += e;
        big |= e == 255u;
    }
    const uint32_t off_pos = lit_pos + L;
    uint32_t nxt = off_pos + 2;
    if ((int)lane < navail && !big && (tok & 15u) == 15u) {
        const uint32_t e = S.win[s][nxt & WM];
        nxt++;
        M += e;
        big |= e == 255u;
    }
    const uint32_t len = L + M;
    const uint32_t sm = __ballot_sync(FULL, (int)lane < navail && big);
    int nv = sm ? __ffs(sm) - 1 : navail;                      // leading entries a batch can take
    // wait for a full batch unless the run ends in a big entry, or the walker cannot go on before this warp has made
    // room in the window
    if (nv > 0 && nv < 32 && sm == 0 && !force) return 0;
    const uint32_t o0 = J.op;
    const uint32_t o = o0 + warp_incl_scan((int)lane < nv ? len : 0u) - ((int)lane < nv ? len : 0u);
    {
        // ... and only as many as fit the staging area and the window (sequences with long literal runs are big in bytes)
        const bool fits = (int)lane < nv && o + len - (o0 & ~15u) <= (uint32_t)(STG - 48) && nxt - (J.ip & ~15u) <= (uint32_t)(W - 32);
        const uint32_t fm = __ballot_sync(FULL, fits);
        const int nfit = fm == FULL ? 32 : __ffs(~fm) - 1;
        if (nfit < nv) nv = nfit;
    }
    if (nv == 0) {
        const long long t0 = STATS_ON ? clock64() : 0;
        special_step(S, s, J, __shfl_sync(FULL, p, 0), parked_last && navail == 1);
        if (STATS_ON && lane == 0) { STAT_ADD(ST_C_SPECIAL, 1); STAT_ADD(ST_C_SPECIAL_CYCLES, clock64() - t0); }
        return 3;
    }
    int err = E_OK;
    // the stream bytes of the whole batch must be in the window
    const uint32_t need = __shfl_sync(FULL, nxt, nv - 1);
    // (a refill that was in flight may land inside refill() and fill the window: no new one can be issued then, and none is needed)
    while (!err && J.whi < need)
        if (!refill(S, s, J, 16) && J.whi < need) err = E_INTERNAL;
    int nproc = 0;
    bool generic = false;
    uint32_t new_op = o0;
    long long tp4 = 0;
    if (!err) {
        const uint8_t *win = S.win[s];
        // 4 stream bytes behind the token: up to 2 literals and the offset of a word-regular sequence
        const uint32_t a = p + 1, a4 = a & ~3u;
        const uint32_t *w32 = reinterpret_cast<const uint32_t *>(win);
        const uint32_t lo = __funnelshift_r(w32[(a4 & WM) >> 2], w32[((a4 + 4) & WM) >> 2], (a & 3u) * 8);
        const uint32_t offv = (lo >> (8 * (L & 3u))) & 0xffffu;
        const uint32_t lit = lo & ~(0xffffffffu << (8 * (L & 3u)));
        const bool wr = (int)lane < nv && !ext && L <= 2 && ((len | offv | o) & 7u) == 0 && (offv - 1u) < o;
        const uint32_t rm = __ballot_sync(FULL, wr && ((o - o0) >> 3) + (len >> 3) <= 32u);
        const int nreg = rm == FULL ? 32 : (__ffs(~rm) - 1);
        if (nreg >= REG_MIN || (nreg == nv)) {
            nproc = nreg;
            err = regular_batch(J, o0, o, len, L, offv, lit, nreg, &new_op);
        } else {
            // up to where a long word-regular run starts (those lanes are better served by the regular path)
            const uint32_t rall = __ballot_sync(FULL, wr);
            const uint32_t run = rall & (rall >> 1) & (rall >> 2) & (rall >> 3);
            const uint32_t run16 = run & (run >> 4) & (run >> 8) & (run >> 12);      // bit i: lanes i..i+15 all regular
            const uint32_t cand = run16 & ~1u & (nv >= 32 ? FULL : ((1u << nv) - 1u));
            nproc = cand ? (__ffs(cand) - 1) : nv;
            generic = true;

            const long long tp1 = STATS_ON ? clock64() : 0;
            // the staging area starts at the 16-byte chunk that holds o0; its bytes below o0 are still there when the
            // warp's last generic batch was this slot's previous one, else they come back from memory
            const int cw = (int)(threadIdx.x >> 5) - NWALK;
            if (lane == 0 && (o0 & 15u) && !(S.stg_slot[cw] == (uint32_t)s && S.stg_op[cw] == o0))
                *reinterpret_cast<uint4 *>(stg) = __ldcg(reinterpret_cast<const uint4 *>(J.dst + (o0 & ~15u)));
            __syncwarp();
            const long long tp2 = STATS_ON ? clock64() : 0;
            err = generic_batch(win, stg, J, o0, lit_pos, off_pos, o, L, M, nproc, &new_op);
            const long long tp3 = STATS_ON ? clock64() : 0;
            if (!err) {
                flush(J, stg, o0 & ~15u, new_op);
                // keep the chunk the next batch starts in
                uint4 keepv = make_uint4(0u, 0u, 0u, 0u);
                if (lane == 0) keepv = *reinterpret_cast<const uint4 *>(stg + ((new_op & ~15u) - (o0 & ~15u)));
                __syncwarp();
                if (lane == 0) { *reinterpret_cast<uint4 *>(stg) = keepv; S.stg_slot[cw] = (uint32_t)s; S.stg_op[cw] = new_op; }
                __syncwarp();
            }
            if (STATS_ON) {
                tp4 = clock64();
                if (lane == 0) { WSTAT_ADD(0, tp1 - tp0); WSTAT_ADD(1, tp2 - tp1); WSTAT_ADD(2, tp3 - tp2); WSTAT_ADD(3, tp4 - tp3); WSTAT_ADD(6, 1); WSTAT_ADD(7, nproc); }
            }
        }
    }
    if (err) { finish_block(J, err); return 1; }   // dropped where it stands; the next block's command flushes the ring
    const uint32_t next_ip = __shfl_sync(FULL, nxt, nproc - 1);
    if (lane == 0) { J.op = new_op; J.ip = next_ip; J.tail += (uint32_t)nproc; st_rlx(&S.tail[s], J.tail); }
    __syncwarp();
    refill_issue(S, s, J, 256);
    if (STATS_ON && tp4 && lane == 0) WSTAT_ADD(4, clock64() - tp4);
    return generic ? 3 : 1;
}

// The slow path takes one batch per call.  After a generic batch or a special sequence, keep going while full batches
// are waiting: more of the same is likely to follow, and a wake-up otherwise costs a sleep period per batch.  (After a
// word-regular batch the burst path takes over again.)
__device__ __noinline__ bool slow_path(V2Smem &S, int s, SlotJob &J, uint8_t *stg, bool force)
{
    bool r = false;
    for (int n = 0; n < 4; n++) {
        const int rr = process_slot(S, s, J, stg, force);
        r |= rr != 0;
        if (rr != 3 || J.state != SLOT_ACTIVE || ((ld_rlx(&S.hint[s]) - J.tail) & H_CNT) < 32u) break;
    }
    return r;
}

// ---- the hot path: a burst of full word-regular batches from one slot, slot state in registers -------------------
// Anything else (partial batches, special entries, sequences that are not word-regular, window shortfalls, errors)
// is handed to process_slot, which works on the slot state in shared memory.
//
// The burst is software-pipelined over two batches.  Stage 1 of batch k+1 (ring entries -> stream bytes ->
// one output word per lane -> classification of every word's source) runs before stage 2 of batch k (resolve
// and store).  A source word is either inside the batch (warp shuffles, dependency waves), inside the previous
// batch (one shuffle from the registers that still hold it), or older -- then it is already in global memory
// and its load is issued in stage 1, a whole batch ahead of its use.
struct Staged {
    uint32_t meta;       // per word lane: offw | Lw << 16
    uint32_t litw;
    uint32_t far_lo, far_hi;
    uint32_t tw;         // uniform: words in the batch, 0 = nothing staged
};
enum { STG_OK = 0, STG_WAIT = 1, STG_SLOW = 2 };   // stage1 verdicts: staged / ring not full yet / needs process_slot

__device__ __forceinline__ int stage1(V2Smem &S, int s, const uint32_t *ring, const uint32_t *w32, const uint8_t *dst, uint32_t origin,
                                      uint32_t &tail, uint32_t op, uint32_t pt, uint32_t whi, uint32_t &ip, Staged &r)
{
    const uint32_t lane = lane_id();
    r.meta = 0; r.litw = 0; r.far_lo = 0; r.far_hi = 0; r.tw = 0;
    const uint32_t h = ld_acq(&S.hint[s]);                              // ring entries below the count are visible
    const uint32_t avail = (h - tail) & H_CNT;
    if (avail < 32u) return (h >> 31) && avail ? STG_SLOW : STG_WAIT;   // partial batch: wait, unless it ends in a special entry
    if (avail == 32u && (h >> 31)) return STG_SLOW;                     // the last entry of a parked walker is special
    const uint32_t p = ring[(tail + lane) & RM];
    // 8 stream bytes from the token on: token, up to 2 literals and the offset of a word-regular sequence
    const uint32_t a4 = p & ~3u, sh = (p & 3u) * 8;
    const uint32_t w0 = w32[(a4 & WM) >> 2], w1 = w32[((a4 + 4) & WM) >> 2];
    const uint32_t b0 = __funnelshift_r(w0, w1, sh);                    // bytes p .. p+3
    const uint32_t lo = __funnelshift_r(b0, w1 >> sh, 8);               // bytes p+1 .. p+4
    const uint32_t L = (b0 >> 4) & 15u;
    const uint32_t len = L + (b0 & 15u) + 4;
    const uint32_t offv = (lo >> (8 * (L & 3u))) & 0xffffu;
    const uint32_t lit = lo & ~(0xffffffffu << (8 * (L & 3u)));
    // output position of my sequence: every length is 8 in most batches, else a warp scan
    uint32_t o = op + 8 * lane;
    if (!__all_sync(FULL, len == 8u)) o = op + warp_incl_scan(len) - len;
    const bool wr = L <= 2 && ((len | offv | o) & 7u) == 0 && (offv - 1u) < o;
    const uint32_t fw = (o - op) >> 3;                             // first output word of my sequence
    const uint32_t rm = __ballot_sync(FULL, wr && fw + (len >> 3) <= 32u);
    const int nreg = rm == FULL ? 32 : (__ffs(~rm) - 1);
    if (nreg < REG_MIN) return STG_SLOW;
    const uint32_t need = __shfl_sync(FULL, p + 3 + L, nreg - 1);  // stream bytes of the batch must be in the window
    const uint32_t tw = __shfl_sync(FULL, fw + (len >> 3), nreg - 1);   // words in the batch (<= 32)
    if (need > whi || op + 8 * tw > origin) return STG_SLOW;
    uint32_t offw = offv, Lw = L, litw = lit;
    if ((int)tw != nreg) {                                         // some sequence spans two words: expand sequences to words
        const uint32_t startmask = __reduce_or_sync(FULL, (int)lane < nreg ? (1u << fw) : 0u);
        int sq = __popc(startmask & (0xffffffffu >> (31 - lane))) - 1;
        if (lane >= tw) sq = 0;
        const uint32_t pk = __shfl_sync(FULL, offv | (L << 16) | (fw << 24), sq);
        const uint32_t lit_s = __shfl_sync(FULL, lit, sq);
        const bool k0 = lane == (pk >> 24);                        // first word of its sequence carries the literals
        offw = pk & 0xffffu;
        Lw = k0 ? ((pk >> 16) & 0xffu) : 0u;
        litw = k0 ? lit_s : 0u;
    }
    r.meta = offw | (Lw << 16);
    r.litw = litw;
    // source older than the previous batch: already in global memory, load it now
    if (lane < tw && (int)lane - (int)(offw >> 3) < -(int)pt) {
        const uint2 far = __ldcg(reinterpret_cast<const uint2 *>(dst + (op + 8 * lane - offw)));
        r.far_lo = far.x;
        r.far_hi = far.y;
    }
    r.tw = tw;
    // the entries and stream bytes of the batch are in registers now: release them to the walker / the refill
    tail += (uint32_t)nreg;
    ip = need;
    if (lane == 0) st_rlx(&S.tail[s], tail);
    return STG_OK;
}

// Window upkeep of a burst, out of line: publish a landed refill, start the next one when 256 bytes are free.
__device__ __noinline__ void window_upkeep(V2Smem &S, int s, SlotJob &J, uint32_t ip)
{
    if (lane_id() == 0) J.ip = ip;
    __syncwarp();
    if (J.pend && !refill_landed(S, s, J, false)) return;
    refill_issue(S, s, J, 256);
}

// CHAINS: the slot's batches are dominated by in-batch chains (sorted / sequential columns: every word copies its
// predecessor); they are collapsed by pointer jumping instead of one dependency wave per word.  The plain flavour
// notices such batches and sets J.chainy, the other one clears it again; the consumer picks the flavour per call.
template <bool CHAINS>
__device__ __noinline__ int burst(V2Smem &S, int s, SlotJob &J)
{
    const uint32_t lane = lane_id();
    uint32_t tail = J.tail, op = J.op, ip = J.ip, whi = J.whi, pend = J.pend;
    uint8_t *const dst = J.dst;
    const uint32_t origin = J.origin;
    const uint32_t whi_max = (J.comp_len + 16) & ~15u;
    const uint32_t *ring = S.ring[s];
    const uint32_t *w32 = reinterpret_cast<const uint32_t *>(S.win[s]);
    bool progress = false;
    unsigned st_batches = 0;
    if (pend && refill_landed(S, s, J, false)) { whi = J.whi; pend = 0; progress = true; }
    uint32_t pt = 0;                       // words of the previous batch still held in prev_lo / prev_hi
    uint32_t prev_lo = 0, prev_hi = 0;
    Staged cur;
    cur.meta = 0; cur.litw = 0; cur.far_lo = 0; cur.far_hi = 0; cur.tw = 0;
    int verdict = STG_WAIT;
    for (int it = 0;; it++) {
        // window upkeep (out of line): a refill is in flight, or 256 bytes of the window are free
        if (it && (pend || (whi < whi_max && (ip & ~15u) + (uint32_t)W - whi >= 256u))) {
            window_upkeep(S, s, J, ip);
            whi = J.whi;
            pend = J.pend;
        }
        // stage 1 of the next batch (the only call site: the loop's first pass has no current batch yet)
        Staged nxt;
        nxt.tw = 0;
        verdict = STG_WAIT;
        if (it < MAX_BURST) verdict = stage1(S, s, ring, w32, dst, origin, tail, op + 8 * cur.tw, cur.tw, whi, ip, nxt);
        if (!cur.tw) {
            cur = nxt;
            if (!cur.tw) break;
            continue;
        }
        // stage 2 of the current batch: resolve every word and store
        const bool wl = lane < cur.tw;
        const uint32_t offw = cur.meta & 0xffffu, Lw = cur.meta >> 16;
        const uint32_t keep = 0xffffffffu << (8 * Lw);            // Lw <= 2
        const int dep = (int)lane - (int)(offw >> 3);             // >= 0: in this batch, >= -pt: previous batch, else memory
        const int pidx = dep + (int)pt;
        const uint32_t qlo = __shfl_sync(FULL, prev_lo, pidx & 31), qhi = __shfl_sync(FULL, prev_hi, pidx & 31);
        bool fin = !wl;
        uint32_t vlo = 0, vhi = 0;
        if (wl && dep < 0) {
            const bool inprev = pidx >= 0;
            vlo = cur.litw | ((inprev ? qlo : cur.far_lo) & keep);
            vhi = inprev ? qhi : cur.far_hi;
            fin = true;
        }
        // in-batch sources
        const uint32_t open = __ballot_sync(FULL, !fin);
        if (open) {
            if (CHAINS) {
                int d = dep;
                uint32_t lt = cur.litw, kp = keep;
                do {   // pointer jumping: (lit, keep) o (lit', keep') = (lit | lit' & keep, keep & keep')
                    const uint32_t finmask = __ballot_sync(FULL, fin);
                    const int j = fin ? (int)lane : d;
                    const uint32_t va = __shfl_sync(FULL, vlo, j), vb = __shfl_sync(FULL, vhi, j);
                    const uint32_t pl = __shfl_sync(FULL, lt, j), pk = __shfl_sync(FULL, kp, j);
                    const int pd = __shfl_sync(FULL, d, j);
                    if (!fin) {
                        if ((finmask >> j) & 1u) { vlo = lt | (va & kp); vhi = vb; fin = true; }
                        else { lt |= pl & kp; kp &= pk; d = pd; }
                    }
                } while (__any_sync(FULL, !fin));
                if (__popc(open) < 8 && lane == 0) J.chainy = 0;
            } else {
                // dependency waves over warp shuffles (the lowest unresolved word always has a resolved producer)
                do {
                    const uint32_t finmask = __ballot_sync(FULL, fin);
                    const int j = fin ? (int)lane : dep;
                    const uint32_t va = __shfl_sync(FULL, vlo, j), vb = __shfl_sync(FULL, vhi, j);
                    if (!fin && ((finmask >> j) & 1u)) { vlo = cur.litw | (va & keep); vhi = vb; fin = true; }
                } while (__any_sync(FULL, !fin));
                if (__popc(open) > 20 && lane == 0) J.chainy = 1;
            }
        }
        if (wl) *reinterpret_cast<uint2 *>(dst + op + 8 * lane) = make_uint2(vlo, vhi);
        __syncwarp();                                             // the stores are ordered before the far loads of later batches
        prev_lo = vlo; prev_hi = vhi; pt = cur.tw;
        op += 8 * cur.tw;
        progress = true;
        st_batches++;
        cur = nxt;
        if (!cur.tw) break;
    }
    if (STATS_ON && lane == 0) STAT_ADD(ST_C_REG, st_batches);
    if (lane == 0) { J.tail = tail; J.op = op; J.ip = ip; }
    __syncwarp();
    return (progress ? 1 : 0) | (verdict == STG_SLOW ? 2 : 0);
}

// Slot upkeep that is not on the hot path: claim the next block for an empty slot, or retire the slot.
__device__ __noinline__ void slot_refresh(V2Smem &S, int s, const DecodeArgs &args, unsigned int *counter, unsigned int first_dynamic,
                                          unsigned int static_job, bool use_static, int &live)
{
    const uint32_t lane = lane_id();
    SlotJob &J = S.job[s];
    const unsigned int njobs = (unsigned int)args.ncols * (unsigned int)args.nblocks;
    unsigned int job = static_job;
    if (!use_static) {
        job = 0;
        if (lane == 0) job = first_dynamic + atomicAdd(counter, 1u);
        job = __shfl_sync(FULL, job, 0);
    }
    if (job >= njobs) {
        if (lane == 0) J.state = SLOT_RETIRED;
        __syncwarp();
        post_cmd(S, s, J, 0u, LIM_EXIT);
        live--;
    } else {
        const long long t0 = STATS_ON ? clock64() : 0;
        start_job(S, s, J, args, job);
        if (STATS_ON && lane == 0) STAT_ADD(ST_C_START_CYCLES, clock64() - t0);
    }
}

__device__ __noinline__ void give_up(V2Smem &S, int c)
{
    for (int k = 0; k < SPC; k++) {
        const int s = c + k * NCONS;
        SlotJob &J = S.job[s];
        if (J.state == SLOT_RETIRED) continue;
        if (J.state == SLOT_ACTIVE && lane_id() == 0) *J.status = E_INTERNAL;
        if (lane_id() == 0) J.state = SLOT_RETIRED;
        __syncwarp();
        post_cmd(S, s, J, 0u, LIM_EXIT);
    }
}

// Consumer warp: owns SPC slots.  The polling loop is kept to a handful of instructions and sleeps when nothing is
// ready -- every issue slot a waiting consumer burns is taken from the walker warps, which pace the whole kernel.
__device__ void consumer(V2Smem &S, const DecodeArgs &args, unsigned int *counter, unsigned int first_dynamic, int c)
{
    const uint32_t lane = lane_id();
    uint8_t *stg = S.stg[c];
    int live = SPC;
    uint32_t first = (1u << SPC) - 1u;       // slots whose first job is the statically assigned one
    long long last_progress = clock64();
    const long long st_t0 = last_progress;
    unsigned int st_polls = 0, st_sleeps = 0;
    long long st_sleep_cycles = 0;
    while (live > 0) {
        bool progress = false, worked = false;   // worked: entries were consumed (anything less does not justify another poll right away)
        st_polls++;
#pragma unroll 1
        for (int k = 0; k < SPC; k++) {
            const int s = c + k * NCONS;
            SlotJob &J = S.job[s];
            const uint32_t st = J.state;
            if (st == SLOT_ACTIVE) {
                const uint32_t h = ld_rlx(&S.hint[s]);
                const uint32_t avail = (h - J.tail) & H_CNT;
                // Wake up for several batches at once (bursts amortise the slot's state and keep far loads a batch
                // ahead) -- but do not wait for entries the walker cannot produce: when its last token sits near the
                // end of the window it is about to stall until this warp has consumed a batch and refilled.
                bool ready = avail >= (uint32_t)READY_MIN || ((h >> 31) && avail >= 1u);
                if (!ready && avail >= 32u) ready = J.whi - S.ring[s][((h & H_CNT) - 1u) & RM] < 128u;
                // the walker needs stream bytes that only fit into the window once entries have been consumed
                const bool force = (h & H_STARVED) && avail >= 1u && J.pend == 0 && J.whi - (J.ip & ~15u) > (uint32_t)(W - 256);
                if (ready || force) {
                    const long long t0 = STATS_ON ? clock64() : 0;
                    const int code = avail < 32u ? 2 : (J.chainy ? burst<true>(S, s, J) : burst<false>(S, s, J));
                    bool r = (code & 1) != 0;
                    const long long t1 = STATS_ON ? clock64() : 0;
                    if (code & 2) r |= slow_path(S, s, J, stg, force);
                    if (STATS_ON && lane == 0) {
                        STAT_ADD(ST_C_PS_CALLS, 1); STAT_ADD(ST_C_PS_CYCLES, t1 - t0);
                        if (code & 2) { STAT_ADD(ST_C_SLOW_CALLS, 1); STAT_ADD(ST_C_SLOW_CYCLES, clock64() - t1); }
                    }
                    progress |= r;
                    worked |= r;
                } else if (J.pend) {
                    // the window is topped up whenever 256 bytes are free: after every batch and whenever a refill lands
                    if (refill_landed(S, s, J, false)) { refill_issue(S, s, J, 256); progress = true; }
                }
            } else if (st == SLOT_EMPTY) {
                // first pass of jobs: spread over CTAs, then warps, then slot levels
                slot_refresh(S, s, args, counter, first_dynamic, ((unsigned int)k * NCONS + (unsigned int)c) * gridDim.x + blockIdx.x,
                             (first >> k) & 1u, live);
                first &= ~(1u << k);
                progress = true;
                worked = true;
            }
        }
        if (progress) last_progress = clock64();
#ifdef DFDB_LZ4_STATS
        if (STATS_ON && st_polls == 3000000u) {   // diagnostics build: a warp that polls this often is stuck -- dump its slots and give up
            for (int k = 0; k < SPC; k++) {
                const int s = c + k * NCONS;
                SlotJob &J = S.job[s];
                if (lane == 0)
                    printf("[lz4 v3 stuck] cta %d cons %d slot %d state %u tail %u hint %08x op %u ip %u whi %u pend %u err %u origin %u comp %u ring[t] %u ring[t+1] %u\n",
                           (int)blockIdx.x, c, s, J.state, J.tail, ld_rlx(&S.hint[s]), J.op, J.ip, J.whi, J.pend, J.err, J.origin, J.comp_len,
                           S.ring[s][J.tail & RM], S.ring[s][(J.tail + 1) & RM]);
                if (J.state == SLOT_RETIRED) continue;
                if (J.state == SLOT_ACTIVE && lane == 0) *J.status = E_INTERNAL;
                if (lane == 0) J.state = SLOT_RETIRED;
                __syncwarp();
                post_cmd(S, s, J, 0u, LIM_EXIT);
            }
            live = 0;
        }
#endif
        if (!worked) {
            const long long ts0 = STATS_ON ? clock64() : 0;
            __nanosleep(IDLE_SLEEP_NS);
            st_sleeps++;
            if (STATS_ON) st_sleep_cycles += clock64() - ts0;
            if (!progress && clock64() - last_progress > 4000000000ll) {
                // watchdog (~2 s without progress): give up on the active slots instead of hanging the device
                give_up(S, c);
                live = 0;
            }
        }
    }
#ifdef DFDB_LZ4_STATS
    if (STATS_ON && lane == 0) {
        unsigned long long *w = g_warp_stats[blockIdx.x * 32 + (threadIdx.x >> 5)];
        STAT_ADD(ST_G_PRO, w[0]); STAT_ADD(ST_G_HEAD, w[1]); STAT_ADD(ST_G_BATCH, w[2]); STAT_ADD(ST_G_FLUSH, w[3]); STAT_ADD(ST_G_EPI, w[4]);
        STAT_ADD(ST_G_WAVES, w[5]); STAT_ADD(ST_C_GEN, w[6]); STAT_ADD(ST_C_GENSEQ, w[7]);
        for (int k = 0; k < 8; k++) w[k] = 0;
    }
#endif
    if (STATS_ON && lane == 0) { STAT_ADD(ST_C_POLLS, st_polls); STAT_ADD(ST_C_SLEEPS, st_sleeps); STAT_ADD(ST_C_LOOP_CYCLES, clock64() - st_t0); STAT_ADD(ST_C_SLEEP_CYCLES, st_sleep_cycles); }
}

}  // namespace

__global__ void __launch_bounds__(V2_THREADS, 1) lz4_decode_v3_kernel(const __grid_constant__ DecodeArgs args, unsigned int *counter,
                                                                           unsigned int first_dynamic)
{
    extern __shared__ __align__(16) uint8_t v3_smem_raw[];
    V2Smem &S = *reinterpret_cast<V2Smem *>(v3_smem_raw + ((1024u - (smem_addr(v3_smem_raw) & 1023u)) & 1023u));
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < NSLOT_PAD; i += blockDim.x) {
        if (i < NCONS) { S.stg_slot[i] = 0xffffffffu; S.stg_op[i] = 0; }
        S.tail[i] = 0; S.whi[i] = 0; S.cmd_seq[i] = 0; S.cmd_p[i] = 0; S.cmd_lim[i] = 0; S.hint[i] = 0; S.cmd_ack[i] = 0; S.cmd_head[i] = 0;
    }
    for (int i = threadIdx.x; i < NSLOT; i += blockDim.x) {
        SlotJob &J = S.job[i];
        J.src = nullptr; J.dst = nullptr; J.status = nullptr;
        J.comp_len = 0; J.origin = 0; J.op = 0; J.ip = 0; J.tail = 0; J.whi = 0; J.state = SLOT_EMPTY; J.err = 0; J.seq = 0; J.pend = 0; J.rphase = 0; J.chainy = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" ::"r"(smem_addr(&S.rbar[i])) : "memory");
    }
    __syncthreads();
    if (warp < NWALK) walker(S, warp * 32 + (int)(threadIdx.x & 31));
    else consumer(S, args, counter, first_dynamic, warp - NWALK);
}

int launch_lz4_decode_v3(const DecodeArgs &args, unsigned int *d_counter, int sm_count, cudaStream_t stream, int cta_limit)
{
    static bool configured = false;
    static unsigned long long *d_stats = nullptr;
#ifdef DFDB_LZ4_STATS
    static const bool want_stats = getenv("DFDB_LZ4_STATS") != nullptr;
    if (want_stats && !d_stats) {
        cudaMalloc(&d_stats, ST_COUNT * 8);
        cudaMemcpyToSymbol(g_stats, &d_stats, sizeof d_stats);
    }
#endif
    if (d_stats) cudaMemsetAsync(d_stats, 0, ST_COUNT * 8, stream);
    if (!configured) {
        if (cudaFuncSetAttribute(lz4_decode_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(V2Smem) + 1024)) != cudaSuccess) return 1;
        configured = true;
    }
    const long long njobs = (long long)args.ncols * args.nblocks;
    if (njobs <= 0) return 0;
    // one persistent CTA per SM; with few jobs, one job per consumer warp before any warp takes a second
    long long ctas = njobs;
    if (ctas > sm_count) ctas = sm_count;
    if (cta_limit > 0 && ctas > cta_limit) ctas = cta_limit;   // leave the other SMs to a concurrent kernel (api.cu: decode / scan overlap)
    if (ctas < 1) ctas = 1;
    // dynamic job ids start after the statically assigned first pass
    const unsigned int first_dynamic = (unsigned int)(ctas * NCONS * SPC);
    cudaMemsetAsync(d_counter, 0, sizeof(unsigned int), stream);
    lz4_decode_v3_kernel<<<(unsigned int)ctas, V2_THREADS, sizeof(V2Smem) + 1024, stream>>>(args, d_counter, first_dynamic);
    if (d_stats) {
        unsigned long long h[ST_COUNT];
        cudaStreamSynchronize(stream);
        cudaMemcpy(h, d_stats, sizeof h, cudaMemcpyDeviceToHost);
        static const char *names[ST_COUNT] = {"w_rounds", "w_commits", "w_ringfull", "w_winempty", "w_parked", "w_sleeps", "c_polls", "c_sleeps",
                                              "c_ps_calls", "c_ps_false", "c_reg", "c_regseq", "c_gen", "c_genseq", "c_special", "c_ps_cycles",
                                              "c_loop_cycles", "c_refills", "w_cycles", "c_start_cycles", "c_slow_calls", "c_slow_cycles",
                                              "c_special_cycles", "c_sleep_cycles", "g_pro", "g_head", "g_batch", "g_flush", "g_epi", "g_waves"};
        fprintf(stderr, "[lz4 v3 stats] jobs=%lld ctas=%lld", njobs, ctas);
        for (int i = 0; i < ST_COUNT; i++) fprintf(stderr, " %s=%llu", names[i], h[i]);
        fprintf(stderr, "\n");
    }
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace dfdb
