// lane_sim.cpp -- host harness for the lane-per-block LZ4 decoder's state machine (csrc/lz4_lane_core.cuh).
//
// Runs ONE lane of lz4_decode_lane.cu on the CPU with the kernel's own schedule (ROUND emit/parse steps through DEPTH piece descriptors, flush of complete
// 128-byte units, window refill of at most 8 chunks of 16 bytes that becomes usable one round later), so that the parse /
// emit logic, the ring and window arithmetic and the LZ4_decompress_safe acceptance rules can be checked against the
// oracle without a GPU (tests/test_lane_core.py).  Test infrastructure: not part of the product library.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../dataframedbs.jl_b200/csrc/lz4_lane_core.cuh"

using namespace dfdb::lane;

namespace {
struct HostMem {
    uint64_t win[WIN_BYTES / 8];
    uint64_t ring[RING_WORDS];
    uint64_t qd[SUBS * DEPTH];
    const uint8_t *out;
    uint32_t flushed;
    long far_loads = 0, far_bad = 0;
    uint64_t win_read(uint32_t pos) const { return win[(pos & (WIN_BYTES - 1)) >> 3]; }
    uint64_t ring_load(uint32_t s) const { return ring[s]; }
    void ring_store(uint32_t s, uint64_t v) { ring[s] = v; }
    void put_data(uint32_t slot, uint64_t d) { qd[slot] = d; }
    void put_far(uint32_t slot, uint32_t src) { qd[slot] = out_load(src); }   // (the kernel's asynchronous copy reads bytes that are final already)
    uint64_t get_data(uint32_t slot) const { return qd[slot]; }
    uint64_t out_load(uint32_t pos)
    {
        far_loads++;
        if (pos + 8 > flushed) far_bad++;     // must never read what is not final in global memory
        uint64_t v;
        memcpy(&v, out + pos, 8);
        return v;
    }
};
}  // namespace

// hot_period: 1 = every step is a full step (general columns); HOT_PERIOD = the schedule of word-regular columns
extern "C" __attribute__((visibility("default")))
int lane_sim_decode2(const uint8_t *comp, int comp_len, uint8_t *out, int origin, long *stats, int hot_period)
{
    if (origin == 0) return (comp_len == 1 && comp[0] == 0) ? 0 : E_SIZE;
    if (comp_len <= 0) return E_TRUNCATED;
    if ((uint32_t)origin >= MAX_POS || (uint32_t)comp_len >= MAX_POS) return E_INTERNAL;
    const uint32_t padded = ((uint32_t)comp_len + 15u) & ~15u;
    std::vector<uint8_t> src(padded + 16, 0xFF);          // hostile padding: nothing past comp_len may influence the result
    memcpy(src.data(), comp, (size_t)comp_len);
    std::vector<uint8_t> obuf((((size_t)origin + 255) & ~(size_t)255) + 256, 0xEE);
    HostMem mem;
    memset(mem.win, 0xAB, sizeof mem.win);
    memset(mem.ring, 0xCD, sizeof mem.ring);
    mem.out = obuf.data();
    mem.flushed = 0;
    Parser P;
    Emitter E;
    P.reset((uint32_t)comp_len, (uint32_t)origin);
    E.reset();
    uint32_t desc[SUBS * DEPTH];
    for (auto &d : desc) d = K_NONE;
    long fast_p = 0, fast_e = 0;
    uint32_t win_req = 0;
    long rounds = 0, pieces = 0, stalls = 0;
    int status = -1;
    for (;;) {
        rounds++;
        P.win_fill = win_req;                               // last round's refill has landed
        {   // this round's refill: whole 16-byte chunks, at most 8, never over bytes the parser still needs; usable next round
            const uint32_t consumed = P.ip & ~15u;
            uint32_t m = (WIN_BYTES - (win_req - consumed)) / 16;
            if (m > 8) m = 8;
            if (m > (padded - win_req) / 16) m = (padded - win_req) / 16;
            for (uint32_t c = 0; c < m; c++) memcpy(reinterpret_cast<uint8_t *>(mem.win) + ((win_req + 16 * c) & (WIN_BYTES - 1)), src.data() + win_req + 16 * c, 16);
            win_req += 16 * m;
        }
        for (int v = 0; v < ROUND; v++) {
            const uint32_t uA = (uint32_t)(SUBS * (v % DEPTH)), uB = uA + 1;
            if (hot_period > 1 && (v % hot_period) != hot_period - 1) {
                // hot step: two plain tokens or nothing
                E.fast2(mem, desc[uA], desc[uB], uA, uB);
                if (P.fast2(mem, E.flushed, uA, uB, desc[uA], desc[uB])) { fast_p += 2; pieces += 2; } else if (!P.finished()) stalls += 2;
                continue;
            }
            // full step, two pieces: sub-slots A and B of queue position v % DEPTH; the word fast path first, the general step otherwise
            for (int sub = 0; sub < SUBS; sub++) {
                const uint32_t u = (uint32_t)(SUBS * (v % DEPTH) + sub);
                if (E.fast(mem, desc[u], u)) fast_e++; else E.step(mem, desc[u], u);
            }
            for (int sub = 0; sub < SUBS; sub++) {
                const uint32_t u = (uint32_t)(SUBS * (v % DEPTH) + sub);
                uint32_t m = K_NONE;
                if (P.fast(mem, E.flushed, u, m)) fast_p++; else m = P.step(mem, E.flushed, u);
                desc[u] = m;
                if (piece_kind(m) != K_NONE) pieces++; else if (!P.finished()) stalls++;
            }
        }
        for (int k = 0; k < SUBS && E.unit_ready(); k++) {   // the kernel flushes at most SUBS units per lane and round
            const uint32_t s = E.flush_slot();
            memcpy(obuf.data() + E.flushed, &mem.ring[s], UNIT_BYTES);
            E.flushed += UNIT_BYTES;
            mem.flushed = E.flushed;
        }
        if (P.st == PS_ERR) { status = (int)P.err; break; }
        if (P.st == PS_END && E.op == P.opp) {
            while (E.flushed < E.op) {                       // the last, partial unit (the slot is padded)
                const uint32_t s = E.flush_slot();
                memcpy(obuf.data() + E.flushed, &mem.ring[s], UNIT_BYTES);
                E.flushed += UNIT_BYTES;
            }
            status = (E.op == (uint32_t)origin && P.ip == P.ip_end) ? E_OK : E_SIZE;
            break;
        }
        if (rounds > (long)origin + (long)comp_len + 64) { status = E_INTERNAL; break; }   // no progress: a bug, not a stream property
    }
    if (status == E_OK) memcpy(out, obuf.data(), (size_t)origin);
    if (mem.far_bad) status = 100;
    if (stats) { stats[0] = rounds; stats[1] = pieces; stats[2] = stalls; stats[3] = mem.far_loads; stats[4] = fast_p; stats[5] = fast_e; }
    return status;
}

extern "C" __attribute__((visibility("default")))
int lane_sim_decode(const uint8_t *comp, int comp_len, uint8_t *out, int origin, long *stats)
{
    return lane_sim_decode2(comp, comp_len, out, origin, stats, 1);
}
