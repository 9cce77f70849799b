// lz4_fused.cuh -- predicate + aggregate folded into a decoder (K3 + K7 inside K1), shared by lz4_decode_spec.cu and
// lz4_decode_lane.cu: apply(::SelectionExecutor) + the Base folds over iterate(::DFColumn)
// (/root/reference/src/tables/selection.jl:133-167, column.jl:102-126) on the words of the predicate column while they are
// still in registers, against an already resident 8-byte column.  The predicate is one closed interval (+ != constants)
// on a non-nullable 8-byte column (ColTest, built by build_tma_args in api.cu); sums are compensated (two-sum) and combined
// in a fixed order, so a given table gives the same bits on every run.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "kernels.cuh"

namespace dfdb {
namespace fused {

constexpr unsigned FULL = 0xffffffffu;

struct LaneAcc {
    double sum_hi, sum_lo, min_f, max_f;
    long long sum_i, min_i, max_i;
    int count, flags;     // flags: 1 = NaN seen, 2 = -0.0 seen, 4 = +0.0 seen
};

__device__ __forceinline__ void acc_reset(LaneAcc &a, bool uns)
{
    a.sum_hi = 0.0; a.sum_lo = 0.0; a.min_f = CUDART_INF; a.max_f = -CUDART_INF;
    a.sum_i = 0; a.min_i = uns ? -1ll : 0x7fffffffffffffffll; a.max_i = uns ? 0ll : (long long)0x8000000000000000ull;
    a.count = 0; a.flags = 0;
}
__device__ __forceinline__ void two_sum_add(double &hi, double &lo, double x)
{
    const double t = hi + x;
    const double bb = t - hi;
    lo += (hi - (t - bb)) + (x - bb);
    hi = t;
}
__device__ __forceinline__ bool lane_test(const LaneFused &F, unsigned long long x)
{
    if (F.test.cls == VC_FLT) {
        const double a = __longlong_as_double((long long)x);
        bool p = (a >= F.test.lo_f) && (a <= F.test.hi_f);
        if (F.test.n_ne) p = p && !(a == F.test.ne_f[0]) && (F.test.n_ne < 2 || !(a == F.test.ne_f[1]));
        if (F.test.nan_passes) p = p || (a != a);
        return p;
    }
    bool p = (x - (unsigned long long)F.test.lo_i) <= ((unsigned long long)F.test.hi_i - (unsigned long long)F.test.lo_i);
    if (F.test.n_ne) p = p && x != (unsigned long long)F.test.ne_i[0] && (F.test.n_ne < 2 || x != (unsigned long long)F.test.ne_i[1]);
    return p;
}
template <int AGG>
__device__ __forceinline__ void acc_add(LaneAcc &a, unsigned long long v, bool uns)
{
    a.count++;
    if (AGG == 2) {
        const double x = __longlong_as_double((long long)v);
        two_sum_add(a.sum_hi, a.sum_lo, x);
        if (x < a.min_f) a.min_f = x;
        if (x > a.max_f) a.max_f = x;
        if (x != x) a.flags |= 1;
        if (x == 0.0) a.flags |= signbit(x) ? 2 : 4;
    } else if (AGG == 1) {
        a.sum_i = (long long)((unsigned long long)a.sum_i + v);
        if (uns) {
            if (v < (unsigned long long)a.min_i) a.min_i = (long long)v;
            if (v > (unsigned long long)a.max_i) a.max_i = (long long)v;
        } else {
            if ((long long)v < a.min_i) a.min_i = (long long)v;
            if ((long long)v > a.max_i) a.max_i = (long long)v;
        }
    }
}
// fold the accumulator of lane `lane ^ d` (same group of 8) into this one; the lower lane keeps (lower, upper) order
template <int AGG>
__device__ __forceinline__ void acc_merge_xor(LaneAcc &a, int d, bool uns, bool upper)
{
    LaneAcc b;
    b.count = __shfl_xor_sync(FULL, a.count, d);
    b.flags = __shfl_xor_sync(FULL, a.flags, d);
    if (AGG == 2) {
        b.sum_hi = __shfl_xor_sync(FULL, a.sum_hi, d);
        b.sum_lo = __shfl_xor_sync(FULL, a.sum_lo, d);
        b.min_f = __shfl_xor_sync(FULL, a.min_f, d);
        b.max_f = __shfl_xor_sync(FULL, a.max_f, d);
        // (lower, upper) order on both sides, so the pair agrees bit for bit
        double hi = upper ? b.sum_hi : a.sum_hi, lo = upper ? b.sum_lo : a.sum_lo;
        const double xh = upper ? a.sum_hi : b.sum_hi, xl = upper ? a.sum_lo : b.sum_lo;
        two_sum_add(hi, lo, xh);
        lo += xl;
        a.sum_hi = hi; a.sum_lo = lo;
        if (b.min_f < a.min_f) a.min_f = b.min_f;
        if (b.max_f > a.max_f) a.max_f = b.max_f;
    } else if (AGG == 1) {
        b.sum_i = __shfl_xor_sync(FULL, a.sum_i, d);
        b.min_i = __shfl_xor_sync(FULL, a.min_i, d);
        b.max_i = __shfl_xor_sync(FULL, a.max_i, d);
        a.sum_i = (long long)((unsigned long long)a.sum_i + (unsigned long long)b.sum_i);
        if (uns) {
            if ((unsigned long long)b.min_i < (unsigned long long)a.min_i) a.min_i = b.min_i;
            if ((unsigned long long)b.max_i > (unsigned long long)a.max_i) a.max_i = b.max_i;
        } else {
            if (b.min_i < a.min_i) a.min_i = b.min_i;
            if (b.max_i > a.max_i) a.max_i = b.max_i;
        }
    }
    a.count += b.count;
    a.flags |= b.flags;
}
template <int AGG>
__device__ __forceinline__ AggPartial acc_to_partial(const LaneAcc &a)
{
    AggPartial p;
    p.count = a.count; p.nmissing = 0; p.sum_i = a.sum_i;
    p.sum_f = a.sum_hi; p.sum_lo = a.sum_lo;
    p.min_i = a.min_i; p.max_i = a.max_i;
    p.min_f = a.min_f; p.max_f = a.max_f;
    if (AGG == 2) {   // signed zeros: -0.0 orders before 0.0 in Julia's min / max
        if (p.min_f == 0.0) p.min_f = (a.flags & 2) ? -0.0 : 0.0;
        if (p.max_f == 0.0) p.max_f = (a.flags & 4) ? 0.0 : -0.0;
    }
    p.has_nan = a.flags & 1;
    p.has_value = (AGG != 0) && a.count > 0;
    return p;
}


}  // namespace fused
}  // namespace dfdb
