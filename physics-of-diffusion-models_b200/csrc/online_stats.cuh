// Online (streaming) Boltzmann statistics of one query row.
//
// State (m, l, a1, a2, aux, idx) with m the running minimum energy and, for e_j = (E_j - m)/T,
//   l = sum exp(-e_j),  a1 = sum exp(-e_j) e_j,  a2 = sum exp(-e_j) e_j^2,  aux = sum exp(-e_j) s_j.
// When the minimum drops by delta*T every term is re-expressed against the new minimum with the
// non-negative shift rule of SURVEY.md section 5 -- no cancellation, same rule as the cross-shard merge:
//   c = exp(-delta);  a2 <- c (a2 + 2 delta a1 + delta^2 l);  a1 <- c (a1 + delta l);  l <- c l.
// The reference computes the same sums after materialising the whole row (utils/stats.py:80-90,
// 282-288); the min-shifted quantities it reports are invariant under the order of accumulation.
#pragma once

#include "pdm_common.cuh"

namespace pdm {

struct RowState {
    float m, l, a1, a2, aux;
    long long idx;   // global dataset index of the minimum (first index on ties)
};

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void state_init(RowState& s) {
    s.m = INFINITY; s.l = 0.f; s.a1 = 0.f; s.a2 = 0.f; s.aux = 0.f; s.idx = -1;
}

// Lower the running minimum to m_new (<= s.m).
__device__ __forceinline__ void state_lower_min(RowState& s, float m_new, float inv_t) {
    if (s.l > 0.f) {
        const float delta = fminf((s.m - m_new) * inv_t, 1.0e15f);
        const float c = fast_exp2(-delta * kLog2e);
        s.a2 = c * fmaf(delta, fmaf(delta, s.l, 2.f * s.a1), s.a2);
        s.a1 = c * fmaf(delta, s.l, s.a1);
        s.l *= c;
        s.aux *= c;
    }
    s.m = m_new;
}

// Add one dataset point with energy E >= s.m.
template <bool kAux>
__device__ __forceinline__ void state_add(RowState& s, float E, float inv_t, float aux_j) {
    const float e = fminf((E - s.m) * inv_t, kMaxE);
    const float w = fast_exp2(-e * kLog2e);
    const float we = w * e;
    s.l += w;
    s.a1 += we;
    s.a2 = fmaf(we, e, s.a2);
    if (kAux) s.aux = fmaf(w, aux_j, s.aux);
}

// Fold a chunk of NC energies (global column index col0 + i*col_step) into the state.
template <int NC, bool kAux>
__device__ __forceinline__ void state_add_chunk(RowState& s, const float (&E)[NC], const float (&aux)[NC],
                                                long long col0, int col_step, float inv_t) {
    float cmin = E[0];
#pragma unroll
    for (int i = 1; i < NC; ++i) cmin = fminf(cmin, E[i]);
    if (cmin >= 0.1f * kBigE && s.l <= 0.f) return;   // nothing but masked columns so far
    if (cmin < s.m) {
        int first = NC - 1;
#pragma unroll
        for (int i = NC - 1; i >= 0; --i) first = (E[i] == cmin) ? i : first;
        s.idx = col0 + (long long)first * col_step;
        state_lower_min(s, cmin, inv_t);
    }
#pragma unroll
    for (int i = 0; i < NC; ++i) state_add<kAux>(s, E[i], inv_t, kAux ? aux[i] : 0.f);
}

// Combine two states of the same row (disjoint dataset subsets).  Ties on m keep the lower index.
__device__ __forceinline__ void state_merge(RowState& a, const RowState& b_in, float inv_t) {
    RowState b = b_in;
    if (b.l <= 0.f) return;
    if (a.l <= 0.f) { a = b; return; }
    if (b.m < a.m || (b.m == a.m && b.idx < a.idx)) a.idx = b.idx;
    const float m = fminf(a.m, b.m);
    state_lower_min(a, m, inv_t);
    state_lower_min(b, m, inv_t);
    a.l += b.l; a.a1 += b.a1; a.a2 += b.a2; a.aux += b.aux;
}

__device__ __forceinline__ void state_store(const RowState& s, float* rec) {
    float4 v0 = make_float4(s.m, s.l, s.a1, s.a2);
    float4 v1 = make_float4(s.aux, __int_as_float((int)(s.idx & 0xffffffffLL)),
                            __int_as_float((int)(s.idx >> 32)), 0.f);
    reinterpret_cast<float4*>(rec)[0] = v0;
    reinterpret_cast<float4*>(rec)[1] = v1;
}

__device__ __forceinline__ void state_load(RowState& s, const float* rec) {
    const float4 v0 = reinterpret_cast<const float4*>(rec)[0];
    const float4 v1 = reinterpret_cast<const float4*>(rec)[1];
    s.m = v0.x; s.l = v0.y; s.a1 = v0.z; s.a2 = v0.w; s.aux = v1.x;
    const unsigned lo = (unsigned)__float_as_int(v1.y);
    const long long hi = (long long)__float_as_int(v1.z);
    s.idx = (hi << 32) | (long long)lo;
}

// ------------------------------------------------------------------------------------------------
// Packed variant for the tensor-core epilogue (sm_100 FADD2 / FMUL2 / FFMA2: two fp32 lanes per issue slot).
// Works on u = 2E (the reference's squared distance, utils/distance.py:21) in log2 units:
//   arg_j = (u_j - m2) * c2  with  m2 = 2m,  c2 = -log2(e) / (2T)   =>   arg_j = -e_j log2(e) <= 0,
//   w_j = 2^arg_j,   l = sum w,   b1 = sum w arg = -log2(e) A1,   b2 = sum w arg^2 = log2(e)^2 A2.
// Every accumulator carries two lanes (even / odd columns) that are added when the record is stored.
// Lowering the minimum by delta' = (m2_old - m2_new)(-c2) >= 0 shifts every arg by -delta':
//   b2 <- c (b2 - 2 delta' b1 + delta'^2 l);  b1 <- c (b1 - delta' l);  l <- c l;   c = 2^-delta'.
struct PackedState {
    float m2;
    float2 l, b1, b2, aux;
    long long idx;
};

__device__ __forceinline__ void packed_init(PackedState& s) {
    s.m2 = INFINITY; s.idx = -1;
    s.l = s.b1 = s.b2 = s.aux = make_float2(0.f, 0.f);
}

__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }

__device__ __forceinline__ void packed_lower_min(PackedState& s, float m2_new, float c2) {
    if (s.l.x > 0.f || s.l.y > 0.f) {
        const float dl = fminf((s.m2 - m2_new) * (-c2), 1.0e15f);
        const float2 c = splat2(fast_exp2(-dl)), d = splat2(dl), nd = splat2(-dl);
        const float2 t = __ffma2_rn(d, s.l, __fmul2_rn(splat2(-2.f), s.b1));        // delta' l - 2 b1
        s.b2 = __fmul2_rn(c, __ffma2_rn(d, t, s.b2));
        s.b1 = __fmul2_rn(c, __ffma2_rn(nd, s.l, s.b1));
        s.l = __fmul2_rn(c, s.l);
        s.aux = __fmul2_rn(c, s.aux);
    }
    s.m2 = m2_new;
}

// Fold 2*NP values u (pairs of consecutive columns, global index col0 + i) into the state.
// kClamp: guard the exponent against overflow (masked columns carry u = 2 kBigE; absurdly small T).
template <int NP, bool kAux, bool kClamp>
__device__ __forceinline__ void packed_add_chunk(PackedState& s, const float2 (&u)[NP], const float2 (&aux)[NP],
                                                 long long col0, float c2) {
    float cmin = fminf(u[0].x, u[0].y);
#pragma unroll
    for (int i = 1; i < NP; ++i) cmin = fminf(fminf(cmin, u[i].x), u[i].y);
    if (kClamp && cmin >= 0.1f * kBigE && s.l.x <= 0.f && s.l.y <= 0.f) return;   // only masked columns so far
    if (cmin < s.m2) {
        int first = 2 * NP - 1;
#pragma unroll
        for (int i = NP - 1; i >= 0; --i) {
            first = (u[i].y == cmin) ? 2 * i + 1 : first;
            first = (u[i].x == cmin) ? 2 * i : first;
        }
        s.idx = col0 + first;
        packed_lower_min(s, cmin, c2);
    }
    const float2 nm = splat2(-s.m2), cv = splat2(c2);
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        float2 arg = __fmul2_rn(__fadd2_rn(u[i], nm), cv);
        if (kClamp) { arg.x = fmaxf(arg.x, -kMaxE); arg.y = fmaxf(arg.y, -kMaxE); }
        const float2 w = make_float2(fast_exp2(arg.x), fast_exp2(arg.y));
        const float2 wa = __fmul2_rn(w, arg);
        s.l = __fadd2_rn(s.l, w);
        s.b1 = __fadd2_rn(s.b1, wa);
        s.b2 = __ffma2_rn(wa, arg, s.b2);
        if (kAux) s.aux = __ffma2_rn(w, aux[i], s.aux);
    }
}

// Record in the natural units of RowState (m, l, A1, A2, AUX, idx).
__device__ __forceinline__ void packed_store(const PackedState& s, float* rec) {
    constexpr float kLn2 = 0.6931471805599453f;
    RowState r;
    r.m = 0.5f * s.m2;
    r.l = s.l.x + s.l.y;
    r.a1 = -kLn2 * (s.b1.x + s.b1.y);
    r.a2 = (kLn2 * kLn2) * (s.b2.x + s.b2.y);
    r.aux = s.aux.x + s.aux.y;
    r.idx = s.idx;
    state_store(r, rec);
}

}  // namespace pdm
