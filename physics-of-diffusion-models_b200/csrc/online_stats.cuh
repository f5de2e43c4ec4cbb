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

}  // namespace pdm
