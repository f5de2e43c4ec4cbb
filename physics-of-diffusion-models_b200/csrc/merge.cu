// Merge of partial Boltzmann records: across the N splits of one launch and across dataset shards
// (GPUs) after an all-gather.  HBM-bound: reads n_src * 32 B per row, writes 8 floats + one int64.
// Finalises the quantities the reference reports (utils/stats.py:83-90, 284-289).
#include "pdm_common.cuh"
#include "online_stats.cuh"

namespace pdm {

// Records are stored record-major, so the threads of a warp read consecutive 32-byte records.  A thread's loads do
// not depend on its exp-heavy merge chain: kChunk records are fetched at a time and the next chunk is already in
// flight while the current one is merged (same merge order as a plain loop -> same bits).
constexpr int kChunk = 4;

struct RecordCursor {          // walks (outer, inner) without a 64-bit division per record
    const float* base;
    int64_t outer_stride, inner_stride, n_inner, i;
    __device__ __forceinline__ const float4* next() {
        const float4* p = reinterpret_cast<const float4*>(base + i * inner_stride);
        if (++i == n_inner) { i = 0; base += outer_stride; }
        return p;
    }
};

__device__ __forceinline__ void gather_row(const float* __restrict__ parts, int64_t row, int64_t n_outer, int64_t outer_stride,
                                           int64_t n_inner, int64_t row_stride, int64_t inner_stride, float it, RowState& acc) {
    state_init(acc);
    const int64_t total = n_outer * n_inner;
    RecordCursor cur{parts + row * row_stride, outer_stride, inner_stride, n_inner, 0};
    float4 nx0[kChunk], nx1[kChunk];
#pragma unroll
    for (int k = 0; k < kChunk; ++k)
        if (k < total) { const float4* p = cur.next(); nx0[k] = __ldg(p); nx1[k] = __ldg(p + 1); }
    for (int64_t t0 = 0; t0 < total; t0 += kChunk) {
        float4 c0[kChunk], c1[kChunk];
#pragma unroll
        for (int k = 0; k < kChunk; ++k) { c0[k] = nx0[k]; c1[k] = nx1[k]; }
#pragma unroll
        for (int k = 0; k < kChunk; ++k)
            if (t0 + kChunk + k < total) { const float4* p = cur.next(); nx0[k] = __ldg(p); nx1[k] = __ldg(p + 1); }
#pragma unroll
        for (int k = 0; k < kChunk; ++k) {
            if (t0 + k >= total) break;
            RowState s;
            s.m = c0[k].x; s.l = c0[k].y; s.a1 = c0[k].z; s.a2 = c0[k].w; s.aux = c1[k].x;
            s.idx = ((long long)__float_as_int(c1[k].z) << 32) | (long long)(unsigned)__float_as_int(c1[k].y);
            state_merge(acc, s, it);
        }
    }
}

// Combine records without finalising: one record per row (what a rank sends to its peers).
__global__ void __launch_bounds__(256, 3) reduce_partials_kernel(const float* __restrict__ parts, int64_t M, int64_t n_outer,
                                                              int64_t outer_stride, int64_t n_inner, int64_t row_stride,
                                                              int64_t inner_stride, const float* __restrict__ inv_temp,
                                                              float* __restrict__ out) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= M) return;
    RowState acc;
    gather_row(parts, row, n_outer, outer_stride, n_inner, row_stride, inner_stride, inv_temp[row], acc);
    state_store(acc, out + row * PDM_PART_STRIDE);
}

__global__ void __launch_bounds__(256, 3) merge_partials_kernel(const float* __restrict__ parts, int64_t M, int64_t n_outer,
                                                             int64_t outer_stride, int64_t n_inner, int64_t row_stride,
                                                             int64_t inner_stride, const float* __restrict__ inv_temp, float log_n,
                                                             float* __restrict__ out, int64_t* __restrict__ argmin) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= M) return;
    RowState acc;
    gather_row(parts, row, n_outer, outer_stride, n_inner, row_stride, inner_stride, inv_temp[row], acc);
    const float inv_l = acc.l > 0.f ? 1.f / acc.l : 0.f;
    const float log_l = logf(acc.l);
    const float mean_e = acc.a1 * inv_l;
    const float mean_e2 = acc.a2 * inv_l;
    out[PDM_OUT_E_MIN * M + row] = acc.m;
    out[PDM_OUT_LOG_L * M + row] = log_l;
    out[PDM_OUT_MEAN_E * M + row] = mean_e;
    out[PDM_OUT_MEAN_E2 * M + row] = mean_e2;
    out[PDM_OUT_VAR_E * M + row] = fmaxf(mean_e2 - mean_e * mean_e, 0.f);
    out[PDM_OUT_AUX_MEAN * M + row] = acc.aux * inv_l;
    out[PDM_OUT_ENTROPY * M + row] = log_l + mean_e - log_n;
    out[PDM_OUT_L * M + row] = acc.l;
    if (argmin) argmin[row] = acc.idx;
}

}  // namespace pdm

using namespace pdm;

extern "C" int pdm_merge_partials(const float* parts, int64_t M, int64_t n_outer, int64_t outer_stride,
                                  int64_t n_inner, int64_t inner_stride, int64_t row_stride, const float* inv_temp,
                                  int64_t n_total, float* out, int64_t* argmin, pdm_stream_t stream) {
    PDM_REQUIRE(parts && inv_temp && out && M >= 0 && n_outer >= 1 && n_inner >= 1 && n_total >= 1,
                "pdm_merge_partials: bad arguments");
    PDM_REQUIRE(row_stride >= PDM_PART_STRIDE && inner_stride >= PDM_PART_STRIDE && (row_stride % 4) == 0 &&
                (inner_stride % 4) == 0 && (outer_stride % 4) == 0,
                "pdm_merge_partials: strides must cover a record and keep 16-byte alignment");
    if (M == 0) return PDM_OK;
    merge_partials_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, as_stream(stream)>>>(
        parts, M, n_outer, outer_stride, n_inner, row_stride, inner_stride, inv_temp, logf((float)n_total), out, argmin);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}

extern "C" int pdm_reduce_partials(const float* parts, int64_t M, int64_t n_outer, int64_t outer_stride,
                                   int64_t n_inner, int64_t inner_stride, int64_t row_stride, const float* inv_temp,
                                   float* out_records, pdm_stream_t stream) {
    PDM_REQUIRE(parts && inv_temp && out_records && M >= 0 && n_outer >= 1 && n_inner >= 1,
                "pdm_reduce_partials: bad arguments");
    PDM_REQUIRE(row_stride >= PDM_PART_STRIDE && inner_stride >= PDM_PART_STRIDE && (row_stride % 4) == 0 &&
                (inner_stride % 4) == 0 && (outer_stride % 4) == 0,
                "pdm_reduce_partials: strides must cover a record and keep 16-byte alignment");
    if (M == 0) return PDM_OK;
    reduce_partials_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, as_stream(stream)>>>(
        parts, M, n_outer, outer_stride, n_inner, row_stride, inner_stride, inv_temp, out_records);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}

// ------------------------------------------------------------------------------------------------
// merge of the top-k epilogue's records: one warp per query row, k selection rounds over the records' candidates
// (records x 8 (value, local index) pairs, L2-resident): round r takes the lexicographic minimum of (value, index)
// strictly above the pair selected in round r-1 -- ascending values, ties by lower dataset index.
// ------------------------------------------------------------------------------------------------
namespace pdm {

__global__ void __launch_bounds__(256) topk_merge_kernel(const float* __restrict__ val, const int32_t* __restrict__ idx, int64_t M,
                                                         int64_t records, int64_t index_offset, int k,
                                                         float* __restrict__ out_val, int64_t* __restrict__ out_idx) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int64_t cands = records * PDM_TOPK_SLOTS;
    float last_v = -INFINITY;
    int last_i = -1;
    for (int r = 0; r < k; ++r) {
        float bv = INFINITY;
        int bi = 0x7fffffff;
        for (int64_t c = lane; c < cands; c += 32) {
            const int64_t rec = c / PDM_TOPK_SLOTS, slot = c - rec * PDM_TOPK_SLOTS;
            const int64_t at = (rec * M + row) * PDM_TOPK_SLOTS + slot;
            const int i = __ldg(idx + at);
            if (i < 0) continue;
            const float v = __ldg(val + at);
            const bool above = v > last_v || (v == last_v && i > last_i);
            if (above && (v < bv || (v == bv && i < bi))) { bv = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        const bool found = bi != 0x7fffffff;
        if (lane == 0) {
            out_val[row * k + r] = found ? bv : INFINITY;
            out_idx[row * k + r] = found ? (int64_t)bi + index_offset : -1;
        }
        if (!found) { bv = INFINITY; bi = 0x7fffffff; }
        last_v = bv; last_i = bi;
    }
}

}  // namespace pdm

extern "C" int pdm_topk_merge(const float* topk_val, const int32_t* topk_idx, int64_t M, int64_t records, int64_t index_offset,
                              int32_t k, float* out_val, int64_t* out_idx, pdm_stream_t stream) {
    PDM_REQUIRE(topk_val && topk_idx && out_val && out_idx && M >= 0 && records >= 1 && k >= 1 && k <= PDM_TOPK_SLOTS,
                "pdm_topk_merge: bad arguments (1 <= k <= %d)", PDM_TOPK_SLOTS);
    if (M == 0) return PDM_OK;
    pdm::topk_merge_kernel<<<(unsigned)pdm::ceil_div(M, 8), 256, 0, pdm::as_stream(stream)>>>(topk_val, topk_idx, M, records,
                                                                                             index_offset, (int)k, out_val, out_idx);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}
