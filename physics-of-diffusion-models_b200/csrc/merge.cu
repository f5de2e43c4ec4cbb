// Merge of partial Boltzmann records: across the N splits of one launch and across dataset shards
// (GPUs) after an all-gather.  HBM-bound: reads n_src * 32 B per row, writes 8 floats + one int64.
// Finalises the quantities the reference reports (utils/stats.py:83-90, 284-289).
#include "pdm_common.cuh"
#include "online_stats.cuh"

namespace pdm {

__device__ __forceinline__ void gather_row(const float* __restrict__ parts, int64_t row, int64_t n_outer, int64_t outer_stride,
                                           int64_t n_inner, int64_t row_stride, int64_t inner_stride, float it, RowState& acc) {
    state_init(acc);
    // the record after the one being merged is already in flight (a thread's loads do not depend on its exp-heavy
    // merge chain); records are stored record-major, so the threads of a warp read consecutive 32-byte records
    const int64_t total = n_outer * n_inner;
    if (total == 0) return;
    float4 n0 = __ldg(reinterpret_cast<const float4*>(parts + row * row_stride));
    float4 n1 = __ldg(reinterpret_cast<const float4*>(parts + row * row_stride) + 1);
    for (int64_t t = 0; t < total; ++t) {
        const float4 c0 = n0, c1 = n1;
        if (t + 1 < total) {
            const int64_t o = (t + 1) / n_inner, i = (t + 1) - o * n_inner;
            const float4* nx = reinterpret_cast<const float4*>(parts + o * outer_stride + row * row_stride + i * inner_stride);
            n0 = __ldg(nx);
            n1 = __ldg(nx + 1);
        }
        RowState s;
        s.m = c0.x; s.l = c0.y; s.a1 = c0.z; s.a2 = c0.w; s.aux = c1.x;
        s.idx = ((long long)__float_as_int(c1.z) << 32) | (long long)(unsigned)__float_as_int(c1.y);
        state_merge(acc, s, it);
    }
}

// Combine records without finalising: one record per row (what a rank sends to its peers).
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ parts, int64_t M, int64_t n_outer,
                                                              int64_t outer_stride, int64_t n_inner, int64_t row_stride,
                                                              int64_t inner_stride, const float* __restrict__ inv_temp,
                                                              float* __restrict__ out) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= M) return;
    RowState acc;
    gather_row(parts, row, n_outer, outer_stride, n_inner, row_stride, inner_stride, inv_temp[row], acc);
    state_store(acc, out + row * PDM_PART_STRIDE);
}

__global__ void __launch_bounds__(256) merge_partials_kernel(const float* __restrict__ parts, int64_t M, int64_t n_outer,
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
