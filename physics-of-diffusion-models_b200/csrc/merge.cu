// Merge of partial Boltzmann records: across the N splits of one launch and across dataset shards
// (GPUs) after an all-gather.  HBM-bound: reads n_src * 32 B per row, writes 8 floats + one int64.
// Finalises the quantities the reference reports (utils/stats.py:83-90, 284-289).
#include "pdm_common.cuh"
#include "online_stats.cuh"

namespace pdm {

__device__ __forceinline__ void gather_row(const float* __restrict__ parts, int64_t row, int64_t n_outer, int64_t outer_stride,
                                           int64_t n_inner, int64_t row_stride, float it, RowState& acc) {
    state_init(acc);
    for (int64_t o = 0; o < n_outer; ++o) {
        const float* base = parts + o * outer_stride + row * row_stride;
        for (int64_t i = 0; i < n_inner; ++i) {
            RowState s;
            state_load(s, base + i * PDM_PART_STRIDE);
            state_merge(acc, s, it);
        }
    }
}

// Combine records without finalising: one record per row (what a rank sends to its peers).
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ parts, int64_t M, int64_t n_outer,
                                                              int64_t outer_stride, int64_t n_inner, int64_t row_stride,
                                                              const float* __restrict__ inv_temp, float* __restrict__ out) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= M) return;
    RowState acc;
    gather_row(parts, row, n_outer, outer_stride, n_inner, row_stride, inv_temp[row], acc);
    state_store(acc, out + row * PDM_PART_STRIDE);
}

__global__ void __launch_bounds__(256) merge_partials_kernel(const float* __restrict__ parts, int64_t M, int64_t n_outer,
                                                             int64_t outer_stride, int64_t n_inner, int64_t row_stride,
                                                             const float* __restrict__ inv_temp, float log_n,
                                                             float* __restrict__ out, int64_t* __restrict__ argmin) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= M) return;
    RowState acc;
    gather_row(parts, row, n_outer, outer_stride, n_inner, row_stride, inv_temp[row], acc);
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
                                  int64_t n_inner, int64_t row_stride, const float* inv_temp, int64_t n_total,
                                  float* out, int64_t* argmin, pdm_stream_t stream) {
    PDM_REQUIRE(parts && inv_temp && out && M >= 0 && n_outer >= 1 && n_inner >= 1 && n_total >= 1,
                "pdm_merge_partials: bad arguments");
    PDM_REQUIRE(row_stride >= n_inner * PDM_PART_STRIDE && (row_stride % 4) == 0 && (outer_stride % 4) == 0,
                "pdm_merge_partials: strides must cover the records and keep 16-byte alignment");
    if (M == 0) return PDM_OK;
    merge_partials_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, as_stream(stream)>>>(
        parts, M, n_outer, outer_stride, n_inner, row_stride, inv_temp, logf((float)n_total), out, argmin);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}

extern "C" int pdm_reduce_partials(const float* parts, int64_t M, int64_t n_outer, int64_t outer_stride,
                                   int64_t n_inner, int64_t row_stride, const float* inv_temp, float* out_records,
                                   pdm_stream_t stream) {
    PDM_REQUIRE(parts && inv_temp && out_records && M >= 0 && n_outer >= 1 && n_inner >= 1,
                "pdm_reduce_partials: bad arguments");
    PDM_REQUIRE(row_stride >= n_inner * PDM_PART_STRIDE && (row_stride % 4) == 0 && (outer_stride % 4) == 0,
                "pdm_reduce_partials: strides must cover the records and keep 16-byte alignment");
    if (M == 0) return PDM_OK;
    reduce_partials_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, as_stream(stream)>>>(
        parts, M, n_outer, outer_stride, n_inner, row_stride, inv_temp, out_records);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}
