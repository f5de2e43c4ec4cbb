// Certified delta posteriors (adaptive precision of the statistics pass); contract in include/pdm_b200.h.
//
// The reference evaluates every (query, point) pair at every temperature (utils/stats.py:80-90, 282-289).  In the
// low-noise part of a schedule the posterior of a row is a delta on its nearest training point to fp32 resolution;
// a ONE-product tensor-core pass (operands rounded to 11 significant bits, |E1 - E| <= delta rigorously) run at a
// fictitious temperature proves it, and the proven rows skip the full-precision pass.
//   screen_temperatures_kernel  1/T' per row                                   (HBM: 8 B read, 4 B written per row)
//   certify_rows_kernel         flags from the merged screening statistics    (8 B read per row)
//   tile_list_kernel            ascending list of row tiles with an unproven row (one block, ballot scan)
//   finalize_rows_kernel        closed form for proven rows; E_min from the split operands in fp64
//                               (one warp per proven row: 8*d B read)
#include "pdm_common.cuh"

#include <cuda_fp8.h>

namespace pdm {

__global__ void __launch_bounds__(256) screen_temperatures_kernel(const float* __restrict__ q_norm,
                                                                  const float* __restrict__ inv_temp, int64_t M,
                                                                  const float* __restrict__ y_norm_max, float g,
                                                                  float e_star, float kappa, float* __restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M) return;
    // operand rounding (2^-10 ||x|| ||y||, with kappa) + the fp32 round-off of the norm expansion itself (two roundings
    // of size 2^-24 (||x||^2 + ||y||^2) in each of the two energies of a gap; matters only for extreme norm ratios)
    const float delta = kappa * 0.0009765625f * sqrtf(q_norm[r]) * sqrtf(__ldg(y_norm_max)) +
                        2.4e-7f * (q_norm[r] + __ldg(y_norm_max));
    const float t = 1.f / inv_temp[r];
    out[r] = e_star / fmaf(g, t, 2.f * delta);
}

__global__ void __launch_bounds__(256) certify_rows_kernel(const float* __restrict__ screen_out, int64_t M, float a1_max,
                                                           uint8_t* __restrict__ flags) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M) return;
    const float l = screen_out[PDM_OUT_L * M + r];
    const float a1 = screen_out[PDM_OUT_MEAN_E * M + r] * l;
    // NaN / inf fail both comparisons: such rows stay on the full-precision path
    flags[r] = (l > 0.99f && l < 1.25f && a1 >= 0.f && a1 < a1_max) ? 1 : 0;
}

// flags[r] = 1 when the posterior of row r is a delta to fp32 resolution: every other weight together is at most one
// ulp of the dominant one (l - 1 <= 2^-23).  NaN fails the comparison: such rows stay on the contraction path.
__global__ void __launch_bounds__(256) delta_flags_kernel(const float* __restrict__ l, int64_t M, uint8_t* __restrict__ flags) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M) return;
    flags[r] = (l[r] - 1.f <= 1.1920929e-7f) ? 1 : 0;
}

// Second stage of a cascade, run over the listed row tiles of the first: rows the first stage left open take the second
// stage's verdict and arg-min; everything else is untouched.  One block per slot of the list, rows_per_tile <= 1024 threads.
__global__ void __launch_bounds__(1024) merge_stage_kernel(const int32_t* __restrict__ tile_list, const int32_t* __restrict__ n_tiles_dev,
                                                           int32_t rows_per_tile, int64_t M, const uint8_t* __restrict__ flags_b,
                                                           const int64_t* __restrict__ arg_b, uint8_t* __restrict__ flags,
                                                           int64_t* __restrict__ arg) {
    if ((int)blockIdx.x >= __ldg(n_tiles_dev)) return;
    for (int i = threadIdx.x; i < rows_per_tile; i += blockDim.x) {
        const int64_t r = (int64_t)__ldg(tile_list + blockIdx.x) * rows_per_tile + i;
        if (r < M && flags[r] == 0) { flags[r] = flags_b[r]; arg[r] = arg_b[r]; }
    }
}

// out[r, :] = src[idx[r] - index_offset, :] for flagged rows whose index falls in this shard, zeros for flagged rows owned
// by another shard (the shards' outputs are summed), untouched otherwise.  One warp per row, 128-bit copies when aligned.
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, int64_t lds, int64_t d,
                                                          const int64_t* __restrict__ idx, int64_t index_offset, int64_t n_local,
                                                          const uint8_t* __restrict__ flags, int64_t M, float* __restrict__ out,
                                                          int64_t ldo, int vec) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= M || (flags && flags[r] == 0)) return;
    const int64_t j = idx[r] - index_offset;
    const bool own = j >= 0 && j < n_local;
    float* o = out + r * ldo;
    const float* s = src + (own ? j : 0) * lds;
    if (vec) {
        for (int64_t k = lane; k < (d >> 2); k += 32)
            reinterpret_cast<float4*>(o)[k] = own ? __ldg(reinterpret_cast<const float4*>(s) + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        for (int64_t k = lane; k < d; k += 32) o[k] = own ? __ldg(s + k) : 0.f;
    }
}

// One block walks the tiles in order; a tile is listed unless every one of its rows is certified.
__global__ void __launch_bounds__(1024) tile_list_kernel(const uint8_t* __restrict__ flags, int64_t M, int32_t rows_per_tile,
                                                         int32_t* __restrict__ tile_list, int32_t* __restrict__ n_out) {
    __shared__ int warp_count[32];
    __shared__ int base;
    const int64_t tiles = ceil_div(M, (int64_t)rows_per_tile);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int64_t t0 = 0; t0 < tiles; t0 += blockDim.x) {
        const int64_t t = t0 + threadIdx.x;
        bool listed = false;
        if (t < tiles) {
            const int64_t r0 = t * rows_per_tile, r1 = min(M, r0 + rows_per_tile);
            if (r1 - r0 == rows_per_tile && rows_per_tile % 16 == 0 && ((reinterpret_cast<uintptr_t>(flags) + r0) & 15) == 0) {
                // a whole tile: independent 128-bit loads, any zero byte lists it (a serial byte scan of a fully certified
                // tile -- the common case of a low-noise sampling step -- took 17 us per launch)
                const uint4* f4 = reinterpret_cast<const uint4*>(flags + r0);
                uint32_t zero = 0;
                for (int i = 0; i < rows_per_tile / 16; ++i) {
                    const uint4 v = __ldg(f4 + i);
                    zero |= ((v.x - 0x01010101u) & ~v.x) | ((v.y - 0x01010101u) & ~v.y) |
                            ((v.z - 0x01010101u) & ~v.z) | ((v.w - 0x01010101u) & ~v.w);
                }
                listed = (zero & 0x80808080u) != 0;
            } else {
                for (int64_t r = r0; r < r1 && !listed; ++r) listed = flags[r] == 0;
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, listed);
        if (lane == 0) warp_count[warp] = __popc(bal);
        __syncthreads();
        int before = base;
        for (int w = 0; w < warp; ++w) before += warp_count[w];
        if (listed) tile_list[before + __popc(bal & ((1u << lane) - 1u))] = (int32_t)t;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += warp_count[w];
            base += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = base;
}

// E4M3 operands for the first stage of the screening cascade: out8[r,k] = e4m3((hi + lo)[r,k] / 16) (|hi| <= 4096, so the
// bytes stay within +-256 of E4M3's +-448), and err[r] = ||(hi + lo)_r - 16 * out8_r|| in the split's scaled units --
// the EXACT rounding deviation of that row (fp64 accumulation), which is what makes the stage's error bound rigorous
// without any assumption about the format.  One warp per row; columns d..ld8-1 are zero.
__global__ void __launch_bounds__(256) split_to_e4m3_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo,
                                                            int64_t ldh, int64_t rows, int64_t d, uint8_t* __restrict__ out8,
                                                            int64_t ld8, float* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= rows) return;
    const __half2* h2 = reinterpret_cast<const __half2*>(hi + r * ldh);
    const __half2* l2 = lo ? reinterpret_cast<const __half2*>(lo + r * ldh) : nullptr;
    uint16_t* o2 = reinterpret_cast<uint16_t*>(out8 + r * ld8);
    double acc = 0.0;
    for (int64_t i = lane; i < (ld8 >> 1); i += 32) {            // ld8 is even; pairs of columns
        float2 v = make_float2(0.f, 0.f);
        if (2 * i < d) {                                         // d is even on the split path (ldh % 8 == 0 == d % 8)
            const float2 a = __half22float2(h2[i]);
            const float2 c = l2 ? __half22float2(l2[i]) : make_float2(0.f, 0.f);
            v = make_float2(a.x + c.x, a.y + c.y);               // hi + lo rounded to fp32: 2^-24 relative, against the
                                                                 // byte's 2^-4 -- inside the round-up factor below
        }
        const __nv_fp8_storage_t b0 = __nv_cvt_float_to_fp8(v.x * 0.0625f, __NV_SATFINITE, __NV_E4M3);
        const __nv_fp8_storage_t b1 = __nv_cvt_float_to_fp8(v.y * 0.0625f, __NV_SATFINITE, __NV_E4M3);
        o2[i] = (uint16_t)b0 | ((uint16_t)b1 << 8);
        const float d0 = __half2float(__half(__nv_cvt_fp8_to_halfraw(b0, __NV_E4M3))) * 16.f;
        const float d1 = __half2float(__half(__nv_cvt_fp8_to_halfraw(b1, __NV_E4M3))) * 16.f;
        const double e0 = (double)v.x - (double)d0, e1 = (double)v.y - (double)d1;
        acc += e0 * e0 + e1 * e1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0 && err) err[r] = (float)(sqrt(acc) * (1.0 + 1e-5));      // rounded up
}

// 1/T' for the E4M3 stage: delta = kappa * (ex * ||y||max + (||x|| + ex) * ey) with the exact per-row deviations.
__global__ void __launch_bounds__(256) screen_temperatures_f8_kernel(const float* __restrict__ q_norm, const float* __restrict__ q_err,
                                                                     const float* __restrict__ q_inv_scale,
                                                                     const float* __restrict__ inv_temp, int64_t M,
                                                                     const float* __restrict__ y_norm_max,
                                                                     const float* __restrict__ y_err_max, float g, float e_star,
                                                                     float kappa, float* __restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M) return;
    const float ex = q_err[r] * q_inv_scale[r], ey = __ldg(y_err_max);
    const float delta = kappa * (ex * sqrtf(__ldg(y_norm_max)) + (sqrtf(q_norm[r]) + ex) * ey) +
                        2.4e-7f * (q_norm[r] + __ldg(y_norm_max));          // + fp32 round-off of the norm expansion
    const float t = 1.f / inv_temp[r];
    out[r] = e_star / fmaf(g, t, 2.f * delta);
}

struct FinalizeParams {
    const uint8_t* flags; const int64_t* screen_argmin; int64_t M, d;
    const uint16_t* q_hi; const uint16_t* q_lo; int64_t ldqh; const float* q_inv_scale; const float* q_norm;
    const uint16_t* y_hi; const uint16_t* y_lo; int64_t ldyh; float y_inv_scale;
    const float* y_norm; const float* y_aux; int64_t index_offset, n_local; float neg_log_n;
    float* out; int64_t* argmin;
};

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 p = __half22float2(h[i]);
        f[2 * i] = p.x; f[2 * i + 1] = p.y;
    }
}

__global__ void __launch_bounds__(256) finalize_rows_kernel(const FinalizeParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= p.M || !p.flags[r]) return;                       // warp-uniform
    const int64_t k = p.screen_argmin[r];
    const int64_t j = k - p.index_offset;
    const int64_t M = p.M;
    if (lane == 0) {
        p.out[PDM_OUT_LOG_L * M + r] = 0.f;
        p.out[PDM_OUT_MEAN_E * M + r] = 0.f;
        p.out[PDM_OUT_MEAN_E2 * M + r] = 0.f;
        p.out[PDM_OUT_VAR_E * M + r] = 0.f;
        p.out[PDM_OUT_ENTROPY * M + r] = p.neg_log_n;
        p.out[PDM_OUT_L * M + r] = 1.f;
        if (p.argmin) p.argmin[r] = k;
    }
    if (j < 0 || j >= p.n_local) {
        // the arg-min lives on another shard: its owner writes E_min and the aux value; the caller combines the
        // shards' rows with MIN (E_min: +inf here) and MAX (aux: -inf here) -- both leave the rows of uncertified
        // queries, which are identical on every shard, as they are
        if (lane == 0) { p.out[PDM_OUT_E_MIN * M + r] = INFINITY; p.out[PDM_OUT_AUX_MEAN * M + r] = -INFINITY; }
        return;
    }
    const uint4* qh = reinterpret_cast<const uint4*>(p.q_hi + r * p.ldqh);
    const uint4* ql = p.q_lo ? reinterpret_cast<const uint4*>(p.q_lo + r * p.ldqh) : nullptr;
    const uint4* yh = reinterpret_cast<const uint4*>(p.y_hi + j * p.ldyh);
    const uint4* yl = p.y_lo ? reinterpret_cast<const uint4*>(p.y_lo + j * p.ldyh) : nullptr;
    const int64_t chunks = ceil_div(p.d, 8);                   // columns d..ld-1 of the split operands are zero
    double acc = 0.0;
    for (int64_t c = lane; c < chunks; c += 32) {
        float a[8], al[8], b[8], bl[8];
        unpack8(__ldg(qh + c), a);
        unpack8(__ldg(yh + c), b);
        if (ql) unpack8(__ldg(ql + c), al);
        if (yl) unpack8(__ldg(yl + c), bl);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const double x = (double)a[i] + (ql ? (double)al[i] : 0.0);
            const double y = (double)b[i] + (yl ? (double)bl[i] : 0.0);
            acc = fma(x, y, acc);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        // the fused pass's formula (stats_tcgen05.cu epilogue) on the exactly rounded dot product
        const float neg2inv = -2.f * p.q_inv_scale[r] * p.y_inv_scale;
        const float u = __fadd_rn(fmaf((float)acc, neg2inv, p.q_norm[r]), p.y_norm[j]);
        p.out[PDM_OUT_E_MIN * M + r] = 0.5f * u;
        p.out[PDM_OUT_AUX_MEAN * M + r] = p.y_aux ? p.y_aux[j] : 0.f;
    }
}

}  // namespace pdm

using namespace pdm;

extern "C" int pdm_screen_temperatures(const float* q_norm, const float* inv_temp, int64_t M, const float* y_norm_max,
                                       float g, float e_star, float kappa, float* inv_temp_screen, pdm_stream_t stream) {
    PDM_REQUIRE(q_norm && inv_temp && y_norm_max && inv_temp_screen && M >= 0, "pdm_screen_temperatures: bad arguments");
    PDM_REQUIRE(g > 0.f && e_star >= 2.f && e_star <= 80.f && kappa >= 1.f,
                "pdm_screen_temperatures: need g > 0, 2 <= e_star <= 80 (e exp(-e) must stay a normal fp32), kappa >= 1");
    if (M == 0) return PDM_OK;
    screen_temperatures_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, as_stream(stream)>>>(q_norm, inv_temp, M, y_norm_max, g,
                                                                                         e_star, kappa, inv_temp_screen);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}

extern "C" int pdm_screen_certify(const float* screen_out, int64_t M, float e_star, int32_t rows_per_tile,
                                  uint8_t* flags, int32_t* tile_list, int32_t* n_tiles_out, pdm_stream_t stream) {
    PDM_REQUIRE(screen_out && flags && tile_list && n_tiles_out && M >= 0 && rows_per_tile >= 1,
                "pdm_screen_certify: bad arguments");
    PDM_REQUIRE(e_star >= 2.f && e_star <= 80.f, "pdm_screen_certify: 2 <= e_star <= 80");
    PDM_REQUIRE(ceil_div(M, (int64_t)rows_per_tile) < (1ll << 31), "pdm_screen_certify: too many row tiles");
    const float a1_max = 0.9f * e_star * expf(-e_star);
    if (M > 0) {
        certify_rows_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, as_stream(stream)>>>(screen_out, M, a1_max, flags);
        PDM_CUDA_CHECK(cudaGetLastError());
    }
    tile_list_kernel<<<1, 1024, 0, as_stream(stream)>>>(flags, M, rows_per_tile, tile_list, n_tiles_out);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}

extern "C" int pdm_delta_tile_list(const float* l, int64_t M, int32_t rows_per_tile, uint8_t* flags, int32_t* tile_list,
                                   int32_t* n_tiles_out, pdm_stream_t stream) {
    PDM_REQUIRE(l && flags && tile_list && n_tiles_out && M >= 0 && rows_per_tile > 0, "pdm_delta_tile_list: bad arguments");
    if (M > 0) {
        delta_flags_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, as_stream(stream)>>>(l, M, flags);
        PDM_CUDA_CHECK(cudaGetLastError());
    }
    tile_list_kernel<<<1, 1024, 0, as_stream(stream)>>>(flags, M, rows_per_tile, tile_list, n_tiles_out);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}

extern "C" int pdm_screen_merge_stage(const int32_t* tile_list, const int32_t* n_tiles_dev, int64_t max_tiles, int32_t rows_per_tile,
                                      int64_t M, const uint8_t* flags_b, const int64_t* arg_b, uint8_t* flags, int64_t* arg,
                                      pdm_stream_t stream) {
    PDM_REQUIRE(tile_list && n_tiles_dev && flags_b && arg_b && flags && arg && M >= 0 && rows_per_tile > 0 && max_tiles >= 0,
                "pdm_screen_merge_stage: bad arguments");
    if (M == 0 || max_tiles == 0) return PDM_OK;
    merge_stage_kernel<<<(unsigned)max_tiles, (unsigned)std::min<int64_t>(1024, round_up(rows_per_tile, 32)), 0, as_stream(stream)>>>(
        tile_list, n_tiles_dev, rows_per_tile, M, flags_b, arg_b, flags, arg);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}

extern "C" int pdm_gather_rows_f32(const float* src, int64_t lds, int64_t n_local, int64_t d, const int64_t* idx,
                                   int64_t index_offset, const uint8_t* flags, int64_t M, float* out, int64_t ldo,
                                   pdm_stream_t stream) {
    PDM_REQUIRE(src && idx && out && n_local > 0 && d > 0 && lds >= d && ldo >= d && M >= 0, "pdm_gather_rows_f32: bad arguments");
    if (M == 0) return PDM_OK;
    const bool vec = d % 4 == 0 && lds % 4 == 0 && ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    gather_rows_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, as_stream(stream)>>>(src, lds, d, idx, index_offset, n_local, flags, M,
                                                                              out, ldo, vec ? 1 : 0);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}

extern "C" int pdm_split_to_e4m3(const uint16_t* hi, const uint16_t* lo, int64_t ldh, int64_t rows, int64_t d,
                                 uint8_t* out8, int64_t ld8, float* err, pdm_stream_t stream) {
    PDM_REQUIRE(hi && out8 && rows >= 0 && d > 0 && d % 2 == 0 && ldh >= d && ldh % 2 == 0 && ld8 >= d && ld8 % 16 == 0,
                "pdm_split_to_e4m3: bad arguments (d even, ld8 a multiple of 16)");
    if (rows == 0) return PDM_OK;
    split_to_e4m3_kernel<<<(unsigned)ceil_div(rows * 32, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const __half*>(hi), reinterpret_cast<const __half*>(lo), ldh, rows, d, out8, ld8, err);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}

extern "C" int pdm_screen_temperatures_f8(const float* q_norm, const float* q_err, const float* q_inv_scale,
                                          const float* inv_temp, int64_t M, const float* y_norm_max, const float* y_err_max,
                                          float g, float e_star, float kappa, float* inv_temp_screen, pdm_stream_t stream) {
    PDM_REQUIRE(q_norm && q_err && q_inv_scale && inv_temp && y_norm_max && y_err_max && inv_temp_screen && M >= 0,
                "pdm_screen_temperatures_f8: bad arguments");
    PDM_REQUIRE(g > 0.f && e_star >= 2.f && e_star <= 80.f && kappa >= 1.f, "pdm_screen_temperatures_f8: bad parameters");
    if (M == 0) return PDM_OK;
    screen_temperatures_f8_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, as_stream(stream)>>>(
        q_norm, q_err, q_inv_scale, inv_temp, M, y_norm_max, y_err_max, g, e_star, kappa, inv_temp_screen);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}

extern "C" int pdm_screen_tile_list(const uint8_t* flags, int64_t M, int32_t rows_per_tile, int32_t* tile_list,
                                    int32_t* n_tiles_out, pdm_stream_t stream) {
    PDM_REQUIRE(flags && tile_list && n_tiles_out && M >= 0 && rows_per_tile >= 1, "pdm_screen_tile_list: bad arguments");
    tile_list_kernel<<<1, 1024, 0, as_stream(stream)>>>(flags, M, rows_per_tile, tile_list, n_tiles_out);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}

extern "C" int pdm_screen_finalize(const uint8_t* flags, const int64_t* screen_argmin, int64_t M, int64_t d,
                                   const uint16_t* q_hi, const uint16_t* q_lo, int64_t ldqh, const float* q_inv_scale,
                                   const float* q_norm,
                                   const uint16_t* y_hi, const uint16_t* y_lo, int64_t ldyh, float y_inv_scale,
                                   const float* y_norm, const float* y_aux, int64_t index_offset, int64_t n_local,
                                   int64_t n_total, float* out, int64_t* argmin, pdm_stream_t stream) {
    PDM_REQUIRE(flags && screen_argmin && q_hi && q_inv_scale && q_norm && y_hi && y_norm && out && M >= 0 && d > 0 &&
                n_local > 0 && n_total >= n_local, "pdm_screen_finalize: bad arguments");
    PDM_REQUIRE(ldqh % 8 == 0 && ldyh % 8 == 0 && ldqh >= d && ldyh >= d &&
                (reinterpret_cast<uintptr_t>(q_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(y_hi) & 15) == 0 &&
                (!q_lo || (reinterpret_cast<uintptr_t>(q_lo) & 15) == 0) && (!y_lo || (reinterpret_cast<uintptr_t>(y_lo) & 15) == 0),
                "pdm_screen_finalize: split operands must be 16-byte aligned with leading dimensions that are multiples of 8");
    if (M == 0) return PDM_OK;
    FinalizeParams p;
    p.flags = flags; p.screen_argmin = screen_argmin; p.M = M; p.d = d;
    p.q_hi = q_hi; p.q_lo = q_lo; p.ldqh = ldqh; p.q_inv_scale = q_inv_scale; p.q_norm = q_norm;
    p.y_hi = y_hi; p.y_lo = y_lo; p.ldyh = ldyh; p.y_inv_scale = y_inv_scale;
    p.y_norm = y_norm; p.y_aux = y_aux; p.index_offset = index_offset; p.n_local = n_local;
    p.neg_log_n = -logf((float)n_total);
    p.out = out; p.argmin = argmin;
    finalize_rows_kernel<<<(unsigned)ceil_div(M * 32, 256), 256, 0, as_stream(stream)>>>(p);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}
