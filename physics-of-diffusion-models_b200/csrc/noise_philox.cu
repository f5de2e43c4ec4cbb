// Noised queries straight into tensor-core operands: the standard-normal draws of torch.randn are regenerated
// in-kernel, element for element, instead of being written to HBM by torch and read back by pdm_prepare_rows.
//
// The reference noises its queries with `torch.randn(...) * t.sqrt() + x0` once per temperature (utils/stats.py:74,
// :273).  On a CUDA device torch.randn(numel) runs a grid-stride kernel of G = 256 * min(SMs * (maxThreadsPerSM/256),
// ceil(numel/256)) threads; thread idx owns the Philox4x32-10 subsequence idx of (seed, offset) and its k-th
// curand_normal4 call fills elements idx + G*ii + 4*G*k, ii = 0..3 (ATen/native/cuda/DistributionTemplates.h:
// calc_execution_policy + distribution_elementwise_grid_stride_kernel).  This kernel walks the same index map with the
// same cuRAND device functions, so the values are bit-identical to torch's; the host verifies that once per device
// against torch.randn itself and falls back to torch.randn + pdm_prepare_rows if it ever is not (pdm_b200/engine.py).
//
// One launch covers n_draws temperatures (blockIdx.y): v = fl(fl(eps * sigma_t) + x0[b, k]) (torch's two roundings),
// operand split with the per-row power-of-two scale derived from the a-priori bound |v| <= max|x0_b| + 6.8 sigma_t
// (|eps| <= sqrt(-2 ln 2^-33) = 6.76 for 32-bit Box-Muller), so no pass over the row is needed before the split.
#include "pdm_common.cuh"

#include <curand_kernel.h>
#include <stdlib.h>

namespace pdm {

struct NoisedParams {
    unsigned long long seed, offset, offset_step;
    long long draw_threads;             // G: threads of the torch.randn launch being reproduced
    const float* x0; long long b, d, ld_x0;
    const float* sigma;                 // [n_draws]
    const float* x0_absmax;             // [b]
    float* x_out; long long ldx;        // optional fp32 rows
    __half* hi; __half* lo; long long ldh; float* inv_scale;   // optional operand split
};

__device__ __forceinline__ float bound_scale(float bound) {
    int e = 0;
    if (!(bound > 0.f) || !(bound < INFINITY)) return 1.f;
    frexpf(bound, &e);                                  // bound = f * 2^e, f in [0.5, 1)
    e = max(-100, min(100, 12 - e));
    return ldexpf(1.f, e);                              // bound * scale in [2^11, 2^12)
}

// inv_scale[t*b + r] = 2^-k of row (t, r), from the bound max|x0_r| + 6.8 sigma_t
__global__ void __launch_bounds__(256) noised_row_scales_kernel(const float* __restrict__ x0_absmax, const float* __restrict__ sigma,
                                                                long long b, long long rows, float* __restrict__ inv_scale) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const long long t = row / b, r = row - t * b;
    inv_scale[row] = 1.f / bound_scale(__fadd_rn(x0_absmax[r], 6.8f * sigma[t]));
}

// 32-bit index arithmetic throughout (one draw stays below 2^31 elements); (row, column) of the four elements a
// curand_normal4 call fills are advanced incrementally instead of divided out.  kX / kSplit select the outputs at
// compile time; ldh == d on the split path, so operand offsets are row * d + k = t * numel + li.
template <bool kX, bool kSplit>
__global__ void __launch_bounds__(256) noised_rows_philox_kernel(NoisedParams p) {
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;                   // < draw_threads by construction
    const unsigned t = blockIdx.y;
    const unsigned numel = (unsigned)(p.b * p.d), G = (unsigned)p.draw_threads, d = (unsigned)p.d;
    const unsigned gq = G / d, gr = G - gq * d;                                   // G = gq * d + gr
    const float sig = p.sigma[t];
    // x0 is contiguous (ld_x0 == d) and so are the outputs (ldx == ldh == d): element li of draw t sits at offset li
    // of the draw's slab, 32-bit offsets from per-draw base pointers
    const float* __restrict__ x0 = p.x0;
    float* __restrict__ xo = kX ? p.x_out + (size_t)t * numel : nullptr;
    __half* __restrict__ hi = kSplit ? p.hi + (size_t)t * numel : nullptr;
    __half* __restrict__ lo = kSplit ? p.lo + (size_t)t * numel : nullptr;
    const float* __restrict__ inv_scale = kSplit ? p.inv_scale + (size_t)t * p.b : nullptr;
    curandStatePhilox4_32_10_t st;
    curand_init(p.seed, (unsigned long long)idx, p.offset + (unsigned long long)t * p.offset_step, &st);
    unsigned li = idx, b = idx / d, k = idx - (idx / d) * d;
    while (li < numel) {                                // same trip structure as torch's rounded_size loop
        const float4 r = curand_normal4(&st);
        const float rv[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            if (li < numel) {
                const float v = __fadd_rn(__fmul_rn(rv[ii], sig), __ldg(x0 + li));
                if (kX) xo[li] = v;
                if (kSplit) {
                    // 1 / inv_scale for a power of two: mirror the exponent field (254 - e), exact
                    const float scale = __uint_as_float(0x7f000000u - __float_as_uint(__ldg(inv_scale + b)));
                    const float vs = v * scale;
                    const __half h = __float2half_rn(vs);
                    hi[li] = h;
                    lo[li] = __float2half_rn(vs - __half2float(h));
                }
            }
            li += G;                                    // next element of this call: G further on
            if (kSplit) { b += gq; k += gr; if (k >= d) { k -= d; ++b; } }
        }
    }
}

// Four adjacent Philox subsequences per thread.  Subsequences 4j .. 4j+3 of torch's launch fill the adjacent elements
// 4j + G*ii + 4G*k + (0..3) (G is a multiple of 256), so a thread that owns all four reads x0 as one float4, looks its
// row and scale up once and stores 8 bytes of hi and of lo (or 16 bytes of fp32) per step instead of four scalar
// round trips -- the per-element index / load / store overhead of the one-subsequence kernel above was about as large
// as Philox + Box-Muller themselves.  Same subsequences, same counters, same cuRAND device functions: the same bits.
// Needs d % 4 == 0 (a group of four never straddles a row) -- guaranteed on the split path (d % 8 == 0).
template <bool kX, bool kSplit>
__global__ void __launch_bounds__(64) noised_rows_philox4_kernel(NoisedParams p) {
    const unsigned idx0 = 4u * (blockIdx.x * blockDim.x + threadIdx.x);          // first subsequence; < draw_threads
    const unsigned t = blockIdx.y;
    const unsigned numel = (unsigned)(p.b * p.d), G = (unsigned)p.draw_threads, d = (unsigned)p.d;
    const unsigned gq = G / d, gr = G - gq * d;
    const float sig = p.sigma[t];
    const float* __restrict__ x0 = p.x0;
    float* __restrict__ xo = kX ? p.x_out + (size_t)t * numel : nullptr;
    __half* __restrict__ hi = kSplit ? p.hi + (size_t)t * numel : nullptr;
    __half* __restrict__ lo = kSplit ? p.lo + (size_t)t * numel : nullptr;
    const float* __restrict__ inv_scale = kSplit ? p.inv_scale + (size_t)t * p.b : nullptr;
    curandStatePhilox4_32_10_t st[4];
#pragma unroll
    for (int s = 0; s < 4; ++s)
        curand_init(p.seed, (unsigned long long)(idx0 + s), p.offset + (unsigned long long)t * p.offset_step, &st[s]);
    unsigned li = idx0, b = idx0 / d, k = idx0 - (idx0 / d) * d;
    while (li < numel) {
        float rv[4][4];                                  // [subsequence][ii]
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const float4 r = curand_normal4(&st[s]);
            rv[s][0] = r.x; rv[s][1] = r.y; rv[s][2] = r.z; rv[s][3] = r.w;
        }
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            if (li < numel) {                            // numel % 4 == 0: the four elements are in or out together
                const float4 x = __ldg(reinterpret_cast<const float4*>(x0 + li));
                float v[4];
                v[0] = __fadd_rn(__fmul_rn(rv[0][ii], sig), x.x);
                v[1] = __fadd_rn(__fmul_rn(rv[1][ii], sig), x.y);
                v[2] = __fadd_rn(__fmul_rn(rv[2][ii], sig), x.z);
                v[3] = __fadd_rn(__fmul_rn(rv[3][ii], sig), x.w);
                if (kX) *reinterpret_cast<float4*>(xo + li) = make_float4(v[0], v[1], v[2], v[3]);
                if (kSplit) {
                    const float scale = __uint_as_float(0x7f000000u - __float_as_uint(__ldg(inv_scale + b)));
                    __half h[4], l[4];
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const float vs = v[s] * scale;
                        h[s] = __float2half_rn(vs);
                        l[s] = __float2half_rn(vs - __half2float(h[s]));
                    }
                    uint2 ph, pl;
                    ph.x = (unsigned)__half_as_ushort(h[0]) | ((unsigned)__half_as_ushort(h[1]) << 16);
                    ph.y = (unsigned)__half_as_ushort(h[2]) | ((unsigned)__half_as_ushort(h[3]) << 16);
                    pl.x = (unsigned)__half_as_ushort(l[0]) | ((unsigned)__half_as_ushort(l[1]) << 16);
                    pl.y = (unsigned)__half_as_ushort(l[2]) | ((unsigned)__half_as_ushort(l[3]) << 16);
                    *reinterpret_cast<uint2*>(hi + li) = ph;
                    *reinterpret_cast<uint2*>(lo + li) = pl;
                }
            }
            li += G;
            if (kSplit) { b += gq; k += gr; if (k >= d) { k -= d; ++b; } }
        }
    }
}

// ||(hi + lo) * inv_scale||^2 per row, fp64 accumulation: the norm of exactly the vector the tensor cores see.
__global__ void __launch_bounds__(256) split_row_norms_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo,
                                                              long long ldh, const float* __restrict__ inv_scale,
                                                              long long rows, long long d, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const __half2* h2 = reinterpret_cast<const __half2*>(hi + row * ldh);
    const __half2* l2 = reinterpret_cast<const __half2*>(lo + row * ldh);
    double acc = 0.0;
    for (long long i = lane; i < (d >> 1); i += 32) {              // d is even on this path (ldh % 8 == 0 == d % 8)
        const float2 a = __half22float2(h2[i]), c = __half22float2(l2[i]);
        const double x = (double)a.x + (double)c.x, y = (double)a.y + (double)c.y;
        acc += x * x + y * y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        const double s = (double)inv_scale[row];
        out[row] = (float)(acc * s * s);
    }
}

__global__ void __launch_bounds__(256) row_absmax_kernel(const float* __restrict__ x, long long rows, long long d, long long ld,
                                                         float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    float m = 0.f;
    for (long long i = lane; i < d; i += 32) m = fmaxf(m, fabsf(__ldg(x + row * ld + i)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) out[row] = m;
}

}  // namespace pdm

using namespace pdm;

extern "C" int pdm_noised_rows_philox(uint64_t seed, uint64_t offset, uint64_t offset_step, int64_t draw_threads,
                                      const float* x0, int64_t b, int64_t d, int64_t ld_x0,
                                      const float* sigma, int64_t n_draws, const float* x0_absmax,
                                      float* x_out, int64_t ldx,
                                      uint16_t* hi, uint16_t* lo, int64_t ldh, float* inv_scale, pdm_stream_t stream) {
    PDM_REQUIRE(x0 && sigma && b > 0 && d > 0 && ld_x0 == d && n_draws >= 0,
                "pdm_noised_rows_philox: bad arguments (x0 must be contiguous: ld_x0 == d)");
    PDM_REQUIRE(draw_threads > 0 && draw_threads % 256 == 0 && draw_threads / 256 <= 0x7fffffffll,
                "pdm_noised_rows_philox: draw_threads must be a positive multiple of 256");
    PDM_REQUIRE(b * d < (1ll << 31), "pdm_noised_rows_philox: one draw must stay below 2^31 elements (torch splits larger ones)");
    PDM_REQUIRE(x_out || hi, "pdm_noised_rows_philox: no output requested");
    PDM_REQUIRE(!x_out || ldx == d, "pdm_noised_rows_philox: x_out must be contiguous (ldx == d)");
    PDM_REQUIRE((hi == nullptr) == (lo == nullptr), "pdm_noised_rows_philox: hi and lo go together");
    PDM_REQUIRE(!hi || (x0_absmax && inv_scale && ldh == d && d % 8 == 0),
                "pdm_noised_rows_philox: the split needs x0_absmax, inv_scale and ldh == d with d %% 8 == 0");
    PDM_REQUIRE(n_draws <= 65535, "pdm_noised_rows_philox: at most 65535 draws per launch");
    if (n_draws == 0) return PDM_OK;
    NoisedParams p{seed, offset, offset_step, draw_threads, x0, b, d, ld_x0, sigma, x0_absmax, x_out, ldx,
                   reinterpret_cast<__half*>(hi), reinterpret_cast<__half*>(lo), ldh, inv_scale};
    if (hi) {
        const int64_t rows = n_draws * b;
        noised_row_scales_kernel<<<(unsigned)ceil_div(rows, 256), 256, 0, as_stream(stream)>>>(x0_absmax, sigma, b, rows, inv_scale);
        PDM_CUDA_CHECK(cudaGetLastError());
    }
    static const bool one_per_thread = getenv("PDM_PHILOX_SCALAR") && atoi(getenv("PDM_PHILOX_SCALAR")) != 0;   // dev knob
    const bool aligned16 = (reinterpret_cast<uintptr_t>(x0) & 15) == 0 && (!x_out || (reinterpret_cast<uintptr_t>(x_out) & 15) == 0) &&
                           (!hi || ((reinterpret_cast<uintptr_t>(hi) & 7) == 0 && (reinterpret_cast<uintptr_t>(lo) & 7) == 0));
    if (d % 4 == 0 && aligned16 && !one_per_thread) {
        // four adjacent subsequences per thread: draw_threads / 4 threads in blocks of 64 = draw_threads / 256 blocks
        dim3 grid4((unsigned)(draw_threads / 256), (unsigned)n_draws);
        if (x_out && hi)  noised_rows_philox4_kernel<true, true><<<grid4, 64, 0, as_stream(stream)>>>(p);
        else if (hi)      noised_rows_philox4_kernel<false, true><<<grid4, 64, 0, as_stream(stream)>>>(p);
        else              noised_rows_philox4_kernel<true, false><<<grid4, 64, 0, as_stream(stream)>>>(p);
        PDM_CUDA_CHECK(cudaGetLastError());
        return PDM_OK;
    }
    dim3 grid((unsigned)(draw_threads / 256), (unsigned)n_draws);
    if (x_out && hi)  noised_rows_philox_kernel<true, true><<<grid, 256, 0, as_stream(stream)>>>(p);
    else if (hi)      noised_rows_philox_kernel<false, true><<<grid, 256, 0, as_stream(stream)>>>(p);
    else              noised_rows_philox_kernel<true, false><<<grid, 256, 0, as_stream(stream)>>>(p);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}

extern "C" int pdm_split_row_norms(const uint16_t* hi, const uint16_t* lo, int64_t ldh, const float* inv_scale,
                                   int64_t rows, int64_t d, float* norms, pdm_stream_t stream) {
    PDM_REQUIRE(hi && lo && inv_scale && norms && rows >= 0 && d > 0 && ldh >= d && d % 2 == 0 && ldh % 2 == 0,
                "pdm_split_row_norms: bad arguments");
    if (rows == 0) return PDM_OK;
    split_row_norms_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const __half*>(hi), reinterpret_cast<const __half*>(lo), ldh, inv_scale, rows, d, norms);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}

extern "C" int pdm_row_absmax_f32(const float* x, int64_t rows, int64_t d, int64_t ld, float* out, pdm_stream_t stream) {
    PDM_REQUIRE(x && out && rows >= 0 && d > 0 && ld >= d, "pdm_row_absmax_f32: bad arguments");
    if (rows == 0) return PDM_OK;
    row_absmax_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, as_stream(stream)>>>(x, rows, d, ld, out);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}
