// HBM-bound preparation kernels: row norms, noising + fp16 hi/lo operand split, dataset transpose,
// column moments, and the energy -> normalised-weight pass.  All are single-pass, coalesced,
// 128-bit vectorised where alignment allows.  See include/pdm_b200.h for the reference lines each replaces.
#include "pdm_common.cuh"
#include "online_stats.cuh"

namespace pdm {

// ------------------------------------------------------------------------------------------------
// small reductions
// ------------------------------------------------------------------------------------------------
template <typename T, typename Op>
__device__ __forceinline__ T warp_reduce(T v, Op op) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, Op op, T identity, T* smem /* >= 32 */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    v = warp_reduce(v, op);
    __syncthreads();
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    T r = (threadIdx.x < nwarp) ? smem[threadIdx.x] : identity;
    if (warp == 0) r = warp_reduce(r, op);
    if (threadIdx.x == 0) smem[0] = r;
    __syncthreads();
    return smem[0];
}

struct OpAddD { __device__ double operator()(double a, double b) const { return a + b; } };
struct OpMaxF { __device__ float operator()(float a, float b) const { return fmaxf(a, b); } };
struct OpMinF { __device__ float operator()(float a, float b) const { return fminf(a, b); } };

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ------------------------------------------------------------------------------------------------
// K1 row norms: one warp per row, fp64 accumulation.
// ------------------------------------------------------------------------------------------------
template <bool kVec>
__global__ void __launch_bounds__(256) row_norms_kernel(const float* __restrict__ x, int64_t rows, int64_t d,
                                                        int64_t ld, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* p = x + row * ld;
    double acc = 0.0;
    if (kVec) {
        const int64_t nv = d >> 2;
        for (int64_t i = lane; i < nv; i += 32) {
            const float4 v = ldg_f4(p + 4 * i);
            acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
        }
    } else {
        for (int64_t i = lane; i < d; i += 32) { const float v = __ldg(p + i); acc += (double)v * v; }
    }
    acc = warp_reduce(acc, OpAddD());
    if (lane == 0) out[row] = (float)acc;
}

// ------------------------------------------------------------------------------------------------
// Row preparation (noising, rescale, norm, fp16 split).  One 256-thread block per row.
// ------------------------------------------------------------------------------------------------
struct PrepParams {
    const float* src; int64_t src_rows; int64_t ld_src;
    const float* noise; int64_t ld_noise; const float* sigma; const float* post;
    int64_t rows; int64_t d; float fixed_scale;
    float* x_out; int64_t ldx; float* norms;
    __half* hi; __half* lo; int64_t ldh; float* inv_scale;
};

__device__ __forceinline__ float prep_value(float s, float n, float sig, float post, bool has_noise, bool has_post) {
    float v = s;
    if (has_noise) v = __fadd_rn(__fmul_rn(n, sig), s);   // randn * sqrt(t) + x0, two roundings
    if (has_post) v = __fmul_rn(v, post);
    return v;
}

__device__ __forceinline__ void split_f16(float v, __half& h, __half& l) {
    h = __float2half_rn(v);
    l = __float2half_rn(v - __half2float(h));
}

template <bool kVec>
__global__ void __launch_bounds__(256) prepare_rows_kernel(PrepParams p) {
    __shared__ double red_d[32];
    __shared__ float red_f[32];
    const int64_t row = blockIdx.x;
    const bool has_noise = p.noise != nullptr, has_post = p.post != nullptr;
    const float* s = p.src + (has_noise ? (row % p.src_rows) : row) * p.ld_src;
    const float* n = has_noise ? p.noise + row * p.ld_noise : nullptr;
    const float sig = has_noise ? p.sigma[row] : 0.f;
    const float post = has_post ? p.post[row] : 1.f;
    float* xo = p.x_out ? p.x_out + row * p.ldx : nullptr;

    double acc = 0.0;
    float amax = 0.f;
    if (kVec) {
        const int64_t nv = p.d >> 2;
        for (int64_t i = threadIdx.x; i < nv; i += blockDim.x) {
            const float4 a = ldg_f4(s + 4 * i);
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_noise) b = ldg_f4(n + 4 * i);
            float4 v;
            v.x = prep_value(a.x, b.x, sig, post, has_noise, has_post);
            v.y = prep_value(a.y, b.y, sig, post, has_noise, has_post);
            v.z = prep_value(a.z, b.z, sig, post, has_noise, has_post);
            v.w = prep_value(a.w, b.w, sig, post, has_noise, has_post);
            acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
            amax = fmaxf(fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
            if (xo) *reinterpret_cast<float4*>(xo + 4 * i) = v;
        }
    } else {
        for (int64_t i = threadIdx.x; i < p.d; i += blockDim.x) {
            const float v = prep_value(__ldg(s + i), has_noise ? __ldg(n + i) : 0.f, sig, post, has_noise, has_post);
            acc += (double)v * v;
            amax = fmaxf(amax, fabsf(v));
            if (xo) xo[i] = v;
        }
    }
    if (p.norms) {
        const double tot = block_reduce(acc, OpAddD(), 0.0, red_d);
        if (threadIdx.x == 0) p.norms[row] = (float)tot;
    }
    if (!p.hi) return;

    float scale = p.fixed_scale;
    if (!(scale > 0.f)) {
        amax = block_reduce(amax, OpMaxF(), 0.f, red_f);
        int e = 0;
        scale = 1.f;
        if (amax > 0.f && amax < INFINITY) {
            frexpf(amax, &e);                       // amax = f * 2^e, f in [0.5, 1)
            e = max(-100, min(100, 12 - e));
            scale = ldexpf(1.f, e);                 // amax * scale in [2^11, 2^12)
        }
    }
    if (threadIdx.x == 0 && p.inv_scale) p.inv_scale[row] = 1.f / scale;

    __half* hi = p.hi + row * p.ldh;
    __half* lo = p.lo + row * p.ldh;
    if (kVec) {
        const int64_t nv = p.ldh >> 2, dv = p.d >> 2;   // d % 4 == 0 on this path
        for (int64_t i = threadIdx.x; i < nv; i += blockDim.x) {
            __half h[4], l[4];
            if (i < dv) {
                const float4 a = ldg_f4(s + 4 * i);
                float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                if (has_noise) b = ldg_f4(n + 4 * i);
                split_f16(prep_value(a.x, b.x, sig, post, has_noise, has_post) * scale, h[0], l[0]);
                split_f16(prep_value(a.y, b.y, sig, post, has_noise, has_post) * scale, h[1], l[1]);
                split_f16(prep_value(a.z, b.z, sig, post, has_noise, has_post) * scale, h[2], l[2]);
                split_f16(prep_value(a.w, b.w, sig, post, has_noise, has_post) * scale, h[3], l[3]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) { h[k] = __float2half_rn(0.f); l[k] = h[k]; }
            }
            *reinterpret_cast<uint2*>(hi + 4 * i) = *reinterpret_cast<uint2*>(h);
            *reinterpret_cast<uint2*>(lo + 4 * i) = *reinterpret_cast<uint2*>(l);
        }
    } else {
        for (int64_t i = threadIdx.x; i < p.ldh; i += blockDim.x) {
            __half h = __float2half_rn(0.f), l = h;
            if (i < p.d) {
                const float v = prep_value(__ldg(s + i), has_noise ? __ldg(n + i) : 0.f, sig, post, has_noise, has_post);
                split_f16(v * scale, h, l);
            }
            hi[i] = h; lo[i] = l;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// global abs-max (non-negative floats order like their bit patterns)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, int64_t rows, int64_t d, int64_t ld,
                                                     unsigned* __restrict__ out_bits) {
    __shared__ float red_f[32];
    float amax = 0.f;
    const int64_t total = rows * d;
    if (ld == d && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {          // contiguous: one flat, 128-bit vectorised sweep
        const int64_t n4 = total >> 2;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
            const float4 v = ldg_f4(x + 4 * i);
            amax = fmaxf(fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
        }
        for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
            amax = fmaxf(amax, fabsf(__ldg(x + i)));
    } else {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
            const int64_t r = i / d, c = i - r * d;
            amax = fmaxf(amax, fabsf(__ldg(x + r * ld + c)));
        }
    }
    amax = block_reduce(amax, OpMaxF(), 0.f, red_f);
    if (threadIdx.x == 0) atomicMax(out_bits, __float_as_uint(amax));
}

// ------------------------------------------------------------------------------------------------
// lattice test: is every y*scale an integer the fp16 hi part holds exactly?  One warp per row:
//   out[0] = max_j ||y_j - rint(y_j*scale)/scale||^2 / ||y_j||^2      out[1] = max |rint(y*scale)|
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lattice_residual_kernel(const float* __restrict__ y, int64_t n, int64_t d, int64_t ld,
                                                               float scale, unsigned* __restrict__ out_bits) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const float* p = y + row * ld;
    const double inv = 1.0 / (double)scale;
    double r2 = 0.0, y2 = 0.0;
    float vmax = 0.f;
    for (int64_t i = lane; i < d; i += 32) {
        const float v = __ldg(p + i);
        const float q = rintf(__fmul_rn(v, scale));        // what rn16(v*scale) holds when |q| <= 2048
        const double r = (double)v - (double)q * inv;
        r2 += r * r;
        y2 += (double)v * v;
        vmax = fmaxf(vmax, fabsf(q));
    }
    r2 = warp_reduce(r2, OpAddD());
    y2 = warp_reduce(y2, OpAddD());
    vmax = warp_reduce(vmax, OpMaxF());
    if (lane == 0) {
        const float ratio = y2 > 0.0 ? (float)(r2 / y2) : (r2 > 0.0 ? INFINITY : 0.f);
        atomicMax(out_bits, __float_as_uint(ratio));
        atomicMax(out_bits + 1, __float_as_uint(vmax));
    }
}

// ------------------------------------------------------------------------------------------------
// transposed fp16 split of the dataset: y (n, d) -> yt_hi/lo (d, ldt)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_split_kernel(const float* __restrict__ y, int64_t n, int64_t d, int64_t ld,
                                                              float scale, __half* __restrict__ th, __half* __restrict__ tl,
                                                              int64_t ldt) {
    __shared__ float tile[32][33];
    const int64_t j0 = (int64_t)blockIdx.x * 32, k0 = (int64_t)blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t j = j0 + r, k = k0 + threadIdx.x;
        tile[r][threadIdx.x] = (j < n && k < d) ? __ldg(y + j * ld + k) * scale : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t k = k0 + r, j = j0 + threadIdx.x;
        if (k < d && j < ldt) {
            __half h, l;
            split_f16(tile[threadIdx.x][r], h, l);
            th[k * ldt + j] = h;
            tl[k * ldt + j] = l;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K11 column moments (fp64 sums per column, global min/max)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_min_f(float* addr, float v) {
    int* a = reinterpret_cast<int*>(addr);
    int old = *a;
    while (__int_as_float(old) > v) { const int assumed = old; old = atomicCAS(a, assumed, __float_as_int(v)); if (old == assumed) break; }
}
__device__ __forceinline__ void atomic_max_f(float* addr, float v) {
    int* a = reinterpret_cast<int*>(addr);
    int old = *a;
    while (__int_as_float(old) < v) { const int assumed = old; old = atomicCAS(a, assumed, __float_as_int(v)); if (old == assumed) break; }
}

__global__ void moments_init_kernel(double* sum, double* sumsq, int64_t d, float* minmax) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < d) { sum[i] = 0.0; sumsq[i] = 0.0; }
    if (i == 0) { minmax[0] = INFINITY; minmax[1] = -INFINITY; }
}

// block = (32 columns, 8 row lanes); grid = (ceil(d/32), row slabs)
__global__ void __launch_bounds__(256) column_moments_kernel(const float* __restrict__ y, int64_t n, int64_t d, int64_t ld,
                                                             double* __restrict__ sum, double* __restrict__ sumsq,
                                                             float* __restrict__ minmax) {
    __shared__ double s1[8][33], s2[8][33];
    __shared__ float red_f[32];
    const int64_t k = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int64_t rows_per = ceil_div(n, (int64_t)gridDim.y);
    const int64_t r0 = (int64_t)blockIdx.y * rows_per, r1 = min(n, r0 + rows_per);
    double a1 = 0.0, a2 = 0.0;
    float mn = INFINITY, mx = -INFINITY;
    if (k < d) {
        for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
            const float v = __ldg(y + r * ld + k);
            a1 += v; a2 += (double)v * v; mn = fminf(mn, v); mx = fmaxf(mx, v);
        }
    }
    s1[threadIdx.y][threadIdx.x] = a1; s2[threadIdx.y][threadIdx.x] = a2;
    __syncthreads();
    if (threadIdx.y == 0 && k < d) {
        for (int i = 1; i < 8; ++i) { a1 += s1[i][threadIdx.x]; a2 += s2[i][threadIdx.x]; }
        atomicAdd(sum + k, a1);
        atomicAdd(sumsq + k, a2);
    }
    // flatten thread index for the block-wide min/max
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    mn = warp_reduce(mn, OpMinF()); mx = warp_reduce(mx, OpMaxF());
    __syncthreads();
    if (lane == 0) { red_f[warp] = mn; red_f[8 + warp] = mx; }
    __syncthreads();
    if (tid == 0) {
        for (int i = 1; i < 8; ++i) { mn = fminf(mn, red_f[i]); mx = fmaxf(mx, red_f[8 + i]); }
        if (mn < INFINITY) atomic_min_f(minmax, mn);
        if (mx > -INFINITY) atomic_max_f(minmax + 1, mx);
    }
}

// ------------------------------------------------------------------------------------------------
// energy -> normalised weights p = exp(-(E - m)/T) / l   (scheduler.py:66-68)
// ------------------------------------------------------------------------------------------------
template <bool kP32>
__global__ void __launch_bounds__(256, 3) weights_kernel(const float* __restrict__ energy, int64_t lde, int64_t M, int64_t N,
                                                      const float* __restrict__ e_min, const float* __restrict__ l,
                                                      const float* __restrict__ inv_temp,
                                                      float* __restrict__ p32, int64_t ldp32,
                                                      __half* __restrict__ ph, __half* __restrict__ pl, int64_t ldph, int vec,
                                                      const int32_t* __restrict__ row_tiles, int32_t rows_per_tile,
                                                      const int32_t* __restrict__ n_tiles_dev, int64_t row0) {
    // blockIdx.y = a row (row0 + y), or with a tile list slot (row0 + y) of the listed tiles' rows; blocks beyond the
    // device-side length of the list (or beyond M) have nothing to do
    int64_t row = row0 + blockIdx.y;
    if (row_tiles) {
        const int64_t t = row / rows_per_tile;
        if (n_tiles_dev && t >= __ldg(n_tiles_dev)) return;
        row = (int64_t)__ldg(row_tiles + t) * rows_per_tile + (row - t * rows_per_tile);
        if (row >= M) return;
    }
    const float m = e_min[row], it = inv_temp[row], inv_l = 1.f / l[row];
    const int64_t width = ph ? ldph : N;
    if (vec) {
        // 8 columns per group: two 128-bit energy loads, one 128-bit store per fp16 plane.  A thread takes kU groups per
        // trip with all of their loads issued before the first exponential (128 bytes in flight per thread: the kernel
        // is HBM-bound, 8 bytes per pair, and a block walks a whole row -- or a few blocks share it when rows are few).
        constexpr int kU = 4;
        const int groups = (int)(width >> 3), n32 = (int)N;             // N < 2^31 (checked by the launcher)
        const int stride = (int)(gridDim.x * blockDim.x);
        const float* er = energy + row * lde;
        __half* phr = ph ? ph + row * ldph : nullptr;
        __half* plr = ph ? pl + row * ldph : nullptr;
        float* p32r = (kP32 && p32) ? p32 + row * ldp32 : nullptr;
        for (int g0 = (int)(blockIdx.x * blockDim.x + threadIdx.x); g0 < groups; g0 += kU * stride) {
            float4 ea[kU], eb[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int g = g0 + u * stride, j0 = g << 3;
                if (g < groups && j0 + 8 <= n32) { ea[u] = ldg_f4(er + j0); eb[u] = ldg_f4(er + j0 + 4); }
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int g = g0 + u * stride, j0 = g << 3;
                if (g < groups && j0 + 8 <= n32) {
                    const float ev[8] = {ea[u].x, ea[u].y, ea[u].z, ea[u].w, eb[u].x, eb[u].y, eb[u].z, eb[u].w};
                    float pv[8];
                    uint32_t hw[4], lw[4];             // the fp16 planes, two columns per word
#pragma unroll
                    for (int k = 0; k < 8; k += 2) {
                        __half h0, l0, h1, l1;
                        pv[k] = fast_exp2(-fminf((ev[k] - m) * it, kMaxE) * kLog2e) * inv_l;
                        pv[k + 1] = fast_exp2(-fminf((ev[k + 1] - m) * it, kMaxE) * kLog2e) * inv_l;
                        split_f16(pv[k] * 16384.f, h0, l0);
                        split_f16(pv[k + 1] * 16384.f, h1, l1);
                        hw[k / 2] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                        lw[k / 2] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
                    }
                    if (phr) {
                        *reinterpret_cast<uint4*>(phr + j0) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                        *reinterpret_cast<uint4*>(plr + j0) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                    }
                    if (kP32 && p32r) {
                        *reinterpret_cast<float4*>(p32r + j0) = make_float4(pv[0], pv[1], pv[2], pv[3]);
                        *reinterpret_cast<float4*>(p32r + j0 + 4) = make_float4(pv[4], pv[5], pv[6], pv[7]);
                    }
                } else if (g < groups) {               // the ragged last group of a row (and the padding up to ldph)
                    for (int k = 0; k < 8; ++k) {
                        float pk = 0.f;
                        if (j0 + k < n32) pk = fast_exp2(-fminf((__ldg(er + j0 + k) - m) * it, kMaxE) * kLog2e) * inv_l;
                        if (phr) {
                            __half h, lo;
                            split_f16(pk * 16384.f, h, lo);
                            phr[j0 + k] = h;
                            plr[j0 + k] = lo;
                        }
                        if (p32r && j0 + k < n32) p32r[j0 + k] = pk;
                    }
                }
            }
        }
        // tail columns of an fp32-only output whose width is not a multiple of 8
        if (!ph && blockIdx.x == 0) {
            for (int64_t j = ((int64_t)groups << 3) + threadIdx.x; j < N; j += blockDim.x) {
                const float ee = fminf((__ldg(energy + row * lde + j) - m) * it, kMaxE);
                p32[row * ldp32 + j] = fast_exp2(-ee * kLog2e) * inv_l;
            }
        }
        return;
    }
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < width; j += (int64_t)gridDim.x * blockDim.x) {
        float p = 0.f;
        if (j < N) {
            const float e = fminf((__ldg(energy + row * lde + j) - m) * it, kMaxE);
            p = fast_exp2(-e * kLog2e) * inv_l;
        }
        if (p32 && j < N) p32[row * ldp32 + j] = p;
        if (ph) {
            __half h, lo;
            split_f16(p * 16384.f, h, lo);
            ph[row * ldph + j] = h;
            pl[row * ldph + j] = lo;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward of the ideal denoiser: with p_j = exp(-e_j)/l, e_j = (E_j - m)/T and s_j = s_scale * S_j (= y_j . g),
//   a = sum_j p_j s_j,   w_j = p_j (s_j - a),   sums = (a, sum_j w_j e_j)            one block per query row
// Centring s before the contraction keeps sum_j w_j y_j = Cov_p(s, y) free of the cancellation
// sum p s y - a x0_hat, which the caller would otherwise amplify by 1/T.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) denoiser_backward_weights_kernel(
    const float* __restrict__ energy, int64_t lde, const float* __restrict__ sdot, int64_t lds, int64_t N,
    const float* __restrict__ e_min, const float* __restrict__ l, const float* __restrict__ inv_temp,
    const float* __restrict__ s_scale, const float* __restrict__ a_in, float* __restrict__ w, int64_t ldw,
    float* __restrict__ sums) {
    __shared__ double red_d[32];
    const int64_t row = blockIdx.x;
    const float m = e_min[row], it = inv_temp[row], inv_l = 1.f / l[row], sc = s_scale ? s_scale[row] : 1.f;
    const float* er = energy + row * lde;
    const float* sr = sdot + row * lds;
    float* wr = w + row * ldw;
    float af;
    if (a_in) {
        af = a_in[row];                      // a summed over every shard of the dataset by the caller
    } else {
        double a = 0.0;
        for (int64_t j = threadIdx.x; j < N; j += blockDim.x) {
            const float e = fminf((__ldg(er + j) - m) * it, kMaxE);
            a += (double)(fast_exp2(-e * kLog2e) * inv_l) * (double)(sc * __ldg(sr + j));
        }
        af = (float)block_reduce(a, OpAddD(), 0.0, red_d);
    }
    double b = 0.0;
    for (int64_t j = threadIdx.x; j < N; j += blockDim.x) {
        const float e = fminf((__ldg(er + j) - m) * it, kMaxE);
        const float wj = fast_exp2(-e * kLog2e) * inv_l * (sc * __ldg(sr + j) - af);
        wr[j] = wj;
        b += (double)wj * e;
    }
    b = block_reduce(b, OpAddD(), 0.0, red_d);
    if (threadIdx.x == 0) { sums[2 * row] = af; sums[2 * row + 1] = (float)b; }
}

// ------------------------------------------------------------------------------------------------
// k smallest entries of every row of a dense (rows, n) matrix, ascending, ties by lower column index.
// One block per row, k selection rounds over the (L2-resident) row: round r takes the lexicographic minimum of
// (value, column) strictly above the pair selected in round r-1.  Replaces the CPU k-NN search of
// utils/stats.py:50-60, 138-146 on the distance tiles of pdm_posterior_stats(energy_out).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) topk_smallest_kernel(const float* __restrict__ x, int64_t ldx, int64_t n, int k,
                                                            float* __restrict__ vals, int64_t* __restrict__ idx) {
    __shared__ float sv[8];
    __shared__ long long si[8];
    __shared__ float last_v_s;
    __shared__ long long last_i_s;
    const int64_t row = blockIdx.x;
    const float* xr = x + row * ldx;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float last_v = -INFINITY;
    long long last_i = -1;
    for (int r = 0; r < k; ++r) {
        float bv = INFINITY;
        long long bi = 0x7fffffffffffffffLL;
        for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
            const float v = __ldg(xr + j);
            const bool above = v > last_v || (v == last_v && j > last_i);
            if (above && (v < bv || (v == bv && j < bi))) { bv = v; bi = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w)
                if (sv[w] < bv || (sv[w] == bv && si[w] < bi)) { bv = sv[w]; bi = si[w]; }
            const bool found = bi != 0x7fffffffffffffffLL;
            vals[row * k + r] = found ? bv : INFINITY;
            idx[row * k + r] = found ? bi : -1;
            last_v_s = bv; last_i_s = bi;
        }
        __syncthreads();
        last_v = last_v_s; last_i = last_i_s;
    }
}

// ------------------------------------------------------------------------------------------------
// Refinement of nearest-neighbour candidates: the selection works on ||x||^2 - 2 x.y + ||y||^2 (utils/distance.py:21),
// whose fp32 round-off is 2^-24 (||x||^2 + ||y||^2) -- for tight clusters that is 1e-3 of the neighbour distance itself,
// while sklearn's kneighbors (utils/stats.py:50-60, 138-146) works in float64.  So the k candidates of a row are
// re-evaluated directly, sum_k (x_k - y_k)^2 with fp64 accumulation, and re-sorted (value, then index).  One warp per row.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) refine_neighbours_kernel(const float* __restrict__ x, int64_t ldx, int64_t M, int64_t d,
                                                                const float* __restrict__ y, int64_t ldy, int64_t n_local,
                                                                int64_t index_offset, int k, float* __restrict__ vals,
                                                                int64_t* __restrict__ idx) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const float* xr = x + row * ldx;
    float v[8];
    long long id[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { v[q] = INFINITY; id[q] = -1; }
    for (int q = 0; q < k; ++q) {
        const long long g = idx[row * k + q];
        const long long j = g - index_offset;
        float val = vals[row * k + q];
        if (g >= 0 && j >= 0 && j < n_local) {
            const float* yr = y + j * ldy;
            double acc = 0.0;
            for (int64_t c = lane; c < d; c += 32) {
                const float df = xr[c] - __ldg(yr + c);
                acc += (double)df * (double)df;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            val = (float)acc;
        }
        // insertion into the sorted prefix (every lane holds the same copy)
        int pos = q;
        v[q] = val; id[q] = g;
#pragma unroll
        for (int t = 7; t > 0; --t) {
            if (t <= pos && g >= 0 && (v[t] < v[t - 1] || (v[t] == v[t - 1] && id[t] < id[t - 1]) || id[t - 1] < 0)) {
                const float tv = v[t]; v[t] = v[t - 1]; v[t - 1] = tv;
                const long long ti = id[t]; id[t] = id[t - 1]; id[t - 1] = ti;
            }
        }
    }
    if (lane == 0) {
        for (int q = 0; q < k; ++q) { vals[row * k + q] = v[q]; idx[row * k + q] = id[q]; }
    }
}

// ------------------------------------------------------------------------------------------------
// one reverse-diffusion update: out = c_x0 * x0_hat + c_xt * xt (+ c_noise * noise), the DDPM / DDIM step of
// diffusion/ddpm_sampling.py:94-110 with the x0 / eps algebra of diffusion/ddpm/ddpm.py:17-20 folded into three
// host-computed coefficients.  HBM-bound: 12-16 B per element.
// ------------------------------------------------------------------------------------------------
// `out` may alias `xt` (the sampler updates its state in place): xt is read with plain loads through a pointer that is
// not declared restrict, every thread reads its elements before it writes them.
__global__ void __launch_bounds__(256) sampler_step_kernel(const float* __restrict__ x0_hat, const float* xt,
                                                           const float* __restrict__ noise, float c_x0, float c_xt, float c_noise,
                                                           float* out, int64_t n, int vec, const float* __restrict__ coef) {
    if (coef) { c_x0 = __ldg(coef); c_xt = __ldg(coef + 1); c_noise = __ldg(coef + 2); }     // CUDA-graphed steps: device scalars
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        const int64_t n4 = n >> 2;
        for (; i < n4; i += stride) {
            const float4 a = ldg_f4(x0_hat + 4 * i), b = *reinterpret_cast<const float4*>(xt + 4 * i);
            float4 r = make_float4(fmaf(c_xt, b.x, c_x0 * a.x), fmaf(c_xt, b.y, c_x0 * a.y), fmaf(c_xt, b.z, c_x0 * a.z),
                                   fmaf(c_xt, b.w, c_x0 * a.w));
            if (noise) {
                const float4 e = ldg_f4(noise + 4 * i);
                r.x = fmaf(c_noise, e.x, r.x); r.y = fmaf(c_noise, e.y, r.y); r.z = fmaf(c_noise, e.z, r.z); r.w = fmaf(c_noise, e.w, r.w);
            }
            *reinterpret_cast<float4*>(out + 4 * i) = r;
        }
        i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // scalar tail
    }
    for (; i < n; i += stride) {
        float r = fmaf(c_xt, xt[i], c_x0 * __ldg(x0_hat + i));
        if (noise) r = fmaf(c_noise, __ldg(noise + i), r);
        out[i] = r;
    }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace pdm

using namespace pdm;

extern "C" int pdm_row_norms_f32(const float* x, int64_t rows, int64_t d, int64_t ld, float* out, pdm_stream_t stream) {
    PDM_REQUIRE(x && out && rows >= 0 && d > 0 && ld >= d, "pdm_row_norms_f32: bad arguments");
    if (rows == 0) return PDM_OK;
    const unsigned grid = (unsigned)ceil_div(rows, 8);
    if (aligned16(x) && d % 4 == 0 && ld % 4 == 0)
        row_norms_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(x, rows, d, ld, out);
    else
        row_norms_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(x, rows, d, ld, out);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}

extern "C" int pdm_prepare_rows(const float* src, int64_t src_rows, int64_t ld_src,
                                const float* noise, int64_t ld_noise, const float* sigma, const float* post,
                                int64_t rows, int64_t d, float fixed_scale,
                                float* x_out, int64_t ldx, float* norms,
                                uint16_t* hi, uint16_t* lo, int64_t ldh, float* inv_scale,
                                pdm_stream_t stream) {
    PDM_REQUIRE(src && rows >= 0 && d > 0 && src_rows > 0 && ld_src >= d, "pdm_prepare_rows: bad source");
    PDM_REQUIRE(!noise || (sigma && ld_noise >= d), "pdm_prepare_rows: noise needs sigma and ld_noise >= d");
    PDM_REQUIRE(noise || src_rows >= rows, "pdm_prepare_rows: src_rows < rows without noise");
    PDM_REQUIRE((hi == nullptr) == (lo == nullptr), "pdm_prepare_rows: hi and lo go together");
    PDM_REQUIRE(!hi || (ldh >= d && ldh % 8 == 0 && inv_scale), "pdm_prepare_rows: ldh must be >= d, a multiple of 8, and inv_scale given");
    PDM_REQUIRE(!x_out || ldx >= d, "pdm_prepare_rows: ldx < d");
    if (rows == 0) return PDM_OK;
    PrepParams p{src, src_rows, ld_src, noise, ld_noise, sigma, post, rows, d, fixed_scale,
                 x_out, ldx, norms, reinterpret_cast<__half*>(hi), reinterpret_cast<__half*>(lo), ldh, inv_scale};
    const bool vec = d % 4 == 0 && ld_src % 4 == 0 && aligned16(src) &&
                     (!noise || (ld_noise % 4 == 0 && aligned16(noise))) &&
                     (!x_out || (ldx % 4 == 0 && aligned16(x_out))) &&
                     (!hi || (aligned16(hi) && aligned16(lo)));
    if (vec) prepare_rows_kernel<true><<<(unsigned)rows, 256, 0, as_stream(stream)>>>(p);
    else     prepare_rows_kernel<false><<<(unsigned)rows, 256, 0, as_stream(stream)>>>(p);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}

extern "C" int pdm_absmax_f32(const float* x, int64_t rows, int64_t d, int64_t ld, float* out, pdm_stream_t stream) {
    PDM_REQUIRE(x && out && rows > 0 && d > 0 && ld >= d, "pdm_absmax_f32: bad arguments");
    PDM_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(float), as_stream(stream)));
    const int64_t total = rows * d;
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(total, 256 * 8), 148 * 16);
    absmax_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, rows, d, ld, reinterpret_cast<unsigned*>(out));
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}

extern "C" int pdm_lattice_residual_f32(const float* y, int64_t n, int64_t d, int64_t ld, float scale, float* out2,
                                        pdm_stream_t stream) {
    PDM_REQUIRE(y && out2 && n > 0 && d > 0 && ld >= d && scale > 0.f, "pdm_lattice_residual_f32: bad arguments");
    PDM_CUDA_CHECK(cudaMemsetAsync(out2, 0, 2 * sizeof(float), as_stream(stream)));
    lattice_residual_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, as_stream(stream)>>>(y, n, d, ld, scale,
                                                                                      reinterpret_cast<unsigned*>(out2));
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}

extern "C" int pdm_transpose_split_f16(const float* y, int64_t n, int64_t d, int64_t ld, float scale,
                                       uint16_t* yt_hi, uint16_t* yt_lo, int64_t ldt, pdm_stream_t stream) {
    PDM_REQUIRE(y && yt_hi && yt_lo && n > 0 && d > 0 && ld >= d && ldt >= n && ldt % 8 == 0 && scale > 0.f,
                "pdm_transpose_split_f16: bad arguments");
    dim3 grid((unsigned)ceil_div(ldt, 32), (unsigned)ceil_div(d, 32)), block(32, 8);
    transpose_split_kernel<<<grid, block, 0, as_stream(stream)>>>(y, n, d, ld, scale, reinterpret_cast<__half*>(yt_hi),
                                                                 reinterpret_cast<__half*>(yt_lo), ldt);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}

extern "C" int pdm_column_moments_f32(const float* y, int64_t n, int64_t d, int64_t ld,
                                      double* sum, double* sumsq, float* minmax, pdm_stream_t stream) {
    PDM_REQUIRE(y && sum && sumsq && minmax && n > 0 && d > 0 && ld >= d, "pdm_column_moments_f32: bad arguments");
    moments_init_kernel<<<(unsigned)ceil_div(d, 256), 256, 0, as_stream(stream)>>>(sum, sumsq, d, minmax);
    const int64_t col_blocks = ceil_div(d, 32);
    const int64_t slabs = std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 64), ceil_div(148 * 8, col_blocks)));
    dim3 grid((unsigned)col_blocks, (unsigned)slabs), block(32, 8);
    column_moments_kernel<<<grid, block, 0, as_stream(stream)>>>(y, n, d, ld, sum, sumsq, minmax);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}

extern "C" int pdm_weights_from_energy_tiles(const float* energy, int64_t lde, int64_t M, int64_t N,
                                             const float* e_min, const float* l, const float* inv_temp,
                                             float* p_f32, int64_t ldp32,
                                             uint16_t* p_hi, uint16_t* p_lo, int64_t ldph,
                                             const int32_t* row_tiles, int32_t rows_per_tile, int64_t n_row_tiles,
                                             const int32_t* n_row_tiles_dev, pdm_stream_t stream) {
    PDM_REQUIRE(energy && e_min && l && inv_temp && M >= 0 && N > 0 && lde >= N, "pdm_weights_from_energy: bad arguments");
    PDM_REQUIRE(p_f32 || p_hi, "pdm_weights_from_energy: no output requested");
    PDM_REQUIRE((p_hi == nullptr) == (p_lo == nullptr), "pdm_weights_from_energy: p_hi and p_lo go together");
    PDM_REQUIRE(!p_hi || (ldph >= N && ldph % 8 == 0), "pdm_weights_from_energy: ldph must be >= N and a multiple of 8");
    PDM_REQUIRE(!p_f32 || ldp32 >= N, "pdm_weights_from_energy: ldp32 < N");
    PDM_REQUIRE(!row_tiles || (rows_per_tile > 0 && n_row_tiles >= 0), "pdm_weights_from_energy_tiles: bad tile list");
    PDM_REQUIRE(row_tiles || !n_row_tiles_dev, "pdm_weights_from_energy_tiles: a device-side count needs a tile list");
    const int64_t slots = row_tiles ? n_row_tiles * rows_per_tile : M;      // rows to visit (upper bound with a device count)
    if (M == 0 || slots == 0) return PDM_OK;
    PDM_REQUIRE(slots <= 65535 * 1024LL, "pdm_weights_from_energy: M too large for one launch");
    PDM_REQUIRE(N < (1ll << 31) - 16 && (!p_hi || ldph < (1ll << 31) - 16), "pdm_weights_from_energy: N must fit in int32");
    const int64_t width = p_hi ? ldph : N;
    const bool vec = lde % 4 == 0 && aligned16(energy) && (!p_hi || (aligned16(p_hi) && aligned16(p_lo))) &&
                     (!p_f32 || (ldp32 % 4 == 0 && aligned16(p_f32)));
    for (int64_t r0 = 0; r0 < slots; r0 += 65535) {
        const int64_t rows = std::min<int64_t>(65535, slots - r0);
        // vectorised path: a block takes 1024 groups of 8 columns per trip and walks its share of the row; rows are split
        // over several blocks only while there are too few rows to fill the device
        const int64_t per_row = vec ? std::max<int64_t>(1, std::min<int64_t>(ceil_div(width, 8 * 1024), ceil_div(148 * 16, rows)))
                                    : std::min<int64_t>(ceil_div(width, 256), 64);
        dim3 grid((unsigned)per_row, (unsigned)rows);
        auto kern = p_f32 ? weights_kernel<true> : weights_kernel<false>;
        kern<<<grid, 256, 0, as_stream(stream)>>>(
            energy, lde, M, N, e_min, l, inv_temp, p_f32, ldp32, reinterpret_cast<__half*>(p_hi), reinterpret_cast<__half*>(p_lo),
            ldph, vec ? 1 : 0, row_tiles, rows_per_tile, n_row_tiles_dev, r0);
        PDM_CUDA_CHECK(cudaGetLastError());
    }
    return PDM_OK;
}

extern "C" int pdm_weights_from_energy(const float* energy, int64_t lde, int64_t M, int64_t N,
                                       const float* e_min, const float* l, const float* inv_temp,
                                       float* p_f32, int64_t ldp32,
                                       uint16_t* p_hi, uint16_t* p_lo, int64_t ldph, pdm_stream_t stream) {
    return pdm_weights_from_energy_tiles(energy, lde, M, N, e_min, l, inv_temp, p_f32, ldp32, p_hi, p_lo, ldph, nullptr, 0, 0,
                                         nullptr, stream);
}

extern "C" int pdm_denoiser_backward_weights(const float* energy, int64_t lde, const float* sdot, int64_t lds,
                                             int64_t M, int64_t N, const float* e_min, const float* l,
                                             const float* inv_temp, const float* s_scale, const float* a_in,
                                             float* w, int64_t ldw, float* sums, pdm_stream_t stream) {
    PDM_REQUIRE(energy && sdot && e_min && l && inv_temp && w && sums && M >= 0 && N > 0 && lde >= N && lds >= N && ldw >= N,
                "pdm_denoiser_backward_weights: bad arguments");
    if (M == 0) return PDM_OK;
    PDM_REQUIRE(M < (1ll << 31), "pdm_denoiser_backward_weights: M too large for one launch");
    denoiser_backward_weights_kernel<<<(unsigned)M, 256, 0, as_stream(stream)>>>(energy, lde, sdot, lds, N, e_min, l, inv_temp,
                                                                                s_scale, a_in, w, ldw, sums);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}

extern "C" int pdm_topk_smallest_f32(const float* x, int64_t ldx, int64_t rows, int64_t n, int32_t k,
                                     float* vals, int64_t* idx, pdm_stream_t stream) {
    PDM_REQUIRE(x && vals && idx && rows >= 0 && n > 0 && ldx >= n && k >= 1 && k <= 1024,
                "pdm_topk_smallest_f32: bad arguments (1 <= k <= 1024)");
    if (rows == 0) return PDM_OK;
    PDM_REQUIRE(rows < (1ll << 31), "pdm_topk_smallest_f32: too many rows for one launch");
    topk_smallest_kernel<<<(unsigned)rows, 256, 0, as_stream(stream)>>>(x, ldx, n, (int)k, vals, idx);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}

extern "C" int pdm_refine_neighbours_f32(const float* x, int64_t ldx, int64_t M, int64_t d, const float* y, int64_t ldy,
                                         int64_t n_local, int64_t index_offset, int32_t k, float* vals, int64_t* idx,
                                         pdm_stream_t stream) {
    PDM_REQUIRE(x && y && vals && idx && M >= 0 && d > 0 && ldx >= d && ldy >= d && n_local > 0 && k >= 1 && k <= 8,
                "pdm_refine_neighbours_f32: bad arguments (1 <= k <= 8)");
    if (M == 0) return PDM_OK;
    refine_neighbours_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, as_stream(stream)>>>(x, ldx, M, d, y, ldy, n_local, index_offset,
                                                                                     (int)k, vals, idx);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}

extern "C" int pdm_sampler_step_f32(const float* x0_hat, const float* xt, const float* noise, float c_x0, float c_xt,
                                    float c_noise, float* out, int64_t n, pdm_stream_t stream) {
    PDM_REQUIRE(x0_hat && xt && out && n >= 0, "pdm_sampler_step_f32: bad arguments");
    if (n == 0) return PDM_OK;
    const bool vec = aligned16(x0_hat) && aligned16(xt) && aligned16(out) && (!noise || aligned16(noise));
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n, 256 * 4), 148 * 16);
    sampler_step_kernel<<<grid, 256, 0, as_stream(stream)>>>(x0_hat, xt, noise, c_x0, c_xt, c_noise, out, n, vec ? 1 : 0, nullptr);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}

extern "C" int pdm_sampler_step_dev_f32(const float* x0_hat, const float* xt, const float* noise, const float* coef,
                                        float* out, int64_t n, pdm_stream_t stream) {
    PDM_REQUIRE(x0_hat && xt && out && coef && n >= 0, "pdm_sampler_step_dev_f32: bad arguments");
    if (n == 0) return PDM_OK;
    const bool vec = aligned16(x0_hat) && aligned16(xt) && aligned16(out) && (!noise || aligned16(noise));
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n, 256 * 4), 148 * 16);
    sampler_step_kernel<<<grid, 256, 0, as_stream(stream)>>>(x0_hat, xt, noise, 0.f, 0.f, 0.f, out, n, vec ? 1 : 0, coef);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}
