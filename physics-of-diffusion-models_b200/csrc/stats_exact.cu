// Exact-fp32 CUDA-core variant of the fused pass (validation variant + small-d path) and the exact
// posterior-mean contraction.  Classic 128x128x16 register-tiled SGEMM (8x8 per thread) whose epilogue
// turns the Gram tile into energies in the reference's op order
//     E = 0.5 * ((||x||^2 - 2 x.y) + ||y||^2)            utils/distance.py:21, utils/stats.py:77
// parks them in shared memory and folds them row by row into the online Boltzmann state
// (online_stats.cuh).  Nothing of size M x N reaches HBM unless energy_out is requested.
#include "pdm_common.cuh"
#include "online_stats.cuh"

namespace pdm {

constexpr int XBM = 128, XBN = 128, XBK = 16;
constexpr int XES = 130;   // row stride of the energy tile: bank = (2*row + col) % 32 is conflict-free
                           // for the (row = lane/2, col = 2c + lane%2) mapping of the row pass

struct ExactParams {
    const float* q; int64_t ldq; const float* y; int64_t ldy;
    const float* q_norm; const float* y_norm; const float* inv_temp; const float* y_aux;
    int64_t M, N, d, index_offset;
    int n_splits;
    float* partials; float* energy_out; int64_t lde; float energy_mult;
};

// ---- tile loaders: 128 rows x 16 k, thread -> (row = tid & 127, k-half = tid >> 7), 8 consecutive k ----
__device__ __forceinline__ void load_rows_k8(const float* base, int64_t ld, int64_t row, int64_t nrows, int64_t k, int64_t kmax,
                                             bool vec_ok, float (&r)[8]) {
    if (row < nrows && vec_ok && k + 8 <= kmax) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(base + row * ld + k));
        const float4 b = __ldg(reinterpret_cast<const float4*>(base + row * ld + k + 4));
        r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = (row < nrows && k + i < kmax) ? __ldg(base + row * ld + k + i) : 0.f;
    }
}

__device__ __forceinline__ void store_rows_k8(float (*s)[XBM], int row, int kh, const float (&r)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s[kh * 8 + i][row] = r[i];
}

__device__ __forceinline__ void mma_tile_16(const float (*As)[XBM], const float (*Bs)[XBN], int ty, int tx, float (&acc)[8][8]) {
#pragma unroll
    for (int k = 0; k < XBK; ++k) {
        float a[8], b[8];
        *reinterpret_cast<float4*>(a)     = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
        *reinterpret_cast<float4*>(b)     = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
}

__device__ __forceinline__ int frag_index(int t, int i) { return (i < 4) ? t * 4 + i : 64 + t * 4 + (i - 4); }

template <bool kAux>
__global__ void __launch_bounds__(256, 2) exact_stats_kernel(ExactParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float (*As)[XBM] = reinterpret_cast<float (*)[XBM]>(smem_raw);
    float (*Bs)[XBN] = reinterpret_cast<float (*)[XBN]>(smem_raw + sizeof(float) * XBK * XBM);
    float* Es = reinterpret_cast<float*>(smem_raw + sizeof(float) * XBK * (XBM + XBN));
    float* yns = Es + XBM * XES;
    float* auxs = yns + XBN;

    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int lrow = tid & 127, lkh = tid >> 7;
    const int64_t m0 = (int64_t)blockIdx.x * XBM;
    const int split = blockIdx.y;
    const int64_t n_tiles = ceil_div(p.N, (int64_t)XBN);
    const bool qvec = (p.ldq % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.q) & 15) == 0);
    const bool yvec = (p.ldy % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0);

    float xn[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = m0 + frag_index(ty, i);
        xn[i] = (r < p.M) ? p.q_norm[r] : 0.f;
    }
    // row-pass ownership: thread -> (row = tid / 2, parity = tid & 1)
    const int prow = tid >> 1, par = tid & 1;
    const int64_t grow = m0 + prow;
    const float inv_t = (p.partials && grow < p.M) ? p.inv_temp[grow] : 1.f;
    RowState st;
    state_init(st);

    for (int64_t t = split; t < n_tiles; t += p.n_splits) {
        const int64_t n0 = t * XBN;
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

        float ra[8], rb[8];
        load_rows_k8(p.q, p.ldq, m0 + lrow, p.M, lkh * 8, p.d, qvec, ra);
        load_rows_k8(p.y, p.ldy, n0 + lrow, p.N, lkh * 8, p.d, yvec, rb);
        for (int64_t k0 = 0; k0 < p.d; k0 += XBK) {
            __syncthreads();                         // previous tile of As/Bs fully consumed
            store_rows_k8(As, lrow, lkh, ra);
            store_rows_k8(Bs, lrow, lkh, rb);
            if (k0 == 0 && tid < XBN) {
                const int64_t c = n0 + tid;
                yns[tid] = (c < p.N) ? p.y_norm[c] : 0.f;
                if (kAux) auxs[tid] = (c < p.N) ? p.y_aux[c] : 0.f;
            }
            __syncthreads();
            if (k0 + XBK < p.d) {
                load_rows_k8(p.q, p.ldq, m0 + lrow, p.M, k0 + XBK + lkh * 8, p.d, qvec, ra);
                load_rows_k8(p.y, p.ldy, n0 + lrow, p.N, k0 + XBK + lkh * 8, p.d, yvec, rb);
            }
            mma_tile_16(As, Bs, ty, tx, acc);
        }

        // energies in the reference's op order; masked columns get a huge energy (weight 0)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = frag_index(ty, i);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = frag_index(tx, j);
                const float dist = __fadd_rn(__fsub_rn(xn[i], 2.f * acc[i][j]), yns[c]);
                Es[r * XES + c] = (n0 + c < p.N) ? 0.5f * dist : kBigE;
            }
        }
        __syncthreads();

        if (p.energy_out) {
            for (int idx = tid; idx < XBM * XBN; idx += 256) {
                const int r = idx >> 7, c = idx & 127;
                if (m0 + r < p.M && n0 + c < p.N)
                    p.energy_out[(m0 + r) * p.lde + n0 + c] = p.energy_mult * Es[r * XES + c];
            }
        }
        if (p.partials) {
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) {
                float e[16], ax[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const int col = 2 * (c0 + c) + par;
                    e[c] = Es[prow * XES + col];
                    ax[c] = kAux ? auxs[col] : 0.f;
                }
                state_add_chunk<16, kAux>(st, e, ax, p.index_offset + n0 + 2 * c0 + par, 2, inv_t);
            }
        }
        // the next tile's first __syncthreads orders these reads before Es is overwritten
    }

    if (p.partials) {
        RowState o;
        o.m = __shfl_xor_sync(0xffffffffu, st.m, 1);
        o.l = __shfl_xor_sync(0xffffffffu, st.l, 1);
        o.a1 = __shfl_xor_sync(0xffffffffu, st.a1, 1);
        o.a2 = __shfl_xor_sync(0xffffffffu, st.a2, 1);
        o.aux = __shfl_xor_sync(0xffffffffu, st.aux, 1);
        o.idx = __shfl_xor_sync(0xffffffffu, st.idx, 1);
        if (par == 0 && grow < p.M) {
            state_merge(st, o, inv_t);
            state_store(st, p.partials + ((int64_t)split * p.M + grow) * PDM_PART_STRIDE);   // record-major
        }
    }
}

// ------------------------------------------------------------------------------------------------
// out (M, d) [+]= P (M, N) @ Y (N, d)   -- exact fp32, same register tiling
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) exact_mean_kernel(const float* __restrict__ P, int64_t ldp, int64_t M, int64_t N,
                                                            const float* __restrict__ Y, int64_t ldy, int64_t d,
                                                            float* __restrict__ out, int64_t ldo, int accumulate) {
    __shared__ __align__(16) float As[XBK][XBM];
    __shared__ __align__(16) float Bs[XBK][XBN];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int lrow = tid & 127, lkh = tid >> 7;
    const int64_t m0 = (int64_t)blockIdx.y * XBM, c0 = (int64_t)blockIdx.x * XBN;
    const bool pvec = (ldp % 4 == 0) && ((reinterpret_cast<uintptr_t>(P) & 15) == 0);
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const int bk = tid >> 4, bc = (tid & 15) * 8;     // B tile: 16 k-rows x 128 columns
    for (int64_t k0 = 0; k0 < N; k0 += XBK) {
        float ra[8], rb[8];
        load_rows_k8(P, ldp, m0 + lrow, M, k0 + lkh * 8, N, pvec, ra);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            rb[i] = (k0 + bk < N && c0 + bc + i < d) ? __ldg(Y + (k0 + bk) * ldy + c0 + bc + i) : 0.f;
        __syncthreads();
        store_rows_k8(As, lrow, lkh, ra);
#pragma unroll
        for (int i = 0; i < 8; ++i) Bs[bk][bc + i] = rb[i];
        __syncthreads();
        mma_tile_16(As, Bs, ty, tx, acc);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = m0 + frag_index(ty, i);
        if (r >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t c = c0 + frag_index(tx, j);
            if (c < d) out[r * ldo + c] = accumulate ? out[r * ldo + c] + acc[i][j] : acc[i][j];
        }
    }
}

constexpr size_t kExactSmem = sizeof(float) * (XBK * (XBM + XBN) + XBM * XES + 2 * XBN);

int launch_exact_stats(const pdm_stats_args& a, cudaStream_t stream) {
    PDM_REQUIRE(a.q && a.y && a.ldq >= a.d && a.ldy >= a.d, "pdm_posterior_stats(exact): fp32 operands q/y missing or ld < d");
    ExactParams p{a.q, a.ldq, a.y, a.ldy, a.q_norm, a.y_norm, a.inv_temp, a.y_aux, a.M, a.N, a.d, a.index_offset,
                  a.n_splits, a.partials, a.energy_out, a.lde, a.energy_mult};
    auto kern = a.y_aux ? exact_stats_kernel<true> : exact_stats_kernel<false>;
    PDM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kExactSmem));
    dim3 grid((unsigned)ceil_div(a.M, (int64_t)XBM), (unsigned)a.n_splits);
    kern<<<grid, 256, kExactSmem, stream>>>(p);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}

}  // namespace pdm

using namespace pdm;

extern "C" int pdm_weighted_mean_exact_f32(const float* p, int64_t ldp, int64_t M, int64_t N,
                                           const float* y, int64_t ldy, int64_t d,
                                           float* out, int64_t ldo, int32_t accumulate, pdm_stream_t stream) {
    PDM_REQUIRE(p && y && out && M >= 0 && N > 0 && d > 0 && ldp >= N && ldy >= d && ldo >= d,
                "pdm_weighted_mean_exact_f32: bad arguments");
    if (M == 0) return PDM_OK;
    dim3 grid((unsigned)ceil_div(d, (int64_t)XBN), (unsigned)ceil_div(M, (int64_t)XBM));
    PDM_REQUIRE(grid.y <= 65535, "pdm_weighted_mean_exact_f32: M too large for one launch");
    exact_mean_kernel<<<grid, 256, 0, as_stream(stream)>>>(p, ldp, M, N, y, ldy, d, out, ldo, accumulate);
    PDM_CUDA_CHECK(cudaGetLastError());
    return PDM_OK;
}
