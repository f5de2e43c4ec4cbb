// Stand-alone C++ host for the C ABI (no Python, no torch): cudaMalloc'd buffers, the planning call, the fused
// statistics pass on tensor cores and on the exact path, the merge -- checked against a double-precision CPU loop.
// Built and run by tests/test_gpu_dropin.py::test_c_abi_without_python (nvcc abi_smoke.cu -lpdm_b200).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "pdm_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %s\n", cudaGetErrorString(e_), #x); return 2; } } while (0)
#define PK(x) do { int r_ = (x); if (r_ != PDM_OK) { printf("pdm error %d at %s: %s\n", r_, #x, pdm_last_error()); return 3; } } while (0)

static float frand(uint64_t& s) { s = s * 6364136223846793005ull + 1442695040888963407ull; return (float)((s >> 33) & 0xFFFFFF) / 8388608.f - 1.f; }

int main() {
    const int64_t N = 1500, d = 320, M = 300;
    int sm = 0, maj = 0, mn = 0;
    PK(pdm_device_info(0, &sm, &maj, &mn));
    printf("device: %d SMs, sm_%d%d, ABI %d\n", sm, maj, mn, pdm_abi_version());
    std::vector<float> y(N * d), x(M * d), temp(M);
    uint64_t seed = 12345;
    for (auto& v : y) v = frand(seed);
    for (int64_t b = 0; b < M; ++b) {
        temp[b] = powf(10.f, -3.f + 6.f * b / (M - 1));
        const int64_t j = (b * 7) % N;
        for (int64_t k = 0; k < d; ++k) x[b * d + k] = y[j * d + k] + sqrtf(temp[b]) * 0.5f * frand(seed);
    }
    std::vector<float> inv_t(M);
    for (int64_t b = 0; b < M; ++b) inv_t[b] = 1.f / temp[b];

    float *dy, *dx, *dyn, *dxn, *dit, *dqinv, *dyinv, *dabs, *dout, *dparts;
    uint16_t *yh, *yl, *qh, *ql;
    int64_t* dargmin;
    CK(cudaMalloc(&dy, N * d * 4)); CK(cudaMalloc(&dx, M * d * 4)); CK(cudaMalloc(&dyn, N * 4)); CK(cudaMalloc(&dxn, M * 4));
    CK(cudaMalloc(&dit, M * 4)); CK(cudaMalloc(&dqinv, M * 4)); CK(cudaMalloc(&dyinv, N * 4)); CK(cudaMalloc(&dabs, 4));
    CK(cudaMalloc(&yh, N * d * 2)); CK(cudaMalloc(&yl, N * d * 2)); CK(cudaMalloc(&qh, M * d * 2)); CK(cudaMalloc(&ql, M * d * 2));
    CK(cudaMalloc(&dout, PDM_OUT_ROWS * M * 4)); CK(cudaMalloc(&dargmin, M * 8));
    CK(cudaMemcpy(dy, y.data(), N * d * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dx, x.data(), M * d * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dit, inv_t.data(), M * 4, cudaMemcpyHostToDevice));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));

    // dataset: norms, global power-of-two scale, operand split; queries: norms + per-row split
    PK(pdm_row_norms_f32(dy, N, d, d, dyn, st));
    PK(pdm_absmax_f32(dy, N, d, d, dabs, st));
    float amax = 0.f;
    CK(cudaMemcpyAsync(&amax, dabs, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    int e = 0;
    frexpf(amax, &e);
    const float yscale = ldexpf(1.f, 12 - e);
    PK(pdm_prepare_rows(dy, N, d, nullptr, 0, nullptr, nullptr, N, d, yscale, nullptr, 0, nullptr, yh, yl, d, dyinv, st));
    PK(pdm_prepare_rows(dx, M, d, nullptr, 0, nullptr, nullptr, M, d, 0.f, nullptr, 0, dxn, qh, ql, d, dqinv, st));

    // CPU reference in double
    std::vector<double> ref_ent(M), ref_emin(M);
    std::vector<int64_t> ref_arg(M);
    for (int64_t b = 0; b < M; ++b) {
        std::vector<double> E(N);
        double m = 1e300;
        for (int64_t j = 0; j < N; ++j) {
            double s = 0;
            for (int64_t k = 0; k < d; ++k) { const double df = (double)x[b * d + k] - y[j * d + k]; s += df * df; }
            E[j] = 0.5 * s;
            if (E[j] < m) { m = E[j]; ref_arg[b] = j; }
        }
        double l = 0, a1 = 0;
        for (int64_t j = 0; j < N; ++j) { const double ee = (E[j] - m) / temp[b]; const double w = exp(-ee); l += w; a1 += w * ee; }
        ref_ent[b] = log(l) + a1 / l - log((double)N);
        ref_emin[b] = m;
    }

    int bad = 0;
    for (int pass = 0; pass < 2; ++pass) {
        pdm_stats_args a = {};
        a.precision = pass == 0 ? PDM_PREC_F16X3 : PDM_PREC_EXACT_F32;
        a.M = M; a.N = N; a.d = d;
        a.q = dx; a.ldq = d; a.y = dy; a.ldy = d;
        a.q_hi = qh; a.q_lo = ql; a.ldqh = d; a.q_inv_scale = dqinv;
        a.y_hi = yh; a.y_lo = yl; a.ldyh = d; a.y_inv_scale = 1.f / yscale;
        a.q_norm = dxn; a.y_norm = dyn; a.inv_temp = dit;
        int64_t nfloats = 0;
        PK(pdm_posterior_stats_plan(&a, 0, &nfloats));
        CK(cudaMalloc(&dparts, nfloats * 4));
        a.partials = dparts;
        PK(pdm_posterior_stats(&a, st));
        PK(pdm_merge_partials(dparts, M, 1, 0, a.records_per_row, M * PDM_PART_STRIDE, PDM_PART_STRIDE, dit, N, dout, dargmin, st));
        std::vector<float> out(PDM_OUT_ROWS * M);
        std::vector<int64_t> arg(M);
        CK(cudaMemcpyAsync(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(arg.data(), dargmin, M * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaFree(dparts));
        double worst = 0;
        for (int64_t b = 0; b < M; ++b) {
            double xn = 0;
            for (int64_t k = 0; k < d; ++k) xn += (double)x[b * d + k] * x[b * d + k];
            const double floor_e = 8 * ldexp(1.0, -24) * (xn + d) / temp[b];
            const double err = fabs(out[PDM_OUT_ENTROPY * M + b] - ref_ent[b]);
            const double tol = fmax(1e-4 * fabs(ref_ent[b]) + 2e-5, 2 * floor_e);
            if (err > tol) { ++bad; printf("row %lld: entropy %g vs %g (tol %g)\n", (long long)b, out[PDM_OUT_ENTROPY * M + b], ref_ent[b], tol); }
            if (fabs(out[PDM_OUT_E_MIN * M + b] - ref_emin[b]) > fmax(1e-4 * ref_emin[b] + 1e-5, floor_e * temp[b])) ++bad;
            if (arg[b] != ref_arg[b]) { ++bad; printf("row %lld: argmin %lld vs %lld\n", (long long)b, (long long)arg[b], (long long)ref_arg[b]); }
            worst = fmax(worst, err);
        }
        printf("%s: records/row %d, worst entropy error %.3e\n", pass == 0 ? "f16x3 (tcgen05)" : "exact fp32", a.records_per_row, worst);
    }
    // invalid arguments come back as status codes with a message, never as a crash
    if (pdm_row_norms_f32(nullptr, 1, 1, 1, nullptr, st) != PDM_ERR_INVALID_ARG || pdm_last_error()[0] == 0) { ++bad; printf("no error for null pointers\n"); }
    printf(bad ? "ABI SMOKE FAILED (%d)\n" : "ABI SMOKE OK\n", bad);
    return bad ? 1 : 0;
}
