/*
 * pdm_b200 -- C ABI of the B200-native empirical-denoiser / thermodynamic-statistics engine.
 *
 * This is the drop-in boundary for ONE hot path of antoniibelyshev/physics-of-diffusion-models:
 * squared distances to all N training points -> min-shifted log-sum-exp partition function ->
 * posterior energy moments -> softmax-weighted posterior mean.  The reference has no FFI layer for
 * this path (its boundary is the Python call signature); each entry point below
 * cites the reference code (file:line, relative to the reference checkout) whose arithmetic it
 * replaces.  The Python mirror of the reference interface
 * (physics-of-diffusion-models_b200/utils/{distance,stats,metric_utils}.py,
 * diffusion/scheduler/scheduler.py) binds these symbols with ctypes; see INTEGRATION.md.
 *
 * Conventions
 *   - extern "C", plain device pointers + int64 sizes, no torch types.  Every pointer is a DEVICE
 *     pointer unless its name ends in `_host`.  Matrices are row-major; `ld*` are leading dimensions
 *     in ELEMENTS.
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*; the caller passes torch's
 *     current stream) and is asynchronous.  No entry point allocates device memory: scratch and
 *     outputs are caller-provided (sizes from the *_plan call).
 *   - Return value: PDM_OK or a negative PDM_ERR_*; pdm_last_error() returns a thread-local message.
 *     Nothing throws.  There is NO CPU fallback: on a machine without an sm_100 device every compute
 *     entry point returns PDM_ERR_CUDA / PDM_ERR_UNSUPPORTED.
 *
 * Unified math.  For query row b (temperature T_b) and dataset row j:
 *     E_bj = 1/2 * ((||x_b||^2 - 2 x_b.y_j) + ||y_j||^2)          (reference op order, never clamped)
 *     m_b  = min_j E_bj,  e_bj = (E_bj - m_b)/T_b,  w_bj = exp(-e_bj)
 *     l = sum w,  A1 = sum w e,  A2 = sum w e^2,  AUX = sum w s_j,  O = sum w y_j
 * A partial record holds (m, l, A1, A2, AUX, argmin) for a subset of dataset rows; records combine with
 * the non-negative shift rule of SURVEY.md section 5 (pdm_merge_partials), which is also how dataset
 * shards on different GPUs are merged.
 */
#ifndef PDM_B200_H
#define PDM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDM_ABI_VERSION 4

#define PDM_OK               0
#define PDM_ERR_INVALID_ARG (-1)
#define PDM_ERR_UNSUPPORTED (-2)
#define PDM_ERR_CUDA        (-3)

/* precision modes of the distance contraction */
#define PDM_PREC_EXACT_F32  0   /* CUDA-core fp32 FMA (validation variant, small-d path)            */
#define PDM_PREC_F16X3      1   /* tcgen05 kind::f16, split operands hi*hi + lo*hi + hi*lo, fp32 acc */
#define PDM_PREC_F16X1      2   /* tcgen05 kind::f16, hi*hi only (11-bit operands; opt-in fast mode) */
#define PDM_PREC_F16X2      3   /* tcgen05 kind::f16, q_hi*y_hi + q_lo*y_hi: for datasets that ARE an fp16
                                   lattice (8-bit images: y*255 is an integer), y_lo == 0 and the third
                                   product of F16X3 vanishes identically -- same accuracy, 2/3 of the MMAs */

#define PDM_PREC_F8X1       4   /* tcgen05 kind::f8f6f4 on E4M3 operands (4-bit significands, twice the MMA rate): the first
                                   stage of the screening cascade only.  q_hi / y_hi point at E4M3 BYTES from
                                   pdm_split_to_e4m3 (ldqh / ldyh in bytes, multiples of 16); q_inv_scale[r] and y_inv_scale
                                   are the factors that turn the byte operands back into values                   */

/* floats per partial record: m, l, A1, A2, AUX, argmin_lo(bits), argmin_hi(bits), reserved */
#define PDM_PART_STRIDE 8

/* rows of the output of pdm_merge_partials (struct-of-arrays, each row has M floats) */
#define PDM_OUT_E_MIN    0   /* m_b                                   utils/stats.py:80,282        */
#define PDM_OUT_LOG_L    1   /* log sum_j exp(-e_bj)  (= logZ')       utils/stats.py:83,284        */
#define PDM_OUT_MEAN_E   2   /* <e> = A1/l                            utils/stats.py:87,288        */
#define PDM_OUT_MEAN_E2  3   /* <e^2> = A2/l                          utils/stats.py:88            */
#define PDM_OUT_VAR_E    4   /* max(<e^2> - <e>^2, 0)                 utils/stats.py:90            */
#define PDM_OUT_AUX_MEAN 5   /* sum_j p_j s_j                         utils/stats.py:101           */
#define PDM_OUT_ENTROPY  6   /* logZ' + <e> - log N                   utils/stats.py:289           */
#define PDM_OUT_L        7   /* l (the normaliser used by the posterior-mean pass)                 */
#define PDM_OUT_ROWS     8

typedef void* pdm_stream_t;

const char* pdm_last_error(void);
int  pdm_abi_version(void);
/* sm count and compute capability of `device`; PDM_ERR_CUDA when there is no usable device. */
int  pdm_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------
 * K1  row norms.  out[r] = sum_k x[r,k]^2 (accumulated in fp64, rounded once to fp32).
 * Replaces norm_sqr (utils/distance.py:9-10); dataset norms are computed ONCE and cached by the
 * caller instead of on every compute_pw_dist_sqr call (utils/distance.py:17-18).
 * ------------------------------------------------------------------------------------------- */
int pdm_row_norms_f32(const float* x, int64_t rows, int64_t d, int64_t ld, float* out, pdm_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Row preparation: noising, optional rescale, row norm, and the fp16 hi/lo operand split.
 *   v[r,k] = (noise ? noise[r,k]*sigma[r] + src[r % src_rows, k] : src[r,k]) * (post ? post[r] : 1)
 * `noise*sigma + src` is evaluated as a rounded multiply followed by a rounded add, the op order of
 * `torch.randn(...) * t.sqrt() + x0` (utils/stats.py:74, :273).  `post` carries 1/sqrt(alpha_bar) for
 * the VP form of the ideal denoiser (diffusion/scheduler/scheduler.py:63-64: ||x - a y||^2 =
 * a^2 ||x/a - y||^2, so the dataset is never rescaled or copied).
 * Outputs (each optional, NULL = skip):
 *   x_out[r,k]      fp32 v                                      (ldx)
 *   norms[r]        ||v_r||^2  (fp64 accumulate)
 *   hi/lo[r,k]      fp16 split of v*2^kr : hi = rn16(v*2^kr), lo = rn16(v*2^kr - hi); columns
 *                   d..ldh-1 are zero-filled.  ldh % 8 == 0.
 *   inv_scale[r]    2^-kr.  kr is chosen per row so that max|v_r|*2^kr is in [2^11, 2^12), unless
 *                   fixed_scale > 0, in which case 2^kr = fixed_scale for every row.
 * ------------------------------------------------------------------------------------------- */
int pdm_prepare_rows(const float* src, int64_t src_rows, int64_t ld_src,
                     const float* noise, int64_t ld_noise, const float* sigma, const float* post,
                     int64_t rows, int64_t d, float fixed_scale,
                     float* x_out, int64_t ldx, float* norms,
                     uint16_t* hi, uint16_t* lo, int64_t ldh, float* inv_scale,
                     pdm_stream_t stream);

/* Noised queries regenerated in-kernel from torch's CUDA Philox stream, straight into operands: for draw t
 * (= one `torch.randn(b, d)` of utils/stats.py:74, :273 at Philox offset `offset + t*offset_step`) and element (r, k)
 *     v = fl(fl(eps * sigma[t]) + x0[r, k])            row t*b + r of the outputs
 * with eps bit-identical to what torch.randn would have written (same thread <-> element map, same cuRAND
 * device functions; `draw_threads` = 256 * grid of that torch launch).  Outputs: x_out fp32 (optional) and/or the
 * fp16 hi/lo split with inv_scale, scaled by the per-row power of two given by the bound max|x0_r| + 6.8 sigma[t]
 * (x0 and the outputs contiguous: ld_x0 == ldx == ldh == d; d % 8 == 0 for the split).  Row norms of the split
 * operands: pdm_split_row_norms.  Replaces the torch.randn launch +
 * pdm_prepare_rows pair (12 B/element less HBM traffic); the host checks bit-identity against torch.randn once. */
int pdm_noised_rows_philox(uint64_t seed, uint64_t offset, uint64_t offset_step, int64_t draw_threads,
                           const float* x0, int64_t b, int64_t d, int64_t ld_x0,
                           const float* sigma, int64_t n_draws, const float* x0_absmax,
                           float* x_out, int64_t ldx,
                           uint16_t* hi, uint16_t* lo, int64_t ldh, float* inv_scale, pdm_stream_t stream);
/* norms[r] = ||(hi_r + lo_r) * inv_scale[r]||^2 (fp64 accumulate): the norm of exactly the vector the MMAs see. */
int pdm_split_row_norms(const uint16_t* hi, const uint16_t* lo, int64_t ldh, const float* inv_scale,
                        int64_t rows, int64_t d, float* norms, pdm_stream_t stream);
/* out[r] = max_k |x[r,k]|. */
int pdm_row_absmax_f32(const float* x, int64_t rows, int64_t d, int64_t ld, float* out, pdm_stream_t stream);

/* max_k |x[r,k]| over the whole matrix -> out[0] (used once per dataset to pick its global 2^k). */
int pdm_absmax_f32(const float* x, int64_t rows, int64_t d, int64_t ld, float* out, pdm_stream_t stream);

/* Lattice test of a dataset against a candidate scale (8-bit images: Normalize(0.5, 0.5)(ToTensor(p)) * 255
 * = 2p - 255 up to fp32 rounding, utils/data.py:43-52 of the reference):
 *   out2[0] = max_j ||y_j - rint(y_j*scale)/scale||^2 / ||y_j||^2,   out2[1] = max |rint(y*scale)|.
 * When out2[1] <= 2048 (fp16 holds the integers exactly) and out2[0] is below fp32 resolution, the lo part
 * of the split carries nothing and PDM_PREC_F16X2 is exact to the same accuracy as PDM_PREC_F16X3. */
int pdm_lattice_residual_f32(const float* y, int64_t n, int64_t d, int64_t ld, float scale, float* out2,
                             pdm_stream_t stream);

/* Transposed fp16 split of the dataset for the posterior-mean contraction:
 *   yt_hi/yt_lo[k, j] = split(y[j,k]*scale), shape (d, ldt), columns N..ldt-1 zero.  ldt % 8 == 0. */
int pdm_transpose_split_f16(const float* y, int64_t n, int64_t d, int64_t ld, float scale,
                            uint16_t* yt_hi, uint16_t* yt_lo, int64_t ldt, pdm_stream_t stream);

/* K11  column moments of the dataset: sum[k] = sum_j y[j,k], sumsq[k] = sum_j y[j,k]^2 (fp64),
 * minmax[0..1] = global min / max.  Replaces the per-batch torch.var / min / max of
 * utils/stats.py:39,66,176 (Tr Sigma_0 = sum_k (sumsq_k - sum_k^2/N)/(N-1)). */
int pdm_column_moments_f32(const float* y, int64_t n, int64_t d, int64_t ld,
                           double* sum, double* sumsq, float* minmax, pdm_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K2-K7 fused pass: distances + online min / log-sum-exp / energy moments.  The M x N distance
 * matrix is never written unless `energy_out` is given.
 * Replaces, for ALL temperatures of a schedule flattened into the M rows of one launch:
 *   compute_pw_dist_sqr              utils/distance.py:13-21
 *   compute_stats_batch inner loop   utils/stats.py:271-289
 *   compute_metric_stats_batch loop  utils/stats.py:71-101
 *   true_posterior_mean_x0 weights   diffusion/scheduler/scheduler.py:64-68
 * ------------------------------------------------------------------------------------------- */
typedef struct pdm_stats_args {
    int32_t precision;          /* PDM_PREC_*                                                      */
    int32_t n_splits;           /* S >= 1: the N range is cut into S interleaved tile sets, each
                                   emitting its own partial record (0 = auto, see pdm_stats_plan)  */
    int32_t m_group;            /* tensor path: M tiles scheduled side by side (0 = auto)          */
    int32_t cta_group;          /* tensor path: 1 or 2 CTAs per MMA (0 = auto = 2)                 */
    int32_t records_per_row;    /* OUT of pdm_posterior_stats_plan: partial records emitted per query
                                   row (S on the exact path, 2*S on the tensor path: one per column
                                   half of a tile)                                                 */
    int32_t reserved0;
    int64_t M, N, d;
    int64_t index_offset;       /* global dataset index of local row 0 (dataset shards)            */
    /* fp32 operands (PDM_PREC_EXACT_F32) */
    const float* q;  int64_t ldq;        /* (M, d) */
    const float* y;  int64_t ldy;        /* (N, d) */
    /* fp16 split operands (tensor path), from pdm_prepare_rows */
    const uint16_t* q_hi; const uint16_t* q_lo; int64_t ldqh;  const float* q_inv_scale;  /* (M) */
    const uint16_t* y_hi; const uint16_t* y_lo; int64_t ldyh;  float y_inv_scale;
    /* per-row / per-column scalars */
    const float* q_norm;        /* (M) ||x||^2                                                     */
    const float* y_norm;        /* (N) ||y||^2                                                     */
    const float* inv_temp;      /* (M) 1/T                                                         */
    const float* y_aux;         /* (N) per-point scalar s_j (utils/stats.py:101) or NULL           */
    /* outputs */
    float*   partials;          /* (records_per_row, M, PDM_PART_STRIDE): record-major, so that both the
                                   kernel's stores and the merge's loads are coalesced over rows   */
    float*   energy_out;        /* optional (M, lde): energy_mult * E  (2.0 gives the squared
                                   distance of utils/distance.py:21, 1.0 the energy)               */
    int64_t  lde;
    float    energy_mult;
    /* screened launches (tensor path only): process only the listed row tiles of 128*cta_group rows
       (tile t = rows [t*128*cta_group, (t+1)*128*cta_group)); records of other rows are left untouched.
       NULL = every row.  pdm_posterior_stats_plan sizes the schedule for n_row_tiles tiles. */
    const int32_t* row_tiles;   /* (n_row_tiles) device, or NULL                                   */
    int64_t  n_row_tiles;
    /* ABI v4: optional DEVICE-side length of row_tiles (one int32): the launch covers the first
       min(n_row_tiles, *n_row_tiles_dev) tiles of the list, so a caller can chain screening pass -> tile list -> full
       pass without reading the count back (no host synchronisation inside a sampling step).  n_row_tiles is then the
       upper bound the schedule is planned for; a count of 0 makes the launch a no-op.                               */
    const int32_t* n_row_tiles_dev;
    /* ABI v4: top-k epilogue (tensor path, no partials / aux): instead of the statistics, every (row, record) keeps the
       PDM_TOPK_SLOTS smallest SQUARED DISTANCES ||x - y_j||^2 (op order of utils/distance.py:21) of its share of the dataset
       in registers and writes them once: topk_val (records_per_row, M, PDM_TOPK_SLOTS) ascending, +inf = empty slot,
       topk_idx the LOCAL dataset row of each (-1 = empty).  pdm_topk_merge combines the records.  Replaces the dense
       N x chunk distance tile + pdm_topk_smallest_f32 of the k-NN searches (utils/stats.py:50-60, 137-146;
       scripts/analyze_cifar_nn.py:37-47): nothing of size M x N reaches HBM.                                         */
    float*   topk_val;
    int32_t* topk_idx;
} pdm_stats_args;

#define PDM_TOPK_SLOTS 8

/* Fills args->n_splits / m_group / cta_group when they are 0, sets args->records_per_row and reports the
 * size of `partials` in floats. */
int pdm_posterior_stats_plan(pdm_stats_args* args, int device, int64_t* partial_floats);
int pdm_posterior_stats(const pdm_stats_args* args, pdm_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Merge of partial records (the only exchange step of the sharded path).
 * Record (outer o, row r, inner i) sits at parts + o*outer_stride + i*inner_stride + r*row_stride floats:
 * `inner` = the records one launch emits per row (pdm_posterior_stats writes them record-major:
 * inner_stride = M*PDM_PART_STRIDE, row_stride = PDM_PART_STRIDE), `outer` = dataset shards (GPUs) after an
 * all-gather.  Writes out[PDM_OUT_ROWS][M] and argmin[M] (int64 global dataset index, first index on ties
 * like torch.min).  n_total = total dataset size N for the entropy's log N.
 * ------------------------------------------------------------------------------------------- */
int pdm_merge_partials(const float* parts, int64_t M, int64_t n_outer, int64_t outer_stride,
                       int64_t n_inner, int64_t inner_stride, int64_t row_stride, const float* inv_temp,
                       int64_t n_total, float* out, int64_t* argmin, pdm_stream_t stream);

/* Same combination without the finalisation: out_records (M, PDM_PART_STRIDE) holds ONE merged record
 * per row.  A rank reduces its own splits with this before the all-gather, so that 32 bytes per query
 * row cross NVLink instead of 32 bytes per (row, split). */
int pdm_reduce_partials(const float* parts, int64_t M, int64_t n_outer, int64_t outer_stride,
                        int64_t n_inner, int64_t inner_stride, int64_t row_stride, const float* inv_temp,
                        float* out_records, pdm_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Certified delta posteriors (adaptive precision; an optimisation the reference has no counterpart of:
 * it evaluates utils/stats.py:80-90, 282-289 for every pair at every temperature).
 * In the low-noise part of a schedule the posterior of a query is a delta on its nearest training point to
 * fp32 resolution.  One tensor-core product (PDM_PREC_F16X1) with a rigorous error bound is enough to PROVE
 * that for a row; proven rows take the closed form (E_min exact, l = 1, <e> = <e^2> = 0, entropy = -log N)
 * and only the row tiles holding an unproven row go through the full-precision pass (row_tiles above).
 *
 * Step 1, pdm_screen_temperatures: the screening pass runs at a fictitious temperature T' per row,
 *     1/T' = e_star / (g*T + 2*delta),   delta = kappa * 2^-10 * ||x|| * max_j ||y_j|| + 2^-22 (||x||^2 + max_j ||y_j||^2)
 * (delta bounds |E1 - E| of the one-product energies: each operand is rounded to 11 significant bits; the second
 * term is the fp32 round-off of the norm expansion itself, which only matters for extreme norm ratios).
 * Step 2, pdm_screen_certify on the merged output of that pass: a row is certified when
 *     l' < 1.25   and   <e>' l' < 0.9 * e_star * exp(-e_star).
 * Every other point j then has (E1_j - E1_min)/T' > e_star  (e exp(-e) decreases for e > 1, and a point with
 * e < 1 would push l' above 1.36), i.e. a true gap E_j - E_min > g*T: its weight is below exp(-g), all of them
 * together below N exp(-g).  With g = 17 + log N that is under half an fp32 ulp of l = 1.  flags[r] = 1 for
 * certified rows; tile_list / n_tiles_out[0] = the ascending list of row tiles (rows_per_tile rows each)
 * that hold at least one uncertified row.
 * Step 3, pdm_screen_finalize, after the full pass over those tiles has been merged into out / argmin:
 * certified rows are overwritten with the closed form; E_min is recomputed for the arg-min of the
 * screening pass from the split operands (fp64 accumulation, then the fused pass's fp32 formula).
 * Row-sharded datasets: the screening records are merged across shards like any others, so every shard holds
 * the same flags; the shard that owns a certified row's arg-min writes E_min and the aux value, the others
 * write +inf / -inf, and the caller combines the E_min rows with MIN and the aux rows with MAX.
 * ------------------------------------------------------------------------------------------- */
int pdm_screen_temperatures(const float* q_norm, const float* inv_temp, int64_t M, const float* y_norm_max,
                            float g, float e_star, float kappa, float* inv_temp_screen, pdm_stream_t stream);
int pdm_screen_certify(const float* screen_out, int64_t M, float e_star, int32_t rows_per_tile,
                       uint8_t* flags, int32_t* tile_list, int32_t* n_tiles_out, pdm_stream_t stream);
/* First stage of the screening cascade: the same certificate from an E4M3 pass (PDM_PREC_F8X1, twice the MMA rate of
 * the fp16 one-product pass).  pdm_split_to_e4m3 rounds the split operands to bytes, out8 = e4m3((hi + lo)/16), and
 * returns the EXACT deviation of every row, err[r] = ||(hi + lo)_r - 16 out8_r|| in the split's scaled units (lo may be
 * NULL); with ex = err_q * q_inv_scale and ey = max_j err_y * y_inv_scale,
 *     |x.y - x8.y8| <= ex ||y|| + (||x|| + ex) ey  =: delta / kappa
 * holds rigorously, whatever the data.  pdm_screen_temperatures_f8 forms 1/T' = e_star / (g T + 2 delta) from it; the pass
 * itself is pdm_posterior_stats with q_hi / y_hi = the byte operands, q_inv_scale = 16 * the split's inverse scales and
 * y_inv_scale = 16 / scale.  Rows it leaves unproven go to the fp16 one-product stage (row_tiles), then to the full pass.
 * pdm_screen_tile_list rebuilds the tile list from a combined flag array. */
int pdm_split_to_e4m3(const uint16_t* hi, const uint16_t* lo, int64_t ldh, int64_t rows, int64_t d,
                      uint8_t* out8, int64_t ld8, float* err, pdm_stream_t stream);
int pdm_screen_temperatures_f8(const float* q_norm, const float* q_err, const float* q_inv_scale,
                               const float* inv_temp, int64_t M, const float* y_norm_max, const float* y_err_max,
                               float g, float e_star, float kappa, float* inv_temp_screen, pdm_stream_t stream);
int pdm_screen_tile_list(const uint8_t* flags, int64_t M, int32_t rows_per_tile, int32_t* tile_list,
                         int32_t* n_tiles_out, pdm_stream_t stream);
/* Second stage of a cascade run over the listed tiles of the first (tile_list, *n_tiles_dev <= max_tiles): rows the
 * first stage left open (flags[r] == 0) take the second stage's verdict and arg-min, in place. */
int pdm_screen_merge_stage(const int32_t* tile_list, const int32_t* n_tiles_dev, int64_t max_tiles, int32_t rows_per_tile,
                           int64_t M, const uint8_t* flags_b, const int64_t* arg_b, uint8_t* flags, int64_t* arg,
                           pdm_stream_t stream);
int pdm_screen_finalize(const uint8_t* flags, const int64_t* screen_argmin, int64_t M, int64_t d,
                        const uint16_t* q_hi, const uint16_t* q_lo, int64_t ldqh, const float* q_inv_scale,
                        const float* q_norm,
                        const uint16_t* y_hi, const uint16_t* y_lo, int64_t ldyh, float y_inv_scale,
                        const float* y_norm, const float* y_aux, int64_t index_offset, int64_t n_local,
                        int64_t n_total, float* out, int64_t* argmin, pdm_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K8 posterior mean  x0_hat_b = sum_j p_bj y_j,  p_bj = exp(-(E_bj - m_b)/T_b)/l_b.
 * Replaces p = exp(-h/(1-ab)); p /= p.sum; p @ data  (diffusion/scheduler/scheduler.py:66-69).
 * Step 1 turns a stored energy tile into normalised weights (fp16 hi/lo split, scaled by 2^14, or
 * fp32); step 2 contracts them with the dataset on tensor cores (f16x3) or CUDA cores (exact).
 * ------------------------------------------------------------------------------------------- */
int pdm_weights_from_energy(const float* energy, int64_t lde, int64_t M, int64_t N,
                            const float* e_min, const float* l, const float* inv_temp,
                            float* p_f32, int64_t ldp32,
                            uint16_t* p_hi, uint16_t* p_lo, int64_t ldph, pdm_stream_t stream);

/* Same over a list of row tiles of `rows_per_tile` rows whose length may live on the device (see
 * pdm_stats_args.n_row_tiles_dev): rows outside the listed tiles are not touched. */
int pdm_weights_from_energy_tiles(const float* energy, int64_t lde, int64_t M, int64_t N,
                                  const float* e_min, const float* l, const float* inv_temp,
                                  float* p_f32, int64_t ldp32,
                                  uint16_t* p_hi, uint16_t* p_lo, int64_t ldph,
                                  const int32_t* row_tiles, int32_t rows_per_tile, int64_t n_row_tiles,
                                  const int32_t* n_row_tiles_dev, pdm_stream_t stream);

/* Delta posteriors (scheduler.py:66-69 at low noise): flags[r] = 1 when l[r] - 1 <= 2^-23, i.e. every weight but the
 * nearest point's sums to at most one ulp -- the posterior mean IS that point; tile_list / n_tiles_out[0] = the row tiles
 * holding a row that is not a delta (the only ones whose weights have to be formed and contracted). */
int pdm_delta_tile_list(const float* l, int64_t M, int32_t rows_per_tile, uint8_t* flags, int32_t* tile_list,
                        int32_t* n_tiles_out, pdm_stream_t stream);
/* out[r, :] = src[idx[r] - index_offset, :] for rows with flags[r] != 0 (flags NULL = every row) whose index lies in
 * [index_offset, index_offset + n_local); zeros for flagged rows owned by another shard (shard outputs are summed);
 * rows with flags[r] == 0 are left as they are. */
int pdm_gather_rows_f32(const float* src, int64_t lds, int64_t n_local, int64_t d, const int64_t* idx,
                        int64_t index_offset, const uint8_t* flags, int64_t M, float* out, int64_t ldo,
                        pdm_stream_t stream);

/* out (M, d) [+]= scale * (A_hi.B_hi^T + A_lo.B_hi^T + A_hi.B_lo^T), A (M, K) and B (d, K) fp16
 * K-major (lda, ldb multiples of 8).  For the posterior mean A = weights, B = transposed dataset,
 * K = N, scale = 2^-14 * y_inv_scale.  accumulate != 0 adds into `out`.  b_lo == NULL drops the third
 * product (lattice datasets, see pdm_lattice_residual_f32). */
int pdm_split_gemm_f16x3(const uint16_t* a_hi, const uint16_t* a_lo, int64_t lda, int64_t M,
                         const uint16_t* b_hi, const uint16_t* b_lo, int64_t ldb, int64_t d, int64_t K,
                         float scale, float* out, int64_t ldo, int32_t accumulate, int32_t cta_group,
                         pdm_stream_t stream);

/* Same contraction over a list of row tiles of 128*cta_group rows of A (length on the host, or on the device through
 * n_row_tiles_dev as in pdm_stats_args); output rows outside the listed tiles are not touched. */
int pdm_split_gemm_f16x3_tiles(const uint16_t* a_hi, const uint16_t* a_lo, int64_t lda, int64_t M,
                               const uint16_t* b_hi, const uint16_t* b_lo, int64_t ldb, int64_t d, int64_t K,
                               float scale, float* out, int64_t ldo, int32_t accumulate, int32_t cta_group,
                               const int32_t* row_tiles, int64_t n_row_tiles, const int32_t* n_row_tiles_dev,
                               pdm_stream_t stream);

/* Exact fp32 CUDA-core version: out (M, d) [+]= P (M, N) @ Y (N, d). */
int pdm_weighted_mean_exact_f32(const float* p, int64_t ldp, int64_t M, int64_t N,
                                const float* y, int64_t ldy, int64_t d,
                                float* out, int64_t ldo, int32_t accumulate, pdm_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * One reverse-diffusion update of the ideal-denoiser sampler (diffusion/ddpm_sampling.py:94-110 with the
 * DDPMPredictions algebra of diffusion/ddpm/ddpm.py:17-20 folded in):
 *     out = c_x0 * x0_hat + c_xt * xt + c_noise * noise        (noise may be NULL: DDIM, or the last DDPM step)
 * DDPM: c_x0 = sqrt(ab') beta / (1 - ab), c_xt = sqrt(alpha) (1 - ab') / (1 - ab), c_noise = sqrt((1 - ab')/(1 - ab) beta),
 *       alpha = ab / ab', beta = 1 - alpha;
 * DDIM: c_x0 = sqrt(ab') - sqrt(1 - ab') sqrt(ab) / sqrt(1 - ab),  c_xt = sqrt(1 - ab') / sqrt(1 - ab).
 * out may alias xt.
 * ------------------------------------------------------------------------------------------- */
int pdm_sampler_step_f32(const float* x0_hat, const float* xt, const float* noise, float c_x0, float c_xt,
                         float c_noise, float* out, int64_t n, pdm_stream_t stream);
/* Same with the three coefficients read from DEVICE memory (coef[0..2] = c_x0, c_xt, c_noise): a step captured once in a
 * CUDA graph is replayed for every noise level of a trajectory, the host only rewrites five device scalars in between. */
int pdm_sampler_step_dev_f32(const float* x0_hat, const float* xt, const float* noise, const float* coef,
                             float* out, int64_t n, pdm_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * k-NN selection on a dense distance tile (pdm_posterior_stats with energy_out, mult 2): the k smallest
 * entries of every row, ascending, ties by lower column index; vals (rows, k), idx (rows, k) int64 (-1 / +inf
 * when a row has fewer than k finite entries).  Replaces sklearn's NearestNeighbors.kneighbors on the CPU in
 * utils/stats.py:50-60, 137-146 (sigma_reg^2 = d_k^2 * scale / D, self included as neighbour 0) and the
 * min / scatter / min flow of scripts/analyze_cifar_nn.py:37-47.
 * ------------------------------------------------------------------------------------------- */
int pdm_topk_smallest_f32(const float* x, int64_t ldx, int64_t rows, int64_t n, int32_t k,
                          float* vals, int64_t* idx, pdm_stream_t stream);

/* Merge of the top-k epilogue's records (see pdm_stats_args.topk_val): for every row the k <= PDM_TOPK_SLOTS smallest
 * (value, global index = local + index_offset) pairs over all records, ascending, ties by lower index; out_val (M, k),
 * out_idx (M, k) int64 (+inf / -1 when the dataset holds fewer than k rows). */
int pdm_topk_merge(const float* topk_val, const int32_t* topk_idx, int64_t M, int64_t records, int64_t index_offset,
                   int32_t k, float* out_val, int64_t* out_idx, pdm_stream_t stream);

/* Exact re-evaluation of nearest-neighbour candidates: vals[r, q] = sum_k (x[r,k] - y[idx[r,q] - index_offset, k])^2 with
 * fp64 accumulation for every candidate that lies in this shard, then each row's k <= 8 candidates are re-sorted (value,
 * then index; empty slots idx < 0 last).  The selection itself runs on the norm expansion of utils/distance.py:21, whose fp32
 * round-off is 2^-24 (||x||^2 + ||y||^2); sklearn's kneighbors, which utils/stats.py:50-60, 138-146 calls, works in
 * float64 -- this step brings the k-NN regulariser sigma_reg^2 = d_k^2 * scale / D to that accuracy. */
int pdm_refine_neighbours_f32(const float* x, int64_t ldx, int64_t M, int64_t d, const float* y, int64_t ldy,
                              int64_t n_local, int64_t index_offset, int32_t k, float* vals, int64_t* idx,
                              pdm_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Backward pass of K8 (autograd through Scheduler.true_posterior_mean_x0, which
 * scripts/optimize_schedule.py:57-91,151 differentiates with respect to the noise schedule).
 * With x0_hat = sum_j p_j y_j, upstream gradient g, s_j = y_j.g and a = sum_j p_j s_j:
 *     d x0_hat / d q . g = Cov_p(s, y) / T   = sum_j p_j (s_j - a) y_j / T            (q = the VE query)
 *     d x0_hat / d T . g = Cov_p(s, e) / T   = sum_j p_j (s_j - a) e_j / T            (e_j = (E_j - m)/T)
 * This entry point turns the stored energy tile and S = g.Y^T (from pdm_split_gemm_f16x3 or
 * pdm_weighted_mean_exact_f32; s_scale[r] is an optional per-row factor on S) into the centred weights
 *     w[r,j] = p_rj (s_rj - a_r)   and   sums[r] = (a_r, sum_j w_rj e_rj);
 * the caller contracts w with the dataset like the forward weights.  (Centring first: the uncentred form
 * sum p s y - a x0_hat cancels to zero at low T and its round-off would be amplified by 1/T.)
 * Row-sharded datasets: a first call (a_in = NULL) yields the local a_r; the caller sums them over the shards
 * and calls again with a_in = the global a_r (e_min, l are the merged, global ones in both calls).
 * ------------------------------------------------------------------------------------------- */
int pdm_denoiser_backward_weights(const float* energy, int64_t lde, const float* sdot, int64_t lds,
                                  int64_t M, int64_t N, const float* e_min, const float* l,
                                  const float* inv_temp, const float* s_scale, const float* a_in,
                                  float* w, int64_t ldw, float* sums, pdm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PDM_B200_H */
