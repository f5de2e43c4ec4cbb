// tcgen05 / TMEM / TMA fused pass for sm_100a.
//
// One warp-specialised persistent kernel computes S = A.B^T on the 5th-generation tensor cores with the
// fp16 hi/lo operand split (TERMS kind::f16 MMAs per k-step: hi.hi + lo.hi [+ hi.lo], exact products, fp32
// accumulation in TMEM; TERMS = 2 when the dataset is an fp16-exact lattice and its lo part is empty) and consumes
// each accumulator tile straight from TMEM in one of two epilogues:
//   EPI_STATS  energies E = 0.5*((|x|^2 - 2 x.y) + |y|^2) -> online min / log-sum-exp / energy moments
//              per query row (online_stats.cuh); the M x N distance matrix never reaches HBM.
//              Replaces utils/distance.py:13-21 + utils/stats.py:71-101, 271-289 + scheduler.py:64-68.
//   EPI_STORE  out = scale * S, the posterior-mean contraction p @ data (scheduler.py:69).
//   EPI_TOPK   the kTopK smallest squared distances per query row and their dataset indices, kept in registers and
//              written once per (row, split, column half): k-NN searches (utils/stats.py:50-60, 137-146; the
//              min / scatter / min flow of scripts/analyze_cifar_nn.py:37-47) without a dense distance tile in HBM.
//
// Per CTA: warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA only with cta_group::2), warp 2 = TMEM
// allocator, warps 4-11 = epilogue (two warps per TMEM lane quarter; thread <-> TMEM lane <-> query row).
// Pipelines: smem full/empty (TMA <-> MMA, kStages deep) and TMEM full/empty (MMA <-> epilogue, 2 accumulator
// stages), all mbarriers.  The accumulator is flushed into fp32 registers (round-to-nearest adds) every flush_kb
// k-blocks: the tensor core's own accumulation truncates.  With cta_group::2 a CTA pair owns a 256-row x
// 256-column tile: each CTA loads its 128 A rows and half of the B rows, the leader issues M=256 MMAs, each CTA's
// epilogue drains its own TMEM half.
//
// Schedule: pair p -> (i = p % m_group, s = p / m_group).  Round r handles row super-tile r*m_group + i;
// within it the pair walks column tiles s, s + n_splits, ...  Pairs that share i re-use the same A tiles
// out of L2 while pairs that share s stream the same B tiles (the producers re-align every few column tiles,
// round_rendezvous, so that this sharing survives a long launch); partial records per (row, s) are merged by
// pdm_merge_partials (the same rule that merges dataset shards across GPUs).
#include "pdm_common.cuh"
#include "online_stats.cuh"
#include "sm100_ptx.cuh"

#include <atomic>
#include <mutex>
#include <stdlib.h>

namespace pdm {
namespace tc {

using namespace ptx;

constexpr int kRowsPerCta = 128;
constexpr int kBlockK = 64;                       // fp16 elements = one 128-byte swizzled row
constexpr int kTileBytes = kRowsPerCta * kBlockK * 2;
constexpr int kAccStages = 2;
constexpr int kEpiWarp0 = 4;                      // warps 0-3: producer / MMA / TMEM alloc / spare
constexpr int kEpiWarps = 8;                      // warps 4-11: two warps per TMEM lane quarter
constexpr int kThreads = 32 * (kEpiWarp0 + kEpiWarps);
constexpr int kRegsControl = 40;                  // setmaxnreg budgets: 128*40 + 256*232 == 384*168 (the launch allocation)
constexpr int kRegsEpilogue = 232;
#ifndef PDM_SPIN_WAIT
#define PDM_SPIN_WAIT 0                           // 1: MMA issuer and epilogue busy-poll their hand-off barriers
                                                  // (measured on B200: slower -- the pollers steal issue slots and power)
#endif
#if PDM_SPIN_WAIT
#define PDM_HANDOFF_WAIT(bar, parity, hint) mbar_wait_spin(bar, parity)
#else
#define PDM_HANDOFF_WAIT(bar, parity, hint) mbar_wait(bar, parity, hint)
#endif
#ifndef PDM_DRAIN_COLS
#define PDM_DRAIN_COLS 32                         // columns per tcgen05.ld of the accumulator drain (32 or 64)
#endif

enum { EPI_STATS = 0, EPI_STORE = 1, EPI_TOPK = 2 };
constexpr int kTopK = 8;                         // slots of the top-k epilogue (k <= 8: the reference's k-NN uses k + 1 = 6)

// Soft rendezvous of the TMA producers at the start of every round.  CTA groups that share a column split stream the
// same dataset tiles; they start a round together and drift apart by a tile or two over its ~50 column tiles, which
// L2 absorbs -- but nothing re-aligns them between rounds, and over a long launch the drift grows until every group
// fetches its own copy of the dataset from HBM (measured: 0.48 MB of HBM reads per query row at 57k rows, 0.73 MB at
// 172k).  Re-aligning once per round costs microseconds.  The wait is bounded (50 us): a group that is not resident
// yet only costs the others that much, never a deadlock.  One counter per round, zeroed by the host before the launch.
constexpr int kMaxSyncRounds = 16384;
__device__ unsigned g_round_sync[kMaxSyncRounds];

__device__ __forceinline__ void round_rendezvous(unsigned* counter, unsigned expected) {
    atomicAdd(counter, 1u);
    const uint64_t t0 = global_timer_ns();
    while (true) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        if (v >= expected || global_timer_ns() - t0 > 50000ull) break;
        __nanosleep(256);
    }
}

// Dev instrumentation (-DPDM_STALL_STATS): nanoseconds each role of every CTA spent blocked on its
// barriers: [0] producer on smem-empty, [1] MMA on smem-full, [2] MMA on TMEM-empty, [3] epilogue warp 4
// on TMEM-full, [4] kernel wall time.  Read + reset with pdm_debug_read_stalls.
#ifdef PDM_STALL_STATS
__device__ unsigned long long g_stall[160][8];
#define PDM_STALL_BEGIN() const uint64_t _st0 = global_timer_ns()
#define PDM_STALL_END(acc) (acc) += global_timer_ns() - _st0
#else
#define PDM_STALL_BEGIN() do {} while (0)
#define PDM_STALL_END(acc) do {} while (0)
#endif

struct GemmParams {
    int64_t M;            // rows of A (queries)
    int64_t ncols;        // rows of B (dataset rows for EPI_STATS, feature columns for EPI_STORE)
    int32_t num_kb;       // number of 64-element k blocks
    int32_t flush_kb;     // k blocks accumulated inside the tensor core before a flush to registers
    uint32_t wait_hint_ns; // suspend-time hint of mbarrier.try_wait
    uint64_t hint_a, hint_b;   // L2 eviction-priority hints of the A (query) and B (dataset) tile loads
    unsigned* round_sync;      // soft rendezvous of the TMA producers (nullptr = off), one counter per rendezvous
    int32_t sync_points;       // counters available
    int32_t sync_tiles;        // column tiles between two rendezvous
    int32_t m_tiles;      // row super-tiles of 128*CG rows at work (all of them, or the length of tile_list)
    const int32_t* tile_list;  // optional: the row super-tiles to process (screened launches); nullptr = 0..m_tiles-1
    const int32_t* m_tiles_dev; // optional device-side length of tile_list (<= m_tiles, which then is its upper bound)
    int32_t n_tiles;      // column tiles of kBlockN
    int32_t m_group, n_splits;
    // EPI_STATS
    const float* q_norm; const float* q_inv_scale; const float* inv_temp;
    const float* y_norm; const float* y_aux; float y_inv_scale; int64_t index_offset;
    float* partials; float* energy_out; int64_t lde; float energy_mult;
    // EPI_STORE
    float* out; int64_t ldo; float out_scale; int32_t accumulate;
    // EPI_TOPK: (records, M, kTopK) squared distances ascending (+inf = empty slot) and LOCAL dataset row indices
    float* topk_val; int32_t* topk_idx;
};

template <int CG, int TERMS>
struct Cfg {
    static constexpr int kBlockN = (CG == 2) ? 256 : 128;
    static constexpr int kColsPerThread = kBlockN / 2;          // two epilogue warps share a lane quarter
    // stage layout: [A_hi][A_lo if TERMS >= 2][B_hi][B_lo if TERMS == 3]
    static constexpr int kTilesPerStage = TERMS + 1;
    static constexpr int kTileBHi = (TERMS >= 2) ? 2 : 1;
    static constexpr int kStageBytes = kTilesPerStage * kTileBytes;
    static constexpr int kStages = (TERMS == 3) ? 3 : (TERMS == 2) ? 4 : 6;
    static constexpr uint32_t kTmemCols = kAccStages * kBlockN;
    static constexpr uint32_t kIdesc = make_idesc_f16(128 * CG, kBlockN);
    static constexpr int kNumBars = 2 * kStages + 2 * kAccStages;
    static constexpr size_t kSmemBytes = 1024 + (size_t)kStages * kStageBytes + 8 * kNumBars + 16;
};

// Register re-balancing between warpgroups (all four warps of a warpgroup execute it).
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// The tensor core adds each MMA's K=16 partial product into the fp32 TMEM accumulator with truncation
// (measured on B200: ~0.35 ulp of bias per MMA, 0.025 absolute on a d=3072 near-neighbour dot product --
// 50x the fp32 SGEMM noise of the reference).  So the accumulator only ever holds `flush_kb` k-blocks
// (12 MMAs at flush_kb = 1); the epilogue warps drain it into fp32 registers with round-to-nearest adds
// while the MMA warp fills the other accumulator stage, and the statistics are computed from the registers.
// F8: one-product mode on E4M3 operands (TERMS == 1 only): a k-block is 128 one-byte elements (the same 128-byte
// swizzled row, the same four K slices of 32 bytes), MMAs are kind::f8f6f4.  Used by the screening cascade.
template <int CG, int TERMS, int EPI, bool AUX, bool F8 = false>
__global__ void __launch_bounds__(kThreads, 1)
fused_gemm_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                  const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                  const GemmParams p) {
    using C = Cfg<CG, TERMS>;
    constexpr int kBlockN = C::kBlockN;
    constexpr int kStages = C::kStages;
    constexpr int CPT = C::kColsPerThread;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    const uint32_t smem_base = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * C::kStageBytes);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + kAccStages + s); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + C::kNumBars);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const int pair = blockIdx.x / CG;

    if (CG == 2) cluster_sync_all();          // both CTAs of the pair are resident before the paired TMEM alloc

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tm_a_hi);
        prefetch_tensormap(&tm_b_hi);
        if (TERMS >= 2) prefetch_tensormap(&tm_a_lo);
        if (TERMS == 3) prefetch_tensormap(&tm_b_lo);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), CG); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < kAccStages; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), CG * kEpiWarps); }
        fence_mbar_init();
    }
    __syncwarp();
    if (warp == 2) {
        tmem_alloc<CG>(smem_u32(const_cast<uint32_t*>(tmem_slot)), C::kTmemCols);
        tmem_relinquish<CG>();
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const bool active = pair < p.m_group * p.n_splits;
    const int mi = pair % p.m_group, sp = pair / p.m_group;
    // a tile list whose length only the device knows (sync-free callers): every role reads the same word once
    const int m_tiles = p.m_tiles_dev ? min(p.m_tiles, __ldg(p.m_tiles_dev)) : p.m_tiles;
    const int n_chunks = (p.num_kb + p.flush_kb - 1) / p.flush_kb;

    if (warp < kEpiWarp0) {
        setmaxnreg_dec<kRegsControl>();
        if (active && warp == 0 && lane == 0) {
            // ===================== TMA producer =====================
            int stage = 0; uint32_t phase = 0;
            unsigned long long st_empty = 0; (void)st_empty;
#ifdef PDM_STALL_STATS
            const uint64_t t_begin = global_timer_ns();
#endif
            for (int mt = mi; mt < m_tiles; mt += p.m_group) {
                const int tile = p.tile_list ? __ldg(p.tile_list + mt) : mt;
                const int32_t a_row = (tile * CG + (int)rank) * kRowsPerCta;
                const int round = mt / p.m_group;
                const int row_groups = min(p.m_group, m_tiles - round * p.m_group);     // row tiles at work this round
                const int per_round = (ceil_div(p.n_tiles, p.n_splits) + p.sync_tiles - 1) / max(1, p.sync_tiles);
                int j = 0;
                for (int nt = sp; nt < p.n_tiles; nt += p.n_splits, ++j) {
                    if (p.round_sync && j % p.sync_tiles == 0 && (round > 0 || j > 0)) {
                        const int point = round * per_round + j / p.sync_tiles;
                        // column splits that still have a tile at step j: s + j*S < n_tiles
                        const int col_groups = min(p.n_splits, p.n_tiles - j * p.n_splits);
                        if (point < p.sync_points) round_rendezvous(p.round_sync + point, (unsigned)(row_groups * col_groups * CG));
                    }
                    const int32_t b_row = nt * kBlockN + (int)rank * kRowsPerCta;
                    for (int kb = 0; kb < p.num_kb; ++kb) {
                        { PDM_STALL_BEGIN(); mbar_wait(empty_bar(stage), phase ^ 1u, p.wait_hint_ns); PDM_STALL_END(st_empty); }
                        const uint32_t dst = smem_base + (uint32_t)stage * C::kStageBytes;
                        const uint32_t fb = full_bar(stage);
                        const int32_t kc = kb * (F8 ? 2 * kBlockK : kBlockK);          // element coordinate of the k-block
                        if (CG == 1) {
                            mbar_arrive_expect_tx(fb, C::kStageBytes);
                            tma_load_2d(dst, &tm_a_hi, fb, kc, a_row, p.hint_a);
                            if (TERMS >= 2) tma_load_2d(dst + kTileBytes, &tm_a_lo, fb, kc, a_row, p.hint_a);
                            tma_load_2d(dst + C::kTileBHi * kTileBytes, &tm_b_hi, fb, kc, b_row, p.hint_b);
                            if (TERMS == 3) tma_load_2d(dst + 3 * kTileBytes, &tm_b_lo, fb, kc, b_row, p.hint_b);
                        } else {
                            if (leader) mbar_arrive_expect_tx(fb, 2u * C::kStageBytes);
                            else mbar_arrive_remote(fb, 0);
                            tma_load_2d_pair(dst, &tm_a_hi, fb, kc, a_row, p.hint_a);
                            if (TERMS >= 2) tma_load_2d_pair(dst + kTileBytes, &tm_a_lo, fb, kc, a_row, p.hint_a);
                            tma_load_2d_pair(dst + C::kTileBHi * kTileBytes, &tm_b_hi, fb, kc, b_row, p.hint_b);
                            if (TERMS == 3) tma_load_2d_pair(dst + 3 * kTileBytes, &tm_b_lo, fb, kc, b_row, p.hint_b);
                        }
                        if (++stage == kStages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
#ifdef PDM_STALL_STATS
            g_stall[blockIdx.x][0] = st_empty;
            g_stall[blockIdx.x][4] = global_timer_ns() - t_begin;
#endif
        } else if (active && warp == 1 && lane == 0 && leader) {
            // ===================== MMA issuer =====================
            int stage = 0; uint32_t phase = 0; uint32_t chunk_iter = 0;
            unsigned long long st_full = 0, st_tempty = 0; (void)st_full; (void)st_tempty;
            for (int mt = mi; mt < m_tiles; mt += p.m_group) {
                for (int nt = sp; nt < p.n_tiles; nt += p.n_splits) {
                    for (int kb0 = 0; kb0 < p.num_kb; kb0 += p.flush_kb, ++chunk_iter) {
                        const uint32_t as = chunk_iter & 1u, aphase = (chunk_iter >> 1) & 1u;
                        // epilogue has drained this accumulator
                        { PDM_STALL_BEGIN(); PDM_HANDOFF_WAIT(tempty_bar(as), aphase ^ 1u, p.wait_hint_ns); PDM_STALL_END(st_tempty); }
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + as * kBlockN;
                        const int kb1 = min(p.num_kb, kb0 + p.flush_kb);
                        for (int kb = kb0; kb < kb1; ++kb) {
                            { PDM_STALL_BEGIN(); PDM_HANDOFF_WAIT(full_bar(stage), phase, p.wait_hint_ns); PDM_STALL_END(st_full); }
                            tc_fence_after();
                            const uint32_t sb = smem_base + (uint32_t)stage * C::kStageBytes;
                            const uint64_t a_hi = make_smem_desc_sw128(sb);
                            const uint64_t a_lo = make_smem_desc_sw128(sb + kTileBytes);
                            const uint64_t b_hi = make_smem_desc_sw128(sb + C::kTileBHi * kTileBytes);
                            const uint64_t b_lo = make_smem_desc_sw128(sb + 3 * kTileBytes);
                            // Small cross terms first: while the accumulator only holds them (2^-11 of the
                            // final magnitude) the tensor core's truncation costs nothing; the four large
                            // hi*hi products then go in last (16 fp16 = 32 bytes = 2 descriptor units per k).
                            if (TERMS >= 2) {
#pragma unroll
                                for (int k = 0; k < kBlockK / 16; ++k) {
                                    const uint64_t ko = (uint64_t)(k * 2);
                                    umma_f16<CG>(d_tmem, a_lo + ko, b_hi + ko, C::kIdesc, (kb > kb0 || k > 0) ? 1u : 0u);
                                    if (TERMS == 3) umma_f16<CG>(d_tmem, a_hi + ko, b_lo + ko, C::kIdesc, 1u);
                                }
                            }
#pragma unroll
                            for (int k = 0; k < kBlockK / 16; ++k) {
                                const uint64_t ko = (uint64_t)(k * 2);
                                if (F8) umma_f8<CG>(d_tmem, a_hi + ko, b_hi + ko, C::kIdesc, (kb > kb0 || k > 0) ? 1u : 0u);
                                else umma_f16<CG>(d_tmem, a_hi + ko, b_hi + ko, C::kIdesc,
                                                  (TERMS >= 2 || kb > kb0 || k > 0) ? 1u : 0u);
                            }
                            umma_commit<CG>(empty_bar(stage));                 // smem slot reusable once these retire
                            if (kb == kb1 - 1) umma_commit<CG>(tfull_bar(as)); // chunk accumulator complete
                            if (++stage == kStages) { stage = 0; phase ^= 1u; }
                        }
                    }
                }
            }
#ifdef PDM_STALL_STATS
            g_stall[blockIdx.x][1] = st_full;
            g_stall[blockIdx.x][2] = st_tempty;
#endif
        }
    } else {
        setmaxnreg_inc<kRegsEpilogue>();
        if (active) {
            // ===================== epilogue (8 warps) =====================
            // Per k-block chunk: TMEM -> registers, round-to-nearest accumulation with packed FADD2.  Per tile:
            // u = (|x|^2 - 2 x.y) + |y|^2 and the online statistics, two columns per instruction
            // (online_stats.cuh, PackedState).
            constexpr int CH = 32;                                // columns per statistics chunk
            constexpr int NP = CH / 2;                            // float2 pairs per chunk
            const int quarter = warp & 3;                         // TMEM lane quarter this warp may read
            const int half = (warp - kEpiWarp0) >> 2;             // which half of the tile's columns
            const int row_in_cta = quarter * 32 + lane;
            const uint32_t t_lane = (uint32_t)(quarter * 32) << 16;
            uint32_t chunk_iter = 0;
            unsigned long long st_tfull = 0, st_stats = 0; (void)st_tfull; (void)st_stats;
            // 128-bit stores of the dense outputs need 16-byte aligned rows
            constexpr bool kDist = EPI == EPI_STATS || EPI == EPI_TOPK;     // the epilogue works on squared distances
            const bool out_vec = kDist
                ? (p.energy_out && (p.lde % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.energy_out) & 15) == 0))
                : ((p.ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0));
            // |y_k| <= 4096 * y_inv_scale by construction of the split, so |y|^2 <= d_pad * (4096 y_inv_scale)^2
            const float yn_bound = (float)(p.num_kb * (F8 ? 2 * kBlockK : kBlockK)) * (4096.f * p.y_inv_scale) * (4096.f * p.y_inv_scale);
            const float half_mult = 0.5f * p.energy_mult;
            for (int mt = mi; mt < m_tiles; mt += p.m_group) {
                const int tile = p.tile_list ? __ldg(p.tile_list + mt) : mt;
                const int64_t grow = (int64_t)(tile * CG + (int)rank) * kRowsPerCta + row_in_cta;
                const bool row_ok = grow < p.M;
                float xn = 0.f, neg2inv = 0.f, c2 = -0.5f * kLog2e;
                bool safe = true;          // (u - 2m) * c2 cannot overflow for this row: no per-element clamp
                PackedState st;
                float tk_v[kTopK];
                int tk_i[kTopK];
                if (EPI == EPI_TOPK) {
#pragma unroll
                    for (int q = 0; q < kTopK; ++q) { tk_v[q] = INFINITY; tk_i[q] = -1; }
                }
                if (kDist) {
                    if (row_ok) {
                        xn = p.q_norm[grow];
                        neg2inv = -2.f * p.q_inv_scale[grow] * p.y_inv_scale;
                        if (EPI == EPI_STATS && p.inv_temp) {
                            const float inv_t = p.inv_temp[grow];
                            c2 = -0.5f * kLog2e * inv_t;
                            safe = inv_t * (xn + yn_bound) < 1.0e29f;
                        }
                    }
                    if (EPI == EPI_STATS) packed_init(st);
                }
                for (int nt = sp; nt < p.n_tiles; nt += p.n_splits) {
                    float2 sums[CPT / 2];
#pragma unroll
                    for (int i = 0; i < CPT / 2; ++i) sums[i] = make_float2(0.f, 0.f);
                    for (int ch = 0; ch < n_chunks; ++ch, ++chunk_iter) {
                        const uint32_t as = chunk_iter & 1u, aphase = (chunk_iter >> 1) & 1u;
                        { PDM_STALL_BEGIN(); PDM_HANDOFF_WAIT(tfull_bar(as), aphase, p.wait_hint_ns); PDM_STALL_END(st_tfull); }
                        tc_fence_after();
                        const uint32_t taddr = tmem_base + t_lane + as * kBlockN + half * CPT;
#if PDM_DRAIN_COLS == 64
#pragma unroll
                        for (int c0 = 0; c0 < CPT; c0 += 64) {
                            uint32_t v[64];
                            tmem_ld_32x64(taddr + c0, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                sums[c0 / 2 + i] = __fadd2_rn(sums[c0 / 2 + i],
                                                              make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])));
                        }
#else
#pragma unroll
                        for (int c0 = 0; c0 < CPT; c0 += 32) {
                            uint32_t v[32];
                            tmem_ld_32x32(taddr + c0, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                sums[c0 / 2 + i] = __fadd2_rn(sums[c0 / 2 + i],
                                                              make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])));
                        }
#endif
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (CG == 1 || leader) mbar_arrive(tempty_bar(as));
                            else mbar_arrive_remote(tempty_bar(as), 0);
                        }
                    }
                    // ---- the tile's Gram entries are complete in registers ----
#ifdef PDM_STALL_STATS
                    const uint64_t t_stats0 = global_timer_ns();
#endif
                    const int64_t nbase = (int64_t)nt * kBlockN + half * CPT;
                    // |y|^2 of the chunk after the current one is loaded while the current one is processed
                    // (one chunk ahead only: the barrier below keeps ptxas from hoisting all of them)
                    float4 yn_next[CH / 4];
                    if (kDist && nbase + CH <= p.ncols) {
#pragma unroll
                        for (int i = 0; i < CH / 4; ++i) yn_next[i] = __ldg(reinterpret_cast<const float4*>(p.y_norm + nbase) + i);
                    }
#pragma unroll
                    for (int c0 = 0; c0 < CPT; c0 += CH) {
                        const int64_t col0 = nbase + c0;
                        float4 yn_cur[CH / 4];
                        if (kDist) {
#pragma unroll
                            for (int i = 0; i < CH / 4; ++i) yn_cur[i] = yn_next[i];
                            asm volatile("" ::: "memory");
                            if (c0 + CH < CPT && col0 + 2 * CH <= p.ncols) {
#pragma unroll
                                for (int i = 0; i < CH / 4; ++i)
                                    yn_next[i] = __ldg(reinterpret_cast<const float4*>(p.y_norm + col0 + CH) + i);
                            }
                        }
                        if (col0 < p.ncols) {
                            const bool full_chunk = col0 + CH <= p.ncols;
                            if (kDist) {
                                float2 u[NP], ax[NP];
                                if (full_chunk) {
                                    const float2 n2 = splat2(neg2inv), xv = splat2(xn);
#pragma unroll
                                    for (int i = 0; i < CH / 4; ++i) {
                                        const float4 y = yn_cur[i];
                                        u[2 * i] = __fadd2_rn(__ffma2_rn(sums[c0 / 2 + 2 * i], n2, xv), make_float2(y.x, y.y));
                                        u[2 * i + 1] = __fadd2_rn(__ffma2_rn(sums[c0 / 2 + 2 * i + 1], n2, xv), make_float2(y.z, y.w));
                                    }
                                    if (AUX) {
                                        const float4* ax4 = reinterpret_cast<const float4*>(p.y_aux + col0);
#pragma unroll
                                        for (int i = 0; i < CH / 4; ++i) {
                                            const float4 y = __ldg(ax4 + i);
                                            ax[2 * i] = make_float2(y.x, y.y);
                                            ax[2 * i + 1] = make_float2(y.z, y.w);
                                        }
                                    }
                                } else {
#pragma unroll
                                    for (int i = 0; i < NP; ++i) {
                                        const bool ok0 = col0 + 2 * i < p.ncols, ok1 = col0 + 2 * i + 1 < p.ncols;
                                        const float y0 = ok0 ? __ldg(p.y_norm + col0 + 2 * i) : 0.f;
                                        const float y1 = ok1 ? __ldg(p.y_norm + col0 + 2 * i + 1) : 0.f;
                                        if (AUX) ax[i] = make_float2(ok0 ? __ldg(p.y_aux + col0 + 2 * i) : 0.f,
                                                                     ok1 ? __ldg(p.y_aux + col0 + 2 * i + 1) : 0.f);
                                        u[i].x = ok0 ? __fadd_rn(fmaf(sums[c0 / 2 + i].x, neg2inv, xn), y0) : 2.f * kBigE;
                                        u[i].y = ok1 ? __fadd_rn(fmaf(sums[c0 / 2 + i].y, neg2inv, xn), y1) : 2.f * kBigE;
                                    }
                                }
                                if (p.energy_out && row_ok) {
                                    float* eo = p.energy_out + grow * p.lde + col0;
                                    if (full_chunk && out_vec) {
#pragma unroll
                                        for (int i = 0; i < CH / 4; ++i)
                                            reinterpret_cast<float4*>(eo)[i] =
                                                make_float4(half_mult * u[2 * i].x, half_mult * u[2 * i].y,
                                                            half_mult * u[2 * i + 1].x, half_mult * u[2 * i + 1].y);
                                    } else {
#pragma unroll
                                        for (int i = 0; i < NP; ++i) {
                                            if (full_chunk || col0 + 2 * i < p.ncols) eo[2 * i] = half_mult * u[i].x;
                                            if (full_chunk || col0 + 2 * i + 1 < p.ncols) eo[2 * i + 1] = half_mult * u[i].y;
                                        }
                                    }
                                }
                                if (EPI == EPI_STATS && p.partials) {
                                    if (full_chunk && safe) packed_add_chunk<NP, AUX, false>(st, u, ax, p.index_offset + col0, c2);
                                    else packed_add_chunk<NP, AUX, true>(st, u, ax, p.index_offset + col0, c2);
                                }
                                if (EPI == EPI_TOPK) {
                                    // Sorted insertion, rare after the first few tiles (k log(N/k) insertions per row in
                                    // expectation): one chunk-minimum test guards the 32 compares.  Columns arrive in
                                    // ascending order within a thread, so the strict '<' keeps the lower index on ties.
                                    float cmin = fminf(u[0].x, u[0].y);
#pragma unroll
                                    for (int i = 1; i < NP; ++i) cmin = fminf(fminf(cmin, u[i].x), u[i].y);
                                    if (cmin < tk_v[kTopK - 1]) {
#pragma unroll
                                        for (int i = 0; i < 2 * NP; ++i) {
                                            const float v = (i & 1) ? u[i / 2].y : u[i / 2].x;
                                            if (v < tk_v[kTopK - 1] && v < kBigE) {
                                                tk_v[kTopK - 1] = v;
                                                tk_i[kTopK - 1] = (int)(col0 + i);
#pragma unroll
                                                for (int q = kTopK - 1; q > 0; --q) {
                                                    if (tk_v[q] < tk_v[q - 1]) {
                                                        const float tv = tk_v[q]; tk_v[q] = tk_v[q - 1]; tk_v[q - 1] = tv;
                                                        const int ti = tk_i[q]; tk_i[q] = tk_i[q - 1]; tk_i[q - 1] = ti;
                                                    }
                                                }
                                            }
                                        }
                                    }
                                }
                            } else if (row_ok) {
                                float* o = p.out + grow * p.ldo + col0;
                                if (full_chunk && out_vec && !p.accumulate) {
#pragma unroll
                                    for (int i = 0; i < CH / 4; ++i) {
                                        const float2 a = sums[c0 / 2 + 2 * i], b = sums[c0 / 2 + 2 * i + 1];
                                        reinterpret_cast<float4*>(o)[i] =
                                            make_float4(p.out_scale * a.x, p.out_scale * a.y, p.out_scale * b.x, p.out_scale * b.y);
                                    }
                                } else {
#pragma unroll
                                    for (int i = 0; i < CH; ++i) {
                                        if (full_chunk || col0 + i < p.ncols) {
                                            const float2 a = sums[c0 / 2 + i / 2];
                                            const float r = p.out_scale * ((i & 1) ? a.y : a.x);
                                            o[i] = p.accumulate ? o[i] + r : r;
                                        }
                                    }
                                }
                            }
                        }
                    }
#ifdef PDM_STALL_STATS
                    st_stats += global_timer_ns() - t_stats0;
#endif
                }
                // one partial record per (row, split, column half)
                if (EPI == EPI_STATS && p.partials && row_ok)
                    packed_store(st, p.partials + ((int64_t)(2 * sp + half) * p.M + grow) * PDM_PART_STRIDE);   // record-major
                if (EPI == EPI_TOPK && row_ok) {
                    const int64_t rec = ((int64_t)(2 * sp + half) * p.M + grow) * kTopK;
                    float4* tv4 = reinterpret_cast<float4*>(p.topk_val + rec);
                    int4* ti4 = reinterpret_cast<int4*>(p.topk_idx + rec);
                    tv4[0] = make_float4(tk_v[0], tk_v[1], tk_v[2], tk_v[3]);
                    tv4[1] = make_float4(tk_v[4], tk_v[5], tk_v[6], tk_v[7]);
                    ti4[0] = make_int4(tk_i[0], tk_i[1], tk_i[2], tk_i[3]);
                    ti4[1] = make_int4(tk_i[4], tk_i[5], tk_i[6], tk_i[7]);
                }
            }
#ifdef PDM_STALL_STATS
            if (warp == kEpiWarp0 && lane == 0) { g_stall[blockIdx.x][3] = st_tfull; g_stall[blockIdx.x][5] = st_stats; }
#endif
        }
    }

    __syncwarp();                             // single-lane roles rejoin their warp before the aligned barrier
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 2) tmem_dealloc<CG>(tmem_base, C::kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode_fn(EncodeTiledFn* out) {
    static std::mutex mu;
    static EncodeTiledFn fn = nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        PDM_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !ptr) {
            set_error("cuTensorMapEncodeTiled is not available from the installed driver");
            return PDM_ERR_UNSUPPORTED;
        }
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    *out = fn;
    return PDM_OK;
}

// fp16 matrix (rows, k) with leading dimension ld -> box of 64 k-elements x 128 rows, 128B swizzle.
// Reads beyond (rows, k) are zero-filled by the TMA unit, so padding never has to be materialised.
static int make_tile_map(CUtensorMap* map, const uint16_t* base, int64_t rows, int64_t k, int64_t ld) {
    EncodeTiledFn fn;
    int rc = get_encode_fn(&fn);
    if (rc != PDM_OK) return rc;
    PDM_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && ld % 8 == 0 && ld >= k,
                "fp16 operand must be 16-byte aligned with ld %% 8 == 0 and ld >= k");
    cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)kRowsPerCta};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<uint16_t*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return PDM_ERR_CUDA; }
    return PDM_OK;
}

// E4M3 matrix (rows, k) of bytes with leading dimension ld -> box of 128 k-elements x 128 rows, 128B swizzle.
static int make_tile_map_u8(CUtensorMap* map, const uint8_t* base, int64_t rows, int64_t k, int64_t ld) {
    EncodeTiledFn fn;
    int rc = get_encode_fn(&fn);
    if (rc != PDM_OK) return rc;
    PDM_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && ld % 16 == 0 && ld >= k,
                "e4m3 operand must be 16-byte aligned with ld %% 16 == 0 and ld >= k");
    cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld};
    cuuint32_t box[2] = {(cuuint32_t)(2 * kBlockK), (cuuint32_t)kRowsPerCta};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (u8) failed with CUresult %d", (int)r); return PDM_ERR_CUDA; }
    return PDM_OK;
}

template <int CG, int TERMS, int EPI, bool AUX, bool F8 = false>
static int launch_variant(const CUtensorMap* maps, const GemmParams& p, int sm_count, cudaStream_t stream) {
    using C = Cfg<CG, TERMS>;
    auto kern = fused_gemm_kernel<CG, TERMS, EPI, AUX, F8>;
    static std::atomic<uint64_t> configured{0};          // one bit per device ordinal (function attributes are per device)
    int dev = 0;
    PDM_CUDA_CHECK(cudaGetDevice(&dev));
    if (!((configured.load(std::memory_order_relaxed) >> (dev & 63)) & 1ull)) {
        PDM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes));
        configured.fetch_or(1ull << (dev & 63), std::memory_order_relaxed);
    }
    const int pairs = sm_count / CG;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(pairs * CG));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = C::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PDM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], p));
    return PDM_OK;
}

template <int EPI, bool AUX>
static int dispatch(int cg, int terms, const CUtensorMap* maps, const GemmParams& p, int sm_count, cudaStream_t stream) {
    if (cg == 2 && terms == 3) return launch_variant<2, 3, EPI, AUX>(maps, p, sm_count, stream);
    if (cg == 1 && terms == 3) return launch_variant<1, 3, EPI, AUX>(maps, p, sm_count, stream);
    if (cg == 2 && terms == 1) return launch_variant<2, 1, EPI, AUX>(maps, p, sm_count, stream);
    if (cg == 1 && terms == 1) return launch_variant<1, 1, EPI, AUX>(maps, p, sm_count, stream);
    if (cg == 2 && terms == 2) return launch_variant<2, 2, EPI, AUX>(maps, p, sm_count, stream);
    if (cg == 1 && terms == 2) return launch_variant<1, 2, EPI, AUX>(maps, p, sm_count, stream);
    set_error("unsupported cta_group %d / terms %d", cg, terms);
    return PDM_ERR_INVALID_ARG;
}

// k-blocks accumulated inside the tensor core between flushes.  Default: at most 16 MMAs per flush
// (f16x3: every k-block = 12 MMAs; f16x2: every second = 16; measured on B200 both stay within ~4 ulp of the
// row norms, the reference's own fp32 error).  The one-product mode rounds its operands to 11 bits (2^-11 relative),
// so 32 MMAs per flush (~1e-6 relative truncation bias) cost it nothing; measured on the 172 032 x 50 000 block:
// 44.1 ms at 4 k-blocks per flush, 40.9 at 8, 40.2 at 16 (f16x3: 119.7).  PDM_FLUSH_KB overrides.
static int flush_kb_setting(int terms) {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PDM_FLUSH_KB");
        v = e ? atoi(e) : 0;
        if (v < 0) v = 0;
    }
    if (v > 0) return v;
    return terms == 3 ? 1 : terms == 2 ? 2 : 8;
}

// L2 eviction priority of the tile loads: PDM_HINT_A / PDM_HINT_B = normal | first | last
static uint64_t evict_hint_setting(const char* name, uint64_t dflt) {
    const char* e = getenv(name);
    if (!e) return dflt;
    if (e[0] == 'f') return ptx::kEvictFirst;
    if (e[0] == 'l') return ptx::kEvictLast;
    return ptx::kEvictNormal;
}

static uint32_t wait_hint_setting() {
    static long v = -1;
    if (v < 0) {
        const char* e = getenv("PDM_WAIT_HINT_NS");
        v = e ? atol(e) : 20000;
        if (v < 0) v = 0;
    }
    return (uint32_t)v;
}

static int require_sm100(DeviceInfo* info) {
    int rc = current_device_info(info);
    if (rc != PDM_OK) return rc;
    if (info->cc_major != 10) {
        set_error("the tensor path needs an sm_100 device (found sm_%d%d); there is no fallback", info->cc_major, info->cc_minor);
        return PDM_ERR_UNSUPPORTED;
    }
    return PDM_OK;
}

// Pick (m_group, n_splits) = (G, S).  Every CTA group walks ceil(m_tiles/G) * ceil(n_tiles/S) tiles: that product is
// the schedule's length (tile quantisation).  The shortest schedule whose G live A tiles fit a 50 MB L2 budget wins
// (ties: the larger G = fewer passes over the dataset); if nothing fits, the shortest overall.  With the producers
// re-aligned every few column tiles (round_rendezvous) the dataset tiles of a split are shared out of L2 for any G, so
// keeping all 74 groups busy beats minimising HBM passes: measured on the 172 032 x 50 000 block, (S,G) = (6,12)
// 119.9 ms, (9,8) 120.8, (5,14) 121.9, (4,16) 125.1 (gpurun_out/exp11.log).
void plan_schedule(int pairs, int64_t m_tiles, int64_t n_tiles, int64_t a_tile_bytes, int* m_group, int* n_splits) {
    const int64_t budget = 50ll << 20;
    const int64_t g_cap = std::max<int64_t>(1, budget / std::max<int64_t>(1, a_tile_bytes));
    int best_g = 0, best_s = 0;
    int64_t best_steps = 0;
    for (int pass = 0; pass < 2 && best_g == 0; ++pass) {         // pass 0: inside the L2 budget; pass 1: anything
        for (int s = 1; s <= pairs && s <= n_tiles; ++s) {
            for (int64_t g = std::min<int64_t>(pairs / s, m_tiles); g >= 1; --g) {
                if (pass == 0 && g > g_cap) continue;
                const int64_t steps = ceil_div(m_tiles, g) * ceil_div(n_tiles, (int64_t)s);
                if (best_g == 0 || steps < best_steps || (steps == best_steps && g > best_g)) {
                    best_g = (int)g; best_s = s; best_steps = steps;
                }
            }
        }
    }
    *m_group = best_g;
    *n_splits = best_s;
}

int launch_tensor_stats(const pdm_stats_args& a, cudaStream_t stream) {
    DeviceInfo info;
    int rc = require_sm100(&info);
    if (rc != PDM_OK) return rc;
    const int cg = a.cta_group;
    const int terms = a.precision == PDM_PREC_F16X3 ? 3 : a.precision == PDM_PREC_F16X2 ? 2 : 1;
    const bool f8 = a.precision == PDM_PREC_F8X1;
    PDM_REQUIRE(cg == 1 || cg == 2, "cta_group must be 1 or 2");
    PDM_REQUIRE(!f8 || !a.y_aux, "the e4m3 mode is a screening pass: no aux accumulator");
    PDM_REQUIRE(a.q_hi && a.y_hi && a.q_inv_scale && a.y_inv_scale > 0.f, "tensor path: q_hi / y_hi / scales missing");
    PDM_REQUIRE(terms < 2 || a.q_lo, "f16x2 / f16x3 need the lo part of the queries");
    PDM_REQUIRE(terms < 3 || a.y_lo, "f16x3 needs the lo part of the dataset");
    PDM_REQUIRE((reinterpret_cast<uintptr_t>(a.y_norm) & 15) == 0 && (!a.y_aux || (reinterpret_cast<uintptr_t>(a.y_aux) & 15) == 0),
                "y_norm / y_aux must be 16-byte aligned");
    PDM_REQUIRE(a.M < (1ll << 31) - 512 && a.N < (1ll << 31) - 512, "M and N must fit in int32 for the TMA coordinates");
    CUtensorMap maps[4];
    if (f8) {       // q_hi / y_hi carry E4M3 bytes, leading dimensions in bytes
        const uint8_t* q8 = reinterpret_cast<const uint8_t*>(a.q_hi);
        const uint8_t* y8 = reinterpret_cast<const uint8_t*>(a.y_hi);
        if ((rc = make_tile_map_u8(&maps[0], q8, a.M, a.d, a.ldqh)) != PDM_OK) return rc;
        if ((rc = make_tile_map_u8(&maps[2], y8, a.N, a.d, a.ldyh)) != PDM_OK) return rc;
        maps[1] = maps[0]; maps[3] = maps[2];
    } else {
    if ((rc = make_tile_map(&maps[0], a.q_hi, a.M, a.d, a.ldqh)) != PDM_OK) return rc;
    if ((rc = make_tile_map(&maps[1], terms >= 2 ? a.q_lo : a.q_hi, a.M, a.d, a.ldqh)) != PDM_OK) return rc;
    if ((rc = make_tile_map(&maps[2], a.y_hi, a.N, a.d, a.ldyh)) != PDM_OK) return rc;
    if ((rc = make_tile_map(&maps[3], terms == 3 ? a.y_lo : a.y_hi, a.N, a.d, a.ldyh)) != PDM_OK) return rc;
    }
    const int block_n = cg == 2 ? 256 : 128;
    GemmParams p = {};
    p.M = a.M; p.ncols = a.N;
    p.num_kb = (int32_t)ceil_div(a.d, f8 ? 2 * kBlockK : kBlockK);
    p.flush_kb = flush_kb_setting(terms);
    p.wait_hint_ns = wait_hint_setting();
    p.hint_a = evict_hint_setting("PDM_HINT_A", kEvictNormal);
    p.hint_b = evict_hint_setting("PDM_HINT_B", kEvictFirst);      // dataset tiles are dead once their split has passed
    {
        static const bool round_sync_on = !(getenv("PDM_ROUND_SYNC") && atoi(getenv("PDM_ROUND_SYNC")) == 0);
        const int64_t rounds = ceil_div(a.row_tiles ? a.n_row_tiles : ceil_div(a.M, (int64_t)kRowsPerCta * cg), (int64_t)a.m_group);
        static const int sync_tiles = getenv("PDM_SYNC_TILES") ? std::max(1, atoi(getenv("PDM_SYNC_TILES"))) : 8;
        const int64_t cols_per_round = ceil_div(ceil_div(a.N, (int64_t)(cg == 2 ? 256 : 128)), (int64_t)a.n_splits);
        const int64_t points = rounds * ceil_div(cols_per_round, (int64_t)sync_tiles);
        if (round_sync_on && points > 1) {
            void* sym = nullptr;
            PDM_CUDA_CHECK(cudaGetSymbolAddress(&sym, g_round_sync));
            p.sync_points = (int32_t)std::min<int64_t>(points, kMaxSyncRounds);
            p.sync_tiles = sync_tiles;
            PDM_CUDA_CHECK(cudaMemsetAsync(sym, 0, sizeof(unsigned) * p.sync_points, stream));
            p.round_sync = static_cast<unsigned*>(sym);
        }
    }
    p.m_tiles = (int32_t)ceil_div(a.M, (int64_t)kRowsPerCta * cg);
    if (a.row_tiles) {
        PDM_REQUIRE(a.n_row_tiles >= 1 && a.n_row_tiles <= p.m_tiles, "n_row_tiles must be between 1 and the number of row tiles");
        p.m_tiles = (int32_t)a.n_row_tiles;
        p.tile_list = a.row_tiles;
        p.m_tiles_dev = a.n_row_tiles_dev;
    }
    p.n_tiles = (int32_t)ceil_div(a.N, block_n);
    p.m_group = a.m_group; p.n_splits = a.n_splits;
    PDM_REQUIRE(p.m_group >= 1 && p.n_splits >= 1 && (int64_t)p.m_group * p.n_splits <= info.sm_count / cg,
                "m_group * n_splits must be between 1 and the number of CTA groups (%d)", info.sm_count / cg);
    p.q_norm = a.q_norm; p.q_inv_scale = a.q_inv_scale; p.inv_temp = a.inv_temp;
    p.y_norm = a.y_norm; p.y_aux = a.y_aux; p.y_inv_scale = a.y_inv_scale; p.index_offset = a.index_offset;
    p.partials = a.partials; p.energy_out = a.energy_out; p.lde = a.lde; p.energy_mult = a.energy_mult;
    if (a.topk_val) {
        PDM_REQUIRE(a.topk_idx && !f8 && terms >= 2 && !a.y_aux && !a.partials,
                    "the top-k epilogue runs in f16x3 / f16x2, without partial records and without aux");
        PDM_REQUIRE((reinterpret_cast<uintptr_t>(a.topk_val) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.topk_idx) & 15) == 0,
                    "topk_val / topk_idx must be 16-byte aligned");
        p.topk_val = a.topk_val; p.topk_idx = a.topk_idx;
        if (cg == 2) return terms == 3 ? launch_variant<2, 3, EPI_TOPK, false>(maps, p, info.sm_count, stream)
                                       : launch_variant<2, 2, EPI_TOPK, false>(maps, p, info.sm_count, stream);
        return terms == 3 ? launch_variant<1, 3, EPI_TOPK, false>(maps, p, info.sm_count, stream)
                          : launch_variant<1, 2, EPI_TOPK, false>(maps, p, info.sm_count, stream);
    }
    if (f8) return cg == 2 ? launch_variant<2, 1, EPI_STATS, false, true>(maps, p, info.sm_count, stream)
                           : launch_variant<1, 1, EPI_STATS, false, true>(maps, p, info.sm_count, stream);
    if (a.y_aux) return dispatch<EPI_STATS, true>(cg, terms, maps, p, info.sm_count, stream);
    return dispatch<EPI_STATS, false>(cg, terms, maps, p, info.sm_count, stream);
}

}  // namespace tc

int launch_exact_stats(const pdm_stats_args& a, cudaStream_t stream);   // stats_exact.cu

}  // namespace pdm

using namespace pdm;

extern "C" int pdm_posterior_stats_plan(pdm_stats_args* a, int device, int64_t* partial_floats) {
    PDM_REQUIRE(a && a->M >= 0 && a->N > 0 && a->d > 0, "pdm_posterior_stats_plan: bad sizes");
    int sm = 0, maj = 0, min_ = 0;
    int rc = pdm_device_info(device, &sm, &maj, &min_);
    if (rc != PDM_OK) return rc;
    if (a->precision == PDM_PREC_EXACT_F32) {
        const int64_t m_tiles = std::max<int64_t>(1, ceil_div(a->M, 128)), n_tiles = ceil_div(a->N, 128);
        if (a->n_splits <= 0)
            a->n_splits = (int32_t)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(n_tiles, 64), ceil_div(4 * sm, m_tiles)));
        a->m_group = 1;
        a->cta_group = 1;
    } else {
        if (a->cta_group <= 0) a->cta_group = 2;
        const int cg = a->cta_group;
        PDM_REQUIRE(cg == 1 || cg == 2, "cta_group must be 1 or 2");
        const int pairs = sm / cg;
        const int64_t m_tiles = a->row_tiles ? std::max<int64_t>(1, a->n_row_tiles) : std::max<int64_t>(1, ceil_div(a->M, 128 * cg));
        const int64_t n_tiles = ceil_div(a->N, cg == 2 ? 256 : 128);
        const int64_t k_pad = round_up(a->d, 64);
        const int64_t a_tile_bytes = a->precision == PDM_PREC_F8X1 ? 128ll * cg * k_pad
                                   : 128ll * cg * k_pad * 2 * (a->precision == PDM_PREC_F16X1 ? 1 : 2);
        int g = a->m_group, s = a->n_splits;
        if (g <= 0 && s <= 0) tc::plan_schedule(pairs, m_tiles, n_tiles, a_tile_bytes, &g, &s);
        else if (g <= 0) g = (int)std::max<int64_t>(1, std::min<int64_t>(pairs / s, m_tiles));
        else if (s <= 0) s = (int)std::max<int64_t>(1, std::min<int64_t>(pairs / g, n_tiles));
        a->m_group = g; a->n_splits = s;
    }
    // the tensor path emits two records per split (one per column half of a tile)
    a->records_per_row = a->n_splits * (a->precision == PDM_PREC_EXACT_F32 ? 1 : 2);
    if (partial_floats) *partial_floats = a->M * a->records_per_row * PDM_PART_STRIDE;
    return PDM_OK;
}

extern "C" int pdm_posterior_stats(const pdm_stats_args* a, pdm_stream_t stream) {
    PDM_REQUIRE(a, "pdm_posterior_stats: null args");
    PDM_REQUIRE(a->M >= 0 && a->N > 0 && a->d > 0, "pdm_posterior_stats: bad sizes");
    PDM_REQUIRE(a->q_norm && a->y_norm, "pdm_posterior_stats: q_norm / y_norm missing");
    PDM_REQUIRE(a->partials || a->energy_out || a->topk_val, "pdm_posterior_stats: no output requested");
    PDM_REQUIRE(!a->topk_val || (a->precision == PDM_PREC_F16X3 || a->precision == PDM_PREC_F16X2),
                "pdm_posterior_stats: the top-k epilogue needs the tensor path (f16x3 / f16x2)");
    PDM_REQUIRE(!a->partials || a->inv_temp, "pdm_posterior_stats: inv_temp missing");
    PDM_REQUIRE(!a->energy_out || a->lde >= a->N, "pdm_posterior_stats: lde < N");
    PDM_REQUIRE(a->n_splits >= 1, "pdm_posterior_stats: n_splits must be planned (>= 1)");
    PDM_REQUIRE(!a->row_tiles || a->precision != PDM_PREC_EXACT_F32, "pdm_posterior_stats: row_tiles needs the tensor path");
    if (a->M == 0 || (a->row_tiles && a->n_row_tiles == 0)) return PDM_OK;
    switch (a->precision) {
        case PDM_PREC_EXACT_F32: return launch_exact_stats(*a, as_stream(stream));
        case PDM_PREC_F16X3:
        case PDM_PREC_F16X2:
        case PDM_PREC_F16X1:
        case PDM_PREC_F8X1: return tc::launch_tensor_stats(*a, as_stream(stream));
        default: set_error("unknown precision %d", a->precision); return PDM_ERR_INVALID_ARG;
    }
}

extern "C" int pdm_split_gemm_f16x3_tiles(const uint16_t* a_hi, const uint16_t* a_lo, int64_t lda, int64_t M,
                                          const uint16_t* b_hi, const uint16_t* b_lo, int64_t ldb, int64_t d, int64_t K,
                                          float scale, float* out, int64_t ldo, int32_t accumulate, int32_t cta_group,
                                          const int32_t* row_tiles, int64_t n_row_tiles, const int32_t* n_row_tiles_dev,
                                          pdm_stream_t stream) {
    PDM_REQUIRE(a_hi && a_lo && b_hi && out && M >= 0 && d > 0 && K > 0 && ldo >= d, "pdm_split_gemm_f16x3: bad arguments");
    PDM_REQUIRE(row_tiles || !n_row_tiles_dev, "pdm_split_gemm_f16x3_tiles: a device-side count needs a tile list");
    const int terms = b_lo ? 3 : 2;
    if (M == 0 || (row_tiles && n_row_tiles == 0)) return PDM_OK;
    DeviceInfo info;
    int rc = tc::require_sm100(&info);
    if (rc != PDM_OK) return rc;
    const int cg = cta_group <= 0 ? 2 : cta_group;
    PDM_REQUIRE(cg == 1 || cg == 2, "cta_group must be 1 or 2");
    PDM_REQUIRE(M < (1ll << 31) - 512 && d < (1ll << 31) - 512, "M and d must fit in int32");
    CUtensorMap maps[4];
    if ((rc = tc::make_tile_map(&maps[0], a_hi, M, K, lda)) != PDM_OK) return rc;
    if ((rc = tc::make_tile_map(&maps[1], a_lo, M, K, lda)) != PDM_OK) return rc;
    if ((rc = tc::make_tile_map(&maps[2], b_hi, d, K, ldb)) != PDM_OK) return rc;
    if ((rc = tc::make_tile_map(&maps[3], b_lo ? b_lo : b_hi, d, K, ldb)) != PDM_OK) return rc;
    const int block_n = cg == 2 ? 256 : 128;
    tc::GemmParams p = {};
    p.M = M; p.ncols = d;
    p.num_kb = (int32_t)ceil_div(K, tc::kBlockK);
    p.flush_kb = tc::flush_kb_setting(terms);
    p.wait_hint_ns = tc::wait_hint_setting();
    p.hint_a = tc::evict_hint_setting("PDM_HINT_A", ptx::kEvictNormal);
    p.hint_b = tc::evict_hint_setting("PDM_HINT_B", ptx::kEvictNormal);
    p.m_tiles = (int32_t)ceil_div(M, (int64_t)tc::kRowsPerCta * cg);
    if (row_tiles) {
        PDM_REQUIRE(n_row_tiles >= 1 && n_row_tiles <= p.m_tiles, "n_row_tiles must be between 1 and the number of row tiles");
        p.m_tiles = (int32_t)n_row_tiles;
        p.tile_list = row_tiles;
        p.m_tiles_dev = n_row_tiles_dev;
    }
    p.n_tiles = (int32_t)ceil_div(d, block_n);
    // every (row tile, column tile) is an independent output tile: spread column tiles first
    const int pairs = info.sm_count / cg;
    p.n_splits = (int32_t)std::min<int64_t>(p.n_tiles, pairs);
    p.m_group = (int32_t)std::max<int64_t>(1, std::min<int64_t>(pairs / p.n_splits, p.m_tiles));
    p.out = out; p.ldo = ldo; p.out_scale = scale; p.accumulate = accumulate;
    return tc::dispatch<tc::EPI_STORE, false>(cg, terms, maps, p, info.sm_count, as_stream(stream));
}

extern "C" int pdm_split_gemm_f16x3(const uint16_t* a_hi, const uint16_t* a_lo, int64_t lda, int64_t M,
                                    const uint16_t* b_hi, const uint16_t* b_lo, int64_t ldb, int64_t d, int64_t K,
                                    float scale, float* out, int64_t ldo, int32_t accumulate, int32_t cta_group,
                                    pdm_stream_t stream) {
    return pdm_split_gemm_f16x3_tiles(a_hi, a_lo, lda, M, b_hi, b_lo, ldb, d, K, scale, out, ldo, accumulate, cta_group,
                                      nullptr, 0, nullptr, stream);
}

#ifdef PDM_STALL_STATS
// dev builds only: copy the per-CTA stall counters (160 x 8 uint64) to the host and clear them
extern "C" int pdm_debug_read_stalls(unsigned long long* out_host) {
    PDM_CUDA_CHECK(cudaDeviceSynchronize());
    PDM_CUDA_CHECK(cudaMemcpyFromSymbol(out_host, pdm::tc::g_stall, sizeof(unsigned long long) * 160 * 8));
    static unsigned long long zeros[160 * 8] = {0};
    PDM_CUDA_CHECK(cudaMemcpyToSymbol(pdm::tc::g_stall, zeros, sizeof(zeros)));
    return PDM_OK;
}
// how many 2-CTA clusters of the f16x3 statistics kernel can be co-resident on the current device
extern "C" int pdm_debug_max_clusters(int* out) {
    using C = pdm::tc::Cfg<2, 3>;
    auto kern = pdm::tc::fused_gemm_kernel<2, 3, pdm::tc::EPI_STATS, false>;
    PDM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(pdm::tc::kThreads); cfg.dynamicSmemBytes = C::kSmemBytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    PDM_CUDA_CHECK(cudaOccupancyMaxActiveClusters(out, kern, &cfg));
    return PDM_OK;
}
#endif
