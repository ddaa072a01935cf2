// CTA-pair resident-operand GEMM core ("pair core"): tcgen05 cta_group::2, one 256 x 256 accumulator tile
// per pair of SMs, one operand parked in shared memory for the whole kernel.
//
// Why: a one-CTA 128 x N tile is bound by shared-memory bandwidth, not by the tensor pipe.  Every
// tcgen05.mma re-reads its A (128 x 16) and B (N x 16) slices from shared memory and every operand byte is
// first written there by TMA: the streaming forward tile (128 x 256 x 64 per 512 tensor cycles) asks for
// 96 KB per k-block = 192 B/clk against ~128 B/clk and measures 68 % tensor-busy; the one-CTA
// resident-operand kernels with N = 128 measure 30-42 % (profiles/r1_v3_*).  With a CTA pair each SM supplies
// only its 128 rows of A and its 128-row half of B for a 128 x 256 slice of the output (64 B/clk of MMA
// reads), and with one operand resident only 16 KB per k-block per SM is written by TMA (32 B/clk):
// 96 B/clk in total, leaving headroom for the epilogue's staging traffic.
//
// Per CTA (rank r of the pair), for K <= 512:
//   resident : 128 rows x K of the reused operand (rows res * 256 + r * 128 ...), loaded once
//   streamed : 128 rows x 64 per stage of the per-tile operand (rows s_row0 + i * 256 + r * 128 ...)
//   TMEM     : two accumulator buffers of 256 fp32 columns; the CTA's 128 lanes are its 128 rows of the tile
// K > 512 (Core::stream_both): the reused operand no longer fits beside the pipeline (128 rows x 1024 bf16 are
// 256 KB), so BOTH operands stream -- a stage holds the k-block of the per-tile operand (16 KB) and the k-block of
// the reused one (16 KB, re-fetched from L2 for every tile; B x D bf16 is a few MB and stays there).  Per SM that is
// 64 B/clk of TMA writes + 64 B/clk of MMA reads, still a third less shared-memory traffic than the one-CTA
// streaming tile (192 B/clk) the wide shapes -- D = 1024 / 1792 / 2816, B = 1024 -- used before.
// P::RES_A == false: the streamed operand is A (its rows are the accumulator rows), the resident one is B.
// P::RES_A == true : the resident operand is A, the streamed one is B (accumulator columns).
// Roles (384 threads): warp 0 TMA producer (both CTAs), warp 1 MMA issuer (leader CTA only), warp 2 TMEM
// allocator, warps 4-11 epilogue (two per TMEM lane quadrant, 128 accumulator columns each).  A policy may add
// P::AUX_WARPS helper warps (12 ...) that run P::aux() beside the GEMM -- the forward uses them to normalise
// and cast the class weights it is about to stream -- and gates each streamed tile on P::acquire_tile().
// Barriers: full[] / tempty[] / res live in the LEADER's shared memory (both CTAs' producers and epilogues
// arrive there through shared::cluster addresses); empty[] / tfull[] exist in both CTAs and are signalled by
// multicast tcgen05.commit.
//
// Reference counterpart: F.linear in arcface.py:47 and the autograd matmuls of loss.backward().
#pragma once
#include "ptx.cuh"

namespace ab {
namespace pr {

constexpr int ROWS = 128;  // rows of either operand held by one CTA
constexpr int BK = 64;
constexpr int ACC_COLS = 256;
constexpr int ACC_BUFS = 2;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = (4 + EPI_WARPS) * 32;
constexpr int TILE_BYTES = ROWS * BK * 2;  // 16 KB
constexpr int STAGING_PER_WARP = 4096;
constexpr int MAX_KBLOCKS = 8;   // resident mode: k-blocks of the parked operand
constexpr int MAX_STAGES = 8;    // barrier slots carved for the pipeline (Core::stages <= MAX_STAGES)

struct Core {
    int kblocks;         // ceil(K / 64); <= MAX_KBLOCKS unless stream_both
    int stream_both = 0; // 1: no resident operand, every stage carries a k-block of both operands
    int stages = 0;      // pipeline stages in use (set by the launcher: P::STAGES, or what fits when streaming)
    int s_blocks;        // 256-row blocks of the streamed operand handled by this launch
    int s_row0;          // TMA row coordinate of block 0 in the streamed tensor map
    int n_res;           // 256-row slices of the resident operand; the number of pairs is a multiple of it
    int contiguous;      // 1: group g walks blocks [g * per, (g + 1) * per); 0: g, g + groups, ...
    int prefetch_tiles;  // L2 prefetch distance in tiles (0 = off)
    // measurements only (nullable): 16 counters per CTA, see WaitProf
    unsigned long long* prof = nullptr;
    int prof_cta = 0;    // index of this launch's / role's first CTA in `prof`
};

// Where a CTA's warps spent their time (cycles), written once at the end of the body when Core::prof is set:
//   [0] producer waiting in acquire_tile   [1] producer waiting for a free stage
//   [2] MMA warp waiting for operands      [3] MMA warp waiting for a free accumulator
//   [4] epilogue warp 4 waiting for an accumulator   [5] epilogue warp 4 inside Epi::tile
//   [6] whole body   [7] tiles
struct WaitProf {
    unsigned long long t0, acc;
    bool on;
    __device__ __forceinline__ explicit WaitProf(bool enabled) : t0(0), acc(0), on(enabled) {}
    __device__ __forceinline__ void begin() { if (on) t0 = clock64(); }
    __device__ __forceinline__ void end() { if (on) acc += clock64() - t0; }
};

struct EpiCtx {
    uint8_t* extra;
    uint32_t staging;  // this warp's 4 KB staging buffer (only if P::STAGING)
    const CUtensorMap* tmC;
    int quad;   // TMEM lane quadrant
    int half;   // which 128 accumulator columns this warp owns
    int lane;
    int rank;   // CTA rank in the pair
    int res;    // resident slice of the pair
    int grp;    // group of the pair (pairs / n_res groups)
    unsigned long long* prof;  // measurements only: this CTA's counters [8..15] for warp 4, else null
};

// bytes of everything but the operand tiles
template <class P>
constexpr size_t smem_fixed_bytes(size_t extra_bytes) {
    return (P::STAGING ? EPI_WARPS * STAGING_PER_WARP : 0) + ((extra_bytes + 15) / 16) * 16 +
           (2 * MAX_STAGES + 2 * ACC_BUFS + 1) * 8 + 16;
}
template <class P>
constexpr size_t smem_bytes(int kblocks, size_t extra_bytes) {  // resident mode
    // no alignment slack: the kernels declare their dynamic shared memory __align__(1024) and trap if it is not
    return static_cast<size_t>(kblocks + P::STAGES) * TILE_BYTES + smem_fixed_bytes<P>(extra_bytes);
}
// pipeline stages a streaming launch gets out of `budget` bytes of shared memory
template <class P>
constexpr int stream_stages(size_t extra_bytes, size_t budget = 227 * 1024) {
    int n = static_cast<int>((budget - smem_fixed_bytes<P>(extra_bytes)) / (2 * TILE_BYTES));
    return n > MAX_STAGES ? MAX_STAGES : n;
}
template <class P>
constexpr size_t smem_bytes(const Core& co, size_t extra_bytes) {
    return co.stream_both ? static_cast<size_t>(co.stages) * 2 * TILE_BYTES + smem_fixed_bytes<P>(extra_bytes)
                          : static_cast<size_t>(co.kblocks + co.stages) * TILE_BYTES + smem_fixed_bytes<P>(extra_bytes);
}
// fills Core::kblocks / stream_both / stages for a contraction of depth K
template <class P>
inline void core_set_k(Core& co, int K, size_t extra_bytes) {
    co.kblocks = (K + BK - 1) / BK;
    co.stream_both = co.kblocks > MAX_KBLOCKS ? 1 : 0;
    co.stages = co.stream_both ? stream_stages<P>(extra_bytes) : P::STAGES;
}

// One 4 KB staging buffer per epilogue warp (32 rows x 128 B, TMA SWIZZLE_128B layout); see gemm_rs.cuh.
struct Stager {
    uint32_t buf;
    int lane;
    __device__ Stager(const EpiCtx& c) : buf(c.staging), lane(c.lane) {}
    __device__ __forceinline__ void acquire() const {
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
    }
    __device__ __forceinline__ void put(int chunk, uint32_t a, uint32_t b, uint32_t c, uint32_t d) const {
        st_shared_v4(buf + lane * 128 + ((chunk ^ (lane & 7)) << 4), a, b, c, d);
    }
    __device__ __forceinline__ void commit(const CUtensorMap* tm, int c0, int c1) const {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            tma_store_2d(tm, buf, c0, c1);
            bulk_commit();
        }
    }
    // the same with an L2 eviction policy on the written lines (0 = none)
    __device__ __forceinline__ void commit(const CUtensorMap* tm, int c0, int c1, uint64_t pol) const {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            if (pol != 0) tma_store_2d_hint(tm, buf, c0, c1, pol);
            else tma_store_2d(tm, buf, c0, c1);
            bulk_commit();
        }
    }
    __device__ __forceinline__ void drain() const {
        if (lane == 0) bulk_wait<0>();
        __syncwarp();
    }
};

// Two 2 KB staging buffers per epilogue warp (32 rows x 64 B, TMA SWIZZLE_64B layout: the 16-byte chunk index
// is XORed with bits 7-8 of the row offset): the next box is written while the TMA still reads the previous one.
struct Stager2 {
    uint32_t base;
    int lane;
    int it;
    __device__ Stager2(const EpiCtx& c) : base(c.staging), lane(c.lane), it(0) {}
    // buffer for the next box; blocks until the store issued two boxes ago has read it out
    __device__ __forceinline__ uint32_t acquire() const {
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
        return base + (it & 1) * 2048;
    }
    __device__ __forceinline__ void put(uint32_t buf, int chunk, uint32_t a, uint32_t b, uint32_t c, uint32_t d) const {
        st_shared_v4(buf + lane * 64 + ((chunk ^ ((lane >> 1) & 3)) << 4), a, b, c, d);
    }
    __device__ __forceinline__ void commit(const CUtensorMap* tm, uint32_t buf, int c0, int c1) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            tma_store_2d(tm, buf, c0, c1);
            bulk_commit();
        }
        ++it;
    }
    __device__ __forceinline__ void drain() const {
        if (lane == 0) bulk_wait<0>();
        __syncwarp();
    }
};

// Optional hooks of a pair policy (a policy derives from this and hides what it needs).
struct PairDefaults {
    static constexpr int AUX_WARPS = 0;                             // helper warps running aux()
    static constexpr int LOW_REGS = 0, EPI_REGS = 0, AUX_REGS = 0;  // setmaxnreg targets (0 = no reallocation)
    // producer warp (all lanes), before the first k-block of tile i of this CTA is fetched
    // (`extra` = the policy's shared memory, the same pointer Epi sees as EpiCtx::extra)
    template <class Prm>
    __device__ static void acquire_tile(const Prm&, uint8_t*, int, int, int) {}
    // one lane of the leader's MMA warp, after every operand byte of tile i has landed in both CTAs
    template <class Prm>
    __device__ static void release_tile(const Prm&, int) {}
    template <class Prm>
    __device__ static void aux(const Prm&, int, int, int) {}
    // warp 3 (otherwise idle), all lanes: (prm, extra, i_begin, i_end, i_step, rank, lane) = this pair's tile walk
    template <class Prm>
    __device__ static void side_warp(const Prm&, uint8_t*, int, int, int, int, int) {}
    // L2 eviction-priority policy for the TMA loads of the streamed operand (0 = none)
    __device__ static uint64_t stream_policy() { return 0ull; }
    // TMA row coordinate of tile i of the streamed operand (the CTA adds rank * ROWS)
    template <class Prm>
    __device__ static int stream_row(const Prm& p, int i) { return p.core.s_row0 + i * 2 * ROWS; }
};

// Registers per thread a kernel of `threads` threads is launched with under __launch_bounds__(threads, 1).
__host__ __device__ constexpr int launch_regs(int threads) {
    int r = 65536 / threads;
    r = r > 255 ? 255 : r;
    return r / 8 * 8;
}
template <int N, int LAUNCH>
__device__ __forceinline__ void setmaxnreg_to() {
    if constexpr (N > LAUNCH) setmaxnreg_inc<N>();
    else if constexpr (N < LAUNCH) setmaxnreg_dec<N>();
}

// The body of a pair kernel: `pair` of `npairs` (consecutive CTA pairs of one launch that share the policy),
// `smem` 1024-aligned dynamic shared memory carved identically in both CTAs.
template <class P>
__device__ __forceinline__ void pair_gemm_body(const CUtensorMap& tmS, const CUtensorMap& tmR, const CUtensorMap& tmC,
                                               const typename P::Params& prm, const int extra_bytes, uint8_t* smem,
                                               const int pair, const int npairs) {
    const Core& co = prm.core;
    const int kblocks = co.kblocks;
    const bool stream = co.stream_both != 0;
    const int STAGES = co.stages;
    const int STAGE_BYTES = stream ? 2 * TILE_BYTES : TILE_BYTES;
    uint8_t* sRes = smem;
    uint8_t* sStage = sRes + (stream ? 0 : kblocks * TILE_BYTES);
    uint8_t* sStaging = sStage + STAGES * STAGE_BYTES;
    uint8_t* sExtra = sStaging + (P::STAGING ? EPI_WARPS * STAGING_PER_WARP : 0);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sExtra + ((extra_bytes + 15) / 16) * 16);
    uint64_t* empty_bar = full_bar + MAX_STAGES;
    uint64_t* tfull_bar = empty_bar + MAX_STAGES;
    uint64_t* tempty_bar = tfull_bar + ACC_BUFS;
    uint64_t* res_bar = tempty_bar + ACC_BUFS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = static_cast<int>(cluster_ctarank());
    const int res = pair % co.n_res;
    const int grp = pair / co.n_res;
    const int ngrp = npairs / co.n_res;
    // tile walk of this pair
    int i_begin, i_end, i_step;
    if (co.contiguous) {
        const int per = (co.s_blocks + ngrp - 1) / ngrp;
        i_begin = grp * per;
        i_end = min(co.s_blocks, i_begin + per);
        i_step = 1;
    } else {
        i_begin = grp;
        i_end = co.s_blocks;
        i_step = ngrp;
    }

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmS);
        tma_prefetch_desc(&tmR);
        if (P::STAGING) tma_prefetch_desc(&tmC);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < MAX_STAGES; ++i) {
            mbar_init(&full_bar[i], 2);   // one arrive.expect_tx per CTA of the pair (used in the leader only)
            mbar_init(&empty_bar[i], 1);  // multicast tcgen05.commit
        }
        for (int i = 0; i < ACC_BUFS; ++i) {
            mbar_init(&tfull_bar[i], 1);                // multicast tcgen05.commit
            mbar_init(&tempty_bar[i], 2 * EPI_WARPS);   // every epilogue warp of both CTAs (leader only)
        }
        mbar_init(res_bar, 2);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc_pair(tmem_slot, ACC_BUFS * ACC_COLS);
        tmem_relinquish_pair();
    }
    P::prologue(prm, sExtra, threadIdx.x, res, rank);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's barriers are initialised before anyone arrives on them remotely
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const bool prof_on = co.prof != nullptr;
    unsigned long long* prof = prof_on ? co.prof + static_cast<size_t>(co.prof_cta + pair * 2 + rank) * 16 : nullptr;
    const unsigned long long body_t0 = prof_on ? clock64() : 0;

    // A policy with helper warps may let the light roles hand registers to the helper warpgroups (P::AUX_REGS > 0;
    // setmaxnreg is executed by whole warpgroups -- warps 0-3, 4-11, 12-... -- at the top of the branch that holds
    // the role's code, so that ptxas allocates each role against its own limit; the pool is what the CTA was
    // launched with, threads x registers, not the whole register file).
    constexpr int LAUNCH_REGS = launch_regs(THREADS + P::AUX_WARPS * 32);
    if (warp < 4) {
    if constexpr (P::AUX_REGS > 0) setmaxnreg_to<P::LOW_REGS, LAUNCH_REGS>();
    if (warp == 0) {
        // ---------------- TMA producer (both CTAs): whole warp walks the schedule, one elected lane issues
        const uint32_t res_bar_leader = mapa_u32(smem_u32(res_bar), 0);
        const int res_row = res * 2 * ROWS + rank * ROWS;
        if (!stream && elect_one()) {
            mbar_expect_tx_cluster(res_bar_leader, kblocks * TILE_BYTES);
            for (int kb = 0; kb < kblocks; ++kb)
                tma_load_2d_pair(sRes + kb * TILE_BYTES, &tmR, res_bar_leader, kb * BK, res_row);
        }
        __syncwarp();
        int stage = 0;
        uint32_t phase = 0;
        WaitProf wp_a(prof_on), wp_b(prof_on);
        const uint64_t s_pol = P::stream_policy();   // L2 eviction priority of the streamed operand's lines (0 = default)
        for (int i = i_begin; i < i_end; i += i_step) {
            const int row = P::stream_row(prm, i) + rank * ROWS;
            wp_a.begin();
            P::acquire_tile(prm, sExtra, i, rank, lane);  // whole warp; returns once this CTA's 128 rows may be fetched
            wp_a.end();
            const int ip = i + co.prefetch_tiles * i_step;
            const bool pf = co.prefetch_tiles > 0 && ip < i_end;
            for (int kb = 0; kb < kblocks; ++kb) {
                wp_b.begin();
                mbar_wait(&empty_bar[stage], phase ^ 1);
                wp_b.end();
                if (elect_one()) {
                    const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
                    mbar_expect_tx_cluster(full_leader, STAGE_BYTES);
                    if (s_pol != 0) tma_load_2d_pair_hint(sStage + stage * STAGE_BYTES, &tmS, full_leader, kb * BK, row, s_pol);
                    else tma_load_2d_pair(sStage + stage * STAGE_BYTES, &tmS, full_leader, kb * BK, row);
                    if (stream) tma_load_2d_pair(sStage + stage * STAGE_BYTES + TILE_BYTES, &tmR, full_leader, kb * BK, res_row);
                    // the n_res pairs that stream the same block share the L2 prefetch work
                    if (pf && (kb % co.n_res) == res)
                        tma_prefetch_2d(&tmS, kb * BK, P::stream_row(prm, ip) + rank * ROWS);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
        if (prof_on && lane == 0) { prof[0] = wp_a.acc; prof[1] = wp_b.acc; }
    } else if (warp == 1) {
        // ---------------- MMA issuer: leader CTA only, whole warp runs the loop, one elected lane issues
        if (rank == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(2 * ROWS, ACC_COLS, false, false);
            const uint32_t sStage_u32 = smem_u32(sStage);
            const uint64_t res_desc0 = make_smem_desc(smem_u32(sRes), 16, 1024);
            if (!stream) {
                mbar_wait_cluster(res_bar, 0);
                tc_fence_after();
            }
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            WaitProf wp_f(prof_on), wp_e(prof_on);
            for (int i = i_begin; i < i_end; i += i_step) {
                wp_e.begin();
                mbar_wait_cluster(&tempty_bar[acc], acc_phase ^ 1);
                wp_e.end();
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * ACC_COLS;
                for (int kb = 0; kb < kblocks; ++kb) {
                    wp_f.begin();
                    mbar_wait_cluster(&full_bar[stage], phase);
                    wp_f.end();
                    tc_fence_after();
                    const uint64_t s_desc = make_smem_desc(sStage_u32 + stage * STAGE_BYTES, 16, 1024);
                    const uint64_t r_desc = stream ? s_desc + static_cast<uint64_t>(TILE_BYTES >> 4)
                                                   : res_desc0 + static_cast<uint64_t>(kb * (TILE_BYTES >> 4));
                    const uint64_t a_desc = P::RES_A ? r_desc : s_desc;
                    const uint64_t b_desc = P::RES_A ? s_desc : r_desc;
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < BK / 16; ++kk)
                            umma_bf16_pair(tmem_d, a_desc + 2 * kk, b_desc + 2 * kk, idesc, (kb | kk) != 0 ? 1u : 0u);
                        umma_commit_pair(&empty_bar[stage], 3);  // frees the stage in both CTAs
                        if (kb == kblocks - 1) {
                            umma_commit_pair(&tfull_bar[acc], 3);
                            P::release_tile(prm, i);
                        }
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == ACC_BUFS) { acc = 0; acc_phase ^= 1; }
            }
            if (prof_on && lane == 0) { prof[2] = wp_f.acc; prof[3] = wp_e.acc; }
        }
    } else if (warp == 3) {
        P::side_warp(prm, sExtra, i_begin, i_end, i_step, rank, lane);
    }
    } else if (warp < 4 + EPI_WARPS) {
        // ---------------- epilogue (both CTAs)
        if constexpr (P::AUX_REGS > 0) setmaxnreg_to<P::EPI_REGS, LAUNCH_REGS>();
        EpiCtx ctx;
        ctx.extra = sExtra;
        ctx.staging = smem_u32(sStaging) + (warp - 4) * STAGING_PER_WARP;
        ctx.tmC = &tmC;
        ctx.quad = warp & 3;
        ctx.half = (warp - 4) >> 2;
        ctx.lane = lane;
        ctx.rank = rank;
        ctx.res = res;
        ctx.grp = grp;
        ctx.prof = (prof_on && warp == 4) ? prof + 8 : nullptr;
        typename P::Epi epi(prm, ctx);
        int acc = 0;
        uint32_t acc_phase = 0;
        if (i_begin < i_end) epi.prefetch(i_begin);
        WaitProf wp_t(prof_on && warp == 4), wp_w(prof_on && warp == 4);
        unsigned long long n_tiles = 0;
        for (int i = i_begin; i < i_end; i += i_step) {
            wp_t.begin();
            mbar_wait(&tfull_bar[acc], acc_phase);
            wp_t.end();
            ++n_tiles;
            wp_w.begin();
            tc_fence_after();
            const uint32_t taddr =
                tmem_base + acc * ACC_COLS + ctx.half * 128 + (static_cast<uint32_t>(ctx.quad * 32) << 16);
            epi.tile(i, i + i_step < i_end ? i + i_step : -1, taddr);  // returns after its last tcgen05.ld completed
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
            wp_w.end();
            if (++acc == ACC_BUFS) { acc = 0; acc_phase ^= 1; }
        }
        epi.finish();
        if (prof_on && warp == 4 && lane == 0) {
            prof[4] = wp_t.acc;
            prof[5] = wp_w.acc;
            prof[6] = clock64() - body_t0;
            prof[7] = n_tiles;
        }
    } else if constexpr (P::AUX_WARPS > 0) {
        if constexpr (P::AUX_REGS > 0) setmaxnreg_to<P::AUX_REGS, LAUNCH_REGS>();
        P::aux(prm, (pair * 2 + rank) * P::AUX_WARPS + (warp - 4 - EPI_WARPS), npairs * 2 * P::AUX_WARPS, lane);
    }

    // neither CTA may leave (or free TMEM) while the peer can still touch its barriers / shared memory / TMEM
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc_pair(tmem_base, ACC_BUFS * ACC_COLS);
}

template <class P>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS + P::AUX_WARPS * 32, 1)
pair_gemm_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmR,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ typename P::Params prm,
                 const int extra_bytes) {
    // 1024-byte alignment (SWIZZLE_128B operand tiles); the two CTAs carve identically
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    uint8_t* smem = smem_raw;
    pair_gemm_body<P>(tmS, tmR, tmC, prm, extra_bytes, smem, blockIdx.x >> 1, gridDim.x >> 1);
}

#ifdef AB_CHECK_CUDA
template <class P>
static int32_t launch_pair(const CUtensorMap& tmS, const CUtensorMap& tmR, const CUtensorMap& tmC,
                           const typename P::Params& prm, int groups, int extra_bytes, cudaStream_t st) {
    const size_t smem = smem_bytes<P>(prm.core, extra_bytes);
    AB_REQUIRE(prm.core.kblocks >= 1 && (prm.core.stream_both || prm.core.kblocks <= MAX_KBLOCKS) &&
                   prm.core.stages >= 2 && prm.core.stages <= MAX_STAGES && smem <= 227 * 1024,
               ARCFACE_B200_E_SHAPE, "CTA-pair kernel: %d k-blocks / %d stages / %zu bytes of shared memory do not fit",
               prm.core.kblocks, prm.core.stages, smem);
    AB_REQUIRE(groups >= 1 && prm.core.n_res >= 1, ARCFACE_B200_E_SHAPE, "empty grid");
    static bool configured[64] = {false};
    int dev = 0;
    AB_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        AB_CHECK_CUDA(cudaFuncSetAttribute(pair_gemm_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured[dev] = true;
    }
    pair_gemm_kernel<P><<<2 * groups * prm.core.n_res, THREADS + P::AUX_WARPS * 32, smem, st>>>(tmS, tmR, tmC, prm,
                                                                                                  extra_bytes);
    AB_CHECK_CUDA(cudaGetLastError());
    return ARCFACE_B200_OK;
}
#endif

}  // namespace pr
}  // namespace ab
