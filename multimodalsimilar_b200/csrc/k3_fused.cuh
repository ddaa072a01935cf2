// Single-launch backward of the ArcFace head ("fused through L2").
//
// The three contractions of the backward -- S^T -> dC^T (recompute + softmax gradient), dW = dC^T . Xhat and
// dXhat = dC . What -- each want a different operand parked in shared memory and a different accumulator
// shape in TMEM (a 128-lane x 512-column fp32 dW or dX tile fills the whole TMEM of an SM), so they cannot
// share one CTA.  Run as three kernels they exchange dC^T through HBM: 1 GB written and 2 GB read back per step
// at C = 1M, B = 512, which makes the backward HBM-bound (profiles/r1_v3_*).  Here the three run as ROLES of
// one persistent launch, each on its own set of SMs, and exchange dC^T through a small ring that lives in L2:
//
//   role DC (CTA pairs, gemm_pair.cuh / BwdDCpT<true>): block j (256 classes) -> ring slot j % R, q[j]
//   role DW (CTA pairs, BwdDWpT<true>)                 : waits for block j, streams its slot, writes dW rows
//   role DX (single CTAs, dx_body below)               : waits for block j, accumulates dC . What in TMEM over
//                                                        every block of its split, one reduce-add at the end
//
// Synchronisation is two counters per block in global memory (struct Ring): `ready` (bumped once per producer
// CTA, by its publisher warp, after the epilogue warps' TMA stores have completed) and `done` (bumped by every
// consumer once the block's last operand byte has landed in its shared memory; the producer of block j + R
// waits for it).  Every role walks its blocks in ascending order and a producer publishes a finished block
// before it waits for a slot, so the smallest unfinished block can always make progress: no deadlock for any
// R >= 1, PROVIDED every CTA of the launch is resident at once -- the launcher sizes the grid with
// cudaOccupancyMaxActiveClusters and the spin loops trap instead of hanging (ptx.cuh, AB_SPIN_LIMIT).
// HBM traffic per step drops from ~8.3 GB to What (1 GB, read once: the dX role and the dW epilogue hit L2)
// plus dW (2 GB written).
//
// Reference counterpart: loss.backward() through arcface.py:45-63.
#pragma once
#include "gemm_core.cuh"
#include "gemm_pair.cuh"

namespace ab {
namespace fz {

constexpr int DX_STAGES = 3;
constexpr int DX_TILE = 256;                         // output tile: 256 batch rows x 256 embedding columns
constexpr int DX_OPER_BYTES = DX_TILE * BLOCK_K * 2;  // 32 KB: one k-block (64 classes) of either operand
constexpr int DX_KB_PER_BLOCK = 256 / BLOCK_K;        // k-blocks per 256-class block

struct DXParams {
    int B, D;
    int n_blocks;           // 256-class blocks
    int m_tiles, dn_tiles;  // output tiles along batch / embedding
    int splits;             // class splits; CTA x handles tile x % (m_tiles * dn_tiles) of split x / (...)
    int part_rows;          // > 0: split sp STORES its tile into rows [sp * part_rows, ...) of the output map (a
                            // [n_parts * part_rows][D] scratch); the warp that completes a 32-row x 256-column region --
                            // the last of the n_parts splits to arrive there, whichever that is -- adds the splits'
                            // copies in split order and writes dXhat: bit-reproducible, no zero-fill, no extra launch.
                            // 0: every split reduce-adds into dXhat [B][D] (order not fixed)
    int n_parts;            // splits that have work (min(splits, n_blocks))
    const float* parts;     // the scratch behind the output map
    int* region_cnt;        // [regions] arrivals per region, zeroed before the launch
    float* dx_out;          // dXhat [B][D]
    Ring ring;
    unsigned long long* prof = nullptr;  // measurements only: 16 counters per CTA (pr::WaitProf layout)
    int prof_cta = 0;
};

constexpr size_t dx_smem_bytes() {
    return static_cast<size_t>(DX_STAGES) * 2 * DX_OPER_BYTES + pr::EPI_WARPS * pr::STAGING_PER_WARP +
           (2 * DX_STAGES + 1) * 8 + 16;
}

// One epilogue warp has stored its 32-row x 256-column region of split `sp`'s tile (TMA stores committed by lane 0).
// Counts the arrival; the warp that is last at this region sums the n_parts copies in split order into dXhat.
__device__ __forceinline__ void dx_region_done(const DXParams& p, int region, int row0, int col0, int lane) {
    int prev = 0;
    if (lane == 0) {
        bulk_wait<0>();          // this warp's stores have been performed
        __threadfence();         // ... and are visible device-wide before the arrival is
        prev = atomicAdd(p.region_cnt + region, 1);
    }
    prev = __shfl_sync(0xffffffffu, prev, 0);
    if (prev != p.n_parts - 1) return;
    __threadfence();             // the other splits' stores (published before their arrivals) are visible to the loads below
    const int64_t part_stride = static_cast<int64_t>(p.part_rows) * p.D;
    const int c = col0 + lane * 8;
#pragma unroll 1
    for (int r = 0; r < 32; ++r) {
        const int row = row0 + r;
        if (row >= p.B) break;
        if (c >= p.D) continue;   // (D % 8 == 0: a lane's eight columns are all inside or all outside)
        const float* src = p.parts + static_cast<int64_t>(row) * p.D + c;
        float4 a0 = __ldcg(reinterpret_cast<const float4*>(src));
        float4 a1 = __ldcg(reinterpret_cast<const float4*>(src + 4));
        for (int sp = 1; sp < p.n_parts; ++sp) {
            const float4 b0 = __ldcg(reinterpret_cast<const float4*>(src + sp * part_stride));
            const float4 b1 = __ldcg(reinterpret_cast<const float4*>(src + sp * part_stride + 4));
            a0.x += b0.x; a0.y += b0.y; a0.z += b0.z; a0.w += b0.w;
            a1.x += b1.x; a1.y += b1.y; a1.z += b1.z; a1.w += b1.w;
        }
        float* dst = p.dx_out + static_cast<int64_t>(row) * p.D + c;
        *reinterpret_cast<float4*>(dst) = a0;
        *reinterpret_cast<float4*>(dst + 4) = a1;
    }
}

// dXhat role.  `x` = index of this CTA among the dX CTAs.  cta_group::1 throughout (the two CTAs of a cluster
// are independent here).
__device__ __forceinline__ void dx_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                                        const DXParams& p, uint8_t* smem, const int x) {
    uint8_t* sA = smem;
    uint8_t* sB = sA + DX_STAGES * DX_OPER_BYTES;
    uint8_t* sStaging = sB + DX_STAGES * DX_OPER_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sStaging + pr::EPI_WARPS * pr::STAGING_PER_WARP);
    uint64_t* empty_bar = full_bar + DX_STAGES;
    uint64_t* tfull_bar = empty_bar + DX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tiles = p.m_tiles * p.dn_tiles;
    const int t = x % tiles;
    const int sp = x / tiles;
    const int m0 = (t % p.m_tiles) * DX_TILE;
    const int n0 = (t / p.m_tiles) * DX_TILE;
    const bool has_work = sp < p.splits && sp < p.n_blocks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < DX_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(tfull_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const bool prof_on = p.prof != nullptr;
    unsigned long long* prof = prof_on ? p.prof + static_cast<size_t>(p.prof_cta + x) * 16 : nullptr;
    const unsigned long long body_t0 = prof_on ? clock64() : 0;

    if (has_work) {
        if (warp == 0) {
            int stage = 0;
            uint32_t phase = 0;
            pr::WaitProf wp_a(prof_on), wp_b(prof_on);
            for (int j = sp; j < p.n_blocks; j += p.splits) {
                wp_a.begin();
                if (lane == 0) wait_counter_ge(p.ring.ready + j, p.ring.ready_target);
                __syncwarp();
                fence_proxy_async_all();
                wp_a.end();
                const int ring_row = (j % p.ring.slots) * 256;
                for (int kb = 0; kb < DX_KB_PER_BLOCK; ++kb) {
                    wp_b.begin();
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    wp_b.end();
                    if (elect_one()) {
                        mbar_expect_tx(&full_bar[stage], 2 * DX_OPER_BYTES);
                        // both operands MN-major: {64 mn x 64 k} boxes landing as [mn / 64][64 k][64]
#pragma unroll
                        for (int c = 0; c < DX_TILE / 64; ++c) {
                            tma_load_2d(sA + stage * DX_OPER_BYTES + c * (BLOCK_K * 128), &tmA, &full_bar[stage],
                                        m0 + c * 64, ring_row + kb * BLOCK_K);
                            tma_load_2d(sB + stage * DX_OPER_BYTES + c * (BLOCK_K * 128), &tmB, &full_bar[stage],
                                        n0 + c * 64, j * 256 + kb * BLOCK_K);
                        }
                    }
                    __syncwarp();
                    if (++stage == DX_STAGES) { stage = 0; phase ^= 1; }
                }
            }
            if (prof_on && lane == 0) { prof[0] = wp_a.acc; prof[1] = wp_b.acc; }
        } else if (warp == 1) {
            constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, DX_TILE, true, true);
            const uint32_t sA_u32 = smem_u32(sA);
            const uint32_t sB_u32 = smem_u32(sB);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t accumulate = 0;
            pr::WaitProf wp_f(prof_on);
            unsigned long long n_blk = 0;
            for (int j = sp; j < p.n_blocks; j += p.splits) {
                const bool last_block = j + p.splits >= p.n_blocks;
                ++n_blk;
                for (int kb = 0; kb < DX_KB_PER_BLOCK; ++kb) {
                    wp_f.begin();
                    mbar_wait(&full_bar[stage], phase);
                    wp_f.end();
                    tc_fence_after();
                    const uint32_t a_base = sA_u32 + stage * DX_OPER_BYTES;
                    const uint32_t b_base = sB_u32 + stage * DX_OPER_BYTES;
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
                            const uint64_t bdesc = operand_desc<true>(b_base, kk);
#pragma unroll
                            for (int ms = 0; ms < 2; ++ms)
                                umma_bf16(tmem_base + ms * DX_TILE, operand_desc<true>(a_base + ms * A_SUB_BYTES, kk), bdesc,
                                          idesc, (accumulate | kk) != 0 ? 1u : 0u);
                        }
                        umma_commit(&empty_bar[stage]);
                        if (kb == DX_KB_PER_BLOCK - 1) {
                            red_relaxed_gpu_add(p.ring.done + j, 1);  // every byte of the block is in shared memory
                            if (last_block) umma_commit(tfull_bar);
                        }
                    }
                    __syncwarp();
                    accumulate = 1;
                    if (++stage == DX_STAGES) { stage = 0; phase ^= 1; }
                }
            }
            if (prof_on && lane == 0) { prof[2] = wp_f.acc; prof[6] = clock64() - body_t0; prof[7] = n_blk; }
        } else if (warp >= 4) {
            // two epilogue warps per TMEM lane quadrant, one per stacked 128-row sub-tile
            const int quad = warp & 3;
            const int ms = (warp - 4) >> 2;
            pr::EpiCtx ctx;
            ctx.staging = smem_u32(sStaging) + (warp - 4) * pr::STAGING_PER_WARP;
            ctx.lane = lane;
            ctx.prof = nullptr;
            pr::Stager stager(ctx);
            mbar_wait(tfull_bar, 0);
            tc_fence_after();
            const int row0 = m0 + ms * BLOCK_M + quad * 32;
            if (row0 < p.B) {  // warp-uniform: otherwise these 32 rows are batch padding
                const uint32_t taddr = tmem_base + ms * DX_TILE + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
                for (int cc = 0; cc < DX_TILE / 32; ++cc) {
                    const int d0 = n0 + cc * 32;
                    if (d0 >= p.D) break;
                    uint32_t v[32];
                    tmem_ld32(taddr + cc * 32, v);
                    tmem_ld_wait();
                    stager.acquire();
#pragma unroll
                    for (int g = 0; g < 8; ++g) stager.put(g, v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        if (p.part_rows > 0) tma_store_2d(&tmC, ctx.staging, d0, sp * p.part_rows + row0);
                        else tma_reduce_add_2d(&tmC, ctx.staging, d0, row0);  // rows >= B / columns >= D are clipped
                        bulk_commit();
                    }
                }
                if (p.part_rows > 0 && p.region_cnt != nullptr) dx_region_done(p, t * 8 + (warp - 4), row0, n0, lane);
                stager.drain();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// dXhat role on a CTA pair (cta_group::2): one 256 (batch) x 512 (embedding) output tile per pair -- each CTA holds its
// 128 batch rows of the tile in all 512 TMEM columns (two N = 256 accumulators).  Why: the one-CTA role above pulls 64 KB
// of operands per 1024 tensor cycles (64 B/clk, the most an SM gets out of L2) and measured 27 % of its time waiting for
// them; a pair shares the dC^T slice (each CTA loads its 128 batch columns: 16 KB per 64-class k-block) and splits each
// 256-column slice of What between the two CTAs (2 x 16 KB), 48 KB per CTA for the same 1024 cycles.
// `x` = index of this pair among the dX pairs: tile x % tiles of class split x / tiles.
constexpr int DXP_STAGES = 4;
constexpr int DXP_A_BYTES = 128 * BLOCK_K * 2;           // 16 KB: this CTA's 128 batch columns of a 64-class k-block
constexpr int DXP_B_BYTES = 128 * BLOCK_K * 2;           // 16 KB: this CTA's 128 embedding columns of one N = 256 slice
constexpr int DXP_STAGE_BYTES = DXP_A_BYTES + 2 * DXP_B_BYTES;

constexpr size_t dxp_smem_bytes() {
    return static_cast<size_t>(DXP_STAGES) * DXP_STAGE_BYTES + pr::EPI_WARPS * pr::STAGING_PER_WARP +
           (2 * DXP_STAGES + 1) * 8 + 16;
}

__device__ __forceinline__ void dx_pair_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                                             const DXParams& p, uint8_t* smem, const int x) {
    uint8_t* sStage = smem;
    uint8_t* sStaging = sStage + DXP_STAGES * DXP_STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sStaging + pr::EPI_WARPS * pr::STAGING_PER_WARP);
    uint64_t* empty_bar = full_bar + DXP_STAGES;
    uint64_t* tfull_bar = empty_bar + DXP_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = static_cast<int>(cluster_ctarank());
    const int tiles = p.m_tiles * p.dn_tiles;   // here dn_tiles counts 512-column slices
    const int t = x % tiles;
    const int sp = x / tiles;
    const int m0 = (t % p.m_tiles) * 256 + rank * 128;   // first batch row of this CTA
    const int n0 = (t / p.m_tiles) * 512;                // first embedding column of the pair's tile
    const int n_nb = (p.D - n0 > 256) ? 2 : 1;           // N = 256 slices of the tile that hold real columns
    const bool has_work = sp < p.splits && sp < p.n_blocks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < DXP_STAGES; ++i) {
            mbar_init(&full_bar[i], 2);   // one arrive.expect_tx per CTA (used in the leader only)
            mbar_init(&empty_bar[i], 1);  // multicast tcgen05.commit
        }
        mbar_init(tfull_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc_pair(tmem_slot, 512);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const bool prof_on = p.prof != nullptr;
    unsigned long long* prof = prof_on ? p.prof + static_cast<size_t>(p.prof_cta + x * 2 + rank) * 16 : nullptr;
    const unsigned long long body_t0 = prof_on ? clock64() : 0;

    if (has_work) {
        if (warp == 0) {
            // ---------------- TMA producer (both CTAs)
            int stage = 0;
            uint32_t phase = 0;
            pr::WaitProf wp_a(prof_on), wp_b(prof_on);
            for (int j = sp; j < p.n_blocks; j += p.splits) {
                wp_a.begin();
                if (lane == 0) wait_counter_ge(p.ring.ready + j, p.ring.ready_target);
                __syncwarp();
                fence_proxy_async_all();
                wp_a.end();
                const int ring_row = (j % p.ring.slots) * 256;
                for (int kb = 0; kb < DX_KB_PER_BLOCK; ++kb) {
                    wp_b.begin();
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    wp_b.end();
                    if (elect_one()) {
                        const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
                        mbar_expect_tx_cluster(full_leader, DXP_A_BYTES + n_nb * DXP_B_BYTES);
                        uint8_t* st = sStage + stage * DXP_STAGE_BYTES;
                        // MN-major operands: {64 mn x 64 k} boxes landing as [mn / 64][64 k][64]
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                            tma_load_2d_pair(st + c * (BLOCK_K * 128), &tmA, full_leader, m0 + c * 64, ring_row + kb * BLOCK_K);
                        for (int nb = 0; nb < n_nb; ++nb) {
#pragma unroll
                            for (int c = 0; c < 2; ++c)
                                tma_load_2d_pair(st + DXP_A_BYTES + nb * DXP_B_BYTES + c * (BLOCK_K * 128), &tmB, full_leader,
                                                 n0 + nb * 256 + rank * 128 + c * 64, j * 256 + kb * BLOCK_K);
                        }
                    }
                    __syncwarp();
                    if (++stage == DXP_STAGES) { stage = 0; phase ^= 1; }
                }
            }
            if (prof_on && lane == 0) { prof[0] = wp_a.acc; prof[1] = wp_b.acc; }
        } else if (warp == 1) {
            // ---------------- MMA issuer (leader CTA)
            if (rank == 0) {
                constexpr uint32_t idesc = make_idesc_bf16(256, 256, true, true);
                const uint32_t sStage_u32 = smem_u32(sStage);
                int stage = 0;
                uint32_t phase = 0;
                uint32_t accumulate = 0;
                pr::WaitProf wp_f(prof_on);
                unsigned long long n_blk = 0;
                for (int j = sp; j < p.n_blocks; j += p.splits) {
                    const bool last_block = j + p.splits >= p.n_blocks;
                    ++n_blk;
                    for (int kb = 0; kb < DX_KB_PER_BLOCK; ++kb) {
                        wp_f.begin();
                        mbar_wait_cluster(&full_bar[stage], phase);
                        wp_f.end();
                        tc_fence_after();
                        const uint32_t a_base = sStage_u32 + stage * DXP_STAGE_BYTES;
                        if (elect_one()) {
#pragma unroll
                            for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
                                const uint64_t adesc = operand_desc<true>(a_base, kk);
                                for (int nb = 0; nb < n_nb; ++nb)
                                    umma_bf16_pair(tmem_base + nb * 256, adesc,
                                                   operand_desc<true>(a_base + DXP_A_BYTES + nb * DXP_B_BYTES, kk), idesc,
                                                   (accumulate | kk) != 0 ? 1u : 0u);
                            }
                            umma_commit_pair(&empty_bar[stage], 3);
                            if (kb == DX_KB_PER_BLOCK - 1) {
                                red_relaxed_gpu_add(p.ring.done + j, 1);  // every byte of the block is in both CTAs' shared memory
                                if (last_block) umma_commit_pair(tfull_bar, 3);
                            }
                        }
                        __syncwarp();
                        accumulate = 1;
                        if (++stage == DXP_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
                if (prof_on && lane == 0) { prof[2] = wp_f.acc; prof[6] = clock64() - body_t0; prof[7] = n_blk; }
            }
        } else if (warp >= 4) {
            // ---------------- epilogue (both CTAs): two warps per TMEM lane quadrant, one per N = 256 slice
            const int quad = warp & 3;
            const int nb = (warp - 4) >> 2;
            pr::EpiCtx ctx;
            ctx.staging = smem_u32(sStaging) + (warp - 4) * pr::STAGING_PER_WARP;
            ctx.lane = lane;
            ctx.prof = nullptr;
            pr::Stager stager(ctx);
            mbar_wait(tfull_bar, 0);
            tc_fence_after();
            const int row0 = m0 + quad * 32;
            if (row0 < p.B && nb < n_nb) {  // warp-uniform: otherwise batch padding / columns past the embedding width
                const uint32_t taddr = tmem_base + nb * 256 + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
                for (int cc = 0; cc < 256 / 32; ++cc) {
                    const int d0 = n0 + nb * 256 + cc * 32;
                    if (d0 >= p.D) break;
                    uint32_t v[32];
                    tmem_ld32(taddr + cc * 32, v);
                    tmem_ld_wait();
                    stager.acquire();
#pragma unroll
                    for (int g = 0; g < 8; ++g) stager.put(g, v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        if (p.part_rows > 0) tma_store_2d(&tmC, ctx.staging, d0, sp * p.part_rows + row0);
                        else tma_reduce_add_2d(&tmC, ctx.staging, d0, row0);  // rows >= B / columns >= D are clipped
                        bulk_commit();
                    }
                }
                if (p.part_rows > 0 && p.region_cnt != nullptr) dx_region_done(p, (t * 2 + rank) * 8 + (warp - 4), row0, n0 + nb * 256, lane);
                stager.drain();
            }
        }
    }

    // neither CTA may leave (or free TMEM) while the peer can still touch its barriers / shared memory / TMEM
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc_pair(tmem_base, 512);
}

constexpr int MAX_PAIRS = 96;

struct FusedParams {
    BwdDCpT<true>::Params dc;
    BwdDWpT<true>::Params dw;
    DXParams dx;
    int n_dc, n_dw, n_dx;  // CTA pairs per role
    int dx_pairs;          // 1: the dX role runs on CTA pairs (dx_pair_body), 0: on single CTAs (dx_body)
    // role (0 dC^T, 1 dW, 2 dX) and index inside the role of every CTA pair of the launch: the roles are
    // interleaved over the launch order so that each gets SMs from every GPC / both dies, whatever way the
    // hardware places consecutive clusters
    uint8_t role[MAX_PAIRS];
    uint8_t index[MAX_PAIRS];
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(pr::THREADS, 1)
bwd_fused_kernel(const __grid_constant__ CUtensorMap tm_w_k, const __grid_constant__ CUtensorMap tm_x_k,
                 const __grid_constant__ CUtensorMap tm_ring_out, const __grid_constant__ CUtensorMap tm_ring_k,
                 const __grid_constant__ CUtensorMap tm_xt_k, const __grid_constant__ CUtensorMap tm_dw_out,
                 const __grid_constant__ CUtensorMap tm_ring_mn, const __grid_constant__ CUtensorMap tm_w_mn,
                 const __grid_constant__ CUtensorMap tm_dx_out, const __grid_constant__ FusedParams prm) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    uint8_t* smem = smem_raw;
    const int pair = blockIdx.x >> 1;
    const int role = prm.role[pair];
    const int idx = prm.index[pair];
    if (role == 0) {
        pr::pair_gemm_body<BwdDCpT<true>>(tm_w_k, tm_x_k, tm_ring_out, prm.dc, BwdDCpT<true>::EXTRA_BYTES, smem, idx,
                                          prm.n_dc);
    } else if (role == 1) {
        pr::pair_gemm_body<BwdDWpT<true>>(tm_ring_k, tm_xt_k, tm_dw_out, prm.dw, BwdDWpT<true>::EXTRA_BYTES, smem, idx,
                                          prm.n_dw);
    } else if (prm.dx_pairs) {
        dx_pair_body(tm_ring_mn, tm_w_mn, tm_dx_out, prm.dx, smem, idx);
    } else {
        dx_body(tm_ring_mn, tm_w_mn, tm_dx_out, prm.dx, smem, idx * 2 + static_cast<int>(cluster_ctarank()));
    }
}

}  // namespace fz
}  // namespace ab
