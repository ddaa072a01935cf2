// K2 -- forward of the ArcFace head: cosine-logit GEMM on tcgen05 with a fused margin / scale /
// online-softmax / argmax epilogue (the B x C logit matrix never reaches HBM), plus the
// materialising variant used by forward_test / the lazy-logits object.
//
// Reference: arcface.py:47 (F.linear of the normalised operands), :58-61 (one-hot blend, * s),
// CrossEntropyLoss + argmax at the call sites (nlp_classifier_train.py:120-123).
//
// Orientation: batch rows on the accumulator rows (TMEM lanes), classes on the columns, so every
// epilogue thread owns one batch row and the running (max, sum-exp, argmax) is a thread-local scan.
#include "host_util.h"
#include "gemm_core.cuh"
#include "gemm_pair.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace ab {

constexpr float LOG2E = 1.4426950408889634f;

// ------------------------------------------------------------------ softmax-statistics epilogue
struct FwdStats {
    static constexpr int BLOCK_N = 256;
    static constexpr int STAGES = 4;
    static constexpr int M_SUB = 1;
    static constexpr int ACC_BUFS = 2;
    static constexpr bool STAGING = false;
    static constexpr bool A_MN = false;  // xhat [B][D]
    static constexpr bool B_MN = false;  // what [C][D]

    struct Params {
        int B, D;
        int C;                   // classes of this shard
        float s;
        const int* label_local;  // nullable: column excluded from the statistics (merged later in fp32)
        float* part_max;
        float* part_sum;
        int* part_arg;
        int m_tiles, n_tiles;
        int groups;       // class ranges; CTA (g, m) handles m-tile m of range g
        int tiles_per_g;  // ceil(n_tiles / groups)
    };

    __device__ static void prologue(const Params&, uint8_t*, int) {}

    struct Sched {
        int m_tile, n, n_end, kblocks;
        __device__ Sched(const Params& p, int cta, int) {
            m_tile = cta % p.m_tiles;
            const int g = cta / p.m_tiles;
            n = g * p.tiles_per_g;
            n_end = min(p.n_tiles, n + p.tiles_per_g);
            if (g >= p.groups) n_end = n;
            kblocks = (p.D + BLOCK_K - 1) / BLOCK_K;
        }
        __device__ bool next(Tile& t) {
            if (n >= n_end) return false;
            t.m0 = m_tile * BLOCK_M;
            t.n0 = n * BLOCK_N;
            t.ka0 = 0;
            t.kb0 = 0;
            t.kblocks = kblocks;
            t.aux = 0;
            ++n;
            return true;
        }
    };

    struct Epi {
        const Params& p;
        int row, g;
        bool active;
        float run_max, sum0, sum1, sum2, sum3;
        int run_arg;
        int lab;
        __device__ Epi(const Params& prm, const EpiCtx& c) : p(prm) {
            const int m_tile = c.cta % p.m_tiles;
            g = c.cta / p.m_tiles;
            row = m_tile * BLOCK_M + c.ew * 32 + c.lane;
            active = (g < p.groups) && (row < p.B);
            run_max = -INFINITY;
            sum0 = sum1 = sum2 = sum3 = 0.f;
            run_arg = 0;
            lab = -1;
            if (active && p.label_local != nullptr) lab = p.label_local[row];
        }
        __device__ void tile(const Tile& t, uint32_t taddr) {
            const float s = p.s;
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(taddr + c * 32, v);
                tmem_ld_wait();
                const int col0 = t.n0 + c * 32;
                if (col0 >= p.C) break;  // warp-uniform: whole chunk past the last class
                float z[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) z[j] = __uint_as_float(v[j]) * s;
                if (col0 + 32 > p.C) {  // warp-uniform tail
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (col0 + j >= p.C) z[j] = -INFINITY;
                }
                const int lr = lab - col0;
                if (lr >= 0 && lr < 32) {  // at most once per row: the label column is left out here and
                                           // merged exactly (fp32 margin logit) by finalize_rows
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j == lr) z[j] = -INFINITY;
                }
                float cm = z[0];
#pragma unroll
                for (int j = 1; j < 32; ++j) cm = fmaxf(cm, z[j]);
                if (cm == -INFINITY) continue;  // nothing but the excluded label / padding in this chunk
                if (cm > run_max) {  // strict: an equal later maximum never displaces the first one
                    int first = 31;
#pragma unroll
                    for (int j = 30; j >= 0; --j)
                        if (z[j] == cm) first = j;
                    run_arg = col0 + first;
                    const float f = ex2((run_max - cm) * LOG2E);  // ex2(-inf) = 0 on the first chunk
                    sum0 *= f; sum1 *= f; sum2 *= f; sum3 *= f;
                    run_max = cm;
                }
                const float mb = run_max * LOG2E;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    sum0 += ex2(fmaf(z[j + 0], LOG2E, -mb));
                    sum1 += ex2(fmaf(z[j + 1], LOG2E, -mb));
                    sum2 += ex2(fmaf(z[j + 2], LOG2E, -mb));
                    sum3 += ex2(fmaf(z[j + 3], LOG2E, -mb));
                }
            }
        }
        __device__ void finish() {
            if (!active) return;
            const int64_t o = static_cast<int64_t>(g) * p.B + row;
            p.part_max[o] = run_max;
            p.part_sum[o] = (sum0 + sum1) + (sum2 + sum3);
            p.part_arg[o] = run_arg;
        }
    };
};

// ------------------------------------------------------------------ softmax statistics on a CTA pair
// (gemm_pair.cuh) resident = 256 batch rows of Xhat (A operand, accumulator lanes: 128 rows per CTA), streamed =
// What rows (256 classes per tile = accumulator columns, each CTA loads 128 of them), K = D <= 512.  Each pair
// walks a contiguous class range; the two epilogue warps of a TMEM lane quadrant take 128 columns each and
// keep separate partial rows (slot = 2 * range + half), merged by combine_partials.
//
// NORM = true additionally runs K1 for the class weights INSIDE this kernel (arcface.py:47, F.normalize of the
// weight): sixteen helper warps per CTA read the fp32 rows, write bf16 what + 1/||w|| and publish one counter
// per 128-row block; the TMA producer of a CTA waits for the counter of the block it is about to fetch, so the
// bf16 rows are read back from L2 and the fp32 weights cross HBM exactly once per step for the forward.
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

#ifndef AB_FWD_AUX_WARPS
#define AB_FWD_AUX_WARPS 16
#endif
__device__ __forceinline__ float4 ldg_stream4_hint(const float* p, uint64_t pol) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}

template <bool NORM>
struct FwdStatsPairT : pr::PairDefaults {
#ifndef AB_FWD_STAGES
#define AB_FWD_STAGES 4
#endif
    static constexpr int STAGES = AB_FWD_STAGES;
    static constexpr bool STAGING = false;
    static constexpr bool RES_A = true;
    static constexpr int NROW = 2 * pr::ROWS;
    static constexpr int AUX_WARPS = NORM ? AB_FWD_AUX_WARPS : 0;
    // 16 helper warps: the CTA is launched with 72 registers per thread (896 threads) and the roles trade
    // registers with setmaxnreg (gemm_pair.cuh): producer / MMA / allocator warps 40, epilogue 96, helpers 64.
    // The helper loop is latency-bound per warp (global load -> shuffle reduction -> sqrt -> convert -> store), so
    // what raises its throughput is MORE warps each holding one row pair, not a deeper pipeline per warp.
#ifndef AB_FWD_AUX_REGS
#define AB_FWD_AUX_REGS 64
#endif
#ifndef AB_FWD_EPI_REGS
#define AB_FWD_EPI_REGS 96
#endif
    static constexpr bool WIDE = NORM && AB_FWD_AUX_WARPS >= 16;
    static constexpr int LOW_REGS = WIDE ? 40 : 0, EPI_REGS = WIDE ? AB_FWD_EPI_REGS : 0, AUX_REGS = WIDE ? AB_FWD_AUX_REGS : 0;
#ifndef AB_FWD_RUN
#define AB_FWD_RUN 8
#endif
    static constexpr int RUN = AB_FWD_RUN;  // consecutive rows a helper warp normalises between two counter updates

    struct Params {
        pr::Core core;
        int B;
        int C;
        float s;
        const int* label_local;
        float* part_max;
        float* part_sum;
        int* part_arg;
        // NORM only
        int D;
        const float* w;          // fp32 class weights [C][D]
        __nv_bfloat16* what;     // out: bf16 normalised weights [C][D]
        float* inv_nw;           // out: 1 / max(||w||, 1e-12)
        int* ready;              // [ceil(C / 128)] rows published per 128-row block (zeroed before the launch)
        int debug;               // measurements only: 1 = the GEMM does not wait for the helper warps
        int pf_dist;             // helper warps prefetch their row pair j + pf_dist into L2 (0 = off)
        int evict_first;         // 1: the fp32 weight stream is loaded with an L2 evict-first policy
    };

    __device__ static void prologue(const Params&, uint8_t*, int, int, int) {}

    __device__ static void acquire_tile(const Params& p, uint8_t*, int i, int rank, int lane) {
        if constexpr (NORM) {
            const int blk = i * 2 + rank;
            const int need = min(pr::ROWS, p.C - blk * pr::ROWS);
            if (need > 0 && p.debug != 1) {
                if (lane == 0) wait_counter_ge(p.ready + blk, need);
                __syncwarp();
                fence_proxy_async_all();  // the rows were written with ordinary stores, TMA reads them
            }
        }
    }

    // two rows of up to 512 floats per lane pair of chunks: lane l owns the 8-float chunks l and l + 32
    struct RowPair {
        float4 v[2][2][2];  // [row][chunk][half]
    };
    __device__ static __forceinline__ void load_pair(const Params& p, int64_t r0, int lane, RowPair& rp, uint64_t pol) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int64_t row = min(r0 + r, static_cast<int64_t>(p.C) - 1);  // clamp: the tail re-reads the last row
            const float* src = p.w + row * p.D;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                const int d = (lane + 32 * ch) * 8;
                if (d < p.D) {
                    if (pol != 0) {
                        rp.v[r][ch][0] = ldg_stream4_hint(src + d, pol);
                        rp.v[r][ch][1] = ldg_stream4_hint(src + d + 4, pol);
                    } else {
                        rp.v[r][ch][0] = ldg_stream4(src + d);
                        rp.v[r][ch][1] = ldg_stream4(src + d + 4);
                    }
                } else {
                    rp.v[r][ch][0] = make_float4(0.f, 0.f, 0.f, 0.f);
                    rp.v[r][ch][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
    }
    __device__ static __forceinline__ void store_pair(const Params& p, int64_t r0, int lane, const RowPair& rp) {
        float ss[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float a = 0.f;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {  // same order as normalize_cast_kernel: bit-identical norms
                const float4 x = rp.v[r][ch][0], y = rp.v[r][ch][1];
                a += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
                a += y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w;
            }
            ss[r] = a;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ss[0] += __shfl_xor_sync(0xffffffffu, ss[0], o);
            ss[1] += __shfl_xor_sync(0xffffffffu, ss[1], o);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int64_t row = r0 + r;
            if (row >= p.C) break;
            const float inv = 1.0f / fmaxf(sqrtf(ss[r]), 1e-12f);
            if (lane == 0) p.inv_nw[row] = inv;
            __nv_bfloat16* dst = p.what + row * p.D;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                const int d = (lane + 32 * ch) * 8;
                if (d < p.D) {
                    const float4 x = rp.v[r][ch][0], y = rp.v[r][ch][1];
                    uint4 o;
                    o.x = pack_bf16x2(x.x * inv, x.y * inv);
                    o.y = pack_bf16x2(x.z * inv, x.w * inv);
                    o.z = pack_bf16x2(y.x * inv, y.y * inv);
                    o.w = pack_bf16x2(y.z * inv, y.w * inv);
#ifdef AB_FWD_WHAT_EVICT_LAST
                    // experiment: ask L2 to keep the normalised rows until the GEMM's TMA has read them back
                    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(dst + d), "r"(o.x),
                                 "r"(o.y), "r"(o.z), "r"(o.w), "l"(l2_policy_evict_last())
                                 : "memory");
#else
                    *reinterpret_cast<uint4*>(dst + d) = o;
#endif
                }
            }
        }
    }

    // helper warp u of nu: runs u, u + nu, ... of RUN consecutive rows, in the order the GEMM consumes them.
    // The warp's rows form one sequence of row pairs j = 0, 1, ...; pair j + 1 is loading into registers while
    // pair j is reduced and stored.  (Measured: a 4-deep register pipeline fed by setmaxnreg, and bulk L2
    // prefetches ahead of the loads, were both slower than this -- the stage is bound by the fence + counter
    // update that ends every run, not by the loads in flight.)
    static constexpr int PPR = RUN / 2;  // pairs per run

    // 512 < D <= 1024 (the RoBERTa-large head): the same 32 registers hold ONE row of up to 1024 floats -- lane l owns
    // the 8-float chunks l, l + 32, l + 64, l + 96 -- so the bytes a warp keeps in flight (4 KB) and the summation
    // order (normalize_cast_kernel<4>'s) are those of the row-pair loop below.
    __device__ static void aux_row1024(const Params& p, int u, int nu, int lane) {
        const int64_t n_runs = (static_cast<int64_t>(p.C) + RUN - 1) / RUN;
        for (int64_t run = u; run < n_runs; run += nu) {
            const int64_t r0 = run * RUN;
            int rows = 0;
#pragma unroll 1
            for (int k = 0; k < RUN; ++k) {
                const int64_t row = r0 + k;
                if (row >= p.C) break;
                ++rows;
                const float* src = p.w + row * p.D;
                float4 v[4][2];
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    const int d = (lane + 32 * ch) * 8;
                    if (d < p.D) {
                        v[ch][0] = ldg_stream4(src + d);
                        v[ch][1] = ldg_stream4(src + d + 4);
                    } else {
                        v[ch][0] = make_float4(0.f, 0.f, 0.f, 0.f);
                        v[ch][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                float ss = 0.f;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    const int d = (lane + 32 * ch) * 8;
                    if (d < p.D) {   // same guard as normalize_cast_kernel: bit-identical sums
                        const float4 x = v[ch][0], y = v[ch][1];
                        ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
                        ss += y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
                const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
                if (lane == 0) p.inv_nw[row] = inv;
                __nv_bfloat16* dst = p.what + row * p.D;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    const int d = (lane + 32 * ch) * 8;
                    if (d < p.D) {
                        const float4 x = v[ch][0], y = v[ch][1];
                        uint4 o;
                        o.x = pack_bf16x2(x.x * inv, x.y * inv);
                        o.y = pack_bf16x2(x.z * inv, x.w * inv);
                        o.z = pack_bf16x2(y.x * inv, y.y * inv);
                        o.w = pack_bf16x2(y.z * inv, y.w * inv);
                        *reinterpret_cast<uint4*>(dst + d) = o;
                    }
                }
            }
            __threadfence();
            __syncwarp();
            if (lane == 0 && rows > 0) red_relaxed_gpu_add(p.ready + (r0 / pr::ROWS), rows);
        }
    }

    __device__ static void aux(const Params& p, int u, int nu, int lane) {
        if constexpr (NORM) {
#ifndef AB_FWD_LEAN
            if (p.D > 512) {
                aux_row1024(p, u, nu, lane);
                return;
            }
#endif
            const int64_t n_runs = (static_cast<int64_t>(p.C) + RUN - 1) / RUN;
            if (u >= n_runs) return;
            const int64_t n_pairs = ((n_runs - u + nu - 1) / nu) * PPR;  // of this warp (even)
            auto row_of = [&](int64_t j) { return (u + (j / PPR) * nu) * RUN + (j % PPR) * 2; };
#ifdef AB_FWD_LEAN
            const uint64_t pol = 0ull;
#else
            const uint64_t pol = p.evict_first ? l2_policy_evict_first() : 0ull;
#endif
            auto fetch = [&](int64_t j, RowPair& rp) {
                if (j < n_pairs) load_pair(p, row_of(j), lane, rp, pol);
            };
            auto retire = [&](int64_t j, const RowPair& rp) {
                const int64_t r = row_of(j);
                store_pair(p, r, lane, rp);
                if (j % PPR == PPR - 1) {  // last pair of a run: publish its rows
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) {
                        const int64_t r0 = r - (RUN - 2);
                        const int rows = static_cast<int>(min(static_cast<int64_t>(RUN), p.C - r0));
                        if (rows > 0) red_relaxed_gpu_add(p.ready + (r0 / pr::ROWS), rows);
                    }
                }
            };
            // DRAM -> L2 prefetch of the pair `pf_dist` ahead: the bytes in flight beyond the two register pairs
            // live in L2, so the loads below see L2 latency instead of queued-DRAM latency
#ifdef AB_FWD_LEAN
            const int pfd = 0;
#else
            const int pfd = p.pf_dist;
#endif
            const uint32_t pair_bytes = static_cast<uint32_t>(2 * p.D * sizeof(float));
            auto ahead = [&](int64_t j) {
                if (pfd > 0 && lane == 0 && j + pfd < n_pairs) {
                    const int64_t r = row_of(j + pfd);
                    if (r + 2 <= p.C) bulk_prefetch_l2(p.w + r * p.D, pair_bytes);
                }
            };
            if (pfd > 0 && lane == 0)
                for (int64_t j = 2; j < pfd && j < n_pairs; ++j) {
                    const int64_t r = row_of(j);
                    if (r + 2 <= p.C) bulk_prefetch_l2(p.w + r * p.D, pair_bytes);
                }
            if constexpr (WIDE) {
                RowPair a;
#ifdef AB_FWD_LATEPUB
                // the fence + counter update that ends a run is issued AFTER the next pair's loads: its latency
                // (every store of the run acknowledged) overlaps the loads' instead of preceding them
                int64_t pending = -1;
                auto publish = [&](int64_t jl) {
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) {
                        const int64_t r0 = row_of(jl) - (RUN - 2);
                        const int rows = static_cast<int>(min(static_cast<int64_t>(RUN), p.C - r0));
                        if (rows > 0) red_relaxed_gpu_add(p.ready + (r0 / pr::ROWS), rows);
                    }
                };
                for (int64_t j = 0; j < n_pairs; ++j) {
                    fetch(j, a);
                    ahead(j);
                    if (pending >= 0) { publish(pending); pending = -1; }
                    store_pair(p, row_of(j), lane, a);
                    if (j % PPR == PPR - 1) pending = j;
                }
                if (pending >= 0) publish(pending);
#else
                for (int64_t j = 0; j < n_pairs; ++j) {
                    fetch(j, a);
                    ahead(j);
                    retire(j, a);
                }
#endif
                return;
            }
            RowPair a, b;
            fetch(0, a);
            for (int64_t j = 0; j < n_pairs; j += 2) {
                fetch(j + 1, b);
                ahead(j);
                retire(j, a);
                fetch(j + 2, a);
                ahead(j + 1);
                retire(j + 1, b);
            }
        }
    }

    struct Epi {
        const Params& p;
        int row, slot, half;
        bool active;
        float run_max, sum0, sum1, sum2, sum3;
        int run_arg;
        int lab;
        __device__ Epi(const Params& prm, const pr::EpiCtx& c) : p(prm) {
            row = c.res * NROW + c.rank * pr::ROWS + c.quad * 32 + c.lane;
            slot = c.grp * 2 + c.half;
            half = c.half;
            active = row < p.B;
            run_max = -INFINITY;
            sum0 = sum1 = sum2 = sum3 = 0.f;
            run_arg = 0;
            lab = -1;
            if (active && p.label_local != nullptr) lab = p.label_local[row];
        }
        __device__ void prefetch(int) {}
        __device__ void tile(int i, int, uint32_t taddr) {
            const float s = p.s;
            const int n0 = i * NROW + half * 128;  // first class of this warp's 128 columns
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                tmem_ld32(taddr + c * 32, v);
                tmem_ld_wait();
                const int col0 = n0 + c * 32;
                if (col0 >= p.C) break;  // warp-uniform: whole chunk past the last class
                float z[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) z[j] = __uint_as_float(v[j]) * s;
                if (col0 + 32 > p.C) {  // warp-uniform tail
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (col0 + j >= p.C) z[j] = -INFINITY;
                }
                const int lr = lab - col0;
                if (lr >= 0 && lr < 32) {  // the label column is merged exactly (fp32) by finalize_rows
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j == lr) z[j] = -INFINITY;
                }
                float cm = z[0];
#pragma unroll
                for (int j = 1; j < 32; ++j) cm = fmaxf(cm, z[j]);
                if (cm == -INFINITY) continue;
                if (cm > run_max) {  // strict: an equal later maximum never displaces the first one
                    int first = 31;
#pragma unroll
                    for (int j = 30; j >= 0; --j)
                        if (z[j] == cm) first = j;
                    run_arg = col0 + first;
                    const float f = ex2((run_max - cm) * LOG2E);
                    sum0 *= f; sum1 *= f; sum2 *= f; sum3 *= f;
                    run_max = cm;
                }
                const float mb = run_max * LOG2E;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    sum0 += ex2(fmaf(z[j + 0], LOG2E, -mb));
                    sum1 += ex2(fmaf(z[j + 1], LOG2E, -mb));
                    sum2 += ex2(fmaf(z[j + 2], LOG2E, -mb));
                    sum3 += ex2(fmaf(z[j + 3], LOG2E, -mb));
                }
            }
        }
        __device__ void finish() {
            if (!active) return;
            const int64_t o = static_cast<int64_t>(slot) * p.B + row;
            p.part_max[o] = run_max;
            p.part_sum[o] = (sum0 + sum1) + (sum2 + sum3);
            p.part_arg[o] = run_arg;
        }
    };
};

using FwdStatsP = FwdStatsPairT<false>;
using FwdStatsPN = FwdStatsPairT<true>;

// ------------------------------------------------------------------ materialising epilogue
struct FwdLogits {
    static constexpr int BLOCK_N = 256;
    static constexpr int STAGES = 4;
    static constexpr int M_SUB = 1;
    static constexpr int ACC_BUFS = 2;
    static constexpr bool STAGING = false;
    static constexpr bool A_MN = false;
    static constexpr bool B_MN = false;

    struct Params {
        int B, D;
        int C;
        float scale;
        const float* z_label;
        const int* label_local;
        float* out;
        int64_t ld_out;
        int m_tiles, n_tiles;
    };

    __device__ static void prologue(const Params&, uint8_t*, int) {}

    struct Sched {
        int idx, total, step, m_tiles, kblocks;
        __device__ Sched(const Params& p, int cta, int ncta) {
            idx = cta;
            step = ncta;
            m_tiles = p.m_tiles;
            total = p.m_tiles * p.n_tiles;
            kblocks = (p.D + BLOCK_K - 1) / BLOCK_K;
        }
        __device__ bool next(Tile& t) {
            if (idx >= total) return false;
            t.m0 = (idx % m_tiles) * BLOCK_M;
            t.n0 = (idx / m_tiles) * BLOCK_N;
            t.ka0 = 0;
            t.kb0 = 0;
            t.kblocks = kblocks;
            t.aux = 0;
            idx += step;
            return true;
        }
    };

    struct Epi {
        const Params& p;
        int ew, lane;
        __device__ Epi(const Params& prm, const EpiCtx& c) : p(prm), ew(c.ew), lane(c.lane) {}
        __device__ void tile(const Tile& t, uint32_t taddr) {
            const int row = t.m0 + ew * 32 + lane;
            const bool rv = row < p.B;
            int lab = -1;
            float zl = 0.f;
            if (rv && p.label_local != nullptr) {
                lab = p.label_local[row];
                zl = p.z_label[row];
            }
            float* orow = p.out + static_cast<int64_t>(rv ? row : 0) * p.ld_out;
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(taddr + c * 32, v);
                tmem_ld_wait();
                const int col0 = t.n0 + c * 32;
                if (col0 >= p.C) break;
                if (rv) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int col = col0 + j;
                        if (col < p.C) orow[col] = (col == lab) ? zl : __uint_as_float(v[j]) * p.scale;
                    }
                }
            }
        }
        __device__ void finish() {}
    };
};

static void fwd_partition(int B, int64_t C, int nsm, int* m_tiles, int* n_tiles, int* groups, int* per) {
    *m_tiles = (B + BLOCK_M - 1) / BLOCK_M;
    *n_tiles = static_cast<int>((C + FwdStats::BLOCK_N - 1) / FwdStats::BLOCK_N);
    int g = nsm / *m_tiles;
    if (g < 1) g = 1;
    if (g > *n_tiles) g = *n_tiles;
    *per = (*n_tiles + g - 1) / g;
    *groups = (*n_tiles + *per - 1) / *per;  // no empty class range
}

// CTA-pair forward: whenever the device has a pair of SMs.  D <= 512 parks the 256-row Xhat slice in shared memory,
// wider embeddings stream both operands (gemm_pair.cuh, Core::stream_both).  Diagnostic builds:
// ARCFACE_B200_FWD_IMPL=generic forces the one-CTA streaming kernel (A/B measurements).
static bool fwd_use_pairs(int D, int nsm) {
    const char* v = diag_env("ARCFACE_B200_FWD_IMPL");
    if (v != nullptr && strcmp(v, "generic") == 0) return false;
    if (v != nullptr && strcmp(v, "pairs") == 0) return nsm >= 2;
    // D <= 512: the batch slice is parked in shared memory and the pair kernel wins by 13-19 %.  Wider embeddings stream
    // both operands, and there the one-CTA streaming kernel measured faster at every BASELINE width (forward of configs
    // 2 / 3-shard / 4: 0.248 / 0.219 / 0.463 ms against 0.268 / 0.223 / 0.476 ms, profiles/r2_dx_pair_exp.log)
    return nsm >= 2 && D <= 512;
}
// The in-kernel weight normaliser holds a row pair (D <= 512) or one row (D <= 1024) in registers.  Wider rows run K1
// as its own launch: a one-row-at-a-time, two-pass helper (second pass from L2) was measured at 0.52 / 0.35 / 0.80 ms for
// the forward of BASELINE configs 2 / 3-shard / 4 against 0.27 / 0.22 / 0.52 ms for K1 + K2 as two launches -- sixteen
// helper warps with one row in flight each do not keep enough bytes in flight.
static bool fwd_norm_in_kernel(int D) {
    if (const char* v = diag_env("ARCFACE_B200_FWD_NORM_MAXD")) return D <= atoi(v);
    // A register row pair covers D <= 512.  The one-register-row variant for D <= 1024 (aux_row1024) is correct and
    // bit-identical but measured SLOWER than K1 + K2 as two launches (0.237 vs 0.214 ms for one rank's shard of the
    // RoBERTa-large head, 0.183 vs 0.166 ms at D = 768): with both GEMM operands streaming, the helper warps' loads
    // and the TMA traffic get in each other's way.  It stays reachable through the diagnostic knob only.
    return D <= 512;
}
static void fwd_pair_partition(int B, int64_t C, int nsm, int* n_res, int* groups, int* s_blocks) {
    *n_res = (B + FwdStatsP::NROW - 1) / FwdStatsP::NROW;
    *s_blocks = static_cast<int>((C + FwdStatsP::NROW - 1) / FwdStatsP::NROW);
    int g = (nsm / 2) / *n_res;
    if (g < 1) g = 1;
    if (g > *s_blocks) g = *s_blocks;
    *groups = g;
}

}  // namespace ab

using namespace ab;

static int32_t check_gemm_shape(const char* who, int32_t B, int32_t D, int64_t C) {
    AB_REQUIRE(B >= 1 && B <= ARCFACE_B200_MAX_BATCH, ARCFACE_B200_E_SHAPE, "%s: B=%d outside [1, %d]", who, B,
               ARCFACE_B200_MAX_BATCH);
    AB_REQUIRE(D >= 8 && D % 8 == 0, ARCFACE_B200_E_SHAPE, "%s: D=%d must be a positive multiple of 8", who, D);
    AB_REQUIRE(C >= 1 && C <= (1ll << 30), ARCFACE_B200_E_SHAPE, "%s: C_local=%lld outside [1, 2^30]", who,
               (long long)C);
    return ARCFACE_B200_OK;
}

template <bool NORM>
static int32_t launch_fwd_pairs(const uint16_t* xhat, const uint16_t* what, const float* w, float* inv_nw, int* ready,
                                const int32_t* label_local, int32_t B, int32_t D, int64_t C_local, float s,
                                float* part_max, float* part_sum, int32_t* part_arg, int32_t n_parts, cudaStream_t st) {
    using P = FwdStatsPairT<NORM>;
    typename P::Params p;
    int groups;
    fwd_pair_partition(B, C_local, sm_count(), &p.core.n_res, &groups, &p.core.s_blocks);
    AB_REQUIRE(n_parts == 2 * groups, ARCFACE_B200_E_WORKSPACE, "forward_stats: n_parts=%d, expected %d", n_parts,
               2 * groups);
    pr::core_set_k<P>(p.core, D, 0);
    p.core.s_row0 = 0;
    // NORM: the helper warps publish rows in ascending order, so the pairs walk the tiles interleaved
    // (tile t at step t / groups); otherwise every pair takes a contiguous class range
    p.core.contiguous = NORM ? 0 : 1;
    p.core.prefetch_tiles = NORM ? 0 : 2;  // NORM: the rows come out of L2 anyway (just written there)
    p.B = B; p.C = static_cast<int>(C_local); p.s = s;
    p.label_local = label_local;
    p.part_max = part_max; p.part_sum = part_sum; p.part_arg = part_arg;
    p.D = D; p.w = w; p.what = reinterpret_cast<__nv_bfloat16*>(const_cast<uint16_t*>(what));
    p.inv_nw = inv_nw; p.ready = ready;
    p.debug = 0;
    if (const char* dbg = diag_env("ARCFACE_B200_FWD_DEBUG")) p.debug = atoi(dbg);
    if (p.debug == 2) p.core.s_blocks = 0;  // measurements only: helper warps alone
    p.pf_dist = 0;
    if (const char* pf = diag_env("ARCFACE_B200_FWD_PF")) p.pf_dist = atoi(pf);
    p.evict_first = 0;
    if (const char* ef = diag_env("ARCFACE_B200_FWD_EVICT")) p.evict_first = atoi(ef);
    CUtensorMap tmS, tmR;
    if (int32_t rc = make_tmap_kmajor(&tmS, what, D, C_local, D, pr::ROWS)) return rc;
    if (int32_t rc = make_tmap_kmajor(&tmR, xhat, D, B, D, pr::ROWS)) return rc;
#ifdef ARCFACE_B200_DIAG
    // diagnostic build only (ARCFACE_B200_FWD_PROF=1): where each role waits; allocates, synchronises, prints to stderr
    static unsigned long long* prof_dev = nullptr;
    const char* pe = diag_env("ARCFACE_B200_FWD_PROF");
    const bool prof = pe != nullptr && atoi(pe) == 1;
    const int n_cta = 2 * groups * p.core.n_res;
    if (prof) {
        if (prof_dev == nullptr) AB_CHECK_CUDA(cudaMalloc(&prof_dev, 512 * 16 * sizeof(unsigned long long)));
        AB_CHECK_CUDA(cudaMemsetAsync(prof_dev, 0, 512 * 16 * sizeof(unsigned long long), st));
        p.core.prof = prof_dev;
        p.core.prof_cta = 0;
    }
#endif
    const int32_t rc = pr::launch_pair<P>(tmS, tmR, tmS, p, groups, 0, st);
#ifdef ARCFACE_B200_DIAG
    if (prof && rc == ARCFACE_B200_OK) {
        static unsigned long long host[512 * 16];
        AB_CHECK_CUDA(cudaStreamSynchronize(st));
        AB_CHECK_CUDA(cudaMemcpy(host, prof_dev, sizeof(host), cudaMemcpyDeviceToHost));
        double a[8] = {0};
        int n = 0;
        for (int c = 0; c < n_cta && c < 512; ++c) {
            for (int k = 0; k < 8; ++k) a[k] += static_cast<double>(host[c * 16 + k]);
            ++n;
        }
        // [2], [3] exist in leader CTAs only (n / 2 of them)
        fprintf(stderr, "[fwd prof] ctas=%d tiles/cta=%.1f body=%.0f cyc | producer: helper rows %.0f, free stage %.0f | "
                        "MMA: operands %.0f, free accumulator %.0f | epilogue: accumulator %.0f, inside tile %.0f\n",
                n, a[7] / n, a[6] / n, a[0] / n, a[1] / n, a[2] / (n / 2), a[3] / (n / 2), a[4] / n, a[5] / n);
    }
#endif
    return rc;
}

extern "C" int32_t arcface_b200_forward_parts(int32_t B, int32_t D, int64_t C_local, int32_t* n_parts) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(n_parts, ARCFACE_B200_E_ARG, "forward_parts: null pointer");
    if (int32_t rc = check_gemm_shape("forward_parts", B, D, C_local)) return rc;
    if (fwd_use_pairs(D, sm_count())) {
        int n_res, g, sb;
        fwd_pair_partition(B, C_local, sm_count(), &n_res, &g, &sb);
        *n_parts = 2 * g;
        return ARCFACE_B200_OK;
    }
    int mt, nt, g, per;
    fwd_partition(B, C_local, sm_count(), &mt, &nt, &g, &per);
    *n_parts = g;
    return ARCFACE_B200_OK;
}

extern "C" int32_t arcface_b200_forward_stats(const uint16_t* xhat, const uint16_t* what,
                                              const int32_t* label_local, int32_t B, int32_t D, int64_t C_local,
                                              float s, float* part_max, float* part_sum, int32_t* part_arg,
                                              int32_t n_parts, void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(xhat && what && part_max && part_sum && part_arg, ARCFACE_B200_E_ARG, "forward_stats: null pointer");
    AB_REQUIRE(s > 0.f, ARCFACE_B200_E_ARG, "forward_stats: scale s must be positive");
    if (int32_t rc = check_gemm_shape("forward_stats", B, D, C_local)) return rc;
    if (fwd_use_pairs(D, sm_count()))
        return launch_fwd_pairs<false>(xhat, what, nullptr, nullptr, nullptr, label_local, B, D, C_local, s, part_max,
                                       part_sum, part_arg, n_parts, static_cast<cudaStream_t>(stream));
    FwdStats::Params p;
    fwd_partition(B, C_local, sm_count(), &p.m_tiles, &p.n_tiles, &p.groups, &p.tiles_per_g);
    AB_REQUIRE(n_parts == p.groups, ARCFACE_B200_E_WORKSPACE, "forward_stats: n_parts=%d, expected %d", n_parts,
               p.groups);
    p.B = B; p.D = D; p.C = static_cast<int>(C_local); p.s = s;
    p.label_local = label_local;
    p.part_max = part_max; p.part_sum = part_sum; p.part_arg = part_arg;
    CUtensorMap tmA, tmB;
    if (int32_t rc = make_tmap_kmajor(&tmA, xhat, D, B, D, BLOCK_M)) return rc;
    if (int32_t rc = make_tmap_kmajor(&tmB, what, D, C_local, D, FwdStats::BLOCK_N)) return rc;
    return launch_gemm<FwdStats>(tmA, tmB, tmA, p, p.groups * p.m_tiles, 0, static_cast<cudaStream_t>(stream));
}

extern "C" int32_t arcface_b200_forward_fused_workspace_bytes(int32_t B, int32_t D, int64_t C_local, size_t* bytes) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(bytes, ARCFACE_B200_E_ARG, "forward_fused_workspace_bytes: null pointer");
    if (int32_t rc = check_gemm_shape("forward_fused_workspace_bytes", B, D, C_local)) return rc;
    *bytes = static_cast<size_t>((C_local + pr::ROWS - 1) / pr::ROWS) * sizeof(int) + 256;
    return ARCFACE_B200_OK;
}

extern "C" int32_t arcface_b200_forward_stats_fused(const uint16_t* xhat, const float* w, const int32_t* label_local,
                                                    int32_t B, int32_t D, int64_t C_local, float s, uint16_t* what,
                                                    float* inv_nw, float* part_max, float* part_sum,
                                                    int32_t* part_arg, int32_t n_parts, void* workspace,
                                                    size_t workspace_bytes, void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(xhat && w && what && inv_nw && part_max && part_sum && part_arg && workspace, ARCFACE_B200_E_ARG,
               "forward_stats_fused: null pointer");
    AB_REQUIRE(s > 0.f, ARCFACE_B200_E_ARG, "forward_stats_fused: scale s must be positive");
    if (int32_t rc = check_gemm_shape("forward_stats_fused", B, D, C_local)) return rc;
    AB_REQUIRE(aligned16(w) && aligned16(what) && aligned16(workspace), ARCFACE_B200_E_LAYOUT,
               "forward_stats_fused: pointers must be 16-byte aligned");
    const char* impl = diag_env("ARCFACE_B200_FWD_IMPL");
    const bool split = impl != nullptr && strcmp(impl, "split") == 0;  // A/B: K1 and K2 as two launches
    if (!fwd_use_pairs(D, sm_count()) || !fwd_norm_in_kernel(D) || split) {
        // shapes the in-kernel normaliser does not cover: the same two steps as separate launches
        if (int32_t rc = arcface_b200_normalize_cast(w, C_local, D, what, inv_nw, nullptr, 0, stream)) return rc;
        return arcface_b200_forward_stats(xhat, what, label_local, B, D, C_local, s, part_max, part_sum, part_arg,
                                          n_parts, stream);
    }
    const size_t flag_bytes = static_cast<size_t>((C_local + pr::ROWS - 1) / pr::ROWS) * sizeof(int);
    AB_REQUIRE(workspace_bytes >= flag_bytes, ARCFACE_B200_E_WORKSPACE, "forward_stats_fused: workspace %zu < required %zu",
               workspace_bytes, flag_bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    AB_CHECK_CUDA(cudaMemsetAsync(workspace, 0, flag_bytes, st));
    return launch_fwd_pairs<true>(xhat, what, w, inv_nw, static_cast<int*>(workspace), label_local, B, D, C_local, s,
                                  part_max, part_sum, part_arg, n_parts, st);
}

extern "C" int32_t arcface_b200_logits(const uint16_t* xhat, const uint16_t* what, const float* z_label,
                                       const int32_t* label_local, int32_t B, int32_t D, int64_t C_local, float scale,
                                       float* out, int64_t ld_out, void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(xhat && what && out, ARCFACE_B200_E_ARG, "logits: null pointer");
    AB_REQUIRE((z_label == nullptr) == (label_local == nullptr), ARCFACE_B200_E_ARG,
               "logits: z_label and label_local must both be given or both be null");
    AB_REQUIRE(ld_out >= C_local, ARCFACE_B200_E_LAYOUT, "logits: ld_out < C_local");
    if (int32_t rc = check_gemm_shape("logits", B, D, C_local)) return rc;
    FwdLogits::Params p;
    p.B = B; p.D = D; p.C = static_cast<int>(C_local); p.scale = scale;
    p.z_label = z_label; p.label_local = label_local; p.out = out; p.ld_out = ld_out;
    p.m_tiles = (B + BLOCK_M - 1) / BLOCK_M;
    p.n_tiles = static_cast<int>((C_local + FwdLogits::BLOCK_N - 1) / FwdLogits::BLOCK_N);
    CUtensorMap tmA, tmB;
    if (int32_t rc = make_tmap_kmajor(&tmA, xhat, D, B, D, BLOCK_M)) return rc;
    if (int32_t rc = make_tmap_kmajor(&tmB, what, D, C_local, D, FwdLogits::BLOCK_N)) return rc;
    const int64_t total = static_cast<int64_t>(p.m_tiles) * p.n_tiles;
    const int grid = static_cast<int>(total < sm_count() ? total : sm_count());
    return launch_gemm<FwdLogits>(tmA, tmB, tmA, p, grid, 0, static_cast<cudaStream_t>(stream));
}
