// Resident-operand streaming GEMM core ("RS core") for the backward of the ArcFace head.
//
// Why a second core: the persistent core in gemm_core.cuh streams BOTH operands of every
// 128 x 256 x 64 step through shared memory (48 KB per 512 tensor cycles = 94 B/clk/SM), which is above
// what L2 can feed 148 SMs (~6.3-6.8 KB/clk chip-wide, ~45 B/clk/SM; B300_MICROARCH.md "LTS throughput
// cap").  Two of the three backward contractions have one SMALL operand that every class block reuses:
//   S^T  = What . Xhat^T     (B operand = a 128-row slice of Xhat,   K = D)
//   dWhat = dC^T . Xhat      (B operand = a 128-row slice of Xhat^T, K = batch)
// so the RS core parks that 128 x K slice in shared memory for the whole kernel (K <= 512: 128 KB) and
// streams only the class-block operand: 16 KB per 256 tensor cycles and each class block is fetched by the
// few CTAs that own the different resident slices at about the same time (merged in L2).
//
// CTA layout (384 threads):
//   warp 0     : TMA producer (one lane): resident slice once, then the streamed 128 x 64 k-blocks through a
//                4-deep mbarrier ring; also issues TMA L2 prefetches a few class blocks ahead (the streamed
//                operand comes from HBM)
//   warp 1     : MMA issuer (one lane): tcgen05.mma kind::f16, M = 128, N = 128, accumulators in TMEM,
//                four accumulator buffers (4 x 128 columns) so the epilogue runs up to 3 tiles behind
//   warp 2     : TMEM allocator
//   warps 4-11 : epilogue, two warps per TMEM lane quadrant (each takes 64 of the 128 columns), thread i of a
//                quadrant owns accumulator row i; output tiles leave through one 4 KB swizzled staging
//                buffer per warp and TMA stores
// Work split: CTA c keeps resident slice (c % n_res) and walks the class blocks i = c / n_res,
// + gridDim / n_res, ... (interleaved, so neighbouring CTAs stream neighbouring blocks).
//
// Reference counterpart: the two autograd matmuls of loss.backward() through arcface.py:47.
#pragma once
#include "ptx.cuh"

namespace ab {
namespace rs {

constexpr int BM = 128;  // streamed rows per tile (accumulator rows / TMEM lanes)
constexpr int BN = 128;  // resident rows (accumulator columns)
constexpr int BK = 64;   // bf16 elements per k-block: one 128-byte swizzle span
constexpr int STAGES = 4;
constexpr int ACC_BUFS = 4;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = (4 + EPI_WARPS) * 32;
constexpr int TILE_BYTES = BM * BK * 2;  // 16 KB: one k-block of either operand
constexpr int STAGING_PER_WARP = 4096;   // 32 rows x 128 B
constexpr int MAX_KBLOCKS = 8;           // resident slice <= 128 KB

struct Core {
    int kblocks;         // ceil(K / 64), <= MAX_KBLOCKS
    int m_blocks;        // 128-row blocks of the streamed operand handled by this launch
    int s_row0;          // TMA row coordinate of block 0 in the streamed tensor map
    int n_res;           // number of resident slices; gridDim.x is a multiple of it
    int prefetch_tiles;  // L2 prefetch distance in tiles of this CTA (0 = off)
};

struct EpiCtx {
    uint8_t* extra;          // policy-defined shared memory (filled by P::prologue)
    uint32_t staging;        // this warp's 4 KB staging buffer (shared-space address, 1024-aligned)
    const CUtensorMap* tmC;  // output tensor map
    int quad;                // TMEM lane quadrant 0..3
    int half;                // which 64 accumulator columns this warp owns (0 / 1)
    int lane;
    int res;                 // resident slice of this CTA
};

constexpr size_t smem_bytes(int kblocks, size_t extra_bytes) {
    return 1024 + static_cast<size_t>(kblocks + STAGES) * TILE_BYTES + EPI_WARPS * STAGING_PER_WARP +
           ((extra_bytes + 15) / 16) * 16 + (2 * STAGES + 2 * ACC_BUFS + 1) * 8 + 16;
}

// One staging buffer per epilogue warp: thread `lane` owns row `lane` (128 B = eight 16-byte chunks,
// swizzled like TMA SWIZZLE_128B so the writes are bank-conflict free); a box leaves with one TMA store.
struct Stager {
    uint32_t buf;
    int lane;
    __device__ Stager(const EpiCtx& c) : buf(c.staging), lane(c.lane) {}
    // the previous store must have finished READING the buffer before it is overwritten
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
    __device__ __forceinline__ void drain() const {
        if (lane == 0) bulk_wait<0>();
        __syncwarp();
    }
};

template <class P>
__global__ void __launch_bounds__(THREADS, 1)
rs_gemm_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmR,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ typename P::Params prm,
               const int extra_bytes) {
    const Core& co = prm.core;
    const int kblocks = co.kblocks;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sRes = smem;
    uint8_t* sStage = sRes + kblocks * TILE_BYTES;
    uint8_t* sStaging = sStage + STAGES * TILE_BYTES;
    uint8_t* sExtra = sStaging + EPI_WARPS * STAGING_PER_WARP;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sExtra + ((extra_bytes + 15) / 16) * 16);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + ACC_BUFS;
    uint64_t* res_bar = tempty_bar + ACC_BUFS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int res = blockIdx.x % co.n_res;
    const int grp = blockIdx.x / co.n_res;
    const int ngrp = gridDim.x / co.n_res;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmS);
        tma_prefetch_desc(&tmR);
        tma_prefetch_desc(&tmC);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < ACC_BUFS; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], EPI_WARPS);
        }
        mbar_init(res_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, ACC_BUFS * BN);
        tmem_relinquish();
    }
    P::prologue(prm, sExtra, threadIdx.x, res);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // whole warp walks the schedule (uniform control flow); one elected lane issues the TMA work
        if (elect_one()) {
            mbar_expect_tx(res_bar, kblocks * TILE_BYTES);
            for (int kb = 0; kb < kblocks; ++kb) tma_load_2d(sRes + kb * TILE_BYTES, &tmR, res_bar, kb * BK, res * BN);
        }
        __syncwarp();
        int stage = 0;
        uint32_t phase = 0;
        for (int i = grp; i < co.m_blocks; i += ngrp) {
            const int row = co.s_row0 + i * BM;
            const int ip = i + co.prefetch_tiles * ngrp;
            const bool pf = co.prefetch_tiles > 0 && ip < co.m_blocks;
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&full_bar[stage], TILE_BYTES);
                    tma_load_2d(sStage + stage * TILE_BYTES, &tmS, &full_bar[stage], kb * BK, row);
                    // the n_res CTAs that stream the same block share the prefetch work
                    if (pf && (kb % co.n_res) == res) tma_prefetch_2d(&tmS, kb * BK, co.s_row0 + ip * BM);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc_bf16(BM, BN, false, false);
        const uint32_t sStage_u32 = smem_u32(sStage);
        const uint64_t res_desc0 = make_smem_desc(smem_u32(sRes), 16, 1024);
        mbar_wait(res_bar, 0);
        tc_fence_after();
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int i = grp; i < co.m_blocks; i += ngrp) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BN;
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint64_t a_desc = make_smem_desc(sStage_u32 + stage * TILE_BYTES, 16, 1024);
                const uint64_t b_desc = res_desc0 + static_cast<uint64_t>(kb * (TILE_BYTES >> 4));
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < BK / 16; ++kk) {
                        // +2 in the address field = 32 bytes = 16 bf16 along K inside the swizzle span
                        umma_bf16(tmem_d, a_desc + 2 * kk, b_desc + 2 * kk, idesc, (kb | kk) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (kb == kblocks - 1) umma_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == ACC_BUFS) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        EpiCtx ctx;
        ctx.extra = sExtra;
        ctx.staging = smem_u32(sStaging) + (warp - 4) * STAGING_PER_WARP;
        ctx.tmC = &tmC;
        ctx.quad = warp & 3;  // a warp may only touch TMEM lanes 32 * (warp % 4) ...
        ctx.half = (warp - 4) >> 2;
        ctx.lane = lane;
        ctx.res = res;
        typename P::Epi epi(prm, ctx);
        int acc = 0;
        uint32_t acc_phase = 0;
        if (grp < co.m_blocks) epi.prefetch(grp);
        for (int i = grp; i < co.m_blocks; i += ngrp) {
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * BN + ctx.half * 64 + (static_cast<uint32_t>(ctx.quad * 32) << 16);
            // i_next lets the policy issue the NEXT tile's global loads before it works on this one
            epi.tile(i, i + ngrp < co.m_blocks ? i + ngrp : -1, taddr);  // returns after its last tcgen05.ld completed
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == ACC_BUFS) { acc = 0; acc_phase ^= 1; }
        }
        epi.finish();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, ACC_BUFS * BN);
}

#ifdef AB_CHECK_CUDA
template <class P>
static int32_t launch_rs(const CUtensorMap& tmS, const CUtensorMap& tmR, const CUtensorMap& tmC,
                         const typename P::Params& prm, int groups, int extra_bytes, cudaStream_t st) {
    const size_t smem = smem_bytes(prm.core.kblocks, extra_bytes);
    AB_REQUIRE(prm.core.kblocks >= 1 && prm.core.kblocks <= MAX_KBLOCKS && smem <= 227 * 1024, ARCFACE_B200_E_SHAPE,
               "resident-operand kernel: %d k-blocks / %zu bytes of shared memory do not fit", prm.core.kblocks, smem);
    AB_REQUIRE(groups >= 1 && prm.core.n_res >= 1, ARCFACE_B200_E_SHAPE, "empty grid");
    static bool configured[64] = {false};
    int dev = 0;
    AB_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        AB_CHECK_CUDA(cudaFuncSetAttribute(rs_gemm_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured[dev] = true;
    }
    rs_gemm_kernel<P><<<groups * prm.core.n_res, THREADS, smem, st>>>(tmS, tmR, tmC, prm, extra_bytes);
    AB_CHECK_CUDA(cudaGetLastError());
    return ARCFACE_B200_OK;
}
#endif

}  // namespace rs
}  // namespace ab
