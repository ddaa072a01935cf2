// Persistent warp-specialised tcgen05 GEMM core shared by every contraction of the ArcFace head
// (cosine logits, dW, dX).  One CTA per SM; inside a CTA:
//   warp 0   : TMA producer  (one lane)  global -> 128B-swizzled smem ring, mbarrier complete_tx
//   warp 1   : MMA issuer    (one lane)  tcgen05.mma kind::f16 bf16 x bf16 -> fp32 in TMEM
//   warp 2   : TMEM allocator / deallocator
//   warps 4-7: epilogue      (tcgen05.ld 32x32b, thread i owns accumulator row i of the 128-row tile)
// The accumulator is double-buffered in TMEM (2 x BLOCK_N columns) so the epilogue of tile i
// overlaps the MMAs of tile i+1.  A policy class supplies the tile schedule and the epilogue.
// Epilogues that write tiles stage them through swizzled shared memory and leave through TMA stores
// (StoreStager), so every global write is a full 128-byte line.
//
// Replaces, in the reference, the cuBLAS/MKL SGEMMs behind F.linear (arcface.py:47) and the two
// autograd matmuls of loss.backward().
#pragma once
#include "ptx.cuh"

namespace ab {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // bf16 elements: one 128-byte swizzle span
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 256;
constexpr int A_SUB_BYTES = BLOCK_M * BLOCK_K * 2;  // one 128-row sub-tile of operand A
constexpr int STAGING_BYTES = 4 * 2 * 4096;  // 4 epilogue warps x 2 buffers x (32 rows x 128 B)

struct Tile {
    int m0;       // first accumulator row (TMA coordinate of operand A along M)
    int n0;       // first accumulator column (TMA coordinate of operand B along N)
    int ka0;      // first K element for operand A
    int kb0;      // first K element for operand B
    int kblocks;  // number of BLOCK_K slices (>= 1)
    int aux;      // policy-defined
};

// What an epilogue object gets to see.
struct EpiCtx {
    uint8_t* extra;          // policy-defined shared memory (filled by P::prologue)
    uint8_t* staging;        // STAGING_BYTES of 1024-aligned shared memory (only if P::STAGING)
    const CUtensorMap* tmC;  // output tensor map (only if P::STAGING)
    int ew;                  // epilogue warp 0..3 == TMEM lane quadrant
    int lane;
    int cta;
};

// A policy may stack M_SUB 128-row sub-tiles that share every B k-block (M_SUB x BLOCK_N accumulator
// columns each) and chooses how many accumulator sets (ACC_BUFS) live in TMEM.
template <class P>
struct SmemLayout {
    static constexpr int A_STAGE_BYTES = P::M_SUB * A_SUB_BYTES;
    static constexpr int B_STAGE_BYTES = P::BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
};

// Dynamic shared memory needed by gemm_kernel<P> (includes 1 KB of alignment slack).
template <class P>
constexpr size_t gemm_smem_bytes(size_t extra_bytes) {
    return 1024 + static_cast<size_t>(P::STAGES) * SmemLayout<P>::STAGE_BYTES + (P::STAGING ? STAGING_BYTES : 0) +
           ((extra_bytes + 15) / 16) * 16 + (2 * P::STAGES + 2 * P::ACC_BUFS) * 8 + 16;
}

// Per-epilogue-warp output staging: two 4 KB buffers (32 rows x 128 B, TMA SWIZZLE_128B layout).  Thread
// `lane` owns row `lane`; a block is written as eight 16-byte chunks per row (bank-conflict free thanks to
// the swizzle) and leaves through one TMA store (or fp32 reduce-add) of a {128 B x 32 rows} box.
struct StoreStager {
    uint32_t base;
    int lane;
    int it;
    __device__ StoreStager(const EpiCtx& c) : base(smem_u32(c.staging) + c.ew * 8192), lane(c.lane), it(0) {}
    // Buffer for the next block; blocks until the store issued two blocks ago has read it out.
    __device__ __forceinline__ uint32_t acquire() {
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
        return base + (it & 1) * 4096;
    }
    __device__ __forceinline__ void put(uint32_t buf, int chunk, uint32_t a, uint32_t b, uint32_t c, uint32_t d) const {
        st_shared_v4(buf + lane * 128 + ((chunk ^ (lane & 7)) << 4), a, b, c, d);
    }
    template <bool REDUCE_ADD>
    __device__ __forceinline__ void commit(const CUtensorMap* tm, uint32_t buf, int c0, int c1) {
        fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA (async proxy)
        __syncwarp();
        if (lane == 0) {
            if constexpr (REDUCE_ADD) tma_reduce_add_2d(tm, buf, c0, c1);
            else tma_store_2d(tm, buf, c0, c1);
            bulk_commit();
        }
        ++it;
    }
    __device__ __forceinline__ void drain() {
        if (lane == 0) bulk_wait<0>();
        __syncwarp();
    }
};

template <bool MN_MAJOR, int TILE_MN>
__device__ __forceinline__ void load_operand(uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, int mn0, int k0) {
    if constexpr (MN_MAJOR) {
        // global [k rows][mn] (mn contiguous): one {64 mn x 64 k} box per 64-wide mn chunk;
        // the tile lands as [mn/64][BLOCK_K][64]
#pragma unroll
        for (int c = 0; c < TILE_MN / 64; ++c) tma_load_2d(dst + c * (BLOCK_K * 128), tm, bar, mn0 + c * 64, k0);
    } else {
        // global viewed as {k (contiguous), mn rows}; lands as [mn rows][64]
        tma_load_2d(dst, tm, bar, k0, mn0);
    }
}

template <bool MN_MAJOR>
__device__ __forceinline__ uint64_t operand_desc(uint32_t stage_base, int kk) {
    if constexpr (MN_MAJOR) {
        return make_smem_desc(stage_base + kk * (UMMA_K * 128), BLOCK_K * 128, 1024);
    } else {
        return make_smem_desc(stage_base + kk * (UMMA_K * 2), 16, 1024);
    }
}

template <class P>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const __grid_constant__ typename P::Params prm,
            const int extra_bytes) {
    constexpr int BLOCK_N = P::BLOCK_N;
    constexpr int STAGES = P::STAGES;
    constexpr int M_SUB = P::M_SUB;
    constexpr int ACC_BUFS = P::ACC_BUFS;
    constexpr int A_STAGE_BYTES = SmemLayout<P>::A_STAGE_BYTES;
    constexpr int B_STAGE_BYTES = SmemLayout<P>::B_STAGE_BYTES;
    constexpr int ACC_COLS = M_SUB * BLOCK_N;            // TMEM columns of one accumulator set
    constexpr uint32_t TMEM_COLS = ACC_BUFS * ACC_COLS;  // 256 or 512: powers of two >= 32
    static_assert(BLOCK_N == 128 || BLOCK_N == 256, "BLOCK_N must be 128 or 256");
    static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "accumulators must fill 256 or 512 TMEM columns");
    static_assert(M_SUB == 1 || P::A_MN, "stacked A sub-tiles are only wired for MN-major A");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = sA + STAGES * A_STAGE_BYTES;
    uint8_t* sStaging = sB + STAGES * B_STAGE_BYTES;  // 1024-aligned: stage sizes are multiples of 1 KB
    uint8_t* sExtra = sStaging + (P::STAGING ? STAGING_BYTES : 0);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sExtra + ((extra_bytes + 15) / 16) * 16);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + ACC_BUFS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + ACC_BUFS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (P::STAGING) tma_prefetch_desc(&tmC);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < ACC_BUFS; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 4);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    P::prologue(prm, sExtra, threadIdx.x);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // whole warp walks the schedule (uniform control flow); one elected lane issues the TMA loads
        typename P::Sched sched(prm, blockIdx.x, gridDim.x);
        Tile t;
        int stage = 0;
        uint32_t phase = 0;
        while (sched.next(t)) {
            for (int kb = 0; kb < t.kblocks; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&full_bar[stage], A_STAGE_BYTES + B_STAGE_BYTES);
                    load_operand<P::A_MN, BLOCK_M * M_SUB>(sA + stage * A_STAGE_BYTES, &tmA, &full_bar[stage], t.m0,
                                                           t.ka0 + kb * BLOCK_K);
                    load_operand<P::B_MN, BLOCK_N>(sB + stage * B_STAGE_BYTES, &tmB, &full_bar[stage], t.n0,
                                                   t.kb0 + kb * BLOCK_K);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // whole warp runs the loop, one elected lane issues: see elect_one() in ptx.cuh
        constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, P::A_MN, P::B_MN);
        typename P::Sched sched(prm, blockIdx.x, gridDim.x);
        Tile t;
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        const uint32_t sA_u32 = smem_u32(sA);
        const uint32_t sB_u32 = smem_u32(sB);
        while (sched.next(t)) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * ACC_COLS;
            for (int kb = 0; kb < t.kblocks; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t a_base = sA_u32 + stage * A_STAGE_BYTES;
                const uint32_t b_base = sB_u32 + stage * B_STAGE_BYTES;
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
                        const uint64_t bdesc = operand_desc<P::B_MN>(b_base, kk);
#pragma unroll
                        for (int ms = 0; ms < M_SUB; ++ms)
                            umma_bf16(tmem_d + ms * BLOCK_N, operand_desc<P::A_MN>(a_base + ms * A_SUB_BYTES, kk), bdesc,
                                      idesc, (kb | kk) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);  // frees this smem slot once the MMAs above retire
                    if (kb == t.kblocks - 1) umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == ACC_BUFS) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        EpiCtx ctx;
        ctx.extra = sExtra;
        ctx.staging = sStaging;
        ctx.tmC = &tmC;
        ctx.ew = warp - 4;  // == warp % 4: the TMEM lane quadrant this warp may read
        ctx.lane = lane;
        ctx.cta = blockIdx.x;
        typename P::Sched sched(prm, blockIdx.x, gridDim.x);
        typename P::Epi epi(prm, ctx);
        Tile t;
        int acc = 0;
        uint32_t acc_phase = 0;
        while (sched.next(t)) {
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * ACC_COLS + (static_cast<uint32_t>(ctx.ew * 32) << 16);
            epi.tile(t, taddr);
            // every tcgen05.ld of this tile has completed (tile() waits on its last load)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == ACC_BUFS) { acc = 0; acc_phase ^= 1; }
        }
        epi.finish();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// Host-side launcher (needs host_util.h included first for the error macros).
#ifdef AB_CHECK_CUDA
template <class P>
static int32_t launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                           const typename P::Params& prm, int grid, int extra_bytes, cudaStream_t st) {
    const size_t smem = gemm_smem_bytes<P>(extra_bytes);
    AB_REQUIRE(smem <= 227 * 1024, ARCFACE_B200_E_SHAPE, "shared memory request %zu exceeds 227 KB", smem);
    AB_REQUIRE(grid >= 1, ARCFACE_B200_E_SHAPE, "empty grid");
    static bool configured[64] = {false};  // the attribute is per function and per device
    int dev = 0;
    AB_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        AB_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured[dev] = true;
    }
    gemm_kernel<P><<<grid, GEMM_THREADS, smem, st>>>(tmA, tmB, tmC, prm, extra_bytes);
    AB_CHECK_CUDA(cudaGetLastError());
    return ARCFACE_B200_OK;
}
#endif

}  // namespace ab
