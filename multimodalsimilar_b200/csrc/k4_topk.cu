// K4 -- fused cosine top-k: for every query row the k largest cosines against C class / catalogue rows and their
// indices, without materialising the B x C cosine matrix.
//
// Reference: the eval loops take torch.argmax / top-k of ArcMarginProduct.forward_test's B x C output
// (arcface.py:65-67, nlp_classifier_train.py:143-156), and the product's retrieval step is a brute-force
// inner-product search over L2-normalised embeddings: faiss.normalize_L2 + IndexFlat(METRIC_INNER_PRODUCT)
// .search(x, k) with k = 13 / 26 / 100 (daodian_infer.py:225-230, 295-302).
//
// Exact, three launches on the shared tcgen05 GEMM core (gemm_core.cuh; any D):
//   1. SegMax  : cosine GEMM whose epilogue keeps only the maximum of every 128-column segment   [B][n_seg]
//      select  : tau[b] = k-th largest segment maximum of row b (radix select).  At most k - 1 segments hold an
//                element > tau, so at most 128 (k - 1) elements exceed it, and at least k elements are >= tau.
//   2. Emit    : the same GEMM again (bit-identical accumulators); the epilogue appends every element > tau, and
//                elements == tau up to k of them, to the row's candidate list (<= 128 k entries).
//   3. sort    : bitonic sort of the row's candidates in shared memory -> k (value, index) pairs, descending,
//                ties by ascending index.
// The same sort kernel merges per-rank lists of a class-sharded catalogue (arcface_b200_topk_merge).
#include "host_util.h"
#include "gemm_core.cuh"

#include <math.h>

namespace ab {

constexpr int TOPK_SEG = 128;   // columns per segment
constexpr int TOPK_MAX_K = 128;

struct TopkCommon {
    static constexpr int BLOCK_N = 256;
    static constexpr int STAGES = 4;
    static constexpr int M_SUB = 1;
    static constexpr int ACC_BUFS = 2;
    static constexpr bool STAGING = false;
    static constexpr bool A_MN = false;  // xhat [B][D]
    static constexpr bool B_MN = false;  // what [C][D]

    struct Params {
        int B, D, C;
        int m_tiles, n_tiles;
        int n_seg;          // ceil(C / 128)
        int ld;             // leading dimension of segmax (>= n_seg)
        float* segmax;      // pass 1 out: [B][ld] (row-major: the select kernel streams a row)
        const float* tau;   // pass 2 in:  [B]
        int k, cap;
        int* cnt;           // [B] candidates appended
        int* eq;            // [B] candidates equal to tau appended
        float* cand_val;    // [B][cap]
        int* cand_idx;      // [B][cap]
    };

    __device__ static void prologue(const Params&, uint8_t*, int) {}

    // every CTA walks tiles cta, cta + grid, ...; m fastest so that the CTAs sharing a What tile run together
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
};

struct TopkSegMax : TopkCommon {
    struct Epi {
        const Params& p;
        int ew, lane;
        __device__ Epi(const Params& prm, const EpiCtx& c) : p(prm), ew(c.ew), lane(c.lane) {}
        __device__ void tile(const Tile& t, uint32_t taddr) {
            const int row = t.m0 + ew * 32 + lane;
#pragma unroll 1
            for (int h = 0; h < BLOCK_N / TOPK_SEG; ++h) {
                float m = -INFINITY;
#pragma unroll 1
                for (int c = 0; c < TOPK_SEG / 32; ++c) {
                    const int col0 = t.n0 + h * TOPK_SEG + c * 32;
                    uint32_t v[32];
                    tmem_ld32(taddr + h * TOPK_SEG + c * 32, v);  // every warp issues every load (uniform)
                    tmem_ld_wait();
                    if (col0 >= p.C) continue;
                    if (col0 + 32 <= p.C) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < p.C) m = fmaxf(m, __uint_as_float(v[j]));
                    }
                }
                const int seg = (t.n0 + h * TOPK_SEG) / TOPK_SEG;
                if (seg < p.n_seg && row < p.B) p.segmax[static_cast<int64_t>(row) * p.ld + seg] = m;
            }
        }
        __device__ void finish() {}
    };
};

struct TopkEmit : TopkCommon {
    struct Epi {
        const Params& p;
        int ew, lane;
        __device__ Epi(const Params& prm, const EpiCtx& c) : p(prm), ew(c.ew), lane(c.lane) {}
        __device__ __forceinline__ void append(int row, float v, int col) const {
            const int pos = atomicAdd(p.cnt + row, 1);
            if (pos < p.cap) {
                p.cand_val[static_cast<int64_t>(row) * p.cap + pos] = v;
                p.cand_idx[static_cast<int64_t>(row) * p.cap + pos] = col;
            }
        }
        __device__ void tile(const Tile& t, uint32_t taddr) {
            const int row = t.m0 + ew * 32 + lane;
            const bool rv = row < p.B;
            const float tau = rv ? p.tau[row] : INFINITY;
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                const int col0 = t.n0 + c * 32;
                uint32_t v[32];
                tmem_ld32(taddr + c * 32, v);
                tmem_ld_wait();
                if (col0 >= p.C) continue;
                float m = -INFINITY;
#pragma unroll
                for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
                if (!(m >= tau)) continue;  // the common case: nothing of this chunk can be a candidate
                // fully unrolled so that v[] stays in registers (a dynamic index would push it to local memory)
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float x = __uint_as_float(v[j]);
                    const int col = col0 + j;
                    if (x >= tau && col < p.C) {
                        if (x > tau) append(row, x, col);
                        else if (atomicAdd(p.eq + row, 1) < p.k) append(row, x, col);
                    }
                }
            }
        }
        __device__ void finish() {}
    };
};

// order-preserving map float -> uint32 (larger float <=> larger key; -0 < +0, NaN not expected)
__device__ __forceinline__ uint32_t fkey(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// tau[b] = k-th largest of segmax[b, 0..n); -inf when n < k.  One CTA per row, 8-bit radix
// select from the most significant byte down.
__global__ void __launch_bounds__(256) topk_select_kernel(const float* __restrict__ segmax, int n, int ld, int k,
                                                          int cached, float* __restrict__ tau) {
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_k;
    extern __shared__ uint32_t row_keys[];  // the row's keys when they fit (cached = 1): read from global once
    const int b = blockIdx.x;
    if (n < k) {
        if (threadIdx.x == 0) tau[b] = -INFINITY;
        return;
    }
    const float* src = segmax + static_cast<int64_t>(b) * ld;
    if (cached)
        for (int i = threadIdx.x; i < n; i += blockDim.x) row_keys[i] = fkey(src[i]);
    if (threadIdx.x == 0) { s_prefix = 0u; s_k = static_cast<unsigned>(k); }
    unsigned mask = 0u;
    for (int shift = 24; shift >= 0; shift -= 8) {
        hist[threadIdx.x] = 0u;
        __syncthreads();
        const unsigned prefix = s_prefix;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const uint32_t key = cached ? row_keys[i] : fkey(src[i]);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned need = s_k, acc = 0u;
            int bin = 255;
            for (; bin > 0; --bin) {
                if (acc + hist[bin] >= need) break;
                acc += hist[bin];
            }
            s_k = need - acc;  // rank inside the chosen bin
            s_prefix = prefix | (static_cast<unsigned>(bin) << shift);
        }
        mask |= 255u << shift;
        __syncthreads();
    }
    if (threadIdx.x == 0) tau[b] = fkey_inv(s_prefix);
}

// Sort one row's candidates (value desc, index asc) and write the first k.  n = min(cnt[b], cap) candidates at
// val[b * cap ...]; indices from idx32 (+ class_offset) or idx64 (global already).  Rows with fewer than k
// candidates are padded with (-inf, -1).
__global__ void __launch_bounds__(512) topk_sort_kernel(const float* __restrict__ val, const int* __restrict__ idx32,
                                                        const int64_t* __restrict__ idx64, const int* __restrict__ cnt,
                                                        int fixed_n, int cap, int k, float scale, int64_t class_offset,
                                                        float* __restrict__ out_val, int64_t* __restrict__ out_idx) {
    extern __shared__ unsigned long long keys[];  // (fkey(value) << 32) | ~position  -> descending sort
    const int b = blockIdx.x;
    int n = cnt != nullptr ? cnt[b] : fixed_n;
    n = n < cap ? n : cap;
    int n_pad = 1;
    while (n_pad < n || n_pad < k) n_pad <<= 1;
    const float* v = val + static_cast<int64_t>(b) * cap;
    for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
        unsigned long long key = 0ull;  // sorts last
        if (i < n) {
            // tie-break by the candidate's class index: smaller index first
            const uint32_t id = idx32 != nullptr ? static_cast<uint32_t>(idx32[static_cast<int64_t>(b) * cap + i])
                                                 : static_cast<uint32_t>(idx64[static_cast<int64_t>(b) * cap + i]);
            key = (static_cast<unsigned long long>(fkey(v[i])) << 32) | static_cast<unsigned long long>(~id);
        }
        keys[i] = key;
    }
    __syncthreads();
    for (int size = 2; size <= n_pad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < n_pad / 2; i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const unsigned long long a = keys[lo], c = keys[hi];
                if (desc ? (a < c) : (a > c)) { keys[lo] = c; keys[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        const unsigned long long key = keys[j];
        if (j < n && key != 0ull) {
            const uint32_t id = ~static_cast<uint32_t>(key & 0xffffffffull);
            const float x = fkey_inv(static_cast<uint32_t>(key >> 32));
            out_val[static_cast<int64_t>(b) * k + j] = x == -INFINITY ? x : x * scale;
            out_idx[static_cast<int64_t>(b) * k + j] =
                x == -INFINITY ? -1 : (idx32 != nullptr ? static_cast<int64_t>(id) + class_offset : static_cast<int64_t>(id));
        } else {
            out_val[static_cast<int64_t>(b) * k + j] = -INFINITY;
            out_idx[static_cast<int64_t>(b) * k + j] = -1;
        }
    }
}

struct TopkPlan {
    int n_seg, ld, cap, m_tiles, n_tiles;
    size_t off_segmax, off_tau, off_cnt, off_eq, off_val, off_idx, total;
};

static TopkPlan topk_plan(int B, int64_t C, int k) {
    TopkPlan pl;
    pl.n_seg = static_cast<int>((C + TOPK_SEG - 1) / TOPK_SEG);
    pl.ld = (pl.n_seg + 31) / 32 * 32;
    const int64_t cap = static_cast<int64_t>(TOPK_SEG) * k;
    const int64_t all = static_cast<int64_t>(pl.n_seg) * TOPK_SEG;
    pl.cap = static_cast<int>(cap < all ? cap : all);
    pl.m_tiles = (B + BLOCK_M - 1) / BLOCK_M;
    pl.n_tiles = static_cast<int>((C + TopkCommon::BLOCK_N - 1) / TopkCommon::BLOCK_N);
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    size_t o = 0;
    pl.off_segmax = o; o = up(o + static_cast<size_t>(B) * pl.ld * 4);
    pl.off_tau = o;    o = up(o + static_cast<size_t>(B) * 4);
    pl.off_cnt = o;    o = up(o + static_cast<size_t>(B) * 4);
    pl.off_eq = o;     o = up(o + static_cast<size_t>(B) * 4);
    pl.off_val = o;    o = up(o + static_cast<size_t>(B) * pl.cap * 4);
    pl.off_idx = o;    o = up(o + static_cast<size_t>(B) * pl.cap * 4);
    pl.total = o;
    return pl;
}

static int32_t topk_check(const char* who, int32_t B, int32_t D, int64_t C, int32_t k) {
    AB_REQUIRE(B >= 1 && B <= ARCFACE_B200_MAX_BATCH, ARCFACE_B200_E_SHAPE, "%s: B=%d outside [1, %d]", who, B,
               ARCFACE_B200_MAX_BATCH);
    AB_REQUIRE(D >= 8 && D % 8 == 0, ARCFACE_B200_E_SHAPE, "%s: D=%d must be a positive multiple of 8", who, D);
    AB_REQUIRE(C >= 1 && C <= (1ll << 30), ARCFACE_B200_E_SHAPE, "%s: C=%lld outside [1, 2^30]", who, (long long)C);
    AB_REQUIRE(k >= 1 && k <= TOPK_MAX_K, ARCFACE_B200_E_SHAPE, "%s: k=%d outside [1, %d]", who, k, TOPK_MAX_K);
    return ARCFACE_B200_OK;
}

static int32_t launch_sort(const float* val, const int* idx32, const int64_t* idx64, const int* cnt, int fixed_n, int cap,
                           int B, int k, float scale, int64_t class_offset, float* out_val, int64_t* out_idx,
                           cudaStream_t st) {
    int n_pad = 1;
    while (n_pad < cap || n_pad < k) n_pad <<= 1;
    const size_t smem = static_cast<size_t>(n_pad) * 8;
    AB_REQUIRE(smem <= 200 * 1024, ARCFACE_B200_E_SHAPE, "top-k sort: %d candidates per row do not fit shared memory", cap);
    static bool configured[64] = {false};
    int dev = 0;
    AB_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        AB_CHECK_CUDA(cudaFuncSetAttribute(topk_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured[dev] = true;
    }
    topk_sort_kernel<<<B, 512, smem, st>>>(val, idx32, idx64, cnt, fixed_n, cap, k, scale, class_offset, out_val, out_idx);
    AB_CHECK_CUDA(cudaGetLastError());
    return ARCFACE_B200_OK;
}

}  // namespace ab

using namespace ab;

extern "C" int32_t arcface_b200_topk_workspace_bytes(int32_t B, int32_t D, int64_t C, int32_t k, size_t* bytes) {
    AB_REQUIRE(bytes, ARCFACE_B200_E_ARG, "topk_workspace_bytes: null pointer");
    if (int32_t rc = topk_check("topk_workspace_bytes", B, D, C, k)) return rc;
    *bytes = topk_plan(B, C, k).total;
    return ARCFACE_B200_OK;
}

extern "C" int32_t arcface_b200_cosine_topk(const uint16_t* xhat, const uint16_t* what, int32_t B, int32_t D, int64_t C,
                                            int32_t k, float scale, int64_t class_offset, float* out_val,
                                            int64_t* out_idx, void* workspace, size_t workspace_bytes, void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(xhat && what && out_val && out_idx && workspace, ARCFACE_B200_E_ARG, "cosine_topk: null pointer");
    if (int32_t rc = topk_check("cosine_topk", B, D, C, k)) return rc;
    const TopkPlan pl = topk_plan(B, C, k);
    AB_REQUIRE(workspace_bytes >= pl.total, ARCFACE_B200_E_WORKSPACE, "cosine_topk: workspace %zu < required %zu",
               workspace_bytes, pl.total);
    AB_REQUIRE(aligned16(workspace), ARCFACE_B200_E_LAYOUT, "cosine_topk: workspace must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    TopkCommon::Params p;
    p.B = B; p.D = D; p.C = static_cast<int>(C);
    p.m_tiles = pl.m_tiles; p.n_tiles = pl.n_tiles;
    p.n_seg = pl.n_seg; p.ld = pl.ld;
    p.segmax = reinterpret_cast<float*>(ws + pl.off_segmax);
    p.tau = reinterpret_cast<const float*>(ws + pl.off_tau);
    p.k = k; p.cap = pl.cap;
    p.cnt = reinterpret_cast<int*>(ws + pl.off_cnt);
    p.eq = reinterpret_cast<int*>(ws + pl.off_eq);
    p.cand_val = reinterpret_cast<float*>(ws + pl.off_val);
    p.cand_idx = reinterpret_cast<int*>(ws + pl.off_idx);
    CUtensorMap tmA, tmB;
    if (int32_t rc = make_tmap_kmajor(&tmA, xhat, D, B, D, BLOCK_M)) return rc;
    if (int32_t rc = make_tmap_kmajor(&tmB, what, D, C, D, TopkCommon::BLOCK_N)) return rc;
    const int64_t total = static_cast<int64_t>(pl.m_tiles) * pl.n_tiles;
    const int grid = static_cast<int>(total < sm_count() ? total : sm_count());
    if (int32_t rc = launch_gemm<TopkSegMax>(tmA, tmB, tmA, p, grid, 0, st)) return rc;
    const int cached = pl.n_seg <= 11 * 1024 ? 1 : 0;  // 44 KB of dynamic shared memory at most (default limit 48 KB)
    topk_select_kernel<<<B, 256, cached ? static_cast<size_t>(pl.n_seg) * 4 : 0, st>>>(
        p.segmax, pl.n_seg, pl.ld, k, cached, reinterpret_cast<float*>(ws + pl.off_tau));
    AB_CHECK_CUDA(cudaGetLastError());
    AB_CHECK_CUDA(cudaMemsetAsync(ws + pl.off_cnt, 0, pl.off_val - pl.off_cnt, st));  // cnt and eq
    if (int32_t rc = launch_gemm<TopkEmit>(tmA, tmB, tmA, p, grid, 0, st)) return rc;
    return launch_sort(p.cand_val, p.cand_idx, nullptr, p.cnt, 0, pl.cap, B, k, scale, class_offset, out_val, out_idx, st);
}

extern "C" int32_t arcface_b200_topk_merge(const float* val, const int64_t* idx, int32_t B, int32_t n_per_row, int32_t k,
                                           float* out_val, int64_t* out_idx, void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(val && idx && out_val && out_idx, ARCFACE_B200_E_ARG, "topk_merge: null pointer");
    AB_REQUIRE(B >= 1 && n_per_row >= 1 && n_per_row <= 16384 && k >= 1 && k <= TOPK_MAX_K, ARCFACE_B200_E_SHAPE,
               "topk_merge: bad shape (B=%d, n_per_row=%d, k=%d)", B, n_per_row, k);
    return launch_sort(val, nullptr, idx, nullptr, n_per_row, n_per_row, B, k, 1.0f, 0, out_val, out_idx,
                       static_cast<cudaStream_t>(stream));
}
