// K3 -- backward of the ArcFace head + softmax cross-entropy (loss.backward() through
// arcface.py:45-63 and CrossEntropyLoss), three tcgen05 GEMMs per class chunk on the shared core:
//
//   BwdDC  S^T = What . Xhat^T   (classes on accumulator rows) -> epilogue recomputes
//          p = exp(s cos - lse) from the saved row statistics, forms dC (label column uses the exact
//          fp32 margin derivative), accumulates q[c] = sum_b dC[b,c] cos[b,c] and writes dC^T (bf16)
//          into an L2-sized scratch chunk [classes][batch].
//   DW     dWhat = dC^T . Xhat   (K = batch) -> epilogue applies the normalise backward
//          dW[c] = (dWhat[c] - q[c] what[c]) * inv_nw[c] and streams fp32 dW.
//   DX     dXhat += dC . What    (K = classes, split across CTAs; both operands MN-major views of
//          the buffers already in memory) -> fp32 vector reductions into dXhat [B][D].
//
// The B x C probability matrix is never materialised: only a bounded chunk (<= ~64 MB, classes x batch
// bf16) lives in the workspace at a time.
#include "host_util.h"
#include "gemm_core.cuh"

#include <math.h>

namespace ab {

constexpr float LOG2E_B = 1.4426950408889634f;

__device__ __forceinline__ uint4 ldg_nc_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// ------------------------------------------------------------------ dC^T producer
struct BwdDC {
    static constexpr int BLOCK_N = 256;  // batch columns per tile
    static constexpr int STAGES = 4;
    static constexpr bool A_MN = false;  // what [C][D]
    static constexpr bool B_MN = false;  // xhat [B][D]

    struct Params {
        int B, D, C;
        int Bp;         // leading dimension of the scratch (multiple of 64)
        int c_begin;    // first class of this chunk (multiple of 128)
        int c_blocks;   // 128-class blocks in this chunk
        int n_tiles;    // ceil(B / 256)
        float s_log2e;  // s * log2(e)
        float coef;     // s * grad_scale
        const float* grad_dev;  // nullable device scalar multiplied into coef
        const float* lse;
        const float* one_minus_p;  // 1 - p_label, cancellation-free (finalize_rows)
        const float* dphi;
        const int* label_local;
        __nv_bfloat16* dct;  // [c_blocks * 128][Bp]
        float* q;            // [C]
    };

    static int extra_bytes(int n_tiles) { return n_tiles * BLOCK_N * 12; }

    // per-batch-column constants: lse * log2e (+inf on padding -> p = 0), label, label-column dC
    __device__ static void prologue(const Params& p, uint8_t* extra, int tid) {
        const int Bpad = p.n_tiles * BLOCK_N;
        float* lse2 = reinterpret_cast<float*>(extra);
        int* lab = reinterpret_cast<int*>(extra + Bpad * 4);
        float* dlab = reinterpret_cast<float*>(extra + Bpad * 8);
        const float coef = p.coef * (p.grad_dev != nullptr ? *p.grad_dev : 1.f);
        for (int b = tid; b < Bpad; b += GEMM_THREADS) {
            if (b < p.B) {
                const float l = p.lse[b];
                const int y = p.label_local[b];
                lse2[b] = l * LOG2E_B;
                lab[b] = y;
                dlab[b] = (y >= 0) ? -coef * p.one_minus_p[b] * p.dphi[b] : 0.f;
            } else {
                lse2[b] = INFINITY;
                lab[b] = -1;
                dlab[b] = 0.f;
            }
        }
    }

    struct Sched {
        int cb, nt, step, c_blocks, n_tiles, c_begin, kblocks;
        __device__ Sched(const Params& p, int cta, int ncta) {
            cb = cta;
            nt = 0;
            step = ncta;
            c_blocks = p.c_blocks;
            n_tiles = p.n_tiles;
            c_begin = p.c_begin;
            kblocks = (p.D + BLOCK_K - 1) / BLOCK_K;
        }
        __device__ bool next(Tile& t) {
            if (cb >= c_blocks) return false;
            t.m0 = c_begin + cb * BLOCK_M;
            t.n0 = nt * BLOCK_N;
            t.ka0 = 0;
            t.kb0 = 0;
            t.kblocks = kblocks;
            t.aux = (nt == 0 ? 1 : 0) | (nt == n_tiles - 1 ? 2 : 0);
            if (++nt == n_tiles) { nt = 0; cb += step; }
            return true;
        }
    };

    struct Epi {
        const Params& p;
        const float* lse2;
        const int* lab;
        const float* dlab;
        int ew, lane;
        float qacc, coef_all;
        __device__ Epi(const Params& prm, uint8_t* extra, int ew_, int lane_, int) : p(prm), ew(ew_), lane(lane_) {
            coef_all = p.coef * (p.grad_dev != nullptr ? *p.grad_dev : 1.f);
            const int Bpad = p.n_tiles * BLOCK_N;
            lse2 = reinterpret_cast<const float*>(extra);
            lab = reinterpret_cast<const int*>(extra + Bpad * 4);
            dlab = reinterpret_cast<const float*>(extra + Bpad * 8);
            qacc = 0.f;
        }
        __device__ void tile(const Tile& t, uint32_t taddr) {
            const int c = t.m0 + ew * 32 + lane;  // class owned by this thread
            const bool cvalid = c < p.C;
            const float coef = cvalid ? coef_all : 0.f;
            const int cmatch = cvalid ? c : -2;
            if (t.aux & 1) qacc = 0.f;
            __nv_bfloat16* orow = p.dct + static_cast<int64_t>(c - p.c_begin) * p.Bp;
#pragma unroll 1
            for (int cc = 0; cc < BLOCK_N / 32; ++cc) {
                uint32_t v[32];
                tmem_ld32(taddr + cc * 32, v);
                tmem_ld_wait();
                const int b0 = t.n0 + cc * 32;
                if (b0 >= p.Bp) break;  // warp-uniform: nothing to store past the padded batch
                uint32_t packed[16];
#pragma unroll
                for (int j4 = 0; j4 < 32; j4 += 4) {
                    const float4 l4 = *reinterpret_cast<const float4*>(lse2 + b0 + j4);
                    const int4 y4 = *reinterpret_cast<const int4*>(lab + b0 + j4);
                    const float ls[4] = {l4.x, l4.y, l4.z, l4.w};
                    const int ys[4] = {y4.x, y4.y, y4.z, y4.w};
                    float dc[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float cosv = __uint_as_float(v[j4 + j]);
                        float d = coef * ex2(fmaf(cosv, p.s_log2e, -ls[j]));
                        if (ys[j] == cmatch) d = dlab[b0 + j4 + j];  // rare: this class is row b's label
                        qacc = fmaf(d, cosv, qacc);
                        dc[j] = d;
                    }
                    packed[j4 / 2] = pack_bf16x2(dc[0], dc[1]);
                    packed[j4 / 2 + 1] = pack_bf16x2(dc[2], dc[3]);
                }
                uint4* o = reinterpret_cast<uint4*>(orow + b0);
#pragma unroll
                for (int k = 0; k < 4; ++k) o[k] = make_uint4(packed[4 * k], packed[4 * k + 1], packed[4 * k + 2], packed[4 * k + 3]);
            }
            if ((t.aux & 2) && cvalid) p.q[c] = qacc;
        }
        __device__ void finish() {}
    };
};

// ------------------------------------------------------------------ dW
struct BwdDW {
    static constexpr int BLOCK_N = 256;  // embedding columns per tile
    static constexpr int STAGES = 4;
    static constexpr bool A_MN = false;  // dC^T chunk [classes][Bp], K = batch contiguous
    static constexpr bool B_MN = false;  // xhat^T [D][ld_t], K = batch contiguous

    struct Params {
        int B, D, C;
        int c_begin, c_blocks;
        int dn_tiles;
        const float* q;
        const float* inv_nw;
        const __nv_bfloat16* what;
        float* dw;
    };

    __device__ static void prologue(const Params&, uint8_t*, int) {}

    struct Sched {
        int idx, total, step, dn_tiles, kblocks;
        __device__ Sched(const Params& p, int cta, int ncta) {
            idx = cta;
            step = ncta;
            dn_tiles = p.dn_tiles;
            total = p.c_blocks * p.dn_tiles;
            kblocks = (p.B + BLOCK_K - 1) / BLOCK_K;
        }
        __device__ bool next(Tile& t) {
            if (idx >= total) return false;
            t.m0 = (idx / dn_tiles) * BLOCK_M;  // chunk-relative class row
            t.n0 = (idx % dn_tiles) * BLOCK_N;
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
        __device__ Epi(const Params& prm, uint8_t*, int ew_, int lane_, int) : p(prm), ew(ew_), lane(lane_) {}
        __device__ void tile(const Tile& t, uint32_t taddr) {
            const int c = p.c_begin + t.m0 + ew * 32 + lane;
            const bool cvalid = c < p.C;
            const float qc = cvalid ? p.q[c] : 0.f;
            const float inw = cvalid ? p.inv_nw[c] : 0.f;
            const int64_t roff = static_cast<int64_t>(cvalid ? c : 0) * p.D;
#pragma unroll 1
            for (int cc = 0; cc < BLOCK_N / 32; ++cc) {
                uint32_t v[32];
                tmem_ld32(taddr + cc * 32, v);
                tmem_ld_wait();
                const int d0 = t.n0 + cc * 32;
                if (d0 >= p.D) break;
                if (cvalid) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int d = d0 + g * 8;
                        if (d < p.D) {  // D % 8 == 0: groups of 8 are all-in or all-out
                            const uint4 w = ldg_nc_u4(p.what + roff + d);
                            float4 o0, o1;
                            o0.x = (__uint_as_float(v[g * 8 + 0]) - qc * bf16_lo(w.x)) * inw;
                            o0.y = (__uint_as_float(v[g * 8 + 1]) - qc * bf16_hi(w.x)) * inw;
                            o0.z = (__uint_as_float(v[g * 8 + 2]) - qc * bf16_lo(w.y)) * inw;
                            o0.w = (__uint_as_float(v[g * 8 + 3]) - qc * bf16_hi(w.y)) * inw;
                            o1.x = (__uint_as_float(v[g * 8 + 4]) - qc * bf16_lo(w.z)) * inw;
                            o1.y = (__uint_as_float(v[g * 8 + 5]) - qc * bf16_hi(w.z)) * inw;
                            o1.z = (__uint_as_float(v[g * 8 + 6]) - qc * bf16_lo(w.w)) * inw;
                            o1.w = (__uint_as_float(v[g * 8 + 7]) - qc * bf16_hi(w.w)) * inw;
                            float4* o = reinterpret_cast<float4*>(p.dw + roff + d);
                            o[0] = o0;
                            o[1] = o1;
                        }
                    }
                }
            }
        }
        __device__ void finish() {}
    };
};

// ------------------------------------------------------------------ dXhat (split over classes)
struct BwdDX {
    static constexpr int BLOCK_N = 256;  // embedding columns per tile
    static constexpr int STAGES = 4;
    static constexpr bool A_MN = true;  // dC^T chunk [classes = K][batch = M contiguous]
    static constexpr bool B_MN = true;  // what [classes = K][D = N contiguous]

    struct Params {
        int B, D;
        int c_begin;
        int m_tiles, dn_tiles, splits;
        int kb_total;      // 64-class slices in this chunk
        int kb_per_split;  // ceil(kb_total / splits); no split is empty
        float* dxhat;
    };

    __device__ static void prologue(const Params&, uint8_t*, int) {}

    struct Sched {
        const Params& p;
        int idx, total, step;
        __device__ Sched(const Params& prm, int cta, int ncta) : p(prm) {
            idx = cta;
            step = ncta;
            total = p.m_tiles * p.dn_tiles * p.splits;
        }
        __device__ bool next(Tile& t) {
            if (idx >= total) return false;
            const int mt = idx % p.m_tiles;
            const int r = idx / p.m_tiles;
            const int dn = r % p.dn_tiles;
            const int sp = r / p.dn_tiles;
            const int kb0 = sp * p.kb_per_split;
            t.m0 = mt * BLOCK_M;
            t.n0 = dn * BLOCK_N;
            t.ka0 = kb0 * BLOCK_K;              // chunk-relative class row in the scratch
            t.kb0 = p.c_begin + kb0 * BLOCK_K;  // absolute class row in what
            t.kblocks = min(p.kb_per_split, p.kb_total - kb0);
            t.aux = 0;
            idx += step;
            return true;
        }
    };

    struct Epi {
        const Params& p;
        int ew, lane;
        __device__ Epi(const Params& prm, uint8_t*, int ew_, int lane_, int) : p(prm), ew(ew_), lane(lane_) {}
        __device__ void tile(const Tile& t, uint32_t taddr) {
            const int row = t.m0 + ew * 32 + lane;
            const bool rv = row < p.B;
            float* orow = p.dxhat + static_cast<int64_t>(rv ? row : 0) * p.D;
#pragma unroll 1
            for (int cc = 0; cc < BLOCK_N / 32; ++cc) {
                uint32_t v[32];
                tmem_ld32(taddr + cc * 32, v);
                tmem_ld_wait();
                const int d0 = t.n0 + cc * 32;
                if (d0 >= p.D) break;
                if (rv) {
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const int d = d0 + g * 4;
                        if (d < p.D)
                            red_add_v4(orow + d, __uint_as_float(v[g * 4 + 0]), __uint_as_float(v[g * 4 + 1]),
                                       __uint_as_float(v[g * 4 + 2]), __uint_as_float(v[g * 4 + 3]));
                    }
                }
            }
        }
        __device__ void finish() {}
    };
};

struct BwdPlan {
    int Bp;             // scratch leading dimension
    int chunk_classes;  // classes per chunk (multiple of 128)
    size_t scratch_off, scratch_bytes, q_off, q_bytes, total;
};

static BwdPlan plan_backward(int B, int64_t C, int nsm) {
    BwdPlan pl;
    pl.Bp = ((B + 63) / 64) * 64;
    if (nsm < 1) nsm = 148;
    // one wave of 128-class blocks per SM, repeated while the chunk stays <= 64 MB (L2-friendly)
    const size_t wave_bytes = static_cast<size_t>(128) * nsm * pl.Bp * 2;
    size_t k = (64u << 20) / wave_bytes;
    if (k < 1) k = 1;
    int64_t chunk = static_cast<int64_t>(128) * nsm * static_cast<int64_t>(k);
    const int64_t c_round = ((C + 127) / 128) * 128;
    if (chunk > c_round) chunk = c_round;
    pl.chunk_classes = static_cast<int>(chunk);
    pl.scratch_off = 0;
    pl.scratch_bytes = static_cast<size_t>(chunk) * pl.Bp * 2;
    pl.q_off = (pl.scratch_bytes + 255) / 256 * 256;
    pl.q_bytes = static_cast<size_t>(C) * 4;
    pl.total = pl.q_off + (pl.q_bytes + 255) / 256 * 256;
    return pl;
}

}  // namespace ab

using namespace ab;

extern "C" int32_t arcface_b200_backward_workspace_bytes(int32_t B, int32_t D, int64_t C_local, size_t* bytes) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(bytes, ARCFACE_B200_E_ARG, "backward_workspace_bytes: null pointer");
    AB_REQUIRE(B >= 1 && B <= ARCFACE_B200_MAX_BATCH && D >= 8 && D % 8 == 0 && C_local >= 1 && C_local <= (1ll << 30),
               ARCFACE_B200_E_SHAPE, "backward_workspace_bytes: bad shape B=%d D=%d C=%lld", B, D, (long long)C_local);
    *bytes = plan_backward(B, C_local, sm_count()).total;
    return ARCFACE_B200_OK;
}

extern "C" int32_t arcface_b200_backward_plan(int32_t B, int32_t D, int64_t C_local, int64_t* chunk_classes,
                                              int32_t* n_chunks) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(chunk_classes && n_chunks, ARCFACE_B200_E_ARG, "backward_plan: null pointer");
    AB_REQUIRE(B >= 1 && B <= ARCFACE_B200_MAX_BATCH && D >= 8 && D % 8 == 0 && C_local >= 1 && C_local <= (1ll << 30),
               ARCFACE_B200_E_SHAPE, "backward_plan: bad shape B=%d D=%d C=%lld", B, D, (long long)C_local);
    const BwdPlan pl = plan_backward(B, C_local, sm_count());
    *chunk_classes = pl.chunk_classes;
    *n_chunks = static_cast<int32_t>((C_local + pl.chunk_classes - 1) / pl.chunk_classes);
    return ARCFACE_B200_OK;
}

extern "C" int32_t arcface_b200_backward(const uint16_t* xhat, const uint16_t* xhat_t, int64_t ld_t,
                                         const uint16_t* what, const float* inv_nw, const float* lse,
                                         const float* one_minus_p, const float* dphi, const int32_t* label_local,
                                         int32_t B, int32_t D, int64_t C_local, float s, float grad_scale,
                                         const float* grad_loss_dev, float* dxhat, float* dw, void* workspace,
                                         size_t workspace_bytes, void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(xhat && xhat_t && what && inv_nw && lse && one_minus_p && dphi && label_local && dxhat && dw && workspace,
               ARCFACE_B200_E_ARG, "backward: null pointer");
    AB_REQUIRE(B >= 1 && B <= ARCFACE_B200_MAX_BATCH, ARCFACE_B200_E_SHAPE, "backward: B=%d outside [1, %d]", B,
               ARCFACE_B200_MAX_BATCH);
    AB_REQUIRE(D >= 8 && D % 8 == 0, ARCFACE_B200_E_SHAPE, "backward: D=%d must be a positive multiple of 8", D);
    AB_REQUIRE(C_local >= 1 && C_local <= (1ll << 30), ARCFACE_B200_E_SHAPE, "backward: bad C_local");
    AB_REQUIRE(ld_t >= B && ld_t % 8 == 0, ARCFACE_B200_E_LAYOUT, "backward: ld_t must be >= B and a multiple of 8");
    AB_REQUIRE(aligned16(dxhat) && aligned16(dw) && aligned16(workspace) && aligned16(what), ARCFACE_B200_E_LAYOUT,
               "backward: pointers must be 16-byte aligned");
    AB_REQUIRE(s > 0.f, ARCFACE_B200_E_ARG, "backward: scale s must be positive");
    const int nsm = sm_count();
    const BwdPlan pl = plan_backward(B, C_local, nsm);
    AB_REQUIRE(workspace_bytes >= pl.total, ARCFACE_B200_E_WORKSPACE, "backward: workspace %zu < required %zu",
               workspace_bytes, pl.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    __nv_bfloat16* dct = reinterpret_cast<__nv_bfloat16*>(ws + pl.scratch_off);
    float* q = reinterpret_cast<float*>(ws + pl.q_off);
    const int C = static_cast<int>(C_local);

    AB_CHECK_CUDA(cudaMemsetAsync(dxhat, 0, static_cast<size_t>(B) * D * sizeof(float), st));

    CUtensorMap tm_w_k, tm_x_k, tm_dct_k, tm_xt_k, tm_dct_mn, tm_w_mn;
    if (int32_t rc = make_tmap_kmajor(&tm_w_k, what, D, C_local, D, BLOCK_M)) return rc;
    if (int32_t rc = make_tmap_kmajor(&tm_x_k, xhat, D, B, D, BwdDC::BLOCK_N)) return rc;
    if (int32_t rc = make_tmap_kmajor(&tm_dct_k, dct, B, pl.chunk_classes, pl.Bp, BLOCK_M)) return rc;
    if (int32_t rc = make_tmap_kmajor(&tm_xt_k, xhat_t, B, D, ld_t, BwdDW::BLOCK_N)) return rc;
    if (int32_t rc = make_tmap_mnmajor(&tm_dct_mn, dct, B, pl.chunk_classes, pl.Bp)) return rc;
    if (int32_t rc = make_tmap_mnmajor(&tm_w_mn, what, D, C_local, D)) return rc;

    const int n_tiles = (B + BwdDC::BLOCK_N - 1) / BwdDC::BLOCK_N;
    const int m_tiles = (B + BLOCK_M - 1) / BLOCK_M;
    const int dn_tiles = (D + 255) / 256;

    for (int64_t c0 = 0; c0 < C_local; c0 += pl.chunk_classes) {
        const int cn = static_cast<int>(C_local - c0 < pl.chunk_classes ? C_local - c0 : pl.chunk_classes);
        const int c_blocks = (cn + BLOCK_M - 1) / BLOCK_M;
        {
            BwdDC::Params p;
            p.B = B; p.D = D; p.C = C; p.Bp = pl.Bp;
            p.c_begin = static_cast<int>(c0); p.c_blocks = c_blocks; p.n_tiles = n_tiles;
            p.s_log2e = s * LOG2E_B; p.coef = s * grad_scale; p.grad_dev = grad_loss_dev;
            p.lse = lse; p.one_minus_p = one_minus_p; p.dphi = dphi; p.label_local = label_local;
            p.dct = dct; p.q = q;
            const int grid = c_blocks < nsm ? c_blocks : nsm;
            if (int32_t rc = launch_gemm<BwdDC>(tm_w_k, tm_x_k, p, grid, BwdDC::extra_bytes(n_tiles), st)) return rc;
        }
        {
            BwdDW::Params p;
            p.B = B; p.D = D; p.C = C;
            p.c_begin = static_cast<int>(c0); p.c_blocks = c_blocks; p.dn_tiles = dn_tiles;
            p.q = q; p.inv_nw = inv_nw; p.what = reinterpret_cast<const __nv_bfloat16*>(what); p.dw = dw;
            const int total = c_blocks * dn_tiles;
            const int grid = total < nsm ? total : nsm;
            if (int32_t rc = launch_gemm<BwdDW>(tm_dct_k, tm_xt_k, p, grid, 0, st)) return rc;
        }
        {
            BwdDX::Params p;
            p.B = B; p.D = D; p.c_begin = static_cast<int>(c0);
            p.m_tiles = m_tiles; p.dn_tiles = dn_tiles;
            p.kb_total = (c_blocks * BLOCK_M) / BLOCK_K;
            int splits = nsm / (m_tiles * dn_tiles);
            if (splits < 1) splits = 1;
            if (splits > p.kb_total) splits = p.kb_total;
            p.kb_per_split = (p.kb_total + splits - 1) / splits;
            p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
            p.dxhat = dxhat;
            const int total = m_tiles * dn_tiles * p.splits;
            const int grid = total < nsm ? total : nsm;
            if (int32_t rc = launch_gemm<BwdDX>(tm_dct_mn, tm_w_mn, p, grid, 0, st)) return rc;
        }
    }
    return ARCFACE_B200_OK;
}
