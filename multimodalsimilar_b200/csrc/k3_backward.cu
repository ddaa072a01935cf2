// K3 -- backward of the ArcFace head + softmax cross-entropy (loss.backward() through
// arcface.py:45-63 and CrossEntropyLoss), three tcgen05 GEMMs per class chunk on the shared core:
//
//   BwdDC  S^T = What . Xhat^T   (classes on accumulator rows) -> epilogue recomputes
//          p = exp(s cos - lse) from the saved row statistics, forms dC (label column: the exact
//          fp32 margin derivative times the cancellation-free 1 - p_label), accumulates
//          q[c] = sum_b dC[b,c] cos[b,c] and writes dC^T (bf16) into an L2-sized scratch chunk
//          [classes][batch].
//   DW     dWhat = dC^T . Xhat   (K = batch) -> epilogue applies the normalise backward
//          dW[c] = (dWhat[c] - q[c] what[c]) * inv_nw[c] and streams fp32 dW.
//   DX     dXhat += dC . What    (K = classes, split across CTAs; both operands MN-major views of
//          the buffers already in memory) -> fp32 TMA reduce-adds into dXhat [B][D].
//
// All three epilogues leave through swizzled shared memory + TMA (full 128-byte lines).
// The B x C probability matrix is never materialised: only a bounded chunk (<= ~64 MB, classes x batch
// bf16) lives in the workspace at a time.
#include "host_util.h"
#include "gemm_core.cuh"
#include "gemm_rs.cuh"
#include "gemm_pair.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef AB_DX_SUM_MODE
#define AB_DX_SUM_MODE 1
#endif
#ifndef AB_DXP_COST
#define AB_DXP_COST 2300.0   // pair-cycles per 256-class block and 256 x 256 of dX output on the CTA-pair role
#endif

namespace ab {

constexpr float LOG2E_B = 1.4426950408889634f;

__device__ __forceinline__ uint4 ldg_nc_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void ldg_nc_u8(const void* p, uint4& a, uint4& b) {  // 32-byte aligned
    asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p));
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// ------------------------------------------------------------------ dC^T producer
struct BwdDC {
    static constexpr int BLOCK_N = 256;  // batch columns per tile
    static constexpr int STAGES = 3;
    static constexpr int M_SUB = 1;
    static constexpr int ACC_BUFS = 2;
    static constexpr bool STAGING = true;
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
        const float* grad_dev;     // nullable device scalar multiplied into coef
        const float* lse;
        const float* one_minus_p;  // 1 - p_label, cancellation-free (finalize_rows)
        const float* dphi;
        const int* label_local;
        float* q;  // [C]
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
                const int y = p.label_local[b];
                lse2[b] = p.lse[b] * LOG2E_B;
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
        const CUtensorMap* tm_out;  // dC^T scratch [chunk classes][Bp] bf16
        StoreStager stager;
        const float* lse2;
        const int* lab;
        const float* dlab;
        int ew, lane;
        float qacc, coef_all;
        __device__ Epi(const Params& prm, const EpiCtx& c) : p(prm), tm_out(c.tmC), stager(c), ew(c.ew), lane(c.lane) {
            coef_all = p.coef * (p.grad_dev != nullptr ? *p.grad_dev : 1.f);
            const int Bpad = p.n_tiles * BLOCK_N;
            lse2 = reinterpret_cast<const float*>(c.extra);
            lab = reinterpret_cast<const int*>(c.extra + Bpad * 4);
            dlab = reinterpret_cast<const float*>(c.extra + Bpad * 8);
            qacc = 0.f;
        }
        // 8 consecutive batch columns -> one 16-byte chunk of bf16
        __device__ __forceinline__ void eight(const uint32_t* v, int b, float coef, int cmatch, uint32_t (&o)[4]) {
            const float4 l0 = *reinterpret_cast<const float4*>(lse2 + b);
            const float4 l1 = *reinterpret_cast<const float4*>(lse2 + b + 4);
            const int4 y0 = *reinterpret_cast<const int4*>(lab + b);
            const int4 y1 = *reinterpret_cast<const int4*>(lab + b + 4);
            const float ls[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
            const int ys[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
            float dc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float cosv = __uint_as_float(v[j]);
                float d = coef * ex2(fmaf(cosv, p.s_log2e, -ls[j]));
                if (ys[j] == cmatch) d = dlab[b + j];  // rare: this class is row b's label
                qacc = fmaf(d, cosv, qacc);
                dc[j] = d;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = pack_bf16x2(dc[2 * j], dc[2 * j + 1]);
        }
        __device__ void tile(const Tile& t, uint32_t taddr) {
            const int c = t.m0 + ew * 32 + lane;  // class owned by this thread
            const bool cvalid = c < p.C;
            const float coef = cvalid ? coef_all : 0.f;
            const int cmatch = cvalid ? c : -2;
            if (t.aux & 1) qacc = 0.f;
            const int row0 = t.m0 - p.c_begin + ew * 32;  // chunk-relative scratch row of this warp's block
#pragma unroll 1
            for (int g = 0; g < BLOCK_N / 64; ++g) {
                const int b0 = t.n0 + g * 64;
                if (b0 >= p.Bp) break;  // warp-uniform: nothing to store past the padded batch
                uint32_t v0[32], v1[32];
                tmem_ld32(taddr + g * 64, v0);
                tmem_ld32(taddr + g * 64 + 32, v1);
                tmem_ld_wait();
                const uint32_t buf = stager.acquire();
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t o[4];
                    eight(v0 + 8 * k, b0 + 8 * k, coef, cmatch, o);
                    stager.put(buf, k, o[0], o[1], o[2], o[3]);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t o[4];
                    eight(v1 + 8 * k, b0 + 32 + 8 * k, coef, cmatch, o);
                    stager.put(buf, 4 + k, o[0], o[1], o[2], o[3]);
                }
                stager.commit<false>(tm_out, buf, b0, row0);
            }
            if ((t.aux & 2) && cvalid) p.q[c] = qacc;
        }
        __device__ void finish() { stager.drain(); }
    };
};

// ------------------------------------------------------------------ dW
struct BwdDW {
    static constexpr int BLOCK_N = 256;  // embedding columns per tile
    static constexpr int STAGES = 4;
    static constexpr int M_SUB = 1;
    static constexpr int ACC_BUFS = 2;
    static constexpr bool STAGING = true;
    static constexpr bool A_MN = false;  // dC^T chunk [classes][Bp], K = batch contiguous
    static constexpr bool B_MN = false;  // xhat^T [D][ld_t], K = batch contiguous

    struct Params {
        int B, D, C;
        int c_begin, c_blocks;
        int dn_tiles;
        const float* q;  // [q_slots][C] partial sums, added up here
        int q_slots;
        const float* inv_nw;
        const __nv_bfloat16* what;
        int64_t ldw;  // row stride of what (D; 3 D in the bf16x3 mode, whose rows start with the hi part)
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
        const CUtensorMap* tm_out;  // dW [C][D] fp32
        StoreStager stager;
        int ew, lane;
        __device__ Epi(const Params& prm, const EpiCtx& c) : p(prm), tm_out(c.tmC), stager(c), ew(c.ew), lane(c.lane) {}
        // the 32 normalised weights what[c, d0 .. d0+32) as 4 x 16 bytes (zeros past D or for padding rows)
        __device__ __forceinline__ void load_w(const __nv_bfloat16* wrow, bool cvalid, int d0, uint4 (&w)[4]) const {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int d = d0 + g * 8;
                w[g] = (cvalid && d < p.D) ? ldg_nc_u4(wrow + d) : make_uint4(0, 0, 0, 0);
            }
        }
        __device__ void tile(const Tile& t, uint32_t taddr) {
            const int crow0 = p.c_begin + t.m0 + ew * 32;
            const int c = crow0 + lane;
            const bool cvalid = c < p.C;
            float qc = 0.f;
            if (cvalid)
                for (int sl = 0; sl < p.q_slots; ++sl) qc += p.q[static_cast<int64_t>(sl) * p.C + c];
            const float inw = cvalid ? p.inv_nw[c] : 0.f;
            const __nv_bfloat16* wrow = p.what + static_cast<int64_t>(cvalid ? c : 0) * p.ldw;
            uint4 wcur[4], wnext[4];
            load_w(wrow, cvalid, t.n0, wcur);
#pragma unroll 1
            for (int cc = 0; cc < BLOCK_N / 32; ++cc) {
                const int d0 = t.n0 + cc * 32;
                if (d0 >= p.D) break;
                uint32_t v[32];
                tmem_ld32(taddr + cc * 32, v);
                if (cc + 1 < BLOCK_N / 32) load_w(wrow, cvalid, d0 + 32, wnext);  // prefetch the next chunk's weights
                tmem_ld_wait();
                const uint32_t buf = stager.acquire();
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint4 w = wcur[g];
                    const float o0 = (__uint_as_float(v[g * 8 + 0]) - qc * bf16_lo(w.x)) * inw;
                    const float o1 = (__uint_as_float(v[g * 8 + 1]) - qc * bf16_hi(w.x)) * inw;
                    const float o2 = (__uint_as_float(v[g * 8 + 2]) - qc * bf16_lo(w.y)) * inw;
                    const float o3 = (__uint_as_float(v[g * 8 + 3]) - qc * bf16_hi(w.y)) * inw;
                    const float o4 = (__uint_as_float(v[g * 8 + 4]) - qc * bf16_lo(w.z)) * inw;
                    const float o5 = (__uint_as_float(v[g * 8 + 5]) - qc * bf16_hi(w.z)) * inw;
                    const float o6 = (__uint_as_float(v[g * 8 + 6]) - qc * bf16_lo(w.w)) * inw;
                    const float o7 = (__uint_as_float(v[g * 8 + 7]) - qc * bf16_hi(w.w)) * inw;
                    stager.put(buf, 2 * g, __float_as_uint(o0), __float_as_uint(o1), __float_as_uint(o2), __float_as_uint(o3));
                    stager.put(buf, 2 * g + 1, __float_as_uint(o4), __float_as_uint(o5), __float_as_uint(o6), __float_as_uint(o7));
                }
                stager.commit<false>(tm_out, buf, d0, crow0);  // rows >= C / columns >= D are clipped by the TMA
#pragma unroll
                for (int g = 0; g < 4; ++g) wcur[g] = wnext[g];
            }
        }
        __device__ void finish() { stager.drain(); }
    };
};

// ------------------------------------------------------------------ dXhat (split over classes)
struct BwdDX {
    static constexpr int BLOCK_N = 256;  // embedding columns per tile
    static constexpr int STAGES = 4;
    static constexpr int M_SUB = 1;
    static constexpr int ACC_BUFS = 2;
    static constexpr bool STAGING = true;
    static constexpr bool A_MN = true;  // dC^T chunk [classes = K][batch = M contiguous]
    static constexpr bool B_MN = true;  // what [classes = K][D = N contiguous]

    struct Params {
        int B, D;
        int c_begin;
        int m_tiles, dn_tiles, splits;
        int kb_total;      // 64-class slices in this chunk
        int kb_per_split;  // ceil(kb_total / splits); no split is empty
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
        const CUtensorMap* tm_out;  // dXhat [B][D] fp32, accumulated with TMA reduce-add
        StoreStager stager;
        int ew, lane;
        __device__ Epi(const Params& prm, const EpiCtx& c) : p(prm), tm_out(c.tmC), stager(c), ew(c.ew), lane(c.lane) {}
        __device__ void tile(const Tile& t, uint32_t taddr) {
            const int row0 = t.m0 + ew * 32;
#pragma unroll 1
            for (int cc = 0; cc < BLOCK_N / 32; ++cc) {
                const int d0 = t.n0 + cc * 32;
                if (d0 >= p.D) break;
                uint32_t v[32];
                tmem_ld32(taddr + cc * 32, v);
                tmem_ld_wait();
                if (row0 >= p.B) continue;  // warp-uniform: this warp's 32 rows are all padding
                const uint32_t buf = stager.acquire();
#pragma unroll
                for (int g = 0; g < 8; ++g) stager.put(buf, g, v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                stager.commit<true>(tm_out, buf, d0, row0);  // rows >= B / columns >= D are clipped by the TMA
            }
        }
        __device__ void finish() { stager.drain(); }
    };
};

// ================================================================== resident-operand fast path
// (gemm_rs.cuh; used when the reused operand fits in shared memory: D <= 512 for dC^T, batch <= 512 for dW)

// ------------------------------------------------------------------ dC^T producer, Xhat slice resident
struct BwdDCr {
    struct Params {
        rs::Core core;  // streamed = what rows (class blocks), resident = 128 batch rows of xhat, K = D
        int B, C, Bp;
        float s_log2e;
        float coef;
        const float* grad_dev;
        const float* lse;
        const float* one_minus_p;
        const float* dphi;
        const int* label_local;
        float* q;  // [2 * n_res][C]: one slot per (batch slice, column half)
    };
    static constexpr int EXTRA_BYTES = rs::BN * 12;

    // constants of this CTA's 128 batch columns: lse * log2e (+inf on padding -> p = 0), label, label-column dC
    __device__ static void prologue(const Params& p, uint8_t* extra, int tid, int res) {
        float* lse2 = reinterpret_cast<float*>(extra);
        int* lab = reinterpret_cast<int*>(extra + rs::BN * 4);
        float* dlab = reinterpret_cast<float*>(extra + rs::BN * 8);
        const float coef = p.coef * (p.grad_dev != nullptr ? *p.grad_dev : 1.f);
        for (int j = tid; j < rs::BN; j += rs::THREADS) {
            const int b = res * rs::BN + j;
            if (b < p.B) {
                const int y = p.label_local[b];
                lse2[j] = p.lse[b] * LOG2E_B;
                lab[j] = y;
                dlab[j] = (y >= 0) ? -coef * p.one_minus_p[b] * p.dphi[b] : 0.f;
            } else {
                lse2[j] = INFINITY;
                lab[j] = -1;
                dlab[j] = 0.f;
            }
        }
    }

    struct Epi {
        const Params& p;
        const CUtensorMap* tm_out;  // dC^T scratch [chunk classes][Bp] bf16
        rs::Stager stager;
        const float* lse2;
        const int* lab;
        const float* dlab;
        int quad, lane, b0, jl0;
        float coef_all;
        float* qslot;
        __device__ Epi(const Params& prm, const rs::EpiCtx& c)
            : p(prm), tm_out(c.tmC), stager(c), quad(c.quad), lane(c.lane) {
            coef_all = p.coef * (p.grad_dev != nullptr ? *p.grad_dev : 1.f);
            jl0 = c.half * 64;              // first of this warp's 64 columns inside the CTA's 128
            b0 = c.res * rs::BN + jl0;      // the same as a batch index
            lse2 = reinterpret_cast<const float*>(c.extra) + jl0;
            lab = reinterpret_cast<const int*>(c.extra + rs::BN * 4) + jl0;
            dlab = reinterpret_cast<const float*>(c.extra + rs::BN * 8) + jl0;
            qslot = p.q + static_cast<int64_t>(c.res * 2 + c.half) * p.C;
        }
        // 8 consecutive batch columns -> one 16-byte chunk of bf16
        __device__ __forceinline__ void eight(const uint32_t* v, int j0, float coef, int cmatch, float& qacc,
                                              uint32_t (&o)[4]) const {
            const float4 l0 = *reinterpret_cast<const float4*>(lse2 + j0);
            const float4 l1 = *reinterpret_cast<const float4*>(lse2 + j0 + 4);
            const int4 y0 = *reinterpret_cast<const int4*>(lab + j0);
            const int4 y1 = *reinterpret_cast<const int4*>(lab + j0 + 4);
            const float ls[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
            const int ys[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
            float dc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float cosv = __uint_as_float(v[j]);
                float d = coef * ex2(fmaf(cosv, p.s_log2e, -ls[j]));
                if (ys[j] == cmatch) d = dlab[j0 + j];  // rare: this class is row b's label
                qacc = fmaf(d, cosv, qacc);
                dc[j] = d;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = pack_bf16x2(dc[2 * j], dc[2 * j + 1]);
        }
        __device__ void prefetch(int) {}
        __device__ void tile(int i, int, uint32_t taddr) {
            const int c = p.core.s_row0 + i * rs::BM + quad * 32 + lane;  // class owned by this thread
            const bool cvalid = c < p.C;
            if (b0 >= p.Bp) {  // warp-uniform: this warp's columns are all batch padding, nothing to store
                if (cvalid) qslot[c] = 0.f;
                return;
            }
            uint32_t v0[32], v1[32];
            tmem_ld32(taddr, v0);
            tmem_ld32(taddr + 32, v1);
            tmem_ld_wait();
            const float coef = cvalid ? coef_all : 0.f;
            const int cmatch = cvalid ? c : -2;
            float q0 = 0.f, q1 = 0.f;
            stager.acquire();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t o[4];
                eight(v0 + 8 * k, 8 * k, coef, cmatch, q0, o);
                stager.put(k, o[0], o[1], o[2], o[3]);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t o[4];
                eight(v1 + 8 * k, 32 + 8 * k, coef, cmatch, q1, o);
                stager.put(4 + k, o[0], o[1], o[2], o[3]);
            }
            stager.commit(tm_out, b0, i * rs::BM + quad * 32);  // scratch rows are chunk-relative
            if (cvalid) qslot[c] = q0 + q1;
        }
        __device__ void finish() { stager.drain(); }
    };
};

// ------------------------------------------------------------------ dW, Xhat^T slice resident
struct BwdDWr {
    struct Params {
        rs::Core core;  // streamed = dC^T scratch rows (class blocks), resident = 128 rows of xhat^T, K = batch
        int C, D;
        int c_begin;    // first class of this chunk
        const float* q;
        int q_slots;
        const float* inv_nw;
        const __nv_bfloat16* what;
        int64_t ldw;
    };
    static constexpr int EXTRA_BYTES = 0;

    __device__ static void prologue(const Params&, uint8_t*, int, int) {}

    struct Epi {
        const Params& p;
        const CUtensorMap* tm_out;  // dW [C][D] fp32
        rs::Stager stager;
        int quad, lane, d0;
        // per-tile operands of the normalise backward, fetched one tile ahead (their HBM / L2 latency would
        // otherwise sit between every tcgen05.ld and its TMA store)
        uint4 wn[8];
        float qn[8];
        float inwn;
        __device__ Epi(const Params& prm, const rs::EpiCtx& c)
            : p(prm), tm_out(c.tmC), stager(c), quad(c.quad), lane(c.lane), d0(c.res * rs::BN + c.half * 64) {}
        __device__ __forceinline__ void prefetch(int i) {
            const int c = p.c_begin + i * rs::BM + quad * 32 + lane;
            const bool cvalid = c < p.C;
            const __nv_bfloat16* wrow = p.what + static_cast<int64_t>(cvalid ? c : 0) * p.ldw;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const int d = d0 + g * 8;
                wn[g] = (cvalid && d < p.D) ? ldg_nc_u4(wrow + d) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int sl = 0; sl < 8; ++sl)
                qn[sl] = (cvalid && sl < p.q_slots) ? __ldg(p.q + static_cast<int64_t>(sl) * p.C + c) : 0.f;
            inwn = cvalid ? __ldg(p.inv_nw + c) : 0.f;
        }
        __device__ void tile(int i, int i_next, uint32_t taddr) {
            const int crow0 = p.c_begin + i * rs::BM + quad * 32;
            // take over the operands fetched for this tile, then start the next tile's loads
            uint4 w[8];
#pragma unroll
            for (int g = 0; g < 8; ++g) w[g] = wn[g];
            const float nq = -(((qn[0] + qn[1]) + (qn[2] + qn[3])) + ((qn[4] + qn[5]) + (qn[6] + qn[7])));
            const float inw = inwn;
            if (d0 >= p.D) return;  // warp-uniform: columns past the embedding width
            uint32_t v[32];
            tmem_ld32(taddr, v);
            if (i_next >= 0) prefetch(i_next);
            tmem_ld_wait();
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                if (hh == 1) {
                    if (d0 + 32 >= p.D) break;
                    tmem_ld32(taddr + 32, v);
                    tmem_ld_wait();
                }
                stager.acquire();
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint4 ww = w[hh * 4 + g];
                    const float o0 = fmaf(nq, bf16_lo(ww.x), __uint_as_float(v[g * 8 + 0])) * inw;
                    const float o1 = fmaf(nq, bf16_hi(ww.x), __uint_as_float(v[g * 8 + 1])) * inw;
                    const float o2 = fmaf(nq, bf16_lo(ww.y), __uint_as_float(v[g * 8 + 2])) * inw;
                    const float o3 = fmaf(nq, bf16_hi(ww.y), __uint_as_float(v[g * 8 + 3])) * inw;
                    const float o4 = fmaf(nq, bf16_lo(ww.z), __uint_as_float(v[g * 8 + 4])) * inw;
                    const float o5 = fmaf(nq, bf16_hi(ww.z), __uint_as_float(v[g * 8 + 5])) * inw;
                    const float o6 = fmaf(nq, bf16_lo(ww.w), __uint_as_float(v[g * 8 + 6])) * inw;
                    const float o7 = fmaf(nq, bf16_hi(ww.w), __uint_as_float(v[g * 8 + 7])) * inw;
                    stager.put(2 * g, __float_as_uint(o0), __float_as_uint(o1), __float_as_uint(o2), __float_as_uint(o3));
                    stager.put(2 * g + 1, __float_as_uint(o4), __float_as_uint(o5), __float_as_uint(o6), __float_as_uint(o7));
                }
                stager.commit(tm_out, d0 + hh * 32, crow0);  // rows >= C / columns >= D are clipped by the TMA
            }
        }
        __device__ void finish() { stager.drain(); }
    };
};

// ================================================================== CTA-pair fast path
// (gemm_pair.cuh: tcgen05 cta_group::2, 256 x 256 accumulator tile per pair of SMs, reused operand resident)

// ------------------------------------------------------------------ dC^T producer on a CTA pair
// streamed = What rows (256 classes per tile, 128 per CTA -> accumulator lanes), resident = 256 batch rows of
// Xhat (accumulator columns), K = D <= 512.
//
// FUSED = true is the producer role of the single-launch backward (k3_fused.cuh): tile i is the global
// 256-class block i, its dC^T goes to slot i % ring of an L2-resident ring instead of a full-size scratch, and
// the epilogue warps gate on / publish per-block counters (struct Ring).
struct Ring {
    int slots;         // ring slots (each 256 classes x Bp bf16)
    int* ready;        // [blocks] epilogue-warp arrivals of the dC^T producers
    int* done;         // [blocks] consumer arrivals (dW pairs + dX CTAs)
    int ready_target;  // arrivals that complete a block
    int done_target;   // arrivals that free its slot
};

// X3 = true (the bf16x3 precision mode, non-fused launches only): dC leaves as a bf16 pair hi + lo laid out
// [hi | hi | lo] in three blocks of Bp columns, so that the dW GEMM over 3 Bp columns against [xhat_hi^T | xhat_lo^T |
// xhat_hi^T] and three dX launches (hi.W_hi, hi.W_lo, lo.W_hi) keep the gradient GEMMs at ~2^-17 relative as well.
template <bool FUSED, bool X3 = false>
struct BwdDCpT : pr::PairDefaults {
    static constexpr int STAGES = 3;  // 3 x 16 KB: the per-column constants below take the fourth stage's room
    static constexpr bool STAGING = true;
    static constexpr bool RES_A = false;
    static constexpr int NCOL = 2 * pr::ROWS;  // batch columns of one pair

    struct Params {
        pr::Core core;
        int B, C, Bp;
        float s_log2e;
        float coef;
        const float* grad_dev;
        const float* lse;
        const float* one_minus_p;
        const float* dphi;
        const int* label_local;
        float* q;  // [2 * n_res][C]: one slot per (batch slice, column half)
        Ring ring;  // FUSED only
    };
#ifdef AB_K3_L2HINTS
    // experiment (measured 1.2 % SLOWER, profiles/r2_l2hint_exp.log): FUSED: the What rows this role streams are read
    // again a ring depth later (dW correction term, dX operand) -- ask L2 to keep them
    __device__ static uint64_t stream_policy() { return FUSED ? l2_policy_evict_last() : 0ull; }
#endif
    static constexpr int BLOOM_WORDS = 128;  // 4096 bits, one per 128-class block (mod 4096)
    static constexpr int PUB_BAR_OFF = NCOL * 12 + BLOOM_WORDS * 4;  // FUSED: mbarrier epilogue warps -> publisher warp
    static constexpr int EXTRA_BYTES = PUB_BAR_OFF + 16;

    // constants of this pair's 256 batch columns: lse * log2e (+inf on padding -> p = 0), label, label-column dC;
    // plus a bloom filter of the 128-class blocks that hold one of these labels: the epilogue only runs the
    // per-element label test on tiles the filter flags (256 labels among C classes: a few percent of the tiles)
    __device__ static void prologue(const Params& p, uint8_t* extra, int tid, int res, int) {
        float* lse2 = reinterpret_cast<float*>(extra);
        int* lab = reinterpret_cast<int*>(extra + NCOL * 4);
        float* dlab = reinterpret_cast<float*>(extra + NCOL * 8);
        unsigned* bloom = reinterpret_cast<unsigned*>(extra + NCOL * 12);
        const float coef = p.coef * (p.grad_dev != nullptr ? *p.grad_dev : 1.f);
        if (FUSED && tid == 0) {
            mbar_init(reinterpret_cast<uint64_t*>(extra + PUB_BAR_OFF), pr::EPI_WARPS);
            fence_barrier_init();
        }
        for (int j = tid; j < BLOOM_WORDS; j += pr::THREADS) bloom[j] = 0u;
        __syncthreads();  // every thread of the CTA runs the prologue
        for (int j = tid; j < NCOL; j += pr::THREADS) {
            const int bb = res * NCOL + j;
            if (bb < p.B) {
                const int yy = p.label_local[bb];
                if (yy >= 0) atomicOr(&bloom[((yy >> 7) & 4095) >> 5], 1u << ((yy >> 7) & 31));
            }
        }
        for (int j = tid; j < NCOL; j += pr::THREADS) {
            const int b = res * NCOL + j;
            if (b < p.B) {
                const int y = p.label_local[b];
                lse2[j] = p.lse[b] * LOG2E_B;
                lab[j] = y;
                dlab[j] = (y >= 0) ? -coef * p.one_minus_p[b] * p.dphi[b] : 0.f;
            } else {
                lse2[j] = INFINITY;
                lab[j] = -1;
                dlab[j] = 0.f;
            }
        }
    }

    // FUSED: the publisher.  For every tile of the pair's walk: wait until the CTA's eight epilogue warps have
    // handed the block over, fence at gpu scope (cumulative over what the mbarrier made visible) and bump the
    // block's ready counter -- one arrival per CTA.
    __device__ static void side_warp(const Params& p, uint8_t* extra, int i_begin, int i_end, int i_step, int, int lane) {
        if constexpr (FUSED) {
            uint64_t* bar = reinterpret_cast<uint64_t*>(extra + PUB_BAR_OFF);
            uint32_t phase = 0;
            for (int i = i_begin; i < i_end; i += i_step) {
                mbar_wait(bar, phase);
                phase ^= 1;
                __threadfence();
                if (lane == 0) red_relaxed_gpu_add(p.ring.ready + i, 1);
                __syncwarp();
            }
        }
    }

    struct Epi {
        const Params& p;
        const CUtensorMap* tm_out;  // dC^T scratch [chunk classes][Bp] bf16
        pr::Stager stager;
        const float* lse2;
        const int* lab;
        const float* dlab;
        const unsigned* bloom;
        uint64_t* pub_bar;
        int quad, lane, rank, jl0, b0;
        int prev;  // FUSED: block whose stores are committed but not yet published
        float coef_all;
        float* qslot;
        unsigned long long* prof;  // measurements only: [0] publish [1] slot wait [2] tmem load [3] math [4] staging wait
        unsigned long long pacc[5];
        __device__ __forceinline__ unsigned long long tick() const { return prof != nullptr ? clock64() : 0ull; }
        __device__ Epi(const Params& prm, const pr::EpiCtx& c)
            : p(prm), tm_out(c.tmC), stager(c), quad(c.quad), lane(c.lane), rank(c.rank), prev(-1), prof(c.prof) {
            for (int k = 0; k < 5; ++k) pacc[k] = 0;
            coef_all = p.coef * (p.grad_dev != nullptr ? *p.grad_dev : 1.f);
            jl0 = c.half * 128;          // first of this warp's 128 columns inside the pair's 256
            b0 = c.res * NCOL + jl0;     // the same as a batch index
            lse2 = reinterpret_cast<const float*>(c.extra) + jl0;
            lab = reinterpret_cast<const int*>(c.extra + NCOL * 4) + jl0;
            dlab = reinterpret_cast<const float*>(c.extra + NCOL * 8) + jl0;
            bloom = reinterpret_cast<const unsigned*>(c.extra + NCOL * 12);
            pub_bar = reinterpret_cast<uint64_t*>(c.extra + PUB_BAR_OFF);
            qslot = p.q + static_cast<int64_t>(c.res * 2 + c.half) * p.C;
        }
        // 8 consecutive batch columns -> one 16-byte chunk of bf16.  LABELS = false skips the label test (the
        // caller knows none of the CTA's 128 classes is a label of these batch columns).
        template <bool LABELS>
        __device__ __forceinline__ void eight(const uint32_t* v, int j0, float coef, int cmatch, float& qa, float& qb,
                                              uint32_t* o, uint32_t* ol) const {
            const float4 l0 = *reinterpret_cast<const float4*>(lse2 + j0);
            const float4 l1 = *reinterpret_cast<const float4*>(lse2 + j0 + 4);
            const float ls[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
            float dc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float cosv = __uint_as_float(v[j]);
                float d = coef * ex2(fmaf(cosv, p.s_log2e, -ls[j]));
                if constexpr (LABELS) {
                    if (lab[j0 + j] == cmatch) d = dlab[j0 + j];  // rare: this class is row b's label
                }
                if (j & 1) qb = fmaf(d, cosv, qb);
                else qa = fmaf(d, cosv, qa);
                dc[j] = d;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                o[j] = pack_bf16x2(dc[2 * j], dc[2 * j + 1]);
                if constexpr (X3)   // lo = bf16(d - hi): the two halves of the packed word are the hi values
                    ol[j] = pack_bf16x2(dc[2 * j] - bf16_lo(o[j]), dc[2 * j + 1] - bf16_hi(o[j]));
            }
        }
        template <bool LABELS>
        __device__ __forceinline__ void group(uint32_t taddr, int g, int row0, float coef, int cmatch, float (&q)[4]) {
            uint32_t v0[32], v1[32];
            const unsigned long long t0 = tick();
            tmem_ld32(taddr + g * 64, v0);
            tmem_ld32(taddr + g * 64 + 32, v1);
            tmem_ld_wait();
            const unsigned long long t1 = tick();
            uint32_t o[32];  // 64 columns of bf16
            uint32_t ol[X3 ? 32 : 1];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                eight<LABELS>(v0 + 8 * k, g * 64 + 8 * k, coef, cmatch, q[0], q[1], o + 4 * k, ol + (X3 ? 4 * k : 0));
#pragma unroll
            for (int k = 0; k < 4; ++k)
                eight<LABELS>(v1 + 8 * k, g * 64 + 32 + 8 * k, coef, cmatch, q[2], q[3], o + 16 + 4 * k,
                              ol + (X3 ? 16 + 4 * k : 0));
            const unsigned long long t2 = tick();
            stager.acquire();  // only now: the previous box has had the whole computation above to leave
            const unsigned long long t3 = tick();
#pragma unroll
            for (int k = 0; k < 8; ++k) stager.put(k, o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
            if constexpr (X3) {
                // the hi box goes to column blocks 0 and 1 (one staging, two TMA stores), the lo box to block 2
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(tm_out, stager.buf, b0 + g * 64, row0);
                    tma_store_2d(tm_out, stager.buf, p.Bp + b0 + g * 64, row0);
                    bulk_commit();
                }
                stager.acquire();
#pragma unroll
                for (int k = 0; k < 8; ++k) stager.put(k, ol[4 * k], ol[4 * k + 1], ol[4 * k + 2], ol[4 * k + 3]);
                stager.commit(tm_out, 2 * p.Bp + b0 + g * 64, row0);
            } else {
                stager.commit(tm_out, b0 + g * 64, row0);  // rows past the chunk are clipped by the TMA
            }
            if (prof != nullptr) { pacc[2] += t1 - t0; pacc[3] += t2 - t1; pacc[4] += t3 - t2; }
        }
        __device__ void prefetch(int) {}
        // FUSED: hand the previous block over to the CTA's publisher warp (side_warp below).  Its TMA stores were
        // committed by lane 0 a whole tile ago, its q values written by every lane; the gpu-scope fence and the
        // counter update -- ~2700 cycles when the epilogue warps did them themselves -- happen off this path.
        __device__ __forceinline__ void hand_over() {
            // the dC^T boxes have been written (completion of the bulk group makes them visible to this thread;
            // the release chain below -- mbarrier arrive, publisher's gpu fence, counter -- carries them on)
            if (lane == 0) bulk_wait<0>();
            __syncwarp();  // every lane's q stores are ordered before lane 0's arrive
            if (lane == 0) mbar_arrive(pub_bar);
        }
        __device__ void tile(int i, int, uint32_t taddr) {
            const int row0 = (FUSED ? (i % p.ring.slots) : i) * NCOL + rank * pr::ROWS + quad * 32;  // scratch row
            const int c0 = p.core.s_row0 + i * NCOL + rank * pr::ROWS;  // first class of this CTA's 128
            const int c = c0 + quad * 32 + lane;                        // class owned by this thread
            const bool cvalid = c < p.C;
            const float coef = cvalid ? coef_all : 0.f;
            const int cmatch = cvalid ? c : -2;
            float q[4] = {0.f, 0.f, 0.f, 0.f};
            if constexpr (FUSED) {
                const unsigned long long t0 = tick();
                if (prev >= 0) hand_over();
                prev = i;
                const unsigned long long t1 = tick();
                if (i >= p.ring.slots) {  // the slot's previous tenant must have been consumed
                    if (lane == 0) wait_counter_ge(p.ring.done + (i - p.ring.slots), p.ring.done_target);
                    __syncwarp();
                }
                if (prof != nullptr) { pacc[0] += t1 - t0; pacc[1] += tick() - t1; }
            }
            // does one of the two 128-class blocks this CTA's classes touch hold a label of our batch columns?
            const int k0 = (c0 >> 7) & 4095, k1 = ((c0 + pr::ROWS - 1) >> 7) & 4095;
            const bool labels = ((bloom[k0 >> 5] >> (k0 & 31)) | (bloom[k1 >> 5] >> (k1 & 31))) & 1u;
#pragma unroll 1
            for (int g = 0; g < 2; ++g) {
                if (b0 + g * 64 >= p.Bp) break;  // warp-uniform: these columns are all batch padding
                if (labels) group<true>(taddr, g, row0, coef, cmatch, q);
                else group<false>(taddr, g, row0, coef, cmatch, q);
            }
            if (cvalid) qslot[c] = (q[0] + q[1]) + (q[2] + q[3]);
        }
        __device__ void finish() {
            if constexpr (FUSED) {
                if (prev >= 0) hand_over();
            }
            if (prof != nullptr && lane == 0)
                for (int k = 0; k < 5; ++k) prof[k] = pacc[k];
            stager.drain();
        }
    };
};
using BwdDCp = BwdDCpT<false>;
using BwdDCp3 = BwdDCpT<false, true>;

// ------------------------------------------------------------------ dW on a CTA pair
// streamed = dC^T scratch rows (256 classes per tile), resident = 256 rows of Xhat^T (embedding columns of the
// accumulator), K = batch <= 512.
// FUSED = true: consumer role of the single-launch backward: the streamed rows come out of the ring once the
// block's counter says every dC^T producer warp has published it; the slot is released when the last operand
// byte of the block has landed in shared memory.
__device__ __forceinline__ float ld_cg_f32(const float* p) {  // data written earlier in this kernel by another SM
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

template <bool FUSED>
struct BwdDWpT : pr::PairDefaults {
    static constexpr int STAGES = 4;
    // The dW rows leave through one 4 KB staging buffer per warp and TMA stores of {32 columns x 32 rows}.
    // Measured alternatives (ARCFACE_B200_BWD_PROF, cycles per tile in the store part of this epilogue; this path:
    // ~4.2k): two 2 KB buffers with 16-column boxes double the number of proxy fences (~640 cycles per box: the
    // fence waits for st.shared traffic that competes with the tensor pipe's operand reads for shared-memory
    // bandwidth); 128-bit global stores straight from registers (thread = class row) are 4x slower still (32
    // half-sectors per warp instruction); round 2: re-reading the staging buffer with eight lanes per row and
    // storing full 128-byte lines with st.global.v4 (no proxy fence, no TMA) ~7.8k; 256-bit stores straight from
    // registers (STG.E.ENL2.256, one full sector per lane) ~8.8k; sixteen epilogue warps (four per lane quadrant, 2 KB
    // boxes of 16 columns, registers traded with setmaxnreg) ~9.8k per tile in total.  Every variant lands at
    // 13-19 bytes per clock and SM.  tools/probes/ explains why: an SM that does nothing else stores 32 B/clk
    // (store_probe), lane pairs writing 64 contiguous bytes per row reach that rate straight from registers
    // (store_pattern_probe; tried here too: 9.8k cycles per tile), but the rate falls to 10-19 B/clk as soon as the same
    // SM also pulls data in (store_load_probe) -- and this role's SMs pull in 192 KB per tile (the dC^T operand and the
    // normalised weights of the correction term) for the 128 KB they write.  A dW tile therefore costs >= ~8k cycles
    // against 4.1k of MMA whatever issues the stores: the role is bound by its SMs' memory port, which is why the
    // role split gives it the largest share; the epilogue's form is not the lever.
    static constexpr bool STAGING = true;
    static constexpr bool RES_A = false;
    static constexpr int NCOL = 2 * pr::ROWS;

    struct Params {
        pr::Core core;
        int C, D;
        int c_begin;  // first class of this chunk (the scratch is chunk-relative)
        const float* q;
        int q_slots;  // <= 8
        const float* inv_nw;
        const __nv_bfloat16* what;
        int64_t ldw;
        float* dw;  // out [C][D] fp32
        Ring ring;  // FUSED only
        int evict_first;  // 1: dW is stored with an L2 evict-first policy (it is never read back by this kernel)
    };
    static constexpr int EXTRA_BYTES = 16;  // FUSED: [0] = 1 + the last block this CTA's producer has acquired

    __device__ static void prologue(const Params&, uint8_t* extra, int tid, int, int) {
        if (tid == 0) *reinterpret_cast<volatile int*>(extra) = 0;
    }
    __device__ static int stream_row(const Params& p, int i) {
        return FUSED ? (i % p.ring.slots) * NCOL : p.core.s_row0 + i * NCOL;
    }
    __device__ static void acquire_tile(const Params& p, uint8_t* extra, int i, int, int lane) {
        if constexpr (FUSED) {
            if (lane == 0) {
                wait_counter_ge(p.ring.ready + i, p.ring.ready_target);
                __threadfence_block();
                *reinterpret_cast<volatile int*>(extra) = i + 1;  // lets the epilogue fetch q[i] ahead of its tile
            }
            __syncwarp();
            fence_proxy_async_all();  // the acquire above -> ordered before the TMA reads of the block
        }
    }
    __device__ static void release_tile(const Params& p, int i) {
        if constexpr (FUSED) red_relaxed_gpu_add(p.ring.done + i, 1);
    }

    struct Epi {
        const Params& p;
        const CUtensorMap* tm_out;  // dW [C][D] fp32, boxes of 32 columns x 32 rows
        pr::Stager stager;
        int quad, lane, rank, d0;
        // operands of the normalise backward, fetched ahead of their use (their L2 latency would otherwise sit
        // between every tcgen05.ld and its TMA store): the normalised weights TWO 32-column groups ahead (across
        // the tile boundary), the per-class scalars one tile ahead (FUSED: at the start of the tile -- q of a later
        // block may not have been produced yet).  The scalars stay raw in registers until they are needed, so
        // that nothing in program order waits on their loads.
        uint4 wa[4], wb[4];  // groups g and g + 1 in flight / ready
        float qn[8];
        float inwn;
        const volatile int* acquired;  // FUSED: see EXTRA_BYTES
        bool have_scalars;
        bool w256;  // the rows of what are 32-byte aligned: the correction term's weights are fetched with 256-bit loads
        uint64_t pol;
        unsigned long long* prof;  // measurements only: [0] scalars [1] tmem load [2] math [3] staging wait [4] store issue
        unsigned long long pacc[5];
        __device__ __forceinline__ unsigned long long tick() const { return prof != nullptr ? clock64() : 0ull; }
        __device__ Epi(const Params& prm, const pr::EpiCtx& c)
            : p(prm), tm_out(c.tmC), stager(c), quad(c.quad), lane(c.lane), rank(c.rank),
              d0(c.res * NCOL + c.half * 128), acquired(reinterpret_cast<const volatile int*>(c.extra)),
              have_scalars(false), prof(c.prof) {
            for (int k = 0; k < 5; ++k) pacc[k] = 0;
            pol = p.evict_first ? l2_policy_evict_first() : 0ull;
            w256 = ((reinterpret_cast<uintptr_t>(p.what) | static_cast<uintptr_t>(p.ldw * 2)) & 31u) == 0;
        }
        __device__ __forceinline__ int class_of(int i) const {
            return p.c_begin + i * NCOL + rank * pr::ROWS + quad * 32 + lane;
        }
        // the 32 normalised weights of group `grp` (0..3) of tile i
        __device__ __forceinline__ void fetch_w(int i, int grp, uint4 (&w)[4]) const {
            const int c = class_of(i);
            const bool cvalid = c < p.C;
            const __nv_bfloat16* wrow = p.what + static_cast<int64_t>(cvalid ? c : 0) * p.ldw;
#ifdef AB_EXP_DW_NOW   // timing experiment only (wrong gradients): no loads for the correction term
#pragma unroll
            for (int g = 0; g < 4; ++g) w[g] = make_uint4(0, 0, 0, static_cast<uint32_t>(reinterpret_cast<uintptr_t>(wrow)) & 1u);
#else
#ifndef AB_DW_NO_W256
            // one full 32-byte sector per lane and instruction (LDG.256) instead of two half sectors: a lane's row is 1 KB
            // away from its neighbour's, so every 16-byte load costs the L1 a whole sector request
            if (w256) {
#pragma unroll
                for (int g = 0; g < 4; g += 2) {
                    const int d = d0 + grp * 32 + g * 8;
                    if (cvalid && d + 16 <= p.D) ldg_nc_u8(wrow + d, w[g], w[g + 1]);
                    else {
                        w[g] = (cvalid && d < p.D) ? ldg_nc_u4(wrow + d) : make_uint4(0, 0, 0, 0);
                        w[g + 1] = (cvalid && d + 8 < p.D) ? ldg_nc_u4(wrow + d + 8) : make_uint4(0, 0, 0, 0);
                    }
                }
                return;
            }
#endif
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int d = d0 + grp * 32 + g * 8;
                w[g] = (cvalid && d < p.D) ? ldg_nc_u4(wrow + d) : make_uint4(0, 0, 0, 0);
            }
#endif
        }
        __device__ __forceinline__ void fetch_scalars(int i) {
            const int c = class_of(i);
            const bool cvalid = c < p.C;
#pragma unroll
            for (int sl = 0; sl < 8; ++sl)
                qn[sl] = (cvalid && sl < p.q_slots) ? ld_cg_f32(p.q + static_cast<int64_t>(sl) * p.C + c) : 0.f;
            inwn = cvalid ? __ldg(p.inv_nw + c) : 0.f;
        }
        __device__ void prefetch(int i) {  // before the first tile
            if constexpr (!FUSED) fetch_scalars(i);
            fetch_w(i, 0, wa);
            fetch_w(i, 1, wb);
        }
        // one 32-column group: w = its weights (consumed), refilled with the group two ahead
        __device__ __forceinline__ void group(int i, int i_next, int grp, uint32_t taddr, int crow0, float nq, float inw,
                                              uint4 (&w)[4]) {
            const int dg = d0 + grp * 32;
            uint4 wc[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) wc[g] = w[g];
            if (dg >= p.D) {  // warp-uniform: columns past the embedding width; only keep the prefetch chain going
                if (grp < 2) fetch_w(i, grp + 2, w);
                else if (i_next >= 0) fetch_w(i_next, grp - 2, w);
                return;
            }
            uint32_t v[32];
            const unsigned long long t0 = tick();
            tmem_ld32(taddr + grp * 32, v);
            // refill: group grp + 2 of this tile, or group grp - 2 of the next one
            if (grp < 2) fetch_w(i, grp + 2, w);
            else if (i_next >= 0) fetch_w(i_next, grp - 2, w);
            tmem_ld_wait();
            const unsigned long long t1 = tick();
            float o[32];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint4 ww = wc[g];
                o[g * 8 + 0] = fmaf(nq, bf16_lo(ww.x), __uint_as_float(v[g * 8 + 0])) * inw;
                o[g * 8 + 1] = fmaf(nq, bf16_hi(ww.x), __uint_as_float(v[g * 8 + 1])) * inw;
                o[g * 8 + 2] = fmaf(nq, bf16_lo(ww.y), __uint_as_float(v[g * 8 + 2])) * inw;
                o[g * 8 + 3] = fmaf(nq, bf16_hi(ww.y), __uint_as_float(v[g * 8 + 3])) * inw;
                o[g * 8 + 4] = fmaf(nq, bf16_lo(ww.z), __uint_as_float(v[g * 8 + 4])) * inw;
                o[g * 8 + 5] = fmaf(nq, bf16_hi(ww.z), __uint_as_float(v[g * 8 + 5])) * inw;
                o[g * 8 + 6] = fmaf(nq, bf16_lo(ww.w), __uint_as_float(v[g * 8 + 6])) * inw;
                o[g * 8 + 7] = fmaf(nq, bf16_hi(ww.w), __uint_as_float(v[g * 8 + 7])) * inw;
            }
            const unsigned long long t2 = tick();
#ifdef AB_EXP_DW_NOSTORE   // timing experiment only (dW is not written): the epilogue without its store path
            float acc_x = 0.f;
#pragma unroll
            for (int k = 0; k < 32; ++k) acc_x += o[k];
            if (acc_x == 1.2345e-30f) stager.put(0, __float_as_uint(acc_x), 0, 0, 0);
            const unsigned long long t3 = tick();
#else
            stager.acquire();  // only now: the previous box has had the tcgen05.ld and the math above to leave
            const unsigned long long t3 = tick();
#pragma unroll
            for (int k = 0; k < 8; ++k)
                stager.put(k, __float_as_uint(o[4 * k]), __float_as_uint(o[4 * k + 1]), __float_as_uint(o[4 * k + 2]),
                           __float_as_uint(o[4 * k + 3]));
            stager.commit(tm_out, dg, crow0, pol);  // rows >= C / columns >= D are clipped by the TMA
#endif
            if (prof != nullptr) { pacc[1] += t1 - t0; pacc[2] += t2 - t1; pacc[3] += t3 - t2; pacc[4] += tick() - t3; }
        }
        __device__ void tile(int i, int i_next, uint32_t taddr) {
            const int crow0 = p.c_begin + i * NCOL + rank * pr::ROWS + quad * 32;
            const unsigned long long ts = tick();
            if constexpr (FUSED) {
                // q[i] is published (the block's accumulator is complete); normally it was fetched a tile ago
                if (!have_scalars) fetch_scalars(i);
            }
            const float nq = -(((qn[0] + qn[1]) + (qn[2] + qn[3])) + ((qn[4] + qn[5]) + (qn[6] + qn[7])));
            const float inw = inwn;
            if (prof != nullptr) pacc[0] += (nq == 123.456f ? 1 : 0) + tick() - ts;  // (uses nq: the loads must have landed)
            if constexpr (!FUSED) {
                if (i_next >= 0) fetch_scalars(i_next);
            } else {
                // the next block's q may be fetched now if this CTA's producer has already seen it published
                have_scalars = false;
                if (i_next >= 0) {
                    int a = 0;
                    if (lane == 0) a = *acquired;
                    a = __shfl_sync(0xffffffffu, a, 0);
                    if (a > i_next) {
                        __threadfence_block();
                        fetch_scalars(i_next);
                        have_scalars = true;
                    }
                }
            }
            group(i, i_next, 0, taddr, crow0, nq, inw, wa);
            group(i, i_next, 1, taddr, crow0, nq, inw, wb);
            group(i, i_next, 2, taddr, crow0, nq, inw, wa);
            group(i, i_next, 3, taddr, crow0, nq, inw, wb);
        }
        __device__ void finish() {
            if (prof != nullptr && lane == 0)
                for (int k = 0; k < 5; ++k) prof[k] = pacc[k];
            stager.drain();
        }
    };
};
using BwdDWp = BwdDWpT<false>;

// ------------------------------------------------------------------ dXhat, 256 x 256 output tile per CTA
// Two stacked 128-row batch sub-tiles share every What k-block (64 KB per 1024 tensor cycles instead of
// 48 KB per 512), the class range is split across CTAs and each CTA keeps its partial product in all 512
// TMEM columns until its range is done: one reduce-add epilogue per CTA.
struct BwdDX2 {
    static constexpr int BLOCK_N = 256;
    static constexpr int STAGES = 3;
    static constexpr int M_SUB = 2;
    static constexpr int ACC_BUFS = 1;
    static constexpr bool STAGING = true;
    static constexpr bool A_MN = true;  // dC^T chunk [classes = K][batch = M contiguous]
    static constexpr bool B_MN = true;  // what [classes = K][D = N contiguous]

    struct Params {
        int B, D;
        int c_begin;
        int m_tiles, dn_tiles, splits;  // m_tiles counts 256-row tiles
        int kb_total;
        int kb_per_split;
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
            t.m0 = mt * (BLOCK_M * M_SUB);
            t.n0 = dn * BLOCK_N;
            t.ka0 = kb0 * BLOCK_K;
            t.kb0 = p.c_begin + kb0 * BLOCK_K;
            t.kblocks = min(p.kb_per_split, p.kb_total - kb0);
            t.aux = 0;
            idx += step;
            return true;
        }
    };

    struct Epi {
        const Params& p;
        const CUtensorMap* tm_out;  // dXhat [B][D] fp32, accumulated with TMA reduce-add
        StoreStager stager;
        int ew, lane;
        __device__ Epi(const Params& prm, const EpiCtx& c) : p(prm), tm_out(c.tmC), stager(c), ew(c.ew), lane(c.lane) {}
        __device__ void tile(const Tile& t, uint32_t taddr) {
#pragma unroll 1
            for (int ms = 0; ms < M_SUB; ++ms) {
                const int row0 = t.m0 + ms * BLOCK_M + ew * 32;
                if (row0 >= p.B) break;  // warp-uniform: these 32 rows (and all later ones) are padding
#pragma unroll 1
                for (int cc = 0; cc < BLOCK_N / 32; ++cc) {
                    const int d0 = t.n0 + cc * 32;
                    if (d0 >= p.D) break;
                    uint32_t v[32];
                    tmem_ld32(taddr + ms * BLOCK_N + cc * 32, v);
                    tmem_ld_wait();
                    const uint32_t buf = stager.acquire();
#pragma unroll
                    for (int g = 0; g < 8; ++g) stager.put(buf, g, v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                    stager.commit<true>(tm_out, buf, d0, row0);  // rows >= B / columns >= D are clipped by the TMA
                }
            }
        }
        __device__ void finish() { stager.drain(); }
    };
};

}  // namespace ab
#include "k3_fused.cuh"
namespace ab {

// dXhat = sum over the class splits' partial tiles, in split order (fixed: the result does not depend on which split
// finished first).  parts: [n_parts][part_stride4 float4], the first n4 float4 of each are the [B][D] rows.
__global__ void __launch_bounds__(256) sum_dx_parts_kernel(const float4* __restrict__ parts, int n_parts, int64_t part_stride4,
                                                           int64_t n4, float4* __restrict__ out) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        float4 a = parts[i];
        for (int s = 1; s < n_parts; ++s) {
            const float4 b = parts[s * part_stride4 + i];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        out[i] = a;
    }
}

struct BwdPlan {
    int Bp;             // scratch leading dimension (batch rounded up to 64)
    int chunk_classes;  // classes per scratch chunk (multiple of 128)
    int n_chunks;
    bool dc_rs, dw_rs, dx2;  // which kernels take the resident-operand / stacked-tile fast path
    bool dc_pair, dw_pair;   // ... on CTA pairs (cta_group::2) instead of single CTAs
    // single-launch backward (k3_fused.cuh): role split in CTA pairs, ring slots, counters
    bool fused;
    bool dx_pairs;           // single-launch backward: the dX role runs on CTA pairs
    int n_dc, n_dw, n_dx, ring_slots, n_blocks;
    size_t cnt_off, cnt_bytes;
    // single-launch backward: per-split dX tiles ([dx_parts][dx_part_rows][D] fp32), summed in split order afterwards
    int dx_parts, dx_part_rows;
    size_t dxp_off, dxp_bytes;
    int q_slots;             // partial-sum slots of q per class
    size_t scratch_off, scratch_bytes, q_off, q_bytes, total;
    int kx;                  // column blocks of the dC^T scratch: 1, or 3 = [hi | hi | lo] (bf16x3 mode on CTA pairs)
};

// CTA pairs of bwd_fused_kernel that can be resident at once on the current device (the roles spin on each
// other's counters, so the whole launch has to be).  0 on failure.
static size_t fused_smem_bytes() {
    // every role is sized against the whole 227 KB (resident mode: MAX_KBLOCKS + STAGES tiles; streaming mode: as
    // many two-tile stages as fit), so the launch simply asks for all of it
    static_assert(pr::smem_bytes<BwdDCpT<true>>(pr::MAX_KBLOCKS, BwdDCpT<true>::EXTRA_BYTES) <= 227 * 1024, "dC^T role");
    static_assert(pr::smem_bytes<BwdDWpT<true>>(pr::MAX_KBLOCKS, BwdDWpT<true>::EXTRA_BYTES) <= 227 * 1024, "dW role");
    static_assert(fz::dx_smem_bytes() <= 227 * 1024, "dX role");
    static_assert(fz::dxp_smem_bytes() <= 227 * 1024, "dX role on CTA pairs");
    return 227 * 1024;
}
static int fused_max_pairs(int nsm) {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    if (cached[dev] == 0) {
        const size_t smem = fused_smem_bytes();
        if (smem > 227 * 1024) return 0;
        if (cudaFuncSetAttribute(fz::bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
            cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(static_cast<unsigned>(nsm / 2 * 2));
        cfg.blockDim = dim3(pr::THREADS);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, fz::bwd_fused_kernel, &cfg) != cudaSuccess) {
            cudaGetLastError();
            n = 0;
        }
        if (n > nsm / 2) n = nsm / 2;
        cached[dev] = n > 0 ? n : -1;
    }
    return cached[dev] > 0 ? cached[dev] : 0;
}

static bool env_is(const char* name, const char* value) {  // diagnostic builds only: always false in the release library
    const char* v = diag_env(name);
    return v != nullptr && strcmp(v, value) == 0;
}

// Kernel selection and scratch chunking.  Environment knobs (diagnostics / A-B measurements only):
//   ARCFACE_B200_BWD_IMPL=generic   use the streaming kernels of gemm_core.cuh for every shape
//   ARCFACE_B200_BWD_IMPL=rs        resident-operand kernels on single CTAs (gemm_rs.cuh) instead of CTA pairs
//   ARCFACE_B200_BWD_IMPL=split     CTA-pair kernels as three launches through a full-size scratch
//   ARCFACE_B200_BWD_SPLIT=a,b,c    CTA pairs given to the dC^T / dW / dX roles of the single-launch backward
//   ARCFACE_B200_BWD_RING=<n>       ring slots of the single-launch backward
//   ARCFACE_B200_BWD_CHUNK_MB=<n>   cap of the dC^T scratch in MiB
// Ds = depth of the S^T recompute (D; 3 D in the bf16x3 mode): only the role split depends on it.
static BwdPlan plan_backward(int B, int D, int64_t C, int nsm, int Ds = 0) {
    if (Ds <= 0) Ds = D;
    BwdPlan pl;
    pl.Bp = ((B + 63) / 64) * 64;
    if (nsm < 1) nsm = 148;
    const bool generic = env_is("ARCFACE_B200_BWD_IMPL", "generic");
    pl.dc_rs = !generic && (Ds + 63) / 64 <= rs::MAX_KBLOCKS;
    pl.dw_rs = !generic && pl.Bp / 64 <= rs::MAX_KBLOCKS;
    pl.dx2 = !generic;
    // CTA pairs for every shape whose q partial sums fit the dW epilogue's eight slots (B <= 1024): the reused operand
    // is parked in shared memory when it fits (D <= 512 for dC^T, B <= 512 for dW) and streamed otherwise
    const bool pairs = !generic && !env_is("ARCFACE_B200_BWD_IMPL", "rs") && nsm >= 2 &&
                       2 * ((B + BwdDCp::NCOL - 1) / BwdDCp::NCOL) <= 8;
    pl.dc_pair = pairs;
    pl.dw_pair = pairs;
    pl.q_slots = pl.dc_pair ? 2 * ((B + BwdDCp::NCOL - 1) / BwdDCp::NCOL) : pl.dc_rs ? 2 * ((B + rs::BN - 1) / rs::BN) : 1;
    // ---- single-launch backward: every role needs its resident operand to fit (D, B <= 512) and the launch
    // needs enough co-resident CTA pairs to give each role a few
    pl.fused = false;
    pl.dx_pairs = false;
    pl.n_dc = pl.n_dw = pl.n_dx = pl.ring_slots = 0;
    pl.n_blocks = static_cast<int>((C + 255) / 256);
    // bf16x3 mode (Ds = 3 D) on CTA pairs: hi/lo dC through a three-block scratch, three launches per chunk (the ring
    // of the single-launch backward holds one block per slot)
    pl.kx = (Ds > D && pl.dc_pair && pl.dw_pair) ? 3 : 1;
    if (pl.dc_pair && pl.dw_pair && pl.kx == 1 && !env_is("ARCFACE_B200_BWD_IMPL", "split")) {
        const int n_res_dc = (B + 255) / 256, n_res_dw = (D + 255) / 256;
        const int tiles = n_res_dc * n_res_dw;         // dX output tiles (256 x 256)
        // dX role: a CTA pair per 256 x 512 tile (k3_fused.cuh dx_pair_body), or -- diagnostic builds,
        // ARCFACE_B200_BWD_DX=cta -- the round-1 role with one CTA per 256 x 256 tile
        // (measured per BASELINE shape, profiles/r2_dx_pair_exp.log: pairs 1-5 % faster except at D = 2816, whose eleven
        // 256-column slices leave the sixth pair tile half empty and make a split twelve pairs coarse: 0.900 vs 0.862 ms)
        pl.dx_pairs = !env_is("ARCFACE_B200_BWD_DX", "cta") && !(n_res_dw % 2 == 1 && n_res_dw > 8);
        if (env_is("ARCFACE_B200_BWD_DX", "pair")) pl.dx_pairs = true;
        const int dx_unit = pl.dx_pairs ? n_res_dc * ((D + 511) / 512) : (tiles + 1) / 2;   // CTA pairs per dX split
        const int pairs = fused_max_pairs(nsm);
        int a = 0, b = 0, c = 0;
        if (const char* v = diag_env("ARCFACE_B200_BWD_SPLIT")) sscanf(v, "%d,%d,%d", &a, &b, &c);
        if (a <= 0 || b <= 0 || c <= 0 || a + b + c > pairs) {
            // Role shares from a cost model in pair-cycles per 256-class block, fitted to the wait profiler
            // (ARCFACE_B200_BWD_PROF) at the north-star shape and checked against role-split sweeps at the BASELINE
            // shapes (profiles/README.md):
            //   dC^T: one 256 x 256 x D tile per 256 batch rows -- 8 D MMA cycles, but never below the ~6.6k cycles its
            //         exp2 / pack / TMA-store epilogue takes;
            //   dW  : one 256 x 256 x B tile per 256 embedding columns -- 8 B MMA cycles, never below the ~9k cycles of
            //         its fp32 store epilogue (the role that small batches starve);
            //   dX  : ~5.05k cycles per block and 256 x 256 output tile on ONE CTA.
            // A role that streams both operands (K > 512) reaches ~70 % of its MMA rate.  The split minimises the
            // slowest role's time per block over the admissible pair counts.
            a = b = c = 0;
            const double mma_dc = 8.0 * (((Ds + 63) / 64) * 64) / ((Ds + 63) / 64 > pr::MAX_KBLOCKS ? 0.7 : 1.0);
            const double mma_dw = 8.0 * pl.Bp / (pl.Bp / 64 > pr::MAX_KBLOCKS ? 0.7 : 1.0);
            const double cost_dc = n_res_dc * (mma_dc > 6650.0 ? mma_dc : 6650.0);
            const double cost_dw = n_res_dw * (mma_dw > 8950.0 ? mma_dw : 8950.0);
            // dX: ~5.05k cycles per block and 256 x 256 tile on one CTA; on a pair ~2.3k pair-cycles per 256 x 256 of output
            const double cost_dx = pl.dx_pairs ? tiles * AB_DXP_COST : tiles * 5053.0 / 2.0;
            double best = 1e300;
            for (int aa = n_res_dc; aa <= pairs; aa += n_res_dc)
                for (int bb = n_res_dw; aa + bb <= pairs; bb += n_res_dw) {
                    const int cc = (pairs - aa - bb) / dx_unit * dx_unit;
                    if (cc < dx_unit) break;
                    const double ta = cost_dc / aa, tb = cost_dw / bb, tc = cost_dx / cc;
                    const double t = ta > tb ? (ta > tc ? ta : tc) : (tb > tc ? tb : tc);
                    if (t < best) { best = t; a = aa; b = bb; c = cc; }
                }
        }
        a = a / n_res_dc * n_res_dc;
        b = b / n_res_dw * n_res_dw;
        c = c / dx_unit * dx_unit;
        if (a >= n_res_dc && b >= n_res_dw && c >= dx_unit && a + b + c <= pairs && pl.n_blocks >= 1) {
            pl.fused = true;
            pl.n_dc = a; pl.n_dw = b; pl.n_dx = c;
            int slots = 64;
            if (const char* v = diag_env("ARCFACE_B200_BWD_RING")) slots = atoi(v);
            const int min_slots = 2 * (a / n_res_dc) + 2;  // producers may run a tile or two ahead of the consumers
            if (slots < min_slots) slots = min_slots;
            if (slots > pl.n_blocks) slots = pl.n_blocks;
            pl.ring_slots = slots;
        }
    }
    if (pl.fused) {
        pl.q_slots = 2 * ((B + BwdDCp::NCOL - 1) / BwdDCp::NCOL);
        pl.chunk_classes = static_cast<int>(((C + 255) / 256) * 256);
        pl.n_chunks = 1;
        pl.scratch_off = 0;
        pl.scratch_bytes = static_cast<size_t>(pl.ring_slots) * 256 * pl.Bp * 2;
        pl.q_off = (pl.scratch_bytes + 255) / 256 * 256;
        pl.q_bytes = static_cast<size_t>(C) * 4 * pl.q_slots;
        pl.cnt_off = pl.q_off + (pl.q_bytes + 255) / 256 * 256;
        // ring counters (two per block) + the dX regions' arrival counters (16 per 256 x 512 of dX at most)
        pl.cnt_bytes = (static_cast<size_t>(pl.n_blocks) * 2 + static_cast<size_t>(((B + 255) / 256) * ((D + 255) / 256) * 16)) *
                       sizeof(int);
        {
            const int n_res_dc = (B + 255) / 256, n_res_dw = (D + 255) / 256;
            const int splits = pl.dx_pairs ? pl.n_dx / (n_res_dc * ((D + 511) / 512)) : (2 * pl.n_dx) / (n_res_dc * n_res_dw);
            pl.dx_parts = splits < pl.n_blocks ? splits : pl.n_blocks;   // splits beyond the block count have no work
            pl.dx_part_rows = n_res_dc * 256;
        }
        pl.dxp_off = pl.cnt_off + (pl.cnt_bytes + 255) / 256 * 256;
        pl.dxp_bytes = static_cast<size_t>(pl.dx_parts) * pl.dx_part_rows * D * sizeof(float);
        pl.total = pl.dxp_off + (pl.dxp_bytes + 255) / 256 * 256;
        return pl;
    }
    pl.cnt_off = pl.cnt_bytes = 0;
    pl.dx_parts = pl.dx_part_rows = 0;
    pl.dxp_off = pl.dxp_bytes = 0;
    const int64_t c_round = ((C + 127) / 128) * 128;
    int64_t chunk;
    size_t cap = generic ? (size_t(64) << 20) : (size_t(2048) << 20);
    if (const char* v = diag_env("ARCFACE_B200_BWD_CHUNK_MB")) {
        const long mb = atol(v);
        if (mb >= 1) cap = static_cast<size_t>(mb) << 20;
    }
    if (generic) {
        // whole waves of 128-class blocks (one per SM) while the chunk stays under the cap (L2-friendly)
        const size_t wave_bytes = static_cast<size_t>(128) * nsm * pl.Bp * 2;
        size_t k = cap / wave_bytes;
        if (k < 1) k = 1;
        chunk = static_cast<int64_t>(128) * nsm * static_cast<int64_t>(k);
        if (chunk > c_round) chunk = c_round;
    } else {
        // the scratch goes through HBM once; as few launches as the cap allows, evenly sized
        int64_t max_chunk = static_cast<int64_t>(cap / (static_cast<size_t>(pl.Bp) * 2 * pl.kx)) / 128 * 128;
        if (max_chunk < 128) max_chunk = 128;
        const int64_t n = (c_round + max_chunk - 1) / max_chunk;
        chunk = ((c_round / 128 + n - 1) / n) * 128;
    }
    pl.chunk_classes = static_cast<int>(chunk);
    pl.n_chunks = static_cast<int>((C + chunk - 1) / chunk);
    pl.scratch_off = 0;
    pl.scratch_bytes = static_cast<size_t>(chunk) * pl.Bp * 2 * pl.kx;
    pl.q_off = (pl.scratch_bytes + 255) / 256 * 256;
    pl.q_bytes = static_cast<size_t>(C) * 4 * pl.q_slots;
    pl.total = pl.q_off + (pl.q_bytes + 255) / 256 * 256;
    return pl;
}

}  // namespace ab

using namespace ab;

static int32_t check_bwd_shape(const char* who, int32_t B, int32_t D, int64_t C_local) {
    AB_REQUIRE(B >= 1 && B <= ARCFACE_B200_MAX_BATCH, ARCFACE_B200_E_SHAPE, "%s: B=%d outside [1, %d]", who, B,
               ARCFACE_B200_MAX_BATCH);
    AB_REQUIRE(D >= 8 && D % 8 == 0, ARCFACE_B200_E_SHAPE, "%s: D=%d must be a positive multiple of 8", who, D);
    AB_REQUIRE(C_local >= 1 && C_local <= (1ll << 30), ARCFACE_B200_E_SHAPE, "%s: C_local=%lld outside [1, 2^30]", who,
               (long long)C_local);
    return ARCFACE_B200_OK;
}

extern "C" int32_t arcface_b200_backward_workspace_bytes(int32_t B, int32_t D, int64_t C_local, size_t* bytes) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(bytes, ARCFACE_B200_E_ARG, "backward_workspace_bytes: null pointer");
    if (int32_t rc = check_bwd_shape("backward_workspace_bytes", B, D, C_local)) return rc;
    // enough for either precision mode (the role split, and with it the ring size, depends on the recompute depth)
    const size_t a = plan_backward(B, D, C_local, sm_count()).total;
    const size_t b = plan_backward(B, D, C_local, sm_count(), 3 * D).total;
    *bytes = a > b ? a : b;
    return ARCFACE_B200_OK;
}

extern "C" int32_t arcface_b200_backward_plan(int32_t B, int32_t D, int64_t C_local, int64_t* chunk_classes,
                                              int32_t* n_chunks) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(chunk_classes && n_chunks, ARCFACE_B200_E_ARG, "backward_plan: null pointer");
    if (int32_t rc = check_bwd_shape("backward_plan", B, D, C_local)) return rc;
    const BwdPlan pl = plan_backward(B, D, C_local, sm_count());
    *chunk_classes = pl.chunk_classes;
    *n_chunks = pl.n_chunks;
    return ARCFACE_B200_OK;
}

extern "C" int32_t arcface_b200_backward_launches(int32_t B, int32_t D, int64_t C_local, int32_t* n_kernels) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(n_kernels, ARCFACE_B200_E_ARG, "backward_launches: null pointer");
    if (int32_t rc = check_bwd_shape("backward_launches", B, D, C_local)) return rc;
    const BwdPlan pl = plan_backward(B, D, C_local, sm_count());
    *n_kernels = pl.fused ? 1 : 3 * pl.n_chunks;
    return ARCFACE_B200_OK;
}

static int32_t backward_impl(const uint16_t* xhat, const uint16_t* xhat_t, int64_t ld_t,
                             const uint16_t* what, const float* inv_nw, const float* lse,
                             const float* one_minus_p, const float* dphi, const int32_t* label_local,
                             int32_t B, int32_t D, int64_t C_local, float s, float grad_scale,
                             const float* grad_loss_dev, float* dxhat, float* dw, void* workspace,
                             size_t workspace_bytes, int32_t prec, void* stream);

extern "C" int32_t arcface_b200_backward(const uint16_t* xhat, const uint16_t* xhat_t, int64_t ld_t,
                                         const uint16_t* what, const float* inv_nw, const float* lse,
                                         const float* one_minus_p, const float* dphi, const int32_t* label_local,
                                         int32_t B, int32_t D, int64_t C_local, float s, float grad_scale,
                                         const float* grad_loss_dev, float* dxhat, float* dw, void* workspace,
                                         size_t workspace_bytes, void* stream) {
    return backward_impl(xhat, xhat_t, ld_t, what, inv_nw, lse, one_minus_p, dphi, label_local, B, D, C_local, s,
                         grad_scale, grad_loss_dev, dxhat, dw, workspace, workspace_bytes, ARCFACE_B200_PREC_BF16, stream);
}

extern "C" int32_t arcface_b200_backward_prec(const uint16_t* xhat, const uint16_t* xhat_t, int64_t ld_t,
                                              const uint16_t* what, const float* inv_nw, const float* lse,
                                              const float* one_minus_p, const float* dphi, const int32_t* label_local,
                                              int32_t B, int32_t D, int64_t C_local, float s, float grad_scale,
                                              const float* grad_loss_dev, float* dxhat, float* dw, void* workspace,
                                              size_t workspace_bytes, int32_t prec, void* stream) {
    AB_REQUIRE(prec == ARCFACE_B200_PREC_BF16 || prec == ARCFACE_B200_PREC_BF16X3, ARCFACE_B200_E_ARG,
               "backward: unknown precision mode %d", prec);
    return backward_impl(xhat, xhat_t, ld_t, what, inv_nw, lse, one_minus_p, dphi, label_local, B, D, C_local, s,
                         grad_scale, grad_loss_dev, dxhat, dw, workspace, workspace_bytes, prec, stream);
}

// prec = ARCFACE_B200_PREC_BF16X3: `xhat` / `what` are the 3 D wide rows of arcface_b200_normalize_cast3 ([hi|hi|lo] /
// [hi|lo|hi]).  The probabilities are recomputed from the three-term product (contraction depth 3 D, the same
// operands the forward used), so dC carries no bf16 logit noise; the two gradient GEMMs run on dC (bf16) and the hi
// parts (xhat_t, the first D columns of `what`).
static int32_t backward_impl(const uint16_t* xhat, const uint16_t* xhat_t, int64_t ld_t,
                             const uint16_t* what, const float* inv_nw, const float* lse,
                             const float* one_minus_p, const float* dphi, const int32_t* label_local,
                             int32_t B, int32_t D, int64_t C_local, float s, float grad_scale,
                             const float* grad_loss_dev, float* dxhat, float* dw, void* workspace,
                             size_t workspace_bytes, int32_t prec, void* stream) {
    const int Ds = prec == ARCFACE_B200_PREC_BF16X3 ? 3 * D : D;   // depth of the S^T recompute
    const int64_t ldw = Ds;                                         // row stride of xhat / what
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(xhat && xhat_t && what && inv_nw && lse && one_minus_p && dphi && label_local && dxhat && dw && workspace,
               ARCFACE_B200_E_ARG, "backward: null pointer");
    if (int32_t rc = check_bwd_shape("backward", B, D, C_local)) return rc;
    AB_REQUIRE(ld_t >= B && ld_t % 8 == 0, ARCFACE_B200_E_LAYOUT, "backward: ld_t must be >= B and a multiple of 8");
    AB_REQUIRE(aligned16(dxhat) && aligned16(dw) && aligned16(workspace) && aligned16(what), ARCFACE_B200_E_LAYOUT,
               "backward: pointers must be 16-byte aligned");
    AB_REQUIRE(s > 0.f, ARCFACE_B200_E_ARG, "backward: scale s must be positive");
    const int nsm = sm_count();
    const BwdPlan pl = plan_backward(B, D, C_local, nsm, Ds);
    AB_REQUIRE(workspace_bytes >= pl.total, ARCFACE_B200_E_WORKSPACE, "backward: workspace %zu < required %zu",
               workspace_bytes, pl.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    __nv_bfloat16* dct = reinterpret_cast<__nv_bfloat16*>(ws + pl.scratch_off);
    float* q = reinterpret_cast<float*>(ws + pl.q_off);
    const int C = static_cast<int>(C_local);

    // How the single-launch backward combines the class splits' dX tiles: 0 = fp32 TMA reduce-adds into a zeroed dXhat
    // (order not fixed), 1 = per-split tiles + a sum kernel in split order, 2 = per-split tiles summed inside the kernel by
    // the last warp to arrive at a region.  1 and 2 are bit-reproducible.
    int dx_sum = AB_DX_SUM_MODE;
    if (const char* v = diag_env("ARCFACE_B200_BWD_DXSUM")) dx_sum = atoi(v);
    if (!pl.fused || dx_sum == 0) AB_CHECK_CUDA(cudaMemsetAsync(dxhat, 0, static_cast<size_t>(B) * D * sizeof(float), st));

    CUtensorMap tm_w_k, tm_x_k, tm_dct_k, tm_xt_k, tm_dct_mn, tm_w_mn, tm_dct_out, tm_dw_out, tm_dx_out;
    if (int32_t rc = make_tmap_kmajor(&tm_w_k, what, Ds, C_local, ldw, BLOCK_M)) return rc;
    if (int32_t rc = make_tmap_kmajor(&tm_x_k, xhat, Ds, B, ldw, (pl.dc_rs || pl.dc_pair) ? rs::BN : BwdDC::BLOCK_N)) return rc;
    const bool x3 = pl.kx == 3;   // hi/lo dC: every dC^T map spans three blocks of Bp columns
    const int64_t ldc = static_cast<int64_t>(pl.Bp) * pl.kx;
    if (x3) {
        AB_REQUIRE(ld_t == pl.Bp, ARCFACE_B200_E_LAYOUT,
                   "backward (bf16x3): xhat_t must be [D][3 * ld_t] with ld_t = %d (the batch rounded up to 64)", pl.Bp);
        // the batch padding inside each block is zero on both sides (written by the dC^T epilogue / the caller)
        if (int32_t rc = make_tmap_kmajor(&tm_dct_k, dct, 3 * pl.Bp, pl.chunk_classes, ldc, BLOCK_M)) return rc;
        if (int32_t rc = make_tmap_kmajor(&tm_xt_k, xhat_t, 3 * pl.Bp, D, 3 * ld_t, rs::BN)) return rc;
    } else {
        if (int32_t rc = make_tmap_kmajor(&tm_dct_k, dct, B, pl.chunk_classes, pl.Bp, BLOCK_M)) return rc;
        if (int32_t rc = make_tmap_kmajor(&tm_xt_k, xhat_t, B, D, ld_t, (pl.dw_rs || pl.dw_pair) ? rs::BN : BwdDW::BLOCK_N)) return rc;
    }
    if (int32_t rc = make_tmap_mnmajor(&tm_dct_mn, dct, B, pl.chunk_classes, ldc)) return rc;
    if (int32_t rc = make_tmap_mnmajor(&tm_w_mn, what, D, C_local, ldw)) return rc;
    if (int32_t rc = make_tmap_store(&tm_dct_out, dct, 2, ldc, pl.chunk_classes, ldc)) return rc;
    if (int32_t rc = make_tmap_store(&tm_dw_out, dw, 4, D, C_local, D)) return rc;
    if (int32_t rc = make_tmap_store(&tm_dx_out, dxhat, 4, D, B, D)) return rc;

    const int n_tiles = (B + BwdDC::BLOCK_N - 1) / BwdDC::BLOCK_N;
    const int dn_tiles = (D + 255) / 256;

    if (pl.fused) {
        int* counters = reinterpret_cast<int*>(ws + pl.cnt_off);
        AB_CHECK_CUDA(cudaMemsetAsync(counters, 0, pl.cnt_bytes, st));
        Ring ring;
        ring.slots = pl.ring_slots;
        ring.ready = counters;
        ring.done = counters + pl.n_blocks;
        const int n_res_dc = (B + 255) / 256, n_res_dw = (D + 255) / 256;
        ring.ready_target = n_res_dc * 2;  // one arrival per producer CTA (its publisher warp)
        const int dn512 = (D + 511) / 512;
        // consumers of a block: one dW pair per 256 embedding columns + the dX CTAs / pairs of the block's split
        ring.done_target = n_res_dw + (pl.dx_pairs ? n_res_dc * dn512 : n_res_dc * n_res_dw);
        fz::FusedParams fp;
        fp.dx_pairs = pl.dx_pairs ? 1 : 0;
        fp.n_dc = pl.n_dc;
        fp.n_dw = pl.n_dw;
        fp.n_dx = pl.n_dx;
        {
            // interleave the roles over the launch order (largest deficit first)
            const int total = pl.n_dc + pl.n_dw + pl.n_dx;
            AB_REQUIRE(total <= fz::MAX_PAIRS, ARCFACE_B200_E_SHAPE, "backward: %d CTA pairs exceed the role table", total);
            const int want[3] = {pl.n_dc, pl.n_dw, pl.n_dx};
            int got[3] = {0, 0, 0};
            const bool blocked = env_is("ARCFACE_B200_BWD_ORDER", "blocked");  // A/B: roles in contiguous ranges
            for (int pi = 0; pi < total; ++pi) {
                int best = -1;
                double best_def = -1e30;
                for (int k = 0; k < 3; ++k) {
                    if (got[k] >= want[k]) continue;
                    const double def = blocked ? -k : static_cast<double>(want[k]) * (pi + 1) / total - got[k];
                    if (def > best_def) { best_def = def; best = k; }
                }
                fp.role[pi] = static_cast<uint8_t>(best);
                fp.index[pi] = static_cast<uint8_t>(got[best]++);
            }
        }
        {
            BwdDCpT<true>::Params& p = fp.dc;
            pr::core_set_k<BwdDCpT<true>>(p.core, Ds, BwdDCpT<true>::EXTRA_BYTES);
            p.core.s_blocks = pl.n_blocks;
            p.core.s_row0 = 0;
            p.core.n_res = n_res_dc;
            p.core.contiguous = 0;
            p.core.prefetch_tiles = 2;
            p.B = B; p.C = C; p.Bp = pl.Bp;
            p.s_log2e = s * LOG2E_B; p.coef = s * grad_scale; p.grad_dev = grad_loss_dev;
            p.lse = lse; p.one_minus_p = one_minus_p; p.dphi = dphi; p.label_local = label_local;
            p.q = q;
            p.ring = ring;
        }
        {
            BwdDWpT<true>::Params& p = fp.dw;
            pr::core_set_k<BwdDWpT<true>>(p.core, pl.Bp, BwdDWpT<true>::EXTRA_BYTES);
            p.core.s_blocks = pl.n_blocks;
            p.core.s_row0 = 0;
            p.core.n_res = n_res_dw;
            p.core.contiguous = 0;
            p.core.prefetch_tiles = 0;  // the ring lives in L2
            p.C = C; p.D = D; p.c_begin = 0;
            p.q = q; p.q_slots = pl.q_slots; p.inv_nw = inv_nw;
            p.what = reinterpret_cast<const __nv_bfloat16*>(what);
            p.ldw = ldw;
            p.dw = dw;
            // dW lines are never read back by this kernel: evict-first keeps them from pushing the ring and the What
            // rows out of L2 (1.481 -> 1.472 ms, alternated twice; ARCFACE_B200_BWD_EVICT=0 in diagnostic builds: off)
            p.evict_first = env_is("ARCFACE_B200_BWD_EVICT", "0") ? 0 : 1;
            p.ring = ring;
        }
        {
            fz::DXParams& p = fp.dx;
            p.B = B; p.D = D;
            p.n_blocks = pl.n_blocks;
            p.m_tiles = n_res_dc;
            if (pl.dx_pairs) {
                p.dn_tiles = dn512;
                p.splits = pl.n_dx / (n_res_dc * dn512);
            } else {
                p.dn_tiles = n_res_dw;
                p.splits = (2 * pl.n_dx) / (n_res_dc * n_res_dw);
            }
            p.part_rows = dx_sum == 0 ? 0 : pl.dx_part_rows;
            p.n_parts = pl.dx_parts;
            p.parts = reinterpret_cast<const float*>(ws + pl.dxp_off);
            p.region_cnt = dx_sum == 2 ? counters + 2 * pl.n_blocks : nullptr;
            p.dx_out = dxhat;
            p.ring = ring;
        }
        // dX: every class split stores its partial tile; the last split to arrive at a region adds them in split order
        // (k3_fused.cuh dx_region_done: bit-reproducible, no zero-fill, no fp32 reduce-adds in L2)
        float* dx_parts = reinterpret_cast<float*>(ws + pl.dxp_off);
        CUtensorMap tm_dxp_out;
        if (int32_t rc = make_tmap_store(&tm_dxp_out, dx_parts, 4, D, static_cast<uint64_t>(pl.dx_parts) * pl.dx_part_rows, D))
            return rc;
        // the ring replaces the full-size scratch: [slots * 256][Bp] bf16
        CUtensorMap tm_ring_out, tm_ring_k, tm_ring_mn;
        const int64_t ring_rows = static_cast<int64_t>(pl.ring_slots) * 256;
        if (int32_t rc = make_tmap_store(&tm_ring_out, dct, 2, pl.Bp, ring_rows, pl.Bp)) return rc;
        if (int32_t rc = make_tmap_kmajor(&tm_ring_k, dct, B, ring_rows, pl.Bp, pr::ROWS)) return rc;
        if (int32_t rc = make_tmap_mnmajor(&tm_ring_mn, dct, B, ring_rows, pl.Bp)) return rc;
        const size_t smem = fused_smem_bytes();
        const int grid = 2 * (pl.n_dc + pl.n_dw + pl.n_dx);
#ifdef ARCFACE_B200_DIAG
        // diagnostic build only: ARCFACE_B200_BWD_PROF=1 prints where each role's warps waited (allocates, synchronises)
        static unsigned long long* prof_dev = nullptr;
        const bool prof = env_is("ARCFACE_B200_BWD_PROF", "1");
        if (prof) {
            if (prof_dev == nullptr) AB_CHECK_CUDA(cudaMalloc(&prof_dev, 2 * fz::MAX_PAIRS * 16 * sizeof(unsigned long long)));
            AB_CHECK_CUDA(cudaMemsetAsync(prof_dev, 0, 2 * fz::MAX_PAIRS * 16 * sizeof(unsigned long long), st));
            fp.dc.core.prof = prof_dev; fp.dc.core.prof_cta = 0;
            fp.dw.core.prof = prof_dev; fp.dw.core.prof_cta = 2 * pl.n_dc;
            fp.dx.prof = prof_dev; fp.dx.prof_cta = 2 * (pl.n_dc + pl.n_dw);
        }
#endif
        fz::bwd_fused_kernel<<<grid, pr::THREADS, smem, st>>>(tm_w_k, tm_x_k, tm_ring_out, tm_ring_k, tm_xt_k, tm_dw_out,
                                                             tm_ring_mn, tm_w_mn, dx_sum == 0 ? tm_dx_out : tm_dxp_out, fp);
        AB_CHECK_CUDA(cudaGetLastError());
        if (dx_sum == 1) {
            const int64_t n4 = static_cast<int64_t>(B) * D / 4;
            const int64_t want = (n4 + 255) / 256;
            sum_dx_parts_kernel<<<static_cast<int>(want < 148 * 8 ? want : 148 * 8), 256, 0, st>>>(
                reinterpret_cast<const float4*>(dx_parts), pl.dx_parts, static_cast<int64_t>(pl.dx_part_rows) * D / 4, n4,
                reinterpret_cast<float4*>(dxhat));
            AB_CHECK_CUDA(cudaGetLastError());
        }
#ifdef ARCFACE_B200_DIAG
        if (prof) {
            static unsigned long long host[2 * fz::MAX_PAIRS * 16];
            AB_CHECK_CUDA(cudaStreamSynchronize(st));
            AB_CHECK_CUDA(cudaMemcpy(host, prof_dev, sizeof(host), cudaMemcpyDeviceToHost));
            const char* names[3] = {"dC^T", "dW", "dX"};
            const int first[4] = {0, 2 * pl.n_dc, 2 * (pl.n_dc + pl.n_dw), grid};
            for (int r = 0; r < 3; ++r) {
                double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                int n = 0;
                for (int c = first[r]; c < first[r + 1]; ++c) {
                    if (r < 2 && ((c - first[r]) & 1)) continue;  // pair roles: the leader CTA carries the MMA counters
                    for (int k = 0; k < 8; ++k) a[k] += static_cast<double>(host[c * 16 + k]);
                    ++n;
                }
                if (n == 0) continue;
                for (int k = 0; k < 8; ++k) a[k] /= n;
                fprintf(stderr,
                        "[bwd prof] %-5s ctas=%d tiles/cta=%.1f body=%.0f cyc | producer: flags %.0f, free stage %.0f | "
                        "mma: operands %.0f, free acc %.0f | epi warp: acc wait %.0f, in tile %.0f | per tile: body %.0f, "
                        "epi %.0f\n",
                        names[r], first[r + 1] - first[r], a[7], a[6], a[0], a[1], a[2], a[3], a[4], a[5],
                        a[7] > 0 ? a[6] / a[7] : 0.0, a[7] > 0 ? a[5] / a[7] : 0.0);
                if (r == 1 && a[7] > 0) {
                    double e[5] = {0, 0, 0, 0, 0};
                    for (int c = first[1]; c < first[2]; c += 2)
                        for (int k = 0; k < 5; ++k) e[k] += static_cast<double>(host[c * 16 + 8 + k]);
                    fprintf(stderr, "[bwd prof] dW epilogue per tile: scalars %.0f, tmem load %.0f, math %.0f, staging wait %.0f, "
                                    "store issue %.0f\n",
                            e[0] / n / a[7], e[1] / n / a[7], e[2] / n / a[7], e[3] / n / a[7], e[4] / n / a[7]);
                }
                if (r == 0 && a[7] > 0) {
                    double e[5] = {0, 0, 0, 0, 0};
                    for (int c = first[0]; c < first[1]; c += 2)
                        for (int k = 0; k < 5; ++k) e[k] += static_cast<double>(host[c * 16 + 8 + k]);
                    fprintf(stderr, "[bwd prof] dC^T epilogue per tile: publish %.0f, slot wait %.0f, tmem load %.0f, math %.0f, "
                                    "staging wait %.0f\n",
                            e[0] / n / a[7], e[1] / n / a[7], e[2] / n / a[7], e[3] / n / a[7], e[4] / n / a[7]);
                }
            }
        }
#endif
        return ARCFACE_B200_OK;
    }

    for (int64_t c0 = 0; c0 < C_local; c0 += pl.chunk_classes) {
        const int cn = static_cast<int>(C_local - c0 < pl.chunk_classes ? C_local - c0 : pl.chunk_classes);
        const int c_blocks = (cn + BLOCK_M - 1) / BLOCK_M;
        // ---- dC^T (and q) for this chunk
        if (x3) {
            BwdDCp3::Params p;
            pr::core_set_k<BwdDCp3>(p.core, Ds, BwdDCp3::EXTRA_BYTES);
            p.core.s_blocks = (cn + BwdDCp3::NCOL - 1) / BwdDCp3::NCOL;
            p.core.s_row0 = static_cast<int>(c0);
            p.core.n_res = (B + BwdDCp3::NCOL - 1) / BwdDCp3::NCOL;
            p.core.contiguous = 0;
            p.core.prefetch_tiles = 2;
            p.B = B; p.C = C; p.Bp = pl.Bp;
            p.s_log2e = s * LOG2E_B; p.coef = s * grad_scale; p.grad_dev = grad_loss_dev;
            p.lse = lse; p.one_minus_p = one_minus_p; p.dphi = dphi; p.label_local = label_local;
            p.q = q;
            int groups = (nsm / 2) / p.core.n_res;
            if (groups < 1) groups = 1;
            if (groups > p.core.s_blocks) groups = p.core.s_blocks;
            if (int32_t rc = pr::launch_pair<BwdDCp3>(tm_w_k, tm_x_k, tm_dct_out, p, groups, BwdDCp3::EXTRA_BYTES, st))
                return rc;
        } else if (pl.dc_pair) {
            BwdDCp::Params p;
            pr::core_set_k<BwdDCp>(p.core, Ds, BwdDCp::EXTRA_BYTES);
            p.core.s_blocks = (cn + BwdDCp::NCOL - 1) / BwdDCp::NCOL;
            p.core.s_row0 = static_cast<int>(c0);
            p.core.n_res = (B + BwdDCp::NCOL - 1) / BwdDCp::NCOL;
            p.core.contiguous = 0;
            p.core.prefetch_tiles = 2;
            p.B = B; p.C = C; p.Bp = pl.Bp;
            p.s_log2e = s * LOG2E_B; p.coef = s * grad_scale; p.grad_dev = grad_loss_dev;
            p.lse = lse; p.one_minus_p = one_minus_p; p.dphi = dphi; p.label_local = label_local;
            p.q = q;
            int groups = (nsm / 2) / p.core.n_res;
            if (groups < 1) groups = 1;
            if (groups > p.core.s_blocks) groups = p.core.s_blocks;
            if (int32_t rc = pr::launch_pair<BwdDCp>(tm_w_k, tm_x_k, tm_dct_out, p, groups, BwdDCp::EXTRA_BYTES, st))
                return rc;
        } else if (pl.dc_rs) {
            BwdDCr::Params p;
            p.core.kblocks = (Ds + rs::BK - 1) / rs::BK;
            p.core.m_blocks = c_blocks;
            p.core.s_row0 = static_cast<int>(c0);
            p.core.n_res = (B + rs::BN - 1) / rs::BN;
            p.core.prefetch_tiles = 2;
            p.B = B; p.C = C; p.Bp = pl.Bp;
            p.s_log2e = s * LOG2E_B; p.coef = s * grad_scale; p.grad_dev = grad_loss_dev;
            p.lse = lse; p.one_minus_p = one_minus_p; p.dphi = dphi; p.label_local = label_local;
            p.q = q;
            int groups = nsm / p.core.n_res;
            if (groups < 1) groups = 1;
            if (groups > c_blocks) groups = c_blocks;
            if (int32_t rc = rs::launch_rs<BwdDCr>(tm_w_k, tm_x_k, tm_dct_out, p, groups, BwdDCr::EXTRA_BYTES, st))
                return rc;
        } else {
            BwdDC::Params p;
            p.B = B; p.D = Ds; p.C = C; p.Bp = pl.Bp;
            p.c_begin = static_cast<int>(c0); p.c_blocks = c_blocks; p.n_tiles = n_tiles;
            p.s_log2e = s * LOG2E_B; p.coef = s * grad_scale; p.grad_dev = grad_loss_dev;
            p.lse = lse; p.one_minus_p = one_minus_p; p.dphi = dphi; p.label_local = label_local;
            p.q = q;
            const int grid = c_blocks < nsm ? c_blocks : nsm;
            if (int32_t rc = launch_gemm<BwdDC>(tm_w_k, tm_x_k, tm_dct_out, p, grid, BwdDC::extra_bytes(n_tiles), st))
                return rc;
        }
        // ---- dW rows of this chunk
        if (pl.dw_pair) {
            BwdDWp::Params p;
            pr::core_set_k<BwdDWp>(p.core, pl.Bp * pl.kx, BwdDWp::EXTRA_BYTES);
            p.core.s_blocks = (cn + BwdDWp::NCOL - 1) / BwdDWp::NCOL;
            p.core.s_row0 = 0;  // the scratch is chunk-relative
            p.core.n_res = (D + BwdDWp::NCOL - 1) / BwdDWp::NCOL;
            p.core.contiguous = 0;
            p.core.prefetch_tiles = 2;
            p.C = C; p.D = D; p.c_begin = static_cast<int>(c0);
            p.q = q; p.q_slots = pl.q_slots; p.inv_nw = inv_nw;
            p.what = reinterpret_cast<const __nv_bfloat16*>(what);
            p.ldw = ldw;
            p.dw = dw;
            // dW lines are never read back by this kernel: evict-first keeps them from pushing the ring and the What
            // rows out of L2 (1.481 -> 1.472 ms, alternated twice; ARCFACE_B200_BWD_EVICT=0 in diagnostic builds: off)
            p.evict_first = env_is("ARCFACE_B200_BWD_EVICT", "0") ? 0 : 1;
            int groups = (nsm / 2) / p.core.n_res;
            if (groups < 1) groups = 1;
            if (groups > p.core.s_blocks) groups = p.core.s_blocks;
            if (int32_t rc = pr::launch_pair<BwdDWp>(tm_dct_k, tm_xt_k, tm_dw_out, p, groups, BwdDWp::EXTRA_BYTES, st))
                return rc;
        } else if (pl.dw_rs) {
            BwdDWr::Params p;
            p.core.kblocks = pl.Bp / rs::BK;
            p.core.m_blocks = c_blocks;
            p.core.s_row0 = 0;  // the scratch is chunk-relative
            p.core.n_res = (D + rs::BN - 1) / rs::BN;
            p.core.prefetch_tiles = 2;
            p.C = C; p.D = D; p.c_begin = static_cast<int>(c0);
            p.q = q; p.q_slots = pl.q_slots; p.inv_nw = inv_nw;
            p.what = reinterpret_cast<const __nv_bfloat16*>(what);
            p.ldw = ldw;
            int groups = nsm / p.core.n_res;
            if (groups < 1) groups = 1;
            if (groups > c_blocks) groups = c_blocks;
            if (int32_t rc = rs::launch_rs<BwdDWr>(tm_dct_k, tm_xt_k, tm_dw_out, p, groups, BwdDWr::EXTRA_BYTES, st))
                return rc;
        } else {
            BwdDW::Params p;
            p.B = B; p.D = D; p.C = C;
            p.c_begin = static_cast<int>(c0); p.c_blocks = c_blocks; p.dn_tiles = dn_tiles;
            p.q = q; p.q_slots = pl.q_slots; p.inv_nw = inv_nw; p.what = reinterpret_cast<const __nv_bfloat16*>(what);
            p.ldw = ldw;
            const int total = c_blocks * dn_tiles;
            const int grid = total < nsm ? total : nsm;
            if (int32_t rc = launch_gemm<BwdDW>(tm_dct_k, tm_xt_k, tm_dw_out, p, grid, 0, st)) return rc;
        }
        // ---- dXhat += this chunk's contribution
        if (pl.dx2) {
            BwdDX2::Params p;
            p.B = B; p.D = D; p.c_begin = static_cast<int>(c0);
            p.m_tiles = (B + 2 * BLOCK_M - 1) / (2 * BLOCK_M); p.dn_tiles = dn_tiles;
            p.kb_total = (c_blocks * BLOCK_M) / BLOCK_K;
            int splits = nsm / (p.m_tiles * dn_tiles);
            if (splits < 1) splits = 1;
            if (splits > p.kb_total) splits = p.kb_total;
            p.kb_per_split = (p.kb_total + splits - 1) / splits;
            p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
            const int total = p.m_tiles * dn_tiles * p.splits;
            const int grid = total < nsm ? total : nsm;
            if (x3) {
                // dXhat += dC_hi . W_hi + dC_hi . W_lo + dC_lo . W_hi: three launches over column-block views of the
                // scratch ([hi | hi | lo]) and of what ([hi | lo | hi]), all reduce-adding into the same dXhat
                const __nv_bfloat16* wb = reinterpret_cast<const __nv_bfloat16*>(what);
                const __nv_bfloat16* a_view[3] = {dct, dct, dct + 2 * pl.Bp};
                const __nv_bfloat16* b_view[3] = {wb, wb + D, wb};
                for (int v = 0; v < 3; ++v) {
                    CUtensorMap ta, tb;
                    if (int32_t rc = make_tmap_mnmajor(&ta, a_view[v], B, pl.chunk_classes, ldc)) return rc;
                    if (int32_t rc = make_tmap_mnmajor(&tb, b_view[v], D, C_local, ldw)) return rc;
                    if (int32_t rc = launch_gemm<BwdDX2>(ta, tb, tm_dx_out, p, grid, 0, st)) return rc;
                }
            } else if (int32_t rc = launch_gemm<BwdDX2>(tm_dct_mn, tm_w_mn, tm_dx_out, p, grid, 0, st)) {
                return rc;
            }
        } else {
            BwdDX::Params p;
            p.B = B; p.D = D; p.c_begin = static_cast<int>(c0);
            p.m_tiles = (B + BLOCK_M - 1) / BLOCK_M; p.dn_tiles = dn_tiles;
            p.kb_total = (c_blocks * BLOCK_M) / BLOCK_K;
            int splits = nsm / (p.m_tiles * dn_tiles);
            if (splits < 1) splits = 1;
            if (splits > p.kb_total) splits = p.kb_total;
            p.kb_per_split = (p.kb_total + splits - 1) / splits;
            p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
            const int total = p.m_tiles * dn_tiles * p.splits;
            const int grid = total < nsm ? total : nsm;
            if (int32_t rc = launch_gemm<BwdDX>(tm_dct_mn, tm_w_mn, tm_dx_out, p, grid, 0, st)) return rc;
        }
    }
    if (x3)   // the dW epilogue projected with what_hi only: add the lo part
        return launch_dw_lo_correction(dw, what, q, pl.q_slots, inv_nw, C_local, D, st);
    return ARCFACE_B200_OK;
}
