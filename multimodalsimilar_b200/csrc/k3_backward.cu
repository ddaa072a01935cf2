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
#include <stdlib.h>
#include <string.h>

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
            const __nv_bfloat16* wrow = p.what + static_cast<int64_t>(cvalid ? c : 0) * p.D;
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
            const __nv_bfloat16* wrow = p.what + static_cast<int64_t>(cvalid ? c : 0) * p.D;
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
struct BwdDCp {
    static constexpr int STAGES = 3;  // 3 x 16 KB: the per-column constants below take the fourth stage's room
    static constexpr int AUX_WARPS = 0;
    static constexpr int LOW_REGS = 0, EPI_REGS = 0, AUX_REGS = 0;  // no helper warps, no register reallocation
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
    };
    static constexpr int EXTRA_BYTES = NCOL * 12;

    // constants of this pair's 256 batch columns: lse * log2e (+inf on padding -> p = 0), label, label-column dC
    __device__ static void prologue(const Params& p, uint8_t* extra, int tid, int res, int) {
        float* lse2 = reinterpret_cast<float*>(extra);
        int* lab = reinterpret_cast<int*>(extra + NCOL * 4);
        float* dlab = reinterpret_cast<float*>(extra + NCOL * 8);
        const float coef = p.coef * (p.grad_dev != nullptr ? *p.grad_dev : 1.f);
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

    __device__ static void acquire_tile(const Params&, int, int, int) {}
    __device__ static void aux(const Params&, int, int, int) {}

    struct Epi {
        const Params& p;
        const CUtensorMap* tm_out;  // dC^T scratch [chunk classes][Bp] bf16
        pr::Stager stager;
        const float* lse2;
        const int* lab;
        const float* dlab;
        int quad, lane, rank, jl0, b0;
        float coef_all;
        float* qslot;
        __device__ Epi(const Params& prm, const pr::EpiCtx& c)
            : p(prm), tm_out(c.tmC), stager(c), quad(c.quad), lane(c.lane), rank(c.rank) {
            coef_all = p.coef * (p.grad_dev != nullptr ? *p.grad_dev : 1.f);
            jl0 = c.half * 128;          // first of this warp's 128 columns inside the pair's 256
            b0 = c.res * NCOL + jl0;     // the same as a batch index
            lse2 = reinterpret_cast<const float*>(c.extra) + jl0;
            lab = reinterpret_cast<const int*>(c.extra + NCOL * 4) + jl0;
            dlab = reinterpret_cast<const float*>(c.extra + NCOL * 8) + jl0;
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
            const int row0 = i * NCOL + rank * pr::ROWS + quad * 32;  // chunk-relative class row of lane 0
            const int c = p.core.s_row0 + row0 + lane;                // class owned by this thread
            const bool cvalid = c < p.C;
            const float coef = cvalid ? coef_all : 0.f;
            const int cmatch = cvalid ? c : -2;
            float q0 = 0.f, q1 = 0.f;
#pragma unroll 1
            for (int g = 0; g < 2; ++g) {
                if (b0 + g * 64 >= p.Bp) break;  // warp-uniform: these columns are all batch padding
                uint32_t v0[32], v1[32];
                tmem_ld32(taddr + g * 64, v0);
                tmem_ld32(taddr + g * 64 + 32, v1);
                tmem_ld_wait();
                stager.acquire();
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t o[4];
                    eight(v0 + 8 * k, g * 64 + 8 * k, coef, cmatch, q0, o);
                    stager.put(k, o[0], o[1], o[2], o[3]);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t o[4];
                    eight(v1 + 8 * k, g * 64 + 32 + 8 * k, coef, cmatch, q1, o);
                    stager.put(4 + k, o[0], o[1], o[2], o[3]);
                }
                stager.commit(tm_out, b0 + g * 64, row0);  // rows past the chunk are clipped by the TMA
            }
            if (cvalid) qslot[c] = q0 + q1;
        }
        __device__ void finish() { stager.drain(); }
    };
};

// ------------------------------------------------------------------ dW on a CTA pair
// streamed = dC^T scratch rows (256 classes per tile), resident = 256 rows of Xhat^T (embedding columns of the
// accumulator), K = batch <= 512.
struct BwdDWp {
    static constexpr int STAGES = 4;
    static constexpr int AUX_WARPS = 0;
    static constexpr int LOW_REGS = 0, EPI_REGS = 0, AUX_REGS = 0;  // no helper warps, no register reallocation
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
    };
    static constexpr int EXTRA_BYTES = 0;

    __device__ static void prologue(const Params&, uint8_t*, int, int, int) {}
    __device__ static void acquire_tile(const Params&, int, int, int) {}
    __device__ static void aux(const Params&, int, int, int) {}

    struct Epi {
        const Params& p;
        const CUtensorMap* tm_out;  // dW [C][D] fp32
        pr::Stager stager;
        int quad, lane, rank, d0;
        // operands of the normalise backward, fetched one 32-column box ahead (their L2 latency would otherwise
        // sit between every tcgen05.ld and its TMA store)
        uint4 wn[4];
        float nqn, inwn;
        __device__ Epi(const Params& prm, const pr::EpiCtx& c)
            : p(prm), tm_out(c.tmC), stager(c), quad(c.quad), lane(c.lane), rank(c.rank),
              d0(c.res * NCOL + c.half * 128) {}
        __device__ __forceinline__ int class_of(int i) const {
            return p.c_begin + i * NCOL + rank * pr::ROWS + quad * 32 + lane;
        }
        __device__ __forceinline__ void fetch_w(int i, int box) {
            const int c = class_of(i);
            const bool cvalid = c < p.C;
            const __nv_bfloat16* wrow = p.what + static_cast<int64_t>(cvalid ? c : 0) * p.D;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int d = d0 + box * 32 + g * 8;
                wn[g] = (cvalid && d < p.D) ? ldg_nc_u4(wrow + d) : make_uint4(0, 0, 0, 0);
            }
        }
        __device__ __forceinline__ void fetch_scalars(int i) {
            const int c = class_of(i);
            const bool cvalid = c < p.C;
            float qs[8];
#pragma unroll
            for (int sl = 0; sl < 8; ++sl)
                qs[sl] = (cvalid && sl < p.q_slots) ? __ldg(p.q + static_cast<int64_t>(sl) * p.C + c) : 0.f;
            nqn = -(((qs[0] + qs[1]) + (qs[2] + qs[3])) + ((qs[4] + qs[5]) + (qs[6] + qs[7])));
            inwn = cvalid ? __ldg(p.inv_nw + c) : 0.f;
        }
        __device__ void prefetch(int i) {
            fetch_scalars(i);
            fetch_w(i, 0);
        }
        __device__ void tile(int i, int i_next, uint32_t taddr) {
            const int crow0 = p.c_begin + i * NCOL + rank * pr::ROWS + quad * 32;
            const float nq = nqn, inw = inwn;
#pragma unroll 1
            for (int box = 0; box < 4; ++box) {
                const int dbox = d0 + box * 32;
                if (dbox >= p.D) {  // warp-uniform: columns past the embedding width
                    if (i_next >= 0) prefetch(i_next);
                    break;
                }
                uint4 w[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) w[g] = wn[g];
                uint32_t v[32];
                tmem_ld32(taddr + box * 32, v);
                // start the next box's operand loads before waiting on this one
                if (box < 3) fetch_w(i, box + 1);
                else if (i_next >= 0) prefetch(i_next);
                tmem_ld_wait();
                stager.acquire();
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint4 ww = w[g];
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
                stager.commit(tm_out, dbox, crow0);  // rows >= C / columns >= D are clipped by the TMA
            }
        }
        __device__ void finish() { stager.drain(); }
    };
};

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

struct BwdPlan {
    int Bp;             // scratch leading dimension (batch rounded up to 64)
    int chunk_classes;  // classes per scratch chunk (multiple of 128)
    int n_chunks;
    bool dc_rs, dw_rs, dx2;  // which kernels take the resident-operand / stacked-tile fast path
    bool dc_pair, dw_pair;   // ... on CTA pairs (cta_group::2) instead of single CTAs
    int q_slots;             // partial-sum slots of q per class
    size_t scratch_off, scratch_bytes, q_off, q_bytes, total;
};

static bool env_is(const char* name, const char* value) {
    const char* v = getenv(name);
    return v != nullptr && strcmp(v, value) == 0;
}

// Kernel selection and scratch chunking.  Environment knobs (diagnostics / A-B measurements only):
//   ARCFACE_B200_BWD_IMPL=generic   use the streaming kernels of gemm_core.cuh for every shape
//   ARCFACE_B200_BWD_IMPL=rs        resident-operand kernels on single CTAs (gemm_rs.cuh) instead of CTA pairs
//   ARCFACE_B200_BWD_CHUNK_MB=<n>   cap of the dC^T scratch in MiB
static BwdPlan plan_backward(int B, int D, int64_t C, int nsm) {
    BwdPlan pl;
    pl.Bp = ((B + 63) / 64) * 64;
    if (nsm < 1) nsm = 148;
    const bool generic = env_is("ARCFACE_B200_BWD_IMPL", "generic");
    pl.dc_rs = !generic && (D + 63) / 64 <= rs::MAX_KBLOCKS;
    pl.dw_rs = !generic && pl.Bp / 64 <= rs::MAX_KBLOCKS;
    pl.dx2 = !generic;
    const bool pairs = !env_is("ARCFACE_B200_BWD_IMPL", "rs") && nsm >= 2;
    pl.dc_pair = pl.dc_rs && pairs;
    pl.dw_pair = pl.dw_rs && pairs;
    pl.q_slots = pl.dc_pair ? 2 * ((B + BwdDCp::NCOL - 1) / BwdDCp::NCOL) : pl.dc_rs ? 2 * ((B + rs::BN - 1) / rs::BN) : 1;
    const int64_t c_round = ((C + 127) / 128) * 128;
    int64_t chunk;
    size_t cap = generic ? (size_t(64) << 20) : (size_t(2048) << 20);
    if (const char* v = getenv("ARCFACE_B200_BWD_CHUNK_MB")) {
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
        int64_t max_chunk = static_cast<int64_t>(cap / (static_cast<size_t>(pl.Bp) * 2)) / 128 * 128;
        if (max_chunk < 128) max_chunk = 128;
        const int64_t n = (c_round + max_chunk - 1) / max_chunk;
        chunk = ((c_round / 128 + n - 1) / n) * 128;
    }
    pl.chunk_classes = static_cast<int>(chunk);
    pl.n_chunks = static_cast<int>((C + chunk - 1) / chunk);
    pl.scratch_off = 0;
    pl.scratch_bytes = static_cast<size_t>(chunk) * pl.Bp * 2;
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
    *bytes = plan_backward(B, D, C_local, sm_count()).total;
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

extern "C" int32_t arcface_b200_backward(const uint16_t* xhat, const uint16_t* xhat_t, int64_t ld_t,
                                         const uint16_t* what, const float* inv_nw, const float* lse,
                                         const float* one_minus_p, const float* dphi, const int32_t* label_local,
                                         int32_t B, int32_t D, int64_t C_local, float s, float grad_scale,
                                         const float* grad_loss_dev, float* dxhat, float* dw, void* workspace,
                                         size_t workspace_bytes, void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(xhat && xhat_t && what && inv_nw && lse && one_minus_p && dphi && label_local && dxhat && dw && workspace,
               ARCFACE_B200_E_ARG, "backward: null pointer");
    if (int32_t rc = check_bwd_shape("backward", B, D, C_local)) return rc;
    AB_REQUIRE(ld_t >= B && ld_t % 8 == 0, ARCFACE_B200_E_LAYOUT, "backward: ld_t must be >= B and a multiple of 8");
    AB_REQUIRE(aligned16(dxhat) && aligned16(dw) && aligned16(workspace) && aligned16(what), ARCFACE_B200_E_LAYOUT,
               "backward: pointers must be 16-byte aligned");
    AB_REQUIRE(s > 0.f, ARCFACE_B200_E_ARG, "backward: scale s must be positive");
    const int nsm = sm_count();
    const BwdPlan pl = plan_backward(B, D, C_local, nsm);
    AB_REQUIRE(workspace_bytes >= pl.total, ARCFACE_B200_E_WORKSPACE, "backward: workspace %zu < required %zu",
               workspace_bytes, pl.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    __nv_bfloat16* dct = reinterpret_cast<__nv_bfloat16*>(ws + pl.scratch_off);
    float* q = reinterpret_cast<float*>(ws + pl.q_off);
    const int C = static_cast<int>(C_local);

    AB_CHECK_CUDA(cudaMemsetAsync(dxhat, 0, static_cast<size_t>(B) * D * sizeof(float), st));

    CUtensorMap tm_w_k, tm_x_k, tm_dct_k, tm_xt_k, tm_dct_mn, tm_w_mn, tm_dct_out, tm_dw_out, tm_dx_out;
    if (int32_t rc = make_tmap_kmajor(&tm_w_k, what, D, C_local, D, BLOCK_M)) return rc;
    if (int32_t rc = make_tmap_kmajor(&tm_x_k, xhat, D, B, D, pl.dc_rs ? rs::BN : BwdDC::BLOCK_N)) return rc;
    if (int32_t rc = make_tmap_kmajor(&tm_dct_k, dct, B, pl.chunk_classes, pl.Bp, BLOCK_M)) return rc;
    if (int32_t rc = make_tmap_kmajor(&tm_xt_k, xhat_t, B, D, ld_t, pl.dw_rs ? rs::BN : BwdDW::BLOCK_N)) return rc;
    if (int32_t rc = make_tmap_mnmajor(&tm_dct_mn, dct, B, pl.chunk_classes, pl.Bp)) return rc;
    if (int32_t rc = make_tmap_mnmajor(&tm_w_mn, what, D, C_local, D)) return rc;
    if (int32_t rc = make_tmap_store(&tm_dct_out, dct, 2, pl.Bp, pl.chunk_classes, pl.Bp)) return rc;
    if (int32_t rc = make_tmap_store(&tm_dw_out, dw, 4, D, C_local, D)) return rc;
    if (int32_t rc = make_tmap_store(&tm_dx_out, dxhat, 4, D, B, D)) return rc;

    const int n_tiles = (B + BwdDC::BLOCK_N - 1) / BwdDC::BLOCK_N;
    const int dn_tiles = (D + 255) / 256;

    for (int64_t c0 = 0; c0 < C_local; c0 += pl.chunk_classes) {
        const int cn = static_cast<int>(C_local - c0 < pl.chunk_classes ? C_local - c0 : pl.chunk_classes);
        const int c_blocks = (cn + BLOCK_M - 1) / BLOCK_M;
        // ---- dC^T (and q) for this chunk
        if (pl.dc_pair) {
            BwdDCp::Params p;
            p.core.kblocks = (D + pr::BK - 1) / pr::BK;
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
            p.core.kblocks = (D + rs::BK - 1) / rs::BK;
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
            p.B = B; p.D = D; p.C = C; p.Bp = pl.Bp;
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
            p.core.kblocks = pl.Bp / pr::BK;
            p.core.s_blocks = (cn + BwdDWp::NCOL - 1) / BwdDWp::NCOL;
            p.core.s_row0 = 0;  // the scratch is chunk-relative
            p.core.n_res = (D + BwdDWp::NCOL - 1) / BwdDWp::NCOL;
            p.core.contiguous = 0;
            p.core.prefetch_tiles = 2;
            p.C = C; p.D = D; p.c_begin = static_cast<int>(c0);
            p.q = q; p.q_slots = pl.q_slots; p.inv_nw = inv_nw;
            p.what = reinterpret_cast<const __nv_bfloat16*>(what);
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
            if (int32_t rc = launch_gemm<BwdDX2>(tm_dct_mn, tm_w_mn, tm_dx_out, p, grid, 0, st)) return rc;
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
    return ARCFACE_B200_OK;
}
