// Row-wise (HBM-bound) kernels of the ArcFace head:
//   K1  normalize_cast       fused row L2-normalise + bf16 cast        (arcface.py:47, both F.normalize calls)
//   label_margin             fp32 label cosine + additive angular margin (arcface.py:49-55 on the label column)
//   combine_partials / finalize_rows   softmax statistics -> lse, argmax, mean CE loss
//   normalize_bwd_x          backward of F.normalize for the embeddings
#include "host_util.h"
#include "ptx.cuh"

#include <math.h>

namespace ab {

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One warp per row; lane l owns the 8-float chunks l, l+32, ... (each 32 B in, 16 B out: 128-bit
// coalesced loads and stores).  NCH chunks per lane live in registers so the row is read once.
template <int NCH>
__global__ void __launch_bounds__(256) normalize_cast_kernel(const float* __restrict__ src, int64_t rows, int D,
                                                             __nv_bfloat16* __restrict__ dst,
                                                             float* __restrict__ inv_norm,
                                                             __nv_bfloat16* __restrict__ dst_t, int64_t ld_t,
                                                             const int64_t* __restrict__ index = nullptr) {
    const int lane = threadIdx.x & 31;
    const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    // index != nullptr: output row `row` is source row index[row] (class sampling: the sampled rows are normalised
    // straight out of the full weight matrix, no fp32 copy of the sub-matrix is ever made)
    const float* rp = src + (index != nullptr ? index[row] : row) * D;
    float4 v[NCH][2];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int d = (lane + 32 * i) * 8;
        if (d < D) {
            v[i][0] = ldg_stream(reinterpret_cast<const float4*>(rp + d));
            v[i][1] = ldg_stream(reinterpret_cast<const float4*>(rp + d + 4));
            ss += v[i][0].x * v[i][0].x + v[i][0].y * v[i][0].y + v[i][0].z * v[i][0].z + v[i][0].w * v[i][0].w;
            ss += v[i][1].x * v[i][1].x + v[i][1].y * v[i][1].y + v[i][1].z * v[i][1].z + v[i][1].w * v[i][1].w;
        }
    }
    ss = warp_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    if (lane == 0) inv_norm[row] = inv;
    __nv_bfloat16* op = dst + row * D;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int d = (lane + 32 * i) * 8;
        if (d < D) {
            uint4 o;
            o.x = pack_bf16x2(v[i][0].x * inv, v[i][0].y * inv);
            o.y = pack_bf16x2(v[i][0].z * inv, v[i][0].w * inv);
            o.z = pack_bf16x2(v[i][1].x * inv, v[i][1].y * inv);
            o.w = pack_bf16x2(v[i][1].z * inv, v[i][1].w * inv);
            *reinterpret_cast<uint4*>(op + d) = o;
            if (dst_t != nullptr) {
                const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    __nv_bfloat16_raw lo, hi;
                    lo.x = static_cast<unsigned short>(w[j] & 0xffffu);
                    hi.x = static_cast<unsigned short>(w[j] >> 16);
                    dst_t[static_cast<int64_t>(d + 2 * j) * ld_t + row] = __nv_bfloat16(lo);
                    dst_t[static_cast<int64_t>(d + 2 * j + 1) * ld_t + row] = __nv_bfloat16(hi);
                }
            }
        }
    }
}

// Any D (multiple of 8): two passes, the second served by L1/L2.
__global__ void __launch_bounds__(256) normalize_cast_generic_kernel(const float* __restrict__ src, int64_t rows, int D,
                                                                     __nv_bfloat16* __restrict__ dst,
                                                                     float* __restrict__ inv_norm,
                                                                     __nv_bfloat16* __restrict__ dst_t, int64_t ld_t,
                                                                     const int64_t* __restrict__ index = nullptr) {
    const int lane = threadIdx.x & 31;
    const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* rp = src + (index != nullptr ? index[row] : row) * D;
    float ss = 0.f;
    for (int d = lane * 4; d < D; d += 128) {
        const float4 a = *reinterpret_cast<const float4*>(rp + d);
        ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
    ss = warp_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    if (lane == 0) inv_norm[row] = inv;
    __nv_bfloat16* op = dst + row * D;
    for (int d = lane * 4; d < D; d += 128) {
        const float4 a = *reinterpret_cast<const float4*>(rp + d);
        uint2 o;
        o.x = pack_bf16x2(a.x * inv, a.y * inv);
        o.y = pack_bf16x2(a.z * inv, a.w * inv);
        *reinterpret_cast<uint2*>(op + d) = o;
        if (dst_t != nullptr) {
            const float f[4] = {a.x * inv, a.y * inv, a.z * inv, a.w * inv};
#pragma unroll
            for (int j = 0; j < 4; ++j) dst_t[static_cast<int64_t>(d + j) * ld_t + row] = __float2bfloat16_rn(f[j]);
        }
    }
}

// K1 for the high-precision mode (precision = 'bf16x3'): every normalised value v is kept as a pair of bf16,
// hi = bf16(v) and lo = bf16(v - hi), and the row is written THREE times along the contraction dimension so that the
// unchanged bf16 tensor-core GEMM over 3 D columns computes  x_hi.w_hi + x_hi.w_lo + x_lo.w_hi  (everything but the
// 2^-18 lo.lo term) in its fp32 accumulator:
//   order 0 (embeddings)   : [hi | hi | lo]
//   order 1 (class weights): [hi | lo | hi]
// Optional transpose of the hi part ([D][ld_t], the dW GEMM's operand).  One warp per row, two passes (the second
// served by L1/L2); a parity mode, not a throughput path.
__global__ void __launch_bounds__(256) normalize_cast3_kernel(const float* __restrict__ src, int64_t rows, int D,
                                                              int order, __nv_bfloat16* __restrict__ dst,
                                                              float* __restrict__ inv_norm,
                                                              __nv_bfloat16* __restrict__ dst_t, int64_t ld_t) {
    const int lane = threadIdx.x & 31;
    const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* rp = src + row * D;
    float ss = 0.f;
    for (int d = lane * 4; d < D; d += 128) {
        const float4 a = *reinterpret_cast<const float4*>(rp + d);
        ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
    ss = warp_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    if (lane == 0) inv_norm[row] = inv;
    __nv_bfloat16* op = dst + row * 3 * static_cast<int64_t>(D);
    const int off_hi2 = order == 0 ? D : 2 * D;   // second copy of hi
    const int off_lo = order == 0 ? 2 * D : D;
    for (int d = lane * 4; d < D; d += 128) {
        const float4 a = *reinterpret_cast<const float4*>(rp + d);
        const float f[4] = {a.x * inv, a.y * inv, a.z * inv, a.w * inv};
        float h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            h[j] = __bfloat162float(__float2bfloat16_rn(f[j]));
            l[j] = f[j] - h[j];
        }
        uint2 hi, lo;
        hi.x = pack_bf16x2(h[0], h[1]);
        hi.y = pack_bf16x2(h[2], h[3]);
        lo.x = pack_bf16x2(l[0], l[1]);
        lo.y = pack_bf16x2(l[2], l[3]);
        *reinterpret_cast<uint2*>(op + d) = hi;
        *reinterpret_cast<uint2*>(op + off_hi2 + d) = hi;
        *reinterpret_cast<uint2*>(op + off_lo + d) = lo;
        if (dst_t != nullptr) {
            // transposed operand of the dW GEMM, three blocks of ld_t columns each: [hi^T | lo^T | hi^T] -- against
            // dC^T laid out [hi | hi | lo] the contraction over 3 ld_t columns is hi.hi + hi.lo + lo.hi again
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                __nv_bfloat16* tp = dst_t + static_cast<int64_t>(d + j) * (3 * ld_t) + row;
                const __nv_bfloat16 hb = __float2bfloat16_rn(h[j]);
                tp[0] = hb;
                tp[ld_t] = __float2bfloat16_rn(l[j]);
                tp[2 * ld_t] = hb;
            }
        }
    }
}

// bf16x3 mode: the dW epilogue subtracts q[c] * what_hi[c] only; this adds the lo part of the projection,
//   dW[c, :] -= inv_nw[c] * q[c] * what_lo[c, :],   what3 rows laid out [hi | lo | hi] (3 D wide), q = sum of q_slots slots.
__global__ void __launch_bounds__(256)
dw_lo_correction_kernel(float* __restrict__ dw, const __nv_bfloat16* __restrict__ what3, const float* __restrict__ q,
                        int q_slots, const float* __restrict__ inv_nw, int64_t C, int D) {
    const int lane = threadIdx.x & 31;
    const int64_t c = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    float qc = 0.f;
    for (int sl = 0; sl < q_slots; ++sl) qc += q[static_cast<int64_t>(sl) * C + c];
    const float f = -qc * inv_nw[c];
    const __nv_bfloat16* lo = what3 + c * 3 * static_cast<int64_t>(D) + D;
    float* row = dw + c * static_cast<int64_t>(D);
    for (int d = lane * 4; d < D; d += 128) {
        const uint2 w = *reinterpret_cast<const uint2*>(lo + d);
        float4 v = *reinterpret_cast<float4*>(row + d);
        v.x = fmaf(f, __uint_as_float(w.x << 16), v.x);
        v.y = fmaf(f, __uint_as_float(w.x & 0xffff0000u), v.y);
        v.z = fmaf(f, __uint_as_float(w.y << 16), v.z);
        v.w = fmaf(f, __uint_as_float(w.y & 0xffff0000u), v.w);
        *reinterpret_cast<float4*>(row + d) = v;
    }
}



// The step before the head in the two-stream model (multimodal_classifier.py:50-56):
//   emb = cat(F.normalize(img_emb), F.normalize(title_emb), dim = 1)
// as ONE pass: one warp per row, both halves normalised and written side by side (the reference runs two norm
// reductions, two divisions and a concat: five launches, each a B x D round trip).  inv1 / inv2 keep 1 / max(||.||, eps)
// for the backward below.
__global__ void __launch_bounds__(256)
two_stream_concat_kernel(const float* __restrict__ a, const float* __restrict__ b, int B, int D1, int D2,
                         float* __restrict__ out, float* __restrict__ inv1, float* __restrict__ inv2) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= B) return;
    const float* ap = a + static_cast<int64_t>(r) * D1;
    const float* bp = b + static_cast<int64_t>(r) * D2;
    float sa = 0.f, sb = 0.f;
    for (int d = lane * 4; d < D1; d += 128) {
        const float4 v = *reinterpret_cast<const float4*>(ap + d);
        sa += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (int d = lane * 4; d < D2; d += 128) {
        const float4 v = *reinterpret_cast<const float4*>(bp + d);
        sb += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    sa = warp_sum(sa);
    sb = warp_sum(sb);
    const float ia = 1.0f / fmaxf(sqrtf(sa), 1e-12f), ib = 1.0f / fmaxf(sqrtf(sb), 1e-12f);
    if (lane == 0) { inv1[r] = ia; inv2[r] = ib; }
    float* op = out + static_cast<int64_t>(r) * (D1 + D2);
    for (int d = lane * 4; d < D1; d += 128) {
        const float4 v = *reinterpret_cast<const float4*>(ap + d);
        *reinterpret_cast<float4*>(op + d) = make_float4(v.x * ia, v.y * ia, v.z * ia, v.w * ia);
    }
    for (int d = lane * 4; d < D2; d += 128) {
        const float4 v = *reinterpret_cast<const float4*>(bp + d);
        *reinterpret_cast<float4*>(op + D1 + d) = make_float4(v.x * ib, v.y * ib, v.z * ib, v.w * ib);
    }
}

// Backward of the above: with e = a normalised half of `emb` and g the matching half of the incoming gradient,
// d(input) = (g - e (e . g)) * inv  (F.normalize's backward), for both halves in one pass.
__global__ void __launch_bounds__(256)
two_stream_concat_bwd_kernel(const float* __restrict__ emb, const float* __restrict__ inv1, const float* __restrict__ inv2,
                             const float* __restrict__ grad, int B, int D1, int D2, float* __restrict__ da,
                             float* __restrict__ db) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= B) return;
    const int D = D1 + D2;
    const float* ep = emb + static_cast<int64_t>(r) * D;
    const float* gp = grad + static_cast<int64_t>(r) * D;
    float ra = 0.f, rb = 0.f;
    for (int d = lane * 4; d < D1; d += 128) {
        const float4 e = *reinterpret_cast<const float4*>(ep + d), g = *reinterpret_cast<const float4*>(gp + d);
        ra += e.x * g.x + e.y * g.y + e.z * g.z + e.w * g.w;
    }
    for (int d = lane * 4; d < D2; d += 128) {
        const float4 e = *reinterpret_cast<const float4*>(ep + D1 + d), g = *reinterpret_cast<const float4*>(gp + D1 + d);
        rb += e.x * g.x + e.y * g.y + e.z * g.z + e.w * g.w;
    }
    ra = warp_sum(ra);
    rb = warp_sum(rb);
    const float ia = inv1[r], ib = inv2[r];
    float* ap = da + static_cast<int64_t>(r) * D1;
    float* bp = db + static_cast<int64_t>(r) * D2;
    for (int d = lane * 4; d < D1; d += 128) {
        const float4 e = *reinterpret_cast<const float4*>(ep + d), g = *reinterpret_cast<const float4*>(gp + d);
        *reinterpret_cast<float4*>(ap + d) = make_float4((g.x - e.x * ra) * ia, (g.y - e.y * ra) * ia,
                                                         (g.z - e.z * ra) * ia, (g.w - e.w * ra) * ia);
    }
    for (int d = lane * 4; d < D2; d += 128) {
        const float4 e = *reinterpret_cast<const float4*>(ep + D1 + d), g = *reinterpret_cast<const float4*>(gp + D1 + d);
        *reinterpret_cast<float4*>(bp + d) = make_float4((g.x - e.x * rb) * ib, (g.y - e.y * rb) * ib,
                                                         (g.z - e.z * rb) * ib, (g.w - e.w * rb) * ib);
    }
}

__global__ void __launch_bounds__(256)
label_margin_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ inv_nx,
                    const float* __restrict__ inv_nw, const int64_t* __restrict__ label, int B, int D, int64_t C_local,
                    int64_t class_offset, int64_t C_total, float s, float cos_m, float sin_m, float th, float mm,
                    int easy, float* __restrict__ t_label, float* __restrict__ z_label, float* __restrict__ dphi,
                    int* __restrict__ label_local, int* __restrict__ bad_flag) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const int64_t y = label[b];
    // plain store (every writer stores the same value): `bad_flag` may be mapped pinned HOST memory, so that the
    // flag reaches the host without a copy node in the step (engine.LabelGuard)
    if ((y < 0 || y >= C_total) && lane == 0 && bad_flag != nullptr) *reinterpret_cast<volatile int*>(bad_flag) = 1;
    const int64_t loc = y - class_offset;
    if (loc < 0 || loc >= C_local) {
        if (lane == 0) {
            t_label[b] = 0.f;
            z_label[b] = 0.f;
            dphi[b] = 0.f;
            label_local[b] = -1;
        }
        return;
    }
    const float* xp = x + static_cast<int64_t>(b) * D;
    const float* wp = w + loc * D;
    float dot = 0.f, ww = 0.f;
    for (int d = lane * 4; d < D; d += 128) {
        const float4 a = *reinterpret_cast<const float4*>(xp + d);
        const float4 c = *reinterpret_cast<const float4*>(wp + d);
        dot += a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w;
        ww += c.x * c.x + c.y * c.y + c.z * c.z + c.w * c.w;
    }
    dot = warp_sum(dot);
    ww = warp_sum(ww);
    if (lane == 0) {
        const float inw = (inv_nw != nullptr) ? inv_nw[loc] : 1.0f / fmaxf(sqrtf(ww), 1e-12f);
        const float t = dot * inv_nx[b] * inw;
        // reference: sqrt(1 - cos^2) unclamped (arcface.py:49); clamped here so |t| = 1 + ulp cannot NaN
        const float sine = sqrtf(fmaxf(0.f, 1.f - t * t));
        const float phi = t * cos_m - sine * sin_m;
        const bool take_phi = easy ? (t > 0.f) : ((t - th) > 0.f);
        const float u = take_phi ? phi : (easy ? t : (t - mm));
        t_label[b] = t;
        z_label[b] = u * s;
        dphi[b] = take_phi ? (cos_m + t * sin_m / fmaxf(sine, 1e-6f)) : 1.f;
        label_local[b] = static_cast<int>(loc);
    }
}

// One warp per batch row: lane l merges parts l, l + 32, ... (independent loads), then the 32 partial
// (max, sum-exp, argmax) triples are merged by shuffles.  Ties on the maximum keep the LOWEST class index
// (torch.argmax), whatever order the parts cover the classes in.
__global__ void __launch_bounds__(256)
combine_partials_kernel(const float* __restrict__ pm, const float* __restrict__ ps, const int* __restrict__ pa,
                        int n_parts, int B, int64_t class_offset, float* __restrict__ row_max,
                        float* __restrict__ row_sum, int64_t* __restrict__ row_arg) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    float M = -INFINITY, S = 0.f;
    int A = 0x7fffffff;
    for (int p = lane; p < n_parts; p += 32) {
        const float m = pm[static_cast<int64_t>(p) * B + b];
        const float s = ps[static_cast<int64_t>(p) * B + b];
        const int a = pa[static_cast<int64_t>(p) * B + b];
        if (m == -INFINITY) continue;  // empty part
        if (m > M) {
            S = S * expf(M - m) + s;  // expf(-inf) = 0 on the first part
            M = m;
            A = a;
        } else {
            S += s * expf(m - M);
            if (m == M && a < A) A = a;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, M, o);
        const float s2 = __shfl_xor_sync(0xffffffffu, S, o);
        const int a2 = __shfl_xor_sync(0xffffffffu, A, o);
        if (m2 > M) {
            S = (M > -INFINITY ? S * expf(M - m2) : 0.f) + s2;
            M = m2;
            A = a2;
        } else if (m2 > -INFINITY) {
            S += s2 * expf(m2 - M);
            if (m2 == M && a2 < A) A = a2;
        }
    }
    if (lane == 0) {
        row_max[b] = M;
        row_sum[b] = S;
        row_arg[b] = static_cast<int64_t>(M > -INFINITY ? A : 0) + class_offset;
    }
}

// Merges the per-rank statistics of the NON-label columns with the fp32 label logit.  Keeping the label
// out of the streamed sum makes 1 - p_label = S_rest / (S_rest + e_label) free of cancellation, which
// matters once the head is trained (p_label -> 1) -- the reference's fp32 softmax - one_hot loses those
// digits.
// One CTA (the mean loss is a fixed-order tree sum: deterministic), up to 1024 threads so that a 512- or 1024-row batch
// is one row per thread; the per-rank values of a row are loaded eight ranks at a time before the merge runs over
// them (the loads are independent, the merge is not): at 8 ranks this kernel sits on the critical path between the
// statistics exchange and the backward.
__global__ void __launch_bounds__(1024)
finalize_rows_kernel(const float* __restrict__ rm, const float* __restrict__ rs, const int64_t* __restrict__ ra,
                     const float* __restrict__ rz, const int64_t* __restrict__ label, int n_ranks, int B,
                     int64_t fstride, int64_t astride, float* __restrict__ lse, int64_t* __restrict__ argmax, float* __restrict__ z_out,
                     float* __restrict__ one_minus_p, float* __restrict__ loss) {
    __shared__ float red[1024];
    float acc = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float Mx = -INFINITY, Sx = 0.f, Z = 0.f;
        int64_t Ax = 0;
        for (int r0 = 0; r0 < n_ranks; r0 += 8) {
            float mv[8], sv[8], zv[8];
            int64_t av[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int r = r0 + k;
                const bool ok = r < n_ranks;
                mv[k] = ok ? rm[r * fstride + b] : -INFINITY;
                sv[k] = ok ? rs[r * fstride + b] : 0.f;
                zv[k] = ok ? rz[r * fstride + b] : 0.f;
                av[k] = ok ? ra[r * astride + b] : 0;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (r0 + k >= n_ranks) break;
                const float m = mv[k];
                const float s = sv[k];
                if (m > Mx) {  // strict: the lower rank (lower class range) keeps ties
                    Sx = Sx * expf(Mx - m) + s;
                    Mx = m;
                    Ax = av[k];
                } else if (m > -INFINITY) {
                    Sx += s * expf(m - Mx);
                }
                Z += zv[k];
            }
        }
        const int64_t y = label[b];
        float l, omp, ce;
        if (Z >= Mx) {  // label logit is the row maximum
            const float ex = (Mx > -INFINITY) ? Sx * expf(Mx - Z) : 0.f;
            ce = log1pf(ex);
            l = Z + ce;
            omp = ex / (1.f + ex);
        } else {
            const float ey = expf(Z - Mx);
            const float S = Sx + ey;
            l = Mx + logf(S);
            ce = (Mx - Z) + logf(S);
            omp = Sx / S;
        }
        lse[b] = l;
        argmax[b] = (Z > Mx || (Z == Mx && y < Ax)) ? y : Ax;
        z_out[b] = Z;
        one_minus_p[b] = omp;
        acc += ce;
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = static_cast<int>(blockDim.x) >> 1; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = red[0] / static_cast<float>(B);
}

__global__ void __launch_bounds__(256)
normalize_bwd_x_kernel(const float* __restrict__ x, const float* __restrict__ inv_nx, const float* __restrict__ dxhat,
                       int B, int D, float* __restrict__ dx) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const float inv = inv_nx[b];
    const float* xp = x + static_cast<int64_t>(b) * D;
    const float* gp = dxhat + static_cast<int64_t>(b) * D;
    float r = 0.f;
    for (int d = lane * 4; d < D; d += 128) {
        const float4 a = *reinterpret_cast<const float4*>(xp + d);
        const float4 g = *reinterpret_cast<const float4*>(gp + d);
        r += (a.x * g.x + a.y * g.y + a.z * g.z + a.w * g.w);
    }
    r = warp_sum(r) * inv;  // xhat . dxhat
    float* op = dx + static_cast<int64_t>(b) * D;
    for (int d = lane * 4; d < D; d += 128) {
        const float4 a = *reinterpret_cast<const float4*>(xp + d);
        const float4 g = *reinterpret_cast<const float4*>(gp + d);
        float4 o;
        o.x = (g.x - r * (a.x * inv)) * inv;
        o.y = (g.y - r * (a.y * inv)) * inv;
        o.z = (g.z - r * (a.z * inv)) * inv;
        o.w = (g.w - r * (a.w * inv)) * inv;
        *reinterpret_cast<float4*>(op + d) = o;
    }
}

// a[i] *= *g, b[i] *= *g; every thread reads the scalar and the whole grid leaves at once when it is exactly 1
// (the common case: loss.backward() on the head's own loss), so the call then costs one empty launch.
__global__ void __launch_bounds__(256)
scale_grads_kernel(float* __restrict__ a, int64_t na4, float* __restrict__ b, int64_t nb4, const float* __restrict__ g) {
    const float f = *g;
    if (f == 1.0f) return;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < na4 + nb4; i += stride) {
        float4* p = i < na4 ? reinterpret_cast<float4*>(a) + i : reinterpret_cast<float4*>(b) + (i - na4);
        float4 v = *p;
        v.x *= f; v.y *= f; v.z *= f; v.w *= f;
        *p = v;
    }
}

// dst = src * (*g) (always written: the caller hands `dst` to autograd and keeps `src` as a static buffer), and
// b *= (*g) in place unless the factor is exactly 1.  One launch for what used to be scale_grads + a device copy.
__global__ void __launch_bounds__(256)
scale_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n4, float* __restrict__ b, int64_t nb4,
                  const float* __restrict__ g) {
    const float f = *g;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t total = f == 1.0f ? n4 : n4 + nb4;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        if (i < n4) {
            float4 v = reinterpret_cast<const float4*>(src)[i];
            v.x *= f; v.y *= f; v.z *= f; v.w *= f;
            reinterpret_cast<float4*>(dst)[i] = v;
        } else {
            float4* p = reinterpret_cast<float4*>(b) + (i - n4);
            float4 v = *p;
            v.x *= f; v.y *= f; v.z *= f; v.w *= f;
            *p = v;
        }
    }
}

// The step's inputs into the packed static buffer of a captured graph: dst = [x (b x D fp32) | labels (b int64)].
__global__ void __launch_bounds__(256)
pack_xy_kernel(const float* __restrict__ x, const int64_t* __restrict__ y, int64_t nx4, int64_t ny, unsigned char* __restrict__ dst) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nx4 + ny; i += stride) {
        if (i < nx4) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(x)[i];
        else reinterpret_cast<int64_t*>(dst + nx4 * 16)[i - nx4] = y[i - nx4];
    }
}

// Fused head optimiser: one AdamW step on the class-weight rows AND the next step's K1
// in the same pass.  Reference: `AdamW(model.classifier.parameters(), lr=1e-2)` + optimizer.step()
// (nlp_classifier_train.py:94-97, 131-133) followed, one forward later, by F.normalize(self.weight) (arcface.py:47).
// Update rule = torch.optim.AdamW (decoupled weight decay):
//   w *= 1 - lr * wd;  m += (1 - b1) (g - m);  v = b2 v + (1 - b2) g^2;  w -= step_size * m / (sqrt(v) / sqrt(bc2) + eps)
// with step_size = lr / (1 - b1^t), bc2 = 1 - b2^t computed by the caller.  One warp per row: the first sweep
// updates w / m / v and accumulates ||w_new||^2, the second re-reads the freshly written row (L1 / L2) and writes
// bf16 what = w_new / max(||w_new||, 1e-12) and inv_nw -- the only extra HBM traffic is the 2-byte what store.
__global__ void __launch_bounds__(256)
adamw_normalize_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                       int64_t rows, int D, float decay, float one_minus_b1, float b2, float one_minus_b2,
                       float step_size, float inv_bc2_sqrt, float eps, __nv_bfloat16* __restrict__ what,
                       float* __restrict__ inv_nw) {
    const int lane = threadIdx.x & 31;
    const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    float* wp = w + row * D;
    const float* gp = g + row * D;
    float* mp = m + row * D;
    float* vp = v + row * D;
    float ss = 0.f;
    for (int d = lane * 4; d < D; d += 128) {
        float4 a = *reinterpret_cast<float4*>(wp + d);
        const float4 gg = *reinterpret_cast<const float4*>(gp + d);
        float4 mm = *reinterpret_cast<float4*>(mp + d);
        float4 vv = *reinterpret_cast<float4*>(vp + d);
        float* af = reinterpret_cast<float*>(&a);
        const float* gf = reinterpret_cast<const float*>(&gg);
        float* mf = reinterpret_cast<float*>(&mm);
        float* vf = reinterpret_cast<float*>(&vv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gj = gf[j];
            float p = af[j] * decay;
            const float mj = mf[j] + one_minus_b1 * (gj - mf[j]);
            const float vj = vf[j] * b2 + one_minus_b2 * (gj * gj);
            const float denom = sqrtf(vj) * inv_bc2_sqrt + eps;
            p = p - step_size * (mj / denom);
            af[j] = p; mf[j] = mj; vf[j] = vj;
            ss += p * p;
        }
        *reinterpret_cast<float4*>(wp + d) = a;
        *reinterpret_cast<float4*>(mp + d) = mm;
        *reinterpret_cast<float4*>(vp + d) = vv;
    }
    if (what == nullptr) return;
    ss = warp_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    if (lane == 0) inv_nw[row] = inv;
    __nv_bfloat16* op = what + row * D;
    for (int d = lane * 4; d < D; d += 128) {
        const float4 a = *reinterpret_cast<const float4*>(wp + d);  // this lane wrote these four values above
        uint2 o;
        o.x = pack_bf16x2(a.x * inv, a.y * inv);
        o.y = pack_bf16x2(a.z * inv, a.w * inv);
        *reinterpret_cast<uint2*>(op + d) = o;
    }
}

}  // namespace ab

using namespace ab;

extern "C" int32_t arcface_b200_adamw_normalize(float* w, const float* grad, float* exp_avg, float* exp_avg_sq,
                                                int64_t rows, int32_t D, double lr, double beta1, double beta2,
                                                double eps, double weight_decay, int64_t step, uint16_t* what,
                                                float* inv_nw, void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(w && grad && exp_avg && exp_avg_sq, ARCFACE_B200_E_ARG, "adamw_normalize: null pointer");
    AB_REQUIRE((what == nullptr) == (inv_nw == nullptr), ARCFACE_B200_E_ARG,
               "adamw_normalize: what and inv_nw must both be given or both be null");
    AB_REQUIRE(rows >= 0 && D >= 8 && D % 8 == 0, ARCFACE_B200_E_SHAPE, "adamw_normalize: D=%d must be a positive multiple of 8", D);
    AB_REQUIRE(step >= 1 && lr >= 0. && beta1 >= 0. && beta1 < 1. && beta2 >= 0. && beta2 < 1. && eps >= 0.,
               ARCFACE_B200_E_ARG, "adamw_normalize: invalid hyper-parameter");
    AB_REQUIRE(aligned16(w) && aligned16(grad) && aligned16(exp_avg) && aligned16(exp_avg_sq) && aligned16(what),
               ARCFACE_B200_E_LAYOUT, "adamw_normalize: pointers must be 16-byte aligned");
    if (rows == 0) return ARCFACE_B200_OK;
    // scalars are formed in double and rounded once, like the Python floats torch.optim.AdamW passes to its kernels
    const double bc1 = 1.0 - pow(beta1, static_cast<double>(step));
    const double bc2 = 1.0 - pow(beta2, static_cast<double>(step));
    const int wpb = 8;
    const int64_t nblk = (rows + wpb - 1) / wpb;
    AB_REQUIRE(nblk < (1ll << 31), ARCFACE_B200_E_SHAPE, "adamw_normalize: too many rows");
    adamw_normalize_kernel<<<static_cast<unsigned>(nblk), wpb * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        w, grad, exp_avg, exp_avg_sq, rows, D, static_cast<float>(1.0 - lr * weight_decay),
        static_cast<float>(1.0 - beta1), static_cast<float>(beta2), static_cast<float>(1.0 - beta2),
        static_cast<float>(lr / bc1), static_cast<float>(1.0 / sqrt(bc2)), static_cast<float>(eps),
        reinterpret_cast<__nv_bfloat16*>(what), inv_nw);
    AB_CHECK_CUDA(cudaGetLastError());
    return ARCFACE_B200_OK;
}

extern "C" int32_t arcface_b200_scale_grads(float* a, int64_t na, float* b, int64_t nb, const float* scale_dev,
                                            void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(scale_dev && (a || na == 0) && (b || nb == 0), ARCFACE_B200_E_ARG, "scale_grads: null pointer");
    AB_REQUIRE(na >= 0 && nb >= 0 && na % 4 == 0 && nb % 4 == 0, ARCFACE_B200_E_SHAPE,
               "scale_grads: element counts must be non-negative multiples of 4");
    AB_REQUIRE(aligned16(a) && aligned16(b), ARCFACE_B200_E_LAYOUT, "scale_grads: pointers must be 16-byte aligned");
    if (na + nb == 0) return ARCFACE_B200_OK;
    const int64_t n4 = (na + nb) / 4;
    const int64_t want = (n4 + 255) / 256;
    const int grid = static_cast<int>(want < 8 * 148 ? want : 8 * 148);
    scale_grads_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, na / 4, b, nb / 4, scale_dev);
    AB_CHECK_CUDA(cudaGetLastError());
    return ARCFACE_B200_OK;
}

extern "C" int32_t arcface_b200_scale_copy(const float* src, float* dst, int64_t n, float* b, int64_t nb,
                                           const float* scale_dev, void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(scale_dev && (n == 0 || (src && dst)) && (b || nb == 0), ARCFACE_B200_E_ARG, "scale_copy: null pointer");
    AB_REQUIRE(n >= 0 && nb >= 0 && n % 4 == 0 && nb % 4 == 0, ARCFACE_B200_E_SHAPE,
               "scale_copy: element counts must be non-negative multiples of 4");
    AB_REQUIRE(aligned16(src) && aligned16(dst) && aligned16(b), ARCFACE_B200_E_LAYOUT,
               "scale_copy: pointers must be 16-byte aligned");
    if (n + nb == 0) return ARCFACE_B200_OK;
    const int64_t n4 = (n + nb) / 4;
    const int64_t want = (n4 + 255) / 256;
    const int grid = static_cast<int>(want < 8 * 148 ? want : 8 * 148);
    scale_copy_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, n / 4, b, nb / 4, scale_dev);
    AB_CHECK_CUDA(cudaGetLastError());
    return ARCFACE_B200_OK;
}

extern "C" int32_t arcface_b200_pack_xy(const float* x, const int64_t* y, int32_t b, int32_t D, void* dst, void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(x && y && dst, ARCFACE_B200_E_ARG, "pack_xy: null pointer");
    AB_REQUIRE(b >= 0 && D >= 4 && D % 4 == 0, ARCFACE_B200_E_SHAPE, "pack_xy: bad shape");
    AB_REQUIRE(aligned16(x) && aligned16(dst) && (reinterpret_cast<uintptr_t>(y) & 7u) == 0, ARCFACE_B200_E_LAYOUT,
               "pack_xy: x / dst must be 16-byte aligned, y 8-byte aligned");
    if (b == 0) return ARCFACE_B200_OK;
    const int64_t nx4 = static_cast<int64_t>(b) * D / 4;
    const int64_t want = (nx4 + b + 255) / 256;
    const int grid = static_cast<int>(want < 2 * 148 ? want : 2 * 148);
    pack_xy_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, nx4, b, static_cast<unsigned char*>(dst));
    AB_CHECK_CUDA(cudaGetLastError());
    return ARCFACE_B200_OK;
}

extern "C" int32_t arcface_b200_normalize_cast(const float* src, int64_t rows, int32_t D, uint16_t* dst,
                                               float* inv_norm, uint16_t* dst_t, int64_t ld_t, void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(src && dst && inv_norm, ARCFACE_B200_E_ARG, "normalize_cast: null pointer");
    AB_REQUIRE(rows >= 0 && D >= 8 && D % 8 == 0, ARCFACE_B200_E_SHAPE, "normalize_cast: D=%d must be a positive multiple of 8", D);
    AB_REQUIRE(aligned16(src) && aligned16(dst), ARCFACE_B200_E_LAYOUT, "normalize_cast: pointers must be 16-byte aligned");
    AB_REQUIRE(dst_t == nullptr || ld_t >= rows, ARCFACE_B200_E_LAYOUT, "normalize_cast: ld_t < rows");
    if (rows == 0) return ARCFACE_B200_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int wpb = 8;
    const int64_t nblk = (rows + wpb - 1) / wpb;
    AB_REQUIRE(nblk < (1ll << 31), ARCFACE_B200_E_SHAPE, "normalize_cast: too many rows");
    dim3 grid(static_cast<unsigned>(nblk)), block(wpb * 32);
    __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst);
    __nv_bfloat16* dt = reinterpret_cast<__nv_bfloat16*>(dst_t);
    const int nch = (D + 255) / 256;
    if (nch <= 1) normalize_cast_kernel<1><<<grid, block, 0, st>>>(src, rows, D, d, inv_norm, dt, ld_t);
    else if (nch <= 2) normalize_cast_kernel<2><<<grid, block, 0, st>>>(src, rows, D, d, inv_norm, dt, ld_t);
    else if (nch <= 4) normalize_cast_kernel<4><<<grid, block, 0, st>>>(src, rows, D, d, inv_norm, dt, ld_t);
    else if (nch <= 8) normalize_cast_kernel<8><<<grid, block, 0, st>>>(src, rows, D, d, inv_norm, dt, ld_t);
    else if (nch <= 12) normalize_cast_kernel<12><<<grid, block, 0, st>>>(src, rows, D, d, inv_norm, dt, ld_t);
    else normalize_cast_generic_kernel<<<grid, block, 0, st>>>(src, rows, D, d, inv_norm, dt, ld_t);
    AB_CHECK_CUDA(cudaGetLastError());
    return ARCFACE_B200_OK;
}

// Class sampling (PartialFC-style): K1 over the sampled rows of the full weight matrix.
extern "C" int32_t arcface_b200_normalize_cast_gather(const float* src, int64_t src_rows, const int64_t* index,
                                                      int64_t rows, int32_t D, uint16_t* dst, float* inv_norm,
                                                      void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(src && index && dst && inv_norm, ARCFACE_B200_E_ARG, "normalize_cast_gather: null pointer");
    AB_REQUIRE(rows >= 0 && src_rows >= 0 && D >= 8 && D % 8 == 0, ARCFACE_B200_E_SHAPE,
               "normalize_cast_gather: D=%d must be a positive multiple of 8", D);
    AB_REQUIRE(aligned16(src) && aligned16(dst) && (reinterpret_cast<uintptr_t>(index) & 7u) == 0, ARCFACE_B200_E_LAYOUT,
               "normalize_cast_gather: src / dst must be 16-byte aligned, index 8-byte aligned");
    if (rows == 0) return ARCFACE_B200_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int wpb = 8;
    const int64_t nblk = (rows + wpb - 1) / wpb;
    AB_REQUIRE(nblk < (1ll << 31), ARCFACE_B200_E_SHAPE, "normalize_cast_gather: too many rows");
    dim3 grid(static_cast<unsigned>(nblk)), block(wpb * 32);
    __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst);
    const int nch = (D + 255) / 256;
    if (nch <= 1) normalize_cast_kernel<1><<<grid, block, 0, st>>>(src, rows, D, d, inv_norm, nullptr, 0, index);
    else if (nch <= 2) normalize_cast_kernel<2><<<grid, block, 0, st>>>(src, rows, D, d, inv_norm, nullptr, 0, index);
    else if (nch <= 4) normalize_cast_kernel<4><<<grid, block, 0, st>>>(src, rows, D, d, inv_norm, nullptr, 0, index);
    else if (nch <= 8) normalize_cast_kernel<8><<<grid, block, 0, st>>>(src, rows, D, d, inv_norm, nullptr, 0, index);
    else if (nch <= 12) normalize_cast_kernel<12><<<grid, block, 0, st>>>(src, rows, D, d, inv_norm, nullptr, 0, index);
    else normalize_cast_generic_kernel<<<grid, block, 0, st>>>(src, rows, D, d, inv_norm, nullptr, 0, index);
    AB_CHECK_CUDA(cudaGetLastError());
    return ARCFACE_B200_OK;
}

// dst[index[r]] = src[r] for r < rows (fp32 rows of D floats, D % 4 == 0): the gradient of the sampled class rows back
// into the full-size gradient.  One warp per row, 128-bit copies.
__global__ void __launch_bounds__(256) scatter_rows_kernel(const float4* __restrict__ src, const int64_t* __restrict__ index,
                                                           int64_t rows, int d4, float4* __restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float4* s = src + row * d4;
    float4* d = dst + index[row] * d4;
    for (int i = lane; i < d4; i += 32) d[i] = ldg_stream(s + i);
}

extern "C" int32_t arcface_b200_scatter_rows(const float* src, const int64_t* index, int64_t rows, int32_t D, float* dst,
                                             int64_t dst_rows, void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(src && index && dst, ARCFACE_B200_E_ARG, "scatter_rows: null pointer");
    AB_REQUIRE(rows >= 0 && rows <= dst_rows && D >= 4 && D % 4 == 0, ARCFACE_B200_E_SHAPE,
               "scatter_rows: %lld rows into %lld, D=%d (multiple of 4)", static_cast<long long>(rows),
               static_cast<long long>(dst_rows), D);
    AB_REQUIRE(aligned16(src) && aligned16(dst) && (reinterpret_cast<uintptr_t>(index) & 7u) == 0, ARCFACE_B200_E_LAYOUT,
               "scatter_rows: src / dst must be 16-byte aligned, index 8-byte aligned");
    if (rows == 0) return ARCFACE_B200_OK;
    const int64_t nblk = (rows + 7) / 8;
    AB_REQUIRE(nblk < (1ll << 31), ARCFACE_B200_E_SHAPE, "scatter_rows: too many rows");
    scatter_rows_kernel<<<static_cast<unsigned>(nblk), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(src), index, rows, D / 4, reinterpret_cast<float4*>(dst));
    AB_CHECK_CUDA(cudaGetLastError());
    return ARCFACE_B200_OK;
}

__global__ void __launch_bounds__(256) accumulate_kernel(float4* __restrict__ dst, const float4* __restrict__ src, int64_t n4) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        float4 a = dst[i];
        const float4 b = ldg_stream(src + i);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        dst[i] = a;
    }
}

extern "C" int32_t arcface_b200_accumulate(float* dst, const float* src, int64_t n, void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(dst && src, ARCFACE_B200_E_ARG, "accumulate: null pointer");
    AB_REQUIRE(n >= 0 && n % 4 == 0, ARCFACE_B200_E_SHAPE, "accumulate: n must be a non-negative multiple of 4");
    AB_REQUIRE(aligned16(dst) && aligned16(src), ARCFACE_B200_E_LAYOUT, "accumulate: pointers must be 16-byte aligned");
    if (n == 0) return ARCFACE_B200_OK;
    const int64_t n4 = n / 4;
    const int64_t want = (n4 + 255) / 256;
    const int grid = static_cast<int>(want < 148 * 16 ? want : 148 * 16);
    accumulate_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<float4*>(dst),
                                                                        reinterpret_cast<const float4*>(src), n4);
    AB_CHECK_CUDA(cudaGetLastError());
    return ARCFACE_B200_OK;
}

extern "C" int32_t arcface_b200_normalize_cast3(const float* src, int64_t rows, int32_t D, int32_t order, uint16_t* dst3,
                                                float* inv_norm, uint16_t* dst_t, int64_t ld_t, void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(src && dst3 && inv_norm, ARCFACE_B200_E_ARG, "normalize_cast3: null pointer");
    AB_REQUIRE(rows >= 0 && D >= 8 && D % 8 == 0, ARCFACE_B200_E_SHAPE, "normalize_cast3: D=%d must be a positive multiple of 8", D);
    AB_REQUIRE(order == 0 || order == 1, ARCFACE_B200_E_ARG, "normalize_cast3: order must be 0 (hi|hi|lo) or 1 (hi|lo|hi)");
    AB_REQUIRE(aligned16(src) && aligned16(dst3), ARCFACE_B200_E_LAYOUT, "normalize_cast3: pointers must be 16-byte aligned");
    AB_REQUIRE(dst_t == nullptr || ld_t >= rows, ARCFACE_B200_E_LAYOUT, "normalize_cast3: ld_t < rows (dst_t is [D][3 * ld_t])");
    if (rows == 0) return ARCFACE_B200_OK;
    const int wpb = 8;
    const int64_t nblk = (rows + wpb - 1) / wpb;
    AB_REQUIRE(nblk < (1ll << 31), ARCFACE_B200_E_SHAPE, "normalize_cast3: too many rows");
    normalize_cast3_kernel<<<static_cast<unsigned>(nblk), wpb * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        src, rows, D, order, reinterpret_cast<__nv_bfloat16*>(dst3), inv_norm, reinterpret_cast<__nv_bfloat16*>(dst_t), ld_t);
    AB_CHECK_CUDA(cudaGetLastError());
    return ARCFACE_B200_OK;
}

int32_t ab::launch_dw_lo_correction(float* dw, const void* what3, const float* q, int q_slots, const float* inv_nw,
                                    int64_t C, int D, cudaStream_t st) {
    if (C == 0) return ARCFACE_B200_OK;
    dw_lo_correction_kernel<<<static_cast<unsigned>((C + 7) / 8), 256, 0, st>>>(
        dw, static_cast<const __nv_bfloat16*>(what3), q, q_slots, inv_nw, C, D);
    AB_CHECK_CUDA(cudaGetLastError());
    return ARCFACE_B200_OK;
}

extern "C" int32_t arcface_b200_two_stream_concat(const float* a, const float* b, int32_t B, int32_t D1, int32_t D2,
                                                  float* out, float* inv1, float* inv2, void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(a && b && out && inv1 && inv2, ARCFACE_B200_E_ARG, "two_stream_concat: null pointer");
    AB_REQUIRE(B >= 0 && D1 >= 4 && D2 >= 4 && D1 % 4 == 0 && D2 % 4 == 0, ARCFACE_B200_E_SHAPE,
               "two_stream_concat: widths %d / %d must be positive multiples of 4", D1, D2);
    AB_REQUIRE(aligned16(a) && aligned16(b) && aligned16(out), ARCFACE_B200_E_LAYOUT,
               "two_stream_concat: pointers must be 16-byte aligned");
    if (B == 0) return ARCFACE_B200_OK;
    two_stream_concat_kernel<<<(B + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, B, D1, D2, out, inv1, inv2);
    AB_CHECK_CUDA(cudaGetLastError());
    return ARCFACE_B200_OK;
}

extern "C" int32_t arcface_b200_two_stream_concat_bwd(const float* emb, const float* inv1, const float* inv2,
                                                      const float* grad, int32_t B, int32_t D1, int32_t D2, float* da,
                                                      float* db, void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(emb && inv1 && inv2 && grad && da && db, ARCFACE_B200_E_ARG, "two_stream_concat_bwd: null pointer");
    AB_REQUIRE(B >= 0 && D1 >= 4 && D2 >= 4 && D1 % 4 == 0 && D2 % 4 == 0, ARCFACE_B200_E_SHAPE,
               "two_stream_concat_bwd: widths %d / %d must be positive multiples of 4", D1, D2);
    AB_REQUIRE(aligned16(emb) && aligned16(grad) && aligned16(da) && aligned16(db), ARCFACE_B200_E_LAYOUT,
               "two_stream_concat_bwd: pointers must be 16-byte aligned");
    if (B == 0) return ARCFACE_B200_OK;
    two_stream_concat_bwd_kernel<<<(B + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(emb, inv1, inv2, grad, B, D1,
                                                                                            D2, da, db);
    AB_CHECK_CUDA(cudaGetLastError());
    return ARCFACE_B200_OK;
}

extern "C" int32_t arcface_b200_label_margin(const float* x, const float* w, const float* inv_nx, const float* inv_nw,
                                             const int64_t* label, int32_t B, int32_t D, int64_t C_local,
                                             int64_t class_offset, int64_t C_total, float s, float cos_m, float sin_m,
                                             float th, float mm, int32_t easy_margin, float* t_label, float* z_label,
                                             float* dphi, int32_t* label_local, int32_t* bad_label_flag, void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(x && w && inv_nx && label && t_label && z_label && dphi && label_local, ARCFACE_B200_E_ARG,
               "label_margin: null pointer");
    AB_REQUIRE(B >= 0 && D >= 8 && D % 8 == 0 && C_local >= 1, ARCFACE_B200_E_SHAPE, "label_margin: bad shape");
    AB_REQUIRE(aligned16(x) && aligned16(w), ARCFACE_B200_E_LAYOUT, "label_margin: x / w must be 16-byte aligned");
    if (B == 0) return ARCFACE_B200_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    label_margin_kernel<<<(B + 7) / 8, 256, 0, st>>>(x, w, inv_nx, inv_nw, label, B, D, C_local, class_offset, C_total,
                                                     s, cos_m, sin_m, th, mm, easy_margin, t_label, z_label, dphi,
                                                     label_local, bad_label_flag);
    AB_CHECK_CUDA(cudaGetLastError());
    return ARCFACE_B200_OK;
}

extern "C" int32_t arcface_b200_combine_partials(const float* part_max, const float* part_sum, const int32_t* part_arg,
                                                 int32_t n_parts, int32_t B, int64_t class_offset, float* row_max,
                                                 float* row_sum, int64_t* row_arg, void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(part_max && part_sum && part_arg && row_max && row_sum && row_arg, ARCFACE_B200_E_ARG,
               "combine_partials: null pointer");
    AB_REQUIRE(n_parts >= 1 && B >= 0, ARCFACE_B200_E_SHAPE, "combine_partials: bad shape");
    if (B == 0) return ARCFACE_B200_OK;
    combine_partials_kernel<<<(B + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        part_max, part_sum, part_arg, n_parts, B, class_offset, row_max, row_sum, row_arg);
    AB_CHECK_CUDA(cudaGetLastError());
    return ARCFACE_B200_OK;
}

extern "C" int32_t arcface_b200_finalize_rows_strided(const float* rows_max, const float* rows_sum,
                                                      const int64_t* rows_arg, const float* rows_z_label,
                                                      const int64_t* label, int32_t n_ranks, int32_t B,
                                                      int64_t rank_stride_f32, int64_t rank_stride_i64, float* lse,
                                                      int64_t* argmax, float* z_label_out, float* one_minus_p,
                                                      float* loss, void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(rows_max && rows_sum && rows_arg && rows_z_label && label && lse && argmax && z_label_out &&
                   one_minus_p && loss,
               ARCFACE_B200_E_ARG, "finalize_rows: null pointer");
    AB_REQUIRE(n_ranks >= 1 && B >= 1, ARCFACE_B200_E_SHAPE, "finalize_rows: bad shape");
    AB_REQUIRE(rank_stride_f32 >= B && rank_stride_i64 >= B, ARCFACE_B200_E_LAYOUT, "finalize_rows: rank stride < B");
    const int threads = B <= 256 ? 256 : (B <= 512 ? 512 : 1024);   // a power of two (tree sum)
    finalize_rows_kernel<<<1, threads, 0, static_cast<cudaStream_t>(stream)>>>(rows_max, rows_sum, rows_arg, rows_z_label,
                                                                          label, n_ranks, B, rank_stride_f32,
                                                                          rank_stride_i64, lse, argmax, z_label_out,
                                                                          one_minus_p, loss);
    AB_CHECK_CUDA(cudaGetLastError());
    return ARCFACE_B200_OK;
}

extern "C" int32_t arcface_b200_finalize_rows(const float* rows_max, const float* rows_sum, const int64_t* rows_arg,
                                              const float* rows_z_label, const int64_t* label, int32_t n_ranks,
                                              int32_t B, float* lse, int64_t* argmax, float* z_label_out,
                                              float* one_minus_p, float* loss, void* stream) {
    return arcface_b200_finalize_rows_strided(rows_max, rows_sum, rows_arg, rows_z_label, label, n_ranks, B, B, B, lse,
                                              argmax, z_label_out, one_minus_p, loss, stream);
}

extern "C" int32_t arcface_b200_normalize_bwd_x(const float* x, const float* inv_nx, const float* dxhat, int32_t B,
                                                int32_t D, float* dx, void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(x && inv_nx && dxhat && dx, ARCFACE_B200_E_ARG, "normalize_bwd_x: null pointer");
    AB_REQUIRE(B >= 0 && D >= 8 && D % 8 == 0, ARCFACE_B200_E_SHAPE, "normalize_bwd_x: bad shape");
    AB_REQUIRE(aligned16(x) && aligned16(dxhat) && aligned16(dx), ARCFACE_B200_E_LAYOUT,
               "normalize_bwd_x: pointers must be 16-byte aligned");
    if (B == 0) return ARCFACE_B200_OK;
    normalize_bwd_x_kernel<<<(B + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, inv_nx, dxhat, B, D, dx);
    AB_CHECK_CUDA(cudaGetLastError());
    return ARCFACE_B200_OK;
}
