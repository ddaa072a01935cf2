// Does an SM's store rate drop when it also pulls data in?  8 store warps (pattern P4 of store_pattern_probe.cu: lane
// pairs write 64 contiguous bytes per row) beside `lw` load warps that stream a private region from L2 / HBM with
// 128-bit loads.  64 active SMs (the size of the backward's dW role).
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ void st8(void* p, float v) {
    unsigned u = __float_as_uint(v);
    asm volatile("st.global.v8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"l"(p), "r"(u) : "memory");
}

__global__ void __launch_bounds__(1024, 1)
k(unsigned char* sbase, const float4* lbase, size_t bytes_per_warp, size_t load_f4_per_warp, int iters, int store_warps,
  int nsm_active, float* sink, unsigned long long* cycles) {
    extern __shared__ unsigned char pad[];
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (smid >= (unsigned)nsm_active) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarp = blockDim.x >> 5;
    const unsigned long long t0 = clock64();
    if (warp < store_warps) {
        unsigned char* wbase = sbase + ((size_t)blockIdx.x * store_warps + warp) * bytes_per_warp;
        const size_t pitch = 2048, boxes = bytes_per_warp / (32 * pitch);
        const float v = (float)smid;
        for (int it = 0; it < iters; ++it)
            for (size_t bx = 0; bx < boxes; ++bx)
                for (int cg = 0; cg < 16; ++cg) {
                    unsigned char* box = wbase + bx * 32 * pitch + cg * 128;
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int h = 0; h < 2; ++h) st8(box + (size_t)(16 * j + (lane >> 1)) * pitch + h * 64 + (lane & 1) * 32, v);
                }
        if (lane == 0 && warp == 0) cycles[blockIdx.x] = clock64() - t0;
    } else {
        const float4* p = lbase + ((size_t)blockIdx.x * (nwarp - store_warps) + (warp - store_warps)) * load_f4_per_warp;
        float acc = 0.f;
        for (int it = 0; it < iters * 64; ++it) {
            // keeps loading until the store warps are done (bounded)
            for (size_t i = lane; i < load_f4_per_warp; i += 32 * 4) {
                const float4 a = p[i], b = p[i + 32], c = p[i + 64], d = p[i + 96];
                acc += a.x + b.y + c.z + d.w;
            }
            if (*(volatile unsigned long long*)&cycles[blockIdx.x] != 0ull) break;
        }
        if (acc == 123.456f) sink[0] = acc;
    }
    if (pad[0] == 123 && threadIdx.x == 9999) sink[1] = 1;
}

int main() {
    int nsm = 0, clk = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const size_t bytes_per_warp = 4u << 20, load_f4 = (2u << 20) / 16;
    unsigned char* sbuf;
    float4* lbuf;
    float* sink;
    unsigned long long* cyc;
    cudaMalloc(&sbuf, bytes_per_warp * 8 * nsm);
    cudaMalloc(&lbuf, (size_t)load_f4 * 16 * 24 * nsm);
    cudaMemset(lbuf, 0, (size_t)load_f4 * 16 * 24 * nsm);
    cudaMalloc(&sink, 8);
    cudaMalloc(&cyc, 8 * nsm);
    for (int lw : {0, 2, 4, 8, 16, 24}) {
        const int active = 64, iters = 2;
        cudaMemset(cyc, 0, 8 * nsm);
        k<<<nsm, (8 + lw) * 32, 200 * 1024>>>(sbuf, lbuf, bytes_per_warp, load_f4, 1, 8, active, sink, cyc);
        cudaMemset(cyc, 0, 8 * nsm);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<<<nsm, (8 + lw) * 32, 200 * 1024>>>(sbuf, lbuf, bytes_per_warp, load_f4, iters, 8, active, sink, cyc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        unsigned long long h[256];
        cudaMemcpy(h, cyc, 8 * nsm, cudaMemcpyDeviceToHost);
        double mean = 0;
        int n = 0;
        for (int i = 0; i < nsm; ++i) if (h[i]) { mean += (double)h[i]; ++n; }
        mean /= (n ? n : 1);
        const double bytes_sm = (double)bytes_per_warp * 8 * iters;
        printf("load warps %2d: store phase %.0f cycles/SM -> %.2f B/clk/SM stored (kernel %.3f ms)\n", lw, mean, bytes_sm / mean, ms);
    }
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
