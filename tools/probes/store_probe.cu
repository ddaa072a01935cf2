// Store-path topology probe: how many bytes per clock can ONE SM push towards L2 / HBM, and is the limit per SM or per
// pair of SMs (TPC)?  One persistent CTA per SM (forced by its shared-memory request); a CTA is active iff its %smid
// passes the chosen predicate; every active CTA streams full 128-byte lines (st.global.v4 from every lane) over a private
// region.  Prints aggregate GB/s and bytes / clock / active SM for each predicate.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/probes/store_probe.cu -o tools/probes/store_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__global__ void __launch_bounds__(512, 1)
store_kernel(float4* base, size_t region_f4, int iters, int mode, int param, int* active_count) {
    extern __shared__ unsigned char pad[];
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    bool active = true;
    if (mode == 1) active = (smid & 1u) == 0u;              // one SM of every consecutive pair
    if (mode == 2) active = smid < (unsigned)param;        // the first `param` SMs
    if (mode == 3) active = (smid % (unsigned)param) == 0u; // every param-th SM
    if (!active) return;
    if (threadIdx.x == 0) atomicAdd(active_count, 1);
    float4* p = base + (size_t)blockIdx.x * region_f4;
    const float4 v = make_float4(1.f, 2.f, 3.f, (float)smid);
    for (int it = 0; it < iters; ++it)
        for (size_t i = threadIdx.x; i < region_f4; i += blockDim.x) p[i] = v;
    if (pad[0] == 123 && threadIdx.x == 9999) p[0] = v;
}

int main() {
    int dev = 0, nsm = 0, clk = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    const size_t region_bytes = 32u << 20;   // per CTA
    const size_t region_f4 = region_bytes / 16;
    float4* buf;
    int* cnt;
    cudaMalloc(&buf, region_bytes * nsm);
    cudaMalloc(&cnt, 4);
    cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct { int mode, param; const char* name; } cases[] = {
        {0, 0, "all SMs"}, {1, 0, "even smid only (one per pair)"}, {2, 74, "first 74 smids"}, {2, 64, "first 64 smids"},
        {2, 32, "first 32 smids"}, {3, 4, "every 4th smid"}, {2, 8, "first 8 smids"}, {3, 16, "every 16th smid"}};
    for (auto& c : cases) {
        const int iters = 4;
        cudaMemset(cnt, 0, 4);
        store_kernel<<<nsm, 512, 200 * 1024>>>(buf, region_f4, 1, c.mode, c.param, cnt);   // warm-up
        cudaMemset(cnt, 0, 4);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        store_kernel<<<nsm, 512, 200 * 1024>>>(buf, region_f4, iters, c.mode, c.param, cnt);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        int active = 0;
        cudaMemcpy(&active, cnt, 4, cudaMemcpyDeviceToHost);
        const double bytes = (double)region_bytes * iters * active;
        printf("%-34s active SMs %3d  %.3f ms  %8.1f GB/s  %6.2f B/clk/SM at %d MHz (nominal max)\n", c.name, active, ms,
               bytes / ms / 1e6, bytes / (ms * 1e-3) / active / (clk * 1e3), clk / 1000);
    }
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
