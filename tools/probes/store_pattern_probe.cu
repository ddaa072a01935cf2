// Which register -> global store patterns reach the per-SM store rate (~32 B/clk, tools/probes/store_probe.cu)?
// Models an epilogue warp that owns 32 rows x 128 B of a row-major fp32 matrix with a 2 KB row pitch:
//   P0  coalesced: every instruction writes 4 full 128-byte lines (lane = 16 B of a line)            [after a transpose]
//   P1  lane = row, 32 B per instruction (st.global.v8): 32 lines touched, one sector each, 4 instr complete a line
//   P2  quad = row: lanes 4i..4i+3 write the four 32-byte sectors of row i (v8): 8 full lines per instruction
//   P3  lane = row, 16 B per instruction (st.global.v4): 8 instructions complete a line
//   P4  pair = row half: lanes 2i, 2i+1 write 64 B of row i (v8): 16 half lines per instruction
// One CTA per SM on the first `nsm_active` SMs, `warps` warps each.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/probes/store_pattern_probe.cu -o tools/probes/store_pattern_probe
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ void st8(void* p, float v) {
    unsigned u = __float_as_uint(v);
    asm volatile("st.global.v8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"l"(p), "r"(u) : "memory");
}
__device__ __forceinline__ void st4(void* p, float v) {
    asm volatile("st.global.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p), "f"(v) : "memory");
}

__global__ void __launch_bounds__(1024, 1)
pattern_kernel(unsigned char* base, size_t bytes_per_warp, int iters, int pattern, int nsm_active) {
    extern __shared__ unsigned char pad[];
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (smid >= (unsigned)nsm_active) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarp = blockDim.x >> 5;
    unsigned char* wbase = base + ((size_t)blockIdx.x * nwarp + warp) * bytes_per_warp;
    const size_t pitch = 2048;                       // row pitch: D = 512 fp32
    const size_t boxes = bytes_per_warp / (32 * pitch);   // a box = 32 rows; the warp walks the 16 column groups of each
    const float v = (float)smid;
    for (int it = 0; it < iters; ++it)
        for (size_t bx = 0; bx < boxes; ++bx)
            for (int cg = 0; cg < 16; ++cg) {        // 16 x 128 B = one 2 KB row
                unsigned char* box = wbase + bx * 32 * pitch + cg * 128;
                if (pattern == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) st4(box + (size_t)(4 * j + (lane >> 3)) * pitch + (lane & 7) * 16, v);
                } else if (pattern == 1) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) st8(box + (size_t)lane * pitch + k * 32, v);
                } else if (pattern == 2) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) st8(box + (size_t)(8 * j + (lane >> 2)) * pitch + (lane & 3) * 32, v);
                } else if (pattern == 3) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) st4(box + (size_t)lane * pitch + k * 16, v);
                } else {
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            st8(box + (size_t)(16 * j + (lane >> 1)) * pitch + h * 64 + (lane & 1) * 32, v);
                }
            }
    if (pad[0] == 123 && threadIdx.x == 9999) wbase[0] = 1;
}

int main() {
    int nsm = 0, clk = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    cudaFuncSetAttribute(pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const size_t bytes_per_warp = 4u << 20;
    unsigned char* buf;
    cudaMalloc(&buf, bytes_per_warp * 32 * nsm);
    const char* names[5] = {"P0 coalesced v4 (4 lines/instr)", "P1 lane=row v8 (32 sectors/instr)", "P2 quad=row v8 (8 lines/instr)",
                            "P3 lane=row v4 (32 half-sectors)", "P4 pair=half row v8"};
    for (int active : {64, 148})
        for (int warps : {8, 16})
            for (int pat = 0; pat < 5; ++pat) {
                const int iters = 2;
                pattern_kernel<<<nsm, warps * 32, 200 * 1024>>>(buf, bytes_per_warp, 1, pat, active);
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0);
                cudaEventCreate(&e1);
                cudaEventRecord(e0);
                pattern_kernel<<<nsm, warps * 32, 200 * 1024>>>(buf, bytes_per_warp, iters, pat, active);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms = 0;
                cudaEventElapsedTime(&ms, e0, e1);
                const double bytes = (double)bytes_per_warp * warps * iters * active;
                printf("SMs %3d warps %2d %-34s %7.3f ms %8.1f GB/s %6.2f B/clk/SM\n", active, warps, names[pat], ms,
                       bytes / ms / 1e6, bytes / (ms * 1e-3) / active / (clk * 1e3));
            }
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
