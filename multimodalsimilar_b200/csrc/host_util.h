// Host-side helpers of libarcface_b200: status codes / thread-local error string, TMA tensor-map
// construction through the driver entry point (no -lcuda link dependency), device queries.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>

#include "../../include/arcface_b200.h"

namespace ab {

int32_t set_error(int32_t code, const char* fmt, ...);
const char* last_error();

#define AB_CHECK_CUDA(expr)                                                                      \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return ab::set_error(ARCFACE_B200_E_CUDA, "%s failed: %s (%s:%d)", #expr,            \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);                    \
    } while (0)

#define AB_REQUIRE(cond, code, ...)                                 \
    do {                                                            \
        if (!(cond)) return ab::set_error((code), __VA_ARGS__);     \
    } while (0)

// Number of SMs of the current device (cached per device); 0 on failure.
int sm_count();
// ARCFACE_B200_OK iff the current device is compute capability 10.x.
int32_t check_arch();

// K-major operand: global [rows][inner] bf16 with row stride `row_stride_elems`; box = {64, box_rows}.
int32_t make_tmap_kmajor(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows,
                         uint64_t row_stride_elems, uint32_t box_rows);
// MN-major operand: global [k_rows][mn] bf16 (mn contiguous); box = {64 mn, 64 k rows}.  A tile is
// fetched as tile_mn / 64 boxes which land as [mn/64][BLOCK_K][64] in shared memory.
int32_t make_tmap_mnmajor(CUtensorMap* out, const void* base, uint64_t mn, uint64_t k_rows,
                          uint64_t row_stride_elems);

// Output tile map for StoreStager: global [rows][cols] of `elem_bytes`-wide elements (2 = bf16, 4 = fp32),
// box = {128 bytes of columns, 32 rows}, SWIZZLE_128B.
int32_t make_tmap_store(CUtensorMap* out, const void* base, int elem_bytes, uint64_t cols, uint64_t rows,
                        uint64_t row_stride_elems);

// Same with a box of {box_bytes of columns, 32 rows}: 128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B (half-size staging
// buffers that can be double-buffered).
int32_t make_tmap_store_box(CUtensorMap* out, const void* base, int elem_bytes, uint64_t cols, uint64_t rows,
                            uint64_t row_stride_elems, int box_bytes);

// Environment knobs exist in diagnostic builds only (make DIAG=1 -> libarcface_b200_diag.so); the release library
// never reads the environment.
#ifdef ARCFACE_B200_DIAG
static inline const char* diag_env(const char* name) { return getenv(name); }
#else
static inline const char* diag_env(const char*) { return nullptr; }
#endif

// rows.cu: dW[c] -= inv_nw[c] * (sum of q slots)[c] * what_lo[c] (bf16x3 backward, see k3_backward.cu)
int32_t launch_dw_lo_correction(float* dw, const void* what3, const float* q, int q_slots, const float* inv_nw,
                                int64_t C, int D, cudaStream_t st);

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace ab
