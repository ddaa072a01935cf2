// Host-side helpers: error string, device queries, TMA tensor maps.
#include "host_util.h"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

namespace ab {

static thread_local char g_err[512] = "";

int32_t set_error(int32_t code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
const char* last_error() { return g_err; }

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        cached[dev] = n;
    }
    return cached[dev];
}

int32_t check_arch() {
    static int ok[64] = {0};  // 0 unknown, 1 ok, -1 bad
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return set_error(ARCFACE_B200_E_CUDA, "cudaGetDevice failed: %s", cudaGetErrorString(e));
    if (dev < 0 || dev >= 64) return set_error(ARCFACE_B200_E_ARCH, "device index %d out of range", dev);
    if (ok[dev] == 0) {
        int major = 0;
        e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
        if (e != cudaSuccess)
            return set_error(ARCFACE_B200_E_CUDA, "cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
        ok[dev] = (major == 10) ? 1 : -1;
    }
    if (ok[dev] < 0)
        return set_error(ARCFACE_B200_E_ARCH, "device %d is not compute capability 10.x (B200 / sm_100a required; "
                                               "there is no fallback path)", dev);
    return ARCFACE_B200_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int32_t make_tmap_kmajor(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows,
                         uint64_t row_stride_elems, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(ARCFACE_B200_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if (!aligned16(base) || (row_stride_elems * 2) % 16 != 0)
        return set_error(ARCFACE_B200_E_LAYOUT, "TMA operand must be 16-byte aligned with a 16-byte multiple row stride");
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {row_stride_elems * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(ARCFACE_B200_E_CUDA, "cuTensorMapEncodeTiled (K-major %llu x %llu) failed: CUresult %d",
                         (unsigned long long)rows, (unsigned long long)inner, (int)r);
    return ARCFACE_B200_OK;
}

int32_t make_tmap_mnmajor(CUtensorMap* out, const void* base, uint64_t mn, uint64_t k_rows,
                          uint64_t row_stride_elems) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(ARCFACE_B200_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if (!aligned16(base) || (row_stride_elems * 2) % 16 != 0)
        return set_error(ARCFACE_B200_E_LAYOUT, "TMA operand must be 16-byte aligned with a 16-byte multiple row stride");
    cuuint64_t dims[2] = {mn, k_rows};
    cuuint64_t strides[1] = {row_stride_elems * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(ARCFACE_B200_E_CUDA, "cuTensorMapEncodeTiled (MN-major %llu x %llu) failed: CUresult %d",
                         (unsigned long long)k_rows, (unsigned long long)mn, (int)r);
    return ARCFACE_B200_OK;
}

int32_t make_tmap_store(CUtensorMap* out, const void* base, int elem_bytes, uint64_t cols, uint64_t rows,
                        uint64_t row_stride_elems) {
    return make_tmap_store_box(out, base, elem_bytes, cols, rows, row_stride_elems, 128);
}

int32_t make_tmap_store_box(CUtensorMap* out, const void* base, int elem_bytes, uint64_t cols, uint64_t rows,
                            uint64_t row_stride_elems, int box_bytes) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(ARCFACE_B200_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if (!aligned16(base) || (row_stride_elems * elem_bytes) % 16 != 0 || (elem_bytes != 2 && elem_bytes != 4))
        return set_error(ARCFACE_B200_E_LAYOUT, "TMA store target must be 16-byte aligned with a 16-byte multiple row stride");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {row_stride_elems * elem_bytes};
    if (box_bytes != 128 && box_bytes != 64)
        return set_error(ARCFACE_B200_E_ARG, "TMA store box must be 64 or 128 bytes wide");
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_bytes / elem_bytes), 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                    const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    box_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(ARCFACE_B200_E_CUDA, "cuTensorMapEncodeTiled (store %llu x %llu) failed: CUresult %d",
                         (unsigned long long)rows, (unsigned long long)cols, (int)r);
    return ARCFACE_B200_OK;
}

}  // namespace ab
