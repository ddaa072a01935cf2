// C-ABI odds and ends of libarcface_b200: version / error string / device check, and the one-call
// host-buffer step (H2D copies -> K1(x) -> label margin -> K1(w)+K2 -> combine -> K3 -> normalise backward ->
// D2H copies) that a host without torch binds.  See include/arcface_b200.h.
#include "host_util.h"

#include <math.h>

using namespace ab;

extern "C" int32_t arcface_b200_version(int32_t* major, int32_t* minor) {
    if (major) *major = ARCFACE_B200_VERSION_MAJOR;
    if (minor) *minor = ARCFACE_B200_VERSION_MINOR;
    return ARCFACE_B200_OK;
}

extern "C" const char* arcface_b200_last_error(void) { return last_error(); }

extern "C" int32_t arcface_b200_device_ok(void) { return check_arch(); }

namespace {

struct StepPlan {
    int Bp;
    int n_parts;
    size_t bwd_bytes, fwd_bytes;
    size_t off_x, off_label, off_xhat, off_xhat_t, off_inv_nx, off_what, off_inv_nw, off_t, off_z, off_dphi,
        off_lab_local, off_flag, off_pmax, off_psum, off_parg, off_rmax, off_rsum, off_rarg, off_lse, off_argmax,
        off_zout, off_omp, off_loss, off_dxhat, off_dx, off_fwd, off_bwd, total;
};

size_t bump(size_t& cur, size_t bytes) {
    const size_t o = cur;
    cur += (bytes + 255) / 256 * 256;
    return o;
}

int32_t plan_step(int32_t B, int32_t D, int64_t C, StepPlan* pl) {
    if (int32_t rc = arcface_b200_forward_parts(B, D, C, &pl->n_parts)) return rc;
    if (int32_t rc = arcface_b200_backward_workspace_bytes(B, D, C, &pl->bwd_bytes)) return rc;
    if (int32_t rc = arcface_b200_forward_fused_workspace_bytes(B, D, C, &pl->fwd_bytes)) return rc;
    pl->Bp = ((B + 63) / 64) * 64;
    size_t cur = 0;
    const size_t b = static_cast<size_t>(B), d = static_cast<size_t>(D), c = static_cast<size_t>(C);
    pl->off_x = bump(cur, b * d * 4);
    pl->off_label = bump(cur, b * 8);
    pl->off_xhat = bump(cur, b * d * 2);
    pl->off_xhat_t = bump(cur, d * pl->Bp * 2);
    pl->off_inv_nx = bump(cur, b * 4);
    pl->off_what = bump(cur, c * d * 2);
    pl->off_inv_nw = bump(cur, c * 4);
    pl->off_t = bump(cur, b * 4);
    pl->off_z = bump(cur, b * 4);
    pl->off_dphi = bump(cur, b * 4);
    pl->off_lab_local = bump(cur, b * 4);
    pl->off_flag = bump(cur, 4);
    pl->off_pmax = bump(cur, b * pl->n_parts * 4);
    pl->off_psum = bump(cur, b * pl->n_parts * 4);
    pl->off_parg = bump(cur, b * pl->n_parts * 4);
    pl->off_rmax = bump(cur, b * 4);
    pl->off_rsum = bump(cur, b * 4);
    pl->off_rarg = bump(cur, b * 8);
    pl->off_lse = bump(cur, b * 4);
    pl->off_argmax = bump(cur, b * 8);
    pl->off_zout = bump(cur, b * 4);
    pl->off_omp = bump(cur, b * 4);
    pl->off_loss = bump(cur, 4);
    pl->off_dxhat = bump(cur, b * d * 4);
    pl->off_dx = bump(cur, b * d * 4);
    pl->off_fwd = bump(cur, pl->fwd_bytes);
    pl->off_bwd = bump(cur, pl->bwd_bytes);
    pl->total = cur;
    return ARCFACE_B200_OK;
}

}  // namespace

extern "C" int32_t arcface_b200_step_workspace_bytes(int32_t B, int32_t D, int64_t C, size_t* bytes) {
    AB_REQUIRE(bytes, ARCFACE_B200_E_ARG, "step_workspace_bytes: null pointer");
    StepPlan pl;
    if (int32_t rc = plan_step(B, D, C, &pl)) return rc;
    *bytes = pl.total;
    return ARCFACE_B200_OK;
}

extern "C" int32_t arcface_b200_step_host(const float* x_host, const int64_t* label_host, const float* w_dev,
                                          int32_t B, int32_t D, int64_t C, float s, float m, int32_t easy_margin,
                                          float grad_loss, float* loss_host, int64_t* argmax_host, float* dx_host,
                                          float* dw_dev, void* device_ws, size_t device_ws_bytes, void* stream) {
    if (int32_t rc = check_arch()) return rc;
    AB_REQUIRE(x_host && label_host && w_dev && loss_host && argmax_host && dx_host && dw_dev && device_ws,
               ARCFACE_B200_E_ARG, "step_host: null pointer");
    StepPlan pl;
    if (int32_t rc = plan_step(B, D, C, &pl)) return rc;
    AB_REQUIRE(device_ws_bytes >= pl.total, ARCFACE_B200_E_WORKSPACE, "step_host: workspace %zu < required %zu",
               device_ws_bytes, pl.total);
    AB_REQUIRE(aligned16(device_ws), ARCFACE_B200_E_LAYOUT, "step_host: workspace must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* ws = static_cast<uint8_t*>(device_ws);
    float* x = reinterpret_cast<float*>(ws + pl.off_x);
    int64_t* label = reinterpret_cast<int64_t*>(ws + pl.off_label);
    uint16_t* xhat = reinterpret_cast<uint16_t*>(ws + pl.off_xhat);
    uint16_t* xhat_t = reinterpret_cast<uint16_t*>(ws + pl.off_xhat_t);
    float* inv_nx = reinterpret_cast<float*>(ws + pl.off_inv_nx);
    uint16_t* what = reinterpret_cast<uint16_t*>(ws + pl.off_what);
    float* inv_nw = reinterpret_cast<float*>(ws + pl.off_inv_nw);
    float* t_label = reinterpret_cast<float*>(ws + pl.off_t);
    float* z_label = reinterpret_cast<float*>(ws + pl.off_z);
    float* dphi = reinterpret_cast<float*>(ws + pl.off_dphi);
    int32_t* lab_local = reinterpret_cast<int32_t*>(ws + pl.off_lab_local);
    int32_t* flag = reinterpret_cast<int32_t*>(ws + pl.off_flag);
    float* pmax = reinterpret_cast<float*>(ws + pl.off_pmax);
    float* psum = reinterpret_cast<float*>(ws + pl.off_psum);
    int32_t* parg = reinterpret_cast<int32_t*>(ws + pl.off_parg);
    float* rmax = reinterpret_cast<float*>(ws + pl.off_rmax);
    float* rsum = reinterpret_cast<float*>(ws + pl.off_rsum);
    int64_t* rarg = reinterpret_cast<int64_t*>(ws + pl.off_rarg);
    float* lse = reinterpret_cast<float*>(ws + pl.off_lse);
    int64_t* argmax = reinterpret_cast<int64_t*>(ws + pl.off_argmax);
    float* zout = reinterpret_cast<float*>(ws + pl.off_zout);
    float* omp = reinterpret_cast<float*>(ws + pl.off_omp);
    float* loss = reinterpret_cast<float*>(ws + pl.off_loss);
    float* dxhat = reinterpret_cast<float*>(ws + pl.off_dxhat);
    float* dx = reinterpret_cast<float*>(ws + pl.off_dx);

    const double md = static_cast<double>(m);
    const float cos_m = static_cast<float>(cos(md)), sin_m = static_cast<float>(sin(md));
    const float th = static_cast<float>(cos(M_PI - md)), mm = static_cast<float>(sin(M_PI - md) * md);

    AB_CHECK_CUDA(cudaMemcpyAsync(x, x_host, static_cast<size_t>(B) * D * 4, cudaMemcpyHostToDevice, st));
    AB_CHECK_CUDA(cudaMemcpyAsync(label, label_host, static_cast<size_t>(B) * 8, cudaMemcpyHostToDevice, st));
    AB_CHECK_CUDA(cudaMemsetAsync(flag, 0, 4, st));
    // the xhat^T padding columns [B, Bp) are never read (TMA extent = B), no need to clear them
    if (int32_t rc = arcface_b200_normalize_cast(x, B, D, xhat, inv_nx, xhat_t, pl.Bp, st)) return rc;
    if (int32_t rc = arcface_b200_label_margin(x, w_dev, inv_nx, nullptr, label, B, D, C, 0, C, s, cos_m, sin_m, th, mm,
                                               easy_margin, t_label, z_label, dphi, lab_local, flag, st))
        return rc;
    // K1 of the class weights runs inside the forward kernel (what / inv_nw are its outputs)
    if (int32_t rc = arcface_b200_forward_stats_fused(xhat, w_dev, lab_local, B, D, C, s, what, inv_nw, pmax, psum, parg,
                                                      pl.n_parts, ws + pl.off_fwd, pl.fwd_bytes, st))
        return rc;
    if (int32_t rc = arcface_b200_combine_partials(pmax, psum, parg, pl.n_parts, B, 0, rmax, rsum, rarg, st)) return rc;
    if (int32_t rc = arcface_b200_finalize_rows(rmax, rsum, rarg, z_label, label, 1, B, lse, argmax, zout, omp, loss, st)) return rc;
    if (int32_t rc = arcface_b200_backward(xhat, xhat_t, pl.Bp, what, inv_nw, lse, omp, dphi, lab_local, B, D, C, s,
                                           grad_loss / static_cast<float>(B), nullptr, dxhat, dw_dev, ws + pl.off_bwd,
                                           pl.bwd_bytes, st))
        return rc;
    if (int32_t rc = arcface_b200_normalize_bwd_x(x, inv_nx, dxhat, B, D, dx, st)) return rc;
    int32_t flag_host = 0;
    AB_CHECK_CUDA(cudaMemcpyAsync(loss_host, loss, 4, cudaMemcpyDeviceToHost, st));
    AB_CHECK_CUDA(cudaMemcpyAsync(argmax_host, argmax, static_cast<size_t>(B) * 8, cudaMemcpyDeviceToHost, st));
    AB_CHECK_CUDA(cudaMemcpyAsync(dx_host, dx, static_cast<size_t>(B) * D * 4, cudaMemcpyDeviceToHost, st));
    AB_CHECK_CUDA(cudaMemcpyAsync(&flag_host, flag, 4, cudaMemcpyDeviceToHost, st));
    AB_CHECK_CUDA(cudaStreamSynchronize(st));
    AB_REQUIRE(flag_host == 0, ARCFACE_B200_E_ARG, "step_host: a label is outside [0, %lld)", (long long)C);
    return ARCFACE_B200_OK;
}
