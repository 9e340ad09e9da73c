// Training losses on the device -- reference: PredictionLoss.forward / gaussian_2d
// (trajnetbaselines/lstm/loss.py:24-91) and CollisionLoss (loss.py:138-162).
//
// One thread per (predicted frame, scene): value and the analytic derivative wrt the five
// outputs of the primary track, so the ~30 elementwise autograd nodes of the reference's
// expression collapse into one launch.  Arithmetic is done in double and rounded once; the
// reductions (mean / per-scene sums) stay with the caller on the [T, B] result.
#include <math_constants.h>

#include "common.cuh"

namespace tb2 {

// -log(0.01 + bg * N(x | mu, 3, 3, 0) + (0.99 - bg) * N(x | mu, s1, s2, rho))
__global__ void prediction_loss_kernel(const float* __restrict__ inputs, const float* __restrict__ targets,
                                       const int* __restrict__ prim, int T, int M, int B, float background_rate,
                                       float* __restrict__ values, float* __restrict__ dinputs) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * B) return;
    const int t = idx / B, b = idx - t * B;
    const size_t row = (size_t)t * M + prim[b];
    const float* in = inputs + row * 5;
    const double mu1 = in[0], mu2 = in[1], s1 = in[2], s2 = in[3], rho = in[4];
    const double n1 = (double)targets[row * 2] - mu1, n2 = (double)targets[row * 2 + 1] - mu2;
    const double two_pi = 6.283185307179586476925286766559;
    // background component: sigma = 3, rho = 0 (loss.py:73-76)
    const double g0 = exp(-(n1 * n1 + n2 * n2) / 18.0) / (two_pi * 9.0);
    const double q = 1.0 - rho * rho, s12 = s1 * s2;
    const double z = (n1 / s1) * (n1 / s1) + (n2 / s2) * (n2 / s2) - 2.0 * rho * n1 * n2 / s12;
    const double g1 = exp(-z / (2.0 * q)) / (two_pi * s12 * sqrt(q));
    const double bg = background_rate, w = 0.99 - bg;
    const double D = 0.01 + bg * g0 + w * g1;
    values[idx] = (float)(-log(D));
    if (dinputs) {
        const double dg0_m1 = g0 * n1 / 9.0, dg0_m2 = g0 * n2 / 9.0;
        const double dg1_m1 = g1 * (n1 / (s1 * s1) - rho * n2 / s12) / q;
        const double dg1_m2 = g1 * (n2 / (s2 * s2) - rho * n1 / s12) / q;
        const double dg1_s1 = g1 * ((n1 * n1 / (s1 * s1 * s1) - rho * n1 * n2 / (s1 * s12)) / q - 1.0 / s1);
        const double dg1_s2 = g1 * ((n2 * n2 / (s2 * s2 * s2) - rho * n1 * n2 / (s2 * s12)) / q - 1.0 / s2);
        const double dg1_r = g1 * (n1 * n2 / (s12 * q) - z * rho / (q * q) + rho / q);
        float* d = dinputs + (size_t)idx * 5;
        d[0] = (float)(-(bg * dg0_m1 + w * dg1_m1) / D);
        d[1] = (float)(-(bg * dg0_m2 + w * dg1_m2) / D);
        d[2] = (float)(-w * dg1_s1 / D);
        d[3] = (float)(-w * dg1_s2 / D);
        d[4] = (float)(-w * dg1_r / D);
    }
}

// L2Loss (loss.py:93-135): per (frame, scene) mean over the two coordinates of the squared error of the
// primary's predicted mean; the x100 multiplier and the mean reductions stay with the caller.
__global__ void l2_loss_kernel(const float* __restrict__ inputs, const float* __restrict__ targets,
                               const int* __restrict__ prim, int T, int M, int B, float* __restrict__ values,
                               float* __restrict__ dinputs) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * B) return;
    const int t = idx / B, b = idx - t * B;
    const size_t row = (size_t)t * M + prim[b];
    const float dx = inputs[row * 5] - targets[row * 2], dy = inputs[row * 5 + 1] - targets[row * 2 + 1];
    values[idx] = 0.5f * (dx * dx + dy * dy);
    if (dinputs) {
        float* d = dinputs + (size_t)idx * 5;
        d[0] = dx; d[1] = dy; d[2] = 0.f; d[3] = 0.f; d[4] = 0.f;
    }
}

// col_wt * sum over frames and neighbours within col_distance of (1 - dist / col_distance); the
// neighbours are constants (detached), NaN coordinates count as -1000 (loss.py:148-161).
__global__ void collision_loss_kernel(const float2* __restrict__ pos, const int* __restrict__ scene_off, int T,
                                      int M, int B, float col_wt, float col_distance,
                                      float* __restrict__ loss_scene, float2* __restrict__ dprim) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * B) return;
    const int t = idx / B, b = idx - t * B;
    const int start = scene_off[b], end = scene_off[b + 1];
    float2 p = pos[(size_t)t * M + start];
    if (isnan(p.x)) p.x = -1000.f;
    if (isnan(p.y)) p.y = -1000.f;
    float loss = 0.f, gx = 0.f, gy = 0.f;
    for (int j = start + 1; j < end; ++j) {
        float2 n = pos[(size_t)t * M + j];
        if (isnan(n.x)) n.x = -1000.f;
        if (isnan(n.y)) n.y = -1000.f;
        const float dx = p.x - n.x, dy = p.y - n.y;
        const float d = sqrtf(dx * dx + dy * dy);
        if (d <= col_distance) {
            loss += col_wt * (1.f - d / col_distance);
            if (d > 0.f) {            // the norm's subgradient at 0 is 0 (as in torch)
                gx -= col_wt * dx / (d * col_distance);
                gy -= col_wt * dy / (d * col_distance);
            }
        }
    }
    loss_scene[idx] = loss;
    if (dprim) dprim[idx] = make_float2(gx, gy);
}

}  // namespace tb2

using namespace tb2;

extern "C" {

int tb2_prediction_loss(const float* inputs, const float* targets, const int32_t* primary_rows, int32_t T,
                        int32_t M, int32_t B, float background_rate, float* values_out, float* dinputs_out,
                        void* stream) {
    TB2_REQUIRE(inputs && targets && primary_rows && values_out, "null argument");
    TB2_REQUIRE(T >= 0 && M >= 0 && B >= 0, "negative size");
    if (T * B == 0) return TB2_OK;
    cudaStream_t st = (cudaStream_t)stream;
    {
        KernelTimer kt("prediction_loss", st);
        prediction_loss_kernel<<<(T * B + 127) / 128, 128, 0, st>>>(inputs, targets, primary_rows, T, M, B,
                                                                    background_rate, values_out, dinputs_out);
    }
    TB2_LAUNCH_CHECK();
    return TB2_OK;
}

int tb2_l2_loss(const float* inputs, const float* targets, const int32_t* primary_rows, int32_t T, int32_t M,
                int32_t B, float* values_out, float* dinputs_out, void* stream) {
    TB2_REQUIRE(inputs && targets && primary_rows && values_out, "null argument");
    TB2_REQUIRE(T >= 0 && M >= 0 && B >= 0, "negative size");
    if (T * B == 0) return TB2_OK;
    cudaStream_t st = (cudaStream_t)stream;
    {
        KernelTimer kt("l2_loss", st);
        l2_loss_kernel<<<(T * B + 127) / 128, 128, 0, st>>>(inputs, targets, primary_rows, T, M, B, values_out,
                                                            dinputs_out);
    }
    TB2_LAUNCH_CHECK();
    return TB2_OK;
}

int tb2_collision_loss(const tb2_layout* l, const float* positions, int32_t T, float col_wt, float col_distance,
                       float* loss_out, float* dprimary_out, void* stream) {
    TB2_REQUIRE(l && positions && loss_out, "null argument");
    if (T * l->B == 0) return TB2_OK;
    cudaStream_t st = (cudaStream_t)stream;
    {
        KernelTimer kt("collision_loss", st);
        collision_loss_kernel<<<(T * l->B + 127) / 128, 128, 0, st>>>(
            (const float2*)positions, l->scene_off, T, l->M, l->B, col_wt, col_distance, loss_out,
            (float2*)dprimary_out);
    }
    TB2_LAUNCH_CHECK();
    return TB2_OK;
}

}  // extern "C"
