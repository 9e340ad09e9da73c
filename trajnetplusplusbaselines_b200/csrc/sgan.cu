// S-GAN generator glue between encoder and decoder -- reference: LSTMGenerator.adding_noise
// (trajnetbaselines/sgan/sgan.py:200-221): h <- cat(ReLU(Linear(H -> H - noise_dim)(h)), z) for
// every track (z is one noise vector shared by all tracks of the call), c unchanged.
#include "common.cuh"

namespace tb2 {

__global__ void __launch_bounds__(128) sgan_add_noise_kernel(const float* __restrict__ W,
                                                             const float* __restrict__ b,
                                                             const float* __restrict__ noise,
                                                             float* __restrict__ h, int M, int H, int nd) {
    extern __shared__ float row_s[];          // [H] the track's hidden state
    const int m = blockIdx.x;
    if (m >= M) return;
    float* hr = h + (size_t)m * H;
    for (int k = threadIdx.x; k < H; k += blockDim.x) row_s[k] = hr[k];
    __syncthreads();
    const int keep = H - nd;
    for (int u = threadIdx.x; u < H; u += blockDim.x) {
        float v;
        if (u < keep) {
            const float* w = W + (size_t)u * H;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            int k = 0;
            for (; k + 3 < H; k += 4) {
                a0 = fmaf(w[k], row_s[k], a0);
                a1 = fmaf(w[k + 1], row_s[k + 1], a1);
                a2 = fmaf(w[k + 2], row_s[k + 2], a2);
                a3 = fmaf(w[k + 3], row_s[k + 3], a3);
            }
            for (; k < H; ++k) a0 = fmaf(w[k], row_s[k], a0);
            v = fmaxf(((a0 + a1) + (a2 + a3)) + b[u], 0.f);
        } else {
            v = noise[u - keep];
        }
        hr[u] = v;
    }
}

// VAE.add_noise at test time (vae/vae.py:87-106): h[m] <- h[m] * ReLU(fc . z[m] + bias), one latent
// sample z[m] per track; c unchanged.
__global__ void __launch_bounds__(128) vae_scale_hidden_kernel(const float* __restrict__ W,
                                                               const float* __restrict__ b,
                                                               const float* __restrict__ z,
                                                               float* __restrict__ h, int M, int H, int L) {
    extern __shared__ float z_s[];            // [L] the track's latent sample
    const int m = blockIdx.x;
    if (m >= M) return;
    for (int k = threadIdx.x; k < L; k += blockDim.x) z_s[k] = z[(size_t)m * L + k];
    __syncthreads();
    for (int u = threadIdx.x; u < H; u += blockDim.x) {
        const float* w = W + (size_t)u * L;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int k = 0;
        for (; k + 3 < L; k += 4) {
            a0 = fmaf(w[k], z_s[k], a0);
            a1 = fmaf(w[k + 1], z_s[k + 1], a1);
            a2 = fmaf(w[k + 2], z_s[k + 2], a2);
            a3 = fmaf(w[k + 3], z_s[k + 3], a3);
        }
        for (; k < L; ++k) a0 = fmaf(w[k], z_s[k], a0);
        h[(size_t)m * H + u] *= fmaxf(((a0 + a1) + (a2 + a3)) + b[u], 0.f);
    }
}

}  // namespace tb2

using namespace tb2;

extern "C" int tb2_vae_scale_hidden(const float* weight, const float* bias, const float* z, float* h, int32_t M,
                                    int32_t H, int32_t latent_dim, void* stream) {
    TB2_REQUIRE(weight && bias && z && h, "null argument");
    TB2_REQUIRE(M >= 0 && H > 0 && latent_dim > 0, "bad sizes");
    if (M == 0) return TB2_OK;
    cudaStream_t st = (cudaStream_t)stream;
    {
        KernelTimer kt("vae_scale_hidden", st);
        vae_scale_hidden_kernel<<<M, 128, (size_t)latent_dim * sizeof(float), st>>>(weight, bias, z, h, M, H,
                                                                                   latent_dim);
    }
    TB2_LAUNCH_CHECK();
    return TB2_OK;
}

extern "C" int tb2_sgan_add_noise(const float* weight, const float* bias, const float* noise, float* h,
                                  int32_t M, int32_t H, int32_t noise_dim, void* stream) {
    TB2_REQUIRE(weight && bias && h && (noise || noise_dim == 0), "null argument");
    TB2_REQUIRE(M >= 0 && H > 0 && noise_dim >= 0 && noise_dim < H, "bad sizes");
    if (M == 0) return TB2_OK;
    cudaStream_t st = (cudaStream_t)stream;
    {
        KernelTimer kt("sgan_add_noise", st);
        sgan_add_noise_kernel<<<M, 128, (size_t)H * sizeof(float), st>>>(weight, bias, noise, h, M, H, noise_dim);
    }
    TB2_LAUNCH_CHECK();
    return TB2_OK;
}
