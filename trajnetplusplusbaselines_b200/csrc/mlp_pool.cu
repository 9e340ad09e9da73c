// HiddenStateMLPPooling on the device (--type hiddenstatemlp, the Social-GAN pooling module).
//
//   pooled_i = max over ALL tracks j of the scene (j = i included) of
//                [ relu(Ws (pos_j - pos_i) + bs) | relu(Wh h_j + bh) | relu(Wv 4 (v_j - v_i) + bv) ]
//              with -100 where an input of the embedding is NaN,
//   out_i    = Wo pooled_i + bo
//   (reference: trajnetbaselines/lstm/non_gridbased_pooling.py:49-58 embed_with_masking, :150-239)
//
// One CTA per scene: the pair terms are 2-input Linears (N^2 x 64 x 2 FMAs), the hidden embedding and the output
// projection are N x 128 x {64, out_dim} -- a few hundred kFLOP per scene, FP32 FFMA, no tensor cores.  The
// reference materialises [B, N, N, 128] tensors and three masked scatters per step.
#include <math_constants.h>

#include "common.cuh"

namespace tb2 {

constexpr int kMpThreads = 256;

struct MlpPoolParams {
    const float2* obs1;
    const float2* obs2;
    const float* hidden;      // [M, H]
    const int* scene_off;
    const float* Ws;          // [ds, 2]
    const float* bs;
    const float* Wv;          // [dv, 2]
    const float* bv;
    const float* WhT;         // [H, dh]
    const float* bh;
    const float* WoT;         // [ds + dh + dv, out_dim]
    const float* bo;
    float* out;               // [M, out_dim]
    int H, ds, dv, dh, out_dim;
};

__global__ void __launch_bounds__(kMpThreads) hidden_mlp_pool_kernel(MlpPoolParams p) {
    extern __shared__ __align__(16) float smem_mp[];
    const int scene = blockIdx.x;
    const int row0 = p.scene_off[scene];
    const int n = p.scene_off[scene + 1] - row0;
    const int D = p.ds + p.dh + p.dv;
    float2* pos = reinterpret_cast<float2*>(smem_mp);              // [n] obs2 (NaN kept)
    float2* vel = pos + n;                                         // [n] obs2 - obs1 (NaN if either is)
    float* hemb = reinterpret_cast<float*>(vel + n);               // [n][dh]
    float* hmax = hemb + (size_t)n * p.dh;                         // [dh]
    float* pooled = hmax + p.dh;                                   // [n][D]
    const int tid = threadIdx.x;
    grid_dep_wait();
    grid_dep_launch();
    for (int j = tid; j < n; j += kMpThreads) {
        const float2 a = p.obs1[row0 + j], b = p.obs2[row0 + j];
        pos[j] = b;
        vel[j] = make_float2(b.x - a.x, b.y - a.y);
    }
    // hidden embedding of every track (a row with a NaN is masked to -100)
    for (int idx = tid; idx < n * p.dh; idx += kMpThreads) {
        const int j = idx / p.dh, k = idx - j * p.dh;
        const float* h = p.hidden + (size_t)(row0 + j) * p.H;
        float acc = 0.f;
        bool bad = false;
        for (int c = 0; c < p.H; ++c) {
            const float hv = __ldg(h + c);
            bad |= isnan(hv);
            acc = fmaf(hv, __ldg(p.WhT + (size_t)c * p.dh + k), acc);
        }
        hemb[idx] = bad ? -100.f : fmaxf(acc + p.bh[k], 0.f);
    }
    __syncthreads();
    for (int k = tid; k < p.dh; k += kMpThreads) {
        float m = -CUDART_INF_F;
        for (int j = 0; j < n; ++j) m = fmaxf(m, hemb[j * p.dh + k]);
        hmax[k] = m;
    }
    __syncthreads();
    // pooled[i] = [spatial | hidden | velocity]
    for (int idx = tid; idx < n * D; idx += kMpThreads) {
        const int i = idx / D, k = idx - i * D;
        float m;
        if (k >= p.ds && k < p.ds + p.dh) {
            m = hmax[k - p.ds];
        } else {
            const bool sp = k < p.ds;
            const int kk = sp ? k : k - p.ds - p.dh;
            const float w0 = sp ? p.Ws[2 * kk] : p.Wv[2 * kk], w1 = sp ? p.Ws[2 * kk + 1] : p.Wv[2 * kk + 1];
            const float b = sp ? p.bs[kk] : p.bv[kk];
            const float2 ci = sp ? pos[i] : vel[i];
            const float scale = sp ? 1.f : 4.f;
            m = -CUDART_INF_F;
            for (int j = 0; j < n; ++j) {
                const float2 cj = sp ? pos[j] : vel[j];
                const float rx = (cj.x - ci.x) * scale, ry = (cj.y - ci.y) * scale;
                float e = -100.f;                                  // embed_with_masking fill value
                if (!(isnan(rx) || isnan(ry))) e = fmaxf(fmaf(ry, w1, fmaf(rx, w0, b)), 0.f);
                m = fmaxf(m, e);
            }
        }
        pooled[idx] = m;
    }
    __syncthreads();
    for (int idx = tid; idx < n * p.out_dim; idx += kMpThreads) {
        const int i = idx / p.out_dim, o = idx - i * p.out_dim;
        const float* pi = pooled + (size_t)i * D;
        float a0 = 0.f, a1 = 0.f;
        int k = 0;
        for (; k + 1 < D; k += 2) {
            a0 = fmaf(pi[k], __ldg(p.WoT + (size_t)k * p.out_dim + o), a0);
            a1 = fmaf(pi[k + 1], __ldg(p.WoT + (size_t)(k + 1) * p.out_dim + o), a1);
        }
        if (k < D) a0 = fmaf(pi[k], __ldg(p.WoT + (size_t)k * p.out_dim + o), a0);
        p.out[(size_t)(row0 + i) * p.out_dim + o] = (a0 + a1) + p.bo[o];
    }
}

int launch_hidden_mlp_pool(const tb2_lstm* m, const tb2_layout* l, const float* hidden, const float* obs1,
                           const float* obs2, float* out, cudaStream_t st) {
    MlpPoolParams p;
    p.obs1 = (const float2*)obs1;
    p.obs2 = (const float2*)obs2;
    p.hidden = hidden;
    p.scene_off = l->scene_off;
    p.Ws = m->mp_Ws; p.bs = m->mp_bs; p.Wv = m->mp_Wv; p.bv = m->mp_bv;
    p.WhT = m->mp_WhT; p.bh = m->mp_bh; p.WoT = m->mp_WoT; p.bo = m->mp_bo;
    p.out = out;
    p.H = m->H;
    p.ds = m->cfg.mlp_dim_spatial; p.dv = m->cfg.mlp_dim_vel; p.dh = m->cfg.mlp_dim_hidden;
    p.out_dim = m->pool_out;
    const int D = p.ds + p.dh + p.dv;
    const size_t smem = ((size_t)l->n_max * (4 + p.dh + D) + p.dh) * sizeof(float) + 16;
    TB2_REQUIRE(smem <= 200 * 1024, "scene too large for the hidden-state MLP pooling kernel");
    static DynSmemConfig configured;
    TB2_CHECK_CUDA(configured.ensure(hidden_mlp_pool_kernel, smem, 48 * 1024));
    {
        KernelTimer kt("hidden_mlp_pool", st);
        launch_pdl(hidden_mlp_pool_kernel, dim3(l->B), dim3(kMpThreads), smem, st, p);
    }
    TB2_LAUNCH_CHECK();
    return TB2_OK;
}

// ------------------------------------------------------------------------------------------
// NearestNeighborMLP on the device (--type nn, reference non_gridbased_pooling.py:64-147).
//
//   for every track i: the n nearest other tracks of its scene by ||pos_j - pos_i|| (absent tracks count as 1000 m,
//   reference :131-132), in ascending distance; per kept neighbour the features [pos_j - pos_i | v_j - v_i] (NaN -> 0,
//   :141; missing neighbours are zero rows, :134-136) go through Linear(2 or 4 -> out_dim / n) + ReLU; the n
//   embeddings are concatenated.  One warp per track: lane = candidate neighbour, n rounds of a warp arg-min.
// ------------------------------------------------------------------------------------------
struct NnPoolParams {
    const float2* obs1;
    const float2* obs2;
    const int* scene_off;
    const float* W;            // [d, 2 or 4]
    const float* b;            // [d]
    float* out;                // [M, n * d]
    int n, d, with_vel;
};

__global__ void __launch_bounds__(256) nn_mlp_pool_kernel(NnPoolParams p) {
    extern __shared__ __align__(16) float smem_nn[];
    const int scene = blockIdx.x;
    const int row0 = p.scene_off[scene];
    const int ns = p.scene_off[scene + 1] - row0;
    float2* pos = reinterpret_cast<float2*>(smem_nn);              // [ns] obs2 (NaN kept)
    float2* vel = pos + ns;                                        // [ns] obs2 - obs1
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    grid_dep_wait();
    grid_dep_launch();
    for (int j = tid; j < ns; j += blockDim.x) {
        const float2 a = p.obs1[row0 + j], b = p.obs2[row0 + j];
        pos[j] = b;
        vel[j] = make_float2(b.x - a.x, b.y - a.y);
    }
    __syncthreads();
    for (int i = warp; i < ns; i += nwarps) {
        const float2 pi = pos[i], vi = vel[i];
        float* out = p.out + (size_t)(row0 + i) * p.n * p.d;
        float last_d = -1.f;
        int last_j = -1;
        for (int k = 0; k < p.n; ++k) {
            // arg-min over the candidates after (last_d, last_j) in (distance, index) order: selection without marking
            float best = CUDART_INF_F;
            int best_j = 0x7fffffff;
            for (int j = lane; j < ns; j += 32) {
                if (j == i) continue;
                const float rx = pos[j].x - pi.x, ry = pos[j].y - pi.y;
                float dist = sqrtf(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)));
                if (isnan(dist)) dist = 1000.f;
                const bool after = dist > last_d || (dist == last_d && j > last_j);
                if (after && (dist < best || (dist == best && j < best_j))) { best = dist; best_j = j; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, off);
                const int oj = __shfl_xor_sync(0xffffffffu, best_j, off);
                if (ob < best || (ob == best && oj < best_j)) { best = ob; best_j = oj; }
            }
            float f[4] = {0.f, 0.f, 0.f, 0.f};                     // fewer than n other tracks: zero features
            if (best_j != 0x7fffffff) {
                last_d = best; last_j = best_j;
                const float2 pj = pos[best_j], vj = vel[best_j];
                f[0] = pj.x - pi.x; f[1] = pj.y - pi.y; f[2] = vj.x - vi.x; f[3] = vj.y - vi.y;
#pragma unroll
                for (int c = 0; c < 4; ++c) f[c] = isnan(f[c]) ? 0.f : (isinf(f[c]) ? copysignf(3.402823466e+38f, f[c]) : f[c]);
            } else {
                last_d = CUDART_INF_F;
            }
            const int in_dim = p.with_vel ? 4 : 2;
            for (int o = lane; o < p.d; o += 32) {
                float acc = p.b[o];
                for (int c = 0; c < in_dim; ++c) acc = fmaf(f[c], p.W[o * in_dim + c], acc);
                out[k * p.d + o] = fmaxf(acc, 0.f);
            }
        }
    }
}

int launch_nn_mlp_pool(const tb2_lstm* m, const tb2_layout* l, const float* obs1, const float* obs2, float* out,
                       cudaStream_t st) {
    NnPoolParams p;
    p.obs1 = (const float2*)obs1;
    p.obs2 = (const float2*)obs2;
    p.scene_off = l->scene_off;
    p.W = m->mp_Ws;
    p.b = m->mp_bs;
    p.out = out;
    p.n = m->cfg.n;
    p.d = m->cfg.mlp_dim_spatial;
    p.with_vel = m->cfg.mlp_dim_vel != 0;
    const size_t smem = (size_t)l->n_max * 4 * sizeof(float) + 16;
    TB2_REQUIRE(smem <= 48 * 1024, "scene too large for the nearest-neighbour pooling kernel");
    {
        KernelTimer kt("nn_mlp_pool", st);
        launch_pdl(nn_mlp_pool_kernel, dim3(l->B), dim3(256), smem, st, p);
    }
    TB2_LAUNCH_CHECK();
    return TB2_OK;
}

}  // namespace tb2
