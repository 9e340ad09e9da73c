// Backward of the recurrence (BPTT) for training -- reference: what autograd does for
// Trainer.train_batch (trajnetbaselines/lstm/trainer.py:229-269) through LSTM.forward
// (lstm/lstm.py:170-264).
//
// Gradient structure exploited (probe-verified on the reference): fed-back
// positions are detached (lstm.py:242-250), and for vanilla / occupancy / directional pooling the
// pooled vector does not depend on any hidden state, so a track's gradient never leaves its own
// LSTM chain.  The caller passes the list of ACTIVE rows (tracks that receive a non-zero upstream
// gradient: the scene primaries for PredictionLoss, loss.py:57,67) and the whole backward runs
// on those R rows only: per step one recompute of the gate pre-activations (R x K x 512), the
// pointwise cell / head backward, two input-gradient GEMMs and the weight-gradient
// accumulations.  Social pooling couples the tracks of a scene through W_enc h_j and is not
// built yet (fails loudly).
//
// All accumulations into parameter gradients are deterministic (one thread owns one output
// element, steps are processed sequentially) except the sparse scatter into the first
// grid-embedding layer's weight gradient, which uses fp32 atomics like PyTorch's own index_put /
// embedding backward.
#include <math_constants.h>

#include "common.cuh"

namespace tb2 {

constexpr int kBH = 128;

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

// X_act[r] = [emb(vel) | pooled | h_prev]  (row gather), K = E + P + H
__global__ void bwd_gather_kernel(const int* __restrict__ rows, int R, const float2* __restrict__ obs1,
                                  const float2* __restrict__ obs2, const float* __restrict__ We,
                                  const float* __restrict__ be, const float* __restrict__ pooled,
                                  const float* __restrict__ h_prev, float* __restrict__ X, int* __restrict__ masked,
                                  int E, int P, int K) {
    const int r = blockIdx.x;
    if (r >= R) return;
    const int m = rows[r];
    const float2 a = obs1[m], b = obs2[m];
    const bool msk = isnan(a.x) || isnan(b.x);
    if (threadIdx.x == 0) masked[r] = msk ? 1 : 0;
    const float vx = (b.x - a.x) * 4.0f, vy = (b.y - a.y) * 4.0f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        float v = 0.f;
        if (!msk) {
            if (k < E) {
                if (k < E - 2) v = fmaxf(fmaf(We[2 * k + 1], vy, fmaf(We[2 * k], vx, be[k])), 0.f);
            } else if (k < E + P) {
                v = pooled[(size_t)m * P + (k - E)];
            } else {
                v = h_prev ? h_prev[(size_t)m * kBH + (k - E - P)] : 0.f;
            }
        }
        X[(size_t)r * K + k] = v;
    }
}

// Pointwise backward of LSTMCell + Hidden2Normal for one active row per CTA (128 threads = units).
//   in : gates_pre [R,512] (with bias), c_prev (state before the step, null = zeros), dh, dc [R,128]
//        upstream dnormal [M,5] of this step (row-indexed by track), Wn, bn
//   out: dgates [R,512], dc (in place: gradient wrt c_prev), hs [R,128] (h of this step),
//        dn_raw [R,5]; dh is overwritten with the part that by-passes the cell for masked rows
__global__ void __launch_bounds__(kBH) bwd_cell_head_kernel(
    const int* __restrict__ rows, const int* __restrict__ masked, const float* __restrict__ gates_pre,
    const float* __restrict__ c_prev, float* __restrict__ dh, float* __restrict__ dc,
    const float* __restrict__ dnormal, const float* __restrict__ Wn, const float* __restrict__ bn,
    float* __restrict__ dgates, float* __restrict__ hs, float* __restrict__ dn_raw, int R) {
    __shared__ float red[5][kBH];
    __shared__ float dn_s[5];
    const int r = blockIdx.x, u = threadIdx.x;
    const int m = rows[r];
    float* dg = dgates + (size_t)r * 4 * kBH;
    if (masked[r]) {   // absent track: state passes through, no parameter gradient (lstm.py:158-166)
#pragma unroll
        for (int g = 0; g < 4; ++g) dg[g * kBH + u] = 0.f;
        hs[(size_t)r * kBH + u] = 0.f;
        if (u < 5) dn_raw[r * 5 + u] = 0.f;
        return;        // dh, dc stay as they are
    }
    const float* gp = gates_pre + (size_t)r * 4 * kBH;
    const float ig = sigm(gp[u]), fg = sigm(gp[kBH + u]), gg = tanhf(gp[2 * kBH + u]), og = sigm(gp[3 * kBH + u]);
    const float cp = c_prev ? c_prev[(size_t)m * kBH + u] : 0.f;
    const float cn = fg * cp + ig * gg;
    const float tc = tanhf(cn);
    const float hn = og * tc;
    hs[(size_t)r * kBH + u] = hn;
    // head: n_raw = Wn h + bn (modules.py:57); recomputed for the sigmoid derivatives
#pragma unroll
    for (int o = 0; o < 5; ++o) red[o][u] = Wn[o * kBH + u] * hn;
    __syncthreads();
    for (int s = kBH / 2; s > 0; s >>= 1) {
        if (u < s) {
#pragma unroll
            for (int o = 0; o < 5; ++o) red[o][u] += red[o][u + s];
        }
        __syncthreads();
    }
    if (u < 5) {
        const float raw = red[u][0] + bn[u];
        float d = dnormal[(size_t)m * 5 + u];
        if (isnan(d)) d = 0.f;
        if (u >= 2) {
            const float sg = sigm(raw);
            d *= (u == 4 ? 0.7f : 0.2f) * sg * (1.f - sg);     // modules.py:60-62
        }
        dn_s[u] = d;
        dn_raw[r * 5 + u] = d;
    }
    __syncthreads();
    float dht = dh[(size_t)r * kBH + u];
#pragma unroll
    for (int o = 0; o < 5; ++o) dht = fmaf(Wn[o * kBH + u], dn_s[o], dht);
    float dct = dc[(size_t)r * kBH + u] + dht * og * (1.f - tc * tc);
    const float dog = dht * tc;
    const float dig = dct * gg, dgg = dct * ig, dfg = dct * cp;
    dg[u] = dig * ig * (1.f - ig);
    dg[kBH + u] = dfg * fg * (1.f - fg);
    dg[2 * kBH + u] = dgg * (1.f - gg * gg);
    dg[3 * kBH + u] = dog * og * (1.f - og);
    dc[(size_t)r * kBH + u] = dct * fg;
    dh[(size_t)r * kBH + u] = 0.f;        // the recurrent part arrives from dgates . W_hh
}

// C[n][k] += sum_r A[r][n] * B[r][k]     A [R, lda] (n < N), B [R, ldb] (k < Kc), C [N, ldc]
__global__ void __launch_bounds__(256) gemm_tn_accum_kernel(const float* __restrict__ A, int lda,
                                                            const float* __restrict__ B, int ldb,
                                                            float* __restrict__ C, int ldc, int R, int N, int Kc) {
    __shared__ float As[16][64 + 1];
    __shared__ float Bs[16][64 + 1];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n0 = blockIdx.y * 64, k0 = blockIdx.x * 64;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int r0 = 0; r0 < R; r0 += 16) {
        for (int idx = tid; idx < 16 * 64; idx += 256) {
            const int rr = idx >> 6, cc = idx & 63;
            const int r = r0 + rr;
            As[rr][cc] = (r < R && n0 + cc < N) ? A[(size_t)r * lda + n0 + cc] : 0.f;
            Bs[rr][cc] = (r < R && k0 + cc < Kc) ? B[(size_t)r * ldb + k0 + cc] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[rr][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[rr][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty * 4 + i;
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k < Kc) C[(size_t)n * ldc + k] += acc[i][j];
        }
    }
}

// out[n] (+ out2[n]) += sum_r A[r][n]
__global__ void colsum_accum_kernel(const float* __restrict__ A, int lda, int R, int N, float* __restrict__ out,
                                    float* __restrict__ out2) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float s = 0.f;
    for (int r = 0; r < R; ++r) s += A[(size_t)r * lda + n];
    out[n] += s;
    if (out2) out2[n] += s;
}

// InputEmbedding backward (modules.py:24-30): d pre[k] = dX[r][k] * (emb > 0); one thread per k
__global__ void bwd_embed_kernel(const int* __restrict__ rows, const int* __restrict__ masked, int R,
                                 const float2* __restrict__ obs1, const float2* __restrict__ obs2,
                                 const float* __restrict__ X, int ldx, const float* __restrict__ dX, int lddx,
                                 float* __restrict__ dWe, float* __restrict__ dbe, int E) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= E - 2) return;
    float gx = 0.f, gy = 0.f, gb = 0.f;
    for (int r = 0; r < R; ++r) {
        if (masked[r]) continue;
        if (X[(size_t)r * ldx + k] > 0.f) {
            const int m = rows[r];
            const float2 a = obs1[m], b = obs2[m];
            const float d = dX[(size_t)r * lddx + k];
            gx = fmaf(d, (b.x - a.x) * 4.0f, gx);
            gy = fmaf(d, (b.y - a.y) * 4.0f, gy);
            gb += d;
        }
    }
    dWe[2 * k] += gx;
    dWe[2 * k + 1] += gy;
    dbe[k] += gb;
}

// First grid-embedding layer backward (one_layer: pooled = relu(W1 grid + b1)), sparse grid:
//   dz[o] = dX[r][E + o] * (pooled > 0);  db1[o] += dz;  dW1[o][c * cells + cell] += dz[o] * val
__global__ void bwd_pool1_kernel(const int* __restrict__ rows, const int* __restrict__ masked, int R,
                                 const float* __restrict__ X, int ldx, const float* __restrict__ dX, int lddx,
                                 const int* __restrict__ win_count, const uint32_t* __restrict__ win_ent,
                                 const float* __restrict__ win_val, int nm1, int C, int cells, int E, int P,
                                 float* __restrict__ dW1, float* __restrict__ dz_out) {
    const int r = blockIdx.x;
    const int m = rows[r];
    const bool msk = masked[r] != 0;
    const int cnt = msk ? 0 : win_count[m];
    for (int o = threadIdx.x; o < P; o += blockDim.x) {
        float dz = 0.f;
        if (!msk && X[(size_t)r * ldx + E + o] > 0.f) dz = dX[(size_t)r * lddx + E + o];
        dz_out[(size_t)r * P + o] = dz;
        if (dz != 0.f) {
            float* wrow = dW1 + (size_t)o * C * cells;
            for (int e = 0; e < cnt; ++e) {
                const uint32_t ent = win_ent[(size_t)m * nm1 + e];
                const int cell = ent >> 16;
                for (int c = 0; c < C; ++c)
                    atomicAdd(&wrow[c * cells + cell], dz * win_val[((size_t)m * nm1 + e) * 2 + c]);
            }
        }
    }
}

__global__ void add_masked_passthrough_kernel(const int* __restrict__ masked, const float* __restrict__ dh_keep,
                                              float* __restrict__ dh_new, int R) {
    // masked rows: dh passes through unchanged (their dgates are zero so dh_new is zero there)
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * kBH) return;
    if (masked[idx / kBH]) dh_new[idx] = dh_keep[idx];
}

__global__ void fill_kernel(float* __restrict__ p, float v, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace tb2

using namespace tb2;

namespace tb2 {
int launch_dense_plain(const float* X, const float* WT, const float* b, float* Y, int M, int K, int N, int relu,
                       cudaStream_t st);
int resolve_step_inputs(const tb2_layout* l, const float* observed, int obs_length, const float* truth,
                        const float* positions, int s, Workspace* ws, const float** o1, const float** o2,
                        int* phase, cudaStream_t st);
}

extern "C" {

size_t tb2_lstm_backward_workspace_bytes(const tb2_lstm* m, int32_t num_active) {
    if (!m || num_active < 0) return 0;
    const size_t R = (size_t)(num_active > 0 ? num_active : 1);
    const size_t K = (size_t)m->K_gate;
    size_t f = 0;
    f += R * K;                 // X_act
    f += R * 512 * 2;           // gates_pre, dgates
    f += R * K;                 // dX (input part, width E + P; padded to K)
    f += R * 128 * 4;           // dh, dh_new, dc, hs
    f += R * 8;                 // dn_raw (5, padded)
    f += R * (size_t)(m->P > 0 ? m->P : 1);   // dz
    f += 1024;                  // zero bias
    f += R;                     // masked flags (int)
    return f * sizeof(float) + 4096;
}

int tb2_lstm_sequence_backward(const tb2_lstm* m, const tb2_layout* l, const tb2_lstm_weights* w,
                               const float* observed, int32_t obs_length, const float* truth, int32_t n_decode,
                               const float* positions, const float* states, const float* d_normals,
                               const int32_t* active_rows, int32_t num_active, const tb2_lstm_grads* g,
                               void* workspace, size_t workspace_bytes, void* bwd_workspace,
                               size_t bwd_workspace_bytes, void* stream) {
    TB2_REQUIRE(m && l && w && g, "null handle");
    TB2_REQUIRE(m->weights_set, "tb2_lstm_set_weights has not been called");
    TB2_REQUIRE(observed && positions && states && d_normals && active_rows, "null argument");
    TB2_REQUIRE(obs_length >= 2 && n_decode >= 0, "need obs_length >= 2 and n_decode >= 0");
    TB2_REQUIRE(m->H == kBH, "hidden_dim must be 128");
    if (m->cfg.pool_type == TB2_POOL_SOCIAL) {
        set_error("training backward through social pooling (hidden-state scatter) is not built yet");
        return TB2_ERR_UNSUPPORTED;
    }
    if (m->cfg.pool_type != TB2_POOL_NONE &&
        (m->n_mlp != 1 || !m->cfg.pool_to_input || m->cfg.constant != 0.f)) {
        set_error("training backward supports one_layer grid embeddings with constant = 0 and pool_to_input");
        return TB2_ERR_UNSUPPORTED;
    }
    TB2_REQUIRE(workspace && workspace_bytes >= carve_workspace(m, l, nullptr, nullptr), "workspace too small");
    TB2_REQUIRE(bwd_workspace && bwd_workspace_bytes >= tb2_lstm_backward_workspace_bytes(m, num_active),
                "backward workspace too small");
    if (num_active == 0) return TB2_OK;
    cudaStream_t st = (cudaStream_t)stream;
    Workspace ws;
    carve_workspace(m, l, workspace, &ws);
    const int R = num_active, K = m->K_gate, E = m->E, P = m->P, EP = E + P;
    const size_t M = (size_t)l->M;
    float* f = (float*)bwd_workspace;
    float* X = f;              f += (size_t)R * K;
    float* gates_pre = f;      f += (size_t)R * 512;
    float* dgates = f;         f += (size_t)R * 512;
    float* dX = f;             f += (size_t)R * K;
    float* dh = f;             f += (size_t)R * 128;
    float* dh_new = f;         f += (size_t)R * 128;
    float* dc = f;             f += (size_t)R * 128;
    float* hs = f;             f += (size_t)R * 128;
    float* dn_raw = f;         f += (size_t)R * 8;
    float* dz = f;             f += (size_t)R * (size_t)(P > 0 ? P : 1);
    float* zero_bias = f;      f += 1024;
    int* masked = (int*)f;
    TB2_CHECK_CUDA(cudaMemsetAsync(dh, 0, (size_t)R * 128 * sizeof(float), st));
    TB2_CHECK_CUDA(cudaMemsetAsync(dc, 0, (size_t)R * 128 * sizeof(float), st));
    TB2_CHECK_CUDA(cudaMemsetAsync(zero_bias, 0, 1024 * sizeof(float), st));
    const int S = obs_length - 1 + n_decode;
    const int nm1 = l->n_max > 1 ? l->n_max - 1 : 1;
    int rc;
    for (int s = S - 1; s >= 0; --s) {
        const float *o1, *o2;
        int phase;
        if ((rc = resolve_step_inputs(l, observed, obs_length, truth, positions, s, &ws, &o1, &o2, &phase, st))) return rc;
        const float* h_prev = s > 0 ? states + ((size_t)(s - 1) * 2 + 0) * M * kBH : nullptr;
        const float* c_prev = s > 0 ? states + ((size_t)(s - 1) * 2 + 1) * M * kBH : nullptr;
        if (m->cfg.pool_type != TB2_POOL_NONE) {
            if ((rc = launch_pool_prepare(m, l, h_prev, o1, o2, 1, 0, 0, &ws, st))) return rc;
            if ((rc = launch_pool_mlp(m, l, &ws, ws.pooled, nullptr, nullptr, st))) return rc;
        }
        bwd_gather_kernel<<<R, 128, 0, st>>>(active_rows, R, (const float2*)o1, (const float2*)o2, m->We, m->be,
                                             ws.pooled, h_prev, X, masked, E, P, K);
        TB2_LAUNCH_CHECK();
        if ((rc = launch_dense_plain(X, m->WgT[phase], m->bg[phase], gates_pre, R, K, 512, 0, st))) return rc;
        bwd_cell_head_kernel<<<R, kBH, 0, st>>>(active_rows, masked, gates_pre, c_prev, dh, dc,
                                                d_normals + (size_t)s * M * 5, m->Wn, m->bn, dgates, hs, dn_raw, R);
        TB2_LAUNCH_CHECK();
        // weight gradients
        float* gWih = phase == TB2_PHASE_ENCODER ? g->encoder_weight_ih : g->decoder_weight_ih;
        float* gWhh = phase == TB2_PHASE_ENCODER ? g->encoder_weight_hh : g->decoder_weight_hh;
        float* gbih = phase == TB2_PHASE_ENCODER ? g->encoder_bias_ih : g->decoder_bias_ih;
        float* gbhh = phase == TB2_PHASE_ENCODER ? g->encoder_bias_hh : g->decoder_bias_hh;
        const float* Wih = phase == TB2_PHASE_ENCODER ? w->encoder_weight_ih : w->decoder_weight_ih;
        const float* Whh = phase == TB2_PHASE_ENCODER ? w->encoder_weight_hh : w->decoder_weight_hh;
        gemm_tn_accum_kernel<<<dim3((EP + 63) / 64, 8), 256, 0, st>>>(dgates, 512, X, K, gWih, EP, R, 512, EP);
        TB2_LAUNCH_CHECK();
        gemm_tn_accum_kernel<<<dim3(2, 8), 256, 0, st>>>(dgates, 512, X + EP, K, gWhh, 128, R, 512, 128);
        TB2_LAUNCH_CHECK();
        colsum_accum_kernel<<<2, 256, 0, st>>>(dgates, 512, R, 512, gbih, gbhh);
        TB2_LAUNCH_CHECK();
        gemm_tn_accum_kernel<<<dim3(2, 1), 256, 0, st>>>(dn_raw, 5, hs, 128, g->hidden2normal_weight, 128, R, 5, 128);
        TB2_LAUNCH_CHECK();
        colsum_accum_kernel<<<1, 32, 0, st>>>(dn_raw, 5, R, 5, g->hidden2normal_bias, nullptr);
        TB2_LAUNCH_CHECK();
        // input gradients: dX = dgates . W_ih  (torch layout [512, E+P] read as WT[k = gate][n])
        if ((rc = launch_dense_plain(dgates, Wih, zero_bias, dX, R, 512, EP, 0, st))) return rc;
        if ((rc = launch_dense_plain(dgates, Whh, zero_bias, dh_new, R, 512, 128, 0, st))) return rc;
        add_masked_passthrough_kernel<<<(R * kBH + 255) / 256, 256, 0, st>>>(masked, dh, dh_new, R);
        TB2_LAUNCH_CHECK();
        bwd_embed_kernel<<<1, 64, 0, st>>>(active_rows, masked, R, (const float2*)o1, (const float2*)o2, X, K, dX, EP,
                                           g->input_embedding_weight, g->input_embedding_bias, E);
        TB2_LAUNCH_CHECK();
        if (m->cfg.pool_type != TB2_POOL_NONE) {
            bwd_pool1_kernel<<<R, 256, 0, st>>>(active_rows, masked, R, X, K, dX, EP, ws.win_count, ws.win_ent,
                                                ws.win_val, nm1, m->C, m->cells, E, P, g->pool_embedding_weight0, dz);
            TB2_LAUNCH_CHECK();
            colsum_accum_kernel<<<(P + 255) / 256, 256, 0, st>>>(dz, P, R, P, g->pool_embedding_bias0, nullptr);
            TB2_LAUNCH_CHECK();
        }
        std::swap(dh, dh_new);
    }
    return TB2_OK;
}

}  // extern "C"
