// Kalman predictor (CPU, float64) -- BASELINE configs[0] is CPU-only, so this entry point is
// host C++ behind the same C ABI.
//
// Replaces the pykalman calls of trajnetbaselines/classical/kalman.py:40-60:
//   KalmanFilter(A = constant-velocity 4x4, C = 2x4, Q = 1e-5 I, R = 0.05^2 I, mu0 = (x0,0,y0,0))
//   .em(obs)            default em_vars: Q, R, mu0, Sigma0; 10 iterations
//   .smooth(obs)        RTS smoother, last smoothed mean = initial_state of the rollout
//   mean of 5 x .sample(n_predict + 1)   -> here: the expectation C A^k x_last plus the fitted
//                       Q, R so the caller can add the reference's sampled noise.
// pykalman is not vendored (parity unpinned, see oracle/classical_oracle.py); the EM follows
// Shumway & Stoffer as pykalman documents it.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

struct M4 { double a[4][4]; };
struct V4 { double a[4]; };

M4 zero4() { M4 r; std::memset(&r, 0, sizeof(r)); return r; }
M4 eye4(double s) { M4 r = zero4(); for (int i = 0; i < 4; ++i) r.a[i][i] = s; return r; }
M4 mul(const M4& x, const M4& y) {
    M4 r = zero4();
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += x.a[i][k] * y.a[k][j]; r.a[i][j] = s; }
    return r;
}
M4 tr(const M4& x) { M4 r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.a[i][j] = x.a[j][i]; return r; }
M4 add(const M4& x, const M4& y) { M4 r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.a[i][j] = x.a[i][j] + y.a[i][j]; return r; }
M4 sub(const M4& x, const M4& y) { M4 r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.a[i][j] = x.a[i][j] - y.a[i][j]; return r; }
V4 mulv(const M4& x, const V4& v) { V4 r; for (int i = 0; i < 4; ++i) { double s = 0; for (int k = 0; k < 4; ++k) s += x.a[i][k] * v.a[k]; r.a[i] = s; } return r; }
M4 outer(const V4& u, const V4& v) { M4 r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.a[i][j] = u.a[i] * v.a[j]; return r; }

// inverse of a 4x4 by Gauss-Jordan with partial pivoting (covariances here are SPD)
bool inv4(const M4& m, M4& out) {
    double w[4][8];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { w[i][j] = m.a[i][j]; w[i][j + 4] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        for (int r = c + 1; r < 4; ++r) if (std::fabs(w[r][c]) > std::fabs(w[piv][c])) piv = r;
        if (std::fabs(w[piv][c]) < 1e-300) return false;
        if (piv != c) for (int j = 0; j < 8; ++j) std::swap(w[piv][j], w[c][j]);
        const double d = w[c][c];
        for (int j = 0; j < 8; ++j) w[c][j] /= d;
        for (int r = 0; r < 4; ++r) if (r != c) { const double f = w[r][c]; if (f != 0.0) for (int j = 0; j < 8; ++j) w[r][j] -= f * w[c][j]; }
    }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out.a[i][j] = w[i][j + 4];
    return true;
}

struct Track {
    int T;
    std::vector<V4> pm, fm, sm;
    std::vector<M4> pc, fc, sc, G;
};

// observation model C = [[1,0,0,0],[0,0,1,0]] is applied by index (state 0 -> x, state 2 -> y)
void filter(const M4& A, const M4& Q, const double R[2][2], const V4& mu0, const M4& S0,
            const double* Z, Track& t) {
    const M4 At = tr(A);
    for (int k = 0; k < t.T; ++k) {
        if (k == 0) { t.pm[k] = mu0; t.pc[k] = S0; }
        else { t.pm[k] = mulv(A, t.fm[k - 1]); t.pc[k] = add(mul(mul(A, t.fc[k - 1]), At), Q); }
        const M4& P = t.pc[k];
        // S = C P C^T + R (2x2), K = P C^T S^-1 (4x2)
        const double s00 = P.a[0][0] + R[0][0], s01 = P.a[0][2] + R[0][1];
        const double s10 = P.a[2][0] + R[1][0], s11 = P.a[2][2] + R[1][1];
        const double det = s00 * s11 - s01 * s10;
        const double i00 = s11 / det, i01 = -s01 / det, i10 = -s10 / det, i11 = s00 / det;
        double K[4][2];
        for (int i = 0; i < 4; ++i) {
            K[i][0] = P.a[i][0] * i00 + P.a[i][2] * i10;
            K[i][1] = P.a[i][0] * i01 + P.a[i][2] * i11;
        }
        const double e0 = Z[2 * k] - t.pm[k].a[0], e1 = Z[2 * k + 1] - t.pm[k].a[2];
        for (int i = 0; i < 4; ++i) t.fm[k].a[i] = t.pm[k].a[i] + K[i][0] * e0 + K[i][1] * e1;
        // fc = P - K C P
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j)
            t.fc[k].a[i][j] = P.a[i][j] - (K[i][0] * P.a[0][j] + K[i][1] * P.a[2][j]);
    }
}

bool smooth(const M4& A, Track& t) {
    const M4 At = tr(A);
    const int T = t.T;
    t.sm[T - 1] = t.fm[T - 1];
    t.sc[T - 1] = t.fc[T - 1];
    for (int k = T - 2; k >= 0; --k) {
        M4 pinv;
        if (!inv4(t.pc[k + 1], pinv)) return false;
        t.G[k] = mul(mul(t.fc[k], At), pinv);
        V4 d;
        for (int i = 0; i < 4; ++i) d.a[i] = t.sm[k + 1].a[i] - t.pm[k + 1].a[i];
        const V4 gd = mulv(t.G[k], d);
        for (int i = 0; i < 4; ++i) t.sm[k].a[i] = t.fm[k].a[i] + gd.a[i];
        t.sc[k] = add(t.fc[k], mul(mul(t.G[k], sub(t.sc[k + 1], t.pc[k + 1])), tr(t.G[k])));
    }
    return true;
}

// One track: EM (Q, R, mu0, Sigma0), smoother, expected rollout.  Returns nullptr or a static error message; tracks are
// independent, so the caller may run them on several host threads (results do not depend on the thread count).
const char* kalman_track(int tr_i, const M4& A, const M4& At, const double* obs, const int64_t* track_offsets, int32_t n_predict,
                         int32_t em_iterations, double* pred_out, double* q_out, double* r_out, double* last_state_out) {
    const int64_t o0 = track_offsets[tr_i];
    const int T = (int)(track_offsets[tr_i + 1] - o0);
    if (T < 2) return "invalid argument: a track needs at least 2 observations (kalman.py:28-29)";
    const double* Z = obs + 2 * o0;
    Track t;
    t.T = T;
    t.pm.resize(T); t.fm.resize(T); t.sm.resize(T);
    t.pc.resize(T); t.fc.resize(T); t.sc.resize(T); t.G.resize(T);
    M4 Q = eye4(1e-5);
    double R[2][2] = {{0.05 * 0.05, 0.0}, {0.0, 0.05 * 0.05}};
    V4 mu0 = {{Z[0], 0.0, Z[1], 0.0}};
    M4 S0 = eye4(1.0);
    for (int it = 0; it < em_iterations; ++it) {
        filter(A, Q, R, mu0, S0, Z, t);
        if (!smooth(A, t)) return "kalman: singular predicted covariance";
        // M-step (pykalman _em_observation_covariance / _em_transition_covariance / initial state)
        double Rn[2][2] = {{0, 0}, {0, 0}};
        for (int k = 0; k < T; ++k) {
            const double e0 = Z[2 * k] - t.sm[k].a[0], e1 = Z[2 * k + 1] - t.sm[k].a[2];
            Rn[0][0] += e0 * e0 + t.sc[k].a[0][0];
            Rn[0][1] += e0 * e1 + t.sc[k].a[0][2];
            Rn[1][0] += e1 * e0 + t.sc[k].a[2][0];
            Rn[1][1] += e1 * e1 + t.sc[k].a[2][2];
        }
        for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) R[i][j] = Rn[i][j] / T;
        M4 Qn = zero4();
        for (int k = 0; k < T - 1; ++k) {
            const V4 ax = mulv(A, t.sm[k]);
            V4 err;
            for (int i = 0; i < 4; ++i) err.a[i] = t.sm[k + 1].a[i] - ax.a[i];
            const M4 pair = mul(t.sc[k + 1], tr(t.G[k]));      // Cov(x_{k+1}, x_k | Z)
            const M4 pa = mul(pair, At);
            M4 term = add(outer(err, err), mul(mul(A, t.sc[k]), At));
            term = add(term, t.sc[k + 1]);
            term = sub(term, pa);
            term = sub(term, tr(pa));
            Qn = add(Qn, term);
        }
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) Q.a[i][j] = Qn.a[i][j] / (T - 1);
        mu0 = t.sm[0];
        S0 = t.sc[0];
    }
    filter(A, Q, R, mu0, S0, Z, t);
    if (!smooth(A, t)) return "kalman: singular predicted covariance";
    V4 x = t.sm[T - 1];
    if (last_state_out) for (int i = 0; i < 4; ++i) last_state_out[(size_t)tr_i * 4 + i] = x.a[i];
    for (int k = 0; k < n_predict; ++k) {
        x = mulv(A, x);
        pred_out[((size_t)tr_i * n_predict + k) * 2 + 0] = x.a[0];
        pred_out[((size_t)tr_i * n_predict + k) * 2 + 1] = x.a[2];
    }
    if (q_out) for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) q_out[(size_t)tr_i * 16 + i * 4 + j] = Q.a[i][j];
    if (r_out) for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) r_out[(size_t)tr_i * 4 + i * 2 + j] = R[i][j];
    return nullptr;
}

}  // namespace

extern "C" int tb2_kalman_predict(const double* obs, const int64_t* track_offsets, int32_t n_tracks,
                                  int32_t n_predict, int32_t em_iterations, double* pred_out,
                                  double* q_out, double* r_out, double* last_state_out) {
    TB2_REQUIRE(obs && track_offsets && pred_out && n_tracks >= 0 && n_predict >= 1 && em_iterations >= 0,
                "bad argument");
    M4 A = eye4(1.0);
    A.a[0][1] = 1.0;
    A.a[2][3] = 1.0;
    const M4 At = tr(A);
    // host threads over contiguous track ranges (TB2_KALMAN_THREADS overrides; small jobs stay on the calling thread)
    int threads = (int)std::thread::hardware_concurrency();
    if (const char* e = getenv("TB2_KALMAN_THREADS")) threads = atoi(e);
    threads = std::max(1, std::min(threads, n_tracks / 64));
    std::vector<const char*> errors((size_t)threads, nullptr);
    auto run = [&](int w) {
        const int lo = (int)((int64_t)n_tracks * w / threads), hi = (int)((int64_t)n_tracks * (w + 1) / threads);
        for (int i = lo; i < hi && !errors[w]; ++i)
            errors[w] = kalman_track(i, A, At, obs, track_offsets, n_predict, em_iterations, pred_out, q_out, r_out, last_state_out);
    };
    if (threads == 1) run(0);
    else {
        std::vector<std::thread> pool;
        for (int w = 1; w < threads; ++w) pool.emplace_back(run, w);
        run(0);
        for (auto& th : pool) th.join();
    }
    for (const char* e : errors)
        if (e) { tb2::set_error(e); return TB2_ERR_INVALID; }
    return TB2_OK;
}
