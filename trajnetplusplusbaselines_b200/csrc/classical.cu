// Classical crowd simulators: social force and ORCA, one persistent kernel each.
//
// The reference builds ONE simulator per scene and crosses Python -> third-party code every
// step (classical/socialforce.py:89-95: 96 x socialforce.Simulator.step(); classical/orca.py:
// 99-119: 97 x rvo2 doStep() + 3 FFI calls per agent per step).  Here one CTA owns a scene for
// the whole rollout: the state lives in shared memory / registers for all steps, every
// pedestrian of the scene is a thread, the scenes of a batch run in lockstep in one launch,
// and only the sampled positions are written to HBM.
//
// Arithmetic follows the un-vendored upstream packages (socialforce v0.1.x: float64; RVO2
// v2.0.x: float) as restated in oracle/classical_oracle.py and oracle/orca_oracle.c -- parity
// vs upstream is UNPINNED (see those headers).  This file is compiled with -fmad=false so the
// CUDA result is comparable operation by operation with the CPU restatement.
#include <math_constants.h>

#include "common.cuh"

namespace tb2 {

// =========================================================================================
// social force (double precision like upstream numpy)
// =========================================================================================
struct SfScene {
    double* px; double* py; double* vx; double* vy; double* ex; double* ey; double* sp;
};

__device__ __forceinline__ double sf_potential(double rx, double ry, double sb, double ebx, double eby,
                                               double dt, double v0, double sigma) {
    // V(r_ab) = v0 exp(-b / sigma), b = 0.5 sqrt((|r| + |r - dt s_b e_b|)^2 - (dt s_b)^2)
    const double n1 = sqrt(rx * rx + ry * ry);
    const double qx = rx - dt * sb * ebx, qy = ry - dt * sb * eby;
    const double n2 = sqrt(qx * qx + qy * qy);
    const double s = n1 + n2;
    const double in_sqrt = s * s - (dt * sb) * (dt * sb);
    const double b = 0.5 * sqrt(in_sqrt);
    return v0 * exp(-b / sigma);
}

// kWarpScenes: every scene has at most 32 pedestrians -> one WARP per scene, 4 scenes per CTA, __syncwarp instead
// of __syncthreads (a CTA of one warp caps an SM at 32 resident warps; the arithmetic per pedestrian is unchanged,
// results are bit-identical to the one-scene-per-CTA form).
constexpr int kScenesPerCta = 4;

template <bool kWarpScenes>
__device__ __forceinline__ void scene_sync() {
    if (kWarpScenes) __syncwarp();
    else __syncthreads();
}

template <bool kWarpScenes>
__global__ void sf_simulate_kernel(const int* __restrict__ scene_off, const double* __restrict__ state,
                                   double* __restrict__ out, int A, int B, int n_max, tb2_sf_params p) {
    extern __shared__ double smem_sf[];
    const int scene = kWarpScenes ? blockIdx.x * kScenesPerCta + (int)(threadIdx.x >> 5) : (int)blockIdx.x;
    if (scene >= B) return;                                  // whole warp (kWarpScenes): no block-wide barrier below
    const int row0 = scene_off[scene];
    const int n = scene_off[scene + 1] - row0;
    const int a = kWarpScenes ? (int)(threadIdx.x & 31) : (int)threadIdx.x;
    double* px = smem_sf + (kWarpScenes ? (size_t)(threadIdx.x >> 5) * n_max * 7 : 0);
    double* py = px + n;
    double* vx = py + n;
    double* vy = vx + n;
    double* ex = vy + n;
    double* ey = ex + n;
    double* sp = ey + n;

    const double dt = (double)p.delta_t, tau = (double)p.tau, v0 = (double)p.v0, sigma = (double)p.sigma;
    const double fd = 1e-3;
    const double cosphi = cos(200.0 / 2.0 / 180.0 * 3.141592653589793);   // FieldOfView(twophi=200)
    double x = 0, y = 0, ux = 0, uy = 0, dx = 0, dy = 0, s0 = 0, smax = 0;
    if (a < n) {
        const double* st = state + (size_t)(row0 + a) * 6;
        x = st[0]; y = st[1]; ux = st[2]; uy = st[3]; dx = st[4]; dy = st[5];
        s0 = sqrt(ux * ux + uy * uy);            // Simulator.__init__: initial_speeds
        smax = 1.3 * s0;                         // max_speeds
    }
    int sample = 0;
    for (int k = 0; k < p.n_steps; ++k) {
        double eax = 0, eay = 0;
        if (a < n) {
            const double gx = dx - x, gy = dy - y;
            const double gn = sqrt(gx * gx + gy * gy);
            eax = gx / gn; eay = gy / gn;        // desired direction (NaN at the destination, as upstream)
            px[a] = x; py[a] = y; vx[a] = ux; vy[a] = uy; ex[a] = eax; ey[a] = eay;
            sp[a] = sqrt(ux * ux + uy * uy);
        }
        scene_sync<kWarpScenes>();
        if (a < n) {
            double Fx = 1.0 / tau * (s0 * eax - ux);
            double Fy = 1.0 / tau * (s0 * eay - uy);
            double sumx = 0.0, sumy = 0.0;
            for (int b = 0; b < n; ++b) {
                double fx = 0.0, fy = 0.0, w = 0.0;
                if (b != a) {
                    const double rx = x - px[b], ry = y - py[b];
                    const double sb = sp[b], ebx = ex[b], eby = ey[b];
                    const double v = sf_potential(rx, ry, sb, ebx, eby, dt, v0, sigma);
                    const double dvdx = (sf_potential(rx + fd, ry, sb, ebx, eby, dt, v0, sigma) - v) / fd;
                    const double dvdy = (sf_potential(rx, ry + fd, sb, ebx, eby, dt, v0, sigma) - v) / fd;
                    fx = -1.0 * dvdx; fy = -1.0 * dvdy;               // f_ab = -grad V
                    const double gx = -fx, gy = -fy;                  // w(e, -f_ab)
                    const bool in_sight = (eax * gx + eay * gy) > sqrt(gx * gx + gy * gy) * cosphi;
                    w = in_sight ? 1.0 : 0.5;
                }
                sumx += w * fx;
                sumy += w * fy;
            }
            Fx += sumx; Fy += sumy;
            const double wx = ux + dt * Fx, wy = uy + dt * Fy;
            const double wn = sqrt(wx * wx + wy * wy);
            const double q = smax / wn;
            const double factor = isnan(q) ? q : (q < 1.0 ? q : 1.0);       // numpy.minimum(1, q)
            ux = wx * factor; uy = wy * factor;
            x = x + ux * dt; y = y + uy * dt;
            if (k % p.sample_every == 0) {
                double* o = out + ((size_t)sample * A + row0 + a) * 2;
                o[0] = x; o[1] = y;
            }
        }
        if (k % p.sample_every == 0) ++sample;
        scene_sync<kWarpScenes>();
    }
}

// =========================================================================================
// ORCA (float like RVO2)
// =========================================================================================
constexpr int kOrcaMaxNeigh = 16;
constexpr float kOrcaEps = 0.00001f;

struct Line { float2 point, dir; };

__device__ __forceinline__ float2 f2(float x, float y) { return make_float2(x, y); }
__device__ __forceinline__ float2 vadd(float2 a, float2 b) { return f2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 vsub(float2 a, float2 b) { return f2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 vmul(float s, float2 a) { return f2(s * a.x, s * a.y); }
__device__ __forceinline__ float vdot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float vdet(float2 a, float2 b) { return a.x * b.y - a.y * b.x; }
__device__ __forceinline__ float vabssq(float2 a) { return vdot(a, a); }
__device__ __forceinline__ float2 vnormalize(float2 a) {
    float l = sqrtf(vabssq(a));
    return f2(a.x / l, a.y / l);
}

__device__ bool orca_lp1(const Line* lines, int line_no, float radius, float2 opt, bool dir_opt, float2& result) {
    const float dp = vdot(lines[line_no].point, lines[line_no].dir);
    const float disc = dp * dp + radius * radius - vabssq(lines[line_no].point);
    if (disc < 0.0f) return false;
    const float sq = sqrtf(disc);
    float t_left = -dp - sq, t_right = -dp + sq;
    for (int i = 0; i < line_no; ++i) {
        const float den = vdet(lines[line_no].dir, lines[i].dir);
        const float num = vdet(lines[i].dir, vsub(lines[line_no].point, lines[i].point));
        if (fabsf(den) <= kOrcaEps) {
            if (num < 0.0f) return false;
            continue;
        }
        const float t = num / den;
        if (den >= 0.0f) t_right = fminf(t_right, t); else t_left = fmaxf(t_left, t);
        if (t_left > t_right) return false;
    }
    if (dir_opt) {
        if (vdot(opt, lines[line_no].dir) > 0.0f) result = vadd(lines[line_no].point, vmul(t_right, lines[line_no].dir));
        else result = vadd(lines[line_no].point, vmul(t_left, lines[line_no].dir));
    } else {
        const float t = vdot(lines[line_no].dir, vsub(opt, lines[line_no].point));
        if (t < t_left) result = vadd(lines[line_no].point, vmul(t_left, lines[line_no].dir));
        else if (t > t_right) result = vadd(lines[line_no].point, vmul(t_right, lines[line_no].dir));
        else result = vadd(lines[line_no].point, vmul(t, lines[line_no].dir));
    }
    return true;
}

__device__ int orca_lp2(const Line* lines, int n, float radius, float2 opt, bool dir_opt, float2& result) {
    if (dir_opt) result = vmul(radius, opt);
    else if (vabssq(opt) > radius * radius) result = vmul(radius, vnormalize(opt));
    else result = opt;
    for (int i = 0; i < n; ++i) {
        if (vdet(lines[i].dir, vsub(lines[i].point, result)) > 0.0f) {
            const float2 tmp = result;
            if (!orca_lp1(lines, i, radius, opt, dir_opt, result)) {
                result = tmp;
                return i;
            }
        }
    }
    return n;
}

__device__ void orca_lp3(const Line* lines, int n, int begin, float radius, float2& result) {
    float distance = 0.0f;
    Line proj[kOrcaMaxNeigh];
    for (int i = begin; i < n; ++i) {
        if (vdet(lines[i].dir, vsub(lines[i].point, result)) > distance) {
            int np = 0;
            for (int j = 0; j < i; ++j) {
                Line l;
                const float d = vdet(lines[i].dir, lines[j].dir);
                if (fabsf(d) <= kOrcaEps) {
                    if (vdot(lines[i].dir, lines[j].dir) > 0.0f) continue;
                    l.point = vmul(0.5f, vadd(lines[i].point, lines[j].point));
                } else {
                    l.point = vadd(lines[i].point,
                                   vmul(vdet(lines[j].dir, vsub(lines[i].point, lines[j].point)) / d, lines[i].dir));
                }
                l.dir = vnormalize(vsub(lines[j].dir, lines[i].dir));
                proj[np++] = l;
            }
            const float2 tmp = result;
            if (orca_lp2(proj, np, radius, f2(-lines[i].dir.y, lines[i].dir.x), true, result) < np) result = tmp;
            distance = vdet(lines[i].dir, vsub(lines[i].point, result));
        }
    }
}

template <bool kWarpScenes>
__global__ void orca_simulate_kernel(const int* __restrict__ scene_off, const float2* __restrict__ pos_in,
                                     const float2* __restrict__ vel_in, const double2* __restrict__ goal_in,
                                     const double* __restrict__ speed_in, float2* __restrict__ out, int A, int B,
                                     int n_max, tb2_orca_params p) {
    extern __shared__ float2 smem_orca[];
    const int scene = kWarpScenes ? blockIdx.x * kScenesPerCta + (int)(threadIdx.x >> 5) : (int)blockIdx.x;
    if (scene >= B) return;
    const int row0 = scene_off[scene];
    const int n = scene_off[scene + 1] - row0;
    const int a = kWarpScenes ? (int)(threadIdx.x & 31) : (int)threadIdx.x;
    float2* pos = smem_orca + (kWarpScenes ? (size_t)(threadIdx.x >> 5) * n_max * 2 : 0);
    float2* vel = pos + n;

    float2 mypos = f2(0.f, 0.f), myvel = f2(0.f, 0.f), pref = f2(0.f, 0.f);
    double2 goal = make_double2(0.0, 0.0);
    double speed = 0.0;
    float maxsp = 0.f;
    if (a < n) {
        mypos = pos_in[row0 + a];
        myvel = vel_in[row0 + a];
        goal = goal_in[row0 + a];
        speed = speed_in[row0 + a];
        maxsp = (float)(1.3 * speed);               // orca.py:36,55  MAX_SPEED_MULTIPLIER
        pos[a] = mypos;
        vel[a] = myvel;
    }
    scene_sync<kWarpScenes>();
    const float inv_th = 1.0f / p.time_horizon;
    const float inv_ts = 1.0f / p.time_step;
    const float cr = p.radius + p.radius;
    const float crsq = cr * cr;
    const int max_nb = p.max_neighbors;
    int sample = 0;
    for (int count = 1; count <= p.n_steps; ++count) {
        float2 newv = myvel;
        if (a < n) {
            int nb[kOrcaMaxNeigh];
            float nd[kOrcaMaxNeigh];
            int nn = 0;
            float range_sq = p.neighbor_dist * p.neighbor_dist;
            for (int b = 0; b < n; ++b) {
                if (b == a) continue;
                const float dsq = vabssq(vsub(mypos, pos[b]));
                if (dsq < range_sq) {
                    if (nn < max_nb) { nb[nn] = b; nd[nn] = dsq; ++nn; }
                    int i = nn - 1;
                    while (i != 0 && dsq < nd[i - 1]) { nb[i] = nb[i - 1]; nd[i] = nd[i - 1]; --i; }
                    nb[i] = b; nd[i] = dsq;
                    if (nn == max_nb) range_sq = nd[nn - 1];
                }
            }
            Line lines[kOrcaMaxNeigh];
            for (int k = 0; k < nn; ++k) {
                const int b = nb[k];
                const float2 rp = vsub(pos[b], mypos);
                const float2 rv = vsub(myvel, vel[b]);
                const float dsq = vabssq(rp);
                Line l;
                float2 u;
                if (dsq > crsq) {
                    const float2 w = vsub(rv, vmul(inv_th, rp));
                    const float wsq = vabssq(w);
                    const float dp1 = vdot(w, rp);
                    if (dp1 < 0.0f && dp1 * dp1 > crsq * wsq) {
                        const float wl = sqrtf(wsq);
                        const float2 uw = f2(w.x / wl, w.y / wl);
                        l.dir = f2(uw.y, -uw.x);
                        u = vmul(cr * inv_th - wl, uw);
                    } else {
                        const float leg = sqrtf(dsq - crsq);
                        if (vdet(rp, w) > 0.0f) {
                            l.dir = f2((rp.x * leg - rp.y * cr) / dsq, (rp.x * cr + rp.y * leg) / dsq);
                        } else {
                            l.dir = f2(-(rp.x * leg + rp.y * cr) / dsq, -(-rp.x * cr + rp.y * leg) / dsq);
                        }
                        const float dp2 = vdot(rv, l.dir);
                        u = vsub(vmul(dp2, l.dir), rv);
                    }
                } else {
                    const float2 w = vsub(rv, vmul(inv_ts, rp));
                    const float wl = sqrtf(vabssq(w));
                    const float2 uw = f2(w.x / wl, w.y / wl);
                    l.dir = f2(uw.y, -uw.x);
                    u = vmul(cr * inv_ts - wl, uw);
                }
                l.point = vadd(myvel, vmul(0.5f, u));
                lines[k] = l;
            }
            float2 res;
            const int fail = orca_lp2(lines, nn, maxsp, pref, false, res);
            if (fail < nn) orca_lp3(lines, nn, fail, maxsp, res);
            newv = res;
        }
        scene_sync<kWarpScenes>();                       // every agent has read the old positions / velocities
        if (a < n) {
            myvel = newv;
            mypos = vadd(mypos, vmul(p.time_step, myvel));
            pos[a] = mypos;
            vel[a] = myvel;
            if (count % p.sample_every == 0) out[(size_t)sample * A + row0 + a] = mypos;
            // orca.py:111-119 (double, like the reference's numpy code)
            const double gx = goal.x - (double)mypos.x, gy = goal.y - (double)mypos.y;
            const double dist = sqrt(gx * gx + gy * gy);
            if (dist < (double)p.end_range) pref = f2(0.f, 0.f);
            else if (dist > speed) pref = f2((float)(speed * gx / dist), (float)(speed * gy / dist));
            else pref = f2((float)gx, (float)gy);
        }
        if (count % p.sample_every == 0) ++sample;
        scene_sync<kWarpScenes>();
    }
}

}  // namespace tb2

using namespace tb2;

extern "C" {

int tb2_sf_simulate(const tb2_layout* l, const tb2_sf_params* p, const double* state, double* out, void* stream) {
    TB2_REQUIRE(l && p && state && out, "null argument");
    TB2_REQUIRE(p->n_steps >= 1 && p->sample_every >= 1, "bad step counts");
    TB2_REQUIRE(l->n_max <= 1024, "scene larger than 1024 pedestrians");
    cudaStream_t st = (cudaStream_t)stream;
    int threads = (l->n_max + 31) / 32 * 32;
    size_t smem = (size_t)l->n_max * 7 * sizeof(double);
    {
        KernelTimer kt("sf_simulate", st);
        if (l->n_max <= 32)
            sf_simulate_kernel<true><<<(l->B + kScenesPerCta - 1) / kScenesPerCta, 32 * kScenesPerCta, smem * kScenesPerCta, st>>>(
                l->scene_off, state, out, l->M, l->B, l->n_max, *p);
        else {
            static DynSmemConfig configured;
            TB2_CHECK_CUDA(configured.ensure(sf_simulate_kernel<false>, smem, 48 * 1024));
            sf_simulate_kernel<false><<<l->B, threads, smem, st>>>(l->scene_off, state, out, l->M, l->B, l->n_max, *p);
        }
    }
    TB2_LAUNCH_CHECK();
    return TB2_OK;
}

int tb2_orca_simulate(const tb2_layout* l, const tb2_orca_params* p, const float* pos, const float* vel,
                      const double* goal, const double* speed, float* out, void* stream) {
    TB2_REQUIRE(l && p && pos && vel && goal && speed && out, "null argument");
    TB2_REQUIRE(p->n_steps >= 1 && p->sample_every >= 1, "bad step counts");
    TB2_REQUIRE(p->max_neighbors >= 1 && p->max_neighbors <= kOrcaMaxNeigh, "max_neighbors must be in [1, 16]");
    TB2_REQUIRE(l->n_max <= 1024, "scene larger than 1024 pedestrians");
    cudaStream_t st = (cudaStream_t)stream;
    int threads = (l->n_max + 31) / 32 * 32;
    size_t smem = (size_t)l->n_max * 2 * sizeof(float2);
    {
        KernelTimer kt("orca_simulate", st);
        if (l->n_max <= 32)
            orca_simulate_kernel<true><<<(l->B + kScenesPerCta - 1) / kScenesPerCta, 32 * kScenesPerCta, smem * kScenesPerCta, st>>>(
                l->scene_off, (const float2*)pos, (const float2*)vel, (const double2*)goal, speed, (float2*)out, l->M, l->B,
                l->n_max, *p);
        else
            orca_simulate_kernel<false><<<l->B, threads, smem, st>>>(l->scene_off, (const float2*)pos, (const float2*)vel,
                                                                    (const double2*)goal, speed, (float2*)out, l->M, l->B,
                                                                    l->n_max, *p);
    }
    TB2_LAUNCH_CHECK();
    return TB2_OK;
}

}  // extern "C"
