// velocity_device.cuh -- block-cooperative planar-flow least squares (one CTA = one frame).
// Shared by ofb_solve_velocity* (fp64 point arrays) and the fused frame-pair path (float tracks).
#pragma once
#include "math3.cuh"

#ifndef OFB_SOLVE_THREADS
#define OFB_SOLVE_THREADS 256
#endif

struct OfbSolveOut { double v[3]; double s[3]; double res; int rank; int count; };

// Per-point contribution. X=(px,py,1), beta = X x (u3 + X x w).
//   SIM  : A_i = (n.X)[X]x, b_i = beta            simulation.py:19-23
//   NODE/EXP: A_i = [X]x,   b_i = beta/(n.X)      node:34-38, evaluate_exp.py:22-26
//   MODULE: A_i = [X]x / dist_i, b_i = A_i u3 / (n.X) = [X]x u3 / (dist_i (n.X)); no gyro term (the caller passes
//           w = 0), right-hand side not scaled by a common height    optical_flow_experiments/of_module.py:141-146
struct OfbPointTerms { double wA, wb, bx, by, bz; };

__device__ __forceinline__ OfbPointTerms ofb_point_terms(int variant, double px, double py, double ux, double uy,
                                                         const double n3[3], const double w3[3], double inv_dist = 1.0)
{
    // a = X x w
    double ax = py * w3[2] - w3[1], ay = w3[0] - px * w3[2], az = px * w3[1] - py * w3[0];
    double cx = ux + ax, cy = uy + ay, cz = az;
    OfbPointTerms t;
    t.bx = py * cz - cy; t.by = cx - px * cz; t.bz = px * cy - py * cx;
    double nx = n3[0] * px + n3[1] * py + n3[2];
    if (variant == OFB_VARIANT_SIM) { t.wA = nx; t.wb = 1.0; }
    else if (variant == OFB_VARIANT_MODULE) { t.wA = inv_dist; t.wb = inv_dist / nx; }
    else { t.wA = 1.0; t.wb = 1.0 / nx; }
    return t;
}

// 1 / dist_i of the older 4-argument of.r_tilde (sensor_precision_experiments/pixhawk_pure_IMU/of_library.py:365-384,
// called at of_module.py:125): dist_i = (n.X) |X x v| / |X x u3|
__device__ __forceinline__ double ofb_module_inv_dist(double px, double py, double ux, double uy, const double n3[3],
                                                      const double v[3])
{
    const double a0 = py * v[2] - v[1], a1 = v[0] - px * v[2], a2 = px * v[1] - py * v[0];      // X x v
    const double b0 = -uy, b1 = ux, b2 = px * uy - py * ux;                                     // X x u3
    const double na = sqrt(a0 * a0 + a1 * a1 + a2 * a2), nb = sqrt(b0 * b0 + b1 * b1 + b2 * b2);
    return nb / ((n3[0] * px + n3[1] * py + n3[2]) * na);
}

__device__ __forceinline__ double ofb_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of K doubles per thread; result valid in every thread.
template <int K>
__device__ __forceinline__ void ofb_block_sum(double (&acc)[K], double* smem /* K*32 doubles */)
{
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = ofb_warp_sum(acc[k]);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < K; ++k) smem[k * 32 + warp] = acc[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double v = lane < nwarps ? smem[k * 32 + lane] : 0.0;
        acc[k] = ofb_warp_sum(v);
    }
}

// Loader concept: begin(f), end(f), load(f, i, px, py, ux, uy) -> bool (false = point skipped),
// dist(f, i, d) -> bool (MODULE variant: true = per-point distance supplied by the caller, false = derive it from the
// prior velocity `vprior` with the 4-argument r_tilde).
template <class Loader>
__device__ OfbSolveOut ofb_block_solve(const Loader& ld, int f, int variant, double d, const double* n3g,
                                       const double* w3g, const double* t3g, const double* vprior = nullptr)
{
    __shared__ double red[11 * 32];
    const bool module = variant == OFB_VARIANT_MODULE;
    double n3[3] = {n3g[0], n3g[1], n3g[2]};
    double w3[3] = {w3g[0], w3g[1], w3g[2]};
    double vp[3] = {0.0, 0.0, 0.0};
    if (module) {                                   // flow is used as given (no gyro term), rhs is not scaled by a height
        w3[0] = w3[1] = w3[2] = 0.0; d = 1.0;
        if (vprior) { vp[0] = vprior[0]; vp[1] = vprior[1]; vp[2] = vprior[2]; }
    }
    // a zero prior would make every derived distance zero (the reference then divides by zero): unit distances instead
    const bool unit_dist = module && vp[0] == 0.0 && vp[1] == 0.0 && vp[2] == 0.0;
    auto inv_dist = [&](int i, double px, double py, double ux, double uy) -> double {
        if (!module) return 1.0;
        double di;
        if (ld.dist(f, i, di)) return 1.0 / di;
        if (unit_dist) return 1.0;
        return ofb_module_inv_dist(px, py, ux, uy, n3, vp);
    };
    int i0 = ld.begin(f), i1 = ld.end(f);
    // M (6), g (3), count
    double acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        double px, py, ux, uy;
        if (!ld.load(f, i, px, py, ux, uy)) continue;
        OfbPointTerms t = ofb_point_terms(variant, px, py, ux, uy, n3, w3, inv_dist(i, px, py, ux, uy));
        double w2 = t.wA * t.wA;
        double xx = px * px, yy = py * py;
        // |X|^2 I - X X^T
        acc[0] += w2 * (yy + 1.0);      // xx
        acc[1] += w2 * (-px * py);      // xy
        acc[2] += w2 * (-px);           // xz
        acc[3] += w2 * (xx + 1.0);      // yy
        acc[4] += w2 * (-py);           // yz
        acc[5] += w2 * (xx + yy);       // zz
        // A^T b = -wA wb (X x beta)
        double s = -t.wA * t.wb;
        acc[6] += s * (py * t.bz - t.by);
        acc[7] += s * (t.bx - px * t.bz);
        acc[8] += s * (px * t.by - py * t.bx);
        acc[9] += 1.0;
    }
    ofb_block_sum<10>(acc, red);
    OfbSolveOut o;
    o.count = (int)(acc[9] + 0.5);
    double M[6] = {acc[0], acc[1], acc[2], acc[3], acc[4], acc[5]};
    double ev[3], q[3][3];
    ofb_jacobi3(M, ev, q);
    // numpy lstsq (rcond=None): singular values below eps*max(M,N)*s_max are treated as zero.
    // s comes from eig(A^T A), so it cannot resolve below ~1e-8*s_max; use 1e-7 as the floor.
    double rows = 3.0 * (double)o.count;
    double rc = 2.220446049250313e-16 * (rows > 3.0 ? rows : 3.0);
    if (rc < 1e-7) rc = 1e-7;
    double smax = ev[0] > 0 ? sqrt(ev[0]) : 0.0;
    int rank = 0;
    double v[3] = {0, 0, 0};
    for (int k = 0; k < 3; ++k) {
        double sk = ev[k] > 0 ? sqrt(ev[k]) : 0.0;
        o.s[k] = sk;
        if (sk > rc * smax && smax > 0) {
            ++rank;
            double c = (q[k][0] * acc[6] + q[k][1] * acc[7] + q[k][2] * acc[8]) * d / ev[k];
            v[0] += c * q[k][0]; v[1] += c * q[k][1]; v[2] += c * q[k][2];
        }
    }
    o.rank = rank;
    // second pass: residual sum ||A_i v - d b_i||^2
    double r[1] = {0};
    for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        double px, py, ux, uy;
        if (!ld.load(f, i, px, py, ux, uy)) continue;
        OfbPointTerms t = ofb_point_terms(variant, px, py, ux, uy, n3, w3, inv_dist(i, px, py, ux, uy));
        double rx = t.wA * (py * v[2] - v[1]) - d * t.wb * t.bx;
        double ry = t.wA * (v[0] - px * v[2]) - d * t.wb * t.by;
        double rz = t.wA * (px * v[1] - py * v[0]) - d * t.wb * t.bz;
        r[0] += rx * rx + ry * ry + rz * rz;
    }
    ofb_block_sum<1>(r, red);
    o.res = r[0];
    // lever arm: v_obs = v' - w x t   (simulation.py:28, evaluate_exp.py:29); NODE has none
    if (variant != OFB_VARIANT_NODE && !module && t3g) {
        v[0] -= w3[1] * t3g[2] - w3[2] * t3g[1];
        v[1] -= w3[2] * t3g[0] - w3[0] * t3g[2];
        v[2] -= w3[0] * t3g[1] - w3[1] * t3g[0];
    }
    o.v[0] = v[0]; o.v[1] = v[1]; o.v[2] = v[2];
    return o;
}
