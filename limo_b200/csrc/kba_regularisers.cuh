// kba_regularisers.cuh -- the f-only residual blocks of the window problem (no landmark involved): ground-plane
// regularisation chain (reference bundle_adjuster_keyframes.cpp:769-818, cost_functors_ceres.hpp:394-438,507-555) and the
// warp-cooperative accumulation of a small residual block into the reduced normal equations.
#pragma once
#include "kba_device.cuh"

namespace kba {

// (I - n n^T / |n|^2) / |n| : Jacobian of FixScaleVectorPlus at delta = 0 (reference local_parameterizations.hpp:146-162)
__device__ inline void dir_plus_jacobian(const double* n, double* P) {
    const double nn = n[0] * n[0] + n[1] * n[1] + n[2] * n[2], inv = 1.0 / sqrt(nn);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) P[3 * i + j] = ((i == j ? 1.0 : 0.0) - n[i] * n[j] / nn) * inv;
}

__device__ inline void dir_plus(const double* n, const double* d, double* o) {
    const double a = n[0] + d[0], b = n[1] + d[1], c = n[2] + d[2];
    const double f = 1.0 / sqrt(a * a + b * b + c * c);
    o[0] = a * f; o[1] = b * f; o[2] = c * f;
}

// d = t_a - R_a R_b^T t_b = (T_a T_b^-1).t ; Ja, Jb: 3x6 local Jacobians (rot | trans) of d w.r.t. poses a and b
__device__ inline void rel_translation(const double* pa, const double* pb, double* d, double* Ja, double* Jb) {
    double Ra[9], Rb[9], c[3], rac[3];
    quat_to_rot<double>(pa, Ra);
    quat_to_rot<double>(pb, Rb);
    const double* tb = pb + 4;
    for (int i = 0; i < 3; ++i) c[i] = Rb[i] * tb[0] + Rb[3 + i] * tb[1] + Rb[6 + i] * tb[2];
    for (int i = 0; i < 3; ++i) rac[i] = Ra[3 * i] * c[0] + Ra[3 * i + 1] * c[1] + Ra[3 * i + 2] * c[2];
    for (int i = 0; i < 3; ++i) d[i] = pa[4 + i] - rac[i];
    if (!Ja) return;
    double Rab[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            Rab[3 * i + j] = Ra[3 * i] * Rb[3 * j] + Ra[3 * i + 1] * Rb[3 * j + 1] + Ra[3 * i + 2] * Rb[3 * j + 2];
    const double X[9] = {0, -rac[2], rac[1], rac[2], 0, -rac[0], -rac[1], rac[0], 0};
    const double Tm[9] = {0, -tb[2], tb[1], tb[2], 0, -tb[0], -tb[1], tb[0], 0};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            Ja[6 * i + j] = 2.0 * X[3 * i + j];
            Ja[6 * i + 3 + j] = (i == j) ? 1.0 : 0.0;
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += Rab[3 * i + k] * Tm[3 * k + j];
            Jb[6 * i + j] = -2.0 * s;
            Jb[6 * i + 3 + j] = -Rab[3 * i + j];
        }
}

// GroundPlaneMotionRegularization (reference cost_functors_ceres.hpp:528-555): r = n0 . normalize((T0 T1^-1).t)
__device__ inline double plane_motion(const double* p0, const double* p1, const double* n0, double* j0, double* j1,
                                      double* jd) {
    double d[3], J0[18], J1[18];
    rel_translation(p0, p1, d, j0 ? J0 : nullptr, J1);
    const double nrm = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    const double u[3] = {d[0] / nrm, d[1] / nrm, d[2] / nrm};
    const double r = n0[0] * u[0] + n0[1] * u[1] + n0[2] * u[2];
    if (j0) {
        const double gq[3] = {(n0[0] - r * u[0]) / nrm, (n0[1] - r * u[1]) / nrm, (n0[2] - r * u[2]) / nrm};
        for (int j = 0; j < 6; ++j) {
            j0[j] = gq[0] * J0[j] + gq[1] * J0[6 + j] + gq[2] * J0[12 + j];
            j1[j] = gq[0] * J1[j] + gq[1] * J1[6 + j] + gq[2] * J1[12 + j];
        }
        double P[9];
        dir_plus_jacobian(n0, P);
        for (int j = 0; j < 3; ++j) jd[j] = u[0] * P[j] + u[1] * P[3 + j] + u[2] * P[6 + j];
    }
    return r;
}

// Warp-cooperative J^T J / J^T r accumulation of one residual block (all 32 lanes pass identical arguments).
// parts: column offsets (or -1 for constant blocks) and sizes; J row-major nres x (sum of sizes), already times sqrt(rho').
template <typename Mat>
__device__ inline void warp_add_block(const Mat& A, double* fdiag, double* g, int nres, const double* r, int nparts,
                                      const int* off, const int* sz, const double* J, int lane) {
    int cols[16];
    double Jc[3][16];
    int m = 0, pos = 0, width = 0;
    for (int a = 0; a < nparts; ++a) width += sz[a];
    for (int a = 0; a < nparts; ++a) {
        for (int c = 0; c < sz[a]; ++c) {
            if (off[a] >= 0) {
                cols[m] = off[a] + c;
                for (int i = 0; i < nres; ++i) Jc[i][m] = J[i * width + pos + c];
                ++m;
            }
        }
        pos += sz[a];
    }
    for (int idx = lane; idx < m * m; idx += 32) {
        const int a = idx / m, b = idx - a * m;
        if (cols[b] > cols[a]) continue;
        double s = 0.0;
        for (int i = 0; i < nres; ++i) s += Jc[i][a] * Jc[i][b];
        A(cols[a], cols[b]) += s;
        if (a == b) {
            fdiag[cols[a]] += s;
            double t = 0.0;
            for (int i = 0; i < nres; ++i) t += Jc[i][a] * r[i];
            g[cols[a]] += t;
        }
    }
    __syncwarp();
}

// Robustified cost of the ground-plane regularisation chain at the given state (TrivialLoss * weight), for windows with
// plane_reg_weight > 0; single thread (called from the LM controller for the candidate, from warp 0 lane 0 at x).
__device__ inline double plane_chain_cost(const WinDesc& wd, const double* pose, const double* plane) {
    const double w = wd.plane_reg_weight;
    double c = 0.0;
    for (int k0 = 0; k0 + 1 < wd.n_kf; ++k0) {
        const double* n0 = plane + 4 * (size_t)(wd.kf_off + k0), *n1 = n0 + 4;
        const double dn[3] = {n1[0] - n0[0], n1[1] - n0[1], n1[2] - n0[2]};
        c += 0.5 * 3.0 * w * (dn[0] * dn[0] + dn[1] * dn[1] + dn[2] * dn[2]);
        const double dd = n1[3] - n0[3];
        c += 0.5 * w * dd * dd;
        const double rm = plane_motion(pose + 7 * (size_t)(wd.kf_off + k0), pose + 7 * (size_t)(wd.kf_off + k0 + 1), n0,
                                       nullptr, nullptr, nullptr);
        c += 0.5 * 2.0 * w * rm * rm;
    }
    for (int k = 0; k < wd.n_kf; ++k) {
        const double* n = plane + 4 * (size_t)(wd.kf_off + k);
        const double e[3] = {0.0 - n[0], 0.0 - n[1], 1.0 - n[2]};
        c += 0.5 * w * (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    }
    return c;
}

}  // namespace kba
