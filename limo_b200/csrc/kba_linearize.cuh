// kba_linearize.cuh -- the whole linearisation of a small window in ONE kernel (fused path, FP64, one camera per keyframe):
//
//   k_linearize = residual / Jacobian evaluation (cost_functors_ceres.hpp:91-155,193-212, as k_eval_obs)
//               + landmark blocks C_j = sum J_l^T J_l, g_j = sum J_l^T r, LM damping, 3x3 Cholesky (as k_landmark_reduce)
//               + V_i = (J_p^T J_l) L^-T of every observation (as k_obs_v2)
//
// Round 1 / early round 2 ran these as three kernels around a materialised J_p (168 B per observation written, then read
// twice: 0.93 ms and 634 B of DRAM traffic per observation for a 148-window batch).  Here nothing of the Jacobian reaches
// memory: a thread evaluates its observation, keeps J_p (18 doubles) in registers across two CTA barriers and writes
// only what the Schur kernel and the back substitution consume -- V_i (144 B per observation) and the landmark's L^-1, z,
// g, lambda.  A rejected LM step (new radius, same x) simply runs the kernel again: re-evaluating is cheaper than re-reading.
//
// Work split: thread = observation (landmark-major order, so a landmark's observations are consecutive threads).  CTA bx
// of a window looks at the observations [224 bx, 224 bx + 256) and OWNS the landmarks whose first observation lies in
// [224 bx, 224 bx + 224); a landmark has at most 32 observations on this path (<= 32 keyframes, one observation per
// keyframe), so every observation of an owned landmark is one of the CTA's 256 -- no tile table, no extra dependent
// load; the threads of landmarks owned by a neighbour idle (12.5 %).  Per-observation contributions to C_j / g_j go
// through shared memory (SoA, conflict-free); the thread of a landmark's FIRST observation sums the segment in
// observation order (the order of the CPU oracle, fixed -> bit-reproducible), factors the damped block and leaves L^-1
// in shared memory for the landmark's other observations.
#pragma once
#include "kba_device.cuh"

namespace kba {

constexpr int kLinTile = 224;   // nominal observations per CTA (see above)
constexpr int kLinThreads = 256;

__global__ void __launch_bounds__(kLinThreads, 2) k_linearize(BatchDev bd, SolveParams sp) {
    const int w = blockIdx.y;
    WinState& st = bd.state[w];
    if (st.phase != PH_ITERATE) return;
    const WinDesc& wd = bd.desc[w];
    const int n_obs = wd.n_obs;
    const int a = blockIdx.x * kLinTile;
    if (a >= n_obs) return;
    const bool lin = st.need_linearize != 0;            // x changed: cost partials and the failure flag are (re)written
    if (!lin && wd.landmarks_fixed) return;             // motion-only window whose step was rejected: nothing depends on the radius
    const int tid = threadIdx.x;
    __shared__ __align__(16) double s_pose[kFusedMaxKf * kPoseStride];
    __shared__ __align__(16) double s_cam[kMaxCam * kCamStride];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ double s_cg[9][kLinThreads];             // per observation: C (6, lower packed) and g (3) contributions
    __shared__ double s_li[6][kLinThreads];             // L^-1 of a landmark, at the slot of its first observation
    __shared__ double s_red[8];
    __shared__ int s_cnt[8];
    const size_t base = (size_t)wd.obs_off;
    const int* lm_ptr = bd.lm_ptr + wd.lm_off + w;
    // my observation: the loads are issued before the staging barrier so that their latency overlaps the bulk copy
    const int o = a + tid;
    const bool have = o < n_obs;
    int j = 0, kf = 0, cam = 0, row0 = -1, p0 = 0, p1 = 0;
    float mu = 0.f, mv = 0.f, md = 0.f;
    if (have) {
        const size_t oo = base + o;
        j = bd.obs_lm[oo];
        kf = bd.obs_kf[oo]; cam = bd.obs_cam[oo]; row0 = bd.obs_row[oo];
        mu = bd.obs_u[oo]; mv = bd.obs_v[oo]; md = bd.obs_d[oo];
        p0 = lm_ptr[j]; p1 = lm_ptr[j + 1];
    }
    const int L = wd.lm_off + j;
    const bool owned = have && p0 >= a && p0 < a + kLinTile;  // else: a neighbour CTA evaluates this observation
    const bool act = owned && bd.lm_active[L] != 0;
    double p[3] = {0, 0, 0}, wgt = 0.0;
    if (act) {
        const double* lmp = bd.lm[st.cur] + 3 * (size_t)L;
        p[0] = lmp[0]; p[1] = lmp[1]; p[2] = lmp[2];
        wgt = bd.lm_weight[L];
    }
    stage_window_bulk(wd, bd.rt[st.cur], bd.cam, s_pose, s_cam, &s_bar);
    // ---- phase 1: evaluate my observation, contributions to its landmark block -> shared memory
    double jp[18];
    bool ok = true;
    double hr = 0.0;
    double cg[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (act) {
        double r[3], jl[9], raw[2];
        ok = eval_observation<double, true>(s_pose + kPoseStride * kf, s_cam + kCamStride * cam, p, (double)mu, (double)mv, (double)md,
                                            wgt, sp.reprojection_thres * sp.reprojection_thres, sp.depth_thres * sp.depth_thres, r, jp,
                                            jl, hr, raw);
        if (ok) {
            cg[0] = jl[0] * jl[0] + jl[3] * jl[3] + jl[6] * jl[6];
            cg[1] = jl[0] * jl[1] + jl[3] * jl[4] + jl[6] * jl[7];
            cg[2] = jl[0] * jl[2] + jl[3] * jl[5] + jl[6] * jl[8];
            cg[3] = jl[1] * jl[1] + jl[4] * jl[4] + jl[7] * jl[7];
            cg[4] = jl[1] * jl[2] + jl[4] * jl[5] + jl[7] * jl[8];
            cg[5] = jl[2] * jl[2] + jl[5] * jl[5] + jl[8] * jl[8];
#pragma unroll
            for (int c = 0; c < 3; ++c) cg[6 + c] = jl[c] * r[0] + jl[3 + c] * r[1] + jl[6 + c] * r[2];
        } else {
            hr = 0.0;
            if (lin) st.eval_failed = 1;  // benign race
        }
    }
#pragma unroll
    for (int q = 0; q < 9; ++q) s_cg[q][tid] = cg[q];
    {   // cost partial of the CTA (fixed-shape reduction) and the observation count of the roofline report
        const double cs = warp_sum(hr);
        const int dn = __reduce_add_sync(0xffffffffu, (act && ok) ? 1 : 0);
        if ((tid & 31) == 0) { s_red[tid >> 5] = cs; s_cnt[tid >> 5] = dn; }
    }
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        int cnt = 0;
        for (int q = 0; q < 8; ++q) { s += s_red[q]; cnt += s_cnt[q]; }
        if (lin) bd.cost_part_x[(size_t)w * bd.cost_parts + blockIdx.x] = s;
        if (cnt) atomicAdd(bd.jac_obs, (unsigned long long)cnt);
    }
    if (wd.landmarks_fixed) return;  // motion-only: the landmark blocks are constant, only the cost at x was needed
    // ---- phase 2: the thread of a landmark's first observation sums the block, damps and factors it
    const int slot = p0 - a;  // shared-memory slot of the landmark's L^-1 (its first observation)
    if (act && o == p0) {
        double c[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
        for (int q = slot; q < p1 - a; ++q) {
#pragma unroll
            for (int e = 0; e < 6; ++e) c[e] += s_cg[e][q];
#pragma unroll
            for (int e = 0; e < 3; ++e) g[e] += s_cg[6 + e][q];
        }
        // the landmark's ground-plane height residual (at most one) is one more row of its Jacobian
        const int gl = (wd.n_gp > 0) ? bd.gp_of_lm[L] : -1;
        const size_t TG = (size_t)bd.tot_gp, G = (size_t)wd.gp_off + (gl >= 0 ? gl : 0);
        double gjl[3] = {0, 0, 0};
        if (gl >= 0) {
            gjl[0] = bd.gp_lin[11 * TG + G]; gjl[1] = bd.gp_lin[12 * TG + G]; gjl[2] = bd.gp_lin[13 * TG + G];
            const double gr = bd.gp_lin[G];
            c[0] += gjl[0] * gjl[0]; c[1] += gjl[0] * gjl[1]; c[2] += gjl[0] * gjl[2];
            c[3] += gjl[1] * gjl[1]; c[4] += gjl[1] * gjl[2]; c[5] += gjl[2] * gjl[2];
            g[0] += gjl[0] * gr; g[1] += gjl[1] * gr; g[2] += gjl[2] * gr;
        }
        // Jacobi scaling (fixed at iteration zero of the solve) and LM damping of the three landmark columns
        const double cd[3] = {c[0], c[3], c[5]};
        double sc[3], lam[3];
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            sc[e] = st.iter0 ? 1.0 / (1.0 + sqrt(cd[e])) : bd.lm_scale[3 * (size_t)L + e];
            const double s2 = sc[e] * sc[e];
            lam[e] = fmin(fmax(cd[e] * s2, sp.min_lm_diagonal), sp.max_lm_diagonal) / (st.radius * s2);
        }
        // Cholesky of C + diag(lam) and the inverse of its factor
        const double a00 = c[0] + lam[0], a11 = c[3] + lam[1], a22 = c[5] + lam[2];
        const double l00 = sqrt(a00);
        const double l10 = c[1] / l00, l20 = c[2] / l00;
        const double d11 = a11 - l10 * l10;
        const double l11 = sqrt(d11);
        const double l21 = (c[4] - l20 * l10) / l11;
        const double d22 = a22 - l20 * l20 - l21 * l21;
        const double l22 = sqrt(d22);
        const bool pd = (a00 > 0.0) && (d11 > 0.0) && (d22 > 0.0);
        if (!pd) {
            st.solve_failed = 1;
#pragma unroll
            for (int e = 0; e < 6; ++e) s_li[e][slot] = 0.0;
        } else {
        const double i00 = 1.0 / l00, i11 = 1.0 / l11, i22 = 1.0 / l22;
        const double i10 = -l10 * i00 * i11;
        const double i21 = -l21 * i11 * i22;
        const double i20 = -(l20 * i00 + l21 * i10) * i22;
        s_li[0][slot] = i00; s_li[1][slot] = i10; s_li[2][slot] = i11; s_li[3][slot] = i20; s_li[4][slot] = i21; s_li[5][slot] = i22;
        double* li = bd.lm_linv + 6 * (size_t)L;
        li[0] = i00; li[1] = i10; li[2] = i11; li[3] = i20; li[4] = i21; li[5] = i22;
        double* zz = bd.lm_z + 3 * (size_t)L;
        zz[0] = i00 * g[0]; zz[1] = i10 * g[0] + i11 * g[1]; zz[2] = i20 * g[0] + i21 * g[1] + i22 * g[2];
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            bd.lm_g[3 * (size_t)L + e] = g[e];
            bd.lm_lambda[3 * (size_t)L + e] = lam[e];
            if (st.iter0) bd.lm_scale[3 * (size_t)L + e] = sc[e];
        }
        if (gl >= 0) {  // V rows of the gp block: E = J_f^T J_l is 10 x 3 (rank one)
            for (int rr = 0; rr < 10; ++rr) {
                const double jf_ = bd.gp_lin[(1 + rr) * TG + G];
                const double e0 = jf_ * gjl[0], e1 = jf_ * gjl[1], e2 = jf_ * gjl[2];
                bd.vgp[(3 * rr + 0) * TG + G] = e0 * i00;
                bd.vgp[(3 * rr + 1) * TG + G] = e0 * i10 + e1 * i11;
                bd.vgp[(3 * rr + 2) * TG + G] = e0 * i20 + e1 * i21 + e2 * i22;
            }
        }
        }
    }
    __syncthreads();
    // ---- phase 3: V_i = (J_p^T J_l) L^-T, J_l = M R(keyframe) with M = J_p[:, 3:6]; W = J_l L^-T first (short chains)
    if (!act || !ok || row0 < 0) return;
    const int ls = slot;
    const double li0 = s_li[0][ls], li1 = s_li[1][ls], li2 = s_li[2][ls], li3 = s_li[3][ls], li4 = s_li[4][ls], li5 = s_li[5][ls];
    const double* R = s_pose + kPoseStride * kf;
    double wm[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double m0 = jp[6 * r + 3], m1 = jp[6 * r + 4], m2 = jp[6 * r + 5];
        const double l0 = m0 * R[0] + m1 * R[3] + m2 * R[6], l1 = m0 * R[1] + m1 * R[4] + m2 * R[7], l2 = m0 * R[2] + m1 * R[5] + m2 * R[8];
        wm[3 * r + 0] = l0 * li0;
        wm[3 * r + 1] = l0 * li1 + l1 * li2;
        wm[3 * r + 2] = l0 * li3 + l1 * li4 + l2 * li5;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {  // 48 contiguous bytes per column; consecutive threads = consecutive observations of the landmark
        double2* out = reinterpret_cast<double2*>(bd.vobs + vobs_index(base, p0, p1, o, c));
#pragma unroll
        for (int h = 0; h < 3; ++h) {
            const int r0 = 2 * h, r1 = 2 * h + 1;  // V[r][c] = sum_k J_p[k][r] W[k][c]
            out[h] = make_double2(jp[r0] * wm[c] + jp[6 + r0] * wm[3 + c] + jp[12 + r0] * wm[6 + c],
                                  jp[r1] * wm[c] + jp[6 + r1] * wm[3 + c] + jp[12 + r1] * wm[6 + c]);
        }
    }
}

}  // namespace kba
