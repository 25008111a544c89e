/*
 * kba_oracle.c -- CPU restatement of limo's keyframe bundle-adjustment window solve (see kba_oracle.h).
 * TEST INFRASTRUCTURE ONLY: never linked into, imported by, or called from the product path.
 *
 * Citations are relative to /root/reference.  [ceres] marks behaviour of ceres-solver 1.13.0, the
 * un-vendored dependency (docker/src/Dockerfile:47), restated from its published algorithm
 * (trust_region_minimizer.cc, levenberg_marquardt_strategy.cc, schur_eliminator_impl.h, loss_function.cc,
 * corrector.cc, local_parameterization.cc) as recorded in SURVEY.md Appendix A.
 */
#include "kba_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>

/* 0: ceres 1.13 order of the tolerance tests (candidate not applied), 1: the <= 1.12 order of SURVEY A.6 (see lm_solve) */
static int g_tolerance_order = 0;
void kbo_set_tolerance_order(int order) { g_tolerance_order = order ? 1 : 0; }

#endif

/* ------------------------------------------------------------------------------------------------ */
/* small linear algebra                                                                              */
/* ------------------------------------------------------------------------------------------------ */

static void cross3(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
static double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void matvec3(const double R[9], const double v[3], double o[3]) {
    for (int i = 0; i < 3; ++i) o[i] = R[3 * i] * v[0] + R[3 * i + 1] * v[1] + R[3 * i + 2] * v[2];
}
static void matTvec3(const double R[9], const double v[3], double o[3]) {
    for (int i = 0; i < 3; ++i) o[i] = R[i] * v[0] + R[3 + i] * v[1] + R[6 + i] * v[2];
}
/* row vector m^T (1x3) times 3x3 R */
static void rowmat3(const double m[3], const double R[9], double o[3]) {
    for (int j = 0; j < 3; ++j) o[j] = m[0] * R[j] + m[1] * R[3 + j] + m[2] * R[6 + j];
}

/* Eigen::Quaternion::toRotationMatrix() WITHOUT normalisation (definitions.hpp:75-83 -> Transform::rotate;
 * SURVEY A.1).  q = (w,x,y,z), R row-major. */
static void quat_to_R(const double q[4], double R[9]) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

/* [ceres] QuaternionParameterization::Plus x IdentityParameterization(3) (cpp:181-182) */
void kbo_pose_plus(const double pose[7], const double delta[6], double out[7]) {
    const double nd = sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
    if (nd > 0.0) {
        const double s = sin(nd) / nd;
        const double qd[4] = {cos(nd), s * delta[0], s * delta[1], s * delta[2]};
        const double* w = pose;
        out[0] = qd[0] * w[0] - qd[1] * w[1] - qd[2] * w[2] - qd[3] * w[3];
        out[1] = qd[0] * w[1] + qd[1] * w[0] + qd[2] * w[3] - qd[3] * w[2];
        out[2] = qd[0] * w[2] - qd[1] * w[3] + qd[2] * w[0] + qd[3] * w[1];
        out[3] = qd[0] * w[3] + qd[1] * w[2] - qd[2] * w[1] + qd[3] * w[0];
    } else {
        for (int i = 0; i < 4; ++i) out[i] = pose[i];
    }
    for (int i = 0; i < 3; ++i) out[4 + i] = pose[4 + i] + delta[3 + i];
}

/* FixScaleVectorPlus with scale 1 (local_parameterizations.hpp:146-162) */
void kbo_dir_plus(const double n[3], const double d[3], double out[3]) {
    const double a = n[0] + d[0], b = n[1] + d[1], c = n[2] + d[2];
    const double f = 1.0 / sqrt(a * a + b * b + c * c);
    out[0] = a * f; out[1] = b * f; out[2] = c * f;
}
/* Jacobian of FixScaleVectorPlus at delta = 0 (what AutoDiffLocalParameterization yields, cpp:190-193):
 * (I - n n^T / |n|^2) / |n|, row-major 3x3. */
static void dir_plus_jacobian(const double n[3], double P[9]) {
    const double nn = dot3(n, n), inv = 1.0 / sqrt(nn);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) P[3 * i + j] = ((i == j ? 1.0 : 0.0) - n[i] * n[j] / nn) * inv;
}

/* ------------------------------------------------------------------------------------------------ */
/* residuals with analytic local Jacobians                                              */
/* ------------------------------------------------------------------------------------------------ */

/* p_C = T_CX * T_XO * p_O (cost_functors_ceres.hpp:116-122); also returns a = R(q) p_O, R, R_c */
static void to_camera(const double pose[7], const double cam[7], const double p[3], double R[9], double Rc[9],
                      double a[3], double pc[3]) {
    quat_to_R(pose, R);
    quat_to_R(cam, Rc);
    matvec3(R, p, a);
    double px[3] = {a[0] + pose[4], a[1] + pose[5], a[2] + pose[6]};
    matvec3(Rc, px, pc);
    pc[0] += cam[4]; pc[1] += cam[5]; pc[2] += cam[6];
}

/* one residual row whose gradient w.r.t. the point in the vehicle frame is m^T:
 * d/d(delta_rot) = -2 (m x a)^T, d/d(delta_t) = m^T, d/d(p_O) = m^T R */
static void row_jacobians(const double m[3], const double a[3], const double R[9], double jp[6], double jl[3]) {
    double c[3];
    cross3(m, a, c);
    jp[0] = -2 * c[0]; jp[1] = -2 * c[1]; jp[2] = -2 * c[2];
    jp[3] = m[0]; jp[4] = m[1]; jp[5] = m[2];
    rowmat3(m, R, jl);
}

int kbo_reprojection(const double pose[7], const double cam[7], const double intr[3], const double p[3], double u,
                     double v, double res[2], double jp[12], double jl[6]) {
    double R[9], Rc[9], a[3], pc[3];
    to_camera(pose, cam, p, R, Rc, a, pc);
    if (!(fabs(pc[2]) >= 0.01)) return 0; /* cost_functors_ceres.hpp:78-83 */
    const double f = intr[0], iz = 1.0 / pc[2];
    const double xn = pc[0] * iz, yn = pc[1] * iz;
    res[0] = f * xn + intr[1] - u; /* :85-86, :151-152 */
    res[1] = f * yn + intr[2] - v;
    if (jp || jl) {
        /* Pi = f/z [[1,0,-x/z],[0,1,-y/z]];  m_row = Pi_row * R_c */
        const double pi0[3] = {f * iz, 0, -f * xn * iz}, pi1[3] = {0, f * iz, -f * yn * iz};
        double m0[3], m1[3], t6[6], t3[3];
        rowmat3(pi0, Rc, m0);
        rowmat3(pi1, Rc, m1);
        row_jacobians(m0, a, R, t6, t3);
        if (jp) memcpy(jp, t6, sizeof t6);
        if (jl) memcpy(jl, t3, sizeof t3);
        row_jacobians(m1, a, R, t6, t3);
        if (jp) memcpy(jp + 6, t6, sizeof t6);
        if (jl) memcpy(jl + 3, t3, sizeof t3);
    }
    return 1;
}

void kbo_depth(const double pose[7], const double cam[7], const double p[3], double d, double res[1], double jp[6],
               double jl[3]) {
    double R[9], Rc[9], a[3], pc[3];
    to_camera(pose, cam, p, R, Rc, a, pc);
    res[0] = pc[2] - d; /* cost_functors_ceres.hpp:207-209 */
    if (jp || jl) {
        double t6[6], t3[3];
        row_jacobians(Rc + 6, a, R, t6, t3);
        if (jp) memcpy(jp, t6, sizeof t6);
        if (jl) memcpy(jl, t3, sizeof t3);
    }
}

void kbo_gp_height(const double pose[7], const double n[3], double dist, const double p[3], double res[1],
                   double jpose[6], double jdir[3], double jdist[1], double jpoint[3]) {
    double R[9], a[3];
    quat_to_R(pose, R);
    matvec3(R, p, a);
    const double px[3] = {a[0] + pose[4], a[1] + pose[5], a[2] + pose[6]};
    res[0] = dot3(n, px) + dist; /* cost_functors_ceres.hpp:370 */
    if (jpose) {
        double t3[3];
        row_jacobians(n, a, R, jpose, jpoint ? jpoint : t3);
    } else if (jpoint) {
        rowmat3(n, R, jpoint);
    }
    if (jdir) {
        double P[9];
        dir_plus_jacobian(n, P);
        rowmat3(px, P, jdir);
    }
    if (jdist) jdist[0] = 1.0;
}

/* d = t_a - R_a R_b^T t_b = (T_a T_b^-1).t with Jacobians w.r.t. the local increments of a and b:
 * dd/d(dr_a) = 2 [R_a c]x, dd/d(dt_a) = I, dd/d(dr_b) = -2 R_a R_b^T [t_b]x, dd/d(dt_b) = -R_a R_b^T,  c = R_b^T t_b */
static void rel_translation(const double pa[7], const double pb[7], double d[3], double Ja[18], double Jb[18]) {
    double Ra[9], Rb[9], c[3], Rac[3];
    quat_to_R(pa, Ra);
    quat_to_R(pb, Rb);
    matTvec3(Rb, pb + 4, c);
    matvec3(Ra, c, Rac);
    for (int i = 0; i < 3; ++i) d[i] = pa[4 + i] - Rac[i];
    if (!Ja) return;
    double Rab[9]; /* R_a R_b^T */
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            Rab[3 * i + j] = Ra[3 * i] * Rb[3 * j] + Ra[3 * i + 1] * Rb[3 * j + 1] + Ra[3 * i + 2] * Rb[3 * j + 2];
    const double* tb = pb + 4;
    const double X[9] = {0, -Rac[2], Rac[1], Rac[2], 0, -Rac[0], -Rac[1], Rac[0], 0};
    const double T[9] = {0, -tb[2], tb[1], tb[2], 0, -tb[0], -tb[1], tb[0], 0};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            Ja[6 * i + j] = 2 * X[3 * i + j];
            Ja[6 * i + 3 + j] = (i == j) ? 1.0 : 0.0;
            double s = 0;
            for (int k = 0; k < 3; ++k) s += Rab[3 * i + k] * T[3 * k + j];
            Jb[6 * i + j] = -2 * s;
            Jb[6 * i + 3 + j] = -Rab[3 * i + j];
        }
}

void kbo_scale_reg(const double pose1[7], const double pose0[7], double scale, double res[1], double j1[6],
                   double j0[6]) {
    double d[3], J1[18], J0[18];
    rel_translation(pose1, pose0, d, (j1 || j0) ? J1 : NULL, J0); /* cost_functors_ceres.hpp:236 */
    const double nrm = sqrt(dot3(d, d));
    res[0] = nrm - scale; /* :238-239 */
    if (j1 || j0) {
        const double u[3] = {d[0] / nrm, d[1] / nrm, d[2] / nrm};
        for (int j = 0; j < 6; ++j) {
            if (j1) j1[j] = u[0] * J1[j] + u[1] * J1[6 + j] + u[2] * J1[12 + j];
            if (j0) j0[j] = u[0] * J0[j] + u[1] * J0[6 + j] + u[2] * J0[12 + j];
        }
    }
}

void kbo_gp_motion(const double pose0[7], const double pose1[7], const double n0[3], double res[1], double j0[6],
                   double j1[6], double jdir[3]) {
    double d[3], J0[18], J1[18];
    rel_translation(pose0, pose1, d, (j0 || j1) ? J0 : NULL, J1); /* cost_functors_ceres.hpp:539 */
    const double nrm = sqrt(dot3(d, d));
    const double u[3] = {d[0] / nrm, d[1] / nrm, d[2] / nrm}; /* :542 */
    res[0] = dot3(n0, u);                                     /* :545 */
    if (j0 || j1) {
        /* dr/dd = n0^T (I - u u^T) / |d| */
        const double nu = dot3(n0, u);
        const double g[3] = {(n0[0] - nu * u[0]) / nrm, (n0[1] - nu * u[1]) / nrm, (n0[2] - nu * u[2]) / nrm};
        for (int j = 0; j < 6; ++j) {
            if (j0) j0[j] = g[0] * J0[j] + g[1] * J0[6 + j] + g[2] * J0[12 + j];
            if (j1) j1[j] = g[0] * J1[j] + g[1] * J1[6 + j] + g[2] * J1[12 + j];
        }
    }
    if (jdir) {
        double P[9];
        dir_plus_jacobian(n0, P);
        rowmat3(u, P, jdir);
    }
}

void kbo_speed_reg(const double pose[7], const double Tob[7], double dt, const double vb[3], double res[3],
                   double jp[18]) {
    /* (T_cur * T_ob).t = R t_ob + t  (cost_functors_ceres.hpp:322-337) */
    double R[9], a[3];
    quat_to_R(pose, R);
    matvec3(R, Tob + 4, a);
    for (int i = 0; i < 3; ++i) res[i] = (a[i] + pose[4 + i]) / dt - vb[i];
    if (jp) {
        const double X[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                jp[6 * i + j] = -2 * X[3 * i + j] / dt;
                jp[6 * i + 3 + j] = (i == j ? 1.0 : 0.0) / dt;
            }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* [ceres] loss functions (loss_function.cc) and first-order corrector (corrector.cc), SURVEY A.3       */
/* ------------------------------------------------------------------------------------------------ */
enum { LOSS_TRIVIAL = 0, LOSS_CAUCHY = 1, LOSS_HUBER = 2 };

static void loss_eval(int kind, double a, double w, double s, double rho[2]) {
    if (kind == LOSS_CAUCHY) {
        const double b = a * a, c = 1.0 / b;
        const double sum = 1.0 + s * c, inv = 1.0 / sum;
        rho[0] = b * log(sum);
        rho[1] = fmax(DBL_MIN, inv);
    } else if (kind == LOSS_HUBER) {
        const double b = a * a;
        if (s > b) {
            const double r = sqrt(s);
            rho[0] = 2.0 * a * r - b;
            rho[1] = fmax(DBL_MIN, a / r);
        } else {
            rho[0] = s;
            rho[1] = 1.0;
        }
    } else {
        rho[0] = s;
        rho[1] = 1.0;
    }
    rho[0] *= w; /* ScaledLoss */
    rho[1] *= w;
}

/* ------------------------------------------------------------------------------------------------ */
/* program = the reduced ceres program of one inner Solve                                            */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    double* pose;  /* n_kf*7 */
    double* plane; /* n_kf*4 */
    double* lm;    /* n_lm*3 */
} state_t;

typedef struct {
    int lm;        /* landmark (e-block) index or -1 */
    int nres, nf;
    int foff[3], fsz[3];
    double r[3];
    double Jf[18]; /* parts consecutive, each nres x fsz row-major */
    double Je[9];  /* nres x 3 */
} rblock;

typedef struct {
    const kba_window* w;
    const kba_options* opt;
    int n_kf, n_lm, n_obs, n_gp;
    unsigned char* lm_active; /* trimmed -> 0 */
    int* gp_of_lm;            /* gp index per landmark or -1 */
    /* layout of this solve */
    int* off_pose;  /* f offset or -1 (constant / absent) */
    int* off_dir;
    int* off_dist;
    unsigned char* pose_in, *dir_in, *dist_in; /* block present in program (even if constant) */
    unsigned char* lm_in;                      /* landmark block in program (variable) */
    int n_f;
    int n_reg;     /* regulariser blocks */
    int n_blocks;  /* n_obs + n_gp + n_reg (slots; inactive ones have nres = 0) */
    rblock* blk;
    int num_threads;
} program;

static void state_alloc(state_t* s, int n_kf, int n_lm) {
    s->pose = (double*)malloc(sizeof(double) * 7 * (n_kf > 0 ? n_kf : 1));
    s->plane = (double*)malloc(sizeof(double) * 4 * (n_kf > 0 ? n_kf : 1));
    s->lm = (double*)malloc(sizeof(double) * 3 * (n_lm > 0 ? n_lm : 1));
}
static void state_free(state_t* s) { free(s->pose); free(s->plane); free(s->lm); }
static void state_copy(state_t* d, const state_t* s, int n_kf, int n_lm) {
    memcpy(d->pose, s->pose, sizeof(double) * 7 * n_kf);
    memcpy(d->plane, s->plane, sizeof(double) * 4 * n_kf);
    memcpy(d->lm, s->lm, sizeof(double) * 3 * n_lm);
}

static int obs_cam(const kba_window* w, int o) { return w->obs_cam ? w->obs_cam[o] : 0; }

/* Which blocks does the (current) problem contain?  A parameter block is in the program iff a residual block
 * references it (removeUnconstraintParameters, robust_solving.cpp:127-137) and it is not constant
 * (cpp:198-219, 722-728, 862). */
static void program_layout(program* P) {
    const kba_window* w = P->w;
    const int K = P->n_kf;
    memset(P->pose_in, 0, K); memset(P->dir_in, 0, K); memset(P->dist_in, 0, K);
    memset(P->lm_in, 0, P->n_lm > 0 ? P->n_lm : 1);
    for (int j = 0; j < P->n_lm; ++j) {
        if (!P->lm_active[j]) continue;
        int has = 0;
        for (int o = w->lm_obs_ptr[j]; o < w->lm_obs_ptr[j + 1]; ++o) { P->pose_in[w->obs_kf[o]] = 1; has = 1; }
        const int g = P->gp_of_lm[j];
        if (g >= 0) {
            const int k = w->gp_kf[g];
            P->pose_in[k] = P->dir_in[k] = P->dist_in[k] = 1;
            has = 1;
        }
        if (has && !w->landmarks_fixed) P->lm_in[j] = 1;
    }
    if (w->scale_weight > 0) P->pose_in[w->scale_kf0] = P->pose_in[w->scale_kf1] = 1;
    if (w->plane_reg_weight > 0 && K > 1)
        for (int k = 0; k < K; ++k) P->pose_in[k] = P->dir_in[k] = P->dist_in[k] = 1;
    if (w->speed_weight > 0) P->pose_in[w->speed_kf] = 1;
    int n = 0;
    for (int k = 0; k < K; ++k) {
        const int fixed = w->kf_fixed[k];
        P->off_pose[k] = (P->pose_in[k] && !fixed) ? n : -1;
        if (P->off_pose[k] >= 0) n += 6;
        P->off_dir[k] = (P->dir_in[k] && !fixed) ? n : -1;
        if (P->off_dir[k] >= 0) n += 3;
        P->off_dist[k] = (P->dist_in[k] && !fixed && !w->plane_dist_fixed) ? n : -1;
        if (P->off_dist[k] >= 0) n += 1;
    }
    P->n_f = n;
}

/* append an f-part to a block; J is nres x sz row-major, already multiplied by sqrt(rho') */
static void blk_add_part(rblock* b, int off, int sz, const double* J) {
    if (off < 0) return; /* constant parameter block: no Jacobian columns */
    int pos = 0;
    for (int a = 0; a < b->nf; ++a) pos += b->nres * b->fsz[a];
    memcpy(b->Jf + pos, J, sizeof(double) * b->nres * sz);
    b->foff[b->nf] = off;
    b->fsz[b->nf] = sz;
    b->nf++;
}

/* Evaluate one observation = LandmarkDepthError block (iff d > 0, cpp:578) + ReprojectionErrorWithQuaternions block,
 * each with ScaledLoss(CauchyLoss(thres), landmark weight) (cpp:587-620).  Rows: (u, v, depth).
 * Returns 0 if the reprojection functor fails.  cost gets 0.5*rho of both blocks.  raw_norm (optional): the
 * un-robustified block norms {reprojection, depth} used by the trimmer. */
static int eval_obs(const program* P, const state_t* x, int o, int lm, int want_jac, double* cost, double r[3],
                    double Jp[18], double Jl[9], double raw_norm[2]) {
    const kba_window* w = P->w;
    const int k = w->obs_kf[o], c = obs_cam(w, o);
    const double u = (double)w->obs_u[o], v = (double)w->obs_v[o]; /* cpp:606-607 float -> double */
    const float df = w->obs_d[o];
    const double wt = w->lm_weight[lm];
    double rr[2], jp[12], jl[6];
    if (!kbo_reprojection(x->pose + 7 * k, w->cam_pose + 7 * c, w->cam_intr + 3 * c, x->lm + 3 * lm, u, v, rr,
                          want_jac ? jp : NULL, want_jac ? jl : NULL))
        return 0;
    double rho[2];
    const double s = rr[0] * rr[0] + rr[1] * rr[1];
    loss_eval(LOSS_CAUCHY, P->opt->reprojection_thres, wt, s, rho);
    *cost += 0.5 * rho[0];
    if (raw_norm) raw_norm[0] = sqrt(s);
    double sq = sqrt(rho[1]);
    r[0] = sq * rr[0];
    r[1] = sq * rr[1];
    r[2] = 0.0;
    if (want_jac) {
        for (int i = 0; i < 12; ++i) Jp[i] = sq * jp[i];
        for (int i = 0; i < 6; ++i) Jl[i] = sq * jl[i];
        for (int i = 0; i < 6; ++i) Jp[12 + i] = 0.0;
        for (int i = 0; i < 3; ++i) Jl[6 + i] = 0.0;
    }
    if (raw_norm) raw_norm[1] = -1.0;
    if (df > 0.0f) {
        double rd[1], jpd[6], jld[3];
        kbo_depth(x->pose + 7 * k, w->cam_pose + 7 * c, x->lm + 3 * lm, (double)df, rd, want_jac ? jpd : NULL,
                  want_jac ? jld : NULL);
        const double sd = rd[0] * rd[0];
        loss_eval(LOSS_CAUCHY, P->opt->depth_thres, wt, sd, rho);
        *cost += 0.5 * rho[0];
        if (raw_norm) raw_norm[1] = fabs(rd[0]);
        sq = sqrt(rho[1]);
        r[2] = sq * rd[0];
        if (want_jac) {
            for (int i = 0; i < 6; ++i) Jp[12 + i] = sq * jpd[i];
            for (int i = 0; i < 3; ++i) Jl[6 + i] = sq * jld[i];
        }
    }
    return 1;
}

/* Evaluate the whole program at x.  With want_jac, fills P->blk (robustified residuals + Jacobian blocks).
 * Returns 0 on evaluation failure.  Block slots: [0,n_obs) observations, [n_obs, n_obs+n_gp) ground plane,
 * then regularisers in the order of SURVEY Appendix B. */
static int program_evaluate(program* P, const state_t* x, int want_jac, double* cost_out) {
    const kba_window* w = P->w;
    const kba_options* opt = P->opt;
    double cost = 0.0;
    int ok = 1;
    /* --- observations (addKeyframeToProblem, cpp:564-627) --- */
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(+ : cost) reduction(&& : ok) num_threads(P->num_threads) if (P->num_threads > 1)
#endif
    for (int j = 0; j < P->n_lm; ++j) {
        if (!P->lm_active[j]) {
            if (want_jac)
                for (int o = w->lm_obs_ptr[j]; o < w->lm_obs_ptr[j + 1]; ++o) P->blk[o].nres = 0;
            continue;
        }
        for (int o = w->lm_obs_ptr[j]; o < w->lm_obs_ptr[j + 1]; ++o) {
            double r[3], Jp[18], Jl[9], c = 0.0;
            if (!eval_obs(P, x, o, j, want_jac, &c, r, Jp, Jl, NULL)) { ok = 0; continue; }
            cost += c;
            if (want_jac) {
                rblock* b = &P->blk[o];
                b->lm = P->lm_in[j] ? j : -1;
                b->nres = (w->obs_d[o] > 0.0f) ? 3 : 2;
                b->nf = 0;
                memcpy(b->r, r, sizeof r);
                blk_add_part(b, P->off_pose[w->obs_kf[o]], 6, Jp); /* first nres rows of the 3x6 */
                memcpy(b->Je, Jl, sizeof(double) * 3 * b->nres);
            }
        }
    }
    if (!ok) return 0;
    /* --- ground plane height residuals (addGroundPlaneResiduals, cpp:517-562): ScaledLoss(Huber(0.1), w) --- */
    for (int g = 0; g < P->n_gp; ++g) {
        const int j = w->gp_lm[g], k = w->gp_kf[g];
        rblock* b = want_jac ? &P->blk[P->n_obs + g] : NULL;
        if (!P->lm_active[j]) { if (b) b->nres = 0; continue; }
        double r[1], jp[6], jd[3], jdist[1], jl[3], rho[2];
        kbo_gp_height(x->pose + 7 * k, x->plane + 4 * k, x->plane[4 * k + 3], x->lm + 3 * j, r, want_jac ? jp : NULL,
                      want_jac ? jd : NULL, want_jac ? jdist : NULL, want_jac ? jl : NULL);
        loss_eval(LOSS_HUBER, opt->gp_huber, w->gp_weight[g], r[0] * r[0], rho);
        cost += 0.5 * rho[0];
        if (b) {
            const double sq = sqrt(rho[1]);
            b->lm = P->lm_in[j] ? j : -1;
            b->nres = 1; b->nf = 0;
            b->r[0] = sq * r[0];
            for (int i = 0; i < 6; ++i) jp[i] *= sq;
            for (int i = 0; i < 3; ++i) { jd[i] *= sq; jl[i] *= sq; }
            jdist[0] *= sq;
            blk_add_part(b, P->off_pose[k], 6, jp);
            blk_add_part(b, P->off_dir[k], 3, jd);
            blk_add_part(b, P->off_dist[k], 1, jdist);
            memcpy(b->Je, jl, sizeof jl);
        }
    }
    /* --- regularisers --- */
    int slot = P->n_obs + P->n_gp;
    if (w->scale_weight > 0) { /* addScaleRegularization, cpp:890-904: TrivialLoss * weight */
        const int k1 = w->scale_kf1, k0 = w->scale_kf0;
        double r[1], j1[6], j0[6], rho[2];
        kbo_scale_reg(x->pose + 7 * k1, x->pose + 7 * k0, w->scale_value, r, want_jac ? j1 : NULL, want_jac ? j0 : NULL);
        loss_eval(LOSS_TRIVIAL, 0, w->scale_weight, r[0] * r[0], rho);
        cost += 0.5 * rho[0];
        if (want_jac) {
            rblock* b = &P->blk[slot];
            const double sq = sqrt(rho[1]);
            b->lm = -1; b->nres = 1; b->nf = 0; b->r[0] = sq * r[0];
            for (int i = 0; i < 6; ++i) { j1[i] *= sq; j0[i] *= sq; }
            blk_add_part(b, P->off_pose[k1], 6, j1);
            blk_add_part(b, P->off_pose[k0], 6, j0);
        }
        slot++;
    }
    if (w->plane_reg_weight > 0 && P->n_kf > 1) { /* addGroundplaneRegularization, cpp:769-818 */
        const double wt = w->plane_reg_weight;
        for (int k0 = 0; k0 + 1 < P->n_kf; ++k0) {
            const int k1 = k0 + 1;
            const double* n0 = x->plane + 4 * k0, *n1 = x->plane + 4 * k1;
            double rho[2];
            { /* VectorDifferenceRegularization(dir1, dir0): r = dir1 - dir0, weight 3w (cpp:779-783) */
                double r[3] = {n1[0] - n0[0], n1[1] - n0[1], n1[2] - n0[2]};
                loss_eval(LOSS_TRIVIAL, 0, 3.0 * wt, dot3(r, r), rho);
                cost += 0.5 * rho[0];
                if (want_jac) {
                    rblock* b = &P->blk[slot];
                    const double sq = sqrt(rho[1]);
                    double P1[9], P0[9];
                    dir_plus_jacobian(n1, P1);
                    dir_plus_jacobian(n0, P0);
                    b->lm = -1; b->nres = 3; b->nf = 0;
                    for (int i = 0; i < 3; ++i) b->r[i] = sq * r[i];
                    for (int i = 0; i < 9; ++i) { P1[i] *= sq; P0[i] *= -sq; }
                    blk_add_part(b, P->off_dir[k1], 3, P1);
                    blk_add_part(b, P->off_dir[k0], 3, P0);
                }
                slot++;
            }
            { /* GroundPlaneDistanceRegularization(dist1, dist0), weight w (cpp:786-790) */
                double r = x->plane[4 * k1 + 3] - x->plane[4 * k0 + 3];
                loss_eval(LOSS_TRIVIAL, 0, wt, r * r, rho);
                cost += 0.5 * rho[0];
                if (want_jac) {
                    rblock* b = &P->blk[slot];
                    const double sq = sqrt(rho[1]);
                    double p1 = sq, p0 = -sq;
                    b->lm = -1; b->nres = 1; b->nf = 0; b->r[0] = sq * r;
                    blk_add_part(b, P->off_dist[k1], 1, &p1);
                    blk_add_part(b, P->off_dist[k0], 1, &p0);
                }
                slot++;
            }
            { /* GroundPlaneMotionRegularization(pose0, pose1, dir0), weight 2w (cpp:794-799) */
                double r[1], j0[6], j1[6], jd[3];
                kbo_gp_motion(x->pose + 7 * k0, x->pose + 7 * k1, n0, r, want_jac ? j0 : NULL, want_jac ? j1 : NULL,
                              want_jac ? jd : NULL);
                loss_eval(LOSS_TRIVIAL, 0, 2.0 * wt, r[0] * r[0], rho);
                cost += 0.5 * rho[0];
                if (want_jac) {
                    rblock* b = &P->blk[slot];
                    const double sq = sqrt(rho[1]);
                    b->lm = -1; b->nres = 1; b->nf = 0; b->r[0] = sq * r[0];
                    for (int i = 0; i < 6; ++i) { j0[i] *= sq; j1[i] *= sq; }
                    for (int i = 0; i < 3; ++i) jd[i] *= sq;
                    blk_add_part(b, P->off_pose[k0], 6, j0);
                    blk_add_part(b, P->off_pose[k1], 6, j1);
                    blk_add_part(b, P->off_dir[k0], 3, jd);
                }
                slot++;
            }
        }
        for (int k = 0; k < P->n_kf; ++k) { /* VectorDifferenceRegularization2((0,0,1)): r = (0,0,1) - dir, weight w (cpp:810-816) */
            const double* n = x->plane + 4 * k;
            double r[3] = {0.0 - n[0], 0.0 - n[1], 1.0 - n[2]}, rho[2];
            loss_eval(LOSS_TRIVIAL, 0, wt, dot3(r, r), rho);
            cost += 0.5 * rho[0];
            if (want_jac) {
                rblock* b = &P->blk[slot];
                const double sq = sqrt(rho[1]);
                double Pn[9];
                dir_plus_jacobian(n, Pn);
                b->lm = -1; b->nres = 3; b->nf = 0;
                for (int i = 0; i < 3; ++i) b->r[i] = sq * r[i];
                for (int i = 0; i < 9; ++i) Pn[i] *= -sq;
                blk_add_part(b, P->off_dir[k], 3, Pn);
            }
            slot++;
        }
    }
    if (w->speed_weight > 0) { /* adjustPoseOnly speed prior, cpp:835-853 */
        double r[3], jp[18], rho[2];
        kbo_speed_reg(x->pose + 7 * w->speed_kf, w->speed_T_origin_before, w->speed_dt, w->speed_v_before, r,
                      want_jac ? jp : NULL);
        loss_eval(LOSS_TRIVIAL, 0, w->speed_weight, dot3(r, r), rho);
        cost += 0.5 * rho[0];
        if (want_jac) {
            rblock* b = &P->blk[slot];
            const double sq = sqrt(rho[1]);
            b->lm = -1; b->nres = 3; b->nf = 0;
            for (int i = 0; i < 3; ++i) b->r[i] = sq * r[i];
            for (int i = 0; i < 18; ++i) jp[i] *= sq;
            blk_add_part(b, P->off_pose[w->speed_kf], 6, jp);
        }
        slot++;
    }
    *cost_out = cost;
    return 1;
}

/* ------------------------------------------------------------------------------------------------ */
/* [ceres] dense Cholesky (Eigen LLT in DenseSchurComplementSolver) on a row-major lower-stored n x n  */
/* ------------------------------------------------------------------------------------------------ */
static int cholesky_lower(double* A, int n) {
    for (int j = 0; j < n; ++j) {
        double d = A[(size_t)j * n + j];
        for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
        if (!(d > 0.0) || !isfinite(d)) return 0;
        d = sqrt(d);
        A[(size_t)j * n + j] = d;
        const double inv = 1.0 / d;
        for (int i = j + 1; i < n; ++i) {
            double s = A[(size_t)i * n + j];
            const double* ai = A + (size_t)i * n, *aj = A + (size_t)j * n;
            for (int k = 0; k < j; ++k) s -= ai[k] * aj[k];
            A[(size_t)i * n + j] = s * inv;
        }
    }
    return 1;
}
static void cholesky_solve(const double* L, int n, double* b) {
    for (int i = 0; i < n; ++i) {
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= L[(size_t)i * n + k] * b[k];
        b[i] = s / L[(size_t)i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (int k = i + 1; k < n; ++k) s -= L[(size_t)k * n + i] * b[k];
        b[i] = s / L[(size_t)i * n + i];
    }
}
/* inverse of a symmetric positive definite 3x3 through its Cholesky factor ([ceres] InvertPSDMatrix<3>) */
static int inv_spd3(const double C[9], double Ci[9]) {
    double L[9];
    memcpy(L, C, sizeof L);
    if (!cholesky_lower(L, 3)) return 0;
    for (int c = 0; c < 3; ++c) {
        double e[3] = {0, 0, 0};
        e[c] = 1.0;
        cholesky_solve(L, 3, e);
        for (int r = 0; r < 3; ++r) Ci[3 * r + c] = e[r];
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------------ */
/* [ceres] Levenberg-Marquardt step through DENSE_SCHUR                                  */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    double* scale_f; /* Jacobi scaling of the f columns (n_f) */
    double* scale_e; /* 3 per landmark */
    double* diag_f;  /* clamped squared column norms of the scaled Jacobian */
    double* diag_e;
    double* grad_f;  /* gradient J~^T r~ (unscaled) */
    double* grad_e;
    double* step_f;  /* trust-region step in SCALED coordinates (y), then delta */
    double* step_e;
    double* S;       /* n_f x n_f */
    double* rhs;     /* n_f */
    double* Cinv;    /* 9 per landmark */
    double* ge;      /* 3 per landmark (scaled J_e^T r) */
} lm_work;

/* squared column norms of the (unscaled) robustified Jacobian and the gradient */
static void column_norms_and_gradient(const program* P, double* cn_f, double* cn_e, double* g_f, double* g_e) {
    memset(cn_f, 0, sizeof(double) * (P->n_f > 0 ? P->n_f : 1));
    memset(g_f, 0, sizeof(double) * (P->n_f > 0 ? P->n_f : 1));
    memset(cn_e, 0, sizeof(double) * 3 * (P->n_lm > 0 ? P->n_lm : 1));
    memset(g_e, 0, sizeof(double) * 3 * (P->n_lm > 0 ? P->n_lm : 1));
    for (int b = 0; b < P->n_blocks; ++b) {
        const rblock* B = &P->blk[b];
        if (B->nres == 0) continue;
        int pos = 0;
        for (int a = 0; a < B->nf; ++a) {
            for (int i = 0; i < B->nres; ++i)
                for (int c = 0; c < B->fsz[a]; ++c) {
                    const double v = B->Jf[pos + i * B->fsz[a] + c];
                    cn_f[B->foff[a] + c] += v * v;
                    g_f[B->foff[a] + c] += v * B->r[i];
                }
            pos += B->nres * B->fsz[a];
        }
        if (B->lm >= 0)
            for (int i = 0; i < B->nres; ++i)
                for (int c = 0; c < 3; ++c) {
                    const double v = B->Je[3 * i + c];
                    cn_e[3 * B->lm + c] += v * v;
                    g_e[3 * B->lm + c] += v * B->r[i];
                }
    }
}

/* accumulate one block's F^T F, F^T r into S / rhs (scaled columns); sign = +1 */
static void add_ff(const program* P, const rblock* B, const double* sf, double* S, double* rhs) {
    const int n = P->n_f;
    int pa = 0;
    for (int a = 0; a < B->nf; ++a) {
        int pb = 0;
        for (int bb = 0; bb < B->nf; ++bb) {
            for (int ca = 0; ca < B->fsz[a]; ++ca)
                for (int cb = 0; cb < B->fsz[bb]; ++cb) {
                    double s = 0;
                    for (int i = 0; i < B->nres; ++i)
                        s += B->Jf[pa + i * B->fsz[a] + ca] * B->Jf[pb + i * B->fsz[bb] + cb];
                    const int ra = B->foff[a] + ca, rb = B->foff[bb] + cb;
                    S[(size_t)ra * n + rb] += s * sf[ra] * sf[rb];
                }
            pb += B->nres * B->fsz[bb];
        }
        for (int ca = 0; ca < B->fsz[a]; ++ca) {
            double s = 0;
            for (int i = 0; i < B->nres; ++i) s += B->Jf[pa + i * B->fsz[a] + ca] * B->r[i];
            rhs[B->foff[a] + ca] += s * sf[B->foff[a] + ca];
        }
        pa += B->nres * B->fsz[a];
    }
}

/* Per-landmark Schur elimination into (S, rhs).  Emax: scratch for E segments. Returns 0 if C_j is not PD. */
static int eliminate_landmark(const program* P, int j, const lm_work* W, double radius, double* S, double* rhs) {
    const kba_window* w = P->w;
    const int n = P->n_f;
    const double* se = W->scale_e + 3 * j;
    /* blocks of this landmark: its observations + optional gp block */
    const int o0 = w->lm_obs_ptr[j], o1 = w->lm_obs_ptr[j + 1];
    const int g = P->gp_of_lm[j];
    const int nb = (o1 - o0) + (g >= 0 ? 1 : 0);
    /* segments: distinct (offset,size) f parts touched */
    int maxseg = 3 * nb + 1;
    int* soff = (int*)malloc(sizeof(int) * maxseg);
    int* ssz = (int*)malloc(sizeof(int) * maxseg);
    double* E = (double*)calloc((size_t)maxseg * 18, sizeof(double)); /* each seg: sz x 3 */
    int nseg = 0;
    double C[9] = {0}, ge[3] = {0};
    for (int c = 0; c < 3; ++c) C[4 * c] = W->diag_e[3 * j + c] / radius; /* D_e^2 */
    for (int t = 0; t < nb; ++t) {
        const rblock* B = (t < o1 - o0) ? &P->blk[o0 + t] : &P->blk[P->n_obs + g];
        if (B->nres == 0 || B->lm != j) continue;
        double Je[9];
        for (int i = 0; i < B->nres; ++i)
            for (int c = 0; c < 3; ++c) Je[3 * i + c] = B->Je[3 * i + c] * se[c];
        for (int a = 0; a < 3; ++a) {
            for (int b = 0; b < 3; ++b) {
                double s = 0;
                for (int i = 0; i < B->nres; ++i) s += Je[3 * i + a] * Je[3 * i + b];
                C[3 * a + b] += s;
            }
            double s = 0;
            for (int i = 0; i < B->nres; ++i) s += Je[3 * i + a] * B->r[i];
            ge[a] += s;
        }
        int pa = 0;
        for (int a = 0; a < B->nf; ++a) {
            int sidx = -1;
            for (int q = 0; q < nseg; ++q)
                if (soff[q] == B->foff[a]) { sidx = q; break; }
            if (sidx < 0) { sidx = nseg++; soff[sidx] = B->foff[a]; ssz[sidx] = B->fsz[a]; }
            for (int ca = 0; ca < B->fsz[a]; ++ca)
                for (int c = 0; c < 3; ++c) {
                    double s = 0;
                    for (int i = 0; i < B->nres; ++i) s += B->Jf[pa + i * B->fsz[a] + ca] * Je[3 * i + c];
                    E[(size_t)sidx * 18 + 3 * ca + c] += s * W->scale_f[B->foff[a] + ca];
                }
            pa += B->nres * B->fsz[a];
        }
        add_ff(P, B, W->scale_f, S, rhs);
    }
    double Ci[9];
    int ok = inv_spd3(C, Ci);
    if (ok) {
        memcpy(W->Cinv + 9 * j, Ci, sizeof Ci);
        memcpy(W->ge + 3 * j, ge, sizeof ge);
        double Cg[3];
        matvec3(Ci, ge, Cg);
        for (int a = 0; a < nseg; ++a) {
            /* EC = E_a * Cinv (sz x 3) */
            double EC[18];
            for (int r = 0; r < ssz[a]; ++r)
                for (int c = 0; c < 3; ++c)
                    EC[3 * r + c] = E[(size_t)a * 18 + 3 * r] * Ci[c] + E[(size_t)a * 18 + 3 * r + 1] * Ci[3 + c] +
                                    E[(size_t)a * 18 + 3 * r + 2] * Ci[6 + c];
            for (int b = 0; b < nseg; ++b)
                for (int r = 0; r < ssz[a]; ++r)
                    for (int c = 0; c < ssz[b]; ++c) {
                        const double* Eb = E + (size_t)b * 18 + 3 * c;
                        S[(size_t)(soff[a] + r) * n + soff[b] + c] -= EC[3 * r] * Eb[0] + EC[3 * r + 1] * Eb[1] + EC[3 * r + 2] * Eb[2];
                    }
            for (int r = 0; r < ssz[a]; ++r)
                rhs[soff[a] + r] -= E[(size_t)a * 18 + 3 * r] * Cg[0] + E[(size_t)a * 18 + 3 * r + 1] * Cg[1] + E[(size_t)a * 18 + 3 * r + 2] * Cg[2];
        }
    }
    free(soff); free(ssz); free(E);
    return ok;
}

/* back-substitution y_e = Cinv (g_e - sum_a E_a^T y_f) for landmark j (scaled coordinates) */
static void backsub_landmark(const program* P, int j, const lm_work* W, double* ye) {
    const kba_window* w = P->w;
    const double* se = W->scale_e + 3 * j;
    const int o0 = w->lm_obs_ptr[j], o1 = w->lm_obs_ptr[j + 1];
    const int g = P->gp_of_lm[j];
    const int nb = (o1 - o0) + (g >= 0 ? 1 : 0);
    double t[3] = {W->ge[3 * j], W->ge[3 * j + 1], W->ge[3 * j + 2]};
    for (int q = 0; q < nb; ++q) {
        const rblock* B = (q < o1 - o0) ? &P->blk[o0 + q] : &P->blk[P->n_obs + g];
        if (B->nres == 0 || B->lm != j) continue;
        /* m_i = sum over f parts Jf_s y_f  (per residual row) */
        double m[3] = {0, 0, 0};
        int pa = 0;
        for (int a = 0; a < B->nf; ++a) {
            for (int i = 0; i < B->nres; ++i)
                for (int c = 0; c < B->fsz[a]; ++c)
                    m[i] += B->Jf[pa + i * B->fsz[a] + c] * W->scale_f[B->foff[a] + c] * W->step_f[B->foff[a] + c];
            pa += B->nres * B->fsz[a];
        }
        for (int c = 0; c < 3; ++c)
            for (int i = 0; i < B->nres; ++i) t[c] -= B->Je[3 * i + c] * se[c] * m[i];
    }
    matvec3(W->Cinv + 9 * j, t, ye);
}

/* Solve (J_s^T J_s + D^2) y = J_s^T r~ by Schur elimination; returns 0 on linear solver failure. */
static int compute_step(program* P, lm_work* W, double radius) {
    const int n = P->n_f;
    memset(W->S, 0, sizeof(double) * (size_t)(n > 0 ? n : 1) * (n > 0 ? n : 1));
    memset(W->rhs, 0, sizeof(double) * (n > 0 ? n : 1));
    int ok = 1;
    int nthreads = P->num_threads;
#ifdef _OPENMP
    if (nthreads > 1) {
        double* Sp = (double*)calloc((size_t)nthreads * ((size_t)n * n + n), sizeof(double));
#pragma omp parallel num_threads(nthreads)
        {
            const int tid = omp_get_thread_num();
            double* St = Sp + (size_t)tid * ((size_t)n * n + n);
            double* rt = St + (size_t)n * n;
            int okl = 1;
#pragma omp for schedule(dynamic, 16)
            for (int j = 0; j < P->n_lm; ++j)
                if (P->lm_in[j] && !eliminate_landmark(P, j, W, radius, St, rt)) okl = 0;
            if (!okl) {
#pragma omp atomic write
                ok = 0;
            }
        }
        for (int t = 0; t < nthreads; ++t) {
            const double* St = Sp + (size_t)t * ((size_t)n * n + n);
            for (size_t i = 0; i < (size_t)n * n; ++i) W->S[i] += St[i];
            for (int i = 0; i < n; ++i) W->rhs[i] += St[(size_t)n * n + i];
        }
        free(Sp);
    } else
#endif
    {
        (void)nthreads;
        for (int j = 0; j < P->n_lm; ++j)
            if (P->lm_in[j] && !eliminate_landmark(P, j, W, radius, W->S, W->rhs)) ok = 0;
    }
    if (!ok) return 0;
    /* blocks without an e-block: regularisers, and every block when the landmarks are constant */
    for (int b = 0; b < P->n_blocks; ++b) {
        const rblock* B = &P->blk[b];
        if (B->nres == 0 || B->lm >= 0) continue;
        add_ff(P, B, W->scale_f, W->S, W->rhs);
    }
    for (int i = 0; i < n; ++i) W->S[(size_t)i * n + i] += W->diag_f[i] / radius;
    if (n > 0) {
        if (!cholesky_lower(W->S, n)) return 0;
        memcpy(W->step_f, W->rhs, sizeof(double) * n);
        cholesky_solve(W->S, n, W->step_f);
    }
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(P->num_threads) if (P->num_threads > 1)
#endif
    for (int j = 0; j < P->n_lm; ++j)
        if (P->lm_in[j]) backsub_landmark(P, j, W, W->step_e + 3 * j);
    /* [ceres] LevenbergMarquardtStrategy: solved J y = r, the step is -y */
    for (int i = 0; i < n; ++i) {
        if (!isfinite(W->step_f[i])) ok = 0;
        W->step_f[i] = -W->step_f[i];
    }
    for (int j = 0; j < P->n_lm; ++j)
        if (P->lm_in[j])
            for (int c = 0; c < 3; ++c) {
                if (!isfinite(W->step_e[3 * j + c])) ok = 0;
                W->step_e[3 * j + c] = -W->step_e[3 * j + c];
            }
    return ok;
}

/* model_cost_change = -(J_s y)^T (r + J_s y / 2)  (trust_region_minimizer.cc ComputeTrustRegionStep) */
static double model_cost_change(const program* P, const lm_work* W) {
    double acc = 0.0;
    for (int b = 0; b < P->n_blocks; ++b) {
        const rblock* B = &P->blk[b];
        if (B->nres == 0) continue;
        double m[3] = {0, 0, 0};
        int pa = 0;
        for (int a = 0; a < B->nf; ++a) {
            for (int i = 0; i < B->nres; ++i)
                for (int c = 0; c < B->fsz[a]; ++c)
                    m[i] += B->Jf[pa + i * B->fsz[a] + c] * W->scale_f[B->foff[a] + c] * W->step_f[B->foff[a] + c];
            pa += B->nres * B->fsz[a];
        }
        if (B->lm >= 0)
            for (int i = 0; i < B->nres; ++i)
                for (int c = 0; c < 3; ++c)
                    m[i] += B->Je[3 * i + c] * W->scale_e[3 * B->lm + c] * W->step_e[3 * B->lm + c];
        for (int i = 0; i < B->nres; ++i) acc += m[i] * (B->r[i] + 0.5 * m[i]);
    }
    return -acc;
}

/* x_plus = Plus(x, delta) over all variable blocks; delta given per f offset / landmark */
static void state_plus(const program* P, const state_t* x, const double* df, const double* de, state_t* out) {
    state_copy(out, x, P->n_kf, P->n_lm);
    for (int k = 0; k < P->n_kf; ++k) {
        if (P->off_pose[k] >= 0) kbo_pose_plus(x->pose + 7 * k, df + P->off_pose[k], out->pose + 7 * k);
        if (P->off_dir[k] >= 0) kbo_dir_plus(x->plane + 4 * k, df + P->off_dir[k], out->plane + 4 * k);
        if (P->off_dist[k] >= 0) out->plane[4 * k + 3] = x->plane[4 * k + 3] + df[P->off_dist[k]];
    }
    for (int j = 0; j < P->n_lm; ++j)
        if (P->lm_in[j])
            for (int c = 0; c < 3; ++c) out->lm[3 * j + c] = x->lm[3 * j + c] + de[3 * j + c];
}

/* squared norm of (a - b) and max-abs over the variable blocks, ambient coordinates */
static void state_diff_norms(const program* P, const state_t* a, const state_t* b, double* sq, double* mx) {
    double s = 0, m = 0;
#define ACC(v) do { const double d_ = (v); s += d_ * d_; if (fabs(d_) > m) m = fabs(d_); } while (0)
    for (int k = 0; k < P->n_kf; ++k) {
        if (P->off_pose[k] >= 0) for (int i = 0; i < 7; ++i) ACC(a->pose[7 * k + i] - (b ? b->pose[7 * k + i] : 0.0));
        if (P->off_dir[k] >= 0) for (int i = 0; i < 3; ++i) ACC(a->plane[4 * k + i] - (b ? b->plane[4 * k + i] : 0.0));
        if (P->off_dist[k] >= 0) ACC(a->plane[4 * k + 3] - (b ? b->plane[4 * k + 3] : 0.0));
    }
    for (int j = 0; j < P->n_lm; ++j)
        if (P->lm_in[j]) for (int i = 0; i < 3; ++i) ACC(a->lm[3 * j + i] - (b ? b->lm[3 * j + i] : 0.0));
#undef ACC
    *sq = s;
    if (mx) *mx = m;
}

typedef struct {
    kba_iteration* log;
    int cap, n;
} iter_log;

static void log_iteration(iter_log* L, int solve_index, int it, double cost, double cost_change, double gmax,
                          double step_norm, double rel, double radius, int valid, int success) {
    if (!L || !L->log || L->n >= L->cap) return;
    kba_iteration* e = &L->log[L->n++];
    e->cost = cost; e->cost_change = cost_change; e->gradient_max_norm = gmax; e->step_norm = step_norm;
    e->relative_decrease = rel; e->trust_region_radius = radius; e->iteration = it; e->solve_index = solve_index;
    e->step_is_valid = valid; e->step_is_successful = success;
}

static double now_sec(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* [ceres] TrustRegionMinimizer::Minimize with LevenbergMarquardtStrategy, Jacobi scaling, monotonic steps
 * (SURVEY A.6).  x is updated in place to the accepted iterate. */
static void ceres_solve(program* P, state_t* x, int max_num_iterations, double max_time, kba_solve_summary* sum,
                        iter_log* L, int solve_index) {
    const kba_options* opt = P->opt;
    const double t_start = now_sec();
    program_layout(P);
    const int n = P->n_f, nl = P->n_lm;
    memset(sum, 0, sizeof *sum);
    for (int j = 0; j < nl; ++j) sum->num_landmarks += P->lm_in[j];

    lm_work W;
    const size_t nn = (size_t)(n > 0 ? n : 1), ne = (size_t)3 * (nl > 0 ? nl : 1);
    W.scale_f = (double*)calloc(nn, sizeof(double)); W.scale_e = (double*)calloc(ne, sizeof(double));
    W.diag_f = (double*)calloc(nn, sizeof(double));  W.diag_e = (double*)calloc(ne, sizeof(double));
    W.grad_f = (double*)calloc(nn, sizeof(double));  W.grad_e = (double*)calloc(ne, sizeof(double));
    W.step_f = (double*)calloc(nn, sizeof(double));  W.step_e = (double*)calloc(ne, sizeof(double));
    W.S = (double*)calloc(nn * nn, sizeof(double));  W.rhs = (double*)calloc(nn, sizeof(double));
    W.Cinv = (double*)calloc(3 * ne, sizeof(double)); W.ge = (double*)calloc(ne, sizeof(double));
    double* cn_f = (double*)calloc(nn, sizeof(double)), *cn_e = (double*)calloc(ne, sizeof(double));
    double* delta_f = (double*)calloc(nn, sizeof(double)), *delta_e = (double*)calloc(ne, sizeof(double));
    state_t cand, gstep;
    state_alloc(&cand, P->n_kf, nl);
    state_alloc(&gstep, P->n_kf, nl);

    double x_cost = 0, radius = opt->initial_trust_region_radius, decrease_factor = 2.0;
    int reuse_diagonal = 0, num_invalid = 0, iteration = 0, step_successful = 0;
    double gmax = 0, x_norm = 0;

    /* ---- IterationZero: EvaluateGradientAndJacobian ---- */
    if (!program_evaluate(P, x, 1, &x_cost)) {
        sum->termination = KBA_TERM_FAILURE; /* "Residual and Jacobian evaluation failed." */
        sum->initial_cost = sum->final_cost = -1.0;
        goto done;
    }
    column_norms_and_gradient(P, cn_f, cn_e, W.grad_f, W.grad_e);
    for (int i = 0; i < n; ++i) W.scale_f[i] = 1.0 / (1.0 + sqrt(cn_f[i])); /* jacobi_scaling, computed once */
    for (int j = 0; j < nl; ++j)
        for (int c = 0; c < 3; ++c) W.scale_e[3 * j + c] = 1.0 / (1.0 + sqrt(cn_e[3 * j + c]));
    {
        double sq;
        state_diff_norms(P, x, NULL, &sq, NULL);
        x_norm = sqrt(sq);
        for (int i = 0; i < n; ++i) delta_f[i] = -W.grad_f[i];
        for (size_t i = 0; i < ne; ++i) delta_e[i] = -W.grad_e[i];
        state_plus(P, x, delta_f, delta_e, &gstep);
        state_diff_norms(P, x, &gstep, &sq, &gmax);
    }
    sum->initial_cost = x_cost;
    sum->final_cost = x_cost;
    sum->num_residual_blocks = 0;
    /* an observation with lidar depth is two ceres residual blocks (depth + reprojection, cpp:587-620) */
    for (int b = 0; b < P->n_blocks; ++b)
        sum->num_residual_blocks += (P->blk[b].nres > 0) + (b < P->n_obs && P->blk[b].nres == 3);
    log_iteration(L, solve_index, 0, x_cost, 0, gmax, 0, 0, radius, 0, 0);

    for (;;) {
        /* ---- FinalizeIterationAndCheckIfMinimizerCanContinue ---- */
        if (max_time > 0 && now_sec() - t_start >= max_time) { sum->termination = KBA_TERM_NO_CONVERGENCE; break; }
        if (iteration >= max_num_iterations) { sum->termination = KBA_TERM_NO_CONVERGENCE; break; }
        if (step_successful && gmax <= opt->gradient_tolerance) { sum->termination = KBA_TERM_CONVERGENCE; break; }
        if (radius <= opt->min_trust_region_radius) { sum->termination = KBA_TERM_CONVERGENCE; break; }
        iteration++;
        step_successful = 0;
        sum->num_iterations = iteration;

        /* ---- ComputeTrustRegionStep (LevenbergMarquardtStrategy::ComputeStep) ---- */
        if (!reuse_diagonal) {
            column_norms_and_gradient(P, cn_f, cn_e, W.grad_f, W.grad_e);
            for (int i = 0; i < n; ++i)
                W.diag_f[i] = fmin(fmax(cn_f[i] * W.scale_f[i] * W.scale_f[i], opt->min_lm_diagonal), opt->max_lm_diagonal);
            for (size_t i = 0; i < ne; ++i)
                W.diag_e[i] = fmin(fmax(cn_e[i] * W.scale_e[i] * W.scale_e[i], opt->min_lm_diagonal), opt->max_lm_diagonal);
        }
        reuse_diagonal = 1;
        int solved = compute_step(P, &W, radius);
        double model_change = 0;
        int valid = 0;
        if (solved) {
            model_change = model_cost_change(P, &W);
            valid = model_change > 0.0;
        }
        if (!valid) {
            /* HandleInvalidStep */
            if (++num_invalid >= opt->max_consecutive_invalid_steps) { sum->termination = KBA_TERM_FAILURE; break; }
            radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = 1; /* StepIsInvalid -> StepRejected(0) */
            log_iteration(L, solve_index, iteration, x_cost, 0, gmax, 0, 0, radius, 0, 0);
            continue;
        }
        num_invalid = 0;
        for (int i = 0; i < n; ++i) delta_f[i] = W.step_f[i] * W.scale_f[i]; /* undo the Jacobi scaling */
        for (size_t i = 0; i < ne; ++i) delta_e[i] = W.step_e[i] * W.scale_e[i];

        /* ---- ComputeCandidatePointAndEvaluateCost ---- */
        state_plus(P, x, delta_f, delta_e, &cand);
        double cand_cost;
        if (!program_evaluate(P, &cand, 0, &cand_cost)) cand_cost = DBL_MAX;

        /* ---- ParameterToleranceReached / FunctionToleranceReached ----
         * Order of ceres 1.13: both tests look at the CANDIDATE and, when one fires, the solve ends with x
         * unchanged.  g_tolerance_order = 1 is the other order SURVEY A.6 names for ceres <= 1.12 ("as recalled"): the
         * same tests, but a candidate that passes the acceptance test is applied before the solve ends.  The reference pins
         * neither (no golden vectors, Ceres not installable); scripts/ceres_order_sensitivity.py measures how far the two
         * move the result -- that is the error bar on "parity with Ceres" this restatement carries. */
        double sq;
        state_diff_norms(P, x, &cand, &sq, NULL);
        const double step_norm = sqrt(sq);
        const double cost_change = x_cost - cand_cost;
        const int tol_param = step_norm <= opt->parameter_tolerance * (x_norm + opt->parameter_tolerance);
        const int tol_func = !tol_param && fabs(cost_change) <= opt->function_tolerance * x_cost;
        if (tol_param || tol_func) {
            sum->termination = KBA_TERM_CONVERGENCE;
            if (g_tolerance_order == 1 && cost_change / model_change > opt->min_relative_decrease) {
                state_copy(x, &cand, P->n_kf, nl);
                if (!program_evaluate(P, x, 1, &x_cost)) { sum->termination = KBA_TERM_FAILURE; break; }
                sum->num_successful_steps++;
                if (x_cost < sum->final_cost) sum->final_cost = x_cost;
                log_iteration(L, solve_index, iteration, x_cost, cost_change, gmax, step_norm, cost_change / model_change, radius, 1, 1);
            } else {
                log_iteration(L, solve_index, iteration, x_cost, tol_param ? 0 : cost_change, gmax, step_norm, 0, radius, 1, 0);
            }
            break;
        }
        /* ---- IsStepSuccessful (monotonic: relative == historical decrease) ---- */
        const double rel = cost_change / model_change;
        if (rel > opt->min_relative_decrease) {
            /* HandleSuccessfulStep */
            state_copy(x, &cand, P->n_kf, nl);
            state_diff_norms(P, x, NULL, &sq, NULL);
            x_norm = sqrt(sq);
            if (!program_evaluate(P, x, 1, &x_cost)) { sum->termination = KBA_TERM_FAILURE; break; }
            column_norms_and_gradient(P, cn_f, cn_e, W.grad_f, W.grad_e);
            for (int i = 0; i < n; ++i) delta_f[i] = -W.grad_f[i];
            for (size_t i = 0; i < ne; ++i) delta_e[i] = -W.grad_e[i];
            state_plus(P, x, delta_f, delta_e, &gstep);
            state_diff_norms(P, x, &gstep, &sq, &gmax);
            step_successful = 1;
            sum->num_successful_steps++;
            /* LevenbergMarquardtStrategy::StepAccepted */
            radius = radius / fmax(1.0 / 3.0, 1.0 - pow(2.0 * rel - 1.0, 3));
            radius = fmin(opt->max_trust_region_radius, radius);
            decrease_factor = 2.0;
            reuse_diagonal = 0;
            if (x_cost < sum->final_cost) sum->final_cost = x_cost;
            log_iteration(L, solve_index, iteration, x_cost, cost_change, gmax, step_norm, rel, radius, 1, 1);
        } else {
            /* HandleUnsuccessfulStep -> StepRejected */
            radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = 1;
            log_iteration(L, solve_index, iteration, cand_cost, cost_change, gmax, step_norm, rel, radius, 1, 0);
        }
    }
done:
    free(W.scale_f); free(W.scale_e); free(W.diag_f); free(W.diag_e); free(W.grad_f); free(W.grad_e);
    free(W.step_f); free(W.step_e); free(W.S); free(W.rhs); free(W.Cinv); free(W.ge);
    free(cn_f); free(cn_e); free(delta_f); free(delta_e);
    state_free(&cand); state_free(&gstep);
}

/* ------------------------------------------------------------------------------------------------ */
/* trimming (robust_solving.cpp:67-125, trimmer_quantile.hpp:40-63)                                   */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { double v; int i; } vi_pair;
static int vi_cmp(const void* a, const void* b) {
    const vi_pair* x = (const vi_pair*)a, *y = (const vi_pair*)b;
    if (x->v < y->v) return -1;
    if (x->v > y->v) return 1;
    return (x->i > y->i) - (x->i < y->i);
}
int kbo_trimmer_quantile(const double* values, int n, double q, unsigned char* rejected) {
    vi_pair* p = (vi_pair*)malloc(sizeof(vi_pair) * (n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) { p[i].v = values[i]; p[i].i = i; rejected[i] = 0; }
    qsort(p, n, sizeof(vi_pair), vi_cmp);
    const int num = (int)((double)n * q); /* trimmer_quantile.hpp:48 */
    for (int i = num; i < n; ++i) rejected[p[i].i] = 1;
    free(p);
    return n - (num < n ? num : n);
}

/* TrimmerFix::getOutliers (robust_optimization/include/robust_optimization/internal/trimmer_fix.hpp:38-47): residuals above a
 * fixed threshold.  Not reachable from BundleAdjusterKeyframes::solve() (which uses the quantile trimmer, cpp:745-758); restated
 * for the reference test Trimmers.TrimmerFix only. */
int kbo_trimmer_fix(const double* values, int n, double threshold, unsigned char* rejected) {
    int cnt = 0;
    for (int i = 0; i < n; ++i) { rejected[i] = values[i] > threshold; cnt += rejected[i]; }
    return cnt;
}

/* one rejection pass over the three residual groups; marks landmarks to remove */
static void trim_round(program* P, const state_t* x, unsigned char* remove) {
    const kba_window* w = P->w;
    const kba_options* opt = P->opt;
    const int nl = P->n_lm;
    double* mx[3];
    for (int g = 0; g < 3; ++g) {
        mx[g] = (double*)malloc(sizeof(double) * (nl > 0 ? nl : 1));
        for (int j = 0; j < nl; ++j) mx[g][j] = -1.0; /* -1: landmark has no block in this group */
    }
    /* calculateResiduals(apply_loss=false) -> block norms -> per-landmark maximum (robust_solving.cpp:16-91) */
    for (int j = 0; j < nl; ++j) {
        if (!P->lm_active[j]) continue;
        for (int o = w->lm_obs_ptr[j]; o < w->lm_obs_ptr[j + 1]; ++o) {
            double r[3], c = 0, raw[2];
            if (!eval_obs(P, x, o, j, 0, &c, r, NULL, NULL, raw)) continue;
            if (raw[1] >= 0 && raw[1] > mx[0][j]) mx[0][j] = raw[1]; /* group 0: depth */
            if (raw[0] > mx[1][j]) mx[1][j] = raw[0];                /* group 1: reprojection */
        }
        const int g = P->gp_of_lm[j];
        if (g >= 0) {
            double r[1];
            const int k = w->gp_kf[g];
            kbo_gp_height(x->pose + 7 * k, x->plane + 4 * k, x->plane[4 * k + 3], x->lm + 3 * j, r, NULL, NULL, NULL, NULL);
            mx[2][j] = fabs(r[0]);
        }
    }
    const double quant[3] = {opt->depth_quantile, opt->reprojection_quantile, opt->gp_quantile};
    double* vals = (double*)malloc(sizeof(double) * (nl > 0 ? nl : 1));
    int* idx = (int*)malloc(sizeof(int) * (nl > 0 ? nl : 1));
    unsigned char* rej = (unsigned char*)malloc(nl > 0 ? nl : 1);
    for (int g = 0; g < 3; ++g) {
        int n = 0;
        for (int j = 0; j < nl; ++j)
            if (mx[g][j] >= 0) { vals[n] = mx[g][j]; idx[n] = j; n++; }
        if (n == 0 || n < opt->min_residual_groups) continue; /* robust_solving.cpp:19-21,109-111 */
        kbo_trimmer_quantile(vals, n, quant[g], rej);
        for (int i = 0; i < n; ++i)
            if (rej[i]) remove[idx[i]] = 1;
    }
    free(vals); free(idx); free(rej);
    for (int g = 0; g < 3; ++g) free(mx[g]);
}

/* ------------------------------------------------------------------------------------------------ */
/* public entry points                                                                               */
/* ------------------------------------------------------------------------------------------------ */
void kbo_default_options(kba_options* o) {
    memset(o, 0, sizeof *o);
    o->depth_thres = 0.16; o->reprojection_thres = 1.6; /* bundle_adjuster_keyframes.hpp:79-89 */
    o->depth_quantile = 0.95; o->reprojection_quantile = 0.95; o->gp_quantile = 1.0; o->gp_huber = 0.1;
    o->num_trim_rounds = -1; o->trim_solver_iterations = 2; o->final_solver_iterations = 100;
    o->min_landmarks_for_trimming = 100; o->min_residual_groups = 30; o->num_rounds_option = 1;
    o->solver_time_sec = 20.0;
    o->function_tolerance = 1e-6; o->gradient_tolerance = 1e-10; o->parameter_tolerance = 1e-8;
    o->initial_trust_region_radius = 1e4; o->max_trust_region_radius = 1e16; o->min_trust_region_radius = 1e-32;
    o->min_relative_decrease = 1e-3; o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32;
    o->max_consecutive_invalid_steps = 5; o->precision = 0;
}

static int count_reg_blocks(const kba_window* w) {
    int n = 0;
    if (w->scale_weight > 0) n += 1;
    if (w->plane_reg_weight > 0 && w->n_kf > 1) n += 3 * (w->n_kf - 1) + w->n_kf;
    if (w->speed_weight > 0) n += 1;
    return n;
}

static int program_init(program* P, const kba_window* w, const kba_options* opt, int num_threads) {
    memset(P, 0, sizeof *P);
    P->w = w; P->opt = opt;
    P->n_kf = w->n_kf; P->n_lm = w->n_lm; P->n_obs = w->n_obs; P->n_gp = w->n_gp;
    const int K = w->n_kf > 0 ? w->n_kf : 1, NL = w->n_lm > 0 ? w->n_lm : 1;
    P->lm_active = (unsigned char*)malloc(NL);
    memset(P->lm_active, 1, NL);
    P->gp_of_lm = (int*)malloc(sizeof(int) * NL);
    for (int j = 0; j < w->n_lm; ++j) P->gp_of_lm[j] = -1;
    for (int g = 0; g < w->n_gp; ++g) P->gp_of_lm[w->gp_lm[g]] = g;
    P->off_pose = (int*)malloc(sizeof(int) * K); P->off_dir = (int*)malloc(sizeof(int) * K);
    P->off_dist = (int*)malloc(sizeof(int) * K);
    P->pose_in = (unsigned char*)malloc(K); P->dir_in = (unsigned char*)malloc(K); P->dist_in = (unsigned char*)malloc(K);
    P->lm_in = (unsigned char*)malloc(NL);
    P->n_reg = count_reg_blocks(w);
    P->n_blocks = w->n_obs + w->n_gp + P->n_reg;
    P->blk = (rblock*)calloc((size_t)(P->n_blocks > 0 ? P->n_blocks : 1), sizeof(rblock));
#ifdef _OPENMP
    P->num_threads = num_threads > 0 ? num_threads : omp_get_max_threads();
#else
    (void)num_threads;
    P->num_threads = 1;
#endif
    return P->blk != NULL;
}
static void program_free(program* P) {
    free(P->lm_active); free(P->gp_of_lm); free(P->off_pose); free(P->off_dir); free(P->off_dist);
    free(P->pose_in); free(P->dir_in); free(P->dist_in); free(P->lm_in); free(P->blk);
}

static int window_valid(const kba_window* w) {
    if (!w || w->n_kf < 0 || w->n_lm < 0 || w->n_obs < 0 || w->n_gp < 0 || w->n_cam < 1) return 0;
    if (!w->kf_pose || !w->kf_fixed || !w->cam_intr || !w->cam_pose) return 0;
    if (w->n_lm > 0 && (!w->lm_pos || !w->lm_weight || !w->lm_obs_ptr)) return 0;
    if (w->n_obs > 0 && (!w->obs_kf || !w->obs_u || !w->obs_v || !w->obs_d)) return 0;
    if (w->n_gp > 0 && (!w->gp_lm || !w->gp_kf || !w->gp_weight || !w->kf_plane)) return 0;
    if (w->plane_reg_weight > 0 && !w->kf_plane) return 0;
    for (int o = 0; o < w->n_obs; ++o) {
        if (w->obs_kf[o] < 0 || w->obs_kf[o] >= w->n_kf) return 0;
        if (w->obs_cam && (w->obs_cam[o] < 0 || w->obs_cam[o] >= w->n_cam)) return 0;
    }
    return 1;
}

static void state_from_window(state_t* x, const kba_window* w) {
    memcpy(x->pose, w->kf_pose, sizeof(double) * 7 * w->n_kf);
    if (w->kf_plane) memcpy(x->plane, w->kf_plane, sizeof(double) * 4 * w->n_kf);
    else for (int k = 0; k < w->n_kf; ++k) { x->plane[4 * k] = 0; x->plane[4 * k + 1] = 0; x->plane[4 * k + 2] = 1; x->plane[4 * k + 3] = 0; }
    if (w->n_lm > 0) memcpy(x->lm, w->lm_pos, sizeof(double) * 3 * w->n_lm);
}

int kbo_solve_window(const kba_window* w, const kba_options* opt, kba_result* res, int num_threads) {
    if (!window_valid(w) || !opt || !res) return KBA_ERR_BAD_ARG;
    const double t0 = now_sec();
    program P;
    program_init(&P, w, opt, num_threads);
    state_t x;
    state_alloc(&x, w->n_kf, w->n_lm);
    state_from_window(&x, w);
    iter_log L = {res->iterations, res->iterations ? res->iterations_capacity : 0, 0};
    res->num_solves = 0;

    /* number_iterations (cpp:740-745 / 864-869) */
    int rounds = opt->num_trim_rounds;
    if (rounds < 0) rounds = (w->n_lm > opt->min_landmarks_for_trimming) ? opt->num_rounds_option : 0;
    unsigned char* remove = (unsigned char*)calloc(w->n_lm > 0 ? w->n_lm : 1, 1);
    for (int r = 0; r < rounds && res->num_solves < KBA_MAX_SOLVES - 1; ++r) {
        kba_solve_summary s;
        ceres_solve(&P, &x, opt->trim_solver_iterations, opt->solver_time_sec, &s, &L, res->num_solves);
        if (s.initial_cost - s.final_cost <= 0.0) /* robust_solving.cpp:172-181 */
            ceres_solve(&P, &x, 3 * opt->trim_solver_iterations, opt->solver_time_sec, &s, &L, res->num_solves);
        res->solves[res->num_solves++] = s;
        memset(remove, 0, w->n_lm > 0 ? w->n_lm : 1);
        trim_round(&P, &x, remove);
        for (int j = 0; j < w->n_lm; ++j)
            if (remove[j]) P.lm_active[j] = 0; /* all residual blocks of the landmark, robust_solving.cpp:199-214 */
    }
    {
        kba_solve_summary s;
        ceres_solve(&P, &x, opt->final_solver_iterations, opt->solver_time_sec, &s, &L, res->num_solves);
        res->solves[res->num_solves++] = s;
    }
    res->initial_cost = res->solves[0].initial_cost;
    res->final_cost = res->solves[res->num_solves - 1].final_cost;
    res->num_iteration_records = L.n;
    if (res->kf_pose) memcpy(res->kf_pose, x.pose, sizeof(double) * 7 * w->n_kf);
    if (res->kf_plane) memcpy(res->kf_plane, x.plane, sizeof(double) * 4 * w->n_kf);
    if (res->lm_pos && w->n_lm > 0) memcpy(res->lm_pos, x.lm, sizeof(double) * 3 * w->n_lm);
    if (res->lm_rejected) for (int j = 0; j < w->n_lm; ++j) res->lm_rejected[j] = !P.lm_active[j];
    res->status = KBA_OK;
    res->time_sec = now_sec() - t0;
    free(remove);
    state_free(&x);
    program_free(&P);
    return KBA_OK;
}

int kbo_eval(const kba_window* w, const kba_options* opt, kba_eval_out* out) {
    if (!window_valid(w) || !opt || !out) return KBA_ERR_BAD_ARG;
    program P;
    program_init(&P, w, opt, 1);
    program_layout(&P);
    state_t x;
    state_alloc(&x, w->n_kf, w->n_lm);
    state_from_window(&x, w);
    double cost = 0;
    int failed = 0;
    for (int j = 0; j < w->n_lm; ++j)
        for (int o = w->lm_obs_ptr[j]; o < w->lm_obs_ptr[j + 1]; ++o) {
            double r[3] = {0, 0, 0}, Jp[18] = {0}, Jl[9] = {0};
            if (!eval_obs(&P, &x, o, j, 1, &cost, r, Jp, Jl, NULL)) { failed = 1; memset(Jp, 0, sizeof Jp); memset(Jl, 0, sizeof Jl); }
            if (w->kf_fixed[w->obs_kf[o]]) memset(Jp, 0, sizeof Jp);
            if (out->residual) memcpy(out->residual + 3 * (size_t)o, r, sizeof r);
            if (out->jac_pose) memcpy(out->jac_pose + 18 * (size_t)o, Jp, sizeof Jp);
            if (out->jac_lm) memcpy(out->jac_lm + 9 * (size_t)o, Jl, sizeof Jl);
        }
    if (out->cost) out->cost[0] = cost;
    if (out->failed) out->failed[0] = failed;
    state_free(&x);
    program_free(&P);
    return KBA_OK;
}

/* Triangulator::triangulate_rays (internal/triangulator.hpp:51-75): solve sum(I - r r^T) p = sum (I - r r^T) t. */
void kbo_triangulate_rays(int n, const double* R_oc, const double* t_oc, const double* rays, double out[3]) {
    double A[9] = {0}, b[3] = {0};
    for (int i = 0; i < n; ++i) {
        double r[3];
        matvec3(R_oc + 9 * i, rays + 3 * i, r);
        double M[9];
        for (int a = 0; a < 3; ++a)
            for (int c = 0; c < 3; ++c) M[3 * a + c] = (a == c ? 1.0 : 0.0) - r[a] * r[c];
        for (int q = 0; q < 9; ++q) A[q] += M[q];
        double Mt[3];
        matvec3(M, t_oc + 3 * i, Mt);
        for (int a = 0; a < 3; ++a) b[a] += Mt[a];
    }
    /* the reference solves with a Jacobi SVD; the matrix is symmetric PSD (PD for >= 2 non-parallel rays), so a
     * symmetric eigen-free solve via Cramer's rule gives the same least-squares solution in the PD case */
    const double det = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
    double inv[9];
    inv[0] = (A[4] * A[8] - A[5] * A[7]) / det; inv[1] = (A[2] * A[7] - A[1] * A[8]) / det; inv[2] = (A[1] * A[5] - A[2] * A[4]) / det;
    inv[3] = (A[5] * A[6] - A[3] * A[8]) / det; inv[4] = (A[0] * A[8] - A[2] * A[6]) / det; inv[5] = (A[2] * A[3] - A[0] * A[5]) / det;
    inv[6] = (A[3] * A[7] - A[4] * A[6]) / det; inv[7] = (A[1] * A[6] - A[0] * A[7]) / det; inv[8] = (A[0] * A[4] - A[1] * A[3]) / det;
    matvec3(inv, b, out);
}

/* Landmark initialisation of BundleAdjusterKeyframes::push() for every landmark of a window (batch restatement, see
 * kba_init_landmarks in include/kba_b200.h): back-projection of the first observation with a lidar depth
 * (bundle_adjuster_keyframes.cpp:332-355), else Triangulator::triangulate_rays over all viewing rays when there are at
 * least two (cpp:125-159,358-382), then the cheirality test (landmark_selection_scheme_cheirality.cpp:22-60). */
void kbo_init_landmarks(const kba_window* w, double* lm_pos_out, unsigned char* flags_out) {
    for (int j = 0; j < w->n_lm; ++j) {
        const int o0 = w->lm_obs_ptr[j], o1 = w->lm_obs_ptr[j + 1];
        double p[3] = {0, 0, 0};
        int created = 0;
        for (int o = o0; o < o1 && !created; ++o) {
            if (w->obs_d[o] < 0.f) continue;
            const int c = w->obs_cam ? w->obs_cam[o] : 0;
            const double* pose = w->kf_pose + 7 * w->obs_kf[o];
            const double* cam = w->cam_pose + 7 * c;
            const double f = w->cam_intr[3 * c], cx = w->cam_intr[3 * c + 1], cy = w->cam_intr[3 * c + 2];
            double R[9], Rc[9];
            quat_to_R(pose, R);
            quat_to_R(cam, Rc);
            const double z = (double)w->obs_d[o];
            const double pc[3] = {((double)w->obs_u[o] - cx) * z / f, ((double)w->obs_v[o] - cy) * z / f, z};
            /* (T_cam_vehicle * T_kf)^-1 * pc = R^T (Rc^T (pc - tc) - t) */
            double a[3], b[3];
            for (int i = 0; i < 3; ++i) a[i] = pc[i] - cam[4 + i];
            for (int i = 0; i < 3; ++i) b[i] = Rc[i] * a[0] + Rc[3 + i] * a[1] + Rc[6 + i] * a[2] - pose[4 + i];
            for (int i = 0; i < 3; ++i) p[i] = R[i] * b[0] + R[3 + i] * b[1] + R[6 + i] * b[2];
            created = 1;
        }
        if (!created && o1 - o0 >= 2) {
            const int n = o1 - o0;
            double* R_oc = (double*)malloc(sizeof(double) * 9 * n);
            double* t_oc = (double*)malloc(sizeof(double) * 3 * n);
            double* rays = (double*)malloc(sizeof(double) * 3 * n);
            for (int q = 0; q < n; ++q) {
                const int o = o0 + q, c = w->obs_cam ? w->obs_cam[o] : 0;
                const double* pose = w->kf_pose + 7 * w->obs_kf[o];
                const double* cam = w->cam_pose + 7 * c;
                const double f = w->cam_intr[3 * c], cx = w->cam_intr[3 * c + 1], cy = w->cam_intr[3 * c + 2];
                double R[9], Rc[9];
                quat_to_R(pose, R);
                quat_to_R(cam, Rc);
                double r[3] = {((double)w->obs_u[o] - cx) / f, ((double)w->obs_v[o] - cy) / f, 1.0};
                const double nr = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
                for (int i = 0; i < 3; ++i) rays[3 * q + i] = r[i] / nr;
                for (int i = 0; i < 3; ++i)  /* R_oc = R^T Rc^T */
                    for (int k = 0; k < 3; ++k)
                        R_oc[9 * q + 3 * i + k] = R[i] * Rc[3 * k] + R[3 + i] * Rc[3 * k + 1] + R[6 + i] * Rc[3 * k + 2];
                double a[3];
                for (int i = 0; i < 3; ++i) a[i] = Rc[i] * cam[4] + Rc[3 + i] * cam[5] + Rc[6 + i] * cam[6] + pose[4 + i];
                for (int i = 0; i < 3; ++i) t_oc[3 * q + i] = -(R[i] * a[0] + R[3 + i] * a[1] + R[6 + i] * a[2]);
            }
            kbo_triangulate_rays(n, R_oc, t_oc, rays, p);
            free(R_oc); free(t_oc); free(rays);
            created = 1;
        }
        int front = created;
        for (int o = o0; o < o1 && front; ++o) {
            const int c = w->obs_cam ? w->obs_cam[o] : 0;
            const double* pose = w->kf_pose + 7 * w->obs_kf[o];
            const double* cam = w->cam_pose + 7 * c;
            double R[9], Rc[9], x[3];
            quat_to_R(pose, R);
            quat_to_R(cam, Rc);
            for (int i = 0; i < 3; ++i) x[i] = R[3 * i] * p[0] + R[3 * i + 1] * p[1] + R[3 * i + 2] * p[2] + pose[4 + i];
            const double zc = Rc[6] * x[0] + Rc[7] * x[1] + Rc[8] * x[2] + cam[6];
            if (zc < 0.0) front = 0;
        }
        lm_pos_out[3 * j] = p[0]; lm_pos_out[3 * j + 1] = p[1]; lm_pos_out[3 * j + 2] = p[2];
        flags_out[j] = (unsigned char)(created | (front << 1));
    }
}

