/*
 * lidar_oracle.c -- CPU restatement of the lidar depth extraction (BASELINE config 4).  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED: the reference code for this step lives in the un-vendored repository johannes-graeter/mono_lidar_depth
 * (install_repos.sh:9, docker/src/Dockerfile:72) and is absent from /root/reference, together with any test or sample
 * data.  This file follows the only in-tree specification, the parameter file
 * demo_keyframe_bundle_adjustment_meta/res/mono_lidar_fusion_parameters.yaml (line numbers below), for the default
 * (non-ground) feature path: pixel-rectangle neighbour search (:5-18), histogram segmentation by depth (:52-63),
 * largest-triangle plane (:167-178), view-ray intersection, global (:97-104) and local relative (:106-114) depth gates,
 * points behind the camera cut (:166).  The RANSAC ground-plane branch (:125-162) is not restated.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "kba_oracle.h"

void kbo_lidar_default_options(kba_lidar_options* o) {
    memset(o, 0, sizeof *o);
    o->image_width = 1242; o->image_height = 375;
    o->rect_width = 6; o->rect_height = 9; o->rect_offset_x = 0; o->rect_offset_y = 0;
    o->hist_bin_width = 0.3; o->hist_min_count = 1; o->min_points = 3;
    o->depth_min = 0; o->depth_max = 100; o->local_rel_tolerance = 0.5;
    o->triangle_crossnorm_min = 0.1; o->viewray_plane_min = 0.1;
}

static void quat_R(const double q[4], double R[9]) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}

typedef struct { float u, v, x, y, z; int idx; } proj_pt;

/* depth of one feature from its neighbour list; shared by every implementation of the specification:
 * all arithmetic in float (the inputs are float), comparisons order-independent (ties by original point index) */
static float feature_depth(const proj_pt* nb, int m, float fu, float fv, const double* intr, const kba_lidar_options* o) {
    if (m < o->min_points) return -1.0f;
    float zmin = nb[0].z, zmax = nb[0].z;
    for (int i = 1; i < m; ++i) { if (nb[i].z < zmin) zmin = nb[i].z; if (nb[i].z > zmax) zmax = nb[i].z; }
    /* histogram segmentation: bins of hist_bin_width from the nearest depth; the nearest local maximum wins */
    enum { NB = 64 };
    int cnt[NB];
    memset(cnt, 0, sizeof cnt);
    const float bw = (float)o->hist_bin_width;
    for (int i = 0; i < m; ++i) {
        int b = (int)floorf((nb[i].z - zmin) / bw);
        if (b > NB - 1) b = NB - 1;
        cnt[b]++;
    }
    int sel = -1;
    for (int b = 0; b < NB; ++b) {
        if (cnt[b] < o->hist_min_count || cnt[b] == 0) continue;
        const int left = b > 0 ? cnt[b - 1] : -1, right = b < NB - 1 ? cnt[b + 1] : -1;
        if (cnt[b] > left && cnt[b] >= right) { sel = b; break; }
    }
    if (sel < 0) return -1.0f;
    /* largest triangle among the points of the selected bin (ties: smallest original indices) */
    int best[3] = {-1, -1, -1};
    float best_area = -1.0f;
    int bi0 = 0, bi1 = 0, bi2 = 0;
    int kept = 0;
    for (int i = 0; i < m; ++i) {
        int b = (int)floorf((nb[i].z - zmin) / bw); if (b > NB - 1) b = NB - 1;
        kept += (b == sel);
    }
    if (kept < o->min_points) return -1.0f;
    for (int i = 0; i < m; ++i) {
        int b = (int)floorf((nb[i].z - zmin) / bw); if (b > NB - 1) b = NB - 1;
        if (b != sel) continue;
        for (int j = 0; j < m; ++j) {
            if (nb[j].idx <= nb[i].idx) continue;
            int bj = (int)floorf((nb[j].z - zmin) / bw); if (bj > NB - 1) bj = NB - 1;
            if (bj != sel) continue;
            for (int k = 0; k < m; ++k) {
                if (nb[k].idx <= nb[j].idx) continue;
                int bk = (int)floorf((nb[k].z - zmin) / bw); if (bk > NB - 1) bk = NB - 1;
                if (bk != sel) continue;
                const float ax = nb[j].x - nb[i].x, ay = nb[j].y - nb[i].y, az = nb[j].z - nb[i].z;
                const float bx = nb[k].x - nb[i].x, by = nb[k].y - nb[i].y, bz = nb[k].z - nb[i].z;
                const float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
                const float area = cx * cx + cy * cy + cz * cz;
                int better = area > best_area;
                if (!better && area == best_area) {
                    const int a0 = nb[i].idx, a1 = nb[j].idx, a2 = nb[k].idx;
                    better = (a0 < bi0) || (a0 == bi0 && (a1 < bi1 || (a1 == bi1 && a2 < bi2)));
                }
                if (better) { best_area = area; best[0] = i; best[1] = j; best[2] = k; bi0 = nb[i].idx; bi1 = nb[j].idx; bi2 = nb[k].idx; }
            }
        }
    }
    if (best[0] < 0) return -1.0f;
    const proj_pt* A = &nb[best[0]], *B = &nb[best[1]], *Cc = &nb[best[2]];
    /* planarity: sine of every inner angle (cross product of the normalised edges) at least the threshold */
    const float e[3][3] = {{B->x - A->x, B->y - A->y, B->z - A->z}, {Cc->x - B->x, Cc->y - B->y, Cc->z - B->z}, {A->x - Cc->x, A->y - Cc->y, A->z - Cc->z}};
    float len[3];
    for (int q = 0; q < 3; ++q) len[q] = sqrtf(e[q][0] * e[q][0] + e[q][1] * e[q][1] + e[q][2] * e[q][2]);
    for (int q = 0; q < 3; ++q) {
        const int r = (q + 1) % 3;
        if (!(len[q] > 0.0f) || !(len[r] > 0.0f)) return -1.0f;
        const float cx = e[q][1] * e[r][2] - e[q][2] * e[r][1], cy = e[q][2] * e[r][0] - e[q][0] * e[r][2], cz = e[q][0] * e[r][1] - e[q][1] * e[r][0];
        if (sqrtf(cx * cx + cy * cy + cz * cz) / (len[q] * len[r]) < (float)o->triangle_crossnorm_min) return -1.0f;
    }
    float nx = e[0][1] * (-e[2][2]) - e[0][2] * (-e[2][1]), ny = e[0][2] * (-e[2][0]) - e[0][0] * (-e[2][2]), nz = e[0][0] * (-e[2][1]) - e[0][1] * (-e[2][0]);
    const float nn = sqrtf(nx * nx + ny * ny + nz * nz);
    nx /= nn; ny /= nn; nz /= nn;
    const float rx = (fu - (float)intr[1]) / (float)intr[0], ry = (fv - (float)intr[2]) / (float)intr[0], rz = 1.0f;
    const float rl = sqrtf(rx * rx + ry * ry + rz * rz);
    const float ndr = nx * rx + ny * ry + nz * rz;
    if (fabsf(ndr) / rl < (float)o->viewray_plane_min) return -1.0f;
    const float depth = (nx * A->x + ny * A->y + nz * A->z) / ndr; /* camera z of the ray / plane intersection */
    if (!(depth >= (float)o->depth_min) || !(depth <= (float)o->depth_max)) return -1.0f;
    if (o->local_rel_tolerance >= 0) {
        float smin = 0, smax = 0; int first = 1;
        for (int i = 0; i < m; ++i) {
            int b = (int)floorf((nb[i].z - zmin) / bw); if (b > NB - 1) b = NB - 1;
            if (b != sel) continue;
            if (first) { smin = smax = nb[i].z; first = 0; }
            if (nb[i].z < smin) smin = nb[i].z;
            if (nb[i].z > smax) smax = nb[i].z;
        }
        const float tol = (float)o->local_rel_tolerance;
        if (depth < smin * (1.0f - tol) || depth > smax * (1.0f + tol)) return -1.0f;
    }
    (void)zmax;
    return depth;
}

int kbo_lidar_depth(const float* cloud, int n_points, int stride, const double* T, const double* intr, const float* feats,
                    int n_feats, const kba_lidar_options* o, float* out) {
    if (!cloud || !T || !intr || !feats || !o || !out || stride < 3) return KBA_ERR_BAD_ARG;
    double Rd[9];
    quat_R(T, Rd);
    float R[9], t[3] = {(float)T[4], (float)T[5], (float)T[6]};
    for (int i = 0; i < 9; ++i) R[i] = (float)Rd[i];
    const float f = (float)intr[0], cx = (float)intr[1], cy = (float)intr[2];
    proj_pt* pts = (proj_pt*)malloc(sizeof(proj_pt) * (n_points > 0 ? n_points : 1));
    int np = 0;
    for (int i = 0; i < n_points; ++i) {
        const float* p = cloud + (size_t)i * stride;
        const float x = R[0] * p[0] + R[1] * p[1] + R[2] * p[2] + t[0];
        const float y = R[3] * p[0] + R[4] * p[1] + R[5] * p[2] + t[1];
        const float z = R[6] * p[0] + R[7] * p[1] + R[8] * p[2] + t[2];
        if (!(z > 0.0f)) continue; /* do_use_cut_behind_camera */
        const float u = f * x / z + cx, v = f * y / z + cy;
        if (!(u >= 0.0f && u < (float)o->image_width && v >= 0.0f && v < (float)o->image_height)) continue;
        pts[np].u = u; pts[np].v = v; pts[np].x = x; pts[np].y = y; pts[np].z = z; pts[np].idx = i;
        np++;
    }
    proj_pt* nb = (proj_pt*)malloc(sizeof(proj_pt) * (np > 0 ? np : 1));
    const float hw = 0.5f * (float)o->rect_width, hh = 0.5f * (float)o->rect_height;
    for (int k = 0; k < n_feats; ++k) {
        const float fu = feats[2 * k], fv = feats[2 * k + 1];
        const float cu = fu + (float)o->rect_offset_x, cv = fv + (float)o->rect_offset_y;
        int m = 0;
        for (int i = 0; i < np; ++i)
            if (fabsf(pts[i].u - cu) <= hw && fabsf(pts[i].v - cv) <= hh) nb[m++] = pts[i];
        out[k] = (m > 96) ? -1.0f : feature_depth(nb, m, fu, fv, intr, o); /* > 96 neighbours: rejected (kernel capacity) */
    }
    free(nb); free(pts);
    return KBA_OK;
}
