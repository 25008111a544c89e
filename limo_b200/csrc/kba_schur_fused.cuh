// kba_schur_fused.cuh -- Schur complement of a small window (<= 184 reduced rows) in ONE warp-specialised kernel:
//
//   producers (4 warps)  : per group of 8 landmarks, V_i = (J_p^T J_l) L^-T of every observation is formed straight into a
//                          shared-memory panel (24 columns x the group's reduced-system rows, column-major, zeros
//                          included), from the materialised J_p (cp.async-prefetched one group ahead), J_l = (translation
//                          columns of J_p) R(keyframe) and the landmark's L^-1; plus the right-hand-side row z_j and the
//                          ground-plane rows.  Nothing of V ever goes to global memory.
//   consumers (12 warps) : Sred += V V^T on the FP64 tensor cores (mma.sync m8n8k4).  The whole lower triangle lives in
//                          the consumers' registers as 16x16 blocks (2x2 tiles: one shared-memory load per DMMA); only
//                          the tile pairs inside the group's row range are multiplied.
//   ring                 : 4 panel stages with full / empty mbarriers, so warps drift up to three groups apart and the
//                          per-group imbalance of the static block -> warp map (scripts/syrk_map_search.py) averages out.
//
// Replaces k_obs_v + k_gp_panel + k_schur_syrk_tma of round 1 (2.2 GB of zero-padded panels written and re-read per pass of
// a 148-window batch).  Included by kba_kernels.cu after dmma(), the mbarrier helpers and gp_row().
#pragma once

namespace kba {

constexpr int kLG = 8;                           // landmarks per group
constexpr int kGC = 3 * kLG;                     // panel columns per group
constexpr int kFStages = 4;
constexpr int kFMaxRs = 196;                     // 184 rows -> row stride 196 (== 4 mod 16)
constexpr int kFStageDoubles = kGC * kFMaxRs;
constexpr int kFObs = 128;                       // observations staged per group = producer lanes
constexpr int kFConsumerWarps = 12;
constexpr int kFMaxKf = 32;

// 16x16 block (linear index bi (bi + 1) / 2 + bj of the 12-row block triangle) owned by each consumer warp: slot 0 is
// the warp's block of row 11 (only systems of more than 176 rows have one), slots 1..6 blocks of rows <= 10.
__constant__ signed char kSyrkMap12[12][7] = {
    {72, 7, 22, 33, 36, 48, 65}, {67, 10, 12, 19, 40, 53, -1}, {73, 1, 13, 17, 42, 50, -1}, {77, 8, 21, 26, 39, 54, -1},
    {75, 3, 18, 34, 49, 57, -1}, {69, 6, 23, 41, 58, 62, -1},  {66, 11, 27, 32, 38, 64, -1}, {74, 9, 20, 30, 37, 52, 55},
    {68, 0, 14, 16, 31, 44, 51}, {76, 2, 24, 29, 43, 45, 60},  {71, 5, 25, 28, 46, 56, 63},  {70, 4, 15, 35, 47, 59, 61}};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_elem(double* dst, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_elem(float* dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void producer_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

constexpr size_t schur_fused_smem() {
    return (size_t)kFStages * kFStageDoubles * sizeof(double) + (size_t)3 * 18 * kFObs * sizeof(double) +
           (size_t)kFMaxKf * kPoseStride * sizeof(double) + 16 * sizeof(uint64_t);
}

// V rows of one observation into the group's panel.  jp(q): entry q of the 3x6 pose block; li: L^-1 (i00; i10 i11; i20
// i21 i22); R: the keyframe's rotation; col0: panel entry (row of the pose block, first column of the landmark).
template <typename JpLoad>
__device__ __forceinline__ void fused_emit(JpLoad jp, const double* __restrict__ li, const double* __restrict__ R,
                                           double* col0, int rs, bool add) {
    double jl[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double m0 = jp(6 * i + 3), m1 = jp(6 * i + 4), m2 = jp(6 * i + 5);
#pragma unroll
        for (int c = 0; c < 3; ++c) jl[3 * i + c] = m0 * R[c] + m1 * R[3 + c] + m2 * R[6 + c];
    }
    const double i00 = li[0], i10 = li[1], i11 = li[2], i20 = li[3], i21 = li[4], i22 = li[5];
    const bool even = ((size_t)col0 & 15) == 0 && (rs & 1) == 0;
#pragma unroll
    for (int r = 0; r < 6; r += 2) {
        double v[2][3];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const double a = jp(r + h), b = jp(6 + r + h), c = jp(12 + r + h);
            const double e0 = a * jl[0] + b * jl[3] + c * jl[6];
            const double e1 = a * jl[1] + b * jl[4] + c * jl[7];
            const double e2 = a * jl[2] + b * jl[5] + c * jl[8];
            v[h][0] = e0 * i00; v[h][1] = e0 * i10 + e1 * i11; v[h][2] = e0 * i20 + e1 * i21 + e2 * i22;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double* q = col0 + (size_t)c * rs + r;
            if (even) {
                double2 t = make_double2(v[0][c], v[1][c]);
                if (add) { const double2 old = *reinterpret_cast<double2*>(q); t.x += old.x; t.y += old.y; }
                *reinterpret_cast<double2*>(q) = t;
            } else {
                q[0] = add ? q[0] + v[0][c] : v[0][c];
                q[1] = add ? q[1] + v[1][c] : v[1][c];
            }
        }
    }
}

// KBA_PROF build: cycles per role (lane 0 of every warp, summed over CTAs into BatchDev::prof) --
//   consumers: [0] waiting for a full panel, [1] multiplying;  producers: [4] waiting for an empty stage, [5] zero fill +
//   copy wait + barrier, [6] scatter + rows + hand-over, [7] (unused)
#ifdef KBA_PROF
#define KBA_PROF_DECL long long pt_ = clock64(), pacc_[3] = {0, 0, 0}
#define KBA_PROF_T0 pt_ = clock64()
#define KBA_PROF_ACC(i) do { const long long n_ = clock64(); pacc_[i] += n_ - pt_; pt_ = n_; } while (0)
#define KBA_PROF_FLUSH(base) do { if (lane == 0 && bd.prof) for (int i_ = 0; i_ < 3; ++i_) atomicAdd(bd.prof + (base) + i_, (unsigned long long)pacc_[i_]); } while (0)
#else
#define KBA_PROF_DECL
#define KBA_PROF_T0
#define KBA_PROF_ACC(i)
#define KBA_PROF_FLUSH(base)
#endif

template <int kSlots, typename TLin>
__global__ void __launch_bounds__(512, 1) k_schur_fused(BatchDev bd) {
    const int w = blockIdx.y;
    const WinState& st = bd.state[w];
    if (st.phase != PH_ITERATE || st.solve_failed) return;
    const WinDesc& wd = bd.desc[w];
    if (wd.landmarks_fixed) return;
    extern __shared__ __align__(128) unsigned char fsm[];
    double* stage = reinterpret_cast<double*>(fsm);
    TLin* jpbuf = reinterpret_cast<TLin*>(fsm + (size_t)kFStages * kFStageDoubles * sizeof(double));
    double* s_pose = reinterpret_cast<double*>(fsm + (size_t)kFStages * kFStageDoubles * sizeof(double) +
                                               (size_t)3 * 18 * kFObs * sizeof(double));
    uint64_t* full = reinterpret_cast<uint64_t*>(s_pose + kFMaxKf * kPoseStride);
    uint64_t* empty = full + kFStages;
    uint64_t* pose_bar = empty + kFStages;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_f = st.n_f, nt = (n_f + 8) >> 3, trhs = n_f >> 3;
    const int per = (wd.n_groups + (int)gridDim.x - 1) / (int)gridDim.x;
    const int g0 = blockIdx.x * per, g1 = min(wd.n_groups, g0 + per);
    const int* grs = bd.grp_rs + wd.grp_off;
    const int* gt0 = bd.grp_t0 + wd.grp_off;
    const int* gt1 = bd.grp_t1 + wd.grp_off;
    if (tid == 0) {
        for (int i = 0; i < kFStages; ++i) { mbar_init(&full[i], 128); mbar_init(&empty[i], kFConsumerWarps); }
        mbar_init(pose_bar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    auto next_group = [&](int g) { while (g < g1 && grs[g] == 0) ++g; return g; };

    if (warp >= kFConsumerWarps) {
        // =============================== producers ===============================
        if (kSlots == 7) reg_dec<56>(); else reg_dec<104>();
        const int ptid = tid - 32 * kFConsumerWarps;
        if (ptid == 0) {  // keyframe rotations (R | t as written next to the poses) with one bulk copy
            const uint32_t bytes = (uint32_t)(wd.n_kf * kPoseStride * sizeof(double));
            mbar_expect_tx(pose_bar, bytes);
            tma_load_1d(s_pose, bd.rt[st.cur] + (size_t)kPoseStride * wd.kf_off, bytes, pose_bar);
        }
        mbar_wait(pose_bar, 0);
        const int* lm_ptr = bd.lm_ptr + wd.lm_off + w;
        const size_t T = (size_t)bd.tot_obs, base = (size_t)wd.obs_off, TG = (size_t)bd.tot_gp;
        const TLin* jpg = reinterpret_cast<const TLin*>(bd.jp);
        // Software pipeline over the groups, three deep, so that no global-memory latency sits on the per-group critical
        // path (a group's panel has to be ready every ~2000 cycles):
        //   P3 (group g+3): observation range of the group                              -> a_*
        //   P2 (group g+2): this lane's observation: row / keyframe / landmark / rank   -> m2_*, J_p by cp.async -> jpbuf
        //   P1 (group g+1): L^-1 of the observation's landmark, the group's tile range, z row and validity -> *1
        //   P0 (group g)  : zero the panel stage, scatter V, right-hand-side and ground-plane rows, hand over
        int a_ob = 0, a_oe = 0, b_ob = 0, b_oe = 0, c_ob = 0, c_oe = 0, d_ob = 0, d_oe = 0;
        int m2_row = -1, m2_kf = 0, m2_lm = 0, m2_rank = 0, m1_row = -1, m1_kf = 0, m1_lm = 0, m1_rank = 0;
        int m0_row = -1, m0_kf = 0, m0_lm = 0, m0_rank = 0;
        int rs1 = 0, t01 = 0, t11 = 0, rs0 = 0, t00 = 0, t10 = 0;
        int zok1 = 0, zok0 = 0, gl1 = -1, gl0 = -1;
        double l1[6], l0[6], z1[3], z0[3];
#pragma unroll
        for (int q = 0; q < 6; ++q) l1[q] = l0[q] = 0.0;
#pragma unroll
        for (int q = 0; q < 3; ++q) z1[q] = z0[q] = 0.0;
        const int zj = ptid / 10, zr = ptid - 10 * zj;  // lanes 0..79: landmark zj of the group, ground-plane row zr (z row: zr == 0)
        int gi = 0;
        KBA_PROF_DECL;
        for (int gg = g0 - 3; gg < g1; ++gg) {
            {   // ---- P3
                const int g = gg + 3;
                if (g >= g0 && g < g1) {
                    a_ob = lm_ptr[g * kLG];
                    a_oe = lm_ptr[min(wd.n_lm, g * kLG + kLG)];
                }
            }
            {   // ---- P2 (b_* = range of group gg + 2)
                const int g = gg + 2;
                m2_row = -1;
                if (g >= g0 && g < g1) {
                    const int o = b_ob + ptid;
                    if (o < b_oe) {
                        const size_t oo = base + o;
                        m2_row = bd.obs_row[oo]; m2_kf = bd.obs_kf[oo]; m2_lm = bd.obs_lm[oo]; m2_rank = bd.obs_rank[oo];
                        TLin* dst = jpbuf + (size_t)(g % 3) * 18 * kFObs + ptid;
#pragma unroll
                        for (int q = 0; q < 18; ++q) cp_async_elem(dst + q * kFObs, jpg + q * T + oo);
                    }
                }
                cp_async_commit();
            }
            {   // ---- P1 (m1_* = this lane's observation of group gg + 1)
                const int g = gg + 1;
                zok1 = 0; gl1 = -1;
                if (g >= g0 && g < g1) {
                    rs1 = grs[g]; t01 = gt0[g]; t11 = gt1[g];
                    if (m1_row >= 0) {
                        const double* lp = bd.lm_linv + 6 * (size_t)(wd.lm_off + m1_lm);
#pragma unroll
                        for (int q = 0; q < 6; ++q) l1[q] = lp[q];
                    }
                    const int j = g * kLG + zj;
                    if (ptid < 80 && j < wd.n_lm) {
                        const int L = wd.lm_off + j;
                        zok1 = bd.lm_active[L] && lm_ptr[j + 1] > lm_ptr[j];
                        if (zr == 0) { const double* zz = bd.lm_z + 3 * (size_t)L; z1[0] = zz[0]; z1[1] = zz[1]; z1[2] = zz[2]; }
                        if (wd.n_gp > 0) gl1 = bd.gp_of_lm[L];
                    }
                }
            }
            if (gg >= g0 && rs0 != 0) {   // ---- P0
                const int g = gg, rs = rs0, t0 = t00, t1 = t10;
                const int slot = gi & (kFStages - 1);
                KBA_PROF_T0;
                if (gi >= kFStages) mbar_wait(&empty[slot], ((gi / kFStages) - 1) & 1);
                KBA_PROF_ACC(0);
                double* sb = stage + (size_t)slot * kFStageDoubles;
                {
                    double2* z = reinterpret_cast<double2*>(sb);
                    for (int i = ptid; i < (kGC / 2) * rs; i += 128) z[i] = make_double2(0.0, 0.0);
                }
                cp_async_wait<2>();  // all but the two newest commit groups (groups g+1, g+2) have landed
                producer_sync();
                KBA_PROF_ACC(1);
                const int j0 = g * kLG;
                const TLin* jb = jpbuf + (size_t)(g % 3) * 18 * kFObs + ptid;
                for (int round = 0; round <= wd.max_rank; ++round) {
                    if (round > 0) producer_sync();
                    if (m0_row >= 0 && m0_rank == round)
                        fused_emit([&](int q) { return (double)jb[q * kFObs]; }, l0, s_pose + kPoseStride * m0_kf,
                                   sb + (size_t)(3 * (m0_lm - j0)) * rs + (m0_row - 8 * t0), rs, round > 0);
                    for (int o = d_ob + kFObs + ptid; o < d_oe; o += kFObs) {  // groups of more than 128 observations (rare)
                        const size_t oo = base + o;
                        const int row = bd.obs_row[oo];
                        if (row < 0 || bd.obs_rank[oo] != round) continue;
                        const int jj = bd.obs_lm[oo] - j0;
                        fused_emit([&](int q) { return (double)jpg[q * T + oo]; }, bd.lm_linv + 6 * (size_t)(wd.lm_off + j0 + jj),
                                   s_pose + kPoseStride * bd.obs_kf[oo], sb + (size_t)(3 * jj) * rs + (row - 8 * t0), rs, round > 0);
                    }
                }
                producer_sync();  // every V row is in place: the ground-plane rows may add onto them
                if (zok0) {       // lanes 0..79 of valid landmarks: z row (zr == 0) and ground-plane row zr
                    double* col = sb + (size_t)(3 * zj) * rs;
                    if (zr == 0) {
                        const int rl = (trhs >= t0 && trhs < t1) ? n_f - 8 * t0 : 8 * (t1 - t0) + (n_f - 8 * trhs);
                        col[rl] = z0[0]; col[rs + rl] = z0[1]; col[2 * rs + rl] = z0[2];
                    }
                    if (gl0 >= 0) {
                        const size_t G = (size_t)wd.gp_off + gl0;
                        const int row = gp_row(bd, wd, bd.gp_kf[G], zr);
                        if (row >= 0) {
                            double* q = col + (row - 8 * t0);
                            q[0] += bd.vgp[(3 * zr + 0) * TG + G];
                            q[rs] += bd.vgp[(3 * zr + 1) * TG + G];
                            q[2 * rs] += bd.vgp[(3 * zr + 2) * TG + G];
                        }
                    }
                }
                mbar_arrive(&full[slot]);  // release: this thread's panel writes are visible to the consumers' acquire
                KBA_PROF_ACC(2);
                ++gi;
            }
            KBA_PROF_T0;
            // ---- shift the pipeline registers
            d_ob = c_ob; d_oe = c_oe; c_ob = b_ob; c_oe = b_oe; b_ob = a_ob; b_oe = a_oe;
            m0_row = m1_row; m0_kf = m1_kf; m0_lm = m1_lm; m0_rank = m1_rank;
            m1_row = m2_row; m1_kf = m2_kf; m1_lm = m2_lm; m1_rank = m2_rank;
            rs0 = rs1; t00 = t01; t10 = t11; zok0 = zok1; gl0 = gl1;
#pragma unroll
            for (int q = 0; q < 6; ++q) l0[q] = l1[q];
#pragma unroll
            for (int q = 0; q < 3; ++q) z0[q] = z1[q];
        }
        cp_async_wait<0>();
        KBA_PROF_FLUSH(4);
        return;
    }

    // =============================== consumers ===============================
    if (kSlots == 7) reg_inc<152>(); else reg_inc<136>();
    const int fr = lane >> 2, fc = lane & 3;
    const int nb2 = (nt + 1) >> 1;
    double acc[kSlots][4][2];
    int my_bi[kSlots], my_bj[kSlots];
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[s][q][0] = acc[s][q][1] = 0.0;
        const int t = kSyrkMap12[warp][s + (7 - kSlots)];
        int bi = 1 << 20, bj = 0;
        if (t >= 0) {
            bi = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
            while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
            while (bi * (bi + 1) / 2 > t) --bi;
            bj = t - bi * (bi + 1) / 2;
        }
        my_bi[s] = bi; my_bj[s] = bj;
    }
    {
        int gi = 0;
        int g = next_group(g0), rs = 0, t0 = 0, t1 = 0;
        if (g < g1) { rs = grs[g]; t0 = gt0[g]; t1 = gt1[g]; }
        KBA_PROF_DECL;
        for (; g < g1; ++gi) {
            const int gn = next_group(g + 1);  // the next group's tile range is requested before this group's panel is awaited
            int rsn = 0, t0n = 0, t1n = 0;
            if (gn < g1) { rsn = grs[gn]; t0n = gt0[gn]; t1n = gt1[gn]; }
            const int slot = gi & (kFStages - 1);
            KBA_PROF_T0;
            mbar_wait(&full[slot], (gi / kFStages) & 1);
            KBA_PROF_ACC(0);
            const double* sb = stage + (size_t)slot * kFStageDoubles + (size_t)fc * rs + fr;
            auto tile_row = [&](int i) -> int {  // panel row of tile i in this group, -1 if the group has no such rows
                if (i >= t0 && i < t1) return 8 * (i - t0);
                if (i == trhs) return 8 * (t1 - t0);
                return -1;
            };
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
                const int bi = my_bi[s], bj = my_bj[s];
                if (bi >= nb2) continue;
                const int ri0 = tile_row(2 * bi), ri1 = tile_row(2 * bi + 1), rj0 = tile_row(2 * bj), rj1 = tile_row(2 * bj + 1);
                if ((ri0 < 0 && ri1 < 0) || (rj0 < 0 && rj1 < 0)) continue;
                const bool diag = bi == bj;
                const double* pa0 = sb + max(ri0, 0);
                const double* pa1 = sb + max(ri1, 0);
                const double* pb0 = sb + max(rj0, 0);
                const double* pb1 = sb + max(rj1, 0);
                if (ri0 >= 0 && ri1 >= 0 && rj0 >= 0 && rj1 >= 0) {
#pragma unroll
                    for (int kk = 0; kk < kGC; kk += 4) {
                        const size_t o = (size_t)kk * rs;
                        const double a0 = pa0[o], a1 = pa1[o], b0 = pb0[o], b1 = pb1[o];
                        dmma(acc[s][0][0], acc[s][0][1], a0, b0);
                        if (!diag) dmma(acc[s][1][0], acc[s][1][1], a0, b1);
                        dmma(acc[s][2][0], acc[s][2][1], a1, b0);
                        dmma(acc[s][3][0], acc[s][3][1], a1, b1);
                    }
                } else {  // a block on the edge of the group's row range: only the tile pairs that exist
                    const bool p00 = ri0 >= 0 && rj0 >= 0, p01 = ri0 >= 0 && rj1 >= 0 && !diag, p10 = ri1 >= 0 && rj0 >= 0,
                               p11 = ri1 >= 0 && rj1 >= 0;
#pragma unroll 2
                    for (int kk = 0; kk < kGC; kk += 4) {
                        const size_t o = (size_t)kk * rs;
                        const double a0 = pa0[o], a1 = pa1[o], b0 = pb0[o], b1 = pb1[o];
                        if (p00) dmma(acc[s][0][0], acc[s][0][1], a0, b0);
                        if (p01) dmma(acc[s][1][0], acc[s][1][1], a0, b1);
                        if (p10) dmma(acc[s][2][0], acc[s][2][1], a1, b0);
                        if (p11) dmma(acc[s][3][0], acc[s][3][1], a1, b1);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
            KBA_PROF_ACC(1);
            g = gn; rs = rsn; t0 = t0n; t1 = t1n;
        }
        KBA_PROF_FLUSH(0);
    }
    double* out = bd.sred + wd.s_off * (size_t)bd.p_split + (size_t)blockIdx.x * wd.nr_cap * wd.nr_cap;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
        const int bi = my_bi[s], bj = my_bj[s];
        if (bi >= nb2) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = 2 * bi + (q >> 1), j = 2 * bj + (q & 1);
            if (i >= nt || j > i) continue;
            double* o = out + (size_t)(8 * i + fr) * wd.nr_cap + 8 * j + 2 * fc;
            o[0] = acc[s][q][0];
            o[1] = acc[s][q][1];
        }
    }
}

}  // namespace kba
