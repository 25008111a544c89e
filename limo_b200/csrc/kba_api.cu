// kba_api.cu -- host side of the C ABI declared in include/kba_b200.h: handle / batch lifetime, packing of caller
// windows into the batch-flat device layout, the pass loop, result download.  No numerical work happens on the host.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "kba_b200.h"
#include "kba_kernels.h"

using namespace kba;

static thread_local std::string g_last_error;
static thread_local bool g_force_host_pack = false;  // kba_eval: needs the host-side observation permutation
static int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
#define CU(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return fail(KBA_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                    \
    } while (0)

struct kba_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    Counters counters;
    bool kernel_timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> ev_pool;  // event pairs around every residual/Jacobian launch (kernel timing)
    int ev_used = 0;
    int sm_count = 148;
    cudaEvent_t ev_block = nullptr;  // blocking-sync event: waiting host threads sleep instead of spinning on a core
    bool blocking_sync = false;      // KBA_BLOCKING_SYNC=1 at kba_create
    // grow-only device workspace of the single-shot entry points (kba_lidar_depth): no cudaMalloc / cudaFree per call
    void* ws = nullptr;
    size_t ws_cap = 0;
};

// Wait for the handle's stream.  With KBA_BLOCKING_SYNC=1 (read at kba_create) the host thread sleeps on a blocking-sync
// event instead of spinning on a core (cudaStreamSynchronize spins under the default scheduling policy): for several
// handles per process / one process per GPU that share the box's cores with the packing threads.
static cudaError_t wait_event(kba_handle*, cudaEvent_t ev) { return cudaEventSynchronize(ev); }
static cudaError_t wait_stream(kba_handle* h) {
    if (!h->blocking_sync) return cudaStreamSynchronize(h->stream);  // lowest latency: the default for a lone handle
    if (!h->ev_block) {
        const cudaError_t e = cudaEventCreateWithFlags(&h->ev_block, cudaEventBlockingSync | cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    const cudaError_t e = cudaEventRecord(h->ev_block, h->stream);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(h->ev_block);
}

// ---- a device + pinned-host buffer pair, filled on the host and uploaded with one async copy ----------------------------
template <typename T>
struct Staged {
    T* h = nullptr;
    T* d = nullptr;
    size_t n = 0;
    int alloc(size_t count, bool host_copy) {
        n = count;
        const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
        if (cudaMalloc(&d, bytes) != cudaSuccess) return 1;
        if (host_copy && cudaMallocHost(&h, bytes) != cudaSuccess) return 1;
        return 0;
    }
    void release() {
        if (d) cudaFree(d);
        if (h) cudaFreeHost(h);
        d = nullptr; h = nullptr;
    }
    cudaError_t upload(cudaStream_t s) { return cudaMemcpyAsync(d, h, std::max<size_t>(n, 1) * sizeof(T), cudaMemcpyHostToDevice, s); }
    cudaError_t download(cudaStream_t s) { return cudaMemcpyAsync(h, d, std::max<size_t>(n, 1) * sizeof(T), cudaMemcpyDeviceToHost, s); }
};

struct kba_batch {
    kba_handle* h = nullptr;
    BatchDev bd{};
    std::vector<WinDesc> desc_h;
    // staged inputs
    Staged<WinDesc> desc;
    Staged<double> pose0, plane0, cam, lm0, lm_weight;
    Staged<uint8_t> kf_fixed;
    Staged<int> lm_ptr, obs_kf, obs_cam, obs_lm, kf_ptr, pm_lm, pm_cam, chunk_lm0, chunk_lm1, chunk_k0, chunk_k1, lm_orig, obs_orig;
    Staged<int> grp_k0, grp_k1;
    // device-side packing (kba_pack.cu): the caller's arrays are uploaded as they are, the sorted layout is built by kernels
    bool device_pack = false;
    Staged<int> r_lm_ptr, r_obs_kf, r_obs_cam, r_gp_lm;
    Staged<float> r_obs_u, r_obs_v, r_obs_d;
    Staged<double> r_lm_pos, r_lm_weight;
    Staged<double> lm_user;            // landmark results in the caller's order (device + pinned)
    Staged<uint8_t> rej_user;
    PackRaw raw;
    Staged<int> obs_rank;
    Staged<float> obs_u, obs_v, obs_d, pm_u, pm_v, pm_d;
    // outputs
    Staged<WinState> state;
    Staged<IterRecord> log;
    Staged<double> pose_out[2], lm_out[2], plane_out[2];
    Staged<int> gp_lm, gp_kf, gp_of_lm, gp_shared;
    Staged<double> gp_weight;
    Staged<uint8_t> lm_active;
    Staged<int> n_active;
    Staged<unsigned long long> jac_obs;
    std::vector<void*> scratch;  // device-only allocations
    LaunchCfg lc;
    size_t h2d_bytes = 0, d2h_bytes = 0;
    float last_solve_ms = 0.f;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_poll = nullptr, ev_poll2 = nullptr;
    // The pass sequence as a CUDA graph (kba_batch_solve): mode 2 = ONE launch per solve, the passes are the body of a conditional
    // WHILE node whose condition the device sets (k_loop_cond); mode 1 = a graph of `check_every` passes launched until the
    // downloaded active-window count is zero.  Rebuilt when anything a kernel receives by value changes (`key`).
    struct SolveGraph {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        std::vector<unsigned char> key;
        Counters per_pass;
        int mode = 0, passes_per_launch = 0;
        bool unusable = false;  // capture / instantiation failed once on this batch: the stream path is used from then on
        void destroy() {
            if (exec) cudaGraphExecDestroy(exec);
            if (graph) cudaGraphDestroy(graph);
            exec = nullptr; graph = nullptr; key.clear();
        }
    } sg;
    Staged<int> loop_pass;  // passes the WHILE node has run (device counter + pinned copy)
    long long solves_done = 0;

    template <typename T>
    int dev_alloc(T** p, size_t count) {
        void* q = nullptr;
        if (cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)) != cudaSuccess) return 1;
        scratch.push_back(q);
        *p = (T*)q;
        return 0;
    }
    void dev_free(void* q) {
        scratch.erase(std::remove(scratch.begin(), scratch.end(), q), scratch.end());
        cudaFree(q);
    }
    void release() {
        desc.release(); pose0.release(); plane0.release(); cam.release(); lm0.release(); lm_weight.release();
        kf_fixed.release(); lm_ptr.release(); obs_kf.release(); obs_cam.release(); obs_lm.release(); kf_ptr.release();
        pm_lm.release(); pm_cam.release(); chunk_lm0.release(); chunk_lm1.release(); chunk_k0.release(); chunk_k1.release();
        lm_orig.release(); obs_orig.release(); obs_rank.release(); obs_u.release(); obs_v.release();
        grp_k0.release(); grp_k1.release();
        r_lm_ptr.release(); r_obs_kf.release(); r_obs_cam.release(); r_gp_lm.release(); r_obs_u.release(); r_obs_v.release();
        r_obs_d.release(); r_lm_pos.release(); r_lm_weight.release(); lm_user.release(); rej_user.release();
        obs_d.release(); pm_u.release(); pm_v.release(); pm_d.release(); state.release(); log.release();
        pose_out[0].release(); pose_out[1].release(); lm_out[0].release(); lm_out[1].release(); lm_active.release();
        n_active.release(); jac_obs.release(); plane_out[0].release(); plane_out[1].release();
        gp_lm.release(); gp_kf.release(); gp_of_lm.release(); gp_weight.release(); gp_shared.release();
        for (void* p : scratch) cudaFree(p);
        scratch.clear();
        if (ev_a) cudaEventDestroy(ev_a);
        if (ev_b) cudaEventDestroy(ev_b);
        if (ev_poll) cudaEventDestroy(ev_poll);
        if (ev_poll2) cudaEventDestroy(ev_poll2);
        sg.destroy();
        loop_pass.release();
    }
};

// persistent, device-resident window (kba_track_*, at the end of this file)
struct kba_track {
    kba_handle* h = nullptr;
    kba_track_caps caps{};
    int n_cam = 0;
    kba_batch* batch = nullptr;            // one window of capacity shape; its raw arrays are filled by the gather kernels
    TrackDev td{};
    int* arena_i[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};      // [buffer][lm, cam]
    float* arena_f[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};  // [buffer][u, v, d]
    int arena_cur = 0, arena_used = 0;
    std::vector<int> m_off, m_cnt;         // host mirror of the arena layout
    std::vector<char> kf_live;
    std::vector<void*> dev;
    Staged<int> p_lm, p_cam, sel_kf, sel_lm, lay;  // pinned staging: one push / one selection / arena layout
    Staged<float> p_u, p_v, p_d;
    Staged<uint8_t> sel_fixed;
    Staged<double> p_dbl;                  // poses / landmark values on their way to the store
    Staged<int> p_slot;                    // ... and the slots they go to
    int push_cap = 0, set_cap = 0;
    int64_t h2d_solve = 0, d2h_solve = 0, h2d_push = 0;
    template <typename T> int alloc(T** p, size_t n) {
        void* q = nullptr;
        if (cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) return 1;
        dev.push_back(q); *p = (T*)q; return 0;
    }
    void point_arena() {
        td.m_lm = arena_i[arena_cur][0]; td.m_cam = arena_i[arena_cur][1];
        td.m_u = arena_f[arena_cur][0]; td.m_v = arena_f[arena_cur][1]; td.m_d = arena_f[arena_cur][2];
    }
};

static int validate_window(const kba_window* w, std::string& why) {
    if (!w) { why = "null window"; return KBA_ERR_BAD_ARG; }
    if (w->n_kf < 0 || w->n_lm < 0 || w->n_obs < 0 || w->n_gp < 0 || w->n_cam < 1) { why = "negative size"; return KBA_ERR_BAD_ARG; }
    if (!w->kf_pose || !w->kf_fixed || !w->cam_intr || !w->cam_pose) { why = "null keyframe/camera array"; return KBA_ERR_BAD_ARG; }
    if (w->n_lm > 0 && (!w->lm_pos || !w->lm_weight || !w->lm_obs_ptr)) { why = "null landmark array"; return KBA_ERR_BAD_ARG; }
    if (w->n_obs > 0 && (!w->obs_kf || !w->obs_u || !w->obs_v || !w->obs_d)) { why = "null observation array"; return KBA_ERR_BAD_ARG; }
    if (w->n_kf > kMaxKf) { why = "more than 128 keyframes per window"; return KBA_ERR_CAPACITY; }
    if (w->n_cam > kMaxCam) { why = "more than 8 cameras per window"; return KBA_ERR_CAPACITY; }
    if ((w->n_gp > 0 || w->plane_reg_weight > 0) && !w->kf_plane) { why = "ground-plane residuals need kf_plane"; return KBA_ERR_BAD_ARG; }
    if (w->n_gp > 0 && (!w->gp_lm || !w->gp_kf || !w->gp_weight)) { why = "null ground-plane array"; return KBA_ERR_BAD_ARG; }
    for (int g = 0; g < w->n_gp; ++g)
        if (w->gp_lm[g] < 0 || w->gp_lm[g] >= w->n_lm || w->gp_kf[g] < 0 || w->gp_kf[g] >= w->n_kf) { why = "ground-plane index out of range"; return KBA_ERR_BAD_ARG; }
    if (w->speed_weight > 0 && (w->speed_kf < 0 || w->speed_kf >= w->n_kf || !(w->speed_dt > 0))) { why = "speed prior: keyframe out of range or dt <= 0"; return KBA_ERR_BAD_ARG; }
    if (w->n_lm > 0) {
        if (w->lm_obs_ptr[0] != 0) { why = "lm_obs_ptr[0] != 0"; return KBA_ERR_BAD_ARG; }
        for (int j = 0; j < w->n_lm; ++j)
            if (w->lm_obs_ptr[j + 1] < w->lm_obs_ptr[j]) { why = "lm_obs_ptr is not non-decreasing"; return KBA_ERR_BAD_ARG; }
        if (w->lm_obs_ptr[w->n_lm] != w->n_obs) { why = "lm_obs_ptr[n_lm] != n_obs"; return KBA_ERR_BAD_ARG; }
    } else if (w->n_obs != 0) { why = "observations without landmarks"; return KBA_ERR_BAD_ARG; }
    if (w->n_gp > 0) {  // at most one ground-plane residual per landmark (one Landmark::is_ground_plane flag, cpp:519-560)
        std::vector<unsigned char> seen((size_t)w->n_lm, 0);
        for (int g = 0; g < w->n_gp; ++g) {
            if (seen[w->gp_lm[g]]) { why = "two ground-plane residuals on one landmark"; return KBA_ERR_BAD_ARG; }
            seen[w->gp_lm[g]] = 1;
        }
    }
    for (int o = 0; o < w->n_obs; ++o) {
        if (w->obs_kf[o] < 0 || w->obs_kf[o] >= w->n_kf) { why = "obs_kf out of range"; return KBA_ERR_BAD_ARG; }
        if (w->obs_cam && (w->obs_cam[o] < 0 || w->obs_cam[o] >= w->n_cam)) { why = "obs_cam out of range"; return KBA_ERR_BAD_ARG; }
    }
    if (w->scale_weight > 0 && (w->scale_kf0 < 0 || w->scale_kf0 >= w->n_kf || w->scale_kf1 < 0 || w->scale_kf1 >= w->n_kf)) {
        why = "scale regulariser keyframe out of range"; return KBA_ERR_BAD_ARG;
    }
    return KBA_OK;
}

static void fill_window(kba_batch* b, int wi, const kba_window* w) {
    const WinDesc& d = b->desc_h[wi];
    memcpy(b->pose0.h + 7 * (size_t)d.kf_off, w->kf_pose, sizeof(double) * 7 * w->n_kf);
    memcpy(b->kf_fixed.h + d.kf_off, w->kf_fixed, w->n_kf);
    for (int k = 0; k < w->n_kf; ++k) {
        double* pl = b->plane0.h + 4 * (size_t)(d.kf_off + k);
        if (w->kf_plane) memcpy(pl, w->kf_plane + 4 * k, 4 * sizeof(double));
        else { pl[0] = 0; pl[1] = 0; pl[2] = 1; pl[3] = 0; }
    }
    for (int c = 0; c < w->n_cam; ++c) {
        double* o = b->cam.h + kCamStride * (size_t)(d.cam_off + c);
        quat_to_rot<double>(w->cam_pose + 7 * c, o);
        o[9] = w->cam_pose[7 * c + 4]; o[10] = w->cam_pose[7 * c + 5]; o[11] = w->cam_pose[7 * c + 6];
        o[12] = w->cam_intr[3 * c]; o[13] = w->cam_intr[3 * c + 1]; o[14] = w->cam_intr[3 * c + 2]; o[15] = 0;
    }
    // Landmarks are stored sorted by (first keyframe, last keyframe): consecutive landmarks then touch the same rows of
    // the reduced system, which is what lets the Schur kernel skip most tiles.  lm_orig maps back to the caller's order.
    const int nl = w->n_lm;
    int* orig = b->lm_orig.h + d.lm_off;
    {
        std::vector<long long> key(nl);
        for (int j = 0; j < nl; ++j) {
            const int o0 = w->lm_obs_ptr[j], o1 = w->lm_obs_ptr[j + 1];
            const int k0 = o1 > o0 ? w->obs_kf[o0] : w->n_kf, k1 = o1 > o0 ? w->obs_kf[o1 - 1] : w->n_kf;
            key[j] = ((long long)k0 << 40) | ((long long)k1 << 20) | 0;
            orig[j] = j;
        }
        std::stable_sort(orig, orig + nl, [&](int a, int c) { return key[a] < key[c]; });
    }
    int* lp = b->lm_ptr.h + d.lm_off + wi;
    int* kp = b->kf_ptr.h + d.kf_off + wi;
    std::fill(kp, kp + w->n_kf + 1, 0);
    int pos = 0, max_rank = 0;
    lp[0] = 0;
    for (int jn = 0; jn < nl; ++jn) {
        const int jo = orig[jn];
        memcpy(b->lm0.h + 3 * (size_t)(d.lm_off + jn), w->lm_pos + 3 * (size_t)jo, 3 * sizeof(double));
        b->lm_weight.h[d.lm_off + jn] = w->lm_weight[jo];
        for (int o = w->lm_obs_ptr[jo]; o < w->lm_obs_ptr[jo + 1]; ++o, ++pos) {
            const size_t e = (size_t)d.obs_off + pos;
            b->obs_kf.h[e] = w->obs_kf[o];
            b->obs_cam.h[e] = w->obs_cam ? w->obs_cam[o] : 0;
            b->obs_lm.h[e] = jn;
            b->obs_u.h[e] = w->obs_u[o]; b->obs_v.h[e] = w->obs_v[o]; b->obs_d.h[e] = w->obs_d[o];
            b->obs_orig.h[e] = o;
            // rank among the observations of this landmark in the same keyframe (multi-camera rigs)
            const int rank = (o > w->lm_obs_ptr[jo] && w->obs_kf[o - 1] == w->obs_kf[o]) ? b->obs_rank.h[e - 1] + 1 : 0;
            b->obs_rank.h[e] = rank;
            max_rank = std::max(max_rank, rank);
            kp[w->obs_kf[o] + 1]++;
        }
        lp[jn + 1] = pos;
    }
    b->desc_h[wi].max_rank = max_rank;
    b->desc.h[wi].max_rank = max_rank;
    // keyframe-major copy (counting sort, stable -> deterministic reduction order)
    for (int k = 0; k < w->n_kf; ++k) kp[k + 1] += kp[k];
    std::vector<int> cur(kp, kp + w->n_kf);
    for (int jn = 0; jn < nl; ++jn)
        for (int o = lp[jn]; o < lp[jn + 1]; ++o) {
            const size_t src = (size_t)d.obs_off + o;
            const size_t e = (size_t)d.obs_off + cur[b->obs_kf.h[src]]++;
            b->pm_lm.h[e] = jn;
            b->pm_cam.h[e] = b->obs_cam.h[src];
            b->pm_u.h[e] = b->obs_u.h[src]; b->pm_v.h[e] = b->obs_v.h[src]; b->pm_d.h[e] = b->obs_d.h[src];
        }
    // ground-plane residuals follow their landmark into the sorted order
    std::vector<int> gp_kf_of_lm(nl, -1);
    {
        std::vector<int> inv(nl);
        for (int jn = 0; jn < nl; ++jn) { inv[orig[jn]] = jn; b->gp_of_lm.h[d.lm_off + jn] = -1; }
        for (int g = 0; g < w->n_gp; ++g) {
            const int jn = inv[w->gp_lm[g]];
            b->gp_lm.h[d.gp_off + g] = jn;
            b->gp_kf.h[d.gp_off + g] = w->gp_kf[g];
            b->gp_weight.h[d.gp_off + g] = w->gp_weight[g];
            b->gp_of_lm.h[d.lm_off + jn] = g;
            gp_kf_of_lm[jn] = w->gp_kf[g];
            int shared = 0;
            for (int o = lp[jn]; o < lp[jn + 1]; ++o) shared |= (b->obs_kf.h[(size_t)d.obs_off + o] == w->gp_kf[g]);
            b->gp_shared.h[d.gp_off + g] = shared;
        }
    }
    for (int c = 0; c < d.n_chunks; ++c) {
        const int j0 = c * 32, j1 = std::min(nl, (c + 1) * 32);
        b->chunk_lm0.h[d.chunk_off + c] = j0;
        b->chunk_lm1.h[d.chunk_off + c] = j1;
        int k0 = w->n_kf, k1 = -1;
        for (int o = lp[j0]; o < lp[j1]; ++o) {
            const int k = b->obs_kf.h[(size_t)d.obs_off + o];
            k0 = std::min(k0, k); k1 = std::max(k1, k);
        }
        for (int j = j0; j < j1; ++j)
            if (gp_kf_of_lm[j] >= 0) { k0 = std::min(k0, gp_kf_of_lm[j]); k1 = std::max(k1, gp_kf_of_lm[j]); }
        b->chunk_k0.h[d.chunk_off + c] = k0;
        b->chunk_k1.h[d.chunk_off + c] = k1;
    }
    for (int c = 0; c < d.n_groups; ++c) {  // 8-landmark groups of the fused Schur kernel
        const int j0 = c * 8, j1 = std::min(nl, (c + 1) * 8);
        int k0 = w->n_kf, k1 = -1;
        for (int o = lp[j0]; o < lp[j1]; ++o) {
            const int k = b->obs_kf.h[(size_t)d.obs_off + o];
            k0 = std::min(k0, k); k1 = std::max(k1, k);
        }
        for (int j = j0; j < j1; ++j)
            if (gp_kf_of_lm[j] >= 0) { k0 = std::min(k0, gp_kf_of_lm[j]); k1 = std::max(k1, gp_kf_of_lm[j]); }
        b->grp_k0.h[d.grp_off + c] = k0;
        b->grp_k1.h[d.grp_off + c] = k1;
    }
}

// device-pack mode: the caller's arrays are copied as they are into the pinned staging buffers (one memcpy per array)
static void fill_window_raw(kba_batch* b, int wi, const kba_window* w) {
    const WinDesc& d = b->desc_h[wi];
    memcpy(b->pose0.h + 7 * (size_t)d.kf_off, w->kf_pose, sizeof(double) * 7 * w->n_kf);
    memcpy(b->kf_fixed.h + d.kf_off, w->kf_fixed, w->n_kf);
    for (int k = 0; k < w->n_kf; ++k) {
        double* pl = b->plane0.h + 4 * (size_t)(d.kf_off + k);
        if (w->kf_plane) memcpy(pl, w->kf_plane + 4 * k, 4 * sizeof(double));
        else { pl[0] = 0; pl[1] = 0; pl[2] = 1; pl[3] = 0; }
    }
    for (int c = 0; c < w->n_cam; ++c) {
        double* o = b->cam.h + kCamStride * (size_t)(d.cam_off + c);
        quat_to_rot<double>(w->cam_pose + 7 * c, o);
        o[9] = w->cam_pose[7 * c + 4]; o[10] = w->cam_pose[7 * c + 5]; o[11] = w->cam_pose[7 * c + 6];
        o[12] = w->cam_intr[3 * c]; o[13] = w->cam_intr[3 * c + 1]; o[14] = w->cam_intr[3 * c + 2]; o[15] = 0;
    }
    const size_t nl = (size_t)w->n_lm, no = (size_t)w->n_obs;
    memcpy(b->r_lm_pos.h + 3 * (size_t)d.lm_off, w->lm_pos, 3 * nl * sizeof(double));
    memcpy(b->r_lm_weight.h + d.lm_off, w->lm_weight, nl * sizeof(double));
    int* lp = b->r_lm_ptr.h + d.lm_off + wi;
    if (nl) memcpy(lp, w->lm_obs_ptr, (nl + 1) * sizeof(int)); else lp[0] = 0;
    memcpy(b->r_obs_kf.h + d.obs_off, w->obs_kf, no * sizeof(int));
    if (w->obs_cam) memcpy(b->r_obs_cam.h + d.obs_off, w->obs_cam, no * sizeof(int));
    else memset(b->r_obs_cam.h + d.obs_off, 0, no * sizeof(int));
    memcpy(b->r_obs_u.h + d.obs_off, w->obs_u, no * sizeof(float));
    memcpy(b->r_obs_v.h + d.obs_off, w->obs_v, no * sizeof(float));
    memcpy(b->r_obs_d.h + d.obs_off, w->obs_d, no * sizeof(float));
    if (w->n_gp) {
        memcpy(b->r_gp_lm.h + d.gp_off, w->gp_lm, w->n_gp * sizeof(int));
        memcpy(b->gp_kf.h + d.gp_off, w->gp_kf, w->n_gp * sizeof(int));
        memcpy(b->gp_weight.h + d.gp_off, w->gp_weight, w->n_gp * sizeof(double));
    }
    int max_rank = 0;  // several cameras of a rig seeing the landmark in one keyframe (observations are sorted by keyframe)
    for (int j = 0; j < w->n_lm; ++j) {
        int rank = 0;
        for (int o = w->lm_obs_ptr[j] + 1; o < w->lm_obs_ptr[j + 1]; ++o) {
            rank = (w->obs_kf[o] == w->obs_kf[o - 1]) ? rank + 1 : 0;
            max_rank = std::max(max_rank, rank);
        }
    }
    b->desc_h[wi].max_rank = max_rank;
    b->desc.h[wi].max_rank = max_rank;
}

// reduced-system rows a window can have given its constant keyframes (k_solve_begin may leave out more)
static int window_rows(const kba_window& w) {
    const bool planes = w.n_gp > 0 || w.plane_reg_weight > 0;
    int n_free = 0;
    for (int k = 0; k < w.n_kf; ++k) n_free += w.kf_fixed[k] ? 0 : 1;
    return (planes ? 10 : 6) * n_free + 1;
}

struct kba_shard_comm;
kba::Exchange kba_shard_exchange(kba_shard_comm* c);  // kba_shard.cu

// How the pass sequence of a solve is issued (KBA_GRAPH, read once):
//   2 (default)  one CUDA graph launch per solve: the passes are the body of a conditional WHILE node, k_loop_cond sets the
//                condition on the device -- no host polling, no pass enqueued after the last window finished, no launch gaps;
//   1            a graph of four passes + the active-window count, launched until the count read back is zero;
//   0            kernel by kernel on the stream (always used with kernel timing on -- event pairs around the linearisation
//                launches --, with KBA_LAUNCH_CHECK, or on the legacy default stream, which cannot be captured).
// A sharded solve (NCCL all-reduces between the kernels) takes mode 1 from the second solve of its batch on, see kba_batch_solve.
static int solve_graph_mode() {
    static const int m = [] {
        const char* e = getenv("KBA_GRAPH");
        const int v = e ? atoi(e) : 2;
        return (v < 0 || v > 2) ? 2 : v;
    }();
    return m;
}

// KBA_SHARD_GRAPH=0: sharded solves stay on the stream path
static bool shard_graph_enabled() {
    static const bool on = [] { const char* e = getenv("KBA_SHARD_GRAPH"); return e ? atoi(e) != 0 : true; }();
    return on;
}

template <typename T>
static void key_append(std::vector<unsigned char>& k, const T& v) {
    const unsigned char* p = reinterpret_cast<const unsigned char*>(&v);
    k.insert(k.end(), p, p + sizeof(T));
}

// (re)builds b->sg for the given kernel arguments; false = not possible here (b->sg.unusable is set, the caller takes the stream path)
static bool build_solve_graph(kba_batch* b, const SolveParams& sp, const LaunchCfg& lc, int mode, int max_passes, int check_every,
                              const std::vector<unsigned char>& key) {
    kba_batch::SolveGraph& g = b->sg;
    g.destroy();
    cudaStream_t s = b->h->stream;
    Counters per{};
    cudaError_t e = cudaSuccess;
    bool capturing = false;
    int rc_pass = 0;
    if (mode == 2) {
        cudaGraphConditionalHandle handle = 0;
        cudaGraphNode_t node = nullptr;
        cudaGraphNodeParams np = {};
        e = cudaGraphCreate(&g.graph, 0);
        if (e == cudaSuccess) e = cudaGraphConditionalHandleCreate(&handle, g.graph, 1, cudaGraphCondAssignDefault);
        if (e == cudaSuccess) {
            np.type = cudaGraphNodeTypeConditional;
            np.conditional.handle = handle;
            np.conditional.type = cudaGraphCondTypeWhile;
            np.conditional.size = 1;
            e = cudaGraphAddNode(&node, g.graph, nullptr, 0, &np);
        }
        if (e == cudaSuccess) e = cudaStreamBeginCaptureToGraph(s, np.conditional.phGraph_out[0], nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
            capturing = true;
            rc_pass = launch_pass(b->bd, sp, lc, &per, s);
            launch_loop_cond(b->bd, (unsigned long long)handle, b->loop_pass.d, max_passes, s);
            cudaGraph_t body = nullptr;
            e = cudaStreamEndCapture(s, &body);
            capturing = false;
        }
        g.passes_per_launch = 0;  // counted on the device
    } else {
        e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
            capturing = true;
            for (int i = 0; i < check_every && !rc_pass; ++i) rc_pass = launch_pass(b->bd, sp, lc, i == 0 ? &per : nullptr, s);
            launch_count_active(b->bd, s);
            b->n_active.download(s);
            e = cudaStreamEndCapture(s, &g.graph);
            capturing = false;
        }
        g.passes_per_launch = check_every;
    }
    if (capturing) { cudaGraph_t junk = nullptr; cudaStreamEndCapture(s, &junk); }
    if (e == cudaSuccess && rc_pass) e = cudaErrorUnknown;
    if (e == cudaSuccess) e = cudaGraphInstantiate(&g.exec, g.graph, 0);
    if (e != cudaSuccess) {
        if (getenv("KBA_GRAPH_VERBOSE")) fprintf(stderr, "kba: solve graph (mode %d) not available: %s -- stream launches\n", mode, cudaGetErrorString(e));
        g.destroy();
        g.unusable = true;
        cudaGetLastError();  // the stream path starts with a clean error state
        return false;
    }
    if (getenv("KBA_GRAPH_VERBOSE")) fprintf(stderr, "kba: solve graph built (mode %d, %lld launches per pass)\n", mode, per.launches_total);
    g.key = key;
    g.per_pass = per;
    g.mode = mode;
    return true;
}

static void add_pass_counters(Counters& c, const Counters& per, long long passes) {
    c.launches_total += per.launches_total * passes; c.launches_jacobian += per.launches_jacobian * passes;
    c.launches_prep += per.launches_prep * passes; c.launches_schur += per.launches_schur * passes;
    c.launches_solve += per.launches_solve * passes; c.launches_backsub += per.launches_backsub * passes;
    c.launches_cost += per.launches_cost * passes; c.launches_update += per.launches_update * passes;
    c.launches_trim += per.launches_trim * passes;
}

extern "C" {

int kba_version(void) { return KBA_VERSION_MAJOR * 100 + KBA_VERSION_MINOR; }
// helpers for the other translation units of the library (not part of the public header)
int kba_internal_stream(kba_handle* h, cudaStream_t* s, int* device) {
    if (!h) return fail(KBA_ERR_BAD_ARG, "null handle");
    *s = h->stream; *device = h->device;
    return KBA_OK;
}
int kba_internal_fail(int code, const char* msg) { return fail(code, msg ? msg : ""); }
// at least `bytes` of device memory owned by the handle, 256-byte aligned, valid until the next call that asks for more
int kba_internal_workspace(kba_handle* h, size_t bytes, void** out) {
    if (!h || !out) return fail(KBA_ERR_BAD_ARG, "null handle");
    if (bytes > h->ws_cap) {
        CU(cudaSetDevice(h->device));
        CU(cudaStreamSynchronize(h->stream));
        if (h->ws) cudaFree(h->ws);
        h->ws = nullptr; h->ws_cap = 0;
        const size_t cap = bytes + bytes / 4 + 4096;
        CU(cudaMalloc(&h->ws, cap));
        h->ws_cap = cap;
    }
    *out = h->ws;
    return KBA_OK;
}
const char* kba_last_error(void) { return g_last_error.c_str(); }

void kba_default_options(kba_options* o) {
    if (!o) return;
    memset(o, 0, sizeof *o);
    o->depth_thres = 0.16; o->reprojection_thres = 1.6;
    o->depth_quantile = 0.95; o->reprojection_quantile = 0.95; o->gp_quantile = 1.0; o->gp_huber = 0.1;
    o->num_trim_rounds = -1; o->trim_solver_iterations = 2; o->final_solver_iterations = 100;
    o->min_landmarks_for_trimming = 100; o->min_residual_groups = 30; o->num_rounds_option = 1;
    o->solver_time_sec = 20.0;
    o->function_tolerance = 1e-6; o->gradient_tolerance = 1e-10; o->parameter_tolerance = 1e-8;
    o->initial_trust_region_radius = 1e4; o->max_trust_region_radius = 1e16; o->min_trust_region_radius = 1e-32;
    o->min_relative_decrease = 1e-3; o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32;
    o->max_consecutive_invalid_steps = 5; o->precision = 0;
}

int kba_create(kba_handle** out, int device) {
    if (!out) return fail(KBA_ERR_BAD_ARG, "null out pointer");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
        return fail(KBA_ERR_CUDA, "no CUDA device available: the kba_b200 library has no CPU fallback");
    if (device < 0 || device >= n) return fail(KBA_ERR_BAD_ARG, "device index out of range");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(KBA_ERR_CUDA, "kba_b200 kernels are built for sm_100a only");
    kba_handle* h = new kba_handle();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    { const char* e = std::getenv("KBA_BLOCKING_SYNC"); h->blocking_sync = e && std::atoi(e) != 0; }
    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
    CU(cudaEventCreate(&h->ev0));
    CU(cudaEventCreate(&h->ev1));
    *out = h;
    return KBA_OK;
}

void kba_destroy(kba_handle* h) {
    if (!h) return;
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_block) cudaEventDestroy(h->ev_block);
    if (h->ws) cudaFree(h->ws);
    for (auto& e : h->ev_pool) cudaEventDestroy(e);
    delete h;
}

int kba_set_stream(kba_handle* h, void* s) {
    if (!h) return fail(KBA_ERR_BAD_ARG, "null handle");
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)s;
    h->own_stream = false;
    return KBA_OK;
}

int kba_get_counters(kba_handle* h, kba_counters* out, int reset) {
    if (!h || !out) return fail(KBA_ERR_BAD_ARG, "null argument");
    const Counters& c = h->counters;
    out->launches_total = c.launches_total; out->launches_jacobian = c.launches_jacobian; out->launches_prep = c.launches_prep;
    out->launches_schur = c.launches_schur; out->launches_solve = c.launches_solve; out->launches_backsub = c.launches_backsub;
    out->launches_cost = c.launches_cost; out->launches_update = c.launches_update; out->launches_trim = c.launches_trim;
    out->ms_jacobian = c.ms_jacobian; out->jacobian_obs = c.jacobian_obs;
    if (reset) h->counters = Counters();
    return KBA_OK;
}

int kba_enable_kernel_timing(kba_handle* h, int on) {
    if (!h) return fail(KBA_ERR_BAD_ARG, "null handle");
    h->kernel_timing = on != 0;
    return KBA_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
int kba_batch_create(kba_handle* h, int32_t n_windows, const kba_window* w, kba_batch** out) {
    if (!h || !w || !out || n_windows <= 0) return fail(KBA_ERR_BAD_ARG, "bad argument to kba_batch_create");
    CU(cudaSetDevice(h->device));
    std::string why;
    for (int i = 0; i < n_windows; ++i) {
        const int rc = validate_window(&w[i], why);
        if (rc != KBA_OK) return fail(rc, "window " + std::to_string(i) + ": " + why);
        if (w[i].n_kf < 3 && !w[i].landmarks_fixed)  // solve() needs 3 keyframes (cpp:630); adjustPoseOnly() has one
            return fail(KBA_ERR_NOT_ENOUGH_KF, "window " + std::to_string(i) + ": fewer than 3 keyframes");
    }
    kba_batch* b = new kba_batch();
    b->h = h;
    BatchDev& bd = b->bd;
    bd.n_win = n_windows;
    b->desc_h.resize(n_windows);
    long long kf = 0, cam = 0, lm = 0, obs = 0, chunks = 0, soff = 0, gp = 0, groups = 0;
    int max_rows = 0, max_rows_free = 0, max_groups = 1;
    int nr_cap_max = 64;
    for (int i = 0; i < n_windows; ++i) {
        WinDesc& d = b->desc_h[i];
        memset(&d, 0, sizeof d);
        d.n_kf = w[i].n_kf; d.n_cam = w[i].n_cam; d.n_lm = w[i].n_lm; d.n_obs = w[i].n_obs; d.n_gp = w[i].n_gp;
        d.kf_off = (int)kf; d.cam_off = (int)cam; d.lm_off = (int)lm; d.obs_off = (int)obs; d.gp_off = (int)gp;
        d.chunk_off = (int)chunks; d.n_chunks = (w[i].n_lm + 31) / 32;
        d.grp_off = (int)groups; d.n_groups = (w[i].n_lm + 7) / 8;
        groups += d.n_groups; max_groups = std::max(max_groups, d.n_groups);
        max_rows_free = std::max(max_rows_free, window_rows(w[i]));
        d.scale_kf0 = w[i].scale_kf0; d.scale_kf1 = w[i].scale_kf1;
        d.scale_weight = w[i].scale_weight; d.scale_value = w[i].scale_value;
        const bool planes = w[i].n_gp > 0 || w[i].plane_reg_weight > 0;
        const int rows = (planes ? 10 : 6) * w[i].n_kf + 1;
        max_rows = std::max(max_rows, rows);
        d.nr_cap = ((rows + 63) / 64) * 64;
        d.plane_reg_weight = w[i].plane_reg_weight; d.plane_dist_fixed = w[i].plane_dist_fixed;
        d.s_off = soff;
        soff += (long long)d.nr_cap * d.nr_cap;
        nr_cap_max = std::max(nr_cap_max, d.nr_cap);
        kf += w[i].n_kf; cam += w[i].n_cam; lm += w[i].n_lm; obs += w[i].n_obs; chunks += d.n_chunks; gp += w[i].n_gp;
        bd.max_obs = std::max(bd.max_obs, w[i].n_obs); bd.max_lm = std::max(bd.max_lm, w[i].n_lm);
        bd.max_kf = std::max(bd.max_kf, w[i].n_kf); bd.max_gp = std::max(bd.max_gp, w[i].n_gp);
    }
    if (obs > 2000000000LL) { b->release(); delete b; return fail(KBA_ERR_CAPACITY, "batch exceeds 2^31 observations"); }
    if (nr_cap_max > 640) {  // shared-memory budget of the panel copies in k_reduced_solve / k_chol_trail (227 KB per CTA)
        b->release();
        delete b;
        return fail(KBA_ERR_CAPACITY, "reduced system larger than 640 rows (106 keyframes, or 63 with ground-plane blocks)");
    }
    bd.tot_kf = kf; bd.tot_cam = cam; bd.tot_lm = lm; bd.tot_obs = obs; bd.tot_chunks = (int)chunks; bd.tot_gp = gp;
    bd.tot_groups = (int)groups;
    bd.nr_cap_max = nr_cap_max;
    b->lc.nr_cap_max = nr_cap_max;
    // split the landmark chunks of each window over several CTAs when the batch alone cannot fill the GPU
    {
        const int nb = nr_cap_max / 64, pairs = nb * (nb + 1) / 2;
        int max_chunks = 1;
        for (auto& d : b->desc_h) max_chunks = std::max(max_chunks, d.n_chunks);
        int p = std::min(16, (6 * h->sm_count + n_windows * pairs - 1) / (n_windows * pairs));
        b->lc.small_syrk = (max_rows <= 184);
        if (b->lc.small_syrk) p = (h->sm_count + n_windows - 1) / n_windows;  // one CTA per SM, each owning all tiles
        bd.p_split = std::max(1, std::min(p, max_chunks));
        // fused small-window path (kba_schur_fused.cuh): no J_l, no global V panels; KBA_FUSED=0 keeps the round-1 kernels
        const char* fe = std::getenv("KBA_FUSED");
        bd.fused = (b->lc.small_syrk && bd.max_kf <= kFusedMaxKf && !(fe && std::atoi(fe) == 0)) ? 1 : 0;
        // KBA_P_SPLIT pins the number of CTAs a window's landmark groups are split over.  The partial Schur sums are folded in
        // a fixed order, so results are bit-reproducible for a given split; the default split follows the batch size.
        if (const char* pe = std::getenv("KBA_P_SPLIT")) { if (std::atoi(pe) > 0) p = std::atoi(pe); }
        // (a CTA of the fused Schur kernel takes at least 8 landmark groups of a window, see schur_split in kba_schur_fused.cuh: the
        // partition of a window then does not depend on how many CTAs the batch was given)
        if (bd.fused) bd.p_split = std::max(1, std::min(p, max_groups));
        else if (std::getenv("KBA_P_SPLIT")) bd.p_split = std::max(1, std::min(p, max_chunks));
        b->lc.fused_slots = (max_rows_free <= 176) ? 6 : 7;
    }
    // cost partials: one per CTA of k_linearize (8 warp tiles each, kba_linearize.cuh: lin_tile_bound) or per 256-observation tile of k_eval_obs
    bd.cost_parts = std::max((bd.max_obs + 255) / 256, (bd.max_obs / 16 + bd.max_lm / 32 + 4 + 7) / 8);
    { const char* le = std::getenv("KBA_LINEARIZE"); b->lc.lin_fused = !(le && std::atoi(le) == 0); }
    { const char* le = std::getenv("KBA_LIN_BLOCKS"); if (le && std::atoi(le) == 3) b->lc.lin_blocks = 3; }
    { const char* le = std::getenv("KBA_LIN_GRID"); if (le) b->lc.lin_grid = std::max(-1, std::atoi(le)); }
    { const char* le = std::getenv("KBA_BS_GRID"); if (le) b->lc.bs_grid = std::max(-1, std::atoi(le)); }
    {  // tuning knobs of the residual/Jacobian kernel (defaults measured on B200)
        auto knob = [](const char* name, int dflt) { const char* e = std::getenv(name); return e ? std::atoi(e) : dflt; };
        bd.eval_tiles_jac = std::max(1, knob("KBA_EVAL_TILES_JAC", 8));
        bd.eval_tiles_cost = std::max(1, knob("KBA_EVAL_TILES_COST", 8));
        bd.eval_min_blocks = knob("KBA_EVAL_MIN_BLOCKS", 2);
        bd.eval_cs = knob("KBA_EVAL_CS", 0);
        bd.solve_row_major = knob("KBA_SOLVE_ROW_MAJOR", 0);
        bd.solve_tiled = (nr_cap_max <= 192 && !bd.solve_row_major) ? 1 : 0;
        // a single SM's FP64 rate bounds the one-CTA factorisation of a large system: with few windows spread it
        const int split_dflt = (!bd.solve_tiled && n_windows <= 16) ? std::max(1, std::min(32, h->sm_count / n_windows)) : 0;
        bd.solve_split = bd.solve_tiled ? 0 : knob("KBA_SOLVE_SPLIT", split_dflt);
    }
    bd.bs_parts = (bd.max_lm + 15) / 16;
    {   // device-side packing: fused batches whose landmark keys fit the sort (KBA_DEVICE_PACK=0: host packing as in round 1)
        const char* pe = std::getenv("KBA_DEVICE_PACK");
        b->device_pack = bd.fused && !g_force_host_pack && bd.max_lm <= pack_max_landmarks() && !(pe && std::atoi(pe) == 0);
    }
    const bool hp = !b->device_pack;  // pinned host mirrors of the sorted layout are only needed when the host builds it
    int bad = 0;
    bad |= b->desc.alloc(n_windows, true);
    bad |= b->pose0.alloc(7 * kf, true); bad |= b->plane0.alloc(4 * kf, true); bad |= b->kf_fixed.alloc(kf, true);
    bad |= b->cam.alloc(kCamStride * cam, true);
    bad |= b->lm0.alloc(3 * lm, hp); bad |= b->lm_weight.alloc(lm, hp); bad |= b->lm_ptr.alloc(lm + n_windows, hp);
    bad |= b->obs_kf.alloc(obs, hp); bad |= b->obs_cam.alloc(obs, hp); bad |= b->obs_lm.alloc(obs, hp);
    bad |= b->obs_u.alloc(obs, hp); bad |= b->obs_v.alloc(obs, hp); bad |= b->obs_d.alloc(obs, hp);
    bad |= b->kf_ptr.alloc(kf + n_windows, hp); bad |= b->pm_lm.alloc(obs, hp); bad |= b->pm_cam.alloc(obs, hp);
    bad |= b->pm_u.alloc(obs, hp); bad |= b->pm_v.alloc(obs, hp); bad |= b->pm_d.alloc(obs, hp);
    bad |= b->chunk_lm0.alloc(chunks, hp); bad |= b->chunk_lm1.alloc(chunks, hp);
    bad |= b->chunk_k0.alloc(chunks, hp); bad |= b->chunk_k1.alloc(chunks, hp);
    bad |= b->lm_orig.alloc(lm, hp); bad |= b->obs_rank.alloc(obs, hp);
    if (hp) bad |= b->obs_orig.alloc(obs, true);
    bad |= b->grp_k0.alloc(groups, hp); bad |= b->grp_k1.alloc(groups, hp);
    if (b->device_pack) {
        bad |= b->r_lm_ptr.alloc(lm + n_windows, true); bad |= b->r_obs_kf.alloc(obs, true); bad |= b->r_obs_cam.alloc(obs, true);
        bad |= b->r_obs_u.alloc(obs, true); bad |= b->r_obs_v.alloc(obs, true); bad |= b->r_obs_d.alloc(obs, true);
        bad |= b->r_lm_pos.alloc(3 * lm, true); bad |= b->r_lm_weight.alloc(lm, true); bad |= b->r_gp_lm.alloc(gp, true);
        bad |= b->lm_user.alloc(3 * lm, true); bad |= b->rej_user.alloc(lm, true);
        bad |= b->dev_alloc(&b->raw.lm_inv, lm);
    }
    bad |= b->dev_alloc(&bd.grp_t0, groups); bad |= b->dev_alloc(&bd.grp_t1, groups); bad |= b->dev_alloc(&bd.grp_rs, groups);
    bad |= b->dev_alloc(&bd.lin_tile, (size_t)(obs / 16) + (size_t)(lm / 32) + 4 * (size_t)n_windows + 4);
#ifdef KBA_PROF
    bad |= b->dev_alloc(&bd.prof, 16);
    if (!bad) cudaMemset(bd.prof, 0, 16 * sizeof(unsigned long long));
#endif
    bad |= b->dev_alloc(&bd.rt[0], (size_t)kPoseStride * kf); bad |= b->dev_alloc(&bd.rt[1], (size_t)kPoseStride * kf);
    {   // dense V panels of the Schur kernels; sized in kba_batch_upload from the chunks' keyframe ranges
        bd.panel_cap = 0;
        bd.vpanel = nullptr;
        bad |= b->dev_alloc(&bd.chunk_poff, chunks); bad |= b->dev_alloc(&bd.chunk_rs, chunks);
    }
    bad |= b->dev_alloc(&bd.chunk_t0, chunks); bad |= b->dev_alloc(&bd.chunk_t1, chunks); bad |= b->dev_alloc(&bd.obs_row, obs);
    bad |= b->state.alloc(n_windows, true); bad |= b->log.alloc((size_t)n_windows * kIterLogCap, true);
    for (int q = 0; q < 2; ++q) { bad |= b->pose_out[q].alloc(7 * kf, true); bad |= b->lm_out[q].alloc(3 * lm, hp); }
    bad |= b->lm_active.alloc(lm, hp); bad |= b->n_active.alloc(1, true); bad |= b->jac_obs.alloc(1, true);
    // device-only scratch
    bad |= b->plane_out[0].alloc(4 * kf, true); bad |= b->plane_out[1].alloc(4 * kf, true);
    bad |= b->gp_lm.alloc(gp, hp); bad |= b->gp_kf.alloc(gp, true); bad |= b->gp_weight.alloc(gp, true); bad |= b->gp_of_lm.alloc(lm, hp); bad |= b->gp_shared.alloc(gp, hp);
    bad |= b->dev_alloc(&bd.gp_lin, 14 * gp); bad |= b->dev_alloc(&bd.vgp, 30 * gp);
    bad |= b->dev_alloc(&bd.gp_cost_x, n_windows); bad |= b->dev_alloc(&bd.gp_cost_c, n_windows);
    bad |= b->dev_alloc(&bd.off_pose, kf); bad |= b->dev_alloc(&bd.off_dir, kf); bad |= b->dev_alloc(&bd.off_dist, kf);
    bad |= b->dev_alloc(&bd.bkf, 27 * kf);
    bad |= b->dev_alloc(&bd.scale_f, (size_t)n_windows * nr_cap_max); bad |= b->dev_alloc(&bd.lambda_f, (size_t)n_windows * nr_cap_max);
    bad |= b->dev_alloc(&bd.grad_f, (size_t)n_windows * nr_cap_max); bad |= b->dev_alloc(&bd.delta_f, (size_t)n_windows * nr_cap_max);
    bad |= b->dev_alloc(&bd.lm_scale, 3 * lm); bad |= b->dev_alloc(&bd.lm_linv, 6 * lm); bad |= b->dev_alloc(&bd.lm_z, 3 * lm);
    bad |= b->dev_alloc(&bd.lm_g, 3 * lm); bad |= b->dev_alloc(&bd.lm_lambda, 3 * lm); bad |= b->dev_alloc(&bd.trim_val, 3 * lm);
    bad |= b->dev_alloc(&bd.trim_reject, lm);
    bad |= b->dev_alloc(&bd.res, 3 * obs); bad |= b->dev_alloc(&bd.jp, 18 * obs);
    if (!bd.fused) bad |= b->dev_alloc(&bd.jl, 9 * obs);  // fused path: J_l is re-formed from J_p by its consumers
    else { bad |= b->dev_alloc(&bd.vobs, 18 * obs); bad |= b->dev_alloc(&bd.lm_run, lm); }  // ... and V is kept per observation, unpadded
    bad |= b->dev_alloc(&bd.cost_part_x, (size_t)n_windows * bd.cost_parts); bad |= b->dev_alloc(&bd.cost_part_c, (size_t)n_windows * bd.cost_parts);
    bad |= b->dev_alloc(&bd.bs_part, (size_t)n_windows * bd.bs_parts * 4);
    bad |= b->dev_alloc(&bd.sred, (size_t)soff * bd.p_split); bad |= b->dev_alloc(&bd.amat, (size_t)soff);
    bad |= b->dev_alloc(&bd.chol_w, (size_t)n_windows * 32 * 32); bad |= b->dev_alloc(&bd.chol_invd, (size_t)n_windows * nr_cap_max);
    if (bad) {
        const std::string msg = std::string("device/pinned allocation failed: ") + cudaGetErrorString(cudaGetLastError());
        b->release(); delete b;
        return fail(KBA_ERR_CUDA, msg);
    }
    bd.desc = b->desc.d; bd.state = b->state.d; bd.log = b->log.d;
    bd.pose0 = b->pose0.d; bd.plane0 = b->plane0.d; bd.pose[0] = b->pose_out[0].d; bd.pose[1] = b->pose_out[1].d;
    bd.kf_fixed = b->kf_fixed.d; bd.cam = b->cam.d;
    bd.lm0 = b->lm0.d; bd.lm[0] = b->lm_out[0].d; bd.lm[1] = b->lm_out[1].d; bd.lm_weight = b->lm_weight.d;
    bd.lm_active = b->lm_active.d; bd.lm_ptr = b->lm_ptr.d;
    bd.obs_kf = b->obs_kf.d; bd.obs_cam = b->obs_cam.d; bd.obs_lm = b->obs_lm.d;
    bd.obs_u = b->obs_u.d; bd.obs_v = b->obs_v.d; bd.obs_d = b->obs_d.d;
    bd.kf_ptr = b->kf_ptr.d; bd.pm_lm = b->pm_lm.d; bd.pm_cam = b->pm_cam.d; bd.pm_u = b->pm_u.d; bd.pm_v = b->pm_v.d; bd.pm_d = b->pm_d.d;
    bd.chunk_lm0 = b->chunk_lm0.d; bd.chunk_lm1 = b->chunk_lm1.d; bd.chunk_k0 = b->chunk_k0.d; bd.chunk_k1 = b->chunk_k1.d;
    bd.lm_orig = b->lm_orig.d;
    bd.obs_rank = b->obs_rank.d;
    bd.grp_k0 = b->grp_k0.d; bd.grp_k1 = b->grp_k1.d;
    bd.plane[0] = b->plane_out[0].d; bd.plane[1] = b->plane_out[1].d;
    bd.gp_lm = b->gp_lm.d; bd.gp_kf = b->gp_kf.d; bd.gp_weight = b->gp_weight.d; bd.gp_of_lm = b->gp_of_lm.d; bd.gp_shared = b->gp_shared.d;
    bd.n_active = b->n_active.d;
    bd.jac_obs = b->jac_obs.d;
    if (b->device_pack) {
        PackRaw& r = b->raw;
        r.lm_ptr = b->r_lm_ptr.d; r.obs_kf = b->r_obs_kf.d; r.obs_cam = b->r_obs_cam.d;
        r.obs_u = b->r_obs_u.d; r.obs_v = b->r_obs_v.d; r.obs_d = b->r_obs_d.d;
        r.lm_pos = b->r_lm_pos.d; r.lm_weight = b->r_lm_weight.d; r.gp_lm = b->r_gp_lm.d; r.obs_orig = nullptr;
    }
    {   // failures from here on must give the allocations back
        cudaError_t e = cudaMemset(bd.jac_obs, 0, sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaEventCreate(&b->ev_a);
        if (e == cudaSuccess) e = cudaEventCreate(&b->ev_b);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->ev_poll, (h->blocking_sync ? cudaEventBlockingSync : 0) | cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->ev_poll2, (h->blocking_sync ? cudaEventBlockingSync : 0) | cudaEventDisableTiming);
        if (e == cudaSuccess && b->loop_pass.alloc(1, true)) e = cudaErrorMemoryAllocation;
        if (e == cudaSuccess) e = configure_kernels(nr_cap_max);
        if (e == cudaSuccess && b->device_pack) e = configure_pack();
        if (e != cudaSuccess) { b->release(); delete b; return fail(KBA_ERR_CUDA, cudaGetErrorString(e)); }
    }
    *out = b;
    const int rc = kba_batch_upload(b, n_windows, w);
    if (rc != KBA_OK) { b->release(); delete b; *out = nullptr; }
    return rc;
}

int kba_batch_upload(kba_batch* b, int32_t n_windows, const kba_window* w) {
    if (!b || !w || n_windows != b->bd.n_win) return fail(KBA_ERR_BAD_ARG, "bad argument to kba_batch_upload");
    CU(cudaSetDevice(b->h->device));  // callers may drive several handles from several host threads
    for (int i = 0; i < n_windows; ++i) {
        const WinDesc& d = b->desc_h[i];
        std::string why;
        const int rc = validate_window(&w[i], why);  // indices of the new contents are checked like at create
        if (rc != KBA_OK) return fail(rc, "kba_batch_upload: window " + std::to_string(i) + ": " + why);
        if (w[i].n_kf != d.n_kf || w[i].n_lm != d.n_lm || w[i].n_obs != d.n_obs || w[i].n_cam != d.n_cam || w[i].n_gp != d.n_gp)
            return fail(KBA_ERR_BAD_ARG, "kba_batch_upload: window shapes (keyframes, cameras, landmarks, observations, ground-plane "
                                         "residuals) differ from kba_batch_create");
        // the reduced system was sized at create: 6 rows per keyframe, 10 with plane blocks
        const bool planes = w[i].n_gp > 0 || w[i].plane_reg_weight > 0;
        if (((planes ? 10 : 6) * w[i].n_kf + 1 + 63) / 64 * 64 > d.nr_cap)
            return fail(KBA_ERR_BAD_ARG, "kba_batch_upload: window " + std::to_string(i) + " needs plane blocks the batch was not created with");
    }
    // packing (landmark sort, observation permutation, keyframe-major copy) is independent per window: host threads
    auto pack = [&](int i) {
        b->desc_h[i].scale_weight = w[i].scale_weight; b->desc_h[i].scale_value = w[i].scale_value;
        b->desc_h[i].scale_kf0 = w[i].scale_kf0; b->desc_h[i].scale_kf1 = w[i].scale_kf1;
        b->desc_h[i].landmarks_fixed = w[i].landmarks_fixed;
        b->desc_h[i].plane_reg_weight = w[i].plane_reg_weight; b->desc_h[i].plane_dist_fixed = w[i].plane_dist_fixed;
        b->desc_h[i].speed_kf = w[i].speed_kf; b->desc_h[i].speed_weight = w[i].speed_weight; b->desc_h[i].speed_dt = w[i].speed_dt;
        memcpy(b->desc_h[i].speed_v_before, w[i].speed_v_before, sizeof(double) * 3);
        memcpy(b->desc_h[i].speed_T_origin_before, w[i].speed_T_origin_before, sizeof(double) * 7);
        b->desc.h[i] = b->desc_h[i];
        if (b->device_pack) fill_window_raw(b, i, &w[i]);
        else fill_window(b, i, &w[i]);
    };
    {
        const char* e = std::getenv("KBA_HOST_THREADS");
        int nt = e ? std::atoi(e) : (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
        nt = std::max(1, std::min(nt, n_windows));
        if (nt == 1) {
            for (int i = 0; i < n_windows; ++i) pack(i);
        } else {
            std::atomic<int> next{0};
            std::vector<std::thread> pool;
            for (int t = 0; t < nt; ++t)
                pool.emplace_back([&] { for (int i = next.fetch_add(1); i < n_windows; i = next.fetch_add(1)) pack(i); });
            for (auto& th : pool) th.join();
        }
    }
    for (int i = 0; i < n_windows; ++i) b->lc.max_rank = std::max(b->lc.max_rank, b->desc_h[i].max_rank);
    if (b->bd.fused) {
        int rows = 0;
        for (int i = 0; i < n_windows; ++i) rows = std::max(rows, window_rows(w[i]));
        b->lc.fused_slots = (rows <= 176) ? 6 : 7;
    }
    if (!b->bd.fused) {   // V panel capacity: per chunk 96 columns x (rows of its keyframe range + right-hand-side tile), see k_solve_begin
        long long need = 0;
        for (int i = 0; i < n_windows; ++i) {
            const WinDesc& d = b->desc_h[i];
            const int rows_per_kf = (w[i].n_gp > 0 || w[i].plane_reg_weight > 0) ? 10 : 6;
            long long tot = 0;
            for (int c = 0; c < d.n_chunks; ++c) {
                const int nk = b->chunk_k1.h[d.chunk_off + c] - b->chunk_k0.h[d.chunk_off + c] + 1;
                if (nk <= 0) continue;
                const int rows = 8 * ((rows_per_kf * nk + 14 + 7) / 8) + 8;
                tot += 96LL * (((rows - 4 + 15) / 16) * 16 + 4);
            }
            need = std::max(need, tot);
        }
        if (need > b->bd.panel_cap) {
            if (b->bd.vpanel) { CU(cudaStreamSynchronize(b->h->stream)); b->dev_free(b->bd.vpanel); b->bd.vpanel = nullptr; }
            b->bd.panel_cap = (need + 1) & ~1LL;
            if (b->dev_alloc(&b->bd.vpanel, (size_t)n_windows * b->bd.panel_cap)) return fail(KBA_ERR_CUDA, "out of device memory (V panels)");
        }
        for (int i = 0; i < n_windows; ++i) { b->desc_h[i].panel_off = (long long)i * b->bd.panel_cap; b->desc.h[i].panel_off = b->desc_h[i].panel_off; }
    }
    cudaStream_t s = b->h->stream;
    if (b->device_pack) {  // the caller's arrays as they are (~33 B per observation), then the packing kernels
        CU(b->desc.upload(s)); CU(b->pose0.upload(s)); CU(b->plane0.upload(s)); CU(b->kf_fixed.upload(s)); CU(b->cam.upload(s));
        CU(b->r_lm_pos.upload(s)); CU(b->r_lm_weight.upload(s)); CU(b->r_lm_ptr.upload(s));
        CU(b->r_obs_kf.upload(s)); CU(b->r_obs_cam.upload(s)); CU(b->r_obs_u.upload(s)); CU(b->r_obs_v.upload(s)); CU(b->r_obs_d.upload(s));
        CU(b->r_gp_lm.upload(s)); CU(b->gp_kf.upload(s)); CU(b->gp_weight.upload(s));
        launch_pack(b->bd, b->raw, s);
        CU(cudaGetLastError());
        const BatchDev& bd = b->bd;
        b->h2d_bytes = sizeof(WinDesc) * bd.n_win + (7 + 4) * 8 * bd.tot_kf + bd.tot_kf + kCamStride * 8 * bd.tot_cam +
                       (3 + 1) * 8 * bd.tot_lm + 4 * (bd.tot_lm + bd.n_win) + (2 * 4 + 3 * 4) * bd.tot_obs + (4 + 4 + 8) * bd.tot_gp;
        return KBA_OK;
    }
    CU(b->desc.upload(s)); CU(b->pose0.upload(s)); CU(b->plane0.upload(s)); CU(b->kf_fixed.upload(s)); CU(b->cam.upload(s));
    CU(b->lm0.upload(s)); CU(b->lm_weight.upload(s)); CU(b->lm_ptr.upload(s));
    CU(b->obs_kf.upload(s)); CU(b->obs_cam.upload(s)); CU(b->obs_lm.upload(s));
    CU(b->obs_u.upload(s)); CU(b->obs_v.upload(s)); CU(b->obs_d.upload(s));
    CU(b->kf_ptr.upload(s)); CU(b->pm_lm.upload(s)); CU(b->pm_cam.upload(s)); CU(b->pm_u.upload(s)); CU(b->pm_v.upload(s)); CU(b->pm_d.upload(s));
    CU(b->chunk_lm0.upload(s)); CU(b->chunk_lm1.upload(s)); CU(b->chunk_k0.upload(s)); CU(b->chunk_k1.upload(s));
    CU(b->grp_k0.upload(s)); CU(b->grp_k1.upload(s));
    CU(b->lm_orig.upload(s)); CU(b->obs_rank.upload(s));
    CU(b->gp_lm.upload(s)); CU(b->gp_kf.upload(s)); CU(b->gp_weight.upload(s)); CU(b->gp_of_lm.upload(s)); CU(b->gp_shared.upload(s));
    const BatchDev& bd = b->bd;
    b->h2d_bytes = sizeof(WinDesc) * bd.n_win + (7 + 4) * 8 * bd.tot_kf + bd.tot_kf + kCamStride * 8 * bd.tot_cam +
                   (3 + 1) * 8 * bd.tot_lm + 4 * (bd.tot_lm + bd.n_win) + (3 * 4 + 3 * 4) * bd.tot_obs +
                   4 * (bd.tot_kf + bd.n_win) + (2 * 4 + 3 * 4) * bd.tot_obs + 8 * bd.tot_chunks + 8 * bd.tot_groups;
    return KBA_OK;
}

static SolveParams make_params(const kba_options* o) {
    SolveParams sp;
    sp.gp_huber = o->gp_huber; sp.gp_quantile = o->gp_quantile;
    sp.depth_thres = o->depth_thres; sp.reprojection_thres = o->reprojection_thres;
    sp.depth_quantile = o->depth_quantile; sp.reprojection_quantile = o->reprojection_quantile;
    sp.function_tolerance = o->function_tolerance; sp.gradient_tolerance = o->gradient_tolerance;
    sp.parameter_tolerance = o->parameter_tolerance; sp.initial_radius = o->initial_trust_region_radius;
    sp.max_radius = o->max_trust_region_radius; sp.min_radius = o->min_trust_region_radius;
    sp.min_relative_decrease = o->min_relative_decrease; sp.min_lm_diagonal = o->min_lm_diagonal;
    sp.max_lm_diagonal = o->max_lm_diagonal; sp.trim_solver_iterations = o->trim_solver_iterations;
    sp.final_solver_iterations = o->final_solver_iterations; sp.min_residual_groups = o->min_residual_groups;
    sp.max_consecutive_invalid_steps = o->max_consecutive_invalid_steps;
    sp.max_solver_time = o->solver_time_sec;
    return sp;
}

int kba_batch_solve(kba_batch* b, const kba_options* opt) {
    if (!b || !opt) return fail(KBA_ERR_BAD_ARG, "bad argument to kba_batch_solve");
    if (opt->precision != 0 && opt->precision != 1) return fail(KBA_ERR_BAD_ARG, "kba_options.precision must be 0 (FP64) or 1 (FP32 linearisation)");
    if (opt->num_trim_rounds > 6 || (opt->num_trim_rounds < 0 && opt->num_rounds_option > 6))
        return fail(KBA_ERR_CAPACITY, "at most 6 trimming rounds (KBA_MAX_SOLVES = 8 inner solves incl. one retry and the final solve)");
    b->bd.precision = opt->precision;
    b->bd.lin1 = (b->bd.fused && b->lc.lin_fused && opt->precision == 0 && b->lc.max_rank == 0) ? 1 : 0;
    kba_handle* h = b->h;
    CU(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const SolveParams sp = make_params(opt);
    LaunchCfg lc = b->lc;
    lc.rounds_override = opt->num_trim_rounds;
    lc.min_landmarks_for_trimming = opt->min_landmarks_for_trimming;
    lc.num_rounds_option = opt->num_rounds_option;
    lc.time_jacobian = h->kernel_timing;
    if (h->kernel_timing && h->ev_pool.empty()) {
        h->ev_pool.resize(1024);
        for (auto& e : h->ev_pool) CU(cudaEventCreate(&e));
    }
    h->ev_used = 0;
    lc.ev_pool = h->ev_pool.data(); lc.ev_cap = (int)h->ev_pool.size(); lc.ev_used = &h->ev_used;
    CU(cudaMemsetAsync(b->bd.jac_obs, 0, sizeof(unsigned long long), s));
    // the evaluation kernels write the cost slots of the CTAs they launch and rely on the others being zero; the slot layout
    // differs between k_linearize and k_eval_obs (a batch may be solved with either: kba_options.precision)
    CU(cudaMemsetAsync(b->bd.cost_part_x, 0, sizeof(double) * (size_t)b->bd.n_win * b->bd.cost_parts, s));
    CU(cudaMemsetAsync(b->bd.cost_part_c, 0, sizeof(double) * (size_t)b->bd.n_win * b->bd.cost_parts, s));
    CU(cudaEventRecord(b->ev_a, s));
    launch_reset(b->bd, lc, s);
    // upper bound on passes: every solve needs (iterations + 2) passes, plus one pass per trimming step
    const int rounds_max = 7;
    const int max_passes = rounds_max * (3 * opt->trim_solver_iterations + 4) + opt->final_solver_iterations + 8;
    const auto t0 = std::chrono::steady_clock::now();
    int check_every = 4;
    bool timed_out = false, poll_pending = false;
    // ---- graph paths (see solve_graph_mode) ----
    int gmode = solve_graph_mode();
    if (h->kernel_timing || launch_check_enabled() || s == nullptr || b->sg.unusable) gmode = 0;
    // Sharded solve: the NCCL all-reduces are captured with the kernels (flat graph, mode 1).  The first solve of a batch runs on
    // the stream so that NCCL sets its connections up outside a capture; the active-window count is read BEFORE the next graph is
    // launched (no look-ahead): it is identical on all ranks, and every rank must launch the same number of graphs.
    const bool lockstep = b->bd.sharded != 0;
    if (lockstep) gmode = (gmode && shard_graph_enabled() && b->solves_done > 0) ? 1 : 0;
    if (gmode) {
        std::vector<unsigned char> key;
        key.reserve(sizeof(BatchDev) + sizeof(SolveParams) + 64);
        key_append(key, b->bd); key_append(key, sp); key_append(key, gmode); key_append(key, max_passes); key_append(key, s);
        key_append(key, lc.nr_cap_max); key_append(key, lc.max_rank); key_append(key, (int)lc.small_syrk); key_append(key, (int)lc.lin_fused);
        key_append(key, lc.lin_blocks); key_append(key, lc.fused_slots); key_append(key, lc.xchg.user);
        key_append(key, lc.lin_grid); key_append(key, lc.bs_grid);
        if (!(b->sg.exec && b->sg.key == key) && !build_solve_graph(b, sp, lc, gmode, max_passes, check_every, key)) gmode = 0;
    }
    if (gmode == 2) {
        CU(cudaMemsetAsync(b->loop_pass.d, 0, sizeof(int), s));
        CU(cudaGraphLaunch(b->sg.exec, s));
        CU(b->loop_pass.download(s));
        CU(cudaEventRecord(b->ev_b, s));
        CU(b->jac_obs.download(s));
        CU(wait_stream(h));
        CU(cudaGetLastError());
        CU(cudaEventElapsedTime(&b->last_solve_ms, b->ev_a, b->ev_b));
        h->counters.jacobian_obs += (long long)b->jac_obs.h[0];
        add_pass_counters(h->counters, b->sg.per_pass, b->loop_pass.h[0]);
        h->counters.launches_total += b->loop_pass.h[0];  // k_loop_cond
        b->solves_done++;
        return KBA_OK;
    }
    if (gmode == 1) {
        cudaEvent_t evs[2] = {b->ev_poll, b->ev_poll2};
        const int n_launch = (max_passes + check_every - 1) / check_every;
        int launched = 0;
        for (int g = 0; g < n_launch; ++g) {
            CU(cudaGraphLaunch(b->sg.exec, s));
            ++launched;
            CU(cudaEventRecord(evs[g & 1], s));
            if (lockstep) {
                CU(wait_event(h, evs[g & 1]));
                if (b->n_active.h[0] == 0) break;
            } else if (g >= 1) {  // the count of the previous launch is read while this one runs (the device never waits for the host)
                CU(wait_event(h, evs[(g - 1) & 1]));
                if (b->n_active.h[0] == 0) break;
                if (opt->solver_time_sec > 0) {
                    const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                    if (el > KBA_MAX_SOLVES * opt->solver_time_sec + 2.0) { timed_out = true; break; }
                }
            }
        }
        if (timed_out) g_last_error = "kba_batch_solve: host safety cap reached, unfinished windows carry KBA_ERR_TIMEOUT in kba_result.status";
        CU(cudaEventRecord(b->ev_b, s));
        CU(b->jac_obs.download(s));
        CU(wait_stream(h));
        CU(cudaGetLastError());
        CU(cudaEventElapsedTime(&b->last_solve_ms, b->ev_a, b->ev_b));
        h->counters.jacobian_obs += (long long)b->jac_obs.h[0];
        add_pass_counters(h->counters, b->sg.per_pass, (long long)launched * check_every);
        h->counters.launches_total += launched;  // k_count_active
        b->solves_done++;
        return KBA_OK;
    }
    for (int pass = 0; pass < max_passes; ++pass) {
        if (launch_pass(b->bd, sp, lc, &h->counters, s)) {  // message set by the exchange
            cudaEventRecord(b->ev_b, s);
            return KBA_ERR_NCCL;
        }
        if (pass < 2) {  // a launch that fails fails in the first pass: do not let the windows spin to the safety cap
            const cudaError_t le = cudaGetLastError();
            if (le != cudaSuccess) { cudaEventRecord(b->ev_b, s); cudaStreamSynchronize(s); return fail(KBA_ERR_CUDA, std::string("kernel launch failed in kba_batch_solve: ") + cudaGetErrorString(le)); }
        }
        // Completion check without draining the queue: every `check_every` passes the active-window count is copied out behind an
        // event, the next passes are enqueued at once, and the count is READ one check later.  The device never waits for the
        // host (a synchronous poll emptied the queue 12-15 times per solve: ~1 ms of a 14 ms single-window solve); the price is
        // up to `check_every` passes enqueued after the last window finished, in which every kernel exits at once.  In a sharded
        // solve the state -- hence the count -- is identical on all ranks, so they still issue the same passes.
        if ((pass + 1) % check_every == 0 || pass + 1 == max_passes) {
            if (poll_pending) {
                CU(wait_event(h, b->ev_poll));
                poll_pending = false;
                if (b->n_active.h[0] == 0) break;
                if (opt->solver_time_sec > 0 && !b->bd.sharded) {  // host safety cap, see below
                    const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                    if (el > KBA_MAX_SOLVES * opt->solver_time_sec + 2.0) { timed_out = true; break; }
                }
            }
            launch_count_active(b->bd, s);
            CU(b->n_active.download(s));
            CU(cudaEventRecord(b->ev_poll, s));
            poll_pending = true;
            // max_solver_time_in_seconds is applied PER INNER SOLVE on the device (k_lm_update, like ceres), which always
            // lets the final solve start; the host only guards against a stuck device with the budget of every possible
            // inner solve.  A sharded solve is collective: its ranks must issue the same passes, so no rank may leave on
            // its own clock (the iteration caps bound it).
        }
    }
    if (timed_out) g_last_error = "kba_batch_solve: host safety cap reached, unfinished windows carry KBA_ERR_TIMEOUT in kba_result.status";
    CU(cudaEventRecord(b->ev_b, s));
    CU(b->jac_obs.download(s));
    CU(wait_stream(h));
    CU(cudaGetLastError());
    CU(cudaEventElapsedTime(&b->last_solve_ms, b->ev_a, b->ev_b));
    h->counters.jacobian_obs += (long long)b->jac_obs.h[0];
    for (int i = 0; i + 1 < h->ev_used; i += 2) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, h->ev_pool[i], h->ev_pool[i + 1]));
        h->counters.ms_jacobian += ms;
    }
    b->solves_done++;
    return KBA_OK;
}

int kba_batch_set_shard(kba_batch* b, kba_shard_comm* comm, int32_t lm_begin, int32_t lm_total) {
    if (!b || !comm) return fail(KBA_ERR_BAD_ARG, "bad argument to kba_batch_set_shard");
    BatchDev& bd = b->bd;
    if (bd.n_win != 1) return fail(KBA_ERR_BAD_ARG, "a sharded batch holds exactly one window (this rank's shard)");
    if (bd.tot_gp > 0) return fail(KBA_ERR_CAPACITY, "sharded windows with ground-plane residuals are not supported");
    if (lm_begin < 0 || lm_total < lm_begin + (int)bd.tot_lm) return fail(KBA_ERR_BAD_ARG, "landmark block outside the window");
    if (bd.sharded) return fail(KBA_ERR_BAD_ARG, "kba_batch_set_shard called twice");
    const WinDesc& d = b->desc_h[0];
    int bad = 0;
    const kba::Exchange xchg = kba_shard_exchange(comm);
    const size_t n_x = (size_t)d.nr_cap * d.nr_cap + (size_t)d.n_kf * 27 + (size_t)bd.cost_parts + 2;
    bad |= b->dev_alloc(&bd.xs, 16 + (size_t)xchg.world);
    bad |= b->dev_alloc(&bd.trim_send, 3 * (size_t)lm_total); bad |= b->dev_alloc(&bd.trim_glob, 3 * (size_t)lm_total);
    bad |= b->dev_alloc(&bd.reject_glob, (size_t)lm_total);
    bad |= b->dev_alloc(&bd.x_send, n_x); bad |= b->dev_alloc(&bd.x_recv, n_x);
    if (bad) return fail(KBA_ERR_CUDA, "out of device memory (shard buffers)");
    cudaStream_t s = b->h->stream;
    CU(cudaMemsetAsync(bd.xs, 0, (16 + (size_t)xchg.world) * sizeof(double), s));
    CU(cudaMemsetAsync(bd.trim_send, 0, 3 * (size_t)lm_total * sizeof(double), s));
    CU(cudaMemsetAsync(bd.x_send, 0, n_x * sizeof(double), s));
    CU(cudaMemsetAsync(bd.x_recv, 0, n_x * sizeof(double), s));
    bd.shard_rank = xchg.rank; bd.shard_world = xchg.world;
    bd.sharded = 1; bd.lm_begin = lm_begin; bd.lm_total = lm_total;
    b->lc.xchg = xchg;
    b->lc.shard_win.nr_cap = d.nr_cap; b->lc.shard_win.n_kf = d.n_kf;
    return KBA_OK;
}

int kba_batch_download(kba_batch* b, kba_result* res) {
    if (!b || !res) return fail(KBA_ERR_BAD_ARG, "bad argument to kba_batch_download");
    CU(cudaSetDevice(b->h->device));
    cudaStream_t s = b->h->stream;
    CU(b->state.download(s)); CU(b->log.download(s));
    CU(b->pose_out[0].download(s)); CU(b->pose_out[1].download(s));
    const BatchDev& bd = b->bd;
    if (b->device_pack) {  // landmarks come back in the caller's order: one kernel, one copy per array
        launch_unpack_landmarks(bd, b->lm_user.d, b->rej_user.d, s);
        CU(b->lm_user.download(s)); CU(b->rej_user.download(s));
    } else {
        CU(b->lm_out[0].download(s)); CU(b->lm_out[1].download(s)); CU(b->lm_active.download(s));
    }
    CU(b->plane_out[0].download(s)); CU(b->plane_out[1].download(s));
    CU(wait_stream(b->h));
    b->d2h_bytes = sizeof(WinState) * bd.n_win + 2 * (7 + 4) * 8 * bd.tot_kf + (b->device_pack ? 1 : 2) * 3 * 8 * bd.tot_lm + bd.tot_lm;
    for (int i = 0; i < bd.n_win; ++i) {
        const WinDesc& d = b->desc_h[i];
        const WinState& st = b->state.h[i];
        kba_result& r = res[i];
        const int cur = st.cur;
        if (r.kf_pose) memcpy(r.kf_pose, b->pose_out[cur].h + 7 * (size_t)d.kf_off, sizeof(double) * 7 * d.n_kf);
        if (r.kf_plane) memcpy(r.kf_plane, b->plane_out[cur].h + 4 * (size_t)d.kf_off, sizeof(double) * 4 * d.n_kf);
        if (b->device_pack) {
            if (r.lm_pos) memcpy(r.lm_pos, b->lm_user.h + 3 * (size_t)d.lm_off, 3 * sizeof(double) * d.n_lm);
            if (r.lm_rejected) memcpy(r.lm_rejected, b->rej_user.h + d.lm_off, d.n_lm);
        } else {
            const int* orig = b->lm_orig.h + d.lm_off;
            if (r.lm_pos)
                for (int j = 0; j < d.n_lm; ++j) memcpy(r.lm_pos + 3 * (size_t)orig[j], b->lm_out[cur].h + 3 * (size_t)(d.lm_off + j), 3 * sizeof(double));
            if (r.lm_rejected) for (int j = 0; j < d.n_lm; ++j) r.lm_rejected[orig[j]] = !b->lm_active.h[d.lm_off + j];
        }
        r.num_solves = st.n_solves;
        for (int q = 0; q < st.n_solves && q < KBA_MAX_SOLVES; ++q) {
            const SolveSummary& ss = st.solves[q];
            kba_solve_summary& o = r.solves[q];
            o.initial_cost = ss.initial_cost; o.final_cost = ss.final_cost; o.num_iterations = ss.num_iterations;
            o.num_successful_steps = ss.num_successful_steps; o.termination = ss.termination;
            o.num_landmarks = ss.num_landmarks; o.num_residual_blocks = ss.num_residual_blocks; o.reserved_ = 0;
        }
        r.initial_cost = st.n_solves > 0 ? st.solves[0].initial_cost : 0.0;
        r.final_cost = st.n_solves > 0 ? st.solves[st.n_solves - 1].final_cost : 0.0;
        r.status = (st.phase == PH_DONE) ? KBA_OK : KBA_ERR_TIMEOUT;  // host safety cap hit (see kba_batch_solve)
        r.time_sec = 1e-3 * b->last_solve_ms;
        int n = 0;
        if (r.iterations) {
            for (; n < st.log_n && n < r.iterations_capacity; ++n) {
                const IterRecord& e = b->log.h[(size_t)i * kIterLogCap + n];
                kba_iteration& o = r.iterations[n];
                o.cost = e.cost; o.cost_change = e.cost_change; o.gradient_max_norm = e.gradient_max_norm;
                o.step_norm = e.step_norm; o.relative_decrease = e.relative_decrease; o.trust_region_radius = e.radius;
                o.iteration = e.iteration; o.solve_index = e.solve_index; o.step_is_valid = e.valid; o.step_is_successful = e.successful;
            }
        }
        r.num_iteration_records = n;
    }
    return KBA_OK;
}

int kba_batch_transfer_bytes(kba_batch* b, int64_t* h2d, int64_t* d2h) {
    if (!b) return fail(KBA_ERR_BAD_ARG, "null batch");
    if (h2d) *h2d = (int64_t)b->h2d_bytes;
    if (d2h) *d2h = (int64_t)b->d2h_bytes;
    return KBA_OK;
}

void kba_batch_destroy(kba_batch* b) {
    if (!b) return;
    cudaStreamSynchronize(b->h->stream);
#ifdef KBA_PROF
    if (b->bd.prof) {
        unsigned long long c[16];
        cudaMemcpy(c, b->bd.prof, sizeof c, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[kba prof] fused Schur kernel, cycles summed over warps: consumers wait %llu multiply %llu | producers "
                "wait-empty %llu set-up %llu copies %llu\n", c[0], c[1], c[4], c[5], c[6]);
    }
#endif
    b->release();
    delete b;
}

int kba_batch_jacobian_pass(kba_batch* b, const kba_options* opt, int32_t repeats, float* ms_out) {
    if (!b || !opt || repeats < 1) return fail(KBA_ERR_BAD_ARG, "bad argument to kba_batch_jacobian_pass");
    kba_handle* h = b->h;
    cudaStream_t s = h->stream;
    const SolveParams sp = make_params(opt);
    if (opt->precision != 0 && opt->precision != 1) return fail(KBA_ERR_BAD_ARG, "kba_options.precision must be 0 or 1");
    b->bd.precision = opt->precision;
    launch_reset(b->bd, b->lc, s);
    launch_force_linearize(b->bd, s);
    CU(cudaEventRecord(b->ev_a, s));
    for (int i = 0; i < repeats; ++i) launch_jacobian_only(b->bd, sp, s);
    CU(cudaEventRecord(b->ev_b, s));
    CU(cudaStreamSynchronize(s));
    CU(cudaGetLastError());
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, b->ev_a, b->ev_b));
    if (ms_out) *ms_out = ms;
    h->counters.launches_total += repeats; h->counters.launches_jacobian += repeats;
    h->counters.ms_jacobian += ms; h->counters.jacobian_obs += (long long)repeats * b->bd.tot_obs;
    CU(cudaMemsetAsync(b->bd.jac_obs, 0, sizeof(unsigned long long), s));
    return KBA_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
int kba_solve_batch(kba_handle* h, int32_t n_windows, const kba_window* w, const kba_options* opt, kba_result* res) {
    if (!h || !w || !opt || !res) return fail(KBA_ERR_BAD_ARG, "null argument to kba_solve_batch");
    kba_batch* b = nullptr;
    int rc = kba_batch_create(h, n_windows, w, &b);
    if (rc != KBA_OK) return rc;
    rc = kba_batch_solve(b, opt);
    if (rc == KBA_OK) rc = kba_batch_download(b, res);
    kba_batch_destroy(b);
    return rc;
}

int kba_solve_window(kba_handle* h, const kba_window* w, const kba_options* opt, kba_result* res) {
    return kba_solve_batch(h, 1, w, opt, res);
}

int kba_eval(kba_handle* h, const kba_window* w, const kba_options* opt, kba_eval_out* out) {
    if (!h || !w || !opt || !out) return fail(KBA_ERR_BAD_ARG, "null argument to kba_eval");
    kba_batch* b = nullptr;
    g_force_host_pack = true;  // the inspection entry point maps observations back to the caller's order on the host
    int rc = kba_batch_create(h, 1, w, &b);
    g_force_host_pack = false;
    if (rc != KBA_OK) return rc;
    float ms;
    rc = kba_batch_jacobian_pass(b, opt, 1, &ms);
    if (rc != KBA_OK) { kba_batch_destroy(b); return rc; }
    const BatchDev& bd = b->bd;
    const size_t n = (size_t)w->n_obs;
    std::vector<double> res_h(3 * n + 1), jp_h(18 * n + 1), jl_h(9 * n + 1), cost_h(bd.cost_parts);
    std::vector<int> offp(w->n_kf);
    WinState st;
    cudaStream_t s = h->stream;
    cudaError_t ce = cudaSuccess;
    auto cp = [&](void* dst, const void* src, size_t bytes) {
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s);
    };
    cp(res_h.data(), bd.res, 3 * n * sizeof(double));
    cp(jp_h.data(), bd.jp, 18 * n * sizeof(double));
    double* jl_dev = bd.jl;
    if (bd.fused) {  // J_l is not materialised on the fused path: expand it on the device the way its consumers do
        if (b->dev_alloc(&jl_dev, 9 * n)) { kba_batch_destroy(b); return fail(KBA_ERR_CUDA, "out of device memory (kba_eval)"); }
        launch_expand_jl(bd, jl_dev, s);
    }
    cp(jl_h.data(), jl_dev, 9 * n * sizeof(double));
    cp(cost_h.data(), bd.cost_part_x, bd.cost_parts * sizeof(double));
    cp(offp.data(), bd.off_pose, w->n_kf * sizeof(int));
    cp(&st, bd.state, sizeof(WinState));
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
    if (ce != cudaSuccess) { kba_batch_destroy(b); return fail(KBA_ERR_CUDA, cudaGetErrorString(ce)); }
    for (size_t e = 0; e < n; ++e) {  // e: internal (sorted) observation slot, o: the caller's observation index
        const size_t o = (size_t)b->obs_orig.h[e];
        const bool fixed = offp[w->obs_kf[o]] < 0;
        // precision 1: the streams hold floats (same component-major layout)
        auto at = [&](const std::vector<double>& v, size_t idx) {
            return opt->precision ? (double)reinterpret_cast<const float*>(v.data())[idx] : v[idx];
        };
        if (out->residual) for (int q = 0; q < 3; ++q) out->residual[3 * o + q] = at(res_h, q * n + e);
        if (out->jac_pose) for (int q = 0; q < 18; ++q) out->jac_pose[18 * o + q] = fixed ? 0.0 : at(jp_h, q * n + e);
        if (out->jac_lm) for (int q = 0; q < 9; ++q) out->jac_lm[9 * o + q] = at(jl_h, q * n + e);
    }
    if (out->cost) { double c = 0; for (int q = 0; q < bd.cost_parts; ++q) c += cost_h[q]; out->cost[0] = c; }
    if (out->failed) out->failed[0] = st.eval_failed;
    kba_batch_destroy(b);
    return KBA_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// persistent, device-resident window (include/kba_b200.h, kba_track_*)
// ---------------------------------------------------------------------------------------------------------------------
void kba_track_destroy(kba_track* t) {
    if (!t) return;
    cudaStreamSynchronize(t->h->stream);
    if (t->batch) kba_batch_destroy(t->batch);
    for (void* p : t->dev) cudaFree(p);
    t->p_lm.release(); t->p_cam.release(); t->sel_kf.release(); t->sel_lm.release(); t->lay.release();
    t->p_u.release(); t->p_v.release(); t->p_d.release(); t->sel_fixed.release(); t->p_dbl.release(); t->p_slot.release();
    delete t;
}

int kba_track_create(kba_handle* h, const kba_track_caps* c, int32_t n_cam, const double* cam_intr, const double* cam_pose, kba_track** out) {
    if (!h || !c || !out || !cam_intr || !cam_pose || n_cam < 1) return fail(KBA_ERR_BAD_ARG, "bad argument to kba_track_create");
    if (c->max_keyframes < 3 || c->max_landmarks < 1 || c->max_measurements < 1 || c->win_keyframes < 3 || c->win_landmarks < 1 ||
        c->win_observations < 1 || c->win_ground < 0 || c->win_ground > c->win_landmarks)
        return fail(KBA_ERR_BAD_ARG, "kba_track_create: capacities");
    if ((c->win_ground > 0 ? 10 : 6) * c->win_keyframes + 1 > 184 || c->win_keyframes > kFusedMaxKf || c->win_landmarks > pack_max_landmarks())
        return fail(KBA_ERR_CAPACITY, "kba_track_create: the stored window must fit the fused path (<= 184 reduced rows: 30 keyframes, "
                                      "18 with ground-plane blocks; <= 32768 landmarks) -- larger windows go through kba_solve_window");
    CU(cudaSetDevice(h->device));
    kba_track* t = new kba_track();
    t->h = h; t->caps = *c; t->n_cam = n_cam;
    // capacity batch from a dummy window of the largest shape.  The observations are spread evenly over the landmarks and, within
    // a landmark, over the keyframes in ascending order: the packing kernels run once on this window (create = upload), and their
    // per-landmark loops (insertion sort of a track, k_track_sort / k_pack_obs) are written for tracks of a few dozen
    // observations -- one landmark carrying all 2^18 of them kept a single GPU thread busy for minutes.
    {
        const int K = c->win_keyframes, L = c->win_landmarks, O = c->win_observations, G = c->win_ground;
        std::vector<double> pose(7 * (size_t)K, 0.0), plane(4 * (size_t)K, 0.0), lmp(3 * (size_t)L, 0.0), lmw(L, 1.0), gw(std::max(G, 1), 1.0);
        std::vector<uint8_t> fixed(K, 0);
        std::vector<int32_t> ptr(L + 1, O), okf(O, 1), gl(std::max(G, 1), 0), gk(std::max(G, 1), 1);
        std::vector<float> u(O, 0.f), v(O, 0.f), d(O, -1.f);
        for (int k = 0; k < K; ++k) { pose[7 * k] = 1.0; plane[4 * k + 2] = 1.0; }
        for (int j = 0; j < L; ++j) lmp[3 * j + 2] = 10.0;
        for (int g = 0; g < G; ++g) gl[g] = g;
        fixed[0] = 1;
        for (int j = 0; j <= L; ++j) ptr[j] = (int32_t)(((long long)O * j) / L);
        for (int j = 0; j < L; ++j) {
            const int n = ptr[j + 1] - ptr[j];
            for (int i = 0; i < n; ++i) okf[ptr[j] + i] = (n <= K) ? i : (int32_t)(((long long)i * K) / n);  // non-decreasing
        }
        kba_window w{};
        w.n_kf = K; w.n_cam = n_cam; w.n_lm = L; w.n_obs = O; w.n_gp = G;
        w.kf_pose = pose.data(); w.kf_fixed = fixed.data(); w.kf_plane = plane.data(); w.cam_intr = cam_intr; w.cam_pose = cam_pose;
        w.lm_pos = lmp.data(); w.lm_weight = lmw.data(); w.lm_obs_ptr = ptr.data(); w.obs_kf = okf.data(); w.obs_u = u.data();
        w.obs_v = v.data(); w.obs_d = d.data(); w.gp_lm = gl.data(); w.gp_kf = gk.data(); w.gp_weight = gw.data();
        w.plane_reg_weight = G > 0 ? 10.0 : 0.0;
        const int rc = kba_batch_create(h, 1, &w, &t->batch);
        if (rc != KBA_OK) { delete t; return rc; }
        if (!t->batch->device_pack) { kba_track_destroy(t); return fail(KBA_ERR_CAPACITY, "kba_track_create: device packing is disabled (KBA_FUSED / KBA_DEVICE_PACK)"); }
    }
    int bad = 0;
    TrackDev& td = t->td;
    td.kf_cap = c->max_keyframes; td.lm_cap = c->max_landmarks; td.m_cap = c->max_measurements;
    bad |= t->alloc(&td.kf_pose, 7 * (size_t)td.kf_cap); bad |= t->alloc(&td.kf_plane, 4 * (size_t)td.kf_cap);
    bad |= t->alloc(&td.m_off, td.kf_cap); bad |= t->alloc(&td.m_cnt, td.kf_cap);
    for (int b2 = 0; b2 < 2; ++b2) {
        for (int q = 0; q < 2; ++q) bad |= t->alloc(&t->arena_i[b2][q], td.m_cap);
        for (int q = 0; q < 3; ++q) bad |= t->alloc(&t->arena_f[b2][q], td.m_cap);
    }
    bad |= t->alloc(&td.lm_pos, 3 * (size_t)td.lm_cap); bad |= t->alloc(&td.lm_weight, td.lm_cap); bad |= t->alloc(&td.sel_index, td.lm_cap);
    bad |= t->alloc(&td.cursor, c->win_landmarks); bad |= t->alloc(&td.key, c->win_observations); bad |= t->alloc(&td.n_depth, 1);
    t->push_cap = std::min(c->max_measurements, 1 << 16);
    bad |= t->p_lm.alloc(t->push_cap, true); bad |= t->p_cam.alloc(t->push_cap, true); bad |= t->p_u.alloc(t->push_cap, true);
    bad |= t->p_v.alloc(t->push_cap, true); bad |= t->p_d.alloc(t->push_cap, true);
    bad |= t->sel_kf.alloc(c->win_keyframes, true); bad |= t->sel_fixed.alloc(c->win_keyframes, true); bad |= t->sel_lm.alloc(c->win_landmarks, true);
    bad |= t->lay.alloc(2 * (size_t)td.kf_cap, true);
    t->set_cap = std::max(c->win_landmarks, 64);  // rows per staged scatter (landmarks or keyframes)
    bad |= t->p_dbl.alloc(7 * (size_t)t->set_cap, true); bad |= t->p_slot.alloc(t->set_cap, true);
    if (bad) { kba_track_destroy(t); return fail(KBA_ERR_CUDA, "kba_track_create: out of memory"); }
    t->point_arena();
    t->m_off.assign(td.kf_cap, 0); t->m_cnt.assign(td.kf_cap, 0); t->kf_live.assign(td.kf_cap, 0);
    cudaStream_t s = h->stream;
    CU(cudaMemsetAsync(td.sel_index, 0xff, sizeof(int) * (size_t)td.lm_cap, s));
    CU(cudaMemsetAsync(td.m_cnt, 0, sizeof(int) * (size_t)td.kf_cap, s));
    CU(cudaMemsetAsync(td.lm_weight, 0, sizeof(double) * (size_t)td.lm_cap, s));
    CU(cudaStreamSynchronize(s));
    *out = t;
    return KBA_OK;
}

static int track_upload_layout(kba_track* t) {  // arena offsets / counts of every keyframe slot
    const int K = t->td.kf_cap;
    memcpy(t->lay.h, t->m_off.data(), K * sizeof(int));
    memcpy(t->lay.h + K, t->m_cnt.data(), K * sizeof(int));
    cudaStream_t s = t->h->stream;
    CU(cudaMemcpyAsync(t->td.m_off, t->lay.h, K * sizeof(int), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(t->td.m_cnt, t->lay.h + K, K * sizeof(int), cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));  // the pinned layout buffer is reused by the next call
    t->h2d_push += 2 * K * (int64_t)sizeof(int);
    return KBA_OK;
}

static int track_compact(kba_track* t) {  // live keyframes copied, in slot order, into the other arena
    const int other = 1 - t->arena_cur;
    cudaStream_t s = t->h->stream;
    int used = 0;
    for (int k = 0; k < t->td.kf_cap; ++k) {
        if (!t->kf_live[k] || t->m_cnt[k] == 0) { if (!t->kf_live[k]) t->m_cnt[k] = 0; continue; }
        const size_t n = (size_t)t->m_cnt[k], o = (size_t)t->m_off[k];
        for (int q = 0; q < 2; ++q) CU(cudaMemcpyAsync(t->arena_i[other][q] + used, t->arena_i[t->arena_cur][q] + o, n * sizeof(int), cudaMemcpyDeviceToDevice, s));
        for (int q = 0; q < 3; ++q) CU(cudaMemcpyAsync(t->arena_f[other][q] + used, t->arena_f[t->arena_cur][q] + o, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
        t->m_off[k] = used;
        used += (int)n;
    }
    t->arena_cur = other; t->arena_used = used;
    t->point_arena();
    return KBA_OK;
}

int kba_track_push_keyframe(kba_track* t, int32_t slot, const double* pose7, const double* plane4, int32_t n, const int32_t* lm,
                            const int32_t* cam, const float* u, const float* v, const float* d) {
    if (!t || !pose7 || n < 0 || slot < 0 || slot >= t->td.kf_cap || (n > 0 && (!lm || !u || !v || !d)))
        return fail(KBA_ERR_BAD_ARG, "bad argument to kba_track_push_keyframe");
    if (t->kf_live[slot]) return fail(KBA_ERR_BAD_ARG, "kba_track_push_keyframe: slot in use (drop it first)");
    for (int i = 0; i < n; ++i)
        if (lm[i] < 0 || lm[i] >= t->td.lm_cap || (cam && (cam[i] < 0 || cam[i] >= t->n_cam))) return fail(KBA_ERR_BAD_ARG, "kba_track_push_keyframe: landmark slot / camera out of range");
    CU(cudaSetDevice(t->h->device));
    if (t->arena_used + n > t->td.m_cap) {
        const int rc = track_compact(t);
        if (rc != KBA_OK) return rc;
        if (t->arena_used + n > t->td.m_cap) return fail(KBA_ERR_CAPACITY, "kba_track_push_keyframe: measurement arena full");
    }
    cudaStream_t s = t->h->stream;
    for (int i0 = 0; i0 < n; i0 += t->push_cap) {  // staged through pinned memory in chunks
        const int m = std::min(t->push_cap, n - i0);
        memcpy(t->p_lm.h, lm + i0, m * sizeof(int));
        if (cam) memcpy(t->p_cam.h, cam + i0, m * sizeof(int)); else memset(t->p_cam.h, 0, m * sizeof(int));
        memcpy(t->p_u.h, u + i0, m * sizeof(float)); memcpy(t->p_v.h, v + i0, m * sizeof(float)); memcpy(t->p_d.h, d + i0, m * sizeof(float));
        const size_t o = (size_t)t->arena_used + i0;
        CU(cudaMemcpyAsync(t->td.m_lm + o, t->p_lm.h, m * sizeof(int), cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(t->td.m_cam + o, t->p_cam.h, m * sizeof(int), cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(t->td.m_u + o, t->p_u.h, m * sizeof(float), cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(t->td.m_v + o, t->p_v.h, m * sizeof(float), cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(t->td.m_d + o, t->p_d.h, m * sizeof(float), cudaMemcpyHostToDevice, s));
        CU(cudaStreamSynchronize(s));
    }
    t->m_off[slot] = t->arena_used; t->m_cnt[slot] = n; t->kf_live[slot] = 1;
    t->arena_used += n;
    t->h2d_push += (int64_t)n * 20 + 11 * 8;
    const int rc = track_upload_layout(t);
    if (rc != KBA_OK) return rc;
    return kba_track_set_keyframe_pose(t, slot, pose7, plane4);
}

int kba_track_drop_keyframe(kba_track* t, int32_t slot) {
    if (!t || slot < 0 || slot >= t->td.kf_cap) return fail(KBA_ERR_BAD_ARG, "bad argument to kba_track_drop_keyframe");
    t->kf_live[slot] = 0;  // the arena space is reclaimed by the next compaction
    return KBA_OK;
}

// rows of `width` doubles into their store slots: staged through pinned memory, one copy + one scatter kernel per chunk
static int track_scatter(kba_track* t, double* dst, int cap_slots, int n, const int32_t* slot, const double* src, int width) {
    cudaStream_t s = t->h->stream;
    for (int i0 = 0; i0 < n; i0 += t->set_cap) {
        const int m = std::min(t->set_cap, n - i0);
        for (int i = 0; i < m; ++i)
            if (slot[i0 + i] < 0 || slot[i0 + i] >= cap_slots) return fail(KBA_ERR_BAD_ARG, "kba_track: slot out of range");
        memcpy(t->p_slot.h, slot + i0, m * sizeof(int));
        memcpy(t->p_dbl.h, src + (size_t)width * i0, (size_t)width * m * sizeof(double));
        CU(cudaMemcpyAsync(t->p_slot.d, t->p_slot.h, m * sizeof(int), cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(t->p_dbl.d, t->p_dbl.h, (size_t)width * m * sizeof(double), cudaMemcpyHostToDevice, s));
        launch_scatter_rows(dst, t->p_slot.d, t->p_dbl.d, m, width, s);
        CU(cudaStreamSynchronize(s));  // the staging buffers are reused
    }
    return KBA_OK;
}

int kba_track_set_keyframe_poses(kba_track* t, int32_t n, const int32_t* slot, const double* pose7s, const double* plane4s) {
    if (!t || n < 0 || (n > 0 && (!slot || !pose7s))) return fail(KBA_ERR_BAD_ARG, "bad argument to kba_track_set_keyframe_poses");
    CU(cudaSetDevice(t->h->device));
    int rc = track_scatter(t, t->td.kf_pose, t->td.kf_cap, n, slot, pose7s, 7);
    if (rc == KBA_OK && plane4s) rc = track_scatter(t, t->td.kf_plane, t->td.kf_cap, n, slot, plane4s, 4);
    return rc;
}

int kba_track_set_keyframe_pose(kba_track* t, int32_t slot, const double* pose7, const double* plane4) {
    static const double kNoPlane[4] = {0., 0., 1., 0.};
    return kba_track_set_keyframe_poses(t, 1, &slot, pose7, plane4 ? plane4 : kNoPlane);
}

int kba_track_set_landmarks(kba_track* t, int32_t n, const int32_t* slot, const double* pos3, const double* weight) {
    if (!t || n < 0 || (n > 0 && !slot)) return fail(KBA_ERR_BAD_ARG, "bad argument to kba_track_set_landmarks");
    CU(cudaSetDevice(t->h->device));
    int rc = KBA_OK;
    if (pos3) rc = track_scatter(t, t->td.lm_pos, t->td.lm_cap, n, slot, pos3, 3);
    if (rc == KBA_OK && weight) rc = track_scatter(t, t->td.lm_weight, t->td.lm_cap, n, slot, weight, 1);
    t->h2d_push += (int64_t)n * ((pos3 ? 24 : 0) + (weight ? 8 : 0) + 4);
    return rc;
}

int kba_track_solve(kba_track* t, int32_t n_kf, const int32_t* kf_slot, const uint8_t* kf_fixed, int32_t n_lm, const int32_t* lm_slot,
                    const kba_window* sel, const kba_options* opt, kba_result* res) {
    if (!t || !kf_slot || !kf_fixed || !lm_slot || !sel || !opt || !res) return fail(KBA_ERR_BAD_ARG, "null argument to kba_track_solve");
    const kba_track_caps& c = t->caps;
    if (n_kf < 3) return fail(KBA_ERR_NOT_ENOUGH_KF, "kba_track_solve: fewer than 3 keyframes");
    if (n_kf > c.win_keyframes || n_lm > c.win_landmarks || n_lm < 0 || sel->n_gp < 0 || sel->n_gp > c.win_ground)
        return fail(KBA_ERR_CAPACITY, "kba_track_solve: window larger than the capacities given to kba_track_create");
    long long n_meas = 0;
    int max_meas = 0, n_free = 0;
    for (int k = 0; k < n_kf; ++k) {
        if (kf_slot[k] < 0 || kf_slot[k] >= t->td.kf_cap || !t->kf_live[kf_slot[k]]) return fail(KBA_ERR_BAD_ARG, "kba_track_solve: keyframe slot not pushed");
        n_meas += t->m_cnt[kf_slot[k]]; max_meas = std::max(max_meas, t->m_cnt[kf_slot[k]]);
        n_free += kf_fixed[k] ? 0 : 1;
    }
    if (n_meas > c.win_observations) return fail(KBA_ERR_CAPACITY, "kba_track_solve: more observations than win_observations");
    for (int j = 0; j < n_lm; ++j)
        if (lm_slot[j] < 0 || lm_slot[j] >= t->td.lm_cap) return fail(KBA_ERR_BAD_ARG, "kba_track_solve: landmark slot out of range");
    for (int g = 0; g < sel->n_gp; ++g)
        if (!sel->gp_lm || !sel->gp_kf || !sel->gp_weight || sel->gp_lm[g] < 0 || sel->gp_lm[g] >= n_lm || sel->gp_kf[g] < 0 || sel->gp_kf[g] >= n_kf)
            return fail(KBA_ERR_BAD_ARG, "kba_track_solve: ground-plane index out of range");
    const bool planes = sel->n_gp > 0 || sel->plane_reg_weight > 0;
    if (planes && c.win_ground == 0) return fail(KBA_ERR_CAPACITY, "kba_track_solve: the track was created without ground-plane capacity");
    kba_batch* b = t->batch;
    CU(cudaSetDevice(t->h->device));
    cudaStream_t s = t->h->stream;
    // ---- the window descriptor of this solve (n_obs is written by the gather kernels)
    WinDesc& d = b->desc_h[0];
    d.n_kf = n_kf; d.n_lm = n_lm; d.n_obs = 0; d.n_gp = sel->n_gp;
    d.n_chunks = (n_lm + 31) / 32; d.n_groups = (n_lm + 7) / 8;
    d.scale_kf0 = sel->scale_kf0; d.scale_kf1 = sel->scale_kf1; d.scale_weight = sel->scale_weight; d.scale_value = sel->scale_value;
    d.plane_reg_weight = sel->plane_reg_weight; d.plane_dist_fixed = sel->plane_dist_fixed; d.landmarks_fixed = 0;
    d.speed_kf = 0; d.speed_weight = 0; d.speed_dt = 1;
    d.max_rank = t->n_cam > 1 ? t->n_cam - 1 : 0;  // a rig may see a landmark from several cameras of one keyframe
    if (d.scale_weight != 0 && (d.scale_kf0 < 0 || d.scale_kf0 >= n_kf || d.scale_kf1 < 0 || d.scale_kf1 >= n_kf))
        return fail(KBA_ERR_BAD_ARG, "kba_track_solve: scale regulariser keyframe out of range");
    b->desc.h[0] = d;
    b->lc.max_rank = d.max_rank;
    b->lc.fused_slots = ((planes ? 10 : 6) * n_free + 1 <= 176) ? 6 : 7;
    memcpy(t->sel_kf.h, kf_slot, n_kf * sizeof(int)); memcpy(t->sel_fixed.h, kf_fixed, n_kf);
    memcpy(t->sel_lm.h, lm_slot, n_lm * sizeof(int));
    if (sel->n_gp) {
        memcpy(b->r_gp_lm.h, sel->gp_lm, sel->n_gp * sizeof(int)); memcpy(b->gp_kf.h, sel->gp_kf, sel->n_gp * sizeof(int));
        memcpy(b->gp_weight.h, sel->gp_weight, sel->n_gp * sizeof(double));
    }
    CU(b->desc.upload(s));
    CU(cudaMemcpyAsync(t->sel_kf.d, t->sel_kf.h, n_kf * sizeof(int), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(t->sel_fixed.d, t->sel_fixed.h, n_kf, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(t->sel_lm.d, t->sel_lm.h, n_lm * sizeof(int), cudaMemcpyHostToDevice, s));
    if (sel->n_gp) { CU(b->r_gp_lm.upload(s)); CU(b->gp_kf.upload(s)); CU(b->gp_weight.upload(s)); }
    t->h2d_solve = (int64_t)sizeof(WinDesc) + n_kf * 5 + n_lm * 4 + sel->n_gp * 16;
    // ---- gather the CSR from the store, pack, solve, write back
    TrackSel ts;
    ts.kf_slot = t->sel_kf.d; ts.kf_fixed = t->sel_fixed.d; ts.lm_slot = t->sel_lm.d; ts.n_kf = n_kf; ts.n_lm = n_lm; ts.max_meas = max_meas;
    ts.auto_scale = sel->scale_weight < 0 ? 1 : 0;
    launch_track_gather(b->bd, b->raw, t->td, ts, s);
    launch_pack(b->bd, b->raw, s);
    CU(cudaGetLastError());
    int rc = kba_batch_solve(b, opt);
    if (rc != KBA_OK) return rc;
    launch_track_writeback(b->bd, t->td, ts, s);
    rc = kba_batch_download(b, res);
    t->d2h_solve = (int64_t)b->d2h_bytes;
    return rc;
}

int kba_track_transfer_bytes(kba_track* t, int64_t* h2d, int64_t* d2h, int64_t* push) {
    if (!t) return fail(KBA_ERR_BAD_ARG, "null track");
    if (h2d) *h2d = t->h2d_solve;
    if (d2h) *d2h = t->d2h_solve;
    if (push) *push = t->h2d_push;
    return KBA_OK;
}

}  // extern "C"

