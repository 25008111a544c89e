// kba_kernels.h -- host-visible launch interface of kba_kernels.cu
#pragma once
#include <cuda_runtime.h>

#include "kba_device.cuh"

namespace kba {

// KBA_LAUNCH_CHECK=1 (debugging): every kernel launch is followed by cudaGetLastError() and a failing one is named on stderr.
// Off by default: kba_batch_solve checks once per solve.
int launch_check_enabled();
void launch_check_report(const char* kernel, cudaError_t e);
#define LCHK(name)                                                                       \
    do {                                                                                 \
        if (kba::launch_check_enabled()) {                                               \
            const cudaError_t lchk_e_ = cudaGetLastError();                              \
            if (lchk_e_ != cudaSuccess) kba::launch_check_report(name, lchk_e_);         \
        }                                                                                \
    } while (0)

struct Counters {
    long long launches_total = 0;
    long long launches_jacobian = 0, launches_prep = 0, launches_schur = 0, launches_solve = 0, launches_backsub = 0,
              launches_cost = 0, launches_update = 0, launches_trim = 0;
    double ms_jacobian = 0.0;
    long long jacobian_obs = 0;
};

// cross-rank reduction hook of the sharded solve: out-of-place all-reduce of `count` doubles on stream s
// (op 0 = sum, 1 = max); returns 0 on success.  Implemented over NCCL in kba_shard.cu.
struct Exchange {
    int (*allreduce)(void* user, const double* send, double* recv, long long count, int op, cudaStream_t s) = nullptr;
    void* user = nullptr;
    int rank = 0, world = 1;
};

struct LaunchCfg {
    int nr_cap_max = 64;
    int max_rank = 0;         // largest observation rank in the batch (multi-camera rigs)
    bool small_syrk = false;  // every window has <= 184 reduced rows: register-resident Schur kernel
    bool lin_fused = true;    // fused path: evaluation + landmark blocks + V rows in one kernel (k_linearize); KBA_LINEARIZE=0: three kernels
    int lin_blocks = 2;       // CTAs per SM k_linearize is compiled for (KBA_LIN_BLOCKS: 2 or 3)
    int lin_grid = -1;        // CTAs per window of k_linearize striding over its tile units (KBA_LIN_GRID; -1: by batch size, 0: one CTA per unit)
    int bs_grid = -1;         // the same for k_backsub_v (KBA_BS_GRID; 0: one CTA per 16 landmarks)
    int fused_slots = 6;      // fused Schur kernel instance: 6 accumulator blocks per warp (<= 176 rows) or 7 (<= 184)
    int rounds_override = -1, min_landmarks_for_trimming = 100, num_rounds_option = 1;
    bool time_jacobian = false;
    cudaEvent_t* ev_pool = nullptr;  // pairs of events bracketing each residual/Jacobian launch
    int ev_cap = 0;
    int* ev_used = nullptr;
    Exchange xchg;  // used when BatchDev::sharded
    struct WinDescHost { int nr_cap = 0, n_kf = 0; } shard_win;  // shapes of the sharded window (host copy)
};

// window arrays exactly as the caller passes them (landmark-major CSR in the caller's landmark order), batch-flat on the
// device: input of the packing kernels (kba_pack.cu)
struct PackRaw {
    const int* lm_ptr = nullptr;      // [tot_lm + n_win]
    const int* obs_kf = nullptr;      // [tot_obs]
    const int* obs_cam = nullptr;
    const float* obs_u = nullptr, *obs_v = nullptr, *obs_d = nullptr;
    const double* lm_pos = nullptr;   // [tot_lm*3]
    const double* lm_weight = nullptr;
    const int* gp_lm = nullptr;       // [tot_gp] caller landmark index
    int* lm_inv = nullptr;            // [tot_lm] scratch: caller index -> sorted position
    int* obs_orig = nullptr;          // [tot_obs] optional: sorted slot -> caller observation index
};
// device-resident track store of a persistent sliding window (kba_track_*, include/kba_b200.h): keyframe poses / planes,
// the measurements of every pushed keyframe in one arena, landmark positions / weights by caller-assigned slot
struct TrackDev {
    double* kf_pose = nullptr;   // [kf_cap*7]
    double* kf_plane = nullptr;  // [kf_cap*4]
    int* m_off = nullptr;        // [kf_cap] first measurement of the keyframe in the arena
    int* m_cnt = nullptr;        // [kf_cap]
    int* m_lm = nullptr;         // [m_cap] landmark slot
    int* m_cam = nullptr;        // [m_cap] camera index
    float* m_u = nullptr, *m_v = nullptr, *m_d = nullptr;
    double* lm_pos = nullptr;    // [lm_cap*3]
    double* lm_weight = nullptr; // [lm_cap]
    int* sel_index = nullptr;    // [lm_cap] position of the landmark in the current selection, -1 otherwise (all -1 between solves)
    int* cursor = nullptr;       // [lm window capacity] scratch
    long long* key = nullptr;    // [obs window capacity] scratch: (keyframe index, arena index) of each gathered observation
    int* n_depth = nullptr;      // [1] gathered observations with a lidar depth (d > 0): the depth residual blocks of the window
    int kf_cap = 0, lm_cap = 0, m_cap = 0;
};
// per-solve selection, device copies of the caller's small lists
struct TrackSel {
    const int* kf_slot = nullptr;    // [n_kf] ascending keyframe id
    const unsigned char* kf_fixed = nullptr;
    const int* lm_slot = nullptr;    // [n_lm] ascending landmark id
    int n_kf = 0, n_lm = 0, max_meas = 0;
    int auto_scale = 0;              // 1: scale-regulariser weight by the reference rule (cpp:703-716) from the gathered window
};
// builds the window's raw CSR (PackRaw inputs of batch `bd`, window 0) from the track store; returns nothing: desc[0].n_obs
// is written on the device
void launch_track_gather(const BatchDev& bd, const PackRaw& raw_out, const TrackDev& td, const TrackSel& sel, cudaStream_t s);
void launch_scatter_rows(double* dst, const int* slot, const double* src, int n, int width, cudaStream_t s);
// results of window 0 back into the track store
void launch_track_writeback(const BatchDev& bd, const TrackDev& td, const TrackSel& sel, cudaStream_t s);

cudaError_t configure_pack();
int pack_max_landmarks();
void launch_pack(const BatchDev& bd, const PackRaw& raw, cudaStream_t s);
void launch_unpack_landmarks(const BatchDev& bd, double* lm_user, unsigned char* rejected_user, cudaStream_t s);

cudaError_t configure_kernels(int nr_cap_max);
void launch_reset(const BatchDev& bd, const LaunchCfg& lc, cudaStream_t s);
int launch_pass(const BatchDev& bd, const SolveParams& sp, const LaunchCfg& lc, Counters* cnt, cudaStream_t s);
void launch_count_active(const BatchDev& bd, cudaStream_t s);
void launch_loop_cond(const BatchDev& bd, unsigned long long handle, int* pass, int max_passes, cudaStream_t s);  // WHILE-node condition
void launch_force_linearize(const BatchDev& bd, cudaStream_t s);
void launch_jacobian_only(const BatchDev& bd, const SolveParams& sp, cudaStream_t s);
void launch_expand_jl(const BatchDev& bd, double* out, cudaStream_t s);  // fused path: J_l as its consumers form it

}  // namespace kba
