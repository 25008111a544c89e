// kba_kernels.h -- host-visible launch interface of kba_kernels.cu
#pragma once
#include <cuda_runtime.h>

#include "kba_device.cuh"

namespace kba {

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
};

struct LaunchCfg {
    int nr_cap_max = 64;
    int max_rank = 0;         // largest observation rank in the batch (multi-camera rigs)
    bool small_syrk = false;  // every window has <= 184 reduced rows: register-resident Schur kernel
    int fused_slots = 6;      // fused Schur kernel instance: 6 accumulator blocks per warp (<= 176 rows) or 7 (<= 184)
    int rounds_override = -1, min_landmarks_for_trimming = 100, num_rounds_option = 1;
    bool time_jacobian = false;
    cudaEvent_t* ev_pool = nullptr;  // pairs of events bracketing each residual/Jacobian launch
    int ev_cap = 0;
    int* ev_used = nullptr;
    Exchange xchg;  // used when BatchDev::sharded
    struct WinDescHost { int nr_cap = 0, n_kf = 0; } shard_win;  // shapes of the sharded window (host copy)
    double *x_sred = nullptr, *x_bkf = nullptr, *x_cost = nullptr;  // window-wide sums (receive buffers of the exchange)
};

cudaError_t configure_kernels(int nr_cap_max);
void launch_reset(const BatchDev& bd, const LaunchCfg& lc, cudaStream_t s);
int launch_pass(const BatchDev& bd, const SolveParams& sp, const LaunchCfg& lc, Counters* cnt, cudaStream_t s);
void launch_count_active(const BatchDev& bd, cudaStream_t s);
void launch_force_linearize(const BatchDev& bd, cudaStream_t s);
void launch_jacobian_only(const BatchDev& bd, const SolveParams& sp, cudaStream_t s);
void launch_expand_jl(const BatchDev& bd, double* out, cudaStream_t s);  // fused path: J_l as its consumers form it

}  // namespace kba
