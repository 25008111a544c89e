// kba_init.cu -- landmark initialisation of BundleAdjusterKeyframes::push() for a whole window on the device
// (SURVEY 8(f) row 2): depth back-projection (reference bundle_adjuster_keyframes.cpp:332-355), least-squares ray
// intersection (cpp:125-159,358-382 + internal/triangulator.hpp:51-75) and the cheirality test of the landmark
// selector (landmark_selection_scheme_cheirality.cpp:22-60).  One thread per landmark walks its CSR row; the tiny
// per-keyframe / per-camera transforms are staged in shared memory.
#include <algorithm>
#include <vector>

#include <cuda_runtime.h>

#include "kba_b200.h"
#include "kba_device.cuh"

extern "C" {
int kba_internal_stream(kba_handle* h, cudaStream_t* s, int* device);
int kba_internal_fail(int code, const char* msg);
}

namespace kba {

// transforms of one (keyframe, camera) pair, origin <- camera: p_o = Roc p_cam + toc
__device__ inline void origin_from_camera(const double* kf /*R|t staged*/, const double* cam /*Rc|tc*/, double Roc[9], double toc[3]) {
    // p_cam = Rc (R p_o + t) + tc  ->  p_o = R^T (Rc^T (p_cam - tc) - t)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Roc[3 * i + j] = kf[i] * cam[3 * j] + kf[3 + i] * cam[3 * j + 1] + kf[6 + i] * cam[3 * j + 2];  // (R^T Rc^T)_ij = sum_k R_ki Rc_jk
    double a[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) a[i] = cam[i] * cam[9] + cam[3 + i] * cam[10] + cam[6 + i] * cam[11] + kf[9 + i];  // Rc^T tc + t
#pragma unroll
    for (int i = 0; i < 3; ++i) toc[i] = -(kf[i] * a[0] + kf[3 + i] * a[1] + kf[6 + i] * a[2]);
}

__global__ void __launch_bounds__(128) k_init_landmarks(int n_kf, int n_cam, int n_lm, const double* __restrict__ kf_pose,
                                                        const double* __restrict__ cam16, const int* __restrict__ lm_ptr,
                                                        const int* __restrict__ obs_kf, const int* __restrict__ obs_cam,
                                                        const float* __restrict__ obs_u, const float* __restrict__ obs_v,
                                                        const float* __restrict__ obs_d, double* __restrict__ lm_out,
                                                        unsigned char* __restrict__ flags) {
    __shared__ double s_pose[kMaxKf * kPoseStride];
    __shared__ double s_cam[kMaxCam * kCamStride];
    for (int k = threadIdx.x; k < n_kf; k += blockDim.x) {
        double R[9];
        quat_to_rot<double>(kf_pose + 7 * k, R);
        for (int i = 0; i < 9; ++i) s_pose[kPoseStride * k + i] = R[i];
        s_pose[kPoseStride * k + 9] = kf_pose[7 * k + 4];
        s_pose[kPoseStride * k + 10] = kf_pose[7 * k + 5];
        s_pose[kPoseStride * k + 11] = kf_pose[7 * k + 6];
    }
    for (int i = threadIdx.x; i < n_cam * kCamStride; i += blockDim.x) s_cam[i] = cam16[i];
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_lm) return;
    const int o0 = lm_ptr[j], o1 = lm_ptr[j + 1];
    double p[3] = {0.0, 0.0, 0.0};
    unsigned char created = 0;
    // (1) the first observation that carries a lidar depth: back-projection (cpp:332-355)
    for (int o = o0; o < o1 && !created; ++o) {
        const float d = obs_d[o];
        if (d < 0.f) continue;
        const double* kf = s_pose + kPoseStride * obs_kf[o];
        const double* cam = s_cam + kCamStride * (obs_cam ? obs_cam[o] : 0);
        const double z = (double)d;
        const double pc[3] = {((double)obs_u[o] - cam[13]) * z / cam[12], ((double)obs_v[o] - cam[14]) * z / cam[12], z};
        double Roc[9], toc[3];
        origin_from_camera(kf, cam, Roc, toc);
#pragma unroll
        for (int i = 0; i < 3; ++i) p[i] = Roc[3 * i] * pc[0] + Roc[3 * i + 1] * pc[1] + Roc[3 * i + 2] * pc[2] + toc[i];
        created = 1;
    }
    // (2) otherwise the point closest to all viewing rays, needs two of them (cpp:125-159, triangulator.hpp:51-75)
    if (!created && o1 - o0 >= 2) {
        double A[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};
        for (int o = o0; o < o1; ++o) {
            const double* kf = s_pose + kPoseStride * obs_kf[o];
            const double* cam = s_cam + kCamStride * (obs_cam ? obs_cam[o] : 0);
            double ray[3] = {((double)obs_u[o] - cam[13]) / cam[12], ((double)obs_v[o] - cam[14]) / cam[12], 1.0};
            const double nrm = sqrt(ray[0] * ray[0] + ray[1] * ray[1] + ray[2] * ray[2]);
            ray[0] /= nrm; ray[1] /= nrm; ray[2] /= nrm;
            double Roc[9], toc[3], r[3];
            origin_from_camera(kf, cam, Roc, toc);
#pragma unroll
            for (int i = 0; i < 3; ++i) r[i] = Roc[3 * i] * ray[0] + Roc[3 * i + 1] * ray[1] + Roc[3 * i + 2] * ray[2];
            double M[9];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int c = 0; c < 3; ++c) M[3 * a + c] = (a == c ? 1.0 : 0.0) - r[a] * r[c];
#pragma unroll
            for (int q = 0; q < 9; ++q) A[q] += M[q];
#pragma unroll
            for (int a = 0; a < 3; ++a) b[a] += M[3 * a] * toc[0] + M[3 * a + 1] * toc[1] + M[3 * a + 2] * toc[2];
        }
        const double det = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
        double inv[9];
        inv[0] = (A[4] * A[8] - A[5] * A[7]) / det; inv[1] = (A[2] * A[7] - A[1] * A[8]) / det; inv[2] = (A[1] * A[5] - A[2] * A[4]) / det;
        inv[3] = (A[5] * A[6] - A[3] * A[8]) / det; inv[4] = (A[0] * A[8] - A[2] * A[6]) / det; inv[5] = (A[2] * A[3] - A[0] * A[5]) / det;
        inv[6] = (A[3] * A[7] - A[4] * A[6]) / det; inv[7] = (A[1] * A[6] - A[0] * A[7]) / det; inv[8] = (A[0] * A[4] - A[1] * A[3]) / det;
#pragma unroll
        for (int i = 0; i < 3; ++i) p[i] = inv[3 * i] * b[0] + inv[3 * i + 1] * b[1] + inv[3 * i + 2] * b[2];
        created = 1;
    }
    // (3) cheirality: in front of every observing camera (landmark_selection_scheme_cheirality.cpp:22-60)
    unsigned char front = created;
    for (int o = o0; o < o1 && front; ++o) {
        const double* kf = s_pose + kPoseStride * obs_kf[o];
        const double* cam = s_cam + kCamStride * (obs_cam ? obs_cam[o] : 0);
        double x[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) x[i] = kf[3 * i] * p[0] + kf[3 * i + 1] * p[1] + kf[3 * i + 2] * p[2] + kf[9 + i];
        const double zc = cam[6] * x[0] + cam[7] * x[1] + cam[8] * x[2] + cam[11];
        if (zc < 0.0) front = 0;
    }
    lm_out[3 * j] = p[0]; lm_out[3 * j + 1] = p[1]; lm_out[3 * j + 2] = p[2];
    flags[j] = (unsigned char)(created | (front << 1));
}

}  // namespace kba

extern "C" int kba_init_landmarks(kba_handle* h, const kba_window* w, double* lm_pos_out, uint8_t* flags_out, float* device_ms) {
    if (!h || !w || !lm_pos_out || !flags_out || w->n_kf < 1 || w->n_kf > kba::kMaxKf || w->n_cam < 1 || w->n_cam > kba::kMaxCam ||
        w->n_lm < 0 || !w->kf_pose || !w->cam_pose || !w->cam_intr || (w->n_lm > 0 && !w->lm_obs_ptr))
        return kba_internal_fail(KBA_ERR_BAD_ARG, "bad argument to kba_init_landmarks");
    cudaStream_t s;
    int device;
    if (kba_internal_stream(h, &s, &device) != KBA_OK) return KBA_ERR_BAD_ARG;
    const int n_obs = w->n_lm > 0 ? w->lm_obs_ptr[w->n_lm] : 0;
    if (n_obs != w->n_obs) return kba_internal_fail(KBA_ERR_BAD_ARG, "kba_init_landmarks: lm_obs_ptr does not end at n_obs");
    std::vector<double> cam16((size_t)w->n_cam * kba::kCamStride);
    for (int c = 0; c < w->n_cam; ++c) {
        double* o = cam16.data() + kba::kCamStride * (size_t)c;
        kba::quat_to_rot<double>(w->cam_pose + 7 * c, o);
        o[9] = w->cam_pose[7 * c + 4]; o[10] = w->cam_pose[7 * c + 5]; o[11] = w->cam_pose[7 * c + 6];
        o[12] = w->cam_intr[3 * c]; o[13] = w->cam_intr[3 * c + 1]; o[14] = w->cam_intr[3 * c + 2]; o[15] = 0;
    }
    cudaError_t err = cudaSuccess;
    auto chk = [&](cudaError_t e) { if (err == cudaSuccess && e != cudaSuccess) err = e; };
    double *d_pose = nullptr, *d_cam = nullptr, *d_out = nullptr;
    int *d_ptr = nullptr, *d_kf = nullptr, *d_cami = nullptr;
    float *d_u = nullptr, *d_v = nullptr, *d_d = nullptr;
    unsigned char* d_flags = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const size_t no = (size_t)std::max(n_obs, 1), nl = (size_t)std::max(w->n_lm, 1);
    chk(cudaSetDevice(device));
    chk(cudaMalloc(&d_pose, sizeof(double) * 7 * w->n_kf)); chk(cudaMalloc(&d_cam, sizeof(double) * cam16.size()));
    chk(cudaMalloc(&d_out, sizeof(double) * 3 * nl)); chk(cudaMalloc(&d_flags, nl));
    chk(cudaMalloc(&d_ptr, sizeof(int) * (nl + 1))); chk(cudaMalloc(&d_kf, sizeof(int) * no));
    if (w->obs_cam) chk(cudaMalloc(&d_cami, sizeof(int) * no));
    chk(cudaMalloc(&d_u, sizeof(float) * no)); chk(cudaMalloc(&d_v, sizeof(float) * no)); chk(cudaMalloc(&d_d, sizeof(float) * no));
    chk(cudaEventCreate(&e0)); chk(cudaEventCreate(&e1));
    if (err == cudaSuccess && w->n_lm > 0) {
        chk(cudaMemcpyAsync(d_pose, w->kf_pose, sizeof(double) * 7 * w->n_kf, cudaMemcpyHostToDevice, s));
        chk(cudaMemcpyAsync(d_cam, cam16.data(), sizeof(double) * cam16.size(), cudaMemcpyHostToDevice, s));
        chk(cudaMemcpyAsync(d_ptr, w->lm_obs_ptr, sizeof(int) * (w->n_lm + 1), cudaMemcpyHostToDevice, s));
        chk(cudaMemcpyAsync(d_kf, w->obs_kf, sizeof(int) * n_obs, cudaMemcpyHostToDevice, s));
        if (w->obs_cam) chk(cudaMemcpyAsync(d_cami, w->obs_cam, sizeof(int) * n_obs, cudaMemcpyHostToDevice, s));
        chk(cudaMemcpyAsync(d_u, w->obs_u, sizeof(float) * n_obs, cudaMemcpyHostToDevice, s));
        chk(cudaMemcpyAsync(d_v, w->obs_v, sizeof(float) * n_obs, cudaMemcpyHostToDevice, s));
        chk(cudaMemcpyAsync(d_d, w->obs_d, sizeof(float) * n_obs, cudaMemcpyHostToDevice, s));
        chk(cudaEventRecord(e0, s));
        kba::k_init_landmarks<<<(w->n_lm + 127) / 128, 128, 0, s>>>(w->n_kf, w->n_cam, w->n_lm, d_pose, d_cam, d_ptr, d_kf, d_cami,
                                                                  d_u, d_v, d_d, d_out, d_flags);
        chk(cudaEventRecord(e1, s));
        chk(cudaMemcpyAsync(lm_pos_out, d_out, sizeof(double) * 3 * w->n_lm, cudaMemcpyDeviceToHost, s));
        chk(cudaMemcpyAsync(flags_out, d_flags, (size_t)w->n_lm, cudaMemcpyDeviceToHost, s));
        chk(cudaStreamSynchronize(s));
        chk(cudaGetLastError());
        if (err == cudaSuccess && device_ms) chk(cudaEventElapsedTime(device_ms, e0, e1));
    }
    cudaFree(d_pose); cudaFree(d_cam); cudaFree(d_out); cudaFree(d_flags); cudaFree(d_ptr); cudaFree(d_kf);
    if (d_cami) cudaFree(d_cami);
    cudaFree(d_u); cudaFree(d_v); cudaFree(d_d);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (err != cudaSuccess) return kba_internal_fail(KBA_ERR_CUDA, cudaGetErrorString(err));
    return KBA_OK;
}
