#include <cuda_runtime.h>
#include <cstdio>
__global__ void k_body(int* c) { if (threadIdx.x == 0) (*c)++; }
__global__ void k_cond(cudaGraphConditionalHandle h, int* c, int n) { if (threadIdx.x == 0) cudaGraphSetConditional(h, *c < n ? 1u : 0u); }
int main() {
    int* c; cudaMalloc(&c, 4); cudaMemset(c, 0, 4);
    cudaStream_t s; cudaStreamCreate(&s);
    cudaGraph_t g; cudaGraphCreate(&g, 0);
    cudaGraphConditionalHandle h;
    cudaError_t e = cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
    printf("handle %s\n", cudaGetErrorString(e));
    cudaGraphNodeParams p = {};
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
    cudaGraphNode_t node;
    e = cudaGraphAddNode(&node, g, nullptr, 0, &p);
    printf("addnode %s\n", cudaGetErrorString(e));
    cudaGraph_t body = p.conditional.phGraph_out[0];
    e = cudaStreamBeginCaptureToGraph(s, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    printf("begin %s\n", cudaGetErrorString(e));
    k_body<<<1, 32, 0, s>>>(c);
    k_cond<<<1, 32, 0, s>>>(h, c, 10);
    e = cudaStreamEndCapture(s, nullptr);
    printf("end %s\n", cudaGetErrorString(e));
    cudaGraphExec_t x; e = cudaGraphInstantiate(&x, g, 0);
    printf("inst %s\n", cudaGetErrorString(e));
    cudaGraphLaunch(x, s); cudaStreamSynchronize(s);
    int hc = 0; cudaMemcpy(&hc, c, 4, cudaMemcpyDeviceToHost);
    printf("count %d (want 10)\n", hc);
    cudaMemset(c, 0, 4); cudaGraphLaunch(x, s); cudaStreamSynchronize(s);
    cudaMemcpy(&hc, c, 4, cudaMemcpyDeviceToHost);
    printf("count %d (want 10)\n", hc);
    return 0;
}
