// Smoke test of the CUDA-graph WHILE node + cudaGraphSetConditional on this driver (the fit loop relies on it):
// the body runs until the kernel inside it clears the condition; prints the trip count (expect 5).
#include <cuda_runtime.h>
#include <cstdio>
__global__ void setc(cudaGraphConditionalHandle h, int* ctr) {
    int v = ++(*ctr);
    cudaGraphSetConditional(h, v < 5 ? 1u : 0u);
}
int main() {
    cudaStream_t s; cudaStreamCreate(&s);
    int* ctr; cudaMalloc(&ctr, 4); cudaMemset(ctr, 0, 4);
    cudaGraph_t g; cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    cudaStreamCaptureStatus st; const cudaGraphNode_t* deps; size_t nd; cudaGraph_t cg;
    cudaStreamGetCaptureInfo_v2(s, &st, nullptr, &cg, &deps, &nd);
    cudaGraphConditionalHandle h; cudaGraphConditionalHandleCreate(&h, cg, 1, cudaGraphCondAssignDefault);
    cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = h;
    p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
    cudaGraphNode_t node; cudaGraphAddNode(&node, cg, deps, nd, &p);
    cudaGraph_t body = p.conditional.phGraph_out[0];
    cudaStream_t s2; cudaStreamCreate(&s2);
    cudaStreamBeginCaptureToGraph(s2, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    setc<<<1,1,0,s2>>>(h, ctr);
    cudaStreamEndCapture(s2, nullptr);
    cudaStreamUpdateCaptureDependencies(s, &node, 1, cudaStreamSetCaptureDependencies);
    cudaStreamEndCapture(s, &g);
    cudaGraphExec_t e; cudaGraphInstantiate(&e, g, 0);
    cudaGraphLaunch(e, s); cudaStreamSynchronize(s);
    int hc; cudaMemcpy(&hc, ctr, 4, cudaMemcpyDeviceToHost); printf("%d %s\n", hc, cudaGetErrorString(cudaGetLastError()));
}
