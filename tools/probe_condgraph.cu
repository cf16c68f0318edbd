#include <cuda_runtime.h>
#include <cstdio>
__global__ void body(int* counter, cudaGraphConditionalHandle h) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int c = atomicAdd(counter, 1) + 1;
    cudaGraphSetConditional(h, c < 10 ? 1 : 0);
  }
}
int main() {
  int* d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
  cudaGraph_t g; cudaGraphCreate(&g, 0);
  cudaGraphConditionalHandle h;
  cudaError_t e = cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
  printf("handle: %s\n", cudaGetErrorString(e));
  cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t node; e = cudaGraphAddNode(&node, g, nullptr, 0, &p); printf("addnode: %s\n", cudaGetErrorString(e));
  cudaGraph_t bodyg = p.conditional.phGraph_out[0];
  cudaStream_t s; cudaStreamCreate(&s);
  cudaStreamBeginCaptureToGraph(s, bodyg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
  body<<<4, 32, 0, s>>>(d, h);
  cudaStreamEndCapture(s, nullptr);
  cudaGraphExec_t ex; e = cudaGraphInstantiate(&ex, g, 0); printf("inst: %s\n", cudaGetErrorString(e));
  cudaGraphLaunch(ex, s); cudaStreamSynchronize(s);
  int hc; cudaMemcpy(&hc, d, 4, cudaMemcpyDeviceToHost); printf("count=%d\n", hc);
}
