// Probe: does this driver / runtime run CUDA-graph conditional WHILE nodes whose condition a kernel sets
// (cudaGraphSetConditional)?  Expected output on the B200 box: "add 0", "inst 0", "ctr 5" (the body ran five times).
// rbv_slice_run's graph mode (rbvfit_b200/csrc/rbv_kernels.cu) is built the same way.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o build/cond_graph_probe tools/micro/cond_graph_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
__global__ void body(int* ctr, cudaGraphConditionalHandle h) {
  int v = atomicAdd(ctr, 1);
  if (v >= 4) cudaGraphSetConditional(h, 0);
}
int main() {
  cudaStream_t st; cudaStreamCreate(&st);
  int* d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
  cudaGraph_t g; cudaGraphCreate(&g, 0);
  cudaGraphConditionalHandle h;
  cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
  cudaGraphNodeParams p = {cudaGraphNodeTypeConditional};
  p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t node; 
  printf("add %d\n", (int)cudaGraphAddNode(&node, g, nullptr, 0, &p));
  cudaGraph_t bg = p.conditional.phGraph_out[0];
  cudaStreamBeginCaptureToGraph(st, bg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
  body<<<1,1,0,st>>>(d, h);
  cudaStreamEndCapture(st, nullptr);
  cudaGraphExec_t ex; printf("inst %d\n", (int)cudaGraphInstantiate(&ex, g, 0));
  cudaGraphLaunch(ex, st); cudaStreamSynchronize(st);
  int v; cudaMemcpy(&v, d, 4, cudaMemcpyDeviceToHost); printf("ctr %d\n", v);
}
