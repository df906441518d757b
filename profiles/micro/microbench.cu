// Latency micro-benchmarks used to budget the latency-bound kernels of the step.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
namespace cg = cooperative_groups;

__global__ void empty_kernel(int *p) { if (p && threadIdx.x == 1234) *p = 1; }

__global__ void pdl_kernel(int *p) {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (p && threadIdx.x == 1234) *p = 1;
}

// pointer chase: each load depends on the previous one
__global__ void chase_kernel(const uint32_t *next, int steps, uint32_t *out, long long *cycles) {
  uint32_t i = threadIdx.x;
  long long t0 = clock64();
  for (int s = 0; s < steps; ++s) i = next[i];
  long long t1 = clock64();
  out[threadIdx.x] = i;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}

__global__ void dadd_kernel(const double *v, int n, double *out, long long *cycles) {
  double acc = v[0];
  long long t0 = clock64();
#pragma unroll 8
  for (int k = 1; k < n; ++k) acc = __dadd_rn(acc, (double)k);
  long long t1 = clock64();
  *out = acc;
  *cycles = t1 - t0;
}

__global__ void gridsync_kernel(int reps, long long *cycles) {
  cg::grid_group g = cg::this_grid();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) g.sync();
  long long t1 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) *cycles = t1 - t0;
}

__global__ void syncthreads_kernel(int reps, long long *cycles) {
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) *cycles = t1 - t0;
}

static float time_graph(cudaStream_t s, cudaGraphExec_t g, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 5; ++i) cudaGraphLaunch(g, s);
  cudaStreamSynchronize(s);
  cudaEventRecord(a, s);
  for (int i = 0; i < reps; ++i) cudaGraphLaunch(g, s);
  cudaEventRecord(b, s);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms * 1000.f / reps;
}

int main() {
  cudaStream_t s; cudaStreamCreate(&s);
  int *flag; cudaMalloc(&flag, 4);
  // 1. chain of N dependent tiny kernels in a graph, with and without PDL
  for (int pdl = 0; pdl < 2; ++pdl) {
    for (int blocks : {1, 128}) {
      const int N = 40;
      cudaGraph_t graph; cudaGraphExec_t exec;
      cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
      for (int k = 0; k < N; ++k) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(128); cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = pdl;
        if (pdl) cudaLaunchKernelEx(&cfg, pdl_kernel, flag);
        else cudaLaunchKernelEx(&cfg, empty_kernel, flag);
      }
      cudaStreamEndCapture(s, &graph);
      cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
      if (e != cudaSuccess) { printf("instantiate failed: %s\n", cudaGetErrorString(e)); continue; }
      float us = time_graph(s, exec, 50);
      printf("graph chain of %d kernels (%d blocks, pdl=%d): %.2f us per kernel\n", N, blocks, pdl, us / N);
    }
  }
  // 2. dependent-load latency: L2-resident (4 MB table) and DRAM (2 GB table)
  for (size_t mb : {4, 2048}) {
    size_t n = mb * 1024 * 1024 / 4;
    uint32_t *h = (uint32_t *)malloc(n * 4);
    uint64_t x = 88172645463325252ull;
    for (size_t i = 0; i < n; ++i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; h[i] = (uint32_t)(x % n); }
    uint32_t *d, *out; long long *cyc;
    cudaMalloc(&d, n * 4); cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; ++rep) chase_kernel<<<1, 32, 0, s>>>(d, 2000, out, cyc);
    cudaStreamSynchronize(s);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent load, %zu MB table: %.0f cycles per load\n", mb, c / 2000.0);
    cudaFree(d); cudaFree(out); cudaFree(cyc); free(h);
  }
  // 3. DADD chain
  {
    double *v, *out; long long *cyc;
    cudaMalloc(&v, 8); cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    cudaMemset(v, 0, 8);
    dadd_kernel<<<1, 1, 0, s>>>(v, 4096, out, cyc);
    dadd_kernel<<<1, 1, 0, s>>>(v, 4096, out, cyc);
    cudaStreamSynchronize(s);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent DADD: %.1f cycles each\n", c / 4095.0);
  }
  // 4. grid.sync with 21 CTAs x 1024 threads, and 148 CTAs x 256
  for (int cfg = 0; cfg < 2; ++cfg) {
    int blocks = cfg ? 148 : 21, threads = cfg ? 256 : 1024, reps = 50;
    long long *cyc; cudaMalloc(&cyc, 8);
    void *args[] = {&reps, &cyc};
    cudaLaunchCooperativeKernel((void *)gridsync_kernel, dim3(blocks), dim3(threads), args, 0, s);
    cudaLaunchCooperativeKernel((void *)gridsync_kernel, dim3(blocks), dim3(threads), args, 0, s);
    cudaStreamSynchronize(s);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("grid.sync %d x %d: %.0f cycles\n", blocks, threads, c / (double)reps);
  }
  {
    long long *cyc; cudaMalloc(&cyc, 8);
    syncthreads_kernel<<<1, 1024, 0, s>>>(100, cyc);
    cudaStreamSynchronize(s);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("__syncthreads (1024 threads): %.0f cycles\n", c / 100.0);
  }
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("clock rate attr %d kHz; last error: %s\n", clk, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
