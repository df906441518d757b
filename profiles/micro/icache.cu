// How expensive is cold straight-line code?  A kernel of N unrolled dependent
// integer ops, run (a) back to back (warm) and (b) alternating with a different big
// kernel on all SMs (cold), single warp.
#include <cuda_runtime.h>
#include <cstdio>

template <int N>
__global__ void straight(unsigned *out, unsigned seed, long long *cyc) {
  unsigned x = seed + threadIdx.x;
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) x = x * 1664525u + (x >> 7) + i;   // ~3 dependent ops
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int N>
__global__ void polluter(unsigned *out, unsigned seed) {
  unsigned x = seed + threadIdx.x + blockIdx.x;
#pragma unroll
  for (int i = 0; i < N; ++i) x = (x ^ (x << 3)) + 0x9e3779b9u * i;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

template <int N>
void run(unsigned *out, long long *cyc, cudaStream_t s) {
  long long c;
  // warm: same kernel repeatedly
  for (int r = 0; r < 5; ++r) straight<N><<<1, 32, 0, s>>>(out, r, cyc);
  cudaStreamSynchronize(s);
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  double warm = (double)c;
  // cold: pollute every SM's instruction cache in between
  double cold = 0;
  for (int r = 0; r < 5; ++r) {
    polluter<6000><<<296, 128, 0, s>>>(out + 1024, r);
    straight<N><<<1, 32, 0, s>>>(out, r, cyc);
    cudaStreamSynchronize(s);
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    cold += c / 5.0;
  }
  printf("N=%5d ops x3 instr: warm %.0f cycles (%.1f/instr)  cold %.0f cycles (%.1f/instr)\n",
         N, warm, warm / (3.0 * N), cold, cold / (3.0 * N));
}

int main() {
  cudaStream_t s; cudaStreamCreate(&s);
  unsigned *out; long long *cyc;
  cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
  run<100>(out, cyc, s);
  run<300>(out, cyc, s);
  run<1000>(out, cyc, s);
  run<3000>(out, cyc, s);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
