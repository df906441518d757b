// Is the per-phase cost of the batch-32 kernels instruction fetch?  A kernel of P
// phases (each ~70 dependent instructions behind a barrier and a data-dependent
// branch, so the fetch is not a straight line) runs R rounds inside ONE launch; the
// first round fetches every phase's code cold, later rounds find it in the SM's
// instruction caches.  Other kernels run between launches, as in the real step.
#include <cuda_runtime.h>
#include <cstdio>

#define PHASE(k)                                                         \
  __syncthreads();                                                       \
  if (sel[k] != 12345) {                                                 \
    _Pragma("unroll") for (int i = 0; i < 24; ++i)                       \
        x = x * (1664525u + k) + (x >> (7 + (k & 3))) + i;               \
    if (threadIdx.x == 0) marks[r * 16 + k] = clock64();                 \
  }

__global__ void phases(unsigned *out, const int *sel, long long *marks, int rounds) {
  unsigned x = threadIdx.x + sel[0];
  for (int r = 0; r < rounds; ++r) {
    if (threadIdx.x == 0) marks[r * 16 + 15] = clock64();
    PHASE(0) PHASE(1) PHASE(2) PHASE(3) PHASE(4) PHASE(5) PHASE(6) PHASE(7)
    PHASE(8) PHASE(9) PHASE(10) PHASE(11)
  }
  out[threadIdx.x] = x;
}

template <int N>
__global__ void polluter(unsigned *out, unsigned seed) {
  unsigned x = seed + threadIdx.x + blockIdx.x;
#pragma unroll
  for (int i = 0; i < N; ++i) x = (x ^ (x << 3)) + 0x9e3779b9u * i;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

int main() {
  unsigned *out; int *sel; long long *marks;
  cudaMalloc(&out, 1 << 22); cudaMalloc(&sel, 64); cudaMalloc(&marks, 16 * 8 * 8);
  cudaMemset(sel, 0, 64);
  long long h[16 * 8];
  for (int trial = 0; trial < 3; ++trial) {
    polluter<4000><<<296, 128>>>(out + 4096, trial);
    polluter<3000><<<296, 128>>>(out + 4096, trial + 7);
    phases<<<1, 672>>>(out, sel, marks, 3);
    cudaDeviceSynchronize();
    cudaMemcpy(h, marks, sizeof(h), cudaMemcpyDeviceToHost);
    for (int r = 0; r < 3; ++r) {
      printf("trial %d round %d: total %lld cycles; per phase:", trial, r,
             h[r * 16 + 11] - h[r * 16 + 15]);
      long long prev = h[r * 16 + 15];
      for (int k = 0; k < 12; ++k) { printf(" %lld", h[r * 16 + k] - prev); prev = h[r * 16 + k]; }
      printf("\n");
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
