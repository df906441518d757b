// What a round of the warp descent costs, piece by piece (one warp, tree warm in L2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include \
//        -I dopamine_b200/csrc -o profiles/micro/descent_bench profiles/micro/descent_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tree.cuh"

using namespace b2r;

__global__ void descent_kernel(const double *heap, int depth, const double *q01, int n,
                               long long *out, long long *cycles) {
  const int lane = threadIdx.x & 31;
  const double total = heap[1];
  long long t_load = 0, t_walk = 0, t_all = 0, t_philox = 0;
  long long acc = 0;
  for (int k = 0; k < n; ++k) {
    long long t0 = clock64();
    const double u = philox_uniform53_fast(1234, 77, (uint64_t)k);
    long long t1 = clock64();
    t_philox += t1 - t0;
    double q = __dmul_rn(q01[k] + u * 1e-30, total);
    int64_t h = 1;
    int level = 0;
    long long ta = clock64();
    while (level < depth) {
      const int K = depth - level < 5 ? depth - level : 5;
      long long a0 = clock64();
      double c = warp_candidate(heap, h, K, lane);
      // consume the load
      if (__double_as_longlong(c) == 0x7fffffffffffffffll) acc++;
      long long a1 = clock64();
      warp_walk(c, K, lane, h, q);
      long long a2 = clock64();
      t_load += a1 - a0;
      t_walk += a2 - a1;
      level += K;
    }
    t_all += clock64() - ta;
    acc += h;
  }
  if (threadIdx.x == 0) {
    out[0] = acc;
    cycles[0] = t_load; cycles[1] = t_walk; cycles[2] = t_all; cycles[3] = t_philox;
  }
}

int main() {
  const int depth = 20;
  const size_t nodes = (size_t)2 << depth;
  std::vector<double> heap(nodes, 0.0);
  srand(1);
  for (size_t i = (size_t)1 << depth; i < nodes; ++i) heap[i] = (rand() % 1000) / 1000.0;
  for (size_t h = ((size_t)1 << depth) - 1; h >= 1; --h) heap[h] = heap[2 * h] + heap[2 * h + 1];
  const int n = 256;
  std::vector<double> q(n);
  for (int k = 0; k < n; ++k) q[k] = (rand() % 100000) / 100000.0;
  double *dheap, *dq;
  long long *dout, *dcyc;
  cudaMalloc(&dheap, nodes * 8); cudaMalloc(&dq, n * 8);
  cudaMalloc(&dout, 64); cudaMalloc(&dcyc, 64);
  cudaMemcpy(dheap, heap.data(), nodes * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dq, q.data(), n * 8, cudaMemcpyHostToDevice);
  for (int rep = 0; rep < 3; ++rep) {
    descent_kernel<<<1, 32>>>(dheap, depth, dq, n, dout, dcyc);
    cudaDeviceSynchronize();
    long long cyc[4];
    cudaMemcpy(cyc, dcyc, 32, cudaMemcpyDeviceToHost);
    printf("rep %d: per descent (depth %d, 4 rounds): load+consume %.0f, walk %.0f, all %.0f "
           "cycles; philox %.0f\n", rep, depth, (double)cyc[0] / n, (double)cyc[1] / n,
           (double)cyc[2] / n, (double)cyc[3] / n);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
