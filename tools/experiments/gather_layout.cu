// Memory-system experiment behind the K1 cell-record layout: random 2^D-corner gathers
// from point records (2 or 4 row pairs of 64 B at 32 B alignment) against one aligned cell record
// (128 B in 2-D, 256 B in 3-D), with the x read and f read-modify-write streams of forces_kernel.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_layout gather_layout.cu && ./gather_layout
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(const double* p, double* o) {
  double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}

// LAYOUT 0: point records; 1: cell records; 2: pair records {rec[p], rec[p+1]} (64 B, aligned) per point
template <int DIM, int CELL>
__global__ void __launch_bounds__(256, 3) gather(const double* __restrict__ rec, long n0, long n1, long n2, long n,
                                                 const double* __restrict__ x, double* __restrict__ f) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double xi[DIM], fo[DIM];
  for (int d = 0; d < DIM; d++) { xi[d] = x[i * DIM + d]; fo[d] = f[i * DIM + d]; }
  long c0 = (long)xi[0], c1 = (long)xi[1], c2 = DIM == 3 ? (long)xi[2] : 0;
  double acc[4] = {0, 0, 0, 0};
  if (CELL == 2) {
    for (int c = 0; c < (1 << (DIM - 1)); c++) {
      long a1 = c1 + (c & 1), a2 = c2 + ((c >> 1) & 1);
      if (a1 == n1) a1 = 0;
      if (DIM == 3 && a2 == n2) a2 = 0;
      const double* p = rec + ((a2 * n1 + a1) * n0 + c0) * 8;
      double r[4], q[4];
      ld32(p, r);
      ld32(p + 4, q);
      for (int k = 0; k < 4; k++) acc[k] += r[k] * (2 * c + 1) + q[k] * (2 * c + 2);
    }
  } else if (CELL == 3) {  // point records in 4x4x4 tiles (2 KB each), tiles in linear order
    for (int c = 0; c < (1 << DIM); c++) {
      long a0 = c0 + (c & 1), a1 = c1 + ((c >> 1) & 1), a2 = c2 + ((c >> 2) & 1);
      if (a0 == n0) a0 = 0;
      if (a1 == n1) a1 = 0;
      if (DIM == 3 && a2 == n2) a2 = 0;
      const long tile = ((a2 >> 2) * (n1 >> 2) + (a1 >> 2)) * (n0 >> 2) + (a0 >> 2);
      const long in = ((a2 & 3) * 4 + (a1 & 3)) * 4 + (a0 & 3);
      double r[4];
      ld32(rec + (tile * 64 + in) * 4, r);
      for (int k = 0; k < 4; k++) acc[k] += r[k] * (c + 1);
    }
  } else if (CELL == 1) {
    const double* p = rec + ((c2 * n1 + c1) * n0 + c0) * (4L << DIM);
    for (int c = 0; c < (1 << DIM); c++) {
      double r[4];
      ld32(p + 4 * c, r);
      for (int k = 0; k < 4; k++) acc[k] += r[k] * (c + 1);
    }
  } else {
    for (int c = 0; c < (1 << DIM); c++) {
      long a0 = c0 + (c & 1), a1 = c1 + ((c >> 1) & 1), a2 = c2 + ((c >> 2) & 1);
      if (a0 == n0) a0 = 0;
      if (a1 == n1) a1 = 0;
      if (DIM == 3 && a2 == n2) a2 = 0;
      double r[4];
      ld32(rec + ((a2 * n1 + a1) * n0 + a0) * 4, r);
      for (int k = 0; k < 4; k++) acc[k] += r[k] * (c + 1);
    }
  }
  for (int d = 0; d < DIM; d++) f[i * DIM + d] = fo[d] - acc[1 + d] - acc[0];
}

__global__ void fill(double* p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 1e-3 * (i % 977);
}
__global__ void rnd(double* x, long n, int dim, double s0, double s1, double s2) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long z = i * 0x9E3779B97F4A7C15ULL + 12345;
  for (int d = 0; d < dim; d++) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z ^= z >> 31;
    double u = (double)(z >> 11) * (1.0 / 9007199254740992.0);
    x[i * dim + d] = u * (d == 0 ? s0 : d == 1 ? s1 : s2) * 0.999999;
  }
}

template <int DIM, int CELL> float run(const double* rec, long n0, long n1, long n2, long n, const double* x, double* f, void* flush) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9;
  for (int it = 0; it < 5; it++) {
    cudaMemsetAsync(flush, 0, 512u << 20);
    cudaEventRecord(a);
    gather<DIM, CELL><<<(unsigned)((n + 255) / 256), 256>>>(rec, n0, n1, n2, n, x, f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (it > 0 && ms < best) best = ms;
  }
  return best;
}

int main() {
  void* flush; cudaMalloc(&flush, 512u << 20);
  {  // C3: 2-D 4096^2, 1e7 points
    long n0 = 4096, n1 = 4096, n = 10000000;
    double *pt, *cell, *x, *f;
    cudaMalloc(&pt, n0 * n1 * 32); cudaMalloc(&cell, n0 * n1 * 128); cudaMalloc(&x, n * 16); cudaMalloc(&f, n * 16);
    fill<<<1184, 256>>>(pt, n0 * n1 * 4); fill<<<1184, 256>>>(cell, n0 * n1 * 16); cudaMemset(f, 0, n * 16);
    rnd<<<(unsigned)((n + 255) / 256), 256>>>(x, n, 2, 4096, 4096, 1);
    printf("2-D 4096^2, 1e7 points: point records %.3f ms, cell records %.3f ms, pair records %.3f ms\n",
           run<2, 0>(pt, n0, n1, 1, n, x, f, flush), run<2, 1>(cell, n0, n1, 1, n, x, f, flush),
           run<2, 2>(cell, n0, n1, 1, n, x, f, flush));
    cudaFree(pt); cudaFree(cell); cudaFree(x); cudaFree(f);
  }
  {  // C4: 3-D 512^3, 1e7 points
    long n0 = 512, n1 = 512, n2 = 512, n = 10000000;
    double *pt, *cell, *x, *f;
    cudaMalloc(&pt, n0 * n1 * n2 * 32);
    if (cudaMalloc(&cell, n0 * n1 * n2 * 256) != cudaSuccess) { printf("no room for 3-D cell records\n"); return 0; }
    cudaMalloc(&x, n * 24); cudaMalloc(&f, n * 24);
    fill<<<1184, 256>>>(pt, n0 * n1 * n2 * 4); fill<<<1184, 256>>>(cell, n0 * n1 * n2 * 32); cudaMemset(f, 0, n * 24);
    rnd<<<(unsigned)((n + 255) / 256), 256>>>(x, n, 3, 512, 512, 512);
    printf("3-D 512^3, 1e7 points: point records %.3f ms, cell records %.3f ms, pair records %.3f ms, 4x4x4-tiled point records %.3f ms\n",
           run<3, 0>(pt, n0, n1, n2, n, x, f, flush), run<3, 1>(cell, n0, n1, n2, n, x, f, flush),
           run<3, 2>(cell, n0, n1, n2, n, x, f, flush), run<3, 3>(pt, n0, n1, n2, n, x, f, flush));
    rnd<<<(unsigned)((n + 255) / 256), 256>>>(x, n, 3, 512, 512, 512);
    printf("   1.25e6 points (one of 8 GPUs): point %.3f ms, tiled %.3f ms\n", run<3, 0>(pt, n0, n1, n2, n / 8, x, f, flush),
           run<3, 3>(pt, n0, n1, n2, n / 8, x, f, flush));
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
