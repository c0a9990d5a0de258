// What would a device-side binning pass in front of K1 cost?  ("tile-ordered K1".)
// Times, for N atoms with DIM coordinates in random order and T coarse tiles, the three passes a tile-ordered
// update_forces needs around the evaluation itself:
//   count   : read x, derive the tile, per-CTA shared-memory histogram -> global counts, tile id kept (2 B/atom)
//   scatter : per-CTA reservation in every tile, xs[pos] = x[i], pos_of[i] = pos      (data moves: the evaluation then
//             streams xs and writes its result in sorted order)
//   unsort  : f[i] -= out_sorted[pos_of[i]] (32 B result record per atom), in the caller's order
// and, for comparison, the indices-only variant's extra traffic in the evaluation: gather x[perm[j]] + RMW f[perm[j]].
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tile_sort_cost tile_sort_cost.cu ; run: ./tile_sort_cost
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t err__ = (x); if (err__ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err__)); exit(1); } } while (0)

constexpr int kThreads = 1024, kItems = 8, kChunk = kThreads * kItems;

__device__ __forceinline__ unsigned mix(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return (unsigned)(z >> 33);
}

template <int DIM>
__global__ void __launch_bounds__(kThreads) count_kernel(long n, const double* __restrict__ x, int T, int per_dim,
                                                         unsigned short* __restrict__ tile_of, int* __restrict__ count) {
  extern __shared__ int hist[];
  for (int t = threadIdx.x; t <= T; t += blockDim.x) hist[t] = 0;
  __syncthreads();
  const long base = (long)blockIdx.x * kChunk;
#pragma unroll
  for (int k = 0; k < kItems; k++) {
    const long i = base + k * kThreads + threadIdx.x;
    if (i < n) {
      int t = 0, mul = 1;
#pragma unroll
      for (int d = 0; d < DIM; d++) {  // coordinates in [0,1): tile = floor(x * per_dim)
        int c = (int)(x[i * DIM + d] * per_dim);
        c = c < 0 ? 0 : (c >= per_dim ? per_dim - 1 : c);
        t += c * mul;
        mul *= per_dim;
      }
      tile_of[i] = (unsigned short)t;
      atomicAdd(&hist[t], 1);
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x)
    if (hist[t]) atomicAdd(&count[t], hist[t]);
}

__global__ void scan_kernel(int T, const int* __restrict__ count, int* __restrict__ start, int* __restrict__ cursor) {
  __shared__ int s[4096];
  for (int t = threadIdx.x; t < T; t += blockDim.x) s[t] = count[t];
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int t = 0; t < T; t++) {
      int c = s[t];
      s[t] = run;
      run += c;
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    start[t] = s[t];
    cursor[t] = 0;
  }
}

template <int DIM>
__global__ void __launch_bounds__(kThreads) scatter_kernel(long n, const double* __restrict__ x, int T,
                                                           const unsigned short* __restrict__ tile_of,
                                                           const int* __restrict__ start, int* __restrict__ cursor,
                                                           double* __restrict__ xs, int* __restrict__ pos_of) {
  extern __shared__ int hist[];  // [0,T): counts, then the CTA's base inside each tile
  for (int t = threadIdx.x; t < T; t += blockDim.x) hist[t] = 0;
  __syncthreads();
  const long base = (long)blockIdx.x * kChunk;
  int tl[kItems], lr[kItems];
#pragma unroll
  for (int k = 0; k < kItems; k++) {
    const long i = base + k * kThreads + threadIdx.x;
    tl[k] = -1;
    if (i < n) {
      tl[k] = tile_of[i];
      lr[k] = atomicAdd(&hist[tl[k]], 1);
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const int c = hist[t];
    hist[t] = c ? start[t] + atomicAdd(&cursor[t], c) : 0;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kItems; k++) {
    const long i = base + k * kThreads + threadIdx.x;
    if (tl[k] >= 0) {
      const int p = hist[tl[k]] + lr[k];
      pos_of[i] = p;
#pragma unroll
      for (int d = 0; d < DIM; d++) xs[(long)p * DIM + d] = x[i * DIM + d];
    }
  }
}

template <int DIM>
__global__ void unsort_kernel(long n, const int* __restrict__ pos_of, const double4* __restrict__ out_sorted,
                              double* __restrict__ f, double* __restrict__ partial) {
  double e = 0.0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const double4 o = out_sorted[pos_of[i]];
    f[i * DIM + 0] -= o.x;
    if (DIM > 1) f[i * DIM + 1] -= o.y;
    if (DIM > 2) f[i * DIM + 2] -= o.z;
    e += o.w;
  }
  if (e == 12345.678) partial[blockIdx.x] = e;
}

// indices-only variant: what the evaluation itself would have to do extra
template <int DIM>
__global__ void perm_touch_kernel(long n, const int* __restrict__ perm, const double* __restrict__ x,
                                  double* __restrict__ f) {
  for (long j = (long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long)gridDim.x * blockDim.x) {
    const long i = perm[j];
    double s = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; d++) s += x[i * DIM + d];
#pragma unroll
    for (int d = 0; d < DIM; d++) f[i * DIM + d] -= s;
  }
}

__global__ void fill_random(long n, int dim, double* x, unsigned long long seed) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n * dim; i += (long)gridDim.x * blockDim.x)
    x[i] = mix(seed + i) * (1.0 / 2147483648.0);
}
__global__ void invert_perm(long n, const int* pos_of, int* perm) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) perm[pos_of[i]] = (int)i;
}

template <int DIM> void run(long n, int per_dim) {
  int T = 1;
  for (int d = 0; d < DIM; d++) T *= per_dim;
  double *x, *xs, *f, *partial;
  double4* out_sorted;
  unsigned short* tile_of;
  int *count, *start, *cursor, *pos_of, *perm;
  char* flush;
  CK(cudaMalloc(&x, n * DIM * 8));
  CK(cudaMalloc(&xs, n * DIM * 8));
  CK(cudaMalloc(&f, n * DIM * 8));
  CK(cudaMalloc(&out_sorted, n * 32));
  CK(cudaMalloc(&tile_of, n * 2));
  CK(cudaMalloc(&count, (T + 1) * 4));
  CK(cudaMalloc(&start, (T + 1) * 4));
  CK(cudaMalloc(&cursor, (T + 1) * 4));
  CK(cudaMalloc(&pos_of, n * 4));
  CK(cudaMalloc(&perm, n * 4));
  CK(cudaMalloc(&partial, 4096 * 8));
  CK(cudaMalloc(&flush, 512u << 20));
  fill_random<<<1184, 256>>>(n, DIM, x, 77);
  CK(cudaMemset(f, 0, n * DIM * 8));
  CK(cudaMemset(out_sorted, 0, n * 32));
  const int chunks = (int)((n + kChunk - 1) / kChunk);
  const size_t sh = (T + 1) * sizeof(int);
  cudaEvent_t e[8];
  for (auto& ev : e) CK(cudaEventCreate(&ev));
  float best[5] = {1e9f, 1e9f, 1e9f, 1e9f, 1e9f};
  for (int rep = 0; rep < 5; rep++) {
    CK(cudaMemset(flush, rep, 512u << 20));
    CK(cudaMemsetAsync(count, 0, (T + 1) * 4));
    CK(cudaEventRecord(e[0]));
    count_kernel<DIM><<<chunks, kThreads, sh>>>(n, x, T, per_dim, tile_of, count);
    CK(cudaEventRecord(e[1]));
    scan_kernel<<<1, 1024>>>(T, count, start, cursor);
    CK(cudaEventRecord(e[2]));
    scatter_kernel<DIM><<<chunks, kThreads, sh>>>(n, x, T, tile_of, start, cursor, xs, pos_of);
    CK(cudaEventRecord(e[3]));
    CK(cudaMemset(flush, rep + 1, 512u << 20));
    CK(cudaEventRecord(e[4]));
    unsort_kernel<DIM><<<148 * 8, 256>>>(n, pos_of, out_sorted, f, partial);
    CK(cudaEventRecord(e[5]));
    invert_perm<<<1184, 256>>>(n, pos_of, perm);
    CK(cudaMemset(flush, rep + 2, 512u << 20));
    CK(cudaEventRecord(e[6]));
    perm_touch_kernel<DIM><<<148 * 8, 256>>>(n, perm, x, f);
    CK(cudaEventRecord(e[7]));
    CK(cudaDeviceSynchronize());
    float t;
    cudaEventElapsedTime(&t, e[0], e[1]); best[0] = fminf(best[0], t);
    cudaEventElapsedTime(&t, e[1], e[2]); best[1] = fminf(best[1], t);
    cudaEventElapsedTime(&t, e[2], e[3]); best[2] = fminf(best[2], t);
    cudaEventElapsedTime(&t, e[4], e[5]); best[3] = fminf(best[3], t);
    cudaEventElapsedTime(&t, e[6], e[7]); best[4] = fminf(best[4], t);
  }
  printf("{\"dim\": %d, \"atoms\": %ld, \"tiles\": %d, \"count_ms\": %.4f, \"scan_ms\": %.4f, \"scatter_ms\": %.4f, "
         "\"unsort_ms\": %.4f, \"data_moving_total_ms\": %.4f, \"indices_only_gather_rmw_ms\": %.4f}\n",
         DIM, n, T, best[0], best[1], best[2], best[3], best[0] + best[1] + best[2] + best[3], best[4]);
  cudaFree(x); cudaFree(xs); cudaFree(f); cudaFree(out_sorted); cudaFree(tile_of); cudaFree(count); cudaFree(start);
  cudaFree(cursor); cudaFree(pos_of); cudaFree(perm); cudaFree(partial); cudaFree(flush);
}

int main() {
  run<2>(10000000, 8);    // C3: 4096^2 points, 8 x 8 tiles of 512^2 points = 8.4 MB of records each
  run<2>(10000000, 16);   //      16 x 16 tiles of 2.1 MB
  run<3>(10000000, 8);    // C4: 512^3 points, 8^3 tiles of 64^3 points = 8.4 MB each
  run<3>(1250000, 8);     // C4 on 8 GPUs: one rank's atoms
  return 0;
}
