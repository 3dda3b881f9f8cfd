// stream_bench.cu -- calibration microbenchmarks for the baseline-pass kernels (not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_bench stream_bench.cu && ./stream_bench
// Measures, for an 80 MB uint32 array (the size of the tiled item-major layout), what a B200 sustains for
//   (a) plain streaming sums with 4-byte and 16-byte loads per thread, one-shot grid vs persistent grid
//   (b) streaming + one / two random 8-byte shared-memory gathers per element
//   (c) (b) + the fp64 arithmetic of the deviation
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void sum4(const uint32_t* __restrict__ p, size_t n, unsigned long long* out) {
  unsigned long long acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) acc += p[i];
  acc += __shfl_xor_sync(~0u, acc, 16); acc += __shfl_xor_sync(~0u, acc, 8); acc += __shfl_xor_sync(~0u, acc, 4);
  acc += __shfl_xor_sync(~0u, acc, 2); acc += __shfl_xor_sync(~0u, acc, 1);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

__global__ void sum16(const uint4* __restrict__ p, size_t n4, unsigned long long* out) {
  unsigned long long acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v = p[i];
    acc += (unsigned long long)v.x + v.y + v.z + v.w;
  }
  acc += __shfl_xor_sync(~0u, acc, 16); acc += __shfl_xor_sync(~0u, acc, 8); acc += __shfl_xor_sync(~0u, acc, 4);
  acc += __shfl_xor_sync(~0u, acc, 2); acc += __shfl_xor_sync(~0u, acc, 1);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// streaming + G random 8-byte gathers from a shared table of TU doubles (+ optional fp64 deviation math)
template <int TU, int G, bool MATH>
__global__ void gather_k(const uint32_t* __restrict__ p, size_t n, const double* __restrict__ tab, double* out) {
  extern __shared__ double s[];
  for (int x = threadIdx.x; x < TU * (G > 1 ? 3 : 1); x += blockDim.x) s[x] = tab[x % TU];
  __syncthreads();
  double acc = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t e = p[i];
    const uint32_t ul = e & (TU - 1);
    const double a = s[ul];
    if (MATH) {
      const uint32_t code = (e >> 16) & 0xffu;
      const double d = fma((double)code, 0.5, -a);
      double inv;
      if (G > 1) inv = s[TU + (d > 0.0 ? ul : ul + TU)]; else inv = 0.25;
      double dev = d * inv;
      dev = (d != 0.0 && (e & 0x80000000u)) ? dev : 0.0;
      acc += dev;
    } else {
      acc += a;
      if (G > 1) acc += s[TU + ul];
    }
  }
  acc += __shfl_xor_sync(~0u, acc, 16); acc += __shfl_xor_sync(~0u, acc, 8); acc += __shfl_xor_sync(~0u, acc, 4);
  acc += __shfl_xor_sync(~0u, acc, 2); acc += __shfl_xor_sync(~0u, acc, 1);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

template <typename F>
float time_it(F f, int reps, void* flush, size_t flush_bytes) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9f;
  for (int r = 0; r < reps; ++r) {
    cudaMemsetAsync(flush, r, flush_bytes);
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  const size_t n = 20u * 1024 * 1024;  // 80 MB of uint32
  uint32_t* d; unsigned long long* o; double* od; double* tab; void* flush;
  CK(cudaMalloc(&d, n * 4)); CK(cudaMalloc(&o, 8)); CK(cudaMalloc(&od, 8)); CK(cudaMalloc(&tab, 8192 * 8)); CK(cudaMalloc(&flush, 256u << 20));
  std::vector<uint32_t> h(n);
  uint32_t x = 12345;
  for (size_t i = 0; i < n; ++i) { x = x * 1664525u + 1013904223u; h[i] = 0x80000000u | (((x >> 8) % 10 + 1) << 16) | ((x >> 12) & 0x1fff); }
  CK(cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice));
  std::vector<double> ht(8192); for (int i = 0; i < 8192; ++i) ht[i] = 3.0 + (i % 100) * 0.01;
  CK(cudaMemcpy(tab, ht.data(), 8192 * 8, cudaMemcpyHostToDevice));
  const double MB = n * 4 / 1e6;
  for (int per_sm : {2, 4, 8, 16}) {
    int grid = 148 * per_sm;
    float t4 = time_it([&] { sum4<<<grid, 256>>>(d, n, o); }, 10, flush, 256u << 20);
    float t16 = time_it([&] { sum16<<<grid, 256>>>((const uint4*)d, n / 4, o); }, 10, flush, 256u << 20);
    printf("persistent grid %4d x256: sum4 %.1f us (%.0f GB/s)  sum16 %.1f us (%.0f GB/s)\n", grid, t4 * 1e3, MB / t4, t16 * 1e3, MB / t16);
  }
  {
    int g4 = (int)((n + 255) / 256), g16 = (int)((n / 4 + 255) / 256);
    float t4 = time_it([&] { sum4<<<g4, 256>>>(d, n, o); }, 10, flush, 256u << 20);
    float t16 = time_it([&] { sum16<<<g16, 256>>>((const uint4*)d, n / 4, o); }, 10, flush, 256u << 20);
    printf("one-shot grid: sum4 %.1f us (%.0f GB/s)  sum16 %.1f us (%.0f GB/s)\n", t4 * 1e3, MB / t4, t16 * 1e3, MB / t16);
  }
  cudaFuncSetAttribute(gather_k<8192, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(gather_k<8192, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(gather_k<8192, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608);
  cudaFuncSetAttribute(gather_k<8192, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608);
  cudaFuncSetAttribute(gather_k<4096, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
  for (int threads : {512, 1024}) {
    int g1 = 148 * (threads == 512 ? 3 : 1), g2 = 148;
    float a = time_it([&] { gather_k<8192, 1, false><<<g1, threads, 65536>>>(d, n, tab, od); }, 10, flush, 256u << 20);
    float b = time_it([&] { gather_k<8192, 1, true><<<g1, threads, 65536>>>(d, n, tab, od); }, 10, flush, 256u << 20);
    float c = time_it([&] { gather_k<8192, 2, false><<<g2, 1024, 196608>>>(d, n, tab, od); }, 10, flush, 256u << 20);
    float e = time_it([&] { gather_k<8192, 2, true><<<g2, 1024, 196608>>>(d, n, tab, od); }, 10, flush, 256u << 20);
    float f = time_it([&] { gather_k<4096, 2, true><<<148 * 2, 1024, 98304>>>(d, n, tab, od); }, 10, flush, 256u << 20);
    printf("threads %4d: 1 gather %.1f us (%.0f GB/s) | 1 gather+math %.1f us (%.0f) | 2 gathers(1024thr,192KB) %.1f us (%.0f) | 2 gathers+math %.1f us (%.0f) | TU4096 2x1024thr 2g+math %.1f us (%.0f)\n",
           threads, a * 1e3, MB / a, b * 1e3, MB / b, c * 1e3, MB / c, e * 1e3, MB / e, f * 1e3, MB / f);
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
