// Store-pattern micro-benchmark (test infrastructure): what HBM write bandwidth does a B200 give for the bilinear-pool
// output pattern?  Y = [B][512][512] fp32 written tile by tile (128 x 128) by persistent CTAs, with different per-instruction
// contiguity, against a plain linear fill and a linear read.   nvcc -arch=sm_100a -O3 -o ubench_store ubench_store.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void fill_linear(float4* y, size_t n4, int cs) {
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    if (cs) __stcs(y + i, v); else y[i] = v;
  }
}
__global__ void read_linear(const float4* x, size_t n4, float* out) {
  float s = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldcs(x + i);
    s += v.x + v.y + v.z + v.w;
  }
  if (s == 123.456f) *out = s;
}
// transposed pattern of the kernel: a warp store = 32 consecutive floats (128 B) of one row; 4 warps -> 512 B; rows 2 KB apart
__global__ void tile_t128(float* y, int B, int cs, const float* x, int xbytes_per_item) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = warp & 3, h = warp >> 2;
  const int items = B * 16;
  float acc = 0.f;
  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    const int b = it >> 4, ti = (it >> 2) & 3, tj = it & 3;
    if (x) {   // emulate the operand reads (L2-resident after the first toucher)
      const float4* xb = reinterpret_cast<const float4*>(x + (size_t)b * 512 * 196);
      for (int i = threadIdx.x; i < xbytes_per_item / 16; i += blockDim.x) { const float4 v = __ldg(xb + (i % (512 * 196 / 4))); acc += v.x; }
    }
    for (int c = 2 * h; c < 2 * h + 2; ++c) {
      float* p = y + (size_t)b * 262144 + (size_t)(tj * 128 + c * 32) * 512 + ti * 128 + q * 32 + lane;
#pragma unroll
      for (int j = 0; j < 32; ++j) { if (cs) __stcs(p + j * 512, (float)j + acc); else p[j * 512] = (float)j + acc; }
    }
  }
}
// row-major float4: a warp store = 512 B of one row (one tile row); 8 warps -> 8 rows
__global__ void tile_v4(float* y, int B, int cs) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = B * 16;
  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    const int b = it >> 4, ti = (it >> 2) & 3, tj = it & 3;
    for (int rr = warp; rr < 128; rr += 8) {
      float4* p = reinterpret_cast<float4*>(y + (size_t)b * 262144 + (size_t)(ti * 128 + rr) * 512 + tj * 128) + lane;
      const float4 v = make_float4(1.f, 2.f, 3.f, (float)rr);
      if (cs) __stcs(p, v); else *p = v;
    }
  }
}
// full rows: item = 32 complete rows (2 KB each) of an image: a warp writes 4 x 512 B of the same row
__global__ void rows_v4(float* y, int B, int cs) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = B * 16;
  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    const int b = it >> 4, r0 = (it & 15) * 32;
    for (int rr = warp; rr < 32; rr += 8) {
      float4* p = reinterpret_cast<float4*>(y + (size_t)b * 262144 + (size_t)(r0 + rr) * 512) + lane;
      const float4 v = make_float4(1.f, 2.f, 3.f, (float)rr);
#pragma unroll
      for (int k = 0; k < 4; ++k) { if (cs) __stcs(p + 32 * k, v); else p[32 * k] = v; }
    }
  }
}

template <class F>
static float timeit(F f, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

int main() {
  const int B = 1024;
  const size_t ybytes = (size_t)B * 262144 * 4, xbytes = (size_t)B * 512 * 196 * 4;
  float *y, *x, *out;
  CK(cudaMalloc(&y, ybytes)); CK(cudaMalloc(&x, xbytes)); CK(cudaMalloc(&out, 4));
  CK(cudaMemset(x, 0, xbytes));
  const double gb = ybytes / 1e9;
  float ms;
  for (int cs = 0; cs < 2; ++cs) {
    ms = timeit([&] { fill_linear<<<148 * 8, 256>>>((float4*)y, ybytes / 16, cs); }, 5);
    printf("fill_linear cs=%d            : %.1f us  %.0f GB/s\n", cs, ms * 1e3, gb / ms * 1e3);
    ms = timeit([&] { tile_t128<<<148, 256>>>(y, B, cs, nullptr, 0); }, 5);
    printf("tile_t128 (kernel pattern) cs=%d: %.1f us  %.0f GB/s\n", cs, ms * 1e3, gb / ms * 1e3);
    ms = timeit([&] { tile_t128<<<296, 256>>>(y, B, cs, nullptr, 0); }, 5);
    printf("tile_t128 2 CTA/SM cs=%d     : %.1f us  %.0f GB/s\n", cs, ms * 1e3, gb / ms * 1e3);
    ms = timeit([&] { tile_v4<<<148, 256>>>(y, B, cs); }, 5);
    printf("tile_v4 (512 B rows) cs=%d   : %.1f us  %.0f GB/s\n", cs, ms * 1e3, gb / ms * 1e3);
    ms = timeit([&] { tile_v4<<<296, 256>>>(y, B, cs); }, 5);
    printf("tile_v4 2 CTA/SM cs=%d       : %.1f us  %.0f GB/s\n", cs, ms * 1e3, gb / ms * 1e3);
    ms = timeit([&] { rows_v4<<<148, 256>>>(y, B, cs); }, 5);
    printf("rows_v4 (2 KB rows) cs=%d    : %.1f us  %.0f GB/s\n", cs, ms * 1e3, gb / ms * 1e3);
    ms = timeit([&] { rows_v4<<<296, 256>>>(y, B, cs); }, 5);
    printf("rows_v4 2 CTA/SM cs=%d       : %.1f us  %.0f GB/s\n", cs, ms * 1e3, gb / ms * 1e3);
    ms = timeit([&] { tile_t128<<<148, 256>>>(y, B, cs, x, 200 * 1024); }, 5);
    printf("tile_t128 + X reads cs=%d    : %.1f us  %.0f GB/s (Y+X algorithmic)\n", cs, ms * 1e3, (ybytes + xbytes) / 1e9 / ms * 1e3);
  }
  ms = timeit([&] { read_linear<<<148 * 8, 256>>>((const float4*)y, ybytes / 16, out); }, 5);
  printf("read_linear                  : %.1f us  %.0f GB/s\n", ms * 1e3, gb / ms * 1e3);
  ms = timeit([&] { CK(cudaMemcpyAsync(y, y + ybytes / 8, ybytes / 2, cudaMemcpyDeviceToDevice)); }, 5);
  printf("memcpy d2d (r+w)             : %.1f us  %.0f GB/s\n", ms * 1e3, gb / ms * 1e3);
  ms = timeit([&] { CK(cudaMemsetAsync(y, 0, ybytes)); }, 5);
  printf("memset                       : %.1f us  %.0f GB/s\n", ms * 1e3, gb / ms * 1e3);
  return 0;
}
