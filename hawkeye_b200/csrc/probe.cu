// Hardware probes used while developing the kernels (not part of the product path or the ABI header):
//   hk_debug_probe_tmem_a : (1) dumps the shared-memory image of an MN-major tile as TMA writes it with
//                           CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; (2) runs D = A^T-staged-in-TMEM . B with the A operand written
//                           to tensor memory by tcgen05.st (lane = m, column = k) and read by tcgen05.mma [tmem], desc.
#include "common.cuh"
#include "host.h"

namespace hk {

__global__ void __launch_bounds__(192, 1)
probe_tmem_a_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const float* Asrc,
                    float* smem_dump, float* D) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                  // 4 boxes of [64 k-rows][32 m] = 4 x 8 KB
  uint8_t* sB = smem + 32768;          // [32 n][64 k] as two K-major k-blocks of [32 rows x 128 B] = 2 x 4 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768 + 8192);
  uint64_t* done = bar + 1;
  uint64_t* a_ready = bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1); mbar_init(done, 1); mbar_init(a_ready, 4);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 128); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0 && lane == 0) {
    mbar_expect_tx(bar, 32768 + 8192);
    for (int j = 0; j < 4; ++j) tma_load_2d(sA + j * 8192, &tmA, bar, j * 32, 0);
    for (int kb = 0; kb < 2; ++kb) tma_load_2d(sB + kb * 4096, &tmB, bar, kb * 32, 0);
  }
  if (warp >= 2) {     // 4 warps: lane quarter = warp % 4
    const int q = warp & 3;
    mbar_wait(bar, 0);
    // (1) raw shared-memory image of the MN-major A tile
    for (int i = (warp - 2) * 32 + lane; i < 8192; i += 128) smem_dump[i] = reinterpret_cast<const float*>(sA)[i];
    // (2) A operand into tensor memory: lane m = 32 q + lane, columns k = 0..63 (read from global: layout-independent)
    const int m = q * 32 + lane;
    float v[32];
    for (int half = 0; half < 2; ++half) {
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] = Asrc[(size_t)(half * 32 + k) * 128 + m];
      tmem_st32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 32 + half * 32, v);   // columns 32..95: A
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(a_ready);
    mbar_wait(done, 0);
    tc_fence_after();
    tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16), v);                        // columns 0..31: D
    tmem_ld_wait();
    for (int n = 0; n < 32; ++n) D[(size_t)m * 32 + n] = v[n];
  } else if (warp == 1) {
    mbar_wait(bar, 0);
    mbar_wait(a_ready, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_tf32(128, 32, 0, 0);
    if (elect_one()) {
      for (int kb = 0; kb < 2; ++kb) {
        const uint64_t bdesc = make_sdesc(smem_u32(sB + kb * 4096), 16, 1024);
        for (int ks = 0; ks < 4; ++ks)
          umma_tf32_ts(tmem_base, tmem_base + 32 + kb * 32 + ks * 8, bdesc + ks * 2, idesc, (kb | ks) ? 1u : 0u);
      }
      umma_commit(done);
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}

}  // namespace hk

using namespace hk;

// Asrc [64][128] (k-major rows of m), Bsrc [32][64] (n rows of k), smem_dump [8192], D [128][32] — all device fp32
extern "C" int hk_debug_probe_tmem_a(const float* Asrc, const float* Bsrc, float* smem_dump, float* D, void* stream) {
  CUtensorMap tmA, tmB;
  int r;
  {
    uint64_t dims[2] = {128, 64};
    uint64_t strides[1] = {128 * 4};
    uint32_t box[2] = {32, 64};
    if ((r = make_tmap(&tmA, Asrc, 2, dims, strides, box, /*mn_major=*/true))) return r;
  }
  {
    uint64_t dims[2] = {64, 32};
    uint64_t strides[1] = {64 * 4};
    uint32_t box[2] = {32, 32};
    if ((r = make_tmap(&tmB, Bsrc, 2, dims, strides, box))) return r;
  }
  cudaFuncSetAttribute(probe_tmem_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  probe_tmem_a_kernel<<<1, 192, 32768 + 8192 + 1024 + 256, (cudaStream_t)stream>>>(tmA, tmB, Asrc, smem_dump, D);
  HK_LAUNCH_CHECK("probe_tmem_a_kernel");
  return 0;
}
