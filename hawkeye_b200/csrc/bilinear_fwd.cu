// K1 — fused bilinear pooling forward for C = 512 (reference model/methods/BCNN.py:13-27):
//     G = X X^T / HW ; z = sqrt(G + 1e-5) ; y = z / max(||z||_2, 1e-12)            X: [B, 512, HW]  ->  y: [B, 512*512]
// as ONE launch of thread-block clusters of four CTAs, one image per cluster at a time.
//
//   * X crosses L2 -> SM once per image: for every 32-column k-block each CTA of the cluster TMA-loads ITS 128-row block
//     of X and MULTICASTS it into the shared memory of all four CTAs (cp.async.bulk.tensor ... .multicast::cluster), so
//     every CTA holds the whole 512 x 32 k-block (64 KB per stage, 3 stages) while L2 serves each byte once.
//   * CTA r computes block-row r of the Gram on tcgen05 (kind::tf32): A = row block r, B = all 512 rows as two N=256
//     operands -> a 128 x 512 fp32 accumulator = the CTA's entire TMEM.
//   * The L2 norm needs no second pass, no pre-kernel and no exchange between CTAs:  ||z||^2 = sum_ij G_ij/HW + C^2 eps
//     = sum_p (sum_c x_cp)^2 / HW + C^2 eps,  and every CTA sees all of X go through its shared memory: four warps
//     accumulate the per-location channel sums s_p from the staged tiles while the tensor core consumes them, so 1/||z||
//     is known before the accumulator is complete.  (Values are summed as the tensor core sees them: low 13 mantissa bits
//     dropped.)  Nothing in the kernel waits on another cluster, and the CTAs of a cluster are co-scheduled by the
//     hardware — there is no residency assumption and no library-owned global state.
//   * Epilogue: 8 warps read TMEM, apply sqrt(.+eps) / ||z||, and write block-COLUMN r of Y — legal because G is
//     symmetric — so the 32 lanes of a warp (= 32 consecutive rows of the accumulator) store 32 consecutive floats of one
//     row of Y: every store instruction is one full 128-byte line, with no shared-memory staging.
//   * shared-memory stages are released cluster-wide: a stage may be overwritten by any CTA's multicast only when all four
//     CTAs are done with it, so the MMA warp commits with tcgen05.commit...multicast::cluster to the `empty` barrier of all
//     four CTAs, and the channel-sum warps arrive remotely (mapa + mbarrier.arrive.shared::cluster); count = 4 + 4.
//
// Algorithmic traffic: 401 408 B read + 1 048 576 B written per image; MMA work 2*512*512*200 flop/image.
#include <stdlib.h>

#include "common.cuh"
#include "host.h"
#include "../../include/hawkeye_b200.h"

namespace hk {

constexpr int CF_C = 512;
constexpr int CF_CLUSTER = 4;
constexpr int CF_STAGES = 3;
constexpr int CF_SLOT = 128 * 128;                 // 16 KB: 128 rows x 32 fp32 (one row block of one k-block)
constexpr int CF_STAGE_BYTES = CF_CLUSTER * CF_SLOT;   // 64 KB: the whole 512-row k-block
constexpr int CF_THREADS = 32 * 14;                // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue, warps 10-13 channel sums
constexpr int CF_SMEM = CF_STAGES * CF_STAGE_BYTES + 1024 /*alignment*/ + 2048 /*barriers, partial sums*/;

struct CfArgs {
  int B, HW;
  float inv_hw, eps;
  float* Y;          // [B][512*512]
  float* inv_norm;   // [B] or null
  int pdl;           // launched with programmatic stream serialization
  int dbg;           // profiling only ($HK_K1_DBG): 1 skip the channel-sum reads, 2 skip the global stores, 4 skip the MMAs
  unsigned long long* trace;   // profiling only: [grid][16] %globaltimer stamps, or null
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_count_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA tile load delivered to the same shared-memory offset (and signalling the same-offset mbarrier) of every CTA in mask
__device__ __forceinline__ void tma_load_3d_mcast(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                                  uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5, "
      "%6}], [%2], %3;" ::"r"(smem_u32(dst)),
      "l"(m), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// tcgen05.commit arriving on the same-offset mbarrier of every CTA in mask once the issued MMAs have retired
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// arrive on the same-offset mbarrier of CTA `cta` of this cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// wait with cluster-scope acquire (the barrier is arrived on by other CTAs of the cluster)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (++spins > HK_SPIN_LIMIT) {
      printf("hawkeye_b200: cluster mbarrier watchdog (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__global__ void __launch_bounds__(CF_THREADS, 1)
bcnn_cluster_fwd_kernel(const __grid_constant__ CUtensorMap tmX, CfArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stages = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + CF_STAGES * CF_STAGE_BYTES);   // [3] TMA (all 4 CTAs) -> consumers
  uint64_t* empty = full + CF_STAGES;       // [3] consumers of ALL four CTAs -> producer (count 8)
  uint64_t* acc_full = empty + CF_STAGES;   // [1] MMA -> epilogue
  uint64_t* acc_empty = acc_full + 1;       // [1] epilogue (8 warps) -> MMA
  uint64_t* norm_ready = acc_empty + 1;     // [2] channel-sum warps -> epilogue, per image parity
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(norm_ready + 2);
  float* inv_box = reinterpret_cast<float*>(tmem_slot + 2);   // [2]
  float* part = inv_box + 2;                                   // [2][4][32] per-warp column sums of a k-block

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = (int)cluster_id_x(), ncl = (int)cluster_count_x();
  const int nk = (a.HW + 31) / 32;
  const uint16_t all = (1u << CF_CLUSTER) - 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < CF_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 2 * CF_CLUSTER); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 8);
    mbar_init(&norm_ready[0], 1);
    mbar_init(&norm_ready[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // every CTA's barriers exist before any peer multicasts into it / arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (a.pdl) {
    // programmatic dependent launch: the next grid may begin its prologue now; this grid must not touch global memory
    // before its predecessor has completed and flushed
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }

  unsigned long long* tr = a.trace ? a.trace + (size_t)blockIdx.x * 16 : nullptr;
  if (tr && threadIdx.x == 0) tr[0] = gtimer();
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: my row block, multicast to all
    if (lane == 0) {
      int kbg = 0;
      for (int img = cid; img < a.B; img += ncl) {
        for (int kb = 0; kb < nk; ++kb, ++kbg) {
          const int s = kbg % CF_STAGES;
          const uint32_t ph = (kbg / CF_STAGES) & 1;
          mbar_wait_cluster(&empty[s], ph ^ 1);            // all four CTAs are done with this stage
          mbar_expect_tx(&full[s], CF_STAGE_BYTES);        // my own copy of the four row blocks
          tma_load_3d_mcast(stages + s * CF_STAGE_BYTES + rank * CF_SLOT, &tmX, &full[s], kb * 32, (int)rank * 128, img, all);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: acc[128 x 512] = X_r . X^T
    const uint32_t idesc = make_idesc_tf32(128, 256, 0, 0);
    const uint64_t desc_tmpl = make_sdesc(0, 16, 1024);
    int kbg = 0, it = 0;
    for (int img = cid; img < a.B; img += ncl, ++it) {
      mbar_wait(acc_empty, (it & 1) ^ 1);                  // the epilogue has drained the previous image
      tc_fence_after();
      for (int kb = 0; kb < nk; ++kb, ++kbg) {
        const int s = kbg % CF_STAGES;
        const uint32_t ph = (kbg / CF_STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (tr && lane == 0 && kb == 0 && it < 2) tr[1 + 4 * it] = gtimer();
        const uint32_t s0 = smem_u32(stages + s * CF_STAGE_BYTES);
        const uint64_t da = desc_tmpl + ((s0 + rank * CF_SLOT) >> 4);
        const uint64_t db0 = desc_tmpl + (s0 >> 4), db1 = desc_tmpl + ((s0 + 2 * CF_SLOT) >> 4);
        const int krem = a.HW - kb * 32;
        const int ksteps = krem >= 32 ? 4 : (krem + 7) / 8;
        if (elect_one()) {
          for (int ks = 0; ks < ksteps && !(a.dbg & 4); ++ks) {
            const uint32_t accum = (kb | ks) ? 1u : 0u;
            umma_tf32_ss(tmem_base, da + ks * 2, db0 + ks * 2, idesc, accum);          // columns   0..255: rows 0..255 of X
            umma_tf32_ss(tmem_base + 256, da + ks * 2, db1 + ks * 2, idesc, accum);    // columns 256..511
          }
          umma_commit_mcast(&empty[s], all);
        }
        __syncwarp();
      }
      if (tr && lane == 0 && it < 2) tr[2 + 4 * it] = gtimer();
      if (elect_one()) umma_commit(acc_full);
      __syncwarp();
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------------ epilogue: 8 warps; warp -> TMEM lane quarter q,
    // column half h.  Accumulator element (row i of my block, column j) goes to Y[j][128 r + i]  (G is symmetric).
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    int it = 0;
    for (int img = cid; img < a.B; img += ncl, ++it) {
      mbar_wait(&norm_ready[it & 1], (it >> 1) & 1);
      const float inv_norm = inv_box[it & 1];
      mbar_wait(acc_full, it & 1);
      tc_fence_after();
      if (tr && threadIdx.x == 64 && it < 2) tr[3 + 4 * it] = gtimer();
      float* ybase = a.Y + (size_t)img * CF_C * CF_C + rank * 128 + q * 32 + lane;
#pragma unroll 1
      for (int c = 8 * h; c < 8 * h + 8; ++c) {
        float v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, v);
        tmem_ld_wait();
        float* y = ybase + (size_t)(c * 32) * CF_C;
        if (a.dbg & 2) {
          float keep = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) keep += tf32_round(sqrt_approx(fmaf(v[j], a.inv_hw, a.eps)) * inv_norm);
          if (keep == 123.456f) y[0] = keep;
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            y[(size_t)j * CF_C] = tf32_round(sqrt_approx(fmaf(v[j], a.inv_hw, a.eps)) * inv_norm);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      if (tr && threadIdx.x == 64 && it < 2) tr[4 + 4 * it] = gtimer();
    }
  } else {
    // ------------------------------------------------------------------ channel sums -> 1/||z||  (4 warps, 128 threads)
    // thread t reads row t of each of the four row blocks (same swizzle phase t&7); a quarter warp touches 8 consecutive
    // rows = 8 distinct 16-byte chunks: conflict-free.
    const int t = threadIdx.x - 320;
    const int w = warp - 10;
    int kbg = 0, it = 0;
    for (int img = cid; img < a.B; img += ncl, ++it) {
      float nsq = 0.f;
      for (int kb = 0; kb < nk; ++kb, ++kbg) {
        const int s = kbg % CF_STAGES;
        const uint32_t ph = (kbg / CF_STAGES) & 1;
        mbar_wait(&full[s], ph);
        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
        const uint8_t* row0 = stages + s * CF_STAGE_BYTES + t * 128;
#pragma unroll
        for (int blk = 0; blk < CF_CLUSTER && !(a.dbg & 1); ++blk) {
#pragma unroll
          for (int lc = 0; lc < 8; ++lc) {
            const float4 v = *reinterpret_cast<const float4*>(row0 + blk * CF_SLOT + ((lc ^ (t & 7)) << 4));
            acc[4 * lc + 0] += __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
            acc[4 * lc + 1] += __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
            acc[4 * lc + 2] += __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
            acc[4 * lc + 3] += __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
          }
        }
        // transposing butterfly: lane L ends with the sum over the warp's 32 threads of acc[L]  (31 shuffles)
#pragma unroll
        for (int off = 16, n = 16; off >= 1; off >>= 1, n >>= 1) {
          const bool upper = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < n; ++i) {
            const float send = upper ? acc[i] : acc[i + n];
            const float keep = upper ? acc[i + n] : acc[i];
            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
        part[((kbg & 1) * 4 + w) * 32 + lane] = acc[0];
        asm volatile("bar.sync 1, 128;" ::: "memory");       // partial sums visible; all 128 threads are done with the stage
        if (t == 0) {
#pragma unroll
          for (uint32_t r = 0; r < CF_CLUSTER; ++r) mbar_arrive_remote(&empty[s], r);
        }
        if (w == 0) {
          const float* pp = part + (kbg & 1) * 128 + lane;
          const float sp = (pp[0] + pp[32]) + (pp[64] + pp[96]);
          nsq = fmaf(sp, sp, nsq);
        }
      }
      if (w == 0) {
        nsq = warp_sum(nsq);
        if (lane == 0) {
          const float nrm = sqrtf(nsq * a.inv_hw + (float)CF_C * (float)CF_C * a.eps);
          const float inn = 1.f / fmaxf(nrm, 1e-12f);
          inv_box[it & 1] = inn;
          mbar_arrive(&norm_ready[it & 1]);
          if (tr && it < 2) tr[9 + it] = gtimer();
          if (rank == 0 && a.inv_norm) a.inv_norm[img] = inn;
        }
        __syncwarp();
      }
    }
  }
  if (tr && threadIdx.x == 0) tr[11] = gtimer();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // no peer is still multicasting into this CTA or arriving on its barriers
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static unsigned long long* g_k1_trace = nullptr;
void set_k1_trace(void* buf) { g_k1_trace = static_cast<unsigned long long*>(buf); }

// x [B,512,HW] -> y [B,512*512]; returns 0, <0 (argument) or >0 (cudaError_t); HK_ERR_UNSUPPORTED if clusters of four
// cannot be scheduled with this much shared memory (the caller then uses the two-kernel path).
// one-time host-side setup (function attribute + cluster occupancy query).  Kept out of the launch path so that the
// first launch may happen inside a CUDA-graph capture, where only stream work is legal.
static int g_max_clusters = -1;
int bcnn_cluster_prepare() {
  if (g_max_clusters >= 0) return 0;
  cudaError_t e = cudaFuncSetAttribute(bcnn_cluster_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CF_SMEM);
  if (e != cudaSuccess) return set_error((int)e, "cudaFuncSetAttribute(bcnn_cluster_fwd): %s", cudaGetErrorString(e));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(CF_THREADS);
  cfg.dynamicSmemBytes = CF_SMEM;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CF_CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int sms = 148, dev = 0, n = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cfg.gridDim = dim3((sms / CF_CLUSTER) * CF_CLUSTER);
  e = cudaOccupancyMaxActiveClusters(&n, bcnn_cluster_fwd_kernel, &cfg);
  if (e != cudaSuccess || n <= 0) {
    (void)cudaGetLastError();
    n = 0;
  }
  if (const char* v = getenv("HK_K1_CLUSTERS")) { const int f = atoi(v); if (f > 0 && f < n) n = f; }
  g_max_clusters = n;
  return 0;
}

int bcnn_cluster_fwd(const CUtensorMap& tmX, float* y, float* inv_norm, int B, int HW, float inv_hw, cudaStream_t stream,
                     bool allow_pdl) {
  int r = bcnn_cluster_prepare();
  if (r) return r;
  const int max_clusters = g_max_clusters;
  cudaError_t e;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(CF_THREADS);
  cfg.dynamicSmemBytes = CF_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CF_CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  if (max_clusters == 0) return set_error(HK_ERR_UNSUPPORTED, "bcnn_cluster_fwd: clusters of %d CTAs cannot be scheduled", CF_CLUSTER);
  static int pdl_env = -1;
  if (pdl_env < 0) { const char* v = getenv("HK_K1_PDL"); pdl_env = v ? atoi(v) : 1; }
  const int pdl = (pdl_env && allow_pdl) ? 1 : 0;
  CfArgs a = {};
  a.B = B; a.HW = HW; a.inv_hw = inv_hw; a.eps = 1e-5f; a.Y = y; a.inv_norm = inv_norm; a.pdl = pdl;
  { const char* v = getenv("HK_K1_DBG"); a.dbg = v ? atoi(v) : 0; }
  a.trace = g_k1_trace;
  const int ncl = B < max_clusters ? B : max_clusters;
  cfg.gridDim = dim3(ncl * CF_CLUSTER);
  cfg.numAttrs = pdl ? 2 : 1;
  e = cudaLaunchKernelEx(&cfg, bcnn_cluster_fwd_kernel, tmX, a);
  if (e != cudaSuccess) return set_error((int)e, "cudaLaunchKernelEx(bcnn_cluster_fwd): %s", cudaGetErrorString(e));
  HK_LAUNCH_CHECK("bcnn_cluster_fwd_kernel");
  return 0;
}

}  // namespace hk

/* profiling aid (not part of the ABI header): per-CTA %globaltimer stamps of the following K1 launches go to buf */
namespace hk { void set_super_trace(void* buf); }
extern "C" void hk_debug_k1_trace(void* buf) { hk::set_k1_trace(buf); hk::set_super_trace(buf); }
