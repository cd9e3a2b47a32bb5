// VGG-16 backbone convolutions (reference model/backbone/vgg.py:56-70: Conv2d 3x3 s1 p1 + bias, ReLU,
// MaxPool2d 2x2) as im2col-free implicit GEMMs on tcgen05 (kind::tf32, fp32 accumulate in TMEM).
//
// Layout: activations are NHWC fp32 inside the backbone.  For a 3x3 tap (kh,kw) the A operand of the
// implicit GEMM is the input window shifted by (kh-1,kw-1); a 4-D TMA box {32 ch, TW, TH, TN} at the
// shifted (possibly negative) coordinate lands as 128 rows x 128 B in 128B-swizzled shared memory —
// exactly the K-major UMMA layout — and TMA's out-of-bounds zero fill *is* the conv padding.
//   fwd   : Y[pix, co]  = sum_{tap,ci} X[pix+tap, ci] * Wf[tap][co][ci]        (+bias, ReLU)
//   dgrad : dX[pix, ci] = sum_{tap,co} dY[pix+tap, co] * Wd[tap][ci][co]       (Wd = flipped/transposed W; * (act>0))
//   wgrad : dW[tap][co][ci] = sum_pix dY[pix, co] * X[pix+tap, ci]             (both operands MN-major; split-K)
#include <stdlib.h>

#include "common.cuh"
#include "host.h"
#include "../../include/hawkeye_b200.h"

namespace hk {

// ------------------------------------------------------------------------------------------------
// weight packing:  W [Cout][Cin][3][3]  ->  Wf [9][Cout][Cin]  and  Wd [9][Cin][Cout] (taps flipped)
// ------------------------------------------------------------------------------------------------
__global__ void pack_weights_kernel(const float* __restrict__ W, float* __restrict__ Wf, float* __restrict__ Wd,
                                    int Cout, int Cin, int round) {
  const size_t n = (size_t)Cout * Cin * 9;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int t = i % 9;
    const int ci = (i / 9) % Cin;
    const int co = i / ((size_t)9 * Cin);
    const float v = round ? tf32_round(W[i]) : W[i];
    if (Wf) Wf[((size_t)t * Cout + co) * Cin + ci] = v;
    if (Wd) Wd[((size_t)(8 - t) * Cin + ci) * Cout + co] = v;
  }
}
// dWp [9][Cout][Cin] -> dW [Cout][Cin][3][3]
__global__ void unpack_wgrad_kernel(const float* __restrict__ dWp, float* __restrict__ dW, int Cout, int Cin, int accumulate) {
  const size_t n = (size_t)Cout * Cin * 9;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int t = i % 9;
    const int ci = (i / 9) % Cin;
    const int co = i / ((size_t)9 * Cin);
    const float v = dWp[((size_t)t * Cout + co) * Cin + ci];
    dW[i] = accumulate ? dW[i] + v : v;
  }
}

// ------------------------------------------------------------------------------------------------
// implicit-GEMM conv (forward and data-gradient)
// ------------------------------------------------------------------------------------------------
struct ConvArgs {
  float* Y;            // [N,H,W,Cout]
  const float* bias;   // [Cout] or null
  const float* mask;   // [N,H,W,Cout] or null: out *= (mask > 0)   (ReLU backward fused into dgrad)
  int N, H, W, Cin, Cout;
  int TW, TH, TN;      // pixel tile (product 128)
  int tiles_w, tiles_h, tiles_n;
  int relu;
  int stride;          // 1 or 2 (N,H,W above are OUTPUT dims; the input map is H*stride x W*stride)
  const float* addend; // [N,H,W,Cout] or null: raw partial sum added to the accumulator first (3xTF32 passes; may alias Y)
  int no_round;        // 1: store fp32 as is (precise mode); 0: round to tf32 (the output feeds another MMA)
  // fused MaxPool2d(2,2) epilogue (vgg.py:59): when P != null the full-resolution map Y is NOT written; the epilogue
  // reduces each 2x2 window across lanes (the pixel tile is a power-of-two patch, so the window partners are lane^1,
  // lane^TW, lane^(TW+1)) and writes the pooled value plus, if code != null, the byte maxpool2x2_bwd_idx reads
  // (bits 0-1 = first maximum in scan order, bit 2 = max > 0).
  float* P;            // [N,H/2,W/2,Cout] (or [N,Cout,H/2,W/2] when pool_nchw)
  unsigned char* code; // [N,H/2,W/2,Cout] or null
  int pool_nchw;
};

// 2x2 max-pool of one 32-channel chunk held as v[32] by the thread of pixel (w, h): window partners are lanes ^1, ^TW and
// ^(TW|1).  Written by the top-left lane of each window.  Bit-identical to maxpool2x2_fwd_idx_kernel on the stored map.
__device__ __forceinline__ void epi_pool_store(const ConvArgs& a, const float* v, int TW, int w, int h, int n, int co,
                                               bool valid) {
  const bool writer = valid && !(w & 1) && !(h & 1);
  const int Ho = a.H >> 1, Wo = a.W >> 1;
  const size_t pp = ((size_t)n * Ho + (h >> 1)) * Wo + (w >> 1);
  float m[32];
  unsigned char cd[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float v0 = v[j];
    const float v1 = __shfl_xor_sync(0xffffffffu, v0, 1);
    const float v2 = __shfl_xor_sync(0xffffffffu, v0, TW);
    const float v3 = __shfl_xor_sync(0xffffffffu, v0, TW | 1);
    float mm = v0;
    unsigned char k = 0;
    if (v1 > mm) { mm = v1; k = 1; }
    if (v2 > mm) { mm = v2; k = 2; }
    if (v3 > mm) { mm = v3; k = 3; }
    m[j] = mm;
    cd[j] = k | (mm > 0.f ? 4 : 0);
  }
  if (!writer) return;
  if (a.code) {
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      pk[j] = (uint32_t)cd[4 * j] | ((uint32_t)cd[4 * j + 1] << 8) | ((uint32_t)cd[4 * j + 2] << 16) | ((uint32_t)cd[4 * j + 3] << 24);
    uint4* cp = reinterpret_cast<uint4*>(a.code + pp * a.Cout + co);
    cp[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    cp[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
  if (!a.pool_nchw) {
    float4* dst = reinterpret_cast<float4*>(a.P + pp * a.Cout + co);
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = make_float4(m[4 * j], m[4 * j + 1], m[4 * j + 2], m[4 * j + 3]);
  } else {
    const size_t hw = (size_t)Ho * Wo;
    float* dst = a.P + ((size_t)n * a.Cout + co) * hw + (size_t)(h >> 1) * Wo + (w >> 1);
#pragma unroll
    for (int j = 0; j < 32; ++j) dst[(size_t)j * hw] = m[j];
  }
}

template <int BN>
struct ConvCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 3 : 4);
  static constexpr int MIN_CTAS = (BN == 256) ? 1 : 2;
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int BN>
__global__ void __launch_bounds__(192, ConvCfg<BN>::MIN_CTAS)
conv3x3_igemm_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, ConvArgs a) {
  using Cfg = ConvCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::STAGES * Cfg::A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + Cfg::STAGES;
  uint64_t* accf = empty + Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accf + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int t = blockIdx.x;
  const int tw = t % a.tiles_w; t /= a.tiles_w;
  const int th = t % a.tiles_h; t /= a.tiles_h;
  const int w0 = tw * a.TW, h0 = th * a.TH, n0 = t * a.TN;
  const int co0 = blockIdx.y * BN;
  const int nchunk = a.Cin / 32;
  const int nk = 9 * nchunk;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(accf, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, BN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % Cfg::STAGES;
        const uint32_t ph = (kb / Cfg::STAGES) & 1;
        const int tap = kb / nchunk, ck = kb - tap * nchunk;
        const int kh = tap / 3, kw = tap - kh * 3;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
        tma_load_4d(sA + s * Cfg::A_BYTES, &tmX, &full[s], ck * 32, w0 * a.stride + kw - 1, h0 * a.stride + kh - 1, n0);
        tma_load_3d(sB + s * Cfg::B_BYTES, &tmW, &full[s], ck * 32, co0, tap);
      }
    }
  } else if (warp == 1) {
    {   // warp-uniform loop (descriptor math on the uniform datapath); tcgen05 issue predicated on one elected lane
      const uint32_t idesc = make_idesc_tf32(128, BN, 0, 0);
      const uint64_t desc_tmpl = make_sdesc(0, 16, 1024);
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % Cfg::STAGES;
        const uint32_t ph = (kb / Cfg::STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint64_t a_base = desc_tmpl + (smem_u32(sA + s * Cfg::A_BYTES) >> 4);
        const uint64_t b_base = desc_tmpl + (smem_u32(sB + s * Cfg::B_BYTES) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_tf32_ss(tmem_base, a_base + ks * 2, b_base + ks * 2, idesc, (kb | ks) ? 1u : 0u);
          umma_commit(&empty[s]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(accf);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int wi = r % a.TW, hi = (r / a.TW) % a.TH, ni = r / (a.TW * a.TH);
    const int w = w0 + wi, h = h0 + hi, n = n0 + ni;
    const bool valid = (w < a.W) && (h < a.H) && (n < a.N);
    const size_t pix = ((size_t)n * a.H + h) * a.W + w;
    mbar_wait(accf, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      float v[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, v);
      tmem_ld_wait();
      const int co = co0 + c * 32;
      if (valid && co < a.Cout) {
        if (a.addend) {
          const float4* ad = reinterpret_cast<const float4*>(a.addend + pix * a.Cout + co);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t = ad[j];
            v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
          }
        }
        if (a.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __ldg(a.bias + co + j);
        }
        if (a.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (a.mask) {
          const float4* m = reinterpret_cast<const float4*>(a.mask + pix * a.Cout + co);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 mm = m[j];
            v[4 * j] = mm.x > 0.f ? v[4 * j] : 0.f;
            v[4 * j + 1] = mm.y > 0.f ? v[4 * j + 1] : 0.f;
            v[4 * j + 2] = mm.z > 0.f ? v[4 * j + 2] : 0.f;
            v[4 * j + 3] = mm.w > 0.f ? v[4 * j + 3] : 0.f;
          }
        }
        if (a.P) {       // (never combined with the precise-mode passes: rounded like the stored map would be)
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = tf32_round(v[j]);
        } else {
        float4* dst = reinterpret_cast<float4*>(a.Y + pix * a.Cout + co);
        if (a.no_round) {
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j] = make_float4(tf32_round(v[4 * j]), tf32_round(v[4 * j + 1]), tf32_round(v[4 * j + 2]),
                                 tf32_round(v[4 * j + 3]));
        }
        }
      }
      // the window reduction is a warp-wide shuffle: every lane takes part, whether or not its own pixel is valid
      if (a.P && co < a.Cout) epi_pool_store(a, v, a.TW, w, h, n, co, valid);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BN);
}

// pick a pixel tile TW x TH x TN with product `target` that tiles W x H (x N) with as little waste as possible
static void pick_tile(int W, int H, int N, int target, int* TW, int* TH, int* TN) {
  int tw = 1;
  while (tw * 2 <= 16 && W % (tw * 2) == 0 && tw * 2 <= target) tw *= 2;
  int th = 1;
  while (th * 2 * tw <= target && H % (th * 2) == 0) th *= 2;
  int tn = target / (tw * th);
  // if the map is tiny (e.g. 2x2) the remainder goes to the batch dimension
  *TW = tw; *TH = th; *TN = tn;
  (void)N;
}

static int make_act_map(CUtensorMap* tm, const float* X, int N, int H, int W, int C, int TW, int TH, int TN,
                        bool mn_major = false, int stride = 1) {
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
  uint64_t strides[3] = {(uint64_t)C * 4, (uint64_t)W * C * 4, (uint64_t)H * W * C * 4};
  // strided traversal: the box spans TW*stride elements and TMA keeps every stride-th one
  uint32_t box[4] = {32, (uint32_t)(TW * stride), (uint32_t)(TH * stride), (uint32_t)TN};
  uint32_t estr[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
  return make_tmap(tm, X, 4, dims, strides, box, mn_major, stride > 1 ? estr : nullptr);
}

template <int BN>
static int launch_conv(const CUtensorMap& tmX, const CUtensorMap& tmW, const ConvArgs& a, cudaStream_t stream) {
  using Cfg = ConvCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_igemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) return set_error((int)e, "cudaFuncSetAttribute(conv<%d>): %s", BN, cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid(a.tiles_w * a.tiles_h * a.tiles_n, (a.Cout + BN - 1) / BN);
  conv3x3_igemm_kernel<BN><<<grid, 192, Cfg::SMEM, stream>>>(tmX, tmW, a);
  HK_LAUNCH_CHECK("conv3x3_igemm_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// implicit-GEMM conv v2 (stride 1, maps with W % 16 == 0 and H % 8 == 0): persistent CTAs, halo reuse, double-buffered TMEM.
//   * pixel tile 16 x 8 of one image (M = 128).  For each (cin chunk, kw) ONE TMA load brings the (8+2) x 16 halo patch
//     (160 rows x 128 B); the three kh taps are the same patch addressed kh*16 rows (2 KB = whole swizzle atoms) further
//     down, so the input crosses the L2->SM fabric 3x (+25 % halo) instead of 9x.
//   * RESIDENT (Cin = 64, BN = 64: VGG conv1_2 and its dgrad): all 9x2 weight tiles (147 KB) stay in shared memory for the
//     life of the CTA; only activations stream.
//   * the epilogue of tile i (4 warps, TMEM set i&1) overlaps the TMA/MMA of tile i+1.
// ------------------------------------------------------------------------------------------------
template <int BN, bool RESIDENT>
struct ConvV2Cfg {
  static constexpr int A_BYTES = 160 * 128;                          // 20 KB halo patch
  static constexpr int B_TILE = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + (RESIDENT ? 0 : 3 * B_TILE);
  static constexpr int STAGES = RESIDENT ? 3 : (BN == 64 ? 4 : 3);
  static constexpr int WRES_BYTES = RESIDENT ? 18 * B_TILE : 0;      // 9 taps x 2 chunks
  static constexpr int EPI_TILE = 4 * 32 * 36 * 4;                   // four epilogue warps x a 32 x 36-float transpose tile
  static constexpr int SMEM = STAGES * STAGE_BYTES + WRES_BYTES + 1024 + 512 + EPI_TILE;
};

template <int BN, bool RESIDENT>
__global__ void __launch_bounds__(192, 1)
conv3x3_igemm_v2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, ConvArgs a,
                        int n_ntiles, int total_tiles) {
  using Cfg = ConvV2Cfg<BN, RESIDENT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stages = smem;
  uint8_t* wres = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(wres + Cfg::WRES_BYTES);
  uint64_t* empty = full + Cfg::STAGES;
  uint64_t* acc_full = empty + Cfg::STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* wbar = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
  float* epi_tiles = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + 512);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunk = a.Cin / 32;
  const int nkb = nchunk * 3;                       // (chunk, kw) steps per tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    mbar_init(wbar, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 2 * BN < 32 ? 32 : 2 * BN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      if (RESIDENT) {   // whole filter bank of this N tile (n_ntiles == 1 by construction): 18 tiles of [BN x 32]
        mbar_expect_tx(wbar, Cfg::WRES_BYTES);
        for (int tap = 0; tap < 9; ++tap)
          for (int ck = 0; ck < 2; ++ck) tma_load_3d(wres + (tap * 2 + ck) * Cfg::B_TILE, &tmW, wbar, ck * 32, 0, tap);
      }
      int kbg = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int nt = t % n_ntiles;
        int pt = t / n_ntiles;
        const int tw = pt % a.tiles_w; pt /= a.tiles_w;
        const int th = pt % a.tiles_h; pt /= a.tiles_h;
        const int w0 = tw * 16, h0 = th * 8, n0 = pt, co0 = nt * BN;
        for (int kb = 0; kb < nkb; ++kb, ++kbg) {
          const int s = kbg % Cfg::STAGES;
          const uint32_t ph = (kbg / Cfg::STAGES) & 1;
          const int ck = kb / 3, kw = kb - ck * 3;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
          uint8_t* st = stages + s * Cfg::STAGE_BYTES;
          tma_load_4d(st, &tmX, &full[s], ck * 32, w0 + kw - 1, h0 - 1, n0);
          if (!RESIDENT) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
              tma_load_3d(st + Cfg::A_BYTES + kh * Cfg::B_TILE, &tmW, &full[s], ck * 32, co0, kh * 3 + kw);
          }
        }
      }
    }
  } else if (warp == 1) {
    {   // warp-uniform loop; tcgen05 issue predicated on one elected lane
      const uint32_t idesc = make_idesc_tf32(128, BN, 0, 0);
      const uint64_t desc_tmpl = make_sdesc(0, 16, 1024);
      if (RESIDENT) { mbar_wait(wbar, 0); tc_fence_after(); }
      int kbg = 0, itl = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++itl) {
        const int set = itl & 1;
        mbar_wait(&acc_empty[set], ((itl >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + set * BN;
        for (int kb = 0; kb < nkb; ++kb, ++kbg) {
          const int s = kbg % Cfg::STAGES;
          const uint32_t ph = (kbg / Cfg::STAGES) & 1;
          const int ck = kb / 3, kw = kb - ck * 3;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          // descriptors = constant template + (address >> 4); per-MMA work is one 64-bit add per operand
          const uint32_t a_addr = smem_u32(stages + s * Cfg::STAGE_BYTES);
          const uint64_t a_base = desc_tmpl + (a_addr >> 4);
          uint64_t b_base[3];
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const uint32_t b_addr = RESIDENT ? smem_u32(wres + ((kh * 3 + kw) * 2 + ck) * Cfg::B_TILE)
                                             : a_addr + Cfg::A_BYTES + kh * Cfg::B_TILE;
            b_base[kh] = desc_tmpl + (b_addr >> 4);
          }
          if (elect_one()) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_tf32_ss(d, a_base + (kh * 128 + ks * 2), b_base[kh] + ks * 2, idesc, (kb | kh | ks) ? 1u : 0u);
            umma_commit(&empty[s]);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(&acc_full[set]);
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int wi = r & 15, hi = r >> 4;
    int itl = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++itl) {
      const int nt = t % n_ntiles;
      int pt = t / n_ntiles;
      const int tw = pt % a.tiles_w; pt /= a.tiles_w;
      const int th = pt % a.tiles_h; pt /= a.tiles_h;
      const int w = tw * 16 + wi, h = th * 8 + hi, n = pt, co0 = nt * BN;
      const size_t pix = ((size_t)n * a.H + h) * a.W + w;
      const int set = itl & 1;
      mbar_wait(&acc_full[set], (itl >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + set * BN + c * 32, v);
        tmem_ld_wait();
        const int co = co0 + c * 32;
        if (a.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __ldg(a.bias + co + j);
        }
        if (a.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (a.mask) {
          const float4* m = reinterpret_cast<const float4*>(a.mask + pix * a.Cout + co);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 mm = __ldg(m + j);
            v[4 * j] = mm.x > 0.f ? v[4 * j] : 0.f;
            v[4 * j + 1] = mm.y > 0.f ? v[4 * j + 1] : 0.f;
            v[4 * j + 2] = mm.z > 0.f ? v[4 * j + 2] : 0.f;
            v[4 * j + 3] = mm.w > 0.f ? v[4 * j + 3] : 0.f;
          }
        }
        if (a.P) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = tf32_round(v[j]);
          epi_pool_store(a, v, 16, w, h, n, co, true);
        } else {
          // thread = pixel owning 32 consecutive channels: a direct float4 store instruction would touch 32 pixels x 16 B
          // (half-written sectors).  Transposed through a warp-private tile, every store instruction writes the full 128 B
          // of four pixels.
          float* tile = epi_tiles + q * (32 * 36);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(tile + lane * 36 + 4 * j) =
                make_float4(tf32_round(v[4 * j]), tf32_round(v[4 * j + 1]), tf32_round(v[4 * j + 2]), tf32_round(v[4 * j + 3]));
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + (lane >> 3);                      // pixel (lane) rr of this warp: wi = rr & 15, hi = 2 q + (rr >> 4)
            const size_t pr = ((size_t)n * a.H + (th * 8 + q * 2 + (rr >> 4))) * a.W + tw * 16 + (rr & 15);
            *reinterpret_cast<float4*>(a.Y + pr * a.Cout + co + (lane & 7) * 4) =
                *reinterpret_cast<const float4*>(tile + rr * 36 + (lane & 7) * 4);
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[set]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * BN < 32 ? 32 : 2 * BN);
}

template <int BN, bool RESIDENT>
static int launch_conv_v2(const float* x, const float* wp, ConvArgs a, int N, int H, int W, cudaStream_t stream) {
  using Cfg = ConvV2Cfg<BN, RESIDENT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_igemm_v2_kernel<BN, RESIDENT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM);
    if (e != cudaSuccess) return set_error((int)e, "cudaFuncSetAttribute(conv_v2<%d>): %s", BN, cudaGetErrorString(e));
    attr_set = true;
  }
  a.TW = 16; a.TH = 8; a.TN = 1;
  a.tiles_w = W / 16; a.tiles_h = H / 8; a.tiles_n = N;
  const int n_ntiles = (a.Cout + BN - 1) / BN;
  const long long total = (long long)a.tiles_w * a.tiles_h * a.tiles_n * n_ntiles;
  HK_REQUIRE(total < (1ll << 31), HK_ERR_UNSUPPORTED, "conv3x3: too many tiles");
  CUtensorMap tmX, tmW;
  int r;
  {
    uint64_t dims[4] = {(uint64_t)a.Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t strides[3] = {(uint64_t)a.Cin * 4, (uint64_t)W * a.Cin * 4, (uint64_t)H * W * a.Cin * 4};
    uint32_t box[4] = {32, 16, 10, 1};
    if ((r = make_tmap(&tmX, x, 4, dims, strides, box))) return r;
  }
  {
    uint64_t dims[3] = {(uint64_t)a.Cin, (uint64_t)a.Cout, 9};
    uint64_t strides[2] = {(uint64_t)a.Cin * 4, (uint64_t)a.Cin * a.Cout * 4};
    uint32_t box[3] = {32, (uint32_t)BN, 1};
    if ((r = make_tmap(&tmW, wp, 3, dims, strides, box))) return r;
  }
  int sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int grid = total < sms ? (int)total : sms;
  conv3x3_igemm_v2_kernel<BN, RESIDENT><<<grid, 192, Cfg::SMEM, stream>>>(tmX, tmW, a, n_ntiles, (int)total);
  HK_LAUNCH_CHECK("conv3x3_igemm_v2_kernel");
  return 0;
}

static int conv3x3_igemm_1x(const float* x, const float* wp, const float* bias, const float* mask, const float* addend,
                            float* y, int N, int H, int W, int Cin, int Cout, int relu, cudaStream_t stream, int stride,
                            bool generic_only, int no_round, float* pooled = nullptr, unsigned char* code = nullptr,
                            int pool_nchw = 0);

// x NHWC [N,H,W,Cin], wp packed [9][Cout][Cin] -> y NHWC [N,H,W,Cout]
int conv3x3_igemm(const float* x, const float* wp, const float* bias, const float* mask, float* y, int N, int H, int W,
                  int Cin, int Cout, int relu, cudaStream_t stream, int stride = 1) {
  if (!precise()) return conv3x3_igemm_1x(x, wp, bias, mask, nullptr, y, N, H, W, Cin, Cout, relu, stream, stride, false, 0);
  // 3xTF32: y = epi(Xh*Wh + Xl*Wh + Xh*Wl), three passes of the same implicit-GEMM kernel chained through `addend`
  HK_REQUIRE(x && wp && y, HK_ERR_ARG, "conv3x3: null pointer");
  const size_t nx = (size_t)N * H * W * Cin, nw = (size_t)9 * Cout * Cin;
  Scratch sx(2 * nx * sizeof(float), stream), sw(2 * nw * sizeof(float), stream);
  HK_REQUIRE(sx.p && sw.p, HK_ERR_DRIVER, "conv3x3 (precise): cudaMallocAsync of the operand halves failed");
  float *xh = sx.f(), *xl = xh + nx, *wh = sw.f(), *wl = wh + nw;
  int r;
  if ((r = tf32_split(x, xh, xl, nx, stream))) return r;
  if ((r = tf32_split(wp, wh, wl, nw, stream))) return r;
  if ((r = conv3x3_igemm_1x(xh, wl, nullptr, nullptr, nullptr, y, N, H, W, Cin, Cout, 0, stream, stride, true, 1))) return r;
  if ((r = conv3x3_igemm_1x(xl, wh, nullptr, nullptr, y, y, N, H, W, Cin, Cout, 0, stream, stride, true, 1))) return r;
  return conv3x3_igemm_1x(xh, wh, bias, mask, y, y, N, H, W, Cin, Cout, relu, stream, stride, true, 1);
}

static int conv3x3_igemm_1x(const float* x, const float* wp, const float* bias, const float* mask, const float* addend,
                            float* y, int N, int H, int W, int Cin, int Cout, int relu, cudaStream_t stream, int stride,
                            bool generic_only, int no_round, float* pooled, unsigned char* code, int pool_nchw) {
  // H, W are the INPUT dims; output is H/stride x W/stride (padding 1)
  const int Hin = H, Win = W;
  if (stride == 2) {
    HK_REQUIRE(H % 2 == 0 && W % 2 == 0, HK_ERR_UNSUPPORTED, "conv3x3 stride 2: even H/W required");
    H /= 2; W /= 2;
  }
  HK_REQUIRE(x && wp && (y || pooled), HK_ERR_ARG, "conv3x3: null pointer");
  HK_REQUIRE(Cin % 32 == 0 && Cout % 32 == 0, HK_ERR_UNSUPPORTED, "conv3x3: Cin=%d Cout=%d must be multiples of 32",
             Cin, Cout);
  HK_REQUIRE(aligned16(x) && aligned16(wp) && (!y || aligned16(y)) && (!mask || aligned16(mask)), HK_ERR_ALIGN,
             "conv3x3: pointer not 16-byte aligned");
  if (pooled)
    HK_REQUIRE(stride == 1 && H % 2 == 0 && W % 2 == 0 && aligned16(pooled) && (!code || aligned16(code)), HK_ERR_UNSUPPORTED,
               "conv3x3 + pool: even H/W, stride 1 and 16-byte aligned outputs required");
  ConvArgs a = {};
  a.P = pooled; a.code = code; a.pool_nchw = pool_nchw;
  a.Y = y; a.bias = bias; a.mask = mask; a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.relu = relu;
  a.stride = stride;
  a.addend = addend; a.no_round = no_round;
  if (!generic_only) {
    static int use_v2 = -1;
    if (use_v2 < 0) { const char* v = getenv("HK_CONV_V2"); use_v2 = v ? atoi(v) : 1; }
    if (use_v2 && stride == 1 && W % 16 == 0 && H % 8 == 0) {
      if (Cin == 64 && Cout == 64) return launch_conv_v2<64, true>(x, wp, a, N, H, W, stream);
      if (Cout <= 64) return launch_conv_v2<64, false>(x, wp, a, N, H, W, stream);
      return launch_conv_v2<128, false>(x, wp, a, N, H, W, stream);
    }
  }
  pick_tile(W, H, N, 128, &a.TW, &a.TH, &a.TN);
  HK_REQUIRE(!pooled || (a.TW >= 2 && a.TH >= 2), HK_ERR_UNSUPPORTED, "conv3x3 + pool: pixel tile %dx%d", a.TW, a.TH);
  a.tiles_w = (W + a.TW - 1) / a.TW; a.tiles_h = (H + a.TH - 1) / a.TH; a.tiles_n = (N + a.TN - 1) / a.TN;
  HK_REQUIRE((long long)a.tiles_w * a.tiles_h * a.tiles_n < (1ll << 31), HK_ERR_UNSUPPORTED, "conv3x3: grid too large");
  CUtensorMap tmX, tmW;
  int r;
  if ((r = make_act_map(&tmX, x, N, Hin, Win, Cin, a.TW, a.TH, a.TN, false, stride))) return r;
  const int BN = Cout <= 64 ? 64 : (Cout <= 128 ? 128 : 256);
  {
    uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, 9};
    uint64_t strides[2] = {(uint64_t)Cin * 4, (uint64_t)Cin * Cout * 4};
    uint32_t box[3] = {32, (uint32_t)BN, 1};
    if ((r = make_tmap(&tmW, wp, 3, dims, strides, box))) return r;
  }
  if (BN == 64) return launch_conv<64>(tmX, tmW, a, stream);
  if (BN == 128) return launch_conv<128>(tmX, tmW, a, stream);
  return launch_conv<256>(tmX, tmW, a, stream);
}

// ------------------------------------------------------------------------------------------------
// weight gradient: one CTA accumulates dWp[tap][co0:co0+128][ci0:ci0+32] for ALL 9 taps over its share of the
// pixels (split-K across CTAs, fp32 atomics at the end).
//   A = dY tile (M = 128 co, MN-major, k = 64 pixels)  — loaded once per pixel tile and reused by the 9 taps;
//   B = X halo patch, one TMA load per kw shift: (TH+2) x TW pixels x 32 ci; the three kh taps are the SAME
//       shared-memory patch addressed kh*TW rows further down (a whole number of 512 B swizzle atoms), so the
//       input is fetched 3x (+halo) instead of 9x;
//   D = 9 accumulators of 128 x 32 fp32 in TMEM (288 columns) + one 128 x 16 accumulator against an all-ones
//       B tile, whose every column is sum_pix dY[pix,co] = the bias gradient (no separate pass over dY).
// ------------------------------------------------------------------------------------------------
struct WgradArgs {
  float* dWp;  // [9][Cout][Cin], pre-zeroed
  float* db;   // [Cout], pre-zeroed, or null
  int N, H, W, Cin, Cout;
  int TW, TH, TN, tiles_w, tiles_h, tiles_n;
  int ksplit;
  int dbg;
};

constexpr int WG_KP = 64;                       // pixels per stage
constexpr int WG_STAGES = 3;
constexpr int WG_A_BYTES = 4 * WG_KP * 128;     // 4 co-blocks of [64 px x 128 B]            = 32 KB
constexpr int WG_B_ONE = 96 * 128;              // one kw patch: (TH+2)*TW*TN = 96 rows      = 12 KB
constexpr int WG_B_BYTES = 3 * WG_B_ONE;
constexpr int WG_STAGE_BYTES = WG_A_BYTES + WG_B_BYTES;   // 68 KB
constexpr int WG_ONES_BYTES = 8 * 128;          // all-ones B tile: 8 k-rows x 128 B (reused for every k-step)
constexpr int WG_BIAS_COL = 448;                 // TMEM column of the 128 x 16 bias-gradient accumulator (taps use 0..287)
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + WG_ONES_BYTES + 1024 + 256;

__global__ void __launch_bounds__(192, 1)
conv3x3_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, WgradArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + WG_STAGES * WG_A_BYTES;
  float* ones = reinterpret_cast<float*>(smem + WG_STAGES * WG_STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES + WG_ONES_BYTES);
  uint64_t* empty = full + WG_STAGES;
  uint64_t* accf = empty + WG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accf + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_ci_tiles = a.Cin / 32;
  const int ci_t = blockIdx.x % n_ci_tiles, co_t = blockIdx.x / n_ci_tiles;
  const int split = blockIdx.y;
  const int co0 = co_t * 128, ci0 = ci_t * 32;
  const bool do_bias = (a.db != nullptr) && (ci_t == 0);
  const long long total_tiles = (long long)a.tiles_w * a.tiles_h * a.tiles_n;
  const long long per = (total_tiles + a.ksplit - 1) / a.ksplit;
  const long long t_begin = per * split;
  const long long t_end = (t_begin + per < total_tiles) ? t_begin + per : total_tiles;
  const int nk = (int)(t_end > t_begin ? t_end - t_begin : 0);
  const int img_rows = a.TH * a.TW;                 // A rows per image in the tile
  const int patch_rows = (a.TH + 2) * a.TW;         // B rows per image in the patch
  const int ksteps_img = img_rows / 8;

  for (int i = threadIdx.x; i < WG_ONES_BYTES / 4; i += blockDim.x) ones[i] = 1.f;
  fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(accf, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (nk > 0) {
    if (warp == 0) {
      if (lane == 0) {
        for (int kb = 0; kb < nk; ++kb) {
          const int s = kb % WG_STAGES;
          const uint32_t ph = (kb / WG_STAGES) & 1;
          int tt = (int)t_begin + kb;
          const int tw = tt % a.tiles_w; tt /= a.tiles_w;
          const int th = tt % a.tiles_h; tt /= a.tiles_h;
          const int w0 = tw * a.TW, h0 = th * a.TH, n0 = tt * a.TN;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], WG_A_BYTES + 3 * patch_rows * a.TN * 128);
          uint8_t* pa = sA + s * WG_A_BYTES;
          uint8_t* pb = sB + s * WG_B_BYTES;
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_4d(pa + j * WG_KP * 128, &tmDY, &full[s], co0 + j * 32, w0, h0, n0);
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) tma_load_4d(pb + kw * WG_B_ONE, &tmX, &full[s], ci0, w0 + kw - 1, h0 - 1, n0);
        }
      }
    } else if (warp == 1) {
      {
        const uint32_t idesc = make_idesc_tf32(128, 96, 1, 1);
        const uint32_t idesc_b = make_idesc_tf32(128, 16, 1, 1);
        const uint64_t ones_desc = make_sdesc_mn(smem_u32(ones), 0);
        // The issuing thread is the bottleneck of this kernel (3 N=96 MMAs of 48 cycles per k-step).  The loop runs
        // warp-uniformly (all 32 lanes) so the descriptor arithmetic is done on the uniform datapath; per-k-step offsets
        // (16-byte units) are computed once; only the tcgen05 issue is predicated on one elected lane.
        uint32_t a_off[8], b_off[8];
        {
          int ks = 0;
          for (int n = 0; n < a.TN; ++n)
            for (int j = 0; j < ksteps_img; ++j, ++ks) {
              a_off[ks] = (uint32_t)((n * img_rows + j * 8) * 128) >> 4;
              b_off[ks] = (uint32_t)((n * patch_rows + j * 8) * 128) >> 4;
            }
        }
        const uint32_t kh_step = (uint32_t)(a.TW * 128) >> 4;
        const uint64_t a_tmpl = make_sdesc_mn(0, WG_KP * 128);
        const uint64_t b_tmpl = make_sdesc_mn(0, WG_B_ONE);
        for (int kb = 0; kb < nk; ++kb) {
          const int s = kb % WG_STAGES;
          const uint32_t ph = (kb / WG_STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint64_t a_base = a_tmpl + (smem_u32(sA + s * WG_A_BYTES) >> 4);
          const uint64_t b_base = b_tmpl + (smem_u32(sB + s * WG_B_BYTES) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint32_t accum = (kb | ks) ? 1u : 0u;
              const uint64_t ad = a_base + a_off[ks];
              const uint64_t bd = b_base + b_off[ks];
              umma_tf32_ss(tmem_base, ad, bd, idesc, accum);
              umma_tf32_ss(tmem_base + 96, ad, bd + kh_step, idesc, accum);
              umma_tf32_ss(tmem_base + 192, ad, bd + 2 * kh_step, idesc, accum);
              if (do_bias) umma_tf32_ss(tmem_base + WG_BIAS_COL, ad, ones_desc, idesc_b, accum);
            }
            umma_commit(&empty[s]);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(accf);
        __syncwarp();
      }
    } else {
      const int q = warp & 3;
      const int co = co0 + q * 32 + lane;
      mbar_wait(accf, 0);
      tc_fence_after();
#pragma unroll 1
      for (int tap = 0; tap < 9; ++tap) {
        float v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + tap * 32, v);
        tmem_ld_wait();
        if (co < a.Cout) {
          float* dst = a.dWp + ((size_t)tap * a.Cout + co) * a.Cin + ci0;
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(dst + j, v[j]);
        }
      }
      if (do_bias) {
        float v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + WG_BIAS_COL, v);
        tmem_ld_wait();
        if (co < a.Cout) atomicAdd(a.db + co, v[0]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// wgrad, version 2: the dY tile (A operand, 128 co x 64 pixels) is reused by all nine taps, yet in the SS form every one of
// the 3 (+1 bias) MMAs of a k-step re-reads it from shared memory: (4 KB A + 3 KB B) per 48-cycle N=96 MMA = 146 B/clk
// against a 128 B/clk shared-memory port — the kernel was shared-memory-bound, not tensor-bound.  Here the four otherwise
// idle epilogue warps transpose each dY tile ONCE from shared memory into tensor memory (thread = co, 64 pixel columns,
// tcgen05.st) and the MMAs take A from TMEM (tcgen05.mma [tmem], b-desc): shared-memory traffic per k-step drops from 21 KB
// to 9 KB + a 4 KB one-off read.  dY is TMA-loaded with the plain 128B swizzle (it is no longer a UMMA smem operand).
// TMEM: tap accumulators 0..287, bias 288..303, A double buffer 320..447.
// ------------------------------------------------------------------------------------------------
constexpr int WG2_BIAS_COL = 288;
constexpr int WG2_A_COL = 320;

__global__ void __launch_bounds__(192, 1)
conv3x3_wgrad_v2_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, WgradArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + WG_STAGES * WG_A_BYTES;
  float* ones = reinterpret_cast<float*>(smem + WG_STAGES * WG_STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES + WG_ONES_BYTES);
  uint64_t* empty = full + WG_STAGES;
  uint64_t* accf = empty + WG_STAGES;
  uint64_t* a_ready = accf + 1;     // [2] stager warps (4) -> MMA: dY tile kb is in TMEM buffer kb & 1
  uint64_t* a_free = a_ready + 2;   // [2] MMA (tcgen05.commit) -> stagers: the MMAs reading buffer kb & 1 have retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_ci_tiles = a.Cin / 32;
  const int ci_t = blockIdx.x % n_ci_tiles, co_t = blockIdx.x / n_ci_tiles;
  const int split = blockIdx.y;
  const int co0 = co_t * 128, ci0 = ci_t * 32;
  const bool do_bias = (a.db != nullptr) && (ci_t == 0);
  const long long total_tiles = (long long)a.tiles_w * a.tiles_h * a.tiles_n;
  const long long per = (total_tiles + a.ksplit - 1) / a.ksplit;
  const long long t_begin = per * split;
  const long long t_end = (t_begin + per < total_tiles) ? t_begin + per : total_tiles;
  const int nk = (int)(t_end > t_begin ? t_end - t_begin : 0);
  const int img_rows = a.TH * a.TW;
  const int patch_rows = (a.TH + 2) * a.TW;
  const int ksteps_img = img_rows / 8;

  for (int i = threadIdx.x; i < WG_ONES_BYTES / 4; i += blockDim.x) ones[i] = 1.f;
  fence_proxy_async();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(accf, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&a_ready[i], 4); mbar_init(&a_free[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (nk > 0) {
    if (warp == 0) {
      if (lane == 0) {
        for (int kb = 0; kb < nk; ++kb) {
          const int s = kb % WG_STAGES;
          const uint32_t ph = (kb / WG_STAGES) & 1;
          int tt = (int)t_begin + kb;
          const int tw = tt % a.tiles_w; tt /= a.tiles_w;
          const int th = tt % a.tiles_h; tt /= a.tiles_h;
          const int w0 = tw * a.TW, h0 = th * a.TH, n0 = tt * a.TN;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], WG_A_BYTES + 3 * patch_rows * a.TN * 128);
          uint8_t* pa = sA + s * WG_A_BYTES;
          uint8_t* pb = sB + s * WG_B_BYTES;
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_4d(pa + j * WG_KP * 128, &tmDY, &full[s], co0 + j * 32, w0, h0, n0);
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) tma_load_4d(pb + kw * WG_B_ONE, &tmX, &full[s], ci0, w0 + kw - 1, h0 - 1, n0);
        }
      }
    } else if (warp == 1) {
      const uint32_t idesc = make_idesc_tf32(128, 96, 0, 1);
      const uint32_t idesc_b = make_idesc_tf32(128, 16, 0, 1);
      const uint64_t ones_desc = make_sdesc_mn(smem_u32(ones), 0);
      uint32_t b_off[8];
      {
        int ks = 0;
        for (int n = 0; n < a.TN; ++n)
          for (int j = 0; j < ksteps_img; ++j, ++ks) b_off[ks] = (uint32_t)((n * patch_rows + j * 8) * 128) >> 4;
      }
      const uint32_t kh_step = (uint32_t)(a.TW * 128) >> 4;
      const uint64_t b_tmpl = make_sdesc_mn(0, WG_B_ONE);
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % WG_STAGES;
        const uint32_t ph = (kb / WG_STAGES) & 1;
        const int ab = kb & 1;
        mbar_wait(&full[s], ph);                       // the X patches of this stage have landed
        mbar_wait(&a_ready[ab], (kb >> 1) & 1);        // ... and its dY tile is in tensor memory
        tc_fence_after();
        const uint64_t b_base = b_tmpl + (smem_u32(sB + s * WG_B_BYTES) >> 4);
        const uint32_t a_tm = tmem_base + WG2_A_COL + ab * WG_KP;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t accum = (kb | ks) ? 1u : 0u;
            const uint32_t ad = a_tm + ks * 8;
            const uint64_t bd = b_base + b_off[ks];
            umma_tf32_ts(tmem_base, ad, bd, idesc, accum);
            umma_tf32_ts(tmem_base + 96, ad, bd + kh_step, idesc, accum);
            umma_tf32_ts(tmem_base + 192, ad, bd + 2 * kh_step, idesc, accum);
            if (do_bias) umma_tf32_ts(tmem_base + WG2_BIAS_COL, ad, ones_desc, idesc_b, accum);
          }
          umma_commit(&empty[s]);
          umma_commit(&a_free[ab]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(accf);
      __syncwarp();
    } else {
      // ---- stagers (main loop), then epilogue.  warp -> TMEM lane quarter q = co block of 32; thread = one co
      const int q = warp & 3;
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % WG_STAGES;
        const uint32_t ph = (kb / WG_STAGES) & 1;
        const int ab = kb & 1;
        mbar_wait(&full[s], ph);
        // co block q of the tile: 64 pixel rows x 128 B, 128B-swizzled (16-byte chunk ^ (row & 7)); lane = word in the row
        const uint8_t* tile = sA + s * WG_A_BYTES + q * (WG_KP * 128);
        float v0[32], v1[32];
#pragma unroll
        for (int k = 0; k < 32; ++k)
          v0[k] = *reinterpret_cast<const float*>(tile + k * 128 + ((((lane >> 2) ^ (k & 7)) << 4) | ((lane & 3) << 2)));
#pragma unroll
        for (int k = 0; k < 32; ++k)
          v1[k] = *reinterpret_cast<const float*>(tile + (32 + k) * 128 + ((((lane >> 2) ^ (k & 7)) << 4) | ((lane & 3) << 2)));
        if (kb >= 2) { mbar_wait(&a_free[ab], ((kb >> 1) - 1) & 1); tc_fence_after(); }
        const uint32_t dst = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + WG2_A_COL + ab * WG_KP;
        tmem_st32(dst, v0);
        tmem_st32(dst + 32, v1);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[ab]);
      }
      const int co = co0 + q * 32 + lane;
      mbar_wait(accf, 0);
      tc_fence_after();
#pragma unroll 1
      for (int tap = 0; tap < 9; ++tap) {
        float v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + tap * 32, v);
        tmem_ld_wait();
        if (co < a.Cout) {
          float* dst = a.dWp + ((size_t)tap * a.Cout + co) * a.Cin + ci0;
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(dst + j, v[j]);
        }
      }
      if (do_bias) {
        float v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + WG2_BIAS_COL, v);
        tmem_ld_wait();
        if (co < a.Cout) atomicAdd(a.db + co, v[0]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// pixel tile for wgrad: TW in {4,8,16} (kh shifts must be whole 512 B swizzle atoms), TW*TH*TN = 64, TH*TW % 8 == 0
static bool pick_wgrad_tile(int W, int H, int* TW, int* TH, int* TN) {
  int tw = 0;
  for (int c : {16, 8, 4}) if (W % c == 0) { tw = c; break; }
  if (!tw) tw = W <= 4 ? 4 : (W <= 8 ? 8 : 16);   // over-wide tile: out-of-range columns are TMA zero fill
  int th = 1;
  while (th * 2 * tw <= 64 && H % (th * 2) == 0) th *= 2;
  int tn = 64 / (tw * th);
  if ((th * tw) % 8 != 0 || (th + 2) * tw * tn > 96) {   // fall back to one image per tile, partial tiles in H allowed
    th = 64 / tw;
    tn = 1;
  }
  *TW = tw; *TH = th; *TN = tn;
  return true;
}

static int conv3x3_wgrad_1x(const float* x, const float* dy, float* dwp, float* db, int N, int H, int W, int Cin, int Cout,
                            cudaStream_t stream, bool zero_dw, bool zero_db);

int conv3x3_wgrad(const float* x, const float* dy, float* dwp, float* db, int N, int H, int W, int Cin, int Cout,
                  cudaStream_t stream, bool zero_db = true) {
  if (!precise()) return conv3x3_wgrad_1x(x, dy, dwp, db, N, H, W, Cin, Cout, stream, true, zero_db);
  // 3xTF32: dW = dYh^T Xh + dYh^T Xl + dYl^T Xh accumulated by the kernel's own atomics; db = sum(dYh) + sum(dYl)
  HK_REQUIRE(x && dy && dwp, HK_ERR_ARG, "conv3x3_wgrad: null pointer");
  const size_t nx = (size_t)N * H * W * Cin, ny = (size_t)N * H * W * Cout;
  Scratch sx(2 * nx * sizeof(float), stream), sy(2 * ny * sizeof(float), stream);
  HK_REQUIRE(sx.p && sy.p, HK_ERR_DRIVER, "conv3x3_wgrad (precise): cudaMallocAsync of the operand halves failed");
  float *xh = sx.f(), *xl = xh + nx, *yh = sy.f(), *yl = yh + ny;
  int r;
  if ((r = tf32_split(x, xh, xl, nx, stream))) return r;
  if ((r = tf32_split(dy, yh, yl, ny, stream))) return r;
  if ((r = conv3x3_wgrad_1x(xh, yh, dwp, db, N, H, W, Cin, Cout, stream, true, zero_db))) return r;
  if ((r = conv3x3_wgrad_1x(xl, yh, dwp, nullptr, N, H, W, Cin, Cout, stream, false, false))) return r;
  return conv3x3_wgrad_1x(xh, yl, dwp, db, N, H, W, Cin, Cout, stream, false, false);
}

static int conv3x3_wgrad_1x(const float* x, const float* dy, float* dwp, float* db, int N, int H, int W, int Cin, int Cout,
                            cudaStream_t stream, bool zero_dw, bool zero_db) {
  HK_REQUIRE(x && dy && dwp, HK_ERR_ARG, "conv3x3_wgrad: null pointer");
  HK_REQUIRE(Cin % 32 == 0 && Cout % 32 == 0, HK_ERR_UNSUPPORTED, "conv3x3_wgrad: Cin=%d Cout=%d unsupported", Cin, Cout);
  WgradArgs a = {};
  a.dWp = dwp; a.db = db; a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout;
  pick_wgrad_tile(W, H, &a.TW, &a.TH, &a.TN);
  HK_REQUIRE((a.TH + 2) * a.TW * a.TN * 128 <= WG_B_ONE, HK_ERR_UNSUPPORTED, "conv3x3_wgrad: halo patch too large");
  a.tiles_w = (W + a.TW - 1) / a.TW; a.tiles_h = (H + a.TH - 1) / a.TH; a.tiles_n = (N + a.TN - 1) / a.TN;
  const long long out_tiles = (long long)((Cout + 127) / 128) * (Cin / 32);
  const long long total_tiles = (long long)a.tiles_w * a.tiles_h * a.tiles_n;
  // split-K factor.  The kernel runs one CTA per SM (204 KB of shared memory), so the grid executes in whole waves of
  // `sms` CTAs: pick the split that fills 1..3 waves best (ties -> fewer waves: fewer partial sums to add atomically).
  // Measured in one process on the whole BCNN step: 1178.7 -> 1235.2 img/s against the first rule ("about two CTAs of work
  // per SM", which left e.g. 304- and 320-CTA grids with a nearly empty third wave); HK_WG_SPLIT=0 restores that rule for A/B runs.
  static int sms = 0, rule = -1;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  if (rule < 0) { const char* v = getenv("HK_WG_SPLIT"); rule = v ? atoi(v) : 1; }
  long long ks = (148 * 2 + out_tiles - 1) / out_tiles;
  if (rule == 1) {
    double best = -1.0;
    for (int w = 1; w <= 3; ++w) {
      long long k = (long long)sms * w / out_tiles;
      if (k < 1) k = 1;
      if (k > total_tiles) k = total_tiles;
      const long long ctas = k * out_tiles;
      const long long waves = (ctas + sms - 1) / sms;
      const double fill = (double)ctas / (double)(waves * sms);
      if (fill > best + 1e-9) { best = fill; ks = k; }
    }
  }
  if (ks > total_tiles) ks = total_tiles;
  if (ks < 1) ks = 1;
  if (ks > 65535) ks = 65535;
  a.ksplit = (int)ks;
  { const char* v = getenv("HK_DBG_WG"); a.dbg = v ? atoi(v) : 0; }
  static int use_v2 = -1;
  if (use_v2 < 0) { const char* v = getenv("HK_WGRAD_V2"); use_v2 = v ? atoi(v) : 1; }
  CUtensorMap tmDY, tmX;
  int r;
  if ((r = make_act_map(&tmDY, dy, N, H, W, Cout, a.TW, a.TH, a.TN, /*mn_major=*/!use_v2))) return r;
  if ((r = make_act_map(&tmX, x, N, H, W, Cin, a.TW, a.TH + 2, a.TN, true))) return r;
  cudaError_t e = cudaSuccess;
  if (zero_dw) {
    e = cudaMemsetAsync(dwp, 0, (size_t)9 * Cout * Cin * sizeof(float), stream);
    if (e != cudaSuccess) return set_error((int)e, "cudaMemsetAsync(dWp): %s", cudaGetErrorString(e));
  }
  if (db && zero_db) {
    e = cudaMemsetAsync(db, 0, (size_t)Cout * sizeof(float), stream);
    if (e != cudaSuccess) return set_error((int)e, "cudaMemsetAsync(db): %s", cudaGetErrorString(e));
  }
  static bool attr_set = false;
  if (!attr_set) {
    e = cudaFuncSetAttribute(conv3x3_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv3x3_wgrad_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
    if (e != cudaSuccess) return set_error((int)e, "cudaFuncSetAttribute(wgrad): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((unsigned)out_tiles, a.ksplit);
  if (use_v2) conv3x3_wgrad_v2_kernel<<<grid, 192, WG_SMEM, stream>>>(tmDY, tmX, a);
  else conv3x3_wgrad_kernel<<<grid, 192, WG_SMEM, stream>>>(tmDY, tmX, a);
  HK_LAUNCH_CHECK("conv3x3_wgrad_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// first layer (Cin = 3; vgg.py:61 in_channels=3): im2col to X27 + one tcgen05 GEMM (see hk_conv3x3_first_fwd)
// ------------------------------------------------------------------------------------------------
// first-layer weight gradient on the tensor cores: materialise the 3x3x3 patches as X27 [pix][32]
// (27 taps, column 27 = 1.0 so that the GEMM's column 27 is the bias gradient, columns 28..31 = 0) and run
// dW^T-partials[s] = dY[pix-range s]^T . X27[pix-range s]  as a batched (split-K) MN-major tcgen05 GEMM.
__global__ void im2col_first_kernel(const float* __restrict__ x, float* __restrict__ x27, int N, int H, int W, int round) {
  const long long total = (long long)N * H * W;
  for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < total;
       pix += (long long)gridDim.x * blockDim.x) {
    const int wq = (int)(pix % W), hq = (int)((pix / W) % H), n = (int)(pix / ((long long)W * H));
    float v[32];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int hh = hq + kh - 1, ww = wq + kw - 1;
          float t = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(x + (((size_t)n * 3 + ci) * H + hh) * W + ww) : 0.f;
          v[ci * 9 + kh * 3 + kw] = round ? tf32_round(t) : t;
        }
    v[27] = 1.f; v[28] = v[29] = v[30] = v[31] = 0.f;
    float4* dst = reinterpret_cast<float4*>(x27 + (size_t)pix * 32);
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
}
// w27[co][0..26] = tf32(w[co][ci][kh][kw]), w27[co][27] = bias[co] (x27 column 27 is 1.0), rest 0
__global__ void pack_first_weights_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                          float* __restrict__ w27, int Cout, int round) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * 32) return;
  const int co = i / 32, r = i % 32;
  const float t = r < 27 ? w[co * 27 + r] : (r == 27 && bias ? bias[co] : 0.f);
  w27[i] = round ? tf32_round(t) : t;
}
// dw[co][r] = sum_s part[s][co][r] (r<27), db[co] = sum_s part[s][co][27]
__global__ void first_wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw,
                                          float* __restrict__ db, int Cout, int S, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * 28) return;
  const int co = i / 28, r = i % 28;
  float s = 0.f;
  for (int k = 0; k < S; ++k) s += part[((size_t)k * Cout + co) * 32 + r];
  if (r < 27) dw[co * 27 + r] = accumulate ? dw[co * 27 + r] + s : s;
  else if (db) db[co] = accumulate ? db[co] + s : s;
}

// ------------------------------------------------------------------------------------------------
// MaxPool2d(2,2) (vgg.py:59) on NHWC; optional NCHW output for the last pool (feeds the pooling head)
// ------------------------------------------------------------------------------------------------
__global__ void maxpool2x2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C,
                                      int out_nchw) {
  const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
  const size_t total = (size_t)N * Ho * Wo * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = i % C4;
    size_t p = i / C4;
    const int wo = p % Wo; p /= Wo;
    const int ho = p % Ho;
    const int n = p / Ho;
    const float4* base = reinterpret_cast<const float4*>(x + (((size_t)n * H + 2 * ho) * W + 2 * wo) * C) + c4;
    const float4 a = base[0], b = base[C4], c = base[(size_t)W * C4], d = base[(size_t)W * C4 + C4];
    float4 m;
    m.x = fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x));
    m.y = fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y));
    m.z = fmaxf(fmaxf(a.z, b.z), fmaxf(c.z, d.z));
    m.w = fmaxf(fmaxf(a.w, b.w), fmaxf(c.w, d.w));
    if (!out_nchw) {
      reinterpret_cast<float4*>(y + (((size_t)n * Ho + ho) * Wo + wo) * C)[c4] = m;
    } else {
      const size_t hw = (size_t)Ho * Wo, o = ((size_t)n * C + c4 * 4) * hw + (size_t)ho * Wo + wo;
      y[o] = m.x; y[o + hw] = m.y; y[o + 2 * hw] = m.z; y[o + 3 * hw] = m.w;
    }
  }
}

// backward: dx[window] = dy routed to the first max in scan order (PyTorch max_pool2d_backward), then multiplied by
// (x > 0): x is the ReLU output feeding the pool, so this also applies the preceding ReLU's backward.
__global__ void maxpool2x2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                      float* __restrict__ dx, int N, int H, int W, int C, int dy_nchw) {
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = (size_t)N * Ho * Wo * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = i % C;
    size_t p = i / C;
    const int wo = p % Wo; p /= Wo;
    const int ho = p % Ho;
    const int n = p / Ho;
    const size_t b00 = (((size_t)n * H + 2 * ho) * W + 2 * wo) * C + c;
    const size_t b01 = b00 + C, b10 = b00 + (size_t)W * C, b11 = b10 + C;
    const float v00 = x[b00], v01 = x[b01], v10 = x[b10], v11 = x[b11];
    const float g = dy_nchw ? dy[((size_t)n * C + c) * Ho * Wo + (size_t)ho * Wo + wo] : dy[i];
    int arg = 0;
    float m = v00;
    if (v01 > m) { m = v01; arg = 1; }
    if (v10 > m) { m = v10; arg = 2; }
    if (v11 > m) { m = v11; arg = 3; }
    const float gm = m > 0.f ? g : 0.f;
    dx[b00] = arg == 0 ? gm : 0.f;
    dx[b01] = arg == 1 ? gm : 0.f;
    dx[b10] = arg == 2 ? gm : 0.f;
    dx[b11] = arg == 3 ? gm : 0.f;
  }
}

// Training variants: the forward also records, per pooled element, ONE byte — bits 0-1 = window position of the first
// maximum in scan order (PyTorch routing), bit 2 = (max > 0), i.e. the mask of the ReLU that precedes the pool — and the
// backward routes dy from that byte alone: it no longer re-reads the four pre-pool activations (1/16 of the bytes).
__global__ void maxpool2x2_fwd_idx_kernel(const float* __restrict__ x, float* __restrict__ y, unsigned char* __restrict__ code,
                                          int N, int H, int W, int C, int out_nchw) {
  const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
  const size_t total = (size_t)N * Ho * Wo * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = i % C4;
    size_t p = i / C4;
    const int wo = p % Wo; p /= Wo;
    const int ho = p % Ho;
    const int n = p / Ho;
    const float4* base = reinterpret_cast<const float4*>(x + (((size_t)n * H + 2 * ho) * W + 2 * wo) * C) + c4;
    const float4 v0 = base[0], v1 = base[C4], v2 = base[(size_t)W * C4], v3 = base[(size_t)W * C4 + C4];
    float4 m = v0;
    uchar4 a = make_uchar4(0, 0, 0, 0);
#define HK_POOL_STEP(V, K)                     \
    if (V.x > m.x) { m.x = V.x; a.x = K; }     \
    if (V.y > m.y) { m.y = V.y; a.y = K; }     \
    if (V.z > m.z) { m.z = V.z; a.z = K; }     \
    if (V.w > m.w) { m.w = V.w; a.w = K; }
    HK_POOL_STEP(v1, 1) HK_POOL_STEP(v2, 2) HK_POOL_STEP(v3, 3)
#undef HK_POOL_STEP
    a.x |= m.x > 0.f ? 4 : 0; a.y |= m.y > 0.f ? 4 : 0; a.z |= m.z > 0.f ? 4 : 0; a.w |= m.w > 0.f ? 4 : 0;
    reinterpret_cast<uchar4*>(code)[i] = a;
    if (!out_nchw) {
      reinterpret_cast<float4*>(y + (((size_t)n * Ho + ho) * Wo + wo) * C)[c4] = m;
    } else {
      const size_t hw = (size_t)Ho * Wo, o = ((size_t)n * C + c4 * 4) * hw + (size_t)ho * Wo + wo;
      y[o] = m.x; y[o + hw] = m.y; y[o + 2 * hw] = m.z; y[o + 3 * hw] = m.w;
    }
  }
}
__global__ void maxpool2x2_bwd_idx_kernel(const unsigned char* __restrict__ code, const float* __restrict__ dy,
                                          float* __restrict__ dx, int N, int H, int W, int C, int dy_nchw) {
  const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
  const size_t total = (size_t)N * Ho * Wo * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = i % C4;
    size_t p = i / C4;
    const int wo = p % Wo; p /= Wo;
    const int ho = p % Ho;
    const int n = p / Ho;
    const uchar4 a = reinterpret_cast<const uchar4*>(code)[i];
    float4 g;
    if (!dy_nchw) {
      g = reinterpret_cast<const float4*>(dy)[i];
    } else {
      const size_t hw = (size_t)Ho * Wo, o = ((size_t)n * C + c4 * 4) * hw + (size_t)ho * Wo + wo;
      g = make_float4(dy[o], dy[o + hw], dy[o + 2 * hw], dy[o + 3 * hw]);
    }
    g.x = (a.x & 4) ? g.x : 0.f; g.y = (a.y & 4) ? g.y : 0.f; g.z = (a.z & 4) ? g.z : 0.f; g.w = (a.w & 4) ? g.w : 0.f;
    float4* base = reinterpret_cast<float4*>(dx + (((size_t)n * H + 2 * ho) * W + 2 * wo) * C) + c4;
#define HK_POOL_OUT(K) make_float4((a.x & 3) == K ? g.x : 0.f, (a.y & 3) == K ? g.y : 0.f, (a.z & 3) == K ? g.z : 0.f, \
                                   (a.w & 3) == K ? g.w : 0.f)
    base[0] = HK_POOL_OUT(0);
    base[C4] = HK_POOL_OUT(1);
    base[(size_t)W * C4] = HK_POOL_OUT(2);
    base[(size_t)W * C4 + C4] = HK_POOL_OUT(3);
#undef HK_POOL_OUT
  }
}

// db[c] = sum over pixels of dy[pix][c]   (NHWC)
__global__ void bias_grad_kernel(const float* __restrict__ dy, float* __restrict__ db, size_t npix, int C) {
  // block handles a slab of pixels; threads stride over channels (coalesced), atomics at the end
  const int c = threadIdx.x % C;
  const int lanes_per_c = blockDim.x / C > 0 ? blockDim.x / C : 1;
  const int sub = threadIdx.x / C;
  if (sub >= lanes_per_c) return;
  const size_t per = (npix + gridDim.x - 1) / gridDim.x;
  const size_t p0 = blockIdx.x * per, p1 = (p0 + per < npix) ? p0 + per : npix;
  for (int cc = c; cc < C; cc += (blockDim.x < C ? blockDim.x : C)) {
    float s = 0.f;
    for (size_t p = p0 + sub; p < p1; p += lanes_per_c) s += dy[p * C + cc];
    atomicAdd(db + cc, s);
  }
}

// elementwise: dy *= (act > 0)
__global__ void relu_mask_kernel(float* __restrict__ dy, const float* __restrict__ act, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 g = reinterpret_cast<float4*>(dy)[i];
    const float4 a = reinterpret_cast<const float4*>(act)[i];
    g.x = a.x > 0.f ? g.x : 0.f; g.y = a.y > 0.f ? g.y : 0.f; g.z = a.z > 0.f ? g.z : 0.f; g.w = a.w > 0.f ? g.w : 0.f;
    reinterpret_cast<float4*>(dy)[i] = g;
  }
}

static inline int grid_for(size_t n, int block) {
  size_t g = (n + block - 1) / block;
  const size_t cap = 148 * 16;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace hk

using namespace hk;

extern "C" {

int hk_conv3x3_pack_weights(const float* w, float* w_fwd, float* w_dgrad, int Cout, int Cin, void* stream) {
  HK_REQUIRE(w && (w_fwd || w_dgrad), HK_ERR_ARG, "hk_conv3x3_pack_weights: null pointer");
  pack_weights_kernel<<<grid_for((size_t)Cout * Cin * 9, 256), 256, 0, (cudaStream_t)stream>>>(w, w_fwd, w_dgrad, Cout, Cin,
                                                                                             precise() ? 0 : 1);
  HK_LAUNCH_CHECK("pack_weights_kernel");
  return 0;
}

int hk_conv3x3_fwd(const float* x, const float* w_packed, const float* bias, float* y, int N, int H, int W, int Cin,
                   int Cout, int relu, void* stream) {
  return conv3x3_igemm(x, w_packed, bias, nullptr, y, N, H, W, Cin, Cout, relu, (cudaStream_t)stream);
}

/* relu(conv3x3(x) + bias) followed by MaxPool2d(2,2) in ONE kernel: the full-resolution map is never written.  pooled:
 * [N,H/2,W/2,Cout] (NHWC) or [N,Cout,H/2,W/2] (out_nchw); code (optional): the byte per pooled element hk_maxpool2x2_bwd_idx
 * consumes.  Bit-identical to hk_conv3x3_fwd + hk_maxpool2x2_fwd_idx.  Single-pass TF32 only (HK_ERR_UNSUPPORTED in precise mode). */
int hk_conv3x3_fwd_pool(const float* x, const float* w_packed, const float* bias, float* pooled, unsigned char* code, int N,
                        int H, int W, int Cin, int Cout, int out_nchw, void* stream) {
  HK_REQUIRE(!precise(), HK_ERR_UNSUPPORTED, "hk_conv3x3_fwd_pool: not available in 3xTF32 mode (use conv + pool)");
  HK_REQUIRE(pooled, HK_ERR_ARG, "hk_conv3x3_fwd_pool: null output");
  return conv3x3_igemm_1x(x, w_packed, bias, nullptr, nullptr, nullptr, N, H, W, Cin, Cout, 1, (cudaStream_t)stream, 1, false,
                          0, pooled, code, out_nchw);
}

int hk_conv3x3_s2_fwd(const float* x, const float* w_packed, const float* bias, float* y, int N, int H, int W, int Cin,
                      int Cout, int relu, void* stream) {
  return conv3x3_igemm(x, w_packed, bias, nullptr, y, N, H, W, Cin, Cout, relu, (cudaStream_t)stream, 2);
}

int hk_conv3x3_dgrad(const float* dy, const float* w_dgrad_packed, const float* relu_mask_act, float* dx, int N, int H,
                     int W, int Cin, int Cout, void* stream) {
  // dgrad is the same implicit GEMM with the roles of Cin/Cout swapped and flipped taps
  return conv3x3_igemm(dy, w_dgrad_packed, nullptr, relu_mask_act, dx, N, H, W, Cout, Cin, 0, (cudaStream_t)stream);
}

size_t hk_conv3x3_wgrad_workspace_bytes(int Cin, int Cout) { return (size_t)9 * Cin * Cout * sizeof(float); }

int hk_conv3x3_wgrad_acc(const float* x, const float* dy, float* dw, float* db, int N, int H, int W, int Cin, int Cout,
                         void* workspace, size_t workspace_bytes, int accumulate, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  HK_REQUIRE(workspace && workspace_bytes >= hk_conv3x3_wgrad_workspace_bytes(Cin, Cout), HK_ERR_WORKSPACE,
             "hk_conv3x3_wgrad: workspace too small");
  HK_REQUIRE(dw, HK_ERR_ARG, "hk_conv3x3_wgrad: null dw");
  float* dwp = static_cast<float*>(workspace);
  int r = conv3x3_wgrad(x, dy, dwp, db, N, H, W, Cin, Cout, stream, /*zero_db=*/!accumulate);
  if (r) return r;
  unpack_wgrad_kernel<<<grid_for((size_t)Cout * Cin * 9, 256), 256, 0, stream>>>(dwp, dw, Cout, Cin, accumulate ? 1 : 0);
  HK_LAUNCH_CHECK("unpack_wgrad_kernel");
  return 0;
}

int hk_conv3x3_wgrad(const float* x, const float* dy, float* dw, float* db, int N, int H, int W, int Cin, int Cout,
                     void* workspace, size_t workspace_bytes, void* stream_) {
  return hk_conv3x3_wgrad_acc(x, dy, dw, db, N, H, W, Cin, Cout, workspace, workspace_bytes, 0, stream_);
}

size_t hk_conv3x3_first_fwd_workspace_bytes(int N, int H, int W, int Cout) {
  return ((size_t)N * H * W * 32 + (size_t)Cout * 32) * sizeof(float);
}

/* y = relu(conv3x3(x) + bias) for the 3-channel input layer: patches are materialised once as X27 [pix][32]
 * (tf32-rounded, column 27 = 1 carries the bias) and the layer is ONE tcgen05 GEMM  y = relu(X27 . W27^T).
 * workspace = X27 followed by W27; X27 (the first N*H*W*32 floats) is what hk_conv3x3_first_wgrad consumes. */
int hk_conv3x3_first_fwd(const float* x_nchw, const float* w, const float* bias, float* y_nhwc, int N, int H, int W,
                         int Cout, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  HK_REQUIRE(x_nchw && w && y_nhwc, HK_ERR_ARG, "hk_conv3x3_first_fwd: null pointer");
  HK_REQUIRE(Cout % 4 == 0 && Cout <= 256, HK_ERR_UNSUPPORTED, "hk_conv3x3_first_fwd: Cout=%d unsupported", Cout);
  HK_REQUIRE(workspace && workspace_bytes >= hk_conv3x3_first_fwd_workspace_bytes(N, H, W, Cout), HK_ERR_WORKSPACE,
             "hk_conv3x3_first_fwd: workspace too small");
  const long long P = (long long)N * H * W;
  HK_REQUIRE(P < (1ll << 31), HK_ERR_UNSUPPORTED, "hk_conv3x3_first_fwd: too many pixels");
  float* x27 = static_cast<float*>(workspace);
  float* w27 = x27 + (size_t)P * 32;
  const int round = precise() ? 0 : 1;
  im2col_first_kernel<<<grid_for((size_t)P, 128), 128, 0, stream>>>(x_nchw, x27, N, H, W, round);
  HK_LAUNCH_CHECK("im2col_first_kernel");
  pack_first_weights_kernel<<<(Cout * 32 + 127) / 128, 128, 0, stream>>>(w, bias, w27, Cout, round);
  HK_LAUNCH_CHECK("pack_first_weights_kernel");
  return hk_gemm_tf32(x27, 0, 32, 0, w27, 0, 32, 0, y_nhwc, Cout, 0, 0, (int)P, Cout, 32, 1, 1.f, nullptr, 0.f, nullptr,
                      0, 0, 0.f, nullptr, 3 /*relu + tf32 round*/, stream_);
}

static int first_wgrad_splits(long long P) {
  for (int S = 592; S > 1; --S)
    if (P % S == 0) return S;
  return 1;
}

size_t hk_conv3x3_first_wgrad_workspace_bytes(int N, int H, int W, int Cout) {
  return (size_t)first_wgrad_splits((long long)N * H * W) * Cout * 32 * sizeof(float);
}

/* dw [Cout,3,3,3], db [Cout] of the input layer from X27 (written by hk_conv3x3_first_fwd) and dy (ReLU-masked):
 * split-K batched MN-major tcgen05 GEMM  partial[s] = dY_s^T . X27_s ; column 27 of the result is the bias grad. */
int hk_conv3x3_first_wgrad_acc(const float* x27, const float* dy_nhwc, float* dw, float* db, int N, int H, int W, int Cout,
                               void* workspace, size_t workspace_bytes, int accumulate, void* stream_);
int hk_conv3x3_first_wgrad(const float* x27, const float* dy_nhwc, float* dw, float* db, int N, int H, int W,
                           int Cout, void* workspace, size_t workspace_bytes, void* stream_) {
  return hk_conv3x3_first_wgrad_acc(x27, dy_nhwc, dw, db, N, H, W, Cout, workspace, workspace_bytes, 0, stream_);
}
int hk_conv3x3_first_wgrad_acc(const float* x27, const float* dy_nhwc, float* dw, float* db, int N, int H, int W, int Cout,
                               void* workspace, size_t workspace_bytes, int accumulate, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  HK_REQUIRE(x27 && dy_nhwc && dw, HK_ERR_ARG, "hk_conv3x3_first_wgrad: null pointer");
  HK_REQUIRE(Cout % 4 == 0 && Cout <= 128, HK_ERR_UNSUPPORTED, "hk_conv3x3_first_wgrad: Cout=%d unsupported", Cout);
  HK_REQUIRE(workspace && workspace_bytes >= hk_conv3x3_first_wgrad_workspace_bytes(N, H, W, Cout), HK_ERR_WORKSPACE,
             "hk_conv3x3_first_wgrad: workspace too small");
  const long long P = (long long)N * H * W;
  const int S = first_wgrad_splits(P);
  const long long Kc = P / S;
  float* part = static_cast<float*>(workspace);
  int r = hk_gemm_tf32(dy_nhwc, 1, Cout, Kc * Cout, x27, 1, 32, Kc * 32, part, 32, (long long)Cout * 32, 0, Cout, 32,
                       (int)Kc, S, 1.f, nullptr, 0.f, nullptr, 0, 0, 0.f, nullptr, 0, stream_);
  if (r) return r;
  first_wgrad_reduce_kernel<<<(Cout * 28 + 127) / 128, 128, 0, stream>>>(part, dw, db, Cout, S, accumulate ? 1 : 0);
  HK_LAUNCH_CHECK("first_wgrad_reduce_kernel");
  return 0;
}

int hk_maxpool2x2_fwd(const float* x_nhwc, float* y, int N, int H, int W, int C, int out_nchw, void* stream) {
  HK_REQUIRE(x_nhwc && y, HK_ERR_ARG, "hk_maxpool2x2_fwd: null pointer");
  HK_REQUIRE(C % 4 == 0 && H % 2 == 0 && W % 2 == 0, HK_ERR_UNSUPPORTED, "hk_maxpool2x2_fwd: C%%4, even H/W required");
  const size_t total = (size_t)N * (H / 2) * (W / 2) * (C / 4);
  maxpool2x2_fwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x_nhwc, y, N, H, W, C, out_nchw);
  HK_LAUNCH_CHECK("maxpool2x2_fwd_kernel");
  return 0;
}

int hk_maxpool2x2_bwd(const float* x_nhwc, const float* dy, float* dx_nhwc, int N, int H, int W, int C, int dy_nchw,
                      void* stream) {
  HK_REQUIRE(x_nhwc && dy && dx_nhwc, HK_ERR_ARG, "hk_maxpool2x2_bwd: null pointer");
  HK_REQUIRE(H % 2 == 0 && W % 2 == 0, HK_ERR_UNSUPPORTED, "hk_maxpool2x2_bwd: even H/W required");
  const size_t total = (size_t)N * (H / 2) * (W / 2) * C;
  maxpool2x2_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x_nhwc, dy, dx_nhwc, N, H, W, C, dy_nchw);
  HK_LAUNCH_CHECK("maxpool2x2_bwd_kernel");
  return 0;
}

int hk_maxpool2x2_fwd_idx(const float* x_nhwc, float* y, unsigned char* code, int N, int H, int W, int C, int out_nchw,
                          void* stream) {
  HK_REQUIRE(x_nhwc && y && code, HK_ERR_ARG, "hk_maxpool2x2_fwd_idx: null pointer");
  HK_REQUIRE(C % 4 == 0 && H % 2 == 0 && W % 2 == 0, HK_ERR_UNSUPPORTED, "hk_maxpool2x2_fwd_idx: C%%4, even H/W required");
  const size_t total = (size_t)N * (H / 2) * (W / 2) * (C / 4);
  maxpool2x2_fwd_idx_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x_nhwc, y, code, N, H, W, C, out_nchw);
  HK_LAUNCH_CHECK("maxpool2x2_fwd_idx_kernel");
  return 0;
}

int hk_maxpool2x2_bwd_idx(const unsigned char* code, const float* dy, float* dx_nhwc, int N, int H, int W, int C,
                          int dy_nchw, void* stream) {
  HK_REQUIRE(code && dy && dx_nhwc, HK_ERR_ARG, "hk_maxpool2x2_bwd_idx: null pointer");
  HK_REQUIRE(C % 4 == 0 && H % 2 == 0 && W % 2 == 0, HK_ERR_UNSUPPORTED, "hk_maxpool2x2_bwd_idx: C%%4, even H/W required");
  const size_t total = (size_t)N * (H / 2) * (W / 2) * (C / 4);
  maxpool2x2_bwd_idx_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(code, dy, dx_nhwc, N, H, W, C, dy_nchw);
  HK_LAUNCH_CHECK("maxpool2x2_bwd_idx_kernel");
  return 0;
}

int hk_relu_mask_inplace(float* dy, const float* act, size_t n, void* stream) {
  HK_REQUIRE(dy && act && n % 4 == 0, HK_ERR_ARG, "hk_relu_mask_inplace: bad args");
  relu_mask_kernel<<<grid_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(dy, act, n / 4);
  HK_LAUNCH_CHECK("relu_mask_kernel");
  return 0;
}

}  // extern "C"
