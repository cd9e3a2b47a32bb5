// Generic batched TF32 GEMM on tcgen05 (UMMA) with TMA-fed, 128B-swizzled shared-memory
// operands and the fp32 accumulator in TMEM.  One 128 x BN output tile per CTA.
//
//   C[b] = alpha_b * (A[b] . B[b]) + diag * I + beta_b * D[b]           (optionally stored transposed)
//
// Both operands may be K-major or MN-major in global memory, so the same kernel serves
//   * Newton-Schulz chains and covariance of Fast MPN-COV (reference model/methods/MPNCOV.py:105-202),
//   * the (S . X) contraction of the bilinear / compact-bilinear backward (BCNN.py:13-27, CBCNN.py:96-135),
//   * 1x1 convolutions in NHWC (model/backbone/resnet.py:29-37).
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2-5 = epilogue.
#include <stdlib.h>

#include "common.cuh"
#include "host.h"
#include "gemm.h"
#include "../../include/hawkeye_b200.h"

namespace hk {


template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_TILE = 4 * 32 * 36 * 4;   // four epilogue warps x a 32 x 36-float transpose tile (coalesced row-major stores)
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_TILE;
};

// Persistent: one CTA per SM walks output tiles t, t+grid, ... (n-tile fastest, so CTAs running at the same time share
// the A rows through L2).  The smem ring and the two TMEM accumulator sets (2 x BN columns) run across tile
// boundaries: the epilogue of tile i overlaps the TMA/MMA of tile i+1.
template <int BN>
__global__ void __launch_bounds__(192, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmAl, const __grid_constant__ CUtensorMap tmBl, int triple, GemmEpi epi,
                 int M, int N, int K, int a_mn, int b_mn, int shareA, int shareB, int mn_sbo, int mn_type, int nstages,
                 int tiles_m, int tiles_n, int total_tiles) {
  // triple: operands are (hi, lo) tf32 pairs (tmA/tmB = hi, tmAl/tmBl = lo) and every k-step issues Ah.Bl, Al.Bh, Ah.Bh
  // into the same accumulator — the 3xTF32 product of the Newton-Schulz chain in ONE launch, no partial sums through memory
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + nstages * Cfg::A_BYTES;
  uint8_t* sAl = smem + nstages * Cfg::STAGE_BYTES;
  uint8_t* sBl = sAl + nstages * Cfg::A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + nstages * Cfg::STAGE_BYTES * (triple ? 2 : 1));
  uint64_t* empty = full + nstages;
  uint64_t* acc_full = empty + nstages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* epi_tiles = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = (K + 31) / 32;
  // triple mode keeps TWO accumulators per tile — the leading product Ah.Bh and the sum of the two correction products —
  // and adds them in the epilogue: the tensor core's accumulator addition truncates, so feeding terms 2^-11 the size of the
  // running sum into it loses them with a bias (the MPN 224x224 reference-gradient test moved by 10x when they shared one)
  const int acc_cols = triple ? 2 * BN : BN;
  const int TCOLS = 2 * acc_cols < 32 ? 32 : 2 * acc_cols;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TCOLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int kbg = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int nt = t % tiles_n;
        const int mt = (t / tiles_n) % tiles_m;
        const int bz = t / (tiles_n * tiles_m);
        const int m0 = mt * 128, n0 = nt * BN;
        const int bza = shareA ? 0 : bz, bzb = shareB ? 0 : bz;
        for (int kb = 0; kb < nk; ++kb, ++kbg) {
          const int s = kbg % nstages;
          const uint32_t ph = (kbg / nstages) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], Cfg::STAGE_BYTES * (triple ? 2 : 1));
          uint8_t* a = sA + s * Cfg::A_BYTES;
          uint8_t* b = sB + s * Cfg::B_BYTES;
          if (triple) {
            uint8_t* al = sAl + s * Cfg::A_BYTES;
            uint8_t* bl = sBl + s * Cfg::B_BYTES;
            if (!a_mn) {
              tma_load_3d(al, &tmAl, &full[s], kb * 32, m0, bza);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) tma_load_3d(al + j * 4096, &tmAl, &full[s], m0 + j * 32, kb * 32, bza);
            }
            if (!b_mn) {
              tma_load_3d(bl, &tmBl, &full[s], kb * 32, n0, bzb);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 32; ++j) tma_load_3d(bl + j * 4096, &tmBl, &full[s], n0 + j * 32, kb * 32, bzb);
            }
          }
          if (!a_mn) {
            tma_load_3d(a, &tmA, &full[s], kb * 32, m0, bza);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) tma_load_3d(a + j * 4096, &tmA, &full[s], m0 + j * 32, kb * 32, bza);
          }
          if (!b_mn) {
            tma_load_3d(b, &tmB, &full[s], kb * 32, n0, bzb);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 32; ++j) tma_load_3d(b + j * 4096, &tmB, &full[s], n0 + j * 32, kb * 32, bzb);
          }
        }
      }
    }
  } else if (warp == 1) {
    {   // warp-uniform loop; tcgen05 issue predicated on one elected lane
      const uint32_t idesc = make_idesc_tf32(128, BN, a_mn, b_mn);
      const uint64_t a_tmpl = a_mn ? make_sdesc(0, 4096, mn_sbo, mn_type) : make_sdesc(0, 16, 1024);
      const uint64_t b_tmpl = b_mn ? make_sdesc(0, 4096, mn_sbo, mn_type) : make_sdesc(0, 16, 1024);
      const uint32_t a_step = a_mn ? (1024u >> 4) : (32u >> 4);
      const uint32_t b_step = b_mn ? (1024u >> 4) : (32u >> 4);
      int kbg = 0, itl = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++itl) {
        const int set = itl & 1;
        mbar_wait(&acc_empty[set], ((itl >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + set * acc_cols;
        const uint32_t d_small = d + BN;
        for (int kb = 0; kb < nk; ++kb, ++kbg) {
          const int s = kbg % nstages;
          const uint32_t ph = (kbg / nstages) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint64_t a_base = a_tmpl + (smem_u32(sA + s * Cfg::A_BYTES) >> 4);
          const uint64_t b_base = b_tmpl + (smem_u32(sB + s * Cfg::B_BYTES) >> 4);
          const int krem = K - kb * 32;
          const int ksteps = krem >= 32 ? 4 : (krem + 7) / 8;
          if (elect_one()) {
            if (triple) {
              const uint64_t al_base = a_tmpl + (smem_u32(sAl + s * Cfg::A_BYTES) >> 4);
              const uint64_t bl_base = b_tmpl + (smem_u32(sBl + s * Cfg::B_BYTES) >> 4);
              for (int ks = 0; ks < ksteps; ++ks) {
                umma_tf32_ss(d_small, a_base + ks * a_step, bl_base + ks * b_step, idesc, (kb | ks) ? 1u : 0u);
                umma_tf32_ss(d_small, al_base + ks * a_step, b_base + ks * b_step, idesc, 1u);
                umma_tf32_ss(d, a_base + ks * a_step, b_base + ks * b_step, idesc, (kb | ks) ? 1u : 0u);
              }
            } else {
              for (int ks = 0; ks < ksteps; ++ks)
                umma_tf32_ss(d, a_base + ks * a_step, b_base + ks * b_step, idesc, (kb | ks) ? 1u : 0u);
            }
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
    int itl = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++itl) {
      const int nt = t % tiles_n;
      const int mt = (t / tiles_n) % tiles_m;
      const int bz = t / (tiles_n * tiles_m);
      const int m0 = mt * 128, n0 = nt * BN;
      const int row = m0 + q * 32 + lane;
      const int set = itl & 1;
      mbar_wait(&acc_full[set], (itl >> 1) & 1);
      tc_fence_after();
      const float alpha = epi.alpha * (epi.alpha_vec ? epi.alpha_vec[bz] : 1.f);
      const float beta = epi.beta * (epi.beta_vec ? epi.beta_vec[bz] : 1.f);
      float* Cb = epi.C + (long long)bz * epi.strideC;
      const float* Db = epi.D ? epi.D + (long long)bz * epi.strideD : nullptr;
      const float* Dlb = (epi.D && epi.D_lo) ? epi.D_lo + (long long)bz * epi.strideD : nullptr;
      const long long ldE = epi.ldE ? epi.ldE : epi.ldc;
      const float* Eb = epi.E ? epi.E + (long long)bz * (epi.ldE ? epi.strideE : epi.strideC) : nullptr;
      float* Clb = epi.C_lo ? epi.C_lo + (long long)bz * epi.strideC : nullptr;
      const bool simple = !Eb && !Clb && !Dlb && epi.diag == 0.f && !epi.trans_c && (!Db || epi.ldd != 0);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + set * acc_cols + c * 32, v);
        tmem_ld_wait();
        if (triple) {
          float sm[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + set * acc_cols + BN + c * 32, sm);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += sm[j];
        }
        const int col0 = n0 + c * 32;
        // fast path (plain scaled store of a full, aligned 32-column chunk): a handful of instructions per element — the
        // general path below costs ~60 and dominates short-K GEMMs such as the first VGG layer (K = 32)
        if (simple && col0 + 32 <= N && (epi.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(Cb) & 15) == 0 &&
            (!Db || ((epi.ldd & 3) == 0 && (reinterpret_cast<uintptr_t>(Db) & 15) == 0))) {
          // A thread owns 32 consecutive columns of ONE row, so a direct float4 store instruction would touch 32 rows x 16 B —
          // half-written 32 B sectors, twice the L2 write transactions (the 1x1-conv GEMMs of the ResNet trunk are
          // output-write-bound).  Transpose the 32 x 32 chunk through a warp-private shared-memory tile instead: every store
          // instruction then writes four full 128 B rows; the optional addend D (residual gradient) is read the same way.
          float* tile = epi_tiles + (warp - 2) * (32 * 36);
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(tile + lane * 36 + j) =
                make_float4(alpha * v[j], alpha * v[j + 1], alpha * v[j + 2], alpha * v[j + 3]);
          __syncwarp();
          const int row0 = m0 + q * 32;
          float4 dadd[8];
          if (Db) {      // all eight addend loads in flight before the first store (C and D may alias as far as the compiler knows)
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int r = it * 4 + (lane >> 3);
              dadd[it] = (row0 + r < M) ? __ldg(reinterpret_cast<const float4*>(Db + (long long)(row0 + r) * epi.ldd + col0 + (lane & 7) * 4))
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + (lane >> 3);
            float4 t = *reinterpret_cast<const float4*>(tile + r * 36 + (lane & 7) * 4);
            if (row0 + r < M) {
              if (Db) {
                const float4 d = dadd[it];
                t.x = fmaf(beta, d.x, t.x); t.y = fmaf(beta, d.y, t.y); t.z = fmaf(beta, d.z, t.z); t.w = fmaf(beta, d.w, t.w);
              }
              if (epi.relu & 1) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
              if (epi.relu & 2) { t.x = tf32_round(t.x); t.y = tf32_round(t.y); t.z = tf32_round(t.z); t.w = tf32_round(t.w); }
              *reinterpret_cast<float4*>(Cb + (long long)(row0 + r) * epi.ldc + col0 + (lane & 7) * 4) = t;
            }
          }
          __syncwarp();
          continue;
        }
        if (row < M && col0 < N) {
          // general epilogue: raw addend E, alpha, diagonal, beta*(D [+ D_lo]), ReLU / rounding, plain / (hi, lo) / transposed
          // store.  A thread owns 32 consecutive columns of one row: whole aligned chunks move as float4 (the Newton-Schulz
          // chain lives on this path; scalar accesses made its epilogues 3-8x longer than the MMAs)
          const bool full = col0 + 32 <= N;
          const float* ep = Eb ? Eb + (long long)row * ldE + col0 : nullptr;
          const float* dp = Db ? Db + (long long)row * epi.ldd + col0 : nullptr;
          const float* dlp = Dlb ? Dlb + (long long)row * epi.ldd + col0 : nullptr;
          auto vec_ok = [&](const void* q) { return full && (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
          if (ep) {
            if (vec_ok(ep)) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 t = *reinterpret_cast<const float4*>(ep + j);
                v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < N) v[j] += ep[j];
            }
          }
          float dv[32];
          if (dp) {
            if (vec_ok(dp) && (!dlp || vec_ok(dlp))) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 t = *reinterpret_cast<const float4*>(dp + j);
                if (dlp) {
                  const float4 u = *reinterpret_cast<const float4*>(dlp + j);
                  t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
                }
                dv[j] = t.x; dv[j + 1] = t.y; dv[j + 2] = t.z; dv[j + 3] = t.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                dv[j] = 0.f;
                if (col0 + j < N) dv[j] = dp[j] + (dlp ? dlp[j] : 0.f);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float o = alpha * v[j];
            if (col0 + j == row) o += epi.diag;
            if (dp) o += beta * dv[j];
            if (epi.relu & 1) o = fmaxf(o, 0.f);
            if (epi.relu & 2) o = tf32_round(o);   // output feeds another tf32 MMA: keep its error unbiased
            v[j] = o;
          }
          if (Clb) {   // (hi, lo) split store; not combined with trans_c
            float* hp = Cb + (long long)row * epi.ldc + col0;
            float* lp = Clb + (long long)row * epi.ldc + col0;
            if (vec_ok(hp) && vec_ok(lp)) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float h0 = tf32_round(v[j]), h1 = tf32_round(v[j + 1]), h2 = tf32_round(v[j + 2]), h3 = tf32_round(v[j + 3]);
                *reinterpret_cast<float4*>(hp + j) = make_float4(h0, h1, h2, h3);
                *reinterpret_cast<float4*>(lp + j) = make_float4(tf32_round(v[j] - h0), tf32_round(v[j + 1] - h1),
                                                                 tf32_round(v[j + 2] - h2), tf32_round(v[j + 3] - h3));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float hi = tf32_round(v[j]);
                if (col0 + j < N) { hp[j] = hi; lp[j] = tf32_round(v[j] - hi); }
              }
            }
          } else if (!epi.trans_c) {
            float* dst = Cb + (long long)row * epi.ldc + col0;
            if (vec_ok(dst)) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < N) dst[j] = v[j];
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < N) Cb[(long long)(col0 + j) * epi.ldc + row] = v[j];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[set]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TCOLS);
}

static int make_operand_map(CUtensorMap* tm, const float* P, int mn_major, long long ld, long long stride, int rows,
                            int K, int batch, int tile_rows, int* share) {
  *share = (stride == 0 || batch == 1);
  uint64_t dims[3], strides[2];
  uint32_t box[3];
  const uint64_t nb = *share ? 1 : (uint64_t)batch;
  const uint64_t bs = *share ? (uint64_t)(mn_major ? K : rows) * ld * 4 : (uint64_t)stride * 4;
  if (!mn_major) {  // [rows][K], K contiguous
    dims[0] = K; dims[1] = rows; dims[2] = nb;
    box[0] = 32; box[1] = tile_rows; box[2] = 1;
  } else {          // [K][rows], rows contiguous
    dims[0] = rows; dims[1] = K; dims[2] = nb;
    box[0] = 32; box[1] = 32; box[2] = 1;
  }
  strides[0] = (uint64_t)ld * 4;
  strides[1] = bs;
  return make_tmap(tm, P, 3, dims, strides, box, mn_major != 0);
}

static int dbg_env(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

template <int BN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmEpi& epi, int M, int N, int K,
                       int batch, int a_mn, int b_mn, int shareA, int shareB, cudaStream_t stream,
                       const CUtensorMap* tmAl = nullptr, const CUtensorMap* tmBl = nullptr) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(umma_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) return set_error((int)e, "cudaFuncSetAttribute(gemm<%d>): %s", BN, cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles_m = (M + 127) / 128, tiles_n = (N + BN - 1) / BN;
  const long long total = (long long)tiles_m * tiles_n * batch;
  if (total >= (1ll << 31)) return set_error(HK_ERR_UNSUPPORTED, "gemm: too many tiles");
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  const int nk = (K + 31) / 32;
  const long long kblocks_per_cta = (long long)nk * ((total + sms - 1) / sms);
  const int triple = tmAl ? 1 : 0;
  const int max_stages = triple ? Cfg::STAGES / 2 : Cfg::STAGES;
  const int nstages = kblocks_per_cta < max_stages ? (int)kblocks_per_cta : max_stages;
  const int smem = nstages * Cfg::STAGE_BYTES * (triple ? 2 : 1) + 1024 + 256 + Cfg::EPI_TILE;
  const int grid = total < sms ? (int)total : sms;
  umma_gemm_kernel<BN><<<grid, 192, smem, stream>>>(tmA, tmB, triple ? *tmAl : tmA, triple ? *tmBl : tmB, triple, epi, M, N, K,
                                                    a_mn, b_mn, shareA, shareB,
                                                    dbg_env("HK_DBG_MN_SBO", 512), dbg_env("HK_DBG_MN_TYPE", 1), nstages,
                                                    tiles_m, tiles_n, (int)total);
  HK_LAUNCH_CHECK("umma_gemm_kernel");
  return 0;
}

__global__ void tf32_split_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float h = tf32_round(v);
    if (hi) hi[i] = h;
    if (lo) lo[i] = tf32_round(v - h);
  }
}

int tf32_split(const float* x, float* hi, float* lo, size_t n, cudaStream_t stream) {
  if (!n) return 0;
  size_t g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  tf32_split_kernel<<<(unsigned)g, 256, 0, stream>>>(x, hi, lo, n);
  HK_LAUNCH_CHECK("tf32_split_kernel");
  return 0;
}

// floats spanned by a strided operand: `batch` matrices of `rows` x `cols` (cols contiguous), leading dimension ld
static size_t operand_extent(long long rows, long long cols, long long ld, long long stride, int batch) {
  return (size_t)((long long)(batch - 1) * (stride > 0 ? stride : 0) + (rows - 1) * ld + cols);
}

// 3xTF32: A.B ~= Ah.Bh + Al.Bh + Ah.Bl with (hi, lo) = tf32 halves of the operands, chained through the epilogue's raw
// addend E so the caller's epilogue (alpha, diag, D, ReLU, transposed store ...) is applied once, to the full sum.
int gemm_tf32_3x(const float* A, int a_mn, long long lda, long long strideA, const float* B, int b_mn, long long ldb,
                        long long strideB, const GemmEpi& epi, int M, int N, int K, int batch, cudaStream_t st) {
  const size_t nA = operand_extent(a_mn ? K : M, a_mn ? M : K, lda, strideA, batch);
  const bool same = (A == B && a_mn == b_mn && lda == ldb && strideA == strideB && M == N);
  const size_t nB = same ? 0 : operand_extent(b_mn ? K : N, b_mn ? N : K, ldb, strideB, batch);
  const size_t nA4 = (nA + 3) & ~size_t(3), nB4 = (nB + 3) & ~size_t(3);     // the lo halves start 16-byte aligned (TMA)
  Scratch sa(2 * nA4 * sizeof(float), st), sb(2 * (nB4 ? nB4 : 4) * sizeof(float), st);
  HK_REQUIRE(sa.p && sb.p, HK_ERR_DRIVER, "gemm (precise): cudaMallocAsync of the operand halves failed");
  float *Ah = sa.f(), *Al = Ah + nA4;
  float *Bh = same ? Ah : sb.f(), *Bl = same ? Al : Bh + nB4;
  int r;
  if ((r = tf32_split(A, Ah, Al, nA, st))) return r;
  if (!same && (r = tf32_split(B, Bh, Bl, nB, st))) return r;
  GemmEpi f = epi;
  f.relu &= ~2;                                                                                             // no rounding
  return gemm_tf32_pair(Ah, Al, a_mn, lda, strideA, Bh, Bl, b_mn, ldb, strideB, f, M, N, K, batch, st);     // one launch
}

int gemm_tf32(const float* A, int a_mn, long long lda, long long strideA, const float* B, int b_mn, long long ldb,
              long long strideB, const GemmEpi& epi, int M, int N, int K, int batch, cudaStream_t stream) {
  if (precise()) {
    HK_REQUIRE(A && B && epi.C, HK_ERR_ARG, "gemm: null pointer");
    HK_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0, HK_ERR_ARG, "gemm: bad shape M=%d N=%d K=%d batch=%d", M, N, K, batch);
    return gemm_tf32_3x(A, a_mn, lda, strideA, B, b_mn, ldb, strideB, epi, M, N, K, batch, stream);
  }
  return gemm_tf32_1x(A, a_mn, lda, strideA, B, b_mn, ldb, strideB, epi, M, N, K, batch, stream);
}

int gemm_tf32_1x(const float* A, int a_mn, long long lda, long long strideA, const float* B, int b_mn, long long ldb,
                 long long strideB, const GemmEpi& epi, int M, int N, int K, int batch, cudaStream_t stream) {
  HK_REQUIRE(A && B && epi.C, HK_ERR_ARG, "gemm: null pointer");
  HK_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0, HK_ERR_ARG, "gemm: bad shape M=%d N=%d K=%d batch=%d", M, N, K, batch);
  HK_REQUIRE(batch <= 65535, HK_ERR_UNSUPPORTED, "gemm: batch %d > 65535", batch);
  CUtensorMap tmA, tmB;
  int shareA, shareB, r;
  const int BN = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  if ((r = make_operand_map(&tmA, A, a_mn, lda, strideA, M, K, batch, 128, &shareA))) return r;
  if ((r = make_operand_map(&tmB, B, b_mn, ldb, strideB, N, K, batch, BN, &shareB))) return r;
  if (BN == 64) return launch_gemm<64>(tmA, tmB, epi, M, N, K, batch, a_mn, b_mn, shareA, shareB, stream);
  if (BN == 128) return launch_gemm<128>(tmA, tmB, epi, M, N, K, batch, a_mn, b_mn, shareA, shareB, stream);
  return launch_gemm<256>(tmA, tmB, epi, M, N, K, batch, a_mn, b_mn, shareA, shareB, stream);
}

// C = epilogue(Ah.Bh + Al.Bh + Ah.Bl) in ONE launch; (Ah, Al) / (Bh, Bl) are tf32 (hi, lo) pairs with identical layouts.
int gemm_tf32_pair(const float* Ah, const float* Al, int a_mn, long long lda, long long strideA, const float* Bh,
                   const float* Bl, int b_mn, long long ldb, long long strideB, const GemmEpi& epi, int M, int N, int K,
                   int batch, cudaStream_t stream) {
  HK_REQUIRE(Ah && Al && Bh && Bl && epi.C, HK_ERR_ARG, "gemm_pair: null pointer");
  HK_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0 && batch <= 65535, HK_ERR_ARG, "gemm_pair: bad shape M=%d N=%d K=%d batch=%d",
             M, N, K, batch);
  CUtensorMap tmA, tmB, tmAl, tmBl;
  int shareA, shareB, r;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  // 128-wide tiles while 256-wide ones would leave SMs idle (Newton-Schulz: n = 256, batch 32 -> 128 tiles instead of 64)
  // two accumulators per tile x double buffering = 4 BN TMEM columns: BN <= 128
  const int BN = N <= 64 ? 64 : 128;
  if ((r = make_operand_map(&tmA, Ah, a_mn, lda, strideA, M, K, batch, 128, &shareA))) return r;
  if ((r = make_operand_map(&tmAl, Al, a_mn, lda, strideA, M, K, batch, 128, &shareA))) return r;
  if ((r = make_operand_map(&tmB, Bh, b_mn, ldb, strideB, N, K, batch, BN, &shareB))) return r;
  if ((r = make_operand_map(&tmBl, Bl, b_mn, ldb, strideB, N, K, batch, BN, &shareB))) return r;
  if (BN == 64) return launch_gemm<64>(tmA, tmB, epi, M, N, K, batch, a_mn, b_mn, shareA, shareB, stream, &tmAl, &tmBl);
  return launch_gemm<128>(tmA, tmB, epi, M, N, K, batch, a_mn, b_mn, shareA, shareB, stream, &tmAl, &tmBl);
}

}  // namespace hk

// same signature as hk_gemm_tf32, always 3xTF32 (callers whose result feeds an exponential: CIN's softmax(-Gram))
extern "C" int hk_gemm_3xtf32(const float* A, int a_mn_major, long long lda, long long strideA, const float* B,
                              int b_mn_major, long long ldb, long long strideB, float* C, long long ldc,
                              long long strideC, int trans_c, int M, int N, int K, int batch, float alpha,
                              const float* alpha_vec, float diag, const float* D, long long ldd, long long strideD,
                              float beta, const float* beta_vec, int relu, void* stream) {
  hk::GemmEpi epi;
  epi.C = C; epi.ldc = ldc; epi.strideC = strideC;
  epi.D = D; epi.ldd = ldd; epi.strideD = strideD;
  epi.alpha_vec = alpha_vec; epi.beta_vec = beta_vec;
  epi.alpha = alpha; epi.beta = beta; epi.diag = diag;
  epi.trans_c = trans_c; epi.relu = relu;
  epi.C_lo = nullptr; epi.D_lo = nullptr; epi.E = nullptr; epi.ldE = 0; epi.strideE = 0;
  HK_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && batch > 0, HK_ERR_ARG, "hk_gemm_3xtf32: bad args");
  return hk::gemm_tf32_3x(A, a_mn_major, lda, strideA, B, b_mn_major, ldb, strideB, epi, M, N, K, batch,
                          static_cast<cudaStream_t>(stream));
}

extern "C" int hk_gemm_tf32(const float* A, int a_mn_major, long long lda, long long strideA, const float* B,
                            int b_mn_major, long long ldb, long long strideB, float* C, long long ldc,
                            long long strideC, int trans_c, int M, int N, int K, int batch, float alpha,
                            const float* alpha_vec, float diag, const float* D, long long ldd, long long strideD,
                            float beta, const float* beta_vec, int relu, void* stream) {
  hk::GemmEpi epi;
  epi.C = C; epi.ldc = ldc; epi.strideC = strideC;
  epi.D = D; epi.ldd = ldd; epi.strideD = strideD;
  epi.alpha_vec = alpha_vec; epi.beta_vec = beta_vec;
  epi.alpha = alpha; epi.beta = beta; epi.diag = diag;
  epi.trans_c = trans_c; epi.relu = relu;
  epi.C_lo = nullptr; epi.D_lo = nullptr; epi.E = nullptr; epi.ldE = 0; epi.strideE = 0;
  return hk::gemm_tf32(A, a_mn_major, lda, strideA, B, b_mn_major, ldb, strideB, epi, M, N, K, batch,
                       static_cast<cudaStream_t>(stream));
}
