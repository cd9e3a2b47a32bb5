// K1 (tile kernel) — fused bilinear pooling forward (reference model/methods/BCNN.py:13-27):
//     G = X X^T / HW ; z = sqrt(G + 1e-5) ; y = z / max(||z||_2, 1e-12)            X: [B, C, HW]  ->  y: [B, C*C]
// ONE launch of persistent CTAs (one per SM) over 128x128 Gram tiles; see the banner above bcnn_gram_fwd_kernel.
//
// Progress guarantee.  The L2 norm of an image needs the sums of all its tiles, which are computed by several CTAs; they
// are exchanged through launch-tagged 64-bit words in a small library-owned table (value and validity travel together).
// That exchange is an OPTIMISATION, never a dependency: a CTA that has polled for GF_POLL_LIMIT rounds without seeing all
// tags (its peers are not resident because another kernel holds their SMs, or a concurrent call overwrote the slots)
// computes the norm itself from X — ||z||^2 = sum_p (sum_c x_cp)^2 / HW + C^2 eps — and carries on.  No CTA ever waits
// unboundedly on another one, so the kernel is correct under any co-scheduling and any number of concurrent calls.
#include <stdlib.h>

#include <atomic>

#include "common.cuh"
#include "host.h"
#include "../../include/hawkeye_b200.h"

namespace hk {

__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

constexpr int GRAM_SLOT = 128 * 128;              // 16 KB: 128 rows x 32 fp32
constexpr int GRAM_STAGE_BYTES = 2 * GRAM_SLOT;   // two row blocks per stage

constexpr int GRAM_CNT_REGIONS = 32;
constexpr int GRAM_CNT_MAXB = 2048;

// Tagged tile-sum slots of the single-launch forward: library-owned, zero at module load, never reset (a slot is valid
// for a launch iff it carries that launch's tag).  Regions are handed out round-robin per call so that calls in flight on
// different streams do not overwrite each other's slots.
__device__ unsigned long long g_gram_slots[(size_t)GRAM_CNT_REGIONS * GRAM_CNT_MAXB * 16];

static unsigned long long* gram_slots(unsigned int* tag) {
  static std::atomic<unsigned> next{0};
  static thread_local int dev_cached = -1;
  static thread_local unsigned long long* base = nullptr;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != dev_cached) {
    void* p = nullptr;
    if (cudaGetSymbolAddress(&p, g_gram_slots) != cudaSuccess) return nullptr;
    base = static_cast<unsigned long long*>(p);
    dev_cached = dev;
  }
  const unsigned n = next.fetch_add(1);
  *tag = n + 1 ? n + 1 : 1;     // never 0 (the initial slot contents)
  return base + (size_t)(n % GRAM_CNT_REGIONS) * GRAM_CNT_MAXB * 16;
}


// =====================================================================================================================
// K1 (third version): fused Gram + sqrt + L2-normalise with NO pre-kernel and half the tensor-core work.
//   item   = one 128x128 Gram tile (bi <= bj) of one image, ONE accumulator (128 TMEM columns, 4-slot ring):
//            off-diagonal tiles are computed once and written twice — block (bj,bi) by the transposed, lane-coalesced
//            direct stores, block (bi,bj) row-major through 128B-swizzled shared memory + TMA bulk-tensor stores;
//   norm   = ||z||^2 = sum_ij G_ij / HW + C^2 eps.  Every item publishes the sum of its tile (x2 off the diagonal) as one
//            launch-tagged 64-bit word; a dedicated warp collects the ipi words of the image.  The exchange is software-
//            pipelined: the sum of item k+1 is published BEFORE item k is normalised and stored, so the cross-CTA latency
//            hides behind a whole tile of stores and the MMA warp runs up to three tiles ahead.
// =====================================================================================================================
constexpr int GF_STAGES = 5;             // shared-memory ring depth allocated (GfArgs::stages of them are used)
constexpr int GF_SLOTS = 16;              // tagged tile-sum slots per image (>= ipi)
constexpr int GF_OUT_BYTES = 128 * 128;   // one 128-row x 32-column fp32 box
constexpr int GF_MAX_ITEMS = 384;         // items per CTA per launch (GRAM_CNT_MAXB images x <= 16 tiles over >= 148 CTAs, with slack)
constexpr int GF_SMEM = GF_STAGES * GRAM_STAGE_BYTES + 4 * GF_OUT_BYTES + 1024 + 512 + 2 * GF_MAX_ITEMS;

struct GfArgs {
  int B, C, HW, nblk;
  float inv_hw, eps;
  float* Y;
  float* inv_norm;
  unsigned long long* slots;   // [B][GF_SLOTS] tagged tile sums {tag:32 | f32 bits:32} (library-owned, never reset)
  unsigned int tag;            // unique per launch, never 0
  int store_mode, x_hint;
  int dbg;               // profiling only (HK_GRAM_DBG): 1 no direct stores, 2 no TMA stores, 4 no norm exchange, 8 no loads/MMA
  unsigned long long* trace;   // profiling only: [grid][16] globaltimer stamps, or null
  const float* X;              // [B][C][HW]: read directly only by the norm fallback
  int poll_limit;              // polls of the slot table before a CTA computes the norm of the image itself
  int stages;                  // 2..GF_STAGES
  int pdl;                     // launched with programmatic stream serialization
  int balance;                 // 1: units-balanced item schedule (see the kernel prologue), 0: plain round-robin
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void gf_item(int t, int nblk, int& bi, int& bj) {
  const int n_off = nblk * (nblk - 1) / 2;
  if (t < n_off) {
    int i = 0;
    while (t >= nblk - 1 - i) { t -= nblk - 1 - i; ++i; }
    bi = i; bj = i + 1 + t;
  } else {
    bi = bj = t - n_off;
  }
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(m),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

constexpr int GF_THREADS = 352;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue, warp 10 norm exchange

// CT: compile-time C (row pitch of Y in floats) so the 32 transposed stores of a chunk use immediate offsets; 0 = runtime C.
template <int CT>
__global__ void __launch_bounds__(GF_THREADS, 1)
bcnn_gram_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, GfArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* outbuf = smem + GF_STAGES * GRAM_STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(outbuf + 4 * GF_OUT_BYTES);
  uint64_t* empty = full + GF_STAGES;
  uint64_t* acc_full = empty + GF_STAGES;      // [4]  MMA -> epilogue
  uint64_t* acc_empty = acc_full + 4;          // [4]  epilogue -> MMA
  uint64_t* sum_ready = acc_empty + 4;         // [4]  epilogue (8 warps) -> norm warp: tile sums of item k in sum_part[k&3]
  uint64_t* norm_ready = sum_ready + 4;        // [4]  norm warp -> epilogue: inv_norm of item k in inv_box[k&3]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(norm_ready + 4);
  float* sum_part = reinterpret_cast<float*>(tmem_slot + 2);   // [4][8]
  float* inv_box = sum_part + 32;                              // [4]
  int* n_my_box = reinterpret_cast<int*>(inv_box + 4);
  uint16_t* sched = reinterpret_cast<uint16_t*>(n_my_box + 1);  // [GF_MAX_ITEMS] item ids (b * ipi + t) of this CTA, in order

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = CT ? CT : a.C;
  const int ipi = a.nblk * (a.nblk + 1) / 2;     // items (tiles bi <= bj) per image
  const int total_items = a.B * ipi;
  const int nk = (a.HW + 31) / 32;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
    for (int s = 0; s < GF_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8);
      mbar_init(&sum_ready[s], 8); mbar_init(&norm_ready[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2 && lane == 0) {
    // Item schedule of this CTA.  Off-diagonal tiles cost two output blocks, diagonal tiles one: off-diagonal items go
    // round-robin over the G CTAs; diagonal items first top up the CTAs that got one off-diagonal item fewer (two each),
    // then continue round-robin (longest-processing-time-first for two job sizes).  Each CTA walks its items in image order,
    // so the tiles of an image are in flight at the same time on all its CTAs and dependencies only point to earlier images.
    const int G = gridDim.x, c = blockIdx.x;
    const int n_off = a.nblk * (a.nblk - 1) / 2;
    const int O = a.B * n_off, D = a.B * a.nblk;
    const int r = O % G, L = G - r;
    const int END = 0x7fffffff;
    int oi = c < O ? c : END;
    int dj = END;
    if (a.balance) {
      if (c >= r && c - r < 2 * L && c - r < D) dj = c - r;
      else if (2 * L + c < D) dj = 2 * L + c;
    } else {
      oi = END;       // plain round-robin over the image-major item list
    }
    int n = 0;
    if (a.balance) {
      while ((oi != END || dj != END) && n < GF_MAX_ITEMS) {
        const int bo = oi != END ? oi / n_off : END, bd = dj != END ? dj / a.nblk : END;
        if (bo <= bd) {
          sched[n++] = (uint16_t)(bo * ipi + (oi - bo * n_off));
          oi = oi + G < O ? oi + G : END;
        } else {
          sched[n++] = (uint16_t)(bd * ipi + n_off + (dj - bd * a.nblk));
          int nx;
          if (dj < 2 * L) {
            nx = dj + L;
            if (nx >= 2 * L) nx = 2 * L + c;
          } else {
            nx = dj + G;
          }
          dj = nx < D ? nx : END;
        }
      }
    } else {
      for (int it = c; it < total_items && n < GF_MAX_ITEMS; it += G) sched[n++] = (uint16_t)it;
    }
    *n_my_box = n;
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_my = *n_my_box;
  if (a.pdl) {
    // programmatic dependent launch: the next grid may begin its prologue as soon as every CTA of this one got here; this
    // grid must not touch global memory before its predecessor has completed and flushed.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  unsigned long long* tr = a.trace ? a.trace + (size_t)blockIdx.x * 16 : nullptr;
  if (tr && threadIdx.x == 0) tr[0] = gtimer();

  if (warp == 0) {
    if (lane == 0) {
      uint64_t policy;
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
      int kbg = 0;
      for (int k = 0; k < n_my && !(a.dbg & 8); ++k) {
        const int it = sched[k];
        const int b = it / ipi;
        int bi, bj;
        gf_item(it - b * ipi, a.nblk, bi, bj);
        for (int kb = 0; kb < nk; ++kb, ++kbg) {
          const int s = kbg % a.stages;
          const uint32_t ph = (kbg / a.stages) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], (bi != bj ? 2 : 1) * GRAM_SLOT);
          uint8_t* st = smem + s * GRAM_STAGE_BYTES;
          if (a.x_hint) {
            tma_load_3d_hint(st, &tmX, &full[s], kb * 32, bi * 128, b, policy);
            if (bi != bj) tma_load_3d_hint(st + GRAM_SLOT, &tmX, &full[s], kb * 32, bj * 128, b, policy);
          } else {
            tma_load_3d(st, &tmX, &full[s], kb * 32, bi * 128, b);
            if (bi != bj) tma_load_3d(st + GRAM_SLOT, &tmX, &full[s], kb * 32, bj * 128, b);
          }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_tf32(128, 128, 0, 0);
    const uint64_t desc_tmpl = make_sdesc(0, 16, 1024);
    int kbg = 0, itl = 0;
    for (; itl < n_my && !(a.dbg & 8); ++itl) {
      const int it = sched[itl];
      const int b = it / ipi;
      int bi, bj;
      gf_item(it - b * ipi, a.nblk, bi, bj);
      const int slot = itl & 3;
      mbar_wait(&acc_empty[slot], ((itl >> 2) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + slot * 128;
      for (int kb = 0; kb < nk; ++kb, ++kbg) {
        const int s = kbg % a.stages;
        const uint32_t ph = (kbg / a.stages) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (tr && kbg == 0 && lane == 0) tr[1] = gtimer();
        const uint32_t s0 = smem_u32(smem + s * GRAM_STAGE_BYTES);
        const uint64_t d0 = desc_tmpl + (s0 >> 4);
        const uint64_t d1 = (bi != bj) ? desc_tmpl + ((s0 + GRAM_SLOT) >> 4) : d0;
        const int krem = a.HW - kb * 32;
        const int ksteps = krem >= 32 ? 4 : (krem + 7) / 8;
        if (elect_one()) {
          for (int ks = 0; ks < ksteps; ++ks) umma_tf32_ss(d, d0 + ks * 2, d1 + ks * 2, idesc, (kb | ks) ? 1u : 0u);
          umma_commit(&empty[s]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&acc_full[slot]);
      __syncwarp();
    }
  } else if (warp == 10) {
    // ------------------------------------------------------------ norm exchange (one warp, off the store path).
    // publish(k): this item's tile sum goes out as ONE 64-bit word {launch tag | f32 bits} into slot t of its image —
    //             value and validity travel together, so no fence, no counter and no reset are needed;
    // resolve(k): poll the image's ipi slots (one coalesced load per poll) until every tag is this launch's, reduce the
    //             values with a fixed shuffle tree (every CTA derives the identical norm), hand 1/||z|| to the epilogue.
    // publish(k+1) precedes resolve(k): the cross-CTA latency hides behind one whole tile of stores.
    auto publish = [&](int k) {
      const int it = sched[k];
      const int b = it / ipi, t = it - b * ipi;
      int bi, bj;
      gf_item(t, a.nblk, bi, bj);
      const int slot = k & 3;
      mbar_wait(&sum_ready[slot], (k >> 2) & 1);
      float v = lane < 8 ? sum_part[slot * 8 + lane] : 0.f;
      v = warp_sum(v);
      if (lane == 0) {
        const float tot = (bi != bj) ? 2.f * v : v;
        const unsigned long long w = ((unsigned long long)a.tag << 32) | (unsigned long long)__float_as_uint(tot);
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(a.slots + (size_t)b * GF_SLOTS + t), "l"(w) : "memory");
      }
      __syncwarp();
    };
    if (n_my > 0) publish(0);
    for (int k = 0; k < n_my; ++k) {
      if (k + 1 < n_my) publish(k + 1);
      const int it = sched[k];
      const int b = it / ipi, t = it - b * ipi;
      const unsigned long long* ps = a.slots + (size_t)b * GF_SLOTS;
      unsigned long long w = 0;
      int spins = 0;
      bool have = false;
      for (; spins < a.poll_limit; ++spins) {
        bool ok = true;
        if (lane < ipi) {
          asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(ps + lane) : "memory");
          ok = (unsigned int)(w >> 32) == a.tag;
        }
        if (__all_sync(0xffffffffu, ok) || (a.dbg & 4)) { have = true; break; }
      }
      float g;
      if (have) {
        g = lane < ipi ? __uint_as_float((unsigned int)w) : 0.f;
        g = warp_sum(g);
      } else {
        // The peers of this image have not published (not resident, or the slots were reused by a concurrent call): do not
        // wait for them.  sum_ij G_ij = sum_p (sum_c x_cp)^2, with x as the tensor core sees it (low 13 mantissa bits dropped).
        const float* xb = a.X + (size_t)b * C * a.HW;
        g = 0.f;
        for (int p0 = 0; p0 < a.HW; p0 += 32) {
          const int p = p0 + lane;
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
          if (p < a.HW) {
            for (int c = 0; c < C; c += 4) {
              s0 += __uint_as_float(__float_as_uint(__ldg(xb + (size_t)(c + 0) * a.HW + p)) & 0xffffe000u);
              s1 += __uint_as_float(__float_as_uint(__ldg(xb + (size_t)(c + 1) * a.HW + p)) & 0xffffe000u);
              s2 += __uint_as_float(__float_as_uint(__ldg(xb + (size_t)(c + 2) * a.HW + p)) & 0xffffe000u);
              s3 += __uint_as_float(__float_as_uint(__ldg(xb + (size_t)(c + 3) * a.HW + p)) & 0xffffe000u);
            }
          }
          const float sp = (s0 + s1) + (s2 + s3);
          g = fmaf(sp, sp, g);
        }
        g = warp_sum(g);
      }
      if (lane == 0) {
        const float nrm = sqrtf(g * a.inv_hw + (float)C * (float)C * a.eps);
        const float inn = 1.f / fmaxf(nrm, 1e-12f);
        inv_box[k & 3] = inn;
        mbar_arrive(&norm_ready[k & 3]);
        if (t == 0 && a.inv_norm) a.inv_norm[b] = inn;
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ epilogue: 8 warps = 2 groups of 4; group h owns accumulator
    // columns [64h, 64h+64) (two 32-column chunks), warp q of a group the TMEM lane quarter q.
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    const int r = q * 32 + lane;       // accumulator row = row of block bi held by this thread
    const bool group_leader = (q == 0 && lane == 0);
    const size_t CC = (size_t)C * C;

    // tile sum of local item k -> sum_part[k&3] (the accumulator stays in TMEM for the store pass)
    auto tile_sum = [&](int k) {
      const int slot = k & 3;
      if (!(a.dbg & 8)) mbar_wait(&acc_full[slot], (k >> 2) & 1);
      tc_fence_after();
      if (tr && k < 4 && threadIdx.x == 64) tr[2 + k] = gtimer();
      float sum = 0.f;
#pragma unroll 1
      for (int c = 2 * h; c < 2 * h + 2; ++c) {
        float v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + slot * 128 + c * 32, v);
        tmem_ld_wait();
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 32; ++j) s4[j & 3] += v[j];
        sum += (s4[0] + s4[1]) + (s4[2] + s4[3]);
      }
      sum = warp_sum(sum);
      if (lane == 0) {
        sum_part[slot * 8 + (warp - 2)] = sum;
        mbar_arrive(&sum_ready[slot]);
      }
    };

    if (n_my > 0) tile_sum(0);
    for (int k = 0; k < n_my; ++k) {
      if (k + 1 < n_my) tile_sum(k + 1);
      const int it = sched[k];
      const int b = it / ipi, t = it - b * ipi;
      int bi, bj;
      gf_item(t, a.nblk, bi, bj);
      const int slot = k & 3;
      const bool off = (bi != bj) && !(a.dbg & 2);
      if (off) {     // staging buffers of this group: the TMA stores of the previous off-diagonal item have drained them
        if (group_leader) bulk_wait_read<0>();
        asm volatile("bar.sync %0, 128;" ::"r"(2 + h) : "memory");
      }
      mbar_wait(&norm_ready[slot], (k >> 2) & 1);
      const float inv_norm = inv_box[slot];
      if (tr && k < 4 && threadIdx.x == 64) tr[6 + k] = gtimer();
#pragma unroll 1
      for (int c = 2 * h; c < 2 * h + 2; ++c) {
        float v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + slot * 128 + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = tf32_round(fast_sqrt(fmaf(v[j], a.inv_hw, a.eps)) * inv_norm);
        // block (bj, bi): transposed — lanes run along a row of Y
        float* y = a.Y + (size_t)b * CC + (size_t)(bj * 128 + c * 32) * C + bi * 128 + r;
        if (a.dbg & 1) {
          float keep = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) keep += v[j];
          if (keep == 123.456f) y[0] = keep;
        } else if (a.store_mode == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) __stcs(y + (size_t)j * C, v[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) y[(size_t)j * C] = v[j];
        }
        if (off) {   // block (bi, bj): row-major via swizzled smem, one TMA store per 128 x 32 box
          uint8_t* row = outbuf + (2 * h + (c & 1)) * GF_OUT_BYTES + r * 128;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4)
            *reinterpret_cast<float4*>(row + ((j4 ^ (r & 7)) << 4)) =
                make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[slot]);     // 8 warp arrivals free the accumulator slot
      if (off) {
        fence_proxy_async();
        asm volatile("bar.sync %0, 128;" ::"r"(2 + h) : "memory");
        if (group_leader) {
          tma_store_3d(&tmY, outbuf + (2 * h) * GF_OUT_BYTES, bj * 128 + 2 * h * 32, bi * 128, b);
          tma_store_3d(&tmY, outbuf + (2 * h + 1) * GF_OUT_BYTES, bj * 128 + (2 * h + 1) * 32, bi * 128, b);
          bulk_commit();
        }
      }
      if (tr && k < 4 && threadIdx.x == 64) tr[10 + k] = gtimer();
    }
    if (group_leader) bulk_wait_all();
    if (tr && threadIdx.x == 64) tr[14] = gtimer();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static int make_y_map(CUtensorMap* tm, const float* Y, int B, int C) {
  uint64_t dims[3] = {(uint64_t)C, (uint64_t)C, (uint64_t)B};
  uint64_t strides[2] = {(uint64_t)C * 4, (uint64_t)C * C * 4};
  uint32_t box[3] = {32, 128, 1};
  return make_tmap(tm, Y, 3, dims, strides, box);
}


static unsigned long long* g_tiles_trace = nullptr;

static int tiles_grid(int total_items) {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return total_items < sms ? total_items : sms;
}

static int tiles_env(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// x [B,C,HW] -> y [B,C*C] (C % 128 == 0, C/128 tiles-per-image (nblk(nblk+1)/2) <= GF_SLOTS); inv_norm [B] receives 1/||z||.
// Returns HK_ERR_UNSUPPORTED when C has more tiles per image than the slot table holds (caller: two-kernel path).
int bcnn_tiles_fwd(const CUtensorMap& tmX_unused, const float* x, float* y, float* inv_norm, int B, int C, int HW,
                   float inv_hw, cudaStream_t stream) {
  (void)tmX_unused;
  const int nblk = C / 128, ipi = nblk * (nblk + 1) / 2;
  if (ipi > GF_SLOTS) return set_error(HK_ERR_UNSUPPORTED, "bcnn_tiles_fwd: C=%d has more than %d tiles per image", C, GF_SLOTS);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(bcnn_gram_fwd_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, GF_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(bcnn_gram_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GF_SMEM);
    if (e != cudaSuccess) return set_error((int)e, "cudaFuncSetAttribute(bcnn_gram_fwd): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  static int pdl = -1, poll = -1, stages = -1, bal = -2, dbg = -1;
  if (pdl < 0) {
    pdl = tiles_env("HK_K1_PDL", 1);
    poll = tiles_env("HK_K1_POLL_LIMIT", 4096);
    stages = tiles_env("HK_K1_STAGES", GF_STAGES);
    if (stages < 2) stages = 2;
    if (stages > GF_STAGES) stages = GF_STAGES;
    bal = tiles_env("HK_K1_BALANCE", -1);
    dbg = tiles_env("HK_K1_DBG", 0);
  }
  GfArgs g = {};
  g.C = C; g.HW = HW; g.nblk = nblk; g.inv_hw = inv_hw; g.eps = 1e-5f; g.store_mode = 1; g.x_hint = 1;
  g.dbg = dbg; g.trace = g_tiles_trace; g.stages = stages; g.pdl = pdl; g.poll_limit = poll;
  int r;
  for (int b0 = 0; b0 < B; b0 += GRAM_CNT_MAXB) {
    const int nb = B - b0 < GRAM_CNT_MAXB ? B - b0 : GRAM_CNT_MAXB;
    CUtensorMap tmx, tmy;
    {
      uint64_t dims[3] = {(uint64_t)HW, (uint64_t)C, (uint64_t)nb};
      uint64_t strides[2] = {(uint64_t)HW * 4, (uint64_t)C * HW * 4};
      uint32_t box[3] = {32, 128, 1};
      if ((r = make_tmap(&tmx, x + (size_t)b0 * C * HW, 3, dims, strides, box))) return r;
    }
    if ((r = make_y_map(&tmy, y + (size_t)b0 * C * C, nb, C))) return r;
    g.B = nb;
    g.Y = y + (size_t)b0 * C * C;
    g.X = x + (size_t)b0 * C * HW;
    g.inv_norm = inv_norm + b0;
    g.slots = gram_slots(&g.tag);
    HK_REQUIRE(g.slots, HK_ERR_DRIVER, "bcnn_tiles_fwd: slot symbol not resolvable");
    const int grid = tiles_grid(nb * ipi);
    // units-balanced schedule only while the launch is a few waves long; long launches keep the plain image-major
    // round-robin, which holds the tiles of an image closer in time
    g.balance = bal >= 0 ? bal : (nb * ipi <= 3 * grid ? 1 : 0);
    HK_REQUIRE((nb * ipi + grid - 1) / grid + 8 <= GF_MAX_ITEMS && nb * ipi <= 65535, HK_ERR_UNSUPPORTED,
               "bcnn_tiles_fwd: item schedule does not fit (B=%d C=%d)", nb, C);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(GF_THREADS);
    cfg.dynamicSmemBytes = GF_SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g.pdl ? 1 : 0;
    cudaError_t le = (C == 512) ? cudaLaunchKernelEx(&cfg, bcnn_gram_fwd_kernel<512>, tmx, tmy, g)
                                : cudaLaunchKernelEx(&cfg, bcnn_gram_fwd_kernel<0>, tmx, tmy, g);
    if (le != cudaSuccess) return set_error((int)le, "cudaLaunchKernelEx(bcnn_gram_fwd): %s", cudaGetErrorString(le));
    HK_LAUNCH_CHECK("bcnn_gram_fwd_kernel");
  }
  return 0;
}

}  // namespace hk
