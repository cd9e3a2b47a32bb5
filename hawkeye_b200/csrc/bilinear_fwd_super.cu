// K1 (super-tile kernel) — fused bilinear pooling forward (reference model/methods/BCNN.py:13-27), C = 512:
//     G = X X^T / HW ; z = sqrt(G + 1e-5) ; y = z / max(||z||_2, 1e-12)            X: [B, 512, HW]  ->  y: [B, 512*512]
//
// Why another K1.  The 128x128-tile kernel (bilinear_fwd_tiles.cu) loads two 128-row blocks of X per tile: 2.0 MB of L2->SM
// traffic per image for 1.0 MB of output, and the ncu captures of round 1/2 show it bound by that ingest (the SM-side fabric,
// ~6.9 TB/s chip-wide), not by HBM.  Here the ten unique tiles of an image are grouped into FOUR items that share operands:
//     item 0 (D): blocks {0,1}   -> tiles (0,0) (0,1) (1,1)     loads X0 X1      (200 KB)   writes 4 output blocks
//     item 1 (D): blocks {2,3}   -> tiles (2,2) (2,3) (3,3)     loads X2 X3      (200 KB)   writes 4 output blocks
//     item 2 (O): row block 0    -> tiles (0,2) (0,3)           loads X0 X2 X3   (300 KB)   writes 4 output blocks
//     item 3 (O): row block 1    -> tiles (1,2) (1,3)           loads X1 X2 X3   (300 KB)   writes 4 output blocks
// 1.0 MB of ingest per image (half), every item stores the same 256 KB, and B = 32 is ONE wave of 128 items.
// Tiles that share the A block and have adjacent B blocks are one N = 256 MMA per k-step (the operand slots of a stage
// are contiguous in shared memory, so a 256-row K-major B operand simply runs on into the next slot).
//
// TMEM: four 128-column slots.  O items take a slot pair, alternating (0,1) / (2,3): two items in flight, so the tile sums
// of item k+1 are published before item k is stored (the norm exchange hides behind the stores, as in the tile kernel).
// D items need three slots and alternate {wide (0,1), single 2} / {wide (2,3), single 0}; the single tile and the diagonal
// tile of the pair are stored first, so the next item's MMAs start half-way through the previous item's stores.
// The grid is a multiple of 4, so a CTA sees a single item type and the four items of an image sit on CTAs 4m..4m+3.
//
// Norm exchange (||z||^2 = sum_ij G_ij / HW + C^2 eps needs all four items of the image), two builds of the same kernel:
//   CL = true : the four CTAs are one thread-block cluster and exchange their item sums through distributed shared memory
//               (st.shared::cluster + remote mbarrier arrive): ~0.3 us, never crosses L2 (the global-memory exchange was
//               measured to take 2-4 us once other CTAs' stores flood the L2 queues), no library-owned state, and
//               co-residency is guaranteed by the cluster launch — nothing to time out;
//   CL = false: the tile kernel's exchange — launch-tagged 64-bit words in a library-owned table, bounded polling and a
//               local closed-form fallback (sum_p (sum_c x_cp)^2 / HW + C^2 eps): an optimisation, never a dependency.
//               Used when clusters of four cannot be scheduled and for batches where 148 plain CTAs beat 132 clustered ones.
#include <stdlib.h>

#include <atomic>

#include "common.cuh"
#include "host.h"
#include "../../include/hawkeye_b200.h"

namespace hk {

namespace {

__device__ __forceinline__ float fast_sqrt_s(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

constexpr int SP_C = 512;
constexpr int SP_SLOT = 128 * 128;                 // 16 KB: 128 rows x 32 fp32 (one 128B-swizzled K-major box)
constexpr int SP_RING_BYTES = 10 * SP_SLOT;        // operand ring: 5 stages of two row blocks (D items) or 3 of three (O items)
constexpr int SP_MAX_STAGES = 5;
constexpr int SP_OUT_BYTES = 128 * 128;            // one 128-row x 32-column fp32 box
constexpr int SP_MAX_ITEMS = 192;                  // items per CTA per launch
constexpr int SP_IPI = 4;                          // items per image
constexpr int SP_SLOTS = 4;                        // tagged sums per image
constexpr int SP_REGIONS = 32;
constexpr int SP_MAXB = 4096;                      // images per launch
constexpr int SP_MAX_ITEMS_BYTES = 4 * 192;
constexpr int SP_SMEM = SP_RING_BYTES + 4 * SP_OUT_BYTES + 1024 + 768 + SP_MAX_ITEMS_BYTES;
#ifndef HK_SP_EPI_WARPS
#define HK_SP_EPI_WARPS 16
#endif
constexpr int SP_EPI = HK_SP_EPI_WARPS;            // epilogue warps: 8 (two 32-column chunks of a tile each) or 16 (one each)
constexpr int SP_CPW = SP_EPI == 8 ? 2 : 1;        // 32-column chunks of a tile per epilogue warp
static_assert(SP_EPI == 8 || SP_EPI == 16, "8 or 16 epilogue warps");
constexpr int SP_NORM_WARP = 2 + SP_EPI;
constexpr int SP_THREADS = 32 * (3 + SP_EPI);      // warp 0 TMA, warp 1 MMA, warps 2..2+SP_EPI-1 epilogue, last warp norm exchange

__device__ unsigned long long g_super_slots[(size_t)SP_REGIONS * SP_MAXB * SP_SLOTS];

unsigned long long* super_slots(unsigned int* tag) {
  static std::atomic<unsigned> next{0};
  static thread_local int dev_cached = -1;
  static thread_local unsigned long long* base = nullptr;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != dev_cached) {
    void* p = nullptr;
    if (cudaGetSymbolAddress(&p, g_super_slots) != cudaSuccess) return nullptr;
    base = static_cast<unsigned long long*>(p);
    dev_cached = dev;
  }
  const unsigned n = next.fetch_add(1);
  *tag = n + 1 ? n + 1 : 1;     // never 0 (the initial slot contents)
  return base + (size_t)(n % SP_REGIONS) * SP_MAXB * SP_SLOTS;
}

struct SpArgs {
  int B, HW;
  float inv_hw, eps;
  float* Y;
  float* inv_norm;
  unsigned long long* slots;   // [B][SP_SLOTS] tagged item sums {tag:32 | f32 bits:32}
  unsigned int tag;
  const float* X;              // read directly only by the norm fallback
  int poll_limit;
  int pdl;
  int dbg;                     // profiling only: 1 no direct stores, 2 no TMA stores, 4 no norm exchange
  int l2_prefetch;             // 1: cp.async.bulk.prefetch.tensor of the next item's operand boxes
  unsigned long long* trace;   // profiling only: [grid][16] %globaltimer stamps, or null
};
__device__ __forceinline__ unsigned long long gtimer_s() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// schedule entry of a local item: it (16) | w (2) << 16 | n (2) << 18 | parity(w) << 20 | parity(w+1) << 21 | parity(n) << 22
//   it = b * 4 + t;  w = first TMEM slot of the wide accumulator (0 or 2);  n = TMEM slot of the D item's third tile (2 / 3)
struct ItemInfo {
  int b, t, type;          // type 0 = D (diagonal super-tile), 1 = O (half of the off-diagonal super-tile)
  int blk[3];              // row blocks of X in operand slots 0..2
  int nsl;                 // operand slots used
  int ntiles;              // 3 (D) or 2 (O)
  int w, n;                // TMEM slots
  uint32_t pw0, pw1, pn;   // use parities of those slots before this item
};
__device__ __forceinline__ ItemInfo decode_item(uint32_t e) {
  ItemInfo i;
  const int it = e & 0xffff;
  i.b = it >> 2;
  i.t = it & 3;
  i.w = (e >> 16) & 3;
  i.n = (e >> 18) & 3;
  i.pw0 = (e >> 20) & 1; i.pw1 = (e >> 21) & 1; i.pn = (e >> 22) & 1;
  if (i.t < 2) {
    i.type = 0; i.blk[0] = 2 * i.t; i.blk[1] = 2 * i.t + 1; i.blk[2] = 0; i.nsl = 2; i.ntiles = 3;
  } else {
    i.type = 1; i.blk[0] = i.t - 2; i.blk[1] = 2; i.blk[2] = 3; i.nsl = 3; i.ntiles = 2;
  }
  return i;
}
// tile q of an item: accumulator slot, its use parity, and the (bi, bj) output block it holds (rows = lanes = block bi)
__device__ __forceinline__ void item_tile(const ItemInfo& i, int q, int& slot, uint32_t& par, int& bi, int& bj) {
  if (i.type == 0) {     // the single-slot tile first, then the diagonal tile of the wide pair: those two slots are what the
                         // next D item is waiting for (see the slot rotation in the scheduler)
    if (q == 0) { slot = i.n; par = i.pn; bi = i.blk[1]; bj = i.blk[1]; }
    else if (q == 1) { slot = i.w; par = i.pw0; bi = i.blk[0]; bj = i.blk[0]; }
    else { slot = i.w + 1; par = i.pw1; bi = i.blk[0]; bj = i.blk[1]; }
  } else {
    if (q == 0) { slot = i.w; par = i.pw0; bi = i.blk[0]; bj = 2; }
    else { slot = i.w + 1; par = i.pw1; bi = i.blk[0]; bj = 3; }
  }
}

__device__ __forceinline__ void tma_store_3d_s(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(m),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_s() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0_s() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all_s() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void cluster_sync_all_s() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store a float into the same-offset shared-memory word of CTA `cta` of this cluster, then arrive (release) on its barrier
__device__ __forceinline__ void dsmem_send(float* word, uint64_t* bar, uint32_t cta, float v) {
  asm volatile(
      "{\n"
      ".reg .b32 ra, rb;\n"
      "mapa.shared::cluster.u32 ra, %0, %2;\n"
      "mapa.shared::cluster.u32 rb, %1, %2;\n"
      "st.shared::cluster.f32 [ra], %3;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [rb];\n"
      "}\n" ::"r"(smem_u32(word)),
      "r"(smem_u32(bar)), "r"(cta), "f"(v)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster_s(uint64_t* bar, uint32_t parity) {
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

template <bool CL>
__global__ void __launch_bounds__(SP_THREADS, 1)
bcnn_super_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, SpArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* outbuf = smem + SP_RING_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(outbuf + 4 * SP_OUT_BYTES);
  uint64_t* empty = full + SP_MAX_STAGES;
  uint64_t* acc_full = empty + SP_MAX_STAGES;      // [4]  MMA -> epilogue, per TMEM slot
  uint64_t* acc_empty = acc_full + 4;          // [4]  epilogue warps -> MMA
  uint64_t* sum_ready = acc_empty + 4;         // [4]  epilogue warps -> norm warp: item sums of item k in sum_part[k&3]
  uint64_t* norm_ready = sum_ready + 4;        // [4]  norm warp -> epilogue: inv_norm of item k in inv_box[k&3]
  uint64_t* peer_ready = norm_ready + 4;       // [4]  CL: the four CTAs of the cluster -> norm warp (item sums in peer_sums)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(peer_ready + 4);
  float* sum_part = reinterpret_cast<float*>(tmem_slot + 2);   // [4][SP_EPI]
  float* inv_box = sum_part + 4 * SP_EPI;                              // [4]
  float* peer_sums = inv_box + 4;                              // [4][4]  CL: item sums of the cluster's CTAs, by rank
  int* n_my_box = reinterpret_cast<int*>(peer_sums + 16);
  uint32_t* sched = reinterpret_cast<uint32_t*>(n_my_box + 1);  // [SP_MAX_ITEMS]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int C = SP_C;
  const int total_items = a.B * SP_IPI;
  const int nk = (a.HW + 31) / 32;
  // the grid is a multiple of 4 CTAs, so this CTA only ever sees one item type: D items (two operand blocks per stage) get
  // a 5-deep ring, O items (three blocks) a 3-deep one
  const int my_nsl = (blockIdx.x & 3) < 2 ? 2 : 3;
  const int stage_bytes = my_nsl * SP_SLOT;
  const int nstages = my_nsl == 2 ? 5 : 3;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
    for (int s = 0; s < SP_MAX_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], SP_EPI);
      mbar_init(&sum_ready[s], SP_EPI); mbar_init(&norm_ready[s], 1);
      mbar_init(&peer_ready[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2 && lane == 0) {
    // items of this CTA (image-major round-robin: the four items of an image run at the same time on four CTAs, so
    // dependencies only ever point to earlier images) and their TMEM slots
    uint32_t uses[4] = {0, 0, 0, 0};
    int o_toggle = 0, n_toggle = 0, n = 0;
    for (int it = blockIdx.x; it < total_items && n < SP_MAX_ITEMS; it += gridDim.x) {
      const int t = it & 3;
      uint32_t w, nn = 0;
      // D items alternate between {wide (0,1), single 2} and {wide (2,3), single 0}: the next D item needs the previous one's
      // single slot and first wide slot, which are stored first, so its MMAs start half-way through the previous stores
      if (t < 2) { w = n_toggle ? 2 : 0; nn = n_toggle ? 0 : 2; n_toggle ^= 1; }
      else { w = o_toggle ? 2 : 0; o_toggle ^= 1; }
      uint32_t e = (uint32_t)it | (w << 16) | (nn << 18) | ((uses[w] & 1) << 20) | ((uses[w + 1] & 1) << 21);
      ++uses[w]; ++uses[w + 1];
      if (t < 2) { e |= (uses[nn] & 1) << 22; ++uses[nn]; }
      sched[n++] = e;
    }
    *n_my_box = n;
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  if (CL) cluster_sync_all_s();      // every CTA's barriers exist before a peer arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_my = *n_my_box;
  if (a.pdl) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  unsigned long long* tr = a.trace ? a.trace + (size_t)blockIdx.x * 16 : nullptr;
  if (tr && threadIdx.x == 0) tr[0] = gtimer_s();

  if (warp == 0) {
    if (lane == 0) {
      uint64_t policy;
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
      int kbg = 0;
      // L2 prefetch runs one item ahead of the shared-memory ring: the ring holds <= 160 KB, a DRAM round trip per refill
      // would make the load phase a latency chain; with the boxes already in L2 the refills are L2 hits
      auto prefetch_item = [&](int k) {
        const ItemInfo it = decode_item(sched[k]);
        for (int kb = 0; kb < nk; ++kb)
          for (int q = 0; q < it.nsl; ++q) tma_prefetch_l2_3d(&tmX, kb * 32, it.blk[q] * 128, it.b);
      };
      if (n_my > 0 && a.l2_prefetch) prefetch_item(0);
      for (int k = 0; k < n_my; ++k) {
        const ItemInfo it = decode_item(sched[k]);
        for (int kb = 0; kb < nk; ++kb, ++kbg) {
          if (kb == (nk > 2 ? 2 : nk - 1) && k + 1 < n_my && a.l2_prefetch) prefetch_item(k + 1);
          const int s = kbg % nstages;
          const uint32_t ph = (kbg / nstages) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], it.nsl * SP_SLOT);
          uint8_t* st = smem + s * stage_bytes;
          for (int q = 0; q < it.nsl; ++q)
            tma_load_3d_hint(st + q * SP_SLOT, &tmX, &full[s], kb * 32, it.blk[q] * 128, it.b, policy);
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc_w = make_idesc_tf32(128, 256, 0, 0);
    const uint32_t idesc_n = make_idesc_tf32(128, 128, 0, 0);
    const uint64_t desc_tmpl = make_sdesc(0, 16, 1024);
    int kbg = 0;
    for (int k = 0; k < n_my; ++k) {
      const ItemInfo it = decode_item(sched[k]);
      mbar_wait(&acc_empty[it.w], it.pw0 ^ 1);
      mbar_wait(&acc_empty[it.w + 1], it.pw1 ^ 1);
      if (it.type == 0) mbar_wait(&acc_empty[it.n], it.pn ^ 1);
      tc_fence_after();
      const uint32_t dw = tmem_base + it.w * 128;
      const uint32_t dn = tmem_base + it.n * 128;
      for (int kb = 0; kb < nk; ++kb, ++kbg) {
        const int s = kbg % nstages;
        const uint32_t ph = (kbg / nstages) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (tr && kb == 0 && k < 2 && lane == 0) tr[1 + 6 * k] = gtimer_s();
        const uint32_t s0 = smem_u32(smem + s * stage_bytes);
        const uint64_t d0 = desc_tmpl + (s0 >> 4);
        const uint64_t d1 = desc_tmpl + ((s0 + SP_SLOT) >> 4);
        const int krem = a.HW - kb * 32;
        const int ksteps = krem >= 32 ? 4 : (krem + 7) / 8;
        if (elect_one()) {
          if (it.type == 0) {
            // wide: X_b0 . [X_b0 ; X_b1]^T  (slots 0,1 contiguous) ; narrow: X_b1 . X_b1^T
            for (int ks = 0; ks < ksteps; ++ks) {
              umma_tf32_ss(dw, d0 + ks * 2, d0 + ks * 2, idesc_w, (kb | ks) ? 1u : 0u);
              umma_tf32_ss(dn, d1 + ks * 2, d1 + ks * 2, idesc_n, (kb | ks) ? 1u : 0u);
            }
          } else {
            // wide: X_a . [X_2 ; X_3]^T  (slots 1,2 contiguous)
            for (int ks = 0; ks < ksteps; ++ks) umma_tf32_ss(dw, d0 + ks * 2, d1 + ks * 2, idesc_w, (kb | ks) ? 1u : 0u);
          }
          umma_commit(&empty[s]);
        }
        __syncwarp();
      }
      if (elect_one()) {
        umma_commit(&acc_full[it.w]);
        umma_commit(&acc_full[it.w + 1]);
        if (it.type == 0) umma_commit(&acc_full[it.n]);
      }
      if (tr && k < 2 && lane == 0) tr[2 + 6 * k] = gtimer_s();
      __syncwarp();
    }
  } else if (warp == SP_NORM_WARP) {
    // ------------------------------------------------------------ norm exchange (see bilinear_fwd_tiles.cu)
    for (int k = 0; k < n_my; ++k) {
      const ItemInfo it = decode_item(sched[k]);
      const int slot = k & 3;
      mbar_wait(&sum_ready[slot], (k >> 2) & 1);
      float v = lane < SP_EPI ? sum_part[slot * SP_EPI + lane] : 0.f;
      v = warp_sum(v);
      float g;
      if (CL) {
        // lane r hands this CTA's item sum to CTA r of the cluster (same image, items 4b..4b+3) and arrives on its barrier;
        // slot reuse is safe: a peer can only be at item k+4 after this CTA has published (hence consumed) items k+1..k+3
        if (lane < 4) dsmem_send(&peer_sums[slot * 4 + (blockIdx.x & 3)], &peer_ready[slot], (uint32_t)lane, v);
        mbar_wait_cluster_s(&peer_ready[slot], (k >> 2) & 1);
        g = (peer_sums[slot * 4 + 0] + peer_sums[slot * 4 + 1]) + (peer_sums[slot * 4 + 2] + peer_sums[slot * 4 + 3]);
      } else {
        if (lane == 0) {
          const unsigned long long w = ((unsigned long long)a.tag << 32) | (unsigned long long)__float_as_uint(v);
          asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(a.slots + (size_t)it.b * SP_SLOTS + it.t), "l"(w) : "memory");
        }
        __syncwarp();
        const unsigned long long* ps = a.slots + (size_t)it.b * SP_SLOTS;
        unsigned long long w = 0;
        bool have = false;
        for (int spins = 0; spins < a.poll_limit; ++spins) {
          bool ok = true;
          if (lane < SP_IPI) {
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(ps + lane) : "memory");
            ok = (unsigned int)(w >> 32) == a.tag;
          }
          if (__all_sync(0xffffffffu, ok) || (a.dbg & 4)) { have = true; break; }
        }
        if (have) {
          g = lane < SP_IPI ? __uint_as_float((unsigned int)w) : 0.f;
          g = warp_sum(g);
        } else {
          // peers not resident / slots reused by a concurrent call: closed form from X, as the tensor core sees it
          const float* xb = a.X + (size_t)it.b * C * a.HW;
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
      }
      if (lane == 0) {
        const float nrm = sqrtf(g * a.inv_hw + (float)C * (float)C * a.eps);
        const float inn = 1.f / fmaxf(nrm, 1e-12f);
        inv_box[slot] = inn;
        mbar_arrive(&norm_ready[slot]);
        if (it.t == 0 && a.inv_norm) a.inv_norm[it.b] = inn;
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ epilogue: SP_EPI warps = groups of 4; group h owns SP_CPW
    // 32-column chunks of every tile, warp q of a group the TMEM lane quarter q.  (16 warps: the store phase is bound by the
    // serial tmem_ld -> math -> store chain of each warp, not by bandwidth — twice the warps, half the time.)
    const int q4 = warp & 3;
    const int h = (warp - 2) >> 2;
    const int r = q4 * 32 + lane;       // accumulator row = row of block bi held by this thread
    const bool group_leader = (q4 == 0 && lane == 0);
    const size_t CC = (size_t)C * C;

    // weighted sum of the tiles of local item k -> sum_part[k&3] (the accumulators stay in TMEM for the store pass)
    auto item_sum = [&](int k) {
      const ItemInfo it = decode_item(sched[k]);
      float sum = 0.f;
      for (int q = 0; q < it.ntiles; ++q) {
        int slot, bi, bj;
        uint32_t par;
        item_tile(it, q, slot, par, bi, bj);
        mbar_wait(&acc_full[slot], par);
        tc_fence_after();
        if (tr && q == 0 && k < 2 && threadIdx.x == 64) tr[3 + 6 * k] = gtimer_s();
        float ts = 0.f;
#pragma unroll 1
        for (int c = SP_CPW * h; c < SP_CPW * h + SP_CPW; ++c) {
          float v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + slot * 128 + c * 32, v);
          tmem_ld_wait();
          float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 32; ++j) s4[j & 3] += v[j];
          ts += (s4[0] + s4[1]) + (s4[2] + s4[3]);
        }
        sum += (bi != bj) ? 2.f * ts : ts;
      }
      sum = warp_sum(sum);
      if (lane == 0) {
        sum_part[(k & 3) * SP_EPI + (warp - 2)] = sum;
        mbar_arrive(&sum_ready[k & 3]);
      }
      if (tr && k < 2 && threadIdx.x == 64) tr[4 + 6 * k] = gtimer_s();
    };
    auto is_o = [&](int k) { return ((sched[k] & 3u) >= 2u); };

    bool summed_next = false;     // item k's sums were already published while item k-1 was stored
    for (int k = 0; k < n_my; ++k) {
      if (!summed_next) item_sum(k);
      // O -> O: disjoint TMEM slot pairs, so the next item's sums go out before this item's stores
      summed_next = (k + 1 < n_my) && is_o(k) && is_o(k + 1);
      if (summed_next) item_sum(k + 1);
      const ItemInfo it = decode_item(sched[k]);
      mbar_wait(&norm_ready[k & 3], (k >> 2) & 1);
      const float inv_norm = inv_box[k & 3];
      if (tr && k < 2 && threadIdx.x == 64) tr[5 + 6 * k] = gtimer_s();
      for (int q = 0; q < it.ntiles; ++q) {
        int slot, bi, bj;
        uint32_t par;
        item_tile(it, q, slot, par, bi, bj);
        const bool off = (bi != bj) && !(a.dbg & 2);
        if (off) {     // staging buffers of this group: the TMA stores of the previous off-diagonal tile have drained them
          if (group_leader) bulk_wait_read0_s();
          asm volatile("bar.sync %0, 128;" ::"r"(2 + h) : "memory");
        }
#pragma unroll 1
        for (int c = SP_CPW * h; c < SP_CPW * h + SP_CPW; ++c) {
          float v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + slot * 128 + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = tf32_round(fast_sqrt_s(fmaf(v[j], a.inv_hw, a.eps)) * inv_norm);
          // block (bj, bi): transposed — lanes run along a row of Y
          float* y = a.Y + (size_t)it.b * CC + (size_t)(bj * 128 + c * 32) * C + bi * 128 + r;
          if (!(a.dbg & 1)) {
#pragma unroll
            for (int j = 0; j < 32; ++j) y[(size_t)j * C] = v[j];
          }
          if (off) {   // block (bi, bj): row-major via swizzled smem, one TMA store per 128 x 32 box
            uint8_t* row = outbuf + c * SP_OUT_BYTES + r * 128;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
              *reinterpret_cast<float4*>(row + ((j4 ^ (r & 7)) << 4)) =
                  make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[slot]);     // SP_EPI warp arrivals free the accumulator slot
        if (off) {
          fence_proxy_async();
          asm volatile("bar.sync %0, 128;" ::"r"(2 + h) : "memory");
          if (group_leader) {
            for (int c = SP_CPW * h; c < SP_CPW * h + SP_CPW; ++c)
              tma_store_3d_s(&tmY, outbuf + c * SP_OUT_BYTES, bj * 128 + c * 32, bi * 128, it.b);
            bulk_commit_s();
          }
        }
        if (tr && k < 2 && q == it.ntiles - 1 && threadIdx.x == 64) tr[6 + 6 * k] = gtimer_s();
      }
    }
    if (group_leader) bulk_wait_all_s();
    if (tr && threadIdx.x == 64) tr[13] = gtimer_s();
  }
  tc_fence_before();
  __syncthreads();
  if (CL) cluster_sync_all_s();      // no peer is still writing into this CTA's shared memory
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

unsigned long long* g_super_trace = nullptr;

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

}  // namespace

void set_super_trace(void* buf) { g_super_trace = static_cast<unsigned long long*>(buf); }

static int g_super_max_clusters = -1;
// one-time host-side setup (function attributes + cluster occupancy query), kept out of the launch path
int bcnn_super_prepare() {
  if (g_super_max_clusters >= 0) return 0;
  cudaError_t e = cudaFuncSetAttribute(bcnn_super_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(bcnn_super_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SMEM);
  if (e != cudaSuccess) return set_error((int)e, "cudaFuncSetAttribute(bcnn_super_fwd): %s", cudaGetErrorString(e));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(SP_THREADS);
  cfg.dynamicSmemBytes = SP_SMEM;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int sms = 148, dev = 0, n = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cfg.gridDim = dim3(sms & ~3);
  e = cudaOccupancyMaxActiveClusters(&n, bcnn_super_fwd_kernel<true>, &cfg);
  if (e != cudaSuccess || n <= 0) {
    (void)cudaGetLastError();
    n = 0;
  }
  g_super_max_clusters = n;
  return 0;
}

bool bcnn_super_one_wave(int B) { return bcnn_super_prepare() == 0 && g_super_max_clusters > 0 && B <= g_super_max_clusters; }

// x [B,512,HW] -> y [B,512*512]; inv_norm [B] receives 1/||z||.  HK_ERR_UNSUPPORTED for other C (caller: tile kernel).
int bcnn_super_fwd(const float* x, float* y, float* inv_norm, int B, int C, int HW, float inv_hw, cudaStream_t stream) {
  if (C != SP_C) return set_error(HK_ERR_UNSUPPORTED, "bcnn_super_fwd: C=%d (only 512)", C);
  int r = bcnn_super_prepare();
  if (r) return r;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  static int pdl = -1, poll = -1, dbg = -1, l2p = 0, clmode = -1;
  if (pdl < 0) {
    l2p = env_int("HK_K1_L2PREFETCH", 0);     // measured: slower (r3 log) — the prefetches queue ahead of the loads
    poll = env_int("HK_K1_POLL_LIMIT", 4096);
    dbg = env_int("HK_K1_DBG", 0);
    clmode = env_int("HK_K1_SUPER_CL", -1);   // 1: always clusters, 0: never, -1: clusters while the launch is <= 2 waves of them
    pdl = env_int("HK_K1_PDL", 1);
  }
  SpArgs g = {};
  g.HW = HW; g.inv_hw = inv_hw; g.eps = 1e-5f; g.pdl = pdl; g.poll_limit = poll; g.dbg = dbg; g.l2_prefetch = l2p; g.trace = g_super_trace;
  for (int b0 = 0; b0 < B; b0 += SP_MAXB) {
    const int nb = B - b0 < SP_MAXB ? B - b0 : SP_MAXB;
    CUtensorMap tmx, tmy;
    {
      uint64_t dims[3] = {(uint64_t)HW, (uint64_t)C, (uint64_t)nb};
      uint64_t strides[2] = {(uint64_t)HW * 4, (uint64_t)C * HW * 4};
      uint32_t box[3] = {32, 128, 1};
      if ((r = make_tmap(&tmx, x + (size_t)b0 * C * HW, 3, dims, strides, box))) return r;
    }
    {
      uint64_t dims[3] = {(uint64_t)C, (uint64_t)C, (uint64_t)nb};
      uint64_t strides[2] = {(uint64_t)C * 4, (uint64_t)C * C * 4};
      uint32_t box[3] = {32, 128, 1};
      if ((r = make_tmap(&tmy, y + (size_t)b0 * C * C, 3, dims, strides, box))) return r;
    }
    g.B = nb;
    g.Y = y + (size_t)b0 * C * C;
    g.X = x + (size_t)b0 * C * HW;
    g.inv_norm = inv_norm ? inv_norm + b0 : nullptr;
    const int items = nb * SP_IPI;                            // always a multiple of 4
    const int mc = g_super_max_clusters;
    const bool cl = mc > 0 && (clmode == 1 || (clmode < 0 && items <= 2 * 4 * mc));
    int grid;
    if (cl) {
      grid = items < 4 * mc ? items : 4 * mc;
    } else {
      grid = items < sms ? items : (sms & ~3);
      g.slots = super_slots(&g.tag);
      HK_REQUIRE(g.slots, HK_ERR_DRIVER, "bcnn_super_fwd: slot symbol not resolvable");
    }
    HK_REQUIRE((items + grid - 1) / grid <= SP_MAX_ITEMS, HK_ERR_UNSUPPORTED, "bcnn_super_fwd: schedule does not fit (B=%d)", nb);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(SP_THREADS);
    cfg.dynamicSmemBytes = SP_SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (cl) {
      attr[na].id = cudaLaunchAttributeClusterDimension;
      attr[na].val.clusterDim.x = 4;
      attr[na].val.clusterDim.y = 1;
      attr[na].val.clusterDim.z = 1;
      ++na;
    }
    if (g.pdl) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    cudaError_t le = cl ? cudaLaunchKernelEx(&cfg, bcnn_super_fwd_kernel<true>, tmx, tmy, g)
                        : cudaLaunchKernelEx(&cfg, bcnn_super_fwd_kernel<false>, tmx, tmy, g);
    if (le != cudaSuccess) return set_error((int)le, "cudaLaunchKernelEx(bcnn_super_fwd): %s", cudaGetErrorString(le));
    HK_LAUNCH_CHECK("bcnn_super_fwd_kernel");
  }
  return 0;
}

}  // namespace hk
