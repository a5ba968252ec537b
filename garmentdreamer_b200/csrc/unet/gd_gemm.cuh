// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = A[M,K] * B[N,K]^T  (fp16 in, fp32 accumulate).
//
// One CTA per 128 x BLOCK_N output tile, 192 threads, warp-specialised:
//   warp 0    : TMA producer  -- cp.async.bulk.tensor (4-D box for A, 3-D box for B, 128B swizzle)
//   warp 1    : MMA issuer    -- one elected lane issues tcgen05.mma (M=128, N=BLOCK_N, K=16);
//                                also owns the TMEM allocation
//   warps 2-5 : epilogue      -- tcgen05.ld (32 lanes x 32 columns per warp), bias / time-embedding /
//                                residual / SiLU / GEGLU / scale, fp16 stores
// smem ring of kStages {A 128x64, B BLOCK_Nx64} fp16 tiles; full/empty mbarriers between producer
// and issuer, tcgen05.commit releases slots and signals the epilogue.
//
// The A operand is addressed through a 4-D tensor map so that the same kernel runs
//   * Linear layers and attention matmuls  (box 64 x 128 x 1 x 1, batched through d2), and
//   * 3x3 / 1x1 convolutions as implicit GEMM over NHWC activations: for tap (dx,dy) the box
//     {64 ch, W, rows, images} is fetched at pixel offset (dx,dy); TMA zero-fills out-of-bounds
//     pixels, which implements the padding. K runs over (tap, channel-block).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "gd_unet.h"

namespace gdu {

constexpr int kBM = 128;
constexpr int kBK = 64;   // 64 fp16 = 128 bytes = one swizzle row
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;  // producer + issuer + epilogue warps

struct GemmKParams {
  int M, N, num_kb, kb_per_tap;
  int mode_conv;  // 0: A coords {k, m0, z, 0};  1: conv taps
  int tap_dx[9], tap_dy[9], tap_c[9];
  int rows_per_image, img_w, rows_box, imgs_box;  // conv tile = imgs_box x rows_box x box_w pixels
  int box_w, tiles_per_row;                       // box_w <= img_w; tiles_per_row = img_w / box_w
  int heads, a_head_k, a_zflat, b_head_k, b_head_n, b_zdim;
  __half* C;
  long long ldc, c_batch_stride, c_head_stride;
  const __half* bias;
  const __half* row_bias;
  long long row_bias_ld;
  const __half* residual;
  float alpha;
  unsigned flags;
  int block_n;
  int stages;
  int m_tiles, n_tiles, total_tiles;
  int stg_bufs;               // staging buffers per epilogue warp (1 or 2; 0 = no staging area)
  int bias_stride;            // floats of bias staging per epilogue warp: 32 x ceil(chunks / 2)
  int b_resident;             // CTA pairs with one N tile: the CTA's half of B (all k-blocks) is loaded once and stays in smem
  int tma_store;              // MODE 0: full 32-column chunks leave through shared memory + TMA store (tmC)
  int ksplit, kb_per_split;   // split-K: tile t covers k-blocks [ks*kb_per_split, ...) and writes fp32 partials
  float* ws;                  // [ksplit][M][N] partial sums (deterministic: summed in order by k_splitk_finalize)
  float* colstats;            // optional [ceil(M/32)][2][N]: per 32-row block and output column, sum and sum of squares of the fp16
                              // values this GEMM stores (GroupNorm statistics of the consumer without another pass over the tensor)
  int a_yscale;               // conv taps: A row coordinate = tile row * a_yscale + tap_dy (2: stride-2 convolution read through a strided
                              // VIEW of the full-resolution input, dims {2C, W/2, H, N}: no space-to-depth copy)
  int c_up2_w;                // > 0: the M rows are the pixels of a half-resolution image of this width and row m = (n, y, x) is stored at
                              // full-resolution pixel (n, 2y, 2x) -- output row 4 (m - x) + 2 x of C (the caller offsets C by the phase
                              // (py, px)): the per-phase GEMMs of a stride-2 data gradient write the upsampled tensor directly
  const float4* gn_coef;      // MODE 4 (GroupNorm-backward producer): [images][N] (ya, yb, ca, cb); p.residual = the GroupNorm input x.
                              // The epilogue stores g = acc * silu'(x*ya + yb) and colstats = per 32-row block sum g | sum g*(x*ca + cb)
  uint32_t halo_bytes;        // bytes of one haloed A slot (1024-aligned)
  int patch_w_tiles, patch_tiles_per_img;   // a_halo == 2: 8-pixel x 16-row patches per image row of patches / per image
  int a_halo;                 // 2 = images narrower than 128 pixels: an M tile is a PATCH of 8 pixels x 16 rows and the A slot holds
                              // its (8+2) x 16 haloed pixels of image rows y0+dy .. (TMA box {64 ch, 10, 16, 1}); tap dx starts dx pixels
                              // in, every image row of the patch is one 8-row UMMA group and the groups are 10 pixels (1280 B) apart
                              // (descriptor SBO = 1280): A traffic / 2.4, and output rows map to pixels patch-wise. 1 =
                              // CTA pairs, 3x3 conv on rows of >= 128 pixels: ONE haloed A tile (130 pixels x 64 ch) per (dy, channel
                              // block) serves the three dx taps through row-shifted UMMA descriptors (A traffic / 3)
};

__device__ __forceinline__ uint32_t s2u(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// One elected lane of a converged warp. Unlike `lane == 0`, ptxas knows exactly one thread is active
// behind this predicate, so operands of the async-unit instructions (UTCHMMA / UTMALDG / UTCBAR need
// uniform registers) are moved with a single R2UR instead of a per-instruction broadcast loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void bar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s2u(b)), "r"(n));
}
__device__ __forceinline__ void bar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s2u(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\t"
      "bra W_%=;\n\t"
      "D_%=:\n\t}" ::"r"(s2u(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
          "r"(s2u(dst)), "l"(map), "r"(s2u(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(s2u(dst)), "l"(map), "r"(s2u(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::
          "l"(map), "r"(s2u(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// K-major, 128B-swizzled operand tile (rows of 128 B, 8-row groups 1024 B apart): UMMA smem
// descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO 1 | SBO 1024>>4 | version 1 | SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t umma_desc_sw128_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
constexpr uint32_t kPatchBytes = 16 * 10 * 128;   // a_halo == 2: 16 image rows x 10 pixels x 128 B
constexpr uint32_t kHaloBytes = 17 * 1024;   // 130 rows x 128 B = 16640 B of haloed A, padded to the 1024-byte atom
// Instruction descriptor (cute::UMMA::InstrDescriptor): F32 accumulate, F16 x F16, K-major A and B.
__device__ __forceinline__ uint32_t umma_idesc_f16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s2u(bar)) : "memory");
}
// ---- CTA-pair (cta_group::2) primitives: two SMs of one TPC compute a 256 x BLOCK_N tile; each CTA
// stages its own 128 rows of A and HALF of the B tile, the leader (cluster rank 0) issues the MMAs.
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cta address -> the leader's copy
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
          "r"(s2u(dst)), "l"(map), "r"(s2u(bar) & kPeerMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(s2u(dst)), "l"(map), "r"(s2u(bar) & kPeerMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {   // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(s2u(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void bar_arrive_leader(uint64_t* bar) {  // arrive on the leader CTA's copy of `bar`
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(s2u(bar) & kPeerMask) : "memory");
}
// Programmatic dependent launch (see launch_pdl in gd_unet.cu)
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_entry() { pdl_trigger(); pdl_wait(); }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
// silu'(y) with sigmoid(y) = 0.5 + 0.5 tanh(y/2): ONE SFU op (tanh.approx, |rel err| <= 2^-11, below the fp16 rounding of
// the result) instead of ex2 + rcp
__device__ __forceinline__ float dsilu_tanh(float y) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * y));
  const float s = fmaf(0.5f, t, 0.5f);
  return s * fmaf(y, 1.0f - s, 1.0f);
}

// Persistent, warp-specialised: grid = min(tiles, SMs). The accumulator is double-buffered in
// TMEM so the epilogue of tile i overlaps the main loop of tile i+1; the TMA ring runs ahead
// across tile boundaries. 10 warps: 0 = TMA producer, 1 = MMA issuer (+TMEM owner), 2..9 =
// epilogue (two warps per TMEM lane quarter, each takes half of the tile's 32-column chunks).
// MODE: 0 plain epilogue (bias / time-embedding / residual / SiLU), 1 GEGLU, 2 transposed store,
// 3 split-K fp32 partials, 4 plain + GroupNorm-backward producer (the data-gradient GEMM in front of a GroupNorm+SiLU
// backward multiplies its output by silu'(y) and emits the two column sums that backward needs: no statistics sweep). One instantiation per mode keeps each kernel's code small (I-cache).
// TWO: CTA-pair version (launched as clusters of 2): tile = 256 x BLOCK_N, per-CTA operand traffic
// 128 x 64 of A + BLOCK_N/2 x 64 of B per k-block instead of 128 x 64 + BLOCK_N x 64.
template <int MODE, bool TWO>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_gemm_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const GemmKParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int BN = p.block_n, S = p.stages;
  const uint32_t a_bytes = kBM * kBK * 2, b_bytes = (uint32_t)(TWO ? BN / 2 : BN) * kBK * 2;   // per CTA
  const uint32_t rank = TWO ? cluster_ctarank() : 0u;
  const int tile0 = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, tile_step = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const bool bres = TWO && p.b_resident;
  if (bres && smem != smem_raw) __trap();   // resident mode is sized without the alignment slack (the base is 1 KB aligned)
  const uint32_t b_slot = (b_bytes + 1023) & ~1023u;
  const bool halo = TWO && p.a_halo;
  const bool patch = TWO && p.a_halo == 2;
  const uint32_t stage_bytes = halo ? (bres ? p.halo_bytes : p.halo_bytes + 3 * b_slot) : (bres ? a_bytes : a_bytes + b_slot);   // resident B: the ring holds A only
  uint8_t* b_res = smem + (size_t)S * stage_bytes;                      // [num_kb][b_slot] when resident
  uint8_t* stg_all = b_res + (bres ? (size_t)p.num_kb * b_slot : 0);    // epilogue staging: kEpiWarps x stg_bufs x 2 KB (1024-aligned)
  uint64_t* full = reinterpret_cast<uint64_t*>(stg_all + (size_t)kEpiWarps * p.stg_bufs * 2048);
  uint64_t* empty = full + S;
  uint64_t* tmem_full = empty + S;      // [2]
  uint64_t* tmem_empty = tmem_full + 2; // [2]
  uint64_t* bres_full = tmem_empty + 2; // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_full + 2);   // +2: keeps the bias area behind it 16-byte aligned

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t acc_cols = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;  // per accumulator
  const int n_tiles = p.n_tiles, m_tiles = p.m_tiles, total = p.total_tiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < S; s++) { bar_init(&full[s], 1); bar_init(&empty[s], 1); }
    for (int s = 0; s < 2; s++) { bar_init(&tmem_full[s], 1); bar_init(&tmem_empty[s], TWO ? 2 * kEpiWarps : kEpiWarps); }
    bar_init(bres_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if (TWO) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(tmem_slot)), "r"(2 * acc_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(tmem_slot)), "r"(2 * acc_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (TWO) cluster_sync_all(); else __syncthreads();   // barriers of BOTH CTAs initialised before any remote signal
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel; its results are visible from here on

  if (warp == 0) {
    if (elect_one()) {
      // ---------------- TMA producer ----------------
      int it = 0;  // ring position, runs across tiles
      if (bres) {   // this CTA's half of the (single) B tile, every k-block, once: it is reused by all M tiles
        if (rank == 0) bar_expect_tx(bres_full, 2u * (uint32_t)p.num_kb * b_bytes);
        for (int kb = 0; kb < p.num_kb; kb++)
          tma_load_3d_2sm(b_res + (size_t)kb * b_slot, &tmB, bres_full, kb * kBK, (int)rank * (BN / 2), 0);
      }
      for (int t = tile0; t < total; t += tile_step) {
        const int ks = t % p.ksplit, tt = t / p.ksplit;
        const int n_blk = tt % n_tiles, z = tt / (n_tiles * m_tiles);
        const int m_blk = TWO ? 2 * ((tt / n_tiles) % m_tiles) + (int)rank : (tt / n_tiles) % m_tiles;   // TWO: m_tiles counts pairs
        const int zh = z % p.heads, zb = z / p.heads;
        const int kb0 = ks * p.kb_per_split, kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        int a_c1, a_c2, a_c3;
        if (p.mode_conv) {
          // tile = imgs_box images x rows_box rows x box_w pixels (box_w < img_w: part of one row)
          const int tiles_per_img = p.rows_per_image / kBM;
          if (patch) {
            const int t_in = m_blk % p.patch_tiles_per_img;
            a_c3 = m_blk / p.patch_tiles_per_img;
            a_c2 = (t_in / p.patch_w_tiles) * 16;
            a_c1 = (t_in % p.patch_w_tiles) * 8;
          } else if (p.imgs_box == 1) {
            const int t_in = m_blk % tiles_per_img;
            a_c3 = m_blk / tiles_per_img;
            a_c2 = (t_in / p.tiles_per_row) * p.rows_box;
            a_c1 = (t_in % p.tiles_per_row) * p.box_w;
          } else { a_c3 = m_blk * p.imgs_box; a_c2 = 0; a_c1 = 0; }
        } else {
          a_c1 = m_blk * kBM; a_c2 = p.a_zflat ? z : zb; a_c3 = 0;
        }
        const int b_c1 = n_blk * BN + (TWO ? (int)rank * (BN / 2) : 0) + zh * p.b_head_n;
        const int b_c2 = p.b_zdim > 1 ? zb : 0;
        if (halo) {
          // 3 x kb_per_tap super-blocks per tile: one 130-pixel row of A (x0-1 .. x0+128, image row y+dy; TMA zero-fills
          // the pixels outside the image = the padding) and the three weight tiles of the taps (dy, dx = -1, 0, 1)
          for (int dyi = 0; dyi < 3; dyi++)
            for (int cb = 0; cb < p.kb_per_tap; cb++, it++) {
              const int s = it % S;
              bar_wait(&empty[s], ((it / S) & 1) ^ 1);
              uint8_t* sa = smem + (size_t)s * stage_bytes;
              uint8_t* sb = sa + p.halo_bytes;
              if (rank == 0) bar_expect_tx(&full[s], 2u * ((patch ? kPatchBytes : 130u * 128u) + (bres ? 0u : 3u * b_bytes)));
              tma_load_4d_2sm(sa, &tmA, &full[s], cb * kBK, a_c1 - 1, a_c2 + dyi - 1, a_c3);
              if (!bres) {
#pragma unroll
                for (int dxi = 0; dxi < 3; dxi++)
                  tma_load_3d_2sm(sb + (size_t)dxi * b_slot, &tmB, &full[s], ((dyi * 3 + dxi) * p.kb_per_tap + cb) * kBK, b_c1, b_c2);
              }
            }
          continue;
        }
        for (int kb = kb0; kb < kb1; kb++, it++) {
          const int s = it % S;
          bar_wait(&empty[s], ((it / S) & 1) ^ 1);
          uint8_t* sa = smem + (size_t)s * stage_bytes;
          uint8_t* sb = sa + a_bytes;
          if (TWO) {
            // both CTAs' copies complete on the LEADER's barrier, which expects the bytes of the pair
            if (rank == 0) bar_expect_tx(&full[s], bres ? 2 * a_bytes : 2 * (a_bytes + b_bytes));
            if (p.mode_conv) {
              const int tap = kb / p.kb_per_tap, cb = kb % p.kb_per_tap;
              tma_load_4d_2sm(sa, &tmA, &full[s], p.tap_c[tap] + cb * kBK, a_c1 + p.tap_dx[tap], a_c2 * p.a_yscale + p.tap_dy[tap], a_c3);
            } else {
              tma_load_4d_2sm(sa, &tmA, &full[s], zh * p.a_head_k + kb * kBK, a_c1, a_c2, a_c3);
            }
            if (!bres) tma_load_3d_2sm(sb, &tmB, &full[s], zh * p.b_head_k + kb * kBK, b_c1, b_c2);
          } else {
            bar_expect_tx(&full[s], a_bytes + b_bytes);
            if (p.mode_conv) {
              const int tap = kb / p.kb_per_tap, cb = kb % p.kb_per_tap;
              tma_load_4d(sa, &tmA, &full[s], p.tap_c[tap] + cb * kBK, a_c1 + p.tap_dx[tap], a_c2 * p.a_yscale + p.tap_dy[tap], a_c3);
            } else {
              tma_load_4d(sa, &tmA, &full[s], zh * p.a_head_k + kb * kBK, a_c1, a_c2, a_c3);
            }
            tma_load_3d(sb, &tmB, &full[s], zh * p.b_head_k + kb * kBK, b_c1, b_c2);
          }
        }
      }
    }
  } else if (warp == 1 && (!TWO || rank == 0)) {
    // ---------------- MMA issuer (TWO: the leader CTA only) ----------------
    const uint32_t idesc = TWO ? (umma_idesc_f16(BN) + ((uint32_t)(kBM >> 4) << 24)) : umma_idesc_f16(BN);   // TWO: M = 256
    int it = 0, lt = 0;
    if (bres) bar_wait(bres_full, 0);
    for (int t = tile0; t < total; t += tile_step, lt++) {
      const int acc = lt & 1;
      bar_wait(&tmem_empty[acc], ((lt >> 1) & 1) ^ 1);  // epilogue drained this accumulator
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tmem_d = tmem_base + (uint32_t)acc * acc_cols;
      const int ks = t % p.ksplit;
      const int kb0 = ks * p.kb_per_split, kb1 = min(p.num_kb, kb0 + p.kb_per_split);
      if (halo) {
        const int nsb = 3 * p.kb_per_tap;
        for (int sbk = 0; sbk < nsb; sbk++, it++) {
          const int s = it % S;
          bar_wait(&full[s], (it / S) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one()) {
            const uint32_t sa = s2u(smem + (size_t)s * stage_bytes);
            // weight tiles: behind the A tile in the stage, or (resident B) tap (dy, dx), channel block cb of the resident copy
            const int dyi = sbk / p.kb_per_tap, cbi = sbk % p.kb_per_tap;
            const uint32_t sb = bres ? s2u(b_res) + (uint32_t)((dyi * 3) * p.kb_per_tap + cbi) * b_slot : sa + p.halo_bytes;
            const uint32_t sb_step = bres ? (uint32_t)p.kb_per_tap * b_slot : b_slot;
#pragma unroll
            for (int dxi = 0; dxi < 3; dxi++) {
              // tap dx reads rows dx .. dx+127 of the 130-row tile: the descriptor simply starts dx rows (128 B each) later. The
              // unit derives the 128B-swizzle XOR from the absolute shared-memory address bits [7:9], exactly like the TMA that
              // wrote the tile, so no base_offset is set (measured: with base_offset = dx the result is wrong).
              const uint64_t da = patch ? umma_desc_sw128_sbo(sa + (uint32_t)dxi * 128u, 1280u) : umma_desc_sw128(sa + (uint32_t)dxi * 128u);
              const uint64_t db = umma_desc_sw128(sb + (uint32_t)dxi * sb_step);
#pragma unroll
              for (int k = 0; k < kBK / 16; k++)
                umma_f16_2sm(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (sbk | dxi | k) ? 1u : 0u);
            }
            umma_commit_2sm(&empty[s]);
            if (sbk == nsb - 1) umma_commit_2sm(&tmem_full[acc]);
          }
          __syncwarp();
        }
        continue;
      }
      for (int kb = kb0; kb < kb1; kb++, it++) {
        const int s = it % S;
        bar_wait(&full[s], (it / S) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t sa = s2u(smem + (size_t)s * stage_bytes);
          const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(bres ? s2u(b_res + (size_t)kb * b_slot) : sa + a_bytes);
          if (TWO) {
#pragma unroll
            for (int k = 0; k < kBK / 16; k++)
              umma_f16_2sm(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, ((kb - kb0) | k) ? 1u : 0u);
            umma_commit_2sm(&empty[s]);                             // frees the slot in both CTAs
            if (kb == kb1 - 1) umma_commit_2sm(&tmem_full[acc]);    // both epilogues
          } else {
#pragma unroll
            for (int k = 0; k < kBK / 16; k++)  // +32 B per K=16 step inside the 128 B swizzle row
              umma_f16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, ((kb - kb0) | k) ? 1u : 0u);
            umma_commit(&empty[s]);                                   // slot free once these MMAs read it
            if (kb == kb1 - 1) umma_commit(&tmem_full[acc]);          // accumulator complete
          }
        }
        __syncwarp();
      }
    }
  } else if (warp >= 2) {
    // ---------------- epilogue: warps 2..9; quarter q = warp & 3, column half = (warp - 2) >> 2 ----
    // Per tile a warp owns up to 4 chunks of 32 columns. Everything that does not depend on the
    // accumulator (bias + time-embedding bias -> per-warp shared memory, residual -> registers) is
    // fetched BEFORE waiting for the MMA, so the global-load latency hides behind the main loop.
    const int q = warp & 3, half = (warp - 2) >> 2;
    constexpr bool geglu = MODE == 1, transposed = MODE == 2, splitk = MODE == 3, gnb = MODE == 4;
    constexpr bool plain = MODE == 0 || MODE == 4;
    const int chunks = (BN + 31) / 32;
    float* sb = reinterpret_cast<float*>(tmem_slot + 4) + (warp - 2) * p.bias_stride;   // [chunks of this warp][32] bias sums
    // MODE 4: (ya, yb, ca, cb) of this warp's columns, behind the bias area: [4 chunks][32] float4 per warp (host reserves it)
    float4* sgn = reinterpret_cast<float4*>(reinterpret_cast<float*>(tmem_slot + 4) + kEpiWarps * p.bias_stride) + (warp - 2) * 128;
    // output staging for the TMA store: 2 x [32 rows x 64 B] per warp, 64B-swizzled, 1024-aligned
    uint8_t* stg_base = stg_all + (size_t)(warp - 2) * p.stg_bufs * 2048;
    uint32_t st_cnt = 0;
    if (plain && p.tma_store && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    const bool vec_ok = (p.ldc & 7) == 0 && plain;
    int lt = 0;
    for (int t = tile0; t < total; t += tile_step, lt++) {
      const int ks = t % p.ksplit, tt = t / p.ksplit;
      const int n_blk = tt % n_tiles, z = tt / (n_tiles * m_tiles);
      const int m_blk = TWO ? 2 * ((tt / n_tiles) % m_tiles) + (int)rank : (tt / n_tiles) % m_tiles;
      const int zh = z % p.heads, zb = z / p.heads;
      const int acc = lt & 1;
      // global output row of tile row r: linear, or (patch tiles) pixel (y0 + r/8, x0 + r%8) of the tile's image
      int prow0 = m_blk * kBM, pstep = 8;
      if (patch) {
        const int t_in = m_blk % p.patch_tiles_per_img;
        prow0 = (m_blk / p.patch_tiles_per_img) * p.rows_per_image + (t_in / p.patch_w_tiles) * 16 * p.img_w + (t_in % p.patch_w_tiles) * 8;
        pstep = p.img_w;
      }
      auto grow = [&](int r) {
        const int g = prow0 + (r >> 3) * pstep + (r & 7);
        if (p.c_up2_w > 0) { const int x2 = g % p.c_up2_w; return 4 * (g - x2) + 2 * x2; }
        return g;
      };
      auto rin = [&](int r) { return m_blk * kBM + r < p.M; };   // row r of this tile exists (also false for the phantom second tile
                                                                 // of a pair when the tile count is odd)
      const int row = grow(q * 32 + lane);
      const bool row_ok = rin(q * 32 + lane);
      const long long coff = (long long)zb * p.c_batch_stride + (long long)zh * p.c_head_stride;
      const int img = p.rows_per_image > 0 ? (m_blk * kBM + q * 32) / p.rows_per_image : 0;  // uniform per warp
      const int nlim = min(p.N, (n_blk + 1) * BN);  // columns of this tile that exist
      // ---- prefetch ----
      __syncwarp();
#pragma unroll
      for (int ci = 0; ci < 4; ci++) {
        if (ci * 32 >= p.bias_stride) break;
        const int n = n_blk * BN + (half + 2 * ci) * 32 + lane;
        float bsum = 0.f;
        if (half + 2 * ci < chunks && n < nlim && !splitk) {
          if (p.bias) bsum += __half2float(p.bias[n]);
          if (p.row_bias && m_blk * kBM < p.M) bsum += __half2float(p.row_bias[(long long)img * p.row_bias_ld + n]);
        }
        sb[ci * 32 + lane] = bsum;
        if constexpr (gnb) {   // one coalesced 512-byte load per chunk; the math loop reads them back as broadcasts
          const bool on = half + 2 * ci < chunks && n < nlim && m_blk * kBM < p.M;
          sgn[ci * 32 + lane] = on ? __ldg(p.gn_coef + (size_t)img * p.N + n) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      // residual of the warp's first chunk now, of chunk c+2 while chunk c is processed (rolled
      // loop: unrolling the chunk body 4x made the kernel 258 KB of SASS and thrashed the I-cache)
      auto fetch_res = [&](int c, uint4 (&dst)[4]) {
        const int n0 = n_blk * BN + c * 32;
        const bool on = p.residual && vec_ok && row_ok && c < chunks && n0 + 32 <= nlim;
        const uint4* res = reinterpret_cast<const uint4*>(p.residual + coff + (long long)row * p.ldc + n0);
#pragma unroll
        for (int u = 0; u < 4; u++) dst[u] = on ? res[u] : make_uint4(0, 0, 0, 0);
      };
      uint4 rnext[4];
      fetch_res(half, rnext);
      __syncwarp();
      // ---- accumulator ----
      bar_wait(&tmem_full[acc], (lt >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int c = half, ci = 0; c < chunks; c += 2, ci++) {
        uint4 rcur[4];
#pragma unroll
        for (int u = 0; u < 4; u++) rcur[u] = rnext[u];
        fetch_res(c + 2, rnext);
        const int c0 = c * 32;
        uint32_t r[32];
        const uint32_t taddr = tmem_base + (uint32_t)acc * acc_cols + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c + 2 >= chunks) {  // last TMEM read of this warp for the tile: release the accumulator
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            if (TWO) bar_arrive_leader(&tmem_empty[acc]);
            else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s2u(&tmem_empty[acc])) : "memory");
          }
        }
        const int n0 = n_blk * BN + c0;
        if (n0 >= nlim) continue;                 // warp-uniform
        if (p.flags & GD_EPI_PROBE_SKIP) continue;   // timing experiment only: epilogue math + stores skipped
        const bool full32 = n0 + 32 <= nlim;
        if (!(plain && p.tma_store && full32) && !row_ok) continue;   // the staged paths need the whole warp
        if constexpr (splitk) {  // raw fp32 partial sums; bias / residual / rounding happen in the finalize kernel
          float* wdst = p.ws + ((size_t)ks * p.M + row) * p.N + n0;
          if (full32) {
#pragma unroll
            for (int u = 0; u < 8; u++)
              reinterpret_cast<uint4*>(wdst)[u] = make_uint4(r[4 * u], r[4 * u + 1], r[4 * u + 2], r[4 * u + 3]);
          } else {
#pragma unroll 1
            for (int j = 0; j < 32; j++) if (n0 + j < nlim) wdst[j] = __uint_as_float(r[j]);
          }
          continue;
        }
        float v[32];
#pragma unroll
        for (int j4 = 0; j4 < 8; j4++) {
          const float4 bq = *reinterpret_cast<const float4*>(sb + ci * 32 + j4 * 4);  // broadcast LDS
          v[4 * j4 + 0] = __uint_as_float(r[4 * j4 + 0]) * p.alpha + bq.x;
          v[4 * j4 + 1] = __uint_as_float(r[4 * j4 + 1]) * p.alpha + bq.y;
          v[4 * j4 + 2] = __uint_as_float(r[4 * j4 + 2]) * p.alpha + bq.z;
          v[4 * j4 + 3] = __uint_as_float(r[4 * j4 + 3]) * p.alpha + bq.w;
        }
        if constexpr (geglu) {  // columns come as 16 values followed by their 16 gates
          __half* dst = p.C + coff + (long long)row * p.ldc + (n0 >> 1);
          __align__(16) __half o[16];
#pragma unroll
          for (int j = 0; j < 16; j++) o[j] = __float2half_rn(v[j] * gelu_erf(v[16 + j]));
          if (full32) {
            reinterpret_cast<uint4*>(dst)[0] = reinterpret_cast<const uint4*>(o)[0];
            reinterpret_cast<uint4*>(dst)[1] = reinterpret_cast<const uint4*>(o)[1];
          }
        } else if constexpr (transposed) {
#pragma unroll
          for (int j = 0; j < 32; j++) {
            const int n = n0 + j;
            if (n < nlim) p.C[coff + (long long)n * p.ldc + row] = __float2half_rn(v[j]);
          }
        } else {
          __half* dst = p.C + coff + (long long)row * p.ldc + n0;
          if (full32 && (p.ldc & 7) == 0) {
            uint4 o[4];
            uint4 ox[4];   // gnb: xh = (x - mean) * rstd of the same elements, fp16 (second staging tile)
            if constexpr (gnb) {
              const __half* rh = reinterpret_cast<const __half*>(&rcur[0]);
              const float4* cf = sgn + ci * 32;   // broadcast LDS.128 per column
#pragma unroll
              for (int u = 0; u < 4; u++) {
                float xh[8];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                  const int j = 8 * u + k;
                  const float4 c4 = cf[j];
                  const float xf = __half2float(rh[j]);
                  v[j] *= dsilu_tanh(fmaf(xf, c4.x, c4.y));
                  xh[k] = fmaf(xf, c4.z, c4.w);
                }
                __half2 h0 = __floats2half2_rn(xh[0], xh[1]), h1 = __floats2half2_rn(xh[2], xh[3]);
                __half2 h2 = __floats2half2_rn(xh[4], xh[5]), h3 = __floats2half2_rn(xh[6], xh[7]);
                ox[u] = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                   *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
              }
            } else {
              if (p.residual) {
                const __half* rh = reinterpret_cast<const __half*>(&rcur[0]);
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] += __half2float(rh[j]);
              }
              if (p.flags & GD_EPI_SILU) {
#pragma unroll 4
                for (int j = 0; j < 32; j++) v[j] = silu(v[j]);
              }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
              __half2 h0 = __floats2half2_rn(v[8 * u], v[8 * u + 1]), h1 = __floats2half2_rn(v[8 * u + 2], v[8 * u + 3]);
              __half2 h2 = __floats2half2_rn(v[8 * u + 4], v[8 * u + 5]), h3 = __floats2half2_rn(v[8 * u + 6], v[8 * u + 7]);
              o[u] = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
            }
            if (p.tma_store == 2) {
              // registers -> swizzled staging tile -> coalesced 16-byte stores: 4 lanes per 64-byte row
              // segment, 8 rows per instruction (every 32-byte sector written whole). Synchronous, so
              // the epilogue never queues behind the operand loads in the TMA unit.
              uint8_t* stg = stg_base;
              __syncwarp();
#pragma unroll
              for (int u = 0; u < 4; u++)
                *reinterpret_cast<uint4*>(stg + lane * 64 + ((u ^ ((lane >> 1) & 3)) << 4)) = o[u];
              if constexpr (gnb) {
#pragma unroll
                for (int u = 0; u < 4; u++)
                  *reinterpret_cast<uint4*>(stg + 2048 + lane * 64 + ((u ^ ((lane >> 1) & 3)) << 4)) = ox[u];
              }
              __syncwarp();
              const int piece = lane & 3;
              const int row0 = m_blk * kBM + q * 32;
              float cs[8], cq[8];   // column sums / sums of squares over this lane's 4 rows (GroupNorm statistics)
#pragma unroll
              for (int u = 0; u < 8; u++) { cs[u] = 0.f; cq[u] = 0.f; }
#pragma unroll
              for (int it = 0; it < 4; it++) {
                const int r = it * 8 + (lane >> 2);
                const uint4 val = *reinterpret_cast<const uint4*>(stg + r * 64 + ((piece ^ ((r >> 1) & 3)) << 4));
                const int orow = grow(q * 32 + r);
                if (rin(q * 32 + r)) {
                  *reinterpret_cast<uint4*>(p.C + coff + (long long)orow * p.ldc + n0 + piece * 8) = val;
                  if constexpr (gnb) {   // sum g | sum g * xh (the GroupNorm backward's two reductions)
                    const uint4 valx = *reinterpret_cast<const uint4*>(stg + 2048 + r * 64 + ((piece ^ ((r >> 1) & 3)) << 4));
                    const __half2* hv = reinterpret_cast<const __half2*>(&val);
                    const __half2* hx = reinterpret_cast<const __half2*>(&valx);
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                      const float2 f = __half22float2(hv[u]), x2 = __half22float2(hx[u]);
                      cs[2 * u] += f.x; cq[2 * u] = fmaf(f.x, x2.x, cq[2 * u]); cs[2 * u + 1] += f.y; cq[2 * u + 1] = fmaf(f.y, x2.y, cq[2 * u + 1]);
                    }
                  } else if (p.colstats) {
                    const __half2* hv = reinterpret_cast<const __half2*>(&val);
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                      const float2 f = __half22float2(hv[u]);
                      cs[2 * u] += f.x; cq[2 * u] += f.x * f.x; cs[2 * u + 1] += f.y; cq[2 * u + 1] += f.y * f.y;
                    }
                  }
                }
              }
              if (p.colstats) {
                // reduce-scatter over the 8 lanes that hold the same 8 columns (lane bits 2..4): 8 + 4 + 2 shuffles; afterwards a
                // lane owns quantity b4 (sum | sum of squares) of columns piece*8 + b3*4 + b2*2 + {0, 1}. Fixed order: deterministic.
                const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
                float a8[8], a4[4], a2[2];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                  const float send = b4 ? cs[u] : cq[u], keep = b4 ? cq[u] : cs[u];
                  a8[u] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                  const float send = b3 ? a8[u] : a8[u + 4], keep = b3 ? a8[u + 4] : a8[u];
                  a4[u] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
#pragma unroll
                for (int u = 0; u < 2; u++) {
                  const float send = b2 ? a4[u] : a4[u + 2], keep = b2 ? a4[u + 2] : a4[u];
                  a2[u] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                }
                const int col = n0 + piece * 8 + (b3 ? 4 : 0) + (b2 ? 2 : 0);
                float* dstats = p.colstats + ((size_t)(row0 >> 5) * 2 + (b4 ? 1 : 0)) * p.N + col;
                if (row0 < p.M) *reinterpret_cast<float2*>(dstats) = make_float2(a2[0], a2[1]);
              }
            } else if (p.tma_store == 1) {
              // registers -> swizzled staging tile -> one bulk tensor store per chunk (coalesced, async;
              // rows beyond M are clipped by the tensor map)
              uint8_t* stg = stg_base + (p.stg_bufs == 2 ? (st_cnt & 1u) * 2048 : 0u);
              const bool leader = elect_one();   // bulk groups are per thread: the same elected lane issues and waits
              if (leader) {   // the buffer's previous store has finished reading it
                if (p.stg_bufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              }
              __syncwarp();
#pragma unroll
              for (int u = 0; u < 4; u++)
                *reinterpret_cast<uint4*>(stg + lane * 64 + ((u ^ ((lane >> 1) & 3)) << 4)) = o[u];
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              __syncwarp();
              if (leader) {
                tma_store_4d(&tmC, stg, n0, m_blk * kBM + q * 32, zh, zb);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              }
              st_cnt++;
            } else if (row_ok) {
#pragma unroll
              for (int u = 0; u < 4; u++) reinterpret_cast<uint4*>(dst)[u] = o[u];
            }
          } else {
            const __half* res = p.residual ? p.residual + coff + (long long)row * p.ldc + n0 : nullptr;
#pragma unroll   // static register indices (a rolled loop would spill v[] to local memory)
            for (int j = 0; j < 32; j++) {
              if (n0 + j < nlim) {
                float x = v[j] + (res ? __half2float(res[j]) : 0.0f);
                dst[j] = __float2half_rn((p.flags & GD_EPI_SILU) ? silu(x) : x);
              }
            }
          }
        }
      }
      if (half >= chunks) {  // this warp had no chunk in the tile (BN <= 32): still release it
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (TWO) bar_arrive_leader(&tmem_empty[acc]);
          else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s2u(&tmem_empty[acc])) : "memory");
        }
      }
    }
  }
  if ((MODE == 0 || MODE == 4) && warp >= 2) {
    if (elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // staging reads + writes done
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (TWO) cluster_sync_all(); else __syncthreads();   // TWO: the peer may still signal this CTA's barriers / read its operands
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (TWO) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * acc_cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * acc_cols) : "memory");
  }
}

}  // namespace gdu
