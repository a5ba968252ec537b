// Shared declarations for the rasteriser kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gd_raster.h"

namespace gd {

constexpr int kTile = GD_TILE;
constexpr int kTilePix = kTile * kTile;  // 256 threads per tile CTA
constexpr int kRecF = 12;                // floats per per-(view,Gaussian) record (48 B)
constexpr int kGradF = 12;               // floats per per-instance gradient row (48 B)
constexpr int kBlk = 256;                // Gaussians per block in the per-Gaussian kernels
constexpr int kSortCap = 4096;           // keys sorted in shared memory per pass
#ifndef GD_SEG
#define GD_SEG 512
#endif
constexpr int kSeg = GD_SEG;                // sorted records per backward work item (list segment of one tile)
constexpr int kCkptF = 5;                // floats per pixel in a forward checkpoint: T, C0, C1, C2, D

// Per-view constants, passed by value as a kernel parameter (<= 32 * 40 B).
struct ViewDev {
  const float* view;
  const float* proj;
  const float* campos;
  float tanfovx, tanfovy, focal_x, focal_y;
};
struct ViewPack {
  ViewDev v[GD_MAX_VIEWS];
};

// Pointers carved out of the three caller-owned state buffers.
struct State {
  // geom
  float* rec;               // [B*P,12]
  uint32_t* tiles_touched;  // [B*P]
  uint32_t* point_offsets;  // [B*P] inclusive
  float* cov3D;             // [P,6]
  uint8_t* clamped;         // [B*P]
  uint32_t* scan_partials;  // [B*ceil(P/256)]
  GdCounters* counters;
  // img
  uint32_t* n_contrib;    // [B*N]
  uint32_t* ranges;       // [B*T,2]
  uint32_t* tile_count;   // [B*T]
  uint32_t* tile_cursor;  // [B*T]
  float4* fin;            // [B*N] forward accumulators (C0, C1, C2, D) before the background term
  float* fin_T;           // [B*N] final transmittance as the forward's running product
  uint32_t* seg_base;     // [B*T] exclusive scan of the interior segment boundaries of each tile
  // binning
  uint64_t* tile_keys;   // [cap]
  uint32_t* point_list;  // [cap]
  float* sorted_rec;     // [cap,12]
  uint32_t* inst_slot;   // [cap]
  float* inst_grad;      // [cap,12]
  float* ckpt;           // [cap/kSeg + 1][kCkptF][256] per-pixel compositor state at segment boundaries
  uint2* items;          // [cap/kSeg + B*T + 1] backward work queue: (global tile, segment)
  size_t geom_bytes, binning_bytes, img_bytes;
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <typename T>
inline void carve(char*& p, T*& out, size_t count) {
  p = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(p), 128));
  out = reinterpret_cast<T*>(p);
  p += count * sizeof(T);
}

inline State carve_state(int P, int W, int H, int B, uint32_t cap, void* geom, void* binning,
                         void* img) {
  State s;
  const size_t BP = (size_t)B * P, N = (size_t)W * H;
  const size_t T = (size_t)((W + kTile - 1) / kTile) * ((H + kTile - 1) / kTile);
  const size_t nblk = (size_t)B * ((P + kBlk - 1) / kBlk);
  char* p = reinterpret_cast<char*>(geom);
  char* p0 = p;
  carve(p, s.rec, BP * kRecF);
  carve(p, s.tiles_touched, BP);
  carve(p, s.point_offsets, BP);
  carve(p, s.cov3D, (size_t)P * 6);
  carve(p, s.clamped, BP);
  carve(p, s.scan_partials, nblk + 1);
  carve(p, s.counters, 1);
  s.geom_bytes = (size_t)(p - p0) + 128;
  p = reinterpret_cast<char*>(img);
  p0 = p;
  carve(p, s.n_contrib, (size_t)B * N);
  carve(p, s.ranges, (size_t)B * T * 2);
  carve(p, s.tile_count, (size_t)B * T);
  carve(p, s.tile_cursor, (size_t)B * T);
  carve(p, s.fin, (size_t)B * N);
  carve(p, s.fin_T, (size_t)B * N);
  carve(p, s.seg_base, (size_t)B * T);
  s.img_bytes = (size_t)(p - p0) + 128;
  p = reinterpret_cast<char*>(binning);
  p0 = p;
  carve(p, s.tile_keys, (size_t)cap);
  carve(p, s.point_list, (size_t)cap);
  carve(p, s.sorted_rec, (size_t)cap * kRecF);
  carve(p, s.inst_slot, (size_t)cap);
  carve(p, s.inst_grad, (size_t)cap * kGradF);
  carve(p, s.ckpt, ((size_t)cap / kSeg + 1) * kCkptF * kTilePix);
  carve(p, s.items, (size_t)cap / kSeg + (size_t)B * T + 1);
  s.binning_bytes = (size_t)(p - p0) + 128;
  return s;
}

// ---- mbarrier + 1-D bulk TMA (cp.async.bulk) helpers -------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"   // suspend-time hint: no hot spinning
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// Contiguous global -> shared bulk copy through the TMA engine; bytes % 16 == 0, 16-B aligned.
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- reference-exact helpers (see oracle/raster_oracle.c for the derivation) --------------
__device__ __forceinline__ float xform_row(const float* m, int r, float x, float y, float z) {
  return __fadd_rn(m[12 + r], __fmaf_rn(z, m[8 + r], __fmaf_rn(x, m[r], __fmul_rn(y, m[4 + r]))));
}
__device__ __forceinline__ float ndc2pix(float v, int S) {
  return __double2float_rn(
      __dmul_rn(__fma_rn(__dadd_rn((double)v, 1.0), (double)S, -1.0), 0.5));
}
__device__ __forceinline__ void get_rect(float px, float py, int max_radius, int gx, int gy,
                                         int& x0, int& y0, int& x1, int& y1) {
  const float r = (float)max_radius;
  x0 = min(gx, max(0, __float2int_rz(__fmul_rn(__fsub_rn(px, r), 0.0625f))));
  y0 = min(gy, max(0, __float2int_rz(__fmul_rn(__fsub_rn(py, r), 0.0625f))));
  x1 = min(gx, max(0, __float2int_rz(__fmul_rn(
                          __fsub_rn(__fadd_rn(__fadd_rn(px, r), 16.0f), 1.0f), 0.0625f))));
  y1 = min(gy, max(0, __float2int_rz(__fmul_rn(
                          __fsub_rn(__fadd_rn(__fadd_rn(py, r), 16.0f), 1.0f), 0.0625f))));
}
__device__ __forceinline__ float pair_power(float dx, float dy, float cx, float cy, float cz) {
  return __fmaf_rn(__fmaf_rn(dx, __fmul_rn(dx, cx), __fmul_rn(dy, __fmul_rn(dy, cz))), -0.5f,
                   -__fmul_rn(dy, __fmul_rn(dx, cy)));
}

// Conservative per-warp culling of a sorted record inside a tile: bit w of the result is 0 only
// if NO pixel of warp w's 8x4 block (tile_pixel) can reach alpha >= 1/255 for this Gaussian, so
// skipping the record for that warp changes nothing (the reference evaluates and rejects it per
// pixel, forward.cu:344-349 / backward.cu:503-508). With the conic Q = [[a,b],[b,c]] and
// d = centre - pixel: alpha = o exp(-0.5 q(d)), q(d) = a dx^2 + 2 b dx dy + c dy^2, so the block is
// culled iff min over the block's rectangle of q exceeds 2 ln(255 o) (0.1 % + 0.01 margin; the float
// error of the kernels' own power/exp evaluation is ~1e-6). q is convex: its minimum over a
// rectangle that does not contain the centre lies on one of the four edges, where it is a clamped
// 1-D parabola. Computed ONCE per instance by k_tile_sort and stored in the sorted record.
__device__ __forceinline__ uint32_t warp_cull_mask(float a, float b, float c, float opacity, float cx, float cy,
                                                   float tile_x0, float tile_y0) {
  const float det = a * c - b * b;
  if (!(det > 0.0f) || !(a > 0.0f) || !(c > 0.0f) || !(opacity > 0.0f) || !(cx == cx) || !(cy == cy))
    return 0xFFu;   // degenerate or NaN: never cull
  const float thr = 2.0f * logf(255.0f * opacity) + 0.01f;
  const float ia = 1.0f / a, ic = 1.0f / c;
  uint32_t m = 0;
#pragma unroll
  for (int w = 0; w < 8; w++) {
    const float bx0 = tile_x0 + (float)((w & 1) << 3), by0 = tile_y0 + (float)((w >> 1) << 2);
    // d ranges over [xl, xh] x [yl, yh]
    const float xl = cx - (bx0 + 7.0f), xh = cx - bx0, yl = cy - (by0 + 3.0f), yh = cy - by0;
    float qmin = 0.0f;
    if (!(xl <= 0.0f && xh >= 0.0f && yl <= 0.0f && yh >= 0.0f)) {
      qmin = 3.0e38f;
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const float X = e ? xh : xl;
        const float y = fminf(yh, fmaxf(yl, -b * X * ic));
        qmin = fminf(qmin, a * X * X + 2.0f * b * X * y + c * y * y);
        const float Y = e ? yh : yl;
        const float x = fminf(xh, fmaxf(xl, -b * Y * ia));
        qmin = fminf(qmin, a * x * x + 2.0f * b * x * Y + c * Y * Y);
      }
    }
    if (!(qmin * 0.999f > thr)) m |= 1u << w;
  }
  return m;
}

// pixel handled by thread t of a tile CTA: warps cover 8x4 pixel blocks (compact footprint).
__device__ __forceinline__ void tile_pixel(int t, int& lx, int& ly) {
  const int w = t >> 5, l = t & 31;
  lx = ((w & 1) << 3) + (l & 7);
  ly = ((w >> 1) << 2) + (l >> 3);
}

}  // namespace gd
