// Backward pass kernels: TMA-staged back-to-front compositor replay with warp-shuffle gradient
// reduction (one 48-byte gradient row per (Gaussian, tile) instance, no atomics), and a fused
// per-Gaussian epilogue (instance gather-reduce + conic/cov2D/cov3D/SH/projection backward).
//
// Follows (behaviour, not code) DGR/cuda_rasterizer/backward.cu:20-601 of the reference.
#include "gd_raster_common.cuh"

namespace gd {

__device__ const float BSH_C0 = 0.28209479177387814f;
__device__ const float BSH_C1 = 0.4886025119029199f;
__device__ const float BSH_C2[5] = {1.0925484305920792f, -1.0925484305920792f,
                                    0.31539156525252005f, -1.0925484305920792f,
                                    0.5462742152960396f};
__device__ const float BSH_C3[7] = {-0.5900435899266435f, 2.890611442640554f,
                                    -0.4570457994644658f, 0.3731763325901154f,
                                    -0.4570457994644658f, 1.445305721320277f,
                                    -0.5900435899266435f};

constexpr int kBwdChunk = 32;  // records per TMA copy and per cross-warp combine
#ifndef GD_BWD_STAGES
#define GD_BWD_STAGES 3
#endif
#ifndef GD_BWD_CTAS
#define GD_BWD_CTAS 4
#endif
constexpr int kBwdStages = GD_BWD_STAGES;
constexpr int kNVal = 10;      // reduced values per (pixel, Gaussian) pair

// Warp reduce-scatter of 10 values in 12 shuffles (5+3+2+1+1) instead of 50: after it, lane
// `l` holds the full warp sum of value slot_of_lane(l) (or padding). Fixed order => deterministic.
__device__ __forceinline__ int slot_of_lane(int lane) {
  if (lane & 1) return -1;  // the xor-1 partner holds the same value
  const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1;
  const int bi = 2 * b2 + b1;  // index within the 3-wide stage (valid < 3)
  if (bi >= 3) return -1;
  const int ai = 3 * b3 + bi;  // index within the 5-wide stage (valid < 5)
  if (ai >= 5) return -1;
  return 5 * b4 + ai;
}
__device__ __forceinline__ float reduce_scatter10(const float (&v)[kNVal], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
  float a[5];
#pragma unroll
  for (int k = 0; k < 5; k++) {
    const float send = b4 ? v[k] : v[k + 5];
    const float keep = b4 ? v[k + 5] : v[k];
    a[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  float b[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float hi = (k + 3 < 5) ? a[(k + 3 < 5) ? k + 3 : 0] : 0.0f;
    const float send = b3 ? a[k] : hi;
    const float keep = b3 ? hi : a[k];
    b[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  float c[2];
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const float hi = (k + 2 < 3) ? b[(k + 2 < 3) ? k + 2 : 0] : 0.0f;
    const float send = b2 ? b[k] : hi;
    const float keep = b2 ? hi : b[k];
    c[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float send = b1 ? c[0] : c[1];
  const float keep = b1 ? c[1] : c[0];
  float d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  return d;
}

// Kernel B1: persistent CTAs (one thread per pixel of a tile) pull work items (tile, list segment)
// from the queue the forward compositor filled. An item replays records [s*kSeg, min(Mx,(s+1)*kSeg))
// of the tile's sorted list back to front, staged by bulk TMA. The per-pixel compositor state at
// the upper end of the segment comes from the forward checkpoint at that boundary:
//   T (transmittance before the boundary record) and, per channel, the normalised colour behind it
//   A = (C_final - C_prefix) / T   (alpha channel: 1 - prod_{j >= boundary}(1 - alpha_j)),
// which is exactly what the reference's back-to-front recurrence (backward.cu:518-556) holds when
// it reaches that record; the top segment starts from (T_final, 0) like the reference. Segments of
// one tile are independent work items, so a 3000-record tile no longer serialises one CTA.
//
// Inside an item the 8 warps are DECOUPLED: there is no block barrier per chunk. A warp waits for
// the chunk's records (mbarrier of the TMA copy), walks the records whose cull mask names its 8x4
// block, leaves its 10 reduced values per record in its own slice of the stage, and arrives on the
// stage's counter. The warp that arrives last sums the slices in warp order (deterministic), writes
// one coalesced 48-byte row per instance and refills the stage with the chunk kBwdStages ahead.
// Warps whose block sees few records run ahead instead of idling at a barrier.
__global__ void __launch_bounds__(kTilePix, GD_BWD_CTAS)
k_render_bwd(int W, int H, int gx, int T, const uint32_t* __restrict__ ranges,
             const uint32_t* __restrict__ seg_base, const uint2* __restrict__ items,
             GdCounters* __restrict__ counters, const float* __restrict__ sorted_rec,
             const float* __restrict__ bg, const float* __restrict__ alphas,
             const uint32_t* __restrict__ n_contrib, const float4* __restrict__ fin,
             const float* __restrict__ fin_T, const float* __restrict__ ckpt,
             const float* __restrict__ dL_dpixels, const float* __restrict__ dL_dpix_depth,
             const float* __restrict__ dL_dalphas, float* __restrict__ inst_grad) {
  constexpr int NW = kTilePix / 32;
  __shared__ __align__(128) float4 s_rec[kBwdStages][kBwdChunk * 3];
  __shared__ __align__(16) float s_part[kBwdStages][NW][kBwdChunk][kGradF];   // [stage][warp][record][value]
  __shared__ uint32_t s_wrote[kBwdStages][NW];   // per warp: records of the chunk it wrote partials for
  __shared__ uint32_t s_cnt[kBwdStages];         // warps that finished the chunk in this stage
  __shared__ __align__(8) uint64_t s_bar[kBwdStages];
  __shared__ int s_max[NW];
  __shared__ uint32_t s_item;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int lx, ly;
  tile_pixel(threadIdx.x, lx, ly);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kBwdStages; s++) { mbar_init(&s_bar[s], 1); s_cnt[s] = 0u; }
    mbar_fence_init();
  }
  const size_t N = (size_t)W * H;
  const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
  const int my_slot = slot_of_lane(lane);
  const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
  const uint32_t n_items = counters->bwd_items;
  uint32_t g = 0;   // chunks this CTA has pushed through the ring so far (stage = g % stages, parity from g / stages)
  for (;;) {
    __syncthreads();   // previous item completely finished (s_item, s_max, ring, last combine)
    if (threadIdx.x == 0) s_item = atomicAdd(&counters->bwd_next, 1u);
    __syncthreads();
    const uint32_t item = s_item;
    if (item >= n_items) break;
    const uint2 it = items[item];
    const int tg = (int)it.x, seg = (int)it.y, b = tg / T, tile = tg % T;
    const uint32_t start = ranges[2 * tg], end = ranges[2 * tg + 1];
    const int n = (int)(end - start);
    const int px = (tile % gx) * kTile + lx, py = (tile / gx) * kTile + ly;
    const bool inside = px < W && py < H;
    const float pfx = (float)px, pfy = (float)py;
    const size_t pix = (size_t)py * W + px;
    const int my_last = inside ? (int)n_contrib[(size_t)b * N + pix] : 0;
    {
      const int wm = __reduce_max_sync(0xffffffffu, my_last);
      if (lane == 0) s_max[warp] = wm;
    }
    __syncthreads();
    int Mx = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) Mx = max(Mx, s_max[w]);
    const int seg_lo = seg * kSeg, seg_hi = min(Mx, seg_lo + kSeg);
    const bool top = seg_lo + kSeg >= Mx;
    // chunk c covers list positions [max(seg_lo, hi_c - chunk), hi_c), hi_c = seg_hi - c*chunk
    const int nchunks = seg_hi > seg_lo ? (seg_hi - seg_lo + kBwdChunk - 1) / kBwdChunk : 0;
    const float* src = sorted_rec + (size_t)start * kRecF;
    float* dst = inst_grad + (size_t)start * kGradF;
    auto issue = [&](int c) {
      const int hi = seg_hi - c * kBwdChunk, lo = max(seg_lo, hi - kBwdChunk);
      const uint32_t bytes = (uint32_t)(hi - lo) * kRecF * 4;
      const int stage = (int)((g + (uint32_t)c) % kBwdStages);
      mbar_expect_tx(&s_bar[stage], bytes);
      tma_load_1d(s_rec[stage], src + (size_t)lo * kRecF, bytes, &s_bar[stage]);
    };
    if (threadIdx.x == 0)
      for (int c = 0; c < kBwdStages && c < nchunks; c++) issue(c);
    if (seg == 0) {   // instances no pixel reached get zero rows (written by the tile's first segment)
      float4* z = reinterpret_cast<float4*>(dst);
      for (int k = Mx * 3 + (int)threadIdx.x; k < n * 3; k += kTilePix) z[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float Tfin = 0.f, dLp0 = 0.f, dLp1 = 0.f, dLp2 = 0.f, dLd = 0.f, dLa = 0.f;
    float Tr = 0.f, A0 = 0.f, A1 = 0.f, A2 = 0.f, Ad = 0.f, Aa = 0.f;
    if (inside && my_last > seg_lo) {
      Tfin = 1.0f - alphas[(size_t)b * N + pix];
      const float* dp = dL_dpixels + (size_t)b * 3 * N;
      dLp0 = dp[pix]; dLp1 = dp[N + pix]; dLp2 = dp[2 * N + pix];
      dLd = dL_dpix_depth[(size_t)b * N + pix];
      dLa = dL_dalphas[(size_t)b * N + pix];
      Tr = Tfin;
      if (!top) {
        const float* ck = ckpt + (size_t)(seg_base[tg] + (uint32_t)seg) * (kCkptF * kTilePix) + threadIdx.x;
        const float4 f = fin[(size_t)b * N + pix];
        const float Tb = ck[0], iT = 1.0f / Tb;
        // the reference reaches this record with T = (1 - alpha_out) / prod_{j >= boundary}(1 - alpha_j): it
        // starts from the ROUNDED 1 - sum(alpha T) (backward.cu:463), whose relative error at opaque pixels
        // (T_final ~ 1e-4) is ~1e-4 -- keep that factor so the gradients agree with it to float rounding
        const float ratio = fin_T[(size_t)b * N + pix] * iT;   // prod_{j >= boundary}(1 - alpha_j)
        Tr = Tfin / ratio;
        A0 = (f.x - ck[kTilePix]) * iT; A1 = (f.y - ck[2 * kTilePix]) * iT; A2 = (f.z - ck[3 * kTilePix]) * iT;
        Ad = (f.w - ck[4 * kTilePix]) * iT;
        Aa = 1.0f - ratio;
      }
    }
    const float bg_dot = bg0 * dLp0 + bg1 * dLp1 + bg2 * dLp2;

    for (int c = 0; c < nchunks; c++) {
      const int stage = (int)((g + (uint32_t)c) % kBwdStages);
      const int hi = seg_hi - c * kBwdChunk, lo = max(seg_lo, hi - kBwdChunk), cnt = hi - lo;
      mbar_wait(&s_bar[stage], (uint32_t)(((g + (uint32_t)c) / kBwdStages) & 1u));
      const float4* r = s_rec[stage];
      uint32_t bits = __ballot_sync(0xffffffffu, lane < cnt && ((__float_as_uint(r[3 * lane + 2].w) >> warp) & 1u));
      uint32_t wrote = 0;
      while (bits) {          // back to front over the records that can touch this warp's pixels
        const int j = 31 - __clz(bits);
        bits &= ~(1u << j);
        const int pos = lo + j;
        const float4 A = r[3 * j], Bq = r[3 * j + 1];
        const float dx = __fsub_rn(Bq.x, pfx), dy = __fsub_rn(Bq.y, pfy);
        const float power = pair_power(dx, dy, A.x, A.y, A.z);
        const float G = expf(power);
        const float alpha = fminf(0.99f, __fmul_rn(A.w, G));
        const bool contrib = (pos < my_last) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
        if (__ballot_sync(0xffffffffu, contrib) == 0u) continue;
        float v[kNVal];
#pragma unroll
        for (int k = 0; k < kNVal; k++) v[k] = 0.0f;
        if (contrib) {
          const float4 Cq = r[3 * j + 2];
          const float om = 1.0f - alpha;
          const float inv = __fdividef(1.0f, om);
          Tr = Tr * inv;
          const float dchannel_dcolor = alpha * Tr;
          float dL_dopa = (Bq.w - A0) * dLp0;
          dL_dopa += (Cq.x - A1) * dLp1;
          dL_dopa += (Cq.y - A2) * dLp2;
          dL_dopa += (Bq.z - Ad) * dLd;
          dL_dopa += (1.0f - Aa) * dLa;
          dL_dopa *= Tr;
          dL_dopa -= Tfin * inv * bg_dot;
          // colour behind the NEXT (nearer) record: the reference's accum_rec recurrence
          A0 = alpha * Bq.w + om * A0; A1 = alpha * Cq.x + om * A1; A2 = alpha * Cq.y + om * A2;
          Ad = alpha * Bq.z + om * Ad; Aa = alpha + om * Aa;
          v[6] = dchannel_dcolor * dLp0;
          v[7] = dchannel_dcolor * dLp1;
          v[8] = dchannel_dcolor * dLp2;
          v[9] = dchannel_dcolor * dLd;
          const float dL_dG = A.w * dL_dopa;
          const float gdx = G * dx, gdy = G * dy;
          const float dG_ddelx = -gdx * A.x - gdy * A.y;
          const float dG_ddely = -gdy * A.z - gdx * A.y;
          v[0] = dL_dG * dG_ddelx * ddelx_dx;
          v[1] = dL_dG * dG_ddely * ddely_dy;
          v[2] = -0.5f * gdx * dx * dL_dG;
          v[3] = -0.5f * gdx * dy * dL_dG;
          v[4] = -0.5f * gdy * dy * dL_dG;
          v[5] = G * dL_dopa;
        }
        const float red = reduce_scatter10(v, lane);
        if (my_slot >= 0) s_part[stage][warp][j][my_slot] = red;
        wrote |= 1u << j;
      }
      // ---- arrive; the last warp of the chunk combines and refills the stage ----
      __syncwarp();
      uint32_t prev = 0;
      if (lane == 0) {
        s_wrote[stage][warp] = wrote;
        __threadfence_block();                       // this warp's partials before its arrival
        prev = atomicAdd(&s_cnt[stage], 1u);
      }
      prev = __shfl_sync(0xffffffffu, prev, 0);
      if (prev == NW - 1) {
        __threadfence_block();                       // the other warps' partials after their arrivals
        if (lane < cnt) {   // lane = record: sum the slices in warp order, one 48-byte row
          float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0;
#pragma unroll
          for (int w = 0; w < NW; w++) {
            if ((s_wrote[stage][w] >> lane) & 1u) {
              const float4* pr = reinterpret_cast<const float4*>(&s_part[stage][w][lane][0]);
              const float4 q0 = pr[0], q1 = pr[1], q2 = pr[2];
              a0.x += q0.x; a0.y += q0.y; a0.z += q0.z; a0.w += q0.w;
              a1.x += q1.x; a1.y += q1.y; a1.z += q1.z; a1.w += q1.w;
              a2.x += q2.x; a2.y += q2.y;
            }
          }
          float4* row = reinterpret_cast<float4*>(dst + (size_t)(lo + lane) * kGradF);
          row[0] = a0; row[1] = a1; row[2] = make_float4(a2.x, a2.y, 0.f, 0.f);
        }
        __syncwarp();
        if (lane == 0) {
          s_cnt[stage] = 0u;
          if (c + kBwdStages < nchunks) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of the stage before the async refill
            issue(c + kBwdStages);
          }
        }
      }
    }
    g += (uint32_t)nchunks;
  }
}

// ---- per-Gaussian epilogue ------------------------------------------------------------------
struct Mat3 {  // GLM convention: m[c][r]
  float m[3][3];
};
__device__ __forceinline__ Mat3 m3mul(const Mat3& a, const Mat3& b) {
  Mat3 o;
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int r = 0; r < 3; r++)
      o.m[c][r] = a.m[0][r] * b.m[c][0] + a.m[1][r] * b.m[c][1] + a.m[2][r] * b.m[c][2];
  return o;
}

__device__ void sh_backward(int deg, int M, const float* pos, const float* campos,
                            const float* sh, uint8_t clamped, const float* dL_dcolor,
                            float* dL_dmean, float* dL_dsh, bool accumulate) {
  const float dox = pos[0] - campos[0], doy = pos[1] - campos[1], doz = pos[2] - campos[2];
  const float len = sqrtf(dox * dox + doy * doy + doz * doz);
  const float x = dox / len, y = doy / len, z = doz / len;
  float dRGB[3];
#pragma unroll
  for (int k = 0; k < 3; k++) dRGB[k] = dL_dcolor[k] * (((clamped >> k) & 1) ? 0.0f : 1.0f);
  float dx[3] = {0, 0, 0}, dy[3] = {0, 0, 0}, dz[3] = {0, 0, 0};
#define SHV(n, k) sh[3 * (n) + (k)]
#define DSH(n, w)                                                      \
  for (int k = 0; k < 3; k++) {                                        \
    const float val = (w)*dRGB[k];                                     \
    if (accumulate) dL_dsh[3 * (n) + k] += val; else dL_dsh[3 * (n) + k] = val; \
  }
  DSH(0, BSH_C0);
  if (deg > 0) {
    DSH(1, -BSH_C1 * y); DSH(2, BSH_C1 * z); DSH(3, -BSH_C1 * x);
    for (int k = 0; k < 3; k++) {
      dx[k] = -BSH_C1 * SHV(3, k); dy[k] = -BSH_C1 * SHV(1, k); dz[k] = BSH_C1 * SHV(2, k);
    }
    if (deg > 1) {
      const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      DSH(4, BSH_C2[0] * xy); DSH(5, BSH_C2[1] * yz); DSH(6, BSH_C2[2] * (2.f * zz - xx - yy));
      DSH(7, BSH_C2[3] * xz); DSH(8, BSH_C2[4] * (xx - yy));
      for (int k = 0; k < 3; k++) {
        dx[k] += BSH_C2[0] * y * SHV(4, k) + BSH_C2[2] * 2.f * -x * SHV(6, k) +
                 BSH_C2[3] * z * SHV(7, k) + BSH_C2[4] * 2.f * x * SHV(8, k);
        dy[k] += BSH_C2[0] * x * SHV(4, k) + BSH_C2[1] * z * SHV(5, k) +
                 BSH_C2[2] * 2.f * -y * SHV(6, k) + BSH_C2[4] * 2.f * -y * SHV(8, k);
        dz[k] += BSH_C2[1] * y * SHV(5, k) + BSH_C2[2] * 2.f * 2.f * z * SHV(6, k) +
                 BSH_C2[3] * x * SHV(7, k);
      }
      if (deg > 2) {
        DSH(9, BSH_C3[0] * y * (3.f * xx - yy)); DSH(10, BSH_C3[1] * xy * z);
        DSH(11, BSH_C3[2] * y * (4.f * zz - xx - yy));
        DSH(12, BSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy));
        DSH(13, BSH_C3[4] * x * (4.f * zz - xx - yy)); DSH(14, BSH_C3[5] * z * (xx - yy));
        DSH(15, BSH_C3[6] * x * (xx - 3.f * yy));
        for (int k = 0; k < 3; k++) {
          dx[k] += (BSH_C3[0] * SHV(9, k) * 3.f * 2.f * xy + BSH_C3[1] * SHV(10, k) * yz +
                    BSH_C3[2] * SHV(11, k) * -2.f * xy + BSH_C3[3] * SHV(12, k) * -3.f * 2.f * xz +
                    BSH_C3[4] * SHV(13, k) * (-3.f * xx + 4.f * zz - yy) +
                    BSH_C3[5] * SHV(14, k) * 2.f * xz + BSH_C3[6] * SHV(15, k) * 3.f * (xx - yy));
          dy[k] += (BSH_C3[0] * SHV(9, k) * 3.f * (xx - yy) + BSH_C3[1] * SHV(10, k) * xz +
                    BSH_C3[2] * SHV(11, k) * (-3.f * yy + 4.f * zz - xx) +
                    BSH_C3[3] * SHV(12, k) * -3.f * 2.f * yz + BSH_C3[4] * SHV(13, k) * -2.f * xy +
                    BSH_C3[5] * SHV(14, k) * -2.f * yz + BSH_C3[6] * SHV(15, k) * -3.f * 2.f * xy);
          dz[k] += (BSH_C3[1] * SHV(10, k) * xy + BSH_C3[2] * SHV(11, k) * 4.f * 2.f * yz +
                    BSH_C3[3] * SHV(12, k) * 3.f * (2.f * zz - xx - yy) +
                    BSH_C3[4] * SHV(13, k) * 4.f * 2.f * xz + BSH_C3[5] * SHV(14, k) * (xx - yy));
        }
      }
    }
  }
#undef SHV
#undef DSH
  (void)M;
  const float ddx = dx[0] * dRGB[0] + dx[1] * dRGB[1] + dx[2] * dRGB[2];
  const float ddy = dy[0] * dRGB[0] + dy[1] * dRGB[1] + dy[2] * dRGB[2];
  const float ddz = dz[0] * dRGB[0] + dz[1] * dRGB[1] + dz[2] * dRGB[2];
  const float sum2 = dox * dox + doy * doy + doz * doz;
  const float inv = 1.0f / sqrtf(sum2 * sum2 * sum2);
  dL_dmean[0] += ((+sum2 - dox * dox) * ddx - doy * dox * ddy - doz * dox * ddz) * inv;
  dL_dmean[1] += (-dox * doy * ddx + (sum2 - doy * doy) * ddy - doz * doy * ddz) * inv;
  dL_dmean[2] += (-dox * doz * ddx - doy * doz * ddy + (sum2 - doz * doz) * ddz) * inv;
}

struct BwdOut {
  float* dL_dmeans2D; float* dL_dcolors; float* dL_dopacity; float* dL_dmeans3D; float* dL_dcov3D;
  float* dL_dsh; float* dL_dscales; float* dL_drotations; float* dL_dconic; float* dL_ddepths;
};

__device__ __forceinline__ void put(float* p, size_t idx, float v, bool acc) {
  if (!p) return;
  if (acc) p[idx] += v; else p[idx] = v;
}

// Kernel B2: one thread per Gaussian, loop over views. Gathers and sums the Gaussian's instance
// rows (in emission order: deterministic), then runs the whole per-Gaussian backward chain.
constexpr int kEpiBlk = 128;   // 6 CTAs of 128 threads per SM (<= 85 registers): 100k Gaussians fit in ONE wave of 148 SMs
__global__ void __launch_bounds__(kEpiBlk, 6)
k_bwd_epilogue(int P, int D, int M, int B, int W, int H, const float* __restrict__ means3D,
               const float* __restrict__ shs, const float* __restrict__ scales, float mod,
               const float* __restrict__ rotations, const float* __restrict__ cov3Ds,
               ViewPack vp, const int* __restrict__ radii,
               const uint32_t* __restrict__ tiles_touched,
               const uint32_t* __restrict__ point_offsets, const uint8_t* __restrict__ clamped,
               const uint32_t* __restrict__ inst_slot, const float* __restrict__ inst_grad,
               const GdCounters* __restrict__ counters, int sum_views, BwdOut out) {
  __shared__ float s_view[GD_MAX_VIEWS][36];
  for (int k = threadIdx.x; k < B * 35; k += blockDim.x) {
    const int b = k / 35, e = k % 35;
    s_view[b][e] = e < 16 ? vp.v[b].view[e] : (e < 32 ? vp.v[b].proj[e - 16] : vp.v[b].campos[e - 32]);
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < P;
  const int lane = threadIdx.x & 31;
  // instance-arena overflow in the forward: nothing was rendered (blank ranges, no instance rows),
  // so every gradient is an exact zero -- never touch inst_slot / inst_grad (uninitialised)
  const bool arena_ok = counters->overflow == 0u;
  float mx = 0, my = 0, mz = 0, c3[6] = {0, 0, 0, 0, 0, 0};
  if (live) {
    mx = means3D[3 * (size_t)i]; my = means3D[3 * (size_t)i + 1]; mz = means3D[3 * (size_t)i + 2];
#pragma unroll
    for (int k = 0; k < 6; k++) c3[k] = cov3Ds[6 * (size_t)i + k];
  }
  for (int b = 0; b < B; b++) {
    const size_t g = (size_t)b * P + i;
    const bool vis = live && arena_ok && radii[g] > 0;
    const uint32_t n = vis ? tiles_touched[g] : 0u;
    const uint32_t off = vis ? point_offsets[g] - n : 0u;
    // ---- gather-reduce the instance rows ----
    float s[kNVal];
#pragma unroll
    for (int k = 0; k < kNVal; k++) s[k] = 0.0f;
    unsigned big = __ballot_sync(0xffffffffu, n > 16u);
    if (n > 0 && n <= 16u) {
      for (uint32_t k = 0; k < n; k++) {
        const float4* row = reinterpret_cast<const float4*>(inst_grad + (size_t)inst_slot[off + k] * kGradF);
        const float4 a = row[0], bq = row[1], cq = row[2];
        s[0] += a.x; s[1] += a.y; s[2] += a.z; s[3] += a.w;
        s[4] += bq.x; s[5] += bq.y; s[6] += bq.z; s[7] += bq.w;
        s[8] += cq.x; s[9] += cq.y;
      }
    }
    while (big) {
      const int srcl = __ffs(big) - 1;
      big &= big - 1;
      const uint32_t n_s = __shfl_sync(0xffffffffu, n, srcl);
      const uint32_t off_s = __shfl_sync(0xffffffffu, off, srcl);
      float t[kNVal];
#pragma unroll
      for (int k = 0; k < kNVal; k++) t[k] = 0.0f;
      for (uint32_t k = lane; k < n_s; k += 32) {
        const float4* row = reinterpret_cast<const float4*>(inst_grad + (size_t)inst_slot[off_s + k] * kGradF);
        const float4 a = row[0], bq = row[1], cq = row[2];
        t[0] += a.x; t[1] += a.y; t[2] += a.z; t[3] += a.w;
        t[4] += bq.x; t[5] += bq.y; t[6] += bq.z; t[7] += bq.w;
        t[8] += cq.x; t[9] += cq.y;
      }
#pragma unroll
      for (int k = 0; k < kNVal; k++) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) t[k] += __shfl_xor_sync(0xffffffffu, t[k], o);
        if (lane == srcl) s[k] = t[k];
      }
    }
    if (!live) continue;
    const bool acc = sum_views && b > 0;
    const size_t o1 = sum_views ? (size_t)i : g;  // row index in the output tensors
    if (!vis) {
      if (!acc) {  // exact zeros, like the reference's zero-initialised outputs
        for (int k = 0; k < 3; k++) { put(out.dL_dmeans2D, 3 * o1 + k, 0.f, false); put(out.dL_dcolors, 3 * o1 + k, 0.f, false); put(out.dL_dmeans3D, 3 * o1 + k, 0.f, false); put(out.dL_dscales, 3 * o1 + k, 0.f, false); }
        put(out.dL_dopacity, o1, 0.f, false); put(out.dL_ddepths, o1, 0.f, false);
        for (int k = 0; k < 6; k++) put(out.dL_dcov3D, 6 * o1 + k, 0.f, false);
        for (int k = 0; k < 4; k++) { put(out.dL_drotations, 4 * o1 + k, 0.f, false); put(out.dL_dconic, 4 * o1 + k, 0.f, false); }
        if (out.dL_dsh) for (int k = 0; k < 3 * M; k++) out.dL_dsh[(size_t)3 * M * o1 + k] = 0.f;
      }
      continue;
    }
    const float* view = s_view[b];
    const float* proj = s_view[b] + 16;
    const float h_x = vp.v[b].focal_x, h_y = vp.v[b].focal_y;
    const float tanfovx = vp.v[b].tanfovx, tanfovy = vp.v[b].tanfovy;
    // raw per-Gaussian sums
    put(out.dL_dmeans2D, 3 * o1, s[0], acc); put(out.dL_dmeans2D, 3 * o1 + 1, s[1], acc);
    put(out.dL_dmeans2D, 3 * o1 + 2, 0.f, acc);
    put(out.dL_dconic, 4 * o1, s[2], acc); put(out.dL_dconic, 4 * o1 + 1, s[3], acc);
    put(out.dL_dconic, 4 * o1 + 2, 0.f, acc); put(out.dL_dconic, 4 * o1 + 3, s[4], acc);
    put(out.dL_dopacity, o1, s[5], acc);
    put(out.dL_dcolors, 3 * o1, s[6], acc); put(out.dL_dcolors, 3 * o1 + 1, s[7], acc);
    put(out.dL_dcolors, 3 * o1 + 2, s[8], acc);
    put(out.dL_ddepths, o1, s[9], acc);
    // ---- conic -> cov2D -> cov3D / mean (reference computeCov2DCUDA) ----
    float t0 = view[0] * mx + view[4] * my + view[8] * mz + view[12];
    float t1 = view[1] * mx + view[5] * my + view[9] * mz + view[13];
    const float t2 = view[2] * mx + view[6] * my + view[10] * mz + view[14];
    const float limx = 1.3f * tanfovx, limy = 1.3f * tanfovy;
    const float txtz = t0 / t2, tytz = t1 / t2;
    t0 = fminf(limx, fmaxf(-limx, txtz)) * t2;
    t1 = fminf(limy, fmaxf(-limy, tytz)) * t2;
    const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    const float J00 = h_x / t2, J02 = -(h_x * t0) / (t2 * t2);
    const float J11 = h_y / t2, J12 = -(h_y * t1) / (t2 * t2);
    // T rows (GLM T[0], T[1]); W[c] = (view[c], view[4+c], view[8+c])
    const float T00 = view[0] * J00 + view[2] * J02, T01 = view[4] * J00 + view[6] * J02,
                T02 = view[8] * J00 + view[10] * J02;
    const float T10 = view[1] * J11 + view[2] * J12, T11 = view[5] * J11 + view[6] * J12,
                T12 = view[9] * J11 + view[10] * J12;
    const float V[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
    float p0[3], p1[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      p0[k] = T00 * V[k][0] + T01 * V[k][1] + T02 * V[k][2];
      p1[k] = T10 * V[k][0] + T11 * V[k][1] + T12 * V[k][2];
    }
    const float a = T00 * p0[0] + T01 * p0[1] + T02 * p0[2] + 0.3f;
    const float bb = T00 * p1[0] + T01 * p1[1] + T02 * p1[2];
    const float c = T10 * p1[0] + T11 * p1[1] + T12 * p1[2] + 0.3f;
    const float dcx = s[2], dcy = s[3], dcz = s[4];
    const float denom = a * c - bb * bb;
    float dL_da = 0, dL_db = 0, dL_dc = 0;
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    float dcov[6] = {0, 0, 0, 0, 0, 0};
    if (denom2inv != 0) {
      dL_da = denom2inv * (-c * c * dcx + 2 * bb * c * dcy + (denom - a * c) * dcz);
      dL_dc = denom2inv * (-a * a * dcz + 2 * a * bb * dcy + (denom - a * c) * dcx);
      dL_db = denom2inv * 2 * (bb * c * dcx - (denom + 2 * bb * bb) * dcy + a * bb * dcz);
      dcov[0] = (T00 * T00 * dL_da + T00 * T10 * dL_db + T10 * T10 * dL_dc);
      dcov[3] = (T01 * T01 * dL_da + T01 * T11 * dL_db + T11 * T11 * dL_dc);
      dcov[5] = (T02 * T02 * dL_da + T02 * T12 * dL_db + T12 * T12 * dL_dc);
      dcov[1] = 2 * T00 * T01 * dL_da + (T00 * T11 + T01 * T10) * dL_db + 2 * T10 * T11 * dL_dc;
      dcov[2] = 2 * T00 * T02 * dL_da + (T00 * T12 + T02 * T10) * dL_db + 2 * T10 * T12 * dL_dc;
      dcov[4] = 2 * T02 * T01 * dL_da + (T01 * T12 + T02 * T11) * dL_db + 2 * T11 * T12 * dL_dc;
    }
#pragma unroll
    for (int k = 0; k < 6; k++) put(out.dL_dcov3D, 6 * o1 + k, dcov[k], acc);
    const float dT00 = 2 * p0[0] * dL_da + p1[0] * dL_db, dT01 = 2 * p0[1] * dL_da + p1[1] * dL_db,
                dT02 = 2 * p0[2] * dL_da + p1[2] * dL_db;
    const float dT10 = 2 * p1[0] * dL_dc + p0[0] * dL_db, dT11 = 2 * p1[1] * dL_dc + p0[1] * dL_db,
                dT12 = 2 * p1[2] * dL_dc + p0[2] * dL_db;
    const float dJ00 = view[0] * dT00 + view[4] * dT01 + view[8] * dT02;
    const float dJ02 = view[2] * dT00 + view[6] * dT01 + view[10] * dT02;
    const float dJ11 = view[1] * dT10 + view[5] * dT11 + view[9] * dT12;
    const float dJ12 = view[2] * dT10 + view[6] * dT11 + view[10] * dT12;
    const float tz = 1.f / t2, tz2 = tz * tz, tz3 = tz2 * tz;
    const float dtx = x_grad_mul * -h_x * tz2 * dJ02;
    const float dty = y_grad_mul * -h_y * tz2 * dJ12;
    const float dtz = -h_x * tz2 * dJ00 - h_y * tz2 * dJ11 + (2 * h_x * t0) * tz3 * dJ02 +
                      (2 * h_y * t1) * tz3 * dJ12;
    float dmean[3] = {view[0] * dtx + view[1] * dty + view[2] * dtz,
                      view[4] * dtx + view[5] * dty + view[6] * dtz,
                      view[8] * dtx + view[9] * dty + view[10] * dtz};
    // ---- mean2D / depth -> mean3D (reference preprocessCUDA backward) ----
    const float m_w = 1.0f / ((proj[3] * mx + proj[7] * my + proj[11] * mz + proj[15]) + 0.0000001f);
    const float mul1 = (proj[0] * mx + proj[4] * my + proj[8] * mz + proj[12]) * m_w * m_w;
    const float mul2 = (proj[1] * mx + proj[5] * my + proj[9] * mz + proj[13]) * m_w * m_w;
    dmean[0] += (proj[0] * m_w - proj[3] * mul1) * s[0] + (proj[1] * m_w - proj[3] * mul2) * s[1];
    dmean[1] += (proj[4] * m_w - proj[7] * mul1) * s[0] + (proj[5] * m_w - proj[7] * mul2) * s[1];
    dmean[2] += (proj[8] * m_w - proj[11] * mul1) * s[0] + (proj[9] * m_w - proj[11] * mul2) * s[1];
    const float mul3 = view[2] * mx + view[6] * my + view[10] * mz + view[14];
    dmean[0] += (view[2] - view[3] * mul3) * s[9];
    dmean[1] += (view[6] - view[7] * mul3) * s[9];
    dmean[2] += (view[10] - view[11] * mul3) * s[9];
    if (shs) {
      const float dcol[3] = {s[6], s[7], s[8]};
      sh_backward(D, M, means3D + 3 * (size_t)i, view + 32, shs + 3 * (size_t)M * i, clamped[g],
                  dcol, dmean, out.dL_dsh + (size_t)3 * M * o1, acc);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) put(out.dL_dmeans3D, 3 * o1 + k, dmean[k], acc);
    if (scales) {  // cov3D -> scale / rotation (reference computeCov3D backward)
      // scalar loads: a caller may pass a slice of a packed buffer that is only 4-byte aligned
      const float4 q = make_float4(rotations[4 * (size_t)i], rotations[4 * (size_t)i + 1], rotations[4 * (size_t)i + 2], rotations[4 * (size_t)i + 3]);
      const float r = q.x, x = q.y, y = q.z, z = q.w;
      Mat3 R = {{{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                 {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                 {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}}};
      const float sc[3] = {mod * scales[3 * (size_t)i], mod * scales[3 * (size_t)i + 1],
                           mod * scales[3 * (size_t)i + 2]};
      Mat3 S = {{{sc[0], 0, 0}, {0, sc[1], 0}, {0, 0, sc[2]}}};
      const Mat3 Mm = m3mul(S, R);
      Mat3 dSig = {{{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                    {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                    {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}}};
      Mat3 dM = m3mul(Mm, dSig);
      float dMt[3][3];  // dMt[c][r] = 2*dM[r][c]
#pragma unroll
      for (int cc = 0; cc < 3; cc++)
#pragma unroll
        for (int rr = 0; rr < 3; rr++) dMt[cc][rr] = 2.0f * dM.m[rr][cc];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        // Rt[k][j] = R[j][k]
        const float ds = R.m[0][k] * dMt[k][0] + R.m[1][k] * dMt[k][1] + R.m[2][k] * dMt[k][2];
        put(out.dL_dscales, 3 * o1 + k, ds, acc);
      }
#pragma unroll
      for (int k = 0; k < 3; k++)
#pragma unroll
        for (int rr = 0; rr < 3; rr++) dMt[k][rr] *= sc[k];
#define D_(a_, b_) dMt[a_][b_]
      const float dq0 = 2 * z * (D_(0, 1) - D_(1, 0)) + 2 * y * (D_(2, 0) - D_(0, 2)) + 2 * x * (D_(1, 2) - D_(2, 1));
      const float dq1 = 2 * y * (D_(1, 0) + D_(0, 1)) + 2 * z * (D_(2, 0) + D_(0, 2)) + 2 * r * (D_(1, 2) - D_(2, 1)) - 4 * x * (D_(2, 2) + D_(1, 1));
      const float dq2 = 2 * x * (D_(1, 0) + D_(0, 1)) + 2 * r * (D_(2, 0) - D_(0, 2)) + 2 * z * (D_(1, 2) + D_(2, 1)) - 4 * y * (D_(2, 2) + D_(0, 0));
      const float dq3 = 2 * r * (D_(0, 1) - D_(1, 0)) + 2 * x * (D_(2, 0) + D_(0, 2)) + 2 * y * (D_(1, 2) + D_(2, 1)) - 4 * z * (D_(1, 1) + D_(0, 0));
#undef D_
      put(out.dL_drotations, 4 * o1, dq0, acc); put(out.dL_drotations, 4 * o1 + 1, dq1, acc);
      put(out.dL_drotations, 4 * o1 + 2, dq2, acc); put(out.dL_drotations, 4 * o1 + 3, dq3, acc);
    }
  }
}

}  // namespace gd
