// Forward pass kernels: per-Gaussian preprocess (+ per-tile counting), scans, tile binning,
// per-tile sort + record gather, TMA-staged alpha compositor.
//
// Follows (behaviour, not code) DGR/cuda_rasterizer/forward.cu:155-381 and
// DGR/cuda_rasterizer/rasterizer_impl.cu:70-138,197-339 of the reference.
#include "gd_raster_common.cuh"

namespace gd {

__device__ const float SH_C0 = 0.28209479177387814f;
__device__ const float SH_C1 = 0.4886025119029199f;
__device__ const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f,
                                   0.31539156525252005f, -1.0925484305920792f,
                                   0.5462742152960396f};
__device__ const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f,
                                   -0.4570457994644658f, 0.3731763325901154f,
                                   -0.4570457994644658f, 1.445305721320277f,
                                   -0.5900435899266435f};

// World-space covariance from scale + (unnormalised) quaternion, rounding like the reference.
__device__ __forceinline__ void cov3d_from_scale_rot(float s0, float s1, float s2, float mod,
                                                     float4 q, float* c) {
  const float sx = __fmul_rn(mod, s0), sy = __fmul_rn(mod, s1), sz = __fmul_rn(mod, s2);
  const float r = q.x, x = q.y, y = q.z, z = q.w;
  const float yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
  const float rz = __fmul_rn(r, z), xz = __fmul_rn(x, z), rx = __fmul_rn(r, x);
  const float yyzz = __fadd_rn(yy, zz);
  const float R00 = __fsub_rn(1.0f, __fadd_rn(yyzz, yyzz));
  const float xy_m_rz = __fmaf_rn(x, y, -rz), xy_p_rz = __fmaf_rn(x, y, rz);
  const float ry_p_xz = __fmaf_rn(r, y, xz), xz_m_ry = __fmaf_rn(-r, y, xz);
  const float yz_m_rx = __fmaf_rn(y, z, -rx), yz_p_rx = __fmaf_rn(y, z, rx);
  const float xx_zz = __fmaf_rn(x, x, zz), xx_yy = __fmaf_rn(x, x, yy);
  const float R01 = __fadd_rn(xy_m_rz, xy_m_rz), R02 = __fadd_rn(ry_p_xz, ry_p_xz);
  const float R10 = __fadd_rn(xy_p_rz, xy_p_rz);
  const float R11 = __fsub_rn(1.0f, __fadd_rn(xx_zz, xx_zz));
  const float R12 = __fadd_rn(yz_m_rx, yz_m_rx);
  const float R20 = __fadd_rn(xz_m_ry, xz_m_ry), R21 = __fadd_rn(yz_p_rx, yz_p_rx);
  const float R22 = __fsub_rn(1.0f, __fadd_rn(xx_yy, xx_yy));
  const float M00 = __fmul_rn(sx, R00), M01 = __fmul_rn(sy, R01), M02 = __fmul_rn(sz, R02);
  const float M10 = __fmul_rn(sx, R10), M11 = __fmul_rn(sy, R11), M12 = __fmul_rn(sz, R12);
  const float M20 = __fmul_rn(sx, R20), M21 = __fmul_rn(sy, R21), M22 = __fmul_rn(sz, R22);
  c[0] = __fmaf_rn(M02, M02, __fmaf_rn(M00, M00, __fmul_rn(M01, M01)));
  c[1] = __fmaf_rn(M12, M02, __fmaf_rn(M10, M00, __fmul_rn(M11, M01)));
  c[2] = __fmaf_rn(M22, M02, __fmaf_rn(M20, M00, __fmul_rn(M21, M01)));
  c[3] = __fmaf_rn(M12, M12, __fmaf_rn(M10, M10, __fmul_rn(M11, M11)));
  c[4] = __fmaf_rn(M22, M12, __fmaf_rn(M20, M10, __fmul_rn(M21, M11)));
  c[5] = __fmaf_rn(M22, M22, __fmaf_rn(M20, M20, __fmul_rn(M21, M21)));
}

// EWA 2-D covariance (a, b, c) with the 0.3 low-pass, rounding like the reference.
__device__ __forceinline__ void cov2d(float x, float y, float z, float focal_x, float focal_y,
                                      float tanfovx, float tanfovy, const float* v,
                                      const float* c, float& a, float& b, float& cc) {
  const float tx = xform_row(v, 0, x, y, z), ty = xform_row(v, 1, x, y, z),
              tz = xform_row(v, 2, x, y, z);
  const float limx = __fmul_rn(tanfovx, 1.3f), limy = __fmul_rn(tanfovy, 1.3f);
  const float txtz = __fdiv_rn(tx, tz), tytz = __fdiv_rn(ty, tz);
  const float cx = fminf(limx, fmaxf(-limx, txtz));
  const float cy = fminf(limy, fmaxf(-limy, tytz));
  const float tz2 = __fmul_rn(tz, tz);
  const float J00 = __fdiv_rn(focal_x, tz);
  const float J02 = __fdiv_rn(__fmul_rn(focal_x, __fmul_rn(cx, -tz)), tz2);
  const float J11 = __fdiv_rn(focal_y, tz);
  const float J12 = __fdiv_rn(__fmul_rn(focal_y, __fmul_rn(cy, -tz)), tz2);
  const float t00 = __fmaf_rn(v[2], J02, __fmul_rn(v[0], J00));
  const float t01 = __fmaf_rn(v[6], J02, __fmul_rn(v[4], J00));
  const float t02 = __fmaf_rn(v[10], J02, __fmul_rn(v[8], J00));
  const float t10 = __fmaf_rn(v[2], J12, __fmul_rn(J11, v[1]));
  const float t11 = __fmaf_rn(v[6], J12, __fmul_rn(J11, v[5]));
  const float t12 = __fmaf_rn(v[10], J12, __fmul_rn(J11, v[9]));
  const float A0 = __fmaf_rn(t02, c[2], __fmaf_rn(t00, c[0], __fmul_rn(t01, c[1])));
  const float B0 = __fmaf_rn(t12, c[2], __fmaf_rn(t10, c[0], __fmul_rn(t11, c[1])));
  const float A1 = __fmaf_rn(t02, c[4], __fmaf_rn(t00, c[1], __fmul_rn(t01, c[3])));
  const float B1 = __fmaf_rn(t12, c[4], __fmaf_rn(t10, c[1], __fmul_rn(t11, c[3])));
  const float A2 = __fmaf_rn(t02, c[5], __fmaf_rn(t00, c[2], __fmul_rn(t01, c[4])));
  const float B2 = __fmaf_rn(t12, c[5], __fmaf_rn(t10, c[2], __fmul_rn(t11, c[4])));
  a = __fadd_rn(__fmaf_rn(t02, A2, __fmaf_rn(t00, A0, __fmul_rn(t01, A1))), 0.3f);
  b = __fmaf_rn(t02, B2, __fmaf_rn(t00, B0, __fmul_rn(t01, B1)));
  cc = __fadd_rn(__fmaf_rn(t12, B2, __fmaf_rn(t10, B0, __fmul_rn(t11, B1))), 0.3f);
}

// SH -> RGB (pre-clamp, +0.5). Degree 0 is the GarmentDreamer configuration.
__device__ void sh_to_rgb(int deg, const float* pos, const float* campos, const float* sh,
                          float* out) {
  float res[3];
#pragma unroll
  for (int k = 0; k < 3; k++) res[k] = __fmul_rn(SH_C0, sh[k]);
  if (deg > 0) {
    const float dx = pos[0] - campos[0], dy = pos[1] - campos[1], dz = pos[2] - campos[2];
    const float len = __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy))));
    const float x = __fdiv_rn(dx, len), y = __fdiv_rn(dy, len), z = __fdiv_rn(dz, len);
    for (int k = 0; k < 3; k++)
      res[k] = res[k] - SH_C1 * y * sh[3 + k] + SH_C1 * z * sh[6 + k] - SH_C1 * x * sh[9 + k];
    if (deg > 1) {
      const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      for (int k = 0; k < 3; k++)
        res[k] = res[k] + SH_C2[0] * xy * sh[12 + k] + SH_C2[1] * yz * sh[15 + k] +
                 SH_C2[2] * (2.0f * zz - xx - yy) * sh[18 + k] + SH_C2[3] * xz * sh[21 + k] +
                 SH_C2[4] * (xx - yy) * sh[24 + k];
      if (deg > 2) {
        for (int k = 0; k < 3; k++)
          res[k] = res[k] + SH_C3[0] * y * (3.0f * xx - yy) * sh[27 + k] +
                   SH_C3[1] * xy * z * sh[30 + k] +
                   SH_C3[2] * y * (4.0f * zz - xx - yy) * sh[33 + k] +
                   SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[36 + k] +
                   SH_C3[4] * x * (4.0f * zz - xx - yy) * sh[39 + k] +
                   SH_C3[5] * z * (xx - yy) * sh[42 + k] +
                   SH_C3[6] * x * (xx - 3.0f * yy) * sh[45 + k];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 3; k++) out[k] = __fadd_rn(res[k], 0.5f);
}

// Calls f(tile_index_in_view, rank) for every tile of the rectangle. Rectangles with more than 8
// tiles are walked by the whole warp, so one huge splat does not serialise its thread.
// Must be called by all 32 lanes of the warp.
template <typename F>
__device__ __forceinline__ void for_each_tile(int n, int x0, int y0, int rw, int gx, F f) {
  const int lane = threadIdx.x & 31;
  unsigned big = __ballot_sync(0xffffffffu, n > 8);
  if (n > 0 && n <= 8) {
    for (int k = 0; k < n; k++) f((y0 + k / rw) * gx + x0 + k % rw, k, true);
  }
  while (big) {
    const int src = __ffs(big) - 1;
    big &= big - 1;
    const int n_s = __shfl_sync(0xffffffffu, n, src);
    const int x0_s = __shfl_sync(0xffffffffu, x0, src);
    const int y0_s = __shfl_sync(0xffffffffu, y0, src);
    const int rw_s = __shfl_sync(0xffffffffu, rw, src);
    for (int k = lane; k < n_s; k += 32)
      f((y0_s + k / rw_s) * gx + x0_s + k % rw_s, k | (src << 24), false);
  }
}

// Kernel 1: one thread per Gaussian, loop over the B views. Writes radii, tiles_touched, the
// 48-byte record, clamp flags, the view-independent cov3D, per-(view,block) partial sums of
// tiles_touched, and counts instances per tile.
__global__ void __launch_bounds__(kBlk)
k_preprocess(int P, int D, int M, int B, int W, int H, int gx, int gy,
             const float* __restrict__ means3D, const float* __restrict__ scales, float mod,
             const float* __restrict__ rotations, const float* __restrict__ opacities,
             const float* __restrict__ shs, const float* __restrict__ cov3D_precomp,
             const float* __restrict__ colors_precomp, ViewPack vp, int* __restrict__ radii,
             float* __restrict__ rec, uint32_t* __restrict__ tiles_touched,
             float* __restrict__ cov3D_out, uint8_t* __restrict__ clamped,
             uint32_t* __restrict__ scan_partials, uint32_t* __restrict__ tile_count) {
  __shared__ float s_view[GD_MAX_VIEWS][36];
  __shared__ uint32_t s_warp[kBlk / 32];
  for (int k = threadIdx.x; k < B * 35; k += blockDim.x) {
    const int b = k / 35, e = k % 35;
    s_view[b][e] = e < 16 ? vp.v[b].view[e] : (e < 32 ? vp.v[b].proj[e - 16] : vp.v[b].campos[e - 32]);
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < P;
  const int T = gx * gy;
  float x = 0, y = 0, z = 0, op = 0, c3[6] = {0, 0, 0, 0, 0, 0};
  if (live) {
    x = means3D[3 * (size_t)i];
    y = means3D[3 * (size_t)i + 1];
    z = means3D[3 * (size_t)i + 2];
    op = opacities[i];
    if (cov3D_precomp) {
#pragma unroll
      for (int k = 0; k < 6; k++) c3[k] = cov3D_precomp[6 * (size_t)i + k];
    } else {
      // scalar loads: a caller may pass a slice of a packed buffer that is only 4-byte aligned
      const float4 q = make_float4(rotations[4 * (size_t)i], rotations[4 * (size_t)i + 1], rotations[4 * (size_t)i + 2], rotations[4 * (size_t)i + 3]);
      cov3d_from_scale_rot(scales[3 * (size_t)i], scales[3 * (size_t)i + 1],
                           scales[3 * (size_t)i + 2], mod, q, c3);
#pragma unroll
      for (int k = 0; k < 6; k++) cov3D_out[6 * (size_t)i + k] = c3[k];
    }
  }
  for (int b = 0; b < B; b++) {
    const float* v = s_view[b];
    const float* pm = s_view[b] + 16;
    int n = 0, rx0 = 0, ry0 = 0, rw = 1, rad = 0;
    const size_t g = (size_t)b * P + i;
    if (live) {
      const float pvz = xform_row(v, 2, x, y, z);
      if (!(pvz <= 0.2f)) {  // NaN passes, as in the reference
        const float hx = xform_row(pm, 0, x, y, z), hy = xform_row(pm, 1, x, y, z),
                    hw = xform_row(pm, 3, x, y, z);
        const float pw = __frcp_rn(__fadd_rn(hw, 0.0000001f));
        const float projx = __fmul_rn(hx, pw), projy = __fmul_rn(hy, pw);
        float a, bb, c;
        cov2d(x, y, z, vp.v[b].focal_x, vp.v[b].focal_y, vp.v[b].tanfovx, vp.v[b].tanfovy, v, c3,
              a, bb, c);
        const float det = __fmaf_rn(a, c, -__fmul_rn(bb, bb));
        if (det != 0.0f) {
          const float det_inv = __frcp_rn(det);
          const float mid = __fmul_rn(__fadd_rn(a, c), 0.5f);
          const float s = __fsqrt_rn(fmaxf(__fmaf_rn(mid, mid, -det), 0.1f));
          const float lam = fmaxf(__fadd_rn(mid, s), __fsub_rn(mid, s));
          const int my_radius = __float2int_ru(__fmul_rn(__fsqrt_rn(lam), 3.0f));
          const float px = ndc2pix(projx, W), py = ndc2pix(projy, H);
          int x0, y0, x1, y1;
          get_rect(px, py, my_radius, gx, gy, x0, y0, x1, y1);
          const int cnt = (x1 - x0) * (y1 - y0);
          if (cnt != 0) {
            float col[3];
            if (colors_precomp) {
              col[0] = colors_precomp[3 * (size_t)i];
              col[1] = colors_precomp[3 * (size_t)i + 1];
              col[2] = colors_precomp[3 * (size_t)i + 2];
            } else {
              float raw[3];
              sh_to_rgb(D, means3D + 3 * (size_t)i, v + 32, shs + 3 * (size_t)M * i, raw);
              uint8_t cl = 0;
#pragma unroll
              for (int k = 0; k < 3; k++) {
                cl |= (raw[k] < 0.0f) ? (1u << k) : 0u;
                col[k] = fmaxf(raw[k], 0.0f);
              }
              clamped[g] = cl;
            }
            n = cnt; rx0 = x0; ry0 = y0; rw = x1 - x0; rad = my_radius;
            float4* r4 = reinterpret_cast<float4*>(rec + g * kRecF);
            r4[0] = make_float4(__fmul_rn(c, det_inv), __fmul_rn(det_inv, -bb),
                                __fmul_rn(a, det_inv), op);
            r4[1] = make_float4(px, py, pvz, col[0]);
            // .z (exclusive instance offset) is filled by k_scatter; .w = packed tile rect
            r4[2] = make_float4(col[1], col[2], 0.0f,
                                __uint_as_float((uint32_t)x0 | ((uint32_t)y0 << 10) |
                                                ((uint32_t)(x1 - x0) << 20)));
          }
        }
      }
      radii[g] = rad;
      tiles_touched[g] = (uint32_t)n;
    }
    // block sum of n for the scan spine
    uint32_t wsum = __reduce_add_sync(0xffffffffu, (uint32_t)n);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = wsum;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t t = 0;
#pragma unroll
      for (int w = 0; w < kBlk / 32; w++) t += s_warp[w];
      scan_partials[(size_t)b * gridDim.x + blockIdx.x] = t;
    }
    __syncthreads();
    uint32_t* tc = tile_count + (size_t)b * T;
    for_each_tile(n, rx0, ry0, rw, gx, [&](int tile, int, bool) { atomicAdd(&tc[tile], 1u); });
  }
}

// Exclusive scan of data[0..n) in place by one 1024-thread block; returns the total.
__device__ uint32_t block_exclusive_scan_inplace(uint32_t* data, int n) {
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_total;
  const int nt = blockDim.x, t = threadIdx.x;
  const int seg = (n + nt - 1) / nt;
  const int lo = min(n, t * seg), hi = min(n, lo + seg);
  uint32_t sum = 0;
  for (int k = lo; k < hi; k++) sum += data[k];
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
    if ((t & 31) >= o) incl += up;
  }
  if ((t & 31) == 31) s_w[t >> 5] = incl;
  __syncthreads();
  if (t < 32) {
    uint32_t w = (t < (nt >> 5)) ? s_w[t] : 0u;
    uint32_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, wi, o);
      if (t >= o) wi += up;
    }
    s_w[t] = wi - w;
    if (t == 31) s_total = wi;
  }
  __syncthreads();
  uint32_t run = s_w[t >> 5] + incl - sum;
  for (int k = lo; k < hi; k++) {
    const uint32_t v = data[k];
    data[k] = run;
    run += v;
  }
  const uint32_t total = s_total;
  __syncthreads();
  return total;
}

// Kernel 2 (2 blocks): block 0 scans the per-block partial sums of tiles_touched; block 1 turns
// the per-tile counts into [start,end) ranges, the counters and zeroed scatter cursors.
__global__ void __launch_bounds__(1024)
k_spine(int nblk, uint32_t* __restrict__ scan_partials, int BT, int T, int B,
        const uint32_t* __restrict__ tile_count, uint32_t* __restrict__ tile_cursor,
        uint32_t* __restrict__ ranges, uint32_t* __restrict__ seg_base,
        GdCounters* __restrict__ counters, uint32_t cap) {
  if (blockIdx.x == 0) {
    const uint32_t total = block_exclusive_scan_inplace(scan_partials, nblk);
    if (threadIdx.x == 0) scan_partials[nblk] = total;
    return;
  }
  // tile_cursor doubles as the exclusive start while scanning
  for (int k = threadIdx.x; k < BT; k += blockDim.x) tile_cursor[k] = tile_count[k];
  __syncthreads();
  const uint32_t total = block_exclusive_scan_inplace(tile_cursor, BT);
  const bool overflow = total > cap;
  for (int k = threadIdx.x; k < BT; k += blockDim.x) {
    const uint32_t s = tile_cursor[k], c = tile_count[k];
    const bool empty = (c == 0) || overflow;
    ranges[2 * k] = empty ? 0u : s;
    ranges[2 * k + 1] = empty ? 0u : s + c;
    if (k % T == 0) counters->view_base[k / T] = s;
  }
  if (threadIdx.x == 0) {
    counters->num_rendered = total;
    counters->overflow = overflow ? 1u : 0u;
    counters->view_base[B] = total;
    counters->bwd_items = 0u;   // the forward compositor queues the backward work items
    counters->bwd_next = 0u;
    counters->depth_max_bits = 0u;
  }
  // checkpoint slots: a tile with c instances has (c-1)/kSeg interior segment boundaries
  for (int k = threadIdx.x; k < BT; k += blockDim.x) {
    const uint32_t c = tile_count[k];
    seg_base[k] = (c == 0 || overflow) ? 0u : (c - 1) / (uint32_t)kSeg;
  }
  __syncthreads();
  block_exclusive_scan_inplace(seg_base, BT);
}

// Kernel 3: finishes the scan (point_offsets), stores each visible Gaussian's exclusive instance
// offset in its record, and scatters (depth, index) keys into the tile bins.
__global__ void __launch_bounds__(kBlk)
k_scatter(int P, int gx, int T, const uint32_t* __restrict__ tiles_touched,
          const uint32_t* __restrict__ scan_partials, uint32_t* __restrict__ point_offsets,
          float* __restrict__ rec, uint32_t* __restrict__ tile_cursor,
          const GdCounters* __restrict__ counters, uint64_t* __restrict__ tile_keys) {
  __shared__ uint32_t s_warp[kBlk / 32];
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < P;
  const size_t g = (size_t)b * P + i;
  const uint32_t n = live ? tiles_touched[g] : 0u;
  uint32_t incl = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += up;
  }
  if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
  __syncthreads();
  uint32_t base = scan_partials[(size_t)b * gridDim.x + blockIdx.x];
  for (int w = 0; w < (threadIdx.x >> 5); w++) base += s_warp[w];
  incl += base;
  if (live) point_offsets[g] = incl;
  int x0 = 0, y0 = 0, rw = 1;
  uint32_t dbits = 0;
  if (n > 0) {
    float4* r4 = reinterpret_cast<float4*>(rec + g * kRecF);
    const float4 r1 = r4[1];
    float4 r2 = r4[2];
    const uint32_t pk = __float_as_uint(r2.w);
    x0 = pk & 1023; y0 = (pk >> 10) & 1023; rw = pk >> 20;
    dbits = __float_as_uint(r1.z);
    r2.z = __uint_as_float(incl - n);
    r4[2] = r2;
  }
  if (counters->overflow) return;  // uniform over the grid
  // keys of warp-cooperative walks need the owner's depth/index: broadcast through shuffles
  uint32_t* cur = tile_cursor + (size_t)b * T;
  const int lane = threadIdx.x & 31;
  unsigned big = __ballot_sync(0xffffffffu, n > 8);
  if (n > 0 && n <= 8) {
    const uint64_t key = ((uint64_t)dbits << 32) | (uint32_t)i;
    for (uint32_t k = 0; k < n; k++) {
      const int tile = (y0 + (int)k / rw) * gx + x0 + (int)k % rw;
      const uint32_t slot = atomicAdd(&cur[tile], 1u);
      tile_keys[slot] = key;
    }
  }
  while (big) {
    const int src = __ffs(big) - 1;
    big &= big - 1;
    const int n_s = (int)__shfl_sync(0xffffffffu, n, src);
    const int x0_s = __shfl_sync(0xffffffffu, x0, src);
    const int y0_s = __shfl_sync(0xffffffffu, y0, src);
    const int rw_s = __shfl_sync(0xffffffffu, rw, src);
    const uint32_t d_s = __shfl_sync(0xffffffffu, dbits, src);
    const uint32_t i_s = (uint32_t)__shfl_sync(0xffffffffu, i, src);
    const uint64_t key = ((uint64_t)d_s << 32) | i_s;
    for (int k = lane; k < n_s; k += 32) {
      const int tile = (y0_s + k / rw_s) * gx + x0_s + k % rw_s;
      const uint32_t slot = atomicAdd(&cur[tile], 1u);
      tile_keys[slot] = key;
    }
  }
}

// ---- per-tile bitonic sort (all-ascending network, works for any n without padding) --------
__device__ __forceinline__ void cmpx(uint64_t* k, int lo, int hi) {
  const uint64_t a = k[lo], b = k[hi];
  if (a > b) { k[lo] = b; k[hi] = a; }
}
// Runs merge sizes kfrom..kto (powers of two) fully inside `keys[0..n)`.
__device__ void bitonic_local(uint64_t* keys, int n, int kfrom, int kto) {
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (int k = kfrom; k <= kto && (k >> 1) < n; k <<= 1) {
    const int half = k >> 1;
    for (int c = threadIdx.x; c < (np2 >> 1); c += blockDim.x) {
      const int blk = c / half, o = c % half;
      const int lo = blk * k + o, hi = blk * k + k - 1 - o;
      if (hi < n) cmpx(keys, lo, hi);
    }
    __syncthreads();
    for (int j = half >> 1; j >= 1; j >>= 1) {
      for (int c = threadIdx.x; c < (np2 >> 1); c += blockDim.x) {
        const int lo = (c / j) * (j << 1) + (c % j), hi = lo + j;
        if (hi < n) cmpx(keys, lo, hi);
      }
      __syncthreads();
    }
  }
}
// Half-cleaner steps j = jfrom..1 only (used after global-memory steps of a large merge).
__device__ void bitonic_clean_local(uint64_t* keys, int n, int jfrom) {
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (int j = jfrom; j >= 1; j >>= 1) {
    for (int c = threadIdx.x; c < (np2 >> 1); c += blockDim.x) {
      const int lo = (c / j) * (j << 1) + (c % j), hi = lo + j;
      if (hi < n) cmpx(keys, lo, hi);
    }
    __syncthreads();
  }
}

// Kernel 4: one CTA per (view, tile). Sorts the tile's keys by (depth bits, Gaussian index),
// then writes point_list, gathers the 48-byte records into sorted order (so the compositors can
// stream them with bulk TMA copies) and records where each unsorted instance ended up.
__global__ void __launch_bounds__(kTilePix)
k_tile_sort(int P, int gx, int T, const uint32_t* __restrict__ ranges,
            uint64_t* __restrict__ tile_keys, const float* __restrict__ rec,
            uint32_t* __restrict__ point_list, float* __restrict__ sorted_rec,
            uint32_t* __restrict__ inst_slot) {
  __shared__ uint64_t s_keys[kSortCap];
  const int tg = blockIdx.x;
  const uint32_t start = ranges[2 * tg], end = ranges[2 * tg + 1];
  const int n = (int)(end - start);
  if (n == 0) return;
  const int b = tg / T, tile = tg % T;
  uint64_t* gk = tile_keys + start;
  if (n <= kSortCap) {
    for (int k = threadIdx.x; k < n; k += blockDim.x) s_keys[k] = gk[k];
    __syncthreads();
    bitonic_local(s_keys, n, 2, kSortCap);
    for (int k = threadIdx.x; k < n; k += blockDim.x) gk[k] = s_keys[k];
  } else {
    // chunks of kSortCap sorted in shared memory, larger merge steps in global memory
    for (int cb = 0; cb < n; cb += kSortCap) {
      const int cn = min(kSortCap, n - cb);
      for (int k = threadIdx.x; k < cn; k += blockDim.x) s_keys[k] = gk[cb + k];
      __syncthreads();
      bitonic_local(s_keys, cn, 2, kSortCap);
      for (int k = threadIdx.x; k < cn; k += blockDim.x) gk[cb + k] = s_keys[k];
      __syncthreads();
    }
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int k = kSortCap << 1; (k >> 1) < n; k <<= 1) {
      const int half = k >> 1;
      for (int c = threadIdx.x; c < (np2 >> 1); c += blockDim.x) {
        const int blk = c / half, o = c % half;
        const int lo = blk * k + o, hi = blk * k + k - 1 - o;
        if (hi < n) cmpx(gk, lo, hi);
      }
      __syncthreads();
      for (int j = half >> 1; j >= kSortCap; j >>= 1) {
        for (int c = threadIdx.x; c < (np2 >> 1); c += blockDim.x) {
          const int lo = (c / j) * (j << 1) + (c % j), hi = lo + j;
          if (hi < n) cmpx(gk, lo, hi);
        }
        __syncthreads();
      }
      for (int cb = 0; cb < n; cb += kSortCap) {
        const int cn = min(kSortCap, n - cb);
        for (int q = threadIdx.x; q < cn; q += blockDim.x) s_keys[q] = gk[cb + q];
        __syncthreads();
        bitonic_clean_local(s_keys, cn, kSortCap >> 1);
        for (int q = threadIdx.x; q < cn; q += blockDim.x) gk[cb + q] = s_keys[q];
        __syncthreads();
      }
    }
  }
  __syncthreads();
  const int tx = tile % gx, ty = tile / gx;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const uint32_t idx = (uint32_t)(gk[k] & 0xffffffffull);
    const size_t g = (size_t)b * P + idx;
    const float4* src = reinterpret_cast<const float4*>(rec + g * kRecF);
    const float4 r0 = src[0], r1 = src[1], r2 = src[2];
    float4* dst = reinterpret_cast<float4*>(sorted_rec + (size_t)(start + k) * kRecF);
    // sorted record: conic.xyz, opacity | px, py, depth, r | g, b, 0, per-warp cull mask (8 bits)
    const uint32_t cull = warp_cull_mask(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, (float)(tx * kTile), (float)(ty * kTile));
    dst[0] = r0; dst[1] = r1; dst[2] = make_float4(r2.x, r2.y, 0.0f, __uint_as_float(cull));
    point_list[start + k] = idx;
    const uint32_t pk = __float_as_uint(r2.w);
    const int x0 = pk & 1023, y0 = (pk >> 10) & 1023, rw = pk >> 20;
    const uint32_t u = __float_as_uint(r2.z) + (uint32_t)((ty - y0) * rw + (tx - x0));
    inst_slot[u] = start + k;
  }
}

// Kernel 5: alpha compositor. One CTA per (view, tile), one thread per pixel. The tile's sorted
// records are contiguous in memory and are streamed into a 4-slot shared-memory ring with bulk
// TMA copies (cp.async.bulk + mbarrier); every thread walks the ring front to back. Each record
// carries the 8-bit mask of the warps (8x4 pixel blocks) it can reach (k_tile_sort), so a warp only
// visits its own records.
// For the backward pass the kernel also leaves (a) the per-pixel accumulators before the background
// term (`fin`), (b) a checkpoint (T, C, D per pixel) every kSeg records and (c) one work item per
// list segment [s*kSeg, (s+1)*kSeg) that some pixel of the tile blended: k_render_bwd replays the
// segments of a tile independently, which turns 1 long tile into several short work items.
constexpr int kFwdChunk = 64;   // records per TMA copy (3 KB)
constexpr int kFwdStages = 4;
static_assert(kSeg % kFwdChunk == 0, "checkpoints are taken at chunk boundaries");

__global__ void __launch_bounds__(kTilePix)
k_render_fwd(int W, int H, int gx, int T, const uint32_t* __restrict__ ranges,
             const float* __restrict__ sorted_rec, const float* __restrict__ bg,
             float* __restrict__ out_color, float* __restrict__ out_depth,
             float* __restrict__ out_alpha, uint32_t* __restrict__ n_contrib,
             float4* __restrict__ fin, float* __restrict__ fin_T, const uint32_t* __restrict__ seg_base,
             float* __restrict__ ckpt,
             uint2* __restrict__ items, GdCounters* __restrict__ counters) {
  __shared__ __align__(128) float4 s_rec[kFwdStages][kFwdChunk * 3];
  __shared__ __align__(8) uint64_t s_bar[kFwdStages];
  __shared__ uint32_t s_max[kTilePix / 32];
  const int tg = blockIdx.x, b = tg / T, tile = tg % T;
  const uint32_t start = ranges[2 * tg], end = ranges[2 * tg + 1];
  const int n = (int)(end - start);
  const int nchunks = (n + kFwdChunk - 1) / kFwdChunk;
  int lx, ly;
  tile_pixel(threadIdx.x, lx, ly);
  const int px = (tile % gx) * kTile + lx, py = (tile / gx) * kTile + ly;
  const bool inside = px < W && py < H;
  const float pfx = (float)px, pfy = (float)py;
  const float* src = sorted_rec + (size_t)start * kRecF;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kFwdStages; s++) mbar_init(&s_bar[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int c = 0; c < kFwdStages && c < nchunks; c++) {
      const uint32_t bytes = (uint32_t)min(kFwdChunk, n - c * kFwdChunk) * kRecF * 4;
      mbar_expect_tx(&s_bar[c], bytes);
      tma_load_1d(s_rec[c], src + (size_t)c * kFwdChunk * kRecF, bytes, &s_bar[c]);
    }
  }
  bool done = !inside;
  float Tr = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, weight = 0.f, Dp = 0.f;
  uint32_t last_contributor = 0;
  int issued = min(kFwdStages, nchunks);  // tracked identically by every thread
  int c = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* ck = ckpt + (size_t)(n > 0 ? seg_base[tg] : 0u) * (kCkptF * kTilePix) + threadIdx.x;
  for (; c < nchunks; c++) {
    const int slot = c % kFwdStages;
    if (c > 0 && (c * kFwdChunk) % kSeg == 0) {   // state after the first c*kFwdChunk records
      float* o = ck + (size_t)((c * kFwdChunk) / kSeg - 1) * (kCkptF * kTilePix);
      o[0] = Tr; o[kTilePix] = C0; o[2 * kTilePix] = C1; o[3 * kTilePix] = C2; o[4 * kTilePix] = Dp;
    }
    mbar_wait(&s_bar[slot], (uint32_t)((c / kFwdStages) & 1));
    const int cnt = min(kFwdChunk, n - c * kFwdChunk);
    {
      // Branch-free body + __syncwarp keeps the 32 pixels of a warp converged: with `continue`
      // in the loop the lanes drift apart and the warp issues each lane group separately.
      // Records whose footprint cannot reach this warp's 8x4 block are skipped warp-uniformly.
      const float4* r = s_rec[slot];
      for (int h = 0; h < cnt; h += 32) {
        const bool mine = h + lane < cnt && ((__float_as_uint(r[3 * (h + lane) + 2].w) >> warp) & 1u);
        uint32_t bits = __ballot_sync(0xffffffffu, mine);
        bool finished = false;
        while (bits) {
        const int j = h + __ffs(bits) - 1;
        bits &= bits - 1;
        if (__ballot_sync(0xffffffffu, !done) == 0u) { finished = true; break; }  // whole warp finished
        const float4 A = r[3 * j], Bq = r[3 * j + 1];
        const float dx = __fsub_rn(Bq.x, pfx), dy = __fsub_rn(Bq.y, pfy);
        const float power = pair_power(dx, dy, A.x, A.y, A.z);
        const float alpha = fminf(0.99f, __fmul_rn(A.w, expf(power)));
        const float test_T = __fmul_rn(Tr, __fsub_rn(1.0f, alpha));
        bool valid = !done && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
        if (valid && test_T < 0.0001f) { done = true; valid = false; }
        if (valid) {
          const float4 Cq = r[3 * j + 2];
          C0 = __fmaf_rn(Tr, __fmul_rn(alpha, Bq.w), C0);
          C1 = __fmaf_rn(Tr, __fmul_rn(alpha, Cq.x), C1);
          C2 = __fmaf_rn(Tr, __fmul_rn(alpha, Cq.y), C2);
          weight = __fmaf_rn(Tr, alpha, weight);
          Dp = __fmaf_rn(Tr, __fmul_rn(alpha, Bq.z), Dp);
          Tr = test_T;
          last_contributor = (uint32_t)(c * kFwdChunk + j + 1);
        }
        __syncwarp();
        }
        if (finished) break;
      }
    }
    const int ndone = __syncthreads_count(done);  // also orders slot reuse after all reads
    if (ndone == kTilePix) { c++; break; }
    if (c + kFwdStages < nchunks) {
      if (threadIdx.x == 0) {
        const int cc = c + kFwdStages;
        const uint32_t bytes = (uint32_t)min(kFwdChunk, n - cc * kFwdChunk) * kRecF * 4;
        mbar_expect_tx(&s_bar[slot], bytes);
        tma_load_1d(s_rec[slot], src + (size_t)cc * kFwdChunk * kRecF, bytes, &s_bar[slot]);
      }
      issued = c + kFwdStages + 1;
    }
  }
  // drain copies still in flight before the CTA's shared memory is released
  if (threadIdx.x == 0)
    for (; c < issued; c++) mbar_wait(&s_bar[c % kFwdStages], (uint32_t)((c / kFwdStages) & 1));
  if (inside) {
    const size_t N = (size_t)W * H, pix = (size_t)py * W + px;
    n_contrib[(size_t)b * N + pix] = last_contributor;
    fin[(size_t)b * N + pix] = make_float4(C0, C1, C2, Dp);
    fin_T[(size_t)b * N + pix] = Tr;
    float* oc = out_color + (size_t)b * 3 * N;
    oc[pix] = __fmaf_rn(Tr, bg[0], C0);
    oc[N + pix] = __fmaf_rn(Tr, bg[1], C1);
    oc[2 * N + pix] = __fmaf_rn(Tr, bg[2], C2);
    out_alpha[(size_t)b * N + pix] = weight;
    out_depth[(size_t)b * N + pix] = Dp;
  }
  if (n > 0) {   // queue the tile's backward work items: one per kSeg records that were blended (at least one)
    const uint32_t wm = __reduce_max_sync(0xffffffffu, last_contributor);
    // depth image maximum over the view batch (the depths are >= 0, so their IEEE bits order like integers)
    const uint32_t dm = __reduce_max_sync(0xffffffffu, inside ? __float_as_uint(fmaxf(Dp, 0.0f)) : 0u);
    if (lane == 0) { s_max[warp] = wm; if (dm) atomicMax(&counters->depth_max_bits, dm); }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t mx = 0;
#pragma unroll
      for (int w = 0; w < kTilePix / 32; w++) mx = max(mx, s_max[w]);
      const uint32_t nseg = mx == 0 ? 1u : (mx + kSeg - 1) / kSeg;
      const uint32_t base = atomicAdd(&counters->bwd_items, nseg);
      for (uint32_t sgi = 0; sgi < nseg; sgi++) items[base + sgi] = make_uint2((uint32_t)tg, nseg - 1 - sgi);
    }
  }
}

__global__ void k_mark_visible(int P, const float* __restrict__ means3D,
                               const float* __restrict__ view, uint8_t* __restrict__ present) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  present[i] = !(xform_row(view, 2, means3D[3 * (size_t)i], means3D[3 * (size_t)i + 1],
                           means3D[3 * (size_t)i + 2]) <= 0.2f);
}

}  // namespace gd
