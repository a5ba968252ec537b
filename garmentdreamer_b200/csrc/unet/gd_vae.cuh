// Kernels around the tcgen05 GEMM for the VAE encoder forward + input-gradient backward
// (encode_images, Garment_3DGS/threestudio/models/guidance/stable_diffusion_guidance.py:160-167,
// differentiated by the SDS loss :424-427). All HBM-bound fp16 NHWC sweeps with fp32 math:
// GroupNorm statistics for tensors of up to 2^20 pixels, GroupNorm(+SiLU) backward, softmax
// backward, batched transposes, depth-to-space, image pre/post-processing, the diagonal-Gaussian
// sampler and its backward.
#pragma once
#include "gd_gemm.cuh"

namespace gdu {

// ---- GroupNorm statistics, finalised: stats[n*groups+g] = (mean, rstd) -----------------------
// One warp per (image, group): lanes stride over the splits (loads in flight), then a fixed-order
// xor-shuffle tree => deterministic.
__device__ __forceinline__ float2 warp_sum_parts(const float2* __restrict__ part, int splits, int lane) {
  float s = 0.f, ss = 0.f;
  for (int k = lane; k < splits; k += 32) {
    const float2 p = part[k];
    s += p.x; ss += p.y;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(~0u, s, o); ss += __shfl_xor_sync(~0u, ss, o); }
  return make_float2(s, ss);
}
__global__ void __launch_bounds__(256)
k_gn_finalize(const float2* __restrict__ part, float2* __restrict__ stats, int total, int splits, float inv_n, float eps) {
  pdl_entry();
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= total) return;
  const float2 t = warp_sum_parts(part + (size_t)i * splits, splits, lane);
  const float mean = t.x * inv_n;
  if (lane == 0) stats[i] = make_float2(mean, rsqrtf(fmaxf(t.y * inv_n - mean * mean, 0.0f) + eps));
}

// y = GN(x) (+SiLU) with finalised statistics. grid (image, pixel chunks).
__global__ void __launch_bounds__(256)
k_gn_apply_final(const __half* __restrict__ x, __half* __restrict__ y, const float2* __restrict__ stats,
                 const __half* __restrict__ gamma, const __half* __restrict__ beta, int HW, int C, int groups,
                 int do_silu, int pix_per_cta) {
  pdl_entry();
  extern __shared__ float2 s_ab[];  // [C] (scale, shift)
  const int n = blockIdx.x, cpg = C / groups;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float2 mr = stats[(size_t)n * groups + c / cpg];
    const float a = mr.y * __half2float(gamma[c]);
    s_ab[c] = make_float2(a, __half2float(beta[c]) - mr.x * a);
  }
  __syncthreads();
  const int C8 = C >> 3;
  const long long p0 = (long long)blockIdx.y * pix_per_cta, p1 = min((long long)HW, p0 + pix_per_cta);
  const uint4* xb = reinterpret_cast<const uint4*>(x + (size_t)n * HW * C);
  uint4* yb = reinterpret_cast<uint4*>(y + (size_t)n * HW * C);
  for (long long i = p0 * C8 + threadIdx.x; i < p1 * C8; i += blockDim.x) {
    const int c0 = (int)(i % C8) << 3;
    uint4 v = xb[i];
    __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const float2 f = __half22float2(h[k]);
      const float2 ab0 = s_ab[c0 + 2 * k], ab1 = s_ab[c0 + 2 * k + 1];
      float a = f.x * ab0.x + ab0.y, b = f.y * ab1.x + ab1.y;
      if (do_silu) { a = silu(a); b = silu(b); }
      h[k] = __floats2half2_rn(a, b);
    }
    yb[i] = v;
  }
}

// ---- GroupNorm (+SiLU) backward --------------------------------------------------------------
// z = act(GN(x)), upstream dz. With xh = (x - mean) * rstd, y = xh*gamma + beta,
// g = dz * act'(y):   dx = rstd * (gamma*g - S1 - xh*S2),  S1 = mean_group(gamma*g),
// S2 = mean_group(gamma*g*xh).
__device__ __forceinline__ float dsilu(float y) { return dsilu_tanh(y); }   // one SFU op (gd_gemm.cuh)
// Pass 1: same sweep as k_gn_stats; per channel sum(g), sum(g*xh) folded with gamma per group.
__global__ void __launch_bounds__(256)
k_gn_bwd_stats(const __half* __restrict__ x, const __half* __restrict__ dz, const float2* __restrict__ stats,
               const __half* __restrict__ gamma, const __half* __restrict__ beta, float2* __restrict__ part,
               int HW, int C, int groups, int splits, int do_silu) {
  pdl_entry();
  __shared__ float s_c[2][2560];
  const int n = blockIdx.x, sp = blockIdx.y;
  const int cpg = C / groups, C8 = C >> 3;
  const int p0 = (int)((long long)HW * sp / splits), p1 = (int)((long long)HW * (sp + 1) / splits);
  const int pix_par = C8 <= 256 ? 256 / C8 : 1;
  const int iters = C8 <= 256 ? 1 : (C8 + 255) / 256;
  const uint4* xb = reinterpret_cast<const uint4*>(x + (size_t)n * HW * C);
  const uint4* gb = reinterpret_cast<const uint4*>(dz + (size_t)n * HW * C);
  for (int it = 0; it < iters; it++) {
    const int chunk = C8 <= 256 ? (int)(threadIdx.x % C8) : (int)threadIdx.x + 256 * it;
    const int pl = C8 <= 256 ? (int)(threadIdx.x / C8) : 0;
    if (pl >= pix_par || chunk >= C8) continue;
    float a[8], b[8], s[8], ss[8];
    float gam[8], bet[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const float2 mr = stats[(size_t)n * groups + (chunk * 8 + k) / cpg];
      a[k] = mr.y; b[k] = -mr.x * mr.y;            // xh = x*a + b
      gam[k] = __half2float(gamma[chunk * 8 + k]); bet[k] = __half2float(beta[chunk * 8 + k]);
      s[k] = 0.f; ss[k] = 0.f;
    }
#pragma unroll 4
    for (int pix = p0 + pl; pix < p1; pix += pix_par) {
      const uint4 xv = xb[(size_t)pix * C8 + chunk], gv = gb[(size_t)pix * C8 + chunk];
      const __half2* xh2 = reinterpret_cast<const __half2*>(&xv);
      const __half2* gh2 = reinterpret_cast<const __half2*>(&gv);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const float2 xf = __half22float2(xh2[k]), gf = __half22float2(gh2[k]);
        const float xh0 = xf.x * a[2 * k] + b[2 * k], xh1 = xf.y * a[2 * k + 1] + b[2 * k + 1];
        float g0 = gf.x, g1 = gf.y;
        if (do_silu) { g0 *= dsilu(xh0 * gam[2 * k] + bet[2 * k]); g1 *= dsilu(xh1 * gam[2 * k + 1] + bet[2 * k + 1]); }
        s[2 * k] += g0; ss[2 * k] += g0 * xh0;
        s[2 * k + 1] += g1; ss[2 * k + 1] += g1 * xh1;
      }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
      s_c[0][pl * C + chunk * 8 + k] = s[k] * gam[k];
      s_c[1][pl * C + chunk * 8 + k] = ss[k] * gam[k];
    }
  }
  __syncthreads();
  if (threadIdx.x < groups) {
    const int g = threadIdx.x;
    float u = 0.f, v = 0.f;
    for (int pl = 0; pl < pix_par; pl++)
      for (int c = g * cpg; c < (g + 1) * cpg; c++) { u += s_c[0][pl * C + c]; v += s_c[1][pl * C + c]; }
    part[((size_t)n * groups + g) * splits + sp] = make_float2(u, v);
  }
}
// Pass 1, fast flavour (C/8 divides 256: the VAE's 128 / 256 / 512 channels): a thread owns ONE 8-channel chunk for the
// whole sweep -- coefficients in registers (xh = x*ca + cb, y = x*ya + yb), no index arithmetic -- with the 16-byte loads of
// the next U pixels in flight during the math. Same partial layout as k_gn_bwd_stats.
template <bool SILU, int U, bool PIPE>
__global__ void __launch_bounds__(256)
k_gn_bwd_stats_fast(const __half* __restrict__ x, const __half* __restrict__ dz, const float2* __restrict__ stats,
                    const __half* __restrict__ gamma, const __half* __restrict__ beta, float2* __restrict__ part,
                    int HW, int C, int groups, int splits) {
  pdl_entry();
  __shared__ float s_c[2][2048];   // [sum g*gamma | sum g*gamma*xh][pl * C + c]   (PP * C = 2048)
  const int n = blockIdx.x, sp = blockIdx.y;
  const int cpg = C / groups, C8 = C >> 3, PP = 256 / C8;
  const int ch = threadIdx.x % C8, pl = threadIdx.x / C8;
  const int p0 = (int)((long long)HW * sp / splits), p1 = (int)((long long)HW * (sp + 1) / splits);
  float ca[8], cb[8], ya[8], yb[8], gam[8], sg[8], sx[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int c = ch * 8 + k;
    const float2 mr = stats[(size_t)n * groups + c / cpg];
    gam[k] = __half2float(gamma[c]);
    ca[k] = mr.y; cb[k] = -mr.x * mr.y;
    ya[k] = ca[k] * gam[k]; yb[k] = fmaf(cb[k], gam[k], __half2float(beta[c]));
    sg[k] = 0.f; sx[k] = 0.f;
  }
  const size_t base = (size_t)n * HW * C;
  const uint4* xb = reinterpret_cast<const uint4*>(x + base) + ch;
  const uint4* gb = reinterpret_cast<const uint4*>(dz + base) + ch;
  // PIPE: the loads of the next U pixels are in flight during the math (2 x 2U x 16 B per thread held in registers);
  // !PIPE: 2U loads issued together, the other resident CTAs cover the math phase.
  uint4 xv[U], gv[U], nxv[PIPE ? U : 1], ngv[PIPE ? U : 1];
  auto fetch = [&](int pp, uint4* X, uint4* G) {
#pragma unroll
    for (int u = 0; u < U; u++) {
      const bool ok = pp + u * PP < p1;
      const size_t o = (size_t)(pp + u * PP) * C8;
      X[u] = ok ? __ldg(xb + o) : make_uint4(0, 0, 0, 0);
      G[u] = ok ? __ldg(gb + o) : make_uint4(0, 0, 0, 0);   // zero gradient: padding pixels add nothing
    }
  };
  if (PIPE) fetch(p0 + pl, nxv, ngv);
  for (int pix = p0 + pl; pix < p1; pix += PP * U) {
    if (PIPE) {
#pragma unroll
      for (int u = 0; u < U; u++) { xv[u] = nxv[u]; gv[u] = ngv[u]; }
      fetch(pix + PP * U, nxv, ngv);
    } else {
      fetch(pix, xv, gv);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const __half2* xh2 = reinterpret_cast<const __half2*>(&xv[u]);
      const __half2* gh2 = reinterpret_cast<const __half2*>(&gv[u]);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const float2 xf = __half22float2(xh2[k]), gf = __half22float2(gh2[k]);
        float g0 = gf.x, g1 = gf.y;
        if (SILU) { g0 *= dsilu(fmaf(xf.x, ya[2 * k], yb[2 * k])); g1 *= dsilu(fmaf(xf.y, ya[2 * k + 1], yb[2 * k + 1])); }
        sg[2 * k] += g0; sx[2 * k] = fmaf(g0, fmaf(xf.x, ca[2 * k], cb[2 * k]), sx[2 * k]);
        sg[2 * k + 1] += g1; sx[2 * k + 1] = fmaf(g1, fmaf(xf.y, ca[2 * k + 1], cb[2 * k + 1]), sx[2 * k + 1]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; k++) {
    s_c[0][pl * C + ch * 8 + k] = sg[k] * gam[k];
    s_c[1][pl * C + ch * 8 + k] = sx[k] * gam[k];
  }
  __syncthreads();
  if ((int)threadIdx.x < groups) {
    const int g = threadIdx.x;
    float u = 0.f, v = 0.f;
    for (int q = 0; q < PP; q++)
      for (int c = g * cpg; c < (g + 1) * cpg; c++) { u += s_c[0][q * C + c]; v += s_c[1][q * C + c]; }
    part[((size_t)n * groups + g) * splits + sp] = make_float2(u, v);
  }
}
__global__ void __launch_bounds__(256)
k_gn_bwd_finalize(const float2* __restrict__ part, float2* __restrict__ bstats, int total, int splits, float inv_n) {
  pdl_entry();
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= total) return;
  const float2 t = warp_sum_parts(part + (size_t)i * splits, splits, lane);
  if (lane == 0) bstats[i] = make_float2(t.x * inv_n, t.y * inv_n);
}
// Pass 2: dx = rstd*(gamma*g - S1 - xh*S2) (+ add).
__global__ void __launch_bounds__(256)
k_gn_bwd_apply(const __half* __restrict__ x, const __half* __restrict__ dz, const __half* __restrict__ add,
               __half* __restrict__ dx, const float2* __restrict__ stats, const float2* __restrict__ bstats,
               const __half* __restrict__ gamma, const __half* __restrict__ beta, int HW, int C, int groups,
               int do_silu, int pix_per_cta) {
  pdl_entry();
  extern __shared__ float4 s_p[];   // [C] (a = rstd, b = -mean*rstd, gamma, beta), then float2 [C] (rstd*S1, rstd*S2)
  float2* s_g = reinterpret_cast<float2*>(s_p + C);
  const int n = blockIdx.x, cpg = C / groups;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float2 mr = stats[(size_t)n * groups + c / cpg], bs = bstats[(size_t)n * groups + c / cpg];
    s_p[c] = make_float4(mr.y, -mr.x * mr.y, __half2float(gamma[c]), __half2float(beta[c]));
    s_g[c] = make_float2(mr.y * bs.x, mr.y * bs.y);
  }
  __syncthreads();
  const int C8 = C >> 3;
  const long long p0 = (long long)blockIdx.y * pix_per_cta, p1 = min((long long)HW, p0 + pix_per_cta);
  const size_t base = (size_t)n * HW * C;
  const uint4* xb = reinterpret_cast<const uint4*>(x + base);
  const uint4* gb = reinterpret_cast<const uint4*>(dz + base);
  const uint4* ab = add ? reinterpret_cast<const uint4*>(add + base) : nullptr;
  uint4* ob = reinterpret_cast<uint4*>(dx + base);
  for (long long i = p0 * C8 + threadIdx.x; i < p1 * C8; i += blockDim.x) {
    const int c0 = (int)(i % C8) << 3;
    const uint4 xv = xb[i], gv = gb[i];
    uint4 av = ab ? ab[i] : make_uint4(0, 0, 0, 0);
    const __half2* xh2 = reinterpret_cast<const __half2*>(&xv);
    const __half2* gh2 = reinterpret_cast<const __half2*>(&gv);
    __half2* ah2 = reinterpret_cast<__half2*>(&av);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const float2 xf = __half22float2(xh2[k]), gf = __half22float2(gh2[k]), af = __half22float2(ah2[k]);
      const float4 q0 = s_p[c0 + 2 * k], q1 = s_p[c0 + 2 * k + 1];
      const float xh0 = xf.x * q0.x + q0.y, xh1 = xf.y * q1.x + q1.y;
      float g0 = gf.x, g1 = gf.y;
      if (do_silu) { g0 *= dsilu(xh0 * q0.z + q0.w); g1 *= dsilu(xh1 * q1.z + q1.w); }
      const float2 sg0 = s_g[c0 + 2 * k], sg1 = s_g[c0 + 2 * k + 1];
      const float d0 = q0.x * q0.z * g0 - sg0.x - xh0 * sg0.y + af.x;
      const float d1 = q1.x * q1.z * g1 - sg1.x - xh1 * sg1.y + af.y;
      ah2[k] = __floats2half2_rn(d0, d1);
    }
    ob[i] = av;
  }
}

// ---- fast sweeps for C/8 dividing 256 (C = 64..2048 powers of two: every VAE layer) ----------
// A thread owns one 8-channel chunk for its whole life (per-channel coefficients in registers,
// no index arithmetic in the loop) and keeps U independent 16-byte loads per tensor in flight.
template <bool SILU, int U>
__global__ void __launch_bounds__(256)
k_gn_apply_fast(const __half* __restrict__ x, __half* __restrict__ y, const float2* __restrict__ stats,
                const __half* __restrict__ gamma, const __half* __restrict__ beta, int HW, int C, int groups, int pix_per_cta) {
  pdl_entry();
  const int n = blockIdx.x, cpg = C / groups, C8 = C >> 3, PP = 256 / C8;
  const int ch = threadIdx.x % C8, pl = threadIdx.x / C8;
  float sa[8], sb[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int c = ch * 8 + k;
    const float2 mr = stats[(size_t)n * groups + c / cpg];
    sa[k] = mr.y * __half2float(gamma[c]);
    sb[k] = __half2float(beta[c]) - mr.x * sa[k];
  }
  const int p0 = blockIdx.y * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
  const uint4* xb = reinterpret_cast<const uint4*>(x + (size_t)n * HW * C) + ch;
  uint4* yb = reinterpret_cast<uint4*>(y + (size_t)n * HW * C) + ch;
  uint4 v[U], nx[U];
#pragma unroll
  for (int u = 0; u < U; u++) nx[u] = (p0 + pl + u * PP < p1) ? xb[(size_t)(p0 + pl + u * PP) * C8] : make_uint4(0, 0, 0, 0);
  for (int pix = p0 + pl; pix < p1; pix += PP * U) {
#pragma unroll
    for (int u = 0; u < U; u++) v[u] = nx[u];
    const int np = pix + PP * U;                       // prefetch the next iteration
#pragma unroll
    for (int u = 0; u < U; u++) if (np + u * PP < p1) nx[u] = xb[(size_t)(np + u * PP) * C8];
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (pix + u * PP >= p1) break;
      __half2* h = reinterpret_cast<__half2*>(&v[u]);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const float2 f = __half22float2(h[k]);
        float a = f.x * sa[2 * k] + sb[2 * k], b = f.y * sa[2 * k + 1] + sb[2 * k + 1];
        if (SILU) { a = silu(a); b = silu(b); }
        h[k] = __floats2half2_rn(a, b);
      }
      yb[(size_t)(pix + u * PP) * C8] = v[u];
    }
  }
}

template <bool SILU, int U>
__global__ void __launch_bounds__(256)
k_gn_bwd_apply_fast(const __half* __restrict__ x, const __half* dz, const __half* __restrict__ add, __half* dx,
                    const float2* __restrict__ stats, const float2* __restrict__ bstats, const __half* __restrict__ gamma,
                    const __half* __restrict__ beta, int HW, int C, int groups, int pix_per_cta) {
  pdl_entry();
  const int n = blockIdx.x, cpg = C / groups, C8 = C >> 3, PP = 256 / C8;
  const int ch = threadIdx.x % C8, pl = threadIdx.x / C8;
  float ca[8], cb[8], cg[8], ce[8], s1[8], s2[8];   // xh = x*ca+cb; y = xh*cg+ce; dx = ca*cg*g - s1 - xh*s2
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int c = ch * 8 + k;
    const float2 mr = stats[(size_t)n * groups + c / cpg], bs = bstats[(size_t)n * groups + c / cpg];
    ca[k] = mr.y; cb[k] = -mr.x * mr.y;
    cg[k] = __half2float(gamma[c]); ce[k] = __half2float(beta[c]);
    s1[k] = mr.y * bs.x; s2[k] = mr.y * bs.y;
  }
  const int p0 = blockIdx.y * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
  const size_t base = (size_t)n * HW * C;
  const uint4* xb = reinterpret_cast<const uint4*>(x + base) + ch;
  const uint4* gb = reinterpret_cast<const uint4*>(dz + base) + ch;
  const uint4* ab = add ? reinterpret_cast<const uint4*>(add + base) + ch : nullptr;
  uint4* ob = reinterpret_cast<uint4*>(dx + base) + ch;
  uint4 xv[U], gv[U], av[U], nxv[U], ngv[U], nav[U];
  auto fetch = [&](int pp, uint4 (&X)[U], uint4 (&G)[U], uint4 (&A)[U]) {
#pragma unroll
    for (int u = 0; u < U; u++) {
      const bool ok = pp + u * PP < p1;
      const size_t o = (size_t)(pp + u * PP) * C8;
      X[u] = ok ? xb[o] : make_uint4(0, 0, 0, 0);
      G[u] = ok ? gb[o] : make_uint4(0, 0, 0, 0);
      A[u] = (ok && ab) ? ab[o] : make_uint4(0, 0, 0, 0);
    }
  };
  fetch(p0 + pl, nxv, ngv, nav);
  for (int pix = p0 + pl; pix < p1; pix += PP * U) {
#pragma unroll
    for (int u = 0; u < U; u++) { xv[u] = nxv[u]; gv[u] = ngv[u]; av[u] = nav[u]; }
    fetch(pix + PP * U, nxv, ngv, nav);                // next iteration's loads in flight during the math
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (pix + u * PP >= p1) break;
      const __half2* xh2 = reinterpret_cast<const __half2*>(&xv[u]);
      const __half2* gh2 = reinterpret_cast<const __half2*>(&gv[u]);
      __half2* ah2 = reinterpret_cast<__half2*>(&av[u]);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const float2 xf = __half22float2(xh2[k]), gf = __half22float2(gh2[k]), af = __half22float2(ah2[k]);
        const float xh0 = xf.x * ca[2 * k] + cb[2 * k], xh1 = xf.y * ca[2 * k + 1] + cb[2 * k + 1];
        float g0 = gf.x, g1 = gf.y;
        if (SILU) { g0 *= dsilu(xh0 * cg[2 * k] + ce[2 * k]); g1 *= dsilu(xh1 * cg[2 * k + 1] + ce[2 * k + 1]); }
        const float d0 = ca[2 * k] * cg[2 * k] * g0 - s1[2 * k] - xh0 * s2[2 * k] + af.x;
        const float d1 = ca[2 * k + 1] * cg[2 * k + 1] * g1 - s1[2 * k + 1] - xh1 * s2[2 * k + 1] + af.y;
        ah2[k] = __floats2half2_rn(d0, d1);
      }
      ob[(size_t)(pix + u * PP) * C8] = av[u];
    }
  }
}

// ---- softmax backward in place: dS = P * (dP - sum_j P_j dP_j); one warp per row -----------------
__global__ void __launch_bounds__(256)
k_softmax_bwd(const __half* __restrict__ P, __half* __restrict__ dP, long long rows, int cols, long long ld) {
  pdl_entry();
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const uint4* p4 = reinterpret_cast<const uint4*>(P + row * ld);
  uint4* d4 = reinterpret_cast<uint4*>(dP + row * ld);
  const int n8 = cols >> 3;   // cols % 8 == 0 (host-checked)
  float dot = 0.f;
  for (int c = lane; c < n8; c += 32) {
    const uint4 pv = p4[c], dv = d4[c];
    const __half2* ph = reinterpret_cast<const __half2*>(&pv);
    const __half2* dh = reinterpret_cast<const __half2*>(&dv);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const float2 a = __half22float2(ph[k]), b = __half22float2(dh[k]);
      dot += a.x * b.x + a.y * b.y;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(~0u, dot, o);
  for (int c = lane; c < n8; c += 32) {
    const uint4 pv = p4[c];
    uint4 dv = d4[c];
    const __half2* ph = reinterpret_cast<const __half2*>(&pv);
    __half2* dh = reinterpret_cast<__half2*>(&dv);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const float2 a = __half22float2(ph[k]), b = __half22float2(dh[k]);
      dh[k] = __floats2half2_rn(a.x * (b.x - dot), a.y * (b.y - dot));
    }
    d4[c] = dv;
  }
}

// ---- batched transpose: x [B, R, C] -> y [B, C, R] (fp16), 64x64 tiles through shared memory ----
__global__ void __launch_bounds__(256)
k_transpose(const __half* __restrict__ x, __half* __restrict__ y, int R, int C) {
  pdl_entry();
  __shared__ __half t[64][66];
  const size_t base = (size_t)blockIdx.z * R * C;
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int j = ty; j < 64; j += 8) {
    const int r = r0 + j, c = c0 + 2 * tx;
    __half2 v = __floats2half2_rn(0.f, 0.f);
    if (r < R && c + 1 < C) v = *reinterpret_cast<const __half2*>(x + base + (size_t)r * C + c);
    else if (r < R && c < C) v = __halves2half2(x[base + (size_t)r * C + c], __float2half_rn(0.f));
    t[j][2 * tx] = __low2half(v); t[j][2 * tx + 1] = __high2half(v);
  }
  __syncthreads();
  for (int j = ty; j < 64; j += 8) {
    const int c = c0 + j, r = r0 + 2 * tx;
    if (c >= C) continue;
    if (r + 1 < R) *reinterpret_cast<__half2*>(y + base + (size_t)c * R + r) = __halves2half2(t[2 * tx][j], t[2 * tx + 1][j]);
    else if (r < R) y[base + (size_t)c * R + r] = t[2 * tx][j];
  }
}

// ---- depth-to-space (inverse of k_space_to_depth): [N,H/2,W/2,4C] -> [N,H,W,C] ---------------
__global__ void k_depth_to_space(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H, int W, int C8) {
  pdl_entry();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)N * H * W * C8;
  if (i >= total) return;
  const int c = (int)(i % C8);
  long long p = i / C8;
  const int ix = (int)(p % W); p /= W;
  const int iy = (int)(p % H);
  const int n = (int)(p / H);
  const int ph = (iy & 1) * 2 + (ix & 1);
  y[i] = x[((((long long)n * (H >> 1) + (iy >> 1)) * (W >> 1) + (ix >> 1)) * 4 + ph) * C8 + c];
}

// ---- image pre-processing: color fp32 NCHW [B,3,H,W] -> fp16 NCHW [B,4,H,W] = a*c+sh | 0 ----
__global__ void k_vae_prep(const float* __restrict__ color, __half* __restrict__ y, int B, long long HW, float a, float sh) {
  pdl_entry();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * 4 * HW) return;
  const long long p = i % HW;
  const int c = (int)((i / HW) % 4), b = (int)(i / (4 * HW));
  y[i] = __float2half_rn(c < 3 ? a * color[((long long)b * 3 + c) * HW + p] + sh : 0.0f);
}
// ---- DiagonalGaussianDistribution.sample() * scaling: moments fp16 NHWC [B,h,w,8] (mean|logvar) ----
__global__ void k_vae_sample(const __half* __restrict__ mom, const float* __restrict__ noise, float* __restrict__ lat,
                             int B, int hw, float scaling) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 4 * hw) return;
  const int p = i % hw, k = (i / hw) % 4, b = i / (4 * hw);
  const __half* m = mom + ((size_t)b * hw + p) * 8;
  const float mean = __half2float(m[k]);
  const float logvar = fminf(fmaxf(__half2float(m[4 + k]), -30.0f), 20.0f);
  // the reference samples in fp16: round the sample like posterior.sample() does
  const float s = __half2float(__float2half_rn(mean + expf(0.5f * logvar) * noise[i]));
  lat[i] = __half2float(__float2half_rn(s * scaling));
}
// Backward of the sampler: grad fp32 NCHW [B,4,h,w] (nan_to_num + clamp(+-clip) applied here,
// stable_diffusion_guidance.py:418-421) -> d moments fp16 NHWC [B,h,w,Cp] (8 real channels, the
// rest zero so that the tensor is a valid K operand of the conv-GEMM). `gscale` = loss scale.
// max |nan_to_num(clamp(g))| per 1024-element block, then the power-of-two loss scale (see gd_vae_grad_scale)
__global__ void __launch_bounds__(256)
k_vae_grad_absmax(const float* __restrict__ grad, long long n, float clip, float* __restrict__ scratch) {
  pdl_entry();
  __shared__ float s_w[8];
  float m = 0.f;
  const long long base = (long long)blockIdx.x * 1024;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const long long i = base + k * 256 + threadIdx.x;
    if (i < n) {
      float g = grad[i];
      if (!(g == g)) g = 0.f;
      g = fminf(fabsf(g), 3.4028235e38f);
      if (clip > 0.f) g = fminf(g, clip);
      m = fmaxf(m, g);
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(~0u, m, o));
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; w++) m = fmaxf(m, s_w[w]);
    scratch[blockIdx.x] = m;
  }
}
__global__ void __launch_bounds__(256)
k_vae_grad_scale(int nblk, const float* __restrict__ scratch, float pre, float target, float* __restrict__ dyn) {
  pdl_entry();
  __shared__ float s_w[8];
  float m = 0.f;
  for (int k = threadIdx.x; k < nblk; k += 256) m = fmaxf(m, scratch[k]);
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(~0u, m, o));
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; w++) m = fmaxf(m, s_w[w]);
    const float top = m * fabsf(pre);
    float e = 0.f;
    if (top > 0.f && top < 3.0e38f) e = floorf(log2f(target / top));
    *dyn = exp2f(fminf(fmaxf(e, -24.f), 24.f));
  }
}
__global__ void k_vae_sample_bwd(const float* __restrict__ grad, const __half* __restrict__ mom,
                                 const float* __restrict__ noise, __half* __restrict__ dmom, int B, int hw, int Cp,
                                 float scaling, float clip, float gscale, const float* __restrict__ dyn) {
  pdl_entry();
  if (dyn) gscale *= *dyn;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * hw) return;
  const int p = i % hw, b = i / hw;
  const __half* m = mom + (size_t)i * 8;
  __half* d = dmom + (size_t)i * Cp;
  __align__(16) __half o[8];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const size_t gi = ((size_t)b * 4 + k) * hw + p;
    float g = grad[gi];
    if (!(g == g)) g = 0.f;
    g = fminf(fmaxf(g, -3.4028235e38f), 3.4028235e38f);
    if (clip > 0.f) g = fminf(fmaxf(g, -clip), clip);
    g *= scaling * gscale;
    const float lv = __half2float(m[4 + k]);
    const bool inside = lv >= -30.0f && lv <= 20.0f;   // clamp passes gradient inside its range
    const float std_ = expf(0.5f * fminf(fmaxf(lv, -30.0f), 20.0f));
    // saturate instead of overflowing to inf (fp16 max 65504): an inf here turns the whole chain into NaN
    o[k] = __float2half_rn(fminf(fmaxf(g, -65504.f), 65504.f));
    o[4 + k] = __float2half_rn(inside ? fminf(fmaxf(g * noise[gi] * 0.5f * std_, -65504.f), 65504.f) : 0.0f);
  }
  *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(o);
  for (int c = 8; c < Cp; c += 8) *reinterpret_cast<uint4*>(d + c) = make_uint4(0, 0, 0, 0);
}
// d image: fp16 NHWC [B,H,W,Cp] (3 real channels) -> fp32 NCHW [B,3,H,W] * scale
// (scale = 2 / loss-scale: d(2c-1)/dc and the unscaling).
__global__ void k_vae_dimg(const __half* __restrict__ dx, float* __restrict__ dcolor, int B, long long HW, int Cp, float scale) {
  pdl_entry();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * HW) return;
  const long long p = i % HW;
  const int b = (int)(i / HW);
  const uint2 v = *reinterpret_cast<const uint2*>(dx + (size_t)i * Cp);
  const __half2* h = reinterpret_cast<const __half2*>(&v);
  const float2 a = __half22float2(h[0]), c = __half22float2(h[1]);
  float* d = dcolor + (size_t)b * 3 * HW + p;
  d[0] = a.x * scale; d[HW] = a.y * scale; d[2 * HW] = c.x * scale;
}

// ---- conv_in as a GEMM: im2col of the 3-channel image, K = 27 taps padded to 64 -------------------
// A[(b,y,x), (ky*3+kx)*3+c] = a*color[b,c,y+ky-1,x+kx-1] + sh (0 outside the image: the conv pads the
// normalised image with zeros). One thread per pixel writes its 128-byte row.
__global__ void __launch_bounds__(256)
k_vae_im2col(const float* __restrict__ color, __half* __restrict__ A, int B, int H, int W, float a, float sh) {
  pdl_entry();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (pixel, 16-byte chunk of its row)
  const long long HW = (long long)H * W, i = t >> 3;
  const int u = (int)(t & 7);
  if (i >= (long long)B * HW) return;
  const int x = (int)(i % W), y = (int)((i / W) % H), b = (int)(i / HW);
  __align__(16) __half v[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int k = u * 8 + j, tap = k / 3, c = k - tap * 3;
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    const bool ok = k < 27 && yy >= 0 && yy < H && xx >= 0 && xx < W;
    v[j] = __float2half_rn(ok ? a * color[((long long)b * 3 + c) * HW + (long long)yy * W + xx] + sh : 0.f);
  }
  reinterpret_cast<uint4*>(A)[t] = *reinterpret_cast<const uint4*>(v);
}
// ---- conv_in data gradient from the per-pixel tap products Z[(b,y,x), (ky*3+kx)*3+c] (fp16, 32 wide):
// dcolor[b,c,y,x] = scale * sum_{ky,kx} Z[(b, y-ky+1, x-kx+1), (ky*3+kx)*3+c]
__global__ void __launch_bounds__(256)
k_vae_dimg_gather(const __half* __restrict__ Z, float* __restrict__ dcolor, int B, int H, int W, float scale,
                  const float* __restrict__ dyn) {
  pdl_entry();
  if (dyn) scale /= *dyn;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long HW = (long long)H * W;
  if (i >= (long long)B * HW) return;
  const int x = (int)(i % W), y = (int)((i / W) % H), b = (int)(i / HW);
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int ky = 0; ky < 3; ky++)
#pragma unroll
    for (int kx = 0; kx < 3; kx++) {
      const int yy = y - ky + 1, xx = x - kx + 1;
      if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
      const __half* z = Z + (((long long)b * H + yy) * W + xx) * 32 + (ky * 3 + kx) * 3;
#pragma unroll
      for (int c = 0; c < 3; c++) acc[c] += __half2float(z[c]);
    }
  float* d = dcolor + (long long)b * 3 * HW + (long long)y * W + x;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    float v = acc[c] * scale;
    if (dyn && !(fabsf(v) <= 3.4028235e38f)) v = 0.f;   // NaN / inf out of an overflowed fp16 chain -> 0
    d[c * HW] = v;
  }
}

// ---- bilinear resize either side of the VAE ------------------------------------------
// F.interpolate(rgb_BCHW, (512, 512), mode="bilinear", align_corners=False) of the reference guidance
// (stable_diffusion_guidance.py:387-396; 1024^2 renders -> 512^2 in the shipped config) and its transpose for the
// gradient. PyTorch's rule: src = max((dst + 0.5) * in/out - 0.5, 0), i0 = floor(src), i1 = min(i0 + 1, in - 1).
__device__ __forceinline__ void bil_src(int d, float scale, int n_in, int& i0, int& i1, float& w1) {
  const float sidx = fmaxf(((float)d + 0.5f) * scale - 0.5f, 0.0f);
  i0 = min((int)sidx, n_in - 1);
  i1 = min(i0 + 1, n_in - 1);
  w1 = sidx - (float)i0;
}
__global__ void __launch_bounds__(256)
k_resize_bilinear(const float* __restrict__ in, float* __restrict__ out, int BC, int Hi, int Wi, int Ho, int Wo) {
  pdl_entry();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)BC * Ho * Wo) return;
  const int x = (int)(i % Wo), y = (int)((i / Wo) % Ho), bc = (int)(i / ((long long)Wo * Ho));
  int y0, y1, x0, x1;
  float wy, wx;
  bil_src(y, (float)Hi / (float)Ho, Hi, y0, y1, wy);
  bil_src(x, (float)Wi / (float)Wo, Wi, x0, x1, wx);
  const float* p = in + (long long)bc * Hi * Wi;
  const float v00 = p[(long long)y0 * Wi + x0], v01 = p[(long long)y0 * Wi + x1];
  const float v10 = p[(long long)y1 * Wi + x0], v11 = p[(long long)y1 * Wi + x1];
  out[i] = (1.f - wy) * ((1.f - wx) * v00 + wx * v01) + wy * ((1.f - wx) * v10 + wx * v11);
}
// Transpose as a GATHER (deterministic, no atomics): an input pixel sums the weights of the few output pixels whose
// 2x2 footprint contains it.
__global__ void __launch_bounds__(256)
k_resize_bilinear_bwd(const float* __restrict__ dout, float* __restrict__ din, int BC, int Hi, int Wi, int Ho, int Wo) {
  pdl_entry();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)BC * Hi * Wi) return;
  const int x = (int)(i % Wi), y = (int)((i / Wi) % Hi), bc = (int)(i / ((long long)Wi * Hi));
  const float sy = (float)Hi / (float)Ho, sx = (float)Wi / (float)Wo;
  // candidate outputs: src in (y - 1, y + 1)  <=>  dst in ((y - 0.5) / s - 0.5, (y + 1.5) / s - 0.5)
  const int ya = max(0, (int)floorf(((float)y - 0.5f) / sy - 0.5f) - 1), yb = min(Ho - 1, (int)ceilf(((float)y + 1.5f) / sy - 0.5f) + 1);
  const int xa = max(0, (int)floorf(((float)x - 0.5f) / sx - 0.5f) - 1), xb = min(Wo - 1, (int)ceilf(((float)x + 1.5f) / sx - 0.5f) + 1);
  const float* g = dout + (long long)bc * Ho * Wo;
  float acc = 0.f;
  for (int oy = ya; oy <= yb; oy++) {
    int y0, y1; float wy;
    bil_src(oy, sy, Hi, y0, y1, wy);
    const float cy = (y0 == y ? 1.f - wy : 0.f) + (y1 == y ? wy : 0.f);
    if (cy == 0.f) continue;
    for (int ox = xa; ox <= xb; ox++) {
      int x0, x1; float wx;
      bil_src(ox, sx, Wi, x0, x1, wx);
      const float cx = (x0 == x ? 1.f - wx : 0.f) + (x1 == x ? wx : 0.f);
      if (cx != 0.f) acc += cy * cx * g[(long long)oy * Wo + ox];
    }
  }
  din[i] = acc;
}

}  // namespace gdu
