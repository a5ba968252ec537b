// Fused Gaussian-parameter kernels either side of the rasteriser.
//
// Reference behaviour (not code): GS/scene/gaussian_model.py:95-115 (activations: exp / sigmoid /
// F.normalize, evaluated per view from GS/gaussian_renderer/__init__.py:53-80), :156-165 (one
// torch.optim.Adam(eps=1e-15) over six parameter groups), :415-419 (densification statistics) and
// TS/systems/GaussianDreamer.py:263-279 (max_radii2D / viewspace gradient accumulation).
// The reference runs ~40 eager kernels per iteration for this; here it is one kernel before the
// rasteriser, one after it, and one for the statistics. All fp32, one thread per Gaussian.
#include "gd_raster_common.cuh"

namespace gd {

// packed struct-of-arrays layout shared with bench.py / parallel.py: xyz 3P | f_dc 3P | opacity P |
// scales 3P | rotation 4P  (14 floats per Gaussian)
__global__ void __launch_bounds__(256)
k_params_activate(int P, const float* __restrict__ xyz, const float* __restrict__ f_dc, const float* __restrict__ opacity,
                  const float* __restrict__ scaling, const float* __restrict__ rotation, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  float* o_xyz = out; float* o_dc = out + 3 * (size_t)P; float* o_op = out + 6 * (size_t)P;
  float* o_sc = out + 7 * (size_t)P; float* o_rot = out + 10 * (size_t)P;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    o_xyz[3 * (size_t)i + k] = xyz[3 * (size_t)i + k];
    o_dc[3 * (size_t)i + k] = f_dc[3 * (size_t)i + k];
    o_sc[3 * (size_t)i + k] = expf(scaling[3 * (size_t)i + k]);                  // torch.exp
  }
  o_op[i] = 1.0f / (1.0f + expf(-opacity[i]));                                   // torch.sigmoid
  const float4 q = *reinterpret_cast<const float4*>(rotation + 4 * (size_t)i);
  const float n = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);   // F.normalize(eps=1e-12)
  float* r = o_rot + 4 * (size_t)i;   // the packed slice starts at float 10P: only 8-byte aligned in general
  r[0] = q.x / n; r[1] = q.y / n; r[2] = q.z / n; r[3] = q.w / n;
}

struct AdamHyper {
  float lr[5];          // xyz, f_dc, opacity, scaling, rotation
  float beta1, beta2, eps;
  float bc1, bc2_sqrt;  // 1 - beta1^t, sqrt(1 - beta2^t)
};
// torch.optim.Adam single-tensor update (no weight decay, no amsgrad):
//   m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g, g, 1-b2); p.addcdiv_(m, sqrt(v)/sqrt(bc2) + eps, -lr/bc1)
__device__ __forceinline__ float adam_step(float p, float g, float& m, float& v, float lr, const AdamHyper& h) {
  m = m + (1.0f - h.beta1) * (g - m);
  v = v * h.beta2 + (1.0f - h.beta2) * g * g;
  const float denom = sqrtf(v) / h.bc2_sqrt + h.eps;
  return p - (lr / h.bc1) * (m / denom);
}
// grad: packed gradient w.r.t. the ACTIVATED parameters (rasteriser backward output, same layout
// as k_params_activate's output); the chain rule through the activations and the Adam update of
// the RAW parameters happen here. exp_avg / exp_avg_sq: packed [14P] optimiser state.
__device__ __forceinline__ void adam_one(int i, int P, float* __restrict__ xyz, float* __restrict__ f_dc, float* __restrict__ opacity,
                                         float* __restrict__ scaling, float* __restrict__ rotation, const float* __restrict__ grad,
                                         float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq, const AdamHyper& h) {
  const size_t o_dc = 3 * (size_t)P, o_op = 6 * (size_t)P, o_sc = 7 * (size_t)P, o_rot = 10 * (size_t)P;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const size_t j = 3 * (size_t)i + k;
    xyz[j] = adam_step(xyz[j], grad[j], exp_avg[j], exp_avg_sq[j], h.lr[0], h);
    f_dc[j] = adam_step(f_dc[j], grad[o_dc + j], exp_avg[o_dc + j], exp_avg_sq[o_dc + j], h.lr[1], h);
    const float s = scaling[j];
    scaling[j] = adam_step(s, grad[o_sc + j] * expf(s), exp_avg[o_sc + j], exp_avg_sq[o_sc + j], h.lr[3], h);   // d exp
  }
  {
    const float x = opacity[i], sg = 1.0f / (1.0f + expf(-x));
    opacity[i] = adam_step(x, grad[o_op + i] * sg * (1.0f - sg), exp_avg[o_op + i], exp_avg_sq[o_op + i], h.lr[2], h);
  }
  {
    const float4 q = *reinterpret_cast<const float4*>(rotation + 4 * (size_t)i);
    const float* gp = grad + o_rot + 4 * (size_t)i;
    const float4 g = make_float4(gp[0], gp[1], gp[2], gp[3]);
    const float nrm = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    const float n = fmaxf(nrm, 1e-12f);
    const float qx = q.x / n, qy = q.y / n, qz = q.z / n, qw = q.w / n;
    // backward of q / max(|q|, eps): (g - qhat (qhat.g)) / n while |q| > eps, g / eps below it
    const float dot = nrm > 1e-12f ? qx * g.x + qy * g.y + qz * g.z + qw * g.w : 0.0f;
    const float d[4] = {(g.x - qx * dot) / n, (g.y - qy * dot) / n, (g.z - qz * dot) / n, (g.w - qw * dot) / n};
    float r[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const size_t j = o_rot + 4 * (size_t)i + k;
      r[k] = adam_step(r[k], d[k], exp_avg[j], exp_avg_sq[j], h.lr[4], h);
    }
    *reinterpret_cast<float4*>(rotation + 4 * (size_t)i) = make_float4(r[0], r[1], r[2], r[3]);
  }
}
__global__ void __launch_bounds__(256)
k_params_adam(int P, float* __restrict__ xyz, float* __restrict__ f_dc, float* __restrict__ opacity,
              float* __restrict__ scaling, float* __restrict__ rotation, const float* __restrict__ grad,
              float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq, AdamHyper h) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  adam_one(i, P, xyz, f_dc, opacity, scaling, rotation, grad, exp_avg, exp_avg_sq, h);
}

// gaussian_model.py:415-419 + GaussianDreamer.py:269-275 for the view batch: visibility = max
// radius over the views > 0; accumulates |d/dmeans2D (summed over views)|_xy, the counter and
// the running maximum screen radius.
__global__ void __launch_bounds__(256)
k_densify_stats(int P, int B, const float* __restrict__ dmeans2D_sum, const int* __restrict__ radii,
                float* __restrict__ xyz_gradient_accum, float* __restrict__ denom, float* __restrict__ max_radii2D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  int r = 0;
  for (int b = 0; b < B; b++) r = max(r, radii[(size_t)b * P + i]);
  if (r <= 0) return;
  const float gx = dmeans2D_sum[3 * (size_t)i], gy = dmeans2D_sum[3 * (size_t)i + 1];
  xyz_gradient_accum[i] += sqrtf(gx * gx + gy * gy);
  denom[i] += 1.0f;
  max_radii2D[i] = fmaxf(max_radii2D[i], (float)r);
}

// ---- sparsity loss on the depth-normalised opacity -----------------------
// Reference behaviour: TS/systems/GaussianDreamer.py:215 (opacity = depths / (depths.max() + 1e-5)) and
// :253-255 (loss_sparsity = mean(sqrt(opacity^2 + 0.01))), differentiated by autograd into the depth
// images. Pass 1: elementwise gradient + per-block partial sums (fixed order); pass 2 (one block): total;
// pass 3: the term that flows through depths.max() lands on the arg-max pixel(s).
__global__ void __launch_bounds__(256)
k_sparsity_grad(long long n, float inv_ntotal, const float* __restrict__ depth, const float* __restrict__ depth_max,
                float lambda, float* __restrict__ dL_ddepth, float* __restrict__ scratch) {
  __shared__ float s_red[3][8];
  const float dmax = *depth_max, inv = 1.0f / (dmax + 1e-5f);
  float a = 0.f, bsum = 0.f, c = 0.f;
  const long long base = (long long)blockIdx.x * 1024;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const long long i = base + k * 256 + threadIdx.x;
    if (i < n) {
      const float d = depth[i], op = d * inv;
      const float f = sqrtf(op * op + 0.01f);
      const float g = lambda * inv_ntotal * op / f;     // dL/d opacity_i
      dL_ddepth[i] = g * inv;
      a += f; bsum += g * d; c += (d == dmax) ? 1.0f : 0.0f;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    a += __shfl_xor_sync(~0u, a, o); bsum += __shfl_xor_sync(~0u, bsum, o); c += __shfl_xor_sync(~0u, c, o);
  }
  if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = a; s_red[1][threadIdx.x >> 5] = bsum; s_red[2][threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) t += s_red[threadIdx.x][w];
    scratch[3 * (size_t)blockIdx.x + threadIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256)
k_sparsity_sum(int nblk, const float* __restrict__ scratch, float* __restrict__ stats) {
  __shared__ double s_red[3][256];
  double t[3] = {0.0, 0.0, 0.0};
  for (int k = threadIdx.x; k < nblk; k += 256)
#pragma unroll
    for (int j = 0; j < 3; j++) t[j] += (double)scratch[3 * (size_t)k + j];
#pragma unroll
  for (int j = 0; j < 3; j++) s_red[j][threadIdx.x] = t[j];
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if ((int)threadIdx.x < o)
#pragma unroll
      for (int j = 0; j < 3; j++) s_red[j][threadIdx.x] += s_red[j][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x < 3) stats[threadIdx.x] = (float)s_red[threadIdx.x][0];
}
__global__ void __launch_bounds__(256)
k_sparsity_finish(long long n, float inv_ntotal, const float* __restrict__ depth, const float* __restrict__ depth_max,
                  float lambda, const float* __restrict__ stats, float* __restrict__ dL_ddepth, float* __restrict__ loss_out) {
  const float dmax = *depth_max, inv = 1.0f / (dmax + 1e-5f);
  const float corr = -stats[1] * inv * inv / fmaxf(stats[2], 1.0f);
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i == 0 && loss_out) loss_out[0] = lambda * stats[0] * inv_ntotal;
  if (i < n && depth[i] == dmax) dL_ddepth[i] += corr;
}
__global__ void __launch_bounds__(256)
k_radii_max(int P, int B, const int* __restrict__ radii, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  int r = radii[i];
  for (int b = 1; b < B; b++) r = max(r, radii[(size_t)b * P + i]);
  out[i] = r;
}

// ---- batched camera construction ----------------------------------------
// Reference behaviour: GS/scene/cameras.py:50-53 + GS/utils/graphics_utils.py:59-101, evaluated on
// the CPU per view per iteration (two 4x4 LU inversions, two H2D copies, a GPU 4x4 inverse). Here:
// one thread per camera; world_view_transform = [[R^T, t],[0,1]]^T (the reference's double inversion
// with translate = 0, scale = 1 is the identity), full_proj = world_view @ P^T, camera_center =
// -(R^T)^-1 t by the 3x3 adjugate. Output row: view 16 | full_proj 16 | campos 3 (row-vector
// convention, i.e. the transposed matrices the rasteriser reads).
struct CamIntrinsics {
  float tan_half_fovx[GD_MAX_VIEWS], tan_half_fovy[GD_MAX_VIEWS];
  float znear, zfar;
};
__global__ void k_cameras_from_c2w(int B, const float* __restrict__ c2w, CamIntrinsics in, float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* m = c2w + 16 * (size_t)b;   // row-major 4x4; R = m[:3,:3], t = m[:3,3]
  float Rt[4][4] = {{m[0], m[4], m[8], m[3]}, {m[1], m[5], m[9], m[7]}, {m[2], m[6], m[10], m[11]}, {0.f, 0.f, 0.f, 1.f}};
  float* o = out + 35 * (size_t)b;
  float wv[4][4];
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int c = 0; c < 4; c++) { wv[r][c] = Rt[c][r]; o[r * 4 + c] = wv[r][c]; }   // transpose
  // P^T of getProjectionMatrix (symmetric frustum: P[0][2] = P[1][2] = 0)
  const float zn = in.znear, zf = in.zfar;
  const float top = in.tan_half_fovy[b] * zn, right = in.tan_half_fovx[b] * zn;
  float PT[4][4] = {{2.0f * zn / (right - (-right)), 0.f, 0.f, 0.f},
                    {0.f, 2.0f * zn / (top - (-top)), 0.f, 0.f},
                    {(right + (-right)) / (right - (-right)), (top + (-top)) / (top - (-top)), zf / (zf - zn), 1.0f},
                    {0.f, 0.f, -(zf * zn) / (zf - zn), 0.f}};
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int c = 0; c < 4; c++) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 4; k++) acc += wv[r][k] * PT[k][c];
      o[16 + r * 4 + c] = acc;
    }
  // camera centre = -(A^-1) t with A = R^T (rows of Rt), by the adjugate
  const float a00 = Rt[0][0], a01 = Rt[0][1], a02 = Rt[0][2], a10 = Rt[1][0], a11 = Rt[1][1], a12 = Rt[1][2];
  const float a20 = Rt[2][0], a21 = Rt[2][1], a22 = Rt[2][2];
  const float c00 = a11 * a22 - a12 * a21, c01 = a02 * a21 - a01 * a22, c02 = a01 * a12 - a02 * a11;
  const float c10 = a12 * a20 - a10 * a22, c11 = a00 * a22 - a02 * a20, c12 = a02 * a10 - a00 * a12;
  const float c20 = a10 * a21 - a11 * a20, c21 = a01 * a20 - a00 * a21, c22 = a00 * a11 - a01 * a10;
  const float inv_det = 1.0f / (a00 * c00 + a01 * c10 + a02 * c20);
  const float tx = Rt[0][3], ty = Rt[1][3], tz = Rt[2][3];
  o[32] = -(c00 * tx + c01 * ty + c02 * tz) * inv_det;
  o[33] = -(c10 * tx + c11 * ty + c12 * tz) * inv_det;
  o[34] = -(c20 * tx + c21 * ty + c22 * tz) * inv_det;
}


// ---- multi-GPU: the gradient exchange over peer memory, fused with the optimiser step (SURVEY.md s.8 row e) --------
// Views are sharded over ranks; the per-Gaussian gradients [17P] (14P parameters | 3P viewspace) and the radii maxima [P]
// of every rank live in a symmetric allocation mapped into every peer (NVLink / NVSwitch). Instead of
// NCCL all-reduce(grad) + all-reduce(radii) + k_densify_stats + k_params_adam (four launches, two stream hops):
//   k_peer_allreduce      rank r sums slice r of the W gradient buffers straight out of the peers' memory (P2P loads in a
//                         fixed rank order, or ONE multimem.ld_reduce per 16 bytes when the allocation has an NVLS multicast
//                         address: the switch adds) and pushes the result into every rank's reduced buffer (P2P stores /
//                         multimem.st). Its entry barrier (flags[0]) is the one cross-rank synchronisation of the step.
//   k_params_adam_peers   waits (flags[1]) until every slice has landed, then densification statistics + Adam on the
//                         local replica. Every rank adds in the same order: replicas stay bit-identical.
// Flags are monotonically increasing epochs written with st.release.sys into the PEER's flag array and polled locally.
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void peer_signal(const GdPeerTable& t, int phase, unsigned epoch) {   // threads 0..world-1
  if ((int)threadIdx.x < t.world) {
    __threadfence_system();
    st_release_sys(t.flags[threadIdx.x] + phase * GD_MAX_PEERS + t.rank, epoch);
  }
}
__device__ __forceinline__ void peer_wait(const GdPeerTable& t, int phase, unsigned epoch) {
  if ((int)threadIdx.x < t.world) {
    const unsigned* f = t.flags[t.rank] + phase * GD_MAX_PEERS + threadIdx.x;
    while ((int)(ld_acquire_sys(f) - epoch) < 0) {}
  }
  __syncthreads();
}
__device__ __forceinline__ float4 ld_peer4(const float* p) {   // peer memory: bypass L1, nothing to reuse
  float4 v;
  asm volatile("ld.global.relaxed.sys.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int4 ld_peer4i(const int* p) {
  int4 v;
  asm volatile("ld.global.relaxed.sys.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__global__ void __launch_bounds__(256)
k_peer_allreduce(const GdPeerTable t, long long g4_per_rank, long long g4_total, long long r4_per_rank, long long r4_total,
                 unsigned epoch, unsigned* __restrict__ counter) {
  __shared__ bool s_last;
  if (blockIdx.x == 0) peer_signal(t, 0, epoch);   // stream order: this rank's backward has retired, its buffers are final
  peer_wait(t, 0, epoch);
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < g4_per_rank) {
    const long long e = (long long)t.rank * g4_per_rank + i;
    if (e < g4_total) {
      float4 acc;
      if (t.mc_grad) {   // NVLS: the switch reduces the W copies and broadcasts the sum
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(acc.x), "=f"(acc.y), "=f"(acc.z), "=f"(acc.w) : "l"(t.mc_grad + 4 * e) : "memory");
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(t.mc_red_grad + 4 * e), "f"(acc.x), "f"(acc.y),
                     "f"(acc.z), "f"(acc.w) : "memory");
      } else {
        float4 v[GD_MAX_PEERS];
#pragma unroll
        for (int w = 0; w < GD_MAX_PEERS; w++)
          if (w < t.world) v[w] = ld_peer4(t.grad[w] + 4 * e);
        acc = v[0];
#pragma unroll
        for (int w = 1; w < GD_MAX_PEERS; w++)
          if (w < t.world) { acc.x += v[w].x; acc.y += v[w].y; acc.z += v[w].z; acc.w += v[w].w; }
#pragma unroll
        for (int w = 0; w < GD_MAX_PEERS; w++)
          if (w < t.world) *reinterpret_cast<float4*>(t.red_grad[w] + 4 * e) = acc;
      }
    }
  } else if (i - g4_per_rank < r4_per_rank) {
    const long long e = (long long)t.rank * r4_per_rank + (i - g4_per_rank);
    if (e < r4_total) {
      int4 m = ld_peer4i(t.radii[0] + 4 * e);
      for (int w = 1; w < t.world; w++) {
        const int4 v = ld_peer4i(t.radii[w] + 4 * e);
        m.x = max(m.x, v.x); m.y = max(m.y, v.y); m.z = max(m.z, v.z); m.w = max(m.w, v.w);
      }
      for (int w = 0; w < t.world; w++) *reinterpret_cast<int4*>(t.red_radii[w] + 4 * e) = m;
    }
  }
  // the last CTA of this rank tells every peer that slice `rank` has landed in its reduced buffer
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(counter, 1u);
    s_last = prev == gridDim.x - 1;
    if (s_last) *counter = 0;
  }
  __syncthreads();
  if (s_last) peer_signal(t, 1, epoch);
}
__global__ void __launch_bounds__(256)
k_params_adam_peers(int P, float* __restrict__ xyz, float* __restrict__ f_dc, float* __restrict__ opacity,
                    float* __restrict__ scaling, float* __restrict__ rotation, const GdPeerTable t, unsigned epoch,
                    float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq, AdamHyper h, int densify,
                    float* __restrict__ xyz_gradient_accum, float* __restrict__ denom, float* __restrict__ max_radii2D) {
  peer_wait(t, 1, epoch);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float* grad = t.red_grad[t.rank];
  if (densify) {
    const int r = t.red_radii[t.rank][i];
    if (r > 0) {
      const float gx = grad[14 * (size_t)P + 3 * (size_t)i], gy = grad[14 * (size_t)P + 3 * (size_t)i + 1];
      xyz_gradient_accum[i] += sqrtf(gx * gx + gy * gy);
      denom[i] += 1.0f;
      max_radii2D[i] = fmaxf(max_radii2D[i], (float)r);
    }
  }
  adam_one(i, P, xyz, f_dc, opacity, scaling, rotation, grad, exp_avg, exp_avg_sq, h);
}

}  // namespace gd
