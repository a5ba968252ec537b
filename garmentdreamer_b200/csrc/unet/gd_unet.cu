// Host side of include/gd_unet.h + the small fused kernels around the tcgen05 GEMM.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "gd_gemm.cuh"
#include "gd_attn.cuh"
#include "gd_vae.cuh"

namespace {
thread_local char g_err[512] = {0};
std::atomic<uint64_t> g_launches{0};
std::atomic<uint64_t> g_pair_launches{0};   // GEMM launches that ran as CTA pairs (cta_group::2)

int fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return GD_UNET_ERR_CUDA;
  }
  return GD_UNET_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
int make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
             const cuuint32_t* box, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(GD_UNET_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides,
                        box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu box %u %u %u",
             (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
             (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], box[1], rank > 2 ? box[2] : 0);
    return GD_UNET_ERR_CUDA;
  }
  return GD_UNET_OK;
}
// Programmatic dependent launch: every kernel of this library triggers its dependents at entry
// (griddepcontrol.launch_dependents) and waits for its predecessors' memory before touching global
// data (griddepcontrol.wait), so launch latency and kernel prologues overlap the previous kernel.
template <typename... KArgs, typename... Args>
void launch_pdl_cluster(int cluster_x, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (cluster_x > 1) {   // thread-block cluster (CTA pair of the cta_group::2 GEMM)
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = cluster_x; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  }
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  launch_pdl_cluster(1, kernel, grid, block, smem, st, static_cast<Args&&>(args)...);
}
// ---- per-device scratch (split-K partials, GroupNorm partial statistics) ------------------------------
// One set per CUDA device, allocated on first use ON THAT DEVICE (or up front by gd_unet_init, which callers
// that capture CUDA graphs run before the capture: cudaMalloc is illegal inside one). The buffers are shared by
// every stream of the device: the library serves ONE stream per device at a time (documented in gd_unet.h).
struct DeviceScratch {
  float* splitk_ws = nullptr;        // [kSplitKFloats]
  float2* gn_part_small = nullptr;   // [kGnSmall]   gd_unet_groupnorm
  float2* gn_part = nullptr;         // [kGnBig]     gd_unet_groupnorm_stats / _bwd
  float2* gn_bstats = nullptr;       // [8192]
};
constexpr size_t kSplitKFloats = (size_t)24 << 20;   // 96 MB
constexpr size_t kGnSmall = (size_t)4096 * 64, kGnBig = (size_t)8192 * 512;
constexpr int kMaxDevices = 64;
DeviceScratch g_scratch[kMaxDevices];
DeviceScratch* device_scratch() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
  DeviceScratch& s = g_scratch[dev];
  if (!s.splitk_ws) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    (void)cap;
    if (cudaMalloc(&s.splitk_ws, kSplitKFloats * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&s.gn_part_small, kGnSmall * sizeof(float2)) != cudaSuccess ||
        cudaMalloc(&s.gn_part, kGnBig * sizeof(float2)) != cudaSuccess ||
        cudaMalloc(&s.gn_bstats, 8192 * sizeof(float2)) != cudaSuccess) {
      cudaGetLastError();
      s = DeviceScratch();
      return nullptr;
    }
  }
  return &s;
}
#define LAUNCH_CHECK(what)              \
  do {                                  \
    const int rc_ = check_launch(what); \
    if (rc_ != GD_UNET_OK) return rc_;  \
  } while (0)
}  // namespace

namespace gdu {

// ---- GroupNorm (+SiLU), NHWC fp16 -------------------------------------------------------------
// Pass 1: grid (image, splits): every CTA sweeps ALL channels of its pixel slice with 16-byte
// loads (thread -> fixed 8-channel chunk(s), strided over pixels), stores per-channel partial sums
// in shared memory and folds them per group in a fixed order (deterministic) into
// part[(n*groups+g)*splits + s].
__global__ void __launch_bounds__(256)
k_gn_stats(const __half* __restrict__ x, float2* __restrict__ part, int HW, int C, int groups, int splits) {
  pdl_entry();
  __shared__ float s_c[2][2560];   // [sum | sumsq][pix_par * C]   (pix_par * C <= 2560)
  const int n = blockIdx.x, sp = blockIdx.y;
  const int cpg = C / groups, C8 = C >> 3;
  const int p0 = (int)((long long)HW * sp / splits), p1 = (int)((long long)HW * (sp + 1) / splits);
  const int pix_par = C8 <= 256 ? 256 / C8 : 1;        // pixels swept in parallel
  const int iters = C8 <= 256 ? 1 : (C8 + 255) / 256;  // chunks per thread when a pixel is wider than the CTA
  const uint4* xb = reinterpret_cast<const uint4*>(x + (size_t)n * HW * C);
  for (int it = 0; it < iters; it++) {
    const int chunk = C8 <= 256 ? (int)(threadIdx.x % C8) : (int)threadIdx.x + 256 * it;
    const int pl = C8 <= 256 ? (int)(threadIdx.x / C8) : 0;
    if (pl >= pix_par || chunk >= C8) continue;
    float s[8], ss[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { s[k] = 0.f; ss[k] = 0.f; }
#pragma unroll 4
    for (int pix = p0 + pl; pix < p1; pix += pix_par) {
      uint4 v = xb[(size_t)pix * C8 + chunk];
      const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const float2 f = __half22float2(h[k]);
        s[2 * k] += f.x; ss[2 * k] += f.x * f.x;
        s[2 * k + 1] += f.y; ss[2 * k + 1] += f.y * f.y;
      }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
      s_c[0][pl * C + chunk * 8 + k] = s[k];
      s_c[1][pl * C + chunk * 8 + k] = ss[k];
    }
  }
  __syncthreads();
  if (threadIdx.x < groups) {
    const int g = threadIdx.x;
    float a = 0.f, b = 0.f;
    for (int pl = 0; pl < pix_par; pl++)
      for (int c = g * cpg; c < (g + 1) * cpg; c++) { a += s_c[0][pl * C + c]; b += s_c[1][pl * C + c]; }
    part[((size_t)n * groups + g) * splits + sp] = make_float2(a, b);
  }
}
// Pass 2: grid (image, pixel chunks): per-channel scale/shift table in shared memory, then a
// vectorised y = x*a + b (+SiLU) sweep over the chunk (8 channels per 16-byte access).
__global__ void __launch_bounds__(256)
k_gn_apply(const __half* __restrict__ x, __half* __restrict__ y, const float2* __restrict__ part,
           const __half* __restrict__ gamma, const __half* __restrict__ beta, int HW, int C, int groups, int splits,
           float eps, int do_silu, int pix_per_cta) {
  pdl_entry();
  extern __shared__ float2 s_ab[];  // [C] (scale, shift)
  __shared__ float2 s_mr[256];      // per group (mean, rstd)
  const int n = blockIdx.x, cpg = C / groups;
  if (threadIdx.x < groups) {
    float s = 0.f, ss = 0.f;
    for (int k = 0; k < splits; k++) {
      const float2 p = part[((size_t)n * groups + threadIdx.x) * splits + k];
      s += p.x; ss += p.y;
    }
    const float inv_n = 1.0f / (float)(HW * cpg);
    const float mean = s * inv_n;
    s_mr[threadIdx.x] = make_float2(mean, rsqrtf(fmaxf(ss * inv_n - mean * mean, 0.0f) + eps));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float2 mr = s_mr[c / cpg];
    const float a = mr.y * __half2float(gamma[c]);
    s_ab[c] = make_float2(a, __half2float(beta[c]) - mr.x * a);
  }
  __syncthreads();
  const int C8 = C >> 3;
  const int p0 = blockIdx.y * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
  const uint4* xb = reinterpret_cast<const uint4*>(x + (size_t)n * HW * C);
  uint4* yb = reinterpret_cast<uint4*>(y + (size_t)n * HW * C);
  for (int i = p0 * C8 + threadIdx.x; i < p1 * C8; i += blockDim.x) {
    const int c0 = (i % C8) << 3;
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

// ---- GroupNorm (+SiLU) in ONE launch for L2-resident tensors: a thread-block cluster per image ----
// Each CTA of the cluster sweeps a pixel slice for the statistics (same sweep as k_gn_stats), the
// per-group partial sums are exchanged through distributed shared memory (mapa + ld.shared::cluster,
// summed in rank order => deterministic), then every CTA normalises its own slice (second read hits
// L2). Replaces k_gn_stats + k_gn_apply (two launches, partials through global memory).
__global__ void __launch_bounds__(256)
k_gn_cluster(const __half* __restrict__ x, __half* __restrict__ y, const __half* __restrict__ gamma,
             const __half* __restrict__ beta, int HW, int C, int groups, float eps, int do_silu) {
  pdl_entry();
  __shared__ float s_c[2][2560];
  __shared__ float2 s_grp[256];    // this CTA's partial (sum, sumsq) per group -- read by the peers
  __shared__ float2 s_mr[256];     // (mean, rstd) per group
  extern __shared__ float2 s_ab[]; // [C] (scale, shift)
  uint32_t rank, cs;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(cs));
  const int n = blockIdx.x / cs;
  const int cpg = C / groups, C8 = C >> 3;
  const int p0 = (int)((long long)HW * rank / cs), p1 = (int)((long long)HW * (rank + 1) / cs);
  const int pix_par = C8 <= 256 ? 256 / C8 : 1;
  const int iters = C8 <= 256 ? 1 : (C8 + 255) / 256;
  const uint4* xb = reinterpret_cast<const uint4*>(x + (size_t)n * HW * C);
  for (int it = 0; it < iters; it++) {
    const int chunk = C8 <= 256 ? (int)(threadIdx.x % C8) : (int)threadIdx.x + 256 * it;
    const int pl = C8 <= 256 ? (int)(threadIdx.x / C8) : 0;
    if (pl >= pix_par || chunk >= C8) continue;
    float s[8], ss[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { s[k] = 0.f; ss[k] = 0.f; }
#pragma unroll 4
    for (int pix = p0 + pl; pix < p1; pix += pix_par) {
      uint4 v = xb[(size_t)pix * C8 + chunk];
      const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const float2 f = __half22float2(h[k]);
        s[2 * k] += f.x; ss[2 * k] += f.x * f.x;
        s[2 * k + 1] += f.y; ss[2 * k + 1] += f.y * f.y;
      }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
      s_c[0][pl * C + chunk * 8 + k] = s[k];
      s_c[1][pl * C + chunk * 8 + k] = ss[k];
    }
  }
  __syncthreads();
  if (threadIdx.x < groups) {
    const int g = threadIdx.x;
    float a = 0.f, b = 0.f;
    for (int pl = 0; pl < pix_par; pl++)
      for (int c = g * cpg; c < (g + 1) * cpg; c++) { a += s_c[0][pl * C + c]; b += s_c[1][pl * C + c]; }
    s_grp[g] = make_float2(a, b);
  }
  cluster_sync_all();   // every CTA's partials are visible cluster-wide
  if (threadIdx.x < groups) {
    const uint32_t local = s2u(&s_grp[threadIdx.x]);
    float s = 0.f, ss = 0.f;
    for (uint32_t k = 0; k < cs; k++) {   // rank order: deterministic
      uint32_t remote;
      float2 pv;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(k));
      asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(pv.x), "=f"(pv.y) : "r"(remote));
      s += pv.x; ss += pv.y;
    }
    const float inv_n = 1.0f / (float)(HW * cpg);
    const float mean = s * inv_n;
    s_mr[threadIdx.x] = make_float2(mean, rsqrtf(fmaxf(ss * inv_n - mean * mean, 0.0f) + eps));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float2 mr = s_mr[c / cpg];
    const float a = mr.y * __half2float(gamma[c]);
    s_ab[c] = make_float2(a, __half2float(beta[c]) - mr.x * a);
  }
  __syncthreads();
  uint4* yb = reinterpret_cast<uint4*>(y + (size_t)n * HW * C);
  for (int i = p0 * C8 + threadIdx.x; i < p1 * C8; i += blockDim.x) {
    const int c0 = (i % C8) << 3;
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
  cluster_sync_all();   // peers may still be reading this CTA's partials
}

// ---- GroupNorm statistics from the column statistics of the producing GEMM(s) ---------------------------------------
// colstats: [rowblocks][2][Cs] fp32 (sum | sum of squares per 32-row block and channel, written by k_gemm_tcgen05's epilogue):
// 1/8 of the bytes of the fp16 tensor they describe. grid (image, splits): a CTA folds the row blocks of its slice of the
// image with coalesced 16-byte loads (thread -> fixed column quad, row blocks strided over `lanes` thread groups), combines
// the lanes and then the channels of a group in a fixed order (deterministic) and writes part[(n*groups+g)*splits + sp] --
// the layout k_gn_stats produces, so the apply / finalize kernels are shared. The C channels are source A's Ca followed by
// source B's Cb (torch.cat of the UNet up path).
__device__ __forceinline__ void colstats_fold_source(const float* __restrict__ st, int Cs, int r0, int r1, float* s_lane,
                                                     float* s_tot, int c_off, int C) {
  const int Q = Cs >> 1;                                  // float4 per row block ([2][Cs] floats)
  const int lanes = Q <= 256 ? 256 / Q : 1, iters = Q <= 256 ? 1 : (Q + 255) / 256;
  for (int it = 0; it < iters; it++) {
    const int q = Q <= 256 ? (int)(threadIdx.x % Q) : (int)threadIdx.x + 256 * it;
    const int l = Q <= 256 ? (int)(threadIdx.x / Q) : 0;
    if (l < lanes && q < Q) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* src = reinterpret_cast<const float4*>(st) + q;
#pragma unroll 4
      for (int rb = r0 + l; rb < r1; rb += lanes) {
        const float4 v = src[(size_t)rb * Q];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      if (Q <= 256) reinterpret_cast<float4*>(s_lane)[l * Q + q] = acc;
      else {   // one lane: straight into the totals ([sum | sumsq][C])
        const int f = q * 4, half = f >= Cs ? 1 : 0, c = f - half * Cs;
        float* d = s_tot + half * C + c_off + c;
        d[0] = acc.x; d[1] = acc.y; d[2] = acc.z; d[3] = acc.w;
      }
    }
  }
  if (Q <= 256) {
    __syncthreads();
    for (int f = threadIdx.x; f < 2 * Cs; f += 256) {
      float a = 0.f;
      for (int l = 0; l < lanes; l++) a += s_lane[l * 2 * Cs + f];
      const int half = f >= Cs ? 1 : 0;
      s_tot[half * C + c_off + (f - half * Cs)] = a;
    }
  }
  __syncthreads();
}
__global__ void __launch_bounds__(256)
k_gn_colstats_reduce(const float* __restrict__ statsA, int Ca, const float* __restrict__ statsB, int Cb, int rb_per_image,
                     int groups, int splits, float2* __restrict__ part, const __half* __restrict__ gamma) {
  pdl_entry();
  __shared__ __align__(16) float s_lane[1024];   // lanes * 2 * Cs <= 1024 floats when Cs <= 512
  __shared__ float s_tot[2 * 2560];              // [sum | sumsq][C]
  const int n = blockIdx.x, sp = blockIdx.y, C = Ca + Cb, cpg = C / groups;
  const int r0 = n * rb_per_image + (int)((long long)rb_per_image * sp / splits);
  const int r1 = n * rb_per_image + (int)((long long)rb_per_image * (sp + 1) / splits);
  colstats_fold_source(statsA, Ca, r0, r1, s_lane, s_tot, 0, C);
  if (Cb > 0) colstats_fold_source(statsB, Cb, r0, r1, s_lane, s_tot, Ca, C);
  if ((int)threadIdx.x < groups) {
    const int g = threadIdx.x;
    float a = 0.f, b = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; c++) {
      const float w = gamma ? __half2float(gamma[c]) : 1.0f;   // backward: sum_c gamma_c * (sum g | sum g xh)
      a = fmaf(s_tot[c], w, a); b = fmaf(s_tot[C + c], w, b);
    }
    part[((size_t)n * groups + g) * splits + sp] = make_float2(a, b);
  }
}
// GroupNorm-backward coefficient table for the MODE 4 GEMM epilogue: per (image, channel) (ya, yb, ca, cb) with
// xh = x*ca + cb = (x - mean) * rstd and y = x*ya + yb = xh*gamma + beta.
__global__ void __launch_bounds__(256)
k_gn_bwd_coef(const float2* __restrict__ stats, const __half* __restrict__ gamma, const __half* __restrict__ beta,
              float4* __restrict__ coef, int N, int C, int groups) {
  pdl_entry();
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= N * C) return;
  const int n = i / C, c = i % C;
  const float2 mr = stats[(size_t)n * groups + c / (C / groups)];
  const float ca = mr.y, cb = -mr.x * mr.y, g = __half2float(gamma[c]);
  coef[i] = make_float4(ca * g, fmaf(cb, g, __half2float(beta[c])), ca, cb);
}

// ---- LayerNorm: one warp per row -----------------------------------------------------------
__global__ void __launch_bounds__(256)
k_layernorm(const __half* __restrict__ x, __half* __restrict__ y, const __half* __restrict__ gamma,
            const __half* __restrict__ beta, int rows, int C, float eps) {
  pdl_entry();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  if ((C & 7) == 0 && C <= 1280) {
    // the row lives in registers between the two passes: up to five 16-byte loads per lane, all issued before the first use
    const int n8 = C >> 3;
    const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)row * C);
    uint4 v[5];
#pragma unroll
    for (int k = 0; k < 5; k++) v[k] = (lane + 32 * k < n8) ? xr[lane + 32 * k] : make_uint4(0, 0, 0, 0);
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int k = 0; k < 5; k++) {
      const __half2* h = reinterpret_cast<const __half2*>(&v[k]);
#pragma unroll
      for (int t = 0; t < 4; t++) {
        const float2 f = __half22float2(h[t]);
        s += f.x + f.y; ss += f.x * f.x + f.y * f.y;
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(~0u, s, o); ss += __shfl_xor_sync(~0u, ss, o); }
    const float mean = s / C, rstd = rsqrtf(fmaxf(ss / C - mean * mean, 0.f) + eps);
    const uint4* g4 = reinterpret_cast<const uint4*>(gamma);
    const uint4* b4 = reinterpret_cast<const uint4*>(beta);
    uint4* yr = reinterpret_cast<uint4*>(y + (size_t)row * C);
#pragma unroll
    for (int k = 0; k < 5; k++) {
      const int c8 = lane + 32 * k;
      if (c8 < n8) {
        const uint4 gv = g4[c8], bv = b4[c8];
        const __half2* h = reinterpret_cast<const __half2*>(&v[k]);
        const __half2* gh = reinterpret_cast<const __half2*>(&gv);
        const __half2* bh = reinterpret_cast<const __half2*>(&bv);
        uint4 o;
        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int t = 0; t < 4; t++) {
          const float2 f = __half22float2(h[t]), ga = __half22float2(gh[t]), be = __half22float2(bh[t]);
          oh[t] = __floats2half2_rn((f.x - mean) * rstd * ga.x + be.x, (f.y - mean) * rstd * ga.y + be.y);
        }
        yr[c8] = o;
      }
    }
    return;
  }
  const __half2* xr = reinterpret_cast<const __half2*>(x + (size_t)row * C);
  __half2* yr = reinterpret_cast<__half2*>(y + (size_t)row * C);
  float s = 0.f, ss = 0.f;
  for (int c = lane; c < (C >> 1); c += 32) {
    const float2 v = __half22float2(xr[c]);
    s += v.x + v.y; ss += v.x * v.x + v.y * v.y;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(~0u, s, o); ss += __shfl_xor_sync(~0u, ss, o); }
  const float mean = s / C, rstd = rsqrtf(fmaxf(ss / C - mean * mean, 0.f) + eps);
  const __half2* g2 = reinterpret_cast<const __half2*>(gamma);
  const __half2* b2 = reinterpret_cast<const __half2*>(beta);
  for (int c = lane; c < (C >> 1); c += 32) {
    const float2 v = __half22float2(xr[c]), ga = __half22float2(g2[c]), be = __half22float2(b2[c]);
    yr[c] = __floats2half2_rn((v.x - mean) * rstd * ga.x + be.x, (v.y - mean) * rstd * ga.y + be.y);
  }
}

// ---- row softmax in place: one warp per row (cols <= 4096) ---------------------------------
__global__ void __launch_bounds__(256)
k_softmax(__half* __restrict__ s, long long rows, int cols, long long ld) {
  pdl_entry();
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  __half* r = s + row * ld;
  float mx = -INFINITY;
  for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, __half2float(r[c]));
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(~0u, mx, o));
  float sum = 0.f;
  for (int c = lane; c < cols; c += 32) sum += __expf(__half2float(r[c]) - mx);
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(~0u, sum, o);
  const float inv = 1.0f / sum;
  for (int c = lane; c < cols; c += 32) r[c] = __float2half_rn(__expf(__half2float(r[c]) - mx) * inv);
}

__global__ void k_geglu(const __half* __restrict__ x, __half* __restrict__ y, long long rows, int H) {
  pdl_entry();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * H) return;
  const long long r = i / H;
  const int j = (int)(i % H);
  const float a = __half2float(x[r * 2 * H + j]), g = __half2float(x[r * 2 * H + H + j]);
  y[i] = __float2half_rn(a * gelu_erf(g));
}
__global__ void k_add(const __half2* __restrict__ a, const __half2* __restrict__ b, __half2* __restrict__ y, long long n2) {
  pdl_entry();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n2) y[i] = __hadd2(a[i], b[i]);
}
__global__ void k_upsample2x(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H, int W, int C8) {
  pdl_entry();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)N * 2 * H * 2 * W * C8;
  if (i >= total) return;
  const int c = (int)(i % C8);
  long long p = i / C8;
  const int ox = (int)(p % (2 * W)); p /= 2 * W;
  const int oy = (int)(p % (2 * H));
  const int n = (int)(p / (2 * H));
  y[i] = x[(((long long)n * H + (oy >> 1)) * W + (ox >> 1)) * C8 + c];
}
__global__ void k_space_to_depth(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H, int W, int C8) {
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
  y[((((long long)n * (H >> 1) + (iy >> 1)) * (W >> 1) + (ix >> 1)) * 4 + ph) * C8 + c] = x[i];
}
__global__ void k_concat(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ y, long long rows,
                         int Ca8, int Cb8) {
  pdl_entry();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int Ct = Ca8 + Cb8;
  if (i >= rows * Ct) return;
  const long long r = i / Ct;
  const int c = (int)(i % Ct);
  y[i] = c < Ca8 ? a[r * Ca8 + c] : b[r * Cb8 + (c - Ca8)];
}
// y[b,n] = act_out(bias[n] + sum_k act_in(x[b,k]) W[n,k]); one warp per output column n.
// The activations (Bm x K fp16, <= 40 KB) are staged in shared memory once per CTA; every lane
// then issues all of its 16-byte weight loads before using them (the kernel is DRAM-latency bound).
template <int KV>  // KV = 16-byte weight vectors per lane (K = KV * 256)
__global__ void __launch_bounds__(256)
k_small_linear(const __half* __restrict__ x, const __half* __restrict__ W, const __half* __restrict__ bias,
               __half* __restrict__ y, int Bm, int K, int N, int silu_in, int silu_out) {
  pdl_entry();
  extern __shared__ __half s_x[];  // [Bm][K]
  for (int i = threadIdx.x; i < Bm * K; i += blockDim.x) {
    float v = __half2float(x[i]);
    if (silu_in) v = silu(v);
    s_x[i] = __float2half_rn(v);
  }
  __syncthreads();
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  uint4 wv[KV];
  const uint4* w4 = reinterpret_cast<const uint4*>(W + (size_t)n * K);
#pragma unroll
  for (int j = 0; j < KV; j++) wv[j] = (j * 32 + lane) * 8 < K ? w4[j * 32 + lane] : make_uint4(0, 0, 0, 0);
  float acc[16];
#pragma unroll
  for (int b = 0; b < 16; b++) acc[b] = 0.f;
#pragma unroll
  for (int j = 0; j < KV; j++) {
    const int k0 = (j * 32 + lane) * 8;
    if (k0 >= K) continue;
    const __half2* wh = reinterpret_cast<const __half2*>(&wv[j]);
#pragma unroll
    for (int b = 0; b < 16; b++) {
      if (b < Bm) {
        const uint4 xv = *reinterpret_cast<const uint4*>(s_x + (size_t)b * K + k0);
        const __half2* xh = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
        for (int t = 0; t < 4; t++) {
          const float2 a = __half22float2(xh[t]), w = __half22float2(wh[t]);
          acc[b] += a.x * w.x + a.y * w.y;
        }
      }
    }
  }
#pragma unroll
  for (int b = 0; b < 16; b++) {
#pragma unroll
    for (int o = 16; o; o >>= 1) acc[b] += __shfl_xor_sync(~0u, acc[b], o);
  }
  if (lane == 0) {
    const float bv = bias ? __half2float(bias[n]) : 0.f;
    for (int b = 0; b < Bm; b++) {
      float v = acc[b] + bv;
      if (silu_out) v = silu(v);
      y[(size_t)b * N + n] = __float2half_rn(v);
    }
  }
}
__global__ void k_timestep_embedding(const float* __restrict__ t, __half* __restrict__ y, int Bm, int dim) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim >> 1;
  if (i >= Bm * half) return;
  const int b = i / half, k = i % half;
  const float tt = __half2float(__float2half_rn(t[b]));  // reference casts t to fp16 first
  const float freq = expf(-logf(10000.0f) * (float)k / (float)half);
  const float a = tt * freq;
  y[(size_t)b * dim + k] = __float2half_rn(cosf(a));         // flip_sin_to_cos: [cos | sin]
  y[(size_t)b * dim + half + k] = __float2half_rn(sinf(a));
}
// conv_in: Cin = 4, NCHW fp16 in -> NHWC fp16 out. Weights staged in shared memory; one thread
// per pixel keeps its 3x3x4 input patch in registers and walks the output channels.
__global__ void __launch_bounds__(128)
k_conv_in(const __half* __restrict__ x, const __half* __restrict__ w, const __half* __restrict__ bias,
          __half* __restrict__ y, int N, int H, int W, int Cout) {
  pdl_entry();
  extern __shared__ __half s_w[];  // [Cout][36] + [Cout] bias
  for (int i = threadIdx.x; i < Cout * 36; i += blockDim.x) s_w[i] = w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) s_w[Cout * 36 + i] = bias[i];
  __syncthreads();
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)N * H * W) return;
  const int px = (int)(pix % W), py = (int)((pix / W) % H), n = (int)(pix / ((long long)W * H));
  float in[36];
#pragma unroll
  for (int ky = 0; ky < 3; ky++)
#pragma unroll
    for (int kx = 0; kx < 3; kx++) {
      const int iy = py + ky - 1, ix = px + kx - 1;
      const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
#pragma unroll
      for (int c = 0; c < 4; c++)
        in[(ky * 3 + kx) * 4 + c] = ok ? __half2float(x[(((size_t)n * 4 + c) * H + iy) * W + ix]) : 0.f;
    }
  __half* yp = y + (size_t)pix * Cout;
  for (int co = 0; co < Cout; co += 8) {
    __align__(16) __half o[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const __half2* wp = reinterpret_cast<const __half2*>(s_w + (co + j) * 36);
      float acc = __half2float(s_w[Cout * 36 + co + j]);
#pragma unroll
      for (int k = 0; k < 18; k++) {
        const float2 ww = __half22float2(wp[k]);
        acc += in[2 * k] * ww.x + in[2 * k + 1] * ww.y;
      }
      o[j] = __float2half_rn(acc);
    }
    *reinterpret_cast<uint4*>(yp + co) = *reinterpret_cast<const uint4*>(o);
  }
}
// conv_in on the tensor cores: im2col rows A[(n,y,x)][(ky*3+kx)*4 + c] (36 of 64 columns, the rest zero) from the fp16 NCHW
// latents; the 3x3x4 -> Cout convolution is then one K = 64 GEMM (k_gemm_tcgen05) whose epilogue also leaves the GroupNorm
// statistics of the first resnet. One thread per (pixel, 16-byte chunk of its 128-byte row).
__global__ void __launch_bounds__(256)
k_unet_im2col4(const __half* __restrict__ x, __half* __restrict__ A, int N, int H, int W) {
  pdl_entry();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long HW = (long long)H * W, i = t >> 3;
  const int u = (int)(t & 7);
  if (i >= (long long)N * HW) return;
  const int px = (int)(i % W), py = (int)((i / W) % H), n = (int)(i / HW);
  __align__(16) __half v[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int k = u * 8 + j, tap = k >> 2, c = k & 3;
    const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
    const bool ok = k < 36 && yy >= 0 && yy < H && xx >= 0 && xx < W;
    v[j] = ok ? x[((long long)n * 4 + c) * HW + (long long)yy * W + xx] : __float2half_rn(0.f);
  }
  reinterpret_cast<uint4*>(A)[t] = *reinterpret_cast<const uint4*>(v);
}
// conv_out tail: the first 4 of `ld` fp16 channels per pixel (NHWC) -> NCHW fp32
__global__ void __launch_bounds__(256)
k_unpack4_nchw(const __half* __restrict__ x, float* __restrict__ y, int N, long long HW, int ld) {
  pdl_entry();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * HW) return;
  const long long n = i / HW, p = i % HW;
  const uint2 v = *reinterpret_cast<const uint2*>(x + i * ld);
  const __half2* h = reinterpret_cast<const __half2*>(&v);
  const float2 a = __half22float2(h[0]), b = __half22float2(h[1]);
  float* yo = y + n * 4 * HW + p;
  yo[0] = a.x; yo[HW] = a.y; yo[2 * HW] = b.x; yo[3 * HW] = b.y;
}
// conv_out: Cout = 4, NHWC fp16 in -> NCHW fp32 out. One warp per pixel.
__global__ void __launch_bounds__(256)
k_conv_out(const __half* __restrict__ x, const __half* __restrict__ w, const __half* __restrict__ bias,
           float* __restrict__ y, int N, int H, int W, int Cin) {
  pdl_entry();
  const long long pix = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pix >= (long long)N * H * W) return;
  const int px = (int)(pix % W), py = (int)((pix / W) % H), n = (int)(pix / ((long long)W * H));
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int ky = 0; ky < 3; ky++) {
    const int iy = py + ky - 1;
    if (iy < 0 || iy >= H) continue;
    for (int kx = 0; kx < 3; kx++) {
      const int ix = px + kx - 1;
      if (ix < 0 || ix >= W) continue;
      const __half2* xp = reinterpret_cast<const __half2*>(x + (((size_t)n * H + iy) * W + ix) * Cin);
      for (int c = lane; c < (Cin >> 1); c += 32) {
        const float2 v = __half22float2(xp[c]);
#pragma unroll
        for (int o = 0; o < 4; o++) {
          const float2 ww = __half22float2(reinterpret_cast<const __half2*>(w + (((size_t)o * 3 + ky) * 3 + kx) * Cin)[c]);
          acc[o] += v.x * ww.x + v.y * ww.y;
        }
      }
    }
  }
#pragma unroll
  for (int o = 0; o < 4; o++) {
#pragma unroll
    for (int s = 16; s; s >>= 1) acc[o] += __shfl_xor_sync(~0u, acc[o], s);
  }
  if (lane < 4) {
    float v = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
    // the reference UNet runs in fp16: its conv_out result is rounded to fp16 before .to(fp32)
    v = __half2float(__float2half_rn(v + __half2float(bias[lane])));
    y[(((size_t)n * 4 + lane) * H + py) * W + px] = v;
  }
}
__global__ void k_add_noise(const float* __restrict__ lat, const float* __restrict__ noise, const float* __restrict__ sa,
                            const float* __restrict__ sb, float* __restrict__ noisy, __half* __restrict__ uin, int B,
                            int reps, int chw) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * chw) return;
  const int b = i / chw;
  const float v = sa[b] * lat[i] + sb[b] * noise[i];
  noisy[i] = v;
  const __half h = __float2half_rn(v);
  for (int r = 0; r < reps; r++) uin[(size_t)r * B * chw + i] = h;
}
__global__ void k_sds_grad(const float* __restrict__ eps, const float* __restrict__ noise, const float* __restrict__ w,
                           float s, float* __restrict__ noise_pred, float* __restrict__ grad, int B, int chw) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * chw) return;
  const int b = i / chw;
  const float et = eps[i], eu = eps[(size_t)B * chw + i];
  const float np = et + s * (et - eu);  // sic: base term is e_text (stable_diffusion_guidance.py:248-251)
  if (noise_pred) noise_pred[i] = np;
  grad[i] = w[b] * (np - noise[i]);
}

// Split-K finalize: y = fp16(alpha * sum_ks ws[ks] + bias + row_bias + residual) (+SiLU).
__global__ void __launch_bounds__(256)
k_splitk_finalize(const float* __restrict__ ws, int ksplit, int M, int N, float alpha, const __half* __restrict__ bias,
                  const __half* __restrict__ row_bias, long long row_bias_ld, int rows_per_image, const __half* __restrict__ residual,
                  __half* __restrict__ C, long long ldc, unsigned flags) {
  pdl_entry();
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= (long long)M * N) return;
  const int row = (int)(i / N), n = (int)(i % N);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = 0; k < ksplit; k++) {
    const float4 v = *reinterpret_cast<const float4*>(ws + (size_t)k * M * N + i);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  float o[4] = {acc.x * alpha, acc.y * alpha, acc.z * alpha, acc.w * alpha};
#pragma unroll
  for (int j = 0; j < 4; j++) {
    if (bias) o[j] += __half2float(bias[n + j]);
    if (row_bias) o[j] += __half2float(row_bias[(long long)(row / rows_per_image) * row_bias_ld + n + j]);
    if (residual) o[j] += __half2float(residual[(long long)row * ldc + n + j]);
    if (flags & GD_EPI_SILU) o[j] = silu(o[j]);
  }
  __half2* dst = reinterpret_cast<__half2*>(C + (long long)row * ldc + n);
  dst[0] = __floats2half2_rn(o[0], o[1]);
  dst[1] = __floats2half2_rn(o[2], o[3]);
}

// Linear stand-in for the VAE encoder (the VAE is outside this build):
// latents[b,k,y,x] = sum_c mix[k][c] * mean_{8x8}(2*color[b,c]-1); and its exact transpose.
__global__ void k_pool_latents(const float* __restrict__ color, const float* __restrict__ mix, float* __restrict__ lat,
                               int B, int H, int W) {
  pdl_entry();
  const int Ho = H >> 3, Wo = W >> 3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Ho * Wo) return;
  const int ox = i % Wo, oy = (i / Wo) % Ho, b = i / (Wo * Ho);
  float m[3];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    float s = 0.f;
    const float* src = color + (((size_t)b * 3 + c) * H + oy * 8) * W + ox * 8;
    for (int dy = 0; dy < 8; dy++) {
      const float4 u = *reinterpret_cast<const float4*>(src + (size_t)dy * W);
      const float4 v = *reinterpret_cast<const float4*>(src + (size_t)dy * W + 4);
      s += u.x + u.y + u.z + u.w + v.x + v.y + v.z + v.w;
    }
    m[c] = 2.0f * (s * (1.0f / 64.0f)) - 1.0f;
  }
#pragma unroll
  for (int k = 0; k < 4; k++)
    lat[(((size_t)b * 4 + k) * Ho + oy) * Wo + ox] = mix[k * 3] * m[0] + mix[k * 3 + 1] * m[1] + mix[k * 3 + 2] * m[2];
}
// dL/dcolor from dL/dlatents (nan_to_num + clamp first, stable_diffusion_guidance.py:418-421),
// scaled by 1/B like loss_sds (:427).
__global__ void k_pool_latents_bwd(const float* __restrict__ grad, const float* __restrict__ mix, float* __restrict__ dcolor,
                                   int B, int H, int W, float clip, float scale) {
  pdl_entry();
  const int Ho = H >> 3, Wo = W >> 3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * 3 * H * W) return;
  const int x = (int)(i % W), y = (int)((i / W) % H), c = (int)((i / ((long long)W * H)) % 3), b = (int)(i / ((long long)3 * W * H));
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    float g = grad[(((size_t)b * 4 + k) * Ho + (y >> 3)) * Wo + (x >> 3)];
    if (!(g == g)) g = 0.f;                          // nan_to_num
    g = fminf(fmaxf(g, -3.4028235e38f), 3.4028235e38f);
    if (clip > 0.f) g = fminf(fmaxf(g, -clip), clip);
    acc += mix[k * 3 + c] * g;
  }
  dcolor[i] = acc * (2.0f / 64.0f) * scale;
}

}  // namespace gdu

extern "C" {

const char* gd_unet_last_error(void) { return g_err; }
uint64_t gd_unet_launch_count(void) { return g_launches.load(); }
uint64_t gd_unet_pair_launch_count(void) { return g_pair_launches.load(); }
const char* gd_unet_version(void) { return "gd_unet 0.2 (sm_100a, tcgen05+TMA)"; }
int gd_unet_init(void) {   // allocate the current device's scratch now (call before capturing a CUDA graph)
  return device_scratch() ? GD_UNET_OK : fail(GD_UNET_ERR_CUDA, "init: per-device scratch allocation failed");
}

int gd_unet_gemm(const GdGemmArgs* a, gd_ustream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !a->A || !a->B || !a->C) return fail(GD_UNET_ERR_INVALID_ARG, "gemm: null pointer");
  if (a->M <= 0 || a->N <= 0 || a->K <= 0 || a->K % gdu::kBK) return fail(GD_UNET_ERR_INVALID_ARG, "gemm: K must be a positive multiple of 64");
  if (a->batch < 1 || a->heads < 1 || a->batch % a->heads) return fail(GD_UNET_ERR_INVALID_ARG, "gemm: bad batch/heads");
  if (a->a_box[0] != gdu::kBK || a->a_box[1] * a->a_box[2] * a->a_box[3] != gdu::kBM)
    return fail(GD_UNET_ERR_INVALID_ARG, "gemm: A box must be 64 x (128 rows)");
  const int ntaps = a->ntaps > 0 ? a->ntaps : 1;
  if (ntaps > 9 || (a->ntaps > 0 && (a->Ck <= 0 || a->Ck % gdu::kBK || a->K != ntaps * a->Ck)))
    return fail(GD_UNET_ERR_INVALID_ARG, "gemm: bad tap description");
  if ((a->flags & GD_EPI_GEGLU) && (a->N % 32 || (a->flags & GD_EPI_TRANSPOSED) || a->residual))
    return fail(GD_UNET_ERR_INVALID_ARG, "gemm: GEGLU epilogue needs N % 32 == 0 and no residual/transposition");
  int BN = a->block_n;
  if (BN <= 0) {
    // persistent grid of 148 CTAs: trade wave quantisation against tile width (arithmetic intensity)
    const int geglu = (a->flags & GD_EPI_GEGLU) ? 1 : 0;
    const long long mt = ((long long)a->M + gdu::kBM - 1) / gdu::kBM * a->batch;
    const bool splitk_ok = !(a->flags & (GD_EPI_GEGLU | GD_EPI_TRANSPOSED)) && a->batch == 1 && a->N % 4 == 0 &&
                           a->c_batch_stride == 0 && a->c_head_stride == 0 && a->K / gdu::kBK >= 24 && a->c_up2_w == 0;
    if (a->N <= 64) BN = (a->N + 15) / 16 * 16;
    else if (splitk_ok && mt * ((a->N + 255) / 256) <= 37) {
      // few output tiles, long K: wide tiles + split-K (below) instead of narrow tiles
      BN = a->N % 256 == 0 ? 256 : a->N % 160 == 0 ? 160 : 128;
      if (BN > (a->N + 15) / 16 * 16) BN = (a->N + 15) / 16 * 16;
    } else {
      const int cands[] = {256, 192, 160, 128, 96, 80, 64};
      double best = -1.0;
      BN = 64;
      for (int c : cands) {
        if ((geglu || a->colstats) && c % 32) continue;   // whole 32-column epilogue chunks
        if (c > ((a->N + 15) / 16 * 16)) continue;
        if (a->N % c && !(c == 128 || c == 64)) continue;   // ragged N only with 128 / 64
        const long long tiles = mt * ((a->N + c - 1) / c);
        const long long waves = (tiles + 147) / 148;
        const double score = ((double)tiles / (double)(waves * 148)) * ((double)c / (double)(c + 128));
        if (score > best) { best = score; BN = c; }
      }
      if (a->N <= 256 && a->N % 16 == 0 && mt >= 148) BN = a->N;  // one n-tile when M alone fills the GPU
    }
    if (geglu && BN % 32) BN = (BN + 31) / 32 * 32;
  }
  if (a->gn_coef && (!a->colstats || !a->residual))
    return fail(GD_UNET_ERR_INVALID_ARG, "gemm: the GroupNorm-backward epilogue needs colstats and residual (= the GroupNorm input)");
  if (BN % 16 || BN < 16 || BN > 256 || ((a->flags & GD_EPI_GEGLU) && BN % 32))
    return fail(GD_UNET_ERR_INVALID_ARG, "gemm: block_n must be a multiple of 16 in 16..256");

  // split-K (under-filled grids with long K) is decided first: it stays on the single-CTA kernel
  int ksplit = 1;
  const int num_kb_all = a->K / gdu::kBK;
  const int m_tiles_all = (a->M + gdu::kBM - 1) / gdu::kBM, n_tiles_all = (a->N + BN - 1) / BN;
  {
    const long long tiles = (long long)m_tiles_all * n_tiles_all * a->batch;
    const bool plain = !(a->flags & (GD_EPI_GEGLU | GD_EPI_TRANSPOSED)) && a->batch == 1 && a->N % 4 == 0 &&
                       a->c_batch_stride == 0 && a->c_head_stride == 0 && a->c_up2_w == 0;   // (the finalize kernel writes linear rows)
    if (plain && a->block_n <= 0 && tiles <= 74 && num_kb_all >= 24) {
      int ks = (int)(148 / tiles);
      if (ks > num_kb_all / 6) ks = num_kb_all / 6;
      if (ks > 16) ks = 16;
      if (ks >= 2 && (size_t)ks * a->M * a->N <= kSplitKFloats) ksplit = ks;
    }
  }
  // CTA pairs (cta_group::2): every GEMM with at least two M tiles per batch entry
  static const bool pair_enabled = []() { const char* e = getenv("GD_GEMM_PAIR"); return !(e && e[0] == '0'); }();
  const bool two = pair_enabled && ksplit == 1 && m_tiles_all >= 2;
  // haloed A tiles (one 130-pixel row per (dy, channel block) instead of three 128-pixel rows): 3x3 stride-1 convolutions over
  // images at least 128 pixels wide, CTA pairs, plain epilogue. GD_GEMM_HALO=0 switches it off (A/B timing).
  static const bool halo_enabled = []() { const char* e = getenv("GD_GEMM_HALO"); return !(e && e[0] == '0'); }();
  bool a_halo = halo_enabled && a->a_yscale != 2 && two && a->ntaps == 9 && a->a_box[1] == gdu::kBM && a->a_box[2] == 1 && a->a_box[3] == 1 && a->Ck % gdu::kBK == 0;
  for (int t = 0; a_halo && t < 9; t++) a_halo = a->tap_dx[t] == t % 3 - 1 && a->tap_dy[t] == t / 3 - 1 && a->tap_c[t] == 0;
  // images narrower than 128 pixels: M tiles become 8 x 16 pixel patches with a haloed A slot per (dy, channel block), see
  // GemmKParams.a_halo == 2. GD_GEMM_PATCH=0 switches it off (A/B timing).
  static const bool patch_enabled = []() { const char* e = getenv("GD_GEMM_PATCH"); return !(e && e[0] == '0'); }();
  static const bool tma_store_env = []() { const char* e = getenv("GD_GEMM_TMA_STORE"); return e && e[0] == '1'; }();
  bool a_patch = patch_enabled && a->a_yscale != 2 && !tma_store_env && !a_halo && two && a->ntaps == 9 && a->a_box[1] < gdu::kBM && a->a_box[3] == 1 &&
                 a->img_w % 8 == 0 && a->img_h % 16 == 0 && a->Ck % gdu::kBK == 0 && !(a->flags & (GD_EPI_GEGLU | GD_EPI_TRANSPOSED));
  for (int t = 0; a_patch && t < 9; t++) a_patch = a->tap_dx[t] == t % 3 - 1 && a->tap_dy[t] == t / 3 - 1 && a->tap_c[t] == 0;
  const size_t halo_bytes = a_patch ? (size_t)gdu::kPatchBytes : (size_t)gdu::kHaloBytes;
  if (a_patch) a_halo = true;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[4] = {(cuuint64_t)a->a_dim[0], (cuuint64_t)a->a_dim[1], (cuuint64_t)a->a_dim[2], (cuuint64_t)a->a_dim[3]};
    cuuint64_t str[3] = {(cuuint64_t)a->a_stride[0], (cuuint64_t)a->a_stride[1], (cuuint64_t)a->a_stride[2]};
    cuuint32_t box[4] = {(cuuint32_t)a->a_box[0], (cuuint32_t)(a->a_box[1] + (a_halo ? 2 : 0)), (cuuint32_t)a->a_box[2], (cuuint32_t)a->a_box[3]};
    if (a_patch) { box[1] = 10; box[2] = 16; box[3] = 1; }
    const int rc = make_map(&tmA, a->A, 4, dims, str, box);
    if (rc != GD_UNET_OK) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)a->b_dim[0], (cuuint64_t)a->b_dim[1], (cuuint64_t)a->b_dim[2]};
    cuuint64_t str[2] = {(cuuint64_t)a->b_stride[0], (cuuint64_t)a->b_stride[1]};
    cuuint32_t box[3] = {(cuuint32_t)gdu::kBK, (cuuint32_t)(two ? BN / 2 : BN), 1};   // a CTA of a pair stages half of the B tile
    const int rc = make_map(&tmB, a->B, 3, dims, str, box);
    if (rc != GD_UNET_OK) return rc;
  }
  gdu::GemmKParams p;
  memset(&p, 0, sizeof(p));
  p.M = a->M; p.N = a->N; p.num_kb = a->K / gdu::kBK;
  p.mode_conv = a->ntaps > 0 ? 1 : 0;
  p.kb_per_tap = p.mode_conv ? a->Ck / gdu::kBK : p.num_kb;
  for (int t = 0; t < 9; t++) { p.tap_dx[t] = a->tap_dx[t]; p.tap_dy[t] = a->tap_dy[t]; p.tap_c[t] = a->tap_c[t]; }
  p.rows_per_image = a->rows_per_image; p.img_w = a->img_w;
  p.rows_box = a->a_box[2]; p.imgs_box = a->a_box[3];
  p.box_w = a->a_box[1]; p.tiles_per_row = 1;
  if (p.mode_conv) {
    if (a->rows_per_image != a->img_w * a->img_h || a->img_h % a->a_box[2] || a->a_box[1] > a->img_w || a->img_w % a->a_box[1])
      return fail(GD_UNET_ERR_INVALID_ARG, "gemm: conv box must tile the image rows");
    if (a->a_box[1] < a->img_w && (a->a_box[2] != 1 || a->a_box[3] != 1))
      return fail(GD_UNET_ERR_INVALID_ARG, "gemm: a box narrower than the image must be one row of one image");
    if (a->a_box[3] > 1 && a->a_box[2] != a->img_h) return fail(GD_UNET_ERR_INVALID_ARG, "gemm: multi-image box must hold whole images");
    p.tiles_per_row = a->img_w / a->a_box[1];
  }
  p.heads = a->heads; p.a_head_k = a->a_head_k; p.a_zflat = a->a_zflat; p.b_head_k = a->b_head_k; p.b_head_n = a->b_head_n; p.b_zdim = a->b_dim[2];
  p.C = reinterpret_cast<__half*>(a->C); p.ldc = a->ldc; p.c_batch_stride = a->c_batch_stride; p.c_head_stride = a->c_head_stride;
  p.bias = reinterpret_cast<const __half*>(a->bias); p.row_bias = reinterpret_cast<const __half*>(a->row_bias);
  p.row_bias_ld = a->row_bias_ld > 0 ? a->row_bias_ld : a->N;
  p.residual = reinterpret_cast<const __half*>(a->residual);
  p.alpha = a->alpha; p.flags = a->flags; p.block_n = BN;
  if (a->c_up2_w < 0 || (a->c_up2_w > 0 && (a->M % a->c_up2_w || a->batch != 1 || (a->flags & (GD_EPI_GEGLU | GD_EPI_TRANSPOSED)) || a->colstats ||
                                            a->residual || a->row_bias)))
    return fail(GD_UNET_ERR_INVALID_ARG, "gemm: c_up2_w needs a plain single-batch GEMM without residual / statistics and M % c_up2_w == 0");
  p.c_up2_w = a->c_up2_w;
  if (a->a_yscale < 0 || a->a_yscale > 2 || (a->a_yscale == 2 && (a->ntaps < 1 || a->a_box[2] != 1 || a->a_box[3] != 1)))
    return fail(GD_UNET_ERR_INVALID_ARG, "gemm: a_yscale = 2 needs a tap GEMM whose A box is part of one image row");
  p.a_yscale = a->a_yscale == 2 ? 2 : 1;
  const size_t b_slot_bytes = (((size_t)(two ? BN / 2 : BN) * gdu::kBK * 2 + 1023) & ~(size_t)1023);
  const size_t stage_bytes = a_halo ? halo_bytes + 3 * b_slot_bytes : (size_t)gdu::kBM * gdu::kBK * 2 + b_slot_bytes;
  p.a_halo = a_patch ? 2 : a_halo ? 1 : 0;
  p.halo_bytes = (uint32_t)halo_bytes;
  p.patch_w_tiles = a_patch ? a->img_w / 8 : 1;
  p.patch_tiles_per_img = a_patch ? (a->img_w / 8) * (a->img_h / 16) : 1;
  // one persistent CTA per SM owns the shared memory: operand ring + epilogue staging + barriers/bias
  const int bias_stride = 32 * ((((BN + 31) / 32) + 1) / 2);   // bias floats per epilogue warp
  p.bias_stride = bias_stride;
  const size_t smem_max = 227 * 1024, fixed_noslack = 256 + 64 + (size_t)gdu::kEpiWarps * bias_stride * sizeof(float) +
                                                     (a->gn_coef ? (size_t)gdu::kEpiWarps * 128 * sizeof(float4) : 0);   // MODE 4 coefficient staging
  const size_t fixed = 1024 + fixed_noslack;
  const bool mode0 = !(a->flags & (GD_EPI_GEGLU | GD_EPI_TRANSPOSED));
  const bool tma_ok = mode0 && (a->ldc % 8) == 0 && (a->c_batch_stride % 8) == 0 && (a->c_head_stride % 8) == 0 &&
                      ((uintptr_t)a->C % 16) == 0 && (a->heads == 1 || a->c_head_stride > 0) &&
                      (a->batch == a->heads || a->c_batch_stride > 0);
  static const bool use_tma_store = []() { const char* e = getenv("GD_GEMM_TMA_STORE"); return e && e[0] == '1'; }();
  // GroupNorm-backward producer epilogue (MODE 4): CTA pairs, staged synchronous store, whole 32-column chunks, one image per
  // 32-row block. When the shape does not allow it the GEMM runs plain (residual ignored) and returns GD_UNET_NO_COLSTATS.
  const bool gnb = a->gn_coef && two && ksplit == 1 && mode0 && tma_ok && !use_tma_store && !(a->flags & GD_EPI_SILU) && a->N % 32 == 0 &&
                   BN % 32 == 0 && a->batch == 1 && a->rows_per_image > 0 && a->rows_per_image % 32 == 0;
  if (a->gn_coef && !gnb) p.residual = nullptr;
  p.gn_coef = gnb ? reinterpret_cast<const float4*>(a->gn_coef) : nullptr;
  int stg_bufs = tma_ok ? ((use_tma_store || gnb) ? 2 : 1) : 0;   // the synchronous staged store needs one buffer per warp (two: g | xh)
  int stages = (int)((smem_max - fixed - (size_t)gdu::kEpiWarps * stg_bufs * 2048) / stage_bytes);
  if (tma_ok && stages < 4) {
    stg_bufs = gnb ? 2 : 1;   // (the GroupNorm-backward epilogue stages two tiles per chunk: g and xh)
    stages = (int)((smem_max - fixed - (size_t)gdu::kEpiWarps * stg_bufs * 2048) / stage_bytes);
  }
  // CTA pairs whose B tile is the same for every output tile (one N tile, no batch): keep this CTA's half of B
  // resident in shared memory for the whole kernel; the ring then streams A only
  // (measured neutral on B200 -- 128->128 conv at 512^2: 367 us resident vs 378 us streamed -- so it is opt-in: GD_GEMM_BRES=1)
  static const int bres_env = []() { const char* e = getenv("GD_GEMM_BRES"); return e ? (e[0] == '1' ? 1 : 0) : -1; }();
  // default: on together with the haloed A tiles (128->128 conv at 4x512^2: 375 us streamed, 282 us with the halo, 256 us with the
  // halo and resident B), off otherwise (measured neutral there: those shapes are A-delivery bound)
  const bool bres_enabled = bres_env >= 0 ? bres_env == 1 : a_halo;
  p.b_resident = 0;
  if (bres_enabled && two && n_tiles_all == 1 && a->batch == 1 && a->heads == 1) {
    const size_t a_bytes = a_halo ? halo_bytes : (size_t)gdu::kBM * gdu::kBK * 2;
    const size_t b_slot = (((size_t)(BN / 2) * gdu::kBK * 2 + 1023) & ~(size_t)1023);
    const size_t res = (size_t)num_kb_all * b_slot, stg = (size_t)gdu::kEpiWarps * stg_bufs * 2048;
    if (res + stg + fixed_noslack + 3 * a_bytes <= smem_max && (size_t)m_tiles_all * a->batch >= 4 * 148) {
      p.b_resident = 1;
      stages = (int)((smem_max - fixed_noslack - stg - res) / a_bytes);
    }
  }
  if (stages > 6) stages = 6;   // deeper rings measured no faster (profiles/r01_gemm_pair_stage_sweep.txt)
  if (stages < 2) stages = 2;
  static const int stages_cap = []() { const char* e = getenv("GD_GEMM_STAGES"); return e ? atoi(e) : 0; }();   // tuning experiments
  if (stages_cap >= 2 && stages > stages_cap) stages = stages_cap;
  p.stages = stages;
  p.stg_bufs = stg_bufs;
  p.m_tiles = two ? (m_tiles_all + 1) / 2 : m_tiles_all;   // pairs of M tiles for the CTA-pair kernel
  p.n_tiles = n_tiles_all;
  p.ksplit = 1; p.kb_per_split = p.num_kb; p.ws = nullptr;
  // split-K partials in fp32, summed in a fixed order by k_splitk_finalize (deterministic)
  if (ksplit > 1) {
    DeviceScratch* sc = device_scratch();
    if (!sc) return fail(GD_UNET_ERR_CUDA, "gemm: split-K workspace (per-device scratch allocation failed)");
    float* ws = sc->splitk_ws;
    p.kb_per_split = (p.num_kb + ksplit - 1) / ksplit;
    p.ksplit = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;  // no empty splits
    p.ws = ws;
  }
  p.total_tiles = p.m_tiles * p.n_tiles * a->batch * p.ksplit;
  const size_t smem = p.b_resident
                          ? (size_t)stages * (a_halo ? halo_bytes : (size_t)gdu::kBM * gdu::kBK * 2) + (size_t)num_kb_all * (((size_t)(BN / 2) * gdu::kBK * 2 + 1023) & ~(size_t)1023) +
                                (size_t)gdu::kEpiWarps * stg_bufs * 2048 + fixed_noslack
                          : stages * stage_bytes + (size_t)gdu::kEpiWarps * stg_bufs * 2048 + fixed;
  // output tensor map for the staged epilogue: [batch][head][M][N], 32 x 32 box, 64B swizzle
  CUtensorMap tmC = tmA;
  p.tma_store = 0;
  if (tma_ok && p.ksplit == 1) {
    const cuuint64_t row_b = (cuuint64_t)a->ldc * 2;
    cuuint64_t dims[4] = {(cuuint64_t)a->N, (cuuint64_t)a->M, (cuuint64_t)a->heads, (cuuint64_t)(a->batch / a->heads)};
    cuuint64_t str[3] = {row_b, a->heads > 1 ? (cuuint64_t)a->c_head_stride * 2 : row_b,
                         a->batch > a->heads ? (cuuint64_t)a->c_batch_stride * 2 : row_b};
    cuuint32_t box[4] = {32, 32, 1, 1};
    const int rc = make_map(&tmC, a->C, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc != GD_UNET_OK) return rc;
    // 2 = staged + coalesced st.global (default); 1 = staged + TMA store (GD_GEMM_TMA_STORE=1)
    p.tma_store = (use_tma_store && a->c_up2_w == 0) ? 1 : 2;
  }
  // fused GroupNorm column statistics need the staged store path and whole 32-column chunks
  const bool colstats_ok = a->colstats && p.tma_store == 2 && a->N % 32 == 0 && BN % 32 == 0 && a->batch == 1 && (!a->gn_coef || gnb);
  p.colstats = colstats_ok ? reinterpret_cast<float*>(a->colstats) : nullptr;
  if (gnb && !colstats_ok) return fail(GD_UNET_ERR_INVALID_ARG, "gemm: internal: GroupNorm-backward epilogue without column statistics");
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    const int lim = 227 * 1024;
    if (cudaFuncSetAttribute(gdu::k_gemm_tcgen05<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim) != cudaSuccess ||
        cudaFuncSetAttribute(gdu::k_gemm_tcgen05<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim) != cudaSuccess ||
        cudaFuncSetAttribute(gdu::k_gemm_tcgen05<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim) != cudaSuccess ||
        cudaFuncSetAttribute(gdu::k_gemm_tcgen05<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim) != cudaSuccess ||
        cudaFuncSetAttribute(gdu::k_gemm_tcgen05<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim) != cudaSuccess ||
        cudaFuncSetAttribute(gdu::k_gemm_tcgen05<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim) != cudaSuccess ||
        cudaFuncSetAttribute(gdu::k_gemm_tcgen05<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim) != cudaSuccess ||
        cudaFuncSetAttribute(gdu::k_gemm_tcgen05<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim) != cudaSuccess)
      return fail(GD_UNET_ERR_CUDA, "gemm: cannot raise dynamic shared memory limit");
  }
  const dim3 blk(gdu::kGemmThreads);
  if (two) {
    g_pair_launches.fetch_add(1, std::memory_order_relaxed);
    const int pairs = num_sms / 2;
    const dim3 grid(2 * (p.total_tiles < pairs ? p.total_tiles : pairs));
    if (a->flags & GD_EPI_GEGLU) launch_pdl_cluster(2, gdu::k_gemm_tcgen05<1, true>, grid, blk, smem, stream, tmA, tmB, tmC, p);
    else if (a->flags & GD_EPI_TRANSPOSED) launch_pdl_cluster(2, gdu::k_gemm_tcgen05<2, true>, grid, blk, smem, stream, tmA, tmB, tmC, p);
    else if (gnb) launch_pdl_cluster(2, gdu::k_gemm_tcgen05<4, true>, grid, blk, smem, stream, tmA, tmB, tmC, p);
    else launch_pdl_cluster(2, gdu::k_gemm_tcgen05<0, true>, grid, blk, smem, stream, tmA, tmB, tmC, p);
  } else {
    const dim3 grid(p.total_tiles < num_sms ? p.total_tiles : num_sms);
    if (p.ksplit > 1) launch_pdl(gdu::k_gemm_tcgen05<3, false>, grid, blk, smem, stream, tmA, tmB, tmC, p);
    else if (a->flags & GD_EPI_GEGLU) launch_pdl(gdu::k_gemm_tcgen05<1, false>, grid, blk, smem, stream, tmA, tmB, tmC, p);
    else if (a->flags & GD_EPI_TRANSPOSED) launch_pdl(gdu::k_gemm_tcgen05<2, false>, grid, blk, smem, stream, tmA, tmB, tmC, p);
    else launch_pdl(gdu::k_gemm_tcgen05<0, false>, grid, blk, smem, stream, tmA, tmB, tmC, p);
  }
  LAUNCH_CHECK("k_gemm_tcgen05");
  if (p.ksplit > 1) {
    const long long n4 = (long long)a->M * a->N / 4;
    launch_pdl(gdu::k_splitk_finalize, dim3((unsigned)((n4 + 255) / 256)), dim3(256), (size_t)(0), (cudaStream_t)(stream), 
        p.ws, p.ksplit, a->M, a->N, a->alpha, p.bias, p.row_bias, p.row_bias_ld, a->rows_per_image > 0 ? a->rows_per_image : 1, p.residual,
        p.C, p.ldc, p.flags);
    LAUNCH_CHECK("k_splitk_finalize");
  }
  return (a->colstats && !colstats_ok) ? GD_UNET_NO_COLSTATS : GD_UNET_OK;
}

int gd_unet_flash_attn(const void* q, const void* k, const void* vt, void* out, int B, int heads, int Tq, int Tk,
                       long long ldq, long long ldk, long long ldv, long long ldo, float scale, gd_ustream_t stream_) {
  return gd_unet_flash_attn_ex(q, k, vt, out, B, heads, Tq, Tk, ldq, ldk, ldv, ldo, 0, scale, stream_);
}
int gd_unet_flash_attn_ex(const void* q, const void* k, const void* vt, void* out, int B, int heads, int Tq, int Tk,
                          long long ldq, long long ldk, long long ldv, long long ldo, long long v_batch_stride, float scale,
                          gd_ustream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!q || !k || !vt || !out || B < 1 || heads < 1 || Tq < 1 || Tk < 1) return fail(GD_UNET_ERR_INVALID_ARG, "flash_attn: bad argument");
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8) || (v_batch_stride % 8))
    return fail(GD_UNET_ERR_INVALID_ARG, "flash_attn: leading dims must be multiples of 8");
  const int C = heads * 64;
  CUtensorMap tmQ, tmK, tmV;
  {
    cuuint64_t d[3] = {(cuuint64_t)C, (cuuint64_t)Tq, (cuuint64_t)B};
    cuuint64_t st[2] = {(cuuint64_t)ldq * 2, (cuuint64_t)Tq * ldq * 2};
    cuuint32_t box[3] = {64, 128, 1};
    int rc = make_map(&tmQ, q, 3, d, st, box);
    if (rc != GD_UNET_OK) return rc;
    cuuint64_t dk[3] = {(cuuint64_t)C, (cuuint64_t)Tk, (cuuint64_t)B};
    cuuint64_t sk[2] = {(cuuint64_t)ldk * 2, (cuuint64_t)Tk * ldk * 2};
    rc = make_map(&tmK, k, 3, dk, sk, box);
    if (rc != GD_UNET_OK) return rc;
    cuuint64_t dv[3] = {(cuuint64_t)Tk, (cuuint64_t)C, (cuuint64_t)B};
    cuuint64_t sv[2] = {(cuuint64_t)ldv * 2, (cuuint64_t)(v_batch_stride > 0 ? v_batch_stride : C * ldv) * 2};
    cuuint32_t boxv[3] = {64, 64, 1};
    rc = make_map(&tmV, vt, 3, dv, sv, boxv);
    if (rc != GD_UNET_OK) return rc;
  }
  gdu::AttnParams p;
  p.Tq = Tq; p.Tk = Tk; p.heads = heads; p.n_kv = (Tk + 127) / 128;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.O = reinterpret_cast<__half*>(out); p.ldo = ldo;
  static const int stagger = []() { const char* e = getenv("GD_ATTN_STAGGER"); return e ? atoi(e) : 800; }();
  p.stagger = stagger;
  const size_t smem = gdu::kQBytes + 2 * gdu::kKBytes + 2 * gdu::kVBytes + 2 * gdu::kPBytes + 1024 + 256 + 1024 + 64;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gdu::k_flash_attn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(gdu::k_flash_attn2, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return fail(GD_UNET_ERR_CUDA, "flash_attn: cannot raise dynamic shared memory limit");
    attr_set = true;
  }
  // long key ranges (self-attention of the 64^2 / 32^2 / 16^2 latent layers): single-pass kernel, two Q tiles per CTA.
  // GD_ATTN_TWO_PASS=1 keeps the two-pass kernel for A/B timing.
  static const bool two_pass_only = []() { const char* e = getenv("GD_ATTN_TWO_PASS"); return e && e[0] == '1'; }();
  static const int attn2_min_kv = []() { const char* e = getenv("GD_ATTN2_MIN_KV"); return e ? atoi(e) : 1; }();
  if (!two_pass_only && p.n_kv >= attn2_min_kv && Tq >= 128) {
    const size_t smem2 = 2 * gdu::kQBytes + gdu::kAttn2Stages * (gdu::kKBytes + gdu::kVBytes) + 1024 + 512;
    launch_pdl(gdu::k_flash_attn2, dim3((Tq + 255) / 256, B * heads), dim3(gdu::kAttn2Threads), smem2, stream, tmQ, tmK, tmV, p);
    LAUNCH_CHECK("k_flash_attn2");
    return GD_UNET_OK;
  }
  launch_pdl(gdu::k_flash_attn, dim3(dim3((Tq + 127) / 128, B * heads)), dim3(gdu::kAttnThreads), (size_t)(smem), (cudaStream_t)(stream), tmQ, tmK, tmV, p);
  LAUNCH_CHECK("k_flash_attn");
  return GD_UNET_OK;
}

int gd_unet_groupnorm(const void* x, void* y, const void* gamma, const void* beta, int N, int HW, int C, int groups,
                      float eps, int silu, gd_ustream_t s) {
  if (C % groups || (C / groups) % 2 || C % 8 || N * groups > 4096 || groups > 256 || C > 2560)
    return fail(GD_UNET_ERR_INVALID_ARG, "groupnorm: channels per group must be even, C % 8 == 0");
  // One launch with a cluster of 8 CTAs per image (DSMEM exchange of the partial sums). Measured SLOWER than the
  // two-launch path at batch 8 (UNet forward 16.6 vs 13.3 ms: 64 CTAs cannot hide the L2 latency), so opt-in only.
  static const bool gn_cluster = []() { const char* e = getenv("GD_GN_CLUSTER"); return e && e[0] == '1'; }();
  if (gn_cluster && N * 8 <= 148 && HW >= 64 && (size_t)N * HW * C * 2 <= ((size_t)96 << 20)) {
    launch_pdl_cluster(8, gdu::k_gn_cluster, dim3(N * 8), dim3(256), (size_t)(sizeof(float2) * C), (cudaStream_t)s, (const __half*)x, (__half*)y,
                       (const __half*)gamma, (const __half*)beta, HW, C, groups, eps, silu);
    LAUNCH_CHECK("k_gn_cluster");
    return GD_UNET_OK;
  }
  // partial statistics live in a small static device buffer (N*groups*splits float2 <= 512 KB)
  DeviceScratch* sc = device_scratch();
  if (!sc) return fail(GD_UNET_ERR_CUDA, "groupnorm: per-device scratch allocation failed");
  float2* part = sc->gn_part_small;
  int splits = (HW * (C / 8) + 8191) / 8192;  // ~32 sixteen-byte loads per thread
  if (splits > 16) splits = 16;
  if (splits < 1) splits = 1;
  while (N * splits < 296 && splits < 64 && HW / (splits * 2) >= 16) splits *= 2;  // at least ~2 CTAs per SM
  if ((size_t)N * groups * splits > kGnSmall) return fail(GD_UNET_ERR_INVALID_ARG, "groupnorm: N * groups * splits exceeds the partial-statistics scratch");
  launch_pdl(gdu::k_gn_stats, dim3(dim3(N, splits)), dim3(256), (size_t)(0), (cudaStream_t)((cudaStream_t)s), (const __half*)x, part, HW, C, groups, splits);
  LAUNCH_CHECK("k_gn_stats");
  int pix_per_cta = (int)((16384 + C - 1) / C);  // ~16k elements per CTA
  if (pix_per_cta < 1) pix_per_cta = 1;
  launch_pdl(gdu::k_gn_apply, dim3(dim3(N, (HW + pix_per_cta - 1) / pix_per_cta)), dim3(256), (size_t)(sizeof(float2) * C), (cudaStream_t)((cudaStream_t)s), 
      (const __half*)x, (__half*)y, part, (const __half*)gamma, (const __half*)beta, HW, C, groups, splits, eps, silu,
      pix_per_cta);
  LAUNCH_CHECK("k_gn_apply");
  return GD_UNET_OK;
}
int gd_unet_layernorm(const void* x, void* y, const void* gamma, const void* beta, int rows, int C, float eps, gd_ustream_t s) {
  if (C % 2) return fail(GD_UNET_ERR_INVALID_ARG, "layernorm: C must be even");
  launch_pdl(gdu::k_layernorm, dim3((rows + 7) / 8), dim3(256), (size_t)(0), (cudaStream_t)((cudaStream_t)s), (const __half*)x, (__half*)y, (const __half*)gamma,
                                                               (const __half*)beta, rows, C, eps);
  LAUNCH_CHECK("k_layernorm");
  return GD_UNET_OK;
}
int gd_unet_softmax(void* sc, long long rows, int cols, long long ld, gd_ustream_t s) {
  launch_pdl(gdu::k_softmax, dim3((unsigned)((rows + 7) / 8)), dim3(256), (size_t)(0), (cudaStream_t)((cudaStream_t)s), (__half*)sc, rows, cols, ld);
  LAUNCH_CHECK("k_softmax");
  return GD_UNET_OK;
}
int gd_unet_geglu(const void* x, void* y, long long rows, int H, gd_ustream_t s) {
  launch_pdl(gdu::k_geglu, dim3((unsigned)((rows * H + 255) / 256)), dim3(256), (size_t)(0), (cudaStream_t)((cudaStream_t)s), (const __half*)x, (__half*)y, rows, H);
  LAUNCH_CHECK("k_geglu");
  return GD_UNET_OK;
}
int gd_unet_add(const void* a, const void* b, void* y, long long n, gd_ustream_t s) {
  if (n % 2) return fail(GD_UNET_ERR_INVALID_ARG, "add: n must be even");
  launch_pdl(gdu::k_add, dim3((unsigned)((n / 2 + 255) / 256)), dim3(256), (size_t)(0), (cudaStream_t)((cudaStream_t)s), (const __half2*)a, (const __half2*)b, (__half2*)y, n / 2);
  LAUNCH_CHECK("k_add");
  return GD_UNET_OK;
}
int gd_unet_upsample2x(const void* x, void* y, int N, int H, int W, int C, gd_ustream_t s) {
  if (C % 8) return fail(GD_UNET_ERR_INVALID_ARG, "upsample: C % 8");
  const long long total = (long long)N * 4 * H * W * (C / 8);
  launch_pdl(gdu::k_upsample2x, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)(0), (cudaStream_t)((cudaStream_t)s), (const uint4*)x, (uint4*)y, N, H, W, C / 8);
  LAUNCH_CHECK("k_upsample2x");
  return GD_UNET_OK;
}
int gd_unet_space_to_depth(const void* x, void* y, int N, int H, int W, int C, gd_ustream_t s) {
  if (C % 8 || H % 2 || W % 2) return fail(GD_UNET_ERR_INVALID_ARG, "space_to_depth: C % 8, even H/W");
  const long long total = (long long)N * H * W * (C / 8);
  launch_pdl(gdu::k_space_to_depth, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)(0), (cudaStream_t)((cudaStream_t)s), (const uint4*)x, (uint4*)y, N, H, W, C / 8);
  LAUNCH_CHECK("k_space_to_depth");
  return GD_UNET_OK;
}
int gd_unet_concat(const void* a, const void* b, void* y, long long rows, int Ca, int Cb, gd_ustream_t s) {
  if (Ca % 8 || Cb % 8) return fail(GD_UNET_ERR_INVALID_ARG, "concat: C % 8");
  const long long total = rows * ((Ca + Cb) / 8);
  launch_pdl(gdu::k_concat, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)(0), (cudaStream_t)((cudaStream_t)s), (const uint4*)a, (const uint4*)b, (uint4*)y, rows, Ca / 8, Cb / 8);
  LAUNCH_CHECK("k_concat");
  return GD_UNET_OK;
}
int gd_unet_small_linear(const void* x, const void* W, const void* bias, void* y, int Bm, int K, int N, int silu_in,
                         int silu_out, gd_ustream_t s) {
  if (Bm < 1 || Bm > 16 || K % 8 || K > 2048) return fail(GD_UNET_ERR_INVALID_ARG, "small_linear: 1 <= Bm <= 16, K % 8 == 0, K <= 2048");
  const size_t smem = (size_t)Bm * K * 2;
  const dim3 grid((N + 7) / 8);
#define GD_SL(KV) launch_pdl(gdu::k_small_linear<KV>, dim3(grid), dim3(256), (size_t)(smem), (cudaStream_t)((cudaStream_t)s), (const __half*)x, (const __half*)W, (const __half*)bias, (__half*)y, Bm, K, N, silu_in, silu_out)
  if (K <= 512) GD_SL(2); else if (K <= 1280) GD_SL(5); else GD_SL(8);
#undef GD_SL
  LAUNCH_CHECK("k_small_linear");
  return GD_UNET_OK;
}
int gd_unet_timestep_embedding(const float* t, void* y, int Bm, int dim, gd_ustream_t s) {
  launch_pdl(gdu::k_timestep_embedding, dim3((Bm * dim / 2 + 127) / 128), dim3(128), (size_t)(0), (cudaStream_t)((cudaStream_t)s), t, (__half*)y, Bm, dim);
  LAUNCH_CHECK("k_timestep_embedding");
  return GD_UNET_OK;
}
int gd_unet_conv_in(const void* x, const void* w, const void* bias, void* y, int N, int H, int W, int Cout, gd_ustream_t s) {
  if (Cout % 8 || Cout > 640) return fail(GD_UNET_ERR_INVALID_ARG, "conv_in: Cout % 8, Cout <= 640");
  const long long pix = (long long)N * H * W;
  launch_pdl(gdu::k_conv_in, dim3((unsigned)((pix + 127) / 128)), dim3(128), (size_t)((size_t)Cout * 37 * 2), (cudaStream_t)((cudaStream_t)s), 
      (const __half*)x, (const __half*)w, (const __half*)bias, (__half*)y, N, H, W, Cout);
  LAUNCH_CHECK("k_conv_in");
  return GD_UNET_OK;
}
int gd_unet_im2col4(const void* x, void* A, int N, int H, int W, gd_ustream_t s) {
  if (!x || !A || N < 1 || H < 1 || W < 1) return fail(GD_UNET_ERR_INVALID_ARG, "im2col4: bad argument");
  const long long t = (long long)N * H * W * 8;
  launch_pdl(gdu::k_unet_im2col4, dim3((unsigned)((t + 255) / 256)), dim3(256), (size_t)0, (cudaStream_t)s, (const __half*)x, (__half*)A, N, H, W);
  LAUNCH_CHECK("k_unet_im2col4");
  return GD_UNET_OK;
}
int gd_unet_unpack4_nchw(const void* x, float* y, int N, long long HW, int ld, gd_ustream_t s) {
  if (!x || !y || N < 1 || HW < 1 || ld < 4 || ld % 4) return fail(GD_UNET_ERR_INVALID_ARG, "unpack4_nchw: ld must be a multiple of 4");
  const long long t = (long long)N * HW;
  launch_pdl(gdu::k_unpack4_nchw, dim3((unsigned)((t + 255) / 256)), dim3(256), (size_t)0, (cudaStream_t)s, (const __half*)x, y, N, HW, ld);
  LAUNCH_CHECK("k_unpack4_nchw");
  return GD_UNET_OK;
}
int gd_unet_conv_out(const void* x, const void* w, const void* bias, float* y, int N, int H, int W, int Cin, gd_ustream_t s) {
  const long long pix = (long long)N * H * W;
  launch_pdl(gdu::k_conv_out, dim3((unsigned)((pix + 7) / 8)), dim3(256), (size_t)(0), (cudaStream_t)((cudaStream_t)s), (const __half*)x, (const __half*)w,
                                                                         (const __half*)bias, y, N, H, W, Cin);
  LAUNCH_CHECK("k_conv_out");
  return GD_UNET_OK;
}
int gd_unet_add_noise(const float* lat, const float* noise, const float* sa, const float* sb, float* noisy, void* uin, int B,
                      int reps, int chw, gd_ustream_t s) {
  launch_pdl(gdu::k_add_noise, dim3((B * chw + 255) / 256), dim3(256), (size_t)(0), (cudaStream_t)((cudaStream_t)s), lat, noise, sa, sb, noisy, (__half*)uin, B, reps, chw);
  LAUNCH_CHECK("k_add_noise");
  return GD_UNET_OK;
}
int gd_unet_pool_latents(const float* color, const float* mix, float* latents, int B, int H, int W, gd_ustream_t s) {
  if (H % 8 || W % 8) return fail(GD_UNET_ERR_INVALID_ARG, "pool_latents: H, W multiples of 8");
  const int n = B * (H / 8) * (W / 8);
  launch_pdl(gdu::k_pool_latents, dim3((n + 127) / 128), dim3(128), (size_t)(0), (cudaStream_t)((cudaStream_t)s), color, mix, latents, B, H, W);
  LAUNCH_CHECK("k_pool_latents");
  return GD_UNET_OK;
}
int gd_unet_pool_latents_bwd(const float* grad, const float* mix, float* dcolor, int B, int H, int W, float clip, float scale,
                             gd_ustream_t s) {
  if (H % 8 || W % 8) return fail(GD_UNET_ERR_INVALID_ARG, "pool_latents_bwd: H, W multiples of 8");
  const long long n = (long long)B * 3 * H * W;
  launch_pdl(gdu::k_pool_latents_bwd, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)(0), (cudaStream_t)((cudaStream_t)s), grad, mix, dcolor, B, H, W, clip, scale);
  LAUNCH_CHECK("k_pool_latents_bwd");
  return GD_UNET_OK;
}
// ---- VAE encoder support (include/gd_unet.h, "VAE" section) -----------------------------------
namespace {
float2* gn_part_buffer() {   // partial statistics: up to 8192 (image, group) x 512 splits, per device
  DeviceScratch* sc = device_scratch();
  return sc ? sc->gn_part : nullptr;
}
// pixels per CTA of the fast sweeps: up to 128 K elements per CTA, but at least ~600 CTAs
int gn_fast_pix_per_cta(int N, int HW, int C, int unroll) {
  const int step = (256 / (C / 8)) * unroll;             // pixels one CTA iteration covers
  long long ppc = 131072 / C;
  const long long fill = ((long long)N * HW + 599) / 600;
  if (ppc > fill) ppc = fill;
  if (ppc < step) ppc = step;
  ppc = (ppc + step - 1) / step * step;
  return (int)ppc;
}
int gn_big_splits(int N, int HW, int C) {
  long long per = 4096;                                    // ~4096 sixteen-byte loads per CTA
  long long splits = ((long long)HW * (C / 8) + per - 1) / per;
  if (splits > 512) splits = 512;
  while (N * splits < 296 && splits < 512 && HW / (splits * 2) >= 16) splits *= 2;
  if (splits < 1) splits = 1;
  return (int)splits;
}
}  // namespace

// y = GN(x) (+SiLU) from the finalised (mean, rstd) table
static int gn_apply_with_table(const void* x, void* y, const void* gamma, const void* beta, const float* stats, int N, int HW, int C,
                               int groups, int silu, cudaStream_t s) {
  const int C8 = C / 8;
  if (256 % C8 == 0) {   // fast sweep: 4 loads in flight per thread
    int pix_per_cta = gn_fast_pix_per_cta(N, HW, C, 2);
    const long long chunks = ((long long)HW + pix_per_cta - 1) / pix_per_cta;
    if (chunks > 65535) return fail(GD_UNET_ERR_INVALID_ARG, "groupnorm_stats: image too large");
    if (silu) launch_pdl(gdu::k_gn_apply_fast<true, 2>, dim3(N, (unsigned)chunks), dim3(256), (size_t)0, s, (const __half*)x, (__half*)y,
                         (const float2*)stats, (const __half*)gamma, (const __half*)beta, HW, C, groups, pix_per_cta);
    else launch_pdl(gdu::k_gn_apply_fast<false, 2>, dim3(N, (unsigned)chunks), dim3(256), (size_t)0, s, (const __half*)x, (__half*)y,
                    (const float2*)stats, (const __half*)gamma, (const __half*)beta, HW, C, groups, pix_per_cta);
    LAUNCH_CHECK("k_gn_apply_fast");
    return GD_UNET_OK;
  }
  int pix_per_cta = (32768 + C - 1) / C;
  const long long chunks = ((long long)HW + pix_per_cta - 1) / pix_per_cta;
  if (chunks > 65535) return fail(GD_UNET_ERR_INVALID_ARG, "groupnorm_stats: image too large");
  launch_pdl(gdu::k_gn_apply_final, dim3(N, (unsigned)chunks), dim3(256), (size_t)(sizeof(float2) * C), s, (const __half*)x,
             (__half*)y, (const float2*)stats, (const __half*)gamma, (const __half*)beta, HW, C, groups, silu, pix_per_cta);
  LAUNCH_CHECK("k_gn_apply_final");
  return GD_UNET_OK;
}

int gd_unet_groupnorm_stats(const void* x, void* y, const void* gamma, const void* beta, float* stats, int N, int HW,
                            int C, int groups, float eps, int silu, gd_ustream_t s_) {
  cudaStream_t s = (cudaStream_t)s_;
  if (!stats) return fail(GD_UNET_ERR_INVALID_ARG, "groupnorm_stats: stats buffer required");
  if (C % groups || (C / groups) % 2 || C % 8 || N * groups > 8192 || groups > 256 || C > 2560)
    return fail(GD_UNET_ERR_INVALID_ARG, "groupnorm_stats: channels per group must be even, C % 8 == 0, C <= 2560");
  float2* part = gn_part_buffer();
  if (!part) return fail(GD_UNET_ERR_CUDA, "groupnorm_stats: cudaMalloc");
  const int splits = gn_big_splits(N, HW, C);
  if ((size_t)N * groups * splits > kGnBig) return fail(GD_UNET_ERR_INVALID_ARG, "groupnorm_stats: N * groups * splits exceeds the scratch");
  launch_pdl(gdu::k_gn_stats, dim3(N, splits), dim3(256), (size_t)0, s, (const __half*)x, part, HW, C, groups, splits);
  LAUNCH_CHECK("k_gn_stats");
  const int total = N * groups;
  launch_pdl(gdu::k_gn_finalize, dim3((total + 7) / 8), dim3(256), (size_t)0, s, (const float2*)part, (float2*)stats, total,
             splits, 1.0f / ((float)HW * (float)(C / groups)), eps);
  LAUNCH_CHECK("k_gn_finalize");
  if (y) return gn_apply_with_table(x, y, gamma, beta, stats, N, HW, C, groups, silu, s);
  return GD_UNET_OK;
}

int gd_unet_groupnorm_colstats(const void* x, void* y, const void* gamma, const void* beta, float* mean_rstd, const float* statsA,
                               int Ca, const float* statsB, int Cb, int N, int HW, int C, int groups, float eps, int silu,
                               gd_ustream_t s_) {
  cudaStream_t s = (cudaStream_t)s_;
  if (!x || !y || !statsA || Ca < 1 || Cb < 0 || (Cb > 0 && !statsB) || Ca + Cb != C)
    return fail(GD_UNET_ERR_INVALID_ARG, "groupnorm_colstats: bad pointers / channel split");
  if (C % groups || (C / groups) % 2 || C % 8 || Ca % 4 || Cb % 4 || groups > 256 || C > 2560 || HW % 32 || N * groups > 8192)
    return fail(GD_UNET_ERR_INVALID_ARG, "groupnorm_colstats: C % 8 == 0, Ca % 4 == 0, Cb % 4 == 0, even channels per group, HW % 32 == 0");
  DeviceScratch* sc = device_scratch();
  if (!sc) return fail(GD_UNET_ERR_CUDA, "groupnorm_colstats: per-device scratch allocation failed");
  float2* part = mean_rstd ? sc->gn_part : sc->gn_part_small;
  const size_t cap = mean_rstd ? kGnBig : kGnSmall;
  const int rbpi = HW / 32;
  // ~2048 sixteen-byte loads per CTA, at least ~2 CTAs per SM when the image has the row blocks for it
  long long splits = ((long long)rbpi * (C / 2) + 2047) / 2048;
  if (splits > 512) splits = 512;
  while (N * splits < 296 && splits * 2 <= rbpi && splits < 512) splits *= 2;
  if (splits > rbpi) splits = rbpi;
  if (splits < 1) splits = 1;
  while ((size_t)N * groups * splits > cap && splits > 1) splits /= 2;
  if ((size_t)N * groups * splits > cap) return fail(GD_UNET_ERR_INVALID_ARG, "groupnorm_colstats: N * groups exceeds the scratch");
  launch_pdl(gdu::k_gn_colstats_reduce, dim3(N, (unsigned)splits), dim3(256), (size_t)0, s, statsA, Ca, statsB, Cb, rbpi, groups,
             (int)splits, part, (const __half*)nullptr);
  LAUNCH_CHECK("k_gn_colstats_reduce");
  if (mean_rstd) {   // VAE flavour: (mean, rstd) table kept for the backward, fast sweeps
    const int total = N * groups;
    launch_pdl(gdu::k_gn_finalize, dim3((total + 7) / 8), dim3(256), (size_t)0, s, (const float2*)part, (float2*)mean_rstd, total,
               (int)splits, 1.0f / ((float)HW * (float)(C / groups)), eps);
    LAUNCH_CHECK("k_gn_finalize");
    return gn_apply_with_table(x, y, gamma, beta, mean_rstd, N, HW, C, groups, silu, s);
  }
  int pix_per_cta = (int)((16384 + C - 1) / C);
  if (pix_per_cta < 1) pix_per_cta = 1;
  launch_pdl(gdu::k_gn_apply, dim3(N, (HW + pix_per_cta - 1) / pix_per_cta), dim3(256), (size_t)(sizeof(float2) * C), s, (const __half*)x,
             (__half*)y, (const float2*)part, (const __half*)gamma, (const __half*)beta, HW, C, groups, (int)splits, eps, silu, pix_per_cta);
  LAUNCH_CHECK("k_gn_apply");
  return GD_UNET_OK;
}

int gd_unet_gn_bwd_coef(const float* stats, const void* gamma, const void* beta, float* coef, int N, int C, int groups, gd_ustream_t s_) {
  if (!stats || !gamma || !beta || !coef || N < 1 || C < 1 || groups < 1 || C % groups)
    return fail(GD_UNET_ERR_INVALID_ARG, "gn_bwd_coef: bad argument");
  launch_pdl(gdu::k_gn_bwd_coef, dim3((unsigned)((N * C + 255) / 256)), dim3(256), (size_t)0, (cudaStream_t)s_, (const float2*)stats,
             (const __half*)gamma, (const __half*)beta, (float4*)coef, N, C, groups);
  LAUNCH_CHECK("k_gn_bwd_coef");
  return GD_UNET_OK;
}
int gd_unet_groupnorm_bwd_g(const void* x, const void* g, const void* add, void* dx, const void* gamma, const void* beta,
                            const float* stats, const float* colstats, int N, int HW, int C, int groups, gd_ustream_t s_) {
  cudaStream_t s = (cudaStream_t)s_;
  if (!x || !g || !dx || !gamma || !beta || !stats || !colstats) return fail(GD_UNET_ERR_INVALID_ARG, "groupnorm_bwd_g: null pointer");
  if (C % groups || C % 8 || N * groups > 8192 || groups > 256 || C > 2048 || HW % 32)
    return fail(GD_UNET_ERR_INVALID_ARG, "groupnorm_bwd_g: C % groups == 0, C % 8 == 0, C <= 2048, HW % 32 == 0");
  float2* part = gn_part_buffer();
  if (!part) return fail(GD_UNET_ERR_CUDA, "groupnorm_bwd_g: cudaMalloc");
  float2* bstats = device_scratch()->gn_bstats;
  const int rbpi = HW / 32;
  long long splits = ((long long)rbpi * (C / 2) + 2047) / 2048;
  if (splits > 512) splits = 512;
  while (N * splits < 296 && splits * 2 <= rbpi && splits < 512) splits *= 2;
  if (splits > rbpi) splits = rbpi;
  if (splits < 1) splits = 1;
  if ((size_t)N * groups * splits > kGnBig) return fail(GD_UNET_ERR_INVALID_ARG, "groupnorm_bwd_g: N * groups * splits exceeds the scratch");
  launch_pdl(gdu::k_gn_colstats_reduce, dim3(N, (unsigned)splits), dim3(256), (size_t)0, s, colstats, C, (const float*)nullptr, 0, rbpi,
             groups, (int)splits, part, (const __half*)gamma);
  LAUNCH_CHECK("k_gn_colstats_reduce");
  const int total = N * groups;
  launch_pdl(gdu::k_gn_bwd_finalize, dim3((total + 7) / 8), dim3(256), (size_t)0, s, (const float2*)part, bstats, total, (int)splits,
             1.0f / ((float)HW * (float)(C / groups)));
  LAUNCH_CHECK("k_gn_bwd_finalize");
  // g already carries silu'(y): the plain (no-activation) apply sweep finishes dx = rstd*(gamma*g - S1 - xh*S2) (+ add)
  if (256 % (C / 8) == 0) {
    const int pix_per_cta = gn_fast_pix_per_cta(N, HW, C, 2);
    const long long chunks = ((long long)HW + pix_per_cta - 1) / pix_per_cta;
    if (chunks > 65535) return fail(GD_UNET_ERR_INVALID_ARG, "groupnorm_bwd_g: image too large");
    launch_pdl(gdu::k_gn_bwd_apply_fast<false, 2>, dim3(N, (unsigned)chunks), dim3(256), (size_t)0, s, (const __half*)x, (const __half*)g,
               (const __half*)add, (__half*)dx, (const float2*)stats, (const float2*)bstats, (const __half*)gamma,
               (const __half*)beta, HW, C, groups, pix_per_cta);
    LAUNCH_CHECK("k_gn_bwd_apply_fast");
    return GD_UNET_OK;
  }
  const int pix_per_cta = (32768 + C - 1) / C;
  const long long chunks = ((long long)HW + pix_per_cta - 1) / pix_per_cta;
  if (chunks > 65535) return fail(GD_UNET_ERR_INVALID_ARG, "groupnorm_bwd_g: image too large");
  launch_pdl(gdu::k_gn_bwd_apply, dim3(N, (unsigned)chunks), dim3(256), (size_t)((sizeof(float4) + sizeof(float2)) * C), s, (const __half*)x,
             (const __half*)g, (const __half*)add, (__half*)dx, (const float2*)stats, (const float2*)bstats, (const __half*)gamma,
             (const __half*)beta, HW, C, groups, 0, pix_per_cta);
  LAUNCH_CHECK("k_gn_bwd_apply");
  return GD_UNET_OK;
}

int gd_unet_groupnorm_bwd(const void* x, const void* dz, const void* add, void* dx, const void* gamma, const void* beta,
                          const float* stats, int N, int HW, int C, int groups, int silu, gd_ustream_t s_) {
  cudaStream_t s = (cudaStream_t)s_;
  if (C % groups || C % 8 || N * groups > 8192 || groups > 256 || C > 2048)
    return fail(GD_UNET_ERR_INVALID_ARG, "groupnorm_bwd: C % groups == 0, C % 8 == 0, C <= 2048");
  float2* part = gn_part_buffer();
  if (!part) return fail(GD_UNET_ERR_CUDA, "groupnorm_bwd: cudaMalloc");
  float2* bstats = device_scratch()->gn_bstats;
  const int splits = gn_big_splits(N, HW, C);
  if ((size_t)N * groups * splits > kGnBig) return fail(GD_UNET_ERR_INVALID_ARG, "groupnorm_bwd: N * groups * splits exceeds the scratch");
  if (256 % (C / 8) == 0) {   // register-resident coefficients, software-pipelined loads. GD_GN_BWD_STATS_PIPE=0: 8 plain loads in
    // flight per thread instead (measured slower: VAE backward 7.48 vs 7.37 ms)
    static const bool pipe = []() { const char* e = getenv("GD_GN_BWD_STATS_PIPE"); return !(e && e[0] == '0'); }();
#define GD_BS(S_, U_, P_) launch_pdl(gdu::k_gn_bwd_stats_fast<S_, U_, P_>, dim3(N, splits), dim3(256), (size_t)0, s, (const __half*)x, \
    (const __half*)dz, (const float2*)stats, (const __half*)gamma, (const __half*)beta, part, HW, C, groups, splits)
    if (pipe) { if (silu) GD_BS(true, 2, true); else GD_BS(false, 2, true); }
    else { if (silu) GD_BS(true, 4, false); else GD_BS(false, 4, false); }
#undef GD_BS
  } else {
    launch_pdl(gdu::k_gn_bwd_stats, dim3(N, splits), dim3(256), (size_t)0, s, (const __half*)x, (const __half*)dz,
               (const float2*)stats, (const __half*)gamma, (const __half*)beta, part, HW, C, groups, splits, silu);
  }
  LAUNCH_CHECK("k_gn_bwd_stats");
  const int total = N * groups;
  launch_pdl(gdu::k_gn_bwd_finalize, dim3((total + 7) / 8), dim3(256), (size_t)0, s, (const float2*)part, bstats, total, splits,
             1.0f / ((float)HW * (float)(C / groups)));
  LAUNCH_CHECK("k_gn_bwd_finalize");
  int pix_per_cta = (32768 + C - 1) / C;
  const long long chunks = ((long long)HW + pix_per_cta - 1) / pix_per_cta;
  if (chunks > 65535) return fail(GD_UNET_ERR_INVALID_ARG, "groupnorm_bwd: image too large");
  if (256 % (C / 8) == 0) {
    pix_per_cta = gn_fast_pix_per_cta(N, HW, C, 2);
    const long long chunks = ((long long)HW + pix_per_cta - 1) / pix_per_cta;
    if (silu) launch_pdl(gdu::k_gn_bwd_apply_fast<true, 2>, dim3(N, (unsigned)chunks), dim3(256), (size_t)0, s, (const __half*)x, (const __half*)dz,
                         (const __half*)add, (__half*)dx, (const float2*)stats, (const float2*)bstats, (const __half*)gamma,
                         (const __half*)beta, HW, C, groups, pix_per_cta);
    else launch_pdl(gdu::k_gn_bwd_apply_fast<false, 2>, dim3(N, (unsigned)chunks), dim3(256), (size_t)0, s, (const __half*)x, (const __half*)dz,
                    (const __half*)add, (__half*)dx, (const float2*)stats, (const float2*)bstats, (const __half*)gamma,
                    (const __half*)beta, HW, C, groups, pix_per_cta);
    LAUNCH_CHECK("k_gn_bwd_apply_fast");
    return GD_UNET_OK;
  }
  launch_pdl(gdu::k_gn_bwd_apply, dim3(N, (unsigned)chunks), dim3(256), (size_t)((sizeof(float4) + sizeof(float2)) * C), s, (const __half*)x,
             (const __half*)dz, (const __half*)add, (__half*)dx, (const float2*)stats, (const float2*)bstats, (const __half*)gamma,
             (const __half*)beta, HW, C, groups, silu, pix_per_cta);
  LAUNCH_CHECK("k_gn_bwd_apply");
  return GD_UNET_OK;
}

int gd_unet_softmax_bwd(const void* P, void* dP, long long rows, int cols, long long ld, gd_ustream_t s) {
  if (cols % 8 || ld % 8) return fail(GD_UNET_ERR_INVALID_ARG, "softmax_bwd: cols and ld must be multiples of 8");
  launch_pdl(gdu::k_softmax_bwd, dim3((unsigned)((rows + 7) / 8)), dim3(256), (size_t)0, (cudaStream_t)s, (const __half*)P, (__half*)dP,
             rows, cols, ld);
  LAUNCH_CHECK("k_softmax_bwd");
  return GD_UNET_OK;
}
int gd_unet_transpose(const void* x, void* y, int B, int R, int C, gd_ustream_t s) {
  if (B < 1 || R < 1 || C < 1 || B > 65535) return fail(GD_UNET_ERR_INVALID_ARG, "transpose: bad shape");
  launch_pdl(gdu::k_transpose, dim3((C + 63) / 64, (R + 63) / 64, B), dim3(256), (size_t)0, (cudaStream_t)s, (const __half*)x, (__half*)y, R, C);
  LAUNCH_CHECK("k_transpose");
  return GD_UNET_OK;
}
int gd_unet_depth_to_space(const void* x, void* y, int N, int H, int W, int C, gd_ustream_t s) {
  if (C % 8 || H % 2 || W % 2) return fail(GD_UNET_ERR_INVALID_ARG, "depth_to_space: C % 8, even H/W");
  const long long total = (long long)N * H * W * (C / 8);
  launch_pdl(gdu::k_depth_to_space, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)0, (cudaStream_t)s, (const uint4*)x, (uint4*)y, N, H, W, C / 8);
  LAUNCH_CHECK("k_depth_to_space");
  return GD_UNET_OK;
}
int gd_vae_prep(const float* color, void* y, int B, int H, int W, float a, float shift, gd_ustream_t s) {
  const long long n = (long long)B * 4 * H * W;
  launch_pdl(gdu::k_vae_prep, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)0, (cudaStream_t)s, color, (__half*)y, B, (long long)H * W, a, shift);
  LAUNCH_CHECK("k_vae_prep");
  return GD_UNET_OK;
}
int gd_vae_im2col(const float* color, void* A, int B, int H, int W, float a, float shift, gd_ustream_t s) {
  const long long n = (long long)B * H * W * 8;
  launch_pdl(gdu::k_vae_im2col, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)0, (cudaStream_t)s, color, (__half*)A, B, H, W, a, shift);
  LAUNCH_CHECK("k_vae_im2col");
  return GD_UNET_OK;
}
int gd_vae_dimg_gather_dyn(const void* Z, float* dcolor, int B, int H, int W, float scale, const float* dyn, gd_ustream_t s) {
  const long long n = (long long)B * H * W;
  launch_pdl(gdu::k_vae_dimg_gather, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)0, (cudaStream_t)s, (const __half*)Z, dcolor, B, H, W, scale, dyn);
  LAUNCH_CHECK("k_vae_dimg_gather");
  return GD_UNET_OK;
}
int gd_vae_dimg_gather(const void* Z, float* dcolor, int B, int H, int W, float scale, gd_ustream_t s) {
  return gd_vae_dimg_gather_dyn(Z, dcolor, B, H, W, scale, nullptr, s);
}
int gd_vae_grad_scale(const float* grad, long long n, float clip, float pre, float target, float* scratch, float* dyn, gd_ustream_t s) {
  if (!grad || n < 1 || !scratch || !dyn || !(target > 0.f)) return fail(GD_UNET_ERR_INVALID_ARG, "vae_grad_scale: bad argument");
  const int nblk = (int)((n + 1023) / 1024);
  launch_pdl(gdu::k_vae_grad_absmax, dim3(nblk), dim3(256), (size_t)0, (cudaStream_t)s, grad, n, clip, scratch);
  LAUNCH_CHECK("k_vae_grad_absmax");
  launch_pdl(gdu::k_vae_grad_scale, dim3(1), dim3(256), (size_t)0, (cudaStream_t)s, nblk, (const float*)scratch, pre, target, dyn);
  LAUNCH_CHECK("k_vae_grad_scale");
  return GD_UNET_OK;
}
int gd_vae_sample(const void* moments, const float* noise, float* latents, int B, int hw, float scaling, gd_ustream_t s) {
  launch_pdl(gdu::k_vae_sample, dim3((B * 4 * hw + 255) / 256), dim3(256), (size_t)0, (cudaStream_t)s, (const __half*)moments, noise, latents, B, hw, scaling);
  LAUNCH_CHECK("k_vae_sample");
  return GD_UNET_OK;
}
int gd_vae_sample_bwd(const float* grad, const void* moments, const float* noise, void* dmoments, int B, int hw, int Cp,
                      float scaling, float clip, float gscale, gd_ustream_t s) {
  return gd_vae_sample_bwd_dyn(grad, moments, noise, dmoments, B, hw, Cp, scaling, clip, gscale, nullptr, s);
}
int gd_vae_sample_bwd_dyn(const float* grad, const void* moments, const float* noise, void* dmoments, int B, int hw, int Cp,
                          float scaling, float clip, float gscale, const float* dyn, gd_ustream_t s) {
  if (Cp < 8 || Cp % 8) return fail(GD_UNET_ERR_INVALID_ARG, "vae_sample_bwd: Cp must be a multiple of 8, >= 8");
  launch_pdl(gdu::k_vae_sample_bwd, dim3((B * hw + 255) / 256), dim3(256), (size_t)0, (cudaStream_t)s, grad, (const __half*)moments, noise,
             (__half*)dmoments, B, hw, Cp, scaling, clip, gscale, dyn);
  LAUNCH_CHECK("k_vae_sample_bwd");
  return GD_UNET_OK;
}
int gd_vae_dimg(const void* dx, float* dcolor, int B, int H, int W, int Cp, float scale, gd_ustream_t s) {
  if (Cp < 4 || Cp % 4) return fail(GD_UNET_ERR_INVALID_ARG, "vae_dimg: Cp must be a multiple of 4");
  const long long n = (long long)B * H * W;
  launch_pdl(gdu::k_vae_dimg, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)0, (cudaStream_t)s, (const __half*)dx, dcolor, B, (long long)H * W, Cp, scale);
  LAUNCH_CHECK("k_vae_dimg");
  return GD_UNET_OK;
}

int gd_resize_bilinear(const float* in, float* out, int BC, int Hi, int Wi, int Ho, int Wo, gd_ustream_t s) {
  if (!in || !out || BC < 1 || Hi < 1 || Wi < 1 || Ho < 1 || Wo < 1) return fail(GD_UNET_ERR_INVALID_ARG, "resize_bilinear: bad argument");
  const long long n = (long long)BC * Ho * Wo;
  launch_pdl(gdu::k_resize_bilinear, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)0, (cudaStream_t)s, in, out, BC, Hi, Wi, Ho, Wo);
  LAUNCH_CHECK("k_resize_bilinear");
  return GD_UNET_OK;
}
int gd_resize_bilinear_bwd(const float* dout, float* din, int BC, int Hi, int Wi, int Ho, int Wo, gd_ustream_t s) {
  if (!dout || !din || BC < 1 || Hi < 1 || Wi < 1 || Ho < 1 || Wo < 1) return fail(GD_UNET_ERR_INVALID_ARG, "resize_bilinear_bwd: bad argument");
  const long long n = (long long)BC * Hi * Wi;
  launch_pdl(gdu::k_resize_bilinear_bwd, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)0, (cudaStream_t)s, dout, din, BC, Hi, Wi, Ho, Wo);
  LAUNCH_CHECK("k_resize_bilinear_bwd");
  return GD_UNET_OK;
}

int gd_unet_sds_grad(const float* eps, const float* noise, const float* w, float gs, float* np, float* grad, int B, int chw,
                     gd_ustream_t s) {
  launch_pdl(gdu::k_sds_grad, dim3((B * chw + 255) / 256), dim3(256), (size_t)(0), (cudaStream_t)((cudaStream_t)s), eps, noise, w, gs, np, grad, B, chw);
  LAUNCH_CHECK("k_sds_grad");
  return GD_UNET_OK;
}

}  // extern "C"
