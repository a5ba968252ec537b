/*
 * gd_raster.h -- C ABI of the B200-native differentiable Gaussian-splat rasteriser.
 *
 * Drop-in boundary for the reference's native module `diff_gaussian_rasterization._C`
 * (reference: Garment_3DGS/gaussiansplatting/submodules/diff-gaussian-rasterization, "DGR/"):
 *
 *   gd_raster_forward   replaces  RasterizeGaussiansCUDA          DGR/rasterize_points.cu:35-119
 *                                 -> CudaRasterizer::Rasterizer::forward
 *                                                                 DGR/cuda_rasterizer/rasterizer_impl.cu:197-339
 *   gd_raster_backward  replaces  RasterizeGaussiansBackwardCUDA  DGR/rasterize_points.cu:121-208
 *                                 -> CudaRasterizer::Rasterizer::backward
 *                                                                 DGR/cuda_rasterizer/rasterizer_impl.cu:343-447
 *   gd_mark_visible     replaces  markVisible                     DGR/rasterize_points.cu:210-229
 *   gd_raster_state_bytes / gd_raster_state_view replace the resize callbacks and the
 *   GeometryState/BinningState/ImageState::fromChunk carving
 *                                                                 DGR/cuda_rasterizer/rasterizer_impl.cu:155-193
 *                                                                 DGR/rasterize_points.cu:27-33
 *
 * Differences from the reference boundary, all deliberate:
 *   - plain pointers and sizes only (no torch types); every pointer is a DEVICE pointer unless
 *     the field name ends in _host;
 *   - B >= 1 camera views are rasterised by ONE call (the reference is called once per view,
 *     Garment_3DGS/threestudio/systems/GaussianDreamer.py:189-191); B = 1 reproduces it;
 *   - the caller owns every buffer; the library never allocates and never synchronises with the
 *     host: the instance count stays on the device (the reference does a blocking cudaMemcpy,
 *     rasterizer_impl.cu:282). The caller sizes the instance arena (`max_rendered`) up front and
 *     reads back `counters` when it wants the count / the overflow flag;
 *   - an explicit cudaStream_t (the reference launches on the legacy default stream).
 *
 * Conventions (identical to the reference): viewmatrix/projmatrix are the TRANSPOSED 4x4
 * matrices, read as matrix[c*4+r]; quaternions are (r,x,y,z) and are not normalised; tiles are
 * 16x16; all floating point data is fp32; images are CHW.
 */
#ifndef GD_RASTER_H_
#define GD_RASTER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GD_MAX_VIEWS 32
#define GD_TILE 16

/* Return codes. Negative = error; gd_last_error() holds a message for the calling thread. */
#define GD_OK 0
#define GD_ERR_INVALID_ARG (-1)
#define GD_ERR_WORKSPACE_TOO_SMALL (-2)
#define GD_ERR_CUDA (-3)
#define GD_ERR_NON_RGB (-4) /* reference: "For non-RGB, provide precomputed Gaussian colors!" */

typedef struct CUstream_st* gd_stream_t; /* == cudaStream_t */

/* One camera view (reference: GaussianRasterizationSettings, DGR/diff_gaussian_rasterization/__init__.py:160-172). */
typedef struct {
  const float* viewmatrix; /* [16] device */
  const float* projmatrix; /* [16] device */
  const float* campos;     /* [3]  device */
  float tanfovx, tanfovy;
} GdView;

/* Device-side counters written by gd_raster_forward (copy back when needed). */
typedef struct {
  uint32_t num_rendered; /* total (Gaussian, tile) instances over all B views */
  uint32_t overflow;     /* 1 if num_rendered > max_rendered: outputs are blank, re-run bigger */
  uint32_t view_base[GD_MAX_VIEWS + 1]; /* first instance of each view in the global list */
  uint32_t bwd_items;    /* backward work items (tile, list segment) queued by the forward compositor */
  uint32_t bwd_next;     /* next item to hand out; reset by every gd_raster_backward call */
  uint32_t depth_max_bits; /* IEEE bits of max over the B depth images (>= 0): `depths.max()` of
                              TS/systems/GaussianDreamer.py:215 without another pass over the images */
} GdCounters;

typedef struct {
  int P, D, M;  /* Gaussians, active SH degree, SH coefficients per Gaussian */
  int W, H, B;  /* image width, height, number of views (1..GD_MAX_VIEWS) */
  const float* background;     /* [3] */
  const float* means3D;        /* [P,3] */
  const float* shs;            /* [P,M,3] or NULL */
  const float* colors_precomp; /* [P,3] or NULL (exactly one of shs / colors_precomp) */
  const float* opacities;      /* [P] */
  const float* scales;         /* [P,3] or NULL */
  float scale_modifier;
  const float* rotations;      /* [P,4] or NULL */
  const float* cov3D_precomp;  /* [P,6] or NULL (exactly one of scales+rotations / cov3D) */
  GdView views[GD_MAX_VIEWS];
  int prefiltered; /* accepted for API compatibility; culled points are skipped either way */
  int debug;       /* non-zero: synchronise after the call and report CUDA errors (CHECK_CUDA) */
  /* outputs */
  float* out_color; /* [B,3,H,W] */
  float* out_depth; /* [B,1,H,W] */
  float* out_alpha; /* [B,1,H,W] */
  int* radii;       /* [B,P] */
  /* caller-owned state, sizes from gd_raster_state_bytes; must survive until backward */
  void* geom_buffer;    size_t geom_bytes;
  void* binning_buffer; size_t binning_bytes;
  void* img_buffer;     size_t img_bytes;
  uint32_t max_rendered; /* capacity of the instance arena the binning buffer was sized for */
} GdFwdArgs;

typedef struct {
  int P, D, M, W, H, B;
  const float* background;
  const float* means3D;
  const float* shs;
  const float* colors_precomp;
  const float* scales;
  float scale_modifier;
  const float* rotations;
  const float* cov3D_precomp;
  GdView views[GD_MAX_VIEWS];
  const int* radii;         /* [B,P] from forward */
  const float* out_alpha;   /* [B,1,H,W] from forward */
  const float* dL_dcolor;   /* [B,3,H,W] */
  const float* dL_ddepth;   /* [B,1,H,W] */
  const float* dL_dalpha;   /* [B,1,H,W] */
  int debug;
  /*
   * Gradient outputs. Every element is WRITTEN (no zero-initialisation needed; Gaussians that
   * were invisible in a view get exact zeros, like the reference's torch::zeros + skipped
   * threads). sum_views = 0: per-view gradients, leading dimension B (B = 1 is the reference
   * layout). sum_views = 1: gradients summed over the B views in view order (what autograd
   * accumulates over the reference's per-view loop), leading dimension dropped.
   */
  int sum_views;
  float* dL_dmeans2D;   /* [B,P,3]  (grad wrt NDC position; z = 0) */
  float* dL_dcolors;    /* [B,P,3]  */
  float* dL_dopacity;   /* [B,P]    */
  float* dL_dmeans3D;   /* [B,P,3]  */
  float* dL_dcov3D;     /* [B,P,6]  */
  float* dL_dsh;        /* [B,P,M,3] or NULL when shs == NULL */
  float* dL_dscales;    /* [B,P,3]  or NULL when scales == NULL */
  float* dL_drotations; /* [B,P,4]  or NULL when scales == NULL */
  float* dL_dconic;     /* [B,P,4] optional (NULL to skip): slots x,y,w used, z = 0 */
  float* dL_ddepths;    /* [B,P]   optional (NULL to skip) */
  void* geom_buffer;    size_t geom_bytes;
  void* binning_buffer; size_t binning_bytes;
  void* img_buffer;     size_t img_bytes;
  uint32_t max_rendered;
} GdBwdArgs;

/* Read-only device pointers into the state buffers (parity tests, debugging). */
typedef struct {
  const float* records;           /* [B*P,12]: conic.xyz, opacity | px, py, depth, r | g, b, -, - */
  const uint32_t* tiles_touched;  /* [B*P] */
  const uint32_t* point_offsets;  /* [B*P] inclusive scan over the view-major concatenation */
  const float* cov3D;             /* [P,6] */
  const uint8_t* clamped;         /* [B*P] bit k = colour channel k was clamped */
  const GdCounters* counters;
  const uint32_t* point_list;     /* [num_rendered] Gaussian index, sorted by (view, tile, depth, index) */
  const uint64_t* tile_keys;      /* [num_rendered] (depth_bits << 32 | index), sorted per tile */
  const float* sorted_records;    /* [num_rendered,12] records gathered in sorted order */
  const uint32_t* instance_slot;  /* [num_rendered] unsorted instance -> sorted position */
  const float* instance_grad;     /* [num_rendered,12] written by backward */
  const uint32_t* ranges;         /* [B*T,2] global [start,end) per tile; (0,0) if empty */
  const uint32_t* n_contrib;      /* [B*H*W] */
} GdStateView;

/* Bytes needed for the three state buffers. */
int gd_raster_state_bytes(int P, int W, int H, int B, uint32_t max_rendered, size_t* geom_bytes,
                          size_t* binning_bytes, size_t* img_bytes);
int gd_raster_state_view(int P, int W, int H, int B, uint32_t max_rendered, void* geom_buffer,
                         void* binning_buffer, void* img_buffer, GdStateView* out);

/* Asynchronous on `stream`; returns GD_OK or a negative error code. */
int gd_raster_forward(const GdFwdArgs* args, gd_stream_t stream);
int gd_raster_backward(const GdBwdArgs* args, gd_stream_t stream);
/* present[i] = (view-space z of means3D[i]) > 0.2 */
int gd_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                    uint8_t* present, gd_stream_t stream);

/* ---- Gaussian parameters: fused activation / optimiser / statistics kernels ----
 * Packed struct-of-arrays layout of the ACTIVATED parameters and of their gradient, 14 floats per
 * Gaussian in one buffer:  xyz 3P | f_dc 3P | opacity P | scales 3P | rotation 4P.
 * gd_params_activate replaces get_xyz/get_features/get_opacity/get_scaling/get_rotation
 *   (GS/scene/gaussian_model.py:95-115: identity, identity, sigmoid, exp, F.normalize), which the
 *   reference re-evaluates per view (GS/gaussian_renderer/__init__.py:53-80).
 * gd_params_adam replaces autograd through those activations + torch.optim.Adam(eps=1e-15) over the
 *   parameter groups of gaussian_model.py:156-165 (lr5 = xyz, f_dc, opacity, scaling, rotation;
 *   step = 1-based optimiser step): RAW parameters and exp_avg / exp_avg_sq [14P] updated in place.
 * gd_densify_stats replaces add_densification_stats (:415-419) and the max_radii2D update of
 *   TS/systems/GaussianDreamer.py:269-275 for a batch of B views (radii i32 [B,P]). */
int gd_params_activate(int P, const float* xyz, const float* f_dc, const float* opacity, const float* scaling,
                       const float* rotation, float* packed_out, gd_stream_t stream);
int gd_params_adam(int P, float* xyz, float* f_dc, float* opacity, float* scaling, float* rotation,
                   const float* packed_grad, float* exp_avg, float* exp_avg_sq, const float* lr5,
                   float beta1, float beta2, float eps, int step, gd_stream_t stream);
int gd_densify_stats(int P, int B, const float* dmeans2D_sum, const int* radii, float* xyz_gradient_accum,
                     float* denom, float* max_radii2D, gd_stream_t stream);

/* Sparsity loss on the depth-normalised opacity and its gradient w.r.t. the depth images:
 *   opacity = depths / (depths.max() + 1e-5)                  TS/systems/GaussianDreamer.py:215
 *   loss    = lambda * mean(sqrt(opacity^2 + 0.01))            :253-255 (mean over n_total elements)
 * depth: DEVICE fp32 [n] (all B views of this rank, n = B*H*W); depth_max: DEVICE fp32 scalar holding the
 * max over the WHOLE view batch (GdCounters.depth_max_bits reinterpreted, after an all-reduce MAX when the
 * views are sharded over ranks); n_total = elements of the whole batch (all ranks).
 * gd_sparsity_grad writes dL_ddepth [n] WITHOUT the term through depths.max() and the three partial sums
 *   stats[0] = sum sqrt(op^2+0.01), stats[1] = sum_pix g_pix * depth_pix, stats[2] = #pixels == max
 * (fixed summation order; sharded views: all-reduce SUM stats before the next call).
 * gd_sparsity_finish adds the max() term, -stats[1] / (max+1e-5)^2 / stats[2] (autograd spreads it evenly
 * over ties), to the arg-max pixels and writes loss_out[0] = lambda * stats[0] / n_total.
 * scratch: DEVICE fp32 [3 * ceil(n / 1024)]. */
int gd_sparsity_grad(long long n, long long n_total, const float* depth, const float* depth_max, float lambda,
                     float* dL_ddepth, float* scratch, float* stats, gd_stream_t stream);
int gd_sparsity_finish(long long n, long long n_total, const float* depth, const float* depth_max, float lambda,
                       const float* stats, float* dL_ddepth, float* loss_out, gd_stream_t stream);
/* out[i] = max over the B views of radii[b][i] (visibility / max_radii2D bookkeeping of a view batch) */
int gd_radii_max(int P, int B, const int* radii, int* out, gd_stream_t stream);

/* ---- Multi-GPU gradient exchange over peer memory, fused with the optimiser step (SURVEY.md s.8 row e) ----
 * The reference is single-GPU; here the view batch is sharded over W <= GD_MAX_PEERS ranks of one NVLink domain and the
 * per-Gaussian results of every rank's rasteriser backward are summed before Adam (what autograd's accumulation over the
 * per-view loop of TS/systems/GaussianDreamer.py:189-219 does inside one process). Every pointer in GdPeerTable is a
 * DEVICE address valid on THIS rank: entry w points into rank w's symmetric allocation (peer-mapped), entry `rank` is local.
 *   grad[w]      fp32 [17P padded to 4]: 14P packed parameter gradients | 3P viewspace gradients (gd_raster_backward output)
 *   radii[w]     i32  [P padded to 4]:   per-Gaussian maximum screen radius over that rank's views (gd_radii_max)
 *   red_grad[w] / red_radii[w]           same shapes; after gd_peer_allreduce on ALL ranks they hold SUM / MAX over ranks
 *   flags[w]     u32  [2][GD_MAX_PEERS]  arrival epochs, zero-initialised; epoch must increase by one per call pair
 *   mc_grad / mc_red_grad                NVLS multicast addresses of grad / red_grad (NULL: peer loads and stores instead)
 * gd_peer_allreduce: cross-rank barrier (entry), reduce-scatter + all-gather of both arrays. counter: LOCAL u32, zero.
 * gd_params_adam_peers: waits for every rank's slice, then gd_densify_stats (if densify; B = 1 on red_radii) and
 *   gd_params_adam on red_grad[rank] in one kernel. Both are asynchronous on `stream`; no host synchronisation. */
#define GD_MAX_PEERS 8
typedef struct {
  int world, rank;
  const float* grad[GD_MAX_PEERS];
  const int* radii[GD_MAX_PEERS];
  float* red_grad[GD_MAX_PEERS];
  int* red_radii[GD_MAX_PEERS];
  unsigned* flags[GD_MAX_PEERS];
  const float* mc_grad;
  float* mc_red_grad;
} GdPeerTable;
int gd_peer_allreduce(int P, const GdPeerTable* t, unsigned epoch, unsigned* counter, gd_stream_t stream);
int gd_params_adam_peers(int P, float* xyz, float* f_dc, float* opacity, float* scaling, float* rotation,
                         const GdPeerTable* t, unsigned epoch, float* exp_avg, float* exp_avg_sq, const float* lr5,
                         float beta1, float beta2, float eps, int step, int densify, float* xyz_gradient_accum,
                         float* denom, float* max_radii2D, gd_stream_t stream);

/* Batched camera construction: replaces Camera.__init__ (GS/scene/cameras.py:50-53,
 * GS/utils/graphics_utils.py:59-101), which the reference runs on the CPU per view per iteration.
 * c2w: DEVICE fp32 [B,4,4] row-major camera-to-world (batch['c2w_3dgs']); tan_half_fovx/y: HOST
 * arrays [B] (the caller needs them on the host anyway for GdView); out35: DEVICE fp32 [B,35] =
 * world_view_transform 16 | full_proj_transform 16 | camera_center 3, in the transposed
 * (row-vector) convention GdView expects. */
int gd_cameras_from_c2w(int B, const float* c2w, const float* tan_half_fovx, const float* tan_half_fovy,
                        float znear, float zfar, float* out35, gd_stream_t stream);

const char* gd_last_error(void);
/* Number of kernels this library has launched so far in this process (bench.py gpu_launches). */
uint64_t gd_launch_count(void);
const char* gd_raster_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GD_RASTER_H_ */
