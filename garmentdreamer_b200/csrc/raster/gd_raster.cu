// Host side of the C ABI declared in include/gd_raster.h (unity build of the kernel files).
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "gd_raster_common.cuh"
#include "gd_raster_forward.cu"
#include "gd_raster_backward.cu"
#include "gd_params.cu"

namespace {
thread_local char g_err[512] = {0};
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, const char* a = "") {
  snprintf(g_err, sizeof(g_err), fmt, a);
  return code;
}
int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return GD_ERR_CUDA;
  }
  return GD_OK;
}
#define GD_LAUNCH_CHECK(what)                  \
  do {                                         \
    const int rc_ = check_launch(what);        \
    if (rc_ != GD_OK) return rc_;              \
  } while (0)

int make_views(const GdView* in, int B, int W, int H, gd::ViewPack& vp) {
  memset(&vp, 0, sizeof(vp));
  for (int b = 0; b < B; b++) {
    if (!in[b].viewmatrix || !in[b].projmatrix || !in[b].campos)
      return fail(GD_ERR_INVALID_ARG, "view %s has a null matrix pointer", "entry");
    vp.v[b].view = in[b].viewmatrix;
    vp.v[b].proj = in[b].projmatrix;
    vp.v[b].campos = in[b].campos;
    vp.v[b].tanfovx = in[b].tanfovx;
    vp.v[b].tanfovy = in[b].tanfovy;
    // reference: rasterizer_impl.cu:223-224
    vp.v[b].focal_y = H / (2.0f * in[b].tanfovy);
    vp.v[b].focal_x = W / (2.0f * in[b].tanfovx);
  }
  return GD_OK;
}
int check_dims(int P, int W, int H, int B) {
  if (P < 0 || W <= 0 || H <= 0) return fail(GD_ERR_INVALID_ARG, "bad P/W/H%s");
  if (B < 1 || B > GD_MAX_VIEWS) return fail(GD_ERR_INVALID_ARG, "B must be in 1..GD_MAX_VIEWS%s");
  if ((W + gd::kTile - 1) / gd::kTile > 1023 || (H + gd::kTile - 1) / gd::kTile > 1023)
    return fail(GD_ERR_INVALID_ARG, "image larger than 1023 tiles per side%s");
  return GD_OK;
}
}  // namespace

extern "C" {

const char* gd_last_error(void) { return g_err; }
uint64_t gd_launch_count(void) { return g_launches.load(); }
const char* gd_raster_version(void) { return "gd_raster 0.1 (sm_100a)"; }

int gd_raster_state_bytes(int P, int W, int H, int B, uint32_t max_rendered, size_t* geom_bytes,
                          size_t* binning_bytes, size_t* img_bytes) {
  const int rc = check_dims(P, W, H, B);
  if (rc != GD_OK) return rc;
  const gd::State s = gd::carve_state(P, W, H, B, max_rendered, nullptr, nullptr, nullptr);
  if (geom_bytes) *geom_bytes = s.geom_bytes;
  if (binning_bytes) *binning_bytes = s.binning_bytes;
  if (img_bytes) *img_bytes = s.img_bytes;
  return GD_OK;
}

int gd_raster_state_view(int P, int W, int H, int B, uint32_t max_rendered, void* geom_buffer,
                         void* binning_buffer, void* img_buffer, GdStateView* out) {
  const int rc = check_dims(P, W, H, B);
  if (rc != GD_OK) return rc;
  if (!out) return fail(GD_ERR_INVALID_ARG, "null out%s");
  const gd::State s = gd::carve_state(P, W, H, B, max_rendered, geom_buffer, binning_buffer, img_buffer);
  out->records = s.rec;
  out->tiles_touched = s.tiles_touched;
  out->point_offsets = s.point_offsets;
  out->cov3D = s.cov3D;
  out->clamped = s.clamped;
  out->counters = s.counters;
  out->point_list = s.point_list;
  out->tile_keys = s.tile_keys;
  out->sorted_records = s.sorted_rec;
  out->instance_slot = s.inst_slot;
  out->instance_grad = s.inst_grad;
  out->ranges = s.ranges;
  out->n_contrib = s.n_contrib;
  return GD_OK;
}

int gd_raster_forward(const GdFwdArgs* a, gd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a) return fail(GD_ERR_INVALID_ARG, "null args%s");
  int rc = check_dims(a->P, a->W, a->H, a->B);
  if (rc != GD_OK) return rc;
  if ((a->shs == nullptr) == (a->colors_precomp == nullptr))
    return fail(GD_ERR_INVALID_ARG, "provide exactly one of shs / colors_precomp%s");
  const bool has_sr = a->scales && a->rotations;
  if (has_sr == (a->cov3D_precomp != nullptr) || (!has_sr && (a->scales || a->rotations)))
    return fail(GD_ERR_INVALID_ARG, "provide exactly one of scales+rotations / cov3D_precomp%s");
  if (a->shs && (a->M < 1 || a->D < 0 || a->D > 3 || (a->D + 1) * (a->D + 1) > a->M))
    return fail(GD_ERR_INVALID_ARG, "SH degree / coefficient count mismatch%s");
  const int P = a->P, W = a->W, H = a->H, B = a->B;
  const int gx = (W + gd::kTile - 1) / gd::kTile, gy = (H + gd::kTile - 1) / gd::kTile;
  const int T = gx * gy;
  size_t gb, bb, ib;
  gd_raster_state_bytes(P, W, H, B, a->max_rendered, &gb, &bb, &ib);
  if (a->geom_bytes < gb || a->binning_bytes < bb || a->img_bytes < ib || !a->geom_buffer ||
      !a->binning_buffer || !a->img_buffer)
    return fail(GD_ERR_WORKSPACE_TOO_SMALL, "state buffer too small%s");
  const gd::State s = gd::carve_state(P, W, H, B, a->max_rendered, a->geom_buffer,
                                      a->binning_buffer, a->img_buffer);
  gd::ViewPack vp;
  rc = make_views(a->views, B, W, H, vp);
  if (rc != GD_OK) return rc;

  cudaMemsetAsync(s.tile_count, 0, sizeof(uint32_t) * (size_t)B * T, stream);
  const int nblkP = (P + gd::kBlk - 1) / gd::kBlk;
  if (P > 0) {
    gd::k_preprocess<<<nblkP, gd::kBlk, 0, stream>>>(
        P, a->D, a->M, B, W, H, gx, gy, a->means3D, a->scales, a->scale_modifier, a->rotations,
        a->opacities, a->shs, a->cov3D_precomp, a->colors_precomp, vp, a->radii, s.rec,
        s.tiles_touched, s.cov3D, s.clamped, s.scan_partials, s.tile_count);
    GD_LAUNCH_CHECK("k_preprocess");
  }
  gd::k_spine<<<2, 1024, 0, stream>>>(B * nblkP, s.scan_partials, B * T, T, B, s.tile_count,
                                      s.tile_cursor, s.ranges, s.seg_base, s.counters, a->max_rendered);
  GD_LAUNCH_CHECK("k_spine");
  if (P > 0) {
    gd::k_scatter<<<dim3(nblkP, B), gd::kBlk, 0, stream>>>(P, gx, T, s.tiles_touched,
                                                          s.scan_partials, s.point_offsets, s.rec,
                                                          s.tile_cursor, s.counters, s.tile_keys);
    GD_LAUNCH_CHECK("k_scatter");
    gd::k_tile_sort<<<B * T, gd::kTilePix, 0, stream>>>(P, gx, T, s.ranges, s.tile_keys, s.rec,
                                                        s.point_list, s.sorted_rec, s.inst_slot);
    GD_LAUNCH_CHECK("k_tile_sort");
  }
  gd::k_render_fwd<<<B * T, gd::kTilePix, 0, stream>>>(W, H, gx, T, s.ranges, s.sorted_rec,
                                                       a->background, a->out_color, a->out_depth,
                                                       a->out_alpha, s.n_contrib, s.fin, s.fin_T, s.seg_base,
                                                       s.ckpt, s.items, s.counters);
  GD_LAUNCH_CHECK("k_render_fwd");
  if (a->debug) {
    const cudaError_t e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return fail(GD_ERR_CUDA, "forward: %s", cudaGetErrorString(e));
  }
  return GD_OK;
}

int gd_raster_backward(const GdBwdArgs* a, gd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a) return fail(GD_ERR_INVALID_ARG, "null args%s");
  int rc = check_dims(a->P, a->W, a->H, a->B);
  if (rc != GD_OK) return rc;
  const int P = a->P, W = a->W, H = a->H, B = a->B;
  if (P == 0) return GD_OK;
  const int gx = (W + gd::kTile - 1) / gd::kTile, gy = (H + gd::kTile - 1) / gd::kTile;
  const int T = gx * gy;
  size_t gb, bb, ib;
  gd_raster_state_bytes(P, W, H, B, a->max_rendered, &gb, &bb, &ib);
  if (a->geom_bytes < gb || a->binning_bytes < bb || a->img_bytes < ib)
    return fail(GD_ERR_WORKSPACE_TOO_SMALL, "state buffer too small%s");
  if (a->shs && !a->dL_dsh) return fail(GD_ERR_INVALID_ARG, "dL_dsh missing%s");
  if (a->scales && (!a->dL_dscales || !a->dL_drotations || !a->rotations))
    return fail(GD_ERR_INVALID_ARG, "dL_dscales / dL_drotations missing%s");
  const gd::State s = gd::carve_state(P, W, H, B, a->max_rendered, a->geom_buffer,
                                      a->binning_buffer, a->img_buffer);
  gd::ViewPack vp;
  rc = make_views(a->views, B, W, H, vp);
  if (rc != GD_OK) return rc;
  // persistent CTAs pull (tile, segment) items from the queue the forward compositor filled
  static int bwd_grid = 0;
  if (!bwd_grid) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gd::k_render_bwd, gd::kTilePix, 0);
    bwd_grid = sms * (per_sm > 0 ? per_sm : 1);
  }
  cudaMemsetAsync(&s.counters->bwd_next, 0, sizeof(uint32_t), stream);
  gd::k_render_bwd<<<bwd_grid, gd::kTilePix, 0, stream>>>(
      W, H, gx, T, s.ranges, s.seg_base, s.items, s.counters, s.sorted_rec, a->background, a->out_alpha,
      s.n_contrib, s.fin, s.fin_T, s.ckpt, a->dL_dcolor, a->dL_ddepth, a->dL_dalpha, s.inst_grad);
  GD_LAUNCH_CHECK("k_render_bwd");
  gd::BwdOut out{a->dL_dmeans2D, a->dL_dcolors, a->dL_dopacity, a->dL_dmeans3D, a->dL_dcov3D,
                 a->shs ? a->dL_dsh : nullptr, a->scales ? a->dL_dscales : nullptr,
                 a->scales ? a->dL_drotations : nullptr, a->dL_dconic, a->dL_ddepths};
  const float* cov3D = a->cov3D_precomp ? a->cov3D_precomp : s.cov3D;
  gd::k_bwd_epilogue<<<(P + gd::kEpiBlk - 1) / gd::kEpiBlk, gd::kEpiBlk, 0, stream>>>(
      P, a->D, a->M, B, W, H, a->means3D, a->shs, a->scales, a->scale_modifier, a->rotations,
      cov3D, vp, a->radii, s.tiles_touched, s.point_offsets, s.clamped, s.inst_slot, s.inst_grad,
      s.counters, a->sum_views, out);
  GD_LAUNCH_CHECK("k_bwd_epilogue");
  if (a->debug) {
    const cudaError_t e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return fail(GD_ERR_CUDA, "backward: %s", cudaGetErrorString(e));
  }
  return GD_OK;
}

int gd_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                    uint8_t* present, gd_stream_t stream_) {
  (void)projmatrix;  // the reference computes p_hom and discards it (auxiliary.h:147-150)
  if (P < 0) return fail(GD_ERR_INVALID_ARG, "bad P%s");
  if (P == 0) return GD_OK;
  gd::k_mark_visible<<<(P + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      P, means3D, viewmatrix, present);
  GD_LAUNCH_CHECK("k_mark_visible");
  return GD_OK;
}

// ---- fused parameter kernels (include/gd_raster.h, "Gaussian parameters" section) --------------
int gd_params_activate(int P, const float* xyz, const float* f_dc, const float* opacity, const float* scaling,
                       const float* rotation, float* packed_out, gd_stream_t stream) {
  if (P < 0 || (P > 0 && (!xyz || !f_dc || !opacity || !scaling || !rotation || !packed_out)))
    return fail(GD_ERR_INVALID_ARG, "params_activate: null pointer%s");
  if (P == 0) return GD_OK;
  gd::k_params_activate<<<(P + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(P, xyz, f_dc, opacity, scaling, rotation, packed_out);
  GD_LAUNCH_CHECK("k_params_activate");
  return GD_OK;
}
int gd_params_adam(int P, float* xyz, float* f_dc, float* opacity, float* scaling, float* rotation, const float* packed_grad,
                   float* exp_avg, float* exp_avg_sq, const float* lr5, float beta1, float beta2, float eps, int step,
                   gd_stream_t stream) {
  if (P < 0 || step < 1 || !lr5 || (P > 0 && (!xyz || !f_dc || !opacity || !scaling || !rotation || !packed_grad || !exp_avg || !exp_avg_sq)))
    return fail(GD_ERR_INVALID_ARG, "params_adam: null pointer or step < 1%s");
  if (P == 0) return GD_OK;
  gd::AdamHyper h;
  for (int k = 0; k < 5; k++) h.lr[k] = lr5[k];
  h.beta1 = beta1; h.beta2 = beta2; h.eps = eps;
  h.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  h.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  gd::k_params_adam<<<(P + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(P, xyz, f_dc, opacity, scaling, rotation, packed_grad,
                                                                                         exp_avg, exp_avg_sq, h);
  GD_LAUNCH_CHECK("k_params_adam");
  return GD_OK;
}
int gd_densify_stats(int P, int B, const float* dmeans2D_sum, const int* radii, float* xyz_gradient_accum, float* denom,
                     float* max_radii2D, gd_stream_t stream) {
  if (P < 0 || B < 1 || (P > 0 && (!dmeans2D_sum || !radii || !xyz_gradient_accum || !denom || !max_radii2D)))
    return fail(GD_ERR_INVALID_ARG, "densify_stats: null pointer%s");
  if (P == 0) return GD_OK;
  gd::k_densify_stats<<<(P + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(P, B, dmeans2D_sum, radii, xyz_gradient_accum, denom, max_radii2D);
  GD_LAUNCH_CHECK("k_densify_stats");
  return GD_OK;
}

static int check_peer_table(const GdPeerTable* t, const char* what) {
  if (!t || t->world < 1 || t->world > GD_MAX_PEERS || t->rank < 0 || t->rank >= t->world) return fail(GD_ERR_INVALID_ARG, what);
  for (int w = 0; w < t->world; w++)
    if (!t->grad[w] || !t->radii[w] || !t->red_grad[w] || !t->red_radii[w] || !t->flags[w]) return fail(GD_ERR_INVALID_ARG, what);
  if ((t->mc_grad == nullptr) != (t->mc_red_grad == nullptr)) return fail(GD_ERR_INVALID_ARG, what);
  return GD_OK;
}
int gd_peer_allreduce(int P, const GdPeerTable* t, unsigned epoch, unsigned* counter, gd_stream_t stream) {
  if (P < 1 || !counter || epoch == 0) return fail(GD_ERR_INVALID_ARG, "peer_allreduce: P >= 1, counter and a non-zero epoch required%s");
  const int rc = check_peer_table(t, "peer_allreduce: bad peer table%s");
  if (rc != GD_OK) return rc;
  const long long g4_total = (17LL * P + 3) / 4, r4_total = ((long long)P + 3) / 4;
  const long long g4 = (g4_total + t->world - 1) / t->world, r4 = (r4_total + t->world - 1) / t->world;
  const long long blocks = (g4 + r4 + 255) / 256;
  gd::k_peer_allreduce<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*t, g4, g4_total, r4, r4_total, epoch, counter);
  GD_LAUNCH_CHECK("k_peer_allreduce");
  return GD_OK;
}
int gd_params_adam_peers(int P, float* xyz, float* f_dc, float* opacity, float* scaling, float* rotation, const GdPeerTable* t,
                         unsigned epoch, float* exp_avg, float* exp_avg_sq, const float* lr5, float beta1, float beta2, float eps,
                         int step, int densify, float* xyz_gradient_accum, float* denom, float* max_radii2D, gd_stream_t stream) {
  if (P < 1 || step < 1 || !lr5 || !xyz || !f_dc || !opacity || !scaling || !rotation || !exp_avg || !exp_avg_sq || epoch == 0)
    return fail(GD_ERR_INVALID_ARG, "params_adam_peers: null pointer, step < 1 or epoch 0%s");
  if (densify && (!xyz_gradient_accum || !denom || !max_radii2D)) return fail(GD_ERR_INVALID_ARG, "params_adam_peers: statistics buffers required%s");
  const int rc = check_peer_table(t, "params_adam_peers: bad peer table%s");
  if (rc != GD_OK) return rc;
  gd::AdamHyper h;
  for (int k = 0; k < 5; k++) h.lr[k] = lr5[k];
  h.beta1 = beta1; h.beta2 = beta2; h.eps = eps;
  h.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  h.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  gd::k_params_adam_peers<<<(P + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      P, xyz, f_dc, opacity, scaling, rotation, *t, epoch, exp_avg, exp_avg_sq, h, densify, xyz_gradient_accum, denom, max_radii2D);
  GD_LAUNCH_CHECK("k_params_adam_peers");
  return GD_OK;
}

int gd_sparsity_grad(long long n, long long n_total, const float* depth, const float* depth_max, float lambda,
                     float* dL_ddepth, float* scratch, float* stats, gd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (n < 1 || n_total < n || !depth || !depth_max || !dL_ddepth || !scratch || !stats)
    return fail(GD_ERR_INVALID_ARG, "sparsity_grad: null pointer or bad element count%s");
  const int nblk = (int)((n + 1023) / 1024);
  gd::k_sparsity_grad<<<nblk, 256, 0, stream>>>(n, 1.0f / (float)n_total, depth, depth_max, lambda, dL_ddepth, scratch);
  GD_LAUNCH_CHECK("k_sparsity_grad");
  gd::k_sparsity_sum<<<1, 256, 0, stream>>>(nblk, scratch, stats);
  GD_LAUNCH_CHECK("k_sparsity_sum");
  return GD_OK;
}
int gd_sparsity_finish(long long n, long long n_total, const float* depth, const float* depth_max, float lambda,
                       const float* stats, float* dL_ddepth, float* loss_out, gd_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (n < 1 || n_total < n || !depth || !depth_max || !dL_ddepth || !stats)
    return fail(GD_ERR_INVALID_ARG, "sparsity_finish: null pointer or bad element count%s");
  gd::k_sparsity_finish<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, 1.0f / (float)n_total, depth, depth_max, lambda, stats,
                                                                        dL_ddepth, loss_out);
  GD_LAUNCH_CHECK("k_sparsity_finish");
  return GD_OK;
}
int gd_radii_max(int P, int B, const int* radii, int* out, gd_stream_t stream) {
  if (P < 0 || B < 1 || (P > 0 && (!radii || !out))) return fail(GD_ERR_INVALID_ARG, "radii_max: null pointer%s");
  if (P == 0) return GD_OK;
  gd::k_radii_max<<<(P + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(P, B, radii, out);
  GD_LAUNCH_CHECK("k_radii_max");
  return GD_OK;
}

int gd_cameras_from_c2w(int B, const float* c2w, const float* tan_half_fovx, const float* tan_half_fovy, float znear, float zfar,
                        float* out35, gd_stream_t stream) {
  if (B < 1 || B > GD_MAX_VIEWS) return fail(GD_ERR_INVALID_ARG, "cameras_from_c2w: B must be in 1..GD_MAX_VIEWS%s");
  if (!c2w || !tan_half_fovx || !tan_half_fovy || !out35) return fail(GD_ERR_INVALID_ARG, "cameras_from_c2w: null pointer%s");
  gd::CamIntrinsics in;
  for (int b = 0; b < B; b++) { in.tan_half_fovx[b] = tan_half_fovx[b]; in.tan_half_fovy[b] = tan_half_fovy[b]; }
  in.znear = znear; in.zfar = zfar;
  gd::k_cameras_from_c2w<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(B, c2w, in, out35);
  GD_LAUNCH_CHECK("k_cameras_from_c2w");
  return GD_OK;
}

}  // extern "C"
