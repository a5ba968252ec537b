// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Thin extern "C" wrapper around the UNMODIFIED reference rasteriser core
// (CudaRasterizer::Rasterizer, /root/reference/.../diff-gaussian-rasterization/
// cuda_rasterizer/rasterizer.h:23-86). The reference sources are compiled where
// they lie by oracle/Makefile; nothing from them is copied into this repo. The
// shim replaces the torch binding (rasterize_points.cu:35-229), which needs ~9
// minutes of torch headers to compile, by plain cudaMalloc'ed state buffers.
//
// It exists so that (a) GPU parity tests can run our kernels against the real
// reference on the same inputs, (b) tests/golden fixtures can be generated, and
// (c) bench.py can time the reference CUDA rasteriser as the "GPU reference".
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <cuda_runtime.h>
#include "rasterizer.h"
#include "rasterizer_impl.h"

namespace {
struct Buf {
  char* p = nullptr;
  size_t cap = 0;
  size_t used = 0;
  char* resize(size_t n) {
    used = n;
    if (n > cap) {
      if (p) cudaFree(p);
      cap = n + n / 4 + 1024;
      if (cudaMalloc(&p, cap) != cudaSuccess) { p = nullptr; cap = 0; }
    }
    return p;
  }
  ~Buf() { if (p) cudaFree(p); }
};
struct Handle {
  Buf geom, binning, img;
  int P = 0, R = 0, W = 0, H = 0;
  char err[256] = {0};
};
}  // namespace

extern "C" {

struct GdRefState {
  // GeometryState (rasterizer_impl.h:33-47)
  float* depths; bool* clamped; int* internal_radii; float* means2D; float* cov3D;
  float* conic_opacity; float* rgb; uint32_t* point_offsets; uint32_t* tiles_touched;
  // BinningState (rasterizer_impl.h:57-67)
  uint64_t* point_list_keys_unsorted; uint64_t* point_list_keys;
  uint32_t* point_list_unsorted; uint32_t* point_list;
  // ImageState (rasterizer_impl.h:49-55)
  uint32_t* ranges; uint32_t* n_contrib;
  int P, R, W, H;
};

void* gdref_create() { return new Handle(); }
void gdref_destroy(void* h) { delete static_cast<Handle*>(h); }
const char* gdref_last_error(void* h) { return static_cast<Handle*>(h)->err; }

int gdref_forward(void* hv, int P, int D, int M, const float* bg, int W, int H,
                  const float* means3D, const float* shs, const float* colors_precomp,
                  const float* opacities, const float* scales, float scale_modifier,
                  const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                  const float* projmatrix, const float* campos, float tanfovx, float tanfovy,
                  int prefiltered, float* out_color, float* out_depth, float* out_alpha,
                  int* radii, int debug) {
  Handle* h = static_cast<Handle*>(hv);
  h->P = P; h->W = W; h->H = H; h->R = 0;
  if (P == 0) return 0;
  try {
    int R = CudaRasterizer::Rasterizer::forward(
        [h](size_t n) { return h->geom.resize(n); },
        [h](size_t n) { return h->binning.resize(n); },
        [h](size_t n) { return h->img.resize(n); },
        P, D, M, bg, W, H, means3D, shs, colors_precomp, opacities, scales, scale_modifier,
        rotations, cov3D_precomp, viewmatrix, projmatrix, campos, tanfovx, tanfovy,
        prefiltered != 0, out_color, out_depth, out_alpha, radii, debug != 0);
    h->R = R;
    return R;
  } catch (const std::exception& e) {
    snprintf(h->err, sizeof(h->err), "%s", e.what());
    return -1;
  }
}

int gdref_backward(void* hv, int P, int D, int M, int R, const float* bg, int W, int H,
                   const float* means3D, const float* shs, const float* colors_precomp,
                   const float* alphas, const float* scales, float scale_modifier,
                   const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                   const float* projmatrix, const float* campos, float tanfovx, float tanfovy,
                   const int* radii, const float* dL_dpix, const float* dL_dpix_depth,
                   const float* dL_dalphas, float* dL_dmean2D, float* dL_dconic,
                   float* dL_dopacity, float* dL_dcolor, float* dL_ddepth, float* dL_dmean3D,
                   float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot, int debug) {
  Handle* h = static_cast<Handle*>(hv);
  if (P == 0) return 0;
  try {
    CudaRasterizer::Rasterizer::backward(
        P, D, M, R, bg, W, H, means3D, shs, colors_precomp, alphas, scales, scale_modifier,
        rotations, cov3D_precomp, viewmatrix, projmatrix, campos, tanfovx, tanfovy, radii,
        h->geom.p, h->binning.p, h->img.p, dL_dpix, dL_dpix_depth, dL_dalphas, dL_dmean2D,
        dL_dconic, dL_dopacity, dL_dcolor, dL_ddepth, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale,
        dL_drot, debug != 0);
    return 0;
  } catch (const std::exception& e) {
    snprintf(h->err, sizeof(h->err), "%s", e.what());
    return -1;
  }
}

void gdref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix,
                        bool* present) {
  if (P) CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, present);
}

// Re-derives the pointers the reference carved out of its opaque byte buffers
// (GeometryState/BinningState/ImageState::fromChunk, rasterizer_impl.cu:155-193).
int gdref_get_state(void* hv, GdRefState* s) {
  Handle* h = static_cast<Handle*>(hv);
  memset(s, 0, sizeof(*s));
  s->P = h->P; s->R = h->R; s->W = h->W; s->H = h->H;
  if (!h->geom.p) return -1;
  char* c = h->geom.p;
  auto g = CudaRasterizer::GeometryState::fromChunk(c, h->P);
  s->depths = g.depths; s->clamped = g.clamped; s->internal_radii = g.internal_radii;
  s->means2D = reinterpret_cast<float*>(g.means2D); s->cov3D = g.cov3D;
  s->conic_opacity = reinterpret_cast<float*>(g.conic_opacity); s->rgb = g.rgb;
  s->point_offsets = g.point_offsets; s->tiles_touched = g.tiles_touched;
  if (h->binning.p) {
    c = h->binning.p;
    auto b = CudaRasterizer::BinningState::fromChunk(c, h->R);
    s->point_list_keys_unsorted = b.point_list_keys_unsorted;
    s->point_list_keys = b.point_list_keys;
    s->point_list_unsorted = b.point_list_unsorted;
    s->point_list = b.point_list;
  }
  if (h->img.p) {
    c = h->img.p;
    auto i = CudaRasterizer::ImageState::fromChunk(c, (size_t)h->W * h->H);
    s->ranges = reinterpret_cast<uint32_t*>(i.ranges);
    s->n_contrib = i.n_contrib;
  }
  return 0;
}

}  // extern "C"
