/*
 * TEST INFRASTRUCTURE -- CPU restatement (plain C) of the reference differentiable
 * Gaussian-splat rasteriser. Never linked into, imported by, or called from the product
 * library (garmentdreamer_b200/csrc). Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it.
 *
 * Reference: /root/reference/Garment_3DGS/gaussiansplatting/submodules/
 *            diff-gaussian-rasterization  (abbreviated DGR/ below)
 *
 * PARITY PIN: the reference ships no tests or golden vectors for this path.
 * This oracle is pinned against outputs of the UNMODIFIED reference CUDA core compiled by
 * oracle/Makefile into oracle/_ref/ and executed on a B200 (tests/golden/make_golden.py wrote
 * tests/golden/*.npz; tests/test_oracle_golden.py checks them on CPU).
 *
 * Floating-point evaluation order. The index path (radii, tile rectangles, depths -> keys ->
 * point_list, ranges) is only bit-exact if every float operation is rounded exactly like the
 * reference's sm_100a binary. nvcc/ptxas contract mul+add into fma in a fixed pattern; that
 * pattern was read from the SASS of DGR/cuda_rasterizer/forward.cu:preprocessCUDA compiled with
 * the flags in oracle/Makefile and is restated here with explicit fmaf(). This file must be
 * compiled with -ffp-contract=off so that the C compiler adds no contraction of its own.
 * Terms that the reference multiplies by a literal 0 (the zero entries of the GLM matrices S and
 * J, forward.cu:89-92,121-124) are dropped: for finite inputs fma(x, 0, y) == y.
 * Division, reciprocal and sqrt are IEEE round-to-nearest on both sides (nvcc default
 * -prec-div=true -prec-sqrt=true). expf is NOT bit-reproducible (the GPU uses MUFU.EX2), so
 * alpha values, colours and n_contrib are compared with a tolerance / mismatch budget.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BLOCK_X 16 /* DGR/cuda_rasterizer/config.h:16-17 */
#define BLOCK_Y 16
#define BLOCK_SIZE (BLOCK_X * BLOCK_Y)

/* DGR/cuda_rasterizer/auxiliary.h:22-39 */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* CUDA float->int conversion saturates and maps NaN to 0 (F2I.TRUNC.NTZ). */
static inline int f2i_trunc(float f) {
  if (!(f == f)) return 0;
  if (f >= 2147483648.0f) return 2147483647;
  if (f <= -2147483648.0f) return (-2147483647 - 1);
  return (int)f;
}

/* auxiliary.h:41-44 -- evaluated in double because of the 1.0 / 0.5 literals; SASS: DADD, DFMA, DMUL. */
static inline float ndc2pix(float v, int S) {
  return (float)(fma((double)v + 1.0, (double)S, -1.0) * 0.5);
}

/* auxiliary.h:46-56. rect = [min, max) in tile units, clamped to the grid. */
static inline void get_rect(float px, float py, int max_radius, int gx, int gy, int* rmin_x,
                            int* rmin_y, int* rmax_x, int* rmax_y) {
  const float r = (float)max_radius;
  *rmin_x = imin(gx, imax(0, f2i_trunc((px - r) * 0.0625f)));
  *rmin_y = imin(gy, imax(0, f2i_trunc((py - r) * 0.0625f)));
  *rmax_x = imin(gx, imax(0, f2i_trunc((((px + r) + 16.0f) - 1.0f) * 0.0625f)));
  *rmax_y = imin(gy, imax(0, f2i_trunc((((py + r) + 16.0f) - 1.0f) * 0.0625f)));
}

/* matrix[c*4+r] column-major 4x4 (auxiliary.h:58-77): row r of M*[x,y,z,1]. */
static inline float xform_row(const float* m, int r, float x, float y, float z) {
  return m[12 + r] + fmaf(z, m[8 + r], fmaf(x, m[r], y * m[4 + r]));
}

/* forward.cu:118-152 computeCov3D; quaternion (r,x,y,z) is NOT normalised (forward.cu:127). */
static void cov3d_from_scale_rot(const float* scale, float mod, const float* q, float* c) {
  const float sx = mod * scale[0], sy = mod * scale[1], sz = mod * scale[2];
  const float r = q[0], x = q[1], y = q[2], z = q[3];
  const float yy = y * y, zz = z * z;
  const float rz = r * z, xz = x * z, rx = r * x;
  const float R00 = 1.0f - ((yy + zz) + (yy + zz));
  const float xy_m_rz = fmaf(x, y, -rz), xy_p_rz = fmaf(x, y, rz);
  const float ry_p_xz = fmaf(r, y, xz), xz_m_ry = fmaf(-r, y, xz);
  const float yz_m_rx = fmaf(y, z, -rx), yz_p_rx = fmaf(y, z, rx);
  const float xx_zz = fmaf(x, x, zz), xx_yy = fmaf(x, x, yy);
  const float R01 = xy_m_rz + xy_m_rz, R02 = ry_p_xz + ry_p_xz;
  const float R10 = xy_p_rz + xy_p_rz, R11 = 1.0f - (xx_zz + xx_zz), R12 = yz_m_rx + yz_m_rx;
  const float R20 = xz_m_ry + xz_m_ry, R21 = yz_p_rx + yz_p_rx, R22 = 1.0f - (xx_yy + xx_yy);
  /* M = S * R (GLM column-major): M[c][r] = s_r * R[c][r] */
  const float M00 = sx * R00, M01 = sy * R01, M02 = sz * R02;
  const float M10 = sx * R10, M11 = sy * R11, M12 = sz * R12;
  const float M20 = sx * R20, M21 = sy * R21, M22 = sz * R22;
  /* Sigma = M^T M, products accumulate as fma(third, fma(first, second-product)) */
  c[0] = fmaf(M02, M02, fmaf(M00, M00, M01 * M01));
  c[1] = fmaf(M12, M02, fmaf(M10, M00, M11 * M01));
  c[2] = fmaf(M22, M02, fmaf(M20, M00, M21 * M01));
  c[3] = fmaf(M12, M12, fmaf(M10, M10, M11 * M11));
  c[4] = fmaf(M22, M12, fmaf(M20, M10, M21 * M11));
  c[5] = fmaf(M22, M22, fmaf(M20, M20, M21 * M21));
}

/* The two non-zero rows of T = W*J (forward.cu:74-113): t0 = T[0], t1 = T[1] in GLM terms. */
static void cov2d_T(float x, float y, float z, float focal_x, float focal_y, float tanfovx,
                    float tanfovy, const float* v, float* t0, float* t1, float* t_out,
                    float* txtz_out, float* tytz_out) {
  const float tx = xform_row(v, 0, x, y, z), ty = xform_row(v, 1, x, y, z),
              tz = xform_row(v, 2, x, y, z);
  const float limx = tanfovx * 1.3f, limy = tanfovy * 1.3f;
  const float txtz = tx / tz, tytz = ty / tz;
  const float cx = fminf(limx, fmaxf(-limx, txtz));
  const float cy = fminf(limy, fmaxf(-limy, tytz));
  const float tz2 = tz * tz;
  const float J00 = focal_x / tz;
  const float J02 = (focal_x * (cx * (-tz))) / tz2;
  const float J11 = focal_y / tz;
  const float J12 = (focal_y * (cy * (-tz))) / tz2;
  t0[0] = fmaf(v[2], J02, v[0] * J00);
  t0[1] = fmaf(v[6], J02, v[4] * J00);
  t0[2] = fmaf(v[10], J02, v[8] * J00);
  t1[0] = fmaf(v[2], J12, J11 * v[1]);
  t1[1] = fmaf(v[6], J12, J11 * v[5]);
  t1[2] = fmaf(v[10], J12, J11 * v[9]);
  if (t_out) { t_out[0] = cx * tz; t_out[1] = cy * tz; t_out[2] = tz; }
  if (txtz_out) *txtz_out = txtz;
  if (tytz_out) *tytz_out = tytz;
}

static void cov2d_from_T(const float* t0, const float* t1, const float* c, float* a, float* b,
                         float* cc) {
  const float A0 = fmaf(t0[2], c[2], fmaf(t0[0], c[0], t0[1] * c[1]));
  const float B0 = fmaf(t1[2], c[2], fmaf(t1[0], c[0], t1[1] * c[1]));
  const float A1 = fmaf(t0[2], c[4], fmaf(t0[0], c[1], t0[1] * c[3]));
  const float B1 = fmaf(t1[2], c[4], fmaf(t1[0], c[1], t1[1] * c[3]));
  const float A2 = fmaf(t0[2], c[5], fmaf(t0[0], c[2], t0[1] * c[4]));
  const float B2 = fmaf(t1[2], c[5], fmaf(t1[0], c[2], t1[1] * c[4]));
  *a = fmaf(t0[2], A2, fmaf(t0[0], A0, t0[1] * A1)) + 0.3f;
  *b = fmaf(t0[2], B2, fmaf(t0[0], B0, t0[1] * B1));
  *cc = fmaf(t1[2], B2, fmaf(t1[0], B0, t1[1] * B1)) + 0.3f;
}

/* forward.cu:20-71 computeColorFromSH. Returns pre-clamp colour + 0.5. */
static void sh_to_rgb(int deg, int M, const float* pos, const float* campos, const float* sh,
                      float* out) {
  float res[3];
  for (int k = 0; k < 3; k++) res[k] = SH_C0 * sh[k];
  if (deg > 0) {
    float dx = pos[0] - campos[0], dy = pos[1] - campos[1], dz = pos[2] - campos[2];
    const float len = sqrtf(fmaf(dz, dz, fmaf(dx, dx, dy * dy)));
    const float x = dx / len, y = dy / len, z = dz / len;
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
  (void)M;
  for (int k = 0; k < 3; k++) out[k] = res[k] + 0.5f;
}

/*
 * forward.cu:155-256 preprocessCUDA. Outputs follow GeometryState (rasterizer_impl.h:33-47).
 * Buffers for culled Gaussians keep whatever the caller put there (the reference leaves them
 * uninitialised); radii and tiles_touched are always written.
 */
void gdo_preprocess(int P, int D, int M, const float* means3D, const float* scales,
                    float scale_modifier, const float* rotations, const float* opacities,
                    const float* shs, const float* cov3D_precomp, const float* colors_precomp,
                    const float* viewmatrix, const float* projmatrix, const float* campos, int W,
                    int H, float tanfovx, float tanfovy, int* radii, float* means2D, float* depths,
                    float* cov3Ds, float* rgb, float* conic_opacity, uint32_t* tiles_touched,
                    uint8_t* clamped) {
  const float focal_y = H / (2.0f * tanfovy); /* rasterizer_impl.cu:223-224 */
  const float focal_x = W / (2.0f * tanfovx);
  const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    radii[i] = 0;
    tiles_touched[i] = 0;
    const float x = means3D[3 * i], y = means3D[3 * i + 1], z = means3D[3 * i + 2];
    const float pvz = xform_row(viewmatrix, 2, x, y, z);
    if (pvz <= 0.2f) continue; /* auxiliary.h:154 near cull only */
    const float hx = xform_row(projmatrix, 0, x, y, z), hy = xform_row(projmatrix, 1, x, y, z),
                hw = xform_row(projmatrix, 3, x, y, z);
    const float pw = 1.0f / (hw + 0.0000001f);
    const float projx = hx * pw, projy = hy * pw;
    const float* c3;
    if (cov3D_precomp) {
      c3 = cov3D_precomp + 6 * (size_t)i;
    } else {
      cov3d_from_scale_rot(scales + 3 * (size_t)i, scale_modifier, rotations + 4 * (size_t)i,
                           cov3Ds + 6 * (size_t)i);
      c3 = cov3Ds + 6 * (size_t)i;
    }
    float t0[3], t1[3], a, b, c;
    cov2d_T(x, y, z, focal_x, focal_y, tanfovx, tanfovy, viewmatrix, t0, t1, 0, 0, 0);
    cov2d_from_T(t0, t1, c3, &a, &b, &c);
    const float det = fmaf(a, c, -(b * b));
    if (det == 0.0f) continue;
    const float det_inv = 1.0f / det;
    const float mid = (a + c) * 0.5f;
    const float s = sqrtf(fmaxf(fmaf(mid, mid, -det), 0.1f));
    const float lam = fmaxf(mid + s, mid - s);
    const float my_radius = ceilf(sqrtf(lam) * 3.0f);
    const float px = ndc2pix(projx, W), py = ndc2pix(projy, H);
    int rx0, ry0, rx1, ry1;
    get_rect(px, py, f2i_trunc(my_radius), gx, gy, &rx0, &ry0, &rx1, &ry1);
    if ((rx1 - rx0) * (ry1 - ry0) == 0) continue;
    if (!colors_precomp) {
      float col[3];
      sh_to_rgb(D, M, means3D + 3 * (size_t)i, campos, shs + 3 * (size_t)M * i, col);
      for (int k = 0; k < 3; k++) {
        clamped[3 * (size_t)i + k] = (col[k] < 0.0f);
        rgb[3 * (size_t)i + k] = fmaxf(col[k], 0.0f);
      }
    }
    depths[i] = pvz;
    radii[i] = f2i_trunc(my_radius);
    means2D[2 * (size_t)i] = px;
    means2D[2 * (size_t)i + 1] = py;
    conic_opacity[4 * (size_t)i + 0] = c * det_inv;
    conic_opacity[4 * (size_t)i + 1] = det_inv * (-b);
    conic_opacity[4 * (size_t)i + 2] = a * det_inv;
    conic_opacity[4 * (size_t)i + 3] = opacities[i];
    tiles_touched[i] = (uint32_t)((ry1 - ry0) * (rx1 - rx0));
  }
}

/* rasterizer_impl.cu:54-66,141-153 markVisible */
void gdo_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present) {
  for (int i = 0; i < P; i++)
    present[i] = !(xform_row(viewmatrix, 2, means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]) <=
                   0.2f);
}

/* rasterizer_impl.cu:35-50 */
uint32_t gdo_higher_msb(uint32_t n) {
  uint32_t msb = sizeof(n) * 4, step = msb;
  while (step > 1) {
    step /= 2;
    if (n >> msb) msb += step; else msb -= step;
  }
  if (n >> msb) msb++;
  return msb;
}

/* rasterizer_impl.cu:278 inclusive scan; returns num_rendered (:282). */
uint32_t gdo_scan(int P, const uint32_t* tiles_touched, uint32_t* point_offsets) {
  uint32_t acc = 0;
  for (int i = 0; i < P; i++) { acc += tiles_touched[i]; point_offsets[i] = acc; }
  return acc;
}

/*
 * rasterizer_impl.cu:70-111 duplicateWithKeys, :304-309 stable radix sort on the low 32+bit key
 * bits, :116-138 identifyTileRanges. Sorting: LSD radix, 8 bits per pass == any stable sort.
 */
void gdo_binning(int P, const float* means2D, const float* depths, const int* radii,
                 const uint32_t* point_offsets, int W, int H, uint32_t R, uint64_t* keys_unsorted,
                 uint32_t* values_unsorted, uint64_t* keys_sorted, uint32_t* point_list,
                 uint32_t* ranges /* [T][2], zeroed here (:311) */) {
  const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
  for (int i = 0; i < P; i++) {
    if (radii[i] <= 0) continue;
    uint32_t off = (i == 0) ? 0 : point_offsets[i - 1];
    int rx0, ry0, rx1, ry1;
    get_rect(means2D[2 * (size_t)i], means2D[2 * (size_t)i + 1], radii[i], gx, gy, &rx0, &ry0, &rx1,
             &ry1);
    uint32_t dbits;
    memcpy(&dbits, &depths[i], 4);
    for (int y = ry0; y < ry1; y++)
      for (int x = rx0; x < rx1; x++) {
        uint64_t key = (uint64_t)(y * gx + x);
        key <<= 32;
        key |= dbits;
        keys_unsorted[off] = key;
        values_unsorted[off] = (uint32_t)i;
        off++;
      }
  }
  const int bit = (int)gdo_higher_msb((uint32_t)(gx * gy));
  const int end_bit = 32 + bit;
  uint64_t* ka = (uint64_t*)malloc(sizeof(uint64_t) * (R + 1));
  uint64_t* kb = (uint64_t*)malloc(sizeof(uint64_t) * (R + 1));
  uint32_t* va = (uint32_t*)malloc(sizeof(uint32_t) * (R + 1));
  uint32_t* vb = (uint32_t*)malloc(sizeof(uint32_t) * (R + 1));
  memcpy(ka, keys_unsorted, sizeof(uint64_t) * R);
  memcpy(va, values_unsorted, sizeof(uint32_t) * R);
  for (int shift = 0; shift < end_bit; shift += 8) {
    const int nb = (end_bit - shift) < 8 ? (end_bit - shift) : 8;
    const uint64_t mask = ((uint64_t)1 << nb) - 1;
    size_t cnt[257];
    memset(cnt, 0, sizeof(cnt));
    for (uint32_t j = 0; j < R; j++) cnt[((ka[j] >> shift) & mask) + 1]++;
    for (int d = 0; d < 256; d++) cnt[d + 1] += cnt[d];
    for (uint32_t j = 0; j < R; j++) {
      size_t dst = cnt[(ka[j] >> shift) & mask]++;
      kb[dst] = ka[j];
      vb[dst] = va[j];
    }
    uint64_t* tk = ka; ka = kb; kb = tk;
    uint32_t* tv = va; va = vb; vb = tv;
  }
  memcpy(keys_sorted, ka, sizeof(uint64_t) * R);
  memcpy(point_list, va, sizeof(uint32_t) * R);
  free(ka); free(kb); free(va); free(vb);
  memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)gx * gy);
  for (uint32_t j = 0; j < R; j++) {
    const uint32_t cur = (uint32_t)(keys_sorted[j] >> 32);
    if (j == 0) ranges[2 * cur] = 0;
    else {
      const uint32_t prev = (uint32_t)(keys_sorted[j - 1] >> 32);
      if (cur != prev) { ranges[2 * prev + 1] = j; ranges[2 * cur] = j; }
    }
    if (j == R - 1) ranges[2 * cur + 1] = R;
  }
}

/*
 * forward.cu:261-381 renderCUDA. power uses the reference's contraction:
 *   fma(fma(dx, dx*cx, dy*(dy*cz)), -0.5, -(dy*(dx*cy)));  colour += T * (alpha*f) as one fma.
 */
static inline float pair_power(float dx, float dy, float cx, float cy, float cz) {
  return fmaf(fmaf(dx, dx * cx, dy * (dy * cz)), -0.5f, -(dy * (dx * cy)));
}

void gdo_render_forward(int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                        const float* means2D, const float* features, const float* depths,
                        const float* conic_opacity, const float* bg, float* out_color,
                        float* out_depth, float* out_alpha, uint32_t* n_contrib) {
  const int gx = (W + BLOCK_X - 1) / BLOCK_X;
#pragma omp parallel for schedule(dynamic, 4)
  for (int py = 0; py < H; py++)
    for (int px = 0; px < W; px++) {
      const int tile = (py / BLOCK_Y) * gx + (px / BLOCK_X);
      const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
      const float pfx = (float)px, pfy = (float)py;
      float T = 1.0f, C[3] = {0, 0, 0}, weight = 0.0f, Dp = 0.0f;
      uint32_t contributor = 0, last_contributor = 0;
      for (uint32_t j = r0; j < r1; j++) {
        contributor++;
        const uint32_t id = point_list[j];
        const float dx = means2D[2 * (size_t)id] - pfx, dy = means2D[2 * (size_t)id + 1] - pfy;
        const float* co = conic_opacity + 4 * (size_t)id;
        const float power = pair_power(dx, dy, co[0], co[1], co[2]);
        if (power > 0.0f) continue;
        const float alpha = fminf(0.99f, co[3] * expf(power));
        if (alpha < 1.0f / 255.0f) continue;
        const float test_T = T * (1.0f - alpha);
        if (test_T < 0.0001f) break; /* done = true */
        for (int ch = 0; ch < 3; ch++) C[ch] = fmaf(T, alpha * features[3 * (size_t)id + ch], C[ch]);
        weight = fmaf(T, alpha, weight);
        Dp = fmaf(T, alpha * depths[id], Dp);
        T = test_T;
        last_contributor = contributor;
      }
      const size_t pix = (size_t)py * W + px;
      n_contrib[pix] = last_contributor;
      for (int ch = 0; ch < 3; ch++) out_color[(size_t)ch * H * W + pix] = fmaf(T, bg[ch], C[ch]);
      out_alpha[pix] = weight; /* forward.cu:378: accumulated weight, not 1-T */
      out_depth[pix] = Dp;
    }
}

/*
 * backward.cu:415-601 renderCUDA (backward). The reference accumulates with float atomicAdd in
 * a non-deterministic order; the oracle accumulates per-pair float contributions in double.
 * acc layout per Gaussian: [0..1] dmean2D.xy, [2..4] dconic x,y,w, [5] dopacity, [6..8] dcolor,
 * [9] ddepth.
 */
void gdo_render_backward(int P, int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                         const float* bg, const float* means2D, const float* conic_opacity,
                         const float* colors, const float* depths, const float* alphas,
                         const uint32_t* n_contrib, const float* dL_dpixels,
                         const float* dL_dpixel_depths, const float* dL_dalphas, float* dL_dmean2D,
                         float* dL_dconic2D, float* dL_dopacity, float* dL_dcolors,
                         float* dL_ddepths) {
  const int gx = (W + BLOCK_X - 1) / BLOCK_X;
  double* acc = (double*)calloc((size_t)P * 10, sizeof(double));
  const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
  /* each OpenMP thread accumulates into a private copy, merged at the end */
#pragma omp parallel
  {
  double* lacc = (double*)calloc((size_t)P * 10, sizeof(double));
#pragma omp for schedule(dynamic, 4)
  for (int py = 0; py < H; py++)
    for (int px = 0; px < W; px++) {
      const int tile = (py / BLOCK_Y) * gx + (px / BLOCK_X);
      const uint32_t r0 = ranges[2 * tile];
      const size_t pix = (size_t)py * W + px;
      const float pfx = (float)px, pfy = (float)py;
      const float T_final = 1.0f - alphas[pix];
      float T = T_final;
      const int last = (int)n_contrib[pix];
      float accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0};
      float dLp[3];
      for (int ch = 0; ch < 3; ch++) dLp[ch] = dL_dpixels[(size_t)ch * H * W + pix];
      const float dLd = dL_dpixel_depths[pix], dLa = dL_dalphas[pix];
      float accum_depth_rec = 0, accum_alpha_rec = 0, last_alpha = 0, last_depth = 0;
      float bg_dot = 0;
      for (int ch = 0; ch < 3; ch++) bg_dot += bg[ch] * dLp[ch];
      for (int pos = last - 1; pos >= 0; pos--) {
        const uint32_t id = point_list[r0 + (uint32_t)pos];
        const float dx = means2D[2 * (size_t)id] - pfx, dy = means2D[2 * (size_t)id + 1] - pfy;
        const float* co = conic_opacity + 4 * (size_t)id;
        const float power = pair_power(dx, dy, co[0], co[1], co[2]);
        if (power > 0.0f) continue;
        const float G = expf(power);
        const float alpha = fminf(0.99f, co[3] * G);
        if (alpha < 1.0f / 255.0f) continue;
        T = T / (1.0f - alpha);
        const float dchannel_dcolor = alpha * T;
        float dL_dopa = 0.0f;
        double* a = lacc + (size_t)id * 10;
        for (int ch = 0; ch < 3; ch++) {
          const float c = colors[3 * (size_t)id + ch];
          accum_rec[ch] = last_alpha * last_color[ch] + (1.0f - last_alpha) * accum_rec[ch];
          last_color[ch] = c;
          dL_dopa += (c - accum_rec[ch]) * dLp[ch];
          a[6 + ch] += (double)(dchannel_dcolor * dLp[ch]);
        }
        const float c_d = depths[id];
        accum_depth_rec = last_alpha * last_depth + (1.0f - last_alpha) * accum_depth_rec;
        last_depth = c_d;
        dL_dopa += (c_d - accum_depth_rec) * dLd;
        a[9] += (double)(dchannel_dcolor * dLd);
        accum_alpha_rec = last_alpha + (1.0f - last_alpha) * accum_alpha_rec;
        dL_dopa += (1.0f - accum_alpha_rec) * dLa;
        dL_dopa *= T;
        last_alpha = alpha;
        dL_dopa += (-T_final / (1.0f - alpha)) * bg_dot;
        const float dL_dG = co[3] * dL_dopa;
        const float gdx = G * dx, gdy = G * dy;
        const float dG_ddelx = -gdx * co[0] - gdy * co[1];
        const float dG_ddely = -gdy * co[2] - gdx * co[1];
        a[0] += (double)(dL_dG * dG_ddelx * ddelx_dx);
        a[1] += (double)(dL_dG * dG_ddely * ddely_dy);
        a[2] += (double)(-0.5f * gdx * dx * dL_dG);
        a[3] += (double)(-0.5f * gdx * dy * dL_dG);
        a[4] += (double)(-0.5f * gdy * dy * dL_dG);
        a[5] += (double)(G * dL_dopa);
      }
    }
#pragma omp critical
  for (size_t k = 0; k < (size_t)P * 10; k++) acc[k] += lacc[k];
  free(lacc);
  }
  for (int i = 0; i < P; i++) {
    const double* a = acc + (size_t)i * 10;
    dL_dmean2D[3 * (size_t)i] = (float)a[0];
    dL_dmean2D[3 * (size_t)i + 1] = (float)a[1];
    dL_dmean2D[3 * (size_t)i + 2] = 0.0f;
    dL_dconic2D[4 * (size_t)i] = (float)a[2];
    dL_dconic2D[4 * (size_t)i + 1] = (float)a[3];
    dL_dconic2D[4 * (size_t)i + 2] = 0.0f; /* slot z never written (backward.cu:593-595) */
    dL_dconic2D[4 * (size_t)i + 3] = (float)a[4];
    dL_dopacity[i] = (float)a[5];
    for (int ch = 0; ch < 3; ch++) dL_dcolors[3 * (size_t)i + ch] = (float)a[6 + ch];
    dL_ddepths[i] = (float)a[9];
  }
  free(acc);
}

/* GLM-style 3x3, m[c][r] */
typedef struct { float m[3][3]; } mat3;
static mat3 m3mul(mat3 a, mat3 b) { /* (a*b)[c][r] = sum_k a[k][r]*b[c][k] */
  mat3 o;
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++)
      o.m[c][r] = a.m[0][r] * b.m[c][0] + a.m[1][r] * b.m[c][1] + a.m[2][r] * b.m[c][2];
  return o;
}
static mat3 m3t(mat3 a) {
  mat3 o;
  for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) o.m[c][r] = a.m[r][c];
  return o;
}

/* backward.cu:20-139 SH backward; adds the view-direction term to dL_dmean. */
static void sh_backward(int deg, int M, const float* pos, const float* campos, const float* sh,
                        const uint8_t* clamped, const float* dL_dcolor, float* dL_dmean,
                        float* dL_dsh) {
  float dox = pos[0] - campos[0], doy = pos[1] - campos[1], doz = pos[2] - campos[2];
  const float len = sqrtf(dox * dox + doy * doy + doz * doz);
  const float x = dox / len, y = doy / len, z = doz / len;
  float dRGB[3];
  for (int k = 0; k < 3; k++) dRGB[k] = dL_dcolor[k] * (clamped[k] ? 0.0f : 1.0f);
  float dx[3] = {0, 0, 0}, dy[3] = {0, 0, 0}, dz[3] = {0, 0, 0};
#define SHV(n, k) sh[3 * (n) + (k)]
#define DSH(n, w) for (int k = 0; k < 3; k++) dL_dsh[3 * (n) + k] = (w) * dRGB[k]
  DSH(0, SH_C0);
  if (deg > 0) {
    DSH(1, -SH_C1 * y); DSH(2, SH_C1 * z); DSH(3, -SH_C1 * x);
    for (int k = 0; k < 3; k++) { dx[k] = -SH_C1 * SHV(3, k); dy[k] = -SH_C1 * SHV(1, k); dz[k] = SH_C1 * SHV(2, k); }
    if (deg > 1) {
      const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      DSH(4, SH_C2[0] * xy); DSH(5, SH_C2[1] * yz); DSH(6, SH_C2[2] * (2.f * zz - xx - yy));
      DSH(7, SH_C2[3] * xz); DSH(8, SH_C2[4] * (xx - yy));
      for (int k = 0; k < 3; k++) {
        dx[k] += SH_C2[0] * y * SHV(4, k) + SH_C2[2] * 2.f * -x * SHV(6, k) + SH_C2[3] * z * SHV(7, k) + SH_C2[4] * 2.f * x * SHV(8, k);
        dy[k] += SH_C2[0] * x * SHV(4, k) + SH_C2[1] * z * SHV(5, k) + SH_C2[2] * 2.f * -y * SHV(6, k) + SH_C2[4] * 2.f * -y * SHV(8, k);
        dz[k] += SH_C2[1] * y * SHV(5, k) + SH_C2[2] * 2.f * 2.f * z * SHV(6, k) + SH_C2[3] * x * SHV(7, k);
      }
      if (deg > 2) {
        DSH(9, SH_C3[0] * y * (3.f * xx - yy)); DSH(10, SH_C3[1] * xy * z);
        DSH(11, SH_C3[2] * y * (4.f * zz - xx - yy));
        DSH(12, SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy));
        DSH(13, SH_C3[4] * x * (4.f * zz - xx - yy)); DSH(14, SH_C3[5] * z * (xx - yy));
        DSH(15, SH_C3[6] * x * (xx - 3.f * yy));
        for (int k = 0; k < 3; k++) {
          dx[k] += (SH_C3[0] * SHV(9, k) * 3.f * 2.f * xy + SH_C3[1] * SHV(10, k) * yz +
                    SH_C3[2] * SHV(11, k) * -2.f * xy + SH_C3[3] * SHV(12, k) * -3.f * 2.f * xz +
                    SH_C3[4] * SHV(13, k) * (-3.f * xx + 4.f * zz - yy) +
                    SH_C3[5] * SHV(14, k) * 2.f * xz + SH_C3[6] * SHV(15, k) * 3.f * (xx - yy));
          dy[k] += (SH_C3[0] * SHV(9, k) * 3.f * (xx - yy) + SH_C3[1] * SHV(10, k) * xz +
                    SH_C3[2] * SHV(11, k) * (-3.f * yy + 4.f * zz - xx) +
                    SH_C3[3] * SHV(12, k) * -3.f * 2.f * yz + SH_C3[4] * SHV(13, k) * -2.f * xy +
                    SH_C3[5] * SHV(14, k) * -2.f * yz + SH_C3[6] * SHV(15, k) * -3.f * 2.f * xy);
          dz[k] += (SH_C3[1] * SHV(10, k) * xy + SH_C3[2] * SHV(11, k) * 4.f * 2.f * yz +
                    SH_C3[3] * SHV(12, k) * 3.f * (2.f * zz - xx - yy) +
                    SH_C3[4] * SHV(13, k) * 4.f * 2.f * xz + SH_C3[5] * SHV(14, k) * (xx - yy));
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
  /* auxiliary.h:105-115 dnormvdv */
  const float sum2 = dox * dox + doy * doy + doz * doz;
  const float inv = 1.0f / sqrtf(sum2 * sum2 * sum2);
  dL_dmean[0] += ((+sum2 - dox * dox) * ddx - doy * dox * ddy - doz * dox * ddz) * inv;
  dL_dmean[1] += (-dox * doy * ddx + (sum2 - doy * doy) * ddy - doz * doy * ddz) * inv;
  dL_dmean[2] += (-dox * doz * ddx - doy * doz * ddy + (sum2 - doz * doz) * ddz) * inv;
}

/*
 * backward.cu:144-274 computeCov2DCUDA + :346-412 preprocessCUDA (+ :278-341 computeCov3D).
 * All dL_* outputs must be zero on entry for Gaussians with radii == 0 (they are skipped).
 */
void gdo_preprocess_backward(int P, int D, int M, const float* means3D, const int* radii,
                             const float* shs, const uint8_t* clamped, const float* scales,
                             const float* rotations, float scale_modifier, const float* cov3Ds,
                             const float* view, const float* proj, int W, int H, float tanfovx,
                             float tanfovy, const float* campos, const float* dL_dmean2D,
                             const float* dL_dconic, float* dL_dmean3D, const float* dL_dcolor,
                             const float* dL_ddepth, float* dL_dcov3D, float* dL_dsh,
                             float* dL_dscale, float* dL_drot) {
  const float h_y = H / (2.0f * tanfovy), h_x = W / (2.0f * tanfovx);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    if (!(radii[i] > 0)) continue;
    const float* c3 = cov3Ds + 6 * (size_t)i;
    const float mx = means3D[3 * i], my = means3D[3 * i + 1], mz = means3D[3 * i + 2];
    const float dcx = dL_dconic[4 * (size_t)i], dcy = dL_dconic[4 * (size_t)i + 1],
                dcz = dL_dconic[4 * (size_t)i + 3];
    float t0[3], t1[3], t[3], txtz, tytz, a, b, c;
    cov2d_T(mx, my, mz, h_x, h_y, tanfovx, tanfovy, view, t0, t1, t, &txtz, &tytz);
    cov2d_from_T(t0, t1, c3, &a, &b, &c);
    const float limx = 1.3f * tanfovx, limy = 1.3f * tanfovy;
    const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.0f : 1.0f;
    const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.0f : 1.0f;
    const float denom = a * c - b * b;
    float dL_da = 0, dL_db = 0, dL_dc = 0;
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    float* dcov = dL_dcov3D + 6 * (size_t)i;
    if (denom2inv != 0) {
      dL_da = denom2inv * (-c * c * dcx + 2 * b * c * dcy + (denom - a * c) * dcz);
      dL_dc = denom2inv * (-a * a * dcz + 2 * a * b * dcy + (denom - a * c) * dcx);
      dL_db = denom2inv * 2 * (b * c * dcx - (denom + 2 * b * b) * dcy + a * b * dcz);
      dcov[0] = (t0[0] * t0[0] * dL_da + t0[0] * t1[0] * dL_db + t1[0] * t1[0] * dL_dc);
      dcov[3] = (t0[1] * t0[1] * dL_da + t0[1] * t1[1] * dL_db + t1[1] * t1[1] * dL_dc);
      dcov[5] = (t0[2] * t0[2] * dL_da + t0[2] * t1[2] * dL_db + t1[2] * t1[2] * dL_dc);
      dcov[1] = 2 * t0[0] * t0[1] * dL_da + (t0[0] * t1[1] + t0[1] * t1[0]) * dL_db + 2 * t1[0] * t1[1] * dL_dc;
      dcov[2] = 2 * t0[0] * t0[2] * dL_da + (t0[0] * t1[2] + t0[2] * t1[0]) * dL_db + 2 * t1[0] * t1[2] * dL_dc;
      dcov[4] = 2 * t0[2] * t0[1] * dL_da + (t0[1] * t1[2] + t0[2] * t1[1]) * dL_db + 2 * t1[1] * t1[2] * dL_dc;
    } else {
      for (int k = 0; k < 6; k++) dcov[k] = 0;
    }
    /* Vrk symmetric: V[r][c] */
    const float V[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
    float p0[3], p1[3]; /* p0[k] = sum_j t0[j] V[k][j] */
    for (int k = 0; k < 3; k++) {
      p0[k] = t0[0] * V[k][0] + t0[1] * V[k][1] + t0[2] * V[k][2];
      p1[k] = t1[0] * V[k][0] + t1[1] * V[k][1] + t1[2] * V[k][2];
    }
    const float dT00 = 2 * p0[0] * dL_da + p1[0] * dL_db, dT01 = 2 * p0[1] * dL_da + p1[1] * dL_db,
                dT02 = 2 * p0[2] * dL_da + p1[2] * dL_db;
    const float dT10 = 2 * p1[0] * dL_dc + p0[0] * dL_db, dT11 = 2 * p1[1] * dL_dc + p0[1] * dL_db,
                dT12 = 2 * p1[2] * dL_dc + p0[2] * dL_db;
    /* W[c][r]: W[0]=(v0,v4,v8) W[1]=(v1,v5,v9) W[2]=(v2,v6,v10) */
    const float dJ00 = view[0] * dT00 + view[4] * dT01 + view[8] * dT02;
    const float dJ02 = view[2] * dT00 + view[6] * dT01 + view[10] * dT02;
    const float dJ11 = view[1] * dT10 + view[5] * dT11 + view[9] * dT12;
    const float dJ12 = view[2] * dT10 + view[6] * dT11 + view[10] * dT12;
    const float tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
    const float dtx = x_grad_mul * -h_x * tz2 * dJ02;
    const float dty = y_grad_mul * -h_y * tz2 * dJ12;
    const float dtz = -h_x * tz2 * dJ00 - h_y * tz2 * dJ11 + (2 * h_x * t[0]) * tz3 * dJ02 +
                      (2 * h_y * t[1]) * tz3 * dJ12;
    /* transformVec4x3Transpose (auxiliary.h:89-97) */
    float dmean[3] = {view[0] * dtx + view[1] * dty + view[2] * dtz,
                      view[4] * dtx + view[5] * dty + view[6] * dtz,
                      view[8] * dtx + view[9] * dty + view[10] * dtz};
    /* preprocessCUDA backward (:346-412) */
    const float m_w = 1.0f / (xform_row(proj, 3, mx, my, mz) + 0.0000001f);
    const float mul1 = (proj[0] * mx + proj[4] * my + proj[8] * mz + proj[12]) * m_w * m_w;
    const float mul2 = (proj[1] * mx + proj[5] * my + proj[9] * mz + proj[13]) * m_w * m_w;
    const float g2x = dL_dmean2D[3 * (size_t)i], g2y = dL_dmean2D[3 * (size_t)i + 1];
    dmean[0] += (proj[0] * m_w - proj[3] * mul1) * g2x + (proj[1] * m_w - proj[3] * mul2) * g2y;
    dmean[1] += (proj[4] * m_w - proj[7] * mul1) * g2x + (proj[5] * m_w - proj[7] * mul2) * g2y;
    dmean[2] += (proj[8] * m_w - proj[11] * mul1) * g2x + (proj[9] * m_w - proj[11] * mul2) * g2y;
    const float mul3 = view[2] * mx + view[6] * my + view[10] * mz + view[14];
    dmean[0] += (view[2] - view[3] * mul3) * dL_ddepth[i];
    dmean[1] += (view[6] - view[7] * mul3) * dL_ddepth[i];
    dmean[2] += (view[10] - view[11] * mul3) * dL_ddepth[i];
    if (shs)
      sh_backward(D, M, means3D + 3 * (size_t)i, campos, shs + 3 * (size_t)M * i,
                  clamped + 3 * (size_t)i, dL_dcolor + 3 * (size_t)i, dmean,
                  dL_dsh + 3 * (size_t)M * i);
    for (int k = 0; k < 3; k++) dL_dmean3D[3 * (size_t)i + k] = dmean[k];
    if (scales) { /* computeCov3D backward (:278-341) */
      const float* q = rotations + 4 * (size_t)i;
      const float r = q[0], x = q[1], y = q[2], z = q[3];
      mat3 R = {{{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                 {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                 {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}}};
      const float s[3] = {scale_modifier * scales[3 * (size_t)i], scale_modifier * scales[3 * (size_t)i + 1],
                          scale_modifier * scales[3 * (size_t)i + 2]};
      mat3 S = {{{s[0], 0, 0}, {0, s[1], 0}, {0, 0, s[2]}}};
      mat3 Mm = m3mul(S, R);
      mat3 dSig = {{{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                    {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                    {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}}};
      mat3 dM = m3mul(Mm, dSig);
      for (int c2 = 0; c2 < 3; c2++) for (int r2 = 0; r2 < 3; r2++) dM.m[c2][r2] *= 2.0f;
      mat3 Rt = m3t(R), dMt = m3t(dM);
      for (int k = 0; k < 3; k++)
        dL_dscale[3 * (size_t)i + k] =
            Rt.m[k][0] * dMt.m[k][0] + Rt.m[k][1] * dMt.m[k][1] + Rt.m[k][2] * dMt.m[k][2];
      for (int k = 0; k < 3; k++) for (int r2 = 0; r2 < 3; r2++) dMt.m[k][r2] *= s[k];
#define D_(a_, b_) dMt.m[a_][b_]
      float* dq = dL_drot + 4 * (size_t)i;
      dq[0] = 2 * z * (D_(0, 1) - D_(1, 0)) + 2 * y * (D_(2, 0) - D_(0, 2)) + 2 * x * (D_(1, 2) - D_(2, 1));
      dq[1] = 2 * y * (D_(1, 0) + D_(0, 1)) + 2 * z * (D_(2, 0) + D_(0, 2)) + 2 * r * (D_(1, 2) - D_(2, 1)) - 4 * x * (D_(2, 2) + D_(1, 1));
      dq[2] = 2 * x * (D_(1, 0) + D_(0, 1)) + 2 * r * (D_(2, 0) - D_(0, 2)) + 2 * z * (D_(1, 2) + D_(2, 1)) - 4 * y * (D_(2, 2) + D_(0, 0));
      dq[3] = 2 * r * (D_(0, 1) - D_(1, 0)) + 2 * x * (D_(2, 0) + D_(0, 2)) + 2 * y * (D_(1, 2) + D_(2, 1)) - 4 * z * (D_(1, 1) + D_(0, 0));
#undef D_
    }
  }
}
