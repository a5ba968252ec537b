/*
 * gd_unet.h -- C ABI of the B200-native operators behind the Stable-Diffusion UNet step of
 * StableDiffusionGuidance.compute_grad_sds
 * (reference: Garment_3DGS/threestudio/models/guidance/stable_diffusion_guidance.py:146-157 and
 * :185-276). The reference executes this step through diffusers 0.19.0 (not in the reference
 * tree; requirements.txt:12) on torch / cuDNN / cuBLAS; here every dense contraction is ONE
 * hand-written sm_100a kernel (tcgen05.mma with TMEM accumulators, operands staged by TMA) and
 * the normalisation / activation / softmax / SDS epilogue steps are small fused CUDA kernels.
 *
 * All pointers are DEVICE pointers. Activations are fp16, channels-last: an image tensor is
 * [N, H, W, C] and a token tensor is [rows, C]; weights are fp16 [N_out, K] (K contiguous),
 * 3x3 convolution weights are [C_out, 3, 3, C_in]. Every call is asynchronous on `stream`
 * and returns GD_UNET_OK or a negative error (message in gd_unet_last_error()).
 */
#ifndef GD_UNET_H_
#define GD_UNET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GD_UNET_OK 0
#define GD_UNET_NO_COLSTATS 1
#define GD_UNET_ERR_INVALID_ARG (-1)
#define GD_UNET_ERR_CUDA (-3)

typedef struct CUstream_st* gd_ustream_t; /* == cudaStream_t */

/* epilogue flags for gd_unet_gemm */
#define GD_EPI_SILU 1u        /* y = silu(y) after bias */
#define GD_EPI_TRANSPOSED 2u  /* store y^T (per batch): element (m,n) at C[n*ldc + m] */
#define GD_EPI_GEGLU 4u       /* weight rows interleaved (value, gate) in blocks of 16: out[m, n/2] */
#define GD_EPI_PROBE_SKIP 0x200u /* timing experiments only (tools/gemm_probe2.py): the epilogue drains TMEM but
                                    computes and stores nothing -- output undefined; never set by the product */

typedef struct {
  /* problem: C[z][M,N] (+)= A[z][M,K] * B[z][N,K]^T, fp16 in, fp32 accumulate, fp16 out */
  int M, N, K;
  int batch;          /* grid z (1 for plain layers; B*heads for attention) */
  int heads;          /* z -> (z / heads, z % heads); 1 when unused */
  /* A operand: 4-D tensor [d3][d2][d1][d0] fp16, d0 contiguous (= K direction) */
  const void* A;
  int a_dim[4];           /* d0..d3 extents in elements */
  long long a_stride[3];  /* byte strides of d1..d3 */
  int a_box[4];           /* TMA box d0..d3; a_box[0] = 64, product of the rest = 128 */
  /* conv mode (ntaps > 1 or explicit taps): K = ntaps * Ck; tap t reads channels
     [tap_c[t], tap_c[t] + Ck) at pixel offset (tap_dx[t], tap_dy[t]); OOB reads are zero */
  int ntaps, Ck;
  int tap_dx[9], tap_dy[9], tap_c[9];
  int rows_per_image;     /* conv: H*W of the output (rows m -> image m / rows_per_image) */
  int img_w, img_h;       /* conv: output width / height */
  /* batched (attention) addressing of A: coordinate d0 += (z % heads) * a_head_k,
     d2 = a_zflat ? z : z / heads */
  int a_head_k, a_zflat;
  /* B operand: 3-D tensor [d2][d1][d0] fp16 (d0 = K direction, d1 = N direction) */
  const void* B;
  int b_dim[3];
  long long b_stride[2];
  int b_head_k, b_head_n; /* d0 += (z % heads) * b_head_k; d1 += (z % heads) * b_head_n */
  /* output */
  void* C;
  long long ldc;          /* row stride in elements */
  long long c_batch_stride, c_head_stride; /* elements; offset = (z/heads)*bs + (z%heads)*hs */
  /* epilogue */
  const void* bias;       /* fp16 [N] or NULL */
  const void* row_bias;   /* fp16 [images, N] or NULL: added per image (time embedding) */
  const void* residual;   /* fp16, same addressing as C, or NULL */
  float alpha;            /* y = alpha * acc (+ bias ...) */
  unsigned flags;
  int block_n;            /* 16..256, multiple of 16; 0 = choose */
  long long row_bias_ld;  /* row stride of row_bias in elements; 0 = N */
  /* optional fused GroupNorm statistics of the OUTPUT (plain epilogue only): fp32 [ceil(M/32)][2][N], per block of 32 output
     rows and column the sum and the sum of squares of the fp16 values stored. The blocks partition the rows of each image
     (rows_per_image % 32 == 0) in tile order -- 32 consecutive rows, or four 8-pixel row segments of a patch tile -- so only
     sums over whole images are defined; that is what gd_unet_groupnorm_colstats consumes. gd_unet_gemm returns GD_UNET_NO_COLSTATS (1) when
     the chosen kernel variant cannot produce them (split-K, ragged N, transposed / GEGLU epilogue): the output tensor is
     complete, the statistics are not written. */
  void* colstats;
  /* optional GroupNorm(+SiLU)-BACKWARD producer epilogue for the data-gradient GEMM in front of that backward: gn_coef =
     fp32 [images][N][4] from gd_unet_gn_bwd_coef, residual = the GroupNorm INPUT x (same layout as C; it is NOT added),
     colstats required. C receives g = acc * silu'(GN(x)) and colstats sum g | sum g*xh per 32-row block and column, which
     gd_unet_groupnorm_bwd_g turns into dx without a statistics sweep over x and dz. Returns GD_UNET_NO_COLSTATS when the
     shape cannot take this epilogue: C then holds the PLAIN product (residual ignored), use gd_unet_groupnorm_bwd. */
  const void* gn_coef;
  /* > 0: the M rows are the pixels (n, y, x) of a half-resolution image of this width and row m is stored at the full-resolution
     pixel (n, 2y, 2x): output row 4*(m - x) + 2*x of C (the caller offsets C by the phase (py, px): + (py*2*w + px) rows). The
     four per-phase GEMMs of a stride-2 convolution's data gradient write the upsampled tensor directly (no depth-to-space). */
  int c_up2_w;
  /* 2: stride-2 convolution without a space-to-depth copy. A is a strided VIEW of the full-resolution NHWC input with dims
     {2C, W/2, H, N} (a pixel pair is one 2C-channel super-pixel); output row y reads input row 2*y + tap_dy, tap_dx shifts
     super-pixels and tap_c selects the pixel of the pair (0 or C). Needs an A box inside one image row (W/2 >= 128). 0/1: off. */
  int a_yscale;
} GdGemmArgs;

int gd_unet_gemm(const GdGemmArgs* args, gd_ustream_t stream);

/* Fused attention, head_dim 64: out[b, :, h*64:(h+1)*64] = softmax(scale * Q_h K_h^T) V_h.
   q [B,Tq,ldq], k [B,Tk,ldk] (head h in columns h*64..), vt = V^T [B, heads*64, ldv] (keys
   contiguous), out [B,Tq,ldo]; leading dimensions in elements, multiples of 8. The score matrix
   stays in TMEM / shared memory. */
int gd_unet_flash_attn(const void* q, const void* k, const void* vt, void* out, int B, int heads, int Tq,
                       int Tk, long long ldq, long long ldk, long long ldv, long long ldo, float scale,
                       gd_ustream_t stream);
/* Same with an explicit batch stride (elements) of vt, for V^T slices of a wider [B, sum C, ldv]
   tensor (all cross-attention V projections of the text computed by one GEMM); 0 = heads*64*ldv. */
int gd_unet_flash_attn_ex(const void* q, const void* k, const void* vt, void* out, int B, int heads, int Tq,
                          int Tk, long long ldq, long long ldk, long long ldv, long long ldo,
                          long long v_batch_stride, float scale, gd_ustream_t stream);

/* GroupNorm over NHWC fp16 (+ optional SiLU). x,y: [N, HW, C]; gamma,beta fp16 [C]. */
int gd_unet_groupnorm(const void* x, void* y, const void* gamma, const void* beta, int N, int HW,
                      int C, int groups, float eps, int silu, gd_ustream_t stream);
/* LayerNorm over the last dim. x,y: [rows, C] fp16. */
int gd_unet_layernorm(const void* x, void* y, const void* gamma, const void* beta, int rows, int C,
                      float eps, gd_ustream_t stream);
/* In-place row softmax of fp16 scores [rows, cols] (row stride ld elements), fp32 math. */
int gd_unet_softmax(void* s, long long rows, int cols, long long ld, gd_ustream_t stream);
/* GEGLU: y[r, j] = x[r, j] * gelu_erf(x[r, H + j]); x [rows, 2H], y [rows, H]. */
int gd_unet_geglu(const void* x, void* y, long long rows, int H, gd_ustream_t stream);
/* y = a + b (fp16, n elements) */
int gd_unet_add(const void* a, const void* b, void* y, long long n, gd_ustream_t stream);
/* nearest-neighbour 2x upsample, NHWC fp16: [N,H,W,C] -> [N,2H,2W,C] */
int gd_unet_upsample2x(const void* x, void* y, int N, int H, int W, int C, gd_ustream_t stream);
/* space-to-depth for the stride-2 3x3 convolution: [N,H,W,C] -> [N,H/2,W/2,4C],
   channel block (py*2+px)*C holds x[2y+py, 2x+px] */
int gd_unet_space_to_depth(const void* x, void* y, int N, int H, int W, int C, gd_ustream_t stream);
/* channel concat NHWC: y[..., :Ca] = a, y[..., Ca:] = b */
int gd_unet_concat(const void* a, const void* b, void* y, long long rows, int Ca, int Cb,
                   gd_ustream_t stream);
/* Small-M linear on CUDA cores (time embedding path): y[b, n] = act_out(sum_k act_in(x[b,k]) *
   W[n,k] + bias[n]); x fp16 [Bm,K], W fp16 [N,K], y fp16 [Bm,N]; Bm <= 16. */
int gd_unet_small_linear(const void* x, const void* W, const void* bias, void* y, int Bm, int K,
                         int N, int silu_in, int silu_out, gd_ustream_t stream);
/* Sinusoidal timestep embedding (flip_sin_to_cos, freq_shift 0): t fp32 [Bm] -> y fp16 [Bm, dim]
   = [cos | sin]; the timestep is first rounded to fp16 like the reference's t.to(fp16). */
int gd_unet_timestep_embedding(const float* t, void* y, int Bm, int dim, gd_ustream_t stream);
/* conv_in: 3x3, pad 1, C_in = 4 (NCHW fp16 input [N,4,H,W]) -> NHWC fp16 [N,H,W,Cout].
   w fp16 [Cout,3,3,4], bias fp16 [Cout]. */
int gd_unet_conv_in(const void* x_nchw, const void* w, const void* bias, void* y, int N, int H,
                    int W, int Cout, gd_ustream_t stream);
/* conv_out: 3x3, pad 1, NHWC fp16 [N,H,W,Cin] -> NCHW fp32 [N,4,H,W]; w fp16 [4,3,3,Cin]. */
int gd_unet_conv_out(const void* x, const void* w, const void* bias, float* y_nchw, int N, int H,
                     int W, int Cin, gd_ustream_t stream);
/* SDS prologue: latents fp32 [B,4,H,W], noise fp32, per-sample sqrt(abar), sqrt(1-abar) ->
   x_t fp32 [B,4,H,W] and the duplicated fp16 NCHW UNet input [reps*B,4,H,W]. */
int gd_unet_add_noise(const float* latents, const float* noise, const float* sqrt_ab,
                      const float* sqrt_1mab, float* latents_noisy, void* unet_in_f16, int B,
                      int reps, int chw, gd_ustream_t stream);
/* SDS epilogue (stable_diffusion_guidance.py:248-265): noise_pred = e_text + s*(e_text-e_uncond),
   grad = w * (noise_pred - noise); eps fp32 [2B,...] ordered (text, uncond). */
int gd_unet_sds_grad(const float* eps, const float* noise, const float* w, float guidance_scale,
                     float* noise_pred, float* grad, int B, int chw, gd_ustream_t stream);

/* ---- VAE encoder forward + input-gradient backward ------------------------------------------
 * encode_images (stable_diffusion_guidance.py:160-167): imgs*2-1 -> AutoencoderKL.encode ->
 * latent_dist.sample() * scaling_factor, differentiated w.r.t. the image by the SDS loss
 * (:424-427). The reference runs diffusers 0.19.0 AutoencoderKL in fp16 under autograd; here the
 * convolutions / attention matmuls are gd_unet_gemm calls (dgrad = the same implicit-GEMM with
 * flipped, transposed weights) and the functions below are the fused sweeps around them. */

/* GroupNorm (+SiLU) for tensors of any size that also returns the statistics the backward
   needs: stats fp32 [N*groups][2] = (mean, rstd). y may be NULL (statistics only). */
int gd_unet_groupnorm_stats(const void* x, void* y, const void* gamma, const void* beta, float* stats,
                            int N, int HW, int C, int groups, float eps, int silu, gd_ustream_t stream);
/* Backward of z = act(GroupNorm(x)) (act = SiLU if silu else identity) w.r.t. x:
   dx = dGN(x; stats)^T (dz * act'(y)) (+ add). x, dz, add (may be NULL), dx: fp16 [N,HW,C]. */
int gd_unet_groupnorm_bwd(const void* x, const void* dz, const void* add, void* dx, const void* gamma,
                          const void* beta, const float* stats, int N, int HW, int C, int groups, int silu,
                          gd_ustream_t stream);
/* In-place softmax backward: dP <- P * (dP - rowsum(P * dP)); fp16 [rows, cols], row stride ld. */
int gd_unet_softmax_bwd(const void* P, void* dP, long long rows, int cols, long long ld, gd_ustream_t stream);
/* Batched transpose fp16: x [B,R,C] -> y [B,C,R]. */
int gd_unet_transpose(const void* x, void* y, int B, int R, int C, gd_ustream_t stream);
/* Inverse of gd_unet_space_to_depth: [N,H/2,W/2,4C] -> [N,H,W,C]. */
int gd_unet_depth_to_space(const void* x, void* y, int N, int H, int W, int C, gd_ustream_t stream);
/* color fp32 NCHW [B,3,H,W] -> fp16 NCHW [B,4,H,W] = (a*color+shift | 0): conv_in input
   ((a, shift) = (2, -1) for images in [0,1], the `imgs * 2 - 1` of encode_images). */
int gd_vae_prep(const float* color_nchw, void* y, int B, int H, int W, float a, float shift,
                gd_ustream_t stream);
/* conv_in (3 -> C, 3x3, pad 1) as a GEMM: im2col rows A fp16 [B*H*W, 64], column (ky*3+kx)*3+c =
   a*color[b,c,y+ky-1,x+kx-1]+shift (0 outside the image; columns 27..63 zero). */
int gd_vae_im2col(const float* color_nchw, void* A, int B, int H, int W, float a, float shift,
                  gd_ustream_t stream);
/* Its data gradient from the per-pixel tap products Z fp16 [B*H*W, 32] (column (ky*3+kx)*3+c):
   dcolor[b,c,y,x] = scale * sum_taps Z[(b, y-ky+1, x-kx+1), (ky*3+kx)*3+c]; fp32 NCHW out. */
int gd_vae_dimg_gather(const void* Z, float* dcolor_nchw, int B, int H, int W, float scale,
                       gd_ustream_t stream);
/* The same with the DYNAMIC loss scale of gd_vae_grad_scale removed (dcolor = scale / *dyn_scale * sum ...) and
   non-finite results replaced by 0 (torch.nan_to_num(nan=0, posinf=0, neginf=0): an overflowed fp16 chain never
   reaches the Gaussians' Adam state). */
int gd_vae_dimg_gather_dyn(const void* Z, float* dcolor_nchw, int B, int H, int W, float scale,
                           const float* dyn_scale, gd_ustream_t stream);
/* Loss scale for the fp16 backward chain, chosen on the device from the gradient that enters it:
   *dyn_scale = 2^floor(log2(target / max|nan_to_num(clamp(grad, +-clip)) * pre|)), clamped to [2^-24, 2^24], 1 when the
   gradient is all zero. n = elements of grad; scratch: DEVICE fp32 [ceil(n/1024)]. With guidance scale 100 and
   grad_clip = None (the reference's Config default) a fixed scale overflows fp16; a power of two is exact. */
int gd_vae_grad_scale(const float* grad, long long n, float clip, float pre, float target, float* scratch,
                      float* dyn_scale, gd_ustream_t stream);
/* gd_vae_sample_bwd with that device-side scale as an extra factor. */
int gd_vae_sample_bwd_dyn(const float* grad, const void* moments, const float* noise, void* dmoments, int B,
                          int hw, int Cp, float scaling, float clip, float gscale, const float* dyn_scale,
                          gd_ustream_t stream);
/* DiagonalGaussianDistribution.sample() * scaling: moments fp16 NHWC [B,hw,8] (mean | logvar,
   logvar clamped to [-30,20]), noise fp32 NCHW [B,4,hw] -> latents fp32 NCHW [B,4,hw]. */
int gd_vae_sample(const void* moments, const float* noise, float* latents, int B, int hw, float scaling,
                  gd_ustream_t stream);
/* Its backward: grad fp32 NCHW (nan_to_num, clamp to +-clip if clip > 0, * scaling * gscale)
   -> d moments fp16 NHWC [B,hw,Cp] (8 real channels, zero padded to Cp). */
int gd_vae_sample_bwd(const float* grad, const void* moments, const float* noise, void* dmoments, int B,
                      int hw, int Cp, float scaling, float clip, float gscale, gd_ustream_t stream);
/* d image: fp16 NHWC [B,H,W,Cp] (3 real channels) -> fp32 NCHW [B,3,H,W] * scale. */
int gd_vae_dimg(const void* dx, float* dcolor_nchw, int B, int H, int W, int Cp, float scale,
                gd_ustream_t stream);

/* GroupNorm (+SiLU) of x fp16 NHWC [N,HW,C] from the column statistics a producing gd_unet_gemm left behind (GdGemmArgs.colstats)
   instead of a statistics pass over x. The C channels are the concatenation of Ca channels described by statsA and Cb by statsB
   (torch.cat([x, skip], 1) of the UNet up path; Cb = 0 / statsB = NULL for a single producer); both cover the same N*HW rows,
   HW % 32 == 0. Writes y; if mean_rstd != NULL also the fp32 (mean, rstd) table [N*groups][2] the backward needs. */
int gd_unet_groupnorm_colstats(const void* x, void* y, const void* gamma, const void* beta, float* mean_rstd,
                               const float* statsA, int Ca, const float* statsB, int Cb, int N, int HW, int C,
                               int groups, float eps, int silu, gd_ustream_t stream);

/* conv_in / conv_out of the UNet (4 latent channels) on the tensor-core GEMM:
   gd_unet_im2col4: x fp16 NCHW [N,4,H,W] -> A fp16 [N*H*W][64], A[p][(ky*3+kx)*4 + c] = x[n, c, y+ky-1, x+kx-1] (zero outside the
     image and in columns 36..63); conv_in is then gd_unet_gemm with the weights laid out [Cout][64] the same way.
   gd_unet_unpack4_nchw: the first 4 of ld fp16 channels per pixel (the 3x3 conv_out run with Cout padded to 16) -> fp32 NCHW. */
int gd_unet_im2col4(const void* x_nchw, void* A, int N, int H, int W, gd_ustream_t stream);
int gd_unet_unpack4_nchw(const void* x_nhwc, float* y_nchw, int N, long long HW, int ld, gd_ustream_t stream);

/* GroupNorm(+SiLU) backward split around the producing data-gradient GEMM (see GdGemmArgs.gn_coef):
   gd_unet_gn_bwd_coef: stats = (mean, rstd) [N*groups][2] of the forward -> coef fp32 [N][C][4] = (ya, yb, ca, cb),
     xh = x*ca + cb, y = x*ya + yb = xh*gamma + beta.
   gd_unet_groupnorm_bwd_g: g = dz*silu'(y) (the GEMM output) and its colstats -> dx = rstd*(gamma*g - S1 - xh*S2) (+ add). */
int gd_unet_gn_bwd_coef(const float* stats, const void* gamma, const void* beta, float* coef, int N, int C, int groups,
                        gd_ustream_t stream);
int gd_unet_groupnorm_bwd_g(const void* x, const void* g, const void* add, void* dx, const void* gamma, const void* beta,
                            const float* stats, const float* colstats, int N, int HW, int C, int groups, gd_ustream_t stream);

/* F.interpolate(x, (Ho, Wo), mode="bilinear", align_corners=False) on fp32 planes [BC, Hi, Wi] -> [BC, Ho, Wo]
   (stable_diffusion_guidance.py:387-396: the rendered batch is resized to 512^2 before encode_images), and its
   transpose dout [BC, Ho, Wo] -> din [BC, Hi, Wi] (a deterministic gather). */
int gd_resize_bilinear(const float* in, float* out, int BC, int Hi, int Wi, int Ho, int Wo, gd_ustream_t stream);
int gd_resize_bilinear_bwd(const float* dout, float* din, int BC, int Hi, int Wi, int Ho, int Wo, gd_ustream_t stream);

/* Bench glue, NOT part of the reference path: a fixed linear stand-in for the VAE encoder that
   the reference runs between the rasteriser and compute_grad_sds (encode_images, :160-167; out of
   this build's scope). latents [B,4,H/8,W/8] = mix[4][3] * mean_8x8(2*color-1); _bwd is its exact
   transpose applied to nan_to_num(clamp(grad, +-clip)) * scale (clip <= 0: no clamp). */
int gd_unet_pool_latents(const float* color_nchw, const float* mix, float* latents, int B, int H, int W,
                         gd_ustream_t stream);
int gd_unet_pool_latents_bwd(const float* grad, const float* mix, float* dcolor_nchw, int B, int H, int W,
                             float clip, float scale, gd_ustream_t stream);

const char* gd_unet_last_error(void);
uint64_t gd_unet_launch_count(void);
/* GEMM launches so far that ran as CTA pairs (tcgen05 cta_group::2, clusters of 2). */
uint64_t gd_unet_pair_launch_count(void);
const char* gd_unet_version(void);
/* Scratch (split-K partials 96 MB, GroupNorm partial statistics) is kept PER CUDA DEVICE and allocated on first
 * use on that device; gd_unet_init() allocates the current device's set up front -- call it before capturing a
 * CUDA graph (cudaMalloc is illegal during capture). The scratch is shared by all streams of a device: issue this
 * library's calls from ONE stream per device at a time. */
int gd_unet_init(void);

#ifdef __cplusplus
}
#endif
#endif /* GD_UNET_H_ */
