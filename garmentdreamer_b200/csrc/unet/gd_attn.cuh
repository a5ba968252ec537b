// Fused attention for sm_100a: O = softmax(scale * Q K^T) V per (batch, head), head_dim 64.
//
// One CTA per 128 query rows of one (batch, head); 192 threads, warp-specialised:
//   warp 0    : TMA producer (Q once; K tiles twice; V^T tiles once)
//   warp 1    : tcgen05.mma issuer  S = Q K_j^T  (M128 N128 K64, TMEM, double buffered)
//                                   O += P_j V_j (M128 N64  K128, TMEM)
//   warps 2-9 : softmax: tcgen05.ld S, P = exp2((S - m) * c) -> fp16 -> shared memory in the
//               128B-swizzled K-major layout the MMA reads as its A operand. Two warps share a
//               TMEM lane quarter and split each 128-key tile into its two 64-key blocks.
// Two passes over the keys: pass 1 only reduces the row maxima m (no exp), pass 2 recomputes S and
// accumulates O and the row sums with the FINAL maximum, so O never needs rescaling in TMEM.
// The S matrix (1.3 GB per 64x64 layer at batch 8) never touches HBM.
#pragma once
#include "gd_gemm.cuh"

namespace gdu {

struct AttnParams {
  int Tq, Tk, heads;
  int n_kv;              // ceil(Tk / 128)
  float scale_log2e;     // softmax scale * log2(e)
  __half* O;             // [B, Tq, ldo]
  long long ldo;
  int stagger;           // k_flash_attn2: cycles the second softmax warpgroup starts late (keeps the two out of lockstep)
};

constexpr int kAttnThreads = 64 + 8 * 32;  // producer, issuer, 8 softmax warps (2 per TMEM lane quarter)
constexpr uint32_t kQBytes = 128 * 64 * 2, kKBytes = 128 * 64 * 2, kVBytes = 2 * 64 * 64 * 2, kPBytes = 2 * 128 * 64 * 2;

__device__ __forceinline__ void bar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s2u(b)) : "memory");
}
__device__ __forceinline__ uint32_t umma_idesc_f16_n(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(kAttnThreads, 1)
k_flash_attn(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
             const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kQBytes;          // 2 stages
  uint8_t* sV = sK + 2 * kKBytes;      // 2 stages x 2 key blocks x [64 d x 64 keys]
  uint8_t* sP = sV + 2 * kVBytes;      // 2 buffers x 2 key blocks x [128 rows x 64 keys]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kPBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;   uint64_t* k_empty = bars + 3;
  uint64_t* v_full = bars + 5;   uint64_t* v_empty = bars + 7;
  uint64_t* s_full = bars + 9;   uint64_t* s_empty = bars + 11;
  uint64_t* p_full = bars + 13;  uint64_t* p_empty = bars + 15;
  uint64_t* o_full = bars + 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blk = blockIdx.x, bh = blockIdx.y, b = bh / p.heads, h = bh % p.heads;
  const int n = p.n_kv, iters = 2 * n;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
    bar_init(q_full, 1);
    for (int s = 0; s < 2; s++) {
      bar_init(&k_full[s], 1); bar_init(&k_empty[s], 1);
      bar_init(&v_full[s], 1); bar_init(&v_empty[s], 1);
      bar_init(&s_full[s], 1); bar_init(&s_empty[s], 8);   // one arrive per softmax warp
      bar_init(&p_full[s], 8); bar_init(&p_empty[s], 1);
    }
    bar_init(o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  const uint32_t tmem_S0 = tmem, tmem_O = tmem + 256;

  if (warp == 0) {
    if (elect_one()) {
      bar_expect_tx(q_full, kQBytes);
      tma_load_3d(sQ, &tmQ, q_full, h * 64, m_blk * 128, b);
      for (int i = 0; i < iters; i++) {
        const int j = i % n, s = i & 1;
        bar_wait(&k_empty[s], ((i >> 1) & 1) ^ 1);
        bar_expect_tx(&k_full[s], kKBytes);
        tma_load_3d(sK + s * kKBytes, &tmK, &k_full[s], h * 64, j * 128, b);
        if (i >= n) {
          const int vi = i - n, vs = vi & 1;
          bar_wait(&v_empty[vs], ((vi >> 1) & 1) ^ 1);
          bar_expect_tx(&v_full[vs], kVBytes);
          tma_load_3d(sV + vs * kVBytes, &tmV, &v_full[vs], j * 128, h * 64, b);
          tma_load_3d(sV + vs * kVBytes + 64 * 64 * 2, &tmV, &v_full[vs], j * 128 + 64, h * 64, b);
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc_s = umma_idesc_f16_n(128), idesc_o = umma_idesc_f16_n(64);
    bar_wait(q_full, 0);
    auto issue_S = [&](int i) {
      const int s = i & 1;
      bar_wait(&k_full[s], (i >> 1) & 1);
      bar_wait(&s_empty[s], ((i >> 1) & 1) ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint64_t da = umma_desc_sw128(s2u(sQ)), db = umma_desc_sw128(s2u(sK + s * kKBytes));
#pragma unroll
        for (int k = 0; k < 4; k++) umma_f16(tmem_S0 + s * 128, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc_s, k ? 1u : 0u);
        umma_commit(&k_empty[s]);
        umma_commit(&s_full[s]);
      }
      __syncwarp();
    };
    issue_S(0);
    for (int i = 0; i < iters; i++) {
      if (i + 1 < iters) issue_S(i + 1);
      if (i >= n) {
        const int vi = i - n, vs = vi & 1;   // P / V buffers are indexed from the start of pass 2
        bar_wait(&p_full[vs], (vi >> 1) & 1);
        bar_wait(&v_full[vs], (vi >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
#pragma unroll
          for (int kb = 0; kb < 2; kb++) {
            const uint64_t da = umma_desc_sw128(s2u(sP + vs * kPBytes + kb * 128 * 64 * 2));
            const uint64_t db = umma_desc_sw128(s2u(sV + vs * kVBytes + kb * 64 * 64 * 2));
#pragma unroll
            for (int k = 0; k < 4; k++)
              umma_f16(tmem_O, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc_o, (vi | kb | k) ? 1u : 0u);
          }
          umma_commit(&p_empty[vs]);
          umma_commit(&v_empty[vs]);
          if (i == iters - 1) umma_commit(o_full);
        }
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;                       // TMEM lane quarter of this warp
    const int hb = (warp - 2) >> 2;               // which 64-key block of every tile this warp owns
    const int r = q * 32 + lane;                  // row inside the 128-row tile
    const int row = m_blk * 128 + r;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float* s_x = reinterpret_cast<float*>(tmem_slot + 4);   // [2][128] exchange between the two halves
    float m = -INFINITY, l = 0.f;
    for (int i = 0; i < iters; i++) {
      const int j = i % n, s = i & 1;
      const int kmax = p.Tk - j * 128 - hb * 64;  // keys >= kmax in this warp's block are padding
      bar_wait(&s_full[s], (i >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (i < n) {
        // ---- pass 1: row maximum over this warp's 64 keys ----
        uint32_t va[32], vb[32];
        tmem_ld32_issue(tmem_S0 + s * 128 + lane_base + hb * 64, va);
        tmem_ld32_issue(tmem_S0 + s * 128 + lane_base + hb * 64 + 32, vb);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 2; c++) {
          uint32_t (&v)[32] = c == 0 ? va : vb;
          if (kmax >= 64) {   // four independent chains instead of one 32-deep FMNMX dependency chain
            float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]), m3 = __uint_as_float(v[3]);
#pragma unroll
            for (int t = 4; t < 32; t += 4) {
              m0 = fmaxf(m0, __uint_as_float(v[t])); m1 = fmaxf(m1, __uint_as_float(v[t + 1]));
              m2 = fmaxf(m2, __uint_as_float(v[t + 2])); m3 = fmaxf(m3, __uint_as_float(v[t + 3]));
            }
            m = fmaxf(m, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
          } else {
#pragma unroll
            for (int t = 0; t < 32; t++)
              if (c * 32 + t < kmax) m = fmaxf(m, __uint_as_float(v[t]));
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) bar_arrive(&s_empty[s]);
        if (i == n - 1) {  // combine the maxima of the two key halves (softmax warps only: barrier 1)
          s_x[hb * 128 + r] = m;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          m = fmaxf(m, s_x[(hb ^ 1) * 128 + r]);
        }
      } else {
        // ---- pass 2: P = exp2(S*c - m*c), row sums, P -> shared memory (A operand of P V) ----
        const int vi = i - n, ps = vi & 1;
        bar_wait(&p_empty[ps], ((vi >> 1) & 1) ^ 1);
        const float mc = m * p.scale_log2e;
        uint8_t* blk = sP + ps * kPBytes + hb * (128 * 64 * 2) + r * 128;
        uint32_t va[32], vb[32];       // both 32-key chunks in flight, one wait, then the S buffer is free
        tmem_ld32_issue(tmem_S0 + s * 128 + lane_base + hb * 64, va);
        tmem_ld32_issue(tmem_S0 + s * 128 + lane_base + hb * 64 + 32, vb);
        tmem_ld_wait();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) bar_arrive(&s_empty[s]);
#pragma unroll
        for (int c = 0; c < 2; c++) {
          uint32_t (&v)[32] = c == 0 ? va : vb;
          float lp[4] = {0.f, 0.f, 0.f, 0.f};                 // independent partial row sums (no 64-deep FADD chain)
#pragma unroll
          for (int g = 0; g < 4; g++) {                       // 4 chunks of 8 keys = 16 bytes
            __align__(16) __half2 hv[4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
              float e0, e1;
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(__uint_as_float(v[g * 8 + 2 * t]), p.scale_log2e, -mc)));
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(__uint_as_float(v[g * 8 + 2 * t + 1]), p.scale_log2e, -mc)));
              if (kmax < 64) {
                const int k0 = c * 32 + g * 8 + 2 * t;
                if (k0 >= kmax) e0 = 0.f;
                if (k0 + 1 >= kmax) e1 = 0.f;
              }
              lp[t] += e0 + e1;
              hv[t] = __floats2half2_rn(e0, e1);
            }
            const int chunk = c * 4 + g;                       // 16-byte chunk index in the 128 B row
            *reinterpret_cast<uint4*>(blk + ((chunk ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(hv);
          }
          l += (lp[0] + lp[1]) + (lp[2] + lp[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> async proxy
        __syncwarp();
        if (lane == 0) bar_arrive(&p_full[ps]);
      }
    }
    // ---- epilogue: O / l; the two warps of a quarter take 32 of the 64 output columns each ----
    asm volatile("bar.sync 1, 256;" ::: "memory");   // everyone has read the pass-1 exchange
    s_x[hb * 128 + r] = l;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    l += s_x[(hb ^ 1) * 128 + r];
    bar_wait(o_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const float inv = 1.0f / l;
    __half* dst = p.O + ((long long)b * p.Tq + row) * p.ldo + h * 64 + hb * 32;
    {
      uint32_t v[32];
      tmem_ld32(tmem_O + lane_base + hb * 32, v);
      if (row < p.Tq) {
        __align__(16) __half o[32];
#pragma unroll
        for (int t = 0; t < 32; t++) o[t] = __float2half_rn(__uint_as_float(v[t]) * inv);
#pragma unroll
        for (int u = 0; u < 4; u++) reinterpret_cast<uint4*>(dst)[u] = reinterpret_cast<const uint4*>(o)[u];
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}


// =====================================================================================================================
// Single-pass variant (online softmax) for self-attention over long key ranges (the 64x64 / 32x32 latent layers).
//
// One CTA per 256 query rows of one (batch, head): TWO 128-row Q tiles ping-pong through the tensor core so that the
// MUFU-bound softmax of one tile overlaps the MMAs of the other (d = 64: one ex2 per 256 tensor flops, the SFU is the
// roof: 16 ex2/clk/SM). 320 threads:
//   warp 0     : TMA producer (Q0, Q1 once; K_j / V^T_j tiles once each, 3 stages)
//   warp 1     : tcgen05.mma issuer   S_q = Q_q K_j^T  (M128 N128 K64, smem x smem -> TMEM)
//                                     O_q += P_q V_j    (M128 N64 K128, A = P_q read from TMEM, B = V^T_j from smem)
//   warps 2-5  : softmax of Q tile 0, warps 6-9: softmax of Q tile 1. A thread owns ONE row: tcgen05.ld of its 128
//                scores, row maximum, P = exp2((S - m) c) -> fp16x2 -> tcgen05.st into the P region of TMEM. No shared
//                memory traffic, no generic->async proxy fence, no exchange between threads.
// The running maximum is LAZY: m only moves when the tile maximum exceeds it by more than 2^8 (P stays below 256, exact
// in fp16's range; l and O are fp32), and only then O_q is rescaled in TMEM by the owning threads (tcgen05.ld/st of the
// row) -- after the first tiles this almost never happens. TMEM: S0 S1 | O0 O1 | P0 P1 = 2x128 + 2x64 + 2x64 = 512 columns.
//
// Measured on B200 (clock64 timeline of one CTA, 64x64 latents, 2780 cycles per pair of 128x128 score tiles): per warp and
// tile ~80 wait for S, ~130 tcgen05.ld, ~330 row maximum, ~1900 ex2 phase, ~150 wait for the previous P V, ~60 tcgen05.st.
// Two bounds sit close together: the SFU (2 x 128 x 128 ex2 at 16 / clk / SM = 2048 cycles; halving the ex2 count halves the
// ex2 phase) and the S -> softmax -> P -> P V barrier chain (~2400 cycles: with half the ex2 work the warps wait for S
// instead). Variants tried and dropped because they did not move the total: a quarter of the ex2 on the FMA pipe
// (Cody-Waite polynomial; +9 instructions per score made the warps issue-bound: 2980 cycles), and sixteen softmax warps
// with two threads per row (same 2800 cycles); ex2.approx.f16x2 on packed exponents (ptxas emits TWO scalar MUFU.EX2.F16 per
// pair, no packed SFU op on sm_100a: 255 -> 317 us); an SFU token that makes the two softmax warps of a scheduler take turns
// in the ex2 phase (one warp alone feeds the pipe at ~13 of 16 ex2/clk: 269 -> 296 us).
constexpr int kAttn2Threads = 64 + 8 * 32;
constexpr int kAttn2Stages = 3;

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kAttn2Threads, 1)
k_flash_attn2(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                              // 2 Q tiles
  uint8_t* sK = sQ + 2 * kQBytes;                  // kAttn2Stages K tiles [128 keys x 64 d]
  uint8_t* sV = sK + kAttn2Stages * kKBytes;       // kAttn2Stages x 2 key blocks x [64 d x 64 keys]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kAttn2Stages * kVBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;    uint64_t* k_empty = k_full + kAttn2Stages;
  uint64_t* v_full = k_empty + kAttn2Stages;  uint64_t* v_empty = v_full + kAttn2Stages;
  uint64_t* s_full = v_empty + kAttn2Stages;  // [2] S_q written by the tensor core
  uint64_t* s_free = s_full + 2;              // [2] S_q read into registers by its 4 softmax warps
  uint64_t* p_full = s_free + 2;              // [2] P_q (and a possibly rescaled O_q) in TMEM
  uint64_t* pv_done = p_full + 2;             // [2] O_q += P_q V_j retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_pair = blockIdx.x, bh = blockIdx.y, b = bh / p.heads, h = bh % p.heads;
  const int n = p.n_kv;
  const int nq = (m_pair * 256 + 128 < p.Tq) ? 2 : 1;   // the last pair of a ragged Tq may hold one tile only

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
    bar_init(q_full, 1);
    for (int s = 0; s < kAttn2Stages; s++) {
      bar_init(&k_full[s], 1); bar_init(&k_empty[s], 1);
      bar_init(&v_full[s], 1); bar_init(&v_empty[s], 1);
    }
    for (int q = 0; q < 2; q++) {
      bar_init(&s_full[q], 1); bar_init(&s_free[q], 4);
      bar_init(&p_full[q], 4); bar_init(&pv_done[q], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  const uint32_t tmem_S = tmem, tmem_O = tmem + 256, tmem_P = tmem + 384;

  if (warp == 0) {
    if (elect_one()) {
      bar_expect_tx(q_full, nq * kQBytes);
      for (int q = 0; q < nq; q++) tma_load_3d(sQ + q * kQBytes, &tmQ, q_full, h * 64, m_pair * 256 + q * 128, b);
      for (int j = 0; j < n; j++) {
        const int s = j % kAttn2Stages, ph = (j / kAttn2Stages) & 1;
        bar_wait(&k_empty[s], ph ^ 1);
        bar_expect_tx(&k_full[s], kKBytes);
        tma_load_3d(sK + s * kKBytes, &tmK, &k_full[s], h * 64, j * 128, b);
        bar_wait(&v_empty[s], ph ^ 1);
        bar_expect_tx(&v_full[s], kVBytes);
        tma_load_3d(sV + s * kVBytes, &tmV, &v_full[s], j * 128, h * 64, b);
        tma_load_3d(sV + s * kVBytes + 64 * 64 * 2, &tmV, &v_full[s], j * 128 + 64, h * 64, b);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc_s = umma_idesc_f16_n(128), idesc_o = umma_idesc_f16_n(64);
    bar_wait(q_full, 0);
    auto issue_S = [&](int j) {
      const int s = j % kAttn2Stages, ph = (j / kAttn2Stages) & 1;
      bar_wait(&k_full[s], ph);
      for (int q = 0; q < nq; q++) {
        if (j > 0) bar_wait(&s_free[q], (j - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint64_t da = umma_desc_sw128(s2u(sQ + q * kQBytes)), db = umma_desc_sw128(s2u(sK + s * kKBytes));
#pragma unroll
          for (int k = 0; k < 4; k++) umma_f16(tmem_S + q * 128, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc_s, k ? 1u : 0u);
          umma_commit(&s_full[q]);
          if (q == nq - 1) umma_commit(&k_empty[s]);
        }
        __syncwarp();
      }
    };
    issue_S(0);
    for (int j = 0; j < n; j++) {
      if (j + 1 < n) issue_S(j + 1);
      const int s = j % kAttn2Stages, ph = (j / kAttn2Stages) & 1;
      bar_wait(&v_full[s], ph);
      for (int q = 0; q < nq; q++) {
        bar_wait(&p_full[q], j & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
#pragma unroll
          for (int kb = 0; kb < 2; kb++) {
            const uint64_t db = umma_desc_sw128(s2u(sV + s * kVBytes + kb * 64 * 64 * 2));
#pragma unroll
            for (int k = 0; k < 4; k++)
              umma_f16_ts(tmem_O + q * 64, tmem_P + q * 64 + kb * 32 + k * 8, db + (uint64_t)(k * 2), idesc_o, (j | kb | k) ? 1u : 0u);
          }
          umma_commit(&pv_done[q]);
          if (q == nq - 1) umma_commit(&v_empty[s]);
        }
        __syncwarp();
      }
    }
  } else {
    const int q = (warp - 2) >> 2;                 // Q tile of this warpgroup
    if (q < nq) {
      const int r = (warp & 3) * 32 + lane;        // row inside the tile = TMEM lane
      const int row = m_pair * 256 + q * 128 + r;
      const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
      const uint32_t tS = tmem_S + q * 128 + lane_base, tO = tmem_O + q * 64 + lane_base, tP = tmem_P + q * 64 + lane_base;
      const float c = p.scale_log2e;
      float m = -INFINITY, l = 0.f;
      for (int j = 0; j < n; j++) {
        bar_wait(&s_full[q], j & 1);
        if (j == 0 && q == 1 && p.stagger > 0) {
          // Both warpgroups would otherwise run in lockstep (S0 and S1 land back to back): their ex2 phases then collide on the
          // SFU and their LDTM / max / STTM phases leave it idle together. Half a period of offset makes one fill the other's gap.
          const long long t0 = clock64();
          while (clock64() - t0 < p.stagger) {}
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t v[128];
        tmem_ld32_issue(tS, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld32_issue(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        tmem_ld32_issue(tS + 64, *reinterpret_cast<uint32_t(*)[32]>(&v[64]));
        tmem_ld32_issue(tS + 96, *reinterpret_cast<uint32_t(*)[32]>(&v[96]));
        tmem_ld_wait();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) bar_arrive(&s_free[q]);
        const int kvalid = p.Tk - j * 128;         // keys >= kvalid of this tile are padding
        if (kvalid < 128) {
#pragma unroll
          for (int t = 0; t < 128; t++)
            if (t >= kvalid) v[t] = 0xff800000u;   // -inf
        }
        float mm[8];                               // 8 independent FMNMX3 chains
#pragma unroll
        for (int u = 0; u < 8; u++) mm[u] = fmaxf(__uint_as_float(v[u]), __uint_as_float(v[u + 8]));
#pragma unroll
        for (int t = 16; t < 128; t += 16) {
#pragma unroll
          for (int u = 0; u < 8; u++) mm[u] = fmaxf(mm[u], fmaxf(__uint_as_float(v[t + u]), __uint_as_float(v[t + u + 8])));
        }
        const float mx = fmaxf(fmaxf(fmaxf(mm[0], mm[1]), fmaxf(mm[2], mm[3])), fmaxf(fmaxf(mm[4], mm[5]), fmaxf(mm[6], mm[7])));
        const bool grow = (mx - m) * c > 8.0f;     // lazy: the reference maximum moves only on a 2^8 overshoot (always at j = 0)
        float factor = 1.0f;
        if (grow) { factor = ex2_approx((m - mx) * c); m = mx; }
        l *= factor;
        const float mc = m * c;
        uint32_t pk[64];
        float lp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int t = 0; t < 64; t++) {
          const float e0 = ex2_approx(fmaf(__uint_as_float(v[2 * t]), c, -mc));
          const float e1 = ex2_approx(fmaf(__uint_as_float(v[2 * t + 1]), c, -mc));
          lp[t & 3] += e0 + e1;
          const __half2 hh = __floats2half2_rn(e0, e1);
          pk[t] = *reinterpret_cast<const uint32_t*>(&hh);
        }
        l += (lp[0] + lp[1]) + (lp[2] + lp[3]);
        if (j > 0) {
          bar_wait(&pv_done[q], (j - 1) & 1);      // P_q is free again and O_q holds all earlier tiles
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (__any_sync(0xffffffffu, grow)) {      // rare: rescale this warp's 32 rows of O_q in TMEM
            uint32_t o[64];
            tmem_ld32_issue(tO, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
            tmem_ld32_issue(tO + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 64; t++) o[t] = __float_as_uint(__uint_as_float(o[t]) * factor);
            tmem_st32(tO, &o[0]);
            tmem_st32(tO + 32, &o[32]);
          }
        }
        tmem_st32(tP, &pk[0]);
        tmem_st32(tP + 32, &pk[32]);
        tmem_st_wait();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) bar_arrive(&p_full[q]);
      }
      // ---- epilogue: O / l ----
      bar_wait(&pv_done[q], (n - 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const float inv = 1.0f / l;
      uint32_t o[64];
      tmem_ld32_issue(tO, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
      tmem_ld32_issue(tO + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
      tmem_ld_wait();
      if (row < p.Tq) {
        __half* dst = p.O + ((long long)b * p.Tq + row) * p.ldo + h * 64;
#pragma unroll
        for (int u = 0; u < 8; u++) {
          __align__(16) __half2 hv[4];
#pragma unroll
          for (int t = 0; t < 4; t++)
            hv[t] = __floats2half2_rn(__uint_as_float(o[u * 8 + 2 * t]) * inv, __uint_as_float(o[u * 8 + 2 * t + 1]) * inv);
          reinterpret_cast<uint4*>(dst)[u] = *reinterpret_cast<const uint4*>(hv);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

}  // namespace gdu
