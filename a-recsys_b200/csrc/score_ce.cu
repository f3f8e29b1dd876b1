// Full-catalog scoring fused with the softmax cross-entropy (K3 + K5 in one pass, and their adjoints)
// on the 5th-generation tensor cores.  Reference semantics: logits = U P^T + beta
// (attributes/embed_attribute.py:148-206 after the pool-first rewrite) followed by
// tf.nn.sparse_softmax_cross_entropy_with_logits (embed_attribute.py:530) and tf.gradients of both.
//
// The [M, N] logit matrix is never written to HBM (16 GB per step at M = 4096, N = 10^6):
//   forward : S tile = U_tile P_tile^T in TMEM -> online log-sum-exp per row -> lse[M]
//   backward: S tile recomputed -> D = g * (softmax - onehot) in registers -> tf32 operand tile in shared
//             memory -> second MMA accumulates dU = D P (rows = users) or dP = D^T U (rows = items) in TMEM.
// One CTA owns 128 rows of the "resident" operand and streams 64-row tiles of the other one:
//   warp 0 = TMA producer (resident tile once; streamed tiles and the second GEMM's B tiles through 2-stage rings)
//   warp 1 = TMEM allocator + MMA issuer (tcgen05.mma.cta_group::1.kind::tf32, S double-buffered in TMEM so that
//            the MMA of tile j+1 overlaps the epilogue of tile j)
//   warps 2..5 = epilogue, one TMEM lane (= row) per thread.
#include "arx_common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int CE_BM = 128;        // rows of the resident operand per CTA (TMEM lanes)
constexpr int CE_BN = 64;         // rows of the streamed operand per tile (= S columns)
constexpr int CE_THREADS = 192;
constexpr float kLog2e = 1.4426950408889634f;

enum { CE_FWD = 0, CE_BWD_U = 1, CE_BWD_P = 2, MW_FWD = 3, MW_BWD_U = 4, MW_BWD_P = 5 };
// MW_* = the same pipeline with the sampled WMRB loss (embed_attribute.py:641-649) in the epilogue:
//   hinge[r, c] = max(0, 1 + S[r, c] - t[r]) over the columns not masked for row r;  loss[r] = log(1 + sum_c hinge)
//   D[r, c] = g[r] / (1 + sum_c hinge[r, c]) * [hinge[r, c] > 0],  d loss / d t[r] = -sum_c D[r, c]
// The positives mask is a dense bit matrix [M, mask_ld] (bit = 1: column excluded), built per step by
// arx_mw_mask_build from the per-user CSR.
__host__ __device__ constexpr bool mode_fwd(int m) { return m == CE_FWD || m == MW_FWD; }
__host__ __device__ constexpr bool mode_rows_items(int m) { return m == CE_BWD_P || m == MW_BWD_P; }
__host__ __device__ constexpr bool mode_mw(int m) { return m >= MW_FWD; }

struct CeParams {
  long long R, S;                 // resident / streamed row counts
  int d;                          // GEMM1 depth = GEMM2 output width (multiple of 32, <= 128)
  int tiles_per_split;            // streamed tiles per blockIdx.y
  const float* beta;              // [N] item bias or nullptr
  const float* lse;               // [M] natural-log log-sum-exp per user row (backward)
  const float* g;                 // [M] d(total loss) / d(row loss)
  const int* tgt;                 // [M] target logit index
  float* part_m;                  // FWD: [nsplit][M] running max (log2 domain)
  float* part_l;                  // FWD: [nsplit][M] sum of 2^(t - max)
  float* out;                     // BWD: dU [M, d] or dP [N, d]
  float* dbeta;                   // BWD_P: [N] or nullptr
  int atomic_out;                 // several splits accumulate into a zeroed output
  // WMRB modes: lse holds the per-row hinge sums, and
  const float* ts;                // [M] target scores
  const uint32_t* mask;           // [M, mask_ld] excluded-column bits
  long long mask_ld;              // words per mask row (multiple of 4)
  float* dts;                     // MW_BWD_U: [M] d/d target score
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename T>
__device__ __forceinline__ T guarded(const T* p, long long i, long long n, T dflt) {
  return (p != nullptr && i < n) ? __ldg(p + i) : dflt;
}

template <int MODE>
__global__ void __launch_bounds__(CE_THREADS, 1)
ce_kernel(const __grid_constant__ CUtensorMap map_r, const __grid_constant__ CUtensorMap map_s,
          const __grid_constant__ CUtensorMap map_b2, const CeParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t r_full, sx_full[2], sx_empty[2], s_full[2], s_empty[2], d_full, d_empty,
      b2_full[2], b2_empty[2], acc_full;
  __shared__ uint32_t tmem_base_s;
  // per-column terms of the current S tile, staged once per tile by the epilogue warps (double-buffered):
  // FWD / BWD_U: [0] = beta[c] * log2(e) (-inf past the catalog); BWD_P: [0] = lse[c] * log2(e), [1] = g[c]
  // MW_BWD_P (single-buffered, guarded by a second barrier — the d = 128 configuration has no shared memory to
  // spare): [0] = 1 - t[c], [1] = g[c] / (1 + hsum[c]), then 4 mask words per column.
  __shared__ __align__(16) uint32_t s_colraw[6 * CE_BN];
  float (*s_cm)[2][CE_BN] = reinterpret_cast<float (*)[2][CE_BN]>(s_colraw);                 // [2][2][CE_BN]
  int (*s_ct)[CE_BN] = reinterpret_cast<int (*)[CE_BN]>(s_colraw + 4 * CE_BN);               // [2][CE_BN]  BWD_P: target[c]
  int (*s_cw)[4] = reinterpret_cast<int (*)[4]>(s_colraw + 2 * CE_BN);                       // [CE_BN][4]  MW_BWD_P

  const int KB = p.d >> 5;                                        // 32-wide k blocks of GEMM1
  const uint32_t r_slab = CE_BM * 128, s_slab = CE_BN * 128;      // [rows x 128 B] slabs
  const uint32_t r_bytes = (uint32_t)KB * r_slab, s_bytes = (uint32_t)KB * s_slab;
  const uint32_t dt_bytes = mode_fwd(MODE) ? 0u : 2u * r_slab;  // D tile: [128 x 64] = 2 slabs
  const uint32_t b2_slab = (uint32_t)p.d * 128;                   // [d rows x 128 B]
  const uint32_t b2_bytes = mode_fwd(MODE) ? 0u : 2u * b2_slab;
  uint8_t* sm_r = smem;
  uint8_t* sm_s = sm_r + r_bytes;                                 // 2 stages
  uint8_t* sm_d = sm_s + 2 * s_bytes;
  uint8_t* sm_b2 = sm_d + dt_bytes;                               // 2 stages

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r0 = (long long)blockIdx.x * CE_BM;
  const long long n_tiles_all = (p.S + CE_BN - 1) / CE_BN;
  const long long t0 = (long long)blockIdx.y * p.tiles_per_split;
  const int nt = (int)max(0ll, min((long long)p.tiles_per_split, n_tiles_all - t0));
  constexpr uint32_t kTmemCols = 256;
  constexpr uint32_t kAccCol = 2 * CE_BN;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&r_full), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&sx_full[s]), 1); mbar_init(smem_u32(&sx_empty[s]), 1);
      mbar_init(smem_u32(&s_full[s]), 1);  mbar_init(smem_u32(&s_empty[s]), 128);
      mbar_init(smem_u32(&b2_full[s]), 1); mbar_init(smem_u32(&b2_empty[s]), 1);
    }
    mbar_init(smem_u32(&d_full), 128); mbar_init(smem_u32(&d_empty), 1); mbar_init(smem_u32(&acc_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && nt > 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_r)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_s)) : "memory");
      if (!mode_fwd(MODE)) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b2)) : "memory");
      const uint32_t rf = smem_u32(&r_full);
      mbar_expect_tx(rf, r_bytes);
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(smem_u32(sm_r + kb * r_slab), &map_r, rf, kb * 32, (int)r0);
      for (int j = 0; j < nt; ++j) {
        const int s = j & 1;
        const uint32_t ph = (uint32_t)(j >> 1) & 1u;
        const int row = (int)((t0 + j) * CE_BN);
        mbar_wait(smem_u32(&sx_empty[s]), ph ^ 1u);
        const uint32_t fb = smem_u32(&sx_full[s]);
        mbar_expect_tx(fb, s_bytes);
        for (int kb = 0; kb < KB; ++kb)
          tma_load_2d(smem_u32(sm_s + s * s_bytes + kb * s_slab), &map_s, fb, kb * 32, row);
        if (!mode_fwd(MODE)) {
          mbar_wait(smem_u32(&b2_empty[s]), ph ^ 1u);
          const uint32_t bb = smem_u32(&b2_full[s]);
          mbar_expect_tx(bb, b2_bytes);
          for (int h = 0; h < 2; ++h)                              // box {32 streamed rows (K of GEMM2), d}
            tma_load_2d(smem_u32(sm_b2 + s * b2_bytes + h * b2_slab), &map_b2, bb, row + 32 * h, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && nt > 0) {
      const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(CE_BN >> 3) << 17) | ((uint32_t)(CE_BM >> 4) << 24);
      const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.d >> 3) << 17) | ((uint32_t)(CE_BM >> 4) << 24);
      mbar_wait(smem_u32(&r_full), 0);
      for (int j = 0; j <= nt; ++j) {
        if (j < nt) {
          const int s = j & 1;
          const uint32_t ph = (uint32_t)(j >> 1) & 1u;
          mbar_wait(smem_u32(&sx_full[s]), ph);
          mbar_wait(smem_u32(&s_empty[s]), ph ^ 1u);               // epilogue drained this TMEM buffer
          tc_fence_after();
          for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = make_desc(smem_u32(sm_r + kb * r_slab) + k * 32, 16, 1024);
              const uint64_t bd = make_desc(smem_u32(sm_s + s * s_bytes + kb * s_slab) + k * 32, 16, 1024);
              umma_tf32(tmem_base + (uint32_t)(s * CE_BN), ad, bd, idesc1, (kb > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(smem_u32(&sx_empty[s]));
          umma_commit(smem_u32(&s_full[s]));
        }
        if (!mode_fwd(MODE) && j >= 1) {
          const int jj = j - 1, s2 = jj & 1;
          mbar_wait(smem_u32(&d_full), (uint32_t)jj & 1u);
          mbar_wait(smem_u32(&b2_full[s2]), (uint32_t)(jj >> 1) & 1u);
          tc_fence_after();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = make_desc(smem_u32(sm_d + h * r_slab) + k * 32, 16, 1024);
              const uint64_t bd = make_desc(smem_u32(sm_b2 + s2 * b2_bytes + h * b2_slab) + k * 32, 16, 1024);
              umma_tf32(tmem_base + kAccCol, ad, bd, idesc2, (jj > 0 || h > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(smem_u32(&d_empty));
          umma_commit(smem_u32(&b2_empty[s2]));
        }
      }
      if (!mode_fwd(MODE)) umma_commit(smem_u32(&acc_full));
    }
  } else {
    // ===================== epilogue: one row (TMEM lane) per thread =====================
    const int q = warp & 3;
    const int rt = q * 32 + lane;                                  // row inside the tile
    const long long row = r0 + rt;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    // rows = users (FWD, BWD_U) or items (BWD_P)
    const long long M = mode_rows_items(MODE) ? p.S : p.R;
    const long long N = mode_rows_items(MODE) ? p.R : p.S;
    float m_run = -INFINITY, l_run = 0.f;                          // FWD
    float row_lse2 = 0.f, row_g = 0.f, row_beta = 0.f, dbeta_acc = 0.f;
    int row_tgt = -1;
    if (MODE == CE_BWD_U) {
      row_lse2 = guarded(p.lse, row, M, 0.f) * kLog2e; row_g = guarded(p.g, row, M, 0.f); row_tgt = guarded(p.tgt, row, M, -1);
    }
    if (mode_rows_items(MODE)) row_beta = guarded(p.beta, row, N, 0.f);
    // WMRB: 1 - t[row] (rows = users), g / (1 + hinge sum), running sums
    float row_1mt = 0.f, row_k = 0.f, mw_sum = 0.f;
    if (MODE == MW_FWD || MODE == MW_BWD_U) row_1mt = 1.f - guarded(p.ts, row, M, 0.f);
    if (MODE == MW_BWD_U) row_k = guarded(p.g, row, M, 0.f) / (1.f + guarded(p.lse, row, M, 0.f));

    const float row_beta2 = row_beta * kLog2e;
    for (int j = 0; j < nt; ++j) {
      const int s = j & 1;
      const long long c0 = (t0 + j) * CE_BN;                        // first streamed row (= S column) of the tile
      if (MODE == MW_BWD_P) asm volatile("bar.sync 1, 128;" ::: "memory");   // everyone is done with the previous tile's terms
      const int cs = (MODE == MW_BWD_P) ? 0 : s;                     // column-term stage
      {   // stage the per-column terms while the MMA of this tile runs (the loads were one LDG per element before)
        const int ci = rt & (CE_BN - 1);
        const long long c = c0 + ci;
        if (MODE == CE_BWD_P) {
          if (rt < CE_BN) {
            s_cm[s][0][ci] = (c < M) ? __ldg(p.lse + c) * kLog2e : 0.f;
            s_cm[s][1][ci] = (c < M) ? __ldg(p.g + c) : 0.f;         // g = 0 zeroes the columns past the batch
          } else {
            s_ct[s][ci] = (c < M) ? __ldg(p.tgt + c) : -1;
          }
        } else if (MODE == MW_BWD_P) {
          if (rt < CE_BN) {
            s_cm[0][0][ci] = (c < M) ? 1.f - __ldg(p.ts + c) : 0.f;
            s_cm[0][1][ci] = (c < M) ? __ldg(p.g + c) / (1.f + __ldg(p.lse + c)) : 0.f;
          } else {                                                   // mask bits of user c for this CTA's 128 item rows
            int4 w = make_int4(0, 0, 0, 0);
            if (c < M && p.mask != nullptr) w = __ldg(reinterpret_cast<const int4*>(p.mask + c * p.mask_ld + (r0 >> 5)));
            *reinterpret_cast<int4*>(&s_cw[ci][0]) = w;
          }
        } else if (mode_mw(MODE)) {
          if (rt < CE_BN) s_cm[s][0][ci] = (c < N) ? (p.beta != nullptr ? __ldg(p.beta + c) : 0.f) : -INFINITY;
        } else if (rt < CE_BN) {
          s_cm[s][0][ci] = (c < N) ? (p.beta != nullptr ? __ldg(p.beta + c) * kLog2e : 0.f) : -INFINITY;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");                 // the four epilogue warps only
      mbar_wait(smem_u32(&s_full[s]), (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      float v[CE_BN];
      tmem_ld32(lane_addr + (uint32_t)(s * CE_BN), v);
      tmem_ld32(lane_addr + (uint32_t)(s * CE_BN + 32), v + 32);
      tc_fence_before();
      mbar_arrive(smem_u32(&s_empty[s]));                          // TMEM buffer may be overwritten

      if (MODE == MW_FWD || MODE == MW_BWD_U) {
        uint32_t w0 = 0u, w1 = 0u;                                   // excluded-column bits of this row for the tile
        if (p.mask != nullptr && row < M) {
          const uint2 ww = __ldg(reinterpret_cast<const uint2*>(p.mask + row * p.mask_ld + (c0 >> 5)));
          w0 = ww.x; w1 = ww.y;
        }
#pragma unroll
        for (int i = 0; i < CE_BN; i += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(&s_cm[s][0][i]);
          const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t ex = (((i + q) < 32 ? w0 : w1) >> ((i + q) & 31)) & 1u;
            const float h = (v[i + q] + bb[q]) + row_1mt;              // -inf past the pool
            if (MODE == MW_FWD) {
              mw_sum += ex ? 0.f : fmaxf(h, 0.f);
            } else {
              const float x = (!ex && h > 0.f) ? row_k : 0.f;
              mw_sum += x;
              v[i + q] = tf32_rn(x);
            }
          }
        }
      } else if (MODE == MW_BWD_P) {
        const int wq = rt >> 5;                                      // mask word of my row inside the staged int4
        const uint32_t bit = 1u << (rt & 31);
#pragma unroll
        for (int i = 0; i < CE_BN; i += 4) {
          const float4 t4 = *reinterpret_cast<const float4*>(&s_cm[cs][0][i]);
          const float4 k4 = *reinterpret_cast<const float4*>(&s_cm[cs][1][i]);
          const float tt[4] = {t4.x, t4.y, t4.z, t4.w}, kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const bool ex = ((uint32_t)s_cw[i + q][wq] & bit) != 0u;
            const float h = (v[i + q] + row_beta) + tt[q];
            const float x = (!ex && h > 0.f) ? kk[q] : 0.f;
            dbeta_acc += x;
            v[i + q] = tf32_rn(x);
          }
        }
      } else if (MODE == CE_FWD) {
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < CE_BN; i += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(&s_cm[s][0][i]);
          v[i] = fmaf(v[i], kLog2e, b4.x); v[i + 1] = fmaf(v[i + 1], kLog2e, b4.y);
          v[i + 2] = fmaf(v[i + 2], kLog2e, b4.z); v[i + 3] = fmaf(v[i + 3], kLog2e, b4.w);
          mx = fmaxf(fmaxf(mx, fmaxf(v[i], v[i + 1])), fmaxf(v[i + 2], v[i + 3]));
        }
        const float m_new = fmaxf(m_run, mx);
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < CE_BN; ++i) acc += ex2(v[i] - m_new);
        l_run = l_run * ex2(m_run - m_new) + acc;
        m_run = m_new;
      } else if (MODE == CE_BWD_U) {
        const int rel = (row_tgt >= c0 && row_tgt < c0 + CE_BN) ? (int)(row_tgt - c0) : -1;
#pragma unroll
        for (int i = 0; i < CE_BN; i += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(&s_cm[s][0][i]);
          const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float e = ex2(fmaf(v[i + q], kLog2e, bb[q]) - row_lse2);       // 0 past the catalog (beta = -inf)
            v[i + q] = tf32_rn(row_g * (e - ((i + q) == rel ? 1.f : 0.f)));
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < CE_BN; i += 4) {
          const float4 l4 = *reinterpret_cast<const float4*>(&s_cm[s][0][i]);
          const float4 g4 = *reinterpret_cast<const float4*>(&s_cm[s][1][i]);
          const int4 t4 = *reinterpret_cast<const int4*>(&s_ct[s][i]);
          const float ll[4] = {l4.x, l4.y, l4.z, l4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
          const int tt[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float e = ex2(fmaf(v[i + q], kLog2e, row_beta2) - ll[q]);
            const float x = gg[q] * (e - ((long long)tt[q] == row ? 1.f : 0.f));
            dbeta_acc += x;
            v[i + q] = tf32_rn(x);
          }
        }
      }
      if (!mode_fwd(MODE)) {
        // D tile as the K-major, 128-byte-swizzled A operand of the second MMA
        mbar_wait(smem_u32(&d_empty), ((uint32_t)j & 1u) ^ 1u);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint8_t* slab = sm_d + h * r_slab + rt * 128;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch)
            *reinterpret_cast<float4*>(slab + ((ch ^ (rt & 7)) << 4)) =
                make_float4(v[h * 32 + ch * 4], v[h * 32 + ch * 4 + 1], v[h * 32 + ch * 4 + 2], v[h * 32 + ch * 4 + 3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(smem_u32(&d_full));
      }
    }

    if (MODE == MW_FWD) {
      if (row < M) p.part_l[(size_t)blockIdx.y * M + row] = mw_sum;
    } else if (MODE == CE_FWD) {
      if (row < M && nt > 0) {
        p.part_m[(size_t)blockIdx.y * M + row] = m_run;
        p.part_l[(size_t)blockIdx.y * M + row] = l_run;
      } else if (row < M) {
        p.part_m[(size_t)blockIdx.y * M + row] = -INFINITY;
        p.part_l[(size_t)blockIdx.y * M + row] = 0.f;
      }
    } else if (nt > 0) {
      mbar_wait(smem_u32(&acc_full), 0);
      tc_fence_after();
      const long long R = p.R;
      for (int cc = 0; cc < KB; ++cc) {
        float a[32];
        tmem_ld32(lane_addr + kAccCol + (uint32_t)(cc * 32), a);
        if (row < R) {
          float* dst = p.out + (size_t)row * p.d + cc * 32;
          if (p.atomic_out) {
            // split accumulation: 128-bit vector reductions (one L2 atomic per 16 bytes instead of one per float)
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(a[i]), "f"(a[i + 1]),
                           "f"(a[i + 2]), "f"(a[i + 3]) : "memory");
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(a[i], a[i + 1], a[i + 2], a[i + 3]);
          }
        }
      }
      if (MODE == MW_BWD_U && p.dts != nullptr && row < R) {
        if (p.atomic_out) atomicAdd(p.dts + row, -mw_sum); else p.dts[row] = -mw_sum;
      }
      if (mode_rows_items(MODE) && p.dbeta != nullptr && row < R) {
        if (p.atomic_out) atomicAdd(p.dbeta + row, dbeta_acc); else p.dbeta[row] = dbeta_acc;
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// lse[r] = ln sum_c exp(logit[r, c]) from the per-split (max, sum) pairs
__global__ void ce_finalize_kernel(const float* __restrict__ part_m, const float* __restrict__ part_l, int nsplit,
                                   long long M, float* __restrict__ lse) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  float m = -INFINITY;
  for (int s = 0; s < nsplit; ++s) m = fmaxf(m, part_m[(size_t)s * M + r]);
  float l = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float pm = part_m[(size_t)s * M + r];
    if (pm > -INFINITY) l += part_l[(size_t)s * M + r] * exp2f(pm - m);
  }
  lse[r] = (m + log2f(l)) / kLog2e;
}

// loss[r] = lse[r] - (U[r] . P[target[r]] + beta[target[r]]): one warp per row, same tf32-rounded operands
// as the tensor-core pass (products of tf32 values are exact in fp32, so only the summation order differs)
__global__ void ce_rowloss_kernel(const float* __restrict__ U, const float* __restrict__ P,
                                  const float* __restrict__ beta, const int* __restrict__ target,
                                  const float* __restrict__ lse, long long M, long long N, int d,
                                  float* __restrict__ loss) {
  const int lane = threadIdx.x & 31;
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= M) return;
  const int t = __ldg(target + r);
  float s = 0.f;
  if (t >= 0 && t < N)
    for (int c = lane; c < d; c += 32) s = fmaf(__ldg(U + r * d + c), __ldg(P + (long long)t * d + c), s);
  s = warp_sum(s);
  if (lane == 0) loss[r] = lse[r] - (s + ((beta != nullptr && t >= 0 && t < N) ? __ldg(beta + t) : 0.f));
}

// hsum[r] = sum over the splits; loss[r] = log(1 + hsum[r])  (embed_attribute.py:649)
__global__ void mw_finalize_kernel(const float* __restrict__ part, int nsplit, long long M, float* __restrict__ hsum,
                                   float* __restrict__ loss) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  float h = 0.f;
  for (int s = 0; s < nsplit; ++s) h += part[(size_t)s * M + r];
  hsum[r] = h;
  loss[r] = logf(1.f + h);
}

// bit (b, col) = 1 for every positive of batch row b's user that sits in the pool (col = pool position >= 0)
__global__ void mw_mask_build_kernel(const int* __restrict__ pos_row, const int* __restrict__ pos_ptr,
                                     const int* __restrict__ pos_idx, long long mb, long long N,
                                     uint32_t* __restrict__ mask, long long mask_ld) {
  const int lane = threadIdx.x & 31;
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= mb) return;
  const long long u = pos_row ? (long long)pos_row[b] : b;
  const int p0 = pos_ptr[u], p1 = pos_ptr[u + 1];
  for (int i = p0 + lane; i < p1; i += 32) {
    const int col = pos_idx[i];
    if (col >= 0 && col < N) atomicOr(mask + b * mask_ld + (col >> 5), 1u << (col & 31));
  }
}

size_t ce_smem_bytes(int mode, int d) {
  const size_t KB = d / 32;
  size_t b = KB * CE_BM * 128 + 2 * KB * CE_BN * 128;
  if (!mode_fwd(mode)) b += 2 * CE_BM * 128 + 2 * 2 * (size_t)d * 128;
  return b + 1024;
}

template <int MODE>
int ce_launch(const CUtensorMap& mr, const CUtensorMap& ms, const CUtensorMap& mb2, const CeParams& p, dim3 grid,
              cudaStream_t st) {
  static int configured = 0;
  const size_t smem = ce_smem_bytes(MODE, p.d);
  if ((int)smem > configured) {
    if (cudaFuncSetAttribute(ce_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return ARX_E_LAUNCH;
    configured = (int)smem;
  }
  ce_kernel<MODE><<<grid, CE_THREADS, smem, st>>>(mr, ms, mb2, p);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

bool ce_shape_ok(int64_t M, int64_t N, int64_t d) {
  return M > 0 && N > 0 && d >= 32 && d <= 128 && (d % 32) == 0 && M < (1ll << 31) && N < (1ll << 31);
}

// how many streamed tiles each CTA takes so that the grid is a few waves of the machine
int ce_split(long long row_tiles, long long stream_tiles, int* nsplit) {
  const long long sms = arx_num_sms();
  const long long want = 3ll * sms;
  long long ns = (want + row_tiles - 1) / row_tiles;
  // >= 8 streamed tiles per CTA amortise the resident-tile load; small problems (the 1024-item sampled pool)
  // go down to 4 and 2 tiles so that most SMs get a CTA (the kernel is latency-bound per tile there)
  long long min_tiles = 8;
  while (min_tiles > 2 && row_tiles * std::max(1ll, stream_tiles / min_tiles) < (3 * sms) / 4) min_tiles >>= 1;
  ns = std::max(1ll, std::min(ns, std::max(1ll, stream_tiles / min_tiles)));
  const long long tps = (stream_tiles + ns - 1) / ns;
  *nsplit = (int)((stream_tiles + tps - 1) / tps);
  return (int)tps;
}

}  // namespace

extern "C" int arx_ce_workspace_floats(int64_t M, int64_t N, int64_t* n_floats) {
  if (!n_floats || M <= 0 || N <= 0) return ARX_E_BADARG;
  int nsplit;
  ce_split((M + CE_BM - 1) / CE_BM, (N + CE_BN - 1) / CE_BN, &nsplit);
  *n_floats = 2ll * nsplit * M;
  return ARX_OK;
}

extern "C" int arx_ce_fwd(const float* U, const float* P, const float* beta, int64_t M, int64_t N, int64_t d,
                          float* workspace, float* lse, void* stream) {
  if (!U || !P || !workspace || !lse) return ARX_E_BADARG;
  if (!ce_shape_ok(M, N, d) || ((uintptr_t)U & 15) || ((uintptr_t)P & 15)) return ARX_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap mr, ms;
  if (!make_map(&mr, U, M, d, CE_BM) || !make_map(&ms, P, N, d, CE_BN)) return ARX_E_UNSUPPORTED;
  int nsplit;
  const long long row_tiles = (M + CE_BM - 1) / CE_BM;
  const int tps = ce_split(row_tiles, (N + CE_BN - 1) / CE_BN, &nsplit);
  CeParams p{};
  p.R = M; p.S = N; p.d = (int)d; p.tiles_per_split = tps; p.beta = beta;
  p.part_m = workspace; p.part_l = workspace + (size_t)nsplit * M;
  int rc = ce_launch<CE_FWD>(mr, ms, ms, p, dim3((unsigned)row_tiles, (unsigned)nsplit), st);
  if (rc != ARX_OK) return rc;
  ce_finalize_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(p.part_m, p.part_l, nsplit, M, lse);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_ce_rowloss(const float* U, const float* P, const float* beta, const int32_t* target,
                              const float* lse, int64_t M, int64_t N, int64_t d, float* loss, void* stream) {
  if (!U || !P || !target || !lse || !loss || M < 0 || N <= 0 || d <= 0) return ARX_E_BADARG;
  if (M == 0) return ARX_OK;
  ce_rowloss_kernel<<<(unsigned)((M + 7) / 8), 256, 0, (cudaStream_t)stream>>>(U, P, beta, target, lse, (long long)M,
                                                                             (long long)N, (int)d, loss);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_ce_bwd(const float* U, const float* P, const float* UT, const float* PT, const float* beta,
                          const float* lse, const float* g, const int32_t* target, int64_t M, int64_t N, int64_t d,
                          float* dU, float* dP, float* dbeta, void* stream) {
  if (!U || !P || !UT || !PT || !lse || !g || !target || !dU || !dP) return ARX_E_BADARG;
  if (!ce_shape_ok(M, N, d) || (M % 4) || (N % 4) || ((uintptr_t)U & 15) || ((uintptr_t)P & 15) ||
      ((uintptr_t)UT & 15) || ((uintptr_t)PT & 15))
    return ARX_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap mu128, mu64, mp128, mp64, mut, mpt;
  if (!make_map(&mu128, U, M, d, CE_BM) || !make_map(&mu64, U, M, d, CE_BN) || !make_map(&mp128, P, N, d, CE_BM) ||
      !make_map(&mp64, P, N, d, CE_BN) || !make_map(&mut, UT, d, M, (int)d) || !make_map(&mpt, PT, d, N, (int)d))
    return ARX_E_UNSUPPORTED;
  int rc;
  {   // dU = D P: rows = users, stream the catalog
    int nsplit;
    const long long row_tiles = (M + CE_BM - 1) / CE_BM;
    const int tps = ce_split(row_tiles, (N + CE_BN - 1) / CE_BN, &nsplit);
    CeParams p{};
    p.R = M; p.S = N; p.d = (int)d; p.tiles_per_split = tps; p.beta = beta; p.lse = lse; p.g = g; p.tgt = target;
    p.out = dU; p.atomic_out = nsplit > 1;
    if (p.atomic_out && cudaMemsetAsync(dU, 0, sizeof(float) * (size_t)M * d, st) != cudaSuccess) return ARX_E_LAUNCH;
    rc = ce_launch<CE_BWD_U>(mu128, mp64, mpt, p, dim3((unsigned)row_tiles, (unsigned)nsplit), st);
    if (rc != ARX_OK) return rc;
  }
  {   // dP = D^T U, dbeta = column sums of D: rows = items, stream the batch
    int nsplit;
    const long long row_tiles = (N + CE_BM - 1) / CE_BM;
    const int tps = ce_split(row_tiles, (M + CE_BN - 1) / CE_BN, &nsplit);
    CeParams p{};
    p.R = N; p.S = M; p.d = (int)d; p.tiles_per_split = tps; p.beta = beta; p.lse = lse; p.g = g; p.tgt = target;
    p.out = dP; p.dbeta = dbeta; p.atomic_out = nsplit > 1;
    if (p.atomic_out) {
      if (cudaMemsetAsync(dP, 0, sizeof(float) * (size_t)N * d, st) != cudaSuccess) return ARX_E_LAUNCH;
      if (dbeta && cudaMemsetAsync(dbeta, 0, sizeof(float) * (size_t)N, st) != cudaSuccess) return ARX_E_LAUNCH;
    }
    rc = ce_launch<CE_BWD_P>(mp128, mu64, mut, p, dim3((unsigned)row_tiles, (unsigned)nsplit), st);
  }
  return rc;
}

// ------------------------------------------------------------------ sampled WMRB ('mw') ----------
extern "C" int arx_mw_mask_words(int64_t N, int64_t* words_per_row) {
  if (!words_per_row || N <= 0) return ARX_E_BADARG;
  *words_per_row = ((N + CE_BM - 1) / CE_BM) * (CE_BM / 32);
  return ARX_OK;
}

extern "C" int arx_mw_mask_build(const int32_t* pos_row, const int32_t* pos_ptr, const int32_t* pos_idx, int64_t mb,
                                 int64_t N, uint32_t* mask, int64_t mask_ld, void* stream) {
  if (!pos_ptr || !pos_idx || !mask || mb < 0 || N <= 0 || mask_ld < (N + 31) / 32) return ARX_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (mb == 0) return ARX_OK;
  if (cudaMemsetAsync(mask, 0, sizeof(uint32_t) * (size_t)mb * mask_ld, st) != cudaSuccess) return ARX_E_LAUNCH;
  mw_mask_build_kernel<<<(unsigned)((mb + 7) / 8), 256, 0, st>>>(pos_row, pos_ptr, pos_idx, (long long)mb, (long long)N,
                                                                 mask, (long long)mask_ld);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_mw_fwd(const float* U, const float* P, const float* beta, const float* tscore, const uint32_t* mask,
                          int64_t mask_ld, int64_t M, int64_t N, int64_t d, float* workspace, float* hsum, float* loss,
                          void* stream) {
  if (!U || !P || !tscore || !workspace || !hsum || !loss) return ARX_E_BADARG;
  if (!ce_shape_ok(M, N, d) || ((uintptr_t)U & 15) || ((uintptr_t)P & 15) || (mask && (mask_ld % 4))) return ARX_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap mr, ms;
  if (!make_map(&mr, U, M, d, CE_BM) || !make_map(&ms, P, N, d, CE_BN)) return ARX_E_UNSUPPORTED;
  int nsplit;
  const long long row_tiles = (M + CE_BM - 1) / CE_BM;
  const int tps = ce_split(row_tiles, (N + CE_BN - 1) / CE_BN, &nsplit);
  CeParams p{};
  p.R = M; p.S = N; p.d = (int)d; p.tiles_per_split = tps; p.beta = beta; p.ts = tscore; p.mask = mask; p.mask_ld = mask_ld;
  p.part_m = workspace; p.part_l = workspace + (size_t)nsplit * M;
  int rc = ce_launch<MW_FWD>(mr, ms, ms, p, dim3((unsigned)row_tiles, (unsigned)nsplit), st);
  if (rc != ARX_OK) return rc;
  mw_finalize_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(p.part_l, nsplit, M, hsum, loss);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_mw_bwd2(const float* U, const float* P, const float* UT, const float* PT, const float* beta,
                           const float* tscore, const uint32_t* mask, int64_t mask_ld, const float* hsum, const float* g,
                           int64_t M, int64_t N, int64_t d, float* dU, float* dP, float* dbeta, float* dts,
                           int outputs_zeroed, void* stream);
extern "C" int arx_mw_bwd(const float* U, const float* P, const float* UT, const float* PT, const float* beta,
                          const float* tscore, const uint32_t* mask, int64_t mask_ld, const float* hsum, const float* g,
                          int64_t M, int64_t N, int64_t d, float* dU, float* dP, float* dbeta, float* dts, void* stream) {
  return arx_mw_bwd2(U, P, UT, PT, beta, tscore, mask, mask_ld, hsum, g, M, N, d, dU, dP, dbeta, dts, 0, stream);
}

// outputs_zeroed != 0: the caller has already zeroed dU, dP, dbeta and dts (e.g. with one memset at the top of the step,
// off the dependent chain): the split accumulation then starts without the four memsets in front of the kernels.
// outputs_zeroed == 2: ALWAYS add into the outputs (also when one split covers a tile): row blocks of a large batch
// accumulate their dP / dbeta into the same buffers (full-catalog WMRB, embed_attribute.py::fused_warp).
extern "C" int arx_mw_bwd2(const float* U, const float* P, const float* UT, const float* PT, const float* beta,
                           const float* tscore, const uint32_t* mask, int64_t mask_ld, const float* hsum, const float* g,
                           int64_t M, int64_t N, int64_t d, float* dU, float* dP, float* dbeta, float* dts,
                           int outputs_zeroed, void* stream) {
  if (!U || !P || !UT || !PT || !tscore || !hsum || !g || !dU || !dP || !dts) return ARX_E_BADARG;
  if (!ce_shape_ok(M, N, d) || (M % 4) || (N % 4) || ((uintptr_t)U & 15) || ((uintptr_t)P & 15) ||
      ((uintptr_t)UT & 15) || ((uintptr_t)PT & 15) || (mask && (mask_ld % 4)))
    return ARX_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap mu128, mu64, mp128, mp64, mut, mpt;
  if (!make_map(&mu128, U, M, d, CE_BM) || !make_map(&mu64, U, M, d, CE_BN) || !make_map(&mp128, P, N, d, CE_BM) ||
      !make_map(&mp64, P, N, d, CE_BN) || !make_map(&mut, UT, d, M, (int)d) || !make_map(&mpt, PT, d, N, (int)d))
    return ARX_E_UNSUPPORTED;
  int rc;
  {   // dU = D P, dts = -row sums of D: rows = users, stream the pool
    int nsplit;
    const long long row_tiles = (M + CE_BM - 1) / CE_BM;
    const int tps = ce_split(row_tiles, (N + CE_BN - 1) / CE_BN, &nsplit);
    CeParams p{};
    p.R = M; p.S = N; p.d = (int)d; p.tiles_per_split = tps; p.beta = beta; p.lse = hsum; p.g = g; p.ts = tscore;
    p.mask = mask; p.mask_ld = mask_ld; p.out = dU; p.dts = dts; p.atomic_out = nsplit > 1 || outputs_zeroed == 2;
    if (p.atomic_out && !outputs_zeroed) {
      if (cudaMemsetAsync(dU, 0, sizeof(float) * (size_t)M * d, st) != cudaSuccess) return ARX_E_LAUNCH;
      if (cudaMemsetAsync(dts, 0, sizeof(float) * (size_t)M, st) != cudaSuccess) return ARX_E_LAUNCH;
    }
    rc = ce_launch<MW_BWD_U>(mu128, mp64, mpt, p, dim3((unsigned)row_tiles, (unsigned)nsplit), st);
    if (rc != ARX_OK) return rc;
  }
  {   // dP = D^T U, dbeta = column sums of D: rows = pool items, stream the batch
    int nsplit;
    const long long row_tiles = (N + CE_BM - 1) / CE_BM;
    const int tps = ce_split(row_tiles, (M + CE_BN - 1) / CE_BN, &nsplit);
    CeParams p{};
    p.R = N; p.S = M; p.d = (int)d; p.tiles_per_split = tps; p.beta = beta; p.lse = hsum; p.g = g; p.ts = tscore;
    p.mask = mask; p.mask_ld = mask_ld; p.out = dP; p.dbeta = dbeta; p.atomic_out = nsplit > 1 || outputs_zeroed == 2;
    if (p.atomic_out && !outputs_zeroed) {
      if (cudaMemsetAsync(dP, 0, sizeof(float) * (size_t)N * d, st) != cudaSuccess) return ARX_E_LAUNCH;
      if (dbeta && cudaMemsetAsync(dbeta, 0, sizeof(float) * (size_t)N, st) != cudaSuccess) return ARX_E_LAUNCH;
    }
    rc = ce_launch<MW_BWD_P>(mp128, mu64, mut, p, dim3((unsigned)row_tiles, (unsigned)nsplit), st);
  }
  return rc;
}
