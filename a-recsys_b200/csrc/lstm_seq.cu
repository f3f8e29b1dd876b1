// LSTM recurrence as ONE persistent kernel per direction (K8): lstm/seqModel.py:99-103,477 — TF-1.0 LSTMCell
// (gate order i, j, f, o on [x, h] W + b; c' = sigmoid(f + forget_bias) c + sigmoid(i) tanh(j); h' = sigmoid(o) tanh(c'))
// statically unrolled over T steps from a zero state.
//
// The x-projection of all T steps is one large tensor-core contraction done beforehand (gemm_tc.cu); what is left is
// the sequential part, h_{t-1} W_h per step, which the round-1 code ran as one GEMM launch + one SIMT gate kernel per
// time step (2 T launches per direction).  Here a CLUSTER of NC = H / 32 CTAs owns 128 batch rows for all T steps:
//   * CTA `rank` holds the W_h columns of ITS 32 hidden units (all four gates: a [128 x H] tf32 B operand, K-major,
//     128-byte swizzled) resident in shared memory for the whole kernel, loaded once by TMA;
//   * per step one thread issues tcgen05.mma (M 128 x N 128 x K H, kind::tf32) from the h_{t-1} tile in shared memory
//     into TMEM; eight epilogue warps read the accumulator with tcgen05.ld, add the x-projection tile (TMA-prefetched
//     one step ahead into shared memory), apply the gate non-linearities, keep c in registers, write the activated
//     gates / c / h to HBM for the backward pass, and write h_t (tf32-rounded) straight into the A-operand tile of
//     EVERY CTA of the cluster through distributed shared memory (st.shared::cluster) — the all-gather of the hidden
//     state never leaves the SMs;
//   * CTAs synchronise per step with cluster-scope mbarriers only (no cluster.sync on the step path).
// The backward kernel mirrors it with a K-split: CTA `rank` contracts its own dZ columns with the matching W_h slice
// into a partial dh tile in TMEM, and the partial tiles are reduce-scattered through distributed shared memory.
#include "arx_common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int LS_BM = 128;            // batch rows per cluster (UMMA M, TMEM lanes)
constexpr int LS_UC = 32;             // hidden units per CTA
constexpr int LS_THREADS = 320;       // warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = epilogue
constexpr int LS_EPI_WARPS = 8;
constexpr uint32_t LS_SLAB = LS_BM * 128;      // [128 rows x 128 B] = one 32-wide k block of an operand tile

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t caddr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(caddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
// 16-byte store into (possibly remote) shared memory that signals completion on an mbarrier of the DESTINATION CTA by
// transaction bytes: data and signal travel together, so the producer needs no fence and no separate arrive (a
// release-arrive at cluster scope compiles to MEMBAR.GPU, which also waits for every outstanding global store).
__device__ __forceinline__ void st_async_f4(uint32_t caddr, float4 v, uint32_t cbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(caddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(cbar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t caddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(caddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT_C:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE_C;\n"
      "bra LAB_WAIT_C;\n"
      "DONE_C:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Explicit shared-space accesses: through a generic pointer the compiler emits LD.E / ST.E, which take the slow
// generic path and are NOT ordered with a following mbarrier arrive by the shared-memory pipeline.
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f4(uint32_t saddr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// 16 consecutive fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// MUFU-based non-linearities (ex2 + rcp): ~2 ulp, far inside the tf32 operand rounding of this path
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

struct LstmSeqParams {
  float* G;            // [T, mb, 4H]  fwd: x-projection + bias in, activated gates out;  bwd: gates in, dZ out
  float* Hs;           // [T+1, mb, H] fwd out (slot t+1 = h_t; slot 0 = zero state, not touched)
  float* Cs;           // [T+1, mb, H] fwd out / bwd in
  const float* dH;     // bwd: [T, mb, H] gradient w.r.t. the outputs h_t
  float* dX0;          // unused
  long long mb;
  int T;
  float forget_bias;
  int reserved;
};

// ------------------------------------------------------------------ forward ---------------------------------
template <int NC>
__global__ void __launch_bounds__(LS_THREADS, 1)
lstm_seq_fwd_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_zx,
                    const LstmSeqParams p) {
  constexpr int H = NC * LS_UC;
  constexpr int KB = NC;                                    // 32-wide k blocks of the h W_h contraction
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;                                       // [KB][128 gate columns][128 B]   B operand (resident)
  uint8_t* sA = sW + KB * LS_SLAB;                          // [KB][128 rows][128 B]           A operand = h_{t-1}
  uint8_t* sZ = sA + KB * LS_SLAB;                          // [4 gates][128 rows][128 B]      x-projection tile of step t
  __shared__ __align__(8) uint64_t w_full, zx_full, zx_empty, a_ready, acc_full, a_free;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (NC > 1) ? cluster_ctarank() : 0u;
  const long long row0 = (long long)(blockIdx.x / NC) * LS_BM;
  const int T = p.T;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&w_full), 1);
    mbar_init(smem_u32(&zx_full), 1);
    mbar_init(smem_u32(&zx_empty), LS_EPI_WARPS);
    mbar_init(smem_u32(&a_ready), 1);                       // one arrive (+ expect_tx) by the MMA thread; data by st.async bytes
    mbar_init(smem_u32(&acc_full), 1);
    mbar_init(smem_u32(&a_free), NC > 1 ? NC - 1 : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (NC > 1) cluster_sync_all();                           // every CTA's barriers exist before any remote arrive
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===================== TMA producer: W_h slice once, then the x-projection tile of every step ===========
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_zx)) : "memory");
      const uint32_t wf = smem_u32(&w_full);
      mbar_expect_tx(wf, KB * LS_SLAB);
      for (int kb = 0; kb < KB; ++kb)
        for (int g = 0; g < 4; ++g)       // box {32 k, 32 gate columns}: rows g*32.. of the k block's slab
          tma_load_2d(smem_u32(sW + kb * LS_SLAB + g * 32 * 128), &map_w, wf, kb * 32, g * H + (int)rank * LS_UC);
      for (int t = 0; t < T; ++t) {
        if (t > 0) mbar_wait(smem_u32(&zx_empty), (uint32_t)(t - 1) & 1u);
        const uint32_t zf = smem_u32(&zx_full);
        mbar_expect_tx(zf, 4 * LS_SLAB);
        for (int g = 0; g < 4; ++g)       // box {32 columns, 128 rows}
          tma_load_2d(smem_u32(sZ + g * LS_SLAB), &map_zx, zf, g * H + (int)rank * LS_UC,
                      (int)((long long)t * p.mb + row0));
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: Z_t = h_{t-1} W_h[:, my gate columns] =================================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(LS_BM >> 4) << 24);
      mbar_wait(smem_u32(&w_full), 0);
      for (int t = 1; t < T; ++t) {
        mbar_expect_tx(smem_u32(&a_ready), (uint32_t)KB * LS_SLAB);        // h_{t-1}: NC slices of 16 KB, sent with st.async
        mbar_wait_cluster(smem_u32(&a_ready), (uint32_t)(t - 1) & 1u);
        fence_async_smem();
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = make_desc(smem_u32(sA + kb * LS_SLAB) + k * 32, 16, 1024);
            const uint64_t bd = make_desc(smem_u32(sW + kb * LS_SLAB) + k * 32, 16, 1024);
            umma_tf32(tmem_base, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(smem_u32(&acc_full));
        if (NC > 1) {
          // my reads of sA are over once the MMAs retire: tell the peers they may overwrite it with h_t
          mbar_wait(smem_u32(&acc_full), (uint32_t)(t - 1) & 1u);
          for (uint32_t r = 0; r < (uint32_t)NC; ++r)
            if (r != rank) mbar_arrive_cluster(map_to_cta(smem_u32(&a_free), r));
        }
      }
    }
  } else {
    // ===================== epilogue: one batch row x 16 hidden units per thread =============================
    const int q = warp & 3;                                  // TMEM lane quarter this warp may access
    const int uh = (warp - 2) >> 2;                          // which 16 of the CTA's 32 units
    const int rt = q * 32 + lane;                            // row inside the tile
    const long long row = row0 + rt;
    const bool row_ok = row < p.mb;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int ucol = (int)rank * LS_UC + uh * 16;            // first hidden unit of this thread (global index)
    float c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = 0.f;
    uint32_t a_dst[NC], a_bar[NC];                           // my 64-byte h segment inside every CTA's sA, and its barrier
#pragma unroll
    for (int r = 0; r < NC; ++r) {
      a_dst[r] = map_to_cta(smem_u32(sA + rank * LS_SLAB + rt * 128), (uint32_t)r);
      a_bar[r] = map_to_cta(smem_u32(&a_ready), (uint32_t)r);
    }

    for (int t = 0; t < T; ++t) {
      // x-projection (+ bias) of my 4 x 16 gate columns from the TMA-staged, 128-byte-swizzled tile
      float z[4][16];
      mbar_wait(smem_u32(&zx_full), (uint32_t)t & 1u);
      const uint32_t zrow0 = smem_u32(sZ) + (uint32_t)rt * 128u;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 v = lds_f4(zrow0 + (uint32_t)g * LS_SLAB + (uint32_t)(((uh * 4 + c4) ^ (rt & 7)) << 4));
          z[g][c4 * 4] = v.x; z[g][c4 * 4 + 1] = v.y; z[g][c4 * 4 + 2] = v.z; z[g][c4 * 4 + 3] = v.w;
        }
      }
      if (t > 0) {
        mbar_wait(smem_u32(&acc_full), (uint32_t)(t - 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float a[16];
          tmem_ld16(lane_addr + (uint32_t)(g * 32 + uh * 16), a);
#pragma unroll
          for (int i = 0; i < 16; ++i) z[g][i] += a[i];
        }
        tc_fence_before();
      } else {
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int i = 0; i < 16; ++i) z[g][i] += 0.f * z[g][i];      // consume the staged values (see below)
      }
      // Release the staging tile only AFTER every staged value has been consumed by an arithmetic instruction: a load
      // that is merely issued may still be in flight when the arrive executes, and the TMA write of step t + 1 then
      // overtakes it (seen on hardware as a rare wrong 16-byte chunk of one gate).
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&zx_empty));       // the producer may stage step t + 1
      float h[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float gi = fast_sigmoid(z[0][i]);
        const float gj = fast_tanh(z[1][i]);
        const float gf = fast_sigmoid(z[2][i] + p.forget_bias);
        const float go = fast_sigmoid(z[3][i]);
        c[i] = gf * c[i] + gi * gj;
        h[i] = tf32_rn(go * fast_tanh(c[i]));                // h only feeds tensor-core contractions: kept tf32-exact
        z[0][i] = gi; z[1][i] = gj; z[2][i] = gf; z[3][i] = go;
      }
      if (t + 1 < T) {
        // h_t -> the A tile of every CTA of the cluster (k block = my rank, 16-byte chunks uh*4 .. uh*4+3), BEFORE the
        // HBM stores below: the exchange is on the recurrence's critical path, the stores are not
        if (NC > 1 && t > 0) mbar_wait_cluster(smem_u32(&a_free), (uint32_t)(t - 1) & 1u);   // peers' MMAs of step t retired
#pragma unroll
        for (int r = 0; r < NC; ++r) {
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const float4 v = row_ok ? make_float4(h[c4 * 4], h[c4 * 4 + 1], h[c4 * 4 + 2], h[c4 * 4 + 3]) : f4_zero();
            st_async_f4(a_dst[r] + (uint32_t)(((uh * 4 + c4) ^ (rt & 7)) << 4), v, a_bar[r]);
          }
        }
      }
      if (row_ok) {
        float* gdst = p.G + ((size_t)t * p.mb + row) * (4 * H) + ucol;
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4)
            st_f4(gdst + g * H + c4 * 4, make_float4(z[g][c4 * 4], z[g][c4 * 4 + 1], z[g][c4 * 4 + 2], z[g][c4 * 4 + 3]));
        float* hdst = p.Hs + ((size_t)(t + 1) * p.mb + row) * H + ucol;
        float* cdst = p.Cs + ((size_t)(t + 1) * p.mb + row) * H + ucol;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          st_f4(hdst + c4 * 4, make_float4(h[c4 * 4], h[c4 * 4 + 1], h[c4 * 4 + 2], h[c4 * 4 + 3]));
          st_f4(cdst + c4 * 4, make_float4(c[c4 * 4], c[c4 * 4 + 1], c[c4 * 4 + 2], c[c4 * 4 + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (NC > 1) cluster_sync_all();                             // no CTA leaves while a peer may still signal it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// ------------------------------------------------------------------ backward --------------------------------
// Per step (t = T-1 .. 0), per CTA (units U = rank*32 .. +32, batch rows of the cluster's tile):
//   dh_t[U]  = dH[t][U] + sum over the cluster of partial_{t+1}[., U]          (reduce-scatter through DSMEM)
//   dc       = dh o (1 - tanh(c_t)^2) + dc_next ;  dZ = (dc j i(1-i), dc i (1-j^2), dc c_{t-1} f(1-f), dh tanh(c_t) o(1-o))
//   partial_t = dZ_t[., my 128 gate columns] * W_h[:, my gate columns]^T   -> [128 rows x H] in TMEM
template <int NC>
__global__ void __launch_bounds__(LS_THREADS, 1)
lstm_seq_bwd_kernel(const __grid_constant__ CUtensorMap map_w, const LstmSeqParams p) {
  constexpr int H = NC * LS_UC;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr uint32_t kWSlab = H * 128;                       // [H units x 128 B] per gate
  uint8_t* sW = smem;                                        // [4 gates][H][128 B]       B operand: W_h[n, g*H + rank*32 + k]
  uint8_t* sA = sW + 4 * kWSlab;                             // [4 gates][128 rows][128 B] A operand: my dZ_t columns
  float* sR = reinterpret_cast<float*>(sA + 4 * LS_SLAB);    // [2][NC-1][128 rows][32]    partial dh tiles from the peers
  __shared__ __align__(8) uint64_t w_full, a_ready, acc_full, recv_full[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (NC > 1) ? cluster_ctarank() : 0u;
  const long long row0 = (long long)(blockIdx.x / NC) * LS_BM;
  const int T = p.T;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&w_full), 1);
    mbar_init(smem_u32(&a_ready), LS_EPI_WARPS);
    mbar_init(smem_u32(&acc_full), 1);
    mbar_init(smem_u32(&recv_full[0]), 1);                  // one arrive (+ expect_tx) locally; data by st.async bytes
    mbar_init(smem_u32(&recv_full[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (NC > 1) cluster_sync_all();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
      const uint32_t wf = smem_u32(&w_full);
      mbar_expect_tx(wf, 4 * kWSlab);
      for (int g = 0; g < 4; ++g)           // box {32 gate columns (K), H units (N)}
        tma_load_2d(smem_u32(sW + g * kWSlab), &map_w, wf, g * H + (int)rank * LS_UC, 0);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(H >> 3) << 17) | ((uint32_t)(LS_BM >> 4) << 24);
      mbar_wait(smem_u32(&w_full), 0);
      for (int it = 0; it + 1 < T; ++it) {
        mbar_wait(smem_u32(&a_ready), (uint32_t)it & 1u);
        fence_async_smem();
        tc_fence_after();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = make_desc(smem_u32(sA + g * LS_SLAB) + k * 32, 16, 1024);
            const uint64_t bd = make_desc(smem_u32(sW + g * kWSlab) + k * 32, 16, 1024);
            umma_tf32(tmem_base, ad, bd, idesc, (g > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(smem_u32(&acc_full));
      }
    }
  } else {
    const int q = warp & 3;
    const int uh = (warp - 2) >> 2;
    const int rt = q * 32 + lane;
    const long long row = row0 + rt;
    const bool row_ok = row < p.mb;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int ucol = (int)rank * LS_UC + uh * 16;
    float dc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) dc[i] = 0.f;

    for (int it = 0; it < T; ++it) {
      const int t = T - 1 - it;
      if (t > 0 && row_ok) {
        // warm L2 with the next step's operands while this one computes (that step's loads then hit L2)
        const size_t rn = (size_t)(t - 1) * p.mb + row;
#pragma unroll
        for (int g = 0; g < 4; ++g) prefetch_l2(p.G + rn * (4 * H) + g * H + ucol);
        prefetch_l2(p.dH + rn * H + ucol);
        prefetch_l2(p.Cs + rn * H + ucol);
      }
      float dh[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) dh[i] = 0.f;
      if (it > 0) {
        // ---- reduce-scatter of the partial dh tiles of step t + 1 ------------------------------------------
        mbar_wait(smem_u32(&acc_full), (uint32_t)(it - 1) & 1u);
        tc_fence_after();
        const int buf = (it - 1) & 1;
#pragma unroll
        for (int r = 0; r < NC; ++r) {
          float a[16];
          tmem_ld16(lane_addr + (uint32_t)(r * LS_UC + uh * 16), a);       // columns of CTA r's units
          if (r == (int)rank) {
#pragma unroll
            for (int i = 0; i < 16; ++i) dh[i] += a[i];
          } else {
            const int slot = (int)rank - ((int)rank > r ? 1 : 0);          // my slot among r's NC-1 senders
            const uint32_t dst = map_to_cta(smem_u32(sR + ((size_t)(buf * (NC - 1) + slot) * LS_BM + rt) * LS_UC + uh * 16),
                                            (uint32_t)r);
            const uint32_t dbar = map_to_cta(smem_u32(&recv_full[buf]), (uint32_t)r);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4)
              st_async_f4(dst + c4 * 16, make_float4(a[c4 * 4], a[c4 * 4 + 1], a[c4 * 4 + 2], a[c4 * 4 + 3]), dbar);
          }
        }
        tc_fence_before();
        if (NC > 1) {
          if (warp == 2 && lane == 0)                       // the peers' (NC - 1) x 16 KB partial tiles of this step
            mbar_expect_tx(smem_u32(&recv_full[buf]), (uint32_t)(NC - 1) * LS_BM * LS_UC * 4u);
          mbar_wait_cluster(smem_u32(&recv_full[buf]), (uint32_t)((it - 1) >> 1) & 1u);
#pragma unroll
          for (int s = 0; s < NC - 1; ++s) {                               // fixed sender order: deterministic
            const uint32_t src = smem_u32(sR + ((size_t)(buf * (NC - 1) + s) * LS_BM + rt) * LS_UC + uh * 16);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const float4 v = lds_f4(src + c4 * 16);
              dh[c4 * 4] += v.x; dh[c4 * 4 + 1] += v.y; dh[c4 * 4 + 2] += v.z; dh[c4 * 4 + 3] += v.w;
            }
          }
        }
      }
      // ---- gate adjoints ---------------------------------------------------------------------------------
      float dz[4][16];
      if (row_ok) {
        const size_t rb = (size_t)t * p.mb + row;
        const float* gsrc = p.G + rb * (4 * H) + ucol;
        const float* dhs = p.dH + rb * H + ucol;
        const float* cs = p.Cs + (rb + (size_t)p.mb) * H + ucol;           // c_t   (slot t + 1)
        const float* cps = p.Cs + rb * H + ucol;                          // c_{t-1} (slot t; zeros at t = 0)
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 gi4 = ld_f4(gsrc + c4 * 4), gj4 = ld_f4(gsrc + H + c4 * 4), gf4 = ld_f4(gsrc + 2 * H + c4 * 4),
                       go4 = ld_f4(gsrc + 3 * H + c4 * 4), dh4 = ld_f4(dhs + c4 * 4), c4v = ld_f4(cs + c4 * 4);
          float4 cp4 = f4_zero();
          if (t > 0) cp4 = ld_f4(cps + c4 * 4);
          const float gi[4] = {gi4.x, gi4.y, gi4.z, gi4.w}, gj[4] = {gj4.x, gj4.y, gj4.z, gj4.w},
                      gf[4] = {gf4.x, gf4.y, gf4.z, gf4.w}, go[4] = {go4.x, go4.y, go4.z, go4.w},
                      dho[4] = {dh4.x, dh4.y, dh4.z, dh4.w}, cc[4] = {c4v.x, c4v.y, c4v.z, c4v.w},
                      cp[4] = {cp4.x, cp4.y, cp4.z, cp4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = c4 * 4 + e;
            const float dht = dh[i] + dho[e];
            const float tc = fast_tanh(cc[e]);
            const float dct = dht * go[e] * (1.0f - tc * tc) + dc[i];
            float g0 = dct * gj[e] * gi[e] * (1.0f - gi[e]);
            float g1 = dct * gi[e] * (1.0f - gj[e] * gj[e]);
            float g2 = dct * cp[e] * gf[e] * (1.0f - gf[e]);
            float g3 = dht * tc * go[e] * (1.0f - go[e]);
            dc[i] = dct * gf[e];
            dz[0][i] = tf32_rn(g0); dz[1][i] = tf32_rn(g1); dz[2][i] = tf32_rn(g2); dz[3][i] = tf32_rn(g3);
          }
        }
      } else {
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int i = 0; i < 16; ++i) dz[g][i] = 0.f;
      }
      if (it + 1 < T) {
        // my dZ columns -> the A tile (k block = gate g, chunks uh*4 ..): the previous MMA has retired (acc_full waited)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t arow = smem_u32(sA + g * LS_SLAB + rt * 128);
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4)
            sts_f4(arow + (uint32_t)(((uh * 4 + c4) ^ (rt & 7)) << 4),
                   make_float4(dz[g][c4 * 4], dz[g][c4 * 4 + 1], dz[g][c4 * 4 + 2], dz[g][c4 * 4 + 3]));
        }
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&a_ready));
      }
      if (row_ok) {                                           // dZ for the three large contractions: off the critical path
        float* gdst = p.G + ((size_t)t * p.mb + row) * (4 * H) + ucol;
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4)
            st_f4(gdst + g * H + c4 * 4, make_float4(dz[g][c4 * 4], dz[g][c4 * 4 + 1], dz[g][c4 * 4 + 2], dz[g][c4 * 4 + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (NC > 1) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// 2-D fp32 tensor map with an explicit box (128-byte swizzle, 32 floats wide)
bool make_map_box(CUtensorMap* map, const float* base, long long rows, long long cols, int box_rows) {
  return make_map(map, base, rows, cols, box_rows);
}

template <int NC>
int launch_fwd(const CUtensorMap& mw, const CUtensorMap& mz, const LstmSeqParams& p, cudaStream_t st) {
  const size_t smem = (size_t)(2 * NC + 4) * LS_SLAB + 1024;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(lstm_seq_fwd_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return ARX_E_LAUNCH;
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(NC * ((p.mb + LS_BM - 1) / LS_BM)));
  cfg.blockDim = dim3(LS_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, lstm_seq_fwd_kernel<NC>, mw, mz, p) != cudaSuccess) return ARX_E_LAUNCH;
  return ARX_OK;
}

template <int NC>
int launch_bwd(const CUtensorMap& mw, const LstmSeqParams& p, cudaStream_t st) {
  constexpr int H = NC * LS_UC;
  const size_t smem = (size_t)4 * H * 128 + 4 * LS_SLAB + (size_t)2 * (NC - 1) * LS_BM * LS_UC * 4 + 1024;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(lstm_seq_bwd_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return ARX_E_LAUNCH;
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(NC * ((p.mb + LS_BM - 1) / LS_BM)));
  cfg.blockDim = dim3(LS_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, lstm_seq_bwd_kernel<NC>, mw, p) != cudaSuccess) return ARX_E_LAUNCH;
  return ARX_OK;
}

bool seq_shape_ok(int64_t T, int64_t mb, int H) {
  return T >= 1 && mb >= 1 && (H == 32 || H == 64 || H == 128) && T * mb < (1ll << 31);
}

}  // namespace

// Forward recurrence of all T steps in one launch.  G [T, mb, 4H]: x-projection + bias on entry, ACTIVATED gates
// (i, j, f, o) on return; WhT [4H, H] = W_h^T, tf32-rounded; Hs / Cs [T+1, mb, H]: slots 1..T are written (slot 0 is the
// caller's zero state).  h is stored tf32-rounded (it only feeds tensor-core contractions on this path).
// ARX_E_UNSUPPORTED unless H is 32, 64 or 128 (the caller then runs the per-step kernels).
extern "C" int arx_lstm_seq_fwd(float* G, const float* WhT, float* Hs, float* Cs, int64_t T, int64_t mb, int H,
                                float forget_bias, void* stream) {
  if (!G || !WhT || !Hs || !Cs || T < 0 || mb < 0 || H < 1) return ARX_E_BADARG;
  if (T == 0 || mb == 0) return ARX_OK;
  if (!seq_shape_ok(T, mb, H) || ((uintptr_t)G & 15) || ((uintptr_t)WhT & 15) || ((uintptr_t)Hs & 15) || ((uintptr_t)Cs & 15))
    return ARX_E_UNSUPPORTED;
  CUtensorMap mw, mz;
  if (!make_map(&mw, WhT, 4ll * H, H, 32)) return ARX_E_UNSUPPORTED;
  if (!make_map(&mz, G, T * mb, 4ll * H, LS_BM)) return ARX_E_UNSUPPORTED;
  LstmSeqParams p{G, Hs, Cs, nullptr, nullptr, (long long)mb, (int)T, forget_bias, 0};
  cudaStream_t st = (cudaStream_t)stream;
  if (H == 32) return launch_fwd<1>(mw, mz, p, st);
  if (H == 64) return launch_fwd<2>(mw, mz, p, st);
  return launch_fwd<4>(mw, mz, p, st);
}

// Backward recurrence of all T steps in one launch.  G [T, mb, 4H]: activated gates on entry, dZ (gradient w.r.t. the
// pre-activations, tf32-rounded) on return; Wh [H, 4H] = W_h, tf32-rounded; Cs as written by arx_lstm_seq_fwd;
// dH [T, mb, H] = gradient w.r.t. the step outputs h_t.
extern "C" int arx_lstm_seq_bwd(float* G, const float* Wh, const float* Cs, const float* dH, int64_t T, int64_t mb,
                                int H, void* stream) {
  if (!G || !Wh || !Cs || !dH || T < 0 || mb < 0 || H < 1) return ARX_E_BADARG;
  if (T == 0 || mb == 0) return ARX_OK;
  if (!seq_shape_ok(T, mb, H) || ((uintptr_t)G & 15) || ((uintptr_t)Wh & 15) || ((uintptr_t)Cs & 15) || ((uintptr_t)dH & 15))
    return ARX_E_UNSUPPORTED;
  CUtensorMap mw;
  if (!make_map(&mw, Wh, H, 4ll * H, H)) return ARX_E_UNSUPPORTED;
  LstmSeqParams p{G, nullptr, const_cast<float*>(Cs), dH, nullptr, (long long)mb, (int)T, 0.f, 1};
  cudaStream_t st = (cudaStream_t)stream;
  if (H == 32) return launch_bwd<1>(mw, p, st);
  if (H == 64) return launch_bwd<2>(mw, p, st);
  return launch_bwd<4>(mw, p, st);
}
