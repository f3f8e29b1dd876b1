// Dense contractions of the hot path on the 5th-generation tensor cores:
//   tcgen05.mma.cta_group::1.kind::tf32 (fp32 operands read as tf32, fp32 accumulate in TMEM),
//   operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) through a 4-stage
//   mbarrier pipeline, accumulator read back with tcgen05.ld for the epilogue.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (one TMEM lane quarter each).
//
// Serves K3 (scores = U * P^T, attributes/embed_attribute.py:171,188 after the pool-first
// rewrite), its adjoints (dU = D * P, dP = D^T * U) and the LSTM gate GEMM
// (lstm/seqModel.py:99, [x,h] * W).  tf32 keeps the 1e-3 parity budget of the north star
// (10-bit mantissa operands, fp32 accumulation); the exact-fp32 SIMT kernel in score.cu stays
// as the parity anchor and the fallback for shapes TMA cannot describe.
#include "arx_common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int BM = 128;          // UMMA M (TMEM lanes)
constexpr int BK = 32;           // tf32 elements per stage row = 128 bytes = one swizzle span
constexpr int UMMA_K = 8;        // tf32: 32 bytes per instruction
constexpr int kStages = 4;
constexpr int kThreads = 192;

struct GemmParams {
  float* C;
  const float* bias_n;
  long long M, N, K, ldc;
  long long k_per_split;             // multiple of BK
  float alpha, beta;
  int atomic_out;                    // split-K: red.add into a zeroed C
};

// A_MN / B_MN: operand stored with the M (resp. N) index contiguous ("MN-major") instead of K.
template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t kABytes = BM * BK * 4;          // 16 KB
  constexpr uint32_t kBBytes = BN * BK * 4;
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], tmem_full_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_epi[4][32 * 33];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long m0 = (long long)blockIdx.y * BM, n0 = (long long)blockIdx.x * BN;
  const long long k_begin = (long long)blockIdx.z * p.k_per_split;
  const long long k_end = min(p.K, k_begin + p.k_per_split);
  const int num_kb = (int)((k_end - k_begin + BK - 1) / BK);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(&tmem_full_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
        const uint32_t fb = smem_u32(&full_bar[s]);
        mbar_expect_tx(fb, kStageBytes);
        const uint32_t sa = smem_u32(smem + (size_t)s * kStageBytes);
        const uint32_t sb = sa + kABytes;
        const int k0 = (int)(k_begin + (long long)kb * BK);
        if (!A_MN) {
          tma_load_2d(sa, &map_a, fb, k0, (int)m0);                       // box {32 k, 128 m}
        } else {
#pragma unroll
          for (int j = 0; j < BM / 32; ++j) tma_load_2d(sa + j * 4096, &map_a, fb, (int)m0 + 32 * j, k0);   // box {32 m, 32 k}
        }
        if (!B_MN) {
          tma_load_2d(sb, &map_b, fb, k0, (int)n0);                       // box {32 k, BN n}
        } else {
#pragma unroll
          for (int j = 0; j < BN / 32; ++j) tma_load_2d(sb + j * 4096, &map_b, fb, (int)n0 + 32 * j, k0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      // instruction descriptor: D=F32, A=B=TF32, majors, N, M
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(smem_u32(&full_bar[s]), ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * kStageBytes);
        const uint32_t sb = sa + kABytes;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // K-major: rows of 128 B, 8-row groups 1024 B apart, +32 B per K step inside the swizzle span.
          // MN-major: 32-wide MN blocks 4096 B apart (LBO), 8-deep K groups 1024 B apart (SBO).
          uint64_t ad = A_MN ? make_desc(sa + k * 1024, 4096, 1024) : make_desc(sa + k * 32, 16, 1024);
          uint64_t bd = B_MN ? make_desc(sb + k * 1024, 4096, 1024) : make_desc(sb + k * 32, 16, 1024);
          umma_tf32(tmem_base, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(smem_u32(&empty_bar[s]));          // frees the smem stage when these MMAs retire
      }
      umma_commit(smem_u32(&tmem_full_bar));           // accumulator complete
    }
  } else {
    // ===== epilogue: TMEM -> registers -> (warp transpose in smem) -> coalesced global =====
    // tcgen05.ld hands thread t the 32 consecutive columns of accumulator row t; writing those
    // directly makes every store instruction touch 32 different rows (32 half-empty sectors).
    // A 32x33 shared-memory transpose per warp turns each store into one 128-byte row segment.
    const int q = warp & 3;                            // TMEM lane quarter this warp may access
    float* sm = s_epi[q];
    if (num_kb > 0) {
      mbar_wait(smem_u32(&tmem_full_bar), 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      float v[32];
      if (num_kb > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) sm[lane * 33 + i] = v[i];
      __syncwarp();
      const long long n = n0 + c * 32 + lane;
      const bool nok = n < p.N;
      const float bias = (nok && p.bias_n && (!p.atomic_out || blockIdx.z == 0)) ? p.bias_n[n] : 0.f;
#pragma unroll 4
      for (int rr = 0; rr < 32; ++rr) {
        const long long row = m0 + q * 32 + rr;
        if (row < p.M && nok) {
          float x = p.alpha * sm[rr * 33 + lane] + bias;
          float* dst = p.C + row * p.ldc + n;
          if (p.atomic_out) atomicAdd(dst, x);
          else {
            if (p.beta != 0.f) x += p.beta * *dst;
            *dst = x;
          }
        }
      }
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

template <int BN, bool A_MN, bool B_MN>
int launch(const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, dim3 grid, cudaStream_t st) {
  constexpr size_t smem = (size_t)kStages * (BM * BK * 4 + BN * BK * 4) + 1024;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(gemm_tc_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
      return ARX_E_LAUNCH;
    configured = true;
  }
  gemm_tc_kernel<BN, A_MN, B_MN><<<grid, kThreads, smem, st>>>(ma, mb, p);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

}  // namespace

// C[m,n] = alpha * op(A) op(B) + bias (+ beta C) on the tensor cores.  Same argument meaning as
// arx_gemm; returns ARX_E_UNSUPPORTED when TMA cannot describe the operands (then call arx_gemm).
extern "C" int arx_gemm_tc(const float* A, const float* B, float* C, int64_t m, int64_t n, int64_t k,
                           int trans_a, int trans_b, const float* bias_n, float alpha, float beta,
                           void* stream) {
  if (!A || !B || !C || m < 0 || n < 0 || k < 0) return ARX_E_BADARG;
  if (m == 0 || n == 0) return ARX_OK;
  // TMA needs 16-byte aligned bases and row pitches
  const long long a_cols = trans_a ? m : k, b_cols = trans_b ? k : n;
  if ((a_cols % 4) || (b_cols % 4) || ((uintptr_t)A & 15) || ((uintptr_t)B & 15) || k == 0) return ARX_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const bool a_mn = trans_a != 0;        // A stored [k, m]
  const bool b_mn = trans_b == 0;        // B stored [k, n]
  // MN-major tf32 operands need the SWIZZLE_128B_BASE32B shared-memory layout, which this kernel
  // does not stage yet: the host transposes such an operand (arx_transpose) and calls K-major.
  if (a_mn || b_mn) return ARX_E_UNSUPPORTED;
  CUtensorMap ma, mb;
  const int BN = (n > 64) ? 128 : 64;
  if (!make_map(&ma, A, a_mn ? k : m, a_mn ? m : k, a_mn ? 32 : BM)) return ARX_E_UNSUPPORTED;
  if (!make_map(&mb, B, b_mn ? k : n, b_mn ? n : k, b_mn ? 32 : BN)) return ARX_E_UNSUPPORTED;
  // split K when the output grid alone cannot fill the machine
  const long long tiles = ((m + BM - 1) / BM) * ((n + BN - 1) / BN);
  long long splits = 1;
  const int sms = arx_num_sms();
  if (tiles < sms && k >= 8 * BK && beta == 0.f) {
    splits = (long long)(sms / tiles);
    if (splits > (long long)k / (4 * BK)) splits = (long long)k / (4 * BK);
    if (splits < 1) splits = 1;
  }
  long long kps = ((k + splits - 1) / splits + BK - 1) / BK * BK;
  splits = (k + kps - 1) / kps;
  GemmParams p{C, bias_n, m, n, k, n, kps, alpha, beta, splits > 1 ? 1 : 0};
  if (splits > 1 && cudaMemsetAsync(C, 0, sizeof(float) * (size_t)m * n, st) != cudaSuccess) return ARX_E_LAUNCH;
  dim3 grid((unsigned)((n + BN - 1) / BN), (unsigned)((m + BM - 1) / BM), (unsigned)splits);
  if (grid.y > 65535u) return ARX_E_UNSUPPORTED;
#define ARX_TC(BN_)                                                                       \
  (a_mn ? (b_mn ? launch<BN_, true, true>(ma, mb, p, grid, st) : launch<BN_, true, false>(ma, mb, p, grid, st)) \
        : (b_mn ? launch<BN_, false, true>(ma, mb, p, grid, st) : launch<BN_, false, false>(ma, mb, p, grid, st)))
  return BN == 128 ? ARX_TC(128) : ARX_TC(64);
#undef ARX_TC
}
