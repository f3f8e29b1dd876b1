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
#include <cuda.h>

namespace {

constexpr int BM = 128;          // UMMA M (TMEM lanes)
constexpr int BK = 32;           // tf32 elements per stage row = 128 bytes = one swizzle span
constexpr int UMMA_K = 8;        // tf32: 32 bytes per instruction
constexpr int kStages = 4;
constexpr int kThreads = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 consecutive fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), SWIZZLE_128B, version 1.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;            // version = 1 (Blackwell)
  d |= (uint64_t)2 << 61;            // layout_type = SWIZZLE_128B
  return d;
}

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

// ---- host side: tensor maps -------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D fp32 tensor [rows, cols] row-major, box {box_cols (=32 -> 128 B), box_rows}, 128-byte swizzle.
bool make_map(CUtensorMap* map, const float* base, long long rows, long long cols, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
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
