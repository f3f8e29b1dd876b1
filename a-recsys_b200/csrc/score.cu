// Catalog scoring (K3/K4), row losses (K5/K6) and top-k (K10) over materialised scores.
// This file holds the exact-fp32 SIMT path: it is the parity anchor and serves the small
// shapes (sampled pool, MovieLens-sized catalogs); the large-catalog path is the
// tcgen05 kernel in score_tc.cu.
//
// Reference semantics: attributes/embed_attribute.py:148-220 (get_prediction after the
// pool-first rewrite, get_target_score), :525-649 (compute_loss and the WMRB family),
// :651-672,:721-747 (positive mask, here a CSR per row), hmf/hmf_model.py:154 (top_k).
#include "arx_common.cuh"
#include <math_constants.h>

namespace {

// ------------------------------------------------------------------ SIMT GEMM -------
// C[m,n] = alpha * sum_k A(m,k) B(k,n) + bias[n] + beta * C ; 64x64 tile, 4x4 per thread.
constexpr int BM = 64, BN = 64, BK = 16;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                 long long M, long long N, long long K, const float* __restrict__ bias_n,
                 float alpha, float beta) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const long long m0 = (long long)blockIdx.y * BM, n0 = (long long)blockIdx.x * BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long k0 = 0; k0 < K; k0 += BK) {
    // A tile: BM x BK
    for (int i = tid; i < BM * BK; i += 256) {
      int mm, kk;
      if (TA) { mm = i % BM; kk = i / BM; } else { kk = i % BK; mm = i / BK; }
      const long long gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < M && gk < K) v = TA ? A[gk * M + gm] : A[gm * K + gk];
      As[kk][mm] = v;
    }
    for (int i = tid; i < BN * BK; i += 256) {
      int nn, kk;
      if (TB) { kk = i % BK; nn = i / BK; } else { nn = i % BN; kk = i / BN; }
      const long long gn = n0 + nn, gk = k0 + kk;
      float v = 0.f;
      if (gn < N && gk < K) v = TB ? B[gn * K + gk] : B[gk * N + gn];
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = alpha * acc[i][j];
      if (bias_n) v += bias_n[gn];
      if (beta != 0.f) v += beta * C[gm * N + gn];
      C[gm * N + gn] = v;
    }
  }
}

// ------------------------------------------------------------------ block reductions -
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) { r = warp_sum(r); if (lane == 0) red[0] = r; }
  __syncthreads();
  return red[0];
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : -CUDART_INF_F;
  if (w == 0) { r = warp_max(r); if (lane == 0) red[0] = r; }
  __syncthreads();
  return red[0];
}
__device__ __forceinline__ long long block_sum_ll(long long v, long long* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(ARX_FULL_MASK, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  long long r = (threadIdx.x < nw) ? red[threadIdx.x] : 0;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(ARX_FULL_MASK, r, o);
    if (lane == 0) red[0] = r;
  }
  __syncthreads();
  return red[0];
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// transform of the masked error sum (embed_attribute.py:580-594); returns l, writes dl/dS
__device__ __forceinline__ float rs_transform(float S, int kind, int lf, float p, float& g) {
  if (kind == ARX_LOSS_WARP || kind == ARX_LOSS_MW) lf = ARX_LF_LOG;
  if (kind == ARX_LOSS_BBPR) lf = ARX_LF_LINEAR;
  switch (lf) {
    case ARX_LF_LOG:    g = 1.0f / (1.0f + S); return logf(1.0f + S);
    case ARX_LF_EXP:  { const float q = powf(p, -S); g = logf(p) * q; return 1.0f - q; }
    case ARX_LF_POLY:   g = p * powf(S, p - 1.0f); return powf(S, p);
    case ARX_LF_POLY2:  g = p * powf(1.0f + S, p - 1.0f); return powf(1.0f + S, p);
    case ARX_LF_LINEAR: g = 1.0f; return S;
    default:            g = 2.0f * S; return S * S;
  }
}

// e(z), e'(z) with z = s_v - s_target (embed_attribute.py:565-578, :615-618, :646-649)
__device__ __forceinline__ float err_fn(float z, int kind, float& de) {
  if (kind == ARX_LOSS_RS_SIG2 || kind == ARX_LOSS_BBPR) {
    const float s = sigmoidf_(z);
    de = s * (1.0f - s);
    return s;
  }
  const float r = z + 1.0f;
  if (r <= 0.f) { de = 0.f; return 0.f; }        // ReluGrad: 0 at 0
  if (kind == ARX_LOSS_RS_SIG) {
    const float s = sigmoidf_(r);
    de = 2.0f * s * (1.0f - s);
    return 2.0f * s - 1.0f;
  }
  de = 1.0f;
  return r;
}

constexpr int kFilterBits = 1 << 15;

// One CTA per row.  Positives of the row are a sorted CSR slice; a hashed bitset in
// shared memory filters the (rare) membership probes, a binary search confirms.
__global__ void __launch_bounds__(256)
loss_rows_kernel(const float* __restrict__ scores, long long V, long long ld,
                 const int* __restrict__ target, const float* __restrict__ target_score,
                 const int* __restrict__ pos_row, const int* __restrict__ pos_ptr,
                 const int* __restrict__ pos_idx, int kind, int lf,
                 float exp_p, const float* __restrict__ row_scale, float* __restrict__ loss,
                 float* __restrict__ dscores, float* __restrict__ dtarget,
                 long long* __restrict__ rank_out) {
  __shared__ float red[32];
  __shared__ long long red_ll[32];
  __shared__ unsigned int filt[kFilterBits / 32];
  const long long b = blockIdx.x;
  const float* x = scores + b * ld;
  float* d = dscores ? dscores + b * ld : nullptr;
  const float scale = row_scale ? row_scale[b] : 1.0f;
  const int tid = threadIdx.x;

  if (kind == ARX_LOSS_CE) {
    const int t = target[b];
    const float xt = x[t];                       // read before any in-place gradient write
    float mx = -CUDART_INF_F;
    for (long long v = tid; v < V; v += blockDim.x) mx = fmaxf(mx, x[v]);
    mx = block_max(mx, red);
    float s = 0.f;
    for (long long v = tid; v < V; v += blockDim.x) s += expf(x[v] - mx);
    s = block_sum(s, red);
    if (tid == 0) loss[b] = mx + logf(s) - xt;
    if (d) {
      const float inv = 1.0f / s;
      for (long long v = tid; v < V; v += blockDim.x) {
        float p = expf(x[v] - mx) * inv;
        if (v == t) p -= 1.0f;
        d[v] = p * scale;
      }
    }
    return;
  }

  // ---- WMRB family ---------------------------------------------------------------
  const long long prow = pos_row ? (long long)pos_row[b] : b;
  const int p0 = pos_ptr ? pos_ptr[prow] : 0;
  const int p1 = pos_ptr ? pos_ptr[prow + 1] : 0;
  const bool exact = V <= kFilterBits;        // bitset is collision-free: no search needed
  for (int i = tid; i < kFilterBits / 32; i += blockDim.x) filt[i] = 0u;
  __syncthreads();
  for (int i = p0 + tid; i < p1; i += blockDim.x) {
    const int pv = pos_idx[i];
    if (pv < 0) continue;
    const unsigned int h = (unsigned int)pv & (kFilterBits - 1);
    atomicOr(&filt[h >> 5], 1u << (h & 31));
  }
  __syncthreads();
  auto masked = [&](long long v) -> bool {
    const unsigned int h = (unsigned int)v & (kFilterBits - 1);
    if (!((filt[h >> 5] >> (h & 31)) & 1u)) return false;
    if (exact) return true;
    if (kind == ARX_LOSS_MW) {                  // pool-position columns are in item order (unsorted, -1 = not in
      for (int i = p0; i < p1; ++i)             // the pool): confirm a (rare) filter hit with a linear scan
        if (pos_idx[i] == (int)v) return true;
      return false;
    }
    int lo = p0, hi = p1;                       // binary search in the sorted slice
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const int pv = pos_idx[mid];
      if (pv == (int)v) return true;
      if (pv < (int)v) lo = mid + 1; else hi = mid;
    }
    return false;
  };
  const int t = (kind == ARX_LOSS_MW) ? -1 : target[b];
  const float xt = (kind == ARX_LOSS_MW) ? target_score[b] : x[t];

  float S = 0.f;
  long long rank = 0;
  for (long long v = tid; v < V; v += blockDim.x) {
    if (masked(v)) continue;
    float de;
    const float xv = x[v];
    S += err_fn(xv - xt, kind, de);
    if (xv - xt > 0.f) rank += 1;
  }
  S = block_sum(S, red);
  if (rank_out) { rank = block_sum_ll(rank, red_ll); if (tid == 0) rank_out[b] = rank; }
  float g;
  const float l = rs_transform(S, kind, lf, exp_p, g);
  if (tid == 0) loss[b] = l;
  if (!d && !dtarget) return;
  const float gs = g * scale;
  float D = 0.f;
  for (long long v = tid; v < V; v += blockDim.x) {
    float dv = 0.f;
    if (!masked(v)) {
      float de;
      err_fn(x[v] - xt, kind, de);
      dv = de * gs;
    }
    D += dv;
    if (d) d[v] = dv;
  }
  D = block_sum(D, red);                        // also orders the d[] writes before the fix-up
  if (tid == 0) {
    if (kind == ARX_LOSS_MW) { if (dtarget) dtarget[b] = -D; }
    else if (d) d[t] -= D;
  }
}

// ------------------------------------------------------------------ row dot ---------
__global__ void rowdot_fwd_kernel(const float* __restrict__ U, const float* __restrict__ P,
                                  const float* __restrict__ beta, long long mb, int dim,
                                  float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long b = warp0; b < mb; b += nwarps) {
    float s = 0.f;
    for (int c = lane; c < dim; c += 32) s = fmaf(U[b * dim + c], P[b * dim + c], s);
    s = warp_sum(s);
    if (lane == 0) out[b] = s + (beta ? beta[b] : 0.f);
  }
}
__global__ void rowdot_bwd_kernel(const float* __restrict__ U, const float* __restrict__ P,
                                  const float* __restrict__ dts, long long mb, int dim,
                                  float* __restrict__ dU, float* __restrict__ dP) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mb * dim) return;
  const long long b = i / dim;
  const float g = dts[b];
  if (dU) dU[i] += g * P[i];
  if (dP) dP[i] = g * U[i];
}

// ------------------------------------------------------------------ top-k -----------
// Monotone key: larger float -> larger uint.  Composite 64-bit key (value, ~index)
// reproduces tf.nn.top_k's order: value descending, ties -> lower index first.
__device__ __forceinline__ unsigned int mono_key(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float mono_inv(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

constexpr int kTopkMax = 1024;

__global__ void __launch_bounds__(256)
topk_rows_kernel(const float* __restrict__ scores, long long V, long long ld, int k,
                 int* __restrict__ idx_out, float* __restrict__ val_out) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long cand[kTopkMax];
  __shared__ unsigned int s_prefix, s_need, s_ngt, s_neq, s_base;
  const long long b = blockIdx.x;
  const float* x = scores + b * ld;
  const int tid = threadIdx.x;
  // radix select (MSB first) of the k-th largest value key
  if (tid == 0) { s_prefix = 0u; s_need = (unsigned int)k; }
  __syncthreads();
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    hist[tid] = 0u;
    __syncthreads();
    const unsigned int prefix = s_prefix;
    const unsigned int himask = (pass == 0) ? 0u : (0xffffffffu << (shift + 8));
    for (long long v = tid; v < V; v += blockDim.x) {
      const unsigned int key = mono_key(x[v]);
      if ((key & himask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned int need = s_need, bin = 255;
      for (;; --bin) {
        const unsigned int c = hist[bin];
        if (c >= need) break;
        need -= c;
        if (bin == 0) break;
      }
      s_need = need;
      s_prefix = prefix | (bin << shift);
    }
    __syncthreads();
  }
  const unsigned int T = s_prefix;          // key of the k-th largest value
  const unsigned int need_eq = s_need;      // how many == T to take (lowest indices)
  if (tid == 0) { s_ngt = 0u; s_neq = 0u; s_base = 0u; }
  __syncthreads();
  // all strictly greater (unordered), then the first need_eq equal ones in index order
  for (long long v = tid; v < V; v += blockDim.x) {
    const unsigned int key = mono_key(x[v]);
    if (key > T) {
      const unsigned int p = atomicAdd(&s_ngt, 1u);
      cand[p] = ((unsigned long long)key << 32) | (unsigned long long)(0xffffffffu - (unsigned int)v);
    }
  }
  __syncthreads();
  const unsigned int ngt = s_ngt;
  for (long long v0 = 0; v0 < V && s_base < need_eq; v0 += blockDim.x) {
    const long long v = v0 + tid;
    const bool eq = (v < V) && (mono_key(x[v]) == T);
    // ordered compaction inside the chunk: warp ballots + serial warp offsets
    const unsigned int bal = __ballot_sync(ARX_FULL_MASK, eq);
    const int lane = tid & 31, w = tid >> 5;
    __syncthreads();
    if (lane == 0) hist[w] = __popc(bal);
    __syncthreads();
    unsigned int off = s_base;
    for (int i = 0; i < w; ++i) off += hist[i];
    const unsigned int pos = off + __popc(bal & ((1u << lane) - 1u));
    if (eq && pos < need_eq)
      cand[ngt + pos] = ((unsigned long long)T << 32) | (unsigned long long)(0xffffffffu - (unsigned int)v);
    __syncthreads();
    if (tid == 0) { unsigned int tot = 0; for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += hist[i]; s_base += tot; }
    __syncthreads();
  }
  __syncthreads();
  // bitonic sort (descending) of k composite keys, padded with 0
  int n2 = 1;
  while (n2 < k) n2 <<= 1;
  for (int i = k + tid; i < n2; i += blockDim.x) cand[i] = 0ull;
  __syncthreads();
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < n2 / 2; i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = cand[lo], c = cand[hi];
        if ((a < c) == desc) { cand[lo] = c; cand[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < k; i += blockDim.x) {
    const unsigned long long c = cand[i];
    idx_out[b * k + i] = (int)(0xffffffffu - (unsigned int)(c & 0xffffffffull));
    if (val_out) val_out[b * k + i] = mono_inv((unsigned int)(c >> 32));
  }
}

}  // namespace

extern "C" int arx_gemm(const float* A, const float* B, float* C, int64_t m, int64_t n, int64_t k,
                        int trans_a, int trans_b, const float* bias_n, float alpha, float beta,
                        void* stream) {
  if (!A || !B || !C || m < 0 || n < 0 || k < 0) return ARX_E_BADARG;
  if (m == 0 || n == 0) return ARX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((n + BN - 1) / BN), (unsigned)((m + BM - 1) / BM));
  if (grid.y > 65535u) return ARX_E_UNSUPPORTED;
  if (trans_a && trans_b) return ARX_E_UNSUPPORTED;
  if (trans_a)      gemm_simt_kernel<true, false><<<grid, 256, 0, st>>>(A, B, C, m, n, k, bias_n, alpha, beta);
  else if (trans_b) gemm_simt_kernel<false, true><<<grid, 256, 0, st>>>(A, B, C, m, n, k, bias_n, alpha, beta);
  else              gemm_simt_kernel<false, false><<<grid, 256, 0, st>>>(A, B, C, m, n, k, bias_n, alpha, beta);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_loss_rows(const float* scores, int64_t mb, int64_t V, int64_t ld,
                             const int32_t* target, const float* target_score,
                             const int32_t* pos_row, const int32_t* pos_ptr,
                             const int32_t* pos_idx, int loss_kind,
                             int loss_func, float exp_p, const float* row_scale, float* loss,
                             float* dscores, float* dtarget, int64_t* rank_out, void* stream) {
  if (!scores || !loss || mb < 0 || V < 1 || ld < V) return ARX_E_BADARG;
  if (loss_kind < ARX_LOSS_CE || loss_kind > ARX_LOSS_MW) return ARX_E_BADARG;
  if (loss_kind == ARX_LOSS_MW ? (target_score == nullptr) : (target == nullptr)) return ARX_E_BADARG;
  if (mb == 0) return ARX_OK;
  loss_rows_kernel<<<(unsigned)mb, 256, 0, (cudaStream_t)stream>>>(
      scores, V, ld, target, target_score, pos_row, pos_ptr, pos_idx, loss_kind, loss_func, exp_p, row_scale,
      loss, dscores, dtarget, (long long*)rank_out);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_rowdot_fwd(const float* U, const float* P, const float* beta, int64_t mb, int dim,
                              float* out, void* stream) {
  if (!U || !P || !out || mb < 0 || dim < 1) return ARX_E_BADARG;
  if (mb == 0) return ARX_OK;
  const int blocks = (int)((mb + 7) / 8);
  rowdot_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(U, P, beta, mb, dim, out);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_rowdot_bwd(const float* U, const float* P, const float* dts, int64_t mb, int dim,
                              float* dU_accum, float* dP, void* stream) {
  if (!U || !P || !dts || mb < 0 || dim < 1) return ARX_E_BADARG;
  if (mb == 0) return ARX_OK;
  const long long tot = mb * dim;
  rowdot_bwd_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(U, P, dts, mb, dim, dU_accum, dP);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_topk_rows(const float* scores, int64_t mb, int64_t V, int64_t ld, int k,
                             int32_t* idx_out, float* val_out, void* stream) {
  if (!scores || !idx_out || mb < 0 || V < 1 || k < 1 || k > kTopkMax || k > V) return ARX_E_BADARG;
  if (mb == 0) return ARX_OK;
  topk_rows_kernel<<<(unsigned)mb, 256, 0, (cudaStream_t)stream>>>(scores, V, ld, k, idx_out, val_out);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}
