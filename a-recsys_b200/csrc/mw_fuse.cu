// Small fused kernels around the sampled-WMRB tensor-core kernels (arx_mw_fwd / arx_mw_bwd) of the HMF `mw` step
// (hmf/hmf_model.py:78,112-115 + embed_attribute.py:208-220,236): the dependent chain between the lookups and the
// scatter-Adagrad was 12 launches of 3-15 us each; these two replace eight of them.
//   arx_mw_prep : u = dropout(u0) ; U_r = tf32(u) ; UT = U_r^T ; tscore = u . Pt + bt ; P_r = tf32(Ps) ; PT = P_r^T
//                 (was arx_scale_mask + 2 x arx_round_tf32 + 2 x arx_transpose + arx_rowdot_fwd)
//   arx_mw_post : du0 = (dU + dts * Pt) * mask / keep ; dPt = dts * u
//                 (was arx_rowdot_bwd + arx_scale_mask)
#include "arx_common.cuh"
#include "philox.cuh"

namespace {

// same rounding as arx_round_tf32 (optim.cu): nearest tf32, ties to even — the fused and the unfused glue agree bit for bit
__device__ __forceinline__ float tf32_round(float x) {
  unsigned int u = __float_as_uint(x);
  if ((u & 0x7f800000u) == 0x7f800000u) return x;
  u += 0x00000fffu + ((u >> 13) & 1u);
  return __uint_as_float(u & 0xffffe000u);
}

// tf.nn.dropout keeps an element when floor(keep + U[0,1)) == 1, i.e. with probability keep
__device__ __forceinline__ float dropout_keep(const unsigned long long* rng, long long elem, float keep) {
  uint32_t o[4];
  philox_words(rng, (unsigned long long)elem >> 2, o);
  return floorf(keep + u01(o[elem & 3]));
}

constexpr int kPrepRows = 32;

// grid: ceil(M / 32) blocks for the batch rows, then ceil(S / 32) blocks for the pool rows; 256 threads
__global__ void __launch_bounds__(256)
mw_prep_kernel(const float* __restrict__ u0, const float* __restrict__ mask, float inv_keep,
               const unsigned long long* __restrict__ rng, float* __restrict__ mask_out,
               const float* __restrict__ Pt, const float* __restrict__ bt, const float* __restrict__ Ps,
               long long M, long long S, int d, float* __restrict__ u, float* __restrict__ U_r,
               float* __restrict__ UT, float* __restrict__ tscore, float* __restrict__ P_r, float* __restrict__ PT) {
  extern __shared__ float tile[];                       // [32][d + 1]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long ublocks = (M + kPrepRows - 1) / kPrepRows;
  const bool is_u = (long long)blockIdx.x < ublocks;
  const long long row0 = (is_u ? (long long)blockIdx.x : (long long)blockIdx.x - ublocks) * kPrepRows;
  const long long R = is_u ? M : S;
  const float* __restrict__ src = is_u ? u0 : Ps;
  float* __restrict__ dst_r = is_u ? U_r : P_r;
  float* __restrict__ dst_t = is_u ? UT : PT;
  const int ld = d + 1;
  if ((d & 3) == 0) {
    // 128-bit path: a lane owns 4 consecutive columns (one Philox block = the same 4 mask values the scalar path
    // draws), the warp's 4 rows are loaded together (all loads of the tile in flight before the first use)
    const int d4 = d >> 2;
    constexpr int RPW = kPrepRows / 8;
    float dot[RPW];
#pragma unroll
    for (int q = 0; q < RPW; ++q) dot[q] = 0.f;
    const int n_it = (d4 + 31) >> 5;                      // uniform trip count: the warp reductions below need every lane
    for (int itc = 0; itc < n_it; ++itc) {
      const int c4 = itc * 32 + lane;
      const bool cok = c4 < d4;
      float4 x[RPW], pt[RPW], mk[RPW];
#pragma unroll
      for (int q = 0; q < RPW; ++q) {
        const long long row = row0 + warp + 8 * q;
        x[q] = f4_zero(); pt[q] = f4_zero(); mk[q] = make_float4(1.f, 1.f, 1.f, 1.f);
        if (cok && row < R) {
          x[q] = ldg_f4(src + row * d + c4 * 4);
          if (is_u && mask != nullptr) mk[q] = ldg_f4(mask + row * d + c4 * 4);
          if (is_u && Pt != nullptr) pt[q] = ldg_f4(Pt + row * d + c4 * 4);
        }
      }
#pragma unroll
      for (int q = 0; q < RPW; ++q) {
        const int r = warp + 8 * q;
        const long long row = row0 + r;
        if (!cok) continue;
        float4 v = x[q];
        if (row < R) {
          if (is_u) {
            if (mask != nullptr) {
              v = make_float4(v.x * inv_keep * mk[q].x, v.y * inv_keep * mk[q].y, v.z * inv_keep * mk[q].z, v.w * inv_keep * mk[q].w);
            } else if (rng != nullptr) {
              uint32_t o[4];
              philox_words(rng, (unsigned long long)(row * d + c4 * 4) >> 2, o);
              const float keep = 1.0f / inv_keep;
              const float4 m = make_float4(floorf(keep + u01(o[0])), floorf(keep + u01(o[1])), floorf(keep + u01(o[2])),
                                           floorf(keep + u01(o[3])));
              st_f4(mask_out + row * d + c4 * 4, m);
              v = make_float4(v.x * inv_keep * m.x, v.y * inv_keep * m.y, v.z * inv_keep * m.z, v.w * inv_keep * m.w);
            }
            st_f4(u + row * d + c4 * 4, v);
            dot[q] = fmaf(v.x, pt[q].x, fmaf(v.y, pt[q].y, fmaf(v.z, pt[q].z, fmaf(v.w, pt[q].w, dot[q]))));
          }
          v = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
          st_f4(dst_r + row * d + c4 * 4, v);
        }
        float* t = tile + r * ld + c4 * 4;
        t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
      }
    }
    if (is_u && tscore != nullptr) {
#pragma unroll
      for (int q = 0; q < RPW; ++q) {
        const long long row = row0 + warp + 8 * q;
        const float dsum = warp_sum(dot[q]);
        if (lane == 0 && row < R) tscore[row] = dsum + (bt != nullptr ? __ldg(bt + row) : 0.f);   // :220
      }
    }
  } else
  for (int r = warp; r < kPrepRows; r += 8) {
    const long long row = row0 + r;
    float dot = 0.f;
    for (int c = lane; c < d; c += 32) {
      float x = 0.f;
      if (row < R) {
        x = __ldg(src + row * d + c);
        if (is_u) {
          if (mask != nullptr) x = x * inv_keep * __ldg(mask + row * d + c);      // tf.nn.dropout: x / keep * mask
          else if (rng != nullptr) {                                              // mask drawn here, kept for the adjoint
            const float mk = dropout_keep(rng, row * d + c, 1.0f / inv_keep);
            mask_out[row * d + c] = mk;
            x = x * inv_keep * mk;
          }
          u[row * d + c] = x;
          if (Pt != nullptr) dot = fmaf(x, __ldg(Pt + row * d + c), dot);
        }
        const float xr = tf32_round(x);
        dst_r[row * d + c] = xr;
        x = xr;
      }
      tile[r * ld + c] = x;
    }
    if (is_u && tscore != nullptr) {
      dot = warp_sum(dot);
      if (lane == 0 && row < R) tscore[row] = dot + (bt != nullptr ? __ldg(bt + row) : 0.f);   // :220
    }
  }
  __syncthreads();
  const long long row = row0 + lane;
  if (dst_t != nullptr && row < R)
    for (int c = warp; c < d; c += 8) dst_t[(long long)c * R + row] = tile[lane * ld + c];
}

__global__ void __launch_bounds__(256)
mw_post_kernel(const float* __restrict__ dU, const float* __restrict__ dts, const float* __restrict__ Pt,
               const float* __restrict__ u, const float* __restrict__ mask, float inv_keep, long long n, int d,
               float* __restrict__ du0, float* __restrict__ dPt, unsigned long long* __restrict__ rng) {
  // the step's dropout draw is over (arx_mw_prep ran before this kernel): advance the Philox step counter
  if (rng != nullptr && blockIdx.x == 0 && threadIdx.x == 0) rng[1] += 1ull;
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    const float g = __ldg(dts + i / d);
    const float4 a = ld_f4(dU + i), p = ldg_f4(Pt + i), uu = ldg_f4(u + i);
    float4 o = make_float4(fmaf(g, p.x, a.x), fmaf(g, p.y, a.y), fmaf(g, p.z, a.z), fmaf(g, p.w, a.w));
    if (mask != nullptr) {
      const float4 m = ldg_f4(mask + i);
      o = make_float4(o.x * inv_keep * m.x, o.y * inv_keep * m.y, o.z * inv_keep * m.z, o.w * inv_keep * m.w);
    }
    st_f4(du0 + i, o);
    st_f4(dPt + i, make_float4(g * uu.x, g * uu.y, g * uu.z, g * uu.w));
  }
}

}  // namespace

extern "C" int arx_mw_prep(const float* u0, const float* mask, float inv_keep, const uint64_t* rng_state,
                           float* mask_out, const float* Pt, const float* bt, const float* Ps, int64_t M, int64_t S,
                           int d, float* u, float* U_r, float* UT, float* tscore, float* P_r, float* PT, void* stream) {
  if (!u0 || !u || !U_r || M < 0 || S < 0 || d < 1 || (S > 0 && (!Ps || !P_r))) return ARX_E_BADARG;
  if (!mask && rng_state && !mask_out) return ARX_E_BADARG;
  if (M == 0 && S == 0) return ARX_OK;
  const size_t smem = (size_t)kPrepRows * (d + 1) * sizeof(float);
  if (smem > 48 * 1024) return ARX_E_UNSUPPORTED;
  const long long blocks = (M + kPrepRows - 1) / kPrepRows + (S + kPrepRows - 1) / kPrepRows;
  mw_prep_kernel<<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(
      u0, mask, inv_keep, reinterpret_cast<const unsigned long long*>(rng_state), mask_out, Pt, bt, Ps, (long long)M,
      (long long)S, d, u, U_r, UT, tscore, P_r, PT);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_mw_post(const float* dU, const float* dts, const float* Pt, const float* u, const float* mask,
                           float inv_keep, int64_t M, int d, float* du0, float* dPt, uint64_t* rng_state, void* stream) {
  if (!dU || !dts || !Pt || !u || !du0 || !dPt || M < 0 || d < 1) return ARX_E_BADARG;
  if (M == 0) return ARX_OK;
  if ((d % 4) || ((uintptr_t)dU & 15) || ((uintptr_t)Pt & 15) || ((uintptr_t)u & 15) || ((uintptr_t)du0 & 15) ||
      ((uintptr_t)dPt & 15) || (mask && ((uintptr_t)mask & 15)))
    return ARX_E_UNSUPPORTED;
  const long long n = (long long)M * d;
  const long long b = (n / 4 + 255) / 256;
  const int grid = (int)std::min<long long>(std::max<long long>(b, 1), (long long)arx_num_sms() * 8);
  mw_post_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dU, dts, Pt, u, mask, inv_keep, n, d, du0, dPt,
                                                         reinterpret_cast<unsigned long long*>(rng_state));
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}
