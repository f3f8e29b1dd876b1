// mulhot_pool: heterogeneous attribute embedding gather + segment-mean (K1+K2) and its
// adjoint with de-duplicated sparse Adagrad (K2b).  HBM-bound integer/row work: no
// tensor cores here — coalesced 128-bit row loads, >= 8 independent rows in flight per
// warp, descriptors staged in shared memory, grids sized in multiples of the SM count.
//
// Reference semantics: attributes/embed_attribute.py:350-417 (_get_embedded),
// attributes/mulhot_index.py:48-67 (batch_slice2 / batch_segids2), autodiff +
// tf.train.AdagradOptimizer at hmf/hmf_model.py:146-151.
#include "arx_common.cuh"

namespace {

constexpr int kMaxAttr = 32;          // one lane per attribute for the (start,len) prefetch
constexpr int kRowsInFlight = 8;      // independent row loads per lane group

// ---- tiny vector abstraction: VEC = 4 (float4, dim % 4 == 0) or 1 (any dim) --------
template <int VEC> struct V;
template <> struct V<4> {
  using T = float4;
  static __device__ __forceinline__ T zero() { return f4_zero(); }
  static __device__ __forceinline__ T ldg(const float* p) { return ldg_f4(p); }
  static __device__ __forceinline__ T ld(const float* p) { return ld_f4(p); }
  static __device__ __forceinline__ void st(float* p, T v) { st_f4(p, v); }
  static __device__ __forceinline__ void add(T& a, T b) { f4_add(a, b); }
  static __device__ __forceinline__ void fma(T& a, float w, T b) { f4_fma(a, w, b); }
  static __device__ __forceinline__ T div(T a, float s) {
    return make_float4(a.x / s, a.y / s, a.z / s, a.w / s);
  }
  static __device__ __forceinline__ T mul(T a, float s) { return f4_scale(a, s); }
  static __device__ __forceinline__ T shfl_xor(T a, int o) {
    return make_float4(__shfl_xor_sync(ARX_FULL_MASK, a.x, o), __shfl_xor_sync(ARX_FULL_MASK, a.y, o),
                       __shfl_xor_sync(ARX_FULL_MASK, a.z, o), __shfl_xor_sync(ARX_FULL_MASK, a.w, o));
  }
  static __device__ __forceinline__ float sumsq(T a) { return a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w; }
  // Adagrad on one vector: acc += g^2; w -= lr * g / sqrt(acc)
  static __device__ __forceinline__ void adagrad(T& w, T& a, T g, float lr) {
    a.x = fmaf(g.x, g.x, a.x); a.y = fmaf(g.y, g.y, a.y); a.z = fmaf(g.z, g.z, a.z); a.w = fmaf(g.w, g.w, a.w);
    w.x -= lr * g.x / sqrtf(a.x); w.y -= lr * g.y / sqrtf(a.y);
    w.z -= lr * g.z / sqrtf(a.z); w.w -= lr * g.w / sqrtf(a.w);
  }
  static __device__ __forceinline__ void sgd(T& w, T g, float lr) {
    w.x -= lr * g.x; w.y -= lr * g.y; w.z -= lr * g.z; w.w -= lr * g.w;
  }
};
template <> struct V<1> {
  using T = float;
  static __device__ __forceinline__ T zero() { return 0.f; }
  static __device__ __forceinline__ T ldg(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ T ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, T v) { *p = v; }
  static __device__ __forceinline__ void add(T& a, T b) { a += b; }
  static __device__ __forceinline__ void fma(T& a, float w, T b) { a = fmaf(w, b, a); }
  static __device__ __forceinline__ T div(T a, float s) { return a / s; }
  static __device__ __forceinline__ T mul(T a, float s) { return a * s; }
  static __device__ __forceinline__ T shfl_xor(T a, int o) { return __shfl_xor_sync(ARX_FULL_MASK, a, o); }
  static __device__ __forceinline__ float sumsq(T a) { return a * a; }
  static __device__ __forceinline__ void adagrad(T& w, T& a, T g, float lr) {
    a = fmaf(g, g, a);
    w -= lr * g / sqrtf(a);
  }
  static __device__ __forceinline__ void sgd(T& w, T g, float lr) { w -= lr * g; }
};

__device__ __forceinline__ void stage_descs(arx_attr_desc* s_attrs, const arx_attr_desc* g_attrs, int n_attr) {
  const int words = n_attr * (int)(sizeof(arx_attr_desc) / 8);
  const unsigned long long* src = reinterpret_cast<const unsigned long long*>(g_attrs);
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(s_attrs);
  for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
}

// Lane f < n_attr fetches the bag (start, length) of attribute f for entity e.
// A categorical attribute is a bag of length 1 starting at e inside features_cat.
__device__ __forceinline__ void fetch_bags(const arx_attr_desc* s_attrs, int n_attr, int lane, int e,
                                           int& my_s, int& my_L) {
  my_s = 0; my_L = 0;
  if (lane < n_attr) {
    if (s_attrs[lane].kind == 1) {
      my_s = __ldg(s_attrs[lane].starts + e);
      my_L = __ldg(s_attrs[lane].lengths + e);
    } else {
      my_s = e; my_L = 1;
    }
  }
}

// ------------------------------------------------------------------ forward ---------
// One warp per entity.  GW lanes cover one row chunk (GW*VEC floats); the 32/GW lane
// groups take different tokens of the bag and are reduced with shuffles.
template <int GW, int VEC>
__global__ void __launch_bounds__(256)
pool_fwd_kernel(const arx_attr_desc* __restrict__ g_attrs, int n_attr, int dim,
                const int* __restrict__ ids, long long n, float* __restrict__ out,
                long long out_stride, int mode, float* __restrict__ bias_out) {
  using VT = typename V<VEC>::T;
  __shared__ arx_attr_desc s_attrs[kMaxAttr];
  stage_descs(s_attrs, g_attrs, n_attr);
  constexpr int NG = 32 / GW;
  const int lane = threadIdx.x & 31;
  const int g = lane / GW, l = lane % GW;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nvec = dim / VEC;
  const float Ff = (float)n_attr;

  for (long long ei = warp0; ei < n; ei += nwarps) {
    const int e = __ldg(ids + ei);
    int my_s, my_L;
    fetch_bags(s_attrs, n_attr, lane, e, my_s, my_L);
    float bias_tot = 0.f;
    for (int c0 = 0; c0 < nvec; c0 += GW) {
      const int col = c0 + l;
      const bool colok = col < nvec;
      VT tot = V<VEC>::zero();
      for (int f = 0; f < n_attr; ++f) {
        const int s = __shfl_sync(ARX_FULL_MASK, my_s, f);
        const int L = __shfl_sync(ARX_FULL_MASK, my_L, f);
        const float* __restrict__ table = s_attrs[f].table;
        const int* __restrict__ values = s_attrs[f].values;
        const float* __restrict__ bias = s_attrs[f].bias;
        const bool want_bias = (c0 == 0) && (bias_out != nullptr) && (bias != nullptr);
        VT acc = V<VEC>::zero();
        float bsum = 0.f;
        for (int j0 = 0; j0 < L; j0 += 32) {
          const int cnt = min(32, L - j0);
          const int tok = (lane < cnt) ? __ldg(values + s + j0 + lane) : 0;
          if (want_bias && lane < cnt) bsum += __ldg(bias + tok);
          for (int jj0 = 0; jj0 < cnt; jj0 += NG * kRowsInFlight) {
            VT v[kRowsInFlight];
#pragma unroll
            for (int u = 0; u < kRowsInFlight; ++u) {
              const int jj = jj0 + u * NG + g;
              const int t = __shfl_sync(ARX_FULL_MASK, tok, jj & 31);
              v[u] = (jj < cnt && colok) ? V<VEC>::ldg(table + (size_t)t * dim + (size_t)col * VEC)
                                         : V<VEC>::zero();
            }
#pragma unroll
            for (int u = 0; u < kRowsInFlight; ++u) V<VEC>::add(acc, v[u]);
          }
        }
#pragma unroll
        for (int o = GW; o < 32; o <<= 1) V<VEC>::add(acc, V<VEC>::shfl_xor(acc, o));
        const float Lf = (float)L;
        acc = V<VEC>::div(acc, Lf);                      // tf.div(embedded_sum, lengs)  :400
        if (want_bias) bias_tot += warp_sum(bsum) / Lf;  // :404-406
        if (mode == ARX_POOL_MEAN) {
          V<VEC>::add(tot, acc);
        } else if (g == 0 && colok) {
          V<VEC>::st(out + ei * out_stride + (size_t)f * dim + (size_t)col * VEC, acc);
        }
      }
      if (mode == ARX_POOL_MEAN && g == 0 && colok)
        V<VEC>::st(out + ei * out_stride + (size_t)col * VEC, V<VEC>::div(tot, Ff));  // reduce_mean :219,:235
    }
    if (bias_out != nullptr && lane == 0) bias_out[ei] = bias_tot / Ff;               // :412
  }
}

// integer part of K2 (mulhot_index.py:48-67)
__global__ void flat_index_kernel(const arx_attr_desc* __restrict__ g_attrs, int attr,
                                  const int* __restrict__ ids, long long n,
                                  const long long* __restrict__ offsets, int* __restrict__ flat_idx,
                                  int* __restrict__ seg_ids) {
  const arx_attr_desc a = g_attrs[attr];
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long ei = warp0; ei < n; ei += nwarps) {
    const int e = ids[ei];
    int s = e, L = 1;
    if (a.kind == 1) { s = a.starts[e]; L = a.lengths[e]; }
    const long long o = offsets[ei];
    for (int j = lane; j < L; j += 32) {
      flat_idx[o + j] = a.values[s + j];
      seg_ids[o + j] = (int)ei;
    }
  }
}

// ------------------------------------------------------------------ backward plan ---
// counters: [0] n_unique  [1] bucket cursor  [2] overflow flag  [3] n_occurrences
__global__ void __launch_bounds__(256)
plan_count_kernel(const arx_attr_desc* __restrict__ g_attrs, int attr_begin, int n_attr,
                  const int* __restrict__ ids, long long n, arx_bwd_plan plan) {
  __shared__ arx_attr_desc s_attrs[kMaxAttr];
  stage_descs(s_attrs, g_attrs + attr_begin, n_attr);
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long ei = warp0; ei < n; ei += nwarps) {
    const int e = __ldg(ids + ei);
    int my_s, my_L;
    fetch_bags(s_attrs, n_attr, lane, e, my_s, my_L);
    int occ = my_L;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) occ += __shfl_xor_sync(ARX_FULL_MASK, occ, o);
    if (lane == 0) atomicAdd(&plan.counters[3], occ);
    for (int f = 0; f < n_attr; ++f) {
      const int s = __shfl_sync(ARX_FULL_MASK, my_s, f);
      const int L = __shfl_sync(ARX_FULL_MASK, my_L, f);
      const int* __restrict__ values = s_attrs[f].values;
      int* touch = s_attrs[f].touch;
      for (int j = lane; j < L; j += 32) {
        const int tok = __ldg(values + s + j);
        const int c = atomicAdd(&touch[tok], 1);
        if (c == 0) {
          const int u = atomicAdd(&plan.counters[0], 1);
          if (u < plan.cap_rows) { plan.uniq_tok[u] = tok; plan.uniq_attr[u] = attr_begin + f; }
          else plan.counters[2] = 1;
        }
      }
    }
  }
}

__global__ void plan_alloc_kernel(const arx_attr_desc* __restrict__ g_attrs, arx_bwd_plan plan) {
  const int nu = (int)min((long long)plan.counters[0], (long long)plan.cap_rows);
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < nu; u += gridDim.x * blockDim.x) {
    const int tok = plan.uniq_tok[u];
    int* touch = g_attrs[plan.uniq_attr[u]].touch;
    const int c = touch[tok];
    const int base = atomicAdd(&plan.counters[1], c);
    plan.row_base[u] = base;
    plan.row_cnt[u] = c;
    touch[tok] = base;                       // becomes the fill cursor of this row
    if ((long long)base + c > plan.cap_occ) plan.counters[2] = 1;
  }
}

__global__ void __launch_bounds__(256)
plan_fill_kernel(const arx_attr_desc* __restrict__ g_attrs, int attr_begin, int n_attr,
                 const int* __restrict__ ids, long long n, int mode, long long row_base,
                 arx_bwd_plan plan) {
  __shared__ arx_attr_desc s_attrs[kMaxAttr];
  stage_descs(s_attrs, g_attrs + attr_begin, n_attr);
  if (plan.counters[2] != 0) return;        // capacity exceeded: leave buckets untouched
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float invF = (mode == ARX_POOL_MEAN) ? 1.0f / (float)n_attr : 1.0f;
  for (long long ei = warp0; ei < n; ei += nwarps) {
    const int e = __ldg(ids + ei);
    int my_s, my_L;
    fetch_bags(s_attrs, n_attr, lane, e, my_s, my_L);
    for (int f = 0; f < n_attr; ++f) {
      const int s = __shfl_sync(ARX_FULL_MASK, my_s, f);
      const int L = __shfl_sync(ARX_FULL_MASK, my_L, f);
      const int* __restrict__ values = s_attrs[f].values;
      int* touch = s_attrs[f].touch;
      const float w = invF / (float)L;
      const int row = (mode == ARX_POOL_MEAN) ? (int)(row_base + ei) : (int)(row_base + ei * n_attr + f);
      for (int j = lane; j < L; j += 32) {
        const int tok = __ldg(values + s + j);
        const int pos = atomicAdd(&touch[tok], 1);
        plan.bucket_src[pos] = row;
        plan.bucket_w[pos] = w;
      }
    }
  }
}

__global__ void plan_reset_kernel(const arx_attr_desc* __restrict__ g_attrs, arx_bwd_plan plan) {
  const int nu = (int)min((long long)plan.counters[0], (long long)plan.cap_rows);
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < nu; u += gridDim.x * blockDim.x)
    g_attrs[plan.uniq_attr[u]].touch[plan.uniq_tok[u]] = 0;
}

// ------------------------------------------------------------------ backward apply --
// One warp per unique (table,row): segment-sum its bucket out of dOut (L2-resident),
// then one read-modify-write of the table row and its accumulator.
template <int VEC>
__global__ void __launch_bounds__(256)
pool_bwd_apply_kernel(const arx_attr_desc* __restrict__ g_attrs, int n_attr, int dim,
                      arx_bwd_plan plan, const float* __restrict__ dout, long long dout_stride,
                      const float* __restrict__ dbias, float lr,
                      const float* __restrict__ grad_scale_dev, int opt, float* __restrict__ rows_out,
                      float* __restrict__ bias_rows_out) {
  using VT = typename V<VEC>::T;
  __shared__ arx_attr_desc s_attrs[kMaxAttr];
  stage_descs(s_attrs, g_attrs, n_attr);
  if (plan.counters[2] != 0) return;
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nvec = dim / VEC;
  const int nu = (int)min((long long)plan.counters[0], (long long)plan.cap_rows);
  const float gs = grad_scale_dev ? __ldg(grad_scale_dev) : 1.0f;

  for (long long u = warp0; u < nu; u += nwarps) {
    const int tok = plan.uniq_tok[u];
    const int f = plan.uniq_attr[u];
    const int base = plan.row_base[u];
    const int cnt = plan.row_cnt[u];
    float gb = 0.f;
    for (int c0 = 0; c0 < nvec; c0 += 32) {
      const int col = c0 + lane;
      const bool colok = col < nvec;
      VT g = V<VEC>::zero();
      for (int k0 = 0; k0 < cnt; k0 += 32) {
        const int kc = min(32, cnt - k0);
        int src = 0; float w = 0.f;
        if (lane < kc) { src = __ldg(plan.bucket_src + base + k0 + lane); w = __ldg(plan.bucket_w + base + k0 + lane); }
        if (c0 == 0 && dbias != nullptr && lane < kc) gb = fmaf(w, __ldg(dbias + src), gb);
        for (int kk0 = 0; kk0 < kc; kk0 += kRowsInFlight) {
          VT v[kRowsInFlight]; float wk[kRowsInFlight];
#pragma unroll
          for (int q = 0; q < kRowsInFlight; ++q) {
            const int kk = kk0 + q;
            const int sk = __shfl_sync(ARX_FULL_MASK, src, kk & 31);
            wk[q] = __shfl_sync(ARX_FULL_MASK, w, kk & 31);
            v[q] = (kk < kc && colok) ? V<VEC>::ldg(dout + (size_t)sk * dout_stride + (size_t)col * VEC)
                                      : V<VEC>::zero();
          }
#pragma unroll
          for (int q = 0; q < kRowsInFlight; ++q) V<VEC>::fma(g, wk[q], v[q]);
        }
      }
      g = V<VEC>::mul(g, gs);
      if (colok) {
        const size_t off = (size_t)tok * dim + (size_t)col * VEC;
        if (opt == ARX_OPT_ADAGRAD) {
          float* wp = s_attrs[f].table + off;
          float* ap = s_attrs[f].table_acc + off;
          VT wv = V<VEC>::ld(wp), av = V<VEC>::ld(ap);
          V<VEC>::adagrad(wv, av, g, lr);
          V<VEC>::st(ap, av);
          V<VEC>::st(wp, wv);
        } else if (opt == ARX_OPT_SGD) {
          float* wp = s_attrs[f].table + off;
          VT wv = V<VEC>::ld(wp);
          V<VEC>::sgd(wv, g, lr);
          V<VEC>::st(wp, wv);
        } else {
          V<VEC>::st(rows_out + (size_t)u * dim + (size_t)col * VEC, g);
        }
      }
    }
    if (dbias != nullptr && s_attrs[f].bias != nullptr) {
      gb = warp_sum(gb) * gs;
      if (lane == 0) {
        if (opt == ARX_OPT_ADAGRAD) {
          float a = s_attrs[f].bias_acc[tok];
          a = fmaf(gb, gb, a);
          s_attrs[f].bias_acc[tok] = a;
          s_attrs[f].bias[tok] -= lr * gb / sqrtf(a);
        } else if (opt == ARX_OPT_SGD) {
          s_attrs[f].bias[tok] -= lr * gb;
        } else if (bias_rows_out != nullptr) {
          bias_rows_out[u] = gb;
        }
      }
    } else if (opt == ARX_OPT_NONE && bias_rows_out != nullptr && lane == 0) {
      bias_rows_out[u] = 0.f;
    }
  }
}

// IndexedSlices part of the global norm: sum over occurrences of w^2 * ||dOut[src]||^2.
__global__ void __launch_bounds__(256)
pool_bwd_sumsq_kernel(int dim, arx_bwd_plan plan, const float* __restrict__ dout,
                      long long dout_stride, const float* __restrict__ dbias,
                      const arx_attr_desc* __restrict__ g_attrs, float* __restrict__ sumsq) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nu = (int)min((long long)plan.counters[0], (long long)plan.cap_rows);
  float part = 0.f;
  for (long long u = warp0; u < nu; u += nwarps) {
    const int f = plan.uniq_attr[u];
    const int base = plan.row_base[u];
    const int cnt = plan.row_cnt[u];
    const bool has_bias = dbias != nullptr && (g_attrs == nullptr || g_attrs[f].bias != nullptr);
    for (int k = 0; k < cnt; ++k) {
      const int src = plan.bucket_src[base + k];
      const float w = plan.bucket_w[base + k];
      float s = 0.f;
      for (int c = lane; c < dim; c += 32) {
        const float v = w * __ldg(dout + (size_t)src * dout_stride + c);
        s = fmaf(v, v, s);
      }
      if (has_bias && lane == 0) { const float b = w * dbias[src]; s = fmaf(b, b, s); }
      part += s;
    }
  }
  part = warp_sum(part);
  if (lane == 0 && part != 0.f) atomicAdd(sumsq, part);
}

inline int pick_grid(long long warps_needed, int threads) {
  const int sms = arx_num_sms();
  const int wpb = threads / 32;
  long long blocks = (warps_needed + wpb - 1) / wpb;
  const long long cap = (long long)sms * (2048 / threads) * 4;   // <= 4 resident waves, grid-stride beyond
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

template <int VEC>
int launch_fwd(const arx_attr_desc* attrs, int n_attr, int dim, const int32_t* ids, int64_t n,
               float* out, int64_t out_stride, int mode, float* bias_out, cudaStream_t st) {
  const int nvec = dim / VEC;
  const int threads = 256;
  const int grid = pick_grid(n, threads);
#define ARX_FWD(GW)                                                                             \
  pool_fwd_kernel<GW, VEC><<<grid, threads, 0, st>>>(attrs, n_attr, dim, ids, (long long)n, out, \
                                                     (long long)out_stride, mode, bias_out)
  if (nvec >= 32) ARX_FWD(32);
  else if (nvec >= 16) ARX_FWD(16);
  else if (nvec >= 8) ARX_FWD(8);
  else if (nvec >= 4) ARX_FWD(4);
  else if (nvec >= 2) ARX_FWD(2);
  else ARX_FWD(1);
#undef ARX_FWD
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

}  // namespace

extern "C" int arx_pool_fwd(const arx_attr_desc* attrs, int n_attr, int dim, const int32_t* ent_ids,
                            int64_t n, float* out, int64_t out_stride, int mode, float* bias_out,
                            void* stream) {
  if (!attrs || !ent_ids || !out || n_attr < 1 || n_attr > kMaxAttr || dim < 1 || n < 0) return ARX_E_BADARG;
  if (mode != ARX_POOL_MEAN && mode != ARX_POOL_CONCAT) return ARX_E_BADARG;
  if (n == 0) return ARX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool v4 = (dim % 4 == 0) && (out_stride % 4 == 0) && (((uintptr_t)out & 15) == 0);
  return v4 ? launch_fwd<4>(attrs, n_attr, dim, ent_ids, n, out, out_stride, mode, bias_out, st)
            : launch_fwd<1>(attrs, n_attr, dim, ent_ids, n, out, out_stride, mode, bias_out, st);
}

extern "C" int arx_mulhot_flat_index(const arx_attr_desc* attrs, int attr, const int32_t* ent_ids,
                                     int64_t n, const int64_t* offsets, int32_t* flat_idx,
                                     int32_t* seg_ids, void* stream) {
  if (!attrs || !ent_ids || !offsets || !flat_idx || !seg_ids || attr < 0 || n < 0) return ARX_E_BADARG;
  if (n == 0) return ARX_OK;
  flat_index_kernel<<<pick_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(
      attrs, attr, ent_ids, (long long)n, (const long long*)offsets, flat_idx, seg_ids);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

static int plan_args_ok(const arx_bwd_plan& plan) {
  return plan.counters && plan.uniq_tok && plan.uniq_attr && plan.row_base && plan.row_cnt &&
         plan.bucket_src && plan.bucket_w && plan.cap_rows >= 1 && plan.cap_occ >= 1;
}

extern "C" int arx_bwd_plan_begin(arx_bwd_plan plan, void* stream) {
  if (!plan_args_ok(plan)) return ARX_E_BADARG;
  if (cudaMemsetAsync(plan.counters, 0, 8 * sizeof(int32_t), (cudaStream_t)stream) != cudaSuccess)
    return ARX_E_LAUNCH;
  return ARX_OK;
}

extern "C" int arx_bwd_plan_count(const arx_attr_desc* attrs, int attr_begin, int n_attr,
                                  const int32_t* ent_ids, int64_t n, arx_bwd_plan plan, void* stream) {
  if (!attrs || !ent_ids || attr_begin < 0 || n_attr < 1 || n_attr > kMaxAttr || n < 0 || !plan_args_ok(plan))
    return ARX_E_BADARG;
  if (n == 0) return ARX_OK;
  plan_count_kernel<<<pick_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(attrs, attr_begin, n_attr, ent_ids,
                                                                        (long long)n, plan);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_bwd_plan_alloc(const arx_attr_desc* attrs, arx_bwd_plan plan, void* stream) {
  if (!attrs || !plan_args_ok(plan)) return ARX_E_BADARG;
  plan_alloc_kernel<<<arx_num_sms() * 4, 256, 0, (cudaStream_t)stream>>>(attrs, plan);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_bwd_plan_fill(const arx_attr_desc* attrs, int attr_begin, int n_attr,
                                 const int32_t* ent_ids, int64_t n, int mode, int64_t row_base,
                                 arx_bwd_plan plan, void* stream) {
  if (!attrs || !ent_ids || attr_begin < 0 || n_attr < 1 || n_attr > kMaxAttr || n < 0 || !plan_args_ok(plan))
    return ARX_E_BADARG;
  if (mode != ARX_POOL_MEAN && mode != ARX_POOL_CONCAT) return ARX_E_BADARG;
  if (n == 0) return ARX_OK;
  plan_fill_kernel<<<pick_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(attrs, attr_begin, n_attr, ent_ids,
                                                                       (long long)n, mode, (long long)row_base, plan);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_bwd_plan_end(const arx_attr_desc* attrs, arx_bwd_plan plan, void* stream) {
  if (!attrs || !plan_args_ok(plan)) return ARX_E_BADARG;
  plan_reset_kernel<<<arx_num_sms() * 4, 256, 0, (cudaStream_t)stream>>>(attrs, plan);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_pool_bwd_plan(const arx_attr_desc* attrs, int n_attr, const int32_t* ent_ids,
                                 int64_t n, int mode, arx_bwd_plan plan, void* stream) {
  int rc;
  if ((rc = arx_bwd_plan_begin(plan, stream)) != ARX_OK) return rc;
  if ((rc = arx_bwd_plan_count(attrs, 0, n_attr, ent_ids, n, plan, stream)) != ARX_OK) return rc;
  if ((rc = arx_bwd_plan_alloc(attrs, plan, stream)) != ARX_OK) return rc;
  if ((rc = arx_bwd_plan_fill(attrs, 0, n_attr, ent_ids, n, mode, 0, plan, stream)) != ARX_OK) return rc;
  return arx_bwd_plan_end(attrs, plan, stream);
}

extern "C" int arx_pool_bwd_apply(const arx_attr_desc* attrs, int n_attr, int dim, arx_bwd_plan plan,
                                  const float* dout, int64_t dout_stride, const float* dbias, float lr,
                                  const float* grad_scale_dev, int opt, float* rows_out,
                                  float* bias_rows_out, void* stream) {
  if (!attrs || !dout || n_attr < 1 || n_attr > kMaxAttr || dim < 1 || !plan_args_ok(plan)) return ARX_E_BADARG;
  if (opt != ARX_OPT_ADAGRAD && opt != ARX_OPT_SGD && opt != ARX_OPT_NONE) return ARX_E_BADARG;
  if (opt == ARX_OPT_NONE && !rows_out) return ARX_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = arx_num_sms() * 8;   // persistent: warps stride over the device-side n_unique
  const bool v4 = (dim % 4 == 0) && (dout_stride % 4 == 0) && (((uintptr_t)dout & 15) == 0);
  if (v4)
    pool_bwd_apply_kernel<4><<<grid, 256, 0, st>>>(attrs, n_attr, dim, plan, dout, (long long)dout_stride,
                                                   dbias, lr, grad_scale_dev, opt, rows_out, bias_rows_out);
  else
    pool_bwd_apply_kernel<1><<<grid, 256, 0, st>>>(attrs, n_attr, dim, plan, dout, (long long)dout_stride,
                                                   dbias, lr, grad_scale_dev, opt, rows_out, bias_rows_out);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_pool_bwd_sumsq(const arx_attr_desc* attrs, int dim, arx_bwd_plan plan, const float* dout,
                                  int64_t dout_stride, const float* dbias, float* sumsq, void* stream) {
  if (!dout || !sumsq || dim < 1 || !plan_args_ok(plan)) return ARX_E_BADARG;
  pool_bwd_sumsq_kernel<<<arx_num_sms() * 4, 256, 0, (cudaStream_t)stream>>>(
      dim, plan, dout, (long long)dout_stride, dbias, attrs, sumsq);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}
