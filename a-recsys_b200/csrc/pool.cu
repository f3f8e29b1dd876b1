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
constexpr int kHeavyMax = 64;         // largest bucket a single warp reduces; larger ones are split into chunks of `heavy`
                                      // entries (arx_set_tuning("heavy", 8..64)) that different warps reduce

// ---- tiny vector abstraction: VEC = 4 (float4, dim % 4 == 0) or 1 (any dim) --------
template <int VEC> struct V;
template <> struct V<4> {
  using T = float4;
  static __device__ __forceinline__ T zero() { return f4_zero(); }
  static __device__ __forceinline__ T ldg(const float* p) { return ldg_f4(p); }
  static __device__ __forceinline__ T ld(const float* p) { return ld_f4(p); }
  static __device__ __forceinline__ T ldcg(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
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
    // rsqrtf = one MUFU.RSQ (<= 2 ulp); the IEEE sqrt + divide sequences cost ~200 instructions
    // per float4 and made this HBM-bound kernel issue-bound (ncu: 67 M warp instructions)
    w.x = fmaf(-lr * g.x, rsqrtf(a.x), w.x); w.y = fmaf(-lr * g.y, rsqrtf(a.y), w.y);
    w.z = fmaf(-lr * g.z, rsqrtf(a.z), w.z); w.w = fmaf(-lr * g.w, rsqrtf(a.w), w.w);
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
  static __device__ __forceinline__ T ldcg(const float* p) { return __ldcg(p); }
  static __device__ __forceinline__ void st(float* p, T v) { *p = v; }
  static __device__ __forceinline__ void add(T& a, T b) { a += b; }
  static __device__ __forceinline__ void fma(T& a, float w, T b) { a = fmaf(w, b, a); }
  static __device__ __forceinline__ T div(T a, float s) { return a / s; }
  static __device__ __forceinline__ T mul(T a, float s) { return a * s; }
  static __device__ __forceinline__ T shfl_xor(T a, int o) { return __shfl_xor_sync(ARX_FULL_MASK, a, o); }
  static __device__ __forceinline__ float sumsq(T a) { return a * a; }
  static __device__ __forceinline__ void adagrad(T& w, T& a, T g, float lr) {
    a = fmaf(g, g, a);
    w = fmaf(-lr * g, rsqrtf(a), w);
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

// Row sharding: arx_attr_desc.reserved = (G << 16) | r means this GPU stores only the
// rows t with t % G == r, at local row t / G (cyclic: balances the Zipf head).  0 = unsharded.
// Kernels skip the rows they do not own; the host sums / exchanges the partial results.
struct Shard { int G, r; };
__device__ __forceinline__ Shard shard_of(const arx_attr_desc& a) {
  Shard s; s.G = a.reserved >> 16; s.r = a.reserved & 0xffff;
  if (a.lengths_full != nullptr) s.G = 0;      // pre-partitioned CSR: every listed row is local already
  return s;
}
// mean divisor of entity e's bag (the full bag length, also when only the owned rows are listed)
__device__ __forceinline__ int full_len(const arx_attr_desc& a, int e, int L) {
  return a.lengths_full != nullptr ? __ldg(a.lengths_full + e) : L;
}
__device__ __forceinline__ bool owns(const Shard& s, int tok) { return s.G <= 1 || (tok % s.G) == s.r; }
__device__ __forceinline__ int local_row(const Shard& s, int tok) { return s.G <= 1 ? tok : tok / s.G; }

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
// one warp per (entity, attribute) pair
// A CTA owns EPB consecutive entities; its 8 warps stride over the EPB*n_attr bags, each
// bag is gathered with >= 8 rows in flight, divided by its length and parked in shared
// memory; after a barrier the attribute mean (fixed order => deterministic) is written
// with 128-bit stores.  9x more warps than one-warp-per-entity: the 1024-entity sampled
// pool still fills all 148 SMs, and no warp walks nine dependent bags in sequence.
template <int GW, int VEC>
__global__ void __launch_bounds__(256, 4)      // 64 registers: 4 CTAs (32 warps) per SM
pool_fwd_kernel(const arx_attr_desc* __restrict__ g_attrs, int n_attr, int dim,
                 const int* __restrict__ ids, long long n, float* __restrict__ out,
                 long long out_stride, int mode, float* __restrict__ bias_out, int epb) {
  using VT = typename V<VEC>::T;
  extern __shared__ __align__(16) float s_pool[];        // [epb][n_attr][dim] then [epb][n_attr] bias
  __shared__ arx_attr_desc s_attrs[kMaxAttr];
  stage_descs(s_attrs, g_attrs, n_attr);
  constexpr int NG = 32 / GW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int g = lane / GW, l = lane % GW;
  const int nvec = dim / VEC;
  float* s_bias = s_pool + (size_t)epb * n_attr * dim;
  const float Ff = (float)n_attr;

  for (long long e0 = (long long)blockIdx.x * epb; e0 < n; e0 += (long long)gridDim.x * epb) {
    const int ne = (int)min((long long)epb, n - e0);
    const int npairs = ne * n_attr;
    for (int p = warp; p < npairs; p += nw) {
      const int el = p / n_attr, f = p - el * n_attr;
      const int e = __ldg(ids + e0 + el);
      int s = e, L = 1;
      if (s_attrs[f].kind == 1) { s = __ldg(s_attrs[f].starts + e); L = __ldg(s_attrs[f].lengths + e); }
      const float* __restrict__ table = s_attrs[f].table;
      const int* __restrict__ values = s_attrs[f].values;
      const float* __restrict__ bias = s_attrs[f].bias;
      const bool want_bias = (bias_out != nullptr) && (bias != nullptr);
      const Shard sh = shard_of(s_attrs[f]);
      const float Lf = (float)(s_attrs[f].kind == 1 ? full_len(s_attrs[f], e, L) : 1);
      float bsum = 0.f;
      for (int c0 = 0; c0 < nvec; c0 += GW) {
        const int col = c0 + l;
        const bool colok = col < nvec;
        VT acc = V<VEC>::zero();
        for (int j0 = 0; j0 < L; j0 += 32) {
          const int cnt = min(32, L - j0);
          int tok = (lane < cnt) ? __ldg(values + s + j0 + lane) : 0;
          tok = owns(sh, tok) ? local_row(sh, tok) : -1;          // -1: row lives on another GPU
          if (want_bias && c0 == 0 && lane < cnt && tok >= 0) bsum += __ldg(bias + tok);
          for (int jj0 = 0; jj0 < cnt; jj0 += NG * kRowsInFlight) {
            VT v[kRowsInFlight];
#pragma unroll
            for (int u = 0; u < kRowsInFlight; ++u) {
              const int jj = jj0 + u * NG + g;
              const int t = __shfl_sync(ARX_FULL_MASK, tok, jj & 31);
              v[u] = (jj < cnt && colok && t >= 0) ? V<VEC>::ldg(table + (size_t)t * dim + (size_t)col * VEC)
                                                   : V<VEC>::zero();
            }
#pragma unroll
            for (int u = 0; u < kRowsInFlight; ++u) V<VEC>::add(acc, v[u]);
          }
        }
#pragma unroll
        for (int o = GW; o < 32; o <<= 1) V<VEC>::add(acc, V<VEC>::shfl_xor(acc, o));
        acc = V<VEC>::div(acc, Lf);                                  // tf.div(embedded_sum, lengs) :400
        if (g == 0 && colok) {
          if (mode == ARX_POOL_MEAN) V<VEC>::st(s_pool + ((size_t)el * n_attr + f) * dim + (size_t)col * VEC, acc);
          else V<VEC>::st(out + (e0 + el) * out_stride + (size_t)f * dim + (size_t)col * VEC, acc);
        }
      }
      if (bias_out != nullptr) {
        const float b = want_bias ? warp_sum(bsum) / Lf : 0.f;       // :404-406
        if (lane == 0) s_bias[el * n_attr + f] = b;
      }
    }
    __syncthreads();
    if (mode == ARX_POOL_MEAN) {
      for (int i = threadIdx.x; i < ne * nvec; i += blockDim.x) {
        const int el = i / nvec, col = i - el * nvec;
        VT tot = V<VEC>::zero();
        for (int f = 0; f < n_attr; ++f)
          V<VEC>::add(tot, V<VEC>::ld(s_pool + ((size_t)el * n_attr + f) * dim + (size_t)col * VEC));
        V<VEC>::st(out + (e0 + el) * out_stride + (size_t)col * VEC, V<VEC>::div(tot, Ff));   // reduce_mean :219,:235
      }
    }
    if (bias_out != nullptr && threadIdx.x < ne) {
      float b = 0.f;
      for (int f = 0; f < n_attr; ++f) b += s_bias[threadIdx.x * n_attr + f];
      bias_out[e0 + threadIdx.x] = b / Ff;                           // :412
    }
    __syncthreads();
  }
}

// ---- flat forward (mean mode, dim = 128 or 256): rows first, entities second --------------------
// The per-bag kernel above walks "ids -> (start,len) -> tokens -> rows" once per bag: four
// dependent round trips for ~12 rows.  Here a CTA takes a group of entities (<= kFlatBags bags,
// <= kFlatRows rows), resolves EVERY row address of the group in two block-wide passes (all loads
// independent), then its 8 warps split the flat row list evenly (load balance by rows, not by
// bags) and stream it 8 rows at a time.  Pooling is linear, so the warp keeps ONE running sum per
// entity: acc += w_row * row with w_row = 1 / (bag length * #attributes) (the mean over the bag,
// :400, and the mean over attributes, :219/:235, folded into the row weight), flushed at entity
// boundaries only (every ~100 rows, not every ~12).  The bias column rides along as w_row * bias[row].
// An entity that straddles two warps' segments is finished from per-warp partial slots in fixed warp
// order, so the result is deterministic.  ncu (profiles/r1_fwd_flat_*.md): the previous per-bag
// flush + 36 IEEE divisions per output float4 made the kernel issue-bound (short-scoreboard stalls on
// the per-row shared-memory lookups), not HBM-bound.
constexpr int kFlatBags = 64;        // bags per CTA pass (pass-0 threads, scan width)
constexpr int kFlatEnt = 16;         // entities per CTA pass
constexpr int kFlatRows = 1024;
constexpr int kFlatU = 8;

// Push mode (row-sharded tables): instead of storing this rank's PARTIAL pooled vectors into a local
// [n, stride] buffer that a reduce-scatter would then sum over the ranks, pass 3 adds them straight into the OWNER
// rank's receive buffer over NVLink peer memory (entity e belongs to rank e / rows_per_rank; red.global.add.v4.f32,
// system scope).  The receive buffers are zero before the step; a device-side barrier (csrc/peer.cu) orders the adds
// before the consumer.  Entities none of whose rows live here cause no traffic at all.
struct PushDesc {
  float* const* peer;        // device array [G]: base of every rank's receive block for this request (NULL = no push)
  float* const* peer_bias;   // device array [G]: every rank's receive vector of the pooled bias, or NULL
  long long rows_per_rank;   // entities per owner rank; 0 = add every entity into ALL G ranks (all-reduce by push)
  long long stride;          // row pitch of the receive blocks, floats
  int G;
};

__device__ __forceinline__ void red_add_f4_sys(float* p, float4 v) {
  asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

template <int CPL>      // float4 columns per lane: dim = 128 * CPL
__device__ __forceinline__ void
pool_fwd_flat_body(const arx_attr_desc* __restrict__ g_attrs, int n_attr, const int* __restrict__ ids,
                   long long n, float* __restrict__ out, long long out_stride,
                   float* __restrict__ bias_out, int epb, long long vblock, long long vgrid, const PushDesc push) {
  constexpr int dim = 128 * CPL;
  extern __shared__ __align__(16) unsigned char s_raw[];
  __shared__ int s_start[kFlatBags], s_len[kFlatBags], s_off[kFlatBags + 1];
  __shared__ float s_w[kFlatBags];
  __shared__ float s_bias[kFlatEnt], s_partb[8][2];
  __shared__ int s_take;
  float* s_pool = reinterpret_cast<float*>(s_raw);                          // [kFlatEnt][dim]
  float* s_part = s_pool + (size_t)kFlatEnt * dim;                          // [8][2][dim]
  const float** s_rowptr = reinterpret_cast<const float**>(s_part + 16 * dim);   // [kFlatRows]
  float* s_roww = reinterpret_cast<float*>(s_rowptr + kFlatRows);           // [kFlatRows] row weight w
  float* s_rowwb = s_roww + kFlatRows;                                      // [kFlatRows] w * bias[row]
  arx_attr_desc* s_attrs = reinterpret_cast<arx_attr_desc*>(s_rowwb + kFlatRows);  // [n_attr]
  stage_descs(s_attrs, g_attrs, n_attr);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float Ff = (float)n_attr;
  const bool want_bias = bias_out != nullptr;

  for (long long g0 = vblock * epb; g0 < n; g0 += vgrid * epb) {
   const int ng = (int)min((long long)epb, n - g0);
   for (int first = 0; first < ng;) {                 // sub-groups that fit kFlatRows rows
    const long long e0 = g0 + first;
    int ne = ng - first;
    int nb = ne * n_attr;
    // -- pass 0: bag extents and row weights ------------------------------------------------------
    if (tid < nb) {
      const int el = tid / n_attr, f = tid - el * n_attr;
      const int e = __ldg(ids + e0 + el);
      int s = e, L = 1, Lf = 1;
      if (s_attrs[f].kind == 1) {
        s = __ldg(s_attrs[f].starts + e); L = __ldg(s_attrs[f].lengths + e); Lf = full_len(s_attrs[f], e, L);
      }
      s_start[tid] = s; s_len[tid] = L;
      s_w[tid] = 1.0f / ((float)Lf * Ff);             // tf.div by the bag length (:400), reduce_mean over attributes (:219,:235)
    }
    __syncthreads();
    if (warp == 0) {                                   // exclusive scan of <= 64 lengths
      const int a = (lane < nb) ? s_len[lane] : 0, b = (lane + 32 < nb) ? s_len[lane + 32] : 0;
      int ia = a, ib = b;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int ta = __shfl_up_sync(ARX_FULL_MASK, ia, o), tb = __shfl_up_sync(ARX_FULL_MASK, ib, o);
        if (lane >= o) { ia += ta; ib += tb; }
      }
      const int tot_a = __shfl_sync(ARX_FULL_MASK, ia, 31);
      if (lane < nb) s_off[lane] = ia - a;
      if (lane + 32 < nb) s_off[lane + 32] = tot_a + ib - b;
      if (lane == 31) s_off[nb] = tot_a + ib;
    }
    __syncthreads();
    if (tid == 0) {                                    // how many whole entities fit the row buffer
      int k = ne;
      while (k > 1 && s_off[k * n_attr] > kFlatRows) --k;
      s_take = k;
    }
    __syncthreads();
    ne = s_take;
    nb = ne * n_attr;
    first += ne;
    const int R = s_off[nb];
    // -- pass 1: every row address of the group (independent loads) ---------------------------
    for (int r = tid; r < R; r += blockDim.x) {
      int lo = 0, hi = nb;                             // bag of row r: last b with off[b] <= r
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_off[mid] <= r) lo = mid; else hi = mid; }
      const int f = lo % n_attr;
      const int tok = __ldg(s_attrs[f].values + s_start[lo] + (r - s_off[lo]));
      const Shard sh = shard_of(s_attrs[f]);
      const bool own = owns(sh, tok);
      const int lr = local_row(sh, tok);
      const float w = s_w[lo];
      s_rowptr[r] = own ? s_attrs[f].table + (size_t)lr * dim : nullptr;
      s_roww[r] = w;
      if (want_bias) s_rowwb[r] = (own && s_attrs[f].bias != nullptr) ? w * __ldg(s_attrs[f].bias + lr) : 0.f;
    }
    __syncthreads();
    // -- pass 2: stream the flat row list, 8 warps x equal segments ------------------------------
    const int seg = (R + 7) >> 3;
    const int ra = min(R, warp * seg), rb = min(R, ra + seg);
    if (ra < rb) {
      float4 acc[CPL];
#pragma unroll
      for (int c = 0; c < CPL; ++c) acc[c] = f4_zero();
      float accb = 0.f;
      int cur = 0;                                     // entity of row ra
      while (cur + 1 < ne && s_off[(cur + 1) * n_attr] <= ra) ++cur;
      int next = s_off[(cur + 1) * n_attr];            // first row of the next entity
      auto flush = [&](int ent) {
        const int o = s_off[ent * n_attr], e = s_off[(ent + 1) * n_attr];
        float* dst; float* dstb;
        if (o >= ra && e <= rb) { dst = s_pool + (size_t)ent * dim; dstb = s_bias + ent; }     // entity wholly mine
        else {
          const int slot = (e > rb) ? 1 : 0;           // continues in the next warp : started before me
          dst = s_part + ((size_t)warp * 2 + slot) * dim; dstb = &s_partb[warp][slot];
        }
#pragma unroll
        for (int c = 0; c < CPL; ++c) { st_f4(dst + (size_t)(c * 32 + lane) * 4, acc[c]); acc[c] = f4_zero(); }
        if (lane == 0) *dstb = accb;
        accb = 0.f;
      };
      for (int r0 = ra; r0 < rb; r0 += kFlatU) {
        const float* p[kFlatU];
#pragma unroll
        for (int u = 0; u < kFlatU; ++u) p[u] = (r0 + u < rb) ? s_rowptr[r0 + u] : nullptr;
        float4 v[kFlatU][CPL];
#pragma unroll
        for (int u = 0; u < kFlatU; ++u) {
#pragma unroll
          for (int c = 0; c < CPL; ++c) v[u][c] = p[u] ? ldg_f4(p[u] + (size_t)(c * 32 + lane) * 4) : f4_zero();
        }
        float wv[kFlatU];
#pragma unroll
        for (int u = 0; u < kFlatU; ++u) wv[u] = (r0 + u < rb) ? s_roww[r0 + u] : 0.f;
#pragma unroll
        for (int u = 0; u < kFlatU; ++u) {
          if (r0 + u < rb) {
            while (r0 + u >= next) { flush(cur); ++cur; next = s_off[(cur + 1) * n_attr]; }
#pragma unroll
            for (int c = 0; c < CPL; ++c) f4_fma(acc[c], wv[u], v[u][c]);
            if (want_bias) accb += s_rowwb[r0 + u];
          }
        }
      }
      flush(cur);
    }
    __syncthreads();
    // -- pass 3: finish the entities that straddle warps -----------------------------------------
    for (int i = tid; i < ne * 32 * CPL; i += blockDim.x) {
      const int el = i / (32 * CPL), col = i - el * 32 * CPL;
      const int o = s_off[el * n_attr], L = s_off[(el + 1) * n_attr] - o;
      float4 tot = f4_zero();
      if (L > 0) {                                       // L == 0 (sharded): none of the entity's rows live here
        const int wlo = o / seg, whi = (o + L - 1) / seg;
        if (wlo == whi) tot = ld_f4(s_pool + (size_t)el * dim + (size_t)col * 4);
        else
          for (int w = wlo; w <= whi; ++w)               // fixed warp order
            f4_add(tot, ld_f4(s_part + ((size_t)w * 2 + (w < whi ? 1 : 0)) * dim + (size_t)col * 4));
      }
      if (push.peer == nullptr) st_f4(out + (e0 + el) * out_stride + (size_t)col * 4, tot);
      else if (L > 0) {
        if (push.rows_per_rank > 0) {
          const long long owner = (e0 + el) / push.rows_per_rank;
          red_add_f4_sys(push.peer[owner] + ((e0 + el) - owner * push.rows_per_rank) * push.stride + (size_t)col * 4, tot);
        } else {
          for (int g = 0; g < push.G; ++g) red_add_f4_sys(push.peer[g] + (e0 + el) * push.stride + (size_t)col * 4, tot);
        }
      }
    }
    if (want_bias && tid < ne) {
      const int o = s_off[tid * n_attr], L = s_off[(tid + 1) * n_attr] - o;
      float bt = 0.f;
      if (L > 0) {
        const int wlo = o / seg, whi = (o + L - 1) / seg;
        if (wlo == whi) bt = s_bias[tid];
        else for (int w = wlo; w <= whi; ++w) bt += s_partb[w][w < whi ? 1 : 0];
      }
      if (push.peer == nullptr) bias_out[e0 + tid] = bt;                                  // :404-412
      else if (L > 0 && push.peer_bias != nullptr) {
        if (push.rows_per_rank > 0) {
          const long long owner = (e0 + tid) / push.rows_per_rank;
          asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(push.peer_bias[owner] + ((e0 + tid) - owner * push.rows_per_rank)),
                       "f"(bt) : "memory");
        } else {
          for (int g = 0; g < push.G; ++g)
            asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(push.peer_bias[g] + (e0 + tid)), "f"(bt) : "memory");
        }
      }
    }
    __syncthreads();
   }
  }
}

template <int CPL>
__global__ void __launch_bounds__(256, CPL == 1 ? 4 : 2)
pool_fwd_flat_kernel(const arx_attr_desc* __restrict__ g_attrs, int n_attr, const int* __restrict__ ids,
                     long long n, float* __restrict__ out, long long out_stride,
                     float* __restrict__ bias_out, int epb) {
  pool_fwd_flat_body<CPL>(g_attrs, n_attr, ids, n, out, out_stride, bias_out, epb, blockIdx.x, gridDim.x, PushDesc{});
}

// Several independent lookups (e.g. the users, the target items and the sampled pool of one training step) in ONE
// launch: each request owns a contiguous range of CTAs.  One launch ramp and one tail instead of one per lookup, and
// the CTAs of a short lookup fill the slots the long one leaves free.
constexpr int kManyReqs = 4;
struct PoolManyParams {
  const arx_attr_desc* attrs[kManyReqs];
  const int* ids[kManyReqs];
  float* out[kManyReqs];
  float* bias_out[kManyReqs];
  long long n[kManyReqs];
  long long out_stride[kManyReqs];
  int n_attr[kManyReqs];
  int epb[kManyReqs];
  int block_end[kManyReqs];        // exclusive prefix of CTAs per request
  PushDesc push[kManyReqs];
  int n_req;
};

template <int CPL>
__global__ void __launch_bounds__(256, CPL == 1 ? 4 : 2)
pool_fwd_flat_many_kernel(const PoolManyParams mp) {
  int r = 0;
  while (r + 1 < mp.n_req && (int)blockIdx.x >= mp.block_end[r]) ++r;
  const int b0 = r == 0 ? 0 : mp.block_end[r - 1];
  pool_fwd_flat_body<CPL>(mp.attrs[r], mp.n_attr[r], mp.ids[r], mp.n[r], mp.out[r], mp.out_stride[r], mp.bias_out[r],
                          mp.epb[r], (long long)blockIdx.x - b0, (long long)(mp.block_end[r] - b0), mp.push[r]);
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
  // one warp per (entity, attribute) bag; first-touch rows claim their slot in the unique
  // list with ONE counter atomic per warp iteration (ballot + popc), not one per row.
  const long long npairs = n * n_attr;
  for (long long p = warp0; p < npairs; p += nwarps) {
    const long long ei = p / n_attr;
    const int f = (int)(p - ei * n_attr);
    const int e = __ldg(ids + ei);
    int s = e, L = 1;
    if (s_attrs[f].kind == 1) { s = __ldg(s_attrs[f].starts + e); L = __ldg(s_attrs[f].lengths + e); }
    const int* __restrict__ values = s_attrs[f].values;
    int* touch = s_attrs[f].touch;
    const Shard sh = shard_of(s_attrs[f]);
    for (int j0 = 0; j0 < L; j0 += 32) {
      const int j = j0 + lane;
      int tok = 0;
      bool first = false;
      if (j < L) {
        tok = __ldg(values + s + j);
        if (owns(sh, tok)) first = (atomicAdd(&touch[tok], 1) == 0);
      }
      const unsigned m = __ballot_sync(ARX_FULL_MASK, first);
      if (m != 0u) {
        int ub = 0;
        if (lane == 0) ub = atomicAdd(&plan.counters[0], __popc(m));
        ub = __shfl_sync(ARX_FULL_MASK, ub, 0);
        if (first) {
          const int u = ub + __popc(m & ((1u << lane) - 1u));
          if (u < plan.cap_rows) { plan.uniq_tok[u] = tok; plan.uniq_attr[u] = attr_begin + f; }
          else plan.counters[2] = 1;
        }
      }
    }
  }
}

__global__ void plan_alloc_kernel(const arx_attr_desc* __restrict__ g_attrs, arx_bwd_plan plan, int kHeavy) {
  const int nu = (int)min((long long)plan.counters[0], (long long)plan.cap_rows);
  const int lane = threadIdx.x & 31;
  for (int u0 = blockIdx.x * blockDim.x + threadIdx.x - lane; u0 < nu; u0 += gridDim.x * blockDim.x) {
    const int u = u0 + lane;
    int tok = 0, c = 0;
    int* touch = nullptr;
    if (u < nu) {
      tok = plan.uniq_tok[u];
      touch = g_attrs[plan.uniq_attr[u]].touch;
      c = touch[tok];
    }
    int incl = c;                                   // warp scan: one cursor atomic per 32 rows
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(ARX_FULL_MASK, incl, o);
      if (lane >= o) incl += t;
    }
    int wb = 0;
    if (lane == 31) wb = atomicAdd(&plan.counters[1], incl);
    wb = __shfl_sync(ARX_FULL_MASK, wb, 31);
    if (u >= nu) continue;
    const int base = wb + incl - c;
    plan.row_base[u] = base;
    plan.row_cnt[u] = c;
    touch[tok] = base;                       // becomes the fill cursor of this row
    if ((long long)base + c > plan.cap_occ) plan.counters[2] = 1;
    // hot rows (Zipf heads: thousands of contributions) are split into kHeavy-sized chunks
    // that different warps reduce; the last chunk to arrive folds the partials and updates.
    if (c > kHeavy) {
      const int nch = (c + kHeavy - 1) / kHeavy;
      const int cb = atomicAdd(&plan.counters[4], nch);
      plan.row_chunk0[u] = cb;
      if ((long long)cb + nch > plan.cap_chunks) { plan.counters[2] = 1; }
      else for (int k = 0; k < nch; ++k) plan.chunk_row[cb + k] = u;
    }
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
  const long long npairs = n * n_attr;
  for (long long p = warp0; p < npairs; p += nwarps) {
    const long long ei = p / n_attr;
    const int f = (int)(p - ei * n_attr);
    const int e = __ldg(ids + ei);
    int s = e, L = 1;
    int Lf = 1;
    if (s_attrs[f].kind == 1) {
      s = __ldg(s_attrs[f].starts + e); L = __ldg(s_attrs[f].lengths + e); Lf = full_len(s_attrs[f], e, L);
    }
    const int* __restrict__ values = s_attrs[f].values;
    int* touch = s_attrs[f].touch;
    const float w = invF / (float)Lf;
    const int row = (mode == ARX_POOL_MEAN) ? (int)(row_base + ei) : (int)(row_base + ei * n_attr + f);
    const Shard sh = shard_of(s_attrs[f]);
    for (int j = lane; j < L; j += 32) {
      const int tok = __ldg(values + s + j);
      if (!owns(sh, tok)) continue;
      const int pos = atomicAdd(&touch[tok], 1);
      plan.bucket_src[pos] = row;
      plan.bucket_w[pos] = w;
    }
  }
}

// ---- block-aggregated plan kernels --------------------------------------------------------------
// The warp-per-bag kernels above issue one global atomic per OCCURRENCE on touch[token]; with Zipf tokens the
// head token of every multi-hot table takes ~8 % of them (32 k same-address atomics per launch at C2), which
// serialise in L2 and stretch these kernels to 35-55 us right where the step's dependent chain needs the SMs
// (tools/trace_step.py).  Here a CTA resolves a group of <= 64 bags block-wide (as pool_fwd_flat_kernel does),
// aggregates its occurrences per (table, token) in a shared-memory hash table and issues ONE global atomic per
// distinct token of the group: the head tokens drop from one atomic per occurrence to one per CTA.
constexpr int kAggBags = 64;
constexpr int kAggRows = 1024;                 // occurrences per hash round
constexpr int kAggSlots = 2048;                // hash slots (load factor <= 0.5)
constexpr unsigned kAggEmpty = 0xffffffffu;

__device__ __forceinline__ int agg_insert(unsigned* keys, unsigned key) {
  unsigned slot = (key * 2654435761u) >> 21;   // 11 bits
  while (true) {
    const unsigned prev = atomicCAS(&keys[slot], kAggEmpty, key);
    if (prev == kAggEmpty || prev == key) return (int)slot;
    slot = (slot + 1) & (kAggSlots - 1);
  }
}

struct AggShared {
  arx_attr_desc attrs[kMaxAttr];
  int start[kAggBags], len[kAggBags], lenf[kAggBags], off[kAggBags + 1];
  unsigned keys[kAggSlots];
  int cnt[kAggSlots];
  int base[kAggSlots];
  int n_first, first_base;
};

// bag extents of entities [e0, e0 + ne) x n_attr attributes -> sh.start/len/lenf/off; returns the row count
__device__ __forceinline__ int agg_resolve(AggShared& sh, int n_attr, const int* __restrict__ ids, long long e0, int ne) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = ne * n_attr;
  if (tid < nb) {
    const int el = tid / n_attr, f = tid - el * n_attr;
    const int e = __ldg(ids + e0 + el);
    int s = e, L = 1, Lf = 1;
    if (sh.attrs[f].kind == 1) {
      s = __ldg(sh.attrs[f].starts + e); L = __ldg(sh.attrs[f].lengths + e); Lf = full_len(sh.attrs[f], e, L);
    }
    sh.start[tid] = s; sh.len[tid] = L; sh.lenf[tid] = Lf;
  }
  __syncthreads();
  if (warp == 0) {
    const int a = (lane < nb) ? sh.len[lane] : 0, b = (lane + 32 < nb) ? sh.len[lane + 32] : 0;
    int ia = a, ib = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int ta = __shfl_up_sync(ARX_FULL_MASK, ia, o), tb = __shfl_up_sync(ARX_FULL_MASK, ib, o);
      if (lane >= o) { ia += ta; ib += tb; }
    }
    const int tot_a = __shfl_sync(ARX_FULL_MASK, ia, 31);
    if (lane < nb) sh.off[lane] = ia - a;
    if (lane + 32 < nb) sh.off[lane + 32] = tot_a + ib - b;
    if (lane == 31) sh.off[nb] = tot_a + ib;
  }
  __syncthreads();
  return sh.off[nb];
}

// row r of the group -> (bag, token); returns false when the row is not owned by this GPU
__device__ __forceinline__ bool agg_row(const AggShared& sh, int n_attr, int nb, int r, int& bag, int& f, int& tok) {
  int lo = 0, hi = nb;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (sh.off[mid] <= r) lo = mid; else hi = mid; }
  bag = lo; f = lo % n_attr;
  tok = __ldg(sh.attrs[f].values + sh.start[lo] + (r - sh.off[lo]));
  return owns(shard_of(sh.attrs[f]), tok);
}

__global__ void __launch_bounds__(256)
plan_count_agg_kernel(const arx_attr_desc* __restrict__ g_attrs, int attr_begin, int n_attr,
                      const int* __restrict__ ids, long long n, arx_bwd_plan plan, int epb) {
  __shared__ AggShared sh;
  stage_descs(sh.attrs, g_attrs + attr_begin, n_attr);
  const int tid = threadIdx.x;
  for (long long e0 = (long long)blockIdx.x * epb; e0 < n; e0 += (long long)gridDim.x * epb) {
    const int ne = (int)min((long long)epb, n - e0);
    const int nb = ne * n_attr;
    const int R = agg_resolve(sh, n_attr, ids, e0, ne);
    for (int rb = 0; rb < R; rb += kAggRows) {
      for (int i = tid; i < kAggSlots; i += blockDim.x) { sh.keys[i] = kAggEmpty; sh.cnt[i] = 0; }
      if (tid == 0) sh.n_first = 0;
      __syncthreads();
      const int re = min(R, rb + kAggRows);
      for (int r = rb + tid; r < re; r += blockDim.x) {
        int bag, f, tok;
        if (agg_row(sh, n_attr, nb, r, bag, f, tok)) atomicAdd(&sh.cnt[agg_insert(sh.keys, ((unsigned)f << 27) | (unsigned)tok)], 1);
      }
      __syncthreads();
      // one global atomic per distinct (table, token) of the group; first-touch rows join the unique list
      int myfirst[kAggSlots / 256];
      int nf = 0;
#pragma unroll
      for (int q = 0; q < kAggSlots / 256; ++q) {          // all global atomics in flight before any result is used
        const int i = q * 256 + tid;
        const unsigned key = sh.keys[i];
        myfirst[q] = 1;
        if (key != kAggEmpty) myfirst[q] = atomicAdd(&sh.attrs[key >> 27].touch[key & 0x7ffffffu], sh.cnt[i]);
      }
#pragma unroll
      for (int q = 0; q < kAggSlots / 256; ++q) {
        if (myfirst[q] == 0) { myfirst[q] = atomicAdd(&sh.n_first, 1); ++nf; }
        else myfirst[q] = -1;
      }
      __syncthreads();
      if (tid == 0 && sh.n_first > 0) sh.first_base = atomicAdd(&plan.counters[0], sh.n_first);
      __syncthreads();
      if (nf > 0) {
#pragma unroll
        for (int q = 0; q < kAggSlots / 256; ++q) {
          if (myfirst[q] < 0) continue;
          const unsigned key = sh.keys[q * 256 + tid];
          const int u = sh.first_base + myfirst[q];
          if (u < plan.cap_rows) { plan.uniq_tok[u] = (int)(key & 0x7ffffffu); plan.uniq_attr[u] = attr_begin + (int)(key >> 27); }
          else plan.counters[2] = 1;
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(256)
plan_fill_agg_kernel(const arx_attr_desc* __restrict__ g_attrs, int attr_begin, int n_attr,
                     const int* __restrict__ ids, long long n, int mode, long long row_base,
                     arx_bwd_plan plan, int epb) {
  __shared__ AggShared sh;
  stage_descs(sh.attrs, g_attrs + attr_begin, n_attr);
  if (plan.counters[2] != 0) return;        // capacity exceeded: leave buckets untouched
  const int tid = threadIdx.x;
  const float invF = (mode == ARX_POOL_MEAN) ? 1.0f / (float)n_attr : 1.0f;
  for (long long e0 = (long long)blockIdx.x * epb; e0 < n; e0 += (long long)gridDim.x * epb) {
    const int ne = (int)min((long long)epb, n - e0);
    const int nb = ne * n_attr;
    const int R = agg_resolve(sh, n_attr, ids, e0, ne);
    for (int rb = 0; rb < R; rb += kAggRows) {
      for (int i = tid; i < kAggSlots; i += blockDim.x) { sh.keys[i] = kAggEmpty; sh.cnt[i] = 0; }
      __syncthreads();
      const int re = min(R, rb + kAggRows);
      int slot_[kAggRows / 256], rank_[kAggRows / 256], row_[kAggRows / 256];
      float w_[kAggRows / 256];
#pragma unroll
      for (int q = 0; q < kAggRows / 256; ++q) {
        const int r = rb + q * 256 + tid;
        slot_[q] = -1;
        if (r < re) {
          int bag, f, tok;
          if (agg_row(sh, n_attr, nb, r, bag, f, tok)) {
            slot_[q] = agg_insert(sh.keys, ((unsigned)f << 27) | (unsigned)tok);
            rank_[q] = atomicAdd(&sh.cnt[slot_[q]], 1);
            const long long ei = e0 + bag / n_attr;
            row_[q] = (mode == ARX_POOL_MEAN) ? (int)(row_base + ei) : (int)(row_base + ei * n_attr + f);
            w_[q] = invF / (float)sh.lenf[bag];
          }
        }
      }
      __syncthreads();
      {                                                           // reserve cnt consecutive bucket slots per token
        int b_[kAggSlots / 256];
#pragma unroll
        for (int q = 0; q < kAggSlots / 256; ++q) {               // independent atomics, all in flight together
          const int i = q * 256 + tid;
          const unsigned key = sh.keys[i];
          b_[q] = 0;
          if (key != kAggEmpty) b_[q] = atomicAdd(&sh.attrs[key >> 27].touch[key & 0x7ffffffu], sh.cnt[i]);
        }
#pragma unroll
        for (int q = 0; q < kAggSlots / 256; ++q) sh.base[q * 256 + tid] = b_[q];
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < kAggRows / 256; ++q) {
        if (slot_[q] < 0) continue;
        const int pos = sh.base[slot_[q]] + rank_[q];
        plan.bucket_src[pos] = row_[q];
        plan.bucket_w[pos] = w_[q];
      }
      __syncthreads();
    }
  }
}

__global__ void plan_reset_kernel(const arx_attr_desc* __restrict__ g_attrs, arx_bwd_plan plan) {
  const int nu = (int)min((long long)plan.counters[0], (long long)plan.cap_rows);
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < nu; u += gridDim.x * blockDim.x)
    g_attrs[plan.uniq_attr[u]].touch[plan.uniq_tok[u]] = 0;
}

// ------------------------------------------------------------------ backward apply --
// (a) hot rows are reduced chunk-by-chunk by many warps (last arriver folds + updates);
// (b) the remaining rows are taken 32 at a time: one coalesced metadata load and one
//     first-bucket-entry load per lane, then four rows per step with their table row,
//     accumulator row and gradient row all in flight together (12 x 512 B per warp).
template <int VEC>
__device__ __forceinline__ typename V<VEC>::T
reduce_bucket(const arx_bwd_plan& plan, int base, int cnt, const float* __restrict__ dout,
              long long stride, int col, bool colok, int lane, const float* __restrict__ dbias, float& gb) {
  using VT = typename V<VEC>::T;
  VT g = V<VEC>::zero();
  float gbl = 0.f;                      // bias column: one bucket entry per lane, reduced at the end
  for (int k0 = 0; k0 < cnt; k0 += 32) {
    const int kc = min(32, cnt - k0);
    int src = 0; float w = 0.f;
    if (lane < kc) {
      src = __ldg(plan.bucket_src + base + k0 + lane); w = __ldg(plan.bucket_w + base + k0 + lane);
      if (dbias != nullptr) gbl = fmaf(w, __ldg(dbias + src), gbl);
    }
    for (int kk0 = 0; kk0 < kc; kk0 += kRowsInFlight) {
      VT v[kRowsInFlight]; float wk[kRowsInFlight];
#pragma unroll
      for (int q = 0; q < kRowsInFlight; ++q) {
        const int kk = kk0 + q;
        const int sk = __shfl_sync(ARX_FULL_MASK, src, kk & 31);
        wk[q] = __shfl_sync(ARX_FULL_MASK, w, kk & 31);
        v[q] = (kk < kc && colok) ? V<VEC>::ldg(dout + (size_t)sk * stride + (size_t)col * VEC) : V<VEC>::zero();
      }
#pragma unroll
      for (int q = 0; q < kRowsInFlight; ++q) V<VEC>::fma(g, wk[q], v[q]);
    }
  }
  gb = (dbias != nullptr) ? warp_sum(gbl) : 0.f;
  return g;
}

__device__ __forceinline__ void bias_update(const arx_attr_desc& a, int tok, float gb, float lr, int opt) {
  if (opt == ARX_OPT_ADAGRAD) {
    const float acc = fmaf(gb, gb, a.bias_acc[tok]);
    a.bias_acc[tok] = acc;
    a.bias[tok] -= lr * gb / sqrtf(acc);
  } else if (opt == ARX_OPT_SGD) {
    a.bias[tok] -= lr * gb;
  }
}

template <int VEC>
__device__ __forceinline__ void row_update(const arx_attr_desc& a, size_t off, typename V<VEC>::T g,
                                           float lr, int opt) {
  using VT = typename V<VEC>::T;
  float* wp = a.table + off;
  VT wv = V<VEC>::ld(wp);
  if (opt == ARX_OPT_ADAGRAD) {
    float* ap = a.table_acc + off;
    VT av = V<VEC>::ld(ap);
    V<VEC>::adagrad(wv, av, g, lr);
    V<VEC>::st(ap, av);
  } else {
    V<VEC>::sgd(wv, g, lr);
  }
  V<VEC>::st(wp, wv);
}

constexpr int kRowsPerStep = 4;

constexpr int kFlatExtra = 128;       // extra bucket entries (beyond the first of each row) a warp pre-loads per 32 rows

template <int VEC, bool FLAT>
__device__ __forceinline__ void
pool_bwd_apply_body(const arx_attr_desc* __restrict__ s_attrs, int dim,
                    const arx_bwd_plan& plan, const float* __restrict__ dout, long long dout_stride,
                    const float* __restrict__ dbias, float lr,
                    const float* __restrict__ grad_scale_dev, int opt, float* __restrict__ rows_out,
                    float* __restrict__ bias_rows_out, int kHeavyArg) {
  using VT = typename V<VEC>::T;
  const int kHeavy = kHeavyArg & 0xff;
  if (plan.counters[2] != 0) return;
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nvec = dim / VEC;
  const int nu = (int)min((long long)plan.counters[0], (long long)plan.cap_rows);
  const int nchunks = (int)min((long long)plan.counters[4], (long long)plan.cap_chunks);
  const float gs = grad_scale_dev ? __ldg(grad_scale_dev) : 1.0f;
  float* __restrict__ part = plan.partials;
  float* __restrict__ part_b = plan.partials + (size_t)plan.cap_chunks * dim;

  // measurement only (arx_set_tuning("apply_phases", 1 | 2 | 3)): bits 8-9 of the argument switch a phase off
  const int phase_off = (kHeavyArg >> 8) & 3;
  // ---- phase 1: chunks of the hot rows (longest work first) ---------------------------
  for (long long ch = warp0; ch < ((phase_off & 1) ? 0 : nchunks); ch += nwarps) {
    const int u = plan.chunk_row[ch];
    const int cb = plan.row_chunk0[u];
    const int k = (int)ch - cb;
    const int rc = plan.row_cnt[u];
    const int base = plan.row_base[u] + k * kHeavy;
    const int cnt = min(kHeavy, rc - k * kHeavy);
    const int nch = (rc + kHeavy - 1) / kHeavy;
    for (int c0 = 0; c0 < nvec; c0 += 32) {
      const int col = c0 + lane;
      const bool colok = col < nvec;
      float gb_unused;
      const VT g = reduce_bucket<VEC>(plan, base, cnt, dout, dout_stride, col, colok, lane, nullptr, gb_unused);
      if (colok) V<VEC>::st(part + (size_t)ch * dim + (size_t)col * VEC, g);
    }
    if (dbias != nullptr) {
      float gb = 0.f;
      for (int kk = lane; kk < cnt; kk += 32)
        gb = fmaf(__ldg(plan.bucket_w + base + kk), __ldg(dbias + __ldg(plan.bucket_src + base + kk)), gb);
      gb = warp_sum(gb);
      if (lane == 0) part_b[ch] = gb;
    }
    __threadfence();
    int last = 0;
    if (lane == 0) last = (atomicAdd(&plan.row_done[u], 1) == nch - 1) ? 1 : 0;
    last = __shfl_sync(ARX_FULL_MASK, last, 0);
    if (!last) continue;
    __threadfence();
    const int f = plan.uniq_attr[u];
    const int tok = local_row(shard_of(s_attrs[f]), plan.uniq_tok[u]);   // row inside this GPU's shard
    for (int c0 = 0; c0 < nvec; c0 += 32) {
      const int col = c0 + lane;
      const bool colok = col < nvec;
      VT g = V<VEC>::zero();
      if (colok) {
        for (int kk0 = 0; kk0 < nch; kk0 += kRowsInFlight) {      // fixed chunk order
          VT v[kRowsInFlight];
#pragma unroll
          for (int q = 0; q < kRowsInFlight; ++q)
            v[q] = (kk0 + q < nch) ? V<VEC>::ldcg(part + (size_t)(cb + kk0 + q) * dim + (size_t)col * VEC)
                                   : V<VEC>::zero();
#pragma unroll
          for (int q = 0; q < kRowsInFlight; ++q) V<VEC>::add(g, v[q]);
        }
        g = V<VEC>::mul(g, gs);
        if (opt == ARX_OPT_NONE) V<VEC>::st(rows_out + (size_t)u * dim + (size_t)col * VEC, g);
        else row_update<VEC>(s_attrs[f], (size_t)tok * dim + (size_t)col * VEC, g, lr, opt);
      }
    }
    if (lane == 0) {
      float gb = 0.f;
      if (dbias != nullptr && s_attrs[f].bias != nullptr) {
        for (int kk = 0; kk < nch; ++kk) gb += __ldcg(part_b + cb + kk);
        gb *= gs;
        bias_update(s_attrs[f], tok, gb, lr, opt);
      }
      if (opt == ARX_OPT_NONE && bias_rows_out != nullptr) bias_rows_out[u] = gb;
      plan.row_done[u] = 0;                                       // self-resetting for the next apply
    }
  }

  // ---- phase 2: all other rows, 32 per warp iteration -----------------------------------
  // Rows are dealt to the warps STRIDED: lane l of warp-iteration `it` takes row l * n_iter + it, so neighbours in the
  // unique list (first-touch order: the ~100 rows of one popular entity, each with as many contributions as the entity
  // has occurrences in the batch) land in ~100 different warps instead of filling three warps with 30 x the average
  // work — those few warps used to set the duration of the item-side launch (Zipf item popularity).
  const int n_iter = (nu + 31) >> 5;
  const bool strided = ((kHeavyArg >> 10) & 1) == 0;
  for (long long it = warp0; it < ((phase_off & 2) ? 0 : n_iter); it += nwarps) {
    const int u = strided ? lane * n_iter + (int)it : (int)it * 32 + lane;
    int tok = 0, f = 0, base = 0, cnt = 0;
    if (u < nu) {
      f = __ldg(plan.uniq_attr + u);
      tok = local_row(shard_of(s_attrs[f]), __ldg(plan.uniq_tok + u));
      base = __ldg(plan.row_base + u); cnt = __ldg(plan.row_cnt + u);
    }
    const bool light = (u < nu) && (cnt <= kHeavy);
    int src0 = 0; float w0 = 0.f;
    if (light) { src0 = __ldg(plan.bucket_src + base); w0 = __ldg(plan.bucket_w + base); }
    // bias column: the lane that owns a row takes its first bucket entry here, the remaining
    // entries are added by reduce_bucket below (one entry per lane, no dependent walk)
    const bool row_bias = light && dbias != nullptr && s_attrs[f].bias != nullptr;
    float gb_mine = row_bias ? w0 * __ldg(dbias + src0) : 0.f;
    const unsigned lightmask = __ballot_sync(ARX_FULL_MASK, light);
    // FLAT: the bucket entries beyond the first of the warp's 32 rows form (up to the gaps of hot rows) one contiguous
    // run of the bucket arrays.  All of them are fetched here in one round of independent, mostly coalesced loads
    // (entry i of the run sits in lane i & 31, slot i >> 5), so that the row steps below never wait on a
    // bucket_src -> gradient-row chain: rows with several contributions (the item side: 2.7 per row) cost one extra
    // wave of gradient loads instead of two dependent round trips per row, one row after the other.
    const int extra = light ? cnt - 1 : 0;
    int incl = extra;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(ARX_FULL_MASK, incl, o);
      if (lane >= o) incl += t;
    }
    const int excl = incl - extra;
    const int n_extra = __shfl_sync(ARX_FULL_MASK, incl, 31);
    const bool flat = FLAT && n_extra <= kFlatExtra;
    int es[kFlatExtra / 32]; float ew[kFlatExtra / 32], eb[kFlatExtra / 32];
    if (flat) {
#pragma unroll
      for (int j = 0; j < kFlatExtra / 32; ++j) {
        const int i = lane + 32 * j;
        int lo = 0;                                   // row (lane) that owns entry i: first lane with incl > i
#pragma unroll
        for (int st = 16; st > 0; st >>= 1) {
          const int t = __shfl_sync(ARX_FULL_MASK, incl, lo + st - 1);
          if (t <= i) lo += st;
        }
        const int rb = __shfl_sync(ARX_FULL_MASK, base, lo), rx = __shfl_sync(ARX_FULL_MASK, excl, lo);
        const int hasb = __shfl_sync(ARX_FULL_MASK, (int)row_bias, lo);
        es[j] = 0; ew[j] = 0.f; eb[j] = 0.f;
        if (i < n_extra) {
          es[j] = __ldg(plan.bucket_src + rb + 1 + (i - rx));
          ew[j] = __ldg(plan.bucket_w + rb + 1 + (i - rx));
          if (hasb) eb[j] = ew[j] * __ldg(dbias + es[j]);
        }
      }
    }
    for (int c0 = 0; c0 < nvec; c0 += 32) {
      const int col = c0 + lane;
      const bool colok = col < nvec;
      for (int r = 0; r < 32; r += kRowsPerStep) {
        if (((lightmask >> r) & ((1u << kRowsPerStep) - 1u)) == 0u) continue;
        VT Ev[kRowsPerStep], Av[kRowsPerStep], g[kRowsPerStep];
        int tq[kRowsPerStep], fq[kRowsPerStep], cq[kRowsPerStep], bq[kRowsPerStep];
#pragma unroll
        for (int q = 0; q < kRowsPerStep; ++q) {
          const int rr = r + q;
          tq[q] = __shfl_sync(ARX_FULL_MASK, tok, rr);
          fq[q] = __shfl_sync(ARX_FULL_MASK, f, rr);
          cq[q] = ((lightmask >> rr) & 1u) ? __shfl_sync(ARX_FULL_MASK, cnt, rr) : 0;   // 0 = skip row
          bq[q] = __shfl_sync(ARX_FULL_MASK, base, rr);
          const int s0 = __shfl_sync(ARX_FULL_MASK, src0, rr);
          const float wq = __shfl_sync(ARX_FULL_MASK, w0, rr);
          const bool ok = (cq[q] > 0) && colok;
          const size_t off = (size_t)tq[q] * dim + (size_t)col * VEC;
          Ev[q] = V<VEC>::zero(); Av[q] = V<VEC>::zero();
          if (ok && opt != ARX_OPT_NONE) {
            Ev[q] = V<VEC>::ld(s_attrs[fq[q]].table + off);
            if (opt == ARX_OPT_ADAGRAD) Av[q] = V<VEC>::ld(s_attrs[fq[q]].table_acc + off);
          }
          g[q] = ok ? V<VEC>::mul(V<VEC>::ldg(dout + (size_t)s0 * dout_stride + (size_t)col * VEC), wq)
                    : V<VEC>::zero();
        }
        if (flat) {
          int ie[kRowsPerStep];                       // running ends of the rows' extra entries inside the run
#pragma unroll
          for (int q = 0; q < kRowsPerStep; ++q) ie[q] = __shfl_sync(ARX_FULL_MASK, incl, r + q);
          const int i_end = ie[kRowsPerStep - 1];
          for (int i0 = __shfl_sync(ARX_FULL_MASK, excl, r); i0 < i_end; i0 += kRowsInFlight) {
            VT v[kRowsInFlight]; float wk[kRowsInFlight]; int qk[kRowsInFlight];
#pragma unroll
            for (int k = 0; k < kRowsInFlight; ++k) {
              const int i = i0 + k;
              const bool ok = i < i_end;
              const int j = i >> 5;                   // warp-uniform slot
              const int sj = j == 0 ? es[0] : (j == 1 ? es[1] : (j == 2 ? es[2] : es[3]));
              const float wj = j == 0 ? ew[0] : (j == 1 ? ew[1] : (j == 2 ? ew[2] : ew[3]));
              const int sk = __shfl_sync(ARX_FULL_MASK, sj, i & 31);
              wk[k] = ok ? __shfl_sync(ARX_FULL_MASK, wj, i & 31) : 0.f;
              qk[k] = 0;
#pragma unroll
              for (int q = 0; q < kRowsPerStep - 1; ++q) qk[k] += (i >= ie[q]) ? 1 : 0;
              v[k] = (ok && colok) ? V<VEC>::ldg(dout + (size_t)sk * dout_stride + (size_t)col * VEC) : V<VEC>::zero();
              if (c0 == 0 && dbias != nullptr) {
                const float bj = j == 0 ? eb[0] : (j == 1 ? eb[1] : (j == 2 ? eb[2] : eb[3]));
                const float bt = __shfl_sync(ARX_FULL_MASK, bj, i & 31);
                if (ok && lane == r + qk[k]) gb_mine += bt;
              }
            }
#pragma unroll
            for (int k = 0; k < kRowsInFlight; ++k) {
#pragma unroll
              for (int q = 0; q < kRowsPerStep; ++q)          // arithmetic select: keeps g[] in registers
                V<VEC>::fma(g[q], qk[k] == q ? wk[k] : 0.f, v[k]);
            }
          }
        } else {
#pragma unroll
        for (int q = 0; q < kRowsPerStep; ++q) {
          if (cq[q] > 1) {                           // warp-uniform: remaining bucket entries
            const bool wb = (c0 == 0) && dbias != nullptr && s_attrs[fq[q]].bias != nullptr;
            float gbx;
            V<VEC>::add(g[q], reduce_bucket<VEC>(plan, bq[q] + 1, cq[q] - 1, dout, dout_stride, col, colok, lane,
                                                 wb ? dbias : nullptr, gbx));
            if (lane == r + q) gb_mine += gbx;
          }
        }
        }
#pragma unroll
        for (int q = 0; q < kRowsPerStep; ++q) {
          if (cq[q] == 0 || !colok) continue;
          const VT gq = V<VEC>::mul(g[q], gs);
          const size_t off = (size_t)tq[q] * dim + (size_t)col * VEC;
          if (opt == ARX_OPT_ADAGRAD) {
            V<VEC>::adagrad(Ev[q], Av[q], gq, lr);
            V<VEC>::st(s_attrs[fq[q]].table_acc + off, Av[q]);
            V<VEC>::st(s_attrs[fq[q]].table + off, Ev[q]);
          } else if (opt == ARX_OPT_SGD) {
            V<VEC>::sgd(Ev[q], gq, lr);
            V<VEC>::st(s_attrs[fq[q]].table + off, Ev[q]);
          } else {
            const size_t urow = strided ? (size_t)(r + q) * n_iter + (size_t)it : (size_t)it * 32 + r + q;   // row of lane r + q
            V<VEC>::st(rows_out + urow * dim + (size_t)col * VEC, gq);
          }
        }
      }
    }
    if (light) {
      if (row_bias) bias_update(s_attrs[f], tok, gb_mine * gs, lr, opt);
      if (opt == ARX_OPT_NONE && bias_rows_out != nullptr) bias_rows_out[u] = row_bias ? gb_mine * gs : 0.f;
    }
  }
}

#ifndef ARX_APPLY_MINB
#define ARX_APPLY_MINB 2
#endif
template <int VEC, bool FLAT>
__global__ void __launch_bounds__(256, ARX_APPLY_MINB)
pool_bwd_apply_kernel(const arx_attr_desc* __restrict__ g_attrs, int n_attr, int dim,
                       arx_bwd_plan plan, const float* __restrict__ dout, long long dout_stride,
                       const float* __restrict__ dbias, float lr,
                       const float* __restrict__ grad_scale_dev, int opt, float* __restrict__ rows_out,
                       float* __restrict__ bias_rows_out, int heavy) {
  __shared__ arx_attr_desc s_attrs[kMaxAttr];
  stage_descs(s_attrs, g_attrs, n_attr);
  pool_bwd_apply_body<VEC, FLAT>(s_attrs, dim, plan, dout, dout_stride, dbias, lr, grad_scale_dev, opt, rows_out, bias_rows_out, heavy);
}

// ---- column-slab apply for the catalog gradient (loss = ce / full-catalog WMRB) ---------------------------------
// The gradient of the pooled catalog, dP [V, d], reaches every table row through ~100 contributions (10^6 items x 97
// rows each at C2: 97 M gathers of 512 B over a 512 MB arena, 4x the L2) — the row-at-a-time kernel above is bound by
// those DRAM gathers.  Adagrad is element-wise, so the update can be done one COLUMN SLAB at a time: the caller lays
// dP out slab-major ([d / 16][V][16], 64 MB per slab at V = 10^6: L2-resident), and pass s updates columns
// [16 s, 16 s + 16) of every touched row from gathers that hit L2; the table / accumulator rows are touched 64 B at a
// time, once per pass.  Lanes: 8 entry slots x 4 float4 columns — one warp instruction gathers 8 contributions of 64 B.
// Rows above `heavy` contributions are split into chunks that different warps reduce (the plan's chunk list, built by
// arx_bwd_plan_alloc_h with the same `heavy`); the last chunk to arrive folds the partials in chunk order and updates.
__device__ __forceinline__ float4 slab_reduce_slots(float4 a) {
#pragma unroll
  for (int o = 4; o < 32; o <<= 1) {
    a.x += __shfl_xor_sync(ARX_FULL_MASK, a.x, o); a.y += __shfl_xor_sync(ARX_FULL_MASK, a.y, o);
    a.z += __shfl_xor_sync(ARX_FULL_MASK, a.z, o); a.w += __shfl_xor_sync(ARX_FULL_MASK, a.w, o);
  }
  return a;
}

// sum over bucket entries [base, base + cnt) of w * dslab[src][4 cl .. 4 cl + 4), reduced over the 8 entry slots (every
// lane of column cl returns the total); gb = sum of w * dbias[src] (warp total) when dbias != NULL
__device__ __forceinline__ float4 slab_bucket(const arx_bwd_plan& plan, int base, int cnt, const float* __restrict__ dslab,
                                              int lane, const float* __restrict__ dbias, float& gb) {
  const int slot = lane >> 2, cl = lane & 3;
  float4 acc = f4_zero();
  float gbl = 0.f;
  for (int e0 = 0; e0 < cnt; e0 += 32) {
    int s_l = -1; float w_l = 0.f;
    if (e0 + lane < cnt) {
      // the bucket arrays are streamed once per pass (0.8 GB at C2): evict-first, so that they do not push the slab
      // (the only data with reuse) out of L2
      s_l = __ldcs(plan.bucket_src + base + e0 + lane); w_l = __ldcs(plan.bucket_w + base + e0 + lane);
      if (dbias != nullptr) gbl = fmaf(w_l, __ldg(dbias + s_l), gbl);
    }
    float4 v[4]; float wj[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int sj = __shfl_sync(ARX_FULL_MASK, s_l, j * 8 + slot);
      wj[j] = __shfl_sync(ARX_FULL_MASK, w_l, j * 8 + slot);
      v[j] = sj >= 0 ? ldg_f4(dslab + (size_t)sj * 16 + cl * 4) : f4_zero();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) f4_fma(acc, wj[j], v[j]);
  }
  gb = dbias != nullptr ? warp_sum(gbl) : 0.f;
  return slab_reduce_slots(acc);
}

// table / accumulator slabs are touched once per pass: streaming (evict-first) loads and stores
__device__ __forceinline__ float4 ldcs_f4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void stcs_f4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

__device__ __forceinline__ void slab_row_update(const arx_attr_desc& a, size_t off, float4 g, float lr, int opt) {
  float* wp = a.table + off;
  float4 wv = ldcs_f4(wp);
  if (opt == ARX_OPT_ADAGRAD) {
    float* ap = a.table_acc + off;
    float4 av = ldcs_f4(ap);
    V<4>::adagrad(wv, av, g, lr);
    stcs_f4(ap, av);
  } else {
    V<4>::sgd(wv, g, lr);
  }
  stcs_f4(wp, wv);
}

__global__ void __launch_bounds__(256, 4)
pool_bwd_apply_slab_kernel(const arx_attr_desc* __restrict__ g_attrs, int n_attr, int dim, arx_bwd_plan plan,
                           const float* __restrict__ dslab, int c0, const float* __restrict__ dbias, float lr,
                           const float* __restrict__ grad_scale_dev, int opt, int heavy) {
  __shared__ arx_attr_desc s_attrs[kMaxAttr];
  stage_descs(s_attrs, g_attrs, n_attr);
  if (plan.counters[2] != 0) return;
  const int lane = threadIdx.x & 31, slot = lane >> 2, cl = lane & 3;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nu = (int)min((long long)plan.counters[0], (long long)plan.cap_rows);
  const int nchunks = (int)min((long long)plan.counters[4], (long long)plan.cap_chunks);
  const float gs = grad_scale_dev ? __ldg(grad_scale_dev) : 1.0f;
  const float* __restrict__ db = (c0 == 0) ? dbias : nullptr;           // the bias column rides with the first slab
  float* __restrict__ part = plan.partials;                             // [cap_chunks][16]
  float* __restrict__ part_b = plan.partials + (size_t)plan.cap_chunks * dim;

  // ---- hot rows: one chunk per warp iteration -------------------------------------------------------------------
  for (long long ch = warp0; ch < nchunks; ch += nwarps) {
    const int u = plan.chunk_row[ch];
    const int cb = plan.row_chunk0[u];
    const int k = (int)ch - cb;
    const int rc = plan.row_cnt[u];
    const int nch = (rc + heavy - 1) / heavy;
    float gb;
    const float4 g = slab_bucket(plan, plan.row_base[u] + k * heavy, min(heavy, rc - k * heavy), dslab, lane, db, gb);
    if (slot == 0) st_f4(part + (size_t)ch * 16 + cl * 4, g);
    if (lane == 0 && db != nullptr) part_b[ch] = gb;
    __threadfence();
    int last = 0;
    if (lane == 0) last = (atomicAdd(&plan.row_done[u], 1) == nch - 1) ? 1 : 0;
    last = __shfl_sync(ARX_FULL_MASK, last, 0);
    if (!last) continue;
    __threadfence();
    const int f = plan.uniq_attr[u];
    const int tok = local_row(shard_of(s_attrs[f]), plan.uniq_tok[u]);
    float4 tot = f4_zero();
    float tb = 0.f;
    for (int k0 = 0; k0 < nch; k0 += 8) {                 // chunk k0 + slot: fixed order per slot, then a fixed tree
      if (k0 + slot < nch) f4_add(tot, V<4>::ldcg(part + (size_t)(cb + k0 + slot) * 16 + cl * 4));
    }
    tot = slab_reduce_slots(tot);
    if (db != nullptr && s_attrs[f].bias != nullptr) {
      for (int kk = lane; kk < nch; kk += 32) tb += __ldcg(part_b + cb + kk);
      tb = warp_sum(tb);
    }
    if (slot == 0) slab_row_update(s_attrs[f], (size_t)tok * dim + c0 + cl * 4, f4_scale(tot, gs), lr, opt);
    if (lane == 0) {
      if (db != nullptr && s_attrs[f].bias != nullptr) bias_update(s_attrs[f], tok, tb * gs, lr, opt);
      plan.row_done[u] = 0;
    }
  }

  // ---- all other rows: 32 per warp iteration (strided over the unique list), one after the other --------------
  const int n_iter = (nu + 31) >> 5;
  for (long long it = warp0; it < n_iter; it += nwarps) {
    const int u = lane * n_iter + (int)it;
    int tok = 0, f = 0, base = 0, cnt = 0;
    if (u < nu) {
      f = __ldg(plan.uniq_attr + u);
      tok = local_row(shard_of(s_attrs[f]), __ldg(plan.uniq_tok + u));
      base = __ldg(plan.row_base + u); cnt = __ldg(plan.row_cnt + u);
    }
    const unsigned lightmask = __ballot_sync(ARX_FULL_MASK, (u < nu) && (cnt <= heavy));
    for (int r = 0; r < 32; ++r) {
      if (!((lightmask >> r) & 1u)) continue;
      const int fr = __shfl_sync(ARX_FULL_MASK, f, r), tr = __shfl_sync(ARX_FULL_MASK, tok, r);
      const int br = __shfl_sync(ARX_FULL_MASK, base, r), cr = __shfl_sync(ARX_FULL_MASK, cnt, r);
      // the row's table / accumulator slab is requested before the bucket walk (independent of it)
      const size_t off = (size_t)tr * dim + c0 + cl * 4;
      float4 wv = f4_zero(), av = f4_zero();
      if (slot == 0) {
        wv = ldcs_f4(s_attrs[fr].table + off);
        if (opt == ARX_OPT_ADAGRAD) av = ldcs_f4(s_attrs[fr].table_acc + off);
      }
      const bool hb = db != nullptr && s_attrs[fr].bias != nullptr;
      float gb;
      float4 g = slab_bucket(plan, br, cr, dslab, lane, hb ? db : nullptr, gb);
      g = f4_scale(g, gs);
      if (slot == 0) {
        if (opt == ARX_OPT_ADAGRAD) { V<4>::adagrad(wv, av, g, lr); stcs_f4(s_attrs[fr].table_acc + off, av); }
        else V<4>::sgd(wv, g, lr);
        stcs_f4(s_attrs[fr].table + off, wv);
      }
      if (lane == 0 && hb) bias_update(s_attrs[fr], tr, gb * gs, lr, opt);
    }
  }
}

// Several table sets (the user and the item tables of one training step) in ONE launch: a warp that has run out of rows
// of the first set goes straight on to the second — one launch ramp and one tail instead of two.
struct ApplySet {
  const arx_attr_desc* attrs;
  const float* dout;
  const float* dbias;
  long long dout_stride;
  arx_bwd_plan plan;
  int n_attr;
};
constexpr int kApplySets = 2;
struct ApplyManyParams {
  ApplySet set[kApplySets];
  const float* grad_scale;
  float lr;
  int n_sets, dim, opt, heavy;
};

__global__ void __launch_bounds__(256, 2)
pool_bwd_apply_many_kernel(const ApplyManyParams mp) {
  __shared__ arx_attr_desc s_attrs_all[kApplySets][kMaxAttr];
#pragma unroll
  for (int si = 0; si < kApplySets; ++si) {
    if (si >= mp.n_sets) break;
    const int words = mp.set[si].n_attr * (int)(sizeof(arx_attr_desc) / 8);
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(mp.set[si].attrs);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(s_attrs_all[si]);
    for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  // constant indices into the parameter block: a run-time index would make the compiler copy the parameters to local
  // memory and read every plan pointer through it inside the inner loops
  pool_bwd_apply_body<4, false>(s_attrs_all[0], mp.dim, mp.set[0].plan, mp.set[0].dout, mp.set[0].dout_stride, mp.set[0].dbias,
                         mp.lr, mp.grad_scale, mp.opt, nullptr, nullptr, mp.heavy);
  if (mp.n_sets > 1)
    pool_bwd_apply_body<4, false>(s_attrs_all[1], mp.dim, mp.set[1].plan, mp.set[1].dout, mp.set[1].dout_stride, mp.set[1].dbias,
                           mp.lr, mp.grad_scale, mp.opt, nullptr, nullptr, mp.heavy);
}

// IndexedSlices part of the global norm: sum over occurrences of w^2 * ||dOut[src]||^2
// (tf.global_norm takes the norm of IndexedSlices.values: duplicates are NOT summed first).
// Flat over the occurrence list (bucket arrays are dense in [0, counters[1])): one warp per
// occurrence, four occurrences in flight; no per-row serial walk (hot rows hold 10^4+ entries).
__global__ void __launch_bounds__(256)
pool_bwd_sumsq_kernel(int dim, arx_bwd_plan plan, const float* __restrict__ dout,
                      long long dout_stride, const float* __restrict__ dbias, float* __restrict__ sumsq) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  if (plan.counters[2] != 0) return;
  const long long nocc = min((long long)plan.counters[1], (long long)plan.cap_occ);
  float part = 0.f;
  for (long long i0 = warp0 * 32; i0 < nocc; i0 += nwarps * 32) {
    const long long i = i0 + lane;
    int src = 0; float w = 0.f;
    if (i < nocc) { src = __ldg(plan.bucket_src + i); w = __ldg(plan.bucket_w + i); }
    if (dbias != nullptr && i < nocc) { const float b = w * __ldg(dbias + src); part = fmaf(b, b, part); }
    const int cnt = (int)min(32ll, nocc - i0);
    for (int k0 = 0; k0 < cnt; k0 += 4) {
      float v[4]; float wk[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int sk = __shfl_sync(ARX_FULL_MASK, src, (k0 + q) & 31);
        wk[q] = (k0 + q < cnt) ? __shfl_sync(ARX_FULL_MASK, w, (k0 + q) & 31) : 0.f;
        float s = 0.f;
        for (int c = lane; c < dim; c += 32) { const float x = __ldg(dout + (size_t)sk * dout_stride + c); s = fmaf(x, x, s); }
        v[q] = s;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) part = fmaf(wk[q] * wk[q], v[q], part);
    }
  }
  part = warp_sum(part);
  if (lane == 0 && part != 0.f) atomicAdd(sumsq, part);
}

// Dense-gradient part of the global norm: ||merged row gradients||^2 over the rows arx_pool_bwd_apply
// wrote with ARX_OPT_NONE (duplicates summed BEFORE the norm), plus the merged bias gradients.
__global__ void __launch_bounds__(256)
rows_sumsq_kernel(const float* __restrict__ rows, const float* __restrict__ brows, arx_bwd_plan plan, int dim,
                  float* __restrict__ sumsq) {
  const long long nu = min((long long)plan.counters[0], (long long)plan.cap_rows);
  const long long n = nu * dim;
  float part = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = __ldg(rows + i);
    part = fmaf(x, x, part);
  }
  if (brows != nullptr)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nu; i += (long long)gridDim.x * blockDim.x) {
      const float x = __ldg(brows + i);
      part = fmaf(x, x, part);
    }
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0 && part != 0.f) atomicAdd(sumsq, part);
}

int g_tune_flat_epb = 0;         // arx_set_tuning("flat_epb", 0 = auto | 1..16): entities per CTA of the flat forward
int g_tune_apply_cps = 4;        // arx_set_tuning("apply_ctas_per_sm", 1..4)
int g_tune_heavy = 64;          // arx_set_tuning("heavy", 8..64): bucket size above which a row is split across warps; must not change
                                // between building a plan and applying it; arx_bwd_plan.cap_chunks >= 2 * cap_occ / heavy + 1
int g_tune_apply_phase_off = 0; // measurement only: bit 0 skips the hot-row chunk phase, bit 1 the row phase (results are then WRONG)
int g_tune_apply_contig = 0;    // arx_set_tuning("apply_contig", 1): A/B switch back to 32 CONSECUTIVE unique rows per warp iteration
int g_tune_apply_flat = 0;      // arx_set_tuning("apply_flat", 0 | 1): pre-loaded flat bucket entries in the row phase of the apply kernel
int g_tune_plan_agg = 3;         // arx_set_tuning("plan_agg", 0 | 1): block-aggregated plan_count / plan_fill (tables < 2^27 rows)

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
               float* out, int64_t out_stride, int mode, float* bias_out, int max_rows, cudaStream_t st) {
  const int nvec = dim / VEC;
  const int threads = 256;
  if (VEC == 4 && mode == ARX_POOL_MEAN && (dim == 128 || dim == 256) && max_rows > 0 && max_rows <= kFlatRows &&
      n_attr <= kFlatBags) {
    // flat row-list kernel (see pool_fwd_flat_kernel): groups of up to kFlatBags bags / kFlatEnt entities /
    // kFlatRows rows per CTA pass; one balanced wave when the batch allows it
    const int cps = (dim == 128) ? 4 : 2;                                        // resident CTAs per SM
    const long long slots = (long long)arx_num_sms() * cps;
    int epb = std::min(kFlatBags / n_attr, kFlatEnt);
    const int even = (int)((n + slots - 1) / slots);                             // entities per CTA for exactly one wave
    if (even >= 1 && even < epb) epb = even;
    if (g_tune_flat_epb > 0 && g_tune_flat_epb < epb) epb = g_tune_flat_epb;
    const size_t smem = (size_t)(kFlatEnt + 16) * dim * 4 + (size_t)kFlatRows * (8 + 8) + (size_t)n_attr * sizeof(arx_attr_desc);
    long long blocks = (n + epb - 1) / epb;
    const long long cap = slots * 4;
    const int grid = (int)(blocks > cap ? cap : blocks);
    // the dynamic size grows with the attribute count: raise the opt-in limit whenever a larger request comes
    // (configuring once with the FIRST request's size made later, wider table sets fail to launch)
    static size_t cfg1 = 0, cfg2 = 0;
    if (dim == 128) {
      if (smem > cfg1) { cudaFuncSetAttribute(pool_fwd_flat_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); cfg1 = smem; }
      pool_fwd_flat_kernel<1><<<grid, threads, smem, st>>>(attrs, n_attr, ids, (long long)n, out, (long long)out_stride, bias_out, epb);
    } else {
      if (smem > cfg2) { cudaFuncSetAttribute(pool_fwd_flat_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); cfg2 = smem; }
      pool_fwd_flat_kernel<2><<<grid, threads, smem, st>>>(attrs, n_attr, ids, (long long)n, out, (long long)out_stride, bias_out, epb);
    }
    ARX_CHECK_LAUNCH();
    return ARX_OK;
  }
  // entities per CTA: 4 when that still gives >= 2 CTAs per SM, fewer for small batches;
  // bounded so the staging buffer stays under the 48 KB static-opt-in-free limit.
  int epb = 4;
  while (epb > 1 && (n + epb - 1) / epb < 2LL * arx_num_sms()) epb >>= 1;
  while (epb > 1 && (size_t)epb * n_attr * (dim + 1) * sizeof(float) > 40 * 1024) epb >>= 1;
  const size_t smem = (size_t)epb * n_attr * (dim + 1) * sizeof(float);
  if (smem > 48 * 1024) return ARX_E_UNSUPPORTED;
  long long blocks = (n + epb - 1) / epb;
  const long long cap = (long long)arx_num_sms() * 32;
  const int grid = (int)(blocks > cap ? cap : blocks);
#define ARX_FWD(GW)                                                                                  \
  pool_fwd_kernel<GW, VEC><<<grid, threads, smem, st>>>(attrs, n_attr, dim, ids, (long long)n, out, \
                                                         (long long)out_stride, mode, bias_out, epb)
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
                            int max_rows_per_entity, void* stream) {
  if (!attrs || !ent_ids || !out || n_attr < 1 || n_attr > kMaxAttr || dim < 1 || n < 0) return ARX_E_BADARG;
  if (mode != ARX_POOL_MEAN && mode != ARX_POOL_CONCAT) return ARX_E_BADARG;
  if (n == 0) return ARX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool v4 = (dim % 4 == 0) && (out_stride % 4 == 0) && (((uintptr_t)out & 15) == 0);
  return v4 ? launch_fwd<4>(attrs, n_attr, dim, ent_ids, n, out, out_stride, mode, bias_out, max_rows_per_entity, st)
            : launch_fwd<1>(attrs, n_attr, dim, ent_ids, n, out, out_stride, mode, bias_out, 0, st);
}

// K1+K2 for several independent lookups in one launch (mean mode, dim 128 / 256, every request within the flat
// kernel's limits); ARX_E_UNSUPPORTED otherwise: the caller then issues one arx_pool_fwd per lookup.
static int pool_fwd_many_impl(const arx_pool_req* reqs, const arx_pool_push* push, int n_req, int dim, void* stream) {
  if (!reqs || n_req < 1 || n_req > kManyReqs) return ARX_E_BADARG;
  if (dim != 128 && dim != 256) return ARX_E_UNSUPPORTED;
  PoolManyParams mp{};
  const int cps = (dim == 128) ? 4 : 2;
  const long long slots = (long long)arx_num_sms() * cps;
  long long total_n = 0;
  int max_attr = 0, k = 0;
  for (int i = 0; i < n_req; ++i) {
    const arx_pool_req& q = reqs[i];
    const bool pushed = push != nullptr && push[i].peer_out != nullptr;
    if (!q.attrs || !q.ent_ids || (!q.out && !pushed) || q.n_attr < 1 || q.n_attr > kMaxAttr || q.n < 0) return ARX_E_BADARG;
    if (pushed && (push[i].rows_per_rank < 0 || push[i].stride < dim || (push[i].stride % 4) || push[i].n_ranks < 1)) return ARX_E_BADARG;
    if (q.n_attr > kFlatBags || q.max_rows_per_entity <= 0 || q.max_rows_per_entity > kFlatRows ||
        (!pushed && ((q.out_stride % 4) || ((uintptr_t)q.out & 15))))
      return ARX_E_UNSUPPORTED;
    total_n += q.n;
    max_attr = std::max(max_attr, q.n_attr);
  }
  if (total_n == 0) return ARX_OK;
  int blocks = 0;
  for (int i = 0; i < n_req; ++i) {
    const arx_pool_req& q = reqs[i];
    if (q.n == 0) continue;
    int epb = std::min(kFlatBags / q.n_attr, kFlatEnt);
    // entities per CTA: as many as fit, but not more than what spreads the whole batch of lookups over one wave
    const int even = (int)((total_n + slots - 1) / slots);
    if (even >= 1 && even < epb) epb = even;
    if (g_tune_flat_epb > 0 && g_tune_flat_epb < epb) epb = g_tune_flat_epb;
    mp.attrs[k] = q.attrs; mp.ids[k] = q.ent_ids; mp.out[k] = q.out; mp.bias_out[k] = q.bias_out;
    mp.n[k] = q.n; mp.out_stride[k] = q.out_stride; mp.n_attr[k] = q.n_attr; mp.epb[k] = epb;
    mp.push[k] = PushDesc{};
    if (push != nullptr && push[i].peer_out != nullptr) {
      mp.push[k].peer = push[i].peer_out; mp.push[k].peer_bias = push[i].peer_bias;
      mp.push[k].rows_per_rank = push[i].rows_per_rank; mp.push[k].stride = push[i].stride; mp.push[k].G = push[i].n_ranks;
      // in push mode the pooled bias goes to the peers' bias vectors; bias_out only says "bias wanted"
      mp.bias_out[k] = push[i].peer_bias ? reinterpret_cast<float*>(16) : nullptr;
    }
    blocks += (int)std::min<long long>((q.n + epb - 1) / epb, slots * 4);
    mp.block_end[k] = blocks;
    ++k;
  }
  mp.n_req = k;
  const size_t smem = (size_t)(kFlatEnt + 16) * dim * 4 + (size_t)kFlatRows * (8 + 8) + (size_t)max_attr * sizeof(arx_attr_desc);
  cudaStream_t st = (cudaStream_t)stream;
  static size_t cfg1 = 0, cfg2 = 0;
  if (dim == 128) {
    if (smem > cfg1) { cudaFuncSetAttribute(pool_fwd_flat_many_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); cfg1 = smem; }
    pool_fwd_flat_many_kernel<1><<<blocks, 256, smem, st>>>(mp);
  } else {
    if (smem > cfg2) { cudaFuncSetAttribute(pool_fwd_flat_many_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); cfg2 = smem; }
    pool_fwd_flat_many_kernel<2><<<blocks, 256, smem, st>>>(mp);
  }
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_pool_fwd_many(const arx_pool_req* reqs, int n_req, int dim, void* stream) {
  return pool_fwd_many_impl(reqs, nullptr, n_req, dim, stream);
}

// The same launch with the results of the requests whose push[i].peer_out is set ADDED into the owner ranks' receive
// buffers over NVLink peer memory (see PushDesc) instead of stored locally: lookup + reduce-scatter in one kernel.
extern "C" int arx_pool_fwd_many_push(const arx_pool_req* reqs, const arx_pool_push* push, int n_req, int dim, void* stream) {
  if (!push) return ARX_E_BADARG;
  return pool_fwd_many_impl(reqs, push, n_req, dim, stream);
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
         plan.bucket_src && plan.bucket_w && plan.chunk_row && plan.row_chunk0 && plan.row_done &&
         plan.partials && plan.cap_rows >= 1 && plan.cap_occ >= 1 && plan.cap_chunks >= 1;
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
  if (g_tune_plan_agg & 1) {
    const int epb = std::max(1, kAggBags / n_attr);
    const long long blocks = (n + epb - 1) / epb;
    const int grid = (int)std::min(blocks, (long long)arx_num_sms() * 16);
    plan_count_agg_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(attrs, attr_begin, n_attr, ent_ids, (long long)n, plan, epb);
  } else
  plan_count_kernel<<<pick_grid(n * n_attr, 256), 256, 0, (cudaStream_t)stream>>>(attrs, attr_begin, n_attr, ent_ids,
                                                                        (long long)n, plan);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_bwd_plan_alloc(const arx_attr_desc* attrs, arx_bwd_plan plan, void* stream) {
  if (!attrs || !plan_args_ok(plan)) return ARX_E_BADARG;
  plan_alloc_kernel<<<arx_num_sms() * 4, 256, 0, (cudaStream_t)stream>>>(attrs, plan, g_tune_heavy);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

// arx_bwd_plan_alloc with an explicit chunk size (the plan and the kernel that applies it must agree on it)
extern "C" int arx_bwd_plan_alloc_h(const arx_attr_desc* attrs, arx_bwd_plan plan, int heavy, void* stream) {
  if (!attrs || !plan_args_ok(plan) || heavy < 8 || heavy > 4096) return ARX_E_BADARG;
  plan_alloc_kernel<<<arx_num_sms() * 4, 256, 0, (cudaStream_t)stream>>>(attrs, plan, heavy);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

// One column slab of the de-duplicated optimizer step: dslab [rows of the gradient arena][16] = columns [c0, c0 + 16)
// of dOut (slab-major copy made by the caller), dim % 16 == 0, ARX_OPT_ADAGRAD / ARX_OPT_SGD.  dbias is applied by
// the pass with c0 == 0.  heavy: the chunk size the plan was allocated with (arx_bwd_plan_alloc_h).
extern "C" int arx_pool_bwd_apply_slab(const arx_attr_desc* attrs, int n_attr, int dim, arx_bwd_plan plan, const float* dslab,
                                       int c0, const float* dbias, float lr, const float* grad_scale_dev, int opt, int heavy,
                                       void* stream) {
  if (!attrs || !dslab || n_attr < 1 || n_attr > kMaxAttr || dim < 16 || (dim % 16) || c0 < 0 || c0 + 16 > dim || (c0 % 16) ||
      !plan_args_ok(plan) || heavy < 8 || heavy > 4096 || (((uintptr_t)dslab) & 15))
    return ARX_E_BADARG;
  if (opt != ARX_OPT_ADAGRAD && opt != ARX_OPT_SGD) return ARX_E_UNSUPPORTED;
  pool_bwd_apply_slab_kernel<<<arx_num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(attrs, n_attr, dim, plan, dslab, c0, dbias, lr,
                                                                                grad_scale_dev, opt, heavy);
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
  if (g_tune_plan_agg & 2) {
    const int epb = std::max(1, kAggBags / n_attr);
    const long long blocks = (n + epb - 1) / epb;
    const int grid = (int)std::min(blocks, (long long)arx_num_sms() * 16);
    plan_fill_agg_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(attrs, attr_begin, n_attr, ent_ids, (long long)n, mode,
                                                                (long long)row_base, plan, epb);
  } else
  plan_fill_kernel<<<pick_grid(n * n_attr, 256), 256, 0, (cudaStream_t)stream>>>(attrs, attr_begin, n_attr, ent_ids,
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
  const int grid = arx_num_sms() * g_tune_apply_cps;   // persistent (2 CTAs/SM): warps stride over the device-side row list
  const bool v4 = (dim % 4 == 0) && (dout_stride % 4 == 0) && (((uintptr_t)dout & 15) == 0);
  if (v4 && g_tune_apply_flat)
    pool_bwd_apply_kernel<4, true><<<grid, 256, 0, st>>>(attrs, n_attr, dim, plan, dout, (long long)dout_stride,
                                                          dbias, lr, grad_scale_dev, opt, rows_out, bias_rows_out, g_tune_heavy | (g_tune_apply_phase_off << 8) | (g_tune_apply_contig << 10));
  else if (v4)
    pool_bwd_apply_kernel<4, false><<<grid, 256, 0, st>>>(attrs, n_attr, dim, plan, dout, (long long)dout_stride,
                                                           dbias, lr, grad_scale_dev, opt, rows_out, bias_rows_out, g_tune_heavy | (g_tune_apply_phase_off << 8) | (g_tune_apply_contig << 10));
  else
    pool_bwd_apply_kernel<1, false><<<grid, 256, 0, st>>>(attrs, n_attr, dim, plan, dout, (long long)dout_stride,
                                                           dbias, lr, grad_scale_dev, opt, rows_out, bias_rows_out, g_tune_heavy | (g_tune_apply_phase_off << 8) | (g_tune_apply_contig << 10));
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

// De-duplicated optimizer step of up to two table sets (e.g. the user and the item tables of one training step) in one
// launch; dim % 4 == 0, ARX_OPT_ADAGRAD / ARX_OPT_SGD.  ARX_E_UNSUPPORTED otherwise (one arx_pool_bwd_apply per set then).
extern "C" int arx_pool_bwd_apply_many(const arx_apply_set* sets, int n_sets, int dim, float lr,
                                       const float* grad_scale_dev, int opt, void* stream) {
  if (!sets || n_sets < 1 || n_sets > kApplySets || dim < 1) return ARX_E_BADARG;
  if (opt != ARX_OPT_ADAGRAD && opt != ARX_OPT_SGD) return ARX_E_UNSUPPORTED;
  if (dim % 4) return ARX_E_UNSUPPORTED;
  ApplyManyParams mp{};
  for (int i = 0; i < n_sets; ++i) {
    const arx_apply_set& q = sets[i];
    if (!q.attrs || !q.dout || q.n_attr < 1 || q.n_attr > kMaxAttr || !plan_args_ok(q.plan)) return ARX_E_BADARG;
    if ((q.dout_stride % 4) || ((uintptr_t)q.dout & 15)) return ARX_E_UNSUPPORTED;
    mp.set[i].attrs = q.attrs; mp.set[i].dout = q.dout; mp.set[i].dbias = q.dbias;
    mp.set[i].dout_stride = q.dout_stride; mp.set[i].plan = q.plan; mp.set[i].n_attr = q.n_attr;
  }
  mp.grad_scale = grad_scale_dev; mp.lr = lr; mp.n_sets = n_sets; mp.dim = dim; mp.opt = opt; mp.heavy = g_tune_heavy;
  const int grid = arx_num_sms() * g_tune_apply_cps;   // persistent: warps stride over the device-side row lists
  pool_bwd_apply_many_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mp);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

// Launch-shape knobs for measurement sweeps (tools/bench_pool.py); defaults are the tuned values.
extern "C" int arx_set_tuning(const char* key, int value) {
  if (!key) return ARX_E_BADARG;
  const auto eq = [&](const char* k) { int i = 0; while (k[i] && key[i] == k[i]) ++i; return k[i] == 0 && key[i] == 0; };
  if (eq("flat_epb")) {
    if (value < 0 || value > kFlatEnt) return ARX_E_BADARG;
    g_tune_flat_epb = value;
    return ARX_OK;
  }
  if (eq("plan_agg")) {
    g_tune_plan_agg = value & 3;          // bit 0: count, bit 1: fill
    return ARX_OK;
  }
  if (eq("heavy")) {
    if (value < 8 || value > kHeavyMax) return ARX_E_BADARG;
    g_tune_heavy = value;
    return ARX_OK;
  }
  if (eq("apply_phase_off")) {
    g_tune_apply_phase_off = value & 3;
    return ARX_OK;
  }
  if (eq("apply_contig")) {
    g_tune_apply_contig = value & 1;
    return ARX_OK;
  }
  if (eq("apply_flat")) {
    g_tune_apply_flat = value & 1;
    return ARX_OK;
  }
  if (eq("apply_ctas_per_sm")) {
    if (value < 1 || value > 8) return ARX_E_BADARG;
    g_tune_apply_cps = value;
    return ARX_OK;
  }
  return ARX_E_BADARG;
}

extern "C" int arx_pool_bwd_sumsq(const arx_attr_desc* attrs, int dim, arx_bwd_plan plan, const float* dout,
                                  int64_t dout_stride, const float* dbias, float* sumsq, int merged,
                                  void* stream) {
  if (!dout || !sumsq || dim < 1 || !plan_args_ok(plan)) return ARX_E_BADARG;
  if (merged) return ARX_E_UNSUPPORTED;     // dense semantics: arx_pool_bwd_apply(ARX_OPT_NONE) + arx_rows_sumsq
  (void)attrs;
  pool_bwd_sumsq_kernel<<<arx_num_sms() * 4, 256, 0, (cudaStream_t)stream>>>(
      dim, plan, dout, (long long)dout_stride, dbias, sumsq);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_rows_sumsq(const float* rows, const float* bias_rows, arx_bwd_plan plan, int dim, float* sumsq,
                              void* stream) {
  if (!rows || !sumsq || dim < 1 || !plan_args_ok(plan)) return ARX_E_BADARG;
  rows_sumsq_kernel<<<arx_num_sms() * 4, 256, 0, (cudaStream_t)stream>>>(rows, bias_rows, plan, dim, sumsq);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}
