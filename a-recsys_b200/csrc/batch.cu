// Device-side batch assembly and negative-pool sampling: what sits immediately upstream of the
// embedding lookups.  In the reference these are Python loops per example (hmf/hmf_model.py:230-260,
// word2vec/data_iterator.py:108-169, lstm/seqModel.py:356-404, utils/prepare_train.py:7-17); a 0.5 ms GPU step cannot be
// fed by 4096 interpreter-level draws.  All kernels are integer gathers: bit-exact against the host functions for the
// same selected indices.
#include "arx_common.cuh"
#include "philox.cuh"

namespace {

inline int grid_for(long long n, int threads = 256) {
  long long b = (n + threads - 1) / threads;
  const long long cap = (long long)arx_num_sms() * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// out_u[b] = users[idx[b]], out_i[b] = items[idx[b]]     (LatentProductModel.get_batch / get_permuted_batch)
__global__ void gather_pairs_kernel(const int* __restrict__ users, const int* __restrict__ items,
                                    const long long* __restrict__ idx, long long n, int* __restrict__ out_u,
                                    int* __restrict__ out_i) {
  for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < n; b += (long long)gridDim.x * blockDim.x) {
    const long long j = idx[b];
    out_u[b] = __ldg(users + j);
    out_i[b] = __ldg(items + j);
  }
}

// CBOW window batch (word2vec/data_iterator.py:108-169): slot b takes the (cursor + b)-th non-PAD stream event as its
// target and draws `ni` of the `window` preceding stream positions as inputs — distinct positions when the user already
// has >= ni events in the window, with replacement otherwise.
__global__ void cbow_window_kernel(const int* __restrict__ users, const int* __restrict__ items,
                                   const int* __restrict__ u_seq_len, const long long* __restrict__ targets,
                                   long long l_seq, long long n_targets, long long cursor, int mb, int ni, int window,
                                   const unsigned long long* __restrict__ rng, int* __restrict__ out_users,
                                   int* __restrict__ out_inputs, int* __restrict__ out_targets) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= mb) return;
  const long long sel = targets[(cursor + b) % n_targets];
  out_users[b] = __ldg(users + sel);
  out_targets[b] = __ldg(items + sel);
  const bool with_repl = __ldg(u_seq_len + sel) < ni || ni > window;
  // up to 64 window offsets: partial Fisher-Yates over a bit mask (without replacement) or plain draws
  unsigned long long used = 0ull;
  for (int k = 0; k < ni; ++k) {
    uint32_t o[4];
    philox_words(rng, ((unsigned long long)b << 8) | (unsigned long long)(k >> 2), o);
    const uint32_t w = o[k & 3];
    int off;
    if (with_repl) {
      off = (int)(((unsigned long long)w * (unsigned long long)window) >> 32);
    } else {
      const int remaining = window - k;
      int r = (int)(((unsigned long long)w * (unsigned long long)remaining) >> 32);     // r-th unused offset
      off = 0;
      for (int c = 0; c < window; ++c) {
        if ((used >> c) & 1ull) continue;
        if (r == 0) { off = c; break; }
        --r;
      }
      used |= 1ull << off;
    }
    long long pos = (sel - window + off) % l_seq;
    if (pos < 0) pos += l_seq;
    out_inputs[(long long)k * mb + b] = __ldg(items + pos);
  }
}

// LSTM batch (lstm/seqModel.py:356-404): slot b holds sequence sel[b] (or nothing when sel[b] < 0):
// inputs = [START] + seq[:-1] + pad, targets = seq + pad, weights = 1 on the sequence, 0 on the padding; time-major.
__global__ void lstm_pad_kernel(const long long* __restrict__ seq_ptr, const int* __restrict__ seq_items,
                                const int* __restrict__ seq_users, const long long* __restrict__ sel, int mb, int T,
                                int start_id, int pad_id, int user_pad, int* __restrict__ out_users,
                                int* __restrict__ out_inputs, int* __restrict__ out_targets,
                                float* __restrict__ out_weights) {
  const long long n = (long long)mb * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i / mb), b = (int)(i - (long long)t * mb);
    const long long s = sel[b];
    long long p0 = 0; int len = 0;
    if (s >= 0) { p0 = seq_ptr[s]; len = (int)(seq_ptr[s + 1] - p0); if (len > T) len = T; }
    if (t == 0) out_users[b] = s >= 0 ? __ldg(seq_users + s) : user_pad;
    int in = pad_id;
    if (t == 0) in = start_id;
    else if (t < len) in = __ldg(seq_items + p0 + t - 1);            // [START] + seq[:-1] + pad (an empty slot: START + pad)
    out_inputs[i] = in;
    out_targets[i] = t < len ? __ldg(seq_items + p0 + t) : pad_id;
    out_weights[i] = t < len ? 1.0f : 0.0f;
  }
}

// Gumbel-top-k keys: key[i] = log p[i] - log(-log U_i).  The n largest keys are a draw of n items WITHOUT replacement
// with probabilities p taken sequentially, i.e. np.random.choice(population, n, replace=False, p=p)
// (utils/prepare_train.py:7-17); the selection itself is arx_topk_rows.
__global__ void gumbel_keys_kernel(const float* __restrict__ logp, long long n, const unsigned long long* __restrict__ rng,
                                   float* __restrict__ keys) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint32_t o[4];
    philox_words(rng, (unsigned long long)i >> 2, o);
    const float u = fmaxf(u01(o[i & 3]), 1e-20f);
    keys[i] = __ldg(logp + i) - __logf(-__logf(u));
  }
}

__global__ void rng_tick_kernel(unsigned long long* rng) { rng[1] += 1ull; }

}  // namespace

extern "C" int arx_gather_pairs(const int32_t* users, const int32_t* items, const int64_t* idx, int64_t n,
                                int32_t* out_users, int32_t* out_items, void* stream) {
  if (!users || !items || !idx || !out_users || !out_items || n < 0) return ARX_E_BADARG;
  if (n == 0) return ARX_OK;
  gather_pairs_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(users, items, (const long long*)idx, (long long)n,
                                                                   out_users, out_items);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_cbow_window_batch(const int32_t* users, const int32_t* items, const int32_t* u_seq_len,
                                     const int64_t* targets, int64_t l_seq, int64_t n_targets, int64_t cursor, int mb,
                                     int ni, int window, uint64_t* rng_state, int32_t* out_users, int32_t* out_inputs,
                                     int32_t* out_targets, void* stream) {
  if (!users || !items || !u_seq_len || !targets || !rng_state || !out_users || !out_inputs || !out_targets)
    return ARX_E_BADARG;
  if (l_seq < 1 || n_targets < 1 || mb < 0 || ni < 1 || window < 1) return ARX_E_BADARG;
  if (window > 64 || ni > 64) return ARX_E_UNSUPPORTED;
  if (mb == 0) return ARX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  cbow_window_kernel<<<(mb + 255) / 256, 256, 0, st>>>(users, items, u_seq_len, (const long long*)targets, (long long)l_seq,
                                                       (long long)n_targets, (long long)cursor, mb, ni, window,
                                                       (const unsigned long long*)rng_state, out_users, out_inputs,
                                                       out_targets);
  rng_tick_kernel<<<1, 1, 0, st>>>((unsigned long long*)rng_state);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_lstm_pad_batch(const int64_t* seq_ptr, const int32_t* seq_items, const int32_t* seq_users,
                                  const int64_t* sel, int mb, int T, int start_id, int pad_id, int user_pad_id,
                                  int32_t* out_users, int32_t* out_inputs, int32_t* out_targets, float* out_weights,
                                  void* stream) {
  if (!seq_ptr || !seq_items || !seq_users || !sel || !out_users || !out_inputs || !out_targets || !out_weights ||
      mb < 0 || T < 1)
    return ARX_E_BADARG;
  if (mb == 0) return ARX_OK;
  lstm_pad_kernel<<<grid_for((long long)mb * T), 256, 0, (cudaStream_t)stream>>>(
      (const long long*)seq_ptr, seq_items, seq_users, (const long long*)sel, mb, T, start_id, pad_id, user_pad_id,
      out_users, out_inputs, out_targets, out_weights);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_gumbel_keys(const float* logp, int64_t n, uint64_t* rng_state, float* keys, void* stream) {
  if (!logp || !rng_state || !keys || n < 0) return ARX_E_BADARG;
  if (n == 0) return ARX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  gumbel_keys_kernel<<<grid_for(n), 256, 0, st>>>(logp, (long long)n, (const unsigned long long*)rng_state, keys);
  rng_tick_kernel<<<1, 1, 0, st>>>((unsigned long long*)rng_state);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}
