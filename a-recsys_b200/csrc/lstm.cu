// LSTM cell pointwise stages (K8): the gate GEMMs run on the tensor cores (gemm_tc.cu:
// x-projection for all T steps as one contraction, h_{t-1} * W_h per step accumulated onto it);
// these kernels are the fused gate non-linearities and state update and their adjoint.
//
// Reference semantics: TF-1.0 LSTMCell as used at lstm/seqModel.py:99-103 — gate order
// i, j, f, o on [x, h] W + b; c' = sigmoid(f + forget_bias) c + sigmoid(i) tanh(j);
// h' = sigmoid(o) tanh(c'); no peepholes, projection or clipping.
#include "arx_common.cuh"

namespace {

__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }

// Z [mb, 4H]: pre-activations in, activated gates (i, j, f, o) out (kept for the backward pass)
__device__ __forceinline__ float round_tf32(float x) {      // nearest tf32 (10-bit mantissa), ties away
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__global__ void lstm_gates_fwd_kernel(float* __restrict__ Z, const float* __restrict__ c_prev,
                                      float* __restrict__ c, float* __restrict__ h, long long mb, int H,
                                      float forget_bias, float* __restrict__ h_tf32) {
  const long long n = mb * H;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
    const long long b = idx / H;
    const int k = (int)(idx - b * H);
    float* z = Z + b * 4 * H;
    const float i = sigm(z[k]);
    const float j = tanhf(z[H + k]);
    const float f = sigm(z[2 * H + k] + forget_bias);
    const float o = sigm(z[3 * H + k]);
    const float cp = c_prev ? c_prev[idx] : 0.f;
    const float cn = f * cp + i * j;
    z[k] = i; z[H + k] = j; z[2 * H + k] = f; z[3 * H + k] = o;
    c[idx] = cn;
    const float hn = o * tanhf(cn);
    h[idx] = hn;
    if (h_tf32) h_tf32[idx] = round_tf32(hn);       // the A operand of the next step's h W_h contraction
  }
}

// G [mb, 4H]: activated gates in, dZ (gradient w.r.t. the pre-activations) out.
// dh = dh_out (+ dh_rec); dc = dh o (1 - tanh(c)^2) (+ dc_next); dc_prev = dc f.
__global__ void lstm_gates_bwd_kernel(float* __restrict__ G, const float* __restrict__ c_prev,
                                      const float* __restrict__ c, const float* __restrict__ dh_out,
                                      const float* __restrict__ dh_rec, const float* __restrict__ dc_next,
                                      float* __restrict__ dc_prev, long long mb, int H, int round_out) {
  const long long n = mb * H;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
    const long long b = idx / H;
    const int k = (int)(idx - b * H);
    float* g = G + b * 4 * H;
    const float i = g[k], j = g[H + k], f = g[2 * H + k], o = g[3 * H + k];
    const float dh = (dh_out ? dh_out[idx] : 0.f) + (dh_rec ? dh_rec[idx] : 0.f);
    const float tc = tanhf(c[idx]);
    const float dc = dh * o * (1.0f - tc * tc) + (dc_next ? dc_next[idx] : 0.f);
    const float cp = c_prev ? c_prev[idx] : 0.f;
    float g0 = dc * j * i * (1.0f - i), g1 = dc * i * (1.0f - j * j), g2 = dc * cp * f * (1.0f - f),
          g3 = dh * tc * o * (1.0f - o);
    if (round_out) { g0 = round_tf32(g0); g1 = round_tf32(g1); g2 = round_tf32(g2); g3 = round_tf32(g3); }
    g[k] = g0; g[H + k] = g1; g[2 * H + k] = g2; g[3 * H + k] = g3;   // dZ only feeds tensor-core contractions
    dc_prev[idx] = dc * f;
  }
}

// y = a * x1 + b * x2 (x2 may be NULL): LSTM input mixing mean([user, item], 0) and its adjoint
__global__ void axpby_rows_kernel(const float* __restrict__ x1, const float* __restrict__ x2_rows,
                                  float a, float b, long long rows, long long rep, int dim,
                                  float* __restrict__ y) {
  // y[r, :] = a * x1[r, :] + b * x2_rows[r % rep, :]   (x2 broadcast over the T time steps)
  const long long n = rows * dim;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
    const long long r = idx / dim;
    const int k = (int)(idx - r * dim);
    float v = a * x1[idx];
    if (x2_rows) v += b * x2_rows[(r % rep) * dim + k];
    y[idx] = v;
  }
}

// out[r, :] = scale * sum_t x[t * rep + r, :]   (adjoint of the broadcast above)
__global__ void sum_over_steps_kernel(const float* __restrict__ x, long long T, long long rep, int dim,
                                      float scale, float* __restrict__ out) {
  const long long n = rep * dim;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
    float s = 0.f;
    for (long long t = 0; t < T; ++t) s += x[t * n + idx];
    out[idx] = s * scale;
  }
}

inline int ew_grid(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)arx_num_sms() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" int arx_lstm_gates_fwd2(float* Z, const float* c_prev, float* c, float* h, float* h_tf32, int64_t mb,
                                   int H, float forget_bias, void* stream) {
  if (!Z || !c || !h || mb < 0 || H < 1) return ARX_E_BADARG;
  if (mb == 0) return ARX_OK;
  lstm_gates_fwd_kernel<<<ew_grid(mb * H), 256, 0, (cudaStream_t)stream>>>(Z, c_prev, c, h, mb, H, forget_bias,
                                                                           h_tf32);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_lstm_gates_fwd(float* Z, const float* c_prev, float* c, float* h, int64_t mb, int H,
                                  float forget_bias, void* stream) {
  return arx_lstm_gates_fwd2(Z, c_prev, c, h, nullptr, mb, H, forget_bias, stream);
}

extern "C" int arx_lstm_gates_bwd2(float* G, const float* c_prev, const float* c, const float* dh_out,
                                   const float* dh_rec, const float* dc_next, float* dc_prev, int64_t mb, int H,
                                   int round_tf32_out, void* stream) {
  if (!G || !c || !dc_prev || mb < 0 || H < 1) return ARX_E_BADARG;
  if (mb == 0) return ARX_OK;
  lstm_gates_bwd_kernel<<<ew_grid(mb * H), 256, 0, (cudaStream_t)stream>>>(G, c_prev, c, dh_out, dh_rec, dc_next,
                                                                           dc_prev, mb, H, round_tf32_out);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_lstm_gates_bwd(float* G, const float* c_prev, const float* c, const float* dh_out,
                                  const float* dh_rec, const float* dc_next, float* dc_prev, int64_t mb, int H,
                                  void* stream) {
  return arx_lstm_gates_bwd2(G, c_prev, c, dh_out, dh_rec, dc_next, dc_prev, mb, H, 0, stream);
}

extern "C" int arx_axpby_rows(const float* x1, const float* x2_rows, float a, float b, int64_t rows,
                              int64_t rep, int dim, float* y, void* stream) {
  if (!x1 || !y || rows < 0 || rep < 1 || dim < 1) return ARX_E_BADARG;
  if (rows == 0) return ARX_OK;
  axpby_rows_kernel<<<ew_grid(rows * dim), 256, 0, (cudaStream_t)stream>>>(x1, x2_rows, a, b, rows, rep, dim, y);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_sum_over_steps(const float* x, int64_t T, int64_t rep, int dim, float scale, float* out,
                                  void* stream) {
  if (!x || !out || T < 0 || rep < 0 || dim < 1) return ARX_E_BADARG;
  if (rep == 0) return ARX_OK;
  sum_over_steps_kernel<<<ew_grid(rep * dim), 256, 0, (cudaStream_t)stream>>>(x, T, rep, dim, scale, out);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}
