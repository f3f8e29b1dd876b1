// Dense optimiser step (K11) and small elementwise helpers of the towers.
// Reference: tf.train.AdagradOptimizer at hmf/hmf_model.py:146-151 (acc0 = 0.1, no eps),
// tf.nn.dropout at attributes/embed_attribute.py:236.
#include "arx_common.cuh"

namespace {

__global__ void dense_update_kernel(float* __restrict__ w, float* __restrict__ acc,
                                    const float* __restrict__ g, long long n, float lr,
                                    const float* __restrict__ gs_dev, int opt) {
  const float gs = gs_dev ? __ldg(gs_dev) : 1.0f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i] * gs;
    if (opt == ARX_OPT_ADAGRAD) {
      const float a = fmaf(gi, gi, acc[i]);
      acc[i] = a;
      w[i] -= lr * gi / sqrtf(a);
    } else {
      w[i] -= lr * gi;
    }
  }
}

__global__ void scale_mask_kernel(const float* __restrict__ x, const float* __restrict__ mask,
                                  float scale, long long n, float* __restrict__ y) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = mask ? x[i] * scale * mask[i] : x[i] * scale;
}

inline int ew_grid(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)arx_num_sms() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" int arx_dense_update(float* w, float* acc, const float* g, int64_t n, float lr,
                                const float* grad_scale_dev, int opt, void* stream) {
  if (!w || !g || n < 0) return ARX_E_BADARG;
  if (opt == ARX_OPT_ADAGRAD && !acc) return ARX_E_BADARG;
  if (opt != ARX_OPT_ADAGRAD && opt != ARX_OPT_SGD) return ARX_E_BADARG;
  if (n == 0) return ARX_OK;
  dense_update_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(w, acc, g, n, lr, grad_scale_dev, opt);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_scale_mask(const float* x, const float* mask, float scale, int64_t n, float* y,
                              void* stream) {
  if (!x || !y || n < 0) return ARX_E_BADARG;
  if (n == 0) return ARX_OK;
  scale_mask_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(x, mask, scale, n, y);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_abi_version(void) { return ARX_ABI_VERSION; }

extern "C" const char* arx_build_info(void) {
  return "libarx_b200 sm_100a nvcc " __DATE__ " " __TIME__;
}
