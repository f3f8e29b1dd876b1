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

// fp32 -> nearest tf32 (round to nearest even on the 13 dropped mantissa bits).  The tensor core
// TRUNCATES fp32 operands to tf32; truncation is biased (every product shrinks, ~1e-3 relative),
// pre-rounding makes the error zero-mean and half as large.
__device__ __forceinline__ float round_tf32(float x) {
  unsigned int u = __float_as_uint(x);
  if ((u & 0x7f800000u) == 0x7f800000u) return x;          // inf / nan untouched
  u += 0x00000fffu + ((u >> 13) & 1u);
  return __uint_as_float(u & 0xffffe000u);
}

__global__ void round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = round_tf32(src[i]);
}

// dst[c, r] = src[r, c]; 32x32 tile through padded shared memory, coalesced both ways
template <bool ROUND>
__global__ void transpose_kernel(const float* __restrict__ src, long long rows, long long cols,
                                 float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const long long c0 = (long long)blockIdx.x * 32, r0 = (long long)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[c * rows + r] = ROUND ? round_tf32(tile[threadIdx.x][i]) : tile[threadIdx.x][i];
  }
}

// out[c] = sum_r x[r, c]  (bias gradients: column sums of the score gradient)
__global__ void colsum_kernel(const float* __restrict__ x, long long rows, long long cols, long long ld,
                              float* __restrict__ out, int accumulate) {
  __shared__ float red[8][33];
  const long long c = (long long)blockIdx.x * 32 + threadIdx.x;
  const long long rchunk = (rows + gridDim.y - 1) / gridDim.y;
  const long long rb = (long long)blockIdx.y * rchunk, re = min(rows, rb + rchunk);
  float s = 0.f;
  if (c < cols)
    for (long long r = rb + threadIdx.y; r < re; r += blockDim.y) s += x[r * ld + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    if (gridDim.y > 1 || accumulate) atomicAdd(out + c, t);
    else out[c] = t;
  }
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

extern "C" int arx_transpose(const float* src, int64_t rows, int64_t cols, float* dst, int round_tf32_out,
                             void* stream) {
  if (!src || !dst || rows < 0 || cols < 0) return ARX_E_BADARG;
  if (rows == 0 || cols == 0) return ARX_OK;
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
  if (grid.y > 65535u) return ARX_E_UNSUPPORTED;
  if (round_tf32_out) transpose_kernel<true><<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(src, rows, cols, dst);
  else transpose_kernel<false><<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(src, rows, cols, dst);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_round_tf32(const float* src, float* dst, int64_t n, void* stream) {
  if (!src || !dst || n < 0) return ARX_E_BADARG;
  if (n == 0) return ARX_OK;
  round_tf32_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(src, dst, n);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_colsum(const float* x, int64_t rows, int64_t cols, int64_t ld, float* out, void* stream) {
  if (!x || !out || rows < 0 || cols < 0 || ld < cols) return ARX_E_BADARG;
  if (cols == 0) return ARX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned gx = (unsigned)((cols + 31) / 32);
  unsigned gy = 1;
  const unsigned want = (unsigned)arx_num_sms() * 4;
  if (gx < want) {
    long long g = (want + gx - 1) / gx, cap = rows / 64;
    if (cap < 1) cap = 1;
    gy = (unsigned)(g < cap ? g : cap);
  }
  if (gy > 1 && cudaMemsetAsync(out, 0, sizeof(float) * (size_t)cols, st) != cudaSuccess) return ARX_E_LAUNCH;
  colsum_kernel<<<dim3(gx, gy), dim3(32, 8), 0, st>>>(x, rows, cols, ld, out, 0);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_abi_version(void) { return ARX_ABI_VERSION; }

extern "C" const char* arx_build_info(void) {
  return "libarx_b200 sm_100a nvcc " __DATE__ " " __TIME__;
}
