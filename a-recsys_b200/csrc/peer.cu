// Peer-memory exchange over NVLink 5 / NVSwitch for the row-sharded path.
//
// One process per GPU.  Every rank allocates its receive buffers with arx_peer_alloc, exports them as CUDA IPC
// handles, and maps the other ranks' buffers (arx_peer_open): the kernels of the step then write partial results
// STRAIGHT into the owner rank's memory — pooled partial sums by red.global.add.v4.f32 from inside the lookup kernel
// (csrc/pool.cu, push mode: replaces reduce-scatter + its [G*mb, w] staging buffer), gradient rows by 128-bit stores
// (arx_peer_push_rows: replaces all-gather / all-reduce) — and one device-side barrier (arx_peer_barrier) per
// exchange phase orders the pushes before the consumer kernels.  No host synchronisation, no NCCL launch on the
// dependent chain; everything is stream-ordered and CUDA-graph capturable.
//
// The reference has no counterpart: it is single-GPU; the nearest analogue is the tower-wise
// `tf.device` placement at lstm/run.py:221-229.
#include "arx_common.cuh"

namespace {

// Flag block of one rank: flags[src] = last epoch rank `src` has arrived at.  Rank r arrives by writing the new epoch
// into flags[r] of EVERY rank (release at system scope, after a system fence that orders the CTA's — and, by stream
// order, every earlier kernel's — peer writes), then waits until all G entries of its own block reached the epoch.
// A bounded spin (≈ timeout_ns) raises *err instead of hanging the GPU when a peer never arrives.
__global__ void peer_barrier_kernel(unsigned int* const* __restrict__ flags, int rank, int G,
                                    unsigned int* __restrict__ epoch, long long timeout_ns, int* __restrict__ err) {
  __shared__ unsigned int s_ep;
  if (threadIdx.x == 0) s_ep = *epoch + 1;
  __syncthreads();
  const unsigned int ep = s_ep;
  const int j = threadIdx.x;
  if (j < G) {
    __threadfence_system();
    unsigned int* dst = flags[j] + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(ep) : "memory");
    const unsigned int* mine = flags[rank] + j;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (true) {
      unsigned int v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int)(v - ep) >= 0) break;
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if ((long long)(t1 - t0) > timeout_ns) { if (err) atomicExch(err, 1 + j); break; }
      __nanosleep(64);
    }
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0) *epoch = ep;
}

// src [rows, width] (row pitch src_stride) -> rows [row0, row0 + rows) of every destination in dst[0..G) (row pitch
// dst_stride), 128-bit accesses; mode 0: plain stores to every rank except `skip` (pass -1 to include all),
// mode 1: red.add to every rank.  One grid-stride loop per destination, destinations interleaved over the CTAs so
// that all NVLink ports are busy from the first wave on.
__global__ void __launch_bounds__(256)
peer_push_rows_kernel(const float* __restrict__ src, long long rows, int width4, long long src_stride,
                      float* const* __restrict__ dst, long long dst_stride, long long row0, int G, int skip, int mode) {
  const long long per = rows * width4;
  const long long total = per * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long slab = i / width4;            // (row, destination) pair: destinations vary fastest across warps
    const int c = (int)(i - slab * width4);
    const int g = (int)(slab % G);
    const long long r = slab / G;
    if (g == skip) continue;
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + r * src_stride) + c);
    float* p = dst[g] + (row0 + r) * dst_stride + (long long)c * 4;
    if (mode == 0) {
      asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    } else {
      asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                   : "memory");
    }
  }
}

// Several row blocks in one launch (the backward exchange of a step: dU | dPt | dts stored, dPs | dbs added): each
// segment owns a contiguous range of CTAs proportional to its size.
constexpr int kPushSegs = 8;
struct PushManyParams {
  arx_peer_seg seg[kPushSegs];
  int block_end[kPushSegs];
  int n_segs, G;
};

__global__ void __launch_bounds__(256) peer_push_many_kernel(const PushManyParams mp) {
  int s = 0;
  while (s + 1 < mp.n_segs && (int)blockIdx.x >= mp.block_end[s]) ++s;
  const int b0 = s == 0 ? 0 : mp.block_end[s - 1];
  const arx_peer_seg& sg = mp.seg[s];
  const int width4 = (int)(sg.width >> 2);
  const int G = mp.G;
  const long long total = sg.rows * width4 * G;
  const long long nthreads = (long long)(mp.block_end[s] - b0) * blockDim.x;
  for (long long i = (long long)(blockIdx.x - b0) * blockDim.x + threadIdx.x; i < total; i += nthreads) {
    const long long slab = i / width4;
    const int c = (int)(i - slab * width4);
    const int g = (int)(slab % G);
    const long long r = slab / G;
    const float4 v = __ldg(reinterpret_cast<const float4*>(sg.src + r * sg.src_stride) + c);
    float* p = sg.dst[g] + (sg.row0 + r) * sg.dst_stride + (long long)c * 4;
    if (sg.mode == 0) {
      asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    } else {
      asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                   : "memory");
    }
  }
}

}  // namespace

extern "C" int arx_peer_push_many(const arx_peer_seg* segs, int n_segs, int G, void* stream) {
  if (!segs || n_segs < 1 || n_segs > kPushSegs || G < 1) return ARX_E_BADARG;
  PushManyParams mp{};
  int blocks = 0, k = 0;
  for (int i = 0; i < n_segs; ++i) {
    const arx_peer_seg& q = segs[i];
    if (!q.src || !q.dst || q.rows < 0 || q.width < 4 || (q.width % 4) || (q.src_stride % 4) || (q.dst_stride % 4) ||
        (((uintptr_t)q.src) & 15) || q.mode < 0 || q.mode > 1)
      return ARX_E_BADARG;
    if (q.rows == 0) continue;
    const long long total = q.rows * (q.width / 4) * G;
    blocks += (int)std::min<long long>((total + 1023) / 1024, (long long)arx_num_sms() * 4);
    mp.seg[k] = q;
    mp.block_end[k] = blocks;
    ++k;
  }
  if (k == 0) return ARX_OK;
  mp.n_segs = k; mp.G = G;
  peer_push_many_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(mp);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_peer_alloc(int64_t bytes, void** ptr) {
  if (!ptr || bytes <= 0) return ARX_E_BADARG;
  void* p = nullptr;
  if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess) { cudaGetLastError(); return ARX_E_CAPACITY; }
  if (cudaMemset(p, 0, (size_t)bytes) != cudaSuccess) { cudaFree(p); cudaGetLastError(); return ARX_E_LAUNCH; }
  if (cudaDeviceSynchronize() != cudaSuccess) { cudaFree(p); cudaGetLastError(); return ARX_E_LAUNCH; }
  *ptr = p;
  return ARX_OK;
}

extern "C" int arx_peer_free(void* ptr) {
  if (!ptr) return ARX_OK;
  return cudaFree(ptr) == cudaSuccess ? ARX_OK : ARX_E_BADARG;
}

extern "C" int arx_peer_export(void* ptr, unsigned char* handle64) {
  if (!ptr || !handle64) return ARX_E_BADARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, ptr) != cudaSuccess) { cudaGetLastError(); return ARX_E_UNSUPPORTED; }
  memcpy(handle64, &h, 64);
  return ARX_OK;
}

extern "C" int arx_peer_open(const unsigned char* handle64, void** ptr) {
  if (!handle64 || !ptr) return ARX_E_BADARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return ARX_E_UNSUPPORTED; }
  *ptr = p;
  return ARX_OK;
}

extern "C" int arx_peer_close(void* ptr) {
  if (!ptr) return ARX_OK;
  return cudaIpcCloseMemHandle(ptr) == cudaSuccess ? ARX_OK : ARX_E_BADARG;
}

extern "C" int arx_peer_barrier(uint32_t* const* flags, int rank, int G, uint32_t* epoch, int64_t timeout_ns, int32_t* err,
                                void* stream) {
  if (!flags || !epoch || G < 1 || G > 32 || rank < 0 || rank >= G || timeout_ns <= 0) return ARX_E_BADARG;
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((unsigned int* const*)flags, rank, G, (unsigned int*)epoch,
                                                          (long long)timeout_ns, (int*)err);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_peer_push_rows(const float* src, int64_t rows, int64_t width, int64_t src_stride, float* const* dst,
                                  int64_t dst_stride, int64_t row0, int G, int skip, int mode, void* stream) {
  if (!src || !dst || rows < 0 || width < 4 || (width % 4) || (src_stride % 4) || (dst_stride % 4) || G < 1 || mode < 0 || mode > 1)
    return ARX_E_BADARG;
  if (((uintptr_t)src & 15) != 0) return ARX_E_BADARG;
  if (rows == 0) return ARX_OK;
  const long long total = rows * (width / 4) * G;
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)arx_num_sms() * 8);
  peer_push_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, (long long)rows, (int)(width / 4), (long long)src_stride,
                                                                dst, (long long)dst_stride, (long long)row0, G, skip, mode);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}
