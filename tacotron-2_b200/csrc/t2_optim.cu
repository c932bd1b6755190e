// t2_optim.cu — fused multi-tensor optimizer step over flat fp32 buffers.
// Replaces wavenet_vocoder/models/wavenet.py:586-613 (clip_by_norm(100) + clip_by_value(5) per tensor, Adam, EMA)
// and tacotron/models/tacotron.py:429-437 (clip_by_global_norm(1.0), Adam). Adam follows tf.train.AdamOptimizer:
//   lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  theta -= lr_t * m / (sqrt(v) + eps)   
#include "../../include/t2b200.h"
#include "t2_common.cuh"

namespace t2 {
namespace {

__device__ __forceinline__ int find_tensor(const long long* __restrict__ offs, int n, long long e) {
  int lo = 0, hi = n;  // offs[lo] <= e < offs[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (offs[mid] <= e) lo = mid; else hi = mid;
  }
  return lo;
}

// per-tensor sum of squares of (grad * grad_scale): one block handles a contiguous chunk
__global__ void sumsq_kernel(const float* __restrict__ g, const long long* __restrict__ offs, int nt, long long n,
                             float gscale, float* __restrict__ norms) {
  const long long chunk = 4096;
  const long long e0 = blockIdx.x * chunk;
  const long long e1 = e0 + chunk < n ? e0 + chunk : n;
  int t = find_tensor(offs, nt, e0);
  float acc = 0.f;
  __shared__ float red[8];
  long long e = e0 + threadIdx.x;
  while (true) {
    const long long tend = offs[t + 1] < e1 ? offs[t + 1] : e1;
    acc = 0.f;
    for (; e < tend; e += blockDim.x) {
      const float v = g[e] * gscale;
      acc += v * v;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int i = 0; i < (blockDim.x >> 5); ++i) s += red[i];
      if (s != 0.f) atomicAdd(norms + t, s);
    }
    __syncthreads();
    if (tend >= e1) break;
    ++t;
    // re-align this thread's cursor to the start of the next tensor
    e = offs[t] + threadIdx.x;
  }
}

struct AdamArgs {
  float* p; const float* g; float* m; float* v; float* ema;
  const long long* offs; int nt; long long n;
  float lr_t, b1, b2, eps, gscale, max_norm, max_value, gclip, ema_decay;
  const float* norms;  // per-tensor sumsq; norms[nt] = global sumsq
};
__global__ void adam_kernel(AdamArgs a) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= a.n) return;
  float g = a.g[e] * a.gscale;
  if (a.gclip > 0.f) {
    const float gn = sqrtf(a.norms[a.nt]);
    g *= a.gclip / fmaxf(gn, a.gclip);
  } else if (a.max_norm > 0.f) {
    const int t = find_tensor(a.offs, a.nt, e);
    const float tn = sqrtf(a.norms[t]);
    g *= a.max_norm / fmaxf(tn, a.max_norm);
  }
  if (a.max_value > 0.f) g = fminf(fmaxf(g, -a.max_value), a.max_value);
  const float m = a.b1 * a.m[e] + (1.f - a.b1) * g;
  const float v = a.b2 * a.v[e] + (1.f - a.b2) * g * g;
  a.m[e] = m;
  a.v[e] = v;
  const float p = a.p[e] - a.lr_t * m / (sqrtf(v) + a.eps);
  a.p[e] = p;
  if (a.ema) a.ema[e] -= (1.f - a.ema_decay) * (a.ema[e] - p);
}
__global__ void total_kernel(float* norms, int nt) {
  float s = 0.f;
  for (int i = threadIdx.x; i < nt; i += 32) s += norms[i];
  s = warp_sum(s);
  if (threadIdx.x == 0) norms[nt] = s;
}

}  // namespace
}  // namespace t2

extern "C" int t2_adam_step(float* d_params, const float* d_grads, float* d_m, float* d_v, float* d_ema,
                            const long long* d_offsets, int n_tensors, long long n_total, float lr, float beta1,
                            float beta2, float eps, int step, float grad_scale, float max_norm, float max_value,
                            float global_norm_clip, float ema_decay, float* d_scratch, void* stream) {
  using namespace t2;
  T2_REQUIRE(d_params && d_grads && d_m && d_v && d_offsets && d_scratch, T2_ERR_INVALID_ARG, "adam: null pointer");
  T2_REQUIRE(step >= 1 && n_tensors >= 1 && n_total >= 1, T2_ERR_INVALID_ARG, "adam: bad step / sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool need_norms = max_norm > 0.f || global_norm_clip > 0.f;
  if (need_norms) {
    T2_CHECK_CUDA(cudaMemsetAsync(d_scratch, 0, (n_tensors + 1) * sizeof(float), st));
    sumsq_kernel<<<(unsigned)((n_total + 4095) / 4096), 256, 0, st>>>(d_grads, d_offsets, n_tensors, n_total, grad_scale, d_scratch);
    if (global_norm_clip > 0.f) total_kernel<<<1, 32, 0, st>>>(d_scratch, n_tensors);
  }
  AdamArgs a;
  a.p = d_params; a.g = d_grads; a.m = d_m; a.v = d_v; a.ema = d_ema; a.offs = d_offsets; a.nt = n_tensors; a.n = n_total;
  a.lr_t = float(double(lr) * sqrt(1.0 - pow(double(beta2), step)) / (1.0 - pow(double(beta1), step)));
  a.b1 = beta1; a.b2 = beta2; a.eps = eps; a.gscale = grad_scale; a.max_norm = max_norm; a.max_value = max_value;
  a.gclip = global_norm_clip; a.ema_decay = ema_decay; a.norms = d_scratch;
  adam_kernel<<<(unsigned)((n_total + 255) / 256), 256, 0, st>>>(a);
  T2_CHECK_CUDA(cudaGetLastError());
  return T2_OK;
}
