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
  // 4 consecutive elements per thread: tensor offsets (and n) are multiples of 4, so they share one tensor
  const long long e = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4;
  if (e >= a.n) return;
  float gs = a.gscale;
  if (a.gclip > 0.f) {
    gs *= a.gclip / fmaxf(sqrtf(a.norms[a.nt]), a.gclip);
  } else if (a.max_norm > 0.f) {
    const int t = find_tensor(a.offs, a.nt, e);
    gs *= a.max_norm / fmaxf(sqrtf(a.norms[t]), a.max_norm);
  }
  const float4 g4 = *reinterpret_cast<const float4*>(a.g + e);
  float4 m4 = *reinterpret_cast<const float4*>(a.m + e);
  float4 v4 = *reinterpret_cast<const float4*>(a.v + e);
  float4 p4 = *reinterpret_cast<const float4*>(a.p + e);
  float4 e4 = a.ema ? *reinterpret_cast<const float4*>(a.ema + e) : make_float4(0, 0, 0, 0);
  float* gp = const_cast<float*>(reinterpret_cast<const float*>(&g4));
  float* mp = reinterpret_cast<float*>(&m4);
  float* vp = reinterpret_cast<float*>(&v4);
  float* pp = reinterpret_cast<float*>(&p4);
  float* ep = reinterpret_cast<float*>(&e4);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float g = gp[i] * gs;
    if (a.max_value > 0.f) g = fminf(fmaxf(g, -a.max_value), a.max_value);
    mp[i] = a.b1 * mp[i] + (1.f - a.b1) * g;
    vp[i] = a.b2 * vp[i] + (1.f - a.b2) * g * g;
    pp[i] = pp[i] - a.lr_t * mp[i] / (sqrtf(vp[i]) + a.eps);
    ep[i] -= (1.f - a.ema_decay) * (ep[i] - pp[i]);
  }
  *reinterpret_cast<float4*>(a.m + e) = m4;
  *reinterpret_cast<float4*>(a.v + e) = v4;
  *reinterpret_cast<float4*>(a.p + e) = p4;
  if (a.ema) *reinterpret_cast<float4*>(a.ema + e) = e4;
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
    sumsq_kernel<<<(unsigned)((n_total + 4095) / 4096), 256, 0, st>>>(d_grads, d_offsets, n_tensors, n_total, grad_scale, d_scratch); t2_count_launch();
    if (global_norm_clip > 0.f) total_kernel<<<1, 32, 0, st>>>(d_scratch, n_tensors); t2_count_launch();
  }
  AdamArgs a;
  a.p = d_params; a.g = d_grads; a.m = d_m; a.v = d_v; a.ema = d_ema; a.offs = d_offsets; a.nt = n_tensors; a.n = n_total;
  a.lr_t = float(double(lr) * sqrt(1.0 - pow(double(beta2), step)) / (1.0 - pow(double(beta1), step)));
  a.b1 = beta1; a.b2 = beta2; a.eps = eps; a.gscale = grad_scale; a.max_norm = max_norm; a.max_value = max_value;
  a.gclip = global_norm_clip; a.ema_decay = ema_decay; a.norms = d_scratch;
  T2_REQUIRE(n_total % 4 == 0, T2_ERR_INVALID_ARG, "adam: flat buffers must hold a multiple of 4 elements (16-byte aligned tensors)");
  adam_kernel<<<(unsigned)((n_total / 4 + 255) / 256), 256, 0, st>>>(a); t2_count_launch();
  T2_CHECK_CUDA(cudaGetLastError());
  return T2_OK;
}
