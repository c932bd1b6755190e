// t2_api.cu — error plumbing of the C-ABI plus the engine-level entry points used by the unit tests.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/t2b200.h"
#include "t2_common.cuh"
#include "t2_gemm.h"

#include <atomic>
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
void t2_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" long long t2_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int t2_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

extern "C" const char* t2_last_error(void) { return g_err; }
extern "C" int t2_abi_version(void) { return T2B200_ABI_VERSION; }

// D[b,t,n] = act( sum_s sum_k A[b, t + shift_s, k] * W[n, s*Kseg + k] + bias[n] ), bf16 in, fp32 accumulate.
extern "C" int t2_dbg_conv_gemm(const void* a, int B, int T, int C, int ld, const int* shifts, int nshift,
                                const void* w, int N, int BN, const float* bias, int relu, void* out_bf16,
                                float* out_f32, void* stream) {
  using namespace t2;
  T2_REQUIRE(nshift >= 1 && nshift <= kMaxSeg, T2_ERR_INVALID_ARG, "nshift out of range");
  T2_REQUIRE(BN == 128 || BN == 256, T2_ERR_UNSUPPORTED_SHAPE, "BN must be 128 or 256");
  ActGemmCall c;
  memset(&c, 0, sizeof(c));
  c.a[0] = make_act(a, C, T, B, 1, ld);
  c.na = 1;
  const int nkb = (C + kBK - 1) / kBK;
  for (int s = 0; s < nshift; ++s) c.seg[s] = Seg{0, shifts[s], 0, nkb, 0, 1};
  c.nseg = nshift;
  c.w = w; c.wN = N; c.wK = nshift * nkb * kBK; c.wL = 1; c.w_layer = 0;
  c.T = T; c.B = B; c.n_tiles = (N + BN - 1) / BN;
  c.epi.ptr[0] = out_bf16; c.epi.ptr[1] = const_cast<float*>(bias); c.epi.ptr[2] = out_f32;
  c.epi.i[0] = N; c.epi.i[1] = relu; c.epi.i[2] = N;
  return launch_act_gemm(EPI_BIAS_ACT, BN, c, static_cast<cudaStream_t>(stream));
}

// dW[m, n] = scale * sum_{b,t} A[b, t + shift_a, m] * Bm[b, t, n]   (fp32 out [Ca, Cb])
extern "C" int t2_dbg_wgrad(const void* a, int Ca, const void* bm, int Cb, int B, int T, int shift_a,
                            float scale, float* out, void* stream) {
  using namespace t2;
  ActT maps[2] = {make_act(a, Ca, T, B), make_act(bm, Cb, T, B)};
  std::vector<WgradTile> tiles;
  for (int m0 = 0; m0 < Ca; m0 += 128)
    for (int n0 = 0; n0 < Cb; n0 += 256) {
      WgradTile t;
      memset(&t, 0, sizeof(t));
      t.a_map = 0; t.a_ch0 = m0; t.a_shift = shift_a; t.a_layer = 0;
      t.b_map = 1; t.b_ch0 = n0; t.b_shift = 0; t.b_layer = 0;
      t.out_off = (long long)m0 * Cb + n0; t.ldc = Cb;
      t.m_valid = Ca - m0 < 128 ? Ca - m0 : 128;
      t.n_valid = Cb - n0 < 256 ? Cb - n0 : 256;
      t.scale = scale; t.accumulate = 0; t.div = nullptr;
      tiles.push_back(t);
    }
  WgradTile* dt = nullptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  T2_CHECK_CUDA(cudaMallocAsync(&dt, tiles.size() * sizeof(WgradTile), st));
  T2_CHECK_CUDA(cudaMemcpyAsync(dt, tiles.data(), tiles.size() * sizeof(WgradTile), cudaMemcpyHostToDevice, st));
  int rc = launch_wgrad(maps, 2, dt, int(tiles.size()), out, T, B, st);
  cudaStreamSynchronize(st);  // debug entry: tiles is a host temporary
  cudaFreeAsync(dt, st);
  return rc;
}

// debug: when non-NULL every act_gemm CTA writes stamps to d_buf[(launch offset + cta) * 16 + slot]: slots 0-6 clock64() at entry,
// setup done, first stage landed, MMAs issued, accumulator ready, epilogue done, teardown; 8-10 %globaltimer (ns) at entry, after the
// programmatic-dependent-launch wait, at exit; 11 = SM id. Each launch advances the offset by its CTA count.
extern "C" int t2_dbg_set_timing_buffer(long long* d_buf) {
  t2::set_timing_buffer(d_buf);
  return T2_OK;
}

// The counter-hash behind every in-kernel dropout / zoneout mask, exported so that a parity test can rebuild the exact
// masks a training step drew and feed them to the CPU oracle: out[i] = hash_uniform32(hash_seed(seed, stream), i0 + i).
// An element is KEPT (dropout) / UPDATED (zoneout) iff out[i] >= rate.
__global__ void rng_uniform_kernel(unsigned long long seed, unsigned int stream, long long i0, long long n, float* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) out[i] = t2::hash_uniform32(t2::hash_seed(seed, stream), (unsigned long long)(i0 + i));
}
extern "C" int t2_rng_uniform_f32(unsigned long long seed, unsigned int stream_id, long long first_index, long long n, float* d_out,
                                  void* stream) {
  T2_REQUIRE(n >= 0 && (n == 0 || d_out != nullptr), T2_ERR_INVALID_ARG, "rng_uniform: bad arguments");
  if (n == 0) return T2_OK;
  rng_uniform_kernel<<<unsigned((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(seed, stream_id, first_index, n, d_out);
  t2_count_launch();
  T2_CHECK_CUDA(cudaGetLastError());
  return T2_OK;
}

// sizeof of the POD structs that cross the C-ABI: lets a binding (ctypes, cgo, ...) assert that its mirror matches this build
extern "C" int t2_struct_size(const char* name) {
  if (!name) return -1;
  if (!strcmp(name, "t2_wn_config_t")) return int(sizeof(t2_wn_config_t));
  if (!strcmp(name, "t2_wn_sizes_t")) return int(sizeof(t2_wn_sizes_t));
  if (!strcmp(name, "t2_taco_config_t")) return int(sizeof(t2_taco_config_t));
  if (!strcmp(name, "t2_cbhg_config_t")) return int(sizeof(t2_cbhg_config_t));
  if (!strcmp(name, "t2_audio_config_t")) return int(sizeof(t2_audio_config_t));
  return -1;
}
