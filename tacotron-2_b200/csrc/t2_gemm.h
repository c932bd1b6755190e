// t2_gemm.h — host-side interface of the tcgen05 GEMM engine (see t2_gemm.cuh for the kernels).
#pragma once
#include <cuda_runtime.h>
#include <string.h>

#include "t2_gemm_types.h"

namespace t2 {

// channels-last bf16 activation tensor [L][B][T][ld]; the first C channels of each row are addressable
struct ActT {
  const void* ptr;
  int C, T, B, L, ld;
};
inline ActT make_act(const void* p, int C, int T, int B, int L = 1, int ld = -1) {
  ActT a;
  a.ptr = p; a.C = C; a.T = T; a.B = B; a.L = L; a.ld = ld < 0 ? C : ld;
  return a;
}

struct ActGemmCall {
  ActT a[4];
  int na;
  Seg seg[kMaxSeg];
  int nseg;
  const void* w;   // packed bf16 weights [wL][wN][wK], K contiguous
  int wN, wK, wL, w_layer, w_k0;
  int T, B;
  int n_tiles;     // grid.y
  int ksplit;      // grid.z: CTAs sharing one output tile over slices of K (0/1 = off; epilogue must accumulate atomically)
  EpiArgs epi;
};

void set_timing_buffer(long long* p);
bool pdl_enabled();
// launch any kernel with the programmatic-dependent-launch attribute (the kernel must call pdl_wait() before it touches
// global memory; see t2_common.cuh)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
int launch_act_gemm(int epi, int BN, const ActGemmCall& c, cudaStream_t stream);
int launch_wgrad(const ActT* maps, int nmaps, const WgradTile* tiles_dev, int ntiles, float* out,
                 int T, int B, cudaStream_t stream);

}  // namespace t2
