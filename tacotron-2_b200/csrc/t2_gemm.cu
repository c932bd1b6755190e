// t2_gemm.cu — host side of the tcgen05 GEMM engine: TMA tensor-map encoding and kernel launches.
#include <stdlib.h>
#include <mutex>

#include "t2_gemm.cuh"
#include "t2_gemm.h"

namespace t2 {

static long long* g_timing_buffer = nullptr;
void set_timing_buffer(long long* p) { g_timing_buffer = p; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 4-D map over a channels-last bf16 activation tensor: dims (C, T, B, L), box (64, rows, 1, 1), 128B swizzle.
static int encode_act_map(CUtensorMap* m, const ActT& a, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  T2_REQUIRE(fn != nullptr, T2_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  T2_REQUIRE(a.ptr && (reinterpret_cast<uintptr_t>(a.ptr) & 15) == 0, T2_ERR_INVALID_ARG,
             "activation pointer must be non-null and 16-byte aligned");
  T2_REQUIRE(a.ld % 8 == 0 && a.C >= 1 && a.C <= a.ld, T2_ERR_UNSUPPORTED_SHAPE,
             "activation row pitch must be a multiple of 8 elements (ld=%d C=%d)", a.ld, a.C);
  cuuint64_t dims[4] = {cuuint64_t(a.C), cuuint64_t(a.T), cuuint64_t(a.B), cuuint64_t(a.L)};
  cuuint64_t strides[3] = {cuuint64_t(a.ld) * 2, cuuint64_t(a.ld) * 2 * a.T,
                           cuuint64_t(a.ld) * 2 * a.T * a.B};
  cuuint32_t box[4] = {64, cuuint32_t(box_rows), 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a.ptr), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  T2_REQUIRE(r == CUDA_SUCCESS, T2_ERR_CUDA, "cuTensorMapEncodeTiled(act) failed: %d (C=%d T=%d B=%d L=%d ld=%d)",
             int(r), a.C, a.T, a.B, a.L, a.ld);
  return T2_OK;
}

// 3-D map over a channels-last bf16 OUTPUT tensor [B][T][C] for the epilogue's TMA stores: box 64 columns x 32 rows (one TMEM lane
// quarter), 128-byte swizzle; rows >= T of an item and columns >= C are clipped by the hardware.
static int encode_out_map(CUtensorMap* m, const void* ptr, int C, int T, int B) {
  EncodeTiledFn fn = get_encode_fn();
  T2_REQUIRE(fn != nullptr, T2_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  T2_REQUIRE(ptr && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && C % 8 == 0, T2_ERR_INVALID_ARG,
             "epilogue output must be non-null, 16-byte aligned, with a row pitch that is a multiple of 8 elements (C=%d)", C);
  cuuint64_t dims[3] = {cuuint64_t(C), cuuint64_t(T), cuuint64_t(B)};
  cuuint64_t strides[2] = {cuuint64_t(C) * 2, cuuint64_t(C) * 2 * T};
  cuuint32_t box[3] = {64, 32, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  T2_REQUIRE(r == CUDA_SUCCESS, T2_ERR_CUDA, "cuTensorMapEncodeTiled(out) failed: %d (C=%d T=%d B=%d)", int(r), C, T, B);
  return T2_OK;
}

// 3-D map over packed bf16 weights [L][N][K]: dims (K, N, L), box (64, rows, 1)
static int encode_wt_map(CUtensorMap* m, const void* w, int N, int K, int L, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  T2_REQUIRE(fn != nullptr, T2_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  T2_REQUIRE(w && (reinterpret_cast<uintptr_t>(w) & 15) == 0, T2_ERR_INVALID_ARG,
             "weight pointer must be non-null and 16-byte aligned");
  T2_REQUIRE(K % 8 == 0, T2_ERR_UNSUPPORTED_SHAPE, "packed weight K must be a multiple of 8 (K=%d)", K);
  cuuint64_t dims[3] = {cuuint64_t(K), cuuint64_t(N), cuuint64_t(L)};
  cuuint64_t strides[2] = {cuuint64_t(K) * 2, cuuint64_t(K) * 2 * N};
  cuuint32_t box[3] = {64, cuuint32_t(box_rows), 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  T2_REQUIRE(r == CUDA_SUCCESS, T2_ERR_CUDA, "cuTensorMapEncodeTiled(weight) failed: %d (N=%d K=%d L=%d)", int(r),
             N, K, L);
  return T2_OK;
}

// programmatic dependent launch for the GEMM kernels (T2_PDL=0 in the environment turns it off for A/B measurements)
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("T2_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

// cluster size along M for the weight-tile multicast (T2_CLUSTER in the environment; default 1 = off: measured on B200 the
// per-layer GEMMs are bound by bytes in flight per SM, not by L2 reads - 2.51 / 2.57 / 2.58 / 2.68 ms per step at 1 / 2 / 4 / 8)
int cluster_pref() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("T2_CLUSTER");
    v = e ? atoi(e) : 1;
    if (v != 1 && v != 2 && v != 4 && v != 8) v = 1;
  }
  return v;
}
static int pick_cluster(int BN, const dim3& grid) {
  int cs = cluster_pref();
  if (BN < 128 || grid.z > 1) return 1;                 // the swapped recurrence GEMMs (BN = 32) and split-K keep single CTAs
  while (cs > 1 && (grid.x % cs != 0 || (BN / cs) % 8 != 0)) cs >>= 1;
  return cs;
}

template <int EPI, int BN, int NT = 1>
static int launch_one(const GemmArgs& g, dim3 grid, int cs, cudaStream_t stream) {
  using Cfg = ActGemmCfg<BN>;
  static bool configured = false;
  if (!configured) {
    T2_CHECK_CUDA(cudaFuncSetAttribute(act_gemm_kernel<EPI, BN, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::kSmemBytes));
    configured = true;
  }
  grid.y = (grid.y + NT - 1) / NT;
  // every launch gets its own slice of the timing buffer (so a captured graph stamps each of its kernel nodes separately)
  if (g.dbg) g_timing_buffer = g.dbg + size_t(grid.x) * grid.y * grid.z * kDbgSlots;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = dim3(kActGemmThreads); cfg.dynamicSmemBytes = Cfg::kSmemBytes; cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cs > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cs; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  T2_CHECK_CUDA(cudaLaunchKernelEx(&cfg, act_gemm_kernel<EPI, BN, NT>, g));
  t2_count_launch();
  return T2_OK;
}

// CTA-pair kernels (tcgen05 cta_group::2): T2_PAIR=0 in the environment keeps the single-CTA kernels
bool pair_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("T2_PAIR"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

template <int EPI, int BN, int NT = 1>
static int launch_pair(const GemmArgs& g, dim3 grid, cudaStream_t stream) {
  using Cfg = ActGemm2Cfg<BN>;
  static bool configured = false;
  if (!configured) {
    T2_CHECK_CUDA(cudaFuncSetAttribute(act_gemm2_kernel<EPI, BN, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  grid.y = (grid.y + NT - 1) / NT;
  if (g.dbg) g_timing_buffer = g.dbg + size_t(grid.x) * grid.y * grid.z * kDbgSlots;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = dim3(kActGemmThreads); cfg.dynamicSmemBytes = Cfg::kSmemBytes; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;       // the cluster shape (2,1,1) is compiled into the kernel
  T2_CHECK_CUDA(cudaLaunchKernelEx(&cfg, act_gemm2_kernel<EPI, BN, NT>, g));
  t2_count_launch();
  return T2_OK;
}

int launch_act_gemm(int epi, int BN, const ActGemmCall& c, cudaStream_t stream) {
  T2_REQUIRE(c.na >= 1 && c.na <= 4 && c.nseg >= 1 && c.nseg <= kMaxSeg, T2_ERR_INVALID_ARG,
             "act_gemm: bad map/segment count (%d, %d)", c.na, c.nseg);
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  for (int i = 0; i < 4; ++i) {
    int rc = encode_act_map(&g.amap[i], c.a[i < c.na ? i : 0], kBM);
    if (rc) return rc;
  }
  dim3 grid((c.T + kBM - 1) / kBM * c.B, c.n_tiles, c.ksplit > 1 ? c.ksplit : 1);
  // CTA pairs when the M tiles pair up (even count, no split-K) and the tile is wide enough for a half to be a whole swizzle atom
  const bool pair = pair_enabled() && BN >= 128 && grid.x % 2 == 0 && grid.z == 1;
  const int cs = pair ? 1 : pick_cluster(BN, grid);
  // weight-tile rows one TMA box fetches: half a tile per CTA of a pair, 1/cs per CTA of a multicast cluster
  int rc = encode_wt_map(&g.bmap, c.w, c.wN, c.wK, c.wL, pair ? BN / 2 : BN / cs);
  if (rc) return rc;
  int ktot = 0;
  for (int s = 0; s < c.nseg; ++s) {
    g.seg[s] = c.seg[s];
    T2_REQUIRE(c.seg[s].map >= 0 && c.seg[s].map < c.na && c.seg[s].nkb > 0 && c.seg[s].nlayers > 0,
               T2_ERR_INVALID_ARG, "act_gemm: bad segment %d", s);
    ktot += c.seg[s].nkb * c.seg[s].nlayers * kBK;
  }
  T2_REQUIRE(c.w_k0 + ktot <= ((c.wK + kBK - 1) / kBK) * kBK, T2_ERR_INVALID_ARG,
             "act_gemm: segments cover K=%d but packed weight has K=%d", ktot, c.wK);
  g.nseg = c.nseg;
  g.T = c.T;
  g.tiles_per_b = (c.T + kBM - 1) / kBM;
  g.b_layer = c.w_layer;
  g.b_k0 = c.w_k0;
  g.dbg = g_timing_buffer;
  g.epi = c.epi;
  // output tensor maps of the epilogues that store through TMA (bf16 mode only; the split-bf16 mode keeps direct stores)
  if (!c.epi.i[11]) {
    const void* outs[3] = {nullptr, nullptr, nullptr};
    int ldo = 0;
    if (epi == EPI_GATE) { outs[0] = c.epi.ptr[0]; outs[1] = c.epi.ptr[1]; outs[2] = c.epi.ptr[2]; ldo = c.epi.i[0]; }
    else if (epi == EPI_RES) { outs[0] = c.epi.ptr[1]; outs[1] = c.epi.ptr[2]; ldo = BN; }
    else if (epi == EPI_GATE_BWD) { outs[0] = c.epi.ptr[2]; ldo = 2 * c.epi.i[0]; }
    else if (epi == EPI_DX) { outs[0] = c.epi.ptr[1]; ldo = BN; }
    for (int i = 0; i < 3; ++i)
      if (outs[i]) {
        rc = encode_out_map(&g.omap[i], outs[i], ldo, c.T, c.B);
        if (rc) return rc;
      }
  }
  if (c.ksplit > 1) {
    T2_REQUIRE(epi == EPI_TOUT && c.ksplit * kBK <= ktot, T2_ERR_INVALID_ARG,
               "act_gemm: split-K needs an atomically accumulating epilogue and at least one k-block per slice");
  }
#define T2_CASE(E, N) \
  if (epi == E && BN == N) return pair ? launch_pair<E, N>(g, grid, stream) : launch_one<E, N>(g, grid, cs, stream);
  if (epi == EPI_GATE && BN == 256 && c.n_tiles % 2 == 0)
    return pair ? launch_pair<EPI_GATE, 256, 2>(g, grid, stream) : launch_one<EPI_GATE, 256, 2>(g, grid, cs, stream);
  T2_CASE(EPI_GATE, 256)
  T2_CASE(EPI_RES, 128)
  T2_CASE(EPI_RES, 256)
  T2_CASE(EPI_BIAS_ACT, 128)
  T2_CASE(EPI_BIAS_ACT, 256)
  T2_CASE(EPI_CE, 256)
  T2_CASE(EPI_SCALE_RELUMASK, 128)
  T2_CASE(EPI_SCALE_RELUMASK, 256)
  T2_CASE(EPI_GATE_BWD, 128)
  T2_CASE(EPI_GATE_BWD, 256)
  T2_CASE(EPI_DX, 128)
  T2_CASE(EPI_DX, 256)
#undef T2_CASE
#define T2_CASE(E, N) \
  if (epi == E && BN == N) return launch_one<E, N>(g, grid, cs, stream);
  T2_CASE(EPI_MOL, 32)
  T2_CASE(EPI_LSTM, 32)
  T2_CASE(EPI_TOUT, 32)
#undef T2_CASE
  return t2_set_error(T2_ERR_UNSUPPORTED_SHAPE, "act_gemm: no kernel for epilogue %d with BN=%d", epi, BN);
}

int launch_wgrad(const ActT* maps, int nmaps, const WgradTile* tiles_dev, int ntiles, float* out, int T,
                 int B, cudaStream_t stream) {
  T2_REQUIRE(nmaps >= 1 && nmaps <= 6 && ntiles >= 1, T2_ERR_INVALID_ARG, "wgrad: bad map/tile count");
  WgradArgs g;
  memset(&g, 0, sizeof(g));
  for (int i = 0; i < 6; ++i) {
    int rc = encode_act_map(&g.map[i], maps[i < nmaps ? i : 0], kBK);
    if (rc) return rc;
  }
  g.tiles = tiles_dev;
  g.out = out;
  g.T = T;
  g.B = B;
  static bool configured = false;
  if (!configured) {
    T2_CHECK_CUDA(cudaFuncSetAttribute(wgrad_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes));
    configured = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(ntiles); cfg.blockDim = dim3(kGemmThreads); cfg.dynamicSmemBytes = kWgSmemBytes; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  T2_CHECK_CUDA(cudaLaunchKernelEx(&cfg, wgrad_gemm_kernel, g));
  t2_count_launch();
  return T2_OK;
}

}  // namespace t2
