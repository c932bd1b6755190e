// t2_wavenet.cu — WaveNet vocoder (teacher-forced train path) on the tcgen05 GEMM engine.
//
// Replaces wavenet_vocoder/models/wavenet.py:650-721 (step), :476-519 (add_loss) and the layers in
// wavenet_vocoder/models/modules.py / mixture.py of the reference. HBM data layout:
//   activations  bf16 channels-last [layer][B][T][channels]  (rows are GEMM-M, channels are GEMM-K / N)
//   parameters   fp32 masters in TensorFlow variable layouts, concatenated (drop-in checkpoint order)
//   packed       bf16 K-major GEMM operand copies of the masters, refreshed after every optimizer step
// Per layer the forward is two GEMM launches:
//   gate : [x(t-2d) | x(t-d) | x(t) | c(t)] (K = 3R + 128) x Wg -> tanh*sigmoid epilogue -> z (+ stashes)
//   out  : z (K = G/2) x Wo -> (o + b + x) * sqrt(.5) epilogue -> x_next
// the skip 1x1 of ALL layers is deferred into one K = L*G/2 GEMM (skips never round-trip through HBM).
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/t2b200.h"
#include "t2_common.cuh"
#include "t2_gemm.h"

namespace t2 {
namespace {

typedef __nv_bfloat16 bf16;

struct PackJob {
  long long src_off;  // fp32 element offset in params ([K][N] row-major == TF [in][out])
  int K, N;
  long long dst_off;  // bf16 element offset in packed
  int dst_ld;
  int transpose;      // 1: dst[rowmap(n)][col0 + k] ; 0: dst[k][col0 + n]
  int col0;
  float scale;
  int perm_gh;        // >0: gate row permutation with this G/2
};
struct ColsumJob {
  long long src_off;  // byte offset in workspace of a bf16 [rows][ld] matrix
  long long rows;
  int C, ld;
  long long dst_off, dst2_off;  // fp32 element offsets in grads (dst2 < 0: none)
  float scale;
  int div_scalar;     // index into ws scalars to divide by (or -1)
};

inline long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

struct ParamT {
  std::string name;
  long long off;
  int ndim;
  int shape[4];
};

struct Layout {
  t2_wn_config_t c;
  int L, R, G, Gh, S, C, O, Q, B, T, Tc, Kg, ldo, Op;
  bool scalar_in, mol;
  float res_scale;
  std::vector<float> skip_scale;
  // params
  std::vector<ParamT> params;
  long long n_params;
  long long p_in_k, p_in_b, p_f1_k, p_f1_b, p_f2_k, p_f2_b;
  std::vector<long long> p_dil_k, p_dil_b, p_c_k, p_c_b, p_s_k, p_s_b, p_o_k, p_o_b, p_up_k, p_up_b;
  // packed (byte offsets)
  long long k_Wg, k_Wo, k_Ws, k_Wf1, k_Wf2, k_WozT, k_WdT, k_WcT, k_Wf1T, k_Wf2T, k_bias_g, k_bias_skip;
  long long packed_bytes;
  // workspace (byte offsets)
  long long w_cup, w_x, w_xd, w_ta, w_sb, w_z, w_h1, w_h2, w_dlog, w_dh2, w_dskip, w_dxin, w_dg, w_dcup;
  long long w_upgrad[2], w_scalars, w_tiles_main, w_tiles_head, w_packjobs, w_colsum, w_tables;
  std::vector<long long> w_upout;
  std::vector<int> up_w;  // width after each upsample layer
  long long workspace_bytes;
  int n_tiles_main, n_tiles_head, n_packjobs, n_colsum;
  std::vector<PackJob> packjobs;
  std::vector<ColsumJob> colsums;
  std::vector<WgradTile> tiles_main, tiles_head;
  int dil(int l) const { return 1 << (l % (L / c.stacks)); }
};

long long add_param(Layout& lo, const std::string& name, std::initializer_list<int> shape) {
  ParamT p;
  p.name = name;
  p.off = lo.n_params;
  p.ndim = int(shape.size());
  long long n = 1;
  int i = 0;
  for (int s : shape) { p.shape[i++] = s; n *= s; }
  for (; i < 4; ++i) p.shape[i] = 1;
  lo.n_params += align_up(n, 4);  // keep every tensor 16-byte aligned inside the flat buffer
  lo.params.push_back(p);
  return p.off;
}

int build_layout(const t2_wn_config_t* cfg, Layout& lo) {
  T2_REQUIRE(cfg != nullptr, T2_ERR_INVALID_ARG, "null config");
  lo.c = *cfg;
  lo.L = cfg->layers; lo.R = cfg->residual_channels; lo.G = cfg->gate_channels; lo.Gh = lo.G / 2;
  lo.S = cfg->skip_out_channels; lo.C = cfg->cin_channels; lo.O = cfg->out_channels;
  lo.Q = cfg->quantize_channels; lo.B = cfg->B; lo.T = cfg->T; lo.Tc = cfg->Tc;
  lo.scalar_in = cfg->input_type != 2;
  lo.mol = lo.scalar_in;
  T2_REQUIRE(lo.L >= 1 && cfg->stacks >= 1 && lo.L % cfg->stacks == 0, T2_ERR_INVALID_ARG, "layers %% stacks != 0");
  T2_REQUIRE(cfg->kernel_size == 3, T2_ERR_UNSUPPORTED_SHAPE, "kernel_size must be 3");
  T2_REQUIRE(lo.R == 128 || lo.R == 256, T2_ERR_UNSUPPORTED_SHAPE, "residual_channels must be 128 or 256 (got %d)", lo.R);
  T2_REQUIRE(lo.S == 128 || lo.S == 256, T2_ERR_UNSUPPORTED_SHAPE, "skip_out_channels must be 128 or 256 (got %d)", lo.S);
  T2_REQUIRE(lo.Gh == 128 || lo.Gh == 256, T2_ERR_UNSUPPORTED_SHAPE, "gate_channels must be 256 or 512 (got %d)", lo.G);
  T2_REQUIRE(lo.C == 0 || (lo.C % 8 == 0 && lo.C <= 128), T2_ERR_UNSUPPORTED_SHAPE, "cin_channels must be 0 or a multiple of 8 <= 128");
  if (lo.mol) {
    T2_REQUIRE(lo.O % 3 == 0 && lo.O <= 30, T2_ERR_UNSUPPORTED_SHAPE, "scalar input needs a MoL head with <= 10 mixtures (out_channels=%d)", lo.O);
  } else {
    T2_REQUIRE(lo.O == 256 && lo.Q == 256, T2_ERR_UNSUPPORTED_SHAPE, "mulaw-quantize needs out_channels == quantize_channels == 256");
  }
  T2_REQUIRE(lo.B >= 1 && lo.T >= 1, T2_ERR_INVALID_ARG, "bad B/T");
  lo.Kg = 3 * lo.R + (lo.C > 0 ? 128 : 0);
  lo.ldo = lo.mol ? 64 : 512;   // row pitch of dlog (bf16): MoL 32 values (+pad so a 64-wide TMA box fits); CE hi|lo pair
  lo.Op = lo.mol ? 64 : 512;    // K of Wf2T (CE: [Wf2^T | Wf2^T] against the hi|lo split of dlog)
  lo.res_scale = cfg->residual_legacy ? float(sqrt(0.5)) : 1.f;
  lo.skip_scale.resize(lo.L);
  for (int l = 0; l < lo.L; ++l) {
    int e = cfg->legacy ? (l == 0 ? lo.L - 1 : lo.L - l) : 0;
    lo.skip_scale[l] = float(pow(sqrt(0.5), e));
  }
  // upsample widths
  lo.up_w.clear();
  if (lo.C > 0 && !cfg->c_pre_upsampled) {
    T2_REQUIRE(cfg->n_upsample >= 1 && cfg->n_upsample <= 4, T2_ERR_INVALID_ARG, "n_upsample out of range");
    T2_REQUIRE(cfg->freq_axis_kernel_size == 3, T2_ERR_UNSUPPORTED_SHAPE, "freq_axis_kernel_size must be 3");
    int w = lo.Tc;
    for (int i = 0; i < cfg->n_upsample; ++i) { w *= cfg->upsample_scales[i]; lo.up_w.push_back(w); }
    T2_REQUIRE(w == lo.T, T2_ERR_INVALID_ARG, "Tc * prod(upsample_scales) = %d != T = %d", w, lo.T);
  }
  // ---- parameters (order == oracle/wavenet.py:param_shapes) ----
  lo.n_params = 0;
  lo.params.clear();
  const int cin = lo.scalar_in ? 1 : lo.Q;
  lo.p_in_k = add_param(lo, "input_convolution/kernel", {1, cin, lo.R});
  lo.p_in_b = add_param(lo, "input_convolution/bias", {lo.R});
  for (int l = 0; l < lo.L; ++l) {
    char p[64];
    snprintf(p, sizeof(p), "ResidualConv1DGLU_%d/", l);
    std::string s(p);
    lo.p_dil_k.push_back(add_param(lo, s + "residual_block_causal_conv/kernel", {3, lo.R, lo.G}));
    lo.p_dil_b.push_back(add_param(lo, s + "residual_block_causal_conv/bias", {lo.G}));
    if (lo.C > 0) {
      lo.p_c_k.push_back(add_param(lo, s + "residual_block_cin_conv/kernel", {1, lo.C, lo.G}));
      lo.p_c_b.push_back(add_param(lo, s + "residual_block_cin_conv/bias", {lo.G}));
    }
    lo.p_s_k.push_back(add_param(lo, s + "residual_block_skip_conv/kernel", {1, lo.Gh, lo.S}));
    lo.p_s_b.push_back(add_param(lo, s + "residual_block_skip_conv/bias", {lo.S}));
    lo.p_o_k.push_back(add_param(lo, s + "residual_block_out_conv/kernel", {1, lo.Gh, lo.R}));
    lo.p_o_b.push_back(add_param(lo, s + "residual_block_out_conv/bias", {lo.R}));
  }
  lo.p_f1_k = add_param(lo, "final_convolution_1/kernel", {1, lo.S, lo.S});
  lo.p_f1_b = add_param(lo, "final_convolution_1/bias", {lo.S});
  lo.p_f2_k = add_param(lo, "final_convolution_2/kernel", {1, lo.S, lo.O});
  lo.p_f2_b = add_param(lo, "final_convolution_2/bias", {lo.O});
  for (size_t i = 0; i < lo.up_w.size(); ++i) {
    char p[64];
    snprintf(p, sizeof(p), "local_conditioning_upsampling_%d/", int(i) + 1);
    std::string s(p);
    const int sc = cfg->upsample_scales[i];
    if (cfg->upsample_type == 0) {
      lo.p_up_k.push_back(add_param(lo, s + "kernel", {3, 3, 1, sc}));
      lo.p_up_b.push_back(add_param(lo, s + "bias", {sc}));
    } else {
      lo.p_up_k.push_back(add_param(lo, s + "kernel", {3, sc, 1, 1}));
      lo.p_up_b.push_back(add_param(lo, s + "bias", {1}));
    }
  }
  // ---- packed ----
  long long o = 0;
  auto takeb = [&](long long bytes) { long long r = o; o = align_up(o + bytes, 256); return r; };
  const long long L = lo.L;
  lo.k_Wg = takeb(L * lo.G * lo.Kg * 2);
  lo.k_Wo = takeb(L * lo.R * lo.Gh * 2);
  lo.k_Ws = takeb((long long)lo.S * L * lo.Gh * 2);
  lo.k_Wf1 = takeb((long long)lo.S * lo.S * 2);
  lo.k_Wf2 = takeb((long long)lo.O * lo.S * 2);
  lo.k_WozT = takeb(L * lo.Gh * (lo.R + lo.S) * 2);
  lo.k_WdT = takeb(L * lo.R * 3 * lo.G * 2);
  lo.k_WcT = takeb((long long)(lo.C > 0 ? lo.C : 8) * L * lo.G * 2);
  lo.k_Wf1T = takeb((long long)lo.S * lo.S * 2);
  lo.k_Wf2T = takeb((long long)lo.S * lo.Op * 2);
  lo.k_bias_g = takeb(L * lo.G * 4);
  lo.k_bias_skip = takeb(lo.S * 4);
  lo.packed_bytes = o;
  // ---- workspace ----
  o = 0;
  const long long BT = (long long)lo.B * lo.T;
  lo.w_cup = takeb(BT * (lo.C > 0 ? lo.C : 8) * 2);
  lo.w_upout.clear();
  for (size_t i = 0; i < lo.up_w.size(); ++i) lo.w_upout.push_back(takeb((long long)lo.B * lo.C * lo.up_w[i] * 4));
  lo.w_upgrad[0] = takeb(BT * (lo.C > 0 ? lo.C : 8) * 4);
  lo.w_upgrad[1] = takeb(BT * (lo.C > 0 ? lo.C : 8) * 4);
  lo.w_x = takeb(L * BT * lo.R * 2);
  lo.w_xd = cfg->dropout > 0.f ? takeb(L * BT * lo.R * 2) : lo.w_x;
  lo.w_ta = takeb(L * BT * lo.Gh * 2);
  lo.w_sb = takeb(L * BT * lo.Gh * 2);
  lo.w_z = takeb(L * BT * lo.Gh * 2);
  lo.w_h1 = takeb(BT * lo.S * 2);
  lo.w_h2 = takeb(BT * lo.S * 2);
  lo.w_dlog = takeb(BT * lo.ldo * 2);
  lo.w_dh2 = takeb(BT * lo.S * 2);
  lo.w_dskip = takeb(BT * lo.S * 2);
  lo.w_dxin = takeb(L * BT * lo.R * 2);
  lo.w_dg = takeb(L * BT * lo.G * 2);
  lo.w_dcup = takeb(BT * (lo.C > 0 ? lo.C : 8) * 4);
  lo.w_scalars = takeb(64 * 4);

  // ---- pack jobs ----
  lo.packjobs.clear();
  auto pj = [&](long long src, int K, int N, long long dst_bytes, int ld, int transpose, int col0, float scale, int perm) {
    PackJob j; j.src_off = src; j.K = K; j.N = N; j.dst_off = dst_bytes / 2; j.dst_ld = ld; j.transpose = transpose;
    j.col0 = col0; j.scale = scale; j.perm_gh = perm; lo.packjobs.push_back(j);
  };
  for (int l = 0; l < lo.L; ++l) {
    const long long wg = lo.k_Wg + (long long)l * lo.G * lo.Kg * 2;
    for (int j = 0; j < 3; ++j) pj(lo.p_dil_k[l] + (long long)j * lo.R * lo.G, lo.R, lo.G, wg, lo.Kg, 1, j * lo.R, 1.f, lo.Gh);
    if (lo.C > 0) pj(lo.p_c_k[l], lo.C, lo.G, wg, lo.Kg, 1, 3 * lo.R, 1.f, lo.Gh);
    pj(lo.p_o_k[l], lo.Gh, lo.R, lo.k_Wo + (long long)l * lo.R * lo.Gh * 2, lo.Gh, 1, 0, 1.f, 0);
    pj(lo.p_s_k[l], lo.Gh, lo.S, lo.k_Ws, lo.L * lo.Gh, 1, l * lo.Gh, lo.skip_scale[l], 0);
    const long long woz = lo.k_WozT + (long long)l * lo.Gh * (lo.R + lo.S) * 2;
    pj(lo.p_o_k[l], lo.Gh, lo.R, woz, lo.R + lo.S, 0, 0, lo.res_scale, 0);
    pj(lo.p_s_k[l], lo.Gh, lo.S, woz, lo.R + lo.S, 0, lo.R, lo.skip_scale[l], 0);
    const long long wd = lo.k_WdT + (long long)l * lo.R * 3 * lo.G * 2;
    for (int j = 0; j < 3; ++j) pj(lo.p_dil_k[l] + (long long)j * lo.R * lo.G, lo.R, lo.G, wd, 3 * lo.G, 0, j * lo.G, 1.f, 0);
    if (lo.C > 0) pj(lo.p_c_k[l], lo.C, lo.G, lo.k_WcT, lo.L * lo.G, 0, l * lo.G, 1.f, 0);
  }
  pj(lo.p_f1_k, lo.S, lo.S, lo.k_Wf1, lo.S, 1, 0, 1.f, 0);
  pj(lo.p_f2_k, lo.S, lo.O, lo.k_Wf2, lo.S, 1, 0, 1.f, 0);
  pj(lo.p_f1_k, lo.S, lo.S, lo.k_Wf1T, lo.S, 0, 0, 1.f, 0);
  pj(lo.p_f2_k, lo.S, lo.O, lo.k_Wf2T, lo.Op, 0, 0, 1.f, 0);
  if (!lo.mol) pj(lo.p_f2_k, lo.S, lo.O, lo.k_Wf2T, lo.Op, 0, 256, 1.f, 0);
  lo.n_packjobs = int(lo.packjobs.size());

  // ---- wgrad tiles ----
  // main maps: 0 xd_all, 1 dg_all, 2 c_up, 3 z_all, 4 dxin_all, 5 dskip
  lo.tiles_main.clear();
  auto wt = [&](std::vector<WgradTile>& v, int am, int ach, int ash, int al, int bm, int bch, int bl, long long off,
                int ldc, int mv, int nv, float scale, const float* div) {
    WgradTile t; memset(&t, 0, sizeof(t));
    t.a_map = am; t.a_ch0 = ach; t.a_shift = ash; t.a_layer = al; t.b_map = bm; t.b_ch0 = bch; t.b_shift = 0; t.b_layer = bl;
    t.out_off = off; t.ldc = ldc; t.m_valid = mv; t.n_valid = nv; t.scale = scale; t.accumulate = 0; t.div = div;
    v.push_back(t);
  };
  for (int l = 0; l < lo.L; ++l) {
    const int d = lo.dil(l);
    for (int j = 0; j < 3; ++j)
      for (int m0 = 0; m0 < lo.R; m0 += 128)
        for (int n0 = 0; n0 < lo.G; n0 += 128)
          wt(lo.tiles_main, 0, m0, -(2 - j) * d, l, 1, n0, l, lo.p_dil_k[l] + (long long)j * lo.R * lo.G + (long long)m0 * lo.G + n0,
             lo.G, 128, 128, 1.f, nullptr);
    if (lo.C > 0)
      for (int n0 = 0; n0 < lo.G; n0 += 128)
        wt(lo.tiles_main, 2, 0, 0, 0, 1, n0, l, lo.p_c_k[l] + n0, lo.G, lo.C, 128, 1.f, nullptr);
    for (int m0 = 0; m0 < lo.Gh; m0 += 128) {
      if (l < lo.L - 1)
        for (int n0 = 0; n0 < lo.R; n0 += 128)
          wt(lo.tiles_main, 3, m0, 0, l, 4, n0, l + 1, lo.p_o_k[l] + (long long)m0 * lo.R + n0, lo.R, 128, 128, lo.res_scale, nullptr);
      for (int n0 = 0; n0 < lo.S; n0 += 128)
        wt(lo.tiles_main, 3, m0, 0, l, 5, n0, 0, lo.p_s_k[l] + (long long)m0 * lo.S + n0, lo.S, 128, 128, lo.skip_scale[l], nullptr);
    }
  }
  // head maps: 0 h1, 1 dh2, 2 h2, 3 dlog
  lo.tiles_head.clear();
  for (int m0 = 0; m0 < lo.S; m0 += 128) {
    for (int n0 = 0; n0 < lo.S; n0 += 128)
      wt(lo.tiles_head, 0, m0, 0, 0, 1, n0, 0, lo.p_f1_k + (long long)m0 * lo.S + n0, lo.S, 128, 128, 1.f, nullptr);
    for (int n0 = 0; n0 < lo.O; n0 += 128)
      for (int part = 0; part < (lo.mol ? 1 : 2); ++part) {  // CE: hi and lo halves of dlog both accumulate (atomics)
        wt(lo.tiles_head, 2, m0, 0, 0, 3, part * 256 + n0, 0, lo.p_f2_k + (long long)m0 * lo.O + n0, lo.O, 128,
           lo.O - n0 < 128 ? lo.O - n0 : 128, 1.f, reinterpret_cast<const float*>(1) /* patched to scalars[1] at init */);
        lo.tiles_head.back().accumulate = 2;
      }
  }
  lo.n_tiles_main = int(lo.tiles_main.size());
  lo.n_tiles_head = int(lo.tiles_head.size());

  // ---- column-sum (bias gradient) jobs ----
  lo.colsums.clear();
  auto cs = [&](long long src, long long rows, int C, int ld, long long dst, long long dst2, float scale, int div) {
    ColsumJob j; j.src_off = src; j.rows = rows; j.C = C; j.ld = ld; j.dst_off = dst; j.dst2_off = dst2; j.scale = scale;
    j.div_scalar = div; lo.colsums.push_back(j);
  };
  for (int l = 0; l < lo.L; ++l) {
    cs(lo.w_dg + (long long)l * BT * lo.G * 2, BT, lo.G, lo.G, lo.p_dil_b[l], lo.C > 0 ? lo.p_c_b[l] : -1, 1.f, -1);
    if (l < lo.L - 1) cs(lo.w_dxin + (long long)(l + 1) * BT * lo.R * 2, BT, lo.R, lo.R, lo.p_o_b[l], -1, lo.res_scale, -1);
    cs(lo.w_dskip, BT, lo.S, lo.S, lo.p_s_b[l], -1, lo.skip_scale[l], -1);
  }
  cs(lo.w_dh2, BT, lo.S, lo.S, lo.p_f1_b, -1, 1.f, -1);
  cs(lo.w_dlog, BT, lo.O, lo.ldo, lo.p_f2_b, -1, 1.f, 1);
  if (!lo.mol) cs(lo.w_dlog + 256 * 2, BT, lo.O, lo.ldo, lo.p_f2_b, -1, 1.f, 1);
  cs(lo.w_dxin, BT, lo.R, lo.R, lo.p_in_b, -1, 1.f, -1);
  lo.n_colsum = int(lo.colsums.size());

  lo.w_tiles_main = takeb((long long)lo.n_tiles_main * sizeof(WgradTile));
  lo.w_tiles_head = takeb((long long)lo.n_tiles_head * sizeof(WgradTile));
  lo.w_packjobs = takeb((long long)lo.n_packjobs * sizeof(PackJob));
  lo.w_colsum = takeb((long long)lo.n_colsum * sizeof(ColsumJob));
  lo.w_tables = takeb((long long)lo.L * (3 * sizeof(long long) + sizeof(float)));
  lo.workspace_bytes = o;
  return T2_OK;
}

// ------------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------------
__global__ void pack_kernel(const float* __restrict__ params, bf16* __restrict__ packed, const PackJob* __restrict__ jobs) {
  const PackJob j = jobs[blockIdx.y];
  const long long n = (long long)j.K * j.N;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int k = int(e / j.N), c = int(e % j.N);
    const float v = params[j.src_off + e] * j.scale;
    long long d;
    if (j.transpose) {
      int row = c;
      if (j.perm_gh > 0) {
        const int half = c / j.perm_gh, idx = c % j.perm_gh;
        row = (idx / 128) * 256 + half * 128 + (idx % 128);
      }
      d = j.dst_off + (long long)row * j.dst_ld + j.col0 + k;
    } else {
      d = j.dst_off + (long long)k * j.dst_ld + j.col0 + c;
    }
    packed[d] = __float2bfloat16(v);
  }
}

// bias_g[l][g] = b_dil + b_cin ; bias_skip[s] = sum_l scale_l * b_skip_l[s]
struct DerivedArgs {
  const float* params;
  float* bias_g;
  float* bias_skip;
  const long long* offs;  // [3L]: dil_b, c_b (or -1), s_b per layer
  const float* scales;    // [L]
  int L, G, S;
};
__global__ void derived_bias_kernel(DerivedArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.L * a.G) {
    const int l = i / a.G, g = i % a.G;
    float v = a.params[a.offs[3 * l] + g];
    if (a.offs[3 * l + 1] >= 0) v += a.params[a.offs[3 * l + 1] + g];
    a.bias_g[i] = v;
  }
  if (i < a.S) {
    float v = 0.f;
    for (int l = 0; l < a.L; ++l) v += a.scales[l] * a.params[a.offs[3 * l + 2] + i];
    a.bias_skip[i] = v;
  }
}

// first (embedding) 1x1 conv: one-hot input == row gather (wavenet.py:705; SURVEY §8a "embedding in disguise")
__global__ void first_conv_kernel(const void* __restrict__ xin, int scalar_in, const float* __restrict__ W,
                                  const float* __restrict__ bias, bf16* __restrict__ x, bf16* __restrict__ xd,
                                  long long npos, int R, float p, unsigned long long seed,
                                  const unsigned long long* __restrict__ step) {
  if (step) seed += *step;
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= npos * R) return;
  const long long pos = e / R;
  const int r = int(e % R);
  float v;
  if (scalar_in) v = static_cast<const float*>(xin)[pos] * W[r] + bias[r];
  else v = W[(long long)static_cast<const int*>(xin)[pos] * R + r] + bias[r];
  x[e] = __float2bfloat16(v);
  if (xd != x && xd != nullptr) {
    const float keep_inv = 1.f / (1.f - p);
    xd[e] = __float2bfloat16(hash_uniform32(hash_seed(seed, 0u), (unsigned long long)e) >= p ? v * keep_inv : 0.f);
  }
}
__global__ void first_conv_bwd_kernel(const void* __restrict__ xin, int scalar_in, const bf16* __restrict__ dx0,
                                      float* __restrict__ dW, long long npos, int R) {
  // one block = 64 positions x R channels; scalar input reduces in registers first
  const int r = threadIdx.x;
  const long long p0 = (long long)blockIdx.x * 64;
  if (scalar_in) {
    float acc = 0.f;
    for (int i = 0; i < 64 && p0 + i < npos; ++i)
      acc += static_cast<const float*>(xin)[p0 + i] * __bfloat162float(dx0[(p0 + i) * R + r]);
    atomicAdd(dW + r, acc);
  } else {
    for (int i = 0; i < 64 && p0 + i < npos; ++i) {
      const int idx = static_cast<const int*>(xin)[p0 + i];
      atomicAdd(dW + (long long)idx * R + r, __bfloat162float(dx0[(p0 + i) * R + r]));
    }
  }
}

__global__ void colsum_kernel(const uint8_t* __restrict__ ws, float* __restrict__ grads, const ColsumJob* __restrict__ jobs,
                              const float* __restrict__ scalars) {
  const ColsumJob j = jobs[blockIdx.y];
  const bf16* src = reinterpret_cast<const bf16*>(ws + j.src_off);
  const long long rows_per = (j.rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * rows_per;
  const long long r1 = r0 + rows_per < j.rows ? r0 + rows_per : j.rows;
  float sc = j.scale;
  if (j.div_scalar >= 0) sc /= fmaxf(scalars[j.div_scalar], 1e-20f);
  // one thread = one pair of adjacent columns (C and ld are even); 4 rows in flight per iteration
  for (int c = 2 * threadIdx.x; c < j.C; c += 2 * blockDim.x) {
    float a0 = 0.f, a1 = 0.f;
    long long r = r0;
    for (; r + 4 <= r1; r += 4) {
      uint32_t u[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) u[i] = __ldg(reinterpret_cast<const uint32_t*>(src + (r + i) * j.ld + c));
#pragma unroll
      for (int i = 0; i < 4; ++i) { a0 += bf16lo(u[i]); a1 += bf16hi(u[i]); }
    }
    for (; r < r1; ++r) {
      const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(src + r * j.ld + c));
      a0 += bf16lo(u); a1 += bf16hi(u);
    }
    a0 *= sc; a1 *= sc;
    atomicAdd(grads + j.dst_off + c, a0);
    if (c + 1 < j.C) atomicAdd(grads + j.dst_off + c + 1, a1);
    if (j.dst2_off >= 0) {
      atomicAdd(grads + j.dst2_off + c, a0);
      if (c + 1 < j.C) atomicAdd(grads + j.dst2_off + c + 1, a1);
    }
  }
}

// ---- conditioning upsampling net (modules.py:539-654 SubPixel, :736-770 ConvTranspose2D) + ReLU -------------
// in [B][H][W] fp32 -> out [B][H][W*s] fp32 (post-ReLU); optional bf16 channels-last copy [B][W*s][H]
__global__ void upsample_fwd_kernel(const float* __restrict__ in, const float* __restrict__ K, const float* __restrict__ bias,
                                    float* __restrict__ out, bf16* __restrict__ out_cl, int B, int H, int W, int s, int type) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long n = (long long)B * H * W * s;
  if (e >= n) return;
  const int Wo = W * s;
  const int xo = int(e % Wo);
  const int h = int((e / Wo) % H);
  const int b = int(e / ((long long)Wo * H));
  const int w = xo / s, k = xo % s;
  const float* ib = in + (long long)b * H * W;
  float acc;
  if (type == 0) {  // SubPixel: 3x3 'same' conv, 1 -> s channels, then periodic shuffle. K [3][3][1][s]
    acc = bias[k];
#pragma unroll
    for (int dh = 0; dh < 3; ++dh) {
      const int hh = h + dh - 1;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int dw = 0; dw < 3; ++dw) {
        const int ww = w + dw - 1;
        if (ww < 0 || ww >= W) continue;
        acc += K[(dh * 3 + dw) * s + k] * ib[hh * W + ww];
      }
    }
  } else {  // Conv2DTranspose kernel (3, s), strides (1, s), 'same': out[h, w*s+k] = sum_q in[h+1-q, w] K[q][k]
    acc = bias[0];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int hh = h + 1 - q;
      if (hh < 0 || hh >= H) continue;
      acc += K[q * s + k] * ib[hh * W + w];
    }
  }
  acc = fmaxf(acc, 0.f);
  out[e] = acc;
  if (out_cl) out_cl[((long long)b * Wo + xo) * H + h] = __float2bfloat16(acc);
}
// d_pre = d_out * (out > 0); accumulates dK, dbias (atomics after a block reduce) and writes d_in.
// d_out is either fp32 [B][H][Wo] (cl = 0) or channels-last fp32 [B][Wo][H] (cl = 1)
__global__ void upsample_bwd_param_kernel(const float* __restrict__ in, const float* __restrict__ out, const float* __restrict__ dout,
                                          int cl, float* __restrict__ dK, float* __restrict__ dbias, int B, int H, int W, int s,
                                          int type) {
  // one block per (k, tap): reduces over all (b, h, w)
  const int k = blockIdx.x;
  const int tap = blockIdx.y;  // SubPixel: 0..8 taps, 9 = bias ; 2D: 0..2 taps, 3 = bias
  const int ntap = type == 0 ? 9 : 3;
  const int Wo = W * s;
  float acc = 0.f;
  const long long n = (long long)B * H * W;
  for (long long e = blockIdx.z * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.z * blockDim.x) {
    const int w = int(e % W), h = int((e / W) % H), b = int(e / ((long long)W * H));
    const int xo = w * s + k;
    const float o = out[((long long)b * H + h) * Wo + xo];
    if (o <= 0.f) continue;
    const float g = cl ? dout[((long long)b * Wo + xo) * H + h] : dout[((long long)b * H + h) * Wo + xo];
    if (tap == ntap) { acc += g; continue; }
    int hh, ww;
    if (type == 0) { hh = h + tap / 3 - 1; ww = w + tap % 3 - 1; }
    else { hh = h + 1 - tap; ww = w; }
    if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
    acc += g * in[((long long)b * H + hh) * W + ww];
  }
  __shared__ float red[32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) {
      if (tap == ntap) {
        if (type == 0) atomicAdd(dbias + k, v); else atomicAdd(dbias, v);
      } else {
        atomicAdd(dK + tap * s + k, v);
      }
    }
  }
}
__global__ void upsample_bwd_input_kernel(const float* __restrict__ out, const float* __restrict__ dout, int cl,
                                          const float* __restrict__ K, float* __restrict__ din, int B, int H, int W, int s, int type) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long n = (long long)B * H * W;
  if (e >= n) return;
  const int w = int(e % W), h = int((e / W) % H), b = int(e / ((long long)W * H));
  const int Wo = W * s;
  float acc = 0.f;
  auto dpre = [&](int hh, int xo) -> float {
    const float o = out[((long long)b * H + hh) * Wo + xo];
    if (o <= 0.f) return 0.f;
    return cl ? dout[((long long)b * Wo + xo) * H + hh] : dout[((long long)b * H + hh) * Wo + xo];
  };
  if (type == 0) {
    for (int dh = 0; dh < 3; ++dh) {
      const int ho = h - dh + 1;
      if (ho < 0 || ho >= H) continue;
      for (int dw = 0; dw < 3; ++dw) {
        const int wo = w - dw + 1;
        if (wo < 0 || wo >= W) continue;
        for (int k = 0; k < s; ++k) acc += dpre(ho, wo * s + k) * K[(dh * 3 + dw) * s + k];
      }
    }
  } else {
    for (int q = 0; q < 3; ++q) {
      const int ho = h - 1 + q;
      if (ho < 0 || ho >= H) continue;
      for (int k = 0; k < s; ++k) acc += dpre(ho, w * s + k) * K[q * s + k];
    }
  }
  din[e] = acc;
}
__global__ void f32_to_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long n) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < n) out[e] = __float2bfloat16(in[e]);
}

// the per-layer gate GEMM: [x(t-2d) | x(t-d) | x(t) | c(t)] x Wg with the tanh*sigmoid epilogue
ActGemmCall make_gate_call(const Layout& lo, uint8_t* ws, const uint8_t* pk, int l, bool save) {
  const long long BT = (long long)lo.B * lo.T;
  const int d = lo.dil(l);
  ActGemmCall g;
  memset(&g, 0, sizeof(g));
  g.a[0] = make_act(ws + lo.w_xd, lo.R, lo.T, lo.B, lo.L);
  g.a[1] = make_act(ws + lo.w_cup, lo.C > 0 ? lo.C : 8, lo.T, lo.B, 1);
  g.na = lo.C > 0 ? 2 : 1;
  g.seg[0] = Seg{0, -2 * d, 0, lo.R / kBK, l, 1};
  g.seg[1] = Seg{0, -d, 0, lo.R / kBK, l, 1};
  g.seg[2] = Seg{0, 0, 0, lo.R / kBK, l, 1};
  g.nseg = 3;
  if (lo.C > 0) { g.seg[3] = Seg{1, 0, 0, 2, 0, 1}; g.nseg = 4; }
  g.w = pk + lo.k_Wg; g.wN = lo.G; g.wK = lo.Kg; g.wL = lo.L; g.w_layer = l; g.w_k0 = 0;
  g.T = lo.T; g.B = lo.B; g.n_tiles = lo.G / 256;
  const long long lofs = (long long)l * BT * lo.Gh;
  g.epi.ptr[0] = save ? reinterpret_cast<bf16*>(ws + lo.w_ta) + lofs : nullptr;
  g.epi.ptr[1] = save ? reinterpret_cast<bf16*>(ws + lo.w_sb) + lofs : nullptr;
  g.epi.ptr[2] = reinterpret_cast<bf16*>(ws + lo.w_z) + lofs;
  g.epi.ptr[3] = const_cast<float*>(reinterpret_cast<const float*>(pk + lo.k_bias_g) + (long long)l * lo.G);
  g.epi.i[0] = lo.Gh;
  return g;
}

ActGemmCall make_out_call(const Layout& lo, uint8_t* ws, const uint8_t* pk, const float* params, int l, float p,
                          unsigned long long seed, const unsigned long long* d_step) {
  const long long BT = (long long)lo.B * lo.T;
  bf16* x_all = reinterpret_cast<bf16*>(ws + lo.w_x);
  bf16* xd_all = reinterpret_cast<bf16*>(ws + lo.w_xd);
  ActGemmCall o;
  memset(&o, 0, sizeof(o));
  o.a[0] = make_act(ws + lo.w_z, lo.Gh, lo.T, lo.B, lo.L); o.na = 1;
  o.seg[0] = Seg{0, 0, 0, lo.Gh / kBK, l, 1}; o.nseg = 1;
  o.w = pk + lo.k_Wo; o.wN = lo.R; o.wK = lo.Gh; o.wL = lo.L; o.w_layer = l;
  o.T = lo.T; o.B = lo.B; o.n_tiles = 1;
  o.epi.ptr[0] = x_all + (long long)l * BT * lo.R;
  o.epi.ptr[1] = x_all + (long long)(l + 1) * BT * lo.R;
  o.epi.ptr[2] = p > 0.f ? xd_all + (long long)(l + 1) * BT * lo.R : nullptr;
  o.epi.ptr[3] = const_cast<float*>(params + lo.p_o_b[l]);
  o.epi.f[0] = lo.res_scale; o.epi.f[1] = p; o.epi.i[1] = l + 1; o.epi.seed = seed;
  o.epi.ptr[7] = const_cast<unsigned long long*>(d_step);
  return o;
}

ActGemmCall make_dz_call(const Layout& lo, uint8_t* ws, const uint8_t* pk, int l) {
  const long long BT = (long long)lo.B * lo.T;
  const bool top = l == lo.L - 1;
  ActGemmCall g;
  memset(&g, 0, sizeof(g));
  g.a[0] = make_act(ws + lo.w_dxin, lo.R, lo.T, lo.B, lo.L);
  g.a[1] = make_act(ws + lo.w_dskip, lo.S, lo.T, lo.B, 1);
  g.na = 2;
  if (top) {
    g.seg[0] = Seg{1, 0, 0, lo.S / kBK, 0, 1}; g.nseg = 1; g.w_k0 = lo.R;
  } else {
    g.seg[0] = Seg{0, 0, 0, lo.R / kBK, l + 1, 1};
    g.seg[1] = Seg{1, 0, 0, lo.S / kBK, 0, 1};
    g.nseg = 2; g.w_k0 = 0;
  }
  const int bn_z = lo.Gh >= 256 ? 256 : 128;
  g.w = pk + lo.k_WozT; g.wN = lo.Gh; g.wK = lo.R + lo.S; g.wL = lo.L; g.w_layer = l;
  g.T = lo.T; g.B = lo.B; g.n_tiles = lo.Gh / bn_z;
  const long long lofs = (long long)l * BT * lo.Gh;
  g.epi.ptr[0] = reinterpret_cast<bf16*>(ws + lo.w_ta) + lofs;
  g.epi.ptr[1] = reinterpret_cast<bf16*>(ws + lo.w_sb) + lofs;
  g.epi.ptr[2] = reinterpret_cast<bf16*>(ws + lo.w_dg) + (long long)l * BT * lo.G;
  g.epi.i[0] = lo.Gh;
  return g;
}

ActGemmCall make_dx_call(const Layout& lo, uint8_t* ws, const uint8_t* pk, int l, float p, unsigned long long seed,
                         const unsigned long long* d_step) {
  const long long BT = (long long)lo.B * lo.T;
  const int d = lo.dil(l);
  const bool top = l == lo.L - 1;
  bf16* dxin = reinterpret_cast<bf16*>(ws + lo.w_dxin);
  ActGemmCall g;
  memset(&g, 0, sizeof(g));
  g.a[0] = make_act(ws + lo.w_dg, lo.G, lo.T, lo.B, lo.L); g.na = 1;
  for (int j = 0; j < 3; ++j) g.seg[j] = Seg{0, (2 - j) * d, 0, lo.G / kBK, l, 1};
  g.nseg = 3;
  g.w = pk + lo.k_WdT; g.wN = lo.R; g.wK = 3 * lo.G; g.wL = lo.L; g.w_layer = l;
  g.T = lo.T; g.B = lo.B; g.n_tiles = 1;
  g.epi.ptr[0] = top ? nullptr : dxin + (long long)(l + 1) * BT * lo.R;
  g.epi.ptr[1] = dxin + (long long)l * BT * lo.R;
  g.epi.f[0] = lo.res_scale; g.epi.f[1] = p; g.epi.i[1] = l; g.epi.seed = seed;
  g.epi.ptr[7] = const_cast<unsigned long long*>(d_step);
  return g;
}

inline dim3 grid1d(long long n, int block = 256) { return dim3((unsigned)((n + block - 1) / block)); }

}  // namespace
}  // namespace t2

using namespace t2;

// ----------------------------------------------------------------------------------------------------------
// C-ABI
// ----------------------------------------------------------------------------------------------------------
extern "C" int t2_wn_sizes(const t2_wn_config_t* cfg, t2_wn_sizes_t* out) {
  Layout lo;
  int rc = build_layout(cfg, lo);
  if (rc) return rc;
  T2_REQUIRE(out != nullptr, T2_ERR_INVALID_ARG, "null out");
  out->n_params = lo.n_params;
  out->packed_bytes = lo.packed_bytes;
  out->workspace_bytes = lo.workspace_bytes;
  out->n_tensors = int(lo.params.size());
  return T2_OK;
}

extern "C" int t2_wn_param_info(const t2_wn_config_t* cfg, int i, char* name, int name_cap, long long* offset,
                                int* ndim, int* shape4) {
  Layout lo;
  int rc = build_layout(cfg, lo);
  if (rc) return rc;
  T2_REQUIRE(i >= 0 && i < int(lo.params.size()), T2_ERR_INVALID_ARG, "tensor index %d out of range", i);
  const ParamT& p = lo.params[i];
  snprintf(name, name_cap, "%s", p.name.c_str());
  *offset = p.off;
  *ndim = p.ndim;
  for (int k = 0; k < 4; ++k) shape4[k] = p.shape[k];
  return T2_OK;
}

extern "C" int t2_wn_init(const t2_wn_config_t* cfg, void* d_packed, void* d_workspace, void* stream) {
  Layout lo;
  int rc = build_layout(cfg, lo);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  T2_CHECK_CUDA(cudaMemsetAsync(d_packed, 0, lo.packed_bytes, st));
  T2_CHECK_CUDA(cudaMemsetAsync(d_workspace, 0, lo.workspace_bytes, st));
  float* scalars = reinterpret_cast<float*>(ws + lo.w_scalars);
  for (auto& t : lo.tiles_head)
    if (t.div != nullptr) t.div = scalars + 1;
  T2_CHECK_CUDA(cudaMemcpyAsync(ws + lo.w_tiles_main, lo.tiles_main.data(), lo.tiles_main.size() * sizeof(WgradTile), cudaMemcpyHostToDevice, st));
  T2_CHECK_CUDA(cudaMemcpyAsync(ws + lo.w_tiles_head, lo.tiles_head.data(), lo.tiles_head.size() * sizeof(WgradTile), cudaMemcpyHostToDevice, st));
  T2_CHECK_CUDA(cudaMemcpyAsync(ws + lo.w_packjobs, lo.packjobs.data(), lo.packjobs.size() * sizeof(PackJob), cudaMemcpyHostToDevice, st));
  T2_CHECK_CUDA(cudaMemcpyAsync(ws + lo.w_colsum, lo.colsums.data(), lo.colsums.size() * sizeof(ColsumJob), cudaMemcpyHostToDevice, st));
  std::vector<long long> offs(3 * lo.L);
  for (int l = 0; l < lo.L; ++l) {
    offs[3 * l] = lo.p_dil_b[l];
    offs[3 * l + 1] = lo.C > 0 ? lo.p_c_b[l] : -1;
    offs[3 * l + 2] = lo.p_s_b[l];
  }
  T2_CHECK_CUDA(cudaMemcpyAsync(ws + lo.w_tables, offs.data(), offs.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
  T2_CHECK_CUDA(cudaMemcpyAsync(ws + lo.w_tables + 3 * lo.L * sizeof(long long), lo.skip_scale.data(), lo.L * sizeof(float),
                                cudaMemcpyHostToDevice, st));
  T2_CHECK_CUDA(cudaStreamSynchronize(st));
  return T2_OK;
}

extern "C" int t2_wn_pack_weights(const t2_wn_config_t* cfg, const float* d_params, void* d_packed,
                                  void* d_workspace, void* stream) {
  Layout lo;
  int rc = build_layout(cfg, lo);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  uint8_t* pk = static_cast<uint8_t*>(d_packed);
  pack_kernel<<<dim3(48, lo.n_packjobs), 256, 0, st>>>(d_params, reinterpret_cast<bf16*>(pk),
                                                       reinterpret_cast<const PackJob*>(ws + lo.w_packjobs)); t2_count_launch();
  T2_CHECK_CUDA(cudaGetLastError());
  long long* d_offs = reinterpret_cast<long long*>(ws + lo.w_tables);
  float* d_scales = reinterpret_cast<float*>(ws + lo.w_tables + 3 * lo.L * sizeof(long long));
  DerivedArgs a;
  a.params = d_params; a.bias_g = reinterpret_cast<float*>(pk + lo.k_bias_g); a.bias_skip = reinterpret_cast<float*>(pk + lo.k_bias_skip);
  a.offs = d_offs; a.scales = d_scales; a.L = lo.L; a.G = lo.G; a.S = lo.S;
  derived_bias_kernel<<<grid1d((long long)lo.L * lo.G), 256, 0, st>>>(a); t2_count_launch();
  T2_CHECK_CUDA(cudaGetLastError());
  return T2_OK;
}

extern "C" int t2_wn_forward(const t2_wn_config_t* cfg, const float* d_params, const void* d_packed,
                             void* d_workspace, const void* d_x, const float* d_c, const void* d_targets,
                             const int* d_lengths, float* d_loss, float* d_logits, int save_for_backward,
                             unsigned long long seed, const unsigned long long* d_step, void* stream) {
  Layout lo;
  int rc = build_layout(cfg, lo);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  const uint8_t* pk = static_cast<const uint8_t*>(d_packed);
  const long long BT = (long long)lo.B * lo.T;
  float* scalars = reinterpret_cast<float*>(ws + lo.w_scalars);
  const float p = cfg->dropout;
  T2_CHECK_CUDA(cudaMemsetAsync(scalars, 0, 16 * sizeof(float), st));

  // 1. conditioning -> c_up (bf16 channels-last)
  bf16* c_up = reinterpret_cast<bf16*>(ws + lo.w_cup);
  if (lo.C > 0) {
    T2_REQUIRE(d_c != nullptr, T2_ERR_INVALID_ARG, "local conditioning enabled but d_c is NULL");
    if (cfg->c_pre_upsampled) {
      f32_to_bf16_kernel<<<grid1d(BT * lo.C), 256, 0, st>>>(d_c, c_up, BT * lo.C); t2_count_launch();
    } else {
      const float* in = d_c;
      int W = lo.Tc;
      for (size_t i = 0; i < lo.up_w.size(); ++i) {
        const int s = cfg->upsample_scales[i];
        float* out = reinterpret_cast<float*>(ws + lo.w_upout[i]);
        const bool last = i + 1 == lo.up_w.size();
        upsample_fwd_kernel<<<grid1d((long long)lo.B * lo.C * W * s), 256, 0, st>>>(
            in, d_params + lo.p_up_k[i], d_params + lo.p_up_b[i], out, last ? c_up : nullptr, lo.B, lo.C, W, s, cfg->upsample_type); t2_count_launch();
        in = out;
        W *= s;
      }
    }
    T2_CHECK_CUDA(cudaGetLastError());
  }
  // 2. first conv
  bf16* x_all = reinterpret_cast<bf16*>(ws + lo.w_x);
  bf16* xd_all = reinterpret_cast<bf16*>(ws + lo.w_xd);
  first_conv_kernel<<<grid1d(BT * lo.R), 256, 0, st>>>(d_x, lo.scalar_in ? 1 : 0, d_params + lo.p_in_k, d_params + lo.p_in_b,
                                                       x_all, p > 0.f ? xd_all : x_all, BT, lo.R, p, seed, d_step); t2_count_launch();
  T2_CHECK_CUDA(cudaGetLastError());
  // 3. residual stack
  bf16* z_all = reinterpret_cast<bf16*>(ws + lo.w_z);
  const ActT a_z = make_act(z_all, lo.Gh, lo.T, lo.B, lo.L);
  for (int l = 0; l < lo.L; ++l) {
    ActGemmCall g = make_gate_call(lo, ws, pk, l, save_for_backward != 0);
    rc = launch_act_gemm(EPI_GATE, 256, g, st);
    if (rc) return rc;
    if (l + 1 < lo.L) {
      ActGemmCall o = make_out_call(lo, ws, pk, d_params, l, p, seed, d_step);
      rc = launch_act_gemm(EPI_RES, lo.R, o, st);
      if (rc) return rc;
    }
  }
  // 4. all skip 1x1s as one K = L*Gh GEMM, + ReLU
  bf16* h1 = reinterpret_cast<bf16*>(ws + lo.w_h1);
  bf16* h2 = reinterpret_cast<bf16*>(ws + lo.w_h2);
  {
    ActGemmCall g;
    memset(&g, 0, sizeof(g));
    g.a[0] = a_z; g.na = 1;
    g.seg[0] = Seg{0, 0, 0, lo.Gh / kBK, 0, lo.L}; g.nseg = 1;
    g.w = pk + lo.k_Ws; g.wN = lo.S; g.wK = lo.L * lo.Gh; g.wL = 1;
    g.T = lo.T; g.B = lo.B; g.n_tiles = 1;
    g.epi.ptr[0] = h1; g.epi.ptr[1] = const_cast<float*>(reinterpret_cast<const float*>(pk + lo.k_bias_skip));
    g.epi.i[0] = lo.S; g.epi.i[1] = 1; g.epi.i[2] = lo.S;
    rc = launch_act_gemm(EPI_BIAS_ACT, lo.S, g, st);
    if (rc) return rc;
  }
  {
    ActGemmCall g;
    memset(&g, 0, sizeof(g));
    g.a[0] = make_act(h1, lo.S, lo.T, lo.B); g.na = 1;
    g.seg[0] = Seg{0, 0, 0, lo.S / kBK, 0, 1}; g.nseg = 1;
    g.w = pk + lo.k_Wf1; g.wN = lo.S; g.wK = lo.S; g.wL = 1;
    g.T = lo.T; g.B = lo.B; g.n_tiles = 1;
    g.epi.ptr[0] = h2; g.epi.ptr[1] = const_cast<float*>(d_params + lo.p_f1_b);
    g.epi.i[0] = lo.S; g.epi.i[1] = 1; g.epi.i[2] = lo.S;
    rc = launch_act_gemm(EPI_BIAS_ACT, lo.S, g, st);
    if (rc) return rc;
  }
  // 5. output projection fused with the loss
  {
    ActGemmCall g;
    memset(&g, 0, sizeof(g));
    g.a[0] = make_act(h2, lo.S, lo.T, lo.B); g.na = 1;
    g.seg[0] = Seg{0, 0, 0, lo.S / kBK, 0, 1}; g.nseg = 1;
    g.w = pk + lo.k_Wf2; g.wN = lo.O; g.wK = lo.S; g.wL = 1;
    g.T = lo.T; g.B = lo.B; g.n_tiles = 1;
    g.epi.ptr[0] = const_cast<void*>(d_targets);
    g.epi.ptr[1] = const_cast<int*>(d_lengths);
    g.epi.ptr[2] = const_cast<float*>(d_params + lo.p_f2_b);
    g.epi.ptr[3] = scalars + 0; g.epi.ptr[4] = scalars + 1;
    g.epi.ptr[5] = save_for_backward ? ws + lo.w_dlog : nullptr;
    g.epi.ptr[6] = d_logits;
    g.epi.i[1] = lo.ldo;
    if (lo.mol) {
      g.epi.f[0] = cfg->log_scale_min;
      g.epi.f[1] = 1.f / float(lo.Q - 1);
      g.epi.f[2] = logf(float(lo.Q - 1) / 2.f);
      g.epi.i[0] = lo.O / 3;
      rc = launch_act_gemm(EPI_MOL, 32, g, st);
    } else {
      rc = launch_act_gemm(EPI_CE, 256, g, st);
    }
    if (rc) return rc;
  }
  if (d_loss) T2_CHECK_CUDA(cudaMemcpyAsync(d_loss, scalars, 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return T2_OK;
}

extern "C" int t2_wn_backward(const t2_wn_config_t* cfg, const float* d_params, const void* d_packed,
                              void* d_workspace, const void* d_x, const float* d_c, float* d_grads,
                              unsigned long long seed, const unsigned long long* d_step, void* stream) {
  Layout lo;
  int rc = build_layout(cfg, lo);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  const uint8_t* pk = static_cast<const uint8_t*>(d_packed);
  const long long BT = (long long)lo.B * lo.T;
  float* scalars = reinterpret_cast<float*>(ws + lo.w_scalars);
  const float p = cfg->dropout;
  T2_CHECK_CUDA(cudaMemsetAsync(d_grads, 0, lo.n_params * sizeof(float), st));
  bf16* h1 = reinterpret_cast<bf16*>(ws + lo.w_h1);
  bf16* h2 = reinterpret_cast<bf16*>(ws + lo.w_h2);
  bf16* dlog = reinterpret_cast<bf16*>(ws + lo.w_dlog);
  bf16* dh2 = reinterpret_cast<bf16*>(ws + lo.w_dh2);
  bf16* dskip = reinterpret_cast<bf16*>(ws + lo.w_dskip);
  bf16* dxin = reinterpret_cast<bf16*>(ws + lo.w_dxin);
  bf16* dg = reinterpret_cast<bf16*>(ws + lo.w_dg);
  bf16* ta_all = reinterpret_cast<bf16*>(ws + lo.w_ta);
  bf16* sb_all = reinterpret_cast<bf16*>(ws + lo.w_sb);
  // head: dh2 = (dlog x Wf2^T) * relu'(h2) / count ; dskip = (dh2 x Wf1^T) * relu'(h1)
  {
    ActGemmCall g;
    memset(&g, 0, sizeof(g));
    g.a[0] = make_act(dlog, lo.ldo, lo.T, lo.B); g.na = 1;
    g.seg[0] = Seg{0, 0, 0, lo.Op / kBK, 0, 1}; g.nseg = 1;
    g.w = pk + lo.k_Wf2T; g.wN = lo.S; g.wK = lo.Op; g.wL = 1;
    g.T = lo.T; g.B = lo.B; g.n_tiles = 1;
    g.epi.ptr[0] = dh2; g.epi.ptr[1] = h2; g.epi.ptr[2] = scalars + 1; g.epi.f[0] = 1.f; g.epi.i[0] = lo.S;
    rc = launch_act_gemm(EPI_SCALE_RELUMASK, lo.S, g, st);
    if (rc) return rc;
  }
  {
    ActGemmCall g;
    memset(&g, 0, sizeof(g));
    g.a[0] = make_act(dh2, lo.S, lo.T, lo.B); g.na = 1;
    g.seg[0] = Seg{0, 0, 0, lo.S / kBK, 0, 1}; g.nseg = 1;
    g.w = pk + lo.k_Wf1T; g.wN = lo.S; g.wK = lo.S; g.wL = 1;
    g.T = lo.T; g.B = lo.B; g.n_tiles = 1;
    g.epi.ptr[0] = dskip; g.epi.ptr[1] = h1; g.epi.ptr[2] = nullptr; g.epi.f[0] = 1.f; g.epi.i[0] = lo.S;
    rc = launch_act_gemm(EPI_SCALE_RELUMASK, lo.S, g, st);
    if (rc) return rc;
  }
  // residual stack, top down
  const ActT a_dxin = make_act(dxin, lo.R, lo.T, lo.B, lo.L);
  const ActT a_dskip = make_act(dskip, lo.S, lo.T, lo.B, 1);
  const ActT a_dg = make_act(dg, lo.G, lo.T, lo.B, lo.L);
  const int bn_z = lo.Gh >= 256 ? 256 : 128;
  for (int l = lo.L - 1; l >= 0; --l) {
    ActGemmCall gz = make_dz_call(lo, ws, pk, l);
    rc = launch_act_gemm(EPI_GATE_BWD, bn_z, gz, st);
    if (rc) return rc;
    ActGemmCall gx = make_dx_call(lo, ws, pk, l, p, seed, d_step);
    rc = launch_act_gemm(EPI_DX, lo.R, gx, st);
    if (rc) return rc;
  }
  // weight gradients: one batched launch for the whole stack, one for the head
  {
    ActT maps[6] = {make_act(ws + lo.w_xd, lo.R, lo.T, lo.B, lo.L), a_dg,
                    make_act(ws + lo.w_cup, lo.C > 0 ? lo.C : 8, lo.T, lo.B, 1),
                    make_act(ws + lo.w_z, lo.Gh, lo.T, lo.B, lo.L), a_dxin, a_dskip};
    rc = launch_wgrad(maps, 6, reinterpret_cast<const WgradTile*>(ws + lo.w_tiles_main), lo.n_tiles_main, d_grads, lo.T, lo.B, st);
    if (rc) return rc;
    ActT hmaps[4] = {make_act(h1, lo.S, lo.T, lo.B), make_act(dh2, lo.S, lo.T, lo.B), make_act(h2, lo.S, lo.T, lo.B),
                     make_act(dlog, lo.ldo, lo.T, lo.B)};
    rc = launch_wgrad(hmaps, 4, reinterpret_cast<const WgradTile*>(ws + lo.w_tiles_head), lo.n_tiles_head, d_grads, lo.T, lo.B, st);
    if (rc) return rc;
  }
  // bias gradients
  colsum_kernel<<<dim3(96, lo.n_colsum), 256, 0, st>>>(ws, d_grads, reinterpret_cast<const ColsumJob*>(ws + lo.w_colsum), scalars); t2_count_launch();
  T2_CHECK_CUDA(cudaGetLastError());
  // first conv
  first_conv_bwd_kernel<<<dim3((unsigned)((BT + 63) / 64)), lo.R, 0, st>>>(d_x, lo.scalar_in ? 1 : 0, dxin, d_grads + lo.p_in_k, BT, lo.R); t2_count_launch();
  T2_CHECK_CUDA(cudaGetLastError());
  // conditioning path
  if (lo.C > 0 && !cfg->c_pre_upsampled) {
    float* dcup = reinterpret_cast<float*>(ws + lo.w_dcup);
    ActGemmCall g;
    memset(&g, 0, sizeof(g));
    g.a[0] = a_dg; g.na = 1;
    g.seg[0] = Seg{0, 0, 0, lo.G / kBK, 0, lo.L}; g.nseg = 1;
    g.w = pk + lo.k_WcT; g.wN = lo.C; g.wK = lo.L * lo.G; g.wL = 1;
    g.T = lo.T; g.B = lo.B; g.n_tiles = 1;
    g.epi.ptr[0] = nullptr; g.epi.ptr[1] = nullptr; g.epi.ptr[2] = dcup;
    g.epi.i[0] = lo.C; g.epi.i[1] = 0; g.epi.i[2] = lo.C;
    rc = launch_act_gemm(EPI_BIAS_ACT, 128, g, st);
    if (rc) return rc;
    const float* dout = dcup;
    int cl = 1, pp = 0;
    for (int i = int(lo.up_w.size()) - 1; i >= 0; --i) {
      const int s = cfg->upsample_scales[i];
      const int W = lo.up_w[i] / s;
      const float* layer_in = i == 0 ? d_c : reinterpret_cast<const float*>(ws + lo.w_upout[i - 1]);
      const float* out = reinterpret_cast<const float*>(ws + lo.w_upout[i]);
      const int ntap = cfg->upsample_type == 0 ? 9 : 3;
      upsample_bwd_param_kernel<<<dim3(s, ntap + 1, 24), 256, 0, st>>>(layer_in, out, dout, cl, d_grads + lo.p_up_k[i], d_grads + lo.p_up_b[i],
                                                                  lo.B, lo.C, W, s, cfg->upsample_type); t2_count_launch();
      if (i > 0) {
        float* din = reinterpret_cast<float*>(ws + lo.w_upgrad[pp]);
        upsample_bwd_input_kernel<<<grid1d((long long)lo.B * lo.C * W), 256, 0, st>>>(out, dout, cl, d_params + lo.p_up_k[i], din, lo.B, lo.C, W, s,
                                                                                    cfg->upsample_type); t2_count_launch();
        dout = din;
        cl = 0;
        pp ^= 1;
      }
      T2_CHECK_CUDA(cudaGetLastError());
    }
  }
  return T2_OK;
}

extern "C" int t2_wn_workspace_tensor(const t2_wn_config_t* cfg, void* d_workspace, const char* name, void** ptr,
                                      long long* count, int* elem_bytes) {
  Layout lo;
  int rc = build_layout(cfg, lo);
  if (rc) return rc;
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  const long long BT = (long long)lo.B * lo.T;
  struct E { const char* n; long long off; long long cnt; int eb; };
  const E table[] = {
      {"c_up", lo.w_cup, BT * lo.C, 2},       {"x", lo.w_x, lo.L * BT * lo.R, 2},    {"xd", lo.w_xd, lo.L * BT * lo.R, 2},
      {"ta", lo.w_ta, lo.L * BT * lo.Gh, 2},  {"sb", lo.w_sb, lo.L * BT * lo.Gh, 2}, {"z", lo.w_z, lo.L * BT * lo.Gh, 2},
      {"h1", lo.w_h1, BT * lo.S, 2},          {"h2", lo.w_h2, BT * lo.S, 2},         {"dlog", lo.w_dlog, BT * lo.ldo, 2},
      {"dh2", lo.w_dh2, BT * lo.S, 2},        {"dskip", lo.w_dskip, BT * lo.S, 2},   {"dxin", lo.w_dxin, lo.L * BT * lo.R, 2},
      {"dg", lo.w_dg, lo.L * BT * lo.G, 2},   {"dc_up", lo.w_dcup, BT * lo.C, 4},    {"scalars", lo.w_scalars, 16, 4},
  };
  for (const E& e : table)
    if (strcmp(e.n, name) == 0) {
      *ptr = ws + e.off; *count = e.cnt; *elem_bytes = e.eb;
      return T2_OK;
    }
  return t2_set_error(T2_ERR_INVALID_ARG, "unknown workspace tensor '%s'", name);
}

// Times `reps` back-to-back launches of one per-layer GEMM of the residual stack with CUDA events on the launching
// stream: which = 0 gate (dilated conv + cin + tanh*sigmoid), 1 out 1x1 + residual, 2 dz + gate backward, 3 dx (data
// gradient of the dilated conv). The workspace must hold the state of a previous forward (+ backward). Synchronises.
extern "C" int t2_wn_time_kernel(const t2_wn_config_t* cfg, const float* d_params, const void* d_packed, void* d_workspace,
                                 int which, int layer, int reps, float* ms_per_launch, void* stream) {
  Layout lo;
  int rc = build_layout(cfg, lo);
  if (rc) return rc;
  T2_REQUIRE(layer >= 0 && layer < lo.L && reps >= 1 && ms_per_launch && which >= 0 && which <= 3, T2_ERR_INVALID_ARG,
             "time_kernel: bad arguments");
  T2_REQUIRE(which != 1 || layer + 1 < lo.L, T2_ERR_INVALID_ARG, "the last layer has no out GEMM");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  const uint8_t* pk = static_cast<const uint8_t*>(d_packed);
  ActGemmCall g;
  int epi, bn;
  if (which == 0) { g = make_gate_call(lo, ws, pk, layer, true); epi = EPI_GATE; bn = 256; }
  else if (which == 1) { g = make_out_call(lo, ws, pk, d_params, layer, cfg->dropout, 1, nullptr); epi = EPI_RES; bn = lo.R; }
  else if (which == 2) { g = make_dz_call(lo, ws, pk, layer); epi = EPI_GATE_BWD; bn = lo.Gh >= 256 ? 256 : 128; }
  else { g = make_dx_call(lo, ws, pk, layer, cfg->dropout, 1, nullptr); epi = EPI_DX; bn = lo.R; }
  cudaEvent_t e0, e1;
  T2_CHECK_CUDA(cudaEventCreate(&e0));
  T2_CHECK_CUDA(cudaEventCreate(&e1));
  rc = launch_act_gemm(epi, bn, g, st);  // warm-up
  if (rc) return rc;
  T2_CHECK_CUDA(cudaEventRecord(e0, st));
  for (int i = 0; i < reps; ++i) {
    rc = launch_act_gemm(epi, bn, g, st);
    if (rc) return rc;
  }
  T2_CHECK_CUDA(cudaEventRecord(e1, st));
  T2_CHECK_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  T2_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  *ms_per_launch = ms / reps;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return T2_OK;
}
