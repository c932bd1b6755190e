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
#include <stdlib.h>
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
  int part;           // 0 / 1: bf16(w) ; 2: bf16(w - bf16(w)), the low half of the split-bf16 operand
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
  bool split;   // split-bf16 forward (t2_wn_config_t.split_bf16)
  int xm;       // channel multiplier of the stored activations: 2 in split mode (hi | lo), else 1
  bool scalar_in, mol, gauss;   // mol: scalar-input head (mixture of logistics, or a single Gaussian when gauss)
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
  long long w_skipsum;
  long long w_upgrad[2], w_scalars, w_tiles_main, w_tiles_head, w_packjobs, w_colsum, w_tables;
  std::vector<long long> w_upout;
  std::vector<int> up_w;  // width after each upsample layer
  long long workspace_bytes;
  int n_tiles_main, n_tiles_head, n_packjobs, n_colsum;
  std::vector<PackJob> packjobs;
  std::vector<ColsumJob> colsums;
  std::vector<WgradTile> tiles_main, tiles_head;
  std::vector<int> tile_start;   // [L + 1]: first entry of tiles_main that belongs to layer l (the table is layer-major)
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
  lo.split = cfg->split_bf16 != 0;
  lo.xm = lo.split ? 2 : 1;
  T2_REQUIRE(!lo.split || cfg->dropout == 0.f, T2_ERR_INVALID_ARG, "split_bf16 (fp32-class forward) needs dropout = 0");
  lo.scalar_in = cfg->input_type != 2;
  lo.mol = lo.scalar_in;
  T2_REQUIRE(lo.L >= 1 && cfg->stacks >= 1 && lo.L % cfg->stacks == 0, T2_ERR_INVALID_ARG, "layers %% stacks != 0");
  T2_REQUIRE(cfg->kernel_size == 3, T2_ERR_UNSUPPORTED_SHAPE, "kernel_size must be 3");
  T2_REQUIRE(lo.R == 128 || lo.R == 256, T2_ERR_UNSUPPORTED_SHAPE, "residual_channels must be 128 or 256 (got %d)", lo.R);
  T2_REQUIRE(lo.S == 128 || lo.S == 256, T2_ERR_UNSUPPORTED_SHAPE, "skip_out_channels must be 128 or 256 (got %d)", lo.S);
  T2_REQUIRE(lo.Gh == 128 || lo.Gh == 256, T2_ERR_UNSUPPORTED_SHAPE, "gate_channels must be 256 or 512 (got %d)", lo.G);
  T2_REQUIRE(lo.C == 0 || (lo.C % 8 == 0 && lo.C <= 128), T2_ERR_UNSUPPORTED_SHAPE, "cin_channels must be 0 or a multiple of 8 <= 128");
  lo.gauss = lo.scalar_in && lo.O == 2;
  if (lo.mol) {
    T2_REQUIRE(lo.gauss || (lo.O % 3 == 0 && lo.O >= 3 && lo.O <= 30), T2_ERR_UNSUPPORTED_SHAPE,
               "scalar input needs out_channels = 2 (Gaussian) or 3 * nr_mix <= 30 (mixture of logistics), got %d", lo.O);
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
  const long long km = lo.split ? 3 : 1;   // split-bf16: every forward K segment becomes [W_hi | W_hi | W_lo]
  lo.k_Wg = takeb(L * lo.G * lo.Kg * km * 2);
  lo.k_Wo = takeb(L * lo.R * lo.Gh * km * 2);
  lo.k_Ws = takeb((long long)lo.S * L * lo.Gh * km * 2);
  lo.k_Wf1 = takeb((long long)lo.S * lo.S * km * 2);
  lo.k_Wf2 = takeb((long long)(lo.O < 32 ? 32 : lo.O) * lo.S * km * 2);
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
  lo.w_cup = takeb(BT * (lo.split ? 256 : (lo.C > 0 ? lo.C : 8)) * 2);     // split: [hi(C) pad 128 | lo(C) pad 128]
  lo.w_upout.clear();
  for (size_t i = 0; i < lo.up_w.size(); ++i) lo.w_upout.push_back(takeb((long long)lo.B * lo.C * lo.up_w[i] * 4));
  lo.w_upgrad[0] = takeb(BT * (lo.C > 0 ? lo.C : 8) * 4);
  lo.w_upgrad[1] = takeb(BT * (lo.C > 0 ? lo.C : 8) * 4);
  lo.w_x = takeb(L * BT * lo.R * 2 * lo.xm);
  lo.w_xd = cfg->dropout > 0.f ? takeb(L * BT * lo.R * 2) : lo.w_x;
  lo.w_ta = takeb(L * BT * lo.Gh * 2);
  lo.w_sb = takeb(L * BT * lo.Gh * 2);
  lo.w_z = takeb(L * BT * lo.Gh * 2 * lo.xm);
  lo.w_h1 = takeb(BT * lo.S * 2 * lo.xm);
  lo.w_h2 = takeb(BT * lo.S * 2 * lo.xm);
  lo.w_dlog = takeb(BT * lo.ldo * 2);
  lo.w_dh2 = takeb(BT * lo.S * 2);
  lo.w_dskip = takeb(BT * lo.S * 2);
  lo.w_dxin = takeb(L * BT * lo.R * 2);
  lo.w_dg = takeb(L * BT * lo.G * 2);
  lo.w_dcup = takeb(BT * (lo.C > 0 ? lo.C : 8) * 4);
  lo.w_scalars = takeb(64 * 4);
  lo.w_skipsum = takeb(lo.S * 4);

  // ---- pack jobs ----
  lo.packjobs.clear();
  auto pj = [&](long long src, int K, int N, long long dst_bytes, int ld, int transpose, int col0, float scale, int perm) {
    PackJob j; j.src_off = src; j.K = K; j.N = N; j.dst_off = dst_bytes / 2; j.dst_ld = ld; j.transpose = transpose;
    j.col0 = col0; j.scale = scale; j.perm_gh = perm; j.part = 0; lo.packjobs.push_back(j);
  };
  // split-bf16 forward operand: the K range [col0, col0 + slot) of the plain layout becomes [W_hi | W_hi | W_lo], slot columns each
  auto pj3 = [&](long long src, int K, int N, long long dst_bytes, int ld, int col_hi, int col_lo, int slot, float scale, int perm) {
    pj(src, K, N, dst_bytes, ld, 1, col_hi, scale, perm);
    pj(src, K, N, dst_bytes, ld, 1, col_hi + slot, scale, perm);
    pj(src, K, N, dst_bytes, ld, 1, col_lo, scale, perm);
    lo.packjobs.back().part = 2;
  };
  for (int l = 0; l < lo.L; ++l) {
    const long long wg = lo.k_Wg + (long long)l * lo.G * lo.Kg * km * 2;
    if (lo.split) {
      const int R = lo.R, Gh = lo.Gh;
      for (int j = 0; j < 3; ++j) pj3(lo.p_dil_k[l] + (long long)j * R * lo.G, R, lo.G, wg, 3 * lo.Kg, j * 3 * R, j * 3 * R + 2 * R, R, 1.f, Gh);
      if (lo.C > 0) pj3(lo.p_c_k[l], lo.C, lo.G, wg, 3 * lo.Kg, 9 * R, 9 * R + 256, 128, 1.f, Gh);
      pj3(lo.p_o_k[l], Gh, R, lo.k_Wo + (long long)l * R * Gh * 3 * 2, 3 * Gh, 0, 2 * Gh, Gh, 1.f, 0);
      // skip GEMM: K runs over [all layers: hi | hi] then [all layers: lo]
      pj3(lo.p_s_k[l], Gh, lo.S, lo.k_Ws, 3 * lo.L * Gh, l * 2 * Gh, 2 * lo.L * Gh + l * Gh, Gh, lo.skip_scale[l], 0);
    } else {
    for (int j = 0; j < 3; ++j) pj(lo.p_dil_k[l] + (long long)j * lo.R * lo.G, lo.R, lo.G, wg, lo.Kg, 1, j * lo.R, 1.f, lo.Gh);
    if (lo.C > 0) pj(lo.p_c_k[l], lo.C, lo.G, wg, lo.Kg, 1, 3 * lo.R, 1.f, lo.Gh);
    pj(lo.p_o_k[l], lo.Gh, lo.R, lo.k_Wo + (long long)l * lo.R * lo.Gh * 2, lo.Gh, 1, 0, 1.f, 0);
    pj(lo.p_s_k[l], lo.Gh, lo.S, lo.k_Ws, lo.L * lo.Gh, 1, l * lo.Gh, lo.skip_scale[l], 0);
    }
    const long long woz = lo.k_WozT + (long long)l * lo.Gh * (lo.R + lo.S) * 2;
    pj(lo.p_o_k[l], lo.Gh, lo.R, woz, lo.R + lo.S, 0, 0, lo.res_scale, 0);
    pj(lo.p_s_k[l], lo.Gh, lo.S, woz, lo.R + lo.S, 0, lo.R, lo.skip_scale[l], 0);
    const long long wd = lo.k_WdT + (long long)l * lo.R * 3 * lo.G * 2;
    for (int j = 0; j < 3; ++j) pj(lo.p_dil_k[l] + (long long)j * lo.R * lo.G, lo.R, lo.G, wd, 3 * lo.G, 0, j * lo.G, 1.f, 0);
    if (lo.C > 0) pj(lo.p_c_k[l], lo.C, lo.G, lo.k_WcT, lo.L * lo.G, 0, l * lo.G, 1.f, 0);
  }
  if (lo.split) {
    pj3(lo.p_f1_k, lo.S, lo.S, lo.k_Wf1, 3 * lo.S, 0, 2 * lo.S, lo.S, 1.f, 0);
    pj3(lo.p_f2_k, lo.S, lo.O, lo.k_Wf2, 3 * lo.S, 0, 2 * lo.S, lo.S, 1.f, 0);
  } else {
    pj(lo.p_f1_k, lo.S, lo.S, lo.k_Wf1, lo.S, 1, 0, 1.f, 0);
    pj(lo.p_f2_k, lo.S, lo.O, lo.k_Wf2, lo.S, 1, 0, 1.f, 0);
  }
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
  const int WN = 256;   // output columns per weight-gradient tile (wgrad_gemm_kernel: kWgBN)
  auto nmin = [&](int rem) { return rem < WN ? rem : WN; };
  lo.tile_start.assign(lo.L + 1, 0);
  for (int l = 0; l < lo.L; ++l) {
    const int d = lo.dil(l);
    lo.tile_start[l] = int(lo.tiles_main.size());
    for (int j = 0; j < 3; ++j)
      for (int m0 = 0; m0 < lo.R; m0 += 128)
        for (int n0 = 0; n0 < lo.G; n0 += WN)
          wt(lo.tiles_main, 0, m0, -(2 - j) * d, l, 1, n0, l, lo.p_dil_k[l] + (long long)j * lo.R * lo.G + (long long)m0 * lo.G + n0,
             lo.G, 128, nmin(lo.G - n0), 1.f, nullptr);
    if (lo.C > 0)
      for (int n0 = 0; n0 < lo.G; n0 += WN)
        wt(lo.tiles_main, 2, 0, 0, 0, 1, n0, l, lo.p_c_k[l] + n0, lo.G, lo.C, nmin(lo.G - n0), 1.f, nullptr);
    for (int m0 = 0; m0 < lo.Gh; m0 += 128) {
      if (l < lo.L - 1)
        for (int n0 = 0; n0 < lo.R; n0 += WN)
          wt(lo.tiles_main, 3, m0, 0, l, 4, n0, l + 1, lo.p_o_k[l] + (long long)m0 * lo.R + n0, lo.R, 128, nmin(lo.R - n0), lo.res_scale, nullptr);
      for (int n0 = 0; n0 < lo.S; n0 += WN)
        wt(lo.tiles_main, 3, m0, 0, l, 5, n0, 0, lo.p_s_k[l] + (long long)m0 * lo.S + n0, lo.S, 128, nmin(lo.S - n0), lo.skip_scale[l], nullptr);
    }
  }
  // head maps: 0 h1, 1 dh2, 2 h2, 3 dlog
  lo.tiles_head.clear();
  for (int m0 = 0; m0 < lo.S; m0 += 128) {
    for (int n0 = 0; n0 < lo.S; n0 += WN)
      wt(lo.tiles_head, 0, m0, 0, 0, 1, n0, 0, lo.p_f1_k + (long long)m0 * lo.S + n0, lo.S, 128, nmin(lo.S - n0), 1.f, nullptr);
    for (int n0 = 0; n0 < lo.O; n0 += WN)
      for (int part = 0; part < (lo.mol ? 1 : 2); ++part) {  // CE: hi and lo halves of dlog both accumulate (atomics)
        wt(lo.tiles_head, 2, m0, 0, 0, 3, part * 256 + n0, 0, lo.p_f2_k + (long long)m0 * lo.O + n0, lo.O, 128,
           nmin(lo.O - n0), 1.f, reinterpret_cast<const float*>(1) /* patched to scalars[1] at init */);
        lo.tiles_head.back().accumulate = 2;
      }
  }
  lo.n_tiles_main = int(lo.tiles_main.size());
  lo.tile_start[lo.L] = lo.n_tiles_main;
  lo.n_tiles_head = int(lo.tiles_head.size());

  // ---- column-sum (bias gradient) jobs ----
  lo.colsums.clear();
  auto cs = [&](long long src, long long rows, int C, int ld, long long dst, long long dst2, float scale, int div) {
    ColsumJob j; j.src_off = src; j.rows = rows; j.C = C; j.ld = ld; j.dst_off = dst; j.dst2_off = dst2; j.scale = scale;
    j.div_scalar = div; lo.colsums.push_back(j);
  };
  // (all other bias gradients are column sums fused into the GEMM epilogues that produce dg / dx / dskip / dh2)
  cs(lo.w_dlog, BT, lo.O, lo.ldo, lo.p_f2_b, -1, 1.f, 1);
  if (!lo.mol) cs(lo.w_dlog + 256 * 2, BT, lo.O, lo.ldo, lo.p_f2_b, -1, 1.f, 1);
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
  // 64x64 tiles through shared memory: float2 reads along the source's fast axis (N), bf16x2 writes along the
  // destination's fast axis (K for the transposing jobs). All offsets / leading dimensions in the job table are even.
  __shared__ float tile[64][65];
  const PackJob j = jobs[blockIdx.y];
  const int tiles_n = (j.N + 63) / 64, tiles_k = (j.K + 63) / 64;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const bool vec_src = ((j.N | int(j.src_off)) & 1) == 0;
  const bool vec_dst = ((j.dst_ld | j.col0 | int(j.dst_off)) & 1) == 0;
  for (int ti = blockIdx.x; ti < tiles_n * tiles_k; ti += gridDim.x) {
    const int k0 = (ti / tiles_n) * 64, n0 = (ti % tiles_n) * 64;
    for (int r = ty; r < 64; r += 8) {
      const int k = k0 + r, n = n0 + 2 * tx;
      float a = 0.f, b = 0.f;
      if (k < j.K) {
        const float* src = params + j.src_off + (long long)k * j.N + n;
        if (vec_src && n + 1 < j.N) { const float2 v = *reinterpret_cast<const float2*>(src); a = v.x; b = v.y; }
        else { if (n < j.N) a = src[0]; if (n + 1 < j.N) b = src[1]; }
      }
      a *= j.scale; b *= j.scale;
      if (j.part == 2) { a -= __bfloat162float(__float2bfloat16(a)); b -= __bfloat162float(__float2bfloat16(b)); }
      tile[r][2 * tx] = a; tile[r][2 * tx + 1] = b;
    }
    __syncthreads();
    if (j.transpose) {
      for (int r = ty; r < 64; r += 8) {
        const int n = n0 + r, k = k0 + 2 * tx;
        if (n < j.N && k < j.K) {
          int row = n;
          if (j.perm_gh > 0) {
            const int half = n / j.perm_gh, idx = n % j.perm_gh;
            row = (idx / 128) * 256 + half * 128 + (idx % 128);
          }
          bf16* dst = packed + j.dst_off + (long long)row * j.dst_ld + j.col0 + k;
          if (vec_dst && k + 1 < j.K) *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(tile[2 * tx][r], tile[2 * tx + 1][r]);
          else { dst[0] = __float2bfloat16(tile[2 * tx][r]); if (k + 1 < j.K) dst[1] = __float2bfloat16(tile[2 * tx + 1][r]); }
        }
      }
    } else {
      for (int r = ty; r < 64; r += 8) {
        const int k = k0 + r, n = n0 + 2 * tx;
        if (n < j.N && k < j.K) {
          bf16* dst = packed + j.dst_off + (long long)k * j.dst_ld + j.col0 + n;
          if (vec_dst && n + 1 < j.N) *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(tile[r][2 * tx], tile[r][2 * tx + 1]);
          else { dst[0] = __float2bfloat16(tile[r][2 * tx]); if (n + 1 < j.N) dst[1] = __float2bfloat16(tile[r][2 * tx + 1]); }
        }
      }
    }
    __syncthreads();
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
                                  const unsigned long long* __restrict__ step, int split) {
  // 8 channels per thread (R % 8 == 0): two float4 loads of the embedding row, one 16-byte store per output
  if (step) seed += *step;
  const long long e8 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int R8 = R >> 3;
  if (e8 >= npos * R8) return;
  const long long pos = e8 / R8;
  const int r = int(e8 % R8) * 8;
  float v[8];
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + r)), b1 = __ldg(reinterpret_cast<const float4*>(bias + r + 4));
  if (scalar_in) {
    const float xv = static_cast<const float*>(xin)[pos];
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(W + r)), w1 = __ldg(reinterpret_cast<const float4*>(W + r + 4));
    v[0] = xv * w0.x + b0.x; v[1] = xv * w0.y + b0.y; v[2] = xv * w0.z + b0.z; v[3] = xv * w0.w + b0.w;
    v[4] = xv * w1.x + b1.x; v[5] = xv * w1.y + b1.y; v[6] = xv * w1.z + b1.z; v[7] = xv * w1.w + b1.w;
  } else {
    const float* row = W + (long long)static_cast<const int*>(xin)[pos] * R + r;
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(row)), w1 = __ldg(reinterpret_cast<const float4*>(row + 4));
    v[0] = w0.x + b0.x; v[1] = w0.y + b0.y; v[2] = w0.z + b0.z; v[3] = w0.w + b0.w;
    v[4] = w1.x + b1.x; v[5] = w1.y + b1.y; v[6] = w1.z + b1.z; v[7] = w1.w + b1.w;
  }
  const long long e = pos * R + r;
  uint4 o;
  if (split) {   // rows are [hi(R) | lo(R)]
    float hi[8], lo8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { hi[j] = __bfloat162float(__float2bfloat16(v[j])); lo8[j] = v[j] - hi[j]; }
    o.x = pack_bf16x2(hi[0], hi[1]); o.y = pack_bf16x2(hi[2], hi[3]); o.z = pack_bf16x2(hi[4], hi[5]); o.w = pack_bf16x2(hi[6], hi[7]);
    *reinterpret_cast<uint4*>(x + pos * 2 * R + r) = o;
    o.x = pack_bf16x2(lo8[0], lo8[1]); o.y = pack_bf16x2(lo8[2], lo8[3]); o.z = pack_bf16x2(lo8[4], lo8[5]); o.w = pack_bf16x2(lo8[6], lo8[7]);
    *reinterpret_cast<uint4*>(x + pos * 2 * R + R + r) = o;
    return;
  }
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(x + e) = o;
  if (xd != x && xd != nullptr) {
    const float keep_inv = 1.f / (1.f - p);
    const uint32_t hs = hash_seed(seed, 0u), thr = uint32_t(p * 65536.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = hash_keep16(hs, (unsigned long long)(e + j), thr) ? v[j] * keep_inv : 0.f;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(xd + e) = o;
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
                                    float* __restrict__ out, bf16* __restrict__ out_cl, int B, int H, int W, int s, int type, int split = 0) {
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
  if (out_cl && split) {   // rows are [hi(H) zero-padded to 128 | lo(H) zero-padded to 128]
    const bf16 hi = __float2bfloat16(acc);
    bf16* row = out_cl + ((long long)b * Wo + xo) * 256;
    row[h] = hi;
    row[128 + h] = __float2bfloat16(acc - __bfloat162float(hi));
  } else if (out_cl) out_cl[((long long)b * Wo + xo) * H + h] = __float2bfloat16(acc);
}
// channels-last fp32 [B][T][C] -> [B][C][T] (tiled transpose) so the upsampling backward reads contiguously
__global__ void cl_to_chw_kernel(const float* __restrict__ in, float* __restrict__ out, int T, int C) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int t = t0 + r, c = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (t < T && c < C) ? in[((long long)b * T + t) * C + c] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int c = c0 + r, t = t0 + threadIdx.x;
    if (t < T && c < C) out[((long long)b * C + c) * T + t] = tile[threadIdx.x][r];
  }
}
// d_pre = d_out * (out > 0); accumulates dK [ntap][s] and dbias. The launch uses gridDim.x * blockDim.x % s == 0, so a
// thread always meets the same sub-pixel phase k = e % s: it sums its taps in registers, the block merges through a
// small shared-memory table (10 atomics per thread, once) and issues one global atomic per table entry.
__global__ void upsample_bwd_param_kernel(const float* __restrict__ in, const float* __restrict__ out, const float* __restrict__ dout,
                                          float* __restrict__ dK, float* __restrict__ dbias, int B, int H, int W, int s, int type) {
  __shared__ float acc[10 * 32];
  const int ntap = type == 0 ? 9 : 3;
  for (int i = threadIdx.x; i < (ntap + 1) * s; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const int Wo = W * s;
  const long long n = (long long)B * H * Wo;
  const long long e0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int k = int(e0 % s);
  float r[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) r[i] = 0.f;
  for (long long e = e0; e < n; e += (long long)gridDim.x * blockDim.x) {
    if (out[e] <= 0.f) continue;
    const float g = dout[e];
    const int xo = int(e % Wo), h = int((e / Wo) % H), b = int(e / ((long long)Wo * H));
    const int w = xo / s;
    r[9] += g;
    const float* ib = in + (long long)b * H * W;
    if (type == 0) {
#pragma unroll
      for (int dh = 0; dh < 3; ++dh) {
        const int hh = h + dh - 1;
        if (hh < 0 || hh >= H) continue;
#pragma unroll
        for (int dw = 0; dw < 3; ++dw) {
          const int ww = w + dw - 1;
          if (ww >= 0 && ww < W) r[dh * 3 + dw] += g * ib[hh * W + ww];
        }
      }
    } else {
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int hh = h + 1 - q;
        if (hh >= 0 && hh < H) r[q] += g * ib[hh * W + w];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 9; ++i)
    if (i < ntap && r[i] != 0.f) atomicAdd(&acc[i * s + k], r[i]);
  if (r[9] != 0.f) atomicAdd(&acc[ntap * s + k], r[9]);
  __syncthreads();
  for (int i = threadIdx.x; i < (ntap + 1) * s; i += blockDim.x) {
    const float v = acc[i];
    if (v == 0.f) continue;
    if (i < ntap * s) atomicAdd(dK + i, v);
    else if (type == 0) atomicAdd(dbias + (i - ntap * s), v);
    else atomicAdd(dbias, v);
  }
}
__global__ void upsample_bwd_input_kernel(const float* __restrict__ out, const float* __restrict__ dout, int cl,
                                          const float* __restrict__ K, float* __restrict__ din, int B, int H, int W, int s, int type) {
  // 4 lanes per input pixel, each walking every 4th sub-pixel phase k; partial sums meet through two shuffles
  const long long e4 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long n = (long long)B * H * W;
  const long long e = e4 >> 2;
  const int part = int(e4 & 3);
  const bool live = e < n;
  const long long ec = live ? e : n - 1;
  const int w = int(ec % W), h = int((ec / W) % H), b = int(ec / ((long long)W * H));
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
        for (int k = part; k < s; k += 4) acc += dpre(ho, wo * s + k) * K[(dh * 3 + dw) * s + k];
      }
    }
  } else {
    for (int q = 0; q < 3; ++q) {
      const int ho = h - 1 + q;
      if (ho < 0 || ho >= H) continue;
      for (int k = part; k < s; k += 4) acc += dpre(ho, w * s + k) * K[q * s + k];
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  if (live && part == 0) din[e] = acc;
}
// skip-conv bias gradients: db_s[l] = skip_scale[l] * column sums of dskip (table layout: offs[3l+2] = skip bias offset)
__global__ void skip_bias_kernel(const float* __restrict__ skipsum, float* __restrict__ grads, const long long* __restrict__ offs,
                                 const float* __restrict__ scales, int L, int S) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L * S) return;
  const int l = i / S, s = i % S;
  grads[offs[3 * l + 2] + s] = scales[l] * skipsum[s];
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
  if (lo.split) {
    // rows [hi | lo]: per tap one segment over both halves against [W_hi | W_hi] and one over the hi half against [W_lo]
    g.a[0] = make_act(ws + lo.w_xd, 2 * lo.R, lo.T, lo.B, lo.L);
    g.a[1] = make_act(ws + lo.w_cup, 256, lo.T, lo.B, 1);
    g.na = lo.C > 0 ? 2 : 1;
    const int shifts[3] = {-2 * d, -d, 0};
    g.nseg = 0;
    for (int j = 0; j < 3; ++j) {
      g.seg[g.nseg++] = Seg{0, shifts[j], 0, 2 * lo.R / kBK, l, 1};
      g.seg[g.nseg++] = Seg{0, shifts[j], 0, lo.R / kBK, l, 1};
    }
    if (lo.C > 0) { g.seg[g.nseg++] = Seg{1, 0, 0, 4, 0, 1}; g.seg[g.nseg++] = Seg{1, 0, 0, 2, 0, 1}; }
    g.w = pk + lo.k_Wg; g.wN = lo.G; g.wK = 3 * lo.Kg; g.wL = lo.L; g.w_layer = l; g.w_k0 = 0;
  } else {
  g.a[0] = make_act(ws + lo.w_xd, lo.R, lo.T, lo.B, lo.L);
  g.a[1] = make_act(ws + lo.w_cup, lo.C > 0 ? lo.C : 8, lo.T, lo.B, 1);
  g.na = lo.C > 0 ? 2 : 1;
  g.seg[0] = Seg{0, -2 * d, 0, lo.R / kBK, l, 1};
  g.seg[1] = Seg{0, -d, 0, lo.R / kBK, l, 1};
  g.seg[2] = Seg{0, 0, 0, lo.R / kBK, l, 1};
  g.nseg = 3;
  if (lo.C > 0) { g.seg[3] = Seg{1, 0, 0, 2, 0, 1}; g.nseg = 4; }
  g.w = pk + lo.k_Wg; g.wN = lo.G; g.wK = lo.Kg; g.wL = lo.L; g.w_layer = l; g.w_k0 = 0;
  }
  g.T = lo.T; g.B = lo.B; g.n_tiles = lo.G / 256;
  const long long lofs = (long long)l * BT * lo.Gh;
  g.epi.ptr[0] = save ? reinterpret_cast<bf16*>(ws + lo.w_ta) + lofs : nullptr;
  g.epi.ptr[1] = save ? reinterpret_cast<bf16*>(ws + lo.w_sb) + lofs : nullptr;
  g.epi.ptr[2] = reinterpret_cast<bf16*>(ws + lo.w_z) + lofs * lo.xm;
  g.epi.i[11] = lo.split ? 1 : 0;
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
  o.a[0] = make_act(ws + lo.w_z, lo.Gh * lo.xm, lo.T, lo.B, lo.L); o.na = 1;
  o.seg[0] = Seg{0, 0, 0, lo.Gh * lo.xm / kBK, l, 1}; o.nseg = 1;
  if (lo.split) { o.seg[1] = Seg{0, 0, 0, lo.Gh / kBK, l, 1}; o.nseg = 2; }
  o.w = pk + lo.k_Wo; o.wN = lo.R; o.wK = lo.Gh * (lo.split ? 3 : 1); o.wL = lo.L; o.w_layer = l;
  o.T = lo.T; o.B = lo.B; o.n_tiles = 1;
  o.epi.ptr[0] = x_all + (long long)l * BT * lo.R * lo.xm;
  o.epi.ptr[1] = x_all + (long long)(l + 1) * BT * lo.R * lo.xm;
  o.epi.i[11] = lo.split ? 1 : 0;
  o.epi.ptr[2] = p > 0.f ? xd_all + (long long)(l + 1) * BT * lo.R : nullptr;
  o.epi.ptr[3] = const_cast<float*>(params + lo.p_o_b[l]);
  o.epi.f[0] = lo.res_scale; o.epi.f[1] = p; o.epi.i[1] = l + 1; o.epi.seed = seed;
  o.epi.ptr[7] = const_cast<unsigned long long*>(d_step);
  return o;
}

ActGemmCall make_dz_call(const Layout& lo, uint8_t* ws, const uint8_t* pk, int l, float* grads) {
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
  g.epi.ptr[3] = grads ? grads + lo.p_dil_b[l] : nullptr;
  g.epi.ptr[4] = (grads && lo.C > 0) ? grads + lo.p_c_b[l] : nullptr;
  g.epi.i[0] = lo.Gh;
  return g;
}

ActGemmCall make_dx_call(const Layout& lo, uint8_t* ws, const uint8_t* pk, int l, float p, unsigned long long seed,
                         const unsigned long long* d_step, float* grads) {
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
  // dx of layer l is the gradient of x_l = output of layer l-1's out-1x1 (or of the first conv): fused bias gradient
  g.epi.ptr[2] = grads ? grads + (l > 0 ? lo.p_o_b[l - 1] : lo.p_in_b) : nullptr;
  g.epi.f[2] = l > 0 ? lo.res_scale : 1.f;
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
  pack_kernel<<<dim3(16, lo.n_packjobs), dim3(32, 8), 0, st>>>(d_params, reinterpret_cast<bf16*>(pk),
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

  T2_REQUIRE(!lo.split || (!save_for_backward && !cfg->c_pre_upsampled), T2_ERR_INVALID_ARG,
             "split_bf16 (fp32-class) mode is forward / loss only (save_for_backward = 0) and needs the upsampling net");
  // 1. conditioning -> c_up (bf16 channels-last)
  bf16* c_up = reinterpret_cast<bf16*>(ws + lo.w_cup);
  if (lo.split) T2_CHECK_CUDA(cudaMemsetAsync(c_up, 0, (size_t)BT * 256 * 2, st));     // the channel padding of both halves must read as zero
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
            in, d_params + lo.p_up_k[i], d_params + lo.p_up_b[i], out, last ? c_up : nullptr, lo.B, lo.C, W, s, cfg->upsample_type,
            lo.split ? 1 : 0); t2_count_launch();
        in = out;
        W *= s;
      }
    }
    T2_CHECK_CUDA(cudaGetLastError());
  }
  // 2. first conv
  bf16* x_all = reinterpret_cast<bf16*>(ws + lo.w_x);
  bf16* xd_all = reinterpret_cast<bf16*>(ws + lo.w_xd);
  first_conv_kernel<<<grid1d(BT * (lo.R / 8)), 256, 0, st>>>(d_x, lo.scalar_in ? 1 : 0, d_params + lo.p_in_k, d_params + lo.p_in_b,
                                                       x_all, p > 0.f ? xd_all : x_all, BT, lo.R, p, seed, d_step, lo.split ? 1 : 0); t2_count_launch();
  T2_CHECK_CUDA(cudaGetLastError());
  // 3. residual stack
  bf16* z_all = reinterpret_cast<bf16*>(ws + lo.w_z);
  const ActT a_z = make_act(z_all, lo.Gh * lo.xm, lo.T, lo.B, lo.L);
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
    g.seg[0] = Seg{0, 0, 0, lo.Gh * lo.xm / kBK, 0, lo.L}; g.nseg = 1;
    if (lo.split) { g.seg[1] = Seg{0, 0, 0, lo.Gh / kBK, 0, lo.L}; g.nseg = 2; }
    g.w = pk + lo.k_Ws; g.wN = lo.S; g.wK = lo.L * lo.Gh * (lo.split ? 3 : 1); g.wL = 1;
    g.T = lo.T; g.B = lo.B; g.n_tiles = 1;
    g.epi.ptr[0] = h1; g.epi.ptr[1] = const_cast<float*>(reinterpret_cast<const float*>(pk + lo.k_bias_skip));
    g.epi.i[0] = lo.S; g.epi.i[1] = 1; g.epi.i[2] = lo.S; g.epi.i[11] = lo.split ? 1 : 0;
    rc = launch_act_gemm(EPI_BIAS_ACT, lo.S, g, st);
    if (rc) return rc;
  }
  {
    ActGemmCall g;
    memset(&g, 0, sizeof(g));
    g.a[0] = make_act(h1, lo.S * lo.xm, lo.T, lo.B); g.na = 1;
    g.seg[0] = Seg{0, 0, 0, lo.S * lo.xm / kBK, 0, 1}; g.nseg = 1;
    if (lo.split) { g.seg[1] = Seg{0, 0, 0, lo.S / kBK, 0, 1}; g.nseg = 2; }
    g.w = pk + lo.k_Wf1; g.wN = lo.S; g.wK = lo.S * (lo.split ? 3 : 1); g.wL = 1;
    g.T = lo.T; g.B = lo.B; g.n_tiles = 1;
    g.epi.ptr[0] = h2; g.epi.ptr[1] = const_cast<float*>(d_params + lo.p_f1_b);
    g.epi.i[0] = lo.S; g.epi.i[1] = 1; g.epi.i[2] = lo.S; g.epi.i[11] = lo.split ? 1 : 0;
    rc = launch_act_gemm(EPI_BIAS_ACT, lo.S, g, st);
    if (rc) return rc;
  }
  // 5. output projection fused with the loss
  {
    ActGemmCall g;
    memset(&g, 0, sizeof(g));
    g.a[0] = make_act(h2, lo.S * lo.xm, lo.T, lo.B); g.na = 1;
    g.seg[0] = Seg{0, 0, 0, lo.S * lo.xm / kBK, 0, 1}; g.nseg = 1;
    if (lo.split) { g.seg[1] = Seg{0, 0, 0, lo.S / kBK, 0, 1}; g.nseg = 2; }
    g.w = pk + lo.k_Wf2; g.wN = lo.O; g.wK = lo.S * (lo.split ? 3 : 1); g.wL = 1;
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
      g.epi.i[0] = lo.gauss ? 0 : lo.O / 3;
      g.epi.i[2] = lo.gauss ? (cfg->cdf_loss ? 2 : 1) : 0;     // 0 mixture of logistics, 1 Gaussian log-density, 2 Gaussian CDF difference
      g.epi.f[3] = cfg->log_scale_min_gauss;
      rc = launch_act_gemm(EPI_MOL, 32, g, st);
    } else {
      rc = launch_act_gemm(EPI_CE, 256, g, st);
    }
    if (rc) return rc;
  }
  if (d_loss) T2_CHECK_CUDA(cudaMemcpyAsync(d_loss, scalars, 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return T2_OK;
}

// side stream + fork/join events for the independent tail of the backward pass (created once per process; T2_SIDE_STREAM=0
// in the environment keeps everything on the caller's stream)
struct SideStream { cudaStream_t s; cudaEvent_t fork, fork2, join; };
static SideStream* side_stream() {
  static SideStream ss;
  static int state = 0;   // 0 unknown, 1 ready, -1 disabled
  if (state == 0) {
    const char* e = getenv("T2_SIDE_STREAM");
    if (e && e[0] == '0') state = -1;
    else if (cudaStreamCreateWithFlags(&ss.s, cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ss.fork2, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming) == cudaSuccess) state = 1;
    else state = -1;
  }
  return state == 1 ? &ss : nullptr;
}

extern "C" int t2_wn_backward(const t2_wn_config_t* cfg, const float* d_params, const void* d_packed,
                              void* d_workspace, const void* d_x, const float* d_c, float* d_grads,
                              unsigned long long seed, const unsigned long long* d_step, void* stream) {
  return t2_wn_backward_phased(cfg, d_params, d_packed, d_workspace, d_x, d_c, d_grads, seed, d_step, -1, 1, stream);
}

// Phased form for data-parallel training: the weight gradients of the residual stack are produced by `n_groups` launches (layer
// groups, top of the parameter buffer first) so that the caller can start the gradient all-reduce of a group's contiguous
// parameter range while the next group's GEMM runs. phase -1: everything in one call (== t2_wn_backward); phase 0: the data-gradient
// chain + head + conditioning tails (no stack weight gradients, side stream NOT yet joined); phase 1 + g: weight gradients of layer
// group g (layers [g*L/n, (g+1)*L/n)); phase 100: join the side stream. wavenet.py:561-593 averages tower gradients after backward.
extern "C" int t2_wn_backward_phased(const t2_wn_config_t* cfg, const float* d_params, const void* d_packed,
                                     void* d_workspace, const void* d_x, const float* d_c, float* d_grads,
                                     unsigned long long seed, const unsigned long long* d_step, int phase, int n_groups, void* stream) {
  Layout lo;
  int rc = build_layout(cfg, lo);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  T2_REQUIRE(!lo.split, T2_ERR_INVALID_ARG, "split_bf16 (fp32-class) mode has no backward pass");
  T2_REQUIRE(n_groups >= 1 && n_groups <= lo.L && (phase == -1 || phase == 100 || (phase >= 0 && phase <= n_groups)), T2_ERR_INVALID_ARG,
             "backward_phased: bad phase %d / n_groups %d", phase, n_groups);
  if (phase >= 1 && phase <= n_groups) {
    uint8_t* ws = static_cast<uint8_t*>(d_workspace);
    const int g = phase - 1;
    const int l0 = int((long long)lo.L * g / n_groups), l1 = int((long long)lo.L * (g + 1) / n_groups);
    const int t0 = lo.tile_start[l0], t1 = lo.tile_start[l1];
    ActT maps[6] = {make_act(ws + lo.w_xd, lo.R, lo.T, lo.B, lo.L), make_act(ws + lo.w_dg, lo.G, lo.T, lo.B, lo.L),
                    make_act(ws + lo.w_cup, lo.C > 0 ? lo.C : 8, lo.T, lo.B, 1),
                    make_act(ws + lo.w_z, lo.Gh, lo.T, lo.B, lo.L), make_act(ws + lo.w_dxin, lo.R, lo.T, lo.B, lo.L),
                    make_act(ws + lo.w_dskip, lo.S, lo.T, lo.B, 1)};
    return launch_wgrad(maps, 6, reinterpret_cast<const WgradTile*>(ws + lo.w_tiles_main) + t0, t1 - t0, d_grads, lo.T, lo.B, st);
  }
  if (phase == 100) {
    SideStream* side = side_stream();
    if (side) {
      T2_CHECK_CUDA(cudaEventRecord(side->join, side->s));
      T2_CHECK_CUDA(cudaStreamWaitEvent(st, side->join, 0));
    }
    return T2_OK;
  }
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  const uint8_t* pk = static_cast<const uint8_t*>(d_packed);
  const long long BT = (long long)lo.B * lo.T;
  float* scalars = reinterpret_cast<float*>(ws + lo.w_scalars);
  const float p = cfg->dropout;
  T2_CHECK_CUDA(cudaMemsetAsync(d_grads, 0, lo.n_params * sizeof(float), st));
  T2_CHECK_CUDA(cudaMemsetAsync(ws + lo.w_skipsum, 0, lo.S * sizeof(float), st));
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
    g.epi.ptr[3] = d_grads + lo.p_f1_b;
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
    g.epi.ptr[3] = ws + lo.w_skipsum;
    rc = launch_act_gemm(EPI_SCALE_RELUMASK, lo.S, g, st);
    if (rc) return rc;
  }
  skip_bias_kernel<<<grid1d((long long)lo.L * lo.S), 256, 0, st>>>(reinterpret_cast<const float*>(ws + lo.w_skipsum), d_grads,
      reinterpret_cast<const long long*>(ws + lo.w_tables), reinterpret_cast<const float*>(ws + lo.w_tables + 3 * lo.L * sizeof(long long)), lo.L, lo.S);
  t2_count_launch();
  // The head's weight gradients (6 tiles with a 15360-long reduction: ~100 us on 6 SMs) and the bias column sums of dlog
  // depend only on the head backward above: they run on the side stream underneath the whole residual-stack chain below,
  // which leaves 28 SMs idle (120 M tiles on 148 SMs).
  SideStream* side = side_stream();
  cudaStream_t sb = st;
  if (side) {
    T2_CHECK_CUDA(cudaEventRecord(side->fork, st));
    T2_CHECK_CUDA(cudaStreamWaitEvent(side->s, side->fork, 0));
    sb = side->s;
  }
  {
    ActT hmaps[4] = {make_act(h1, lo.S, lo.T, lo.B), make_act(dh2, lo.S, lo.T, lo.B), make_act(h2, lo.S, lo.T, lo.B),
                     make_act(dlog, lo.ldo, lo.T, lo.B)};
    rc = launch_wgrad(hmaps, 4, reinterpret_cast<const WgradTile*>(ws + lo.w_tiles_head), lo.n_tiles_head, d_grads, lo.T, lo.B, sb);
    if (rc) return rc;
    colsum_kernel<<<dim3(96, lo.n_colsum), 256, 0, sb>>>(ws, d_grads, reinterpret_cast<const ColsumJob*>(ws + lo.w_colsum), scalars); t2_count_launch();
    T2_CHECK_CUDA(cudaGetLastError());
  }
  // residual stack, top down
  const ActT a_dxin = make_act(dxin, lo.R, lo.T, lo.B, lo.L);
  const ActT a_dskip = make_act(dskip, lo.S, lo.T, lo.B, 1);
  const ActT a_dg = make_act(dg, lo.G, lo.T, lo.B, lo.L);
  const int bn_z = lo.Gh >= 256 ? 256 : 128;
  for (int l = lo.L - 1; l >= 0; --l) {
    ActGemmCall gz = make_dz_call(lo, ws, pk, l, d_grads);
    rc = launch_act_gemm(EPI_GATE_BWD, bn_z, gz, st);
    if (rc) return rc;
    ActGemmCall gx = make_dx_call(lo, ws, pk, l, p, seed, d_step, d_grads);
    rc = launch_act_gemm(EPI_DX, lo.R, gx, st);
    if (rc) return rc;
  }
  // From here on two independent tails: (A) the batched weight-gradient GEMM of the stack (fills the machine), (B) the
  // conditioning path (K = L*G data-gradient GEMM, transposes, upsampling-net backward) + first-conv gradient: latency-bound
  // small kernels. (B) continues on the side stream (fork/join through events, capturable into the caller's CUDA graph).
  if (side) {
    T2_CHECK_CUDA(cudaEventRecord(side->fork2, st));
    T2_CHECK_CUDA(cudaStreamWaitEvent(side->s, side->fork2, 0));
  }
  // weight gradients of the stack: one batched launch
  if (phase == -1) {
    ActT maps[6] = {make_act(ws + lo.w_xd, lo.R, lo.T, lo.B, lo.L), a_dg,
                    make_act(ws + lo.w_cup, lo.C > 0 ? lo.C : 8, lo.T, lo.B, 1),
                    make_act(ws + lo.w_z, lo.Gh, lo.T, lo.B, lo.L), a_dxin, a_dskip};
    rc = launch_wgrad(maps, 6, reinterpret_cast<const WgradTile*>(ws + lo.w_tiles_main), lo.n_tiles_main, d_grads, lo.T, lo.B, st);
    if (rc) return rc;
  }
  // first conv
  first_conv_bwd_kernel<<<dim3((unsigned)((BT + 63) / 64)), lo.R, 0, sb>>>(d_x, lo.scalar_in ? 1 : 0, dxin, d_grads + lo.p_in_k, BT, lo.R); t2_count_launch();
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
    rc = launch_act_gemm(EPI_BIAS_ACT, 128, g, sb);
    if (rc) return rc;
    // dc_up arrives channels-last from the GEMM: transpose once into [B][C][T]
    float* dchw = reinterpret_cast<float*>(ws + lo.w_upgrad[1]);
    cl_to_chw_kernel<<<dim3((lo.T + 31) / 32, (lo.C + 31) / 32, lo.B), dim3(32, 8), 0, sb>>>(dcup, dchw, lo.T, lo.C); t2_count_launch();
    const float* dout = dchw;
    int pp = 0;
    for (int i = int(lo.up_w.size()) - 1; i >= 0; --i) {
      const int s = cfg->upsample_scales[i];
      const int W = lo.up_w[i] / s;
      T2_REQUIRE(s <= 32, T2_ERR_UNSUPPORTED_SHAPE, "upsample scale > 32");
      const float* layer_in = i == 0 ? d_c : reinterpret_cast<const float*>(ws + lo.w_upout[i - 1]);
      const float* out = reinterpret_cast<const float*>(ws + lo.w_upout[i]);
      upsample_bwd_param_kernel<<<s * ((296 + s - 1) / s), 256, 0, sb>>>(layer_in, out, dout, d_grads + lo.p_up_k[i], d_grads + lo.p_up_b[i], lo.B, lo.C, W, s,
                                                     cfg->upsample_type); t2_count_launch();
      if (i > 0) {
        float* din = reinterpret_cast<float*>(ws + lo.w_upgrad[pp]);
        upsample_bwd_input_kernel<<<grid1d(4LL * lo.B * lo.C * W), 256, 0, sb>>>(out, dout, 0, d_params + lo.p_up_k[i], din, lo.B, lo.C, W, s,
                                                                                    cfg->upsample_type); t2_count_launch();
        dout = din;
        pp ^= 1;
      }
      T2_CHECK_CUDA(cudaGetLastError());
    }
  }
  if (side && phase == -1) {
    T2_CHECK_CUDA(cudaEventRecord(side->join, sb));
    T2_CHECK_CUDA(cudaStreamWaitEvent(st, side->join, 0));
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
  T2_REQUIRE(layer >= 0 && layer < lo.L && reps >= 1 && ms_per_launch && which >= 0 && which <= 4, T2_ERR_INVALID_ARG,
             "time_kernel: bad arguments");
  T2_REQUIRE(which != 1 || layer + 1 < lo.L, T2_ERR_INVALID_ARG, "the last layer has no out GEMM");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  const uint8_t* pk = static_cast<const uint8_t*>(d_packed);
  ActGemmCall g;
  int epi, bn;
  if (which == 0 || which == 4) { g = make_gate_call(lo, ws, pk, layer, which == 0); epi = EPI_GATE; bn = 256; }   // 4: without the tanh / sigmoid stashes
  else if (which == 1) { g = make_out_call(lo, ws, pk, d_params, layer, cfg->dropout, 1, nullptr); epi = EPI_RES; bn = lo.R; }
  else if (which == 2) { g = make_dz_call(lo, ws, pk, layer, nullptr); epi = EPI_GATE_BWD; bn = lo.Gh >= 256 ? 256 : 128; }
  else { g = make_dx_call(lo, ws, pk, layer, cfg->dropout, 1, nullptr, nullptr); epi = EPI_DX; bn = lo.R; }
  // `reps` back-to-back launches captured into ONE CUDA graph on a private stream and replayed, so that the figure is the
  // device-side time per launch (as in the captured training step) and not the host's launch rate
  T2_CHECK_CUDA(cudaStreamSynchronize(st));
  cudaStream_t ps;
  T2_CHECK_CUDA(cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
  cudaEvent_t e0, e1;
  T2_CHECK_CUDA(cudaEventCreate(&e0));
  T2_CHECK_CUDA(cudaEventCreate(&e1));
  rc = launch_act_gemm(epi, bn, g, ps);  // warm-up (also sets the kernel attributes outside the capture)
  if (rc) return rc;
  T2_CHECK_CUDA(cudaStreamSynchronize(ps));
  cudaGraph_t graph;
  cudaGraphExec_t exec;
  T2_CHECK_CUDA(cudaStreamBeginCapture(ps, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < reps && rc == 0; ++i) rc = launch_act_gemm(epi, bn, g, ps);
  cudaError_t ce = cudaStreamEndCapture(ps, &graph);
  if (rc) return rc;
  T2_CHECK_CUDA(ce);
  T2_CHECK_CUDA(cudaGraphInstantiate(&exec, graph, 0));
  T2_CHECK_CUDA(cudaGraphLaunch(exec, ps));   // warm replay
  T2_CHECK_CUDA(cudaEventRecord(e0, ps));
  T2_CHECK_CUDA(cudaGraphLaunch(exec, ps));
  T2_CHECK_CUDA(cudaEventRecord(e1, ps));
  T2_CHECK_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  T2_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  *ms_per_launch = ms / reps;
  cudaGraphExecDestroy(exec);
  cudaGraphDestroy(graph);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaStreamDestroy(ps);
  return T2_OK;
}

// =========================================================================================================
// Fast-WaveNet autoregressive synthesis (wavenet_vocoder/models/wavenet.py:724-911, modules.py:273-303)
// =========================================================================================================
// One persistent kernel generates the whole utterance. A thread-block CLUSTER of CS CTAs owns `NI` batch items;
// every layer's output channels are split across the cluster's CTAs, bf16 weights stream from L2 (the 27.6 MB of
// the paper-width model stay L2-resident), activations live in shared memory as fp32 and are exchanged between
// CTAs through distributed shared memory + one cluster barrier per stage (2 per layer). The reference's
// convolution queues (O(d) concat shift per step, modules.py:285-288) become ring buffers in global memory.
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace t2 {
namespace {

constexpr int kArThreads = 512;
__device__ long long* g_ar_dbg = nullptr;   // tools only: clock64 stamps of (t = 64, l = 7) in CTA 0
#define AR_STAMP(i) do { if (g_ar_dbg && blockIdx.x == 0 && threadIdx.x == 0 && t == 64 && l == 7) g_ar_dbg[i] = clock64(); } while (0)
constexpr int kArMaxItems = 4;

struct ArLayout {
  int CS, ZC, RC, SC, FC, OC, K1;
  long long per_rank_layer;        // bf16 elements per (layer, rank): 2*ZC*K1 + (RC+SC)*Gh
  long long o_head1, o_head2;      // element offsets of the head blocks (after all layers)
  long long n_weights;             // bf16 elements
  long long o_bias;                // byte offset of the fp32 bias block
  long long packed_bytes;
  long long ring_slots_total;      // sum over layers of slots
  long long workspace_bytes, w_cup, w_upout_base, w_ring, w_ringoff;
  std::vector<int> ring_slots, ring_off;
};

int build_ar_layout(const Layout& lo, int CS, ArLayout& a) {
  T2_REQUIRE(CS == 1 || CS == 2 || CS == 4 || CS == 8 || CS == 16, T2_ERR_INVALID_ARG, "cluster size must be 1,2,4,8,16");
  a.CS = CS;
  a.ZC = lo.Gh / CS; a.RC = lo.R / CS; a.SC = lo.S / CS; a.FC = lo.S / CS; a.OC = (lo.O + CS - 1) / CS;
  a.K1 = 3 * lo.R + lo.C;
  a.per_rank_layer = 2LL * a.ZC * a.K1 + (long long)(a.RC + a.SC) * lo.Gh;
  a.o_head1 = a.per_rank_layer * CS * lo.L;
  a.o_head2 = a.o_head1 + (long long)CS * a.FC * lo.S;
  a.n_weights = a.o_head2 + (long long)CS * a.OC * lo.S;
  a.o_bias = align_up(a.n_weights * 2, 256);
  // biases fp32: per layer [G gate | R out], then skip_total [S], f1 [S], f2 [O padded to CS*OC]
  a.packed_bytes = a.o_bias + 4LL * ((long long)lo.L * (lo.G + lo.R) + 2 * lo.S + CS * a.OC);
  a.ring_slots.clear(); a.ring_off.clear();
  long long off = 0;
  for (int l = 0; l < lo.L; ++l) {
    int need = 2 * lo.dil(l) + 1, s = 1;
    while (s < need) s <<= 1;
    a.ring_slots.push_back(s);
    a.ring_off.push_back(int(off));
    off += s;
  }
  a.ring_slots_total = off;
  long long o = 0;
  auto takeb = [&](long long bytes) { long long r = o; o = align_up(o + bytes, 256); return r; };
  const long long BT = (long long)lo.B * lo.T;
  a.w_cup = takeb(BT * (lo.C > 0 ? lo.C : 8) * 2);
  a.w_upout_base = o;
  for (size_t i = 0; i < lo.up_w.size(); ++i) takeb((long long)lo.B * lo.C * lo.up_w[i] * 4);
  a.w_ring = takeb((long long)lo.B * off * lo.R * 4);
  a.w_ringoff = takeb(2LL * lo.L * 4);
  a.workspace_bytes = o;
  return T2_OK;
}

struct ArPackArgs {
  const float* params;
  bf16* w;
  float* bias;
  const long long* offs;   // per layer: dil_k, dil_b, c_k, c_b, s_k, s_b, o_k, o_b  (8 per layer); then f1_k,f1_b,f2_k,f2_b
  const float* skip_scale;
  int L, R, G, Gh, S, C, O, CS, ZC, RC, SC, FC, OC, K1;
  long long per_rank_layer, o_head1, o_head2, n_weights;
};
__global__ void ar_pack_kernel(ArPackArgs a) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < a.n_weights) {
    float v = 0.f;
    if (e < a.o_head1) {
      const int lr = int(e / a.per_rank_layer);
      const int l = lr / a.CS, r = lr % a.CS;
      long long i = e % a.per_rank_layer;
      const long long* o = a.offs + 8 * l;
      if (i < 2LL * a.ZC * a.K1) {
        const int row = int(i / a.K1), k = int(i % a.K1);
        const int ch = (row < a.ZC) ? r * a.ZC + row : a.Gh + r * a.ZC + (row - a.ZC);
        if (k < 3 * a.R) v = a.params[o[0] + (long long)(k / a.R) * a.R * a.G + (long long)(k % a.R) * a.G + ch];
        else v = a.params[o[2] + (long long)(k - 3 * a.R) * a.G + ch];
      } else {
        i -= 2LL * a.ZC * a.K1;
        const int row = int(i / a.Gh), k = int(i % a.Gh);
        if (row < a.RC) v = a.params[o[6] + (long long)k * a.R + r * a.RC + row];
        else v = a.params[o[4] + (long long)k * a.S + r * a.SC + (row - a.RC)] * a.skip_scale[l];
      }
    } else if (e < a.o_head2) {
      const long long i = e - a.o_head1;
      const int row = int(i / a.S), k = int(i % a.S);   // row = r*FC + j == output channel
      v = a.params[a.offs[8 * a.L + 0] + (long long)k * a.S + row];
    } else {
      const long long i = e - a.o_head2;
      const int row = int(i / a.S), k = int(i % a.S);
      if (row < a.O) v = a.params[a.offs[8 * a.L + 2] + (long long)k * a.O + row];
    }
    a.w[e] = __float2bfloat16(v);
  }
  // biases
  const long long nb = (long long)a.L * (a.G + a.R) + 2 * a.S + a.CS * a.OC;
  if (e < nb) {
    float v = 0.f;
    const long long lg = (long long)a.L * (a.G + a.R);
    if (e < lg) {
      const int l = int(e / (a.G + a.R)), j = int(e % (a.G + a.R));
      const long long* o = a.offs + 8 * l;
      if (j < a.G) v = a.params[o[1] + j] + (a.C > 0 ? a.params[o[3] + j] : 0.f);
      else v = a.params[o[7] + (j - a.G)];
    } else if (e < lg + a.S) {
      const int s = int(e - lg);
      for (int l = 0; l < a.L; ++l) v += a.skip_scale[l] * a.params[a.offs[8 * l + 5] + s];
    } else if (e < lg + 2 * a.S) {
      v = a.params[a.offs[8 * a.L + 1] + (e - lg - a.S)];
    } else {
      const int j = int(e - lg - 2 * a.S);
      if (j < a.O) v = a.params[a.offs[8 * a.L + 3] + j];
    }
    a.bias[e] = v;
  }
}

struct ArArgs {
  const bf16* w;            // slice-major weights
  const float* bias;
  const float* in_k;        // input_convolution kernel fp32 [cin][R]
  const float* in_b;
  const bf16* c_up;         // [B][T][C]
  float* ring;              // [B][ring_slots_total][R]
  const int* ring_off;      // [L] then ring_slots [L]
  const void* initial;      // int32 [B] or f32 [B]
  const void* test_inputs;  // nullable: int32 / f32 [B][T]
  const float* u_a;         // MoL: [B][T][nm] mixture-selection uniforms; categorical: [B][T]; nullable
  const float* u_b;         // MoL: [B][T] logistic uniforms; nullable
  void* out_samples;        // int32 / f32 [B][T]
  float* out_raw;           // nullable [B][T][O]
  unsigned long long seed;
  int B, T, L, R, G, Gh, S, C, O, Q, scalar_in, layers_per_stack;
  int CS, ZC, RC, SC, FC, OC, K1;
  long long per_rank_layer, o_head1, o_head2;
  float res_scale, log_scale_min, log_scale_min_gauss;
  int items_per_cluster;
  int prefetch;             // 1: this CTA's per-layer weight slice is double-buffered in shared memory (bulk async copies)
};

// y[it][o] = sum_k W[o][k] * x[it][k] for o in [0, nout): all outputs of a pass run side by side - a group of GS lanes
// (GS = 512 / nout rounded down to a power of two, <= 32) owns one output row and strides its 16-byte chunks, the partial
// sums meet in log2(GS) shuffles. (The warp-per-output form serialised 2 rows x 5 shuffle levels x NI per warp.)
template <int NI>
__device__ __forceinline__ void ar_matvec(const bf16* __restrict__ W, int nout, int K, const float* __restrict__ x /*[NI][ldx]*/,
                                          int ldx, float* __restrict__ y /*[NI][ldy]*/, int ldy, int ni) {
  int gs = 32;
  while (gs > 1 && gs * nout > kArThreads) gs >>= 1;
  const int per_pass = kArThreads / gs;
  const int sub = threadIdx.x % gs;
  const int nchunk = K >> 3;
  for (int o0 = 0; o0 < nout; o0 += per_pass) {
    const int o = o0 + threadIdx.x / gs;
    const bool live = o < nout;
    float acc[NI];
#pragma unroll
    for (int it = 0; it < NI; ++it) acc[it] = 0.f;
    if (live) {
      const uint4* wr = reinterpret_cast<const uint4*>(W + (size_t)o * K);
#pragma unroll 2
      for (int c = sub; c < nchunk; c += gs) {
        const uint4 u = wr[c];          // generic load: the slice is either in L2 (global) or prefetched into shared memory
        const float w0 = bf16lo(u.x), w1 = bf16hi(u.x), w2 = bf16lo(u.y), w3 = bf16hi(u.y);
        const float w4 = bf16lo(u.z), w5 = bf16hi(u.z), w6 = bf16lo(u.w), w7 = bf16hi(u.w);
#pragma unroll
        for (int it = 0; it < NI; ++it) {
          if (it < ni) {
            const float4 p = *reinterpret_cast<const float4*>(x + it * ldx + c * 8);
            const float4 q = *reinterpret_cast<const float4*>(x + it * ldx + c * 8 + 4);
            acc[it] += w0 * p.x + w1 * p.y + w2 * p.z + w3 * p.w + w4 * q.x + w5 * q.y + w6 * q.z + w7 * q.w;
          }
        }
      }
    }
#pragma unroll
    for (int it = 0; it < NI; ++it) {
      if (it < ni) {      // block-uniform
        float v = acc[it];
        for (int m = gs >> 1; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        if (live && sub == 0) y[it * ldy + o] = v;
      }
    }
  }
}

template <int NI>
__global__ void __launch_bounds__(kArThreads, 1) wn_ar_kernel(ArArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = int(cluster.block_rank());
  const int cid = blockIdx.x / a.CS;
  const int item0 = cid * a.items_per_cluster;
  int ni = a.B - item0;
  if (ni > a.items_per_cluster) ni = a.items_per_cluster;
  if (ni < 0) ni = 0;   // surplus clusters still take part in no barriers of other clusters; they just idle through
  extern __shared__ __align__(16) float sm[];
  const int ld1 = (a.K1 + 3) & ~3;
  const int nbs = 2 * a.ZC + a.RC;                  // bias slice per layer: a rows | b rows | residual-out rows
  float* in1 = sm;                                  // [NI][ld1]  : x(t-2d) | x(t-d) | x(t) | c(t)
  float* zbuf = in1 + NI * ld1;                     // [NI][Gh]   : full z vector (pulled from the cluster)
  float* xbuf = zbuf + NI * a.Gh;                   // [NI][R]    : current layer input (full vector)
  float* zsl = xbuf + NI * a.R;                     // [NI][ZC]   : this CTA's z slice, read remotely by the cluster
  float* xsl = zsl + NI * a.ZC;                     // [NI][RC]   : this CTA's slice of the next layer input, read remotely
  float* loc = xsl + NI * a.RC;                     // [NI][2*ZC] : this CTA's gate pre-activations / stage-2 outputs
  float* skip = loc + NI * (2 * a.ZC > a.RC + a.SC ? 2 * a.ZC : a.RC + a.SC);   // [NI][SC] running skip sum (slice)
  float* hbuf = skip + NI * a.SC;                   // [NI][S]    : head activations (full vector)
  float* obuf = hbuf + NI * a.S;                    // [NI][CS*OC]: network output (full vector)
  float* bsl = obuf + NI * a.CS * a.OC;             // [L][nbs]   : this CTA's bias slices (cluster.sync flushes L1: keep them here)
  float* cvec = bsl + a.L * nbs;                    // [NI][C]    : conditioning frame of the current step
  int* rofs = reinterpret_cast<int*>(cvec + NI * ((a.C + 3) & ~3));   // [2L+1] ring offsets / slots / total
  float* cur = reinterpret_cast<float*>(rofs + ((2 * a.L + 1 + 3) & ~3));   // [NI] current input sample (scalar) or index
  // weight prefetch: two slots of per_rank_layer bf16 + two mbarriers behind the activation buffers (16-byte aligned)
  uint64_t* wbar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(cur + NI) + 15) & ~uintptr_t(15));
  bf16* wslot = reinterpret_cast<bf16*>(wbar + 2);
  const uint32_t wbytes = uint32_t(a.per_rank_layer * 2);
  const int tid = threadIdx.x;
  if (a.prefetch && tid == 0) {
    mbar_init(&wbar[0], 1); mbar_init(&wbar[1], 1);
    fence_barrier_init();
  }
  const float* bias_all = a.bias;
  const float* b_skip = bias_all + (long long)a.L * (a.G + a.R);
  const float* b_f1 = b_skip + a.S;
  const float* b_f2 = b_f1 + a.S;
  const int nm = a.O / 3;

  for (int i = tid; i < a.L * nbs; i += kArThreads) {
    const int l = i / nbs, j = i % nbs;
    const float* bgl = bias_all + (long long)l * (a.G + a.R);
    bsl[i] = j < a.ZC ? bgl[rank * a.ZC + j] : (j < 2 * a.ZC ? bgl[a.Gh + rank * a.ZC + (j - a.ZC)] : bgl[a.G + rank * a.RC + (j - 2 * a.ZC)]);
  }
  for (int i = tid; i < 2 * a.L + 1; i += kArThreads) rofs[i] = a.ring_off[i];
  for (int i = tid; i < NI; i += kArThreads)
    if (i < ni) cur[i] = a.scalar_in ? static_cast<const float*>(a.initial)[item0 + i]
                                     : float(static_cast<const int*>(a.initial)[item0 + i]);
  __syncthreads();
  if (a.prefetch && tid == 0) {   // slice of layer 0 -> slot 0
    mbar_expect_tx(&wbar[0], wbytes);
    bulk_load_1d(wslot, a.w + (long long)rank * a.per_rank_layer, wbytes, &wbar[0]);
  }
  uint32_t wphase = 0;            // bit s = parity to wait for on slot s
  long long seq = 0;              // (t, l) sequence number: slot = seq & 1
  // ring taps of the NEXT layer are fetched into registers one layer ahead (their producers ran >= one time step ago)
  constexpr int kTapRegs = (NI * 2 * 512 + kArThreads - 1) / kArThreads;   // R <= 512
  float tapv[kTapRegs];
  auto fetch_taps = [&](int tt0, int l) {
    const int d = 1 << (l % a.layers_per_stack);
    const int slots = rofs[a.L + l];
    const long long roff = rofs[l];
#pragma unroll
    for (int j = 0; j < kTapRegs; ++j) {
      const int i = tid + j * kArThreads;
      float v = 0.f;
      if (i < ni * 2 * a.R) {
        const int it = i / (2 * a.R), k = i % (2 * a.R);
        const int tap = k / a.R, r = k % a.R;
        const int tt = tt0 - (2 - tap) * d;
        if (tt >= 0) v = __ldcg(a.ring + ((long long)(item0 + it) * rofs[2 * a.L] + roff + (tt & (slots - 1))) * a.R + r);
      }
      tapv[j] = v;
    }
  };
  fetch_taps(0, 0);

  for (int t = 0; t < a.T; ++t) {
    // ---- first conv: x0 = W_in[idx] + b (one-hot) or x * w + b (scalar); every CTA builds the full vector ----
    for (int i = tid; i < ni * a.R; i += kArThreads) {
      const int it = i / a.R, r = i % a.R;
      float v;
      if (a.scalar_in) v = cur[it] * a.in_k[r] + a.in_b[r];
      else v = a.in_k[(long long)int(cur[it]) * a.R + r] + a.in_b[r];
      xbuf[it * a.R + r] = v;
    }
    for (int i = tid; i < ni * a.SC; i += kArThreads) skip[i] = 0.f;
    for (int i = tid; i < ni * a.C; i += kArThreads)     // conditioning frame of this step, once (not once per layer)
      cvec[(i / a.C) * ((a.C + 3) & ~3) + i % a.C] = __bfloat162float(a.c_up[((long long)(item0 + i / a.C) * a.T + t) * a.C + i % a.C]);
    __syncthreads();
    for (int l = 0; l < a.L; ++l) {
      const int slots = rofs[a.L + l];
      const long long roff = rofs[l];
      AR_STAMP(0);
      // ---- gather the stage-1 input: taps (prefetched registers), current x, conditioning ----
#pragma unroll
      for (int j = 0; j < kTapRegs; ++j) {
        const int i = tid + j * kArThreads;
        if (i < ni * 2 * a.R) in1[(i / (2 * a.R)) * ld1 + i % (2 * a.R)] = tapv[j];
      }
      for (int i = tid; i < ni * (ld1 - 2 * a.R); i += kArThreads) {
        const int it = i / (ld1 - 2 * a.R), k = 2 * a.R + i % (ld1 - 2 * a.R);
        float v = 0.f;
        if (k < 3 * a.R) v = xbuf[it * a.R + (k - 2 * a.R)];
        else if (k < a.K1) v = cvec[it * ((a.C + 3) & ~3) + (k - 3 * a.R)];
        in1[it * ld1 + k] = v;
      }
      __syncthreads();
      AR_STAMP(1);
      // rank 0 publishes x_l(t) into the ring for later steps (read back no earlier than step t + d)
      if (rank == 0)
        for (int i = tid; i < ni * a.R; i += kArThreads) {
          const int it = i / a.R, r = i % a.R;
          a.ring[((long long)(item0 + it) * rofs[2 * a.L] + roff + (t & (slots - 1))) * a.R + r] = xbuf[i];
        }
      // ---- stage 1: gate pre-activations for this CTA's ZC z-channels (a rows then b rows) ----
      const bf16* w1 = a.w + ((long long)l * a.CS + rank) * a.per_rank_layer;
      if (a.prefetch) {
        const int slot = int(seq & 1);
        mbar_wait(&wbar[slot], (wphase >> slot) & 1u);       // this layer's slice has landed
        wphase ^= 1u << slot;
        w1 = wslot + (long long)slot * a.per_rank_layer;
        if (tid == 0 && (t + 1 < a.T || l + 1 < a.L)) {      // next layer's slice -> the other slot (free since the last cluster.sync)
          const int ln = l + 1 < a.L ? l + 1 : 0;
          mbar_expect_tx(&wbar[slot ^ 1], wbytes);
          bulk_load_1d(wslot + (long long)(slot ^ 1) * a.per_rank_layer, a.w + ((long long)ln * a.CS + rank) * a.per_rank_layer, wbytes,
                       &wbar[slot ^ 1]);
        }
      }
      ++seq;
      AR_STAMP(2);
      ar_matvec<NI>(w1, 2 * a.ZC, a.K1, in1, ld1, loc, 2 * a.ZC, ni);
      __syncthreads();
      AR_STAMP(3);
      // taps of the next layer (or of layer 0 at the next time step): issued now, consumed after two cluster barriers
      if (l + 1 < a.L) fetch_taps(t, l + 1);
      else if (t + 1 < a.T) fetch_taps(t + 1, 0);
      const float* bg = bsl + l * nbs;
      for (int i = tid; i < ni * a.ZC; i += kArThreads) {
        const int it = i / a.ZC, j = i % a.ZC;
        const float av = loc[it * 2 * a.ZC + j] + bg[j];
        const float bv = loc[it * 2 * a.ZC + a.ZC + j] + bg[a.ZC + j];
        zsl[it * a.ZC + j] = tanhf_(av) * sigmoidf_(bv);      // local slice; the cluster PULLS it after the barrier
      }
      AR_STAMP(4);
      cluster.sync();
      AR_STAMP(5);
      // pull the full z vector: 16-byte pieces from the owners' slices through distributed shared memory
      for (int i = tid; i < ni * (a.Gh >> 2); i += kArThreads) {
        const int it = i / (a.Gh >> 2), ch = (i % (a.Gh >> 2)) << 2;
        const int r = ch / a.ZC, j = ch % a.ZC;
        *reinterpret_cast<float4*>(zbuf + it * a.Gh + ch) = *reinterpret_cast<const float4*>(cluster.map_shared_rank(zsl, r) + it * a.ZC + j);
      }
      __syncthreads();
      // ---- stage 2: this CTA's RC residual-out channels and SC skip channels ----
      const bf16* w2 = w1 + 2LL * a.ZC * a.K1;
      ar_matvec<NI>(w2, a.RC + a.SC, a.Gh, zbuf, a.Gh, loc, a.RC + a.SC, ni);
      __syncthreads();
      AR_STAMP(6);
      for (int i = tid; i < ni * (a.RC + a.SC); i += kArThreads) {
        const int it = i / (a.RC + a.SC), j = i % (a.RC + a.SC);
        const float v = loc[it * (a.RC + a.SC) + j];
        if (j < a.RC) xsl[it * a.RC + j] = (v + bg[2 * a.ZC + j] + xbuf[it * a.R + rank * a.RC + j]) * a.res_scale;
        else skip[it * a.SC + (j - a.RC)] += v;   // scale_l folded into the packed skip weights
      }
      AR_STAMP(7);
      cluster.sync();
      AR_STAMP(8);
      if (l + 1 < a.L) {
        for (int i = tid; i < ni * (a.R >> 2); i += kArThreads) {
          const int it = i / (a.R >> 2), ch = (i % (a.R >> 2)) << 2;
          const int r = ch / a.RC, j = ch % a.RC;
          *reinterpret_cast<float4*>(xbuf + it * a.R + ch) = *reinterpret_cast<const float4*>(cluster.map_shared_rank(xsl, r) + it * a.RC + j);
        }
        __syncthreads();
      }
      AR_STAMP(9);
    }
    // ---- head: relu(skips + b) -> f1 -> relu -> f2 ----
    for (int i = tid; i < ni * a.SC; i += kArThreads) {
      const int it = i / a.SC, j = i % a.SC, ch = rank * a.SC + j;
      const float v = fmaxf(skip[i] + b_skip[ch], 0.f);
      for (int r = 0; r < a.CS; ++r) cluster.map_shared_rank(hbuf, r)[it * a.S + ch] = v;
    }
    cluster.sync();
    ar_matvec<NI>(a.w + a.o_head1 + (long long)rank * a.FC * a.S, a.FC, a.S, hbuf, a.S, loc, a.FC, ni);
    __syncthreads();
    cluster.sync();   // every CTA has finished reading hbuf (h1) before it is overwritten with h2
    for (int i = tid; i < ni * a.FC; i += kArThreads) {
      const int it = i / a.FC, j = i % a.FC, ch = rank * a.FC + j;
      const float v = fmaxf(loc[it * a.FC + j] + b_f1[ch], 0.f);
      for (int r = 0; r < a.CS; ++r) cluster.map_shared_rank(hbuf, r)[it * a.S + ch] = v;
    }
    cluster.sync();
    ar_matvec<NI>(a.w + a.o_head2 + (long long)rank * a.OC * a.S, a.OC, a.S, hbuf, a.S, loc, a.OC, ni);
    __syncthreads();
    for (int i = tid; i < ni * a.OC; i += kArThreads) {
      const int it = i / a.OC, j = i % a.OC, ch = rank * a.OC + j;
      const float v = loc[it * a.OC + j] + b_f2[ch];
      for (int r = 0; r < a.CS; ++r) cluster.map_shared_rank(obuf, r)[it * a.CS * a.OC + ch] = v;
    }
    cluster.sync();
    // ---- sample (identically in every CTA: same inputs, same uniforms) ----
    if (tid < 32 * NI) {
      const int it = tid >> 5, lane = tid & 31;
      if (it < ni) {
        const int bi = item0 + it;
        const float* y = obuf + it * a.CS * a.OC;
        if (rank == 0 && a.out_raw)
          for (int j = lane; j < a.O; j += 32) a.out_raw[((long long)bi * a.T + t) * a.O + j] = y[j];
        float nxt;
        if (a.scalar_in && a.O == 2) {
          // sample_from_gaussian (gaussian.py:39-52): x = mean + exp(max(log_scale, min)) * n, clipped to [-1, 1]; the standard-normal
          // draw n is injected through u_b or made by Box-Muller from two counter-hash uniforms
          float n;
          if (a.u_b) n = a.u_b[(long long)bi * a.T + t];
          else {
            const float u1 = 1e-7f + (1.f - 2e-7f) * hash_uniform(a.seed, ((unsigned long long)bi * a.T + t) * 16 + 14);
            const float u2 = hash_uniform(a.seed, ((unsigned long long)bi * a.T + t) * 16 + 15);
            n = sqrtf(-2.f * __logf(u1)) * __cosf(6.28318530718f * u2);
          }
          const float x = y[0] + __expf(fmaxf(y[1], a.log_scale_min_gauss)) * n;
          nxt = fminf(fmaxf(x, -1.f), 1.f);
          if (rank == 0 && lane == 0) static_cast<float*>(a.out_samples)[(long long)bi * a.T + t] = nxt;
        } else if (a.scalar_in) {
          // sample_from_discretized_mix_logistic (mixture.py:76-107): Gumbel-max over the mixture logits
          float best = -INFINITY;
          int bk = 0;
          for (int k = 0; k < nm; ++k) {
            float u = a.u_a ? a.u_a[((long long)bi * a.T + t) * nm + k]
                            : 1e-5f + (1.f - 2e-5f) * hash_uniform(a.seed, ((unsigned long long)bi * a.T + t) * 16 + k);
            const float g = y[k] - __logf(-__logf(u));
            if (g > best) { best = g; bk = k; }
          }
          const float mean = y[nm + bk];
          const float ls = fmaxf(y[2 * nm + bk], a.log_scale_min);
          const float u = a.u_b ? a.u_b[(long long)bi * a.T + t]
                                : 1e-5f + (1.f - 2e-5f) * hash_uniform(a.seed, ((unsigned long long)bi * a.T + t) * 16 + 15);
          float x = mean + __expf(ls) * (__logf(u) - __logf(1.f - u));
          nxt = fminf(fmaxf(x, -1.f), 1.f);
          if (rank == 0 && lane == 0) static_cast<float*>(a.out_samples)[(long long)bi * a.T + t] = nxt;
        } else {
          // categorical sample from softmax(logits) by inverse CDF (tf.multinomial on raw logits, wavenet.py:865)
          float mx = -INFINITY;
          for (int j = lane; j < a.O; j += 32) mx = fmaxf(mx, y[j]);
          mx = warp_max(mx);
          float se = 0.f;
          for (int j = lane; j < a.O; j += 32) se += __expf(y[j] - mx);
          se = warp_sum(se);
          const float u = (a.u_a ? a.u_a[(long long)bi * a.T + t]
                                 : hash_uniform(a.seed, (unsigned long long)bi * a.T + t)) * se;
          // lane-blocked scan: each lane owns O/32 consecutive classes
          const int per = (a.O + 31) / 32;
          float part = 0.f;
          for (int j = lane * per; j < (lane + 1) * per && j < a.O; ++j) part += __expf(y[j] - mx);
          float incl = part;
          for (int o = 1; o < 32; o <<= 1) {
            const float nb = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += nb;
          }
          const float excl = incl - part;
          int pick = -1;
          if (u >= excl && u < incl) {
            float c = excl;
            pick = lane * per;
            for (int j = lane * per; j < (lane + 1) * per && j < a.O; ++j) { c += __expf(y[j] - mx); pick = j; if (c > u) break; }
          }
          pick = __reduce_max_sync(0xffffffffu, pick);
          if (pick < 0) pick = a.O - 1;
          nxt = float(pick);
          if (rank == 0 && lane == 0) static_cast<int*>(a.out_samples)[(long long)bi * a.T + t] = pick;
        }
        if (a.test_inputs)
          nxt = a.scalar_in ? static_cast<const float*>(a.test_inputs)[(long long)bi * a.T + t]
                            : float(static_cast<const int*>(a.test_inputs)[(long long)bi * a.T + t]);
        if (lane == 0) cur[it] = nxt;
      }
    }
    __syncthreads();
  }
}

}  // namespace
}  // namespace t2

typedef struct {
  long long packed_bytes, workspace_bytes;
} t2_wn_ar_sizes_t_;

// tools only: device buffer of 16 int64 receiving clock64() stamps of one AR layer pass (NULL = off)
extern "C" int t2_dbg_ar_stamps(long long* d_buf) {
  T2_CHECK_CUDA(cudaMemcpyToSymbol(t2::g_ar_dbg, &d_buf, sizeof(d_buf)));
  return T2_OK;
}

extern "C" int t2_wn_ar_sizes(const t2_wn_config_t* cfg, int cluster_size, long long* packed_bytes, long long* workspace_bytes) {
  Layout lo;
  int rc = build_layout(cfg, lo);
  if (rc) return rc;
  ArLayout a;
  rc = build_ar_layout(lo, cluster_size, a);
  if (rc) return rc;
  T2_REQUIRE(lo.Gh % cluster_size == 0 && lo.R % cluster_size == 0 && lo.S % cluster_size == 0 && lo.R <= lo.S + 0,
             T2_ERR_UNSUPPORTED_SHAPE, "channel counts must divide by the cluster size and R <= S");
  *packed_bytes = a.packed_bytes;
  *workspace_bytes = a.workspace_bytes;
  return T2_OK;
}

extern "C" int t2_wn_ar_pack(const t2_wn_config_t* cfg, int cluster_size, const float* d_params, void* d_packed_ar,
                             void* d_workspace, void* stream) {
  Layout lo;
  int rc = build_layout(cfg, lo);
  if (rc) return rc;
  ArLayout a;
  rc = build_ar_layout(lo, cluster_size, a);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  std::vector<long long> offs(8 * lo.L + 4);
  for (int l = 0; l < lo.L; ++l) {
    long long* o = &offs[8 * l];
    o[0] = lo.p_dil_k[l]; o[1] = lo.p_dil_b[l]; o[2] = lo.C > 0 ? lo.p_c_k[l] : 0; o[3] = lo.C > 0 ? lo.p_c_b[l] : 0;
    o[4] = lo.p_s_k[l]; o[5] = lo.p_s_b[l]; o[6] = lo.p_o_k[l]; o[7] = lo.p_o_b[l];
  }
  offs[8 * lo.L] = lo.p_f1_k; offs[8 * lo.L + 1] = lo.p_f1_b; offs[8 * lo.L + 2] = lo.p_f2_k; offs[8 * lo.L + 3] = lo.p_f2_b;
  // tables go through a temporary device allocation (this entry point synchronises; it runs once per checkpoint)
  long long* d_offs = nullptr;
  float* d_scale = nullptr;
  T2_CHECK_CUDA(cudaMallocAsync(&d_offs, offs.size() * sizeof(long long), st));
  T2_CHECK_CUDA(cudaMallocAsync(&d_scale, lo.L * sizeof(float), st));
  T2_CHECK_CUDA(cudaMemcpyAsync(d_offs, offs.data(), offs.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
  T2_CHECK_CUDA(cudaMemcpyAsync(d_scale, lo.skip_scale.data(), lo.L * sizeof(float), cudaMemcpyHostToDevice, st));
  ArPackArgs p;
  p.params = d_params; p.w = static_cast<bf16*>(d_packed_ar);
  p.bias = reinterpret_cast<float*>(static_cast<uint8_t*>(d_packed_ar) + a.o_bias);
  p.offs = d_offs; p.skip_scale = d_scale;
  p.L = lo.L; p.R = lo.R; p.G = lo.G; p.Gh = lo.Gh; p.S = lo.S; p.C = lo.C; p.O = lo.O; p.CS = a.CS; p.ZC = a.ZC; p.RC = a.RC;
  p.SC = a.SC; p.FC = a.FC; p.OC = a.OC; p.K1 = a.K1; p.per_rank_layer = a.per_rank_layer; p.o_head1 = a.o_head1;
  p.o_head2 = a.o_head2; p.n_weights = a.n_weights;
  ar_pack_kernel<<<grid1d(a.n_weights), 256, 0, st>>>(p); t2_count_launch();
  T2_CHECK_CUDA(cudaGetLastError());
  // ring tables
  std::vector<int> rt(2 * lo.L + 1);
  for (int l = 0; l < lo.L; ++l) { rt[l] = a.ring_off[l]; rt[lo.L + l] = a.ring_slots[l]; }
  rt[2 * lo.L] = int(a.ring_slots_total);
  T2_CHECK_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(d_workspace) + a.w_ringoff, rt.data(), rt.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  T2_CHECK_CUDA(cudaStreamSynchronize(st));
  cudaFreeAsync(d_offs, st);
  cudaFreeAsync(d_scale, st);
  return T2_OK;
}

extern "C" int t2_wn_ar_generate(const t2_wn_config_t* cfg, int cluster_size, const float* d_params, const void* d_packed_ar,
                                 void* d_workspace, const float* d_c, const void* d_initial, const void* d_test_inputs,
                                 const float* d_u_a, const float* d_u_b, unsigned long long seed, void* d_out_samples,
                                 float* d_out_raw, void* stream) {
  Layout lo;
  int rc = build_layout(cfg, lo);
  if (rc) return rc;
  ArLayout al;
  rc = build_ar_layout(lo, cluster_size, al);
  if (rc) return rc;
  T2_REQUIRE(lo.C > 0, T2_ERR_UNSUPPORTED_SHAPE, "AR synthesis needs local conditioning");
  T2_REQUIRE(lo.R <= lo.S, T2_ERR_UNSUPPORTED_SHAPE, "AR synthesis needs residual_channels <= skip_out_channels");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  const uint8_t* pk = static_cast<const uint8_t*>(d_packed_ar);
  const long long BT = (long long)lo.B * lo.T;
  // conditioning -> c_up (bf16 channels-last), same kernels as the training path
  bf16* c_up = reinterpret_cast<bf16*>(ws + al.w_cup);
  if (cfg->c_pre_upsampled) {
    f32_to_bf16_kernel<<<grid1d(BT * lo.C), 256, 0, st>>>(d_c, c_up, BT * lo.C); t2_count_launch();
  } else {
    const float* in = d_c;
    int W = lo.Tc;
    long long o = al.w_upout_base;
    for (size_t i = 0; i < lo.up_w.size(); ++i) {
      const int s = cfg->upsample_scales[i];
      float* out = reinterpret_cast<float*>(ws + o);
      o = align_up(o + (long long)lo.B * lo.C * lo.up_w[i] * 4, 256);
      const bool last = i + 1 == lo.up_w.size();
      upsample_fwd_kernel<<<grid1d((long long)lo.B * lo.C * W * s), 256, 0, st>>>(
          in, d_params + lo.p_up_k[i], d_params + lo.p_up_b[i], out, last ? c_up : nullptr, lo.B, lo.C, W, s, cfg->upsample_type);
      t2_count_launch();
      in = out;
      W *= s;
    }
  }
  T2_CHECK_CUDA(cudaGetLastError());
  ArArgs a;
  memset(&a, 0, sizeof(a));
  a.w = reinterpret_cast<const bf16*>(pk);
  a.bias = reinterpret_cast<const float*>(pk + al.o_bias);
  a.in_k = d_params + lo.p_in_k; a.in_b = d_params + lo.p_in_b;
  a.c_up = c_up;
  a.ring = reinterpret_cast<float*>(ws + al.w_ring);
  a.ring_off = reinterpret_cast<const int*>(ws + al.w_ringoff);
  a.initial = d_initial; a.test_inputs = d_test_inputs; a.u_a = d_u_a; a.u_b = d_u_b;
  a.out_samples = d_out_samples; a.out_raw = d_out_raw; a.seed = seed;
  a.B = lo.B; a.T = lo.T; a.L = lo.L; a.R = lo.R; a.G = lo.G; a.Gh = lo.Gh; a.S = lo.S; a.C = lo.C; a.O = lo.O; a.Q = lo.Q;
  a.scalar_in = lo.scalar_in ? 1 : 0; a.layers_per_stack = lo.L / cfg->stacks;
  a.CS = al.CS; a.ZC = al.ZC; a.RC = al.RC; a.SC = al.SC; a.FC = al.FC; a.OC = al.OC; a.K1 = al.K1;
  a.per_rank_layer = al.per_rank_layer; a.o_head1 = al.o_head1; a.o_head2 = al.o_head2;
  a.res_scale = lo.res_scale; a.log_scale_min = cfg->log_scale_min; a.log_scale_min_gauss = cfg->log_scale_min_gauss;
  // clusters: as many as fit on the device, but never more than batch items
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int n_clusters = sms / al.CS;
  if (n_clusters > lo.B) n_clusters = lo.B;
  if (n_clusters < 1) n_clusters = 1;
  int ipc = (lo.B + n_clusters - 1) / n_clusters;
  while (ipc > kArMaxItems) { ++n_clusters; ipc = (lo.B + n_clusters - 1) / n_clusters; }  // more clusters than fit run in waves
  n_clusters = (lo.B + ipc - 1) / ipc;
  a.items_per_cluster = ipc;
  const int ld1 = (al.K1 + 3) & ~3;
  const int locw = 2 * al.ZC > al.RC + al.SC ? 2 * al.ZC : al.RC + al.SC;
  T2_REQUIRE(al.ZC % 4 == 0 && al.RC % 4 == 0, T2_ERR_UNSUPPORTED_SHAPE, "AR synthesis: channel slices per CTA must be multiples of 4");
  const int NIt = ipc <= 1 ? 1 : (ipc <= 2 ? 2 : kArMaxItems);     // kernel instantiation: items per cluster pass
  size_t smem = sizeof(float) * (size_t(NIt) * (ld1 + lo.Gh + lo.R + al.ZC + al.RC + locw + al.SC + lo.S + al.CS * al.OC + ((lo.C + 3) & ~3) + 1) +
                                 size_t(lo.L) * (2 * al.ZC + al.RC) + ((2 * lo.L + 1 + 3) & ~3)) + 64;
  // double-buffered shared-memory copy of this CTA's per-layer weight slice when it fits next to the activations
  // (paper widths: 70.6 KB per slice at cluster size 16); otherwise the slices stream from L2 as before
  const size_t wslots = 2 * size_t(al.per_rank_layer) * 2 + 64;
  a.prefetch = (al.per_rank_layer % 8 == 0 && smem + wslots <= 232448 - 1024) ? 1 : 0;
  if (const char* e = getenv("T2_AR_PREFETCH")) { if (e[0] == '0') a.prefetch = 0; }
  if (a.prefetch) smem += wslots;
  void (*kern)(ArArgs) = NIt == 1 ? wn_ar_kernel<1> : (NIt == 2 ? wn_ar_kernel<2> : wn_ar_kernel<kArMaxItems>);
  T2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  if (al.CS > 8) T2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  T2_CHECK_CUDA(cudaMemsetAsync(ws + al.w_ring, 0, (size_t)lo.B * al.ring_slots_total * lo.R * 4, st));
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.gridDim = dim3(n_clusters * al.CS);
  lc.blockDim = dim3(kArThreads);
  lc.dynamicSmemBytes = smem;
  lc.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = al.CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = 1;
  T2_CHECK_CUDA(cudaLaunchKernelEx(&lc, kern, a));
  t2_count_launch();
  return T2_OK;
}
