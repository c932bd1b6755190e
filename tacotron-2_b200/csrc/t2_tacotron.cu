// t2_tacotron.cu — Tacotron-2 mel-spectrogram predictor (training graph, teacher forcing, r = 1) on the GEMM engine.
//
// Replaces tacotron/models/tacotron.py:104-200 (graph assembly), :315-354 (losses), modules.py (conv1d / BN blocks,
// ZoneoutLSTMCell, Prenet, projections, Postnet), attention.py (location-sensitive attention) and
// Architecture_wrappers.py:169-213 (decoder step order) of the reference.
//
// Mapping onto the B200:
//   * everything that is batched over time runs on act_gemm_kernel: the k=5 'same' convolutions are 5 row-shifted
//     K-segments (zero padding = TMA out-of-bounds fill), prenet / LSTM input projections / frame+stop projections /
//     attention keys are plain 1x1 GEMMs; weight gradients of all of them go through wgrad_gemm_kernel.
//   * the recurrences (BiLSTM encoder, 2-layer decoder LSTM) use the SWAPPED GEMM: the permuted recurrent weight
//     matrix is the 128-row M operand, the batch is N = 32, and the LSTM cell + zoneout is the epilogue (EPI_LSTM).
//     Backward-through-time uses the transposed weights the same way (EPI_TOUT) and stashes gate gradients so that
//     every recurrent weight gradient is ONE wgrad GEMM over all time steps afterwards.
//   * one attention CTA per batch item per step (query projection, location conv, energies, masked softmax, context).
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
inline long long al256(long long v) { return (v + 255) / 256 * 256; }
inline dim3 g1(long long n, int block = 256) { return dim3((unsigned)((n + block - 1) / block)); }

struct PT { std::string name; long long off; int ndim; int shape[4]; bool trainable, reg; };

struct ConvL {   // one conv + BN block
  int cin, cout, k, act;   // act: 1 relu, 2 tanh, 0 none
  long long p_k, p_b, p_gamma, p_beta, p_mm, p_mv;
  long long k_w, k_wT;     // packed forward [cout][k*cinp], packed dgrad [cinp][k*cout]
  int cinp;                // cin rounded up to 64
  long long w_y, w_x;      // workspace: y (post-activation, pre-BN) bf16 [B][T][cout], x = block output bf16
  long long w_stats;       // fp32 [4][cout]: sum / sumsq (fwd) then mean / rstd ; [2][cout] bwd sums
  int stream;              // dropout hash stream
};

struct TL {  // layout
  t2_taco_config_t c;
  int B, Ti, To, E, C, H, D, A, F, KA, P1, P2, M, PC, NS;
  std::vector<PT> params;
  long long n_params;
  long long p_emb, p_elk[2], p_elb[2], p_mem, p_qry, p_lck, p_lcb, p_lfl, p_v, p_ba, p_p1k, p_p1b, p_p2k, p_p2b;
  long long p_l1k, p_l1b, p_l2k, p_l2b, p_fk, p_fb, p_sk, p_sb, p_ppk, p_ppb;
  std::vector<ConvL> enc, post;
  // packed (bytes)
  long long k_encWx[2], k_encWr[2], k_encWrT[2], k_encWxT, k_mem, k_memT, k_p1, k_p1T, k_p2, k_p2T, k_l1x, k_l1xT, k_l1r, k_l1rT;
  long long k_l2, k_l2T, k_proj, k_projT, k_pp, k_ppT, k_qT, packed_bytes;
  // workspace (bytes)
  long long w_emb, w_encpre[2], w_ench[2], w_encc[2], w_encg[2], w_enct[2], w_memory, w_values, w_keys, w_decin, w_pn1, w_pn2;
  long long w_pre1, w_S1, w_S2, w_PI, w_c1, w_c2, w_g1, w_g2, w_t1, w_t2, w_cum, w_alpha, w_projo, w_decbm, w_decf, w_stop;
  long long w_resid, w_mel, w_scal, w_zero, w_tlen;
  // backward
  long long w_dmel, w_dY, w_ddec_tm, w_dPI, w_dh1ext, w_dh2ext, w_dhs1, w_dhs2, w_dcs1, w_dcs2, w_dg1, w_dg2, w_dgstep;
  long long w_dctxl, w_dctx_all, w_dq_all, w_dcum, w_cumrun, w_dkeys, w_dvalues, w_attacc, w_dpn2, w_dpn1;
  long long w_dencpre[2], w_dx3, w_encdh[2], w_encdc[2], w_encdg, w_demb, w_tiles, w_packjobs, w_regtab;
  long long w_ddecf, w_encdgall[2], w_dkeysb, w_dz, w_attU;
  std::vector<int> tile_off, tile_cnt;  // per wgrad launch (fixed order, see build_tiles)
  long long workspace_bytes;
  int n_packjobs, n_reg;
};

struct PJ { long long src_off; int K, N; long long dst_off; int dst_ld, transpose, col0; float scale; int perm_h; int part; };   // part 2: bf16(w - bf16(w))

long long addp(TL& lo, const std::string& name, std::initializer_list<int> shape, bool trainable = true) {
  PT p; p.name = name; p.off = lo.n_params; p.ndim = int(shape.size());
  long long n = 1; int i = 0;
  for (int s : shape) { p.shape[i++] = s; n *= s; }
  for (; i < 4; ++i) p.shape[i] = 1;
  p.trainable = trainable;
  p.reg = trainable && name.find("bias") == std::string::npos && name.find("_projection") == std::string::npos &&
          name.find("inputs_embedding") == std::string::npos && name.find("LSTM") == std::string::npos;
  lo.n_params += (n + 3) / 4 * 4;
  lo.params.push_back(p);
  return p.off;
}

int build(const t2_taco_config_t* cfg, TL& lo, std::vector<PJ>* jobs_out) {
  T2_REQUIRE(cfg != nullptr, T2_ERR_INVALID_ARG, "null config");
  lo.c = *cfg;
  lo.B = cfg->B; lo.Ti = cfg->T_in; lo.To = cfg->T_out; lo.E = cfg->embedding_dim; lo.C = cfg->enc_conv_channels;
  lo.H = cfg->encoder_lstm_units; lo.D = cfg->decoder_lstm_units; lo.A = cfg->attention_dim; lo.F = cfg->attention_filters;
  lo.KA = cfg->attention_kernel; lo.P1 = cfg->prenet1; lo.P2 = cfg->prenet2; lo.M = cfg->num_mels; lo.PC = cfg->postnet_channels;
  lo.NS = cfg->n_symbols;
  T2_REQUIRE(lo.B >= 1 && lo.B <= 256 && lo.Ti >= 1 && lo.To >= 1, T2_ERR_INVALID_ARG, "bad B / T_in / T_out");
  T2_REQUIRE(lo.E % 64 == 0 && lo.C % 64 == 0 && lo.PC % 64 == 0 && lo.P1 % 64 == 0 && lo.P2 % 64 == 0, T2_ERR_UNSUPPORTED_SHAPE,
             "channel counts must be multiples of 64");
  T2_REQUIRE(lo.H % 32 == 0 && lo.D % 32 == 0 && (4 * lo.H) % 128 == 0 && (2 * lo.H) % 64 == 0, T2_ERR_UNSUPPORTED_SHAPE, "LSTM sizes");
  T2_REQUIRE(lo.A % 64 == 0 && lo.A <= 128 && lo.F <= 32 && lo.KA % 2 == 1 && lo.KA <= 31, T2_ERR_UNSUPPORTED_SHAPE, "attention sizes");
  T2_REQUIRE(lo.M % 8 == 0 && lo.M + 1 <= 128 && lo.Ti <= 1024, T2_ERR_UNSUPPORTED_SHAPE, "num_mels / T_in");
  T2_REQUIRE(cfg->enc_conv_layers >= 1 && cfg->enc_conv_layers <= 8 && cfg->postnet_layers >= 1 && cfg->postnet_layers <= 8,
             T2_ERR_INVALID_ARG, "layer counts");
  // ---- parameters (order == oracle/tacotron.py:param_shapes) ----
  lo.n_params = 0; lo.params.clear(); lo.enc.clear(); lo.post.clear();
  lo.p_emb = addp(lo, "inputs_embedding", {lo.NS, lo.E});
  auto conv_params = [&](ConvL& L, const std::string& pre) {
    L.p_k = addp(lo, pre + "kernel", {L.k, L.cin, L.cout}); L.p_b = addp(lo, pre + "bias", {L.cout});
    L.p_gamma = addp(lo, pre + "gamma", {L.cout}); L.p_beta = addp(lo, pre + "beta", {L.cout});
    L.p_mm = addp(lo, pre + "moving_mean", {L.cout}, false); L.p_mv = addp(lo, pre + "moving_variance", {L.cout}, false);
  };
  int cin = lo.E;
  for (int i = 0; i < cfg->enc_conv_layers; ++i) {
    ConvL L; L.cin = cin; L.cout = lo.C; L.k = cfg->enc_conv_kernel; L.act = 1; L.stream = 10 + i;
    char b[64]; snprintf(b, sizeof(b), "encoder_convolutions/conv_layer_%d/", i + 1);
    conv_params(L, b); lo.enc.push_back(L); cin = lo.C;
  }
  const char* dn[2] = {"fw", "bw"};
  for (int d = 0; d < 2; ++d) {
    lo.p_elk[d] = addp(lo, std::string("encoder_LSTM/") + dn[d] + "/kernel", {lo.C + lo.H, 4 * lo.H});
    lo.p_elb[d] = addp(lo, std::string("encoder_LSTM/") + dn[d] + "/bias", {4 * lo.H});
  }
  lo.p_mem = addp(lo, "attention/memory_layer/kernel", {2 * lo.H, lo.A});
  lo.p_qry = addp(lo, "attention/query_layer/kernel", {lo.D, lo.A});
  lo.p_lck = addp(lo, "attention/location_features_convolution/kernel", {lo.KA, 1, lo.F});
  lo.p_lcb = addp(lo, "attention/location_features_convolution/bias", {lo.F});
  lo.p_lfl = addp(lo, "attention/location_features_layer/kernel", {lo.F, lo.A});
  lo.p_v = addp(lo, "attention/attention_variable_projection", {lo.A});
  lo.p_ba = addp(lo, "attention/attention_bias", {lo.A});
  lo.p_p1k = addp(lo, "decoder_prenet/dense_1/kernel", {lo.M, lo.P1}); lo.p_p1b = addp(lo, "decoder_prenet/dense_1/bias", {lo.P1});
  lo.p_p2k = addp(lo, "decoder_prenet/dense_2/kernel", {lo.P1, lo.P2}); lo.p_p2b = addp(lo, "decoder_prenet/dense_2/bias", {lo.P2});
  const int K1 = lo.P2 + 2 * lo.H + lo.D, K2 = 2 * lo.D;
  lo.p_l1k = addp(lo, "decoder_LSTM/cell_1/kernel", {K1, 4 * lo.D}); lo.p_l1b = addp(lo, "decoder_LSTM/cell_1/bias", {4 * lo.D});
  lo.p_l2k = addp(lo, "decoder_LSTM/cell_2/kernel", {K2, 4 * lo.D}); lo.p_l2b = addp(lo, "decoder_LSTM/cell_2/bias", {4 * lo.D});
  const int PIK = lo.D + 2 * lo.H;
  lo.p_fk = addp(lo, "linear_transform_projection/kernel", {PIK, lo.M}); lo.p_fb = addp(lo, "linear_transform_projection/bias", {lo.M});
  lo.p_sk = addp(lo, "stop_token_projection/kernel", {PIK, 1}); lo.p_sb = addp(lo, "stop_token_projection/bias", {1});
  cin = lo.M;
  for (int i = 0; i < cfg->postnet_layers; ++i) {
    ConvL L; L.cin = cin; L.cout = lo.PC; L.k = cfg->postnet_kernel; L.act = (i + 1 < cfg->postnet_layers) ? 2 : 0; L.stream = 30 + i;
    char b[64]; snprintf(b, sizeof(b), "postnet_convolutions/conv_layer_%d/", i + 1);
    conv_params(L, b); lo.post.push_back(L); cin = lo.PC;
  }
  lo.p_ppk = addp(lo, "postnet_projection/kernel", {lo.PC, lo.M}); lo.p_ppb = addp(lo, "postnet_projection/bias", {lo.M});

  // ---- packed operands + pack jobs ----
  std::vector<PJ> jobs;
  long long o = 0;
  auto takeb = [&](long long bytes) { long long r = o; o = al256(o + bytes); return r; };
  auto pj = [&](long long src, int K, int N, long long dst_bytes, int ld, int tr, int col0, int perm = 0) {
    PJ j; j.src_off = src; j.K = K; j.N = N; j.dst_off = dst_bytes / 2; j.dst_ld = ld; j.transpose = tr; j.col0 = col0; j.scale = 1.f;
    j.perm_h = perm; j.part = 0; jobs.push_back(j);
  };
  const bool split = cfg->split_bf16 != 0;
  // split-bf16 operand: the K slot [col, col + slot) of the plain layout becomes [W_hi | W_hi | W_lo]
  auto pj3 = [&](long long src, int K, int N, long long dst_bytes, int ld, int col, int slot) {
    pj(src, K, N, dst_bytes, ld, 1, col);
    pj(src, K, N, dst_bytes, ld, 1, col + slot);
    pj(src, K, N, dst_bytes, ld, 1, col + 2 * slot);
    jobs.back().part = 2;
  };
  auto conv_pack = [&](ConvL& L) {
    L.cinp = (L.cin + 63) / 64 * 64;
    L.k_w = takeb(2LL * L.cout * L.k * L.cinp * (split ? 3 : 1));
    L.k_wT = takeb(2LL * L.cinp * L.k * L.cout);
    for (int j = 0; j < L.k; ++j) {
      if (split) pj3(L.p_k + (long long)j * L.cin * L.cout, L.cin, L.cout, L.k_w, 3 * L.k * L.cinp, 3 * j * L.cinp, L.cinp);
      else
      pj(L.p_k + (long long)j * L.cin * L.cout, L.cin, L.cout, L.k_w, L.k * L.cinp, 1, j * L.cinp);      // fwd: [cout][tap j | cin]
      pj(L.p_k + (long long)j * L.cin * L.cout, L.cin, L.cout, L.k_wT, L.k * L.cout, 0, j * L.cout);     // dgrad: [cin][tap j | cout]
    }
  };
  for (auto& L : lo.enc) conv_pack(L);
  for (auto& L : lo.post) conv_pack(L);
  for (int d = 0; d < 2; ++d) {
    lo.k_encWx[d] = takeb(2LL * 4 * lo.H * lo.C * (split ? 3 : 1));            // [4H][C]  input projection (natural gate order)
    if (split) pj3(lo.p_elk[d], lo.C, 4 * lo.H, lo.k_encWx[d], 3 * lo.C, 0, lo.C);
    else pj(lo.p_elk[d], lo.C, 4 * lo.H, lo.k_encWx[d], lo.C, 1, 0);
    lo.k_encWr[d] = takeb(2LL * 4 * lo.H * lo.H);            // [4H perm][H] recurrent, rows permuted for EPI_LSTM
    pj(lo.p_elk[d] + (long long)lo.C * 4 * lo.H, lo.H, 4 * lo.H, lo.k_encWr[d], lo.H, 1, 0, lo.H);
    lo.k_encWrT[d] = takeb(2LL * lo.H * 4 * lo.H);           // [H][4H] for the backward step
    pj(lo.p_elk[d] + (long long)lo.C * 4 * lo.H, lo.H, 4 * lo.H, lo.k_encWrT[d], 4 * lo.H, 0, 0);
  }
  lo.k_encWxT = takeb(2LL * lo.C * 8 * lo.H);                // [C][fw 4H | bw 4H]
  for (int d = 0; d < 2; ++d) pj(lo.p_elk[d], lo.C, 4 * lo.H, lo.k_encWxT, 8 * lo.H, 0, d * 4 * lo.H);
  lo.k_mem = takeb(2LL * lo.A * 2 * lo.H); pj(lo.p_mem, 2 * lo.H, lo.A, lo.k_mem, 2 * lo.H, 1, 0);
  lo.k_memT = takeb(2LL * 2 * lo.H * lo.A); pj(lo.p_mem, 2 * lo.H, lo.A, lo.k_memT, lo.A, 0, 0);
  const int Mp = 128;  // mel channels padded for TMA boxes
  lo.k_p1 = takeb(2LL * lo.P1 * Mp); pj(lo.p_p1k, lo.M, lo.P1, lo.k_p1, Mp, 1, 0);
  lo.k_p1T = takeb(2LL * lo.M * lo.P1); pj(lo.p_p1k, lo.M, lo.P1, lo.k_p1T, lo.P1, 0, 0);
  lo.k_p2 = takeb(2LL * lo.P2 * lo.P1); pj(lo.p_p2k, lo.P1, lo.P2, lo.k_p2, lo.P1, 1, 0);
  lo.k_p2T = takeb(2LL * lo.P1 * lo.P2); pj(lo.p_p2k, lo.P1, lo.P2, lo.k_p2T, lo.P2, 0, 0);
  lo.k_l1x = takeb(2LL * 4 * lo.D * lo.P2); pj(lo.p_l1k, lo.P2, 4 * lo.D, lo.k_l1x, lo.P2, 1, 0);
  lo.k_l1xT = takeb(2LL * lo.P2 * 4 * lo.D); pj(lo.p_l1k, lo.P2, 4 * lo.D, lo.k_l1xT, 4 * lo.D, 0, 0);
  const int K1r = 2 * lo.H + lo.D;
  lo.k_l1r = takeb(2LL * 4 * lo.D * K1r); pj(lo.p_l1k + (long long)lo.P2 * 4 * lo.D, K1r, 4 * lo.D, lo.k_l1r, K1r, 1, 0, lo.D);
  lo.k_l1rT = takeb(2LL * K1r * 4 * lo.D); pj(lo.p_l1k + (long long)lo.P2 * 4 * lo.D, K1r, 4 * lo.D, lo.k_l1rT, 4 * lo.D, 0, 0);
  lo.k_l2 = takeb(2LL * 4 * lo.D * K2); pj(lo.p_l2k, K2, 4 * lo.D, lo.k_l2, K2, 1, 0, lo.D);
  lo.k_l2T = takeb(2LL * K2 * 4 * lo.D); pj(lo.p_l2k, K2, 4 * lo.D, lo.k_l2T, 4 * lo.D, 0, 0);
  lo.k_proj = takeb(2LL * 128 * PIK);                         // rows 0..M-1 frame projection, row M stop projection
  pj(lo.p_fk, PIK, lo.M, lo.k_proj, PIK, 1, 0);
  { PJ j; j.src_off = lo.p_sk; j.K = PIK; j.N = 1; j.dst_off = lo.k_proj / 2 + (long long)lo.M * PIK; j.dst_ld = PIK; j.transpose = 1; j.col0 = 0;
    j.scale = 1.f; j.perm_h = 0; j.part = 0; jobs.push_back(j); }
  lo.k_projT = takeb(2LL * PIK * 128);                        // [PIK][128]: cols 0..M-1 Wf, col M Ws
  pj(lo.p_fk, PIK, lo.M, lo.k_projT, 128, 0, 0);
  pj(lo.p_sk, PIK, 1, lo.k_projT, 128, 0, lo.M);
  lo.k_pp = takeb(2LL * 128 * lo.PC * (split ? 3 : 1));
  if (split) pj3(lo.p_ppk, lo.PC, lo.M, lo.k_pp, 3 * lo.PC, 0, lo.PC);
  else pj(lo.p_ppk, lo.PC, lo.M, lo.k_pp, lo.PC, 1, 0);
  lo.k_ppT = takeb(2LL * lo.PC * 128); pj(lo.p_ppk, lo.PC, lo.M, lo.k_ppT, 128, 0, 0);
  lo.k_qT = takeb(2LL * lo.A * lo.D); pj(lo.p_qry, lo.D, lo.A, lo.k_qT, lo.D, 1, 0);   // [A][D] for the attention kernel
  lo.packed_bytes = o;
  lo.n_packjobs = int(jobs.size());

  // ---- workspace ----
  o = 0;
  const long long B = lo.B, Ti = lo.Ti, To = lo.To;
  const long long xm = split ? 2 : 1;       // split-bf16: stored conv-stack activations are [hi | lo]; pre-batch-norm activations fp32
  lo.w_emb = takeb(B * Ti * lo.E * 2 * xm);
  auto conv_ws = [&](ConvL& L, long long T) {
    L.w_y = takeb(B * T * L.cout * (split ? 4 : 2)); L.w_x = takeb(B * T * L.cout * 2 * xm); L.w_stats = takeb(8LL * L.cout * 4);
  };
  for (auto& L : lo.enc) conv_ws(L, Ti);
  for (int d = 0; d < 2; ++d) {
    lo.w_encpre[d] = takeb(B * Ti * 4 * lo.H * 4);
    lo.w_ench[d] = takeb((Ti + 1) * B * lo.H * 2);      // h_state history, slot s = state after s processed steps
    lo.w_encc[d] = takeb((Ti + 1) * B * lo.H * 4);
    lo.w_encg[d] = takeb(Ti * B * 4 * lo.H * 2);
    lo.w_enct[d] = takeb(Ti * B * lo.H * 2);
  }
  lo.w_memory = takeb(B * Ti * 2 * lo.H * 2);
  lo.w_values = takeb(B * Ti * 2 * lo.H * 2);
  lo.w_keys = takeb(B * Ti * lo.A * 4);
  lo.w_decin = takeb(B * To * lo.M * 2);                  // time-major [To][B][M]
  lo.w_pn1 = takeb(To * B * lo.P1 * 2);
  lo.w_pn2 = takeb(To * B * lo.P2 * 2);
  lo.w_pre1 = takeb(To * B * 4 * lo.D * 4);
  lo.w_S1 = takeb((To + 1) * B * K1r * 2);
  lo.w_S2 = takeb((To + 1) * B * K2 * 2);
  lo.w_PI = takeb(To * B * PIK * 2);
  lo.w_c1 = takeb((To + 1) * B * lo.D * 4); lo.w_c2 = takeb((To + 1) * B * lo.D * 4);
  lo.w_g1 = takeb(To * B * 4 * lo.D * 2); lo.w_g2 = takeb(To * B * 4 * lo.D * 2);
  lo.w_t1 = takeb(To * B * lo.D * 2); lo.w_t2 = takeb(To * B * lo.D * 2);
  lo.w_cum = takeb(B * Ti * 4);
  lo.w_alpha = takeb(To * B * Ti * 4);
  lo.w_projo = takeb(To * B * 128 * 4);
  lo.w_decbm = takeb(split ? B * To * 256 * 2 : B * To * lo.M * 2);     /* split: [hi(M) padded to 128 | lo(M) padded to 128] */ lo.w_decf = takeb(B * To * lo.M * 4); lo.w_stop = takeb(B * To * 4);
  for (auto& L : lo.post) conv_ws(L, To);
  lo.w_tlen = takeb(B * 4);
  lo.w_resid = takeb(B * To * 128 * 4); lo.w_mel = takeb(B * To * lo.M * 4);
  lo.w_scal = takeb(64 * 4);
  // backward
  lo.w_dmel = takeb(B * To * 128 * 2);                    // bf16 [B][To][128] (cols >= M zero)
  lo.w_dY = takeb(2 * B * (To > Ti ? To : Ti) * (lo.PC > lo.C ? lo.PC : lo.C) * 2);   // ping-pong activation grads bf16
  lo.w_ddec_tm = takeb(To * B * 128 * 2);
  lo.w_dPI = takeb(To * B * PIK * 4);
  lo.w_dh1ext = takeb(B * lo.D * 4); lo.w_dh2ext = takeb(B * lo.D * 4);
  lo.w_dhs1 = takeb(B * lo.D * 4); lo.w_dhs2 = takeb(B * lo.D * 4); lo.w_dcs1 = takeb(B * lo.D * 4); lo.w_dcs2 = takeb(B * lo.D * 4);
  lo.w_dg1 = takeb(To * B * 4 * lo.D * 2); lo.w_dg2 = takeb(To * B * 4 * lo.D * 2);
  lo.w_dgstep = 0;
  lo.w_dctxl = takeb(B * 2 * lo.H * 4);
  lo.w_dctx_all = takeb(To * B * 2 * lo.H * 2);
  lo.w_dq_all = takeb(To * B * lo.A * 2);
  lo.w_dcum = takeb(B * Ti * 4); lo.w_cumrun = takeb(B * Ti * 4);
  lo.w_dkeys = takeb(B * Ti * lo.A * 4);
  lo.w_dvalues = takeb(B * Ti * 2 * lo.H * 4);
  lo.w_attacc = takeb(B * (lo.KA + 2) * lo.A * 4);
  lo.w_attU = takeb((2 * lo.KA + 4) * lo.A * 4);
  lo.w_dpn2 = takeb(To * B * lo.P2 * 2); lo.w_dpn1 = takeb(To * B * lo.P1 * 2);
  for (int d = 0; d < 2; ++d) {
    lo.w_dencpre[d] = takeb(B * Ti * 4 * lo.H * 2);       // bf16 gate grads [B][Ti][4H] (batch-major, = dpre)
    lo.w_encdh[d] = takeb(B * lo.H * 4); lo.w_encdc[d] = takeb(B * lo.H * 4);
  }
  lo.w_encdg = takeb(B * 4 * lo.H * 2);
  lo.w_ddecf = takeb(B * To * lo.M * 4);
  for (int d = 0; d < 2; ++d) lo.w_encdgall[d] = takeb(Ti * B * 4 * lo.H * 2);
  lo.w_dkeysb = takeb(B * Ti * lo.A * 2);
  lo.w_dz = takeb(To * B * (lo.P1 > lo.P2 ? lo.P1 : lo.P2) * 2);
  lo.w_dx3 = takeb(B * Ti * lo.C * 2);
  lo.w_demb = takeb(B * Ti * lo.E * 2);
  lo.w_tiles = takeb(16384 * sizeof(WgradTile));
  lo.w_packjobs = takeb((long long)jobs.size() * sizeof(PJ));
  lo.n_reg = 0;
  for (auto& p : lo.params) lo.n_reg += p.reg ? 1 : 0;
  lo.w_regtab = takeb((long long)lo.n_reg * 2 * sizeof(long long));
  lo.w_zero = takeb(B * 4096 * 4);
  lo.workspace_bytes = o;
  if (jobs_out) jobs_out->swap(jobs);
  return T2_OK;
}

// ------------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------------
__global__ void tpack_kernel(const float* __restrict__ params, bf16* __restrict__ packed, const PJ* __restrict__ jobs) {
  // 64x64 tiles through shared memory: float2 reads along the source's fast axis (N), bf16x2 writes along the destination's
  // fast axis (K for the transposing jobs); scalar fallbacks when an offset / leading dimension is odd
  __shared__ float tile[64][65];
  const PJ j = jobs[blockIdx.y];
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
          if (j.perm_h > 0) {  // gate-major column n = g*H + u  ->  EPI_LSTM row (u/32)*128 + g*32 + u%32
            const int g = n / j.perm_h, u = n % j.perm_h;
            row = (u / 32) * 128 + g * 32 + (u % 32);
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

__global__ void embed_fwd_kernel(const int* __restrict__ idx, const float* __restrict__ table, bf16* __restrict__ out, long long npos, int E,
                                 int split) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= npos * E) return;
  const float v = table[(long long)idx[e / E] * E + e % E];
  if (!split) { out[e] = __float2bfloat16(v); return; }
  const bf16 hi = __float2bfloat16(v);
  bf16* row = out + (e / E) * 2 * E + e % E;       // rows [hi(E) | lo(E)]
  row[0] = hi; row[E] = __float2bfloat16(v - __bfloat162float(hi));
}
__global__ void embed_bwd_kernel(const int* __restrict__ idx, const bf16* __restrict__ dx, float* __restrict__ dtable, long long npos, int E) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= npos * E) return;
  atomicAdd(dtable + (long long)idx[e / E] * E + e % E, __bfloat162float(dx[e]));
}

// per-channel sum / sum of squares of y [rows][C] (bf16) -> stats[0..C), stats[C..2C)
__device__ __forceinline__ float ld_act(const bf16* p, long long i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ float ld_act(const float* p, long long i) { return p[i]; }
template <typename TY>
__global__ void bn_stats_kernel(const TY* __restrict__ y, float* __restrict__ stats, long long rows, int C) {
  const long long per = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * per, r1 = r0 + per < rows ? r0 + per : rows;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f, q = 0.f;
    for (long long r = r0; r < r1; ++r) { const float v = ld_act(y, r * C + c); s += v; q += v * v; }
    atomicAdd(stats + c, s); atomicAdd(stats + C + c, q);
  }
}
// x = dropout(((y - mean) * rstd) * gamma + beta); writes mean / rstd into stats[2C..4C), updates the moving stats
template <typename TY>
__global__ void bn_apply_kernel(const TY* __restrict__ y, bf16* __restrict__ x, float* __restrict__ stats, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float* __restrict__ mm, float* __restrict__ mv, long long rows, int C,
                                int training, float p, unsigned long long seed, const unsigned long long* step, int stream, int split) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= rows * C) return;
  const int c = int(e % C);
  float mean, rstd;
  if (training) {
    mean = stats[c] / float(rows);
    const float var = fmaxf(stats[C + c] / float(rows) - mean * mean, 0.f);
    rstd = rsqrtf(var + 1e-3f);
    if (e < C) {
      stats[2 * C + c] = mean; stats[3 * C + c] = rstd;
      mm[c] = 0.99f * mm[c] + 0.01f * mean; mv[c] = 0.99f * mv[c] + 0.01f * var;
    }
  } else { mean = mm[c]; rstd = rsqrtf(mv[c] + 1e-3f); }
  float v = (ld_act(y, e) - mean) * rstd * gamma[c] + beta[c];
  if (training && p > 0.f) {
    if (step) seed += *step;
    v = hash_uniform32(hash_seed(seed, uint32_t(stream)), (unsigned long long)e) >= p ? v / (1.f - p) : 0.f;
  }
  if (!split) { x[e] = __float2bfloat16(v); return; }
  const bf16 hi = __float2bfloat16(v);
  bf16* row = x + (e / C) * 2 * C + c;             // rows [hi(C) | lo(C)]
  row[0] = hi; row[C] = __float2bfloat16(v - __bfloat162float(hi));
}
// backward sums: sg[c] = sum g, sgx[c] = sum g * xhat   (g = dout * dropout mask)
__global__ void bn_bwd_stats_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ y, const float* __restrict__ stats,
                                    float* __restrict__ bsum, long long rows, int C, float p, unsigned long long seed,
                                    const unsigned long long* step, int stream) {
  const long long per = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * per, r1 = r0 + per < rows ? r0 + per : rows;
  if (step) seed += *step;
  const uint32_t hs = hash_seed(seed, uint32_t(stream));
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float mean = stats[2 * C + c], rstd = stats[3 * C + c];
    float s = 0.f, q = 0.f;
    for (long long r = r0; r < r1; ++r) {
      float g = __bfloat162float(dout[r * C + c]);
      if (p > 0.f) g = hash_uniform32(hs, (unsigned long long)(r * C + c)) >= p ? g / (1.f - p) : 0.f;
      s += g; q += g * (__bfloat162float(y[r * C + c]) - mean) * rstd;
    }
    atomicAdd(bsum + c, s); atomicAdd(bsum + C + c, q);
  }
}
__global__ void bn_bwd_apply_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ y, const float* __restrict__ stats,
                                    const float* __restrict__ bsum, const float* __restrict__ gamma, bf16* __restrict__ dpre,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows, int C, int act, float p,
                                    unsigned long long seed, const unsigned long long* step, int stream) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= rows * C) return;
  const int c = int(e % C);
  if (step) seed += *step;
  const float mean = stats[2 * C + c], rstd = stats[3 * C + c];
  float g = __bfloat162float(dout[e]);
  if (p > 0.f) g = hash_uniform32(hash_seed(seed, uint32_t(stream)), (unsigned long long)e) >= p ? g / (1.f - p) : 0.f;
  const float yv = __bfloat162float(y[e]);
  const float xhat = (yv - mean) * rstd;
  float dy = gamma[c] * rstd * (g - bsum[c] / float(rows) - xhat * bsum[C + c] / float(rows));
  if (act == 1) dy = yv > 0.f ? dy : 0.f;
  else if (act == 2) dy *= (1.f - yv * yv);
  dpre[e] = __float2bfloat16(dy);
  if (e < C) { dgamma[c] += bsum[C + c]; dbeta[c] += bsum[c]; }
}

// column sums of a bf16 [rows][ld] matrix (first C columns) into fp32 dst (+=), scaled
__global__ void colsum_bf16_kernel(const bf16* __restrict__ src, long long rows, int C, int ld, float* __restrict__ dst, float scale) {
  const long long per = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * per, r1 = r0 + per < rows ? r0 + per : rows;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (long long r = r0; r < r1; ++r) s += __bfloat162float(src[r * ld + c]);
    atomicAdd(dst + c, s * scale);
  }
}

// memory [B][Ti][2H] -> values = memory * mask (BahdanauAttention memory masking)
__global__ void mask_values_kernel(const bf16* __restrict__ mem, const int* __restrict__ lens, bf16* __restrict__ vals, int B, int Ti, int C2) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= (long long)B * Ti * C2) return;
  const int t = int((e / C2) % Ti), b = int(e / ((long long)C2 * Ti));
  vals[e] = t < lens[b] ? mem[e] : __float2bfloat16(0.f);
}
// decoder inputs, time-major: dec_in[t][b][:] = t == 0 ? 0 : target[b][t-1][:]   (helpers.py:62-128, r = 1)
__global__ void decin_kernel(const float* __restrict__ tgt, bf16* __restrict__ out, int B, int To, int M) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= (long long)To * B * M) return;
  const int m = int(e % M), b = int((e / M) % B), t = int(e / ((long long)M * B));
  out[e] = __float2bfloat16(t == 0 ? 0.f : tgt[((long long)b * To + t - 1) * M + m]);
}

// ---- location-sensitive attention, one CTA per batch item per decoder step (attention.py:169-226) -------------
// The location branch conv1d(k=31, 1 -> F) followed by dense(F -> A) is linear in the cumulative alignments, so it is
// evaluated as ONE 31-tap filter bank U[k][a] = sum_f K[k][f] Wl[f][a] with offset u0[a] = sum_f bK[f] Wl[f][a] + b_a[a]
// (built once per forward by att_prep_kernel); the backward pass differentiates through the same factorisation.
constexpr int kAttThreads = 512;
__device__ long long* g_att_dbg = nullptr;   // optional phase stamps (tools only): [0..15] forward, [16..31] backward
#define ATT_STAMP(i) do { if (g_att_dbg && blockIdx.x == 0 && threadIdx.x == 0) g_att_dbg[i] = clock64(); } while (0)
__global__ void att_prep_kernel(const float* __restrict__ K, const float* __restrict__ bK, const float* __restrict__ Wl,
                                const float* __restrict__ ba, float* __restrict__ U, int KA, int F, int A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (KA + 1) * A) return;
  const int k = i / A, a = i % A;
  float acc = 0.f;
  if (k < KA) { for (int f = 0; f < F; ++f) acc += K[k * F + f] * Wl[f * A + a]; }
  else { acc = ba[a]; for (int f = 0; f < F; ++f) acc += bK[f] * Wl[f * A + a]; }
  U[i] = acc;   // rows 0..KA-1: U, row KA: u0
}
struct AttArgs {
  const bf16* h2out; int ld_h2;            // query source: PI_all[t][b][0:D]
  const bf16* WqT;                          // [A][D] bf16
  const float* U;                           // [KA + 1][A]
  const float* v;
  const float* keys;                        // [B][Ti][A] fp32
  const bf16* values;                       // [B][Ti][C2]
  const int* lens;
  float* cum;                               // [B][Ti] running cumulative alignments (in/out)
  float* alpha;                             // [B][Ti] output for this step
  bf16* ctx_a; int ld_a;                    // context -> S1_all[t+1][b][0:C2]
  bf16* ctx_b; int ld_b;                    // context -> PI_all[t][b][D:]
  int B, Ti, D, A, KA, C2;
};
// q[a] = sum_k h[k] WqT[a][k]: one warp per output row (two rows in flight), lanes stride the row in 16-byte pieces so
// that every load instruction reads 512 contiguous bytes (the 4-threads-per-output form touched 32 sectors per load)
// A length-D fp32 vector that is dotted against 16-byte bf16 pieces lives in shared memory in a SPLIT layout: element
// 8p + x of the vector sits at xs[(x >> 2) * (D/2) + 4p + (x & 3)], so lane p reads two float4 at a 16-byte lane stride
// (conflict-free); the natural layout (32-byte lane stride) made every one of these loads an 8-way bank conflict.
__device__ __forceinline__ int split8(int i, int D) { return ((i >> 2) & 1) * (D >> 1) + ((i >> 3) << 2) + (i & 3); }
__device__ __forceinline__ float dot8s(const uint4 u, const float* __restrict__ xs, int p, int D) {
  const float4 lo = *reinterpret_cast<const float4*>(xs + 4 * p), hi = *reinterpret_cast<const float4*>(xs + (D >> 1) + 4 * p);
  return bf16lo(u.x) * lo.x + bf16hi(u.x) * lo.y + bf16lo(u.y) * lo.z + bf16hi(u.y) * lo.w + bf16lo(u.z) * hi.x +
         bf16hi(u.z) * hi.y + bf16lo(u.w) * hi.z + bf16hi(u.w) * hi.w;
}
// q[a] = sum_k h[k] WqT[a][k]: one warp per output row (two rows in flight), lanes stride the row in 16-byte pieces so
// that every load instruction reads 512 contiguous bytes; hs is in the split layout above
__device__ __forceinline__ void att_query(const bf16* __restrict__ WqT, const float* __restrict__ hs, float* __restrict__ q, int A, int D) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NW = kAttThreads / 32;
  const int n16 = D >> 3;   // 16-byte pieces per row
  for (int o = warp; o < A; o += 2 * NW) {
    const int o2 = o + NW;
    const uint4* w0 = reinterpret_cast<const uint4*>(WqT + (long long)o * D);
    const uint4* w1 = reinterpret_cast<const uint4*>(WqT + (long long)(o2 < A ? o2 : o) * D);
    float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
    for (int i = lane; i < n16; i += 32) {
      const uint4 u0 = __ldg(w0 + i), u1 = __ldg(w1 + i);
      a0 += dot8s(u0, hs, i, D);
      a1 += dot8s(u1, hs, i, D);
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1);
    if (lane == 0) { q[o] = a0; if (o2 < A) q[o2] = a1; }
  }
}
// ---- the location filter bank on tensor cores -------------------------------------------------------------------
// pl[j][n] = u0[n] + sum_k cum[j + k - half] U[k][n] is a [T_in x 32] Toeplitz matrix times the [32 x A] filter bank
// (row KA of the bank is the offset u0, matched by a column of ones): per batch item 160 x 32 x 128 - far too small for a
// tcgen05 tile pipeline, so it runs as warp-level mma.sync.m16n8k8 TF32 (fp32 accumulate) straight out of shared memory;
// the Toeplitz operand is never materialised (fragments read cum[j + k]). The same instruction computes the three
// products of the backward pass (dU = T^T dE, P = dE U^T for dcum). The scalar FMA form was issue/shared-memory bound:
// 20 k (forward) / 55 k (backward) cycles per step at T_in = 160 (tools/att_phases.py).
__device__ __forceinline__ uint32_t f2tf32(float f) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(f));
  return r;
}
// D(16x8) += A(16x8, row) * B(8x8, col); g = lane >> 2, t = lane & 3:
//   a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4) | b0 (k=t, n=g) b1 (k=t+4, n=g) | c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__host__ __device__ inline int att_ti16(int Ti) { return (Ti + 15) & ~15; }
__host__ __device__ inline int att_cumlen(int Ti, int KA) { return (att_ti16(Ti) + KA + 16 + 3) & ~3; }   // zero-padded cum window
__host__ __device__ inline int att_erows(int Ti, int KA) { return (att_ti16(Ti) + 2 * (KA / 2) + 15) & ~15; } // rows of dE (backward)
constexpr int kAttPad = 8;   // row padding (floats) of the filter bank / dE tiles in shared memory: conflict-free fragments
// Toeplitz element T[j][k]: cum window for the taps, a column of ones for the offset row, zero beyond
__device__ __forceinline__ float toep(const float* __restrict__ cum, int j, int k, int KA) {
  return k < KA ? cum[j + k] : (k == KA ? 1.f : 0.f);
}
// pl for the 16 rows j0.. and the 64 channels n0..: acc[nt] = C fragment of n-tile nt (8 channels each)
__device__ __forceinline__ void loc_tile(const float* __restrict__ Us, int AP, const float* __restrict__ cum, int KA, int j0, int n0,
                                         float (&acc)[8][4]) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int k = ks * 8 + t;
    uint32_t af[4];
    af[0] = f2tf32(toep(cum, j0 + g, k, KA)); af[1] = f2tf32(toep(cum, j0 + g + 8, k, KA));
    af[2] = f2tf32(toep(cum, j0 + g, k + 4, KA)); af[3] = f2tf32(toep(cum, j0 + g + 8, k + 4, KA));
    const bool v0 = k <= KA, v1 = k + 4 <= KA;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int n = n0 + nt * 8 + g;
      const uint32_t b0 = v0 ? f2tf32(Us[k * AP + n]) : 0u, b1 = v1 ? f2tf32(Us[(k + 4) * AP + n]) : 0u;
      mma_tf32(acc[nt], af, b0, b1);
    }
  }
}
inline size_t att_fwd_smem(int Ti, int KA, int A, int D, int C2) {
  return sizeof(float) * (size_t)((KA + 1) * (A + kAttPad) + att_cumlen(Ti, KA) + A + ((Ti + 3) & ~3) + D + 8 * C2 + 32) + 64;
}
__global__ void __launch_bounds__(kAttThreads) att_fwd_kernel(AttArgs a) {
  extern __shared__ __align__(16) float sm[];
  pdl_wait();
  pdl_launch_dependents();
  ATT_STAMP(0);
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Ti = a.Ti, A = a.A, AP = A + kAttPad, half = a.KA / 2, NW = kAttThreads / 32;
  const int Tip = (Ti + 3) & ~3, cumlen = att_cumlen(Ti, a.KA);
  float* Us = sm;                       // [(KA+1)][AP] filter bank, row KA = offset u0
  float* cum = Us + (a.KA + 1) * AP;    // [cumlen] zero-padded halo
  float* q = cum + cumlen;              // [A]
  float* e = q + A;                     // [Ti]
  float* hs = e + Tip;                  // [D] query source as fp32
  float* part = hs + a.D;               // [8][C2] context partials
  float* red = part + 8 * a.C2;         // [32]
  for (int i = tid; i < (a.KA + 1) * A; i += kAttThreads) Us[(i / A) * AP + (i % A)] = a.U[i];
  for (int i = tid; i < cumlen; i += kAttThreads) {
    const int j = i - half;
    cum[i] = (j >= 0 && j < Ti) ? a.cum[(long long)b * Ti + j] : 0.f;
  }
  for (int i = tid; i < a.D; i += kAttThreads) hs[split8(i, a.D)] = __bfloat162float(a.h2out[(long long)b * a.ld_h2 + i]);
  for (int i = tid; i < Tip; i += kAttThreads) e[i] = 0.f;
  __syncthreads();
  ATT_STAMP(1);
  att_query(a.WqT, hs, q, A, a.D);
  __syncthreads();
  ATT_STAMP(2);
  const int len = a.lens[b];
  {
    // energies: units of 16 memory rows x 64 channels; e[j] += sum over the unit's channels of v tanh(keys + q + pl)
    const int g = lane >> 2, t = lane & 3;
    const int n_nh = A >> 6, n_units = ((len + 15) >> 4) * n_nh;
    for (int u = warp; u < n_units; u += NW) {
      const int j0 = (u / n_nh) * 16, n0 = (u % n_nh) * 64;
      float acc[8][4];
      loc_tile(Us, AP, cum, a.KA, j0, n0, acc);
      const int r0 = j0 + g, r1 = r0 + 8;
      const float* k0p = a.keys + ((long long)b * Ti + (r0 < len ? r0 : 0)) * A + n0 + 2 * t;
      const float* k1p = a.keys + ((long long)b * Ti + (r1 < len ? r1 : 0)) * A + n0 + 2 * t;
      float2 ky0[8], ky1[8];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) { ky0[nt] = __ldg(reinterpret_cast<const float2*>(k0p + nt * 8)); ky1[nt] = __ldg(reinterpret_cast<const float2*>(k1p + nt * 8)); }
      float e0 = 0.f, e1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int c = n0 + nt * 8 + 2 * t;
        const float2 qq = *reinterpret_cast<const float2*>(q + c);
        const float2 vv = __ldg(reinterpret_cast<const float2*>(a.v + c));
        e0 += vv.x * tanhf_(ky0[nt].x + qq.x + acc[nt][0]) + vv.y * tanhf_(ky0[nt].y + qq.y + acc[nt][1]);
        e1 += vv.x * tanhf_(ky1[nt].x + qq.x + acc[nt][2]) + vv.y * tanhf_(ky1[nt].y + qq.y + acc[nt][3]);
      }
      e0 += __shfl_xor_sync(0xffffffffu, e0, 1); e0 += __shfl_xor_sync(0xffffffffu, e0, 2);
      e1 += __shfl_xor_sync(0xffffffffu, e1, 1); e1 += __shfl_xor_sync(0xffffffffu, e1, 2);
      if (t == 0) {
        if (r0 < len) atomicAdd(&e[r0], e0);
        if (r1 < len) atomicAdd(&e[r1], e1);
      }
    }
  }
  __syncthreads();
  ATT_STAMP(3);
  float mx = -INFINITY;
  for (int j = tid; j < len; j += kAttThreads) mx = fmaxf(mx, e[j]);
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < kAttThreads / 32; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float s = 0.f;
  for (int j = tid; j < Ti; j += kAttThreads) { const float p = j < len ? __expf(e[j] - mx) : 0.f; e[j] = p; s += p; }
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  s = 0.f;
  for (int w = 0; w < kAttThreads / 32; ++w) s += red[w];
  const float inv = 1.f / s;
  for (int j = tid; j < Ti; j += kAttThreads) {
    const float al = e[j] * inv;
    e[j] = al;
    a.alpha[(long long)b * Ti + j] = al;
    a.cum[(long long)b * Ti + j] = cum[j + half] + al;
  }
  __syncthreads();
  ATT_STAMP(4);
  // context = alpha . values: 8 row groups x (C2/8) column chunks of 8 channels
  {
    const int nch = a.C2 >> 3;                 // uint4 chunks per row
    const int rg = tid / nch, ch = tid % nch;  // kAttThreads >= 8 * nch for C2 <= 512
    if (rg < 8) {
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      const uint4* vp = reinterpret_cast<const uint4*>(a.values + (long long)b * Ti * a.C2) + ch;
#pragma unroll 4
      for (int j = rg; j < len; j += 8) {
        const uint4 u = __ldg(vp + (long long)j * nch);
        const float al = e[j];
        acc[0] += al * bf16lo(u.x); acc[1] += al * bf16hi(u.x); acc[2] += al * bf16lo(u.y); acc[3] += al * bf16hi(u.y);
        acc[4] += al * bf16lo(u.z); acc[5] += al * bf16hi(u.z); acc[6] += al * bf16lo(u.w); acc[7] += al * bf16hi(u.w);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) part[rg * a.C2 + ch * 8 + i] = acc[i];
    }
  }
  __syncthreads();
  ATT_STAMP(5);
  for (int c = tid; c < a.C2; c += kAttThreads) {
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) acc += part[r * a.C2 + c];
    const bf16 r16 = __float2bfloat16(acc);
    if (a.ctx_a) a.ctx_a[(long long)b * a.ld_a + c] = r16;
    a.ctx_b[(long long)b * a.ld_b + c] = r16;
  }
  ATT_STAMP(6);
}

// ---- output heads / losses ---------------------------------------------------------------------------------------
// projo [To][B][128] fp32 (cols 0..M-1 frames, col M stop logit) -> clipped decoder output (batch-major), stop logits,
// loss sums: scal[0] += sum (dec - tgt)^2, scal[2] += sum BCE(stop)
__global__ void dec_finish_kernel(const float* __restrict__ projo, const float* __restrict__ tgt, const float* __restrict__ stop_tgt,
                                  bf16* __restrict__ dec_bm, float* __restrict__ dec_f, float* __restrict__ stop, float* __restrict__ scal,
                                  int B, int To, int M, int clip, float lo, float hi, int split, const int* __restrict__ tlen, float pos_w) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  float l0 = 0.f, l2 = 0.f, nz = 0.f;
  if (e < (long long)B * To * (M + 1)) {
    const int m = int(e % (M + 1)), t = int((e / (M + 1)) % To), b = int(e / ((long long)(M + 1) * To));
    const bool live = !tlen || t < tlen[b];     // mask_decoder: frames past the target length do not count
    const float v = projo[((long long)t * B + b) * 128 + m];
    if (m < M) {
      const float d = clip ? fminf(fmaxf(v, lo), hi) : v;
      const long long o = ((long long)b * To + t) * M + m;
      dec_f[o] = d;
      if (!split) dec_bm[o] = __float2bfloat16(d);
      else {   // rows [hi(M) zero-padded to 128 | lo(M) zero-padded to 128] (the buffer is cleared once at init)
        const bf16 h = __float2bfloat16(d);
        bf16* row = dec_bm + ((long long)b * To + t) * 256 + m;
        row[0] = h; row[128] = __float2bfloat16(d - __bfloat162float(h));
      }
      if (tgt && live) { const float df = d - tgt[o]; l0 = df * df; }
    } else {
      stop[(long long)b * To + t] = v;
      if (stop_tgt) {
        const float z = stop_tgt[(long long)b * To + t];
        if (!tlen) l2 = fmaxf(v, 0.f) - v * z + log1pf(__expf(-fabsf(v)));
        else if (live) {   // tf.nn.weighted_cross_entropy_with_logits, then / count_nonzero(masked loss) (modules.py:450-455)
          l2 = (1.f - z) * v + (1.f + (pos_w - 1.f) * z) * (log1pf(__expf(-fabsf(v))) + fmaxf(-v, 0.f));
          nz = l2 != 0.f ? 1.f : 0.f;
        }
      }
    }
  }
  l0 = warp_sum(l0); l2 = warp_sum(l2); nz = warp_sum(nz);
  if ((threadIdx.x & 31) == 0) {
    if (l0 != 0.f) atomicAdd(scal + 0, l0);
    if (l2 != 0.f) atomicAdd(scal + 2, l2);
    if (nz != 0.f) atomicAdd(scal + 4, nz);
  }
}
// mel = clip(dec + residual); scal[1] += sum (mel - tgt)^2
__global__ void mel_finish_kernel(const float* __restrict__ dec_f, const float* __restrict__ resid, const float* __restrict__ tgt,
                                  float* __restrict__ mel, float* __restrict__ scal, long long npos, int M, int clip, float lo, float hi,
                                  const int* __restrict__ tlen, int To) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  float l = 0.f;
  if (e < npos * M) {
    const long long pos = e / M; const int m = int(e % M);
    float v = dec_f[e] + resid[pos * 128 + m];
    if (clip) v = fminf(fmaxf(v, lo), hi);
    mel[e] = v;
    if (tgt && (!tlen || int(pos % To) < tlen[pos / To])) { const float d = v - tgt[e]; l = d * d; }
  }
  l = warp_sum(l);
  if ((threadIdx.x & 31) == 0 && l != 0.f) atomicAdd(scal + 1, l);
}
__global__ void reg_loss_kernel(const float* __restrict__ params, const long long* __restrict__ tab, int n, float* __restrict__ scal) {
  const long long off = tab[2 * blockIdx.y], len = tab[2 * blockIdx.y + 1];
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < len; i += (long long)gridDim.x * blockDim.x) {
    const float v = params[off + i]; s += v * v;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0 && s != 0.f) atomicAdd(scal + 3, 0.5f * s);
}

__global__ void proj_bias_kernel(float* p, const float* fb, const float* sb, long long rows, int M) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= rows * (M + 1)) return;
  const int m = int(e % (M + 1));
  p[(e / (M + 1)) * 128 + m] += m < M ? fb[m] : sb[0];
}
// normalisers -> s[5] (mel terms), s[6] (stop term); out (nullable) = the four normalised loss terms
__global__ void loss_norm_kernel(float* s, float* out, float n_mel, float n_stop, float regw, const int* tlen, int B, int To, int M) {
  if (tlen) {
    float frames = 0.f;
    for (int b = 0; b < B; ++b) frames += float(tlen[b] < To ? tlen[b] : To);
    n_mel = fmaxf(frames * float(M), 1.f);      // count_nonzero of the broadcast mask (tf.losses.mean_squared_error weights)
    n_stop = fmaxf(s[4], 1.f);                  // count_nonzero of the masked stop-token losses
  }
  s[5] = n_mel; s[6] = n_stop;
  if (out) { out[0] = s[0] / n_mel; out[1] = s[1] / n_mel; out[2] = s[2] / n_stop; out[3] = s[3] * regw; }
}

// generic helper: 1x1 / k-tap conv GEMM through the engine
// split != 0 ("fp32-class" conv stacks): the input rows are [hi | lo], each half `C` channels zero-padded to Cp = ceil(C / 64) * 64; per
// tap one segment over both halves against [W_hi | W_hi] and one over the hi half against [W_lo] (wK = 3 * ntaps * Cp); a bf16 output is
// written as [hi(ldo) | lo(ldo)]
int conv_gemm(const void* a, int C, long long T, int Bn, const void* w, int N, int wK, int ntaps, const int* shifts, int BN, float* bias,
              int act, void* out_bf16, float* out_f32, int ldo, int nvalid, float pdrop, int stream_id, unsigned long long seed,
              const unsigned long long* d_step, cudaStream_t st, int split = 0) {
  ActGemmCall g;
  memset(&g, 0, sizeof(g));
  const int nkb = (C + kBK - 1) / kBK;
  if (split) {
    const int Cp = nkb * kBK;
    g.a[0] = make_act(a, 2 * Cp, int(T), Bn, 1, 2 * Cp); g.na = 1;
    g.nseg = 0;
    for (int s = 0; s < ntaps; ++s) {
      g.seg[g.nseg++] = Seg{0, shifts ? shifts[s] : 0, 0, 2 * nkb, 0, 1};
      g.seg[g.nseg++] = Seg{0, shifts ? shifts[s] : 0, 0, nkb, 0, 1};
    }
    T2_REQUIRE(g.nseg <= kMaxSeg && pdrop == 0.f, T2_ERR_UNSUPPORTED_SHAPE, "split conv_gemm: too many taps / dropout in the epilogue");
    g.epi.i[11] = 1;
  } else {
  g.a[0] = make_act(a, C, int(T), Bn, 1, C); g.na = 1;
  for (int s = 0; s < ntaps; ++s) g.seg[s] = Seg{0, shifts ? shifts[s] : 0, 0, nkb, 0, 1};
  g.nseg = ntaps;
  }
  g.w = w; g.wN = N; g.wK = wK; g.wL = 1;
  g.T = int(T); g.B = Bn; g.n_tiles = (nvalid + BN - 1) / BN;
  g.epi.ptr[0] = out_bf16; g.epi.ptr[1] = bias; g.epi.ptr[2] = out_f32; g.epi.ptr[7] = const_cast<unsigned long long*>(d_step);
  g.epi.i[0] = ldo; g.epi.i[1] = act; g.epi.i[2] = nvalid; g.epi.i[3] = stream_id; g.epi.f[1] = pdrop; g.epi.seed = seed;
  return launch_act_gemm(EPI_BIAS_ACT, BN, g, st);
}

struct StepCtx {
  const TL* lo; uint8_t* ws; const uint8_t* pk; const float* params; cudaStream_t st; unsigned long long seed;
  const unsigned long long* d_step; int training;
};

// one LSTM step on the swapped GEMM: gates^T = Wrec[4H perm][K] x S[B][K]^T (+ pre / bias) -> cell + zoneout
int lstm_step(const StepCtx& s, const void* wrec, int H, int K, const void* state, int B, const float* pre, int pre_stride, const float* bias,
              const float* c_prev, float* c_out, const bf16* h_prev, int ld_hp, bf16* h_state, int ld_hs, bf16* h_out, int ld_ho,
              bf16* gst, bf16* tst, const int* lens, int t, int stream_id, float zone) {
  ActGemmCall g;
  memset(&g, 0, sizeof(g));
  g.a[0] = make_act(wrec, K, 4 * H, 1, 1, K); g.na = 1;
  g.seg[0] = Seg{0, 0, 0, K / kBK, 0, 1}; g.nseg = 1;
  g.w = state; g.wN = B; g.wK = K; g.wL = 1;
  g.T = 4 * H; g.B = 1; g.n_tiles = (B + 31) / 32;
  g.epi.ptr[0] = const_cast<float*>(pre); g.epi.ptr[1] = const_cast<float*>(bias); g.epi.ptr[2] = const_cast<float*>(c_prev); g.epi.ptr[3] = c_out;
  g.epi.ptr[4] = const_cast<bf16*>(h_prev); g.epi.ptr[5] = h_state; g.epi.ptr[6] = h_out; g.epi.ptr[7] = gst; g.epi.ptr[8] = tst;
  g.epi.ptr[9] = const_cast<int*>(lens); g.epi.ptr[10] = const_cast<unsigned long long*>(s.d_step);
  g.epi.i[0] = H; g.epi.i[1] = B; g.epi.i[3] = pre_stride; g.epi.i[4] = ld_hp; g.epi.i[5] = ld_hs; g.epi.i[6] = ld_ho; g.epi.i[7] = t;
  g.epi.i[8] = stream_id; g.epi.i[9] = s.training; g.epi.f[0] = zone; g.epi.seed = s.seed;
  return launch_act_gemm(EPI_LSTM, 32, g, s.st);
}

int conv_block_fwd(const StepCtx& s, const ConvL& L, const void* x_in, long long T, int training) {
  const TL& lo = *s.lo;
  int shifts[8];
  for (int j = 0; j < L.k; ++j) shifts[j] = j - (L.k - 1) / 2;
  const int split = lo.c.split_bf16;
  bf16* y = reinterpret_cast<bf16*>(s.ws + L.w_y);
  float* yf = reinterpret_cast<float*>(s.ws + L.w_y);     // split mode keeps the pre-batch-norm activation in fp32
  int rc = conv_gemm(x_in, L.cin, T, lo.B, s.pk + L.k_w, L.cout, L.k * L.cinp * (split ? 3 : 1), L.k, shifts, L.cout % 256 == 0 ? 256 : 128,
                     const_cast<float*>(s.params + L.p_b), L.act, split ? nullptr : y,
                     split ? yf : nullptr, L.cout, L.cout, 0.f, 0, 0, nullptr, s.st, split);
  if (rc) return rc;
  float* stats = reinterpret_cast<float*>(s.ws + L.w_stats);
  const long long rows = (long long)lo.B * T;
  if (training) {
    T2_CHECK_CUDA(cudaMemsetAsync(stats, 0, 2 * L.cout * sizeof(float), s.st));
    if (split) bn_stats_kernel<float><<<64, 256, 0, s.st>>>(yf, stats, rows, L.cout);
    else bn_stats_kernel<bf16><<<64, 256, 0, s.st>>>(y, stats, rows, L.cout);
    t2_count_launch();
  }
  float* pp = const_cast<float*>(s.params);
  if (split)
    bn_apply_kernel<float><<<g1(rows * L.cout), 256, 0, s.st>>>(yf, reinterpret_cast<bf16*>(s.ws + L.w_x), stats, s.params + L.p_gamma, s.params + L.p_beta,
                                                         pp + L.p_mm, pp + L.p_mv, rows, L.cout, training, lo.c.dropout_rate, s.seed, s.d_step, L.stream, 1);
  else
    bn_apply_kernel<bf16><<<g1(rows * L.cout), 256, 0, s.st>>>(y, reinterpret_cast<bf16*>(s.ws + L.w_x), stats, s.params + L.p_gamma, s.params + L.p_beta,
                                                       pp + L.p_mm, pp + L.p_mv, rows, L.cout, training, lo.c.dropout_rate, s.seed, s.d_step,
                                                       L.stream, 0);
  t2_count_launch();
  T2_CHECK_CUDA(cudaGetLastError());
  return T2_OK;
}


// ======================================================================================================
// backward
// ======================================================================================================
// fixed-order list of weight-gradient GEMM launches; tiles live in the workspace (uploaded by t2_taco_init)
struct WgL { std::vector<WgradTile> tiles; };
void wg_tile(std::vector<WgradTile>& v, int am, int ach, int ash, int bm, int bch, long long off, int ldc, int mv, int nv, float scale = 1.f) {
  WgradTile t; memset(&t, 0, sizeof(t));
  t.a_map = am; t.a_ch0 = ach; t.a_shift = ash; t.b_map = bm; t.b_ch0 = bch; t.out_off = off; t.ldc = ldc;
  t.m_valid = mv; t.n_valid = nv; t.scale = scale; t.accumulate = 0; t.div = nullptr;
  v.push_back(t);
}
// dense [Ca x Cb] gradient of a 1x1 map: A channels [a0, a0+Ca), B channels [b0, b0+Cb)
void wg_dense(std::vector<WgradTile>& v, int am, int a0, int Ca, int bm, int b0, int Cb, long long off, int ldc, int shift = 0) {
  for (int m0 = 0; m0 < Ca; m0 += 128)
    for (int n0 = 0; n0 < Cb; n0 += 256)   // wgrad_gemm_kernel tiles: 128 x (up to) 256
      wg_tile(v, am, a0 + m0, shift, bm, b0 + n0, off + (long long)m0 * ldc + n0, ldc, Ca - m0 < 128 ? Ca - m0 : 128, Cb - n0 < 256 ? Cb - n0 : 256);
}
enum { WG_PP = 0, WG_POST0 = 1 /* .. +postnet layers */ };
void build_tiles(const TL& lo, std::vector<WgL>& L) {
  L.clear();
  const int H = lo.H, D = lo.D, K1r = 2 * H + D, K2 = 2 * D, PIK = D + 2 * H;
  auto conv = [&](const ConvL& c) {
    WgL w;
    for (int j = 0; j < c.k; ++j) wg_dense(w.tiles, 0, 0, c.cin, 1, 0, c.cout, c.p_k + (long long)j * c.cin * c.cout, c.cout, j - (c.k - 1) / 2);
    L.push_back(w);
  };
  { WgL w; wg_dense(w.tiles, 0, 0, lo.PC, 1, 0, lo.M, lo.p_ppk, lo.M); L.push_back(w); }            // 0: postnet projection
  for (int i = int(lo.post.size()) - 1; i >= 0; --i) conv(lo.post[i]);                                // 1..: postnet convs (reverse)
  { WgL w; wg_dense(w.tiles, 0, 0, PIK, 1, 0, lo.M, lo.p_fk, lo.M); wg_dense(w.tiles, 0, 0, PIK, 1, lo.M, 1, lo.p_sk, 1); L.push_back(w); }  // proj
  { WgL w;                                                                                               // decoder LSTMs + prenet-to-LSTM
    wg_dense(w.tiles, 0, 0, K2, 1, 0, 4 * D, lo.p_l2k, 4 * D);                                          // maps: 0 S2, 1 dg2, 2 S1, 3 dg1, 4 pn2
    wg_dense(w.tiles, 2, 0, K1r, 3, 0, 4 * D, lo.p_l1k + (long long)lo.P2 * 4 * D, 4 * D);
    wg_dense(w.tiles, 4, 0, lo.P2, 3, 0, 4 * D, lo.p_l1k, 4 * D);
    L.push_back(w); }
  { WgL w;                                                                                               // prenet + query layer
    wg_dense(w.tiles, 0, 0, lo.P1, 1, 0, lo.P2, lo.p_p2k, lo.P2);                                       // maps: 0 pn1, 1 dz2, 2 decin, 3 dz1, 4 PI, 5 dq_all
    wg_dense(w.tiles, 2, 0, lo.M, 3, 0, lo.P1, lo.p_p1k, lo.P1);
    wg_dense(w.tiles, 4, 0, D, 5, 0, lo.A, lo.p_qry, lo.A);
    L.push_back(w); }
  { WgL w; wg_dense(w.tiles, 0, 0, 2 * H, 1, 0, lo.A, lo.p_mem, lo.A); L.push_back(w); }               // memory layer: values x dkeys
  for (int d = 0; d < 2; ++d) {                                                                          // encoder LSTM d
    WgL w;  // maps: 0 h history (time-major), 1 gate grads time-major, 2 x3 (batch-major), 3 gate grads batch-major
    wg_dense(w.tiles, 0, 0, H, 1, 0, 4 * H, lo.p_elk[d] + (long long)lo.C * 4 * H, 4 * H);
    L.push_back(w);
    WgL w2; wg_dense(w2.tiles, 0, 0, lo.C, 1, 0, 4 * H, lo.p_elk[d], 4 * H); L.push_back(w2);
  }
  for (int i = int(lo.enc.size()) - 1; i >= 0; --i) conv(lo.enc[i]);
}

// loss seeds: dmel = 2 (mel - tgt) / N * [not clipped] (bf16, 128-col padded) ; ddec_direct = dmel + 2 (dec - tgt) / N
__global__ void loss_seed_kernel(const float* __restrict__ dec_f, const float* __restrict__ resid, const float* __restrict__ mel,
                                 const float* __restrict__ tgt, bf16* __restrict__ dmel, float* __restrict__ ddec, long long npos, int M,
                                 int clip, float lo, float hi, const int* __restrict__ tlen, int To, const float* __restrict__ scal,
                                 const float* __restrict__ extra) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= npos * 128) return;
  const long long pos = e / 128; const int m = int(e % 128);
  float g = 0.f;
  if (m < M) {
    const long long o = pos * M + m;
    const float n = scal[5];
    const float raw = dec_f[o] + resid[pos * 128 + m];
    const bool masked = tlen && int(pos % To) >= tlen[pos / To];             // masked frame: no mel-loss gradient
    g = masked ? 0.f : 2.f * (mel[o] - tgt[o]) / n;
    if (extra) g += extra[o];          // gradient of the post-processing net w.r.t. the clipped mel_outputs (not masked: its convs / GRU mix frames)
    if (clip && (raw < lo || raw > hi)) g = 0.f;
    ddec[o] = g + (masked ? 0.f : 2.f * (dec_f[o] - tgt[o]) / n);
  }
  dmel[e] = __float2bfloat16(g);
}
// ddec_tm[t][b][0..M) = (ddec_direct + ddec_post)[b][t][:] * [decoder clip inactive] ; col M = d BCE / d stop logit
__global__ void ddec_tm_kernel(const float* __restrict__ ddec, const bf16* __restrict__ dpost, const float* __restrict__ projo,
                               const float* __restrict__ stop_tgt, bf16* __restrict__ out, int B, int To, int M, int clip, float lo, float hi,
                               const int* __restrict__ tlen, float pos_w, const float* __restrict__ scal) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= (long long)To * B * 128) return;
  const int m = int(e % 128), b = int((e / 128) % B), t = int(e / (128LL * B));
  float g = 0.f;
  if (m < M) {
    const long long o = ((long long)b * To + t) * M + m;
    g = ddec[o] + __bfloat162float(dpost[o]);
    const float raw = projo[((long long)t * B + b) * 128 + m];
    if (clip && (raw < lo || raw > hi)) g = 0.f;
  } else if (m == M) {
    const float x = projo[((long long)t * B + b) * 128 + M];
    const float z = stop_tgt[(long long)b * To + t];
    if (!tlen) g = (1.f / (1.f + __expf(-x)) - z) / scal[6];
    else if (t < tlen[b]) g = ((1.f - z) - (1.f + (pos_w - 1.f) * z) / (1.f + __expf(x))) / scal[6];     // d/dx of the weighted CE
  }
  out[e] = __float2bfloat16(g);
}
// d(pre-activation) of relu + inverted dropout given the stored post-dropout output y: dz = y > 0 ? d / keep : 0
__global__ void relu_drop_bwd_kernel(const bf16* __restrict__ d, const bf16* __restrict__ y, bf16* __restrict__ dz, long long n, float p) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < n) dz[e] = __float2bfloat16(__bfloat162float(y[e]) > 0.f ? __bfloat162float(d[e]) / (1.f - p) : 0.f);
}
__global__ void f32_to_bf16_k(const float* __restrict__ in, bf16* __restrict__ out, long long n) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < n) out[e] = __float2bfloat16(in[e]);
}
__global__ void reg_grad_kernel(const float* __restrict__ params, float* __restrict__ grads, const long long* __restrict__ tab, float w) {
  const long long off = tab[2 * blockIdx.y], len = tab[2 * blockIdx.y + 1];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < len; i += (long long)gridDim.x * blockDim.x)
    grads[off + i] += w * params[off + i];
}

// backward of one LSTM cell + zoneout (see EPI_LSTM): produces the pre-activation gate gradients and the state grads
struct CellBwd {
  float* dh_ext; long long ld_ext;            // grad wrt the un-zoned output h_new: dh_ext[b*ld_ext + u]
  int zero_ext;                                // clear dh_ext after reading (its producer accumulates atomically, split-K)
  float* dhs; float* dcs;                      // [B][H] running grads wrt the carried (zoned) state (in/out)
  const bf16* gst; const bf16* tst; const float* c_prev;
  bf16* dg_a; long long ld_a;                  // gate grads, gate-major [4H] per batch row (row stride ld_a)
  bf16* dg_b; long long ld_b;                  // optional second copy
  const int* lens; int t, B, H, stream;
  float zone; unsigned long long seed; const unsigned long long* step;
};
__global__ void lstm_cell_bwd_kernel(CellBwd a) {
  pdl_wait();
  pdl_launch_dependents();
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= (long long)a.B * a.H) return;
  const int b = int(e / a.H), u = int(e % a.H);
  unsigned long long seed = a.seed + (a.step ? *a.step : 0ull);
  const bool live = a.lens ? (a.t < a.lens[b]) : true;
  float dzi = 0.f, dzj = 0.f, dzf = 0.f, dzo = 0.f;
  if (live) {
    const bf16* g = a.gst + (long long)b * 4 * a.H;
    const float gi = __bfloat162float(g[u]), gj = __bfloat162float(g[a.H + u]), gf = __bfloat162float(g[2 * a.H + u]),
                go = __bfloat162float(g[3 * a.H + u]);
    const float tc = __bfloat162float(a.tst[e]);
    const uint64_t idx = (uint64_t(a.t) * a.B + b) * a.H + u;
    const bool mc = a.zone <= 0.f || hash_uniform32(hash_seed(seed, uint32_t(a.stream) * 2u), idx) >= a.zone;
    const bool mh = a.zone <= 0.f || hash_uniform32(hash_seed(seed, uint32_t(a.stream) * 2u + 1u), idx) >= a.zone;
    const float dhs = a.dhs[e], dcs = a.dcs[e];
    const float dh_new = a.dh_ext[(long long)b * a.ld_ext + u] + (mh ? dhs : 0.f);
    if (a.zero_ext) a.dh_ext[(long long)b * a.ld_ext + u] = 0.f;
    const float dc_new = (mc ? dcs : 0.f) + dh_new * go * (1.f - tc * tc);
    dzo = dh_new * tc * go * (1.f - go);
    dzi = dc_new * gj * gi * (1.f - gi);
    dzj = dc_new * gi * (1.f - gj * gj);
    dzf = dc_new * a.c_prev[e] * gf * (1.f - gf);
    a.dcs[e] = dc_new * gf + (mc ? 0.f : dcs);
    a.dhs[e] = mh ? 0.f : dhs;
  }
  bf16* o = a.dg_a + (long long)b * a.ld_a;
  o[u] = __float2bfloat16(dzi); o[a.H + u] = __float2bfloat16(dzj); o[2 * a.H + u] = __float2bfloat16(dzf); o[3 * a.H + u] = __float2bfloat16(dzo);
  if (a.dg_b) {
    bf16* o2 = a.dg_b + (long long)b * a.ld_b;
    o2[u] = __float2bfloat16(dzi); o2[a.H + u] = __float2bfloat16(dzj); o2[2 * a.H + u] = __float2bfloat16(dzf); o2[3 * a.H + u] = __float2bfloat16(dzo);
  }
}

// backward step GEMM: dS^T [K rows][B] = W^T-packed [K][4H] x dgates [B][4H]^T, rows split over two fp32 destinations
int lstm_bwd_gemm(const StepCtx& s, const void* wT, int K, int H4, const void* dg, int B, float* dst0, int rows0, int ld0, int acc0, float* dst1,
                  int ld1, int acc1, int ksplit) {
  ActGemmCall g;
  memset(&g, 0, sizeof(g));
  g.a[0] = make_act(wT, H4, K, 1, 1, H4); g.na = 1;
  g.seg[0] = Seg{0, 0, 0, H4 / kBK, 0, 1}; g.nseg = 1;
  g.w = dg; g.wN = B; g.wK = H4; g.wL = 1;
  g.T = K; g.B = 1; g.n_tiles = (B + 31) / 32;
  // few output tiles (K / 128 <= 16) but a long reduction (4H): slice the reduction over `ksplit` CTAs per tile
  g.ksplit = ksplit;
  g.epi.ptr[0] = dst0; g.epi.ptr[1] = dst1;
  g.epi.i[0] = rows0; g.epi.i[1] = ld0; g.epi.i[2] = acc0; g.epi.i[3] = K; g.epi.i[4] = ld1; g.epi.i[5] = acc1; g.epi.i[6] = B;
  return launch_act_gemm(EPI_TOUT, 32, g, s.st);
}

// ---- attention backward, one CTA per batch item per step ----------------------------------------------------------
struct AttBwd {
  const bf16* h2out; int ld_h2; const bf16* WqT; const float* Wq;   // Wq fp32 [D][A]
  const float* U; const float* v;
  const float* keys; const bf16* values; const int* lens;
  const float* alpha;      // [B][Ti] of this step
  float* cumrun;           // [B][Ti]: cum_t on entry, cum_{t-1} on exit
  float* dcum;             // [B][Ti] running grad wrt cum_t (in) / cum_{t-1} (out)
  const float* dPI; int ld_dPI;   // dPI_all[t]: [B][PIK] fp32
  float* dctxl;            // [B][C2] grad wrt ctx_t from LSTM-1 of step t+1 (read, then cleared for the split-K accumulation)
  float* dh2ext;           // [B][D] out: grad wrt the un-zoned LSTM-2 output of this step
  bf16* dctx_save;         // [B][C2]
  bf16* dq_save;           // [B][A]
  float* dkeys;            // [B][Ti][A] accumulated
  float* acc;              // per item: dU [(KA+1)][A] (row KA = d u0) | dv [A]
  int B, Ti, D, A, KA, C2;
};
inline size_t att_bwd_smem(int Ti, int KA, int A, int D, int C2) {
  const int Tip = (Ti + 3) & ~3;
  return sizeof(float) * (size_t)((KA + 1) * (A + kAttPad) + att_cumlen(Ti, KA) + A + 4 * Tip + D + C2 + 2 * A +
                                  att_erows(Ti, KA) * (A + kAttPad) + 32) + 64;
}
__global__ void __launch_bounds__(kAttThreads) att_bwd_kernel(AttBwd a) {
  extern __shared__ __align__(16) float sm[];
  pdl_wait();
  pdl_launch_dependents();
  ATT_STAMP(16);
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int Ti = a.Ti, A = a.A, AP = A + kAttPad, half = a.KA / 2, NW = kAttThreads / 32;
  const int Tip = (Ti + 3) & ~3, cumlen = att_cumlen(Ti, a.KA), RE = att_erows(Ti, a.KA);
  float* Us = sm;                         // [(KA+1)][AP]
  float* cum = Us + (a.KA + 1) * AP;      // [cumlen] cum_{t-1}, zero-padded
  float* q = cum + cumlen;                // [A]
  float* al = q + A;                      // [Ti]
  float* de = al + Tip;                   // [Ti]
  float* dcs = de + Tip;                  // [Ti] incoming dcum_t
  float* dca = dcs + Tip;                 // [Ti] location-path contribution to dcum_{t-1}
  float* hs = dca + Tip;                  // [D]
  float* dctx = hs + a.D;                 // [C2]
  float* dq = dctx + a.C2;                // [A]
  float* dv = dq + A;                     // [A]
  float* dE = dv + A;                     // [RE][AP]: position j lives in row j + half; everything else stays zero
  float* red = dE + RE * AP;              // [32]
  const int len = a.lens[b];
  for (int i = tid; i < (a.KA + 1) * A; i += kAttThreads) Us[(i / A) * AP + (i % A)] = a.U[i];
  for (int i = tid; i < cumlen; i += kAttThreads) {
    const int j = i - half;
    float cp = 0.f;
    if (j >= 0 && j < Ti) {
      const float aj = a.alpha[(long long)b * Ti + j];
      al[j] = aj;
      cp = a.cumrun[(long long)b * Ti + j] - aj;
      a.cumrun[(long long)b * Ti + j] = cp;
      dcs[j] = a.dcum[(long long)b * Ti + j];
      dca[j] = 0.f;
    }
    cum[i] = cp;
  }
  for (int i = tid; i < a.D; i += kAttThreads) hs[split8(i, a.D)] = __bfloat162float(a.h2out[(long long)b * a.ld_h2 + i]);
  for (int c = tid; c < a.C2; c += kAttThreads) {
    const float gg = a.dPI[(long long)b * a.ld_dPI + a.D + c] + a.dctxl[(long long)b * a.C2 + c];
    a.dctxl[(long long)b * a.C2 + c] = 0.f;
    dctx[split8(c, a.C2)] = gg;            // split layout: dotted against 16-byte bf16 pieces of the values rows
    a.dctx_save[(long long)b * a.C2 + c] = __float2bfloat16(gg);
  }
  for (int i = tid; i < A; i += kAttThreads) { dq[i] = 0.f; dv[i] = 0.f; }
  for (int i = tid * 4; i < RE * AP; i += kAttThreads * 4) *reinterpret_cast<float4*>(dE + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  ATT_STAMP(17);
  att_query(a.WqT, hs, q, A, a.D);
  ATT_STAMP(18);
  // d alpha[j] = dctx . values[j] + dcum[j] : one warp per row, 8-wide bf16 loads, four rows in flight
  float part = 0.f;
  {
    const int nch = a.C2 >> 3;
    const uint4* vb = reinterpret_cast<const uint4*>(a.values + (long long)b * Ti * a.C2);
    for (int j = warp; j < len; j += 4 * NW) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int ch = lane; ch < nch; ch += 32) {
        uint4 u[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int jr = j + r * NW;
          u[r] = jr < len ? __ldg(vb + (long long)jr * nch + ch) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r] += dot8s(u[r], dctx, ch, a.C2);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = warp_sum(acc[r]);
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int jr = j + r * NW;
          if (jr < len) { const float da = acc[r] + dcs[jr]; de[jr] = da; part += al[jr] * da; }
        }
      }
    }
    for (int j = len + tid; j < Ti; j += kAttThreads) de[j] = 0.f;
  }
  if (lane == 0) red[warp] = part;
  __syncthreads();
  ATT_STAMP(19);
  float dot = 0.f;
  for (int w = 0; w < NW; ++w) dot += red[w];
  __syncthreads();
  for (int j = tid; j < Ti; j += kAttThreads) de[j] = j < len ? al[j] * (de[j] - dot) : 0.f;
  __syncthreads();
  // energies backward: units of 16 rows x 64 channels (pl recomputed on the tensor cores, see loc_tile)
  {
    const int n_nh = A >> 6, n_units = ((len + 15) >> 4) * n_nh;
    for (int u = warp; u < n_units; u += NW) {
      const int j0 = (u / n_nh) * 16, n0 = (u % n_nh) * 64;
      float acc[8][4];
      loc_tile(Us, AP, cum, a.KA, j0, n0, acc);
      const int r0 = j0 + g, r1 = r0 + 8;
      const bool ok0 = r0 < len, ok1 = r1 < len;
      const long long kr0 = ((long long)b * Ti + (ok0 ? r0 : 0)) * A + n0 + 2 * t, kr1 = ((long long)b * Ti + (ok1 ? r1 : 0)) * A + n0 + 2 * t;
      float2 ky0[8], ky1[8];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        ky0[nt] = __ldg(reinterpret_cast<const float2*>(a.keys + kr0 + nt * 8));
        ky1[nt] = __ldg(reinterpret_cast<const float2*>(a.keys + kr1 + nt * 8));
      }
      const float de0 = ok0 ? de[r0] : 0.f, de1 = ok1 ? de[r1] : 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int c = n0 + nt * 8 + 2 * t;
        const float2 qq = *reinterpret_cast<const float2*>(q + c);
        const float2 vv = __ldg(reinterpret_cast<const float2*>(a.v + c));
        const float t00 = tanhf_(ky0[nt].x + qq.x + acc[nt][0]), t01 = tanhf_(ky0[nt].y + qq.y + acc[nt][1]);
        const float t10 = tanhf_(ky1[nt].x + qq.x + acc[nt][2]), t11 = tanhf_(ky1[nt].y + qq.y + acc[nt][3]);
        const float2 d0 = make_float2(de0 * vv.x * (1.f - t00 * t00), de0 * vv.y * (1.f - t01 * t01));
        const float2 d1 = make_float2(de1 * vv.x * (1.f - t10 * t10), de1 * vv.y * (1.f - t11 * t11));
        // column sums over the unit's 16 rows: this lane's two rows, then the 8 row groups (lanes with equal t)
        float sq0 = d0.x + d1.x, sq1 = d0.y + d1.y, sv0 = de0 * t00 + de1 * t10, sv1 = de0 * t01 + de1 * t11;
#pragma unroll
        for (int m = 4; m < 32; m <<= 1) {
          sq0 += __shfl_xor_sync(0xffffffffu, sq0, m); sq1 += __shfl_xor_sync(0xffffffffu, sq1, m);
          sv0 += __shfl_xor_sync(0xffffffffu, sv0, m); sv1 += __shfl_xor_sync(0xffffffffu, sv1, m);
        }
        if (g == 0) { atomicAdd(&dq[c], sq0); atomicAdd(&dq[c + 1], sq1); atomicAdd(&dv[c], sv0); atomicAdd(&dv[c + 1], sv1); }
        if (ok0) {
          atomicAdd(reinterpret_cast<float2*>(a.dkeys + kr0 + nt * 8), d0);    // accumulates over decoder steps, no read-back
          *reinterpret_cast<float2*>(dE + (r0 + half) * AP + c) = d0;
        }
        if (ok1) {
          atomicAdd(reinterpret_cast<float2*>(a.dkeys + kr1 + nt * 8), d1);
          *reinterpret_cast<float2*>(dE + (r1 + half) * AP + c) = d1;
        }
      }
    }
  }
  __syncthreads();
  ATT_STAMP(20);
  float* accp = a.acc + (long long)b * ((a.KA + 2) * A);
  for (int c = tid; c < A; c += kAttThreads) {
    a.dq_save[(long long)b * A + c] = __float2bfloat16(dq[c]);
    accp[a.KA * A + c] += dq[c];          // d u0 (also the gradient of attention_bias)
    accp[(a.KA + 1) * A + c] += dv[c];
  }
  // dU[k][c] += sum_j cum_{t-1}[j + k - half] dE[j][c] = (T^T dE)[k][c]: M = 32 taps (2 m-tiles), N = A (one n-tile per
  // warp pass), K = memory rows
  {
    const int nks = (len + 7) >> 3;
    for (int nt = warp; nt < (A >> 3); nt += NW) {
      float acc[2][4];
#pragma unroll
      for (int m = 0; m < 2; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;
      for (int ks = 0; ks < nks; ++ks) {
        const int j = ks * 8 + t;
        const uint32_t b0 = f2tf32(dE[(j + half) * AP + nt * 8 + g]), b1 = f2tf32(dE[(j + 4 + half) * AP + nt * 8 + g]);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          const int k = m * 16 + g;
          uint32_t af[4];
          af[0] = f2tf32(k < a.KA ? cum[j + k] : 0.f); af[1] = f2tf32(k + 8 < a.KA ? cum[j + k + 8] : 0.f);
          af[2] = f2tf32(k < a.KA ? cum[j + 4 + k] : 0.f); af[3] = f2tf32(k + 8 < a.KA ? cum[j + 4 + k + 8] : 0.f);
          mma_tf32(acc[m], af, b0, b1);
        }
      }
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const int k0 = m * 16 + g, k1 = k0 + 8, c = nt * 8 + 2 * t;
        if (k0 < a.KA) atomicAdd(reinterpret_cast<float2*>(accp + k0 * A + c), make_float2(acc[m][0], acc[m][1]));
        if (k1 < a.KA) atomicAdd(reinterpret_cast<float2*>(accp + k1 * A + c), make_float2(acc[m][2], acc[m][3]));
      }
    }
  }
  ATT_STAMP(21);
  // dcum_{t-1}[i] = dcum_t[i] + sum_k P[i - k + 2*half][k] with P = dE U^T ([RE rows] x [32 taps], K = A): one m-tile of
  // dE rows per warp pass, the anti-diagonal sums go through shared-memory atomics
  {
    for (int mt = warp; mt < (RE >> 4); mt += NW) {
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      if (mt * 16 >= len + 2 * half) continue;     // rows past the last written position are zero (warp-uniform)
      float acc[4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
      for (int ks = 0; ks < (A >> 3); ++ks) {
        const int c = ks * 8 + t;
        uint32_t af[4];
        af[0] = f2tf32(dE[r0 * AP + c]); af[1] = f2tf32(dE[r1 * AP + c]);
        af[2] = f2tf32(dE[r0 * AP + c + 4]); af[3] = f2tf32(dE[r1 * AP + c + 4]);
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const int tap = n * 8 + g;
          const uint32_t b0 = tap < a.KA ? f2tf32(Us[tap * AP + c]) : 0u, b1 = tap < a.KA ? f2tf32(Us[tap * AP + c + 4]) : 0u;
          mma_tf32(acc[n], af, b0, b1);
        }
      }
#pragma unroll
      for (int n = 0; n < 4; ++n) {
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const int tap = n * 8 + 2 * t + (x & 1), r = (x & 2) ? r1 : r0;
          const int i = r + tap - 2 * half;
          if (tap < a.KA && i >= 0 && i < Ti) atomicAdd(&dca[i], acc[n][x]);
        }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < Ti; i += kAttThreads) a.dcum[(long long)b * Ti + i] = dcs[i] + dca[i];
  ATT_STAMP(22);
  // dh2ext = dPI[:, 0:D] + dq . Wq^T   (bf16 [A][D] copy of the query weights: coalesced along D)
  for (int k2 = tid; k2 < (a.D >> 1); k2 += kAttThreads) {
    const float2 gg = *reinterpret_cast<const float2*>(a.dPI + (long long)b * a.ld_dPI + 2 * k2);
    float a0 = gg.x, a1 = gg.y;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a.WqT) + k2;
#pragma unroll 16
    for (int c = 0; c < A; ++c) {
      const uint32_t u = __ldg(w + (long long)c * (a.D >> 1));
      a0 += dq[c] * bf16lo(u); a1 += dq[c] * bf16hi(u);
    }
    *reinterpret_cast<float2*>(a.dh2ext + (long long)b * a.D + 2 * k2) = make_float2(a0, a1);
  }
  ATT_STAMP(23);
}
// after the loop: reduce the per-item accumulators over the batch and push dU / du0 / dv through the U = K . Wl
// factorisation: dK = dU Wl^T, dWl = K^T dU + bK (x) du0, dbK = Wl du0, d attention_bias = du0, dv
__global__ void att_finish_kernel(const float* __restrict__ acc, const float* __restrict__ K, const float* __restrict__ bK,
                                  const float* __restrict__ Wl, float* __restrict__ grads, int B, int KA, int F, int A, long long o_k,
                                  long long o_bk, long long o_wl, long long o_v, long long o_ba, float* __restrict__ scratch) {
  // phase 1 (all blocks): reduce over the batch into scratch [(KA+2)][A]
  const int n = (KA + 2) * A;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += acc[(long long)b * n + i];
    scratch[i] = s;
  }
}
__global__ void att_finish2_kernel(const float* __restrict__ dU, const float* __restrict__ K, const float* __restrict__ bK,
                                   const float* __restrict__ Wl, float* __restrict__ grads, int KA, int F, int A, long long o_k, long long o_bk,
                                   long long o_wl, long long o_v, long long o_ba) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float* du0 = dU + KA * A;
  const float* dv = dU + (KA + 1) * A;
  if (i < KA * F) {            // dK[k][f] = sum_a dU[k][a] Wl[f][a]
    const int k = i / F, f = i % F;
    float s = 0.f;
    for (int c = 0; c < A; ++c) s += dU[k * A + c] * Wl[f * A + c];
    grads[o_k + i] += s;
  } else if (i < KA * F + F * A) {   // dWl[f][a] = sum_k K[k][f] dU[k][a] + bK[f] du0[a]
    const int j = i - KA * F, f = j / A, c = j % A;
    float s = bK[f] * du0[c];
    for (int k = 0; k < KA; ++k) s += K[k * F + f] * dU[k * A + c];
    grads[o_wl + j] += s;
  } else if (i < KA * F + F * A + F) {   // dbK[f] = sum_a Wl[f][a] du0[a]
    const int f = i - KA * F - F * A;
    float s = 0.f;
    for (int c = 0; c < A; ++c) s += Wl[f * A + c] * du0[c];
    grads[o_bk + f] += s;
  } else if (i < KA * F + F * A + F + A) {
    const int c = i - KA * F - F * A - F;
    grads[o_v + c] += dv[c];
    grads[o_ba + c] += du0[c];
  }
}
// dvalues[b][j][c] += sum_t alpha[t][b][j] * dctx[t][b][c]; then apply the memory mask in place
__global__ void dvalues_ctx_kernel(const float* __restrict__ alpha, const bf16* __restrict__ dctx, const int* __restrict__ lens,
                                   float* __restrict__ dvalues, int B, int Ti, int To, int C2) {
  const int b = blockIdx.y, j = blockIdx.x;
  const bool live = j < lens[b];
  for (int c = threadIdx.x; c < C2; c += blockDim.x) {
    float acc = dvalues[((long long)b * Ti + j) * C2 + c];
    if (live) for (int t = 0; t < To; ++t) acc += alpha[((long long)t * B + b) * Ti + j] * __bfloat162float(dctx[((long long)t * B + b) * C2 + c]);
    dvalues[((long long)b * Ti + j) * C2 + c] = live ? acc : 0.f;
  }
}

int conv_block_bwd(const StepCtx& s, const ConvL& L, const void* x_in, long long T, const bf16* dout, bf16* dpre, bf16* dx, float* grads,
                   const WgradTile* tiles, int ntiles) {
  const TL& lo = *s.lo;
  const long long rows = (long long)lo.B * T;
  float* stats = reinterpret_cast<float*>(s.ws + L.w_stats);
  float* bsum = stats + 4 * L.cout;
  const bf16* y = reinterpret_cast<const bf16*>(s.ws + L.w_y);
  T2_CHECK_CUDA(cudaMemsetAsync(bsum, 0, 2 * L.cout * sizeof(float), s.st));
  bn_bwd_stats_kernel<<<64, 256, 0, s.st>>>(dout, y, stats, bsum, rows, L.cout, lo.c.dropout_rate, s.seed, s.d_step, L.stream); t2_count_launch();
  bn_bwd_apply_kernel<<<g1(rows * L.cout), 256, 0, s.st>>>(dout, y, stats, bsum, s.params + L.p_gamma, dpre, grads + L.p_gamma, grads + L.p_beta, rows,
                                                           L.cout, L.act, lo.c.dropout_rate, s.seed, s.d_step, L.stream); t2_count_launch();
  colsum_bf16_kernel<<<64, 256, 0, s.st>>>(dpre, rows, L.cout, L.cout, grads + L.p_b, 1.f); t2_count_launch();
  T2_CHECK_CUDA(cudaGetLastError());
  ActT maps[2] = {make_act(x_in, L.cin, int(T), lo.B), make_act(dpre, L.cout, int(T), lo.B)};
  int rc = launch_wgrad(maps, 2, tiles, ntiles, grads, int(T), lo.B, s.st);
  if (rc) return rc;
  if (dx) {
    int shifts[8];
    for (int j = 0; j < L.k; ++j) shifts[j] = (L.k - 1) / 2 - j;
    rc = conv_gemm(dpre, L.cout, T, lo.B, s.pk + L.k_wT, L.cin, L.k * L.cout, L.k, shifts, L.cin % 256 == 0 ? 256 : 128, nullptr, 0, dx, nullptr, L.cin,
                   L.cin, 0.f, 0, 0, nullptr, s.st);
    if (rc) return rc;
  }
  return T2_OK;
}

}  // namespace
}  // namespace t2

using namespace t2;

extern "C" int t2_taco_sizes(const t2_taco_config_t* cfg, long long* n_params, long long* packed_bytes, long long* workspace_bytes,
                             int* n_tensors) {
  TL lo;
  int rc = build(cfg, lo, nullptr);
  if (rc) return rc;
  *n_params = lo.n_params; *packed_bytes = lo.packed_bytes; *workspace_bytes = lo.workspace_bytes; *n_tensors = int(lo.params.size());
  return T2_OK;
}

extern "C" int t2_taco_param_info(const t2_taco_config_t* cfg, int i, char* name, int cap, long long* offset, int* ndim, int* shape4,
                                  int* trainable) {
  TL lo;
  int rc = build(cfg, lo, nullptr);
  if (rc) return rc;
  T2_REQUIRE(i >= 0 && i < int(lo.params.size()), T2_ERR_INVALID_ARG, "tensor index out of range");
  const PT& p = lo.params[i];
  snprintf(name, cap, "%s", p.name.c_str());
  *offset = p.off; *ndim = p.ndim; *trainable = p.trainable ? 1 : 0;
  for (int k = 0; k < 4; ++k) shape4[k] = p.shape[k];
  return T2_OK;
}

extern "C" int t2_taco_init(const t2_taco_config_t* cfg, void* d_packed, void* d_workspace, void* stream) {
  TL lo;
  std::vector<PJ> jobs;
  int rc = build(cfg, lo, &jobs);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  T2_CHECK_CUDA(cudaMemsetAsync(d_packed, 0, lo.packed_bytes, st));
  T2_CHECK_CUDA(cudaMemsetAsync(d_workspace, 0, lo.workspace_bytes, st));
  T2_CHECK_CUDA(cudaMemcpyAsync(ws + lo.w_packjobs, jobs.data(), jobs.size() * sizeof(PJ), cudaMemcpyHostToDevice, st));
  std::vector<long long> reg;
  for (auto& p : lo.params)
    if (p.reg) { long long n = 1; for (int k = 0; k < p.ndim; ++k) n *= p.shape[k]; reg.push_back(p.off); reg.push_back(n); }
  T2_CHECK_CUDA(cudaMemcpyAsync(ws + lo.w_regtab, reg.data(), reg.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
  std::vector<WgL> wl;
  build_tiles(lo, wl);
  std::vector<WgradTile> all;
  for (auto& w : wl) all.insert(all.end(), w.tiles.begin(), w.tiles.end());
  T2_REQUIRE(all.size() <= 16384, T2_ERR_UNSUPPORTED_SHAPE, "too many weight-gradient tiles (%d)", int(all.size()));
  T2_CHECK_CUDA(cudaMemcpyAsync(ws + lo.w_tiles, all.data(), all.size() * sizeof(WgradTile), cudaMemcpyHostToDevice, st));
  T2_CHECK_CUDA(cudaStreamSynchronize(st));
  return T2_OK;
}

extern "C" int t2_taco_pack_weights(const t2_taco_config_t* cfg, const float* d_params, void* d_packed, void* d_workspace, void* stream) {
  TL lo;
  int rc = build(cfg, lo, nullptr);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  tpack_kernel<<<dim3(32, lo.n_packjobs), dim3(32, 8), 0, st>>>(d_params, static_cast<bf16*>(d_packed),
                                                        reinterpret_cast<const PJ*>(static_cast<uint8_t*>(d_workspace) + lo.w_packjobs));
  t2_count_launch();
  T2_CHECK_CUDA(cudaGetLastError());
  return T2_OK;
}

// The two directions of the encoder BiLSTM are independent chains of small launches (8-16 CTAs each): the backward
// direction runs on a side stream (fork / join through events; capturable; T2_SIDE_STREAM=0 keeps one stream).
struct TacoSide { cudaStream_t s; cudaEvent_t fork, join; };
static TacoSide* taco_side() {
  static TacoSide ss;
  static int state = 0;   // 0 unknown, 1 ready, -1 disabled
  if (state == 0) {
    const char* e = getenv("T2_SIDE_STREAM");
    if (e && e[0] == '0') state = -1;
    else if (cudaStreamCreateWithFlags(&ss.s, cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming) == cudaSuccess) state = 1;
    else state = -1;
  }
  return state == 1 ? &ss : nullptr;
}

// ---- encoder: embedding -> conv blocks -> BiLSTM -> masked values -> attention keys (tacotron.py:113-131) ----
static int encoder_fwd(const StepCtx& s, const int* d_inputs, const int* d_input_lengths, int training) {
  const TL& lo = *s.lo;
  uint8_t* ws = s.ws; const uint8_t* pk = s.pk; const float* d_params = s.params; cudaStream_t st = s.st;
  const int B = lo.B, Ti = lo.Ti, H = lo.H;
  int rc;
  bf16* emb = reinterpret_cast<bf16*>(ws + lo.w_emb);
  embed_fwd_kernel<<<g1((long long)B * Ti * lo.E), 256, 0, st>>>(d_inputs, d_params + lo.p_emb, emb, (long long)B * Ti, lo.E, lo.c.split_bf16); t2_count_launch();
  const void* x = emb;
  for (auto& L : lo.enc) { rc = conv_block_fwd(s, L, x, Ti, training); if (rc) return rc; x = ws + L.w_x; }
  TacoSide* side = taco_side();
  if (side) {
    T2_CHECK_CUDA(cudaEventRecord(side->fork, st));
    T2_CHECK_CUDA(cudaStreamWaitEvent(side->s, side->fork, 0));
  }
  for (int d = 0; d < 2; ++d) {
    cudaStream_t sx = (d == 1 && side) ? side->s : st;
    StepCtx sc = s; sc.st = sx;
    float* pre = reinterpret_cast<float*>(ws + lo.w_encpre[d]);
    rc = conv_gemm(x, lo.C, Ti, B, pk + lo.k_encWx[d], 4 * H, lo.C * (lo.c.split_bf16 ? 3 : 1), 1, nullptr, 256, const_cast<float*>(d_params) + lo.p_elb[d], 0,
                   nullptr, pre, 4 * H, 4 * H, 0.f, 0, 0, nullptr, sx, lo.c.split_bf16);
    if (rc) return rc;
    bf16* hh = reinterpret_cast<bf16*>(ws + lo.w_ench[d]);
    float* cc = reinterpret_cast<float*>(ws + lo.w_encc[d]);
    T2_CHECK_CUDA(cudaMemsetAsync(hh, 0, (size_t)B * H * 2, sx));
    T2_CHECK_CUDA(cudaMemsetAsync(cc, 0, (size_t)B * H * 4, sx));
    bf16* memory = reinterpret_cast<bf16*>(ws + lo.w_memory);
    for (int sidx = 0; sidx < Ti; ++sidx) {
      const int t = d == 0 ? sidx : Ti - 1 - sidx;
      rc = lstm_step(sc, pk + lo.k_encWr[d], H, H, hh + (long long)sidx * B * H, B, pre + (long long)t * 4 * H, Ti * 4 * H, nullptr,
                     cc + (long long)sidx * B * H, cc + (long long)(sidx + 1) * B * H, hh + (long long)sidx * B * H, H,
                     hh + (long long)(sidx + 1) * B * H, H, memory + (long long)t * 2 * H + d * H, Ti * 2 * H,
                     reinterpret_cast<bf16*>(ws + lo.w_encg[d]) + (long long)sidx * B * 4 * H,
                     reinterpret_cast<bf16*>(ws + lo.w_enct[d]) + (long long)sidx * B * H, d_input_lengths, t, 52 + d, lo.c.zoneout_rate);
      if (rc) return rc;
    }
  }
  if (side) {
    T2_CHECK_CUDA(cudaEventRecord(side->join, side->s));
    T2_CHECK_CUDA(cudaStreamWaitEvent(st, side->join, 0));
  }
  bf16* values = reinterpret_cast<bf16*>(ws + lo.w_values);
  mask_values_kernel<<<g1((long long)B * Ti * 2 * H), 256, 0, st>>>(reinterpret_cast<bf16*>(ws + lo.w_memory), d_input_lengths, values, B, Ti, 2 * H);
  t2_count_launch();
  float* keys = reinterpret_cast<float*>(ws + lo.w_keys);
  return conv_gemm(values, 2 * H, Ti, B, pk + lo.k_mem, lo.A, 2 * H, 1, nullptr, 128, nullptr, 0, nullptr, keys, lo.A, lo.A, 0.f, 0, 0, nullptr, st);
}

struct DecBufs { bf16 *S1, *S2, *PI, *values; float *c1, *c2, *cum, *attU, *keys, *pre1; size_t att_smem; int K1r, K2, PIK; };
// zero initial decoder state (Architecture_wrappers.py:134-167) + the merged location filter bank
static int decoder_reset(const StepCtx& s, DecBufs& d) {
  const TL& lo = *s.lo;
  uint8_t* ws = s.ws; const float* d_params = s.params; cudaStream_t st = s.st;
  const int B = lo.B, Ti = lo.Ti, H = lo.H, D = lo.D;
  d.K1r = 2 * H + D; d.K2 = 2 * D; d.PIK = D + 2 * H;
  d.S1 = reinterpret_cast<bf16*>(ws + lo.w_S1); d.S2 = reinterpret_cast<bf16*>(ws + lo.w_S2); d.PI = reinterpret_cast<bf16*>(ws + lo.w_PI);
  d.c1 = reinterpret_cast<float*>(ws + lo.w_c1); d.c2 = reinterpret_cast<float*>(ws + lo.w_c2);
  d.cum = reinterpret_cast<float*>(ws + lo.w_cum); d.attU = reinterpret_cast<float*>(ws + lo.w_attU);
  d.keys = reinterpret_cast<float*>(ws + lo.w_keys); d.values = reinterpret_cast<bf16*>(ws + lo.w_values);
  d.pre1 = reinterpret_cast<float*>(ws + lo.w_pre1);
  T2_CHECK_CUDA(cudaMemsetAsync(d.S1, 0, (size_t)B * d.K1r * 2, st));
  T2_CHECK_CUDA(cudaMemsetAsync(d.S2, 0, (size_t)B * d.K2 * 2, st));
  T2_CHECK_CUDA(cudaMemsetAsync(d.c1, 0, (size_t)B * D * 4, st));
  T2_CHECK_CUDA(cudaMemsetAsync(d.c2, 0, (size_t)B * D * 4, st));
  T2_CHECK_CUDA(cudaMemsetAsync(d.cum, 0, (size_t)B * Ti * 4, st));
  att_prep_kernel<<<g1((lo.KA + 1) * lo.A), 256, 0, st>>>(d_params + lo.p_lck, d_params + lo.p_lcb, d_params + lo.p_lfl, d_params + lo.p_ba, d.attU,
                                                       lo.KA, lo.F, lo.A); t2_count_launch();
  d.att_smem = att_fwd_smem(Ti, lo.KA, lo.A, D, 2 * H);
  T2_CHECK_CUDA(cudaFuncSetAttribute(att_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(d.att_smem)));
  T2_CHECK_CUDA(cudaGetLastError());
  return T2_OK;
}
// decoder step t: LSTM-1 (prenet part of its gates precomputed in pre1[t]), LSTM-2, attention (Architecture_wrappers.py:169-213)
static int decoder_step(const StepCtx& s, const DecBufs& d, const int* d_input_lengths, int t) {
  const TL& lo = *s.lo;
  uint8_t* ws = s.ws; const uint8_t* pk = s.pk; const float* d_params = s.params; cudaStream_t st = s.st;
  const int B = lo.B, Ti = lo.Ti, H = lo.H, D = lo.D, K1r = d.K1r, K2 = d.K2, PIK = d.PIK;
  bf16* S1t = d.S1 + (long long)t * B * K1r;
  bf16* S1n = d.S1 + (long long)(t + 1) * B * K1r;
  bf16* S2t = d.S2 + (long long)t * B * K2;
  bf16* S2n = d.S2 + (long long)(t + 1) * B * K2;
  bf16* PIt = d.PI + (long long)t * B * PIK;
  // LSTM 1: state operand [ctx_{t-1} | h1_{t-1}], input part precomputed in pre1[t]
  int rc = lstm_step(s, pk + lo.k_l1r, D, K1r, S1t, B, d.pre1 + (long long)t * B * 4 * D, 4 * D, nullptr, d.c1 + (long long)t * B * D,
                     d.c1 + (long long)(t + 1) * B * D, S1t + 2 * H, K1r, S1n + 2 * H, K1r, S2t, K2,
                     reinterpret_cast<bf16*>(ws + lo.w_g1) + (long long)t * B * 4 * D, reinterpret_cast<bf16*>(ws + lo.w_t1) + (long long)t * B * D,
                     nullptr, t, 54, lo.c.zoneout_rate);
  if (rc) return rc;
  // LSTM 2: state operand [h1out_t | h2_{t-1}]
  rc = lstm_step(s, pk + lo.k_l2, D, K2, S2t, B, nullptr, 0, d_params + lo.p_l2b, d.c2 + (long long)t * B * D, d.c2 + (long long)(t + 1) * B * D,
                 S2t + D, K2, S2n + D, K2, PIt, PIK, reinterpret_cast<bf16*>(ws + lo.w_g2) + (long long)t * B * 4 * D,
                 reinterpret_cast<bf16*>(ws + lo.w_t2) + (long long)t * B * D, nullptr, t, 55, lo.c.zoneout_rate);
  if (rc) return rc;
  AttArgs a;
  a.h2out = PIt; a.ld_h2 = PIK; a.WqT = reinterpret_cast<const bf16*>(pk + lo.k_qT);
  a.U = d.attU; a.v = d_params + lo.p_v;
  a.keys = d.keys; a.values = d.values; a.lens = d_input_lengths; a.cum = d.cum;
  a.alpha = reinterpret_cast<float*>(ws + lo.w_alpha) + (long long)t * B * Ti;
  a.ctx_a = S1n; a.ld_a = K1r; a.ctx_b = PIt + D; a.ld_b = PIK;
  a.B = B; a.Ti = Ti; a.D = D; a.A = lo.A; a.KA = lo.KA; a.C2 = 2 * H;
  T2_CHECK_CUDA(launch_pdl(att_fwd_kernel, dim3(B), dim3(kAttThreads), d.att_smem, st, a)); t2_count_launch();
  return T2_OK;
}

// forward + losses. d_inputs int32 [B][T_in]; d_input_lengths int32 [B]; d_mel_targets fp32 [B][T_out][M];
// d_stop_targets fp32 [B][T_out]. d_loss fp32[4] = {before, after, stop, regularisation} (already normalised).
extern "C" int t2_taco_forward(const t2_taco_config_t* cfg, float* d_params, const void* d_packed, void* d_workspace, const int* d_inputs,
                               const int* d_input_lengths, const float* d_mel_targets, const float* d_stop_targets, float* d_loss,
                               int training, unsigned long long seed, const unsigned long long* d_step, void* stream) {
  TL lo;
  int rc = build(cfg, lo, nullptr);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  const uint8_t* pk = static_cast<const uint8_t*>(d_packed);
  StepCtx s{&lo, ws, pk, d_params, st, seed, d_step, training};
  const int B = lo.B, Ti = lo.Ti, To = lo.To, H = lo.H, D = lo.D;
  float* scal = reinterpret_cast<float*>(ws + lo.w_scal);
  T2_CHECK_CUDA(cudaMemsetAsync(scal, 0, 16 * sizeof(float), st));
  const int* tlen = lo.c.mask_decoder ? reinterpret_cast<const int*>(ws + lo.w_tlen) : nullptr;     // t2_taco_set_target_lengths
  rc = encoder_fwd(s, d_inputs, d_input_lengths, training);
  if (rc) return rc;
  const void* x = nullptr;
  // ---- decoder: everything that does not depend on the recurrence is batched over time (teacher forcing) ----
  bf16* decin = reinterpret_cast<bf16*>(ws + lo.w_decin);
  decin_kernel<<<g1((long long)To * B * lo.M), 256, 0, st>>>(d_mel_targets, decin, B, To, lo.M); t2_count_launch();
  const long long TB = (long long)To * B;
  bf16* pn1 = reinterpret_cast<bf16*>(ws + lo.w_pn1);
  bf16* pn2 = reinterpret_cast<bf16*>(ws + lo.w_pn2);
  rc = conv_gemm(decin, lo.M, TB, 1, pk + lo.k_p1, lo.P1, 128, 1, nullptr, lo.P1 >= 256 ? 256 : 128, d_params + lo.p_p1b, 1, pn1, nullptr, lo.P1,
                 lo.P1, lo.c.dropout_rate, 20, seed, d_step, st);
  if (rc) return rc;
  rc = conv_gemm(pn1, lo.P1, TB, 1, pk + lo.k_p2, lo.P2, lo.P1, 1, nullptr, lo.P2 >= 256 ? 256 : 128, d_params + lo.p_p2b, 1, pn2, nullptr, lo.P2,
                 lo.P2, lo.c.dropout_rate, 21, seed, d_step, st);
  if (rc) return rc;
  float* pre1 = reinterpret_cast<float*>(ws + lo.w_pre1);
  rc = conv_gemm(pn2, lo.P2, TB, 1, pk + lo.k_l1x, 4 * D, lo.P2, 1, nullptr, 256, d_params + lo.p_l1b, 0, nullptr, pre1, 4 * D, 4 * D, 0.f, 0, 0,
                 nullptr, st);
  if (rc) return rc;
  const int PIK = D + 2 * H;
  DecBufs db;
  rc = decoder_reset(s, db);
  if (rc) return rc;
  bf16* PI = db.PI;
  for (int t = 0; t < To; ++t) { rc = decoder_step(s, db, d_input_lengths, t); if (rc) return rc; }
  T2_CHECK_CUDA(cudaGetLastError());
  // frame + stop projections for all steps at once
  float* projo = reinterpret_cast<float*>(ws + lo.w_projo);
  {
    // bias vector [M frames | 1 stop] lives in two parameter tensors: add them in the finishing kernel instead
    rc = conv_gemm(PI, PIK, TB, 1, pk + lo.k_proj, lo.M + 1, PIK, 1, nullptr, 128, nullptr, 0, nullptr, projo, 128, lo.M + 1, 0.f, 0, 0, nullptr, st);
    if (rc) return rc;
  }
  // add the projection biases in place (tiny), then clip / losses
  proj_bias_kernel<<<g1(TB * (lo.M + 1)), 256, 0, st>>>(projo, d_params + lo.p_fb, d_params + lo.p_sb, TB, lo.M); t2_count_launch();
  const float lo_c = -lo.c.max_abs_value - lo.c.lower_bound_decay, hi_c = lo.c.max_abs_value;
  bf16* dec_bm = reinterpret_cast<bf16*>(ws + lo.w_decbm);
  float* dec_f = reinterpret_cast<float*>(ws + lo.w_decf);
  dec_finish_kernel<<<g1((long long)B * To * (lo.M + 1)), 256, 0, st>>>(projo, d_mel_targets, d_stop_targets, dec_bm, dec_f,
                                                                        reinterpret_cast<float*>(ws + lo.w_stop), scal, B, To, lo.M,
                                                                        lo.c.clip_outputs, lo_c, hi_c, lo.c.split_bf16, tlen, lo.c.cross_entropy_pos_weight); t2_count_launch();
  // ---- postnet ----
  x = dec_bm;
  for (auto& L : lo.post) { rc = conv_block_fwd(s, L, x, To, training); if (rc) return rc; x = ws + L.w_x; }
  float* resid = reinterpret_cast<float*>(ws + lo.w_resid);
  rc = conv_gemm(x, lo.PC, To, B, pk + lo.k_pp, lo.M, lo.PC * (lo.c.split_bf16 ? 3 : 1), 1, nullptr, 128, d_params + lo.p_ppb, 0, nullptr, resid, 128, lo.M,
                 0.f, 0, 0, nullptr, st, lo.c.split_bf16);
  if (rc) return rc;
  mel_finish_kernel<<<g1((long long)B * To * lo.M), 256, 0, st>>>(dec_f, resid, d_mel_targets, reinterpret_cast<float*>(ws + lo.w_mel), scal,
                                                                  (long long)B * To, lo.M, lo.c.clip_outputs, lo_c, hi_c, tlen, To); t2_count_launch();
  reg_loss_kernel<<<dim3(8, lo.n_reg), 256, 0, st>>>(d_params, reinterpret_cast<const long long*>(ws + lo.w_regtab), lo.n_reg, scal); t2_count_launch();
  T2_CHECK_CUDA(cudaGetLastError());
  loss_norm_kernel<<<1, 1, 0, st>>>(scal, d_loss, float((long long)B * To * lo.M), float((long long)B * To), lo.c.reg_weight, tlen, B, To, lo.M);
  t2_count_launch();
  return T2_OK;
}

// ======================================================================================================
// free-running synthesis (TacoTestHelper, helpers.py:6-59; tacotron.py:150-200 with is_training = False)
// ======================================================================================================
// projo[t] += bias; next decoder input = the raw (un-clipped) frame just predicted (helpers.py:56)
__global__ void proj_bias_feedback_kernel(float* __restrict__ p, const float* __restrict__ fb, const float* __restrict__ sb, bf16* __restrict__ next_in,
                                          int B, int M) {
  pdl_wait();
  pdl_launch_dependents();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * (M + 1)) return;
  const int m = e % (M + 1), b = e / (M + 1);
  const float v = p[b * 128 + m] + (m < M ? fb[m] : sb[0]);
  p[b * 128 + m] = v;
  if (m < M && next_in) next_in[b * M + m] = __float2bfloat16(v);
}

// encoder + zero decoder state + go frame. inputs int32 [B][T_in], lengths int32 [B].
extern "C" int t2_taco_infer_begin(const t2_taco_config_t* cfg, float* d_params, const void* d_packed, void* d_workspace, const int* d_inputs,
                                   const int* d_input_lengths, void* stream) {
  TL lo;
  int rc = build(cfg, lo, nullptr);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  StepCtx s{&lo, ws, static_cast<const uint8_t*>(d_packed), d_params, st, 0ull, nullptr, 0};
  rc = encoder_fwd(s, d_inputs, d_input_lengths, 0);
  if (rc) return rc;
  DecBufs db;
  rc = decoder_reset(s, db);
  if (rc) return rc;
  T2_CHECK_CUDA(cudaMemsetAsync(ws + lo.w_decin, 0, (size_t)lo.B * lo.M * 2, st));   // go frame (helpers.py:31)
  return T2_OK;
}

// decoder steps [t_begin, t_end) of the free-running loop; t_end <= cfg->T_out (= max_iters). After the call
// workspace "stop_logits_tm" [T_out][B] (col M of the projection rows) holds the stop logits of the finished steps:
// the host applies the helper's rule (all rows round(sigmoid) == 1, helpers.py:40-54) and decides whether to continue.
extern "C" int t2_taco_infer_steps(const t2_taco_config_t* cfg, float* d_params, const void* d_packed, void* d_workspace,
                                   const int* d_input_lengths, int t_begin, int t_end, unsigned long long seed, void* stream) {
  TL lo;
  int rc = build(cfg, lo, nullptr);
  if (rc) return rc;
  T2_REQUIRE(t_begin >= 0 && t_begin <= t_end && t_end <= lo.To, T2_ERR_INVALID_ARG, "bad step range [%d, %d)", t_begin, t_end);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  const uint8_t* pk = static_cast<const uint8_t*>(d_packed);
  StepCtx s{&lo, ws, pk, d_params, st, seed, nullptr, 0};
  const int B = lo.B, D = lo.D, H = lo.H, PIK = D + 2 * H;
  DecBufs db;
  db.K1r = 2 * H + D; db.K2 = 2 * D; db.PIK = PIK;
  db.S1 = reinterpret_cast<bf16*>(ws + lo.w_S1); db.S2 = reinterpret_cast<bf16*>(ws + lo.w_S2); db.PI = reinterpret_cast<bf16*>(ws + lo.w_PI);
  db.c1 = reinterpret_cast<float*>(ws + lo.w_c1); db.c2 = reinterpret_cast<float*>(ws + lo.w_c2);
  db.cum = reinterpret_cast<float*>(ws + lo.w_cum); db.attU = reinterpret_cast<float*>(ws + lo.w_attU);
  db.keys = reinterpret_cast<float*>(ws + lo.w_keys); db.values = reinterpret_cast<bf16*>(ws + lo.w_values);
  db.pre1 = reinterpret_cast<float*>(ws + lo.w_pre1);
  db.att_smem = att_fwd_smem(lo.Ti, lo.KA, lo.A, D, 2 * H);
  bf16* decin = reinterpret_cast<bf16*>(ws + lo.w_decin);
  bf16* pn1 = reinterpret_cast<bf16*>(ws + lo.w_pn1);
  bf16* pn2 = reinterpret_cast<bf16*>(ws + lo.w_pn2);
  float* projo = reinterpret_cast<float*>(ws + lo.w_projo);
  for (int t = t_begin; t < t_end; ++t) {
    const unsigned long long seed_t = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(t + 1);   // fresh prenet masks per step
    bf16* x0 = decin + (long long)t * B * lo.M;
    bf16* x1 = pn1 + (long long)t * B * lo.P1;
    bf16* x2 = pn2 + (long long)t * B * lo.P2;
    rc = conv_gemm(x0, lo.M, B, 1, pk + lo.k_p1, lo.P1, 128, 1, nullptr, lo.P1 >= 256 ? 256 : 128, d_params + lo.p_p1b, 1, x1, nullptr, lo.P1,
                   lo.P1, lo.c.dropout_rate, 20, seed_t, nullptr, st);
    if (rc) return rc;
    rc = conv_gemm(x1, lo.P1, B, 1, pk + lo.k_p2, lo.P2, lo.P1, 1, nullptr, lo.P2 >= 256 ? 256 : 128, d_params + lo.p_p2b, 1, x2, nullptr, lo.P2,
                   lo.P2, lo.c.dropout_rate, 21, seed_t, nullptr, st);
    if (rc) return rc;
    rc = conv_gemm(x2, lo.P2, B, 1, pk + lo.k_l1x, 4 * D, lo.P2, 1, nullptr, 256, d_params + lo.p_l1b, 0, nullptr,
                   db.pre1 + (long long)t * B * 4 * D, 4 * D, 4 * D, 0.f, 0, 0, nullptr, st);
    if (rc) return rc;
    rc = decoder_step(s, db, d_input_lengths, t);
    if (rc) return rc;
    float* pt = projo + (long long)t * B * 128;
    rc = conv_gemm(db.PI + (long long)t * B * PIK, PIK, B, 1, pk + lo.k_proj, lo.M + 1, PIK, 1, nullptr, 128, nullptr, 0, nullptr, pt, 128, lo.M + 1,
                   0.f, 0, 0, nullptr, st);
    if (rc) return rc;
    T2_CHECK_CUDA(launch_pdl(proj_bias_feedback_kernel, dim3(g1((long long)B * (lo.M + 1))), dim3(256), 0, st, pt, d_params + lo.p_fb,
                             d_params + lo.p_sb, t + 1 < lo.To ? decin + (long long)(t + 1) * B * lo.M : (bf16*)nullptr, B, lo.M));
    t2_count_launch();
  }
  T2_CHECK_CUDA(cudaGetLastError());
  return T2_OK;
}

// clip, postnet (inference batch-norm) and residual over the T_used decoded frames. Results (compact, batch-major):
// workspace "decoder_output" / "mel_outputs" [B][T_used][M], "stop_logits" [B][T_used]; alignments stay [T_out][B][T_in].
extern "C" int t2_taco_infer_finish(const t2_taco_config_t* cfg, float* d_params, const void* d_packed, void* d_workspace, int T_used, void* stream) {
  TL lo;
  int rc = build(cfg, lo, nullptr);
  if (rc) return rc;
  T2_REQUIRE(T_used >= 1 && T_used <= lo.To, T2_ERR_INVALID_ARG, "T_used %d outside [1, %d]", T_used, lo.To);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  const uint8_t* pk = static_cast<const uint8_t*>(d_packed);
  StepCtx s{&lo, ws, pk, d_params, st, 0ull, nullptr, 0};
  const int B = lo.B;
  float* scal = reinterpret_cast<float*>(ws + lo.w_scal);
  const float lo_c = -lo.c.max_abs_value - lo.c.lower_bound_decay, hi_c = lo.c.max_abs_value;
  bf16* dec_bm = reinterpret_cast<bf16*>(ws + lo.w_decbm);
  float* dec_f = reinterpret_cast<float*>(ws + lo.w_decf);
  dec_finish_kernel<<<g1((long long)B * T_used * (lo.M + 1)), 256, 0, st>>>(reinterpret_cast<float*>(ws + lo.w_projo), nullptr, nullptr, dec_bm, dec_f,
                                                                            reinterpret_cast<float*>(ws + lo.w_stop), scal, B, T_used, lo.M,
                                                                            lo.c.clip_outputs, lo_c, hi_c, lo.c.split_bf16, nullptr, 1.f); t2_count_launch();
  const void* x = dec_bm;
  for (auto& L : lo.post) { rc = conv_block_fwd(s, L, x, T_used, 0); if (rc) return rc; x = ws + L.w_x; }
  float* resid = reinterpret_cast<float*>(ws + lo.w_resid);
  rc = conv_gemm(x, lo.PC, T_used, B, pk + lo.k_pp, lo.M, lo.PC * (lo.c.split_bf16 ? 3 : 1), 1, nullptr, 128, d_params + lo.p_ppb, 0, nullptr, resid, 128, lo.M,
                 0.f, 0, 0, nullptr, st, lo.c.split_bf16);
  if (rc) return rc;
  mel_finish_kernel<<<g1((long long)B * T_used * lo.M), 256, 0, st>>>(dec_f, resid, nullptr, reinterpret_cast<float*>(ws + lo.w_mel), scal,
                                                                      (long long)B * T_used, lo.M, lo.c.clip_outputs, lo_c, hi_c, nullptr, T_used); t2_count_launch();
  T2_CHECK_CUDA(cudaGetLastError());
  return T2_OK;
}

// tools only: device buffer of 32 int64 that receives clock64() phase stamps of the attention kernels (NULL = off)
extern "C" int t2_dbg_att_stamps(long long* d_buf) {
  T2_CHECK_CUDA(cudaMemcpyToSymbol(g_att_dbg, &d_buf, sizeof(d_buf)));
  return T2_OK;
}

extern "C" int t2_taco_workspace_tensor(const t2_taco_config_t* cfg, void* d_workspace, const char* name, void** ptr, long long* count,
                                        int* elem_bytes) {
  TL lo;
  int rc = build(cfg, lo, nullptr);
  if (rc) return rc;
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  const long long B = lo.B, Ti = lo.Ti, To = lo.To;
  struct E { const char* n; long long off, cnt; int eb; };
  const E table[] = {
      {"memory", lo.w_memory, B * Ti * 2 * lo.H, 2}, {"keys", lo.w_keys, B * Ti * lo.A, 4}, {"alignments", lo.w_alpha, To * B * Ti, 4},
      {"decoder_output", lo.w_decf, B * To * lo.M, 4}, {"mel_outputs", lo.w_mel, B * To * lo.M, 4}, {"stop_logits", lo.w_stop, B * To, 4},
      {"projection_rows", lo.w_projo, To * B * 128, 4}, {"enc_conv_out", lo.enc.back().w_x, B * Ti * lo.C, 2}, {"prenet", lo.w_pn2, To * B * lo.P2, 2}, {"proj_in", lo.w_PI, To * B * (lo.D + 2 * lo.H), 2},
  };
  for (const E& e : table)
    if (strcmp(e.n, name) == 0) { *ptr = ws + e.off; *count = e.cnt; *elem_bytes = e.eb; return T2_OK; }
  // per-layer conv-block tensors: "enc_conv_y<i>" / "enc_conv_x<i>" / "post_conv_y<i>" / "post_conv_x<i>" (y = conv + activation,
  // x = batch norm + dropout; bf16 [B][T][C]) and the postnet projection "postnet_residual" (fp32 [B][T_out][128])
  if (strcmp(name, "postnet_residual") == 0) { *ptr = ws + lo.w_resid; *count = B * To * 128; *elem_bytes = 4; return T2_OK; }
  for (int which = 0; which < 2; ++which) {
    const std::vector<ConvL>& v = which == 0 ? lo.enc : lo.post;
    const char* pre = which == 0 ? "enc_conv_" : "post_conv_";
    const size_t pl = strlen(pre);
    if (strncmp(name, pre, pl) == 0 && (name[pl] == 'x' || name[pl] == 'y') && name[pl + 1] >= '0' && name[pl + 1] <= '9') {
      const int i = name[pl + 1] - '0';
      if (i < int(v.size())) {
        *ptr = ws + (name[pl] == 'x' ? v[i].w_x : v[i].w_y);
        *count = B * (which == 0 ? Ti : To) * v[i].cout; *elem_bytes = 2;
        return T2_OK;
      }
    }
  }
  return t2_set_error(T2_ERR_INVALID_ARG, "unknown workspace tensor '%s'", name);
}

extern "C" int t2_taco_set_target_lengths(const t2_taco_config_t* cfg, void* d_workspace, const int* d_target_lengths, void* stream) {
  TL lo;
  int rc = build(cfg, lo, nullptr);
  if (rc) return rc;
  T2_REQUIRE(d_target_lengths != nullptr, T2_ERR_INVALID_ARG, "null target lengths");
  T2_CHECK_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(d_workspace) + lo.w_tlen, d_target_lengths, lo.B * sizeof(int), cudaMemcpyDeviceToDevice,
                                static_cast<cudaStream_t>(stream)));
  return T2_OK;
}

// backward of the last t2_taco_forward(training=1): writes d(total loss)/d(theta) for every trainable tensor
extern "C" int t2_taco_backward(const t2_taco_config_t* cfg, const float* d_params, const void* d_packed, void* d_workspace,
                                const int* d_inputs, const int* d_input_lengths, const float* d_mel_targets, const float* d_stop_targets,
                                float* d_grads, unsigned long long seed, const unsigned long long* d_step, void* stream) {
  return t2_taco_backward_ex(cfg, d_params, d_packed, d_workspace, d_inputs, d_input_lengths, d_mel_targets, d_stop_targets, d_grads, nullptr, seed,
                             d_step, stream);
}

extern "C" int t2_taco_backward_ex(const t2_taco_config_t* cfg, const float* d_params, const void* d_packed, void* d_workspace,
                                   const int* d_inputs, const int* d_input_lengths, const float* d_mel_targets, const float* d_stop_targets,
                                   float* d_grads, const float* d_mel_outputs_grad, unsigned long long seed, const unsigned long long* d_step,
                                   void* stream) {
  TL lo;
  int rc = build(cfg, lo, nullptr);
  if (rc) return rc;
  T2_REQUIRE(!lo.c.split_bf16, T2_ERR_INVALID_ARG, "split_bf16 (fp32-class conv stacks) mode has no backward pass");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  const uint8_t* pk = static_cast<const uint8_t*>(d_packed);
  StepCtx s{&lo, ws, pk, d_params, st, seed, d_step, 1};
  const int B = lo.B, Ti = lo.Ti, To = lo.To, H = lo.H, D = lo.D, M = lo.M, A = lo.A;
  const int K1r = 2 * H + D, K2 = 2 * D, PIK = D + 2 * H;
  const long long TB = (long long)To * B, BTo = (long long)B * To;
  std::vector<WgL> wl;
  build_tiles(lo, wl);
  std::vector<int> toff(wl.size());
  { int o = 0; for (size_t i = 0; i < wl.size(); ++i) { toff[i] = o; o += int(wl[i].tiles.size()); } }
  const WgradTile* tiles = reinterpret_cast<const WgradTile*>(ws + lo.w_tiles);
  auto TILES = [&](int i) { return tiles + toff[i]; };
  auto NT = [&](int i) { return int(wl[i].tiles.size()); };
  int li = 0;  // running wgrad-launch index (must follow build_tiles order)
  T2_CHECK_CUDA(cudaMemsetAsync(d_grads, 0, lo.n_params * sizeof(float), st));
  const float lo_c = -lo.c.max_abs_value - lo.c.lower_bound_decay, hi_c = lo.c.max_abs_value;
  bf16* dY0 = reinterpret_cast<bf16*>(ws + lo.w_dY);
  bf16* dY1 = dY0 + (long long)B * (To > Ti ? To : Ti) * (lo.PC > lo.C ? lo.PC : lo.C);
  bf16* dmel = reinterpret_cast<bf16*>(ws + lo.w_dmel);
  float* ddecf = reinterpret_cast<float*>(ws + lo.w_ddecf);
  bf16* dec_bm = reinterpret_cast<bf16*>(ws + lo.w_decbm);
  const int* tlen = lo.c.mask_decoder ? reinterpret_cast<const int*>(ws + lo.w_tlen) : nullptr;
  const float* scal = reinterpret_cast<const float*>(ws + lo.w_scal);      // [5], [6]: the loss normalisers of the forward pass
  // ---- loss seeds + postnet ----
  loss_seed_kernel<<<g1(BTo * 128), 256, 0, st>>>(reinterpret_cast<float*>(ws + lo.w_decf), reinterpret_cast<float*>(ws + lo.w_resid),
                                                  reinterpret_cast<float*>(ws + lo.w_mel), d_mel_targets, dmel, ddecf, BTo, M, lo.c.clip_outputs,
                                                  lo_c, hi_c, tlen, To, scal, d_mel_outputs_grad); t2_count_launch();
  rc = conv_gemm(dmel, 128, To, B, pk + lo.k_ppT, lo.PC, 128, 1, nullptr, lo.PC % 256 == 0 ? 256 : 128, nullptr, 0, dY0, nullptr, lo.PC, lo.PC, 0.f, 0, 0,
                 nullptr, st);
  if (rc) return rc;
  {
    ActT maps[2] = {make_act(ws + lo.post.back().w_x, lo.PC, To, B), make_act(dmel, 128, To, B)};
    rc = launch_wgrad(maps, 2, TILES(li), NT(li), d_grads, To, B, st); if (rc) return rc; ++li;
    colsum_bf16_kernel<<<64, 256, 0, st>>>(dmel, BTo, M, 128, d_grads + lo.p_ppb, 1.f); t2_count_launch();
  }
  bf16* ddec_post = reinterpret_cast<bf16*>(ws + lo.w_dz);
  for (int i = int(lo.post.size()) - 1; i >= 0; --i) {
    const void* xin = i > 0 ? (const void*)(ws + lo.post[i - 1].w_x) : (const void*)dec_bm;
    rc = conv_block_bwd(s, lo.post[i], xin, To, dY0, dY1, i > 0 ? dY0 : ddec_post, d_grads, TILES(li), NT(li));
    if (rc) return rc;
    ++li;
  }
  // ---- projections ----
  bf16* ddec_tm = reinterpret_cast<bf16*>(ws + lo.w_ddec_tm);
  float* projo = reinterpret_cast<float*>(ws + lo.w_projo);
  ddec_tm_kernel<<<g1(TB * 128), 256, 0, st>>>(ddecf, ddec_post, projo, d_stop_targets, ddec_tm, B, To, M, lo.c.clip_outputs, lo_c, hi_c, tlen,
                                               lo.c.cross_entropy_pos_weight, scal);
  t2_count_launch();
  float* dPI = reinterpret_cast<float*>(ws + lo.w_dPI);
  rc = conv_gemm(ddec_tm, 128, TB, 1, pk + lo.k_projT, PIK, 128, 1, nullptr, PIK % 256 == 0 ? 256 : 128, nullptr, 0, nullptr, dPI, PIK, PIK, 0.f, 0, 0,
                 nullptr, st);
  if (rc) return rc;
  bf16* PI = reinterpret_cast<bf16*>(ws + lo.w_PI);
  {
    ActT maps[2] = {make_act(PI, PIK, int(TB), 1), make_act(ddec_tm, 128, int(TB), 1)};
    rc = launch_wgrad(maps, 2, TILES(li), NT(li), d_grads, int(TB), 1, st); if (rc) return rc; ++li;
    colsum_bf16_kernel<<<64, 256, 0, st>>>(ddec_tm, TB, M, 128, d_grads + lo.p_fb, 1.f); t2_count_launch();
    colsum_bf16_kernel<<<64, 256, 0, st>>>(ddec_tm + M, TB, 1, 128, d_grads + lo.p_sb, 1.f); t2_count_launch();
  }
  // ---- decoder: backward through time ----
  float* dh1ext = reinterpret_cast<float*>(ws + lo.w_dh1ext);
  float* dh2ext = reinterpret_cast<float*>(ws + lo.w_dh2ext);
  float* dhs1 = reinterpret_cast<float*>(ws + lo.w_dhs1);
  float* dhs2 = reinterpret_cast<float*>(ws + lo.w_dhs2);
  float* dcs1 = reinterpret_cast<float*>(ws + lo.w_dcs1);
  float* dcs2 = reinterpret_cast<float*>(ws + lo.w_dcs2);
  float* dctxl = reinterpret_cast<float*>(ws + lo.w_dctxl);
  float* dcum = reinterpret_cast<float*>(ws + lo.w_dcum);
  float* cumrun = reinterpret_cast<float*>(ws + lo.w_cumrun);
  float* dkeys = reinterpret_cast<float*>(ws + lo.w_dkeys);
  float* attacc = reinterpret_cast<float*>(ws + lo.w_attacc);
  const int nacc = (lo.KA + 2) * A;
  for (float* p : {dhs1, dhs2, dcs1, dcs2}) T2_CHECK_CUDA(cudaMemsetAsync(p, 0, (size_t)B * D * 4, st));
  T2_CHECK_CUDA(cudaMemsetAsync(dctxl, 0, (size_t)B * 2 * H * 4, st));
  T2_CHECK_CUDA(cudaMemsetAsync(dh1ext, 0, (size_t)B * D * 4, st));
  const int ks_dec = (4 * D / kBK) >= 16 ? 8 : 1, ks_enc = (4 * H / kBK) >= 16 ? 8 : 1;   // k-blocks per slice >= 2
  T2_CHECK_CUDA(cudaMemsetAsync(dcum, 0, (size_t)B * Ti * 4, st));
  T2_CHECK_CUDA(cudaMemsetAsync(dkeys, 0, (size_t)B * Ti * A * 4, st));
  T2_CHECK_CUDA(cudaMemsetAsync(attacc, 0, (size_t)B * nacc * 4, st));
  T2_CHECK_CUDA(cudaMemcpyAsync(cumrun, ws + lo.w_cum, (size_t)B * Ti * 4, cudaMemcpyDeviceToDevice, st));
  bf16* dg1 = reinterpret_cast<bf16*>(ws + lo.w_dg1);
  bf16* dg2 = reinterpret_cast<bf16*>(ws + lo.w_dg2);
  bf16* dctx_all = reinterpret_cast<bf16*>(ws + lo.w_dctx_all);
  bf16* dq_all = reinterpret_cast<bf16*>(ws + lo.w_dq_all);
  const size_t ab_smem = att_bwd_smem(Ti, lo.KA, A, D, 2 * H);
  const float* attU = reinterpret_cast<const float*>(ws + lo.w_attU);
  T2_CHECK_CUDA(cudaFuncSetAttribute(att_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(ab_smem)));
  for (int t = To - 1; t >= 0; --t) {
    AttBwd a;
    a.h2out = PI + (long long)t * B * PIK; a.ld_h2 = PIK; a.WqT = reinterpret_cast<const bf16*>(pk + lo.k_qT); a.Wq = d_params + lo.p_qry;
    a.U = attU; a.v = d_params + lo.p_v;
    a.keys = reinterpret_cast<const float*>(ws + lo.w_keys); a.values = reinterpret_cast<const bf16*>(ws + lo.w_values); a.lens = d_input_lengths;
    a.alpha = reinterpret_cast<const float*>(ws + lo.w_alpha) + (long long)t * B * Ti; a.cumrun = cumrun; a.dcum = dcum;
    a.dPI = dPI + (long long)t * B * PIK; a.ld_dPI = PIK; a.dctxl = dctxl; a.dh2ext = dh2ext;
    a.dctx_save = dctx_all + (long long)t * B * 2 * H; a.dq_save = dq_all + (long long)t * B * A; a.dkeys = dkeys; a.acc = attacc;
    a.B = B; a.Ti = Ti; a.D = D; a.A = A; a.KA = lo.KA; a.C2 = 2 * H;
    T2_CHECK_CUDA(launch_pdl(att_bwd_kernel, dim3(B), dim3(kAttThreads), ab_smem, st, a)); t2_count_launch();
    CellBwd c2;
    c2.dh_ext = dh2ext; c2.zero_ext = 0; c2.ld_ext = D; c2.dhs = dhs2; c2.dcs = dcs2;
    c2.gst = reinterpret_cast<const bf16*>(ws + lo.w_g2) + (long long)t * B * 4 * D; c2.tst = reinterpret_cast<const bf16*>(ws + lo.w_t2) + (long long)t * B * D;
    c2.c_prev = reinterpret_cast<const float*>(ws + lo.w_c2) + (long long)t * B * D;
    c2.dg_a = dg2 + (long long)t * B * 4 * D; c2.ld_a = 4 * D; c2.dg_b = nullptr; c2.ld_b = 0; c2.lens = nullptr; c2.t = t; c2.B = B; c2.H = D; c2.stream = 55;
    c2.zone = lo.c.zoneout_rate; c2.seed = seed; c2.step = d_step;
    T2_CHECK_CUDA(launch_pdl(lstm_cell_bwd_kernel, dim3(g1((long long)B * D)), dim3(256), 0, st, c2)); t2_count_launch();
    rc = lstm_bwd_gemm(s, pk + lo.k_l2T, K2, 4 * D, c2.dg_a, B, dh1ext, D, D, 2, dhs2, D, 2, ks_dec);   // dh1ext: zeroed by its consumer
    if (rc) return rc;
    CellBwd c1 = c2;
    c1.dh_ext = dh1ext; c1.zero_ext = 1; c1.dhs = dhs1; c1.dcs = dcs1;
    c1.gst = reinterpret_cast<const bf16*>(ws + lo.w_g1) + (long long)t * B * 4 * D; c1.tst = reinterpret_cast<const bf16*>(ws + lo.w_t1) + (long long)t * B * D;
    c1.c_prev = reinterpret_cast<const float*>(ws + lo.w_c1) + (long long)t * B * D; c1.dg_a = dg1 + (long long)t * B * 4 * D; c1.stream = 54;
    T2_CHECK_CUDA(launch_pdl(lstm_cell_bwd_kernel, dim3(g1((long long)B * D)), dim3(256), 0, st, c1)); t2_count_launch();
    rc = lstm_bwd_gemm(s, pk + lo.k_l1rT, K1r, 4 * D, c1.dg_a, B, dctxl, 2 * H, 2 * H, 2, dhs1, D, 2, ks_dec);  // dctxl: zeroed by att_bwd
    if (rc) return rc;
  }
  T2_CHECK_CUDA(cudaGetLastError());
  // ---- recurrent / prenet weight gradients: one wgrad GEMM over all steps ----
  bf16* pn1 = reinterpret_cast<bf16*>(ws + lo.w_pn1);
  bf16* pn2 = reinterpret_cast<bf16*>(ws + lo.w_pn2);
  {
    ActT maps[5] = {make_act(ws + lo.w_S2, K2, int(TB), 1), make_act(dg2, 4 * D, int(TB), 1), make_act(ws + lo.w_S1, K1r, int(TB), 1),
                    make_act(dg1, 4 * D, int(TB), 1), make_act(pn2, lo.P2, int(TB), 1)};
    rc = launch_wgrad(maps, 5, TILES(li), NT(li), d_grads, int(TB), 1, st); if (rc) return rc; ++li;
    colsum_bf16_kernel<<<64, 256, 0, st>>>(dg2, TB, 4 * D, 4 * D, d_grads + lo.p_l2b, 1.f); t2_count_launch();
    colsum_bf16_kernel<<<64, 256, 0, st>>>(dg1, TB, 4 * D, 4 * D, d_grads + lo.p_l1b, 1.f); t2_count_launch();
  }
  bf16* dpn2 = reinterpret_cast<bf16*>(ws + lo.w_dpn2);
  bf16* dpn1 = reinterpret_cast<bf16*>(ws + lo.w_dpn1);
  bf16* dz2 = reinterpret_cast<bf16*>(ws + lo.w_dz);
  rc = conv_gemm(dg1, 4 * D, TB, 1, pk + lo.k_l1xT, lo.P2, 4 * D, 1, nullptr, lo.P2 % 256 == 0 ? 256 : 128, nullptr, 0, dpn2, nullptr, lo.P2, lo.P2, 0.f, 0,
                 0, nullptr, st);
  if (rc) return rc;
  relu_drop_bwd_kernel<<<g1(TB * lo.P2), 256, 0, st>>>(dpn2, pn2, dz2, TB * lo.P2, lo.c.dropout_rate); t2_count_launch();
  rc = conv_gemm(dz2, lo.P2, TB, 1, pk + lo.k_p2T, lo.P1, lo.P2, 1, nullptr, lo.P1 % 256 == 0 ? 256 : 128, nullptr, 0, dpn1, nullptr, lo.P1, lo.P1, 0.f, 0, 0,
                 nullptr, st);
  if (rc) return rc;
  relu_drop_bwd_kernel<<<g1(TB * lo.P1), 256, 0, st>>>(dpn1, pn1, dpn1, TB * lo.P1, lo.c.dropout_rate); t2_count_launch();
  {
    ActT maps[6] = {make_act(pn1, lo.P1, int(TB), 1), make_act(dz2, lo.P2, int(TB), 1), make_act(ws + lo.w_decin, M, int(TB), 1),
                    make_act(dpn1, lo.P1, int(TB), 1), make_act(PI, PIK, int(TB), 1), make_act(dq_all, A, int(TB), 1)};
    rc = launch_wgrad(maps, 6, TILES(li), NT(li), d_grads, int(TB), 1, st); if (rc) return rc; ++li;
    colsum_bf16_kernel<<<64, 256, 0, st>>>(dz2, TB, lo.P2, lo.P2, d_grads + lo.p_p2b, 1.f); t2_count_launch();
    colsum_bf16_kernel<<<64, 256, 0, st>>>(dpn1, TB, lo.P1, lo.P1, d_grads + lo.p_p1b, 1.f); t2_count_launch();
  }
  {
    float* scratch = reinterpret_cast<float*>(ws + lo.w_attU) + (lo.KA + 1) * A;   // [(KA+2)][A] reduced accumulators
    att_finish_kernel<<<g1(nacc), 256, 0, st>>>(attacc, d_params + lo.p_lck, d_params + lo.p_lcb, d_params + lo.p_lfl, d_grads, B, lo.KA, lo.F, A,
                                                lo.p_lck, lo.p_lcb, lo.p_lfl, lo.p_v, lo.p_ba, scratch); t2_count_launch();
    att_finish2_kernel<<<g1(lo.KA * lo.F + lo.F * A + lo.F + A), 256, 0, st>>>(scratch, d_params + lo.p_lck, d_params + lo.p_lcb, d_params + lo.p_lfl,
                                                                             d_grads, lo.KA, lo.F, A, lo.p_lck, lo.p_lcb, lo.p_lfl, lo.p_v,
                                                                             lo.p_ba); t2_count_launch();
  }
  // ---- attention memory: keys / values ----
  bf16* dkeysb = reinterpret_cast<bf16*>(ws + lo.w_dkeysb);
  f32_to_bf16_k<<<g1((long long)B * Ti * A), 256, 0, st>>>(dkeys, dkeysb, (long long)B * Ti * A); t2_count_launch();
  float* dvalues = reinterpret_cast<float*>(ws + lo.w_dvalues);
  rc = conv_gemm(dkeysb, A, Ti, B, pk + lo.k_memT, 2 * H, A, 1, nullptr, (2 * H) % 256 == 0 ? 256 : 128, nullptr, 0, nullptr, dvalues, 2 * H, 2 * H, 0.f, 0, 0,
                 nullptr, st);
  if (rc) return rc;
  {
    ActT maps[2] = {make_act(ws + lo.w_values, 2 * H, Ti, B), make_act(dkeysb, A, Ti, B)};
    rc = launch_wgrad(maps, 2, TILES(li), NT(li), d_grads, Ti, B, st); if (rc) return rc; ++li;
  }
  dvalues_ctx_kernel<<<dim3(Ti, B), 256, 0, st>>>(reinterpret_cast<float*>(ws + lo.w_alpha), dctx_all, d_input_lengths, dvalues, B, Ti, To, 2 * H);
  t2_count_launch();
  // ---- encoder BiLSTM, backward through time ----
  const void* x3 = ws + lo.enc.back().w_x;
  TacoSide* side = taco_side();
  if (side) {
    T2_CHECK_CUDA(cudaEventRecord(side->fork, st));
    T2_CHECK_CUDA(cudaStreamWaitEvent(side->s, side->fork, 0));
  }
  for (int d = 0; d < 2; ++d) {
    cudaStream_t sx = (d == 1 && side) ? side->s : st;
    StepCtx sc = s; sc.st = sx;
    float* edh = reinterpret_cast<float*>(ws + lo.w_encdh[d]);
    float* edc = reinterpret_cast<float*>(ws + lo.w_encdc[d]);
    T2_CHECK_CUDA(cudaMemsetAsync(edh, 0, (size_t)B * H * 4, sx));
    T2_CHECK_CUDA(cudaMemsetAsync(edc, 0, (size_t)B * H * 4, sx));
    bf16* dgall = reinterpret_cast<bf16*>(ws + lo.w_encdgall[d]);
    bf16* dpre = reinterpret_cast<bf16*>(ws + lo.w_dencpre[d]);
    for (int sidx = Ti - 1; sidx >= 0; --sidx) {
      const int t = d == 0 ? sidx : Ti - 1 - sidx;
      CellBwd c;
      c.zero_ext = 0; c.dh_ext = dvalues + (long long)t * 2 * H + d * H; c.ld_ext = (long long)Ti * 2 * H; c.dhs = edh; c.dcs = edc;
      c.gst = reinterpret_cast<const bf16*>(ws + lo.w_encg[d]) + (long long)sidx * B * 4 * H;
      c.tst = reinterpret_cast<const bf16*>(ws + lo.w_enct[d]) + (long long)sidx * B * H;
      c.c_prev = reinterpret_cast<const float*>(ws + lo.w_encc[d]) + (long long)sidx * B * H;
      c.dg_a = dgall + (long long)sidx * B * 4 * H; c.ld_a = 4 * H; c.dg_b = dpre + (long long)t * 4 * H; c.ld_b = (long long)Ti * 4 * H;
      c.lens = d_input_lengths; c.t = t; c.B = B; c.H = H; c.stream = 52 + d; c.zone = lo.c.zoneout_rate; c.seed = seed; c.step = d_step;
      T2_CHECK_CUDA(launch_pdl(lstm_cell_bwd_kernel, dim3(g1((long long)B * H)), dim3(256), 0, sx, c)); t2_count_launch();
      rc = lstm_bwd_gemm(sc, pk + lo.k_encWrT[d], H, 4 * H, c.dg_a, B, edh, H, H, 2, nullptr, 0, 0, ks_enc);
      if (rc) return rc;
    }
    {
      ActT maps[2] = {make_act(ws + lo.w_ench[d], H, Ti * B, 1), make_act(dgall, 4 * H, Ti * B, 1)};
      rc = launch_wgrad(maps, 2, TILES(li), NT(li), d_grads, Ti * B, 1, sx); if (rc) return rc; ++li;
      ActT maps2[2] = {make_act(x3, lo.C, Ti, B), make_act(dpre, 4 * H, Ti, B)};
      rc = launch_wgrad(maps2, 2, TILES(li), NT(li), d_grads, Ti, B, sx); if (rc) return rc; ++li;
      colsum_bf16_kernel<<<64, 256, 0, sx>>>(dpre, (long long)B * Ti, 4 * H, 4 * H, d_grads + lo.p_elb[d], 1.f); t2_count_launch();
    }
  }
  if (side) {
    T2_CHECK_CUDA(cudaEventRecord(side->join, side->s));
    T2_CHECK_CUDA(cudaStreamWaitEvent(st, side->join, 0));
  }
  // dx3 = dpre_fw x Wx_fw^T + dpre_bw x Wx_bw^T
  bf16* dx3 = reinterpret_cast<bf16*>(ws + lo.w_dx3);
  {
    ActGemmCall g;
    memset(&g, 0, sizeof(g));
    g.a[0] = make_act(ws + lo.w_dencpre[0], 4 * H, Ti, B); g.a[1] = make_act(ws + lo.w_dencpre[1], 4 * H, Ti, B); g.na = 2;
    g.seg[0] = Seg{0, 0, 0, 4 * H / kBK, 0, 1}; g.seg[1] = Seg{1, 0, 0, 4 * H / kBK, 0, 1}; g.nseg = 2;
    g.w = pk + lo.k_encWxT; g.wN = lo.C; g.wK = 8 * H; g.wL = 1;
    g.T = Ti; g.B = B; g.n_tiles = lo.C / (lo.C % 256 == 0 ? 256 : 128);
    g.epi.ptr[0] = dx3; g.epi.i[0] = lo.C; g.epi.i[1] = 0; g.epi.i[2] = lo.C;
    rc = launch_act_gemm(EPI_BIAS_ACT, lo.C % 256 == 0 ? 256 : 128, g, st);
    if (rc) return rc;
  }
  // ---- encoder conv blocks + embedding ----
  {
    const bf16* dout = dx3;   // a block's output gradient is dead once its BN backward has run, so dx may overwrite it
    bf16* demb = reinterpret_cast<bf16*>(ws + lo.w_demb);
    for (int i = int(lo.enc.size()) - 1; i >= 0; --i) {
      const void* xin = i > 0 ? (const void*)(ws + lo.enc[i - 1].w_x) : (const void*)(ws + lo.w_emb);
      bf16* dx = i > 0 ? dY0 : demb;
      rc = conv_block_bwd(s, lo.enc[i], xin, Ti, dout, dY1, dx, d_grads, TILES(li), NT(li));
      if (rc) return rc;
      ++li;
      dout = dx;
    }
    embed_bwd_kernel<<<g1((long long)B * Ti * lo.E), 256, 0, st>>>(d_inputs, demb, d_grads + lo.p_emb, (long long)B * Ti, lo.E); t2_count_launch();
  }
  // ---- L2 regulariser (tacotron.py:343-345) ----
  reg_grad_kernel<<<dim3(8, lo.n_reg), 256, 0, st>>>(d_params, d_grads, reinterpret_cast<const long long*>(ws + lo.w_regtab), lo.c.reg_weight);
  t2_count_launch();
  T2_CHECK_CUDA(cudaGetLastError());
  return T2_OK;
}
