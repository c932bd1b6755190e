// t2_cbhg.cu — the CBHG post-processing network + linear-spectrogram head of the Tacotron graph (predict_linear = True,
// the reference's DEFAULT: hparams.py:175).
//
// Replaces tacotron/models/tacotron.py:203-219 (CBHG_postnet -> cbhg_linear_specs_projection -> clip), :323-330 (linear loss) and
// tacotron/models/modules.py:4-16 (HighwayNet), :19-78 (CBHG), :457-485 (MaskedLinearLoss). Forward and backward.
//
// Dataflow (rows = b * T + t, batch-major; N = B * T):
//   mel_outputs fp32 [N][M] -> bf16
//   conv bank: k = 1..K convolutions M -> CC ('same' padding, the extra pad of an even kernel on the right) + bias + ReLU, each one a
//     tap-shifted GEMM on the tcgen05 engine writing its 128-column slice of Y [N][K*CC]; every layer's batch norm runs on its slice
//   max-pool (2, stride 1, 'same': max(x[t], x[t+1]))
//   proj1 (k = 3, K*CC -> PJ, ReLU, BN), proj2 (k = 3, PJ -> M, linear, BN), + mel_outputs, dense M -> HU
//   NH highway layers: one GEMM with N = 2 HU ([H | T] pre-activations) + an elementwise kernel
//   bidirectional GRU (HU -> RU per direction, whole padded sequence): the input projections of all steps are ONE GEMM (N = 6 RU);
//     the recurrence runs in a persistent kernel - one CTA per (4 batch items, direction) keeps the recurrent weights (bf16) and the
//     state (fp32) in shared memory for all T steps
//   linear projection 2 RU -> num_freq (GEMM), clip, L1 loss with half of the weight on the bins below 2 kHz
// Backward mirrors it: elementwise / batch-norm backward kernels, data-gradient GEMMs with reversed taps, a persistent BPTT kernel
// for the GRU, and the weight gradients of every layer as tiles of the batched wgrad GEMM (reduction over positions).
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
constexpr int kMaxBank = 16, kMaxHw = 8;
constexpr int kGruItems = 4;          // batch items per CTA of the recurrent kernels
constexpr int kGruThreads = 256;

struct CPT { std::string name; long long off; int ndim; int shape[4]; bool trainable, reg; };
struct CConv {
  int cin, cout, k, act;               // act: 1 relu, 0 none
  long long p_k, p_b, p_g, p_be, p_mm, p_mv;
  int cinp, coutp;                     // channels rounded up to 64 (K slots of the packed operands)
  long long k_w, k_wT;                 // packed forward [cout][k * cinp], packed dgrad [cin rows][k * coutp]
};
struct PJ { long long src_off; int K, N; long long dst_off; int dst_ld, transpose, col0; };

struct CL {
  t2_cbhg_config_t c;
  int B, T, M, K, CC, KC, PJc, PK, NH, HU, RU, NF, NFP, NFR;
  long long N;
  std::vector<CPT> params;
  long long n_params;
  std::vector<CConv> bank;
  CConv proj1, proj2;
  bool has_dense;
  long long p_dk, p_db, p_hk[kMaxHw][2], p_hb[kMaxHw][2], p_gk[2], p_gb[2], p_ck[2], p_cb[2], p_lk, p_lb;
  // packed operands (bytes)
  long long k_dense, k_denseT, k_hw[kMaxHw], k_hwT[kMaxHw], k_gx, k_gxT, k_lin, k_linT, k_bankT[3], packed_bytes;
  int grp_first[4], grp_taps[3], n_grp;   // bank layers [grp_first[g], grp_first[g+1]) share one data-gradient GEMM (<= 16 taps)
  // workspace (bytes)
  long long w_x0, w_Y, w_Xb, w_P, w_stb, w_Y1, w_X1, w_st1, w_Y2, w_st2, w_hin, w_hf[kMaxHw + 1], w_hb[kMaxHw + 1], w_HT[kMaxHw];
  long long w_XP, w_out, w_gr[2], w_gu[2], w_gc[2], w_grh[2], w_lin, w_scal, w_tlen;
  long long w_dlin, w_dout, w_dXP, w_dh, w_dhb, w_dHT, w_dhin, w_dY2b, w_d1, w_d2, w_dP, w_dbank, w_dx0[3], w_bsum, w_tiles, w_jobs, w_regtab;
  long long workspace_bytes;
  int n_jobs, n_reg;
};

long long addp(CL& lo, const std::string& name, std::initializer_list<int> shape, bool trainable = true) {
  CPT p; p.name = name; p.off = lo.n_params; p.ndim = int(shape.size());
  long long n = 1; int i = 0;
  for (int s : shape) { p.shape[i++] = s; n *= s; }
  for (; i < 4; ++i) p.shape[i] = 1;
  p.trainable = trainable;
  // tacotron.py:343-345: no 'bias', 'Bias', '_projection', 'inputs_embedding', 'RNN', 'LSTM' in the variable name
  p.reg = trainable && name.find("bias") == std::string::npos && name.find("_projection") == std::string::npos &&
          name.find("RNN") == std::string::npos;
  lo.n_params += (n + 3) / 4 * 4;
  lo.params.push_back(p);
  return p.off;
}

int build(const t2_cbhg_config_t* cfg, CL& lo, std::vector<PJ>* jobs_out) {
  T2_REQUIRE(cfg != nullptr, T2_ERR_INVALID_ARG, "null CBHG config");
  lo.c = *cfg;
  lo.B = cfg->B; lo.T = cfg->T; lo.M = cfg->num_mels; lo.K = cfg->kernels; lo.CC = cfg->conv_channels; lo.KC = lo.K * lo.CC;
  lo.PJc = cfg->projection; lo.PK = cfg->projection_kernel_size; lo.NH = cfg->highwaynet_layers; lo.HU = cfg->highway_units;
  lo.RU = cfg->rnn_units; lo.NF = cfg->num_freq;
  lo.NFP = (lo.NF + 7) / 8 * 8;                 // row pitch of the bf16 gradient of the linear outputs
  lo.NFR = (lo.NF + 127) / 128 * 128;           // rows of the packed projection (whole 128-column output tiles)
  lo.N = (long long)lo.B * lo.T;
  T2_REQUIRE(lo.B >= 1 && lo.T >= 2, T2_ERR_UNSUPPORTED_SHAPE, "CBHG: B >= 1, T >= 2");
  T2_REQUIRE(lo.M % 8 == 0 && lo.M <= 128, T2_ERR_UNSUPPORTED_SHAPE, "CBHG: num_mels must be a multiple of 8, <= 128");
  T2_REQUIRE(lo.K >= 1 && lo.K <= 8 && lo.CC == 128, T2_ERR_UNSUPPORTED_SHAPE, "CBHG: 1..8 bank kernels of 128 channels");
  T2_REQUIRE(cfg->pool_size == 2, T2_ERR_UNSUPPORTED_SHAPE, "CBHG: pool_size must be 2");
  T2_REQUIRE(lo.PJc % 128 == 0 && lo.PJc <= 512 && lo.PK >= 1 && lo.PK <= 7 && lo.PK % 2 == 1, T2_ERR_UNSUPPORTED_SHAPE, "CBHG: projection sizes");
  T2_REQUIRE(lo.NH >= 1 && lo.NH <= kMaxHw && lo.HU == 128 && lo.RU == 128, T2_ERR_UNSUPPORTED_SHAPE,
             "CBHG: 1..8 highway layers, 128 highway / GRU units");
  T2_REQUIRE(lo.NF >= 8 && lo.NF <= 4096, T2_ERR_UNSUPPORTED_SHAPE, "CBHG: num_freq");
  lo.n_params = 0; lo.params.clear(); lo.bank.clear();
  const std::string P = "CBHG_postnet/";
  auto conv_params = [&](CConv& L, const std::string& pre) {
    L.p_k = addp(lo, pre + "kernel", {L.k, L.cin, L.cout}); L.p_b = addp(lo, pre + "bias", {L.cout});
    L.p_g = addp(lo, pre + "gamma", {L.cout}); L.p_be = addp(lo, pre + "beta", {L.cout});
    L.p_mm = addp(lo, pre + "moving_mean", {L.cout}, false); L.p_mv = addp(lo, pre + "moving_variance", {L.cout}, false);
    L.cinp = (L.cin + 63) / 64 * 64; L.coutp = (L.cout + 63) / 64 * 64;
  };
  for (int k = 1; k <= lo.K; ++k) {
    CConv L; L.cin = lo.M; L.cout = lo.CC; L.k = k; L.act = 1;
    char b[64]; snprintf(b, sizeof(b), "conv_bank/conv1d_%d/", k);
    conv_params(L, P + b); lo.bank.push_back(L);
  }
  lo.proj1.cin = lo.KC; lo.proj1.cout = lo.PJc; lo.proj1.k = lo.PK; lo.proj1.act = 1; conv_params(lo.proj1, P + "proj1/");
  lo.proj2.cin = lo.PJc; lo.proj2.cout = lo.M; lo.proj2.k = lo.PK; lo.proj2.act = 0; conv_params(lo.proj2, P + "proj2/");
  lo.has_dense = lo.M != lo.HU;
  T2_REQUIRE(lo.has_dense, T2_ERR_UNSUPPORTED_SHAPE, "CBHG: num_mels == highway_units (no dense layer) is not implemented");
  lo.p_dk = addp(lo, P + "dense/kernel", {lo.M, lo.HU}); lo.p_db = addp(lo, P + "dense/bias", {lo.HU});
  for (int i = 0; i < lo.NH; ++i) {
    char b[64];
    for (int j = 0; j < 2; ++j) {
      snprintf(b, sizeof(b), "highwaynet_%d/%s/", i + 1, j == 0 ? "H" : "T");
      lo.p_hk[i][j] = addp(lo, P + b + "kernel", {lo.HU, lo.HU}); lo.p_hb[i][j] = addp(lo, P + b + "bias", {lo.HU});
    }
  }
  const char* dn[2] = {"forward_RNN/", "backward_RNN/"};
  for (int d = 0; d < 2; ++d) {
    lo.p_gk[d] = addp(lo, P + dn[d] + "gates/kernel", {lo.HU + lo.RU, 2 * lo.RU}); lo.p_gb[d] = addp(lo, P + dn[d] + "gates/bias", {2 * lo.RU});
    lo.p_ck[d] = addp(lo, P + dn[d] + "candidate/kernel", {lo.HU + lo.RU, lo.RU}); lo.p_cb[d] = addp(lo, P + dn[d] + "candidate/bias", {lo.RU});
  }
  lo.p_lk = addp(lo, "cbhg_linear_specs_projection/kernel", {2 * lo.RU, lo.NF}); lo.p_lb = addp(lo, "cbhg_linear_specs_projection/bias", {lo.NF});

  // ---- packed operands ----
  std::vector<PJ> jobs;
  long long o = 0;
  auto takeb = [&](long long bytes) { long long r = o; o = al256(o + bytes); return r; };
  auto pj = [&](long long src, int K, int N, long long dst_bytes, int ld, int tr, int col0) {
    PJ j; j.src_off = src; j.K = K; j.N = N; j.dst_off = dst_bytes / 2; j.dst_ld = ld; j.transpose = tr; j.col0 = col0; jobs.push_back(j);
  };
  auto conv_pack = [&](CConv& L, bool with_t) {
    const int rows_f = (L.cout + 127) / 128 * 128, rows_t = (L.cin + 127) / 128 * 128;
    L.k_w = takeb(2LL * rows_f * L.k * L.cinp);
    L.k_wT = with_t ? takeb(2LL * rows_t * L.k * L.coutp) : 0;
    for (int j = 0; j < L.k; ++j) {
      pj(L.p_k + (long long)j * L.cin * L.cout, L.cin, L.cout, L.k_w, L.k * L.cinp, 1, j * L.cinp);      // fwd [cout][tap j | cin]
      if (with_t) pj(L.p_k + (long long)j * L.cin * L.cout, L.cin, L.cout, L.k_wT, L.k * L.coutp, 0, j * L.coutp);   // dgrad [cin][tap j | cout]
    }
  };
  for (auto& L : lo.bank) conv_pack(L, false);
  conv_pack(lo.proj1, true); conv_pack(lo.proj2, true);
  // data gradient of the bank: layers grouped greedily so that one GEMM holds <= 16 (layer, tap) segments; operand [M rows][taps * CC]
  lo.n_grp = 0; lo.grp_first[0] = 0;
  for (int l = 0, taps = 0; l < lo.K; ++l) {
    if (taps + (l + 1) > kMaxSeg) { lo.grp_taps[lo.n_grp++] = taps; lo.grp_first[lo.n_grp] = l; taps = 0; }
    taps += l + 1;
    if (l == lo.K - 1) { lo.grp_taps[lo.n_grp++] = taps; lo.grp_first[lo.n_grp] = lo.K; }
  }
  T2_REQUIRE(lo.n_grp <= 3, T2_ERR_UNSUPPORTED_SHAPE, "CBHG: conv bank too large for three data-gradient groups");
  for (int g = 0; g < lo.n_grp; ++g) {
    lo.k_bankT[g] = takeb(2LL * 128 * lo.grp_taps[g] * lo.CC);
    int slot = 0;
    for (int l = lo.grp_first[g]; l < lo.grp_first[g + 1]; ++l)
      for (int j = 0; j < lo.bank[l].k; ++j, ++slot)
        pj(lo.bank[l].p_k + (long long)j * lo.M * lo.CC, lo.M, lo.CC, lo.k_bankT[g], lo.grp_taps[g] * lo.CC, 0, slot * lo.CC);
  }
  const int Mp = 128;
  lo.k_dense = takeb(2LL * lo.HU * Mp); pj(lo.p_dk, lo.M, lo.HU, lo.k_dense, Mp, 1, 0);
  lo.k_denseT = takeb(2LL * 128 * lo.HU); pj(lo.p_dk, lo.M, lo.HU, lo.k_denseT, lo.HU, 0, 0);
  for (int i = 0; i < lo.NH; ++i) {
    lo.k_hw[i] = takeb(2LL * 2 * lo.HU * lo.HU);             // rows [H units | T units][K = HU]
    pj(lo.p_hk[i][0], lo.HU, lo.HU, lo.k_hw[i], lo.HU, 1, 0);
    { PJ j; j.src_off = lo.p_hk[i][1]; j.K = lo.HU; j.N = lo.HU; j.dst_off = lo.k_hw[i] / 2 + (long long)lo.HU * lo.HU; j.dst_ld = lo.HU; j.transpose = 1; j.col0 = 0; jobs.push_back(j); }
    lo.k_hwT[i] = takeb(2LL * lo.HU * 2 * lo.HU);            // [HU in][H units | T units]
    pj(lo.p_hk[i][0], lo.HU, lo.HU, lo.k_hwT[i], 2 * lo.HU, 0, 0);
    pj(lo.p_hk[i][1], lo.HU, lo.HU, lo.k_hwT[i], 2 * lo.HU, 0, lo.HU);
  }
  // GRU input projections: output columns [fw gates 2RU | fw cand RU | bw gates 2RU | bw cand RU], K = HU (the first HU kernel rows)
  const int XPW = 6 * lo.RU;
  lo.k_gx = takeb(2LL * XPW * lo.HU);
  lo.k_gxT = takeb(2LL * lo.HU * XPW);
  for (int d = 0; d < 2; ++d) {
    { PJ j; j.src_off = lo.p_gk[d]; j.K = lo.HU; j.N = 2 * lo.RU; j.dst_off = lo.k_gx / 2 + (long long)(d * 3 * lo.RU) * lo.HU; j.dst_ld = lo.HU; j.transpose = 1; j.col0 = 0; jobs.push_back(j); }
    { PJ j; j.src_off = lo.p_ck[d]; j.K = lo.HU; j.N = lo.RU; j.dst_off = lo.k_gx / 2 + (long long)(d * 3 * lo.RU + 2 * lo.RU) * lo.HU; j.dst_ld = lo.HU; j.transpose = 1; j.col0 = 0; jobs.push_back(j); }
    pj(lo.p_gk[d], lo.HU, 2 * lo.RU, lo.k_gxT, XPW, 0, d * 3 * lo.RU);
    pj(lo.p_ck[d], lo.HU, lo.RU, lo.k_gxT, XPW, 0, d * 3 * lo.RU + 2 * lo.RU);
  }
  lo.k_lin = takeb(2LL * lo.NFR * 2 * lo.RU); pj(lo.p_lk, 2 * lo.RU, lo.NF, lo.k_lin, 2 * lo.RU, 1, 0);
  const int NFK = (lo.NF + 63) / 64 * 64;
  lo.k_linT = takeb(2LL * 2 * lo.RU * NFK); pj(lo.p_lk, 2 * lo.RU, lo.NF, lo.k_linT, NFK, 0, 0);
  lo.packed_bytes = o;
  lo.n_jobs = int(jobs.size());

  // ---- workspace ----
  o = 0;
  const long long N = lo.N;
  lo.w_x0 = takeb(N * lo.M * 2);
  lo.w_Y = takeb(N * lo.KC * 2); lo.w_Xb = takeb(N * lo.KC * 2); lo.w_P = takeb(N * lo.KC * 2); lo.w_stb = takeb(8LL * lo.KC * 4);
  lo.w_Y1 = takeb(N * lo.PJc * 2); lo.w_X1 = takeb(N * lo.PJc * 2); lo.w_st1 = takeb(8LL * lo.PJc * 4);
  lo.w_Y2 = takeb(N * lo.M * 4); lo.w_st2 = takeb(8LL * 128 * 4);
  lo.w_hin = takeb(N * lo.M * 2);
  for (int i = 0; i <= lo.NH; ++i) { lo.w_hf[i] = takeb(N * lo.HU * 4); lo.w_hb[i] = takeb(N * lo.HU * 2); }
  for (int i = 0; i < lo.NH; ++i) lo.w_HT[i] = takeb(N * 2 * lo.HU * 2);
  lo.w_XP = takeb(N * XPW * 4);
  lo.w_out = takeb(N * 2 * lo.RU * 2);
  for (int d = 0; d < 2; ++d) { lo.w_gr[d] = takeb(N * lo.RU * 2); lo.w_gu[d] = takeb(N * lo.RU * 2); lo.w_gc[d] = takeb(N * lo.RU * 2); lo.w_grh[d] = takeb(N * lo.RU * 2); }
  lo.w_lin = takeb(N * lo.NFP * 4);            // fp32 [N][NFP]: padded pitch (the epilogue stores whole float4s)
  lo.w_scal = takeb(64 * 4);
  lo.w_tlen = takeb(lo.B * 4);
  // backward
  lo.w_dlin = takeb(N * lo.NFP * 2);
  lo.w_dout = takeb(N * 2 * lo.RU * 4);
  lo.w_dXP = takeb(N * XPW * 2);
  lo.w_dh = takeb(N * lo.HU * 4); lo.w_dhb = takeb(N * lo.HU * 2);
  lo.w_dHT = takeb(N * 2 * lo.HU * 2);
  lo.w_dhin = takeb(N * lo.M * 4);
  lo.w_dY2b = takeb(N * 128 * 2);
  lo.w_d1 = takeb(N * lo.PJc * 2); lo.w_d2 = takeb(N * lo.PJc * 2);
  lo.w_dP = takeb(N * lo.KC * 2); lo.w_dbank = takeb(N * lo.KC * 2);
  for (int i = 0; i < 3; ++i) lo.w_dx0[i] = takeb(N * 128 * 4);
  lo.w_bsum = takeb(2LL * lo.KC * 4);
  lo.w_tiles = takeb(4096 * sizeof(WgradTile));
  lo.w_jobs = takeb((long long)jobs.size() * sizeof(PJ));
  lo.n_reg = 0;
  for (auto& p : lo.params) lo.n_reg += p.reg ? 1 : 0;
  lo.w_regtab = takeb((long long)lo.n_reg * 2 * sizeof(long long));
  lo.workspace_bytes = o;
  if (jobs_out) jobs_out->swap(jobs);
  return T2_OK;
}

// ------------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------------
// fp32 [K][N] (TensorFlow [in][out]) -> bf16; transpose: dst[n][col0 + k] (K-major rows per output), else dst[k][col0 + n]
__global__ void cpack_kernel(const float* __restrict__ params, bf16* __restrict__ packed, const PJ* __restrict__ jobs) {
  const PJ j = jobs[blockIdx.y];
  const long long n_el = (long long)j.K * j.N;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n_el; e += (long long)gridDim.x * blockDim.x) {
    const int k = int(e / j.N), n = int(e % j.N);
    const bf16 v = __float2bfloat16(params[j.src_off + e]);
    if (j.transpose) packed[j.dst_off + (long long)n * j.dst_ld + j.col0 + k] = v;
    else packed[j.dst_off + (long long)k * j.dst_ld + j.col0 + n] = v;
  }
}
__global__ void f32_to_bf16_k(const float* __restrict__ in, bf16* __restrict__ out, long long n) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < n) out[e] = __float2bfloat16(in[e]);
}
__device__ __forceinline__ float ldv(const bf16* p, long long i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ float ldv(const float* p, long long i) { return p[i]; }
// Batch-norm kernels work on a COLUMN SLICE [c0, c0 + C) of row-pitch-ld matrices (the conv bank keeps its K layers side by side in
// one [N][K*CC] matrix but every layer owns its own gamma / beta / moving tensors); the statistics buffer has four sections of Ct
// floats: sum | sum of squares | mean | rstd, indexed by the absolute column.
template <typename TY>
__global__ void bn_stats_k(const TY* __restrict__ y, int ld, int c0, float* __restrict__ stats, int Ct, long long rows, int C) {
  const long long per = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * per, r1 = r0 + per < rows ? r0 + per : rows;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f, q = 0.f;
    for (long long r = r0; r < r1; ++r) { const float v = ldv(y, r * ld + c0 + c); s += v; q += v * v; }
    atomicAdd(stats + c0 + c, s); atomicAdd(stats + Ct + c0 + c, q);
  }
}
// batch norm (tf.layers.batch_normalization: eps 1e-3, biased batch variance, momentum 0.99): x = (y - mean) rstd gamma + beta.
// Writes mean / rstd into the statistics buffer and updates the moving statistics in training mode.
// Outputs: bf16 xb (same pitch / slice as y) and / or fp32 xf (dense [rows][C], + add).
template <typename TY>
__global__ void bn_apply_k(const TY* __restrict__ y, int ld, int c0, bf16* __restrict__ xb, float* __restrict__ xf, const float* __restrict__ add,
                           float* __restrict__ stats, int Ct, const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ mm,
                           float* __restrict__ mv, long long rows, int C, int training) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= rows * C) return;
  const int c = int(e % C);
  const long long r = e / C;
  float mean, rstd;
  if (training) {
    mean = stats[c0 + c] / float(rows);
    const float var = fmaxf(stats[Ct + c0 + c] / float(rows) - mean * mean, 0.f);
    rstd = rsqrtf(var + 1e-3f);
    if (e < C) {
      stats[2 * Ct + c0 + c] = mean; stats[3 * Ct + c0 + c] = rstd;
      mm[c] = 0.99f * mm[c] + 0.01f * mean; mv[c] = 0.99f * mv[c] + 0.01f * var;
    }
  } else { mean = mm[c]; rstd = rsqrtf(mv[c] + 1e-3f); }
  float v = (ldv(y, r * ld + c0 + c) - mean) * rstd * gamma[c] + beta[c];
  if (add) v += add[e];
  if (xb) xb[r * ld + c0 + c] = __float2bfloat16(v);
  if (xf) xf[e] = v;
}
// backward sums: bsum[c0 + c] = sum g, bsum[Ct + c0 + c] = sum g * xhat   (g: pitch ldg, same column slice)
template <typename TG, typename TY>
__global__ void bn_bwd_stats_k(const TG* __restrict__ g, int ldg, const TY* __restrict__ y, int ld, int c0, const float* __restrict__ stats, int Ct,
                               float* __restrict__ bsum, long long rows, int C) {
  const long long per = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * per, r1 = r0 + per < rows ? r0 + per : rows;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float mean = stats[2 * Ct + c0 + c], rstd = stats[3 * Ct + c0 + c];
    float s = 0.f, q = 0.f;
    for (long long r = r0; r < r1; ++r) {
      const float gv = ldv(g, r * ldg + c0 + c);
      s += gv; q += gv * (ldv(y, r * ld + c0 + c) - mean) * rstd;
    }
    atomicAdd(bsum + c0 + c, s); atomicAdd(bsum + Ct + c0 + c, q);
  }
}
// d(pre-activation) = act'(y) gamma rstd (g - mean(g) - xhat mean(g xhat)) -> dpre (bf16, pitch ldd, same column slice)
template <typename TG, typename TY>
__global__ void bn_bwd_apply_k(const TG* __restrict__ g, int ldg, const TY* __restrict__ y, int ld, int c0, const float* __restrict__ stats, int Ct,
                               const float* __restrict__ bsum, const float* __restrict__ gamma, bf16* __restrict__ dpre, int ldd,
                               float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows, int C, int act) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= rows * C) return;
  const int c = int(e % C);
  const long long r = e / C;
  const float mean = stats[2 * Ct + c0 + c], rstd = stats[3 * Ct + c0 + c];
  const float yv = ldv(y, r * ld + c0 + c);
  const float xhat = (yv - mean) * rstd;
  float dy = gamma[c] * rstd * (ldv(g, r * ldg + c0 + c) - bsum[c0 + c] / float(rows) - xhat * bsum[Ct + c0 + c] / float(rows));
  if (act == 1) dy = yv > 0.f ? dy : 0.f;
  dpre[r * ldd + c0 + c] = __float2bfloat16(dy);
  if (e < C) { dgamma[c] += bsum[Ct + c0 + c]; dbeta[c] += bsum[c0 + c]; }
}
// column sums of [rows][ld] (first C columns) added to dst
template <typename TS>
__global__ void colsum_k(const TS* __restrict__ src, long long rows, int C, int ld, float* __restrict__ dst) {
  const long long per = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * per, r1 = r0 + per < rows ? r0 + per : rows;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (long long r = r0; r < r1; ++r) s += ldv(src, r * ld + c);
    atomicAdd(dst + c, s);
  }
}
// tf.layers.max_pooling1d(pool 2, stride 1, 'same'): out[t] = max(x[t], x[t + 1]) (last step: x[t])
__global__ void maxpool_fwd_k(const bf16* __restrict__ x, bf16* __restrict__ out, long long N, int T, int C) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= N * C) return;
  const long long r = e / C;
  const int t = int(r % T);
  float v = __bfloat162float(x[e]);
  if (t + 1 < T) v = fmaxf(v, __bfloat162float(x[e + C]));
  out[e] = __float2bfloat16(v);
}
// gradient routing: x[t] receives dout[t] when it is the (first) maximum of window t and dout[t - 1] when it beats x[t - 1]
__global__ void maxpool_bwd_k(const bf16* __restrict__ x, const bf16* __restrict__ dout, bf16* __restrict__ dx, long long N, int T, int C) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= N * C) return;
  const long long r = e / C;
  const int t = int(r % T);
  const float v = __bfloat162float(x[e]);
  float g = 0.f;
  if (t + 1 >= T || v >= __bfloat162float(x[e + C])) g += __bfloat162float(dout[e]);
  if (t > 0 && v > __bfloat162float(x[e - C])) g += __bfloat162float(dout[e - C]);
  dx[e] = __float2bfloat16(g);
}
// highway layer (modules.py:12-16): pre [N][2HU] = [H pre-activation | T pre-activation] (biases added here);
// h' = relu(H) sigmoid(T) + h (1 - sigmoid(T)). Stashes relu(H) | sigmoid(T) in bf16 for the backward pass.
__global__ void highway_fwd_k(const float* __restrict__ pre, const float* __restrict__ bh, const float* __restrict__ bt, const float* __restrict__ h,
                              float* __restrict__ hf, bf16* __restrict__ hb, bf16* __restrict__ HT, long long N, int HU) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= N * HU) return;
  const long long r = e / HU; const int c = int(e % HU);
  const float Hh = fmaxf(pre[r * 2 * HU + c] + bh[c], 0.f);
  const float Tt = 1.f / (1.f + __expf(-(pre[r * 2 * HU + HU + c] + bt[c])));
  const float v = Hh * Tt + h[e] * (1.f - Tt);
  hf[e] = v; hb[e] = __float2bfloat16(v);
  if (HT) { HT[r * 2 * HU + c] = __float2bfloat16(Hh); HT[r * 2 * HU + HU + c] = __float2bfloat16(Tt); }
}
// dh' -> d[H pre | T pre] (bf16, GEMM operand) and the carry part dh * (1 - T) written to dcarry (fp32)
__global__ void highway_bwd_k(const float* __restrict__ dh, const bf16* __restrict__ HT, const float* __restrict__ h, bf16* __restrict__ dHT,
                              float* __restrict__ dcarry, long long N, int HU) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= N * HU) return;
  const long long r = e / HU; const int c = int(e % HU);
  const float Hh = __bfloat162float(HT[r * 2 * HU + c]), Tt = __bfloat162float(HT[r * 2 * HU + HU + c]);
  const float g = dh[e];
  dHT[r * 2 * HU + c] = __float2bfloat16(Hh > 0.f ? g * Tt : 0.f);
  dHT[r * 2 * HU + HU + c] = __float2bfloat16(g * (Hh - h[e]) * Tt * (1.f - Tt));
  dcarry[e] = g * (1.f - Tt);
}
// out fp32 += a (fp32) ; optional bf16 copy
__global__ void add_k(float* __restrict__ acc, const float* __restrict__ a, bf16* __restrict__ outb, long long n) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= n) return;
  const float v = acc[e] + a[e];
  acc[e] = v;
  if (outb) outb[e] = __float2bfloat16(v);
}
// linear outputs: clip (tacotron.py:218-219), L1 loss with priority on the low bins (:323-330 / MaskedLinearLoss), gradient seed.
// scal[0] += sum |t - o| * w over all bins, scal[1] += the same over bins < n_prio; normalisers are applied by the caller-side kernel.
__global__ void lin_finish_k(float* __restrict__ lin, const float* __restrict__ tgt, bf16* __restrict__ dlin, float* __restrict__ scal, long long N,
                             int T, int NF, int NFP, int n_prio, int clip, float lo, float hi, const int* __restrict__ tlen) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  float l_all = 0.f, l_low = 0.f;
  const float n_all = scal[8], n_low = scal[9];
  if (e < N * NFP) {
    const long long r = e / NFP; const int f = int(e % NFP);
    float g = 0.f;
    if (f < NF) {
      const long long o = r * NF + f;           // targets are dense [N][NF]; the outputs have pitch NFP
      const float raw = lin[e];
      const float v = clip ? fminf(fmaxf(raw, lo), hi) : raw;
      lin[e] = v;
      if (tgt) {
        const bool live = !tlen || int(r % T) < tlen[r / T];
        if (live) {
          const float d = v - tgt[o];
          l_all = fabsf(d);
          if (f < n_prio) l_low = l_all;
          const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
          g = sg * (0.5f / n_all + (f < n_prio ? 0.5f / n_low : 0.f));
          if (clip && (raw < lo || raw > hi)) g = 0.f;
        }
      }
    }
    if (dlin) dlin[e] = __float2bfloat16(g);
  }
  l_all = warp_sum(l_all); l_low = warp_sum(l_low);
  if ((threadIdx.x & 31) == 0 && tgt) { atomicAdd(scal + 0, l_all); atomicAdd(scal + 1, l_low); }
}
__global__ void reg_loss_k(const float* __restrict__ params, const long long* __restrict__ tab, int n, float* __restrict__ scal) {
  const long long off = tab[2 * blockIdx.y], cnt = tab[2 * blockIdx.y + 1];
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < cnt; i += (long long)gridDim.x * blockDim.x) { const float v = params[off + i]; s += v * v; }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) atomicAdd(scal + 2, 0.5f * s);
}
__global__ void reg_grad_k(const float* __restrict__ params, float* __restrict__ grads, const long long* __restrict__ tab, float w) {
  const long long off = tab[2 * blockIdx.y], cnt = tab[2 * blockIdx.y + 1];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < cnt; i += (long long)gridDim.x * blockDim.x) grads[off + i] += w * params[off + i];
}
// normalisers of the two L1 means: plain = (N NF, N n_prio); masked (MaskedLinearLoss) = sum(mask) for BOTH terms
__global__ void lin_norm_k(float* __restrict__ scal, const int* __restrict__ tlen, int B, int T, int NF, int n_prio) {
  if (!tlen) { scal[8] = float((long long)B * T) * NF; scal[9] = float((long long)B * T) * n_prio; return; }
  long long n = 0;
  for (int b = 0; b < B; ++b) n += tlen[b] < T ? tlen[b] : T;
  scal[8] = scal[9] = fmaxf(float(n) * NF, 1.f);
}
__global__ void loss_out_k(const float* __restrict__ scal, float* __restrict__ out, float regw) {
  out[0] = 0.5f * scal[0] / scal[8] + 0.5f * scal[1] / scal[9];
  out[1] = scal[2] * regw;
}
// dmel_out[r][m] = sum of the three partial data gradients of the conv bank + the residual path (d highway_in) ; rows of 128
__global__ void dmel_k(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c, const float* __restrict__ dhin,
                       float* __restrict__ out, long long N, int M) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= N * M) return;
  const long long r = e / M; const int m = int(e % M);
  out[e] = a[r * 128 + m] + b[r * 128 + m] + c[r * 128 + m] + dhin[e];
}

// ------------------------------------------------------------------------------------------------------
// GRU recurrence (tf.nn.rnn_cell.GRUCell inside bidirectional_dynamic_rnn, modules.py:34-35,69-75)
//   r, u = sigmoid(xg + h Wg_h) ; c = tanh(xc + (r h) Wc_h) ; h' = u h + (1 - u) c        (xg, xc: input projections incl. biases)
// One CTA = kGruItems batch items of one direction for all T steps; recurrent weights live in shared memory as bf16 PAIRS along k
// ([k/2][col] of bf16x2) so that a 4-byte load feeds two FMAs; state in fp32.
// ------------------------------------------------------------------------------------------------------
struct GruArgs {
  const float* params; long long p_gk[2], p_ck[2], p_gb[2], p_cb[2];
  const float* XP;          // [N][6RU] fp32
  bf16* out;                // [N][2RU]: h of direction d in columns [d RU, (d+1) RU)
  bf16 *r[2], *u[2], *c[2], *rh[2];   // stashes [N][RU] (nullable)
  int B, T, HU, RU;
};
constexpr int kRU = 128;
__global__ void __launch_bounds__(kGruThreads, 1) gru_fwd_kernel(GruArgs a) {
  extern __shared__ __align__(16) uint8_t gsm[];
  uint32_t* Wg = reinterpret_cast<uint32_t*>(gsm);                 // [64][256] bf16x2 (k pairs)
  uint32_t* Wc = Wg + 64 * 256;                                    // [64][128]
  float* h = reinterpret_cast<float*>(Wc + 64 * 128);              // [4][128]
  float* rhs = h + kGruItems * kRU;                                // [4][128]
  float* us = rhs + kGruItems * kRU;                               // [4][128]
  const int d = blockIdx.y, b0 = blockIdx.x * kGruItems, tid = threadIdx.x;
  const float* gk = a.params + a.p_gk[d] + (long long)a.HU * 2 * kRU;     // recurrent rows of the gates kernel [RU][2RU]
  const float* ck = a.params + a.p_ck[d] + (long long)a.HU * kRU;          // recurrent rows of the candidate kernel [RU][RU]
  for (int i = tid; i < 64 * 256; i += kGruThreads) {
    const int kp = i / 256, col = i % 256;
    Wg[i] = pack_bf16x2(gk[(2 * kp) * 256 + col], gk[(2 * kp + 1) * 256 + col]);
  }
  for (int i = tid; i < 64 * 128; i += kGruThreads) {
    const int kp = i / 128, col = i % 128;
    Wc[i] = pack_bf16x2(ck[(2 * kp) * 128 + col], ck[(2 * kp + 1) * 128 + col]);
  }
  for (int i = tid; i < kGruItems * kRU; i += kGruThreads) h[i] = 0.f;
  __syncthreads();
  const int XPW = 6 * kRU;
  const int j2 = tid & 127, half = tid >> 7;     // phase 2: column j2 of items {2 half, 2 half + 1}
  const float bg = a.params[a.p_gb[d] + tid], bc = a.params[a.p_cb[d] + j2];
  for (int s = 0; s < a.T; ++s) {
    const int t = d == 0 ? s : a.T - 1 - s;
    // phase 1: gate column `tid` (r: 0..127, u: 128..255) of all items
    float acc[kGruItems];
#pragma unroll
    for (int i = 0; i < kGruItems; ++i) acc[i] = b0 + i < a.B ? a.XP[((long long)(b0 + i) * a.T + t) * XPW + d * 3 * kRU + tid] + bg : 0.f;
#pragma unroll 4
    for (int kp = 0; kp < 64; ++kp) {
      const uint32_t w = Wg[kp * 256 + tid];
      const float w0 = bf16lo(w), w1 = bf16hi(w);
#pragma unroll
      for (int i = 0; i < kGruItems; ++i) {
        const float2 hv = *reinterpret_cast<const float2*>(h + i * kRU + 2 * kp);
        acc[i] += hv.x * w0 + hv.y * w1;
      }
    }
#pragma unroll
    for (int i = 0; i < kGruItems; ++i) {
      const float g = 1.f / (1.f + __expf(-acc[i]));
      if (tid < kRU) {
        const float rhv = g * h[i * kRU + tid];
        rhs[i * kRU + tid] = rhv;
        if (a.r[d] && b0 + i < a.B) {
          const long long o = ((long long)(b0 + i) * a.T + t) * kRU + tid;
          a.r[d][o] = __float2bfloat16(g); a.rh[d][o] = __float2bfloat16(rhv);
        }
      } else {
        us[i * kRU + tid - kRU] = g;
        if (a.u[d] && b0 + i < a.B) a.u[d][((long long)(b0 + i) * a.T + t) * kRU + tid - kRU] = __float2bfloat16(g);
      }
    }
    __syncthreads();
    // phase 2: candidate column j2 for two items, then the state update
    float cc[2];
#pragma unroll
    for (int q = 0; q < 2; ++q)
      cc[q] = b0 + 2 * half + q < a.B ? a.XP[((long long)(b0 + 2 * half + q) * a.T + t) * XPW + d * 3 * kRU + 2 * kRU + j2] + bc : 0.f;
#pragma unroll 4
    for (int kp = 0; kp < 64; ++kp) {
      const uint32_t w = Wc[kp * 128 + j2];
      const float w0 = bf16lo(w), w1 = bf16hi(w);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float2 v = *reinterpret_cast<const float2*>(rhs + (2 * half + q) * kRU + 2 * kp);
        cc[q] += v.x * w0 + v.y * w1;
      }
    }
    float hn[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int i = 2 * half + q;
      const float cv = tanhf(cc[q]);
      const float uv = us[i * kRU + j2];
      hn[q] = uv * h[i * kRU + j2] + (1.f - uv) * cv;
      const long long row = (long long)(b0 + i) * a.T + t;
      if (b0 + i < a.B) {
        a.out[row * 2 * kRU + d * kRU + j2] = __float2bfloat16(hn[q]);
        if (a.c[d]) a.c[d][row * kRU + j2] = __float2bfloat16(cv);
      }
    }
    __syncthreads();          // every thread has read h / us / rhs of this step
#pragma unroll
    for (int q = 0; q < 2; ++q) h[(2 * half + q) * kRU + j2] = hn[q];
    __syncthreads();
  }
}

// BPTT of the recurrence: walks the steps in reverse processing order, carries dh in shared memory, writes the gradients of the
// pre-activations [dr_pre | du_pre | dc_pre] (bf16) into dXP (the operand of the input-projection dgrad / wgrad GEMMs).
struct GruBwdArgs {
  const float* params; long long p_gk[2], p_ck[2];
  const float* dout;        // [N][2RU] fp32: upstream gradient of the outputs
  const bf16* out;          // [N][2RU] forward outputs (h_prev of a step = the output of the previously processed step)
  const bf16 *r[2], *u[2], *c[2];
  bf16* dXP;                // [N][6RU]
  int B, T, HU, RU;
};
__global__ void __launch_bounds__(kGruThreads, 1) gru_bwd_kernel(GruBwdArgs a) {
  extern __shared__ __align__(16) uint8_t gsm[];
  uint32_t* WgT = reinterpret_cast<uint32_t*>(gsm);                // [128 (j pairs of 256 gate cols)][128 k] : bf16x2 over gate columns j
  uint32_t* WcT = WgT + 128 * 128;                                 // [64 (j pairs of 128 cand cols)][128 k]
  float* dh = reinterpret_cast<float*>(WcT + 64 * 128);            // [4][128] carried gradient
  float* dcp = dh + kGruItems * kRU;                               // [4][128]
  float* dgp = dcp + kGruItems * kRU;                              // [4][256]: dr_pre | du_pre
  const int d = blockIdx.y, b0 = blockIdx.x * kGruItems, tid = threadIdx.x;
  const float* gk = a.params + a.p_gk[d] + (long long)a.HU * 2 * kRU;
  const float* ck = a.params + a.p_ck[d] + (long long)a.HU * kRU;
  for (int i = tid; i < 128 * 128; i += kGruThreads) {
    const int jp = i / 128, k = i % 128;
    WgT[i] = pack_bf16x2(gk[k * 256 + 2 * jp], gk[k * 256 + 2 * jp + 1]);
  }
  for (int i = tid; i < 64 * 128; i += kGruThreads) {
    const int jp = i / 128, k = i % 128;
    WcT[i] = pack_bf16x2(ck[k * 128 + 2 * jp], ck[k * 128 + 2 * jp + 1]);
  }
  for (int i = tid; i < kGruItems * kRU; i += kGruThreads) dh[i] = 0.f;
  __syncthreads();
  const int XPW = 6 * kRU;
  const int k2 = tid & 127, half = tid >> 7;     // thread owns unit k2 of items {2 half, 2 half + 1}
  for (int s = a.T - 1; s >= 0; --s) {
    const int t = d == 0 ? s : a.T - 1 - s;
    const int tp = d == 0 ? t - 1 : t + 1;       // time index of the previously processed step (h_prev)
    float g[2], hp[2], uv[2], rv[2], du[2], part[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int i = 2 * half + q;
      const bool live = b0 + i < a.B;               // a batch that is not a multiple of the CTA's item count: idle lanes carry zeros
      const long long row = live ? (long long)(b0 + i) * a.T + t : 0;
      g[q] = live ? dh[i * kRU + k2] + a.dout[row * 2 * kRU + d * kRU + k2] : 0.f;
      hp[q] = (live && s > 0) ? __bfloat162float(a.out[((long long)(b0 + i) * a.T + tp) * 2 * kRU + d * kRU + k2]) : 0.f;
      uv[q] = live ? __bfloat162float(a.u[d][row * kRU + k2]) : 0.f;
      rv[q] = live ? __bfloat162float(a.r[d][row * kRU + k2]) : 0.f;
      const float cv = live ? __bfloat162float(a.c[d][row * kRU + k2]) : 0.f;
      du[q] = g[q] * (hp[q] - cv);
      const float dc = g[q] * (1.f - uv[q]);
      const float dcpv = dc * (1.f - cv * cv);
      dcp[i * kRU + k2] = dcpv;
      if (live) a.dXP[row * XPW + d * 3 * kRU + 2 * kRU + k2] = __float2bfloat16(dcpv);
      part[q] = g[q] * uv[q];
    }
    __syncthreads();
    // d(r h)[k2] = sum_j dc_pre[j] Wc_h[k2][j]
    float drh[2] = {0.f, 0.f};
#pragma unroll 4
    for (int jp = 0; jp < 64; ++jp) {
      const uint32_t w = WcT[jp * 128 + k2];
      const float w0 = bf16lo(w), w1 = bf16hi(w);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float2 v = *reinterpret_cast<const float2*>(dcp + (2 * half + q) * kRU + 2 * jp);
        drh[q] += v.x * w0 + v.y * w1;
      }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int i = 2 * half + q;
      const long long row = (long long)(b0 + i) * a.T + t;
      const float drp = drh[q] * hp[q] * rv[q] * (1.f - rv[q]);
      const float dup = du[q] * uv[q] * (1.f - uv[q]);
      dgp[i * 2 * kRU + k2] = drp; dgp[i * 2 * kRU + kRU + k2] = dup;
      if (b0 + i < a.B) {
        a.dXP[row * XPW + d * 3 * kRU + k2] = __float2bfloat16(drp);
        a.dXP[row * XPW + d * 3 * kRU + kRU + k2] = __float2bfloat16(dup);
      }
      part[q] += drh[q] * rv[q];
    }
    __syncthreads();
    // dh_prev[k2] += sum_j [dr_pre | du_pre][j] Wg_h[k2][j]
#pragma unroll 4
    for (int jp = 0; jp < 128; ++jp) {
      const uint32_t w = WgT[jp * 128 + k2];
      const float w0 = bf16lo(w), w1 = bf16hi(w);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float2 v = *reinterpret_cast<const float2*>(dgp + (2 * half + q) * 2 * kRU + 2 * jp);
        part[q] += v.x * w0 + v.y * w1;
      }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) dh[(2 * half + q) * kRU + k2] = part[q];   // only this thread reads / writes dh[i][k2]
    __syncthreads();          // dcp / dgp are rewritten by the next step
  }
}

// ------------------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------------------
// out[pos][n] = act(sum_taps sum_k a[pos + shift][k0 + k] w[n][tap * Cp + k] + bias[n]) on the tcgen05 engine (EPI_BIAS_ACT)
int gemm(const void* a, int C, int ld, int k0, long long T, int Bn, const void* w, int wN, int wK, int ntaps, const int* shifts, int BN,
         const float* bias, int act, void* out_bf16, float* out_f32, int ldo, int nvalid, cudaStream_t st, const int* k0s = nullptr, int Ctot = 0) {
  ActGemmCall g;
  memset(&g, 0, sizeof(g));
  const int nkb = (C + kBK - 1) / kBK;
  g.a[0] = make_act(a, Ctot > 0 ? Ctot : k0 + C, int(T), Bn, 1, ld); g.na = 1;
  T2_REQUIRE(ntaps <= kMaxSeg, T2_ERR_UNSUPPORTED_SHAPE, "CBHG gemm: too many taps");
  for (int s = 0; s < ntaps; ++s) g.seg[s] = Seg{0, shifts ? shifts[s] : 0, k0s ? k0s[s] : k0, nkb, 0, 1};
  g.nseg = ntaps;
  g.w = w; g.wN = wN; g.wK = wK; g.wL = 1;
  g.T = int(T); g.B = Bn; g.n_tiles = (nvalid + BN - 1) / BN;
  g.epi.ptr[0] = out_bf16; g.epi.ptr[1] = const_cast<float*>(bias); g.epi.ptr[2] = out_f32;
  g.epi.i[0] = ldo; g.epi.i[1] = act; g.epi.i[2] = nvalid;
  return launch_act_gemm(EPI_BIAS_ACT, BN, g, st);
}

void wg_tile(std::vector<WgradTile>& v, int am, int ach, int ash, int bm, int bch, long long off, int ldc, int mv, int nv) {
  WgradTile t; memset(&t, 0, sizeof(t));
  t.a_map = am; t.a_ch0 = ach; t.a_shift = ash; t.b_map = bm; t.b_ch0 = bch; t.out_off = off; t.ldc = ldc;
  t.m_valid = mv; t.n_valid = nv; t.scale = 1.f; t.accumulate = 0; t.div = nullptr;
  v.push_back(t);
}
void wg_dense(std::vector<WgradTile>& v, int am, int a0, int Ca, int bm, int b0, int Cb, long long off, int ldc, int shift = 0) {
  for (int m0 = 0; m0 < Ca; m0 += 128)
    for (int n0 = 0; n0 < Cb; n0 += 256)
      wg_tile(v, am, a0 + m0, shift, bm, b0 + n0, off + (long long)m0 * ldc + n0, ldc, Ca - m0 < 128 ? Ca - m0 : 128, Cb - n0 < 256 ? Cb - n0 : 256);
}
inline int conv_shift(int k, int j) { return j - (k - 1) / 2; }
// wgrad launches, in the order t2_cbhg_backward issues them
enum { WG_LIN = 0, WG_GRU = 1, WG_HW0 = 2 /* NH launches, last highway layer first */ };
void build_tiles(const CL& lo, std::vector<std::vector<WgradTile>>& L) {
  L.clear();
  const int RU = lo.RU, HU = lo.HU;
  { std::vector<WgradTile> w; wg_dense(w, 0, 0, 2 * RU, 1, 0, lo.NF, lo.p_lk, lo.NF); L.push_back(w); }      // maps: 0 rnn out, 1 dlin
  { std::vector<WgradTile> w;     // maps: 0 h_last (bf16 [N][HU]), 1 dXP, 2 rnn out, 3 rh fw, 4 rh bw
    for (int d = 0; d < 2; ++d) {
      wg_dense(w, 0, 0, HU, 1, d * 3 * RU, 2 * RU, lo.p_gk[d], 2 * RU);                                       // input rows of the gates kernel
      wg_dense(w, 0, 0, HU, 1, d * 3 * RU + 2 * RU, RU, lo.p_ck[d], RU);                                      // input rows of the candidate kernel
      wg_dense(w, 2, d * RU, RU, 1, d * 3 * RU, 2 * RU, lo.p_gk[d] + (long long)HU * 2 * RU, 2 * RU, d == 0 ? -1 : 1);   // h_prev x d gates
      wg_dense(w, 3 + d, 0, RU, 1, d * 3 * RU + 2 * RU, RU, lo.p_ck[d] + (long long)HU * RU, RU);             // (r h_prev) x d cand
    }
    L.push_back(w); }
  for (int i = lo.NH - 1; i >= 0; --i) {   // maps: 0 h_i (bf16), 1 dHT
    std::vector<WgradTile> w;
    wg_dense(w, 0, 0, HU, 1, 0, HU, lo.p_hk[i][0], HU);
    wg_dense(w, 0, 0, HU, 1, HU, HU, lo.p_hk[i][1], HU);
    L.push_back(w);
  }
  { std::vector<WgradTile> w; wg_dense(w, 0, 0, lo.M, 1, 0, HU, lo.p_dk, HU); L.push_back(w); }               // dense: hin x dh0
  auto conv = [&](const CConv& c, int a_ch0, int b_ch0) {
    std::vector<WgradTile> w;
    for (int j = 0; j < c.k; ++j) wg_dense(w, 0, a_ch0, c.cin, 1, b_ch0, c.cout, c.p_k + (long long)j * c.cin * c.cout, c.cout, conv_shift(c.k, j));
    L.push_back(w);
  };
  conv(lo.proj2, 0, 0);     // maps: 0 X1, 1 dY2b
  conv(lo.proj1, 0, 0);     // maps: 0 P, 1 d1
  { std::vector<WgradTile> w;  // bank: maps 0 x0, 1 dbank (channel block k-1)
    for (int k = 1; k <= lo.K; ++k) {
      const CConv& c = lo.bank[k - 1];
      for (int j = 0; j < c.k; ++j) wg_dense(w, 0, 0, c.cin, 1, (k - 1) * lo.CC, c.cout, c.p_k + (long long)j * c.cin * c.cout, c.cout, conv_shift(c.k, j));
    }
    L.push_back(w); }
}

size_t gru_fwd_smem() { return (64 * 256 + 64 * 128) * 4 + 3 * kGruItems * kRU * 4; }
size_t gru_bwd_smem() { return (128 * 128 + 64 * 128) * 4 + (2 * kGruItems * kRU + kGruItems * 2 * kRU) * 4; }

}  // namespace
}  // namespace t2

using namespace t2;

extern "C" int t2_cbhg_sizes(const t2_cbhg_config_t* cfg, long long* n_params, long long* packed_bytes, long long* workspace_bytes, int* n_tensors) {
  CL lo;
  int rc = build(cfg, lo, nullptr);
  if (rc) return rc;
  if (n_params) *n_params = lo.n_params;
  if (packed_bytes) *packed_bytes = lo.packed_bytes;
  if (workspace_bytes) *workspace_bytes = lo.workspace_bytes;
  if (n_tensors) *n_tensors = int(lo.params.size());
  return T2_OK;
}

extern "C" int t2_cbhg_param_info(const t2_cbhg_config_t* cfg, int i, char* name, int cap, long long* offset, int* ndim, int* shape4, int* trainable) {
  CL lo;
  int rc = build(cfg, lo, nullptr);
  if (rc) return rc;
  T2_REQUIRE(i >= 0 && i < int(lo.params.size()) && name && cap > 0, T2_ERR_INVALID_ARG, "cbhg_param_info: bad index");
  const CPT& p = lo.params[i];
  snprintf(name, cap, "%s", p.name.c_str());
  if (offset) *offset = p.off;
  if (ndim) *ndim = p.ndim;
  if (shape4) for (int k = 0; k < 4; ++k) shape4[k] = p.shape[k];
  if (trainable) *trainable = p.trainable ? 1 : 0;
  return T2_OK;
}

extern "C" int t2_cbhg_init(const t2_cbhg_config_t* cfg, void* d_packed, void* d_workspace, void* stream) {
  CL lo;
  std::vector<PJ> jobs;
  int rc = build(cfg, lo, &jobs);
  if (rc) return rc;
  T2_REQUIRE(d_packed && d_workspace, T2_ERR_INVALID_ARG, "cbhg_init: null buffers");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  T2_CHECK_CUDA(cudaMemsetAsync(d_packed, 0, lo.packed_bytes, st));
  T2_CHECK_CUDA(cudaMemsetAsync(d_workspace, 0, lo.workspace_bytes, st));
  T2_CHECK_CUDA(cudaMemcpyAsync(ws + lo.w_jobs, jobs.data(), jobs.size() * sizeof(PJ), cudaMemcpyHostToDevice, st));
  std::vector<std::vector<WgradTile>> wl;
  build_tiles(lo, wl);
  std::vector<WgradTile> all;
  for (auto& w : wl) all.insert(all.end(), w.begin(), w.end());
  T2_REQUIRE(all.size() <= 4096, T2_ERR_UNSUPPORTED_SHAPE, "cbhg: too many wgrad tiles (%d)", int(all.size()));
  T2_CHECK_CUDA(cudaMemcpyAsync(ws + lo.w_tiles, all.data(), all.size() * sizeof(WgradTile), cudaMemcpyHostToDevice, st));
  std::vector<long long> tab;
  for (auto& p : lo.params)
    if (p.reg) { long long n = 1; for (int k = 0; k < p.ndim; ++k) n *= p.shape[k]; tab.push_back(p.off); tab.push_back(n); }
  if (!tab.empty()) T2_CHECK_CUDA(cudaMemcpyAsync(ws + lo.w_regtab, tab.data(), tab.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
  T2_CHECK_CUDA(cudaFuncSetAttribute(gru_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(gru_fwd_smem())));
  T2_CHECK_CUDA(cudaFuncSetAttribute(gru_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(gru_bwd_smem())));
  T2_CHECK_CUDA(cudaStreamSynchronize(st));
  return T2_OK;
}

extern "C" int t2_cbhg_pack_weights(const t2_cbhg_config_t* cfg, const float* d_params, void* d_packed, void* d_workspace, void* stream) {
  CL lo;
  int rc = build(cfg, lo, nullptr);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cpack_kernel<<<dim3(32, lo.n_jobs), 256, 0, st>>>(d_params, static_cast<bf16*>(d_packed),
                                                    reinterpret_cast<const PJ*>(static_cast<uint8_t*>(d_workspace) + lo.w_jobs));
  t2_count_launch();
  T2_CHECK_CUDA(cudaGetLastError());
  return T2_OK;
}

extern "C" int t2_cbhg_set_target_lengths(const t2_cbhg_config_t* cfg, void* d_workspace, const int* d_target_lengths, void* stream) {
  CL lo;
  int rc = build(cfg, lo, nullptr);
  if (rc) return rc;
  T2_REQUIRE(d_workspace && d_target_lengths, T2_ERR_INVALID_ARG, "cbhg_set_target_lengths: null pointer");
  T2_CHECK_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(d_workspace) + lo.w_tlen, d_target_lengths, lo.B * sizeof(int), cudaMemcpyDeviceToDevice,
                                static_cast<cudaStream_t>(stream)));
  return T2_OK;
}

namespace {
struct Ctx { const CL* lo; uint8_t* ws; const uint8_t* pk; float* params; cudaStream_t st; int training; };
template <typename T> T* W(const Ctx& s, long long off) { return reinterpret_cast<T*>(s.ws + off); }

// conv (+ bias, activation) into `y` (bf16 [N][ldo] column slice or fp32), batch-norm statistics are taken by the caller
int conv_fwd(const Ctx& s, const CConv& L, const void* x, int ld_x, bf16* y_b, float* y_f, int ldo) {
  int shifts[16];
  for (int j = 0; j < L.k; ++j) shifts[j] = conv_shift(L.k, j);
  const int BN = L.cout % 256 == 0 ? 256 : 128;
  return gemm(x, L.cin, ld_x, 0, s.lo->T, s.lo->B, s.pk + L.k_w, (L.cout + 127) / 128 * 128, L.k * L.cinp, L.k, shifts, BN, s.params + L.p_b, L.act, y_b, y_f,
              ldo, L.cout, s.st);
}
}  // namespace

extern "C" int t2_cbhg_forward(const t2_cbhg_config_t* cfg, float* d_params, const void* d_packed, void* d_workspace, const float* d_mel,
                               const float* d_linear_targets, float* d_loss, int training, void* stream) {
  CL lo;
  int rc = build(cfg, lo, nullptr);
  if (rc) return rc;
  T2_REQUIRE(d_params && d_packed && d_workspace && d_mel, T2_ERR_INVALID_ARG, "cbhg_forward: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Ctx s{&lo, static_cast<uint8_t*>(d_workspace), static_cast<const uint8_t*>(d_packed), d_params, st, training};
  const long long N = lo.N;
  const int T = lo.T, B = lo.B, M = lo.M, HU = lo.HU, RU = lo.RU, KC = lo.KC, PJc = lo.PJc;
  float* scal = W<float>(s, lo.w_scal);
  T2_CHECK_CUDA(cudaMemsetAsync(scal, 0, 16 * sizeof(float), st));
  bf16* x0 = W<bf16>(s, lo.w_x0);
  f32_to_bf16_k<<<g1(N * M), 256, 0, st>>>(d_mel, x0, N * M); t2_count_launch();
  // ---- conv bank (each layer writes its 128-column slice) + per-layer batch norm ----
  bf16* Y = W<bf16>(s, lo.w_Y);
  bf16* Xb = W<bf16>(s, lo.w_Xb);
  float* stb = W<float>(s, lo.w_stb);
  if (training) T2_CHECK_CUDA(cudaMemsetAsync(stb, 0, 2LL * KC * sizeof(float), st));
  for (int k = 1; k <= lo.K; ++k) {
    const CConv& L = lo.bank[k - 1];
    const int c0 = (k - 1) * lo.CC;
    rc = conv_fwd(s, L, x0, M, Y + c0, nullptr, KC);
    if (rc) return rc;
    if (training) { bn_stats_k<bf16><<<64, 128, 0, st>>>(Y, KC, c0, stb, KC, N, lo.CC); t2_count_launch(); }
    bn_apply_k<bf16><<<g1(N * lo.CC), 256, 0, st>>>(Y, KC, c0, Xb, nullptr, nullptr, stb, KC, d_params + L.p_g, d_params + L.p_be, d_params + L.p_mm,
                                                    d_params + L.p_mv, N, lo.CC, training); t2_count_launch();
  }
  bf16* P = W<bf16>(s, lo.w_P);
  maxpool_fwd_k<<<g1(N * KC), 256, 0, st>>>(Xb, P, N, T, KC); t2_count_launch();
  // ---- projections ----
  bf16* Y1 = W<bf16>(s, lo.w_Y1); bf16* X1 = W<bf16>(s, lo.w_X1); float* st1 = W<float>(s, lo.w_st1);
  rc = conv_fwd(s, lo.proj1, P, KC, Y1, nullptr, PJc);
  if (rc) return rc;
  if (training) {
    T2_CHECK_CUDA(cudaMemsetAsync(st1, 0, 2LL * PJc * sizeof(float), st));
    bn_stats_k<bf16><<<64, 256, 0, st>>>(Y1, PJc, 0, st1, PJc, N, PJc); t2_count_launch();
  }
  bn_apply_k<bf16><<<g1(N * PJc), 256, 0, st>>>(Y1, PJc, 0, X1, nullptr, nullptr, st1, PJc, d_params + lo.proj1.p_g, d_params + lo.proj1.p_be,
                                                d_params + lo.proj1.p_mm, d_params + lo.proj1.p_mv, N, PJc, training); t2_count_launch();
  float* Y2 = W<float>(s, lo.w_Y2); float* st2 = W<float>(s, lo.w_st2);
  rc = conv_fwd(s, lo.proj2, X1, PJc, nullptr, Y2, M);
  if (rc) return rc;
  if (training) {
    T2_CHECK_CUDA(cudaMemsetAsync(st2, 0, 2LL * M * sizeof(float), st));
    bn_stats_k<float><<<64, 128, 0, st>>>(Y2, M, 0, st2, M, N, M); t2_count_launch();
  }
  // highway input = BN(proj2) + mel_outputs (modules.py:59); the fp32 sum goes through w_dhin (free until the backward pass)
  float* hin_f = W<float>(s, lo.w_dhin);
  bf16* hin = W<bf16>(s, lo.w_hin);
  bn_apply_k<float><<<g1(N * M), 256, 0, st>>>(Y2, M, 0, nullptr, hin_f, d_mel, st2, M, d_params + lo.proj2.p_g, d_params + lo.proj2.p_be,
                                               d_params + lo.proj2.p_mm, d_params + lo.proj2.p_mv, N, M, training); t2_count_launch();
  f32_to_bf16_k<<<g1(N * M), 256, 0, st>>>(hin_f, hin, N * M); t2_count_launch();
  // ---- dense to the highway width, highway layers ----
  rc = gemm(hin, M, M, 0, T, B, s.pk + lo.k_dense, HU, 128, 1, nullptr, 128, d_params + lo.p_db, 0, W<bf16>(s, lo.w_hb[0]), W<float>(s, lo.w_hf[0]), HU, HU, st);
  if (rc) return rc;
  float* pre = W<float>(s, lo.w_XP);       // [N][2HU] scratch (the GRU input projections overwrite it afterwards)
  for (int i = 0; i < lo.NH; ++i) {
    rc = gemm(W<bf16>(s, lo.w_hb[i]), HU, HU, 0, T, B, s.pk + lo.k_hw[i], 2 * HU, HU, 1, nullptr, 256, nullptr, 0, nullptr, pre, 2 * HU, 2 * HU, st);
    if (rc) return rc;
    highway_fwd_k<<<g1(N * HU), 256, 0, st>>>(pre, d_params + lo.p_hb[i][0], d_params + lo.p_hb[i][1], W<float>(s, lo.w_hf[i]), W<float>(s, lo.w_hf[i + 1]),
                                              W<bf16>(s, lo.w_hb[i + 1]), training ? W<bf16>(s, lo.w_HT[i]) : nullptr, N, HU); t2_count_launch();
  }
  // ---- bidirectional GRU ----
  const int XPW = 6 * RU;
  float* XP = W<float>(s, lo.w_XP);
  rc = gemm(W<bf16>(s, lo.w_hb[lo.NH]), HU, HU, 0, T, B, s.pk + lo.k_gx, XPW, HU, 1, nullptr, 256, nullptr, 0, nullptr, XP, XPW, XPW, st);
  if (rc) return rc;
  {
    GruArgs a;
    memset(&a, 0, sizeof(a));
    a.params = d_params;
    for (int d = 0; d < 2; ++d) {
      a.p_gk[d] = lo.p_gk[d]; a.p_ck[d] = lo.p_ck[d]; a.p_gb[d] = lo.p_gb[d]; a.p_cb[d] = lo.p_cb[d];
      if (training) { a.r[d] = W<bf16>(s, lo.w_gr[d]); a.u[d] = W<bf16>(s, lo.w_gu[d]); a.c[d] = W<bf16>(s, lo.w_gc[d]); a.rh[d] = W<bf16>(s, lo.w_grh[d]); }
    }
    a.XP = XP; a.out = W<bf16>(s, lo.w_out); a.B = B; a.T = T; a.HU = HU; a.RU = RU;
    gru_fwd_kernel<<<dim3((B + kGruItems - 1) / kGruItems, 2), kGruThreads, gru_fwd_smem(), st>>>(a); t2_count_launch();
    T2_CHECK_CUDA(cudaGetLastError());
  }
  // ---- linear projection, clip, loss ----
  float* lin = W<float>(s, lo.w_lin);
  rc = gemm(W<bf16>(s, lo.w_out), 2 * RU, 2 * RU, 0, T, B, s.pk + lo.k_lin, lo.NFR, 2 * RU, 1, nullptr, 128, d_params + lo.p_lb, 0, nullptr, lin, lo.NFP, lo.NF, st);
  if (rc) return rc;
  const int* tlen = lo.c.mask_decoder ? W<int>(s, lo.w_tlen) : nullptr;
  const float lo_c = -lo.c.max_abs_value - lo.c.lower_bound_decay, hi_c = lo.c.max_abs_value;
  lin_norm_k<<<1, 1, 0, st>>>(scal, tlen, B, T, lo.NF, lo.c.n_priority_freq); t2_count_launch();
  lin_finish_k<<<g1(N * lo.NFP), 256, 0, st>>>(lin, d_linear_targets, (training && d_linear_targets) ? W<bf16>(s, lo.w_dlin) : nullptr, scal, N, T, lo.NF,
                                               lo.NFP, lo.c.n_priority_freq, lo.c.clip_outputs, lo_c, hi_c, tlen); t2_count_launch();
  if (d_loss) {
    if (lo.n_reg > 0) { reg_loss_k<<<dim3(8, lo.n_reg), 256, 0, st>>>(d_params, W<long long>(s, lo.w_regtab), lo.n_reg, scal); t2_count_launch(); }
    loss_out_k<<<1, 1, 0, st>>>(scal, d_loss, lo.c.reg_weight); t2_count_launch();
  }
  T2_CHECK_CUDA(cudaGetLastError());
  return T2_OK;
}

extern "C" int t2_cbhg_backward(const t2_cbhg_config_t* cfg, const float* d_params, const void* d_packed, void* d_workspace, const float* d_mel,
                                float* d_grads, float* d_mel_grad, void* stream) {
  CL lo;
  int rc = build(cfg, lo, nullptr);
  if (rc) return rc;
  T2_REQUIRE(d_params && d_packed && d_workspace && d_mel && d_grads && d_mel_grad, T2_ERR_INVALID_ARG, "cbhg_backward: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Ctx s{&lo, static_cast<uint8_t*>(d_workspace), static_cast<const uint8_t*>(d_packed), const_cast<float*>(d_params), st, 1};
  const long long N = lo.N;
  const int T = lo.T, B = lo.B, M = lo.M, HU = lo.HU, RU = lo.RU, KC = lo.KC, PJc = lo.PJc, XPW = 6 * lo.RU;
  std::vector<std::vector<WgradTile>> wl;
  build_tiles(lo, wl);
  std::vector<int> toff(wl.size());
  { int o = 0; for (size_t i = 0; i < wl.size(); ++i) { toff[i] = o; o += int(wl[i].size()); } }
  const WgradTile* tiles = W<WgradTile>(s, lo.w_tiles);
  int li = 0;
  auto wgrad = [&](const ActT* maps, int nmaps) {
    int r = launch_wgrad(maps, nmaps, tiles + toff[li], int(wl[li].size()), d_grads, T, B, st);
    ++li;
    return r;
  };
  T2_CHECK_CUDA(cudaMemsetAsync(d_grads, 0, lo.n_params * sizeof(float), st));
  // ---- linear projection ----
  bf16* dlin = W<bf16>(s, lo.w_dlin);
  bf16* out = W<bf16>(s, lo.w_out);
  float* dout = W<float>(s, lo.w_dout);
  const int NFK = (lo.NF + 63) / 64 * 64;
  rc = gemm(dlin, lo.NF, lo.NFP, 0, T, B, s.pk + lo.k_linT, 2 * RU, NFK, 1, nullptr, 256, nullptr, 0, nullptr, dout, 2 * RU, 2 * RU, st);
  if (rc) return rc;
  { ActT maps[2] = {make_act(out, 2 * RU, T, B), make_act(dlin, lo.NF, T, B, 1, lo.NFP)}; rc = wgrad(maps, 2); if (rc) return rc; }
  colsum_k<bf16><<<64, 256, 0, st>>>(dlin, N, lo.NF, lo.NFP, d_grads + lo.p_lb); t2_count_launch();
  // ---- GRU ----
  bf16* dXP = W<bf16>(s, lo.w_dXP);
  {
    GruBwdArgs a;
    memset(&a, 0, sizeof(a));
    a.params = d_params;
    for (int d = 0; d < 2; ++d) {
      a.p_gk[d] = lo.p_gk[d]; a.p_ck[d] = lo.p_ck[d];
      a.r[d] = W<bf16>(s, lo.w_gr[d]); a.u[d] = W<bf16>(s, lo.w_gu[d]); a.c[d] = W<bf16>(s, lo.w_gc[d]);
    }
    a.dout = dout; a.out = out; a.dXP = dXP; a.B = B; a.T = T; a.HU = HU; a.RU = RU;
    gru_bwd_kernel<<<dim3((B + kGruItems - 1) / kGruItems, 2), kGruThreads, gru_bwd_smem(), st>>>(a); t2_count_launch();
    T2_CHECK_CUDA(cudaGetLastError());
  }
  {
    ActT maps[5] = {make_act(W<bf16>(s, lo.w_hb[lo.NH]), HU, T, B), make_act(dXP, XPW, T, B), make_act(out, 2 * RU, T, B),
                    make_act(W<bf16>(s, lo.w_grh[0]), RU, T, B), make_act(W<bf16>(s, lo.w_grh[1]), RU, T, B)};
    rc = wgrad(maps, 5); if (rc) return rc;
  }
  for (int d = 0; d < 2; ++d) {
    colsum_k<bf16><<<64, 256, 0, st>>>(dXP + d * 3 * RU, N, 2 * RU, XPW, d_grads + lo.p_gb[d]); t2_count_launch();
    colsum_k<bf16><<<64, 128, 0, st>>>(dXP + d * 3 * RU + 2 * RU, N, RU, XPW, d_grads + lo.p_cb[d]); t2_count_launch();
  }
  float* dh = W<float>(s, lo.w_dh);
  rc = gemm(dXP, XPW, XPW, 0, T, B, s.pk + lo.k_gxT, HU, XPW, 1, nullptr, 128, nullptr, 0, nullptr, dh, HU, HU, st);
  if (rc) return rc;
  // ---- highway layers (last first) ----
  bf16* dHT = W<bf16>(s, lo.w_dHT);
  float* dcar = W<float>(s, lo.w_XP);        // [N][HU] fp32 scratch (the forward input projections are no longer needed)
  for (int i = lo.NH - 1; i >= 0; --i) {
    highway_bwd_k<<<g1(N * HU), 256, 0, st>>>(dh, W<bf16>(s, lo.w_HT[i]), W<float>(s, lo.w_hf[i]), dHT, dcar, N, HU); t2_count_launch();
    rc = gemm(dHT, 2 * HU, 2 * HU, 0, T, B, s.pk + lo.k_hwT[i], HU, 2 * HU, 1, nullptr, 128, nullptr, 0, nullptr, dh, HU, HU, st);
    if (rc) return rc;
    add_k<<<g1(N * HU), 256, 0, st>>>(dh, dcar, i == 0 ? W<bf16>(s, lo.w_dhb) : nullptr, N * HU); t2_count_launch();
    { ActT maps[2] = {make_act(W<bf16>(s, lo.w_hb[i]), HU, T, B), make_act(dHT, 2 * HU, T, B)}; rc = wgrad(maps, 2); if (rc) return rc; }
    colsum_k<bf16><<<64, 128, 0, st>>>(dHT, N, HU, 2 * HU, d_grads + lo.p_hb[i][0]); t2_count_launch();
    colsum_k<bf16><<<64, 128, 0, st>>>(dHT + HU, N, HU, 2 * HU, d_grads + lo.p_hb[i][1]); t2_count_launch();
  }
  // ---- dense ----
  bf16* dhb = W<bf16>(s, lo.w_dhb);
  float* dhin = W<float>(s, lo.w_dhin);
  rc = gemm(dhb, HU, HU, 0, T, B, s.pk + lo.k_denseT, 128, HU, 1, nullptr, 128, nullptr, 0, nullptr, dhin, M, M, st);
  if (rc) return rc;
  { ActT maps[2] = {make_act(W<bf16>(s, lo.w_hin), M, T, B), make_act(dhb, HU, T, B)}; rc = wgrad(maps, 2); if (rc) return rc; }
  colsum_k<bf16><<<64, 128, 0, st>>>(dhb, N, HU, HU, d_grads + lo.p_db); t2_count_launch();
  // ---- proj2 (BN, linear) ----
  float* bsum = W<float>(s, lo.w_bsum);
  bf16* dY2b = W<bf16>(s, lo.w_dY2b);
  T2_CHECK_CUDA(cudaMemsetAsync(bsum, 0, 2LL * KC * sizeof(float), st));
  bn_bwd_stats_k<float, float><<<64, 128, 0, st>>>(dhin, M, W<float>(s, lo.w_Y2), M, 0, W<float>(s, lo.w_st2), M, bsum, N, M); t2_count_launch();
  bn_bwd_apply_k<float, float><<<g1(N * M), 256, 0, st>>>(dhin, M, W<float>(s, lo.w_Y2), M, 0, W<float>(s, lo.w_st2), M, bsum, d_params + lo.proj2.p_g, dY2b, 128,
                                                          d_grads + lo.proj2.p_g, d_grads + lo.proj2.p_be, N, M, 0); t2_count_launch();
  int sh[16];
  bf16* d2 = W<bf16>(s, lo.w_d2);
  for (int j = 0; j < lo.PK; ++j) sh[j] = -conv_shift(lo.PK, j);
  rc = gemm(dY2b, lo.proj2.coutp, 128, 0, T, B, s.pk + lo.proj2.k_wT, (PJc + 127) / 128 * 128, lo.PK * lo.proj2.coutp, lo.PK, sh, 256, nullptr, 0, d2, nullptr, PJc, PJc, st);
  if (rc) return rc;
  { ActT maps[2] = {make_act(W<bf16>(s, lo.w_X1), PJc, T, B), make_act(dY2b, M, T, B, 1, 128)}; rc = wgrad(maps, 2); if (rc) return rc; }
  colsum_k<bf16><<<64, 128, 0, st>>>(dY2b, N, M, 128, d_grads + lo.proj2.p_b); t2_count_launch();
  // ---- proj1 (ReLU, BN) ----
  bf16* d1 = W<bf16>(s, lo.w_d1);
  T2_CHECK_CUDA(cudaMemsetAsync(bsum, 0, 2LL * KC * sizeof(float), st));
  bn_bwd_stats_k<bf16, bf16><<<64, 256, 0, st>>>(d2, PJc, W<bf16>(s, lo.w_Y1), PJc, 0, W<float>(s, lo.w_st1), PJc, bsum, N, PJc); t2_count_launch();
  bn_bwd_apply_k<bf16, bf16><<<g1(N * PJc), 256, 0, st>>>(d2, PJc, W<bf16>(s, lo.w_Y1), PJc, 0, W<float>(s, lo.w_st1), PJc, bsum, d_params + lo.proj1.p_g, d1, PJc,
                                                          d_grads + lo.proj1.p_g, d_grads + lo.proj1.p_be, N, PJc, 1); t2_count_launch();
  bf16* dP = W<bf16>(s, lo.w_dP);
  rc = gemm(d1, PJc, PJc, 0, T, B, s.pk + lo.proj1.k_wT, KC, lo.PK * lo.proj1.coutp, lo.PK, sh, 256, nullptr, 0, dP, nullptr, KC, KC, st);
  if (rc) return rc;
  { ActT maps[2] = {make_act(W<bf16>(s, lo.w_P), KC, T, B), make_act(d1, PJc, T, B)}; rc = wgrad(maps, 2); if (rc) return rc; }
  colsum_k<bf16><<<64, 256, 0, st>>>(d1, N, PJc, PJc, d_grads + lo.proj1.p_b); t2_count_launch();
  // ---- max-pool, conv bank ----
  bf16* dbank = W<bf16>(s, lo.w_dbank);
  maxpool_bwd_k<<<g1(N * KC), 256, 0, st>>>(W<bf16>(s, lo.w_Xb), dP, dbank, N, T, KC); t2_count_launch();
  T2_CHECK_CUDA(cudaMemsetAsync(bsum, 0, 2LL * KC * sizeof(float), st));
  bf16* dpre = dP;                              // pre-activation gradients of the bank reuse the (consumed) dP buffer
  for (int k = 1; k <= lo.K; ++k) {
    const CConv& L = lo.bank[k - 1];
    const int c0 = (k - 1) * lo.CC;
    bn_bwd_stats_k<bf16, bf16><<<64, 128, 0, st>>>(dbank, KC, W<bf16>(s, lo.w_Y), KC, c0, W<float>(s, lo.w_stb), KC, bsum, N, lo.CC); t2_count_launch();
    bn_bwd_apply_k<bf16, bf16><<<g1(N * lo.CC), 256, 0, st>>>(dbank, KC, W<bf16>(s, lo.w_Y), KC, c0, W<float>(s, lo.w_stb), KC, bsum, d_params + L.p_g, dpre, KC,
                                                              d_grads + L.p_g, d_grads + L.p_be, N, lo.CC, 1); t2_count_launch();
    colsum_k<bf16><<<64, 128, 0, st>>>(dpre + c0, N, lo.CC, KC, d_grads + L.p_b); t2_count_launch();
  }
  { ActT maps[2] = {make_act(W<bf16>(s, lo.w_x0), M, T, B), make_act(dpre, KC, T, B)}; rc = wgrad(maps, 2); if (rc) return rc; }
  for (int g = 0; g < 3; ++g) {
    float* dx = W<float>(s, lo.w_dx0[g]);
    if (g >= lo.n_grp) { T2_CHECK_CUDA(cudaMemsetAsync(dx, 0, N * 128 * sizeof(float), st)); continue; }
    int shifts[16], k0s[16], n = 0;
    for (int l = lo.grp_first[g]; l < lo.grp_first[g + 1]; ++l)
      for (int j = 0; j < lo.bank[l].k; ++j, ++n) { shifts[n] = -conv_shift(lo.bank[l].k, j); k0s[n] = l * lo.CC; }
    rc = gemm(dpre, lo.CC, KC, 0, T, B, s.pk + lo.k_bankT[g], 128, n * lo.CC, n, shifts, 128, nullptr, 0, nullptr, dx, 128, M, st, k0s, KC);
    if (rc) return rc;
  }
  dmel_k<<<g1(N * M), 256, 0, st>>>(W<float>(s, lo.w_dx0[0]), W<float>(s, lo.w_dx0[1]), W<float>(s, lo.w_dx0[2]), dhin, d_mel_grad, N, M); t2_count_launch();
  if (lo.n_reg > 0 && lo.c.reg_weight != 0.f) {
    reg_grad_k<<<dim3(8, lo.n_reg), 256, 0, st>>>(d_params, d_grads, W<long long>(s, lo.w_regtab), lo.c.reg_weight); t2_count_launch();
  }
  T2_CHECK_CUDA(cudaGetLastError());
  return T2_OK;
}

extern "C" int t2_cbhg_workspace_tensor(const t2_cbhg_config_t* cfg, void* d_workspace, const char* name, void** ptr, long long* count) {
  CL lo;
  int rc = build(cfg, lo, nullptr);
  if (rc) return rc;
  T2_REQUIRE(d_workspace && name && ptr, T2_ERR_INVALID_ARG, "cbhg_workspace_tensor: null pointer");
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  const std::string n(name);
  long long off = -1, cnt = 0;
  if (n == "linear_outputs") { off = lo.w_lin; cnt = lo.N * lo.NFP; }           // fp32 [B][T][num_freq rounded up to 8] (row pitch!)
  else if (n == "rnn_outputs") { off = lo.w_out; cnt = lo.N * 2 * lo.RU; }     // bf16 [B][T][2 RU]
  else if (n == "highway_input") { off = lo.w_hin; cnt = lo.N * lo.M; }        // bf16 [B][T][M]
  else if (n == "bank_outputs") { off = lo.w_Xb; cnt = lo.N * lo.KC; }         // bf16 [B][T][K CC] (after batch norm)
  T2_REQUIRE(off >= 0, T2_ERR_INVALID_ARG, "cbhg_workspace_tensor: unknown tensor '%s'", name);
  *ptr = ws + off;
  if (count) *count = cnt;
  return T2_OK;
}
