/* t2b200.h — C-ABI of libt2b200.so, the sm_100a compute library behind the Tacotron-2 hot paths.
 *
 * The reference (Rayhane-mamah/Tacotron-2) has NO FFI / plugin boundary: its hot paths are Python methods
 * that build TensorFlow-1 graph nodes. This header therefore DEFINES the boundary; each entry point names the
 * reference function (file:line under the reference tree) whose arithmetic it replaces.
 *
 * Conventions
 *   - every function returns 0 (T2_OK) or a negative T2_ERR_* code; t2_last_error() returns a thread-local
 *     message for the last failure on the calling thread;
 *   - all pointers named d_* / documented "device" are DEVICE pointers owned by the caller; the library
 *     never allocates persistent device memory (the *_sizes queries say how much the caller must provide);
 *   - all work is enqueued on the cudaStream_t passed as `void* stream`; no host synchronisation unless
 *     documented; entry points are re-entrant per stream;
 *   - no torch / C++ types cross this boundary.
 */
#ifndef T2B200_H_
#define T2B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define T2B200_ABI_VERSION 2

#define T2_OK 0
#define T2_ERR_INVALID_ARG (-1)
#define T2_ERR_UNSUPPORTED_SHAPE (-2)
#define T2_ERR_CUDA (-3)
#define T2_ERR_NCCL (-4)

const char* t2_last_error(void);
int t2_abi_version(void);
/* sizeof() of a POD struct of this header by name ("t2_wn_config_t", "t2_wn_sizes_t", "t2_taco_config_t", "t2_cbhg_config_t",
 * "t2_audio_config_t"), -1 for an unknown name: a binding asserts that its mirror of the struct matches the library it loaded */
int t2_struct_size(const char* name);
/* kernels launched (or captured) by this library so far in this process */
long long t2_launch_count(void);

/* ---- engine-level test hooks (tests/test_gemm_engine.py) --------------------------------------------- */
/* bf16 dilated-conv-as-GEMM on the tcgen05 engine: out[b,t,n] = act(sum_s sum_k a[b,t+shift_s,k] w[n,s*Kp+k] + bias[n])
 * (Kp = C rounded up to 64). Replaces tf.layers.Conv1D as used by wavenet_vocoder/models/modules.py:206-224,320. */
int t2_dbg_conv_gemm(const void* d_a, int B, int T, int C, int ld, const int* shifts, int nshift,
                     const void* d_w, int N, int BN, const float* d_bias, int relu, void* d_out_bf16,
                     float* d_out_f32, void* stream);
/* debug: non-NULL => every GEMM CTA records 16 int64 stamps (clock64 phases, %globaltimer at entry / after the dependent-launch
 * wait / exit, SM id) at d_buf[(launch_offset + cta)*16 + slot], consecutive launches append; NULL disables */
int t2_dbg_set_timing_buffer(long long* d_buf);
/* weight-gradient GEMM: out[m,n] = scale * sum_{b,t} a[b,t+shift_a,m] * bm[b,t,n]  (fp32 [Ca,Cb]); synchronises. */
int t2_dbg_wgrad(const void* d_a, int Ca, const void* d_bm, int Cb, int B, int T, int shift_a, float scale,
                 float* d_out, void* stream);


/* ---- WaveNet vocoder: teacher-forced training path ---------------------------------------------------------
 * Replaces wavenet_vocoder/models/wavenet.py:650-721 (WaveNet.step), :476-519 (add_loss), the layers of
 * wavenet_vocoder/models/modules.py:184-521,539-654,736-817 and wavenet_vocoder/models/mixture.py:18-74.
 * Field names follow the reference's hparams.py:187-228. */
typedef struct {
  int layers, stacks, residual_channels, gate_channels, skip_out_channels, kernel_size;
  int cin_channels;          /* 80 (num_mels) or 0 = no local conditioning */
  int out_channels;          /* 256 (mu-law softmax), 3*nr_mix (MoL) or 2 (single Gaussian: mean, log-scale) */
  int quantize_channels;     /* 256 or 65536 */
  int input_type;            /* 0 'raw', 1 'mulaw', 2 'mulaw-quantize' */
  int legacy, residual_legacy;
  int upsample_type;         /* 0 'SubPixel', 1 '2D' (ConvTranspose2D) */
  int n_upsample;
  int upsample_scales[4];
  int freq_axis_kernel_size;
  float dropout;             /* wavenet_dropout */
  float log_scale_min;
  int B, T, Tc;              /* per-GPU batch, samples per item, conditioning frames per item */
  int c_pre_upsampled;       /* 1: conditioning is given at sample rate [B, T, cin] fp32 (skip upsample net) */
  float log_scale_min_gauss; /* single-Gaussian head (out_channels == 2): clamp of the predicted log-scale (hparams.py:196) */
  int cdf_loss;              /* Gaussian head: 1 = log(CDF+ - CDF-) loss, 0 = log-density (gaussian.py:18-33) */
  int split_bf16;            /* 1 = "fp32-class" forward: activations and weights travel as bf16 hi + lo pairs (3 tensor-core products per
                              * contraction, fp32 accumulate; ~2^-17 relative operand error instead of 2^-9). Forward / loss only, dropout 0:
                              * the parity mode that shows the bf16-mode deviation from the reference's fp32 graph is storage rounding. */
} t2_wn_config_t;

typedef struct {
  long long n_params;        /* fp32 master parameters (TF variable layouts, concatenated) */
  long long packed_bytes;    /* bf16 GEMM-operand copies + derived fp32 biases */
  long long workspace_bytes; /* activations saved for backward, gradients of activations, tables */
  int n_tensors;
} t2_wn_sizes_t;

int t2_wn_sizes(const t2_wn_config_t* cfg, t2_wn_sizes_t* out);
/* i-th parameter tensor: TF-style name, offset into the flat buffer, shape */
int t2_wn_param_info(const t2_wn_config_t* cfg, int i, char* name, int name_cap, long long* offset, int* ndim,
                     int* shape4);
/* one-time: zero the packed buffer / workspace and upload the static job tables (synchronises the stream) */
int t2_wn_init(const t2_wn_config_t* cfg, void* d_packed, void* d_workspace, void* stream);
/* fp32 masters -> bf16 operand layouts (run after every optimizer step) */
int t2_wn_pack_weights(const t2_wn_config_t* cfg, const float* d_params, void* d_packed, void* d_workspace,
                       void* stream);
/* forward + loss. d_x: int32 [B,T] mu-law indices (input_type 2) or fp32 [B,T] samples; d_c: fp32 [B,cin,Tc]
 * (or [B,T,cin] when c_pre_upsampled); d_targets: int32 / fp32 [B,T]; d_lengths int32 [B].
 * d_loss: fp32[2] = {sum of masked losses, normaliser} (loss = [0]/[1]); d_logits: optional fp32 [B,T,ldo]
 * (ldo = 256, or 32 for MoL). save_for_backward=0 skips the backward stashes. Dropout masks are a pure function of
 * (seed + *d_step, layer, position, channel); d_step (device u64, nullable) lets a replayed CUDA graph advance. */
int t2_wn_forward(const t2_wn_config_t* cfg, const float* d_params, const void* d_packed, void* d_workspace,
                  const void* d_x, const float* d_c, const void* d_targets, const int* d_lengths, float* d_loss,
                  float* d_logits, int save_for_backward, unsigned long long seed,
                  const unsigned long long* d_step, void* stream);
/* backward of the last t2_wn_forward(save_for_backward=1): writes all parameter gradients (d loss / d theta) */
int t2_wn_backward(const t2_wn_config_t* cfg, const float* d_params, const void* d_packed, void* d_workspace,
                   const void* d_x, const float* d_c, float* d_grads, unsigned long long seed,
                   const unsigned long long* d_step, void* stream);
/* Phased backward for data-parallel training (wavenet.py:561-593: tower gradients are averaged after backward). The weight
 * gradients of the residual stack come from `n_groups` launches over layer groups whose parameters are CONTIGUOUS ranges of the
 * flat gradient buffer, so the caller can all-reduce group g while group g+1 computes. phase -1: everything (== t2_wn_backward);
 * 0: data-gradient chain, head, conditioning tails; 1 + g: weight gradients of layers [g*L/n, (g+1)*L/n); 100: join of the
 * library's side stream (call after the last group, before reading gradients of the head / upsampling / first-conv tensors). */
int t2_wn_backward_phased(const t2_wn_config_t* cfg, const float* d_params, const void* d_packed, void* d_workspace,
                          const void* d_x, const float* d_c, float* d_grads, unsigned long long seed,
                          const unsigned long long* d_step, int phase, int n_groups, void* stream);
/* measurement hook for bench.py's roofline leg: average device time (CUDA events on `stream`) of `reps` launches of
 * one per-layer GEMM over the state left in the workspace by the last forward/backward. which: 0 gate GEMM, 1 out
 * GEMM, 2 dz + gate-backward GEMM, 3 dx GEMM. Re-running 1 / 3 rewrites x[l+1] / dx[l] with identical values.
 * Synchronises. */
int t2_wn_time_kernel(const t2_wn_config_t* cfg, const float* d_params, const void* d_packed, void* d_workspace,
                      int which, int layer, int reps, float* ms_per_launch, void* stream);
/* debug / test access to workspace tensors by name ("x", "z", "c_up", "h1", "dg", ...): returns device pointer,
 * element count and element size */
int t2_wn_workspace_tensor(const t2_wn_config_t* cfg, void* d_workspace, const char* name, void** ptr,
                           long long* count, int* elem_bytes);


/* ---- WaveNet vocoder: Fast-WaveNet autoregressive synthesis ---------------------------------------------------
 * Replaces WaveNet.incremental (wavenet_vocoder/models/wavenet.py:724-911), CausalConv1D incremental path
 * (modules.py:273-303) and sample_from_discretized_mix_logistic (mixture.py:76-107). cfg->B = synthesis batch,
 * cfg->T = samples to generate. cluster_size in {1,2,4,8,16}: CTAs per thread-block cluster sharing one batch group. */
int t2_wn_ar_sizes(const t2_wn_config_t* cfg, int cluster_size, long long* packed_bytes, long long* workspace_bytes);
/* fp32 masters -> slice-major bf16 synthesis weights + ring tables; run once per checkpoint; synchronises */
int t2_wn_ar_pack(const t2_wn_config_t* cfg, int cluster_size, const float* d_params, void* d_packed_ar,
                  void* d_workspace, void* stream);
/* d_c: fp32 [B,cin,Tc] (feeder-normalised mels); d_initial: int32[B] (mu-law index, 127 = silence) or fp32[B];
 * d_test_inputs: NULL or int32/fp32 [B,T] teacher-forcing inputs (the reference's wavenet_synth_debug path);
 * d_u_a / d_u_b: NULL (on-device counter RNG from `seed`) or injected uniforms in (0,1): MoL d_u_a [B,T,nr_mix] mixture
 * selection and d_u_b [B,T] logistic draw; mu-law d_u_a [B,T]. d_out_samples: int32 / fp32 [B,T];
 * d_out_raw: NULL or fp32 [B,T,out_channels] network outputs (what the reference collects in tower_y_hat_eval). */
int t2_wn_ar_generate(const t2_wn_config_t* cfg, int cluster_size, const float* d_params, const void* d_packed_ar,
                      void* d_workspace, const float* d_c, const void* d_initial, const void* d_test_inputs,
                      const float* d_u_a, const float* d_u_b, unsigned long long seed, void* d_out_samples,
                      float* d_out_raw, void* stream);


/* ---- Tacotron-2 mel predictor: training graph (teacher forcing, outputs_per_step = 1, predict_linear = False) ----
 * Replaces tacotron/models/tacotron.py:104-200,315-354, tacotron/models/modules.py:81-455,
 * tacotron/models/attention.py:38-226 and Architecture_wrappers.py:169-213. Names follow hparams.py:121-176,238-283. */
typedef struct {
  int B, T_in, T_out;
  int n_symbols, num_mels, embedding_dim;
  int enc_conv_layers, enc_conv_kernel, enc_conv_channels, encoder_lstm_units;
  int attention_dim, attention_filters, attention_kernel;
  int prenet1, prenet2, decoder_lstm_units;
  int postnet_layers, postnet_kernel, postnet_channels;
  int clip_outputs;
  float dropout_rate;        /* tacotron_dropout_rate (conv blocks in training; prenet always) */
  float zoneout_rate;        /* tacotron_zoneout_rate */
  float reg_weight;          /* tacotron_reg_weight */
  float max_abs_value, lower_bound_decay;
  int split_bf16;            /* 1 = "fp32-class" convolution stacks: the embedding, the encoder conv blocks + the BiLSTM input projection and
                              * the postnet conv blocks + projection run on bf16 hi + lo operand pairs with fp32 pre-batch-norm activations
                              * (forward / losses only). The recurrences (LSTMs, attention) keep bf16 operands / fp32 state. */
  int mask_decoder;          /* 1 = masked losses (tacotron/models/modules.py:412-455): MSE terms over the frames t < targets_lengths[b]
                              * (sum / count_nonzero of the mask), stop-token loss = weighted sigmoid CE over the same frames divided by the
                              * number of NON-ZERO masked terms; the lengths come from t2_taco_set_target_lengths */
  float cross_entropy_pos_weight;  /* pos_weight of tf.nn.weighted_cross_entropy_with_logits (masked stop-token loss only) */
} t2_taco_config_t;

int t2_taco_sizes(const t2_taco_config_t* cfg, long long* n_params, long long* packed_bytes, long long* workspace_bytes,
                  int* n_tensors);
int t2_taco_param_info(const t2_taco_config_t* cfg, int i, char* name, int name_cap, long long* offset, int* ndim,
                       int* shape4, int* trainable);
int t2_taco_init(const t2_taco_config_t* cfg, void* d_packed, void* d_workspace, void* stream);
int t2_taco_pack_weights(const t2_taco_config_t* cfg, const float* d_params, void* d_packed, void* d_workspace, void* stream);
/* forward + losses. d_inputs int32 [B,T_in] (0-padded character ids), d_input_lengths int32 [B], d_mel_targets fp32
 * [B,T_out,num_mels] (padded with -max_abs_value), d_stop_targets fp32 [B,T_out]. d_loss fp32[4] = {before MSE, after
 * MSE, stop-token CE, L2 regularisation}; total = sum. training=1: batch-norm batch statistics (moving stats in
 * d_params are updated), conv dropout, stochastic zoneout. */
int t2_taco_forward(const t2_taco_config_t* cfg, float* d_params, const void* d_packed, void* d_workspace,
                    const int* d_inputs, const int* d_input_lengths, const float* d_mel_targets,
                    const float* d_stop_targets, float* d_loss, int training, unsigned long long seed,
                    const unsigned long long* d_step, void* stream);
/* mask_decoder = 1: copies the B target lengths (device int32) into the workspace; call before t2_taco_forward (stream-ordered,
 * capturable). Replaces the `targets_lengths` placeholder of tacotron/models/tacotron.py:28 / tacotron/feeder.py:207. */
int t2_taco_set_target_lengths(const t2_taco_config_t* cfg, void* d_workspace, const int* d_target_lengths, void* stream);
/* backward of the last t2_taco_forward(training=1): d(total loss)/d(theta) into the flat gradient buffer (non-trainable
 * batch-norm moving statistics get 0) */
int t2_taco_backward(const t2_taco_config_t* cfg, const float* d_params, const void* d_packed, void* d_workspace,
                     const int* d_inputs, const int* d_input_lengths, const float* d_mel_targets,
                     const float* d_stop_targets, float* d_grads, unsigned long long seed,
                     const unsigned long long* d_step, void* stream);
/* same, plus an extra upstream gradient on the clipped mel_outputs (fp32 [B][T_out][num_mels], NULL = none): the CBHG head's
 * t2_cbhg_backward output (tacotron.py:203-219 hangs the post-processing net on mel_outputs) */
int t2_taco_backward_ex(const t2_taco_config_t* cfg, const float* d_params, const void* d_packed, void* d_workspace,
                     const int* d_inputs, const int* d_input_lengths, const float* d_mel_targets,
                     const float* d_stop_targets, float* d_grads, const float* d_mel_outputs_grad, unsigned long long seed,
                     const unsigned long long* d_step, void* stream);
/* Free-running synthesis (TacoTestHelper, tacotron/models/helpers.py:6-59; tacotron.py:150-200 with is_training = False:
 * inference batch-norm, deterministic zoneout blend, prenet dropout still on). cfg->T_out is max_iters (hparams.py:138).
 *   t2_taco_infer_begin   encoder, zero decoder state, go frame
 *   t2_taco_infer_steps   decoder steps [t_begin, t_end): each feeds back its own raw frame (helpers.py:56). The stop logit
 *                         of step t is workspace "projection_rows"[t][b][num_mels]; the CALLER applies the stop rule
 *                         (every row round(sigmoid) == 1, helpers.py:40-54; r = 1) between chunks - no host sync inside.
 *   t2_taco_infer_finish  clip, postnet, residual over the first T_used frames; workspace "decoder_output" /
 *                         "mel_outputs" are then COMPACT [B][T_used][num_mels], "stop_logits" [B][T_used]. */
int t2_taco_infer_begin(const t2_taco_config_t* cfg, float* d_params, const void* d_packed, void* d_workspace,
                        const int* d_inputs, const int* d_input_lengths, void* stream);
int t2_taco_infer_steps(const t2_taco_config_t* cfg, float* d_params, const void* d_packed, void* d_workspace,
                        const int* d_input_lengths, int t_begin, int t_end, unsigned long long seed, void* stream);
int t2_taco_infer_finish(const t2_taco_config_t* cfg, float* d_params, const void* d_packed, void* d_workspace, int T_used,
                         void* stream);
/* tools only: clock64() phase stamps of the attention kernels into a device buffer of 32 int64 (NULL turns them off) */
/* Counter-hash RNG behind the in-kernel dropout / zoneout masks (SURVEY.md §7 "stochastic ops in parity"):
 * d_out[i] = U[0,1) drawn for element (first_index + i) of hash stream `stream_id` under `seed` (= the seed passed to
 * t2_taco_forward / t2_wn_forward plus the device step counter). Element kept / updated iff d_out[i] >= rate.
 * Tacotron stream ids: encoder conv dropout 10+i, prenet 20 / 21, postnet conv dropout 30+i (element = linear index of
 * the layer output), zoneout (c, h) = 2*s, 2*s+1 with s = 52 / 53 (encoder fw / bw), 54 / 55 (decoder LSTM 1 / 2) and
 * element = (t*B + b)*H + unit. Replaces tf.layers.dropout / tf.nn.dropout draws of tacotron/models/modules.py:133-134,249,389. */
int t2_rng_uniform_f32(unsigned long long seed, unsigned int stream_id, long long first_index, long long n, float* d_out,
                       void* stream);
int t2_dbg_att_stamps(long long* d_buf);
int t2_dbg_ar_stamps(long long* d_buf);    /* same for one layer pass of the AR synthesis kernel (16 int64) */
int t2_taco_workspace_tensor(const t2_taco_config_t* cfg, void* d_workspace, const char* name, void** ptr,
                             long long* count, int* elem_bytes);

/* ---- Tacotron: CBHG post-processing net + linear-spectrogram head (predict_linear = True, the reference default) ----
 * Replaces tacotron/models/tacotron.py:203-219 (CBHG_postnet, cbhg_linear_specs_projection, clip), :323-330 / the
 * MaskedLinearLoss of tacotron/models/modules.py:457-485, and modules.py:4-78 (HighwayNet, CBHG: conv bank, max-pool, projections,
 * highway layers, bidirectional GRU over the whole padded sequence). Field names follow hparams.py:64,162-169.
 * It is a separate engine chained behind t2_taco_forward: its input is the Tacotron workspace tensor "mel_outputs", its backward
 * returns d(linear loss + its regulariser)/d(mel_outputs), which t2_taco_backward_ex adds to the mel loss seed. Parameters, gradients
 * and Adam moments are sub-ranges of the caller's flat buffers (one optimizer step / one global-norm clip over both engines). */
typedef struct {
  int B, T;                  /* batch items, decoder steps (= mel frames, >= 2) */
  int num_mels;              /* 80 */
  int kernels;               /* cbhg_kernels: convolution bank sizes 1..kernels (<= 8) */
  int conv_channels;         /* cbhg_conv_channels (128) */
  int pool_size;             /* cbhg_pool_size (2) */
  int projection;            /* cbhg_projection (256); the second projection maps back to num_mels */
  int projection_kernel_size;/* cbhg_projection_kernel_size (3) */
  int highwaynet_layers;     /* cbhg_highwaynet_layers (4) */
  int highway_units;         /* cbhg_highway_units (128) */
  int rnn_units;             /* cbhg_rnn_units (128): GRU units per direction */
  int num_freq;              /* 1025 */
  int n_priority_freq;       /* int(2000 / (sample_rate / 2) * num_freq): bins carrying the second half of the L1 weight */
  int clip_outputs, mask_decoder;
  float max_abs_value, lower_bound_decay, reg_weight;
} t2_cbhg_config_t;
int t2_cbhg_sizes(const t2_cbhg_config_t* cfg, long long* n_params, long long* packed_bytes, long long* workspace_bytes, int* n_tensors);
int t2_cbhg_param_info(const t2_cbhg_config_t* cfg, int i, char* name, int name_cap, long long* offset, int* ndim, int* shape4,
                       int* trainable);
int t2_cbhg_init(const t2_cbhg_config_t* cfg, void* d_packed, void* d_workspace, void* stream);          /* synchronises */
int t2_cbhg_pack_weights(const t2_cbhg_config_t* cfg, const float* d_params, void* d_packed, void* d_workspace, void* stream);
/* mask_decoder = 1: the B target lengths (device int32), as t2_taco_set_target_lengths */
int t2_cbhg_set_target_lengths(const t2_cbhg_config_t* cfg, void* d_workspace, const int* d_target_lengths, void* stream);
/* d_mel: fp32 [B][T][num_mels] (the clipped mel_outputs); d_linear_targets fp32 [B][T][num_freq] or NULL (inference);
 * d_loss[0] = linear loss, d_loss[1] = reg_weight * sum l2_loss(CBHG kernels) (NULL allowed). training = 1: batch statistics
 * (+ moving-average update), stashes for the backward pass. Linear outputs: workspace tensor "linear_outputs", fp32 rows of
 * pitch (num_freq rounded up to a multiple of 8), the first num_freq columns valid. */
int t2_cbhg_forward(const t2_cbhg_config_t* cfg, float* d_params, const void* d_packed, void* d_workspace, const float* d_mel,
                    const float* d_linear_targets, float* d_loss, int training, void* stream);
/* backward of the last t2_cbhg_forward(training = 1 with targets): gradients into d_grads (this engine's range of the flat
 * buffer; overwritten), d(loss)/d(mel_outputs) into d_mel_grad fp32 [B][T][num_mels] */
int t2_cbhg_backward(const t2_cbhg_config_t* cfg, const float* d_params, const void* d_packed, void* d_workspace, const float* d_mel,
                     float* d_grads, float* d_mel_grad, void* stream);
int t2_cbhg_workspace_tensor(const t2_cbhg_config_t* cfg, void* d_workspace, const char* name, void** ptr, long long* count);

/* ---- optimizer: tf.train.AdamOptimizer + per-tensor clip_by_norm/clip_by_value + EMA ------------------------
 * Replaces wavenet.py:586-613 (and tacotron.py:429-437 with global_norm_clip > 0).
 * d_offsets: int64 [n_tensors + 1] element offsets of the tensors inside the flat buffers.
 * grad_scale multiplies every gradient first (1/world_size after an NCCL sum all-reduce).
 * max_norm <= 0 disables per-tensor norm clipping; max_value <= 0 disables value clipping;
 * global_norm_clip > 0 applies tf.clip_by_global_norm instead. d_ema may be NULL. d_scratch: fp32 [n_tensors+1]. */
int t2_adam_step(float* d_params, const float* d_grads, float* d_m, float* d_v, float* d_ema,
                 const long long* d_offsets, int n_tensors, long long n_total, float lr, float beta1, float beta2,
                 float eps, int step, float grad_scale, float max_norm, float max_value, float global_norm_clip,
                 float ema_decay, float* d_scratch, void* stream);

/* ---- audio front-end ----------------------------------------------------------------------------------------
 * Replaces datasets/audio.py:22-25,61-77,178-182,225-270 and wavenet_vocoder/util.py:30-129. Field names follow
 * hparams.py:63-111. */
typedef struct {
  int sample_rate, n_fft, hop_size, win_size, num_mels;
  float fmin, fmax;
  float magnitude_power;
  float min_level_db, ref_level_db, max_abs_value;
  int symmetric_mels, allow_clipping_in_normalization, signal_normalization;
} t2_audio_config_t;

int t2_stft_mel_plan_bytes(const t2_audio_config_t* cfg, long long* bytes);
/* builds twiddles, the periodic Hann window and the sparse Slaney mel filterbank (librosa.filters.mel restated in
 * fp64 on the host) into d_plan; synchronises the stream */
int t2_stft_mel_plan_init(const t2_audio_config_t* cfg, void* d_plan, void* stream);
int t2_stft_mel_frames(const t2_audio_config_t* cfg, int n_samples);
/* melspectrogram (and optionally linearspectrogram) of B clips of n_samples fp32 samples.
 * sample fed to the STFT = gain * (x[n] - preemphasis * x[n-1]); preemphasis = 0, gain = 1 is the plain
 * datasets/audio.py:melspectrogram. d_mel: fp32 [B][frames][num_mels] (time_major=1, the layout the preprocessor
 * saves, datasets/preprocessor.py:158) or [B][num_mels][frames] (time_major=0, what audio.melspectrogram returns);
 * d_linear: NULL or fp32 [B][frames][n_fft/2+1] / [B][n_fft/2+1][frames]. */
int t2_stft_mel_f32(const t2_audio_config_t* cfg, const void* d_plan, const float* d_wav, int B, int n_samples,
                    float preemphasis, float gain, float* d_mel, float* d_linear, int time_major, void* stream);
/* dense Slaney mel filterbank the fused kernel uses (librosa.filters.mel as called by datasets/audio.py:243-246): HOST double
 * [num_mels][n_fft/2 + 1]; its pseudo-inverse is the reference's _mel_to_linear (audio.py:231-241) */
int t2_mel_basis_f64(const t2_audio_config_t* cfg, double* h_basis);
/* Griffin-Lim phase reconstruction on the GPU: replaces datasets/audio.py:151-161 (_griffin_lim: librosa istft / stft iterations)
 * and :163-176 (the TF-graph variant). d_mag: fp32 [B][frames][n_fft/2+1] magnitudes (already raised to hparams.power);
 * d_phase_io: optional float2 [B][frames][bins] unit phases (in: initial phases, out: final) - NULL draws exp(2 pi i u) from the
 * counter hash under `seed` (the reference draws np.random.rand); iters = hparams.griffin_lim_iters re-estimation rounds (iters + 1
 * inverse transforms); d_wav: fp32 [B][hop * (frames - 1)] (librosa.istft length, centre-trimmed). Workspace: t2_griffin_lim_bytes. */
int t2_griffin_lim_bytes(const t2_audio_config_t* cfg, int B, int frames, long long* bytes);
int t2_griffin_lim_f32(const t2_audio_config_t* cfg, const void* d_plan, const float* d_mag, float* d_phase_io, int B, int frames,
                       int iters, unsigned long long seed, void* d_workspace, float* d_wav, void* stream);
int t2_preemphasis_f32(const float* d_x, float* d_y, int B, int n_samples, float k, void* stream);
/* mu-law (mu forced to 255 like util.py:48,67,99,127); quantise truncates toward zero */
int t2_mulaw_quantize_f32_i32(const float* d_in, int* d_out, long long n, void* stream);
int t2_inv_mulaw_quantize_i32_f32(const int* d_in, float* d_out, long long n, void* stream);
int t2_mulaw_f32(const float* d_in, float* d_out, long long n, void* stream);
int t2_inv_mulaw_f32(const float* d_in, float* d_out, long long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* T2B200_H_ */
