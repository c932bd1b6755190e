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

#define T2B200_ABI_VERSION 1

#define T2_OK 0
#define T2_ERR_INVALID_ARG (-1)
#define T2_ERR_UNSUPPORTED_SHAPE (-2)
#define T2_ERR_CUDA (-3)
#define T2_ERR_NCCL (-4)

const char* t2_last_error(void);
int t2_abi_version(void);

/* ---- engine-level test hooks (tests/test_gemm_engine.py) --------------------------------------------- */
/* bf16 dilated-conv-as-GEMM on the tcgen05 engine: out[b,t,n] = act(sum_s sum_k a[b,t+shift_s,k] w[n,s*Kp+k] + bias[n])
 * (Kp = C rounded up to 64). Replaces tf.layers.Conv1D as used by wavenet_vocoder/models/modules.py:206-224,320. */
int t2_dbg_conv_gemm(const void* d_a, int B, int T, int C, int ld, const int* shifts, int nshift,
                     const void* d_w, int N, int BN, const float* d_bias, int relu, void* d_out_bf16,
                     float* d_out_f32, void* stream);
/* weight-gradient GEMM: out[m,n] = scale * sum_{b,t} a[b,t+shift_a,m] * bm[b,t,n]  (fp32 [Ca,Cb]); synchronises. */
int t2_dbg_wgrad(const void* d_a, int Ca, const void* d_bm, int Cb, int B, int T, int shift_a, float scale,
                 float* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* T2B200_H_ */
