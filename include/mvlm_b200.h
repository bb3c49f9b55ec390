/* mvlm_b200 -- C-ABI of the B200-native multi-view landmarking hot path.
 *
 * Every entry point replaces one stage of the reference's
 * Pipeline.predict_one_file (src/mvlm/pipeline/general_pipeline.py:67-131);
 * the stage it replaces is cited at each declaration (paths relative to the
 * reference tree).  Conventions:
 *   - plain C symbols, device pointers unless the name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - return 0 on success, <0 on error (MVLM_E_*); mvlm_last_error() gives the
 *     message of the last failing call on the calling thread;
 *   - no hidden device allocation in stage calls: the caller owns workspaces
 *     sized by the matching *_workspace_bytes() function.  Only the
 *     *_create() functions allocate (and *_destroy() frees).
 */
#ifndef MVLM_B200_H
#define MVLM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVLM_OK 0
#define MVLM_E_INVALID (-1)
#define MVLM_E_CUDA (-2)
#define MVLM_E_UNSUPPORTED (-3)

const char* mvlm_last_error(void);
int mvlm_version(void);
/* Number of kernels this library launched since the last reset (bench.py's gpu_launches). */
long long mvlm_launch_count(int reset);

/* ------------------------------------------------------------------------- */
/* Stage 2 building block: one convolution of the stacked-hourglass CNN.      */
/* Replaces torch.nn.Conv2d (+BatchNorm2d/ReLU/residual/cat) call sites:      */
/*   src/mvlm/prediction/paulsenpredictor.py:251-273, :385-402, :404-432      */
/* Activations NHWC bf16; weights packed by mvlm_pack_conv_weight().          */
/* ------------------------------------------------------------------------- */
typedef struct mvlm_conv_args {
  const void* in;      /* bf16 NHWC, channel stride in_cs */
  int n, h, w, cin, in_cs;
  const void* wpacked; /* bf16 [cout_pad][kw][kh][cin] */
  int cout_pad, n_tile, kh, kw, y_off0, x_off0;
  const float* bias;
  const float* pre_scale;
  const float* pre_shift;
  void* out_pre;
  int pre_cs, pre_co;
  const void* res1;
  int res1_cs, res1_co;
  const void* res2;
  int res2_cs, res2_co;
  void* out_raw;
  int raw_cs, raw_co;
  const float* post_scale;
  const float* post_shift;
  void* out_post;
  int post_cs, post_co;
  float* out_f32;                  /* NCHW fp32 (n, cout_real, h*up_sy, w*up_sx) */
  unsigned long long* argmax_keys; /* [n*cout_real], see mvlm_peaks_from_keys */
  int cout_real;
  int up_sy, up_sx, up_py, up_px;
} mvlm_conv_args;

int mvlm_conv2d_bf16(const mvlm_conv_args* args, void* stream);

/* fp32 OIHW (device) -> packed bf16 [cout_pad][kw][kh][cin_pad] (device), zero padded. */
int mvlm_pack_conv_weight(const float* w_oihw, int cout, int cin, int kh, int kw, int cout_pad,
                          int cin_pad, void* out_bf16, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MVLM_B200_H */
