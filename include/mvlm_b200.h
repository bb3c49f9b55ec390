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
  const float* mid_scale; /* optional: v = relu(v*mid_scale+mid_shift) right after the bias */
  const float* mid_shift;
  int pool2;              /* 1: out_raw / out_post are the 2x2 max-pooled (half resolution) tensors */
  const void* res_up;     /* optional bf16 (n, h/2, w/2, up_cs): v += nearest-x2 up-sampled res_up (with res1/res2) */
  int up_cs, up_co;
} mvlm_conv_args;

int mvlm_conv2d_bf16(const mvlm_conv_args* args, void* stream);

/* Debug aid: when dev_buf (mvlm_debug_conv_profile_ints() int64, device) is non-NULL the following conv launches record per-CTA role
 * stall cycles there: [0] producer wait A-empty, [1] producer wait B-empty, [2] MMA wait operands,
 * [3] MMA wait accumulator-free, [4] MMA total, [5] epilogue wait accumulator-full, [6] epilogue total,
 * [7] producer total.  NULL switches it off. */
void mvlm_debug_conv_profile(long long* dev_buf);
/* number of int64 the buffer given to mvlm_debug_conv_profile must hold (role counters + CTA-0 tile timeline) */
int mvlm_debug_conv_profile_ints(void);
/* Debug experiment switch: 1 = tensor-pipe-only timing (no TMA loads; outputs are garbage), 0 = normal. */
void mvlm_debug_conv_mode(int mode);

/* fp32 OIHW (device) -> packed bf16 [cout_pad][kw][kh][cin_pad] (device), zero padded. */
int mvlm_pack_conv_weight(const float* w_oihw, int cout, int cin, int kh, int kw, int cout_pad,
                          int cin_pad, void* out_bf16, void* stream);

/* ------------------------------------------------------------------------- */
/* Scan loader (host, multi-threaded): Wavefront OBJ -> flat arrays.            */
/* Replaces vtkOBJReader in obj_to_actor, src/mvlm/utils/utils3d.py:16-21       */
/* (error "does not contain any points" :20-21).  Output vertices are the       */
/* unique (position, vt) pairs in ascending order, faces fan-triangulated.      */
/* n_threads <= 0: one thread per host core (at most 16).                       */
/* ------------------------------------------------------------------------- */
typedef struct mvlm_obj mvlm_obj;
int mvlm_obj_load(const char* path, int n_threads, mvlm_obj** out);
int mvlm_obj_counts(const mvlm_obj* obj, int* n_verts, int* n_tris, int* has_uv);
/* verts f32[n_verts*3], uvs f32[n_verts*2] (may be NULL), tris i32[n_tris*3]: host memory */
int mvlm_obj_copy(const mvlm_obj* obj, float* verts, float* uvs, int32_t* tris);
void mvlm_obj_free(mvlm_obj* obj);

/* Texture decode on the GPU (nvJPEG); replaces vtkJPEGReader in obj_to_actor, src/mvlm/utils/utils3d.py:26-36.  */
/* data/len: the compressed file in host memory; out_rgb_dev: device buffer u8[height][width][3], row 0 = top.    */
/* Thread-safe (one decoder state per calling thread); asynchronous on `stream` after the host-side parse.        */
int mvlm_jpeg_info(const uint8_t* data, size_t len, int* width, int* height);
int mvlm_jpeg_decode_rgb(const uint8_t* data, size_t len, uint8_t* out_rgb_dev, int width, int height, void* stream);

/* ------------------------------------------------------------------------- */
/* Stage 1: batched multi-view orthographic rasteriser.                       */
/* Replaces ObjVTKRenderer3D.render_3d_multi_rgb_geometry_depth               */
/*   src/mvlm/utils/render3d.py:114-177 (+ camera :53-59,:136,:150-152, depth  */
/*   encoder :73-77,:166-170, flip :177, /255 :191) and obj_to_actor's         */
/*   material src/mvlm/utils/utils3d.py:26-64.                                 */
/* rot: (V,9) float64 row-major R = Ry(ry)*Rx(rx)*Rz(rz) per view.             */
/* channel_mode: 0 RGB+depth, 1 geometry+depth, 2 RGB, 3 depth, 4 geometry.    */
/* tex: (Th,Tw,tex_channels) u8, tex_channels 3 (RGB) or 4 (RGBA, one 4-byte load  */
/* per texel).  workspace: mvlm_raster_workspace_bytes(...) bytes = the packed      */
/* depth|triangle-id keys of all pixels + the per-(view, vertex) window coordinates */
/* (each vertex is transformed once per view).  Any of the four outputs may be NULL.*/
/* ------------------------------------------------------------------------- */
size_t mvlm_raster_workspace_bytes(int n_views, int h, int w, int n_verts);
int mvlm_raster_multiview(const float* verts, int n_verts, const float* uvs, const int32_t* tris, int n_tris,
                          const uint8_t* tex, int tex_h, int tex_w, int tex_channels, const double* rot,
                          int n_views, int h, int w, int channel_mode, void* workspace, size_t workspace_bytes,
                          uint8_t* out_u8 /* (V,H,W,4) */, float* out_f32 /* (V,H,W,C) */,
                          int32_t* out_tri_id /* (V,H,W) */, float* out_depth /* (V,H,W) */,
                          void* stream);

/* ------------------------------------------------------------------------- */
/* Stage 2: stacked-hourglass heat-map CNN (MVLMModel),                       */
/*   src/mvlm/prediction/paulsenpredictor.py:364-432 driven by                 */
/*   predict_landmarks_from_images :167-217.                                   */
/* names/ptrs: the torch state_dict (fp32 device tensors, reference key names).*/
/* forward: exactly one of img_u8 (V,H,W,4 packed u8 from the rasteriser) or    */
/* img_f32 (V,H,W,cin fp32 in [0,1], the reference's image_stack layout).       */
/* out_heatmaps (V,L,H,W) fp32 and out_peaks (L,V,3) fp32 are each optional.    */
/* ------------------------------------------------------------------------- */
typedef struct mvlm_hourglass mvlm_hourglass;
size_t mvlm_hourglass_workspace_bytes(int n_landmarks, int cin, int n_views, int h, int w);
double mvlm_hourglass_flops_per_view(int n_landmarks, int cin, int h, int w);
/* numels[i] = element count of tensor i (or NULL: unchecked).  Like load_state_dict (paulsenpredictor.py:102,108)
 * a missing key or a tensor whose size does not fit this model's layer is an error (MVLM_E_INVALID), never a
 * silent misread. */
int mvlm_hourglass_create(const char* const* names, const void* const* ptrs, const long long* numels,
                          int n_entries, int n_landmarks, int cin, int n_views, int h, int w, void* workspace,
                          size_t workspace_bytes, mvlm_hourglass** out);
int mvlm_hourglass_forward(mvlm_hourglass* net, const uint8_t* img_u8, const float* img_f32,
                           float* out_heatmaps, float* out_peaks, void* stream);
/* Same result as mvlm_hourglass_forward; the ~155 launches are captured into a CUDA graph per distinct
 * (img, out) pointer tuple on first use and replayed afterwards (buffers must stay valid and unchanged). */
int mvlm_hourglass_forward_graph(mvlm_hourglass* net, const uint8_t* img_u8, const float* img_f32,
                                 float* out_heatmaps, float* out_peaks, void* stream);
/* Peak selection of the out_peaks of forward / forward_graph (PaulsenModel.selection_method, paulsenpredictor.py:55,
 * 112-158): 0 = "simple" (the fused arg-max), 1 = "moment" (centre of mass of the 31x31 window around it, :129-156).
 * "moment" stays on the fused path: the window's heat-map values are re-evaluated from the last layer's input by a
 * small kernel, so the (V,L,H,W) fp32 heat maps are not written to memory for it either. */
int mvlm_hourglass_set_selection_method(mvlm_hourglass* net, int method);
/* View-split hand-off (SURVEY.md 8b "mvlm_allgather_peaks", 8e): when one scan's views are split over ranks, each rank
 * runs the network on its block of views with the fused arg-max of the last convolution writing its keys
 * (u64 = ordered value << 32 | ~index, n_views x n_landmarks) STRAIGHT into this rank's slot of the all-gather buffer
 * keys[world][slot_views][n_landmarks]; the host layer issues ONE in-place NCCL all-gather on that buffer
 * (torch.distributed's communicator: mvlm_b200/sharding.py::allgather_keys) and ONE kernel turns the gathered keys of
 * ALL views into peaks (L, V, 3), mapping view v to (rank, row) by the contiguous split of sharding.split_views.
 * No padding copy, no per-rank scatter. */
int mvlm_hourglass_forward_keys(mvlm_hourglass* net, const uint8_t* img_u8, const float* img_f32, uint64_t* out_keys,
                                void* stream);
int mvlm_peaks_from_gathered_keys(const uint64_t* keys, int n_views, int n_landmarks, int w, int world, int slot_views,
                                  float* out_peaks /* (L, V, 3) */, void* stream);
int mvlm_hourglass_num_launches(const mvlm_hourglass* net);
/* number of dataflow segments of the plan (csrc/conv_flow.cuh): runs of layers executed by one persistent launch
 * over small view batches so that their intermediate tensors stay in L2; 0 = every layer is its own launch */
int mvlm_hourglass_num_segments(const mvlm_hourglass* net);
/* Debug aids: per-op mean milliseconds over `reps` passes (ms_out[num_launches], host memory; out_peaks (L,V,3)
 * device), optionally the conv kernel's per-role stall cycles (roles_out[num_launches*8] host doubles, layout of
 * mvlm_debug_conv_profile, mean over CTAs), and a one-line description of op `op`. */
int mvlm_debug_hourglass_profile(mvlm_hourglass* net, const uint8_t* img_u8, const float* img_f32,
                                 float* out_peaks, int reps, float* ms_out, double* roles_out, int trace_op,
                                 long long* trace_out /* 64 x 16 int64 timeline of CTA 0 of conv op trace_op, or NULL */,
                                 void* stream);
int mvlm_debug_hourglass_describe(const mvlm_hourglass* net, int op, char* buf, int buf_len);
/* layer-wise parity probes: "r3", "hg1", "sum_temp", "x10" -> NHWC bf16 tensor in the workspace */
int mvlm_hourglass_probe(const mvlm_hourglass* net, const char* name, const void** ptr, int* h, int* w, int* c);
void mvlm_hourglass_destroy(mvlm_hourglass* net);

/* ------------------------------------------------------------------------- */
/* Stage 3: heat-map peaks.  Replaces find_heat_map_maxima /                   */
/*   find_maxima_in_batch_of_heatmaps, paulsenpredictor.py:112-165.            */
/* method 0 = "simple", 1 = "moment".  heatmaps (V,L,H,W) fp32 -> (L,V,3).     */
/* ------------------------------------------------------------------------- */
int mvlm_heatmap_peaks(const float* heatmaps, int n_views, int n_landmarks, int h, int w, int method,
                       float* out_peaks, void* stream);

/* ------------------------------------------------------------------------- */
/* Stage 4: peaks -> view rays.  Replaces Estimator3D.estimate_landmark_lines  */
/*   src/mvlm/utils/estimator3d.py:31-90.  rot as in mvlm_raster_multiview.    */
/* ------------------------------------------------------------------------- */
int mvlm_rays_from_peaks(const float* peaks /* (L,V,3) */, const double* rot /* (V,9) */, int n_landmarks,
                         int n_views, int image_size, double* out_starts, double* out_ends, void* stream);

/* ------------------------------------------------------------------------- */
/* Stage 5a: filter + RANSAC + least squares.  Replaces                        */
/*   estimate_landmarks_from_lines estimator3d.py:158-183 (filters :140-155,   */
/*   RANSAC :92-137, LSQ utils3d.py:99-124).  mode 0 = quantile, 1 = absolute. */
/* draws: (L,H,8) uint32 seeded hypothesis table, line index = draw mod n.     */
/* ------------------------------------------------------------------------- */
size_t mvlm_consensus_workspace_bytes(int n_landmarks, int n_views, int n_hyp);
int mvlm_consensus(const float* peaks, const double* starts, const double* ends, int n_landmarks,
                   int n_views, int mode, double threshold_quantile, float threshold_absolute,
                   const uint32_t* draws, int n_hyp, double dist_thres, void* workspace,
                   size_t workspace_bytes, double* out_landmarks /* (L,3) */, double* out_errors /* (L) */,
                   int32_t* out_nlines /* (L) or NULL */, void* stream);

/* ------------------------------------------------------------------------- */
/* Stage 5b: snap to the mesh surface.  Replaces                               */
/*   Estimator3D.project_landmarks_to_surface estimator3d.py:252-285.          */
/* ------------------------------------------------------------------------- */
size_t mvlm_snap_workspace_bytes(int n_landmarks, int n_tris);
int mvlm_snap_to_mesh(const float* verts, const int32_t* tris, int n_tris, const double* landmarks,
                      int n_landmarks, void* workspace, size_t workspace_bytes, double* out /* (L,3) */,
                      int32_t* out_tri /* (L) or NULL */, void* stream);

/* Same stage with a spatial index, for large meshes or a mesh queried more than once (the reference     */
/* builds a vtkCellLocator inside every call, estimator3d.py:258-262).  mvlm_snap_grid_build fills `grid` */
/* (mvlm_snap_grid_bytes(n_tris) bytes of device memory, 16-byte aligned) with a uniform grid over the    */
/* triangle centroids, entirely on the device; mvlm_snap_grid_query returns exactly what                  */
/* mvlm_snap_to_mesh returns (same triangle, same point); landmarks many cells away from the surface are  */
/* handed to the full scan inside the same call.  out_stats (2L ints or NULL): point-triangle tests and   */
/* shells walked per landmark (-1 = full scan).                                                           */
size_t mvlm_snap_grid_bytes(int n_tris);
size_t mvlm_snap_grid_query_workspace_bytes(int n_landmarks, int n_tris);
int mvlm_snap_grid_build(const float* verts, const int32_t* tris, int n_tris, void* grid, size_t grid_bytes,
                         void* stream);
int mvlm_snap_grid_query(const float* verts, const int32_t* tris, int n_tris, const void* grid,
                         size_t grid_bytes, const double* landmarks, int n_landmarks, void* workspace,
                         size_t workspace_bytes, double* out /* (L,3) */, int32_t* out_tri /* (L) or NULL */,
                         int32_t* out_stats /* (L,2) or NULL */, void* stream);
/* debug: grid dims[3] + oversize-list length, cell edge + largest binned triangle radius (synchronises) */
int mvlm_debug_snap_grid_describe(const void* grid, int32_t* dims_nover /* [4] */, double* edge_tau /* [2] */,
                                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MVLM_B200_H */
