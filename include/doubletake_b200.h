/*
 * doubletake_b200 -- C ABI of the B200-native plane-sweep MVS depth engine.
 *
 * The reference (nianticlabs/doubletake) has no FFI: its "plugin API" for this path is Python duck-typing on the
 * nn.Module attributes of the LightningModule (model.cost_volume / model.cost_volume_net / model.depth_decoder,
 * swapped the way utils/model_utils.py:30-34 swaps in the "fast" cost volume).  This header is the C boundary a
 * maintainer binds underneath those attributes (ctypes stub in INTEGRATION.md).  Plain pointers and sizes only:
 * no torch types, no allocation inside, no exceptions across the ABI.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 unless stated; all tensors dense/contiguous in the stated layout
 *   - every entry point enqueues work on `stream` and returns immediately (0 = ok, <0 = error code below;
 *     dtb200_last_error() gives the message for the calling thread)
 *   - thread-safe per stream; process-global state is limited to the per-thread error string, the launch counter, per-device
 *     one-time kernel attributes (std::call_once) and the development switches of dtb200_debug_set
 */
#ifndef DOUBLETAKE_B200_H
#define DOUBLETAKE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* dtb200_stream_t; /* cudaStream_t */

#define DTB200_OK 0
#define DTB200_ERR_INVALID (-1)   /* bad argument / unsupported shape */
#define DTB200_ERR_CUDA (-2)      /* CUDA runtime error on launch */
#define DTB200_ERR_UNSUPPORTED (-3)

#define DTB200_ABI_VERSION 1
#define DTB200_MAX_VIEWS 16

int dtb200_abi_version(void);
/* development only: kernel-variant switches of the tensor-core conv pipeline (0 = normal operation); the timing
 * knock-outs (bits 8 and up, wrong results by design) are ignored unless DTB200_DEVELOPMENT=1 is in the environment */
int dtb200_debug_set(int flags);
/* development only: timeline of the split16 ("tch") conv kernels.  `device_pairs` = capacity x 2 uint64 in device memory, caller
 * initialises every pair to (UINT64_MAX, 0); each later tch conv launch takes the next pair (launch order, fixed at graph
 * capture) and stamps %globaltimer: earliest CTA start, latest CTA end.  NULL turns it off.  tools/graph_trace.py */
int dtb200_debug_trace(uint64_t* device_pairs, int32_t capacity);
const char* dtb200_last_error(void);
/* number of kernels this library has launched from the calling process (bench.py's gpu_launches claim) */
uint64_t dtb200_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * Layout helpers (boundary transposes; the reference hands NCHW tensors, the kernels gather channels-last)
 * ---------------------------------------------------------------------------------------------------------- */
int dtb200_nchw_to_nhwc(const float* src, float* dst, int n, int c, int h, int w, dtb200_stream_t stream);
int dtb200_nhwc_to_nchw(const float* src, float* dst, int n, int c, int h, int w, dtb200_stream_t stream);
/* math = TCH keeps activations in the "split16" layout: (N, H, W, 2, C) fp16 -- per pixel C values big = fp16(x) followed by
 * C values small = fp16((x - big) * 2048); 4C bytes per pixel like fp32, 22 significant bits, clamped to the fp16 range.
 * c must be a multiple of 8.  A split16 tensor occupies exactly the bytes of the (N, H, W, C) fp32 tensor. */
int dtb200_nchw_to_split16(const float* src, void* dst, int n, int c, int h, int w, dtb200_stream_t stream);
int dtb200_split16_to_nchw(const void* src, float* dst, int n, int c, int h, int w, dtb200_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Plane-sweep cost volumes.
 *
 * Replaces, in one launch per batch, the per-plane Python loop of
 *   CostVolumeManager.build_cost_volume + forward          (modules/cost_volume.py:219-363)        kind = DOT
 *   FeatureVolumeManager.build_cost_volume + forward       (modules/feature_volume.py:81-356)      kind = MLP
 *   FeatureMeshHintVolumeManager.build_cost_volume+forward (modules/mesh_hint_volume.py:84-439)    kind = MLP_HINT
 * i.e. BackprojectDepth/Project3D (utils/geometry_utils.py:55-93), F.grid_sample warp, dot product + mask,
 * metadata assembly (26K+20 channels, order of mesh_hint_volume.py:343-367), MLP 26K+20 -> 128 -> 128 -> 1,
 * hint MLP 3 -> 12 -> 12 -> 1, arg-max over planes -> lowest_cost depth, and the overall source-view mask.
 * ---------------------------------------------------------------------------------------------------------- */
#define DTB200_VOLUME_DOT 0
#define DTB200_VOLUME_MLP 1
#define DTB200_VOLUME_MLP_HINT 2

/* math modes: EXACT = fp32 FMA in the reference operation order (bit-reproducible, CUDA cores);
 * TC3X = tcgen05 tensor cores with the 3xTF32 split (fp32-class accuracy, ~2^-21 per product);
 * TCH  = tcgen05 kind::f16 with a 2-term fp16 split of both operands (big + small, three MMAs per product at twice the
 *        TF32 rate and half the operand bytes; the same ~2^-21 accuracy class) -- the default of the benchmark */
#define DTB200_MATH_EXACT 0
#define DTB200_MATH_TC3X 1
#define DTB200_MATH_TCH 2

typedef struct {
  int32_t kind;          /* DTB200_VOLUME_* */
  int32_t math;          /* DTB200_MATH_* (DOT ignores it) */
  int32_t batch, views, channels, height, width, planes; /* B K C H W D ; C must be 16 */
  /* inputs */
  const float* cur_feats;       /* (B,C,H,W)   NCHW, as the reference passes it */
  const float* src_feats_nhwc;  /* (B,K,H,W,C) channels-last staging of the reference's (B,K,C,H,W) */
  const float* src_extrinsics;  /* (B,K,4,4) src_cam_T_cur_cam, row-major */
  const float* src_poses;       /* (B,K,4,4) cur_cam_T_src_cam */
  const float* src_Ks;          /* (B,K,4,4) intrinsics at matching resolution */
  const float* cur_invK;        /* (B,4,4) */
  const float* plane_depths;    /* (B,D) plane depths, or (B,D,H,W) when planes_per_pixel != 0 */
  int32_t planes_per_pixel;
  /* hint branch (MLP_HINT only): maps at hint_height x hint_width, nearest-resampled to HxW in-kernel
   * (mesh_hint_volume.py:186-204).  depth_hint may hold NaN where hint_mask == 0. */
  const float* depth_hint;      /* (B,1,Hh,Wh) */
  const float* hint_weights;    /* (B,1,Hh,Wh) */
  const float* hint_mask;       /* (B,1,Hh,Wh) float 0/1 */
  int32_t hint_height, hint_width;
  /* weights, PyTorch Linear layout (out,in) row-major: keys mlp.net.{0,2,4}, hint_mlp.net.{0,2,4} */
  const float* w1; const float* b1;   /* (128, 26K+20), (128) */
  const float* w2; const float* b2;   /* (128,128), (128) */
  const float* w3; const float* b3;   /* (1,128), (1) */
  const float* hw1; const float* hb1; /* (12,3), (12) */
  const float* hw2; const float* hb2; /* (12,12), (12) */
  const float* hw3; const float* hb3; /* (1,12), (1) */
  /* outputs */
  float* volume;        /* (B,D,H,W) */
  float* lowest_cost;   /* (B,H,W) plane depth at arg-max (first max wins); may be NULL */
  int32_t* best_index;  /* (B,H,W) arg-max plane index; may be NULL */
  uint8_t* mask_views;  /* (B,K,H,W) last-plane depth-valid & in-bounds per view; may be NULL */
  uint8_t* mask_any;    /* (B,H,W) any_k(depth-valid) & any_k(in-bounds) on the last plane; may be NULL */
  /* math = TC3X / TCH: device scratch holding the MLP weights re-tiled for the tensor cores (TCH: plus 8 bytes per pixel
   * of arg-max keys), size from dtb200_cost_volume_workspace_bytes.  dtb200_cost_volume_prepare fills the weight part;
   * set workspace_prepared = 1 to reuse it on later calls with the same weights (otherwise every call re-tiles first).
   * NULL / 0 for EXACT. */
  void* workspace;
  uint64_t workspace_bytes;
  int32_t workspace_prepared;
} dtb200_cost_volume_params;

uint64_t dtb200_cost_volume_workspace_bytes(const dtb200_cost_volume_params* p);
int dtb200_cost_volume_prepare(const dtb200_cost_volume_params* p, dtb200_stream_t stream);
int dtb200_cost_volume(const dtb200_cost_volume_params* p, dtb200_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fused 2-D convolution, channels-last.  One descriptor covers every conv of
 *   BasicBlock (modules/layers.py:33-94), CVEncoder (modules/networks.py:88-117),
 *   DepthDecoderPP (modules/networks.py:20-85) and SkipDecoderRegression (modules/networks_fast.py:6-141):
 *     out = act( conv_{k x k, stride}( concat_c[ resample_i(src_i) ] ) + bias (+ residual) )
 * with torch.cat (networks.py:78,114; networks_fast.py:39) and the x2 upsamples (bilinear align_corners=False,
 * utils/generic_utils.py:95-104; nearest, networks_fast.py:38) applied on load, never materialised.
 * ---------------------------------------------------------------------------------------------------------- */
#define DTB200_RESAMPLE_NONE 0
#define DTB200_RESAMPLE_BILINEAR_UP2 1
#define DTB200_RESAMPLE_NEAREST_UP2 2

#define DTB200_ACT_NONE 0
#define DTB200_ACT_LEAKY 1 /* slope in act_slope */
#define DTB200_ACT_ELU 2   /* alpha = 1 */

#define DTB200_CONV_MAX_SRC 3

/* math = TCH: src[] and residual are split16 tensors (see dtb200_nchw_to_split16), dst is written as split16 when out_c is a
 * multiple of 64 and as plain fp32 NHWC otherwise (the 1-channel heads); x2-resampled sources must be materialised first
 * with a ksize = 0 descriptor (split16 -> split16), as for TC3X. */
typedef struct {
  int32_t math;                 /* DTB200_MATH_* */
  int32_t batch;
  int32_t in_h, in_w;           /* conv-input spatial size (after resampling): out = floor((in + 2*pad - k)/stride) + 1 */
  int32_t out_h, out_w, out_c;  /* output spatial size and channels */
  int32_t ksize, stride;        /* 3 (pad 1) or 1 (pad 0); stride 1 or 2 */
  int32_t num_src;
  const float* src[DTB200_CONV_MAX_SRC];      /* NHWC; spatial = conv-input size, or half of it when upsampled */
  int32_t src_c[DTB200_CONV_MAX_SRC];
  int32_t src_resample[DTB200_CONV_MAX_SRC];  /* DTB200_RESAMPLE_* */
  const float* weight;   /* packed by dtb200_pack_conv_weight for `math` */
  const float* bias;     /* (out_c) or NULL */
  const float* residual; /* NHWC (B,out_h,out_w,out_c) added before the activation, or NULL */
  int32_t act;
  float act_slope;
  float* dst;            /* NHWC (B,out_h,out_w,out_c) */
  /* scratch for split-K partial sums (math = TC3X on small maps); size from dtb200_conv_workspace_bytes, may be NULL
   * when that returns 0.  Launches of one sequence run in stream order, so one buffer can serve all of them. */
  void* workspace;
  uint64_t workspace_bytes;
} dtb200_conv_params;

/* ksize == 0 is a pure resample descriptor: dst = resample(src[0]) with src_resample[0] in {BILINEAR_UP2, NEAREST_UP2},
 * out_c == src_c[0], out size == in size (the up-sampled size); weight/bias/residual ignored.  The tensor-core plans
 * materialise an up-sampled map once with it instead of interpolating inside every consumer. */
uint64_t dtb200_conv_workspace_bytes(const dtb200_conv_params* p);

/* floats needed for the packed copy of an (out_c, in_c, k, k) OIHW weight */
uint64_t dtb200_packed_conv_weight_floats(int32_t math, int32_t out_c, int32_t in_c, int32_t ksize);
/* device->device repack of a PyTorch OIHW weight into the layout `math` consumes */
int dtb200_pack_conv_weight(int32_t math, const float* oihw, float* packed, int32_t out_c, int32_t in_c,
                            int32_t ksize, dtb200_stream_t stream);
/* Same, for a conv whose input is the concatenation of num_src sources with src_c[] channels each (the layout of the
 * descriptor it will be used with).  The tensor-core layout cuts every source into its own 32-channel K blocks, so the
 * packed weights depend on the split; the single-source forms above are the num_src = 1 case. */
uint64_t dtb200_packed_conv_weight_floats_srcs(int32_t math, int32_t out_c, int32_t num_src, const int32_t* src_c,
                                               int32_t ksize);
int dtb200_pack_conv_weight_srcs(int32_t math, const float* oihw, float* packed, int32_t out_c, int32_t num_src,
                                 const int32_t* src_c, int32_t ksize, dtb200_stream_t stream);
int dtb200_conv2d(const dtb200_conv_params* p, dtb200_stream_t stream);
/* Launch a whole network (an array of conv descriptors, in order) from one native call. */
int dtb200_conv2d_sequence(const dtb200_conv_params* ops, int32_t count, dtb200_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Compiled networks.  A descriptor array is analysed into its dependency DAG (from the buffers every op reads and
 * writes: sources, residual, dst, split-K workspace) and captured once into a CUDA graph in which independent layers
 * -- the parallel right/diag/up blocks of a DepthDecoderPP column (modules/networks.py:65-85), a BasicBlock's skip
 * projection next to its conv1 (modules/layers.py:77-94) -- run concurrently, so the idle SMs of one layer's last
 * tile round are taken by the next layer.  Results are bit-identical to dtb200_conv2d_sequence.
 * The descriptors' buffers (and packed weights) must stay allocated and in place while the graph is alive; ops that
 * may overlap need disjoint split-K workspaces (any two ops whose workspace ranges intersect are serialised).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct dtb200_conv_graph dtb200_conv_graph;
/* max_lanes: upper bound on concurrently runnable branches (capture streams); 1 = a linear graph */
int dtb200_conv_graph_create(const dtb200_conv_params* ops, int32_t count, int32_t max_lanes, dtb200_conv_graph** out);
int dtb200_conv_graph_launch(dtb200_conv_graph* g, dtb200_stream_t stream);
/* any pointer may be NULL: descriptors, kernels per replay, graph edges, lanes used, longest dependency chain */
int dtb200_conv_graph_info(const dtb200_conv_graph* g, int32_t* ops, int32_t* kernels, int32_t* edges, int32_t* lanes,
                           int32_t* depth);
void dtb200_conv_graph_destroy(dtb200_conv_graph* g);
/* Host-only part of the above (no CUDA call; usable without a GPU): lane and dependency level of every op and the
 * transitively reduced dependency lists in CSR form (dep_offsets has count+1 entries).  Output pointers may be NULL. */
int dtb200_conv_graph_analyze(const dtb200_conv_params* ops, int32_t count, int32_t max_lanes, int32_t* lane_of,
                              int32_t* level_of, int32_t* dep_offsets, int32_t* deps, int32_t deps_capacity);

/* Relative camera poses of DepthModel.forward (experiment_modules/doubletake_model.py:341-349):
 *   src_cam_T_cur_cam[b,k] = src_cam_T_world[b,k] @ cur_world_T_cam[b]   (the managers' src_extrinsics)
 *   cur_cam_T_src_cam[b,k] = cur_cam_T_world[b] @ src_world_T_cam[b,k]   (the managers' src_poses)
 * all row-major 4x4; src tensors (B,K,4,4), cur tensors (B,4,4). */
int dtb200_relative_poses(const float* src_cam_T_world, const float* src_world_T_cam, const float* cur_cam_T_world,
                          const float* cur_world_T_cam, float* src_cam_T_cur_cam, float* cur_cam_T_src_cam,
                          int batch, int views, dtb200_stream_t stream);

/* out = exp(in), elementwise (experiment_modules/doubletake_model.py:410-418) */
int dtb200_exp(const float* src, float* dst, uint64_t count, dtb200_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * TSDF fusion of predicted depth maps and sampling of the fused confidence: the hint-production step on the other side
 * of the hot path.  Replaces reference tools/tsdf.py:
 *   TSDFFuser.integrate_depth (:414-558)  -> dtb200_tsdf_integrate
 *   TSDF.sample_tsdf          (:277-337)  -> dtb200_tsdf_sample
 * The volume is the reference's: fp16 TSDF values and weights of shape (X, Y, Z), Z fastest, dims multiples of 8
 * (TSDF.VOX_MOD); voxel (i,j,k) sits at fp16(origin + (i,j,k) * voxel_size) (TSDF.generate_voxel_coords, :155-166), or
 * at caller-supplied fp16 coordinates (3, X, Y, Z) (TSDF.from_file).  All arithmetic reproduces the reference's fp16
 * torch ops: fp32 evaluation, one fp16 rounding per op.
 *
 * One call integrates up to DTB200_TSDF_MAX_FRAMES depth maps IN ORDER in a single pass over the volume (a voxel's
 * update depends only on its own previous state), so the volume crosses HBM once per batch instead of once per frame.
 * Per frame the caller supplies, as fp32 numbers that hold fp16 values exactly (computed with the reference's own
 * small host-side ops): P = (K @ cam_T_world)[:3] and the frustum bounding box of get_frustum_bounds (:15-50).
 * ---------------------------------------------------------------------------------------------------------- */
#define DTB200_TSDF_MAX_FRAMES 8
#define DTB200_TSDF_SEMANTICS_ATEN_CPU 0  /* fp16 grid_sample as ATen's CPU build evaluates it (pinned by the fixtures) */
#define DTB200_TSDF_SEMANTICS_ATEN_CUDA 1 /* ... as ATen's CUDA build does: index arithmetic in fp32 (opmath_t), saturating
                                             index cast -- pinned against torch 2.11 CUDA on the GPU box */
#define DTB200_TSDF_SEMANTICS_ATEN_CUDA_HALF_INDEX 2 /* older ATen CUDA builds: the un-normalised index is rounded to fp16
                                                        before nearbyint (restated from source, not executable here) */

typedef struct dtb200_tsdf_frame {
  const void* depth;       /* fp16 (img_h, img_w) depth map; <= 0 = no measurement */
  const uint8_t* mask;     /* optional (img_h, img_w) bytes: 0 = pixel invalid (reads as depth -1), or NULL */
  float P[12];             /* rows 0..2 of fp16(K @ cam_T_world) */
  float box_min[3], box_max[3]; /* fp16 frustum bounds in world space; only voxels strictly inside are touched */
} dtb200_tsdf_frame;

typedef struct dtb200_tsdf_integrate_params {
  void* values;            /* fp16 (X, Y, Z), updated in place */
  void* weights;           /* fp16 (X, Y, Z), updated in place */
  const void* voxel_coords; /* fp16 (3, X, Y, Z) or NULL = generated from origin / voxel_size */
  float origin[3];         /* fp32 origin (TSDF.from_bounds keeps it in fp32 until the coordinates are rounded) */
  float voxel_size;
  int32_t dims[3];
  int32_t vox_begin[3], vox_end[3]; /* index box [begin, end) the launch scans (z bounds multiples of 8): any conservative
                                       cover of the frames' frustum boxes; {0,0,0} / dims scans the whole volume */
  int32_t img_h, img_w;
  int32_t num_frames;
  int32_t semantics;       /* DTB200_TSDF_SEMANTICS_* */
  float min_depth;         /* TSDFFuser.min_depth (0.5) */
  float depth_range;       /* max_depth - min_depth, evaluated in double by the caller */
  float max_depth_h;       /* fp16(max_depth) */
  float truncation;        /* fp32(truncation_size * voxel_size) */
  float trunc_check_h;     /* fp16(-truncation) or fp16(-1.5 * truncation) (extended_neg_truncation) */
  dtb200_tsdf_frame frames[DTB200_TSDF_MAX_FRAMES];
} dtb200_tsdf_integrate_params;

int dtb200_tsdf_integrate(const dtb200_tsdf_integrate_params* p, dtb200_stream_t stream);

/* out[n] = volume sampled at world point n (fp32 (N,3)); volume = fp16 (X, Y, Z); dims and origin_h are HOST arrays of 3
 * (origin_h = fp16(origin) widened to fp32, as TSDF.origin is stored);
 * mode 0 = trilinear ("bilinear" on a 5-D input), 1 = nearest; align_corners=True, zeros padding, fp32 coordinates
 * (the reference's CPU branch, tools/tsdf.py:323-334). */
int dtb200_tsdf_sample(const void* volume, const int32_t* dims, const float* origin_h, float voxel_size,
                       const float* world_points, float* out, int64_t num_points, int32_t mode, dtb200_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Matching-feature encoder: reference modules/networks.py:138-189 ResnetMatchingEncoder --
 *   conv 7x7/2 + BN + ReLU (net.0-2)            -> dtb200_encoder_stem (BatchNorm folded into weight / bias by the caller)
 *   MaxPool2d(2,1) + BlurPool(4,2) | MaxPool2d(3,2,1) (net.3) -> dtb200_encoder_pool
 *   ResNet layer1, 1x1 conv, 3x3 conv           -> ordinary dtb200_conv2d descriptors (BatchNorm folded)
 *   InstanceNorm2d [+ LeakyReLU] (net.6-7, net.9) -> dtb200_instance_norm, which can also write a replicate-padded map (so
 *       that the padding_mode="replicate" 3x3 conv runs as a zero-padded conv on the enlarged map and is cropped again by
 *       the next call's src_border) and the final features in the cost-volume kernels' layouts.
 * ---------------------------------------------------------------------------------------------------------- */
#define DTB200_LAYOUT_F32 0      /* (N, H, W, C) fp32 */
#define DTB200_LAYOUT_SPLIT16 1  /* (N, H, W, 2, C) fp16 big | small, see dtb200_nchw_to_split16 */

/* dst (N, H/2, W/2, 64) fp32 NHWC = relu(conv7x7/2(image) + bias); weight_oihw (64,3,7,7) and bias (64) with the BatchNorm
 * folded in; packed_weight: scratch of 147*64 floats (re-packed on every call, 9408 floats) */
int dtb200_encoder_stem(const float* image_nchw, const float* weight_oihw, const float* bias, float* packed_weight,
                        float* dst_nhwc, int n, int h, int w, dtb200_stream_t stream);
/* src (N, h, w, c) fp32 NHWC -> dst in `dst_layout`; variant 0: MaxPool2d(2, stride 1) + BlurPool(filt 4, stride 2, reflect),
 * output ((h-1+3-4)/2+1, ...) = (h/2, w/2) for even sizes; variant 1: MaxPool2d(3, stride 2, padding 1) */
int dtb200_encoder_pool(const float* src_nhwc, void* dst, int32_t dst_layout, int n, int h, int w, int c, int32_t variant,
                        dtb200_stream_t stream);

typedef struct dtb200_instance_norm_params {
  const void* src;        /* (N, H + 2*src_border, W + 2*src_border, src_channels) in src_layout; the border is skipped */
  int32_t src_layout, src_channels, src_border;
  int32_t batch, height, width;
  int32_t channels;       /* the first `channels` (multiple of 4) channels of src are normalised and written */
  float eps;              /* 1e-5 */
  int32_t act;            /* DTB200_ACT_NONE or DTB200_ACT_LEAKY */
  float act_slope;
  void* dst;              /* (N, H + 2*dst_border, W + 2*dst_border, channels) in dst_layout, border = replicate padding; or NULL */
  int32_t dst_layout, dst_border;
  float* dst_nchw;        /* optional (N, channels, H, W) fp32 copy (needs dst_border == 0); or NULL */
  float* stats;           /* scratch, N * channels * 2 floats (mean, 1/sqrt(var + eps)) */
} dtb200_instance_norm_params;

int dtb200_instance_norm(const dtb200_instance_norm_params* p, dtb200_stream_t stream);

/* Rendered-depth hint of the incremental loop by ray casting the fused TSDF (mesh-free): one kernel
 * replaces reference marching cubes (tools/marching_cubes/marching_cubes.cu:164-424) + the mesh depth rasteriser
 * (utils/rendering_utils.py:25-53) + BackprojectDepth + TSDF.sample_tsdf + the threshold / NaN / mask rules of
 * test_incremental.py:215-252.  Per hint pixel: the first front-facing zero crossing of the TSDF along the camera ray between
 * two observed samples, its camera-space depth, and the fused confidence sampled there exactly as dtb200_tsdf_sample does. */
typedef struct dtb200_tsdf_raycast_params {
  const void* values;        /* fp16 (X, Y, Z) */
  const void* weights;       /* fp16 (X, Y, Z) */
  int32_t dims[3];
  float origin_h[3];         /* fp16(origin) widened, as TSDF.origin is stored */
  float voxel_size;
  const float* invK;         /* DEVICE (B,4,4) inverse intrinsics at the hint resolution (invK_s0_b44) */
  const float* world_T_cam;  /* DEVICE (B,4,4) camera pose */
  int32_t batch, height, width;
  float z_near, z_far;       /* marching range along camera z */
  int32_t max_steps;         /* hard bound on marching steps per ray */
  float weight_threshold;    /* 0.025 (test_incremental.py:244) */
  float* depth_hint;         /* (B,1,H,W): depth or NaN                 -> cur_data["depth_hint_b1hw"] */
  float* hint_mask;          /* (B,1,H,W): 1 where depth_hint is valid   -> cur_data["depth_hint_mask_b1hw"] */
  float* sampled_weights;    /* (B,1,H,W): confidence, 0 where invalid   -> cur_data["sampled_weights_b1hw"] */
} dtb200_tsdf_raycast_params;

int dtb200_tsdf_raycast(const dtb200_tsdf_raycast_params* p, dtb200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DOUBLETAKE_B200_H */
