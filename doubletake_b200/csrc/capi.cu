// C-ABI glue: error string, launch counter, layout transposes, elementwise exp, conv dispatch.
#include "common.cuh"

namespace dtb200 {

thread_local char g_error[512] = "";
std::atomic<uint64_t> g_launches{0};

int launch_conv_simt(const dtb200_conv_params& p, int in_c_total, cudaStream_t stream);                  // conv_simt.cu
int launch_pack_simt(const float* oihw, float* packed, int out_c, int in_c, int ksize, cudaStream_t s);  // conv_simt.cu
int launch_conv_tc(const dtb200_conv_params& p, int in_c_total, cudaStream_t stream);                    // conv_tc.cu
int launch_pack_tc(const float* oihw, float* packed, int out_c, int num_src, const int32_t* src_c, int ksize,
                   cudaStream_t s);                                                                      // conv_tc.cu
uint64_t packed_floats_tc(int out_c, int num_src, const int32_t* src_c, int ksize);                      // conv_tc.cu
uint64_t conv_tc_workspace_bytes(const dtb200_conv_params& p, int in_c_total);                           // conv_tc.cu
int launch_resample_copy(const dtb200_conv_params& p, cudaStream_t stream);                              // conv_tc.cu
int conv_tc_debug_set(int flags);                                                                        // conv_tc.cu
int launch_conv_tch(const dtb200_conv_params& p, int in_c_total, cudaStream_t stream);                   // conv_tch.cu
int launch_pack_tch(const float* oihw, float* packed, int out_c, int num_src, const int32_t* src_c, int ksize,
                    cudaStream_t s);                                                                     // conv_tch.cu
uint64_t packed_floats_tch(int out_c, int num_src, const int32_t* src_c, int ksize);                     // conv_tch.cu
uint64_t conv_tch_workspace_bytes(const dtb200_conv_params& p);                                          // conv_tch.cu
int conv_tch_trace_set(unsigned long long* buf, int capacity);                                            // conv_tch.cu
int launch_resample_copy_h(const dtb200_conv_params& p, cudaStream_t stream);                            // conv_tch.cu
int launch_split16_transpose(const void* src, void* dst, int n, int c, int hw, bool to_split, cudaStream_t stream);  // conv_tch.cu

// (N,C,H,W) <-> (N,H,W,C): 32x32 smem tile transpose of the (C, H*W) matrix of each sample.
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  __shared__ float tile[32][33];
  const float* s = src + (long long)blockIdx.z * rows * cols;
  float* d = dst + (long long)blockIdx.z * rows * cols;
  int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = s[(long long)r * cols + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) d[(long long)c * rows + r] = tile[threadIdx.x][i];
  }
}

static int transpose(const float* src, float* dst, int n, int rows, int cols, cudaStream_t stream) {
  if (!src || !dst || n < 1 || rows < 1 || cols < 1) return fail(DTB200_ERR_INVALID, "transpose: bad arguments%s");
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32), n);
  transpose_kernel<<<grid, dim3(32, 8), 0, stream>>>(src, dst, rows, cols);
  return check_launch("transpose_kernel");
}

__global__ void exp_kernel(const float* __restrict__ src, float* __restrict__ dst, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    dst[i] = expf(src[i]);
}

// DepthModel.forward pre-amble (experiment_modules/doubletake_model.py:341-349): per (b,k)
//   src_cam_T_cur_cam = src_cam_T_world @ cur_world_T_cam ;  cur_cam_T_src_cam = cur_cam_T_world @ src_world_T_cam
__global__ void relative_pose_kernel(const float* __restrict__ src_cam_T_world, const float* __restrict__ src_world_T_cam,
                                     const float* __restrict__ cur_cam_T_world, const float* __restrict__ cur_world_T_cam,
                                     float* __restrict__ ext, float* __restrict__ pose, int batch, int views) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per output element of both products
  int total = batch * views * 16;
  if (i >= total) return;
  int e = i & 15, bk = i >> 4, b = bk / views;
  int r = e >> 2, c = e & 3;
  const float* A = src_cam_T_world + bk * 16;
  const float* Bm = cur_world_T_cam + b * 16;
  float acc = DT_MUL(A[r * 4 + 0], Bm[0 * 4 + c]);
  acc = DT_FMA(A[r * 4 + 1], Bm[1 * 4 + c], acc);
  acc = DT_FMA(A[r * 4 + 2], Bm[2 * 4 + c], acc);
  acc = DT_FMA(A[r * 4 + 3], Bm[3 * 4 + c], acc);
  ext[i] = acc;
  const float* Cm = cur_cam_T_world + b * 16;
  const float* Dm = src_world_T_cam + bk * 16;
  acc = DT_MUL(Cm[r * 4 + 0], Dm[0 * 4 + c]);
  acc = DT_FMA(Cm[r * 4 + 1], Dm[1 * 4 + c], acc);
  acc = DT_FMA(Cm[r * 4 + 2], Dm[2 * 4 + c], acc);
  acc = DT_FMA(Cm[r * 4 + 3], Dm[3 * 4 + c], acc);
  pose[i] = acc;
}

static int conv_total_in_c(const dtb200_conv_params& p, int& total) {
  if (p.num_src < 1 || p.num_src > DTB200_CONV_MAX_SRC) return fail(DTB200_ERR_INVALID, "conv: num_src must be 1..3%s");
  total = 0;
  for (int s = 0; s < p.num_src; ++s) {
    if (!p.src[s] || p.src_c[s] < 1) return fail(DTB200_ERR_INVALID, "conv: null/empty source%s");
    if (p.src_resample[s] != DTB200_RESAMPLE_NONE && ((p.in_h & 1) || (p.in_w & 1)))
      return fail(DTB200_ERR_INVALID, "conv: x2-upsampled source needs an even conv-input size%s");
    total += p.src_c[s];
  }
  if (!p.weight || !p.dst) return fail(DTB200_ERR_INVALID, "conv: null weight/dst%s");
  if (!((p.ksize == 3 || p.ksize == 1) && (p.stride == 1 || p.stride == 2)))
    return fail(DTB200_ERR_UNSUPPORTED, "conv: only k in {1,3}, stride in {1,2}%s");
  int pad = p.ksize / 2;
  if (p.out_h != (p.in_h + 2 * pad - p.ksize) / p.stride + 1 || p.out_w != (p.in_w + 2 * pad - p.ksize) / p.stride + 1)
    return fail(DTB200_ERR_INVALID, "conv: out size does not match in size/stride%s");
  if (p.batch < 1 || p.out_c < 1) return fail(DTB200_ERR_INVALID, "conv: empty shape%s");
  return DTB200_OK;
}

}  // namespace dtb200

using namespace dtb200;

extern "C" int dtb200_debug_set(int flags) { return conv_tc_debug_set(flags); }
extern "C" int dtb200_debug_trace(uint64_t* device_pairs, int32_t capacity) {
  return conv_tch_trace_set(reinterpret_cast<unsigned long long*>(device_pairs), capacity);
}
extern "C" int dtb200_abi_version(void) { return DTB200_ABI_VERSION; }
extern "C" const char* dtb200_last_error(void) { return g_error; }
extern "C" uint64_t dtb200_launch_count(void) { return g_launches.load(); }

extern "C" int dtb200_nchw_to_nhwc(const float* src, float* dst, int n, int c, int h, int w, dtb200_stream_t stream) {
  return transpose(src, dst, n, c, h * w, (cudaStream_t)stream);
}
extern "C" int dtb200_nhwc_to_nchw(const float* src, float* dst, int n, int c, int h, int w, dtb200_stream_t stream) {
  return transpose(src, dst, n, h * w, c, (cudaStream_t)stream);
}

extern "C" int dtb200_nchw_to_split16(const float* src, void* dst, int n, int c, int h, int w, dtb200_stream_t stream) {
  if (c % 8 != 0) return fail(DTB200_ERR_INVALID, "nchw_to_split16: channels must be a multiple of 8, got %s%lld", "", c);
  return launch_split16_transpose(src, dst, n, c, h * w, true, (cudaStream_t)stream);
}
extern "C" int dtb200_split16_to_nchw(const void* src, float* dst, int n, int c, int h, int w, dtb200_stream_t stream) {
  return launch_split16_transpose(src, dst, n, c, h * w, false, (cudaStream_t)stream);
}

extern "C" int dtb200_relative_poses(const float* src_cam_T_world, const float* src_world_T_cam,
                                     const float* cur_cam_T_world, const float* cur_world_T_cam,
                                     float* src_cam_T_cur_cam, float* cur_cam_T_src_cam, int batch, int views,
                                     dtb200_stream_t stream) {
  if (!src_cam_T_world || !src_world_T_cam || !cur_cam_T_world || !cur_world_T_cam || !src_cam_T_cur_cam ||
      !cur_cam_T_src_cam || batch < 1 || views < 1)
    return fail(DTB200_ERR_INVALID, "relative_poses: bad arguments%s");
  int total = batch * views * 16;
  relative_pose_kernel<<<ceil_div(total, 128), 128, 0, (cudaStream_t)stream>>>(
      src_cam_T_world, src_world_T_cam, cur_cam_T_world, cur_world_T_cam, src_cam_T_cur_cam, cur_cam_T_src_cam, batch, views);
  return check_launch("relative_pose_kernel");
}

extern "C" int dtb200_exp(const float* src, float* dst, uint64_t count, dtb200_stream_t stream) {
  if (!src || !dst) return fail(DTB200_ERR_INVALID, "exp: null pointer%s");
  if (count == 0) return DTB200_OK;
  uint64_t blocks = (count + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  exp_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, count);
  return check_launch("exp_kernel");
}

extern "C" uint64_t dtb200_packed_conv_weight_floats(int32_t math, int32_t out_c, int32_t in_c, int32_t ksize) {
  return dtb200_packed_conv_weight_floats_srcs(math, out_c, 1, &in_c, ksize);
}

extern "C" uint64_t dtb200_packed_conv_weight_floats_srcs(int32_t math, int32_t out_c, int32_t num_src, const int32_t* src_c,
                                                          int32_t ksize) {
  if (!src_c || num_src < 1 || num_src > DTB200_CONV_MAX_SRC) return 0;
  if (math == DTB200_MATH_TC3X) return packed_floats_tc(out_c, num_src, src_c, ksize);
  if (math == DTB200_MATH_TCH) return packed_floats_tch(out_c, num_src, src_c, ksize);
  uint64_t in_c = 0;
  for (int s = 0; s < num_src; ++s) in_c += src_c[s];
  return (uint64_t)out_c * in_c * ksize * ksize;
}

extern "C" int dtb200_pack_conv_weight(int32_t math, const float* oihw, float* packed, int32_t out_c, int32_t in_c,
                                       int32_t ksize, dtb200_stream_t stream) {
  return dtb200_pack_conv_weight_srcs(math, oihw, packed, out_c, 1, &in_c, ksize, stream);
}

extern "C" int dtb200_pack_conv_weight_srcs(int32_t math, const float* oihw, float* packed, int32_t out_c, int32_t num_src,
                                            const int32_t* src_c, int32_t ksize, dtb200_stream_t stream) {
  if (!oihw || !packed || !src_c || out_c < 1 || num_src < 1 || num_src > DTB200_CONV_MAX_SRC || (ksize != 1 && ksize != 3))
    return fail(DTB200_ERR_INVALID, "pack_conv_weight: bad arguments%s");
  int in_c = 0;
  for (int s = 0; s < num_src; ++s) {
    if (src_c[s] < 1) return fail(DTB200_ERR_INVALID, "pack_conv_weight: empty source%s");
    in_c += src_c[s];
  }
  if (math == DTB200_MATH_TC3X) return launch_pack_tc(oihw, packed, out_c, num_src, src_c, ksize, (cudaStream_t)stream);
  if (math == DTB200_MATH_TCH) return launch_pack_tch(oihw, packed, out_c, num_src, src_c, ksize, (cudaStream_t)stream);
  if (math == DTB200_MATH_EXACT) return launch_pack_simt(oihw, packed, out_c, in_c, ksize, (cudaStream_t)stream);
  return fail(DTB200_ERR_INVALID, "pack_conv_weight: unknown math mode%s");
}

extern "C" uint64_t dtb200_conv_workspace_bytes(const dtb200_conv_params* p) {
  if (!p || p->ksize == 0) return 0;
  if (p->math == DTB200_MATH_TCH) return conv_tch_workspace_bytes(*p);
  if (p->math != DTB200_MATH_TC3X) return 0;
  int total = 0;
  for (int s = 0; s < p->num_src && s < DTB200_CONV_MAX_SRC; ++s) total += p->src_c[s];
  return conv_tc_workspace_bytes(*p, total);
}

extern "C" int dtb200_conv2d(const dtb200_conv_params* p, dtb200_stream_t stream) {
  if (!p) return fail(DTB200_ERR_INVALID, "conv: null params%s");
  if (p->ksize == 0)
    return p->math == DTB200_MATH_TCH ? launch_resample_copy_h(*p, (cudaStream_t)stream) : launch_resample_copy(*p, (cudaStream_t)stream);
  int total = 0;
  int rc = conv_total_in_c(*p, total);
  if (rc != DTB200_OK) return rc;
  if (p->math == DTB200_MATH_TC3X) return launch_conv_tc(*p, total, (cudaStream_t)stream);
  if (p->math == DTB200_MATH_TCH) return launch_conv_tch(*p, total, (cudaStream_t)stream);
  if (p->math == DTB200_MATH_EXACT) return launch_conv_simt(*p, total, (cudaStream_t)stream);
  return fail(DTB200_ERR_INVALID, "conv: unknown math mode%s");
}

extern "C" int dtb200_conv2d_sequence(const dtb200_conv_params* ops, int32_t count, dtb200_stream_t stream) {
  if (!ops || count < 0) return fail(DTB200_ERR_INVALID, "conv sequence: bad arguments%s");
  for (int i = 0; i < count; ++i) {
    int rc = dtb200_conv2d(&ops[i], stream);
    if (rc != DTB200_OK) return rc;
  }
  return DTB200_OK;
}
