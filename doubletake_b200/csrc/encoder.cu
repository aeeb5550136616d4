// Matching-feature encoder pieces that the fused-conv descriptors do not cover (reference
// modules/networks.py:138-189 ResnetMatchingEncoder):
//   enc_stem_kernel      conv 7x7 / stride 2 (3 -> 64) with the BatchNorm folded into weights / bias, ReLU       (net.0-2)
//   enc_pool_kernel      MaxPool2d(2, stride 1) + BlurPool(4, stride 2, reflect) of antialiased_cnns, or torchvision's
//                        MaxPool2d(3, stride 2, padding 1)                                                       (net.3)
//   inorm_*_kernel       InstanceNorm2d (eps 1e-5, no affine) [+ LeakyReLU], writing either a replicate-padded map (the input
//                        of the padding_mode="replicate" 3x3 conv, which then runs as a plain zero-padded conv on the enlarged
//                        map) or the final features in the two layouts the cost-volume kernels read: current view NCHW,
//                        source views channels-last                                                              (net.6-7, net.9)
// The ResNet layer1 blocks, the 1x1 and the 3x3 convs run as ordinary conv descriptors (BatchNorm folded on the host).
// Layouts: DTB200_LAYOUT_F32 = (N,H,W,C) fp32, DTB200_LAYOUT_SPLIT16 = (N,H,W,2,C) fp16 big | small (math = TCH).
#include "common.cuh"
#include "conv_common.cuh"
#include "tc_common.cuh"

namespace dtb200 {

using tc::join_half;
using tc::split_half2;

// ---------------------------------------------------------------------------------------------------------------------
// stem: block = 16 x 8 output pixels x 64 channels; the 37 x 21 x 3 input patch and the 147 x 64 weights sit in shared memory;
// a thread owns 4 pixels (a 2 x 2 quad) x 8 channels.  Accumulation order per output: channel, ky, kx ascending, then bias.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kStemTW = 16, kStemTH = 8;
constexpr int kStemPW = 2 * kStemTW + 5, kStemPH = 2 * kStemTH + 5;   // 37 x 21 input patch

__global__ void __launch_bounds__(256) enc_stem_kernel(const float* __restrict__ image, const float* __restrict__ wpack /* [147][64] */,
                                                       const float* __restrict__ bias, float* __restrict__ dst, int H, int W, int OH,
                                                       int OW) {
  extern __shared__ float stem_smem[];
  float* s_w = stem_smem;                       // [147][64]
  float* s_in = stem_smem + 147 * 64;           // [3][kStemPH][kStemPW + 1]
  constexpr int kRow = kStemPW + 1;
  const int n = blockIdx.z;
  const int ox0 = blockIdx.x * kStemTW, oy0 = blockIdx.y * kStemTH;
  const int tid = threadIdx.x;
  for (int i = tid; i < 147 * 64; i += 256) s_w[i] = wpack[i];
  const int ix0 = 2 * ox0 - 3, iy0 = 2 * oy0 - 3;
  for (int i = tid; i < 3 * kStemPH * kStemPW; i += 256) {
    const int c = i / (kStemPH * kStemPW), r = i - c * kStemPH * kStemPW;
    const int py = r / kStemPW, px = r - py * kStemPW;
    const int y = iy0 + py, x = ix0 + px;
    s_in[(c * kStemPH + py) * kRow + px] = (y >= 0 && y < H && x >= 0 && x < W) ? image[((size_t)(n * 3 + c) * H + y) * W + x] : 0.f;
  }
  __syncthreads();
  const int cg = tid & 7;            // channels 8cg .. 8cg+7
  const int qx = (tid >> 3) & 7, qy = tid >> 6;   // 8 x 4 quads of 2 x 2 pixels
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int c = 0; c < 3; ++c)
    for (int ky = 0; ky < 7; ++ky)
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const float* wv = s_w + ((c * 7 + ky) * 7 + kx) * 64 + cg * 8;
        const float4 w0 = *reinterpret_cast<const float4*>(wv), w1 = *reinterpret_cast<const float4*>(wv + 4);
        const float ws[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int py = 2 * (2 * qy + (i >> 1)) + ky, px = 2 * (2 * qx + (i & 1)) + kx;
          const float v = s_in[(c * kStemPH + py) * kRow + px];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = DT_FMA(v, ws[j], acc[i][j]);
        }
      }
  const float4 b0 = ld4(bias + cg * 8), b1 = ld4(bias + cg * 8 + 4);
  const float bs[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int oy = oy0 + 2 * qy + (i >> 1), ox = ox0 + 2 * qx + (i & 1);
    if (oy >= OH || ox >= OW) continue;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaxf(DT_ADD(acc[i][j], bs[j]), 0.f);
    float* d = dst + (((size_t)n * OH + oy) * OW + ox) * 64 + cg * 8;
    *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(d + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// OIHW (64, 3, 7, 7) with BatchNorm folded by the host -> [c][ky][kx][64]
__global__ void enc_pack_stem_kernel(const float* __restrict__ oihw, float* __restrict__ packed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 147 * 64) return;
  const int o = i & 63, t = i >> 6;   // t = (c*7 + ky)*7 + kx
  packed[i] = oihw[o * 147 + t];
}

// ---------------------------------------------------------------------------------------------------------------------
// pool: one thread per (output pixel, 4 channels)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 max4(float4 a, float4 b) { return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w)); }

__device__ __forceinline__ void store4_layout(void* dst, int layout, size_t pix, int C, int c, float4 v) {
  if (layout == DTB200_LAYOUT_SPLIT16) {
    uint32_t b0, s0, b1, s1;
    split_half2(v.x, v.y, b0, s0);
    split_half2(v.z, v.w, b1, s1);
    uint8_t* d = reinterpret_cast<uint8_t*>(dst) + pix * (size_t)C * 4 + (size_t)c * 2;
    *reinterpret_cast<uint2*>(d) = make_uint2(b0, b1);
    *reinterpret_cast<uint2*>(d + (size_t)C * 2) = make_uint2(s0, s1);
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + pix * C + c) = v;
  }
}
__device__ __forceinline__ float4 load4_layout(const void* src, int layout, size_t pix, int C, int c) {
  if (layout == DTB200_LAYOUT_SPLIT16) {
    const uint8_t* s = reinterpret_cast<const uint8_t*>(src) + pix * (size_t)C * 4 + (size_t)c * 2;
    const uint2 b = __ldg(reinterpret_cast<const uint2*>(s)), sm = __ldg(reinterpret_cast<const uint2*>(s + (size_t)C * 2));
    const __half2 b0 = *reinterpret_cast<const __half2*>(&b.x), b1 = *reinterpret_cast<const __half2*>(&b.y);
    const __half2 s0 = *reinterpret_cast<const __half2*>(&sm.x), s1 = *reinterpret_cast<const __half2*>(&sm.y);
    return make_float4(join_half(__low2half(b0), __low2half(s0)), join_half(__high2half(b0), __high2half(s0)),
                       join_half(__low2half(b1), __low2half(s1)), join_half(__high2half(b1), __high2half(s1)));
  }
  return ld4(reinterpret_cast<const float*>(src) + pix * C + c);
}

// variant 0: MaxPool2d(2, stride 1) then BlurPool(filt 4, stride 2, ReflectionPad2d(1, 2, 1, 2)); variant 1: MaxPool2d(3, 2, 1)
__global__ void enc_pool_kernel(const float* __restrict__ src, void* __restrict__ dst, int dst_layout, int N, int H, int W, int C,
                                int OH, int OW, int variant) {
  const int c4 = C / 4;
  const long long total = (long long)N * OH * OW * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4) * 4;
    const long long pix = i / c4;
    const int ox = (int)(pix % OW);
    const long long r = pix / OW;
    const int oy = (int)(r % OH), n = (int)(r / OH);
    const float* base = src + (size_t)n * H * W * C + c;
    auto at = [&](int y, int x) { return ld4(base + ((size_t)y * W + x) * C); };
    float4 out;
    if (variant == 1) {
      out = make_float4(-3.4e38f, -3.4e38f, -3.4e38f, -3.4e38f);
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
          const int y = 2 * oy - 1 + ky, x = 2 * ox - 1 + kx;
          if (y >= 0 && y < H && x >= 0 && x < W) out = max4(out, at(y, x));
        }
    } else {
      const int MH = H - 1, MW = W - 1;   // max-pooled size
      const float f[4] = {1.f, 3.f, 3.f, 1.f};
      out = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int ky = 0; ky < 4; ++ky) {
        int my = 2 * oy + ky - 1;          // row in the reflection-padded max-pooled map, pad (top 1, bottom 2)
        my = my < 0 ? -my : (my >= MH ? 2 * MH - 2 - my : my);
        for (int kx = 0; kx < 4; ++kx) {
          int mx = 2 * ox + kx - 1;
          mx = mx < 0 ? -mx : (mx >= MW ? 2 * MW - 2 - mx : mx);
          const float4 m = max4(max4(at(my, mx), at(my, mx + 1)), max4(at(my + 1, mx), at(my + 1, mx + 1)));
          const float w = DT_DIV(DT_MUL(f[ky], f[kx]), 64.f);
          out.x = DT_FMA(m.x, w, out.x), out.y = DT_FMA(m.y, w, out.y), out.z = DT_FMA(m.z, w, out.z), out.w = DT_FMA(m.w, w, out.w);
        }
      }
    }
    store4_layout(dst, dst_layout, (size_t)pix, C, c, out);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// InstanceNorm2d: statistics per (sample, channel) over the H x W interior of a (possibly bordered) map -- two passes, mean
// then centred sum of squares (biased variance, as torch) -- then the normalising pass
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) inorm_stats_kernel(const dtb200_instance_norm_params p, float* __restrict__ stats /* [N][C][2] */) {
  // block = (sample n, group of 4 channels); 256 threads stride the pixels
  const int n = blockIdx.y, c = blockIdx.x * 4;
  const int SW = p.width + 2 * p.src_border, SH = p.height + 2 * p.src_border;
  const size_t base = (size_t)n * SH * SW;
  __shared__ float red[4][256];
  const int HW = p.height * p.width;
  auto reduce4 = [&](float4 v) -> float4 {
    red[0][threadIdx.x] = v.x, red[1][threadIdx.x] = v.y, red[2][threadIdx.x] = v.z, red[3][threadIdx.x] = v.w;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if ((int)threadIdx.x < s) {
#pragma unroll
        for (int j = 0; j < 4; ++j) red[j][threadIdx.x] += red[j][threadIdx.x + s];
      }
      __syncthreads();
    }
    const float4 out = make_float4(red[0][0], red[1][0], red[2][0], red[3][0]);
    __syncthreads();
    return out;
  };
  auto pixel = [&](int i) -> size_t {
    const int y = i / p.width, x = i - y * p.width;
    return base + (size_t)(y + p.src_border) * SW + x + p.src_border;
  };
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = threadIdx.x; i < HW; i += 256) {
    const float4 v = load4_layout(p.src, p.src_layout, pixel(i), p.src_channels, c);
    s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
  }
  s = reduce4(s);
  const float inv = 1.f / (float)HW;
  const float4 mean = make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv);
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = threadIdx.x; i < HW; i += 256) {
    const float4 v = load4_layout(p.src, p.src_layout, pixel(i), p.src_channels, c);
    const float dx = v.x - mean.x, dy = v.y - mean.y, dz = v.z - mean.z, dw = v.w - mean.w;
    q.x = fmaf(dx, dx, q.x), q.y = fmaf(dy, dy, q.y), q.z = fmaf(dz, dz, q.z), q.w = fmaf(dw, dw, q.w);
  }
  q = reduce4(q);
  if (threadIdx.x == 0) {
    float* o = stats + ((size_t)n * p.channels + c) * 2;
    const float m[4] = {mean.x, mean.y, mean.z, mean.w}, v[4] = {q.x * inv, q.y * inv, q.z * inv, q.w * inv};
#pragma unroll
    for (int j = 0; j < 4; ++j) o[2 * j] = m[j], o[2 * j + 1] = rsqrtf(v[j] + p.eps);
  }
}

__global__ void inorm_apply_kernel(const dtb200_instance_norm_params p, const float* __restrict__ stats) {
  const int c4 = p.channels / 4;
  const int SW = p.width + 2 * p.src_border, SH = p.height + 2 * p.src_border;
  const int DW = p.width + 2 * p.dst_border, DH = p.height + 2 * p.dst_border;
  const long long total = (long long)p.batch * DH * DW * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4) * 4;
    const long long pix = i / c4;
    const int dx = (int)(pix % DW);
    const long long r = pix / DW;
    const int dy = (int)(r % DH), n = (int)(r / DH);
    // replicate padding of the destination: border pixels copy the nearest interior pixel
    const int y = min(max(dy - p.dst_border, 0), p.height - 1), x = min(max(dx - p.dst_border, 0), p.width - 1);
    const size_t spix = ((size_t)n * SH + y + p.src_border) * SW + x + p.src_border;
    float4 v = load4_layout(p.src, p.src_layout, spix, p.src_channels, c);
    const float* st = stats + ((size_t)n * p.channels + c) * 2;
    v.x = (v.x - st[0]) * st[1], v.y = (v.y - st[2]) * st[3], v.z = (v.z - st[4]) * st[5], v.w = (v.w - st[6]) * st[7];
    if (p.act == DTB200_ACT_LEAKY) {
      v.x = v.x > 0.f ? v.x : v.x * p.act_slope, v.y = v.y > 0.f ? v.y : v.y * p.act_slope;
      v.z = v.z > 0.f ? v.z : v.z * p.act_slope, v.w = v.w > 0.f ? v.w : v.w * p.act_slope;
    }
    if (p.dst) store4_layout(p.dst, p.dst_layout, (size_t)pix, p.channels, c, v);
    if (p.dst_nchw) {   // (N, C, H, W) fp32 (no border)
      float* d = p.dst_nchw + ((size_t)n * p.channels + c) * p.height * p.width + (size_t)y * p.width + x;
      const size_t cs = (size_t)p.height * p.width;
      d[0] = v.x, d[cs] = v.y, d[2 * cs] = v.z, d[3 * cs] = v.w;
    }
  }
}

}  // namespace dtb200

using namespace dtb200;

extern "C" int dtb200_encoder_stem(const float* image_nchw, const float* weight_oihw, const float* bias, float* packed_weight,
                                   float* dst_nhwc, int n, int h, int w, dtb200_stream_t stream) {
  if (!image_nchw || !weight_oihw || !bias || !packed_weight || !dst_nhwc || n < 1 || h < 1 || w < 1)
    return fail(DTB200_ERR_INVALID, "encoder_stem: bad arguments%s");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  enc_pack_stem_kernel<<<(147 * 64 + 255) / 256, 256, 0, s>>>(weight_oihw, packed_weight);
  int rc = check_launch("enc_pack_stem_kernel");
  if (rc != DTB200_OK) return rc;
  const int oh = (h + 6 - 7) / 2 + 1, ow = (w + 6 - 7) / 2 + 1;
  const size_t smem = (147 * 64 + 3 * kStemPH * (kStemPW + 1)) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(enc_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  dim3 grid((ow + kStemTW - 1) / kStemTW, (oh + kStemTH - 1) / kStemTH, n);
  enc_stem_kernel<<<grid, 256, smem, s>>>(image_nchw, packed_weight, bias, dst_nhwc, h, w, oh, ow);
  return check_launch("enc_stem_kernel");
}

extern "C" int dtb200_encoder_pool(const float* src_nhwc, void* dst, int32_t dst_layout, int n, int h, int w, int c, int32_t variant,
                                   dtb200_stream_t stream) {
  if (!src_nhwc || !dst || n < 1 || h < 3 || w < 3 || c < 4 || (c & 3) || (variant != 0 && variant != 1) ||
      (dst_layout != DTB200_LAYOUT_F32 && dst_layout != DTB200_LAYOUT_SPLIT16) || (dst_layout == DTB200_LAYOUT_SPLIT16 && (c & 7)))
    return fail(DTB200_ERR_INVALID, "encoder_pool: bad arguments%s");
  const int oh = variant == 1 ? (h + 2 - 3) / 2 + 1 : (h - 1 + 3 - 4) / 2 + 1;
  const int ow = variant == 1 ? (w + 2 - 3) / 2 + 1 : (w - 1 + 3 - 4) / 2 + 1;
  const long long total = (long long)n * oh * ow * (c / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  enc_pool_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src_nhwc, dst, dst_layout, n, h, w, c, oh, ow, variant);
  return check_launch("enc_pool_kernel");
}

extern "C" int dtb200_instance_norm(const dtb200_instance_norm_params* p, dtb200_stream_t stream) {
  if (!p || !p->src || !p->stats || (!p->dst && !p->dst_nchw)) return fail(DTB200_ERR_INVALID, "instance_norm: null argument%s");
  if (p->batch < 1 || p->height < 1 || p->width < 1 || p->channels < 4 || (p->channels & 3) || p->src_channels < p->channels ||
      p->src_border < 0 || p->dst_border < 0)
    return fail(DTB200_ERR_INVALID, "instance_norm: bad shape%s");
  if ((p->src_layout == DTB200_LAYOUT_SPLIT16 && (p->src_channels & 7)) || (p->dst && p->dst_layout == DTB200_LAYOUT_SPLIT16 && (p->channels & 7)))
    return fail(DTB200_ERR_INVALID, "instance_norm: split16 maps need a multiple of 8 channels%s");
  if (p->dst_nchw && p->dst_border != 0) return fail(DTB200_ERR_INVALID, "instance_norm: the NCHW output has no border%s");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  inorm_stats_kernel<<<dim3(p->channels / 4, p->batch), 256, 0, s>>>(*p, p->stats);
  int rc = check_launch("inorm_stats_kernel");
  if (rc != DTB200_OK) return rc;
  const long long total = (long long)p->batch * (p->height + 2 * p->dst_border) * (p->width + 2 * p->dst_border) * (p->channels / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  inorm_apply_kernel<<<blocks, 256, 0, s>>>(*p, p->stats);
  return check_launch("inorm_apply_kernel");
}
