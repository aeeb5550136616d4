// Fused channels-last convolution, math = EXACT (fp32 FMA on CUDA cores), sm_100a.
//
//   dst = act( conv_{k x k, stride, pad k/2}( concat_c[ resample_i(src_i) ] ) + bias (+ residual) )
//
// One kernel covers every conv of BasicBlock (modules/layers.py:77-94), CVEncoder (modules/networks.py:110-117),
// DepthDecoderPP (modules/networks.py:65-85) and SkipDecoderRegression (modules/networks_fast.py:17-141).  torch.cat and
// the x2 upsamples (bilinear align_corners=False: utils/generic_utils.py:95-104; nearest: networks_fast.py:38) happen in
// the im2col loader, so the concatenated / upsampled tensors of the reference never exist in HBM.
//
// Implicit GEMM: M = 8x8 output pixels, N = 64 output channels, K = taps x concatenated input channels in chunks of 8;
// 256 threads, 4x4 register tile each; double-buffered shared memory with register prefetch (one barrier per chunk).
// Accumulation order per output: taps (ky,kx) ascending, sources in order, channels ascending; then bias, residual, act.
#include "common.cuh"
#include "conv_common.cuh"

namespace dtb200 {

constexpr int kTM = 64, kTN = 64, kKC = 8;
constexpr int kAS = kTM + 4;  // padded row strides (16-B aligned)
constexpr int kBS = kTN + 4;

__global__ void __launch_bounds__(256) conv_simt_kernel(const dtb200_conv_params p, int in_c_total) {
  const int in_h = p.in_h, in_w = p.in_w;
  __shared__ __align__(16) float As[2][kKC][kAS];
  __shared__ __align__(16) float Bs[2][kKC][kBS];

  const int tid = threadIdx.x;
  const int tiles_x = ceil_div(p.out_w, 8);
  const int tile = blockIdx.x;
  const int ty0 = (tile / tiles_x) * 8, tx0 = (tile % tiles_x) * 8;
  const int n0 = blockIdx.y * kTN;
  const int b = blockIdx.z;

  SrcView sv[DTB200_CONV_MAX_SRC];
#pragma unroll
  for (int s = 0; s < DTB200_CONV_MAX_SRC; ++s) {
    sv[s].ptr = p.src[s];
    sv[s].c = s < p.num_src ? p.src_c[s] : 0;
    sv[s].resample = p.src_resample[s];
    const bool up = sv[s].resample != DTB200_RESAMPLE_NONE;
    sv[s].h = up ? in_h / 2 : in_h;
    sv[s].w = up ? in_w / 2 : in_w;
  }

  // loader roles: threads 0..127 fetch A (pixel m = t/2, channel half t%2), threads 128..255 fetch B
  const bool a_loader = tid < 128;
  const int lm = tid >> 1, lh = tid & 1;
  const int l_oy = ty0 + (lm >> 3), l_ox = tx0 + (lm & 7);
  const int lb_row = (tid - 128) >> 4, lb_col = ((tid - 128) & 15) * 4;
  const int pad = p.ksize / 2;
  const int taps = p.ksize * p.ksize;
  const int chunks_per_tap = in_c_total / kKC;
  const int steps = taps * chunks_per_tap;

  // loader cursor
  int cur_tap = 0, cur_src = 0, cur_c = 0, cur_cglobal = 0;
  auto fetch = [&]() -> float4 {
    float4 v;
    if (a_loader) {
      int ky = cur_tap / p.ksize, kx = cur_tap - ky * p.ksize;
      int iy = l_oy * p.stride + ky - pad, ix = l_ox * p.stride + kx - pad;
      bool pix_ok = (l_oy < p.out_h) && (l_ox < p.out_w);
      v = pix_ok ? load_input4(sv[cur_src], b, iy, ix, in_h, in_w, cur_c + lh * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      const float* w = p.weight + ((long long)cur_tap * in_c_total + cur_cglobal + lb_row) * p.out_c + n0 + lb_col;
      v = ld4(w);
    }
    // advance cursor to the next chunk
    cur_c += kKC;
    cur_cglobal += kKC;
    if (cur_c >= sv[cur_src].c) {
      cur_c = 0;
      ++cur_src;
      if (cur_src >= p.num_src) {
        cur_src = 0;
        cur_cglobal = 0;
        ++cur_tap;
      }
    }
    return v;
  };
  auto stash = [&](int buf, float4 v) {
    if (a_loader) {
      As[buf][lh * 4 + 0][lm] = v.x;
      As[buf][lh * 4 + 1][lm] = v.y;
      As[buf][lh * 4 + 2][lm] = v.z;
      As[buf][lh * 4 + 3][lm] = v.w;
    } else {
      *reinterpret_cast<float4*>(&Bs[buf][lb_row][lb_col]) = v;
    }
  };

  const int tr = tid >> 4, tc = tid & 15;  // 4 pixels (m = 4tr..4tr+3) x 4 channels (n = 4tc..4tc+3)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  stash(0, fetch());
  __syncthreads();
  for (int step = 0; step < steps; ++step) {
    const int buf = step & 1;
    float4 nxt;
    const bool more = step + 1 < steps;
    if (more) nxt = fetch();
#pragma unroll
    for (int k = 0; k < kKC; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[buf][k][tr * 4]);
      float4 w = *reinterpret_cast<const float4*>(&Bs[buf][k][tc * 4]);
      const float as[4] = {a.x, a.y, a.z, a.w};
      const float ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = DT_FMA(as[i], ws[j], acc[i][j]);
    }
    if (more) stash(buf ^ 1, nxt);
    __syncthreads();
  }

  // epilogue: bias, residual, activation, NHWC store
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.bias) {
    float4 bv = ld4(p.bias + n0 + tc * 4);
    bias[0] = bv.x, bias[1] = bv.y, bias[2] = bv.z, bias[3] = bv.w;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = tr * 4 + i;
    int oy = ty0 + (m >> 3), ox = tx0 + (m & 7);
    if (oy >= p.out_h || ox >= p.out_w) continue;
    long long o = (((long long)b * p.out_h + oy) * p.out_w + ox) * p.out_c + n0 + tc * 4;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = DT_ADD(acc[i][j], bias[j]);
    if (p.residual) {
      float4 r = ld4(p.residual + o);
      v[0] = DT_ADD(v[0], r.x), v[1] = DT_ADD(v[1], r.y), v[2] = DT_ADD(v[2], r.z), v[3] = DT_ADD(v[3], r.w);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = activate(v[j], p.act, p.act_slope);
    *reinterpret_cast<float4*>(p.dst + o) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// 1x1 conv to few output channels (the log-depth heads: networks.py:58-61, networks_fast.py:102-131).
// One warp per pixel; lanes stride the input channels; fixed shuffle tree.
__global__ void __launch_bounds__(256) conv_head_kernel(const dtb200_conv_params p, int in_c_total, long long pixels) {
  const int lane = threadIdx.x & 31;
  const long long pix = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (pix >= pixels) return;
  for (int n = 0; n < p.out_c; ++n) {
    float s = 0.f;
    int cg = 0;
    for (int si = 0; si < p.num_src; ++si) {
      const float* x = p.src[si] + pix * p.src_c[si];
      for (int c = lane; c < p.src_c[si]; c += 32) s = DT_FMA(__ldg(x + c), __ldg(p.weight + (long long)(cg + c) * p.out_c + n), s);
      cg += p.src_c[si];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = DT_ADD(s, __shfl_xor_sync(0xffffffffu, s, o));
    if (lane == 0) {
      float v = p.bias ? DT_ADD(s, p.bias[n]) : s;
      if (p.residual) v = DT_ADD(v, p.residual[pix * p.out_c + n]);
      p.dst[pix * p.out_c + n] = activate(v, p.act, p.act_slope);
    }
  }
}

// OIHW (out_c, in_c, k, k) -> [tap][in_c][out_c]
__global__ void pack_weight_simt_kernel(const float* __restrict__ oihw, float* __restrict__ packed, int out_c, int in_c,
                                        int taps) {
  long long n = (long long)out_c * in_c * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int o = (int)(i % out_c);
    long long r = i / out_c;
    int c = (int)(r % in_c);
    int t = (int)(r / in_c);
    packed[i] = oihw[((long long)o * in_c + c) * taps + t];
  }
}

int launch_conv_simt(const dtb200_conv_params& p, int in_c_total, cudaStream_t stream) {
  if (p.ksize == 1 && p.stride == 1 && p.out_c < 64) {
    for (int s = 0; s < p.num_src; ++s)
      if (p.src_resample[s] != DTB200_RESAMPLE_NONE)
        return fail(DTB200_ERR_UNSUPPORTED, "conv head: resampled sources not supported%s");
    long long pixels = (long long)p.batch * p.out_h * p.out_w;
    conv_head_kernel<<<(unsigned)((pixels + 7) / 8), 256, 0, stream>>>(p, in_c_total, pixels);
    return check_launch("conv_head_kernel");
  }
  if (p.out_c % kTN != 0)
    return fail(DTB200_ERR_UNSUPPORTED, "conv (exact): out_c must be a multiple of 64 (or a <64-channel 1x1 head), got %s%lld", "",
                p.out_c);
  for (int s = 0; s < p.num_src; ++s)
    if (p.src_c[s] % kKC != 0)
      return fail(DTB200_ERR_UNSUPPORTED, "conv (exact): every source needs a multiple of 8 channels, got %s%lld", "", p.src_c[s]);
  dim3 grid(ceil_div(p.out_h, 8) * ceil_div(p.out_w, 8), p.out_c / kTN, p.batch);
  conv_simt_kernel<<<grid, 256, 0, stream>>>(p, in_c_total);
  return check_launch("conv_simt_kernel");
}

int launch_pack_simt(const float* oihw, float* packed, int out_c, int in_c, int ksize, cudaStream_t stream) {
  long long n = (long long)out_c * in_c * ksize * ksize;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_weight_simt_kernel<<<blocks, 256, 0, stream>>>(oihw, packed, out_c, in_c, ksize * ksize);
  return check_launch("pack_weight_simt_kernel");
}

}  // namespace dtb200
