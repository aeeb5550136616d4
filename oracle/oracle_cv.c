/* Plain-C restatement of the reference's plane-sweep cost volumes.
 * TEST INFRASTRUCTURE ONLY: loaded by tests/ through oracle/oracle_cv.py, never by doubletake_b200/.
 *
 * One scalar loop nest per (batch, pixel, plane), every step a separately rounded fp32 operation in the order the reference's
 * torch ops evaluate them (compile with -ffp-contract=off).  Reference (src/doubletake/):
 *   backproject      utils/geometry_utils.py:34-39,55-63      X = d * (invK[:3,:3] @ (x+0.5, y+0.5, 1))
 *   project          utils/geometry_utils.py:77-93            P = K_src @ src_cam_T_cur_cam; z' = z + 1e-8; s = |z|>1e-8 ? 1/z' : 1
 *   warp             modules/mesh_hint_volume.py:238-249      g = 2*pix*(1/W,1/H) - 1; ATen grid_sampler_2d bilinear / zeros /
 *                                                             align_corners=False in the CPU kernel's evaluation order
 *   dot + mask       modules/cost_volume.py:300-313, mesh_hint_volume.py:334-340
 *   rays / angle     modules/feature_volume.py:262-300 (F.normalize eps 1e-12, cosine_similarity eps 1e-5),
 *                    utils/geometry_utils.py:178-182
 *   pose distance    utils/geometry_utils.py:187-199
 *   MLP input order  modules/mesh_hint_volume.py:343-367      26K+20 channels
 *   MLP              modules/networks.py:120-135              Linear-LReLU(0.01)-Linear-LReLU(0.01)-Linear
 *   hint             modules/mesh_hint_volume.py:186-214,373-386
 *   arg-max          modules/cost_volume.py:317-320,356-361   first maximum wins, NaN is maximal (torch.argmax)
 *   masks            modules/cost_volume.py:73-94, mesh_hint_volume.py:273-287,818-822 (last plane only)
 * Pinned against the fixtures produced by executing the reference (tests/test_oracle_cv_c.py). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define HIDDEN 128
#define MAX_VIEWS 16

typedef struct {
  int32_t kind; /* 0 dot-product volume, 1 metadata-MLP volume, 2 MLP + depth hint */
  int32_t B, K, C, H, W, D;
  const float *cur_feats, *src_feats;                /* (B,C,H,W), (B,K,C,H,W) */
  const float *src_extrinsics, *src_poses, *src_Ks;  /* (B,K,4,4): src_cam_T_cur_cam, cur_cam_T_src_cam, K at matching res */
  const float *cur_invK;                             /* (B,4,4) */
  const float *planes;                               /* (D) plane depths */
  const float *w1, *b1, *w2, *b2, *w3, *b3;          /* mlp.net.{0,2,4} */
  const float *hw1, *hb1, *hw2, *hb2, *hw3, *hb3;    /* hint_mlp.net.{0,2,4} */
  const float *hint_depth, *hint_weights, *hint_mask; /* (B,1,hint_h,hint_w) */
  int32_t hint_h, hint_w;
  float *volume;       /* (B,D,H,W) */
  int32_t *index;      /* (B,H,W) arg-max plane */
  float *lowest;       /* (B,H,W) depth of the arg-max plane */
  uint8_t *mask_views; /* (B,K,H,W) depth-valid && in-bounds at the last plane, or NULL */
  uint8_t *mask_any;   /* (B,H,W) any-view depth-valid && any-view in-bounds at the last plane, or NULL */
} orc_cv_params;

static float leaky(float x) { return x > 0.f ? x : x * 0.01f; }

static int better(float v, int i, float bv, int bi) {
  const int vn = isnan(v), bn = isnan(bv);
  if (vn != bn) return vn;
  if (!vn && v != bv) return v > bv;
  return i < bi;
}

/* ATen grid_sampler_2d (bilinear, zeros padding, align_corners=False), one channel vector of C floats at pixel coords
 * (u, v) of a (C,H,W) map.  Evaluation order of ATen's vectorised CPU kernel (GridSamplerKernel.cpp: ComputeLocation +
 * ApplyGridSample<bilinear>), which is what the reference executed on CPU -- and therefore the fixtures -- follow:
 * i = (g + 1) * (size / 2) - 0.5; w = i - floor(i), e = 1 - w (same for n, s); taps nw = s*e, ne = s*w, sw = n*e, se = n*w;
 * out = ((nw_val*nw + ne_val*ne) + sw_val*sw) + se_val*se with out-of-image taps read as 0.
 * (The scalar / CUDA form ((g+1)*size-1)/2 with weights (x1-i)(y1-i) differs from it by an ulp of the coordinate.) */
static void bilinear(const float* src, int C, int H, int W, float u, float v, float inv_w, float inv_h, float* out) {
  const float gx = (2.f * u) * inv_w - 1.f, gy = (2.f * v) * inv_h - 1.f;
  const float ix = (gx + 1.f) * ((float)W / 2.f) - 0.5f, iy = (gy + 1.f) * ((float)H / 2.f) - 0.5f;
  const float x0 = floorf(ix), y0 = floorf(iy), x1 = x0 + 1.f, y1 = y0 + 1.f;
  const float ww = ix - x0, e = 1.f - ww, n = iy - y0, so = 1.f - n;
  const float w[4] = {so * e, so * ww, n * e, n * ww};
  const float tx[4] = {x0, x1, x0, x1}, ty[4] = {y0, y0, y1, y1};
  long off[4];
  int ok[4];
  for (int t = 0; t < 4; ++t) {
    /* NaN / huge coordinates fail every comparison */
    ok[t] = tx[t] >= 0.f && tx[t] <= (float)(W - 1) && ty[t] >= 0.f && ty[t] <= (float)(H - 1);
    off[t] = ok[t] ? (long)ty[t] * W + (long)tx[t] : 0;
  }
  for (int c = 0; c < C; ++c) {
    const float* s = src + (long)c * H * W;
    const float a = (ok[0] ? s[off[0]] : 0.f) * w[0], b = (ok[1] ? s[off[1]] : 0.f) * w[1];
    const float cc = (ok[2] ? s[off[2]] : 0.f) * w[2], d = (ok[3] ? s[off[3]] : 0.f) * w[3];
    out[c] = ((a + b) + cc) + d;
  }
}

static void linear(const float* x, int in, const float* w, const float* b, int out, float* y) {
  for (int o = 0; o < out; ++o) {
    float acc = 0.f;
    for (int i = 0; i < in; ++i) acc = acc + x[i] * w[(long)o * in + i];
    y[o] = acc + b[o];
  }
}

int orc_cost_volume(const orc_cv_params* p) {
  const int B = p->B, K = p->K, C = p->C, H = p->H, W = p->W, D = p->D;
  if (K < 1 || K > MAX_VIEWS || C < 1 || C > 64) return -1;
  const long HW = (long)H * W;
  const int F = (C + 10) * K + C + 4; /* 26K+20 at C = 16 */
  const float inv_w = 1.f / (float)W, inv_h = 1.f / (float)H; /* uv_scale, mesh_hint_volume.py:142-146 */

  for (int b = 0; b < B; ++b) {
    /* per-view constants */
    float P[MAX_VIEWS][12], tsrc[MAX_VIEWS][3], comb[MAX_VIEWS], rm[MAX_VIEWS], tm[MAX_VIEWS];
    for (int k = 0; k < K; ++k) {
      const float* Ks = p->src_Ks + ((long)b * K + k) * 16;
      const float* E = p->src_extrinsics + ((long)b * K + k) * 16;
      const float* T = p->src_poses + ((long)b * K + k) * 16;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) {
          float acc = Ks[i * 4 + 0] * E[0 * 4 + j];
          for (int l = 1; l < 4; ++l) acc = acc + Ks[i * 4 + l] * E[l * 4 + j];
          P[k][i * 4 + j] = acc;
        }
      tsrc[k][0] = T[3], tsrc[k][1] = T[7], tsrc[k][2] = T[11];
      const float trace = (T[0] + T[5]) + T[10];
      rm[k] = sqrtf(2.f * (1.f - fminf(3.f, trace) / 3.f));
      tm[k] = sqrtf((tsrc[k][0] * tsrc[k][0] + tsrc[k][1] * tsrc[k][1]) + tsrc[k][2] * tsrc[k][2]);
      comb[k] = sqrtf(tm[k] * tm[k] + rm[k] * rm[k]);
    }
    const float* invK = p->cur_invK + (long)b * 16;

#pragma omp parallel for schedule(dynamic, 64)
    for (long pix = 0; pix < HW; ++pix) {
      const int y = (int)(pix / W), x = (int)(pix % W);
      float feat[(64 + 10) * MAX_VIEWS + 64 + 4], h1[HIDDEN], h2[HIDDEN], warped[64], cur[64];
      for (int c = 0; c < C; ++c) cur[c] = p->cur_feats[((long)b * C + c) * HW + pix];
      const float px = (float)x + 0.5f, py = (float)y + 0.5f;
      float r[3];
      for (int i = 0; i < 3; ++i) r[i] = (invK[i * 4 + 0] * px + invK[i * 4 + 1] * py) + invK[i * 4 + 2];
      float hint_d = 0.f, hint_w = 0.f;
      int hint_ok = 0;
      if (p->kind == 2) { /* F.interpolate(mode="nearest") to matching res: source index floor(dst * in/out) */
        int sy = (int)floorf((float)y * ((float)p->hint_h / (float)H)), sx = (int)floorf((float)x * ((float)p->hint_w / (float)W));
        if (sy > p->hint_h - 1) sy = p->hint_h - 1;
        if (sx > p->hint_w - 1) sx = p->hint_w - 1;
        const long o = ((long)b * p->hint_h + sy) * p->hint_w + sx;
        hint_ok = p->hint_mask[o] != 0.f;
        hint_d = p->hint_depth[o];
        hint_w = hint_ok ? p->hint_weights[o] : 0.f;
      }
      float best = 0.f;
      int besti = -1;
      for (int d = 0; d < D; ++d) {
        const float depth = p->planes[d];
        const float X[3] = {depth * r[0], depth * r[1], depth * r[2]};
        float nc = sqrtf((X[0] * X[0] + X[1] * X[1]) + X[2] * X[2]);
        nc = fmaxf(nc, 1e-12f);
        const float rc[3] = {X[0] / nc, X[1] / nc, X[2] / nc};
        float n1 = fmaxf(sqrtf((rc[0] * rc[0] + rc[1] * rc[1]) + rc[2] * rc[2]), 1e-5f);
        float sum = 0.f;
        int any_d = 0, any_b = 0;
        /* channel offsets, mesh_hint_volume.py:343-367 */
        const int oCur = C * K, oMask = oCur + C, oDepth = oMask + K, oPlane = oDepth + K, oDot = oPlane + 1;
        const int oAngle = oDot + K, oRayCur = oAngle + K, oRaySrc = oRayCur + 3, oComb = oRaySrc + 3 * K;
        const int oRm = oComb + K, oTm = oRm + K;
        for (int k = 0; k < K; ++k) {
          float q[3];
          for (int i = 0; i < 3; ++i)
            q[i] = ((P[k][i * 4 + 0] * X[0] + P[k][i * 4 + 1] * X[1]) + P[k][i * 4 + 2] * X[2]) + P[k][i * 4 + 3];
          const float zp = q[2] + 1e-8f;
          const float s = fabsf(q[2]) > 1e-8f ? 1.f / zp : 1.f;
          const float u = q[0] * s, v = q[1] * s;
          bilinear(p->src_feats + ((long)b * K + k) * C * HW, C, H, W, u, v, inv_w, inv_h, warped);
          const int depth_ok = zp > 0.f;
          float dot = 0.f;
          for (int c = 0; c < C; ++c) dot = dot + warped[c] * cur[c];
          dot = dot * (depth_ok ? 1.f : 0.f);
          sum = sum + dot;
          if (d == D - 1) {
            const int inb = (u > 2.f) && (u < (float)(W - 2)) && (v > 2.f) && (v < (float)(H - 2));
            any_d |= depth_ok, any_b |= inb;
            if (p->mask_views) p->mask_views[((long)b * K + k) * HW + pix] = (uint8_t)(depth_ok && inb);
          }
          if (p->kind == 0) continue;
          memcpy(feat + C * k, warped, sizeof(float) * C);
          feat[oMask + k] = depth_ok ? 1.f : 0.f;
          feat[oDepth + k] = zp;
          feat[oDot + k] = dot;
          const float yv[3] = {X[0] - tsrc[k][0], X[1] - tsrc[k][1], X[2] - tsrc[k][2]};
          const float sc = fmaxf(sqrtf((yv[0] * yv[0] + yv[1] * yv[1]) + yv[2] * yv[2]), 1e-12f);
          const float rs[3] = {yv[0] / sc, yv[1] / sc, yv[2] / sc};
          const float n2 = fmaxf(sqrtf((rs[0] * rs[0] + rs[1] * rs[1]) + rs[2] * rs[2]), 1e-5f);
          feat[oAngle + k] = ((rc[0] / n1) * (rs[0] / n2) + (rc[1] / n1) * (rs[1] / n2)) + (rc[2] / n1) * (rs[2] / n2);
          feat[oRaySrc + 3 * k + 0] = rs[0], feat[oRaySrc + 3 * k + 1] = rs[1], feat[oRaySrc + 3 * k + 2] = rs[2];
          feat[oComb + k] = comb[k], feat[oRm + k] = rm[k], feat[oTm + k] = tm[k];
        }
        if (d == D - 1 && p->mask_any) p->mask_any[(long)b * HW + pix] = (uint8_t)(any_d && any_b);
        float score = sum;
        if (p->kind != 0) {
          memcpy(feat + oCur, cur, sizeof(float) * C);
          feat[oPlane] = depth;
          feat[oRayCur + 0] = rc[0], feat[oRayCur + 1] = rc[1], feat[oRayCur + 2] = rc[2];
          linear(feat, F, p->w1, p->b1, HIDDEN, h1);
          for (int i = 0; i < HIDDEN; ++i) h1[i] = leaky(h1[i]);
          linear(h1, HIDDEN, p->w2, p->b2, HIDDEN, h2);
          for (int i = 0; i < HIDDEN; ++i) h2[i] = leaky(h2[i]);
          linear(h2, HIDDEN, p->w3, p->b3, 1, &score);
          if (p->kind == 2) {
            const float in[3] = {score, hint_ok ? fabsf(hint_d - depth) : -1.f, hint_w};
            float a[12], c2[12];
            linear(in, 3, p->hw1, p->hb1, 12, a);
            for (int i = 0; i < 12; ++i) a[i] = leaky(a[i]);
            linear(a, 12, p->hw2, p->hb2, 12, c2);
            for (int i = 0; i < 12; ++i) c2[i] = leaky(c2[i]);
            linear(c2, 12, p->hw3, p->hb3, 1, &score);
          }
        }
        p->volume[((long)b * D + d) * HW + pix] = score;
        if (besti < 0 || better(score, d, best, besti)) best = score, besti = d;
      }
      if (p->index) p->index[(long)b * HW + pix] = besti;
      if (p->lowest) p->lowest[(long)b * HW + pix] = p->planes[besti];
    }
  }
  return 0;
}
