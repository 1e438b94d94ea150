/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the step's arithmetic, independent of torch.
 *
 * A second, library-free statement of what the reference computes, used by tests/ to cross-check the
 * torch-based oracle (oracle/conv3d_oracle.py) on small cases: it pins the MEANING of the operators the
 * reference delegates to torch (cross-correlation orientation, NCDHW flatten order, Linear layout).
 * Scalar loops, double accumulation, float results.  Never linked into the product.
 * Pinned against outputs of the UNMODIFIED reference: composed by tests/test_c_oracle.py into the whole step (forward, the
 * returned and logged losses, backward, two Adam steps), the two-tower forward and the Conv3dMaxPool block, these functions
 * reproduce the golden vectors of tests/golden/ (oracle/make_golden.py) to 1e-5.
 *
 *   ora_sat_normalise : predict_pv_yield/netcdf_dataset.py:96-101
 *   ora_conv3d_relu   : nn.Conv3d(k=3, padding=0) + F.relu, predict_pv_yield/models/conv3d/model.py:80-90,117-120
 *   ora_linear        : nn.Linear (+ optional ReLU), model.py:92-103,125-152
 *   ora_l1_loss       : (y_hat - y).abs().mean(), predict_pv_yield/models/base_model.py:99
 *   ora_conv3d_pad    : nn.Conv3d(k=3, padding=(pt, ph, ph)): models/conv3d/model_sat_nwp.py:101-133 (pt = 1, ph = 0),
 *                       models/perceiver/perceiver_conv3d_nwp_sat.py:42-57 (pt = ph = 1)
 *   ora_conv3d_dgrad / ora_conv3d_wgrad : the autograd of those convolutions (aten::convolution_backward), written as
 *                       the transposed scatter / the correlation of input and output gradient
 *   ora_linear_bwd    : autograd of nn.Linear
 *   ora_maxpool3d     : nn.MaxPool3d(3, stride=(1,2,2), padding=1) with aten's arg-max rule (first maximum in scan order,
 *                       NaN wins), perceiver_conv3d_nwp_sat.py:49-51, and its backward (gradient routed to the arg-max)
 *   ora_adam_step     : torch.optim.Adam(lr) defaults, models/base_model.py:255-257 (bias-corrected, eps outside the sqrt)
 */
#include <math.h>
#include <stdint.h>

void ora_sat_normalise(const int16_t* x, float* y, const float* mean, const float* std, int B, int C, long thw) {
  for (long p = 0; p < (long)B * C; ++p) {
    const int c = (int)(p % C);
    for (long i = 0; i < thw; ++i) {
      volatile float d = (float)x[p * thw + i] - mean[c]; /* rounded to fp32 before the division */
      y[p * thw + i] = d / std[c];
    }
  }
}

/* x [B,Ci,T,H,W], w [Co,Ci,3,3,3], bias [Co] -> y [B,Co,T-2,H-2,W-2] */
void ora_conv3d_relu(const float* x, const float* w, const float* bias, float* y, int B, int Ci, int T, int H, int W,
                     int Co, int relu) {
  const int To = T - 2, Ho = H - 2, Wo = W - 2;
  for (int b = 0; b < B; ++b)
    for (int co = 0; co < Co; ++co)
      for (int t = 0; t < To; ++t)
        for (int h = 0; h < Ho; ++h)
          for (int v = 0; v < Wo; ++v) {
            double s = bias ? (double)bias[co] : 0.0;
            for (int ci = 0; ci < Ci; ++ci)
              for (int kt = 0; kt < 3; ++kt)
                for (int kh = 0; kh < 3; ++kh)
                  for (int kw = 0; kw < 3; ++kw)
                    s += (double)x[(((long)(b * Ci + ci) * T + t + kt) * H + h + kh) * W + v + kw] *
                         (double)w[(((long)(co * Ci + ci) * 3 + kt) * 3 + kh) * 3 + kw];
            float r = (float)s;
            if (relu && r < 0.f) r = 0.f;
            y[(((long)(b * Co + co) * To + t) * Ho + h) * Wo + v] = r;
          }
}

/* x [B,I], w [O,I], bias [O] -> y [B,O] */
void ora_linear(const float* x, const float* w, const float* bias, float* y, int B, long I, int O, int relu) {
  for (int b = 0; b < B; ++b)
    for (int o = 0; o < O; ++o) {
      double s = bias ? (double)bias[o] : 0.0;
      for (long i = 0; i < I; ++i) s += (double)x[b * I + i] * (double)w[o * I + i];
      float r = (float)s;
      if (relu && r < 0.f) r = 0.f;
      y[(long)b * O + o] = r;
    }
}

float ora_l1_loss(const float* y_hat, const float* y, long n) {
  double s = 0.0;
  for (long i = 0; i < n; ++i) s += fabs((double)y_hat[i] - (double)y[i]);
  return (float)(s / (double)n);
}

/* x [B,Ci,T,H,W], w [Co,Ci,3,3,3] -> y [B,Co,T+2pt-2,H+2ph-2,W+2ph-2]; implicit zero padding */
void ora_conv3d_pad(const float* x, const float* w, const float* bias, float* y, int B, int Ci, int T, int H, int W, int Co,
                    int pt, int ph, int relu) {
  const int To = T + 2 * pt - 2, Ho = H + 2 * ph - 2, Wo = W + 2 * ph - 2;
  for (int b = 0; b < B; ++b)
    for (int co = 0; co < Co; ++co)
      for (int t = 0; t < To; ++t)
        for (int h = 0; h < Ho; ++h)
          for (int v = 0; v < Wo; ++v) {
            double s = bias ? (double)bias[co] : 0.0;
            for (int ci = 0; ci < Ci; ++ci)
              for (int kt = 0; kt < 3; ++kt)
                for (int kh = 0; kh < 3; ++kh)
                  for (int kw = 0; kw < 3; ++kw) {
                    const int ti = t + kt - pt, hi = h + kh - ph, wi = v + kw - ph;
                    if (ti < 0 || ti >= T || hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
                    s += (double)x[(((long)(b * Ci + ci) * T + ti) * H + hi) * W + wi] *
                         (double)w[(((long)(co * Ci + ci) * 3 + kt) * 3 + kh) * 3 + kw];
                  }
            float r = (float)s;
            if (relu && r < 0.f) r = 0.f;
            y[(((long)(b * Co + co) * To + t) * Ho + h) * Wo + v] = r;
          }
}

/* gx[b,ci,ti,hi,wi] = sum_{co,k} gz[b,co,ti-kt+pt,...] * w[co,ci,kt,kh,kw]: every output gradient scattered to its taps */
void ora_conv3d_dgrad(const float* gz, const float* w, float* gx, int B, int Ci, int T, int H, int W, int Co, int pt, int ph) {
  const int To = T + 2 * pt - 2, Ho = H + 2 * ph - 2, Wo = W + 2 * ph - 2;
  for (int b = 0; b < B; ++b)
    for (int ci = 0; ci < Ci; ++ci)
      for (int ti = 0; ti < T; ++ti)
        for (int hi = 0; hi < H; ++hi)
          for (int wi = 0; wi < W; ++wi) {
            double s = 0.0;
            for (int co = 0; co < Co; ++co)
              for (int kt = 0; kt < 3; ++kt)
                for (int kh = 0; kh < 3; ++kh)
                  for (int kw = 0; kw < 3; ++kw) {
                    const int t = ti - kt + pt, h = hi - kh + ph, v = wi - kw + ph;
                    if (t < 0 || t >= To || h < 0 || h >= Ho || v < 0 || v >= Wo) continue;
                    s += (double)gz[(((long)(b * Co + co) * To + t) * Ho + h) * Wo + v] *
                         (double)w[(((long)(co * Ci + ci) * 3 + kt) * 3 + kh) * 3 + kw];
                  }
            gx[(((long)(b * Ci + ci) * T + ti) * H + hi) * W + wi] = (float)s;
          }
}

/* dw[co,ci,kt,kh,kw] = sum_{b,t,h,v} gz[b,co,t,h,v] * x[b,ci,t+kt-pt,h+kh-ph,v+kw-ph];  db[co] = sum gz */
void ora_conv3d_wgrad(const float* x, const float* gz, float* dw, float* db, int B, int Ci, int T, int H, int W, int Co, int pt,
                      int ph) {
  const int To = T + 2 * pt - 2, Ho = H + 2 * ph - 2, Wo = W + 2 * ph - 2;
  for (int co = 0; co < Co; ++co) {
    double sb = 0.0;
    for (int b = 0; b < B; ++b)
      for (long i = 0; i < (long)To * Ho * Wo; ++i) sb += (double)gz[(long)(b * Co + co) * To * Ho * Wo + i];
    if (db) db[co] = (float)sb;
    for (int ci = 0; ci < Ci; ++ci)
      for (int kt = 0; kt < 3; ++kt)
        for (int kh = 0; kh < 3; ++kh)
          for (int kw = 0; kw < 3; ++kw) {
            double s = 0.0;
            for (int b = 0; b < B; ++b)
              for (int t = 0; t < To; ++t)
                for (int h = 0; h < Ho; ++h)
                  for (int v = 0; v < Wo; ++v) {
                    const int ti = t + kt - pt, hi = h + kh - ph, wi = v + kw - ph;
                    if (ti < 0 || ti >= T || hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
                    s += (double)gz[(((long)(b * Co + co) * To + t) * Ho + h) * Wo + v] *
                         (double)x[(((long)(b * Ci + ci) * T + ti) * H + hi) * W + wi];
                  }
            dw[(((long)(co * Ci + ci) * 3 + kt) * 3 + kh) * 3 + kw] = (float)s;
          }
  }
}

/* gy [B,O] (gradient w.r.t. the PRE-activation), x [B,I], w [O,I] -> gx [B,I], dw [O,I], db [O] */
void ora_linear_bwd(const float* gy, const float* x, const float* w, float* gx, float* dw, float* db, int B, long I, int O) {
  for (int b = 0; b < B; ++b)
    for (long i = 0; i < I; ++i) {
      double s = 0.0;
      for (int o = 0; o < O; ++o) s += (double)gy[(long)b * O + o] * (double)w[o * I + i];
      gx[b * I + i] = (float)s;
    }
  for (int o = 0; o < O; ++o) {
    double sb = 0.0;
    for (int b = 0; b < B; ++b) sb += (double)gy[(long)b * O + o];
    db[o] = (float)sb;
    for (long i = 0; i < I; ++i) {
      double s = 0.0;
      for (int b = 0; b < B; ++b) s += (double)gy[(long)b * O + o] * (double)x[b * I + i];
      dw[o * I + i] = (float)s;
    }
  }
}

/* x [P,T,H,W] (P = B*C planes) -> y, idx [P,T,Ho,Wo] with Ho = (H-1)/2+1, Wo = (W-1)/2+1: window 3x3x3, stride (1,2,2),
 * padding 1 (padded cells never win); idx = flat index inside the plane's T*H*W of the FIRST maximum in (t,h,w) scan
 * order, a NaN always replaces the running maximum (aten max_pool3d_with_indices) */
void ora_maxpool3d(const float* x, float* y, long* idx, long P, int T, int H, int W) {
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  for (long p = 0; p < P; ++p)
    for (int t = 0; t < T; ++t)
      for (int h = 0; h < Ho; ++h)
        for (int v = 0; v < Wo; ++v) {
          float best = -INFINITY;
          long arg = -1;
          for (int ti = t - 1; ti <= t + 1; ++ti)
            for (int hi = 2 * h - 1; hi <= 2 * h + 1; ++hi)
              for (int wi = 2 * v - 1; wi <= 2 * v + 1; ++wi) {
                if (ti < 0 || ti >= T || hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
                const long o = ((long)ti * H + hi) * W + wi;
                const float c = x[p * T * H * W + o];
                if (arg < 0 || c > best || isnan(c)) { best = c; arg = o; }
              }
          const long q = ((p * T + t) * Ho + h) * Wo + v;
          y[q] = best;
          idx[q] = arg;
        }
}

/* gx (zero-initialised here) [P,T,H,W] += gy routed to the arg-max of every window */
void ora_maxpool3d_bwd(const float* gy, const long* idx, float* gx, long P, int T, int H, int W) {
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long plane = (long)T * H * W, oplane = (long)T * Ho * Wo;
  for (long i = 0; i < P * plane; ++i) gx[i] = 0.f;
  for (long p = 0; p < P; ++p)
    for (long q = 0; q < oplane; ++q) gx[p * plane + idx[p * oplane + q]] += gy[p * oplane + q];
}

/* one Adam step in place (step = 1, 2, ...), fp32 arithmetic in torch's order of operations */
void ora_adam_step(float* p, const float* g, float* m, float* v, long n, float lr, float b1, float b2, float eps, int step) {
  const double bc1 = 1.0 - pow((double)b1, step), bc2 = 1.0 - pow((double)b2, step);
  const float step_size = (float)((double)lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  for (long i = 0; i < n; ++i) {
    m[i] = m[i] + (g[i] - m[i]) * (1.f - b1);        /* exp_avg.lerp_(grad, 1 - beta1) */
    v[i] = v[i] * b2 + (1.f - b2) * g[i] * g[i];     /* exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2) */
    const float denom = sqrtf(v[i]) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (m[i] / denom);
  }
}
