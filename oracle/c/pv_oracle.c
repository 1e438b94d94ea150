/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the step's arithmetic, independent of torch.
 *
 * A second, library-free statement of what the reference computes, used by tests/ to cross-check the
 * torch-based oracle (oracle/conv3d_oracle.py) on small cases: it pins the MEANING of the operators the
 * reference delegates to torch (cross-correlation orientation, NCDHW flatten order, Linear layout).
 * Scalar loops, double accumulation, float results.  Never linked into the product.
 *
 *   ora_sat_normalise : predict_pv_yield/netcdf_dataset.py:96-101
 *   ora_conv3d_relu   : nn.Conv3d(k=3, padding=0) + F.relu, predict_pv_yield/models/conv3d/model.py:80-90,117-120
 *   ora_linear        : nn.Linear (+ optional ReLU), model.py:92-103,125-152
 *   ora_l1_loss       : (y_hat - y).abs().mean(), predict_pv_yield/models/base_model.py:99
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
