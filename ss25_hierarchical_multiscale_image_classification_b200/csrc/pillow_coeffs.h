// Host-side restatement of Pillow's bilinear (triangle) resample coefficients for the
// integer scales the HiPAC path uses (P/224 in {2,4,8}).
//
// Third-party algorithm (not vendored in the reference): Pillow src/libImaging/Resample.c,
// precompute_coeffs() + normalize_coeffs_8bpc(); reached from the reference through
// transforms.Resize((224,224)) (src/main.py:814).  For an integer scale f every interior
// output pixel has the same 2f-tap window; only output 0 and output 223 are clamped to
// 3f/2 taps and renormalised.  build() recomputes the generic per-output tables exactly as
// Pillow does (double arithmetic, C truncation) and verifies that structure.
#pragma once
#include <math.h>
#include <stdint.h>
#include <vector>

namespace hipac {

constexpr int kPrecisionBits = 32 - 8 - 2;

struct PillowCoeffs {
  int scale = 0;
  int32_t interior[16];  // 2*scale taps, window starts at scale*i - scale/2
  int32_t left[12];      // 3*scale/2 taps, output 0, window starts at 0
  int32_t right[12];     // 3*scale/2 taps, output 223, window starts at 223*scale - scale/2
  bool ok = false;
};

inline void pillow_generic(int in_size, int out_size, std::vector<int>& xmin, std::vector<int>& cnt,
                           std::vector<std::vector<int32_t>>& kk) {
  double scale = (double)in_size / out_size;
  double filterscale = scale < 1.0 ? 1.0 : scale;
  double support = 1.0 * filterscale;
  xmin.assign(out_size, 0);
  cnt.assign(out_size, 0);
  kk.assign(out_size, {});
  double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; xx++) {
    double center = (xx + 0.5) * scale;
    int lo = (int)(center - support + 0.5);
    if (lo < 0) lo = 0;
    int hi = (int)(center + support + 0.5);
    if (hi > in_size) hi = in_size;
    int n = hi - lo;
    std::vector<double> w(n);
    double ww = 0.0;
    for (int x = 0; x < n; x++) {
      double a = (x + lo - center + 0.5) * ss;
      if (a < 0.0) a = -a;
      double v = a < 1.0 ? 1.0 - a : 0.0;
      w[x] = v;
      ww += v;
    }
    kk[xx].resize(n);
    for (int x = 0; x < n; x++) {
      double v = w[x];
      if (ww != 0.0) v /= ww;
      kk[xx][x] = v < 0 ? (int32_t)(-0.5 + v * (1 << kPrecisionBits)) : (int32_t)(0.5 + v * (1 << kPrecisionBits));
    }
    xmin[xx] = lo;
    cnt[xx] = n;
  }
}

inline PillowCoeffs build_pillow_coeffs(int scale) {
  PillowCoeffs c;
  c.scale = scale;
  if (scale != 2 && scale != 4 && scale != 8) return c;
  const int out = 224, in = 224 * scale;
  std::vector<int> xmin, cnt;
  std::vector<std::vector<int32_t>> kk;
  pillow_generic(in, out, xmin, cnt, kk);
  bool ok = true;
  const int ni = 2 * scale, ne = 3 * scale / 2;
  ok = ok && xmin[0] == 0 && cnt[0] == ne;
  ok = ok && xmin[out - 1] == (out - 1) * scale - scale / 2 && cnt[out - 1] == ne;
  for (int i = 1; i < out - 1 && ok; i++) {
    ok = ok && xmin[i] == scale * i - scale / 2 && cnt[i] == ni;
    for (int t = 0; t < ni && ok; t++) ok = ok && kk[i][t] == kk[1][t];
  }
  if (!ok) return c;
  for (int t = 0; t < 16; t++) c.interior[t] = t < ni ? kk[1][t] : 0;
  for (int t = 0; t < 12; t++) c.left[t] = t < ne ? kk[0][t] : 0;
  for (int t = 0; t < 12; t++) c.right[t] = t < ne ? kk[out - 1][t] : 0;
  c.ok = true;
  return c;
}

}  // namespace hipac
