// Host-only checks of the half-split path's planning code (no GPU needed; compiled with nvcc, run on the CPU):
//   * hs_make_parity: the parity-class decomposition of a strided dgrad must reproduce, for every destination
//     pixel, exactly the (tap, source pixel) pairs of the generic rule  source = (dest + pad - tap) / stride
//   * hs_shift_from_bits / hs_pow2: scaled maxima land in [2^14, 2^15), clamps, zero
#include <cstdio>
#include <cstdlib>
#include <set>
#include <tuple>

#include "../../curvlinops_b200/csrc/hs_gemm.cuh"

using namespace curv;

static int fails = 0;
#define EXPECT(c)                                                        \
  do {                                                                   \
    if (!(c)) { printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); ++fails; } \
  } while (0)

static Geom dgrad_geom(int B, int H, int W, int Cin, int Cout, int k, int s, int p) {
  Geom q;
  const int Ho = (H + 2 * p - k) / s + 1, Wo = (W + 2 * p - k) / s + 1;
  q.B = B; q.Hs = Ho; q.Ws = Wo; q.Cs = Cout; q.Hd = H; q.Wd = W;
  q.KH = k; q.KW = k; q.sh = s; q.sw = s; q.ph = p; q.pw = p; q.mode = 1;
  q.N = Cin; q.Nd = Cin; q.Kd = k * k * Cout; q.M = B * H * W;
  return q;
}

static void check_parity(const Geom& g) {
  HsParity par;
  const bool ok = hs_make_parity(g, par);
  EXPECT(ok);
  if (!ok) return;
  // reference: for each destination pixel the set of (tap, hs, ws)
  typedef std::tuple<int, int, int, int, int> Rec;  // hd, wd, tap, hs, ws
  std::set<Rec> ref, got;
  for (int hd = 0; hd < g.Hd; ++hd)
    for (int wd = 0; wd < g.Wd; ++wd)
      for (int kh = 0; kh < g.KH; ++kh)
        for (int kw = 0; kw < g.KW; ++kw) {
          const int th = hd + g.ph - kh, tw = wd + g.pw - kw;
          if (th < 0 || tw < 0 || th % g.sh || tw % g.sw) continue;
          const int hs = th / g.sh, ws = tw / g.sw;
          if (hs >= g.Hs || ws >= g.Ws) continue;
          ref.insert(Rec(hd, wd, kh * g.KW + kw, hs, ws));
        }
  long long pixels = 0;
  int tiles = 0;
  for (int c = 0; c < par.nclass; ++c) {
    EXPECT(par.tile0[c] == tiles);
    tiles += ceil_div(g.B * par.Hc[c] * par.Wc[c], TC_BM);
    pixels += (long long)par.Hc[c] * par.Wc[c];
    if (c > 0) EXPECT(par.ntap[c] <= par.ntap[c - 1]);  // heaviest first
    for (int i = 0; i < par.Hc[c]; ++i)
      for (int j = 0; j < par.Wc[c]; ++j)
        for (int t = 0; t < par.ntap[c]; ++t) {
          const int hs = i + par.dh[c][t], ws = j + par.dw[c][t];
          if (hs < 0 || hs >= g.Hs || ws < 0 || ws >= g.Ws) continue;
          got.insert(Rec(g.sh * i + par.oh[c], g.sw * j + par.ow[c], par.tap[c][t], hs, ws));
        }
  }
  EXPECT(par.tile0[par.nclass] == tiles);
  EXPECT(pixels == (long long)g.Hd * g.Wd);  // the classes partition the destination grid
  EXPECT(ref == got);
}

int main() {
  check_parity(dgrad_geom(2, 56, 56, 64, 128, 3, 2, 1));
  check_parity(dgrad_geom(1, 15, 13, 64, 64, 3, 2, 1));
  check_parity(dgrad_geom(1, 14, 14, 64, 128, 1, 2, 0));
  check_parity(dgrad_geom(1, 9, 11, 128, 64, 5, 2, 2));
  check_parity(dgrad_geom(1, 12, 12, 64, 64, 2, 2, 0));
  {  // not applicable: stride 1, C % 64 != 0, stride 3x3 = 9 classes
    HsParity par;
    EXPECT(!hs_make_parity(dgrad_geom(1, 8, 8, 64, 64, 3, 1, 1), par));
    EXPECT(!hs_make_parity(dgrad_geom(1, 8, 8, 64, 24, 3, 2, 1), par));
    EXPECT(!hs_make_parity(dgrad_geom(1, 9, 9, 64, 64, 3, 3, 1), par));
  }
  // scales
  EXPECT(hs_shift_from_bits(0u) == 0);
  const float vals[] = {1.0f, 0.75f, 3.1e-5f, 7.7e3f, 1e-30f, 1e30f, 65504.f};
  for (float v : vals) {
    union { float f; uint32_t u; } c;
    c.f = v;
    const int sh = hs_shift_from_bits(c.u);
    EXPECT(sh >= -HS_SH_CLAMP && sh <= HS_SH_CLAMP);
    const float scaled = v * hs_pow2(sh);
    if (sh > -HS_SH_CLAMP && sh < HS_SH_CLAMP) EXPECT(scaled >= 16384.f && scaled < 32768.f);
    EXPECT(hs_pow2(sh) * hs_pow2(-sh) == 1.0f);
    EXPECT(scaled < 65504.f || sh == -HS_SH_CLAMP);
  }
  printf(fails ? "FAILED: %d\n" : "ALL PASS\n", fails);
  return fails ? 1 : 0;
}
