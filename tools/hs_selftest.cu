// Standalone check of the half-split tcgen05 kernels (hs_gemm.cuh) against a CPU double reference (GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o /tmp/hs_selftest tools/hs_selftest.cu -lcuda
//   /tmp/hs_selftest
// Exercises: absmax/split, weight images, gather GEMM (forward, multi-segment slots, stride-1 and stride-2
// dgrad, C % 64 != 0), multi-slot wgrad (1..8 slots, ragged pixel ranges, several splits).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cstring>

#include "../curvlinops_b200/csrc/hs_gemm.cuh"

using namespace curv;

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e__ = (x);                                                                 \
    if (e__ != cudaSuccess) {                                                              \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__);     \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

static uint32_t rng_state = 12345u;
static float frand() {  // U(-1, 1)
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xffff) / 32768.0f - 1.0f;
}

template <class T>
static T* dev(const std::vector<T>& h) {
  T* d;
  CK(cudaMalloc(&d, h.size() * sizeof(T) + 256));
  CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

// value gathered for GEMM row m, reduction index r
static double gather_ref(const Geom& g, const float* A, int m, int r) {
  const int tap = r / g.Cs, c = r % g.Cs;
  const int kh = tap / g.KW, kw = tap % g.KW;
  const int b = m / (g.Hd * g.Wd), rem = m % (g.Hd * g.Wd);
  const int hd = rem / g.Wd, wd = rem % g.Wd;
  int hs, ws;
  if (g.mode == 0) { hs = hd * g.sh - g.ph + kh; ws = wd * g.sw - g.pw + kw; }
  else {
    const int th = hd + g.ph - kh, tw = wd + g.pw - kw;
    if (th < 0 || tw < 0 || th % g.sh || tw % g.sw) return 0.0;
    hs = th / g.sh; ws = tw / g.sw;
  }
  if (hs < 0 || hs >= g.Hs || ws < 0 || ws >= g.Ws) return 0.0;
  return A[(((long long)b * g.Hs + hs) * g.Ws + ws) * g.Cs + c];
}

static int n_fail = 0;
static int g_debug = 0;
static bool g_tma = true;
// HS_PLANES=1: bf16 mode of the kernels (one bf16 plane per operand, one MMA per product).  The test data is rounded to
// bf16 on the host first, so the device-side rounding is exact and the same tight tolerances apply (x10 for the
// longer TMEM accumulation chains of that mode).
static int g_planes = 2;
static float tolx() { return g_planes == 1 ? 10.f : 1.f; }
static void q(std::vector<float>& v) {
  if (g_planes != 1) return;
  for (auto& x : v) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0x7fffu + ((u >> 16) & 1u)) & 0xffff0000u;  // round to nearest even bf16
    memcpy(&x, &u, 4);
  }
}

// slots: slot 0 primal, 1..K tangents.  a_has_slots / with_wt choose the segments.
static void test_gather(const char* name, Geom g, int K, int a_has_slots, int with_wt, int slot0, int accumulate,
                        int with_bias) {
  const int nA = a_has_slots ? 1 + K : 1;
  const long long a_elems = (long long)g.B * g.Hs * g.Ws * g.Cs;
  const long long w_elems = (long long)g.N * g.Kd;
  const long long o_elems = (long long)g.M * g.Nd;
  std::vector<float> A(a_elems * nA), W(w_elems * (1 + K)), out0(o_elems * (1 + K)), bias(g.Nd * (1 + K));
  for (int s = 0; s < nA; ++s) {
    const float sc = powf(10.f, -1.5f * s);  // very different magnitudes per slot
    for (long long i = 0; i < a_elems; ++i) A[s * a_elems + i] = frand() * sc * ((i % 97) == 0 ? 1e-4f : 1.f);
  }
  for (int s = 0; s <= K; ++s) {
    const float sc = 0.05f * powf(7.f, (float)(s % 3));
    for (long long i = 0; i < w_elems; ++i) W[s * w_elems + i] = frand() * sc;
  }
  for (auto& v : out0) v = frand();
  for (auto& v : bias) v = frand();
  q(A); q(W);
  float *dA = dev(A), *dW = dev(W), *dout = dev(out0), *dbias = dev(bias);
  const int nslots = 1 + K - slot0;
  // scales + planes + images
  std::vector<uint32_t> zero(64, 0);
  uint32_t *abits = dev(zero), *wbits = dev(zero);
  __half *Ah, *Al, *Wimg;
  CK(cudaMalloc(&Ah, a_elems * nA * 2 + 256)); CK(cudaMalloc(&Al, a_elems * nA * 2 + 256));
  const long long img = hs_image_halves(g.Nd, g.Kd, g_planes);
  CK(cudaMalloc(&Wimg, img * (1 + K) * 2 + 256));
  const int a_first = a_has_slots ? slot0 : 0;          // first A slot stored in the planes
  const int a_cnt = a_has_slots ? 1 + K - slot0 : 1;
  if (hs_launch_absmax(dA + a_first * a_elems, a_elems, a_elems, abits + a_first, a_cnt, 0)) exit(3);
  if (a_has_slots && slot0 > 0 && with_wt) {  // segment 2 gathers slot 0: not supported by the plane layout here
    printf("bad test config\n"); exit(3);
  }
  if (hs_launch_split(dA + a_first * a_elems, a_elems, a_elems, Ah, g_planes == 1 ? nullptr : Al, a_elems, abits + a_first, a_cnt, 0)) exit(3);
  if (hs_launch_absmax(dW, w_elems, w_elems, wbits, 1 + K, 0)) exit(3);
  if (hs_launch_pack_image(dW, w_elems, Wimg, img, g.N, g.Nd, g.Kd, 1 + K, wbits, 0, g_planes)) exit(3);
  HsGatherArgs a;
  memset(&a, 0, sizeof(a));
  a.planes = g_planes;
  a.g = g; a.Ah = Ah; a.Al = Al; a.A_slot = a_elems; a.a_slot_base = a_first; a.a_has_slots = a_has_slots;
  a.a_bits = abits; a.W_img = Wimg; a.Wt_img = with_wt ? Wimg + img : nullptr; a.Wt_img_slot = img;
  a.w_bits = wbits; a.bias = with_bias ? dbias : nullptr; a.bias_t = with_bias ? dbias + g.Nd : nullptr;
  a.bias_slot = g.Nd; a.out = dout; a.out_slot = o_elems; a.slot0 = slot0; a.accumulate = accumulate;
  int rc = hs_launch_gather_gemm(a, nslots, 0, true, g_tma ? a_cnt : 0);
  if (rc) { printf("%s: launch rc=%d\n", name, rc); exit(3); }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: KERNEL ERROR %s\n", name, cudaGetErrorString(e)); exit(4); }
  std::vector<float> got(out0.size());
  CK(cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost));
  double worst = 0;
  for (int slot = slot0; slot <= K; ++slot) {
    double maxref = 0, maxerr = 0;
    std::vector<double> row1(g.Kd), row2(g.Kd);
    const bool first_is_act = slot == 0 || a_has_slots;
    const bool second = slot > 0 && with_wt;
    for (int m = 0; m < g.M; ++m) {
      for (int r = 0; r < g.Kd; ++r) {
        row1[r] = first_is_act ? gather_ref(g, A.data() + (long long)slot * a_elems, m, r) : 0.0;
        row2[r] = second ? gather_ref(g, A.data(), m, r) : 0.0;
      }
      for (int n = 0; n < g.N; ++n) {
        double acc = 0;
        if (first_is_act)
          for (int r = 0; r < g.Kd; ++r) acc += row1[r] * W[(long long)n * g.Kd + r];
        if (second)
          for (int r = 0; r < g.Kd; ++r) acc += row2[r] * W[slot * w_elems + (long long)n * g.Kd + r];
        if (with_bias) acc += bias[slot * g.Nd + n];
        if (accumulate) acc += out0[slot * o_elems + (long long)m * g.Nd + n];
        const double gv = got[slot * o_elems + (long long)m * g.Nd + n];
        maxref = fmax(maxref, fabs(acc));
        maxerr = fmax(maxerr, fabs(acc - gv));
      }
    }
    worst = fmax(worst, maxerr / (maxref + 1e-300));
  }
  const bool ok = worst < 2e-6 * tolx();
  printf("%-44s M=%6d N=%4d Kd=%5d slots=%d  max rel err %.3e  %s\n", name, g.M, g.N, g.Kd, nslots, worst,
         ok ? "PASS" : "FAIL");
  if (!ok) ++n_fail;
  cudaFree(dA); cudaFree(dW); cudaFree(dout); cudaFree(dbias); cudaFree(abits); cudaFree(wbits);
  cudaFree(Ah); cudaFree(Al); cudaFree(Wimg);
}

static Geom conv_geom(int B, int H, int W, int Cin, int Cout, int k, int s, int p) {
  Geom g;
  g.B = B; g.Hs = H; g.Ws = W; g.Cs = Cin; g.Hd = (H + 2 * p - k) / s + 1; g.Wd = (W + 2 * p - k) / s + 1;
  g.KH = k; g.KW = k; g.sh = s; g.sw = s; g.ph = p; g.pw = p; g.mode = 0;
  g.N = Cout; g.Nd = Cout; g.Kd = k * k * Cin; g.M = B * g.Hd * g.Wd;
  return g;
}
static Geom dgrad_geom(const Geom& f) {  // source = grad of out, destination = grad of in
  Geom q = f;
  q.Hs = f.Hd; q.Ws = f.Wd; q.Cs = f.Nd; q.Hd = f.Hs; q.Wd = f.Ws; q.mode = 1;
  q.N = f.Cs; q.Nd = f.Cs; q.Kd = f.KH * f.KW * f.Nd; q.M = f.B * f.Hs * f.Ws;
  return q;
}

// N-stacked shared-activation kernel: out_k = gather(A_0) W_k^T for slots slot_lo .. slot_lo + nslots - 1
static void test_stack(const char* name, Geom g, int K, int slot_lo, int nslots, int accumulate, int with_bias) {
  const long long a_elems = (long long)g.B * g.Hs * g.Ws * g.Cs;
  const long long w_elems = (long long)g.N * g.Kd;
  const long long o_elems = (long long)g.M * g.Nd;
  std::vector<float> A(a_elems), W(w_elems * (1 + K)), out0(o_elems * (1 + K)), bias(g.Nd * (1 + K));
  for (auto& v : A) v = frand() * 2.f;
  for (int s = 0; s <= K; ++s) {
    const float sc = 0.05f * powf(9.f, (float)(s % 4)) * (s == 2 ? 1e-6f : 1.f);
    for (long long i = 0; i < w_elems; ++i) W[s * w_elems + i] = frand() * sc;
  }
  for (auto& v : out0) v = frand();
  for (auto& v : bias) v = frand();
  q(A); q(W);
  float *dA = dev(A), *dW = dev(W), *dout = dev(out0), *dbias = dev(bias);
  std::vector<uint32_t> zero(64, 0);
  uint32_t *abits = dev(zero), *wbits = dev(zero);
  __half *Ah, *Al, *Wimg;
  CK(cudaMalloc(&Ah, a_elems * 2 + 256)); CK(cudaMalloc(&Al, a_elems * 2 + 256));
  const long long img = hs_image_halves(g.Nd, g.Kd, g_planes);
  CK(cudaMalloc(&Wimg, img * (1 + K) * 2 + 256));
  if (hs_launch_absmax(dA, 0, a_elems, abits, 1, 0)) exit(3);
  if (hs_launch_split(dA, 0, a_elems, Ah, g_planes == 1 ? nullptr : Al, 0, abits, 1, 0)) exit(3);
  if (hs_launch_absmax(dW, w_elems, w_elems, wbits, 1 + K, 0)) exit(3);
  if (hs_launch_pack_image(dW, w_elems, Wimg, img, g.N, g.Nd, g.Kd, 1 + K, wbits, 0, g_planes)) exit(3);
  HsStackArgs a;
  memset(&a, 0, sizeof(a));
  a.planes = g_planes;
  a.g = g; a.Ah = Ah; a.Al = Al; a.a_bits = abits; a.W_img = Wimg; a.Wt_img = Wimg + img; a.Wt_img_slot = img;
  a.w_bits = wbits; a.bias = with_bias ? dbias : nullptr; a.bias_t = with_bias ? dbias + g.Nd : nullptr;
  a.bias_slot = g.Nd; a.out = dout; a.out_slot = o_elems; a.slot_lo = slot_lo; a.nslots = nslots;
  a.accumulate = accumulate;
  int rc = hs_launch_gather_stack(a, 0);
  if (rc) { printf("%s: launch rc=%d\n", name, rc); exit(3); }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: KERNEL ERROR %s\n", name, cudaGetErrorString(e)); exit(4); }
  std::vector<float> got(out0.size());
  CK(cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost));
  double worst = 0;
  std::vector<double> row(g.Kd);
  for (int slot = 0; slot <= K; ++slot) {
    const bool touched = slot >= slot_lo && slot < slot_lo + nslots;
    double maxref = 0, maxerr = 0;
    for (int m = 0; m < g.M; ++m) {
      for (int r = 0; r < g.Kd; ++r) row[r] = gather_ref(g, A.data(), m, r);
      for (int n = 0; n < g.N; ++n) {
        double acc = 0;
        if (touched) {
          for (int r = 0; r < g.Kd; ++r) acc += row[r] * W[slot * w_elems + (long long)n * g.Kd + r];
          if (with_bias) acc += bias[slot * g.Nd + n];
          if (accumulate) acc += out0[slot * o_elems + (long long)m * g.Nd + n];
        } else {
          acc = out0[slot * o_elems + (long long)m * g.Nd + n];  // untouched slots must stay as they were
        }
        const double gv = got[slot * o_elems + (long long)m * g.Nd + n];
        maxref = fmax(maxref, fabs(acc));
        maxerr = fmax(maxerr, fabs(acc - gv));
      }
    }
    worst = fmax(worst, maxerr / (maxref + 1e-300));
  }
  const bool ok = worst < 2e-6 * tolx();
  printf("%-44s M=%6d N=%4d Kd=%5d slots=%d..%d  max rel err %.3e  %s\n", name, g.M, g.N, g.Kd, slot_lo,
         slot_lo + nslots - 1, worst, ok ? "PASS" : "FAIL");
  if (!ok) ++n_fail;
  cudaFree(dA); cudaFree(dW); cudaFree(dout); cudaFree(dbias); cudaFree(abits); cudaFree(wbits);
  cudaFree(Ah); cudaFree(Al); cudaFree(Wimg);
}

static void test_wgrad(const char* name, Geom g, int NS, int slot0, int nsplit, double tol = 2e-6) {
  const long long i_elems = (long long)g.B * g.Hs * g.Ws * g.Cs;
  const int Ng = g.Nd;
  const long long g_elems = (long long)g.M * Ng;
  std::vector<float> In(i_elems), G(g_elems * NS);
  for (auto& v : In) v = frand() * 3.f;
  for (int s = 0; s < NS; ++s) {
    const float sc = powf(10.f, -2.f * s);
    for (long long i = 0; i < g_elems; ++i) G[s * g_elems + i] = frand() * sc;
  }
  q(In); q(G);
  tol *= tolx();
  float *dIn = dev(In), *dG = dev(G);
  std::vector<uint32_t> zero(64, 0);
  uint32_t *ibits = dev(zero), *gbits = dev(zero);
  __half *Ih, *Il, *Gh, *Gl;
  CK(cudaMalloc(&Ih, i_elems * 2 + 256)); CK(cudaMalloc(&Il, i_elems * 2 + 256));
  CK(cudaMalloc(&Gh, g_elems * NS * 2 + 256)); CK(cudaMalloc(&Gl, g_elems * NS * 2 + 256));
  if (hs_launch_absmax(dIn, 0, i_elems, ibits, 1, 0)) exit(3);
  if (hs_launch_split(dIn, 0, i_elems, Ih, g_planes == 1 ? nullptr : Il, 0, ibits, 1, 0)) exit(3);
  // G slots carry absolute slot indices slot0 .. slot0+NS-1 (their scale bits live at gbits[slot])
  if (hs_launch_absmax(dG, g_elems, g_elems, gbits + slot0, NS, 0)) exit(3);
  if (hs_launch_split(dG, g_elems, g_elems, Gh, g_planes == 1 ? nullptr : Gl, g_elems, gbits + slot0, NS, 0)) exit(3);
  HsWgradArgs a;
  memset(&a, 0, sizeof(a));
  a.planes = g_planes;
  a.g = g; a.Gh = Gh; a.Gl = Gl; a.G_slot = g_elems; a.Ng = Ng; a.g_bits = gbits; a.Ih = Ih; a.Il = Il;
  a.i_bits = ibits; a.nslots = NS; a.slot0 = slot0;
  a.m_per_split = (ceil_div(g.M, nsplit) + 15) / 16 * 16;
  a.nsplit = ceil_div(g.M, a.m_per_split);
  const long long w_elems = (long long)g.N * g.Kd;
  float* dpart;
  CK(cudaMalloc(&dpart, (size_t)a.nsplit * NS * w_elems * 4 + 256));
  CK(cudaMemset(dpart, 0xff, (size_t)a.nsplit * NS * w_elems * 4));
  a.partial = dpart;
  int rc = hs_launch_wgrad(a, 0, g_tma);
  if (rc) { printf("%s: launch rc=%d\n", name, rc); exit(3); }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: KERNEL ERROR %s\n", name, cudaGetErrorString(e)); exit(4); }
  std::vector<float> part((size_t)a.nsplit * NS * w_elems);
  CK(cudaMemcpy(part.data(), dpart, part.size() * 4, cudaMemcpyDeviceToHost));
  double worst = 0;
  std::vector<double> ref(w_elems);
  for (int s = 0; s < NS; ++s) {
    std::fill(ref.begin(), ref.end(), 0.0);
    for (int m = 0; m < g.M; ++m)
      for (int r = 0; r < g.Kd; ++r) {
        const double x = gather_ref(g, In.data(), m, r);
        if (x == 0.0) continue;
        for (int n = 0; n < g.N; ++n) ref[(long long)n * g.Kd + r] += x * G[s * g_elems + (long long)m * Ng + n];
      }
    double maxref = 0, maxerr = 0;
    for (long long i = 0; i < w_elems; ++i) {
      double gv = 0;
      for (int sp = 0; sp < a.nsplit; ++sp) gv += part[((size_t)sp * NS + s) * w_elems + i];
      maxref = fmax(maxref, fabs(ref[i]));
      maxerr = fmax(maxerr, fabs(ref[i] - gv));
    }
    worst = fmax(worst, maxerr / (maxref + 1e-300));
  }
  const bool ok = worst < tol;
  printf("%-44s M=%6d N=%4d Kd=%5d NS=%d splits=%d  max rel err %.3e  %s\n", name, g.M, g.N, g.Kd, NS, a.nsplit,
         worst, ok ? "PASS" : "FAIL");
  if (!ok) ++n_fail;
  cudaFree(dIn); cudaFree(dG); cudaFree(ibits); cudaFree(gbits); cudaFree(Ih); cudaFree(Il); cudaFree(Gh);
  cudaFree(Gl); cudaFree(dpart);
}

// timing of the kernels at ResNet-18 shapes (B = 128, K = 8); no reference, prints ms and algorithmic TFLOP/s
static int g_reps = 3;  // 0: one launch per kernel, no warm-up (ncu capture)
static void bench_layer(const char* name, int B, int H, int C, int Cout, int k, int stride) {
  const int K = 8, pad = k / 2;
  Geom f = conv_geom(B, H, H, C, Cout, k, stride, pad);
  Geom q = dgrad_geom(f);
  const long long a_elems = (long long)B * H * H * C, o_elems = (long long)f.M * Cout;
  float *dA, *dO, *dW, *dpart;
  CK(cudaMalloc(&dA, a_elems * (1 + K) * 4)); CK(cudaMalloc(&dO, o_elems * (1 + K) * 4));
  const long long w_elems = (long long)Cout * f.Kd;
  CK(cudaMalloc(&dW, w_elems * (1 + K) * 4));
  CK(cudaMemset(dA, 0, a_elems * (1 + K) * 4)); CK(cudaMemset(dO, 0, o_elems * (1 + K) * 4));
  CK(cudaMemset(dW, 0, w_elems * (1 + K) * 4));
  std::vector<uint32_t> zero(64, 0);
  uint32_t *abits = dev(zero), *wbits = dev(zero), *obits = dev(zero);
  __half *Ah, *Al, *Oh, *Ol, *Wimg, *Wtimg;
  CK(cudaMalloc(&Ah, a_elems * (1 + K) * 2)); CK(cudaMalloc(&Al, a_elems * (1 + K) * 2));
  CK(cudaMalloc(&Oh, o_elems * (1 + K) * 2)); CK(cudaMalloc(&Ol, o_elems * (1 + K) * 2));
  CK(cudaMemset(Ah, 0, a_elems * (1 + K) * 2)); CK(cudaMemset(Al, 0, a_elems * (1 + K) * 2));
  CK(cudaMemset(Oh, 0, o_elems * (1 + K) * 2)); CK(cudaMemset(Ol, 0, o_elems * (1 + K) * 2));
  const long long img = hs_image_halves(f.Nd, f.Kd, g_planes), timg = hs_image_halves(q.Nd, q.Kd, g_planes);
  CK(cudaMalloc(&Wimg, img * (1 + K) * 2)); CK(cudaMalloc(&Wtimg, timg * 2));
  CK(cudaMemset(Wimg, 0, img * (1 + K) * 2)); CK(cudaMemset(Wtimg, 0, timg * 2));
  const int nsplit = std::max(1, std::min(64, (2 * 148) / (ceil_div(f.Kd, 128) * ceil_div(Cout, 64))));
  const int mps = (ceil_div(f.M, nsplit) + 15) / 16 * 16;
  CK(cudaMalloc(&dpart, (size_t)ceil_div(f.M, mps) * K * w_elems * 4));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](auto fn) {
    if (g_reps > 0) { fn(); CK(cudaDeviceSynchronize()); }
    const int reps = g_reps > 0 ? g_reps : 1;
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) fn();
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
  };
  const double F = 2.0 * f.M * Cout * f.Kd;
  // split pass (absmax + split of 1+K input slots)
  float t_split = timeit([&] {
    hs_launch_absmax(dA, a_elems, a_elems, abits, 1 + K, 0);
    hs_launch_split(dA, a_elems, a_elems, Ah, g_planes == 1 ? nullptr : Al, a_elems, abits, 1 + K, 0);
  });
  HsGatherArgs a;
  memset(&a, 0, sizeof(a));
  a.planes = g_planes;
  a.g = f; a.Ah = Ah; a.Al = Al; a.A_slot = a_elems; a.a_has_slots = 1; a.a_bits = abits; a.W_img = Wimg;
  a.Wt_img = Wimg + img; a.Wt_img_slot = img; a.w_bits = wbits; a.out = dO; a.out_slot = o_elems; a.debug = g_debug;
  float t_fwd = timeit([&] { hs_launch_gather_gemm(a, 1 + K, 0, true, g_tma ? 1 + K : 0); });
  HsGatherArgs d;
  memset(&d, 0, sizeof(d));
  d.planes = g_planes;
  d.g = q; d.Ah = Oh; d.Al = Ol; d.A_slot = o_elems; d.a_slot_base = 1; d.a_has_slots = 1; d.a_bits = obits;
  d.W_img = Wtimg; d.w_bits = wbits; d.out = dA; d.out_slot = a_elems; d.slot0 = 1; d.debug = g_debug;
  float t_dgr = timeit([&] { hs_launch_gather_gemm(d, K, 0, true, g_tma ? K : 0); });
  HsWgradArgs w;
  memset(&w, 0, sizeof(w));
  w.planes = g_planes;
  w.g = f; w.Gh = Oh; w.Gl = Ol; w.G_slot = o_elems; w.Ng = Cout; w.g_bits = obits; w.Ih = Ah; w.Il = Al;
  w.i_bits = abits; w.partial = dpart; w.nslots = K; w.slot0 = 1; w.m_per_split = mps; w.nsplit = ceil_div(f.M, mps); w.debug = g_debug;
  float t_wgr = timeit([&] { hs_launch_wgrad(w, 0, g_tma); });
  printf("%-28s split %6.3f ms | fwd(1+2K) %7.3f ms %6.1f TF/s | dgrad(K) %7.3f ms %6.1f TF/s | wgrad(K) %7.3f ms %6.1f TF/s\n",
         name, t_split, t_fwd, F * (1 + 2 * K) / t_fwd / 1e9, t_dgr, F * K / t_dgr / 1e9, t_wgr, F * K / t_wgr / 1e9);
  cudaFree(dA); cudaFree(dO); cudaFree(dW); cudaFree(dpart); cudaFree(abits); cudaFree(wbits); cudaFree(obits);
  cudaFree(Ah); cudaFree(Al); cudaFree(Oh); cudaFree(Ol); cudaFree(Wimg); cudaFree(Wtimg);
}

int main(int argc, char** argv) {
  const int which = argc > 1 ? atoi(argv[1]) : 3;
  g_debug = argc > 2 ? atoi(argv[2]) : 0;
  g_tma = argc > 3 ? atoi(argv[3]) != 0 : true;
  if (argc > 3) g_hs_gather_tma_mode = atoi(argv[3]);
  if (getenv("HS_PLANES") && atoi(getenv("HS_PLANES")) == 1) { g_planes = 1; printf("bf16 mode: one plane, one MMA per product\n"); }
  if (g_debug) printf("debug knob = %d (timings only, results invalid)\n", g_debug);
  if (hs_ready() <= 0) { printf("half-split kernels unavailable on this device\n"); return 1; }
  if (which & 1) {
    // forward, Cin = 64 (fast path), BN = 64: primal + 2 tangents with both segments
    test_gather("fwd 3x3 s1 C64->64, K=2, both segs, bias", conv_geom(2, 12, 12, 64, 64, 3, 1, 1), 2, 1, 1, 0, 0, 1);
    // BN = 128, two n tiles, tangents only through the weights (input without tangent)
    test_gather("fwd 3x3 s2 C64->256, K=3, Wt only", conv_geom(3, 13, 13, 64, 256, 3, 2, 1), 3, 0, 1, 0, 0, 0);
    // BN = 128 single tile, 1x1 stride 2 (downsample), accumulate
    test_gather("fwd 1x1 s2 C128->128, K=2, acc", conv_geom(2, 14, 14, 128, 128, 1, 2, 0), 2, 1, 1, 0, 1, 0);
    // generic chunk path: Cin = 24 (not a multiple of 64), N = 40 (ragged)
    test_gather("fwd 3x3 s1 C24->40, K=1 (generic taps)", conv_geom(2, 9, 9, 24, 40, 3, 1, 1), 1, 1, 1, 0, 0, 1);
    // long reduction (odd number of stages), one slot
    test_gather("fwd 3x3 s1 C192->64, primal only", conv_geom(1, 10, 10, 192, 64, 3, 1, 1), 0, 0, 0, 0, 0, 0);
    // dgrad stride 1 (linear addressing), slots 1..3
    test_gather("dgrad 3x3 s1 C64<-128, slots 1..3", dgrad_geom(conv_geom(2, 12, 12, 64, 128, 3, 1, 1)), 3, 1, 0, 1, 0, 0);
    // dgrad stride 2 (generic addressing), accumulate
    test_gather("dgrad 3x3 s2 C64<-128, slots 1..2, acc", dgrad_geom(conv_geom(2, 14, 14, 64, 128, 3, 2, 1)), 2, 1, 0, 1, 1, 0);
    test_gather("dgrad 1x1 s2 C64<-128, slots 1..2", dgrad_geom(conv_geom(2, 14, 14, 64, 128, 1, 2, 0)), 2, 1, 0, 1, 0, 0);
    test_gather("dgrad 3x3 s2 C24<-40 (generic strided)", dgrad_geom(conv_geom(2, 11, 11, 24, 40, 3, 2, 1)), 2, 1, 0, 1, 0, 0);
    test_gather("dgrad 3x3 s2 p1 C64<-64 odd size, parity", dgrad_geom(conv_geom(2, 15, 13, 64, 64, 3, 2, 1)), 2, 1, 0, 1, 1, 0);
    // many tiles: persistent loop with more tiles than SMs
    test_gather("fwd 3x3 s1 C64->64, K=8, big M", conv_geom(8, 28, 28, 64, 64, 3, 1, 1), 8, 1, 1, 0, 0, 0);
  }
  if (which & 32) {
    test_stack("stack 7x7 s2 C8->64 slots 0..8 (stem)", conv_geom(2, 30, 30, 8, 64, 7, 2, 3), 8, 0, 9, 0, 1);
    test_stack("stack 3x3 s1 C64->64 slots 1..8 acc", conv_geom(2, 12, 12, 64, 64, 3, 1, 1), 8, 1, 8, 1, 0);
    test_stack("stack 3x3 s2 C64->256 slots 1..5", conv_geom(3, 13, 13, 64, 256, 3, 2, 1), 5, 1, 5, 0, 0);
    test_stack("stack 3x3 s1 C24->40 slots 0..2 bias", conv_geom(2, 9, 9, 24, 40, 3, 1, 1), 2, 0, 3, 0, 1);
    test_stack("stack 3x3 s1 C128->64 slots 2..3 long K", conv_geom(1, 20, 20, 128, 64, 3, 1, 1), 4, 2, 2, 1, 0);
    test_stack("stack 3x3 s1 C64->64 slots 1..8 big M", conv_geom(8, 28, 28, 64, 64, 3, 1, 1), 8, 1, 8, 0, 0);
  }
  if (which & 2) {
    test_wgrad("wgrad 3x3 s1 C64->64 NS=8", conv_geom(2, 12, 12, 64, 64, 3, 1, 1), 8, 1, 2);
    test_wgrad("wgrad 3x3 s1 C64->128 NS=3", conv_geom(2, 12, 12, 64, 128, 3, 1, 1), 3, 1, 3);
    test_wgrad("wgrad 3x3 s2 C128->128 NS=5", conv_geom(3, 13, 13, 128, 128, 3, 2, 1), 5, 1, 1);
    test_wgrad("wgrad 1x1 s2 C64->128 NS=4", conv_geom(2, 14, 14, 64, 128, 1, 2, 0), 4, 0, 2);
    test_wgrad("wgrad 3x3 s1 C24->40 NS=1 (ragged)", conv_geom(2, 9, 9, 24, 40, 3, 1, 1), 1, 1, 2);
    test_wgrad("wgrad 3x3 s1 C64->64 NS=8 long (flush)", conv_geom(8, 36, 36, 64, 64, 3, 1, 1), 8, 1, 2, 5e-5);
    test_wgrad("wgrad same, 2048-pixel splits (engine plan)", conv_geom(8, 36, 36, 64, 64, 3, 1, 1), 8, 1, 6, 1.5e-5);
  }
  if (which & 16) test_wgrad("wgrad 3x3 s1 C64->64 NS=8 long (flush)", conv_geom(8, 36, 36, 64, 64, 3, 1, 1), 8, 1, 2, 5e-5);
    test_wgrad("wgrad same, 2048-pixel splits (engine plan)", conv_geom(8, 36, 36, 64, 64, 3, 1, 1), 8, 1, 6, 1.5e-5);
  if (which & 8) {  // ncu capture: one launch per kernel
    g_reps = 0;
    bench_layer("layer1 3x3 C64->64 @56", 128, 56, 64, 64, 3, 1);
    bench_layer("layer3 3x3 C256->256 @14", 128, 14, 256, 256, 3, 1);
  }
  if (which & 64) {  // N-stacked kernel at the stem / layer1 shapes (zero data)
    auto bench_stack = [&](const char* name, int B, int H, int C, int Cout, int k, int s, int slot_lo, int ns) {
      const int K = 8;
      Geom f = conv_geom(B, H, H, C, Cout, k, s, k / 2);
      const long long a_elems = (long long)B * H * H * C, o_elems = (long long)f.M * Cout;
      const long long img = hs_image_halves(f.Nd, f.Kd, g_planes);
      __half *Ah, *Al, *Wimg; float* dO;
      CK(cudaMalloc(&Ah, a_elems * 2)); CK(cudaMalloc(&Al, a_elems * 2)); CK(cudaMalloc(&Wimg, img * (1 + K) * 2));
      CK(cudaMalloc(&dO, o_elems * (1 + K) * 4));
      CK(cudaMemset(Ah, 0, a_elems * 2)); CK(cudaMemset(Al, 0, a_elems * 2)); CK(cudaMemset(Wimg, 0, img * (1 + K) * 2));
      std::vector<uint32_t> zero(64, 0);
      uint32_t *abits = dev(zero), *wbits = dev(zero);
      HsStackArgs a;
      memset(&a, 0, sizeof(a));
      a.planes = g_planes;
      a.g = f; a.Ah = Ah; a.Al = Al; a.a_bits = abits; a.W_img = Wimg; a.Wt_img = Wimg + img; a.Wt_img_slot = img;
      a.w_bits = wbits; a.out = dO; a.out_slot = o_elems; a.slot_lo = slot_lo; a.nslots = ns;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      hs_launch_gather_stack(a, 0); CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      for (int i = 0; i < 3; ++i) hs_launch_gather_stack(a, 0);
      cudaEventRecord(e1); CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
      const double F = 2.0 * f.M * Cout * (double)(k * k * C) * ns;
      printf("%-34s %7.3f ms  %6.1f TF/s executed\n", name, ms, F / ms / 1e9);
      cudaFree(Ah); cudaFree(Al); cudaFree(Wimg); cudaFree(dO); cudaFree(abits); cudaFree(wbits);
    };
    bench_stack("stem 7x7 s2 C8->64 @224, 9 slots", 128, 224, 8, 64, 7, 2, 0, 9);
    bench_stack("layer1 3x3 C64->64 @56, slots 1..8", 128, 56, 64, 64, 3, 1, 1, 8);
    bench_stack("layer2 3x3 C128->128 @28, slots 1..8", 128, 28, 128, 128, 3, 1, 1, 8);
    bench_stack("layer3 3x3 C256->256 @14, slots 1..8", 128, 14, 256, 256, 3, 1, 1, 8);
  }
  if (which & 4) {
    bench_layer("layer1 3x3 C64->64 @56", 128, 56, 64, 64, 3, 1);
    bench_layer("layer2 3x3 C128->128 @28", 128, 28, 128, 128, 3, 1);
    bench_layer("layer3 3x3 C256->256 @14", 128, 14, 256, 256, 3, 1);
    bench_layer("layer4 3x3 C512->512 @7", 128, 7, 512, 512, 3, 1);
    bench_layer("layer2.0 3x3 s2 C64->128 @56", 128, 56, 64, 128, 3, 2);
  }
  printf(n_fail ? "FAILED: %d\n" : "ALL PASS\n", n_fail);
  return n_fail ? 1 : 0;
}
