// Two-sided Kronecker-factor apply on the tcgen05 contraction kernel (included at the end of engine.cu):
//     Y[a][B][z] = sum_{a', b} G[a][a'] X[a'][b][z] A[B][b]             (kronecker.py:141-153 'abZ,Aa,Bb->ABZ')
// for X, Y of shape [d_out, d_in, K] (K minor) -- the per-layer block of KFAC / its damped inverse, and (with the
// eigenvector matrices as factors) the two rotations of the EKFAC apply (eigh.py:98-104).
//
// wgrad_gemm_hs computes D[i][n] = sum_m In[m][i] * Gm[m][n], i.e. In^T Gm with BOTH operands stored reduction-row
// major (MN-major tcgen05 descriptors), and stores D transposed ([n][i]).  Written with the TRANSPOSED factors
// Gt = G^T, At = A^T (symmetric Kronecker factors and their inverses are their own transposes) the apply is two
// such products with no data transposition in between:
//   step 1   In = Gt [d_out x d_out],  Gm = X [d_out x (d_in K)]      ->  stored [(b,z)][a]  = T^T,  T = G X
//   step 2   In = At [d_in x d_in],    Gm = T^T rows (b, z), slot z   ->  stored [z][a][B]   = Y_z = T_z A^T
// and the split-K finish kernel of the weight gradients writes Y K-minor.  Operands are fp16 hi/lo planes (fp32-grade,
// three MMAs per product) or one bf16 plane (bf16 operators).  The "slots" of the kernel are column blocks of X in
// step 1 (N = 256 MMAs) and the K columns in step 2.

namespace curv {

// max |x| over n floats -> bits[0..copies) (bit pattern; bits must be zero before)
__global__ void __launch_bounds__(256) kron_absmax_kernel(const float* __restrict__ x, long long n,
                                                         uint32_t* __restrict__ bits, int copies) {
  float m = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(__ldg(x + i)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float wm[8];
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = fmaxf(m, wm[w]);
    if (m > 0.f)
      for (int c = 0; c < copies; ++c) atomicMax(bits + c, __float_as_uint(m));
  }
}

// src [rows][cols] fp32 (row stride ld_src) -> planes [rows][ld] (ld % 8 == 0, zero padded), scale from bits[0]
__global__ void __launch_bounds__(256) kron_split_pad_kernel(const float* __restrict__ src, long long ld_src, int rows,
                                                            int cols, __half* __restrict__ hi, __half* __restrict__ lo,
                                                            int ld, const uint32_t* __restrict__ bits) {
  const float sc = hs_pow2(hs_shift_from_bits(bits[0]));
  const int chunks = ld >> 3;
  const long long total = (long long)rows * chunks;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(e % chunks);
    const long long r = e / chunks;
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = ch * 8 + j < cols ? __ldg(src + r * ld_src + ch * 8 + j) : 0.f;
    const float4 v0 = make_float4(x[0], x[1], x[2], x[3]), v1 = make_float4(x[4], x[5], x[6], x[7]);
    if (lo == nullptr) { reinterpret_cast<uint4*>(hi)[e] = hs_bf16x8(v0, v1, sc); continue; }
    uint4 h, l;
    hs_split8(v0, v1, sc, h, l);
    reinterpret_cast<uint4*>(hi)[e] = h;
    reinterpret_cast<uint4*>(lo)[e] = l;
  }
}

// out[i] = sum_s part[s * n + i], fixed order (split-K partials of step 1)
__global__ void __launch_bounds__(256) kron_sum_splits_kernel(const float* __restrict__ part, int nsplit, long long n,
                                                             float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) s += __ldg(part + sp * n + i);
    out[i] = s;
  }
}

}  // namespace curv

struct KronTcPlan {
  int ldG, ldA, ldX, W8, NS1, nsplit1, mps1, nsplit2, mps2;
  long long part1_elems, off_P1;
  long long halves_G, halves_A, halves_X, halves_T;   // per plane, incl. slack
  long long t_elems, part2_elems;                     // floats
  long long off_bits, off_G, off_A, off_X, off_T, off_Tp, off_P2, total_bytes;
};
static KronTcPlan kron_tc_plan(int d_out, int d_in, int K, int planes) {
  KronTcPlan p;
  p.ldG = pad8(d_out); p.ldA = pad8(d_in); p.ldX = pad8(d_in * K);
  p.W8 = 64 * ceil_div(p.ldX, 512);
  p.NS1 = ceil_div(p.ldX, p.W8);
  auto slack = [](long long h) { return align_up(h + 8192, 128); };
  p.halves_G = slack((long long)d_out * p.ldG);
  p.halves_A = slack((long long)d_in * p.ldA);
  p.halves_X = slack((long long)d_out * p.ldX);
  p.t_elems = align_up((long long)p.NS1 * p.W8 * p.ldG + 1024, 64);  // T^T: rows (b, z) [+ block padding], ld = ldG
  {  // step 1 reduces over d_out rows: short chains there too (see below), partials summed by kron_sum_splits_kernel
    long long ns1 = planes == 1 ? 1 : ceil_div(d_out, 64);
    const long long cap1 = std::max<long long>(1, (1LL << 27) / p.t_elems);
    if (ns1 > cap1) ns1 = cap1;
    p.mps1 = (int)((ceil_div(d_out, (int)ns1) + 15) / 16 * 16);
    p.nsplit1 = ceil_div(d_out, p.mps1);
    p.part1_elems = p.nsplit1 > 1 ? (long long)p.nsplit1 * p.t_elems : 0;
  }
  p.halves_T = slack(p.t_elems);
  const long long tiles = (long long)ceil_div(p.ldA, 128) * ceil_div(p.ldG, 64);
  long long want = (2 * 148 + tiles - 1) / tiles;
  // Short accumulation chains: the tensor core adds into its fp32 accumulator with truncation, a bias of ~0.25 ulp
  // per MMA that does not average out, and Kronecker factors make it visible -- the rows of a softmax layer's gradient
  // covariance sum to zero, so G X A^T cancels to ~1/30 of its terms (measured on the fc block of ResNet-18: 1.3e-4
  // of the result with 171-row chains = 33 MMAs; strict-fp32 FMA accumulation: 2e-6).  64 rows = 12 MMAs per chain;
  // the extra split-K partials are a few GB/s-milliseconds on these small matrices (capped at 1 GiB of scratch).
  const long long by_len = ceil_div(d_in, planes == 1 ? 4096 : 64);
  long long ns = want > by_len ? want : by_len;
  const long long maxsplit = ceil_div(d_in, planes == 1 ? 256 : 32);
  const long long cap = std::max<long long>(1, (1LL << 28) / ((long long)K * d_out * p.ldA));
  if (ns > maxsplit) ns = maxsplit;
  if (ns > cap) ns = cap;
  if (ns < 1) ns = 1;
  p.mps2 = (int)((ceil_div(d_in, (int)ns) + 15) / 16 * 16);
  p.nsplit2 = ceil_div(d_in, p.mps2);
  p.part2_elems = (long long)p.nsplit2 * K * d_out * p.ldA + 64;
  long long o = 0;
  auto take = [&](long long bytes) { long long r = o; o = align_up(o + bytes, 1024); return r; };
  p.off_bits = take(64 * 4);
  p.off_G = take(p.halves_G * 2 * planes);
  p.off_A = take(p.halves_A * 2 * planes);
  p.off_X = take(p.halves_X * 2 * planes);
  p.off_T = take(p.t_elems * 4);
  p.off_Tp = take(p.halves_T * 2 * planes);
  p.off_P2 = take(p.part2_elems * 4);
  p.off_P1 = take(p.part1_elems * 4);
  p.total_bytes = o;
  return p;
}

extern "C" size_t curv_kron_apply_tc_workspace(int d_out, int d_in, int K, int planes) {
  if (d_out < 1 || d_in < 1 || K < 1 || K > 8) return 0;
  return (size_t)kron_tc_plan(d_out, d_in, K, planes == 1 ? 1 : 2).total_bytes;
}

extern "C" int curv_kron_apply_tc(const float* Gt, const float* At, int d_out, int d_in, int K, const float* X,
                                  float* Y, int planes, void* ws, size_t ws_bytes, void* stream) {
  if (!Gt || !At || !X || !Y || d_out < 1 || d_in < 1 || K < 1 || K > 8)
    return fail(CURV_ERR_INVALID, "curv_kron_apply_tc: bad arguments (two factors, 1 <= K <= 8)");
  planes = planes == 1 ? 1 : 2;
  if (hs_ready() <= 0) return fail(CURV_ERR_CUDA, "no CUDA device / tcgen05 kernels unavailable: curvb200 has no CPU fallback");
  const KronTcPlan p = kron_tc_plan(d_out, d_in, K, planes);
  if (!ws || ws_bytes < (size_t)p.total_bytes) return fail(CURV_ERR_WORKSPACE, "curv_kron_apply_tc: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  char* base = (char*)ws;
  uint32_t* bits = (uint32_t*)(base + p.off_bits);  // [0] G, [8] A, [16..24) X, [32..40) T
  auto hi = [&](long long off) { return (__half*)(base + off); };
  auto lo = [&](long long off, long long halves) { return planes == 1 ? (__half*)nullptr : (__half*)(base + off) + halves; };
  CHECK_CUDA(cudaMemsetAsync(bits, 0, 64 * 4, st));
  // slack regions of the planes are read (never used) by blocks that reach past a row end: keep them finite
  CHECK_CUDA(cudaMemsetAsync(base + p.off_G, 0, (size_t)(p.off_T - p.off_G), st));
  CHECK_CUDA(cudaMemsetAsync(base + p.off_Tp, 0, (size_t)(p.off_P2 - p.off_Tp), st));
  CHECK_CUDA(cudaMemsetAsync(Y, 0, sizeof(float) * (size_t)d_out * d_in * K, st));
  if (planes == 2) {
    kron_absmax_kernel<<<grid1d((long long)d_out * d_out), 256, 0, st>>>(Gt, (long long)d_out * d_out, bits, 1);
    kron_absmax_kernel<<<grid1d((long long)d_in * d_in), 256, 0, st>>>(At, (long long)d_in * d_in, bits + 8, 1);
    kron_absmax_kernel<<<grid1d((long long)d_out * d_in * K), 256, 0, st>>>(X, (long long)d_out * d_in * K, bits + 16, 8);
    g_launches += 3;
  }
  kron_split_pad_kernel<<<grid1d((long long)d_out * (p.ldG / 8)), 256, 0, st>>>(Gt, d_out, d_out, d_out, hi(p.off_G),
                                                                             lo(p.off_G, p.halves_G), p.ldG, bits);
  kron_split_pad_kernel<<<grid1d((long long)d_in * (p.ldA / 8)), 256, 0, st>>>(At, d_in, d_in, d_in, hi(p.off_A),
                                                                            lo(p.off_A, p.halves_A), p.ldA, bits + 8);
  kron_split_pad_kernel<<<grid1d((long long)d_out * (p.ldX / 8)), 256, 0, st>>>(
      X, (long long)d_in * K, d_out, d_in * K, hi(p.off_X), lo(p.off_X, p.halves_X), p.ldX, bits + 16);
  g_launches += 3;
  LAUNCH_CHECK();
  // ---- step 1: T^T[(b,z)][a] = sum_a' X[a'][(b,z)] Gt[a'][a]
  float* Tt = (float*)(base + p.off_T);
  {
    HsWgradArgs a;
    memset(&a, 0, sizeof(a));
    Geom& g = a.g;
    g.B = d_out; g.Hs = g.Ws = g.Hd = g.Wd = 1; g.Cs = p.ldG; g.KH = g.KW = 1; g.sh = g.sw = 1;
    g.mode = 0; g.N = p.W8; g.Nd = p.W8; g.Kd = p.ldG; g.M = d_out;
    a.Gh = hi(p.off_X); a.Gl = lo(p.off_X, p.halves_X); a.G_slot = p.W8; a.G_ld = p.ldX; a.Ng = p.W8;
    a.g_bits = bits + 16; a.Ih = hi(p.off_G); a.Il = lo(p.off_G, p.halves_G); a.i_bits = bits;
    a.partial = p.nsplit1 > 1 ? (float*)(base + p.off_P1) : Tt;
    a.nsplit = p.nsplit1; a.nslots = p.NS1; a.slot0 = 0; a.m_per_split = p.mps1;
    a.planes = planes;
    ProfScope prof(1, 2.0 * d_out * (double)d_out * d_in * K, st);
    if (hs_launch_wgrad(a, st)) return fail(CURV_ERR_CUDA, "Kronecker apply: step-1 contraction launch failed");
    ++g_launches;
  }
  if (p.nsplit1 > 1) {
    const long long n1 = (long long)p.NS1 * p.W8 * p.ldG;  // one split's partial: [NS1][W8][ldG]
    kron_sum_splits_kernel<<<grid1d(n1), 256, 0, st>>>((const float*)(base + p.off_P1), p.nsplit1, n1, Tt);
    LAUNCH_CHECK();
  }
  const long long t_used = (long long)d_in * K * p.ldG;  // rows (b, z) of T^T that exist
  if (planes == 2) {
    kron_absmax_kernel<<<grid1d(t_used), 256, 0, st>>>(Tt, t_used, bits + 32, 8);
    ++g_launches;
  }
  kron_split_pad_kernel<<<grid1d((long long)d_in * K * (p.ldG / 8)), 256, 0, st>>>(
      Tt, p.ldG, d_in * K, p.ldG, hi(p.off_Tp), lo(p.off_Tp, p.halves_T), p.ldG, bits + 32);
  LAUNCH_CHECK();
  // ---- step 2: Y_z[a][B] = sum_b At[b][B] T^T[(b,z)][a]
  float* part = (float*)(base + p.off_P2);
  {
    HsWgradArgs a;
    memset(&a, 0, sizeof(a));
    Geom& g = a.g;
    g.B = d_in; g.Hs = g.Ws = g.Hd = g.Wd = 1; g.Cs = p.ldA; g.KH = g.KW = 1; g.sh = g.sw = 1;
    g.mode = 0; g.N = d_out; g.Nd = p.ldG; g.Kd = p.ldA; g.M = d_in;
    a.Gh = hi(p.off_Tp); a.Gl = lo(p.off_Tp, p.halves_T); a.G_slot = p.ldG; a.G_ld = (long long)K * p.ldG; a.Ng = p.ldG;
    a.g_bits = bits + 32; a.Ih = hi(p.off_A); a.Il = lo(p.off_A, p.halves_A); a.i_bits = bits + 8;
    a.partial = part; a.nsplit = p.nsplit2; a.nslots = K; a.slot0 = 0; a.m_per_split = p.mps2;
    a.planes = planes;
    ProfScope prof(1, 2.0 * d_out * (double)d_in * d_in * K, st);
    if (hs_launch_wgrad(a, st)) return fail(CURV_ERR_CUDA, "Kronecker apply: step-2 contraction launch failed");
    ++g_launches;
  }
  return launch_wgrad_finish(part, p.nsplit2, K, 0, d_out, d_in, p.ldA, 1, Y, 0, K, 0, 1.f,
                             (long long)d_out * p.ldA, st);
}
