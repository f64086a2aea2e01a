// Two-sided Kronecker-factor apply on the tcgen05 contraction kernels (included at the end of engine.cu):
//     Y[a][B][z] = sum_{a', b} G[a][a'] X[a'][b][z] A[B][b]             (kronecker.py:141-153 'abZ,Aa,Bb->ABZ')
// for X, Y of shape [d_out, d_in, K] (K minor) -- the per-layer block of KFAC / its damped inverse, and (with the
// eigenvector matrices as factors) the two rotations of the EKFAC apply (eigh.py:98-104).  Any square G, A.
//
// Both products are plain GEMMs  out[m][n] = sum_r Act[m][r] W[n][r]  = a 1x1 "convolution" of gather_gemm_hs
// (K-major operands, TMA-fed, TMEM accumulation in chunks of HS_FLUSH stages summed in registers with
// round-to-nearest adds -- which is what keeps the tensor core's truncating accumulator harmless here: Kronecker
// factors of a softmax layer make G X cancel to ~1/30 of its terms):
//   step 1   T_z = G X_z        Act = G (one shared slot),  W_z = X_z^T as weight images, one per column z
//   step 2   Y_z = T_z A^T      Act = T_z (slot z),         W   = A     as weight image
// Operands are fp16 hi/lo planes / images (fp32-grade, three MMAs per product) also for bf16 operators (bf16-rounded
// operands gave 4e-2 on a softmax layer's block).  The operand forms of the FACTORS (planes of G, image of A, their
// scale words) depend only on the factors: they live in a caller-owned buffer and are built once per operator.
// Round-2 history: a first version ran both products on the MN-major wgrad kernel with split-K partials; its
// accumulation chains had to be cut to 64 rows against the truncation bias, the drains then dominated (8.9 ms for the
// 21 ResNet-18 blocks against 2.0 ms for the cuBLAS-backed reference einsum).

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
    uint4 h, l;
    hs_split8(make_float4(x[0], x[1], x[2], x[3]), make_float4(x[4], x[5], x[6], x[7]), sc, h, l);
    reinterpret_cast<uint4*>(hi)[e] = h;
    reinterpret_cast<uint4*>(lo)[e] = l;
  }
}

// Weight image (layout of hs_pack_image_kernel: blocks (tn, kc) = [hi: BN rows x 64 halves, 128B-swizzled][lo]) of the
// matrix  Wm[n][k] = src[n * sn + k * sk]  (n < N, k < Kreal; zero beyond), one image per grid.y slot (src += slot *
// s_slot, dst += slot * dst_slot), scale bits[slot].  (sn, sk) = (ld, 1): row-major source; = (K, d_in K): X_z^T.
__global__ void __launch_bounds__(256) kron_pack_image_kernel(const float* __restrict__ src, long long sn, long long sk,
                                                             long long s_slot, __half* __restrict__ dst, long long dst_slot,
                                                             int N, int Kreal, int BN, int tiles_n, int nchunks,
                                                             const uint32_t* __restrict__ bits) {
  src += blockIdx.y * s_slot;
  dst += blockIdx.y * dst_slot;
  const float sc = hs_pow2(hs_shift_from_bits(bits[blockIdx.y]));
  const long long total = (long long)tiles_n * nchunks * BN * 8;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e & 7);
    long long t = e >> 3;
    const int r = (int)(t % BN); t /= BN;
    const int kc = (int)(t % nchunks);
    const int tn = (int)(t / nchunks);
    const int n = tn * BN + r, k = kc * HS_BK + c * 8;
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = (n < N && k + j < Kreal) ? __ldg(src + n * sn + (k + j) * sk) : 0.f;
    uint4 h, l;
    hs_split8(make_float4(x[0], x[1], x[2], x[3]), make_float4(x[4], x[5], x[6], x[7]), sc, h, l);
    __half* blk = dst + ((long long)tn * nchunks + kc) * (2 * BN * HS_BK);
    const int o = ((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) >> 1;
    *reinterpret_cast<uint4*>(blk + o) = h;
    *reinterpret_cast<uint4*>(blk + BN * HS_BK + o) = l;
  }
}

// Y[(a * d_in + b) * K + z] = Yz[z][a][b]  (row stride ld of Yz)
__global__ void kron_interleave_kernel(const float* __restrict__ Yz, long long slot, int ld, int d_out, int d_in, int K,
                                       float* __restrict__ Y) {
  const long long total = (long long)d_out * d_in * K;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int z = (int)(e % K);
    const long long ab = e / K;
    const int b = (int)(ab % d_in);
    const long long a = ab / d_in;
    Y[e] = __ldg(Yz + z * slot + a * ld + b);
  }
}

}  // namespace curv

static inline int pad64(int x) { return (x + 63) & ~63; }

struct KronTcPlan {
  int ldG, ldT, NdY;                // row strides: G planes (= reduction of step 1), T (= reduction of step 2), Y_z
  long long g_halves, a_img_halves;  // factor operands (per plane / whole image)
  long long x_img_halves;           // per column image of X_z^T
  long long t_elems, y_elems;       // per column
  long long f_bits, f_G, f_A, factor_bytes;
  long long off_bits, off_X, off_T, off_Tp, off_Y, total_bytes;
};
static KronTcPlan kron_tc_plan(int d_out, int d_in, int K) {
  KronTcPlan p;
  p.ldG = pad64(d_out); p.ldT = pad64(d_in); p.NdY = pad4(d_in);
  p.g_halves = align_up((long long)d_out * p.ldG + 8192, 128);
  p.a_img_halves = hs_image_halves(pad4(d_in), p.ldT);       // A as weights [N = d_in][Kd = ldT]
  p.x_img_halves = hs_image_halves(p.ldT, p.ldG);            // X_z^T as weights [N = d_in (rows up to ldT)][Kd = ldG]
  p.t_elems = (long long)d_out * p.ldT;
  p.y_elems = (long long)d_out * p.NdY;
  long long o = 0;
  auto take = [&](long long bytes) { long long r = o; o = align_up(o + bytes, 1024); return r; };
  p.f_bits = take(64 * 4);
  p.f_G = take(p.g_halves * 2 * 2);
  p.f_A = take(p.a_img_halves * 2);
  p.factor_bytes = o;
  o = 0;
  p.off_bits = take(64 * 4);
  p.off_X = take(p.x_img_halves * 2 * K);
  p.off_T = take(p.t_elems * 4 * K);
  p.off_Tp = take((align_up(p.t_elems * K + 8192, 128)) * 2 * 2);
  p.off_Y = take(p.y_elems * 4 * K);
  p.total_bytes = o;
  return p;
}

extern "C" size_t curv_kron_apply_tc_workspace(int d_out, int d_in, int K) {
  if (d_out < 1 || d_in < 1 || K < 1 || K > 8) return 0;
  return (size_t)kron_tc_plan(d_out, d_in, K).total_bytes;
}
extern "C" size_t curv_kron_apply_tc_factor_bytes(int d_out, int d_in) {
  if (d_out < 1 || d_in < 1) return 0;
  return (size_t)kron_tc_plan(d_out, d_in, 1).factor_bytes;
}

extern "C" int curv_kron_apply_tc(const float* G, const float* A, int d_out, int d_in, int K, const float* X,
                                  float* Y, void* factor_ws, size_t factor_bytes, int factors_ready, void* ws,
                                  size_t ws_bytes, void* stream) {
  if (!G || !A || !X || !Y || d_out < 1 || d_in < 1 || K < 1 || K > 8)
    return fail(CURV_ERR_INVALID, "curv_kron_apply_tc: bad arguments (two factors, 1 <= K <= 8)");
  if (hs_ready() <= 0)
    return fail(CURV_ERR_CUDA, "no CUDA device / tcgen05 kernels unavailable: curvb200 has no CPU fallback");
  const KronTcPlan p = kron_tc_plan(d_out, d_in, K);
  if (!ws || ws_bytes < (size_t)p.total_bytes || !factor_ws || factor_bytes < (size_t)p.factor_bytes)
    return fail(CURV_ERR_WORKSPACE, "curv_kron_apply_tc: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  char* fb = (char*)factor_ws;
  char* base = (char*)ws;
  uint32_t* fbits = (uint32_t*)(fb + p.f_bits);  // [0] G, [8] A
  __half* Gh = (__half*)(fb + p.f_G);
  __half* Gl = Gh + p.g_halves;
  __half* Aimg = (__half*)(fb + p.f_A);
  uint32_t* bits = (uint32_t*)(base + p.off_bits);  // [0..8) X columns, [16..24) T columns
  const int BN_T = tc_bn(p.ldT), BN_Y = tc_bn(pad4(d_in));
  if (!factors_ready) {  // operand forms of the factors: once per operator
    CHECK_CUDA(cudaMemsetAsync(fb, 0, (size_t)p.factor_bytes, st));
    kron_absmax_kernel<<<grid1d((long long)d_out * d_out), 256, 0, st>>>(G, (long long)d_out * d_out, fbits, 1);
    kron_absmax_kernel<<<grid1d((long long)d_in * d_in), 256, 0, st>>>(A, (long long)d_in * d_in, fbits + 8, 1);
    kron_split_pad_kernel<<<grid1d((long long)d_out * (p.ldG / 8)), 256, 0, st>>>(G, d_out, d_out, d_out, Gh, Gl, p.ldG,
                                                                               fbits);
    const int tiles_n = ceil_div(pad4(d_in), BN_Y), nchunks = p.ldT / HS_BK;
    kron_pack_image_kernel<<<dim3(hs_grid((long long)tiles_n * nchunks * BN_Y * 8), 1), 256, 0, st>>>(
        A, d_in, 1, 0, Aimg, 0, d_in, d_in, BN_Y, tiles_n, nchunks, fbits + 8);
    g_launches += 4;
    LAUNCH_CHECK();
  }
  CHECK_CUDA(cudaMemsetAsync(bits, 0, 64 * 4, st));
  // ---- X: per-column scale (one word for all columns: they come from one vector), images of X_z^T
  kron_absmax_kernel<<<grid1d((long long)d_out * d_in * K), 256, 0, st>>>(X, (long long)d_out * d_in * K, bits, 8);
  {
    const int tiles_n = ceil_div(p.ldT, BN_T), nchunks = p.ldG / HS_BK;
    // W_z[n = b][k = a'] = X[(a' * d_in + b) * K + z]:  sn = K, sk = d_in * K, slot stride 1
    kron_pack_image_kernel<<<dim3(hs_grid((long long)tiles_n * nchunks * BN_T * 8), K), 256, 0, st>>>(
        X, K, (long long)d_in * K, 1, (__half*)(base + p.off_X), p.x_img_halves, d_in, d_out, BN_T, tiles_n, nchunks,
        bits);
  }
  g_launches += 2;
  LAUNCH_CHECK();
  // ---- step 1: T_z[a][b] = sum_a' G[a][a'] X_z[a'][b]      (shared activation G, one weight image per column)
  float* T = (float*)(base + p.off_T);
  {
    HsGatherArgs h;
    memset(&h, 0, sizeof(h));
    Geom& g = h.g;
    g.B = d_out; g.Hs = g.Ws = g.Hd = g.Wd = 1; g.Cs = p.ldG; g.KH = g.KW = 1; g.sh = g.sw = 1; g.mode = 0;
    g.N = d_in; g.Nd = p.ldT; g.Kd = p.ldG; g.M = d_out;
    h.Ah = Gh; h.Al = Gl; h.A_slot = (long long)d_out * p.ldG; h.a_slot_base = 0; h.a_has_slots = 0; h.a_bits = fbits;
    h.W_img = (const __half*)(base + p.off_X);
    h.Wt_img = K > 1 ? (const __half*)(base + p.off_X) + p.x_img_halves : nullptr;
    h.Wt_img_slot = p.x_img_halves;
    h.w_bits = bits;
    h.out = T; h.out_slot = p.t_elems; h.slot0 = 0; h.accumulate = 0; h.planes = 2;
    ProfScope prof(0, 2.0 * d_out * (double)d_out * d_in * K, st);
    if (hs_launch_gather_gemm(h, K, st, false, 1)) return fail(CURV_ERR_CUDA, "Kronecker apply: step-1 launch failed");
    ++g_launches;
  }
  // ---- T planes (per-column scales)
  __half* Th = (__half*)(base + p.off_Tp);
  __half* Tl = Th + align_up(p.t_elems * K + 8192, 128);
  CHECK_CUDA(cudaMemsetAsync(Th + p.t_elems * K, 0, 8192 * 2, st));
  CHECK_CUDA(cudaMemsetAsync(Tl + p.t_elems * K, 0, 8192 * 2, st));
  if (hs_launch_absmax(T, p.t_elems, p.t_elems, bits + 16, K, st) ||
      hs_launch_split(T, p.t_elems, p.t_elems, Th, Tl, p.t_elems, bits + 16, K, st))
    return fail(CURV_ERR_CUDA, "Kronecker apply: split of the intermediate failed");
  g_launches += 2;
  // ---- step 2: Y_z[a][B] = sum_b T_z[a][b] A[B][b]
  float* Yz = (float*)(base + p.off_Y);
  {
    HsGatherArgs h;
    memset(&h, 0, sizeof(h));
    Geom& g = h.g;
    g.B = d_out; g.Hs = g.Ws = g.Hd = g.Wd = 1; g.Cs = p.ldT; g.KH = g.KW = 1; g.sh = g.sw = 1; g.mode = 0;
    g.N = d_in; g.Nd = p.NdY; g.Kd = p.ldT; g.M = d_out;
    h.Ah = Th; h.Al = Tl; h.A_slot = p.t_elems; h.a_slot_base = 0; h.a_has_slots = 1; h.a_bits = bits + 16;
    h.W_img = Aimg; h.Wt_img = nullptr; h.w_bits = fbits + 8;
    h.out = Yz; h.out_slot = p.y_elems; h.slot0 = 0; h.accumulate = 0; h.planes = 2;
    ProfScope prof(0, 2.0 * d_out * (double)d_in * d_in * K, st);
    if (hs_launch_gather_gemm(h, K, st, false, K)) return fail(CURV_ERR_CUDA, "Kronecker apply: step-2 launch failed");
    ++g_launches;
  }
  kron_interleave_kernel<<<grid1d((long long)d_out * d_in * K), 256, 0, st>>>(Yz, p.y_elems, p.NdY, d_out, d_in, K, Y);
  LAUNCH_CHECK();
  return CURV_OK;
}
