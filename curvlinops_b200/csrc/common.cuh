// Shared declarations of the curvb200 engine (host + device).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace curv {

// Geometry of one implicit-GEMM convolution (all tensors NHWC, channels padded to a multiple of 8:
// one 16-byte chunk of an fp16 plane = 8 channels of one pixel, see hs_gemm.cuh).
//   source tensor  : [B][Hs][Ws][Cs]            (what is gathered)
//   destination    : [B][Hd][Wd][Nd]            (one GEMM row per destination pixel)
//   mode 0 (forward / wgrad gather): source pixel = dest*stride - pad + tap
//   mode 1 (dgrad gather)          : source pixel = (dest + pad - tap)/stride, if divisible
struct Geom {
  int B, Hs, Ws, Cs;
  int Hd, Wd;
  int KH, KW, sh, sw, ph, pw;
  int mode;
  int N;   // valid rows of the weight matrix (real destination channels)
  int Nd;  // padded destination channels (row stride of the destination)
  int Kd;  // KH*KW*Cs, reduction length
  int M;   // B*Hd*Wd
};

// Slot algebra shared by all bilinear ops:  slot 0 = primal,  slot k>=1 = k-th tangent/cotangent.
//   out_0 = op(A_0, W)
//   out_k = op(A_k, W) [if A has slots]  +  op(A_0, Wt_k) [if Wt != null]
struct GatherGemmArgs {
  Geom g;
  const float* A;
  long long A_slot;
  int a_has_slots;
  const float* W;
  const float* Wt;
  long long Wt_slot;
  // tcgen05 path only: pre-split / pre-swizzled images of W and Wt (tc_gemm.cuh), null -> SIMT path
  const float* W_img;
  const float* Wt_img;
  long long Wt_img_slot;
  const float* bias;    // [Nd] added to slot 0 (may be null)
  const float* bias_t;  // [(k-1)*bias_slot + n] added to slot k (may be null)
  long long bias_slot;
  float* out;
  long long out_slot;
  int slot0;  // first slot computed; gridDim.y = number of slots
  int accumulate;
  int debug;  // perf experiments only: 1 = producers skip the global loads, 2 = skip split + smem stores
};

// wgrad:  D_k[n][tap][c] = sum_m G_k[m][n] * In_0[src(m,tap)][c]   (+ G_0 * In_k in R-op mode)
struct WgradArgs {
  Geom g;  // mode 0 geometry of the forward conv; (Hd,Wd) is the grid of G
  const float* G;
  long long G_slot;
  int Ng;  // padded channels of G
  const float* In;
  long long In_slot;
  int second_seg;  // 1: add (G_0, In_k) (Hessian R-op); needs In slots
  float* partial;  // [split][nslots][N][Kd]
  int nsplit, nslots, slot0;
  int m_per_split;
};

inline __host__ __device__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline __host__ __device__ int pad4(int c) { return (c + 3) & ~3; }
inline __host__ __device__ int pad8(int c) { return (c + 7) & ~7; }

// ---- half-split operand scales (hs_gemm.cuh): one power of two per (tensor, slot)
constexpr int HS_SH_TARGET = 14;  // scaled maximum in [2^14, 2^15)
constexpr int HS_SH_CLAMP = 60;   // |sh| <= 60: products of two inverse scales stay normal floats

// scale exponent sh (s = 2^sh) from the bit pattern of the slot's absolute maximum
__host__ __device__ inline int hs_shift_from_bits(uint32_t bits) {
  if (bits == 0u) return 0;  // all-zero slot
  const int e = (int)((bits >> 23) & 0xffu) - 127;
  int sh = HS_SH_TARGET - e;
  if (sh > HS_SH_CLAMP) sh = HS_SH_CLAMP;
  if (sh < -HS_SH_CLAMP) sh = -HS_SH_CLAMP;
  return sh;
}
__host__ __device__ inline float hs_pow2(int sh) {
#ifdef __CUDA_ARCH__
  return __uint_as_float((uint32_t)(sh + 127) << 23);
#else
  union { uint32_t u; float f; } c;
  c.u = (uint32_t)(sh + 127) << 23;
  return c.f;
#endif
}

}  // namespace curv
