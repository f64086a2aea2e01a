// "Half-split" tcgen05 path: fp32-grade contractions on the fp16 tensor pipe (kind::f16), sm_100a only.
//
// Every fp32 operand x of a contraction is stored as two fp16 planes of the power-of-two scaled value,
//     s*x = hi + lo,   hi = rn_fp16(s*x),  lo = rn_fp16(s*x - hi)          (22 significant bits)
// with one scale s = 2^sh per (tensor, slot), chosen from the slot's absolute maximum so that
// max|s*x| lies in [2^14, 2^15).  The product is evaluated as
//     a*b ~= (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo) / (s_a*s_b)               (dropped lo*lo ~ 2^-22)
// i.e. three fp16 MMAs accumulated in fp32 tensor memory - the same error budget as the 3xTF32 scheme of
// tc_gemm.cuh at twice the MMA rate and half the shared-memory / L2 operand bytes per reduction element.
// Elements smaller than 2^-17 of their slot's maximum lose relative (not absolute) accuracy: the absolute
// error of an operand is bounded by 2^-39 of the slot maximum.
//
// Because the operands are already in their final format in global memory, the producers are pure copies: one TMA
// im2col box per plane and stage for the gathered activations (cuTensorMapEncodeIm2col: 128 pixels x 64 channels of
// one filter tap, SWIZZLE_128B = the K-major UMMA tile; zero fill of padding / ragged edges by the hardware), one
// 3-d tiled TMA box per plane for the cotangent tiles of all slots in the wgrad, one cp.async.bulk for the
// pre-swizzled weight block; shapes TMA cannot express (C % 64 != 0) use 16-byte cp.async with zero fill,
// completion signalled on the stage's mbarrier (cp.async.mbarrier.arrive.noinc) - no register staging, no ALU work.
// The MMA warp runs its loop converged and one elect.sync lane issues the MMAs + commits of a stage.
//
//   hs_absmax_kernel / hs_split_kernel   fp32 [slot][n] -> scale bits + hi/lo planes (the BN kernels of
//                                        elementwise.cuh write planes directly where a bound gives the scale)
//   hs_pack_image_kernel                 packed fp32 weights [N][Kd] -> K-major SWIZZLE_128B image (hi, lo)
//   gather_gemm_hs<BN>                   forward conv + tangents, dgrad   (contract of gather_gemm_tc); main and
//                                        cross-term accumulators in separate TMEM buffers, parity-class
//                                        decomposition of strided dgrads
//   gather_gemm_hs_stack<BN>             shared-activation segment of a group of slots stacked along N (the stem)
//   wgrad_gemm_hs                        weight gradients of all K slots  (contract of wgrad_gemm_tc_ms)
// tools/hs_selftest.cu checks all of them against a CPU double reference; tests/host/hs_host_test.cu the host
// planning code.  The `debug` fields of the argument structs are timing experiments (results invalid).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "tc_gemm.cuh"

namespace curv {

#ifndef CURV_DISABLE_TC

constexpr int HS_BK = 64;        // fp16 reduction elements per stage = one 128-byte swizzle row
constexpr int HS_FLUSH = 6;      // K stages per main accumulation chunk of gather_gemm_hs (24 MMAs per chain)
// ---------------------------------------------------------------------------------------------------
// absmax + split (streaming, HBM-bound)
// ---------------------------------------------------------------------------------------------------
// bits[blockIdx.y] = max over slot blockIdx.y of |x| (as the uint bit pattern; atomicMax on non-negative
// floats is order independent, hence deterministic).  bits must be zero before the first use.
__global__ void __launch_bounds__(256) hs_absmax_kernel(const float* __restrict__ x, long long slot_stride,
                                                        long long n4, uint32_t* __restrict__ bits) {
  const int slot = blockIdx.y;
  const float4* p = reinterpret_cast<const float4*>(x + (long long)slot * slot_stride);
  float m = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(p + i);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float wm[8];
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = fmaxf(m, wm[w]);
    if (m > 0.f && __float_as_uint(m) > *reinterpret_cast<volatile uint32_t*>(bits + slot))
      atomicMax(bits + slot, __float_as_uint(m));
  }
}

__device__ __forceinline__ uint32_t hs_pack2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
// 8 consecutive floats -> one 16-byte chunk of the hi plane and one of the lo plane
__device__ __forceinline__ void hs_split8(const float4& v0, const float4& v1, float s, uint4& hi, uint4& lo) {
  const float x[8] = {v0.x * s, v0.y * s, v0.z * s, v0.w * s, v1.x * s, v1.y * s, v1.z * s, v1.w * s};
  __half h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    h[i] = __float2half_rn(x[i]);
    l[i] = __float2half_rn(x[i] - __half2float(h[i]));
  }
  hi = make_uint4(hs_pack2(h[0], h[1]), hs_pack2(h[2], h[3]), hs_pack2(h[4], h[5]), hs_pack2(h[6], h[7]));
  lo = make_uint4(hs_pack2(l[0], l[1]), hs_pack2(l[2], l[3]), hs_pack2(l[4], l[5]), hs_pack2(l[6], l[7]));
}

// bf16 mode (one plane, one MMA per product; BASELINE bf16 configs): 8 consecutive floats -> one 16-byte chunk of
// the single bf16 plane (stored through the __half* plane pointers: the planes are raw 16-bit words to the copy paths)
__device__ __forceinline__ uint4 hs_bf16x8(const float4& v0, const float4& v1, float s) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v0.x * s, v0.y * s), b = __floats2bfloat162_rn(v0.z * s, v0.w * s);
  const __nv_bfloat162 c = __floats2bfloat162_rn(v1.x * s, v1.y * s), d = __floats2bfloat162_rn(v1.z * s, v1.w * s);
  return make_uint4(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b),
                    *reinterpret_cast<const uint32_t*>(&c), *reinterpret_cast<const uint32_t*>(&d));
}

// x [slot][n] fp32 -> hi/lo [slot][n] fp16 (n = 8*n8), scale from bits[slot].  grid.y = slots.
// lo == nullptr: bf16 mode, only the (bf16) hi plane is written
__global__ void __launch_bounds__(256) hs_split_kernel(const float* __restrict__ x, long long slot_stride,
                                                       __half* __restrict__ hi, __half* __restrict__ lo,
                                                       long long out_slot_stride, long long n8,
                                                       const uint32_t* __restrict__ bits) {
  const int slot = blockIdx.y;
  const float s = hs_pow2(hs_shift_from_bits(bits[slot]));
  const float4* p = reinterpret_cast<const float4*>(x + (long long)slot * slot_stride);
  uint4* ph = reinterpret_cast<uint4*>(hi + (long long)blockIdx.y * out_slot_stride);
  uint4* pl = reinterpret_cast<uint4*>(lo + (long long)blockIdx.y * out_slot_stride);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 v0 = __ldg(p + 2 * i), v1 = __ldg(p + 2 * i + 1);
    if (lo == nullptr) { ph[i] = hs_bf16x8(v0, v1, s); continue; }
    uint4 h, l;
    hs_split8(v0, v1, s, h, l);
    ph[i] = h;
    pl[i] = l;
  }
}

// Weight image for gather_gemm_hs: src [N][Kd] fp32 row-major (slots along grid.y, scale bits[slot]).
//   block (tn, kc) = [hi plane: BN rows x 128 B (64 fp16), 128B-swizzled][lo plane], blocks ordered tn-major.
//   planes == 1 (bf16 mode): blocks hold the single bf16 plane only.
__global__ void hs_pack_image_kernel(const float* __restrict__ src, long long src_slot,
                                     __half* __restrict__ dst, long long dst_slot, int N, int Kd, int BN,
                                     int tiles_n, int nchunks, const uint32_t* __restrict__ bits, int planes) {
  src += blockIdx.y * src_slot;
  dst += blockIdx.y * dst_slot;
  const float s = hs_pow2(hs_shift_from_bits(bits[blockIdx.y]));
  const long long total = (long long)tiles_n * nchunks * BN * 8;  // 16-byte chunks per plane
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e & 7);
    long long t = e >> 3;
    const int r = (int)(t % BN); t /= BN;
    const int kc = (int)(t % nchunks);
    const int tn = (int)(t / nchunks);
    const int n = tn * BN + r, k = kc * HS_BK + c * 8;
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
    if (n < N && k < Kd) {
      const float4* q = reinterpret_cast<const float4*>(src + (long long)n * Kd + k);
      v0 = __ldg(q);
      if (k + 4 < Kd) v1 = __ldg(q + 1);
    }
    __half* blk = dst + ((long long)tn * nchunks + kc) * (planes * BN * HS_BK);
    const int o = ((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) >> 1;  // in halves
    if (planes == 1) { *reinterpret_cast<uint4*>(blk + o) = hs_bf16x8(v0, v1, s); continue; }
    uint4 h, l;
    hs_split8(v0, v1, s, h, l);
    *reinterpret_cast<uint4*>(blk + o) = h;
    *reinterpret_cast<uint4*>(blk + BN * HS_BK + o) = l;
  }
}

// Tangent-weight images straight from the columns of V (no packed fp32 intermediate): the rows of V that belong to a
// conv weight [N][C][taps] hold its K tangent directions side by side (K minor, row stride ldk).
//   hs_vcol_absmax_kernel      bits[k] = max_rows |V[row][k]|  (the per-(tensor, column) scale of the images)
//   hs_pack_image_cols_kernel  image k (dst + k*dst_slot) of the weight matrix [N][taps*Cp] of column k, same block /
//                              swizzle layout as hs_pack_image_kernel; the 8 elements of a 16-byte chunk are channels
//                              cp0..cp0+7 of one filter tap, i.e. 8 rows of V whose K columns are read together
__global__ void __launch_bounds__(256) hs_vcol_absmax_kernel(const float* __restrict__ V, long long rows, int ldk,
                                                            int K, uint32_t* __restrict__ bits) {
  float m[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) m[k] = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows;
       i += (long long)gridDim.x * blockDim.x) {
    const float* r = V + i * ldk;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < K) m[k] = fmaxf(m[k], fabsf(__ldg(r + k)));
  }
  __shared__ float wm[8][8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float v = m[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v = fmaxf(v, wm[w][threadIdx.x]);
    if (v > 0.f && __float_as_uint(v) > *reinterpret_cast<volatile uint32_t*>(bits + threadIdx.x))
      atomicMax(bits + threadIdx.x, __float_as_uint(v));
  }
}

__global__ void __launch_bounds__(256) hs_pack_image_cols_kernel(const float* __restrict__ V, int ldk, int K,
                                                                __half* __restrict__ dst, long long dst_slot, int N,
                                                                int C, int taps, int Cp, int BN, int tiles_n,
                                                                int nchunks, const uint32_t* __restrict__ bits,
                                                                int planes) {
  const int Kd = taps * Cp;
  const long long total = (long long)tiles_n * nchunks * BN * 8;  // 16-byte chunks per plane
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e & 7);
    long long t = e >> 3;
    const int r = (int)(t % BN); t /= BN;
    const int kc = (int)(t % nchunks);
    const int tn = (int)(t / nchunks);
    const int n = tn * BN + r, k = kc * HS_BK + c * 8;
    const int tap = k / Cp, cp0 = k - tap * Cp;
    const bool ok = n < N && k < Kd;
    const float* src = V + (((long long)n * C + cp0) * taps + tap) * ldk;  // row of element (n, cp0, tap)
    const long long rstep = (long long)taps * ldk;                           // next channel
    const long long blk_off = ((long long)tn * nchunks + kc) * (planes * BN * HS_BK);
    const int o = ((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) >> 1;  // in halves
    for (int slot = 0; slot < K; ++slot) {
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = (ok && cp0 + j < C) ? __ldg(src + j * rstep + slot) : 0.f;
      const float sc = hs_pow2(hs_shift_from_bits(bits[slot]));
      const float4 v0 = make_float4(x[0], x[1], x[2], x[3]), v1 = make_float4(x[4], x[5], x[6], x[7]);
      __half* blk = dst + (long long)slot * dst_slot + blk_off;
      if (planes == 1) { *reinterpret_cast<uint4*>(blk + o) = hs_bf16x8(v0, v1, sc); continue; }
      uint4 h, l;
      hs_split8(v0, v1, sc, h, l);
      *reinterpret_cast<uint4*>(blk + o) = h;
      *reinterpret_cast<uint4*>(blk + BN * HS_BK + o) = l;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------
// one lane of the (converged) warp, chosen by the hardware: tcgen05.mma / commit are issued under this predicate
// from warp-uniform control flow, which lets the compiler emit them without per-lane uniformisation loops
__device__ __forceinline__ bool hs_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// 16-byte async copy global -> shared, zero fill when !ok (the source address is then never dereferenced)
__device__ __forceinline__ void hs_cp16(uint32_t dst, const void* src, bool ok) {
  const uint32_t n = ok ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
// this thread's prior cp.async complete -> one (pre-counted) arrival on the mbarrier
__device__ __forceinline__ void hs_cp_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (fp16 inputs, fp32 accumulate)
__device__ __forceinline__ void hs_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// instruction descriptor: D = f32, A = B = f16, M = 128, N; majors: 0 = K-major, 1 = MN-major
// bf16 != 0: A = B = bf16 (format code 1 in bits 7-9 / 10-12)
__host__ __device__ constexpr uint32_t hs_idesc(int N, int a_mn_major, int b_mn_major, int bf16 = 0) {
  return (1u << 4) | ((uint32_t)(bf16 ? 1 : 0) << 7) | ((uint32_t)(bf16 ? 1 : 0) << 10) |
         ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}
// MN-major 16-bit operand, SWIZZLE_128B: atoms of 8 reduction rows x 128 bytes (64 MN-contiguous fp16),
// 16-byte chunk index XOR row-in-atom.  lbo = stride between 64-wide MN atoms, sbo = between 8-row K atoms.
__device__ __forceinline__ uint64_t hs_mnmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}

// Strided dgrad (mode 1, stride > 1): of the KH*KW filter taps only those with kh = (hd + ph) mod sh and
// kw = (wd + pw) mod sw reach a source pixel, so the destination pixels fall into sh*sw parity classes,
// each with its own (smaller) tap list.  In class coordinates (hd = sh*i + oh, wd = sw*j + ow) the gather is
// a stride-1 gather source(i + dh, j + dw): the kernel runs one dense GEMM per class instead of multiplying
// 1 - 1/(sh*sw) zeros.  Built on the host by hs_make_parity; needs Cs % 64 == 0 (one tap per K stage).
constexpr int HS_PAR_TAPS = 16;
struct HsParity {
  int nclass;  // 0: off
  int oh[4], ow[4], Hc[4], Wc[4];
  int tile0[5];  // first m-tile of each class (prefix sums), tile0[nclass] = total
  int ntap[4];
  signed char dh[4][HS_PAR_TAPS], dw[4][HS_PAR_TAPS];
  unsigned char tap[4][HS_PAR_TAPS];  // index of the tap in the weight image (kh*KW + kw)
};

struct HsGatherArgs {
  Geom g;
  const __half* Ah;       // hi plane of the gathered tensor, [slot][B*Hs*Ws*Cs]
  const __half* Al;       // lo plane
  long long A_slot;       // elements between slots
  int a_slot_base;        // absolute slot index of the first slot stored in the planes
  int a_has_slots;
  const uint32_t* a_bits;  // absmax bits of the gathered tensor, indexed by absolute slot
  const __half* W_img;     // image of W (hs_pack_image_kernel)
  const __half* Wt_img;    // images of the tangent weights, slot k at (k-1)*Wt_img_slot (may be null)
  long long Wt_img_slot;   // halves
  const uint32_t* w_bits;  // [0] = W, [k] = tangent weight k
  const float* bias;
  const float* bias_t;
  long long bias_slot;
  float* out;
  long long out_slot;
  int slot0;
  int accumulate;
  int debug;  // perf experiments only: 1 = producers skip the copies, 2 = the MMA lane skips the MMAs
  int flush;  // K stages per TMEM accumulation chunk (set by the launcher: g_hs_flush)
  unsigned int* out_bits;  // nullable: fused absmax of the written result, word [slot] (see elementwise.cuh)
  HsParity par;
  // TMA im2col path of the gathered operand (set by hs_launch_gather_gemm when Cs % 64 == 0): 4-d im2col maps
  // (C, W, H, slots*B) of the two planes, one pair per parity class (class 0 = the whole tensor otherwise):
  // box = 64 channels x 128 pixels, SWIZZLE_128B = exactly the K-major A tile of a stage.  The tile's first
  // pixel is (w, h) = (tma_w0 + x*tma_sw, tma_h0 + y*tma_sh) for destination / class coordinates (x, y);
  // the filter tap enters as the instruction's im2col offsets (tma_offw/h; per class in parity mode).
  int use_tma;
  int planes;  // 2 (0 = default): fp16 hi/lo planes; 1: one bf16 plane (Al unused), one MMA per product
  int tma_w0[4], tma_h0[4], tma_sw, tma_sh;
  unsigned char tma_offw[4][HS_PAR_TAPS], tma_offh[4][HS_PAR_TAPS];
  alignas(64) CUtensorMap tmA[4][2];  // [class][plane]
};

// ---------------------------------------------------------------------------------------------------
// gather GEMM:  out[m][n] = sum_seg (1 / (s_A s_W)) sum_r gather(A_seg)[m][r] * W_seg[n][r]
// Same CTA organisation as gather_gemm_tc (17 warps, persistent, smem ring + TMEM double buffer); a TMEM
// accumulation chunk (TC_FLUSH stages) never straddles two segments because they carry different scales.
// Requires Cs % 8 == 0 (a 16-byte chunk = 8 channels of one filter tap).
// ---------------------------------------------------------------------------------------------------
struct HsTile { int slot, m0, tn, cls, T; };  // T = K stages per segment

// Shared-memory map of gather_gemm_hs: the stage ring of TcCfg<BN> plus the barriers of the TMEM accumulators.
// All 512 tensor-memory columns are used, as 512 / BN buffers of BN columns:
//   buffers 0, 1        "cross" accumulators  D_x += a_lo b_hi + a_hi b_lo   (one per segment, double buffered)
//   buffers 2 .. NB-1   "main" accumulator ring  D_m += a_hi b_hi            (one per chunk of `flush` stages)
// Why two kinds: the tensor core truncates when it adds into the fp32 accumulator (~0.25 ulp(D) of bias per MMA,
// compounding over the ~60 layer applications of a product), so an accumulation CHAIN must stay short.  The
// cross terms are 2^-10 of the result: their chain may span a whole segment (drained once), and keeping them out
// of the main accumulator cuts the main chain to one MMA per K step - chunks can be 3x longer for the same bias,
// and draining a chunk (64 KB of TMEM reads that compete with the MMAs) happens 3x less often.
// PL = operand planes per tensor: 2 = fp16 hi/lo (half-split, three MMAs per product), 1 = one bf16 plane (bf16
// arithmetic: ONE MMA per product, no cross accumulators; the freed shared memory deepens the stage ring)
template <int BN, int PL>
struct HsCfg {
  static constexpr int A_BYTES = TC_BM * 128;  // per plane
  static constexpr int B_BYTES = BN * 128;     // per plane
  static constexpr int STAGE_BYTES = PL * (A_BYTES + B_BYTES);
  static constexpr int STAGES = PL == 2 ? (BN == 128 ? 3 : 4) : (BN == 128 ? 6 : 7);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 512 /*barriers*/;
};

template <int BN, int PL>
struct HsSmem {
  using Cfg = HsCfg<BN, PL>;
  static constexpr int NB = 512 / BN;
  static constexpr int NM = NB - 2;  // main ring
  uint32_t base, bar_base;
  __device__ explicit HsSmem(uint8_t* raw) {
    base = (smem_u32(raw) + 1023u) & ~1023u;
    bar_base = base + Cfg::STAGES * Cfg::STAGE_BYTES;
  }
  __device__ uint32_t full(int s) const { return bar_base + 8u * s; }
  __device__ uint32_t empty(int s) const { return bar_base + 8u * (Cfg::STAGES + s); }
  __device__ uint32_t tfull(int a) const { return bar_base + 8u * (2 * Cfg::STAGES + a); }
  __device__ uint32_t tempty(int a) const { return bar_base + 8u * (2 * Cfg::STAGES + NM + a); }
  __device__ uint32_t cfull(int a) const { return bar_base + 8u * (2 * Cfg::STAGES + 2 * NM + a); }
  __device__ uint32_t cempty(int a) const { return bar_base + 8u * (2 * Cfg::STAGES + 2 * NM + 2 + a); }
  __device__ uint32_t tmem_slot() const { return bar_base + 8u * (2 * Cfg::STAGES + 2 * NM + 4); }
  __device__ uint32_t stageA(int s) const { return base + s * Cfg::STAGE_BYTES; }
  __device__ uint32_t stageB(int s) const { return base + s * Cfg::STAGE_BYTES + PL * Cfg::A_BYTES; }
};
static_assert(8 * (2 * 7 + 2 * 6 + 4 + 1) <= 512, "barrier block of HsSmem must fit the 512 bytes reserved by HsCfg");

template <int BN, int PL>
__device__ __forceinline__ uint32_t hs_prologue(const HsSmem<BN, PL>& S, uint8_t* raw, int full_count) {
  using Cfg = HsCfg<BN, PL>;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(S.full(s), full_count); mbar_init(S.empty(s), 1); }
    for (int a = 0; a < HsSmem<BN, PL>::NM; ++a) { mbar_init(S.tfull(a), 1); mbar_init(S.tempty(a), 256); }
    for (int a = 0; a < 2; ++a) { mbar_init(S.cfull(a), 1); mbar_init(S.cempty(a), 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(S.tmem_slot()), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *reinterpret_cast<volatile uint32_t*>(raw + (S.tmem_slot() - smem_u32(raw)));
}
__device__ __forceinline__ void hs_teardown(uint32_t tmem_base) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// 4-d im2col TMA load global -> shared, completion on an mbarrier (complete_tx::bytes)
__device__ __forceinline__ void hs_tma_load_im2col(uint32_t dst, const CUtensorMap* map, int c, int w, int h, int n,
                                                   unsigned short offw, unsigned short offh, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4, %5}], [%6], {%7, %8};"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c), "r"(w), "r"(h), "r"(n), "r"(bar), "h"(offw), "h"(offh)
      : "memory");
}

template <int BN, int PL>
__global__ void __launch_bounds__(TC_THREADS, 1) gather_gemm_hs(const __grid_constant__ HsGatherArgs p, int nslots) {
  using Cfg = HsCfg<BN, PL>;  // stage: 128 rows x 128 B per plane of A, BN rows x 128 B per plane of W
  constexpr int STAGES = Cfg::STAGES;
  constexpr int NM = HsSmem<BN, PL>::NM;  // main accumulator ring (TMEM buffers 2..), buffers 0/1 = cross terms
  extern __shared__ uint8_t smem_raw[];
  const HsSmem<BN, PL> S(smem_raw);
  const Geom& g = p.g;
  const HsParity& par = p.par;
  const bool parity = par.nclass > 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = ceil_div(g.Nd, BN);
  const int tiles_m = parity ? par.tile0[par.nclass] : ceil_div(g.M, TC_BM);
  const int ntiles = tiles_m * tiles_n * nslots;
  const int nchunks = ceil_div(g.Kd, HS_BK);
  const int cpt = g.Cs / HS_BK;  // K stages per filter tap (parity mode)
  __shared__ unsigned int s_slotmax[40];  // per-CTA absmax of the written result per slot (out_bits)
  if (threadIdx.x < 40) s_slotmax[threadIdx.x] = 0u;
  // full barrier: TMA mode = the single expect_tx arrival of the issuing thread
  const uint32_t tmem_base = hs_prologue<BN, PL>(S, smem_raw, p.use_tma == 1 ? 1 : TC_PRODUCERS + 1);

  auto decode_tile = [&](int tile, HsTile& t) {
    const int si = tile % nslots;
    const int rest = tile / nslots;
    t.tn = rest % tiles_n;
    const int mt = rest / tiles_n;
    t.slot = p.slot0 + si;
    if (parity) {
      int cls = 0;
      while (cls + 1 < par.nclass && mt >= par.tile0[cls + 1]) ++cls;
      t.cls = cls; t.m0 = (mt - par.tile0[cls]) * TC_BM; t.T = par.ntap[cls] * cpt;
    } else {
      t.cls = 0; t.m0 = mt * TC_BM; t.T = nchunks;
    }
  };
  auto num_segments = [&](int slot) {
    return slot == 0 ? 1 : (p.a_has_slots ? 1 : 0) + (p.Wt_img != nullptr ? 1 : 0);
  };
  // segment s of a slot: which A slot is gathered, which weight (0 = W, k = tangent k)
  auto segment_ids = [&](int slot, int s, int& a_slot, int& w_id) {
    const bool first_is_act = (slot == 0) || p.a_has_slots;
    if (s == 0 && first_is_act) { a_slot = slot; w_id = 0; }
    else { a_slot = 0; w_id = slot; }
  };

  if (warp >= 5 && warp < 13 && p.use_tma == 1) {
    // ------------------------------------------------------------------ producer, TMA mode: ONE thread
    // per stage: two im2col boxes (hi / lo plane: 128 pixels x 64 channels of one filter tap) + the weight block
    if (warp == 5 && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t bytes = PL * (Cfg::A_BYTES + Cfg::B_BYTES);
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        HsTile t;
        decode_tile(tile, t);
        const int nseg = num_segments(t.slot);
        // first pixel of the tile in destination (or class) coordinates
        int x, y, bimg;
        if (parity) {
          const int hw = par.Hc[t.cls] * par.Wc[t.cls];
          bimg = t.m0 / hw;
          const int rem = t.m0 - bimg * hw;
          y = rem / par.Wc[t.cls]; x = rem - y * par.Wc[t.cls];
        } else {
          bimg = t.m0 / (g.Hd * g.Wd);
          const int rem = t.m0 - bimg * (g.Hd * g.Wd);
          y = rem / g.Wd; x = rem - y * g.Wd;
        }
        const int w0 = p.tma_w0[t.cls] + x * p.tma_sw, h0 = p.tma_h0[t.cls] + y * p.tma_sh;
        for (int seg = 0; seg < nseg; ++seg) {
          int a_slot, w_id;
          segment_ids(t.slot, seg, a_slot, w_id);
          const int n0 = (a_slot - p.a_slot_base) * g.B + bimg;
          const __half* Wimg = (w_id == 0) ? p.W_img : p.Wt_img + (long long)(w_id - 1) * p.Wt_img_slot;
          int ti = 0, cb = 0;  // tap position (in the class list / kh*KW + kw), channel chunk
          for (int kc = 0; kc < t.T; ++kc) {
            const int wblk = parity ? par.tap[t.cls][ti] * cpt + cb / HS_BK : kc;
            mbar_wait(S.empty(stage), phase ^ 1);
            if (p.debug & 1) {
              mbar_arrive(S.full(stage));
            } else {
              asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(S.full(stage)),
                           "r"(bytes)
                           : "memory");
              const uint32_t sA = S.stageA(stage);
              const unsigned short ow = p.tma_offw[t.cls][ti], oh = p.tma_offh[t.cls][ti];
              hs_tma_load_im2col(sA, &p.tmA[t.cls][0], cb, w0, h0, n0, ow, oh, S.full(stage));
              if (PL == 2)
                hs_tma_load_im2col(sA + Cfg::A_BYTES, &p.tmA[t.cls][1], cb, w0, h0, n0, ow, oh, S.full(stage));
              const __half* wsrc = Wimg + ((long long)t.tn * nchunks + wblk) * (PL * BN * HS_BK);
              asm volatile(
                  "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                      S.stageB(stage)),
                  "l"(wsrc), "r"((uint32_t)(PL * Cfg::B_BYTES)), "r"(S.full(stage))
                  : "memory");
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            cb += HS_BK;
            if (cb >= g.Cs) { cb = 0; ++ti; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 5 && warp < 13) {
    // ------------------------------------------------------------------ producers (pure cp.async)
    const int pt = threadIdx.x - 5 * 32;     // 0..255
    const int a_c = pt & 7, a_r0 = pt >> 3;  // chunk a_c (8 channels) of rows a_r0 + 32 i
    const uint32_t a_off = (uint32_t)((a_r0 >> 3) * 1024 + (a_r0 & 7) * 128 + ((a_c ^ (a_r0 & 7)) << 4));
    const bool fast = (g.Cs % HS_BK) == 0;  // a 128-byte K row never straddles two filter taps
    const bool linear = parity || g.mode == 0 || (g.sh == 1 && g.sw == 1);
    const int sgn = g.mode == 0 ? 1 : -1;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      HsTile t;
      decode_tile(tile, t);
      const int nseg = num_segments(t.slot);
      bool m_ok[4];
      int ah[4], aw[4], rowoff[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = t.m0 + a_r0 + 32 * i;
        if (parity) {
          const int hw = par.Hc[t.cls] * par.Wc[t.cls];
          m_ok[i] = m < g.B * hw;
          const int mm = m_ok[i] ? m : 0;
          const int bimg = mm / hw;
          const int rem = mm - bimg * hw;
          ah[i] = rem / par.Wc[t.cls]; aw[i] = rem - ah[i] * par.Wc[t.cls];
          rowoff[i] = ((bimg * g.Hs + ah[i]) * g.Ws + aw[i]) * g.Cs;
        } else {
          m_ok[i] = m < g.M;
          const int mm = m_ok[i] ? m : 0;
          const int bimg = mm / (g.Hd * g.Wd);
          const int rem = mm - bimg * (g.Hd * g.Wd);
          const int hd = rem / g.Wd, wd = rem - hd * g.Wd;
          ah[i] = g.mode == 0 ? hd * g.sh - g.ph : hd + g.ph;
          aw[i] = g.mode == 0 ? wd * g.sw - g.pw : wd + g.pw;
          // linear: element offset of (b, ah, aw, 0); generic: pixel offset of image b
          rowoff[i] = linear ? ((bimg * g.Hs + ah[i]) * g.Ws + aw[i]) * g.Cs : bimg * g.Hs * g.Ws;
        }
        if (!m_ok[i]) ah[i] = -(1 << 20);  // fails every bounds test
      }
      // hybrid mode (use_tma == 2): the hi plane of the stage comes by TMA im2col (thread 0), only the lo plane
      // by cp.async: the two paths have separate throughput limits (TMA row rate / LSU miss path)
      const bool hybrid = PL == 2 && p.use_tma == 2;
      int tw0 = 0, th0 = 0, tb = 0;
      if (hybrid && pt == 0) {
        int x, y;
        if (parity) {
          const int hw = par.Hc[t.cls] * par.Wc[t.cls];
          tb = t.m0 / hw;
          const int rem = t.m0 - tb * hw;
          y = rem / par.Wc[t.cls]; x = rem - y * par.Wc[t.cls];
        } else {
          tb = t.m0 / (g.Hd * g.Wd);
          const int rem = t.m0 - tb * (g.Hd * g.Wd);
          y = rem / g.Wd; x = rem - y * g.Wd;
        }
        tw0 = p.tma_w0[t.cls] + x * p.tma_sw; th0 = p.tma_h0[t.cls] + y * p.tma_sh;
      }
      for (int seg = 0; seg < nseg; ++seg) {
        int a_slot, w_id;
        segment_ids(t.slot, seg, a_slot, w_id);
        const __half* Ah = p.Ah + (long long)(a_slot - p.a_slot_base) * p.A_slot;
        const __half* Al = p.Al + (long long)(a_slot - p.a_slot_base) * p.A_slot;
        const __half* Wimg = (w_id == 0) ? p.W_img : p.Wt_img + (long long)(w_id - 1) * p.Wt_img_slot;
        int kh = 0, kw = 0, cb = 0, ti = 0;  // ti: position in the class tap list (parity mode)
        for (int kc = 0; kc < t.T; ++kc) {
          // (dy, dx): pixel displacement of this stage's tap, c: first channel of this thread's chunk,
          // wblk: K-block of the weight image
          int dy, dx, c = cb + a_c * 8, wblk = kc;
          bool rok = true;
          if (parity) {
            dy = par.dh[t.cls][ti]; dx = par.dw[t.cls][ti];
            wblk = par.tap[t.cls][ti] * cpt + cb / HS_BK;
          } else if (fast) {
            dy = sgn * kh; dx = sgn * kw;
          } else {  // the 16-byte chunk decides its own filter tap
            const int r = kc * HS_BK + a_c * 8;
            const int tap = r / g.Cs;
            c = r - tap * g.Cs;
            const int kh_ = tap / g.KW;
            dy = sgn * kh_; dx = sgn * (tap - kh_ * g.KW);
            rok = r < g.Kd;
          }
          mbar_wait(S.empty(stage), phase ^ 1);
          const uint32_t sA = S.stageA(stage);
          if (pt == 0) {
            const uint32_t bytes = PL * Cfg::B_BYTES;
            const __half* wsrc = Wimg + ((long long)t.tn * nchunks + wblk) * (PL * BN * HS_BK);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(S.full(stage)),
                         "r"(bytes + (hybrid ? (uint32_t)Cfg::A_BYTES : 0u))
                         : "memory");
            if (hybrid)
              hs_tma_load_im2col(sA, &p.tmA[t.cls][0], cb, tw0, th0, (a_slot - p.a_slot_base) * g.B + tb,
                                 p.tma_offw[t.cls][ti], p.tma_offh[t.cls][ti], S.full(stage));
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    S.stageB(stage)),
                "l"(wsrc), "r"(bytes), "r"(S.full(stage))
                : "memory");
          }
          if (p.debug & 1) {
            mbar_arrive(S.full(stage));
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            cb += HS_BK;
            if (cb >= g.Cs) { cb = 0; ++ti; if (++kw == g.KW) { kw = 0; ++kh; } }
            continue;
          }
          if (linear) {
            const int tapoff = (dy * g.Ws + dx) * g.Cs + c;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int hs = ah[i] + dy, ws = aw[i] + dx;
              const bool ok = rok && (unsigned)hs < (unsigned)g.Hs && (unsigned)ws < (unsigned)g.Ws;
              const long long eo = ok ? (long long)(rowoff[i] + tapoff) : 0;
              const uint32_t o = sA + a_off + (uint32_t)(i * 4096);
              if (!hybrid) hs_cp16(o, Ah + eo, ok);
              if (PL == 2) hs_cp16(o + Cfg::A_BYTES, Al + eo, ok);
            }
          } else {  // strided dgrad, generic: source pixel = (dest + pad - tap) / stride when divisible
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int th = ah[i] + dy, tw = aw[i] + dx;  // dy, dx = -kh, -kw
              bool ok = rok && m_ok[i] && th >= 0 && tw >= 0;
              const int hs = th / g.sh, ws = tw / g.sw;
              ok = ok && hs * g.sh == th && ws * g.sw == tw && hs < g.Hs && ws < g.Ws;
              const long long eo = ok ? ((long long)(rowoff[i] + hs * g.Ws + ws) * g.Cs + c) : 0;
              const uint32_t o = sA + a_off + (uint32_t)(i * 4096);
              hs_cp16(o, Ah + eo, ok);
              if (PL == 2) hs_cp16(o + Cfg::A_BYTES, Al + eo, ok);
            }
          }
          hs_cp_arrive(S.full(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          cb += HS_BK;
          if (cb >= g.Cs) { cb = 0; ++ti; if (++kw == g.KW) { kw = 0; ++kh; } }
        }
      }
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp runs the (uniform) loop and waits on the barriers; one elected lane issues the MMAs and
    // the commits of a stage in one predicated block.
    {
      constexpr uint32_t idesc = hs_idesc(BN, 0, 0, PL == 1);
      int stage = 0, acc = 0, xb = 0;
      uint32_t phase = 0, acc_phase = 0, xb_phase = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        HsTile t;
        decode_tile(tile, t);
        const int nseg = num_segments(t.slot);
        for (int seg = 0; seg < nseg && t.T > 0; ++seg) {
          if (PL == 2) mbar_wait(S.cempty(xb), xb_phase ^ 1);  // cross accumulator of this segment drained
          const uint32_t d_cross = tmem_base + (uint32_t)(xb * BN);
          for (int t0 = 0; t0 < t.T; t0 += p.flush) {  // one chunk of the main accumulation
            mbar_wait(S.tempty(acc), acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_main = tmem_base + (uint32_t)((2 + acc) * BN);
            const int T = min(p.flush, t.T - t0);
            for (int it = 0; it < T; ++it) {
              mbar_wait(S.full(stage), phase);
              tc_fence_after();
              const uint32_t sA = S.stageA(stage), sB = S.stageB(stage);
              const uint64_t dAh = make_kmajor_sw128_desc(sA), dAl = make_kmajor_sw128_desc(sA + Cfg::A_BYTES);
              const uint64_t dBh = make_kmajor_sw128_desc(sB), dBl = make_kmajor_sw128_desc(sB + Cfg::B_BYTES);
              if (hs_elect_one()) {
                if (p.use_tma != 1) fence_async_proxy();  // cp.async (generic proxy) writes -> tensor-core reads
                if (!(p.debug & 2)) {
#pragma unroll
                  for (int ks = 0; ks < HS_BK / 16; ++ks) {
                    const uint64_t adv = (uint64_t)((ks * 32) >> 4);  // +32 bytes inside the 128-byte row
                    hs_mma_f16(d_main, dAh + adv, dBh + adv, idesc, (it | ks) != 0 ? 1u : 0u);
                    if (PL == 2) {
                      hs_mma_f16(d_cross, dAl + adv, dBh + adv, idesc, (t0 | it | ks) != 0 ? 1u : 0u);
                      hs_mma_f16(d_cross, dAh + adv, dBl + adv, idesc, 1u);
                    }
                  }
                }
                tc_commit(S.empty(stage));                       // frees the smem stage when these MMAs retire
                if (it + 1 == T) {
                  tc_commit(S.tfull(acc));                       // main chunk complete -> epilogue
                  if (PL == 2 && t0 + T == t.T) tc_commit(S.cfull(xb));  // segment complete -> cross terms to the epilogue
                }
              }
              __syncwarp();
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == NM) { acc = 0; acc_phase ^= 1; }
          }
          if (++xb == 2) { xb = 0; xb_phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 0-3, 13-16)
    constexpr int HALF = BN / 2;
    const int egrp = warp >= 13 ? 1 : 0;
    const int quad = warp & 3;
    int acc = 0, xb = 0;
    uint32_t acc_phase = 0, xb_phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      HsTile t;
      decode_tile(tile, t);
      const int n0 = t.tn * BN + egrp * HALF;
      const int nseg = num_segments(t.slot);
      float accv[HALF];
#pragma unroll
      for (int j = 0; j < HALF; ++j) accv[j] = 0.f;
      // drain one TMEM buffer (this warp's 32 lanes x HALF columns) into accv, scaled by the segment's inverse scale
      auto drain = [&](int buf, float inv) {
#pragma unroll
        for (int c0 = 0; c0 < HALF; c0 += 16) {
          uint32_t r[16];
          tc_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BN + egrp * HALF + c0), r);
#pragma unroll
          for (int j = 0; j < 16; ++j) accv[c0 + j] = fmaf(__uint_as_float(r[j]), inv, accv[c0 + j]);
        }
        tc_fence_before();
      };
      for (int seg = 0; seg < nseg && t.T > 0; ++seg) {
        int a_slot, w_id;
        segment_ids(t.slot, seg, a_slot, w_id);
        const float inv = hs_pow2(-hs_shift_from_bits(__ldg(p.a_bits + a_slot)) -
                                  hs_shift_from_bits(__ldg(p.w_bits + w_id)));
        for (int t0 = 0; t0 < t.T; t0 += p.flush) {
          mbar_wait(S.tfull(acc), acc_phase);
          tc_fence_after();
          drain(2 + acc, inv);
          mbar_arrive(S.tempty(acc));
          if (++acc == NM) { acc = 0; acc_phase ^= 1; }
        }
        if (PL == 2) {
          mbar_wait(S.cfull(xb), xb_phase);
          tc_fence_after();
          drain(xb, inv);
          mbar_arrive(S.cempty(xb));
        }
        if (++xb == 2) { xb = 0; xb_phase ^= 1; }
      }
      const float* bias = (t.slot == 0) ? p.bias
                                        : (p.bias_t ? p.bias_t + (long long)(t.slot - 1) * p.bias_slot : nullptr);
      float* outp = p.out + (long long)t.slot * p.out_slot;
      int m = t.m0 + quad * 32 + lane;  // GEMM row -> destination pixel
      bool m_ok = m < g.M;
      float vmax = 0.f;
      if (parity) {
        const int hw = par.Hc[t.cls] * par.Wc[t.cls];
        m_ok = m < g.B * hw;
        const int mm = m_ok ? m : 0;
        const int bimg = mm / hw;
        const int rem = mm - bimg * hw;
        const int i = rem / par.Wc[t.cls], j = rem - i * par.Wc[t.cls];
        m = (bimg * g.Hd + g.sh * i + par.oh[t.cls]) * g.Wd + g.sw * j + par.ow[t.cls];
      }
      if (m_ok) {
#pragma unroll
        for (int j = 0; j < HALF / 4; ++j) {
          const int n = n0 + j * 4;
          if (n >= g.Nd) continue;
          float4 v = make_float4(accv[j * 4 + 0], accv[j * 4 + 1], accv[j * 4 + 2], accv[j * 4 + 3]);
          if (bias) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + n));
            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
          }
          float4* dst = reinterpret_cast<float4*>(outp + (long long)m * g.Nd + n);
          if (p.accumulate) {
            const float4 o = *dst;
            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
          }
          *dst = v;
          vmax = fmaxf(vmax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
      }
      if (p.out_bits) {  // one shared-memory atomic per warp and tile, one global atomic per CTA and slot at the end
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        if (lane == 0 && vmax > 0.f) atomicMax(&s_slotmax[t.slot], __float_as_uint(vmax));
      }
    }
  }
  hs_teardown(tmem_base);
  if (p.out_bits && threadIdx.x < 40 && s_slotmax[threadIdx.x] != 0u)
    atomicMax(p.out_bits + threadIdx.x, s_slotmax[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------------
// N-stacked gather GEMM for the SHARED-activation segment:  out_k[m][n] (+)= sum_r gather(A_0)[m][r] * W_k[n][r]
// for a GROUP of slots k at once (the tangent term conv(a, Wdot_k) of every layer, and all slots of a layer whose
// input carries no tangent - the stem).  The gathered tile of the one shared operand is staged ONCE per group and
// multiplied against the weight blocks of the group's slots stacked along N (N = slots * BN <= 256): the
// activation gather - the measured limit of the C_out = 64 layers - is paid once per 256 / BN slots, and one
// N = 256 MMA reads the A tile from shared memory once for all of them.
//   tile    = 128 rows x (group of G = 256 / BN slots) x one n-tile of BN channels; K stage = 64
//   smem    = 2 stages x (A hi/lo 32 KB + B hi/lo 64 KB)
//   TMEM    = main accumulator (hi*hi, columns 0..255) + cross accumulator (lo*hi + hi*lo, columns 256..511):
//             main is drained every `flush` stages (short truncation chains, see HsSmem), cross once per tile
//   warps   = 0-15 epilogue (TMEM lane quadrant w & 3, column group w >> 2 of 64 columns), 16 MMA,
//             17-19 producers (16-byte cp.async with zero fill; weights by cp.async.bulk, 2 per slot)
// Forward geometry (mode 0) only.  Requires Cs % 8 == 0.
// ---------------------------------------------------------------------------------------------------
struct HsStackArgs {
  Geom g;
  const __half* Ah;        // planes of the shared operand (ONE slot)
  const __half* Al;
  const uint32_t* a_bits;  // [0]
  const __half* W_img;     // image of W (slot 0)
  const __half* Wt_img;    // images of the tangent weights, slot k at (k-1)*Wt_img_slot
  long long Wt_img_slot;
  const uint32_t* w_bits;  // [0] = W, [k] = tangent weight k
  const float* bias;
  const float* bias_t;
  long long bias_slot;
  float* out;
  long long out_slot;
  int slot_lo, nslots;     // slots slot_lo .. slot_lo + nslots - 1
  int accumulate;
  int flush;
  int planes;              // 2 (0 = default) or 1 (bf16), see HsGatherArgs
};

constexpr int HSN_THREADS = 20 * 32;  // 20 warps: 65536 / 640 leaves 96 registers per thread for the epilogue
constexpr int HSN_PRODUCERS = 96;
constexpr int HSN_A_BYTES = TC_BM * 128;            // per plane
constexpr int HSN_B_BYTES = 256 * 128;              // per plane (up to 256 stacked rows)
template <int PL>
struct HsnCfg {  // PL = 2: 2 stages of 96 KB; PL = 1 (bf16): 4 stages of 48 KB
  static constexpr int STAGES = PL == 2 ? 2 : 4;
  static constexpr int STAGE_BYTES = PL * (HSN_A_BYTES + HSN_B_BYTES);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int BN, int PL>
__global__ void __launch_bounds__(HSN_THREADS, 1) gather_gemm_hs_stack(const HsStackArgs p) {
  constexpr int HSN_STAGES = HsnCfg<PL>::STAGES;
  constexpr int HSN_STAGE_BYTES = HsnCfg<PL>::STAGE_BYTES;
  constexpr int G = 256 / BN;  // slots per group
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = sbase + HSN_STAGES * HSN_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (HSN_STAGES + s); };
  const uint32_t tfull = bar_base + 8u * (2 * HSN_STAGES), tempty = tfull + 8u, cfull = tfull + 16u,
                 cempty = tfull + 24u, tmem_slot = tfull + 32u;
  const Geom& g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = ceil_div(g.Nd, BN);
  const int tiles_m = ceil_div(g.M, TC_BM);
  const int ngroups = ceil_div(p.nslots, G);
  const int ntiles = tiles_m * tiles_n * ngroups;
  const int nchunks = ceil_div(g.Kd, HS_BK);

  if (threadIdx.x == 0) {
    for (int s = 0; s < HSN_STAGES; ++s) { mbar_init(full_bar(s), HSN_PRODUCERS + 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull, 1); mbar_init(tempty, 512); mbar_init(cfull, 1); mbar_init(cempty, 512);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  // tile -> (group, m0, tn); the groups of one (m-tile, n-tile) are adjacent: they gather the same rows
  auto decode_tile = [&](int tile, int& grp, int& m0, int& tn) {
    grp = tile % ngroups;
    const int rest = tile / ngroups;
    tn = rest % tiles_n;
    m0 = (rest / tiles_n) * TC_BM;
  };

  if (warp >= 17) {
    // ------------------------------------------------------------------ producers
    const int pt = threadIdx.x - 17 * 32;     // 0..95
    const int a_c = pt & 7, a_r0 = pt >> 3;   // chunk a_c (8 channels) of rows a_r0 + 12 i < 128, i < 11
    constexpr int NR = 11;
    const bool fast = (g.Cs % HS_BK) == 0;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int grp, m0, tn;
      decode_tile(tile, grp, m0, tn);
      const int s_first = p.slot_lo + grp * G;
      const int cnt = min(G, p.slot_lo + p.nslots - s_first);
      int ah[NR], aw[NR], rowoff[NR];
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const int m = m0 + a_r0 + 12 * i;
        const bool ok = m < g.M && a_r0 + 12 * i < TC_BM;
        const int mm = ok ? m : 0;
        const int bimg = mm / (g.Hd * g.Wd);
        const int rem = mm - bimg * (g.Hd * g.Wd);
        const int hd = rem / g.Wd, wd = rem - hd * g.Wd;
        ah[i] = ok ? hd * g.sh - g.ph : -(1 << 20);  // a row beyond M fails every bounds test
        aw[i] = wd * g.sw - g.pw;
        rowoff[i] = ((bimg * g.Hs + hd * g.sh - g.ph) * g.Ws + aw[i]) * g.Cs;
      }
      int kh = 0, kw = 0, cb = 0;
      for (int kc = 0; kc < nchunks; ++kc) {
        int dy = kh, dx = kw, c = cb + a_c * 8;
        bool rok = true;
        if (!fast) {  // the 16-byte chunk decides its own filter tap
          const int r = kc * HS_BK + a_c * 8;
          const int tap = r / g.Cs;
          c = r - tap * g.Cs;
          dy = tap / g.KW; dx = tap - dy * g.KW;
          rok = r < g.Kd;
        }
        mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t sA = sbase + stage * HSN_STAGE_BYTES;
        const uint32_t sB = sA + PL * HSN_A_BYTES;
        if (pt == 0) {  // weight blocks of the group's slots, stacked along N: hi planes, then lo planes
          const uint32_t blk = (uint32_t)BN * 128u;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_bar(stage)),
                       "r"((uint32_t)PL * blk * (uint32_t)cnt)
                       : "memory");
          for (int j = 0; j < cnt; ++j) {
            const int s = s_first + j;
            const __half* Wimg = (s == 0) ? p.W_img : p.Wt_img + (long long)(s - 1) * p.Wt_img_slot;
            const __half* wsrc = Wimg + ((long long)tn * nchunks + kc) * (PL * BN * HS_BK);
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    sB + (uint32_t)j * blk),
                "l"(wsrc), "r"(blk), "r"(full_bar(stage))
                : "memory");
            if (PL == 2)
              asm volatile(
                  "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                      sB + (uint32_t)HSN_B_BYTES + (uint32_t)j * blk),
                  "l"(wsrc + BN * HS_BK), "r"(blk), "r"(full_bar(stage))
                  : "memory");
          }
        }
        const int tapoff = (dy * g.Ws + dx) * g.Cs + c;
#pragma unroll
        for (int i = 0; i < NR; ++i) {
          const int r = a_r0 + 12 * i;
          if (r >= TC_BM) continue;
          const int hs = ah[i] + dy, ws = aw[i] + dx;
          const bool ok = rok && (unsigned)hs < (unsigned)g.Hs && (unsigned)ws < (unsigned)g.Ws;
          const long long eo = ok ? (long long)(rowoff[i] + tapoff) : 0;
          const uint32_t o = sA + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((a_c ^ (r & 7)) << 4));
          hs_cp16(o, p.Ah + eo, ok);
          if (PL == 2) hs_cp16(o + HSN_A_BYTES, p.Al + eo, ok);
        }
        hs_cp_arrive(full_bar(stage));
        if (++stage == HSN_STAGES) { stage = 0; phase ^= 1; }
        cb += HS_BK;
        if (cb >= g.Cs) { cb = 0; if (++kw == g.KW) { kw = 0; ++kh; } }
      }
    }
  } else if (warp == 16) {
    // ------------------------------------------------------------------ MMA issuer (uniform loop, elected lane)
    int stage = 0;
    uint32_t phase = 0, tphase = 0, cphase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int grp, m0, tn;
      decode_tile(tile, grp, m0, tn);
      const int cnt = min(G, p.nslots - grp * G);
      const uint32_t idesc = hs_idesc(cnt * BN, 0, 0, PL == 1);
      if (PL == 2) mbar_wait(cempty, cphase ^ 1);  // cross accumulator of the previous tile drained
      for (int t0 = 0; t0 < nchunks; t0 += p.flush) {
        mbar_wait(tempty, tphase ^ 1);  // main accumulator of the previous chunk drained
        tc_fence_after();
        const int T = min(p.flush, nchunks - t0);
        for (int it = 0; it < T; ++it) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sA = sbase + stage * HSN_STAGE_BYTES, sB = sA + PL * HSN_A_BYTES;
          const uint64_t dAh = make_kmajor_sw128_desc(sA), dAl = make_kmajor_sw128_desc(sA + HSN_A_BYTES);
          const uint64_t dBh = make_kmajor_sw128_desc(sB), dBl = make_kmajor_sw128_desc(sB + HSN_B_BYTES);
          if (hs_elect_one()) {
            fence_async_proxy();  // cp.async (generic proxy) writes -> tensor-core reads
#pragma unroll
            for (int ks = 0; ks < HS_BK / 16; ++ks) {
              const uint64_t adv = (uint64_t)((ks * 32) >> 4);
              hs_mma_f16(tmem_base, dAh + adv, dBh + adv, idesc, (it | ks) != 0 ? 1u : 0u);
              if (PL == 2) {
                hs_mma_f16(tmem_base + 256u, dAl + adv, dBh + adv, idesc, (t0 | it | ks) != 0 ? 1u : 0u);
                hs_mma_f16(tmem_base + 256u, dAh + adv, dBl + adv, idesc, 1u);
              }
            }
            tc_commit(empty_bar(stage));
            if (it + 1 == T) {
              tc_commit(tfull);
              if (PL == 2 && t0 + T == nchunks) tc_commit(cfull);
            }
          }
          __syncwarp();
          if (++stage == HSN_STAGES) { stage = 0; phase ^= 1; }
        }
        tphase ^= 1;
      }
      cphase ^= 1;
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 0-15)
    const int quad = warp & 3, cg = warp >> 2;  // TMEM lanes 32*quad.., accumulator columns 64*cg .. +63
    const int j = (cg * 64) / BN;               // slot of the group these columns belong to
    const int ncol = (cg * 64) % BN;            // first channel inside the slot's n-tile
    const int sh_a = hs_shift_from_bits(__ldg(p.a_bits));
    uint32_t tphase = 0, cphase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int grp, m0, tn;
      decode_tile(tile, grp, m0, tn);
      const int s_first = p.slot_lo + grp * G;
      const int cnt = min(G, p.slot_lo + p.nslots - s_first);
      const bool active = j < cnt;
      const int slot = s_first + (active ? j : 0);
      const float inv = hs_pow2(-sh_a - hs_shift_from_bits(__ldg(p.w_bits + slot)));
      float accv[64];
#pragma unroll
      for (int q = 0; q < 64; ++q) accv[q] = 0.f;
      auto drain = [&](uint32_t col0) {
        if (active) {
#pragma unroll
          for (int c0 = 0; c0 < 64; c0 += 16) {
            uint32_t r[16];
            tc_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + col0 + (uint32_t)(cg * 64 + c0), r);
#pragma unroll
            for (int q = 0; q < 16; ++q) accv[c0 + q] = fmaf(__uint_as_float(r[q]), inv, accv[c0 + q]);
          }
        }
        tc_fence_before();
      };
      for (int t0 = 0; t0 < nchunks; t0 += p.flush) {
        mbar_wait(tfull, tphase);
        tc_fence_after();
        drain(0u);
        mbar_arrive(tempty);
        tphase ^= 1;
      }
      if (PL == 2) {
        mbar_wait(cfull, cphase);
        tc_fence_after();
        drain(256u);
        mbar_arrive(cempty);
      }
      cphase ^= 1;
      if (!active) continue;
      const float* bias = (slot == 0) ? p.bias : (p.bias_t ? p.bias_t + (long long)(slot - 1) * p.bias_slot : nullptr);
      float* outp = p.out + (long long)slot * p.out_slot;
      const int m = m0 + quad * 32 + lane;
      if (m < g.M) {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int n = tn * BN + ncol + q * 4;
          if (n >= g.Nd) continue;
          float4 v = make_float4(accv[q * 4 + 0], accv[q * 4 + 1], accv[q * 4 + 2], accv[q * 4 + 3]);
          if (bias) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + n));
            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
          }
          float4* dst = reinterpret_cast<float4*>(outp + (long long)m * g.Nd + n);
          if (p.accumulate) {
            const float4 o = *dst;
            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
          }
          *dst = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// host: parity classes of a strided dgrad geometry; returns false (par.nclass = 0) when not applicable
static inline bool hs_make_parity(const Geom& g, HsParity& par) {
  memset(&par, 0, sizeof(par));
  if (g.mode != 1 || (g.sh == 1 && g.sw == 1) || g.sh * g.sw > 4 || g.Cs % HS_BK != 0) return false;
  struct Cls { int oh, ow, Hc, Wc, ntap; signed char dh[HS_PAR_TAPS], dw[HS_PAR_TAPS]; unsigned char tap[HS_PAR_TAPS]; };
  Cls cls[4];
  int nc = 0;
  for (int ch = 0; ch < g.sh; ++ch)
    for (int cw = 0; cw < g.sw; ++cw) {
      Cls c;
      memset(&c, 0, sizeof(c));
      c.oh = ((ch - g.ph) % g.sh + g.sh) % g.sh;
      c.ow = ((cw - g.pw) % g.sw + g.sw) % g.sw;
      c.Hc = c.oh < g.Hd ? (g.Hd - c.oh + g.sh - 1) / g.sh : 0;
      c.Wc = c.ow < g.Wd ? (g.Wd - c.ow + g.sw - 1) / g.sw : 0;
      if (c.Hc == 0 || c.Wc == 0) continue;
      for (int kh = ch; kh < g.KH; kh += g.sh)
        for (int kw = cw; kw < g.KW; kw += g.sw) {
          if (c.ntap == HS_PAR_TAPS) return false;
          const int dh = (c.oh + g.ph - kh) / g.sh, dw = (c.ow + g.pw - kw) / g.sw;  // exact divisions
          if (dh < -127 || dh > 127 || dw < -127 || dw > 127) return false;
          c.dh[c.ntap] = (signed char)dh; c.dw[c.ntap] = (signed char)dw;
          c.tap[c.ntap] = (unsigned char)(kh * g.KW + kw);
          ++c.ntap;
        }
      cls[nc++] = c;
    }
  if (nc == 0) return false;
  for (int a = 0; a < nc; ++a)  // heaviest classes first (static round-robin tile schedule)
    for (int b = a + 1; b < nc; ++b)
      if (cls[b].ntap > cls[a].ntap) { Cls tmp = cls[a]; cls[a] = cls[b]; cls[b] = tmp; }
  par.nclass = nc;
  int tiles = 0;
  for (int a = 0; a < nc; ++a) {
    par.oh[a] = cls[a].oh; par.ow[a] = cls[a].ow; par.Hc[a] = cls[a].Hc; par.Wc[a] = cls[a].Wc;
    par.ntap[a] = cls[a].ntap;
    memcpy(par.dh[a], cls[a].dh, HS_PAR_TAPS); memcpy(par.dw[a], cls[a].dw, HS_PAR_TAPS);
    memcpy(par.tap[a], cls[a].tap, HS_PAR_TAPS);
    par.tile0[a] = tiles;
    tiles += ceil_div(g.B * cls[a].Hc * cls[a].Wc, TC_BM);
  }
  par.tile0[nc] = tiles;
  return true;
}

// ---------------------------------------------------------------------------------------------------
// Multi-slot wgrad:  D_k[i][n] = (1 / (s_In s_Gk)) sum_m In[m][i] * G_k[m][n]   for ALL slots k of a tile.
//   tile = 128 columns i (tap*Cs + c) x 64 channels n x one split of the pixel range
//   A operand (M = 128): gathered primal input, MN-major;  B operand: the 64-channel G tiles of up to 4 slots
//   side by side (N = 64 * slots-in-group <= 256), MN-major, so ONE MMA feeds 4 accumulators and the A tile
//   is read from shared memory once per 4 slots.  The 8 accumulators [128 x 64] fill the 512 TMEM columns.
//   stage = 16 pixels (one K = 16 MMA step), 40 KB:
//     [In hi 4K][In lo 4K][G hi: 8 slots x 2K][G lo: 8 slots x 2K]      (slot group g = slots 4g .. 4g+3)
//   MN-major SWIZZLE_128B: chunk c (16 B = 8 elements) of pixel row r of MN atom a (64 elements) at
//     a*2048 + (r>>3)*1024 + (r&7)*128 + ((c ^ (r&7)) << 4)          (LBO = 2048, SBO = 1024)
// TMEM is single-buffered; every HSW_FLUSH stages (4096 pixels) the epilogue adds the chunk into the split's
// partial in global memory (same thread, same address, fixed order) - bounds the truncation-bias chain.
// Requires Cs % 8 == 0 and Ng % 8 == 0.
// ---------------------------------------------------------------------------------------------------
struct HsWgradArgs {
  Geom g;                  // mode 0 geometry of the forward conv; (Hd, Wd) is the grid of G
  const __half* Gh;        // cotangent planes [slot - slot0][M*Ng]  (slot0 first)
  const __half* Gl;
  long long G_slot;        // elements
  long long G_ld;          // row stride of G in elements (0: Ng).  With G_ld > Ng and G_slot = Ng the "slots" are
                           // adjacent column blocks of ONE matrix: the Gram matrices of kfac.cuh
  int Ng;
  const uint32_t* g_bits;  // indexed by absolute slot
  const __half* Ih;        // primal input planes [B*Hs*Ws*Cs]
  const __half* Il;
  const uint32_t* i_bits;  // [0]
  float* partial;          // [split][nslots][N][Kd]
  int nsplit, nslots, slot0, m_per_split;
  int debug;               // perf experiments only: 1 = producers skip the copies, 2 = no MMAs
  // TMA path of the G operand (set by hs_launch_wgrad): 3-d tiled maps (Ng, M, slots) of the two planes, box
  // (64 channels, 16 pixels, all slots) with SWIZZLE_128B = per slot one MN atom x two K atoms of the stage.
  int use_tma;
  int planes;              // 2 (0 = default) or 1 (bf16, Gl / Il unused), see HsGatherArgs
  alignas(64) CUtensorMap tmGh;
  alignas(64) CUtensorMap tmGl;
};

constexpr int HSW_ROWS = 16;
constexpr int HSW_A_BYTES = 128 * HSW_ROWS * 2;   // 4 KB per plane
constexpr int HSW_B_BYTES = 256 * HSW_ROWS * 2;   // 8 KB per plane and slot group
template <int PL>
struct HswCfg {  // PL = 2: 5 stages of 40 KB; PL = 1 (bf16): 7 stages of 20 KB (one gather warp per ring slot: warps 6-12)
  static constexpr int STAGE_BYTES = PL * (HSW_A_BYTES + 2 * HSW_B_BYTES);
  static constexpr int STAGES = PL == 2 ? 5 : 7;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};
constexpr int HSW_FLUSH = 256;  // stages (of 16 pixels) per TMEM accumulation chunk

// 3-d tiled TMA load global -> shared, completion on an mbarrier (complete_tx::bytes)
__device__ __forceinline__ void hs_tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                               uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}

template <int PL>
__global__ void __launch_bounds__(TC_THREADS, 1) wgrad_gemm_hs(const __grid_constant__ HsWgradArgs p) {
  constexpr int HSW_STAGES = HswCfg<PL>::STAGES;
  constexpr int HSW_STAGE_BYTES = HswCfg<PL>::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = sbase + HSW_STAGES * HSW_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (HSW_STAGES + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * HSW_STAGES);
  const uint32_t tempty_bar = bar_base + 8u * (2 * HSW_STAGES + 1);
  const uint32_t tmem_slot = bar_base + 8u * (2 * HSW_STAGES + 2);

  const Geom& g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NS = p.nslots;  // 1..8
  const int tiles_i = ceil_div(g.Kd, TC_BM), tiles_j = ceil_div(p.Ng, 64);
  const int ntiles = tiles_i * tiles_j * p.nsplit;

  if (threadIdx.x == 0) {
    for (int s = 0; s < HSW_STAGES; ++s) {
      // TMA mode: 32 lanes of the stage's gather warp + the expect_tx arrival of the TMA issuer
      mbar_init(full_bar(s), p.use_tma ? 33 : TC_PRODUCERS);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  auto decode_tile = [&](int tile, int& split, int& i0, int& j0) {
    const int tj = tile % tiles_j;
    const int rest = tile / tiles_j;
    const int ti = rest % tiles_i;
    split = rest / tiles_i;
    i0 = ti * TC_BM; j0 = tj * 64;
  };
  auto stages_of = [&](int split) {
    const int mb = split * p.m_per_split;
    const int me = min(g.M, mb + p.m_per_split);
    return ceil_div(max(0, me - mb), HSW_ROWS);
  };

  if (warp >= 5 && warp < 13 && p.use_tma) {
    // ------------------------------------------------------------------ producers, TMA mode
    // warp 5, lane 0: the G tiles of every stage by TMA (one box per plane: 64 channels x 16 pixels x NS slots).
    // warps 6-10    : the gathered-input tile, ONE WARP PER RING SLOT (stage q of this CTA belongs to warp
    //                 q % HSW_STAGES, i.e. a warp always refills the same smem slot: it sees every phase of that
    //                 slot's barriers in order, which parity waits require), so the address arithmetic of a
    //                 stage (pixel decode, bounds) is paid once per 16 copies and a warp has HSW_STAGES stage
    //                 periods to hide its latency.  Warps 11-12 idle.
    //                 lane -> pixel row lane >> 1, column half lane & 1 (64 columns = 8 chunks, both planes).
    if (warp == 5) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t bytes = (uint32_t)NS * (uint32_t)PL * 2048u;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
          int split, i0, j0;
          decode_tile(tile, split, i0, j0);
          const int mb = split * p.m_per_split;
          const int nst = stages_of(split);
          for (int st = 0; st < nst; ++st) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            const uint32_t sB = sbase + stage * HSW_STAGE_BYTES + PL * HSW_A_BYTES;
            if (p.debug & 1) {
              mbar_arrive(full_bar(stage));
            } else {
              asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_bar(stage)),
                           "r"(bytes)
                           : "memory");
              hs_tma_load_3d(sB, &p.tmGh, j0, mb + st * HSW_ROWS, 0, full_bar(stage));
              if (PL == 2) hs_tma_load_3d(sB + 2 * HSW_B_BYTES, &p.tmGl, j0, mb + st * HSW_ROWS, 0, full_bar(stage));
            }
            if (++stage == HSW_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
      __syncwarp();
    } else if (warp < 6 + HSW_STAGES) {
      const int gw = warp - 6;               // ring slot owned by this warp
      const int ir = lane >> 1, half = lane & 1;
      const bool uniform = (g.Cs % 64) == 0 || g.KH * g.KW == 1;  // the 64 columns of a half lie inside one filter tap
      const uint32_t offRow = (uint32_t)(half * 2048 + (ir >> 3) * 1024 + (ir & 7) * 128);
      int q0 = 0;  // stages of this CTA before the current tile
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int split, i0, j0;
        decode_tile(tile, split, i0, j0);
        const int mb = split * p.m_per_split;
        const int me = min(g.M, mb + p.m_per_split);
        const int nst = stages_of(split);
        const int col0 = i0 + half * 64;
        int tap0 = 0, ic0 = 0, kh0 = 0, kw0 = 0;
        if (uniform && col0 < g.Kd) { tap0 = col0 / g.Cs; ic0 = col0 - tap0 * g.Cs; kh0 = tap0 / g.KW; kw0 = tap0 - kh0 * g.KW; }
        int st = (gw - q0 % HSW_STAGES + HSW_STAGES) % HSW_STAGES;  // first stage of this tile owned by this warp
        for (; st < nst; st += HSW_STAGES) {
          const int q = q0 + st;
          const int stage = gw;
          const uint32_t phase = (uint32_t)((q / HSW_STAGES) & 1);
          const int mi = mb + st * HSW_ROWS + ir;
          const bool mok = mi < me;
          const int mm = mok ? mi : 0;
          const int bimg = mm / (g.Hd * g.Wd);
          const int rem = mm - bimg * (g.Hd * g.Wd);
          const int hd = rem / g.Wd, wd = rem - hd * g.Wd;
          const int h0 = hd * g.sh - g.ph, w0 = wd * g.sw - g.pw;
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sA = sbase + stage * HSW_STAGE_BYTES + offRow;
          if (p.debug & 1) {
            mbar_arrive(full_bar(stage));
            continue;
          }
          if (uniform) {
            const int hs = h0 + kh0, ws = w0 + kw0;
            const bool ok = mok && col0 < g.Kd && hs >= 0 && hs < g.Hs && ws >= 0 && ws < g.Ws;
            const long long eo = ok ? (((long long)bimg * g.Hs + hs) * g.Ws + ws) * g.Cs + ic0 : 0;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint32_t o = sA + (uint32_t)((c ^ (ir & 7)) << 4);
              hs_cp16(o, p.Ih + eo + (ok ? c * 8 : 0), ok);
              if (PL == 2) hs_cp16(o + HSW_A_BYTES, p.Il + eo + (ok ? c * 8 : 0), ok);
            }
          } else {
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
              const int col = col0 + c * 8;
              const int tap = col / g.Cs;
              const int ic = col - tap * g.Cs;
              const int kh = tap / g.KW, kw = tap - kh * g.KW;
              const int hs = h0 + kh, ws = w0 + kw;
              const bool ok = mok && col < g.Kd && hs >= 0 && hs < g.Hs && ws >= 0 && ws < g.Ws;
              const long long eo = ok ? (((long long)bimg * g.Hs + hs) * g.Ws + ws) * g.Cs + ic : 0;
              const uint32_t o = sA + (uint32_t)((c ^ (ir & 7)) << 4);
              hs_cp16(o, p.Ih + eo, ok);
              if (PL == 2) hs_cp16(o + HSW_A_BYTES, p.Il + eo, ok);
            }
          }
          hs_cp_arrive(full_bar(stage));
        }
        q0 += nst;
      }
    }
  } else if (warp >= 5 && warp < 13) {
    // ------------------------------------------------------------------ producers (pure cp.async)
    const int pt = threadIdx.x - 5 * 32;
    // In operand: pixel row ir = pt >> 4, chunk ic = pt & 15 (columns i0 + 8 ic .. +7), both planes
    const int ir = pt >> 4, ic = pt & 15;
    const uint32_t offI = (uint32_t)((ic >> 3) * 2048 + (ir >> 3) * 1024 + (ir & 7) * 128 + (((ic & 7) ^ (ir & 7)) << 4));
    // G operand: plane gp = pt >> 7 (0 hi, 1 lo), pixel row gr = (pt >> 3) & 15, chunk gc = pt & 7 (8 channels)
    const int gp = pt >> 7, gr = (pt >> 3) & 15, gc = pt & 7;
    const uint32_t offG = (uint32_t)((gr >> 3) * 1024 + (gr & 7) * 128 + ((gc ^ (gr & 7)) << 4));
    const __half* Gplane = gp ? p.Gl : p.Gh;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int split, i0, j0;
      decode_tile(tile, split, i0, j0);
      const int mb = split * p.m_per_split;
      const int me = min(g.M, mb + p.m_per_split);
      const int nst = stages_of(split);
      const int col = i0 + ic * 8;
      const bool iok = col < g.Kd;
      const int tap = iok ? col / g.Cs : 0;
      const int icn = col - tap * g.Cs;
      const int ikh = tap / g.KW, ikw = tap - ikh * g.KW;
      const int ch = j0 + gc * 8;
      const bool chok = ch < p.Ng;
      for (int st = 0; st < nst; ++st) {
        // gathered input pixel of this thread's In row
        const int mi = mb + st * HSW_ROWS + ir;
        bool okI = iok && mi < me;
        long long eoI = 0;
        if (okI) {
          const int bimg = mi / (g.Hd * g.Wd);
          const int rem = mi - bimg * (g.Hd * g.Wd);
          const int hd = rem / g.Wd, wd = rem - hd * g.Wd;
          const int hs = hd * g.sh - g.ph + ikh, ws = wd * g.sw - g.pw + ikw;
          okI = hs >= 0 && hs < g.Hs && ws >= 0 && ws < g.Ws;
          if (okI) eoI = (((long long)bimg * g.Hs + hs) * g.Ws + ws) * g.Cs + icn;
        }
        const int mg = mb + st * HSW_ROWS + gr;
        const bool okG = chok && mg < me;
        const long long eoG = okG ? (long long)mg * (p.G_ld ? p.G_ld : p.Ng) + ch : 0;
        mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t sA = sbase + stage * HSW_STAGE_BYTES;
        const uint32_t sB = sA + PL * HSW_A_BYTES;
        if (p.debug & 1) {
          mbar_arrive(full_bar(stage));
          if (++stage == HSW_STAGES) { stage = 0; phase ^= 1; }
          continue;
        }
        hs_cp16(sA + offI, p.Ih + eoI, okI);
        if (PL == 2) hs_cp16(sA + HSW_A_BYTES + offI, p.Il + eoI, okI);
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          if (s < NS && (PL == 2 || gp == 0)) {
            // slot s: group s >> 2, MN atom s & 3 inside the group's plane
            const uint32_t o = sB + (uint32_t)(gp * 2 * HSW_B_BYTES + s * 2048) + offG;
            hs_cp16(o, Gplane + (long long)s * p.G_slot + eoG, okG);
          }
        }
        hs_cp_arrive(full_bar(stage));
        if (++stage == HSW_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ MMA issuer
    // whole warp in the uniform loop, one elected lane issues the MMAs + commits of a stage (see gather_gemm_hs)
    {
      const int n0 = NS >= 4 ? 256 : 64 * NS;            // N of slot group 0
      const int n1 = NS > 4 ? 64 * (NS - 4) : 0;         // N of slot group 1
      const uint32_t idesc0 = hs_idesc(n0, 1, 1, PL == 1);
      const uint32_t idesc1 = hs_idesc(n1 > 0 ? n1 : 64, 1, 1, PL == 1);
      int stage = 0;
      uint32_t phase = 0, tphase = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int split, i0, j0;
        decode_tile(tile, split, i0, j0);
        const int nst = stages_of(split);
        for (int c0 = 0; c0 < max(nst, 1); c0 += HSW_FLUSH) {
          mbar_wait(tempty_bar, tphase ^ 1);  // previous chunk / tile drained
          tc_fence_after();
          const int cend = min(nst, c0 + HSW_FLUSH);
          for (int st = c0; st < cend; ++st) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sA = sbase + stage * HSW_STAGE_BYTES;
            const uint32_t sB = sA + PL * HSW_A_BYTES;
            const uint64_t dAh = hs_mnmajor_desc(sA, 2048, 1024), dAl = hs_mnmajor_desc(sA + HSW_A_BYTES, 2048, 1024);
            // the two slot groups accumulate into independent TMEM regions: interleave them so that consecutive
            // MMAs never depend on each other's accumulator
            const uint64_t dBh0 = hs_mnmajor_desc(sB, 2048, 1024), dBl0 = hs_mnmajor_desc(sB + 2 * HSW_B_BYTES, 2048, 1024);
            const uint64_t dBh1 = hs_mnmajor_desc(sB + HSW_B_BYTES, 2048, 1024);
            const uint64_t dBl1 = hs_mnmajor_desc(sB + 3 * HSW_B_BYTES, 2048, 1024);
            const uint32_t accf = st != c0 ? 1u : 0u;
            if (hs_elect_one()) {
              fence_async_proxy();  // cp.async (generic proxy) writes of the gathered tile -> tensor-core reads
              if (!(p.debug & 2)) {
                if (PL == 2) {
                  hs_mma_f16(tmem_base, dAl, dBh0, idesc0, accf);
                  if (n1 > 0) hs_mma_f16(tmem_base + 256u, dAl, dBh1, idesc1, accf);
                  hs_mma_f16(tmem_base, dAh, dBl0, idesc0, 1u);
                  if (n1 > 0) hs_mma_f16(tmem_base + 256u, dAh, dBl1, idesc1, 1u);
                }
                hs_mma_f16(tmem_base, dAh, dBh0, idesc0, PL == 2 ? 1u : accf);
                if (n1 > 0) hs_mma_f16(tmem_base + 256u, dAh, dBh1, idesc1, PL == 2 ? 1u : accf);
              }
              tc_commit(empty_bar(stage));
              if (st + 1 == cend) tc_commit(tfull_bar);
            }
            __syncwarp();
            if (++stage == HSW_STAGES) { stage = 0; phase ^= 1; }
          }
          if (cend <= c0 && hs_elect_one()) tc_commit(tfull_bar);  // empty tile: the epilogue still writes zeros
          __syncwarp();
          tphase ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 0-3, 13-16)
    const int egrp = warp >= 13 ? 1 : 0;  // slots with (s & 1) == egrp
    const int quad = warp & 3;
    const int sh_in = hs_shift_from_bits(__ldg(p.i_bits));
    uint32_t tphase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int split, i0, j0;
      decode_tile(tile, split, i0, j0);
      const int nst = stages_of(split);
      const int i = i0 + quad * 32 + lane;
      for (int c0 = 0; c0 < max(nst, 1); c0 += HSW_FLUSH) {
        mbar_wait(tfull_bar, tphase);
        tc_fence_after();
        for (int s = egrp; s < NS; s += 2) {
          const float inv = hs_pow2(-sh_in - hs_shift_from_bits(__ldg(p.g_bits + p.slot0 + s)));
          float* outp = p.partial + ((long long)split * p.nslots + s) * (long long)g.N * g.Kd;
#pragma unroll 1
          for (int cc = 0; cc < 64; cc += 16) {
            uint32_t r[16];
            tc_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(s * 64 + cc), r);
            if (i < g.Kd) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int n = j0 + cc + j;
                if (n < g.N) {
                  float* dst = outp + (long long)n * g.Kd + i;
                  float v = nst == 0 ? 0.f : __uint_as_float(r[j]) * inv;
                  if (c0 > 0) v += *dst;  // same thread wrote it in the previous chunk: fixed order
                  *dst = v;
                }
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(tempty_bar);
        tphase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
// shapes the half-split kernels accept (the engine decides *when* to use them)
static inline bool hs_gather_shape_ok(const Geom& g) {
  if (g.Cs % 8 != 0 || g.Nd % 4 != 0) return false;
  if ((long long)g.B * g.Hs * g.Ws * g.Cs >= (1LL << 31) - (1LL << 24)) return false;  // 32-bit row offsets
  return true;
}
static inline bool hs_wgrad_shape_ok(const Geom& g, int Ng) { return g.Cs % 8 == 0 && Ng % 8 == 0; }
// size (halves) of the weight image of an [N][Kd] matrix
static inline long long hs_image_halves(int Nd, int Kd, int planes = 2) {
  const int BN = tc_bn(Nd);
  return (long long)ceil_div(Nd, BN) * ceil_div(Kd, HS_BK) * (planes * BN * HS_BK);
}

static int hs_ready() {
  static int ready = -1;
  if (ready == -1) {
    ready = 0;
    if (tc_sm_count() > 0) {
      bool ok = true;
      auto attr = [&](auto* fn, int bytes) {
        ok = ok && cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess;
      };
      attr(gather_gemm_hs<128, 2>, HsCfg<128, 2>::SMEM_BYTES);
      attr(gather_gemm_hs<64, 2>, HsCfg<64, 2>::SMEM_BYTES);
      attr(gather_gemm_hs<128, 1>, HsCfg<128, 1>::SMEM_BYTES);
      attr(gather_gemm_hs<64, 1>, HsCfg<64, 1>::SMEM_BYTES);
      attr(wgrad_gemm_hs<2>, HswCfg<2>::SMEM_BYTES);
      attr(wgrad_gemm_hs<1>, HswCfg<1>::SMEM_BYTES);
      attr(gather_gemm_hs_stack<64, 2>, HsnCfg<2>::SMEM_BYTES);
      attr(gather_gemm_hs_stack<128, 2>, HsnCfg<2>::SMEM_BYTES);
      attr(gather_gemm_hs_stack<64, 1>, HsnCfg<1>::SMEM_BYTES);
      attr(gather_gemm_hs_stack<128, 1>, HsnCfg<1>::SMEM_BYTES);
      ready = ok ? 1 : 0;
    }
  }
  return ready;
}

static inline int hs_grid(long long total, int threads = 256) {
  long long b = (total + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > 148 * 8) b = 148 * 8;
  return (int)b;
}

// absmax of nslots slots (x = first slot, n elements each) into bits[0..nslots); bits must have been zeroed
static inline int hs_launch_absmax(const float* x, long long slot_stride, long long n, uint32_t* bits,
                                   int nslots, cudaStream_t st) {
  hs_absmax_kernel<<<dim3(hs_grid(n / 4), nslots), 256, 0, st>>>(x, slot_stride, n / 4, bits);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
// split nslots slots (x = first slot, scale bits[slot]) into planes
static inline int hs_launch_split(const float* x, long long slot_stride, long long n, __half* hi, __half* lo,
                                  long long out_slot_stride, const uint32_t* bits, int nslots, cudaStream_t st) {
  hs_split_kernel<<<dim3(hs_grid(n / 8), nslots), 256, 0, st>>>(x, slot_stride, hi, lo, out_slot_stride, n / 8,
                                                                bits);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
// pack [N][Kd] fp32 (nslots matrices src_slot apart, scales bits[slot]) into weight images
static inline int hs_launch_pack_image(const float* src, long long src_slot, __half* dst, long long dst_slot,
                                       int N, int Nd, int Kd, int nslots, const uint32_t* bits,
                                       cudaStream_t st, int planes = 2) {
  const int BN = tc_bn(Nd);
  const int tiles_n = ceil_div(Nd, BN), nchunks = ceil_div(Kd, HS_BK);
  const long long total = (long long)tiles_n * nchunks * BN * 8;
  hs_pack_image_kernel<<<dim3(hs_grid(total), nslots), 256, 0, st>>>(src, src_slot, dst, dst_slot, N, Kd, BN,
                                                                    tiles_n, nchunks, bits, planes);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// images of the K tangent weights of a conv [N][C][taps] from the K columns of V (rows = N*C*taps, stride ldk);
// bits[k] receives / holds the absmax of column k (computed here unless planes == 1: bf16 planes are unscaled)
static inline int hs_launch_pack_image_cols(const float* V, int ldk, int K, __half* dst, long long dst_slot, int N,
                                            int C, int taps, int Cp, int Nd, uint32_t* bits, cudaStream_t st,
                                            int planes) {
  if (K < 1 || K > 8) return 1;
  const int Kd = taps * Cp;
  const int BN = tc_bn(Nd);
  const int tiles_n = ceil_div(Nd, BN), nchunks = ceil_div(Kd, HS_BK);
  const long long rows = (long long)N * C * taps;
  if (planes != 1) hs_vcol_absmax_kernel<<<hs_grid(rows), 256, 0, st>>>(V, rows, ldk, K, bits);
  const long long total = (long long)tiles_n * nchunks * BN * 8;
  hs_pack_image_cols_kernel<<<hs_grid(total), 256, 0, st>>>(V, ldk, K, dst, dst_slot, N, C, taps, Cp, BN, tiles_n,
                                                           nchunks, bits, planes);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// returns 0 on success, >0 on a CUDA error, <0 if unavailable
typedef CUresult (*hs_encode_im2col_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline hs_encode_im2col_fn hs_encode_im2col() {
  static hs_encode_im2col_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<hs_encode_im2col_fn>(ptr);
    (void)cudaGetLastError();
  }
  return fn;
}
// im2col map over an fp16 plane [images][H][W][C]: box = 64 channels x 128 pixels, base pixels inside
// [lower, dim + upper) traversed with the given strides
static inline bool hs_make_im2col_map(CUtensorMap* map, const __half* base, int C, int W, int H, long long images,
                                      int lw, int lh, int uw, int uh, int sw, int sh, bool bf16 = false) {
  hs_encode_im2col_fn enc = hs_encode_im2col();
  if (!enc) return false;
  if (lw < -128 || lw > 127 || lh < -128 || lh > 127 || uw < -128 || uw > 127 || uh < -128 || uh > 127) return false;
  if (sw < 1 || sw > 8 || sh < 1 || sh > 8 || images < 1 || images >= (1LL << 31)) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)images};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const int lower[2] = {lw, lh}, upper[2] = {uw, uh};
  const cuuint32_t estr[4] = {1, (cuuint32_t)sw, (cuuint32_t)sh, 1};
  return enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
             const_cast<__half*>(base), dims, strides, lower, upper, 64,
             TC_BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// fills the TMA fields of a (parity-resolved) argument block; false -> the cp.async producers are used
static inline bool hs_setup_gather_tma(HsGatherArgs& a, int a_plane_slots) {
  const Geom& g = a.g;
  a.use_tma = 0;
  if (g.Cs % HS_BK != 0 || a.A_slot != (long long)g.B * g.Hs * g.Ws * g.Cs) return false;
  const long long images = (long long)a_plane_slots * g.B;
  const int npl = a.planes == 1 ? 1 : 2;
  if (a.par.nclass > 0) {
    for (int c = 0; c < a.par.nclass; ++c) {
      if (a.par.ntap[c] == 0) continue;  // class without taps: never loaded
      int lh = 127, lw = 127;
      for (int t = 0; t < a.par.ntap[c]; ++t) { lh = std::min(lh, (int)a.par.dh[c][t]); lw = std::min(lw, (int)a.par.dw[c][t]); }
      for (int t = 0; t < a.par.ntap[c]; ++t) {
        a.tma_offh[c][t] = (unsigned char)(a.par.dh[c][t] - lh);
        a.tma_offw[c][t] = (unsigned char)(a.par.dw[c][t] - lw);
      }
      a.tma_w0[c] = lw; a.tma_h0[c] = lh;
      const int uw = a.par.Wc[c] - g.Ws + lw, uh = a.par.Hc[c] - g.Hs + lh;
      for (int pl = 0; pl < npl; ++pl)
        if (!hs_make_im2col_map(&a.tmA[c][pl], pl ? a.Al : a.Ah, g.Cs, g.Ws, g.Hs, images, lw, lh, uw, uh, 1, 1,
                                npl == 1))
          return false;
    }
    a.tma_sw = 1; a.tma_sh = 1;
  } else if (g.mode == 0) {
    if (g.KH * g.KW > HS_PAR_TAPS) return false;
    for (int kh = 0; kh < g.KH; ++kh)
      for (int kw = 0; kw < g.KW; ++kw) { a.tma_offh[0][kh * g.KW + kw] = (unsigned char)kh; a.tma_offw[0][kh * g.KW + kw] = (unsigned char)kw; }
    a.tma_w0[0] = -g.pw; a.tma_h0[0] = -g.ph; a.tma_sw = g.sw; a.tma_sh = g.sh;
    for (int pl = 0; pl < npl; ++pl)
      if (!hs_make_im2col_map(&a.tmA[0][pl], pl ? a.Al : a.Ah, g.Cs, g.Ws, g.Hs, images, -g.pw, -g.ph,
                              g.pw - (g.KW - 1), g.ph - (g.KH - 1), g.sw, g.sh, npl == 1))
        return false;
  } else if (g.sh == 1 && g.sw == 1) {  // stride-1 dgrad: source = dest + pad - tap = (dest + lower) + (K-1-tap)
    if (g.KH * g.KW > HS_PAR_TAPS) return false;
    const int lw = -(g.KW - 1 - g.pw), lh = -(g.KH - 1 - g.ph);
    for (int kh = 0; kh < g.KH; ++kh)
      for (int kw = 0; kw < g.KW; ++kw) {
        a.tma_offh[0][kh * g.KW + kw] = (unsigned char)(g.KH - 1 - kh);
        a.tma_offw[0][kh * g.KW + kw] = (unsigned char)(g.KW - 1 - kw);
      }
    a.tma_w0[0] = lw; a.tma_h0[0] = lh; a.tma_sw = 1; a.tma_sh = 1;
    for (int pl = 0; pl < npl; ++pl)
      if (!hs_make_im2col_map(&a.tmA[0][pl], pl ? a.Al : a.Ah, g.Cs, g.Ws, g.Hs, images, lw, lh, g.Wd - g.Ws + lw,
                              g.Hd - g.Hs + lh, 1, 1, npl == 1))
        return false;
  } else {
    return false;
  }
  a.use_tma = 1;
  return true;
}

// producer mode of the gathered operand when TMA applies: 1 = both planes by TMA im2col, 2 = hybrid (hi plane by
// TMA, lo plane by cp.async), 0 = cp.async only
static int g_hs_gather_tma_mode = 1;
// K stages (64 reduction elements each) per TMEM accumulation chunk of gather_gemm_hs: a chunk is one chain of
// 12 * flush truncating tensor-core accumulations; chunks are summed in registers with round-to-nearest adds.
// Environment override CURV_HS_FLUSH (accuracy / speed experiments).
static int hs_flush() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("CURV_HS_FLUSH");
    v = e ? atoi(e) : HS_FLUSH;
    if (v < 1 || v > 64) v = HS_FLUSH;
  }
  return v;
}

// a_plane_slots: number of slots stored in the planes (bounds the TMA tensor); 0 -> no TMA
static inline int hs_launch_gather_gemm(const HsGatherArgs& a_in, int nslots, cudaStream_t st,
                                        bool allow_parity = true, int a_plane_slots = 0) {
  if (hs_ready() <= 0) return -1;
  const int sms = tc_sm_count();
  HsGatherArgs a = a_in;
  const Geom& g = a.g;
  if (!allow_parity || !hs_make_parity(g, a.par)) a.par.nclass = 0;
  a.use_tma = 0;
  a.flush = a.planes == 1 ? 64 : hs_flush();  // bf16: truncation bias is far below the format's own rounding
  if (a_plane_slots > 0 && g_hs_gather_tma_mode > 0 && hs_setup_gather_tma(a, a_plane_slots))
    a.use_tma = g_hs_gather_tma_mode;
  const int tiles_m = a.par.nclass > 0 ? a.par.tile0[a.par.nclass] : ceil_div(g.M, TC_BM);
  if (tc_bn(g.Nd) == 128) {
    const int ntiles = tiles_m * ceil_div(g.Nd, 128) * nslots;
    if (a.planes == 1)
      gather_gemm_hs<128, 1><<<ntiles < sms ? ntiles : sms, TC_THREADS, HsCfg<128, 1>::SMEM_BYTES, st>>>(a, nslots);
    else
      gather_gemm_hs<128, 2><<<ntiles < sms ? ntiles : sms, TC_THREADS, HsCfg<128, 2>::SMEM_BYTES, st>>>(a, nslots);
  } else {
    const int ntiles = tiles_m * ceil_div(g.Nd, 64) * nslots;
    if (a.planes == 1)
      gather_gemm_hs<64, 1><<<ntiles < sms ? ntiles : sms, TC_THREADS, HsCfg<64, 1>::SMEM_BYTES, st>>>(a, nslots);
    else
      gather_gemm_hs<64, 2><<<ntiles < sms ? ntiles : sms, TC_THREADS, HsCfg<64, 2>::SMEM_BYTES, st>>>(a, nslots);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point query: no link-time dependency on libcuda,
// so the library still loads (and reports "no CUDA device") on a machine without a driver
typedef CUresult (*hs_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                       CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                       CUtensorMapFloatOOBfill);
static inline hs_encode_tiled_fn hs_encode_tiled() {
  static hs_encode_tiled_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<hs_encode_tiled_fn>(ptr);
    (void)cudaGetLastError();
  }
  return fn;
}

// 3-d tiled tensor map over an fp16 plane [slots][rows][width] with box (64, 16, slots), SWIZZLE_128B
static inline bool hs_make_plane_map(CUtensorMap* map, const __half* base, int width, long long rows, int slots,
                                     long long slot_stride_elems, bool bf16 = false, long long row_stride_elems = 0) {
  hs_encode_tiled_fn enc = hs_encode_tiled();
  if (!enc) return false;
  if (row_stride_elems == 0) row_stride_elems = width;
  if ((row_stride_elems * 2) % 16 != 0) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)rows, (cuuint64_t)slots};
  const cuuint64_t strides[2] = {(cuuint64_t)row_stride_elems * 2, (cuuint64_t)slot_stride_elems * 2};
  const cuuint32_t box[3] = {64, (cuuint32_t)HSW_ROWS, (cuuint32_t)slots};
  const cuuint32_t estr[3] = {1, 1, 1};
  if (slots > 1 && strides[1] % 16 != 0) return false;
  return enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
             const_cast<__half*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// shared-activation segment of slots [slot_lo, slot_lo + nslots): returns 0 on success, >0 on a CUDA error
static inline int hs_launch_gather_stack(const HsStackArgs& a_in, cudaStream_t st) {
  if (hs_ready() <= 0) return -1;
  const int sms = tc_sm_count();
  HsStackArgs a = a_in;
  a.flush = a.planes == 1 ? 64 : hs_flush();
  const Geom& g = a.g;
  if (g.mode != 0 || g.Cs % 8 != 0 || a.nslots < 1) return -1;
  if (a.planes != 1 && ceil_div(g.Kd, HS_BK) <= 8 && !getenv("CURV_HS_FLUSH")) a.flush = 8;  // short reductions: one main chunk per tile
  const int BN = tc_bn(g.Nd), G = 256 / BN;
  const int ntiles = ceil_div(g.M, TC_BM) * ceil_div(g.Nd, BN) * ceil_div(a.nslots, G);
  const int grid = ntiles < sms ? ntiles : sms;
  if (a.planes == 1) {
    if (BN == 128) gather_gemm_hs_stack<128, 1><<<grid, HSN_THREADS, HsnCfg<1>::SMEM_BYTES, st>>>(a);
    else gather_gemm_hs_stack<64, 1><<<grid, HSN_THREADS, HsnCfg<1>::SMEM_BYTES, st>>>(a);
  } else {
    if (BN == 128) gather_gemm_hs_stack<128, 2><<<grid, HSN_THREADS, HsnCfg<2>::SMEM_BYTES, st>>>(a);
    else gather_gemm_hs_stack<64, 2><<<grid, HSN_THREADS, HsnCfg<2>::SMEM_BYTES, st>>>(a);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

static inline int hs_launch_wgrad(const HsWgradArgs& a_in, cudaStream_t st, bool allow_tma = true) {
  if (hs_ready() <= 0) return -1;
  if (a_in.nslots < 1 || a_in.nslots > 8) return -1;
  const int sms = tc_sm_count();
  HsWgradArgs a = a_in;
  a.use_tma = 0;
  const bool bf = a.planes == 1;
  if (allow_tma && a.g.M >= HSW_ROWS &&
      hs_make_plane_map(&a.tmGh, a.Gh, a.Ng, a.g.M, a.nslots, a.G_slot, bf, a.G_ld) &&
      (bf || hs_make_plane_map(&a.tmGl, a.Gl, a.Ng, a.g.M, a.nslots, a.G_slot, false, a.G_ld)))
    a.use_tma = 1;
  const int ntiles = ceil_div(a.g.Kd, TC_BM) * ceil_div(a.Ng, 64) * a.nsplit;
  if (bf) wgrad_gemm_hs<1><<<ntiles < sms ? ntiles : sms, TC_THREADS, HswCfg<1>::SMEM_BYTES, st>>>(a);
  else wgrad_gemm_hs<2><<<ntiles < sms ? ntiles : sms, TC_THREADS, HswCfg<2>::SMEM_BYTES, st>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

#endif  // CURV_DISABLE_TC

}  // namespace curv
