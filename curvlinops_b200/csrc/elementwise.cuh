// Streaming (HBM-bound) kernels of the layer program: layout conversion, parameter packing,
// BatchNorm(eval)/activation/pooling/residual forward+tangent and adjoint passes, loss-Hessian
// apply and the deterministic finishing reductions.  All tensors are [slot][rows][Cp] fp32 with
// Cp % 4 == 0, so every access is a coalesced float4.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace curv {

enum ActKind { ACT_RELU = 0, ACT_SIGMOID = 1, ACT_TANH = 2, ACT_GELU = 3 };  // GELU: exact (erf) form

__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4add(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 f4mul(float4 a, float4 b) {
  return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}
__device__ __forceinline__ float4 f4fma(float4 a, float4 b, float4 c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}

// Fused absolute-maximum tracking for the half-split GEMM path (hs_gemm.cuh): a kernel that produces a GEMM operand
// folds max|value| of everything it wrote into dst (uint bit pattern of a non-negative float; atomicMax is order
// independent, hence deterministic).  One atomic per CTA.  Must be reached by all threads of the block.
__device__ __forceinline__ float f4absmax(float m, const float4& v) {
  return fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
}
__device__ __forceinline__ void block_absmax_commit(float m, unsigned int* dst) {
  __shared__ float absmax_red[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) absmax_red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 1; w < nw; ++w) m = fmaxf(m, absmax_red[w]);
    // thousands of CTAs target the same word: only those that would raise it pay for the atomic
    if (m > 0.f && __float_as_uint(m) > *reinterpret_cast<volatile unsigned int*>(dst))
      atomicMax(dst, __float_as_uint(m));
  }
}

// fp16 hi/lo planes of 4 scaled floats at float4-index i of a plane pair.  Adjacent lanes (i even / odd, both
// active: the element count is even and warps are index-aligned) exchange halves so that every lane issues ONE
// 16-byte store (even lanes the hi plane, odd lanes the lo plane) instead of two 8-byte ones.
__device__ __forceinline__ void store_split_planes(__half* __restrict__ ph, __half* __restrict__ pl, long long i,
                                                   const float4& r, float sc) {
  // packed conversions: one cvt.rn.f16x2.f32 per pair (the kernels that call this are issue-bound otherwise)
  const float2 a = make_float2(r.x * sc, r.y * sc), b = make_float2(r.z * sc, r.w * sc);
  if (pl == nullptr) {  // bf16 mode: ONE plane of bf16 values (one 8-byte store per lane, adjacent lanes contiguous)
    const __nv_bfloat162 ba = __floats2bfloat162_rn(a.x, a.y), bb = __floats2bfloat162_rn(b.x, b.y);
    reinterpret_cast<uint2*>(ph)[i] = make_uint2(*reinterpret_cast<const unsigned int*>(&ba),
                                                 *reinterpret_cast<const unsigned int*>(&bb));
    return;
  }
  const __half2 ha = __float22half2_rn(a), hb2 = __float22half2_rn(b);
  const float2 fa = __half22float2(ha), fb = __half22float2(hb2);
  const __half2 la = __float22half2_rn(make_float2(a.x - fa.x, a.y - fa.y));
  const __half2 lb2 = __float22half2_rn(make_float2(b.x - fb.x, b.y - fb.y));
  const unsigned int h0 = *reinterpret_cast<const unsigned int*>(&ha), h1 = *reinterpret_cast<const unsigned int*>(&hb2);
  const unsigned int l0 = *reinterpret_cast<const unsigned int*>(&la), l1 = *reinterpret_cast<const unsigned int*>(&lb2);
  const unsigned int mask = __activemask();
  const bool odd = (i & 1) != 0;
  // even lane sends its lo words and receives the neighbour's hi words; odd lane the other way round
  const unsigned int s0 = odd ? h0 : l0, s1 = odd ? h1 : l1;
  const unsigned int r0 = __shfl_xor_sync(mask, s0, 1), r1 = __shfl_xor_sync(mask, s1, 1);
  if (!odd) reinterpret_cast<uint4*>(ph)[i >> 1] = make_uint4(h0, h1, r0, r1);
  else reinterpret_cast<uint4*>(pl)[i >> 1] = make_uint4(r0, r1, l0, l1);
}

// ------------------------------------------------------------------ layout / packing
// X [B][C][H][W] -> out [B][H][W][Cp]
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ X, float* __restrict__ out, int B, int C,
                                    int H, int W, int Cp) {
  long long total = (long long)B * H * W * Cp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % Cp);
    long long pix = i / Cp;
    int w = (int)(pix % W);
    long long r = pix / W;
    int h = (int)(r % H);
    int b = (int)(r / H);
    out[i] = c < C ? __ldg(X + (((long long)b * C + c) * H + h) * W + w) : 0.f;
  }
}

// out[b][c][k] (ldk) <- act[slot k][b][c]   (JVP output)
__global__ void export_pred_kernel(const float* __restrict__ act, long long slot_stride, int slot0,
                                   float* __restrict__ out, int B, int C, int Cp, int K, int ldk, int k0) {
  long long total = (long long)B * C * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int k = (int)(i % K);
    long long r = i / K;
    int c = (int)(r % C);
    int b = (int)(r / C);
    out[((long long)b * C + c) * ldk + k0 + k] = act[(slot0 + k) * slot_stride + (long long)b * Cp + c];
  }
}
// act[slot k][b][c] <- in[b][c][k]  (VJP seed), pad lanes zeroed
__global__ void import_pred_kernel(float* __restrict__ act, long long slot_stride, int slot0,
                                   const float* __restrict__ in, int B, int C, int Cp, int K, int ldk, int k0) {
  long long total = (long long)B * Cp * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % Cp);
    long long r = i / Cp;
    int b = (int)(r % B);
    int k = (int)(r / B);
    act[(slot0 + k) * slot_stride + (long long)b * Cp + c] =
        c < C ? __ldg(in + ((long long)b * C + c) * ldk + k0 + k) : 0.f;
  }
}

// Weight [N][C][KH][KW] (element stride `es`, e.g. ldk for a column of V) ->
//   Wk [n][tap][cp]   (rows = N,  row length = taps*Cp)      if transposed == 0
//   Wt [c][tap][np]   (rows = C,  row length = taps*Np)      if transposed == 1
// grid.y = column index k (src += k, dst += k*dst_slot)
__global__ void pack_weight_kernel(const float* __restrict__ src, long long es, float* __restrict__ dst,
                                   long long dst_slot, int N, int C, int KH, int KW, int Cp, int Np,
                                   int transposed) {
  const int taps = KH * KW;
  const long long total = transposed ? (long long)C * taps * Np : (long long)N * taps * Cp;
  src += blockIdx.y;
  dst += blockIdx.y * dst_slot;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int n, c, tap;
    if (!transposed) {
      c = (int)(i % Cp); long long r = i / Cp; tap = (int)(r % taps); n = (int)(r / taps);
    } else {
      n = (int)(i % Np); long long r = i / Np; tap = (int)(r % taps); c = (int)(r / taps);
    }
    float v = 0.f;
    if (n < N && c < C) v = __ldg(src + (((long long)n * C + c) * taps + tap) * es);
    dst[i] = v;
  }
}

// dst[i] = i < n ? v : 0 for i < np
__global__ void fill_kernel(float* __restrict__ dst, int n, float v, int np) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < np; i += gridDim.x * blockDim.x) dst[i] = i < n ? v : 0.f;
}

// vector [n] (element stride es) -> dst[k][Np] zero padded.  grid.y = k
__global__ void pack_vec_kernel(const float* __restrict__ src, long long es, float* __restrict__ dst,
                                long long dst_slot, int N, int Np) {
  src += blockIdx.y;
  dst += blockIdx.y * dst_slot;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Np; i += gridDim.x * blockDim.x)
    dst[i] = i < N ? __ldg(src + (long long)i * es) : 0.f;
}

// BatchNorm(eval) coefficients.  coef layout: [(1+K)][2][Cp]:
//   slot 0: (s, t) with s = gamma*invstd, t = beta - mean*s
//   slot k: (sdot_k, tdot_k) = (gdot_k*invstd, bdot_k - mean*sdot_k)
// aux [2][Cp] = (invstd, mean).   gamma/beta null -> 1 / 0.  gdot/bdot: columns of V (stride ldk).
__global__ void affine_prep_kernel(const float* gamma, const float* beta, const float* mean,
                                   const float* var, float eps, const float* gdot, const float* bdot,
                                   long long ldk, int K, int C, int Cp, float* coef, float* aux,
                                   unsigned int* smax_bits) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < Cp; c += gridDim.x * blockDim.x) {
    float invstd = 0.f, mu = 0.f, s = 0.f, tt = 0.f;
    if (c < C) {
      invstd = 1.0f / sqrtf(var[c] + eps);
      mu = mean[c];
      float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
      s = g * invstd;
      tt = b - mu * s;
    }
    coef[c] = s; coef[Cp + c] = tt;
    aux[c] = invstd; aux[Cp + c] = mu;
    // coefficient maxima for the bounds of the planes modes: words [0] = max|s|, [1] = max|t|, [2k], [2k+1] =
    // max|sdot_k|, max|tdot_k|
    if (smax_bits && s != 0.f) atomicMax(smax_bits, __float_as_uint(fabsf(s)));
    if (smax_bits && tt != 0.f) atomicMax(smax_bits + 1, __float_as_uint(fabsf(tt)));
    for (int k = 0; k < K; ++k) {
      float sd = 0.f, td = 0.f;
      if (c < C) {
        sd = gdot ? gdot[(long long)c * ldk + k] * invstd : 0.f;
        td = (bdot ? bdot[(long long)c * ldk + k] : 0.f) - mu * sd;
      }
      coef[(long long)(1 + k) * 2 * Cp + c] = sd;
      coef[(long long)(1 + k) * 2 * Cp + Cp + c] = td;
      if (smax_bits && sd != 0.f) atomicMax(smax_bits + 2 * (k + 1), __float_as_uint(fabsf(sd)));
      if (smax_bits && td != 0.f) atomicMax(smax_bits + 2 * (k + 1) + 1, __float_as_uint(fabsf(td)));
    }
  }
}

// ------------------------------------------------------------------ forward + tangent
// y_0 = s*x_0 + t ;  y_k = s*x_k + sdot_k*x_0 + tdot_k     grid.y = slot
// grid.y = slot GROUP: group g handles the tangent slots 1 + 8g .. 8 + 8g (and, for g = 0, the primal slot 0), so the
// primal input - needed by every slot for the tangent of the coefficients and for the ReLU mask - is read once per
// group instead of once per slot.
template <bool PLANES>
__global__ void __launch_bounds__(256, PLANES ? 3 : 4) affine_fwd_kernel(const float* __restrict__ x, long long x_slot, int x_has_slots,
                                                         const float* __restrict__ coef, int coef_has_tan,
                                                         float* __restrict__ y, long long y_slot, long long rows,
                                                         int Cp, int relu, int nslots,
                                                         unsigned int* __restrict__ amax, __half* __restrict__ ph,
                                                         __half* __restrict__ pl, long long plane_slot,
                                                         const unsigned int* __restrict__ in_bits,
                                                         const unsigned int* __restrict__ cbits,
                                                         int write_fp32_tangents) {
  // Planes mode (ph != null; half-split GEMM path): the output is ALSO written as the fp16 hi/lo planes the next
  // convolution gathers (slot s at ph/pl + s * plane_slot), scaled by a power of two derived from the bound
  //   |y_0| <= max|s| max|x_0| + max|t|,   |y_k| <= max|s| max|x_k| + max|sdot_k| max|x_0| + max|tdot_k|
  // (in_bits: absmax words of x by slot, cbits: coefficient maxima of affine_prep_kernel); the bound's bit pattern
  // replaces the exact maximum in amax[slot].  The fp32 tangent slots are skipped when nobody else reads them.
  const int grp = blockIdx.y;
  const int k_lo = 1 + 8 * grp;  // first tangent slot of the group
  const int C4 = Cp >> 2;
  const long long total = rows * C4;
  const long long xs4 = x_slot >> 2, ys4 = y_slot >> 2;
  const float4* x0 = reinterpret_cast<const float4*>(x);
  float4* y0 = reinterpret_cast<float4*>(y);
  const float4* s4 = reinterpret_cast<const float4*>(coef);
  const float4* t4 = reinterpret_cast<const float4*>(coef + Cp);
  float am[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) am[j] = 0.f;
  float psc[9];  // plane scales of slot 0 and of the group's tangent slots
  if (PLANES) {
    const float smax = __uint_as_float(__ldg(cbits)), x0max = __uint_as_float(__ldg(in_bits));
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const int slot = j == 0 ? 0 : k_lo + j - 1;
      float bound = 0.f;
      if (slot == 0) bound = smax * x0max + __uint_as_float(__ldg(cbits + 1));
      else if (slot < nslots)
        bound = (x_has_slots ? smax * __uint_as_float(__ldg(in_bits + slot)) : 0.f) +
                (coef_has_tan ? __uint_as_float(__ldg(cbits + 2 * slot)) * x0max + __uint_as_float(__ldg(cbits + 2 * slot + 1))
                              : 0.f);
      psc[j] = hs_pow2(hs_shift_from_bits(__float_as_uint(bound)));
      if (blockIdx.x == 0 && threadIdx.x == 0 && slot < nslots && (j > 0 || grp == 0)) amax[slot] = __float_as_uint(bound);
    }
  }
  auto store_planes = [&](int slot, long long i, const float4& r, float sc) {
    store_split_planes(ph + (long long)slot * plane_slot, pl ? pl + (long long)slot * plane_slot : nullptr, i, r, sc);
  };
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    const float4 sv = __ldg(s4 + c4), xv = __ldg(x0 + i);
    const float4 pre = f4fma(sv, xv, __ldg(t4 + c4));
    if (grp == 0) {
      const float4 r = relu ? make_float4(fmaxf(pre.x, 0.f), fmaxf(pre.y, 0.f), fmaxf(pre.z, 0.f), fmaxf(pre.w, 0.f))
                            : pre;
      y0[i] = r;
      if (PLANES) store_planes(0, i, r, psc[0]);
      else am[0] = f4absmax(am[0], r);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int slot = k_lo + j;
      if (slot < nslots) {
        float4 r = f4zero();
        if (x_has_slots) r = f4mul(sv, __ldg(x0 + slot * xs4 + i));
        if (coef_has_tan) {
          const float4* sd4 = reinterpret_cast<const float4*>(coef + (long long)slot * 2 * Cp);
          r = f4add(r, f4fma(__ldg(sd4 + c4), xv, __ldg(sd4 + C4 + c4)));
        }
        if (relu)  // fused ReLU: mask with the sign of the primal pre-activation
          r = make_float4(pre.x > 0.f ? r.x : 0.f, pre.y > 0.f ? r.y : 0.f, pre.z > 0.f ? r.z : 0.f,
                          pre.w > 0.f ? r.w : 0.f);
        if (!PLANES || write_fp32_tangents) y0[slot * ys4 + i] = r;
        if (PLANES) store_planes(slot, i, r, psc[1 + j]);
        else am[1 + j] = f4absmax(am[1 + j], r);
      }
    }
  }
  if (amax && !PLANES) {
    if (grp == 0) block_absmax_commit(am[0], amax);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      __syncthreads();
      if (k_lo + j < nslots) block_absmax_commit(am[1 + j], amax + k_lo + j);
    }
  }
}

__device__ __forceinline__ float act_apply(int kind, float x) {
  if (kind == ACT_RELU) return fmaxf(x, 0.f);
  if (kind == ACT_SIGMOID) return 1.f / (1.f + expf(-x));
  if (kind == ACT_GELU) return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
  return tanhf(x);
}
// derivative expressed through the OUTPUT y = phi(x) -- except GELU, whose derivatives take the INPUT x (the callers
// hand these functions the primal input for that kind)
__device__ __forceinline__ float act_d1(int kind, float y) {
  if (kind == ACT_GELU)  // Phi(x) + x phi(x)
    return 0.5f * (1.f + erff(y * 0.70710678118654752f)) + y * 0.3989422804014327f * expf(-0.5f * y * y);
  if (kind == ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (kind == ACT_SIGMOID) return y * (1.f - y);
  return 1.f - y * y;
}
__device__ __forceinline__ float act_d2(int kind, float y) {
  if (kind == ACT_RELU) return 0.f;
  if (kind == ACT_GELU) return 0.3989422804014327f * expf(-0.5f * y * y) * (2.f - y * y);  // phi(x) (2 - x^2)
  if (kind == ACT_SIGMOID) return y * (1.f - y) * (1.f - 2.f * y);
  return -2.f * y * (1.f - y * y);
}

// y_0 = phi(x_0);  y_k = phi'(x_0) * x_k.   Slot 0 must be computed before slots >= 1:
// launch 1 handles slot 0 (grid.y = 1, slot0 = 0), launch 2 the tangents (slot0 = 1).
__global__ void act_fwd_kernel(int kind, const float* __restrict__ x, long long x_slot,
                               float* __restrict__ y, long long y_slot, long long n4, int slot0) {
  const int slot = slot0 + blockIdx.y;
  const float4* xs = reinterpret_cast<const float4*>(x + slot * x_slot);
  const float4* y0 = reinterpret_cast<const float4*>(kind == ACT_GELU ? x : y);  // what act_d1 is evaluated at
  float4* yo = reinterpret_cast<float4*>(y + slot * y_slot);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 v = __ldg(xs + i);
    if (slot == 0) {
      v = make_float4(act_apply(kind, v.x), act_apply(kind, v.y), act_apply(kind, v.z), act_apply(kind, v.w));
    } else {
      float4 p = y0[i];
      v = make_float4(v.x * act_d1(kind, p.x), v.y * act_d1(kind, p.y), v.z * act_d1(kind, p.z),
                      v.w * act_d1(kind, p.w));
    }
    yo[i] = v;
  }
}

// gx_k (+)= phi'(y_0) * gy_k   [+ phi''(y_0) * xdot_k * gy_0  (R-op, slot >= 1, kind != relu)]
__global__ void act_bwd_kernel(int kind, const float* __restrict__ gy, long long gy_slot,
                               const float* __restrict__ y0, float* __restrict__ gx, long long gx_slot,
                               const float* __restrict__ xdot, long long xdot_slot,
                               long long n4, int slot0, int accumulate) {
  const int slot = slot0 + blockIdx.y;
  const float4* g = reinterpret_cast<const float4*>(gy + slot * gy_slot);
  const float4* g0 = reinterpret_cast<const float4*>(gy);
  const float4* p4 = reinterpret_cast<const float4*>(y0);
  const float4* xd = xdot ? reinterpret_cast<const float4*>(xdot + slot * xdot_slot) : nullptr;
  float4* o = reinterpret_cast<float4*>(gx + slot * gx_slot);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 v = __ldg(g + i), p = __ldg(p4 + i);
    float4 r = make_float4(v.x * act_d1(kind, p.x), v.y * act_d1(kind, p.y), v.z * act_d1(kind, p.z),
                           v.w * act_d1(kind, p.w));
    if (xd != nullptr && slot > 0 && kind != ACT_RELU) {
      float4 d = __ldg(xd + i), q = __ldg(g0 + i);
      r.x += act_d2(kind, p.x) * d.x * q.x; r.y += act_d2(kind, p.y) * d.y * q.y;
      r.z += act_d2(kind, p.z) * d.z * q.z; r.w += act_d2(kind, p.w) * d.w * q.w;
    }
    if (accumulate) r = f4add(r, o[i]);
    o[i] = r;
  }
}

// dst_slot (+)= scale * src_slot  (residual add forward/backward, copies).  grid.y = slot
__global__ void axpy_slots_kernel(const float* __restrict__ src, long long src_slot,
                                  float* __restrict__ dst, long long dst_slot, long long n4, int slot0,
                                  float scale, int accumulate) {
  const int slot = slot0 + blockIdx.y;
  const float4* s = reinterpret_cast<const float4*>(src + slot * src_slot);
  float4* d = reinterpret_cast<float4*>(dst + slot * dst_slot);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 v = __ldg(s + i);
    v = make_float4(v.x * scale, v.y * scale, v.z * scale, v.w * scale);
    if (accumulate) v = f4add(v, d[i]);
    d[i] = v;
  }
}

// fused residual join + ReLU:  y_0 = relu(a_0 + b_0),  y_k = [a_0 + b_0 > 0] (a_k + b_k).  grid.y = slot
// grid.y = slot group (see affine_fwd_kernel): the primal sum a_0 + b_0 (the ReLU mask) is formed once per group
__global__ void __launch_bounds__(256, 4) add_relu_fwd_kernel(const float* __restrict__ a, long long a_slot,
                                                           int a_has_slots, const float* __restrict__ b,
                                                           long long b_slot, int b_has_slots, float* __restrict__ y,
                                                           long long y_slot, long long n4, int nslots,
                                                           unsigned int* __restrict__ amax) {
  const int grp = blockIdx.y;
  const int k_lo = 1 + 8 * grp;
  const long long as4 = a_slot >> 2, bs4 = b_slot >> 2, ys4 = y_slot >> 2;
  const float4* a0 = reinterpret_cast<const float4*>(a);
  const float4* b0 = reinterpret_cast<const float4*>(b);
  float4* y0 = reinterpret_cast<float4*>(y);
  float am[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) am[j] = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 p = f4add(__ldg(a0 + i), __ldg(b0 + i));
    if (grp == 0) {
      const float4 r = make_float4(fmaxf(p.x, 0.f), fmaxf(p.y, 0.f), fmaxf(p.z, 0.f), fmaxf(p.w, 0.f));
      y0[i] = r;
      am[0] = f4absmax(am[0], r);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int slot = k_lo + j;
      if (slot < nslots) {
        float4 r = f4zero();
        if (a_has_slots) r = __ldg(a0 + slot * as4 + i);
        if (b_has_slots) r = f4add(r, __ldg(b0 + slot * bs4 + i));
        r = make_float4(p.x > 0.f ? r.x : 0.f, p.y > 0.f ? r.y : 0.f, p.z > 0.f ? r.z : 0.f, p.w > 0.f ? r.w : 0.f);
        y0[slot * ys4 + i] = r;
        am[1 + j] = f4absmax(am[1 + j], r);
      }
    }
  }
  if (amax) {
    if (grp == 0) block_absmax_commit(am[0], amax);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      __syncthreads();
      if (k_lo + j < nslots) block_absmax_commit(am[1 + j], amax + k_lo + j);
    }
  }
}
// adjoint: g = [y_0 > 0] gy_k ;  ga_k (+)= g ;  gb_k (+)= g   (either destination may be null)
// grid.y = groups of 8 cotangent slots: the primal output y0 (the ReLU mask) is read once per group
__global__ void __launch_bounds__(256, 4) add_relu_bwd_kernel(const float* __restrict__ gy, long long gy_slot,
                                                           const float* __restrict__ y0, float* __restrict__ ga,
                                                           long long ga_slot, int acc_a, float* __restrict__ gb,
                                                           long long gb_slot, int acc_b, long long n4, int slot0,
                                                           int nslots, unsigned int* __restrict__ amax_a,
                                                           unsigned int* __restrict__ amax_b) {
  // amax_a / amax_b (nullable, indexed by absolute slot): fused absmax of what is written to ga / gb
  float am[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) am[j] = 0.f;
  const int s_lo = slot0 + 8 * blockIdx.y;
  const int s_hi = min(slot0 + nslots, s_lo + 8);
  const long long gs4 = gy_slot >> 2, as4 = ga_slot >> 2, bs4 = gb_slot >> 2;
  const float4* g = reinterpret_cast<const float4*>(gy);
  const float4* p4 = reinterpret_cast<const float4*>(y0);
  float4* oa = reinterpret_cast<float4*>(ga);
  float4* ob = reinterpret_cast<float4*>(gb);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 p = __ldg(p4 + i);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int slot = s_lo + j;
      if (slot < s_hi) {
        const float4 v = __ldg(g + slot * gs4 + i);
        const float4 r = make_float4(p.x > 0.f ? v.x : 0.f, p.y > 0.f ? v.y : 0.f, p.z > 0.f ? v.z : 0.f,
                                     p.w > 0.f ? v.w : 0.f);
        if (oa) oa[slot * as4 + i] = acc_a ? f4add(r, oa[slot * as4 + i]) : r;
        if (ob) ob[slot * bs4 + i] = acc_b ? f4add(r, ob[slot * bs4 + i]) : r;
        am[j] = f4absmax(am[j], r);  // without accumulation both outputs equal r
      }
    }
  }
  if (amax_a || amax_b) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      __syncthreads();
      if (s_lo + j < s_hi) {
        if (amax_a) block_absmax_commit(am[j], amax_a + s_lo + j);
        if (amax_a && amax_b) __syncthreads();
        if (amax_b) block_absmax_commit(am[j], amax_b + s_lo + j);
      }
    }
  }
}

__global__ void zero_slots_kernel(float* __restrict__ dst, long long dst_slot, long long n4, int slot0) {
  float4* d = reinterpret_cast<float4*>(dst + (slot0 + blockIdx.y) * dst_slot);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x)
    d[i] = f4zero();
}

// max pooling: slot 0 computes max + argmax tap (uint8 per element); slot k gathers x_k[argmax].
// One thread per (output pixel, 4 channels), 32-bit index arithmetic (the host checks the element count).
__global__ void maxpool_fwd_kernel(const float* __restrict__ x, long long x_slot, float* __restrict__ y,
                                   long long y_slot, unsigned char* __restrict__ idx, int B, int Hs,
                                   int Ws, int Hd, int Wd, int Cp, int KH, int KW, int sh, int sw, int ph,
                                   int pw, int slot0, unsigned int* __restrict__ amax) {
  const int slot = slot0 + blockIdx.y;
  float am = 0.f;
  const float4* xs = reinterpret_cast<const float4*>(x + slot * x_slot);
  float4* ys = reinterpret_cast<float4*>(y + slot * y_slot);
  uchar4* ix = reinterpret_cast<uchar4*>(idx);
  const unsigned C4 = (unsigned)Cp >> 2;
  const unsigned total = (unsigned)B * Hd * Wd * C4;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned c4 = i % C4;
    unsigned pix = i / C4;
    const int wd = (int)(pix % (unsigned)Wd);
    pix /= (unsigned)Wd;
    const int hd = (int)(pix % (unsigned)Hd);
    const unsigned b = pix / (unsigned)Hd;
    const int h0 = hd * sh - ph, w0 = wd * sw - pw;
    if (slot == 0) {
      float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      uchar4 bi = make_uchar4(255, 255, 255, 255);
      for (int kh = 0; kh < KH; ++kh) {
        const int hs = h0 + kh;
        if (hs < 0 || hs >= Hs) continue;
        for (int kw = 0; kw < KW; ++kw) {
          const int ws = w0 + kw;
          if (ws < 0 || ws >= Ws) continue;
          const float4 v = __ldg(xs + ((size_t)(b * Hs + hs) * Ws + ws) * C4 + c4);
          const unsigned char tap = (unsigned char)(kh * KW + kw);
          if (v.x > best.x || bi.x == 255) { best.x = v.x; bi.x = tap; }
          if (v.y > best.y || bi.y == 255) { best.y = v.y; bi.y = tap; }
          if (v.z > best.z || bi.z == 255) { best.z = v.z; bi.z = tap; }
          if (v.w > best.w || bi.w == 255) { best.w = v.w; bi.w = tap; }
        }
      }
      ys[i] = best;
      ix[i] = bi;
      am = f4absmax(am, best);
    } else {
      const uchar4 bi = ix[i];
      auto pick = [&](unsigned char t, int lane) {
        const int kh = t / KW, kw = t - kh * KW;
        return __ldg(reinterpret_cast<const float*>(xs + ((size_t)(b * Hs + h0 + kh) * Ws + w0 + kw) * C4 + c4) + lane);
      };
      float4 v;
      if (bi.x == bi.y && bi.x == bi.z && bi.x == bi.w) {  // common case: one tap wins all four channels
        const int kh = bi.x / KW, kw = bi.x - kh * KW;
        v = __ldg(xs + ((size_t)(b * Hs + h0 + kh) * Ws + w0 + kw) * C4 + c4);
      } else {
        v = make_float4(pick(bi.x, 0), pick(bi.y, 1), pick(bi.z, 2), pick(bi.w, 3));
      }
      ys[i] = v;
      am = f4absmax(am, v);
    }
  }
  if (amax && !ph) block_absmax_commit(am, amax + slot);
}

// gx[pixel] (+)= sum over output windows whose argmax is this pixel of gy (gather form: deterministic).
// One thread per (input pixel, 4 channels) and up to 8 cotangent slots (grid.y = groups of 8 slots): the window
// arithmetic and the argmax bytes are shared by the slots; only the <= ceil(K/s)^2 windows that contain the
// pixel are visited.
__global__ void __launch_bounds__(256, 3) maxpool_bwd_kernel(const float* __restrict__ gy, long long gy_slot,
                                                          float* __restrict__ gx, long long gx_slot,
                                                          const unsigned char* __restrict__ idx, int B, int Hs,
                                                          int Ws, int Hd, int Wd, int Cp, int KH, int KW, int sh,
                                                          int sw, int ph, int pw, int slot0, int nslots,
                                                          int accumulate, unsigned int* __restrict__ amax) {
  float am[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) am[s] = 0.f;
  const int sfirst = blockIdx.y * 8;
  const int ns = min(8, nslots - sfirst);
  const float4* g = reinterpret_cast<const float4*>(gy + (long long)(slot0 + sfirst) * gy_slot);
  float4* o = reinterpret_cast<float4*>(gx + (long long)(slot0 + sfirst) * gx_slot);
  const long long g4 = gy_slot >> 2, o4 = gx_slot >> 2;
  const uchar4* ix = reinterpret_cast<const uchar4*>(idx);
  const unsigned C4 = (unsigned)Cp >> 2;
  const unsigned total = (unsigned)B * Hs * Ws * C4;  // 32-bit index arithmetic (checked on the host)
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned c4 = i % C4;
    unsigned pix = i / C4;
    const int ws = (int)(pix % (unsigned)Ws);
    pix /= (unsigned)Ws;
    const int hs = (int)(pix % (unsigned)Hs);
    const unsigned b = pix / (unsigned)Hs;
    // windows hd with hd*sh - ph <= hs <= hd*sh - ph + KH - 1
    int hd_lo = hs + ph - KH + 1; hd_lo = hd_lo <= 0 ? 0 : (hd_lo + sh - 1) / sh;
    int hd_hi = (hs + ph) / sh; if (hd_hi > Hd - 1) hd_hi = Hd - 1;
    int wd_lo = ws + pw - KW + 1; wd_lo = wd_lo <= 0 ? 0 : (wd_lo + sw - 1) / sw;
    int wd_hi = (ws + pw) / sw; if (wd_hi > Wd - 1) wd_hi = Wd - 1;
    float4 acc[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) acc[s] = f4zero();
    for (int hd = hd_lo; hd <= hd_hi; ++hd) {
      const int kh = hs + ph - hd * sh;
      for (int wd = wd_lo; wd <= wd_hi; ++wd) {
        const int tap = kh * KW + (ws + pw - wd * sw);
        const size_t oi = ((size_t)(b * Hd + hd) * Wd + wd) * C4 + c4;
        const uchar4 t = __ldg(ix + oi);
        if (t.x != tap && t.y != tap && t.z != tap && t.w != tap) continue;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          if (s < ns) {
            const float4 v = __ldg(g + s * g4 + oi);
            if (t.x == tap) acc[s].x += v.x;
            if (t.y == tap) acc[s].y += v.y;
            if (t.z == tap) acc[s].z += v.z;
            if (t.w == tap) acc[s].w += v.w;
          }
        }
      }
    }
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      if (s < ns) {
        float4 a = acc[s];
        if (accumulate) a = f4add(a, o[s * o4 + i]);
        o[s * o4 + i] = a;
        am[s] = f4absmax(am[s], a);
      }
    }
  }
  if (amax) {
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      __syncthreads();
      if (s < ns) block_absmax_commit(am[s], amax + slot0 + sfirst + s);
    }
  }
}

// global average pool: y[b][c] = mean_{hw} x[b][hw][c].  one thread per (b, c4).  grid.y = slot
__global__ void avgpool_fwd_kernel(const float* __restrict__ x, long long x_slot, float* __restrict__ y,
                                   long long y_slot, int B, int HW, int Cp, int slot0) {
  const int slot = slot0 + blockIdx.y;
  const int C4 = Cp >> 2;
  const float4* xs = reinterpret_cast<const float4*>(x + slot * x_slot);
  float4* ys = reinterpret_cast<float4*>(y + slot * y_slot);
  const float inv = 1.f / (float)HW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * C4; i += gridDim.x * blockDim.x) {
    int c4 = i % C4, b = i / C4;
    float4 acc = f4zero();
    for (int p = 0; p < HW; ++p) acc = f4add(acc, __ldg(xs + ((long long)b * HW + p) * C4 + c4));
    ys[i] = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
  }
}
__global__ void avgpool_bwd_kernel(const float* __restrict__ gy, long long gy_slot,
                                   float* __restrict__ gx, long long gx_slot, int B, int HW, int Cp,
                                   int slot0, int accumulate) {
  const int slot = slot0 + blockIdx.y;
  const int C4 = Cp >> 2;
  const float4* g = reinterpret_cast<const float4*>(gy + slot * gy_slot);
  float4* o = reinterpret_cast<float4*>(gx + slot * gx_slot);
  const float inv = 1.f / (float)HW;
  const long long total = (long long)B * HW * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c4 = (int)(i % C4);
    int b = (int)(i / ((long long)HW * C4));
    float4 v = __ldg(g + (long long)b * C4 + c4);
    v = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
    if (accumulate) v = f4add(v, o[i]);
    o[i] = v;
  }
}

// ------------------------------------------------------------------ adjoint of the affine map + reductions
// Deterministic column reduction helper: every thread owns a fixed channel group c4 = t % C4 and
// walks rows  rl, rl+RPP, ...  of its CTA's row range; the CTA then sums its row lanes in order.
// partial layout: [chunk][slot_idx][NV][Cp]
//
// affine_bwd:  gx_k (+)= s * gy_k  [+ sdot_k * gy_0 (R-op)];
//              partial v0 = sum gy_k * xhat_0 [+ gy_0 * xdot_k * invstd (R-op)]  (gamma grad)
//              partial v1 = sum gy_k                                             (beta grad)
// colsum (bias grad): same kernel with coef == nullptr: only v1 is produced, gx untouched.
__global__ void __launch_bounds__(256) affine_bwd_kernel(
    const float* __restrict__ gy, long long gy_slot, const float* __restrict__ x0,
    const float* __restrict__ xdot, long long xdot_slot, const float* __restrict__ coef,
    const float* __restrict__ aux, float* __restrict__ gx, long long gx_slot, int write_gx,
    float* __restrict__ partial, int want_partial, long long rows, int Cp, int rows_per_cta,
    int slot0, int nslots, int rop, int accumulate, int relu, unsigned int* __restrict__ amax,
    __half* __restrict__ ph, __half* __restrict__ pl, long long plane_slot,
    const unsigned int* __restrict__ in_bits, const unsigned int* __restrict__ smax_bits, int write_fp32) {
  // Planes mode (ph != null; half-split GEMM path): gx is ALSO (write_fp32) or ONLY written as the fp16 hi/lo
  // planes the producing convolution's wgrad / dgrad read, at slot (slot - slot0) of ph / pl.  The scale comes from
  // the bound  max|gx_k| <= max_c|s_c| * max|gy_k|  (in_bits: absmax words of gy by absolute slot, smax_bits: one
  // word); the bound's bit pattern is stored in amax[slot] in place of the exact maximum - any upper bound is a
  // valid scale, it only moves the 2^-39 absolute floor of the split.
  extern __shared__ float red[];  // [RPP][2][ctile*4]
  float am = 0.f;  // max |gx| written by this thread (amax: fused absmax for the half-split GEMMs)
  const int C4 = Cp >> 2;
  const int ctile4 = min(C4, 256);
  const int RPP = 256 / ctile4;
  const int t = threadIdx.x;
  const int rl = t / ctile4, cl = t - rl * ctile4;
  const bool active = rl < RPP;
  const int slot_idx = blockIdx.y;
  const int slot = slot0 + slot_idx;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  const float4* g = reinterpret_cast<const float4*>(gy + slot * gy_slot);
  const float4* g0 = reinterpret_cast<const float4*>(gy);
  const float4* xp = reinterpret_cast<const float4*>(x0);
  const float4* xd = (rop && xdot) ? reinterpret_cast<const float4*>(xdot + slot * xdot_slot) : nullptr;
  float4* o = (write_gx && write_fp32) ? reinterpret_cast<float4*>(gx + slot * gx_slot) : nullptr;
  float pscale = 0.f;
  uint2* oh = nullptr;
  uint2* ol = nullptr;
  if (ph) {
    const float bound = __uint_as_float(__ldg(smax_bits)) * __uint_as_float(__ldg(in_bits + slot));
    pscale = hs_pow2(hs_shift_from_bits(__float_as_uint(bound)));
    oh = reinterpret_cast<uint2*>(ph + (long long)slot_idx * plane_slot);
    ol = pl ? reinterpret_cast<uint2*>(pl + (long long)slot_idx * plane_slot) : nullptr;  // null: bf16 mode
    if (blockIdx.x == 0 && threadIdx.x == 0) amax[slot] = __float_as_uint(bound);
  }
  for (int cbase = 0; cbase < C4; cbase += ctile4) {
    const int c4 = cbase + cl;
    const bool cok = active && c4 < C4;
    float4 s = f4zero(), sd = f4zero(), istd = f4zero(), mu = f4zero(), tt = f4zero();
    if (cok && coef) {
      s = __ldg(reinterpret_cast<const float4*>(coef) + c4);
      tt = __ldg(reinterpret_cast<const float4*>(coef + Cp) + c4);
      istd = __ldg(reinterpret_cast<const float4*>(aux) + c4);
      mu = __ldg(reinterpret_cast<const float4*>(aux + Cp) + c4);
      if (rop && slot > 0) sd = __ldg(reinterpret_cast<const float4*>(coef + (long long)slot * 2 * Cp) + c4);
    }
    float4 a0 = f4zero(), a1 = f4zero();
    if (cok) {
      for (long long r = r0 + rl; r < r1; r += RPP) {
        long long i = r * C4 + c4;
        float4 v = __ldg(g + i);
        if (relu && coef) {  // fused ReLU adjoint: cotangent passes where the primal pre-activation is > 0
          const float4 pre = f4fma(s, __ldg(xp + i), tt);
          v = make_float4(pre.x > 0.f ? v.x : 0.f, pre.y > 0.f ? v.y : 0.f, pre.z > 0.f ? v.z : 0.f,
                          pre.w > 0.f ? v.w : 0.f);
        }
        a1 = f4add(a1, v);
        if (coef) {
          float4 xv = __ldg(xp + i);
          float4 xh = make_float4((xv.x - mu.x) * istd.x, (xv.y - mu.y) * istd.y, (xv.z - mu.z) * istd.z,
                                  (xv.w - mu.w) * istd.w);
          a0 = f4fma(v, xh, a0);
          float4 rr = f4mul(s, v);
          if (rop && slot > 0) {
            float4 q = __ldg(g0 + i);
            rr = f4fma(sd, q, rr);
            if (xd) a0 = f4fma(q, f4mul(__ldg(xd + i), istd), a0);
          }
          if (o) {
            if (accumulate) rr = f4add(rr, o[i]);
            o[i] = rr;
            am = f4absmax(am, rr);
          }
          if (oh) store_split_planes(reinterpret_cast<__half*>(oh), reinterpret_cast<__half*>(ol), i, rr, pscale);
        }
      }
    }
    if (want_partial) {
      // ordered in-CTA reduction over row lanes
      if (active) {
        float* r0p = red + ((rl * 2 + 0) * ctile4 + cl) * 4;
        float* r1p = red + ((rl * 2 + 1) * ctile4 + cl) * 4;
        r0p[0] = a0.x; r0p[1] = a0.y; r0p[2] = a0.z; r0p[3] = a0.w;
        r1p[0] = a1.x; r1p[1] = a1.y; r1p[2] = a1.z; r1p[3] = a1.w;
      }
      __syncthreads();
      if (rl == 0 && cok) {
        float s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
        for (int q = 0; q < RPP; ++q)
          for (int e = 0; e < 4; ++e) {
            s0[e] += red[((q * 2 + 0) * ctile4 + cl) * 4 + e];
            s1[e] += red[((q * 2 + 1) * ctile4 + cl) * 4 + e];
          }
        float* pp = partial + (((long long)blockIdx.x * nslots + slot_idx) * 2) * Cp;
        for (int e = 0; e < 4; ++e) { pp[c4 * 4 + e] = s0[e]; pp[Cp + c4 * 4 + e] = s1[e]; }
      }
      __syncthreads();
    }
  }
  if (amax) block_absmax_commit(am, amax + slot);
}

// ------------------------------------------------------------------ LayerNorm over the channels of a pixel / token
//   xhat = (x - mean) * rstd,  y_0 = gamma xhat + beta
//   y_k  = gamma * rstd (xdot_k - mean(xdot_k) - xhat mean(xhat xdot_k)) + gammadot_k xhat + betadot_k
// One warp per row (channel reductions by warp shuffles), Cp <= 1024: lane l holds channels 4 l + 128 j.
// coef [(1+K)][2][Cp]: slot 0 = (gamma, beta), slot k = (gammadot_k, betadot_k); aux [rows][2] = (mean, rstd).
constexpr int LN_MAXJ = 8;
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, long long x_slot, int x_has_slots,
                                                          const float* __restrict__ coef, int coef_has_tan,
                                                          float* __restrict__ y, long long y_slot, float* __restrict__ aux,
                                                          long long rows, int C, int Cp, float eps, int nslots) {
  const int lane = threadIdx.x & 31;
  const int nj = (Cp + 127) / 128;
  const float invC = 1.f / (float)C;
  for (long long r = blockIdx.x * 8LL + (threadIdx.x >> 5); r < rows; r += gridDim.x * 8LL) {
    float4 xh[LN_MAXJ];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAXJ; ++j) {
      const int c = lane * 4 + 128 * j;
      xh[j] = (j < nj && c < Cp) ? __ldg(reinterpret_cast<const float4*>(x + r * Cp + c)) : f4zero();
      s += (xh[j].x + xh[j].y) + (xh[j].z + xh[j].w);  // pad lanes hold zeros
    }
    const float mean = warp_sum(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAXJ; ++j) {
      const int c = lane * 4 + 128 * j;
      if (j < nj && c < Cp) {
        float4 d = make_float4(c < C ? xh[j].x - mean : 0.f, c + 1 < C ? xh[j].y - mean : 0.f,
                               c + 2 < C ? xh[j].z - mean : 0.f, c + 3 < C ? xh[j].w - mean : 0.f);
        xh[j] = d;
        q += (d.x * d.x + d.y * d.y) + (d.z * d.z + d.w * d.w);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * invC + eps);
    if (lane == 0) { aux[2 * r] = mean; aux[2 * r + 1] = rstd; }
#pragma unroll
    for (int j = 0; j < LN_MAXJ; ++j) {
      const int c = lane * 4 + 128 * j;
      if (j < nj && c < Cp) {
        xh[j] = make_float4(xh[j].x * rstd, xh[j].y * rstd, xh[j].z * rstd, xh[j].w * rstd);
        const float4 g = __ldg(reinterpret_cast<const float4*>(coef + c)), b = __ldg(reinterpret_cast<const float4*>(coef + Cp + c));
        *reinterpret_cast<float4*>(y + r * Cp + c) = f4fma(g, xh[j], b);
      }
    }
    for (int k = 1; k < nslots; ++k) {
      float4 xd[LN_MAXJ];
      float m1 = 0.f, m2 = 0.f;
      if (x_has_slots) {
#pragma unroll
        for (int j = 0; j < LN_MAXJ; ++j) {
          const int c = lane * 4 + 128 * j;
          xd[j] = (j < nj && c < Cp) ? __ldg(reinterpret_cast<const float4*>(x + k * x_slot + r * Cp + c)) : f4zero();
          m1 += (xd[j].x + xd[j].y) + (xd[j].z + xd[j].w);
          m2 += (xd[j].x * xh[j].x + xd[j].y * xh[j].y) + (xd[j].z * xh[j].z + xd[j].w * xh[j].w);
        }
        m1 = warp_sum(m1) * invC;
        m2 = warp_sum(m2) * invC;
      }
#pragma unroll
      for (int j = 0; j < LN_MAXJ; ++j) {
        const int c = lane * 4 + 128 * j;
        if (j < nj && c < Cp) {
          float4 o = f4zero();
          if (x_has_slots) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(coef + c));
            o = make_float4(c < C ? g.x * rstd * (xd[j].x - m1 - xh[j].x * m2) : 0.f,
                            c + 1 < C ? g.y * rstd * (xd[j].y - m1 - xh[j].y * m2) : 0.f,
                            c + 2 < C ? g.z * rstd * (xd[j].z - m1 - xh[j].z * m2) : 0.f,
                            c + 3 < C ? g.w * rstd * (xd[j].w - m1 - xh[j].w * m2) : 0.f);
          }
          if (coef_has_tan) {
            const float* ck = coef + (long long)k * 2 * Cp;
            o = f4add(o, f4fma(__ldg(reinterpret_cast<const float4*>(ck + c)), xh[j],
                               __ldg(reinterpret_cast<const float4*>(ck + Cp + c))));
          }
          *reinterpret_cast<float4*>(y + k * y_slot + r * Cp + c) = o;
        }
      }
    }
  }
}

// gx_k (+)= rstd (gh - mean(gh) - xhat mean(gh xhat)),  gh = gamma gy_k;  partial[chunk][k][0][c] = sum_rows gy_k xhat
// (gamma grad), partial[chunk][k][1][c] = sum_rows gy_k (beta grad).  grid = (chunks, slots), 8 warps = 8 rows at a time
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ gy, long long gy_slot,
                                                          const float* __restrict__ x0, const float* __restrict__ aux,
                                                          const float* __restrict__ coef, float* __restrict__ gx,
                                                          long long gx_slot, int write_gx, int accumulate,
                                                          float* __restrict__ partial, int want_partial, long long rows,
                                                          int C, int Cp, int rows_per_cta, int slot0, int nslots) {
  __shared__ float red[2 * 1024];  // [2][Cp]: the warps add their parameter partials one after the other (fixed order)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nj = (Cp + 127) / 128;
  const float invC = 1.f / (float)C;
  const int k = blockIdx.y, slot = slot0 + k;
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  float4 pg[LN_MAXJ], pb[LN_MAXJ];
#pragma unroll
  for (int j = 0; j < LN_MAXJ; ++j) { pg[j] = f4zero(); pb[j] = f4zero(); }
  for (long long r = r0 + warp; r < r1; r += 8) {
    const float mean = aux[2 * r], rstd = aux[2 * r + 1];
    float4 xh[LN_MAXJ], gh[LN_MAXJ];
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAXJ; ++j) {
      const int c = lane * 4 + 128 * j;
      if (j < nj && c < Cp) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x0 + r * Cp + c));
        const float4 g = __ldg(reinterpret_cast<const float4*>(gy + slot * gy_slot + r * Cp + c));
        xh[j] = make_float4(c < C ? (xv.x - mean) * rstd : 0.f, c + 1 < C ? (xv.y - mean) * rstd : 0.f,
                            c + 2 < C ? (xv.z - mean) * rstd : 0.f, c + 3 < C ? (xv.w - mean) * rstd : 0.f);
        pg[j] = f4fma(g, xh[j], pg[j]);
        pb[j] = f4add(pb[j], g);
        gh[j] = f4mul(__ldg(reinterpret_cast<const float4*>(coef + c)), g);
        m1 += (gh[j].x + gh[j].y) + (gh[j].z + gh[j].w);
        m2 += (gh[j].x * xh[j].x + gh[j].y * xh[j].y) + (gh[j].z * xh[j].z + gh[j].w * xh[j].w);
      } else { xh[j] = f4zero(); gh[j] = f4zero(); }
    }
    m1 = warp_sum(m1) * invC;
    m2 = warp_sum(m2) * invC;
    if (write_gx) {
#pragma unroll
      for (int j = 0; j < LN_MAXJ; ++j) {
        const int c = lane * 4 + 128 * j;
        if (j < nj && c < Cp) {
          float4 o = make_float4(c < C ? rstd * (gh[j].x - m1 - xh[j].x * m2) : 0.f,
                                 c + 1 < C ? rstd * (gh[j].y - m1 - xh[j].y * m2) : 0.f,
                                 c + 2 < C ? rstd * (gh[j].z - m1 - xh[j].z * m2) : 0.f,
                                 c + 3 < C ? rstd * (gh[j].w - m1 - xh[j].w * m2) : 0.f);
          float4* dst = reinterpret_cast<float4*>(gx + slot * gx_slot + r * Cp + c);
          if (accumulate) o = f4add(o, *dst);
          *dst = o;
        }
      }
    }
  }
  if (!want_partial) return;
  for (int w = 0; w < 8; ++w) {
    if (warp == w) {
#pragma unroll
      for (int j = 0; j < LN_MAXJ; ++j) {
        const int c = lane * 4 + 128 * j;
        if (j < nj && c < Cp) {
          float4* a = reinterpret_cast<float4*>(red + c);
          float4* b = reinterpret_cast<float4*>(red + 1024 + c);
          *a = w == 0 ? pg[j] : f4add(*a, pg[j]);
          *b = w == 0 ? pb[j] : f4add(*b, pb[j]);
        }
      }
    }
    __syncthreads();
  }
  float* dst = partial + ((long long)blockIdx.x * nslots + k) * 2 * Cp;  // [chunk][k][which][c]
  for (int e = threadIdx.x; e < 2 * Cp; e += blockDim.x) {
    const int which = e / Cp, c = e - which * Cp;
    dst[e] = red[which * 1024 + c];
  }
}

// out[(off + c)*ldk + k0 + k] += alpha * sum_chunks partial[chunk][k][which][c]
// (slot indices kskip .. nslots-1 of the partial buffer map to columns 0 .. nslots-kskip-1)
// One WARP per output: the lanes stride over the chunks (a few hundred dependent loads per output were pure latency
// with one thread per output: 41 launches x 30 us per C2 step), then a shuffle tree in a fixed order (deterministic).
__global__ void __launch_bounds__(256) vec_grad_finish_kernel(const float* __restrict__ partial, int nchunks, int nslots,
                                                            int kskip, int which, int C, int Cp,
                                                            float* __restrict__ out, long long off, int ldk, int k0,
                                                            float alpha) {
  const int nk = nslots - kskip;
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= C * nk) return;
  const int k = i % nk, c = i / nk;
  const float* q = partial + ((long long)(kskip + k) * 2 + which) * Cp + c;
  const long long stride = (long long)nslots * 2 * Cp;
  float a = 0.f;
  for (int ch = lane; ch < nchunks; ch += 32) a += __ldg(q + (long long)ch * stride);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) out[(off + c) * ldk + k0 + k] += alpha * a;
}

// First level of the split-K sum when there are many splits of a small weight tensor (layer1: 196 splits of 37 K
// weights - one thread per weight would walk 196 x 8 strided loads with a handful of warps per SM):
// partial[ch * L][j] = sum_{sp in [ch * L, min(nsplit, (ch + 1) * L))} partial[sp][j], j < split_elems, fixed order;
// grid (x, chunks).  The finish kernel then sums the chunk heads (split_step = L).
__global__ void __launch_bounds__(256) wgrad_presum_kernel(float* __restrict__ partial, int nsplit, int L,
                                                          long long split_elems) {
  const int s0 = blockIdx.y * L, s1 = min(nsplit, s0 + L);
  float4* base = reinterpret_cast<float4*>(partial + (long long)s0 * split_elems);
  const long long n4 = split_elems >> 2;  // split_elems is a multiple of 4 (Cp % 8 == 0)
  for (long long j = blockIdx.x * 256LL + threadIdx.x; j < n4; j += (long long)gridDim.x * 256) {
    float4 a = base[j];
    int sp = s0 + 1;
    for (; sp + 2 <= s1; sp += 2) {
      const float4 b = __ldcs(reinterpret_cast<const float4*>(partial + (long long)sp * split_elems) + j);
      const float4 c = __ldcs(reinterpret_cast<const float4*>(partial + (long long)(sp + 1) * split_elems) + j);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
    }
    for (; sp < s1; ++sp) {
      const float4 b = __ldcs(reinterpret_cast<const float4*>(partial + (long long)sp * split_elems) + j);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    base[j] = a;
  }
}

// out[(off + ((n*C + c)*taps + tap))*ldk + k0 + k] += alpha * sum_splits partial[split][k][n][tap][cp]
// One thread per weight element (output row): it sums the splits of every column in a fixed order (reads coalesced
// over the warp for each (split, column)) and updates the nk <= 8 adjacent floats of its row with one 16- or 32-byte
// read-modify-write when the row is aligned.  (Round 2 tried a CTA tile transposed through shared memory, 6.2 ms per
// C2 step, and a 32 x nk register transpose, 2.7 ms, against 2.3 ms for the column-per-warp kernel this replaces,
// whose 4-byte updates used one eighth of every 32-byte sector.)
// split_step: distance between the splits to sum, in splits (> 1 after wgrad_presum_kernel)
__global__ void __launch_bounds__(256) wgrad_finish_kernel(const float* __restrict__ partial, int nsplit, int nslots,
                                                          int kskip, int N, int C, int Cp, int taps,
                                                          float* __restrict__ out, long long off, int ldk, int k0,
                                                          float alpha, int nrows, int split_step) {
  const long long per = (long long)N * taps * Cp;
  const long long per_slot = (long long)(nrows > 0 ? nrows : N) * taps * Cp;
  const long long stride = (long long)nslots * per_slot * split_step;
  const int nk = nslots - kskip;
  const bool vec = (ldk % 4 == 0) && (k0 % 4 == 0) && (nk % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < per; i += (long long)gridDim.x * 256) {
    const int c = (int)(i % Cp);
    if (c >= C) continue;
    const long long r = i / Cp;
    const int tap = (int)(r % taps);
    const int n = (int)(r / taps);
    const float* q = partial + (long long)kskip * per_slot + i;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) {
      const float* qs = q + (long long)sp * stride;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < nk) acc[k] += __ldg(qs + (long long)k * per_slot);
    }
    float* o = out + (off + ((long long)n * C + c) * taps + tap) * ldk + k0;
    if (vec) {
#pragma unroll
      for (int k = 0; k < 8; k += 4) {
        if (k < nk) {
          float4 v = *reinterpret_cast<float4*>(o + k);
          v.x += alpha * acc[k]; v.y += alpha * acc[k + 1]; v.z += alpha * acc[k + 2]; v.w += alpha * acc[k + 3];
          *reinterpret_cast<float4*>(o + k) = v;
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < nk) o[k] += alpha * acc[k];
    }
  }
}

// ------------------------------------------------------------------ loss Hessian apply
// One CTA (128 threads) per sample.  f = logits slot 0 [B][Cp]; u_k = slot k (Jv); result written to
// g (+ slot stride) slot k in place of the tangent:   hu_k = scale * (p*u - p*(p.u))      (CE)
//                                                     hu_k = 2*scale*u                     (MSE)
//                                                     hu_k = scale*s(1-s)*u                (BCE)
//                                                     hu_k = scale * sum_m g_m (g_m . u)   (MC)
// Hessian mode additionally writes slot 0 of g: dl/df = scale*(p - onehot(y)) etc.
__device__ __forceinline__ float block_reduce_sum128(float v, float* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0] + sh[1] + sh[2] + sh[3];
  return r;
}
__device__ __forceinline__ float block_reduce_max128(float v, float* sh) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  return fmaxf(fmaxf(sh[0], sh[1]), fmaxf(sh[2], sh[3]));
}

__global__ void __launch_bounds__(128) loss_hessian_kernel(
    int loss, int mc_samples, const float* __restrict__ f, const float* __restrict__ u,
    long long u_slot, float* __restrict__ g, long long g_slot, const void* __restrict__ y,
    const float* __restrict__ mc_grad, int C, int Cp, int K, float scale, int write_grad0) {
  __shared__ float sh[4];
  extern __shared__ float prob[];  // [Cp]
  const int b = blockIdx.x, t = threadIdx.x;
  const float* fb = f + (long long)b * Cp;
  if (loss == 0 /*CE*/) {
    float mx = -INFINITY;
    for (int c = t; c < C; c += 128) mx = fmaxf(mx, fb[c]);
    mx = block_reduce_max128(mx, sh);
    float sum = 0.f;
    for (int c = t; c < C; c += 128) { float e = expf(fb[c] - mx); prob[c] = e; sum += e; }
    sum = block_reduce_sum128(sum, sh);
    float inv = 1.f / sum;
    for (int c = t; c < C; c += 128) prob[c] *= inv;
  } else if (loss == 2 /*BCE*/) {
    for (int c = t; c < C; c += 128) prob[c] = 1.f / (1.f + expf(-fb[c]));
  }
  __syncthreads();
  if (write_grad0) {
    float* g0 = g + (long long)b * Cp;
    if (loss == 0) {
      long long lab = reinterpret_cast<const long long*>(y)[b];
      for (int c = t; c < Cp; c += 128) g0[c] = c < C ? scale * (prob[c] - (c == lab ? 1.f : 0.f)) : 0.f;
    } else if (loss == 1) {
      const float* yb = reinterpret_cast<const float*>(y) + (long long)b * C;
      for (int c = t; c < Cp; c += 128) g0[c] = c < C ? 2.f * scale * (fb[c] - yb[c]) : 0.f;
    } else {
      const float* yb = reinterpret_cast<const float*>(y) + (long long)b * C;
      for (int c = t; c < Cp; c += 128) g0[c] = c < C ? scale * (prob[c] - yb[c]) : 0.f;
    }
  }
  for (int k = 1; k <= K; ++k) {
    const float* uk = u + k * u_slot + (long long)b * Cp;
    float* gk = g + k * g_slot + (long long)b * Cp;
    if (mc_samples > 0) {
      // rank-M estimate; accumulate into registers per channel owned by this thread
      // (C <= 128*8 channels per thread loop handled generically through a second pass)
      // NOTE: u_k and g_k alias in GGN mode, so all dots are computed before any write.
      // dots for all m are computed first and kept in shared memory (M <= 32).
      __shared__ float dots[32];
      for (int m = 0; m < mc_samples; ++m) {
        const float* gm = mc_grad + ((long long)b * mc_samples + m) * C;
        float d = 0.f;
        for (int c = t; c < C; c += 128) d += gm[c] * uk[c];
        d = block_reduce_sum128(d, sh);
        if (t == 0) dots[m] = d;
      }
      __syncthreads();
      for (int c = t; c < Cp; c += 128) {
        float r = 0.f;
        if (c < C)
          for (int m = 0; m < mc_samples; ++m) r += mc_grad[((long long)b * mc_samples + m) * C + c] * dots[m];
        gk[c] = scale * r;
      }
    } else if (loss == 0) {
      float d = 0.f;
      for (int c = t; c < C; c += 128) d += prob[c] * uk[c];
      d = block_reduce_sum128(d, sh);
      for (int c = t; c < Cp; c += 128) gk[c] = c < C ? scale * prob[c] * (uk[c] - d) : 0.f;
    } else if (loss == 1) {
      for (int c = t; c < Cp; c += 128) gk[c] = c < C ? 2.f * scale * uk[c] : 0.f;
    } else {
      for (int c = t; c < Cp; c += 128) gk[c] = c < C ? scale * prob[c] * (1.f - prob[c]) * uk[c] : 0.f;
    }
    __syncthreads();
  }
}

}  // namespace curv
