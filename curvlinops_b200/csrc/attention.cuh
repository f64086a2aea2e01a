// Multi-head self-attention core (CURV_OP_ATTENTION): in = packed projections [B, T, 3E] (q | k | v, head h at columns
// h*d .. h*d + d of each third), out = [B, T, E]; per example and head
//     P = softmax(Q K^T / sqrt(d)),  O = P V
// (torch.nn.functional.scaled_dot_product_attention without mask / dropout, the op nn.MultiheadAttention lowers to;
// the reference differentiates it with torch.func, ggn.py:61-71).  Tangent and adjoint:
//     dS = (dQ K^T + Q dK^T) / sqrt(d),  dP = P o (dS - rowsum(P o dS)),  dO = dP V + P dV
//     gV = P^T gO,  gP = gO V^T,  gS = P o (gP - rowsum(P o gP)),  gQ = gS K / sqrt(d),  gK = gS^T Q / sqrt(d)
// P is kept from the forward sweep (B * heads * T * T floats per attention node).  First version: exact fp32 FMA GEMMs
// batched over (slot, example, head) with three-level strides, warp-per-row softmax kernels.  The core is ~4 % of a
// ViT-B/16's forward FLOPs (T = 197, d = 64); moving it onto the tensor cores is listed in DESIGN.md.
#pragma once
#include <cuda_bf16.h>

namespace curv {

struct Bgemm3 {
  int transA, transB, M, N, Kd;
  float alpha, beta;
  const float* A; int lda; long long sA[3];
  const float* B; int ldb; long long sB[3];
  float* C; int ldc; long long sC[3];
  int n1, n2;  // blockIdx.z = (i0 * n1 + i1) * n2 + i2
  // optional second product added to the first (C = alpha (op(A) op(B) + op(A2) op(B2)) + beta C): the two terms of a
  // product rule (dQ K^T + Q dK^T, dP V + P dV) in ONE pass over C.  A2 == nullptr: absent.  Same M, N, own Kd2.
  int transA2, transB2, Kd2;
  const float* A2; int lda2; long long sA2[3];
  const float* B2; int ldb2; long long sB2[3];
};

__device__ __forceinline__ uint32_t attn_pack_bf16(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);  // .x (low half) = the element with the smaller k
  return *reinterpret_cast<const uint32_t*>(&v);
}

// op(A) [M, Kd] . op(B) [Kd, N]; transA: A stored [Kd, M]; transB: B stored [N, Kd].  64 x 64 tiles, scalar loads.
// MMA = false: exact fp32 FMAs (fp32 operators).  MMA = true (bf16 operators): the staged fp32 tile is rounded to bf16
// while the fragments are built and multiplied with mma.sync.m16n8k16 (fp32 accumulation) -- 8 warps x (16 x 32) of the
// tile; the operands are activations on both sides, 197 x 64 per head, too small and too many for the weight-image
// tcgen05 kernels of hs_gemm.cuh.
template <bool MMA>
__global__ void __launch_bounds__(256) attn_bgemm_kernel(const Bgemm3 p) {
  constexpr int BM = 64, BN = 64, BK = 16, LDA = BM + 4, LDB = BN + 4;
  __shared__ __align__(16) float As[BK][LDA];
  __shared__ __align__(16) float Bs[BK][LDB];
  const int z = blockIdx.z;
  const int i2 = z % p.n2, i1 = (z / p.n2) % p.n1, i0 = z / (p.n2 * p.n1);
  const float* A = p.A + i0 * p.sA[0] + i1 * p.sA[1] + i2 * p.sA[2];
  const float* B = p.B + i0 * p.sB[0] + i1 * p.sB[1] + i2 * p.sB[2];
  float* C = p.C + i0 * p.sC[0] + i1 * p.sC[1] + i2 * p.sC[2];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < p.Kd; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = t + i * 256;
      int r, m;
      if (p.transA) { m = e & 63; r = e >> 6; } else { r = e & 15; m = e >> 4; }
      const int gm = m0 + m, gr = k0 + r;
      float v = 0.f;
      if (gm < p.M && gr < p.Kd)
        v = p.transA ? __ldg(A + (long long)gr * p.lda + gm) : __ldg(A + (long long)gm * p.lda + gr);
      As[r][m] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = t + i * 256;
      int r, n;
      if (p.transB) { r = e & 15; n = e >> 4; } else { n = e & 63; r = e >> 6; }
      const int gn = n0 + n, gr = k0 + r;
      float v = 0.f;
      if (gn < p.N && gr < p.Kd)
        v = p.transB ? __ldg(B + (long long)gn * p.ldb + gr) : __ldg(B + (long long)gr * p.ldb + gn);
      Bs[r][n] = v;
    }
    __syncthreads();
    if (MMA) {
      const int w = t >> 5, lane = t & 31, g = lane >> 2, q = lane & 3;
      const int mw = (w >> 1) * 16, nw = (w & 1) * 32;
      const uint32_t a0 = attn_pack_bf16(As[2 * q][mw + g], As[2 * q + 1][mw + g]);
      const uint32_t a1 = attn_pack_bf16(As[2 * q][mw + g + 8], As[2 * q + 1][mw + g + 8]);
      const uint32_t a2 = attn_pack_bf16(As[2 * q + 8][mw + g], As[2 * q + 9][mw + g]);
      const uint32_t a3 = attn_pack_bf16(As[2 * q + 8][mw + g + 8], As[2 * q + 9][mw + g + 8]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = nw + 8 * j + g;
        const uint32_t b0 = attn_pack_bf16(Bs[2 * q][n], Bs[2 * q + 1][n]);
        const uint32_t b1 = attn_pack_bf16(Bs[2 * q + 8][n], Bs[2 * q + 9][n]);
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(acc[j][0]), "+f"(acc[j][1]), "+f"(acc[j][2]), "+f"(acc[j][3])
            : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      }
    } else {
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
  if (MMA) {  // accumulator j: rows mw + g (+8), columns nw + 8 j + 2 q (+1)
    const int w = t >> 5, lane = t & 31, g = lane >> 2, q = lane & 3;
    const int mw = (w >> 1) * 16, nw = (w & 1) * 32;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int m = m0 + mw + g + (e >> 1) * 8, n = n0 + nw + 8 * j + 2 * q + (e & 1);
        if (m >= p.M || n >= p.N) continue;
        float* c = C + (long long)m * p.ldc + n;
        float v = p.alpha * acc[j][e];
        if (p.beta != 0.f) v += p.beta * *c;
        *c = v;
      }
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float* c = C + (long long)m * p.ldc + n;
      float v = p.alpha * acc[i][j];
      if (p.beta != 0.f) v += p.beta * *c;
      *c = v;
    }
  }
}

// bf16 operators, whole problem per CTA: one (slot, example, head) product with M, N, Kd <= 224.  Both operands are
// read ONCE from global memory, rounded to bf16 and kept in shared memory with the reduction index contiguous
// (As[m][k], Bs[n][k], row pitch Kp + 8 halves: the m16n8k16 fragment loads are conflict-free); 8 warps sweep the
// 16 x 32 output strips with mma.sync (fp32 accumulation).  Against the 64 x 64-tile kernel above, which re-fetches
// every operand once per output tile and waits for each 16-deep slice: 86 -> see DESIGN.md section 3.8.
constexpr int ATTN_SG_THREADS = 1024;
// dst[r][k] (bf16, row pitch `pitch`, zero-padded to Rp x Kp) <- src;  K_CONTIG: src[r * ld + k], else src[k * ld + r]
template <bool K_CONTIG>
__device__ __forceinline__ void attn_stage(__nv_bfloat16* __restrict__ dst, int pitch, const float* __restrict__ src,
                                           int ld, int R, int Rp, int Kd, int Kp, int t) {
  const bool vec = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (ld % 4 == 0);
  if (K_CONTIG) {
    const int groups = Kp >> 2, total = Rp * groups;
#pragma unroll 4
    for (int e = t; e < total; e += ATTN_SG_THREADS) {
      const int r = e / groups, k = (e - r * groups) << 2;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < R && k < Kd) {
        const float* q = src + (long long)r * ld + k;
        if (vec && k + 3 < Kd) v = __ldg(reinterpret_cast<const float4*>(q));
        else {
          v.x = __ldg(q);
          if (k + 1 < Kd) v.y = __ldg(q + 1);
          if (k + 2 < Kd) v.z = __ldg(q + 2);
          if (k + 3 < Kd) v.w = __ldg(q + 3);
        }
      }
      __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(dst + r * pitch + k);
      o[0] = __floats2bfloat162_rn(v.x, v.y);
      o[1] = __floats2bfloat162_rn(v.z, v.w);
    }
  } else {
    const int groups = Rp >> 2, total = Kp * groups;
#pragma unroll 4
    for (int e = t; e < total; e += ATTN_SG_THREADS) {
      const int k = e / groups, r = (e - k * groups) << 2;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < Kd && r < R) {
        const float* q = src + (long long)k * ld + r;
        if (vec && r + 3 < R) v = __ldg(reinterpret_cast<const float4*>(q));
        else {
          v.x = __ldg(q);
          if (r + 1 < R) v.y = __ldg(q + 1);
          if (r + 2 < R) v.z = __ldg(q + 2);
          if (r + 3 < R) v.w = __ldg(q + 3);
        }
      }
      dst[r * pitch + k] = __float2bfloat16_rn(v.x);
      dst[(r + 1) * pitch + k] = __float2bfloat16_rn(v.y);
      dst[(r + 2) * pitch + k] = __float2bfloat16_rn(v.z);
      dst[(r + 3) * pitch + k] = __float2bfloat16_rn(v.w);
    }
  }
}
  // 32 warps: the staging loads of a lone resident CTA need the parallelism
__global__ void __launch_bounds__(ATTN_SG_THREADS) attn_sgemm_kernel(const Bgemm3 p) {
  extern __shared__ __align__(16) unsigned char attn_smem[];
  const int Mp = (p.M + 15) & ~15, Np = (p.N + 31) & ~31, Kp1 = (p.Kd + 15) & ~15;
  const int Kp = Kp1 + (p.A2 ? ((p.Kd2 + 15) & ~15) : 0), pitch = Kp + 8;  // the second product extends the reduction
  __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(attn_smem);
  __nv_bfloat16* Bs = As + (size_t)Mp * pitch;
  const int z = blockIdx.x;
  const int i2 = z % p.n2, i1 = (z / p.n2) % p.n1, i0 = z / (p.n2 * p.n1);
  const float* A = p.A + i0 * p.sA[0] + i1 * p.sA[1] + i2 * p.sA[2];
  const float* B = p.B + i0 * p.sB[0] + i1 * p.sB[1] + i2 * p.sB[2];
  float* C = p.C + i0 * p.sC[0] + i1 * p.sC[1] + i2 * p.sC[2];
  const int t = threadIdx.x;
  // ---- stage both operands: 16-byte loads along the contiguous axis of the source where it is aligned
  if (p.transA) attn_stage<false>(As, pitch, A, p.lda, p.M, Mp, p.Kd, Kp1, t);
  else attn_stage<true>(As, pitch, A, p.lda, p.M, Mp, p.Kd, Kp1, t);
  if (p.transB) attn_stage<true>(Bs, pitch, B, p.ldb, p.N, Np, p.Kd, Kp1, t);
  else attn_stage<false>(Bs, pitch, B, p.ldb, p.N, Np, p.Kd, Kp1, t);
  if (p.A2) {
    const float* A2 = p.A2 + i0 * p.sA2[0] + i1 * p.sA2[1] + i2 * p.sA2[2];
    const float* B2 = p.B2 + i0 * p.sB2[0] + i1 * p.sB2[1] + i2 * p.sB2[2];
    const int Kp2 = Kp - Kp1;
    if (p.transA2) attn_stage<false>(As + Kp1, pitch, A2, p.lda2, p.M, Mp, p.Kd2, Kp2, t);
    else attn_stage<true>(As + Kp1, pitch, A2, p.lda2, p.M, Mp, p.Kd2, Kp2, t);
    if (p.transB2) attn_stage<true>(Bs + Kp1, pitch, B2, p.ldb2, p.N, Np, p.Kd2, Kp2, t);
    else attn_stage<false>(Bs + Kp1, pitch, B2, p.ldb2, p.N, Np, p.Kd2, Kp2, t);
  }
  __syncthreads();
  const int w = t >> 5, lane = t & 31, g = lane >> 2, q = lane & 3;
  const int mtiles = Mp / 16, nstrips = Np / 32;
  for (int tile = w; tile < mtiles * nstrips; tile += ATTN_SG_THREADS / 32) {
    const int m0 = (tile % mtiles) * 16, n0 = (tile / mtiles) * 32;
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
    const __nv_bfloat16* a_lo = As + (m0 + g) * pitch + 2 * q;
    const __nv_bfloat16* a_hi = a_lo + 8 * pitch;
    for (int k0 = 0; k0 < Kp; k0 += 16) {
      const uint32_t a0 = *reinterpret_cast<const uint32_t*>(a_lo + k0);
      const uint32_t a1 = *reinterpret_cast<const uint32_t*>(a_hi + k0);
      const uint32_t a2 = *reinterpret_cast<const uint32_t*>(a_lo + k0 + 8);
      const uint32_t a3 = *reinterpret_cast<const uint32_t*>(a_hi + k0 + 8);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat16* bp = Bs + (n0 + 8 * j + g) * pitch + 2 * q + k0;
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(bp);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(bp + 8);
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(acc[j][0]), "+f"(acc[j][1]), "+f"(acc[j][2]), "+f"(acc[j][3])
            : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int m = m0 + g + (e >> 1) * 8, n = n0 + 8 * j + 2 * q + (e & 1);
        if (m >= p.M || n >= p.N) continue;
        float* c = C + (long long)m * p.ldc + n;
        float v = p.alpha * acc[j][e];
        if (p.beta != 0.f) v += p.beta * *c;
        *c = v;
      }
  }
}
static inline size_t attn_sgemm_smem(int M, int N, int Kd, int Kd2 = 0) {
  const int Mp = (M + 15) & ~15, Np = (N + 31) & ~31, Kp = ((Kd + 15) & ~15) + (Kd2 > 0 ? ((Kd2 + 15) & ~15) : 0);
  return (size_t)(Mp + Np) * (Kp + 8) * 2;
}

// in-place softmax of `rows` rows of length T, row pitch Tp (one warp per row)
__global__ void __launch_bounds__(256) attn_softmax_kernel(float* __restrict__ S, long long rows, int T, int Tp) {
  const long long row = blockIdx.x * 8LL + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float* s = S + row * Tp;
  float mx = -INFINITY;
  for (int j = lane; j < T; j += 32) mx = fmaxf(mx, s[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int j = lane; j < T; j += 32) {
    const float e = expf(s[j] - mx);
    s[j] = e;
    sum += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  for (int j = lane; j < T; j += 32) s[j] *= inv;
}

// D <- P o (D - rowsum(P o D)) in place; D has nslots * rows_per_slot rows, P rows_per_slot rows (shared by the slots)
__global__ void __launch_bounds__(256) attn_softmax_jvp_kernel(const float* __restrict__ P, float* __restrict__ D,
                                                              long long rows_per_slot, long long rows, int T,
                                                              int Tp) {
  const long long row = blockIdx.x * 8LL + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* p = P + (row % rows_per_slot) * Tp;
  float* d = D + row * Tp;
  float dot = 0.f;
  for (int j = lane; j < T; j += 32) dot = fmaf(p[j], d[j], dot);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  for (int j = lane; j < T; j += 32) d[j] = p[j] * (d[j] - dot);
}

}  // namespace curv
