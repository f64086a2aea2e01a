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

namespace curv {

struct Bgemm3 {
  int transA, transB, M, N, Kd;
  float alpha, beta;
  const float* A; int lda; long long sA[3];
  const float* B; int ldb; long long sB[3];
  float* C; int ldc; long long sC[3];
  int n1, n2;  // blockIdx.z = (i0 * n1 + i1) * n2 + i2
};

// op(A) [M, Kd] . op(B) [Kd, N]; transA: A stored [Kd, M]; transB: B stored [N, Kd].  64 x 64 tiles, scalar loads.
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
    __syncthreads();
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

// in-place softmax of `rows` rows of length T (one warp per row)
__global__ void __launch_bounds__(256) attn_softmax_kernel(float* __restrict__ S, long long rows, int T) {
  const long long row = blockIdx.x * 8LL + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float* s = S + row * T;
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
                                                              long long rows_per_slot, long long rows, int T) {
  const long long row = blockIdx.x * 8LL + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* p = P + (row % rows_per_slot) * T;
  float* d = D + row * T;
  float dot = 0.f;
  for (int j = lane; j < T; j += 32) dot = fmaf(p[j], d[j], dot);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  for (int j = lane; j < T; j += 32) d[j] = p[j] * (d[j] - dot);
}

}  // namespace curv
