// fp32 SIMT implicit-GEMM kernels (exact IEEE fp32 FMA path).
//
// These are the always-available contraction kernels: they serve the small / oddly shaped layers
// (MLP test nets, 1x1 maps, fc) and are the numerical cross-check of the tcgen05 path (tc_gemm.cuh).
// Tiling: BM x BN output tile per 256-thread CTA, 16-deep reduction chunks, register-prefetch double
// buffering through shared memory, (BM/16) x (BN/16) outputs per thread.
#pragma once
#include "common.cuh"

namespace curv {

__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

// Inner product stage shared by both kernels: acc += As[kk][rows] (x) Bs[kk][cols].
template <int BM, int BN, int LDA, int LDB>
__device__ __forceinline__ void simt_tile_fma(const float (*As)[LDA], const float (*Bs)[LDB], int tx,
                                              int ty, float (&acc)[BM / 16][BN / 16]) {
  constexpr int TM = BM / 16, TN = BN / 16;
#pragma unroll
  for (int kk = 0; kk < 16; ++kk) {
    float a[TM], b[TN];
#pragma unroll
    for (int i = 0; i < TM / 4; ++i) {
      float4 v = *reinterpret_cast<const float4*>(&As[kk][i * (BM / 2) + ty * 4]);
      a[i * 4 + 0] = v.x; a[i * 4 + 1] = v.y; a[i * 4 + 2] = v.z; a[i * 4 + 3] = v.w;
    }
#pragma unroll
    for (int j = 0; j < TN / 4; ++j) {
      float4 v = *reinterpret_cast<const float4*>(&Bs[kk][j * (BN / 2) + tx * 4]);
      b[j * 4 + 0] = v.x; b[j * 4 + 1] = v.y; b[j * 4 + 2] = v.z; b[j * 4 + 3] = v.w;
    }
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

// ---------------------------------------------------------------------------------------------
// gather GEMM:  out[m][n] = sum_r gather(A)[m][r] * W[n][r]     (forward conv, dgrad, Linear)
// grid.x = m-tiles * n-tiles (n fastest), grid.y = slots
// ---------------------------------------------------------------------------------------------
template <int BM, int BN>
__global__ void __launch_bounds__(256) gather_gemm_simt(const GatherGemmArgs p) {
  constexpr int BK = 16;
  constexpr int TM = BM / 16, TN = BN / 16;
  constexpr int LDA = BM + 4, LDB = BN + 4;
  constexpr int AR = BM / 64, BR = BN / 64;
  __shared__ __align__(16) float As[2][BK][LDA];
  __shared__ __align__(16) float Bs[2][BK][LDB];

  const Geom& g = p.g;
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int tiles_n = ceil_div(g.Nd, BN);
  const int tile_m = blockIdx.x / tiles_n, tile_n = blockIdx.x - tile_m * tiles_n;
  const int slot = p.slot0 + blockIdx.y;
  const int m0 = tile_m * BM, n0 = tile_n * BN;

  // segments of this slot
  const float* segA[2];
  const float* segB[2];
  int nseg = 0;
  if (slot == 0) {
    segA[0] = p.A; segB[0] = p.W; nseg = 1;
  } else {
    if (p.a_has_slots) { segA[nseg] = p.A + (long long)slot * p.A_slot; segB[nseg] = p.W; ++nseg; }
    if (p.Wt != nullptr) { segA[nseg] = p.A; segB[nseg] = p.Wt + (long long)(slot - 1) * p.Wt_slot; ++nseg; }
  }

  // per-thread loader rows
  const int lr = t >> 2;   // 0..63
  const int l4 = t & 3;    // which float4 of the 16-wide chunk
  int a_h[AR], a_w[AR];
  long long a_base[AR];
  bool a_ok[AR];
#pragma unroll
  for (int i = 0; i < AR; ++i) {
    int m = m0 + lr + i * 64;
    a_ok[i] = m < g.M;
    int mm = a_ok[i] ? m : 0;
    int b = mm / (g.Hd * g.Wd);
    int rem = mm - b * (g.Hd * g.Wd);
    int hd = rem / g.Wd, wd = rem - hd * g.Wd;
    if (g.mode == 0) { a_h[i] = hd * g.sh - g.ph; a_w[i] = wd * g.sw - g.pw; }
    else             { a_h[i] = hd + g.ph;        a_w[i] = wd + g.pw; }
    a_base[i] = (long long)b * g.Hs * g.Ws;
  }
  int b_n[BR];
#pragma unroll
  for (int i = 0; i < BR; ++i) b_n[i] = n0 + lr + i * 64;

  const int nchunks = ceil_div(g.Kd, BK);
  const int T = nseg * nchunks;

  float4 ra[AR], rb[BR];
  auto load_global = [&](int it) {
    int seg = it / nchunks;
    int r = (it - seg * nchunks) * BK + l4 * 4;
    const float* Ap = segA[seg];
    const float* Bp = segB[seg];
    bool rok = r < g.Kd;
    int tap = rok ? r / g.Cs : 0;
    int c = r - tap * g.Cs;
    int kh = tap / g.KW, kw = tap - kh * g.KW;
#pragma unroll
    for (int i = 0; i < AR; ++i) {
      int hs, ws;
      bool ok = rok && a_ok[i];
      if (g.mode == 0) {
        hs = a_h[i] + kh; ws = a_w[i] + kw;
      } else {
        int th = a_h[i] - kh, tw = a_w[i] - kw;
        ok = ok && th >= 0 && tw >= 0;
        hs = th / g.sh; ws = tw / g.sw;
        ok = ok && (hs * g.sh == th) && (ws * g.sw == tw);
      }
      ok = ok && hs >= 0 && hs < g.Hs && ws >= 0 && ws < g.Ws;
      ra[i] = ok ? ldg4(Ap + ((a_base[i] + (long long)hs * g.Ws + ws) * g.Cs + c))
                 : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < BR; ++i) {
      bool ok = rok && b_n[i] < g.N;
      rb[i] = ok ? ldg4(Bp + ((long long)b_n[i] * g.Kd + r)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int i = 0; i < AR; ++i) {
      int row = lr + i * 64;
      As[buf][l4 * 4 + 0][row] = ra[i].x; As[buf][l4 * 4 + 1][row] = ra[i].y;
      As[buf][l4 * 4 + 2][row] = ra[i].z; As[buf][l4 * 4 + 3][row] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < BR; ++i) {
      int row = lr + i * 64;
      Bs[buf][l4 * 4 + 0][row] = rb[i].x; Bs[buf][l4 * 4 + 1][row] = rb[i].y;
      Bs[buf][l4 * 4 + 2][row] = rb[i].z; Bs[buf][l4 * 4 + 3][row] = rb[i].w;
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  if (T > 0) {
    load_global(0);
    store_smem(0);
    __syncthreads();
    for (int it = 0; it < T; ++it) {
      const int buf = it & 1;
      if (it + 1 < T) load_global(it + 1);
      simt_tile_fma<BM, BN, LDA, LDB>(As[buf], Bs[buf], tx, ty, acc);
      if (it + 1 < T) store_smem(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue
  const float* bias = (slot == 0) ? p.bias
                                  : (p.bias_t ? p.bias_t + (long long)(slot - 1) * p.bias_slot : nullptr);
  float* outp = p.out + (long long)slot * p.out_slot;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + (i / 4) * (BM / 2) + ty * 4 + (i & 3);
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN / 4; ++j) {
      int n = n0 + j * (BN / 2) + tx * 4;
      if (n >= g.Nd) continue;
      float4 v = make_float4(acc[i][j * 4 + 0], acc[i][j * 4 + 1], acc[i][j * 4 + 2], acc[i][j * 4 + 3]);
      if (bias) {
        float4 bv = ldg4(bias + n);
        v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
      }
      float4* dst = reinterpret_cast<float4*>(outp + (long long)m * g.Nd + n);
      if (p.accumulate) {
        float4 o = *dst;
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      *dst = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// wgrad GEMM:  D[n][j] = sum_m G[m][n] * gather(In)[m][j],  j = tap*Cs + c
// grid.x = n-tiles * j-tiles (j fastest), grid.y = slots, grid.z = splits over m
// ---------------------------------------------------------------------------------------------
template <int BM, int BN>
__global__ void __launch_bounds__(256) wgrad_gemm_simt(const WgradArgs p) {
  constexpr int BK = 16;
  constexpr int TM = BM / 16, TN = BN / 16;
  constexpr int LDA = BM + 4, LDB = BN + 4;
  constexpr int AF = BM / 4, BF = BN / 4;       // float4 per row
  constexpr int ARP = 256 / AF, BRP = 256 / BF;  // rows per pass
  constexpr int AP = BK / ARP, BP = BK / BRP;    // passes
  __shared__ __align__(16) float As[2][BK][LDA];
  __shared__ __align__(16) float Bs[2][BK][LDB];

  const Geom& g = p.g;
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int tiles_j = ceil_div(g.Kd, BN);
  const int tile_n = blockIdx.x / tiles_j, tile_j = blockIdx.x - tile_n * tiles_j;
  const int slot_idx = blockIdx.y;
  const int slot = p.slot0 + slot_idx;
  const int split = blockIdx.z;
  const int n0 = tile_n * BM, j0 = tile_j * BN;
  const int m_begin = split * p.m_per_split;
  const int m_end = min(g.M, m_begin + p.m_per_split);

  const float* segG[2];
  const float* segI[2];
  int nseg = 0;
  segG[0] = p.G + (long long)slot * p.G_slot; segI[0] = p.In; nseg = 1;
  if (p.second_seg && slot > 0) { segG[1] = p.G; segI[1] = p.In + (long long)slot * p.In_slot; nseg = 2; }

  const int a_r = t / AF, a_n4 = t - a_r * AF;
  const int b_r = t / BF, b_j4 = t - b_r * BF;
  const int an = n0 + a_n4 * 4;
  const bool an_ok = an < p.Ng;
  const int bj = j0 + b_j4 * 4;
  const bool bj_ok = bj < g.Kd;
  const int tap = bj_ok ? bj / g.Cs : 0;
  const int bc = bj - tap * g.Cs;
  const int kh = tap / g.KW, kw = tap - kh * g.KW;

  const int msz = max(0, m_end - m_begin);
  const int nchunks = ceil_div(msz, BK);
  const int T = nseg * nchunks;

  float4 ra[AP], rb[BP];
  auto load_global = [&](int it) {
    int seg = it / nchunks;
    int mc = m_begin + (it - seg * nchunks) * BK;
    const float* Gp = segG[seg];
    const float* Ip = segI[seg];
#pragma unroll
    for (int i = 0; i < AP; ++i) {
      int m = mc + a_r + i * ARP;
      bool ok = an_ok && m < m_end;
      ra[i] = ok ? ldg4(Gp + ((long long)m * p.Ng + an)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < BP; ++i) {
      int m = mc + b_r + i * BRP;
      bool ok = bj_ok && m < m_end;
      int mm = ok ? m : 0;
      int b = mm / (g.Hd * g.Wd);
      int rem = mm - b * (g.Hd * g.Wd);
      int hd = rem / g.Wd, wd = rem - hd * g.Wd;
      int hs = hd * g.sh - g.ph + kh, ws = wd * g.sw - g.pw + kw;
      ok = ok && hs >= 0 && hs < g.Hs && ws >= 0 && ws < g.Ws;
      rb[i] = ok ? ldg4(Ip + ((((long long)b * g.Hs + hs) * g.Ws + ws) * g.Cs + bc))
                 : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int i = 0; i < AP; ++i)
      *reinterpret_cast<float4*>(&As[buf][a_r + i * ARP][a_n4 * 4]) = ra[i];
#pragma unroll
    for (int i = 0; i < BP; ++i)
      *reinterpret_cast<float4*>(&Bs[buf][b_r + i * BRP][b_j4 * 4]) = rb[i];
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  if (T > 0) {
    load_global(0);
    store_smem(0);
    __syncthreads();
    for (int it = 0; it < T; ++it) {
      const int buf = it & 1;
      if (it + 1 < T) load_global(it + 1);
      simt_tile_fma<BM, BN, LDA, LDB>(As[buf], Bs[buf], tx, ty, acc);
      if (it + 1 < T) store_smem(buf ^ 1);
      __syncthreads();
    }
  }

  float* outp = p.partial + ((long long)split * p.nslots + slot_idx) * (long long)g.N * g.Kd;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int n = n0 + (i / 4) * (BM / 2) + ty * 4 + (i & 3);
    if (n >= g.N) continue;
#pragma unroll
    for (int j = 0; j < TN / 4; ++j) {
      int jj = j0 + j * (BN / 2) + tx * 4;
      if (jj >= g.Kd) continue;
      *reinterpret_cast<float4*>(outp + (long long)n * g.Kd + jj) =
          make_float4(acc[i][j * 4 + 0], acc[i][j * 4 + 1], acc[i][j * 4 + 2], acc[i][j * 4 + 3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// plain dense GEMM on row-major matrices: C = alpha * op(A) op(B) + beta*C   (Kronecker apply etc.)
// op(A) is [M, Kd], op(B) is [Kd, N].  transA: A stored [Kd, M]; transB: B stored [N, Kd].
// Scalar loads (arbitrary leading dimensions / alignment), 64x64 tiles.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dense_gemm_simt(int transA, int transB, int M, int N, int Kd,
                                                       float alpha, const float* __restrict__ A, int lda,
                                                       const float* __restrict__ B, int ldb, float beta,
                                                       float* __restrict__ C, int ldc,
                                                       long long strideA, long long strideB,
                                                       long long strideC) {
  constexpr int BM = 64, BN = 64, BK = 16, LDA = BM + 4, LDB = BN + 4;
  __shared__ __align__(16) float As[BK][LDA];
  __shared__ __align__(16) float Bs[BK][LDB];
  A += blockIdx.z * strideA; B += blockIdx.z * strideB; C += blockIdx.z * strideC;
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < Kd; k0 += BK) {
    // A tile: 64 x 16 = 1024 elements, 4 per thread
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = t + i * 256;
      int r, m;
      if (transA) { m = e & 63; r = e >> 6; } else { r = e & 15; m = e >> 4; }
      int gm = m0 + m, gr = k0 + r;
      float v = 0.f;
      if (gm < M && gr < Kd) v = transA ? __ldg(A + (long long)gr * lda + gm) : __ldg(A + (long long)gm * lda + gr);
      As[r][m] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = t + i * 256;
      int r, n;
      if (transB) { r = e & 15; n = e >> 4; } else { n = e & 63; r = e >> 6; }
      int gn = n0 + n, gr = k0 + r;
      float v = 0.f;
      if (gn < N && gr < Kd) v = transB ? __ldg(B + (long long)gn * ldb + gr) : __ldg(B + (long long)gr * ldb + gn);
      Bs[r][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float* c = C + (long long)m * ldc + n;
      float v = alpha * acc[i][j];
      if (beta != 0.f) v += beta * *c;
      *c = v;
    }
  }
}

}  // namespace curv
