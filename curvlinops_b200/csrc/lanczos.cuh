// Full re-orthogonalisation of a Lanczos vector against the m previous ones (the eigensolver consumer of the GGN /
// Hessian products, BASELINE.json configs[4]; the reference hands this to ARPACK on the host through
// `to_scipy()`, _torch_base.py:560-592).  HBM-bound: one round is two streaming passes over Q[:m] ([m, n] fp32 rows):
//   c = Q[:m] w          lanczos_dots_kernel   (+ lanczos_dots_finish_kernel: fixed-order sum of the block partials)
//   w -= Q[:m]^T c       lanczos_update_kernel
// No atomics: every dot product is a two-level sum in a fixed order, so results are bit-wise repeatable.
// cuBLAS gemv on these shapes (m ~ 30 rows of 25.6 M columns) parallelises over the m outputs and ran at a few
// percent of the memory roofline; these kernels tile the long axis over the whole grid instead.
#pragma once

namespace curv {

constexpr int LZ_ROWS = 4;      // rows of Q per sweep over a block's chunk (w is re-read from L2 once per LZ_ROWS rows)
constexpr int LZ_THREADS = 256;

// grid: nb blocks; block b owns elements [b * chunk, min(n, (b+1) * chunk)); chunk % 4 == 0.
// partial[j * nb + b] = sum over the chunk of Q[j, i] * w[i]
template <bool VEC>
__global__ void __launch_bounds__(LZ_THREADS) lanczos_dots_kernel(const float* __restrict__ Q, long long ldq, int m,
                                                                  const float* __restrict__ w, long long n,
                                                                  long long chunk, float* __restrict__ partial) {
  __shared__ float red[LZ_THREADS / 32][LZ_ROWS];
  const long long lo = (long long)blockIdx.x * chunk;
  const long long hi = min(n, lo + chunk);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j0 = 0; j0 < m; j0 += LZ_ROWS) {
    float acc[LZ_ROWS];
#pragma unroll
    for (int r = 0; r < LZ_ROWS; ++r) acc[r] = 0.f;
    const float* q[LZ_ROWS];
#pragma unroll
    for (int r = 0; r < LZ_ROWS; ++r) q[r] = Q + (long long)min(j0 + r, m - 1) * ldq;
    if (VEC) {
      for (long long i = lo + 4ll * threadIdx.x; i < hi; i += 4ll * LZ_THREADS) {
        if (i + 4 <= hi) {
          const float4 wv = *reinterpret_cast<const float4*>(w + i);
#pragma unroll
          for (int r = 0; r < LZ_ROWS; ++r) {
            const float4 qv = __ldcs(reinterpret_cast<const float4*>(q[r] + i));
            acc[r] += qv.x * wv.x + qv.y * wv.y + qv.z * wv.z + qv.w * wv.w;
          }
        } else {
          for (long long t = i; t < hi; ++t)
#pragma unroll
            for (int r = 0; r < LZ_ROWS; ++r) acc[r] += q[r][t] * w[t];
        }
      }
    } else {
      for (long long i = lo + threadIdx.x; i < hi; i += LZ_THREADS) {
        const float wv = w[i];
#pragma unroll
        for (int r = 0; r < LZ_ROWS; ++r) acc[r] += q[r][i] * wv;
      }
    }
#pragma unroll
    for (int r = 0; r < LZ_ROWS; ++r) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
      if (lane == 0) red[warp][r] = acc[r];
    }
    __syncthreads();
    if (threadIdx.x < LZ_ROWS && j0 + threadIdx.x < m) {
      float s = 0.f;
#pragma unroll
      for (int x = 0; x < LZ_THREADS / 32; ++x) s += red[x][threadIdx.x];
      partial[(long long)(j0 + threadIdx.x) * gridDim.x + blockIdx.x] = s;
    }
    __syncthreads();
  }
}

// one warp per row j: c[j] = sum_b partial[j, b] (lane-strided, then a shuffle tree: fixed order);
// coeff_out[j] (+)= c[j] so that the caller reads the total coefficient over the rounds
__global__ void lanczos_dots_finish_kernel(const float* __restrict__ partial, int nb, int m, float* __restrict__ c,
                                           float* __restrict__ coeff_out, int accumulate) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= m) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int b = lane; b < nb; b += 32) s += partial[(long long)j * nb + b];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    c[j] = s;
    if (coeff_out) coeff_out[j] = accumulate ? coeff_out[j] + s : s;
  }
}

// w[i] -= sum_j c[j] Q[j, i]; c staged in shared memory (m floats)
template <bool VEC>
__global__ void __launch_bounds__(LZ_THREADS) lanczos_update_kernel(const float* __restrict__ Q, long long ldq, int m,
                                                                    const float* __restrict__ c, float* __restrict__ w,
                                                                    long long n) {
  extern __shared__ float cs[];
  for (int j = threadIdx.x; j < m; j += LZ_THREADS) cs[j] = c[j];
  __syncthreads();
  if (VEC) {
    const long long i = 4ll * ((long long)blockIdx.x * LZ_THREADS + threadIdx.x);
    if (i >= n) return;
    if (i + 4 <= n) {
      float4 a = *reinterpret_cast<const float4*>(w + i);
      int j = 0;
      for (; j + 4 <= m; j += 4) {
        float4 qv[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) qv[r] = __ldcs(reinterpret_cast<const float4*>(Q + (long long)(j + r) * ldq + i));
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float cj = cs[j + r];
          a.x -= cj * qv[r].x; a.y -= cj * qv[r].y; a.z -= cj * qv[r].z; a.w -= cj * qv[r].w;
        }
      }
      for (; j < m; ++j) {
        const float4 qv = __ldcs(reinterpret_cast<const float4*>(Q + (long long)j * ldq + i));
        const float cj = cs[j];
        a.x -= cj * qv.x; a.y -= cj * qv.y; a.z -= cj * qv.z; a.w -= cj * qv.w;
      }
      *reinterpret_cast<float4*>(w + i) = a;
    } else {
      for (long long t = i; t < n; ++t) {
        float a = w[t];
        for (int j = 0; j < m; ++j) a -= cs[j] * Q[(long long)j * ldq + t];
        w[t] = a;
      }
    }
  } else {
    const long long i = (long long)blockIdx.x * LZ_THREADS + threadIdx.x;
    if (i >= n) return;
    float a = w[i];
    for (int j = 0; j < m; ++j) a -= cs[j] * Q[(long long)j * ldq + i];
    w[i] = a;
  }
}

static int lanczos_blocks(long long n) {
  // 4 resident CTAs per SM; never less than ~4K elements per block
  long long nb = 148ll * 4;
  const long long cap = (n + 4095) / 4096;
  if (nb > cap) nb = cap;
  return (int)(nb < 1 ? 1 : nb);
}

}  // namespace curv

extern "C" long long curv_lanczos_reorth_workspace(int m, long long n) {
  return (long long)sizeof(float) * ((long long)m * curv::lanczos_blocks(n) + m);
}

extern "C" int curv_lanczos_reorth(const float* Q, long long ldq, int m, float* w, long long n, int rounds,
                                   float* coeff, void* ws, long long ws_bytes, void* stream) {
  using namespace curv;
  if (m < 0 || n <= 0 || rounds < 1 || (m > 0 && (!Q || !w || !ws))) return fail(CURV_ERR_INVALID, "lanczos_reorth: bad args");
  if (m == 0) return CURV_OK;
  if (m > 12000) return fail(CURV_ERR_INVALID, "lanczos_reorth: at most 12000 previous vectors");
  if (ws_bytes < curv_lanczos_reorth_workspace(m, n)) return fail(CURV_ERR_WORKSPACE, "lanczos_reorth: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = lanczos_blocks(n);
  long long chunk = (n + nb - 1) / nb;
  chunk = (chunk + 3) / 4 * 4;
  float* partial = (float*)ws;
  float* c = partial + (long long)m * nb;
  const bool vec = (ldq % 4 == 0) && (((uintptr_t)Q | (uintptr_t)w) % 16 == 0);
  for (int r = 0; r < rounds; ++r) {
    if (vec) lanczos_dots_kernel<true><<<nb, LZ_THREADS, 0, st>>>(Q, ldq, m, w, n, chunk, partial);
    else lanczos_dots_kernel<false><<<nb, LZ_THREADS, 0, st>>>(Q, ldq, m, w, n, chunk, partial);
    LAUNCH_CHECK();
    lanczos_dots_finish_kernel<<<(m + 7) / 8, 256, 0, st>>>(partial, nb, m, c, coeff, r > 0);
    LAUNCH_CHECK();
    if (vec) {
      const long long nthr = (n + 3) / 4;
      lanczos_update_kernel<true><<<(unsigned)((nthr + LZ_THREADS - 1) / LZ_THREADS), LZ_THREADS, sizeof(float) * m, st>>>(
          Q, ldq, m, c, w, n);
    } else {
      lanczos_update_kernel<false><<<(unsigned)((n + LZ_THREADS - 1) / LZ_THREADS), LZ_THREADS, sizeof(float) * m, st>>>(
          Q, ldq, m, c, w, n);
    }
    LAUNCH_CHECK();
  }
  return CURV_OK;
}
