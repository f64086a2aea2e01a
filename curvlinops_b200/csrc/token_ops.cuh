// Token-sequence glue ops of a vision transformer (class token, position embedding, class-token read-out); values are
// [slot][B][T][C] fp32 with C == Cp.  Parameters enter as broadcast operands: their tangent is the column of V, their
// gradient the sum of the cotangent over the batch (fixed order: one thread per element, loop over the examples).
#pragma once

namespace curv {

// out[s][b][0][c] = cls (s == 0) | V column s-1 of the class token (or 0);  out[s][b][1 + t][c] = in[s][b][t][c] (or 0)
__global__ void clscat_fwd_kernel(const float* __restrict__ in, long long in_slot, int in_has_slots,
                                  const float* __restrict__ cls, const float* __restrict__ vcol, int ldk,
                                  float* __restrict__ out, long long out_slot, int B, int T0, int C, int nslots) {
  const long long per = (long long)B * (T0 + 1) * C;
  const int s = blockIdx.y;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long r = i / C;
    const int t = (int)(r % (T0 + 1));
    const long long b = r / (T0 + 1);
    float v;
    if (t == 0) v = s == 0 ? __ldg(cls + c) : (vcol ? __ldg(vcol + (long long)c * ldk + (s - 1)) : 0.f);
    else v = (s == 0 || in_has_slots) ? __ldg(in + (long long)s * in_slot + (b * T0 + (t - 1)) * C + c) : 0.f;
    out[(long long)s * out_slot + i] = v;
  }
}

// gin[s][b][t][c] (+)= g[s][b][1 + t][c]
__global__ void clscat_bwd_kernel(const float* __restrict__ g, long long g_slot, float* __restrict__ gin,
                                  long long gin_slot, int B, int T0, int C, int s0, int accumulate) {
  const long long per = (long long)B * T0 * C;
  const int s = s0 + blockIdx.y;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long r = i / C;
    const int t = (int)(r % T0);
    const long long b = r / T0;
    const float v = __ldg(g + (long long)s * g_slot + (b * (T0 + 1) + t + 1) * C + c);
    float* o = gin + (long long)s * gin_slot + i;
    *o = accumulate ? *o + v : v;
  }
}

// out[(off + e) * ldk + k0 + k] += alpha * sum_b g[s0 + k][b][row0 + e / C ... ]: e indexes [rows][C] of one example,
// example stride `bstride`, first element `first` (class token: rows = 1, first = 0; position embedding: rows = T)
__global__ void batch_sum_grad_kernel(const float* __restrict__ g, long long g_slot, int B, long long bstride,
                                      long long elems, int s0, int K, float* __restrict__ out, long long off, int ldk,
                                      int k0, float alpha) {
  const long long total = elems * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const long long e = i / K;
    const float* q = g + (long long)(s0 + k) * g_slot + e;
    float a = 0.f;
    for (int b = 0; b < B; ++b) a += __ldg(q + b * bstride);
    out[(off + e) * ldk + k0 + k] += alpha * a;
  }
}

// out[s][b][t][c] = in[s][b][t][c] (or 0) + (s == 0 ? pos[t][c] : V column s-1 (or 0))
__global__ void posadd_fwd_kernel(const float* __restrict__ in, long long in_slot, int in_has_slots,
                                  const float* __restrict__ pos, const float* __restrict__ vcol, int ldk,
                                  float* __restrict__ out, long long out_slot, int B, long long tc, int nslots) {
  const long long per = (long long)B * tc;
  const int s = blockIdx.y;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i % tc;
    float v = (s == 0 || in_has_slots) ? __ldg(in + (long long)s * in_slot + i) : 0.f;
    v += s == 0 ? __ldg(pos + e) : (vcol ? __ldg(vcol + e * ldk + (s - 1)) : 0.f);
    out[(long long)s * out_slot + i] = v;
  }
}

// out[s][b][c] = in[s][b][t0][c]
__global__ void toksel_fwd_kernel(const float* __restrict__ in, long long in_slot, float* __restrict__ out,
                                  long long out_slot, int B, int T, int C, int t0) {
  const long long per = (long long)B * C;
  const int s = blockIdx.y;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long b = i / C;
    out[(long long)s * out_slot + i] = __ldg(in + (long long)s * in_slot + (b * T + t0) * C + c);
  }
}

// gin[s][b][t][c] (+)= (t == t0) ? g[s][b][c] : 0
__global__ void toksel_bwd_kernel(const float* __restrict__ g, long long g_slot, float* __restrict__ gin,
                                  long long gin_slot, int B, int T, int C, int t0, int s0, int accumulate) {
  const long long per = (long long)B * T * C;
  const int s = s0 + blockIdx.y;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long r = i / C;
    const int t = (int)(r % T);
    const long long b = r / T;
    const float v = t == t0 ? __ldg(g + (long long)s * g_slot + b * C + c) : 0.f;
    float* o = gin + (long long)s * gin_slot + i;
    *o = accumulate ? *o + v : v;
  }
}

}  // namespace curv
