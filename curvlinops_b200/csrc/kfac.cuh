// KFAC-expand Kronecker-factor accumulation (included at the end of engine.cu).
//
//   A_l += wA/S_l * sum_{n,s} a~ a~^T     a~ = im2col patch of the layer input, (c, kh, kw) order like
//                                         F.unfold, optionally with a trailing 1 (joint weight+bias)
//   G_l += wG * sum_{v,n,s} g g^T         g = grad_outputs[v] back-propagated to the layer output
// (reference curvlinops/computers/kfac_hooks.py:318-393, kfac_math.py:47-203).
//
// Both are Gram matrices  X^T X  of a tall matrix X [rows, width]; they run on the wgrad contraction
// kernels (tcgen05 when the layer is large enough, fp32 SIMT otherwise) with a 1x1 "convolution"
// geometry and the deterministic split-reduce finish.  A needs the patch matrix: it is materialised once
// per layer in (c, kh, kw) order (so no permutation of the factor is needed afterwards).

namespace curv {

// patches[m][c*taps + tap] = in[src(m, tap)][c];  column Kreal = 1 if joint; columns up to widthp = 0
__global__ void im2col_kernel(const float* __restrict__ in, float* __restrict__ patches, Geom g, int C,
                              int joint, int widthp) {
  const int taps = g.KH * g.KW;
  const int Kreal = C * taps;
  const long long total = (long long)g.M * widthp;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(e % widthp);
    const int m = (int)(e / widthp);
    float v = 0.f;
    if (col < Kreal) {
      const int c = col / taps, tap = col - c * taps;
      const int kh = tap / g.KW, kw = tap - kh * g.KW;
      const int b = m / (g.Hd * g.Wd);
      const int rem = m - b * (g.Hd * g.Wd);
      const int hd = rem / g.Wd, wd = rem - hd * g.Wd;
      const int hs = hd * g.sh - g.ph + kh, ws = wd * g.sw - g.pw + kw;
      if (hs >= 0 && hs < g.Hs && ws >= 0 && ws < g.Ws)
        v = __ldg(in + (((long long)b * g.Hs + hs) * g.Ws + ws) * g.Cs + c);
    } else if (col == Kreal && joint) {
      v = 1.f;
    }
    patches[e] = v;
  }
}

// F[i][j] += w * sum_splits partial[s][i][j]   (F dense [width][width], partial rows have stride widthp)
__global__ void gram_finish_kernel(const float* __restrict__ partial, int nsplit, int width, int widthp,
                                   float* __restrict__ F, float w) {
  const long long total = (long long)width * width;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e % width);
    const int i = (int)(e / width);
    float s = 0.f;
    for (int sp = 0; sp < nsplit; ++sp)
      s += __ldg(partial + ((long long)sp * width + i) * widthp + j);
    F[e] += w * s;
  }
}

// ---- tensor-core Gram path (round 2): the tall matrix lives as operand planes (fp16 hi/lo, or one bf16 plane in
// bf16 programs) and X^T X runs on wgrad_gemm_hs: the gathered "input" operand is X itself (1x1 geometry, 128
// columns i per tile) and the "cotangent slots" are up to 8 adjacent column blocks of W8 columns of the SAME
// matrix (G_ld = ld, G_slot = W8), so one N = 256 MMA multiplies a 128-column tile against 256 other columns.

// patch planes P[m][tap*C + c] = s * in[src(m, tap)][c] (columns in (tap, c) order: 8-column chunks are channel
// runs of one tap when C % 8 == 0; the factor is permuted to F.unfold's (c, tap) order by gram_hs_finish_kernel),
// column taps*C = s * 1 if joint, up to ld zero.  lo == nullptr: one bf16 plane.
__global__ void __launch_bounds__(256) im2col_planes_kernel(const float* __restrict__ in, __half* __restrict__ hi,
                                                           __half* __restrict__ lo, Geom g, int C, int joint, int ld,
                                                           const uint32_t* __restrict__ bits) {
  const int taps = g.KH * g.KW;
  const int Kreal = C * taps;
  const int chunks = ld >> 3;
  const float sc = hs_pow2(hs_shift_from_bits(bits[0]));
  const long long total = (long long)g.M * chunks;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(e % chunks);
    const int m = (int)(e / chunks);
    const int b = m / (g.Hd * g.Wd);
    const int rem = m - b * (g.Hd * g.Wd);
    const int hd = rem / g.Wd, wd = rem - hd * g.Wd;
    float x[8];
    if ((C & 7) == 0 && ch * 8 < Kreal) {  // 8 consecutive channels of one filter tap: two 16-byte loads
      const int col = ch * 8;
      const int tap = col / C, c = col - tap * C;
      const int kh = tap / g.KW, kw = tap - kh * g.KW;
      const int hs = hd * g.sh - g.ph + kh, ws = wd * g.sw - g.pw + kw;
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
      if (hs >= 0 && hs < g.Hs && ws >= 0 && ws < g.Ws) {
        const float4* q = reinterpret_cast<const float4*>(in + (((long long)b * g.Hs + hs) * g.Ws + ws) * g.Cs + c);
        v0 = __ldg(q); v1 = __ldg(q + 1);
      }
      if (lo == nullptr) { reinterpret_cast<uint4*>(hi)[e] = hs_bf16x8(v0, v1, sc); continue; }
      uint4 h, l;
      hs_split8(v0, v1, sc, h, l);
      reinterpret_cast<uint4*>(hi)[e] = h;
      reinterpret_cast<uint4*>(lo)[e] = l;
      continue;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = ch * 8 + j;
      float v = 0.f;
      if (col < Kreal) {
        const int tap = col / C, c = col - tap * C;
        const int kh = tap / g.KW, kw = tap - kh * g.KW;
        const int hs = hd * g.sh - g.ph + kh, ws = wd * g.sw - g.pw + kw;
        if (hs >= 0 && hs < g.Hs && ws >= 0 && ws < g.Ws)
          v = __ldg(in + (((long long)b * g.Hs + hs) * g.Ws + ws) * g.Cs + c);
      } else if (col == Kreal && joint) {
        v = 1.f;
      }
      x[j] = v;
    }
    const float4 v0 = make_float4(x[0], x[1], x[2], x[3]), v1 = make_float4(x[4], x[5], x[6], x[7]);
    if (lo == nullptr) { reinterpret_cast<uint4*>(hi)[e] = hs_bf16x8(v0, v1, sc); continue; }
    uint4 h, l;
    hs_split8(v0, v1, sc, h, l);
    reinterpret_cast<uint4*>(hi)[e] = h;
    reinterpret_cast<uint4*>(lo)[e] = l;
  }
}

// dst[0..n) = max(*src, floor_bits)   (bit patterns of non-negative floats; src may be null)
__global__ void hs_bits_fill_kernel(uint32_t* dst, int n, const uint32_t* src, uint32_t floor_bits) {
  const uint32_t v = src ? max(*src, floor_bits) : floor_bits;
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = v;
}

// F[r][q] += w * sum_splits partial[sp][col(r)][col(q)], partial split = [rows_per_split][ld] floats.
// taps > 1: F is in (c, tap) order, the planes in (tap, c) order (joint column last in both).
__global__ void gram_hs_finish_kernel(const float* __restrict__ partial, int nsplit, long long split_elems, int ld,
                                      int width, int C, int taps, float* __restrict__ F, float w) {
  const long long total = (long long)width * width;
  const int Kreal = C * taps;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(e % width);
    const int r = (int)(e / width);
    const int cr = r < Kreal ? (r % taps) * C + r / taps : r;
    const int cq = q < Kreal ? (q % taps) * C + q / taps : q;
    float s = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) s += __ldg(partial + sp * split_elems + (long long)cr * ld + cq);
    F[e] += w * s;
  }
}

}  // namespace curv

// plan of the tensor-core Gram: column block width W8, blocks NS, splits
struct GramHsPlan { int W8, NS, nsplit, m_per_split; long long partial_elems; };
static GramHsPlan gram_hs_plan(long long rows, int ld, int planes) {
  GramHsPlan p;
  p.W8 = 64 * ceil_div(ld, 512);
  p.NS = ceil_div(ld, p.W8);
  const long long tiles = (long long)ceil_div(ld, 128) * (p.W8 / 64);
  long long want = (2 * 148 + tiles - 1) / tiles;
  const long long by_len = (rows + (planes == 1 ? 4096 : 2048) - 1) / (planes == 1 ? 4096 : 2048);
  long long ns = want > by_len ? want : by_len;
  const long long maxsplit = (rows + 255) / 256;
  if (ns > maxsplit) ns = maxsplit;
  if (ns < 1) ns = 1;
  p.m_per_split = (int)(((rows + ns - 1) / ns + 15) / 16 * 16);
  p.nsplit = (int)((rows + p.m_per_split - 1) / p.m_per_split);
  p.partial_elems = (long long)p.nsplit * p.NS * p.W8 * ld;
  return p;
}
static long long gram_hs_plan_elems(long long rows, int ld, int planes) {
  return gram_hs_plan(rows, ld, planes).partial_elems + 64;
}
// F[width][width] += w * X^T X for X given as planes [rows][ld] (scale word sbits, replicated in sbits[0..8));
// (C, taps): column permutation of gram_hs_finish_kernel (taps = 1: none)
static int gram_hs(const __half* Xh, const __half* Xl, int planes, long long rows, int width, int ld,
                   const uint32_t* sbits, int C, int taps, float* F, float w, float* partial,
                   long long partial_elems, cudaStream_t st) {
  if (rows >= (1LL << 31) - 4096) return fail(CURV_ERR_INVALID, "too many rows for a Gram matrix");
  const GramHsPlan pl = gram_hs_plan(rows, ld, planes);
  if (pl.partial_elems > partial_elems)
    return fail(CURV_ERR_WORKSPACE, "KFAC scratch too small (program not created with the kfac flag?)");
  HsWgradArgs a;
  memset(&a, 0, sizeof(a));
  Geom& g = a.g;
  g.B = (int)rows; g.Hs = g.Ws = g.Hd = g.Wd = 1; g.Cs = ld; g.KH = g.KW = 1; g.sh = g.sw = 1;
  g.ph = g.pw = 0; g.mode = 0; g.N = pl.W8; g.Nd = pl.W8; g.Kd = ld; g.M = (int)rows;
  a.Gh = Xh; a.Gl = Xl; a.G_slot = pl.W8; a.G_ld = ld; a.Ng = pl.W8; a.g_bits = sbits;
  a.Ih = Xh; a.Il = Xl; a.i_bits = sbits;
  a.partial = partial; a.nsplit = pl.nsplit; a.nslots = pl.NS; a.slot0 = 0; a.m_per_split = pl.m_per_split;
  a.planes = planes;
  {
    ProfScope prof(1, 2.0 * (double)rows * width * width, st);
    if (hs_launch_wgrad(a, st)) return fail(CURV_ERR_CUDA, "tensor-core Gram launch failed");
    ++g_launches;
  }
  gram_hs_finish_kernel<<<grid1d((long long)width * width), 256, 0, st>>>(
      partial, pl.nsplit, (long long)pl.NS * pl.W8 * ld, ld, width, C, taps, F, w);
  LAUNCH_CHECK();
  return CURV_OK;
}

// Gram matrix of X [rows][widthp] (first `width` columns real): F += w * X^T X
static int gram_accumulate(const float* X, long long rows, int width, int widthp, float* F, float w,
                           float* partial, long long partial_elems, cudaStream_t st) {
  if (rows >= (1LL << 31)) return fail(CURV_ERR_INVALID, "too many rows for a Gram matrix");
  WgradArgs a;
  memset(&a, 0, sizeof(a));
  Geom& g = a.g;
  g.B = (int)rows; g.Hs = g.Ws = g.Hd = g.Wd = 1; g.Cs = widthp; g.KH = g.KW = 1; g.sh = g.sw = 1;
  g.ph = g.pw = 0; g.mode = 0; g.N = width; g.Nd = widthp; g.Kd = widthp; g.M = (int)rows;
  a.G = X; a.G_slot = 0; a.Ng = widthp; a.In = X; a.In_slot = 0; a.second_seg = 0;
  a.nslots = 1; a.slot0 = 0;
  const int bm = width > 64 ? 128 : 64, bn = widthp > 64 ? 128 : 64;
  int nsplit = gram_nsplit(rows, width, widthp);
  if ((long long)nsplit * width * widthp > partial_elems)
    return fail(CURV_ERR_WORKSPACE, "KFAC scratch too small (program not created with the kfac flag?)");
  a.m_per_split = (int)(((rows + nsplit - 1) / nsplit + 31) / 32 * 32);
  a.nsplit = (int)((rows + a.m_per_split - 1) / a.m_per_split);
  a.partial = partial;
  int rc = launch_wgrad(a, bm, bn, st, 2.0 * (double)rows * width * width);
  if (rc) return rc;
  gram_finish_kernel<<<grid1d((long long)width * width), 256, 0, st>>>(partial, a.nsplit, width, widthp, F, w);
  LAUNCH_CHECK();
  return CURV_OK;
}

extern "C" int curv_kfac_accumulate_batch(curv_program* P, const void* const* param_ptrs,
                                          const void* const* const_ptrs, const void* X,
                                          const int* layer_nodes, int n_layers, float* const* A_ptrs,
                                          float* const* G_ptrs, const int* joint_bias,
                                          const float* grad_outputs, int V, float wA, float wG,
                                          void* workspace, size_t workspace_bytes, void* stream) {
  if (!P) return fail(CURV_ERR_INVALID, "null program");
  if (!(P->hessian & 2)) return fail(CURV_ERR_INVALID, "program was not created with the kfac flag (2)");
  if (workspace_bytes < P->ws_bytes || !workspace) return fail(CURV_ERR_WORKSPACE, "workspace too small");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(CURV_ERR_CUDA, "no CUDA device: curvb200 has no CPU fallback");
  Ctx c;
  c.P = P; c.ws = (float*)workspace; c.pp = param_ptrs; c.cp = const_ptrs; c.V = nullptr; c.out = nullptr;
  c.K = 0; c.ldk = 0; c.k0 = 0; c.alpha = 0.f; c.st = (cudaStream_t)stream; c.kind = CURV_KIND_VJP;
  c.rop = false;
  cudaStream_t st = c.st;
  int rc;
  // the forward convolutions and the dgrad of the seed back-propagation run on the half-split kernels; the Gram
  // matrices (one slot, patch matrix operand) stay on the 3xTF32 / SIMT wgrad kernels
  std::vector<char> hs_valid;
  if (g_tc_mode && !(g_tc_disable & 32) && P->hs1_elems > 0 && hs_ready() > 0) {
    c.hs = true;
    c.planes = (P->hessian & 4) ? 1 : 2;
    hs_valid.assign((size_t)P->hsbits_count, 0);
    c.hs_valid = &hs_valid;
    CHECK_CUDA(cudaMemsetAsync(c.hsbits(), 0, (size_t)P->hsbits_count * sizeof(uint32_t), c.st));
  }
  if ((rc = prepare_params(c, false))) return rc;
  if ((rc = forward(c, X, 0))) return rc;
  float* scratch = c.ws + P->scratch_off;
  // ---- input covariances
  for (int i = 0; i < n_layers; ++i) {
    if (!A_ptrs || !A_ptrs[i]) continue;
    if (layer_nodes[i] < 0 || layer_nodes[i] >= (int)P->nodes.size() ||
        P->nodes[layer_nodes[i]].d.op != CURV_OP_CONV)
      return fail(CURV_ERR_INVALID, "layer_nodes must reference CONV nodes");
    const Node& n = P->nodes[layer_nodes[i]];
    const Value& vi = P->values[n.d.in0];
    const Geom& g = n.fwd;
    const int width = vi.C * g.KH * g.KW + (joint_bias[i] ? 1 : 0);
    if (c.hs && g.M >= 256 && width >= 16 && !(g_tc_disable & 1024)) {
      // tensor-core Gram: patch matrix as operand planes in the scratch region, partials behind it
      const int ld = pad8(width);
      const long long plane_halves = align_up((long long)g.M * ld + 1024, 128);  // + slack: blocks read past a row end
      const long long plane_floats = plane_halves * c.planes / 2;
      const GramHsPlan pl = gram_hs_plan(g.M, ld, c.planes);
      if (plane_floats + pl.partial_elems > P->scratch_elems)
        return fail(CURV_ERR_WORKSPACE, "KFAC scratch too small for the patch planes");
      __half* ph = reinterpret_cast<__half*>(scratch);
      __half* plo = c.planes == 1 ? nullptr : ph + plane_halves;
      const int ni = layer_nodes[i];
      uint32_t* sb = c.hsbits() + c.bits_node(ni) + (1 + P->kmax);  // 8 copies of the patch scale
      if ((rc = hs_absmax(c, c.act(n.d.in0), 0, vi.slot_elems, c.bits_act(n.d.in0), 1))) return rc;
      hs_bits_fill_kernel<<<1, 32, 0, st>>>(sb, 8, c.planes == 1 ? nullptr : c.hsbits() + c.bits_act(n.d.in0),
                                            (c.planes == 1 || !joint_bias[i]) ? 0u : 0x3f800000u);
      LAUNCH_CHECK();
      im2col_planes_kernel<<<grid1d((long long)g.M * (ld / 8)), 256, 0, st>>>(c.act(n.d.in0), ph, plo, g, vi.C,
                                                                            joint_bias[i] ? 1 : 0, ld, sb);
      LAUNCH_CHECK();
      const float w = wA / (float)(g.Hd * g.Wd);
      if ((rc = gram_hs(ph, plo, c.planes, g.M, width, ld, sb, vi.C, g.KH * g.KW, A_ptrs[i], w,
                        scratch + plane_floats, P->scratch_elems - plane_floats, st)))
        return rc;
      continue;
    }
    const int widthp = pad4(width);
    const long long pelems = (long long)g.M * widthp;
    if (pelems > P->scratch_elems) return fail(CURV_ERR_WORKSPACE, "KFAC scratch too small for patches");
    im2col_kernel<<<grid1d(pelems), 256, 0, st>>>(c.act(n.d.in0), scratch, g, vi.C, joint_bias[i] ? 1 : 0,
                                                  widthp);
    LAUNCH_CHECK();
    const float w = wA / (float)(g.Hd * g.Wd);
    if ((rc = gram_accumulate(scratch, g.M, width, widthp, A_ptrs[i], w, scratch + align_up(pelems, 64),
                              P->scratch_elems - align_up(pelems, 64), st)))
      return rc;
  }
  // ---- gradient covariances: back-propagate the V seed vectors, kmax at a time
  const int last = P->nodes.back().d.out;
  const Value& vl = P->values[last];
  if (V > 0 && G_ptrs && vl.tan) {
    std::vector<float*> gmap(P->nodes.size(), nullptr);
    for (int i = 0; i < n_layers; ++i)
      if (G_ptrs[i]) gmap[layer_nodes[i]] = G_ptrs[i];
    for (int v0 = 0; v0 < V; v0 += P->kmax) {
      const int kk = V - v0 < P->kmax ? V - v0 : P->kmax;
      import_pred_kernel<<<grid1d((long long)P->B * vl.Cp * kk), 256, 0, st>>>(
          c.grad(last), vl.slot_elems, 1, grad_outputs, P->B, vl.C, vl.Cp, kk, V, v0);
      LAUNCH_CHECK();
      c.K = kk;
      c.kfac_G = gmap.data();
      c.kfac_wG = wG;
      if (c.hs) {  // new seeds: the cotangent maxima must be recomputed (atomicMax over the old words = valid bounds)
        for (size_t v = 0; v < P->values.size(); ++v)
          for (int sl = 0; sl <= P->kmax; ++sl) hs_valid[(size_t)c.bits_grad((int)v) + sl] = 0;
      }
      if ((rc = backward(c, kk))) return rc;
    }
  }
  return CURV_OK;
}
