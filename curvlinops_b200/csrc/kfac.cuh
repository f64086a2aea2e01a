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

}  // namespace curv

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
