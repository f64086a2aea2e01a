// EKFAC eigenvalue correction on the device (included at the end of engine.cu):
//     lambda_l[i][j] += w * sum_{v, n} ( sum_s (g_{v,n,s} Q_g)[i] * (a~_{n,s} Q_a)[j] )^2
// (reference curvlinops/computers/ekfac_hooks.py:25-238: per-example gradients in the Kronecker eigenbasis, squared,
// summed over examples and back-propagated vectors).  Round 1 did this on the host (F.unfold, a Python loop over the
// examples); here it is four launches per layer and back-propagated vector, all on the tcgen05 kernels:
//   1. patch planes of the layer input (im2col_planes_kernel of kfac.cuh, (tap, channel) column order)
//   2. at = patches . Q_a   and   gt = g . Q_g     two K-major GEMMs (gather_gemm_hs, 1x1 geometry); the rows of Q_a
//      are permuted to the plane order while its weight image is packed
//   3. per-example contraction  E_n = gt_n^T at_n  on wgrad_gemm_hs with the pixel range split AT THE EXAMPLE
//      BOUNDARIES (one split = one example, or an integer fraction of one if an example has more than 2048
//      positions): the split-K partials ARE the per-example gradients; column blocks of gt as slots (N = 256 MMAs)
//   4. ekfac_finish_kernel: sum the chunks of an example, square, sum over examples into lambda (fixed order)
// fp32 operators only (operands as fp16 hi/lo planes); bias-only groups stay on the host (a column sum).

namespace curv {

// lam[r][j] += w * sum_n ( sum_{c < cps} partial[(n * cps + c) * split_elems + r * ld + j] )^2
// perm_taps > 1: lam is in the canonical (channel, tap) column order while the partials are in the (tap, channel)
// order of the patch planes (the un-rotated use: exact / MC GGN diagonal); columns >= perm_C * perm_taps (joint) stay.
__global__ void __launch_bounds__(256) ekfac_finish_kernel(const float* __restrict__ partial, int nsamples, int cps,
                                                          long long split_elems, int ld, int d_out, int width,
                                                          float* __restrict__ lam, float w, int perm_C, int perm_taps) {
  const long long total = (long long)d_out * width;
  const int CT = perm_C * perm_taps;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    int j = (int)(e % width);
    if (perm_taps > 1 && j < CT) j = (j % perm_taps) * perm_C + j / perm_taps;
    const long long r = e / width;
    const float* q = partial + r * ld + j;
    float acc = 0.f;
    for (int n = 0; n < nsamples; ++n) {
      float s = 0.f;
      for (int c = 0; c < cps; ++c) s += __ldg(q + ((long long)n * cps + c) * split_elems);
      acc = fmaf(s, s, acc);
    }
    lam[e] += w * acc;
  }
}

// Weight image of  Wm[n][k] = Q[row(k)][n]  (n < N eigenvectors, k < Kreal plane columns), Q row-major [Kreal][N]:
// row(k) maps the (tap, channel) plane order to Q_a's F.unfold (channel, tap) row order (joint column last); taps = 1:
// identity.  Layout of hs_pack_image_kernel.
__global__ void __launch_bounds__(256) ekfac_pack_q_kernel(const float* __restrict__ Q, int N, int Kreal, int C, int taps,
                                                          __half* __restrict__ dst, int BN, int tiles_n, int nchunks,
                                                          const uint32_t* __restrict__ bits) {
  const float sc = hs_pow2(hs_shift_from_bits(bits[0]));
  const int CT = C * taps;
  const long long total = (long long)tiles_n * nchunks * BN * 8;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e & 7);
    long long t = e >> 3;
    const int r = (int)(t % BN); t /= BN;
    const int kc = (int)(t % nchunks);
    const int tn = (int)(t / nchunks);
    const int n = tn * BN + r, k0 = kc * HS_BK + c * 8;
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + j;
      float v = 0.f;
      if (n < N && k < Kreal) {
        const int row = k < CT ? (k % C) * taps + k / C : k;
        v = __ldg(Q + (long long)row * N + n);
      }
      x[j] = v;
    }
    uint4 h, l;
    hs_split8(make_float4(x[0], x[1], x[2], x[3]), make_float4(x[4], x[5], x[6], x[7]), sc, h, l);
    __half* blk = dst + ((long long)tn * nchunks + kc) * (2 * BN * HS_BK);
    const int o = ((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) >> 1;
    *reinterpret_cast<uint4*>(blk + o) = h;
    *reinterpret_cast<uint4*>(blk + BN * HS_BK + o) = l;
  }
}

}  // namespace curv

// chunks per example so that a chunk has at most 2048 positions and examples split evenly (0: not possible)
static int ekfac_chunks_per_sample(int S) {
  for (int c = 1; c <= 64; ++c)
    if (S % c == 0 && S / c <= 2048) return c;
  return 0;
}

struct EkfacPlan {
  int width, ldp, lda, ldg, W8, NS, cps;
  long long p_halves, qa_img, qg_img, at_elems, gt_elems, at_halves, gt_halves, split_elems, partial_elems;
  long long o_bits, o_P, o_QA, o_QG, o_atF, o_atP, o_gtF, o_gtP, o_part, total_floats;
};
// identity: no rotation (Q_a = Q_g = I) - the contraction reads the patch planes and the cotangent planes directly
static EkfacPlan ekfac_plan(long long M, int B, int S, int width, int Cp_out, bool identity = false) {
  EkfacPlan p;
  p.width = width;
  p.ldp = (width + 63) & ~63;       // patch planes = reduction of the at-GEMM: multiples of 64 take the TMA producers
  p.lda = identity ? p.ldp : pad8(width);  // rows of at (output of that GEMM, gathered operand of the contraction)
  p.ldg = pad8(Cp_out);             // gt rows
  p.W8 = 64 * ceil_div(p.ldg, 512);
  p.NS = ceil_div(p.ldg, p.W8);
  p.cps = ekfac_chunks_per_sample(S);
  p.p_halves = align_up(M * p.ldp + 8192, 128);
  p.qa_img = identity ? 0 : hs_image_halves(p.lda, p.ldp);
  p.qg_img = identity ? 0 : hs_image_halves(p.ldg, Cp_out);
  p.at_elems = identity ? 0 : M * p.lda;
  p.gt_elems = identity ? 0 : M * p.ldg;
  p.at_halves = align_up(p.at_elems + 8192, 128);
  p.gt_halves = align_up(p.gt_elems + 8192, 128);
  p.split_elems = (long long)p.NS * p.W8 * p.lda;
  p.partial_elems = (long long)B * (p.cps > 0 ? p.cps : 1) * p.split_elems + 64;
  long long o = 0;  // in floats
  auto take = [&](long long floats) { long long r = o; o = align_up(o + floats, 256); return r; };
  p.o_bits = take(64);
  p.o_P = take(p.p_halves);        // 2 planes x 2 bytes = 1 float per element
  p.o_QA = take(p.qa_img / 2 + 1);
  p.o_QG = take(p.qg_img / 2 + 1);
  p.o_atF = take(p.at_elems);
  p.o_atP = take(p.at_halves);
  p.o_gtF = take(p.gt_elems);
  p.o_gtP = take(p.gt_halves);
  p.o_part = take(p.partial_elems);
  p.total_floats = o;
  return p;
}
static long long ekfac_scratch_elems(long long M, int B, int S, int d_in_width, int Cp_out) {
  const long long a = ekfac_plan(M, B, S, d_in_width, Cp_out).total_floats;
  const long long b = ekfac_plan(M, B, S, d_in_width, Cp_out, true).total_floats;
  return (a > b ? a : b) + 64;
}

// one layer, one back-propagated vector (cotangent slot `slot_index` of the planes in hs1, absolute slot `abs_slot`)
static int ekfac_layer(const Ctx& c, const Node& n, const EkfacEntry& en, int slot_index, int abs_slot) {
  curv_program* P = c.P;
  cudaStream_t st = c.st;
  const EkfacJob& job = *c.ekfac;
  const Value& vi = P->values[n.d.in0];
  const Value& vo = P->values[n.d.out];
  const Geom& g = n.fwd;
  const int S = vo.H * vo.W;
  const int Cin = en.bias_only ? 0 : vi.C;            // bias group: the patch is the ones column alone
  const int joint = (en.bias_only || en.joint) ? 1 : 0;
  const int width = Cin * g.KH * g.KW + joint;
  const bool identity = en.identity != 0;
  const EkfacPlan p = ekfac_plan(g.M, P->B, S, width, vo.Cp, identity);
  if (identity && pad8(vo.Cp) != vo.Cp) return fail(CURV_ERR_UNSUPPORTED, "GGN diagonal: channel padding");
  if (!c.hs || c.planes != 2) return fail(CURV_ERR_UNSUPPORTED, "EKFAC correction needs the fp32 tensor-core path");
  if (p.cps == 0) return fail(CURV_ERR_UNSUPPORTED, "EKFAC correction: output map cannot be split at example boundaries");
  if (p.total_floats > P->scratch_elems) return fail(CURV_ERR_WORKSPACE, "EKFAC scratch too small");
  float* sc = c.ws + P->scratch_off;
  uint32_t* bits = reinterpret_cast<uint32_t*>(sc + p.o_bits);  // [0] patches, [1] Qa, [2] Qg, [8..16) at, [16..24) gt
  __half* Ph = reinterpret_cast<__half*>(sc + p.o_P);
  __half* Pl = Ph + p.p_halves;
  __half* QAimg = reinterpret_cast<__half*>(sc + p.o_QA);
  __half* QGimg = reinterpret_cast<__half*>(sc + p.o_QG);
  float* atF = sc + p.o_atF;
  float* gtF = sc + p.o_gtF;
  __half* atH = reinterpret_cast<__half*>(sc + p.o_atP);
  __half* atL = atH + p.at_halves;
  __half* gtH = reinterpret_cast<__half*>(sc + p.o_gtP);
  __half* gtL = gtH + p.gt_halves;
  float* part = sc + p.o_part;
  int rc;
  CHECK_CUDA(cudaMemsetAsync(bits, 0, 64 * 4, st));
  // ---- 1. patch planes (scale: absmax of the layer input, at least 1 for the joint ones column)
  if ((rc = hs_absmax(c, c.act(n.d.in0), 0, vi.slot_elems, c.bits_act(n.d.in0), 1))) return rc;
  hs_bits_fill_kernel<<<1, 32, 0, st>>>(bits, 1, c.hsbits() + c.bits_act(n.d.in0), joint ? 0x3f800000u : 0u);
  CHECK_CUDA(cudaMemsetAsync(Ph + g.M * p.ldp, 0, 8192 * 2, st));
  CHECK_CUDA(cudaMemsetAsync(Pl + g.M * p.ldp, 0, 8192 * 2, st));
  im2col_planes_kernel<<<grid1d((long long)g.M * (p.ldp / 8)), 256, 0, st>>>(c.act(n.d.in0), Ph, Pl, g, Cin, joint,
                                                                           p.ldp, bits);
  if (!identity) {
  // ---- 2a. at = patches . Q_a (rows of Q_a permuted to the plane order)
  kron_absmax_kernel<<<grid1d((long long)width * width), 256, 0, st>>>(en.Qa, (long long)width * width, bits + 1, 1);
  {
    const int BN = tc_bn(p.lda), tiles_n = ceil_div(p.lda, BN), nchunks = ceil_div(p.ldp, HS_BK);
    ekfac_pack_q_kernel<<<hs_grid((long long)tiles_n * nchunks * BN * 8), 256, 0, st>>>(
        en.Qa, width, width, Cin, g.KH * g.KW, QAimg, BN, tiles_n, nchunks, bits + 1);
  }
  g_launches += 4;
  LAUNCH_CHECK();
  auto gemm = [&](const __half* Ah, const __half* Al, long long rows, int Cs, const uint32_t* abits,
                  const __half* Wimg, const uint32_t* wbits, int N, int Nd, float* out, double flops) -> int {
    HsGatherArgs h;
    memset(&h, 0, sizeof(h));
    Geom& q = h.g;
    q.B = (int)rows; q.Hs = q.Ws = q.Hd = q.Wd = 1; q.Cs = Cs; q.KH = q.KW = 1; q.sh = q.sw = 1; q.mode = 0;
    q.N = N; q.Nd = Nd; q.Kd = Cs; q.M = (int)rows;
    h.Ah = Ah; h.Al = Al; h.A_slot = rows * Cs; h.a_slot_base = 0; h.a_has_slots = 0; h.a_bits = abits;
    h.W_img = Wimg; h.Wt_img = nullptr; h.w_bits = wbits;
    h.out = out; h.out_slot = rows * Nd; h.slot0 = 0; h.accumulate = 0; h.planes = 2;
    ProfScope prof(0, flops, st);
    if (hs_launch_gather_gemm(h, 1, st, false, 1)) return fail(CURV_ERR_CUDA, "EKFAC rotation GEMM launch failed");
    ++g_launches;
    return CURV_OK;
  };
  if ((rc = gemm(Ph, Pl, g.M, p.ldp, bits, QAimg, bits + 1, width, p.lda, atF, 2.0 * g.M * (double)width * width)))
    return rc;
  // ---- 2b. gt = g . Q_g
  kron_absmax_kernel<<<grid1d((long long)vo.C * vo.C), 256, 0, st>>>(en.Qg, (long long)vo.C * vo.C, bits + 2, 1);
  {
    const int BN = tc_bn(p.ldg), tiles_n = ceil_div(p.ldg, BN), nchunks = ceil_div(vo.Cp, HS_BK);
    ekfac_pack_q_kernel<<<hs_grid((long long)tiles_n * nchunks * BN * 8), 256, 0, st>>>(
        en.Qg, vo.C, vo.C, vo.C, 1, QGimg, BN, tiles_n, nchunks, bits + 2);
  }
  g_launches += 2;
  LAUNCH_CHECK();
  if ((rc = gemm(c.hs1_hi() + (long long)slot_index * vo.slot_elems, c.hs1_lo() + (long long)slot_index * vo.slot_elems,
                 g.M, vo.Cp, c.hsbits() + c.bits_grad(n.d.out) + abs_slot, QGimg, bits + 2, vo.C, p.ldg, gtF,
                 2.0 * g.M * (double)vo.C * vo.C)))
    return rc;
  // ---- planes of at and gt
  CHECK_CUDA(cudaMemsetAsync(atH + p.at_elems, 0, 8192 * 2, st));
  CHECK_CUDA(cudaMemsetAsync(atL + p.at_elems, 0, 8192 * 2, st));
  CHECK_CUDA(cudaMemsetAsync(gtH + p.gt_elems, 0, 8192 * 2, st));
  CHECK_CUDA(cudaMemsetAsync(gtL + p.gt_elems, 0, 8192 * 2, st));
  kron_absmax_kernel<<<grid1d(p.at_elems), 256, 0, st>>>(atF, p.at_elems, bits + 8, 1);
  kron_absmax_kernel<<<grid1d(p.gt_elems), 256, 0, st>>>(gtF, p.gt_elems, bits + 16, 8);
  if (hs_launch_split(atF, 0, p.at_elems, atH, atL, 0, bits + 8, 1, st) ||
      hs_launch_split(gtF, 0, p.gt_elems, gtH, gtL, 0, bits + 16, 1, st))
    return fail(CURV_ERR_CUDA, "EKFAC: split of the rotated operands failed");
  g_launches += 4;
  LAUNCH_CHECK();
  }
  // ---- 3. per-example contraction: splits at the example boundaries
  {
    HsWgradArgs a;
    memset(&a, 0, sizeof(a));
    Geom& q = a.g;
    q.B = g.M; q.Hs = q.Ws = q.Hd = q.Wd = 1; q.Cs = p.lda; q.KH = q.KW = 1; q.sh = q.sw = 1; q.mode = 0;
    q.N = p.W8; q.Nd = p.W8; q.Kd = p.lda; q.M = g.M;
    a.Gh = gtH; a.Gl = gtL; a.G_slot = p.W8; a.G_ld = p.ldg; a.Ng = p.W8; a.g_bits = bits + 16;
    a.Ih = atH; a.Il = atL; a.i_bits = bits + 8;
    if (identity) {  // un-rotated: the patch planes and the cotangent planes of this slot are the operands
      a.Gh = c.hs1_hi() + (long long)slot_index * vo.slot_elems;
      a.Gl = c.hs1_lo() + (long long)slot_index * vo.slot_elems;
      hs_bits_fill_kernel<<<1, 32, 0, st>>>(bits + 16, 8, c.hsbits() + c.bits_grad(n.d.out) + abs_slot, 0u);
      LAUNCH_CHECK();
      a.Ih = Ph; a.Il = Pl; a.i_bits = bits;
    }
    a.partial = part; a.nsplit = P->B * p.cps; a.nslots = p.NS; a.slot0 = 0; a.m_per_split = S / p.cps;
    a.planes = 2;
    ProfScope prof(1, 2.0 * g.M * (double)vo.C * width, st);
    if (hs_launch_wgrad(a, st)) return fail(CURV_ERR_CUDA, "EKFAC per-example contraction launch failed");
    ++g_launches;
  }
  // ---- 4. square and sum over the examples
  ekfac_finish_kernel<<<grid1d((long long)vo.C * width), 256, 0, st>>>(part, P->B, p.cps, p.split_elems, p.lda, vo.C,
                                                                     width, en.lam, job.w, identity ? Cin : 0,
                                                                     identity ? g.KH * g.KW : 1);
  LAUNCH_CHECK();
  return CURV_OK;
}

extern "C" int curv_ekfac_correction_batch(curv_program* P, const void* const* param_ptrs,
                                           const void* const* const_ptrs, const void* X, const int* layer_nodes,
                                           int n_layers, const float* const* QA_ptrs, const float* const* QG_ptrs,
                                           float* const* lambda_ptrs, const int* joint_bias, const float* grad_outputs,
                                           int V, float w, void* workspace, size_t workspace_bytes, void* stream) {
  if (!P) return fail(CURV_ERR_INVALID, "null program");
  if (!(P->hessian & 2)) return fail(CURV_ERR_INVALID, "program was not created with the kfac flag (2)");
  if (P->hessian & 4) return fail(CURV_ERR_UNSUPPORTED, "EKFAC correction runs in fp32 programs");
  if (workspace_bytes < P->ws_bytes || !workspace) return fail(CURV_ERR_WORKSPACE, "workspace too small");
  if (!layer_nodes || !QA_ptrs || !QG_ptrs || !lambda_ptrs || !joint_bias || V < 1 || !grad_outputs)
    return fail(CURV_ERR_INVALID, "curv_ekfac_correction_batch: bad arguments");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(CURV_ERR_CUDA, "no CUDA device: curvb200 has no CPU fallback");
  Ctx c;
  c.P = P; c.ws = (float*)workspace; c.pp = param_ptrs; c.cp = const_ptrs; c.V = nullptr; c.out = nullptr;
  c.K = 0; c.ldk = 0; c.k0 = 0; c.alpha = 0.f; c.st = (cudaStream_t)stream; c.kind = CURV_KIND_VJP;
  c.rop = false;
  std::vector<char> hs_valid;
  if (!(g_tc_mode && !(g_tc_disable & 32) && P->hs1_elems > 0 && hs_ready() > 0))
    return fail(CURV_ERR_UNSUPPORTED, "EKFAC correction needs the tensor-core path (no layer of this program uses it)");
  c.hs = true;
  c.planes = 2;
  hs_valid.assign((size_t)P->hsbits_count, 0);
  c.hs_valid = &hs_valid;
  CHECK_CUDA(cudaMemsetAsync(c.hsbits(), 0, (size_t)P->hsbits_count * sizeof(uint32_t), c.st));
  EkfacJob job;
  job.has.assign(P->nodes.size(), 0);
  job.w = w;
  for (int i = 0; i < n_layers; ++i) {
    const int ni = layer_nodes[i];
    if (ni < 0 || ni >= (int)P->nodes.size() || P->nodes[ni].d.op != CURV_OP_CONV)
      return fail(CURV_ERR_INVALID, "layer_nodes must reference CONV nodes");
    // joint_bias[i] & 3: 0 weight group, 1 joint weight + bias group, 2 bias-only group (QA_ptrs[i] = [[1]]);
    // | 4: no rotation (Q_a = Q_g = I, the pointers are ignored): lambda = sum of squared per-example gradients, the
    // exact / MC GGN diagonal of the layer (curvlinops/computers/ggn_diagonal.py:49-75)
    const int ident = (joint_bias[i] & 4) ? 1 : 0, jb = joint_bias[i] & 3;
    if (!lambda_ptrs[i] || (!ident && !QG_ptrs[i])) continue;
    if (!ident && !QA_ptrs[i])
      return fail(CURV_ERR_INVALID, "curv_ekfac_correction_batch: QA missing (bias groups take [[1]])");
    job.entries.push_back({ni, QA_ptrs[i], QG_ptrs[i], lambda_ptrs[i], jb == 1, jb == 2, ident});
    job.has[ni] = 1;
  }
  int rc;
  if ((rc = prepare_params(c, false))) return rc;
  if ((rc = forward(c, X, 0))) return rc;
  const int last = P->nodes.back().d.out;
  const Value& vl = P->values[last];
  if (!vl.tan) return CURV_OK;
  std::vector<float*> gmap(P->nodes.size(), nullptr);  // non-null array: KFAC-style sweep without parameter gradients
  c.kfac_G = gmap.data();
  c.ekfac = &job;
  for (int v0 = 0; v0 < V; v0 += P->kmax) {
    const int kk = V - v0 < P->kmax ? V - v0 : P->kmax;
    import_pred_kernel<<<grid1d((long long)P->B * vl.Cp * kk), 256, 0, c.st>>>(
        c.grad(last), vl.slot_elems, 1, grad_outputs, P->B, vl.C, vl.Cp, kk, V, v0);
    LAUNCH_CHECK();
    c.K = kk;
    for (size_t v = 0; v < P->values.size(); ++v)
      for (int sl = 0; sl <= P->kmax; ++sl) hs_valid[(size_t)c.bits_grad((int)v) + sl] = 0;
    if ((rc = backward(c, kk))) return rc;
  }
  return CURV_OK;
}
