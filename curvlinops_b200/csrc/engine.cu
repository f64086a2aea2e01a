// curvb200 engine: layer program, workspace plan and the two fused sweeps
//   forward + Jv      (primal slot 0, tangent slots 1..K)
//   backward + J^T    (cotangent slots 1..K, written over the dead tangents in GGN mode)
// behind the C ABI of include/curvb200.h.  Host code only plans and launches; all arithmetic is in the
// kernels of hs_gemm.cuh (default tensor-core path) / tc_gemm.cuh / gemm_simt.cuh / elementwise.cuh.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/curvb200.h"
#include "common.cuh"
#include "elementwise.cuh"
#include "gemm_simt.cuh"
#include "attention.cuh"
#include "token_ops.cuh"
#include "tc_gemm.cuh"
#include "hs_gemm.cuh"

using namespace curv;

// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static long long g_launches = 0;
static int g_tc_mode = 1;       // 0 off, 1 auto, 2 forced (tests)
static int g_tc_disable = 0;    // debug bitmask: 1 = no tcgen05 gather GEMM, 2 = no tcgen05 wgrad GEMM,
                                // 4 = gather GEMM with the old two-threads-per-row producer mapping,
                                // 8, 16 = producer experiments, 64 = BN kernels never write planes directly,
                                // 128 = no N-stacked kernel for shared-activation layers (the stem),
                                // 32 = no half-split (fp16 hi/lo) kernels:
                                // everything eligible runs the 3xTF32 kernels instead

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CHECK_CUDA(expr)                                                                      \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return fail(CURV_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));        \
  } while (0)
#define LAUNCH_CHECK()                                                                        \
  do {                                                                                        \
    ++g_launches;                                                                             \
    cudaError_t e__ = cudaGetLastError();                                                     \
    if (e__ != cudaSuccess)                                                                   \
      return fail(CURV_ERR_CUDA, std::string("kernel launch failed at ") + __FILE__ + ":" +   \
                                     std::to_string(__LINE__) + ": " + cudaGetErrorString(e__)); \
  } while (0)

static inline int grid1d(long long total, int threads = 256) {
  long long b = (total + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > 148 * 16) b = 148 * 16;
  return (int)b;
}

// ------------------------------------------------------------------------------------------------
struct Value {
  int C, Cp, H, W;
  bool tan;
  long long slot_elems;  // B*H*W*Cp
  int nslots;
  long long act_off;   // float offset in the workspace
  long long grad_off;  // == act_off in GGN mode
};

struct Node {
  curv_node_desc d;
  // CONV
  Geom fwd, dgr;
  long long wk_off = -1, wt_off = -1, wkt_off = -1, wtt_off = -1;  // packed weight, transposed, tangents
  long long wsize = 0, wtsize = 0;
  // tcgen05 weight images (forward geometry: W, tangents; dgrad geometry: W^T, tangents)
  long long wimg_off = -1, wimgt_off = -1, wtimg_off = -1, wtimgt_off = -1;
  long long wimg_size = 0, wtimg_size = 0;
  long long bias_off = -1, biast_off = -1;
  int nsplit = 1, m_per_split = 0;
  int wbm = 64, wbn = 64;
  // AFFINE
  long long coef_off = -1, aux_off = -1;
  int rows_per_cta = 0, nchunks = 0;
  // MAXPOOL
  long long idx_off = -1;  // float offset
};

struct curv_program {
  int B, kmax, hessian;
  std::vector<Value> values;
  std::vector<Node> nodes;
  std::vector<curv_param_desc> params;
  long long scratch_off = 0, scratch_elems = 0;
  // half-split path (hs_gemm.cuh): plane scratch R1 (all slots of one tensor), R2 (one primal slot), both in
  // floats (hi + lo fp16 planes = 4 bytes per element), and the absmax bit patterns:
  //   value v, act slot s  -> bits[(2 v) (1+kmax) + s],  grad slot s -> bits[(2 v + 1) (1+kmax) + s]
  //   node i: 2 (1+kmax) words at bits[(2 nvalues + 2 i) (1+kmax)]: conv: weight 0 / tangent k -> word k;
  //           affine: coefficient maxima (affine_prep_kernel)
  long long hs1_off = 0, hs1_elems = 0, hs2_off = 0, hs2_elems = 0, hsbits_off = 0, hsbits_count = 0;
  size_t ws_bytes = 0;
};

static long long align_up(long long x, long long a) { return (x + a - 1) / a * a; }

// split count / scratch need of a Gram matrix X^T X (X: rows x widthp, `width` real columns); kfac.cuh
static int gram_nsplit(long long rows, int width, int widthp) {
  const int bm = width > 64 ? 128 : 64, bn = widthp > 64 ? 128 : 64;
  long long tiles = (long long)ceil_div(width, bm) * ceil_div(widthp, bn);
  long long want = (2 * 148 + tiles - 1) / tiles;
  long long maxsplit = (rows + 511) / 512;
  long long ns = want < 1 ? 1 : (want > maxsplit ? maxsplit : want);
  return (int)(ns > 32 ? 32 : (ns < 1 ? 1 : ns));
}
static long long gram_partial_elems(long long rows, int width, int widthp) {
  return (long long)gram_nsplit(rows, width, widthp) * width * widthp + 64;
}

static long long gram_hs_plan_elems(long long rows, int ld, int planes);  // kfac.cuh
static long long ekfac_scratch_elems(long long M, int B, int S, int d_in_width, int Cp_out);  // ekfac.cuh

extern "C" const char* curv_last_error(void) { return g_err.c_str(); }
extern "C" int curv_abi_version(void) { return CURV_ABI_VERSION; }
extern "C" long long curv_launch_count(void) { return g_launches; }
// a host mirror that replays a captured CUDA graph of n of this library's launches reports them here
extern "C" void curv_add_launch_count(long long n) { g_launches += n; }
extern "C" int curv_set_tensor_core_mode(int mode) {
  int old = g_tc_mode | (g_tc_disable << 4);
  g_tc_mode = mode & 3;
  g_tc_disable = (mode >> 4) & 0xFFF;  // 512 (mode bit 0x2000): tangent-weight images through the packed fp32 copy,
                                      // 1024 (0x4000): KFAC Gram matrices on the SIMT / 3xTF32 kernels
  return old;
}

// state that changes which kernels a call launches (host mirrors key their CUDA-graph caches on it):
// tensor-core mode word (as accepted by curv_set_tensor_core_mode) | profiling flag << 16
extern "C" int curv_launch_config(void);

extern "C" int curv_program_create(const curv_value_desc* values, int n_values,
                                   const curv_node_desc* nodes, int n_nodes,
                                   const curv_param_desc* params, int n_params, int batch, int kmax,
                                   int hessian, curv_program** out) {
  if (!values || !nodes || !out || n_values < 1 || n_nodes < 1 || batch < 1 || kmax < 1 || kmax > 32)
    return fail(CURV_ERR_INVALID, "curv_program_create: bad arguments (need 1 <= kmax <= 32)");
  auto* P = new curv_program();
  P->B = batch; P->kmax = kmax; P->hessian = hessian;
  P->params.assign(params, params + n_params);
  long long off = 0;  // in floats
  auto alloc = [&](long long elems) { long long o = off; off = align_up(off + elems, 64); return o; };
  for (int i = 0; i < n_values; ++i) {
    Value v;
    v.C = values[i].C; v.H = values[i].H; v.W = values[i].W; v.Cp = pad8(v.C);
    v.tan = values[i].has_tangent != 0;
    if (v.C < 1 || v.H < 1 || v.W < 1) { delete P; return fail(CURV_ERR_INVALID, "bad value shape"); }
    v.slot_elems = (long long)batch * v.H * v.W * v.Cp;
    v.nslots = v.tan ? 1 + kmax : 1;
    v.act_off = alloc(v.slot_elems * v.nslots);
    v.grad_off = ((hessian & 1) && v.tan) ? alloc(v.slot_elems * v.nslots) : v.act_off;
    P->values.push_back(v);
  }
  long long scratch = 64;
  for (int i = 0; i < n_nodes; ++i) {
    Node n;
    n.d = nodes[i];
    const curv_node_desc& d = n.d;
    auto bad_id = [&](int id) { return id < 0 || id >= n_values; };
    if (d.op != CURV_OP_INPUT && (bad_id(d.in0) || bad_id(d.out))) {
      delete P; return fail(CURV_ERR_INVALID, "node references an unknown value");
    }
    if (d.op == CURV_OP_CONV) {
      const Value& vi = P->values[d.in0];
      const Value& vo = P->values[d.out];
      Geom g;
      g.B = batch; g.Hs = vi.H; g.Ws = vi.W; g.Cs = vi.Cp; g.Hd = vo.H; g.Wd = vo.W;
      g.KH = d.kh; g.KW = d.kw; g.sh = d.sh; g.sw = d.sw; g.ph = d.ph; g.pw = d.pw; g.mode = 0;
      g.N = vo.C; g.Nd = vo.Cp; g.Kd = d.kh * d.kw * vi.Cp; g.M = batch * vo.H * vo.W;
      if ((vi.H + 2 * d.ph - d.kh) / d.sh + 1 != vo.H || (vi.W + 2 * d.pw - d.kw) / d.sw + 1 != vo.W) {
        delete P; return fail(CURV_ERR_INVALID, "conv geometry does not match value shapes");
      }
      n.fwd = g;
      Geom q = g;  // dgrad: source = grad of out, destination = grad of in
      q.Hs = vo.H; q.Ws = vo.W; q.Cs = vo.Cp; q.Hd = vi.H; q.Wd = vi.W; q.mode = 1;
      q.N = vi.C; q.Nd = vi.Cp; q.Kd = d.kh * d.kw * vo.Cp; q.M = batch * vi.H * vi.W;
      n.dgr = q;
      n.wsize = (long long)g.N * g.Kd;
      n.wtsize = (long long)q.N * q.Kd;
      n.wk_off = alloc(n.wsize);
      if (vi.tan) n.wt_off = alloc(n.wtsize);
      if (d.p0 >= 0) {
        n.wkt_off = alloc(n.wsize * kmax);
        if ((hessian & 1) && vi.tan) n.wtt_off = alloc(n.wtsize * kmax);
      }
      n.wimg_size = tc_image_elems(g.N, g.Nd, g.Kd);
      n.wtimg_size = tc_image_elems(q.N, q.Nd, q.Kd);
      if (n.wimg_size > 0) {
        n.wimg_off = alloc(n.wimg_size);
        if (vi.tan) n.wtimg_off = alloc(n.wtimg_size);
        if (d.p0 >= 0) {
          n.wimgt_off = alloc(n.wimg_size * kmax);
          if ((hessian & 1) && vi.tan) n.wtimgt_off = alloc(n.wtimg_size * kmax);
        }
      }
      if (d.p1 >= 0 || d.c1 >= 0) n.bias_off = alloc(vo.Cp);
      if (d.p1 >= 0) n.biast_off = alloc((long long)vo.Cp * kmax);
      // wgrad tiling / split plan
      if (d.p0 >= 0) {
        n.wbm = g.N > 64 ? 128 : 64;
        n.wbn = g.Kd > 64 ? 128 : 64;
        int nsl = kmax + ((hessian & 1) ? 1 : 0);
        // multi-slot tcgen05 wgrad: the slots live inside one tile
        long long tiles = (long long)ceil_div(g.Kd, 128) * ceil_div(vo.Cp, 64);
        // Split count of the pixel reduction.  Lower bound: at most 2048 pixels (384 MMAs per accumulator on the fp16
        // path) per split, so the tensor core's truncating accumulation stays below ~1e-5 and no tile needs an in-kernel
        // TMEM flush (measured, round 2: flushing in the kernel instead -- HSW_FLUSH = 128 stages, splits only to fill
        // the SMs -- halves the partial traffic but doubles the kernel time: the single-buffered accumulators stall the
        // MMA warp for every read-modify-write drain); bf16 operators (tolerance 1e-2) take chains of 4096 pixels =
        // the kernel's flush interval (16384-pixel chains were measured: 6.3 -> 9.5 ms of wgrad per bf16 C2 step,
        // the drains again).  Above the bound the count is chosen by a cost model: whole waves of
        // tiles x splits work items over the 148 SMs, an item costing its pixels plus a fixed prologue / epilogue
        // (~768 pixels' worth), plus the partial traffic of the split-K sums.
        {
          const int lower = std::max(1, ceil_div(g.M, (hessian & 4) ? 4096 : 2048));
          const int upper = std::max(lower, std::min(ceil_div(g.M, 512), 1024));
          const double px_ns = (hessian & 4) ? 14.0 * nsl : 42.0 * nsl / 4.0;  // per pixel of one work item
          const double split_ns = (double)n.wsize * nsl * 8.0 / 6.5e3;          // write + read of one split's partials
          double best = 1e300;
          n.nsplit = lower;
          for (int ns = lower; ns <= upper; ++ns) {
            const double waves = (double)((tiles * ns + 147) / 148);
            const double cost = waves * ((double)ceil_div(g.M, ns) + 768.0) * px_ns + ns * split_ns;
            if (cost < best * 0.999) { best = cost; n.nsplit = ns; }
          }
        }
        n.m_per_split = (ceil_div(g.M, n.nsplit) + 15) / 16 * 16;
        n.nsplit = ceil_div(g.M, n.m_per_split);
        long long need = (long long)n.nsplit * nsl * n.wsize;
        if (need > scratch) scratch = need;
      }
      if (hessian & 2) {  // KFAC: patch matrix + Gram partials (A factor), Gram partials of the cotangents (G)
        const int width = vi.C * d.kh * d.kw + 1, widthp = pad4(width);
        long long needA = align_up((long long)g.M * widthp, 64) + gram_partial_elems(g.M, width, widthp);
        long long needG = gram_partial_elems((long long)g.M * kmax, vo.C, vo.Cp);
        if (needA > scratch) scratch = needA;
        if (needG > scratch) scratch = needG;
        // tensor-core Gram path (kfac.cuh): patch planes (2 bytes per element and plane) + partials
        const int planes = (hessian & 4) ? 1 : 2, ld = pad8(width);
        long long needH = align_up((long long)g.M * ld + 1024, 128) * planes / 2 + gram_hs_plan_elems(g.M, ld, planes);
        long long needHG = gram_hs_plan_elems(g.M, vo.Cp, planes);
        if (needH > scratch) scratch = needH;
        if (needHG > scratch) scratch = needHG;
        if (!(hessian & 4)) {  // EKFAC eigenvalue correction (fp32 operators): rotated operands + per-example partials
          long long needE = ekfac_scratch_elems(g.M, batch, vo.H * vo.W, width, vo.Cp);
          if (needE > scratch) scratch = needE;
        }
      }
      if (d.p1 >= 0) {  // bias grad = column sum of the output cotangent
        long long rows = g.M;
        n.rows_per_cta = (int)((rows + 2 * 148 - 1) / (2 * 148));
        if (n.rows_per_cta < 64) n.rows_per_cta = 64;
        n.nchunks = (int)((rows + n.rows_per_cta - 1) / n.rows_per_cta);
        long long need = (long long)n.nchunks * (kmax + 1) * 2 * vo.Cp;
        if (need > scratch) scratch = need;
      }
    } else if (d.op == CURV_OP_AFFINE) {
      const Value& vi = P->values[d.in0];
      n.coef_off = alloc((long long)(1 + kmax) * 2 * vi.Cp);
      n.aux_off = alloc(2LL * vi.Cp);
      long long rows = (long long)batch * vi.H * vi.W;
      n.rows_per_cta = (int)((rows + 2 * 148 - 1) / (2 * 148));
      if (n.rows_per_cta < 64) n.rows_per_cta = 64;
      n.nchunks = (int)((rows + n.rows_per_cta - 1) / n.rows_per_cta);
      long long need = (long long)n.nchunks * (kmax + 1) * 2 * vi.Cp;
      if (need > scratch) scratch = need;
    } else if (d.op == CURV_OP_LAYERNORM) {
      const Value& vi = P->values[d.in0];
      if (hessian & 1) { delete P; return fail(CURV_ERR_UNSUPPORTED, "LayerNorm is not supported by the Hessian R-op program"); }
      if (vi.Cp > 1024) { delete P; return fail(CURV_ERR_UNSUPPORTED, "LayerNorm over more than 1024 channels"); }
      n.coef_off = alloc((long long)(1 + kmax) * 2 * vi.Cp);
      long long rows = (long long)batch * vi.H * vi.W;
      n.aux_off = alloc(2 * rows);
      n.rows_per_cta = (int)((rows + 2 * 148 - 1) / (2 * 148));
      if (n.rows_per_cta < 32) n.rows_per_cta = 32;
      n.nchunks = (int)((rows + n.rows_per_cta - 1) / n.rows_per_cta);
      long long need = (long long)n.nchunks * (kmax + 1) * 2 * vi.Cp;
      if (need > scratch) scratch = need;
    } else if (d.op == CURV_OP_RESHAPE || d.op == CURV_OP_CLSCAT || d.op == CURV_OP_POSADD || d.op == CURV_OP_TOKSEL) {
      Value& vi = P->values[d.in0];
      Value& vo = P->values[d.out];
      if (hessian & 1) { delete P; return fail(CURV_ERR_UNSUPPORTED, "token ops are not supported by the Hessian R-op program"); }
      if (vi.C != vi.Cp || vo.C != vi.C) { delete P; return fail(CURV_ERR_INVALID, "token ops need a channel count that is a multiple of 8"); }
      bool ok = true;
      if (d.op == CURV_OP_RESHAPE) {
        ok = vi.slot_elems == vo.slot_elems && vi.tan == vo.tan;
        if (ok) { vo.act_off = vi.act_off; vo.grad_off = vi.grad_off; }  // alias: same elements, other (H, W)
      } else if (d.op == CURV_OP_CLSCAT) {
        ok = vi.H == 1 && vo.H == 1 && vo.W == vi.W + 1 && (d.p0 >= 0 || d.c0 >= 0);
      } else if (d.op == CURV_OP_POSADD) {
        ok = vi.H == 1 && vo.H == 1 && vo.W == vi.W && (d.p0 >= 0 || d.c0 >= 0);
      } else {
        ok = vi.H == 1 && vo.H == 1 && vo.W == 1 && d.kw >= 0 && d.kw < vi.W;
      }
      if (!ok) { delete P; return fail(CURV_ERR_INVALID, "token op: value shapes do not match"); }
    } else if (d.op == CURV_OP_ATTENTION) {
      const Value& vi = P->values[d.in0];
      const Value& vo = P->values[d.out];
      if (hessian & 1) { delete P; return fail(CURV_ERR_UNSUPPORTED, "attention is not supported by the Hessian R-op program"); }
      if (d.kh < 1 || vo.C % d.kh != 0 || vi.C != 3 * vo.C || vi.H != 1 || vo.H != 1 || vi.W != vo.W ||
          vi.Cp != vi.C || vo.Cp != vo.C) {
        delete P; return fail(CURV_ERR_INVALID, "attention node: in0 must be [B, T, 3E], out [B, T, E], E a multiple of 8 and of the head count");
      }
      const long long pt = (long long)batch * d.kh * vi.W * ((vi.W + 3) & ~3);  // softmax matrices of one slot (rows padded to 16 bytes)
      n.aux_off = alloc(pt);
      if (pt * kmax > scratch) scratch = pt * kmax;
    } else if (d.op == CURV_OP_MAXPOOL) {
      const Value& vo = P->values[d.out];
      n.idx_off = alloc((vo.slot_elems + 3) / 4);  // float offset of a byte buffer (1 byte per element)
    } else if (d.op == CURV_OP_ADD) {
      if (bad_id(d.in1)) { delete P; return fail(CURV_ERR_INVALID, "add node needs two inputs"); }
    }
    if ((d.op == CURV_OP_ADD || d.op == CURV_OP_AFFINE) && d.kh == 2 && (hessian & 1)) {
      delete P; return fail(CURV_ERR_INVALID, "fused ReLU nodes are not supported by the Hessian R-op program");
    }
    P->nodes.push_back(n);
  }
  {
    const Value& last = P->values[P->nodes.back().d.out];
    if (last.H != 1 || last.W != 1) {
      delete P; return fail(CURV_ERR_UNSUPPORTED, "prediction must be [batch, C] (H = W = 1)");
    }
  }
  P->scratch_elems = scratch;
  P->scratch_off = alloc(scratch);
  {  // half-split planes for the sweeps (the Hessian R-op keeps K + 1 cotangent slots)
    for (const Node& n : P->nodes) {
      if (n.d.op != CURV_OP_CONV) continue;
      const Value& vi = P->values[n.d.in0];
      const Value& vo = P->values[n.d.out];
      if (!hs_gather_shape_ok(n.fwd)) continue;
      const long long fwd_need = vi.slot_elems * vi.nslots;
      const long long bwd_need = (vo.Cp % 8 == 0) ? vo.slot_elems * (kmax + ((hessian & 1) ? 1 : 0)) : 0;
      P->hs1_elems = std::max(P->hs1_elems, std::max(fwd_need, bwd_need));
      P->hs2_elems = std::max(P->hs2_elems, vi.slot_elems);
    }
    if (P->hs1_elems > 0) {
      P->hs1_off = alloc(P->hs1_elems);
      P->hs2_off = alloc(P->hs2_elems);
      P->hsbits_count = (long long)(2 * n_values + 2 * n_nodes) * (1 + kmax) + 16;  // + 16 words of per-launch scratch
      P->hsbits_off = alloc(P->hsbits_count);
    }
  }
  P->ws_bytes = (size_t)off * sizeof(float);
  *out = P;
  return CURV_OK;
}

extern "C" void curv_program_destroy(curv_program* p) { delete p; }
extern "C" size_t curv_program_workspace_bytes(const curv_program* p) { return p ? p->ws_bytes : 0; }
extern "C" int curv_program_value_layout(const curv_program* p, int id, size_t* byte_offset,
                                         size_t* slot_bytes, int* cp) {
  if (!p || id < 0 || id >= (int)p->values.size()) return fail(CURV_ERR_INVALID, "bad value id");
  const Value& v = p->values[id];
  if (byte_offset) *byte_offset = (size_t)v.act_off * 4;
  if (slot_bytes) *slot_bytes = (size_t)v.slot_elems * 4;
  if (cp) *cp = v.Cp;
  return CURV_OK;
}

// ------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------
// Optional per-launch timing of the contraction kernels with CUDA events on the launching stream
// (bench.py `roofline`): class 0 = gather GEMM (forward / dgrad), 1 = wgrad GEMM.
struct ProfRec { int cls; double flops; cudaEvent_t e0, e1; };
static int g_profile = 0;
static std::vector<ProfRec> g_prof;
struct ProfScope {
  ProfRec r; bool on; cudaStream_t st;
  ProfScope(int cls, double flops, cudaStream_t s) : on(g_profile != 0), st(s) {
    if (!on) return;
    r.cls = cls; r.flops = flops;
    cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, st);
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(r.e1, st);
    g_prof.push_back(r);
  }
};
extern "C" int curv_launch_config(void) { return g_tc_mode | (g_tc_disable << 4) | (g_profile ? (1 << 16) : 0); }
extern "C" int curv_profile_enable(int on) {
  for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  g_prof.clear();
  g_profile = on;
  return CURV_OK;
}
// sums over the launches recorded since curv_profile_enable(1): ms[c], flops[c], count[c] for c = 0, 1
extern "C" int curv_profile_read(double* ms, double* flops, long long* count) {
  for (int c = 0; c < 2; ++c) { ms[c] = 0; flops[c] = 0; count[c] = 0; }
  for (auto& r : g_prof) {
    if (cudaEventSynchronize(r.e1) != cudaSuccess) return fail(CURV_ERR_CUDA, "profile event sync failed");
    float t = 0.f;
    cudaEventElapsedTime(&t, r.e0, r.e1);
    if (r.cls > 1) continue;
    ms[r.cls] += t; flops[r.cls] += r.flops; count[r.cls] += 1;
  }
  return CURV_OK;
}
// same for one class: 0 gather GEMM, 1 wgrad GEMM, 2 half-split absmax/split passes
extern "C" int curv_profile_read_class(int cls, double* ms, double* flops, long long* count) {
  *ms = 0; *flops = 0; *count = 0;
  for (auto& r : g_prof) {
    if (r.cls != cls) continue;
    if (cudaEventSynchronize(r.e1) != cudaSuccess) return fail(CURV_ERR_CUDA, "profile event sync failed");
    float t = 0.f;
    cudaEventElapsedTime(&t, r.e0, r.e1);
    *ms += t; *flops += r.flops; *count += 1;
  }
  return CURV_OK;
}
// algorithmic FLOPs of one (slot, segment) contraction of a conv with forward geometry f
static double conv_flops(const Geom& f, int cin_real) {
  return 2.0 * (double)f.M * f.N * f.KH * f.KW * cin_real;
}

static int launch_gather_gemm(const GatherGemmArgs& a, int nslots, cudaStream_t st, double flops = 0) {
  if (nslots <= 0) return CURV_OK;
  const Geom& g = a.g;
  ProfScope prof(0, flops, st);
  if (g_tc_mode && a.W_img != nullptr) {
    GatherGemmArgs a2 = a;
    a2.debug = (g_tc_disable >> 3) & 3;
    int rc = tc_launch_gather_gemm(a2, nslots, st, (g_tc_disable & 4) == 0);
    if (rc == 0) { ++g_launches; return CURV_OK; }
    if (rc > 0) return fail(CURV_ERR_CUDA, "tcgen05 gather GEMM launch failed");
    // rc < 0: not eligible after all -> SIMT
  }
  if (g.M >= 4096 && g.Nd > 64) {
    dim3 grid(ceil_div(g.M, 128) * ceil_div(g.Nd, 128), nslots);
    gather_gemm_simt<128, 128><<<grid, 256, 0, st>>>(a);
  } else if (g.M >= 4096) {
    dim3 grid(ceil_div(g.M, 128) * ceil_div(g.Nd, 64), nslots);
    gather_gemm_simt<128, 64><<<grid, 256, 0, st>>>(a);
  } else {
    dim3 grid(ceil_div(g.M, 64) * ceil_div(g.Nd, 64), nslots);
    gather_gemm_simt<64, 64><<<grid, 256, 0, st>>>(a);
  }
  LAUNCH_CHECK();
  return CURV_OK;
}

static int launch_wgrad(const WgradArgs& a, int bm, int bn, cudaStream_t st, double flops = 0) {
  const Geom& g = a.g;
  ProfScope prof(1, flops, st);
  if (g_tc_mode && !(g_tc_disable & 2) && tc_wgrad_eligible(g, g_tc_mode)) {
    int rc = tc_launch_wgrad(a, st);
    if (rc == 0) { ++g_launches; return CURV_OK; }
    if (rc > 0) return fail(CURV_ERR_CUDA, "tcgen05 wgrad GEMM launch failed");
  }
  dim3 grid(ceil_div(g.N, bm) * ceil_div(g.Kd, bn), a.nslots, a.nsplit);
  if (bm == 128 && bn == 128) wgrad_gemm_simt<128, 128><<<grid, 256, 0, st>>>(a);
  else if (bm == 128) wgrad_gemm_simt<128, 64><<<grid, 256, 0, st>>>(a);
  else if (bn == 128) wgrad_gemm_simt<64, 128><<<grid, 256, 0, st>>>(a);
  else wgrad_gemm_simt<64, 64><<<grid, 256, 0, st>>>(a);
  LAUNCH_CHECK();
  return CURV_OK;
}

// split-K finish of a weight gradient: partial [split][slot][n][tap][cp] -> out rows of the parameter (K minor)
// nrows: rows per slot of the partials (0: N; larger when the slots were column blocks padded to the block width).
// The partials are consumed (many splits are first summed in place, chunk by chunk).
static int launch_wgrad_finish(float* partial, int nsplit, int ns, int kskip, int N, int C, int Cp, int taps,
                               float* out, long long off, int ldk, int k0, float alpha, long long wsize,
                               cudaStream_t st, int nrows = 0) {
  if (ns - kskip > 8) return fail(CURV_ERR_INVALID, "wgrad finish: more than 8 columns");
  const long long split_elems = (long long)ns * (nrows > 0 ? nrows : N) * taps * Cp;
  int step = 1;
  if (nsplit >= 12 && split_elems % 4 == 0) {
    int L = 2;
    while ((long long)L * L < nsplit) ++L;
    const int nch = ceil_div(nsplit, L);
    const int gx = std::max(1, std::min(grid1d(split_elems / 4), (148 * 16) / nch));
    wgrad_presum_kernel<<<dim3(gx, nch), 256, 0, st>>>(partial, nsplit, L, split_elems);
    LAUNCH_CHECK();
    step = L;
    nsplit = nch;
  }
  wgrad_finish_kernel<<<grid1d((long long)N * taps * Cp), 256, 0, st>>>(partial, nsplit, ns, kskip, N, C, Cp, taps, out,
                                                                        off, ldk, k0, alpha, nrows, step);
  LAUNCH_CHECK();
  return CURV_OK;
}

struct Ctx {
  curv_program* P;
  float* ws;
  const void* const* pp;  // differentiated parameter pointers
  const void* const* cp;  // constant pointers
  const float* V;
  float* out;
  int K, ldk, k0;
  float alpha;
  cudaStream_t st;
  int kind;
  bool rop;  // Hessian R-op: slot 0 of the cotangent storage holds the plain backward
  float** kfac_G = nullptr;  // KFAC mode: per node, the G factor to accumulate into (or null); no param grads
  float kfac_wG = 0.f;
  // EKFAC eigenvalue-correction mode (ekfac.cuh): per node the eigenvector matrices and the lambda to accumulate into
  const struct EkfacJob* ekfac = nullptr;
  // streaming calls (curv_matmat_batch_sync): per-parameter events.  v_ready[p]: the rows of V of parameter p are
  // on the device (waited for before the node that owns p is prepared, lazily, inside the forward sweep);
  // out_done[p]: recorded once the rows of `out` of parameter p are final (after the node's finish kernels).
  void* const* v_ready = nullptr;
  void* const* out_done = nullptr;
  // half-split path: enabled per call; hs_valid[e] = absmax entry e already computed in this call
  bool hs = false;
  int planes = 2;  // operand planes of the tcgen05 contraction kernels: 2 = fp16 hi/lo (fp32-grade), 1 = bf16 (program flag 4)
  std::vector<char>* hs_valid = nullptr;
  uint32_t* hsbits() const { return reinterpret_cast<uint32_t*>(ws + P->hsbits_off); }
  int bits_act(int v) const { return (2 * v) * (1 + P->kmax); }
  int bits_grad(int v) const { return (2 * v + 1) * (1 + P->kmax); }
  int bits_node(int ni) const { return (2 * (int)P->values.size() + 2 * ni) * (1 + P->kmax); }
  __half* hs1_hi() const { return reinterpret_cast<__half*>(ws + P->hs1_off); }
  __half* hs1_lo() const { return reinterpret_cast<__half*>(ws + P->hs1_off) + P->hs1_elems; }
  __half* hs2_hi() const { return reinterpret_cast<__half*>(ws + P->hs2_off); }
  __half* hs2_lo() const { return reinterpret_cast<__half*>(ws + P->hs2_off) + P->hs2_elems; }
  float* act(int v, int slot = 0) const {
    const Value& x = P->values[v];
    return ws + x.act_off + (long long)slot * x.slot_elems;
  }
  float* grad(int v, int slot = 0) const {
    const Value& x = P->values[v];
    return ws + x.grad_off + (long long)slot * x.slot_elems;
  }
  const float* param(int i) const { return (const float*)pp[i]; }
  const float* cst(int i) const { return (const float*)cp[i]; }
  const float* vcol(int p) const { return V + P->params[p].offset * ldk + k0; }
  float* ocol(int p) const { return out + P->params[p].offset * ldk + k0; }
};

// ---- half-split path: which contractions of a conv node run on the fp16 hi/lo kernels (hs_gemm.cuh)
// size thresholds of the half-split kernels: two full 128-row tiles are enough (a mini-batch sharded over 8 GPUs
// leaves 16 x 7 x 7 = 784 rows in the last ResNet stage); smaller problems (fc, MLPs) stay on the SIMT kernels
static bool hs_big_enough(const Geom& g) { return g_tc_mode >= 2 || (g.M >= 256 && g.Kd >= 64 && g.N >= 16); }
static bool hs_fwd_ok(const Ctx& c, const Node& n) {
  return c.hs && n.wimg_off >= 0 && hs_gather_shape_ok(n.fwd) && hs_big_enough(n.fwd);
}
static bool hs_dgr_ok(const Ctx& c, const Node& n) {
  return c.hs && n.wtimg_off >= 0 && hs_gather_shape_ok(n.dgr) && hs_big_enough(n.dgr);
}
static bool hs_wgr_ok(const Ctx& c, const Node& n, int ns) {
  const Value& vo = c.P->values[n.d.out];
  return c.hs && !(g_tc_disable & 2) && ns >= 1 && ns <= 8 && hs_wgrad_shape_ok(n.fwd, vo.Cp) &&
         hs_big_enough(n.fwd);
}
// absmax of `count` slots (x = first of them) into bits entries [entry, entry + count), once per call
static int hs_absmax(const Ctx& c, const float* x, long long slot_stride, long long n, int entry, int count) {
  bool need = false;
  for (int i = 0; i < count; ++i) need = need || !(*c.hs_valid)[entry + i];
  if (!need) return CURV_OK;
  if (c.planes == 1) {  // bf16 planes need no scale: an absmax word that nobody tracked stays 0 = scale 1
    for (int i = 0; i < count; ++i) (*c.hs_valid)[entry + i] = 1;
    return CURV_OK;
  }
  ProfScope prof(2, 0, c.st);
  if (hs_launch_absmax(x, slot_stride, n, c.hsbits() + entry, count, c.st))
    return fail(CURV_ERR_CUDA, "half-split absmax launch failed");
  ++g_launches;
  for (int i = 0; i < count; ++i) (*c.hs_valid)[entry + i] = 1;
  return CURV_OK;
}
// absmax entries [entry, entry + count) will be produced by the kernel about to be launched (fused tracking):
// returns the bits pointer to hand to it (nullptr when the half-split path is off)
static unsigned int* hs_fused_absmax(const Ctx& c, int entry, int count) {
  if (!c.hs) return nullptr;
  for (int i = 0; i < count; ++i) (*c.hs_valid)[entry + i] = 1;
  return c.hsbits() + entry;
}
static int hs_split(const Ctx& c, const float* x, long long slot_stride, long long n, __half* hi, __half* lo,
                    int entry, int count) {
  ProfScope prof(2, 0, c.st);
  if (hs_launch_split(x, slot_stride, n, hi, c.planes == 1 ? nullptr : lo, n, c.hsbits() + entry, count, c.st))
    return fail(CURV_ERR_CUDA, "half-split split launch failed");
  ++g_launches;
  return CURV_OK;
}

// pack parameters (and, if with_tangents, the K columns of V) into the engine's layouts
static int prepare_node(const Ctx& c, Node& n, bool with_tangents) {
  curv_program* P = c.P;
  cudaStream_t st = c.st;
  {
    const curv_node_desc& d = n.d;
    if (c.v_ready && with_tangents) {  // streaming call: the columns of V of this node's parameters must have landed
      for (int p : {d.p0, d.p1})
        if ((d.op == CURV_OP_CONV || d.op == CURV_OP_AFFINE || d.op == CURV_OP_LAYERNORM || d.op == CURV_OP_CLSCAT ||
             d.op == CURV_OP_POSADD) && p >= 0 && c.v_ready[p])
          CHECK_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)c.v_ready[p], 0));
    }
    if (d.op == CURV_OP_CONV) {
      const Geom& g = n.fwd;
      const Value& vi = P->values[d.in0];
      const float* w = d.p0 >= 0 ? c.param(d.p0) : c.cst(d.c0);
      int Cin = vi.C;
      pack_weight_kernel<<<dim3(grid1d(n.wsize), 1), 256, 0, st>>>(w, 1, c.ws + n.wk_off, 0, g.N, Cin, g.KH,
                                                                  g.KW, vi.Cp, g.Nd, 0);
      LAUNCH_CHECK();
      if (n.wt_off >= 0) {
        pack_weight_kernel<<<dim3(grid1d(n.wtsize), 1), 256, 0, st>>>(w, 1, c.ws + n.wt_off, 0, g.N, Cin,
                                                                     g.KH, g.KW, vi.Cp, g.Nd, 1);
        LAUNCH_CHECK();
      }
      // the half-split / bf16 forward reads the tangent weights only as images, which are built straight from the
      // columns of V below: no packed fp32 copy of the K tangent weights
      const bool images_from_v = hs_fwd_ok(c, n) && c.K <= 8 && !(g_tc_disable & 512);
      if (with_tangents && d.p0 >= 0 && !images_from_v) {
        pack_weight_kernel<<<dim3(grid1d(n.wsize), c.K), 256, 0, st>>>(
            c.vcol(d.p0), c.ldk, c.ws + n.wkt_off, n.wsize, g.N, Cin, g.KH, g.KW, vi.Cp, g.Nd, 0);
        LAUNCH_CHECK();
      }
      if (with_tangents && d.p0 >= 0 && n.wtt_off >= 0 && c.rop) {  // R-op: transposed tangent weights (dgrad term)
        pack_weight_kernel<<<dim3(grid1d(n.wtsize), c.K), 256, 0, st>>>(
            c.vcol(d.p0), c.ldk, c.ws + n.wtt_off, n.wtsize, g.N, Cin, g.KH, g.KW, vi.Cp, g.Nd, 1);
        LAUNCH_CHECK();
      }
      const int ni = (int)(&n - P->nodes.data());
      const bool hsf = hs_fwd_ok(c, n), hsd = vi.tan && hs_dgr_ok(c, n);
      if (hsf || hsd) {  // fp16 hi/lo weight images (scale = one power of two per weight tensor / tangent)
        const int e0 = c.bits_node(ni);
        const bool wt = with_tangents && d.p0 >= 0;
        int rc = hs_absmax(c, c.ws + n.wk_off, 0, n.wsize, e0, 1);
        if (!rc && wt && hsf && !images_from_v) rc = hs_absmax(c, c.ws + n.wkt_off, n.wsize, n.wsize, e0 + 1, c.K);
        if (rc) return rc;
        int bad = 0;
        if (hsf) {
          bad |= hs_launch_pack_image(c.ws + n.wk_off, 0, reinterpret_cast<__half*>(c.ws + n.wimg_off), 0, g.N,
                                      g.Nd, g.Kd, 1, c.hsbits() + e0, st, c.planes);
          ++g_launches;
          if (wt && images_from_v) {
            bad |= hs_launch_pack_image_cols(c.vcol(d.p0), c.ldk, c.K, reinterpret_cast<__half*>(c.ws + n.wimgt_off),
                                             hs_image_halves(g.Nd, g.Kd, c.planes), g.N, Cin, g.KH * g.KW, vi.Cp,
                                             g.Nd, c.hsbits() + e0 + 1, st, c.planes);
            g_launches += c.planes == 1 ? 1 : 2;
            for (int k = 0; k < c.K; ++k) (*c.hs_valid)[e0 + 1 + k] = 1;
          } else if (wt) {
            bad |= hs_launch_pack_image(c.ws + n.wkt_off, n.wsize, reinterpret_cast<__half*>(c.ws + n.wimgt_off),
                                        hs_image_halves(g.Nd, g.Kd, c.planes), g.N, g.Nd, g.Kd, c.K,
                                        c.hsbits() + e0 + 1, st, c.planes);
            ++g_launches;
          }
        }
        if (hsd) {
          const Geom& q = n.dgr;
          bad |= hs_launch_pack_image(c.ws + n.wt_off, 0, reinterpret_cast<__half*>(c.ws + n.wtimg_off), 0, q.N,
                                      q.Nd, q.Kd, 1, c.hsbits() + e0, st, c.planes);
          ++g_launches;
          if (c.rop && wt && n.wtimgt_off >= 0 && n.wtt_off >= 0) {  // R-op dgrad term  dW_k^T . delta_0
            if (!images_from_v || !hsf) rc = hs_absmax(c, c.ws + n.wtt_off, n.wtsize, n.wtsize, e0 + 1, c.K);
            if (rc) return rc;
            bad |= hs_launch_pack_image(c.ws + n.wtt_off, n.wtsize, reinterpret_cast<__half*>(c.ws + n.wtimgt_off),
                                        hs_image_halves(q.Nd, q.Kd, c.planes), q.N, q.Nd, q.Kd, c.K,
                                        c.hsbits() + e0 + 1, st, c.planes);
            ++g_launches;
          }
        }
        if (bad) return fail(CURV_ERR_CUDA, "half-split weight image packing failed");
      }
      if (g_tc_mode && n.wimg_off >= 0 && !(hsf && (hsd || !vi.tan))) {  // 3xTF32 tcgen05 weight images
        const Geom& q = n.dgr;
        int bad = 0;
        if (!hsf && tc_gather_eligible(g, g_tc_mode)) {
          bad |= tc_pack_image(c.ws + n.wk_off, 0, c.ws + n.wimg_off, 0, g.N, g.Nd, g.Kd, 1, st) > 0;
          ++g_launches;
          if (with_tangents && d.p0 >= 0) {
            bad |= tc_pack_image(c.ws + n.wkt_off, n.wsize, c.ws + n.wimgt_off, n.wimg_size, g.N, g.Nd, g.Kd,
                                 c.K, st) > 0;
            ++g_launches;
          }
        }
        if (!hsd && n.wtimg_off >= 0 && tc_gather_eligible(q, g_tc_mode)) {
          bad |= tc_pack_image(c.ws + n.wt_off, 0, c.ws + n.wtimg_off, 0, q.N, q.Nd, q.Kd, 1, st) > 0;
          ++g_launches;
          if (with_tangents && n.wtimgt_off >= 0) {
            bad |= tc_pack_image(c.ws + n.wtt_off, n.wtsize, c.ws + n.wtimgt_off, n.wtimg_size, q.N, q.Nd,
                                 q.Kd, c.K, st) > 0;
            ++g_launches;
          }
        }
        if (bad) return fail(CURV_ERR_CUDA, "tcgen05 weight image packing failed");
      }
      if (n.bias_off >= 0) {
        const float* b = d.p1 >= 0 ? c.param(d.p1) : c.cst(d.c1);
        pack_vec_kernel<<<dim3(grid1d(g.Nd), 1), 256, 0, st>>>(b, 1, c.ws + n.bias_off, 0, g.N, g.Nd);
        LAUNCH_CHECK();
      }
      if (with_tangents && n.biast_off >= 0) {
        pack_vec_kernel<<<dim3(grid1d(g.Nd), c.K), 256, 0, st>>>(c.vcol(d.p1), c.ldk, c.ws + n.biast_off,
                                                                g.Nd, g.N, g.Nd);
        LAUNCH_CHECK();
      }
    } else if (d.op == CURV_OP_LAYERNORM) {  // coef slot 0 = (gamma, beta) [defaults 1, 0], slot k = their tangents
      const Value& vi = P->values[d.in0];
      const float* gamma = d.p0 >= 0 ? c.param(d.p0) : (d.c0 >= 0 ? c.cst(d.c0) : nullptr);
      const float* beta = d.p1 >= 0 ? c.param(d.p1) : (d.c1 >= 0 ? c.cst(d.c1) : nullptr);
      float* coef = c.ws + n.coef_off;
      fill_kernel<<<grid1d(vi.Cp), 256, 0, st>>>(coef, vi.C, 1.f, vi.Cp);  // gamma default 1 (pad lanes 0)
      LAUNCH_CHECK();
      CHECK_CUDA(cudaMemsetAsync(coef + vi.Cp, 0, sizeof(float) * vi.Cp * (1 + 2 * (size_t)(with_tangents ? c.K : 0)), st));
      if (gamma) { pack_vec_kernel<<<dim3(grid1d(vi.Cp), 1), 256, 0, st>>>(gamma, 1, coef, 0, vi.C, vi.Cp); LAUNCH_CHECK(); }
      if (beta) { pack_vec_kernel<<<dim3(grid1d(vi.Cp), 1), 256, 0, st>>>(beta, 1, coef + vi.Cp, 0, vi.C, vi.Cp); LAUNCH_CHECK(); }
      if (with_tangents && d.p0 >= 0) {
        pack_vec_kernel<<<dim3(grid1d(vi.Cp), c.K), 256, 0, st>>>(c.vcol(d.p0), c.ldk, coef + 2 * vi.Cp, 2 * vi.Cp, vi.C, vi.Cp);
        LAUNCH_CHECK();
      }
      if (with_tangents && d.p1 >= 0) {
        pack_vec_kernel<<<dim3(grid1d(vi.Cp), c.K), 256, 0, st>>>(c.vcol(d.p1), c.ldk, coef + 3 * vi.Cp, 2 * vi.Cp, vi.C, vi.Cp);
        LAUNCH_CHECK();
      }
    } else if (d.op == CURV_OP_AFFINE) {
      const Value& vi = P->values[d.in0];
      const float* gamma = d.p0 >= 0 ? c.param(d.p0) : (d.c0 >= 0 ? c.cst(d.c0) : nullptr);
      const float* beta = d.p1 >= 0 ? c.param(d.p1) : (d.c1 >= 0 ? c.cst(d.c1) : nullptr);
      const float* gd = (with_tangents && d.p0 >= 0) ? c.vcol(d.p0) : nullptr;
      const float* bd = (with_tangents && d.p1 >= 0) ? c.vcol(d.p1) : nullptr;
      affine_prep_kernel<<<grid1d(vi.Cp), 256, 0, st>>>(
          gamma, beta, c.cst(d.c2), c.cst(d.c3), d.eps, gd, bd, c.ldk, with_tangents ? c.K : 0, vi.C, vi.Cp,
          c.ws + n.coef_off, c.ws + n.aux_off,
          c.hs ? c.hsbits() + c.bits_node((int)(&n - P->nodes.data())) : nullptr);  // max_c |s_c| (AFFINE nodes own word 0)
      LAUNCH_CHECK();
    }
  }
  return CURV_OK;
}
// all nodes up front (default), or lazily node by node inside the forward sweep (streaming calls, so that the
// upload of later parameters' columns overlaps the first layers)
static int prepare_params(const Ctx& c, bool with_tangents) {
  if (c.v_ready) return CURV_OK;  // lazily, in forward()
  for (Node& n : c.P->nodes) {
    int rc = prepare_node(c, n, with_tangents);
    if (rc) return rc;
  }
  return CURV_OK;
}
// parameter rows of `out` owned by node n are final: signal the streaming caller
static int signal_out_done(const Ctx& c, const Node& n) {
  if (!c.out_done) return CURV_OK;
  for (int p : {n.d.p0, n.d.p1})
    if ((n.d.op == CURV_OP_CONV || n.d.op == CURV_OP_AFFINE || n.d.op == CURV_OP_LAYERNORM || n.d.op == CURV_OP_CLSCAT ||
         n.d.op == CURV_OP_POSADD) && p >= 0 && c.out_done[p])
      CHECK_CUDA(cudaEventRecord((cudaEvent_t)c.out_done[p], c.st));
  return CURV_OK;
}

// ---- multi-head attention core (attention.cuh): batched over (slot, example, head) with three-level strides
static bool g_attn_mma = false;  // set per call: bf16 operators multiply on the tensor cores (mma.sync, bf16 operands)
static int attn_gemm(cudaStream_t st, Bgemm3 p, int n0) {
  if (n0 < 1) return CURV_OK;
  const long long nz = (long long)n0 * p.n1 * p.n2;
  if (nz > 65535) return fail(CURV_ERR_UNSUPPORTED, "attention: more than 65535 (slot, example, head) triples per launch");
  constexpr size_t SMEM_MAX = 227 * 1024;
  if (p.A2 && !(g_attn_mma && attn_sgemm_smem(p.M, p.N, p.Kd, p.Kd2) <= SMEM_MAX)) {
    // the fused two-term kernel does not apply: two launches, the second accumulating
    Bgemm3 a = p, b = p;
    a.A2 = nullptr; a.B2 = nullptr;
    b.A = p.A2; b.lda = p.lda2; b.transA = p.transA2; b.B = p.B2; b.ldb = p.ldb2; b.transB = p.transB2; b.Kd = p.Kd2;
    for (int i = 0; i < 3; ++i) { b.sA[i] = p.sA2[i]; b.sB[i] = p.sB2[i]; }
    b.A2 = nullptr; b.B2 = nullptr; b.beta = 1.f;
    int rc = attn_gemm(st, a, n0);
    return rc ? rc : attn_gemm(st, b, n0);
  }
  const dim3 grid(ceil_div(p.N, 64), ceil_div(p.M, 64), (unsigned)nz);
  const size_t smem = attn_sgemm_smem(p.M, p.N, p.Kd, p.A2 ? p.Kd2 : 0);
  if (g_attn_mma && smem <= SMEM_MAX) {  // whole (slot, example, head) product per CTA
    static bool attr_set = false;
    if (!attr_set) {
      CHECK_CUDA(cudaFuncSetAttribute(attn_sgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
      attr_set = true;
    }
    attn_sgemm_kernel<<<(unsigned)nz, ATTN_SG_THREADS, smem, st>>>(p);
    LAUNCH_CHECK();
    return CURV_OK;
  }
  if (g_attn_mma) attn_bgemm_kernel<true><<<grid, 256, 0, st>>>(p);
  else attn_bgemm_kernel<false><<<grid, 256, 0, st>>>(p);
  LAUNCH_CHECK();
  return CURV_OK;
}
struct AttnDims {
  int T, Tp, E, H, dh, B;  // Tp: row pitch of the softmax-shaped buffers
  long long lin, lout, pt;  // slot strides of in / out, softmax elements per slot
  float scale;
};
static AttnDims attn_dims(const Ctx& c, const Node& n) {
  const Value& vi = c.P->values[n.d.in0];
  const Value& vo = c.P->values[n.d.out];
  AttnDims a;
  a.T = vi.W; a.Tp = (a.T + 3) & ~3; a.E = vo.C; a.H = n.d.kh; a.dh = a.E / a.H; a.B = c.P->B;
  a.lin = vi.slot_elems; a.lout = vo.slot_elems; a.pt = (long long)a.B * a.H * a.T * a.Tp;
  a.scale = 1.f / sqrtf((float)a.dh);
  return a;
}
// operand descriptors: activations [slot][b][t][3E or E] (head h at column offset col0 + h * dh), softmax-shaped
// buffers [slot][b][h][T][T]
static void attn_act(const AttnDims& a, const float* base, int ld, long long slot_stride, int col0, const float*& ptr,
                     int& ldo, long long s[3]) {
  ptr = base + col0; ldo = ld; s[0] = slot_stride; s[1] = (long long)a.T * ld; s[2] = a.dh;
}
static void attn_sm(const AttnDims& a, const float* base, bool per_slot, const float*& ptr, int& ldo, long long s[3]) {
  ptr = base; ldo = a.Tp; s[0] = per_slot ? a.pt : 0; s[1] = (long long)a.H * a.T * a.Tp; s[2] = (long long)a.T * a.Tp;
}
static int attention_forward(const Ctx& c, const Node& n, int K) {
  cudaStream_t st = c.st;
  g_attn_mma = (c.P->hessian & 4) && g_tc_mode && !(g_tc_disable & 32);
  const AttnDims a = attn_dims(c, n);
  const int E = a.E, T = a.T, ldi = 3 * E;
  const float* in = c.act(n.d.in0);
  float* out = c.act(n.d.out);
  float* Pm = c.ws + n.aux_off;
  float* D = c.ws + c.P->scratch_off;
  int rc;
  Bgemm3 g;
  memset(&g, 0, sizeof(g));
  g.n1 = a.B; g.n2 = a.H;
  // S = scale Q K^T
  g.transA = 0; g.transB = 1; g.M = T; g.N = T; g.Kd = a.dh; g.alpha = a.scale; g.beta = 0.f;
  attn_act(a, in, ldi, 0, 0, g.A, g.lda, g.sA);
  attn_act(a, in, ldi, 0, E, g.B, g.ldb, g.sB);
  { const float* q; attn_sm(a, Pm, false, q, g.ldc, g.sC); g.C = Pm; }
  if ((rc = attn_gemm(st, g, 1))) return rc;
  attn_softmax_kernel<<<(unsigned)(((long long)a.B * a.H * T + 7) / 8), 256, 0, st>>>(Pm, (long long)a.B * a.H * T, T, a.Tp);
  LAUNCH_CHECK();
  // O = P V
  g.transA = 0; g.transB = 0; g.M = T; g.N = a.dh; g.Kd = T; g.alpha = 1.f; g.beta = 0.f;
  attn_sm(a, Pm, false, g.A, g.lda, g.sA);
  attn_act(a, in, ldi, 0, 2 * E, g.B, g.ldb, g.sB);
  { const float* q; attn_act(a, out, E, 0, 0, q, g.ldc, g.sC); g.C = out; }
  if ((rc = attn_gemm(st, g, 1))) return rc;
  if (K < 1) return CURV_OK;
  if (a.pt * K > c.P->scratch_elems) return fail(CURV_ERR_WORKSPACE, "attention scratch too small");
  const float* tin = in + a.lin;    // tangent slots 1..K
  float* tout = out + a.lout;
  // dS = scale (dQ K^T + Q dK^T)
  g.transA = 0; g.transB = 1; g.M = T; g.N = T; g.Kd = a.dh; g.alpha = a.scale; g.beta = 0.f;
  attn_act(a, tin, ldi, a.lin, 0, g.A, g.lda, g.sA);
  attn_act(a, in, ldi, 0, E, g.B, g.ldb, g.sB);
  { const float* q; attn_sm(a, D, true, q, g.ldc, g.sC); g.C = D; }
  g.transA2 = 0; g.transB2 = 1; g.Kd2 = a.dh;                       // + Q dK^T in the same pass
  attn_act(a, in, ldi, 0, 0, g.A2, g.lda2, g.sA2);
  attn_act(a, tin, ldi, a.lin, E, g.B2, g.ldb2, g.sB2);
  if ((rc = attn_gemm(st, g, K))) return rc;
  const long long rows = (long long)K * a.B * a.H * T;
  attn_softmax_jvp_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(Pm, D, (long long)a.B * a.H * T, rows, T, a.Tp);
  LAUNCH_CHECK();
  // dO = dP V + P dV
  g.transA = 0; g.transB = 0; g.M = T; g.N = a.dh; g.Kd = T; g.alpha = 1.f; g.beta = 0.f;
  attn_sm(a, D, true, g.A, g.lda, g.sA);
  attn_act(a, in, ldi, 0, 2 * E, g.B, g.ldb, g.sB);
  { const float* q; attn_act(a, tout, E, a.lout, 0, q, g.ldc, g.sC); g.C = tout; }
  g.transA2 = 0; g.transB2 = 0; g.Kd2 = T;                          // + P dV in the same pass
  attn_sm(a, Pm, false, g.A2, g.lda2, g.sA2);
  attn_act(a, tin, ldi, a.lin, 2 * E, g.B2, g.ldb2, g.sB2);
  return attn_gemm(st, g, K);
}
// cotangent slots [s0, s0 + ns) of out -> the same slots of in (accumulated if `accumulate`)
static int attention_backward(const Ctx& c, const Node& n, int s0, int ns, int accumulate) {
  cudaStream_t st = c.st;
  g_attn_mma = (c.P->hessian & 4) && g_tc_mode && !(g_tc_disable & 32);
  const AttnDims a = attn_dims(c, n);
  const int E = a.E, T = a.T, ldi = 3 * E;
  const float* in = c.act(n.d.in0);            // primal q | k | v
  const float* go = c.grad(n.d.out, s0);
  float* gi = c.grad(n.d.in0, s0);
  float* Pm = c.ws + n.aux_off;
  float* D = c.ws + c.P->scratch_off;
  if (a.pt * ns > c.P->scratch_elems) return fail(CURV_ERR_WORKSPACE, "attention scratch too small");
  const float beta = accumulate ? 1.f : 0.f;
  int rc;
  Bgemm3 g;
  memset(&g, 0, sizeof(g));
  g.n1 = a.B; g.n2 = a.H;
  // gP = gO V^T
  g.transA = 0; g.transB = 1; g.M = T; g.N = T; g.Kd = a.dh; g.alpha = 1.f; g.beta = 0.f;
  attn_act(a, go, E, a.lout, 0, g.A, g.lda, g.sA);
  attn_act(a, in, ldi, 0, 2 * E, g.B, g.ldb, g.sB);
  { const float* q; attn_sm(a, D, true, q, g.ldc, g.sC); g.C = D; }
  if ((rc = attn_gemm(st, g, ns))) return rc;
  // gV = P^T gO
  g.transA = 1; g.transB = 0; g.M = T; g.N = a.dh; g.Kd = T; g.alpha = 1.f; g.beta = beta;
  attn_sm(a, Pm, false, g.A, g.lda, g.sA);
  attn_act(a, go, E, a.lout, 0, g.B, g.ldb, g.sB);
  { const float* q; attn_act(a, gi, ldi, a.lin, 2 * E, q, g.ldc, g.sC); g.C = gi + 2 * E; }
  if ((rc = attn_gemm(st, g, ns))) return rc;
  // gS = P o (gP - rowsum(P o gP))
  const long long rows = (long long)ns * a.B * a.H * T;
  attn_softmax_jvp_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(Pm, D, (long long)a.B * a.H * T, rows, T, a.Tp);
  LAUNCH_CHECK();
  // gQ = scale gS K
  g.transA = 0; g.transB = 0; g.M = T; g.N = a.dh; g.Kd = T; g.alpha = a.scale; g.beta = beta;
  attn_sm(a, D, true, g.A, g.lda, g.sA);
  attn_act(a, in, ldi, 0, E, g.B, g.ldb, g.sB);
  { const float* q; attn_act(a, gi, ldi, a.lin, 0, q, g.ldc, g.sC); g.C = gi; }
  if ((rc = attn_gemm(st, g, ns))) return rc;
  // gK = scale gS^T Q
  g.transA = 1;
  attn_sm(a, D, true, g.A, g.lda, g.sA);
  attn_act(a, in, ldi, 0, 0, g.B, g.ldb, g.sB);
  { const float* q; attn_act(a, gi, ldi, a.lin, E, q, g.ldc, g.sC); g.C = gi + E; }
  return attn_gemm(st, g, ns);
}

// forward sweep: primal only (K = 0) or primal + K tangents
static int forward(const Ctx& c, const void* X, int K) {
  curv_program* P = c.P;
  cudaStream_t st = c.st;
  int planes_of = -1;  // value whose slots currently sit in the hs1 planes (written by affine_fwd, planes mode)
  // consumers per value: a BN output read only by the next convolution needs no fp32 tangent slots
  std::vector<int> nuse(P->values.size(), 0);
  for (const Node& q : P->nodes) {
    if (q.d.op == CURV_OP_INPUT) continue;
    ++nuse[q.d.in0];
    if (q.d.op == CURV_OP_ADD) ++nuse[q.d.in1];
  }
  for (Node& n : P->nodes) {
    const curv_node_desc& d = n.d;
    if (d.op == CURV_OP_INPUT) {
      const Value& v = P->values[d.out];
      nchw_to_nhwc_kernel<<<grid1d(v.slot_elems), 256, 0, st>>>((const float*)X, c.act(d.out), P->B, v.C,
                                                                v.H, v.W, v.Cp);
      LAUNCH_CHECK();
      continue;
    }
    const Value& vi = P->values[d.in0];
    const Value& vo = P->values[d.out];
    const int nsl = (vo.tan && K > 0) ? 1 + K : 1;
    if (c.v_ready) {
      int rc = prepare_node(c, n, K > 0);
      if (rc) return rc;
    }
    switch (d.op) {
      case CURV_OP_CONV: {
        double fl = conv_flops(n.fwd, vi.C) *
                    (1 + (nsl - 1) * ((vi.tan ? 1 : 0) + (d.p0 >= 0 ? 1 : 0)));
        if (hs_fwd_ok(c, n)) {  // fp16 hi/lo planes of the input slots, then the half-split gather GEMM
          const int nin = (vi.tan && K > 0) ? 1 + K : 1;
          const int ea = c.bits_act(d.in0);
          if (planes_of != d.in0) {  // (else the producing BN kernel already wrote them)
            int rc = hs_absmax(c, c.act(d.in0), vi.slot_elems, vi.slot_elems, ea, nin);
            if (!rc) rc = hs_split(c, c.act(d.in0), vi.slot_elems, vi.slot_elems, c.hs1_hi(), c.hs1_lo(), ea, nin);
            if (rc) return rc;
          }
          planes_of = -1;
          if (!vi.tan && nsl > 1 && d.p0 >= 0 && !(g_tc_disable & 128)) {
            // the layer input carries no tangent (the stem): every slot gathers the same operand -> N-stacked
            // kernel, the gather is staged once per group of 256 / BN slots
            HsStackArgs q;
            memset(&q, 0, sizeof(q));
            q.g = n.fwd;
            q.Ah = c.hs1_hi(); q.Al = c.hs1_lo(); q.a_bits = c.hsbits() + ea;
            q.W_img = reinterpret_cast<const __half*>(c.ws + n.wimg_off);
            q.Wt_img = reinterpret_cast<const __half*>(c.ws + n.wimgt_off);
            q.Wt_img_slot = hs_image_halves(n.fwd.Nd, n.fwd.Kd, c.planes);
            q.planes = c.planes;
            q.w_bits = c.hsbits() + c.bits_node((int)(&n - P->nodes.data()));
            q.bias = n.bias_off >= 0 ? c.ws + n.bias_off : nullptr;
            q.bias_t = n.biast_off >= 0 ? c.ws + n.biast_off : nullptr; q.bias_slot = vo.Cp;
            q.out = c.act(d.out); q.out_slot = vo.slot_elems; q.slot_lo = 0; q.nslots = nsl; q.accumulate = 0;
            ProfScope prof(0, fl, st);
            if (hs_launch_gather_stack(q, st)) return fail(CURV_ERR_CUDA, "half-split stacked gather GEMM launch failed");
            ++g_launches;
            break;
          }
          HsGatherArgs h;
          memset(&h, 0, sizeof(h));
          h.g = n.fwd;
          h.Ah = c.hs1_hi(); h.Al = c.hs1_lo(); h.A_slot = vi.slot_elems; h.a_slot_base = 0;
          h.a_has_slots = vi.tan ? 1 : 0; h.a_bits = c.hsbits() + ea;
          h.W_img = reinterpret_cast<const __half*>(c.ws + n.wimg_off);
          h.Wt_img = (d.p0 >= 0 && K > 0) ? reinterpret_cast<const __half*>(c.ws + n.wimgt_off) : nullptr;
          h.Wt_img_slot = hs_image_halves(n.fwd.Nd, n.fwd.Kd, c.planes);
          h.planes = c.planes;
          h.w_bits = c.hsbits() + c.bits_node((int)(&n - P->nodes.data()));
          h.bias = n.bias_off >= 0 ? c.ws + n.bias_off : nullptr;
          h.bias_t = n.biast_off >= 0 ? c.ws + n.biast_off : nullptr; h.bias_slot = vo.Cp;
          h.out = c.act(d.out); h.out_slot = vo.slot_elems; h.slot0 = 0; h.accumulate = 0;
          h.out_bits = hs_fused_absmax(c, c.bits_act(d.out), nsl);  // absmax of the conv output (bounds downstream)
          ProfScope prof(0, fl, st);
          if (hs_launch_gather_gemm(h, nsl, st, true, nin)) return fail(CURV_ERR_CUDA, "half-split gather GEMM launch failed");
          ++g_launches;
          break;
        }
        GatherGemmArgs a;
        memset(&a, 0, sizeof(a));
        a.g = n.fwd;
        a.A = c.act(d.in0); a.A_slot = vi.slot_elems; a.a_has_slots = vi.tan ? 1 : 0;
        a.W = c.ws + n.wk_off;
        a.Wt = (d.p0 >= 0) ? c.ws + n.wkt_off : nullptr; a.Wt_slot = n.wsize;
        if (g_tc_mode && !(g_tc_disable & 1) && n.wimg_off >= 0 && tc_gather_eligible(n.fwd, g_tc_mode)) {
          a.W_img = c.ws + n.wimg_off;
          a.Wt_img = (d.p0 >= 0 && K > 0) ? c.ws + n.wimgt_off : nullptr; a.Wt_img_slot = n.wimg_size;
        }
        a.bias = n.bias_off >= 0 ? c.ws + n.bias_off : nullptr;
        a.bias_t = n.biast_off >= 0 ? c.ws + n.biast_off : nullptr; a.bias_slot = vo.Cp;
        a.out = c.act(d.out); a.out_slot = vo.slot_elems;
        a.slot0 = 0; a.accumulate = 0;
        int rc = launch_gather_gemm(a, nsl, st, fl);
        if (rc) return rc;
        break;
      }
      case CURV_OP_AFFINE: {
        long long rows = (long long)P->B * vi.H * vi.W;
        // Planes mode: the only reader of this output is the convolution that comes next and it runs on the
        // half-split kernels -> write its fp16 hi/lo operand planes here (scale from a bound, see the kernel) and
        // skip the fp32 tangent slots and the separate split pass.  (GGN-type sweeps only: the cotangent slots
        // of the backward sweep reuse the tangent storage, nothing reads the fp32 tangents later.)
        const int ni = (int)(&n - P->nodes.data());
        bool planes = false;
        if (c.hs && K > 0 && !c.rop && !(g_tc_disable & 64) && ni + 1 < (int)P->nodes.size() && nuse[d.out] == 1 &&
            (c.kind == CURV_KIND_GGN || c.kind == CURV_KIND_GGN_MC)) {
          const Node& cn = P->nodes[ni + 1];
          planes = cn.d.op == CURV_OP_CONV && cn.d.in0 == d.out && hs_fwd_ok(c, cn) &&
                   vo.slot_elems * nsl <= P->hs1_elems && nsl == 1 + K;
        }
        if (planes) {  // the bound needs the absmax of the input slots (fused by the producing conv, else a pass)
          int rc = hs_absmax(c, c.act(d.in0), vi.slot_elems, vi.slot_elems, c.bits_act(d.in0), vi.tan ? nsl : 1);
          if (rc) return rc;
        }
        {
          const dim3 grid(grid1d(rows * (vi.Cp / 4)), nsl > 1 ? (nsl + 6) / 8 : 1);
          unsigned int* am = hs_fused_absmax(c, c.bits_act(d.out), nsl);
          if (planes)
            affine_fwd_kernel<true><<<grid, 256, 0, st>>>(
                c.act(d.in0), vi.slot_elems, vi.tan ? 1 : 0, c.ws + n.coef_off, (d.p0 >= 0 || d.p1 >= 0) ? 1 : 0,
                c.act(d.out), vo.slot_elems, rows, vi.Cp, d.kh == 2 ? 1 : 0, nsl, am, c.hs1_hi(),
                c.planes == 1 ? nullptr : c.hs1_lo(),
                vo.slot_elems, c.hsbits() + c.bits_act(d.in0), c.hsbits() + c.bits_node(ni), 0);
          else
            affine_fwd_kernel<false><<<grid, 256, 0, st>>>(
                c.act(d.in0), vi.slot_elems, vi.tan ? 1 : 0, c.ws + n.coef_off, (d.p0 >= 0 || d.p1 >= 0) ? 1 : 0,
                c.act(d.out), vo.slot_elems, rows, vi.Cp, d.kh == 2 ? 1 : 0, nsl, am, nullptr, nullptr, 0, nullptr,
                nullptr, 1);
        }
        LAUNCH_CHECK();
        planes_of = planes ? d.out : -1;
        break;
      }
      case CURV_OP_LAYERNORM: {
        const long long rows = (long long)P->B * vi.H * vi.W;
        layernorm_fwd_kernel<<<(int)std::min<long long>((rows + 7) / 8, 148 * 8), 256, 0, st>>>(
            c.act(d.in0), vi.slot_elems, vi.tan ? 1 : 0, c.ws + n.coef_off, (d.p0 >= 0 || d.p1 >= 0) ? 1 : 0,
            c.act(d.out), vo.slot_elems, c.ws + n.aux_off, rows, vi.C, vi.Cp, d.eps, nsl);
        LAUNCH_CHECK();
        break;
      }
      case CURV_OP_RELU:
      case CURV_OP_SIGMOID:
      case CURV_OP_GELU:
      case CURV_OP_TANH: {
        int kind = d.op == CURV_OP_RELU ? ACT_RELU : (d.op == CURV_OP_SIGMOID ? ACT_SIGMOID
                                                      : (d.op == CURV_OP_GELU ? ACT_GELU : ACT_TANH));
        long long n4 = vo.slot_elems / 4;
        act_fwd_kernel<<<dim3(grid1d(n4), 1), 256, 0, st>>>(kind, c.act(d.in0), vi.slot_elems, c.act(d.out),
                                                            vo.slot_elems, n4, 0);
        LAUNCH_CHECK();
        if (nsl > 1) {
          act_fwd_kernel<<<dim3(grid1d(n4), nsl - 1), 256, 0, st>>>(kind, c.act(d.in0), vi.slot_elems,
                                                                    c.act(d.out), vo.slot_elems, n4, 1);
          LAUNCH_CHECK();
        }
        break;
      }
      case CURV_OP_ADD: {
        const Value& vj = P->values[d.in1];
        long long n4 = vo.slot_elems / 4;
        if (d.kh == 2) {  // fused residual join + ReLU
          add_relu_fwd_kernel<<<dim3(grid1d(n4), nsl > 1 ? (nsl + 6) / 8 : 1), 256, 0, st>>>(
              c.act(d.in0), vi.slot_elems, vi.tan ? 1 : 0, c.act(d.in1), vj.slot_elems, vj.tan ? 1 : 0,
              c.act(d.out), vo.slot_elems, n4, nsl, hs_fused_absmax(c, c.bits_act(d.out), nsl));
          LAUNCH_CHECK();
          break;
        }
        // primal
        axpy_slots_kernel<<<dim3(grid1d(n4), 1), 256, 0, st>>>(c.act(d.in0), vi.slot_elems, c.act(d.out),
                                                               vo.slot_elems, n4, 0, 1.f, 0);
        LAUNCH_CHECK();
        axpy_slots_kernel<<<dim3(grid1d(n4), 1), 256, 0, st>>>(c.act(d.in1), vj.slot_elems, c.act(d.out),
                                                               vo.slot_elems, n4, 0, 1.f, 1);
        LAUNCH_CHECK();
        if (nsl > 1) {
          bool first = true;
          if (vi.tan) {
            axpy_slots_kernel<<<dim3(grid1d(n4), nsl - 1), 256, 0, st>>>(
                c.act(d.in0), vi.slot_elems, c.act(d.out), vo.slot_elems, n4, 1, 1.f, 0);
            LAUNCH_CHECK();
            first = false;
          }
          if (vj.tan) {
            axpy_slots_kernel<<<dim3(grid1d(n4), nsl - 1), 256, 0, st>>>(
                c.act(d.in1), vj.slot_elems, c.act(d.out), vo.slot_elems, n4, 1, 1.f, first ? 0 : 1);
            LAUNCH_CHECK();
          }
        }
        break;
      }
      case CURV_OP_MAXPOOL: {
        unsigned char* idx = reinterpret_cast<unsigned char*>(c.ws + n.idx_off);
        if (vi.slot_elems / 4 >= (1LL << 31) || vo.slot_elems / 4 >= (1LL << 31))
          return fail(CURV_ERR_UNSUPPORTED, "max-pool tensors beyond 2^33 elements per slot are not supported");
        maxpool_fwd_kernel<<<dim3(grid1d(vo.slot_elems / 4), 1), 256, 0, st>>>(
            c.act(d.in0), vi.slot_elems, c.act(d.out), vo.slot_elems, idx, P->B, vi.H, vi.W, vo.H, vo.W,
            vo.Cp, d.kh, d.kw, d.sh, d.sw, d.ph, d.pw, 0, hs_fused_absmax(c, c.bits_act(d.out), nsl));
        LAUNCH_CHECK();
        if (nsl > 1) {
          maxpool_fwd_kernel<<<dim3(grid1d(vo.slot_elems / 4), nsl - 1), 256, 0, st>>>(
              c.act(d.in0), vi.slot_elems, c.act(d.out), vo.slot_elems, idx, P->B, vi.H, vi.W, vo.H, vo.W,
              vo.Cp, d.kh, d.kw, d.sh, d.sw, d.ph, d.pw, 1, c.hs ? c.hsbits() + c.bits_act(d.out) : nullptr);
          LAUNCH_CHECK();
        }
        break;
      }
      case CURV_OP_RESHAPE:
        break;  // the output aliases the input
      case CURV_OP_CLSCAT: {
        const float* cls = d.p0 >= 0 ? c.param(d.p0) : c.cst(d.c0);
        clscat_fwd_kernel<<<dim3(grid1d(vo.slot_elems), nsl), 256, 0, st>>>(
            c.act(d.in0), vi.slot_elems, vi.tan ? 1 : 0, cls, (d.p0 >= 0 && K > 0) ? c.vcol(d.p0) : nullptr, c.ldk,
            c.act(d.out), vo.slot_elems, P->B, vi.W, vi.C, nsl);
        LAUNCH_CHECK();
        break;
      }
      case CURV_OP_POSADD: {
        const float* pos = d.p0 >= 0 ? c.param(d.p0) : c.cst(d.c0);
        posadd_fwd_kernel<<<dim3(grid1d(vo.slot_elems), nsl), 256, 0, st>>>(
            c.act(d.in0), vi.slot_elems, vi.tan ? 1 : 0, pos, (d.p0 >= 0 && K > 0) ? c.vcol(d.p0) : nullptr, c.ldk,
            c.act(d.out), vo.slot_elems, P->B, (long long)vi.W * vi.C, nsl);
        LAUNCH_CHECK();
        break;
      }
      case CURV_OP_TOKSEL: {
        toksel_fwd_kernel<<<dim3(grid1d(vo.slot_elems), nsl), 256, 0, st>>>(
            c.act(d.in0), vi.slot_elems, c.act(d.out), vo.slot_elems, P->B, vi.W, vi.C, d.kw);
        LAUNCH_CHECK();
        break;
      }
      case CURV_OP_ATTENTION: {
        int rc = attention_forward(c, n, vi.tan ? nsl - 1 : 0);
        if (rc) return rc;
        if (nsl > 1 && !vi.tan)
          CHECK_CUDA(cudaMemsetAsync(c.act(d.out, 1), 0, sizeof(float) * vo.slot_elems * (nsl - 1), st));
        break;
      }
      case CURV_OP_AVGPOOL: {
        avgpool_fwd_kernel<<<dim3(grid1d((long long)P->B * vo.Cp / 4), nsl), 256, 0, st>>>(
            c.act(d.in0), vi.slot_elems, c.act(d.out), vo.slot_elems, P->B, vi.H * vi.W, vi.Cp, 0);
        LAUNCH_CHECK();
        break;
      }
      default:
        return fail(CURV_ERR_UNSUPPORTED, "unsupported op in forward sweep");
    }
  }
  return CURV_OK;
}

static int gram_accumulate(const float* X, long long rows, int width, int widthp, float* F, float w,
                           float* partial, long long partial_elems, cudaStream_t st);
namespace curv {
__global__ void hs_bits_fill_kernel(uint32_t* dst, int n, const uint32_t* src, uint32_t floor_bits);
}
static int gram_hs(const __half* Xh, const __half* Xl, int planes, long long rows, int width, int ld,
                   const uint32_t* sbits, int C, int taps, float* F, float w, float* partial,
                   long long partial_elems, cudaStream_t st);
// EKFAC eigenvalue correction (ekfac.cuh)
struct EkfacEntry {
  int node;
  const float *Qa, *Qg;
  float* lam;
  int joint;      // ones column appended to the patches (joint weight + bias group)
  int bias_only;  // bias group: the "patch" is the ones column alone (Qa = [[1]], lambda is [d_out, 1])
  int identity;   // no rotation: squared per-example gradients (GGN diagonal)
};
struct EkfacJob {
  std::vector<EkfacEntry> entries;
  std::vector<char> has;  // per node: some entry collects it
  float w = 0.f;
};
struct Ctx;
static int ekfac_layer(const Ctx& c, const struct Node& n, const EkfacEntry& e, int slot_index, int abs_slot);

// backward sweep over cotangent slots [s0, s0+ns) of the grad storage.
//   GGN / VJP: s0 = 1, ns = K.      Hessian R-op: s0 = 0, ns = K+1 (slot 0 = plain backward).
//   param_out: accumulate parameter-space results into c.out (columns k0..k0+K-1)
static int backward(const Ctx& c, int K) {
  curv_program* P = c.P;
  cudaStream_t st = c.st;
  const bool rop = c.rop;
  const int s0 = rop ? 0 : 1;
  const int ns = rop ? K + 1 : K;
  const int kskip = rop ? 1 : 0;
  std::vector<char> ginit(P->values.size(), 0);
  ginit[P->nodes.back().d.out] = 1;
  float* scratch = c.ws + P->scratch_off;
  int planes_of = -1;  // value whose cotangent slots currently sit in the hs1 planes (written by affine_bwd)
  // number of kernels that have written (accumulated into) the cotangent of a value: a fused absmax word is only
  // trusted while its value has a single writer
  std::vector<int> nwrites(P->values.size(), 0);
  auto mark_written = [&](int v) { ginit[v] = 1; ++nwrites[v]; };
  for (int ni = (int)P->nodes.size() - 1; ni >= 0; --ni) {
    Node& n = P->nodes[ni];
    const curv_node_desc& d = n.d;
    if (d.op == CURV_OP_INPUT) continue;
    const Value& vi = P->values[d.in0];
    const Value& vo = P->values[d.out];
    if (!vo.tan) continue;
    if (!ginit[d.out]) continue;  // value does not influence the prediction (dead branch)
    switch (d.op) {
      case CURV_OP_CONV: {
        const Geom& g = n.fwd;
        // KFAC: Gram matrix of the cotangent slots instead of parameter gradients; on the tensor cores (from the
        // operand planes the dgrad reads as well) when the layer is large enough, else fp32 SIMT / 3xTF32
        const bool kfac_g = c.kfac_G != nullptr && c.kfac_G[ni] != nullptr;
        const bool hs_g = kfac_g && c.hs && g.M >= 256 && vo.C >= 16 && vo.Cp % 8 == 0 && !(g_tc_disable & 1024) &&
                          vo.slot_elems * ns <= P->hs1_elems;
        if (kfac_g && !hs_g) {
          int rc = gram_accumulate(c.grad(d.out, 1), (long long)K * g.M, vo.C, vo.Cp, c.kfac_G[ni],
                                   c.kfac_wG, scratch, P->scratch_elems, st);
          if (rc) return rc;
        }
        const int nidx = ni;
        const int wns = rop ? K : ns;  // slots whose weight gradient is wanted
        const bool hs_w = c.kfac_G == nullptr && d.p0 >= 0 && hs_wgr_ok(c, n, wns) &&
                          vo.slot_elems * ns <= P->hs1_elems;
        const bool hs_d = vi.tan && hs_dgr_ok(c, n);
        const bool ek = c.ekfac != nullptr && c.ekfac->has[ni];
        const int eg = c.bits_grad(d.out);
        if ((hs_w || hs_d || hs_g || ek) && planes_of != d.out) {  // fp16 hi/lo planes of the cotangent slots (wgrad + dgrad)
          int rc = hs_absmax(c, c.grad(d.out, s0), vo.slot_elems, vo.slot_elems, eg + s0, ns);
          if (!rc) rc = hs_split(c, c.grad(d.out, s0), vo.slot_elems, vo.slot_elems, c.hs1_hi(), c.hs1_lo(),
                                 eg + s0, ns);
          if (rc) return rc;
        }
        planes_of = -1;
        if (hs_g) {  // one Gram per slot (the slots carry different scales)
          uint32_t* sb = c.hsbits() + c.bits_node(nidx) + (1 + P->kmax);
          for (int sl = 0; sl < ns; ++sl) {
            hs_bits_fill_kernel<<<1, 32, 0, st>>>(sb, 8, c.hsbits() + eg + s0 + sl, 0u);
            LAUNCH_CHECK();
            int rc = gram_hs(c.hs1_hi() + (long long)sl * vo.slot_elems,
                             c.planes == 1 ? nullptr : c.hs1_lo() + (long long)sl * vo.slot_elems, c.planes, g.M, vo.C,
                             vo.Cp, sb, vo.C, 1, c.kfac_G[nidx], c.kfac_wG, scratch, P->scratch_elems, st);
            if (rc) return rc;
          }
        }
        if (ek) {  // EKFAC: per-example gradients in the Kronecker eigenbasis, squared and summed (one slot at a time)
          for (const EkfacEntry& e : c.ekfac->entries) {
            if (e.node != nidx) continue;
            for (int sl = 0; sl < ns; ++sl) {
              int rc = ekfac_layer(c, n, e, sl, s0 + sl);
              if (rc) return rc;
            }
          }
        }
        if (hs_w) {  // weight gradients of all slots on the half-split kernel
          const int ea = c.bits_act(d.in0);
          int rc = hs_absmax(c, c.act(d.in0), 0, vi.slot_elems, ea, 1);
          if (!rc) rc = hs_split(c, c.act(d.in0), 0, vi.slot_elems, c.hs2_hi(), c.hs2_lo(), ea, 1);
          if (rc) return rc;
          // R-op: the weight gradient of the plain backward (slot 0) is not part of H v; the planes hold slot 0 first
          const int wskip = rop ? 1 : 0;
          HsWgradArgs h;
          memset(&h, 0, sizeof(h));
          h.g = g;
          h.Gh = c.hs1_hi() + (long long)wskip * vo.slot_elems; h.Gl = c.hs1_lo() + (long long)wskip * vo.slot_elems;
          h.G_slot = vo.slot_elems; h.Ng = vo.Cp; h.g_bits = c.hsbits() + eg;
          h.Ih = c.hs2_hi(); h.Il = c.hs2_lo(); h.i_bits = c.hsbits() + ea;
          h.partial = scratch; h.nsplit = n.nsplit; h.nslots = wns; h.slot0 = s0 + wskip; h.m_per_split = n.m_per_split;
          h.planes = c.planes;
          {
            ProfScope prof(1, conv_flops(g, vi.C) * wns, st);
            if (hs_launch_wgrad(h, st)) return fail(CURV_ERR_CUDA, "half-split wgrad GEMM launch failed");
            ++g_launches;
          }
          {
            int rc2 = launch_wgrad_finish(scratch, n.nsplit, wns, 0, g.N, vi.C, vi.Cp, g.KH * g.KW, c.out,
                                          P->params[d.p0].offset, c.ldk, c.k0, c.alpha, n.wsize, st);
            if (rc2) return rc2;
          }
          if (rop && vi.tan) {
            // R-op second term  delta_0^T . da_k  (k = 1..K): the plain-backward cotangent is the shared operand, the
            // tangent activations change per column -> one launch per column with the COLUMN BLOCKS of delta_0 as the
            // kernel's slots (full-width MMAs, as the Gram kernels of kfac.cuh)
            const int W8 = 64 * ceil_div(vo.Cp, 512), NS = ceil_div(vo.Cp, W8);
            uint32_t* sb = c.hsbits() + (P->hsbits_count - 16);
            hs_bits_fill_kernel<<<1, 32, 0, st>>>(sb, 8, c.planes == 1 ? nullptr : c.hsbits() + eg, 0u);
            LAUNCH_CHECK();
            for (int k = 1; k <= K; ++k) {
              rc = hs_absmax(c, c.act(d.in0, k), 0, vi.slot_elems, ea + k, 1);
              if (!rc) rc = hs_split(c, c.act(d.in0, k), 0, vi.slot_elems, c.hs2_hi(), c.hs2_lo(), ea + k, 1);
              if (rc) return rc;
              HsWgradArgs a;
              memset(&a, 0, sizeof(a));
              a.g = g; a.g.N = W8; a.g.Nd = W8;
              a.Gh = c.hs1_hi(); a.Gl = c.hs1_lo(); a.G_slot = W8; a.G_ld = vo.Cp; a.Ng = W8; a.g_bits = sb;
              a.Ih = c.hs2_hi(); a.Il = c.hs2_lo(); a.i_bits = c.hsbits() + ea + k;
              a.partial = scratch; a.nsplit = n.nsplit; a.nslots = NS; a.slot0 = 0; a.m_per_split = n.m_per_split;
              a.planes = c.planes;
              {
                ProfScope prof(1, conv_flops(g, vi.C), st);
                if (hs_launch_wgrad(a, st)) return fail(CURV_ERR_CUDA, "half-split R-op wgrad GEMM launch failed");
                ++g_launches;
              }
              int rc2 = launch_wgrad_finish(scratch, n.nsplit, 1, 0, g.N, vi.C, vi.Cp, g.KH * g.KW, c.out,
                                            P->params[d.p0].offset, c.ldk, c.k0 + k - 1, c.alpha, n.wsize, st,
                                            NS * W8);
              if (rc2) return rc2;
            }
          }
        } else if (c.kfac_G == nullptr && d.p0 >= 0) {  // weight gradient
          WgradArgs a;
          memset(&a, 0, sizeof(a));
          a.g = g;
          a.G = c.grad(d.out); a.G_slot = vo.slot_elems; a.Ng = vo.Cp;
          a.In = c.act(d.in0); a.In_slot = vi.slot_elems;
          a.second_seg = (rop && vi.tan) ? 1 : 0;
          a.partial = scratch; a.nsplit = n.nsplit; a.nslots = ns; a.slot0 = s0;
          a.m_per_split = n.m_per_split;
          int rc = launch_wgrad(a, n.wbm, n.wbn, st, conv_flops(g, vi.C) * ns * (a.second_seg ? 2 : 1));
          if (rc) return rc;
          {
            int rc2 = launch_wgrad_finish(scratch, n.nsplit, ns, kskip, g.N, vi.C, vi.Cp, g.KH * g.KW, c.out,
                                          P->params[d.p0].offset, c.ldk, c.k0, c.alpha, n.wsize, st);
            if (rc2) return rc2;
          }
        }
        if (c.kfac_G == nullptr && d.p1 >= 0) {  // bias gradient: column sums of the cotangent
          long long rows = g.M;
          affine_bwd_kernel<<<dim3(n.nchunks, ns), 256, 8192, st>>>(
              c.grad(d.out), vo.slot_elems, nullptr, nullptr, 0, nullptr, nullptr, nullptr, 0, 0, scratch, 1,
              rows, vo.Cp, n.rows_per_cta, s0, ns, 0, 0, 0, nullptr, nullptr, nullptr, 0, nullptr, nullptr, 1);
          LAUNCH_CHECK();
          vec_grad_finish_kernel<<<ceil_div(vo.C * K, 8), 256, 0, st>>>(
              scratch, n.nchunks, ns, kskip, 1, vo.C, vo.Cp, c.out, P->params[d.p1].offset, c.ldk, c.k0,
              c.alpha);
          LAUNCH_CHECK();
        }
        if (hs_d) {  // data gradient on the half-split kernel
          HsGatherArgs h;
          memset(&h, 0, sizeof(h));
          h.g = n.dgr;
          h.Ah = c.hs1_hi(); h.Al = c.hs1_lo(); h.A_slot = vo.slot_elems; h.a_slot_base = s0; h.a_has_slots = 1;
          h.a_bits = c.hsbits() + eg;
          h.W_img = reinterpret_cast<const __half*>(c.ws + n.wtimg_off);
          if (rop && d.p0 >= 0 && n.wtimgt_off >= 0) {  // R-op: + dW_k^T . delta_0 (second segment of slot k >= 1)
            h.Wt_img = reinterpret_cast<const __half*>(c.ws + n.wtimgt_off);
            h.Wt_img_slot = hs_image_halves(n.dgr.Nd, n.dgr.Kd, c.planes);
          }
          h.w_bits = c.hsbits() + c.bits_node(nidx);
          h.out = c.grad(d.in0); h.out_slot = vi.slot_elems; h.slot0 = s0; h.accumulate = ginit[d.in0];
          h.planes = c.planes;
          if (!ginit[d.in0]) {  // first writer: track the absmax of the data gradient for its consumers
            hs_fused_absmax(c, c.bits_grad(d.in0) + s0, ns);
            h.out_bits = c.hsbits() + c.bits_grad(d.in0);
          }
          {
            ProfScope prof(0, conv_flops(g, vi.C) * (ns + (h.Wt_img ? K : 0)), st);
            if (hs_launch_gather_gemm(h, ns, st, true, ns)) return fail(CURV_ERR_CUDA, "half-split dgrad GEMM launch failed");
            ++g_launches;
          }
          mark_written(d.in0);
        } else if (vi.tan) {  // data gradient
          GatherGemmArgs a;
          memset(&a, 0, sizeof(a));
          a.g = n.dgr;
          a.A = c.grad(d.out); a.A_slot = vo.slot_elems; a.a_has_slots = 1;
          a.W = c.ws + n.wt_off;
          a.Wt = (rop && n.wtt_off >= 0) ? c.ws + n.wtt_off : nullptr; a.Wt_slot = n.wtsize;
          if (g_tc_mode && !(g_tc_disable & 1) && n.wtimg_off >= 0 && tc_gather_eligible(n.dgr, g_tc_mode)) {
            a.W_img = c.ws + n.wtimg_off;
            a.Wt_img = (rop && n.wtimgt_off >= 0) ? c.ws + n.wtimgt_off : nullptr; a.Wt_img_slot = n.wtimg_size;
          }
          a.out = c.grad(d.in0); a.out_slot = vi.slot_elems;
          a.slot0 = s0; a.accumulate = ginit[d.in0];
          int rc = launch_gather_gemm(a, ns, st, conv_flops(g, vi.C) * ns * (a.Wt ? 2 : 1));
          if (rc) return rc;
          mark_written(d.in0);
        }
        break;
      }
      case CURV_OP_AFFINE: {
        long long rows = (long long)P->B * vi.H * vi.W;
        int want_partial = (d.p0 >= 0 || d.p1 >= 0) ? 1 : 0;
        if (!vi.tan && !want_partial) break;
        // the cotangent written here is final when this is its first (and then only) writer: its absmax feeds
        // the half-split GEMMs of the producing conv (kernel indexes the bits by absolute slot)
        unsigned int* amax = nullptr;
        if (c.hs && vi.tan && !ginit[d.in0]) {
          hs_fused_absmax(c, c.bits_grad(d.in0) + s0, ns);
          amax = c.hsbits() + c.bits_grad(d.in0);
        }
        // Planes mode: when the convolution that produced this value is the very next node of the sweep and both
        // of its contractions run on the half-split kernels, the cotangent is written straight as their fp16
        // hi/lo planes (scaled by a bound derived from the absmax of the incoming cotangent), and the fp32 copy -
        // which nobody else reads - is skipped together with the separate split pass.
        bool planes = false;
        if (amax && !rop && c.kfac_G == nullptr && ni > 0) {
          const Node& cn = P->nodes[ni - 1];
          if (cn.d.op == CURV_OP_CONV && cn.d.out == d.in0 && cn.d.p1 < 0) {
            const bool need_w = cn.d.p0 >= 0, need_d = P->values[cn.d.in0].tan;
            const bool ok_w = !need_w || hs_wgr_ok(c, cn, ns), ok_d = !need_d || hs_dgr_ok(c, cn);
            planes = (need_w || need_d) && ok_w && ok_d && vi.slot_elems * ns <= P->hs1_elems &&
                     !(g_tc_disable & 64);
          }
        }
        if (planes) {  // the bound needs the absmax of the incoming cotangent (fused by its producer, else a pass)
          if (nwrites[d.out] != 1)
            for (int sl = 0; sl < ns; ++sl) (*c.hs_valid)[c.bits_grad(d.out) + s0 + sl] = 0;
          int rc = hs_absmax(c, c.grad(d.out, s0), vo.slot_elems, vo.slot_elems, c.bits_grad(d.out) + s0, ns);
          if (rc) return rc;
        }
        affine_bwd_kernel<<<dim3(n.nchunks, ns), 256, 8192, st>>>(
            c.grad(d.out), vo.slot_elems, c.act(d.in0), (rop && vi.tan) ? c.act(d.in0) : nullptr,
            vi.slot_elems, c.ws + n.coef_off, c.ws + n.aux_off, vi.tan ? c.grad(d.in0) : nullptr,
            vi.slot_elems, vi.tan ? 1 : 0, scratch, want_partial, rows, vi.Cp, n.rows_per_cta, s0, ns,
            rop ? 1 : 0, ginit[d.in0], d.kh == 2 ? 1 : 0, amax, planes ? c.hs1_hi() : nullptr,
            (planes && c.planes == 2) ? c.hs1_lo() : nullptr, vi.slot_elems, c.hsbits() + c.bits_grad(d.out),
            c.hsbits() + c.bits_node(ni), planes ? 0 : 1);
        LAUNCH_CHECK();
        planes_of = planes ? d.in0 : -1;
        if (vi.tan) mark_written(d.in0);
        if (d.p0 >= 0) {
          vec_grad_finish_kernel<<<ceil_div(vi.C * K, 8), 256, 0, st>>>(
              scratch, n.nchunks, ns, kskip, 0, vi.C, vi.Cp, c.out, P->params[d.p0].offset, c.ldk, c.k0,
              c.alpha);
          LAUNCH_CHECK();
        }
        if (d.p1 >= 0) {
          vec_grad_finish_kernel<<<ceil_div(vi.C * K, 8), 256, 0, st>>>(
              scratch, n.nchunks, ns, kskip, 1, vi.C, vi.Cp, c.out, P->params[d.p1].offset, c.ldk, c.k0,
              c.alpha);
          LAUNCH_CHECK();
        }
        break;
      }
      case CURV_OP_LAYERNORM: {
        const long long rows = (long long)P->B * vi.H * vi.W;
        const int want_partial = (d.p0 >= 0 || d.p1 >= 0) ? 1 : 0;
        if (!vi.tan && !want_partial) break;
        layernorm_bwd_kernel<<<dim3(n.nchunks, ns), 256, 0, st>>>(
            c.grad(d.out), vo.slot_elems, c.act(d.in0), c.ws + n.aux_off, c.ws + n.coef_off,
            vi.tan ? c.grad(d.in0) : nullptr, vi.slot_elems, vi.tan ? 1 : 0, ginit[d.in0], scratch, want_partial, rows,
            vi.C, vi.Cp, n.rows_per_cta, s0, ns);
        LAUNCH_CHECK();
        if (vi.tan) mark_written(d.in0);
        if (c.kfac_G == nullptr) {
          for (int which = 0; which < 2; ++which) {
            const int pi = which == 0 ? d.p0 : d.p1;
            if (pi < 0) continue;
            vec_grad_finish_kernel<<<ceil_div(vi.C * K, 8), 256, 0, st>>>(
                scratch, n.nchunks, ns, kskip, which, vi.C, vi.Cp, c.out, P->params[pi].offset, c.ldk, c.k0, c.alpha);
            LAUNCH_CHECK();
          }
        }
        break;
      }
      case CURV_OP_RELU:
      case CURV_OP_SIGMOID:
      case CURV_OP_GELU:
      case CURV_OP_TANH: {
        if (!vi.tan) break;
        int kind = d.op == CURV_OP_RELU ? ACT_RELU : (d.op == CURV_OP_SIGMOID ? ACT_SIGMOID
                                                      : (d.op == CURV_OP_GELU ? ACT_GELU : ACT_TANH));
        long long n4 = vo.slot_elems / 4;
        act_bwd_kernel<<<dim3(grid1d(n4), ns), 256, 0, st>>>(
            kind, c.grad(d.out), vo.slot_elems, kind == ACT_GELU ? c.act(d.in0) : c.act(d.out), c.grad(d.in0), vi.slot_elems,
            rop ? c.act(d.in0) : nullptr, vi.slot_elems, n4, s0, ginit[d.in0]);
        LAUNCH_CHECK();
        mark_written(d.in0);
        break;
      }
      case CURV_OP_ADD: {
        const Value& vj = P->values[d.in1];
        long long n4 = vo.slot_elems / 4;
        if (d.kh == 2) {  // fused residual join + ReLU
          add_relu_bwd_kernel<<<dim3(grid1d(n4), (ns + 7) / 8), 256, 0, st>>>(
              c.grad(d.out), vo.slot_elems, c.act(d.out), vi.tan ? c.grad(d.in0) : nullptr, vi.slot_elems,
              ginit[d.in0], vj.tan ? c.grad(d.in1) : nullptr, vj.slot_elems, ginit[d.in1], n4, s0, ns,
              (c.hs && vi.tan && !ginit[d.in0]) ? (hs_fused_absmax(c, c.bits_grad(d.in0) + s0, ns),
                                                   c.hsbits() + c.bits_grad(d.in0)) : nullptr,
              (c.hs && vj.tan && !ginit[d.in1]) ? (hs_fused_absmax(c, c.bits_grad(d.in1) + s0, ns),
                                                   c.hsbits() + c.bits_grad(d.in1)) : nullptr);
          LAUNCH_CHECK();
          if (vi.tan) mark_written(d.in0);
          if (vj.tan) mark_written(d.in1);
          break;
        }
        if (vi.tan) {
          axpy_slots_kernel<<<dim3(grid1d(n4), ns), 256, 0, st>>>(c.grad(d.out), vo.slot_elems, c.grad(d.in0),
                                                                  vi.slot_elems, n4, s0, 1.f, ginit[d.in0]);
          LAUNCH_CHECK();
          mark_written(d.in0);
        }
        if (vj.tan) {
          axpy_slots_kernel<<<dim3(grid1d(n4), ns), 256, 0, st>>>(c.grad(d.out), vo.slot_elems, c.grad(d.in1),
                                                                  vj.slot_elems, n4, s0, 1.f, ginit[d.in1]);
          LAUNCH_CHECK();
          mark_written(d.in1);
        }
        break;
      }
      case CURV_OP_MAXPOOL: {
        if (!vi.tan) break;
        unsigned char* idx = reinterpret_cast<unsigned char*>(c.ws + n.idx_off);
        maxpool_bwd_kernel<<<dim3(grid1d(vi.slot_elems / 4), (ns + 7) / 8), 256, 0, st>>>(
            c.grad(d.out), vo.slot_elems, c.grad(d.in0), vi.slot_elems, idx, P->B, vi.H, vi.W, vo.H, vo.W,
            vo.Cp, d.kh, d.kw, d.sh, d.sw, d.ph, d.pw, s0, ns, ginit[d.in0],
            (c.hs && !ginit[d.in0]) ? (hs_fused_absmax(c, c.bits_grad(d.in0) + s0, ns),
                                       c.hsbits() + c.bits_grad(d.in0)) : nullptr);
        LAUNCH_CHECK();
        mark_written(d.in0);
        break;
      }
      case CURV_OP_RESHAPE:
        if (vi.tan) mark_written(d.in0);  // same storage: the cotangent of the output IS the input's
        break;
      case CURV_OP_CLSCAT:
      case CURV_OP_POSADD: {
        const bool cat = d.op == CURV_OP_CLSCAT;
        if (d.p0 >= 0 && c.kfac_G == nullptr) {  // parameter gradient: cotangent summed over the batch
          const long long elems = cat ? vo.C : (long long)vo.W * vo.C;
          batch_sum_grad_kernel<<<grid1d(elems * ns), 256, 0, st>>>(
              c.grad(d.out), vo.slot_elems, P->B, (long long)vo.W * vo.C, elems, s0, ns, c.out,
              P->params[d.p0].offset, c.ldk, c.k0, c.alpha);
          LAUNCH_CHECK();
        }
        if (!vi.tan) break;
        if (cat) {
          clscat_bwd_kernel<<<dim3(grid1d(vi.slot_elems), ns), 256, 0, st>>>(
              c.grad(d.out), vo.slot_elems, c.grad(d.in0), vi.slot_elems, P->B, vi.W, vi.C, s0, ginit[d.in0]);
        } else {
          axpy_slots_kernel<<<dim3(grid1d(vo.slot_elems / 4), ns), 256, 0, st>>>(
              c.grad(d.out), vo.slot_elems, c.grad(d.in0), vi.slot_elems, vo.slot_elems / 4, s0, 1.f, ginit[d.in0]);
        }
        LAUNCH_CHECK();
        mark_written(d.in0);
        break;
      }
      case CURV_OP_TOKSEL: {
        if (!vi.tan) break;
        toksel_bwd_kernel<<<dim3(grid1d(vi.slot_elems), ns), 256, 0, st>>>(
            c.grad(d.out), vo.slot_elems, c.grad(d.in0), vi.slot_elems, P->B, vi.W, vi.C, d.kw, s0, ginit[d.in0]);
        LAUNCH_CHECK();
        mark_written(d.in0);
        break;
      }
      case CURV_OP_ATTENTION: {
        if (!vi.tan) break;
        int rc = attention_backward(c, n, s0, ns, ginit[d.in0]);
        if (rc) return rc;
        mark_written(d.in0);
        break;
      }
      case CURV_OP_AVGPOOL: {
        if (!vi.tan) break;
        avgpool_bwd_kernel<<<dim3(grid1d(vi.slot_elems / 4), ns), 256, 0, st>>>(
            c.grad(d.out), vo.slot_elems, c.grad(d.in0), vi.slot_elems, P->B, vi.H * vi.W, vi.Cp, s0,
            ginit[d.in0]);
        LAUNCH_CHECK();
        mark_written(d.in0);
        break;
      }
      default:
        return fail(CURV_ERR_UNSUPPORTED, "unsupported op in backward sweep");
    }
    {
      int rc = signal_out_done(c, n);
      if (rc) return rc;
    }
  }
  return CURV_OK;
}

static int matmat_batch_impl(curv_program* P, int kind, int loss, const void* const* param_ptrs,
                             const void* const* const_ptrs, const void* X, const void* y, const float* mc_grad,
                             int mc_samples, const float* V, float* out, int K, int ldk, int k0, float loss_scale,
                             float alpha, void* workspace, size_t workspace_bytes, void* stream,
                             void* const* v_ready, void* const* out_done);

extern "C" int curv_matmat_batch(curv_program* P, int kind, int loss, const void* const* param_ptrs,
                                 const void* const* const_ptrs, const void* X, const void* y,
                                 const float* mc_grad, int mc_samples, const float* V, float* out, int K,
                                 int ldk, int k0, float loss_scale, float alpha, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  return matmat_batch_impl(P, kind, loss, param_ptrs, const_ptrs, X, y, mc_grad, mc_samples, V, out, K, ldk, k0,
                           loss_scale, alpha, workspace, workspace_bytes, stream, nullptr, nullptr);
}

// Streaming variant: v_ready[p] / out_done[p] are cudaEvent_t handles per parameter (either array or single entries
// may be NULL).  The columns of V of parameter p are only read after v_ready[p] (the node that owns p is prepared
// lazily inside the forward sweep), and out_done[p] is recorded on `stream` as soon as the rows of parameter p in
// `out` are final, so a caller can overlap the upload of V and the download of the result with the sweeps.
extern "C" int curv_matmat_batch_sync(curv_program* P, int kind, int loss, const void* const* param_ptrs,
                                      const void* const* const_ptrs, const void* X, const void* y,
                                      const float* mc_grad, int mc_samples, const float* V, float* out, int K,
                                      int ldk, int k0, float loss_scale, float alpha, void* workspace,
                                      size_t workspace_bytes, void* stream, void* const* v_ready,
                                      void* const* out_done) {
  return matmat_batch_impl(P, kind, loss, param_ptrs, const_ptrs, X, y, mc_grad, mc_samples, V, out, K, ldk, k0,
                           loss_scale, alpha, workspace, workspace_bytes, stream, v_ready, out_done);
}

static int matmat_batch_impl(curv_program* P, int kind, int loss, const void* const* param_ptrs,
                                 const void* const* const_ptrs, const void* X, const void* y,
                                 const float* mc_grad, int mc_samples, const float* V, float* out, int K,
                                 int ldk, int k0, float loss_scale, float alpha, void* workspace,
                                 size_t workspace_bytes, void* stream, void* const* v_ready,
                                 void* const* out_done) {
  if (!P) return fail(CURV_ERR_INVALID, "null program");
  if (kind != CURV_KIND_FORWARD && (K < 1 || K > P->kmax))
    return fail(CURV_ERR_INVALID, "K must be in [1, kmax]");
  if (workspace_bytes < P->ws_bytes || !workspace) return fail(CURV_ERR_WORKSPACE, "workspace too small");
  if (kind == CURV_KIND_HESSIAN && !(P->hessian & 1))
    return fail(CURV_ERR_INVALID, "program was not created with hessian=1");
  if (kind == CURV_KIND_GGN_MC && (mc_samples < 1 || mc_samples > 32 || !mc_grad))
    return fail(CURV_ERR_INVALID, "MC mode needs 1 <= mc_samples <= 32 and mc_grad");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(CURV_ERR_CUDA, "no CUDA device: curvb200 has no CPU fallback");
  Ctx c;
  c.P = P; c.ws = (float*)workspace; c.pp = param_ptrs; c.cp = const_ptrs; c.V = V; c.out = out;
  c.K = K; c.ldk = ldk; c.k0 = k0; c.alpha = alpha; c.st = (cudaStream_t)stream; c.kind = kind;
  c.rop = kind == CURV_KIND_HESSIAN;
  c.v_ready = v_ready; c.out_done = out_done;
  std::vector<char> hs_valid;
  if (g_tc_mode && !(g_tc_disable & 32) && !(c.rop && (g_tc_disable & 2048)) && P->hs1_elems > 0 && hs_ready() > 0) {
    c.hs = true;
    c.planes = (P->hessian & 4) ? 1 : 2;
    hs_valid.assign((size_t)P->hsbits_count, 0);
    c.hs_valid = &hs_valid;
    CHECK_CUDA(cudaMemsetAsync(c.hsbits(), 0, (size_t)P->hsbits_count * sizeof(uint32_t), c.st));
  }
  const int last = P->nodes.back().d.out;
  const Value& vl = P->values[last];
  int rc;
  if (kind == CURV_KIND_FORWARD) {
    if ((rc = prepare_params(c, false))) return rc;
    return forward(c, X, 0);
  }
  if (!vl.tan) return CURV_OK;  // prediction independent of the selected parameters: zero matrix
  if (kind == CURV_KIND_VJP) {
    if ((rc = prepare_params(c, false))) return rc;
    if ((rc = forward(c, X, 0))) return rc;
    import_pred_kernel<<<grid1d((long long)P->B * vl.Cp * K), 256, 0, c.st>>>(
        c.grad(last), vl.slot_elems, 1, V, P->B, vl.C, vl.Cp, K, ldk, k0);
    LAUNCH_CHECK();
    return backward(c, K);
  }
  if ((rc = prepare_params(c, true))) return rc;
  if ((rc = forward(c, X, K))) return rc;
  if (kind == CURV_KIND_JVP) {
    export_pred_kernel<<<grid1d((long long)P->B * vl.C * K), 256, 0, c.st>>>(
        c.act(last), vl.slot_elems, 1, out, P->B, vl.C, vl.Cp, K, ldk, k0);
    LAUNCH_CHECK();
    return CURV_OK;
  }
  loss_hessian_kernel<<<P->B, 128, vl.Cp * sizeof(float), c.st>>>(
      loss, kind == CURV_KIND_GGN_MC ? mc_samples : 0, c.act(last), c.act(last), vl.slot_elems, c.grad(last),
      vl.slot_elems, y, mc_grad, vl.C, vl.Cp, K, loss_scale, c.rop ? 1 : 0);
  LAUNCH_CHECK();
  return backward(c, K);
}

// ------------------------------------------------------------------------------------------------
// dense helpers (Kronecker / eigen-basis apply)
// ------------------------------------------------------------------------------------------------
static int dense_gemm(int tA, int tB, int M, int N, int Kd, float alpha, const float* A, int lda,
                      const float* B, int ldb, float beta, float* C, int ldc, int batch, long long sA,
                      long long sB, long long sC, cudaStream_t st) {
  if (M <= 0 || N <= 0) return CURV_OK;
  dim3 grid(ceil_div(N, 64), ceil_div(M, 64), batch);
  dense_gemm_simt<<<grid, 256, 0, st>>>(tA, tB, M, N, Kd, alpha, A, lda, B, ldb, beta, C, ldc, sA, sB, sC);
  LAUNCH_CHECK();
  return CURV_OK;
}

extern "C" int curv_gemm(int transA, int transB, int M, int N, int Kd, float alpha, const float* A,
                         int lda, const float* B, int ldb, float beta, float* C, int ldc, void* stream) {
  return dense_gemm(transA, transB, M, N, Kd, alpha, A, lda, B, ldb, beta, C, ldc, 1, 0, 0, 0,
                    (cudaStream_t)stream);
}

// `batch` independent products C_b = alpha op(A_b) op(B_b) + beta C_b with element strides sA / sB / sC between them
// (one launch; the per-example contractions of the EKFAC eigenvalue correction, ekfac_hooks.py:215-236)
extern "C" int curv_gemm_batched(int transA, int transB, int M, int N, int Kd, float alpha, const float* A, int lda,
                                 long long sA, const float* B, int ldb, long long sB, float beta, float* C, int ldc,
                                 long long sC, int batch, void* stream) {
  if (batch < 1 || batch > 65535) return fail(CURV_ERR_INVALID, "curv_gemm_batched: 1 <= batch <= 65535");
  return dense_gemm(transA, transB, M, N, Kd, alpha, A, lda, B, ldb, beta, C, ldc, batch, sA, sB, sC,
                    (cudaStream_t)stream);
}

// X, Y: [d_out][d_in][K] (K minor).  Y = G X A^T per column.
//   step 1: T[a][(b,z)] = sum_a' G[a][a'] X[a'][(b,z)]           plain GEMM, N = d_in*K
//   step 2: Y[a][B][z]  = sum_b  A[B][b] T[a][b][z]              batched over a: A (d_in x d_in) @ T_a (d_in x K)
extern "C" int curv_kron_apply(const float* G, const float* A, int d_out, int d_in, int K, const float* X,
                               float* Y, float* tmp, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const float* T = X;
  int rc;
  if (G) {
    float* dst = A ? tmp : Y;
    if ((rc = dense_gemm(0, 0, d_out, d_in * K, d_out, 1.f, G, d_out, X, d_in * K, 0.f, dst, d_in * K, 1, 0, 0,
                         0, st)))
      return rc;
    T = dst;
  }
  if (A) {
    if ((rc = dense_gemm(0, 0, d_in, K, d_in, 1.f, A, d_in, T, K, 0.f, Y, K, d_out, 0, (long long)d_in * K,
                         (long long)d_in * K, st)))
      return rc;
  } else if (!G) {
    CHECK_CUDA(cudaMemcpyAsync(Y, X, sizeof(float) * d_out * d_in * K, cudaMemcpyDeviceToDevice, st));
  }
  return CURV_OK;
}

__global__ void eig_scale_kernel(float* __restrict__ T, const float* __restrict__ lam, float damping,
                                 int inverse, long long n, int K) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n * K;
       i += (long long)gridDim.x * blockDim.x) {
    float l = lam[i / K];
    T[i] *= inverse ? 1.f / (l + damping) : l;
  }
}

// Y = (Qg (x) Qa) diag(scale(lambda)) (Qg (x) Qa)^T X   with lambda [d_out][d_in]
extern "C" int curv_eigh_apply(const float* Qg, const float* Qa, const float* lambda, float damping,
                               int inverse, int d_out, int d_in, int K, const float* X, float* Y,
                               float* tmp, float* tmp2, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  // T1 = Qg^T X  (rows)
  const float* cur = X;
  if (Qg) {
    if ((rc = dense_gemm(1, 0, d_out, d_in * K, d_out, 1.f, Qg, d_out, cur, d_in * K, 0.f, tmp, d_in * K, 1, 0,
                         0, 0, st)))
      return rc;
    cur = tmp;
  }
  // T2_a = Qa^T T1_a
  if (Qa) {
    if ((rc = dense_gemm(1, 0, d_in, K, d_in, 1.f, Qa, d_in, cur, K, 0.f, tmp2, K, d_out, 0,
                         (long long)d_in * K, (long long)d_in * K, st)))
      return rc;
  } else {
    CHECK_CUDA(cudaMemcpyAsync(tmp2, cur, sizeof(float) * d_out * d_in * K, cudaMemcpyDeviceToDevice, st));
  }
  eig_scale_kernel<<<grid1d((long long)d_out * d_in * K), 256, 0, st>>>(tmp2, lambda, damping, inverse,
                                                                       (long long)d_out * d_in, K);
  LAUNCH_CHECK();
  // back: T3_a = Qa T2_a ; Y = Qg T3
  const float* back = tmp2;
  if (Qa) {
    float* dst = Qg ? tmp : Y;
    if ((rc = dense_gemm(0, 0, d_in, K, d_in, 1.f, Qa, d_in, back, K, 0.f, dst, K, d_out, 0,
                         (long long)d_in * K, (long long)d_in * K, st)))
      return rc;
    back = dst;
  }
  if (Qg) {
    if ((rc = dense_gemm(0, 0, d_out, d_in * K, d_out, 1.f, Qg, d_out, back, d_in * K, 0.f, Y, d_in * K, 1, 0,
                         0, 0, st)))
      return rc;
  } else if (back != Y) {
    CHECK_CUDA(cudaMemcpyAsync(Y, back, sizeof(float) * d_out * d_in * K, cudaMemcpyDeviceToDevice, st));
  }
  return CURV_OK;
}

#include "kfac.cuh"
#include "kron_tc.cuh"
#include "ekfac.cuh"
#include "lanczos.cuh"
