// tcgen05 tensor-core path of the gather GEMM (forward conv + K tangents, dgrad), sm_100a only.
//
//   out[m][n] = sum_seg sum_r gather(A_seg)[m][r] * W_seg[n][r]        (same contract as gemm_simt.cuh)
//
// fp32 in, fp32 out, fp32-grade accuracy on the TF32 tensor pipe by the 3xTF32 split
//   a = a_hi + a_lo,  b = b_hi + b_lo   (hi = top 19 bits, lo = remainder, both exactly TF32)
//   a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo            (dropped a_lo*b_lo ~ 2^-22 relative)
// accumulated in fp32 in tensor memory.
//
// Persistent, warp-specialised CTA (544 threads, one CTA per SM):
//   warps 0-3, 13-16  epilogue: two groups of 4 warps, each owning half of the BN accumulator columns
//   warp  4           MMA     : one elected lane issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8), 12 per stage
//   warps 5-12        producers: implicit-im2col gather, global -> registers -> hi/lo split ->
//                          st.shared into the 128B-swizzled UMMA layout (software "TMA": the gather is per
//                          16-byte channel group, which TMA tiles cannot express for padded / strided /
//                          transposed-stride convolutions); weights arrive by cp.async.bulk
// Pipelines: smem ring (full/empty mbarriers, STAGES deep), TMEM double buffer (tmem_full/empty).
//
// Accuracy: the tensor core adds into its fp32 accumulator with truncation, one truncation per MMA.
// Left alone that is a bias of ~0.5 ulp(D) per instruction, ~1e-5 relative for a 3x3x64 reduction, and it
// compounds over the ~60 layer applications of a GGN product (measured 4e-4 on ResNet-18).  Therefore a
// TMEM accumulator only ever holds TC_FLUSH stages (64 reduction elements): the epilogue warps drain each
// chunk with tcgen05.ld and sum the chunks in registers with round-to-nearest fp32 adds, overlapped with
// the MMAs of the next chunk through the TMEM double buffer.
// Every mbarrier wait is bounded and traps instead of hanging the GPU.
#pragma once
#include "common.cuh"

namespace curv {

#ifndef CURV_DISABLE_TC

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;  // fp32 elements = one 128-byte swizzle row
constexpr int TC_THREADS = 544;   // 17 warps, see the role table below
constexpr int TC_FLUSH = 2;       // smem stages per TMEM accumulation chunk (see tc epilogue)
constexpr int TC_PRODUCERS = 256;

template <int BN>
struct TcCfg {
  static constexpr int STAGES = BN == 128 ? 3 : 4;
  static constexpr int A_BYTES = TC_BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = 2 * BN;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps (CUDA error) instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t i = 0; i < (1u << 26); ++i)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
}
__device__ __forceinline__ void fence_async_proxy() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// One lane of the (converged) warp, chosen by the hardware.  tcgen05.mma / commit are issued under this predicate
// from warp-uniform control flow: issued from inside `if (lane == 0)` nvcc wraps every UTCxMMA in an
// ELECT / BRA.U.ANY uniformisation loop and the single issuing thread becomes the bottleneck.
__device__ __forceinline__ bool tc_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);  // start address
  d |= (uint64_t)1 << 16;                      // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;            // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
  return d;
}
// MN-major tf32 operand tile.  For 32-bit MN-major operands the only legal swizzled layout is
// SWIZZLE_128B_BASE32B (CUTLASS sm100_common.inl:92): atoms of 4 reduction rows x 128 bytes (32
// MN-contiguous fp32), XOR of the 32-byte chunk index (address bits 5-6) with the row-in-atom (bits 7-8).
// A stage holds 32 reduction rows: row r of MN atom a lives at a*4096 + r*128 (K atoms 512 B apart).
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(4096 >> 4) << 16;  // leading byte offset: stride between 32-wide MN atoms
  d |= (uint64_t)(512 >> 4) << 32;   // stride byte offset: stride between 4-row K atoms
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;            // SWIZZLE_128B_BASE32B
  return d;
}
// instruction descriptor: D=f32, A=B=tf32, both K-major, M=128, N=BN
__host__ __device__ constexpr uint32_t make_tf32_idesc(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

__device__ __forceinline__ void split_tf32(float4 v, float4& hi, float4& lo) {
  auto h = [](float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); };
  hi = make_float4(h(v.x), h(v.y), h(v.z), h(v.w));
  lo = make_float4(h(v.x - hi.x), h(v.y - hi.y), h(v.z - hi.z), h(v.w - hi.w));
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// ---------------------------------------------------------------------------------------------------
// shared CTA prologue / epilogue pieces
// ---------------------------------------------------------------------------------------------------
template <int BN>
struct TcSmem {
  using Cfg = TcCfg<BN>;
  uint32_t base, bar_base;
  __device__ explicit TcSmem(uint8_t* raw) {
    base = (smem_u32(raw) + 1023u) & ~1023u;
    bar_base = base + Cfg::STAGES * Cfg::STAGE_BYTES;
  }
  __device__ uint32_t full(int s) const { return bar_base + 8u * s; }
  __device__ uint32_t empty(int s) const { return bar_base + 8u * (Cfg::STAGES + s); }
  __device__ uint32_t tfull(int a) const { return bar_base + 8u * (2 * Cfg::STAGES + a); }
  __device__ uint32_t tempty(int a) const { return bar_base + 8u * (2 * Cfg::STAGES + 2 + a); }
  __device__ uint32_t tmem_slot() const { return bar_base + 8u * (2 * Cfg::STAGES + 4); }
  __device__ uint32_t stageA(int s) const { return base + s * Cfg::STAGE_BYTES; }
  __device__ uint32_t stageB(int s) const { return base + s * Cfg::STAGE_BYTES + 2 * Cfg::A_BYTES; }
};

// barrier init + TMEM allocation; returns the TMEM base address.  full_count = arrivals per stage.
template <int BN>
__device__ __forceinline__ uint32_t tc_prologue(const TcSmem<BN>& S, uint8_t* raw, int full_count) {
  using Cfg = TcCfg<BN>;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(S.full(s), full_count); mbar_init(S.empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(S.tfull(a), 1); mbar_init(S.tempty(a), 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(S.tmem_slot()),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *reinterpret_cast<volatile uint32_t*>(raw + (S.tmem_slot() - smem_u32(raw)));
}
template <int BN>
__device__ __forceinline__ void tc_teardown(uint32_t tmem_base) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)TcCfg<BN>::TMEM_COLS)
                 : "memory");
  }
}

// MMA issuer loop body for one chunk of T stages (whole warp, one elected lane issues): 3xTF32 = 12 MMAs per
// 32-deep stage; the last stage also commits the chunk's tfull barrier.
// DESC(addr) builds the smem descriptor; K_ADV = descriptor start-address advance (16-byte units) per K=8.
template <int BN, bool MN_MAJOR>
__device__ __forceinline__ void tc_issue_tile(const TcSmem<BN>& S, uint32_t d_tmem, int T, int& stage,
                                              uint32_t& phase, uint32_t tfull_bar) {
  using Cfg = TcCfg<BN>;
  constexpr uint32_t idesc = make_tf32_idesc(BN) | (MN_MAJOR ? ((1u << 15) | (1u << 16)) : 0u);
  for (int it = 0; it < T; ++it) {
    mbar_wait(S.full(stage), phase);
    tc_fence_after();
    const uint32_t sA = S.stageA(stage), sB = S.stageB(stage);
    uint64_t dAh, dAl, dBh, dBl;
    if (MN_MAJOR) {
      dAh = make_mnmajor_sw128_desc(sA); dAl = make_mnmajor_sw128_desc(sA + Cfg::A_BYTES);
      dBh = make_mnmajor_sw128_desc(sB); dBl = make_mnmajor_sw128_desc(sB + Cfg::B_BYTES);
    } else {
      dAh = make_kmajor_sw128_desc(sA); dAl = make_kmajor_sw128_desc(sA + Cfg::A_BYTES);
      dBh = make_kmajor_sw128_desc(sB); dBl = make_kmajor_sw128_desc(sB + Cfg::B_BYTES);
    }
    if (tc_elect_one()) {
#pragma unroll
      for (int ks = 0; ks < TC_BK / 8; ++ks) {
        // K-major: +32 bytes inside the 128-byte row; MN-major: next 8-row group (+1024 bytes)
        const uint64_t adv = (uint64_t)((MN_MAJOR ? ks * 1024 : ks * 32) >> 4);
        tc_mma_tf32(d_tmem, dAl + adv, dBh + adv, idesc, (it | ks) != 0 ? 1u : 0u);
        tc_mma_tf32(d_tmem, dAh + adv, dBl + adv, idesc, 1u);
        tc_mma_tf32(d_tmem, dAh + adv, dBh + adv, idesc, 1u);
      }
      tc_commit(S.empty(stage));  // frees the smem stage when these MMAs retire
      if (it + 1 == T) tc_commit(tfull_bar);  // chunk complete -> epilogue
    }
    __syncwarp();
    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
  }
}

// ---------------------------------------------------------------------------------------------------
// gather GEMM (forward conv + tangents, dgrad):  out[m][n] = sum_seg sum_r gather(A)[m][r] * W[n][r]
// A: gathered by the producer warps (hi/lo split on the fly).  W: pre-split, pre-swizzled "UMMA image"
// (pack_umma_kmajor_kernel), one cp.async.bulk per stage straight into shared memory.
// ---------------------------------------------------------------------------------------------------
template <int BN, bool COAL>
__global__ void __launch_bounds__(TC_THREADS, 1) gather_gemm_tc(const GatherGemmArgs p, int nslots) {
  using Cfg = TcCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const TcSmem<BN> S(smem_raw);
  const Geom& g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = ceil_div(g.Nd, BN);
  const int tiles_m = ceil_div(g.M, TC_BM);
  const int ntiles = tiles_m * tiles_n * nslots;
  const int nchunks = ceil_div(g.Kd, TC_BK);
  const uint32_t tmem_base = tc_prologue<BN>(S, smem_raw, TC_PRODUCERS + 1);

  auto decode_tile = [&](int tile, int& slot, int& m0, int& tn) {
    int si = tile % nslots;
    int rest = tile / nslots;
    tn = rest % tiles_n;
    slot = p.slot0 + si; m0 = (rest / tiles_n) * TC_BM;
  };
  // segment s of a slot: activation base and weight image base
  auto segment = [&](int slot, int s, const float*& A, const float*& Wimg) {
    const bool first_is_act = (slot == 0) || p.a_has_slots;
    if (s == 0 && first_is_act) {
      A = p.A + (long long)slot * p.A_slot; Wimg = p.W_img;
    } else {
      A = p.A; Wimg = p.Wt_img + (long long)(slot - 1) * p.Wt_img_slot;
    }
  };
  auto num_segments = [&](int slot) {
    return slot == 0 ? 1 : (p.a_has_slots ? 1 : 0) + (p.Wt_img != nullptr ? 1 : 0);
  };

  if (warp >= 5 && warp < 13) {
    // ------------------------------------------------------------------ producers
    const int pt = threadIdx.x - 5 * 32;  // 0..255
    // Thread -> data mapping (template COAL):
    //  COAL: 8 consecutive lanes read the 8 16-byte chunks of ONE 128-byte K row (a warp request = 4 rows =
    //        4 cache lines instead of 32 sectors in 16-32 lines); thread -> chunk a_c of rows a_r0 + 32 i.
    //        Addresses are LINEAR in (row, tap): element offset = rowoff_i + tapoff with
    //        rowoff_i = ((b*Hs + ah_i)*Ws + aw_i)*Cs fixed per tile and tapoff = +-(kh*Ws + kw)*Cs + c per
    //        stage (uniform over rows), so a load costs one add and a bounds test on the packed (ah, aw).
    //        Strided dgrad ((dest + pad - tap)/stride) is not linear and takes the generic path.
    //  !COAL: two threads per row, 4 consecutive chunks (64 B) each (kept for A/B measurements).
    constexpr int NR = COAL ? 4 : 1;               // rows per thread
    const int a_c = pt & 7, a_r0 = pt >> 3;         // COAL
    const int a_row = pt >> 1, a_c0 = (pt & 1) * 4; // !COAL
    const uint32_t a_off = COAL ? (uint32_t)((a_r0 >> 3) * 1024 + (a_r0 & 7) * 128 + ((a_c ^ (a_r0 & 7)) << 4))
                                : (uint32_t)((a_row >> 3) * 1024 + (a_row & 7) * 128);
    const bool fast = (g.Cs % TC_BK) == 0;  // a 128-byte K row never straddles two filter taps
    const bool linear = g.mode == 0 || (g.sh == 1 && g.sw == 1);
    // iteration state
    int tile = blockIdx.x, slot = 0, m0 = 0, tn = 0, nseg = 0, seg = 0, kc = 0;
    int kh = 0, kw = 0, cb = 0;
    bool m_ok[NR];
    int ah[NR], aw[NR];
    int rowoff[NR];  // element offset of (b, ah, aw, channel 0) inside the slot (fits 32 bits: checked on the host)
    const float* Ap = nullptr;
    const float* Wimg = nullptr;
    auto enter_tile = [&]() {
      decode_tile(tile, slot, m0, tn);
      nseg = num_segments(slot);
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const int m = m0 + (COAL ? a_r0 + 32 * i : a_row);
        m_ok[i] = m < g.M;
        const int mm = m_ok[i] ? m : 0;
        const int bimg = mm / (g.Hd * g.Wd);
        const int rem = mm - bimg * (g.Hd * g.Wd);
        const int hd = rem / g.Wd, wd = rem - hd * g.Wd;
        ah[i] = g.mode == 0 ? hd * g.sh - g.ph : hd + g.ph;
        aw[i] = g.mode == 0 ? wd * g.sw - g.pw : wd + g.pw;
        rowoff[i] = linear ? ((bimg * g.Hs + ah[i]) * g.Ws + aw[i]) * g.Cs : bimg * g.Hs * g.Ws;
        if (!m_ok[i]) ah[i] = -(1 << 20);  // fails every bounds test
      }
      seg = 0; kc = 0; kh = 0; kw = 0; cb = 0;
      segment(slot, 0, Ap, Wimg);
    };
    // generic (non-linear) source pixel of row i for tap (kh_, kw_)
    auto source_pixel = [&](int i, int kh_, int kw_, int& hs, int& ws) -> bool {
      bool ok = m_ok[i];
      if (g.mode == 0) { hs = ah[i] + kh_; ws = aw[i] + kw_; }
      else {
        const int th = ah[i] - kh_, tw = aw[i] - kw_;
        ok = ok && th >= 0 && tw >= 0;
        hs = th / g.sh; ws = tw / g.sw;
        ok = ok && hs * g.sh == th && ws * g.sw == tw;
      }
      return ok && hs >= 0 && hs < g.Hs && ws >= 0 && ws < g.Ws;
    };
    auto issue = [&](float4 (&v)[4]) {
      if (p.debug & 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = make_float4(1.f, 1.f, 1.f, 1.f);
        return;
      }
      if (COAL) {
        int kh_ = kh, kw_ = kw, c = cb + a_c * 4;
        bool rok = true;
        if (!fast) {  // the 16-byte chunk decides its own filter tap (C_in not a multiple of 32: the stem)
          const int r = kc * TC_BK + a_c * 4;
          const int tap = r / g.Cs;
          c = r - tap * g.Cs;
          kh_ = tap / g.KW; kw_ = tap - kh_ * g.KW;
          rok = r < g.Kd;
        }
        if (linear) {
          const int sgn = g.mode == 0 ? 1 : -1;
          const int tapoff = sgn * (kh_ * g.Ws + kw_) * g.Cs + c;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int hs = ah[i % NR] + sgn * kh_, ws = aw[i % NR] + sgn * kw_;
            const bool ok = rok && (unsigned)hs < (unsigned)g.Hs && (unsigned)ws < (unsigned)g.Ws;
            v[i] = ok ? __ldg(reinterpret_cast<const float4*>(Ap + (rowoff[i % NR] + tapoff)))
                      : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            int hs, ws;
            const bool ok = rok && source_pixel(i % NR, kh_, kw_, hs, ws);
            v[i] = ok ? __ldg(reinterpret_cast<const float4*>(
                            Ap + ((long long)(rowoff[i % NR] + hs * g.Ws + ws) * g.Cs + c)))
                      : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      } else if (fast) {
        int hs, ws;
        const bool ok = source_pixel(0, kh, kw, hs, ws);
        const long long pix = linear ? (long long)(rowoff[0] / g.Cs) - (long long)(ah[0] * g.Ws + aw[0]) : rowoff[0];
        const float* rowp = Ap + ((pix + (long long)hs * g.Ws + ws) * g.Cs + cb + a_c0 * 4);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          v[j] = ok ? __ldg(reinterpret_cast<const float4*>(rowp) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        const long long pix = linear ? (long long)(rowoff[0] / g.Cs) - (long long)(ah[0] * g.Ws + aw[0]) : rowoff[0];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = kc * TC_BK + (a_c0 + j) * 4;
          const int tap = r / g.Cs;
          const int c = r - tap * g.Cs;
          const int kh_ = tap / g.KW, kw_ = tap - kh_ * g.KW;
          int hs, ws;
          const bool ok = r < g.Kd && source_pixel(0, kh_, kw_, hs, ws);
          v[j] = ok ? __ldg(reinterpret_cast<const float4*>(Ap + ((pix + (long long)hs * g.Ws + ws) * g.Cs + c)))
                    : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    // advance to the next (tile, seg, kc); returns false when this CTA is done
    auto advance = [&]() -> bool {
      ++kc;
      cb += TC_BK;
      if (cb >= g.Cs) { cb = 0; if (++kw == g.KW) { kw = 0; ++kh; } }
      if (kc < nchunks) return true;
      kc = 0; kh = 0; kw = 0; cb = 0;
      if (++seg < nseg) { segment(slot, seg, Ap, Wimg); return true; }
      tile += gridDim.x;
      if (tile >= ntiles) return false;
      enter_tile();
      return true;
    };

    // Software pipeline, TWO iterations deep in registers: the producers are bound by global-load latency
    // (ncu: long-scoreboard stalls dominate), one stage of lookahead is shorter than an L2 miss.
    auto wblock = [&]() { return Wimg + ((long long)tn * nchunks + kc) * (2 * BN * TC_BK); };
    int stage = 0;
    uint32_t phase = 0;
    float4 q0[4], q1[4], q2[4];
    const float *w0 = nullptr, *w1 = nullptr, *w2 = nullptr;
    bool h0 = tile < ntiles, h1 = false, h2 = false;
    if (h0) {
      enter_tile();
      w0 = wblock(); issue(q0);
      h1 = advance();
      if (h1) { w1 = wblock(); issue(q1); }
    }
    while (h0) {
      h2 = h1 ? advance() : false;
      if (h2) { w2 = wblock(); issue(q2); }
      mbar_wait(S.empty(stage), phase ^ 1);
      const uint32_t sA = S.stageA(stage);
      if (pt == 0) {
        const uint32_t bytes = 2 * Cfg::B_BYTES;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(S.full(stage)),
                     "r"(bytes)
                     : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                S.stageB(stage)),
            "l"(w0), "r"(bytes), "r"(S.full(stage))
            : "memory");
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (p.debug & 2) continue;
        float4 hi, lo;
        split_tf32(q0[i], hi, lo);
        const uint32_t o = COAL ? a_off + (uint32_t)(i * 4096)
                                : a_off + (uint32_t)(((a_c0 + i) ^ (a_row & 7)) << 4);
        sts128(sA + o, hi);
        sts128(sA + Cfg::A_BYTES + o, lo);
      }
      fence_async_proxy();
      mbar_arrive(S.full(stage));
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
#pragma unroll
      for (int i = 0; i < 4; ++i) { q0[i] = q1[i]; q1[i] = q2[i]; }
      w0 = w1; w1 = w2; h0 = h1; h1 = h2;
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ MMA issuer
    {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int slot, m0, tn;
        decode_tile(tile, slot, m0, tn);
        const int T = num_segments(slot) * nchunks;
        for (int t0 = 0; t0 < T; t0 += TC_FLUSH) {  // one TMEM accumulation chunk
          mbar_wait(S.tempty(acc), acc_phase ^ 1);
          tc_fence_after();
          tc_issue_tile<BN, false>(S, tmem_base + (uint32_t)(acc * BN), min(TC_FLUSH, T - t0), stage, phase,
                                   S.tfull(acc));  // chunk complete -> epilogue
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 0-3, 13-16)
    constexpr int HALF = BN / 2;
    const int egrp = warp >= 13 ? 1 : 0;     // which half of the accumulator columns
    const int quad = warp & 3;               // TMEM lane quadrant this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int slot, m0, tn;
      decode_tile(tile, slot, m0, tn);
      const int n0 = tn * BN + egrp * HALF;
      const int T = num_segments(slot) * nchunks;
      float accv[HALF];
#pragma unroll
      for (int j = 0; j < HALF; ++j) accv[j] = 0.f;
      for (int t0 = 0; t0 < T; t0 += TC_FLUSH) {
        mbar_wait(S.tfull(acc), acc_phase);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < HALF; c0 += 16) {
          uint32_t r[16];
          tc_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + egrp * HALF + c0), r);
#pragma unroll
          for (int j = 0; j < 16; ++j) accv[c0 + j] += __uint_as_float(r[j]);
        }
        tc_fence_before();
        mbar_arrive(S.tempty(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      const float* bias = (slot == 0) ? p.bias
                                      : (p.bias_t ? p.bias_t + (long long)(slot - 1) * p.bias_slot : nullptr);
      float* outp = p.out + (long long)slot * p.out_slot;
      const int m = m0 + quad * 32 + lane;
      if (m < g.M) {
#pragma unroll
        for (int j = 0; j < HALF / 4; ++j) {
          const int n = n0 + j * 4;
          if (n >= g.Nd) continue;
          float4 v = make_float4(accv[j * 4 + 0], accv[j * 4 + 1], accv[j * 4 + 2], accv[j * 4 + 3]);
          if (bias) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + n));
            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
          }
          float4* dst = reinterpret_cast<float4*>(outp + (long long)m * g.Nd + n);
          if (p.accumulate) {
            const float4 o = *dst;
            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
          }
          *dst = v;
        }
      }
    }
  }
  tc_teardown<BN>(tmem_base);
}

// Pre-split, pre-swizzled weight image for gather_gemm_tc:
//   block (tn, kc) = [hi plane: BN rows x 128 B, 128B-swizzled][lo plane], blocks ordered tn-major.
// src: [N][Kd] row-major (the SIMT-layout packed weights).  grid.y = slot.
__global__ void pack_umma_kmajor_kernel(const float* __restrict__ src, long long src_slot,
                                        float* __restrict__ dst, long long dst_slot, int N, int Kd, int BN,
                                        int tiles_n, int nchunks) {
  src += blockIdx.y * src_slot;
  dst += blockIdx.y * dst_slot;
  const long long total = (long long)tiles_n * nchunks * BN * 8;  // 16-byte chunks per plane
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e & 7);
    long long t = e >> 3;
    const int r = (int)(t % BN); t /= BN;
    const int kc = (int)(t % nchunks);
    const int tn = (int)(t / nchunks);
    const int n = tn * BN + r, k = kc * TC_BK + c * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N && k < Kd) v = __ldg(reinterpret_cast<const float4*>(src + (long long)n * Kd + k));
    float4 hi, lo;
    split_tf32(v, hi, lo);
    float* blk = dst + ((long long)tn * nchunks + kc) * (2 * BN * TC_BK);
    const int o = ((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) >> 2;
    *reinterpret_cast<float4*>(blk + o) = hi;
    *reinterpret_cast<float4*>(blk + BN * TC_BK + o) = lo;
  }
}

// ---------------------------------------------------------------------------------------------------
// wgrad GEMM on tcgen05:  D[i][j] = sum_m P[m][i] * Q[m][j]  over a split of the pixel range,
// both operands MN-major (the reduction index m = pixel is the smem row).
//   swap == 0:  P = G (i = output channel n, 128 per tile),  Q = gathered input (j = tap*Cs + c)
//   swap == 1:  P = gathered input (i = tap*Cs + c),         Q = G (j = n)      [for C_out <= 64]
// partial[(split*nslots + slot_idx)][n][tap*Cs + c]
// ---------------------------------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1) wgrad_gemm_tc(const WgradArgs p, int swap) {
  using Cfg = TcCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const TcSmem<BN> S(smem_raw);
  const Geom& g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile grid: i over (channels | Kd) in 128s, j over (Kd | channels) in BNs
  const int ext_i = swap ? g.Kd : g.N;
  const int ext_j = swap ? p.Ng : g.Kd;
  const int tiles_i = ceil_div(ext_i, TC_BM), tiles_j = ceil_div(ext_j, BN);
  const int ntiles = tiles_i * tiles_j * p.nslots * p.nsplit;
  const uint32_t tmem_base = tc_prologue<BN>(S, smem_raw, TC_PRODUCERS);

  auto decode_tile = [&](int tile, int& slot_idx, int& split, int& i0, int& j0) {
    slot_idx = tile % p.nslots;
    int rest = tile / p.nslots;
    const int tj = rest % tiles_j; rest /= tiles_j;
    const int ti = rest % tiles_i;
    split = rest / tiles_i;
    i0 = ti * TC_BM; j0 = tj * BN;
  };
  auto num_segments = [&](int slot) { return (p.second_seg && slot > 0) ? 2 : 1; };
  auto stages_of = [&](int split) {
    const int mb = split * p.m_per_split;
    const int me = min(g.M, mb + p.m_per_split);
    return ceil_div(max(0, me - mb), TC_BK);
  };

  if (warp >= 5 && warp < 13) {
    // ------------------------------------------------------------------ producers
    const int pt = threadIdx.x - 5 * 32;
    const int krow = pt >> 3;            // pixel row of the 32-pixel stage
    const int cg = pt & 7;               // chunk group: chunks cg*4 .. cg*4+3 of a 128-wide operand
    // channel-operand (G) and gather-operand (In) extents in this tile
    constexpr int GCH = 4;               // chunks per thread for a 128-wide operand
    // smem offsets of this thread's chunks inside an MN-major plane
    auto smem_off = [&](int c) {         // c = 16-byte chunk index along MN (0..31)
      const int c8 = c & 7;
      const int csw = ((((c8 >> 1) ^ (krow & 3)) << 1) | (c8 & 1));  // Swizzle<2,5,2>
      return (uint32_t)((c >> 3) * 4096 + krow * 128 + (csw << 4));
    };
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int slot_idx, split, i0, j0;
      decode_tile(tile, slot_idx, split, i0, j0);
      const int slot = p.slot0 + slot_idx;
      const int mb = split * p.m_per_split;
      const int me = min(g.M, mb + p.m_per_split);
      const int nst = stages_of(split);
      const int nseg = num_segments(slot);
      // G operand: channel offset / width of this tile;  In operand: reduction-column offset
      const int ch0 = swap ? j0 : i0;          // first channel
      const int chw = swap ? BN : TC_BM;       // channels in the tile (operand width)
      const int col0 = swap ? i0 : j0;         // first tap*Cs+c column
      const int colw = swap ? TC_BM : BN;
      const uint32_t planeG = swap ? 2 * Cfg::A_BYTES : 0;     // G goes to the B planes when swapped
      const uint32_t planeI = swap ? 0 : 2 * Cfg::A_BYTES;
      const uint32_t loG = swap ? Cfg::B_BYTES : Cfg::A_BYTES;
      const uint32_t loI = swap ? Cfg::A_BYTES : Cfg::B_BYTES;
      // gather-operand chunk descriptors of this thread (fixed for the tile)
      int ikh[GCH], ikw[GCH], ic[GCH];
      bool iok[GCH];
#pragma unroll
      for (int q = 0; q < GCH; ++q) {
        const int c = cg * 4 + q;                  // chunk index in the operand (0..31)
        const int col = col0 + c * 4;
        iok[q] = (c * 4 < colw) && col < g.Kd;
        const int tap = iok[q] ? col / g.Cs : 0;
        ic[q] = col - tap * g.Cs;
        ikh[q] = tap / g.KW; ikw[q] = tap - ikh[q] * g.KW;
      }
      for (int seg = 0; seg < nseg; ++seg) {
        const float* Gp = (seg == 0) ? p.G + (long long)slot * p.G_slot : p.G;
        const float* Ip = (seg == 0) ? p.In : p.In + (long long)slot * p.In_slot;
        for (int st = 0; st < nst; ++st) {
          const int m = mb + st * TC_BK + krow;
          const bool mok = m < me;
          float4 vg[GCH], vi[GCH];
          // G chunks (pure streaming)
#pragma unroll
          for (int q = 0; q < GCH; ++q) {
            const int c = cg * 4 + q;
            const int ch = ch0 + c * 4;
            const bool ok = mok && (c * 4 < chw) && ch < p.Ng;
            vg[q] = ok ? __ldg(reinterpret_cast<const float4*>(Gp + (long long)m * p.Ng + ch))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          // gathered input chunks
          const int mm = mok ? m : 0;
          const int bimg = mm / (g.Hd * g.Wd);
          const int rem = mm - bimg * (g.Hd * g.Wd);
          const int hd = rem / g.Wd, wd = rem - hd * g.Wd;
          const int h0 = hd * g.sh - g.ph, w0 = wd * g.sw - g.pw;
#pragma unroll
          for (int q = 0; q < GCH; ++q) {
            const int hs = h0 + ikh[q], ws = w0 + ikw[q];
            const bool ok = mok && iok[q] && hs >= 0 && hs < g.Hs && ws >= 0 && ws < g.Ws;
            vi[q] = ok ? __ldg(reinterpret_cast<const float4*>(
                             Ip + ((((long long)bimg * g.Hs + hs) * g.Ws + ws) * g.Cs + ic[q])))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          mbar_wait(S.empty(stage), phase ^ 1);
          const uint32_t sbase = S.stageA(stage);
#pragma unroll
          for (int q = 0; q < GCH; ++q) {
            const int c = cg * 4 + q;
            float4 hi, lo;
            if (c * 4 < chw) {
              split_tf32(vg[q], hi, lo);
              const uint32_t o = smem_off(c);
              sts128(sbase + planeG + o, hi);
              sts128(sbase + planeG + loG + o, lo);
            }
            if (c * 4 < colw) {
              split_tf32(vi[q], hi, lo);
              const uint32_t o = smem_off(c);
              sts128(sbase + planeI + o, hi);
              sts128(sbase + planeI + loI + o, lo);
            }
          }
          fence_async_proxy();
          mbar_arrive(S.full(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 4) {
    {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int slot_idx, split, i0, j0;
        decode_tile(tile, slot_idx, split, i0, j0);
        const int T = num_segments(p.slot0 + slot_idx) * stages_of(split);
        for (int t0 = 0; t0 < T; t0 += TC_FLUSH) {
          mbar_wait(S.tempty(acc), acc_phase ^ 1);
          tc_fence_after();
          tc_issue_tile<BN, true>(S, tmem_base + (uint32_t)(acc * BN), min(TC_FLUSH, T - t0), stage, phase,
                                   S.tfull(acc));
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 0-3, 13-16)
    constexpr int HALF = BN / 2;
    const int egrp = warp >= 13 ? 1 : 0;
    const int quad = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int slot_idx, split, i0, j0;
      decode_tile(tile, slot_idx, split, i0, j0);
      const int T = num_segments(p.slot0 + slot_idx) * stages_of(split);
      float accv[HALF];
#pragma unroll
      for (int j = 0; j < HALF; ++j) accv[j] = 0.f;
      for (int t0 = 0; t0 < T; t0 += TC_FLUSH) {
        mbar_wait(S.tfull(acc), acc_phase);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < HALF; c0 += 16) {
          uint32_t r[16];
          tc_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + egrp * HALF + c0), r);
#pragma unroll
          for (int j = 0; j < 16; ++j) accv[c0 + j] += __uint_as_float(r[j]);
        }
        tc_fence_before();
        mbar_arrive(S.tempty(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      float* outp = p.partial + ((long long)split * p.nslots + slot_idx) * (long long)g.N * g.Kd;
      const int i = i0 + quad * 32 + lane;
      const int jbase = j0 + egrp * HALF;
      if (!swap) {
        if (i < g.N) {
#pragma unroll
          for (int j = 0; j < HALF / 4; ++j) {
            const int col = jbase + j * 4;
            if (col >= g.Kd) continue;
            *reinterpret_cast<float4*>(outp + (long long)i * g.Kd + col) =
                make_float4(accv[j * 4 + 0], accv[j * 4 + 1], accv[j * 4 + 2], accv[j * 4 + 3]);
          }
        }
      } else {
        if (i < g.Kd) {
#pragma unroll
          for (int j = 0; j < HALF; ++j) {
            const int n = jbase + j;
            if (n < g.N) outp[(long long)n * g.Kd + i] = accv[j];
          }
        }
      }
    }
  }
  tc_teardown<BN>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------
// Multi-slot wgrad (the GGN/VJP case: one segment):  D_k[i][n] = sum_m In[m][i] * G_k[m][n]  for ALL slots k
// of a tile at once.  The gathered-input operand (128 columns of tap*Cs+c, the expensive im2col gather) is
// staged ONCE per 16-pixel stage and multiplied against the 64-channel G tile of every slot: the NS <= 8
// accumulators [128 x 64] fill the 512 TMEM columns exactly.  Operand traffic per FLOP drops 2.5x versus one
// slot per tile (the previous kernel was bound by operand re-reads: 21 FLOP per loaded byte).
//   smem stage (80 KB, 2 stages): In hi/lo 2 x 8 KB, then per slot G hi/lo 2 x 4 KB
//   MN-major SWIZZLE_128B_BASE32B: chunk c (16 B) of pixel row r at (c>>3)*2048 + r*128 + swz*16
// TMEM is single-buffered (all 512 columns hold accumulators), so draining it stalls the MMAs for a few
// microseconds; the accumulation is therefore chunked coarsely: every MS_FLUSH = 256 stages (4096 pixels,
// 1536 MMAs per accumulator) the epilogue adds the chunk into the split's partial in global memory with
// round-to-nearest adds (same thread, same address, fixed order).  That bounds the tensor core's
// truncation bias (one ~0.5 ulp truncation per MMA; unchunked it reached 9e-5 at B=128).
// ---------------------------------------------------------------------------------------------------
constexpr int MS_ROWS = 16;                                  // pixels per stage
constexpr int MS_A_BYTES = 128 * MS_ROWS * 4;                // 8 KB per plane
constexpr int MS_B_BYTES = 64 * MS_ROWS * 4;                 // 4 KB per plane and slot
constexpr int MS_STAGE_BYTES = 2 * MS_A_BYTES + 8 * 2 * MS_B_BYTES;  // 80 KB
constexpr int MS_STAGES = 2;
constexpr int MS_FLUSH = 256;                                // stages per TMEM accumulation chunk (2048 pixels)
constexpr int MS_SMEM_BYTES = MS_STAGES * MS_STAGE_BYTES + 1024 + 256;

__device__ __forceinline__ uint64_t make_mnmajor_b32_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;  // stride between 32-wide MN atoms
  d |= (uint64_t)(512 >> 4) << 32;        // stride between 4-row K atoms
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                 // SWIZZLE_128B_BASE32B
  return d;
}

__global__ void __launch_bounds__(TC_THREADS, 1) wgrad_gemm_tc_ms(const WgradArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = sbase + MS_STAGES * MS_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MS_STAGES + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * MS_STAGES);
  const uint32_t tempty_bar = bar_base + 8u * (2 * MS_STAGES + 1);
  const uint32_t tmem_slot = bar_base + 8u * (2 * MS_STAGES + 2);

  const Geom& g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NS = p.nslots;  // 1..8 slots, all resident in TMEM
  const int tiles_i = ceil_div(g.Kd, TC_BM), tiles_j = ceil_div(p.Ng, 64);
  const int ntiles = tiles_i * tiles_j * p.nsplit;

  if (threadIdx.x == 0) {
    for (int s = 0; s < MS_STAGES; ++s) { mbar_init(full_bar(s), TC_PRODUCERS); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  auto decode_tile = [&](int tile, int& split, int& i0, int& j0) {
    const int tj = tile % tiles_j;
    int rest = tile / tiles_j;
    const int ti = rest % tiles_i;
    split = rest / tiles_i;
    i0 = ti * TC_BM; j0 = tj * 64;
  };
  auto stages_of = [&](int split) {
    const int mb = split * p.m_per_split;
    const int me = min(g.M, mb + p.m_per_split);
    return ceil_div(max(0, me - mb), MS_ROWS);
  };

  if (warp >= 5 && warp < 13) {
    // ------------------------------------------------------------------ producers
    const int pt = threadIdx.x - 5 * 32;
    const int krow = pt >> 4;            // pixel row of the 16-pixel stage
    const int cq = pt & 15;              // G: chunk cq (16 chunks = 64 channels); In: chunks 2cq, 2cq+1
    auto smem_off = [&](int c) {         // c = 16-byte chunk index along MN
      const int c8 = c & 7;
      const int csw = ((((c8 >> 1) ^ (krow & 3)) << 1) | (c8 & 1));  // Swizzle<2,5,2>
      return (uint32_t)((c >> 3) * 2048 + krow * 128 + (csw << 4));
    };
    const uint32_t offI0 = smem_off(2 * cq), offI1 = smem_off(2 * cq + 1), offG = smem_off(cq);
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int split, i0, j0;
      decode_tile(tile, split, i0, j0);
      const int mb = split * p.m_per_split;
      const int me = min(g.M, mb + p.m_per_split);
      const int nst = stages_of(split);
      // gathered-input chunks of this thread: columns i0 + (2cq + q)*4
      int ikh[2], ikw[2], ic[2];
      bool iok[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int col = i0 + (2 * cq + q) * 4;
        iok[q] = col < g.Kd;
        const int tap = iok[q] ? col / g.Cs : 0;
        ic[q] = col - tap * g.Cs;
        ikh[q] = tap / g.KW; ikw[q] = tap - ikh[q] * g.KW;
      }
      const int ch = j0 + cq * 4;
      const bool chok = ch < p.Ng;
      for (int st = 0; st < nst; ++st) {
        const int m = mb + st * MS_ROWS + krow;
        const bool mok = m < me;
        float4 vi[2], vg[8];
        const int mm = mok ? m : 0;
        const int bimg = mm / (g.Hd * g.Wd);
        const int rem = mm - bimg * (g.Hd * g.Wd);
        const int hd = rem / g.Wd, wd = rem - hd * g.Wd;
        const int h0 = hd * g.sh - g.ph, w0 = wd * g.sw - g.pw;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int hs = h0 + ikh[q], ws = w0 + ikw[q];
          const bool ok = mok && iok[q] && hs >= 0 && hs < g.Hs && ws >= 0 && ws < g.Ws;
          vi[q] = ok ? __ldg(reinterpret_cast<const float4*>(
                           p.In + ((((long long)bimg * g.Hs + hs) * g.Ws + ws) * g.Cs + ic[q])))
                     : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float* grow = p.G + (long long)p.slot0 * p.G_slot + (long long)mm * p.Ng + ch;
#pragma unroll
        for (int s = 0; s < 8; ++s)
          vg[s] = (s < NS && mok && chok) ? __ldg(reinterpret_cast<const float4*>(grow + (long long)s * p.G_slot))
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
        mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t sA = sbase + stage * MS_STAGE_BYTES;
        const uint32_t sB = sA + 2 * MS_A_BYTES;
        float4 hi, lo;
        split_tf32(vi[0], hi, lo);
        sts128(sA + offI0, hi); sts128(sA + MS_A_BYTES + offI0, lo);
        split_tf32(vi[1], hi, lo);
        sts128(sA + offI1, hi); sts128(sA + MS_A_BYTES + offI1, lo);
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          if (s < NS) {
            split_tf32(vg[s], hi, lo);
            sts128(sB + s * 2 * MS_B_BYTES + offG, hi);
            sts128(sB + s * 2 * MS_B_BYTES + MS_B_BYTES + offG, lo);
          }
        }
        fence_async_proxy();
        mbar_arrive(full_bar(stage));
        if (++stage == MS_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_tf32_idesc(64) | (1u << 15) | (1u << 16);
      int stage = 0;
      uint32_t phase = 0, tphase = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int split, i0, j0;
        decode_tile(tile, split, i0, j0);
        const int nst = stages_of(split);
        // accumulation chunks of MS_FLUSH stages: bounded truncation chains (see header), an empty tile
        // still signals once so that the epilogue writes its zeros
        for (int c0 = 0; c0 < max(nst, 1); c0 += MS_FLUSH) {
          mbar_wait(tempty_bar, tphase ^ 1);  // previous chunk / tile drained
          tc_fence_after();
          const int cend = min(nst, c0 + MS_FLUSH);
          for (int st = c0; st < cend; ++st) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sA = sbase + stage * MS_STAGE_BYTES;
            const uint32_t sB = sA + 2 * MS_A_BYTES;
            const uint64_t dAh = make_mnmajor_b32_desc(sA, 2048), dAl = make_mnmajor_b32_desc(sA + MS_A_BYTES, 2048);
#pragma unroll
            for (int ks = 0; ks < MS_ROWS / 8; ++ks) {
              const uint64_t adv = (uint64_t)((ks * 1024) >> 4);
              for (int s = 0; s < NS; ++s) {
                const uint64_t dBh = make_mnmajor_b32_desc(sB + s * 2 * MS_B_BYTES, 2048);
                const uint64_t dBl = make_mnmajor_b32_desc(sB + s * 2 * MS_B_BYTES + MS_B_BYTES, 2048);
                const uint32_t d = tmem_base + (uint32_t)(s * 64);
                tc_mma_tf32(d, dAl + adv, dBh + adv, idesc, ((st - c0) | ks) != 0 ? 1u : 0u);
                tc_mma_tf32(d, dAh + adv, dBl + adv, idesc, 1u);
                tc_mma_tf32(d, dAh + adv, dBh + adv, idesc, 1u);
              }
            }
            tc_commit(empty_bar(stage));
            if (++stage == MS_STAGES) { stage = 0; phase ^= 1; }
          }
          tc_commit(tfull_bar);
          tphase ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 0-3, 13-16)
    const int egrp = warp >= 13 ? 1 : 0;  // slots with (s & 1) == egrp
    const int quad = warp & 3;
    uint32_t tphase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int split, i0, j0;
      decode_tile(tile, split, i0, j0);
      const int nst = stages_of(split);
      const int i = i0 + quad * 32 + lane;
      for (int c0 = 0; c0 < max(nst, 1); c0 += MS_FLUSH) {
        mbar_wait(tfull_bar, tphase);
        tc_fence_after();
        for (int s = egrp; s < NS; s += 2) {
          float* outp = p.partial + ((long long)split * p.nslots + s) * (long long)g.N * g.Kd;
#pragma unroll 1
          for (int cc = 0; cc < 64; cc += 16) {
            uint32_t r[16];
            tc_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(s * 64 + cc), r);
            if (i < g.Kd) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int n = j0 + cc + j;
                if (n < g.N) {
                  float* dst = outp + (long long)n * g.Kd + i;
                  float v = nst == 0 ? 0.f : __uint_as_float(r[j]);
                  if (c0 > 0) v += *dst;  // same thread wrote it in the previous chunk: fixed order
                  *dst = v;
                }
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(tempty_bar);
        tphase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static inline int tc_bn(int width) { return width > 64 ? 128 : 64; }

static inline bool tc_gather_eligible(const Geom& g, int mode) {
  if ((long long)g.B * g.Hs * g.Ws * g.Cs >= (1LL << 31) - (1LL << 24)) return false;  // 32-bit row offsets
  if (mode >= 2) return true;  // forced (tests): every shape is legal, small ones just waste tiles
  // big enough to fill 128-row tiles; everything else stays on the SIMT kernels
  return g.M >= 1024 && g.Kd >= 32 && g.Nd >= 16;
}
static inline bool tc_wgrad_eligible(const Geom& g, int mode) {
  if (mode >= 2) return true;
  return g.M >= 2048 && g.Kd >= 32 && g.N >= 16;
}
// size (floats) of the UMMA weight image of an [N][Kd] matrix
static inline long long tc_image_elems(int N, int Nd, int Kd) {
  const int BN = tc_bn(Nd);
  return (long long)ceil_div(Nd, BN) * ceil_div(Kd, TC_BK) * (2 * BN * TC_BK);
}

static int tc_sm_count() {
  static int sm_count = -1;
  if (sm_count == -1) {
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
    sm_count = prop.major == 10 ? prop.multiProcessorCount : 0;  // tcgen05 needs sm_100
    if (sm_count > 0) {
      bool ok = cudaFuncSetAttribute(gather_gemm_tc<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     TcCfg<128>::SMEM_BYTES) == cudaSuccess;
      ok = ok && cudaFuncSetAttribute(gather_gemm_tc<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      TcCfg<64>::SMEM_BYTES) == cudaSuccess;
      ok = ok && cudaFuncSetAttribute(gather_gemm_tc<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      TcCfg<128>::SMEM_BYTES) == cudaSuccess;
      ok = ok && cudaFuncSetAttribute(gather_gemm_tc<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      TcCfg<64>::SMEM_BYTES) == cudaSuccess;
      ok = ok && cudaFuncSetAttribute(wgrad_gemm_tc<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      TcCfg<128>::SMEM_BYTES) == cudaSuccess;
      ok = ok && cudaFuncSetAttribute(wgrad_gemm_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      TcCfg<64>::SMEM_BYTES) == cudaSuccess;
      ok = ok && cudaFuncSetAttribute(wgrad_gemm_tc_ms, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      MS_SMEM_BYTES) == cudaSuccess;
      if (!ok) sm_count = 0;
    }
  }
  return sm_count;
}

// pack [N][Kd] (slots along grid.y) into the UMMA image layout used by gather_gemm_tc
static inline int tc_pack_image(const float* src, long long src_slot, float* dst, long long dst_slot, int N,
                                int Nd, int Kd, int nslots, cudaStream_t st) {
  const int BN = tc_bn(Nd);
  const int tiles_n = ceil_div(Nd, BN), nchunks = ceil_div(Kd, TC_BK);
  long long total = (long long)tiles_n * nchunks * BN * 8;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_umma_kmajor_kernel<<<dim3(blocks, nslots), 256, 0, st>>>(src, src_slot, dst, dst_slot, N, Kd, BN,
                                                              tiles_n, nchunks);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// returns 0 on success, >0 on a CUDA error, <0 if the problem should use the SIMT path
static inline int tc_launch_gather_gemm(const GatherGemmArgs& a, int nslots, cudaStream_t st, bool coal) {
  const int sms = tc_sm_count();
  if (sms <= 0 || a.W_img == nullptr) return -1;
  const Geom& g = a.g;
  if (tc_bn(g.Nd) == 128) {
    int ntiles = ceil_div(g.M, TC_BM) * ceil_div(g.Nd, 128) * nslots;
    int grid = ntiles < sms ? ntiles : sms;
    if (coal) gather_gemm_tc<128, true><<<grid, TC_THREADS, TcCfg<128>::SMEM_BYTES, st>>>(a, nslots);
    else gather_gemm_tc<128, false><<<grid, TC_THREADS, TcCfg<128>::SMEM_BYTES, st>>>(a, nslots);
  } else {
    int ntiles = ceil_div(g.M, TC_BM) * ceil_div(g.Nd, 64) * nslots;
    int grid = ntiles < sms ? ntiles : sms;
    if (coal) gather_gemm_tc<64, true><<<grid, TC_THREADS, TcCfg<64>::SMEM_BYTES, st>>>(a, nslots);
    else gather_gemm_tc<64, false><<<grid, TC_THREADS, TcCfg<64>::SMEM_BYTES, st>>>(a, nslots);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

static inline int tc_launch_wgrad(const WgradArgs& a, cudaStream_t st) {
  const int sms = tc_sm_count();
  if (sms <= 0) return -1;
  const Geom& g = a.g;
  if (!a.second_seg && a.nslots >= 2 && a.nslots <= 8) {  // all slots of a tile resident in TMEM
    int ntiles = ceil_div(g.Kd, TC_BM) * ceil_div(a.Ng, 64) * a.nsplit;
    wgrad_gemm_tc_ms<<<ntiles < sms ? ntiles : sms, TC_THREADS, MS_SMEM_BYTES, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
  }
  const int swap = g.N <= 64 ? 1 : 0;
  if (swap) {
    int ntiles = ceil_div(g.Kd, TC_BM) * ceil_div(a.Ng, 64) * a.nslots * a.nsplit;
    wgrad_gemm_tc<64><<<ntiles < sms ? ntiles : sms, TC_THREADS, TcCfg<64>::SMEM_BYTES, st>>>(a, 1);
  } else {
    int ntiles = ceil_div(g.N, TC_BM) * ceil_div(g.Kd, 128) * a.nslots * a.nsplit;
    wgrad_gemm_tc<128><<<ntiles < sms ? ntiles : sms, TC_THREADS, TcCfg<128>::SMEM_BYTES, st>>>(a, 0);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

#else
static inline bool tc_gather_eligible(const Geom&, int) { return false; }
static inline bool tc_wgrad_eligible(const Geom&, int) { return false; }
static inline long long tc_image_elems(int, int, int) { return 0; }
static inline int tc_pack_image(const float*, long long, float*, long long, int, int, int, int, cudaStream_t) { return -1; }
static inline int tc_launch_gather_gemm(const GatherGemmArgs&, int, cudaStream_t, bool) { return -1; }
static inline int tc_launch_wgrad(const WgradArgs&, cudaStream_t) { return -1; }
#endif

}  // namespace curv
