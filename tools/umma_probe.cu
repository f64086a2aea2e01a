// Probe of tcgen05.mma.kind::tf32 shared-memory operand layouts (run on the GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_probe tools/umma_probe.cu && /tmp/umma_probe
// D[128 x 64] = A[128 x 8] * B[64 x 8]^T with A, B stored MN-major in several candidate layouts.
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Variant { int layout_type; int lbo, sbo; int writer; int a_major, b_major; int ts; };

// writer: byte offset of element (mn, k) inside an operand tile with MN extent `mnext`
__device__ __host__ inline int elem_off(int writer, int mn, int k, int lbo, int sbo) {
  switch (writer) {
    case 0:  // K-major SW128: row = mn (128 B = 32 k), 8-row groups sbo apart
    { int c = (k >> 2) ^ (mn & 7); return (mn >> 3) * sbo + (mn & 7) * 128 + c * 16 + (k & 3) * 4; }
    case 1:  // MN-major SW128 (Swizzle<3,4,3>): row = k (8 rows), 32 mn per 128 B, atoms lbo apart
    { int c = ((mn & 31) >> 2) ^ (k & 7); return (mn >> 5) * lbo + (k >> 3) * sbo + (k & 7) * 128 + c * 16 + (mn & 3) * 4; }
    case 2:  // MN-major SW128_BASE32B (Swizzle<2,5,2>): atoms of 4 k rows
    { int c8 = (mn & 31) >> 2; int c = (((c8 >> 1) ^ (k & 3)) << 1) | (c8 & 1);
      return (mn >> 5) * lbo + (k >> 2) * sbo + (k & 3) * 128 + c * 16 + (mn & 3) * 4; }
    case 3:  // MN-major no swizzle (interleave): core = 8 k-rows x 16 B; mn/4 -> sbo, k/8 -> lbo
      return (mn >> 2) * sbo + (k >> 3) * lbo + (k & 7) * 16 + (mn & 3) * 4;
    case 4:  // MN-major no swizzle, lbo/sbo roles swapped
      return (mn >> 2) * lbo + (k >> 3) * sbo + (k & 7) * 16 + (mn & 3) * 4;
  }
  return 0;
}

__global__ void probe(Variant v, const float* A, const float* B, float* D) {
  // A: [128][8] (m, k), B: [64][8] (n, k) plain row-major in global
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* sm = raw + (base - smem_u32(raw));
  uint8_t* sA = sm; uint8_t* sB = sm + 32768;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) ((float*)sm)[i] = 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
    int m = i / 8, k = i % 8;
    *(float*)(sA + elem_off(v.writer, m, k, v.lbo, v.sbo)) = A[i];
  }
  for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
    int n = i / 8, k = i % 8;
    *(float*)(sB + elem_off(v.writer, n, k, v.lbo, v.sbo)) = B[i];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem = tmem_slot;
  if (v.ts) {  // A operand from tensor memory: lane = row m, columns 64..71 = k
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    uint32_t r[8];
    for (int k = 0; k < 8; ++k) r[k] = __float_as_uint(A[(w * 32 + l) * 8 + k]);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(
                     tmem + ((uint32_t)(w * 32) << 16) + 64u),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    auto desc = [&](uint32_t addr) {
      uint64_t d = 0;
      d |= (uint64_t)((addr >> 4) & 0x3FFF);
      d |= (uint64_t)((v.lbo >> 4) & 0x3FFF) << 16;
      d |= (uint64_t)((v.sbo >> 4) & 0x3FFF) << 32;
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)v.layout_type << 61;
      return d;
    };
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)v.a_major << 15) | ((uint32_t)v.b_major << 16) |
                     ((64u >> 3) << 17) | ((128u >> 4) << 24);
    uint64_t da = desc(base), db = desc(base + 32768);
    if (v.ts) {
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem), "r"(tmem + 64u),
                   "l"(db), "r"(idesc), "r"(0u) : "memory");
    } else {
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db),
                   "r"(idesc), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait
  uint32_t ok = 0;
  for (int i = 0; i < (1 << 22) && !ok; ++i)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < 4) {
    for (int c0 = 0; c0 < 64; c0 += 8) {
      uint32_t r[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) D[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
}

int main() {
  std::vector<float> A(128 * 8), B(64 * 8), D(128 * 64), R(128 * 64);
  for (int i = 0; i < 128 * 8; ++i) A[i] = (float)((i * 37 % 17) - 8) / 4.f;   // exactly representable in tf32
  for (int i = 0; i < 64 * 8; ++i) B[i] = (float)((i * 53 % 13) - 6) / 2.f;
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
    float s = 0; for (int k = 0; k < 8; ++k) s += A[m * 8 + k] * B[n * 8 + k]; R[m * 64 + n] = s; }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  Variant vs[] = {
      {2, 16, 1024, 0, 0, 0, 0},   // control: K-major SW128
      {2, 16, 1024, 0, 0, 0, 1},   // TS mode: A from TMEM (lane = row, column = k), B K-major SW128
      {2, 4096, 1024, 1, 1, 1, 0},    // MN SW128, lbo = atom stride, sbo = k-group stride
      {2, 1024, 4096, 1, 1, 1, 0},    //   (writer uses lbo for atoms; try both orders)
      {1, 4096, 512, 2, 1, 1, 0},     // MN SW128_BASE32B
      {1, 512, 4096, 2, 1, 1, 0},
      {0, 128, 256, 3, 1, 1, 0},      // MN no swizzle: sbo between mn cores (128 B apart), lbo between k groups
      {0, 4096, 128, 3, 1, 1, 0},
      {0, 128, 4096, 4, 1, 1, 0},
      {0, 256, 128, 4, 1, 1, 0},
  };
  for (auto& v : vs) {
    cudaMemset(dD, 0, D.size() * 4);
    probe<<<1, 128, 70000>>>(v, dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0, nz = 0;
    for (int i = 0; i < 128 * 64; ++i) { err = fmax(err, fabs(D[i] - R[i])); nz += D[i] != 0; }
    printf("ts=%d type=%d lbo=%d sbo=%d writer=%d major=%d%d : %s max_err=%g nonzero=%g  D[0..3]=%g %g %g %g ref=%g %g %g %g\n",
           v.ts, v.layout_type, v.lbo, v.sbo, v.writer, v.a_major, v.b_major, cudaGetErrorString(e), err, nz, D[0], D[1], D[2], D[3],
           R[0], R[1], R[2], R[3]);
    if (e != cudaSuccess) break;
  }
  return 0;
}
