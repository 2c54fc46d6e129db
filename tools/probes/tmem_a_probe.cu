// Probe: tcgen05.mma (kind::f16) with the A operand in tensor memory.  Which TMEM layout does A take?
// Hypothesis: row r of A in lane r, K elements packed two per 32-bit column (element 2j in the low half of column j).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_a_probe tmem_a_probe.cu ; run on a B200.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../boosting-fine-grained-feature-fusion-in-3d-point-cloud-registration_b200/csrc/tc_ptx.cuh"
using namespace kpreg::tc;

constexpr int M = 128, N = 32, KSTEPS = 2, K = 16 * KSTEPS;

__global__ void __launch_bounds__(128) k_probe(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ d, int swap) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // B [N rows x 32 halves] K-major, SWIZZLE_64B box
  for (int i = tid; i < N * 32; i += 128) {
    const int n = i / 32, k = i % 32;
    const uint32_t off = (uint32_t)n * 64u + (uint32_t)(((k >> 3) ^ ((n >> 1) & 3)) << 4) + (uint32_t)(k & 7) * 2u;
    *reinterpret_cast<__half*>(bp + off) = __float2half_rn(k < K ? b[n * K + k] : 0.f);
  }
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_off = (uint32_t)(32 * warp) << 16;
  // A row tid -> TMEM lane tid, columns 32 .. 32 + K/2
  {
    uint32_t r[16];
    for (int j = 0; j < K / 2; ++j) {
      const __half lo = __float2half_rn(a[tid * K + 2 * j]), hi = __float2half_rn(a[tid * K + 2 * j + 1]);
      const __half2 p = swap ? __halves2half2(hi, lo) : __halves2half2(lo, hi);
      r[j] = *reinterpret_cast<const uint32_t*>(&p);
    }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(tmem + lane_off + 32u),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                 "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  if (tid == 0) {
    tcgen05_fence_after();
    const uint32_t idesc = make_instr_desc_f16(M, N);
    const uint64_t bdesc = make_smem_desc_sw64(base);
    for (int ks = 0; ks < KSTEPS; ++ks) {
      const uint32_t a_addr = tmem + 32u + 8u * ks;  // 8 columns per 16 halves
      const uint64_t bd = bdesc + (uint64_t)(ks * 2);
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
          "}" ::"r"(tmem), "r"(a_addr), "l"(bd), "r"(idesc), "r"(ks ? 1u : 0u)
          : "memory");
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tcgen05_fence_after();
  uint32_t v[32];
  tmem_ld_32x32b_x32(tmem + lane_off, v);
  tmem_ld_wait();
  for (int j = 0; j < N; ++j) d[tid * N + j] = __uint_as_float(v[j]);
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64));
}

int main() {
  std::vector<float> a(M * K), b(N * K), d(M * N), ref(M * N);
  uint32_t st = 12345u;
  auto rnd = [&]() { st = st * 1664525u + 1013904223u; return (int)((st >> 16) % 13) - 6; };
  for (int i = 0; i < M * K; ++i) a[i] = (float)rnd();
  for (int i = 0; i < N * K; ++i) b[i] = (float)rnd();
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0.f;
      for (int k = 0; k < K; ++k) s += a[m * K + k] * b[n * K + k];
      ref[m * N + n] = s;
    }
  float *da, *db, *dd;
  cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dd, d.size() * 4);
  cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
  for (int swap = 0; swap < 2; ++swap) {
    cudaMemset(dd, 0, d.size() * 4);
    k_probe<<<1, 128, 8192>>>(da, db, dd, swap);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < M * N; ++i) bad += d[i] != ref[i];
    printf("swap=%d: %s, mismatches %d of %d; d[0..3] = %g %g %g %g  ref %g %g %g %g; d[row 77] %g ref %g\n", swap, cudaGetErrorString(e), bad, M * N,
           d[0], d[1], d[2], d[3], ref[0], ref[1], ref[2], ref[3], d[77 * N + 5], ref[77 * N + 5]);
  }
  return 0;
}
