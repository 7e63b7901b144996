// Microbenchmark: how fast can one SM gather scattered 64-byte rows?
//   mode 0: LDG.128, 4 lanes per row (what the SpMM does), rows summed in registers
//   mode 1: cp.async.bulk (TMA engine, UBLKCP) 64 B per row into shared memory, completion on an mbarrier
//   mode 2: cp.async (LDGSTS) 16 B per lane, 4 lanes per row, into shared memory
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int TB = 256;
constexpr int ROWB = 64;  // bytes per row

__global__ void __launch_bounds__(TB) k_ldg(const int32_t* __restrict__ idx, const double2* __restrict__ P, int64_t nper, double* out) {
  const int l = threadIdx.x & 3, grp = threadIdx.x >> 2;
  const int32_t* my = idx + (int64_t)blockIdx.x * nper;
  double a = 0, b = 0;
  for (int64_t i = grp; i + 3 * (TB / 4) < nper; i += 4 * (TB / 4)) {
    const int32_t c0 = my[i], c1 = my[i + TB / 4], c2 = my[i + 2 * (TB / 4)], c3 = my[i + 3 * (TB / 4)];
    const double2 x0 = P[(int64_t)c0 * 4 + l], x1 = P[(int64_t)c1 * 4 + l], x2 = P[(int64_t)c2 * 4 + l], x3 = P[(int64_t)c3 * 4 + l];
    a += x0.x + x1.x + x2.x + x3.x; b += x0.y + x1.y + x2.y + x3.y;
  }
  if (a + b == 1.2345) out[0] = a;
}

// mode 3: 256-bit loads (ld.global.v4.f64, sm_100+), 2 lanes per row
__global__ void __launch_bounds__(TB) k_ldg256(const int32_t* __restrict__ idx, const double* __restrict__ P, int64_t nper, double* out) {
  const int l = threadIdx.x & 1, grp = threadIdx.x >> 1;
  const int32_t* my = idx + (int64_t)blockIdx.x * nper;
  double a = 0, b = 0;
  constexpr int G = TB / 2;
  for (int64_t i = grp; i + 3 * G < nper; i += 4 * G) {
    const int32_t c0 = my[i], c1 = my[i + G], c2 = my[i + 2 * G], c3 = my[i + 3 * G];
    double x[4][4];
    const double* p0 = P + (int64_t)c0 * 8 + l * 4; const double* p1 = P + (int64_t)c1 * 8 + l * 4;
    const double* p2 = P + (int64_t)c2 * 8 + l * 4; const double* p3 = P + (int64_t)c3 * 8 + l * 4;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x[0][0]), "=d"(x[0][1]), "=d"(x[0][2]), "=d"(x[0][3]) : "l"(p0));
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x[1][0]), "=d"(x[1][1]), "=d"(x[1][2]), "=d"(x[1][3]) : "l"(p1));
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x[2][0]), "=d"(x[2][1]), "=d"(x[2][2]), "=d"(x[2][3]) : "l"(p2));
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x[3][0]), "=d"(x[3][1]), "=d"(x[3][2]), "=d"(x[3][3]) : "l"(p3));
    for (int q = 0; q < 4; q++) { a += x[q][0] + x[q][2]; b += x[q][1] + x[q][3]; }
  }
  if (a + b == 1.2345) out[0] = a;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int STAGES, int RPS>  // rows per stage
__global__ void __launch_bounds__(TB) k_bulk(const int32_t* __restrict__ idx, const char* __restrict__ P, int64_t nper, double* out) {
  extern __shared__ __align__(128) char sm[];
  __shared__ uint64_t bar[STAGES];
  const int32_t* my = idx + (int64_t)blockIdx.x * nper;
  if (threadIdx.x == 0)
    for (int s = 0; s < STAGES; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(1));
  asm volatile("fence.mbarrier_init.release.cluster;");
  __syncthreads();
  const int64_t nst = nper / RPS;
  double acc = 0;
  auto issue = [&](int64_t st) {
    const int s = st % STAGES;
    if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(RPS * ROWB) : "memory");
    __syncwarp();
    for (int r = threadIdx.x; r < RPS; r += TB) {
      const int32_t c = my[st * RPS + r];
      asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + ((size_t)s * RPS + r) * ROWB)),
                   "l"(P + (int64_t)c * ROWB), "r"(ROWB), "r"(smem_u32(&bar[s]))
                   : "memory");
    }
  };
  // NOTE: thread 0's expect_tx must precede the copies of other warps: do a block barrier per stage issue (cheap here)
  for (int64_t st = 0; st < STAGES - 1 && st < nst; st++) { issue(st); __syncthreads(); }
  for (int64_t st = 0; st < nst; st++) {
    if (st + STAGES - 1 < nst) issue(st + STAGES - 1);
    const int s = st % STAGES;
    const uint32_t parity = (st / STAGES) & 1;
    uint32_t ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar[s])), "r"(parity) : "memory");
    // consume: every thread reads 16 B of "its" rows (4 lanes per row)
    for (int r = threadIdx.x >> 2; r < RPS; r += TB / 4) {
      const double2 x = *reinterpret_cast<const double2*>(sm + ((size_t)s * RPS + r) * ROWB + (threadIdx.x & 3) * 16);
      acc += x.x + x.y;
    }
    __syncthreads();
  }
  if (acc == 1.2345) out[0] = acc;
}

template <int STAGES, int RPS>
__global__ void __launch_bounds__(TB) k_ldgsts(const int32_t* __restrict__ idx, const char* __restrict__ P, int64_t nper, double* out) {
  extern __shared__ __align__(128) char sm[];
  const int32_t* my = idx + (int64_t)blockIdx.x * nper;
  const int l = threadIdx.x & 3;
  const int64_t nst = nper / RPS;
  double acc = 0;
  auto issue = [&](int64_t st) {
    const int s = st % STAGES;
    for (int r = threadIdx.x >> 2; r < RPS; r += TB / 4) {
      const int32_t c = my[st * RPS + r];
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sm + ((size_t)s * RPS + r) * ROWB + l * 16)), "l"(P + (int64_t)c * ROWB + l * 16) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int64_t st = 0; st < STAGES - 1; st++) issue(st);
  for (int64_t st = 0; st < nst; st++) {
    if (st + STAGES - 1 < nst) issue(st + STAGES - 1); else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
    __syncthreads();
    const int s = st % STAGES;
    for (int r = threadIdx.x >> 2; r < RPS; r += TB / 4) {
      const double2 x = *reinterpret_cast<const double2*>(sm + ((size_t)s * RPS + r) * ROWB + l * 16);
      acc += x.x + x.y;
    }
    __syncthreads();
  }
  if (acc == 1.2345) out[0] = acc;
}

int main(int argc, char** argv) {
  const int64_t nrows = argc > 1 ? atoll(argv[1]) : 4800000;     // rows of P (64 B each)
  const int64_t nper = argc > 2 ? atoll(argv[2]) : 262144;       // gathers per CTA
  const int local = argc > 3 ? atoi(argv[3]) : 0;                // 0 = uniform random, else window of this many rows around a moving centre
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
  const int nsm = pr.multiProcessorCount;
  char* P; CK(cudaMalloc(&P, nrows * ROWB)); CK(cudaMemset(P, 0, nrows * ROWB));
  double* out; CK(cudaMalloc(&out, 8));
  for (int cps = 1; cps <= 4; cps *= 2) {
    const int grid = nsm * cps;
    std::vector<int32_t> h((size_t)grid * nper);
    uint64_t s = 88172645463325252ull;
    for (int b = 0; b < grid; b++)
      for (int64_t i = 0; i < nper; i++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        int64_t c;
        if (local) { const int64_t centre = (int64_t)((double)b / grid * nrows) + i / 28; c = (centre + (int64_t)(s % (uint64_t)local)) % nrows; }
        else c = (int64_t)(s % (uint64_t)nrows);
        h[(size_t)b * nper + i] = (int32_t)c;
      }
    int32_t* idx; CK(cudaMalloc(&idx, h.size() * 4)); CK(cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto time = [&](const char* name, auto launch) {
      launch(); CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0)); for (int i = 0; i < 5; i++) launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 5;
      CK(cudaGetLastError());
      const double cyc = ms * 1e-3 * pr.clockRate * 1e3;
      printf("cps=%d %-28s %.3f ms  %.2f cyc/row/SM  %.0f GB/s\n", cps, name, ms, cyc / (nper * cps), (double)grid * nper * ROWB / ms / 1e6);
    };
    time("ldg128 4 lanes/row", [&] { k_ldg<<<grid, TB>>>(idx, (const double2*)P, nper, out); });
    time("ldg256 2 lanes/row", [&] { k_ldg256<<<grid, TB>>>(idx, (const double*)P, nper, out); });
    {
      constexpr int ST = 4, RPS = 256;
      CK(cudaFuncSetAttribute(k_bulk<ST, RPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST * RPS * ROWB));
      time("cp.async.bulk 64B, 4x256", [&] { k_bulk<ST, RPS><<<grid, TB, ST * RPS * ROWB>>>(idx, P, nper, out); });
    }
    {
      constexpr int ST = 8, RPS = 256;
      CK(cudaFuncSetAttribute(k_bulk<ST, RPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST * RPS * ROWB));
      if (cps <= 1) time("cp.async.bulk 64B, 8x256", [&] { k_bulk<ST, RPS><<<grid, TB, ST * RPS * ROWB>>>(idx, P, nper, out); });
    }
    {
      constexpr int ST = 4, RPS = 256;
      CK(cudaFuncSetAttribute(k_ldgsts<ST, RPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST * RPS * ROWB));
      time("cp.async 16B x4, 4x256", [&] { k_ldgsts<ST, RPS><<<grid, TB, ST * RPS * ROWB>>>(idx, P, nper, out); });
    }
    CK(cudaFree(idx));
  }
  return 0;
}
