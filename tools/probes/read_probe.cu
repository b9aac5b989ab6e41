// Read-bandwidth probes (not product code): how fast can a B200 stream 2.1 GB of fp32 rows?
//   mode 0: grid-stride ld.global.nc.v4 reduction, full occupancy
//   mode 1: the stream kernel's skeleton: persistent CTAs, 4 groups x 128 threads, one 2x16 KiB TMA stage
//           per group, prefetch distance one row, LDS.128 of the whole stage, one named barrier per row
//   mode 2: as mode 1 but two stages per group of 2x8 KiB (half rows), prefetch distance two half-rows
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o read_probe read_probe.cu ; run: ./read_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned long long* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma(void* dst, const void* src, uint32_t bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

__global__ void probe_ldg(const float4* __restrict__ a, const float4* __restrict__ b, size_t n4, float* out) {
  float acc = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v, w;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a + i));
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(w.x), "=f"(w.y), "=f"(w.z), "=f"(w.w) : "l"(b + i));
    acc += v.x + v.y + v.z + v.w + w.x + w.y + w.z + w.w;
  }
  if (acc == 12345.678f) out[0] = acc;
}

template <int STAGES, int ROWF>  // ROWF floats per tensor per stage
__global__ void __launch_bounds__(512, 1) probe_tma(const float* __restrict__ a, const float* __restrict__ b, long long units, float* out) {
  extern __shared__ __align__(128) unsigned char raw[];
  const int g = threadIdx.x >> 7, tg = threadIdx.x & 127;
  float* base = reinterpret_cast<float*>(raw) + (size_t)g * (STAGES * 2 * ROWF + 64);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(base + STAGES * 2 * ROWF);
  const long long G = (long long)gridDim.x * 4, first = (long long)g * gridDim.x + blockIdx.x;
  if (tg == 0) for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
  asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");
  auto issue = [&](long long u, int s) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect(&bars[s], 2 * ROWF * 4);
    tma(base + s * 2 * ROWF, a + u * ROWF, ROWF * 4, &bars[s]);
    tma(base + s * 2 * ROWF + ROWF, b + u * ROWF, ROWF * 4, &bars[s]);
  };
  if (tg == 0) for (int s = 0; s < STAGES; ++s) if (first + s * G < units) issue(first + s * G, s);
  float acc = 0.f;
  long long it = 0;
  for (long long u = first; u < units; u += G, ++it) {
    const int s = it % STAGES;
    mbar_wait(&bars[s], (it / STAGES) & 1);
    const float* st = base + s * 2 * ROWF;
#pragma unroll
    for (int i = 0; i < 2 * ROWF / 512; ++i) {
      const float4 v = *reinterpret_cast<const float4*>(st + 4 * (128 * i + tg));
      acc += v.x + v.y + v.z + v.w;
    }
    asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");
    if (tg == 0 && u + STAGES * G < units) issue(u + STAGES * G, s);
  }
  if (acc == 12345.678f) out[0] = acc;
}

int main(int argc, char** argv) {
  const long long rows = 65536, K = 4096;
  const size_t n = (size_t)rows * K;
  float *a, *b, *out;
  cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4); cudaMalloc(&out, 4);
  cudaMemset(a, 0, n * 4); cudaMemset(b, 0, n * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto time = [&](const char* name, auto&& launch) {
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) launch();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 20;
    printf("%-44s %.3f ms  %.0f GB/s  (%s)\n", name, ms, 2.0 * n * 4 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  };
  if (argc > 1) {  // sustained mode: 30 chunks of 100 launches of the best plain-load configuration and of the TMA skeleton
    auto kt = probe_tma<1, 4096>; const int smem_t = 4 * (1 * 2 * 4096 + 64) * 4;
    cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_t);
    for (int mode = 0; mode < 2; ++mode) {
      printf("sustained %s:", mode ? "tma 1 stage" : "ldg.v4 grid=1184 block=512");
      for (int c = 0; c < 30; ++c) {
        cudaEventRecord(e0);
        for (int i = 0; i < 100; ++i) {
          if (mode) kt<<<148, 512, smem_t>>>(a, b, rows, out);
          else probe_ldg<<<1184, 512>>>((const float4*)a, (const float4*)b, n / 4, out);
        }
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf(" %.0f", 2.0 * n * 4 / (ms / 100) / 1e6);
      }
      printf(" GB/s\n");
    }
    return 0;
  }
  for (int ctas : {148 * 4, 148 * 8, 148 * 16})
    for (int thr : {256, 512}) {
      char nm[64]; snprintf(nm, 64, "ldg.v4 grid=%d block=%d", ctas, thr);
      time(nm, [&] { probe_ldg<<<ctas, thr>>>((const float4*)a, (const float4*)b, n / 4, out); });
    }
  {
    auto k = probe_tma<1, 4096>; const int smem = 4 * (1 * 2 * 4096 + 64) * 4;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    time("tma 1 stage x (2x16 KiB) per group", [&] { k<<<148, 512, smem>>>(a, b, rows, out); });
  }
  {
    auto k = probe_tma<2, 2048>; const int smem = 4 * (2 * 2 * 2048 + 64) * 4;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    time("tma 2 stages x (2x8 KiB) per group", [&] { k<<<148, 512, smem>>>(a, b, rows * 2, out); });
  }
  {
    auto k = probe_tma<3, 2048>; const int smem = 4 * (3 * 2 * 2048 + 64) * 4;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    time("tma 3 stages x (2x8 KiB) per group", [&] { k<<<148, 512, smem>>>(a, b, rows * 2, out); });
  }
  {
    auto k = probe_tma<4, 1024>; const int smem = 4 * (4 * 2 * 1024 + 64) * 4;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    time("tma 4 stages x (2x4 KiB) per group", [&] { k<<<148, 512, smem>>>(a, b, rows * 4, out); });
  }
  return 0;
}
