// Gather micro-benchmark for sm_100a: what one L1/LSU "wavefront" costs for the access shapes the neighbor-list
// kernels use.  Prints cycles per warp-instruction per SM and lanes served per cycle per SM for each pattern.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gather_bench.bin gather.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ double4 ld256(const double4 *p)
{
  double4 r;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ double2 ld128(const double2 *p)
{
  double2 r;
  asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ double ld64(const double *p)
{
  double r;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}

// idx: per-thread index stream [iters][nthreads_total] (coalesced reads), MODE selects the load width
template <int MODE, int UNROLL>
__global__ void __launch_bounds__(256) k_gather(const char *__restrict__ tab, const int *__restrict__ idx, int iters,
                                                double *__restrict__ out)
{
  const size_t nt = (size_t) gridDim.x * blockDim.x;
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0;
  for (int it = 0; it < iters; it += UNROLL) {
    int j[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) j[u] = __ldcs(idx + (size_t) (it + u) * nt + t);
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      if (MODE == 256) {
        double4 v = ld256((const double4 *) (tab + 32 * (size_t) j[u]));
        acc += v.x + v.w;
      } else if (MODE == 128) {
        double2 v = ld128((const double2 *) (tab + 32 * (size_t) j[u]));
        acc += v.x + v.y;
      } else if (MODE == 64) {
        acc += ld64((const double *) (tab + 32 * (size_t) j[u]));
      } else if (MODE == 512) {    // 64-byte row as two 32-byte sectors
        double4 v = ld256((const double4 *) (tab + 64 * (size_t) (j[u] >> 1)));
        double4 w = ld256((const double4 *) (tab + 64 * (size_t) (j[u] >> 1) + 32));
        acc += v.x + w.w;
      }
    }
  }
  if (acc == 1.2345) out[t] = acc;
}

// shared-memory gathers from a table of `rows` 32-byte rows (LDS.128 x2) or 16-byte rows (LDS.128)
template <int BYTES, int UNROLL>
__global__ void __launch_bounds__(256) k_lds(const char *__restrict__ tab, int rows, const int *__restrict__ idx,
                                             int iters, double *__restrict__ out)
{
  extern __shared__ __align__(16) char sm[];
  for (int k = threadIdx.x; k < rows * BYTES / 16; k += blockDim.x) ((double2 *) sm)[k] = ((const double2 *) tab)[k];
  __syncthreads();
  const size_t nt = (size_t) gridDim.x * blockDim.x;
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0;
  for (int it = 0; it < iters; it += UNROLL) {
    int j[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) j[u] = __ldcs(idx + (size_t) (it + u) * nt + t) % rows;
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const double2 *p = (const double2 *) (sm + (size_t) BYTES * j[u]);
      double2 a = p[0];
      acc += a.x + a.y;
      if (BYTES == 32) {
        double2 b = p[1];
        acc += b.x + b.y;
      }
    }
  }
  if (acc == 1.2345) out[t] = acc;
}

// FP64 atomics (RED.ADD.F64), 3 per lane per iteration at index j (the f[j] scatter of a half-list force kernel)
template <int UNROLL>
__global__ void __launch_bounds__(256) k_red(double *__restrict__ f, const int *__restrict__ idx, int iters)
{
  const size_t nt = (size_t) gridDim.x * blockDim.x;
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  for (int it = 0; it < iters; it += UNROLL) {
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const int j = __ldcs(idx + (size_t) (it + u) * nt + t);
      atomicAdd(f + 4 * (size_t) j, 1.0);
      atomicAdd(f + 4 * (size_t) j + 1, 1.0);
      atomicAdd(f + 4 * (size_t) j + 2, 1.0);
    }
  }
}

struct Pattern {
  const char *name;
  int group;        // lanes that share one 128-byte line (1 = fully random, 4 = consecutive sectors of one line)
  int same;         // 1: all lanes of a group read the SAME sector
  size_t rows;      // table rows (32 B) addressed
};

int main(int argc, char **argv)
{
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp pr;
  CK(cudaGetDeviceProperties(&pr, dev));
  const int sms = pr.multiProcessorCount;
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
  printf("device %s, %d SMs, clock attr %.0f MHz\n", pr.name, sms, clk_khz / 1000.0);
  const int ctas = sms * 8, threads = 256, iters = 256;
  const size_t nt = (size_t) ctas * threads;
  const size_t big_rows = (size_t) 1 << 26;      // 2 GiB of 32-byte rows: DRAM-resident
  const size_t mid_rows = (size_t) 2400000;      // 77 MB: one GPU's positions (L2-resident)
  const size_t small_rows = (size_t) 160000;     // 5 MB: spline tables
  char *tab;
  CK(cudaMalloc(&tab, big_rows * 32));
  CK(cudaMemset(tab, 0, big_rows * 32));
  int *idx;
  CK(cudaMalloc(&idx, nt * iters * sizeof(int)));
  double *out;
  CK(cudaMalloc(&out, nt * sizeof(double) + 4 * 32 * mid_rows));
  std::vector<int> h(nt * iters);
  std::mt19937_64 rng(1234);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));

  const Pattern pats[] = {
      {"random sector, 2 GiB table", 1, 0, big_rows},
      {"random sector, 77 MB table (positions)", 1, 0, mid_rows},
      {"random sector, 5 MB table (splines)", 1, 0, small_rows},
      {"2 lanes = consecutive sectors, 77 MB", 2, 0, mid_rows},
      {"4 lanes = one 128 B line, 77 MB", 4, 0, mid_rows},
      {"8 lanes = two consecutive lines, 77 MB", 8, 0, mid_rows},
      {"32 lanes consecutive (coalesced), 77 MB", 32, 0, mid_rows},
      {"4 lanes same sector, 77 MB", 4, 1, mid_rows},
      {"32 lanes same sector (broadcast), 77 MB", 32, 1, mid_rows},
      {"neighbor-row-like runs (len 1-8, unaligned), 77 MB", -1, 0, mid_rows},
  };
  auto fill = [&](const Pattern &p) {
    for (int it = 0; it < iters; it++)
      for (size_t w = 0; w < nt / 32; w++) {
        int *row = &h[(size_t) it * nt + w * 32];
        if (p.group == -1) {
          int l = 0;
          while (l < 32) {
            int len = 1 + (int) (rng() % 8);
            size_t base = rng() % (p.rows - 16);
            for (int k = 0; k < len && l < 32; k++) row[l++] = (int) (base + k);
          }
          continue;
        }
        for (int g = 0; g < 32; g += p.group) {
          size_t base = (rng() % (p.rows / p.group)) * p.group;
          for (int k = 0; k < p.group; k++) row[g + k] = (int) (base + (p.same ? 0 : k));
        }
      }
    CK(cudaMemcpy(idx, h.data(), nt * iters * sizeof(int), cudaMemcpyHostToDevice));
  };
  auto report = [&](const char *what, const char *name, float ms, double instr_per_thread) {
    // warp-instructions per SM = ctas/sms * (threads/32) * instr ; cycles = ms * clock
    const double winstr_sm = (double) ctas / sms * (threads / 32) * instr_per_thread;
    const double cyc = ms * 1e-3 * clk_khz * 1e3;
    printf("%-10s %-52s %8.3f ms  %7.2f cyc/warp-instr/SM  %6.3f lanes/cyc/SM  %7.1f G lanes/s\n", what, name, ms,
           cyc / winstr_sm, 32.0 * winstr_sm / cyc, 32.0 * winstr_sm * sms / (ms * 1e-3) / 1e9);
  };
  for (const Pattern &p : pats) {
    fill(p);
#define RUN(MODE, LABEL, NI)                                                        \
  {                                                                                 \
    k_gather<MODE, 4><<<ctas, threads>>>(tab, idx, iters, out);                     \
    CK(cudaEventRecord(e0));                                                        \
    k_gather<MODE, 4><<<ctas, threads>>>(tab, idx, iters, out);                     \
    CK(cudaEventRecord(e1));                                                        \
    CK(cudaEventSynchronize(e1));                                                   \
    float ms;                                                                       \
    CK(cudaEventElapsedTime(&ms, e0, e1));                                          \
    report(LABEL, p.name, ms, (double) iters * NI);                                 \
  }
    RUN(256, "LDG.256", 1)
    RUN(128, "LDG.128", 1)
    RUN(64, "LDG.64", 1)
    if (p.group == 1) RUN(512, "2xLDG.256", 2)
  }
  // shared memory: 6656-row tables
  {
    Pattern p = {"random row of a 6656-row table", 1, 0, 6656};
    fill(p);
    const int rows = 6656;
    CK(cudaFuncSetAttribute(k_lds<32, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, rows * 32));
    CK(cudaFuncSetAttribute(k_lds<16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, rows * 16));
    for (int b = 0; b < 2; b++) {
      float ms;
      const int c2 = sms * (b ? 2 : 1);
      const size_t nt2 = (size_t) c2 * threads;
      (void) nt2;
      if (b == 0) {
        k_lds<32, 4><<<ctas, threads, rows * 32>>>(tab, rows, idx, iters, out);
        CK(cudaEventRecord(e0));
        k_lds<32, 4><<<ctas, threads, rows * 32>>>(tab, rows, idx, iters, out);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        report("LDS 32B", p.name, ms, (double) iters);
      } else {
        k_lds<16, 4><<<ctas, threads, rows * 16>>>(tab, rows, idx, iters, out);
        CK(cudaEventRecord(e0));
        k_lds<16, 4><<<ctas, threads, rows * 16>>>(tab, rows, idx, iters, out);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        report("LDS 16B", p.name, ms, (double) iters);
      }
    }
  }
  // atomics
  {
    Pattern p = {"3 x RED.F64 per lane, random atom of 2.4 M", 1, 0, mid_rows};
    fill(p);
    double *f = out + nt;
    k_red<4><<<ctas, threads>>>(f, idx, iters);
    CK(cudaEventRecord(e0));
    k_red<4><<<ctas, threads>>>(f, idx, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    report("RED.F64x3", p.name, ms, (double) iters);
    Pattern q = {"3 x RED.F64 per lane, neighbor-row-like runs", -1, 0, mid_rows};
    fill(q);
    k_red<4><<<ctas, threads>>>(f, idx, iters);
    CK(cudaEventRecord(e0));
    k_red<4><<<ctas, threads>>>(f, idx, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    report("RED.F64x3", q.name, ms, (double) iters);
  }
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
