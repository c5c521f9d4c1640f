// probe_bench.cu -- microbenchmarks behind the design choices of scan.cu:
// how many random table probes per SM-cycle the different paths sustain on a B200.
//   ldg   : ld.global.cg.v4 / .v2 / .u32 at random 16-byte buckets of a table of 2^k bytes
//   tex   : the same through a texture object (tex1Dfetch<uint4>)
//   dsmem : random 4-byte ld.shared::cluster out of a bitmap spread over the CTAs of a cluster
//   lds   : the same from the CTA's own shared memory (cluster size 1)
//   tma   : 16-byte cp.async.bulk global->shared, issued by every lane
// Every CTA = 1024 threads, grid = #SMs (one CTA per SM, like scan_kernel).
#include <cooperative_groups.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}

template <int W>  // W = bytes per probe: 4, 8, 16
__global__ void __launch_bounds__(1024, 1) ldg_kernel(const uint4 *tab, uint32_t mask, int iters, uint32_t *sink) {
  uint32_t h = mix(blockIdx.x * 1024 + threadIdx.x + 1), acc = 0;
  for (int i = 0; i < iters; ++i) {
    uint32_t a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { h = h * 0x9E3779B1u + 0x85EBCA6Bu; a[u] = (h >> 4) & mask; }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint4 *p = tab + a[u];
      if (W == 16) { uint4 v; asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); acc += v.x ^ v.y ^ v.z ^ v.w; }
      else if (W == 8) { uint2 v; asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p)); acc += v.x ^ v.y; }
      else { uint32_t v; asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p)); acc += v; }
    }
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

__global__ void __launch_bounds__(1024, 1) tex_kernel(cudaTextureObject_t tex, uint32_t mask, int iters, uint32_t *sink) {
  uint32_t h = mix(blockIdx.x * 1024 + threadIdx.x + 1), acc = 0;
  for (int i = 0; i < iters; ++i) {
    uint32_t a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { h = h * 0x9E3779B1u + 0x85EBCA6Bu; a[u] = (h >> 4) & mask; }
#pragma unroll
    for (int u = 0; u < 4; ++u) { const uint4 v = tex1Dfetch<uint4>(tex, (int)a[u]); acc += v.x ^ v.y ^ v.z ^ v.w; }
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

// bitmap of `words_per_cta` 32-bit words in every CTA of the cluster; a probe picks CTA = hash bits, word = hash bits
__global__ void __launch_bounds__(1024, 1) dsmem_kernel(uint32_t words_per_cta, int iters, uint32_t *sink, int local_only) {
  extern __shared__ uint32_t bm[];
  cg::cluster_group cl = cg::this_cluster();
  const uint32_t csz = cl.num_blocks();
  for (uint32_t i = threadIdx.x; i < words_per_cta; i += 1024) bm[i] = mix(i + blockIdx.x);
  cl.sync();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(bm);
  uint32_t h = mix(blockIdx.x * 1024 + threadIdx.x + 1), acc = 0;
  const uint32_t wmask = words_per_cta - 1;
  for (int i = 0; i < iters; ++i) {
    uint32_t a[4], r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { h = h * 0x9E3779B1u + 0x85EBCA6Bu; a[u] = base + (((h >> 8) & wmask) << 2); r[u] = local_only ? cl.block_rank() : (h >> 28) % csz; }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      uint32_t ra, v;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a[u]), "r"(r[u]));
      asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(ra));
      acc += v;
    }
  }
  if (acc == 0x12345678u) sink[0] = acc;
  cl.sync();
}

__global__ void __launch_bounds__(1024, 1) lds_kernel(uint32_t words, int iters, uint32_t *sink) {
  extern __shared__ uint32_t bm[];
  for (uint32_t i = threadIdx.x; i < words; i += 1024) bm[i] = mix(i + blockIdx.x);
  __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(bm);
  uint32_t h = mix(blockIdx.x * 1024 + threadIdx.x + 1), acc = 0;
  const uint32_t wmask = words - 1;
  for (int i = 0; i < iters; ++i) {
    uint32_t a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { h = h * 0x9E3779B1u + 0x85EBCA6Bu; a[u] = base + (((h >> 8) & wmask) << 2); }
#pragma unroll
    for (int u = 0; u < 4; ++u) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a[u])); acc += v; }
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

// every lane issues 16-byte bulk copies into its own 16-byte slot; one mbarrier per warp
__global__ void __launch_bounds__(1024, 1) tma_kernel(const uint4 *tab, uint32_t mask, int iters, uint32_t *sink) {
  __shared__ __align__(16) uint4 slot[1024][2];
  __shared__ __align__(8) uint64_t bar[32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar[warp]);
  if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
  __syncthreads();
  uint32_t h = mix(blockIdx.x * 1024 + threadIdx.x + 1), acc = 0, par = 0;
  for (int i = 0; i < iters; ++i) {
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(32u * 2u * 16u) : "memory");
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      h = h * 0x9E3779B1u + 0x85EBCA6Bu;
      const uint4 *p = tab + ((h >> 4) & mask);
      const uint32_t d = (uint32_t)__cvta_generic_to_shared(&slot[threadIdx.x][u]);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];" ::"r"(d), "l"(p), "r"(b) : "memory");
    }
    uint32_t ok;
    do {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(b), "r"(par) : "memory");
    } while (!ok);
    par ^= 1;
    acc += slot[threadIdx.x][0].x ^ slot[threadIdx.x][1].y;
    __syncwarp();
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  printf("device %s, %d SMs, max clock %d MHz\n", prop.name, sms, clk_khz / 1000);
  uint32_t *sink; CK(cudaMalloc(&sink, 64));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double ghz = clk_khz / 1e6;
  auto report = [&](const char *name, double probes, float ms) {
    printf("%-34s %8.3f ms  %7.2f Gprobe/s  %6.3f probes/SM-cycle(@max clk)  %6.2f cyc/warp-probe\n", name, ms, probes / ms / 1e6,
           probes / (ms * 1e-3) / (sms * ghz * 1e9), 32.0 / (probes / (ms * 1e-3) / (sms * ghz * 1e9)));
  };
  // ---- global tables
  const size_t max_bytes = size_t(256) << 20;
  uint4 *tab; CK(cudaMalloc(&tab, max_bytes)); CK(cudaMemset(tab, 1, max_bytes));
  const int iters = 256;
  for (int mb : {1, 4, 8, 16, 32, 64, 128, 256}) {
    const uint32_t mask = (uint32_t)((size_t(mb) << 20) / 16 - 1);
    const double probes = double(sms) * 1024 * iters * 4;
    char nm[64];
    for (int w : {16, 8, 4}) {
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        if (w == 16) ldg_kernel<16><<<sms, 1024>>>(tab, mask, iters, sink);
        else if (w == 8) ldg_kernel<8><<<sms, 1024>>>(tab, mask, iters, sink);
        else ldg_kernel<4><<<sms, 1024>>>(tab, mask, iters, sink);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      }
      snprintf(nm, sizeof nm, "ldg.cg %2dB table %3d MiB", w, mb);
      report(nm, probes, time_ms(e0, e1));
    }
    // texture
    cudaResourceDesc rd{}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = tab;
    rd.res.linear.desc = cudaCreateChannelDesc<uint4>(); rd.res.linear.sizeInBytes = size_t(mb) << 20;
    cudaTextureDesc td{}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0)); tex_kernel<<<sms, 1024>>>(tex, mask, iters, sink); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    }
    snprintf(nm, sizeof nm, "tex1Dfetch 16B table %3d MiB", mb);
    report(nm, probes, time_ms(e0, e1));
    CK(cudaDestroyTextureObject(tex));
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0)); tma_kernel<<<sms, 1024>>>(tab, mask, iters / 2, sink); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    }
    snprintf(nm, sizeof nm, "tma bulk 16B table %3d MiB", mb);
    report(nm, double(sms) * 1024 * (iters / 2) * 2, time_ms(e0, e1));
  }
  // ---- shared / distributed shared memory bitmaps
  const uint32_t words = 32768; // 128 KiB per CTA
  CK(cudaFuncSetAttribute(lds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, words * 4));
  CK(cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, words * 4));
  CK(cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  const int it2 = 1024;
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(e0)); lds_kernel<<<sms, 1024, words * 4>>>(words, it2, sink); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  }
  report("lds random u32, 128 KiB", double(sms) * 1024 * it2 * 4, time_ms(e0, e1));
  for (int csz : {1, 2, 4, 8, 16}) {
    for (int local_only = 0; local_only < 2; ++local_only) {
      cudaLaunchConfig_t cfg{}; cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = words * 4; cfg.attrs = at; cfg.numAttrs = 1;
      int maxc = 0;
      cfg.gridDim = dim3(csz);
      cudaError_t e = cudaOccupancyMaxActiveClusters(&maxc, dsmem_kernel, &cfg);
      if (e != cudaSuccess) { printf("cluster %d: %s\n", csz, cudaGetErrorString(e)); cudaGetLastError(); break; }
      const int grid = maxc * csz;
      cfg.gridDim = dim3(grid);
      float ms = 0;
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        CK(cudaLaunchKernelEx(&cfg, dsmem_kernel, words, it2, sink, local_only));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); ms = time_ms(e0, e1);
      }
      char nm[96]; snprintf(nm, sizeof nm, "dsmem u32 cluster %2d (%3d CTAs)%s", csz, grid, local_only ? " own" : "");
      // per-SM rate: scale by the CTAs that ran
      const double probes = double(grid) * 1024 * it2 * 4;
      printf("%-34s %8.3f ms  %7.2f Gprobe/s  %6.3f probes/SM-cycle(@max clk, per active SM)\n", nm, ms, probes / ms / 1e6,
             probes / (ms * 1e-3) / (grid * ghz * 1e9));
    }
  }
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
