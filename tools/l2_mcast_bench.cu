// Micro-benchmark: how many bytes per clock do TMA loads deliver into the SMs out of L2, and does multicast raise it?
//   mode 0  every CTA loads distinct 32 KB boxes                         (delivered = L2 reads)
//   mode 1  the CTAs of a cluster load the SAME box, each for itself      (unicast duplicates: L2-side dedup?)
//   mode 2  each CTA loads 1/C of the box and multicasts it to the cluster (delivered = C x L2 reads)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/l2_mcast tools/l2_mcast_bench.cu
// Run:   gpurun_out/l2_mcast            (prints bytes/clk per SM and chip-wide for cluster sizes 1, 2, 4, 8)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../3dgan_b200/csrc/ptx.cuh"
using namespace b200;

constexpr int kStages = 4;
constexpr int kBoxRows = 256, kBoxCols = 64;            // 256 x 64 bf16 = 32 KB, 128-byte swizzled rows
constexpr int kBoxBytes = kBoxRows * kBoxCols * 2;

struct Params {
  CUtensorMap tm_full;     // box 64 x 256
  CUtensorMap tm_part[4];  // box 64 x (256 / C) for C = 1, 2, 4, 8
  int mode, csz, iters, rows_total;
};

__global__ void __launch_bounds__(128, 1) bench_kernel(const __grid_constant__ Params p, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kBoxBytes);
  uint64_t* empty = full + kStages;                       // multicast: every CTA of the cluster has consumed the stage
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), p.csz);
    }
    fence_mbar_init();
  }
  __syncthreads();
  if (p.csz > 1) cluster_sync_all();
  const int cluster_id = blockIdx.x / p.csz;
  const int nclusters = gridDim.x / p.csz;
  const int boxes_total = p.rows_total / kBoxRows;
  long long t0 = 0;
  if (threadIdx.x == 0) {
    t0 = clock64();
    uint32_t par = 0;
    int s = 0;
    const int log2c = p.csz == 1 ? 0 : (p.csz == 2 ? 1 : (p.csz == 4 ? 2 : 3));
    const int part_rows = kBoxRows / p.csz;
    const bool mc = p.mode == 2 && p.csz > 1;
    for (int it = 0; it < p.iters + kStages; ++it) {
      if (it >= kStages) {                       // consume the stage loaded kStages iterations ago ...
        mbar_wait(smem_u32(&full[s]), par);
        if (mc) {                                // ... tell every CTA of the cluster, and wait until all have
          for (int c = 0; c < p.csz; ++c) mbar_arrive_cluster(smem_u32(&empty[s]), c);
          if (it < p.iters) mbar_wait(smem_u32(&empty[s]), par);
        }
      }
      if (it < p.iters) {
        const uint32_t bar = smem_u32(&full[s]);
        const uint32_t dst = smem_u32(smem) + s * kBoxBytes;
        mbar_arrive_expect_tx(bar, kBoxBytes);
        if (p.mode == 0) {
          const int box = (blockIdx.x + it * gridDim.x) % boxes_total;
          tma_load_2d(dst, &p.tm_full, bar, 0, box * kBoxRows);
        } else if (p.mode == 1 || p.csz == 1) {
          const int box = (cluster_id + it * nclusters) % boxes_total;
          tma_load_2d(dst, &p.tm_full, bar, 0, box * kBoxRows);
        } else {
          const int box = (cluster_id + it * nclusters) % boxes_total;
          const uint16_t mask = (uint16_t)((1u << p.csz) - 1u);
          tma_load_2d_mc(dst + rank * part_rows * kBoxCols * 2, &p.tm_part[log2c], bar, 0,
                         box * kBoxRows + rank * part_rows, mask);
        }
      }
      if (++s == kStages) { s = 0; if (it >= kStages) par ^= 1; }
    }
  }
  __syncthreads();
  if (p.csz > 1) cluster_sync_all();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  const int rows_total = 256 * 256;                       // 65536 rows x 64 cols bf16 = 8 MB: stays in L2
  void* buf;
  cudaMalloc(&buf, (size_t)rows_total * kBoxCols * 2);
  cudaMemset(buf, 0, (size_t)rows_total * kBoxCols * 2);
  long long* d_cycles;
  cudaMalloc(&d_cycles, 8);
  Params p;
  auto make = [&](CUtensorMap* tm, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)kBoxCols, (cuuint64_t)rows_total}, str[1] = {(cuuint64_t)kBoxCols * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBoxCols, (cuuint32_t)box_rows}, es[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  };
  make(&p.tm_full, kBoxRows);
  for (int i = 0; i < 4; ++i) make(&p.tm_part[i], kBoxRows >> i);
  p.rows_total = rows_total;
  p.iters = 2000;
  const size_t smem = kStages * kBoxBytes + 128 + 1024;
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int csz : {1, 2, 4, 8}) {
    for (int mode = 0; mode < 3; ++mode) {
      if (csz == 1 && mode != 0) continue;
      p.mode = mode; p.csz = csz;
      const int grid = sms / csz * csz;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int max_clusters = 0;
      cudaOccupancyMaxActiveClusters(&max_clusters, bench_kernel, &cfg);
      float best = 1e30f;
      long long cyc = 0;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        cudaError_t err = cudaLaunchKernelEx(&cfg, bench_kernel, p, d_cycles);
        cudaEventRecord(e1);
        cudaError_t err2 = cudaDeviceSynchronize();
        if (err != cudaSuccess || err2 != cudaSuccess) { printf("csz %d mode %d: %s / %s\n", csz, mode, cudaGetErrorString(err), cudaGetErrorString(err2)); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
        cudaMemcpy(&cyc, d_cycles, 8, cudaMemcpyDeviceToHost);
      }
      const double delivered = (double)grid * p.iters * kBoxBytes;
      const double reads = mode == 2 ? delivered / csz : delivered;
      printf("cluster %d mode %d (%s): grid %d (max co-resident clusters %d)  %.3f ms  %lld cycles  delivered %.1f B/clk/SM = %.0f B/clk chip (%.2f TB/s); L2 reads requested %.0f B/clk chip\n",
             csz, mode, mode == 0 ? "distinct" : (mode == 1 ? "same box, unicast" : "multicast"), grid, max_clusters, best, cyc,
             delivered / grid / cyc, delivered / cyc, delivered / (best * 1e-3) / 1e12, reads / cyc);
    }
  }
  return 0;
}
