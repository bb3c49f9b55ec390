// Experiment (GPU box): is the ~6300 B/cycle chip-wide TMA ingest cap on the L2 (LTS) side, i.e. does
// TMA multicast within a thread-block cluster raise the bytes DELIVERED per SM when all SMs read the same
// (L2-resident) weight tiles?  Every CTA streams the same 1 MB buffer as 16 KB tiles into a 12-slot ring.
//   C = 1: plain cp.async.bulk.tensor (each CTA loads every tile itself)
//   C = 2, 4, 8: each CTA of a cluster loads 1/C of every tile with .multicast::cluster to all C CTAs
// build: nvcc -gencode arch=compute_100a,code=sm_100a -I mvlm_b200/csrc tools/exp_tma_multicast.cu
//        -o tools/_bin/exp_tma_multicast -lcuda
#include <cooperative_groups.h>

#include <cstdio>
#include <vector>

#include "common.cuh"

using namespace mvlm;
namespace cg = cooperative_groups;

constexpr int kSlots = 12;
constexpr int kTileBytes = 16384;  // 128 rows x 128 B

struct Args {
  CUtensorMap tm[4];  // box rows 128, 64, 32, 16 (cluster size 1, 2, 4, 8)
  int rounds, n_tiles, csize, desync;
  long long* cycles;
};

__device__ __forceinline__ void tma_load_2d_mc(const void* tmap, uint64_t* bar, void* dst, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(ptx::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) mc_kernel(const __grid_constant__ Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t full[kSlots];
  cg::cluster_group cluster = cg::this_cluster();
  const int C = a.csize;
  const int rank = C > 1 ? static_cast<int>(cluster.block_rank()) : 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlots; ++i) ptx::mbar_init(&full[i], 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  if (C > 1) cluster.sync();
  const long long t0 = clock64();
  const int tm_i = C == 1 ? 0 : (C == 2 ? 1 : (C == 4 ? 2 : 3));
  const int rows = 128 / C;
  uint32_t parity = 0;
  // desync: every cluster starts at a different tile, so that plain loads of different clusters cannot be merged in L2
  int tile = a.desync ? static_cast<int>((blockIdx.x / C) * 7u % a.n_tiles) : 0;
  for (int r = 0; r < a.rounds; ++r) {
    // every CTA of the cluster has consumed the previous round: slots may be overwritten by any of them
    if (C > 1) cluster.sync(); else __syncthreads();
    if (threadIdx.x == 0) {
      for (int s = 0; s < kSlots; ++s) {
        ptx::mbar_expect_tx(&full[s], kTileBytes);
        uint8_t* dst = smem + s * kTileBytes + rank * rows * 128;
        if (C == 1)
          ptx::tma_load_2d(&a.tm[0], &full[s], dst, 0, tile * 128);
        else
          tma_load_2d_mc(&a.tm[tm_i], &full[s], dst, 0, tile * 128 + rank * rows, static_cast<uint16_t>((1u << C) - 1));
        tile = tile + 1 == a.n_tiles ? 0 : tile + 1;
      }
      for (int s = 0; s < kSlots; ++s) ptx::mbar_wait(&full[s], parity);
    }
    parity ^= 1;
  }
  __syncthreads();
  if (C > 1) cluster.sync();
  if (threadIdx.x == 0) a.cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
namespace mvlm {
void set_error(const char*, ...) {}
void count_launch(int) {}
}  // namespace mvlm

int main() {
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || !sym) return 2;
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(sym);
  const int n_tiles = 64;  // 1 MB: L2 resident
  __nv_bfloat16* dW;
  cudaMalloc(&dW, static_cast<size_t>(n_tiles) * kTileBytes);
  cudaMemset(dW, 0, static_cast<size_t>(n_tiles) * kTileBytes);
  long long* dCyc;
  cudaMalloc(&dCyc, 148 * 8);
  Args a;
  for (int i = 0; i < 4; ++i) {
    cuuint64_t gdim[2] = {64, (cuuint64_t)n_tiles * 128};
    cuuint64_t gstr[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)(128 >> i)}, es[2] = {1, 1};
    if (enc(&a.tm[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dW, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE))
      return 3;
  }
  a.n_tiles = n_tiles;
  a.rounds = 400;
  a.cycles = dCyc;
  const int smem = kSlots * kTileBytes + 2048;
  cudaFuncSetAttribute(mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(mc_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int desync : {0, 1})
  for (int C : {1, 2, 4, 8}) {
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    cfg.gridDim = dim3(148 / C * C);
    cudaError_t oe = cudaOccupancyMaxActiveClusters(&max_clusters, mc_kernel, &cfg);
    const int grid = (oe == cudaSuccess && max_clusters > 0 ? (max_clusters < 148 / C ? max_clusters : 148 / C) : 148 / C) * C;
    cfg.gridDim = dim3(grid);
    a.csize = C;
    a.desync = desync;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      cudaError_t le = cudaLaunchKernelEx(&cfg, mc_kernel, a);
      cudaEventRecord(e1);
      cudaError_t se = cudaDeviceSynchronize();
      if (le != cudaSuccess || se != cudaSuccess) {
        printf("C=%d: launch %s / sync %s\n", C, cudaGetErrorString(le), cudaGetErrorString(se));
        return 5;
      }
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> cyc(148);
    cudaMemcpy(cyc.data(), dCyc, grid * 8, cudaMemcpyDeviceToHost);
    double mean_cyc = 0;
    for (int i = 0; i < grid; ++i) mean_cyc += cyc[i];
    mean_cyc /= grid;
    const double bytes_per_cta = static_cast<double>(a.rounds) * kSlots * kTileBytes;
    printf("desync %d cluster %d: grid %3d (max active clusters %d)  %.3f ms  delivered %.1f B/cycle/SM, %.2f TB/s aggregate\n", desync, C, grid,
           max_clusters, ms, bytes_per_cta / mean_cyc, bytes_per_cta * grid / (ms * 1e-3) / 1e12);
  }
  return 0;
}
