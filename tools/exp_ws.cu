// Experiment (GPU box): tcgen05.mma.ws (M = 32 / 64 / 128) against the plain form: (a) which TMEM lanes hold the 64 output rows,
// (b) cycles per instruction against M=128 at N=256, K=16 (operands resident in shared memory, no TMA traffic).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -I mvlm_b200/csrc tools/exp_ws.cu -o tools/_bin/exp_ws -lcuda
#include <cstdio>
#include <vector>

#include "common.cuh"

using namespace mvlm;

struct Args {
  CUtensorMap tm_a, tm_x;
  float* out;  // [128][256]
  float* frag; // [32 threads][16 regs]: tcgen05.ld.16x256b.x4 of lanes 0..15, columns 64..95, by warp 0
  long long* cycles;
  int m, n_mma, ws, n;
};


__device__ __forceinline__ void umma_ws_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.ws.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* sA = smem;          // 128 x 128 B
  uint8_t* sX = smem + 16384;  // 256 x 128 B
  __shared__ uint64_t bar_full, bar_done;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar_full, 1);
    ptx::mbar_init(&bar_done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_base_s, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0 && ptx::elect_one()) {
    ptx::mbar_expect_tx(&bar_full, 16384 + 32768);
    ptx::tma_load_2d(&a.tm_a, &bar_full, sA, 0, 0);
    ptx::tma_load_2d(&a.tm_x, &bar_full, sX, 0, 0);
    ptx::mbar_wait(&bar_full, 0);
    ptx::tc_fence_after();
    const uint32_t idesc = ptx::umma_idesc_bf16(a.m, a.n);
    const long long t0 = clock64();
    for (int i = 0; i < a.n_mma; ++i) {
      const int kk = i & 3;
      if (a.ws)
        umma_ws_bf16(tmem, ptx::umma_desc_sw128(ptx::smem_u32(sA) + kk * 32), ptx::umma_desc_sw128(ptx::smem_u32(sX) + kk * 32),
                     idesc, i > 0 ? 1u : 0u);
      else
        ptx::umma_bf16(tmem, ptx::umma_desc_sw128(ptx::smem_u32(sA) + kk * 32), ptx::umma_desc_sw128(ptx::smem_u32(sX) + kk * 32),
                       idesc, i > 0 ? 1u : 0u);
    }
    ptx::umma_commit(&bar_done);
    ptx::mbar_wait(&bar_done, 0);
    a.cycles[0] = clock64() - t0;
  }
  __syncwarp();
  ptx::mbar_wait(&bar_done, 0);
  ptx::tc_fence_after();
  for (int c = 0; c < 256; c += 16) {
    uint32_t v[16];
    ptx::tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 16; ++j) a.out[(warp * 32 + lane) * 256 + c + j] = __uint_as_float(v[j]);
  }
  if (warp == 0) {
    uint32_t f[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(f[0]), "=r"(f[1]), "=r"(f[2]), "=r"(f[3]), "=r"(f[4]), "=r"(f[5]), "=r"(f[6]), "=r"(f[7]), "=r"(f[8]),
          "=r"(f[9]), "=r"(f[10]), "=r"(f[11]), "=r"(f[12]), "=r"(f[13]), "=r"(f[14]), "=r"(f[15])
        : "r"(tmem + 64)
        : "memory");
    ptx::tmem_ld_wait();
    for (int j = 0; j < 16; ++j) a.frag[lane * 16 + j] = __uint_as_float(f[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 256);
  }
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
  // A[m][k] = (m + 1) if k == 0 else 0 ; X[n][k] = 1 if k == 0 -> D[m][n] = m + 1 for every n (one K=16 step, kk = 0)
  std::vector<__nv_bfloat16> hA(128 * 64), hX(256 * 64);
  for (int m = 0; m < 128; ++m)
    for (int kk = 0; kk < 64; ++kk) hA[m * 64 + kk] = __float2bfloat16(kk == 0 ? float(m + 1) : (kk == 1 ? 1.f / 256.f : 0.f));
  for (int n = 0; n < 256; ++n)
    for (int kk = 0; kk < 64; ++kk) hX[n * 64 + kk] = __float2bfloat16(kk == 0 ? 1.f : (kk == 1 ? float(n) : 0.f));
  __nv_bfloat16 *dA, *dX;
  float* dOut;
  long long* dCyc;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dX, hX.size() * 2);
  cudaMalloc(&dOut, 128 * 256 * 4);
  cudaMalloc(&dCyc, 8);
  float* dFrag;
  cudaMalloc(&dFrag, 32 * 16 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice);
  Args a;
  {
    cuuint64_t gdim[2] = {64, 128};
    cuuint64_t gstr[1] = {128};
    cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
    if (enc(&a.tm_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE))
      return 3;
    cuuint64_t gdim2[2] = {64, 256};
    cuuint32_t box2[2] = {64, 256};
    if (enc(&a.tm_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dX, gdim2, gstr, box2, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE))
      return 4;
  }
  a.out = dOut;
  a.frag = dFrag;
  a.cycles = dCyc;
  const int smem = 16384 + 32768 + 2048;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> h(128 * 256);
  struct Cfg { int m, ws, n; };
  const Cfg cfgs[] = {{128, 0, 256}, {64, 0, 256}, {128, 1, 256}, {64, 1, 256}, {32, 1, 256}, {64, 1, 128}, {32, 1, 128}, {64, 1, 64}, {32, 1, 64}};
  for (const Cfg& c : cfgs) {
    a.m = c.m; a.ws = c.ws; a.n = c.n;
    a.n_mma = 1;
    cudaMemset(dOut, 0xff, 128 * 256 * 4);  // NaN pattern = untouched
    k<<<1, 128, smem>>>(a);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("M=%d ws=%d N=%d: error %s\n", c.m, c.ws, c.n, cudaGetErrorString(cudaGetLastError())); return 5; }
    cudaMemcpy(h.data(), dOut, h.size() * 4, cudaMemcpyDeviceToHost);
    // value = (row + 1) + col / 256 -> decode (row, col) per TMEM (lane, column); print the map compactly:
    // for each 32-lane group: row range and column range found in TMEM columns [0, 16) and the last written column
    printf("M=%d ws=%d N=%d layout:\n", c.m, c.ws, c.n);
    for (int g = 0; g < 4; ++g) {
      for (int l : {0, 1, 15, 16, 31}) {
        const int lane = g * 32 + l;
        int last = -1;
        for (int col = 0; col < 256; ++col) if (h[lane * 256 + col] == h[lane * 256 + col]) last = col;
        printf("  lane %3d:", lane);
        for (int col : {0, 1, 63, 64, 127, 128, 255}) {
          const float v = h[lane * 256 + col];
          if (v != v) { printf(" c%d=--", col); continue; }
          const int row = static_cast<int>(v) - 1, cc = static_cast<int>((v - static_cast<int>(v)) * 256.f + 0.5f);
          printf(" c%d=(%d,%d)", col, row, cc);
        }
        printf(" last=%d\n", last);
      }
    }
    for (int n_mma : {64, 512}) {
      a.n_mma = n_mma;
      long long best = 1ll << 60;
      for (int rep = 0; rep < 3; ++rep) {
        k<<<1, 128, smem>>>(a);
        cudaDeviceSynchronize();
        long long cy;
        cudaMemcpy(&cy, dCyc, 8, cudaMemcpyDeviceToHost);
        if (cy < best) best = cy;
      }
      printf("  M=%d ws=%d N=%d K=16: %d MMAs in %lld cycles = %.1f cycles per MMA\n", c.m, c.ws, c.n, n_mma, best, double(best) / n_mma);
    }
  }
  return 0;
}
