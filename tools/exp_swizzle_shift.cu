// Experiment (GPU box): does a tcgen05 K-major SWIZZLE_128B B-operand descriptor work when
//   * the start address is shifted by whole 128-byte rows (not 1024-byte aligned), and
//   * the stride between 8-row groups (SBO) is not a multiple of 1024 bytes (10 pixels = 1280 B)?
// If yes, one (10 px x 34 rows x 64 ch) halo tile serves all nine 3x3 taps of an 8 px x 32 row
// output tile (today: three 16 px x 18 row tiles, one per horizontal tap).
//
// A = [I_64 ; 0] (128 x 64), so D[m][n] = B-row n, channel m: the result shows which smem row each
// B row actually read.  build: nvcc -gencode arch=compute_100a,code=sm_100a -I mvlm_b200/csrc
//        tools/exp_swizzle_shift.cu -o gpurun_out/exp_swizzle_shift -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

using namespace mvlm;

constexpr int kRows = 34, kPx = 10;

struct Args {
  CUtensorMap tm_a, tm_x;
  float* out;  // [128][256]
  long long* cyc;
  int reps;
  int kx, ky, base_mode, n, ch;  // ch = 64 | 32 | 16 channels per pixel row (SWIZZLE_128B | 64B | 32B)
};

__global__ void __launch_bounds__(128, 1) exp_kernel(const __grid_constant__ Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* sA = smem;                 // 128 x 128 B
  uint8_t* sX = smem + 16384;         // 340 x 128 B
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
    const uint32_t rb = static_cast<uint32_t>(a.ch) * 2u;  // bytes per row
    const uint64_t layout = a.ch == 64 ? 2ull : (a.ch == 32 ? 4ull : 6ull);
    ptx::mbar_expect_tx(&bar_full, 128 * rb + kRows * kPx * rb);
    ptx::tma_load_2d(&a.tm_a, &bar_full, sA, 0, 0);
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        :
        : "r"(ptx::smem_u32(sX)), "l"(reinterpret_cast<uint64_t>(&a.tm_x)), "r"(ptx::smem_u32(&bar_full)), "r"(0),
          "r"(0), "r"(0)
        : "memory");
    ptx::mbar_wait(&bar_full, 0);
    ptx::tc_fence_after();
    if (a.reps < 0) {  // TMA timing: the halo box again and again (L2 hits), one at a time / four in flight
      uint32_t par = 1;
      const long long t0 = clock64();
      for (int r = 0; r < -a.reps; ++r) {
        ptx::mbar_expect_tx(&bar_full, kRows * kPx * rb);
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
            :
            : "r"(ptx::smem_u32(sX)), "l"(reinterpret_cast<uint64_t>(&a.tm_x)), "r"(ptx::smem_u32(&bar_full)), "r"(0),
              "r"(0), "r"(0)
            : "memory");
        ptx::mbar_wait(&bar_full, par);
        par ^= 1;
      }
      a.cyc[0] = clock64() - t0;
      ptx::umma_commit(&bar_done);
    } else {
    const uint32_t idesc = ptx::umma_idesc_bf16(128, a.n);
    const uint32_t xa = ptx::smem_u32(sX) + static_cast<uint32_t>(a.ky * kPx + a.kx) * rb;
    const uint64_t base_off = a.base_mode == 1 ? ((xa >> 7) & 7u) : 0u;
    const long long t0 = clock64();
    for (int rep = 0; rep < a.reps; ++rep)
    for (int k = 0; k < a.ch / 16; ++k) {
      const uint64_t da = static_cast<uint64_t>(((ptx::smem_u32(sA) + k * 32) >> 4) & 0x3FFF) | (1ull << 16) |
                          (static_cast<uint64_t>((8 * rb) >> 4) << 32) | (1ull << 46) | (layout << 61);
      const uint64_t db = static_cast<uint64_t>(((xa + k * 32) >> 4) & 0x3FFF) | (1ull << 16) |
                          (static_cast<uint64_t>((kPx * rb) >> 4) << 32) | (1ull << 46) | (base_off << 49) | (layout << 61);
      ptx::umma_bf16(tmem, da, db, idesc, (k > 0 || rep > 0) ? 1u : 0u);
    }
    ptx::umma_commit(&bar_done);
    if (a.reps > 1) {
      ptx::mbar_wait(&bar_done, 0);
      a.cyc[0] = clock64() - t0;
    }
    }
  }
  __syncwarp();
  ptx::mbar_wait(&bar_done, 0);
  ptx::tc_fence_after();
  for (int c = 0; c < a.n; c += 16) {
    uint32_t v[16];
    ptx::tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 16; ++j) a.out[(warp * 32 + lane) * 256 + c + j] = __uint_as_float(v[j]);
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
  int all_ok = 0, total = 0;
  for (int ch : {64, 32, 16}) {
    const int kCh = ch;
    // A = [I_ch ; 0] (128 x ch): D[m][n] = B-row n, channel m
    std::vector<__nv_bfloat16> hA(128 * kCh), hX(kRows * kPx * kCh);
    for (int m = 0; m < 128; ++m)
      for (int k = 0; k < kCh; ++k) hA[m * kCh + k] = __float2bfloat16(m == k ? 1.f : 0.f);
    // integers < 256 are exact in bf16
    for (int p = 0; p < kRows * kPx; ++p)
      for (int c = 0; c < kCh; ++c) hX[p * kCh + c] = __float2bfloat16(static_cast<float>((p * 7 + c * 3) % 251));
    __nv_bfloat16 *dA, *dX;
    float* dOut;
    cudaMalloc(&dA, hA.size() * 2);
    cudaMalloc(&dX, hX.size() * 2);
    cudaMalloc(&dOut, 128 * 256 * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice);
    Args a;
    const CUtensorMapSwizzle sw = ch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (ch == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    {
      cuuint64_t gdim[2] = {(cuuint64_t)kCh, 128};
      cuuint64_t gstr[1] = {(cuuint64_t)kCh * 2};
      cuuint32_t box[2] = {(cuuint32_t)kCh, 128}, es[2] = {1, 1};
      if (enc(&a.tm_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE))
        return 3;
    }
    {
      cuuint64_t gdim[3] = {(cuuint64_t)kCh, kPx, kRows};
      cuuint64_t gstr[2] = {(cuuint64_t)kCh * 2, (cuuint64_t)kCh * 2 * kPx};
      cuuint32_t box[3] = {(cuuint32_t)kCh, kPx, kRows}, es[3] = {1, 1, 1};
      if (enc(&a.tm_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dX, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE))
        return 4;
    }
    a.out = dOut;
    a.ch = ch;
    a.reps = 1;
    long long* dCyc;
    cudaMalloc(&dCyc, 8);
    a.cyc = dCyc;
    const int smem = 16384 + kRows * kPx * 128 + 2048;
    cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> h(128 * 256);
    for (int n_cols : {256, 64})
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
          const int base_mode = 0;
          a.kx = kx; a.ky = ky; a.base_mode = base_mode; a.n = n_cols;
          cudaMemset(dOut, 0xff, 128 * 256 * 4);
          exp_kernel<<<1, 128, smem>>>(a);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("kx=%d ky=%d base_mode=%d: CUDA error %s\n", kx, ky, base_mode, cudaGetErrorString(e));
            return 5;
          }
          cudaMemcpy(h.data(), dOut, h.size() * 4, cudaMemcpyDeviceToHost);
          int bad = 0, first_bad = -1;
          for (int n = 0; n < n_cols; ++n) {
            const int p = (n / 8 + ky) * kPx + (n % 8) + kx;
            for (int m = 0; m < kCh; ++m) {
              const float want = static_cast<float>((p * 7 + m * 3) % 251);
              if (h[m * 256 + n] != want) {
                if (first_bad < 0) first_bad = n * 64 + m;
                ++bad;
              }
            }
          }
          ++total;
          all_ok += bad == 0;
          printf("ch=%d N=%d kx=%d ky=%d: %s (%d of %d mismatches", ch, n_cols, kx, ky, bad ? "MISMATCH" : "ok", bad, n_cols * kCh);
          if (bad) {
            const int n = first_bad / 64, m = first_bad % 64;
            printf("; first at n=%d m=%d got %g", n, m, h[m * 256 + n]);
          }
          printf(")\n");
        }
    // timing: 64 / (ch/16) repetitions = 64 MMAs back to back per variant, tap (1,1) and tap (0,0)
    for (int tap = 0; tap < 2; ++tap) {
      a.kx = tap; a.ky = tap; a.base_mode = 0; a.n = 256; a.reps = 256 / (ch / 16);
      exp_kernel<<<1, 128, smem>>>(a);
      cudaDeviceSynchronize();
      long long c = 0;
      cudaMemcpy(&c, dCyc, 8, cudaMemcpyDeviceToHost);
      printf("ch=%d (row %d B) N=256 tap(%d,%d): %.1f cycles per MMA (256 MMAs)\n", ch, ch * 2, tap, tap, double(c) / 256.0);
    }
    a.reps = -64;
    exp_kernel<<<1, 128, smem>>>(a);
    cudaDeviceSynchronize();
    {
      long long c = 0;
      cudaMemcpy(&c, dCyc, 8, cudaMemcpyDeviceToHost);
      printf("ch=%d: TMA box (%d ch x 10 px x 34 rows = %d B, %d-byte rows), one at a time, L2-hot: %.0f cycles per load\n", ch, ch,
             340 * ch * 2, ch * 2, double(c) / 64.0);
    }
    a.reps = 1;
    cudaFree(dA); cudaFree(dX); cudaFree(dOut);
  }
  printf("%d of %d variants exact\n", all_ok, total);
  return 0;
}
