// Experiment (GPU box): what makes tcgen05.mma.ws slower inside the conv kernel than back to back (84.5 cycles)?
// One CTA, operands in (uninitialised) shared memory, M = 64 .ws or M = 128 plain, N = 256, K = 16 per instruction.
// variant bits: 1 fence::after_thread_sync before every tap, 2 A operand rotates over 9 slots (one per tap), 4 B operand
// shifted per tap inside a 10 x 34 pixel halo tile (SBO = 1280), 8 an mbarrier try_wait (already completed) per tap,
// 16 tcgen05.commit to an mbarrier per tap
// build: nvcc -gencode arch=compute_100a,code=sm_100a -I mvlm_b200/csrc tools/exp_ws2.cu -o tools/_bin/exp_ws2 -lcuda
#include <cstdio>
#include <vector>

#include "common.cuh"

using namespace mvlm;

struct Args {
  long long* cycles;
  int m, ws, n_taps, k_per_tap, variant;
};

__global__ void __launch_bounds__(128, 1) k(const Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* sA = smem;              // 9 slots x 16 KB
  uint8_t* sX = smem + 9 * 16384;  // halo tile 10 x 34 x 128 B = 43520
  __shared__ uint64_t bar_done, bar_dummy, bar_sink;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar_done, 1);
    ptx::mbar_init(&bar_dummy, 1);
    ptx::mbar_init(&bar_sink, 1);
    ptx::fence_mbar_init();
  }
  for (int i = threadIdx.x; i < (9 * 16384 + 44032) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_base_s, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0 && ptx::elect_one()) {
    ptx::mbar_arrive(&bar_dummy);  // phase 0 complete: waits on parity 0 pass at once
    const uint32_t idesc = ptx::umma_idesc_bf16(a.m, 256);
    const uint32_t sbo_b = (a.variant & 4) ? 1280u : 1024u;
    const uint64_t a_hi = static_cast<uint64_t>((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
    const uint64_t b_hi = static_cast<uint64_t>((sbo_b >> 4) | (1u << 14) | (2u << 29)) << 32;
    const uint32_t a_lo0 = ((ptx::smem_u32(sA) >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t b_lo0 = ((ptx::smem_u32(sX) >> 4) & 0x3FFFu) | (1u << 16);
    const long long t0 = clock64();
    uint32_t accumulate = 0;
    for (int t = 0; t < a.n_taps; ++t) {
      const int tap = t % 9;
      if (a.variant & 8) ptx::mbar_wait(&bar_dummy, 0);
      if (a.variant & 1) ptx::tc_fence_after();
      const uint32_t w_lo = a_lo0 + ((a.variant & 2) ? tap * (16384 >> 4) : 0);
      const uint32_t x_lo = b_lo0 + ((a.variant & 4) ? (((tap % 3) * 10 + tap / 3) * 128) >> 4 : 0);
      for (int kk = 0; kk < a.k_per_tap; ++kk) {
        if (a.ws) ptx::umma_ws_bf16(tmem, a_hi | (w_lo + 2 * kk), b_hi | (x_lo + 2 * kk), idesc, (kk == 0) ? accumulate : 1u);
        else ptx::umma_bf16(tmem, a_hi | (w_lo + 2 * kk), b_hi | (x_lo + 2 * kk), idesc, (kk == 0) ? accumulate : 1u);
      }
      accumulate = 1;
      if (a.variant & 16) ptx::umma_commit(&bar_sink);
    }
    ptx::umma_commit(&bar_done);
    ptx::mbar_wait(&bar_done, 0);
    a.cycles[0] = clock64() - t0;
  }
  __syncwarp();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 256);
  }
}

namespace mvlm {
void set_error(const char*, ...) {}
void count_launch(int) {}
}  // namespace mvlm

int main() {
  cudaFree(0);
  long long* dCyc;
  cudaMalloc(&dCyc, 8);
  const int smem = 9 * 16384 + 44032 + 2048;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  struct Cfg { int m, ws; };
  for (const Cfg c : {Cfg{64, 1}, Cfg{32, 1}, Cfg{128, 0}}) {
    for (int kpt : {4, 2, 1}) {
      for (int variant : {0, 1, 2, 4, 6, 8, 16, 7, 15, 31}) {
        Args a;
        a.cycles = dCyc; a.m = c.m; a.ws = c.ws; a.k_per_tap = kpt; a.variant = variant; a.n_taps = 900;
        long long best = 1ll << 60;
        for (int rep = 0; rep < 3; ++rep) {
          k<<<1, 128, smem>>>(a);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 5; }
          long long cy;
          cudaMemcpy(&cy, dCyc, 8, cudaMemcpyDeviceToHost);
          if (cy < best) best = cy;
        }
        printf("M=%3d ws=%d k/tap=%d variant=%2d: %.1f cycles per MMA, %.1f per tap\n", c.m, c.ws, kpt, variant,
               double(best) / (a.n_taps * kpt), double(best) / a.n_taps);
      }
    }
  }
  return 0;
}
