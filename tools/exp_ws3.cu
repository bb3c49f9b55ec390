// Experiment (GPU box): issue cost of the conv kernel's tap loop.  One CTA, operands in shared memory (contents do not
// matter), per "tile" 9 taps x KPT K-steps accumulate into one TMEM tile.  UNROLL = 1: taps and K-steps fully unrolled
// with compile-time descriptor offsets; UNROLL = 0: the loop nest of conv_umma.cu (kx, ky loops, runtime slot index).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -I mvlm_b200/csrc tools/exp_ws3.cu -o tools/_bin/exp_ws3 -lcuda
#include <cstdio>

#include "common.cuh"

using namespace mvlm;

struct Args {
  long long* cycles;
  int n_tiles, kw, kh, halo_px, n_wslots;
};

template <bool WS, int M, int KPT, int UNROLL, int ROWB = 128, int SPIN = 0, int HS = 0>
__global__ void __launch_bounds__(128 + 32 * SPIN, 1) k(const Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* sA = smem;              // 9 slots x 16 KB
  uint8_t* sX = smem + 9 * 16384;  // halo tile 10 x 34 x 128 B = 43520
  __shared__ uint64_t bar_done, t_full[2], t_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar_done, 1);
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&t_full[i], 1); ptx::mbar_init(&t_empty[i], 1); }
    ptx::fence_mbar_init();
  }
  for (int i = threadIdx.x; i < (9 * 16384 + 44032) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_base_s, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0 && ptx::elect_one()) {
    const uint32_t idesc = ptx::umma_idesc_bf16(M, 256);
    constexpr uint32_t swz = ROWB == 128 ? 2u : (ROWB == 64 ? 4u : 6u);
    const uint64_t a_hi = static_cast<uint64_t>(((8u * ROWB) >> 4) | (1u << 14) | (swz << 29)) << 32;
    const uint64_t b_hi = static_cast<uint64_t>(((10u * ROWB) >> 4) | (1u << 14) | (swz << 29)) << 32;
    const uint32_t a_lo0 = ((ptx::smem_u32(sA) >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t b_lo0 = ((ptx::smem_u32(sX) >> 4) & 0x3FFFu) | (1u << 16);
    const long long t0 = clock64();
    int acc = 0;
    uint32_t pacc = 0;
    for (int t = 0; t < a.n_tiles; ++t) {
      if (HS == 2) ptx::mbar_wait(&t_empty[acc], pacc ^ 1);
      const uint32_t d = tmem + acc * 256;
      if (UNROLL) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t w_lo = a_lo0 + tap * (16384 >> 4);
          const uint32_t x_lo = b_lo0 + ((((tap % 3) * 10 + tap / 3) * ROWB) >> 4);
#pragma unroll
          for (int kk = 0; kk < KPT; ++kk) {
            const uint32_t en = (tap == 0 && kk == 0) ? 0u : 1u;
            if (WS) ptx::umma_ws_bf16(d, a_hi | (w_lo + 2 * kk), b_hi | (x_lo + 2 * kk), idesc, en);
            else ptx::umma_bf16(d, a_hi | (w_lo + 2 * kk), b_hi | (x_lo + 2 * kk), idesc, en);
          }
        }
      } else {
        uint32_t accumulate = 0;
        int sw = 0;
        for (int kx = 0; kx < a.kw; ++kx) {
          for (int ky = 0; ky < a.kh; ++ky) {
            const uint32_t w_lo = ((ptx::smem_u32(sA + sw * 16384) >> 4) & 0x3FFFu) | (1u << 16);
            const uint32_t x_lo = b_lo0 + ((static_cast<uint32_t>(ky * a.halo_px + kx) * ROWB) >> 4);
#pragma unroll
            for (int kk = 0; kk < KPT; ++kk) {
              if (WS) ptx::umma_ws_bf16(d, a_hi | (w_lo + 2 * kk), b_hi | (x_lo + 2 * kk), idesc, (kk == 0) ? accumulate : 1u);
              else ptx::umma_bf16(d, a_hi | (w_lo + 2 * kk), b_hi | (x_lo + 2 * kk), idesc, (kk == 0) ? accumulate : 1u);
            }
            accumulate = 1;
            if (++sw == a.n_wslots) sw = 0;
          }
        }
      }
      if (HS >= 1) ptx::umma_commit(&t_full[acc]);
      if (++acc == 2) { acc = 0; pacc ^= 1; }
    }
    ptx::umma_commit(&bar_done);
    ptx::mbar_wait(&bar_done, 0);
    a.cycles[0] = clock64() - t0;
  }
  if (HS == 2 && warp == 1) {
    int acc = 0;
    uint32_t pacc = 0;
    for (int t = 0; t < a.n_tiles; ++t) {
      ptx::mbar_wait(&t_full[acc], pacc);
      ptx::tc_fence_after();
      ptx::tc_fence_before();
      __syncwarp();
      if ((threadIdx.x & 31) == 0) ptx::mbar_arrive(&t_empty[acc]);
      if (++acc == 2) { acc = 0; pacc ^= 1; }
    }
  }
  __syncwarp();
  if (SPIN && warp >= 4) ptx::mbar_wait(&bar_done, 0);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

namespace mvlm {
void set_error(const char*, ...) {}
void count_launch(int) {}
}  // namespace mvlm

template <bool WS, int M, int KPT, int UNROLL, int ROWB = 128, int SPIN = 0, int HS = 0>
void run(long long* dCyc) {
  const int smem = 9 * 16384 + 44032 + 2048;
  cudaFuncSetAttribute(k<WS, M, KPT, UNROLL, ROWB, SPIN, HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  Args a;
  a.cycles = dCyc; a.n_tiles = 100; a.kw = 3; a.kh = 3; a.halo_px = 10; a.n_wslots = 9;
  long long best = 1ll << 60;
  for (int rep = 0; rep < 3; ++rep) {
    k<WS, M, KPT, UNROLL, ROWB, SPIN, HS><<<1, 128 + 32 * SPIN, smem>>>(a);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); exit(5); }
    long long cy;
    cudaMemcpy(&cy, dCyc, 8, cudaMemcpyDeviceToHost);
    if (cy < best) best = cy;
  }
  printf("hs=%d rowb=%d spin=%d ws=%d M=%3d k/tap=%d unroll=%d: %.1f cycles per MMA, %.1f per tap, %.0f per tile\n", HS, ROWB, SPIN, int(WS), M, KPT, UNROLL,
         double(best) / (a.n_tiles * 9 * KPT), double(best) / (a.n_tiles * 9), double(best) / a.n_tiles);
}

int main() {
  cudaFree(0);
  long long* dCyc;
  cudaMalloc(&dCyc, 8);
  run<true, 64, 1, 0, 32, 0, 0>(dCyc);
  run<true, 64, 1, 0, 32, 0, 1>(dCyc);
  run<true, 64, 1, 0, 32, 0, 2>(dCyc);
  run<true, 64, 2, 0, 64, 0, 1>(dCyc);
  run<true, 64, 2, 0, 64, 0, 2>(dCyc);
  run<true, 64, 4, 0, 128, 0, 1>(dCyc);
  run<true, 64, 4, 0, 128, 0, 2>(dCyc);
  run<false, 128, 4, 0, 128, 0, 2>(dCyc);
  return 0;
}
