// tcgen05/TMEM implicit-GEMM convolution, see conv_umma.cuh for the contract.
//
// GEMM orientation ("weights are the A operand"):
//     D[cout (M=128 TMEM lanes) x pixels (N=256 TMEM columns)] += W[cout x K] * X[K x pixels]
// A single-CTA tcgen05.mma streams its A operand (128 rows x K=16) from shared memory at ~64 B/clk,
// i.e. >= ~75 cycles per instruction however small N is (measured: 79/84/94 cycles at N=32/64/128
// against a 16/32/64-cycle math floor).  With the 128 output channels as A and a whole 256-pixel
// tile as B (N=256) the same instruction does 128 cycles of math (measured 145), so the A stream is
// hidden -- see tools/exp_mma_only.py and DESIGN.md section 4.
//
// One persistent CTA per SM, 11 warps:
//   warp 0      TMA producer   X: ONE halo tile (8+KW-1 px x 32+KH-1 rows x 64 ch) per cin-chunk
//   warp 10     TMA producer   W: one [<=128 cout][64 cin] tile per (cin-chunk, kx, ky)
//   warp 1      MMA issuer     one elected lane issues tcgen05.mma (M=128, N=256, K=16); owns TMEM
//   warps 2..9  epilogue       tcgen05.ld (lane = channel, 16 pixels) -> smem transpose ->
//                              bias/BN/ReLU/residual on (pixel, 8 channels) vectors -> 16-byte stores
//
// Output tile = 8 px x 32 rows (N = 256; 16/8/4 rows for maps lower than 32).  The halo tile of a
// 64-channel chunk is stored pixel-major (128-byte rows, SWIZZLE_128B); the B operand of tap (kx, ky)
// is a descriptor into that SAME tile: start address shifted by (ky*(8+KW-1) + kx) rows of 128 bytes,
// stride between 8-row groups (SBO) = one halo row of (8+KW-1)*128 bytes.  Neither is a multiple of
// the 1024-byte swizzle atom: the hardware applies the swizzle XOR to the absolute shared-memory
// address (measured, tools/exp_swizzle_shift.cu: all taps / all swizzle widths bit-exact with base
// offset 0), so one TMA load serves all KW*KH taps -- 1.33x the tile's own pixels instead of 3.4x
// (one load per horizontal tap), which matters because the L2 -> SM fabric is the binding resource
// (10 TB/s on the 256->128 layers before this change).  Zero padding, ragged edges and maps smaller
// than the tile come from TMA out-of-bounds zero fill.
//
// Accumulators: 2 pipeline stages x 256 fp32 columns of TMEM: the epilogue of tile i overlaps the
// MMAs of tile i+1.
#include "conv_umma.cuh"

#include <stdlib.h>

#include <algorithm>

namespace mvlm {

namespace {

constexpr int kThreads = 352;                          // warps: 0 halo TMA, 1 MMA, 2..9 epilogue, 10 weight TMA
constexpr int kEpiWarps = 8;
constexpr int kTileW = 8;                               // output tile: 8 px wide ...
constexpr int kMaxTileH = 32;                           // ... and up to 32 rows high (N = 256)
constexpr int kMTile = 128;                             // output channels per CTA tile (UMMA M)
constexpr int kMaxHSlots = 4;                           // halo (activation) ring, depth chosen per layer
constexpr int kMaxWSlots = 12;                          // weight ring (16 KB slots at M = 128, 8 KB at M = 64)
constexpr int kPoolBytes = 198 * 1024;                  // both rings
constexpr int kStageFloats = 32 * 20;                   // per-warp transpose buffer: 16 px x (32 ch + 4 pad) | 32 px x (16 + 4)
constexpr int kMaxCout = 256;
constexpr int kTraceTiles = 64;                        // debug timeline: tiles traced on CTA 0

struct __align__(8) Barriers {
  uint64_t h_full[kMaxHSlots];
  uint64_t h_empty[kMaxHSlots];
  uint64_t w_full[kMaxWSlots];
  uint64_t w_empty[kMaxWSlots];
  uint64_t t_full[2];
  uint64_t t_empty[2];
  uint32_t tmem_base;
};
static_assert(sizeof(Barriers) <= 512, "barrier block");

// Per-output-channel epilogue parameters staged once per CTA in shared memory.
struct EpiParams {
  float bias[kMaxCout];
  float mid_s[kMaxCout];
  float mid_t[kMaxCout];
  float pre_s[kMaxCout];
  float pre_t[kMaxCout];
  float post_s[kMaxCout];
  float post_t[kMaxCout];
};

constexpr int kSmemBytes = kPoolBytes + kEpiWarps * kStageFloats * 4 + 512 + static_cast<int>(sizeof(EpiParams)) +
                           1024 /*align*/;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ uint32_t order_f32(float f) {
  const uint32_t b = __float_as_uint(f);
  // negative: ~b, else b | 0x80000000  ==  b ^ (sign-extended sign | 0x80000000)
  return b ^ (static_cast<uint32_t>(static_cast<int32_t>(b) >> 31) | 0x80000000u);
}

__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(f[2], f[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(f[4], f[5]);
  __nv_bfloat162 d = __floats2bfloat162_rn(f[6], f[7]);
  uint4 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  r.z = *reinterpret_cast<uint32_t*>(&c);
  r.w = *reinterpret_cast<uint32_t*>(&d);
  return r;
}
// relu + round to bf16 in one instruction per pair (cvt.rn.relu.bf16x2.f32: first source -> upper half)
__device__ __forceinline__ uint32_t relu_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// relu(f * s + t) -> bf16 x 8
__device__ __forceinline__ uint4 affine_relu_pack8(const float (&f)[8], const float (&sc)[8], const float (&sh)[8]) {
  uint4 r;
  r.x = relu_bf16x2(fmaf(f[0], sc[0], sh[0]), fmaf(f[1], sc[1], sh[1]));
  r.y = relu_bf16x2(fmaf(f[2], sc[2], sh[2]), fmaf(f[3], sc[3], sh[3]));
  r.z = relu_bf16x2(fmaf(f[4], sc[4], sh[4]), fmaf(f[5], sc[5], sh[5]));
  r.w = relu_bf16x2(fmaf(f[6], sc[6], sh[6]), fmaf(f[7], sc[7], sh[7]));
  return r;
}
__device__ __forceinline__ void add8(const uint4& q, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __bfloat1622float2(h[j]);
    f[2 * j] += t.x;
    f[2 * j + 1] += t.y;
  }
}
__device__ __forceinline__ uint4 max_bf16x8(const uint4& a, const uint4& b) {
  uint4 r;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int j = 0; j < 4; ++j) pr[j] = __hmax2(pa[j], pb[j]);
  return r;
}
__device__ __forceinline__ void lds8(const float* src, float (&v)[8]) {
  const float4 a = reinterpret_cast<const float4*>(src)[0];
  const float4 b = reinterpret_cast<const float4*>(src)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// role timing: wait on an mbarrier and add the stalled cycles to `acc` when profiling is on
__device__ __forceinline__ void timed_wait(uint64_t* b, uint32_t parity, bool prof, long long& acc) {
  if (!prof) { ptx::mbar_wait(b, parity); return; }
  const long long t0 = clock64();
  ptx::mbar_wait(b, parity);
  acc += clock64() - t0;
}

struct TileCoord {
  int mt, tx, ty, img;
};
__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int t) {
  TileCoord c;
  c.mt = t % p.n_nt;
  int r = t / p.n_nt;
  c.tx = r % p.tiles_x;
  r /= p.tiles_x;
  c.ty = r % p.tiles_y;
  c.img = r / p.tiles_y;
  return c;
}

// Epilogue feature flags (template parameter F): code for a feature is only generated when its bit is set,
// which keeps the per-row instruction count of the hot variants low (the epilogue is issue-bound).
enum : int { F_PRE = 1, F_RES1 = 2, F_RES2 = 4, F_RAW = 8, F_POST = 16, F_F32 = 32 /* fp32 NCHW map */, F_ARGMAX = 64,
              F_MID = 128 /* affine + ReLU right after the bias */, F_POOL = 256 /* raw/post at half resolution */,
              F_UP = 512 /* + nearest-x2 up-sampled half-resolution tensor */,
              F_M64 = 1024 /* cout <= 64: UMMA M = 64, 16 channels per TMEM lane group */ };
constexpr int F_HEAD = F_F32 | F_ARGMAX;

template <int F>
__global__ void __launch_bounds__(kThreads, 1) conv_umma_kernel(const __grid_constant__ ConvParams p) {
  constexpr bool ARGMAX = (F & F_ARGMAX) != 0;  // arg-max variants use the contiguous tile schedule
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment: TMA and UMMA agree on the SWIZZLE_128B XOR pattern through the absolute address
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* h_slots = smem;
  uint8_t* w_slots = smem + p.n_hslots * p.h_slot_bytes;
  float* stage_all = reinterpret_cast<float*>(smem + kPoolBytes);
  Barriers* bar = reinterpret_cast<Barriers*>(reinterpret_cast<uint8_t*>(stage_all) + kEpiWarps * kStageFloats * 4);
  EpiParams* ep = reinterpret_cast<EpiParams*>(reinterpret_cast<uint8_t*>(bar) + 512);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const ConvShape& s = p.s;
  const ConvEpilogue& e = p.e;
  const int n_chunks = (s.cin + 63) >> 6;
  const int n_hslots = p.n_hslots, n_wslots = p.n_wslots;
  const int halo_px = kTileW + s.kw - 1;            // pixels per halo row
  const int halo_rows = p.tile_h + s.kh - 1;
  // cout <= 64 runs with UMMA M = 64 (same cycles per instruction as M = 128, measured in tools/exp_m64.cu, but
  // half the weight bytes through shared memory, which is the binding resource): output row i sits in TMEM lane
  // 32 * (i / 16) + i % 16, i.e. every lane group (every epilogue warp) holds 16 channels in its lanes 0..15.
  constexpr bool M64 = (F & F_M64) != 0;
  constexpr int kM = M64 ? 64 : kMTile;
  const int w_rows = s.cout_pad < kM ? s.cout_pad : kM;  // weight rows actually loaded per tile
  // cout < M: the weight rows are replicated `rep` times along M, so that all four TMEM lane groups (and
  // therefore all eight epilogue warps) hold the same channels and split the tile's pixel rows instead.
  const int rep = kM / w_rows;
  const uint32_t w_slot_bytes = static_cast<uint32_t>(kM) * 128u;
  // stationary weights: all (chunk, tap) tiles of the layer fit in the ring -> loaded once per CTA, never released
  const bool w_stat = p.w_stationary != 0;

  // Tile schedule shared by the three roles.  Default: round-robin (neighbouring CTAs work on neighbouring
  // tiles -> halo / weight reuse in L2).  ARGMAX: contiguous ranges, so that a CTA stays within one image
  // for ~40 tiles and keeps its running arg-max in registers.
  int t_begin, t_end, t_step;
  if (ARGMAX) {
    const int per = (p.total_tiles + gridDim.x - 1) / gridDim.x;
    t_begin = blockIdx.x * per;
    t_end = min(p.total_tiles, t_begin + per);
    t_step = 1;
  } else {
    t_begin = blockIdx.x; t_end = p.total_tiles; t_step = gridDim.x;
  }

  for (int i = threadIdx.x; i < s.cout_pad; i += kThreads) {
    ep->bias[i] = e.bias ? e.bias[i] : 0.f;
    ep->mid_s[i] = e.mid_scale ? e.mid_scale[i] : 1.f;
    ep->mid_t[i] = e.mid_scale ? e.mid_shift[i] : 0.f;
    ep->pre_s[i] = e.out_pre ? e.pre_scale[i] : 0.f;
    ep->pre_t[i] = e.out_pre ? e.pre_shift[i] : 0.f;
    ep->post_s[i] = e.out_post ? e.post_scale[i] : 0.f;
    ep->post_t[i] = e.out_post ? e.post_shift[i] : 0.f;
  }
  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&p.tm_a);
    ptx::tma_prefetch_desc(&p.tm_b);
    if (p.tail) { ptx::tma_prefetch_desc(&p.tm_a2); ptx::tma_prefetch_desc(&p.tm_b2); }
    for (int i = 0; i < n_hslots; ++i) {
      ptx::mbar_init(&bar->h_full[i], 1);
      ptx::mbar_init(&bar->h_empty[i], 1);
    }
    for (int i = 0; i < n_wslots; ++i) {
      ptx::mbar_init(&bar->w_full[i], 1);
      ptx::mbar_init(&bar->w_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bar->t_full[i], 1);
      ptx::mbar_init(&bar->t_empty[i], kEpiWarps);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(&bar->tmem_base, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  // Programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch, per-channel parameters,
  // none of which another launch of the plan writes) may overlap the tail of the previous kernel in the stream;
  // every read or write of an activation tensor comes after this wait.  The next kernel may start its own
  // prologue as soon as SMs free up.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const uint32_t tmem_base = bar->tmem_base;
  const bool prof = p.prof != nullptr;
  const long long t_kernel0 = clock64();
  long long w0 = 0, w1 = 0;
  // per-tile timeline of CTA 0 (first kTraceTiles tiles), cycles since kernel start:
  //   [0] producer: halo load of chunk 0 issued   [1] producer: last weight load issued
  //   [2] MMA: accumulator stage acquired         [3] MMA: first halo tile landed      [4] MMA: all MMAs issued
  //   [5] epilogue warp 2: accumulator ready      [6] epilogue warp 2: stage released
  //   epilogue warp 2, unit r = 0 / 1: [8 + 4r] accumulator columns in registers, [9 + 4r] transposed to shared
  //   memory, [10 + 4r] outputs computed and stores issued, [11 + 4r] unit done (rolling prefetch issued)
  long long* const trace = (prof && blockIdx.x == 0) ? p.prof + kNumSMs * 8 : nullptr;
  int trace_i = 0;
#define MVLM_TRACE(slot)                                                                  \
  do {                                                                                    \
    if (trace && trace_i < kTraceTiles) trace[trace_i * 16 + (slot)] = clock64() - t_kernel0; \
  } while (0)

  if (warp == 0) {
    // ===================== TMA producer: activations =====================
    // Its own thread, so that the halo prefetch runs n_hslots chunks ahead of the MMAs whatever the state of the
    // weight ring (measured: behind the weight loads in one queue the next halo tile was issued ~4k cycles before
    // it was needed, about the DRAM + queueing latency of the load, and the tensor pipe waited for it every tile).
    if (ptx::elect_one() && p.debug_mode == 0) {
      int sh = 0;
      uint32_t ph = 0;
      for (int t = t_begin; t < t_end; t += t_step) {
        const TileCoord tc = decode_tile(p, t);
        const int x0 = tc.tx * kTileW + s.x_off0;
        const int y0 = tc.ty * p.tile_h + s.y_off0;
        for (int c = 0; c < n_chunks; ++c) {
          // the last chunk may be a narrow tail (16 / 32 channels) with its own tensor map: rows of 32 / 64 bytes
          const bool is_tail = p.tail != 0 && c == n_chunks - 1;
          const uint32_t row_b = is_tail ? static_cast<uint32_t>(p.tail) * 2u : 128u;  // bytes per pixel
          timed_wait(&bar->h_empty[sh], ph ^ 1, prof, w0);
          ptx::mbar_expect_tx(&bar->h_full[sh], static_cast<uint32_t>(halo_rows * halo_px) * row_b);
          // one halo tile for all KW x KH taps of this chunk
          ptx::tma_load_4d(is_tail ? &p.tm_a2 : &p.tm_a, &bar->h_full[sh], h_slots + sh * p.h_slot_bytes, c * 64, x0, y0, tc.img);
          if (c == 0) MVLM_TRACE(0);
          if (++sh == n_hslots) { sh = 0; ph ^= 1; }
        }
        ++trace_i;
      }
      if (prof) { p.prof[blockIdx.x * 8 + 0] = w0; p.prof[blockIdx.x * 8 + 7] = clock64() - t_kernel0; }
    }
  } else if (warp == 10) {
    // ===================== TMA producer: weights =====================
    if (ptx::elect_one() && p.debug_mode == 0) {
      int sw = 0;
      uint32_t pw = 0;
      for (int t = t_begin; t < t_end; t += t_step) {
        const int mt = t % p.n_nt;
        for (int c = 0; c < n_chunks; ++c) {
          const bool is_tail = p.tail != 0 && c == n_chunks - 1;
          const void* tmb = is_tail ? &p.tm_b2 : &p.tm_b;
          const uint32_t row_b = is_tail ? static_cast<uint32_t>(p.tail) * 2u : 128u;  // bytes per weight row
          const uint32_t wb = static_cast<uint32_t>(w_rows) * row_b;
          for (int tap = 0; tap < s.kw * s.kh; ++tap) {  // tap = kx * KH + ky
            if (!w_stat) timed_wait(&bar->w_empty[sw], pw ^ 1, prof, w1);
            ptx::mbar_expect_tx(&bar->w_full[sw], wb * rep);
            for (int q = 0; q < rep; ++q)  // small cout: the same rows again for the other TMEM lane groups
              ptx::tma_load_2d(tmb, &bar->w_full[sw], w_slots + sw * w_slot_bytes + q * w_rows * row_b,
                               tap * s.cin + c * 64, mt * kM);
            if (++sw == n_wslots) { sw = 0; pw ^= 1; }
          }
        }
        MVLM_TRACE(1);
        ++trace_i;
        if (w_stat) break;  // one pass over the layer's weights is all there is
      }
      if (prof) p.prof[blockIdx.x * 8 + 1] = w1;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // elect.sync (not `lane == 0`): the compiler then knows a single lane is active and emits the
    // UTCHMMA / UTCBAR uniform-datapath instructions without a per-lane serialisation loop.
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::umma_idesc_bf16(kM, p.tile_h * kTileW);
      // descriptor = {lo: start>>4 | LBO, hi: SBO | version | swizzle}; only lo changes per MMA.
      //   A (weights): 8-row groups 8 rows apart.  B (halo tile): 8-row group = the 8 pixels of one image row,
      //   groups one halo row (halo_px pixels) apart.
      int sh = 0, sw = 0;
      uint32_t ph = 0, pw = 0;
      int acc = 0;
      uint32_t pacc = 0;
      for (int t = t_begin; t < t_end; t += t_step) {
        timed_wait(&bar->t_empty[acc], pacc ^ 1, prof, w1);
        ptx::tc_fence_after();
        MVLM_TRACE(2);
        const uint32_t d = tmem_base + static_cast<uint32_t>(acc * 256);
        uint32_t accumulate = 0;
        for (int c = 0; c < n_chunks; ++c) {
          const int rem = s.cin - c * 64;
          const int nk = rem >= 64 ? 4 : (rem >> 4);
          // tail chunk: SWIZZLE_32B (16 ch: 32-byte rows) or SWIZZLE_64B (32 ch: 64-byte rows)
          const bool is_tail = p.tail != 0 && c == n_chunks - 1;
          const uint32_t row_b = is_tail ? static_cast<uint32_t>(p.tail) * 2u : 128u;
          const uint32_t swz = !is_tail ? 2u : (p.tail == 16 ? 6u : 4u);
          const uint64_t a_hi = static_cast<uint64_t>(((8u * row_b) >> 4) | (1u << 14) | (swz << 29)) << 32;
          const uint64_t b_hi = static_cast<uint64_t>(((static_cast<uint32_t>(halo_px) * row_b) >> 4) | (1u << 14) | (swz << 29)) << 32;
          if (p.debug_mode == 0) timed_wait(&bar->h_full[sh], ph, prof, w0);
          if (c == 0) MVLM_TRACE(3);
          const uint32_t h_lo = ((ptx::smem_u32(h_slots + sh * p.h_slot_bytes) >> 4) & 0x3FFFu) | (1u << 16);
          for (int kx = 0; kx < s.kw; ++kx) {
            for (int ky = 0; ky < s.kh; ++ky) {
              if (p.debug_mode == 0) timed_wait(&bar->w_full[sw], w_stat ? 0u : pw, prof, w0);
              ptx::tc_fence_after();
              const uint32_t w_lo = ((ptx::smem_u32(w_slots + sw * w_slot_bytes) >> 4) & 0x3FFFu) | (1u << 16);
              // tap (kx, ky) = the same halo tile shifted by ky halo rows + kx pixels
              const uint32_t x_lo = h_lo + ((static_cast<uint32_t>(ky * halo_px + kx) * row_b) >> 4);
              if (nk == 4) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  ptx::umma_bf16(d, a_hi | (w_lo + 2 * k), b_hi | (x_lo + 2 * k), idesc, (k == 0) ? accumulate : 1u);
              } else {
                for (int k = 0; k < nk; ++k)
                  ptx::umma_bf16(d, a_hi | (w_lo + 2 * k), b_hi | (x_lo + 2 * k), idesc, (k == 0) ? accumulate : 1u);
              }
              accumulate = 1;
              if (p.debug_mode == 0 && !w_stat) ptx::umma_commit(&bar->w_empty[sw]);
              if (++sw == n_wslots) { sw = 0; pw ^= 1; }
            }
          }
          if (p.debug_mode == 0) ptx::umma_commit(&bar->h_empty[sh]);
          if (++sh == n_hslots) { sh = 0; ph ^= 1; }
        }
        ptx::umma_commit(&bar->t_full[acc]);
        MVLM_TRACE(4);
        ++trace_i;
        if (++acc == 2) { acc = 0; pacc ^= 1; }
      }
      if (prof) { p.prof[blockIdx.x * 8 + 2] = w0; p.prof[blockIdx.x * 8 + 3] = w1; p.prof[blockIdx.x * 8 + 4] = clock64() - t_kernel0; }
    }
  } else {
    // ===================== epilogue =====================
    // Accumulator column n = 8 * (row in tile) + (pixel in row).  One "unit" (one TMEM load, one transpose) =
    //   M = 128: 16 columns = 2 image rows x 8 px of the warp's 32 channels,
    //   M = 64 : 32 columns = 4 image rows x 8 px of the warp's 16 channels (lanes 0..15 of the lane group),
    // i.e. 512 values either way.  The epilogue is latency-bound (two warps per scheduler, dependent
    // TMEM -> shared -> registers -> global chain per unit), so fewer, fatter units matter.
    constexpr int kChGrp = M64 ? 16 : 32;      // channels per TMEM lane group
    constexpr int kPitch = M64 ? 20 : 36;      // floats per pixel row of the transpose buffer
    constexpr int kCols = M64 ? 32 : 16;       // accumulator columns per unit
    constexpr int kUnitRows = kCols / 8;       // image rows per unit
    constexpr int kPassStep = M64 ? 2 : 1;     // image rows between my pixel in pass 0 and in pass 1
    constexpr bool kFrag = M64 && (F & F_HEAD) == 0;  // M = 64 accumulators read with the 16x256b shape
    static_assert(kCols * kPitch <= kStageFloats, "transpose buffer");
    const int ew = warp - 2;
    const int lane_grp = warp & 3;   // TMEM lanes this warp may read: 32*(warp%4)..
    const int n_cgrp = 4 / rep;      // distinct channel groups along M
    const int cgrp = lane_grp % n_cgrp;
    const int replica = lane_grp / n_cgrp;
    const int n_units = max(1, p.tile_h / kUnitRows);
    const int upw = max(1, n_units / (2 * rep));              // units per warp
    const int u_begin = ((ew >> 2) * rep + replica) * upw;    // first unit handled by this warp
    float* stage = stage_all + ew * kStageFloats;
    int acc = 0;
    uint32_t pacc = 0;
    // HEAD: lane = channel -> one running (ordered value, ~index) pair per thread
    uint32_t best_hi = 0u, best_lo = 0u;
    int cur_img = -1;
    const int oh = s.h * e.up_sy, ow = s.w * e.up_sx;
    // channel-major role (TMEM load, bias, transpose store): lane = channel; M = 64 -> lanes 0..15 only.
    // pixel-major role after the transpose, two passes per unit:
    //   M = 128: lane -> (pixel column pj = lane/4, channels (lane%4)*8 .. +7), unit row ip in pass ip
    //   M = 64 : lane -> (pixel column pj = (lane/2)%8, channels (lane%2)*8 .. +7), unit row lane/16 + 2*ip in pass ip
    const bool cm_lane = !M64 || lane < 16;
    const int pj = M64 ? ((lane >> 1) & 7) : (lane >> 2);
    const int my_i = M64 ? (lane >> 4) : 0;
    const int cq = M64 ? (lane & 1) * 8 : (lane & 3) * 8;
    constexpr bool kBf16Out = (F & (F_PRE | F_RAW | F_POST)) != 0;
    // byte strides of one image row in every tensor the epilogue touches: an access of unit r, pass ip is then
    // (per-tile base pointer) + (kUnitRows * r + kPassStep * ip) * stride with a compile-time row offset
    const uint32_t rs_pre = static_cast<uint32_t>(s.w * e.pre_cs) * 2u, rs_res1 = static_cast<uint32_t>(s.w * e.res1_cs) * 2u;
    const uint32_t rs_res2 = static_cast<uint32_t>(s.w * e.res2_cs) * 2u;
    const uint32_t rs_up = static_cast<uint32_t>((s.w >> 1) * e.up_cs) * 2u;
    // F_POOL: raw / post live at half resolution
    const uint32_t rs_raw = static_cast<uint32_t>(((F & F_POOL) ? (s.w >> 1) : s.w) * e.raw_cs) * 2u;
    const uint32_t rs_post = static_cast<uint32_t>(((F & F_POOL) ? (s.w >> 1) : s.w) * e.post_cs) * 2u;
    // Residual inputs do not depend on the accumulator: they are fetched in BATCHES before they are needed -- all
    // units of a tile while its MMAs still run when they fit in ~64 registers, else half of them then and the other
    // half once the first are consumed.  (Loads issued one by one while earlier ones are being consumed do not work:
    // the few hardware scoreboards are shared, so every use then waits for the newest load; measured.)
    constexpr bool kHasRes = (F & (F_RES1 | F_RES2 | F_UP)) != 0;
    constexpr int kMaxUpw = M64 ? 4 : 8;  // units per warp at N = 256
    constexpr int kResRegs = 4 * (2 * (((F & F_RES1) ? 1 : 0) + ((F & F_RES2) ? 1 : 0)) + (M64 ? 2 : 1) * ((F & F_UP) ? 1 : 0));
    constexpr int kPref = !kHasRes ? 1 : (kResRegs * kMaxUpw <= 64 ? kMaxUpw : kMaxUpw / 2);  // units per batch
    uint4 r1[kPref][2], r2[(F & F_RES2) ? kPref : 1][2], ru[(F & F_UP) ? kPref : 1][M64 ? 2 : 1];
    for (int t = t_begin; t < t_end; t += t_step) {
      const TileCoord tc = decode_tile(p, t);
      const int m0 = tc.mt * kM;
      const int c_lane = m0 + cgrp * kChGrp + lane;       // channel-major role: my output channel
      const int c0 = m0 + cgrp * kChGrp + cq;             // pixel-major role: first of my 8 channels
      const bool grp_active = m0 + cgrp * kChGrp < s.cout_pad;
      const int y_first = tc.ty * p.tile_h + kUnitRows * u_begin;  // first image row handled by this warp
      const bool rows_active = u_begin < n_units && y_first < s.h;
      const bool ch_ok = c0 < s.cout_pad;  // weight rows beyond cout_pad are never loaded
      const int xa = tc.tx * kTileW + pj;
      const bool vx = ch_ok && xa < s.w;
      // element index of my pixel in pass 0 of the first unit: (img, y_first + my_i, xa); 32-bit: pixel count x
      // channel stride < 2^31 (checked in conv_plan)
      const uint32_t pix0 = (static_cast<uint32_t>(tc.img) * s.h + y_first + my_i) * s.w + xa;
      // element index of the half-resolution pixel (img, y_first/2, xa/2): F_POOL outputs, F_UP input
      const uint32_t ppix0 = (static_cast<uint32_t>(tc.img) * (s.h >> 1) + (y_first >> 1)) * (s.w >> 1) + (xa >> 1);
      if ((F & F_ARGMAX) && tc.img != cur_img) {
        if (cur_img >= 0 && best_hi != 0u && cm_lane && c_lane < e.cout_real)
          atomicMax(e.argmax_keys + static_cast<size_t>(cur_img) * e.cout_real + c_lane,
                    (static_cast<unsigned long long>(best_hi) << 32) | best_lo);
        best_hi = 0u; best_lo = 0u;
        cur_img = tc.img;
      }
      // per-tile base pointers of my (pixel, 8 channels) in the first unit
      const uint32_t opix0 = (F & F_POOL) ? ppix0 : pix0;
      uint8_t* const b_pre = (F & F_PRE) ? reinterpret_cast<uint8_t*>(e.out_pre + e.pre_co + c0 + static_cast<size_t>(pix0) * e.pre_cs) : nullptr;
      uint8_t* const b_raw = (F & F_RAW) ? reinterpret_cast<uint8_t*>(e.out_raw + e.raw_co + c0 + static_cast<size_t>(opix0) * e.raw_cs) : nullptr;
      uint8_t* const b_post = (F & F_POST) ? reinterpret_cast<uint8_t*>(e.out_post + e.post_co + c0 + static_cast<size_t>(opix0) * e.post_cs) : nullptr;
      // rows at / below my first pixel that exist in the image (0 when my pixel column / channels do not):
      // unit r, pass ip is valid iff kUnitRows * r + kPassStep * ip < n_rows_ok
      const int n_rows_ok = vx ? s.h - y_first - my_i : 0;
      const uint8_t* const b_res1 = (F & F_RES1) ? reinterpret_cast<const uint8_t*>(e.res1 + e.res1_co + c0 + static_cast<size_t>(pix0) * e.res1_cs) : nullptr;
      const uint8_t* const b_res2 = (F & F_RES2) ? reinterpret_cast<const uint8_t*>(e.res2 + e.res2_co + c0 + static_cast<size_t>(pix0) * e.res2_cs) : nullptr;
      const uint8_t* const b_up = (F & F_UP) ? reinterpret_cast<const uint8_t*>(e.res_up + e.up_co + c0 + static_cast<size_t>(ppix0) * e.up_cs) : nullptr;
      auto prefetch_batch = [&](int u0) {  // units u0 .. u0 + kPref - 1 (u0 compile-time after unrolling)
#pragma unroll
        for (int q = 0; q < kPref; ++q) {
          const int u = u0 + q;
          if (u < upw) {
#pragma unroll
            for (int ip = 0; ip < 2; ++ip) {
              const int row = kUnitRows * u + kPassStep * ip;
              if (row < n_rows_ok) {
                if (F & F_RES1) r1[q][ip] = *reinterpret_cast<const uint4*>(b_res1 + static_cast<size_t>(row * rs_res1));
                if (F & F_RES2) r2[q][ip] = *reinterpret_cast<const uint4*>(b_res2 + static_cast<size_t>(row * rs_res2));
                // nearest x2: rows 2k, 2k+1 and columns xa, xa^1 all read low-res pixel (k, xa/2);
                // M = 128: both passes share one low-res row per unit, M = 64: pass ip reads low-res row 2u + ip
                if ((F & F_UP) && (M64 || ip == 0))
                  ru[q][M64 ? ip : 0] = *reinterpret_cast<const uint4*>(b_up + static_cast<size_t>((row >> 1) * rs_up));
              }
            }
          }
        }
      };
      if (kHasRes && grp_active && rows_active) prefetch_batch(0);
      // per-channel parameters of my 8 channels (pixel-major role) and my channel (channel-major role)
      float pre_s[8], pre_t[8], post_s[8], post_t[8];
      if (F & F_PRE) { lds8(ep->pre_s + (ch_ok ? c0 : 0), pre_s); lds8(ep->pre_t + (ch_ok ? c0 : 0), pre_t); }
      if (F & F_POST) { lds8(ep->post_s + (ch_ok ? c0 : 0), post_s); lds8(ep->post_t + (ch_ok ? c0 : 0), post_t); }
      const float bias_c = ep->bias[c_lane < kMaxCout ? c_lane : 0];
      const float mid_s_c = ep->mid_s[c_lane < kMaxCout ? c_lane : 0], mid_t_c = ep->mid_t[c_lane < kMaxCout ? c_lane : 0];
      timed_wait(&bar->t_full[acc], pacc, prof, w0);
      ptx::tc_fence_after();
      if (warp == 2 && lane == 0) MVLM_TRACE(5);
      if (grp_active && rows_active) {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) +
                               static_cast<uint32_t>(acc * 256 + u_begin * kCols);
        // one register buffer: the load of unit r+1 is issued as soon as unit r has left the registers
        // M = 64 without channel-major consumers: the 16 data lanes of the lane group are read with the 16x256b shape,
        // which spreads them over all 32 threads (2 channels x 16 pixels each): half the registers and, above all,
        // full 32-lane transpose stores -- the shared-memory pipe is what bounds these layers
        uint32_t vr[kFrag ? 16 : kCols];
        auto tmem_load = [&](int u) {
          if constexpr (kFrag) ptx::tmem_ld_16x256b_x4(taddr + u * kCols, vr);
          else if constexpr (M64) ptx::tmem_ld32(taddr + u * kCols, vr);
          else ptx::tmem_ld16(taddr + u * kCols, vr);
        };
        tmem_load(0);
#pragma unroll
        for (int r = 0; r < kMaxUpw; ++r) {
          if (r < upw) {
            if (kHasRes && kPref < kMaxUpw && r == kPref) prefetch_batch(kPref);  // second batch
            const int y = y_first + kUnitRows * r;  // first image row of the unit
            ptx::tmem_ld_wait();
            if (r < 2 && warp == 2 && lane == 0) MVLM_TRACE(8 + 4 * r);
            if (y < s.h) {
              if (F & F_HEAD) {
                // channel-major consumers: lane = channel c_lane, vr[j] = pixel (y + j/8, x0 + j%8)
                const int x0 = tc.tx * kTileW;
                const int nvx = s.w - x0;  // >= 8 for interior tiles
                if (cm_lane && c_lane < e.cout_real) {
#pragma unroll
                  for (int i = 0; i < kUnitRows; ++i) {
                    if (y + i < s.h) {
                      const int oy = (y + i) * e.up_sy + e.up_py;
                      const uint32_t idx0 = static_cast<uint32_t>(oy * ow + x0 * e.up_sx + e.up_px);
                      if (F & F_ARGMAX) {
                        // first maximum of the row's 8 pixels by a tournament over ordered keys (pixel index grows
                        // with j, so "the later one only if strictly greater" keeps the first), then ONE merge with
                        // the running best: a serial compare-select chain per pixel made this epilogue latency-bound
                        uint32_t k8[8];
                        uint32_t j8[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                          k8[j] = j < nvx ? order_f32(__uint_as_float(vr[8 * i + j]) + bias_c) : 0u;
                          j8[j] = static_cast<uint32_t>(j);
                        }
#pragma unroll
                        for (int st = 1; st < 8; st *= 2) {
#pragma unroll
                          for (int j = 0; j < 8; j += 2 * st) {
                            const bool take = k8[j + st] > k8[j];
                            k8[j] = take ? k8[j + st] : k8[j];
                            j8[j] = take ? j8[j + st] : j8[j];
                          }
                        }
                        const uint32_t lo = 0xFFFFFFFFu - (idx0 + j8[0] * static_cast<uint32_t>(e.up_sx));
                        if (k8[0] > best_hi || (k8[0] == best_hi && lo > best_lo)) {
                          best_hi = k8[0];
                          best_lo = lo;
                        }
                      }
                      if (F & F_F32) {
                        float* dst = e.out_f32 + (static_cast<size_t>(tc.img) * e.cout_real + c_lane) * oh * ow + idx0;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                          if (j < nvx) dst[j * e.up_sx] = __uint_as_float(vr[8 * i + j]) + bias_c;
                      }
                    }
                  }
                }
              }
              if (kBf16Out) {
                // bias in the channel-major role (one register), then transpose the unit through shared memory:
                // row = pixel, 36 (20)-float pitch (conflict-free STS.32; LDS.128 conflict-free at 36)
                __syncwarp();
                if constexpr (kFrag) {
                  // register 4k + 2h + e = (channel lane/4 + 8h, pixel 8k + 2(lane%4) + e); conflict-free at pitch 20
                  const int cb = m0 + cgrp * kChGrp + (lane >> 2);
                  const int p0 = 2 * (lane & 3);
#pragma unroll
                  for (int h = 0; h < 2; ++h) {
                    const int ch = (cb + 8 * h) < kMaxCout ? cb + 8 * h : 0;
                    const float bias_h = ep->bias[ch];
                    const float ms = (F & F_MID) ? ep->mid_s[ch] : 1.f, mt_ = (F & F_MID) ? ep->mid_t[ch] : 0.f;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
#pragma unroll
                      for (int e2 = 0; e2 < 2; ++e2) {
                        float f = __uint_as_float(vr[4 * k + 2 * h + e2]) + bias_h;
                        if (F & F_MID) f = fmaxf(fmaf(f, ms, mt_), 0.f);
                        stage[(8 * k + p0 + e2) * kPitch + (lane >> 2) + 8 * h] = f;
                      }
                    }
                  }
                } else if (cm_lane) {
#pragma unroll
                  for (int j = 0; j < kCols; ++j) {
                    float f = __uint_as_float(vr[j]) + bias_c;
                    if (F & F_MID) f = fmaxf(fmaf(f, mid_s_c, mid_t_c), 0.f);
                    stage[j * kPitch + lane] = f;
                  }
                }
              }
            }
            if (r + 1 < upw) tmem_load(r + 1);  // next unit in flight while this one is post-processed
            if (y < s.h) {
              if (kBf16Out) {
                __syncwarp();
                if (r < 2 && warp == 2 && lane == 0) MVLM_TRACE(9 + 4 * r);
                uint4 pool_cur[2];
#pragma unroll
                for (int ip = 0; ip < 2; ++ip) {
                  const int row = kUnitRows * r + kPassStep * ip;  // my image row in this pass, relative to my first
                  const bool valid = row < n_rows_ok;
                  float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                  if (valid) {
                    lds8(stage + (pj + 8 * (my_i + kPassStep * ip)) * kPitch + cq, f);
                    if (F & F_PRE)
                      *reinterpret_cast<uint4*>(b_pre + static_cast<size_t>(row * rs_pre)) = affine_relu_pack8(f, pre_s, pre_t);
                    if (F & F_RES1) add8(r1[r % kPref][ip], f);
                    if (F & F_RES2) add8(r2[r % kPref][ip], f);
                    if (F & F_UP) add8(ru[r % kPref][M64 ? ip : 0], f);
                  }
                  if (F & F_POOL) {
                    pool_cur[ip] = pack8(f);
                  } else if (valid) {
                    if (F & F_RAW) *reinterpret_cast<uint4*>(b_raw + static_cast<size_t>(row * rs_raw)) = pack8(f);
                    if (F & F_POST)
                      *reinterpret_cast<uint4*>(b_post + static_cast<size_t>(row * rs_post)) = affine_relu_pack8(f, post_s, post_t);
                  }
                }
                if (r < 2 && warp == 2 && lane == 0) MVLM_TRACE(10 + 4 * r);
                if (F & F_POOL) {
                  // 2x2 max-pool of the bf16-rounded values; shuffles run on all lanes.
                  //   M = 128: vertical partner = my other pass, horizontal (pixel column pj ^ 1) = lane ^ 4,
                  //            one pooled row per unit;
                  //   M = 64 : vertical partner = lane ^ 16 (same pass), horizontal = lane ^ 2, pass ip is pooled row
                  //            2r + ip of my part of the tile.
#pragma unroll
                  for (int pp = 0; pp < (M64 ? 2 : 1); ++pp) {
                    uint4 m = pool_cur[pp];
                    if (M64) {
                      uint4 o;
                      o.x = __shfl_xor_sync(0xffffffffu, m.x, 16);
                      o.y = __shfl_xor_sync(0xffffffffu, m.y, 16);
                      o.z = __shfl_xor_sync(0xffffffffu, m.z, 16);
                      o.w = __shfl_xor_sync(0xffffffffu, m.w, 16);
                      m = max_bf16x8(m, o);
                    } else {
                      m = max_bf16x8(m, pool_cur[1]);
                    }
                    uint4 o;
                    o.x = __shfl_xor_sync(0xffffffffu, m.x, M64 ? 2 : 4);
                    o.y = __shfl_xor_sync(0xffffffffu, m.y, M64 ? 2 : 4);
                    o.z = __shfl_xor_sync(0xffffffffu, m.z, M64 ? 2 : 4);
                    o.w = __shfl_xor_sync(0xffffffffu, m.w, M64 ? 2 : 4);
                    m = max_bf16x8(m, o);
                    // H, W even: row y+1 and column xa+1 exist whenever (y, xa) does
                    const int prow = (kUnitRows / 2) * r + pp;  // pooled row relative to y_first / 2
                    if (2 * prow < n_rows_ok && my_i == 0 && (pj & 1) == 0) {
                      if (F & F_RAW) *reinterpret_cast<uint4*>(b_raw + static_cast<size_t>(prow * rs_raw)) = m;
                      if (F & F_POST) {
                        float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        add8(m, g);
                        *reinterpret_cast<uint4*>(b_post + static_cast<size_t>(prow * rs_post)) = affine_relu_pack8(g, post_s, post_t);
                      }
                    }
                  }
                }
                if (r < 2 && warp == 2 && lane == 0) MVLM_TRACE(11 + 4 * r);
              }
            }
          }
        }
      }
      // all tcgen05.ld of this accumulator stage have completed (wait::ld above)
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar->t_empty[acc]);
      if (warp == 2 && lane == 0) MVLM_TRACE(6);
      ++trace_i;
      if (++acc == 2) { acc = 0; pacc ^= 1; }
    }
    if ((F & F_ARGMAX) && cur_img >= 0 && best_hi != 0u) {
      const int c_lane = cgrp * kChGrp + lane;  // arg-max convs have a single M tile
      if (cm_lane && c_lane < e.cout_real)
        atomicMax(e.argmax_keys + static_cast<size_t>(cur_img) * e.cout_real + c_lane,
                  (static_cast<unsigned long long>(best_hi) << 32) | best_lo);
    }
    if (prof && warp == 2 && lane == 0) { p.prof[blockIdx.x * 8 + 5] = w0; p.prof[blockIdx.x * 8 + 6] = clock64() - t_kernel0; }
  }

#undef MVLM_TRACE
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess || !sym)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

// programmatic dependent launch of consecutive conv kernels (MVLM_CONV_NO_PDL=1 switches it off)
const bool g_pdl = getenv("MVLM_CONV_NO_PDL") == nullptr;

template <int F>
int launch_t(const ConvParams& p, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    MVLM_CHECK_CUDA(cudaFuncSetAttribute(conv_umma_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MVLM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_umma_kernel<F>, p));
  count_launch();
  MVLM_CHECK_CUDA(cudaGetLastError());
  return MVLM_OK;
}

long long* g_prof_buf = nullptr;
int g_debug_mode = 0;

}  // namespace

int conv_plan(const ConvShape& s, const ConvEpilogue& e, ConvParams* out) {
  MVLM_REQUIRE(s.in && s.wpacked, "conv_plan: null input/weights");
  MVLM_REQUIRE(s.n > 0 && s.h > 0 && s.w > 0, "conv_plan: bad image dims %d %d %d", s.n, s.h, s.w);
  MVLM_REQUIRE(s.cin >= 16 && s.cin % 16 == 0, "conv_plan: cin=%d must be a multiple of 16", s.cin);
  MVLM_REQUIRE(s.in_cs >= s.cin && s.in_cs % 8 == 0, "conv_plan: in_cs=%d invalid", s.in_cs);
  MVLM_REQUIRE(s.cout_pad >= 16 && s.cout_pad % 16 == 0, "conv_plan: cout_pad=%d must be a multiple of 16", s.cout_pad);
  MVLM_REQUIRE(s.cout_pad <= kMTile || s.cout_pad % kMTile == 0,
               "conv_plan: cout_pad=%d must be <= 128 or a multiple of 128", s.cout_pad);
  MVLM_REQUIRE(s.cout_pad <= kMaxCout, "conv_plan: cout_pad=%d exceeds %d", s.cout_pad, kMaxCout);
  MVLM_REQUIRE(!e.argmax_keys || s.cout_pad <= kMTile, "conv_plan: fused arg-max needs cout_pad <= 128");
  MVLM_REQUIRE(!e.pool2 || (s.h % 2 == 0 && s.w % 2 == 0 && (e.out_raw || e.out_post) && !e.out_f32 && !e.argmax_keys),
               "conv_plan: pool2 needs even H, W and a bf16 raw/post output");
  MVLM_REQUIRE(!e.res_up || (s.h % 2 == 0 && s.w % 2 == 0 && !e.pool2 && !e.out_f32 && !e.argmax_keys),
               "conv_plan: res_up needs even H, W and bf16 outputs");
  MVLM_REQUIRE(!e.mid_scale || (e.mid_shift && !e.out_f32 && !e.argmax_keys), "conv_plan: mid affine needs mid_shift and bf16 outputs");
  MVLM_REQUIRE(!((e.argmax_keys || e.out_f32) && (e.res1 || e.res2 || e.out_pre || e.out_raw || e.out_post)),
               "conv_plan: fp32 / arg-max outputs cannot be combined with bf16 outputs or residual inputs");
  {
    const long long npix = 1ll * s.n * s.h * s.w;
    const int max_cs = std::max(std::max(std::max(e.pre_cs, e.raw_cs), std::max(std::max(e.post_cs, e.res1_cs), e.res2_cs)), e.up_cs);
    MVLM_REQUIRE(npix * std::max(max_cs, 1) < (1ll << 31), "conv_plan: tensor too large for 32-bit element offsets");
  }
  MVLM_REQUIRE(s.kh >= 1 && s.kh <= 3 && s.kw >= 1 && s.kw <= 3, "conv_plan: kernel %dx%d unsupported", s.kh, s.kw);
  MVLM_REQUIRE((reinterpret_cast<uintptr_t>(s.in) & 15) == 0 && (reinterpret_cast<uintptr_t>(s.wpacked) & 15) == 0,
               "conv_plan: pointers must be 16-byte aligned");
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("conv_plan: cuTensorMapEncodeTiled entry point unavailable");
    return MVLM_E_CUDA;
  }
  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.s = s;
  p.e = e;
  // output tile: 8 px x tile_h rows (N = 8 * tile_h columns), tile_h = 32 unless the map is lower
  const int tile_h = s.h >= 32 ? kMaxTileH : (s.h >= 16 ? 16 : (s.h >= 8 ? 8 : 4));
  p.tile_h = tile_h;
  {
    // ring depths: the halo slot holds the whole (8+KW-1) x (tile_h+KH-1) pixel tile of one 64-channel chunk
    const int hbytes = (kTileW + s.kw - 1) * (tile_h + s.kh - 1) * 128;
    p.h_slot_bytes = (hbytes + 1023) & ~1023;
    // cout <= 64 -> UMMA M = 64, 8 KB weight slots; all (chunk, tap) tiles resident when they fit the ring
    const int m_tile = s.cout_pad <= 64 ? 64 : kMTile;
    const int wslot = m_tile * 128;
    const int n_wtiles = ((s.cin + 63) / 64) * s.kh * s.kw;
    int nh = s.kh * s.kw == 1 ? 3 : 2, nw = s.kh * s.kw == 1 ? 6 : (m_tile == 64 ? kMaxWSlots : 7);
    p.w_stationary = (m_tile == 64 && n_wtiles <= kMaxWSlots && getenv("MVLM_CONV_NO_STATIONARY") == nullptr) ? 1 : 0;
    if (p.w_stationary) nw = n_wtiles;
    if (const char* env = getenv("MVLM_CONV_RING")) {  // experiment knob: "halo_slots,weight_slots"
      int a = 0, b = 0;
      if (sscanf(env, "%d,%d", &a, &b) == 2 && a >= 2 && a <= kMaxHSlots && b >= 2 && b <= kMaxWSlots && !p.w_stationary) { nh = a; nw = b; }
    }
    while (nh * p.h_slot_bytes + nw * wslot > kPoolBytes && nw > 2 && !p.w_stationary) --nw;
    while (nh * p.h_slot_bytes + nw * wslot > kPoolBytes && nh > 2) --nh;
    MVLM_REQUIRE(nh * p.h_slot_bytes + nw * wslot <= kPoolBytes, "conv_plan: operand rings do not fit");
    p.n_hslots = nh;
    p.n_wslots = nw;
  }
  {
    // X: (C, W, H, N) bf16, box (64, 8+KW-1, tile_h+KH-1, 1), 128-byte swizzle, OOB -> 0
    cuuint64_t gdim[4] = {(cuuint64_t)s.cin, (cuuint64_t)s.w, (cuuint64_t)s.h, (cuuint64_t)s.n};
    cuuint64_t gstr[3] = {(cuuint64_t)s.in_cs * 2, (cuuint64_t)s.in_cs * 2 * s.w,
                          (cuuint64_t)s.in_cs * 2 * s.w * s.h};
    cuuint32_t box[4] = {64, (cuuint32_t)(kTileW + s.kw - 1), (cuuint32_t)(tile_h + s.kh - 1), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&p.tm_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(s.in), gdim, gstr,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv_plan: cuTensorMapEncodeTiled(X) failed with %d (cin=%d w=%d h=%d n=%d cs=%d)", (int)r, s.cin,
                s.w, s.h, s.n, s.in_cs);
      return MVLM_E_CUDA;
    }
  }
  {
    // W: (K = KW*KH*cin, cout_pad) bf16, box (64, min(cout_pad,128))
    const cuuint64_t ktot = (cuuint64_t)s.kw * s.kh * s.cin;
    cuuint64_t gdim[2] = {ktot, (cuuint64_t)s.cout_pad};
    cuuint64_t gstr[1] = {ktot * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)(s.cout_pad < kMTile ? s.cout_pad : kMTile)};  // cout_pad <= 64: M = 64 tile
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&p.tm_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(s.wpacked), gdim,
                     gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv_plan: cuTensorMapEncodeTiled(W) failed with %d (ktot=%llu cout_pad=%d)", (int)r,
                (unsigned long long)ktot, s.cout_pad);
      return MVLM_E_CUDA;
    }
  }
  p.tail = (s.cin % 64 == 16 || s.cin % 64 == 32) ? s.cin % 64 : 0;
  if (p.tail) {
    const CUtensorMapSwizzle sw = p.tail == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B;
    cuuint64_t gdim[4] = {(cuuint64_t)s.cin, (cuuint64_t)s.w, (cuuint64_t)s.h, (cuuint64_t)s.n};
    cuuint64_t gstr[3] = {(cuuint64_t)s.in_cs * 2, (cuuint64_t)s.in_cs * 2 * s.w, (cuuint64_t)s.in_cs * 2 * s.w * s.h};
    cuuint32_t box[4] = {(cuuint32_t)p.tail, (cuuint32_t)(kTileW + s.kw - 1), (cuuint32_t)(tile_h + s.kh - 1), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&p.tm_a2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(s.in), gdim, gstr, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const cuuint64_t ktot = (cuuint64_t)s.kw * s.kh * s.cin;
    cuuint64_t gdim2[2] = {ktot, (cuuint64_t)s.cout_pad};
    cuuint64_t gstr2[1] = {ktot * 2};
    cuuint32_t box2[2] = {(cuuint32_t)p.tail, (cuuint32_t)(s.cout_pad < kMTile ? s.cout_pad : kMTile)};
    CUresult r2 = enc(&p.tm_b2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(s.wpacked), gdim2,
                      gstr2, box2, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
      set_error("conv_plan: cuTensorMapEncodeTiled(tail %d) failed with %d / %d", p.tail, (int)r, (int)r2);
      return MVLM_E_CUDA;
    }
  }
  p.tiles_x = ceil_div(s.w, kTileW);
  p.tiles_y = ceil_div(s.h, tile_h);
  p.n_nt = s.cout_pad <= 64 ? 1 : ceil_div(s.cout_pad, kMTile);
  p.total_tiles = s.n * p.tiles_x * p.tiles_y * p.n_nt;
  p.prof = nullptr;
  p.debug_mode = 0;
  *out = p;
  return MVLM_OK;
}

void conv_set_profile_buffer(long long* dev_buf) { g_prof_buf = dev_buf; }
void conv_set_debug_mode(int mode) { g_debug_mode = mode; }

int conv_launch(const ConvParams& p_in, cudaStream_t stream) {
  ConvParams p = p_in;
  p.prof = g_prof_buf;
  p.debug_mode = g_debug_mode;
  const ConvEpilogue& e = p.e;
  const int f = (e.out_pre ? F_PRE : 0) | (e.res1 ? F_RES1 : 0) | (e.res2 ? F_RES2 : 0) | (e.out_raw ? F_RAW : 0) |
                (e.out_post ? F_POST : 0) | (e.out_f32 ? F_F32 : 0) | (e.argmax_keys ? F_ARGMAX : 0) |
                (e.mid_scale ? F_MID : 0) | (e.pool2 ? F_POOL : 0) | (e.res_up ? F_UP : 0);
  // cout <= 64 layers run the M = 64 variant (conv_plan chose the ring geometry accordingly)
  const bool m64 = p.s.cout_pad <= 64;
#define MVLM_CASE_BOTH(FLAGS) \
  case (FLAGS): return m64 ? launch_t<(FLAGS) | F_M64>(p, stream) : launch_t<(FLAGS)>(p, stream)
#define MVLM_CASE_128(FLAGS) \
  case (FLAGS):              \
    if (!m64) return launch_t<(FLAGS)>(p, stream); \
    break
  switch (f) {
    // the combinations the network plan uses (hourglass.cu)
    MVLM_CASE_BOTH(F_PRE | F_RES1 | F_RAW | F_POST);  // RB conv1/2
    MVLM_CASE_BOTH(F_PRE | F_RES1 | F_RAW);
    MVLM_CASE_BOTH(F_RES1 | F_RAW | F_POST);          // RB conv3
    MVLM_CASE_BOTH(F_RES1 | F_RAW);
    MVLM_CASE_BOTH(F_RAW);                            // resample, conv6/10
    MVLM_CASE_128(F_PRE);                             // conv5, conv9
    MVLM_CASE_128(F_RES1 | F_RES2 | F_RAW | F_POST);  // conv7
    MVLM_CASE_128(F_ARGMAX);                          // conv11 phases
    MVLM_CASE_BOTH(F_F32);
    MVLM_CASE_BOTH(F_F32 | F_ARGMAX);
    MVLM_CASE_BOTH(F_MID | F_PRE | F_POST);           // stem (conv1)
    MVLM_CASE_BOTH(F_POOL | F_PRE | F_RES1 | F_RAW | F_POST);  // conv2 block: pooled outputs only
    MVLM_CASE_BOTH(F_POOL | F_RES1 | F_RAW | F_POST);
    MVLM_CASE_BOTH(F_POOL | F_RAW);
    MVLM_CASE_128(F_PRE | F_RES1 | F_RES2 | F_RAW | F_POST);
    // skip-branch ResidualBlocks with the up-sampled low path added in the epilogue (hourglass up path)
    MVLM_CASE_BOTH(F_UP | F_PRE | F_RES1 | F_RAW | F_POST);
    MVLM_CASE_BOTH(F_UP | F_PRE | F_RES1 | F_RAW);
    MVLM_CASE_BOTH(F_UP | F_RES1 | F_RAW | F_POST);
    MVLM_CASE_BOTH(F_UP | F_RES1 | F_RAW);
  }
#undef MVLM_CASE_BOTH
#undef MVLM_CASE_128
  set_error("conv_launch: unsupported epilogue combination 0x%x (cout_pad %d)", f, p.s.cout_pad);
  return MVLM_E_UNSUPPORTED;
}

}  // namespace mvlm
