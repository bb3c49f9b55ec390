// tcgen05/TMEM implicit-GEMM convolution, see conv_umma.cuh for the contract.
//
// One persistent CTA per SM, 10 warps:
//   warp 0      TMA producer   (A: halo tile per (cin-chunk, kx); B: one
//                               weight tile per (cin-chunk, kx, ky))
//   warp 1      MMA issuer     (one elected lane issues tcgen05.mma; owns TMEM)
//   warps 2..9  epilogue       (tcgen05.ld -> bias/BN/ReLU/residual -> global)
//
// A CTA tile is 16x16 output pixels (two 128-row UMMA sub-tiles of 8 image
// rows each) x N_TILE output channels.  For a fixed input-channel chunk (64
// channels = one 128-byte swizzle row) and a fixed horizontal tap kx the
// producer loads ONE (16+KH-1)-row halo tile; the KH vertical taps are
// descriptors into that tile shifted by whole image rows (2048 B, so the
// 1024-byte swizzle-atom alignment is preserved).  A traffic is therefore
// KW*(16+KH-1)/16 tile loads per chunk instead of KW*KH.
// Zero padding comes from TMA out-of-bounds fill (negative / past-the-end
// coordinates), also for feature maps smaller than the 16x16 tile.
//
// Accumulators: 2 pipeline stages x 2 sub-tiles x N_TILE fp32 columns of TMEM,
// so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "conv_umma.cuh"

namespace mvlm {

namespace {

constexpr int kThreads = 320;
constexpr int kEpiWarps = 8;
constexpr int kTileW = 16;
constexpr int kTileH = 16;
constexpr int kASlots = 3;
constexpr int kARowBytes = kTileW * 128;               // one image row of the tile: 16 px x 64 ch bf16
constexpr int kASlotBytes = (kTileH + 2) * kARowBytes;  // up to 18 halo rows
constexpr int kSmemBudget = 227 * 1024;

template <int N_TILE>
struct Cfg {
  static constexpr int kBSlotBytes = N_TILE * 128;
  static constexpr int kBSlots = (N_TILE >= 128) ? 6 : 8;
  static constexpr int kSmemBytes = kASlots * kASlotBytes + kBSlots * kBSlotBytes + 1024 /*align*/ +
                                    256 /*barriers*/ + 5 * 256 * 4 /*EpiParams*/;
  static_assert(kSmemBytes <= kSmemBudget, "shared memory budget");
  static_assert(N_TILE % 16 == 0 && N_TILE >= 16 && N_TILE <= 128, "UMMA N");
};

struct __align__(8) Barriers {
  uint64_t a_full[kASlots];
  uint64_t a_empty[kASlots];
  uint64_t b_full[8];
  uint64_t b_empty[8];
  uint64_t t_full[2];
  uint64_t t_empty[2];
  uint32_t tmem_base;
};
static_assert(sizeof(Barriers) <= 256, "barrier block");

__device__ __forceinline__ uint32_t order_f32(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ uint4 pack16_lo(const float (&f)[16]) {
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
__device__ __forceinline__ uint4 pack16_hi(const float (&f)[16]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(f[8], f[9]);
  __nv_bfloat162 b = __floats2bfloat162_rn(f[10], f[11]);
  __nv_bfloat162 c = __floats2bfloat162_rn(f[12], f[13]);
  __nv_bfloat162 d = __floats2bfloat162_rn(f[14], f[15]);
  uint4 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  r.z = *reinterpret_cast<uint32_t*>(&c);
  r.w = *reinterpret_cast<uint32_t*>(&d);
  return r;
}
__device__ __forceinline__ void store16(__nv_bfloat16* dst, const float (&f)[16]) {
  uint4* p = reinterpret_cast<uint4*>(dst);
  p[0] = pack16_lo(f);
  p[1] = pack16_hi(f);
}
__device__ __forceinline__ void add16(const __nv_bfloat16* src, float (&f)[16]) {
  const uint4* p = reinterpret_cast<const uint4*>(src);
  uint4 q[2] = {p[0], p[1]};
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(q);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float2 t = __bfloat1622float2(h[j]);
    f[2 * j] += t.x;
    f[2 * j + 1] += t.y;
  }
}

// Per-output-channel epilogue parameters staged once per CTA in shared memory (LDS broadcast instead of
// 80 global loads per 16-channel step): bias, pre scale/shift, post scale/shift for all N tiles.
constexpr int kMaxCout = 256;
struct EpiParams {
  float bias[kMaxCout];
  float pre_s[kMaxCout];
  float pre_t[kMaxCout];
  float post_s[kMaxCout];
  float post_t[kMaxCout];
};

__device__ __forceinline__ void lds16(const float* src, float (&v)[16]) {
  const float4* p = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 q = p[j];
    v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
  }
}

// role timing: wait on an mbarrier and add the stalled cycles to `acc` when profiling is on
__device__ __forceinline__ void timed_wait(uint64_t* b, uint32_t parity, bool prof, long long& acc) {
  if (!prof) { ptx::mbar_wait(b, parity); return; }
  const long long t0 = clock64();
  ptx::mbar_wait(b, parity);
  acc += clock64() - t0;
}

template <int N_TILE, bool ARGMAX>
__global__ void __launch_bounds__(kThreads, 1) conv_umma_kernel(const __grid_constant__ ConvParams p) {
  using C = Cfg<N_TILE>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by SWIZZLE_128B atoms (TMA and UMMA agree on
  // the XOR pattern only relative to 1024-byte aligned addresses).
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* a_slots = smem;
  uint8_t* b_slots = smem + kASlots * kASlotBytes;
  Barriers* bar = reinterpret_cast<Barriers*>(b_slots + C::kBSlots * C::kBSlotBytes);
  EpiParams* ep = reinterpret_cast<EpiParams*>(reinterpret_cast<uint8_t*>(bar) + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const ConvShape& s = p.s;
  const ConvEpilogue& e = p.e;
  const int n_chunks = (s.cin + 63) >> 6;
  const int halo_rows = kTileH + s.kh - 1;
  const uint32_t a_bytes = static_cast<uint32_t>(halo_rows) * kARowBytes;

  // Tile schedule shared by the three roles.  Default: round-robin (neighbouring CTAs work on
  // neighbouring tiles -> halo / weight reuse in L2).  ARGMAX: contiguous ranges, so that a CTA stays
  // within one image for ~40 tiles and keeps its running arg-max in registers.
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
    ep->pre_s[i] = e.out_pre ? e.pre_scale[i] : 0.f;
    ep->pre_t[i] = e.out_pre ? e.pre_shift[i] : 0.f;
    ep->post_s[i] = e.out_post ? e.post_scale[i] : 0.f;
    ep->post_t[i] = e.out_post ? e.post_shift[i] : 0.f;
  }
  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&p.tm_a);
    ptx::tma_prefetch_desc(&p.tm_b);
    for (int i = 0; i < kASlots; ++i) {
      ptx::mbar_init(&bar->a_full[i], 1);
      ptx::mbar_init(&bar->a_empty[i], 1);
    }
    for (int i = 0; i < C::kBSlots; ++i) {
      ptx::mbar_init(&bar->b_full[i], 1);
      ptx::mbar_init(&bar->b_empty[i], 1);
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
  const uint32_t tmem_base = bar->tmem_base;
  const bool prof = p.prof != nullptr;
  const long long t_kernel0 = clock64();
  long long w0 = 0, w1 = 0;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (ptx::elect_one() && p.debug_mode != 1) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int t = t_begin; t < t_end; t += t_step) {
        const int nt = t % p.n_nt;
        int r = t / p.n_nt;
        const int tx = r % p.tiles_x;
        r /= p.tiles_x;
        const int ty = r % p.tiles_y;
        const int img = r / p.tiles_y;
        const int x0 = tx * kTileW + s.x_off0;
        const int y0 = ty * kTileH + s.y_off0;
        for (int c = 0; c < n_chunks; ++c) {
          for (int kx = 0; kx < s.kw; ++kx) {
            timed_wait(&bar->a_empty[sa], pa ^ 1, prof, w0);
            ptx::mbar_expect_tx(&bar->a_full[sa], a_bytes);
            ptx::tma_load_4d(&p.tm_a, &bar->a_full[sa], a_slots + sa * kASlotBytes, c * 64, x0 + kx,
                             y0, img);
            if (++sa == kASlots) { sa = 0; pa ^= 1; }
            for (int ky = 0; ky < s.kh; ++ky) {
              timed_wait(&bar->b_empty[sb], pb ^ 1, prof, w1);
              ptx::mbar_expect_tx(&bar->b_full[sb], C::kBSlotBytes);
              ptx::tma_load_2d(&p.tm_b, &bar->b_full[sb], b_slots + sb * C::kBSlotBytes,
                               (kx * s.kh + ky) * s.cin + c * 64, nt * N_TILE);
              if (++sb == C::kBSlots) { sb = 0; pb ^= 1; }
            }
          }
        }
      }
      if (prof) { p.prof[blockIdx.x * 8 + 0] = w0; p.prof[blockIdx.x * 8 + 1] = w1; p.prof[blockIdx.x * 8 + 7] = clock64() - t_kernel0; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // elect.sync (not `lane == 0`): the compiler then knows a single lane is active and emits the
    // UTCHMMA / UTCBAR uniform-datapath instructions without a per-lane serialisation loop.
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, N_TILE);
      // descriptor = {lo: start>>4 | LBO, hi: SBO | version | SWIZZLE_128B}; only lo changes per MMA
      constexpr uint32_t kDescHi = 64u | (1u << 14) | (2u << 29);
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int acc = 0;
      uint32_t pacc = 0;
      for (int t = t_begin; t < t_end; t += t_step) {
        const int ty = (t / p.n_nt / p.tiles_x) % p.tiles_y;
        // maps of height <= 8 (hourglass levels 8^2 .. 1^2) only populate the first 128-row sub-tile
        const int n_sub = (ty * kTileH + 8 < s.h) ? 2 : 1;
        timed_wait(&bar->t_empty[acc], pacc ^ 1, prof, w1);
        ptx::tc_fence_after();
        const uint32_t d_base = tmem_base + static_cast<uint32_t>(acc * 2 * N_TILE);
        uint32_t accumulate = 0;
        for (int c = 0; c < n_chunks; ++c) {
          const int rem = s.cin - c * 64;
          const int nk = rem >= 64 ? 4 : (rem >> 4);
          for (int kx = 0; kx < s.kw; ++kx) {
            if (p.debug_mode != 1) timed_wait(&bar->a_full[sa], pa, prof, w0);
            const uint32_t a_lo = ((ptx::smem_u32(a_slots + sa * kASlotBytes) >> 4) & 0x3FFFu) | (1u << 16);
            for (int ky = 0; ky < s.kh; ++ky) {
              if (p.debug_mode != 1) timed_wait(&bar->b_full[sb], pb, prof, w0);
              ptx::tc_fence_after();
              const uint32_t b_lo = ((ptx::smem_u32(b_slots + sb * C::kBSlotBytes) >> 4) & 0x3FFFu) | (1u << 16);
              for (int sub = 0; sub < n_sub; ++sub) {
                // vertical tap / sub-tile = whole image rows of the halo tile: (sub*8+ky) * 2048 B >> 4
                const uint32_t a_sub = a_lo + static_cast<uint32_t>((sub * 8 + ky) * (kARowBytes >> 4));
                const uint32_t d = d_base + sub * N_TILE;
                if (nk == 4) {
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    ptx::umma_bf16(d, (static_cast<uint64_t>(kDescHi) << 32) | (a_sub + 2 * k),
                                   (static_cast<uint64_t>(kDescHi) << 32) | (b_lo + 2 * k), idesc,
                                   (k == 0) ? accumulate : 1u);
                } else {
                  for (int k = 0; k < nk; ++k)
                    ptx::umma_bf16(d, (static_cast<uint64_t>(kDescHi) << 32) | (a_sub + 2 * k),
                                   (static_cast<uint64_t>(kDescHi) << 32) | (b_lo + 2 * k), idesc,
                                   (k == 0) ? accumulate : 1u);
                }
              }
              accumulate = 1;
              if (p.debug_mode != 1) ptx::umma_commit(&bar->b_empty[sb]);
              if (++sb == C::kBSlots) { sb = 0; pb ^= 1; }
            }
            if (p.debug_mode != 1) ptx::umma_commit(&bar->a_empty[sa]);
            if (++sa == kASlots) { sa = 0; pa ^= 1; }
          }
        }
        ptx::umma_commit(&bar->t_full[acc]);
        if (++acc == 2) { acc = 0; pacc ^= 1; }
      }
      if (prof) { p.prof[blockIdx.x * 8 + 2] = w0; p.prof[blockIdx.x * 8 + 3] = w1; p.prof[blockIdx.x * 8 + 4] = clock64() - t_kernel0; }
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 2;
    const int lane_grp = warp & 3;  // TMEM lanes this warp may read: 32*(warp%4)..
    const int half = ew >> 2;
    constexpr int kChunks = N_TILE / 16;
    constexpr int kHalf0 = (kChunks + 1) / 2;
    const int ch_begin = half == 0 ? 0 : kHalf0;
    const int ch_end = half == 0 ? kHalf0 : kChunks;
    const int m = lane_grp * 32 + lane;
    int acc = 0;
    uint32_t pacc = 0;
    // ARGMAX: running (ordered value, ~index) per owned channel, kept across the tiles of one image
    constexpr int kKeys = ARGMAX ? kHalf0 * 16 : 1;
    uint32_t best_hi[kKeys], best_lo[kKeys];
    int cur_img = -1;
#pragma unroll
    for (int i = 0; i < kKeys; ++i) { best_hi[i] = 0u; best_lo[i] = 0u; }
    auto flush = [&](int img) {
      if (!ARGMAX || img < 0) return;
#pragma unroll
      for (int ci = 0; ci < kHalf0; ++ci) {
        const int ch = ch_begin + ci;
        if (ch >= ch_end) continue;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t hi = best_hi[ARGMAX ? ci * 16 + j : 0], lo = best_lo[ARGMAX ? ci * 16 + j : 0];
          const uint32_t mhi = __reduce_max_sync(0xffffffffu, hi);
          const uint32_t mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
          const int cidx = ch * 16 + j;  // N tile 0 only: arg-max convs have a single N tile
          if (lane == 0 && cidx < e.cout_real && mhi != 0u)
            atomicMax(e.argmax_keys + static_cast<size_t>(img) * e.cout_real + cidx,
                      (static_cast<unsigned long long>(mhi) << 32) | mlo);
          best_hi[ARGMAX ? ci * 16 + j : 0] = 0u;
          best_lo[ARGMAX ? ci * 16 + j : 0] = 0u;
        }
      }
    };
    for (int t = t_begin; t < t_end; t += t_step) {
      const int nt = t % p.n_nt;
      int r = t / p.n_nt;
      const int tx = r % p.tiles_x;
      r /= p.tiles_x;
      const int ty = r % p.tiles_y;
      const int img = r / p.tiles_y;
      const int n_sub = (ty * kTileH + 8 < s.h) ? 2 : 1;
      if (ARGMAX && img != cur_img) {
        flush(cur_img);
        cur_img = img;
      }
      timed_wait(&bar->t_full[acc], pacc, prof, w0);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < n_sub; ++sub) {
        const int y = ty * kTileH + sub * 8 + (m >> 4);
        const int x = tx * kTileW + (m & 15);
        const bool valid = (y < s.h) && (x < s.w);
        const size_t pix = (static_cast<size_t>(img) * s.h + y) * s.w + x;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) +
                               static_cast<uint32_t>(acc * 2 * N_TILE + sub * N_TILE);
        const int oy = y * e.up_sy + e.up_py;
        const int ox = x * e.up_sx + e.up_px;
        const int oh = s.h * e.up_sy, ow = s.w * e.up_sx;
#pragma unroll
        for (int ci = 0; ci < kHalf0; ++ci) {
          const int ch = ch_begin + ci;
          if (ch >= ch_end) continue;
          const int c0 = nt * N_TILE + ch * 16;
          // residual loads first: their latency overlaps the TMEM load
          uint4 r1[2], r2[2];
          const bool has_r1 = e.res1 && valid, has_r2 = e.res2 && valid;
          if (has_r1) {
            const uint4* q = reinterpret_cast<const uint4*>(e.res1 + pix * e.res1_cs + e.res1_co + c0);
            r1[0] = q[0]; r1[1] = q[1];
          }
          if (has_r2) {
            const uint4* q = reinterpret_cast<const uint4*>(e.res2 + pix * e.res2_cs + e.res2_co + c0);
            r2[0] = q[0]; r2[1] = q[1];
          }
          uint32_t v[16];
          ptx::tmem_ld16(taddr + ch * 16, v);
          float f[16], pb_[16];
          lds16(ep->bias + c0, pb_);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + pb_[j];
          if (e.out_pre && valid) {
            float sc[16], sh[16], g[16];
            lds16(ep->pre_s + c0, sc);
            lds16(ep->pre_t + c0, sh);
#pragma unroll
            for (int j = 0; j < 16; ++j) g[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
            store16(e.out_pre + pix * e.pre_cs + e.pre_co + c0, g);
          }
          if (has_r1) {
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(r1);
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float2 q = __bfloat1622float2(h[j]); f[2 * j] += q.x; f[2 * j + 1] += q.y; }
          }
          if (has_r2) {
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(r2);
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float2 q = __bfloat1622float2(h[j]); f[2 * j] += q.x; f[2 * j + 1] += q.y; }
          }
          if (e.out_raw && valid) store16(e.out_raw + pix * e.raw_cs + e.raw_co + c0, f);
          if (e.out_post && valid) {
            float sc[16], sh[16], g[16];
            lds16(ep->post_s + c0, sc);
            lds16(ep->post_t + c0, sh);
#pragma unroll
            for (int j = 0; j < 16; ++j) g[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
            store16(e.out_post + pix * e.post_cs + e.post_co + c0, g);
          }
          if (e.out_f32 && valid) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int c = c0 + j;
              if (c < e.cout_real)
                e.out_f32[((static_cast<size_t>(img) * e.cout_real + c) * oh + oy) * ow + ox] = f[j];
            }
          }
          if (ARGMAX && valid) {
            const uint32_t inv_idx = 0xFFFFFFFFu - static_cast<uint32_t>(oy * ow + ox);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const uint32_t hi = order_f32(f[j]);
              uint32_t& bh = best_hi[ARGMAX ? ci * 16 + j : 0];
              uint32_t& bl = best_lo[ARGMAX ? ci * 16 + j : 0];
              if (hi > bh || (hi == bh && inv_idx > bl)) { bh = hi; bl = inv_idx; }
            }
          }
        }
      }
      // all tcgen05.ld of this accumulator stage have completed (wait::ld above)
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar->t_empty[acc]);
      if (++acc == 2) { acc = 0; pacc ^= 1; }
    }
    flush(cur_img);
    if (prof && warp == 2 && lane == 0) { p.prof[blockIdx.x * 8 + 5] = w0; p.prof[blockIdx.x * 8 + 6] = clock64() - t_kernel0; }
  }

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

template <int N_TILE, bool ARGMAX>
int launch_t(const ConvParams& p, cudaStream_t stream) {
  using C = Cfg<N_TILE>;
  static bool configured = false;
  if (!configured) {
    MVLM_CHECK_CUDA(cudaFuncSetAttribute(conv_umma_kernel<N_TILE, ARGMAX>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    configured = true;
  }
  const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
  conv_umma_kernel<N_TILE, ARGMAX><<<grid, kThreads, C::kSmemBytes, stream>>>(p);
  count_launch();
  MVLM_CHECK_CUDA(cudaGetLastError());
  return MVLM_OK;
}

}  // namespace

int conv_plan(const ConvShape& s, const ConvEpilogue& e, ConvParams* out) {
  MVLM_REQUIRE(s.in && s.wpacked, "conv_plan: null input/weights");
  MVLM_REQUIRE(s.n > 0 && s.h > 0 && s.w > 0, "conv_plan: bad image dims %d %d %d", s.n, s.h, s.w);
  MVLM_REQUIRE(s.cin >= 16 && s.cin % 16 == 0, "conv_plan: cin=%d must be a multiple of 16", s.cin);
  MVLM_REQUIRE(s.in_cs >= s.cin && s.in_cs % 8 == 0, "conv_plan: in_cs=%d invalid", s.in_cs);
  MVLM_REQUIRE(s.n_tile == 32 || s.n_tile == 64 || s.n_tile == 80 || s.n_tile == 96 || s.n_tile == 128,
               "conv_plan: n_tile=%d unsupported", s.n_tile);
  MVLM_REQUIRE(s.cout_pad > 0 && s.cout_pad % s.n_tile == 0, "conv_plan: cout_pad=%d not a multiple of n_tile=%d",
               s.cout_pad, s.n_tile);
  MVLM_REQUIRE(s.cout_pad <= kMaxCout, "conv_plan: cout_pad=%d exceeds %d", s.cout_pad, kMaxCout);
  MVLM_REQUIRE(!e.argmax_keys || s.cout_pad == s.n_tile, "conv_plan: fused arg-max needs a single N tile");
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
  {
    // A: (C, W, H, N) bf16, box (64, 16, 16+KH-1, 1), 128-byte swizzle, OOB -> 0
    cuuint64_t gdim[4] = {(cuuint64_t)s.cin, (cuuint64_t)s.w, (cuuint64_t)s.h, (cuuint64_t)s.n};
    cuuint64_t gstr[3] = {(cuuint64_t)s.in_cs * 2, (cuuint64_t)s.in_cs * 2 * s.w,
                          (cuuint64_t)s.in_cs * 2 * s.w * s.h};
    cuuint32_t box[4] = {64, (cuuint32_t)kTileW, (cuuint32_t)(kTileH + s.kh - 1), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&p.tm_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(s.in), gdim, gstr,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv_plan: cuTensorMapEncodeTiled(A) failed with %d (cin=%d w=%d h=%d n=%d cs=%d)", (int)r, s.cin,
                s.w, s.h, s.n, s.in_cs);
      return MVLM_E_CUDA;
    }
  }
  {
    // B: (K = KW*KH*cin, cout_pad) bf16, box (64, N_TILE)
    const cuuint64_t ktot = (cuuint64_t)s.kw * s.kh * s.cin;
    cuuint64_t gdim[2] = {ktot, (cuuint64_t)s.cout_pad};
    cuuint64_t gstr[1] = {ktot * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)s.n_tile};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&p.tm_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(s.wpacked), gdim,
                     gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv_plan: cuTensorMapEncodeTiled(B) failed with %d (ktot=%llu cout_pad=%d)", (int)r,
                (unsigned long long)ktot, s.cout_pad);
      return MVLM_E_CUDA;
    }
  }
  p.tiles_x = ceil_div(s.w, kTileW);
  p.tiles_y = ceil_div(s.h, kTileH);
  p.n_nt = s.cout_pad / s.n_tile;
  p.total_tiles = s.n * p.tiles_x * p.tiles_y * p.n_nt;
  p.prof = nullptr;
  p.debug_mode = 0;
  *out = p;
  return MVLM_OK;
}

static long long* g_prof_buf = nullptr;
void conv_set_profile_buffer(long long* dev_buf) { g_prof_buf = dev_buf; }
static int g_debug_mode = 0;
void conv_set_debug_mode(int mode) { g_debug_mode = mode; }

int conv_launch(const ConvParams& p_in, cudaStream_t stream) {
  ConvParams p = p_in;
  p.prof = g_prof_buf;
  p.debug_mode = g_debug_mode;
  const bool am = p.e.argmax_keys != nullptr;
  switch (p.s.n_tile) {
    case 32: return am ? launch_t<32, true>(p, stream) : launch_t<32, false>(p, stream);
    case 64: return am ? launch_t<64, true>(p, stream) : launch_t<64, false>(p, stream);
    case 80: return am ? launch_t<80, true>(p, stream) : launch_t<80, false>(p, stream);
    case 96: return am ? launch_t<96, true>(p, stream) : launch_t<96, false>(p, stream);
    case 128: return am ? launch_t<128, true>(p, stream) : launch_t<128, false>(p, stream);
  }
  set_error("conv_launch: n_tile=%d unsupported", p.s.n_tile);
  return MVLM_E_UNSUPPORTED;
}

}  // namespace mvlm
