// Multi-layer dataflow launch of the tcgen05 convolution pipeline, see conv_flow.cuh.
#include "conv_flow.cuh"

#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <set>

namespace mvlm {

namespace {

using namespace epi;

// fixed operand rings (the per-layer kernel sizes them per layer; here tiles of different layers follow each
// other through the same slots): 2 halo slots of (8+2) x (32+2) px x 128 B, 7 weight slots of 128 rows x 128 B
// warps 0..7 epilogue (two warpgroups), 8 dependency tracker + halo TMA, 9 MMA issuer, 10 weight TMA, 11 completion
// signaller (one warpgroup): setmaxnreg moves registers from the third warpgroup to the first two
constexpr int kFlowThreads = 384;
constexpr int kWarpTracker = 8, kWarpMma = 9, kWarpWeights = 10, kWarpSignal = 11;
constexpr int kEpiRegs = 224, kAuxRegs = 56;  // 2 x 128 x 224 + 128 x 56 = 64512 of the SM's 65536 registers
constexpr int kFlowTraceSkip = 256;          // debug timeline: items of CTA 0 skipped before tracing starts
constexpr int kHSlots = 2;
constexpr int kWSlots = 7;
constexpr int kHSlotBytes = 44032;
constexpr int kWSlotBytes = 16384;
constexpr int kPoolBytes = kHSlots * kHSlotBytes + kWSlots * kWSlotBytes;
static_assert((kTileW + 2) * (kMaxTileH + 2) * 128 <= kHSlotBytes && kHSlotBytes % 1024 == 0, "halo slot");

struct __align__(8) Barriers {
  uint64_t h_full[kHSlots];
  uint64_t h_empty[kHSlots];
  uint64_t w_full[kWSlots];
  uint64_t w_empty[kWSlots];
  uint64_t t_full[2];
  uint64_t t_empty[2];
  uint32_t tmem_base;
  int deps_ok;  // ordinal (1-based) of the last item of this CTA whose dependencies the producer has seen satisfied
  int epi_count[kEpiWarps];  // items of this CTA each epilogue warp has finished (all its stores issued)
};
static_assert(sizeof(Barriers) <= 512, "barrier block");

// per-layer parameters staged in shared memory at kernel start
struct LayerSm {
  ConvShape s;
  ConvEpilogue e;
  ChannelParams cp;
  int kind, f, tile_h, tail;
  int ring_in, ring_pre, ring_raw, ring_post, ring_res1, ring_res2, ring_up;
  int ring_aux;
};
static_assert(sizeof(LayerSm) % 4 == 0, "copied word by word");

constexpr int kSmemBytes = kPoolBytes + kEpiWarps * kStageFloats * 4 + 512 +
                           kFlowMaxLayers * static_cast<int>(sizeof(LayerSm)) + 1024 /*align*/;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_acquire_cta_shared(const int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(ptx::smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_cta_shared(int* p, int v) {
  asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(ptx::smem_u32(p)), "r"(v) : "memory");
}

struct Item {
  int layer, img, mt, tx, ty, group;
};
__device__ __forceinline__ Item decode_item(const int4& q) {
  Item it;
  it.layer = q.x & 0xffff;
  it.img = static_cast<int>(static_cast<unsigned>(q.x) >> 16);
  it.mt = q.y & 0xff;
  it.tx = (q.y >> 8) & 0xff;
  it.ty = static_cast<int>(static_cast<unsigned>(q.y) >> 16);
  it.group = q.z;
  return it;
}

// bounded spin: a scheduling bug must trap (kernel error) instead of hanging the GPU box
#define MVLM_FLOW_SPIN_GUARD(t0, what)                                                                     \
  do {                                                                                                     \
    if (clock64() - (t0) > 4000000000LL) {                                                                 \
      printf("mvlm: conv_flow %s timed out (block %d thread %d)\n", what, (int)blockIdx.x, (int)threadIdx.x); \
      __trap();                                                                                            \
    }                                                                                                      \
  } while (0)

// every epilogue variant the plan uses (the same list as conv_launch): X(flags) once per M = 128 / M = 64 form
#define MVLM_FLOW_VARIANTS_BOTH(X)            \
  X(F_PRE | F_RES1 | F_RAW | F_POST)          \
  X(F_PRE | F_RES1 | F_RAW)                   \
  X(F_RES1 | F_RAW | F_POST)                  \
  X(F_RES1 | F_RAW)                           \
  X(F_RAW)                                    \
  X(F_MID | F_PRE | F_POST)                   \
  X(F_POOL | F_PRE | F_RES1 | F_RAW | F_POST) \
  X(F_POOL | F_RES1 | F_RAW | F_POST)         \
  X(F_POOL | F_RAW)                           \
  X(F_UP | F_PRE | F_RES1 | F_RAW | F_POST)   \
  X(F_UP | F_PRE | F_RES1 | F_RAW)            \
  X(F_UP | F_RES1 | F_RAW | F_POST)           \
  X(F_UP | F_RES1 | F_RAW)                    \
  X(F_POST2 | F_PRE | F_RES1 | F_RAW | F_POST) \
  X(F_POST2 | F_RES1 | F_RAW | F_POST)        \
  X(F_POOLX | F_PRE | F_RES1 | F_RAW | F_POST) \
  X(F_POOLX | F_RES1 | F_RAW | F_POST)
#define MVLM_FLOW_VARIANTS_128(X)     \
  X(F_PRE)                            \
  X(F_RES1 | F_RES2 | F_RAW | F_POST) \
  X(F_PRE | F_RES1 | F_RES2 | F_RAW | F_POST) \
  X(F_POOLX | F_RES1 | F_RES2 | F_RAW | F_POST)

__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __bfloat1622float2(h[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}

// Element-wise items on the 256 epilogue threads: one 8 px x kFlowEltRows tile of the OUTPUT, all channels, as
// (pixel, 8 channels) 16-byte vectors; the arithmetic is that of pool2_act_kernel / bn_relu_kernel (eltwise.cu).
// (not inlined: inside the epilogue warps' loop its load buffers upset the register allocation of the conv variants,
// which then spill their residual prefetch buffers)
template <bool pool>
__device__ __noinline__ void elt_tile(const LayerSm& L, const Item& it, const int tid) {
  const int c8 = L.s.cin >> 3;
  const int hi = L.s.h, wi = L.s.w;
  const int ho = pool ? hi >> 1 : hi, wo = pool ? wi >> 1 : wi;
  const uint4* in = reinterpret_cast<const uint4*>(L.s.in);
  uint4* out_raw = reinterpret_cast<uint4*>(L.e.out_raw);
  uint4* out_act = reinterpret_cast<uint4*>(L.e.out_post);
  const int img_in = it.img % L.ring_in;
  const int img_raw = it.img % L.ring_raw, img_act = it.img % L.ring_post;
  const int n_vec = kTileW * kFlowEltRows * c8;
  // output vectors per thread in flight (16 / 8 loads): the loop is latency-bound (L2 round trips) otherwise
  constexpr int kU = pool ? 4 : 8;
  for (int v0 = tid; v0 < n_vec; v0 += kU * kEpiWarps * 32) {
    uint4 q[kU][pool ? 4 : 1];
    size_t o_idx[kU];
    bool ok[kU];
    int cvs[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int v = v0 + u * kEpiWarps * 32;
      const int cv = v % c8;
      const int px = v / c8;
      const int x = it.tx * kTileW + (px & 7);
      const int y = it.ty * kFlowEltRows + (px >> 3);
      ok[u] = v < n_vec && x < wo && y < ho;
      cvs[u] = cv;
      o_idx[u] = (static_cast<size_t>(y) * wo + x) * c8 + cv;
      if (!ok[u]) continue;
      if (pool) {
        const size_t base = ((static_cast<size_t>(img_in) * hi + 2 * y) * wi + 2 * x) * c8 + cv;
        q[u][0] = __ldcg(in + base);
        q[u][1] = __ldcg(in + base + c8);
        q[u][2] = __ldcg(in + base + static_cast<size_t>(wi) * c8);
        q[u][3] = __ldcg(in + base + static_cast<size_t>(wi) * c8 + c8);
      } else {
        q[u][0] = __ldcg(in + ((static_cast<size_t>(img_in) * hi + y) * wi + x) * c8 + cv);
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (!ok[u]) continue;
      float m[8];
      unpack8(q[u][0], m);
      if (pool) {
        float a[8], b[8];
        unpack8(q[u][1], a);
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], a[j]);
        unpack8(q[u][2], a);
        unpack8(q[u][3], b);
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], fmaxf(a[j], b[j]));
      }
      const size_t plane = static_cast<size_t>(ho) * wo * c8;
      if (out_raw) out_raw[img_raw * plane + o_idx[u]] = pack8(m);
      if (out_act) {
        float sc[8], sh[8], o[8];
        lds8(L.cp.post_s + cvs[u] * 8, sc);
        lds8(L.cp.post_t + cvs[u] * 8, sh);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaxf(fmaf(m[j], sc[j], sh[j]), 0.f);
        out_act[img_act * plane + o_idx[u]] = pack8(o);
      }
    }
  }
}

// One epilogue variant, inlined into the epilogue warps' switch.  (As separate non-inlined functions the variants
// kept their residual prefetch buffers on the stack -- every prefetched line was stored to local memory as soon as it
// arrived, one L2 round trip per load -- and inlined under the kernel-wide 168-register cap the widest variants
// spill; the epilogue warpgroups therefore raise their register allocation with setmaxnreg, see the kernel.)
template <int F>
__device__ __forceinline__ void flow_epilogue(const LayerSm* L, const int mt, const int tx, const int ty, const int img,
                                              uint64_t* t_full, const uint32_t parity, const uint32_t tmem_acc,
                                              float* stage, const int ew, const int lane, long long* wait_cycles,
                                              long long* trace_row, const long long trace_t0) {
  const TileCoord tc = {mt, tx, ty, img};
  // image -> slot of each tensor's buffer; the division only runs for ring buffers shorter than the image index
  auto slot = [img](int ring) __attribute__((always_inline)) { return img < ring ? img : img % ring; };
  const ImageSlots is = {slot(L->ring_pre), slot(L->ring_raw), slot(L->ring_post), slot(L->ring_res1),
                         slot(L->ring_res2), slot(L->ring_up), slot(L->ring_aux), slot(L->ring_aux)};
  ArgmaxState am;
  EpiTrace tr;
  tr.trace = trace_row; tr.trace_i = 0; tr.t0 = trace_t0; tr.on = trace_row != nullptr;
  long long w0 = 0;
  epilogue_tile<F, true>(L->s, L->e, L->tile_h, L->cp, tc, is, t_full, parity, tmem_acc, stage, ew, ew & 3, lane, am,
                         wait_cycles != nullptr, w0, tr);
  if (wait_cycles) *wait_cycles += w0;
}

// (launched with 168 registers per thread, the most 12 warps can have; re-balanced at run time)
__global__ void __launch_bounds__(kFlowThreads, 1)
conv_flow_kernel(const FlowLayer* __restrict__ layers, const void* __restrict__ layers_sm, const int n_layers,
                 const int4* __restrict__ items,
                 const int n_items, const FlowGroup* __restrict__ groups, unsigned int* __restrict__ done,
                 long long* __restrict__ prof_buf, const int debug_flags) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment: TMA and UMMA agree on the SWIZZLE_128B XOR pattern through the absolute address
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* h_slots = smem;
  uint8_t* w_slots = smem + kHSlots * kHSlotBytes;
  float* stage_all = reinterpret_cast<float*>(smem + kPoolBytes);
  Barriers* bar = reinterpret_cast<Barriers*>(reinterpret_cast<uint8_t*>(stage_all) + kEpiWarps * kStageFloats * 4);
  LayerSm* lsm = reinterpret_cast<LayerSm*>(reinterpret_cast<uint8_t*>(bar) + 512);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // stage the layers' parameters (written by the host before the launch, constant during it)
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(layers_sm);
    uint32_t* dst = reinterpret_cast<uint32_t*>(lsm);
    const int n_words = n_layers * static_cast<int>(sizeof(LayerSm) / 4);
    for (int i = threadIdx.x; i < n_words; i += kFlowThreads) dst[i] = __ldg(src + i);
  }
  if (warp == kWarpTracker && lane == 0) {
    for (int i = 0; i < kHSlots; ++i) {
      ptx::mbar_init(&bar->h_full[i], 1);
      ptx::mbar_init(&bar->h_empty[i], 1);
    }
    for (int i = 0; i < kWSlots; ++i) {
      ptx::mbar_init(&bar->w_full[i], 1);
      ptx::mbar_init(&bar->w_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bar->t_full[i], 1);
      ptx::mbar_init(&bar->t_empty[i], kEpiWarps);
    }
    bar->deps_ok = 0;
    for (int i = 0; i < kEpiWarps; ++i) bar->epi_count[i] = 0;
    ptx::fence_mbar_init();
  }
  if (warp == kWarpMma) {
    ptx::tmem_alloc(&bar->tmem_base, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  // programmatic dependent launch: everything above may overlap the tail of the previous kernel in the stream
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const uint32_t tmem_base = bar->tmem_base;
  const int step = static_cast<int>(gridDim.x);
  // optional role counters, 8 x int64 per CTA (cycles): [0] tracker: dependency waits, [1] tracker: halo slot waits,
  // [2] MMA: operand waits, [3] MMA: accumulator waits, [4] MMA: loop total, [5] epilogue warp 2: accumulator-ready
  // waits, [6] epilogue warp 2: dependency-flag waits, [7] epilogue warp 2: loop total
  const bool prof = prof_buf != nullptr;
  long long* const pc = prof ? prof_buf + blockIdx.x * 8 : nullptr;
  const long long t_kernel0 = clock64();
  long long w0 = 0, w1 = 0;

  // register re-allocation (all threads of a warpgroup execute it): the epilogue variants want up to ~200 registers,
  // the single-thread roles a few dozen
  if (warp >= kEpiWarps) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kAuxRegs));
  if (warp == kWarpTracker) {
    // ===================== dependency tracker + TMA producer: activations =====================
    if (ptx::elect_one()) {
      int sh = 0;
      uint32_t ph = 0;
      int g_ok = -1;
      int k = 0;
      int4 nxt = blockIdx.x < n_items ? __ldg(items + blockIdx.x) : make_int4(0, 0, 0, 0);
      for (int i = blockIdx.x; i < n_items; i += step) {
        const Item it = decode_item(nxt);
        if (i + step < n_items) nxt = __ldg(items + i + step);
        if (it.group != g_ok && !(debug_flags & 1)) {
          // all tiles of the groups this one depends on have been stored (and their stores made visible at GPU
          // scope before the count): producers of my inputs / residuals, earlier readers and writers of my outputs
          const FlowGroup& G = groups[it.group];
          const int nd = G.n_deps;
          for (int d = 0; d < nd; ++d) {
            const unsigned int* ctr = done + G.dep[d];
            const unsigned int need = static_cast<unsigned int>(G.dep_need[d]);
            if (ld_acquire_gpu(ctr) < need) {
              const long long t0 = clock64();
              while (ld_acquire_gpu(ctr) < need) MVLM_FLOW_SPIN_GUARD(t0, "dependency wait");
              w0 += clock64() - t0;
            }
          }
          g_ok = it.group;
          // the data was written through the generic proxy by other CTAs; the TMA reads below use the async proxy
          asm volatile("fence.proxy.async.global;" ::: "memory");
        }
        ++k;
        st_release_cta_shared(&bar->deps_ok, k);  // epilogue warps may now read this item's residual inputs
        const LayerSm& L = lsm[it.layer];
        if (L.kind != FLOW_CONV) continue;
        const ConvShape& s = L.s;
        const FlowLayer& GL = layers[it.layer];
        const int n_chunks = (s.cin + 63) >> 6;
        const int halo_px = kTileW + s.kw - 1, halo_rows = L.tile_h + s.kh - 1;
        const int x0 = it.tx * kTileW + s.x_off0;
        const int y0 = it.ty * L.tile_h + s.y_off0;
        const int img_in = it.img % L.ring_in;
        for (int c = 0; c < n_chunks; ++c) {
          // the last chunk may be a narrow tail (16 / 32 channels) with its own tensor map: rows of 32 / 64 bytes
          const bool is_tail = L.tail != 0 && c == n_chunks - 1;
          const uint32_t row_b = is_tail ? static_cast<uint32_t>(L.tail) * 2u : 128u;  // bytes per pixel
          timed_wait(&bar->h_empty[sh], ph ^ 1, prof, w1);
          ptx::mbar_expect_tx(&bar->h_full[sh], static_cast<uint32_t>(halo_rows * halo_px) * row_b);
          // one halo tile for all KW x KH taps of this chunk
          ptx::tma_load_4d(is_tail ? &GL.p.tm_a2 : &GL.p.tm_a, &bar->h_full[sh], h_slots + sh * kHSlotBytes, c * 64, x0, y0, img_in);
          if (++sh == kHSlots) { sh = 0; ph ^= 1; }
        }
      }
      if (prof) { pc[0] = w0; pc[1] = w1; }
    }
  } else if (warp == kWarpWeights) {
    // ===================== TMA producer: weights =====================
    if (ptx::elect_one()) {
      int sw = 0;
      uint32_t pw = 0;
      int4 nxt = blockIdx.x < n_items ? __ldg(items + blockIdx.x) : make_int4(0, 0, 0, 0);
      for (int i = blockIdx.x; i < n_items; i += step) {
        const Item it = decode_item(nxt);
        if (i + step < n_items) nxt = __ldg(items + i + step);
        const LayerSm& L = lsm[it.layer];
        if (L.kind != FLOW_CONV) continue;
        const ConvShape& s = L.s;
        const FlowLayer& GL = layers[it.layer];
        const int kM = (L.f & F_M64) ? 64 : kMTile;
        const int w_rows = s.cout_pad < kM ? s.cout_pad : kM;
        const int n_chunks = (s.cin + 63) >> 6;
        for (int c = 0; c < n_chunks; ++c) {
          const bool is_tail = L.tail != 0 && c == n_chunks - 1;
          const void* tmb = is_tail ? &GL.p.tm_b2 : &GL.p.tm_b;
          const uint32_t row_b = is_tail ? static_cast<uint32_t>(L.tail) * 2u : 128u;  // bytes per weight row
          const uint32_t wb = static_cast<uint32_t>(w_rows) * row_b;
          for (int tap = 0; tap < s.kw * s.kh; ++tap) {  // tap = kx * KH + ky
            ptx::mbar_wait(&bar->w_empty[sw], pw ^ 1);
            ptx::mbar_expect_tx(&bar->w_full[sw], wb);
            ptx::tma_load_2d(tmb, &bar->w_full[sw], w_slots + sw * kWSlotBytes, tap * s.cin + c * 64, it.mt * kM);
            if (++sw == kWSlots) { sw = 0; pw ^= 1; }
          }
        }
      }
    }
  } else if (warp == kWarpMma) {
    // ===================== MMA issuer =====================
    if (ptx::elect_one()) {
      int sh = 0, sw = 0;
      uint32_t ph = 0, pw = 0;
      int acc = 0;
      uint32_t pacc = 0;
      int4 nxt = blockIdx.x < n_items ? __ldg(items + blockIdx.x) : make_int4(0, 0, 0, 0);
      for (int i = blockIdx.x; i < n_items; i += step) {
        const Item it = decode_item(nxt);
        if (i + step < n_items) nxt = __ldg(items + i + step);
        const LayerSm& L = lsm[it.layer];
        if (L.kind != FLOW_CONV) continue;
        const ConvShape& s = L.s;
        // cout <= 64: tcgen05.mma.ws with M = 64 / 32 (see conv_umma.cu)
        const bool ws = (L.f & F_M64) != 0;
        const uint32_t idesc = ptx::umma_idesc_bf16(ws ? (s.cout_pad <= 32 ? 32 : 64) : kMTile, L.tile_h * kTileW);
        const int n_chunks = (s.cin + 63) >> 6;
        const int halo_px = kTileW + s.kw - 1;
        timed_wait(&bar->t_empty[acc], pacc ^ 1, prof, w1);
        ptx::tc_fence_after();
        const uint32_t d = tmem_base + static_cast<uint32_t>(acc * 256);
        uint32_t accumulate = 0;
        for (int c = 0; c < n_chunks; ++c) {
          const int rem = s.cin - c * 64;
          const int nk = rem >= 64 ? 4 : (rem >> 4);
          // tail chunk: SWIZZLE_32B (16 ch: 32-byte rows) or SWIZZLE_64B (32 ch: 64-byte rows)
          const bool is_tail = L.tail != 0 && c == n_chunks - 1;
          const uint32_t row_b = is_tail ? static_cast<uint32_t>(L.tail) * 2u : 128u;
          const uint32_t swz = !is_tail ? 2u : (L.tail == 16 ? 6u : 4u);
          const uint64_t a_hi = static_cast<uint64_t>(((8u * row_b) >> 4) | (1u << 14) | (swz << 29)) << 32;
          const uint64_t b_hi = static_cast<uint64_t>(((static_cast<uint32_t>(halo_px) * row_b) >> 4) | (1u << 14) | (swz << 29)) << 32;
          timed_wait(&bar->h_full[sh], ph, prof, w0);
          const uint32_t h_lo = ((ptx::smem_u32(h_slots + sh * kHSlotBytes) >> 4) & 0x3FFFu) | (1u << 16);
          for (int kx = 0; kx < s.kw; ++kx) {
            for (int ky = 0; ky < s.kh; ++ky) {
              timed_wait(&bar->w_full[sw], pw, prof, w0);
              ptx::tc_fence_after();
              const uint32_t w_lo = ((ptx::smem_u32(w_slots + sw * kWSlotBytes) >> 4) & 0x3FFFu) | (1u << 16);
              // tap (kx, ky) = the same halo tile shifted by ky halo rows + kx pixels
              const uint32_t x_lo = h_lo + ((static_cast<uint32_t>(ky * halo_px + kx) * row_b) >> 4);
              if (ws) {
                for (int kk = 0; kk < nk; ++kk)
                  ptx::umma_ws_bf16(d, a_hi | (w_lo + 2 * kk), b_hi | (x_lo + 2 * kk), idesc, (kk == 0) ? accumulate : 1u);
              } else if (nk == 4) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  ptx::umma_bf16(d, a_hi | (w_lo + 2 * kk), b_hi | (x_lo + 2 * kk), idesc, (kk == 0) ? accumulate : 1u);
              } else {
                for (int kk = 0; kk < nk; ++kk)
                  ptx::umma_bf16(d, a_hi | (w_lo + 2 * kk), b_hi | (x_lo + 2 * kk), idesc, (kk == 0) ? accumulate : 1u);
              }
              accumulate = 1;
              ptx::umma_commit(&bar->w_empty[sw]);
              if (++sw == kWSlots) { sw = 0; pw ^= 1; }
            }
          }
          ptx::umma_commit(&bar->h_empty[sh]);
          if (++sh == kHSlots) { sh = 0; ph ^= 1; }
        }
        ptx::umma_commit(&bar->t_full[acc]);
        if (++acc == 2) { acc = 0; pacc ^= 1; }
      }
      if (prof) { pc[2] = w0; pc[3] = w1; pc[4] = clock64() - t_kernel0; }
    }
  } else if (warp == kWarpSignal) {
    // ===================== completion signaller =====================
    // An epilogue warp only notes (CTA scope) that it has issued the stores of its part of an item.  This thread
    // waits until all eight have, makes those stores visible at GPU scope with ONE fence (cumulative over what it
    // observed through the CTA-scope acquire) and counts the tile on its group's counter.  A fence + atomic in every
    // epilogue warp (first version) stalls each of them for the write-acknowledge latency once per tile: the
    // segments ran 2.3x slower than the per-layer launches.
    if (ptx::elect_one()) {
      int k = 0;
      int4 nxt = blockIdx.x < n_items ? __ldg(items + blockIdx.x) : make_int4(0, 0, 0, 0);
      for (int i = blockIdx.x; i < n_items; i += step) {
        const int group = nxt.z;
        if (i + step < n_items) nxt = __ldg(items + i + step);
        ++k;
        const long long t0 = clock64();
        for (;;) {
          int lo = k;
#pragma unroll
          for (int w = 0; w < kEpiWarps; ++w) lo = min(lo, ld_acquire_cta_shared(&bar->epi_count[w]));
          if (lo >= k) break;
          __nanosleep(100);
          MVLM_FLOW_SPIN_GUARD(t0, "signaller wait");
        }
        if (debug_flags & 1) continue;  // timing experiment: no publication (and no dependency waits above)
        if (debug_flags & 2) {
          asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(done + group) : "memory");
        } else {
          __threadfence();
          atomicAdd(done + group, 1u);
        }
      }
    }
  }
  } else {
    // ===================== epilogue / element-wise items =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
    float* stage = stage_all + warp * kStageFloats;
    int acc = 0;
    uint32_t pacc = 0;
    int k = 0;
    int4 nxt = blockIdx.x < n_items ? __ldg(items + blockIdx.x) : make_int4(0, 0, 0, 0);
    for (int i = blockIdx.x; i < n_items; i += step) {
      const Item it = decode_item(nxt);
      if (i + step < n_items) nxt = __ldg(items + i + step);
      ++k;
      const long long t_item0 = prof ? clock64() : 0;
      // the tracker (warp 0) has seen this item's dependencies satisfied; normally it is tiles ahead
      if (ld_acquire_cta_shared(&bar->deps_ok) < k) {
        const long long t0 = clock64();
        while (ld_acquire_cta_shared(&bar->deps_ok) < k) MVLM_FLOW_SPIN_GUARD(t0, "epilogue dependency wait");
        w1 += clock64() - t0;
      }
      const LayerSm& L = lsm[it.layer];
      // timeline of CTA 0, warp 2 (items kFlowTraceSkip .. +kTraceTiles): [0] item fetched, [1] dependencies seen,
      // [2] layer, [3] item ordinal, [5] accumulator ready, [8..15] first two units (conv_umma.cu), [6] item done
      long long* trow = nullptr;
      if (prof && blockIdx.x == 0 && warp == 0 && lane == 0 && k > kFlowTraceSkip && k <= kFlowTraceSkip + kTraceTiles) {
        trow = prof_buf + kNumSMs * 8 + (k - 1 - kFlowTraceSkip) * 16;
        trow[0] = t_item0 - t_kernel0; trow[1] = clock64() - t_kernel0; trow[2] = it.layer; trow[3] = k;
      }
      if (L.kind == FLOW_CONV) {
        const uint32_t tmem_acc = tmem_base + static_cast<uint32_t>(acc * 256);
        switch (L.f) {
#define MVLM_FLOW_CASE(FLAGS)                                                                                      \
  case (FLAGS):                                                                                                    \
    flow_epilogue<(FLAGS)>(&L, it.mt, it.tx, it.ty, it.img, &bar->t_full[acc], pacc, tmem_acc, stage, warp, lane, \
                           prof ? &w0 : nullptr, trow, t_kernel0);                                                  \
    break;
#define MVLM_FLOW_CASE_M64(FLAGS) MVLM_FLOW_CASE((FLAGS) | F_M64)
          MVLM_FLOW_VARIANTS_BOTH(MVLM_FLOW_CASE)
          MVLM_FLOW_VARIANTS_BOTH(MVLM_FLOW_CASE_M64)
          MVLM_FLOW_VARIANTS_128(MVLM_FLOW_CASE)
#undef MVLM_FLOW_CASE
#undef MVLM_FLOW_CASE_M64
          default:
            // flow_build_segment rejects other masks; keep the pipeline moving if one slips through
            ptx::mbar_wait(&bar->t_full[acc], pacc);
            ptx::tc_fence_after();
            break;
        }
        // all tcgen05.ld of this accumulator stage have completed (wait::ld in the epilogue)
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar->t_empty[acc]);
        if (++acc == 2) { acc = 0; pacc ^= 1; }
      } else {
        if (L.kind == FLOW_POOL) elt_tile<true>(L, it, warp * 32 + lane);
        else elt_tile<false>(L, it, warp * 32 + lane);
        __syncwarp();
      }
      // this warp's part of the item is stored (the signaller publishes the item once all eight say so)
      __syncwarp();
      if (lane == 0) st_release_cta_shared(&bar->epi_count[warp], k);
      if (trow) trow[6] = clock64() - t_kernel0;
    }
    if (prof && warp == 0 && lane == 0) { pc[5] = w0; pc[6] = w1; pc[7] = clock64() - t_kernel0; }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

bool supported_mask(int f) {
  switch (f) {
#define MVLM_FLOW_CASE(FLAGS) case (FLAGS):
#define MVLM_FLOW_CASE_M64(FLAGS) case ((FLAGS) | F_M64):
    MVLM_FLOW_VARIANTS_BOTH(MVLM_FLOW_CASE)
    MVLM_FLOW_VARIANTS_BOTH(MVLM_FLOW_CASE_M64)
    MVLM_FLOW_VARIANTS_128(MVLM_FLOW_CASE)
#undef MVLM_FLOW_CASE
#undef MVLM_FLOW_CASE_M64
    return true;
  }
  return false;
}

}  // namespace

int flow_build_segment(const std::vector<FlowLayerDesc>& layers, int n_views, int batch, int interleave,
                       std::vector<void*>* owned, FlowSegment* out) {
  const int nl = static_cast<int>(layers.size());
  MVLM_REQUIRE(nl > 0 && nl <= kFlowMaxLayers, "flow: %d layers in one segment (max %d)", nl, kFlowMaxLayers);
  MVLM_REQUIRE(n_views > 0 && n_views < 65536 && batch > 0 && interleave > 0, "flow: bad batching %d/%d/%d", n_views,
               batch, interleave);
  // tensors read / written by every layer, identified by their base pointers
  std::vector<std::vector<const void*>> reads(nl), writes(nl);
  for (int l = 0; l < nl; ++l) {
    const FlowLayer& L = layers[l].layer;
    const ConvShape& s = L.p.s;
    const ConvEpilogue& e = L.p.e;
    MVLM_REQUIRE(layers[l].tiles_x > 0 && layers[l].tiles_x < 256 && layers[l].tiles_y > 0 && layers[l].n_nt > 0 &&
                     layers[l].n_nt < 256,
                 "flow: layer %d has a bad tile grid", l);
    if (L.kind == FLOW_CONV) {
      MVLM_REQUIRE(supported_mask(L.f), "flow: layer %d has an unsupported epilogue mask 0x%x", l, L.f);
      MVLM_REQUIRE(!e.out_f32 && !e.argmax_keys, "flow: head layers run through conv_launch");
    } else {
      MVLM_REQUIRE(s.in && s.cin % 8 == 0 && (e.out_raw || e.out_post), "flow: bad element-wise layer %d", l);
    }
    reads[l] = {s.in, e.res1, e.res2, e.res_up};
    writes[l] = {e.out_pre, e.out_raw, e.out_post, e.out_aux1, e.out_aux2};
  }
  auto touches = [](const std::vector<const void*>& a, const std::vector<const void*>& b) {
    for (const void* x : a)
      if (x)
        for (const void* y : b)
          if (x == y) return true;
    return false;
  };
  // layer-level dependencies (same batch): read-after-write, write-after-read, write-after-write; transitively
  // implied ones are dropped
  std::vector<std::set<int>> dep(nl), closure(nl);
  for (int l = 0; l < nl; ++l) {
    for (int q = l - 1; q >= 0; --q) {
      const bool need = touches(reads[l], writes[q]) || touches(writes[l], reads[q]) || touches(writes[l], writes[q]);
      if (need && !closure[l].count(q)) {
        dep[l].insert(q);
        closure[l].insert(q);
        closure[l].insert(closure[q].begin(), closure[q].end());
      }
    }
    MVLM_REQUIRE(static_cast<int>(dep[l].size()) <= kFlowMaxDeps, "flow: layer %d has %zu dependencies", l, dep[l].size());
  }
  const int n_batches = ceil_div(n_views, batch);
  std::vector<FlowItem> items;
  std::vector<FlowGroup> groups;
  std::vector<int> gid(static_cast<size_t>(nl) * n_batches, -1), gneed(static_cast<size_t>(nl) * n_batches, 0);
  for (int b0 = 0; b0 < n_batches; b0 += interleave) {
    for (int l = 0; l < nl; ++l) {
      const FlowLayerDesc& D = layers[l];
      for (int b = b0; b < std::min(n_batches, b0 + interleave); ++b) {
        const int img0 = b * batch, img1 = std::min(n_views, img0 + batch);
        FlowGroup G;
        memset(&G, 0, sizeof(G));
        for (int q : dep[l]) {
          G.dep[G.n_deps] = gid[static_cast<size_t>(q) * n_batches + b];
          G.dep_need[G.n_deps] = gneed[static_cast<size_t>(q) * n_batches + b];
          ++G.n_deps;
        }
        const int g = static_cast<int>(groups.size());
        int count = 0;
        for (int img = img0; img < img1; ++img)
          for (int ty = 0; ty < D.tiles_y; ++ty)
            for (int tx = 0; tx < D.tiles_x; ++tx)
              for (int mt = 0; mt < D.n_nt; ++mt) {
                FlowItem it;
                it.layer_img = l | (img << 16);
                it.tile = mt | (tx << 8) | (ty << 16);
                it.group = g;
                it.pad = 0;
                items.push_back(it);
                ++count;
              }
        gid[static_cast<size_t>(l) * n_batches + b] = g;
        gneed[static_cast<size_t>(l) * n_batches + b] = count;
        groups.push_back(G);
      }
    }
  }
  // every dependency points backwards in the list: with round-robin item assignment and all CTAs resident, the
  // smallest unfinished item can always run, so the launch cannot deadlock
  for (size_t g = 0; g < groups.size(); ++g)
    for (int d = 0; d < groups[g].n_deps; ++d)
      MVLM_REQUIRE(groups[g].dep[d] >= 0 && groups[g].dep[d] < static_cast<int>(g), "flow: forward dependency");
  FlowSegment seg;
  seg.n_layers = nl;
  seg.n_items = static_cast<int>(items.size());
  seg.n_groups = static_cast<int>(groups.size());
  std::vector<FlowLayer> dl(nl);
  std::vector<LayerSm> ds(nl);
  for (int l = 0; l < nl; ++l) {
    const FlowLayer& g = layers[l].layer;
    dl[l] = g;
    LayerSm& d = ds[l];
    memset(&d, 0, sizeof(d));
    d.s = g.p.s; d.e = g.p.e; d.cp = g.cp;
    d.kind = g.kind; d.f = g.f; d.tile_h = g.p.tile_h; d.tail = g.p.tail;
    d.ring_in = std::max(1, g.ring_in); d.ring_pre = std::max(1, g.ring_pre); d.ring_raw = std::max(1, g.ring_raw);
    d.ring_post = std::max(1, g.ring_post); d.ring_res1 = std::max(1, g.ring_res1);
    d.ring_res2 = std::max(1, g.ring_res2); d.ring_up = std::max(1, g.ring_up);
    d.ring_aux = std::max(1, g.ring_aux);
  }
  void* p = nullptr;
  MVLM_CHECK_CUDA(cudaMalloc(&p, sizeof(LayerSm) * nl));
  owned->push_back(p);
  seg.layers_sm = p;
  MVLM_CHECK_CUDA(cudaMemcpy(p, ds.data(), sizeof(LayerSm) * nl, cudaMemcpyHostToDevice));
  MVLM_CHECK_CUDA(cudaMalloc(&p, sizeof(FlowLayer) * nl));
  owned->push_back(p);
  seg.layers = static_cast<FlowLayer*>(p);
  MVLM_CHECK_CUDA(cudaMemcpy(p, dl.data(), sizeof(FlowLayer) * nl, cudaMemcpyHostToDevice));
  MVLM_CHECK_CUDA(cudaMalloc(&p, sizeof(FlowItem) * items.size()));
  owned->push_back(p);
  seg.items = static_cast<FlowItem*>(p);
  MVLM_CHECK_CUDA(cudaMemcpy(p, items.data(), sizeof(FlowItem) * items.size(), cudaMemcpyHostToDevice));
  MVLM_CHECK_CUDA(cudaMalloc(&p, sizeof(FlowGroup) * groups.size()));
  owned->push_back(p);
  seg.groups = static_cast<FlowGroup*>(p);
  MVLM_CHECK_CUDA(cudaMemcpy(p, groups.data(), sizeof(FlowGroup) * groups.size(), cudaMemcpyHostToDevice));
  MVLM_CHECK_CUDA(cudaMalloc(&p, sizeof(unsigned int) * groups.size()));
  owned->push_back(p);
  seg.done = static_cast<unsigned int*>(p);
  MVLM_CHECK_CUDA(cudaMemset(p, 0, sizeof(unsigned int) * groups.size()));
  *out = seg;
  return MVLM_OK;
}

namespace {
thread_local long long* g_flow_prof = nullptr;
// experiment switches (MVLM_FLOW_DEBUG): 1 = no dependency waits / no publication (timing only, results invalid),
// 2 = red.release instead of fence + atomicAdd
const int g_flow_debug = getenv("MVLM_FLOW_DEBUG") ? atoi(getenv("MVLM_FLOW_DEBUG")) : 0;
}
void flow_set_profile_buffer(long long* dev_buf) { g_flow_prof = dev_buf; }

int flow_launch(const FlowSegment& seg, cudaStream_t stream) {
  MVLM_REQUIRE(seg.layers && seg.items && seg.groups && seg.done && seg.n_items > 0, "flow_launch: empty segment");
  // the attribute is per device: one flag per device ordinal
  static std::atomic<bool> configured[kMaxDevices];
  int dev = 0;
  MVLM_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices || !configured[dev].load(std::memory_order_acquire)) {
    MVLM_CHECK_CUDA(cudaFuncSetAttribute(conv_flow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    if (dev >= 0 && dev < kMaxDevices) configured[dev].store(true, std::memory_order_release);
  }
  // the counters of the previous call are stale
  MVLM_CHECK_CUDA(cudaMemsetAsync(seg.done, 0, sizeof(unsigned int) * seg.n_groups, stream));
  // every CTA must be resident (items wait for items of other CTAs): one CTA per SM, never more CTAs than SMs
  const int n_sms = sm_count();
  const int grid = seg.n_items < n_sms ? seg.n_items : n_sms;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kFlowThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  MVLM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_flow_kernel, static_cast<const FlowLayer*>(seg.layers),
                                     static_cast<const void*>(seg.layers_sm), seg.n_layers,
                                     reinterpret_cast<const int4*>(seg.items), seg.n_items,
                                     static_cast<const FlowGroup*>(seg.groups), seg.done, g_flow_prof, g_flow_debug));
  count_launch();
  MVLM_CHECK_CUDA(cudaGetLastError());
  return MVLM_OK;
}

}  // namespace mvlm
