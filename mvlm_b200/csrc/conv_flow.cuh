// Dataflow execution of several consecutive layers of the CNN plan in ONE persistent launch.
//
// The per-layer kernel (conv_umma.cu) runs a layer over all V views before the next layer starts, so at
// V = 100 every activation tensor (210 MB .. 1.7 GB) streams through HBM between its producer and its
// consumer.  Here a launch covers a SEGMENT of the plan (e.g. the stem and the three 256^2 / 128^2
// ResidualBlocks, reference paulsenpredictor.py:405-413) as a list of work items
//     item = (layer, output tile)            grouped as  group = (layer, batch of B views),
// ordered batch-major with K batches interleaved layer by layer.  Every CTA takes items round-robin
// from the list; a group's items may start once the groups it depends on (its producers for the same
// views, plus write-after-read / write-after-write predecessors of the buffers it writes) have counted
// all their tiles in `done[group]`.  A tensor written for batch b is consumed within a few hundred items,
// i.e. while it is still in the 126 MB L2, and -- with ring buffers of a few batches instead of V views per
// intermediate tensor -- it is overwritten there before it is ever written back.
//
// The tile pipeline (TMA halo / weight producers, tcgen05 MMA issuer, eight epilogue warps) is the one of
// conv_umma.cu with a fixed operand-ring geometry; the epilogue is the shared epi::epilogue_tile.  Two
// element-wise item kinds (2x2 max-pool + BatchNorm/ReLU copy, BatchNorm/ReLU copy) run on the epilogue
// warps alone, so that no separate launch separates them from their producers.
#pragma once
#include <vector>

#include "conv_epilogue.cuh"

namespace mvlm {

constexpr int kFlowMaxLayers = 18;  // layers per segment (their parameters are staged in shared memory)
constexpr int kFlowMaxDeps = 10;
constexpr int kFlowEltRows = 8;  // element-wise items cover 8 px x 8 rows of their output (short items: a long one
                                 // delays every group that waits for its group)

enum FlowKind : int { FLOW_CONV = 0, FLOW_POOL = 1, FLOW_BNRELU = 2 };

// One layer of a segment (device-resident array).  FLOW_POOL / FLOW_BNRELU reuse the conv fields:
// s.in / s.h / s.w / s.cin = input tensor (cin = channel count = channel stride), e.out_raw = pooled raw
// output (or null), e.out_post = BatchNorm+ReLU'd output (or null), cp.post_s / post_t = its scale / shift.
struct alignas(64) FlowLayer {
  ConvParams p;
  int kind;
  int f;                  // epilogue feature mask incl. F_M64 (FLOW_CONV)
  epi::ChannelParams cp;  // never-null per-channel arrays in global memory
  // images held by each tensor's buffer (ring buffers hold fewer than the layer processes): image i lives in
  // slot i % ring
  int ring_in, ring_pre, ring_raw, ring_post, ring_res1, ring_res2, ring_up, ring_aux;
};

// 16-byte work item: x = layer | img << 16, y = mt | tx << 8 | ty << 16, z = group, w = unused
struct FlowItem {
  int layer_img, tile, group, pad;
};

struct FlowGroup {
  int n_deps;
  int dep[kFlowMaxDeps];       // group indices (counter slots) this group waits for ...
  int dep_need[kFlowMaxDeps];  // ... until done[dep] reaches this count (the tiles of that group)
};

struct FlowSegment {
  FlowLayer* layers = nullptr;  // device (tensor maps are read from here by the TMA unit)
  void* layers_sm = nullptr;    // device: the per-layer parameter block every CTA stages in shared memory
  int n_layers = 0;
  FlowItem* items = nullptr;  // device
  int n_items = 0;
  FlowGroup* groups = nullptr;  // device
  int n_groups = 0;
  unsigned int* done = nullptr;  // device, n_groups counters, zeroed before every launch
};

// Host description of one layer for flow_build_segment.
struct FlowLayerDesc {
  FlowLayer layer;                // p.tm_* / shapes as planned by conv_plan (or the element-wise fields)
  int tiles_x = 0, tiles_y = 0, n_nt = 1;  // tile grid of one image
};

// Builds items / groups for `layers` over `n_views` images in batches of `batch` views, `interleave` batches in
// lock step, derives the dependencies from the tensors' base pointers, and uploads everything (cudaMalloc'd
// pointers are appended to `owned`).  `done` must hold at least n_groups counters.
int flow_build_segment(const std::vector<FlowLayerDesc>& layers, int n_views, int batch, int interleave,
                       std::vector<void*>* owned, FlowSegment* out);
int flow_launch(const FlowSegment& seg, cudaStream_t stream);
// Debug: when non-null, the next flow_launch calls record per-CTA role counters there (148 x 8 int64, see
// conv_flow.cu).
void flow_set_profile_buffer(long long* dev_buf);

}  // namespace mvlm
