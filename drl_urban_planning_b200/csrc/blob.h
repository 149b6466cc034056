// Packed, unpadded graph blob shared by the host packer (pack_host.cpp) and the device kernels.
//
// One blob = one contiguous byte range = one H2D copy.  It replaces the reference's list of per-state
// 9-tuples + per-array tensorfy (urban_planning_agent.py:16-20) and batch_data (state_encoder.py:163-177).
// Layout (every section 16-byte aligned, offsets in bytes from the blob start):
//
//   BlobHeader                       128 B
//   GraphDesc  desc[count]           64 B each
//   float      x[sum_n][24]          node features, rows padded 23 -> 24 floats
//   float      numerical[count][52]
//   float      cur[count][24]        current-node features, padded
//   uint16     rowptr[...]           per graph n+1 entries (CSR over the symmetrised adjacency), padded to 8
//   uint16     order[...]            per graph the pull schedule: rounds x 16 warps x 8 node ids (0xFFFF = none),
//                                    8-node groups of similar degree, assigned to the 16 warps longest-first (LPT)
//   uint32     adj[...]              per graph 2e directed entries: neighbour | (slot+1) << 16 | first << 31, padded to 4
//                                    slot = rank of the entry's undirected edge among the land-use candidates (15 bits);
//                                    first = 1 when the row's node is the edge's first endpoint (edge_index[j][0]): the
//                                    rl-mlp encoder picks an endpoint by position (state_encoder.py:269-276)
//   uint32     cand_uv[...]          per graph k candidates of the active stage: u | v << 16 (edges) or node id
//   int32      cand_idx[...]         original edge / node index of each candidate (action ids), padded to 4
#pragma once
#include <stdint.h>

namespace upb {

constexpr uint32_t kBlobMagic = 0x55504232u;  // "UPB2"
constexpr int kNodeStride = 24;
constexpr int kNumDim = 52;
constexpr int kPullWarps = 16;      // warps of the fused kernel (NT / 32); the pull schedule is laid out for them
constexpr int kPullGroup = 8;       // nodes per warp-task (4 lanes per node)
constexpr uint16_t kNoNode = 0xFFFFu;
constexpr uint32_t kAdjFirst = 0x80000000u;   // adj entry: the row's node is the first endpoint of the edge
constexpr uint32_t kAdjSlotMask = 0x7FFFu;    // (entry >> 16) & kAdjSlotMask = candidate slot + 1

struct BlobHeader {
  uint32_t magic;
  int32_t count;
  int32_t n_cap, e_cap;
  uint64_t total_bytes;
  uint64_t off_desc, off_x, off_num, off_cur, off_rowptr, off_adj, off_cand_uv, off_cand_idx;
  uint64_t sum_n, sum_e, sum_k;
  uint64_t off_order;   // uint16: per graph pull schedule (see GraphDesc::ord_off)
  uint64_t reserved[1];
};
static_assert(sizeof(BlobHeader) == 128, "BlobHeader must be 128 bytes");

struct GraphDesc {
  int32_t n, e, stage, k;
  int32_t x_row;     // first row of this graph in x
  int32_t rp_off;    // first element in rowptr (uint16 units, multiple of 8)
  int32_t adj_off;   // first element in adj (uint32 units, multiple of 4)
  int32_t cand_off;  // first element in cand_uv / cand_idx (multiple of 4)
  int32_t cost;      // work estimate used for static scheduling
  int32_t ord_off;   // first element of this graph's pull schedule in `order` (uint16 units, multiple of 8)
  int32_t ord_rounds;// rounds of the pull schedule: order[ord_off + (r*16 + w)*8 + j] = j-th node of warp w in round r
  int32_t pad[5];
};
static_assert(sizeof(GraphDesc) == 64, "GraphDesc must be 64 bytes");

}  // namespace upb
