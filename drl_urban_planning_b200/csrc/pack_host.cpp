// Host-side packer: reference 9-array states -> one unpadded blob (see blob.h).  Pure CPU, no CUDA.
// Replaces tensorfy + batch_data of the reference (urban_planning_agent.py:16-20, state_encoder.py:163-177).
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/upb200.h"
#include "blob.h"
#include "errors.h"

namespace upb {

namespace {

struct Counts {
  int n, e, k, stage;
};

inline uint64_t align16(uint64_t v) { return (v + 15) & ~uint64_t(15); }

struct StateView {
  const float* numerical;
  const float* node_features;
  const int64_t* edge_index;
  const float* current_node;
  const uint8_t* node_mask;
  const uint8_t* edge_mask;
  const uint8_t* land_use_mask;
  const uint8_t* road_mask;
  const float* stage;
};

inline StateView view(const void* const* arrays, int i) {
  const void* const* a = arrays + 9 * (size_t)i;
  return StateView{(const float*)a[0],   (const float*)a[1],   (const int64_t*)a[2],
                   (const float*)a[3],   (const uint8_t*)a[4], (const uint8_t*)a[5],
                   (const uint8_t*)a[6], (const uint8_t*)a[7], (const float*)a[8]};
}

inline int count_prefix(const uint8_t* m, int cap, bool* is_prefix) {
  int c = 0;
  for (int i = 0; i < cap; ++i) c += m[i] != 0;
  bool ok = true;
  for (int i = 0; i < c; ++i) ok &= (m[i] != 0);
  *is_prefix = ok;
  return c;
}

// Validates one state against the layout contract and returns its sizes.  Returns nullptr or an error text.
const char* measure_one(const StateView& s, int n_cap, int e_cap, Counts* out) {
  for (int j = 0; j < 9; ++j)
    if (((const void* const*)&s)[j] == nullptr) return "null array pointer";
  bool pn, pe;
  const int n = count_prefix(s.node_mask, n_cap, &pn);
  const int e = count_prefix(s.edge_mask, e_cap, &pe);
  if (!pn || !pe) return "node_mask / edge_mask must be prefix masks (observation_extractor.py:60-66)";
  if (n < 1) return "a state needs at least one node";
  int stage;
  if (s.stage[0] != 0.f && s.stage[1] == 0.f) stage = 0;
  else if (s.stage[1] != 0.f && s.stage[0] == 0.f) stage = 1;
  else return "stage must be one-hot on 'land_use' or 'road' (stored states are pre-step states)";
  for (int j = 0; j < e; ++j) {
    const int64_t u = s.edge_index[2 * j], v = s.edge_index[2 * j + 1];
    if (u < 0 || v < 0 || u >= n || v >= n) return "a real edge joins a padded node";
  }
  int k = 0;
  if (stage == 0) {
    for (int j = 0; j < e_cap; ++j)
      if (s.land_use_mask[j]) {
        if (j >= e) return "land_use_mask marks a padded edge";
        ++k;
      }
  } else {
    for (int i = 0; i < n_cap; ++i)
      if (s.road_mask[i]) {
        if (i >= n) return "road_mask marks a padded node";
        ++k;
      }
  }
  if (k > 65534) return "more than 65534 action candidates";
  *out = Counts{n, e, k, stage};
  return nullptr;
}

struct Plan {
  std::vector<Counts> counts;
  std::vector<GraphDesc> desc;
  BlobHeader hdr;
};

template <class F>
void parallel_for(int count, int threads, F&& fn) {
  // spawning a thread costs tens of microseconds: a handful of them saturates the memory system for this copy-bound
  // work, a hundred (hardware_concurrency on a big host) would cost more than the packing itself
  if (threads <= 0) threads = (int)std::min(12u, std::max(1u, std::thread::hardware_concurrency()));
  threads = std::min(threads, std::max(1, count / 16));
  if (threads <= 1) {
    for (int i = 0; i < count; ++i) fn(i);
    return;
  }
  std::atomic<int> next{0};
  std::vector<std::thread> pool;
  auto worker = [&]() {
    for (;;) {
      const int lo = next.fetch_add(8);
      if (lo >= count) break;
      const int hi = std::min(count, lo + 8);
      for (int i = lo; i < hi; ++i) fn(i);
    }
  };
  for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
  worker();
  for (auto& th : pool) th.join();
}

int make_plan(int count, const void* const* arrays, int n_cap, int e_cap, int threads, Plan* plan) {
  if (count < 0 || (count > 0 && arrays == nullptr)) return set_error(UPB_ERR_ARG, "pack: bad count / arrays");
  if (n_cap < 1 || n_cap > 65535 || e_cap < 0 || 2 * (int64_t)e_cap > 65535)
    return set_error(UPB_ERR_ARG, "pack: caps must satisfy n_cap <= 65535 and 2*e_cap <= 65535");
  plan->counts.assign(count, Counts{0, 0, 0, 0});
  std::atomic<int> bad{-1};
  std::vector<const char*> why(count, nullptr);
  parallel_for(count, threads, [&](int i) {
    why[i] = measure_one(view(arrays, i), n_cap, e_cap, &plan->counts[i]);
    if (why[i]) {
      int expected = -1;
      bad.compare_exchange_strong(expected, i);
    }
  });
  if (bad.load() >= 0) {
    int first = -1;
    for (int i = 0; i < count; ++i)
      if (why[i]) { first = i; break; }
    char buf[256];
    snprintf(buf, sizeof(buf), "pack: state %d: %s", first, why[first]);
    return set_error(UPB_ERR_FORMAT, buf);
  }
  plan->desc.assign(count, GraphDesc{});
  uint64_t rows = 0, rp = 0, adj = 0, cand = 0, sum_e = 0, ord = 0;
  for (int i = 0; i < count; ++i) {
    const Counts& c = plan->counts[i];
    GraphDesc& d = plan->desc[i];
    d.n = c.n; d.e = c.e; d.stage = c.stage; d.k = c.k;
    d.x_row = (int32_t)rows;
    d.rp_off = (int32_t)rp;
    d.adj_off = (int32_t)adj;
    d.cand_off = (int32_t)cand;
    d.cost = 4 * c.e + c.n + 64;
    d.ord_off = (int32_t)ord;
    d.ord_rounds = ((c.n + kPullGroup - 1) / kPullGroup + kPullWarps - 1) / kPullWarps + 1;   // upper bound, see fill_one
    ord += (uint64_t)d.ord_rounds * kPullWarps * kPullGroup;
    rows += c.n;
    rp += (uint64_t)((c.n + 1 + 7) & ~7);
    adj += (uint64_t)((2 * c.e + 3) & ~3);
    cand += (uint64_t)((c.k + 3) & ~3);
    sum_e += c.e;
    if (rows > 0x7fffffffull || adj > 0x7fffffffull) return set_error(UPB_ERR_CAPACITY, "pack: blob too large");
  }
  BlobHeader& h = plan->hdr;
  memset(&h, 0, sizeof(h));
  h.magic = kBlobMagic;
  h.count = count;
  h.n_cap = n_cap;
  h.e_cap = e_cap;
  uint64_t off = sizeof(BlobHeader);
  h.off_desc = off;      off = align16(off + sizeof(GraphDesc) * (uint64_t)count);
  h.off_x = off;         off = align16(off + rows * kNodeStride * sizeof(float));
  h.off_num = off;       off = align16(off + (uint64_t)count * kNumDim * sizeof(float));
  h.off_cur = off;       off = align16(off + (uint64_t)count * kNodeStride * sizeof(float));
  h.off_rowptr = off;    off = align16(off + rp * sizeof(uint16_t));
  h.off_order = off;     off = align16(off + ord * sizeof(uint16_t));
  h.off_adj = off;       off = align16(off + adj * sizeof(uint32_t));
  h.off_cand_uv = off;   off = align16(off + cand * sizeof(uint32_t));
  h.off_cand_idx = off;  off = align16(off + cand * sizeof(int32_t));
  h.total_bytes = off;
  h.sum_n = rows;
  h.sum_e = sum_e;
  uint64_t sk = 0;
  for (const Counts& c : plan->counts) sk += c.k;
  h.sum_k = sk;
  return UPB_OK;
}

void fill_one(const StateView& s, const GraphDesc& d, const BlobHeader& h, uint8_t* blob) {
  const int n = d.n, e = d.e;
  float* x = (float*)(blob + h.off_x) + (size_t)d.x_row * kNodeStride;
  for (int i = 0; i < n; ++i) {
    memcpy(x + (size_t)i * kNodeStride, s.node_features + (size_t)i * UPB_NODE_DIM, UPB_NODE_DIM * sizeof(float));
    x[(size_t)i * kNodeStride + UPB_NODE_DIM] = 0.f;
  }
  uint16_t* rp = (uint16_t*)(blob + h.off_rowptr) + d.rp_off;
  uint32_t* adj = (uint32_t*)(blob + h.off_adj) + d.adj_off;
  uint32_t* cuv = (uint32_t*)(blob + h.off_cand_uv) + d.cand_off;
  int32_t* cidx = (int32_t*)(blob + h.off_cand_idx) + d.cand_off;
  // degree count -> CSR row pointers over the symmetrised adjacency
  std::vector<int> pos(n + 1, 0);
  for (int j = 0; j < e; ++j) {
    pos[(int)s.edge_index[2 * j] + 1]++;
    pos[(int)s.edge_index[2 * j + 1] + 1]++;
  }
  for (int i = 0; i < n; ++i) pos[i + 1] += pos[i];
  for (int i = 0; i <= n; ++i) rp[i] = (uint16_t)pos[i];
  for (int i = n + 1; i < ((n + 1 + 7) & ~7); ++i) rp[i] = (uint16_t)pos[n];
  {  // Pull schedule.  Nodes sorted by descending degree are cut into groups of 8 (one warp-task: 4 lanes per node,
     // trip count = the group's largest degree); groups are dealt to the 16 warps longest-first onto the least
     // loaded warp (LPT), so all warps finish a pull phase at about the same time.
    uint16_t* ord = (uint16_t*)(blob + h.off_order) + d.ord_off;
    const int slots = d.ord_rounds * kPullWarps * kPullGroup;
    for (int i = 0; i < slots; ++i) ord[i] = kNoNode;
    std::vector<int> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return pos[a + 1] - pos[a] > pos[b + 1] - pos[b]; });
    const int groups = (n + kPullGroup - 1) / kPullGroup;
    int load[kPullWarps] = {0}, used[kPullWarps] = {0};
    for (int gi = 0; gi < groups; ++gi) {
      const int first = idx[gi * kPullGroup];
      const int cost = (pos[first + 1] - pos[first] + 1) / 2 + 2;     // trips of two neighbours + fixed overhead
      int best = -1;
      for (int w = 0; w < kPullWarps; ++w)
        if (used[w] < d.ord_rounds && (best < 0 || load[w] < load[best])) best = w;
      load[best] += cost;
      const int r = used[best]++;
      for (int j = 0; j < kPullGroup && gi * kPullGroup + j < n; ++j)
        ord[(r * kPullWarps + best) * kPullGroup + j] = (uint16_t)idx[gi * kPullGroup + j];
    }
  }
  int slot = 0;
  for (int j = 0; j < e; ++j) {
    const uint32_t u = (uint32_t)s.edge_index[2 * j], v = (uint32_t)s.edge_index[2 * j + 1];
    uint32_t tag = 0;
    if (d.stage == 0 && s.land_use_mask[j]) {
      cuv[slot] = u | (v << 16);
      cidx[slot] = j;
      tag = (uint32_t)(slot + 1) << 16;
      ++slot;
    }
    adj[pos[u]++] = v | tag;
    adj[pos[v]++] = u | tag;
  }
  for (int a = 2 * e; a < ((2 * e + 3) & ~3); ++a) adj[a] = 0;
  if (d.stage == 1) {
    for (int i = 0; i < n; ++i)
      if (s.road_mask[i]) {
        cuv[slot] = (uint32_t)i;
        cidx[slot] = i;
        ++slot;
      }
  }
  for (int c = slot; c < ((d.k + 3) & ~3); ++c) { cuv[c] = 0; cidx[c] = 0; }
}

}  // namespace

}  // namespace upb

using namespace upb;

extern "C" int upb_pack_measure(int count, const void* const* state_arrays, int n_cap, int e_cap, int threads,
                                uint64_t* blob_bytes) {
  if (!blob_bytes) return set_error(UPB_ERR_ARG, "pack_measure: blob_bytes is null");
  Plan plan;
  int rc = make_plan(count, state_arrays, n_cap, e_cap, threads, &plan);
  if (rc != UPB_OK) return rc;
  *blob_bytes = plan.hdr.total_bytes;
  return UPB_OK;
}

extern "C" int upb_pack_fill(int count, const void* const* state_arrays, int n_cap, int e_cap, int threads,
                             void* blob_host, uint64_t blob_bytes) {
  if (!blob_host || ((uintptr_t)blob_host & 15)) return set_error(UPB_ERR_ARG, "pack_fill: blob must be 16-byte aligned");
  Plan plan;
  int rc = make_plan(count, state_arrays, n_cap, e_cap, threads, &plan);
  if (rc != UPB_OK) return rc;
  if (blob_bytes < plan.hdr.total_bytes) return set_error(UPB_ERR_CAPACITY, "pack_fill: blob buffer too small");
  uint8_t* blob = (uint8_t*)blob_host;
  memcpy(blob, &plan.hdr, sizeof(BlobHeader));
  if (count > 0) memcpy(blob + plan.hdr.off_desc, plan.desc.data(), sizeof(GraphDesc) * (size_t)count);
  float* num = (float*)(blob + plan.hdr.off_num);
  float* cur = (float*)(blob + plan.hdr.off_cur);
  parallel_for(count, threads, [&](int i) {
    const StateView s = view(state_arrays, i);
    memcpy(num + (size_t)i * kNumDim, s.numerical, kNumDim * sizeof(float));
    memcpy(cur + (size_t)i * kNodeStride, s.current_node, UPB_NODE_DIM * sizeof(float));
    cur[(size_t)i * kNodeStride + UPB_NODE_DIM] = 0.f;
    fill_one(s, plan.desc[i], plan.hdr, blob);
  });
  return UPB_OK;
}

extern "C" int upb_blob_info(const void* blob_host, uint64_t blob_bytes, int* count, int32_t* per_graph4) {
  if (!blob_host || blob_bytes < sizeof(BlobHeader)) return set_error(UPB_ERR_ARG, "blob_info: bad blob");
  const BlobHeader* h = (const BlobHeader*)blob_host;
  if (h->magic != kBlobMagic || h->total_bytes > blob_bytes) return set_error(UPB_ERR_FORMAT, "blob_info: not a blob");
  if (count) *count = h->count;
  if (per_graph4) {
    const GraphDesc* d = (const GraphDesc*)((const uint8_t*)blob_host + h->off_desc);
    for (int i = 0; i < h->count; ++i) {
      per_graph4[4 * i + 0] = d[i].n;
      per_graph4[4 * i + 1] = d[i].e;
      per_graph4[4 * i + 2] = d[i].k;
      per_graph4[4 * i + 3] = d[i].stage;
    }
  }
  return UPB_OK;
}
