// alz_comm.cu — multi-GPU window merge: one rank per GPU, events
// pre-partitioned by alz_owner_rank(saddr), tables replicated. At flush every
// rank contributes its live edges and ONE ncclAllReduce sums the per-edge
// accumulators, after which every rank holds the whole graph in canonical
// (ascending packed key) order — integer sums, so bit-exact for any rank count.
//
//   1. all-gather the local edge counts, then the local sorted keys (padded)
//   2. sort + unique the gathered keys  -> canonical dictionary, same on all ranks
//   3. scatter local rows into a zeroed canonical array [n_can x 35 u64]
//        word 0..2  = count, err5xx, lat_sum ;  word 3..34 = hist cells packed 2 x u32
//      (an edge is owned by one rank, every other rank adds 0, so the packed u32
//       halves cannot carry into each other)
//   4. ncclAllReduce(sum, u64) on that array   <- the single exchange step
//   5. unpack into alz_edge_out rows
//
// NCCL is loaded lazily with dlopen so that single-GPU users (and the Go agent
// on a box without NCCL) never need libnccl.so.2.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

#include "alz_handle.h"

using namespace alz;

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

bool load_nccl() {
  if (g_nccl.lib) return true;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return false;
#define SYM(field, name)                                             \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(lib, name)); \
  if (!g_nccl.field) return false;
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllGather, "ncclAllGather");
  SYM(AllReduce, "ncclAllReduce");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  g_nccl.lib = lib;
  return true;
}

constexpr int kCanWords = 3 + ALZ_NB / 2;  // u64 words per canonical edge row

// out[i] = keys[i] for i < n, kEmptyKey padding up to n_pad
__global__ void pad_keys_kernel(const uint64_t* __restrict__ keys, uint32_t n, uint64_t* __restrict__ out,
                                uint32_t n_pad) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += stride)
    out[i] = i < n ? keys[i] : kEmptyKey;
}

// unique of the sorted gathered keys: flag run heads, exclusive-scan the flags, scatter the heads
__global__ void flag_heads_kernel(const uint64_t* __restrict__ sorted, uint32_t n, uint32_t* __restrict__ flags) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t k = sorted[i];
    flags[i] = (k != kEmptyKey && (i == 0 || sorted[i - 1] != k)) ? 1u : 0u;
  }
}
__global__ void scatter_heads_kernel(const uint64_t* __restrict__ sorted, const uint32_t* __restrict__ flags,
                                     const uint32_t* __restrict__ pos, uint32_t n, uint64_t* __restrict__ out) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    if (flags[i]) out[pos[i]] = sorted[i];
}

__device__ __forceinline__ uint32_t lower_bound_u64(const uint64_t* __restrict__ a, uint32_t n, uint64_t k) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (a[mid] < k) lo = mid + 1; else hi = mid; }
  return lo;
}

// union of R sorted, pairwise disjoint key lists (segment r = gathered[r * pad .. + counts[r])):
// out[sum of lower bounds] = key. *dup is set if a key occurs in two lists.
__global__ void __launch_bounds__(256) merge_disjoint_kernel(const uint64_t* __restrict__ gathered, uint32_t pad,
                                                             uint32_t R, const uint32_t* __restrict__ counts,
                                                             uint64_t* __restrict__ out, uint32_t* __restrict__ dup) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < pad * R; t += stride) {
    const uint32_t r = t / pad, i = t - r * pad;
    if (i >= counts[r]) continue;
    const uint64_t k = gathered[t];
    uint32_t pos = i;
    for (uint32_t q = 0; q < R; ++q) {
      if (q == r) continue;
      const uint64_t* seg = gathered + (size_t)q * pad;
      const uint32_t lb = lower_bound_u64(seg, counts[q], k);
      if (lb < counts[q] && seg[lb] == k) *dup = 1u;
      pos += lb;
    }
    out[pos] = k;
  }
}

// local live rows -> canonical array (zeroed beforehand); a warp per local edge; rows are zeroed
__global__ void __launch_bounds__(256) scatter_canonical_kernel(AccTable edges, const uint64_t* __restrict__ keys,
                                                                const uint32_t* __restrict__ rows, uint32_t n_local,
                                                                const uint64_t* __restrict__ can_keys, uint32_t n_can,
                                                                uint64_t* __restrict__ can) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_local; i += warps_per_grid) {
    const uint32_t row = rows[i];
    const uint32_t pos = lower_bound_u64(can_keys, n_can, keys[i]);   // always present
    uint64_t* dst = can + (size_t)pos * kCanWords;
    const uint32_t c0 = edges.hist[(size_t)row * ALZ_NB + 2 * lane];
    const uint32_t c1 = edges.hist[(size_t)row * ALZ_NB + 2 * lane + 1];
    dst[3 + lane] = ((uint64_t)c1 << 32) | c0;
    edges.hist[(size_t)row * ALZ_NB + 2 * lane] = 0u;
    edges.hist[(size_t)row * ALZ_NB + 2 * lane + 1] = 0u;
    if (lane == 0) {
      dst[0] = edges.count[row]; dst[1] = edges.err5xx[row]; dst[2] = edges.lat_sum[row];
      edges.count[row] = 0ull; edges.err5xx[row] = 0ull; edges.lat_sum[row] = 0ull;
    }
  }
}

__global__ void __launch_bounds__(256) unpack_canonical_kernel(const uint64_t* __restrict__ can_keys,
                                                               const uint64_t* __restrict__ can, uint32_t n_can,
                                                               alz_edge_out* __restrict__ out) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_can; i += warps_per_grid) {
    const uint64_t* src = can + (size_t)i * kCanWords;
    alz_edge_out* o = &out[i];
    const uint64_t w = src[3 + lane];
    o->hist[2 * lane] = (uint32_t)w;
    o->hist[2 * lane + 1] = (uint32_t)(w >> 32);
    if (lane == 0) {
      uint8_t ft, tt; uint32_t f, t;
      unpack_edge_key(can_keys[i], &ft, &f, &tt, &t);
      o->from_type = ft; o->to_type = tt;
      for (int k = 0; k < 6; ++k) o->_pad[k] = 0;
      o->from = f; o->to = t;
      o->count = src[0]; o->err5xx = src[1]; o->lat_sum_ns = src[2];
    }
  }
}

}  // namespace

struct alz_comm_state {
  ncclComm_t comm = nullptr;
  uint32_t* d_counts = nullptr;   // [nranks]
  uint32_t* h_counts = nullptr;   // pinned
  uint64_t* d_gather = nullptr;   // [nranks * pad] gathered keys, then sorted copy behind it
  uint64_t* d_sorted = nullptr;
  uint32_t* d_flags = nullptr;
  uint32_t* d_pos = nullptr;
  uint64_t* d_can_keys = nullptr; // [max_edges]
  uint64_t* d_can = nullptr;      // [max_edges * kCanWords]
  uint32_t* d_iota = nullptr;
  uint32_t* d_vals = nullptr;
  void* d_tmp = nullptr;
  size_t tmp_bytes = 0;
  size_t gather_cap = 0;          // keys
  uint64_t allreduce_bytes = 0;   // of the last window
};

#define CK(expr)                                                                       \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      h->last_err = std::string(#expr) + ": " + cudaGetErrorString(_e);                \
      return ALZ_E_CUDA;                                                               \
    }                                                                                  \
  } while (0)
#define NK(expr)                                                                       \
  do {                                                                                 \
    ncclResult_t _r = (expr);                                                          \
    if (_r != ncclSuccess) {                                                           \
      h->last_err = std::string(#expr) + ": " + g_nccl.GetErrorString(_r);             \
      return ALZ_E_NCCL;                                                               \
    }                                                                                  \
  } while (0)

extern "C" int alz_comm_unique_id(void* out_id) {
  if (!out_id) return ALZ_E_INVAL;
  if (!load_nccl()) return ALZ_E_NCCL;
  static_assert(sizeof(ncclUniqueId) <= ALZ_COMM_ID_BYTES, "ncclUniqueId larger than ALZ_COMM_ID_BYTES");
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return ALZ_E_NCCL;
  memset(out_id, 0, ALZ_COMM_ID_BYTES);
  memcpy(out_id, &id, sizeof(id));
  return ALZ_OK;
}

extern "C" int alz_comm_init(alz_handle* h, int nranks, int rank, const void* id_bytes) {
  if (!h || !id_bytes || nranks < 1 || rank < 0 || rank >= nranks) return ALZ_E_INVAL;
  std::lock_guard<std::mutex> g(h->mu);
  if (h->comm) return ALZ_E_STATE;
  if (!load_nccl()) { h->last_err = "dlopen(libnccl.so.2) failed"; return ALZ_E_NCCL; }
  CK(cudaSetDevice(h->device));
  alz_comm_state* c = new alz_comm_state();
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof(id));
  ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
  if (r != ncclSuccess) { h->last_err = g_nccl.GetErrorString(r); delete c; return ALZ_E_NCCL; }
  h->comm = c;
  h->comm_nranks = nranks;
  h->comm_rank = rank;
  const size_t me = h->cfg.max_edges;
  c->gather_cap = me;  // every rank holds the merged graph, so max_edges bounds the gathered keys too
  CK(cudaMalloc(&c->d_counts, sizeof(uint32_t) * nranks));
  CK(cudaMallocHost(&c->h_counts, sizeof(uint32_t) * nranks));
  CK(cudaMalloc(&c->d_gather, me * 8 * 2));
  c->d_sorted = c->d_gather + me;
  CK(cudaMalloc(&c->d_flags, me * 4));
  CK(cudaMalloc(&c->d_pos, me * 4));
  CK(cudaMalloc(&c->d_can_keys, me * 8));
  CK(cudaMalloc(&c->d_can, me * kCanWords * 8));
  CK(cudaMalloc(&c->d_iota, me * 4));
  CK(cudaMalloc(&c->d_vals, me * 4));
  c->tmp_bytes = std::max(sort_pairs_temp_bytes((uint32_t)me), scan_temp_bytes((uint32_t)me));
  CK(cudaMalloc(&c->d_tmp, c->tmp_bytes));
  return ALZ_OK;
}

void alz_internal_free_comm(alz_handle* h) {
  alz_comm_state* c = h->comm;
  if (!c) return;
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  cudaFree(c->d_counts); if (c->h_counts) cudaFreeHost(c->h_counts);
  cudaFree(c->d_gather); cudaFree(c->d_flags); cudaFree(c->d_pos); cudaFree(c->d_can_keys);
  cudaFree(c->d_can); cudaFree(c->d_iota); cudaFree(c->d_vals); cudaFree(c->d_tmp);
  delete c;
  h->comm = nullptr;
}

// Called by alz_window_flush_device after prepare_flush(): local live edges are
// sorted in d_keys[1] (keys) / d_rows[1] (rows), h->n_live of them.
int alz_internal_merge_ranks(alz_handle* h, int local_rc) {
  alz_comm_state* c = h->comm;
  if (!c || h->comm_nranks <= 1) return ALZ_E_UNSUPPORTED;
  if (local_rc != ALZ_OK) return local_rc;
  const int R = h->comm_nranks;
  cudaStream_t s = h->stream;
  const unsigned grid = (unsigned)h->sms * 4;

  // 1. counts
  CK(cudaMemcpyAsync(c->d_counts + h->comm_rank, &h->n_live, 4, cudaMemcpyHostToDevice, s));
  NK(g_nccl.AllGather(c->d_counts + h->comm_rank, c->d_counts, 1, ncclUint32, c->comm, s));
  CK(cudaMemcpyAsync(c->h_counts, c->d_counts, 4 * R, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  uint32_t pad = 0;
  uint64_t total = 0;
  for (int r = 0; r < R; ++r) { pad = std::max(pad, c->h_counts[r]); total += c->h_counts[r]; }
  if (total == 0) {
    h->last_n_edges = 0;
    h->windows++;
    CK(cudaMemsetAsync(h->edges.dict, 0xFF, ((size_t)h->edges.dict_mask + 1) * sizeof(DictEnt), s));
    CK(cudaMemsetAsync(h->edges.n_rows, 0, 4, s));
    return ALZ_OK;
  }
  if ((uint64_t)pad * R > c->gather_cap) return ALZ_E_CAPACITY;
  const uint32_t n_g = pad * (uint32_t)R;

  // 2. keys: pad, all-gather, merge. Each rank's list is sorted and, when the caller partitioned by
  //    alz_owner_rank, the lists are disjoint: a key's place in the union is then the sum of its lower
  //    bounds in the R lists - no sort. A key found on two ranks (caller routed one source to two ranks)
  //    raises a flag and the general path (sort + unique) runs instead; the sums are right either way.
  pad_keys_kernel<<<grid, 256, 0, s>>>(h->d_keys[1], h->n_live, c->d_sorted, pad);
  NK(g_nccl.AllGather(c->d_sorted, c->d_gather, pad, ncclUint64, c->comm, s));
  CK(cudaMemsetAsync(c->d_flags, 0, 4, s));
  merge_disjoint_kernel<<<grid, 256, 0, s>>>(c->d_gather, pad, (uint32_t)R, c->d_counts, c->d_can_keys, c->d_flags);
  uint32_t dup = 0;
  CK(cudaMemcpyAsync(&dup, c->d_flags, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  uint32_t n_can = (uint32_t)total;
  if (dup) {
    launch_iota(c->d_iota, n_g, h->sms, s);
    sort_pairs(c->d_tmp, c->tmp_bytes, c->d_gather, c->d_sorted, c->d_iota, c->d_vals, n_g, s);
    flag_heads_kernel<<<grid, 256, 0, s>>>(c->d_sorted, n_g, c->d_flags);
    exclusive_scan_u32(c->d_tmp, c->tmp_bytes, c->d_flags, c->d_pos, n_g, s);
    scatter_heads_kernel<<<grid, 256, 0, s>>>(c->d_sorted, c->d_flags, c->d_pos, n_g, c->d_can_keys);
    uint32_t last[2];
    CK(cudaMemcpyAsync(&last[0], c->d_pos + (n_g - 1), 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&last[1], c->d_flags + (n_g - 1), 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    n_can = last[0] + last[1];
  }
  if (n_can > h->cfg.max_edges) return ALZ_E_CAPACITY;

  // 3. scatter local rows into the zeroed canonical array
  const size_t can_bytes = (size_t)n_can * kCanWords * 8;
  CK(cudaMemsetAsync(c->d_can, 0, can_bytes, s));
  if (h->n_live)
    scatter_canonical_kernel<<<grid * 2, 256, 0, s>>>(h->edges, h->d_keys[1], h->d_rows[1], h->n_live,
                                                      c->d_can_keys, n_can, c->d_can);
  // 4. the single exchange step
  NK(g_nccl.AllReduce(c->d_can, c->d_can, (size_t)n_can * kCanWords, ncclUint64, ncclSum, c->comm, s));
  c->allreduce_bytes = can_bytes;
  // 5. unpack; local edge table back to empty
  unpack_canonical_kernel<<<grid * 2, 256, 0, s>>>(c->d_can_keys, c->d_can, n_can, h->d_out);
  CK(cudaGetLastError());
  CK(cudaMemsetAsync(h->edges.dict, 0xFF, ((size_t)h->edges.dict_mask + 1) * sizeof(DictEnt), s));
  CK(cudaMemsetAsync(h->edges.n_rows, 0, 4, s));
  h->last_n_edges = n_can;
  h->windows++;
  return ALZ_OK;
}
