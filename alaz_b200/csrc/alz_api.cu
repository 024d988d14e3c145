// alz_api.cu — the C ABI of libalazgpu (include/alazgpu.h): handle, HBM layout,
// host staging and the launch sequence of each entry point. Host side C++; the
// reference's Go caller reaches it through cgo.
//
// There is no CPU fallback anywhere in this file: without a CUDA device
// alz_create fails with ALZ_E_NODEVICE.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "alz_handle.h"

using namespace alz;

#define CK(expr)                                                                       \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      h->last_err = std::string(#expr) + ": " + cudaGetErrorString(_e);                \
      return ALZ_E_CUDA;                                                               \
    }                                                                                  \
  } while (0)

static uint32_t next_pow2(uint64_t x) {
  uint64_t p = 1;
  while (p < x) p <<= 1;
  return (uint32_t)std::min<uint64_t>(p, 1ull << 31);
}

// pinned buffers handed to callers (alz_pinned_alloc): submits from these skip
// the staging memcpy
static std::mutex g_pin_mu;
static std::vector<std::pair<const char*, size_t>> g_pinned;
static bool is_lib_pinned(const void* p, size_t bytes) {
  std::lock_guard<std::mutex> g(g_pin_mu);
  const char* c = (const char*)p;
  for (auto& r : g_pinned)
    if (c >= r.first && c + bytes <= r.first + r.second) return true;
  return false;
}

static int alloc_table(alz_handle* h, AccTable* t, uint32_t max_rows, uint32_t* n_rows_dev, bool with_count) {
  t->max_rows = max_rows;
  t->n_rows = n_rows_dev;
  const uint32_t dict_cap = next_pow2(2ull * max_rows);
  t->dict_mask = dict_cap - 1;
  const size_t rows = (size_t)max_rows + 1;
  CK(cudaMalloc(&t->dict, (size_t)dict_cap * sizeof(DictEnt)));
  CK(cudaMalloc(&t->row_key, rows * 8));
  CK(cudaMalloc(&t->lat_sum, rows * 8));
  CK(cudaMalloc(&t->err5xx, rows * 8));
  t->count = nullptr;
  t->row_cnt = nullptr;
  t->row_aux = nullptr;
  if (with_count) CK(cudaMalloc(&t->count, rows * 8));
  else {
    CK(cudaMalloc(&t->row_cnt, rows * 4)); CK(cudaMemsetAsync(t->row_cnt, 0, rows * 4, h->stream));
    CK(cudaMalloc(&t->row_aux, rows * 4));
  }
  CK(cudaMalloc(&t->hist, rows * ALZ_NB * 4));
  CK(cudaMemsetAsync(t->dict, 0xFF, (size_t)dict_cap * sizeof(DictEnt), h->stream));
  CK(cudaMemsetAsync(t->lat_sum, 0, rows * 8, h->stream));
  CK(cudaMemsetAsync(t->err5xx, 0, rows * 8, h->stream));
  if (with_count) CK(cudaMemsetAsync(t->count, 0, rows * 8, h->stream));
  CK(cudaMemsetAsync(t->hist, 0, rows * ALZ_NB * 4, h->stream));
  return ALZ_OK;
}
static void free_table(AccTable* t) {
  cudaFree(t->dict); cudaFree(t->row_key); cudaFree(t->lat_sum); cudaFree(t->err5xx); cudaFree(t->count);
  cudaFree(t->row_cnt); cudaFree(t->row_aux);
  cudaFree(t->hist);
  memset(t, 0, sizeof(*t));
}
// all keys out of the dictionary, row allocator back to zero (rows were zeroed by fold / gather)
static int clear_dict(alz_handle* h, AccTable* t) {
  CK(cudaMemsetAsync(t->dict, 0xFF, ((size_t)t->dict_mask + 1) * sizeof(DictEnt), h->stream));
  CK(cudaMemsetAsync(t->n_rows, 0, 4, h->stream));
  return ALZ_OK;
}

extern "C" const char* alz_strerror(int s) {
  switch (s) {
    case ALZ_OK: return "ok";
    case ALZ_E_INVAL: return "invalid argument";
    case ALZ_E_NOMEM: return "out of memory";
    case ALZ_E_CUDA: return "CUDA error";
    case ALZ_E_NODEVICE: return "no CUDA device (libalazgpu has no CPU fallback)";
    case ALZ_E_CAPACITY: return "capacity exceeded";
    case ALZ_E_STATE: return "invalid state";
    case ALZ_E_NCCL: return "NCCL error";
    case ALZ_E_UNSUPPORTED: return "unsupported";
    default: return "unknown status";
  }
}

extern "C" const char* alz_last_cuda_error(alz_handle* h) { return h ? h->last_err.c_str() : ""; }

extern "C" int alz_create(const alz_config* cfg, alz_handle** out) {
  if (!cfg || !out || cfg->abi_version != ALZ_ABI_VERSION) return ALZ_E_INVAL;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return ALZ_E_NODEVICE;
  if (cfg->device < 0 || cfg->device >= ndev) return ALZ_E_INVAL;
  alz_handle* h = new (std::nothrow) alz_handle();
  if (!h) return ALZ_E_NOMEM;
  h->cfg = *cfg;
  if (h->cfg.max_endpoints == 0) h->cfg.max_endpoints = 1u << 16;
  if (h->cfg.max_pairs == 0) h->cfg.max_pairs = 1u << 20;
  if (h->cfg.max_edges == 0) h->cfg.max_edges = h->cfg.max_pairs;
  if (h->cfg.max_batch == 0) h->cfg.max_batch = 1u << 22;
  h->device = cfg->device;
  auto fail = [&](int rc) { alz_destroy(h); return rc; };
#define CKC(expr)                                                                      \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      fprintf(stderr, "libalazgpu: %s: %s\n", #expr, cudaGetErrorString(_e));          \
      return fail(_e == cudaErrorMemoryAllocation ? ALZ_E_NOMEM : ALZ_E_CUDA);         \
    }                                                                                  \
  } while (0)
  CKC(cudaSetDevice(h->device));
  cudaDeviceProp prop;
  CKC(cudaGetDeviceProperties(&prop, h->device));
  h->sms = prop.multiProcessorCount;
  CKC(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  CKC(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  h->stream = h->own_stream;
  for (int b = 0; b < 2; ++b) {
    CKC(cudaEventCreateWithFlags(&h->ev_copied[b], cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&h->ev_consumed[b], cudaEventDisableTiming));
  }
  CKC(cudaEventCreateWithFlags(&h->ev_tmp, cudaEventDisableTiming));

  h->ep_cap = next_pow2(2ull * h->cfg.max_endpoints);
  CKC(cudaMalloc(&h->d_ep, (size_t)h->ep_cap * sizeof(EpEntry)));
  CKC(cudaMemsetAsync(h->d_ep, 0, (size_t)h->ep_cap * sizeof(EpEntry), h->stream));
  CKC(cudaMalloc(&h->d_ctr, sizeof(Counters)));
  CKC(cudaMemsetAsync(h->d_ctr, 0, sizeof(Counters), h->stream));
  CKC(cudaMalloc(&h->d_hot, 2 * sizeof(HotState)));
  CKC(cudaMemsetAsync(h->d_hot, 0, 2 * sizeof(HotState), h->stream));
  int rc;
  if (!(h->cfg.flags & ALZ_CFG_EAGER_JOIN)) {
    if ((rc = alloc_table(h, &h->pairs_fwd, h->cfg.max_pairs, &h->d_ctr->fwd_rows, false)) != ALZ_OK) return fail(rc);
    if ((rc = alloc_table(h, &h->pairs_rev, std::max<uint32_t>(1024u, h->cfg.max_pairs >> 1), &h->d_ctr->rev_rows,
                          false)) != ALZ_OK) return fail(rc);
  }
  if ((rc = alloc_table(h, &h->edges, h->cfg.max_edges, &h->d_ctr->edge_rows, true)) != ALZ_OK) return fail(rc);
  CKC(cudaMallocHost(&h->h_ctr, sizeof(Counters)));
  for (int b = 0; b < 2; ++b) {
    CKC(cudaMalloc(&h->d_keys[b], (size_t)h->cfg.max_edges * 8));
    CKC(cudaMalloc(&h->d_rows[b], (size_t)h->cfg.max_edges * 4));
  }
  h->sort_tmp_bytes = sort_pairs_temp_bytes(h->cfg.max_edges);
  CKC(cudaMalloc(&h->d_sort_tmp, h->sort_tmp_bytes));
  CKC(cudaMalloc(&h->d_out, (size_t)h->cfg.max_edges * sizeof(alz_edge_out)));
  CKC(cudaStreamSynchronize(h->stream));
#undef CKC
  *out = h;
  return ALZ_OK;
}

extern "C" int alz_destroy(alz_handle* h) {
  if (!h) return ALZ_E_INVAL;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  alz_internal_free_extensions(h);
  free_table(&h->pairs_fwd); free_table(&h->pairs_rev); free_table(&h->edges);
  cudaFree(h->d_ep); cudaFree(h->d_ctr); cudaFree(h->d_hot);
  if (h->h_ctr) cudaFreeHost(h->h_ctr);
  for (int b = 0; b < 2; ++b) {
    cudaFree(h->d_keys[b]); cudaFree(h->d_rows[b]);
    cudaFree(h->d_stage[b]);
    if (h->h_stage[b]) cudaFreeHost(h->h_stage[b]);
    if (h->ev_copied[b]) cudaEventDestroy(h->ev_copied[b]);
    if (h->ev_consumed[b]) cudaEventDestroy(h->ev_consumed[b]);
  }
  if (h->ev_tmp) cudaEventDestroy(h->ev_tmp);
  cudaFree(h->d_raw_stage);
  if (h->h_raw_stage) cudaFreeHost(h->h_raw_stage);
  cudaFree(h->d_sort_tmp); cudaFree(h->d_out);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  delete h;
  return ALZ_OK;
}

extern "C" int alz_set_stream(alz_handle* h, void* s) {
  if (!h) return ALZ_E_INVAL;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  h->stream = s ? (cudaStream_t)s : h->own_stream;
  return ALZ_OK;
}

extern "C" int alz_sync(alz_handle* h) {
  if (!h) return ALZ_E_INVAL;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->copy_stream));
  CK(cudaStreamSynchronize(h->stream));
  return ALZ_OK;
}

// ---- pinned host memory for callers ----------------------------------------------
extern "C" int alz_pinned_alloc(size_t bytes, void** out) {
  if (!out || bytes == 0) return ALZ_E_INVAL;
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) return ALZ_E_NOMEM;
  std::lock_guard<std::mutex> g(g_pin_mu);
  g_pinned.emplace_back((const char*)p, bytes);
  *out = p;
  return ALZ_OK;
}
extern "C" int alz_pinned_free(void* p) {
  if (!p) return ALZ_E_INVAL;
  {
    std::lock_guard<std::mutex> g(g_pin_mu);
    for (size_t i = 0; i < g_pinned.size(); ++i)
      if (g_pinned[i].first == (const char*)p) { g_pinned.erase(g_pinned.begin() + i); break; }
  }
  return cudaFreeHost(p) == cudaSuccess ? ALZ_OK : ALZ_E_CUDA;
}

// ---- join build side -----------------------------------------------------------------
extern "C" int alz_table_upsert(alz_handle* h, int table, uint32_t ip, uint32_t id) {
  if (!h || (table != ALZ_TABLE_POD && table != ALZ_TABLE_SVC) || id >= (1u << 29)) return ALZ_E_INVAL;
  HostEp& e = h->ep_host[ip];
  if (table == ALZ_TABLE_POD) { e.state |= kEpPod; e.pod = id; }   // persist.go:55-65
  else { e.state |= kEpSvc; e.svc = id; }                          // persist.go:114-124
  h->ep_dirty = true;
  return ALZ_OK;
}
extern "C" int alz_table_erase(alz_handle* h, int table, uint32_t ip) {
  if (!h || (table != ALZ_TABLE_POD && table != ALZ_TABLE_SVC)) return ALZ_E_INVAL;
  auto it = h->ep_host.find(ip);
  if (it == h->ep_host.end()) return ALZ_OK;                       // delete of a missing key: no-op
  it->second.state &= ~(table == ALZ_TABLE_POD ? kEpPod : kEpSvc); // persist.go:66-70, :125-129
  if (it->second.state == 0) h->ep_host.erase(it);
  h->ep_dirty = true;
  return ALZ_OK;
}

int alz_internal_fold(alz_handle* h) {
  if (h->cfg.flags & ALZ_CFG_EAGER_JOIN) return ALZ_OK;
  if (h->pending_since_fold == 0) return ALZ_OK;
  // the join on distinct pairs; it also leaves per-pair counts behind, from which the next
  // ingest launches learn which pairs are hot (alz_ingest.cu)
  CK(cudaMemsetAsync(h->d_hot[0].bins, 0, sizeof(h->d_hot[0].bins), h->stream));
  CK(cudaMemsetAsync(h->d_hot[1].bins, 0, sizeof(h->d_hot[1].bins), h->stream));
  launch_fold_pairs(h->pairs_fwd, false, h->d_ep, h->ep_cap - 1, h->edges, h->d_ctr, h->d_hot[0].bins, h->sms, h->stream);
  launch_fold_pairs(h->pairs_rev, true, h->d_ep, h->ep_cap - 1, h->edges, h->d_ctr, h->d_hot[1].bins, h->sms, h->stream);
  launch_hot_select(h->pairs_fwd, &h->d_hot[0], false, h->sms, h->stream);
  launch_hot_select(h->pairs_rev, &h->d_hot[1], true, h->sms, h->stream);
  CK(cudaGetLastError());
  int rc = clear_dict(h, &h->pairs_fwd);
  if (rc == ALZ_OK) rc = clear_dict(h, &h->pairs_rev);
  h->pending_since_fold = 0;
  return rc;
}

extern "C" int alz_table_commit(alz_handle* h) {
  if (!h) return ALZ_E_INVAL;
  CK(cudaSetDevice(h->device));
  if (!h->ep_dirty) return ALZ_OK;
  if (h->ep_host.size() > h->cfg.max_endpoints) return ALZ_E_CAPACITY;
  // events already submitted are joined through the tables they arrived under
  int rc = alz_internal_fold(h);
  if (rc != ALZ_OK) return rc;
  std::vector<EpEntry> tab(h->ep_cap);
  memset(tab.data(), 0, tab.size() * sizeof(EpEntry));
  const uint32_t mask = h->ep_cap - 1;
  for (auto& kv : h->ep_host) {
    uint32_t slot = hash32(kv.first) & mask;
    while (tab[slot].state & kEpOcc) slot = (slot + 1) & mask;
    tab[slot].ip = kv.first;
    tab[slot].state = kEpOcc | kv.second.state;
    tab[slot].pod = kv.second.pod;
    tab[slot].svc = kv.second.svc;
  }
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpy(h->d_ep, tab.data(), tab.size() * sizeof(EpEntry), cudaMemcpyHostToDevice));
  h->ep_dirty = false;
  return ALZ_OK;
}

// ---- ingest ------------------------------------------------------------------------------
static int ingest_device(alz_handle* h, const alz_l7_rec* d, uint64_t n) {
  if (h->cfg.flags & ALZ_CFG_EAGER_JOIN)
    launch_ingest_eager(d, n, h->d_ep, h->ep_cap - 1, h->edges, h->d_ctr, h->sms, h->stream);
  else
  {
    if (h->cfg.flags & ALZ_CFG_NO_SMEM_CACHE)
      launch_ingest_pairs_v1(d, n, h->pairs_fwd, h->pairs_rev, h->d_ctr, h->sms, h->stream);
    else
      launch_ingest_pairs_v4(d, n, h->pairs_fwd, h->pairs_rev, h->d_ctr, &h->d_hot[0], &h->d_hot[1], h->d_ep,
                             h->ep_cap - 1, h->sms, h->stream);
  }
  CK(cudaGetLastError());
  h->events_in += n;
  h->pending_since_fold += n;
  // pair histograms are u32: fold before any bucket could wrap
  if (h->pending_since_fold >= (1ull << 31)) return alz_internal_fold(h);
  return ALZ_OK;
}

extern "C" int alz_submit_l7_device(alz_handle* h, const alz_l7_rec* d, size_t n) {
  if (!h || (!d && n)) return ALZ_E_INVAL;
  if (((uintptr_t)d & 31u) != 0) return ALZ_E_INVAL;  // 256-bit loads
  CK(cudaSetDevice(h->device));
  return ingest_device(h, d, n);
}

static int ensure_stage(alz_handle* h) {
  if (h->d_stage[0]) return ALZ_OK;
  const size_t bytes = (size_t)h->cfg.max_batch * sizeof(alz_l7_rec);
  for (int b = 0; b < 2; ++b) {
    CK(cudaMalloc(&h->d_stage[b], bytes));
    CK(cudaMallocHost(&h->h_stage[b], bytes));
  }
  return ALZ_OK;
}

extern "C" int alz_submit_l7(alz_handle* h, const alz_l7_rec* recs, size_t n) {
  if (!h || (!recs && n)) return ALZ_E_INVAL;
  CK(cudaSetDevice(h->device));
  int rc = ensure_stage(h);
  if (rc != ALZ_OK) return rc;
  const bool direct = is_lib_pinned(recs, n * sizeof(alz_l7_rec));
  size_t done = 0;
  while (done < n) {
    const size_t m = std::min<size_t>(h->cfg.max_batch, n - done);
    const int b = (int)(h->stage_turn++ & 1u);
    const void* src = recs + done;
    if (!direct) {
      CK(cudaEventSynchronize(h->ev_copied[b]));   // previous H2D out of this pinned buffer is done
      memcpy(h->h_stage[b], recs + done, m * sizeof(alz_l7_rec));
      src = h->h_stage[b];
    }
    CK(cudaStreamWaitEvent(h->copy_stream, h->ev_consumed[b], 0));  // kernel that read d_stage[b]
    CK(cudaMemcpyAsync(h->d_stage[b], src, m * sizeof(alz_l7_rec), cudaMemcpyHostToDevice, h->copy_stream));
    CK(cudaEventRecord(h->ev_copied[b], h->copy_stream));
    CK(cudaStreamWaitEvent(h->stream, h->ev_copied[b], 0));
    rc = ingest_device(h, h->d_stage[b], m);
    if (rc != ALZ_OK) return rc;
    CK(cudaEventRecord(h->ev_consumed[b], h->stream));
    done += m;
  }
  if (direct) {  // the caller may reuse its pinned buffer once we return
    CK(cudaStreamSynchronize(h->copy_stream));
  }
  return ALZ_OK;
}

extern "C" int alz_submit_l7_raw(alz_handle* h, const void* raw, size_t n) {
  if (!h || (!raw && n)) return ALZ_E_INVAL;
  CK(cudaSetDevice(h->device));
  int rc = ensure_stage(h);
  if (rc != ALZ_OK) return rc;
  // raw chunk: as many samples as fit the byte size of one compact staging buffer
  const size_t chunk = std::max<size_t>(1, ((size_t)h->cfg.max_batch * sizeof(alz_l7_rec)) / ALZ_BPF_L7_EVENT_SIZE);
  if (!h->d_raw_stage) {
    CK(cudaMalloc(&h->d_raw_stage, chunk * ALZ_BPF_L7_EVENT_SIZE));
    CK(cudaMallocHost(&h->h_raw_stage, chunk * ALZ_BPF_L7_EVENT_SIZE));
  }
  const bool direct = is_lib_pinned(raw, n * ALZ_BPF_L7_EVENT_SIZE);
  const uint8_t* p = (const uint8_t*)raw;
  size_t done = 0;
  while (done < n) {
    const size_t m = std::min(chunk, n - done);
    const void* src = p + done * ALZ_BPF_L7_EVENT_SIZE;
    // single raw staging buffer: wait until the compaction kernel has consumed it
    CK(cudaStreamSynchronize(h->stream));
    if (!direct) { memcpy(h->h_raw_stage, src, m * ALZ_BPF_L7_EVENT_SIZE); src = h->h_raw_stage; }
    CK(cudaMemcpyAsync(h->d_raw_stage, src, m * ALZ_BPF_L7_EVENT_SIZE, cudaMemcpyHostToDevice, h->stream));
    launch_compact_raw(h->d_raw_stage, m, h->d_stage[0], h->sms, h->stream);
    rc = ingest_device(h, h->d_stage[0], m);
    if (rc != ALZ_OK) return rc;
    done += m;
  }
  CK(cudaStreamSynchronize(h->stream));
  return ALZ_OK;
}

// ---- window result ---------------------------------------------------------------------------
static int read_counters(alz_handle* h) {
  CK(cudaMemcpyAsync(h->h_ctr, h->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return ALZ_OK;
}

// fold + sort of the live edge keys (no reset). After it h->n_live edges sit in d_keys[1]/d_rows[1].
static int prepare_flush(alz_handle* h) {
  int rc = alz_internal_fold(h);
  if (rc != ALZ_OK) return rc;
  rc = read_counters(h);
  if (rc != ALZ_OK) return rc;
  h->n_live = h->h_ctr->edge_rows;
  if (h->n_live > h->cfg.max_edges) { h->n_live = h->cfg.max_edges; return ALZ_E_CAPACITY; }
  if (h->n_live) {
    launch_iota(h->d_rows[0], h->n_live, h->sms, h->stream);
    sort_pairs(h->d_sort_tmp, h->sort_tmp_bytes, h->edges.row_key, h->d_keys[1], h->d_rows[0], h->d_rows[1],
               h->n_live, h->stream);
    CK(cudaGetLastError());
  }
  return ALZ_OK;
}

static int finish_flush(alz_handle* h) {
  launch_gather_edges(h->edges, h->d_keys[1], h->d_rows[1], h->n_live, h->d_out, true, h->sms, h->stream);
  CK(cudaGetLastError());
  int rc = clear_dict(h, &h->edges);
  if (rc != ALZ_OK) return rc;
  h->last_n_edges = h->n_live;
  h->windows++;
  return ALZ_OK;
}

extern "C" int alz_window_flush_device(alz_handle* h, const alz_edge_out** dev_edges, size_t* n_out) {
  if (!h || !n_out) return ALZ_E_INVAL;
  CK(cudaSetDevice(h->device));
  int rc = prepare_flush(h);
  *n_out = h->n_live;
  if (rc != ALZ_OK) return rc;
  const bool lost = h->h_ctr->capacity_events != 0;
  rc = alz_internal_merge_ranks(h);  // multi-GPU: canonical merge + all-reduce (alz_comm.cu); no-op at 1 rank
  if (rc == ALZ_E_UNSUPPORTED) rc = finish_flush(h);
  if (rc != ALZ_OK) return rc;
  *n_out = h->last_n_edges;
  if (dev_edges) *dev_edges = h->d_out;
  return lost ? ALZ_E_CAPACITY : ALZ_OK;
}

extern "C" int alz_window_flush(alz_handle* h, alz_edge_out* out, size_t cap, size_t* n_out) {
  if (!h || !n_out || (!out && cap)) return ALZ_E_INVAL;
  CK(cudaSetDevice(h->device));
  if (h->comm_nranks <= 1) {
    // size check first so that a too-small buffer keeps the window intact
    int rc = prepare_flush(h);
    *n_out = h->n_live;
    if (rc != ALZ_OK) return rc;
    if (h->n_live > cap) return ALZ_E_CAPACITY;
    const bool lost = h->h_ctr->capacity_events != 0;
    rc = finish_flush(h);
    if (rc != ALZ_OK) return rc;
    if (h->n_live) CK(cudaMemcpyAsync(out, h->d_out, (size_t)h->n_live * sizeof(alz_edge_out),
                                      cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return lost ? ALZ_E_CAPACITY : ALZ_OK;
  }
  const alz_edge_out* d = nullptr;
  int rc = alz_window_flush_device(h, &d, n_out);
  if (rc != ALZ_OK && rc != ALZ_E_CAPACITY) return rc;
  if (*n_out > cap) return ALZ_E_CAPACITY;
  if (*n_out) CK(cudaMemcpyAsync(out, d, *n_out * sizeof(alz_edge_out), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return rc;
}

extern "C" int alz_get_stats(alz_handle* h, alz_stats* st) {
  if (!h || !st) return ALZ_E_INVAL;
  CK(cudaSetDevice(h->device));
  int rc = read_counters(h);
  if (rc != ALZ_OK) return rc;
  memset(st, 0, sizeof(*st));
  st->events_in = h->events_in;
  st->not_request = h->h_ctr->not_request;
  st->src_unresolved = h->h_ctr->src_unresolved;   // complete once pending pairs are folded
  st->rows_emitted = h->events_in - st->not_request - st->src_unresolved - h->h_ctr->capacity_events;
  st->pairs_live = (uint64_t)h->h_ctr->fwd_rows + h->h_ctr->rev_rows;
  st->edges_live = h->h_ctr->edge_rows;
  st->tcp_events_in = h->tcp_events_in;
  st->tcp_localhost_dropped = h->tcp_localhost_dropped;
  return ALZ_OK;
}

// make pending pairs visible in the edge accumulators (and in src_unresolved)
extern "C" int alz_fold(alz_handle* h) {
  if (!h) return ALZ_E_INVAL;
  CK(cudaSetDevice(h->device));
  return alz_internal_fold(h);
}

extern "C" uint32_t alz_owner_rank(uint32_t saddr, uint32_t nranks) { return owner_rank(saddr, nranks); }

// ---- device memory helpers + synthetic stream (include/alazgpu_synth.h) ---------------------------
extern "C" int alz_dev_alloc(alz_handle* h, size_t bytes, void** out) {
  if (!h || !out) return ALZ_E_INVAL;
  CK(cudaSetDevice(h->device));
  cudaError_t e = cudaMalloc(out, bytes);
  if (e != cudaSuccess) { h->last_err = cudaGetErrorString(e); return ALZ_E_NOMEM; }
  return ALZ_OK;
}
extern "C" int alz_dev_free(alz_handle* h, void* p) {
  if (!h) return ALZ_E_INVAL;
  CK(cudaSetDevice(h->device));
  CK(cudaFree(p));
  return ALZ_OK;
}
extern "C" int alz_memcpy_h2d(alz_handle* h, void* dst, const void* src, size_t bytes) {
  if (!h) return ALZ_E_INVAL;
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return ALZ_OK;
}
extern "C" int alz_memcpy_d2h(alz_handle* h, void* dst, const void* src, size_t bytes) {
  if (!h) return ALZ_E_INVAL;
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return ALZ_OK;
}

struct alz_synth_dev {
  alz_synth_view view;  // device pointers
  void* bufs[6];
};

extern "C" int alz_synth_dev_create(alz_handle* h, const alz_synth_topo* t, alz_synth_dev** out) {
  if (!h || !t || !out) return ALZ_E_INVAL;
  CK(cudaSetDevice(h->device));
  alz_synth_dev* d = new (std::nothrow) alz_synth_dev();
  if (!d) return ALZ_E_NOMEM;
  memset(d, 0, sizeof(*d));
  d->view = t->view;
  const size_t E = t->n_edges;
  const void* src[6] = {t->edge_saddr, t->edge_daddr, t->edge_flags, t->alias_thresh, t->alias_idx, t->lat_q};
  const size_t bytes[6] = {E * 4, E * 4, E, E * 4, E * 4, (ALZ_SYNTH_LATQ + 1) * 8};
  for (int i = 0; i < 6; ++i) {
    CK(cudaMalloc(&d->bufs[i], bytes[i]));
    CK(cudaMemcpy(d->bufs[i], src[i], bytes[i], cudaMemcpyHostToDevice));
  }
  d->view.edge_saddr = (const uint32_t*)d->bufs[0];
  d->view.edge_daddr = (const uint32_t*)d->bufs[1];
  d->view.edge_flags = (const uint8_t*)d->bufs[2];
  d->view.alias_thresh = (const uint32_t*)d->bufs[3];
  d->view.alias_idx = (const uint32_t*)d->bufs[4];
  d->view.lat_q = (const uint64_t*)d->bufs[5];
  *out = d;
  return ALZ_OK;
}
extern "C" int alz_synth_dev_fill(alz_handle* h, alz_synth_dev* d, uint64_t first, uint64_t n, alz_l7_rec* dev_out) {
  if (!h || !d || (!dev_out && n)) return ALZ_E_INVAL;
  CK(cudaSetDevice(h->device));
  launch_synth(d->view, first, n, dev_out, h->sms, h->stream);
  CK(cudaGetLastError());
  return ALZ_OK;
}
// First `want` events of the global stream (indices first, first+1, ...) that rank `rank` of `nranks` owns,
// scanning in chunks; *n_scanned = how many global events were looked at. Order within dev_out is arbitrary.
extern "C" int alz_synth_dev_fill_owned(alz_handle* h, alz_synth_dev* d, uint64_t first, uint32_t nranks, uint32_t rank,
                                        alz_l7_rec* dev_out, uint64_t want, uint64_t* n_written, uint64_t* n_scanned) {
  if (!h || !d || !dev_out || !n_written || !n_scanned || nranks == 0 || rank >= nranks) return ALZ_E_INVAL;
  CK(cudaSetDevice(h->device));
  unsigned long long* d_cnt = nullptr;
  CK(cudaMalloc(&d_cnt, 8));
  CK(cudaMemsetAsync(d_cnt, 0, 8, h->stream));
  const uint64_t chunk = std::max<uint64_t>(1u << 20, want / 4);
  unsigned long long got = 0;
  uint64_t scanned = 0;
  while (got < want && scanned < want * (uint64_t)nranks * 64ull) {
    launch_synth_owned(d->view, first + scanned, chunk, nranks, rank, dev_out, want, d_cnt, h->sms, h->stream);
    scanned += chunk;
    CK(cudaMemcpyAsync(&got, d_cnt, 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  cudaFree(d_cnt);
  *n_written = got < want ? got : want;
  *n_scanned = scanned;
  return ALZ_OK;
}
extern "C" int alz_synth_dev_destroy(alz_handle* h, alz_synth_dev* d) {
  if (!h || !d) return ALZ_E_INVAL;
  cudaSetDevice(h->device);
  for (int i = 0; i < 6; ++i) cudaFree(d->bufs[i]);
  delete d;
  return ALZ_OK;
}
