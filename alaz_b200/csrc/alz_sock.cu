// alz_sock.cu — tcp_state sink and the temporal socket join.
//
// Replaces, for events whose in-event 5-tuple is empty (get_sock miss, ebpf/c/l7.c:313-314):
//   processTcpConnect      aggregator/data.go:404-506   -> alz_submit_tcp   (host, exact order semantics)
//   SocketLine.AddValue    aggregator/sock_num_line.go:62-80, 311-322
//   SocketLine.GetValue    aggregator/sock_num_line.go:82-158 -> alz_sock_lookup (device, one thread per query)
//
// tcp_state events are two per connection and AddValue's dedupe depends on arrival order
// ("equal to the LAST element"), so the timelines are maintained on the host exactly as the
// reference does, sequentially. The join itself — (pid, fd, timestamp) -> SockInfo for a batch of
// L7 events — is the data-parallel part: the timelines are flattened into one time-sorted array
// with an open-addressed (pid,fd) index in HBM and each query does the reference's binary search
// and its open/closed-gap rules on the device.
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "alz_handle.h"

using namespace alz;

namespace {

struct SockRec {       // one TimestampedSocket (sock_num_line.go:23-27); LastMatch is GC-only state
  uint64_t ts;
  uint32_t open;       // SockInfo != nil
  uint32_t saddr, daddr;
  uint16_t sport, dport;
  uint32_t pad;
};
static_assert(sizeof(SockRec) == 32, "SockRec layout");

struct LineEnt {       // index entry: (pid, fd) -> segment of the flat array
  uint64_t fd;
  uint32_t pid;
  uint32_t used;
  uint32_t off, len;
  uint64_t pad;
};
static_assert(sizeof(LineEnt) == 32, "LineEnt layout");

struct LineKey {
  uint32_t pid;
  uint64_t fd;
  bool operator==(const LineKey& o) const { return pid == o.pid && fd == o.fd; }
};
struct LineKeyHash {
  size_t operator()(const LineKey& k) const { return (size_t)hash64(((uint64_t)k.pid << 40) ^ k.fd ^ 0x9E3779B97F4A7C15ull); }
};

__host__ __device__ inline uint32_t line_slot(uint32_t pid, uint64_t fd, uint32_t mask) {
  return (uint32_t)hash64(((uint64_t)pid << 40) ^ fd ^ 0x9E3779B97F4A7C15ull) & mask;
}

constexpr uint64_t kOneMinuteNs = 60ull * 1000000000ull;
constexpr uint32_t kLocalhost = 0x7F000001u;   // "127.0.0.1" (data.go:409, :455)

// SocketLine.GetValue on the flattened timeline
__global__ void __launch_bounds__(256) sock_lookup_kernel(const LineEnt* __restrict__ index, uint32_t mask,
                                                          const SockRec* __restrict__ recs,
                                                          const alz_sock_query* __restrict__ q, uint32_t n,
                                                          alz_sock_result* __restrict__ out) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    alz_sock_result r;
    r.found = 0; r.saddr = 0; r.daddr = 0; r.sport = 0; r.dport = 0;
    const uint32_t pid = q[i].pid;
    const uint64_t fd = q[i].fd, ts = q[i].timestamp_ns;
    // findRelatedSocket: SocketMaps[pid].M[fd] (data.go:1407-1429)
    uint32_t slot = line_slot(pid, fd, mask);
    const SockRec* v = nullptr;
    uint32_t len = 0;
    for (;;) {
      const LineEnt e = index[slot];
      if (!e.used) break;
      if (e.pid == pid && e.fd == fd) { v = recs + e.off; len = e.len; break; }
      slot = (slot + 1u) & mask;
    }
    const SockRec* hit = nullptr;
    if (v != nullptr && len != 0u) {                                    // :86-88 empty line -> error
      uint32_t lo = 0, hi = len;                                         // sort.Search(!(Timestamp < ts)) :90-92
      while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (!(v[mid].ts < ts)) hi = mid; else lo = mid + 1; }
      const uint32_t idx = lo;
      if (idx == len) {                                                  // after the last entry, :94-105
        if (!v[len - 1].open) {
          if (idx >= 2u && v[idx - 2].open && (ts - v[idx - 2].ts) < kOneMinuteNs) hit = &v[idx - 2];
        } else hit = &v[len - 1];
      } else if (idx == 0u) {                                            // before the first entry, :107-119
        if (v[0].open) hit = &v[0];
      } else if (!v[idx - 1].open) {                                     // matched a close, :123-153
        if (idx >= 2u && v[idx - 2].open && v[idx].open && v[idx - 2].daddr == v[idx].daddr &&
            v[idx - 2].dport == v[idx].dport)
          hit = (ts - v[idx - 2].ts < v[idx].ts - ts) ? &v[idx - 2] : &v[idx];
      } else hit = &v[idx - 1];                                          // :155-157
    }
    if (hit != nullptr) {
      r.found = 1; r.saddr = hit->saddr; r.daddr = hit->daddr; r.sport = hit->sport; r.dport = hit->dport;
    }
    out[i] = r;
  }
}

}  // namespace

struct alz_sock_state {
  std::unordered_map<LineKey, std::vector<SockRec>, LineKeyHash> lines;
  bool dirty = true;
  LineEnt* d_index = nullptr;
  SockRec* d_recs = nullptr;
  uint32_t index_cap = 0;
  size_t recs_cap = 0;
  alz_sock_query* d_q = nullptr;
  alz_sock_result* d_out = nullptr;
  size_t q_cap = 0;
};

#define CK(expr)                                                                       \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      h->last_err = std::string(#expr) + ": " + cudaGetErrorString(_e);                \
      return ALZ_E_CUDA;                                                               \
    }                                                                                  \
  } while (0)

// SocketLine.AddValue: skip when equal to the last element's open socket, else sorted insert
static void add_value(std::vector<SockRec>& v, uint64_t ts, const alz_tcp_rec* si) {
  if (!v.empty() && si != nullptr) {
    const SockRec& last = v.back();
    if (last.open && last.saddr == si->saddr && last.sport == si->sport && last.daddr == si->daddr &&
        last.dport == si->dport)
      return;                                                            // sock_num_line.go:70-78
  }
  size_t lo = 0, hi = v.size();                                          // insertIntoSortedSlice :311-322
  while (lo < hi) { const size_t mid = (lo + hi) / 2; if (v[mid].ts >= ts) hi = mid; else lo = mid + 1; }
  SockRec r;
  memset(&r, 0, sizeof r);
  r.ts = ts;
  if (si != nullptr) { r.open = 1; r.saddr = si->saddr; r.daddr = si->daddr; r.sport = si->sport; r.dport = si->dport; }
  v.insert(v.begin() + (ptrdiff_t)lo, r);
}

extern "C" int alz_submit_tcp(alz_handle* h, const alz_tcp_rec* recs, size_t n) {
  if (!h || (!recs && n)) return ALZ_E_INVAL;
  std::lock_guard<std::mutex> g(h->mu);
  if (!h->sock) h->sock = new alz_sock_state();
  alz_sock_state* s = h->sock;
  for (size_t i = 0; i < n; ++i) {
    const alz_tcp_rec& d = recs[i];
    h->tcp_events_in++;
    if (d.type != 1u && d.type != 5u) continue;                          // only ESTABLISHED / CLOSED are handled
    if (d.saddr == kLocalhost || d.daddr == kLocalhost) { h->tcp_localhost_dropped++; continue; }
    const LineKey k{d.pid, d.fd};
    if (d.type == 1u) {                                                  // EVENT_TCP_ESTABLISHED, data.go:406-449
      add_value(s->lines[k], d.timestamp_ns, &d);                        // line created on first use (:417-437)
      s->dirty = true;
    } else {                                                             // EVENT_TCP_CLOSED, :450-478
      auto it = s->lines.find(k);
      if (it == s->lines.end()) continue;                                // no line: ignored (:471-473)
      add_value(it->second, d.timestamp_ns, nullptr);
      s->dirty = true;
    }
  }
  return ALZ_OK;
}

static uint32_t pow2_at_least(size_t x) { uint32_t p = 16; while (p < x) p <<= 1; return p; }

static int upload_lines(alz_handle* h) {
  alz_sock_state* s = h->sock;
  size_t total = 0;
  for (auto& kv : s->lines) total += kv.second.size();
  const uint32_t cap = pow2_at_least(2 * s->lines.size() + 1);
  std::vector<LineEnt> index(cap);
  memset(index.data(), 0, cap * sizeof(LineEnt));
  std::vector<SockRec> flat;
  flat.reserve(total);
  for (auto& kv : s->lines) {
    uint32_t slot = line_slot(kv.first.pid, kv.first.fd, cap - 1);
    while (index[slot].used) slot = (slot + 1) & (cap - 1);
    index[slot].used = 1; index[slot].pid = kv.first.pid; index[slot].fd = kv.first.fd;
    index[slot].off = (uint32_t)flat.size(); index[slot].len = (uint32_t)kv.second.size();
    flat.insert(flat.end(), kv.second.begin(), kv.second.end());
  }
  if (cap > s->index_cap) { cudaFree(s->d_index); CK(cudaMalloc(&s->d_index, (size_t)cap * sizeof(LineEnt))); }
  s->index_cap = cap;
  if (flat.size() > s->recs_cap) {
    cudaFree(s->d_recs);
    s->recs_cap = flat.size() * 2 + 16;
    CK(cudaMalloc(&s->d_recs, s->recs_cap * sizeof(SockRec)));
  }
  CK(cudaMemcpyAsync(s->d_index, index.data(), (size_t)cap * sizeof(LineEnt), cudaMemcpyHostToDevice, h->stream));
  if (!flat.empty())
    CK(cudaMemcpyAsync(s->d_recs, flat.data(), flat.size() * sizeof(SockRec), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));   // host vectors die here
  s->dirty = false;
  return ALZ_OK;
}

extern "C" int alz_sock_lookup(alz_handle* h, const alz_sock_query* q, size_t n, alz_sock_result* out) {
  if (!h || (!q && n) || (!out && n)) return ALZ_E_INVAL;
  if (n == 0) return ALZ_OK;
  std::lock_guard<std::mutex> g(h->mu);
  CK(cudaSetDevice(h->device));
  if (!h->sock) h->sock = new alz_sock_state();
  alz_sock_state* s = h->sock;
  if (s->dirty) { int rc = upload_lines(h); if (rc != ALZ_OK) return rc; }
  if (n > s->q_cap) {
    cudaFree(s->d_q); cudaFree(s->d_out);
    s->q_cap = n;
    CK(cudaMalloc(&s->d_q, n * sizeof(alz_sock_query)));
    CK(cudaMalloc(&s->d_out, n * sizeof(alz_sock_result)));
  }
  CK(cudaMemcpyAsync(s->d_q, q, n * sizeof(alz_sock_query), cudaMemcpyHostToDevice, h->stream));
  sock_lookup_kernel<<<(unsigned)h->sms * 4, 256, 0, h->stream>>>(s->d_index, s->index_cap - 1, s->d_recs, s->d_q,
                                                                (uint32_t)n, s->d_out);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, s->d_out, n * sizeof(alz_sock_result), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return ALZ_OK;
}

void alz_internal_free_sock(alz_handle* h) {
  alz_sock_state* s = h->sock;
  if (!s) return;
  cudaFree(s->d_index); cudaFree(s->d_recs); cudaFree(s->d_q); cudaFree(s->d_out);
  delete s;
  h->sock = nullptr;
}
