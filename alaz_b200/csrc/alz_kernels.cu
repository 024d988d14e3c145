// alz_kernels.cu — sm_100a kernels of the aggregation hot path.
//
// Plan: the reference resolves every event through the IP->UID
// tables and then emits one row (aggregator/data.go:827-870, :1244). Grouping
// commutes with that join, so the default plan here reduces the stream per
// socket pair (saddr,daddr,direction) first — one dictionary probe and two
// reductions per event — and joins only the DISTINCT pairs against the tables
// (fold_pairs_kernel). ALZ_CFG_EAGER_JOIN keeps the textbook plan (join every
// event, then reduce) for comparison; both give bit-identical edges.
#include "alz_kernels.cuh"

namespace alz {

// ---------------------------------------------------------------------------
// ingest, default plan: events -> per-socket-pair accumulators
// ---------------------------------------------------------------------------
template <int UNROLL>
__global__ void __launch_bounds__(256) ingest_pairs_kernel(const alz_l7_rec* __restrict__ recs, uint64_t n,
                                                           AccTable fwd, AccTable rev, Counters* ctr) {
  uint32_t not_request = 0, inserted = 0, lost = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < n; i += stride * UNROLL) {
    Rec r[UNROLL];
    bool live[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const uint64_t j = i + (uint64_t)u * stride;
      live[u] = j < n;
      if (live[u]) r[u] = load_rec(recs + j);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (!live[u]) continue;
      const uint32_t proto = rec_protocol(r[u]), mf = rec_mflags(r[u]);
      if (!emits_request_row(proto, mf)) { ++not_request; continue; }
      const uint64_t key = ((uint64_t)rec_saddr(r[u]) << 32) | rec_daddr(r[u]);
      const AccTable& t = is_reversed(proto, mf) ? rev : fwd;
      const uint32_t slot = find_or_insert(t, key, &inserted);
      if (slot == 0xFFFFFFFFu) { ++lost; continue; }
      const uint64_t dur = rec_duration(r[u]);
      atomicAdd(&t.hist[(size_t)slot * ALZ_NB + latency_bucket(dur)], 1u);
      atomicAdd((unsigned long long*)&t.lat_sum[slot], (unsigned long long)dur);
      if (is_5xx(proto, rec_status(r[u]))) atomicAdd((unsigned long long*)&t.err5xx[slot], 1ull);
    }
  }
  // warp-aggregate the rare counters
  for (int o = 16; o > 0; o >>= 1) {
    not_request += __shfl_xor_sync(0xFFFFFFFFu, not_request, o);
    inserted += __shfl_xor_sync(0xFFFFFFFFu, inserted, o);
    lost += __shfl_xor_sync(0xFFFFFFFFu, lost, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (not_request) atomicAdd(&ctr->not_request, (unsigned long long)not_request);
    if (inserted) atomicAdd(&ctr->pairs_inserted, (unsigned long long)inserted);
    if (lost) atomicAdd(&ctr->capacity_events, (unsigned long long)lost);
  }
}

// ---------------------------------------------------------------------------
// ingest, eager plan: join every event, then reduce per edge
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ingest_eager_kernel(const alz_l7_rec* __restrict__ recs, uint64_t n,
                                                           const EpEntry* __restrict__ ep, uint32_t ep_mask,
                                                           AccTable edges, Counters* ctr) {
  uint32_t not_request = 0, unresolved = 0, inserted = 0, lost = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const Rec r = load_rec(recs + i);
    const uint32_t proto = rec_protocol(r), mf = rec_mflags(r);
    if (!emits_request_row(proto, mf)) { ++not_request; continue; }
    uint64_t ekey;
    if (!resolve_edge(ep, ep_mask, rec_saddr(r), rec_daddr(r), is_reversed(proto, mf), &ekey)) {
      ++unresolved; continue;
    }
    const uint32_t slot = find_or_insert(edges, ekey, &inserted);
    if (slot == 0xFFFFFFFFu) { ++lost; continue; }
    const uint64_t dur = rec_duration(r);
    atomicAdd(&edges.hist[(size_t)slot * ALZ_NB + latency_bucket(dur)], 1u);
    atomicAdd((unsigned long long*)&edges.lat_sum[slot], (unsigned long long)dur);
    atomicAdd((unsigned long long*)&edges.count[slot], 1ull);
    if (is_5xx(proto, rec_status(r))) atomicAdd((unsigned long long*)&edges.err5xx[slot], 1ull);
  }
  for (int o = 16; o > 0; o >>= 1) {
    not_request += __shfl_xor_sync(0xFFFFFFFFu, not_request, o);
    unresolved += __shfl_xor_sync(0xFFFFFFFFu, unresolved, o);
    inserted += __shfl_xor_sync(0xFFFFFFFFu, inserted, o);
    lost += __shfl_xor_sync(0xFFFFFFFFu, lost, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (not_request) atomicAdd(&ctr->not_request, (unsigned long long)not_request);
    if (unresolved) atomicAdd(&ctr->src_unresolved, (unsigned long long)unresolved);
    if (inserted) atomicAdd(&ctr->edges_inserted, (unsigned long long)inserted);
    if (lost) atomicAdd(&ctr->capacity_events, (unsigned long long)lost);
  }
}

// ---------------------------------------------------------------------------
// fold: the hash join proper, on distinct socket pairs. One warp per pair row:
// resolve (saddr,daddr) -> edge, add the row into the edge accumulators, and
// return the row to the empty state.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fold_pairs_kernel(AccTable pairs, bool rev, const EpEntry* __restrict__ ep,
                                                         uint32_t ep_mask, AccTable edges, Counters* ctr) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row <= pairs.cap; row += warps_per_grid) {
    uint64_t key = pairs.keys[row];
    uint32_t h0 = pairs.hist[(size_t)row * ALZ_NB + lane];
    uint32_t h1 = pairs.hist[(size_t)row * ALZ_NB + 32u + lane];
    uint64_t cnt = (uint64_t)h0 + h1;
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    if (row == pairs.cap) {
      if (cnt == 0) continue;       // sentinel row unused
      key = kEmptyKey;              // it stands for saddr = daddr = 255.255.255.255
    } else if (key == kEmptyKey) {
      continue;
    }
    uint64_t ekey = 0;
    uint32_t eslot = 0xFFFFFFFFu;
    bool ok = false;
    if (lane == 0) {
      ok = resolve_edge(ep, ep_mask, (uint32_t)(key >> 32), (uint32_t)key, rev, &ekey);
      if (ok) {
        uint32_t ins = 0;
        eslot = find_or_insert(edges, ekey, &ins);
        if (ins) atomicAdd(&ctr->edges_inserted, 1ull);
        if (eslot == 0xFFFFFFFFu) atomicAdd(&ctr->capacity_events, (unsigned long long)cnt);
      } else {
        atomicAdd(&ctr->src_unresolved, (unsigned long long)cnt);
      }
    }
    ok = __shfl_sync(0xFFFFFFFFu, ok ? 1 : 0, 0) != 0;
    eslot = __shfl_sync(0xFFFFFFFFu, eslot, 0);
    if (ok && eslot != 0xFFFFFFFFu) {
      if (h0) atomicAdd(&edges.hist[(size_t)eslot * ALZ_NB + lane], h0);
      if (h1) atomicAdd(&edges.hist[(size_t)eslot * ALZ_NB + 32u + lane], h1);
      if (lane == 0) {
        atomicAdd((unsigned long long*)&edges.count[eslot], (unsigned long long)cnt);
        atomicAdd((unsigned long long*)&edges.lat_sum[eslot], (unsigned long long)pairs.lat_sum[row]);
        const uint64_t e = pairs.err5xx[row];
        if (e) atomicAdd((unsigned long long*)&edges.err5xx[eslot], (unsigned long long)e);
      }
    }
    // row back to empty
    pairs.hist[(size_t)row * ALZ_NB + lane] = 0u;
    pairs.hist[(size_t)row * ALZ_NB + 32u + lane] = 0u;
    if (lane == 0) {
      if (row != pairs.cap) pairs.keys[row] = kEmptyKey;
      pairs.lat_sum[row] = 0ull;
      pairs.err5xx[row] = 0ull;
    }
  }
}

// ---------------------------------------------------------------------------
// flush: live edges -> (key, row) list; after the sort, gather rows into the
// caller-facing layout and return them to the empty state.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) compact_edges_kernel(AccTable edges, uint64_t* out_keys, uint32_t* out_rows,
                                                            Counters* ctr) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t row = blockIdx.x * blockDim.x + threadIdx.x; row <= edges.cap; row += stride) {
    const uint64_t key = edges.keys[row];
    const bool live = (row == edges.cap) ? (edges.count[row] != 0ull) : (key != kEmptyKey);
    const uint32_t m = __ballot_sync(__activemask(), live);
    if (!live) continue;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t leader = __ffs(m) - 1u;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(&ctr->n_live, (unsigned long long)__popc(m));
    base = __shfl_sync(m, base, leader);
    const uint32_t pos = (uint32_t)base + __popc(m & ((1u << lane) - 1u));
    out_keys[pos] = key;
    out_rows[pos] = row;
  }
}

__global__ void __launch_bounds__(256) gather_edges_kernel(AccTable edges, const uint64_t* __restrict__ keys,
                                                           const uint32_t* __restrict__ rows, uint32_t n_live,
                                                           alz_edge_out* __restrict__ out, bool reset) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_live; i += warps_per_grid) {
    const uint32_t row = rows[i];
    alz_edge_out* o = &out[i];
    const uint32_t h0 = edges.hist[(size_t)row * ALZ_NB + lane];
    const uint32_t h1 = edges.hist[(size_t)row * ALZ_NB + 32u + lane];
    o->hist[lane] = h0;
    o->hist[32u + lane] = h1;
    if (lane == 0) {
      uint8_t ft, tt; uint32_t f, t;
      unpack_edge_key(keys[i], &ft, &f, &tt, &t);
      o->from_type = ft; o->to_type = tt;
      for (int k = 0; k < 6; ++k) o->_pad[k] = 0;
      o->from = f; o->to = t;
      o->count = edges.count[row];
      o->err5xx = edges.err5xx[row];
      o->lat_sum_ns = edges.lat_sum[row];
    }
    if (reset) {
      edges.hist[(size_t)row * ALZ_NB + lane] = 0u;
      edges.hist[(size_t)row * ALZ_NB + 32u + lane] = 0u;
      if (lane == 0) {
        if (row != edges.cap) edges.keys[row] = kEmptyKey;
        edges.count[row] = 0ull; edges.err5xx[row] = 0ull; edges.lat_sum[row] = 0ull;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// raw perf samples (1096-B struct l7_event, ebpf/l7_req/l7.go:345-369) -> 32-B
// records. One thread per sample; the 72 useful bytes sit in the first 36 and
// the last 36 bytes of each sample.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) compact_raw_kernel(const uint8_t* __restrict__ raw, uint64_t n,
                                                          alz_l7_rec* __restrict__ out) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint8_t* p = raw + i * ALZ_BPF_L7_EVENT_SIZE;  // 1096 = 8 * 137: 8-byte aligned
    const uint64_t write_time = *reinterpret_cast<const uint64_t*>(p + 8);
    const uint32_t status = *reinterpret_cast<const uint32_t*>(p + 20);
    const uint64_t duration = *reinterpret_cast<const uint64_t*>(p + 24);
    const uint32_t pm = *reinterpret_cast<const uint32_t*>(p + 32);      // protocol, method, pad
    const uint32_t fl = *reinterpret_cast<const uint32_t*>(p + 1064);    // read_complete, failed, is_tls
    const uint32_t saddr = *reinterpret_cast<const uint32_t*>(p + 1076);
    const uint32_t sport = *reinterpret_cast<const uint16_t*>(p + 1080);
    const uint32_t daddr = *reinterpret_cast<const uint32_t*>(p + 1084);
    const uint32_t dport = *reinterpret_cast<const uint16_t*>(p + 1088);
    const uint32_t protocol = pm & 0xFFu, method = (pm >> 8) & 0xFFu;
    const uint32_t is_tls = (fl >> 16) & 0xFFu;
    uint4 lo, hi;
    lo.x = saddr; lo.y = daddr; lo.z = sport | (dport << 16);
    lo.w = (status > 65535u ? 65535u : status) | (protocol << 16) |
           (((method & ALZ_MF_METHOD_MASK) | (is_tls ? ALZ_MF_TLS : 0u)) << 24);
    hi.x = (uint32_t)duration; hi.y = (uint32_t)(duration >> 32);
    hi.z = (uint32_t)write_time; hi.w = (uint32_t)(write_time >> 32);
    uint4* o = reinterpret_cast<uint4*>(out + i);
    o[0] = lo; o[1] = hi;
  }
}

// synthetic stream on the device (bench/test support; alaz_b200/synth/alz_synth.h)
__global__ void __launch_bounds__(256) synth_kernel(alz_synth_view v, uint64_t first, uint64_t n,
                                                    alz_l7_rec* __restrict__ out) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    alz_l7_rec r;
    alz_synth_event(&v, first + i, &r);
    out[i] = r;
  }
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
static inline unsigned grid_for(int sms, int per_sm) { return (unsigned)(sms * per_sm); }

void launch_ingest_pairs(const alz_l7_rec* recs, uint64_t n, const AccTable& fwd, const AccTable& rev,
                         Counters* ctr, int sms, cudaStream_t s) {
  if (n == 0) return;
  ingest_pairs_kernel<4><<<grid_for(sms, 8), 256, 0, s>>>(recs, n, fwd, rev, ctr);
}
void launch_ingest_eager(const alz_l7_rec* recs, uint64_t n, const EpEntry* ep, uint32_t ep_mask,
                         const AccTable& edges, Counters* ctr, int sms, cudaStream_t s) {
  if (n == 0) return;
  ingest_eager_kernel<<<grid_for(sms, 8), 256, 0, s>>>(recs, n, ep, ep_mask, edges, ctr);
}
void launch_fold_pairs(const AccTable& pairs, bool rev, const EpEntry* ep, uint32_t ep_mask,
                       const AccTable& edges, Counters* ctr, int sms, cudaStream_t s) {
  fold_pairs_kernel<<<grid_for(sms, 8), 256, 0, s>>>(pairs, rev, ep, ep_mask, edges, ctr);
}
void launch_compact_edges(const AccTable& edges, uint64_t* keys, uint32_t* rows, Counters* ctr, int sms,
                          cudaStream_t s) {
  compact_edges_kernel<<<grid_for(sms, 8), 256, 0, s>>>(edges, keys, rows, ctr);
}
void launch_gather_edges(const AccTable& edges, const uint64_t* keys, const uint32_t* rows, uint32_t n_live,
                         alz_edge_out* out, bool reset, int sms, cudaStream_t s) {
  if (n_live == 0) return;
  gather_edges_kernel<<<grid_for(sms, 8), 256, 0, s>>>(edges, keys, rows, n_live, out, reset);
}
void launch_compact_raw(const uint8_t* raw, uint64_t n, alz_l7_rec* out, int sms, cudaStream_t s) {
  if (n == 0) return;
  compact_raw_kernel<<<grid_for(sms, 8), 256, 0, s>>>(raw, n, out);
}
void launch_synth(const alz_synth_view& v, uint64_t first, uint64_t n, alz_l7_rec* out, int sms, cudaStream_t s) {
  if (n == 0) return;
  synth_kernel<<<grid_for(sms, 8), 256, 0, s>>>(v, first, n, out);
}

}  // namespace alz
