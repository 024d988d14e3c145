// alz_kernels.cu — sm_100a kernels of the aggregation hot path.
//
// Plan: the reference resolves every event through the IP->UID
// tables and then emits one row (aggregator/data.go:827-870, :1244). Grouping
// commutes with that join, so the default plan here reduces the stream per
// socket pair (saddr,daddr,direction) first — one dictionary probe and two
// reductions per event — and joins only the DISTINCT pairs against the tables
// (fold_pairs_kernel). ALZ_CFG_EAGER_JOIN keeps the textbook plan (join every
// event, then reduce) for comparison; both give bit-identical edges.
//
// Control flow in the per-event loops is kept warp-convergent on purpose: every
// lane walks the same statements under predicates and meets at __syncwarp()
// between the divergent dictionary probes and the reductions. The first
// version let lanes leave the probe loops on their own paths and the
// reductions then issued with ~2 active lanes per instruction
// (profiles/r1_v2_ingest_ncu.txt).
#include <algorithm>

#include "alz_kernels.cuh"

namespace alz {

// per-event fields every plan needs
struct Ev {
  uint64_t key;   // make_pair_key(saddr, daddr)
  uint64_t dur;
  uint32_t bucket;
  bool act;       // the reference would hand a row to PersistRequest (before resolve)
  bool rev;       // ReverseDirection applies
  bool err;       // counts as 5xx
  bool hostkey;   // daddr is a Host-header id (ALZ_PROTO_F_HOSTKEY)
};
__device__ __forceinline__ Ev decode(const Rec& r, bool live) {
  Ev e;
  const uint32_t proto = rec_protocol(r), mf = rec_mflags(r);
  e.act = live && emits_request_row(proto, mf);
  e.rev = is_reversed(proto, mf);
  e.hostkey = rec_hostkey(r);
  e.key = make_pair_key(rec_saddr(r), rec_daddr(r));
  e.dur = rec_duration(r);
  e.bucket = latency_bucket(e.dur);
  e.err = is_5xx(proto, rec_status(r));
  return e;
}

__device__ __forceinline__ void global_accumulate(const AccTable& t, uint32_t row, uint32_t bucket, uint64_t dur,
                                                  bool err) {
  atomicAdd(&t.hist[(size_t)row * ALZ_NB + bucket], 1u);
  atomicAdd((unsigned long long*)&t.lat_sum[row], (unsigned long long)dur);
  if (err) atomicAdd((unsigned long long*)&t.err5xx[row], 1ull);
}

// the same into a pair row (sectored layout, alz_device.cuh); v1 plan: one lane per event, three instructions
__device__ __forceinline__ void pair_accumulate(const AccTable& t, uint32_t row, uint32_t bucket, uint64_t dur, bool err) {
  uint64_t* sc = pair_sect(t, row, bucket);
  atomicAdd(reinterpret_cast<uint32_t*>(sc) + (bucket & 3u), 1u);
  atomicAdd((unsigned long long*)(sc + 2), (unsigned long long)dur);
  if (err) atomicAdd((unsigned long long*)(sc + 3), 1ull);
}

__device__ __forceinline__ void flush_thread_counters(Counters* ctr, uint32_t not_request, uint32_t unresolved,
                                                      uint32_t lost) {
  for (int o = 16; o > 0; o >>= 1) {
    not_request += __shfl_xor_sync(0xFFFFFFFFu, not_request, o);
    unresolved += __shfl_xor_sync(0xFFFFFFFFu, unresolved, o);
    lost += __shfl_xor_sync(0xFFFFFFFFu, lost, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (not_request) atomicAdd(&ctr->not_request, (unsigned long long)not_request);
    if (unresolved) atomicAdd(&ctr->src_unresolved, (unsigned long long)unresolved);
    if (lost) atomicAdd(&ctr->capacity_events, (unsigned long long)lost);
  }
}

// ---------------------------------------------------------------------------
// ingest v1 (ALZ_CFG_NO_SMEM_CACHE): global reductions only
// ---------------------------------------------------------------------------
template <int UNROLL>
__global__ void __launch_bounds__(256) ingest_pairs_kernel(const alz_l7_rec* __restrict__ recs, uint64_t n,
                                                           AccTable pairs, Counters* ctr, const EpEntry* __restrict__ ep,
                                                           uint32_t ep_mask) {
  uint32_t not_request = 0, lost = 0, unresolved = 0;
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride * UNROLL) {
    Rec r[UNROLL];
    bool live[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const uint64_t j = base + (uint64_t)u * stride + lane;
      live[u] = j < n;
      if (live[u]) r[u] = load_rec(recs + j);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const Ev e = decode(r[u], live[u]);
      not_request += (live[u] && !e.act) ? 1u : 0u;
      uint32_t row = kLostRow;
      if (e.act) row = find_or_insert_pair(pairs, e.key, e.hostkey ? kPairHost : e.rev ? kPairRev : kPairFwd, ep, ep_mask);
      __syncwarp();
      if (e.act) {
        if (row == kDropRow) ++unresolved;
        else if (row >= kLostRow) ++lost;
        else pair_accumulate(pairs, row, e.bucket, e.dur, e.err);
      }
    }
  }
  flush_thread_counters(ctr, not_request, unresolved, lost);
}

// ---------------------------------------------------------------------------
// ingest, eager plan: join every event, then reduce per edge
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ingest_eager_kernel(const alz_l7_rec* __restrict__ recs, uint64_t n,
                                                           const EpEntry* __restrict__ ep, uint32_t ep_mask,
                                                           AccTable edges, Counters* ctr) {
  uint32_t not_request = 0, unresolved = 0, lost = 0;
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
    const uint64_t j = base + lane;
    const bool live = j < n;
    Rec r = {};
    if (live) r = load_rec(recs + j);
    const Ev e = decode(r, live);
    not_request += (live && !e.act) ? 1u : 0u;
    uint64_t ekey = 0;
    bool ok = false;
    if (e.act) ok = e.hostkey ? resolve_edge_hostkey(ep, ep_mask, pair_saddr(e.key), pair_daddr(e.key), &ekey)
                              : resolve_edge(ep, ep_mask, pair_saddr(e.key), pair_daddr(e.key), e.rev, &ekey);
    __syncwarp();
    unresolved += (e.act && !ok) ? 1u : 0u;
    uint32_t row = kLostRow;
    if (ok) row = find_or_insert(edges, ekey);
    __syncwarp();
    if (ok) {
      if (row >= kLostRow) ++lost;
      else {
        global_accumulate(edges, row, e.bucket, e.dur, e.err);
        atomicAdd((unsigned long long*)&edges.count[row], 1ull);
      }
    }
  }
  flush_thread_counters(ctr, not_request, unresolved, lost);
}

// ---------------------------------------------------------------------------
// fold: the hash join proper, on distinct socket pairs. One warp per pair row:
// resolve (saddr,daddr) -> edge, add the row into the edge accumulators and
// zero it. The caller clears the pair dictionary afterwards.
// ---------------------------------------------------------------------------
// pass 1, a thread per pair row: resolve and find the edge row. All the dependent
// table/dictionary probes of a pair sit in one thread, 32 pairs per warp in flight
// (the one-warp-per-row version spent its time in lane 0's probe chain).
__global__ void __launch_bounds__(256) fold_resolve_kernel(AccTable pairs, const EpEntry* __restrict__ ep,
                                                           uint32_t ep_mask, AccTable edges, HotState* hot) {
  // the hot-list bookkeeping of this fold starts from zero (fold_pairs_kernel fills the bins, hot_emit
  // the list; both run after this kernel, the ingest launches that read the list ran before it)
  if (blockIdx.x == 0 && hot != nullptr) {
    if (threadIdx.x < 128) hot->bins[threadIdx.x] = 0u;
    if (threadIdx.x == 0) { hot->n_a = 0u; hot->n_b = 0u; hot->n_s = 0u; }
  }
  const uint32_t n_rows = min(*pairs.n_rows, pairs.max_rows);
  const uint32_t stride = gridDim.x * blockDim.x;
  // rows [0, n_rows) plus the sentinel rows (index n_rows + kind stands for row max_rows + kind)
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rows + kPairKinds; i += stride) {
    const bool sentinel = i >= n_rows;
    const uint32_t row = sentinel ? pairs.max_rows + (i - n_rows) : i;
    const uint64_t key = sentinel ? kEmptyKey : pairs.row_key[row];
    const uint32_t kind = sentinel ? (i - n_rows) : pairs.row_kind[row];
    if (sentinel) {                                            // the sentinel rows exist even when unused
      uint64_t any = 0;
      for (uint32_t g = 0; g < kSectPerRow; ++g) { const uint64_t* sc = pairs.sect + ((size_t)row * kSectPerRow + g) * 4u; any |= sc[0] | sc[1]; }
      if (any == 0ull) { pairs.row_aux[row] = kDropRow; continue; }
    }
    uint64_t ekey = 0;
    uint32_t erow = kDropRow;                                  // source is not a pod: data.go:829-832
    const bool ok = kind == kPairHost ? resolve_edge_hostkey(ep, ep_mask, pair_saddr(key), pair_daddr(key), &ekey)
                                      : resolve_edge(ep, ep_mask, pair_saddr(key), pair_daddr(key), kind == kPairRev, &ekey);
    if (ok) erow = find_or_insert(edges, ekey);
    pairs.row_aux[row] = erow;
  }
}

// pass 2, eight lanes per pair row (four rows per warp in flight): add the row into its edge row and
// zero it. Each lane owns 8 histogram cells = two 16-byte loads; the one-warp-per-row version was bound
// by its chain of dependent L2 round trips (profiles/r1_final_launches.csv: 0.15 ms per launch).
__global__ void __launch_bounds__(256) fold_pairs_kernel(AccTable pairs, AccTable edges, Counters* ctr,
                                                         uint32_t* __restrict__ hot_bins) {
  const uint32_t lane = threadIdx.x & 31u, sl = lane & 7u;
  const uint32_t groups_per_grid = (gridDim.x * blockDim.x) >> 3;
  const uint32_t n_rows = min(*pairs.n_rows, pairs.max_rows);
  const uint32_t n_iter = (n_rows + kPairKinds + groups_per_grid - 1u) / groups_per_grid;   // same trip count for every lane
  uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  // rows [0, n_rows) plus the sentinel rows (index n_rows + kind stands for row max_rows + kind)
  for (uint32_t it = 0; it < n_iter; ++it, i += groups_per_grid) {
    const bool valid = i < n_rows + kPairKinds;
    const uint32_t row = !valid ? 0u : (i >= n_rows) ? pairs.max_rows + (i - n_rows) : i;
    // the lane's two sectors = 64 contiguous bytes: cells 8sl..8sl+3 | lat, 5xx partials | cells 8sl+4..8sl+7 | partials
    uint4* cells = reinterpret_cast<uint4*>(pairs.sect + ((size_t)row * kSectPerRow + sl * 2u) * 4u);
    uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a, pa = a, pb = a;
    if (valid) { a = cells[0]; pa = cells[1]; b = cells[2]; pb = cells[3]; }
    uint64_t lat = ((((uint64_t)pa.y << 32) | pa.x) + (((uint64_t)pb.y << 32) | pb.x));
    uint64_t e5 = ((((uint64_t)pa.w << 32) | pa.z) + (((uint64_t)pb.w << 32) | pb.z));
    uint64_t cnt = (uint64_t)a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    for (int o = 1; o < 8; o <<= 1) {
      lat += __shfl_xor_sync(0xFFFFFFFFu, lat, o);
      e5 += __shfl_xor_sync(0xFFFFFFFFu, e5, o);
    }
    cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, 1);
    cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, 2);
    cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, 4);
    if (pairs.row_base != nullptr) {   // every lane of the warp takes part in the shuffles (uniform branch)
      // the 16-bucket window (aligned to 4 buckets) that holds most of the row's events: group sums of 4 cells,
      // two candidates per lane, arg-max over the row's eight lanes
      const uint64_t g0 = (uint64_t)a.x + a.y + a.z + a.w, g1 = (uint64_t)b.x + b.y + b.z + b.w;   // groups 2sl, 2sl+1
      const uint32_t seg = lane & ~7u;
      const uint64_t n0 = __shfl_sync(0xFFFFFFFFu, g0, seg | ((sl + 1u) & 7u)), n1 = __shfl_sync(0xFFFFFFFFu, g1, seg | ((sl + 1u) & 7u));
      const uint64_t m0 = __shfl_sync(0xFFFFFFFFu, g0, seg | ((sl + 2u) & 7u));
      // candidate j covers groups j..j+3 (j <= 12); lanes past the end contribute nothing
      uint64_t c0 = sl <= 6u ? g0 + g1 + n0 + n1 : 0ull;                    // j = 2 sl       (sl = 6 -> j = 12)
      uint64_t c1 = sl <= 5u ? g1 + n0 + n1 + m0 : 0ull;                    // j = 2 sl + 1   (sl = 5 -> j = 11)
      uint64_t best = c0 >= c1 ? (c0 << 4) | (2u * sl) : (c1 << 4) | (2u * sl + 1u);   // counts < 2^37: room for 4 index bits
      for (int o = 1; o < 8; o <<= 1) { const uint64_t x = __shfl_xor_sync(0xFFFFFFFFu, best, o); best = x > best ? x : best; }
      if (sl == 0 && valid && row < pairs.max_rows) pairs.row_base[row] = (uint8_t)(best & 15u);
    }
    if (!valid || cnt == 0) continue;  // unused sentinel row (allocated rows always hold >= 1 event)
    if (sl == 0 && row < pairs.max_rows && pairs.row_cnt != nullptr) {    // feedback for the next ingest
      const uint32_t c32 = cnt > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)cnt;
      pairs.row_cnt[row] = c32;
      if (hot_bins != nullptr && pairs.row_kind[row] == kPairFwd) atomicAdd(&hot_bins[count_bin(c32)], 1u);
    }
    const uint32_t erow = pairs.row_aux[row];
    if (erow < kDropRow) {
      uint32_t* dst = edges.hist + (size_t)erow * ALZ_NB + sl * 8u;
      if (a.x) atomicAdd(dst + 0, a.x);
      if (a.y) atomicAdd(dst + 1, a.y);
      if (a.z) atomicAdd(dst + 2, a.z);
      if (a.w) atomicAdd(dst + 3, a.w);
      if (b.x) atomicAdd(dst + 4, b.x);
      if (b.y) atomicAdd(dst + 5, b.y);
      if (b.z) atomicAdd(dst + 6, b.z);
      if (b.w) atomicAdd(dst + 7, b.w);
      if (sl == 0) {
        atomicAdd((unsigned long long*)&edges.count[erow], (unsigned long long)cnt);
        atomicAdd((unsigned long long*)&edges.lat_sum[erow], (unsigned long long)lat);
        if (e5) atomicAdd((unsigned long long*)&edges.err5xx[erow], (unsigned long long)e5);
      }
    } else if (sl == 0) {
      if (erow == kDropRow) atomicAdd(&ctr->src_unresolved, (unsigned long long)cnt);
      else atomicAdd(&ctr->fold_lost_events, (unsigned long long)cnt);
    }
    // zero what was not zero: a pair's latencies fall into a few of its 16 sectors, the rest was never written
    if (a.x | a.y | a.z | a.w | pa.x | pa.y | pa.z | pa.w) { cells[0] = make_uint4(0u, 0u, 0u, 0u); cells[1] = make_uint4(0u, 0u, 0u, 0u); }
    if (b.x | b.y | b.z | b.w | pb.x | pb.y | pb.z | pb.w) { cells[2] = make_uint4(0u, 0u, 0u, 0u); cells[3] = make_uint4(0u, 0u, 0u, 0u); }
  }
}

// ---------------------------------------------------------------------------
// flush: rows [0, n) of the edge table are the live edges. iota -> sort by key
// (alz_sort.cu) -> gather into the caller-facing layout, zeroing the rows.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) iota_kernel(uint32_t* out, uint32_t n) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = i;
}

__global__ void __launch_bounds__(256) gather_edges_kernel(AccTable edges, const uint64_t* __restrict__ keys,
                                                           const uint32_t* __restrict__ rows, uint32_t n_live,
                                                           alz_edge_out* __restrict__ out, bool reset) {
  // eight lanes per edge, four edges per warp in flight; a lane moves 8 histogram cells (2 x 16 bytes)
  const uint32_t sl = threadIdx.x & 7u;
  const uint32_t groups_per_grid = (gridDim.x * blockDim.x) >> 3;
  for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; i < n_live; i += groups_per_grid) {
    const uint32_t row = rows[i];
    alz_edge_out* o = &out[i];
    uint4* cells = reinterpret_cast<uint4*>(edges.hist + (size_t)row * ALZ_NB + sl * 8u);
    uint4* dst = reinterpret_cast<uint4*>(&o->hist[sl * 8u]);   // hist sits at byte 40 of a 296-byte row: 8-byte aligned only
    const uint4 a = cells[0], b = cells[1];
    uint2* d2 = reinterpret_cast<uint2*>(dst);
    d2[0] = make_uint2(a.x, a.y); d2[1] = make_uint2(a.z, a.w);
    d2[2] = make_uint2(b.x, b.y); d2[3] = make_uint2(b.z, b.w);
    if (sl == 0) {
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
      cells[0] = make_uint4(0u, 0u, 0u, 0u);
      cells[1] = make_uint4(0u, 0u, 0u, 0u);
      if (sl == 0) { edges.count[row] = 0ull; edges.err5xx[row] = 0ull; edges.lat_sum[row] = 0ull; }
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

// events first..first+n of the global stream, keeping only those owned by `rank`
// (alz_owner_rank: hash of saddr). Appends in arbitrary order; stops writing at cap.
__global__ void __launch_bounds__(256) synth_owned_kernel(alz_synth_view v, uint64_t first, uint64_t n, uint32_t nranks,
                                                          uint32_t rank, alz_l7_rec* __restrict__ out, uint64_t cap,
                                                          unsigned long long* n_written) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
    const uint64_t i = base + lane;
    alz_l7_rec r;
    bool mine = false;
    if (i < n) {
      alz_synth_event(&v, first + i, &r);
      mine = owner_rank(r.saddr, nranks) == rank;
    }
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, mine);
    if (m == 0u) continue;
    unsigned long long pos = 0;
    if (lane == 0) pos = atomicAdd(n_written, (unsigned long long)__popc(m));
    pos = __shfl_sync(0xFFFFFFFFu, pos, 0) + __popc(m & ((1u << lane) - 1u));
    if (mine && pos < cap) out[pos] = r;
  }
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
static inline unsigned grid_for(int sms, int per_sm) { return (unsigned)(sms * per_sm); }

void launch_ingest_pairs_v1(const alz_l7_rec* recs, uint64_t n, const AccTable& pairs, Counters* ctr,
                            const EpEntry* ep, uint32_t ep_mask, int sms, cudaStream_t s) {
  if (n == 0) return;   // global reductions only (ALZ_CFG_NO_SMEM_CACHE; kept for the ncu comparison)
  ingest_pairs_kernel<4><<<grid_for(sms, 8), 256, 0, s>>>(recs, n, pairs, ctr, ep, ep_mask);
}
void launch_ingest_eager(const alz_l7_rec* recs, uint64_t n, const EpEntry* ep, uint32_t ep_mask,
                         const AccTable& edges, Counters* ctr, int sms, cudaStream_t s) {
  if (n == 0) return;
  ingest_eager_kernel<<<grid_for(sms, 8), 256, 0, s>>>(recs, n, ep, ep_mask, edges, ctr);
}
void launch_fold_resolve(const AccTable& pairs, const EpEntry* ep, uint32_t ep_mask, const AccTable& edges,
                         Counters* ctr, HotState* hot, int sms, cudaStream_t s) {
  (void)ctr;
  fold_resolve_kernel<<<grid_for(sms, 4), 256, 0, s>>>(pairs, ep, ep_mask, edges, hot);
}
void launch_fold_add(const AccTable& pairs, const AccTable& edges, Counters* ctr, HotState* hot, int sms,
                     cudaStream_t s) {
  fold_pairs_kernel<<<grid_for(sms, 8), 256, 0, s>>>(pairs, edges, ctr, hot ? hot->bins : nullptr);
}
// changed slots of the endpoint table (alz_table_commit): patch[i] = {slot, pad x3, entry}
__global__ void __launch_bounds__(256) ep_patch_kernel(EpEntry* __restrict__ tab, const uint4* __restrict__ patch, uint32_t n) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint4 hd = patch[2 * i], e = patch[2 * i + 1];
    *reinterpret_cast<uint4*>(&tab[hd.x]) = e;
  }
}
void launch_ep_patch(EpEntry* tab, const void* patch, uint32_t n, int sms, cudaStream_t s) {
  if (n == 0) return;
  const unsigned grid = (unsigned)std::min<uint32_t>((n + 255u) / 256u, (uint32_t)sms * 4u);
  ep_patch_kernel<<<grid, 256, 0, s>>>(tab, (const uint4*)patch, n);
}
void launch_iota(uint32_t* out, uint32_t n, int sms, cudaStream_t s) {
  if (n == 0) return;
  iota_kernel<<<grid_for(sms, 4), 256, 0, s>>>(out, n);
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
void launch_synth_owned(const alz_synth_view& v, uint64_t first, uint64_t n, uint32_t nranks, uint32_t rank,
                        alz_l7_rec* out, uint64_t cap, unsigned long long* n_written, int sms, cudaStream_t s) {
  if (n == 0) return;
  synth_owned_kernel<<<grid_for(sms, 8), 256, 0, s>>>(v, first, n, nranks, rank, out, cap, n_written);
}

}  // namespace alz
