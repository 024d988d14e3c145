// alz_ingest.cu — the dominant kernel: l7 events -> per-socket-pair accumulators
// (DESIGN.md §3 step 1, §5). One persistent CTA of 1024 threads per SM.
//
// Each CTA keeps a private table of hot socket pairs in shared memory (the stream
// is Zipf-skewed; without it the hottest pairs serialise in one L2 slice):
//   * 4-way buckets with 16-bit tags: a lookup is one 8-byte load of the four tags and one
//     key load for the lanes whose tag matched (two random LDS.128 per lane made the shared-
//     memory pipe the limiter), no probe loop, every lane walks the same instructions;
//   * which pairs are hot is fed back from the previous fold (hot_select kernels
//     below): tier A = the ~64 hottest, inserted first so they cannot lose a
//     bucket race, tier B = the rest up to the table size. With no history (first
//     window) pairs are admitted first-come while buckets have room;
//   * racing admissions may duplicate a key inside a bucket: harmless, both rows
//     are added into the global table when the CTA drains.
// Cold pairs go to the global dictionary: the home slot of all events of a thread
// is fetched before any is consumed (memory-level parallelism) and a hit there is
// reduced at once. Everything else — a new pair, a collision, an unresolvable
// source — is rare (~4 % of events) and loop-shaped, so it is NOT run inline: the
// lane pushes the event onto its warp's queue in shared memory (ballot-compacted,
// no atomics) and the warp runs the slow path for 32 queued events at a time.
// The main loop therefore stays converged: with the slow path inline, lanes
// came back from it in separate groups and the whole loop body issued ~1.55x
// (profiles/r1_v4b_ingest_ncu.txt: 20 of 32 lanes active on the loop's own code).
#include <cstdlib>

#include "alz_kernels.cuh"

namespace alz {

namespace {

constexpr int kLatSub = 4;              // latency sub-accumulators per row, picked by lane: every event of a pair adds to
                                        // its latency sum, so a hot pair's lanes would all serialise on one word
constexpr int kRowWords = ALZ_NB + 2 * kLatSub + 1;   // 64 hist cells, 4 x (lat_lo, lat_hi), err5xx = 73 (odd stride)
constexpr uint32_t kFwdBuckets = 128, kRevBuckets = 16, kWays = 4;
constexpr uint32_t kSlots = (kFwdBuckets + kRevBuckets) * kWays;
constexpr uint32_t kQueue = 64;          // slow-path queue entries per warp (ring)

struct Smem {
  uint64_t* keys;   // [kSlots]  bucket-major, 4 keys per bucket
  uint16_t* tags;   // [kSlots]  16 hash bits per way (0 = free): a lookup reads these 8 bytes, then one key
  uint32_t* fill;   // [kFwdBuckets + kRevBuckets] ways handed out
  uint32_t* rows;   // [kSlots * kRowWords]
};

__device__ __forceinline__ uint32_t bucket_of(uint32_t h, bool rv) {
  const uint32_t hb = h >> 20;   // high bits: independent of the dictionary's low-bit slot
  return rv ? kFwdBuckets + (hb & (kRevBuckets - 1u)) : (hb & (kFwdBuckets - 1u));
}
__device__ __forceinline__ uint32_t tag_of(uint32_t h) { return (h & 0xFFFFu) | 1u; }

// slot of key in its bucket or -1. One 8-byte load of the four tags, then the key of the first way whose
// tag matches (only lanes with a match load it). A second way with an equal tag, or a tag that is visible
// before its key, reads as a miss: the event then takes the global path, which is always correct.
// *room = the bucket still has a free way.
__device__ __forceinline__ int smem_lookup(const Smem& s, uint32_t bucket, uint64_t key, uint32_t tag, bool* room) {
  const uint2 t = *reinterpret_cast<const uint2*>(&s.tags[bucket * kWays]);
  const uint32_t t0 = t.x & 0xFFFFu, t1 = t.x >> 16, t2 = t.y & 0xFFFFu, t3 = t.y >> 16;
  *room = t3 == 0u;              // ways fill in order 0..3
  int w = -1;
  w = (t3 == tag) ? 3 : w;
  w = (t2 == tag) ? 2 : w;
  w = (t1 == tag) ? 1 : w;
  w = (t0 == tag) ? 0 : w;
  if (w >= 0 && s.keys[bucket * kWays + (uint32_t)w] != key) w = -1;
  return w < 0 ? -1 : (int)(bucket * kWays) + w;
}

// claim a free way of the bucket for key; -1 if the bucket is full
__device__ __forceinline__ int smem_admit(const Smem& s, uint32_t bucket, uint64_t key, uint32_t tag) {
  const uint32_t w = atomicAdd(&s.fill[bucket], 1u);
  if (w >= kWays) return -1;
  s.keys[bucket * kWays + w] = key;
  // no fence here on purpose: a reader that sees the tag before the key takes it for a miss (still correct),
  // and a __threadfence_block() anywhere in this kernel makes nvcc emit every global reduction as ATOMG
  // (with return) instead of REDG: +70 % kernel time (profiles/r1_v6_tagfence_ncu.txt)
  *reinterpret_cast<volatile uint16_t*>(&s.tags[bucket * kWays + w]) = (uint16_t)tag;
  return (int)(bucket * kWays + w);
}

__device__ __forceinline__ void smem_accumulate(const Smem& s, int slot, uint32_t bucket, uint64_t dur, bool err) {
  uint32_t* row = s.rows + (size_t)slot * kRowWords;
  atomicAdd(&row[bucket], 1u);
  uint32_t* lat = row + ALZ_NB + 2 * (threadIdx.x & (kLatSub - 1));
  const uint32_t lo = (uint32_t)dur;
  const uint32_t old = atomicAdd(&lat[0], lo);
  const uint32_t hi = (uint32_t)(dur >> 32) + ((old + lo < old) ? 1u : 0u);
  if (hi) atomicAdd(&lat[1], hi);
  if (err) atomicAdd(&row[ALZ_NB + 2 * kLatSub], 1u);
}

struct Ev {
  uint64_t key, dur;
  uint32_t bucket;
  bool act, rev, err;
};
// processL7's switch as bit tests (aggregator/data.go:1364-1383): request rows for
// HTTP(1) AMQP(2) POSTGRES(3) REDIS(5) MYSQL(7) MONGO(8); the SQL/Mongo ones unless rejected
__device__ __forceinline__ Ev decode(const Rec& r, bool live) {
  Ev e;
  const uint32_t p = rec_protocol(r), mf = rec_mflags(r);
  const bool row = p <= 8u && ((0x1AEu >> p) & 1u);
  const bool sql = p <= 8u && ((0x188u >> p) & 1u);
  e.act = live && row && !(sql && (mf & ALZ_MF_PAYLOAD_REJECT));
  e.rev = (mf & ALZ_MF_METHOD_MASK) == 2u && (p == ALZ_PROTO_AMQP || p == ALZ_PROTO_REDIS);  // DELIVER / PUSHED_EVENT
  e.key = ((uint64_t)rec_saddr(r) << 32) | rec_daddr(r);
  e.dur = rec_duration(r);
  e.bucket = latency_bucket(e.dur);
  e.err = is_5xx(p, rec_status(r));
  return e;
}

// streaming load: read once, keep it out of L1 and first in line for L2 eviction so the
// accumulator rows stay resident
__device__ __forceinline__ Rec load_rec_stream(const alz_l7_rec* p, uint64_t policy) {
  Rec r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
               : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]),
                 "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
               : "l"(p), "l"(policy));
  return r;
}

// tier A then tier B of one table's hot list into its buckets
__device__ __forceinline__ void preload_hot(const Smem& s, const HotState* hot, bool rv) {
  if (hot == nullptr) return;
  const uint32_t na = min(hot->n_a, (uint32_t)kHotA), nb = min(hot->n_b, (uint32_t)kHotB);
  for (uint32_t i = threadIdx.x; i < na; i += blockDim.x) {
    const uint64_t k = hot->keys_a[i];
    if (k != kEmptyKey) { const uint32_t hh = pair_hash(k); smem_admit(s, bucket_of(hh, rv), k, tag_of(hh)); }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) {
    const uint64_t k = hot->keys_b[i];
    if (k != kEmptyKey) { const uint32_t hh = pair_hash(k); smem_admit(s, bucket_of(hh, rv), k, tag_of(hh)); }
  }
}

// private rows [first, first + count) into global table g; a warp per slot
__device__ __forceinline__ void smem_drain(const Smem& s, uint32_t first, uint32_t count, const AccTable& g,
                                           const EpEntry* __restrict__ ep, uint32_t ep_mask, uint32_t* lost,
                                           uint32_t* unresolved) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (uint32_t i = warp; i < count; i += nwarps) {
    const uint32_t slot = first + i;
    const uint64_t key = s.keys[slot];
    if (key == kEmptyKey) continue;   // warp-uniform
    const uint32_t* row = s.rows + (size_t)slot * kRowWords;
    const uint32_t h0 = row[lane], h1 = row[32u + lane];
    uint32_t c = h0 + h1;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if (c == 0u) continue;            // preloaded but never hit in this launch: no global row needed
    uint32_t grow = 0;
    if (lane == 0) grow = find_or_insert_pair(g, key, ep, ep_mask);
    grow = __shfl_sync(0xFFFFFFFFu, grow, 0);
    if (grow >= kDropRow) {   // source is not a pod (dropped like the reference does) or capacity
      if (lane == 0) { if (grow == kDropRow) *unresolved += c; else *lost += c; }
      continue;
    }
    if (h0) atomicAdd(&g.hist[(size_t)grow * ALZ_NB + lane], h0);
    if (h1) atomicAdd(&g.hist[(size_t)grow * ALZ_NB + 32u + lane], h1);
    if (lane == 0) {
      uint64_t lat = 0;
      for (int q = 0; q < kLatSub; ++q) lat += ((uint64_t)row[ALZ_NB + 2 * q + 1] << 32) + row[ALZ_NB + 2 * q];
      if (lat) atomicAdd((unsigned long long*)&g.lat_sum[grow], (unsigned long long)lat);
      const uint32_t er = row[ALZ_NB + 2 * kLatSub];
      if (er) atomicAdd((unsigned long long*)&g.err5xx[grow], (unsigned long long)er);
    }
  }
}

// the warp's slow path: lane i takes queue entry (head + i) for i < count
__device__ __forceinline__ void slow_path_32(const uint64_t* q_key, const uint64_t* q_dur, const uint32_t* q_meta,
                                             uint32_t head, uint32_t count, const AccTable& fwd, const AccTable& rev,
                                             const EpEntry* __restrict__ ep, uint32_t ep_mask, uint32_t* lost,
                                             uint32_t* unresolved) {
  const uint32_t lane = threadIdx.x & 31u;
  const bool mine = lane < count;
  const uint32_t pos = (head + lane) & (kQueue - 1u);
  uint64_t key = 0, dur = 0;
  uint32_t meta = 0, row = kLostRow;
  if (mine) { key = q_key[pos]; dur = q_dur[pos]; meta = q_meta[pos]; }
  const bool rv = (meta & 0x100u) != 0u;
  if (mine) row = find_or_insert_pair(rv ? rev : fwd, key, ep, ep_mask);
  __syncwarp();
  if (mine) {
    if (row >= kDropRow) { if (row == kDropRow) *unresolved += 1u; else *lost += 1u; }
    else {
      const AccTable& t = rv ? rev : fwd;
      atomicAdd(&t.hist[(size_t)row * ALZ_NB + (meta & 0xFFu)], 1u);
      atomicAdd((unsigned long long*)&t.lat_sum[row], (unsigned long long)dur);
      if (meta & 0x200u) atomicAdd((unsigned long long*)&t.err5xx[row], 1ull);
    }
  }
  __syncwarp();
}

// a fetched dictionary home slot waiting to be used
struct Pend {
  uint4 ent;       // the 16-byte DictEnt as loaded
  uint64_t key, dur;
  uint32_t meta;   // bit 31 valid, bit 9 5xx, bit 8 reversed, bits 0..7 latency bucket
};

// use the probes: a hit reduces into its row at once, anything else joins the warp's slow-path queue
template <int kUnroll>
__device__ __forceinline__ void consume_probes(const Pend* p, const AccTable& fwd, const AccTable& rev,
                                               const EpEntry* __restrict__ ep, uint32_t ep_mask, uint64_t* q_key,
                                               uint64_t* q_dur, uint32_t* q_meta, uint32_t& q_head, uint32_t& q_count,
                                               uint32_t* lost, uint32_t* unresolved) {
  const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const bool g = (p[u].meta & 0x80000000u) != 0u;
    const bool rv = (p[u].meta & 0x100u) != 0u;
    const uint64_t k = ((uint64_t)p[u].ent.y << 32) | p[u].ent.x;
    const bool hit = g && k == p[u].key && p[u].ent.z < kDropRow && p[u].key != kEmptyKey;
    if (hit) {
      const AccTable& t = rv ? rev : fwd;
      atomicAdd(&t.hist[(size_t)p[u].ent.z * ALZ_NB + (p[u].meta & 0xFFu)], 1u);
      atomicAdd((unsigned long long*)&t.lat_sum[p[u].ent.z], (unsigned long long)p[u].dur);
      if (p[u].meta & 0x200u) atomicAdd((unsigned long long*)&t.err5xx[p[u].ent.z], 1ull);
    }
    const bool slow = g && !hit;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, slow);
    if (slow) {
      const uint32_t pos = (q_head + q_count + __popc(m & ((1u << lane) - 1u))) & (kQueue - 1u);
      q_key[pos] = p[u].key;
      q_dur[pos] = p[u].dur;
      q_meta[pos] = p[u].meta & 0x3FFu;
    }
    q_count += __popc(m);
    __syncwarp();
    if (q_count >= 32u) {
      slow_path_32(q_key, q_dur, q_meta, q_head, 32u, fwd, rev, ep, ep_mask, lost, unresolved);
      q_head = (q_head + 32u) & (kQueue - 1u);
      q_count -= 32u;
    }
  }
}

template <int kThreads, int kUnroll, bool kPrefetch, bool kPipe>
__global__ void __launch_bounds__(kThreads, 1) ingest_pairs_v4_kernel(const alz_l7_rec* __restrict__ recs, uint64_t n,
                                                                      AccTable fwd, AccTable rev, Counters* ctr,
                                                                      const HotState* hot_fwd, const HotState* hot_rev,
                                                                      const EpEntry* __restrict__ ep, uint32_t ep_mask) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Smem s;
  s.keys = reinterpret_cast<uint64_t*>(smem_raw);
  s.tags = reinterpret_cast<uint16_t*>(smem_raw + (size_t)kSlots * 8);
  s.fill = reinterpret_cast<uint32_t*>(smem_raw + (size_t)kSlots * 10);
  s.rows = s.fill + (kFwdBuckets + kRevBuckets);
  // per-warp slow-path queues behind the table
  uint8_t* qbase = reinterpret_cast<uint8_t*>(s.rows + (size_t)kSlots * kRowWords) + (size_t)(threadIdx.x >> 5) * kQueue * 20;
  uint64_t* q_key = reinterpret_cast<uint64_t*>(qbase);
  uint64_t* q_dur = q_key + kQueue;
  uint32_t* q_meta = reinterpret_cast<uint32_t*>(q_dur + kQueue);
  uint32_t q_head = 0, q_count = 0;
  for (uint32_t i = threadIdx.x; i < kSlots; i += kThreads) { s.keys[i] = kEmptyKey; s.tags[i] = 0; }
  for (uint32_t i = threadIdx.x; i < kFwdBuckets + kRevBuckets; i += kThreads) s.fill[i] = 0u;
  for (uint32_t i = threadIdx.x; i < kSlots * kRowWords; i += kThreads) s.rows[i] = 0u;
  __syncthreads();
  preload_hot(s, hot_fwd, false);
  preload_hot(s, hot_rev, true);
  __syncthreads();

  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));

  uint32_t not_request = 0, lost = 0, unresolved = 0;
  Pend pend[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) { pend[u].ent = make_uint4(0u, 0u, 0u, 0u); pend[u].key = 0; pend[u].dur = 0; pend[u].meta = 0u; }
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t stride = (uint64_t)gridDim.x * kThreads;
  const uint64_t first = (uint64_t)blockIdx.x * kThreads + (threadIdx.x & ~31u);
  Rec nxt[kUnroll];
  bool nlive[kUnroll];
  if (kPrefetch) {   // records of the first iteration
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint64_t j = first + (uint64_t)u * stride + lane;
      nlive[u] = j < n;
      nxt[u] = Rec{};
      if (nlive[u]) nxt[u] = load_rec_stream(recs + j, policy);
    }
  }
  for (uint64_t base = first; base < n; base += stride * kUnroll) {
    __syncwarp();
    Rec r[kUnroll];
    bool live[kUnroll];
    if (kPrefetch) {
      // take this iteration's records, put the next iteration's loads in flight before any processing
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) { r[u] = nxt[u]; live[u] = nlive[u]; }
      const uint64_t nb = base + stride * kUnroll;
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const uint64_t j = nb + (uint64_t)u * stride + lane;
        nlive[u] = j < n;
        if (nlive[u]) nxt[u] = load_rec_stream(recs + j, policy);
      }
    } else {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const uint64_t j = base + (uint64_t)u * stride + lane;
        live[u] = j < n;
        r[u] = Rec{};
        if (live[u]) r[u] = load_rec_stream(recs + j, policy);
      }
    }
    Ev e[kUnroll];
    int ss[kUnroll];
    uint32_t sb[kUnroll], hh[kUnroll];
    bool room[kUnroll];
    // stage 1: decode, shared-memory lookup (no loop, no divergence)
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      e[u] = decode(r[u], live[u]);
      not_request += (live[u] && !e[u].act) ? 1u : 0u;
      hh[u] = pair_hash(e[u].key);
      sb[u] = bucket_of(hh[u], e[u].rev);
      ss[u] = smem_lookup(s, sb[u], e[u].key, tag_of(hh[u]), &room[u]);
      if (!e[u].act || e[u].key == kEmptyKey) ss[u] = -1;
    }
    // stage 2: first-come admission of misses while their bucket has room (rare once warm)
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (room[u] && e[u].act && ss[u] < 0 && e[u].key != kEmptyKey)
        ss[u] = smem_admit(s, sb[u], e[u].key, tag_of(hh[u]));
    __syncwarp();
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (ss[u] >= 0) smem_accumulate(s, ss[u], e[u].bucket, e[u].dur, e[u].err);
    // stage 3: the rest goes to the global dictionary. The home slots are fetched now and, with kPipe,
    // consumed one iteration later, so the L2 round trip of the probe hides behind a whole iteration
    Pend cur[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const bool g = e[u].act && ss[u] < 0;
      const AccTable& t = e[u].rev ? rev : fwd;
      const uint32_t home = hh[u] & t.dict_mask;
      cur[u].key = e[u].key;
      cur[u].dur = e[u].dur;
      cur[u].meta = g ? (0x80000000u | e[u].bucket | (e[u].rev ? 0x100u : 0u) | (e[u].err ? 0x200u : 0u)) : 0u;
      cur[u].ent = make_uint4(0u, 0u, 0u, 0u);
      if (g) cur[u].ent = __ldcg(reinterpret_cast<const uint4*>(&t.dict[home]));
    }
    if (kPipe) {
      consume_probes<kUnroll>(pend, fwd, rev, ep, ep_mask, q_key, q_dur, q_meta, q_head, q_count, &lost, &unresolved);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) pend[u] = cur[u];
    } else {
      consume_probes<kUnroll>(cur, fwd, rev, ep, ep_mask, q_key, q_dur, q_meta, q_head, q_count, &lost, &unresolved);
    }
  }
  if (kPipe) consume_probes<kUnroll>(pend, fwd, rev, ep, ep_mask, q_key, q_dur, q_meta, q_head, q_count, &lost, &unresolved);
  if (q_count) slow_path_32(q_key, q_dur, q_meta, q_head, q_count, fwd, rev, ep, ep_mask, &lost, &unresolved);
  __syncthreads();
  smem_drain(s, 0u, kFwdBuckets * kWays, fwd, ep, ep_mask, &lost, &unresolved);
  smem_drain(s, kFwdBuckets * kWays, kRevBuckets * kWays, rev, ep, ep_mask, &lost, &unresolved);
  for (int o = 16; o > 0; o >>= 1) {
    not_request += __shfl_xor_sync(0xFFFFFFFFu, not_request, o);
    lost += __shfl_xor_sync(0xFFFFFFFFu, lost, o);
    unresolved += __shfl_xor_sync(0xFFFFFFFFu, unresolved, o);
  }
  if (lane == 0) {
    if (not_request) atomicAdd(&ctr->not_request, (unsigned long long)not_request);
    if (lost) atomicAdd(&ctr->capacity_events, (unsigned long long)lost);
    if (unresolved) atomicAdd(&ctr->src_unresolved, (unsigned long long)unresolved);
  }
}

// ---- hot-pair feedback: after a fold, pick the pairs that took the most events -----
// fold_pairs_kernel left row_cnt[row] and a 128-bin (quarter-octave) histogram of the
// counts; thresholds = the lowest bins that keep tier A <= kHotA and A+B <= target.
__global__ void hot_pick_kernel(HotState* hot, uint32_t target_total) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  uint32_t cum = 0, thr_a = 128, thr_b = 128;
  for (int b = 127; b >= 0; --b) {
    cum += hot->bins[b];
    if (cum <= (uint32_t)kHotA) thr_a = (uint32_t)b;
    if (cum <= target_total) thr_b = (uint32_t)b;
  }
  hot->thr_a = thr_a;
  hot->thr_b = thr_b;
  hot->n_a = 0;
  hot->n_b = 0;
}
__global__ void __launch_bounds__(256) hot_emit_kernel(AccTable pairs, HotState* hot) {
  const uint32_t n_rows = min(*pairs.n_rows, pairs.max_rows);
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t row = blockIdx.x * blockDim.x + threadIdx.x; row < n_rows; row += stride) {
    const uint32_t c = pairs.row_cnt[row];
    if (c == 0u) continue;
    const uint32_t b = count_bin(c);
    if (b >= hot->thr_a) {
      const uint32_t p = atomicAdd(&hot->n_a, 1u);
      if (p < (uint32_t)kHotA) hot->keys_a[p] = pairs.row_key[row];
    } else if (b >= hot->thr_b) {
      const uint32_t p = atomicAdd(&hot->n_b, 1u);
      if (p < (uint32_t)kHotB) hot->keys_b[p] = pairs.row_key[row];
    }
  }
}

}  // namespace

template <int kThreads, int kUnroll, bool kPrefetch, bool kPipe>
static void launch_variant(const alz_l7_rec* recs, uint64_t n, const AccTable& fwd, const AccTable& rev, Counters* ctr,
                           const HotState* hot_fwd, const HotState* hot_rev, const EpEntry* ep, uint32_t ep_mask,
                           int sms, cudaStream_t s) {
  const size_t smem = (size_t)kSlots * 10 + (size_t)(kFwdBuckets + kRevBuckets) * 4 + (size_t)kSlots * kRowWords * 4 +
                      (size_t)(kThreads / 32) * kQueue * 20;
  // per device (a process may drive several GPUs), so not cached in a static
  cudaFuncSetAttribute(ingest_pairs_v4_kernel<kThreads, kUnroll, kPrefetch, kPipe>,
                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ingest_pairs_v4_kernel<kThreads, kUnroll, kPrefetch, kPipe><<<(unsigned)sms, kThreads, smem, s>>>(
      recs, n, fwd, rev, ctr, hot_fwd, hot_rev, ep, ep_mask);
}

void launch_ingest_pairs_v4(const alz_l7_rec* recs, uint64_t n, const AccTable& fwd, const AccTable& rev,
                            Counters* ctr, const HotState* hot_fwd, const HotState* hot_rev, const EpEntry* ep,
                            uint32_t ep_mask, int sms, cudaStream_t s) {
  if (n == 0) return;
  // ALZ_INGEST_VARIANT: tuning knob for profiling runs (default = the measured best)
  static const int variant = [] { const char* v = getenv("ALZ_INGEST_VARIANT"); return v ? atoi(v) : 0; }();
  switch (variant) {
    case 1: launch_variant<1024, 2, false, true>(recs, n, fwd, rev, ctr, hot_fwd, hot_rev, ep, ep_mask, sms, s); break;
    case 2: launch_variant<768, 2, true, true>(recs, n, fwd, rev, ctr, hot_fwd, hot_rev, ep, ep_mask, sms, s); break;
    case 3: launch_variant<768, 2, false, true>(recs, n, fwd, rev, ctr, hot_fwd, hot_rev, ep, ep_mask, sms, s); break;
    case 4: launch_variant<1024, 2, true, false>(recs, n, fwd, rev, ctr, hot_fwd, hot_rev, ep, ep_mask, sms, s); break;
    case 5: launch_variant<1024, 2, true, true>(recs, n, fwd, rev, ctr, hot_fwd, hot_rev, ep, ep_mask, sms, s); break;
    default: launch_variant<1024, 2, false, false>(recs, n, fwd, rev, ctr, hot_fwd, hot_rev, ep, ep_mask, sms, s); break;
  }
}

// after fold_pairs_kernel(pairs, ..., hot->bins): choose next window's hot list
void launch_hot_select(const AccTable& pairs, HotState* hot, bool rev, int sms, cudaStream_t s) {
  const uint32_t target = rev ? (kRevBuckets * kWays * 7u) / 8u : (kFwdBuckets * kWays * 7u) / 8u;
  hot_pick_kernel<<<1, 32, 0, s>>>(hot, target);
  hot_emit_kernel<<<(unsigned)sms * 4, 256, 0, s>>>(pairs, hot);
}

}  // namespace alz
