// alz_ingest.cu — the dominant kernel: l7 events -> per-socket-pair accumulators
// (DESIGN.md §3 step 1, §5). One persistent CTA per SM.
//
// Data movement: every warp owns a two-stage ring in shared memory and streams its
// chunks of the record array into it with TMA bulk copies (cp.async.bulk + mbarrier
// complete_tx, issued by lane 0, L2 evict-first). Lanes copy their records from the
// ring into registers, the stage is handed back to the TMA as soon as those loads have
// returned, and the next two chunks are in flight while the warp works: no per-lane address arithmetic, no
// scoreboard stall on the first use of a streamed record, no cross-warp barrier
// anywhere in the main loop.
//
// Work per event, three tiers:
//   hot   the event's socket pair is in the CTA's shared-memory table (the stream is
//         Zipf-skewed: ~70 % of events): one direct-mapped probe (fingerprint + row,
//         verified against the row's key), one shared histogram increment, one shared
//         add of the latency. Which pairs are hot is fed back from the previous fold
//         (hot_select kernels below); 1/8 of the rows stay free for first-come
//         admission, which is also what the very first window runs on.
//   cold  everything else is NOT handled inline: the lane pushes the event onto its
//         warp's queue (ballot-compacted, no atomics) and the warp runs the cold path
//         for 32 queued events at a time with all lanes busy — global dictionary probe
//         of the home slot, then reductions (REDG) into the pair's row in L2. The hot
//         loop therefore carries no global-memory code and stays converged; r1's
//         kernel issued the cold path's instructions for every warp iteration
//         (profiles/r1_final_ingest_ncu.txt: 247 instructions per event).
//         The probes are cp.async copies into shared memory and are consumed when the
//         next batch is requested, so their L2/DRAM round trip is off the critical path.
//   slow  a cold event whose home slot does not hold its pair (new pair, collision,
//         unresolvable source: ~7 % of the cold events) is queued once more and 32 of
//         them at a time walk find_or_insert_pair; inline, nearly every cold batch would
//         run that loop for a lane or two.
// Reversed rows (AMQP DELIVER / REDIS PUSHED_EVENT, ~3 % of events) always go cold.
#include <cstdlib>

#include "alz_kernels.cuh"

namespace alz {

// Cycle accounting per section of the main loop, summed over warps (profiling build only: -DALZ_INGEST_PROF,
// alaz_b200/build.py --prof -> libalazgpu_prof.so; read with alz_debug_ingest_prof). Tells where a warp's time
// goes in a REAL run — ncu's kernel replay flushes the caches between passes and so shows every dictionary
// probe as a DRAM miss.
#ifdef ALZ_INGEST_PROF
__device__ unsigned long long g_ingest_prof[8];
#define PROF_DECL unsigned long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pc0 = 0, pc1 = 0; (void)pc0; (void)pc1
#define PROF_NOW() clock64()
#define PROF_ADD(i, v) pt[i] += (unsigned long long)(v)
#define PROF_FLUSH() do { if ((threadIdx.x & 31u) == 0) for (int i_ = 0; i_ < 8; ++i_) atomicAdd(&g_ingest_prof[i_], pt[i_]); } while (0)
#else
#define PROF_DECL long long pc0 = 0, pc1 = 0; (void)pc0; (void)pc1
#define PROF_NOW() 0ll
#define PROF_ADD(i, v)
#define PROF_FLUSH()
#endif

namespace {

constexpr int kLatSub = 4;                 // latency sub-accumulators per row, picked by lane: every event of a pair adds to
                                           // its latency sum, so a hot pair's lanes would all serialise on one word
constexpr int kRowWords = ALZ_NB + 2 * kLatSub + 1;   // 64 hist cells, 4 x (lat_lo, lat_hi), err5xx = 73 (odd stride)
constexpr uint32_t kTab = 4096;            // direct-mapped lookup entries: fingerprint (hash bits 19..0, bit 0 forced) << 12 | row
constexpr uint32_t kBusy = 0xFFFu;         // entry whose row field is no row: claimed, not (or never) published
constexpr uint32_t kSlowQ = 64;            // slow queue entries per warp (ring)
constexpr uint32_t kQBytes = 24;           // queue entry: key u64, dur u64, meta u32, pad
constexpr uint32_t kSmemMax = 232448;      // 227 KB per CTA on sm_100

template <int kWarps, int kU, int kRecWords>
struct Layout {
  // cold queue entries per warp (ring): up to 31 left over from the last iteration + 32 * kU new ones
  static constexpr uint32_t kColdQ = (32 * kU + 31 <= 64) ? 64u : (32 * kU + 31 <= 128) ? 128u : 256u;
  static constexpr uint32_t kChunk = 32u * kU;                          // events per chunk (per warp per iteration)
  static constexpr uint32_t kChunkBytes = kChunk * kRecWords * 4u;
  static constexpr uint32_t kRing = (uint32_t)kWarps * 2u * kChunkBytes;
  static constexpr uint32_t kBars = kRing;                              // kWarps * 2 mbarriers
  static constexpr uint32_t kTabOff = kBars + (uint32_t)kWarps * 16u;
  static constexpr uint32_t kColdOff = kTabOff + kTab * 4u;
  static constexpr uint32_t kSlowOff = kColdOff + (uint32_t)kWarps * kColdQ * kQBytes;
  static constexpr uint32_t kProbeOff = kSlowOff + (uint32_t)kWarps * kSlowQ * kQBytes;   // per warp: 32 x 16-B DictEnt
  static constexpr uint32_t kMisc = kProbeOff + (uint32_t)kWarps * 512u;                  // row allocator
  static constexpr uint32_t kRowKeys = kMisc + 16u;
  static constexpr uint32_t kFixed = kRowKeys + 8u;                     // + (kRows + 1) * 8 + kRows * kRowWords * 4
  static constexpr uint32_t kRows = ((kSmemMax - kFixed) / (kRowWords * 4u + 8u)) / 32u * 32u;
  static constexpr uint32_t kRowsOff = kRowKeys + (kRows + 1u) * 8u;
  static constexpr uint32_t kBytes = kRowsOff + kRows * kRowWords * 4u;
  static constexpr uint32_t kPreload = kRows - kRows / 8u;              // rows the hot list may take
  static_assert(kBytes <= kSmemMax, "shared memory layout too large");
  static_assert(kRows < kBusy, "row field is 12 bits");
};

// ---- TMA / mbarrier (PTX) -------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
// global -> shared bulk copy, completion counted on the mbarrier; L2 evict-first so that the stream does not
// push the accumulator rows out of L2
__device__ __forceinline__ void tma_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
// global reductions are written as PTX red so that no fence elsewhere in the kernel can turn them into
// returning atomics (r1 found nvcc emitting ATOMG for every atomicAdd once a __threadfence_block() was present)
__device__ __forceinline__ void red_add_u32(uint32_t* p, uint32_t v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_u64(uint64_t* p, uint64_t v) {
  asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// 16-byte global -> shared copy that no register waits on (LDGSTS); L2 only (the dictionary is written by other CTAs)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// shared-memory reductions under a predicate, written as PTX so that the hot tier is one straight line of code
// (with C++ ifs nvcc branches around each atomic and the two events of a lane cannot overlap)
__device__ __forceinline__ void red_shared_add_if(bool p, uint32_t addr, uint32_t v) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q red.shared.add.u32 [%0], %1;\n\t}"
               ::"r"(addr), "r"(v), "r"((uint32_t)p) : "memory");
}
__device__ __forceinline__ uint32_t atom_shared_add_if(bool p, uint32_t addr, uint32_t v) {
  uint32_t old = 0;
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q atom.shared.add.u32 %0, [%1], %2;\n\t}"
               : "+r"(old) : "r"(addr), "r"(v), "r"((uint32_t)p) : "memory");
  return old;
}
__device__ __forceinline__ bool elect_one() {   // one lane of the (converged) warp
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0u;
}
__device__ __forceinline__ uint32_t shr_clamp(uint32_t v, uint32_t by) {   // PTX shr: amounts > 31 give 0
  uint32_t r;
  asm("shr.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(by));
  return r;
}

// docs/SPEC.md §4 through the float exponent: round-toward-zero keeps floor(log2 d) and the next mantissa bit
// exact for every u64, so bits >> 22 = 2 * (127 + o) + bit(o-1); one conversion instead of a 64-bit clz chain
__device__ __forceinline__ uint32_t latency_bucket_rz(uint64_t d) {
  const int b = (int)(__float_as_uint(__ull2float_rz(d)) >> 22) - 2 * (127 + 8);
  return (uint32_t)min(max(b, 0), ALZ_NB - 1);
}

// processL7's switch (aggregator/data.go:1364-1383) as a 3-bit class per protocol value, packed in one word:
// bit 0 a request row is built (HTTP 1, AMQP 2, POSTGRES 3, REDIS 5, MYSQL 7, MONGO 8), bit 1 the row is dropped
// when the payload parser rejected it (POSTGRES, MYSQL, MONGO), bit 2 method 2 reverses the row (AMQP DELIVER,
// REDIS PUSHED_EVENT). Protocol values > 8 shift everything out: class 0 = no row.
constexpr uint32_t proto_class(uint32_t p) {
  return ((0x1AEu >> p) & 1u) | (((0x188u >> p) & 1u) << 1) | (((0x024u >> p) & 1u) << 2);
}
constexpr uint32_t kProtoLut = proto_class(0) | proto_class(1) << 3 | proto_class(2) << 6 | proto_class(3) << 9 |
                               proto_class(4) << 12 | proto_class(5) << 15 | proto_class(6) << 18 |
                               proto_class(7) << 21 | proto_class(8) << 24;

// hash of the per-CTA table only: two multiply-adds; the index comes from its top bits, the fingerprint from the
// rest. Weak low bits only cost a wasted verify (the row's key decides). The dictionary hash (pair_hash) is computed
// for cold events only.
__device__ __forceinline__ uint32_t table_hash(uint64_t key) {
  return (uint32_t)key * 0x9E3779B1u + (uint32_t)(key >> 32) * 0x85EBCA6Bu;
}
constexpr uint32_t kTabShift = 20;         // index = hash >> 20 (12 bits)

struct Shared {
  uint32_t* tab;      // [kTab]
  uint64_t* rowkey;   // [kRows + 1], entry kRows = kEmptyKey (never a hit)
  uint32_t* rows;     // [kRows * kRowWords]
  uint32_t* n_rows;   // rows handed out
};

__device__ __forceinline__ uint32_t tab_fp(uint32_t h) { return (h << 12) | 0x1000u; }   // never 0 in bits 31..12
__device__ __forceinline__ uint32_t tab_entry(uint32_t h, uint32_t row) { return tab_fp(h) | row; }

// claim the direct-mapped slot of `key` and give it a row; false if the slot is taken or the rows are used up.
// `publish_fenced`: other warps are probing concurrently, so the row's key must be visible before the entry
__device__ __forceinline__ bool smem_admit(const Shared& s, uint64_t key, uint32_t h, uint32_t limit, bool publish_fenced) {
  const uint32_t idx = h >> kTabShift;
  if (atomicCAS(&s.tab[idx], 0u, kBusy) != 0u) return false;
  const uint32_t row = atomicAdd(s.n_rows, 1u);
  if (row >= limit) return false;                       // slot stays kBusy: reads as a miss for everyone
  s.rowkey[row] = key;
  if (publish_fenced) __threadfence_block();
  *reinterpret_cast<volatile uint32_t*>(&s.tab[idx]) = tab_entry(h, row);
  return true;
}

// one warp's queue of deferred events in shared memory: a ring of 24-byte entries {key u64, dur u64, meta u32}
// filled by ballot compaction. meta: bits 0..5 latency bucket, bit 8 reversed row, bit 9 counts as 5xx, bit 10 host-keyed
template <uint32_t kCap>
struct Queue {
  uint8_t* base;
  uint32_t head, count;
  __device__ __forceinline__ void bind(uint8_t* b) { base = b; head = 0; count = 0; }
  __device__ __forceinline__ uint8_t* at(uint32_t i) const { return base + ((head + i) & (kCap - 1u)) * kQBytes; }
  // every lane calls; lanes with `want` append their event
  __device__ __forceinline__ uint32_t push(bool want, uint64_t k, uint64_t d, uint32_t m, uint32_t lane_lt) {
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, want);
    if (want) {
      uint8_t* e = at(count + __popc(mask & lane_lt));
      *reinterpret_cast<uint64_t*>(e) = k;
      *reinterpret_cast<uint64_t*>(e + 8) = d;
      *reinterpret_cast<uint32_t*>(e + 16) = m;
    }
    const uint32_t added = __popc(mask);
    count += added;
    return added;
  }
  __device__ __forceinline__ void get(uint32_t i, uint64_t* k, uint64_t* d, uint32_t* m) const {
    const uint8_t* e = at(i);
    *k = *reinterpret_cast<const uint64_t*>(e);
    *d = *reinterpret_cast<const uint64_t*>(e + 8);
    *m = *reinterpret_cast<const uint32_t*>(e + 16);
  }
  __device__ __forceinline__ void pop(uint32_t n) { head = (head + n) & (kCap - 1u); count -= n; }
};

// slow tier, 32 queued events at a time: the pair is new to the dictionary, or its home slot is taken by another
// pair, or its source is no pod (dropped like the reference does, aggregator/data.go:829-832)
template <uint32_t kRows>
__device__ __forceinline__ void slow_batch(Queue<kSlowQ>& q, uint32_t count, const AccTable& t, const Shared& s,
                                           const EpEntry* __restrict__ ep, uint32_t ep_mask, uint32_t* lost,
                                           uint32_t* unresolved) {
  const uint32_t lane = threadIdx.x & 31u;
  const bool mine = lane < count;
  if (mine) {
    uint64_t key, dur;
    uint32_t meta;
    q.get(lane, &key, &dur, &meta);
    const uint32_t kind = (meta >> 8) & 1u ? kPairRev : (meta & 0x400u) ? kPairHost : kPairFwd;
    const uint32_t row = find_or_insert_pair(t, key, kind, ep, ep_mask);
    if (row >= kDropRow) { if (row == kDropRow) *unresolved += 1u; else *lost += 1u; }
    else {
      red_add_u32(&t.hist[(size_t)row * ALZ_NB + (meta & 0x3Fu)], 1u);
      red_add_u64(&t.lat_sum[row], dur);
      if (meta & 0x200u) red_add_u64(&t.err5xx[row], 1ull);
      // a pair the dictionary did not know yet: give it a private row while some are left (first-come)
      if (kind == kPairFwd && key != kEmptyKey && *reinterpret_cast<volatile uint32_t*>(s.n_rows) < kRows)
        smem_admit(s, key, table_hash(key), kRows, true);
    }
  }
  __syncwarp();
  q.pop(count);
}

// cold tier, step 1: request the dictionary home slots of the first `count` queued events. The 16-byte entries
// are copied straight into the warp's probe buffer in shared memory (cp.async): nothing waits on them until the
// batch is consumed, which happens when the NEXT batch is ready to be requested — one to two iterations later,
// so the L2/DRAM round trip of the probes is off the warp's critical path. (Holding the probes in registers did
// not work: handing them from one loop trip to the next needs a move, and the move waits for the load.)
template <uint32_t kColdQ>
__device__ __forceinline__ void cold_issue(const Queue<kColdQ>& q, uint32_t count, const AccTable& t, uint32_t probe_a) {
  const uint32_t lane = threadIdx.x & 31u;
  if (lane < count) {
    uint64_t key, dur;
    uint32_t meta;
    q.get(lane, &key, &dur, &meta);
    const bool rv = (meta & 0x100u) != 0u, hk = (meta & 0x400u) != 0u;
    const DictEnt* dict = rv ? t.dict_rev : hk ? t.dict_host : t.dict;
    const uint32_t mask = rv ? t.dict_rev_mask : hk ? t.dict_host_mask : t.dict_mask;
    cp_async16(probe_a + lane * 16u, &dict[pair_hash(key) & mask]);
  }
  cp_async_commit();
}
// cold tier, step 2: the probes have landed. A home-slot hit is reduced into its row at once, anything else joins
// the slow queue.
template <uint32_t kRows, uint32_t kColdQ>
__device__ __forceinline__ void cold_consume(Queue<kColdQ>& q, uint32_t count, const uint4* probe, Queue<kSlowQ>& slow,
                                             const AccTable& t, const Shared& s, const EpEntry* __restrict__ ep,
                                             uint32_t ep_mask, uint32_t lane_lt, uint32_t* lost, uint32_t* unresolved,
                                             unsigned long long* t_wait, unsigned long long* t_slow) {
  const uint32_t lane = threadIdx.x & 31u;
  const long long w0 = PROF_NOW();
  cp_async_wait_all();
  *t_wait += (unsigned long long)(PROF_NOW() - w0);
  const bool valid = lane < count;
  uint64_t key = 0, dur = 0;
  uint32_t meta = 0;
  uint4 ent = make_uint4(0u, 0u, kNoRow, 0u);
  if (valid) { q.get(lane, &key, &dur, &meta); ent = probe[lane]; }
  const bool home = valid && ent.x == (uint32_t)key && ent.y == (uint32_t)(key >> 32) && ent.z < kDropRow && key != kEmptyKey;
  if (home) {
    red_add_u32(&t.hist[(size_t)ent.z * ALZ_NB + (meta & 0x3Fu)], 1u);
    red_add_u64(&t.lat_sum[ent.z], dur);
    if (meta & 0x200u) red_add_u64(&t.err5xx[ent.z], 1ull);
  }
  slow.push(valid && !home, key, dur, meta & 0x7FFu, lane_lt);
  __syncwarp();
  q.pop(count);
  if (slow.count >= 32u) {
    const long long s0 = PROF_NOW();
    slow_batch<kRows>(slow, 32u, t, s, ep, ep_mask, lost, unresolved);
    *t_slow += (unsigned long long)(PROF_NOW() - s0);
  }
}

// private rows into the global table; a warp per row
template <uint32_t kRows>
__device__ __forceinline__ void smem_drain(const Shared& s, const AccTable& g, const EpEntry* __restrict__ ep,
                                           uint32_t ep_mask, uint32_t* lost, uint32_t* unresolved) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const uint32_t used = min(*s.n_rows, kRows);
  for (uint32_t r = warp; r < used; r += nwarps) {
    const uint64_t key = s.rowkey[r];
    if (key == kEmptyKey) continue;   // warp-uniform
    const uint32_t* row = s.rows + (size_t)r * kRowWords;
    const uint32_t h0 = row[lane], h1 = row[32u + lane];
    uint32_t c = h0 + h1;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if (c == 0u) continue;            // preloaded but never hit in this launch: no global row needed
    uint32_t grow = 0;
    if (lane == 0) grow = find_or_insert_pair(g, key, kPairFwd, ep, ep_mask);
    grow = __shfl_sync(0xFFFFFFFFu, grow, 0);
    if (grow >= kDropRow) {   // source is not a pod (dropped like the reference does) or capacity
      if (lane == 0) { if (grow == kDropRow) *unresolved += c; else *lost += c; }
      continue;
    }
    if (h0) red_add_u32(&g.hist[(size_t)grow * ALZ_NB + lane], h0);
    if (h1) red_add_u32(&g.hist[(size_t)grow * ALZ_NB + 32u + lane], h1);
    if (lane == 0) {
      uint64_t lat = 0;
      for (int q = 0; q < kLatSub; ++q) lat += ((uint64_t)row[ALZ_NB + 2 * q + 1] << 32) + row[ALZ_NB + 2 * q];
      if (lat) red_add_u64(&g.lat_sum[grow], lat);
      const uint32_t er = row[ALZ_NB + 2 * kLatSub];
      if (er) red_add_u64(&g.err5xx[grow], (uint64_t)er);
    }
  }
}

// kRecWords = 8: alz_l7_rec (32 B). kRecWords = 4: alz_l7_rec16 (16 B; durations >= 2^32 ns sit in dur_ovf)
template <int kWarps, int kU, int kRecWords>
__global__ void __launch_bounds__(kWarps * 32, 1)
ingest_pairs_v6_kernel(const uint32_t* __restrict__ recs, uint64_t n, AccTable pairs, Counters* ctr,
                       const HotState* __restrict__ hot, const EpEntry* __restrict__ ep, uint32_t ep_mask,
                       const uint64_t* __restrict__ dur_ovf) {
  using L = Layout<kWarps, kU, kRecWords>;
  constexpr uint32_t kRows = L::kRows;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t lane_lt = (1u << lane) - 1u;
  Shared s;
  s.tab = reinterpret_cast<uint32_t*>(smem_raw + L::kTabOff);
  s.n_rows = reinterpret_cast<uint32_t*>(smem_raw + L::kMisc);
  s.rowkey = reinterpret_cast<uint64_t*>(smem_raw + L::kRowKeys);
  s.rows = reinterpret_cast<uint32_t*>(smem_raw + L::kRowsOff);
  constexpr uint32_t kColdQ = L::kColdQ;
  Queue<kColdQ> cold;
  Queue<kSlowQ> slow;
  cold.bind(smem_raw + L::kColdOff + (size_t)warp * kColdQ * kQBytes);
  slow.bind(smem_raw + L::kSlowOff + (size_t)warp * kSlowQ * kQBytes);
  const uint4* probe = reinterpret_cast<const uint4*>(smem_raw + L::kProbeOff + (size_t)warp * 512u);
  const uint32_t probe_a = smem_u32(probe);
  const uint32_t rows_a = smem_u32(s.rows);
  const uint8_t* ring = smem_raw + (size_t)warp * 2u * L::kChunkBytes;
  const uint32_t ring_a = smem_u32(ring);
  const uint32_t bar_a = smem_u32(smem_raw + L::kBars + warp * 16u);

  // chunks of this warp: c, c + stride, ... (a chunk = 32 * kU consecutive records); all but possibly the last
  // chunk of the array are full. 32-bit chunk numbers: n < 2^37 events per launch (the ABI layer splits above).
  const uint32_t n_chunks = (uint32_t)((n + L::kChunk - 1u) / L::kChunk);
  const uint32_t c_stride = gridDim.x * kWarps;
  const uint32_t c_first = blockIdx.x * kWarps + warp;
  const uint32_t tail = (uint32_t)(n - (uint64_t)(n_chunks - 1u) * L::kChunk);   // events in the last chunk, 1..kChunk
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  // producer state (used by the elected lane): the chunk two iterations ahead and its address
  uint32_t c_next = c_first;
  const uint32_t* src_next = recs + (uint64_t)c_first * (L::kChunk * kRecWords);
  const uint64_t src_step = (uint64_t)c_stride * (L::kChunk * kRecWords);        // in words
  // `dep` is always 0 but computed from the words just loaded out of the stage (see the main loop): the copy cannot
  // be issued before those loads have returned
  auto issue = [&](uint32_t stage, uint32_t dep) {   // one lane; requests chunk c_next if there is one
    if (c_next < n_chunks) {
      const uint32_t bytes = (c_next == n_chunks - 1u ? tail : L::kChunk) * (uint32_t)kRecWords * 4u + dep;
      mbar_expect_tx(bar_a + stage * 8u, bytes);
      tma_load(ring_a + stage * L::kChunkBytes, src_next, bytes, bar_a + stage * 8u, policy);
    }
  };
  // the first two chunks are requested before the table is even built
  if (lane == 0) {
    mbar_init(bar_a, 1u);
    mbar_init(bar_a + 8u, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    issue(0u, 0u);
    c_next += c_stride; src_next += src_step;
    issue(1u, 0u);
  }
  c_next += 2u * c_stride - (lane == 0 ? c_stride : 0u);   // every lane tracks the producer state, so any lane can be elected
  src_next += 2u * src_step - (lane == 0 ? src_step : 0u);

  for (uint32_t i = threadIdx.x; i < kTab; i += kWarps * 32) s.tab[i] = 0u;
  for (uint32_t i = threadIdx.x; i <= kRows; i += kWarps * 32) s.rowkey[i] = kEmptyKey;
  for (uint32_t i = threadIdx.x; i < kRows * kRowWords; i += kWarps * 32) s.rows[i] = 0u;
  if (threadIdx.x == 0) *s.n_rows = 0u;
  __syncthreads();
  if (hot != nullptr) {   // tier A first so that the hottest pairs cannot lose a slot to a cooler one
    const uint32_t na = min(hot->n_a, (uint32_t)kHotA);
    const uint32_t nb = min(min(hot->n_b, (uint32_t)(kHotMax - kHotA)), L::kPreload - min(na, L::kPreload));
    for (uint32_t i = threadIdx.x; i < na; i += kWarps * 32) {
      const uint64_t k = hot->keys[i];
      if (k != kEmptyKey) smem_admit(s, k, table_hash(k), L::kPreload, false);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nb; i += kWarps * 32) {
      const uint64_t k = hot->keys[kHotA + i];
      if (k != kEmptyKey) smem_admit(s, k, table_hash(k), L::kPreload, false);
    }
    __syncthreads();
    if (threadIdx.x == 0 && *s.n_rows > L::kPreload) *s.n_rows = L::kPreload;   // failed claims past the limit
  }
  __syncthreads();

  const uint32_t zero = (uint32_t)(n >> 63);   // n < 2^37
  uint32_t lost = 0, unresolved = 0;
  uint32_t probing = 0;                // events at the head of the cold queue whose probes are in flight
  uint32_t it = 0, n_hit = 0, n_live = 0, n_cold = 0;
  PROF_DECL;
  unsigned long long t_wait = 0, t_slow = 0;
  const long long p_begin = PROF_NOW();
  for (uint32_t c = c_first; c < n_chunks; c += c_stride, ++it) {
    const uint32_t stage = it & 1u;
    pc0 = PROF_NOW();
    mbar_wait(bar_a + stage * 8u, (it >> 1) & 1u);
    pc1 = PROF_NOW();
    PROF_ADD(1, pc1 - pc0);
    PROF_ADD(6, 1);
    // records of this chunk into registers (lane l takes records l, l + 32, ...)
    uint32_t w[kU][kRecWords];
    const uint8_t* st = ring + stage * L::kChunkBytes;
    uint32_t seen = 0;
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const uint4* p = reinterpret_cast<const uint4*>(st + ((size_t)u * 32u + lane) * (kRecWords * 4u));
      const uint4 a = p[0];
      w[u][0] = a.x; w[u][1] = a.y; w[u][2] = a.z; w[u][3] = a.w;
      seen ^= a.x;
      if (kRecWords == 8) {
        const uint2 b = *reinterpret_cast<const uint2*>(p + 1);   // duration; write_time is not read here
        w[u][4] = b.x; w[u][5] = b.y;
        seen ^= b.x;
      }
    }
    // The stage goes back to the TMA only when the loads above have RETURNED: the byte count of the copy is made
    // to depend on the loaded words (`zero` is a run-time 0 the compiler cannot see through), so the copy's issue
    // waits on their scoreboard. A __syncwarp() alone orders the instructions, not the completion of the
    // shared-memory loads, and under load the TMA write of the next chunk overtook the duration loads of this one
    // (right keys with the wrong latencies).
    __syncwarp();
    if (elect_one()) issue(stage, seen & zero);
    c_next += c_stride; src_next += src_step;
    const uint32_t n_here = (c == n_chunks - 1u) ? tail : L::kChunk;
    n_live += n_here;

    // hot tier for all kU events of the lane: straight-line code (independent chains), then one pass over the
    // cold queue
    uint64_t key[kU], dur[kU];
    uint32_t meta[kU];
    bool coldf[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const bool live = (uint32_t)u * 32u + lane < n_here;
      const uint32_t mw = (kRecWords == 8) ? w[u][3] : w[u][2];   // status | protocol << 16 | method_flags << 24
      uint32_t p = __byte_perm(mw, 0u, 0x4442u);                  // protocol byte, flag bits still on
      if (kRecWords == 8) dur[u] = ((uint64_t)w[u][5] << 32) | w[u][4];
      else {
        dur[u] = w[u][3];
        if (p & ALZ_REC16_DUR_OVERFLOW) dur[u] = live ? __ldg(&dur_ovf[w[u][3]]) : 0ull;
      }
      const bool hk = (p & ALZ_PROTO_F_HOSTKEY) != 0u;             // daddr is a Host-header id: own key space, cold tier
      p &= 0x3Fu;
      const uint32_t cls = shr_clamp(kProtoLut, 3u * p);
      // a row is built unless the class says "payload parser decides" and the parser said no (bit 30 of mw)
      const bool act = live && (cls & 1u) && !((cls & 2u) && (mw & ((uint32_t)ALZ_MF_PAYLOAD_REJECT << 24)));
      const bool rv = (cls & 4u) && (mw & ((uint32_t)ALZ_MF_METHOD_MASK << 24)) == (2u << 24);   // DELIVER / PUSHED_EVENT
      const bool err = p == ALZ_PROTO_HTTP && ((mw & 0xFFFFu) - 500u) < 100u;
      key[u] = ((uint64_t)w[u][1] << 32) | w[u][0];               // make_pair_key: the record's first two words as they lie
      const uint32_t bucket = latency_bucket_rz(dur[u]);
      meta[u] = bucket | (rv ? 0x100u : 0u) | (err ? 0x200u : 0u) | (hk ? 0x400u : 0u);

      // direct-mapped probe, verified against the row's key
      const uint32_t h = table_hash(key[u]);
      const uint32_t x = s.tab[h >> kTabShift] ^ tab_fp(h);
      const uint32_t r = min(x, kRows);
      const bool hit = act && !rv && !hk && x < kRows && s.rowkey[r] == key[u];
      // the row's reductions, each under the hit predicate
      const uint32_t row_a = rows_a + r * (kRowWords * 4u);
      const uint32_t lat_a = row_a + (ALZ_NB + 2u * (lane & (kLatSub - 1u))) * 4u;
      const uint32_t lo = (uint32_t)dur[u], dhi = (uint32_t)(dur[u] >> 32);
      red_shared_add_if(hit, row_a + bucket * 4u, 1u);
      const uint32_t old = atom_shared_add_if(hit, lat_a, lo);
      const bool carry = old > ~lo;                                  // out of the low word
      red_shared_add_if(hit && (carry || dhi != 0u), lat_a + 4u, dhi + (carry ? 1u : 0u));
      red_shared_add_if(hit && err, row_a + (ALZ_NB + 2 * kLatSub) * 4u, 1u);
      n_hit += hit ? 1u : 0u;
      coldf[u] = act && !hit;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) n_cold += cold.push(coldf[u], key[u], dur[u], meta[u], lane_lt);
    __syncwarp();
    pc0 = PROF_NOW();
    PROF_ADD(5, pc0 - pc1);
    // a batch is consumed when the next one is ready to be requested
    while (cold.count - probing >= 32u) {
      if (probing) {
        cold_consume<kRows>(cold, probing, probe, slow, pairs, s, ep, ep_mask, lane_lt, &lost, &unresolved, &t_wait, &t_slow);
        probing = 0;
        PROF_ADD(7, 1);
      }
      cold_issue(cold, 32u, pairs, probe_a);
      probing = 32u;
    }
    PROF_ADD(3, PROF_NOW() - pc0);
  }
  if (probing) cold_consume<kRows>(cold, probing, probe, slow, pairs, s, ep, ep_mask, lane_lt, &lost, &unresolved, &t_wait, &t_slow);
  if (cold.count) {
    const uint32_t rest = cold.count;
    cold_issue(cold, rest, pairs, probe_a);
    cold_consume<kRows>(cold, rest, probe, slow, pairs, s, ep, ep_mask, lane_lt, &lost, &unresolved, &t_wait, &t_slow);
  }
  PROF_ADD(0, PROF_NOW() - p_begin);
  PROF_ADD(2, t_wait);
  PROF_ADD(4, t_slow);
  PROF_FLUSH();
  while (slow.count) slow_batch<kRows>(slow, min(slow.count, 32u), pairs, s, ep, ep_mask, &lost, &unresolved);
  __syncthreads();
  smem_drain<kRows>(s, pairs, ep, ep_mask, &lost, &unresolved);
  for (int o = 16; o > 0; o >>= 1) {
    n_hit += __shfl_xor_sync(0xFFFFFFFFu, n_hit, o);
    lost += __shfl_xor_sync(0xFFFFFFFFu, lost, o);
    unresolved += __shfl_xor_sync(0xFFFFFFFFu, unresolved, o);
  }
  // events that built no request row = events seen - hot hits - cold pushes (the last two are counted anyway)
  const uint32_t not_request = n_live - n_cold - n_hit;
  if (lane == 0) {
    if (not_request) atomicAdd(&ctr->not_request, (unsigned long long)not_request);
    if (lost) atomicAdd(&ctr->capacity_events, (unsigned long long)lost);
    if (unresolved) atomicAdd(&ctr->src_unresolved, (unsigned long long)unresolved);
  }
}

// ---- hot-pair feedback: after a fold, pick the pairs that took the most events -----
// fold_pairs_kernel left row_cnt[row] and a 128-bin (quarter-octave) histogram of the counts of the
// forward rows; thresholds = the lowest bins that keep tier A <= kHotA and A+B <= target. Every block
// derives the same two thresholds from the bins itself (128 adds) instead of a separate launch.
__global__ void __launch_bounds__(256) hot_emit_kernel(AccTable pairs, HotState* hot, uint32_t target_total) {
  __shared__ uint32_t s_thr[2];
  if (threadIdx.x == 0) {
    uint32_t cum = 0, thr_a = 128, thr_b = 128;
    for (int b = 127; b >= 0; --b) {
      cum += hot->bins[b];
      if (cum <= (uint32_t)kHotA) thr_a = (uint32_t)b;
      if (cum <= target_total) thr_b = (uint32_t)b;
    }
    s_thr[0] = thr_a; s_thr[1] = thr_b;
    if (blockIdx.x == 0) { hot->thr_a = thr_a; hot->thr_b = thr_b; }
  }
  __syncthreads();
  const uint32_t thr_a = s_thr[0], thr_b = s_thr[1];
  const uint32_t n_rows = min(*pairs.n_rows, pairs.max_rows);
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t row = blockIdx.x * blockDim.x + threadIdx.x; row < n_rows; row += stride) {
    const uint32_t c = pairs.row_cnt[row];
    if (c == 0u || pairs.row_kind[row] != kPairFwd) continue;   // only forward pairs enter the per-CTA table
    const uint32_t b = count_bin(c);
    if (b >= thr_a) {
      const uint32_t p = atomicAdd(&hot->n_a, 1u);
      if (p < (uint32_t)kHotA) hot->keys[p] = pairs.row_key[row];
    } else if (b >= thr_b) {
      const uint32_t p = atomicAdd(&hot->n_b, 1u);
      if (p < (uint32_t)(kHotMax - kHotA)) hot->keys[kHotA + p] = pairs.row_key[row];
    }
  }
}

template <int kWarps, int kU, int kRecWords>
void launch_variant(const void* recs, uint64_t n, const AccTable& pairs, Counters* ctr, const HotState* hot,
                    const EpEntry* ep, uint32_t ep_mask, const uint64_t* dur_ovf, int sms, cudaStream_t s) {
  using L = Layout<kWarps, kU, kRecWords>;
  // per device (a process may drive several GPUs), so not cached in a static
  cudaFuncSetAttribute(ingest_pairs_v6_kernel<kWarps, kU, kRecWords>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)L::kBytes);
  ingest_pairs_v6_kernel<kWarps, kU, kRecWords><<<(unsigned)sms, kWarps * 32, L::kBytes, s>>>(
      (const uint32_t*)recs, n, pairs, ctr, hot, ep, ep_mask, dur_ovf);
}

}  // namespace

uint32_t ingest_table_rows() { return Layout<16, 2, 8>::kRows; }

void launch_ingest_pairs_v6(const alz_l7_rec* recs, uint64_t n, const AccTable& pairs, Counters* ctr,
                            const HotState* hot, const EpEntry* ep, uint32_t ep_mask, int sms, cudaStream_t s) {
  if (n == 0) return;
  // ALZ_INGEST_SHAPE: CTA shape for profiling runs (default = the measured best)
  static const int shape = [] { const char* v = getenv("ALZ_INGEST_SHAPE"); return v ? atoi(v) : 0; }();
  switch (shape) {
    case 1: launch_variant<32, 1, 8>(recs, n, pairs, ctr, hot, ep, ep_mask, nullptr, sms, s); break;
    case 2: launch_variant<12, 2, 8>(recs, n, pairs, ctr, hot, ep, ep_mask, nullptr, sms, s); break;
    case 3: launch_variant<20, 2, 8>(recs, n, pairs, ctr, hot, ep, ep_mask, nullptr, sms, s); break;
    default: launch_variant<16, 2, 8>(recs, n, pairs, ctr, hot, ep, ep_mask, nullptr, sms, s); break;
  }
}

void launch_ingest_pairs_v6_rec16(const alz_l7_rec16* recs, uint64_t n, const uint64_t* dur_ovf, const AccTable& pairs,
                                  Counters* ctr, const HotState* hot, const EpEntry* ep, uint32_t ep_mask, int sms,
                                  cudaStream_t s) {
  if (n == 0) return;
  launch_variant<16, 2, 4>(recs, n, pairs, ctr, hot, ep, ep_mask, dur_ovf, sms, s);
}

// after fold_pairs_kernel(pairs, ..., hot->bins): choose next window's hot list
void launch_hot_select(const AccTable& pairs, HotState* hot, int sms, cudaStream_t s) {
  hot_emit_kernel<<<(unsigned)sms * 2, 256, 0, s>>>(pairs, hot, Layout<16, 2, 8>::kPreload);
}

}  // namespace alz

// profiling build only: cycles per section of the ingest main loop since the last call (see g_ingest_prof)
extern "C" int alz_debug_ingest_prof(unsigned long long* out8) {
#ifdef ALZ_INGEST_PROF
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cudaDeviceSynchronize() != cudaSuccess) return ALZ_E_CUDA;
  if (cudaMemcpyFromSymbol(out8, alz::g_ingest_prof, sizeof(z)) != cudaSuccess) return ALZ_E_CUDA;
  if (cudaMemcpyToSymbol(alz::g_ingest_prof, z, sizeof(z)) != cudaSuccess) return ALZ_E_CUDA;
  return ALZ_OK;
#else
  (void)out8;
  return ALZ_E_UNSUPPORTED;
#endif
}
