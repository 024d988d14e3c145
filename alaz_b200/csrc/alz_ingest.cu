// alz_ingest.cu — the dominant kernel: l7 events -> per-socket-pair accumulators
// (DESIGN.md §3 step 1, §5). One persistent CTA per SM.
//
// What bounds it (profiles/r2_ingest_sections.txt, measured with the instrumented build on warm
// caches): neither DRAM nor instruction issue but the SM's load/store path (L1TEX): every shared-
// memory wavefront and every scattered global request (a dictionary probe, a reduction into L2)
// takes a slot there, and round 2's first kernel spent ~105 of them per 32 events. This version is
// organised around spending fewer:
//   * records come straight from L2 into registers (one 32-byte load per record, evict-first);
//     the TMA engine is used to PREFETCH the stream into L2 three iterations ahead
//     (cp.async.bulk.prefetch.L2), so the loads find their lines in L2 and nothing is staged
//     through shared memory (a staged copy costs a write and a read of the data path per record);
//   * the per-CTA table of hot socket pairs holds a 16-bucket WINDOW of the latency histogram per
//     pair in 16-bit cells (52 bytes a row instead of 292), which is where a pair's latencies
//     fall (the window is chosen from the pair's own histogram at the previous fold); ~2700 pairs
//     fit instead of 224 and ~80 % of the events end there: a hit is two probes of a 2-choice
//     direct-mapped index, one key load, two shared reductions;
//   * everything else (a cold pair, a latency outside the row's window, a reversed or host-keyed
//     row) goes to the global pair table: one dictionary probe and two reductions (REDG) per event,
//     issued inline under predicates; only events whose home slot does not hold their pair (new
//     pair, collision, unresolvable source) are queued in shared memory and walk the dictionary
//     32 at a time.
// 16-bit cells stay exact: a cell that reaches 0x7FFF is spilled into the global table by the lane
// that saw it (bit 15 is head room for the increments that race with the spill).
#include <cstdlib>

#include "alz_kernels.cuh"

namespace alz {

// Cycle accounting per section of the main loop, summed over warps (profiling build only: -DALZ_INGEST_PROF,
// alaz_b200/build.py --prof -> libalazgpu_prof.so; read with alz_debug_ingest_prof). Tells where a warp's time
// goes in a REAL run — ncu's kernel replay flushes the caches between passes unless told otherwise.
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

constexpr uint32_t kRowWords = 13;         // 8 words = 16 x u16 histogram cells, err5xx u32, 2 x (lat_lo, lat_hi); odd stride
constexpr uint32_t kCellSpill = 0x7FFFu;   // a 16-bit cell seen at this value is spilled (bit 15 = head room)
constexpr uint32_t kTab = 8192;            // index entries: fingerprint 16 | window base 4 | row 12, two choices per key
constexpr uint32_t kTabShift = 19;
constexpr uint32_t kRowMask = 0xFFFu;
constexpr uint32_t kBusy = kRowMask;       // entry whose row field is no row: claimed, not (or never) published
constexpr uint32_t kSlowQ = 64;            // slow queue entries per warp (ring)
constexpr uint32_t kQBytes = 16;           // {key u64, dur_lo u32, meta u32}
constexpr uint32_t kSmemMax = 232448;      // 227 KB per CTA on sm_100
constexpr uint32_t kPrefetchAhead = 3;     // iterations

template <int kWarps>
struct Layout {
  static constexpr uint32_t kTabOff = 0;
  static constexpr uint32_t kSlowOff = kTabOff + kTab * 4u;
  static constexpr uint32_t kMisc = kSlowOff + (uint32_t)kWarps * kSlowQ * kQBytes;   // row allocator
  static constexpr uint32_t kRowKeys = kMisc + 16u;
  static constexpr uint32_t kPerRow = 8u + kRowWords * 4u + 1u;         // key, cells, window base
  static constexpr uint32_t kRowsRaw = (kSmemMax - kRowKeys - 64u) / kPerRow - 1u;
  static constexpr uint32_t kRows = (kRowsRaw < 4064u ? kRowsRaw : 4064u) / 32u * 32u;
  static constexpr uint32_t kRowsOff = kRowKeys + (kRows + 1u) * 8u;    // row kRows: scratch target of clamped indices
  static constexpr uint32_t kBaseOff = kRowsOff + (kRows + 1u) * kRowWords * 4u;
  static constexpr uint32_t kBytes = kBaseOff + ((kRows + 1u + 15u) / 16u) * 16u;
  static constexpr uint32_t kPreload = kRows - kRows / 8u;              // rows the hot list may take
  static_assert(kBytes <= kSmemMax, "shared memory layout too large");
  static_assert(kRows < kBusy, "row field is 12 bits");
};

// ---- PTX helpers ----------------------------------------------------------------------------
// global reductions are written as PTX red so that no fence elsewhere in the kernel can turn them into
// returning atomics (r1 found nvcc emitting ATOMG for every atomicAdd once a __threadfence_block() was present)
__device__ __forceinline__ void red_add_u32(uint32_t* p, uint32_t v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_u64(uint64_t* p, uint64_t v) {
  asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
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
// the TMA engine pulls `bytes` of the stream into L2 ahead of the loads (no shared memory involved)
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(p), "r"(bytes), "l"(policy) : "memory");
}
// streaming load of one 32-byte record: read once, keep it out of L1 and first in line for L2 eviction
__device__ __forceinline__ void load_rec32(const uint32_t* p, uint64_t policy, uint32_t (&w)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
               : "l"(p), "l"(policy));
}
__device__ __forceinline__ void load_rec16(const uint32_t* p, uint64_t policy, uint32_t (&w)[4]) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p), "l"(policy));
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

// hash of the per-CTA table only: two multiply-adds; index 1 from its top bits, index 2 from the top bits of
// one more multiply, the fingerprint from its low bits. Weak low bits only cost a wasted key load (the row's key
// decides). The dictionary hash (pair_hash) is computed for cold events only.
__device__ __forceinline__ uint32_t table_hash(uint64_t key) {
  return (uint32_t)key * 0x9E3779B1u + (uint32_t)(key >> 32) * 0x85EBCA6Bu;
}
__device__ __forceinline__ uint32_t tab_idx1(uint32_t h) { return h >> kTabShift; }
__device__ __forceinline__ uint32_t tab_idx2(uint32_t h) { return (h * 0xC2B2AE35u) >> kTabShift; }
__device__ __forceinline__ uint32_t tab_fp(uint32_t h) { return (h << 16) | 0x10000u; }   // never 0 in bits 31..16
__device__ __forceinline__ uint32_t tab_entry(uint32_t h, uint32_t base4, uint32_t row) { return tab_fp(h) | (base4 << 12) | row; }

struct Shared {
  uint32_t* tab;      // [kTab]
  uint64_t* rowkey;   // [kRows + 1], entry kRows = kEmptyKey (never a hit)
  uint32_t* rows;     // [(kRows + 1) * kRowWords]
  uint8_t* rowbase;   // [kRows + 1] first bucket / 4 of the row's window (also in its index entry)
  uint32_t* n_rows;   // rows handed out
};

// claim one of the key's two index slots and give it a row; false if both are taken or the rows are used up.
// `publish_fenced`: other warps are probing concurrently, so the row's key must be visible before the entry
__device__ __forceinline__ bool smem_admit(const Shared& s, uint64_t key, uint32_t h, uint32_t base4, uint32_t limit,
                                           bool publish_fenced) {
  uint32_t idx = tab_idx1(h);
  if (atomicCAS(&s.tab[idx], 0u, kBusy) != 0u) {
    idx = tab_idx2(h);
    if (atomicCAS(&s.tab[idx], 0u, kBusy) != 0u) return false;
  }
  const uint32_t row = atomicAdd(s.n_rows, 1u);
  if (row >= limit) return false;                       // slot stays kBusy: reads as a miss for everyone
  s.rowkey[row] = key;
  s.rowbase[row] = (uint8_t)base4;
  if (publish_fenced) __threadfence_block();
  *reinterpret_cast<volatile uint32_t*>(&s.tab[idx]) = tab_entry(h, base4, row);
  return true;
}
// window for a pair admitted without history: centred on the bucket of the event that brought it in
__device__ __forceinline__ uint32_t base4_around(uint32_t bucket) { return (uint32_t)min(max((int)bucket - 6, 0), 48) >> 2; }

// one warp's queue of slow events in shared memory: a ring of 16-byte entries {key u64, dur_lo u32, meta u32}
// filled by ballot compaction. meta: bits 0..5 latency bucket, 6..7 pair kind, bit 8 counts as 5xx, bits 9..31
// the duration's high word (events whose duration does not fit, >= 2^55 ns, are handled on the spot)
struct SlowQueue {
  uint8_t* base;
  uint32_t head, count;
  __device__ __forceinline__ void bind(uint8_t* b) { base = b; head = 0; count = 0; }
  __device__ __forceinline__ uint4* at(uint32_t i) const {
    return reinterpret_cast<uint4*>(base + ((head + i) & (kSlowQ - 1u)) * kQBytes);
  }
  __device__ __forceinline__ void push(bool want, uint64_t k, uint32_t dlo, uint32_t m, uint32_t lane_lt) {
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, want);
    if (want) *at(count + __popc(mask & lane_lt)) = make_uint4((uint32_t)k, (uint32_t)(k >> 32), dlo, m);
    count += __popc(mask);
  }
  __device__ __forceinline__ void pop(uint32_t n) { head = (head + n) & (kSlowQ - 1u); count -= n; }
};

// the global path for one event whose pair row is known
__device__ __forceinline__ void global_add(const AccTable& t, uint32_t row, uint32_t bucket, uint64_t dur, bool err) {
  red_add_u32(&t.hist[(size_t)row * ALZ_NB + bucket], 1u);
  red_add_u64(&t.lat_sum[row], dur);
  if (err) red_add_u64(&t.err5xx[row], 1ull);
}

// slow tier for one event (walks the dictionary): the pair is new to it, or its home slot is taken by another
// pair, or its source is no pod (dropped like the reference does, aggregator/data.go:829-832)
template <uint32_t kRows>
__device__ __forceinline__ void slow_one(uint64_t key, uint64_t dur, uint32_t bucket, uint32_t kind, bool err,
                                         const AccTable& t, const Shared& s, const EpEntry* __restrict__ ep,
                                         uint32_t ep_mask, uint32_t* lost, uint32_t* unresolved) {
  const uint32_t row = find_or_insert_pair(t, key, kind, ep, ep_mask);
  if (row >= kDropRow) { if (row == kDropRow) *unresolved += 1u; else *lost += 1u; return; }
  global_add(t, row, bucket, dur, err);
  // a pair the per-CTA table does not hold: give it a private row while some are left (first-come)
  if (kind == kPairFwd && key != kEmptyKey && *reinterpret_cast<volatile uint32_t*>(s.n_rows) < kRows)
    smem_admit(s, key, table_hash(key), base4_around(bucket), kRows, true);
}
template <uint32_t kRows>
__device__ __forceinline__ void slow_batch(SlowQueue& q, uint32_t count, const AccTable& t, const Shared& s,
                                           const EpEntry* __restrict__ ep, uint32_t ep_mask, uint32_t* lost,
                                           uint32_t* unresolved) {
  const uint32_t lane = threadIdx.x & 31u;
  if (lane < count) {
    const uint4 e = *q.at(lane);
    const uint64_t key = ((uint64_t)e.y << 32) | e.x;
    const uint64_t dur = ((uint64_t)(e.w >> 9) << 32) | e.z;
    slow_one<kRows>(key, dur, e.w & 0x3Fu, (e.w >> 6) & 3u, (e.w & 0x100u) != 0u, t, s, ep, ep_mask, lost, unresolved);
  }
  __syncwarp();
  q.pop(count);
}

// kRecWords = 8: alz_l7_rec (32 B). kRecWords = 4: alz_l7_rec16 (16 B; durations >= 2^32 ns sit in dur_ovf)
template <int kWarps, int kRecWords>
__global__ void __launch_bounds__(kWarps * 32, 1)
ingest_pairs_v7_kernel(const uint32_t* __restrict__ recs, uint64_t n, AccTable pairs, Counters* ctr,
                       const HotState* __restrict__ hot, const EpEntry* __restrict__ ep, uint32_t ep_mask,
                       const uint64_t* __restrict__ dur_ovf) {
  using L = Layout<kWarps>;
  constexpr uint32_t kRows = L::kRows;
  constexpr int kU = 2;
  constexpr uint32_t kChunk = 32u * kU;                                  // events per warp per iteration
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t lane_lt = (1u << lane) - 1u;
  Shared s;
  s.tab = reinterpret_cast<uint32_t*>(smem_raw + L::kTabOff);
  s.n_rows = reinterpret_cast<uint32_t*>(smem_raw + L::kMisc);
  s.rowkey = reinterpret_cast<uint64_t*>(smem_raw + L::kRowKeys);
  s.rows = reinterpret_cast<uint32_t*>(smem_raw + L::kRowsOff);
  s.rowbase = smem_raw + L::kBaseOff;
  SlowQueue slow;
  slow.bind(smem_raw + L::kSlowOff + (size_t)warp * kSlowQ * kQBytes);

  // chunks of this warp: c, c + stride, ... (a chunk = 64 consecutive records); 32-bit chunk numbers:
  // n < 2^37 events per launch (the ABI layer splits above)
  const uint32_t n_chunks = (uint32_t)((n + kChunk - 1u) / kChunk);
  const uint32_t c_stride = gridDim.x * kWarps;
  const uint32_t c_first = blockIdx.x * kWarps + warp;
  const uint32_t tail = (uint32_t)(n - (uint64_t)(n_chunks - 1u) * kChunk);   // events in the last chunk, 1..kChunk
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  constexpr uint32_t kChunkWords = kChunk * kRecWords;
  auto prefetch_chunk = [&](uint32_t c) {   // one lane
    if (c < n_chunks)
      prefetch_l2(recs + (uint64_t)c * kChunkWords, (c == n_chunks - 1u ? tail : kChunk) * (uint32_t)kRecWords * 4u, policy);
  };
  if (lane == 0)
    for (uint32_t a = 0; a < kPrefetchAhead; ++a) prefetch_chunk(c_first + a * c_stride);

  for (uint32_t i = threadIdx.x; i < kTab; i += kWarps * 32) s.tab[i] = 0u;
  for (uint32_t i = threadIdx.x; i <= kRows; i += kWarps * 32) { s.rowkey[i] = kEmptyKey; s.rowbase[i] = 0; }
  for (uint32_t i = threadIdx.x; i < (kRows + 1u) * kRowWords; i += kWarps * 32) s.rows[i] = 0u;
  if (threadIdx.x == 0) *s.n_rows = 0u;
  __syncthreads();
  if (hot != nullptr) {   // tier A first so that the hottest pairs cannot lose their slots to cooler ones
    const uint32_t na = min(hot->n_a, (uint32_t)kHotA);
    const uint32_t nb = min(min(hot->n_b, (uint32_t)(kHotMax - kHotA)), L::kPreload - min(na, L::kPreload));
    for (uint32_t i = threadIdx.x; i < na; i += kWarps * 32) {
      const uint64_t k = hot->keys[i];
      if (k != kEmptyKey) smem_admit(s, k, table_hash(k), min((uint32_t)hot->base[i], 12u), L::kPreload, false);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nb; i += kWarps * 32) {
      const uint64_t k = hot->keys[kHotA + i];
      if (k != kEmptyKey) smem_admit(s, k, table_hash(k), min((uint32_t)hot->base[kHotA + i], 12u), L::kPreload, false);
    }
    __syncthreads();
    if (threadIdx.x == 0 && *s.n_rows > L::kPreload) *s.n_rows = L::kPreload;   // failed claims past the limit
  }
  __syncthreads();

  uint32_t lost = 0, unresolved = 0, n_hit = 0, n_live = 0, n_cold = 0;
  PROF_DECL;
  const long long p_begin = PROF_NOW();
  for (uint32_t c = c_first; c < n_chunks; c += c_stride) {
    pc0 = PROF_NOW();
    PROF_ADD(6, 1);
    const uint32_t n_here = (c == n_chunks - 1u) ? tail : kChunk;
    n_live += n_here;
    if (elect_one()) prefetch_chunk(c + kPrefetchAhead * c_stride);
    // this iteration's records (lane l takes records l and l + 32 of the chunk)
    uint32_t w[kU][kRecWords];
    bool live[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      live[u] = (uint32_t)u * 32u + lane < n_here;
#pragma unroll
      for (int k = 0; k < kRecWords; ++k) w[u][k] = 0u;
      const uint32_t* p = recs + ((uint64_t)c * kChunk + (uint32_t)u * 32u + lane) * kRecWords;
      if (live[u]) { if constexpr (kRecWords == 8) load_rec32(p, policy, w[u]); else load_rec16(p, policy, w[u]); }
    }

    // hot tier for both events of the lane, the dictionary probe of a cold one requested right away
    uint64_t key[kU], dur[kU];
    uint32_t meta[kU];      // bits 0..5 bucket, 6..7 kind, 8 err, 31 cold
    uint4 ent[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const uint32_t mw = (kRecWords == 8) ? w[u][3] : w[u][2];   // status | protocol << 16 | method_flags << 24
      uint32_t p = __byte_perm(mw, 0u, 0x4442u);                  // protocol byte, flag bits still on
      if (kRecWords == 8) dur[u] = ((uint64_t)w[u][5] << 32) | w[u][4];
      else {
        dur[u] = w[u][3];
        if (p & ALZ_REC16_DUR_OVERFLOW) dur[u] = live[u] ? __ldg(&dur_ovf[w[u][3]]) : 0ull;
      }
      const bool hk = (p & ALZ_PROTO_F_HOSTKEY) != 0u;             // daddr is a Host-header id: own key space
      p &= 0x3Fu;
      const uint32_t cls = shr_clamp(kProtoLut, 3u * p);
      // a row is built unless the class says "payload parser decides" and the parser said no (bit 30 of mw)
      const bool act = live[u] && (cls & 1u) && !((cls & 2u) && (mw & ((uint32_t)ALZ_MF_PAYLOAD_REJECT << 24)));
      const bool rv = (cls & 4u) && (mw & ((uint32_t)ALZ_MF_METHOD_MASK << 24)) == (2u << 24);   // DELIVER / PUSHED_EVENT
      const bool err = p == ALZ_PROTO_HTTP && ((mw & 0xFFFFu) - 500u) < 100u;
      key[u] = ((uint64_t)w[u][1] << 32) | w[u][0];               // make_pair_key: the record's first two words as they lie
      const uint32_t bucket = latency_bucket_rz(dur[u]);
      const uint32_t kind = hk ? kPairHost : rv ? kPairRev : kPairFwd;

      // per-CTA table: two index probes, the matching entry names the row and its histogram window
      const uint32_t h = table_hash(key[u]);
      const uint32_t fp = tab_fp(h);
      const uint32_t x1 = s.tab[tab_idx1(h)] ^ fp, x2 = s.tab[tab_idx2(h)] ^ fp;
      const uint32_t x = x1 < 0x10000u ? x1 : x2;                    // upper 16 bits 0: the fingerprint matched
      const uint32_t r = min(x & kRowMask, kRows);
      const uint32_t d = bucket - ((x >> 12) & 15u) * 4u;            // cell of this latency in the row's window
      const bool hit = act && kind == kPairFwd && x < 0x10000u && (x & kRowMask) < kRows && d < 16u && s.rowkey[r] == key[u];
      uint32_t* row = s.rows + r * kRowWords;
      if (hit) {
        const uint32_t sh = (d & 1u) * 16u;
        const uint32_t old = atomicAdd(&row[d >> 1], 1u << sh);
        if (((old >> sh) & 0xFFFFu) == kCellSpill) {
          // this lane took the cell to 0x8000: move 0x8000 counts into the global table (rare: a pair with more
          // than 32767 events in one bucket within one launch of one CTA)
          const uint32_t grow = find_or_insert_pair(pairs, key[u], kPairFwd, ep, ep_mask);
          if (grow < kDropRow) red_add_u32(&pairs.hist[(size_t)grow * ALZ_NB + bucket], 0x8000u);
          else if (grow == kDropRow) unresolved += 0x8000u; else lost += 0x8000u;
          atomicSub(&row[d >> 1], 0x8000u << sh);
        }
        uint32_t* lat = row + 9u + 2u * (lane & 1u);
        const uint32_t lo = (uint32_t)dur[u], dhi = (uint32_t)(dur[u] >> 32);
        const uint32_t oldl = atomicAdd(&lat[0], lo);
        const bool carry = oldl > ~lo;                                   // out of the low word
        if (carry || dhi != 0u) atomicAdd(&lat[1], dhi + (carry ? 1u : 0u));
        if (err) atomicAdd(&row[8], 1u);
        ++n_hit;
      }
      const bool cold = act && !hit;
      meta[u] = bucket | (kind << 6) | (err ? 0x100u : 0u) | (cold ? 0x80000000u : 0u);
      ent[u] = make_uint4(0u, 0u, kNoRow, 0u);
      if (cold) ent[u] = __ldcg(reinterpret_cast<const uint4*>(&pairs.dict_of(kind)[pair_hash(key[u]) & pairs.mask_of(kind)]));
    }
    pc1 = PROF_NOW();
    PROF_ADD(5, pc1 - pc0);
    // cold tier: a home-slot hit is reduced into its row at once, anything else joins the slow queue
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const bool cold = (meta[u] & 0x80000000u) != 0u;
      n_cold += __popc(__ballot_sync(0xFFFFFFFFu, cold));
      const uint32_t bucket = meta[u] & 0x3Fu, kind = (meta[u] >> 6) & 3u;
      const bool err = (meta[u] & 0x100u) != 0u;
      const bool home = cold && ent[u].x == (uint32_t)key[u] && ent[u].y == (uint32_t)(key[u] >> 32) && ent[u].z < kDropRow &&
                        key[u] != kEmptyKey;
      if (home) global_add(pairs, ent[u].z, bucket, dur[u], err);
      const uint32_t dhi = (uint32_t)(dur[u] >> 32);
      const bool huge = dhi >= (1u << 23);                               // does not fit the queue entry: rare beyond words
      if (cold && !home && huge) slow_one<kRows>(key[u], dur[u], bucket, kind, err, pairs, s, ep, ep_mask, &lost, &unresolved);
      slow.push(cold && !home && !huge, key[u], (uint32_t)dur[u], (meta[u] & 0x1FFu) | (dhi << 9), lane_lt);
      __syncwarp();
      if (slow.count >= 32u) {
        const long long s0 = PROF_NOW();
        slow_batch<kRows>(slow, 32u, pairs, s, ep, ep_mask, &lost, &unresolved);
        PROF_ADD(4, PROF_NOW() - s0);
        PROF_ADD(7, 1);
      }
    }
    PROF_ADD(3, PROF_NOW() - pc1);
  }
  while (slow.count) slow_batch<kRows>(slow, min(slow.count, 32u), pairs, s, ep, ep_mask, &lost, &unresolved);
  PROF_ADD(0, PROF_NOW() - p_begin);
  PROF_FLUSH();
  __syncthreads();

  // drain the private rows into the global table. Phase 1, a thread per row: find (or create) the pair's global
  // row — the dependent dictionary probes of up to 32 rows per warp overlap; the row number replaces the key's
  // low word. Phase 2, a half-warp per row: add the window's cells, 5xx count and latency sums.
  const uint32_t used = min(*s.n_rows, kRows);
  for (uint32_t r = threadIdx.x; r < used; r += kWarps * 32) {
    const uint64_t key = s.rowkey[r];
    const uint32_t* row = s.rows + r * kRowWords;
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < (int)kRowWords; ++k) any |= row[k];
    uint32_t grow = kNoRow;                       // never hit in this launch: no global row needed
    if (key != kEmptyKey && any != 0u) grow = find_or_insert_pair(pairs, key, kPairFwd, ep, ep_mask);
    reinterpret_cast<uint32_t*>(&s.rowkey[r])[0] = grow;
  }
  __syncthreads();
  const uint32_t hl = lane & 15u, half = lane >> 4;
  const uint32_t n_pass = (used + kWarps * 2u - 1u) / (kWarps * 2u);     // same trip count for both halves of a warp
  for (uint32_t ps = 0; ps < n_pass; ++ps) {
    const uint32_t r = ps * kWarps * 2u + warp * 2u + half;
    const bool valid = r < used;
    const uint32_t grow = valid ? reinterpret_cast<const uint32_t*>(&s.rowkey[r])[0] : kNoRow;
    const uint32_t* row = s.rows + (valid ? r : kRows) * kRowWords;
    const uint32_t cnt = (row[hl >> 1] >> ((hl & 1u) * 16u)) & 0xFFFFu;
    uint32_t tot = cnt;
    for (int o = 8; o > 0; o >>= 1) tot += __shfl_xor_sync(0xFFFFFFFFu, tot, o);
    if (grow == kNoRow) continue;
    if (grow >= kDropRow) {   // source is not a pod any more (dropped like the reference does) or capacity
      if (hl == 0) { if (grow == kDropRow) unresolved += tot; else lost += tot; }
      continue;
    }
    const uint32_t base = (uint32_t)s.rowbase[r] * 4u;
    if (cnt) red_add_u32(&pairs.hist[(size_t)grow * ALZ_NB + base + hl], cnt);
    if (hl == 0) {
      const uint64_t lat = (((uint64_t)row[10] << 32) + row[9]) + (((uint64_t)row[12] << 32) + row[11]);
      if (lat) red_add_u64(&pairs.lat_sum[grow], lat);
      if (row[8]) red_add_u64(&pairs.err5xx[grow], (uint64_t)row[8]);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    n_hit += __shfl_xor_sync(0xFFFFFFFFu, n_hit, o);
    lost += __shfl_xor_sync(0xFFFFFFFFu, lost, o);
    unresolved += __shfl_xor_sync(0xFFFFFFFFu, unresolved, o);
  }
  // events that built no request row = events seen - hot hits - cold events (the last two are counted anyway)
  const uint32_t not_request = n_live - n_cold - n_hit;
  if (lane == 0) {
    if (not_request) atomicAdd(&ctr->not_request, (unsigned long long)not_request);
    if (lost) atomicAdd(&ctr->capacity_events, (unsigned long long)lost);
    if (unresolved) atomicAdd(&ctr->src_unresolved, (unsigned long long)unresolved);
  }
}

// ---- hot-pair feedback: after a fold, pick the pairs that took the most events -----
// fold_pairs_kernel left row_cnt[row], row_base[row] and a 128-bin (quarter-octave) histogram of the counts of
// the forward rows; thresholds = the lowest bins that keep tier A <= kHotA and A+B <= target. Every block derives
// the same two thresholds from the bins itself (128 adds) instead of a separate launch.
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
      if (p < (uint32_t)kHotA) { hot->keys[p] = pairs.row_key[row]; hot->base[p] = pairs.row_base[row]; }
    } else if (b >= thr_b) {
      const uint32_t p = atomicAdd(&hot->n_b, 1u);
      if (p < (uint32_t)(kHotMax - kHotA)) { hot->keys[kHotA + p] = pairs.row_key[row]; hot->base[kHotA + p] = pairs.row_base[row]; }
    }
  }
}

template <int kWarps, int kRecWords>
void launch_variant(const void* recs, uint64_t n, const AccTable& pairs, Counters* ctr, const HotState* hot,
                    const EpEntry* ep, uint32_t ep_mask, const uint64_t* dur_ovf, int sms, cudaStream_t s) {
  using L = Layout<kWarps>;
  // per device (a process may drive several GPUs), so not cached in a static
  cudaFuncSetAttribute(ingest_pairs_v7_kernel<kWarps, kRecWords>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)L::kBytes);
  ingest_pairs_v7_kernel<kWarps, kRecWords><<<(unsigned)sms, kWarps * 32, L::kBytes, s>>>(
      (const uint32_t*)recs, n, pairs, ctr, hot, ep, ep_mask, dur_ovf);
}

constexpr int kDefaultWarps = 32;

}  // namespace

uint32_t ingest_table_rows() { return Layout<kDefaultWarps>::kRows; }

void launch_ingest_pairs(const alz_l7_rec* recs, uint64_t n, const AccTable& pairs, Counters* ctr,
                         const HotState* hot, const EpEntry* ep, uint32_t ep_mask, int sms, cudaStream_t s) {
  if (n == 0) return;
  // ALZ_INGEST_SHAPE: CTA shape for profiling runs (default = the measured best)
  static const int shape = [] { const char* v = getenv("ALZ_INGEST_SHAPE"); return v ? atoi(v) : 0; }();
  switch (shape) {
    case 1: launch_variant<24, 8>(recs, n, pairs, ctr, hot, ep, ep_mask, nullptr, sms, s); break;
    case 2: launch_variant<16, 8>(recs, n, pairs, ctr, hot, ep, ep_mask, nullptr, sms, s); break;
    default: launch_variant<kDefaultWarps, 8>(recs, n, pairs, ctr, hot, ep, ep_mask, nullptr, sms, s); break;
  }
}

void launch_ingest_pairs_rec16(const alz_l7_rec16* recs, uint64_t n, const uint64_t* dur_ovf, const AccTable& pairs,
                               Counters* ctr, const HotState* hot, const EpEntry* ep, uint32_t ep_mask, int sms,
                               cudaStream_t s) {
  if (n == 0) return;
  launch_variant<kDefaultWarps, 4>(recs, n, pairs, ctr, hot, ep, ep_mask, dur_ovf, sms, s);
}

// after fold_pairs_kernel(pairs, ..., hot->bins): choose next window's hot list
void launch_hot_select(const AccTable& pairs, HotState* hot, int sms, cudaStream_t s) {
  hot_emit_kernel<<<(unsigned)sms * 2, 256, 0, s>>>(pairs, hot, Layout<kDefaultWarps>::kPreload);
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
