// alz_device.cuh — device-side building blocks of libalazgpu (sm_100a).
//
// Semantics restated from the reference (file:line = getanteon/alaz @ 828b997f):
//   row-emit rule      aggregator/data.go:1364-1383 (processL7 switch), :1252-1255,
//                      :1288-1292, :1328-1332 (payload parse failure => no row)
//   direction reversal aggregator/data.go:1110-1112 (AMQP DELIVER), :1151-1153
//                      (REDIS PUSHED_EVENT); datastore/dto.go:246-251
//   resolve            aggregator/data.go:827-870 (setFromToV2)
// Everything else (edge key, buckets, accumulators) is docs/SPEC.md.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/alazgpu.h"

namespace alz {

constexpr uint64_t kEmptyKey = ~0ull;   // open-addressing empty marker (u64 tables)
constexpr uint32_t kMaxProbe = 4096;    // linear-probe bound before "capacity"

// ---- endpoint table: the join's build side, one entry per IPv4 ---------------
// (ClusterInfo.PodIPToPodUid + ServiceIPToServiceUid, aggregator/cluster.go:15-16,
// merged: one probe answers "is it a pod" and "is it a service")
struct __align__(16) EpEntry {
  uint32_t ip;
  uint32_t state;  // bit0 occupied, bit1 has pod id, bit2 has service id
  uint32_t pod;
  uint32_t svc;
};
constexpr uint32_t kEpOcc = 1u, kEpPod = 2u, kEpSvc = 4u;

__host__ __device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint64_t hash64(uint64_t x) {
  x ^= x >> 33; x *= 0xFF51AFD7ED558CCDull; x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ull; x ^= x >> 33;
  return x;
}

// ---- packed edge key (docs/SPEC.md §3) -----------------------------------------
// One end of every edge is the pod that setFromToV2 resolved from saddr.
//   bit 63      rev: 0 = pod is From, 1 = pod is To (row was reversed)
//   bits 62..61 type of the other end (ALZ_NODE_*)
//   bits 60..32 pod id (< 2^29)
//   bits 31..0  other end: pod/service id, or raw IPv4 for outbound
// Canonical form: a reversed pod<->pod row is stored with rev = 0 and the ends
// swapped, so equal (From,To) always give equal keys.
__host__ __device__ __forceinline__ uint64_t make_edge_key(uint32_t pod_id, uint32_t other_type,
                                                           uint32_t other, bool rev) {
  if (rev && other_type == ALZ_NODE_POD) { uint32_t t = pod_id; pod_id = other; other = t; rev = false; }
  return ((uint64_t)(rev ? 1u : 0u) << 63) | ((uint64_t)other_type << 61) |
         ((uint64_t)(pod_id & 0x1FFFFFFFu) << 32) | (uint64_t)other;
}
__host__ __device__ __forceinline__ void unpack_edge_key(uint64_t k, uint8_t* from_type, uint32_t* from,
                                                         uint8_t* to_type, uint32_t* to) {
  const bool rev = (k >> 63) != 0;
  const uint8_t ot = (uint8_t)((k >> 61) & 3u);
  const uint32_t pod = (uint32_t)((k >> 32) & 0x1FFFFFFFu);
  const uint32_t other = (uint32_t)k;
  if (!rev) { *from_type = ALZ_NODE_POD; *from = pod; *to_type = ot; *to = other; }
  else      { *from_type = ot; *from = other; *to_type = ALZ_NODE_POD; *to = pod; }
}

// ---- per-record rules ------------------------------------------------------------
// Does processL7 hand a row to PersistRequest for this record (before resolve)?
__host__ __device__ __forceinline__ bool emits_request_row(uint32_t protocol, uint32_t method_flags) {
  switch (protocol) {
    case ALZ_PROTO_HTTP: case ALZ_PROTO_AMQP: case ALZ_PROTO_REDIS:
      return true;
    case ALZ_PROTO_POSTGRES: case ALZ_PROTO_MYSQL: case ALZ_PROTO_MONGO:
      return (method_flags & ALZ_MF_PAYLOAD_REJECT) == 0;
    default:  // HTTP2 (frame pairing), KAFKA (PersistKafkaEvent), UNKNOWN
      return false;
  }
}
__host__ __device__ __forceinline__ bool is_reversed(uint32_t protocol, uint32_t method_flags) {
  const uint32_t m = method_flags & ALZ_MF_METHOD_MASK;
  return (protocol == ALZ_PROTO_AMQP && m == ALZ_AMQP_DELIVER) ||
         (protocol == ALZ_PROTO_REDIS && m == ALZ_REDIS_PUSHED_EVENT);
}
// Protocol is "HTTP" or "HTTPS" (HTTPS = HTTP && tls, data.go:1240-1242)
__host__ __device__ __forceinline__ bool is_5xx(uint32_t protocol, uint32_t status) {
  return protocol == ALZ_PROTO_HTTP && status >= 500u && status < 600u;
}
// docs/SPEC.md §4: two sub-buckets per octave over [2^8, 2^40), clamped
__device__ __forceinline__ uint32_t latency_bucket(uint64_t d) {
  if (d < 256ull) return 0u;
  const uint32_t o = 63u - (uint32_t)__clzll((long long)d);
  if (o >= 40u) return ALZ_NB - 1u;
  return 2u * (o - 8u) + (uint32_t)((d >> (o - 1u)) & 1ull);
}

// ---- open-addressed u64 dictionary with accumulator rows ---------------------------
struct AccTable {
  uint64_t* keys;     // [cap + 1]; row `cap` is the overflow/sentinel row
  uint64_t* lat_sum;  // [cap + 1]
  uint64_t* err5xx;   // [cap + 1]
  uint64_t* count;    // [cap + 1] (edge tables only; pair tables derive it from hist)
  uint32_t* hist;     // [(cap + 1) * ALZ_NB]
  uint32_t cap;       // power of two
};

// returns the row of `key`, inserting it if absent; cap = table full (counted by caller)
__device__ __forceinline__ uint32_t find_or_insert(const AccTable& t, uint64_t key, uint32_t* inserted) {
  if (key == kEmptyKey) return t.cap;  // the one key that collides with the marker
  const uint32_t mask = t.cap - 1u;
  uint32_t slot = (uint32_t)hash64(key) & mask;
  for (uint32_t p = 0; p < kMaxProbe; ++p) {
    uint64_t k = __ldcg(&t.keys[slot]);
    if (k == key) return slot;
    if (k == kEmptyKey) {
      const uint64_t old = atomicCAS((unsigned long long*)&t.keys[slot], (unsigned long long)kEmptyKey,
                                     (unsigned long long)key);
      if (old == kEmptyKey) { if (inserted) *inserted += 1u; return slot; }
      if (old == key) return slot;
    }
    slot = (slot + 1u) & mask;
  }
  return 0xFFFFFFFFu;
}

// probe the endpoint table; returns state bits (0 = absent) and ids
__device__ __forceinline__ uint32_t ep_lookup(const EpEntry* __restrict__ tab, uint32_t mask, uint32_t ip,
                                              uint32_t* pod, uint32_t* svc) {
  uint32_t slot = hash32(ip) & mask;
  for (;;) {
    const uint4 e = __ldg(reinterpret_cast<const uint4*>(&tab[slot]));
    if ((e.y & kEpOcc) == 0u) return 0u;
    if (e.x == ip) { *pod = e.z; *svc = e.w; return e.y; }
    slot = (slot + 1u) & mask;
  }
}

// setFromToV2 on integers: false = "error finding pod with sockets saddr" (drop)
__device__ __forceinline__ bool resolve_edge(const EpEntry* __restrict__ tab, uint32_t mask, uint32_t saddr,
                                             uint32_t daddr, bool rev, uint64_t* edge_key) {
  uint32_t pod, svc;
  const uint32_t s = ep_lookup(tab, mask, saddr, &pod, &svc);
  if ((s & kEpPod) == 0u) return false;                 // data.go:829-832
  const uint32_t from_pod = pod;
  const uint32_t d = ep_lookup(tab, mask, daddr, &pod, &svc);
  uint32_t ot, ov;
  if (d & kEpSvc) { ot = ALZ_NODE_SVC; ov = svc; }       // :840-843 service first
  else if (d & kEpPod) { ot = ALZ_NODE_POD; ov = pod; }  // :845-849
  else { ot = ALZ_NODE_OUTBOUND; ov = daddr; }           // :862 raw daddr
  *edge_key = make_edge_key(from_pod, ot, ov, rev);
  return true;
}

// 32-byte record, one 256-bit load (LDG.E.256 on sm_100a), streaming: bypass L1
struct Rec { uint32_t w[8]; };
__device__ __forceinline__ Rec load_rec(const alz_l7_rec* p) {
  Rec r;
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]),
                 "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
               : "l"(p));
  return r;
}
// field accessors of the packed words (layout of alz_l7_rec)
__device__ __forceinline__ uint32_t rec_saddr(const Rec& r) { return r.w[0]; }
__device__ __forceinline__ uint32_t rec_daddr(const Rec& r) { return r.w[1]; }
__device__ __forceinline__ uint32_t rec_status(const Rec& r) { return r.w[3] & 0xFFFFu; }
__device__ __forceinline__ uint32_t rec_protocol(const Rec& r) { return (r.w[3] >> 16) & 0xFFu; }
__device__ __forceinline__ uint32_t rec_mflags(const Rec& r) { return r.w[3] >> 24; }
__device__ __forceinline__ uint64_t rec_duration(const Rec& r) { return ((uint64_t)r.w[5] << 32) | r.w[4]; }

// device-side counters (one cache line per handle)
struct Counters {
  unsigned long long not_request;
  unsigned long long src_unresolved;
  unsigned long long pairs_inserted;
  unsigned long long edges_inserted;
  unsigned long long capacity_events;  // events/pairs lost to a full dictionary
  unsigned long long n_live;           // scratch for compaction
  unsigned long long pad[2];
};

}  // namespace alz
