/*
 * alz_synth.h — deterministic synthetic l7_req stream (SURVEY.md §8d,
 * docs/SPEC.md §7). Test/bench infrastructure shared by the oracle, the CPU
 * baseline and the GPU bench: event i is a pure function of (topology, i),
 * computed with integer arithmetic only, so gcc on the host and nvcc on the
 * device produce bit-identical records from the same tables.
 *
 * Shape mirrors the reference simulator (main_benchmark_test.go:383-532,
 * testconfig/config1.json): pods = 2 x services, random pod->target edges, a
 * fixed per-event HTTP-like record — but deterministic, skewed (Zipf 1.1 over
 * the edge set) and with the drop/reversal branches exercised.
 */
#ifndef ALZ_SYNTH_H
#define ALZ_SYNTH_H

#include <stdint.h>
#include "../../include/alazgpu.h"

#if defined(__CUDACC__)
#define ALZ_HD __host__ __device__ __forceinline__
#else
#define ALZ_HD static inline
#endif

#define ALZ_SYNTH_LATQ 4096 /* lognormal quantile table has LATQ+1 knots */

/* protocol mixes */
#define ALZ_MIX_SURVEY 0 /* HTTP 90 / REDIS 5 / AMQP 5 */
#define ALZ_MIX_ALL 1    /* every branch of processL7 incl. non-row protocols */

/* read-only view used by the per-event function (host or device pointers) */
typedef struct alz_synth_view {
  uint64_t seed;
  uint64_t t0_ns;
  uint32_t dt_ns;
  uint32_t mix;
  uint32_t n_edges;
  uint32_t n_unknown;      /* pool of source IPs that are in no table */
  uint32_t unknown_base;   /* first IPv4 of that pool */
  uint32_t _pad;
  const uint32_t* edge_saddr;
  const uint32_t* edge_daddr;
  const uint8_t* edge_flags;     /* bit0: anomalous edge */
  const uint32_t* alias_thresh;  /* Vose alias table for Zipf(1.1) over edges */
  const uint32_t* alias_idx;
  const uint64_t* lat_q;         /* LATQ+1 lognormal(ln 2e6, 1) quantiles, ns */
} alz_synth_view;

ALZ_HD uint64_t alz_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

ALZ_HD void alz_synth_event(const alz_synth_view* v, uint64_t i, alz_l7_rec* out) {
  const uint64_t base = alz_splitmix64(v->seed ^ (i * 0xD1342543DE82EF95ull));
  const uint64_t r0 = alz_splitmix64(base + 0);
  const uint64_t r1 = alz_splitmix64(base + 1);
  const uint64_t r2 = alz_splitmix64(base + 2);
  const uint64_t r3 = alz_splitmix64(base + 3);
  const uint64_t r4 = alz_splitmix64(base + 4);

  /* edge ~ Zipf(1.1) by alias sampling, integer only */
  uint32_t j = (uint32_t)(((r0 >> 32) * (uint64_t)v->n_edges) >> 32);
  uint32_t e = ((uint32_t)r0 < v->alias_thresh[j]) ? j : v->alias_idx[j];
  uint32_t saddr = v->edge_saddr[e];
  uint32_t daddr = v->edge_daddr[e];
  const uint32_t anomalous = v->edge_flags[e] & 1u;

  /* 0.1 % saddr == 0 (get_sock miss, ebpf/c/l7.c:313-314); 2 % unknown source
   * (drop rule, aggregator/data.go:829-832) */
  uint32_t fate = (uint32_t)(r1 % 100000u);
  if (fate < 100u) saddr = 0u;
  else if (fate < 2100u) saddr = v->unknown_base + (uint32_t)((r1 >> 32) % v->n_unknown);

  /* protocol / method / tls */
  uint32_t y = (uint32_t)(r2 % 1000u);
  uint32_t w = (uint32_t)(r2 >> 32);
  uint32_t proto, method = 0, flags = 0, dport;
  if (v->mix == ALZ_MIX_SURVEY) {
    proto = (y < 900u) ? ALZ_PROTO_HTTP : (y < 950u) ? ALZ_PROTO_REDIS : ALZ_PROTO_AMQP;
  } else {
    proto = (y < 700u) ? ALZ_PROTO_HTTP : (y < 760u) ? ALZ_PROTO_REDIS
          : (y < 820u) ? ALZ_PROTO_AMQP : (y < 860u) ? ALZ_PROTO_POSTGRES
          : (y < 890u) ? ALZ_PROTO_MYSQL : (y < 920u) ? ALZ_PROTO_MONGO
          : (y < 950u) ? ALZ_PROTO_HTTP2 : (y < 980u) ? ALZ_PROTO_KAFKA
          : ALZ_PROTO_UNKNOWN;
  }
  switch (proto) {
    case ALZ_PROTO_HTTP:
      method = (w % 10u < 7u) ? 1u : 1u + (w % 9u);
      if ((w >> 16) % 10u < 3u) flags |= ALZ_MF_TLS;
      dport = (flags & ALZ_MF_TLS) ? 443u : 80u;
      break;
    case ALZ_PROTO_REDIS:
      method = (w % 20u == 0u) ? ALZ_REDIS_PUSHED_EVENT
             : (w % 20u < 17u) ? ALZ_REDIS_COMMAND : ALZ_REDIS_PING;
      dport = 6379u;
      break;
    case ALZ_PROTO_AMQP:
      method = (w & 1u) ? ALZ_AMQP_DELIVER : ALZ_AMQP_PUBLISH;
      dport = 5672u;
      break;
    case ALZ_PROTO_POSTGRES:
      method = 2u + (w & 1u);
      if ((w >> 8) % 10u == 0u) flags |= ALZ_MF_PAYLOAD_REJECT;
      dport = 5432u;
      break;
    case ALZ_PROTO_MYSQL:
      method = 1u + (w & 3u);
      if (method == 1u && (w >> 8) % 10u == 0u) flags |= ALZ_MF_PAYLOAD_REJECT;
      dport = 3306u;
      break;
    case ALZ_PROTO_MONGO:
      if ((w >> 8) % 20u == 0u) flags |= ALZ_MF_PAYLOAD_REJECT;
      dport = 27017u;
      break;
    case ALZ_PROTO_HTTP2: method = 1u + (w & 1u); dport = 8080u; break;
    case ALZ_PROTO_KAFKA: method = 1u + (w & 1u); dport = 9092u; break;
    default: dport = 9u; break;
  }

  /* status + duration */
  uint32_t z = (uint32_t)(r3 % 100u);
  uint32_t status;
  if (proto == ALZ_PROTO_HTTP) {
    if (anomalous) status = (z < 80u) ? 200u : ((z & 1u) ? 500u : 503u);
    else status = (z < 94u) ? 200u : (z < 97u) ? 404u : (z < 99u) ? 500u : 503u;
  } else {
    status = (z == 0u) ? 2u : 1u;
  }
  uint32_t q = (uint32_t)(r3 >> 32);
  uint32_t qi = q >> 20, qf = q & 0xFFFFFu;
  uint64_t lo = v->lat_q[qi], hi = v->lat_q[qi + 1];
  uint64_t dur = lo + (((hi - lo) * (uint64_t)qf) >> 20);
  if (anomalous) dur *= 10u;

  out->saddr = saddr;
  out->daddr = daddr;
  out->sport = (uint16_t)(32768u + (uint32_t)(r4 % 28232u));
  out->dport = (uint16_t)dport;
  out->status = (uint16_t)status;
  out->protocol = (uint8_t)proto;
  out->method_flags = (uint8_t)(method | flags);
  out->duration_ns = dur;
  out->write_time_ns = v->t0_ns + i * (uint64_t)v->dt_ns + ((r4 >> 32) % v->dt_ns);
}

/* ---- host-side topology builder (alz_synth_topo.c) ------------------------- */
#ifdef __cplusplus
extern "C" {
#endif

typedef struct alz_synth_topo {
  uint32_t n_services, n_pods, n_edges, n_outbound;
  uint32_t* pod_ip;   /* pod id k  -> IPv4 (10.0.0.0/8) */
  uint32_t* svc_ip;   /* svc id j  -> ClusterIP (172.16.0.0/12) */
  uint32_t* out_ip;   /* outbound hosts, in no table (203.0.0.0/8) */
  uint32_t* edge_saddr;
  uint32_t* edge_daddr;
  uint8_t* edge_flags;
  uint32_t* alias_thresh;
  uint32_t* alias_idx;
  uint64_t* lat_q;
  alz_synth_view view; /* host pointers into the arrays above */
} alz_synth_topo;

/* S services, P = 2S pods, E = 10S edges (80 % ->svc, 15 % ->pod, 5 % ->outbound),
 * 1 % anomalous edges. Returns NULL on allocation failure. */
alz_synth_topo* alz_synth_topo_create(uint32_t n_services, uint64_t seed, uint32_t mix);
void alz_synth_topo_destroy(alz_synth_topo* t);
/* fill recs[0..n) with events first..first+n (host) */
void alz_synth_fill(const alz_synth_topo* t, uint64_t first, uint64_t n, alz_l7_rec* recs);

#ifdef __cplusplus
}
#endif
#endif /* ALZ_SYNTH_H */
