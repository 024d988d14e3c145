/*
 * alz_fastcpu.c — the "fair" CPU arm: the same resolve/emit/reduce semantics
 * as alz_oracle.c, written the way one would write it for speed on host cores — integer keys, flat
 * open-addressed tables, per-thread accumulators, threads pinned one per core, parallel merge. TEST/BENCH
 * INFRASTRUCTURE ONLY (same rules as alz_oracle.h): it exists so that bench.py's cpu_baseline is not only
 * the allocation-heavy restatement of the Go data structures. tests/test_oracle.py holds it bit-exact to the
 * faithful restatement; the product never links it.
 *
 * Semantics restated (file:line = getanteon/alaz @ 828b997f): processL7 switch aggregator/data.go:1364-1383;
 * payload-parse drops :1252-1255, :1288-1292, :1328-1332; setFromToV2 :827-870; ReverseDirection
 * datastore/dto.go:246-251 (:1110-1112, :1151-1153); edge/accumulators docs/SPEC.md §3-§4.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>

#include "alz_oracle.h"

typedef struct fep { uint32_t ip, state, pod, svc; } fep; /* state: 1 occupied, 2 pod, 4 svc */
typedef struct facc {
  uint64_t key; /* packed edge key, SPEC §3; ~0 = empty */
  uint64_t count, err5xx, lat_sum;
  uint32_t hist[ALZ_NB];
} facc;
typedef struct fmap { facc* e; size_t cap, len; } fmap;

struct orc_fast {
  fep* tab; uint32_t mask; size_t n_ep;
  fmap result;
  alz_stats st;
};

static uint32_t h32(uint32_t x) { x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16; return x; }
static uint64_t h64(uint64_t x) { x ^= x >> 33; x *= 0xFF51AFD7ED558CCDull; x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ull; x ^= x >> 33; return x; }

static void fmap_init(fmap* m, size_t cap) {
  m->cap = cap; m->len = 0;
  m->e = (facc*)malloc(cap * sizeof(facc));
  for (size_t i = 0; i < cap; i++) m->e[i].key = ~0ull;
}
static facc* fmap_get(fmap* m, uint64_t key);
static void fmap_grow(fmap* m) {
  fmap old = *m;
  fmap_init(m, old.cap * 2);
  for (size_t i = 0; i < old.cap; i++)
    if (old.e[i].key != ~0ull) { facc* d = fmap_get(m, old.e[i].key); *d = old.e[i]; }
  free(old.e);
}
static facc* fmap_get(fmap* m, uint64_t key) { /* find or insert (zeroed) */
  if ((m->len + 1) * 2 > m->cap) fmap_grow(m);
  size_t i = (size_t)h64(key) & (m->cap - 1);
  for (;;) {
    facc* a = &m->e[i];
    if (a->key == key) return a;
    if (a->key == ~0ull) { memset(a, 0, sizeof *a); a->key = key; m->len++; return a; }
    i = (i + 1) & (m->cap - 1);
  }
}

orc_fast* orc_fast_create(uint32_t max_endpoints) {
  orc_fast* f = (orc_fast*)calloc(1, sizeof *f);
  uint32_t cap = 1024;
  while (cap < 2u * max_endpoints) cap <<= 1;
  f->tab = (fep*)calloc(cap, sizeof(fep));
  f->mask = cap - 1;
  fmap_init(&f->result, 1024);
  return f;
}
void orc_fast_destroy(orc_fast* f) { if (!f) return; free(f->tab); free(f->result.e); free(f); }

/* ADD/UPDATE only (the bench loads the tables once); same overwrite rule as persist.go:55-65, :114-124 */
void orc_fast_table_upsert(orc_fast* f, int table, uint32_t ip, uint32_t id) {
  uint32_t i = h32(ip) & f->mask;
  while ((f->tab[i].state & 1u) && f->tab[i].ip != ip) i = (i + 1) & f->mask;
  f->tab[i].ip = ip; f->tab[i].state |= 1u;
  if (table == ALZ_TABLE_POD) { f->tab[i].state |= 2u; f->tab[i].pod = id; }
  else { f->tab[i].state |= 4u; f->tab[i].svc = id; }
}
static const fep* ep_find(const orc_fast* f, uint32_t ip) {
  uint32_t i = h32(ip) & f->mask;
  for (;;) {
    const fep* e = &f->tab[i];
    if (!(e->state & 1u)) return NULL;
    if (e->ip == ip) return e;
    i = (i + 1) & f->mask;
  }
}

/* packed edge key, docs/SPEC.md §3 */
static uint64_t edge_key(uint32_t pod, uint32_t ot, uint32_t ov, int rev) {
  if (rev && ot == ALZ_NODE_POD) { uint32_t t = pod; pod = ov; ov = t; rev = 0; }
  return ((uint64_t)(rev ? 1u : 0u) << 63) | ((uint64_t)ot << 61) | ((uint64_t)(pod & 0x1FFFFFFFu) << 32) | ov;
}

typedef struct fworker {
  const orc_fast* f; const alz_l7_rec* recs; size_t n;
  fmap local; alz_stats st;
  struct fworker* all; int nworkers, id; fmap* shard_out;
  uint32_t* order;   /* slots of the live entries of `local`, grouped by the merge thread that owns the key */
  size_t* group;     /* [nworkers + 1] offsets into order */
} fworker;

static int owner_of(uint64_t key, int nworkers) { return (int)((h64(key) >> 40) % (uint64_t)nworkers); }

static void pin_to(int cpu) {
  cpu_set_t s; CPU_ZERO(&s); CPU_SET(cpu % CPU_SETSIZE, &s);
  pthread_setaffinity_np(pthread_self(), sizeof s, &s); /* best effort */
}

static void process_range(const orc_fast* f, const alz_l7_rec* recs, size_t n, fmap* m, alz_stats* st) {
  for (size_t i = 0; i < n; i++) {
    const alz_l7_rec* d = &recs[i];
    st->events_in++;
    const uint32_t p = d->protocol, mf = d->method_flags;
    const int row = p <= 8u && ((0x1AEu >> p) & 1u);
    const int sql = p <= 8u && ((0x188u >> p) & 1u);
    if (!row || (sql && (mf & ALZ_MF_PAYLOAD_REJECT))) { st->not_request++; continue; }
    const fep* s = ep_find(f, d->saddr);
    if (!s || !(s->state & 2u)) { st->src_unresolved++; continue; }
    const fep* t = ep_find(f, d->daddr);
    uint32_t ot, ov;
    if (t && (t->state & 4u)) { ot = ALZ_NODE_SVC; ov = t->svc; }
    else if (t && (t->state & 2u)) { ot = ALZ_NODE_POD; ov = t->pod; }
    else { ot = ALZ_NODE_OUTBOUND; ov = d->daddr; }
    const int rev = (mf & ALZ_MF_METHOD_MASK) == 2u && (p == ALZ_PROTO_AMQP || p == ALZ_PROTO_REDIS);
    facc* a = fmap_get(m, edge_key(s->pod, ot, ov, rev));
    a->count++;
    if (p == ALZ_PROTO_HTTP && d->status >= 500 && d->status < 600) a->err5xx++;
    a->lat_sum += d->duration_ns;
    a->hist[orc_bucket(d->duration_ns)]++;
    st->rows_emitted++;
  }
}
static void* fwork_main(void* p) {
  fworker* w = (fworker*)p;
  pin_to(w->id);
  process_range(w->f, w->recs, w->n, &w->local, &w->st);
  /* hand-over for the merge: a counting sort of the live slots by owner, so that every merge thread reads only
   * the entries it owns (scanning all maps in every merge thread is quadratic in the thread count) */
  const int T = w->nworkers;
  w->group = (size_t*)calloc((size_t)T + 1, sizeof(size_t));
  w->order = (uint32_t*)malloc((w->local.len ? w->local.len : 1) * sizeof(uint32_t));
  for (size_t i = 0; i < w->local.cap; i++)
    if (w->local.e[i].key != ~0ull) w->group[owner_of(w->local.e[i].key, T) + 1]++;
  for (int t = 0; t < T; t++) w->group[t + 1] += w->group[t];
  size_t* at = (size_t*)malloc((size_t)T * sizeof(size_t));
  memcpy(at, w->group, (size_t)T * sizeof(size_t));
  for (size_t i = 0; i < w->local.cap; i++)
    if (w->local.e[i].key != ~0ull) w->order[at[owner_of(w->local.e[i].key, T)]++] = (uint32_t)i;
  free(at);
  return NULL;
}
/* phase 2: thread id merges the keys it owns out of every worker's map */
static void* fmerge_main(void* p) {
  fworker* w = (fworker*)p;
  pin_to(w->id);
  fmap* out = w->shard_out;
  for (int t = 0; t < w->nworkers; t++) {
    const fworker* src = &w->all[t];
    for (size_t k = src->group[w->id]; k < src->group[w->id + 1]; k++) {
      const facc* a = &src->local.e[src->order[k]];
      facc* d = fmap_get(out, a->key);
      d->count += a->count; d->err5xx += a->err5xx; d->lat_sum += a->lat_sum;
      for (int b = 0; b < ALZ_NB; b++) d->hist[b] += a->hist[b];
    }
  }
  return NULL;
}

void orc_fast_process(orc_fast* f, const alz_l7_rec* recs, size_t n, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  fworker* w = (fworker*)calloc((size_t)nthreads, sizeof(fworker));
  pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
  fmap* shards = (fmap*)calloc((size_t)nthreads, sizeof(fmap));
  const size_t per = (n + (size_t)nthreads - 1) / (size_t)nthreads;
  for (int t = 0; t < nthreads; t++) {
    size_t b = per * (size_t)t, e = b + per; if (b > n) b = n; if (e > n) e = n;
    w[t].f = f; w[t].recs = recs + b; w[t].n = e - b; w[t].all = w; w[t].nworkers = nthreads; w[t].id = t;
    fmap_init(&w[t].local, 4096);
    fmap_init(&shards[t], 1024);
    w[t].shard_out = &shards[t];
    pthread_create(&th[t], NULL, fwork_main, &w[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, fmerge_main, &w[t]);
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  for (int t = 0; t < nthreads; t++) {
    for (size_t i = 0; i < shards[t].cap; i++) {
      const facc* a = &shards[t].e[i];
      if (a->key == ~0ull) continue;
      facc* d = fmap_get(&f->result, a->key);
      d->count += a->count; d->err5xx += a->err5xx; d->lat_sum += a->lat_sum;
      for (int b = 0; b < ALZ_NB; b++) d->hist[b] += a->hist[b];
    }
    free(shards[t].e); free(w[t].local.e); free(w[t].order); free(w[t].group);
    f->st.events_in += w[t].st.events_in; f->st.rows_emitted += w[t].st.rows_emitted;
    f->st.not_request += w[t].st.not_request; f->st.src_unresolved += w[t].st.src_unresolved;
  }
  free(shards); free(w); free(th);
}

static int fedge_cmp(const void* pa, const void* pb) {
  const alz_edge_out* a = (const alz_edge_out*)pa; const alz_edge_out* b = (const alz_edge_out*)pb;
  if (a->from_type != b->from_type) return a->from_type < b->from_type ? -1 : 1;
  if (a->from != b->from) return a->from < b->from ? -1 : 1;
  if (a->to_type != b->to_type) return a->to_type < b->to_type ? -1 : 1;
  if (a->to != b->to) return a->to < b->to ? -1 : 1;
  return 0;
}
size_t orc_fast_edges(orc_fast* f, alz_edge_out* out, size_t cap) {
  size_t n = 0;
  for (size_t i = 0; i < f->result.cap; i++) {
    const facc* a = &f->result.e[i];
    if (a->key == ~0ull) continue;
    if (n < cap) {
      alz_edge_out* r = &out[n]; memset(r, 0, sizeof *r);
      const int rev = (int)(a->key >> 63);
      const uint8_t ot = (uint8_t)((a->key >> 61) & 3u);
      const uint32_t pod = (uint32_t)((a->key >> 32) & 0x1FFFFFFFu), ov = (uint32_t)a->key;
      if (!rev) { r->from_type = ALZ_NODE_POD; r->from = pod; r->to_type = ot; r->to = ov; }
      else { r->from_type = ot; r->from = ov; r->to_type = ALZ_NODE_POD; r->to = pod; }
      r->count = a->count; r->err5xx = a->err5xx; r->lat_sum_ns = a->lat_sum;
      memcpy(r->hist, a->hist, sizeof r->hist);
    }
    n++;
  }
  qsort(out, n < cap ? n : cap, sizeof(alz_edge_out), fedge_cmp);
  return n;
}
void orc_fast_reset(orc_fast* f) {
  free(f->result.e); fmap_init(&f->result, 1024); memset(&f->st, 0, sizeof f->st);
}
void orc_fast_stats(orc_fast* f, alz_stats* st) { *st = f->st; st->edges_live = f->result.len; }
