/*
 * alz_synth_topo.c — host-side builder of the synthetic cluster topology and
 * the sampling tables used by alz_synth_event (alz_synth.h). Test/bench
 * infrastructure; shape after the reference simulator
 * (main_benchmark_test.go:383-532: pods, services, random pod->svc edges).
 */
#include "alz_synth.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static uint32_t perm_bits(uint32_t k, uint32_t bits) {
  /* bijection on [0, 2^bits): odd multiply, xorshift, odd multiply */
  const uint32_t mask = (bits >= 32) ? 0xFFFFFFFFu : ((1u << bits) - 1u);
  uint32_t x = (k * 2654435761u) & mask;
  x ^= x >> (bits / 2);
  x = (x * 0x5BD1E995u) & mask;
  return x;
}

/* Acklam's rational approximation of the standard normal quantile */
static double norm_ppf(double p) {
  static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02,
                             -2.759285104469687e+02, 1.383577518672690e+02,
                             -3.066479806614716e+01, 2.506628277459239e+00};
  static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02,
                             -1.556989798598866e+02, 6.680131188771972e+01,
                             -1.328068155288572e+01};
  static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01,
                             -2.400758277161838e+00, -2.549732539343734e+00,
                             4.374664141464968e+00, 2.938163982698783e+00};
  static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01,
                             2.445134137142996e+00, 3.754408661907416e+00};
  const double plow = 0.02425, phigh = 1 - plow;
  double q, r;
  if (p < plow) {
    q = sqrt(-2 * log(p));
    return (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
           ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
  }
  if (p > phigh) {
    q = sqrt(-2 * log(1 - p));
    return -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
           ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
  }
  q = p - 0.5;
  r = q * q;
  return (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
         (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1);
}

/* Vose alias method for weights w[k] = (k+1)^-1.1 */
static int build_alias(uint32_t n, uint32_t* thresh, uint32_t* alias) {
  double* p = (double*)malloc(sizeof(double) * n);
  uint32_t* small = (uint32_t*)malloc(sizeof(uint32_t) * n);
  uint32_t* large = (uint32_t*)malloc(sizeof(uint32_t) * n);
  if (!p || !small || !large) { free(p); free(small); free(large); return -1; }
  double sum = 0;
  for (uint32_t k = 0; k < n; k++) { p[k] = pow((double)k + 1.0, -1.1); sum += p[k]; }
  uint32_t ns = 0, nl = 0;
  for (uint32_t k = 0; k < n; k++) {
    p[k] = p[k] * (double)n / sum;
    if (p[k] < 1.0) small[ns++] = k; else large[nl++] = k;
  }
  for (uint32_t k = 0; k < n; k++) { thresh[k] = 0xFFFFFFFFu; alias[k] = k; }
  while (ns > 0 && nl > 0) {
    uint32_t s = small[--ns], l = large[--nl];
    double t = p[s] * 4294967296.0;
    thresh[s] = (t >= 4294967295.0) ? 0xFFFFFFFFu : (uint32_t)t;
    alias[s] = l;
    p[l] = (p[l] + p[s]) - 1.0;
    if (p[l] < 1.0) small[ns++] = l; else large[nl++] = l;
  }
  free(p); free(small); free(large);
  return 0;
}

alz_synth_topo* alz_synth_topo_create(uint32_t S, uint64_t seed, uint32_t mix) {
  if (S == 0 || S > (1u << 20)) return NULL;
  alz_synth_topo* t = (alz_synth_topo*)calloc(1, sizeof(*t));
  if (!t) return NULL;
  const uint32_t P = 2 * S, E = 10 * S;
  const uint32_t O = (S / 10 < 16) ? 16 : S / 10;
  t->n_services = S; t->n_pods = P; t->n_edges = E; t->n_outbound = O;
  t->pod_ip = (uint32_t*)malloc(sizeof(uint32_t) * P);
  t->svc_ip = (uint32_t*)malloc(sizeof(uint32_t) * S);
  t->out_ip = (uint32_t*)malloc(sizeof(uint32_t) * O);
  t->edge_saddr = (uint32_t*)malloc(sizeof(uint32_t) * E);
  t->edge_daddr = (uint32_t*)malloc(sizeof(uint32_t) * E);
  t->edge_flags = (uint8_t*)malloc(E);
  t->alias_thresh = (uint32_t*)malloc(sizeof(uint32_t) * E);
  t->alias_idx = (uint32_t*)malloc(sizeof(uint32_t) * E);
  t->lat_q = (uint64_t*)malloc(sizeof(uint64_t) * (ALZ_SYNTH_LATQ + 1));
  if (!t->pod_ip || !t->svc_ip || !t->out_ip || !t->edge_saddr || !t->edge_daddr ||
      !t->edge_flags || !t->alias_thresh || !t->alias_idx || !t->lat_q) {
    alz_synth_topo_destroy(t);
    return NULL;
  }
  for (uint32_t k = 0; k < P; k++) t->pod_ip[k] = 0x0A000000u | perm_bits(k, 24);
  for (uint32_t j = 0; j < S; j++) t->svc_ip[j] = 0xAC100000u | perm_bits(j, 20);
  for (uint32_t m = 0; m < O; m++) t->out_ip[m] = 0xCB000000u | perm_bits(m, 24);

  for (uint32_t e = 0; e < E; e++) {
    uint64_t h1 = alz_splitmix64(seed * 3 + 0x1000000000ull + e);
    uint64_t h2 = alz_splitmix64(h1);
    uint64_t h3 = alz_splitmix64(h2);
    uint32_t src = (uint32_t)(h1 % P);
    uint32_t tsel = (uint32_t)((h1 >> 32) % 100u);
    uint32_t dst_ip;
    if (tsel < 80u) dst_ip = t->svc_ip[h2 % S];
    else if (tsel < 95u) {
      uint32_t dp = (uint32_t)(h2 % P);
      if (dp == src) dp = (dp + 1) % P;
      dst_ip = t->pod_ip[dp];
    } else dst_ip = t->out_ip[h2 % O];
    t->edge_saddr[e] = t->pod_ip[src];
    t->edge_daddr[e] = dst_ip;
    t->edge_flags[e] = (uint8_t)((h3 % 100u) == 0u);
  }
  if (build_alias(E, t->alias_thresh, t->alias_idx) != 0) {
    alz_synth_topo_destroy(t);
    return NULL;
  }
  const double mu = log(2e6), sigma = 1.0;
  uint64_t prev = 1;
  for (uint32_t j = 0; j <= ALZ_SYNTH_LATQ; j++) {
    double p = ((double)j + 0.5) / (double)(ALZ_SYNTH_LATQ + 1);
    double x = exp(mu + sigma * norm_ppf(p));
    uint64_t qv = (uint64_t)llround(x);
    if (qv < prev) qv = prev;
    t->lat_q[j] = qv;
    prev = qv;
  }
  alz_synth_view* v = &t->view;
  v->seed = seed;
  v->t0_ns = 1000000000000ull;
  v->dt_ns = 100;
  v->mix = mix;
  v->n_edges = E;
  v->n_unknown = (S / 10 < 16) ? 16 : S / 10;
  v->unknown_base = 0x64400000u; /* 100.64.0.0/10: in no table */
  v->edge_saddr = t->edge_saddr;
  v->edge_daddr = t->edge_daddr;
  v->edge_flags = t->edge_flags;
  v->alias_thresh = t->alias_thresh;
  v->alias_idx = t->alias_idx;
  v->lat_q = t->lat_q;
  return t;
}

void alz_synth_topo_destroy(alz_synth_topo* t) {
  if (!t) return;
  free(t->pod_ip); free(t->svc_ip); free(t->out_ip);
  free(t->edge_saddr); free(t->edge_daddr); free(t->edge_flags);
  free(t->alias_thresh); free(t->alias_idx); free(t->lat_q);
  free(t);
}

void alz_synth_fill(const alz_synth_topo* t, uint64_t first, uint64_t n, alz_l7_rec* recs) {
  for (uint64_t i = 0; i < n; i++) alz_synth_event(&t->view, first + i, &recs[i]);
}
