// alz_ext_stubs.cu — entry points whose implementation file is not part of this
// build (alaz_b200/build.py defines ALZ_HAVE_<X> when alz_<x>.cu is compiled in).
#include "alz_handle.h"

#ifndef ALZ_HAVE_COMM
int alz_internal_merge_ranks(alz_handle*) { return ALZ_E_UNSUPPORTED; }
void alz_internal_free_comm(alz_handle*) {}
extern "C" int alz_comm_unique_id(void*) { return ALZ_E_UNSUPPORTED; }
extern "C" int alz_comm_init(alz_handle*, int, int, const void*) { return ALZ_E_UNSUPPORTED; }
#endif
#ifndef ALZ_HAVE_GNN
void alz_internal_free_gnn(alz_handle*) {}
extern "C" int alz_gnn_score(alz_handle*, float*, size_t, size_t*) { return ALZ_E_UNSUPPORTED; }
extern "C" int alz_edge_quantiles(const alz_edge_out*, const double*, size_t, double*) { return ALZ_E_UNSUPPORTED; }
#endif
#ifndef ALZ_HAVE_SOCK
void alz_internal_free_sock(alz_handle*) {}
extern "C" int alz_submit_tcp(alz_handle*, const alz_tcp_rec*, size_t) { return ALZ_E_UNSUPPORTED; }
extern "C" int alz_sock_lookup(alz_handle*, const alz_sock_query*, size_t, alz_sock_result*) { return ALZ_E_UNSUPPORTED; }
#endif

void alz_internal_free_extensions(alz_handle* h) {
  alz_internal_free_comm(h);
  alz_internal_free_gnn(h);
  alz_internal_free_sock(h);
}
