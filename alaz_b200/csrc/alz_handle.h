// alz_handle.h — the handle behind the C ABI: everything one GPU owns.
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <cuda_runtime.h>

#include "../../include/alazgpu.h"
#include "../../include/alazgpu_synth.h"
#include "alz_kernels.cuh"

struct HostEp {
  uint32_t state = 0;  // kEpPod | kEpSvc
  uint32_t pod = 0, svc = 0;
};

struct alz_gnn_state;   // alz_gnn.cu
struct alz_sock_state;  // alz_sock.cu
struct alz_comm_state;  // alz_comm.cu

struct alz_handle {
  alz_config cfg{};
  int device = 0, sms = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr}, ev_tmp = nullptr;
  std::string last_err;

  // join build side: host mirror (ClusterInfo maps) + device open-addressed table
  std::unordered_map<uint32_t, HostEp> ep_host;
  bool ep_dirty = false;
  alz::EpEntry* d_ep = nullptr;
  uint32_t ep_cap = 0;

  // accumulators
  alz::AccTable pairs_fwd{}, pairs_rev{}, edges{};
  alz::Counters* d_ctr = nullptr;
  alz::Counters* h_ctr = nullptr;  // pinned
  alz::HotState* d_hot = nullptr;  // [2]: forward pairs, reversed pairs

  // flush scratch
  uint64_t* d_keys[2] = {nullptr, nullptr};
  uint32_t* d_rows[2] = {nullptr, nullptr};
  void* d_sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;
  alz_edge_out* d_out = nullptr;
  uint32_t n_live = 0, last_n_edges = 0;

  // host staging (double buffered)
  alz_l7_rec* h_stage[2] = {nullptr, nullptr};
  alz_l7_rec* d_stage[2] = {nullptr, nullptr};
  uint8_t* h_raw_stage = nullptr;
  uint8_t* d_raw_stage = nullptr;
  uint64_t stage_turn = 0;

  uint64_t events_in = 0, pending_since_fold = 0, windows = 0;
  uint64_t tcp_events_in = 0, tcp_localhost_dropped = 0;

  int comm_nranks = 1, comm_rank = 0;
  alz_comm_state* comm = nullptr;
  alz_gnn_state* gnn = nullptr;
  alz_sock_state* sock = nullptr;
};

int alz_internal_fold(alz_handle* h);
// multi-GPU merge of the prepared (sorted) live edges; ALZ_E_UNSUPPORTED = single rank,
// caller finishes the flush locally
int alz_internal_merge_ranks(alz_handle* h);
void alz_internal_free_extensions(alz_handle* h);
void alz_internal_free_comm(alz_handle* h);
void alz_internal_free_gnn(alz_handle* h);
void alz_internal_free_sock(alz_handle* h);
