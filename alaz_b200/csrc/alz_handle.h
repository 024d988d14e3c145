// alz_handle.h — the handle behind the C ABI: everything one GPU owns.
#pragma once
#include <cstdint>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>
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

// One host->device staging slot: pinned host buffer + device buffer. A submitting thread owns the slot
// (mu) from the copy into the pinned buffer until its H2D and kernel are enqueued.
struct StageSlot {
  std::mutex mu;
  void* h = nullptr;
  void* d = nullptr;
  void* d_aux = nullptr;                  // raw path: compacted 32-B records
  cudaEvent_t copied = nullptr;           // H2D out of h done -> h may be rewritten
  cudaEvent_t consumed = nullptr;         // kernel that read d (and d_aux) done -> d may be rewritten
};
constexpr int kStageSlots = 4;
constexpr int kRawSlots = 2;

struct alz_handle {
  alz_config cfg{};
  int device = 0, sms = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev_tmp = nullptr, ev_count = nullptr;
  std::string last_err;

  // Serialises everything that enqueues on `stream` or touches the bookkeeping below. Submitting threads
  // take it only around their enqueue; flush / commit / stats hold it for the whole call.
  std::mutex mu;
  std::mutex turn_mu;                     // staging slot hand-out
  uint64_t stage_turn = 0, raw_turn = 0;
  StageSlot stage[kStageSlots];
  StageSlot raw[kRawSlots];
  size_t raw_chunk = 0;                   // samples per raw slot

  // join build side: host mirror of the ClusterInfo maps and of the device's open-addressed table;
  // a commit patches only the slots that changed (alz_table_commit)
  std::unordered_map<uint32_t, HostEp> ep_host;
  std::vector<uint32_t> ep_dirty_ips;
  std::vector<alz::EpEntry> ep_tab;       // host mirror, ep_cap entries
  std::vector<uint8_t> ep_touched;        // per slot: changed since the last upload
  std::vector<uint32_t> ep_touched_list;
  alz::EpEntry* d_ep = nullptr;
  uint32_t ep_cap = 0;
  std::vector<uint8_t> bloom_cnt;         // per filter bit: how many pod addresses set it (saturating at 255)
  bool bloom_dirty = false;
  uint32_t* d_bloom = nullptr;            // ALZ_BLOOM_WORDS words
  uint32_t* h_bloom = nullptr;            // pinned staging
  void* h_patch = nullptr;                // pinned: (slot, entry) records of one commit
  void* d_patch = nullptr;
  size_t patch_cap = 0;
  cudaEvent_t ev_patch = nullptr;

  // accumulators
  alz::AccTable pairs{}, edges{};
  alz::Counters* d_ctr = nullptr;
  alz::Counters* h_ctr = nullptr;  // pinned
  alz::HotState* d_hot = nullptr;

  // flush scratch
  uint64_t* d_keys[2] = {nullptr, nullptr};
  uint32_t* d_rows[2] = {nullptr, nullptr};
  void* d_sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;
  alz_edge_out* d_out = nullptr;
  uint32_t n_live = 0, last_n_edges = 0;
  uint64_t lost_reported = 0;             // capacity_events already reported by an earlier flush

  // time-cut windows (alz_window_clock)
  bool win_on = false;
  uint64_t win_off = 0, win_len = 0;      // first_user - first_kernel (mod 2^64), window length
  uint64_t* d_win = nullptr;              // device WinClock {lo, len, ready}
  alz_l7_rec* d_defer[2] = {nullptr, nullptr};   // records of later windows (ping-pong)
  uint32_t defer_cap = 0;
  int defer_cur = 0;
  uint64_t deferred_last = 0;

  uint64_t events_in = 0, pending_since_fold = 0, windows = 0;
  uint64_t launches = 0;                  // own kernels launched (stats)
  cudaEvent_t ev_t[3] = {nullptr, nullptr, nullptr};   // flush start / local part done / merge done (timing events)
  bool ev_t_valid = false;
  uint64_t collective_bytes_last = 0;
  uint64_t tcp_events_in = 0, tcp_localhost_dropped = 0;

  int comm_nranks = 1, comm_rank = 0;
  alz_comm_state* comm = nullptr;
  alz_gnn_state* gnn = nullptr;
  alz_sock_state* sock = nullptr;
};

int alz_internal_fold(alz_handle* h);
// device records -> the ingest kernel of the current mode; caller holds h->mu and has set the device
int alz_internal_ingest(alz_handle* h, const alz_l7_rec* d_recs, size_t n);
// multi-GPU merge of the prepared (sorted) live edges; ALZ_E_UNSUPPORTED = single rank,
// caller finishes the flush locally. local_rc: this rank's own status so far (every rank enters the
// collectives even when its local preparation failed, and all ranks return the same failure)
int alz_internal_merge_ranks(alz_handle* h, int local_rc);
void alz_internal_free_comm(alz_handle* h);
void alz_internal_free_gnn(alz_handle* h);
void alz_internal_free_sock(alz_handle* h);
