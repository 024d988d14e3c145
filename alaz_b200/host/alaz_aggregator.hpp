// alaz_aggregator.hpp — host side of the path above the C ABI, in C++ (the reference is Go and no Go
// toolchain exists here; INTEGRATION.md shows the cgo equivalent). It mirrors the aggregator's seam:
// the same event types go in (only the fields this path reads), a DataStore-shaped sink comes out.
//
//   reference                                              here
//   NewAggregator(ctx, ct, k8sChan, events, ...)           alaz::Aggregator(ds, cfg)      aggregator/data.go:135-140
//   processk8s -> processPod / processSvc                  ProcessK8s(msg)                aggregator/persist.go:25-131
//   processEbpf -> processL7                               ProcessL7(ev)                  aggregator/data.go:310-337, 1364
//   processEbpfTcp -> processTcpConnect                    ProcessTcpConnect(ev)          aggregator/data.go:404-506
//   ds.PersistRequest(row) per event                       ds->PersistEdgeWindow(edges)   datastore/datastore.go:13
#pragma once
#include <cstdint>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/alazgpu.h"

namespace alaz {

// ebpf/l7_req/l7.go:396-421 (fields read on this path)
struct L7Event {
  uint64_t Fd = 0;
  uint32_t Pid = 0;
  uint32_t Status = 0;
  uint64_t Duration = 0;
  std::string Protocol;   // "HTTP", "AMQP", ... (l7.go:32-42)
  bool Tls = false;
  std::string Method;     // "GET", "DELIVER", "PUSHED_EVENT", ...
  uint64_t WriteTimeNs = 0;
  uint32_t Saddr = 0, Daddr = 0;
  uint16_t Sport = 0, Dport = 0;
  bool PayloadRejected = false;   // the Go-side SQL/Mongo parser returned an error (data.go:1252, 1288, 1328)
  std::string Payload;            // HTTP only: Payload[0:PayloadSize], read for its Host header (data.go:1213)
};

// ebpf/tcp_state/tcp.go:75-84
struct TcpConnectEvent {
  uint64_t Fd = 0, Timestamp = 0;
  std::string Type_;      // "EVENT_TCP_ESTABLISHED", "EVENT_TCP_CLOSED", ... (tcp.go:27-60)
  uint32_t Pid = 0;
  uint16_t SPort = 0, DPort = 0;
  std::string SAddr, DAddr;   // dotted quad (tcp.go:241-242)
};

// k8s/informer.go:236-240, reduced to what processPod/processSvc write into ClusterInfo
struct K8sResourceMessage {
  std::string ResourceType;   // "Pod" | "Service" (k8s/informer.go:26-36)
  std::string EventType;      // "Add" | "Update" | "Delete"
  std::string UID;
  std::string IP;             // pod.Status.PodIP or service.Spec.ClusterIP; empty pod IP is skipped (persist.go:37-40)
};

struct EdgeWindow {
  std::string FromType, FromUID, ToType, ToUID;   // "pod" | "service" | "outbound" (data.go:42-44)
  uint64_t Count = 0, Err5xx = 0, LatSumNs = 0;
  uint32_t Hist[ALZ_NB] = {0};
  float Score = 0.f;                              // GNN anomaly score when requested
};

// datastore/dto.go AliveConnection, as sendOpenConnection fills it (aggregator/data.go:1649-1675)
struct AliveConnection {
  int64_t CheckTime = 0;   // unix ms
  std::string FromIP, FromType, FromUID;
  uint16_t FromPort = 0;
  std::string ToIP, ToType, ToUID;
  uint16_t ToPort = 0;
};

// the sink, next to DataStore.PersistRequest (datastore/datastore.go:13)
class DataStore {
 public:
  virtual ~DataStore() = default;
  virtual int PersistEdgeWindow(const std::vector<EdgeWindow>& edges) = 0;
  virtual int PersistAliveConnection(const AliveConnection&) { return 0; }   // datastore/datastore.go:19
};

struct AggregatorConfig {
  int Device = 0;
  uint32_t MaxEndpoints = 1u << 16;
  uint32_t MaxPairs = 1u << 18;
  size_t BatchSize = 1u << 16;   // records per alz_submit_l7
};

class Aggregator {
 public:
  Aggregator(DataStore* ds, const AggregatorConfig& cfg);
  ~Aggregator();
  Aggregator(const Aggregator&) = delete;
  Aggregator& operator=(const Aggregator&) = delete;

  bool Ok() const { return h_ != nullptr; }
  const std::string& LastError() const { return err_; }
  // events dropped because a submit failed (capacity, CUDA error): the reference only logs such errors too
  // (aggregator/data.go:1244-1247), but the count is kept
  uint64_t DroppedEvents() const { return dropped_events_; }

  // All Process* methods and Flush may be called from several threads (the reference runs 4*NumCPU processL7
  // workers, aggregator/data.go:230-232); the batch buffer is guarded by one mutex, the library below is
  // thread-safe on its own.

  void ProcessK8s(const K8sResourceMessage& m);
  void ProcessL7(const L7Event& e);
  void ProcessTcpConnect(const TcpConnectEvent& e);
  // close the window: rows grouped by edge go to ds->PersistEdgeWindow. Returns 0 or an alz_status.
  int Flush(bool with_scores = false);
  int Stats(alz_stats* out);
  // one tick of clearSocketLines (aggregator/data.go:1681-1716, every 120 s there): with send_alive (the
  // reference's SEND_ALIVE_TCP_CONNECTIONS) every open connection with a pod at its source goes to
  // ds->PersistAliveConnection, then every socket line is garbage-collected (SocketLine.DeleteUnused)
  int ClearSocketLines(bool send_alive, int64_t check_time_ms = 0);

  static uint32_t ParseIPv4(const std::string& s, bool* ok);   // "a.b.c.d" -> the integer IntToIPv4 takes
  static std::string FormatIPv4(uint32_t ip);
  // parseHttpPayload's hostHeader (aggregator/data.go:508-531): "" when the reference finds none
  static std::string ParseHttpHostHeader(const std::string& payload);

 private:
  int SubmitBatch();   // requires mu_
  struct Interner {     // UID <-> dense id (< 2^29), ids recycled one window after their last IP mapping went away
    std::unordered_map<std::string, uint32_t> ids;
    std::vector<std::string> names;
    std::vector<uint32_t> refs;                       // IP entries currently mapping to the id
    std::unordered_map<uint32_t, uint32_t> ip_to_id;  // mirror of the device table's value per IP
    std::vector<uint32_t> free_ids, pending_free;
    uint32_t Acquire(const std::string& uid);
    void Map(uint32_t ip, uint32_t id);
    void Unmap(uint32_t ip);
    void EndWindow();
  };

  DataStore* ds_;
  alz_handle* h_ = nullptr;
  std::string err_;
  AggregatorConfig cfg_;
  alz_l7_rec* batch_ = nullptr;   // pinned (alz_pinned_alloc)
  size_t batch_n_ = 0;
  bool tables_dirty_ = false;
  std::mutex mu_;
  uint64_t dropped_events_ = 0;
  Interner pods_, svcs_;
  std::unordered_map<std::string, uint32_t> host_ids_;   // Host header text -> dense id (ALZ_NODE_OUTBOUND_HOST)
  std::vector<std::string> host_names_;
  std::vector<alz_edge_out> out_;
  std::vector<float> scores_;
  std::vector<alz_alive_conn> alive_;
};

}  // namespace alaz
