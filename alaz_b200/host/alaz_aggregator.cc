// alaz_aggregator.cc — see alaz_aggregator.hpp. Plain C++ over the C ABI; no CUDA here.
#include "alaz_aggregator.hpp"

#include <cstdio>
#include <cstring>

#include "../../include/alazgpu_synth.h"   // alz_pinned_alloc / alz_pinned_free

namespace alaz {

namespace {
// reverse of L7ProtocolConversion.String(), ebpf/l7_req/l7.go:48-71
uint8_t ProtocolEnum(const std::string& p) {
  if (p == "HTTP") return ALZ_PROTO_HTTP;
  if (p == "AMQP") return ALZ_PROTO_AMQP;
  if (p == "POSTGRES") return ALZ_PROTO_POSTGRES;
  if (p == "HTTP2") return ALZ_PROTO_HTTP2;
  if (p == "REDIS") return ALZ_PROTO_REDIS;
  if (p == "KAFKA") return ALZ_PROTO_KAFKA;
  if (p == "MYSQL") return ALZ_PROTO_MYSQL;
  if (p == "MONGO") return ALZ_PROTO_MONGO;
  return ALZ_PROTO_UNKNOWN;
}
// reverse of the per-protocol method conversions, l7.go:204-325
uint8_t MethodEnum(uint8_t proto, const std::string& m) {
  static const char* http[] = {"", "GET", "POST", "PUT", "PATCH", "DELETE", "HEAD", "CONNECT", "OPTIONS", "TRACE"};
  switch (proto) {
    case ALZ_PROTO_HTTP:
      for (int i = 1; i <= 9; ++i) if (m == http[i]) return (uint8_t)i;
      return 0;
    case ALZ_PROTO_AMQP: return m == "PUBLISH" ? 1 : m == "DELIVER" ? 2 : 0;
    case ALZ_PROTO_POSTGRES: return m == "CLOSE_OR_TERMINATE" ? 1 : m == "SIMPLE_QUERY" ? 2 : m == "EXTENDED_QUERY" ? 3 : 0;
    case ALZ_PROTO_HTTP2: return m == "CLIENT_FRAME" ? 1 : m == "SERVER_FRAME" ? 2 : 0;
    case ALZ_PROTO_REDIS: return m == "COMMAND" ? 1 : m == "PUSHED_EVENT" ? 2 : m == "PING" ? 3 : 0;
    case ALZ_PROTO_KAFKA: return m == "PRODUCE_REQUEST" ? 1 : m == "FETCH_RESPONSE" ? 2 : 0;
    case ALZ_PROTO_MYSQL: return m == "TEXT_QUERY" ? 1 : m == "PREPARE_STMT" ? 2 : m == "EXEC_STMT" ? 3 : m == "STMT_CLOSE" ? 4 : 0;
    default: return 0;
  }
}
const char* NodeTypeName(uint8_t t) {   // aggregator/data.go:42-44 (a Host-header keyed destination is OUTBOUND too)
  return t == ALZ_NODE_POD ? "pod" : t == ALZ_NODE_SVC ? "service" : "outbound";
}
}  // namespace

uint32_t Aggregator::ParseIPv4(const std::string& s, bool* ok) {
  unsigned a, b, c, d;
  char tail;
  const bool good = sscanf(s.c_str(), "%u.%u.%u.%u%c", &a, &b, &c, &d, &tail) == 4 && a < 256 && b < 256 && c < 256 && d < 256;
  if (ok) *ok = good;
  return good ? (a << 24) | (b << 16) | (c << 8) | d : 0u;
}
std::string Aggregator::FormatIPv4(uint32_t ip) {
  char buf[16];
  snprintf(buf, sizeof buf, "%u.%u.%u.%u", ip >> 24, (ip >> 16) & 255u, (ip >> 8) & 255u, ip & 255u);
  return buf;
}

// parseHttpPayload, aggregator/data.go:508-531, the hostHeader part: the request is split on "\n"; among the
// lines after the first, the first one that starts with "Host:" AND splits on single spaces into >= 2 parts
// yields parts[1] minus a trailing "\r". A "Host:" line with fewer parts is skipped and the scan goes on.
std::string Aggregator::ParseHttpHostHeader(const std::string& payload) {
  size_t i = payload.find('\n');
  while (i != std::string::npos) {
    const size_t b = i + 1;
    const size_t e = payload.find('\n', b);
    const size_t len = (e == std::string::npos ? payload.size() : e) - b;
    if (len >= 5 && payload.compare(b, 5, "Host:") == 0) {
      const size_t sp = payload.find(' ', b);
      if (sp != std::string::npos && sp < b + len) {            // at least two parts
        size_t q = sp + 1, end = payload.find(' ', q);
        if (end == std::string::npos || end > b + len) end = b + len;
        std::string host = payload.substr(q, end - q);
        if (!host.empty() && host.back() == '\r') host.pop_back();
        return host;
      }
    }
    i = e;
  }
  return std::string();
}

Aggregator::Aggregator(DataStore* ds, const AggregatorConfig& cfg) : ds_(ds), cfg_(cfg) {
  alz_config c;
  memset(&c, 0, sizeof c);
  c.abi_version = ALZ_ABI_VERSION;
  c.device = cfg.Device;
  c.max_endpoints = cfg.MaxEndpoints;
  c.max_pairs = cfg.MaxPairs;
  c.max_batch = (uint32_t)cfg.BatchSize;
  if (cfg_.BatchSize == 0) cfg_.BatchSize = 1;
  const int rc = alz_create(&c, &h_);
  if (rc != ALZ_OK) { err_ = alz_strerror(rc); h_ = nullptr; return; }
  void* p = nullptr;
  if (alz_pinned_alloc(cfg_.BatchSize * sizeof(alz_l7_rec), &p) != ALZ_OK) {
    err_ = "alz_pinned_alloc";   // no batch buffer: the aggregator is unusable, Ok() says so
    alz_destroy(h_);
    h_ = nullptr;
    return;
  }
  batch_ = static_cast<alz_l7_rec*>(p);
  out_.resize(cfg.MaxPairs);
}

Aggregator::~Aggregator() {
  if (batch_) alz_pinned_free(batch_);
  if (h_) alz_destroy(h_);
}

uint32_t Aggregator::Interner::Acquire(const std::string& uid) {
  auto it = ids.find(uid);
  if (it != ids.end()) return it->second;
  uint32_t id;
  if (!free_ids.empty()) { id = free_ids.back(); free_ids.pop_back(); names[id] = uid; }
  else { id = (uint32_t)names.size(); names.push_back(uid); refs.push_back(0); }
  ids.emplace(uid, id);
  return id;
}
void Aggregator::Interner::Map(uint32_t ip, uint32_t id) {
  auto it = ip_to_id.find(ip);
  if (it != ip_to_id.end()) { if (it->second == id) return; Unmap(ip); }
  ip_to_id[ip] = id;
  refs[id]++;
}
void Aggregator::Interner::Unmap(uint32_t ip) {
  auto it = ip_to_id.find(ip);
  if (it == ip_to_id.end()) return;
  const uint32_t id = it->second;
  ip_to_id.erase(it);
  // no IP resolves to the id any more: edges of the open window may still carry it, so it is recycled one
  // window later (a UID whose stale IP entry is still in the table keeps its id, like the reference keeps the
  // map entry: ADD/UPDATE never delete the old IP, aggregator/persist.go:55-65)
  if (--refs[id] == 0) pending_free.push_back(id);
}
void Aggregator::Interner::EndWindow() {
  for (uint32_t id : pending_free)
    if (refs[id] == 0) { ids.erase(names[id]); free_ids.push_back(id); }
  pending_free.clear();
}

// processPod / processSvc: ADD and UPDATE write map[ip] = uid, DELETE removes (persist.go:55-71, 114-130)
void Aggregator::ProcessK8s(const K8sResourceMessage& m) {
  if (!h_) return;
  std::lock_guard<std::mutex> g(mu_);
  const bool pod = m.ResourceType == "Pod";
  if (!pod && m.ResourceType != "Service") return;   // other kinds are only relayed to the backend
  if (m.IP.empty()) return;                          // persist.go:37-40
  bool ok = false;
  const uint32_t ip = ParseIPv4(m.IP, &ok);
  if (!ok) return;                                   // e.g. ClusterIP "None"
  // events already batched were resolved by the reference with the tables as they were: submit first
  if (batch_n_) SubmitBatch();
  const int table = pod ? ALZ_TABLE_POD : ALZ_TABLE_SVC;
  Interner& in = pod ? pods_ : svcs_;
  if (m.EventType == "Add" || m.EventType == "Update") {
    const uint32_t id = in.Acquire(m.UID);
    const int rc = alz_table_upsert(h_, table, ip, id);
    if (rc != ALZ_OK) { err_ = std::string("alz_table_upsert: ") + alz_strerror(rc); return; }
    in.Map(ip, id);
  } else if (m.EventType == "Delete") {
    alz_table_erase(h_, table, ip);
    in.Unmap(ip);
  } else {
    return;
  }
  tables_dirty_ = true;
}

int Aggregator::SubmitBatch() {
  if (!h_) return ALZ_E_STATE;
  if (tables_dirty_) {
    const int rc = alz_table_commit(h_);
    if (rc != ALZ_OK) {   // events batched behind a failed commit are dropped, never left to pile up
      dropped_events_ += batch_n_;
      batch_n_ = 0;
      err_ = std::string("alz_table_commit: ") + alz_strerror(rc);
      return rc;
    }
    tables_dirty_ = false;
  }
  if (batch_n_ == 0) return ALZ_OK;
  const int rc = alz_submit_l7(h_, batch_, batch_n_);
  if (rc != ALZ_OK) { dropped_events_ += batch_n_; err_ = std::string("alz_submit_l7: ") + alz_strerror(rc); }
  batch_n_ = 0;   // the batch is gone either way: a failed submit must not let the buffer overrun
  return rc;
}

// the reader side of processL7: one compact record per event; the switch itself runs on the device
void Aggregator::ProcessL7(const L7Event& e) {
  if (!h_) return;
  std::lock_guard<std::mutex> g(mu_);
  if (tables_dirty_ || batch_n_ >= cfg_.BatchSize) SubmitBatch();   // table changes precede the events that follow them
  if (batch_n_ >= cfg_.BatchSize) return;                            // cannot happen (SubmitBatch empties the batch)
  alz_l7_rec& r = batch_[batch_n_++];
  r.saddr = e.Saddr; r.daddr = e.Daddr; r.sport = e.Sport; r.dport = e.Dport;
  r.status = e.Status > 65535u ? 65535u : (uint16_t)e.Status;
  r.protocol = ProtocolEnum(e.Protocol);
  r.method_flags = (uint8_t)((MethodEnum(r.protocol, e.Method) & ALZ_MF_METHOD_MASK) | (e.Tls ? ALZ_MF_TLS : 0) |
                             (e.PayloadRejected ? ALZ_MF_PAYLOAD_REJECT : 0));
  r.duration_ns = e.Duration;
  r.write_time_ns = e.WriteTimeNs;
  // setFromToV2's third-party branch (data.go:851-854): an HTTP destination that is neither service nor pod is
  // keyed by the request's Host header when there is one. The maps consulted are this adapter's mirror of the
  // tables it has upserted, in the same order as the events, so the decision is the reference's.
  if (r.protocol == ALZ_PROTO_HTTP && !e.Payload.empty() && !svcs_.ip_to_id.count(e.Daddr) && !pods_.ip_to_id.count(e.Daddr)) {
    const std::string host = ParseHttpHostHeader(e.Payload);
    if (!host.empty()) {
      bool is_ip = false;
      const uint32_t ip = ParseIPv4(host, &is_ip);
      if (is_ip) r.daddr = ip;   // a header that is itself a dotted quad is the same node as that raw daddr (same ToUID string)
      else {
        auto it = host_ids_.find(host);
        if (it == host_ids_.end()) { it = host_ids_.emplace(host, (uint32_t)host_names_.size()).first; host_names_.push_back(host); }
        r.daddr = it->second;
        r.protocol |= ALZ_PROTO_F_HOSTKEY;
      }
    }
  }
  if (batch_n_ >= cfg_.BatchSize) SubmitBatch();
}

void Aggregator::ProcessTcpConnect(const TcpConnectEvent& e) {
  if (!h_) return;
  alz_tcp_rec r;
  memset(&r, 0, sizeof r);
  r.fd = e.Fd; r.timestamp_ns = e.Timestamp; r.pid = e.Pid; r.sport = e.SPort; r.dport = e.DPort;
  r.saddr = ParseIPv4(e.SAddr, nullptr);
  r.daddr = ParseIPv4(e.DAddr, nullptr);
  // TcpStateConversion.String(), ebpf/tcp_state/tcp.go:27-60
  r.type = e.Type_ == "EVENT_TCP_ESTABLISHED" ? 1u : e.Type_ == "EVENT_TCP_CONNECT_FAILED" ? 2u
         : e.Type_ == "EVENT_TCP_LISTEN" ? 3u : e.Type_ == "EVENT_TCP_LISTEN_CLOSED" ? 4u
         : e.Type_ == "EVENT_TCP_CLOSED" ? 5u : 0u;
  alz_submit_tcp(h_, &r, 1);
}

int Aggregator::Flush(bool with_scores) {
  if (!h_) return ALZ_E_STATE;
  std::lock_guard<std::mutex> g(mu_);
  int rc = SubmitBatch();
  if (rc != ALZ_OK) return rc;
  size_t n = 0;
  rc = alz_window_flush(h_, out_.data(), out_.size(), &n);
  if (rc != ALZ_OK) return rc;
  if (with_scores && n) {
    scores_.resize(n);
    size_t ns = 0;
    rc = alz_gnn_score(h_, scores_.data(), n, &ns);
    if (rc != ALZ_OK) return rc;
  }
  std::vector<EdgeWindow> edges(n);
  for (size_t i = 0; i < n; ++i) {
    const alz_edge_out& o = out_[i];
    EdgeWindow& w = edges[i];
    auto uid = [&](uint8_t t, uint32_t v) -> std::string {
      if (t == ALZ_NODE_POD) return v < pods_.names.size() ? pods_.names[v] : std::string("?");
      if (t == ALZ_NODE_SVC) return v < svcs_.names.size() ? svcs_.names[v] : std::string("?");
      if (t == ALZ_NODE_OUTBOUND_HOST) return v < host_names_.size() ? host_names_[v] : std::string("?");   // data.go:852
      return FormatIPv4(v);   // outbound: the raw daddr string (data.go:862)
    };
    w.FromType = NodeTypeName(o.from_type); w.FromUID = uid(o.from_type, o.from);
    w.ToType = NodeTypeName(o.to_type);     w.ToUID = uid(o.to_type, o.to);
    w.Count = o.count; w.Err5xx = o.err5xx; w.LatSumNs = o.lat_sum_ns;
    memcpy(w.Hist, o.hist, sizeof w.Hist);
    if (with_scores) w.Score = scores_[i];
  }
  pods_.EndWindow();
  svcs_.EndWindow();
  return ds_ ? ds_->PersistEdgeWindow(edges) : ALZ_OK;
}

int Aggregator::Stats(alz_stats* out) { return h_ ? alz_get_stats(h_, out) : ALZ_E_STATE; }

int Aggregator::ClearSocketLines(bool send_alive, int64_t check_time_ms) {
  if (!h_) return ALZ_E_STATE;
  std::lock_guard<std::mutex> g(mu_);
  if (send_alive) {
    if (tables_dirty_) {   // the export resolves against the committed tables
      int rc = alz_table_commit(h_);
      if (rc != ALZ_OK) return rc;
      tables_dirty_ = false;
    }
    size_t n = 0;
    if (alive_.empty()) alive_.resize(1u << 12);
    int rc = alz_sock_alive(h_, alive_.data(), alive_.size(), &n);
    if (rc == ALZ_E_CAPACITY) {   // n = the rows there are
      alive_.resize(n + n / 4);
      rc = alz_sock_alive(h_, alive_.data(), alive_.size(), &n);
    }
    if (rc != ALZ_OK) return rc;
    for (size_t i = 0; i < n && ds_; ++i) {
      const alz_alive_conn& c = alive_[i];
      AliveConnection ac;
      ac.CheckTime = check_time_ms;
      ac.FromIP = FormatIPv4(c.from_ip); ac.FromType = "pod"; ac.FromPort = c.from_port;
      ac.FromUID = c.from_id < pods_.names.size() ? pods_.names[c.from_id] : std::string("?");
      ac.ToIP = FormatIPv4(c.to_ip); ac.ToPort = c.to_port;
      ac.ToType = NodeTypeName(c.to_type);
      if (c.to_type == ALZ_NODE_POD) ac.ToUID = c.to_id < pods_.names.size() ? pods_.names[c.to_id] : std::string("?");
      else if (c.to_type == ALZ_NODE_SVC) ac.ToUID = c.to_id < svcs_.names.size() ? svcs_.names[c.to_id] : std::string("?");
      else ac.ToUID = ac.ToIP;                                             // data.go:1672-1673
      ds_->PersistAliveConnection(ac);
    }
  }
  return alz_sock_gc(h_);
}

}  // namespace alaz
