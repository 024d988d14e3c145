// sim_test.cc — the reference's load test (main_benchmark_test.go:84-150 TestSimulation +
// testconfig/config1.json) replayed against the C++ host adapter: 100 pods, 50 services,
// 20 random pod->service edges, edgeRate x testDuration HTTP events per edge
// (Status 200, Duration 50, payload "GET /user HTTP1.1"), one tcp ESTABLISHED per edge,
// MockDataStore counting what reaches the sink.
//
// Differences, on purpose: events carry Saddr/Daddr (today's resolver needs them,
// aggregator/data.go:1760-1767; the reference's simulator predates that and no longer
// compiles), nothing is dropped on full channels, so the assertion is exact
// (== duration*edgeCount*edgeRate rows, and the per-edge split) instead of ">= 90 %".
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "../../alaz_b200/host/alaz_aggregator.hpp"

struct SimulatorConfig {   // testconfig/config1.json
  int testDuration = 15, podCount = 100, serviceCount = 50, edgeCount = 20, edgeRate = 10000;
};

struct MockDataStore : alaz::DataStore {   // main_benchmark_test.go:639-678
  uint64_t ReadyToBeSendReq = 0, Windows = 0;
  std::map<std::string, uint64_t> perEdge;
  std::map<std::string, float> score;
  int PersistEdgeWindow(const std::vector<alaz::EdgeWindow>& edges) override {
    Windows++;
    for (const auto& e : edges) {
      ReadyToBeSendReq += e.Count;
      perEdge[e.FromType + ":" + e.FromUID + "->" + e.ToType + ":" + e.ToUID] += e.Count;
      score[e.FromUID + "->" + e.ToUID] = e.Score;
    }
    return 0;
  }
  std::vector<alaz::AliveConnection> alive;
  int PersistAliveConnection(const alaz::AliveConnection& c) override { alive.push_back(c); return 0; }
};

static uint64_t rng_state = 0x5EED5EEDull;
static uint32_t rnd(uint32_t n) {
  rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
  return (uint32_t)((rng_state >> 33) % n);
}

#define CHECK(cond, ...)                                             \
  do {                                                               \
    if (!(cond)) { fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); return 1; } \
  } while (0)

int main(int argc, char** argv) {
  SimulatorConfig conf;
  if (argc > 1) conf.testDuration = atoi(argv[1]);
  MockDataStore ds;
  alaz::AggregatorConfig ac;
  alaz::Aggregator a(&ds, ac);
  CHECK(a.Ok(), "aggregator: %s", a.LastError().c_str());

  // Simulator.Setup: pods, services (K8sResourceMessage ADD), then edges
  struct FakePod { std::string Name, IP, Uid; uint32_t Pid; };
  std::vector<FakePod> pods;
  std::vector<std::pair<std::string, std::string>> svcs;   // uid, ip
  for (int i = 0; i < conf.podCount; ++i) {
    FakePod p{"pod-" + std::to_string(i), "10.1." + std::to_string(i / 250) + "." + std::to_string(1 + i % 250),
              "pod-uid-" + std::to_string(i), 1000u + (uint32_t)i};
    pods.push_back(p);
    a.ProcessK8s({"Pod", "Add", p.Uid, p.IP});
  }
  a.ProcessK8s({"Pod", "Add", "pod-without-ip", ""});        // persist.go:37-40: skipped
  for (int i = 0; i < conf.serviceCount; ++i) {
    svcs.emplace_back("svc-uid-" + std::to_string(i), "172.20.0." + std::to_string(1 + i));
    a.ProcessK8s({"Service", "Add", svcs.back().first, svcs.back().second});
  }
  a.ProcessK8s({"Service", "Add", "headless", "None"});      // not an IP: ignored
  struct Traffic { int pod, svc; uint64_t fd; };
  std::vector<Traffic> edges;
  std::map<std::string, uint64_t> expected;
  for (int i = 0; i < conf.edgeCount; ++i) {
    Traffic t{(int)rnd(conf.podCount), (int)rnd(conf.serviceCount), 3u + (uint64_t)i};
    edges.push_back(t);
    // tcpEstablish (main_benchmark_test.go:622-633)
    alaz::TcpConnectEvent c;
    c.Fd = t.fd; c.Timestamp = 1000 + (uint64_t)i; c.Type_ = "EVENT_TCP_ESTABLISHED"; c.Pid = pods[t.pod].Pid;
    c.SAddr = pods[t.pod].IP; c.DAddr = svcs[t.svc].second;
    a.ProcessTcpConnect(c);
  }
  // httpTraffic: edgeRate events per second per edge for testDuration seconds
  const uint64_t perEdge = (uint64_t)conf.testDuration * conf.edgeRate;
  bool ok = false;
  for (int sec = 0; sec < conf.testDuration; ++sec) {
    for (const Traffic& t : edges) {
      alaz::L7Event e;
      e.Fd = t.fd; e.Pid = pods[t.pod].Pid; e.Status = 200; e.Duration = 50; e.Protocol = "HTTP"; e.Method = "GET";
      e.Saddr = alaz::Aggregator::ParseIPv4(pods[t.pod].IP, &ok);
      e.Daddr = alaz::Aggregator::ParseIPv4(svcs[t.svc].second, &ok);
      e.Sport = 40000; e.Dport = 80;
      for (int k = 0; k < conf.edgeRate; ++k) { e.WriteTimeNs = (uint64_t)sec * 1000000000ull + (uint64_t)k; a.ProcessL7(e); }
    }
    // a 1 s window, like the ticker INTEGRATION.md describes; scores on the last one
    CHECK(a.Flush(sec == conf.testDuration - 1) == 0, "flush failed");
  }
  for (const Traffic& t : edges)
    expected["pod:" + pods[t.pod].Uid + "->service:" + svcs[t.svc].first] += perEdge;

  const uint64_t expectedTotalReqProcessed = (uint64_t)conf.testDuration * conf.edgeCount * conf.edgeRate;
  CHECK(ds.ReadyToBeSendReq == expectedTotalReqProcessed, "rows %" PRIu64 " != expected %" PRIu64, ds.ReadyToBeSendReq,
        expectedTotalReqProcessed);
  CHECK(ds.Windows == (uint64_t)conf.testDuration, "windows %" PRIu64, ds.Windows);
  CHECK(ds.perEdge == expected, "per-edge split differs (%zu vs %zu edges)", ds.perEdge.size(), expected.size());
  for (auto& kv : ds.score) CHECK(kv.second > 0.f && kv.second < 1.f, "score out of range");

  // DELETE a source pod: its events are dropped like setFromToV2 does (data.go:829-832); UPDATE re-keys an edge
  const Traffic& t0 = edges[0];
  a.ProcessK8s({"Pod", "Delete", pods[t0.pod].Uid, pods[t0.pod].IP});
  alaz::L7Event e;
  e.Status = 503; e.Duration = 1000; e.Protocol = "HTTP"; e.Method = "POST"; e.Tls = true;
  e.Saddr = alaz::Aggregator::ParseIPv4(pods[t0.pod].IP, &ok);
  e.Daddr = alaz::Aggregator::ParseIPv4(svcs[t0.svc].second, &ok);
  for (int k = 0; k < 1000; ++k) a.ProcessL7(e);
  a.ProcessK8s({"Pod", "Add", "reborn-uid", pods[t0.pod].IP});
  for (int k = 0; k < 10; ++k) a.ProcessL7(e);
  ds.perEdge.clear();
  CHECK(a.Flush() == 0, "flush failed");
  CHECK(ds.perEdge.size() == 1 && ds.perEdge.begin()->first == "pod:reborn-uid->service:" + svcs[t0.svc].first &&
            ds.perEdge.begin()->second == 10, "table change semantics");
  alz_stats st;
  CHECK(a.Stats(&st) == 0, "stats");
  CHECK(st.src_unresolved == 1000, "src_unresolved %" PRIu64, (uint64_t)st.src_unresolved);
  CHECK(st.tcp_events_in == (uint64_t)conf.edgeCount, "tcp events");
  // clearSocketLines tick with SEND_ALIVE_TCP_CONNECTIONS (data.go:1681-1716): every tcpEstablish above is still
  // open; the one whose pod was deleted and re-added under a new UID reports the new UID
  CHECK(a.ClearSocketLines(true, 1234) == 0, "ClearSocketLines: %s", a.LastError().c_str());
  CHECK(ds.alive.size() == (size_t)conf.edgeCount, "alive connections %zu", ds.alive.size());
  size_t reborn = 0;
  for (const auto& c : ds.alive) {
    CHECK(c.FromType == "pod" && c.ToType == "service" && c.CheckTime == 1234, "alive row types");
    CHECK(c.ToUID.rfind("svc-uid-", 0) == 0 && c.ToIP.rfind("172.20.0.", 0) == 0, "alive row destination");
    reborn += c.FromUID == "reborn-uid";
  }
  size_t on_t0 = 0;
  for (const Traffic& t : edges) on_t0 += t.pod == t0.pod;
  CHECK(reborn == on_t0, "alive rows of the re-added pod: %zu vs %zu", reborn, on_t0);
  printf("sim ok: %" PRIu64 " rows over %d windows, %zu edges\n", expectedTotalReqProcessed, conf.testDuration,
         expected.size());
  return 0;
}
