// host_unit_test.cc — CPU-only checks of the C++ adapter's conversions (no device needed): the
// string <-> enum / IPv4 mappings must be the inverse of the reference's conversions
// (ebpf/l7_req/l7.go:48-71, :204-325; aggregator/data.go:1751-1758) and an adapter without a
// GPU must fail loudly, not fall back.
#include <cstdio>
#include <string>

#include "../../alaz_b200/host/alaz_aggregator.hpp"

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

struct NullStore : alaz::DataStore {
  int PersistEdgeWindow(const std::vector<alaz::EdgeWindow>&) override { return 0; }
};

int main() {
  bool ok = false;
  CHECK(alaz::Aggregator::ParseIPv4("10.0.0.1", &ok) == 0x0A000001u && ok);
  CHECK(alaz::Aggregator::ParseIPv4("255.255.255.255", &ok) == 0xFFFFFFFFu && ok);
  CHECK(alaz::Aggregator::ParseIPv4("172.16.0.300", &ok) == 0u && !ok);
  CHECK(alaz::Aggregator::ParseIPv4("None", &ok) == 0u && !ok);       // headless service ClusterIP
  CHECK(alaz::Aggregator::ParseIPv4("1.2.3.4x", &ok) == 0u && !ok);
  CHECK(alaz::Aggregator::FormatIPv4(0x08080808u) == "8.8.8.8");
  for (uint32_t ip : {0u, 0x7F000001u, 0xC0A80164u, 0xFFFFFFFFu})
    CHECK(alaz::Aggregator::ParseIPv4(alaz::Aggregator::FormatIPv4(ip), &ok) == ip && ok);
  // parseHttpPayload's Host header (aggregator/data.go:508-531); the same vectors pin the oracle in tests/test_oracle.py
  using A = alaz::Aggregator;
  CHECK(A::ParseHttpHostHeader("GET /x HTTP/1.1\r\nHost: example.com\r\nAccept: */*\r\n\r\n") == "example.com");
  CHECK(A::ParseHttpHostHeader("GET /x HTTP/1.1\nHost: a.b:8080\n\n") == "a.b:8080");
  CHECK(A::ParseHttpHostHeader("GET /x HTTP/1.1\r\nAccept: */*\r\n\r\n") == "");
  CHECK(A::ParseHttpHostHeader("Host: first.line\r\nX: y\r\n") == "");                   // lines[0] is never looked at
  CHECK(A::ParseHttpHostHeader("GET / HTTP/1.1\r\nHost:nospace.com\r\nHost: second.com\r\n") == "second.com");
  CHECK(A::ParseHttpHostHeader("GET / HTTP/1.1\r\nHost:  two.spaces\r\n") == "");          // parts[1] is empty
  CHECK(A::ParseHttpHostHeader("GET / HTTP/1.1\r\nhost: lower.case\r\n") == "");          // HasPrefix is case-sensitive
  CHECK(A::ParseHttpHostHeader("GET / HTTP/1.1\r\nHost: h.com extra words\r\n") == "h.com");
  CHECK(A::ParseHttpHostHeader("GET / HTTP/1.1\r\nX-Host: no\r\nHost: yes.com") == "yes.com"); // last line without newline
  CHECK(A::ParseHttpHostHeader("") == "");
  NullStore ds;
  alaz::AggregatorConfig cfg;
  alaz::Aggregator a(&ds, cfg);
  if (!a.Ok()) {
    // no CUDA device here: every entry point must be inert and report the reason
    CHECK(a.LastError().find("no CUDA device") != std::string::npos);
    CHECK(a.Flush() == ALZ_E_STATE);
    alz_stats st;
    CHECK(a.Stats(&st) == ALZ_E_STATE);
    a.ProcessK8s({"Pod", "Add", "u", "10.0.0.1"});
    a.ProcessL7(alaz::L7Event{});
    printf("host unit ok (no device: adapter refused to start: %s)\n", a.LastError().c_str());
  } else {
    printf("host unit ok (device present)\n");
  }
  return 0;
}
