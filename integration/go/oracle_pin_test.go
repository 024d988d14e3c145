// oracle_pin_test.go — pins this repo's CPU oracle against the REAL reference aggregator.
//
// Where: copy (or symlink) this file into getanteon/alaz's `aggregator/` directory (it is an internal test of
// package aggregator: it calls the unexported processL7 and fills ClusterInfo directly, so no CRI socket, no
// eBPF and no backend are needed) and run, from the reference checkout:
//
//	ALZ_GOLDEN=/path/to/alaz-b200/tests/golden/resolve_branches.json go test ./aggregator -run TestOraclePin -v
//
// What: every event of tests/golden/resolve_branches.json (one per branch of processL7 / setFromToV2 /
// ReverseDirection, aggregator/data.go:1364-1383, :827-870, :1110-1112, :1151-1153, :1240-1242) is pushed
// through Aggregator.processL7 with a counting DataStore (the shape of MockDataStore,
// main_benchmark_test.go:661-678); the rows PersistRequest receives are grouped by
// (FromType, FromUID, ToType, ToUID) exactly as docs/SPEC.md §3 defines an edge, and count / err5xx /
// lat_sum / histogram are compared with the file's `expect_edges` — the same expectations
// oracle/alz_oracle.c, oracle/ref_py.py and the CUDA path are held to. A green run turns the oracle's
// "PARITY UNPINNED" status for resolve/emit into "pinned by the reference itself".
//
// This repo's build image has no Go toolchain (SURVEY.md §8c), so this file has never been compiled here;
// it is written against the reference at 828b997f.
package aggregator

import (
	"context"
	"encoding/json"
	"fmt"
	"net"
	"os"
	"sort"
	"strconv"
	"sync"
	"testing"

	"github.com/ddosify/alaz/datastore"
	"github.com/ddosify/alaz/ebpf/l7_req"
	"k8s.io/apimachinery/pkg/types"
)

type pinEvent struct {
	ID     string      `json:"id"`
	Proto  interface{} `json:"proto"` // protocol name, or a raw number for "no such protocol"
	Method int         `json:"method"`
	Flags  []string    `json:"flags"`
	Saddr  string      `json:"saddr"`
	Daddr  string      `json:"daddr"`
	Status uint32      `json:"status"`
	Dur    uint64      `json:"dur"`
}
type pinEdge struct {
	From   []interface{}     `json:"from"` // ["pod"|"svc"|"outbound", id or ip]
	To     []interface{}     `json:"to"`
	Count  uint64            `json:"count"`
	Err5xx uint64            `json:"err5xx"`
	LatSum uint64            `json:"lat_sum"`
	Hist   map[string]uint32 `json:"hist"`
}
type pinFile struct {
	Pods     map[string]int `json:"pods"`
	Services map[string]int `json:"services"`
	Events   []pinEvent     `json:"events"`
	Expect   []pinEdge      `json:"expect_edges"`
}

type pinAcc struct {
	count, err5xx, latSum uint64
	hist                  [64]uint32
}

// countingDS implements datastore.DataStore; only PersistRequest does anything.
type countingDS struct {
	mu    sync.Mutex
	edges map[string]*pinAcc
}

func (c *countingDS) PersistPod(datastore.Pod, string) error                 { return nil }
func (c *countingDS) PersistService(datastore.Service, string) error         { return nil }
func (c *countingDS) PersistReplicaSet(datastore.ReplicaSet, string) error   { return nil }
func (c *countingDS) PersistDeployment(datastore.Deployment, string) error   { return nil }
func (c *countingDS) PersistEndpoints(datastore.Endpoints, string) error     { return nil }
func (c *countingDS) PersistContainer(datastore.Container, string) error     { return nil }
func (c *countingDS) PersistDaemonSet(datastore.DaemonSet, string) error     { return nil }
func (c *countingDS) PersistStatefulSet(datastore.StatefulSet, string) error { return nil }
func (c *countingDS) PersistKafkaEvent(*datastore.KafkaEvent) error          { return nil }
func (c *countingDS) PersistAliveConnection(*datastore.AliveConnection) error {
	return nil
}

// docs/SPEC.md §4
func pinBucket(d uint64) int {
	if d < 256 {
		return 0
	}
	o := 0
	for t := d; t > 1; t >>= 1 {
		o++
	}
	if o >= 40 {
		return 63
	}
	return 2*(o-8) + int((d>>(uint(o)-1))&1)
}

func (c *countingDS) PersistRequest(r *datastore.Request) error {
	key := fmt.Sprintf("%s|%s|%s|%s", r.FromType, r.FromUID, r.ToType, r.ToUID)
	c.mu.Lock()
	defer c.mu.Unlock()
	a := c.edges[key]
	if a == nil {
		a = &pinAcc{}
		c.edges[key] = a
	}
	a.count++
	if (r.Protocol == "HTTP" || r.Protocol == "HTTPS") && r.StatusCode >= 500 && r.StatusCode < 600 {
		a.err5xx++
	}
	a.latSum += r.Latency
	a.hist[pinBucket(r.Latency)]++
	return nil
}

func ipToU32(s string) uint32 {
	ip := net.ParseIP(s).To4()
	return uint32(ip[0])<<24 | uint32(ip[1])<<16 | uint32(ip[2])<<8 | uint32(ip[3])
}

// the strings L7Prog.Consume produces (ebpf/l7_req/l7.go:48-71, :204-325, :712-734)
func protoString(p interface{}) string {
	if s, ok := p.(string); ok {
		return s
	}
	return "Unknown"
}
func methodString(proto string, m int) string {
	switch proto {
	case l7_req.L7_PROTOCOL_HTTP:
		return l7_req.HTTPMethodConversion(m).String()
	case l7_req.L7_PROTOCOL_AMQP:
		return l7_req.RabbitMQMethodConversion(m).String()
	case l7_req.L7_PROTOCOL_POSTGRES:
		return l7_req.PostgresMethodConversion(m).String()
	case l7_req.L7_PROTOCOL_HTTP2:
		return l7_req.Http2MethodConversion(m).String()
	case l7_req.L7_PROTOCOL_REDIS:
		return l7_req.RedisMethodConversion(m).String()
	case l7_req.L7_PROTOCOL_KAFKA:
		return l7_req.KafkaMethodConversion(m).String()
	case l7_req.L7_PROTOCOL_MYSQL:
		return l7_req.MySQLMethodConversion(m).String()
	}
	return "Unknown"
}

func nodeKey(n []interface{}) (string, string) {
	kind := n[0].(string)
	switch kind {
	case "pod":
		return POD, "pod-" + strconv.Itoa(int(n[1].(float64)))
	case "svc":
		return SVC, "svc-" + strconv.Itoa(int(n[1].(float64)))
	}
	return OUTBOUND, n[1].(string)
}

func TestOraclePin(t *testing.T) {
	path := os.Getenv("ALZ_GOLDEN")
	if path == "" {
		t.Skip("ALZ_GOLDEN not set (path to tests/golden/resolve_branches.json)")
	}
	raw, err := os.ReadFile(path)
	if err != nil {
		t.Fatal(err)
	}
	var g pinFile
	if err := json.Unmarshal(raw, &g); err != nil {
		t.Fatal(err)
	}
	ds := &countingDS{edges: map[string]*pinAcc{}}
	a := &Aggregator{
		ctx: context.Background(),
		ds:  ds,
		clusterInfo: &ClusterInfo{ // what processPod / processSvc write (aggregator/persist.go:55-71, :114-130)
			PodIPToPodUid:         map[string]types.UID{},
			ServiceIPToServiceUid: map[string]types.UID{},
		},
		pgStmts:    map[string]string{},
		mySqlStmts: map[string]string{},
	}
	for ip, id := range g.Pods {
		a.clusterInfo.PodIPToPodUid[ip] = types.UID("pod-" + strconv.Itoa(id))
	}
	for ip, id := range g.Services {
		a.clusterInfo.ServiceIPToServiceUid[ip] = types.UID("svc-" + strconv.Itoa(id))
	}
	for _, e := range g.Events {
		proto := protoString(e.Proto)
		ev := &l7_req.L7Event{
			Status: e.Status, Duration: e.Dur, Protocol: proto, Method: methodString(proto, e.Method),
			Saddr: ipToU32(e.Saddr), Daddr: ipToU32(e.Daddr), Sport: 40000, Dport: 80, WriteTimeNs: 1000,
		}
		payload := "GET /x HTTP/1.1\r\n\r\n" // no Host header: outbound stays keyed by the raw daddr (data.go:862)
		for _, f := range e.Flags {
			switch f {
			case "tls":
				ev.Tls = true
			case "reject":
				// what ALZ_MF_PAYLOAD_REJECT stands for: a payload the reference's own parser refuses
				// (SQL text without a keyword, data.go:1440-1443 / :1495-1497; Mongo garbage, :1252-1255)
				payload = "\x00\x00\x00\x00not sql at all"
			}
		}
		if proto == l7_req.L7_PROTOCOL_POSTGRES || proto == l7_req.L7_PROTOCOL_MYSQL {
			if payload[0] != 0 {
				payload = "Q\x00\x00\x00\x10SELECT 1;\x00" // a simple query the keyword regex accepts
			}
		}
		copy(ev.Payload[:], payload)
		ev.PayloadSize = uint32(len(payload))
		ev.PayloadReadComplete = true
		a.processL7(context.Background(), ev)
	}
	// compare
	want := map[string]pinEdge{}
	for _, e := range g.Expect {
		ft, fu := nodeKey(e.From)
		tt, tu := nodeKey(e.To)
		want[fmt.Sprintf("%s|%s|%s|%s", ft, fu, tt, tu)] = e
	}
	var keys []string
	for k := range ds.edges {
		keys = append(keys, k)
	}
	sort.Strings(keys)
	for _, k := range keys {
		got := ds.edges[k]
		w, ok := want[k]
		if !ok {
			t.Errorf("reference emitted an edge the oracle does not expect: %s (count %d)", k, got.count)
			continue
		}
		if got.count != w.Count || got.err5xx != w.Err5xx || got.latSum != w.LatSum {
			t.Errorf("%s: reference count/err5xx/lat = %d/%d/%d, oracle expects %d/%d/%d", k, got.count, got.err5xx,
				got.latSum, w.Count, w.Err5xx, w.LatSum)
		}
		for b := 0; b < 64; b++ {
			if got.hist[b] != w.Hist[strconv.Itoa(b)] {
				t.Errorf("%s: hist[%d] = %d, oracle expects %d", k, b, got.hist[b], w.Hist[strconv.Itoa(b)])
			}
		}
		delete(want, k)
	}
	for k := range want {
		t.Errorf("oracle expects an edge the reference did not emit: %s", k)
	}
}
