"""Second, independent restatement of the reference resolve/emit path in plain
Python (dicts and strings, small cases only). TEST INFRASTRUCTURE ONLY.

It exists to cross-check oracle/alz_oracle.c on random small cases and against
the hand-derived golden vectors; it shares no code with the C oracle or with
the product. PARITY UNPINNED for resolve/emit (see oracle/alz_oracle.h).
"""

# ebpf/l7_req/l7.go:48-71
_PROTO = {0: "UNKNOWN", 1: "HTTP", 2: "AMQP", 3: "POSTGRES", 4: "HTTP2",
          5: "REDIS", 6: "KAFKA", 7: "MYSQL", 8: "MONGO"}
# ebpf/l7_req/l7.go:204-325, switch at :712-734
_METHOD = {
    "HTTP": {1: "GET", 2: "POST", 3: "PUT", 4: "PATCH", 5: "DELETE", 6: "HEAD",
             7: "CONNECT", 8: "OPTIONS", 9: "TRACE"},
    "AMQP": {1: "PUBLISH", 2: "DELIVER"},
    "POSTGRES": {1: "CLOSE_OR_TERMINATE", 2: "SIMPLE_QUERY", 3: "EXTENDED_QUERY"},
    "HTTP2": {1: "CLIENT_FRAME", 2: "SERVER_FRAME"},
    "REDIS": {1: "COMMAND", 2: "PUSHED_EVENT", 3: "PING"},
    "KAFKA": {1: "PRODUCE_REQUEST", 2: "FETCH_RESPONSE"},
    "MYSQL": {1: "TEXT_QUERY", 2: "PREPARE_STMT", 3: "EXEC_STMT", 4: "STMT_CLOSE"},
}
NB = 64


def ip_string(x):  # IntToIPv4(x).String(), aggregator/data.go:1751-1767
    return "%d.%d.%d.%d" % ((x >> 24) & 255, (x >> 16) & 255, (x >> 8) & 255, x & 255)


def bucket(d):  # docs/SPEC.md §4
    if d < 256:
        return 0
    o = d.bit_length() - 1
    if o >= 40:
        return NB - 1
    return 2 * (o - 8) + ((d >> (o - 1)) & 1)


class Aggregator:
    def __init__(self):
        self.pod_ip_to_uid = {}   # cluster.go:15
        self.svc_ip_to_uid = {}   # cluster.go:16
        self.groups = {}
        self.stats = dict(events_in=0, rows_emitted=0, not_request=0, src_unresolved=0)

    def set_from_to_v2(self, row):  # aggregator/data.go:827-870
        uid = self.pod_ip_to_uid.get(row["FromIP"])
        if uid is None:
            return False
        row["FromUID"], row["FromType"] = uid, "pod"
        s = self.svc_ip_to_uid.get(row["ToIP"])
        if s is not None:
            row["ToUID"], row["ToType"] = s, "service"
            return True
        p = self.pod_ip_to_uid.get(row["ToIP"])
        if p is not None:
            row["ToUID"], row["ToType"] = p, "pod"
            return True
        row["ToUID"], row["ToType"] = row["ToIP"], "outbound"  # :862, DNS treated as failing
        return True

    def process_l7(self, rec):  # aggregator/data.go:1364-1383
        self.stats["events_in"] += 1
        proto = _PROTO.get(int(rec["protocol"]), "Unknown")
        mf = int(rec["method_flags"])
        method = _METHOD.get(proto, {}).get(mf & 0x3F, "Unknown")
        tls, reject = bool(mf & 0x80), bool(mf & 0x40)
        if proto not in ("HTTP", "AMQP", "REDIS", "POSTGRES", "MYSQL", "MONGO"):
            self.stats["not_request"] += 1
            return
        if proto in ("POSTGRES", "MYSQL", "MONGO") and reject:
            self.stats["not_request"] += 1
            return
        row = dict(Latency=int(rec["duration_ns"]), FromIP=ip_string(int(rec["saddr"])),
                   ToIP=ip_string(int(rec["daddr"])), Protocol=proto, Tls=tls,
                   StatusCode=int(rec["status"]), Method=method)
        if not self.set_from_to_v2(row):
            self.stats["src_unresolved"] += 1
            return
        if (proto == "AMQP" and method == "DELIVER") or (proto == "REDIS" and method == "PUSHED_EVENT"):
            # Request.ReverseDirection, datastore/dto.go:246-251
            row["FromIP"], row["ToIP"] = row["ToIP"], row["FromIP"]
            row["FromUID"], row["ToUID"] = row["ToUID"], row["FromUID"]
            row["FromType"], row["ToType"] = row["ToType"], row["FromType"]
        if proto == "HTTP" and tls:
            row["Protocol"] = "HTTPS"
        self.persist_request(row)

    def persist_request(self, row):
        key = (row["FromType"], row["FromUID"], row["ToType"], row["ToUID"])
        g = self.groups.setdefault(key, dict(count=0, err5xx=0, lat_sum=0, hist=[0] * NB))
        g["count"] += 1
        if row["Protocol"] in ("HTTP", "HTTPS") and 500 <= row["StatusCode"] < 600:
            g["err5xx"] += 1
        g["lat_sum"] += row["Latency"]
        g["hist"][bucket(row["Latency"])] += 1
        self.stats["rows_emitted"] += 1
