"""Python mirror of include/alazgpu.h: record layouts and constants.

Pure data definitions (numpy dtypes + ctypes structs); no compute. The sizes
are asserted against the C header in tests/test_abi.py.
"""
import ctypes as C
import os

import numpy as np

ABI_VERSION = int(os.environ.get("ALZ_ABI_VERSION", "2"))   # override only to drive an older build for A/B timing
NB = 64
BPF_L7_EVENT_SIZE = 1096
COMM_ID_BYTES = 128

# status codes
OK, E_INVAL, E_NOMEM, E_CUDA, E_NODEVICE, E_CAPACITY, E_STATE, E_NCCL, E_UNSUPPORTED = (
    0, -1, -2, -3, -4, -5, -6, -7, -8)

# ebpf/l7_req/l7.go:19-29
PROTO_UNKNOWN, PROTO_HTTP, PROTO_AMQP, PROTO_POSTGRES, PROTO_HTTP2 = 0, 1, 2, 3, 4
PROTO_REDIS, PROTO_KAFKA, PROTO_MYSQL, PROTO_MONGO = 5, 6, 7, 8
AMQP_PUBLISH, AMQP_DELIVER = 1, 2
REDIS_COMMAND, REDIS_PUSHED_EVENT, REDIS_PING = 1, 2, 3
MF_METHOD_MASK, MF_PAYLOAD_REJECT, MF_TLS = 0x3F, 0x40, 0x80
NODE_POD, NODE_SVC, NODE_OUTBOUND, NODE_OUTBOUND_HOST = 0, 1, 2, 3
PROTO_F_HOSTKEY = 0x40
TABLE_POD, TABLE_SVC = 0, 1
CFG_EAGER_JOIN = 0x1
CFG_NO_SMEM_CACHE = 0x2
MIX_SURVEY, MIX_ALL = 0, 1

L7_REC = np.dtype([
    ("saddr", "<u4"), ("daddr", "<u4"), ("sport", "<u2"), ("dport", "<u2"),
    ("status", "<u2"), ("protocol", "u1"), ("method_flags", "u1"),
    ("duration_ns", "<u8"), ("write_time_ns", "<u8"),
])
assert L7_REC.itemsize == 32

L7_REC16 = np.dtype([
    ("saddr", "<u4"), ("daddr", "<u4"), ("status", "<u2"), ("protocol", "u1"), ("method_flags", "u1"),
    ("duration_ns", "<u4"),
])
assert L7_REC16.itemsize == 16
REC16_DUR_OVERFLOW = 0x80

TCP_REC = np.dtype([
    ("fd", "<u8"), ("timestamp_ns", "<u8"), ("pid", "<u4"), ("saddr", "<u4"),
    ("daddr", "<u4"), ("sport", "<u2"), ("dport", "<u2"), ("type", "<u4"), ("_pad", "<u4"),
])
assert TCP_REC.itemsize == 40

# struct tcp_event as the perf ring carries it (ebpf/c/struct.h:2-12), 64 B with tail padding
BPF_TCP_EVENT = np.dtype({"names": ["fd", "timestamp", "type", "pid", "sport", "dport", "saddr", "daddr"],
                          "formats": ["<u8", "<u8", "<u4", "<u4", "<u2", "<u2", ("u1", (16,)), ("u1", (16,))],
                          "offsets": [0, 8, 16, 20, 24, 26, 28, 44], "itemsize": 64})
SOCK_QUERY = np.dtype([("fd", "<u8"), ("timestamp_ns", "<u8"), ("pid", "<u4"), ("_pad", "<u4")])
assert SOCK_QUERY.itemsize == 24
SOCK_RESULT = np.dtype([("found", "<u4"), ("saddr", "<u4"), ("daddr", "<u4"),
                        ("sport", "<u2"), ("dport", "<u2")])
assert SOCK_RESULT.itemsize == 16
ALIVE_CONN = np.dtype([("from_ip", "<u4"), ("from_id", "<u4"), ("to_ip", "<u4"), ("to_id", "<u4"),
                       ("from_port", "<u2"), ("to_port", "<u2"), ("to_type", "u1"), ("_pad", "u1", (3,))])
assert ALIVE_CONN.itemsize == 24

EDGE_OUT = np.dtype([
    ("from_type", "u1"), ("to_type", "u1"), ("_pad", "u1", (6,)),
    ("from", "<u4"), ("to", "<u4"),
    ("count", "<u8"), ("err5xx", "<u8"), ("lat_sum_ns", "<u8"),
    ("hist", "<u4", (NB,)),
])
assert EDGE_OUT.itemsize == 296


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32), ("device", C.c_int32),
        ("max_endpoints", C.c_uint32), ("max_pairs", C.c_uint32),
        ("max_edges", C.c_uint32), ("max_batch", C.c_uint32),
        ("flags", C.c_uint32), ("_reserved", C.c_uint32 * 9),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("events_in", C.c_uint64), ("rows_emitted", C.c_uint64),
        ("not_request", C.c_uint64), ("src_unresolved", C.c_uint64),
        ("pairs_live", C.c_uint64), ("edges_live", C.c_uint64),
        ("tcp_events_in", C.c_uint64), ("tcp_localhost_dropped", C.c_uint64),
        ("capacity_events", C.c_uint64), ("windows", C.c_uint64),
        ("kernel_launches", C.c_uint64), ("collective_bytes_last", C.c_uint64),
        ("flush_local_us_last", C.c_uint64), ("merge_us_last", C.c_uint64),
        ("late_events", C.c_uint64), ("deferred_events", C.c_uint64),
    ]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_ if not n.startswith("_")}


class SockStats(C.Structure):
    """alz_sock_stats_t."""
    _fields_ = [(n, C.c_uint64) for n in ("lines", "pool_records", "pool_garbage", "syncs", "sync_ops", "sync_bytes",
                                          "repools", "joined_events")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class SynthView(C.Structure):
    """alz_synth_view (alaz_b200/synth/alz_synth.h)."""
    _fields_ = [
        ("seed", C.c_uint64), ("t0_ns", C.c_uint64), ("dt_ns", C.c_uint32),
        ("mix", C.c_uint32), ("n_edges", C.c_uint32), ("n_unknown", C.c_uint32),
        ("unknown_base", C.c_uint32), ("_pad", C.c_uint32),
        ("edge_saddr", C.c_void_p), ("edge_daddr", C.c_void_p), ("edge_flags", C.c_void_p),
        ("alias_thresh", C.c_void_p), ("alias_idx", C.c_void_p), ("lat_q", C.c_void_p),
    ]


class SynthTopo(C.Structure):
    """alz_synth_topo (alaz_b200/synth/alz_synth.h)."""
    _fields_ = [
        ("n_services", C.c_uint32), ("n_pods", C.c_uint32),
        ("n_edges", C.c_uint32), ("n_outbound", C.c_uint32),
        ("pod_ip", C.POINTER(C.c_uint32)), ("svc_ip", C.POINTER(C.c_uint32)),
        ("out_ip", C.POINTER(C.c_uint32)),
        ("edge_saddr", C.POINTER(C.c_uint32)), ("edge_daddr", C.POINTER(C.c_uint32)),
        ("edge_flags", C.POINTER(C.c_uint8)),
        ("alias_thresh", C.POINTER(C.c_uint32)), ("alias_idx", C.POINTER(C.c_uint32)),
        ("lat_q", C.POINTER(C.c_uint64)),
        ("view", SynthView),
    ]


def ip(s: str) -> int:
    """Dotted quad -> the u32 the reference feeds IntToIPv4 (first octet in MSB)."""
    a, b, c, d = (int(x) for x in s.split("."))
    return (a << 24) | (b << 16) | (c << 8) | d
