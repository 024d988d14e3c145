"""CPU, world_size 2 over gloo: the multi-rank window merge as specified in SURVEY.md §8e /
alaz_b200/csrc/alz_comm.cu — shard events by alz_owner_rank(saddr), reduce each shard
independently, build the canonical key list by all-gather + sort + unique, scatter into a
zeroed canonical array and ONE all-reduce(sum). The result on every rank must equal the
single-rank oracle, bit for bit. (The shards are reduced by the CPU oracle here; the CUDA
shards are covered by tests/test_gpu_multi.py.)"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def pack_key(e):
    """alz_device.cuh make_edge_key on the oracle's (from,to) representation."""
    ft, f, tt, t = int(e["from_type"]), int(e["from"]), int(e["to_type"]), int(e["to"])
    if ft == 0:          # pod is From (canonical for pod->pod too)
        rev, pod, ot, ov = 0, f, tt, t
    else:                # reversed row whose From is a service / outbound host
        rev, pod, ot, ov = 1, t, ft, f
    return (rev << 63) | (ot << 61) | (pod << 32) | ov


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    from alaz_b200 import abi, capi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = capi.load()
    t = ol.Topo(300, seed=17, mix=abi.MIX_ALL)
    ev = t.events(0, 200_000)
    owner = np.array([L.alz_owner_rank(int(s), world) for s in np.unique(ev["saddr"])])
    own_of = dict(zip(np.unique(ev["saddr"]).tolist(), owner.tolist()))
    mine = np.array([own_of[int(s)] == rank for s in ev["saddr"]])
    o = ol.Oracle()
    o.load_tables(t.pod_ip, t.svc_ip)
    o.process(ev[mine])
    local = o.edges()
    keys = np.array([pack_key(e) for e in local], dtype=np.uint64)
    order = np.argsort(keys)
    keys, local = keys[order], local[order]
    # 1. counts, 2. padded key all-gather, sort + unique
    cnt = torch.tensor([len(keys)], dtype=torch.int64)
    cnts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    pad = int(max(c.item() for c in cnts))
    send = np.full(pad, np.iinfo(np.uint64).max, dtype=np.uint64)
    send[: len(keys)] = keys
    bufs = [torch.zeros(pad, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(bufs, torch.from_numpy(send.view(np.int64)))
    allk = np.concatenate([b.numpy().view(np.uint64) for b in bufs])
    can = np.unique(allk[allk != np.iinfo(np.uint64).max])
    # 3. scatter into the zeroed canonical array [n_can x 35] u64 (hist packed 2 x u32)
    arr = np.zeros((len(can), 35), dtype=np.uint64)
    pos = np.searchsorted(can, keys)
    arr[pos, 0], arr[pos, 1], arr[pos, 2] = local["count"], local["err5xx"], local["lat_sum_ns"]
    h = local["hist"].astype(np.uint64)
    arr[pos, 3:] = h[:, 0::2] | (h[:, 1::2] << np.uint64(32))
    # 4. the single all-reduce
    tt = torch.from_numpy(arr.view(np.int64))
    dist.all_reduce(tt, op=dist.ReduceOp.SUM)
    merged = tt.numpy().view(np.uint64)
    q.put((rank, can.tobytes(), merged.tobytes(), int(mine.sum())))
    dist.destroy_process_group()


def test_two_rank_merge_equals_single_rank_oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    from alaz_b200 import abi
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=180) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == res[1][1] and res[0][2] == res[1][2], "ranks disagree after the all-reduce"
    assert 0 < res[0][3] and 0 < res[1][3] and res[0][3] + res[1][3] == 200_000
    t = ol.Topo(300, seed=17, mix=abi.MIX_ALL)
    o = ol.Oracle()
    o.load_tables(t.pod_ip, t.svc_ip)
    o.process(t.events(0, 200_000))
    exp = o.edges()
    ek = np.array([pack_key(e) for e in exp], dtype=np.uint64)
    order = np.argsort(ek)
    ek, exp = ek[order], exp[order]
    can = np.frombuffer(res[0][1], dtype=np.uint64)
    merged = np.frombuffer(res[0][2], dtype=np.uint64).reshape(len(can), 35)
    assert np.array_equal(can, ek)
    assert np.array_equal(merged[:, 0], exp["count"])
    assert np.array_equal(merged[:, 1], exp["err5xx"])
    assert np.array_equal(merged[:, 2], exp["lat_sum_ns"])
    h = exp["hist"].astype(np.uint64)
    assert np.array_equal(merged[:, 3:], h[:, 0::2] | (h[:, 1::2] << np.uint64(32)))
