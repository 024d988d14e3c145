"""GPU: GNN anomaly pass (docs/SPEC.md §6) against the numpy float64 restatement, tolerance
1e-5 relative (+1e-6 absolute) on the edge scores as north_star states; quantiles 1e-12."""
import ctypes as C

import numpy as np
import pytest

import gnn_ref
import oracle_lib as ol
from alaz_b200 import abi, capi

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-6


def _run(S, N, mix, seed, max_pairs=1 << 17):
    t = ol.Topo(S, seed=seed, mix=mix)
    h = capi.Handle(max_endpoints=4 * S, max_pairs=max_pairs)
    h.load_tables(t.pod_ip, t.svc_ip)
    step = 8_000_000
    for i in range(0, N, step):
        h.submit(t.events(i, min(step, N - i)))
    edges = h.flush()
    scores = np.zeros(len(edges), dtype=np.float32)
    n = C.c_size_t(0)
    h._ck(h.L.alz_gnn_score(h.h, scores.ctypes.data_as(C.c_void_p), len(scores), C.byref(n)), "alz_gnn_score")
    assert n.value == len(edges)
    keys = np.zeros(2 * len(edges), dtype=np.uint64)
    h2 = np.zeros((2 * len(edges), 64), dtype=np.float32)
    nv = C.c_size_t(0)
    h._ck(h.L.alz_gnn_nodes(h.h, keys.ctypes.data_as(C.c_void_p), h2.ctypes.data_as(C.c_void_p), len(keys),
                            C.byref(nv)), "alz_gnn_nodes")
    h.close()
    return edges, scores, keys[: nv.value], h2[: nv.value]


@pytest.mark.parametrize("S,N,mix", [(200, 300_000, abi.MIX_ALL), (2000, 2_000_000, abi.MIX_SURVEY)])
def test_gnn_scores_match_float64_reference(S, N, mix):
    edges, scores, keys, h2 = _run(S, N, mix, seed=1000 + S)
    nodes, h2_ref, ref = gnn_ref.run(edges)
    assert np.array_equal(keys, nodes)
    assert np.allclose(h2, h2_ref, rtol=1e-4, atol=1e-4), float(np.abs(h2 - h2_ref).max())
    assert np.all(np.isfinite(scores))
    err = np.abs(scores.astype(np.float64) - ref)
    assert np.all(err <= ATOL + RTOL * np.abs(ref)), (float(err.max()), int(err.argmax()))
    assert scores.std() > 1e-4    # not a constant


def test_gnn_at_config3_scale():
    """BASELINE.json configs[2] names a 50k-service graph: most of a million edges, several hundred thousand nodes.
    The tensor-core layer runs thousands of 128-row tiles with gathers that miss L2; same tolerance."""
    edges, scores, keys, h2 = _run(50_000, 40_000_000, abi.MIX_SURVEY, seed=77, max_pairs=1 << 21)
    assert len(edges) > 500_000, len(edges)
    nodes, h2_ref, ref = gnn_ref.run(edges)
    assert np.array_equal(keys, nodes)
    assert np.allclose(h2, h2_ref, rtol=1e-4, atol=1e-4), float(np.abs(h2 - h2_ref).max())
    err = np.abs(scores.astype(np.float64) - ref)
    assert np.all(err <= ATOL + RTOL * np.abs(ref)), (float(err.max()), int(err.argmax()))


def test_gnn_on_empty_window_and_quantiles():
    h = capi.Handle(max_endpoints=64, max_pairs=256)
    h.commit()
    assert len(h.flush()) == 0
    n = C.c_size_t(7)
    h._ck(h.L.alz_gnn_score(h.h, None, 0, C.byref(n)), "alz_gnn_score")
    assert n.value == 0
    e = np.zeros(1, dtype=abi.EDGE_OUT)
    e["hist"][0][4] = 10
    e["hist"][0][5] = 10
    qs = np.array([0.25, 0.5, 0.75, 0.99, 1.0])
    out = np.zeros(5)
    assert h.L.alz_edge_quantiles(e.ctypes.data_as(C.c_void_p), qs.ctypes.data_as(C.c_void_p), 5,
                                  out.ctypes.data_as(C.c_void_p)) == 0
    ref = np.array([ol.quantile(e["hist"][0], q) for q in qs])
    assert np.allclose(out, ref, rtol=1e-12)
    h.close()
