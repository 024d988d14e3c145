"""CPU: the C-ABI library loads, exports every symbol include/alazgpu.h declares,
the Python mirror matches the C layouts, and the product fails loudly without a GPU."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from alaz_b200 import abi, capi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "alazgpu.h")


def _declared_functions():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(alz_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    lib = build.build()
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    declared = _declared_functions()
    assert declared, "no declarations parsed"
    missing = [f for f in declared if f not in exported]
    assert not missing, f"declared in alazgpu.h but not exported: {missing}"
    assert sorted(capi.EXPORTS) == declared


def test_struct_layouts_match_header():
    prog = r"""
    #include <stdio.h>
    #include <stddef.h>
    #include "alazgpu.h"
    int main(void){
      printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(alz_l7_rec), sizeof(alz_tcp_rec), sizeof(alz_sock_query),
             sizeof(alz_sock_result), sizeof(alz_edge_out), sizeof(alz_config), sizeof(alz_stats));
      printf("%zu %zu %zu %zu\n", offsetof(alz_l7_rec,status), offsetof(alz_l7_rec,duration_ns),
             offsetof(alz_edge_out,count), offsetof(alz_edge_out,hist));
      return 0; }
    """
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        l1, l2 = subprocess.check_output([exe], text=True).strip().splitlines()
    import ctypes as C
    sizes = [int(x) for x in l1.split()]
    assert sizes == [abi.L7_REC.itemsize, abi.TCP_REC.itemsize, abi.SOCK_QUERY.itemsize,
                     abi.SOCK_RESULT.itemsize, abi.EDGE_OUT.itemsize, C.sizeof(abi.Config), C.sizeof(abi.Stats)]
    offs = [int(x) for x in l2.split()]
    assert offs == [abi.L7_REC.fields["status"][1], abi.L7_REC.fields["duration_ns"][1],
                    abi.EDGE_OUT.fields["count"][1], abi.EDGE_OUT.fields["hist"][1]]


def test_owner_rank_is_a_pure_function_of_saddr():
    L = capi.load()
    rng = np.random.default_rng(0)
    ips = rng.integers(0, 2**32, 2000, dtype=np.uint64)
    for n in (1, 2, 4, 8):
        r = np.array([L.alz_owner_rank(int(x), n) for x in ips])
        assert r.min() >= 0 and r.max() < n
        if n > 1:
            assert len(set(r.tolist())) == n
    assert all(L.alz_owner_rank(int(x), 1) == 0 for x in ips[:10])


def test_host_generator_in_product_lib_equals_oracle_lib_copy():
    import oracle_lib as ol
    a = capi.Topo(64, seed=99, mix=abi.MIX_ALL)
    b = ol.Topo(64, seed=99, mix=abi.MIX_ALL)
    assert np.array_equal(a.pod_ip, b.pod_ip) and np.array_equal(a.svc_ip, b.svc_ip)
    assert a.events(5, 5000).tobytes() == b.events(5, 5000).tobytes()
    a.close()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.AlzError) as e:
        capi.Handle()
    assert e.value.status == abi.E_NODEVICE


def test_product_never_references_the_oracle():
    # the oracle is a checker: nothing under alaz_b200/ or include/ may name it
    bad = []
    for base in ("alaz_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp", ".cc")):
                    s = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"oracle_lib|alz_oracle|liboracle|orc_process|ref_py", s):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_cpp_host_adapter_conversions_and_loud_failure():
    exe = build.build_host_unit_test()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host unit ok" in r.stdout


def test_pack_l7_is_lossless_for_what_the_reduce_reads():
    """alz_pack_l7 is host code (no GPU needed): 32-B records -> 16-B records + overflow durations (docs/SPEC.md §10)."""
    import ctypes as C
    rng = np.random.default_rng(3)
    n = 5000
    ev = np.zeros(n, dtype=abi.L7_REC)
    ev["saddr"] = rng.integers(0, 2**32, n)
    ev["daddr"] = rng.integers(0, 2**32, n)
    ev["sport"] = rng.integers(0, 65536, n)
    ev["dport"] = rng.integers(0, 65536, n)
    ev["status"] = rng.integers(0, 65536, n)
    ev["protocol"] = rng.integers(0, 256, n)
    ev["method_flags"] = rng.integers(0, 256, n)
    dur = rng.integers(0, 2**32, n).astype(np.uint64)
    big = rng.random(n) < 0.1
    dur[big] = rng.integers(2**32, 2**63, int(big.sum())).astype(np.uint64)      # >= 4.29 s: side array
    dur[0], dur[1] = 2**32 - 1, 2**32                                            # the boundary
    big[0], big[1] = False, True
    ev["duration_ns"] = dur
    ev["write_time_ns"] = rng.integers(0, 2**63, n)
    r16, ovf = capi.pack_l7(ev)
    assert len(ovf) == int(big.sum())
    assert np.array_equal(r16["saddr"], ev["saddr"]) and np.array_equal(r16["daddr"], ev["daddr"])
    assert np.array_equal(r16["status"], ev["status"]) and np.array_equal(r16["method_flags"], ev["method_flags"])
    flag = (r16["protocol"] & abi.REC16_DUR_OVERFLOW) != 0
    assert np.array_equal(flag, big)
    back = r16["duration_ns"].astype(np.uint64)
    back[flag] = ovf[r16["duration_ns"][flag]]                                   # the index into the side array
    assert np.array_equal(back, dur)
    # protocol: bits 0..5 the value, bit 6 ALZ_PROTO_F_HOSTKEY travels; a value with bit 7 set is no protocol at all
    p16 = r16["protocol"] & 0x7F
    hi = (ev["protocol"] & 0x80) != 0
    assert np.array_equal(p16[~hi], ev["protocol"][~hi] & 0x7F)
    assert np.all((p16[hi] & 0x3F) == 0x3F)
    # a side array that is too small is an error, not a truncation
    L = capi.load()
    out = np.zeros(n, dtype=abi.L7_REC16)
    small = np.zeros(1, dtype=np.uint64)
    assert L.alz_pack_l7(ev.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p), small.ctypes.data_as(C.c_void_p), 1) == -1
