"""GPU: the C++ host adapter (alaz_b200/host) driven by the reference's own load-test scenario
(main_benchmark_test.go TestSimulation, testconfig/config1.json), built with g++ against libalazgpu.so."""
import os
import subprocess

import pytest

from alaz_b200 import build

pytestmark = pytest.mark.gpu


def test_simulation_through_cpp_adapter():
    exe = build.build_sim_test()
    r = subprocess.run([exe, "15"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "sim ok: 3000000 rows over 15 windows" in r.stdout
