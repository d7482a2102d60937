"""-m gpu: the drop-in proof.  oracle/_ref/dropin_test is tests/cpp/dropin_test.cu compiled (in the build container,
where /root/reference exists) against the UNMODIFIED MatX headers plus include/matx_b200/executor.h: every statement
runs on matx::cudaExecutor (the reference's CUB path) and on matx::b200Executor (libmatx_b200.so) and is compared."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "dropin_test")


def test_matx_statements_through_b200_executor_match_reference_cuda_executor():
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/dropin_test not built (needs /root/reference at build time)")
    r = subprocess.run([BIN], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    print(r.stdout)
    lines = r.stdout.splitlines()
    summary = [ln for ln in lines if ln.startswith("SUMMARY")]
    assert summary, r.stdout[-2000:]
    fails = [ln for ln in lines if ln.startswith("FAIL")]
    assert not fails and r.returncode == 0, "\n".join(fails)
    # the statements were served by native kernels, not by the fallback
    native = [ln for ln in lines if ln.startswith("PASS") and "kernel=red_" in ln or "kernel=ew" in ln or "kernel=var_" in ln]
    assert len(native) >= 15, r.stdout
    fallback = [ln for ln in lines if "shift" in ln]
    assert fallback and "(reference fallback)" in fallback[0]
