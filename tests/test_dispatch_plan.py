"""CPU: the host-side dispatch policy of libmatx_b200 (view collapsing, kernel-family choice, launch geometry) on the
BASELINE.json configurations at full size, through the library's plan-only mode (MXB_PLAN_ONLY=1: no CUDA call, nothing
launched, nothing computed — see api.cu).  Runs in a subprocess so the mode never leaks into a process that computes."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def plans(built_lib):
    env = dict(os.environ, MXB_PLAN_ONLY="1")
    for k in list(env):
        if k.startswith("MXB_TUNE") or k.startswith("MXB_VAR") or k.startswith("MXB_OUTER"):
            del env[k]
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "plan_probe.py")], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def geom(s):
    return {k: int(v) for k, v in (kv.split("=") for kv in s.split("|") if "=" in kv)}


def test_headline_config_is_a_split_row_reduction_on_a_persistent_grid(plans):
    k = plans["c2.sum"]
    assert k.startswith("red_inner|") and "|sum|f32|V8|U4|T0|aot" in k, k          # 32-byte loads, four in flight
    assert geom(k) == {"grid": 148 * 4, "block": 256, "smem": 0}, k        # one resident wave (4 CTAs per SM at 64 registers), chunks dealt dynamically
    assert "|argmax|" in plans["c2.argmax"] and plans["c2.argmax"].startswith("red_inner|")


def test_rows_pick_warp_or_cta_teams(plans):
    assert "|T1|aot" in plans["c1.fma_sum"] and plans["c1.fma_sum"].startswith("red_inner|")      # 16 KB rows: a warp per row
    assert "|T0|aot" in plans["c3.mean"]                                                          # 64 KB rows: a CTA per row
    assert "|T1|aot" in plans["c3.argmax_abs2"] and "|argmax|" in plans["c3.argmax_abs2"]         # arg ops: warp team up to 128 KB
    assert "|T1|" in plans["rowsum.short64"]


def test_variance_routing(plans):
    assert plans["c3.var"].startswith("var_tma|"), plans["c3.var"]                 # 64 KB rows of a plain tensor: TMA-staged two-pass
    g = geom(plans["c3.var"])
    assert g["smem"] > 128 * 1024 and g["grid"] == 148, plans["c3.var"]            # >= 2 row buffers, one CTA per SM
    assert plans["rowvar.1024"].startswith("var_group|"), plans["rowvar.1024"]     # rows in registers, lanes per row
    assert plans["rowvar.short64"].startswith("red_inner|") and "|var|" in plans["rowvar.short64"]   # 16..128 elements: one-pass op
    assert plans["rowvar.4096x65536"].startswith("red_inner|") and "|var|" in plans["rowvar.4096x65536"]   # 256 KB rows: one read
    assert plans["full.var"].startswith("red_inner|") and "|var|f32|V4|U4|T0|aot" in plans["full.var"]
    assert plans["colvar.4096x65536"].startswith("red_outer_tma|") and "|var|" in plans["colvar.4096x65536"]


def test_strided_reduce_dims_take_the_tma_tiles(plans):
    k = plans["c5.permuted_sum"]
    assert k.startswith("red_outer_tma|") and "|sum|bf16|V8|U1|T0|aot" in k, k
    g = geom(k)
    # 2 CTAs per SM, 256 consumer threads + the producer warp, 3 stages of 32 KB + 128 B of mbarriers + the partials
    assert g["grid"] == 296 and g["block"] == 288 and 3 * 32768 + 128 <= g["smem"] <= 112 * 1024, k
    assert plans["colsum.4096x65536"].startswith("red_outer_tma|")
    assert plans["colsum.tall"].startswith("red_outer|")          # 1000 columns: too few strips for the ring -> LDG walker with split R


def test_tma_eligibility_edges(plans):
    assert plans["tma.narrow_rows"].startswith("red_outer|"), plans["tma.narrow_rows"]            # < 32 chunks per row: LDG walker
    assert plans["tma.short_reduce_dim"].startswith("red_outer|"), plans["tma.short_reduce_dim"]  # fewer than 64 reduce rows
    k = plans["tma.fused_expression"]
    assert k.startswith("red_outer|") and "|V4|" in k, k                                          # expressions keep the LDG walker
    k = plans["tma.unaligned"]
    assert not k.startswith("red_outer_tma") and "|V1|" in k, k                                   # no 16-byte alignment: scalar walk


def test_elementwise_scan_and_select_families(plans):
    assert plans["ew.vector_add"].startswith("ew|") and plans["ew.vector_add"].split("|grid")[0].endswith("aot")
    assert plans["ew_tr.permute"].startswith("ew_tr|")
    assert plans["scan.rows"].startswith("scan|")
    k = plans["find.values"]
    # 1-D view: the single-pass kernel over warp tiles, one resident wave (3 CTAs per SM), static shared memory only
    assert k.startswith("select|") and "|f32|V4|U4|T3|aot" in k and geom(k) == {"grid": 148 * 3, "block": 256, "smem": 0}, k
    k = plans["find.strided_idx"]
    assert k.startswith("select|") and "|i32|V1|U4|T4|" in k, k    # every other column of a matrix collapses to ONE strided dim: scalar walk, single pass
