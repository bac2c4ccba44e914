"""CPU-only: the partitioner of the grid-resident solver (variant 3) keeps the invariants its kernel
relies on, for every graph size the benchmarks and the parity tests use.  Runs the library's own
host-side verifier (fb_grid_plan_verify) -- no device, no compute."""
import numpy as np
import pytest

from flame_ros_b200 import synth
from helpers import small_graph


@pytest.mark.parametrize("parts", [1, 2, 3, 7, 16])
def test_small_graph_tables(capi, parts):
    g = small_graph()
    rc, st = capi.grid_plan_verify(g["pos"], g["edges"], parts)
    assert rc == 0
    assert st["parts"] == parts
    if parts == 1:
        assert st["dup_edges"] == 0 and st["max_halo"] == 0 and st["boundary"] == 0
    else:
        assert st["dup_edges"] > 0 and st["max_halo"] > 0


@pytest.mark.parametrize("parts", [1, 2, 5, 16])
def test_small_graph_tables_cluster_transport(capi, parts):
    g = small_graph(20, 15, 160, 120, seed=9)
    rc, st = capi.grid_plan_verify(g["pos"], g["edges"], parts, cluster=True)
    assert rc == 0 and st["parts"] == parts


@pytest.mark.parametrize("cfg,parts,cluster", [("C2", 18, False), ("C2", 37, False), ("C2", 52, False),
                                               ("C4", 148, False), ("C4", 296, False),
                                               ("C2", 8, True), ("C2", 10, True), ("C2", 16, True)])
def test_benchmark_graph_tables(capi, cfg, parts, cluster):
    g = synth.s_graph(cfg)
    V, E = len(g["pos"]), len(g["edges"])
    rc, st = capi.grid_plan_verify(g["pos"], g["edges"], parts, cluster)
    assert rc == 0
    if cluster:
        assert st["max_edges"] <= 2048 and st["max_own"] <= 1024 and st["smem_bytes"] <= 200 * 1024
        return
    # compact parts: the cut stays a small fraction of the edges, the load is balanced
    assert st["dup_edges"] < 0.35 * E
    assert st["max_own"] <= 1.35 * V / parts + 8
    assert st["max_edges"] <= 1024 and st["max_own"] <= 512 and st["smem_bytes"] <= 100 * 1024


def test_too_few_parts_is_reported_not_mangled(capi):
    g = synth.s_graph("C2")
    rc, _ = capi.grid_plan_verify(g["pos"], g["edges"], 4)   # 1250 vertices per part > 512
    assert rc == 1
    g4 = synth.s_graph("C4")
    rc, _ = capi.grid_plan_verify(g4["pos"], g4["edges"], 16, cluster=True)   # 20k vertices need > 16 CTAs
    assert rc == 1


def test_degenerate_graphs(capi):
    one = np.array([[3.0, 4.0]], np.float32)
    rc, st = capi.grid_plan_verify(one, np.zeros((0, 2), np.int32), 1)
    assert rc == 0 and st["max_own"] == 1
    # more parts than vertices: empty parts are legal
    g = small_graph(4, 3, 64, 48, seed=1)
    rc, st = capi.grid_plan_verify(g["pos"], g["edges"], 16)
    assert rc == 0
    # isolated vertex + collinear duplicates of a coordinate
    pos = np.array([[0, 0], [10, 0], [20, 0], [30, 0], [5, 40]], np.float32)
    edges = np.array([[0, 1], [1, 2], [2, 3]], np.int32)
    for parts in (1, 2, 3):
        rc, _ = capi.grid_plan_verify(pos, edges, parts)
        assert rc == 0
