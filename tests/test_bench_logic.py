"""Host-side logic of bench.py that can be checked without a GPU: how timed steps are grouped into CUDA graphs."""
import math
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


@pytest.mark.parametrize("K,W0", [(20, 5), (4000, 50), (320, 24), (64, 8), (8, 3), (7, 3), (100, 50), (2, 1)])
def test_graph_replay_keeps_the_rotation(K, W0):
    """Replaying graphs of `chunk` steps in bench.py's order visits, step by step, the same source set / emitter set / bus buffer /
    reduce buffer as a plain loop over the step index would (the pipelined form depends on it: the launch that streams block k has
    prepared block k + 1 for the NEXT launch, whichever graph that is in)."""
    chunk = bench.steps_per_graph(K)
    assert K % chunk == 0
    _check_rotation(K, W0, chunk)


@pytest.mark.parametrize("K,chunk", [(20, 20), (20, 10), (20, 5), (4000, 32), (60, 3)])
def test_longer_graphs_keep_the_rotation_too(K, chunk):
    _check_rotation(K, 5, chunk)


def _check_rotation(K, W0, chunk):
    n_graphs = bench.graphs_per_cycle(chunk)
    W = max(3, W0, bench.N_SETS, n_graphs * chunk)
    W = ((W + chunk - 1) // chunk) * chunk
    graphs = [[g * chunk + j for j in range(chunk)] for g in range(n_graphs)]
    assert (n_graphs * chunk) % bench.ROTATION == 0
    replayed = []
    for k0, n in ((0, W), (W, K)):
        for k in range(k0, k0 + n, chunk):
            replayed += graphs[(k // chunk) % len(graphs)]
    assert len(replayed) == W + K
    assert {(k // chunk) % len(graphs) for k in range(0, W, chunk)} == set(range(len(graphs)))  # the warm-up launches every graph
    for actual, captured in enumerate(replayed):
        for period in (bench.N_SETS, bench.NB, 2):
            assert actual % period == captured % period
            assert (actual + 1) % period == (captured + 1) % period  # the block each step prepares


def test_steps_per_graph_choices():
    assert bench.steps_per_graph(20) == 4
    assert bench.steps_per_graph(4000) == 8
    assert bench.steps_per_graph(7) == 1
    assert bench.graphs_per_cycle(20) == 2 and bench.graphs_per_cycle(32) == 1 and bench.graphs_per_cycle(5) == 8 and bench.graphs_per_cycle(1) == 8
    assert math.gcd(bench.ROTATION, bench.NB) == bench.NB
