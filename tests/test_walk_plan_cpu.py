"""Host-side logic of the cluster-sharded encoder walk (csrc/encoder_walk.cu) on the CPU: the row plan lrce_encoder_walk
derives (rows per cluster, groups, ring slots, shared memory) and the size of the packed weight buffer — pure host
arithmetic of the C ABI, no GPU call."""
import ctypes
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lrce_b200  # noqa: E402
from lrce_b200 import _lib  # noqa: E402


def plan(rows, max_clusters):
    out = [ctypes.c_int() for _ in range(5)]
    rc = _lib.lib().lrce_encoder_walk_plan(rows, max_clusters, *[ctypes.addressof(o) for o in out])
    return rc, [o.value for o in out]


@pytest.mark.parametrize("max_clusters", [1, 2, 4, 7, 8, 9])
def test_walk_plan_covers_every_row(max_clusters):
    for rows in list(range(1, 130)) + [160, 161, 255, 256, 1000]:
        rc, (rpc, groups, clusters, slots, smem) = plan(rows, max_clusters)
        assert rc == 0, (rows, max_clusters)
        assert 1 <= rpc <= 8                                  # N = 8 columns of the swap-AB MMA
        assert groups * rpc >= rows > (groups - 1) * rpc      # every row in exactly one group, no empty group
        assert clusters == min(groups, max_clusters)
        assert 3 <= slots <= 8 and smem <= 232448             # the ring never starves; fits the 227 KB of a B200 CTA
        passes = -(-groups // clusters)
        assert passes == -(-rows // (8 * max_clusters))       # no more passes over the weights than 8 rows per cluster force


def test_walk_plan_configs_of_the_reference():
    # configs[1] / [2]: 32 clips per GPU on the 7 clusters a B200 keeps resident -> one pass, 5 rows per cluster
    assert plan(32, 7)[1][:3] == [5, 7, 7]
    # configs[3] (multiple choice, 5 candidates): 160 rows -> 3 passes of 8 rows per cluster
    assert plan(160, 7)[1][:3] == [8, 20, 7]
    # a single row (a B=1 request) runs on one cluster
    assert plan(1, 7)[1][:3] == [1, 1, 1]


def test_walk_plan_rejects_bad_arguments():
    assert plan(0, 7)[0] != 0 and plan(8, 0)[0] != 0
    assert b"lrce_encoder_walk_plan" in _lib.lib().lrce_last_error()


@pytest.mark.parametrize("n_layers,n_out", [(12, 1000), (12, 1500), (12, 1), (2, 4096)])
def test_walk_pack_bytes(n_layers, n_out):
    """weight stream (16 ranks x 6336 rows of 128 B per layer) + head stream (64-row tiles of the 16-way split, zero padded)
    + parameter blocks (672 fp32 per layer and rank) + gamma3 / beta3 of the last layer + head bias"""
    mt = -(-(-(-n_out // 16)) // 64)
    want = n_layers * 16 * 6336 * 128 + 16 * mt * 12 * 64 * 128 + n_layers * 16 * 672 * 4 + 2 * 768 * 4 + 16 * 64 * mt * 4
    assert _lib.lib().lrce_encoder_walk_pack_bytes(n_layers, n_out) == want
    assert want % 128 == 0
