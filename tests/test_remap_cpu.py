"""CPU test of the integer index maps the attention kernel's TMA loads are built from (csrc/remap.cuh, host side): the
slot order, the four-box decomposition of a (shifted) window and the mask classes, against the oracle's restatement of
the reference's window_partition / roll / compute_mask (oracle/lrce_oracle.py, pinned to the reference's golden tables
in test_oracle_golden.py). The same functions run on the device: test_kernels_gpu.py checks them there."""
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

import lrce_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpu", "remap_check.cu")


@pytest.fixture(scope="module")
def remap_check(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("remap") / "remap_check")
    subprocess.run([nvcc, "-O1", "-std=c++17", "-o", exe, SRC], check=True)
    return exe


def _tables(exe, H, W, sh, sw):
    out = subprocess.run([exe, str(H), str(W), str(sh), str(sw)], check=True, capture_output=True, text=True).stdout
    rows = [np.array(l.split(), dtype=np.int64) for l in out.strip().split("\n")]
    return rows[0], rows[1], rows[2::3], rows[3::3], rows[4::3]


@pytest.mark.parametrize("H,W", [(56, 56), (28, 56), (14, 14), (7, 7)])
@pytest.mark.parametrize("shift", [(0, 0), (3, 3), (3, 0), (0, 3)])
def test_box_decomposition_matches_window_partition(remap_check, H, W, shift):
    if H == 7 and shift != (0, 0):
        pytest.skip("a clamped axis is never shifted (video_swin_ori.py:91-104)")
    key_slot, slot_tok, box_tok, src_tok, region = _tables(remap_check, H, W, *shift)
    # slot order: a bijection between the 147 window tokens and the 147 valid slots, 13 pads
    assert sorted(key_slot) == sorted(np.nonzero(slot_tok >= 0)[0]) and (slot_tok < 0).sum() == 13
    assert np.array_equal(slot_tok[key_slot], np.arange(147))
    gather = O.window_gather_index((3, H, W), (3, 7, 7), (0,) + shift).numpy()
    ids = O.shift_region_ids((3, H, W), (3, 7, 7), (0,) + shift).numpy() if any(shift) else None
    for win in range(gather.shape[0]):
        # closed-form gather == the oracle's window_partition(roll(x, -shift)) table (bit-exact)
        assert np.array_equal(src_tok[win], gather[win])
        # what the four TMA boxes deliver into slot s is exactly the source token of the window token living in slot s
        valid = slot_tok >= 0
        assert np.array_equal(box_tok[win][valid], gather[win][slot_tok[valid]])
        assert (box_tok[win][~valid] == -1).all()
        if ids is not None:
            # tokens attend each other iff their region ids agree: the kernel's ids induce the same partition as the oracle's
            same_k = region[win][:, None] == region[win][None, :]
            same_o = ids[win][:, None] == ids[win][None, :]
            assert np.array_equal(same_k, same_o)
            # ... and the mask is constant per (row class, key class): classes are the slot ranges [0,48) [48,88) [88,128) [128,160)
            cls = (key_slot >= 48).astype(int) + (key_slot >= 88) + (key_slot >= 128)
            for a in range(4):
                for b in range(4):
                    blk = same_o[np.ix_(cls == a, cls == b)]
                    assert blk.all() or not blk.any()
