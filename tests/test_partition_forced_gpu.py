"""The two-level tile partition (binning.cu: partition_two_level) only serves tile counts above 2048 by
default. This test re-runs the batched-path identity tests (instance lists, tile ranges, images against the
per-view API; empty views; mask back-projection; the semantic channel) in a child process with the partition
FORCED on every tile count and 16-tile groups (DGE_PART2=2 DGE_PART2_SHIFT=4: the switches are read once per
process), so that small images exercise several tile groups, partial last groups and empty segments."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SELECT = "bit_identical or empty_views or backprojection or semantic"


def run_forced(shift):
    env = dict(os.environ, DGE_PART2="2", DGE_PART2_SHIFT=str(shift))
    return subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_fit_gpu.py"), "-x", "-q",
                           "-m", "gpu", "-k", SELECT, "-p", "no:cacheprovider"],
                          cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)


@pytest.mark.gpu
@pytest.mark.parametrize("shift", [4, 8])
def test_two_level_partition_forced_on_every_tile_count(cuda, shift):
    r = run_forced(shift)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "skipped" not in r.stdout, r.stdout[-500:]
