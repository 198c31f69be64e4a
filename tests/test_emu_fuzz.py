"""A bounded slice of the emulator fuzzers in the regular CPU suite: tools/fuzz_emulated*.py for a few seconds each with a
fixed seed (random sizes, PML depths, programs, pass depths, vector widths, row partitions, step splits, uploaded state,
DFT, slabs, fused halo exchange, streamed plans -- every configuration bit-for-bit against the numpy oracle).  Longer
runs with other seeds are a command away; what they found so far is listed in DESIGN.md section 2."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tool,seconds,least", [("fuzz_emulated_1d.py", 4, 50), ("fuzz_emulated.py", 7, 3),
                                               ("fuzz_emulated_2d_modes.py", 7, 3)])
def test_bounded_fuzz_finds_nothing(tool, seconds, least):
    from tests.emu import build_emu
    build_emu.build_library()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), str(seconds), "4"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    last = r.stdout.strip().splitlines()[-1]
    assert " 0 failures" in last, r.stdout[-3000:]
    assert int(last.split()[0]) >= least, last            # it really ran configurations
