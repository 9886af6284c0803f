"""Launch overlap must not change results.  Most kernels are launched with programmatic stream serialization (their CTAs become resident
while the predecessor drains); every kernel is deterministic, so a full-size 7B prefill whose launches execute back to back on the device
must give BITWISE the logits of the same prefill with every launch fully serialised (DLB200_NO_PDL=15).  This is the model-level guard
for the class of bug described in profiles/r2_stream_k_workspace_race.md.  The switch is read once per process, hence two child
processes (tools/pdl_bitwise.py, reduced configuration list)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(900)
def test_prefill_logits_bitwise_identical_with_and_without_launch_overlap(tmp_path):
    tool = os.path.join(ROOT, "tools", "pdl_bitwise.py")
    outs = []
    for tag, extra in (("on", {}), ("off", {"DLB200_NO_PDL": "15"})):
        out = str(tmp_path / f"pdl_{tag}.pt")
        env = dict(os.environ, PDL_BITWISE_QUICK="1", **extra)
        r = subprocess.run([sys.executable, tool, out], cwd=ROOT, env=env, capture_output=True, text=True, timeout=400)
        assert r.returncode == 0 and "saved 6" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
        outs.append(out)
    r = subprocess.run([sys.executable, tool, "--compare"] + outs, cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "6/6 configurations bitwise identical" in r.stdout, r.stdout[-2000:] + r.stderr[-1000:]
