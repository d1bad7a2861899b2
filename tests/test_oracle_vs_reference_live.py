"""Live randomised cross-check of the oracle against the unmodified reference (tests/golden/live_check.py).
Runs only where /root/reference exists (the build container); skipped on the GPU box.  A subprocess keeps the
reference's top-level modules (`impl`, `datasets`) out of this test session's import state."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("GLASS_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "impl")), reason="reference tree not present on this machine")
def test_oracle_matches_reference_on_random_configurations():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "live_check.py"), "10"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith(("ok", "FAIL"))]
    assert r.returncode == 0 and len(lines) == 10, r.stdout[-3000:] + r.stderr[-2000:]
