"""Mutated filter sources and mutated IR text must be rejected with an error or accepted -- never crash the library, which
lives inside the host application (GIMP, the command line).  Each fuzzer runs in its own process so that a crash is a test
failure with an exit code, not the end of the test run (tools/fuzz_frontend.py, tools/fuzz_ir_loader.py take any seed / count)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tool,seed,count", [("fuzz_frontend.py", 11, 400), ("fuzz_ir_loader.py", 12, 400)])
def test_mutated_input_never_crashes(tool, seed, count):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), str(seed), str(count)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=900)
    assert r.returncode == 0, "exit code %d\n%s" % (r.returncode, r.stdout[-2000:])
    assert "rejected" in r.stdout
