"""Device message tier (rings + persistent server kernel + credit counter).

The cases live in ``tests/rings_cases.py`` and run in a child interpreter with
``CUDA_MODULE_LOADING=EAGER``: a resident persistent kernel deadlocks against CUDA's lazy module loading
(the first launch of a not-yet-loaded kernel synchronises the context), and eager loading must be
selected before CUDA initialises -- which is not something a test inside a long-lived pytest process
can guarantee.  Keeping it out of the main process also keeps the main process's start-up lazy."""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_device_message_tier_cases_in_eager_child_process():
    env = dict(os.environ, CUDA_MODULE_LOADING="EAGER", PYTHONUNBUFFERED="1")
    cmd = [sys.executable, "-m", "pytest", os.path.join(REPO, "tests", "rings_cases.py"), "-x", "-q",
           "-p", "no:cacheprovider", "--timeout", "120", "--timeout-method=thread"]
    r = subprocess.run(cmd, cwd=REPO, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-2000:]
