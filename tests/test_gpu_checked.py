"""The bounds-checked twin of the library (DYS_CHECKS=1: every index the day kernels form is range-checked on the device;
a failed check ends the launch and surfaces as DYS_ERR_ABORTED with the check's code) runs the windowed workloads -- single
run, replicas, device lockdown policy, and (with two GPUs) the peer exchange -- without a single check firing.
compute-sanitizer is not available on the development pool; this is the stand-in SURVEY.md section 5 asks for."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _checked_lib():
    from dysturbd_b200 import _build
    return _build.build(checked=True)


def test_checked_build_single_gpu():
    env = dict(os.environ, DYS_LIB=_checked_lib(), DYS_SPIN_TIMEOUT_MS="60000")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_run.py"), "--version"], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "checked build" in r.stdout and "single-GPU windows ok" in r.stdout, r.stdout[-2000:]


def test_checked_build_two_gpu_peer_exchange():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, DYS_LIB=_checked_lib(), DYS_SPIN_TIMEOUT_MS="60000")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29800 + os.getpid() % 100), os.path.join(ROOT, "tools", "sanitize_run.py"), "--peer"]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "peer windows ok: mode peer" in r.stdout, r.stdout[-3000:]
