"""The convolution engine has run-time variants that the default configuration never takes on the bench network:
the run-time ("generic") epilogue, the single-issuer schedule, the two-CTAs-per-SM variant and the plain K3 path for
the final layer.  Each is selected by an environment variable read once per process, so the kernel parity tests are
re-run in a child process per variant."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = [
    {"B200SEG_TC_GENERIC_EPILOGUE": "1"},
    {"B200SEG_TC_ISSUERS": "1"},
    {"B200SEG_TC_ISSUERS": "3"},
    {"B200SEG_TC_VARIANT": "1"},
    {"B200SEG_TC_NO_K3T": "1"},
]


@pytest.mark.gpu
@pytest.mark.parametrize("env", VARIANTS, ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()))
def test_conv_parity_under_variant(env):
    child_env = dict(os.environ, **env)
    proc = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_kernels.py"),
                           os.path.join(ROOT, "tests", "test_gpu_models.py"), "-q", "-x", "-m", "gpu", "-k",
                           "conv_tc or network or patch_predict", "-p", "no:cacheprovider"],
                          cwd=ROOT, env=child_env, capture_output=True, text=True, timeout=900)
    tail = (proc.stdout + proc.stderr)[-2000:]
    assert proc.returncode == 0, f"variant {env} failed:\n{tail}"
    assert " passed" in proc.stdout, tail
