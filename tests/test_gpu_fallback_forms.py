"""-m gpu: the round-1 forms of the layout kernels stay in the library as the fall-back for unaligned buffers / odd channel
counts (and as the A/B reference, ASN_GLUE=0).  They are selected once per process, so the oracle parity tests of the two
kernel families are re-run in a child process with ASN_GLUE=0: both forms must pass the same gates."""
import os
import subprocess
import sys

from gpu_util import gpu

pytestmark = gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_round1_layout_kernels_pass_the_same_parity_gates():
    env = dict(os.environ, ASN_GLUE="0", PYTHONWARNINGS="ignore")
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "--no-header", "-p", "no:cacheprovider",
           os.path.join(ROOT, "tests", "test_gpu_aspp.py") + "::test_aspp_oracle",
           os.path.join(ROOT, "tests", "test_gpu_fcd.py") + "::test_fcd_oracle"]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-1000:]
