"""The kernel-experiment libraries of `make variants` (pbrt_v2_b200/ab/) and the product library under its development
switches (the kernels the default path replaced), each against the GPU parity suite in its own process.  They are NOT the
product path (nothing selects them by default), so their outcome is reported without deciding the suite: xpassed = the
variant reproduces the oracle on the GPU, xfailed = it does not."""
import os
import subprocess
import sys

import pytest

import pbrt_v2_b200 as pb

pytestmark = pytest.mark.gpu

# (library or None = the product library, extra environment, test files)
CASES = [
    ("pairs", "ab/libpbrtgpu_pairs.so", {}, ["tests/test_gpu_parity.py"]),
    # what the default path (vertex wavefront sorted by next action + k_vertex<2>) replaced or did not adopt stays
    # selectable (A/B baselines): every combination must still reproduce the oracle
    ("vertex-wavefront-unsorted", None, {"PBRTGPU_WSORT": "0"}, ["tests/test_gpu_parity.py", "tests/test_gpu_specular.py"]),
    ("vertex-wavefront+chain-wavefront", None, {"PBRTGPU_WAVE": "3"}, ["tests/test_gpu_parity.py", "tests/test_gpu_specular.py"]),
    ("request-wavefront+vertex2", None, {"PBRTGPU_WAVE": "0"}, ["tests/test_gpu_parity.py", "tests/test_gpu_specular.py"]),
    ("request-wavefront+chain-wavefront", None, {"PBRTGPU_WAVE": "2"}, ["tests/test_gpu_parity.py"]),
    ("request-wavefront+k_path2", None, {"PBRTGPU_WAVE": "0", "PBRTGPU_VERTEX": "0"}, ["tests/test_gpu_parity.py"]),
    ("megakernel+k_path2", None, {"PBRTGPU_WAVE": "0", "PBRTGPU_VERTEX": "0", "PBRTGPU_WAVEFRONT": "0"}, ["tests/test_gpu_parity.py"]),
    ("vertex1+vertex2", None, {"PBRTGPU_WAVE": "0", "PBRTGPU_VERTEX": "3"}, ["tests/test_gpu_parity.py"]),
]


@pytest.mark.xfail(strict=False, reason="kernel experiment / A-B baseline: not the product path")
@pytest.mark.parametrize("name,lib,extra_env,files", CASES, ids=[c[0] for c in CASES])
def test_variant_library_reproduces_the_oracle(name, lib, extra_env, files):
    if lib and not os.path.exists(os.path.join(pb.REPO, "pbrt_v2_b200", lib)):
        pytest.skip("variant library not built (make variants)")
    env = dict(os.environ, PBRTGPU_SCRATCH_GB="16", **extra_env)
    if lib:
        env["PBRTGPU_LIB"] = lib
    r = subprocess.run([sys.executable, "-m", "pytest"] + files + ["-x", "-q", "-m", "gpu", "-p", "no:cacheprovider"],
                       cwd=pb.REPO, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    tail = "\n".join(r.stdout.splitlines()[-15:])
    print("\n[%s]\n%s" % (name, tail))
    assert r.returncode == 0, tail
