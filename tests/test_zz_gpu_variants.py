"""The kernel-experiment libraries of `make variants` (pbrt_v2_b200/ab/) against the GPU parity suite, each in its own
process with PBRTGPU_LIB pointing at it.  They are CPU-verified experiments that have not been timed or run on hardware
yet and are NOT part of the product (nothing loads them by default), so their outcome is reported without deciding the
suite: xpassed = the variant reproduces the oracle on the GPU, xfailed = it does not (yet)."""
import os
import subprocess
import sys

import pytest

import pbrt_v2_b200 as pb

pytestmark = pytest.mark.gpu

CASES = [
    ("pairs", ["tests/test_gpu_parity.py"], []),
    ("tables2", ["tests/test_gpu_parity.py"], []),
    ("kdrec", ["tests/test_gpu_parity.py"], []),
    ("pairs_tables2", ["tests/test_gpu_parity.py"], []),
    # the specular build must still reproduce every non-specular scene, and render mirror / glass like the oracle
    ("specular", ["tests/test_gpu_parity.py", "tests/test_gpu_specular.py"],
     ["--deselect", "tests/test_gpu_parity.py::test_specular_materials_are_refused_loudly"]),
]


@pytest.mark.xfail(strict=False, reason="kernel experiment: CPU-verified, first hardware run; not part of the product")
@pytest.mark.parametrize("name,files,extra", CASES, ids=[c[0] for c in CASES])
def test_variant_library_reproduces_the_oracle(name, files, extra):
    lib = os.path.join("ab", "libpbrtgpu_%s.so" % name)
    if not os.path.exists(os.path.join(pb.REPO, "pbrt_v2_b200", lib)):
        pytest.skip("variant library not built (make variants)")
    env = dict(os.environ, PBRTGPU_LIB=lib, PBRTGPU_SCRATCH_GB="16")
    r = subprocess.run([sys.executable, "-m", "pytest"] + files + ["-x", "-q", "-m", "gpu", "-p", "no:cacheprovider"] + extra,
                       cwd=pb.REPO, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    tail = "\n".join(r.stdout.splitlines()[-15:])
    print("\n[%s]\n%s" % (name, tail))
    assert r.returncode == 0, tail
