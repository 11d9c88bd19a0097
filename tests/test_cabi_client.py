"""The C ABI used from plain C: tests/cabi_client.c compiles as C99 against include/vbsky_b200.h, links against
libvbsky_b200.so with gcc alone (CPU check) and, on a GPU, reproduces the oracle's numbers without Python in the loop."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GCC = "/usr/bin/gcc"


def _build(tmp_path):
    exe = str(tmp_path / "cabi_client")
    libdir = os.path.join(ROOT, "vbsky_b200")
    subprocess.check_call([GCC, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cabi_client.c"), "-o", exe, "-L", libdir, "-lvbsky_b200",
                           f"-Wl,-rpath,{libdir}"])
    return exe


def test_header_is_c99_and_client_links(tmp_path):
    import vbsky_b200  # noqa: F401  (the library must have been built)

    assert os.path.exists(_build(tmp_path))


@pytest.mark.gpu
def test_c_client_matches_oracle(tmp_path, cuda):
    import oracle.prune as oprune
    import oracle.substitution as osub
    import oracle.tree_data as otd

    out = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    got = np.array([[float(x) for x in line.split()] for line in out.stdout.strip().splitlines()])
    td = otd.TreeData(np.array([0, 1, 3, 2, 4]), np.array([3, 3, 4, 4, -1]), np.array([[-1, -1]] * 3 + [[1, 0], [3, 2]]), np.array([1.0, 0.0, 0.5]))
    masks = np.array([[1, 1, 1], [1, 2, 1], [4, 4, 8], [15, 1, 2]], dtype=np.uint8)
    part = ((masks[..., None] >> np.arange(4, dtype=np.uint8)) & 1).astype(float)
    counts = np.array([1.0, 2.0, 1.0, 3.0])
    for u, bl in enumerate(([0.1, 0.2, 0.25, 0.05], [0.3, 0.01, 0.02, 0.4])):
        tot, g, _ = oprune.data_loglik_and_grad(np.array(bl), osub.HKY(2.7), part, counts, td)
        assert got[u, 0] == pytest.approx(tot, rel=1e-9)
        np.testing.assert_allclose(got[u, 1:], g, rtol=1e-6)
