"""Shared fixtures.

* ``oracle``  -- the reference's own sources compiled over the Embree shim
                 (oracle/_ref/liboracle_ref.so).  TEST INFRASTRUCTURE: only tests,
                 ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may load it.
* ``emu``     -- test-only host emulation of the device code (tests/emu), CPU tests only.
* ``qz``      -- the product: host library over the CUDA C ABI (GPU tests only).
"""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from quetzalcoatlus_b200.harness import Harness  # noqa: E402

ORACLE_LIB = ROOT / "oracle" / "_ref" / "liboracle_ref.so"
EMU_LIB = ROOT / "tests" / "emu" / "_build" / "libqz_emu_harness.so"
REFERENCE = Path("/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _make(directory: Path) -> None:
    subprocess.run(["make", "-C", str(directory)], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)


@pytest.fixture(scope="session")
def oracle() -> Harness:
    if REFERENCE.exists():
        _make(ROOT / "oracle")  # no-op when up to date
    if not ORACLE_LIB.exists():
        pytest.skip("oracle/_ref/liboracle_ref.so is not built and /root/reference is absent")
    return Harness(ORACLE_LIB, "orc_")


@pytest.fixture(scope="session")
def emu() -> Harness:
    # QZ_EMU_LIB: another build of the emulation, e.g. one with -fsanitize=address,undefined (tests/emu/Makefile: sanitize)
    override = os.environ.get("QZ_EMU_LIB")
    if override:
        return Harness(Path(override), "qzh_")
    _make(ROOT / "tests" / "emu")
    return Harness(EMU_LIB, "qzh_")


@pytest.fixture(scope="session")
def qz() -> Harness:
    from quetzalcoatlus_b200 import load_harness

    return load_harness()


@pytest.fixture(scope="session")
def small_mesh(tmp_path_factory) -> str:
    """A 4000-triangle instance of the synthetic obj_viewer mesh (tools/gen_mesh.py: the bulky "dragon" body that fills
    a third of the obj_viewer frame; the full 1M-triangle instance is pinned by tests/golden/paths_obj_viewer_1m.npz)."""
    sys.path.insert(0, str(ROOT / "tools"))
    import gen_mesh

    path = tmp_path_factory.mktemp("mesh") / "dragon_small.obj"
    pos, nrm, tris = gen_mesh.dragon_mesh(4000)
    gen_mesh.write_obj(str(path), pos, nrm, tris)
    return str(path)
