import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF = os.environ.get("CUTRACE_REF", "/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The plain-C restatement (oracle/cutrace_oracle.c), built on demand."""
    from oracle import pyoracle as po

    if not po.have_oracle():
        po.build()
    return po


def load_golden_scene(name):
    from cutrace_b200.scene import FlatScene

    return FlatScene.load(os.path.join(GOLDEN, "scenes", f"{name}.npz"))


GOLDEN_CASES = {
    "triangle": "triangle_20x20.npz",
    "sphere_plane": "sphere_plane_160x90.npz",
    "mirror": "mirror_160x90.npz",
    "bunny": "bunny_96x54.npz",
}
