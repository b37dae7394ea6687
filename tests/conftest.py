import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import cusift_b200 as csb
        ctx = csb.Context(0, 1)
        ctx.close()
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not errored) on a box without a GPU, unless they were asked for with -m gpu."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if gpu_items and not _has_gpu():
        skip = pytest.mark.skip(reason="no CUDA GPU on this box (run with -m gpu on the B200)")
        for it in gpu_items:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def gpu_ctx():
    """One product-library context for the whole GPU session (fails loudly without a GPU)."""
    import cusift_b200 as csb
    ctx = csb.Context(0, 4)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def workdir(tmp_path_factory):
    return tmp_path_factory.mktemp("csb_work")
