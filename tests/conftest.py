import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _ensure_library():
    """libfegnn.so is a build artefact (git-ignored): a fresh checkout has none.  Build it once (nvcc cross-compiles
    sm_100a without a GPU, ~40 s) so that importing fastegnn_b200 -- which has no fallback -- works in the tests."""
    lib = os.path.join(ROOT, "fastegnn_b200", "_C", "libfegnn.so")
    src = os.path.join(ROOT, "fastegnn_b200", "csrc")
    newest = max(os.path.getmtime(os.path.join(src, f)) for f in os.listdir(src))
    if os.path.exists(lib) and os.path.getmtime(lib) >= newest:
        return                          # up to date (a stale library silently tests yesterday's kernels)
    try:
        import __graft_entry__ as entry
        entry.build()
    except Exception as exc:            # the tests that need the library then fail with its own loud ImportError
        print(f"conftest: could not build libfegnn.so: {exc}", file=sys.stderr)


_ensure_library()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
