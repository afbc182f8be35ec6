"""pytest configuration: the ``gpu`` marker and import paths."""

from __future__ import annotations

import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "tests" / "golden"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config: pytest.Config) -> None:
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config: pytest.Config, items: list[pytest.Item]) -> None:
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        # a kernel that never returns must end the run, not hang it: the watchdog thread exits the process (a signal
        # would wait for the blocked CUDA call to come back)
        for item in items:
            if "gpu" in item.keywords and item.get_closest_marker("timeout") is None:
                item.add_marker(pytest.mark.timeout(600, method="thread"))
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class KernelForm:
    """Kernel-selection switches of the library (``tl_set_option``) for the parity tests; reset after the test."""

    _NAMES = {
        "TL_MODE": ("kernel_form", lambda v: {"ent": 0, "grid": 1, "": -1}[v]),
        "TL_CHUNK": ("pass_agents", int),
        "TL_TPC": ("towns_per_cta", int),
        "TL_GENERIC": ("generic_window", lambda v: 1 if v else 0),
        "TL_NO_MIRROR_FUSE": ("mirror_pass", lambda v: 1 if v else 0),
        "TL_NO_EVEN_GRID": ("even_grid", lambda v: 0 if v else 1),
    }

    def setenv(self, name: str, value: str) -> None:
        """Same call shape as ``monkeypatch.setenv`` for the switches that used to be environment variables."""
        from townlet_b200 import capi

        option, convert = self._NAMES[name]
        capi.set_option(option, convert(value))


@pytest.fixture
def kernel_form():
    from townlet_b200 import capi

    yield KernelForm()
    capi.reset_options()
