"""Random worlds through the reference and through this repo's oracle / drop-in classes (build container only).

Small editions of the three fuzz modes of tests/golden/make_golden.py, which import the reference from a writable
copy: ``fuzz`` (oracle vs build_batch + decay + lifecycle + reward), ``fuzz-world`` (the world-dynamics phases vs
apply_actions + process_actions + RelationshipService.decay) and ``fuzz-dropin`` (B200ObservationService /
B200RewardEngine / install_need_decay beside the reference classes, metadata included).  Larger runs:
``TL_FUZZ_WORLDS=900 TL_GOLDEN_ONLY=fuzz python tests/golden/make_golden.py``.
"""

from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

import pytest

REF = Path(os.environ.get("TOWNLET_REFERENCE", "/root/reference"))
HERE = Path(__file__).resolve().parent
pytestmark = pytest.mark.skipif(not (REF / "src" / "townlet").exists(), reason="reference checkout not available")


@pytest.mark.parametrize("mode,worlds", [("fuzz", 45), ("fuzz-world", 30), ("fuzz-dropin", 30), ("fuzz-gae", 100)])
def test_random_worlds_agree_with_the_reference(mode, worlds) -> None:
    env = dict(os.environ, TL_GOLDEN_ONLY=mode, TL_FUZZ_WORLDS=str(worlds), TL_FUZZ_SEED="20261020")
    proc = subprocess.run([sys.executable, str(HERE / "golden" / "make_golden.py")], env=env, capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0, proc.stdout[-1500:] + proc.stderr[-3000:]
    assert f"{mode}: {worlds}" in proc.stdout, proc.stdout[-1500:]
