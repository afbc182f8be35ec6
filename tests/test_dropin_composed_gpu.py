"""GPU + reference on ONE box: the drop-in classes with the real CUDA engine beside the real reference classes.

The CPU suite runs these comparisons with the oracle standing in for the kernels (tests/test_dropin_reference.py);
the GPU suite runs the kernels on fake worlds.  Here both halves meet: ``B200ObservationService`` /
``B200RewardEngine`` / ``install_need_decay`` call ``libtownlet_b200.so`` through ``HostEngine`` while
``WorldObservationService`` / ``RewardEngine`` / ``WorldState._apply_need_decay`` of the unmodified reference run on
twin worlds — observation dictionaries with metadata equality, reward breakdowns, decayed needs.  The reference comes
from ``oracle/_ref`` (``oracle/make_ref.sh``; git-ignored, ships with the gpurun snapshot) or ``/root/reference``.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

HERE = Path(__file__).resolve().parent
_candidates = [Path(os.environ["TOWNLET_REFERENCE"])] if os.environ.get("TOWNLET_REFERENCE") else []
_candidates += [Path("/root/reference"), HERE.parent / "oracle" / "_ref"]
REF = next((c for c in _candidates if (c / "src" / "townlet").is_dir()), None)
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(REF is None, reason="no reference tree (oracle/make_ref.sh)")]


@pytest.fixture(scope="module")
def ref(tmp_path_factory):
    """A writable copy of the reference on sys.path; the drop-in classes keep their real (CUDA) HostEngine."""
    work = tmp_path_factory.mktemp("ref")
    shutil.copytree(REF, work / "ref", symlinks=True)
    os.system(f"chmod -R u+w {work / 'ref'}")
    cwd = os.getcwd()
    os.chdir(work / "ref")
    sys.path.insert(0, str(work / "ref" / "src"))
    import townlet_b200.observation_service as obs_mod
    from townlet_b200.engine import HostEngine

    assert obs_mod.HostEngine is HostEngine, "the composed test must run the CUDA engine"
    yield work
    os.chdir(cwd)
    sys.path.remove(str(work / "ref" / "src"))
    for name in [m for m in sys.modules if m == "townlet" or m.startswith("townlet.")]:
        del sys.modules[name]  # later test modules run without the reference, as on a box that has none


# the comparisons themselves are the CPU suite's, collected here against the fixture above
from test_dropin_reference import (  # noqa: E402,F401
    test_need_decay_hook_matches_reference,
    test_observation_service_matches_reference,
    test_reference_stored_goldens_through_the_service,
    test_relationship_stage_required,
    test_reward_engine_matches_reference,
)


def test_random_twin_worlds_through_the_cuda_engine() -> None:
    """``fuzz-dropin`` of tests/golden/make_golden.py with the CUDA engine: random worlds (unbounded coordinates, stacked
    agents, pending resets, every variant) through the reference classes and the drop-in classes, a few ticks each."""
    env = dict(os.environ, TL_GOLDEN_ONLY="fuzz-dropin", TL_FUZZ_WORLDS="40", TL_FUZZ_SEED="20261021", TL_DROPIN_ENGINE="cuda",
               TOWNLET_REFERENCE=str(REF))
    proc = subprocess.run([sys.executable, str(HERE / "golden" / "make_golden.py")], env=env, capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0, proc.stdout[-1500:] + proc.stderr[-3000:]
    assert "fuzz-dropin: 40" in proc.stdout, proc.stdout[-1500:]
