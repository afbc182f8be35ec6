"""The REFERENCE's own hot-path tests, run unmodified against the drop-in classes.

Where the reference checkout exists (the build container), copy it to a writable place and run its pytest files for
this path — observation baselines / builders / social snippet, reward engine, need decay, lifecycle, the world suites,
the ``SimulationLoop`` suites under tests/core and the rollout-capture suites — with ``tests/refsuite/b200_plugin.py`` loaded first.  The plugin swaps
``WorldObservationService``, ``RewardEngine`` and ``WorldState._apply_need_decay`` for ``B200ObservationService``,
``B200RewardEngine`` and ``install_need_decay`` wherever the reference binds those names.  On the CPU
box the engine underneath is the C oracle (host logic under test); with ``TL_REFSUITE_ENGINE=cuda`` on a GPU box it is
the CUDA library.  The only failure allowed is the one the reference has on a clean checkout (the missing
``tmp/wp-c/phase4_baseline_shapes.json`` of tests/test_observation_builder_parity.py).
"""

from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

REF = Path(os.environ.get("TOWNLET_REFERENCE", "/root/reference"))
HERE = Path(__file__).resolve().parent
pytestmark = pytest.mark.skipif(not (REF / "src" / "townlet").exists(), reason="reference checkout not available")

REFERENCE_TEST_FILES = (
    "tests/test_observation_baselines.py",
    "tests/test_observation_builder.py",
    "tests/test_observation_builder_compact.py",
    "tests/test_observation_builder_full.py",
    "tests/test_observation_builder_parity.py",
    "tests/test_observations_social_snippet.py",
    "tests/test_reward_engine.py",
    "tests/test_reward_dto.py",
    "tests/test_need_decay.py",
    "tests/test_lifecycle_manager.py",
    "tests/world",  # incl. test_world_context_observe.py
    "tests/core",  # SimulationLoop with dummy ports, DTO parity across ticks
    "tests/test_rollout_capture.py",
    "tests/test_rollout_buffer.py",
    "tests/policy/test_training_orchestrator_capture.py",
)
KNOWN_REFERENCE_FAILURE = "test_observation_builder_matches_baseline_snapshot"


def test_reference_hot_path_tests_pass_with_the_dropins(tmp_path) -> None:
    work = tmp_path / "ref"
    shutil.copytree(REF, work, symlinks=True)
    os.system(f"chmod -R u+w {work}")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([str(HERE / "refsuite"), str(work / "src"), env.get("PYTHONPATH", "")])
    env.setdefault("TL_REFSUITE_ENGINE", "oracle")
    proc = subprocess.run(
        [sys.executable, "-m", "pytest", "-p", "b200_plugin", "-p", "no:cacheprovider", "-o", "addopts=", "-q",
         *REFERENCE_TEST_FILES],
        cwd=work, env=env, capture_output=True, text=True, timeout=900,
    )
    tail = proc.stdout[-3000:]
    summary = re.search(r"(?:(\d+) failed, )?(\d+) passed", proc.stdout)
    assert summary, tail + proc.stderr[-2000:]
    failed, passed = int(summary.group(1) or 0), int(summary.group(2))
    failures = re.findall(r"^FAILED (\S+)", proc.stdout, flags=re.M)
    assert all(KNOWN_REFERENCE_FAILURE in name for name in failures), tail
    assert failed <= 1 and passed >= 160, tail


@pytest.mark.skipif(not os.environ.get("TL_REFSUITE_FULL"), reason="set TL_REFSUITE_FULL=1: the reference's whole suite takes ~3 min")
def test_whole_reference_suite_passes_with_the_dropins(tmp_path) -> None:
    """Every test file of the reference (as its CI runs them: ``-k "not rollout_capture"``, .github/workflows/ci.yml)
    with the drop-in classes patched in — 893 passed, 2 skipped when this was written."""
    work = tmp_path / "ref"
    shutil.copytree(REF, work, symlinks=True)
    os.system(f"chmod -R u+w {work}")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([str(HERE / "refsuite"), str(work / "src"), env.get("PYTHONPATH", "")])
    env.setdefault("TL_REFSUITE_ENGINE", "oracle")
    proc = subprocess.run(
        [sys.executable, "-m", "pytest", "-p", "b200_plugin", "-p", "no:cacheprovider", "-o", "addopts=", "-q",
         "--deselect", f"tests/test_observation_builder_parity.py::{KNOWN_REFERENCE_FAILURE}", "-k", "not rollout_capture", "tests"],
        cwd=work, env=env, capture_output=True, text=True, timeout=2400,
    )
    summary = re.search(r"(?:(\d+) failed, )?(\d+) passed", proc.stdout)
    assert summary and not summary.group(1) and int(summary.group(2)) >= 850, proc.stdout[-3000:]
