"""pytest plugin for running the REFERENCE's own tests with the townlet_b200 drop-in classes patched in.

Loaded with ``-p b200_plugin`` from a writable copy of the reference (tests/test_reference_suite.py).  It swaps

* ``WorldObservationService`` (world/observations/service.py:27) for ``B200ObservationService`` wherever the
  reference binds the name (the service module and ``world/grid.py:309``),
* ``RewardEngine`` (rewards/engine.py:16) for ``B200RewardEngine`` (also where ``core/sim_loop.py:273`` binds it),
* ``WorldState._apply_need_decay`` (world/grid.py:1439) through ``install_need_decay`` after every construction,

before any test module is imported.  On the CPU build box the CUDA library is replaced by the C oracle through the
test-only ``OracleHostEngine`` (set ``TL_REFSUITE_ENGINE=cuda`` on a GPU box to run the kernels instead).
"""

import os
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[2]
for path in (REPO, REPO / "tests"):
    if str(path) not in sys.path:
        sys.path.insert(0, str(path))

import townlet_b200.observation_service as obs_mod  # noqa: E402
import townlet_b200.reward_engine as rew_mod  # noqa: E402

if os.environ.get("TL_REFSUITE_ENGINE", "oracle") != "cuda":
    from oracle_engine import OracleHostEngine  # noqa: E402

    obs_mod.HostEngine = OracleHostEngine
    rew_mod.HostEngine = OracleHostEngine

import townlet.rewards.engine as ref_rewards  # noqa: E402
import townlet.world.grid as ref_grid  # noqa: E402
import townlet.world.observations.service as ref_service  # noqa: E402

ref_service.WorldObservationService = obs_mod.B200ObservationService
ref_grid.WorldObservationService = obs_mod.B200ObservationService
ref_rewards.RewardEngine = rew_mod.B200RewardEngine
for name in ("townlet.rewards", "townlet.core.sim_loop"):
    try:
        module = __import__(name, fromlist=["RewardEngine"])
    except Exception:  # pragma: no cover - optional binding sites
        continue
    if hasattr(module, "RewardEngine"):
        module.RewardEngine = rew_mod.B200RewardEngine

_original_post_init = ref_grid.WorldState.__post_init__


def _post_init_with_device_decay(self, *args, **kwargs):
    _original_post_init(self, *args, **kwargs)
    rew_mod.install_need_decay(self)


ref_grid.WorldState.__post_init__ = _post_init_with_device_decay
