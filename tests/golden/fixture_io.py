"""Golden fixture (de)serialisation shared by make_golden.py and the tests.

A fixture is one ``.npz``: the batch arena bytes + sizes, the observation /
reward specs as JSON, the phase mask and tick count, and the reference's
outputs (``ref_*`` arrays with a leading tick axis).
"""

from __future__ import annotations

import json
from dataclasses import asdict
from pathlib import Path

import numpy as np

from townlet_b200.params import DynamicsSpec, ObsSpec, RewardSpec
from townlet_b200.soa import BatchLayout, TownBatch

GOLDEN_DIR = Path(__file__).resolve().parent


def save_fixture(
    path: Path,
    batch: TownBatch,
    obs: ObsSpec,
    rew: RewardSpec,
    phases: int,
    n_ticks: int,
    ref: dict[str, np.ndarray],
    note: str = "",
    dyn: DynamicsSpec | None = None,
    actions: np.ndarray | None = None,
) -> None:
    lay = batch.layout
    meta = {
        "n_towns": lay.n_towns,
        "n_agents": lay.n_agents,
        "n_entities": lay.n_entities,
        "max_edges": lay.max_edges,
        "embed_words": lay.embed_words,
        "grid_w": batch.grid_w,
        "grid_h": batch.grid_h,
        "has_social_in": batch.has_social_in,
        "has_personality": batch.has_personality,
        "obs": asdict(obs),
        "rew": asdict(rew),
        "phases": int(phases),
        "n_ticks": int(n_ticks),
        "note": note,
    }
    if dyn is not None:
        meta["dyn"] = asdict(dyn)
    arrays = {f"ref_{k}": v for k, v in ref.items()}
    if actions is not None:
        arrays["in_actions"] = np.ascontiguousarray(actions, dtype=np.uint32)
    np.savez_compressed(path, arena=batch.arena, meta=np.array(json.dumps(meta)), **arrays)


def load_fixture(path: Path):
    data = np.load(path, allow_pickle=False)
    meta = json.loads(str(data["meta"]))
    layout = BatchLayout.make(meta["n_towns"], meta["n_agents"], meta["n_entities"], meta["max_edges"], meta["embed_words"])
    arena = np.array(data["arena"], dtype=np.uint8)
    legacy = arena.size < layout.total_bytes  # stored before TlTownBatch.ent_holder (the last array of the arena) existed
    if legacy:
        arena = np.concatenate([arena, np.zeros(layout.total_bytes - arena.size, dtype=np.uint8)])
    batch = TownBatch(layout, meta["grid_w"], meta["grid_h"], arena)
    if legacy:
        batch.ent_holder[...] = -2  # no reservation slots
    batch.has_social_in = meta["has_social_in"]
    batch.has_personality = meta["has_personality"]
    obs_kwargs = dict(meta["obs"])
    obs_kwargs["object_channels"] = tuple(obs_kwargs["object_channels"])
    obs = ObsSpec(**obs_kwargs)
    rew = RewardSpec(**meta["rew"])
    ref = {k[4:]: data[k] for k in data.files if k.startswith("ref_")}
    if "dyn" in meta:  # world-dynamics fixtures (tests/golden/world/)
        meta["dyn_spec"] = DynamicsSpec(**meta["dyn"])
        meta["actions"] = np.array(data["in_actions"]) if "in_actions" in data.files else None
    return batch, obs, rew, meta, ref


def fixture_paths() -> list[Path]:
    return sorted(GOLDEN_DIR.glob("*.npz"))


def world_fixture_paths() -> list[Path]:
    """Fixtures that also drive the world-dynamics phases (actions, ledger decay)."""
    return sorted((GOLDEN_DIR / "world").glob("*.npz"))
