"""CPU: the C oracle (oracle/townlet_oracle.c) against the reference-generated golden vectors.

This pins the oracle: every fixture under tests/golden/ was produced by
running the reference itself (tests/golden/make_golden.py), four of them are
the reference's own stored goldens (baseline_{hybrid,full,compact}.npz,
social_snippet_gold.npz).
"""

from __future__ import annotations

import numpy as np
import pytest
from fixture_io import fixture_paths, load_fixture
from helpers import assert_kpi_close, assert_obs_equal, assert_refreshed_state, assert_reward_close

from oracle.oracle import oracle_tick
from townlet_b200 import capi

FIXTURES = fixture_paths()


def test_kpi_rows_are_pinned_to_the_reference() -> None:
    """Row K: the synthetic-town fixtures carry KPI rows produced by the reference's telemetry / stability classes."""
    pinned = [p for p in FIXTURES if "kpi" in load_fixture(p)[4]]
    assert len(pinned) >= 5
    _, _, _, _, ref = load_fixture(pinned[0])
    assert ref["kpi"].shape[-1] == capi.TL_N_KPI and ref["kpi"][..., capi.KPI_NAMES.index("starving")].max() >= 1


def test_fixtures_present() -> None:
    names = {p.stem for p in FIXTURES}
    for required in ("ref_baseline_hybrid", "ref_baseline_full", "ref_baseline_compact", "ref_social_snippet"):
        assert required in names


@pytest.mark.parametrize("path", FIXTURES, ids=[p.stem for p in FIXTURES])
def test_oracle_matches_reference(path) -> None:
    batch, obs, rew, meta, ref = load_fixture(path)
    batch.validate(len(obs.object_channels))
    out = oracle_tick(batch, obs, rew, meta["phases"], meta["n_ticks"])
    assert_obs_equal(out, ref, path.stem)
    assert_reward_close(out, ref, path.stem)
    assert_kpi_close(out, ref, batch, path.stem)
    assert_refreshed_state(batch, ref, path.stem)
    if "needs" in ref:  # state after the last tick
        np.testing.assert_allclose(batch.needs, ref["needs"][-1], rtol=1e-6, atol=1e-15)
    if "episode_total" in ref:
        np.testing.assert_allclose(batch.episode_total, ref["episode_total"][-1], rtol=1e-6, atol=1e-12)


def test_oracle_threads_agree() -> None:
    path = next(p for p in FIXTURES if p.stem == "synth_hybrid_48x8")
    batch, obs, rew, meta, _ = load_fixture(path)
    b1, b4 = batch.copy(), batch.copy()
    o1 = oracle_tick(b1, obs, rew, capi.TL_PHASE_ALL, 3, n_threads=1)
    o4 = oracle_tick(b4, obs, rew, capi.TL_PHASE_ALL, 3, n_threads=4)
    for key in o1:
        assert np.array_equal(o1[key], o4[key], equal_nan=True), key
    assert np.array_equal(b1.arena, b4.arena)
