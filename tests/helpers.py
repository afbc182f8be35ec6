"""Shared comparison helpers for oracle / kernel parity tests."""

from __future__ import annotations

import numpy as np

from townlet_b200 import capi

REL_TOL = 1e-6  # north_star: need and reward floats within 1e-6 relative


def assert_obs_equal(got: dict, ref: dict, what: str) -> None:
    """Maps bit-exact; features bit-exact as float32; integer summaries exact."""
    if "map" in ref:
        g, r = np.asarray(got["map"]), np.asarray(ref["map"])
        assert g.shape == r.shape, (what, g.shape, r.shape)
        bad = np.nonzero(g.view(np.uint32) != r.view(np.uint32))
        assert bad[0].size == 0, f"{what}: {bad[0].size} map cells differ, first at {tuple(b[0] for b in bad)}"
    if "features" in ref:
        g, r = np.asarray(got["features"]), np.asarray(ref["features"])
        assert g.shape == r.shape, (what, g.shape, r.shape)
        bad = np.nonzero(g.view(np.uint32) != r.view(np.uint32))
        if bad[0].size:
            idx = tuple(b[0] for b in bad)
            raise AssertionError(
                f"{what}: {bad[0].size} features differ, first at {idx}: got {g[idx]!r} ref {r[idx]!r}"
            )
    if "summary" in ref:
        g, r = np.asarray(got["summary"]), np.asarray(ref["summary"])
        assert np.array_equal(g[..., :6], r[..., :6]), f"{what}: summary rows differ"


def assert_reward_close(got: dict, ref: dict, what: str) -> None:
    if "comps" in ref and "comps" in got:
        g, r = np.asarray(got["comps"]), np.asarray(ref["comps"])
        np.testing.assert_allclose(g[..., :15], r[..., :15], rtol=REL_TOL, atol=1e-12, err_msg=f"{what}: reward components")
        flags = [capi.COMP_NAMES.index("guardrail_active"), capi.COMP_NAMES.index("clipped")]
        assert np.array_equal(g[..., flags], r[..., flags]), f"{what}: guardrail/clipped flags"
    if "reward" in got and "comps" in ref:
        np.testing.assert_allclose(
            np.asarray(got["reward"], dtype=np.float64), np.asarray(ref["comps"])[..., 0], rtol=REL_TOL, atol=1e-9, err_msg=f"{what}: reward f32"
        )
    if "terminated" in ref and "terminated" in got:
        assert np.array_equal(np.asarray(got["terminated"]), np.asarray(ref["terminated"])), f"{what}: terminated flags"


def assert_kpi_close(got: dict, ref: dict, batch, what: str) -> None:
    """Per-town KPI rows against the rows the reference's own telemetry / stability code produced (SURVEY.md §8a row K;
    tests/golden/make_golden.py: reference_kpi_row).  The rows are float32 of float64 accumulations: 1e-5 relative.  The
    reward mean and POPULATION variance the reference's RewardVarianceAnalyzer reports follow from the sum, the sum of
    squares and the town's agent count."""
    if "kpi" not in ref or "kpi" not in got:
        return
    g, r = np.asarray(got["kpi"], dtype=np.float64), np.asarray(ref["kpi"], dtype=np.float64)
    assert g.shape == r.shape, (what, g.shape, r.shape)
    for k, name in enumerate(capi.KPI_NAMES):
        np.testing.assert_allclose(g[..., k], r[..., k], rtol=1e-5, atol=1e-7, err_msg=f"{what}: KPI {name}")
    count = np.diff(np.asarray(batch.town_agent_off)).astype(np.float64)
    mean = g[..., capi.KPI_NAMES.index("reward_sum")] / count
    variance = g[..., capi.KPI_NAMES.index("reward_sumsq")] / count - mean**2
    np.testing.assert_allclose(mean, ref["kpi_reward_mean"], rtol=1e-5, atol=1e-7, err_msg=f"{what}: reward mean")
    np.testing.assert_allclose(variance, ref["kpi_reward_variance"], rtol=1e-4, atol=1e-7, err_msg=f"{what}: reward variance")


def assert_refreshed_state(batch, ref: dict, what: str) -> None:
    """TL_PHASE_RESERVATIONS fixtures: the entity words and the holders' flag after the launch equal a second ingest of the
    reference world after its own refresh_reservations."""
    if "ent_after" not in ref:
        return
    assert np.array_equal(np.asarray(batch.ent), ref["ent_after"][-1]), f"{what}: entity words after the refresh"
    bit = capi.TL_F_RESERVATION
    assert np.array_equal(np.asarray(batch.flags) & bit, ref["flags_after"][-1] & bit), f"{what}: TL_F_RESERVATION after the refresh"
