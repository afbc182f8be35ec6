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
