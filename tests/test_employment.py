"""Employment shift state machine: the C oracle against the reference-generated trajectory
(tests/golden/employment/ref_employment.npz, produced by running EmploymentEngine.apply_job_state for 250 ticks), and the
CUDA step against the fixture and against the oracle on larger random towns."""

from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import pytest

from townlet_b200 import capi
from townlet_b200.employment import EmploymentArrays, EmploymentSpec
from townlet_b200.soa import BatchLayout, TownBatch, pack_pos

FIXTURE = Path(__file__).resolve().parent / "golden" / "employment" / "ref_employment.npz"
ORIGIN = (-3, -2)  # the town's origin in world coordinates: job locations are given in world coordinates


def _load():
    d = np.load(FIXTURE, allow_pickle=False)
    raw = json.loads(str(d["spec"]))
    raw["job_location"] = tuple(tuple(loc) if loc is not None else None for loc in raw["job_location"])
    spec = EmploymentSpec(**{k: tuple(v) if isinstance(v, list) else v for k, v in raw.items()})
    return d, spec


def _town(n_agents: int, wallet: np.ndarray) -> TownBatch:
    batch = TownBatch(BatchLayout.make(1, n_agents, n_agents), 16, 16)
    batch.town_agent_off[:] = [0, n_agents]
    batch.town_ent_off[:] = [0, n_agents]
    batch.wallet[:] = wallet
    return batch


def _set_tick(batch: TownBatch, tick: int, pos_world: np.ndarray) -> None:
    batch.town_tick[0] = tick
    batch.pos[:] = pack_pos(pos_world[:, 0] - ORIGIN[0], pos_world[:, 1] - ORIGIN[1])


def _compare(d, t: int, batch: TownBatch, arrays: EmploymentArrays, events: np.ndarray, what: str) -> None:
    for name, _, _ in capi.EMPLOYMENT_STATE_FIELDS:
        want, got = d[f"emp_{name}"][t], arrays.arrays[name]
        if name in ("shift_started_tick", "shift_end_tick"):  # meaningful while a shift is open
            mask = (d["emp_ctx_flags"][t] & (1 << 10)) != 0
            want, got = want[mask], got[mask]
        elif name == "last_present_tick":
            mask = (d["emp_ctx_flags"][t] & (1 << 11)) != 0
            want, got = want[mask], got[mask]
        elif name == "attendance_samples":
            live = np.arange(want.shape[1])[None, :] < d["emp_attendance_count"][t][:, None]
            want, got = np.where(live, want, 0.0), np.where(live, got, 0.0)
        elif name == "absence_events":
            live = np.arange(want.shape[1])[None, :] < d["emp_absence_count"][t][:, None]
            want, got = np.where(live, want, 0), np.where(live, got, 0)
        assert np.array_equal(want, got), f"{what} tick {t}: context key {name}: {got} vs {want}"
    # float64 results of the same operations in the same order: bit-exact
    for key, got in (("wallet", batch.wallet), ("wages_withheld", batch.wages_withheld), ("attendance", batch.attendance),
                     ("wages_paid", batch.wages_paid), ("punctuality", batch.punctuality)):
        assert np.array_equal(d[key][t], got), f"{what} tick {t}: {key}: {got} vs {d[key][t]}"
    assert np.array_equal(d["lateness"][t], batch.lateness), f"{what} tick {t}: lateness_counter"
    flags = batch.flags
    assert np.array_equal(d["on_shift"][t], flags & 1), f"{what} tick {t}: on_shift"
    assert np.array_equal(d["shift_code"][t], (flags >> capi.TL_F_SHIFT_SHIFT) & 7), f"{what} tick {t}: shift state"
    assert np.array_equal(d["events"][t].astype(np.uint32), events.astype(np.uint32)), f"{what} tick {t}: events"


def test_oracle_follows_the_reference_tick_by_tick() -> None:
    from oracle.oracle import oracle_employment_step

    d, spec = _load()
    n_ticks, n = d["pos"].shape[:2]
    batch = _town(n, d["initial_wallet"])
    arrays = EmploymentArrays(n)
    arrays.job[:] = d["initial_job"]
    origin = np.array([ORIGIN], dtype=np.int32)
    for t in range(n_ticks):
        _set_tick(batch, t, d["pos"][t])
        events = oracle_employment_step(batch, spec, arrays, 1, 0, origin)
        _compare(d, t, batch, arrays, events[0], "oracle")
    assert set(np.unique(d["emp_state"])) == {1, 2, 3, 4, 5, 6} and int(d["events"].max()) > 0


@pytest.mark.gpu
def test_kernel_follows_the_reference_tick_by_tick() -> None:
    import torch

    from townlet_b200.employment import DeviceEmployment
    from townlet_b200.engine import DeviceTownBatch

    d, spec = _load()
    n_ticks, n = d["pos"].shape[:2]
    batch = _town(n, d["initial_wallet"])
    arrays = EmploymentArrays(n)
    arrays.job[:] = d["initial_job"]
    dev = DeviceTownBatch(batch, None, None)
    emp = DeviceEmployment(dev, spec, arrays, town_origin=np.array([ORIGIN], dtype=np.int32))
    for t in range(n_ticks):
        _set_tick(batch, t, d["pos"][t])
        dev.field("town_tick").copy_(torch.from_numpy(batch.town_tick))
        dev.field("pos").copy_(torch.from_numpy(batch.pos))
        events = emp.step(1, tick_offset=0)
        torch.cuda.synchronize()
        state = dev.download_state(full=True)
        _compare(d, t, state, emp.download(), events[0].cpu().numpy(), "kernel")


@pytest.mark.gpu
def test_kernel_equals_oracle_on_random_towns_and_multi_tick_launches() -> None:
    import torch

    from oracle.oracle import oracle_employment_step
    from townlet_b200.employment import DeviceEmployment
    from townlet_b200.engine import DeviceTownBatch
    from townlet_b200.params import ObsSpec
    from townlet_b200.synth import synth_batch

    rng = np.random.default_rng(9)
    spec = EmploymentSpec(
        job_ids=("a", "b", "c"), job_start=(10, 25, 40), job_end=(30, 45, 40), job_wage=(0.02, 0.03, 0.01), job_lateness_penalty=(0.1, 0.2, 0.0),
        job_location=((3, 3), None, (7, 2)), arrival_buffer=4, ticks_per_day=60, grace_ticks=2, absent_cutoff=9, absence_slack=4,
        attendance_window=3, late_tick_penalty=0.005, absence_penalty=0.2,
    )
    batch = synth_batch(300, 16, 7, ObsSpec())
    batch.town_tick[:] = rng.integers(0, 50, batch.n_towns)
    arrays = EmploymentArrays(batch.n_agents)
    arrays.job[:] = rng.integers(-1, 3, batch.n_agents)
    ref_batch, ref_arrays = batch.copy(), arrays.copy()
    dev = DeviceTownBatch(batch, None, None)
    emp = DeviceEmployment(dev, spec, arrays)
    seen_states, seen_events = set(), 0
    for step in range(40):  # positions change between launches; a launch steps 1-5 ticks
        n_ticks = int(rng.integers(1, 6))
        x, y = rng.choice([3, 7, 1], batch.n_agents), rng.choice([3, 2, 5], batch.n_agents)
        ref_batch.pos[:] = pack_pos(x, y)
        dev.field("pos").copy_(torch.from_numpy(ref_batch.pos))
        events = emp.step(n_ticks, tick_offset=0)
        want = oracle_employment_step(ref_batch, spec, ref_arrays, n_ticks, 0)
        ref_batch.town_tick[:] += n_ticks  # the tick kernel would advance it
        dev.field("town_tick").copy_(torch.from_numpy(ref_batch.town_tick))
        assert np.array_equal(events.cpu().numpy().view(np.uint32), want), step
        seen_states |= set(np.unique(ref_arrays.state).tolist())
        seen_events |= int(np.bitwise_or.reduce(want, axis=None))
    state, got = dev.download_state(full=True), emp.download()
    for name in ("wallet", "wages_withheld", "wages_paid", "punctuality", "attendance", "lateness", "flags"):
        assert np.array_equal(getattr(state, name), getattr(ref_batch, name)), name
    for name, _, _ in capi.EMPLOYMENT_STATE_FIELDS:
        assert np.array_equal(got.arrays[name], ref_arrays.arrays[name]), name
    assert len(seen_states) >= 5 and seen_events == 31 and (got.attendance_count > 0).any() and (got.absence_count > 0).any()
