"""Employment shift state machine on the device (``tl_employment_step``) — host-side mirror.

The reference's ``EmploymentEngine.apply_job_state`` (src/townlet/world/employment.py:63-142) keeps one context
dictionary per agent (``context_defaults``, :167-193).  Here every key of that dictionary is one array
(``EmploymentArrays`` ↔ ``TlEmploymentState``), the job table and the employment knobs are ``EmploymentSpec``
(↔ ``TlEmploymentParams``), and ``DeviceEmployment`` runs the step on a ``DeviceTownBatch``.  The two coworker ledger
updates and event emission stay on the host: the step reports them as per-agent event bits (``EVENT_BITS``).
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any

import numpy as np

from . import capi

TL_MAX_JOBS = capi.TL_MAX_JOBS
TL_MAX_ATTENDANCE_WINDOW = capi.TL_MAX_ATTENDANCE_WINDOW
TL_MAX_ABSENCE_EVENTS = capi.TL_MAX_ABSENCE_EVENTS

STATE_CODE = {"pre_shift": 0, "idle": 1, "await_start": 2, "on_time": 3, "late": 4, "absent": 5, "post_shift": 6}
# boolean context keys -> bits of ctx_flags (include/townlet_b200.h: TL_EMP_F_*)
FLAG_BITS = {
    "late_penalty_applied": 1 << 0,
    "absence_penalty_applied": 1 << 1,
    "late_event_emitted": 1 << 2,
    "absence_event_emitted": 1 << 3,
    "departure_event_emitted": 1 << 4,
    "late_help_event_emitted": 1 << 5,
    "took_shift_event_emitted": 1 << 6,
    "late_counter_recorded": 1 << 7,
    "ever_on_time": 1 << 8,
    "shift_outcome_recorded": 1 << 9,
}
FLAG_SHIFT_STARTED, FLAG_LAST_PRESENT = 1 << 10, 1 << 11
EVENT_BITS = {
    "shift_late_start": 1,
    "shift_absent": 2,
    "shift_departed_early": 4,
    "employment_helped_when_late": 8,
    "employment_took_my_shift": 16,
}
# snapshot.shift_state -> the flag word's shift code (what the feature one-hot reads, encoders/features.py:151-169)
SHIFT_FLAG_CODE = {"on_time": 1, "late": 2, "absent": 3, "post_shift": 4}


TlEmploymentParams = capi.TlEmploymentParams
STATE_FIELDS = capi.EMPLOYMENT_STATE_FIELDS
TlEmploymentState = capi.TlEmploymentState


@dataclass(frozen=True)
class EmploymentSpec:
    """The job table and the employment knobs of a ``SimulationConfig`` (what ``apply_job_state`` reads, :63-70)."""

    job_ids: tuple[str, ...] = ()
    job_start: tuple[int, ...] = ()
    job_end: tuple[int, ...] = ()
    job_wage: tuple[float, ...] = ()
    job_lateness_penalty: tuple[float, ...] = ()
    job_location: tuple[tuple[int, int] | None, ...] = ()
    arrival_buffer: int = 20
    ticks_per_day: int = 1440
    grace_ticks: int = 5
    absent_cutoff: int = 30
    absence_slack: int = 20
    attendance_window: int = 3
    late_tick_penalty: float = 0.005
    absence_penalty: float = 0.2

    @staticmethod
    def from_config(config: Any) -> "EmploymentSpec":
        jobs = config.jobs
        emp = config.employment
        default_wage = float(config.economy.get("wage_income", 0.0))
        ids = tuple(jobs.keys())
        if len(ids) > TL_MAX_JOBS:
            raise ValueError(f"at most {TL_MAX_JOBS} jobs")
        start = tuple(int(jobs[j].start_tick) for j in ids)
        # end = spec.end_tick or spec.start_tick, never before the start (employment.py:92-96)
        end = tuple(max(int(jobs[j].end_tick or jobs[j].start_tick), int(jobs[j].start_tick)) for j in ids)
        return EmploymentSpec(
            job_ids=ids,
            job_start=start,
            job_end=end,
            job_wage=tuple(float(jobs[j].wage_rate or default_wage) for j in ids),
            job_lateness_penalty=tuple(float(jobs[j].lateness_penalty) for j in ids),
            job_location=tuple(tuple(int(c) for c in jobs[j].location) if jobs[j].location else None for j in ids),
            arrival_buffer=int(config.behavior.job_arrival_buffer),
            ticks_per_day=max(1, int(emp.exit_review_window)),
            grace_ticks=int(emp.grace_ticks),
            absent_cutoff=int(emp.absent_cutoff),
            absence_slack=int(emp.absence_slack),
            attendance_window=max(1, int(emp.attendance_window)),
            late_tick_penalty=float(emp.late_tick_penalty),
            absence_penalty=float(emp.absence_penalty),
        )

    def struct(self) -> TlEmploymentParams:
        p = TlEmploymentParams()
        p.n_jobs = len(self.job_ids)
        p.arrival_buffer, p.ticks_per_day = self.arrival_buffer, self.ticks_per_day
        p.grace_ticks, p.absent_cutoff, p.absence_slack = self.grace_ticks, self.absent_cutoff, self.absence_slack
        p.attendance_window = self.attendance_window
        p.late_tick_penalty, p.absence_penalty = self.late_tick_penalty, self.absence_penalty
        for k in range(len(self.job_ids)):
            p.job_start[k], p.job_end[k] = self.job_start[k], self.job_end[k]
            p.job_wage[k], p.job_lateness_penalty[k] = self.job_wage[k], self.job_lateness_penalty[k]
            loc = self.job_location[k]
            p.job_has_location[k] = 0 if loc is None else 1
            if loc is not None:
                p.job_x[k], p.job_y[k] = loc
        return p


@dataclass
class EmploymentArrays:
    """One array per key of the reference's employment context, ``n_agents`` rows (host copy)."""

    n_agents: int
    arrays: dict[str, np.ndarray] = field(default_factory=dict)

    def __post_init__(self) -> None:
        if not self.arrays:
            for name, dtype, shape in STATE_FIELDS:
                self.arrays[name] = np.zeros((self.n_agents, *shape), dtype=dtype)
            self.arrays["job"][...] = -1
            self.arrays["current_day"][...] = -1  # context_defaults, employment.py:171

    def __getattr__(self, name: str) -> np.ndarray:
        arrays = self.__dict__.get("arrays")
        if arrays is not None and name in arrays:
            return arrays[name]
        raise AttributeError(name)

    def copy(self) -> "EmploymentArrays":
        return EmploymentArrays(self.n_agents, {k: v.copy() for k, v in self.arrays.items()})

    def struct(self, pointers: dict[str, int], town_origin_ptr: int | None) -> TlEmploymentState:
        s = TlEmploymentState()
        for name, _, _ in STATE_FIELDS:
            setattr(s, name, pointers[name])
        s.town_origin = town_origin_ptr
        return s

    def load_context(self, row: int, ctx: dict[str, Any], snapshot: Any, spec: EmploymentSpec) -> None:
        """Row ``row`` from a reference context dictionary and the agent's snapshot (``EmploymentEngine._state``)."""
        a = self.arrays
        job_id = getattr(snapshot, "job_id", None)
        a["job"][row] = spec.job_ids.index(job_id) if job_id in spec.job_ids else -1
        a["state"][row] = STATE_CODE[ctx["state"]]
        a["current_day"][row] = ctx["current_day"]
        a["late_ticks"][row] = ctx["late_ticks"]
        flags = 0
        for key, bit in FLAG_BITS.items():
            if ctx[key]:
                flags |= bit
        if ctx["shift_started_tick"] is not None:
            flags |= FLAG_SHIFT_STARTED
            a["shift_started_tick"][row] = ctx["shift_started_tick"]
            a["shift_end_tick"][row] = ctx["shift_end_tick"] if ctx["shift_end_tick"] is not None else 0
        if ctx["last_present_tick"] is not None:
            flags |= FLAG_LAST_PRESENT
            a["last_present_tick"][row] = ctx["last_present_tick"]
        a["ctx_flags"][row] = flags
        a["scheduled_ticks"][row] = ctx["scheduled_ticks"]
        a["on_time_ticks"][row] = ctx["on_time_ticks"]
        samples = list(ctx["attendance_samples"])
        a["attendance_count"][row] = len(samples)
        a["attendance_samples"][row, : len(samples)] = samples
        events = list(ctx["absence_events"])[-TL_MAX_ABSENCE_EVENTS:]
        a["absence_count"][row] = len(events)
        a["absence_events"][row, : len(events)] = events
        a["late_ticks_today"][row] = getattr(snapshot, "late_ticks_today", 0)
        a["absent_shifts_7d"][row] = getattr(snapshot, "absent_shifts_7d", 0)
        a["wages_earned"][row] = int(getattr(snapshot, "inventory", {}).get("wages_earned", 0))


class DeviceEmployment:
    """The employment context of a ``DeviceTownBatch`` resident on the device; ``step`` = ``tl_employment_step``."""

    def __init__(self, dev: Any, spec: EmploymentSpec, arrays: EmploymentArrays | None = None, town_origin: np.ndarray | None = None) -> None:
        import torch

        self.dev = dev
        self.spec = spec
        self.lib = capi.load_library()
        host = arrays or EmploymentArrays(dev.n_agents)
        if host.n_agents != dev.n_agents:
            raise ValueError("employment arrays and the batch disagree on the number of agents")
        self.tensors = {name: torch.from_numpy(np.ascontiguousarray(host.arrays[name]).view(np.int32 if host.arrays[name].dtype == np.uint32 else host.arrays[name].dtype)).to(dev.device)
                        for name, _, _ in STATE_FIELDS}
        origin = np.zeros((dev.n_towns, 2), dtype=np.int32) if town_origin is None else np.ascontiguousarray(town_origin, dtype=np.int32)
        self.town_origin = torch.from_numpy(origin).to(dev.device)
        self._params = spec.struct()
        self._state = host.struct({k: v.data_ptr() for k, v in self.tensors.items()}, self.town_origin.data_ptr())

    def step(self, n_ticks: int = 1, tick_offset: int = 1, events: bool = True):
        """``n_ticks`` steps of ``apply_job_state`` on the batch; returns the event words ``[n_ticks, N]`` (or None)."""
        import torch

        from .engine import _stream_ptr

        out = torch.zeros((n_ticks, self.dev.n_agents), dtype=torch.int32, device=self.dev.device) if events else None
        with torch.cuda.device(self.dev.device):
            status = self.lib.tl_employment_step(
                C.byref(self.dev._struct), C.byref(self._params), C.byref(self._state), out.data_ptr() if out is not None else None,
                n_ticks, tick_offset, _stream_ptr(self.dev.device),
            )
        capi.check(status)
        return out

    def download(self) -> EmploymentArrays:
        arrays = {}
        for name, dtype, _ in STATE_FIELDS:
            arrays[name] = self.tensors[name].cpu().numpy().view(dtype).copy()
        return EmploymentArrays(self.dev.n_agents, arrays)
