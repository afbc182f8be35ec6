"""ctypes wrapper of ``oracle/liboracle.so`` (townlet_oracle.c) — TEST INFRASTRUCTURE ONLY.

Parity status: pinned against reference-generated golden vectors
(tests/golden/, tests/test_oracle_golden.py).  Never imported by the product
package ``townlet_b200``; it only borrows the package's ctypes struct
definitions so that oracle and kernels consume the identical batch bytes.
"""

from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from townlet_b200 import capi
from townlet_b200.params import DynamicsSpec, ObsSpec, RewardSpec
from townlet_b200.soa import TownBatch

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle.so"
_lib = None


def build(force: bool = False) -> Path:
    src = HERE / "townlet_oracle.c"
    hdr = HERE.parent / "include" / "townlet_b200.h"
    if force or not LIB.exists() or LIB.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["make", "-C", str(HERE), "-B", "liboracle.so"], check=True, capture_output=True)
    return LIB


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB))
        _lib.tlo_tick_world.restype = C.c_int
        _lib.tlo_tick_world.argtypes = [
            C.POINTER(capi.TlTownBatch),
            C.POINTER(capi.TlObsParams),
            C.POINTER(capi.TlRewardParams),
            C.POINTER(capi.TlDynamicsParams),
            C.POINTER(capi.TlObsOut),
            C.POINTER(capi.TlRewardOut),
            C.c_uint32,
            C.c_int32,
            C.c_int32,
        ]
    return _lib


def oracle_gae(rewards: np.ndarray, values: np.ndarray, dones: np.ndarray, gamma: float, gae_lambda: float):
    """GAE on time-major float32 arrays ``[T, N]`` (values ``[T or T+1, N]``); returns (advantages, returns)."""
    lib = load()
    lib.tlo_gae.restype = C.c_int
    lib.tlo_gae.argtypes = [C.c_void_p] * 5 + [C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double]
    rewards = np.ascontiguousarray(rewards, dtype=np.float32)
    values = np.ascontiguousarray(values, dtype=np.float32)
    dones = np.ascontiguousarray(dones, dtype=np.float32)
    steps, agents = rewards.shape
    adv = np.zeros_like(rewards)
    ret = np.zeros_like(rewards)
    status = lib.tlo_gae(
        rewards.ctypes.data, values.ctypes.data, dones.ctypes.data, adv.ctypes.data, ret.ctypes.data,
        steps, agents, int(values.shape[0] == steps + 1), gamma, gae_lambda,
    )
    if status != 0:
        raise RuntimeError(f"oracle tlo_gae failed with {status}")
    return adv, ret


def oracle_tick(
    batch: TownBatch,
    obs: ObsSpec | None,
    rew: RewardSpec | None,
    phases: int,
    n_ticks: int = 1,
    n_threads: int = 1,
    want_comps: bool = True,
    dyn: DynamicsSpec | None = None,
    actions: np.ndarray | None = None,
) -> dict[str, np.ndarray]:
    """Run the oracle IN PLACE on ``batch`` (host arena); returns the output arrays.

    Outputs carry a leading tick axis of length ``n_ticks``.  ``dyn`` / ``actions`` (uint32 words,
    ``[n_agents]`` or ``[n_ticks, n_agents]``) feed the world-dynamics phases.
    """
    lib = load()
    n, t = batch.n_agents, batch.n_towns
    out: dict[str, np.ndarray] = {}
    bs = batch.struct()
    op = rp = None
    oo = ro = None
    if phases & capi.TL_PHASE_OBS or (phases & capi.TL_PHASE_ADVANCE and obs is not None):
        assert obs is not None
        op = obs.struct()  # the oracle ignores LUT pointers
    if phases & capi.TL_PHASE_OBS:
        assert obs is not None
        out["map"] = np.full((n_ticks, n, *obs.map_shape), np.nan, dtype=np.float32)
        out["features"] = np.full((n_ticks, n, obs.n_features), np.nan, dtype=np.float32)
        out["summary"] = np.zeros((n_ticks, n, capi.TL_N_SUMMARY), dtype=np.int32)
        oo = capi.TlObsOut()
        oo.map = out["map"].ctypes.data
        oo.features = out["features"].ctypes.data
        oo.summary = out["summary"].ctypes.data
        oo.map_tick_stride = n * obs.map_elems
        oo.features_tick_stride = n * obs.n_features
        oo.summary_tick_stride = n * capi.TL_N_SUMMARY
    if phases & (capi.TL_PHASE_DECAY | capi.TL_PHASE_REWARD):
        assert rew is not None
        rp = rew.struct()
        out["reward"] = np.zeros((n_ticks, n), dtype=np.float32)
        out["terminated"] = np.zeros((n_ticks, n), dtype=np.uint8)
        out["kpi"] = np.zeros((n_ticks, t, capi.TL_N_KPI), dtype=np.float32)
        ro = capi.TlRewardOut()
        ro.reward = out["reward"].ctypes.data
        ro.terminated = out["terminated"].ctypes.data
        ro.kpi = out["kpi"].ctypes.data
        ro.reward_tick_stride = n
        ro.terminated_tick_stride = n
        ro.kpi_tick_stride = t * capi.TL_N_KPI
        if want_comps:
            out["comps"] = np.zeros((n_ticks, n, capi.TL_N_COMPS), dtype=np.float64)
            ro.comps = out["comps"].ctypes.data
            ro.comps_tick_stride = n * capi.TL_N_COMPS
    dp = None
    if phases & (capi.TL_PHASE_ACTIONS | capi.TL_PHASE_SOCIAL_DECAY):
        assert dyn is not None
        dp = dyn.struct()
        if phases & capi.TL_PHASE_ACTIONS:
            assert actions is not None and actions.shape[-1] == n
            actions = np.ascontiguousarray(actions, dtype=np.uint32)
            dp.actions = actions.ctypes.data
            dp.actions_tick_stride = n if actions.ndim == 2 else 0
    status = lib.tlo_tick_world(
        C.byref(bs),
        C.byref(op) if op is not None else None,
        C.byref(rp) if rp is not None else None,
        C.byref(dp) if dp is not None else None,
        C.byref(oo) if oo is not None else None,
        C.byref(ro) if ro is not None else None,
        phases,
        n_ticks,
        n_threads,
    )
    if status != 0:
        raise RuntimeError(f"oracle tlo_tick failed with {status}")
    return out


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) on uint32
    arrays; returns the four output words.  Known-answer vectors of the paper's Random123 library pin it in
    tests/test_gae.py."""
    m0, m1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & np.uint64(0xFFFFFFFF) for c in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = m0 * c0, m1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & mask, (k1 + np.uint64(0xBB67AE85)) & mask
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def oracle_sample_actions(heads: np.ndarray, n_actions: int, seed: int, counter: int = 0, offset: int = 0):
    """CPU restatement of ``tl_sample_actions`` (include/townlet_b200.h): float32 log_softmax of the first ``n_actions``
    columns (what the reference computes per agent, policy/runner.py:478-488) and one Gumbel-max categorical sample
    per row from Philox4x32-10 uniforms.  Returns ``(actions, log_probs, log_softmax, uniforms)``."""
    x = np.asarray(heads, dtype=np.float32)[:, :n_actions]
    n = x.shape[0]
    top = x.max(axis=1, keepdims=True)
    total = np.zeros((n, 1), dtype=np.float32)
    for j in range(n_actions):  # the kernel's summation order
        total[:, 0] += np.exp(x[:, j] - top[:, 0], dtype=np.float32)
    logp = x - (top + np.log(total, dtype=np.float32))
    rows = np.arange(n, dtype=np.uint64)
    u = np.empty((n, n_actions), dtype=np.float32)
    for d in range((n_actions + 3) // 4):
        c = np.uint64(4 * (counter + offset) + d)
        words = philox4x32_10(np.full(n, c & np.uint64(0xFFFFFFFF)), np.full(n, c >> np.uint64(32)), rows & np.uint64(0xFFFFFFFF),
                              rows >> np.uint64(32), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        for lane in range(4):
            if 4 * d + lane < n_actions:
                u[:, 4 * d + lane] = ((words[lane] >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0**-23)
    score = logp - np.log(-np.log(u, dtype=np.float32), dtype=np.float32)
    actions = score.argmax(axis=1)
    return actions.astype(np.int64), logp[np.arange(n), actions], logp, u


def oracle_employment_step(batch: TownBatch, spec, arrays, n_ticks: int = 1, tick_offset: int = 0, town_origin: np.ndarray | None = None) -> np.ndarray:
    """``tlo_employment_step``: ``apply_job_state`` restated in C on the host arrays (``batch`` and ``arrays`` are updated in
    place); returns the event words ``[n_ticks, n_agents]``."""
    lib = load()
    lib.tlo_employment_step.restype = C.c_int
    lib.tlo_employment_step.argtypes = [C.POINTER(capi.TlTownBatch), C.POINTER(capi.TlEmploymentParams), C.POINTER(capi.TlEmploymentState),
                                        C.c_void_p, C.c_int32, C.c_int32]
    origin = np.zeros((batch.n_towns, 2), dtype=np.int32) if town_origin is None else np.ascontiguousarray(town_origin, dtype=np.int32)
    bs = batch.struct()
    params = spec.struct()
    state = arrays.struct({name: arrays.arrays[name].ctypes.data for name, _, _ in capi.EMPLOYMENT_STATE_FIELDS}, origin.ctypes.data)
    events = np.zeros((n_ticks, batch.n_agents), dtype=np.uint32)
    status = lib.tlo_employment_step(C.byref(bs), C.byref(params), C.byref(state), events.ctypes.data, n_ticks, tick_offset)
    if status != 0:
        raise RuntimeError(f"oracle tlo_employment_step failed with {status}")
    return events
