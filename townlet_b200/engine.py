"""Device-resident town batch and launch wrappers around the C ABI.

PyTorch is plumbing here: it owns device memory and the CUDA stream.  The
arithmetic happens in ``libtownlet_b200.so`` (``csrc/townlet_b200.cu``); when
that library is missing every entry point raises (no CPU fallback).
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import capi
from .params import DynamicsSpec, ObsSpec, RewardSpec
from .soa import MUTABLE_FIELDS, TownBatch

_TORCH_DTYPES = {
    "f8": torch.float64,
    "f4": torch.float32,
    "i4": torch.int32,
    "u4": torch.int32,  # bit-identical storage; torch lacks arithmetic uint32
    "u1": torch.uint8,
    "u8": torch.int64,
}


def _stream_ptr(device: torch.device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


@dataclass
class TickOutputs:
    """Device tensors written by one launch (leading axis = ticks)."""

    map: torch.Tensor | None = None
    features: torch.Tensor | None = None
    summary: torch.Tensor | None = None
    reward: torch.Tensor | None = None
    comps: torch.Tensor | None = None
    terminated: torch.Tensor | None = None
    kpi: torch.Tensor | None = None
    # optional bfloat16 mirror of ``map`` for an in-loop policy: ``[ticks, N, width]`` (or ``[N, width]``, overwritten every
    # tick), rows zero-padded to a multiple of 8 elements; may be a column slice of a wider buffer
    # (``TlObsOut.map_bf16``, device launches only)
    map_bf16: torch.Tensor | None = None
    features_bf16: torch.Tensor | None = None  # the same mirror of ``features``


class TickPlan:
    """A planned launch (``tl_plan_create``): kernel form, grid and parameter block are resolved once, a launch is
    one ``cudaLaunchKernel``.  ``out`` is a ring of tick slots; ``launch(slot)`` writes ``n_ticks`` ticks from ``slot``."""

    def __init__(self, dev: "DeviceTownBatch", handle: int, out: TickOutputs, n_ticks: int, keep: tuple) -> None:
        self.dev = dev
        self.lib = dev.lib
        self.handle = handle
        self.out = out
        self.n_ticks = n_ticks
        self._keep = keep  # tensors the plan's parameter block points at
        first = next(t for t in (out.map, out.reward, out.features, out.kpi) if t is not None)
        self.ring_slots = int(first.shape[0])

    def launch(self, slot: int = 0) -> None:
        capi.check(self.lib.tl_plan_launch(self.handle, slot, _stream_ptr(self.dev.device)))

    def launch_many(self, n_launches: int, slot0: int = 0, timed: bool = False) -> None:
        """``n_launches`` back-to-back launches walking the output ring, issued by one C call."""
        capi.check(
            self.lib.tl_plan_launch_many(self.handle, slot0, n_launches, self.ring_slots, 1 if timed else 0, _stream_ptr(self.dev.device))
        )

    def launch_ms(self, n_launches: int) -> np.ndarray:
        """Per-launch durations (ms) of the last ``launch_many(..., timed=True)``; synchronise the stream first."""
        ms = np.zeros(n_launches, dtype=np.float32)
        capi.check(self.lib.tl_plan_launch_ms(self.handle, ms.ctypes.data_as(C.POINTER(C.c_float)), n_launches))
        return ms

    @property
    def grid(self) -> tuple[int, int, int]:
        g, t, s = C.c_int32(), C.c_int32(), C.c_int32()
        capi.check(self.lib.tl_plan_grid(self.handle, C.byref(g), C.byref(t), C.byref(s)))
        return g.value, t.value, s.value

    def close(self) -> None:
        if self.handle:
            self.lib.tl_plan_destroy(self.handle)
            self.handle = None

    def __del__(self) -> None:  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass


class DeviceTownBatch:
    """A ``TownBatch`` resident in HBM plus the look-up tables of one ``ObsSpec``."""

    def __init__(
        self,
        batch: TownBatch,
        obs: ObsSpec | None = None,
        rew: RewardSpec | None = None,
        device: torch.device | str = "cuda:0",
        dyn: DynamicsSpec | None = None,
    ) -> None:
        self.lib = capi.load_library()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise capi.TownletB200Error("townlet_b200 runs on CUDA devices only (no CPU fallback)")
        self.host = batch
        self.obs = obs
        self.rew = rew
        self.layout = batch.layout
        host_t = torch.from_numpy(batch.arena[: batch.layout.total_bytes])
        self.arena = host_t.to(self.device, non_blocking=False)
        self._struct = batch.struct(self.arena.data_ptr())
        self._luts: dict[str, torch.Tensor] = {}
        self._obs_struct: capi.TlObsParams | None = None
        self._rew_struct: capi.TlRewardParams | None = rew.struct() if rew is not None else None
        self.dyn = dyn
        self._dyn_struct: capi.TlDynamicsParams | None = dyn.struct() if dyn is not None else None
        if obs is not None:
            for name, table in obs.luts().items():
                self._luts[name] = torch.from_numpy(np.ascontiguousarray(table)).to(self.device)
            self._obs_struct = obs.struct({k: v.data_ptr() for k, v in self._luts.items()})

    # sizes ----------------------------------------------------------------
    @property
    def n_agents(self) -> int:
        return self.layout.n_agents

    @property
    def n_towns(self) -> int:
        return self.layout.n_towns

    def field(self, name: str) -> torch.Tensor:
        """Typed device view of one SoA field (shares the arena's memory)."""
        lay = self.layout
        dtype = lay.dtypes[name]
        shape = lay.shapes[name]
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        off = lay.offsets[name]
        return self.arena[off : off + nbytes].view(_TORCH_DTYPES[dtype.str[1:]]).view(*shape)

    def upload(self, batch: TownBatch | None = None) -> None:
        """Host → device copy of the whole arena (one copy)."""
        src = batch or self.host
        self.arena.copy_(torch.from_numpy(src.arena[: self.layout.total_bytes]), non_blocking=True)

    def download_state(self, full: bool | None = None) -> TownBatch:
        """Device → host copy of the mutable fields into ``self.host`` (one copy).

        With a ``DynamicsSpec`` (or ``full=True``) the world-dynamics phases may have rewritten positions, entity
        words and ledger rows as well, so the whole arena comes back.
        """
        if full is None:
            full = self.dyn is not None
        n = self.layout.total_bytes if full else self.layout.mutable_bytes
        self.host.arena[:n] = self.arena[:n].cpu().numpy()
        return self.host

    # outputs --------------------------------------------------------------
    def alloc_outputs(
        self,
        phases: int,
        n_ticks: int = 1,
        *,
        summary: bool = False,
        comps: bool = False,
        kpi: bool = True,
        reward: bool = True,
        terminated: bool = True,
        map_f32: bool = True,
    ) -> TickOutputs:
        """``map_f32=False`` leaves ``out.map`` unset: the caller supplies ``map_bf16`` rows and the map is delivered as
        bfloat16 only (``TlObsOut.map == NULL``)."""
        out = TickOutputs()
        n, t, dev = self.n_agents, self.n_towns, self.device
        if phases & capi.TL_PHASE_OBS:
            assert self.obs is not None
            if map_f32:
                out.map = torch.empty((n_ticks, n, *self.obs.map_shape), dtype=torch.float32, device=dev)
            out.features = torch.empty((n_ticks, n, self.obs.n_features), dtype=torch.float32, device=dev)
            if summary:
                out.summary = torch.empty((n_ticks, n, capi.TL_N_SUMMARY), dtype=torch.int32, device=dev)
        if phases & (capi.TL_PHASE_DECAY | capi.TL_PHASE_REWARD):
            if reward:
                out.reward = torch.zeros((n_ticks, n), dtype=torch.float32, device=dev)
            if comps:
                out.comps = torch.zeros((n_ticks, n, capi.TL_N_COMPS), dtype=torch.float64, device=dev)
            if terminated:
                out.terminated = torch.zeros((n_ticks, n), dtype=torch.uint8, device=dev)
            if kpi:
                out.kpi = torch.zeros((n_ticks, t, capi.TL_N_KPI), dtype=torch.float32, device=dev)
        return out

    @staticmethod
    def _out_structs(out: TickOutputs, tick0: int = 0) -> tuple[capi.TlObsOut, capi.TlRewardOut]:
        oo = capi.TlObsOut()
        ro = capi.TlRewardOut()

        def bind(struct, name, tensor):
            if tensor is None:
                return
            view = tensor[tick0:]
            setattr(struct, name, C.c_void_p(view.data_ptr()))
            setattr(struct, f"{name}_tick_stride", int(tensor.stride(0)))

        bind(oo, "map", out.map)
        bind(oo, "features", out.features)
        bind(oo, "summary", out.summary)
        for name in ("map_bf16", "features_bf16"):
            mirror = getattr(out, name)
            if mirror is None:
                continue
            # rows may be views into a wider buffer (both mirrors side by side in one combined row)
            if mirror.dtype != torch.bfloat16 or mirror.dim() not in (2, 3) or mirror.stride(-1) != 1:
                raise capi.TownletB200Error(f"{name} must be bfloat16 rows [N, width] or [ticks, N, width] with unit element stride")
            view = mirror[tick0:] if mirror.dim() == 3 else mirror
            setattr(oo, name, C.c_void_p(view.data_ptr()))
            setattr(oo, f"{name}_tick_stride", int(mirror.stride(0)) if mirror.dim() == 3 else 0)
            setattr(oo, f"{name}_row", int(mirror.stride(-2)))
            setattr(oo, f"{name}_width", int(mirror.shape[-1]))
        bind(ro, "reward", out.reward)
        bind(ro, "comps", out.comps)
        bind(ro, "terminated", out.terminated)
        bind(ro, "kpi", out.kpi)
        return oo, ro

    # launches -------------------------------------------------------------
    def _dyn_pointer(self, phases: int, n_ticks: int, actions: torch.Tensor | None):
        if not phases & (capi.TL_PHASE_ACTIONS | capi.TL_PHASE_SOCIAL_DECAY):
            return None
        if self._dyn_struct is None:
            raise capi.TownletB200Error("the world-dynamics phases need a DynamicsSpec")
        self._dyn_struct.actions = None
        self._dyn_struct.actions_tick_stride = 0
        if phases & capi.TL_PHASE_ACTIONS:
            if actions is None or actions.device != self.arena.device or actions.element_size() != 4:
                raise capi.TownletB200Error("TL_PHASE_ACTIONS needs a 32-bit action tensor on the batch's device")
            if not actions.is_contiguous() or actions.shape[-1] != self.n_agents:
                raise capi.TownletB200Error("actions must be contiguous [n_agents] or [n_ticks, n_agents]")
            if actions.dim() == 2 and actions.shape[0] != n_ticks:
                raise capi.TownletB200Error("actions has the wrong number of ticks")
            self._dyn_struct.actions = actions.data_ptr()
            self._dyn_struct.actions_tick_stride = self.n_agents if actions.dim() == 2 else 0
        return C.byref(self._dyn_struct)

    def plan(
        self,
        out: TickOutputs,
        phases: int = capi.TL_PHASE_ALL,
        n_ticks: int = 1,
        actions: torch.Tensor | None = None,
    ) -> TickPlan:
        """Resolve a launch of ``n_ticks`` ticks once (``tl_plan_create``); ``out`` is a ring of tick slots."""
        oo, ro = self._out_structs(out, 0)
        obs_p = C.byref(self._obs_struct) if self._obs_struct is not None else None
        rew_p = C.byref(self._rew_struct) if self._rew_struct is not None else None
        dyn_p = self._dyn_pointer(phases, n_ticks, actions)
        with torch.cuda.device(self.device):
            handle = self.lib.tl_plan_create(C.byref(self._struct), obs_p, rew_p, dyn_p, C.byref(oo), C.byref(ro), phases, n_ticks)
        if not handle:
            capi.check(-1)
        return TickPlan(self, handle, out, n_ticks, (actions, self.arena, tuple(self._luts.values())))

    def tick(
        self,
        out: TickOutputs,
        phases: int = capi.TL_PHASE_ALL,
        n_ticks: int = 1,
        tick0: int = 0,
        actions: torch.Tensor | None = None,
    ) -> None:
        """One launch advancing ``n_ticks`` ticks; outputs go to ``out[tick0 : tick0 + n_ticks]``.

        ``actions`` (``TL_PHASE_ACTIONS``): device action words, ``[n_agents]`` (applied every tick of the
        launch) or ``[n_ticks, n_agents]``, int32 / uint32 bit patterns of ``params.pack_actions``.
        """
        oo, ro = self._out_structs(out, tick0)
        obs_p = C.byref(self._obs_struct) if self._obs_struct is not None else None
        rew_p = C.byref(self._rew_struct) if self._rew_struct is not None else None
        dyn_p = self._dyn_pointer(phases, n_ticks, actions)
        with torch.cuda.device(self.device):
            status = self.lib.tl_tick_world(
                C.byref(self._struct), obs_p, rew_p, dyn_p, C.byref(oo), C.byref(ro), phases, n_ticks,
                _stream_ptr(self.device),
            )
        capi.check(status)

    def obs_build(self, out: TickOutputs) -> None:
        """``WorldObservationService.build_batch`` for every agent of the batch (service.py:201)."""
        oo, _ = self._out_structs(out)
        assert self._obs_struct is not None
        with torch.cuda.device(self.device):
            status = self.lib.tl_obs_build(C.byref(self._struct), C.byref(self._obs_struct), C.byref(oo), _stream_ptr(self.device))
        capi.check(status)

    def decay_reward(self, out: TickOutputs, phases: int = capi.TL_PHASE_DECAY | capi.TL_PHASE_REWARD) -> None:
        """Need decay (grid.py:1439) and / or ``RewardEngine.compute`` (engine.py:34) in place."""
        _, ro = self._out_structs(out)
        assert self._rew_struct is not None
        with torch.cuda.device(self.device):
            status = self.lib.tl_decay_reward(
                C.byref(self._struct), C.byref(self._rew_struct), C.byref(ro), phases, _stream_ptr(self.device)
            )
        capi.check(status)


class HostEngine:
    """Host-buffer path: every array (batch, look-up tables, outputs) is HOST memory.

    ``tick`` is one synchronous step (``tl_host_tick_world``: H2D + kernels + D2H + sync) — what the drop-in
    services use.  ``submit`` / ``wait`` are the pipelined form (``tl_host_submit`` / ``tl_host_wait``): up to
    ``capi.TL_HOST_SLOTS`` steps in flight, each slot keeping its batch resident on the device, the map crossing
    the bus as float32, uint8 or bits (``wire``) and arriving as the reference's float32 tensor (expanded by the
    context's host threads) or as the wire bytes themselves (``map_format="wire"``).  Buffers may be pinned
    (``pin=True`` allocates them with torch) so copies run at full PCIe rate and without staging.
    """

    def __init__(self, obs: ObsSpec | None, rew: RewardSpec | None, device: int = 0, threads: int | None = None) -> None:
        self.lib = capi.load_library()
        self.obs = obs
        self.rew = rew
        self.ctx = self.lib.tl_context_create(device)
        if not self.ctx:
            capi.check(-2)
        if threads is not None:
            capi.check(self.lib.tl_context_set_threads(self.ctx, int(threads)))
        self._luts = {k: np.ascontiguousarray(v) for k, v in obs.luts().items()} if obs is not None else {}
        self._obs_struct = obs.struct({k: v.ctypes.data for k, v in self._luts.items()}) if obs is not None else None
        self._rew_struct = rew.struct() if rew is not None else None
        self._pinned: list[torch.Tensor] = []
        self._inflight: dict[int, tuple] = {}  # slot -> ctypes objects and arrays the pending step points at

    def close(self) -> None:
        if self.ctx:
            self.lib.tl_context_destroy(self.ctx)
            self.ctx = None

    def __del__(self) -> None:  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    def pinned_array(self, shape: tuple[int, ...], dtype: np.dtype | str) -> np.ndarray:
        """numpy view of a pinned torch buffer (kept alive by the engine)."""
        np_dtype = np.dtype(dtype)
        nbytes = max(1, int(np.prod(shape, dtype=np.int64)) * np_dtype.itemsize)
        t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        self._pinned.append(t)
        return t.numpy()[: int(np.prod(shape, dtype=np.int64)) * np_dtype.itemsize].view(np_dtype).reshape(shape)

    def pinned_batch(self, batch: TownBatch) -> TownBatch:
        """A copy of ``batch`` whose arena is page-locked (crosses the bus as one direct copy)."""
        arena = self.pinned_array((batch.layout.total_bytes,), np.uint8)
        arena[:] = batch.arena[: batch.layout.total_bytes]
        out = TownBatch(batch.layout, batch.grid_w, batch.grid_h, arena)
        out.has_social_in, out.has_personality = batch.has_social_in, batch.has_personality
        return out

    def wire_row_bytes(self, wire: str) -> int:
        """Bytes per agent row of the map on the wire (``map_format="wire"`` buffers are ``[ticks, N, row]`` uint8)."""
        assert self._obs_struct is not None
        row = int(self.lib.tl_wire_row_bytes(C.byref(self._obs_struct), capi.TL_WIRE[wire]))
        capi.check(row)
        return row

    def alloc_outputs(
        self, batch: TownBatch, phases: int, n_ticks: int = 1, *, summary: bool = True, comps: bool = True, pin: bool = False,
        map_wire: str | None = None,
    ) -> dict[str, np.ndarray]:
        """``map_wire``: allocate ``map`` as the wire bytes ``[ticks, N, row]`` uint8 (``map_format="wire"``) instead of float32."""
        make = self.pinned_array if pin else (lambda shape, dtype: np.zeros(shape, dtype=dtype))
        n, t = batch.n_agents, batch.n_towns
        out: dict[str, np.ndarray] = {}
        if phases & capi.TL_PHASE_OBS:
            assert self.obs is not None
            if map_wire is not None and map_wire != "f32":
                out["map"] = make((n_ticks, n, self.wire_row_bytes(map_wire)), np.uint8)
            else:
                out["map"] = make((n_ticks, n, *self.obs.map_shape), np.float32)
            out["features"] = make((n_ticks, n, self.obs.n_features), np.float32)
            if summary:
                out["summary"] = make((n_ticks, n, capi.TL_N_SUMMARY), np.int32)
        if phases & (capi.TL_PHASE_DECAY | capi.TL_PHASE_REWARD):
            out["reward"] = make((n_ticks, n), np.float32)
            out["terminated"] = make((n_ticks, n), np.uint8)
            out["kpi"] = make((n_ticks, t, capi.TL_N_KPI), np.float32)
            if comps:
                out["comps"] = make((n_ticks, n, capi.TL_N_COMPS), np.float64)
        return out

    def _out_structs(self, out: dict[str, np.ndarray], map_bytes: bool) -> tuple[capi.TlObsOut, capi.TlRewardOut]:
        oo = capi.TlObsOut()
        ro = capi.TlRewardOut()
        for struct, names in ((oo, ("map", "features", "summary")), (ro, ("reward", "comps", "terminated", "kpi"))):
            for name in names:
                arr = out.get(name)
                if arr is None:
                    continue
                setattr(struct, name, C.c_void_p(arr.ctypes.data))
                per = 1 if (name == "map" and map_bytes) else arr.itemsize  # wire-format maps state their tick stride in bytes
                setattr(struct, f"{name}_tick_stride", int(arr.strides[0] // per))
        return oo, ro

    def _dyn_struct(self, batch: TownBatch, phases: int, n_ticks: int, dyn: DynamicsSpec | None, actions: np.ndarray | None):
        if not phases & (capi.TL_PHASE_ACTIONS | capi.TL_PHASE_SOCIAL_DECAY):
            return None, None
        if dyn is None:
            raise capi.TownletB200Error("the world-dynamics phases need a DynamicsSpec")
        ds = dyn.struct()
        if phases & capi.TL_PHASE_ACTIONS:
            if actions is None or actions.shape[-1] != batch.n_agents or actions.dtype.itemsize != 4:
                raise capi.TownletB200Error("TL_PHASE_ACTIONS needs 32-bit action words [n_agents] or [n_ticks, n_agents]")
            actions = np.ascontiguousarray(actions)
            if actions.ndim == 2 and actions.shape[0] != n_ticks:
                raise capi.TownletB200Error("actions has the wrong number of ticks")
            ds.actions = actions.ctypes.data
            ds.actions_tick_stride = batch.n_agents if actions.ndim == 2 else 0
        return ds, actions

    def tick(
        self,
        batch: TownBatch,
        out: dict[str, np.ndarray],
        phases: int,
        n_ticks: int = 1,
        dyn: DynamicsSpec | None = None,
        actions: np.ndarray | None = None,
    ) -> None:
        """Run on HOST buffers; ``batch`` state fields are updated in place.

        The world-dynamics phases take a ``DynamicsSpec`` and (``TL_PHASE_ACTIONS``) host action words
        ``[n_agents]`` or ``[n_ticks, n_agents]`` (uint32 / int32)."""
        oo, ro = self._out_structs(out, False)
        bs = batch.struct()
        ds, actions = self._dyn_struct(batch, phases, n_ticks, dyn, actions)
        status = self.lib.tl_host_tick_world(
            self.ctx,
            C.byref(bs),
            C.byref(self._obs_struct) if self._obs_struct is not None else None,
            C.byref(self._rew_struct) if self._rew_struct is not None else None,
            C.byref(ds) if ds is not None else None,
            C.byref(oo),
            C.byref(ro),
            phases,
            n_ticks,
        )
        capi.check(status)

    def make_step(
        self,
        batch: TownBatch,
        out: dict[str, np.ndarray],
        phases: int = capi.TL_PHASE_ALL,
        n_ticks: int = 1,
        *,
        wire: str = "f32",
        map_format: str = "f32",
        dirty: "int | tuple[str, ...] | None" = None,
        writeback: bool = True,
        dyn: DynamicsSpec | None = None,
        actions: np.ndarray | None = None,
    ) -> tuple:
        """Build the ``TlHostStep`` of one step once; ``submit(slot, step=...)`` then costs one C call.  The step keeps
        references to the arrays it points at."""
        if map_format not in ("f32", "wire"):
            raise ValueError("map_format is 'f32' or 'wire'")
        as_wire = map_format == "wire"
        oo, ro = self._out_structs(out, as_wire)
        bs = batch.struct()
        ds, actions = self._dyn_struct(batch, phases, n_ticks, dyn, actions)
        if dirty is None:
            mask = capi.TL_DIRTY_ALL
        elif isinstance(dirty, int):
            mask = dirty
        else:
            mask = 0
            for name in dirty:
                mask |= 1 << capi.FIELD_INDEX[name]
        st = capi.TlHostStep()
        st.batch = C.pointer(bs)
        st.obs = C.pointer(self._obs_struct) if self._obs_struct is not None else None
        st.rew = C.pointer(self._rew_struct) if self._rew_struct is not None else None
        st.dyn = C.pointer(ds) if ds is not None else None
        st.obs_out = C.pointer(oo)
        st.rew_out = C.pointer(ro)
        st.phases = phases
        st.n_ticks = n_ticks
        st.dirty_fields = mask
        st.wire = capi.TL_WIRE[wire]
        st.map_format = capi.TL_MAP_WIRE if as_wire else capi.TL_MAP_F32
        st.writeback = 1 if writeback else 0
        return (st, bs, ds, oo, ro, batch, out, actions)

    def submit(self, slot: int, batch: TownBatch | None = None, out: dict[str, np.ndarray] | None = None, *args, step: tuple | None = None,
               **kwargs) -> None:
        """Enqueue one step on ``slot`` (``tl_host_submit``) and return; ``wait(slot)`` completes it.

        Either the arguments of ``make_step`` (``batch, out, phases, n_ticks, wire=..., map_format=..., dirty=...,
        writeback=..., dyn=..., actions=...``) or a prebuilt ``step``.  ``dirty``: ``None`` uploads every array; a tuple
        of field names (or a bit mask, ``capi.FIELD_INDEX``) uploads only those and uses the slot's resident copy of the
        rest.  ``map_format="wire"`` expects ``out["map"]`` as the wire bytes (``alloc_outputs(map_wire=wire)``).  The
        arrays must stay alive until ``wait`` returns (the engine keeps references)."""
        if step is None:
            step = self.make_step(batch, out, *args, **kwargs)
        capi.check(self.lib.tl_host_submit(self.ctx, slot, C.byref(step[0])))
        self._inflight[slot] = step

    def wait(self, slot: int) -> None:
        """Complete the step pending on ``slot``: afterwards its outputs (and written-back state) are in the host arrays."""
        try:
            capi.check(self.lib.tl_host_wait(self.ctx, slot))
        finally:
            self._inflight.pop(slot, None)

    @property
    def last_h2d_bytes(self) -> int:
        return int(self.lib.tl_context_last_h2d_bytes(self.ctx))

    @property
    def last_d2h_bytes(self) -> int:
        return int(self.lib.tl_context_last_d2h_bytes(self.ctx))


def expand_wire_map(wire_map: np.ndarray, obs: ObsSpec, wire: str) -> np.ndarray:
    """numpy restatement of the host expander (tests, and callers that keep the wire bytes): ``[..., N, row]`` uint8
    wire rows -> the reference's float32 ``[..., N, C, W, W]`` map (``full``: path channels from the look-up table)."""
    c, w, _ = obs.map_shape
    w2 = w * w
    cells = (4 if obs.variant == "full" else c) * w2
    lead = wire_map.shape[:-1]
    if wire == "u8":
        flat = wire_map[..., :cells].astype(np.float32)
    elif wire == "bits":
        flat = np.unpackbits(wire_map, axis=-1, bitorder="little")[..., :cells].astype(np.float32)
    else:
        raise ValueError("wire is 'u8' or 'bits'")
    if obs.variant == "full":
        path = np.broadcast_to(np.asarray(obs.luts()["path_lut"], dtype=np.float32).reshape(2 * w2), (*lead, 2 * w2))
        flat = np.concatenate([flat, path], axis=-1)
    return flat.reshape(*lead, c, w, w)
