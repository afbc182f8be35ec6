"""Device-resident rollout capture: tick kernels + batched in-loop policy forward + GAE scan.

The step immediately downstream of the hot path.  The reference converts every
observation tensor to nested Python lists (world/dto/factory.py:231-235), re-stacks them per agent
(policy/replay.py:550-636), runs one batch-size-1 forward per agent per tick (policy/runner.py:444-488) and
computes GAE with a Python double loop (policy/backends/pytorch/ppo_utils.py:55-67).  Here the tick kernel
writes straight into ``[T, N, ...]`` rollout tensors, the policy runs once per tick over all agents of all
towns (plain torch modules -> cuBLAS; the network is not part of the hot path), and GAE is one scan kernel.

A step pairs an observation, the action sampled on it, and the reward / done of the tick that applied that action
(policy/trajectory_service.py:127-138).  With ``act_on_world`` the next tick applies the sampled moves on the
device; otherwise action application stays with the caller and the actions are only recorded.
"""

from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import torch
from torch import nn

from . import capi
from .engine import DeviceTownBatch, TickOutputs, _stream_ptr


class PolicyMLP(nn.Module):
    """Same architecture as ``ConflictAwarePolicyNetwork`` (src/townlet/policy/models.py:36-69):
    flattened-map and feature encoders (Linear + ReLU), a shared Linear + ReLU, policy and value heads."""

    def __init__(self, feature_dim: int, map_shape: tuple[int, int, int], action_dim: int, hidden_dim: int = 256) -> None:
        super().__init__()
        flat_map = map_shape[0] * map_shape[1] * map_shape[2]
        self.map_encoder = nn.Sequential(nn.Flatten(), nn.Linear(flat_map, hidden_dim), nn.ReLU())
        self.feature_encoder = nn.Sequential(nn.Linear(feature_dim, hidden_dim), nn.ReLU())
        self.shared = nn.Sequential(nn.Linear(hidden_dim * 2, hidden_dim), nn.ReLU())
        self.policy_head = nn.Linear(hidden_dim, action_dim)
        self.value_head = nn.Linear(hidden_dim, 1)

    def forward(self, map_tensor: torch.Tensor, features: torch.Tensor,
                inputs_bf16: torch.Tensor | None = None) -> tuple[torch.Tensor, torch.Tensor]:
        """``inputs_bf16``: the combined padded bfloat16 input rows (``padded_input_buffer``) when the tick kernel already
        wrote them (``TickOutputs.map_bf16`` / ``.features_bf16`` = ``input_views(buffer)``); the bfloat16 path then
        skips its cast passes."""
        if (
            map_tensor.is_cuda
            and map_tensor.dtype == torch.float32
            and not torch.is_grad_enabled()
            and torch.is_autocast_enabled()
            and torch.get_autocast_dtype("cuda") == torch.bfloat16
        ):
            return self._forward_bf16_padded(map_tensor, features, inputs_bf16)
        hidden = torch.cat([self.map_encoder(map_tensor), self.feature_encoder(features)], dim=-1)
        hidden = self.shared(hidden)
        return self.policy_head(hidden), self.value_head(hidden).squeeze(-1)

    def _pad_widths(self) -> tuple[int, int]:
        pad = lambda k: (k + 63) // 64 * 64
        return pad(self.map_encoder[1].in_features), pad(self.feature_encoder[0].in_features)

    def padded_input_buffer(self, n: int, device: torch.device) -> torch.Tensor:
        """``[n, pad64(C·W·W) + pad64(F)]`` bfloat16: one combined input row per agent, map columns then features."""
        return torch.zeros((n, sum(self._pad_widths())), dtype=torch.bfloat16, device=device)

    def input_views(self, buffer: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        """The map and feature columns of a combined input buffer (what the tick kernel fills)."""
        pm, _ = self._pad_widths()
        return buffer[:, :pm], buffer[:, pm:]

    def _forward_bf16_padded(self, map_tensor: torch.Tensor, features: torch.Tensor, inputs_bf16: torch.Tensor | None = None,
                             raw_heads: bool = False):
        """The same network for bfloat16-autocast inference.  The two input widths (4·11·11 = 484 and 81) are not
        multiples of 8, which pushes cuBLAS onto unaligned fall-back kernels; the inputs are cast (what autocast would
        do) into ONE zero-padded row per agent — map columns, then features — either by the tick kernel itself
        (``inputs_bf16``) or by ``tl_cast_pad_bf16_strided``.  Both encoders are then one product with a block-diagonal
        weight (bias + ReLU in the epilogue) whose output IS the concatenation the shared layer reads, the shared layer
        one product with the same epilogue, and the policy and value heads one product: three GEMMs, no other pass
        (measured against two encoder products + a split shared layer + ReLU: 87 -> 70 µs at 65 536 agents,
        scripts/policy_gemm_variants.py)."""
        n = map_tensor.shape[0]
        flat, feat = self.map_encoder[1].in_features, self.feature_encoder[0].in_features
        hidden = self.map_encoder[1].out_features
        actions = self.policy_head.out_features
        dev = map_tensor.device
        bf16 = torch.bfloat16
        pm, pf = self._pad_widths()
        if inputs_bf16 is None:
            if getattr(self, "_pad_key", None) != (n, dev):
                self._pad_in = self.padded_input_buffer(n, dev)
                self._pad_key = (n, dev)
            x = self._pad_in
            lib = capi.load_library()
            stream = _stream_ptr(dev)
            src_map, src_feat = map_tensor.reshape(n, flat), features.contiguous()
            with torch.cuda.device(dev):
                capi.check(lib.tl_cast_pad_bf16_strided(src_map.data_ptr(), x.data_ptr(), n, flat, pm, pm + pf, stream))
                capi.check(lib.tl_cast_pad_bf16_strided(src_feat.data_ptr(), x[:, pm:].data_ptr(), n, feat, pf, pm + pf, stream))
        else:
            x = inputs_bf16
            if x.shape != (n, pm + pf) or x.dtype != bf16 or not x.is_contiguous():
                raise ValueError("inputs_bf16 must be the contiguous buffer of padded_input_buffer(n, device)")
        versions = tuple(p._version for p in self.parameters())
        if getattr(self, "_pad_versions", None) != versions or self._w["enc"].device != dev:
            def padded_t(weight: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
                w = torch.zeros((rows, cols), dtype=bf16, device=dev)  # [in, out], zero rows / columns for the padding
                w[: weight.shape[1], : weight.shape[0]] = weight.detach().t().to(bf16)
                return w

            enc = torch.zeros((pm + pf, 2 * hidden), dtype=bf16, device=dev)  # block diagonal: map rows | feature rows
            enc[:flat, :hidden] = self.map_encoder[1].weight.detach().t().to(bf16)
            enc[pm : pm + feat, hidden:] = self.feature_encoder[0].weight.detach().t().to(bf16)
            heads = (actions + 1 + 15) // 16 * 16  # policy logits and the value in ONE product, padded to 16 columns
            head_w = torch.cat([self.policy_head.weight.detach(), self.value_head.weight.detach()], dim=0)
            head_b = torch.zeros(heads, dtype=bf16, device=dev)
            head_b[: actions + 1] = torch.cat([self.policy_head.bias.detach(), self.value_head.bias.detach()]).to(bf16)
            self._w = {
                "enc": enc,
                "b_enc": torch.cat([self.map_encoder[1].bias.detach(), self.feature_encoder[0].bias.detach()]).to(bf16),
                "shared": self.shared[0].weight.detach().t().contiguous().to(bf16),  # reads cat([map, feat]) as produced
                "b_shared": self.shared[0].bias.detach().to(bf16),
                "heads": padded_t(head_w, hidden, heads),
                "b_heads": head_b,
            }
            self._pad_versions = versions
        w = self._w
        # bias + ReLU in the GEMM epilogue (cuBLASLt) instead of a separate pass over the activations
        fused = getattr(torch, "_addmm_activation", None)  # cuBLASLt epilogue; plain addmm + relu_ where torch lacks it
        if fused is not None:
            shared = fused(w["b_shared"], fused(w["b_enc"], x, w["enc"]), w["shared"])
        else:
            shared = torch.addmm(w["b_shared"], torch.addmm(w["b_enc"], x, w["enc"]).relu_(), w["shared"]).relu_()
        out = torch.addmm(w["b_heads"], shared, w["heads"])
        if raw_heads:
            return out  # [n, pad16(actions + 1)] bfloat16: logits, value, zero columns (what tl_sample_actions reads)
        return out[:, :actions], out[:, actions]


class FusedPolicyWeights:
    """The weights of a ``PolicyMLP`` in the layout ``tl_policy_step`` reads: bfloat16 ``[out, in]`` matrices with the input
    dimension zero-padded (map 484 -> 512, features 81 -> 128), the policy and value heads stacked and padded to 16 rows,
    float32 biases.  Raises ``TownletB200Error`` when the network is not the 256-wide reference shape."""

    def __init__(self, policy: "PolicyMLP", device: torch.device) -> None:
        hidden = policy.map_encoder[1].out_features
        flat, feat = policy.map_encoder[1].in_features, policy.feature_encoder[0].in_features
        actions = policy.policy_head.out_features
        pm, pf = policy._pad_widths()
        if hidden != 256 or pm != 512 or pf != 128 or actions + 1 > 16 or policy.shared[0].in_features != 512:
            raise capi.TownletB200Error("tl_policy_step is built for the reference shape: hidden 256, <= 512 map inputs, <= 128 features")
        bf16 = torch.bfloat16

        def padded(weight: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
            w = torch.zeros((rows, cols), dtype=bf16, device=device)
            w[: weight.shape[0], : weight.shape[1]] = weight.detach().to(device=device, dtype=bf16)
            return w.contiguous()

        self.n_actions = actions
        self.w_map = padded(policy.map_encoder[1].weight, hidden, pm)
        self.w_feat = padded(policy.feature_encoder[0].weight, hidden, pf)
        self.w_shared = padded(policy.shared[0].weight, hidden, 2 * hidden)
        self.w_heads = padded(torch.cat([policy.policy_head.weight.detach(), policy.value_head.weight.detach()], dim=0), 16, hidden)
        # the biases are rounded to bfloat16 first, as the autocast path holds them, then kept in float32
        rounded = lambda b: b.detach().to(device=device, dtype=bf16).float()
        self.b_enc = torch.cat([rounded(policy.map_encoder[1].bias), rounded(policy.feature_encoder[0].bias)]).contiguous()
        self.b_shared = rounded(policy.shared[0].bias).contiguous()
        self.b_heads = torch.zeros(16, dtype=torch.float32, device=device)
        self.b_heads[: actions + 1] = torch.cat([rounded(policy.policy_head.bias), rounded(policy.value_head.bias)])
        self.struct = capi.TlPolicyWeights()
        for name in ("w_map", "w_feat", "w_shared", "w_heads", "b_enc", "b_shared", "b_heads"):
            setattr(self.struct, name, getattr(self, name).data_ptr())
        self.struct.hidden, self.struct.map_in, self.struct.feat_in, self.struct.n_actions = hidden, pm, pf, actions
        self.versions = tuple(p._version for p in policy.parameters())


def policy_step(inputs_bf16: torch.Tensor, weights: FusedPolicyWeights, seed: int, counter: torch.Tensor | None = None, offset: int = 0,
                out: tuple[torch.Tensor, torch.Tensor, torch.Tensor] | None = None, heads_out: torch.Tensor | None = None):
    """``tl_policy_step``: encoders -> shared layer -> heads -> log_softmax -> categorical sample for every row of the padded
    bfloat16 input buffer, in one tcgen05 kernel.  Returns ``(actions int64, log_probs float32, values float32)``."""
    if not inputs_bf16.is_cuda or inputs_bf16.dtype != torch.bfloat16 or inputs_bf16.dim() != 2 or inputs_bf16.stride(1) != 1:
        raise capi.TownletB200Error("policy_step needs a CUDA bfloat16 [rows, >= 640] buffer with unit column stride (no CPU fallback)")
    n = inputs_bf16.shape[0]
    if out is None:
        out = (torch.empty(n, dtype=torch.int64, device=inputs_bf16.device), torch.empty(n, dtype=torch.float32, device=inputs_bf16.device),
               torch.empty(n, dtype=torch.float32, device=inputs_bf16.device))
    actions, log_probs, values = out
    if heads_out is not None and (heads_out.dtype != torch.float32 or heads_out.shape != (n, 16) or not heads_out.is_contiguous()):
        raise ValueError("heads_out must be contiguous float32 [rows, 16]")
    lib = capi.load_library()
    with torch.cuda.device(inputs_bf16.device):
        status = lib.tl_policy_step(
            inputs_bf16.data_ptr(), n, int(inputs_bf16.stride(0)), C.byref(weights.struct), seed & (2**64 - 1),
            counter.data_ptr() if counter is not None else None, offset, actions.data_ptr(), log_probs.data_ptr(), values.data_ptr(),
            heads_out.data_ptr() if heads_out is not None else None, _stream_ptr(inputs_bf16.device),
        )
    capi.check(status)
    return actions, log_probs, values


def gae(rewards: torch.Tensor, values: torch.Tensor, dones: torch.Tensor, gamma: float, gae_lambda: float):
    """``compute_gae`` on time-major CUDA tensors ``[T, N]`` (``values`` ``[T, N]`` or ``[T + 1, N]``)."""
    if rewards.ndim != 2 or dones.ndim != 2:
        raise ValueError("Rewards and dones must be 2D tensors (timesteps, agents)")
    steps, agents = rewards.shape
    if values.ndim != 2 or values.shape[0] not in (steps, steps + 1) or values.shape[1] != agents:
        raise ValueError("values must align with rewards (same length) or provide a bootstrap row")
    if not rewards.is_cuda:
        raise capi.TownletB200Error("townlet_b200.rollout.gae runs on CUDA tensors only (no CPU fallback)")
    rewards, values, dones = (t.contiguous().float() for t in (rewards, values, dones))
    advantages = torch.empty_like(rewards)
    returns = torch.empty_like(rewards)
    lib = capi.load_library()
    with torch.cuda.device(rewards.device):
        status = lib.tl_gae(
            rewards.data_ptr(), values.data_ptr(), dones.data_ptr(), advantages.data_ptr(), returns.data_ptr(),
            steps, agents, int(values.shape[0] == steps + 1), float(gamma), float(gae_lambda), _stream_ptr(rewards.device),
        )
    capi.check(status)
    return advantages, returns


def sample_actions(heads: torch.Tensor, n_actions: int, seed: int, counter: torch.Tensor | None = None, offset: int = 0,
                   with_value: bool = False, out: tuple[torch.Tensor, ...] | None = None):
    """One categorical sample per row of ``heads[:, :n_actions]`` (float32 or bfloat16 logits) in ONE kernel
    (``tl_sample_actions``): float32 log_softmax, Gumbel-max on counter-based Philox uniforms, the action's log
    probability and, ``with_value``, column ``n_actions`` as the value.  The same distribution as
    ``Categorical(logits=...).sample()`` / ``log_softmax(...)[action]`` (policy/runner.py:478-488).  ``counter`` is an
    optional int64 device scalar added to ``offset`` at run time, so a replayed CUDA graph draws fresh numbers once
    the caller advances it.  ``out``: contiguous ``(actions int64, log_probs float32[, values float32])`` rows to fill."""
    if not heads.is_cuda:
        raise capi.TownletB200Error("townlet_b200.rollout.sample_actions runs on CUDA tensors only (no CPU fallback)")
    if heads.dim() != 2 or heads.stride(1) != 1 or heads.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("heads must be [rows, columns] float32 or bfloat16 with unit column stride")
    if counter is not None and (counter.dtype != torch.int64 or counter.device != heads.device or counter.numel() != 1):
        raise ValueError("counter must be one int64 on the device of heads")
    n = heads.shape[0]
    if out is not None:
        actions, log_probs = out[0], out[1]
        values = out[2] if with_value else None
        for t, dtype in ((actions, torch.int64), (log_probs, torch.float32), (values, torch.float32)):
            if t is not None and (t.dtype != dtype or t.shape != (n,) or not t.is_contiguous() or t.device != heads.device):
                raise ValueError("out rows must be contiguous [rows] int64 / float32 / float32 on the device of heads")
    else:
        actions = torch.empty(n, dtype=torch.int64, device=heads.device)
        log_probs = torch.empty(n, dtype=torch.float32, device=heads.device)
        values = torch.empty(n, dtype=torch.float32, device=heads.device) if with_value else None
    lib = capi.load_library()
    with torch.cuda.device(heads.device):
        status = lib.tl_sample_actions(
            heads.data_ptr(), capi.TL_DTYPE_BF16 if heads.dtype == torch.bfloat16 else capi.TL_DTYPE_F32, n,
            int(heads.stride(0)) if n > 1 else int(heads.shape[1]), n_actions, seed & (2**64 - 1),
            counter.data_ptr() if counter is not None else None, offset,
            actions.data_ptr(), log_probs.data_ptr(), values.data_ptr() if values is not None else None,
            _stream_ptr(heads.device),
        )
    capi.check(status)
    return (actions, log_probs, values) if with_value else (actions, log_probs)


@dataclass
class Rollout:
    """``[T, N, ...]`` tensors of one capture, all on the device.

    Row ``t`` pairs the observation the policy saw (``map[t]``, ``features[t]``), the action it chose on it
    (``actions[t]``, ``log_probs[t]``, ``values[t]``) and the reward / done flag of the tick that APPLIED that action
    (the tick after the observation) — the pairing of the reference's trajectory frames
    (policy/trajectory_service.py:127-138).  ``values[T]`` is the policy's value of the final observation.
    """

    map: torch.Tensor
    features: torch.Tensor
    rewards: torch.Tensor
    dones: torch.Tensor
    values: torch.Tensor  # [T + 1, N], last row = the policy's value of the observation after the last tick
    actions: torch.Tensor
    log_probs: torch.Tensor
    advantages: torch.Tensor
    returns: torch.Tensor
    kpi: torch.Tensor  # [T, towns, 8], of the ticks that produced `rewards`

    def replay_batch(self, agents: slice | None = None, bootstrap: str = "value") -> dict[str, "np.ndarray"]:
        """The arrays of the reference's ``ReplayBatch`` (policy/replay.py:161-243), one sample per agent.

        What ``frames_to_replay_sample`` (policy/replay.py:550-636) builds per agent from the envelope's nested
        lists and ``build_batch`` (replay.py:245-276) stacks — ``maps [B, T, C, W, W]``, ``features [B, T, F]``,
        ``actions [B, T]`` int64, ``old_log_probs`` / ``rewards [B, T]`` float32, ``dones [B, T]`` bool,
        ``value_preds [B, T + 1]`` — taken straight from the time-major device rollout (one transpose, one D2H).
        ``bootstrap="value"`` keeps the policy's value of the final observation in the last ``value_preds`` column;
        ``"repeat"`` repeats the last step's value as the reference does (replay.py:603-607).
        """
        import numpy as np

        sel = agents if agents is not None else slice(None)

        def agent_major(t: torch.Tensor) -> "np.ndarray":
            return t[:, sel].transpose(0, 1).contiguous().cpu().numpy()

        values = self.values.float()
        steps = self.rewards.shape[0]
        if bootstrap == "repeat" or values.shape[0] == steps:
            values = torch.cat([values[:steps], values[steps - 1 : steps]], dim=0)
        elif bootstrap != "value":
            raise ValueError("bootstrap must be 'value' or 'repeat'")
        return {
            # (a bfloat16 map — RolloutCapture(map_dtype="bf16") — is flat [T, N, C*W*W]; callers reshape with the spec's map_shape)
            "maps": agent_major(self.map.float()),
            "features": agent_major(self.features.float()),
            "actions": agent_major(self.actions.long()).astype(np.int64),
            "old_log_probs": agent_major(self.log_probs.float()),
            "value_preds": agent_major(values),
            "rewards": agent_major(self.rewards.float()),
            "dones": agent_major(self.dones) != 0,
        }


@dataclass
class GraphedRollout:
    """A captured rollout: ``replay()`` advances the batch by its ticks and refreshes ``rollout`` in place."""

    graph: torch.cuda.CUDAGraph
    rollout: Rollout
    library_launches: int = 0  # kernels of libtownlet_b200 inside one replay (counted while capturing)

    def replay(self) -> Rollout:
        self.graph.replay()
        return self.rollout


@dataclass
class _Carry:
    """What one capture hands to the next: the last observation and the action already chosen on it."""

    map: torch.Tensor
    features: torch.Tensor
    action: torch.Tensor
    log_prob: torch.Tensor
    value: torch.Tensor


class RolloutCapture:
    """Collect ``n_ticks`` steps of every agent of a device batch with the policy in the loop.

    Step ``t``: the policy reads observation ``t`` and samples action ``t``; the NEXT tick applies it (with
    ``act_on_world``) and produces reward ``t``, done ``t`` and observation ``t + 1``.  A capture of ``n_ticks``
    steps therefore holds ``n_ticks + 1`` observations (``slots_for``); the last one — and the action already
    sampled on it — opens the next capture, so consecutive captures form one unbroken trajectory and every tick
    and every policy step is run exactly once.  The very first capture runs one priming tick to get observation 0.
    """

    # What a sampled action index does when the policy drives the world (``act_on_world=True``): 0 waits, 1-4 move one
    # tile north / south / east / west (clamped to the grid), anything above is a kind the host services would have
    # to resolve and is recorded as an unsuccessful request.  The reference has no fixed action space (its policy
    # runtime numbers action payloads as they appear, policy/runner.py:254-258), so this table is this wrapper's.
    ACTION_KIND = (3, 1, 1, 1, 1)  # TL_ACT_WAIT, TL_ACT_MOVE x 4
    ACTION_DX = (0, 0, 0, 1, -1)
    ACTION_DY = (0, -1, 1, 0, 0)

    def __init__(self, dev: DeviceTownBatch, policy: nn.Module, gamma: float = 0.99, gae_lambda: float = 0.95,
                 autocast_dtype: torch.dtype | None = torch.bfloat16, act_on_world: bool = False, map_dtype: str = "f32") -> None:
        """``map_dtype="bf16"``: the rollout keeps the map observation ONCE, as the bfloat16 rows the policy reads (exact: the
        hybrid map's cells are 0 or 1), instead of a float32 tensor plus that mirror — 2 056 instead of 3 992 bytes written
        per agent-step.  Needs the fused policy step (bfloat16 ``PolicyMLP``); ``alloc_outputs`` then allocates no ``map``
        and ``Rollout.map`` is a bfloat16 ``[T, N, C·W·W]`` view (``replay_batch`` casts it back, exactly)."""
        assert dev.obs is not None and dev.rew is not None
        if map_dtype not in ("f32", "bf16"):
            raise ValueError("map_dtype must be 'f32' or 'bf16'")
        self.map_dtype = map_dtype
        self._ring: torch.Tensor | None = None  # map_dtype="bf16": [slots, N, padded row] bfloat16, one row set per tick slot
        self.dev = dev
        self.policy = policy.to(dev.device).eval()
        self.gamma, self.gae_lambda = gamma, gae_lambda
        self.autocast_dtype = autocast_dtype
        self.act_on_world = act_on_world
        self.fuse_map_cast = True  # False: the policy step casts the float32 observation itself (A/B runs, tests)
        self.fuse_sampling = True  # False: log_softmax / Gumbel-max / gather as separate torch kernels (A/B runs)
        self.fuse_policy = True  # False: three cuBLASLt products + tl_sample_actions instead of tl_policy_step (A/B runs)
        self._fused_weights: FusedPolicyWeights | None = None
        self.seed = int(torch.initial_seed())  # Philox key of tl_sample_actions
        self._draws = torch.zeros(1, dtype=torch.int64, device=dev.device)  # policy steps sampled so far (device counter)
        self._map_bf16: torch.Tensor | None = None
        self._feat_bf16: torch.Tensor | None = None
        self._inputs_bf16: torch.Tensor | None = None  # the combined rows the two mirrors are views of
        self._carry: _Carry | None = None
        if act_on_world:
            if dev.dyn is None:
                raise capi.TownletB200Error("act_on_world needs a device batch built with a DynamicsSpec")
            table = lambda values, fill: torch.tensor(list(values) + [fill] * 11, dtype=torch.int32, device=dev.device)[:16]
            self._kind = table(self.ACTION_KIND, 4)  # TL_ACT_REQUEST for the indices without a device meaning
            self._dx, self._dy = table(self.ACTION_DX, 0), table(self.ACTION_DY, 0)
            self._words = torch.zeros(dev.n_agents, dtype=torch.int32, device=dev.device)

    @staticmethod
    def slots_for(n_ticks: int) -> int:
        """Tick slots the output tensors of a capture of ``n_ticks`` steps need (one more observation than steps)."""
        return n_ticks + 1

    def alloc_outputs(self, n_ticks: int, **kwargs) -> TickOutputs:
        """Output tensors for captures of ``n_ticks`` steps (without a float32 map when ``map_dtype="bf16"``)."""
        return self.dev.alloc_outputs(capi.TL_PHASE_ALL, self.slots_for(n_ticks), map_f32=self.map_dtype == "f32", **kwargs)

    def _rows_ring(self, slots: int) -> torch.Tensor:
        """``map_dtype="bf16"``: the combined bfloat16 input rows of every tick slot (the rollout's map observation)."""
        if self._mirror() is None or self._fused_policy() is None:
            raise capi.TownletB200Error("map_dtype='bf16' needs the bfloat16 PolicyMLP and the fused policy step")
        if self._ring is None or self._ring.shape[0] < slots:
            self._ring = torch.zeros((slots, *self._inputs_bf16.shape), dtype=torch.bfloat16, device=self.dev.device)
        return self._ring

    def describe(self) -> dict:
        """What the policy step is made of (for the bench record)."""
        fused = isinstance(self.policy, PolicyMLP) and self.autocast_dtype == torch.bfloat16 and self.fuse_map_cast
        return {
            "network": "ConflictAwarePolicyNetwork shape (policy/models.py:36-69): map / feature encoders 256, shared 256, 9 logits + value",
            "arithmetic": "bfloat16 inputs and weights, float32 accumulation" if self.autocast_dtype == torch.bfloat16 else "float32",
            "inputs": "bfloat16 rows written by the tick kernel" if fused else "float32 observation cast by the policy step",
            "map_delivery": "bfloat16 rows only (exact for the 0/1 map), features float32 + bfloat16" if self.map_dtype == "bf16"
            else "float32 tensor + bfloat16 rows for the policy",
            "sampling": "tl_sample_actions (one kernel: log_softmax + Gumbel-max + gather)" if self.fuse_sampling else "torch ops",
            "kernel": "tl_policy_step: one tcgen05 kernel per tick (TMA operands, TMEM accumulators, activations in shared memory, "
            "sampling in the epilogue)" if (fused and self._fused_policy() is not None) else "three cuBLASLt products",
        }

    def _action_words(self, act: torch.Tensor) -> torch.Tensor:
        """Action words (include/townlet_b200.h: TL_ACT_PACK) of the sampled indices, from the agents' positions."""
        idx = act.clamp(max=15).to(torch.int64)
        kind, dx, dy = self._kind[idx], self._dx[idx], self._dy[idx]
        pos = self.dev.field("pos")
        x = (pos & 0xFFFF).clamp(max=self.dev.host.grid_w - 1)
        y = ((pos >> 16) & 0xFFFF).clamp(max=self.dev.host.grid_h - 1)
        nx = (x + dx).clamp(0, self.dev.host.grid_w - 1)
        ny = (y + dy).clamp(0, self.dev.host.grid_h - 1)
        is_move = (kind == 1).to(torch.int32)
        return kind | (is_move << 5) | (nx << 6) | (ny << 16) | (1 << 26)

    @torch.no_grad()
    def _forward(self, map_t: torch.Tensor, feat_t: torch.Tensor, map_bf16: torch.Tensor | None = None):
        if self.autocast_dtype is not None:
            with torch.autocast("cuda", dtype=self.autocast_dtype):
                if map_bf16 is not None:
                    logits, value = self.policy(map_t, feat_t, self._inputs_bf16)
                else:
                    logits, value = self.policy(map_t, feat_t)
        else:
            logits, value = self.policy(map_t, feat_t)
        return logits.float(), value.float()

    def _mirror(self) -> torch.Tensor | None:
        """The padded bfloat16 rows the tick kernel writes beside the float32 map when the policy is the bfloat16
        ``PolicyMLP`` (one buffer, overwritten every tick: the policy consumes it before the next tick runs)."""
        if not (isinstance(self.policy, PolicyMLP) and self.autocast_dtype == torch.bfloat16 and self.fuse_map_cast):
            return None
        if self._map_bf16 is None:
            self._inputs_bf16 = self.policy.padded_input_buffer(self.dev.n_agents, self.dev.device)
            self._map_bf16, self._feat_bf16 = self.policy.input_views(self._inputs_bf16)
        return self._map_bf16

    def _new_rows(self, n_ticks: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        n, device = self.dev.n_agents, self.dev.device
        return (torch.empty((n_ticks + 1, n), dtype=torch.int64, device=device),
                torch.empty((n_ticks + 1, n), dtype=torch.float32, device=device),
                torch.empty((n_ticks + 1, n), dtype=torch.float32, device=device))

    def _tick(self, out: TickOutputs, slot: int) -> None:
        """One tick into ``slot``; with ``act_on_world`` it first applies the action words chosen on the previous
        observation (moves), then decays needs and ledgers."""
        if self.act_on_world:
            self.dev.tick(out, capi.TL_PHASE_WORLD, 1, tick0=slot, actions=self._words)
        else:
            self.dev.tick(out, capi.TL_PHASE_ALL, 1, tick0=slot)

    def _policy_step(self, out: TickOutputs, slot: int, draw: int, rows, mirror) -> None:
        """The policy on the observation in ``slot`` (just written by a tick): action, log-probability and value go to
        row ``slot`` of ``rows``; with ``act_on_world`` the action words of the next tick are derived from the action."""
        actions, log_probs, values = rows
        # categorical sample by the Gumbel-max trick (argmax of log-probabilities minus log of Exp(1) noise): the
        # same distribution as Categorical(logits).sample(), and no host sync
        fused = self._fused_policy() if (self.fuse_sampling and mirror is not None) else None
        if fused is not None:
            # the whole step — encoders, shared layer, heads, log_softmax, sample — in one tcgen05 kernel reading the rows the
            # tick kernel just wrote
            rows_in = self._ring[slot] if self.map_dtype == "bf16" else self._inputs_bf16
            act = policy_step(rows_in, fused, self.seed, self._draws, draw, out=(actions[slot], log_probs[slot], values[slot]))[0]
        elif self.fuse_sampling and mirror is not None:
            # bfloat16 PolicyMLP: one kernel reads the fused heads product and writes action, log-prob and value
            # straight into the rollout rows
            heads = self.policy._forward_bf16_padded(out.map[slot], out.features[slot], self._inputs_bf16, raw_heads=True)
            act = sample_actions(heads, self.policy.policy_head.out_features, self.seed, self._draws, draw,
                                 with_value=True, out=(actions[slot], log_probs[slot], values[slot]))[0]
        else:
            logits, value = self._forward(out.map[slot], out.features[slot], mirror)
            if self.fuse_sampling and logits.shape[1] <= capi.TL_MAX_ACTIONS:
                act = sample_actions(logits.contiguous(), logits.shape[1], self.seed, self._draws, draw,
                                     out=(actions[slot], log_probs[slot]))[0]
            else:
                logp = torch.log_softmax(logits, dim=-1)
                act = (logp - torch.empty_like(logp).exponential_().log_()).argmax(dim=-1)
                actions[slot], log_probs[slot] = act, logp.gather(1, act.unsqueeze(1)).squeeze(1)
            values[slot] = value
        if self.act_on_world:
            self._words.copy_(self._action_words(act))

    def _fused_policy(self) -> "FusedPolicyWeights | None":
        """The weights for ``tl_policy_step`` when the policy is the 256-wide bfloat16 ``PolicyMLP`` (rebuilt when a parameter changes)."""
        if not (self.fuse_policy and isinstance(self.policy, PolicyMLP) and self.autocast_dtype == torch.bfloat16):
            return None
        versions = tuple(p._version for p in self.policy.parameters())
        if self._fused_weights is None or self._fused_weights.versions != versions:
            try:
                self._fused_weights = FusedPolicyWeights(self.policy, self.dev.device)
            except capi.TownletB200Error:
                self.fuse_policy = False  # another layer shape: the library products
                return None
        return self._fused_weights

    def _with_mirrors(self, out: TickOutputs):
        mirror = self._mirror()
        saved = (out.map_bf16, out.features_bf16)
        if self.map_dtype == "bf16":  # one set of rows per tick slot: they ARE the stored map observation
            ring = self._rows_ring(out.features.shape[0])
            pm = self._map_bf16.shape[1]
            out.map_bf16, out.features_bf16 = ring[:, :, :pm], ring[:, :, pm:]
            return mirror, saved
        out.map_bf16 = mirror
        out.features_bf16 = self._feat_bf16 if mirror is not None else None
        return mirror, saved

    def _obs_slot(self, out: TickOutputs, slot: int) -> torch.Tensor:
        """The stored map observation of a tick slot (float32 tensor, or the bfloat16 rows with ``map_dtype="bf16"``)."""
        return self._ring[slot] if self.map_dtype == "bf16" else out.map[slot]

    @torch.no_grad()
    def prime(self, out: TickOutputs, n_ticks: int, rows=None) -> None:
        """The opening observation of a trajectory: one tick into slot ``n_ticks`` (where every capture leaves its last
        observation) and the policy's first action on it.  The reward of this tick belongs to whatever acted before
        the trajectory and is not recorded."""
        rows = rows or self._new_rows(n_ticks)
        mirror, saved = self._with_mirrors(out)
        try:
            self._tick(out, n_ticks)
            self._policy_step(out, n_ticks, 0, rows, mirror)
        finally:
            out.map_bf16, out.features_bf16 = saved
        self._draws += 1
        self._carry = _Carry(self._obs_slot(out, n_ticks), out.features[n_ticks], rows[0][n_ticks], rows[1][n_ticks], rows[2][n_ticks])

    def build_graph(self, n_ticks: int, out: TickOutputs | None = None) -> "GraphedRollout":
        """Capture the whole ``n_ticks`` rollout (tick launches + policy steps + GAE) in ONE CUDA graph.

        A tick is a handful of short kernels (the tick kernel, the policy's GEMMs, the sampling kernel); launched
        eagerly their CPU launch cost dominates the tick.  Replaying the graph keeps the device busy; the state —
        the towns, the last observation, the action chosen on it, the draw counter — lives in device memory, so every
        replay continues the trajectory where the previous one stopped, and the first replay starts exactly where an
        eager ``capture`` on the same batch would.
        """
        out = out or self.alloc_outputs(n_ticks)
        rows = self._new_rows(n_ticks)
        lib = capi.load_library()
        stream = torch.cuda.Stream(self.dev.device)
        stream.wait_stream(torch.cuda.current_stream(self.dev.device))
        with torch.cuda.stream(stream):  # warm-up outside the capture: lazy cuBLAS / allocator initialisation
            arena = self.dev.arena.clone()
            draws = self._draws.clone()
            words = self._words.clone() if self.act_on_world else None
            carry = self._carry
            self.capture(min(2, n_ticks), out)
            # everything the warm-up advanced goes back: the graph starts from the caller's state
            self.dev.arena.copy_(arena)
            self._draws.copy_(draws)
            if words is not None:
                self._words.copy_(words)
            self._carry = carry
            if self._carry is None:
                self.prime(out, n_ticks, rows)  # observation 0 and its action, into the slot / rows the replays carry over
            else:  # continue an eager trajectory: its last observation moves into the graph's carry slot
                c = self._carry
                if self.map_dtype == "bf16":
                    self._rows_ring(out.features.shape[0])
                self._obs_slot(out, n_ticks).copy_(c.map), out.features[n_ticks].copy_(c.features)
                rows[0][n_ticks].copy_(c.action), rows[1][n_ticks].copy_(c.log_prob), rows[2][n_ticks].copy_(c.value)
                self._carry = _Carry(self._obs_slot(out, n_ticks), out.features[n_ticks], rows[0][n_ticks], rows[1][n_ticks], rows[2][n_ticks])
        torch.cuda.current_stream(self.dev.device).wait_stream(stream)
        torch.cuda.synchronize(self.dev.device)
        graph = torch.cuda.CUDAGraph()
        before = lib.tl_launch_count()
        with torch.cuda.graph(graph):
            rollout = self.capture(n_ticks, out, rows)
        return GraphedRollout(graph, rollout, int(lib.tl_launch_count() - before))

    @torch.no_grad()
    def capture(self, n_ticks: int, out: TickOutputs | None = None, rows=None) -> Rollout:
        """``n_ticks`` steps; ``out`` needs ``slots_for(n_ticks)`` tick slots (slot ``t`` = observation ``t``)."""
        dev = self.dev
        out = out or self.alloc_outputs(n_ticks)
        if out.features is None or out.features.shape[0] < n_ticks + 1 or (self.map_dtype == "f32" and out.map is None):
            raise capi.TownletB200Error(f"a capture of {n_ticks} steps needs outputs with {n_ticks + 1} tick slots (slots_for)")
        if self.map_dtype == "bf16":
            self._rows_ring(out.features.shape[0])
        rows = rows or self._new_rows(n_ticks)
        actions, log_probs, values = rows
        if self._carry is None:
            self.prime(out, n_ticks, rows)
        c = self._carry
        if c.map.data_ptr() != self._obs_slot(out, 0).data_ptr():  # observation 0 = the last observation of the previous capture
            self._obs_slot(out, 0).copy_(c.map), out.features[0].copy_(c.features)
        actions[0].copy_(c.action), log_probs[0].copy_(c.log_prob), values[0].copy_(c.value)
        mirror, saved = self._with_mirrors(out)
        try:
            for t in range(n_ticks):
                self._tick(out, t + 1)  # applies action t; reward / done of step t and observation t + 1 land in slot t + 1
                self._policy_step(out, t + 1, t, rows, mirror)
        finally:
            out.map_bf16, out.features_bf16 = saved
        self._draws += n_ticks  # on the device: the next capture (or graph replay) draws fresh numbers
        self._carry = _Carry(self._obs_slot(out, n_ticks), out.features[n_ticks], actions[n_ticks], log_probs[n_ticks], values[n_ticks])
        rewards = out.reward[1 : n_ticks + 1]
        dones = out.terminated[1 : n_ticks + 1].float()
        # values[n_ticks] is the policy's value of the final observation: the bootstrap row of GAE
        advantages, returns = gae(rewards, values, dones, self.gamma, self.gae_lambda)
        if self.map_dtype == "bf16":  # [T, N, C*W*W] bfloat16: the map columns of the stored rows
            maps = self._ring[:n_ticks, :, : math.prod(dev.obs.map_shape)]
        else:
            maps = out.map[:n_ticks]
        return Rollout(maps, out.features[:n_ticks], rewards, dones, values, actions[:n_ticks],
                       log_probs[:n_ticks], advantages, returns, out.kpi[1 : n_ticks + 1])
