"""CPU: layout, parameters, look-up tables, ingest and synthetic generator."""

from __future__ import annotations

import math

import numpy as np
import pytest
from fake_world import FakeWorld, fake_config

from townlet_b200 import capi
from townlet_b200.hashing import id_embedding_words, last_action_hash
from townlet_b200.ingest import ingest_town, rows_to_batch
from townlet_b200.params import ObsSpec, RewardSpec
from townlet_b200.soa import BatchLayout, TownBatch, pack_entity, pack_pos, unpack_pos
from townlet_b200.synth import assemble, synth_batch, synth_town, tile_batch


def test_feature_layout_matches_survey_appendix_b() -> None:
    spec = ObsSpec()
    names = spec.feature_names
    assert len(names) == 81 and spec.social_base == 33 and spec.social_vector_length == 48
    assert names[:3] == ("need_hunger", "need_hygiene", "need_energy") and names[16] == "ctx_reset_flag"
    assert names[29:33] == ("neighbor_agent_ratio", "neighbor_object_ratio", "reserved_tile_ratio", "nearest_agent_distance")
    assert names[33] == "social_slot0_d0" and names[-1] == "social_rivalry_max"
    targets = ObsSpec(include_targets=True, personality_channels=True)
    assert targets.n_features == 96 and targets.landmark_base == 33 and targets.personality_base == 45
    assert targets.landmark_slices["stove"] == (36, 39)
    assert ObsSpec(variant="full").map_shape == (6, 11, 11)
    assert ObsSpec(variant="compact", object_channels=("stove",)).map_channels == ("self", "agents", "objects", "reservations", "object:stove", "walkable")
    with pytest.raises(ValueError, match="relationships stage"):
        ObsSpec(relationships_stage="OFF")
    with pytest.raises(ValueError):
        ObsSpec(top_friends=5, top_rivals=4)


def test_luts_hold_the_reference_expressions() -> None:
    luts = ObsSpec().luts()
    assert luts["time_lut"].shape == (1440, 2) and luts["time_lut"][360, 0] == np.float32(math.sin(math.tau * 0.25))
    assert luts["progress_lut"][1440] == 1.0 and luts["progress_lut"][720] == np.float32(0.5)
    assert luts["agent_ratio_lut"][2] == np.float32(2 / 120) and luts["tile_ratio_lut"][3] == np.float32(3 / 121)
    assert luts["nearest_lut"][4] == np.float32(2 / 5) and luts["nearest_lut"][50] == 1.0
    assert luts["embed_lut"][255] == 1.0 and luts["embed_lut"][1] == np.float32(1) / np.float32(255)
    assert luts["path_lut"][0, 5 * 11 + 6] == 1.0 and luts["path_lut"][1, 6 * 11 + 5] == 1.0


@pytest.mark.parametrize("radius", range(1, 11))
def test_hypot_is_a_function_of_the_squared_distance(radius) -> None:
    """the float64 metadata distance can be derived from d^2 alone."""
    seen: dict[int, float] = {}
    for dy in range(-radius, radius + 1):
        for dx in range(-radius, radius + 1):
            value = float(np.hypot(dx, dy))
            assert seen.setdefault(dx * dx + dy * dy, value) == value


def test_sqrt_equals_hypot_after_float32_store() -> None:
    """The device uses sqrt of the exact integer d^2 for landmark distances (features.py:350-355)."""
    r = np.arange(-128, 129)
    dx, dy = np.meshgrid(r, r)
    d2 = (dx * dx + dy * dy).astype(np.float64)
    hyp = np.hypot(dx, dy)
    sq = np.sqrt(d2)
    assert np.array_equal(hyp.astype(np.float32), sq.astype(np.float32))
    nz = d2 > 0
    assert np.array_equal((dx[nz] / hyp[nz]).astype(np.float32), (dx[nz] / sq[nz]).astype(np.float32))


def test_hashing_known_answers() -> None:
    assert last_action_hash("") == 0.0
    assert np.float32(last_action_hash("wait")) == np.float32(0.23899886)
    words = id_embedding_words("bob", 4)
    import hashlib

    assert words.tobytes() == hashlib.blake2s(b"bob").digest()


def test_pack_roundtrips_and_limits() -> None:
    x, y = np.array([-5, 0, 300]), np.array([7, -32768, 32767])
    ux, uy = unpack_pos(pack_pos(x, y))
    assert np.array_equal(ux, x) and np.array_equal(uy, y)
    with pytest.raises(ValueError):
        pack_pos(np.array([40000]), np.array([0]))
    with pytest.raises(ValueError):
        pack_entity(np.array([1024]), np.array([0]), capi.TL_KIND_AGENT)
    w = int(pack_entity(np.array([3]), np.array([9]), capi.TL_KIND_OBJECT, 2, 1, 1, 1)[0])
    assert (w & 1023, (w >> 10) & 1023, (w >> 20) & 3, (w >> 22) & 7, (w >> 25) & 15, (w >> 29) & 1, (w >> 30) & 1) == (3, 9, 1, 2, 1, 1, 1)


def test_layout_is_aligned_and_mutable_fields_lead() -> None:
    lay = BatchLayout.make(7, 61, 200, 6, 1)
    assert all(off % 256 == 0 for off in lay.offsets.values())
    assert max(lay.offsets[n] for n in ("needs", "episode_total", "term_block_tick", "episode_tick", "flags", "town_tick")) < lay.mutable_bytes
    assert lay.offsets["ent"] >= lay.mutable_bytes
    b = TownBatch(lay, 48, 48)
    assert np.all(b.term_block_tick == capi.TL_TERM_BLOCK_NONE)
    s = b.struct()
    assert s.needs == b.arena.ctypes.data + lay.offsets["needs"]


def test_synthetic_batch_is_deterministic_and_valid() -> None:
    spec = ObsSpec(variant="compact", object_channels=("stove", "bed"))
    a, b = synth_batch(5, 48, 10, spec), synth_batch(5, 48, 10, spec)
    assert np.array_equal(a.arena, b.arena)
    a.validate(2)
    assert a.n_agents == 50 and a.n_entities >= 50 + 5 * 12
    kinds = (a.ent >> 20) & 3
    assert np.all(kinds[: 10] == capi.TL_KIND_AGENT)
    t = tile_batch(a, 3)
    t.validate(2)
    assert t.n_towns == 15 and np.array_equal(t.pos[50:100], a.pos)
    sub = t.slice_towns(5, 10)
    assert np.array_equal(sub.needs, a.needs) and np.array_equal(sub.ent, a.ent)
    bad = a.copy()
    bad.ent[0] = pack_entity(np.array([500]), np.array([0]), capi.TL_KIND_AGENT)[0]
    with pytest.raises(ValueError, match="outside the declared town grid"):
        bad.validate()


def test_ingest_of_a_fake_world_equals_the_assembler() -> None:
    spec = ObsSpec()
    town = synth_town(11, 48, 8, terminated_frac=0.3)
    world = FakeWorld(town, fake_config())
    terminated = {aid: bool(town.terminated[i]) for i, aid in enumerate(town.agent_ids)}
    reasons = {aid: "faint" for aid, flag in terminated.items() if flag}
    rows = ingest_town(world, spec, terminated=terminated, reasons=reasons, slots={a: i for i, a in enumerate(town.agent_ids)}, origin=(0, 0))
    got = rows_to_batch([rows], spec, grid=(48, 48))
    want = assemble([town], spec)
    for name, view in want.views().items():
        assert np.array_equal(getattr(got, name), view), name
    # without a pinned origin the town is shifted to its bounding box
    free = ingest_town(world, spec)
    assert free.origin == (int(min(town.positions[:, 0].min(), town.object_positions[:, 0].min())), int(min(town.positions[:, 1].min(), town.object_positions[:, 1].min())))


def test_reward_spec_struct_tables() -> None:
    p = RewardSpec(personality_scaling=True).struct()
    assert p.need_mult[2][1] == 1.1 and p.need_mult[3][2] == 0.9 and p.need_mult[4][0] == 0.95 and p.need_mult[1][0] == 1.0
    assert p.reward_bias[2][0] == 1.2 and p.reward_bias[3][1] == 1.25 and p.reward_bias[4][2] == 1.1
    assert p.decay_personality == 1 and p.block_window == 10 and p.faint_threshold == 0.03


# ---- property tests (hypothesis) --------------------------------------------------------------
from hypothesis import given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 1023), st.integers(0, 1023), st.integers(0, 2), st.integers(0, 4), st.integers(0, 4), st.booleans(), st.booleans())
def test_entity_word_fields_roundtrip(x, y, kind, ltype, ctype, lm, tile) -> None:
    from townlet_b200.soa import pack_entity

    w = int(pack_entity(x, y, kind, ltype, ctype, int(lm), int(tile)))
    assert (w & 1023, (w >> 10) & 1023, (w >> 20) & 3, (w >> 22) & 7, (w >> 25) & 15, (w >> 29) & 1, (w >> 30) & 1) == (
        x, y, kind, ltype, ctype, int(lm), int(tile))


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 15), st.booleans(), st.booleans(), st.integers(0, 1023), st.integers(0, 1023), st.integers(0, 63))
def test_action_word_fields_roundtrip(kind, success, has_pos, x, y, duration) -> None:
    from townlet_b200.params import pack_actions

    w = int(pack_actions(kind, int(success), int(has_pos), x, y, duration))
    assert (w & 15, (w >> 4) & 1, (w >> 5) & 1, (w >> 6) & 1023, (w >> 16) & 1023, w >> 26) == (
        kind, int(success), int(has_pos), x, y, duration)


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 9), st.integers(1, 7), st.integers(1, 5), st.data())
def test_town_slices_partition_any_batch(n_towns, agents, world_size, data) -> None:
    """Any split into contiguous blocks of towns keeps every field, and the blocks validate on their own."""
    from townlet_b200.sharding import shard_batch
    from townlet_b200.synth import assemble, synth_town

    towns = [synth_town(data.draw(st.integers(0, 500)), 16, data.draw(st.integers(1, agents))) for _ in range(n_towns)]
    full = assemble(towns, ObsSpec())
    parts = [shard_batch(full, r, world_size) for r in range(world_size)]
    assert sum(p.n_towns for p in parts) == full.n_towns
    for name in ("pos", "flags", "ent", "tie_count", "riv_score", "town_tick"):
        assert np.array_equal(np.concatenate([np.asarray(getattr(p, name)).reshape(-1) for p in parts]),
                              np.asarray(getattr(full, name)).reshape(-1)), name
    assert np.array_equal(np.concatenate([p.needs for p in parts], axis=1), full.needs)
    for p in parts:
        if p.n_towns:
            p.validate()


def test_bf16_mirror_views_bind_their_row_stride_and_width() -> None:
    """TickOutputs.map_bf16 / .features_bf16 may be column slices of one combined buffer: the struct carries the row
    stride, the written width and the tick stride separately (TlObsOut, ABI v3)."""
    import torch

    from townlet_b200 import capi
    from townlet_b200.engine import DeviceTownBatch, TickOutputs

    both = torch.zeros((3, 10, 648), dtype=torch.bfloat16)
    out = TickOutputs(map_bf16=both[..., :512], features_bf16=both[..., 512:640])
    oo, _ = DeviceTownBatch._out_structs(out, tick0=1)
    assert (oo.map_bf16_row, oo.map_bf16_width, oo.map_bf16_tick_stride) == (648, 512, 6480)
    assert (oo.features_bf16_row, oo.features_bf16_width, oo.features_bf16_tick_stride) == (648, 128, 6480)
    assert oo.map_bf16 == both[1].data_ptr() and oo.features_bf16 == both[1].data_ptr() + 2 * 512
    single = TickOutputs(map_bf16=torch.zeros((10, 512), dtype=torch.bfloat16))
    oo, _ = DeviceTownBatch._out_structs(single)
    assert (oo.map_bf16_row, oo.map_bf16_width, oo.map_bf16_tick_stride) == (512, 512, 0) and not oo.features_bf16
    for bad in (torch.zeros((10, 512), dtype=torch.float16), torch.zeros((10, 512, 2), dtype=torch.bfloat16)[..., 0],
                torch.zeros(512, dtype=torch.bfloat16)):
        with pytest.raises(capi.TownletB200Error, match="map_bf16"):
            DeviceTownBatch._out_structs(TickOutputs(map_bf16=bad))


def test_policy_helpers_refuse_cpu_tensors() -> None:
    """No CPU fallback: the sampling and GAE wrappers raise on host tensors instead of computing there."""
    import torch

    from townlet_b200 import capi
    from townlet_b200.rollout import PolicyMLP, gae, sample_actions

    with pytest.raises(capi.TownletB200Error, match="CUDA"):
        sample_actions(torch.zeros(4, 9), 9, seed=1)
    with pytest.raises(capi.TownletB200Error, match="CUDA"):
        gae(torch.zeros(2, 3), torch.zeros(2, 3), torch.zeros(2, 3), 0.99, 0.95)
    policy = PolicyMLP(81, (4, 11, 11), 9)
    buf = policy.padded_input_buffer(5, torch.device("cpu"))
    pm, pf = policy.input_views(buf)
    assert buf.shape == (5, 640) and pm.shape == (5, 512) and pf.shape == (5, 128) and pf.data_ptr() == buf.data_ptr() + 1024
