"""GPU: the pipelined host-buffer path (tl_host_submit / tl_host_wait) — wire formats, resident batches, slots.

Everything is compared with the reference-generated fixtures, the oracle, or the synchronous path, bit for bit."""

from __future__ import annotations

import ctypes as C

import numpy as np
import pytest
from fixture_io import fixture_paths, load_fixture
from helpers import assert_kpi_close, assert_obs_equal, assert_reward_close

from townlet_b200 import capi
from townlet_b200.params import ObsSpec, RewardSpec
from townlet_b200.synth import synth_batch

pytestmark = pytest.mark.gpu
FIXTURES = fixture_paths()


def _binary_map(obs: ObsSpec) -> bool:
    """Does every integer channel of this variant hold 0 / 1 only (what TL_WIRE_BITS can carry)?"""
    return obs.variant != "compact" or obs.normalize_counts


@pytest.mark.parametrize("wire", ["u8", "bits"])
@pytest.mark.parametrize("path", FIXTURES, ids=[p.stem for p in FIXTURES])
def test_narrow_wire_formats_deliver_the_reference_tensors(path, wire) -> None:
    """Every reference-generated fixture through the uint8 / bit wire: the float32 tensors that arrive are the reference's."""
    from townlet_b200.engine import HostEngine, expand_wire_map

    batch, obs, rew, meta, ref = load_fixture(path)
    if not meta["phases"] & capi.TL_PHASE_OBS:
        pytest.skip("no observation in this fixture")
    eng = HostEngine(obs, rew, threads=3)
    out = eng.alloc_outputs(batch, meta["phases"], meta["n_ticks"])
    counts_above_one = "map" in ref and not _binary_map(obs) and float(ref["map"].max()) > 1.0
    eng.submit(0, batch, out, meta["phases"], meta["n_ticks"], wire=wire)
    if wire == "bits" and counts_above_one:
        with pytest.raises(capi.TownletB200Error, match="wire format"):
            eng.wait(0)
        eng.close()
        return
    eng.wait(0)
    assert_obs_equal(out, ref, f"{path.stem}/{wire}")
    assert_reward_close(out, ref, f"{path.stem}/{wire}")
    assert_kpi_close(out, ref, batch, f"{path.stem}/{wire}")
    # the same step delivering the wire bytes themselves; the numpy expander restates the host threads' work
    raw = eng.alloc_outputs(load_fixture(path)[0], meta["phases"], meta["n_ticks"], map_wire=wire)
    fresh = load_fixture(path)[0]
    eng.submit(1, fresh, raw, meta["phases"], meta["n_ticks"], wire=wire, map_format="wire")
    eng.wait(1)
    assert raw["map"].dtype == np.uint8 and raw["map"].shape[-1] == eng.wire_row_bytes(wire)
    assert np.array_equal(expand_wire_map(raw["map"], obs, wire).view(np.uint32), np.asarray(out["map"]).view(np.uint32))
    assert np.array_equal(raw["features"].view(np.uint32), out["features"].view(np.uint32))
    eng.close()


@pytest.mark.parametrize("pin", [False, True], ids=["pageable", "pinned"])
@pytest.mark.parametrize("variant,wire", [("hybrid", "f32"), ("hybrid", "u8"), ("hybrid", "bits"), ("full", "u8"), ("full", "bits"),
                                          ("compact", "u8"), ("compact", "bits")])
def test_pipelined_slots_equal_the_synchronous_path(variant, wire, pin) -> None:
    """Two batches alternate over two slots for several multi-tick steps; every step equals the oracle."""
    from oracle.oracle import oracle_tick
    from townlet_b200.engine import HostEngine

    obs, rew = ObsSpec(variant=variant), RewardSpec(fuse_faint=True)
    eng = HostEngine(obs, rew, threads=4)
    phases, n_ticks = capi.TL_PHASE_ALL, 2
    batches = [synth_batch(40, 48, 8, obs, first_town=100 * k) for k in range(2)]
    refs = [b.copy() for b in batches]
    if pin:
        batches = [eng.pinned_batch(b) for b in batches]
    outs = [eng.alloc_outputs(b, phases, n_ticks, pin=pin) for b in batches]
    eng.submit(0, batches[0], outs[0], phases, n_ticks, wire=wire)
    for step in range(6):
        cur, nxt = step % 2, (step + 1) % 2
        if step + 1 < 6:  # slot `nxt` was waited for in the previous iteration
            eng.submit(nxt, batches[nxt], outs[nxt], phases, n_ticks, wire=wire)
        eng.wait(cur)
        ref = oracle_tick(refs[cur], obs, rew, phases, n_ticks)
        assert_obs_equal(outs[cur], ref, f"step {step}")
        assert_reward_close(outs[cur], ref, f"step {step}")
        np.testing.assert_allclose(outs[cur]["kpi"], ref["kpi"], rtol=1e-5, atol=1e-6)
        assert np.array_equal(batches[cur].needs, refs[cur].needs) and np.array_equal(batches[cur].flags, refs[cur].flags)
        assert np.array_equal(batches[cur].town_tick, refs[cur].town_tick)
    assert eng.last_d2h_bytes > 0 and eng.last_h2d_bytes >= batches[0].layout.total_bytes
    eng.close()


def test_resident_batch_uploads_only_the_dirty_arrays() -> None:
    """The slot keeps the batch on the device: a step with an empty dirty set continues from the device state, a step
    that marks one array dirty sees the host's new values for it and the device's state for the rest."""
    from oracle.oracle import oracle_tick
    from townlet_b200.engine import HostEngine

    obs, rew = ObsSpec(), RewardSpec(fuse_faint=True)
    eng = HostEngine(obs, rew)
    batch = synth_batch(24, 48, 8, obs)
    ref_batch = batch.copy()
    out = eng.alloc_outputs(batch, capi.TL_PHASE_ALL, 1)
    eng.submit(2, batch, out, writeback=False)  # first step of the slot: everything goes up whatever the mask
    eng.wait(2)
    full_upload = eng.last_h2d_bytes
    ref = oracle_tick(ref_batch, obs, rew, capi.TL_PHASE_ALL, 1)
    assert_obs_equal(out, ref, "step 0")
    assert not np.array_equal(batch.needs, ref_batch.needs)  # no write-back: the host copy is stale by one tick

    eng.submit(2, batch, out, dirty=(), writeback=False)  # nothing changed on the host: the device state advances
    eng.wait(2)
    assert eng.last_h2d_bytes < full_upload // 4  # look-up tables only (~21 KB)
    ref = oracle_tick(ref_batch, obs, rew, capi.TL_PHASE_ALL, 1)
    assert_obs_equal(out, ref, "step 1 (resident)")
    assert_reward_close(out, ref, "step 1 (resident)")

    rng = np.random.default_rng(5)
    batch.social_in[...] = rng.uniform(-0.05, 0.05, batch.social_in.shape)  # the host reduces this tick's chat events
    batch.wallet[...] += 1.0
    ref_batch.social_in[...] = batch.social_in
    ref_batch.wallet[...] = batch.wallet
    eng.submit(2, batch, out, dirty=("social_in", "wallet"), writeback=True)
    eng.wait(2)
    assert eng.last_h2d_bytes < full_upload // 3
    ref = oracle_tick(ref_batch, obs, rew, capi.TL_PHASE_ALL, 1)
    assert_obs_equal(out, ref, "step 2 (two dirty arrays)")
    assert_reward_close(out, ref, "step 2 (two dirty arrays)")
    # write-back brings the three ticks of device state home
    assert np.array_equal(batch.needs, ref_batch.needs) and np.array_equal(batch.episode_total, ref_batch.episode_total)
    assert np.array_equal(batch.town_tick, ref_batch.town_tick)

    # a batch of another size in the same slot: the resident copy is dropped, everything is uploaded
    other = synth_batch(9, 48, 8, obs, first_town=500)
    other_ref = other.copy()
    out2 = eng.alloc_outputs(other, capi.TL_PHASE_ALL, 1)
    eng.submit(2, other, out2, dirty=())
    eng.wait(2)
    assert_obs_equal(out2, oracle_tick(other_ref, obs, rew, capi.TL_PHASE_ALL, 1), "other batch")
    eng.close()


def test_separately_allocated_arrays_are_copied_one_by_one_and_nothing_else_is_touched() -> None:
    """A foreign binding with one malloc per array (no arena declared): guard bytes around every array stay intact."""
    from oracle.oracle import oracle_tick
    from townlet_b200.engine import HostEngine
    from townlet_b200.soa import FIELD_NAMES

    obs, rew = ObsSpec(), RewardSpec(fuse_faint=True)
    batch = synth_batch(12, 48, 8, obs)
    ref_batch = batch.copy()
    eng = HostEngine(obs, rew)
    guard = 96
    holders = {}
    struct = batch.struct()
    for name in FIELD_NAMES:
        src = getattr(batch, name)
        raw = np.full(src.nbytes + 2 * guard, 0xA5, dtype=np.uint8)
        raw[guard : guard + src.nbytes] = src.reshape(-1).view(np.uint8)
        holders[name] = raw
        setattr(struct, name, C.c_void_p(raw.ctypes.data + guard))
    struct.arena_base = None
    struct.arena_bytes = 0
    out = eng.alloc_outputs(batch, capi.TL_PHASE_ALL, 2)
    oo, ro = eng._out_structs(out, False)
    capi.check(eng.lib.tl_host_tick(eng.ctx, C.byref(struct), C.byref(eng._obs_struct), C.byref(eng._rew_struct), C.byref(oo), C.byref(ro),
                                    capi.TL_PHASE_ALL, 2))
    ref = oracle_tick(ref_batch, obs, rew, capi.TL_PHASE_ALL, 2)
    assert_obs_equal(out, ref, "separate arrays")
    for name, raw in holders.items():
        assert (raw[:guard] == 0xA5).all() and (raw[-guard:] == 0xA5).all(), f"guard bytes around {name} were overwritten"
    needs = holders["needs"][guard:-guard].view(np.float64).reshape(3, -1)
    assert np.array_equal(needs, ref_batch.needs)  # the mutated arrays came back, each on its own
    # an array outside a declared arena is an argument error, not a guess
    struct.arena_base = C.c_void_p(holders["needs"].ctypes.data)
    struct.arena_bytes = 64
    status = eng.lib.tl_host_tick(eng.ctx, C.byref(struct), C.byref(eng._obs_struct), C.byref(eng._rew_struct), C.byref(oo), C.byref(ro),
                                  capi.TL_PHASE_ALL, 1)
    assert status == -1 and b"arena" in eng.lib.tl_last_error()
    eng.close()


def test_submit_argument_errors() -> None:
    from townlet_b200.engine import HostEngine

    obs, rew = ObsSpec(), RewardSpec(fuse_faint=True)
    eng = HostEngine(obs, rew)
    batch = synth_batch(4, 24, 4, obs)
    out = eng.alloc_outputs(batch, capi.TL_PHASE_ALL, 1)
    with pytest.raises(capi.TownletB200Error, match="slot"):
        eng.submit(capi.TL_HOST_SLOTS, batch, out)
    eng.submit(1, batch, out)
    with pytest.raises(capi.TownletB200Error, match="pending"):
        eng.submit(1, batch, out)
    eng.wait(1)
    eng.wait(1)  # nothing pending: a no-op
    with pytest.raises(KeyError):
        eng.submit(0, batch, out, wire="f16")
    assert eng.wire_row_bytes("u8") == 484 and eng.wire_row_bytes("bits") == 64 and eng.wire_row_bytes("f32") == 1936
    eng.close()


def test_world_phases_through_the_pipelined_path() -> None:
    from oracle.oracle import oracle_tick
    from townlet_b200.engine import HostEngine
    from townlet_b200.params import DynamicsSpec
    from townlet_b200.synth import random_actions

    obs, rew, dyn = ObsSpec(), RewardSpec(fuse_faint=True), DynamicsSpec()
    batch = synth_batch(20, 32, 6, obs)
    ref_batch = batch.copy()
    words = random_actions(batch, 3, seed=3)
    eng = HostEngine(obs, rew)
    out = eng.alloc_outputs(batch, capi.TL_PHASE_WORLD, 3)
    eng.submit(0, batch, out, capi.TL_PHASE_WORLD, 3, wire="bits", dyn=dyn, actions=words)
    eng.wait(0)
    ref = oracle_tick(ref_batch, obs, rew, capi.TL_PHASE_WORLD, 3, dyn=dyn, actions=words)
    assert_obs_equal(out, ref, "world")
    assert np.array_equal(batch.pos, ref_batch.pos) and np.array_equal(batch.ent, ref_batch.ent)
    assert np.array_equal(batch.riv_score, ref_batch.riv_score) and np.array_equal(batch.tie_count, ref_batch.tie_count)
    eng.close()
