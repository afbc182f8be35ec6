"""CPU: the C-ABI library loads and exports every symbol include/townlet_b200.h declares."""

from __future__ import annotations

import ctypes as C
import re

import pytest

from townlet_b200 import capi


@pytest.fixture(scope="module")
def lib():
    capi.build_library()
    return capi.load_library()


def test_header_symbols_are_exported(lib) -> None:
    header = (capi.INCLUDE / "townlet_b200.h").read_text()
    declared = set(re.findall(r"\b(tl_[a-z0-9_]+)\s*\(", header)) - {"tl_stream_t"}
    assert declared == set(capi.EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None, name


def test_struct_layouts_match_the_header(lib) -> None:
    for idx, struct in enumerate((capi.TlTownBatch, capi.TlObsParams, capi.TlRewardParams, capi.TlObsOut, capi.TlRewardOut, capi.TlDynamicsParams)):
        assert lib.tl_abi_struct_size(idx) == C.sizeof(struct), struct.__name__
    assert lib.tl_abi_struct_size(99) == -1
    assert lib.tl_abi_version() == capi.TL_ABI_VERSION == int(
        re.search(r"#define TL_ABI_VERSION (\d+)", (capi.INCLUDE / "townlet_b200.h").read_text()).group(1)
    )


def test_argument_errors_are_reported_without_a_gpu(lib) -> None:
    """Validation happens before any CUDA call, so it is testable on the CPU box."""
    b = capi.TlTownBatch()
    b.n_towns, b.n_agents, b.n_entities, b.max_edges, b.embed_words, b.grid_w, b.grid_h = 1, 1, 1, 99, 1, 16, 16
    assert lib.tl_tick(C.byref(b), None, None, None, None, capi.TL_PHASE_DECAY, 1, None) == -1
    assert b"max_edges" in lib.tl_last_error()
    b.max_edges = 6
    assert lib.tl_tick(C.byref(b), None, None, None, None, capi.TL_PHASE_DECAY, 1, None) == -1
    assert b"reward params" in lib.tl_last_error()
    assert lib.tl_tick(None, None, None, None, None, capi.TL_PHASE_DECAY, 1, None) == -1
    assert lib.tl_decay_reward(C.byref(b), None, None, capi.TL_PHASE_OBS, None) == -1
    # the world-dynamics phases need TlDynamicsParams, and action words when actions are requested
    assert lib.tl_tick(C.byref(b), None, None, None, None, capi.TL_PHASE_SOCIAL_DECAY, 1, None) == -1
    assert b"tl_tick_world" in lib.tl_last_error()
    d = capi.TlDynamicsParams()
    assert lib.tl_tick_world(C.byref(b), None, None, C.byref(d), None, None, capi.TL_PHASE_ACTIONS, 1, None) == -1
    assert b"action words" in lib.tl_last_error()
    assert lib.tl_tick_world(C.byref(b), None, None, C.byref(d), None, None, 1 << 9, 1, None) == -1
    assert lib.tl_host_tick(None, C.byref(b), None, None, None, None, capi.TL_PHASE_SOCIAL_DECAY, 1) == -1
    assert b"tl_host_tick_world" in lib.tl_last_error()


def test_missing_library_fails_loudly(monkeypatch, tmp_path) -> None:
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setenv("TOWNLET_B200_LIB", str(tmp_path / "absent.so"))
    with pytest.raises(capi.TownletB200Error, match="no CPU fallback"):
        capi.load_library()
