"""ctypes mirror of ``include/townlet_b200.h`` and loader of the CUDA library.

The structures below are field-for-field copies of the C header; the layout
is checked against the compiled library in ``tests/test_capi_symbols.py``
(``tl_abi_struct_size``).  There is no CPU fallback: ``load_library`` raises
when the shared object is missing.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = Path(__file__).resolve().parent / "csrc"
LIB_PATH = Path(__file__).resolve().parent / "lib" / "libtownlet_b200.so"
INCLUDE = ROOT / "include"

TL_ABI_VERSION = 4
TL_MAX_ACTIONS = 16
TL_DTYPE_F32, TL_DTYPE_BF16 = 0, 1
TL_N_NEEDS = 3
TL_MAX_EDGES = 8
TL_MAX_SOCIAL_SLOTS = 8
TL_MAX_PROFILES = 8
TL_N_COMPS = 16
TL_N_KPI = 8
TL_N_SUMMARY = 8
TL_N_BASE_FEATURES = 33
TL_MAX_CTYPES = 4
TL_MAX_WINDOW = 21
TL_MAX_GRID = 1024

TL_VARIANT = {"hybrid": 0, "full": 1, "compact": 2}

TL_PHASE_DECAY = 1
TL_PHASE_REWARD = 2
TL_PHASE_OBS = 4
TL_PHASE_ADVANCE = 8
TL_PHASE_ALL = 15
TL_PHASE_ACTIONS = 16  # needs TlDynamicsParams (tl_tick_world)
TL_PHASE_SOCIAL_DECAY = 32
TL_PHASE_RESERVATIONS = 64  # reservation slots and TL_F_RESERVATION re-derived from TlTownBatch.ent_holder every tick
TL_PHASE_WORLD = TL_PHASE_ALL | TL_PHASE_ACTIONS | TL_PHASE_SOCIAL_DECAY

# action kinds of the action word (include/townlet_b200.h: TL_ACT_*)
TL_N_ACT_KINDS = 16
ACT_KINDS = ("", "move", "noop", "wait", "request", "start", "release", "chat", "blocked")

TL_KIND_AGENT = 0
TL_KIND_OBJECT = 1
TL_KIND_RESERVED = 2
TL_KIND_EMPTY = 3

TL_F_ON_SHIFT = 1 << 0
TL_F_SHIFT_SHIFT = 1
TL_F_LAST_SUCCESS = 1 << 4
TL_F_IN_QUEUE = 1 << 5
TL_F_RESERVATION = 1 << 6
TL_F_TERMINATED = 1 << 7
TL_F_REASON_SHIFT = 8
TL_F_REASON_MASK = 3 << 8
TL_F_CTX_RESET = 1 << 10
TL_F_PROFILE_SHIFT = 11
TL_F_PROFILE_MASK = 7 << 11
TL_F_FAINTED = 1 << 14
TL_TERM_BLOCK_NONE = -(2**31)

COMP_NAMES = (
    "total",
    "survival",
    "needs_penalty",
    "social_bonus",
    "social_penalty",
    "social_avoidance",
    "social",
    "wage",
    "punctuality",
    "terminal_penalty",
    "clip_adjustment",
    "homeostasis",
    "work",
    "guardrail_active",
    "clipped",
    "reserved",
)
KPI_NAMES = (
    "reward_sum",
    "reward_sumsq",
    "hunger_min",
    "hunger_mean",
    "starving",
    "terminated",
    "lateness_mean",
    "wallet_mean",
)
SUMMARY_NAMES = (
    "other_agents",
    "objects",
    "reserved",
    "nearest_d2",
    "filled_slots",
    "relation_source",
    "reserved0",
    "reserved1",
)

_p = C.c_void_p


class TlTownBatch(C.Structure):
    _fields_ = [
        ("n_towns", C.c_int32),
        ("n_agents", C.c_int32),
        ("n_entities", C.c_int32),
        ("max_edges", C.c_int32),
        ("embed_words", C.c_int32),
        ("grid_w", C.c_int32),
        ("grid_h", C.c_int32),
        ("reserved0", C.c_int32),
        ("town_agent_off", _p),
        ("town_ent_off", _p),
        ("town_tick", _p),
        ("ent", _p),
        ("pos", _p),
        ("needs", _p),
        ("wallet", _p),
        ("attendance", _p),
        ("wages_withheld", _p),
        ("wages_paid", _p),
        ("punctuality", _p),
        ("episode_total", _p),
        ("social_in", _p),
        ("term_block_tick", _p),
        ("lateness", _p),
        ("last_action_duration", _p),
        ("episode_tick", _p),
        ("slot", _p),
        ("flags", _p),
        ("last_action_hash", _p),
        ("personality", _p),
        ("tie_count", _p),
        ("tie_embed", _p),
        ("tie_trust", _p),
        ("tie_fam", _p),
        ("tie_riv", _p),
        ("riv_count", _p),
        ("riv_embed", _p),
        ("riv_score", _p),
        ("ent_holder", _p),
        ("arena_base", _p),
        ("arena_bytes", C.c_int64),
    ]


class TlObsParams(C.Structure):
    _fields_ = [
        ("variant", C.c_int32),
        ("window", C.c_int32),
        ("n_channels", C.c_int32),
        ("n_features", C.c_int32),
        ("n_ctypes", C.c_int32),
        ("normalize_counts", C.c_int32),
        ("ticks_per_day", C.c_int32),
        ("max_slots", C.c_int32),
        ("rivalry_max_edges", C.c_int32),
        ("top_friends", C.c_int32),
        ("top_rivals", C.c_int32),
        ("embed_dim", C.c_int32),
        ("include_aggregates", C.c_int32),
        ("landmark_base", C.c_int32),
        ("personality_base", C.c_int32),
        ("social_base", C.c_int32),
        ("path_hint_ltype", C.c_int32),
        ("reserved0", C.c_int32),
        ("avoid_threshold", C.c_double),
        ("time_lut", _p),
        ("progress_lut", _p),
        ("slot_lut", _p),
        ("agent_ratio_lut", _p),
        ("tile_ratio_lut", _p),
        ("nearest_lut", _p),
        ("path_lut", _p),
        ("embed_lut", _p),
    ]


class TlRewardParams(C.Structure):
    _fields_ = [
        ("decay_rate", C.c_double * TL_N_NEEDS),
        ("need_weight", C.c_double * TL_N_NEEDS),
        ("survival_tick", C.c_double),
        ("wage_rate", C.c_double),
        ("punctuality_bonus", C.c_double),
        ("faint_penalty", C.c_double),
        ("eviction_penalty", C.c_double),
        ("clip_per_tick", C.c_double),
        ("clip_per_episode", C.c_double),
        ("faint_threshold", C.c_double),
        ("starvation_threshold", C.c_double),
        ("need_mult", (C.c_double * TL_N_NEEDS) * TL_MAX_PROFILES),
        ("reward_bias", (C.c_double * 3) * TL_MAX_PROFILES),
        ("block_window", C.c_int32),
        ("decay_personality", C.c_int32),
        ("reward_personality", C.c_int32),
        ("fuse_faint", C.c_int32),
    ]


class TlObsOut(C.Structure):
    _fields_ = [
        ("map", _p),
        ("features", _p),
        ("summary", _p),
        ("map_tick_stride", C.c_int64),
        ("features_tick_stride", C.c_int64),
        ("summary_tick_stride", C.c_int64),
        ("map_bf16", _p),
        ("map_bf16_tick_stride", C.c_int64),
        ("map_bf16_row", C.c_int32),
        ("features_bf16_row", C.c_int32),
        ("features_bf16", _p),
        ("features_bf16_tick_stride", C.c_int64),
        ("map_bf16_width", C.c_int32),
        ("features_bf16_width", C.c_int32),
    ]


class TlRewardOut(C.Structure):
    _fields_ = [
        ("reward", _p),
        ("comps", _p),
        ("terminated", _p),
        ("kpi", _p),
        ("reward_tick_stride", C.c_int64),
        ("comps_tick_stride", C.c_int64),
        ("terminated_tick_stride", C.c_int64),
        ("kpi_tick_stride", C.c_int64),
    ]


class TlDynamicsParams(C.Structure):
    _fields_ = [
        ("trust_decay", C.c_double),
        ("familiarity_decay", C.c_double),
        ("tie_rivalry_decay", C.c_double),
        ("rivalry_decay_per_tick", C.c_double),
        ("rivalry_min", C.c_double),
        ("rivalry_max", C.c_double),
        ("rivalry_eviction_threshold", C.c_double),
        ("actions", _p),
        ("actions_tick_stride", C.c_int64),
        ("action_hash", C.c_float * TL_N_ACT_KINDS),
    ]


class TlHostStep(C.Structure):
    _fields_ = [
        ("batch", C.POINTER(TlTownBatch)),
        ("obs", C.POINTER(TlObsParams)),
        ("rew", C.POINTER(TlRewardParams)),
        ("dyn", C.POINTER(TlDynamicsParams)),
        ("obs_out", C.POINTER(TlObsOut)),
        ("rew_out", C.POINTER(TlRewardOut)),
        ("phases", C.c_uint32),
        ("n_ticks", C.c_int32),
        ("dirty_fields", C.c_uint64),
        ("wire", C.c_int32),
        ("map_format", C.c_int32),
        ("writeback", C.c_int32),
        ("reserved0", C.c_int32),
    ]


class TlPolicyWeights(C.Structure):
    _fields_ = [
        ("w_map", _p),
        ("w_feat", _p),
        ("w_shared", _p),
        ("w_heads", _p),
        ("b_enc", _p),
        ("b_shared", _p),
        ("b_heads", _p),
        ("hidden", C.c_int32),
        ("map_in", C.c_int32),
        ("feat_in", C.c_int32),
        ("n_actions", C.c_int32),
    ]


TL_MAX_JOBS = 16
TL_MAX_ATTENDANCE_WINDOW = 14
TL_MAX_ABSENCE_EVENTS = 16


class TlEmploymentParams(C.Structure):
    _fields_ = [
        ("n_jobs", C.c_int32),
        ("arrival_buffer", C.c_int32),
        ("ticks_per_day", C.c_int32),
        ("grace_ticks", C.c_int32),
        ("absent_cutoff", C.c_int32),
        ("absence_slack", C.c_int32),
        ("attendance_window", C.c_int32),
        ("reserved0", C.c_int32),
        ("late_tick_penalty", C.c_double),
        ("absence_penalty", C.c_double),
        ("job_start", C.c_int32 * TL_MAX_JOBS),
        ("job_end", C.c_int32 * TL_MAX_JOBS),
        ("job_has_location", C.c_int32 * TL_MAX_JOBS),
        ("job_x", C.c_int32 * TL_MAX_JOBS),
        ("job_y", C.c_int32 * TL_MAX_JOBS),
        ("job_wage", C.c_double * TL_MAX_JOBS),
        ("job_lateness_penalty", C.c_double * TL_MAX_JOBS),
    ]


EMPLOYMENT_STATE_FIELDS = (
    ("job", 'i4', ()),
    ("state", 'i4', ()),
    ("current_day", 'i4', ()),
    ("late_ticks", 'i4', ()),
    ("ctx_flags", 'u4', ()),
    ("shift_started_tick", 'i4', ()),
    ("shift_end_tick", 'i4', ()),
    ("last_present_tick", 'i4', ()),
    ("scheduled_ticks", 'i4', ()),
    ("on_time_ticks", 'i4', ()),
    ("attendance_samples", 'f8', (TL_MAX_ATTENDANCE_WINDOW,)),
    ("attendance_count", 'i4', ()),
    ("absence_events", 'i4', (TL_MAX_ABSENCE_EVENTS,)),
    ("absence_count", 'i4', ()),
    ("late_ticks_today", 'i4', ()),
    ("absent_shifts_7d", 'i4', ()),
    ("wages_earned", 'i4', ()),
)


class TlEmploymentState(C.Structure):
    _fields_ = [(name, C.c_void_p) for name, _, _ in EMPLOYMENT_STATE_FIELDS] + [("town_origin", C.c_void_p)]


TL_HOST_SLOTS = 4
TL_WIRE = {"f32": 0, "u8": 1, "bits": 2}
TL_MAP_F32, TL_MAP_WIRE = 0, 1
TL_DIRTY_ALL = 2**64 - 1
# index of every TlTownBatch array in TlHostStep.dirty_fields (declaration order of the struct)
FIELD_INDEX = {name: i for i, (name, _) in enumerate(TlTownBatch._fields_[8:38])}

EXPORTED_SYMBOLS = (
    "tl_last_error",
    "tl_abi_version",
    "tl_abi_struct_size",
    "tl_device_sm_count",
    "tl_obs_build",
    "tl_decay_reward",
    "tl_tick",
    "tl_tick_world",
    "tl_launch_count",
    "tl_plan_create",
    "tl_plan_destroy",
    "tl_plan_launch",
    "tl_plan_launch_many",
    "tl_plan_launch_ms",
    "tl_plan_grid",
    "tl_set_option",
    "tl_reset_options",
    "tl_gae",
    "tl_cast_pad_bf16",
    "tl_cast_pad_bf16_strided",
    "tl_sample_actions",
    "tl_policy_step",
    "tl_employment_step",
    "tl_context_create",
    "tl_context_destroy",
    "tl_context_set_threads",
    "tl_wire_row_bytes",
    "tl_host_submit",
    "tl_host_wait",
    "tl_host_tick",
    "tl_host_tick_world",
    "tl_context_last_h2d_bytes",
    "tl_context_last_d2h_bytes",
)


class TownletB200Error(RuntimeError):
    """Raised when a C-ABI entry point returns a negative status."""


def nvcc_command(out: Path = LIB_PATH) -> list[str]:
    return [
        "nvcc",
        "-gencode",
        "arch=compute_100a,code=sm_100a",
        "-O3",
        "-fmad=false",  # float64 reward guardrail flags must round exactly like the reference (no FMA contraction)
        "-lineinfo",
        "-std=c++17",
        "-Xcompiler",
        "-fPIC",
        "-shared",
        f"-I{INCLUDE}",
        "-o",
        str(out),
        str(CSRC / "townlet_b200.cu"),
    ]


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    src = CSRC / "townlet_b200.cu"
    deps = [src, INCLUDE / "townlet_b200.h", *sorted(CSRC.glob("*.cuh"))]
    if LIB_PATH.exists() and not force:
        if all(LIB_PATH.stat().st_mtime >= d.stat().st_mtime for d in deps):
            return LIB_PATH
    LIB_PATH.parent.mkdir(parents=True, exist_ok=True)
    cmd = nvcc_command()
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise TownletB200Error(f"nvcc failed:\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        print(proc.stderr)
    return LIB_PATH


_lib: C.CDLL | None = None


def load_library() -> C.CDLL:
    """Load ``libtownlet_b200.so``; raise loudly when it is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("TOWNLET_B200_LIB", LIB_PATH))
    if not path.exists():
        raise TownletB200Error(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(townlet_b200 has no CPU fallback)"
        )
    lib = C.CDLL(str(path))
    lib.tl_last_error.restype = C.c_char_p
    lib.tl_abi_version.restype = C.c_int
    lib.tl_abi_struct_size.restype = C.c_int
    lib.tl_abi_struct_size.argtypes = [C.c_int]
    lib.tl_device_sm_count.restype = C.c_int
    lib.tl_launch_count.restype = C.c_int64
    lib.tl_obs_build.restype = C.c_int
    lib.tl_obs_build.argtypes = [C.POINTER(TlTownBatch), C.POINTER(TlObsParams), C.POINTER(TlObsOut), _p]
    lib.tl_decay_reward.restype = C.c_int
    lib.tl_decay_reward.argtypes = [
        C.POINTER(TlTownBatch),
        C.POINTER(TlRewardParams),
        C.POINTER(TlRewardOut),
        C.c_uint32,
        _p,
    ]
    lib.tl_tick.restype = C.c_int
    lib.tl_tick.argtypes = [
        C.POINTER(TlTownBatch),
        C.POINTER(TlObsParams),
        C.POINTER(TlRewardParams),
        C.POINTER(TlObsOut),
        C.POINTER(TlRewardOut),
        C.c_uint32,
        C.c_int32,
        _p,
    ]
    lib.tl_tick_world.restype = C.c_int
    lib.tl_tick_world.argtypes = [
        C.POINTER(TlTownBatch),
        C.POINTER(TlObsParams),
        C.POINTER(TlRewardParams),
        C.POINTER(TlDynamicsParams),
        C.POINTER(TlObsOut),
        C.POINTER(TlRewardOut),
        C.c_uint32,
        C.c_int32,
        _p,
    ]
    lib.tl_plan_create.restype = _p
    lib.tl_plan_create.argtypes = [
        C.POINTER(TlTownBatch),
        C.POINTER(TlObsParams),
        C.POINTER(TlRewardParams),
        C.POINTER(TlDynamicsParams),
        C.POINTER(TlObsOut),
        C.POINTER(TlRewardOut),
        C.c_uint32,
        C.c_int32,
    ]
    lib.tl_plan_destroy.restype = None
    lib.tl_plan_destroy.argtypes = [_p]
    lib.tl_plan_launch.restype = C.c_int
    lib.tl_plan_launch.argtypes = [_p, C.c_int64, _p]
    lib.tl_plan_launch_many.restype = C.c_int
    lib.tl_plan_launch_many.argtypes = [_p, C.c_int64, C.c_int32, C.c_int64, C.c_int32, _p]
    lib.tl_plan_launch_ms.restype = C.c_int
    lib.tl_plan_launch_ms.argtypes = [_p, C.POINTER(C.c_float), C.c_int32]
    lib.tl_plan_grid.restype = C.c_int
    lib.tl_plan_grid.argtypes = [_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.tl_set_option.restype = C.c_int
    lib.tl_set_option.argtypes = [C.c_char_p, C.c_int32]
    lib.tl_reset_options.restype = None
    lib.tl_reset_options.argtypes = []
    lib.tl_gae.restype = C.c_int
    lib.tl_gae.argtypes = [_p, _p, _p, _p, _p, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, _p]
    lib.tl_cast_pad_bf16.restype = C.c_int
    lib.tl_cast_pad_bf16.argtypes = [_p, _p, C.c_int64, C.c_int32, C.c_int32, _p]
    lib.tl_cast_pad_bf16_strided.restype = C.c_int
    lib.tl_cast_pad_bf16_strided.argtypes = [_p, _p, C.c_int64, C.c_int32, C.c_int32, C.c_int64, _p]
    lib.tl_sample_actions.restype = C.c_int
    lib.tl_sample_actions.argtypes = [_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_uint64, _p, C.c_int64, _p, _p, _p, _p]
    lib.tl_policy_step.restype = C.c_int
    lib.tl_policy_step.argtypes = [_p, C.c_int64, C.c_int64, C.POINTER(TlPolicyWeights), C.c_uint64, _p, C.c_int64, _p, _p, _p, _p, _p]
    lib.tl_employment_step.restype = C.c_int
    lib.tl_employment_step.argtypes = [C.POINTER(TlTownBatch), C.POINTER(TlEmploymentParams), C.POINTER(TlEmploymentState), _p, C.c_int32, C.c_int32, _p]
    lib.tl_context_create.restype = _p
    lib.tl_context_create.argtypes = [C.c_int]
    lib.tl_context_destroy.restype = None
    lib.tl_context_destroy.argtypes = [_p]
    lib.tl_context_set_threads.restype = C.c_int
    lib.tl_context_set_threads.argtypes = [_p, C.c_int32]
    lib.tl_wire_row_bytes.restype = C.c_int64
    lib.tl_wire_row_bytes.argtypes = [C.POINTER(TlObsParams), C.c_int32]
    lib.tl_host_submit.restype = C.c_int
    lib.tl_host_submit.argtypes = [_p, C.c_int32, C.POINTER(TlHostStep)]
    lib.tl_host_wait.restype = C.c_int
    lib.tl_host_wait.argtypes = [_p, C.c_int32]
    lib.tl_host_tick.restype = C.c_int
    lib.tl_host_tick.argtypes = [
        _p,
        C.POINTER(TlTownBatch),
        C.POINTER(TlObsParams),
        C.POINTER(TlRewardParams),
        C.POINTER(TlObsOut),
        C.POINTER(TlRewardOut),
        C.c_uint32,
        C.c_int32,
    ]
    lib.tl_host_tick_world.restype = C.c_int
    lib.tl_host_tick_world.argtypes = [
        _p,
        C.POINTER(TlTownBatch),
        C.POINTER(TlObsParams),
        C.POINTER(TlRewardParams),
        C.POINTER(TlDynamicsParams),
        C.POINTER(TlObsOut),
        C.POINTER(TlRewardOut),
        C.c_uint32,
        C.c_int32,
    ]
    lib.tl_context_last_h2d_bytes.restype = C.c_int64
    lib.tl_context_last_h2d_bytes.argtypes = [_p]
    lib.tl_context_last_d2h_bytes.restype = C.c_int64
    lib.tl_context_last_d2h_bytes.argtypes = [_p]
    if lib.tl_abi_version() != TL_ABI_VERSION:
        raise TownletB200Error("libtownlet_b200.so ABI version mismatch")
    for idx, struct in enumerate((TlTownBatch, TlObsParams, TlRewardParams, TlObsOut, TlRewardOut, TlDynamicsParams, TlHostStep, TlPolicyWeights, TlEmploymentParams, TlEmploymentState)):
        if lib.tl_abi_struct_size(idx) != C.sizeof(struct):
            raise TownletB200Error(
                f"struct layout mismatch for {struct.__name__}: C {lib.tl_abi_struct_size(idx)} vs ctypes {C.sizeof(struct)}"
            )
    _lib = lib
    return lib


def set_option(name: str, value: int) -> None:
    """``tl_set_option``: kernel-selection switches for parity tests and A/B runs (see the header)."""
    check(load_library().tl_set_option(name.encode(), int(value)))


def reset_options() -> None:
    load_library().tl_reset_options()


def check(status: int) -> None:
    if status < 0:
        lib = load_library()
        msg = lib.tl_last_error()
        raise TownletB200Error(f"townlet_b200 error {status}: {msg.decode() if msg else '?'}")
