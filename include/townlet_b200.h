/*
 * townlet_b200.h — C ABI of the B200-native Townlet per-tick agent hot path.
 *
 * The reference (tachyon-beep/townlet) is pure Python and has no FFI of its
 * own; its seams for this path are Python protocols.  This header is the
 * boundary a maintainer binds (ctypes / cffi) underneath those protocols:
 *
 *   tl_obs_build      replaces WorldObservationService.build_batch
 *                     (src/townlet/world/observations/service.py:201-233) and
 *                     the encoders it drives (encoders/map.py:39,151,
 *                     encoders/features.py:43, encoders/social.py:27).
 *   tl_decay_reward   replaces WorldState._apply_need_decay
 *                     (src/townlet/world/grid.py:1439-1455), the faint
 *                     predicate of LifecycleManager.evaluate
 *                     (src/townlet/lifecycle/manager.py:39-45) and
 *                     RewardEngine.compute (src/townlet/rewards/engine.py:34-178).
 *   tl_tick           one launch that advances n_ticks ticks in the order of
 *                     SimulationLoop.step (src/townlet/core/sim_loop.py:632):
 *                     decay -> faint -> reward -> episode_tick -> observe.
 *   tl_host_tick      the same through HOST buffers (copies inside the call).
 *   TlTownBatch       the structure-of-arrays layout replacing AgentSnapshot
 *                     (world/agents/snapshot.py:19-44), InteractiveObject
 *                     (world/grid.py:121-128), the relationship / rivalry
 *                     ledgers (world/relationships.py:40, world/rivalry.py:36)
 *                     and the spatial index (world/spatial.py:14).
 *
 * Conventions: every pointer in TlTownBatch / TlObsOut / TlRewardOut and the
 * LUT pointers of TlObsParams is caller-owned DEVICE memory for the device
 * entry points (tl_obs_build, tl_decay_reward, tl_tick) and caller-owned HOST
 * memory for tl_host_tick.  Nothing is allocated inside the device entry
 * points; they are asynchronous on the given stream.  Return value 0 = ok,
 * negative = error (text from tl_last_error(), thread-local).  There is no
 * CPU fallback: without a CUDA device every compute entry returns TL_ERR_CUDA.
 */
#ifndef TOWNLET_B200_H
#define TOWNLET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TL_ABI_VERSION 4

#define TL_N_NEEDS 3          /* hunger, hygiene, energy (snapshot.py:15) */
#define TL_MAX_EDGES 8        /* upper bound for max_edges (ledger capacity) */
#define TL_MAX_SOCIAL_SLOTS 8 /* config/observations.py:43-44 */
#define TL_MAX_PROFILES 8
#define TL_N_COMPS 16         /* reward component row, see TL_COMP_* */
#define TL_N_KPI 8            /* per-town KPI row, see TL_KPI_* */
#define TL_N_SUMMARY 8        /* per-agent integer summary row, see TL_SUM_* */
#define TL_N_BASE_FEATURES 33 /* service.py:92-129 */
#define TL_MAX_CTYPES 4       /* compact object:<type> channels */
#define TL_MAX_WINDOW 21

/* error codes */
#define TL_OK 0
#define TL_ERR_ARG (-1)
#define TL_ERR_CUDA (-2)
#define TL_ERR_UNSUPPORTED (-3)

/* observation variants (config ObservationVariant) */
#define TL_VARIANT_HYBRID 0
#define TL_VARIANT_FULL 1
#define TL_VARIANT_COMPACT 2

/* phases of tl_tick */
#define TL_PHASE_DECAY 1u
#define TL_PHASE_REWARD 2u
#define TL_PHASE_OBS 4u
#define TL_PHASE_ADVANCE 8u /* tick += 1, episode_tick = (episode_tick+1) % tpd (world/core/context.py:283-289) */
/* phases upstream of the hot path; they need TlDynamicsParams (tl_tick_world) */
#define TL_PHASE_ACTIONS 16u      /* apply the tick's action words: moves + outcome bookkeeping
                                     (world/actions.py:44-105, world/systems/affordances.py:105-229) */
#define TL_PHASE_SOCIAL_DECAY 32u /* RelationshipService.decay (world/agents/relationships_service.py:165-191) */
#define TL_PHASE_RESERVATIONS 64u /* refresh_reservations (world/systems/queues.py:14-31,55-86; world/spatial.py:106-119):
                                     the reserved-tile entity words and the agents' TL_F_RESERVATION flag are re-derived
                                     every tick from TlTownBatch.ent_holder (who holds which object, as the host's queue
                                     manager granted it) instead of being ingested */

/*
 * Action word (uint32), one per agent and tick; what process_actions leaves on the agent snapshot.
 *   bits  0- 3  kind   TL_ACT_*; 0 = no action this tick (nothing is recorded)
 *   bit   4     success as resolved by the host services for kinds other than move
 *                      (queue grant, affordance start / release, chat); ignored for a move with a position
 *   bit   5     has_position (move only; a move without one is recorded as unsuccessful, affordances.py:157)
 *   bits  6-15  x      town-relative target tile of a move
 *   bits 16-25  y
 *   bits 26-31  duration 0..63 (action["duration"], default 1)
 */
#define TL_ACT_NONE 0u
#define TL_ACT_MOVE 1u
#define TL_ACT_NOOP 2u
#define TL_ACT_WAIT 3u
#define TL_ACT_REQUEST 4u
#define TL_ACT_START 5u
#define TL_ACT_RELEASE 6u
#define TL_ACT_CHAT 7u
#define TL_ACT_BLOCKED 8u
#define TL_N_ACT_KINDS 16
#define TL_ACT_KIND(a) ((a) & 15u)
#define TL_ACT_SUCCESS(a) (((a) >> 4) & 1u)
#define TL_ACT_HAS_POS(a) (((a) >> 5) & 1u)
#define TL_ACT_X(a) ((int)(((a) >> 6) & 1023u))
#define TL_ACT_Y(a) ((int)(((a) >> 16) & 1023u))
#define TL_ACT_DURATION(a) ((int)((a) >> 26))
#define TL_ACT_PACK(kind, success, has_pos, x, y, duration)                                        \
    (((uint32_t)(kind) & 15u) | (((uint32_t)(success) & 1u) << 4) | (((uint32_t)(has_pos) & 1u) << 5) | \
     (((uint32_t)(x) & 1023u) << 6) | (((uint32_t)(y) & 1023u) << 16) | (((uint32_t)(duration) & 63u) << 26))

/*
 * Entity word (uint32), one per thing that occupies a tile of a town.
 * Coordinates are relative to the town's origin chosen at ingest, so that
 * 0 <= x < grid_w and 0 <= y < grid_h.
 *   bits  0- 9  x
 *   bits 10-19  y
 *   bits 20-21  kind   0 agent occupancy (cache.py:28-31)
 *                      1 object (cache.py:33-38 / observe.py:109-147)
 *                      2 reserved tile (world/spatial.py:117-120)
 *                      3 empty slot: ignored by every reader (a reservation slot whose object is free)
 *   bits 22-24  ltype  landmark code, exact match on object_type:
 *                      0 other, 1 fridge, 2 stove, 3 bed, 4 shower
 *   bits 25-28  ctype  1 + index into compact.object_channels, 0 = none
 *   bit  29     lm     object is a nearest-of-type candidate
 *   bit  30     tile   object is present in the position lookup
 * Per town, entities are stored agents first (one per agent, agent order),
 * then objects in insertion order, then reserved tiles.  With TL_PHASE_RESERVATIONS the reserved tiles are
 * reservation SLOTS, one per object that has a position, in object order (kind 2 while reserved, 3 while free,
 * x / y = the object's position), followed by any reserved tile no object accounts for; ent_holder tells them apart.
 */
#define TL_ENT_X(e) ((int)((e) & 1023u))
#define TL_ENT_Y(e) ((int)(((e) >> 10) & 1023u))
#define TL_ENT_KIND(e) (((e) >> 20) & 3u)
#define TL_ENT_LTYPE(e) (((e) >> 22) & 7u)
#define TL_ENT_CTYPE(e) (((e) >> 25) & 15u)
#define TL_ENT_LM(e) (((e) >> 29) & 1u)
#define TL_ENT_TILE(e) (((e) >> 30) & 1u)
#define TL_ENT_PACK(x, y, kind, ltype, ctype, lm, tile)                              \
    (((uint32_t)(x) & 1023u) | (((uint32_t)(y) & 1023u) << 10) | ((uint32_t)(kind) << 20) | \
     ((uint32_t)(ltype) << 22) | ((uint32_t)(ctype) << 25) | ((uint32_t)(lm) << 29) |   \
     ((uint32_t)(tile) << 30))
#define TL_KIND_AGENT 0u
#define TL_KIND_OBJECT 1u
#define TL_KIND_RESERVED 2u
#define TL_KIND_EMPTY 3u
#define TL_MAX_GRID 1024

/* per-agent flag word */
#define TL_F_ON_SHIFT (1u << 0)
#define TL_F_SHIFT_SHIFT 1 /* bits 1-3: 0 pre_shift/unknown 1 on_time 2 late 3 absent 4 post_shift */
#define TL_F_SHIFT_MASK (7u << 1)
#define TL_F_LAST_SUCCESS (1u << 4)
#define TL_F_IN_QUEUE (1u << 5)
#define TL_F_RESERVATION (1u << 6)
#define TL_F_TERMINATED (1u << 7)
#define TL_F_REASON_SHIFT 8 /* bits 8-9: 0 none 1 faint 2 eviction 3 other */
#define TL_F_REASON_MASK (3u << 8)
#define TL_F_CTX_RESET (1u << 10) /* pending ctx-reset request: reported and CLEARED by the observation phase (grid.py:932-936) */
#define TL_F_PROFILE_SHIFT 11 /* bits 11-13: 0 none, 1 balanced, 2 socialite, 3 industrious, 4 stoic */
#define TL_F_PROFILE_MASK (7u << 11)
#define TL_F_FAINTED (1u << 14) /* written by the fused faint predicate each tick; effective terminated = TERMINATED | FAINTED */

#define TL_TERM_BLOCK_NONE INT32_MIN

/* reward component row (double[TL_N_COMPS]); names follow dto/rewards.py:17-60 */
enum {
    TL_COMP_TOTAL = 0,
    TL_COMP_SURVIVAL,
    TL_COMP_NEEDS_PENALTY,
    TL_COMP_SOCIAL_BONUS,
    TL_COMP_SOCIAL_PENALTY,
    TL_COMP_SOCIAL_AVOIDANCE,
    TL_COMP_SOCIAL,
    TL_COMP_WAGE,
    TL_COMP_PUNCTUALITY,
    TL_COMP_TERMINAL_PENALTY,
    TL_COMP_CLIP_ADJUSTMENT,
    TL_COMP_HOMEOSTASIS,
    TL_COMP_WORK,
    TL_COMP_GUARDRAIL_ACTIVE, /* 0.0 / 1.0 */
    TL_COMP_CLIPPED,          /* 0.0 / 1.0 */
    TL_COMP_RESERVED
};

/* per-town KPI row (float[TL_N_KPI]) */
enum {
    TL_KPI_REWARD_SUM = 0,
    TL_KPI_REWARD_SUMSQ,
    TL_KPI_HUNGER_MIN,
    TL_KPI_HUNGER_MEAN,
    TL_KPI_STARVING,   /* count(hunger <= starvation_threshold), starvation.py:132-139 */
    TL_KPI_TERMINATED, /* count */
    TL_KPI_LATENESS_MEAN, /* publisher.py:2671-2694 */
    TL_KPI_WALLET_MEAN
};

/* per-agent integer summary row (int32[TL_N_SUMMARY]); host derives the
 * float64 LocalSummary metadata from it (encoders/map.py:127-146) */
enum {
    TL_SUM_OTHER_AGENTS = 0,
    TL_SUM_OBJECTS,
    TL_SUM_RESERVED,
    TL_SUM_NEAREST_D2, /* 0 = none */
    TL_SUM_FILLED_SLOTS,
    TL_SUM_RELATION_SOURCE, /* 0 empty, 1 relationships, 2 rivalry_fallback */
    TL_SUM_RESERVED0,
    TL_SUM_RESERVED1
};

/* base feature indices, fixed by service.py:92-129 */
enum {
    TL_FEAT_NEED_HUNGER = 0,
    TL_FEAT_NEED_HYGIENE,
    TL_FEAT_NEED_ENERGY,
    TL_FEAT_WALLET,
    TL_FEAT_LATENESS,
    TL_FEAT_ON_SHIFT,
    TL_FEAT_TIME_SIN,
    TL_FEAT_TIME_COS,
    TL_FEAT_ATTENDANCE,
    TL_FEAT_WAGES_WITHHELD,
    TL_FEAT_SHIFT_PRE, /* ..+4 */
    TL_FEAT_SLOT_NORM = 15,
    TL_FEAT_CTX_RESET,
    TL_FEAT_EPISODE_PROGRESS,
    TL_FEAT_RIVALRY_MAX,
    TL_FEAT_RIVALRY_AVOID,
    TL_FEAT_LAST_HASH,
    TL_FEAT_LAST_SUCCESS,
    TL_FEAT_LAST_DURATION,
    TL_FEAT_RESERVATION,
    TL_FEAT_IN_QUEUE,
    TL_FEAT_PATH_NORTH, /* south, east, west follow */
    TL_FEAT_AGENT_RATIO = 29,
    TL_FEAT_OBJECT_RATIO,
    TL_FEAT_RESERVED_RATIO,
    TL_FEAT_NEAREST
};

/*
 * Structure-of-arrays town batch.  T towns, N agents (town-major: the agents
 * of town t are rows town_agent_off[t] .. town_agent_off[t+1]-1, in the
 * reference's agent iteration order), E entities.  K = max_edges is the row
 * stride of the tie and rivalry tables, EW = embed_words the number of
 * uint64 words per id embedding (ceil(embed_dim / 8)).
 */
typedef struct TlTownBatch {
    int32_t n_towns;
    int32_t n_agents;
    int32_t n_entities;
    int32_t max_edges;   /* K <= TL_MAX_EDGES */
    int32_t embed_words; /* EW in 1..4 */
    int32_t grid_w;      /* tiles; every town fits in grid_w x grid_h */
    int32_t grid_h;
    int32_t reserved0;

    /* towns */
    const int32_t* town_agent_off; /* [T+1] */
    const int32_t* town_ent_off;   /* [T+1] */
    int32_t* town_tick;            /* [T] world.tick */

    /* entities */
    uint32_t* ent; /* [E]; the agents' own words are rewritten by TL_PHASE_ACTIONS (spatial index, world/spatial.py:62-83) */

    /* agents: identity / geometry */
    int32_t* pos; /* [N] window centre snapshot.position: (x & 0xffff) | (y << 16), int16 each, town-relative */

    /* agents: float64 state (grid.py / snapshot.py fields stay float64) */
    double* needs;                /* [3][N] hunger, hygiene, energy */
    const double* wallet;         /* [N] */
    const double* attendance;     /* [N] attendance_ratio */
    const double* wages_withheld; /* [N] */
    const double* wages_paid;     /* [N] employment context wages_paid (context.py:57) */
    const double* punctuality;    /* [N] employment context punctuality (context.py:58) */
    double* episode_total;        /* [N] RewardEngine._episode_totals (engine.py:28) */
    const double* social_in;      /* [3][N] social_bonus, social_penalty, avoidance (engine.py:54-63); may be NULL */

    /* agents: integer state */
    int32_t* term_block_tick;            /* [N] RewardEngine._termination_block or TL_TERM_BLOCK_NONE */
    const int32_t* lateness;             /* [N] lateness_counter */
    int32_t* last_action_duration;       /* [N] */
    int32_t* episode_tick;               /* [N] */
    const int32_t* slot;                 /* [N] embedding slot (host allocator, embedding.py:35) */
    uint32_t* flags;                     /* [N] TL_F_* */
    float* last_action_hash;             /* [N] blake2s(last_action_id)[:4]/2^32 (features.py:194-200) */
    const float* personality;            /* [3][N] extroversion, forgiveness, ambition; may be NULL */

    /* relationship ties in ledger insertion order (social.py:171-186); rows are decayed and compacted in
     * place by TL_PHASE_SOCIAL_DECAY (entries at and beyond the count are zero afterwards) */
    uint8_t* tie_count;  /* [N] */
    uint64_t* tie_embed; /* [N][K][EW] blake2s(other_id) digest bytes, little-endian words */
    double* tie_trust;   /* [N][K] */
    double* tie_fam;     /* [N][K] */
    double* tie_riv;     /* [N][K] */

    /* rivalry edges as returned by rivalry_top(agent, K): already sorted (rivalry.py:101-110) */
    uint8_t* riv_count;  /* [N] */
    uint64_t* riv_embed; /* [N][K][EW] */
    double* riv_score;   /* [N][K] */

    /* ABI v4, TL_PHASE_RESERVATIONS: per entity, for a reservation slot the town-local index of the agent the queue
     * manager has granted the slot's object to (QueueManager.active_agent, world/queue/manager.py) or -1 while the
     * object is free; -2 for every other entity.  May be NULL when the phase is not requested. */
    const int32_t* ent_holder; /* [E] */

    /* ABI v4, host-buffer entry points only: when every array above lies inside ONE caller-owned block
     * [arena_base, arena_base + arena_bytes) the batch crosses the bus as a single copy (the layout of soa.py).
     * NULL / 0 = the arrays are separate allocations: one copy per array, and nothing outside an array is
     * ever read or written.  Ignored by the device entry points. */
    const void* arena_base;
    int64_t arena_bytes;
} TlTownBatch;

/* Observation parameters derived from SimulationConfig (config/observations.py). */
typedef struct TlObsParams {
    int32_t variant;          /* TL_VARIANT_* */
    int32_t window;           /* odd; hybrid.local_window or compact.map_window */
    int32_t n_channels;       /* 4 hybrid, 6 full, 5 + n_ctypes compact */
    int32_t n_features;       /* F */
    int32_t n_ctypes;         /* compact object:<type> channels, <= TL_MAX_CTYPES */
    int32_t normalize_counts; /* compact */
    int32_t ticks_per_day;    /* hybrid.time_ticks_per_day (also for compact/full) */
    int32_t max_slots;        /* embedding_allocator.max_slots */
    int32_t rivalry_max_edges;/* conflict.rivalry.max_edges */
    int32_t top_friends;
    int32_t top_rivals;
    int32_t embed_dim;
    int32_t include_aggregates;
    int32_t landmark_base;    /* feature index of fridge_dx, or -1 */
    int32_t personality_base; /* feature index of personality_extroversion, or -1 */
    int32_t social_base;      /* feature index of social_slot0_d0, or -1 when the snippet is disabled */
    int32_t path_hint_ltype;  /* 2 = stove (features.py:264) */
    int32_t reserved0;
    double avoid_threshold;   /* conflict.rivalry.avoid_threshold */
    /* float32 look-up tables computed on the host with the reference's own
     * float64 expressions, so device results are bit-identical by construction */
    const float* time_lut;        /* [tpd][2] sin, cos (features.py:142-148) */
    const float* progress_lut;    /* [tpd+1] clamp01(e/tpd) (features.py:180-188); index min(e, tpd) */
    const float* slot_lut;        /* [max_slots] slot/max_slots (features.py:172-176) */
    const float* agent_ratio_lut; /* [W*W] min(1, n/max(1,W*W-1)); index min(n, W*W-1) (map.py:127-128) */
    const float* tile_ratio_lut;  /* [W*W+1] min(1, n/W*W); index min(n, W*W) (map.py:129-130) */
    const float* nearest_lut;     /* [2*r*r+1] clamp01(hypot/max(1,r)) by squared distance (map.py:131-134) */
    const float* path_lut;        /* [2][W*W] dx/|d|, dy/|d| (map.py:115-125); full variant only */
    const float* embed_lut;       /* [256] float32(b) / 255.0f in float32 (social.py:236-242) */
} TlObsParams;

/* Reward / decay parameters derived from config.rewards (config/rewards.py). */
typedef struct TlRewardParams {
    double decay_rate[TL_N_NEEDS];  /* rewards.decay_rates; 0 when the key is absent */
    double need_weight[TL_N_NEEDS]; /* rewards.needs_weights */
    double survival_tick;
    double wage_rate;
    double punctuality_bonus;
    double faint_penalty;
    double eviction_penalty;
    double clip_per_tick;
    double clip_per_episode;
    double faint_threshold;      /* 0.03, lifecycle/manager.py:41 */
    double starvation_threshold; /* stability.starvation.hunger_threshold */
    double need_mult[TL_MAX_PROFILES][TL_N_NEEDS];   /* profile need multipliers (agents/models.py:61-85) */
    double reward_bias[TL_MAX_PROFILES][3];          /* social, employment, survival */
    int32_t block_window;        /* clip.no_positive_within_death_ticks */
    int32_t decay_personality;   /* WorldState._personality_reward_enabled (grid.py:244) */
    int32_t reward_personality;  /* RewardEngine._reward_scaling_enabled (engine.py:30) */
    int32_t fuse_faint;          /* rewrite TL_F_FAINTED = (hunger <= faint_threshold) after decay; reason faint unless one is set */
} TlRewardParams;

/* Observation outputs.  Row strides are in elements; per-tick strides apply in tl_tick. */
typedef struct TlObsOut {
    float* map;        /* [N][C][W][W]; may be NULL when map_bf16 is given and the kernel writes that row itself (entity
                          form, hybrid variant, window 11): the map is then delivered once, as bfloat16 — exact, its cells
                          are 0 or 1 — which takes 1 936 of the 3 992 bytes per agent-step off a rollout with the policy in
                          the loop; any other form answers TL_ERR_UNSUPPORTED */
    float* features;   /* [N][F] */
    int32_t* summary;  /* [N][TL_N_SUMMARY]; may be NULL */
    int64_t map_tick_stride;      /* elements between consecutive ticks (0 = overwrite) */
    int64_t features_tick_stride;
    int64_t summary_tick_stride;
    /* ABI v3, device entry points only (tl_host_tick ignores them): optional bfloat16 mirrors of `map` and `features`
     * for a policy that reads the observation right after the tick (the reference casts per agent
     * under autocast, policy/runner.py:444-488).  Row a of tick t = the flattened tile (feature vector) of agent a,
     * round-to-nearest-even, followed by zeros up to *_width elements; rows are *_row elements apart (row >= width, so
     * both mirrors can share one buffer of combined rows).  Widths, row and tick strides are multiples of 8 elements,
     * the bases 16-byte aligned; width 0 means width = row. */
    uint16_t* map_bf16;            /* [N][map_bf16_row]; may be NULL */
    int64_t map_bf16_tick_stride;  /* elements between consecutive ticks (0 = overwrite) */
    int32_t map_bf16_row;
    int32_t features_bf16_row;
    uint16_t* features_bf16;       /* [N][features_bf16_row]; may be NULL */
    int64_t features_bf16_tick_stride;
    int32_t map_bf16_width;        /* >= C*W*W */
    int32_t features_bf16_width;   /* >= F */
} TlObsOut;

typedef struct TlRewardOut {
    float* reward;        /* [N] total as float32; may be NULL */
    double* comps;        /* [N][TL_N_COMPS]; may be NULL */
    uint8_t* terminated;  /* [N] effective terminated flag of the tick; may be NULL */
    float* kpi;           /* [T][TL_N_KPI]; may be NULL */
    int64_t reward_tick_stride;
    int64_t comps_tick_stride;
    int64_t terminated_tick_stride;
    int64_t kpi_tick_stride;
} TlRewardOut;

/*
 * World dynamics around the hot path, in WorldContext.tick order
 * (world/core/context.py:229-289): actions are applied first, the ledgers decay in the
 * "relationships" system step after need decay, both before reward and observation.
 */
typedef struct TlDynamicsParams {
    /* RelationshipLedger.decay (world/relationships.py:78-89): a value moves towards 0 by its decay
     * (no change when the decay is <= 0); a tie whose three values are 0.0 afterwards is evicted */
    double trust_decay;
    double familiarity_decay;
    double tie_rivalry_decay;
    /* RivalryLedger.decay(ticks=1) (world/rivalry.py:68-83): clamp(v - decay_per_tick, min, max);
     * an edge at or below the eviction threshold is evicted */
    double rivalry_decay_per_tick;
    double rivalry_min;
    double rivalry_max;
    double rivalry_eviction_threshold;
    /* action words of the launch, [n_ticks][N] (tick stride in elements, 0 = the same row every tick) */
    const uint32_t* actions;
    int64_t actions_tick_stride;
    /* last_action_id feature per kind code: blake2s(kind)[:4] / 2^32 as float32 (features.py:194-200) */
    float action_hash[TL_N_ACT_KINDS];
} TlDynamicsParams;

typedef void* tl_stream_t; /* cudaStream_t */

const char* tl_last_error(void);
int tl_abi_version(void);
/* sizeof of the idx-th ABI struct (0 TlTownBatch, 1 TlObsParams, 2 TlRewardParams,
 * 3 TlObsOut, 4 TlRewardOut, 5 TlDynamicsParams, 6 TlHostStep, 7 TlPolicyWeights, 8 TlEmploymentParams, 9 TlEmploymentState) so bindings can verify their layout; -1 otherwise */
int tl_abi_struct_size(int idx);
/* number of SMs of the current device, or a negative error */
int tl_device_sm_count(void);

/* O1-O8: observation build for every agent of the batch (no state change except consuming TL_F_CTX_RESET). */
int tl_obs_build(const TlTownBatch* batch, const TlObsParams* obs, const TlObsOut* out, tl_stream_t stream);

/* D1 + T1 + R1 + K: need decay, faint predicate, reward with guardrails, per-town KPIs. */
int tl_decay_reward(const TlTownBatch* batch, const TlRewardParams* rew, const TlRewardOut* out,
                    uint32_t phases, tl_stream_t stream);

/* Fused tick loop: for each of n_ticks ticks run the phases of the mask in
 * SimulationLoop.step order inside one launch.  obs / obs_out may be NULL
 * when TL_PHASE_OBS is not requested, rew / rew_out likewise. */
int tl_tick(const TlTownBatch* batch, const TlObsParams* obs, const TlRewardParams* rew,
            const TlObsOut* obs_out, const TlRewardOut* rew_out, uint32_t phases, int32_t n_ticks,
            tl_stream_t stream);

/* tl_tick plus the world-dynamics phases (TL_PHASE_ACTIONS, TL_PHASE_SOCIAL_DECAY, TL_PHASE_RESERVATIONS); per tick:
 * actions -> reservation refresh -> need decay -> ledger decay -> faint -> reward -> episode_tick -> observe.
 * dyn may be NULL when neither TL_PHASE_ACTIONS nor TL_PHASE_SOCIAL_DECAY is requested. */
int tl_tick_world(const TlTownBatch* batch, const TlObsParams* obs, const TlRewardParams* rew,
                  const TlDynamicsParams* dyn, const TlObsOut* obs_out, const TlRewardOut* rew_out,
                  uint32_t phases, int32_t n_ticks, tl_stream_t stream);

/* Number of kernels launched by this library since load (the bench reports it). */
int64_t tl_launch_count(void);

/*
 * Planned launches (ABI v4).  tl_tick_world validates its arguments, picks the kernel form, grid and shared-memory
 * carve-up and builds the parameter block on every call; a plan does that ONCE, after which a launch is a single
 * cudaLaunchKernel — what a rollout loop (the in-loop caller of the reference, core/sim_loop.py:632, called once per
 * tick) wants.  Every output pointer of obs_out / rew_out is the base of a ring of tick slots spaced by its
 * *_tick_stride; tl_plan_launch(plan, s, stream) writes the launch's n_ticks ticks to slots s .. s + n_ticks - 1.
 * The batch, parameter and LUT pointers are captured by value: they must stay valid while the plan lives.
 * Device pointers; asynchronous on the stream; legal inside CUDA-graph stream capture.  NULL on error.
 */
typedef struct TlPlan TlPlan;
TlPlan* tl_plan_create(const TlTownBatch* batch, const TlObsParams* obs, const TlRewardParams* rew,
                       const TlDynamicsParams* dyn, const TlObsOut* obs_out, const TlRewardOut* rew_out,
                       uint32_t phases, int32_t n_ticks);
void tl_plan_destroy(TlPlan* plan);
int tl_plan_launch(const TlPlan* plan, int64_t tick_slot, tl_stream_t stream);
/* n_launches back-to-back launches from one host call: launch i writes to slot ((slot0 / n_ticks + i) mod
 * (ring_slots / n_ticks)) * n_ticks.  timed != 0 records a CUDA event before the first and after every launch;
 * after the stream has been synchronised tl_plan_launch_ms returns the n per-launch durations in milliseconds. */
int tl_plan_launch_many(TlPlan* plan, int64_t slot0, int32_t n_launches, int64_t ring_slots, int32_t timed,
                        tl_stream_t stream);
int tl_plan_launch_ms(const TlPlan* plan, float* ms, int32_t n);
/* grid size, threads per block and dynamic shared memory the plan launches with (any pointer may be NULL) */
int tl_plan_grid(const TlPlan* plan, int32_t* grid, int32_t* threads, int32_t* smem_bytes);

/*
 * Kernel-selection switches for parity tests and A/B measurements; process-wide, read when a launch is planned
 * (never per launch, never from the environment).  Names and values:
 *   "kernel_form"     -1 by entity count (default), 0 entity-list form with role warps, 1 dense shared-memory grid form
 *   "pass_agents"     entity form: agents per pass, 1..32 (0 = default)
 *   "towns_per_cta"   entity form: towns per CTA, 1..16 (0 = chosen from the batch shape)
 *   "generic_window"  1 = runtime-window kernels instead of the compile-time window specialisations
 *   "mirror_pass"     1 = bfloat16 mirrors by a cast pass after the launch instead of inside the kernel
 *   "even_grid"       0 = do not round small grids up to a multiple of the SM count (default 1)
 *   "policy_pair"     1 (default) = tl_policy_step runs CTA pairs (a cluster of two, tcgen05 cta_group::2: each CTA stages half
 *                     of every weight chunk); 0 = single CTAs
 *   "policy_multicast" 1 = tl_policy_step runs clusters of two single-CTA pipelines that share the weight stream: each CTA
 *                     loads half of every weight chunk and TMA-multicasts it into both rings (with policy_pair = 0 only)
 *   "stream_stores"   1 = the host expander of the narrow wire formats writes the float32 map with non-temporal
 *                     stores (default 0: ordinary stores, the rows stay in the last-level cache for the consumer)
 */
int tl_set_option(const char* name, int32_t value);
void tl_reset_options(void);

/* Generalised advantage estimation over a device rollout (replaces the
 * Python double loop of compute_gae, src/townlet/policy/backends/pytorch/ppo_utils.py:55-67).
 * Layout is time-major as the tick kernels write it: rewards[T][N], dones[T][N] (1.0 = episode ended),
 * values[T + (bootstrap ? 1 : 0)][N]; outputs advantages[T][N], returns[T][N].  float32 arithmetic in
 * the reference's operation order.  Device pointers, asynchronous on the stream. */
int tl_gae(const float* rewards, const float* values, const float* dones, float* advantages, float* returns,
           int32_t n_steps, int32_t n_agents, int32_t bootstrap, double gamma, double gae_lambda, tl_stream_t stream);

/* float32 rows [n_rows][width] -> bfloat16 rows [n_rows][padded_width], zero padded (padded_width even, >= width):
 * the low-precision copy of the observation tensors the batched policy step consumes (the cast
 * PolicyRuntime's autocast would do, policy/runner.py:444-488, fused with the padding that aligns the GEMM's reduction
 * width).  Device pointers, asynchronous on the stream; round to nearest even. */
int tl_cast_pad_bf16(const float* src, void* dst, int64_t n_rows, int32_t width, int32_t padded_width, tl_stream_t stream);
/* the same with destination rows dst_stride elements apart (even, >= padded_width) */
int tl_cast_pad_bf16_strided(const float* src, void* dst, int64_t n_rows, int32_t width, int32_t padded_width,
                             int64_t dst_stride, tl_stream_t stream);

/* Sampling step of the batched policy (the reference takes log_softmax of the logits and the log
 * probability of the action per agent at batch size 1, policy/runner.py:478-488).  heads: [n_rows][row] float32
 * (TL_DTYPE_F32) or bfloat16 (TL_DTYPE_BF16) rows whose first n_actions columns are the policy logits and, when
 * `values` is not NULL, whose column n_actions is the value.  Per row: float32 log_softmax, one categorical sample by
 * the Gumbel-max trick (argmax_j logp_j - log(-log u_j)), outputs actions[r] (int64), log_probs[r] = logp of the action,
 * values[r].  u_j = ((bits >> 9) + 0.5) * 2^-23 with bits = lane j % 4 of Philox4x32-10(counter = (draw, r), key = seed),
 * draw = 4 * ((counter ? *counter : 0) + offset) + j / 4: `counter` is an optional DEVICE int64 the caller advances between
 * rollouts (read at run time, so a replayed CUDA graph draws fresh numbers), `offset` the step inside the rollout.
 * Device pointers, asynchronous on the stream. */
#define TL_MAX_ACTIONS 16
#define TL_DTYPE_F32 0
#define TL_DTYPE_BF16 1
int tl_sample_actions(const void* heads, int32_t heads_dtype, int64_t n_rows, int32_t row, int32_t n_actions, uint64_t seed,
                      const int64_t* counter, int64_t offset, int64_t* actions, float* log_probs, float* values,
                      tl_stream_t stream);

/*
 * Employment shift state machine on the device (ABI v4): EmploymentEngine.apply_job_state
 * (world/employment.py:63-142 with its helpers :229-455), the `employment` step of the world systems
 * (world/systems/__init__.py:9-19; between need decay and the ledger decay).  One thread walks the agents of one town in
 * order — the walk is sequential by construction: an agent's coworker look-up (:466-480) sees the states already
 * updated this tick for earlier agents and last tick's for later ones.  Per agent it keeps the reference's employment
 * context (TlEmploymentState, one array per key of context_defaults, :167-193) and rewrites the snapshot fields the hot
 * path reads: wallet, wages_withheld, lateness, attendance, flags (on_shift, shift state), wages_paid, punctuality.
 * What stays on the host: the two coworker ledger updates (world.update_relationship for "employment_help" /
 * "employment_shift_taken", with capacity pruning) and event emission — the kernel reports them as per-agent event bits
 * (TL_EMP_EV_*) and the host applies them; job assignment (assign_jobs_to_agents) and the exit queue.
 * `tick_offset` is added to town_tick: 1 when the step runs before the tl_tick launch that advances the tick.
 */
#define TL_MAX_JOBS 16
#define TL_MAX_ATTENDANCE_WINDOW 14 /* employment.attendance_window <= 14 (config/world_config.py:38) */
#define TL_MAX_ABSENCE_EVENTS 16
/* ctx["state"] codes */
#define TL_EMP_PRE_SHIFT 0
#define TL_EMP_IDLE 1
#define TL_EMP_AWAIT_START 2
#define TL_EMP_ON_TIME 3
#define TL_EMP_LATE 4
#define TL_EMP_ABSENT 5
#define TL_EMP_POST_SHIFT 6
/* boolean context keys, bits of TlEmploymentState.ctx_flags */
#define TL_EMP_F_LATE_PENALTY (1u << 0)
#define TL_EMP_F_ABSENCE_PENALTY (1u << 1)
#define TL_EMP_F_LATE_EVENT (1u << 2)
#define TL_EMP_F_ABSENCE_EVENT (1u << 3)
#define TL_EMP_F_DEPARTURE_EVENT (1u << 4)
#define TL_EMP_F_LATE_HELP_EVENT (1u << 5)
#define TL_EMP_F_TOOK_SHIFT_EVENT (1u << 6)
#define TL_EMP_F_LATE_COUNTER (1u << 7)
#define TL_EMP_F_EVER_ON_TIME (1u << 8)
#define TL_EMP_F_OUTCOME_RECORDED (1u << 9)
#define TL_EMP_F_SHIFT_STARTED (1u << 10) /* shift_started_tick is not None */
#define TL_EMP_F_LAST_PRESENT (1u << 11)  /* last_present_tick is not None */
/* events of a tick, bits of events_out[tick][agent] (the reference emits them through _emit_event) */
#define TL_EMP_EV_LATE_START 1u       /* shift_late_start */
#define TL_EMP_EV_ABSENT 2u           /* shift_absent */
#define TL_EMP_EV_DEPARTED_EARLY 4u   /* shift_departed_early */
#define TL_EMP_EV_HELPED_WHEN_LATE 8u /* employment_helped_when_late: the host applies the coworker ledger update */
#define TL_EMP_EV_TOOK_MY_SHIFT 16u   /* employment_took_my_shift: likewise */

typedef struct TlEmploymentParams {
    int32_t n_jobs;               /* config.jobs in order; job 0 is the default for an agent without one (:83-89) */
    int32_t arrival_buffer;       /* behavior.job_arrival_buffer */
    int32_t ticks_per_day;        /* max(1, employment.exit_review_window) */
    int32_t grace_ticks;
    int32_t absent_cutoff;
    int32_t absence_slack;
    int32_t attendance_window;    /* max(1, employment.attendance_window) <= TL_MAX_ATTENDANCE_WINDOW */
    int32_t reserved0;
    double late_tick_penalty;
    double absence_penalty;
    int32_t job_start[TL_MAX_JOBS];
    int32_t job_end[TL_MAX_JOBS];          /* end_tick or start_tick, never before the start (:92-96) */
    int32_t job_has_location[TL_MAX_JOBS];
    int32_t job_x[TL_MAX_JOBS];            /* world coordinates (town_origin is subtracted on the device) */
    int32_t job_y[TL_MAX_JOBS];
    double job_wage[TL_MAX_JOBS];          /* spec.wage_rate or economy["wage_income"] (:97) */
    double job_lateness_penalty[TL_MAX_JOBS];
} TlEmploymentParams;

typedef struct TlEmploymentState {
    int32_t* job;               /* [N] index into the job table, -1 = none */
    int32_t* state;             /* [N] TL_EMP_* */
    int32_t* current_day;       /* [N] */
    int32_t* late_ticks;        /* [N] */
    uint32_t* ctx_flags;        /* [N] TL_EMP_F_* */
    int32_t* shift_started_tick;/* [N] valid with TL_EMP_F_SHIFT_STARTED */
    int32_t* shift_end_tick;    /* [N] */
    int32_t* last_present_tick; /* [N] valid with TL_EMP_F_LAST_PRESENT */
    int32_t* scheduled_ticks;   /* [N] */
    int32_t* on_time_ticks;     /* [N] */
    double* attendance_samples; /* [N][TL_MAX_ATTENDANCE_WINDOW] oldest first (deque(maxlen=window)) */
    int32_t* attendance_count;  /* [N] */
    int32_t* absence_events;    /* [N][TL_MAX_ABSENCE_EVENTS] ticks, oldest first */
    int32_t* absence_count;     /* [N] */
    int32_t* late_ticks_today;  /* [N] snapshot.late_ticks_today */
    int32_t* absent_shifts_7d;  /* [N] snapshot.absent_shifts_7d */
    int32_t* wages_earned;      /* [N] snapshot.inventory["wages_earned"] */
    const int32_t* town_origin; /* [T][2] world coordinates of the town-relative tile (0, 0) */
} TlEmploymentState;

/* n_ticks steps of apply_job_state for every town of the batch (tick = town_tick + tick_offset + step).  The batch arrays
 * wallet, wages_withheld, lateness, attendance, wages_paid, punctuality and flags are rewritten.  events_out may be NULL;
 * otherwise [n_ticks][N] words of TL_EMP_EV_* bits.  Device pointers, asynchronous on the stream. */
int tl_employment_step(const TlTownBatch* batch, const TlEmploymentParams* params, const TlEmploymentState* state,
                       uint32_t* events_out, int32_t n_ticks, int32_t tick_offset, tl_stream_t stream);

/*
 * The whole batched policy step in ONE kernel on the tensor cores (ABI v4).  The reference evaluates
 * ConflictAwarePolicyNetwork (policy/models.py:36-69) per agent at batch size 1 and samples with torch
 * (policy/runner.py:444-488).  tl_policy_step takes the padded bfloat16 input rows the tick kernel writes
 * (TlObsOut.map_bf16 / features_bf16 side by side: 512 map columns, then 128 feature columns, rows `input_row` elements
 * apart) and per 128-agent tile chains map / feature encoders -> shared layer -> policy + value heads with tcgen05.mma
 * (operands staged by TMA, accumulators in tensor memory, hidden activations kept in shared memory), then takes the
 * float32 log_softmax and samples exactly as tl_sample_actions does.  Weights are bfloat16 in torch.nn.Linear layout
 * [out][in], the input dimension zero-padded (484 -> 512, 81 -> 128), heads = policy rows then the value row, zero rows up
 * to 16; biases float32.  heads_out (optional, [n_rows][16] float32) receives the logits / value row for tests.
 * Returns TL_ERR_UNSUPPORTED for other layer sizes (callers fall back to their own network code).
 */
typedef struct TlPolicyWeights {
    const uint16_t* w_map;    /* [256][512] */
    const uint16_t* w_feat;   /* [256][128] */
    const uint16_t* w_shared; /* [256][512]: columns 0-255 read the map encoder's output, 256-511 the feature encoder's */
    const uint16_t* w_heads;  /* [16][256] */
    const float* b_enc;       /* [512] map encoder bias, then feature encoder bias */
    const float* b_shared;    /* [256] */
    const float* b_heads;     /* [16] */
    int32_t hidden;           /* 256 */
    int32_t map_in;           /* 512 */
    int32_t feat_in;          /* 128 */
    int32_t n_actions;        /* policy logits; the value is row n_actions of w_heads */
} TlPolicyWeights;
int tl_policy_step(const void* inputs, int64_t n_rows, int64_t input_row, const TlPolicyWeights* weights, uint64_t seed,
                   const int64_t* counter, int64_t offset, int64_t* actions, float* log_probs, float* values, float* heads_out,
                   tl_stream_t stream);

/*
 * Host-buffer entries: every pointer (batch, LUTs, outputs) is HOST memory — what a binding under
 * WorldContext.observe (world/core/context.py:163-168) calls when the world state lives in host memory.
 *
 * A context owns TL_HOST_SLOTS independent slots (stream, device arena, device outputs, pinned staging).  A step
 * is SUBMITTED to a slot (host->device copies, kernels and device->host copies are enqueued, nothing waits) and
 * later WAITED for; steps in different slots overlap — the upload of one rides the bus while the download of
 * another comes back (PCIe is full duplex) — so a caller that keeps two batches in flight (two halves of its
 * towns, or two independent vectors of towns) pays max(up, down) per step instead of up + kernel + down.
 *
 * The observation map is 84 % of the bytes that come back and every cell of its integer channels is a small
 * non-negative integer (0 / 1 in hybrid and full, counts in compact).  `wire` selects how the map crosses the bus:
 *   TL_WIRE_F32   float32 cells as they are (4 bytes per cell);
 *   TL_WIRE_U8    one byte per cell of the integer channels;
 *   TL_WIRE_BITS  one bit per cell (0 / 1 channels only; a count above 1 fails the step with TL_ERR_UNSUPPORTED).
 * The two agent-independent path channels of the full variant (encoders/map.py:115-125) never cross the bus in the
 * narrow formats: the host writes them from obs->path_lut.  `map_format` selects what obs_out->map receives:
 *   TL_MAP_F32    the reference's float32 [N][C][W][W] tensor, bit-identical whatever the wire (the narrow formats
 *                 are expanded by the context's host threads inside tl_host_wait);
 *   TL_MAP_WIRE   the wire bytes themselves: rows of tl_wire_row_bytes() bytes per agent — uint8 [N][row] cells in
 *                 channel-major order, or packed bits (cell i = bit i % 32 of little-endian word i / 32), integer
 *                 channels only.  obs_out->map_tick_stride is then in BYTES.
 */
#define TL_HOST_SLOTS 4
#define TL_WIRE_F32 0
#define TL_WIRE_U8 1
#define TL_WIRE_BITS 2
#define TL_MAP_F32 0
#define TL_MAP_WIRE 1

/* index of every array of TlTownBatch in TlHostStep.dirty_fields (declaration order) */
enum {
    TL_FIELD_TOWN_AGENT_OFF = 0, TL_FIELD_TOWN_ENT_OFF, TL_FIELD_TOWN_TICK, TL_FIELD_ENT, TL_FIELD_POS, TL_FIELD_NEEDS,
    TL_FIELD_WALLET, TL_FIELD_ATTENDANCE, TL_FIELD_WAGES_WITHHELD, TL_FIELD_WAGES_PAID, TL_FIELD_PUNCTUALITY,
    TL_FIELD_EPISODE_TOTAL, TL_FIELD_SOCIAL_IN, TL_FIELD_TERM_BLOCK_TICK, TL_FIELD_LATENESS, TL_FIELD_LAST_ACTION_DURATION,
    TL_FIELD_EPISODE_TICK, TL_FIELD_SLOT, TL_FIELD_FLAGS, TL_FIELD_LAST_ACTION_HASH, TL_FIELD_PERSONALITY, TL_FIELD_TIE_COUNT,
    TL_FIELD_TIE_EMBED, TL_FIELD_TIE_TRUST, TL_FIELD_TIE_FAM, TL_FIELD_TIE_RIV, TL_FIELD_RIV_COUNT, TL_FIELD_RIV_EMBED,
    TL_FIELD_RIV_SCORE, TL_FIELD_ENT_HOLDER, TL_N_FIELDS
};
#define TL_DIRTY_ALL (~(uint64_t)0)

typedef struct TlHostStep {
    const TlTownBatch* batch;   /* HOST arrays */
    const TlObsParams* obs;     /* HOST look-up tables; may be NULL without TL_PHASE_OBS / TL_PHASE_ADVANCE */
    const TlRewardParams* rew;  /* may be NULL without TL_PHASE_DECAY / TL_PHASE_REWARD */
    const TlDynamicsParams* dyn;/* may be NULL without the world-dynamics phases; dyn->actions is HOST memory */
    const TlObsOut* obs_out;    /* HOST buffers (the bfloat16 mirrors are device-only and ignored) */
    const TlRewardOut* rew_out; /* HOST buffers */
    uint32_t phases;
    int32_t n_ticks;
    /* The slot keeps the batch RESIDENT on the device between steps.  Bit i set = array i (TL_FIELD_*) changed on the
     * host since this slot's previous step and is uploaded; clear = the device copy (as the previous steps left it,
     * including the state the ticks advanced) is used.  TL_DIRTY_ALL uploads everything; a slot's first step, or a
     * step whose sizes differ from the resident batch, always does. */
    uint64_t dirty_fields;
    int32_t wire;       /* TL_WIRE_* */
    int32_t map_format; /* TL_MAP_* */
    int32_t writeback;  /* 1 = the arrays the phases mutate are copied back into the host batch; 0 = they stay on the device */
    int32_t reserved0;
} TlHostStep;

typedef struct TlContext TlContext;
TlContext* tl_context_create(int device);
void tl_context_destroy(TlContext* ctx);
/* host threads that expand narrow wire formats and copy staged outputs inside tl_host_wait (default: up to 8) */
int tl_context_set_threads(TlContext* ctx, int32_t n_threads);
/* bytes per agent row of the map on the wire (TL_MAP_WIRE buffers are [n_ticks][N][row]); negative on error */
int64_t tl_wire_row_bytes(const TlObsParams* obs, int32_t wire);
/* enqueue one step on slot 0 <= slot < TL_HOST_SLOTS; the slot must not have a step pending.  Every host buffer of
 * the step must stay valid until tl_host_wait returns.  Pinned (page-locked) buffers are copied directly; pageable
 * ones go through the context's pinned staging and are filled inside tl_host_wait. */
int tl_host_submit(TlContext* ctx, int32_t slot, const TlHostStep* step);
/* wait for the slot's pending step, expand / scatter what was staged; after it returns every output of the step
 * (and, with writeback, the mutated state arrays) is in the caller's buffers */
int tl_host_wait(TlContext* ctx, int32_t slot);
/* one synchronous step on slot 0: float32 wire, everything uploaded, state written back (ABI v1-v3 behaviour) */
int tl_host_tick(TlContext* ctx, const TlTownBatch* host_batch, const TlObsParams* obs,
                 const TlRewardParams* rew, const TlObsOut* host_obs_out, const TlRewardOut* host_rew_out,
                 uint32_t phases, int32_t n_ticks);
/* tl_host_tick plus the world-dynamics phases: dyn->actions is HOST memory too; positions, the agents' entity words,
 * the last-action fields and the ledgers are copied back with the rest of the mutated state. */
int tl_host_tick_world(TlContext* ctx, const TlTownBatch* host_batch, const TlObsParams* obs,
                       const TlRewardParams* rew, const TlDynamicsParams* dyn, const TlObsOut* host_obs_out,
                       const TlRewardOut* host_rew_out, uint32_t phases, int32_t n_ticks);
/* bytes moved over the bus by the last submitted step */
int64_t tl_context_last_h2d_bytes(const TlContext* ctx);
int64_t tl_context_last_d2h_bytes(const TlContext* ctx);

#ifdef __cplusplus
}
#endif
#endif /* TOWNLET_B200_H */
