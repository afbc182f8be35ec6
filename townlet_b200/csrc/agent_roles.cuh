// agent_roles.cuh — the thread-per-agent parts of the tick, shared by the role-parallel entity kernel
// (tick_roles.cuh) and the dense grid kernel (townlet_b200.cu).  Included inside the anonymous namespace
// of townlet_b200.cu; uses KParams, RewardIn / reward_agent, pos_x / pos_y.

// -DTL_ROLE_TIMING: two CTAs print the cycles each role spends per tick in its role body and at the pass
// barrier + sweep, and CTA 0 the parts of its state role (scripts/role_timing.sh; never part of the shipped library)
#ifdef TL_ROLE_TIMING
#define TL_TIMING(...) __VA_ARGS__
__device__ long long g_state_parts[8];
#define TL_MARK_BEGIN() long long tl_mark_t = clock64()
#define TL_MARK(k)                                                         \
    do {                                                                   \
        const long long tl_now = clock64();                                \
        if (blockIdx.x == 0 && threadIdx.x == 0) g_state_parts[k] += tl_now - tl_mark_t; \
        tl_mark_t = tl_now;                                                \
    } while (0)
#else
#define TL_TIMING(...)
#define TL_MARK_BEGIN()
#define TL_MARK(k)
#endif

struct StateRoleOut {
    double need0 = 0.0;         // hunger after decay (KPI)
    double reward_total = 0.0;  // reward of the tick (KPI)
    double kpi_wallet = 0.0;
    int kpi_late = 0;
    bool terminated = false;    // effective terminated flag of the tick
};

// What process_actions leaves on ONE agent (world/systems/affordances.py:124-229 after
// world/actions.py:44-105).  A move with a position relocates the agent — snapshot.position and the spatial
// index (world/spatial.py:62-83), i.e. the agent's own entity word — and succeeds; every other kind carries
// the success the host services resolved; apply_affordance_outcome (world/affordances/core.py:198-206)
// records kind, success and duration.  The caller's copies of the fields are updated along with memory
// (flags are left to the caller, which rewrites them anyway).
__device__ __forceinline__ void apply_action_agent(const KParams& p, int a, int self_e, uint32_t w, uint32_t& fl,
                                                   int& duration, float& last_hash, int32_t& pw) {
    const unsigned kind = TL_ACT_KIND(w);
    bool success = TL_ACT_SUCCESS(w) != 0;
    if (kind == TL_ACT_MOVE) {
        success = TL_ACT_HAS_POS(w) != 0;
        if (success) {
            const uint32_t x = (uint32_t)TL_ACT_X(w), y = (uint32_t)TL_ACT_Y(w);
            pw = (int32_t)(x | (y << 16));
            p.b.pos[a] = pw;
            p.b.ent[self_e] = (p.b.ent[self_e] & ~0xfffffu) | x | (y << 10);
        }
    }
    fl = (fl & ~TL_F_LAST_SUCCESS) | (success ? TL_F_LAST_SUCCESS : 0u);
    duration = TL_ACT_DURATION(w);
    last_hash = p.dyn.action_hash[kind];
    p.b.last_action_duration[a] = duration;
    p.b.last_action_hash[a] = last_hash;
}

// The position part of an action word on top of a packed position / entity word: what every reader of
// pos / ent uses inside a launch that applies actions, so that it does not depend on whether the owning
// thread has written the move back yet.
__device__ __forceinline__ int32_t moved_pos(int32_t pw, uint32_t act_w) {
    if (TL_ACT_KIND(act_w) == TL_ACT_MOVE && TL_ACT_HAS_POS(act_w))
        return (int32_t)((uint32_t)TL_ACT_X(act_w) | ((uint32_t)TL_ACT_Y(act_w) << 16));
    return pw;
}
__device__ __forceinline__ uint32_t moved_ent(uint32_t ew, uint32_t act_w) {
    if (TL_ACT_KIND(act_w) == TL_ACT_MOVE && TL_ACT_HAS_POS(act_w))
        return (ew & ~0xfffffu) | (uint32_t)TL_ACT_X(act_w) | ((uint32_t)TL_ACT_Y(act_w) << 10);
    return ew;
}

// refresh_reservations for ONE town by one warp (world/systems/queues.py:55-86 driven per tick by queues.step, :14-31):
// the reference walks the objects in order and, for each, adds the object's tile to the reservation set when the queue
// manager has an active agent for it and discards the tile otherwise (world/spatial.py:106-119) — so a tile ends up
// reserved exactly when the LAST object standing on it is held.  Here every object with a position owns a reservation
// slot in the town's entity list (ent_holder >= -1, in object order); the slot becomes a reserved-tile word when its
// object is the last slot on the tile and is held, an empty word otherwise.  Agents that hold any object get
// TL_F_RESERVATION (features.py:215-219), all others lose it.  Entities with ent_holder == -2 are left alone.
__device__ __forceinline__ void refresh_reservations_town(const KParams& p, int abeg, int aend, int ebeg, int eend, int lane) {
    const unsigned FULL = 0xffffffffu;
    for (int a = abeg + lane; a < aend; a += 32) p.b.flags[a] &= ~TL_F_RESERVATION;
    __syncwarp();
    const int n_agents = aend - abeg;
    const int e0 = ebeg + n_agents;  // the agents' own words lead the list
#pragma unroll 1
    for (int base = e0; base < eend; base += 32) {
        const int e = base + lane;
        const bool in = e < eend;
        const uint32_t w = in ? p.b.ent[e] : 0u;
        const int holder = in ? __ldg(p.b.ent_holder + e) : -2;
        const bool slot = holder >= -1;
        const uint32_t tile = w & 0xfffffu;
        // a later slot on the same tile decides instead of this one: first within the 32 entities of this step ...
        const unsigned peers = __match_any_sync(FULL, slot ? tile : (0x100000u | (unsigned)lane));
        bool later = slot && (peers >> lane) > 1u;
        // ... then in the rest of the list (towns with more than 32 objects)
#pragma unroll 1
        for (int nb = base + 32; nb < eend; nb += 32) {
            const int ne = nb + lane;
            const bool nin = ne < eend;
            const uint32_t nw = nin ? p.b.ent[ne] : 0u;
            const bool nslot = nin && __ldg(p.b.ent_holder + ne) >= -1;
            const uint32_t ntile = nslot ? (nw & 0xfffffu) : 0xffffffffu;
#pragma unroll 1
            for (int i = 0; i < 32; ++i) {
                const uint32_t other = __shfl_sync(FULL, ntile, i);  // every lane takes part, whatever its own state
                later = later || (slot && other == tile);
            }
        }
        if (slot) {
            const uint32_t kind = (holder >= 0 && !later) ? TL_KIND_RESERVED : TL_KIND_EMPTY;
            const uint32_t nw = (w & ~(3u << 20)) | (kind << 20);
            if (nw != w) p.b.ent[e] = nw;
            if (holder >= 0 && holder < n_agents) atomicOr(p.b.flags + abeg + holder, TL_F_RESERVATION);
        }
    }
}

// The scalar inputs of ONE agent for one tick (AgentSnapshot fields, world/agents/snapshot.py:19-44, the
// employment context and the reward engine's per-agent state).
struct StateIn {
    double need[3];
    double wallet, attendance, wages_withheld;
    RewardIn rin;
    uint32_t fl;
    int episode_tick, lateness, duration, slot_idx;
    float last_hash;
    int32_t pw;
};

// Staging rows of the state role in shared memory: kStateF64 rows of 32 doubles, then kStateI32 rows of
// 32 four-byte values (lane = agent of the pass).
constexpr int kStateF64 = 12, kStateI32 = 8;
constexpr int kStateStageBytes = 32 * (8 * kStateF64 + 4 * kStateI32);

// Straight from global memory (the grid form).
template <bool MUT>
__device__ __forceinline__ StateIn load_state_in(const KParams& p, int a) {
    const size_t N = (size_t)p.b.n_agents;
    StateIn in;
    in.need[0] = p.b.needs[a];
    in.need[1] = p.b.needs[N + a];
    in.need[2] = p.b.needs[2 * N + a];
    in.fl = p.b.flags[a];
    in.episode_tick = p.b.episode_tick[a];
    in.rin = load_reward_in(p, a);
    in.wallet = __ldg(p.b.wallet + a);
    in.lateness = __ldg(p.b.lateness + a);
    in.attendance = __ldg(p.b.attendance + a);
    in.wages_withheld = in.rin.wages_withheld;
    in.duration = ld_state<MUT>(p.b.last_action_duration + a);
    in.slot_idx = __ldg(p.b.slot + a);
    in.last_hash = ld_state<MUT>(p.b.last_action_hash + a);
    in.pw = ld_state<MUT>(p.b.pos + a);
    return in;
}

// Asynchronous global -> shared copies of the same inputs for agent `a` into column `lane` of the
// staging rows; read them back with staged_state_in after cp_async_wait_all() (same thread).
__device__ __forceinline__ void prefetch_state_in(const KParams& p, unsigned char* stage, int a, int lane) {
    const size_t N = (size_t)p.b.n_agents;
    double* d = reinterpret_cast<double*>(stage) + lane;
    int32_t* w = reinterpret_cast<int32_t*>(stage + 32 * 8 * kStateF64) + lane;
    cp_async8(d + 0 * 32, p.b.needs + a);
    cp_async8(d + 1 * 32, p.b.needs + N + a);
    cp_async8(d + 2 * 32, p.b.needs + 2 * N + a);
    cp_async8(d + 3 * 32, p.b.wallet + a);
    cp_async8(d + 4 * 32, p.b.attendance + a);
    cp_async8(d + 5 * 32, p.b.wages_withheld + a);
    cp_async8(d + 6 * 32, p.b.wages_paid + a);
    cp_async8(d + 7 * 32, p.b.punctuality + a);
    cp_async8(d + 8 * 32, p.b.episode_total + a);
    if (p.b.social_in) {
        cp_async8(d + 9 * 32, p.b.social_in + a);
        cp_async8(d + 10 * 32, p.b.social_in + N + a);
        cp_async8(d + 11 * 32, p.b.social_in + 2 * N + a);
    }
    cp_async4(w + 0 * 32, p.b.flags + a);
    cp_async4(w + 1 * 32, p.b.episode_tick + a);
    cp_async4(w + 2 * 32, p.b.term_block_tick + a);
    cp_async4(w + 3 * 32, p.b.lateness + a);
    cp_async4(w + 4 * 32, p.b.last_action_duration + a);
    cp_async4(w + 5 * 32, p.b.slot + a);
    cp_async4(w + 6 * 32, p.b.last_action_hash + a);
    cp_async4(w + 7 * 32, p.b.pos + a);
}

__device__ __forceinline__ StateIn staged_state_in(const KParams& p, const unsigned char* stage, int lane) {
    const double* d = reinterpret_cast<const double*>(stage) + lane;
    const int32_t* w = reinterpret_cast<const int32_t*>(stage + 32 * 8 * kStateF64) + lane;
    StateIn in;
    in.need[0] = d[0 * 32];
    in.need[1] = d[1 * 32];
    in.need[2] = d[2 * 32];
    in.wallet = d[3 * 32];
    in.attendance = d[4 * 32];
    in.wages_withheld = d[5 * 32];
    in.rin.wages_withheld = in.wages_withheld;
    in.rin.wages_paid = d[6 * 32];
    in.rin.punctuality = d[7 * 32];
    in.rin.episode_total = d[8 * 32];
    in.rin.sb = in.rin.sp = in.rin.sa = 0.0;
    if (p.b.social_in) {
        in.rin.sb = d[9 * 32];
        in.rin.sp = d[10 * 32];
        in.rin.sa = d[11 * 32];
    }
    in.fl = (uint32_t)w[0 * 32];
    in.episode_tick = w[1 * 32];
    in.rin.block = w[2 * 32];
    in.lateness = w[3 * 32];
    in.duration = w[4 * 32];
    in.slot_idx = w[5 * 32];
    in.last_hash = __int_as_float(w[6 * 32]);
    in.pw = w[7 * 32];
    return in;
}

// slot / max_slots for a slot outside the host table (never with the reference's allocator): kept out of line so the
// float64 division does not sit in the hot instruction stream
__device__ __noinline__ float slot_norm_out_of_table(int slot_idx, int max_slots) {
    return (float)((double)slot_idx / (double)max_slots);
}

// Nearest object of every landmark type (world/observe.py:109-147) and the landmark features
// (_encode_landmarks, features.py:333-356) for configurations with include_targets; out of line, the default
// configuration never runs it.  Returns the path-hint type's squared distance (or -1) and offsets.
struct NearestHint {
    int d2, dx, dy;
};
template <bool MUT>
__device__ __noinline__ NearestHint landmarks_all_types(const uint32_t* ent, int e_obj, int e_end, int cx, int cy, int hint_ltype,
                                                        float* dst0) {
    int best_d2[5] = {-1, -1, -1, -1, -1};
    int best_dx[5] = {0, 0, 0, 0, 0}, best_dy[5] = {0, 0, 0, 0, 0};
#pragma unroll 1
    for (int e = e_obj; e < e_end; ++e) {
        const uint32_t w = ld_state<MUT>(ent + e);
        if (TL_ENT_KIND(w) != TL_KIND_OBJECT || !TL_ENT_LM(w)) continue;
        const int lt = (int)TL_ENT_LTYPE(w);
        if (lt == 0) continue;
        const int dx = TL_ENT_X(w) - cx, dy = TL_ENT_Y(w) - cy;
        const int d2 = dx * dx + dy * dy;
#pragma unroll
        for (int l = 1; l <= 4; ++l) {
            if (l == lt && (best_d2[l] < 0 || d2 < best_d2[l])) {
                best_d2[l] = d2;
                best_dx[l] = dx;
                best_dy[l] = dy;
            }
        }
    }
    NearestHint hint{-1, 0, 0};
#pragma unroll
    for (int l = 1; l <= 4; ++l) {
        if (l == hint_ltype) hint = NearestHint{best_d2[l], best_dx[l], best_dy[l]};
        float* dst = dst0 + 3 * (l - 1);
        if (best_d2[l] < 0) {
            dst[0] = dst[1] = dst[2] = 0.0f;
        } else {
            // sqrt of an exact integer: equals np.hypot after the float32 store
            const double distance = sqrt((double)best_d2[l]);
            const double norm = distance > 0.0 ? distance : 1.0;
            dst[0] = (float)((double)best_dx[l] / norm);
            dst[1] = (float)((double)best_dy[l] / norm);
            dst[2] = (float)distance;
        }
    }
    return hint;
}

// Role "state" for ONE agent: float64 need decay (world/grid.py:1439-1455), faint predicate
// (lifecycle/manager.py:39-45), episode tick (world/core/context.py:283-289), reward with guardrails
// (rewards/engine.py:71-172) and, when `f` is not null, the scalar features, path hint, landmarks and
// personality channels of encode_feature_vector (encoders/features.py:43-404).
// MUT: the world-dynamics phases rewrite pos / ent / the last-action fields inside the launch, so they
// are read with plain loads; otherwise through the read-only path.
template <bool MUT>
__device__ __forceinline__ void state_role_agent(const KParams& p, int it, int a, int tick, int town_abeg, int town_aend,
                                                 int town_ebeg, int town_eend, uint32_t act_w, const StateIn& in,
                                                 float* f, StateRoleOut& out) {
    const TlObsParams& o = p.obs;
    const int N = p.b.n_agents;
    TL_MARK_BEGIN();
    const bool do_decay = p.phases & TL_PHASE_DECAY;
    const bool do_reward = p.phases & TL_PHASE_REWARD;
    const bool do_obs = (p.phases & TL_PHASE_OBS) && f != nullptr;
    const bool do_advance = p.phases & TL_PHASE_ADVANCE;
    const bool do_state = do_decay || do_reward;
    double& need0 = out.need0;
    double& reward_total = out.reward_total;
    double& kpi_wallet = out.kpi_wallet;
    int& kpi_late = out.kpi_late;
    bool& terminated = out.terminated;
    double need[3] = {in.need[0], in.need[1], in.need[2]};
    uint32_t fl = in.fl;
    int episode_tick = in.episode_tick;
    const RewardIn& rin = in.rin;
    const double wallet = in.wallet;
    const int lateness = in.lateness;
    kpi_wallet = wallet;
    kpi_late = lateness;
    const double attendance = in.attendance, wages_withheld = in.wages_withheld;
    int duration = in.duration;
    const int slot_idx = in.slot_idx;
    float last_hash = in.last_hash;
    int32_t pw = in.pw;
    if (MUT && TL_ACT_KIND(act_w) != TL_ACT_NONE) {
        // TL_PHASE_ACTIONS, applied by the agent's own thread (the tiles role derives the same positions from
        // the action words, so no barrier is needed between this write-back and its reads)
        apply_action_agent(p, a, town_ebeg + (a - town_abeg), act_w, fl, duration, last_hash, pw);
        p.b.flags[a] = fl;
    }
    const unsigned profile = (fl & TL_F_PROFILE_MASK) >> TL_F_PROFILE_SHIFT;
    if (do_decay) {  // world/grid.py:1439-1455
#pragma unroll
        for (int k = 0; k < TL_N_NEEDS; ++k) {
            double mult = 1.0;
            if (p.rew.decay_personality && profile != 0) {
                mult = p.rew.need_mult[profile][k];
                if (mult <= 0.0) mult = 1.0;
            }
            const double adjusted = p.rew.decay_rate[k] * mult;
            need[k] = fmax(0.0, need[k] - adjusted);
            p.b.needs[(size_t)k * N + a] = need[k];
        }
    }
    if (do_state && p.rew.fuse_faint) {  // lifecycle/manager.py:39-45
        fl = (fl & ~TL_F_FAINTED) | (need[0] <= p.rew.faint_threshold ? TL_F_FAINTED : 0u);
        p.b.flags[a] = fl;
    }
    terminated = (fl & (TL_F_TERMINATED | TL_F_FAINTED)) != 0;
    unsigned reason = (fl & TL_F_REASON_MASK) >> TL_F_REASON_SHIFT;
    if (reason == 0 && (fl & TL_F_FAINTED)) reason = 1;
    if (do_advance) {  // world/core/context.py:283-289
        const int span = max(1, o.ticks_per_day);
        // (episode_tick + 1) % span; the division only runs for a value outside [0, span) (never after the first tick)
        episode_tick += 1;
        if ((unsigned)episode_tick >= (unsigned)span) {
            episode_tick %= span;
            if (episode_tick < 0) episode_tick += span;  // Python's modulo is never negative
        }
        p.b.episode_tick[a] = episode_tick;
    }
    TL_MARK(1);
    if (do_reward) {
        RewardResult rr;
        reward_agent(p.rew, rin, tick, need, fl, terminated, reason, rr);
        p.b.term_block_tick[a] = rr.block;
        p.b.episode_total[a] = rr.cumulative;
        reward_total = rr.total;
        if (p.ro.reward) p.ro.reward[(size_t)it * p.ro.reward_tick_stride + a] = (float)rr.total;
        if (p.ro.comps) {
            double* dst = p.ro.comps + (size_t)it * p.ro.comps_tick_stride + (size_t)a * TL_N_COMPS;
#pragma unroll
            for (int i = 0; i < TL_N_COMPS; i += 2)
                *reinterpret_cast<double2*>(dst + i) = make_double2(rr.comps[i], rr.comps[i + 1]);
        }
    }
    if (do_state && p.ro.terminated)
        p.ro.terminated[(size_t)it * p.ro.terminated_tick_stride + a] = terminated ? 1 : 0;
    need0 = need[0];
    TL_MARK(2);

    if (do_obs) {  // encoders/features.py:43-404
        f[TL_FEAT_NEED_HUNGER] = (float)need[0];
        f[TL_FEAT_NEED_HYGIENE] = (float)need[1];
        f[TL_FEAT_NEED_ENERGY] = (float)need[2];
        f[TL_FEAT_WALLET] = (float)wallet;
        f[TL_FEAT_LATENESS] = (float)lateness;
        f[TL_FEAT_ON_SHIFT] = (fl & TL_F_ON_SHIFT) ? 1.0f : 0.0f;
        int m = tick % o.ticks_per_day;
        if (m < 0) m += o.ticks_per_day;
        const float2 sc = __ldg(reinterpret_cast<const float2*>(o.time_lut + 2 * m));
        f[TL_FEAT_TIME_SIN] = sc.x;
        f[TL_FEAT_TIME_COS] = sc.y;
        f[TL_FEAT_ATTENDANCE] = (float)attendance;
        f[TL_FEAT_WAGES_WITHHELD] = (float)wages_withheld;
        unsigned st = (fl & TL_F_SHIFT_MASK) >> TL_F_SHIFT_SHIFT;
        if (st > 4) st = 0;
#pragma unroll
        for (unsigned k = 0; k < 5; ++k) f[TL_FEAT_SHIFT_PRE + k] = (k == st) ? 1.0f : 0.0f;
        f[TL_FEAT_SLOT_NORM] = (slot_idx >= 0 && slot_idx < o.max_slots) ? __ldg(o.slot_lut + slot_idx)
                                                                         : slot_norm_out_of_table(slot_idx, o.max_slots);
        f[TL_FEAT_CTX_RESET] = ((fl & TL_F_CTX_RESET) || terminated) ? 1.0f : 0.0f;  // service.py:224-229
        if (fl & TL_F_CTX_RESET) {  // a pending request is reported once (consume_ctx_reset_requests, grid.py:932-936)
            fl &= ~TL_F_CTX_RESET;
            p.b.flags[a] = fl;
        }
        f[TL_FEAT_EPISODE_PROGRESS] =
            episode_tick <= 0 ? 0.0f : __ldg(o.progress_lut + min(episode_tick, o.ticks_per_day));
        f[TL_FEAT_LAST_HASH] = last_hash;
        f[TL_FEAT_LAST_SUCCESS] = (fl & TL_F_LAST_SUCCESS) ? 1.0f : 0.0f;
        f[TL_FEAT_LAST_DURATION] = (float)duration;
        f[TL_FEAT_RESERVATION] = (fl & TL_F_RESERVATION) ? 1.0f : 0.0f;
        f[TL_FEAT_IN_QUEUE] = (fl & TL_F_IN_QUEUE) ? 1.0f : 0.0f;
        // nearest object per landmark type: first minimum of the squared
        // distance in insertion order (world/observe.py:109-147)
        TL_MARK(3);
        const int cx = pos_x(pw), cy = pos_y(pw);
        const bool all_types = o.landmark_base >= 0;
        const int e_obj = town_ebeg + (town_aend - town_abeg);
        int hd2 = -1, hdx = 0, hdy = 0;
        if (!all_types) {
            // object, nearest-of-type candidate, landmark type of the path hint: one masked compare
            const uint32_t lm_mask = (3u << 20) | (7u << 22) | (1u << 29);
            const uint32_t lm_want = (TL_KIND_OBJECT << 20) | ((uint32_t)o.path_hint_ltype << 22) | (1u << 29);
            // branch-free body: the lanes of a warp scan different towns, so a `continue` would diverge every step
#pragma unroll 4
            for (int e = e_obj; e < town_eend; ++e) {
                const uint32_t w = ld_state<MUT>(p.b.ent + e);
                const int dx = TL_ENT_X(w) - cx, dy = TL_ENT_Y(w) - cy;
                const int d2 = dx * dx + dy * dy;
                const bool better = (w & lm_mask) == lm_want && (hd2 < 0 || d2 < hd2);
                hd2 = better ? d2 : hd2;
                hdx = better ? dx : hdx;
                hdy = better ? dy : hdy;
            }
        } else {
            const NearestHint hint = landmarks_all_types<MUT>(p.b.ent, e_obj, town_eend, cx, cy, o.path_hint_ltype, f + o.landmark_base);
            hd2 = hint.d2, hdx = hint.dx, hdy = hint.dy;
        }
        TL_MARK(4);
        {  // _encode_path_hint, features.py:251-286
            // The reference divides in float64 and stores float32: float32(round64(a / b)).  For integers
            // 0 <= a <= b < 2^24 that equals the correctly rounded float32 quotient: a float32 rounding boundary
            // m (25 significant bits) other than a / b itself is at least (a / b) * 2^-25 / b away from a / b,
            // round64 moves a / b by at most (a / b) * 2^-53, so both sit on the same side of every boundary.
            // One float32 division instead of two float64 ones (checked against the oracle's float64 form).
            float north = 0.0f, south = 0.0f, east = 0.0f, west = 0.0f;
            if (hd2 > 0) {
                const float dn = (float)(abs(hdx) + abs(hdy));
                // max(0, -dy)/norm etc.: one of each opposite pair is exactly 0.0
                const float ns = __fdiv_rn((float)abs(hdy), dn), ew = __fdiv_rn((float)abs(hdx), dn);
                north = hdy < 0 ? ns : 0.0f;
                south = hdy > 0 ? ns : 0.0f;
                east = hdx > 0 ? ew : 0.0f;
                west = hdx < 0 ? ew : 0.0f;
            }
            f[TL_FEAT_PATH_NORTH + 0] = north;
            f[TL_FEAT_PATH_NORTH + 1] = south;
            f[TL_FEAT_PATH_NORTH + 2] = east;
            f[TL_FEAT_PATH_NORTH + 3] = west;
        }
        if (o.personality_base >= 0) {  // features.py:359-380
#pragma unroll
            for (int k = 0; k < 3; ++k)
                f[o.personality_base + k] = p.b.personality ? __ldg(p.b.personality + (size_t)k * N + a) : 0.0f;
        }
        TL_MARK(5);
    }
}

// Role "social" for ONE agent: rivalry features (encoders/features.py:232-248) and the social snippet
// (encoders/social.py:27-254).  kt / kf / ktr are the agent's tie rows (trust, familiarity, rivalry),
// `score` its rivalry-ledger row, temb / remb the id-embedding words of the two tables; the rows may live
// in shared memory (staged by the caller) or in global memory.  `elut` is the byte -> float32 table of
// the id embedding.
__device__ __forceinline__ void social_role_agent(const KParams& p, int n_riv, int n_t, const double* kt,
                                                  const double* kf, const double* ktr, const double* score,
                                                  const uint64_t* temb, const uint64_t* remb, const float* elut,
                                                  float* f, int& filled, int& source) {
    const TlObsParams& o = p.obs;
    const int EW = p.b.embed_words;
    filled = 0;
    source = 0;
    {  // rivalry_top(agent, max_edges): rows are pre-sorted (features.py:232-248)
        const int nr = min(n_riv, o.rivalry_max_edges);
        int avoid = 0;
        double mx = 0.0;
#pragma unroll 1
        for (int j = 0; j < nr; ++j) {
            const double v = score[j];
            if (j == 0) mx = v;
            avoid += (v >= o.avoid_threshold) ? 1 : 0;
        }
        f[TL_FEAT_RIVALRY_MAX] = (float)mx;
        f[TL_FEAT_RIVALRY_AVOID] = (float)avoid;
    }
    if (o.social_base >= 0) {  // encoders/social.py:27-254
        const int slot_dim = o.embed_dim + 3;
        const int total_slots = o.top_friends + o.top_rivals;
        float* srow = f + o.social_base;
        const bool from_ties = n_t > 0;  // _resolve_relationships, social.py:165-209
        const int n_rel = from_ties ? n_t : min(n_riv, o.top_rivals);
        source = from_ties ? 1 : (n_rel > 0 ? 2 : 0);
        const double* kr = from_ties ? ktr : score;
        const uint64_t* ke = from_ties ? temb : remb;
        // sorted(..., reverse=True)[:k] is stable: repeated FIRST argmax (social.py:137-156).
        // In the rivalry fallback trust = familiarity = 0, so friends are the first entries.
        unsigned chosen = 0, sel = 0;
        const int n_friends = min(n_rel, o.top_friends);
        const int n_rivals = min(o.top_rivals, n_rel - n_friends);
        filled = n_friends + n_rivals;
        if (from_ties) {
#pragma unroll 1
            for (int s = 0; s < n_friends; ++s) {  // friends by trust + familiarity
                int best = -1;
                double bk = 0.0;
#pragma unroll 1
                for (int j = 0; j < n_rel; ++j) {
                    const double key = kt[j] + kf[j];
                    if (!((chosen >> j) & 1u) && (best < 0 || key > bk)) {
                        best = j;
                        bk = key;
                    }
                }
                chosen |= 1u << best;
                sel |= (unsigned)best << (4 * s);
            }
        } else {  // every friend key is 0.0: the first entries in order
#pragma unroll 1
            for (int s = 0; s < n_friends; ++s) sel |= (unsigned)s << (4 * s);
            chosen = (1u << n_friends) - 1u;
        }
#pragma unroll 1
        for (int s = n_friends; s < filled; ++s) {  // rivals by rivalry among the rest
            int best = -1;
            double bk = 0.0;
#pragma unroll 1
            for (int j = 0; j < n_rel; ++j) {
                const double key = kr[j];
                if (!((chosen >> j) & 1u) && (best < 0 || key > bk)) {
                    best = j;
                    bk = key;
                }
            }
            chosen |= 1u << best;
            sel |= (unsigned)best << (4 * s);
        }
        double tsum = 0.0, rsum = 0.0, tmax = 0.0, rmax = 0.0;
#pragma unroll 1
        for (int s = 0; s < total_slots; ++s) {
            float* dst = srow + s * slot_dim;
            if (s >= filled) {  // _empty_relationship_entry, social.py:226-233
                if (o.embed_dim == 8) {
#pragma unroll
                    for (int d = 0; d < 11; ++d) dst[d] = 0.0f;
                } else {
#pragma unroll 1
                    for (int d = 0; d < slot_dim; ++d) dst[d] = 0.0f;
                }
                continue;
            }
            const int j = (sel >> (4 * s)) & 15;
            const double trust = from_ties ? kt[j] : 0.0;
            const double fam = from_ties ? kf[j] : 0.0;
            const double riv = kr[j];
            if (o.embed_dim == 8) {  // _embed_agent_id, social.py:236-242; bytes via PRMT
                const uint64_t word = ke[j * EW];
                const unsigned lo = (unsigned)word, hi = (unsigned)(word >> 32);
                dst[0] = elut[__byte_perm(lo, 0, 0x4440)];
                dst[1] = elut[__byte_perm(lo, 0, 0x4441)];
                dst[2] = elut[__byte_perm(lo, 0, 0x4442)];
                dst[3] = elut[__byte_perm(lo, 0, 0x4443)];
                dst[4] = elut[__byte_perm(hi, 0, 0x4440)];
                dst[5] = elut[__byte_perm(hi, 0, 0x4441)];
                dst[6] = elut[__byte_perm(hi, 0, 0x4442)];
                dst[7] = elut[__byte_perm(hi, 0, 0x4443)];
            } else {
                uint64_t word = 0;
#pragma unroll 1
                for (int d = 0; d < o.embed_dim; ++d) {
                    if ((d & 7) == 0) word = ke[j * EW + (d >> 3)];
                    dst[d] = elut[(unsigned)(word >> (8 * (d & 7))) & 255u];
                }
            }
            dst[o.embed_dim + 0] = (float)trust;
            dst[o.embed_dim + 1] = (float)fam;
            dst[o.embed_dim + 2] = (float)riv;
            tsum += trust;  // numpy sums left to right below 8 values
            rsum += riv;
            tmax = s == 0 ? trust : fmax(tmax, trust);
            rmax = s == 0 ? riv : fmax(rmax, riv);
        }
        if (o.include_aggregates) {  // _compute_aggregates, social.py:90-105, 245-254
            float* agg = srow + total_slots * slot_dim;
            if (filled > 0) {
                if (filled == 8) {  // numpy's unrolled pairwise block of 8
                    double tv[8], rv[8];
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        const int j = (sel >> (4 * s)) & 15;
                        tv[s] = from_ties ? kt[j] : 0.0;
                        rv[s] = kr[j];
                    }
                    tsum = ((tv[0] + tv[1]) + (tv[2] + tv[3])) + ((tv[4] + tv[5]) + (tv[6] + tv[7]));
                    rsum = ((rv[0] + rv[1]) + (rv[2] + rv[3])) + ((rv[4] + rv[5]) + (rv[6] + rv[7]));
                }
                agg[0] = (float)(tsum / (double)filled);
                agg[1] = (float)tmax;
                agg[2] = (float)(rsum / (double)filled);
                agg[3] = (float)rmax;
            } else {
                agg[0] = agg[1] = agg[2] = agg[3] = 0.0f;
            }
        }
    }
}

// TL_PHASE_ACTIONS for the agents of one CTA's towns [t0, t0 + n_towns), all threads of the CTA (the grid
// form, which stages the entity words of a tick before any agent is served).  The caller puts a block
// barrier between this and any read of pos / ent / the last-action fields.
__device__ __forceinline__ void apply_actions_cta(const KParams& p, int it, int t0, int n_towns, int tid, int n_threads) {
    const int a_begin = p.b.town_agent_off[t0], a_end = p.b.town_agent_off[t0 + n_towns];
    const uint32_t* act = p.dyn.actions + (size_t)it * p.dyn.actions_tick_stride;
#pragma unroll 1
    for (int a = a_begin + tid; a < a_end; a += n_threads) {
        const uint32_t w = __ldg(act + a);
        if (TL_ACT_KIND(w) == TL_ACT_NONE) continue;
        int s = 0;
        while (s + 1 < n_towns && a >= p.b.town_agent_off[t0 + s + 1]) ++s;
        const int self_e = p.b.town_ent_off[t0 + s] + (a - p.b.town_agent_off[t0 + s]);
        uint32_t fl = p.b.flags[a];
        int duration;
        float last_hash;
        int32_t pw;
        apply_action_agent(p, a, self_e, w, fl, duration, last_hash, pw);
        p.b.flags[a] = fl;
    }
}

// _decay_value, world/relationships.py:155-162
__device__ __forceinline__ double decay_value(double value, double decay) {
    if (decay <= 0.0) return value;
    if (value > 0.0) return fmax(0.0, value - decay);
    if (value < 0.0) return fmin(0.0, value + decay);
    return 0.0;
}

// TL_PHASE_SOCIAL_DECAY for ONE agent, in place on its rows (shared or global memory):
// RivalryLedger.decay(ticks=1) (world/rivalry.py:68-83) then RelationshipLedger.decay
// (world/relationships.py:78-89), the order of RelationshipService.decay
// (world/agents/relationships_service.py:165-191).  Evicted entries leave the row, the rest keep their
// order, the freed tail is zeroed.  n_riv / n_t come in clamped to the row stride and go out updated.
__device__ __forceinline__ void social_decay_agent(const KParams& p, int& n_riv, int& n_t, double* kt, double* kf,
                                                   double* ktr, double* score, uint64_t* temb, uint64_t* remb) {
    const TlDynamicsParams& d = p.dyn;
    const int EW = p.b.embed_words;
    {
        int w = 0;
#pragma unroll 1
        for (int j = 0; j < n_riv; ++j) {
            const double v = fmax(d.rivalry_min, fmin(d.rivalry_max, score[j] - d.rivalry_decay_per_tick));
            if (v <= d.rivalry_eviction_threshold) continue;
            score[w] = v;
            if (w != j)
                for (int k = 0; k < EW; ++k) remb[w * EW + k] = remb[j * EW + k];
            ++w;
        }
#pragma unroll 1
        for (int j = w; j < n_riv; ++j) {
            score[j] = 0.0;
            for (int k = 0; k < EW; ++k) remb[j * EW + k] = 0ull;
        }
        n_riv = w;
    }
    {
        int w = 0;
#pragma unroll 1
        for (int j = 0; j < n_t; ++j) {
            const double tr = decay_value(kt[j], d.trust_decay);
            const double fa = decay_value(kf[j], d.familiarity_decay);
            const double rv = decay_value(ktr[j], d.tie_rivalry_decay);
            if (tr == 0.0 && fa == 0.0 && rv == 0.0) continue;
            kt[w] = tr;
            kf[w] = fa;
            ktr[w] = rv;
            if (w != j)
                for (int k = 0; k < EW; ++k) temb[w * EW + k] = temb[j * EW + k];
            ++w;
        }
#pragma unroll 1
        for (int j = w; j < n_t; ++j) {
            kt[j] = kf[j] = ktr[j] = 0.0;
            for (int k = 0; k < EW; ++k) temb[j * EW + k] = 0ull;
        }
        n_t = w;
    }
}
