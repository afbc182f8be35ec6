// tick_roles.cuh — entity-list form of the tick kernel with role-parallel warps
// (included by townlet_b200.cu inside its anonymous namespace; uses KParams,
// reward_agent, KpiAcc, pos_x/pos_y, kEntCache).
//
// One CTA of FOUR warps owns a group of consecutive towns whose agents fill one
// pass of up to 32 agents.  The warps take ROLES that run concurrently, so the
// latency of a pass is the longest role instead of the sum of all of them:
//
//   warp 0  "state"   one THREAD per agent: float64 need decay, faint predicate,
//                     reward with guardrails, per-town KPI partials by segmented
//                     warp shuffles, the scalar features, path hint / landmarks;
//   warp 1  "social"  one THREAD per agent: rivalry features and the social
//                     snippet (stable top-k by repeated first-argmax);
//   warp 2,3 "tiles"  the WARP serves one agent at a time (even / odd agents of
//                     the pass): template tile with 16-byte coalesced stores,
//                     then one ENTITY per lane patches its cell; local summary
//                     by REDUX over the same tests.
//
// One block barrier per pass hands the double-buffered feature rows to a
// coalesced 16-byte sweep done by all four warps.
//
// Inputs of the thread-per-agent roles arrive by cp.async: the relation rows of
// a pass (social) and the ~20 per-agent scalars (state) are requested at the end
// of the previous pass, after that pass's stores to the same arrays, so their
// global-memory latency overlaps the barrier wait and the sweep.  They are still
// re-read from global memory every pass; nothing is kept across ticks.
//
// Template parameters: VARIANT (hybrid / full / compact), WT (compile-time
// window for the common sizes, 0 = any odd window up to 21), WORLD (the launch
// runs the world-dynamics phases — action words, ledger decay — whose state is
// rewritten inside the launch and therefore read with plain loads), MIRROR (the
// tiles role also writes the bfloat16 mirror TlObsOut.map_bf16: a second, padded
// template row and one 2-byte store beside every patch; static hybrid window only,
// every other form takes the cast pass after the launch).


#ifndef TL_TILE_WARPS
#define TL_TILE_WARPS 2
#endif
constexpr int kTileWarps = TL_TILE_WARPS;  // warps in the tiles role
constexpr int kRoleWarps = 2 + kTileWarps;
constexpr int kRoleCtas = 512 / (32 * kRoleWarps);  // resident CTAs per SM the register budget is set for (128 registers)

// (the instances that also write the policy's bfloat16 rows take 96 registers and a fifth resident CTA per SM: their batches are
// rollout-sized, several waves of CTAs per tick: +1.7 % on configs[4], +3.6 % with bfloat16-only map rows; on configs[1], whose grid
// is one wave of four CTAs per SM, the same trade measured -9 %)
template <int VARIANT, int WT, bool WORLD, bool MIRROR = false>
__global__ void __launch_bounds__(32 * kRoleWarps, MIRROR ? kRoleCtas + 1 : kRoleCtas) tick_kernel_roles(const __grid_constant__ KParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* s_feat = reinterpret_cast<float*>(smem + p.off_feat);  // [2][chunk * F + 4]
    float* s_tmpl = reinterpret_cast<float*>(smem + p.off_map);
    KpiAcc* s_kpi_all = reinterpret_cast<KpiAcc*>(smem + p.off_meta);  // [2][towns_per_cta], by tick parity
    KpiRaw* s_kpi_raw = reinterpret_cast<KpiRaw*>(smem + p.off_kpi_raw);  // [2][32], by tick parity (KPI hand-off)

    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int T = p.b.n_towns;
    // the first n_big_ctas CTAs own towns_per_cta consecutive towns, the others one fewer
    const int cta = blockIdx.x;
    const int t0 = cta * p.towns_per_cta - max(0, cta - p.n_big_ctas);
    const int n_towns = min(p.towns_per_cta - (cta >= p.n_big_ctas ? 1 : 0), T - t0);
    if (n_towns <= 0) return;

    const bool do_decay = p.phases & TL_PHASE_DECAY;
    const bool do_reward = p.phases & TL_PHASE_REWARD;
    const bool do_obs = p.phases & TL_PHASE_OBS;
    const bool do_advance = p.phases & TL_PHASE_ADVANCE;
    const bool do_state = do_decay || do_reward;
    const bool do_kpi = do_state && p.ro.kpi != nullptr;
    // WORLD: the launch runs world-dynamics phases (they rewrite pos / ent / ledger rows: plain loads)
    const bool do_actions = WORLD && (p.phases & TL_PHASE_ACTIONS);
    const bool do_social_decay = WORLD && (p.phases & TL_PHASE_SOCIAL_DECAY);
    const bool do_reservations = WORLD && (p.phases & TL_PHASE_RESERVATIONS);

    const TlObsParams& o = p.obs;
    constexpr bool kStatic = WT > 0;
    static_assert(!MIRROR || (kStatic && VARIANT == TL_VARIANT_HYBRID), "the bfloat16 mirror is built for the static hybrid window");
    constexpr int kChannels = VARIANT == TL_VARIANT_HYBRID ? 4 : (VARIANT == TL_VARIANT_FULL ? 6 : 5);
    constexpr int kMapElems = kStatic ? kChannels * WT * WT : 4;
    const int W = kStatic ? WT : o.window;
    const int R = W >> 1;
    const int W2 = W * W;
    const int F = o.n_features;
    const int n_ct = kStatic ? 0 : o.n_ctypes;
    const int map_elems = kStatic ? kMapElems : o.n_channels * W2;
    const int chunk_cap = p.chunk;  // agents per pass (<= 32)
    const int centre_tile = R * W + R;
    const int walk_off = (4 + n_ct) * W2;  // compact: first float of the walkable channel
    const int feat_buf_floats = (chunk_cap * F + 4 + 3) & ~3;

    // ---- tile templates, one per reachable 16-byte misalignment of a global tile ----------
    if (do_obs && p.tmpl_mask) {
#pragma unroll 1
        for (int m = 0; m < 4; ++m) {
            if (!((p.tmpl_mask >> m) & 1)) continue;
            float* t = s_tmpl + p.tmpl_off[m] + m;
#pragma unroll 1
            for (int k = tid; k < map_elems; k += 32 * kRoleWarps) {
                float v = (k == centre_tile) ? 1.0f : 0.0f;  // self channel (map.py:69 / :185-186)
                if (VARIANT == TL_VARIANT_FULL && k >= 4 * W2) v = __ldg(o.path_lut + (k - 4 * W2));  // map.py:115-125
                if (VARIANT == TL_VARIANT_COMPACT && k >= walk_off) v = (k - walk_off == centre_tile) ? 0.0f : 1.0f;  // map.py:231-241
                t[k] = v;
            }
        }
    }

    if (do_obs) {  // shared copy of the byte -> float32 table of the id embedding
        float* elut = reinterpret_cast<float*>(smem + p.off_social) +
                      (size_t)chunk_cap * p.b.max_edges * (8 + 4 * p.b.embed_words);
        for (int k = tid; k < 256; k += 32 * kRoleWarps) elut[k] = __ldg(o.embed_lut + k);
    }

    // every warp: lane s holds the offsets of town t0 + s (s <= n_towns)
    int my_aoff = 0, my_eoff = 0, my_tick = 0;
    if (lane <= n_towns) {
        my_aoff = p.b.town_agent_off[t0 + lane];
        my_eoff = p.b.town_ent_off[t0 + lane];
    }
    if (lane < n_towns) my_tick = p.b.town_tick[t0 + lane];
    const int a_begin = __shfl_sync(FULL, my_aoff, 0);
    const int a_end = __shfl_sync(FULL, my_aoff, n_towns);
    const int my_aoff_next = __shfl_down_sync(FULL, my_aoff, 1);  // lane s < n_towns: end of town t0 + s
    const int n_chunks = (a_end - a_begin + chunk_cap - 1) / chunk_cap;
    // Small single-pass CTAs leave the tile warps waiting for the state role: there the state role only deposits each
    // agent's KPI terms in shared memory and the KPI-output warp sums them per town after the pass barrier
    // (one sequential float64 sum per town instead of a segmented shuffle reduction on the critical path).
    const bool kpi_handoff = p.kpi_out_warp != 0 && n_chunks == 1;

    // tiles role: entity words of the town being served (lane l holds entities cached_eb + l + 32 k)
    int cached_eb = -1;
    uint32_t wc[kEntCache] = {0u, 0u, 0u, 0u};
    int pass = 0;
    int slot = 0, town_abeg = 0, town_aend = 0, town_ebeg = 0, town_eend = 0, kpi_steps = 32;
    // social role: rows and edge counts of the coming pass were requested at the end of the previous one
    bool rows_in_flight = false, state_in_flight = false;
    int pre_n_riv = 0, pre_n_t = 0;
    __syncthreads();

    TL_TIMING(long long t_role = 0, t_rest = 0; const long long t_begin = clock64();)
#pragma unroll 1
    for (int it = 0; it < p.n_ticks; ++it) {
        cached_eb = -1;  // every tick re-reads the entity words of its towns
        // world/core/context.py:253-257: actions land before any system or observer runs.  No barrier: the state
        // role writes its agent's move back, and every reader of pos / ent re-derives it from the action words.
        const uint32_t* act = do_actions ? p.dyn.actions + (size_t)it * p.dyn.actions_tick_stride : nullptr;
        if (do_reservations) {
            // queues.step runs before any observer (world/systems/__init__.py order): the state warp rewrites the
            // reservation slots and the holders' flags of the CTA's towns, one barrier hands them to every role
            if (warp == 0) {
#pragma unroll 1
                for (int s = 0; s < n_towns; ++s)
                    refresh_reservations_town(p, __shfl_sync(FULL, my_aoff, s), __shfl_sync(FULL, my_aoff, s + 1),
                                              __shfl_sync(FULL, my_eoff, s), __shfl_sync(FULL, my_eoff, s + 1), lane);
            }
            __syncthreads();
        }
        // KPI partials of this tick; the other parity is being written out by a tile warp (see below)
        KpiAcc* s_kpi = s_kpi_all + (it & 1) * p.towns_per_cta;
        if (warp == 0) {
            if (do_advance && lane < n_towns) {  // core/sim_loop.py:635,659
                my_tick += 1;
                p.b.town_tick[t0 + lane] = my_tick;
            }
            if (do_kpi && !kpi_handoff && lane < n_towns) {
                KpiAcc z;
                z.rsum = z.rsq = z.hsum = z.late = z.wallet = 0.0;
                z.hmin = 1e300;
                z.starving = z.term = z.count = 0;
                s_kpi[lane] = z;
            }
            __syncwarp();
        }

#pragma unroll 1
        for (int chunk = 0; chunk < n_chunks; ++chunk, ++pass) {
            const int c_begin = a_begin + chunk * chunk_cap;
            const int c_n = min(chunk_cap, a_end - c_begin);
            float* feat_g0 = do_obs ? p.oo.features + (size_t)it * p.oo.features_tick_stride + (size_t)c_begin * F : nullptr;
            const int feat_mis = do_obs ? (int)((reinterpret_cast<uintptr_t>(feat_g0) >> 2) & 3) : 0;
            // double buffered rows; each mirrors the 16-byte misalignment of the global rows
            float* feat_s0 = s_feat + (pass & 1) * feat_buf_floats + feat_mis;

            TL_TIMING(const long long tt0 = clock64();)
            const bool active = lane < c_n;
            const int a = c_begin + (active ? lane : 0);
            // town of this lane's agent; with a single pass per tick the answer is the same every tick
            if (n_chunks > 1 || it == 0) {
                slot = 0;
#pragma unroll 1
                for (int s = 1; s < n_towns; ++s) slot += (a >= __shfl_sync(FULL, my_aoff, s)) ? 1 : 0;
                town_abeg = __shfl_sync(FULL, my_aoff, slot);
                town_aend = __shfl_sync(FULL, my_aoff, slot + 1);
                town_ebeg = __shfl_sync(FULL, my_eoff, slot);
                town_eend = __shfl_sync(FULL, my_eoff, slot + 1);
                // segmented KPI reduction needs offsets below the longest town segment of the pass only
                const int seg_len = __reduce_max_sync(FULL, active ? min(town_aend, c_begin + c_n) - max(town_abeg, c_begin) : 1);
                kpi_steps = 1;
                while (kpi_steps < seg_len) kpi_steps <<= 1;
            }

            if (warp == 0) {
                // =================== role "state": one thread per agent ==================
                const int tick = __shfl_sync(FULL, my_tick, slot);
                // the inputs of this pass were requested at the end of the previous one (same lane = same column)
                unsigned char* s_state = smem + p.off_state;
                TL_MARK_BEGIN();
                if (!state_in_flight && active) prefetch_state_in(p, s_state, a, lane);
                cp_async_wait_all();
                TL_MARK(0);
                StateRoleOut so;
                if (active) {
                    StateIn sin = staged_state_in(p, s_state, lane);
                    if (do_reservations) sin.fl = p.b.flags[a];  // the staged word predates this tick's refresh
                    state_role_agent<WORLD>(p, it, a, tick, town_abeg, town_aend, town_ebeg, town_eend,
                                            do_actions ? __ldg(act + a) : 0u, sin, do_obs ? feat_s0 + lane * F : nullptr, so);
                }
                TL_TIMING(tl_mark_t = clock64();)
                const double need0 = so.need0, reward_total = so.reward_total, kpi_wallet = so.kpi_wallet;
                const bool terminated = so.terminated;
                const int kpi_late = so.kpi_late;
                // ---- per-town KPIs: hand-off, or segmented warp-shuffle reduction ----
                if (do_kpi && kpi_handoff) {
                    if (active) {
                        KpiRaw r;
                        r.reward = reward_total, r.hunger = need0, r.wallet = kpi_wallet, r.late = kpi_late;
                        r.flags = (need0 <= p.rew.starvation_threshold ? 1 : 0) | (terminated ? 2 : 0);
                        s_kpi_raw[(it & 1) * 32 + lane] = r;
                    }
                } else if (do_kpi) {
                    const int seg = active ? slot : -1 - lane;  // inactive lanes are their own segment
                    double v_r = reward_total, v_q = reward_total * reward_total, v_h = active ? need0 : 0.0;
                    double v_m = active ? need0 : 1e300, v_w = kpi_wallet;
                    int v_l = kpi_late;
                    // starving | terminated | count packed in 10-bit fields
                    int v_c = active ? ((need0 <= p.rew.starvation_threshold ? 1 : 0) | (terminated ? 1 << 10 : 0) | (1 << 20)) : 0;
#pragma unroll 1
                    for (int off = 1; off < kpi_steps; off <<= 1) {
                        const int oseg = __shfl_down_sync(FULL, seg, off);
                        const double o_r = __shfl_down_sync(FULL, v_r, off), o_q = __shfl_down_sync(FULL, v_q, off);
                        const double o_h = __shfl_down_sync(FULL, v_h, off), o_m = __shfl_down_sync(FULL, v_m, off);
                        const double o_w = __shfl_down_sync(FULL, v_w, off);
                        const int o_l = __shfl_down_sync(FULL, v_l, off), o_c = __shfl_down_sync(FULL, v_c, off);
                        if (lane + off < 32 && oseg == seg) {
                            v_r += o_r;
                            v_q += o_q;
                            v_h += o_h;
                            v_m = fmin(v_m, o_m);
                            v_w += o_w;
                            v_l += o_l;
                            v_c += o_c;
                        }
                    }
                    const int prev_seg = __shfl_up_sync(FULL, seg, 1);
                    if (active && (lane == 0 || prev_seg != seg)) {  // segment head; passes arrive in order
                        KpiAcc& k = s_kpi[slot];
                        k.rsum += v_r;
                        k.rsq += v_q;
                        k.hsum += v_h;
                        k.hmin = fmin(k.hmin, v_m);
                        k.wallet += v_w;
                        k.late += (double)v_l;
                        k.starving += v_c & 1023;
                        k.term += (v_c >> 10) & 1023;
                        k.count += v_c >> 20;
                    }
                }
                TL_MARK(6);
                // The inputs of the NEXT pass start their trip now, after this pass's stores to the same
                // arrays; they are re-read from global memory every pass, only earlier.
                state_in_flight = false;
                {
                    int ncb = -1;
                    if (chunk + 1 < n_chunks) ncb = c_begin + chunk_cap;
                    else if (it + 1 < p.n_ticks) ncb = a_begin;
                    if (ncb >= 0) {
                        if (lane < min(chunk_cap, a_end - ncb)) prefetch_state_in(p, s_state, ncb + lane, lane);
                        state_in_flight = true;
                    }
                }
            } else if (warp == 1) {
                // =================== role "social": one thread per agent =================
                // The relation rows of the pass are one contiguous block per table: stage them in
                // shared memory with coalesced loads (one round trip), then select from there.
                const int K = p.b.max_edges;
                const int EW = p.b.embed_words;
                double* s_trust = reinterpret_cast<double*>(smem + p.off_social);
                double* s_fam = s_trust + chunk_cap * K;
                double* s_riv = s_fam + chunk_cap * K;
                double* s_score = s_riv + chunk_cap * K;
                uint64_t* s_temb = reinterpret_cast<uint64_t*>(s_score + chunk_cap * K);
                uint64_t* s_remb = s_temb + chunk_cap * K * EW;
                const float* s_elut = reinterpret_cast<const float*>(s_remb + chunk_cap * K * EW);  // byte / 255.0f
                const bool stage_ties = o.social_base >= 0 || do_social_decay;
                const bool stage_any = do_obs || do_social_decay;
                const size_t r0 = (size_t)c_begin * K;
                const int n_rows = c_n * K;
                const int rbase = lane * K;  // row of this agent inside the staged block
                // Asynchronous global -> shared copies of the relation rows (and register loads of the edge
                // counts) of the pass that starts at agent `cb`: every copy is in flight at once.
                auto prefetch_rows = [&](int cb, int cn) {
                    const size_t q0 = (size_t)cb * K;
                    const int nr = cn * K;
#pragma unroll 2
                    for (int i = lane; i < nr; i += 32) {
                        cp_async8(s_score + i, p.b.riv_score + q0 + i);
                        if (stage_ties) {
                            cp_async8(s_trust + i, p.b.tie_trust + q0 + i);
                            cp_async8(s_fam + i, p.b.tie_fam + q0 + i);
                            cp_async8(s_riv + i, p.b.tie_riv + q0 + i);
                        }
                    }
                    if (stage_ties) {
#pragma unroll 2
                        for (int i = lane; i < nr * EW; i += 32) {
                            cp_async8(s_temb + i, p.b.tie_embed + q0 * EW + i);
                            cp_async8(s_remb + i, p.b.riv_embed + q0 * EW + i);
                        }
                    }
                    pre_n_riv = pre_n_t = 0;
                    if (lane < cn) {  // raw counts: nothing may wait for these loads before the next pass uses them
                        pre_n_riv = ld_state<WORLD>(p.b.riv_count + cb + lane);
                        pre_n_t = ld_state<WORLD>(p.b.tie_count + cb + lane);
                    }
                };
                if (stage_any && !rows_in_flight) prefetch_rows(c_begin, c_n);  // first pass of the launch
                cp_async_wait_all();
                __syncwarp();
                int n_riv = min(pre_n_riv, K), n_t = min(pre_n_t, K);  // counts are clamped to the row stride
                if (do_social_decay) {
                    // ledger decay in the staged rows, then the block goes back with coalesced stores
                    const int n_riv0 = n_riv, n_t0 = n_t;
                    if (active) {
                        social_decay_agent(p, n_riv, n_t, s_trust + rbase, s_fam + rbase, s_riv + rbase, s_score + rbase,
                                           s_temb + rbase * EW, s_remb + rbase * EW);
                        if (n_riv != n_riv0) p.b.riv_count[a] = (uint8_t)n_riv;
                        if (n_t != n_t0) p.b.tie_count[a] = (uint8_t)n_t;
                    }
                    const bool riv_moved = __any_sync(FULL, n_riv != n_riv0), tie_moved = __any_sync(FULL, n_t != n_t0);
                    // a table goes back when its values decay or an entry moved; an untouched table stays as it is
                    const bool w_score = p.dyn.rivalry_decay_per_tick != 0.0 || riv_moved;
                    const bool w_trust = p.dyn.trust_decay > 0.0 || tie_moved, w_fam = p.dyn.familiarity_decay > 0.0 || tie_moved;
                    const bool w_riv = p.dyn.tie_rivalry_decay > 0.0 || tie_moved;
                    __syncwarp();
#pragma unroll 2
                    for (int i = lane; i < n_rows; i += 32) {
                        if (w_score) p.b.riv_score[r0 + i] = s_score[i];
                        if (w_trust) p.b.tie_trust[r0 + i] = s_trust[i];
                        if (w_fam) p.b.tie_fam[r0 + i] = s_fam[i];
                        if (w_riv) p.b.tie_riv[r0 + i] = s_riv[i];
                    }
                    if (riv_moved || tie_moved) {
#pragma unroll 2
                        for (int i = lane; i < n_rows * EW; i += 32) {
                            if (tie_moved) p.b.tie_embed[r0 * EW + i] = s_temb[i];
                            if (riv_moved) p.b.riv_embed[r0 * EW + i] = s_remb[i];
                        }
                    }
                }
                if (active && do_obs) {
                    float* f = feat_s0 + lane * F;
                    int filled = 0, source = 0;
                    social_role_agent(p, n_riv, n_t, s_trust + rbase, s_fam + rbase, s_riv + rbase, s_score + rbase,
                                      s_temb + rbase * EW, s_remb + rbase * EW, s_elut, f, filled, source);
                    if (p.oo.summary) {
                        int* srow_i = p.oo.summary + (size_t)it * p.oo.summary_tick_stride + (size_t)a * TL_N_SUMMARY;
                        *reinterpret_cast<int4*>(srow_i + 4) = make_int4(filled, source, 0, 0);
                    }
                }
                // The rows of the NEXT pass (next block of this tick, or the first block of the next tick) start
                // their trip now and land while this warp waits at the barrier and sweeps: they are re-read from
                // global memory every pass, only earlier.
                rows_in_flight = false;
                if (stage_any) {
                    int ncb = -1;
                    if (chunk + 1 < n_chunks) ncb = c_begin + chunk_cap;
                    else if (it + 1 < p.n_ticks) ncb = a_begin;
                    if (ncb >= 0) {
                        __syncwarp();  // every lane is done with the staged rows (and their write-back)
                        prefetch_rows(ncb, min(chunk_cap, a_end - ncb));
                        rows_in_flight = true;
                    }
                }
            } else if (do_obs) {
                // =================== role "tiles": the warp serves one agent at a time ====
                // Lane i keeps the integer summary of agent i; the look-ups run thread-per-agent afterwards.
                const int parity = warp - 2;
                int32_t pw = active ? ld_state<WORLD>(p.b.pos + a) : 0;
                if (do_actions && active) pw = moved_pos(pw, __ldg(act + a));
                int my_others = 0, my_objs = 0, my_res = 0, my_d2 = 0;
                float* mrow_g = p.oo.map + (size_t)it * p.oo.map_tick_stride + (size_t)(c_begin + parity) * map_elems;
                // MIRROR: the same tile as bfloat16 (0 -> 0x0000, 1 -> 0x3f80: the cast is exact), rows padded with zeros
                const int brow = MIRROR ? p.oo.map_bf16_row : 0;
                uint16_t* brow_g = (MIRROR && p.oo.map_bf16)
                                       ? p.oo.map_bf16 + (size_t)it * p.oo.map_bf16_tick_stride + (size_t)(c_begin + parity) * brow
                                       : nullptr;
                constexpr bool kAlignedRows = kStatic && (kMapElems & 3) == 0;
                const bool f32_map = !MIRROR || p.oo.map != nullptr;  // (MIRROR: the float32 tile is optional)
                const int map_mis0 = (int)((reinterpret_cast<uintptr_t>(mrow_g) >> 2) & 3);
#pragma unroll 1
                for (int i = parity; i < c_n; i += kTileWarps, mrow_g += kTileWarps * map_elems, brow_g += (MIRROR && brow_g) ? kTileWarps * brow : 0) {
                    const int32_t pwi = __shfl_sync(FULL, pw, i);
                    const int cxi = pos_x(pwi), cyi = pos_y(pwi);
                    const int eb = __shfl_sync(FULL, town_ebeg, i), ee = __shfl_sync(FULL, town_eend, i);
                    const int ab = __shfl_sync(FULL, town_abeg, i);
                    const int self_e = eb + (c_begin + i - ab);
                    // entity word e of the served town; the first (town's agent count) words are its agents
                    const int n_town_agents = do_actions ? __shfl_sync(FULL, town_aend, i) - ab : 0;
                    auto ent_word = [&](int e) -> uint32_t {
                        if (e >= ee) return 0u;
                        const uint32_t w = ld_state<WORLD>(p.b.ent + e);
                        return (do_actions && e - eb < n_town_agents) ? moved_ent(w, __ldg(act + ab + (e - eb))) : w;
                    };
                    if (eb != cached_eb) {  // the warp moved on to the next town: fetch its entity words
                        cached_eb = eb;
#pragma unroll
                        for (int k = 0; k < kEntCache; ++k) wc[k] = ent_word(eb + 32 * k + lane);
                    }
                    // ---- template tile: 16-byte vectorised coalesced stores -------------
                    // rows of a multiple of four floats all sit at the misalignment of the first one
                    const int map_mis = kAlignedRows ? map_mis0 : (int)((reinterpret_cast<uintptr_t>(mrow_g) >> 2) & 3);
                    const int head = (4 - map_mis) & 3;
                    const int nvec = (map_elems - head) >> 2;
                    float4* g4 = reinterpret_cast<float4*>(mrow_g + head);
                    if (MIRROR && !f32_map) {
                        // the map is delivered as its bfloat16 row only (TlObsOut.map == NULL)
                    } else if (kStatic && VARIANT == TL_VARIANT_HYBRID && map_mis == 0) {
                        // all zeros except the self cell: flat float centre_tile (60 for W = 11) sits in
                        // float4 number 15, component 0, which lane 15 writes in its first store
                        static_assert(!kStatic || VARIANT != TL_VARIANT_HYBRID || (((WT / 2) * WT + WT / 2) & 3) == 0, "self cell component");
                        const float self0 = (lane == (centre_tile >> 2)) ? 1.0f : 0.0f;
#pragma unroll
                        for (int k = 0; k < (kMapElems / 4 + 31) / 32; ++k) {
                            const int v = lane + 32 * k;
                            if (v < kMapElems / 4) {
                                const float4 z = make_float4(k == 0 ? self0 : 0.0f, 0.0f, 0.0f, 0.0f);
                                if (MIRROR) __stcs(g4 + v, z);  // streamed: the bfloat16 rows should be what stays in L2 for the policy step
                                else g4[v] = z;
                            }
                        }
                    } else {
                        const float* t = s_tmpl + p.tmpl_off[map_mis] + map_mis;
                        const float4* s4 = reinterpret_cast<const float4*>(t + head);
                        if (kStatic) {
#pragma unroll
                            for (int k = 0; k < (kMapElems / 4 + 31) / 32; ++k) {
                                const int v = lane + 32 * k;
                                if (v < nvec) g4[v] = s4[v];
                            }
                        } else {
#pragma unroll 1
                            for (int v = lane; v < nvec; v += 32) g4[v] = s4[v];
                        }
                        if (lane < head) mrow_g[lane] = t[lane];
                        const int tail0 = head + (nvec << 2);
                        if (lane < map_elems - tail0) mrow_g[tail0 + lane] = t[tail0 + lane];
                    }
                    if (MIRROR && brow_g) {  // padded bfloat16 row: zeros and the self cell, 16-byte stores of 8 elements
                        uint4* b4 = reinterpret_cast<uint4*>(brow_g);
                        const int nb = p.oo.map_bf16_width >> 3;
                        const unsigned one = 0x3f80u << (16 * (centre_tile & 1));
                        const int cw = (centre_tile & 7) >> 1;
#pragma unroll 1
                        for (int v = lane; v < nb; v += 32) {
                            const unsigned self = v == (centre_tile >> 3) ? one : 0u;
                            b4[v] = make_uint4(cw == 0 ? self : 0u, cw == 1 ? self : 0u, cw == 2 ? self : 0u, cw == 3 ? self : 0u);
                        }
                    }
                    __syncwarp();  // orders the template stores before the patches of other lanes

                    // ---- patches: one entity per lane (cache.py:16-41 + map.py:80-125) ----
                    int n_ag = 0, n_obj = 0, n_res = 0, centre_has = 0;
                    unsigned nearest = 0xffffffffu;
                    auto serve = [&](uint32_t w, int e, bool may_skip) {
                        const bool valid = e < ee;
                        const int dx = TL_ENT_X(w) - cxi, dy = TL_ENT_Y(w) - cyi;
                        const bool inwin = valid && (unsigned)(dx + R) <= (unsigned)(2 * R) && (unsigned)(dy + R) <= (unsigned)(2 * R);
                        // later passes of a large town rarely touch the window: one vote skips the rest
                        if (may_skip && !__any_sync(FULL, inwin)) return;
                        const int tile = (dy + R) * W + (dx + R);
                        const unsigned kind = TL_ENT_KIND(w);
                        const bool is_agent = inwin && kind == TL_KIND_AGENT;
                        const bool is_obj = inwin && kind == TL_KIND_OBJECT && TL_ENT_TILE(w);
                        const bool is_res = inwin && kind == TL_KIND_RESERVED;
                        // LocalSummary (map.py:88-103): one vote per count, so every lane holds the warp's totals
                        n_ag += __popc(__ballot_sync(FULL, is_agent));
                        n_obj += __popc(__ballot_sync(FULL, is_obj));
                        n_res += __popc(__ballot_sync(FULL, is_res));
                        const int d2 = dx * dx + dy * dy;
                        centre_has |= __ballot_sync(FULL, is_agent && d2 == 0) != 0u;
                        if (is_agent && d2 != 0) nearest = min(nearest, (unsigned)d2);
                        if (VARIANT == TL_VARIANT_COMPACT) {
                            const bool is_other = is_agent && e != self_e;  // map.py:198
                            const bool norm = o.normalize_counts != 0;
                            if (is_other) {
                                if (norm) mrow_g[W2 + tile] = 1.0f;
                                else atomicAdd(mrow_g + W2 + tile, 1.0f);
                            }
                            if (is_obj) {
                                if (norm) mrow_g[2 * W2 + tile] = 1.0f;
                                else atomicAdd(mrow_g + 2 * W2 + tile, 1.0f);
                                const int ct = (int)TL_ENT_CTYPE(w);
                                if (ct > 0 && ct <= n_ct) {  // map.py:214-229
                                    if (norm) mrow_g[(3 + ct) * W2 + tile] = 1.0f;
                                    else atomicAdd(mrow_g + (3 + ct) * W2 + tile, 1.0f);
                                }
                            }
                            if (is_res) mrow_g[3 * W2 + tile] = 1.0f;
                            if (is_other || is_obj || is_res) mrow_g[walk_off + tile] = 0.0f;
                        } else {
                            if (!MIRROR || f32_map) {
                                if (is_agent) mrow_g[W2 + tile] = 1.0f;  // includes the observer (map.py:109-110)
                                if (is_obj) mrow_g[2 * W2 + tile] = 1.0f;
                                if (is_res) mrow_g[3 * W2 + tile] = 1.0f;
                            }
                            if (MIRROR && brow_g) {
                                if (is_agent) brow_g[W2 + tile] = 0x3f80;
                                if (is_obj) brow_g[2 * W2 + tile] = 0x3f80;
                                if (is_res) brow_g[3 * W2 + tile] = 0x3f80;
                            }
                        }
                    };
                    // towns of up to 32 entities (the common case) need one pass and one branch
                    serve(wc[0], eb + lane, false);
                    if (ee - eb > 32) {
                        serve(wc[1], eb + 32 + lane, true);
                        if (ee - eb > 64) {
                            serve(wc[2], eb + 64 + lane, true);
                            if (ee - eb > 96) {
                                serve(wc[3], eb + 96 + lane, true);
#pragma unroll 1
                                for (int e0 = eb + 32 * kEntCache; e0 < ee; e0 += 32) {
                                    serve(ent_word(e0 + lane), e0 + lane, true);
                                }
                            }
                        }
                    }
                    // the counts are warp totals already (votes); the nearest distance needs one reduction
                    const int centre = centre_has;
                    nearest = __reduce_min_sync(FULL, nearest);
                    if (lane == i) {
                        my_others = n_ag - centre;  // self excluded on the centre tile only
                        my_objs = n_obj;
                        my_res = n_res;
                        my_d2 = nearest == 0xffffffffu ? 0 : (int)nearest;
                    }
                }
                // local summary features, thread-per-agent (features.py:289-330 via the host LUTs)
                if (active && (lane % kTileWarps) == parity) {
                    float* frow = feat_s0 + lane * F;
                    frow[TL_FEAT_AGENT_RATIO] = __ldg(o.agent_ratio_lut + min(my_others, W2 - 1));
                    frow[TL_FEAT_OBJECT_RATIO] = __ldg(o.tile_ratio_lut + min(my_objs, W2));
                    frow[TL_FEAT_RESERVED_RATIO] = __ldg(o.tile_ratio_lut + min(my_res, W2));
                    frow[TL_FEAT_NEAREST] = __ldg(o.nearest_lut + my_d2);
                    if (p.oo.summary) {
                        int* srow_i = p.oo.summary + (size_t)it * p.oo.summary_tick_stride + (size_t)a * TL_N_SUMMARY;
                        *reinterpret_cast<int4*>(srow_i) = make_int4(my_others, my_objs, my_res, my_d2);
                    }
                }
            }
            TL_TIMING(const long long tt1 = clock64();)
            if (!do_obs) continue;  // CTA-uniform
            __syncthreads();  // the rows of this pass are complete
            {  // feature rows of the whole pass: one coalesced 16-byte sweep by all four warps
                const int total = c_n * F;
                const int head = min((4 - feat_mis) & 3, total);
                if (tid < head) feat_g0[tid] = feat_s0[tid];
                const int nvec = (total - head) >> 2;
                float4* g4 = reinterpret_cast<float4*>(feat_g0 + head);
                const float4* s4 = reinterpret_cast<const float4*>(feat_s0 + head);
#pragma unroll 4
                for (int v = tid; v < nvec; v += 32 * kRoleWarps) {
                    if (MIRROR) __stcs(g4 + v, s4[v]);
                    else g4[v] = s4[v];
                }
                const int tail0 = head + (nvec << 2);
                if (tid < total - tail0) feat_g0[tail0 + tid] = feat_s0[tail0 + tid];
            }
            if (MIRROR && p.oo.features_bf16) {  // the same rows as padded bfloat16 (round to nearest even), 8 elements per store
                const int frow = p.oo.features_bf16_row, per_row = p.oo.features_bf16_width >> 3;
                uint16_t* b0 = p.oo.features_bf16 + (size_t)it * p.oo.features_bf16_tick_stride + (size_t)c_begin * frow;
#pragma unroll 1
                for (int v = tid; v < c_n * per_row; v += 32 * kRoleWarps) {
                    const int ag = v / per_row, e0 = (v - ag * per_row) << 3;
                    const float* src = feat_s0 + ag * F;
                    unsigned w[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int e = e0 + 2 * k;
                        const __nv_bfloat162 pair = __floats2bfloat162_rn(e < F ? src[e] : 0.0f, e + 1 < F ? src[e + 1] : 0.0f);
                        w[k] = *reinterpret_cast<const unsigned*>(&pair);
                    }
                    *reinterpret_cast<uint4*>(b0 + (size_t)ag * frow + e0) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            TL_TIMING(t_role += tt1 - tt0; t_rest += clock64() - tt1;)
            // no second barrier: the next pass fills the other buffer, and the barrier of
            // that pass orders this sweep before the buffer is written again
        }

        // Small passes leave the tile warps idle while the state role is the long pole; the host then lets a tile warp
        // write the KPI rows out (after the barrier of the tick's last pass), while the state warp already starts the
        // next tick on the other parity.
        if (do_kpi && warp == (do_obs ? p.kpi_out_warp : 0)) {
            __syncwarp();
            if (lane < n_towns) {
                KpiAcc k;
                if (kpi_handoff) {  // this lane's town: sum its agents' terms (they all sit in the tick's single pass)
                    k.rsum = k.rsq = k.hsum = k.late = k.wallet = 0.0;
                    k.hmin = 1e300;
                    k.starving = k.term = k.count = 0;
                    const KpiRaw* raw = s_kpi_raw + (it & 1) * 32;
                    const int i0 = my_aoff - a_begin, i1 = my_aoff_next - a_begin;
#pragma unroll 1
                    for (int i = i0; i < i1; ++i) {
                        const KpiRaw r = raw[i];
                        k.rsum += r.reward;
                        k.rsq += r.reward * r.reward;
                        k.hsum += r.hunger;
                        k.hmin = fmin(k.hmin, r.hunger);
                        k.wallet += r.wallet;
                        k.late += (double)r.late;
                        k.starving += r.flags & 1;
                        k.term += (r.flags >> 1) & 1;
                        k.count += 1;
                    }
                } else {
                    k = s_kpi[lane];
                }
                float* dst = p.ro.kpi + (size_t)it * p.ro.kpi_tick_stride + (size_t)(t0 + lane) * TL_N_KPI;
                const double inv = k.count ? 1.0 / (double)k.count : 0.0;  // one division for the three means
                dst[TL_KPI_REWARD_SUM] = (float)k.rsum;
                dst[TL_KPI_REWARD_SUMSQ] = (float)k.rsq;
                dst[TL_KPI_HUNGER_MIN] = k.count ? (float)k.hmin : 0.0f;
                dst[TL_KPI_HUNGER_MEAN] = (float)(k.hsum * inv);
                dst[TL_KPI_STARVING] = (float)k.starving;
                dst[TL_KPI_TERMINATED] = (float)k.term;
                dst[TL_KPI_LATENESS_MEAN] = (float)(k.late * inv);
                dst[TL_KPI_WALLET_MEAN] = (float)(k.wallet * inv);
            }
            __syncwarp();
        }
    }
    TL_TIMING(if (blockIdx.x == 0 && tid == 0 && p.n_ticks > 1) {
        printf("TL_ROLE_TIMING state parts per tick: wait-inputs %lld decay+faint+tick %lld reward %lld scalar-features %lld landmark-scan %lld "
               "path-hint %lld kpi %lld\n", g_state_parts[0] / p.n_ticks, g_state_parts[1] / p.n_ticks, g_state_parts[2] / p.n_ticks,
               g_state_parts[3] / p.n_ticks, g_state_parts[4] / p.n_ticks, g_state_parts[5] / p.n_ticks, g_state_parts[6] / p.n_ticks);
        for (int k = 0; k < 8; ++k) g_state_parts[k] = 0;
    })
    TL_TIMING(if ((blockIdx.x == 0 || blockIdx.x == gridDim.x / 2) && lane == 0 && p.n_ticks > 1)
                  printf("TL_ROLE_TIMING cta %d warp %d (%s) role %lld barrier+sweep %lld total %lld cycles per tick\n", blockIdx.x, warp,
                         warp == 0 ? "state" : (warp == 1 ? "social" : "tiles"), t_role / p.n_ticks, t_rest / p.n_ticks,
                         (clock64() - t_begin) / p.n_ticks);)
}
