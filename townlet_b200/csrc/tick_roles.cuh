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
// coalesced 16-byte sweep done by all four warps.  See tick_ent.cuh for the
// single-warp form this was derived from (kept for very small launches).

#ifndef TL_TILE_WARPS
#define TL_TILE_WARPS 2
#endif
constexpr int kTileWarps = TL_TILE_WARPS;  // warps in the tiles role
constexpr int kRoleWarps = 2 + kTileWarps;

template <int VARIANT, int WT>
__global__ void __launch_bounds__(32 * kRoleWarps, 512 / (32 * kRoleWarps)) tick_kernel_roles(const __grid_constant__ KParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* s_feat = reinterpret_cast<float*>(smem + p.off_feat);  // [2][chunk * F + 4]
    float* s_tmpl = reinterpret_cast<float*>(smem + p.off_map);
    KpiAcc* s_kpi = reinterpret_cast<KpiAcc*>(smem + p.off_meta);

    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int T = p.b.n_towns;
    const int N = p.b.n_agents;
    const int t0 = blockIdx.x * p.towns_per_cta;
    const int n_towns = min(p.towns_per_cta, T - t0);
    if (n_towns <= 0) return;

    const bool do_decay = p.phases & TL_PHASE_DECAY;
    const bool do_reward = p.phases & TL_PHASE_REWARD;
    const bool do_obs = p.phases & TL_PHASE_OBS;
    const bool do_advance = p.phases & TL_PHASE_ADVANCE;
    const bool do_state = do_decay || do_reward;
    const bool do_kpi = do_state && p.ro.kpi != nullptr;

    const TlObsParams& o = p.obs;
    constexpr bool kStatic = WT > 0;
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
    const int n_chunks = (a_end - a_begin + chunk_cap - 1) / chunk_cap;

    // tiles role: entity words of the town being served (lane l holds entities cached_eb + l + 32 k)
    int cached_eb = -1;
    uint32_t wc[kEntCache] = {0u, 0u, 0u, 0u};
    int pass = 0;
    int slot = 0, town_abeg = 0, town_aend = 0, town_ebeg = 0, town_eend = 0, kpi_steps = 32;
    __syncthreads();

#pragma unroll 1
    for (int it = 0; it < p.n_ticks; ++it) {
        cached_eb = -1;  // every tick re-reads the entity words of its towns
        if (warp == 0) {
            if (do_advance && lane < n_towns) {  // core/sim_loop.py:635,659
                my_tick += 1;
                p.b.town_tick[t0 + lane] = my_tick;
            }
            if (do_kpi && lane < n_towns) {
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
                double need0 = 0.0, reward_total = 0.0, kpi_wallet = 0.0;
                bool terminated = false;
                int kpi_late = 0;
                if (active) {
                    // the agent's scalar inputs are loaded up front so the loads overlap
                    double need[3];
                    need[0] = p.b.needs[a];
                    need[1] = p.b.needs[(size_t)N + a];
                    need[2] = p.b.needs[2 * (size_t)N + a];
                    uint32_t fl = p.b.flags[a];
                    int episode_tick = p.b.episode_tick[a];
                    RewardIn rin;
                    if (do_reward) rin = load_reward_in(p, a);
                    const double wallet = __ldg(p.b.wallet + a);
                    const int lateness = __ldg(p.b.lateness + a);
                    kpi_wallet = wallet;
                    kpi_late = lateness;
                    double attendance = 0.0, wages_withheld = 0.0;
                    int duration = 0, slot_idx = 0;
                    float last_hash = 0.0f;
                    int32_t pw = 0;
                    if (do_obs) {
                        attendance = __ldg(p.b.attendance + a);
                        wages_withheld = __ldg(p.b.wages_withheld + a);
                        duration = __ldg(p.b.last_action_duration + a);
                        slot_idx = __ldg(p.b.slot + a);
                        last_hash = __ldg(p.b.last_action_hash + a);
                        pw = __ldg(p.b.pos + a);
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
                        episode_tick = (episode_tick + 1) % span;
                        p.b.episode_tick[a] = episode_tick;
                    }
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

                    if (do_obs) {  // encoders/features.py:43-404
                        float* f = feat_s0 + lane * F;
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
                        f[TL_FEAT_SLOT_NORM] = (slot_idx >= 0 && slot_idx < o.max_slots)
                                                   ? __ldg(o.slot_lut + slot_idx)
                                                   : (float)((double)slot_idx / (double)o.max_slots);
                        f[TL_FEAT_CTX_RESET] = ((fl & TL_F_CTX_RESET) || terminated) ? 1.0f : 0.0f;  // service.py:224-229
                        f[TL_FEAT_EPISODE_PROGRESS] =
                            episode_tick <= 0 ? 0.0f : __ldg(o.progress_lut + min(episode_tick, o.ticks_per_day));
                        f[TL_FEAT_LAST_HASH] = last_hash;
                        f[TL_FEAT_LAST_SUCCESS] = (fl & TL_F_LAST_SUCCESS) ? 1.0f : 0.0f;
                        f[TL_FEAT_LAST_DURATION] = (float)duration;
                        f[TL_FEAT_RESERVATION] = (fl & TL_F_RESERVATION) ? 1.0f : 0.0f;
                        f[TL_FEAT_IN_QUEUE] = (fl & TL_F_IN_QUEUE) ? 1.0f : 0.0f;
                        // nearest object per landmark type: first minimum of the squared
                        // distance in insertion order (world/observe.py:109-147)
                        const int cx = pos_x(pw), cy = pos_y(pw);
                        const bool all_types = o.landmark_base >= 0;
                        const int e_obj = town_ebeg + (town_aend - town_abeg);
                        int hd2 = -1, hdx = 0, hdy = 0;
                        if (!all_types) {
#pragma unroll 1
                            for (int e = e_obj; e < town_eend; ++e) {
                                const uint32_t w = __ldg(p.b.ent + e);
                                if (TL_ENT_KIND(w) != TL_KIND_OBJECT || !TL_ENT_LM(w) || (int)TL_ENT_LTYPE(w) != o.path_hint_ltype)
                                    continue;
                                const int dx = TL_ENT_X(w) - cx, dy = TL_ENT_Y(w) - cy;
                                const int d2 = dx * dx + dy * dy;
                                if (hd2 < 0 || d2 < hd2) {
                                    hd2 = d2;
                                    hdx = dx;
                                    hdy = dy;
                                }
                            }
                        } else {
                            int best_d2[5] = {-1, -1, -1, -1, -1};
                            int best_dx[5] = {0, 0, 0, 0, 0}, best_dy[5] = {0, 0, 0, 0, 0};
#pragma unroll 1
                            for (int e = e_obj; e < town_eend; ++e) {
                                const uint32_t w = __ldg(p.b.ent + e);
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
#pragma unroll
                            for (int l = 1; l <= 4; ++l) {  // _encode_landmarks, features.py:333-356
                                if (l == o.path_hint_ltype) {
                                    hd2 = best_d2[l];
                                    hdx = best_dx[l];
                                    hdy = best_dy[l];
                                }
                                float* dst = f + o.landmark_base + 3 * (l - 1);
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
                        }
                        {  // _encode_path_hint, features.py:251-286
                            double north = 0.0, south = 0.0, east = 0.0, west = 0.0;
                            if (hd2 > 0) {
                                const double dn = (double)(abs(hdx) + abs(hdy));
                                // max(0, -dy)/norm etc.: one of each opposite pair is exactly 0.0, so two
                                // IEEE divisions give the four values
                                const double ns = (double)abs(hdy) / dn, ew = (double)abs(hdx) / dn;
                                north = hdy < 0 ? ns : 0.0;
                                south = hdy > 0 ? ns : 0.0;
                                east = hdx > 0 ? ew : 0.0;
                                west = hdx < 0 ? ew : 0.0;
                            }
                            f[TL_FEAT_PATH_NORTH + 0] = (float)north;
                            f[TL_FEAT_PATH_NORTH + 1] = (float)south;
                            f[TL_FEAT_PATH_NORTH + 2] = (float)east;
                            f[TL_FEAT_PATH_NORTH + 3] = (float)west;
                        }
                        if (o.personality_base >= 0) {  // features.py:359-380
#pragma unroll
                            for (int k = 0; k < 3; ++k)
                                f[o.personality_base + k] = p.b.personality ? __ldg(p.b.personality + (size_t)k * N + a) : 0.0f;
                        }
                    }
                }
                // ---- per-town KPIs: segmented warp-shuffle reduction ----------------
                if (do_kpi) {
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
                if (do_obs) {
                    const size_t r0 = (size_t)c_begin * K;
                    const int n_rows = c_n * K;
#pragma unroll 2
                    for (int i = lane; i < n_rows; i += 32) {
                        const double v_score = __ldg(p.b.riv_score + r0 + i);
                        if (o.social_base >= 0) {
                            const double v_t = __ldg(p.b.tie_trust + r0 + i), v_f = __ldg(p.b.tie_fam + r0 + i);
                            const double v_r = __ldg(p.b.tie_riv + r0 + i);
                            s_trust[i] = v_t;
                            s_fam[i] = v_f;
                            s_riv[i] = v_r;
                        }
                        s_score[i] = v_score;
                    }
                    if (o.social_base >= 0) {
#pragma unroll 2
                        for (int i = lane; i < n_rows * EW; i += 32) {
                            const uint64_t v_t = __ldg(p.b.tie_embed + r0 * EW + i), v_r = __ldg(p.b.riv_embed + r0 * EW + i);
                            s_temb[i] = v_t;
                            s_remb[i] = v_r;
                        }
                    }
                }
                __syncwarp();
                if (active && do_obs) {
                    float* f = feat_s0 + lane * F;
                    const int rbase = lane * K;  // row of this agent inside the staged block
                    const int n_riv = min((int)__ldg(p.b.riv_count + a), K);  // counts are clamped to the row stride
                    const int n_t = min((int)__ldg(p.b.tie_count + a), K);
                    {  // rivalry_top(agent, max_edges): rows are pre-sorted (features.py:232-248)
                        const int nr = min(n_riv, o.rivalry_max_edges);
                        int avoid = 0;
                        double mx = 0.0;
#pragma unroll 1
                        for (int j = 0; j < nr; ++j) {
                            const double v = s_score[rbase + j];
                            if (j == 0) mx = v;
                            avoid += (v >= o.avoid_threshold) ? 1 : 0;
                        }
                        f[TL_FEAT_RIVALRY_MAX] = (float)mx;
                        f[TL_FEAT_RIVALRY_AVOID] = (float)avoid;
                    }
                    int filled = 0, source = 0;
                    if (o.social_base >= 0) {  // encoders/social.py:27-254
                        const int slot_dim = o.embed_dim + 3;
                        const int total_slots = o.top_friends + o.top_rivals;
                        float* srow = f + o.social_base;
                        const bool from_ties = n_t > 0;  // _resolve_relationships, social.py:165-209
                        const int n_rel = from_ties ? n_t : min(n_riv, o.top_rivals);
                        source = from_ties ? 1 : (n_rel > 0 ? 2 : 0);
                        const double* kt = s_trust + rbase;
                        const double* kf = s_fam + rbase;
                        const double* kr = (from_ties ? s_riv : s_score) + rbase;
                        const uint64_t* ke = (from_ties ? s_temb : s_remb) + rbase * EW;
                        // sorted(..., reverse=True)[:k] is stable: repeated FIRST argmax (social.py:137-156).
                        // In the rivalry fallback trust = familiarity = 0, so friends are the first entries.
                        unsigned chosen = 0, sel = 0;
                        const int n_friends = min(n_rel, o.top_friends);
                        const int n_rivals = min(o.top_rivals, n_rel - n_friends);
                        filled = n_friends + n_rivals;
#pragma unroll 1
                        for (int s = 0; s < filled; ++s) {
                            const bool friend_pass = s < n_friends;
                            int best = -1;
                            double bk = 0.0;
#pragma unroll 1
                            for (int j = 0; j < n_rel; ++j) {
                                if ((chosen >> j) & 1u) continue;
                                const double key = friend_pass ? (from_ties ? kt[j] + kf[j] : 0.0) : kr[j];
                                if (best < 0 || key > bk) {
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
#pragma unroll 1
                                for (int d = 0; d < slot_dim; ++d) dst[d] = 0.0f;
                                continue;
                            }
                            const int j = (sel >> (4 * s)) & 15;
                            const double trust = from_ties ? kt[j] : 0.0;
                            const double fam = from_ties ? kf[j] : 0.0;
                            const double riv = kr[j];
                            if (o.embed_dim == 8) {  // _embed_agent_id, social.py:236-242; bytes via PRMT
                                const uint64_t word = ke[j * EW];
                                const unsigned lo = (unsigned)word, hi = (unsigned)(word >> 32);
                                dst[0] = s_elut[__byte_perm(lo, 0, 0x4440)];
                                dst[1] = s_elut[__byte_perm(lo, 0, 0x4441)];
                                dst[2] = s_elut[__byte_perm(lo, 0, 0x4442)];
                                dst[3] = s_elut[__byte_perm(lo, 0, 0x4443)];
                                dst[4] = s_elut[__byte_perm(hi, 0, 0x4440)];
                                dst[5] = s_elut[__byte_perm(hi, 0, 0x4441)];
                                dst[6] = s_elut[__byte_perm(hi, 0, 0x4442)];
                                dst[7] = s_elut[__byte_perm(hi, 0, 0x4443)];
                            } else {
                                uint64_t word = 0;
#pragma unroll 1
                                for (int d = 0; d < o.embed_dim; ++d) {
                                    if ((d & 7) == 0) word = ke[j * EW + (d >> 3)];
                                    dst[d] = s_elut[(unsigned)(word >> (8 * (d & 7))) & 255u];
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
                    if (p.oo.summary) {
                        int* srow_i = p.oo.summary + (size_t)it * p.oo.summary_tick_stride + (size_t)a * TL_N_SUMMARY;
                        *reinterpret_cast<int4*>(srow_i + 4) = make_int4(filled, source, 0, 0);
                    }
                }
            } else if (do_obs) {
                // =================== role "tiles": the warp serves one agent at a time ====
                // Lane i keeps the integer summary of agent i; the look-ups run thread-per-agent afterwards.
                const int parity = warp - 2;
                const int32_t pw = active ? __ldg(p.b.pos + a) : 0;
                int my_others = 0, my_objs = 0, my_res = 0, my_d2 = 0;
                float* mrow_g = p.oo.map + (size_t)it * p.oo.map_tick_stride + (size_t)(c_begin + parity) * map_elems;
#pragma unroll 1
                for (int i = parity; i < c_n; i += kTileWarps, mrow_g += kTileWarps * map_elems) {
                    const int32_t pwi = __shfl_sync(FULL, pw, i);
                    const int cxi = pos_x(pwi), cyi = pos_y(pwi);
                    const int eb = __shfl_sync(FULL, town_ebeg, i), ee = __shfl_sync(FULL, town_eend, i);
                    const int self_e = eb + (c_begin + i - __shfl_sync(FULL, town_abeg, i));
                    if (eb != cached_eb) {  // the warp moved on to the next town: fetch its entity words
                        cached_eb = eb;
#pragma unroll
                        for (int k = 0; k < kEntCache; ++k) {
                            const int e = eb + 32 * k + lane;
                            wc[k] = e < ee ? __ldg(p.b.ent + e) : 0u;
                        }
                    }
                    // ---- template tile: 16-byte vectorised coalesced stores -------------
                    const int map_mis = (int)((reinterpret_cast<uintptr_t>(mrow_g) >> 2) & 3);
                    const int head = (4 - map_mis) & 3;
                    const int nvec = (map_elems - head) >> 2;
                    float4* g4 = reinterpret_cast<float4*>(mrow_g + head);
                    if (kStatic && VARIANT == TL_VARIANT_HYBRID && map_mis == 0) {
                        // all zeros except the self cell: flat float centre_tile (60 for W = 11) sits in
                        // float4 number 15, component 0, which lane 15 writes in its first store
                        static_assert(!kStatic || VARIANT != TL_VARIANT_HYBRID || (((WT / 2) * WT + WT / 2) & 3) == 0, "self cell component");
                        const float self0 = (lane == (centre_tile >> 2)) ? 1.0f : 0.0f;
#pragma unroll
                        for (int k = 0; k < (kMapElems / 4 + 31) / 32; ++k) {
                            const int v = lane + 32 * k;
                            if (v < kMapElems / 4) g4[v] = make_float4(k == 0 ? self0 : 0.0f, 0.0f, 0.0f, 0.0f);
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
                        // LocalSummary (map.py:88-103)
                        n_ag += is_agent ? 1 : 0;
                        n_obj += is_obj ? 1 : 0;
                        n_res += is_res ? 1 : 0;
                        const int d2 = dx * dx + dy * dy;
                        if (is_agent) {
                            if (d2 == 0) centre_has = 1;
                            else nearest = min(nearest, (unsigned)d2);
                        }
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
                            if (is_agent) mrow_g[W2 + tile] = 1.0f;  // includes the observer (map.py:109-110)
                            if (is_obj) mrow_g[2 * W2 + tile] = 1.0f;
                            if (is_res) mrow_g[3 * W2 + tile] = 1.0f;
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
                                    const int e = e0 + lane;
                                    serve(e < ee ? __ldg(p.b.ent + e) : 0u, e, true);
                                }
                            }
                        }
                    }
                    // three warp reductions: packed counts (agents | objects << 10 | reserved << 20, each
                    // below 1024 because a window holds fewer than 1024 entities), centre flag, nearest
                    int counts, centre;
                    if (ee - eb < 1024) {
                        counts = __reduce_add_sync(FULL, n_ag | (n_obj << 10) | (n_res << 20));
                        n_ag = counts & 1023;
                        n_obj = (counts >> 10) & 1023;
                        n_res = (counts >> 20) & 1023;
                    } else {
                        n_ag = __reduce_add_sync(FULL, n_ag);
                        n_obj = __reduce_add_sync(FULL, n_obj);
                        n_res = __reduce_add_sync(FULL, n_res);
                    }
                    centre = __reduce_or_sync(FULL, (unsigned)centre_has);
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
                for (int v = tid; v < nvec; v += 32 * kRoleWarps) g4[v] = s4[v];
                const int tail0 = head + (nvec << 2);
                if (tid < total - tail0) feat_g0[tail0 + tid] = feat_s0[tail0 + tid];
            }
            // no second barrier: the next pass fills the other buffer, and the barrier of
            // that pass orders this sweep before the buffer is written again
        }

        if (do_kpi && warp == 0) {
            __syncwarp();
            if (lane < n_towns) {
                const KpiAcc k = s_kpi[lane];
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
}
