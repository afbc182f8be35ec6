// tick_grid.cuh — dense shared-memory grid form of the tick kernel
// (included by townlet_b200.cu inside its anonymous namespace; uses KParams, KpiAcc, seg_reduce,
// state_role_agent / social_role_agent, pos_x / pos_y).
//
// One CTA of 8 warps owns up to kMaxTownsPerCta towns whose occupancy grids fit in shared memory
// together.  Every tick the entity words are scattered into a uint16 cell grid (agents : 8 bits,
// objects : 7 bits, reserved : 1 bit; cache.py:16-41), then agents are served in passes of 32:
//
//   phase A  warp 0 runs the "state" role and warp 1 the "social" role, one THREAD per agent
//            (agent_roles.cuh, the same code the role-parallel entity kernel runs);
//   phase B  one WARP per agent gathers the window from the cell grid (encoders/map.py:39-243),
//            reduces the local summary with REDUX and stores tile and feature row with 16-byte
//            coalesced stores.
//
// The cost of this form grows with window tiles, not with entities, so the host picks it for towns
// with many entities per town (launch(): more than 384 on average) or when TL_MODE=grid.

struct ChunkMeta {
    int32_t cx, cy;      // window centre
    int32_t ox, oy;      // occupancy tile of the agent itself
    int32_t town_slot;   // index of the agent's town inside the CTA
    int32_t pad[3];
};

struct TownMeta {
    int32_t agent_begin, agent_end;
    int32_t ent_begin, ent_end;
    int32_t tick;
    int32_t pad;
};

template <int VARIANT>
__global__ void __launch_bounds__(kThreads) tick_kernel(const __grid_constant__ KParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ TownMeta s_town[kMaxTownsPerCta];
    __shared__ KpiAcc s_kpi[kMaxTownsPerCta];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int T = p.b.n_towns;
    const int t0 = blockIdx.x * p.towns_per_cta;
    const int n_towns = min(p.towns_per_cta, T - t0);
    if (n_towns <= 0) return;

    const bool do_decay = p.phases & TL_PHASE_DECAY;
    const bool do_reward = p.phases & TL_PHASE_REWARD;
    const bool do_obs = p.phases & TL_PHASE_OBS;
    const bool do_advance = p.phases & TL_PHASE_ADVANCE;
    const bool do_state = do_decay || do_reward;
    const bool do_actions = p.phases & TL_PHASE_ACTIONS;
    const bool do_social_decay = p.phases & TL_PHASE_SOCIAL_DECAY;

    uint16_t* s_grid = reinterpret_cast<uint16_t*>(smem + p.off_grid);
    uint16_t* s_ctype = p.off_ctype >= 0 ? reinterpret_cast<uint16_t*>(smem + p.off_ctype) : nullptr;
    float* s_feat = reinterpret_cast<float*>(smem + p.off_feat);
    float* s_map = reinterpret_cast<float*>(smem + p.off_map);
    ChunkMeta* s_meta = reinterpret_cast<ChunkMeta*>(smem + p.off_meta);

    const TlObsParams& o = p.obs;
    const int W = o.window;
    const int R = W >> 1;
    const int W2 = W * W;
    const int F = o.n_features;
    const int GW = p.b.grid_w, GH = p.b.grid_h;
    const int map_elems = o.n_channels * W2;

    if (tid < n_towns) {
        TownMeta m;
        m.agent_begin = p.b.town_agent_off[t0 + tid];
        m.agent_end = p.b.town_agent_off[t0 + tid + 1];
        m.ent_begin = p.b.town_ent_off[t0 + tid];
        m.ent_end = p.b.town_ent_off[t0 + tid + 1];
        m.tick = p.b.town_tick[t0 + tid];
        m.pad = 0;
        s_town[tid] = m;
    }
    __syncthreads();
    const int a_begin = s_town[0].agent_begin;
    const int a_end = s_town[n_towns - 1].agent_end;
    const int e_begin = s_town[0].ent_begin;
    const int e_end = s_town[n_towns - 1].ent_end;
    const int n_chunks = (a_end - a_begin + kChunk - 1) / kChunk;

    // per-lane window coordinates of tile (lane + 32 k) are advanced incrementally
    const int row0 = lane / W, col0 = lane - row0 * W;
    const int drow = 32 / W, dcol = 32 - drow * W;

    for (int it = 0; it < p.n_ticks; ++it) {
        // ---- tick bookkeeping (core/sim_loop.py:635,659) ---------------------
        if (tid < n_towns) {
            if (do_advance) {
                s_town[tid].tick += 1;
                p.b.town_tick[t0 + tid] = s_town[tid].tick;
            }
            KpiAcc z;
            z.rsum = z.rsq = z.hsum = z.late = z.wallet = 0.0;
            z.hmin = 1e300;
            z.starving = z.term = z.count = 0;
            s_kpi[tid] = z;
        }
        // world/core/context.py:253-257: actions land before any system or observer runs
        if (do_actions) apply_actions_cta(p, it, t0, n_towns, tid, kThreads);
        if (p.phases & TL_PHASE_RESERVATIONS) {  // queues.step: one warp per town refreshes the reservation slots
            if (do_actions) __syncthreads();  // both rewrite the agents' flag words
            for (int s = warp; s < n_towns; s += kWarps)
                refresh_reservations_town(p, s_town[s].agent_begin, s_town[s].agent_end, s_town[s].ent_begin, s_town[s].ent_end, lane);
        }
        // ---- stage the town grids (cache.py:16-41) ---------------------------
        if (do_obs) {
            const int words = (n_towns * p.grid_tiles + 1) >> 1;
            uint32_t* g32 = reinterpret_cast<uint32_t*>(s_grid);
            for (int i = tid; i < words; i += kThreads) g32[i] = 0u;
            if (s_ctype) {
                uint32_t* c32 = reinterpret_cast<uint32_t*>(s_ctype);
                for (int i = tid; i < words; i += kThreads) c32[i] = 0u;
            }
        }
        __syncthreads();
        if (do_obs) {
            for (int e = e_begin + tid; e < e_end; e += kThreads) {
                const uint32_t w = p.b.ent[e];
                int slot = 0;
#pragma unroll
                for (int s = 1; s < kMaxTownsPerCta; ++s)
                    if (s < n_towns && e >= s_town[s].ent_begin) slot = s;
                if (TL_ENT_X(w) >= GW || TL_ENT_Y(w) >= GH) continue;  // outside the declared grid: ignored, never out of bounds
                const int idx = slot * p.grid_tiles + TL_ENT_Y(w) * GW + TL_ENT_X(w);
                const unsigned kind = TL_ENT_KIND(w);
                uint32_t* word = reinterpret_cast<uint32_t*>(s_grid) + (idx >> 1);
                const int sh = (idx & 1) * 16;
                if (kind == TL_KIND_AGENT) {
                    atomicAdd(word, 1u << sh);
                } else if (kind == TL_KIND_OBJECT) {
                    if (TL_ENT_TILE(w)) {
                        atomicAdd(word, 0x100u << sh);
                        const unsigned ct = TL_ENT_CTYPE(w);
                        if (s_ctype && ct > 0 && (int)ct <= o.n_ctypes)
                            atomicAdd(reinterpret_cast<uint32_t*>(s_ctype) + (idx >> 1), 1u << (sh + 4 * (ct - 1)));
                    }
                } else if (kind == TL_KIND_RESERVED) {
                    atomicOr(word, 0x8000u << sh);
                }
            }
        }
        __syncthreads();

        for (int chunk = 0; chunk < n_chunks; ++chunk) {
            const int c_begin = a_begin + chunk * kChunk;
            const int c_n = min(kChunk, a_end - c_begin);
            // feature staging mirrors the 16-byte misalignment of the global rows
            float* feat_g0 = do_obs ? p.oo.features + (size_t)it * p.oo.features_tick_stride + (size_t)c_begin * F : nullptr;
            const int feat_mis = do_obs ? (int)((reinterpret_cast<uintptr_t>(feat_g0) >> 2) & 3) : 0;
            float* feat_s0 = s_feat + feat_mis;

            // =================== phase A: one thread per agent ==================
            if (warp < 2) {
                const bool active = lane < c_n;
                const int a = c_begin + (active ? lane : 0);
                int slot = 0;
#pragma unroll
                for (int s = 1; s < kMaxTownsPerCta; ++s)
                    if (s < n_towns && a >= s_town[s].agent_begin) slot = s;
                const TownMeta tm = s_town[slot];
                if (warp == 0) {
                    StateRoleOut so;
                    if (active)
                        state_role_agent<true>(p, it, a, tm.tick, tm.agent_begin, tm.agent_end, tm.ent_begin, tm.ent_end,
                                         0u /* actions were applied before the grid was staged */,
                                         load_state_in<true>(p, a), do_obs ? feat_s0 + lane * F : nullptr, so);
                    // ---- per-town KPIs: segmented warp-shuffle reduction ------------
                    if (do_state && p.ro.kpi) {
                        const int seg = active ? slot : -1 - lane;  // inactive lanes are their own segment
                        const double hunger = so.need0;
                        auto add = [](double x, double y) { return x + y; };
                        auto addi = [](int x, int y) { return x + y; };
                        auto mn = [](double x, double y) { return fmin(x, y); };
                        const double rsum = seg_reduce(so.reward_total, seg, lane, add);
                        const double rsq = seg_reduce(so.reward_total * so.reward_total, seg, lane, add);
                        const double hsum = seg_reduce(active ? hunger : 0.0, seg, lane, add);
                        const double hmin = seg_reduce(active ? hunger : 1e300, seg, lane, mn);
                        const double lsum = seg_reduce((double)so.kpi_late, seg, lane, add);
                        const double wsum = seg_reduce(so.kpi_wallet, seg, lane, add);
                        const int starving = seg_reduce((int)(active && hunger <= p.rew.starvation_threshold), seg, lane, addi);
                        const int term = seg_reduce((int)so.terminated, seg, lane, addi);
                        const int cnt = seg_reduce((int)active, seg, lane, addi);
                        const int prev_seg = __shfl_up_sync(0xffffffffu, seg, 1);
                        if (active && (lane == 0 || prev_seg != seg)) {  // segment head; chunks arrive in order
                            KpiAcc& k = s_kpi[slot];
                            k.rsum += rsum;
                            k.rsq += rsq;
                            k.hsum += hsum;
                            k.hmin = fmin(k.hmin, hmin);
                            k.late += lsum;
                            k.wallet += wsum;
                            k.starving += starving;
                            k.term += term;
                            k.count += cnt;
                        }
                    }
                    if (do_obs && active) {
                        const int32_t pw = p.b.pos[a];
                        const uint32_t self_e = p.b.ent[tm.ent_begin + (a - tm.agent_begin)];
                        ChunkMeta cm;
                        cm.cx = pos_x(pw);
                        cm.cy = pos_y(pw);
                        cm.ox = TL_ENT_X(self_e);
                        cm.oy = TL_ENT_Y(self_e);
                        cm.town_slot = slot;
                        cm.pad[0] = cm.pad[1] = cm.pad[2] = 0;
                        s_meta[lane] = cm;
                    }
                } else if ((do_obs || do_social_decay) && active) {
                    // rows are read straight from global memory: this form is not bound by the social role
                    const size_t K = (size_t)p.b.max_edges, EW = (size_t)p.b.embed_words;
                    int n_riv = min((int)p.b.riv_count[a], (int)K);  // counts are clamped to the row stride
                    int n_t = min((int)p.b.tie_count[a], (int)K);
                    const size_t r = (size_t)a * K;
                    if (do_social_decay) {
                        social_decay_agent(p, n_riv, n_t, p.b.tie_trust + r, p.b.tie_fam + r, p.b.tie_riv + r,
                                           p.b.riv_score + r, p.b.tie_embed + r * EW, p.b.riv_embed + r * EW);
                        p.b.riv_count[a] = (uint8_t)n_riv;
                        p.b.tie_count[a] = (uint8_t)n_t;
                    }
                    if (do_obs) {
                        int filled = 0, source = 0;
                        social_role_agent(p, n_riv, n_t, p.b.tie_trust + r, p.b.tie_fam + r, p.b.tie_riv + r,
                                          p.b.riv_score + r, p.b.tie_embed + r * EW, p.b.riv_embed + r * EW, o.embed_lut,
                                          feat_s0 + lane * F, filled, source);
                        if (p.oo.summary) {
                            int* srow_i = p.oo.summary + (size_t)it * p.oo.summary_tick_stride + (size_t)a * TL_N_SUMMARY;
                            *reinterpret_cast<int4*>(srow_i + TL_SUM_FILLED_SLOTS) = make_int4(filled, source, 0, 0);
                        }
                    }
                }
            }
            if (!do_obs) continue;  // uniform for the CTA
            __syncthreads();

            // =================== phase B: one warp per agent =====================
            for (int i = warp; i < c_n; i += kWarps) {
                const int a = c_begin + i;
                const ChunkMeta cm = s_meta[i];
                const uint16_t* grid = s_grid + cm.town_slot * p.grid_tiles;
                const uint16_t* cgrid = s_ctype ? s_ctype + cm.town_slot * p.grid_tiles : nullptr;
                float* frow = feat_s0 + i * F;

                // -- window gather (encoders/map.py:39-148, 151-243) -----------
                float* mrow_g = p.oo.map + (size_t)it * p.oo.map_tick_stride + (size_t)a * map_elems;
                const int map_mis = (int)((reinterpret_cast<uintptr_t>(mrow_g) >> 2) & 3);
                float* st = s_map + warp * p.map_stride + map_mis;
                int others_sum = 0, objects_sum = 0, reserved_sum = 0;
                unsigned nearest = 0xffffffffu;
                int row = row0, col = col0;
                for (int tile = lane; tile < W2; tile += 32) {
                    const int dx = col - R, dy = row - R;
                    const int x = cm.cx + dx, y = cm.cy + dy;
                    const bool inb = (unsigned)x < (unsigned)GW && (unsigned)y < (unsigned)GH;
                    const unsigned w = inb ? grid[y * GW + x] : 0u;
                    const int agents = w & 0xff, objects = (w >> 8) & 0x7f, reserved = w >> 15;
                    const bool centre = (dx == 0 && dy == 0);
                    // LocalSummary (map.py:88-103): self excluded on the centre tile only
                    const int cnt = agents - ((centre && agents > 0) ? 1 : 0);
                    others_sum += cnt;
                    objects_sum += objects;
                    reserved_sum += reserved;
                    if (cnt > 0 && !centre) nearest = min(nearest, (unsigned)(dx * dx + dy * dy));
                    st[tile] = centre ? 1.0f : 0.0f;  // map.py:69 / :185-186
                    if (VARIANT == TL_VARIANT_COMPACT) {
                        const bool own = (x == cm.ox && y == cm.oy);
                        const int others = agents - (own ? 1 : 0);  // map.py:198
                        const bool norm = o.normalize_counts != 0;
                        st[W2 + tile] = (float)(norm ? min(others, 1) : others);
                        st[2 * W2 + tile] = (float)(norm ? min(objects, 1) : objects);
                        st[3 * W2 + tile] = (float)reserved;
                        const unsigned cw = (cgrid && inb) ? cgrid[y * GW + x] : 0u;
                        for (int k = 0; k < o.n_ctypes; ++k) {
                            const int c = (cw >> (4 * k)) & 15;
                            st[(4 + k) * W2 + tile] = (float)(norm ? min(c, 1) : c);
                        }
                        const bool blockd = others > 0 || objects > 0 || reserved || centre;  // map.py:231-241
                        st[(4 + o.n_ctypes) * W2 + tile] = blockd ? 0.0f : 1.0f;
                    } else {
                        st[W2 + tile] = agents > 0 ? 1.0f : 0.0f;  // includes the observer (map.py:109-110)
                        st[2 * W2 + tile] = objects > 0 ? 1.0f : 0.0f;
                        st[3 * W2 + tile] = (float)reserved;
                        if (VARIANT == TL_VARIANT_FULL) {
                            st[4 * W2 + tile] = o.path_lut[tile];
                            st[5 * W2 + tile] = o.path_lut[W2 + tile];
                        }
                    }
                    col += dcol;
                    row += drow;
                    if (col >= W) {
                        col -= W;
                        row += 1;
                    }
                }
                others_sum = __reduce_add_sync(0xffffffffu, others_sum);
                objects_sum = __reduce_add_sync(0xffffffffu, objects_sum);
                reserved_sum = __reduce_add_sync(0xffffffffu, reserved_sum);
                nearest = __reduce_min_sync(0xffffffffu, nearest);
                const int nearest_d2 = nearest == 0xffffffffu ? 0 : (int)nearest;
                if (lane == 0) {  // _encode_local_summary, features.py:289-330 via host LUTs
                    frow[TL_FEAT_AGENT_RATIO] = o.agent_ratio_lut[min(others_sum, W2 - 1)];
                    frow[TL_FEAT_OBJECT_RATIO] = o.tile_ratio_lut[min(objects_sum, W2)];
                    frow[TL_FEAT_RESERVED_RATIO] = o.tile_ratio_lut[min(reserved_sum, W2)];
                    frow[TL_FEAT_NEAREST] = o.nearest_lut[nearest_d2];
                }
                __syncwarp();

                // -- 16-byte vectorised coalesced stores -----------------------
                {
                    const int head = (4 - map_mis) & 3;
                    if (lane < head) mrow_g[lane] = st[lane];
                    const int nvec = (map_elems - head) >> 2;
                    float4* g4 = reinterpret_cast<float4*>(mrow_g + head);
                    const float4* s4 = reinterpret_cast<const float4*>(st + head);
                    for (int v = lane; v < nvec; v += 32) g4[v] = s4[v];
                    const int tail0 = head + (nvec << 2);
                    if (lane < map_elems - tail0) mrow_g[tail0 + lane] = st[tail0 + lane];
                }
                {
                    float* frow_g = feat_g0 + (size_t)i * F;
                    const int mis = (int)((reinterpret_cast<uintptr_t>(frow_g) >> 2) & 3);
                    const int head = min((4 - mis) & 3, F);
                    if (lane < head) frow_g[lane] = frow[lane];
                    const int nvec = (F - head) >> 2;
                    float4* g4 = reinterpret_cast<float4*>(frow_g + head);
                    const float4* s4 = reinterpret_cast<const float4*>(frow + head);
                    for (int v = lane; v < nvec; v += 32) g4[v] = s4[v];
                    const int tail0 = head + (nvec << 2);
                    if (lane < F - tail0) frow_g[tail0 + lane] = frow[tail0 + lane];
                }
                if (p.oo.summary && lane < TL_SUM_FILLED_SLOTS) {  // the social role wrote the upper half of the row
                    int v = 0;
                    if (lane == TL_SUM_OTHER_AGENTS) v = others_sum;
                    else if (lane == TL_SUM_OBJECTS) v = objects_sum;
                    else if (lane == TL_SUM_RESERVED) v = reserved_sum;
                    else if (lane == TL_SUM_NEAREST_D2) v = nearest_d2;
                    p.oo.summary[(size_t)it * p.oo.summary_tick_stride + (size_t)a * TL_N_SUMMARY + lane] = v;
                }
            }
            __syncthreads();  // staging rows are reused by the next chunk
        }

        // ---- per-town KPI rows --------------------------------------------------
        if (do_state && p.ro.kpi) {
            __syncthreads();
            if (tid < n_towns) {
                const KpiAcc k = s_kpi[tid];
                float* dst = p.ro.kpi + (size_t)it * p.ro.kpi_tick_stride + (size_t)(t0 + tid) * TL_N_KPI;
                const double n = (double)k.count;
                dst[TL_KPI_REWARD_SUM] = (float)k.rsum;
                dst[TL_KPI_REWARD_SUMSQ] = (float)k.rsq;
                dst[TL_KPI_HUNGER_MIN] = k.count ? (float)k.hmin : 0.0f;
                dst[TL_KPI_HUNGER_MEAN] = k.count ? (float)(k.hsum / n) : 0.0f;
                dst[TL_KPI_STARVING] = (float)k.starving;
                dst[TL_KPI_TERMINATED] = (float)k.term;
                dst[TL_KPI_LATENESS_MEAN] = k.count ? (float)(k.late / n) : 0.0f;
                dst[TL_KPI_WALLET_MEAN] = k.count ? (float)(k.wallet / n) : 0.0f;
            }
        }
        __syncthreads();  // state written this tick is visible to the CTA's next tick
    }
}
