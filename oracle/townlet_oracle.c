/*
 * townlet_oracle.c — CPU restatement of Townlet's per-tick agent hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA
 * kernels in townlet_b200/csrc; it is imported only by tests/,
 * __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of
 * bench.py.  The product path never calls it and there is no CPU fallback.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle_golden.py)
 * against golden vectors produced by running the reference itself
 * (tests/golden/make_golden.py imports /root/reference), which include the
 * reference's own fixtures tests/data/observations/baseline_{hybrid,full,compact}.npz
 * and social_snippet_gold.npz.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * src/townlet/ of tachyon-beep/townlet).  It works on the same
 * structure-of-arrays batch (include/townlet_b200.h) as the kernels, but
 * deliberately shares no arithmetic with them: no look-up tables (it calls
 * libm sin/cos/hypot/pow exactly where the reference calls math/numpy), dense
 * per-town lookup tables in place of the reference's dicts, one agent at a
 * time in the reference's iteration order.
 */
#include "../include/townlet_b200.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- town cache: world/observations/cache.py:16-41 --------------------- */
typedef struct {
    int gw, gh, n_ctypes;
    uint16_t* agents;  /* agent_lookup: number of ids on the tile */
    uint16_t* objects; /* object_lookup: number of ids on the tile */
    uint8_t* reserved; /* reservation_tiles membership */
    uint16_t* ctype;   /* [n_ctypes][gw*gh] objects per compact channel type */
} TownCache;

static int cache_alloc(TownCache* c, int gw, int gh, int n_ctypes) {
    size_t n = (size_t)gw * (size_t)gh;
    c->gw = gw;
    c->gh = gh;
    c->n_ctypes = n_ctypes;
    c->agents = (uint16_t*)calloc(n, sizeof(uint16_t));
    c->objects = (uint16_t*)calloc(n, sizeof(uint16_t));
    c->reserved = (uint8_t*)calloc(n, 1);
    c->ctype = (uint16_t*)calloc(n * (size_t)(n_ctypes > 0 ? n_ctypes : 1), sizeof(uint16_t));
    return (c->agents && c->objects && c->reserved && c->ctype) ? 0 : -1;
}

static void cache_free(TownCache* c) {
    free(c->agents);
    free(c->objects);
    free(c->reserved);
    free(c->ctype);
}

/* build_local_cache, cache.py:28-40.  Entities were laid out by the ingest in
 * the order the reference iterates them; clearing only touched tiles keeps the
 * cost O(entities) like the dict build. */
static void cache_fill(TownCache* c, const uint32_t* ent, int n, int sign) {
    size_t plane = (size_t)c->gw * (size_t)c->gh;
    for (int i = 0; i < n; ++i) {
        uint32_t e = ent[i];
        size_t idx = (size_t)TL_ENT_Y(e) * (size_t)c->gw + (size_t)TL_ENT_X(e);
        switch (TL_ENT_KIND(e)) {
            case TL_KIND_AGENT:
                c->agents[idx] = (uint16_t)(c->agents[idx] + sign);
                break;
            case TL_KIND_OBJECT:
                if (TL_ENT_TILE(e)) {
                    c->objects[idx] = (uint16_t)(c->objects[idx] + sign);
                    unsigned ct = TL_ENT_CTYPE(e);
                    if (ct > 0 && (int)ct <= c->n_ctypes)
                        c->ctype[(ct - 1) * plane + idx] = (uint16_t)(c->ctype[(ct - 1) * plane + idx] + sign);
                }
                break;
            case TL_KIND_RESERVED:
                c->reserved[idx] = (uint8_t)(sign > 0 ? 1 : 0);
                break;
            default:
                break;
        }
    }
}

static inline int in_grid(const TownCache* c, int x, int y) {
    return x >= 0 && y >= 0 && x < c->gw && y < c->gh;
}

/* ---- LocalSummary: encoders/map.py:24-36 -------------------------------- */
typedef struct {
    int other_agents, object_total, reserved_tiles, nearest_d2;
    double agent_ratio, object_ratio, reserved_ratio, nearest_distance, nearest_norm;
} LocalSummary;

/* encode_map_tensor, encoders/map.py:39-148.  n_channels is 0 (summary only),
 * 4 (self, agents, objects, reservations) or 6 (+ path_dx, path_dy). */
static void encode_map_tensor(int n_channels, int cx, int cy, int radius, const TownCache* c, float* tensor,
                              LocalSummary* s) {
    int window = radius * 2 + 1;
    int total_tiles = window * window;
    if (n_channels > 0) {
        memset(tensor, 0, sizeof(float) * (size_t)n_channels * (size_t)total_tiles);
        tensor[0 * total_tiles + radius * window + radius] = 1.0f; /* map.py:69 */
    }
    int other_agents = 0, object_total = 0, reserved_tiles = 0, nearest_d2 = 0;
    double nearest = -1.0; /* None */
    for (int dy = -radius; dy <= radius; ++dy) {
        for (int dx = -radius; dx <= radius; ++dx) {
            int x = cx + dx, y = cy + dy;
            int agents = 0, objects = 0, reserved = 0;
            if (in_grid(c, x, y)) {
                size_t idx = (size_t)y * (size_t)c->gw + (size_t)x;
                agents = c->agents[idx];
                objects = c->objects[idx];
                reserved = c->reserved[idx];
            }
            if (agents) { /* map.py:88-100 */
                int count = agents;
                if (dx == 0 && dy == 0) count -= 1;
                other_agents += count > 0 ? count : 0;
                if (count > 0) {
                    double distance = hypot((double)dx, (double)dy);
                    if (distance > 0 && (nearest < 0 || distance < nearest)) {
                        nearest = distance;
                        nearest_d2 = dx * dx + dy * dy;
                    }
                }
            }
            object_total += objects;
            if (reserved) reserved_tiles += 1;
            if (n_channels == 0) continue;
            int tile = (dy + radius) * window + (dx + radius);
            if (agents) tensor[1 * total_tiles + tile] = 1.0f;
            if (objects) tensor[2 * total_tiles + tile] = 1.0f;
            if (reserved) tensor[3 * total_tiles + tile] = 1.0f;
            if (n_channels >= 6) { /* map.py:115-125 */
                double distance = hypot((double)dx, (double)dy);
                if (distance > 0) {
                    tensor[4 * total_tiles + tile] = (float)((double)dx / distance);
                    tensor[5 * total_tiles + tile] = (float)((double)dy / distance);
                }
            }
        }
    }
    int max_tiles = total_tiles - 1 > 1 ? total_tiles - 1 : 1; /* map.py:127 */
    int tt = total_tiles > 1 ? total_tiles : 1;
    s->agent_ratio = fmin(1.0, (double)other_agents / (double)max_tiles);
    s->object_ratio = fmin(1.0, (double)object_total / (double)tt);
    s->reserved_ratio = fmin(1.0, (double)reserved_tiles / (double)tt);
    if (nearest < 0) {
        s->nearest_norm = 0.0;
        s->nearest_distance = 0.0;
    } else {
        int r1 = radius > 1 ? radius : 1;
        s->nearest_norm = fmax(0.0, fmin(1.0, nearest / (double)r1));
        s->nearest_distance = nearest;
    }
    s->other_agents = other_agents;
    s->object_total = object_total;
    s->reserved_tiles = reserved_tiles;
    s->nearest_d2 = nearest_d2;
}

/* encode_compact_map, encoders/map.py:151-243.  Channels: self, agents,
 * objects, reservations, object:<type>*, walkable (service.py:170-180). */
static void encode_compact_map(int cx, int cy, int ox, int oy, int radius, const TownCache* c, int n_ctypes,
                               int normalize, float* tensor) {
    int window = radius * 2 + 1;
    int tiles = window * window;
    int n_channels = 5 + n_ctypes;
    size_t plane = (size_t)c->gw * (size_t)c->gh;
    memset(tensor, 0, sizeof(float) * (size_t)n_channels * (size_t)tiles);
    tensor[0 * tiles + radius * window + radius] = 1.0f; /* map.py:185-186 */
    for (int dy = -radius; dy <= radius; ++dy) {
        for (int dx = -radius; dx <= radius; ++dx) {
            int x = cx + dx, y = cy + dy;
            int tile = (dy + radius) * window + (dx + radius);
            int agents = 0, objects = 0, reserved = 0;
            size_t idx = 0;
            int inside = in_grid(c, x, y);
            if (inside) {
                idx = (size_t)y * (size_t)c->gw + (size_t)x;
                agents = c->agents[idx];
                objects = c->objects[idx];
                reserved = c->reserved[idx];
            }
            /* other_agents = ids on the tile different from the observer (map.py:198);
             * the observer is listed exactly once, on its occupancy tile (ox, oy). */
            int others = agents - ((x == ox && y == oy) ? 1 : 0);
            if (others > 0) {
                double v = (double)others;
                if (normalize) v = fmin(1.0, v);
                tensor[1 * tiles + tile] = (float)v;
            }
            if (objects > 0) {
                double v = (double)objects;
                if (normalize) v = fmin(1.0, v);
                tensor[2 * tiles + tile] = (float)v;
            }
            if (reserved) tensor[3 * tiles + tile] = 1.0f;
            for (int k = 0; k < n_ctypes; ++k) { /* map.py:214-229 */
                int cnt = inside ? c->ctype[(size_t)k * plane + idx] : 0;
                if (cnt > 0) {
                    double v = (double)cnt;
                    if (normalize) v = fmin(1.0, v);
                    tensor[(4 + k) * tiles + tile] = (float)v;
                }
            }
            double walkable = 1.0; /* map.py:231-241 */
            if (others > 0 || objects > 0 || reserved || (dx == 0 && dy == 0)) walkable = 0.0;
            tensor[(4 + n_ctypes) * tiles + tile] = (float)walkable;
        }
    }
}

/* find_nearest_object_of_type, world/observe.py:109-147: first minimum of the
 * squared distance over objects in insertion order. */
static int find_nearest(const uint32_t* ent, int n_ent, unsigned ltype, int ox, int oy, int* tx, int* ty) {
    long best = -1;
    for (int i = 0; i < n_ent; ++i) {
        uint32_t e = ent[i];
        if (TL_ENT_KIND(e) != TL_KIND_OBJECT || !TL_ENT_LM(e) || TL_ENT_LTYPE(e) != ltype) continue;
        long dx = TL_ENT_X(e) - ox, dy = TL_ENT_Y(e) - oy;
        long d2 = dx * dx + dy * dy;
        if (best < 0 || d2 < best) {
            best = d2;
            *tx = TL_ENT_X(e);
            *ty = TL_ENT_Y(e);
        }
    }
    return best >= 0;
}

static inline int pos_x(int32_t p) { return (int)(int16_t)(p & 0xffff); }
static inline int pos_y(int32_t p) { return (int)(int16_t)((uint32_t)p >> 16); }

/* numpy add.reduce over a contiguous float64 vector of n <= 8 values, as
 * np.mean uses it (encoders/social.py:245-254): plain left-to-right below 8,
 * the 8-accumulator pairwise block at exactly 8. */
static double np_sum_le8(const double* a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    return ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
}

/* stable descending order of keys[0..n) — Python sorted(..., reverse=True) */
static void stable_desc_order(const double* keys, int n, int* order) {
    for (int i = 0; i < n; ++i) order[i] = i;
    for (int i = 1; i < n; ++i) {
        int v = order[i];
        int j = i - 1;
        while (j >= 0 && keys[order[j]] < keys[v]) {
            order[j + 1] = order[j];
            --j;
        }
        order[j + 1] = v;
    }
}

/* encode_social_vector + _collect_social_slots + _resolve_relationships,
 * encoders/social.py:27-209.  Writes social_len floats at out[]. */
static void encode_social(const TlTownBatch* b, const TlObsParams* p, int a, float* out, int* filled, int* source) {
    int K = b->max_edges, EW = b->embed_words;
    int total_slots = p->top_friends + p->top_rivals;
    int slot_dim = p->embed_dim + 3;
    int len = total_slots * slot_dim + (p->include_aggregates ? 4 : 0);
    memset(out, 0, sizeof(float) * (size_t)len);
    *filled = 0;
    *source = 0;
    if (total_slots <= 0) return;

    /* _resolve_relationships, social.py:165-209 */
    int n = 0;
    double trust[TL_MAX_EDGES], fam[TL_MAX_EDGES], riv[TL_MAX_EDGES];
    const uint64_t* emb[TL_MAX_EDGES];
    int nt = b->tie_count[a];
    if (nt > 0) {
        n = nt;
        for (int j = 0; j < n; ++j) {
            trust[j] = b->tie_trust[(size_t)a * K + j];
            fam[j] = b->tie_fam[(size_t)a * K + j];
            riv[j] = b->tie_riv[(size_t)a * K + j];
            emb[j] = b->tie_embed + ((size_t)a * K + j) * EW;
        }
        *source = 1;
    } else {
        int nr = b->riv_count[a];
        n = nr < p->top_rivals ? nr : p->top_rivals; /* rivalry_top(agent, limit=top_rivals) */
        for (int j = 0; j < n; ++j) {
            trust[j] = 0.0;
            fam[j] = 0.0;
            riv[j] = b->riv_score[(size_t)a * K + j];
            emb[j] = b->riv_embed + ((size_t)a * K + j) * EW;
        }
        *source = n > 0 ? 2 : 0;
    }

    /* _collect_social_slots, social.py:137-163 */
    double key[TL_MAX_EDGES];
    int forder[TL_MAX_EDGES], rorder[TL_MAX_EDGES], chosen[TL_MAX_SOCIAL_SLOTS], is_friend[TL_MAX_EDGES];
    for (int j = 0; j < n; ++j) key[j] = trust[j] + fam[j];
    stable_desc_order(key, n, forder);
    stable_desc_order(riv, n, rorder);
    int n_slots = 0;
    memset(is_friend, 0, sizeof(is_friend));
    for (int j = 0; j < n && j < p->top_friends; ++j) {
        chosen[n_slots++] = forder[j];
        is_friend[forder[j]] = 1;
    }
    int n_rivals = 0;
    for (int j = 0; j < n && n_rivals < p->top_rivals; ++j) {
        if (is_friend[rorder[j]]) continue;
        chosen[n_slots++] = rorder[j];
        ++n_rivals;
    }

    /* encode_social_vector, social.py:72-115 */
    double tvals[TL_MAX_SOCIAL_SLOTS], rvals[TL_MAX_SOCIAL_SLOTS];
    int offset = 0;
    for (int s = 0; s < n_slots; ++s) {
        int j = chosen[s];
        const uint8_t* bytes = (const uint8_t*)emb[j];
        for (int d = 0; d < p->embed_dim; ++d) /* _embed_agent_id, social.py:236-242: float32 division */
            out[offset + d] = (float)bytes[d % 32] / 255.0f;
        offset += p->embed_dim;
        out[offset++] = (float)trust[j];
        out[offset++] = (float)fam[j];
        out[offset++] = (float)riv[j];
        tvals[s] = trust[j];
        rvals[s] = riv[j];
    }
    if (p->include_aggregates) {
        offset = total_slots * slot_dim;
        if (n_slots > 0) { /* _compute_aggregates, social.py:245-254 */
            double tmax = tvals[0], rmax = rvals[0];
            for (int s = 1; s < n_slots; ++s) {
                if (tvals[s] > tmax) tmax = tvals[s];
                if (rvals[s] > rmax) rmax = rvals[s];
            }
            out[offset + 0] = (float)(np_sum_le8(tvals, n_slots) / (double)n_slots);
            out[offset + 1] = (float)tmax;
            out[offset + 2] = (float)(np_sum_le8(rvals, n_slots) / (double)n_slots);
            out[offset + 3] = (float)rmax;
        }
    }
    *filled = n_slots;
}

static inline double clamp01(double v) { return fmax(0.0, fmin(1.0, v)); }

/* encode_feature_vector and its sub-encoders, encoders/features.py:43-404,
 * plus the social splice and ctx-reset handling of service.py:224-229,313-315. */
static void encode_features(const TlTownBatch* b, const TlObsParams* p, int t, int a, const LocalSummary* ls,
                            int terminated, float* f, int32_t* summary) {
    int N = b->n_agents;
    memset(f, 0, sizeof(float) * (size_t)p->n_features);
    /* _encode_needs, features.py:115-123 */
    f[TL_FEAT_NEED_HUNGER] = (float)b->needs[0 * (size_t)N + a];
    f[TL_FEAT_NEED_HYGIENE] = (float)b->needs[1 * (size_t)N + a];
    f[TL_FEAT_NEED_ENERGY] = (float)b->needs[2 * (size_t)N + a];
    /* _encode_wallet_and_employment, features.py:126-139 */
    uint32_t fl = b->flags[a];
    f[TL_FEAT_WALLET] = (float)b->wallet[a];
    f[TL_FEAT_LATENESS] = (float)(double)b->lateness[a];
    f[TL_FEAT_ON_SHIFT] = (fl & TL_F_ON_SHIFT) ? 1.0f : 0.0f;
    f[TL_FEAT_ATTENDANCE] = (float)b->attendance[a];
    f[TL_FEAT_WAGES_WITHHELD] = (float)b->wages_withheld[a];
    /* _encode_time, features.py:142-148 */
    {
        int tpd = p->ticks_per_day;
        int tick = b->town_tick[t];
        int m = tick % tpd;
        if (m < 0) m += tpd; /* Python modulo */
        double phase = (double)m / (double)tpd;
        const double tau = 6.283185307179586;
        f[TL_FEAT_TIME_SIN] = (float)sin(tau * phase);
        f[TL_FEAT_TIME_COS] = (float)cos(tau * phase);
    }
    /* _encode_shift_state, features.py:151-169 */
    {
        unsigned st = (fl & TL_F_SHIFT_MASK) >> TL_F_SHIFT_SHIFT;
        if (st > 4) st = 0;
        for (unsigned k = 0; k < 5; ++k) f[TL_FEAT_SHIFT_PRE + k] = (k == st) ? 1.0f : 0.0f;
    }
    /* _encode_embedding_slot, features.py:172-177 */
    f[TL_FEAT_SLOT_NORM] = (float)((double)b->slot[a] / (double)p->max_slots);
    f[TL_FEAT_CTX_RESET] = 0.0f;
    /* _encode_episode_progress, features.py:180-188 */
    {
        int len = p->ticks_per_day > 1 ? p->ticks_per_day : 1;
        double progress = (double)b->episode_tick[a] / (double)len;
        f[TL_FEAT_EPISODE_PROGRESS] = (float)clamp01(progress);
    }
    /* _encode_last_action, features.py:191-206 */
    f[TL_FEAT_LAST_HASH] = b->last_action_hash[a];
    f[TL_FEAT_LAST_SUCCESS] = (fl & TL_F_LAST_SUCCESS) ? 1.0f : 0.0f;
    f[TL_FEAT_LAST_DURATION] = (float)(double)b->last_action_duration[a];
    /* _encode_environmental_flags, features.py:209-229 */
    f[TL_FEAT_RESERVATION] = (fl & TL_F_RESERVATION) ? 1.0f : 0.0f;
    f[TL_FEAT_IN_QUEUE] = (fl & TL_F_IN_QUEUE) ? 1.0f : 0.0f;
    /* _encode_rivalry, features.py:232-248: rivalry_top(agent, max_edges) */
    {
        int K = b->max_edges;
        int nr = b->riv_count[a];
        if (nr > p->rivalry_max_edges) nr = p->rivalry_max_edges;
        double mx = nr > 0 ? b->riv_score[(size_t)a * K] : 0.0;
        int avoid = 0;
        for (int j = 0; j < nr; ++j)
            if (b->riv_score[(size_t)a * K + j] >= p->avoid_threshold) ++avoid;
        f[TL_FEAT_RIVALRY_MAX] = (float)mx;
        f[TL_FEAT_RIVALRY_AVOID] = (float)(double)avoid;
    }
    /* _encode_path_hint, features.py:251-286 */
    int e0 = b->town_ent_off[t], e1 = b->town_ent_off[t + 1];
    int cx = pos_x(b->pos[a]), cy = pos_y(b->pos[a]);
    {
        int tx = 0, ty = 0;
        double north = 0, south = 0, east = 0, west = 0;
        if (find_nearest(b->ent + e0, e1 - e0, (unsigned)p->path_hint_ltype, cx, cy, &tx, &ty)) {
            int dx = tx - cx, dy = ty - cy;
            int norm = abs(dx) + abs(dy);
            if (norm != 0) {
                north = fmax(0.0, (double)-dy) / (double)norm;
                south = fmax(0.0, (double)dy) / (double)norm;
                east = fmax(0.0, (double)dx) / (double)norm;
                west = fmax(0.0, (double)-dx) / (double)norm;
            }
        }
        f[TL_FEAT_PATH_NORTH + 0] = (float)north;
        f[TL_FEAT_PATH_NORTH + 1] = (float)south;
        f[TL_FEAT_PATH_NORTH + 2] = (float)east;
        f[TL_FEAT_PATH_NORTH + 3] = (float)west;
    }
    /* _encode_local_summary, features.py:289-330 */
    f[TL_FEAT_AGENT_RATIO] = (float)ls->agent_ratio;
    f[TL_FEAT_OBJECT_RATIO] = (float)ls->object_ratio;
    f[TL_FEAT_RESERVED_RATIO] = (float)ls->reserved_ratio;
    f[TL_FEAT_NEAREST] = (float)ls->nearest_norm;
    /* _encode_landmarks, features.py:333-356 */
    if (p->landmark_base >= 0) {
        for (unsigned l = 0; l < 4; ++l) {
            int tx = 0, ty = 0;
            float* dst = f + p->landmark_base + 3 * (int)l;
            if (!find_nearest(b->ent + e0, e1 - e0, l + 1, cx, cy, &tx, &ty)) {
                dst[0] = dst[1] = dst[2] = 0.0f;
                continue;
            }
            int dx = tx - cx, dy = ty - cy;
            double distance = hypot((double)dx, (double)dy);
            double norm = distance > 0 ? distance : 1.0;
            dst[0] = (float)((double)dx / norm);
            dst[1] = (float)((double)dy / norm);
            dst[2] = (float)distance;
        }
    }
    /* _encode_personality, features.py:359-380 */
    if (p->personality_base >= 0 && b->personality) {
        for (int k = 0; k < 3; ++k) f[p->personality_base + k] = b->personality[(size_t)k * N + a];
    }
    /* social splice, service.py:300-315 */
    int filled = 0, source = 0;
    if (p->social_base >= 0) encode_social(b, p, a, f + p->social_base, &filled, &source);
    /* ctx reset, service.py:224-229: a pending request is reported once (consume_ctx_reset_requests, grid.py:932-936) */
    if ((fl & TL_F_CTX_RESET) || terminated) f[TL_FEAT_CTX_RESET] = 1.0f;
    if (fl & TL_F_CTX_RESET) b->flags[a] = fl & ~TL_F_CTX_RESET;

    if (summary) {
        summary[TL_SUM_OTHER_AGENTS] = ls->other_agents;
        summary[TL_SUM_OBJECTS] = ls->object_total;
        summary[TL_SUM_RESERVED] = ls->reserved_tiles;
        summary[TL_SUM_NEAREST_D2] = ls->nearest_d2;
        summary[TL_SUM_FILLED_SLOTS] = filled;
        summary[TL_SUM_RELATION_SOURCE] = source;
        summary[TL_SUM_RESERVED0] = 0;
        summary[TL_SUM_RESERVED1] = 0;
    }
}

/* WorldState._apply_need_decay, world/grid.py:1439-1455 */
static void need_decay(const TlTownBatch* b, const TlRewardParams* r, int a) {
    int N = b->n_agents;
    unsigned profile = (b->flags[a] & TL_F_PROFILE_MASK) >> TL_F_PROFILE_SHIFT;
    for (int k = 0; k < TL_N_NEEDS; ++k) {
        double multiplier = 1.0;
        if (r->decay_personality && profile != 0) {
            multiplier = r->need_mult[profile][k];
            if (multiplier <= 0.0) multiplier = 1.0;
        }
        double adjusted = r->decay_rate[k] * multiplier;
        double* v = &b->needs[(size_t)k * N + a];
        *v = fmax(0.0, *v - adjusted);
    }
}

/* RewardEngine._apply_reward_biases, rewards/engine.py:220-255, for the three
 * built-in bias groups (social, employment, survival). */
static double scale_bias(double value, double factor) {
    if (factor <= 0.0) return value;
    if (value >= 0.0) return value * factor;
    return value / factor;
}

/* RewardEngine.compute body for one agent, rewards/engine.py:71-172 */
static double reward_agent(const TlTownBatch* b, const TlRewardParams* r, int t, int a, int terminated,
                           unsigned reason, double* comps) {
    int N = b->n_agents;
    int tick = b->town_tick[t];
    double c[TL_N_COMPS];
    memset(c, 0, sizeof(c));
    double total = r->survival_tick;
    c[TL_COMP_SURVIVAL] = r->survival_tick;
    double needs_penalty = 0.0;
    for (int k = 0; k < TL_N_NEEDS; ++k) { /* engine.py:76-80 */
        double deficit = 1.0 - b->needs[(size_t)k * N + a];
        needs_penalty += r->need_weight[k] * pow(fmax(0.0, deficit), 2.0);
    }
    total -= needs_penalty;
    c[TL_COMP_NEEDS_PENALTY] = -needs_penalty;
    double sb = 0.0, sp = 0.0, sa = 0.0;
    if (b->social_in) {
        sb = b->social_in[0 * (size_t)N + a];
        sp = b->social_in[1 * (size_t)N + a];
        sa = b->social_in[2 * (size_t)N + a];
    }
    total += sb + sp + sa; /* engine.py:87 */
    c[TL_COMP_SOCIAL_BONUS] = sb;
    c[TL_COMP_SOCIAL_PENALTY] = sp;
    c[TL_COMP_SOCIAL_AVOIDANCE] = sa;
    c[TL_COMP_SOCIAL] = sb + sp + sa;
    double wage = 0.0; /* _compute_wage_bonus, engine.py:375-386 */
    if (r->wage_rate > 0.0) wage = b->wages_paid[a] - b->wages_withheld[a];
    total += wage;
    c[TL_COMP_WAGE] = wage;
    double punct = 0.0; /* _compute_punctuality_bonus, engine.py:388-398 */
    if (r->punctuality_bonus > 0.0) punct = r->punctuality_bonus * b->punctuality[a];
    total += punct;
    c[TL_COMP_PUNCTUALITY] = punct;
    unsigned profile = (b->flags[a] & TL_F_PROFILE_MASK) >> TL_F_PROFILE_SHIFT;
    if (r->reward_personality && profile != 0) { /* engine.py:104-106, 220-255 */
        static const int groups[3][3] = {{TL_COMP_SOCIAL_BONUS, TL_COMP_SOCIAL_PENALTY, TL_COMP_SOCIAL_AVOIDANCE},
                                         {TL_COMP_WAGE, TL_COMP_PUNCTUALITY, -1},
                                         {TL_COMP_SURVIVAL, -1, -1}};
        double delta = 0.0;
        for (int g = 0; g < 3; ++g) {
            double factor = r->reward_bias[profile][g];
            if (factor == 1.0) continue;
            for (int i = 0; i < 3; ++i) {
                int idx = groups[g][i];
                if (idx < 0) continue;
                double base = c[idx];
                double adj = scale_bias(base, factor);
                c[idx] = adj;
                delta += adj - base;
            }
            if (g == 0) c[TL_COMP_SOCIAL] = c[TL_COMP_SOCIAL_BONUS] + c[TL_COMP_SOCIAL_PENALTY] + c[TL_COMP_SOCIAL_AVOIDANCE];
        }
        total += delta;
    }
    double penalty = 0.0; /* _compute_terminal_penalty, engine.py:400-413 */
    if (terminated) {
        if (reason == 1) penalty = r->faint_penalty;
        else if (reason == 2) penalty = r->eviction_penalty;
    }
    total += penalty;
    c[TL_COMP_TERMINAL_PENALTY] = penalty;
    double pre_guard = total;
    /* termination block bookkeeping, engine.py:113-122 with the prune of :363-373 */
    int32_t block = b->term_block_tick[a];
    if (r->block_window < 0) block = TL_TERM_BLOCK_NONE;
    else if (block != TL_TERM_BLOCK_NONE && (int64_t)tick - (int64_t)block > (int64_t)r->block_window)
        block = TL_TERM_BLOCK_NONE;
    if (terminated) block = tick;
    int blocked = r->block_window >= 0 && block != TL_TERM_BLOCK_NONE &&
                  (int64_t)tick - (int64_t)block <= (int64_t)r->block_window;
    b->term_block_tick[a] = block;
    if (terminated || blocked) {
        if (total > 0.0) total = 0.0;
    }
    double guard_adjust = total - pre_guard;
    double post = fmax(fmin(total, r->clip_per_tick), -r->clip_per_tick); /* engine.py:126-128 */
    double tick_clip_adjust = post - total;
    total = post;
    /* _reset_episode_totals, engine.py:415-424: totals of terminated agents are dropped first */
    double prev = terminated ? 0.0 : b->episode_total[a];
    double cumulative = prev + total;
    double episode_clip_adjust = 0.0;
    if (r->clip_per_episode > 0) { /* engine.py:130-139 */
        cumulative = fmax(fmin(cumulative, r->clip_per_episode), -r->clip_per_episode);
        double new_total = cumulative - prev;
        episode_clip_adjust = new_total - total;
        total = new_total;
    }
    b->episode_total[a] = cumulative;
    c[TL_COMP_CLIP_ADJUSTMENT] = guard_adjust + tick_clip_adjust + episode_clip_adjust;
    c[TL_COMP_TOTAL] = total;
    c[TL_COMP_HOMEOSTASIS] = c[TL_COMP_SURVIVAL] + c[TL_COMP_NEEDS_PENALTY];
    c[TL_COMP_WORK] = c[TL_COMP_WAGE] + c[TL_COMP_PUNCTUALITY];
    c[TL_COMP_GUARDRAIL_ACTIVE] = guard_adjust != 0.0 ? 1.0 : 0.0;
    c[TL_COMP_CLIPPED] = (tick_clip_adjust != 0.0 || episode_clip_adjust != 0.0) ? 1.0 : 0.0;
    if (comps) memcpy(comps, c, sizeof(c));
    return total;
}

/* What process_actions leaves on one agent (world/systems/affordances.py:124-229 after
 * world/actions.py:44-105): a move with a position relocates the agent (snapshot.position and the
 * spatial index, world/spatial.py:62-83) and succeeds; every other kind carries the success the host
 * services resolved; apply_affordance_outcome (world/affordances/core.py:198-206) records kind,
 * success and duration. */
static void apply_action(const TlTownBatch* b, const TlDynamicsParams* d, int t, int a, uint32_t w) {
    unsigned kind = TL_ACT_KIND(w);
    if (kind == TL_ACT_NONE) return;
    int success = (int)TL_ACT_SUCCESS(w);
    if (kind == TL_ACT_MOVE) {
        success = 0;
        if (TL_ACT_HAS_POS(w)) {
            int x = TL_ACT_X(w), y = TL_ACT_Y(w);
            b->pos[a] = (int32_t)(((uint32_t)x & 0xffffu) | ((uint32_t)y << 16));
            uint32_t* self_e = b->ent + b->town_ent_off[t] + (a - b->town_agent_off[t]);
            *self_e = (*self_e & ~0xfffffu) | (uint32_t)x | ((uint32_t)y << 10);
            success = 1;
        }
    }
    b->flags[a] = (b->flags[a] & ~TL_F_LAST_SUCCESS) | (success ? TL_F_LAST_SUCCESS : 0u);
    b->last_action_duration[a] = TL_ACT_DURATION(w);
    b->last_action_hash[a] = d->action_hash[kind];
}

/* _decay_value, world/relationships.py:155-162 */
static double decay_value(double value, double decay) {
    if (decay <= 0.0) return value;
    if (value > 0) return fmax(0.0, value - decay);
    if (value < 0) return fmin(0.0, value + decay);
    return 0.0;
}

/* RelationshipService.decay for one agent (world/agents/relationships_service.py:165-191): the rivalry
 * ledger first (RivalryLedger.decay, world/rivalry.py:68-83), then the relationship ledger
 * (RelationshipLedger.decay, world/relationships.py:78-89).  Evicted entries leave the row, the rest
 * keep their order; the freed tail is zeroed. */
static void social_decay(const TlTownBatch* b, const TlDynamicsParams* d, int a) {
    const int K = b->max_edges, EW = b->embed_words;
    const size_t r = (size_t)a * K;
    {
        int n = b->riv_count[a] < K ? b->riv_count[a] : K, w = 0;
        const double amount = d->rivalry_decay_per_tick * 1.0; /* decay_per_tick * float(ticks) */
        for (int j = 0; j < n; ++j) {
            double v = b->riv_score[r + j] - amount;
            v = fmax(d->rivalry_min, fmin(d->rivalry_max, v)); /* _clamp, world/rivalry.py:131-134 */
            if (v <= d->rivalry_eviction_threshold) continue;
            b->riv_score[r + w] = v;
            if (w != j) memcpy(b->riv_embed + (r + w) * EW, b->riv_embed + (r + j) * EW, sizeof(uint64_t) * EW);
            ++w;
        }
        for (int j = w; j < n; ++j) {
            b->riv_score[r + j] = 0.0;
            memset(b->riv_embed + (r + j) * EW, 0, sizeof(uint64_t) * EW);
        }
        b->riv_count[a] = (uint8_t)w;
    }
    {
        int n = b->tie_count[a] < K ? b->tie_count[a] : K, w = 0;
        for (int j = 0; j < n; ++j) {
            const double tr = decay_value(b->tie_trust[r + j], d->trust_decay);
            const double fa = decay_value(b->tie_fam[r + j], d->familiarity_decay);
            const double rv = decay_value(b->tie_riv[r + j], d->tie_rivalry_decay); /* `minimum` is unused there */
            if (tr == 0.0 && fa == 0.0 && rv == 0.0) continue;
            b->tie_trust[r + w] = tr;
            b->tie_fam[r + w] = fa;
            b->tie_riv[r + w] = rv;
            if (w != j) memcpy(b->tie_embed + (r + w) * EW, b->tie_embed + (r + j) * EW, sizeof(uint64_t) * EW);
            ++w;
        }
        for (int j = w; j < n; ++j) {
            b->tie_trust[r + j] = b->tie_fam[r + j] = b->tie_riv[r + j] = 0.0;
            memset(b->tie_embed + (r + j) * EW, 0, sizeof(uint64_t) * EW);
        }
        b->tie_count[a] = (uint8_t)w;
    }
}

/* refresh_reservations, world/systems/queues.py:55-86 (run every tick by queues.step, :14-31): the objects are
 * walked in order; an object with an active agent adds its tile to the reservation set and records the holder, a free
 * object discards the tile (world/spatial.py:106-119).  The set is replayed on a scratch list of tiles, then written
 * back as the slot words: a tile that is in the set is carried by the last slot standing on it.  Holders get
 * TL_F_RESERVATION (encoders/features.py:215-219). */
static void refresh_reservations(const TlTownBatch* b, int t) {
    int a0 = b->town_agent_off[t], a1 = b->town_agent_off[t + 1];
    int e0 = b->town_ent_off[t] + (a1 - a0), e1 = b->town_ent_off[t + 1];
    for (int a = a0; a < a1; ++a) b->flags[a] &= ~TL_F_RESERVATION;
    int n = e1 - e0;
    if (n <= 0) return;
    uint32_t* tiles = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n); /* the reservation set, as a list of tiles */
    int n_tiles = 0;
    for (int e = e0; e < e1; ++e) {
        int holder = b->ent_holder[e];
        if (holder < -1) continue; /* not a reservation slot */
        uint32_t tile = b->ent[e] & 0xfffffu;
        int at = -1;
        for (int i = 0; i < n_tiles; ++i)
            if (tiles[i] == tile) at = i;
        if (holder >= 0) { /* _reservation_tiles.add(position); active_reservations[object] = agent */
            if (at < 0) tiles[n_tiles++] = tile;
            if (holder < a1 - a0) b->flags[a0 + holder] |= TL_F_RESERVATION;
        } else if (at >= 0) { /* _reservation_tiles.discard(position) */
            tiles[at] = tiles[--n_tiles];
        }
    }
    for (int e = e0; e < e1; ++e) {
        if (b->ent_holder[e] < -1) continue;
        uint32_t tile = b->ent[e] & 0xfffffu;
        int reserved = 0, last = 1;
        for (int i = 0; i < n_tiles; ++i) reserved |= tiles[i] == tile;
        for (int l = e + 1; l < e1; ++l)
            if (b->ent_holder[l] >= -1 && (b->ent[l] & 0xfffffu) == tile) last = 0;
        uint32_t kind = (reserved && last) ? TL_KIND_RESERVED : TL_KIND_EMPTY;
        b->ent[e] = (b->ent[e] & ~(3u << 20)) | (kind << 20);
    }
    free(tiles);
}

/* One town, one tick, in SimulationLoop.step order (core/sim_loop.py:632-958). */
static void town_tick(const TlTownBatch* b, const TlObsParams* p, const TlRewardParams* r, const TlDynamicsParams* d,
                      const TlObsOut* oo, const TlRewardOut* ro, uint32_t phases, int t, int it, TownCache* cache,
                      float* map_scratch) {
    int a0 = b->town_agent_off[t], a1 = b->town_agent_off[t + 1];
    int N = b->n_agents;
    if (phases & TL_PHASE_ADVANCE) b->town_tick[t] += 1; /* sim_loop.py:635,659 */
    if ((phases & TL_PHASE_ACTIONS) && d->actions) /* world/core/context.py:253-257 */
        for (int a = a0; a < a1; ++a) apply_action(b, d, t, a, d->actions[(size_t)it * d->actions_tick_stride + a]);
    if ((phases & TL_PHASE_RESERVATIONS) && b->ent_holder) refresh_reservations(b, t); /* queues.step, first of the systems */
    double k_rsum = 0, k_rsq = 0, k_hmin = INFINITY, k_hsum = 0, k_late = 0, k_wallet = 0;
    int k_starving = 0, k_term = 0;
    for (int a = a0; a < a1; ++a) {
        if (phases & TL_PHASE_DECAY) need_decay(b, r, a);
        if (phases & TL_PHASE_SOCIAL_DECAY) social_decay(b, d, a); /* systems order, world/systems/__init__.py:9-19 */
        uint32_t fl = b->flags[a];
        if (r && r->fuse_faint && (phases & (TL_PHASE_DECAY | TL_PHASE_REWARD))) {
            /* LifecycleManager.evaluate, lifecycle/manager.py:39-45: recomputed every tick */
            fl = (fl & ~TL_F_FAINTED) | (b->needs[a] <= r->faint_threshold ? TL_F_FAINTED : 0u);
            b->flags[a] = fl;
        }
        int terminated = (fl & (TL_F_TERMINATED | TL_F_FAINTED)) != 0;
        unsigned reason = (fl & TL_F_REASON_MASK) >> TL_F_REASON_SHIFT;
        if (reason == 0 && (fl & TL_F_FAINTED)) reason = 1; /* reasons.setdefault(agent, "faint") */
        if (phases & TL_PHASE_ADVANCE) { /* world/core/context.py:283-289 */
            int span = p ? (p->ticks_per_day > 1 ? p->ticks_per_day : 1) : 1;
            int e = (b->episode_tick[a] + 1) % span;
            b->episode_tick[a] = e < 0 ? e + span : e; /* Python's modulo is never negative */
        }
        if (phases & TL_PHASE_REWARD) {
            double* comps = (ro && ro->comps) ? ro->comps + (size_t)it * ro->comps_tick_stride + (size_t)a * TL_N_COMPS : NULL;
            double total = reward_agent(b, r, t, a, terminated, reason, comps);
            if (ro && ro->reward) ro->reward[(size_t)it * ro->reward_tick_stride + a] = (float)total;
            k_rsum += total;
            k_rsq += total * total;
        }
        if (ro && ro->terminated && (phases & (TL_PHASE_DECAY | TL_PHASE_REWARD)))
            ro->terminated[(size_t)it * ro->terminated_tick_stride + a] = (uint8_t)terminated;
        double hunger = b->needs[a];
        k_hmin = fmin(k_hmin, hunger);
        k_hsum += hunger;
        if (r && hunger <= r->starvation_threshold) ++k_starving;
        k_term += terminated;
        k_late += (double)b->lateness[a];
        k_wallet += b->wallet[a];
    }
    if (ro && ro->kpi && (phases & (TL_PHASE_DECAY | TL_PHASE_REWARD))) {
        float* k = ro->kpi + (size_t)it * ro->kpi_tick_stride + (size_t)t * TL_N_KPI;
        int n = a1 - a0;
        k[TL_KPI_REWARD_SUM] = (float)k_rsum;
        k[TL_KPI_REWARD_SUMSQ] = (float)k_rsq;
        k[TL_KPI_HUNGER_MIN] = n ? (float)k_hmin : 0.0f;
        k[TL_KPI_HUNGER_MEAN] = n ? (float)(k_hsum / n) : 0.0f;
        k[TL_KPI_STARVING] = (float)k_starving;
        k[TL_KPI_TERMINATED] = (float)k_term;
        k[TL_KPI_LATENESS_MEAN] = n ? (float)(k_late / n) : 0.0f;
        k[TL_KPI_WALLET_MEAN] = n ? (float)(k_wallet / n) : 0.0f;
    }
    if (!(phases & TL_PHASE_OBS)) return;

    /* WorldObservationService.build_batch, world/observations/service.py:201-233 */
    int e0 = b->town_ent_off[t], e1 = b->town_ent_off[t + 1];
    cache_fill(cache, b->ent + e0, e1 - e0, +1);
    int radius = p->window / 2;
    size_t map_elems = (size_t)p->n_channels * p->window * p->window;
    for (int a = a0; a < a1; ++a) {
        int cx = pos_x(b->pos[a]), cy = pos_y(b->pos[a]);
        float* map = oo->map ? oo->map + (size_t)it * oo->map_tick_stride + (size_t)a * map_elems : map_scratch;
        LocalSummary ls;
        if (p->variant == TL_VARIANT_COMPACT) { /* service.py:394-423 */
            encode_map_tensor(0, cx, cy, radius, cache, NULL, &ls);
            uint32_t self_e = b->ent[e0 + (a - a0)];
            encode_compact_map(cx, cy, TL_ENT_X(self_e), TL_ENT_Y(self_e), radius, cache, p->n_ctypes,
                               p->normalize_counts, map);
        } else {
            encode_map_tensor(p->variant == TL_VARIANT_FULL ? 6 : 4, cx, cy, radius, cache, map, &ls);
        }
        int terminated = (b->flags[a] & (TL_F_TERMINATED | TL_F_FAINTED)) != 0;
        float* feat = oo->features + (size_t)it * oo->features_tick_stride + (size_t)a * p->n_features;
        int32_t* summary = oo->summary ? oo->summary + (size_t)it * oo->summary_tick_stride + (size_t)a * TL_N_SUMMARY : NULL;
        encode_features(b, p, t, a, &ls, terminated, feat, summary);
    }
    cache_fill(cache, b->ent + e0, e1 - e0, -1);
    (void)N;
}

/* Oracle entry point: same argument meaning as tl_tick, every pointer HOST
 * memory.  n_threads > 1 shards towns over OpenMP threads (towns are
 * independent; used by the CPU baseline timing only). */
int tlo_tick_world(const TlTownBatch* b, const TlObsParams* p, const TlRewardParams* r, const TlDynamicsParams* d,
                   const TlObsOut* oo, const TlRewardOut* ro, uint32_t phases, int32_t n_ticks, int32_t n_threads) {
    if (!b || n_ticks < 0) return TL_ERR_ARG;
    if ((phases & (TL_PHASE_ACTIONS | TL_PHASE_SOCIAL_DECAY)) && !d) return TL_ERR_ARG;
    if ((phases & TL_PHASE_OBS) && (!p || !oo || !oo->features)) return TL_ERR_ARG;
    if ((phases & (TL_PHASE_DECAY | TL_PHASE_REWARD)) && !r) return TL_ERR_ARG;
    if (b->max_edges > TL_MAX_EDGES) return TL_ERR_ARG;
    if (n_threads < 1) n_threads = 1;
    int status = 0;
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads)
#endif
    {
        TownCache cache;
        int n_ctypes = p ? p->n_ctypes : 0;
        size_t map_elems = p ? (size_t)p->n_channels * p->window * p->window : 1;
        float* scratch = (float*)malloc(sizeof(float) * (map_elems ? map_elems : 1));
        if (cache_alloc(&cache, b->grid_w, b->grid_h, n_ctypes) != 0 || !scratch) {
            status = TL_ERR_ARG;
        } else {
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
            for (int t = 0; t < b->n_towns; ++t)
                for (int it = 0; it < n_ticks; ++it) town_tick(b, p, r, d, oo, ro, phases, t, it, &cache, scratch);
        }
        cache_free(&cache);
        free(scratch);
    }
    return status;
}

int tlo_tick(const TlTownBatch* b, const TlObsParams* p, const TlRewardParams* r, const TlObsOut* oo,
             const TlRewardOut* ro, uint32_t phases, int32_t n_ticks, int32_t n_threads) {
    return tlo_tick_world(b, p, r, NULL, oo, ro, phases, n_ticks, n_threads);
}

/* compute_gae, policy/backends/pytorch/ppo_utils.py:19-67, on the time-major layout of the rollout
 * (the reference indexes [batch, t]; here [t][agent]).  float32 like the reference's tensors; the Python
 * float scalars gamma and gamma * gae_lambda enter float32 tensor ops rounded to float32. */
int tlo_gae(const float* rewards, const float* values, const float* dones, float* advantages, float* returns,
            int32_t n_steps, int32_t n_agents, int32_t bootstrap, double gamma, double gae_lambda) {
    if (!rewards || !values || !dones || !advantages || !returns || n_steps < 0 || n_agents < 0) return TL_ERR_ARG;
    const float g = (float)gamma, gl = (float)(gamma * gae_lambda);
    for (int b = 0; b < n_agents; ++b) {
        volatile float gae = 0.0f; /* volatile: keep every intermediate rounded to float32 */
        for (int t = n_steps - 1; t >= 0; --t) {
            size_t i = (size_t)t * (size_t)n_agents + (size_t)b;
            volatile float mask = 1.0f - dones[i];
            float next_value = bootstrap ? values[(size_t)(t + 1) * (size_t)n_agents + (size_t)b] : values[i];
            volatile float t1 = g * next_value;
            volatile float t2 = t1 * mask;
            volatile float t3 = rewards[i] + t2;
            volatile float delta = t3 - values[i];
            volatile float c1 = gl * mask;
            volatile float c2 = c1 * gae;
            gae = delta + c2;
            advantages[i] = gae;
            returns[i] = gae + values[i];
        }
    }
    return TL_OK;
}

int tlo_abi_version(void) { return TL_ABI_VERSION; }

/* ---------------------------------------------------------------------------
 * Employment shift state machine: EmploymentEngine.apply_job_state,
 * src/townlet/world/employment.py:63-142, restated helper by helper (the ctx
 * dictionary of context_defaults, :167-193, is one array per key in
 * TlEmploymentState).  TEST INFRASTRUCTURE like the rest of this file.
 * ------------------------------------------------------------------------- */
typedef struct {
    const TlTownBatch* b;
    const TlEmploymentParams* p;
    const TlEmploymentState* s;
    uint32_t events;
} EmpCtx;

static void emp_set_shift_state(const TlTownBatch* b, int a, int state, int on_shift) {
    /* snapshot.shift_state / snapshot.on_shift -> the flag word the features read (features.py:133, 151-169) */
    unsigned code = state == TL_EMP_ON_TIME ? 1u : state == TL_EMP_LATE ? 2u : state == TL_EMP_ABSENT ? 3u : state == TL_EMP_POST_SHIFT ? 4u : 0u;
    b->flags[a] = (b->flags[a] & ~(TL_F_ON_SHIFT | TL_F_SHIFT_MASK)) | (on_shift ? TL_F_ON_SHIFT : 0u) | (code << TL_F_SHIFT_SHIFT);
}

static void emp_finalize_shift(EmpCtx* e, int a) { /* :420-455 */
    const TlEmploymentState* s = e->s;
    if (!(s->ctx_flags[a] & TL_EMP_F_SHIFT_STARTED) || (s->ctx_flags[a] & TL_EMP_F_OUTCOME_RECORDED)) {
        emp_set_shift_state(e->b, a, TL_EMP_POST_SHIFT, 0);
        return;
    }
    int scheduled = s->scheduled_ticks[a] > 1 ? s->scheduled_ticks[a] : 1;
    double attendance_value = (double)s->on_time_ticks[a] / (double)scheduled;
    int window = e->p->attendance_window > 1 ? e->p->attendance_window : 1;
    double* samples = s->attendance_samples + (size_t)a * TL_MAX_ATTENDANCE_WINDOW;
    int n = s->attendance_count[a];
    if (n == window) { /* deque(maxlen=window).append */
        memmove(samples, samples + 1, sizeof(double) * (size_t)(n - 1));
        n -= 1;
    }
    samples[n++] = attendance_value;
    s->attendance_count[a] = n;
    double total = 0.0;
    for (int i = 0; i < n; ++i) total += samples[i];
    ((double*)e->b->attendance)[a] = total / (double)n;
    s->ctx_flags[a] |= TL_EMP_F_OUTCOME_RECORDED;
    s->state[a] = TL_EMP_POST_SHIFT;
    emp_set_shift_state(e->b, a, TL_EMP_POST_SHIFT, 0);
    s->ctx_flags[a] &= ~(TL_EMP_F_SHIFT_STARTED | TL_EMP_F_LAST_PRESENT); /* shift_started_tick = shift_end_tick = last_present_tick = None */
    s->ctx_flags[a] &= ~(TL_EMP_F_LATE_PENALTY | TL_EMP_F_ABSENCE_PENALTY | TL_EMP_F_LATE_EVENT | TL_EMP_F_ABSENCE_EVENT |
                         TL_EMP_F_DEPARTURE_EVENT | TL_EMP_F_LATE_COUNTER);
    s->late_ticks_today[a] = s->late_ticks[a];
    s->late_ticks[a] = 0;
}

static void emp_idle_state(EmpCtx* e, int a) { /* :229-243 */
    if (e->s->state[a] != TL_EMP_PRE_SHIFT && e->s->state[a] != TL_EMP_IDLE) emp_finalize_shift(e, a);
    e->s->state[a] = TL_EMP_IDLE;
    emp_set_shift_state(e->b, a, TL_EMP_IDLE, 0);
}

static int emp_coworkers_on_shift(EmpCtx* e, int a0, int a1, int a) { /* :466-480: the list is only tested for emptiness here */
    int count = 0;
    for (int o = a0; o < a1; ++o) {
        if (o == a || e->s->job[o] != e->s->job[a]) continue;
        if (e->s->state[o] == TL_EMP_ON_TIME || e->s->state[o] == TL_EMP_LATE) ++count;
    }
    return count;
}

static void emp_apply_agent(EmpCtx* e, int t, int a0, int a1, int a, int tick) {
    const TlTownBatch* b = e->b;
    const TlEmploymentParams* p = e->p;
    const TlEmploymentState* s = e->s;
    double* wallet = (double*)b->wallet;
    double* wages_withheld = (double*)b->wages_withheld;
    double* wages_paid = (double*)b->wages_paid;
    int tpd = p->ticks_per_day > 1 ? p->ticks_per_day : 1;
    int current_day = tick / tpd;
    if (tick % tpd != 0 && tick < 0) --current_day; /* Python floor division */
    if (s->current_day[a] != current_day) { /* :73-78 */
        s->current_day[a] = current_day;
        s->late_ticks[a] = 0;
        wages_paid[a] = 0.0;
        s->late_ticks_today[a] = 0;
    }
    if (s->job[a] < 0 || s->job[a] >= p->n_jobs) { /* :80-86 */
        if (p->n_jobs <= 0) {
            emp_idle_state(e, a);
            return;
        }
        s->job[a] = 0;
    }
    int job = s->job[a];
    int start = p->job_start[job], end = p->job_end[job];
    double wage_rate = p->job_wage[job], lateness_penalty = p->job_lateness_penalty[job];
    int ox = s->town_origin ? s->town_origin[2 * t] : 0, oy = s->town_origin ? s->town_origin[2 * t + 1] : 0;
    int at_required_location = !p->job_has_location[job] || (pos_x(b->pos[a]) + ox == p->job_x[job] && pos_y(b->pos[a]) + oy == p->job_y[job]);
    int32_t* absence = s->absence_events + (size_t)a * TL_MAX_ABSENCE_EVENTS;
    while (s->absence_count[a] > 0 && (long long)tick - absence[0] > 7ll * tpd) { /* :103-107 popleft */
        memmove(absence, absence + 1, sizeof(int32_t) * (size_t)(s->absence_count[a] - 1));
        s->absence_count[a] -= 1;
    }
    s->absent_shifts_7d[a] = s->absence_count[a];
    if (tick < start - p->arrival_buffer) {
        emp_idle_state(e, a);
        return;
    }
    if (tick < start) { /* _employment_prepare_state :245-248 */
        s->state[a] = TL_EMP_AWAIT_START;
        emp_set_shift_state(b, a, TL_EMP_AWAIT_START, 0);
        return;
    }
    if (tick > end) {
        emp_finalize_shift(e, a);
        return;
    }
    /* _employment_begin_shift :250-268 */
    if (!(s->ctx_flags[a] & TL_EMP_F_SHIFT_STARTED) || s->shift_started_tick[a] != start) {
        s->shift_started_tick[a] = start;
        s->shift_end_tick[a] = end;
        s->scheduled_ticks[a] = end - start + 1 > 1 ? end - start + 1 : 1;
        s->on_time_ticks[a] = 0;
        s->late_ticks[a] = 0;
        wages_paid[a] = 0.0;
        s->ctx_flags[a] = TL_EMP_F_SHIFT_STARTED;
    }
    /* _employment_determine_state :270-301 */
    int ticks_since_start = tick - start;
    int state = s->state[a];
    if (at_required_location) {
        s->last_present_tick[a] = tick;
        s->ctx_flags[a] |= TL_EMP_F_LAST_PRESENT;
        if ((s->ctx_flags[a] & TL_EMP_F_EVER_ON_TIME) || ticks_since_start <= p->grace_ticks) {
            state = TL_EMP_ON_TIME;
            s->ctx_flags[a] |= TL_EMP_F_EVER_ON_TIME;
        } else {
            state = TL_EMP_LATE;
        }
    } else {
        if (ticks_since_start <= p->grace_ticks) state = TL_EMP_LATE;
        else if (ticks_since_start > p->absent_cutoff) state = TL_EMP_ABSENT;
        else if ((s->ctx_flags[a] & TL_EMP_F_LAST_PRESENT) && tick - s->last_present_tick[a] > p->absence_slack) state = TL_EMP_ABSENT;
        else state = TL_EMP_LATE;
    }
    /* _employment_apply_state_effects :303-418 */
    int previous_state = s->state[a];
    s->state[a] = state;
    int eligible_for_wage = (state == TL_EMP_ON_TIME || state == TL_EMP_LATE) && at_required_location;
    int coworkers = emp_coworkers_on_shift(e, a0, a1, a);
    int previous_working = previous_state == TL_EMP_ON_TIME || previous_state == TL_EMP_LATE;
    if (state == TL_EMP_ON_TIME) s->on_time_ticks[a] += 1;
    if (state == TL_EMP_LATE) {
        s->late_ticks[a] += 1;
        s->late_ticks_today[a] += 1;
        if (!(s->ctx_flags[a] & TL_EMP_F_LATE_PENALTY) && !previous_working) {
            wallet[a] = fmax(0.0, wallet[a] - lateness_penalty);
            s->ctx_flags[a] |= TL_EMP_F_LATE_PENALTY;
            if (!(s->ctx_flags[a] & TL_EMP_F_LATE_COUNTER)) {
                ((int32_t*)b->lateness)[a] += 1;
                s->ctx_flags[a] |= TL_EMP_F_LATE_COUNTER;
            }
        }
        if (!(s->ctx_flags[a] & TL_EMP_F_LATE_EVENT)) {
            e->events |= TL_EMP_EV_LATE_START;
            s->ctx_flags[a] |= TL_EMP_F_LATE_EVENT;
        }
        if (!at_required_location) {
            if (p->late_tick_penalty != 0.0) wallet[a] = fmax(0.0, wallet[a] - p->late_tick_penalty);
            wages_withheld[a] += wage_rate;
        }
        if (coworkers && !(s->ctx_flags[a] & TL_EMP_F_LATE_HELP_EVENT)) {
            e->events |= TL_EMP_EV_HELPED_WHEN_LATE;
            s->ctx_flags[a] |= TL_EMP_F_LATE_HELP_EVENT;
        }
    }
    if (state == TL_EMP_ABSENT) {
        if (!(s->ctx_flags[a] & TL_EMP_F_ABSENCE_PENALTY)) {
            wallet[a] = fmax(0.0, wallet[a] - p->absence_penalty);
            s->ctx_flags[a] |= TL_EMP_F_ABSENCE_PENALTY;
        }
        wages_withheld[a] += wage_rate;
        if (!(s->ctx_flags[a] & TL_EMP_F_ABSENCE_EVENT)) {
            e->events |= TL_EMP_EV_ABSENT;
            s->ctx_flags[a] |= TL_EMP_F_ABSENCE_EVENT;
            if (s->absence_count[a] == TL_MAX_ABSENCE_EVENTS) {
                memmove(absence, absence + 1, sizeof(int32_t) * (TL_MAX_ABSENCE_EVENTS - 1));
                s->absence_count[a] -= 1;
            }
            absence[s->absence_count[a]++] = tick;
            s->absent_shifts_7d[a] = s->absence_count[a];
        }
        if (coworkers && !(s->ctx_flags[a] & TL_EMP_F_TOOK_SHIFT_EVENT)) {
            e->events |= TL_EMP_EV_TOOK_MY_SHIFT;
            s->ctx_flags[a] |= TL_EMP_F_TOOK_SHIFT_EVENT;
        }
    } else if (state == TL_EMP_ON_TIME || state == TL_EMP_LATE) {
        s->ctx_flags[a] &= ~TL_EMP_F_ABSENCE_PENALTY;
    }
    if (eligible_for_wage && state != TL_EMP_ABSENT) {
        emp_set_shift_state(b, a, state, 1);
        wallet[a] += wage_rate;
        s->wages_earned[a] += 1;
        wages_paid[a] += wage_rate;
    } else {
        emp_set_shift_state(b, a, state, 0);
        if (previous_working && state != TL_EMP_ON_TIME && state != TL_EMP_LATE) {
            if (!(s->ctx_flags[a] & TL_EMP_F_DEPARTURE_EVENT) && previous_state != TL_EMP_ABSENT) {
                e->events |= TL_EMP_EV_DEPARTED_EARLY;
                s->ctx_flags[a] |= TL_EMP_F_DEPARTURE_EVENT;
            }
        }
    }
}

int tlo_employment_step(const TlTownBatch* b, const TlEmploymentParams* p, const TlEmploymentState* s, uint32_t* events_out,
                        int32_t n_ticks, int32_t tick_offset) {
    if (!b || !p || !s || n_ticks < 0) return TL_ERR_ARG;
    for (int it = 0; it < n_ticks; ++it) {
        for (int t = 0; t < b->n_towns; ++t) {
            int a0 = b->town_agent_off[t], a1 = b->town_agent_off[t + 1];
            int tick = b->town_tick[t] + tick_offset + it;
            for (int a = a0; a < a1; ++a) {
                EmpCtx e = {b, p, s, 0u};
                emp_apply_agent(&e, t, a0, a1, a, tick);
                /* employment_context_punctuality :215-222 */
                int scheduled = s->scheduled_ticks[a] > 1 ? s->scheduled_ticks[a] : 1;
                double value = (double)s->on_time_ticks[a] / (double)scheduled;
                ((double*)b->punctuality)[a] = fmax(0.0, fmin(1.0, value));
                if (events_out) events_out[(size_t)it * b->n_agents + a] = e.events;
            }
        }
    }
    return TL_OK;
}
