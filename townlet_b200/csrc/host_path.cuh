// host_path.cuh — the host-buffer entry points (included at the end of townlet_b200.cu; uses validate, prepare,
// issue, cuda_ok, set_error, g_launches).
//
// What crosses the bus per step and how:
//   up    the batch arrays the caller marks dirty (one copy when the batch is one declared arena), the look-up
//         tables gathered into one pinned block (one copy), the action words;
//   down  the map in its wire format (float32 / uint8 / bits), the feature rows, rewards, flags, KPI rows and the
//         state arrays the phases mutated — each straight into the caller's buffer when that buffer is pinned,
//         else into the slot's pinned staging block, from where tl_host_wait scatters it with the context's threads.
// Steps of different slots run on different streams, so the upload of one overlaps the download of another.

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <thread>
#include <vector>

#if defined(__SSE2__) || defined(__x86_64__)
#include <emmintrin.h>
#define TL_HAVE_SSE2 1
#else
#define TL_HAVE_SSE2 0
#endif

namespace {

// ---- device: float32 map rows -> wire rows ------------------------------------------------------------------
// One thread per output word of 4 cells.  A cell that is not an integer in 0..255 raises *flag.
__global__ void __launch_bounds__(256) pack_map_u8_kernel(const float* __restrict__ map, uint32_t* __restrict__ wire, long long n_rows,
                                                          int row_elems, int wire_cells, int words_per_row, int* __restrict__ flag) {
    const long long total = n_rows * words_per_row;
    bool bad = false;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long row = idx / words_per_row;
        const int w = (int)(idx - row * words_per_row);
        const float* src = map + row * row_elems;
        uint32_t word = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = 4 * w + k;
            const float v = c < wire_cells ? __ldg(src + c) : 0.0f;
            const unsigned byte = (unsigned)fminf(fmaxf(v, 0.0f), 255.0f);
            bad |= !(v == (float)byte);
            word |= byte << (8 * k);
        }
        wire[idx] = word;
    }
    if (bad) atomicOr(flag, 1);
}

// One warp per output word of 32 cells (coalesced reads, ballot).  A cell other than 0 or 1 raises *flag.
__global__ void __launch_bounds__(256) pack_map_bits_kernel(const float* __restrict__ map, uint32_t* __restrict__ wire, long long n_rows,
                                                            int row_elems, int wire_cells, int words_per_row, int* __restrict__ flag) {
    const long long total = n_rows * words_per_row;
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    bool bad = false;
    for (long long idx = warp0; idx < total; idx += n_warps) {
        const long long row = idx / words_per_row;
        const int w = (int)(idx - row * words_per_row);
        const int c = 32 * w + lane;
        const float v = c < wire_cells ? __ldg(map + row * row_elems + c) : 0.0f;
        bad |= !(v == 0.0f || v == 1.0f);
        const unsigned word = __ballot_sync(0xffffffffu, v != 0.0f);
        if (lane == 0) wire[idx] = word;
    }
    if (bad) atomicOr(flag, 2);
}

// ---- host: wire rows -> float32 rows ------------------------------------------------------------------------
void expand_u8_cells(const uint8_t* src, float* dst, int n, bool stream) {
    int i = 0;
#if TL_HAVE_SSE2
    const bool aligned = stream && (((uintptr_t)dst) & 15) == 0;
    const __m128i zero = _mm_setzero_si128();
    for (; i + 16 <= n; i += 16) {
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
        const __m128i lo = _mm_unpacklo_epi8(b, zero), hi = _mm_unpackhi_epi8(b, zero);
        const __m128 f0 = _mm_cvtepi32_ps(_mm_unpacklo_epi16(lo, zero)), f1 = _mm_cvtepi32_ps(_mm_unpackhi_epi16(lo, zero));
        const __m128 f2 = _mm_cvtepi32_ps(_mm_unpacklo_epi16(hi, zero)), f3 = _mm_cvtepi32_ps(_mm_unpackhi_epi16(hi, zero));
        if (aligned) {  // streaming stores (optional): the rows are written once and read by someone else later
            _mm_stream_ps(dst + i, f0), _mm_stream_ps(dst + i + 4, f1), _mm_stream_ps(dst + i + 8, f2), _mm_stream_ps(dst + i + 12, f3);
        } else {
            _mm_storeu_ps(dst + i, f0), _mm_storeu_ps(dst + i + 4, f1), _mm_storeu_ps(dst + i + 8, f2), _mm_storeu_ps(dst + i + 12, f3);
        }
    }
#endif
    for (; i < n; ++i) dst[i] = (float)src[i];
}

struct BitTable {
    alignas(16) float f[256][8];
    BitTable() {
        for (int b = 0; b < 256; ++b)
            for (int k = 0; k < 8; ++k) f[b][k] = (b >> k) & 1 ? 1.0f : 0.0f;
    }
};
const BitTable& bit_table() {
    static const BitTable t;
    return t;
}

void expand_bit_cells(const uint8_t* src, float* dst, int n, bool stream) {  // cell i = bit i % 8 of byte i / 8 (little-endian words)
    const BitTable& t = bit_table();
    int i = 0;
#if TL_HAVE_SSE2
    const bool aligned = stream && (((uintptr_t)dst) & 15) == 0;
    for (; i + 8 <= n; i += 8) {
        const float* e = t.f[src[i >> 3]];
        const __m128 a = _mm_load_ps(e), b = _mm_load_ps(e + 4);
        if (aligned) _mm_stream_ps(dst + i, a), _mm_stream_ps(dst + i + 4, b);
        else _mm_storeu_ps(dst + i, a), _mm_storeu_ps(dst + i + 4, b);
    }
#else
    for (; i + 8 <= n; i += 8) memcpy(dst + i, t.f[src[i >> 3]], 32);
#endif
    for (; i < n; ++i) dst[i] = (src[i >> 3] >> (i & 7)) & 1 ? 1.0f : 0.0f;
}

// ---- a small pool: the caller and n - 1 workers share the tasks of one parallel region ------------------------
// Workers spin on the region counter for about a millisecond after a region before they go to sleep on the condition
// variable: in a pipelined loop the next region is ~0.1-0.3 ms away and a futex wake-up per worker would cost as much
// as the region itself; an idle context sleeps.
class Pool {
public:
    explicit Pool(int n_threads) : n_(n_threads < 1 ? 1 : n_threads) {
        for (int i = 1; i < n_; ++i) workers_.emplace_back([this] { work(); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lock(mu_);
            stop_.store(true);
            epoch_.fetch_add(1);
        }
        cv_.notify_all();
        for (auto& w : workers_) w.join();
    }
    int size() const { return n_; }
    void run(int n_tasks, const std::function<void(int)>& fn) {
        if (n_tasks <= 0) return;
        if (n_ == 1 || n_tasks == 1) {
            for (int i = 0; i < n_tasks; ++i) fn(i);
            return;
        }
        fn_ = &fn;
        n_tasks_ = n_tasks;
        next_.store(0);
        left_.store(n_tasks);
        open_.store(true);
        epoch_.fetch_add(1, std::memory_order_release);  // publishes fn_ / n_tasks_ / the counters
        if (sleepers_.load() > 0) {
            std::lock_guard<std::mutex> lock(mu_);
            cv_.notify_all();
        }
        drain();
        // the region ends when every task is done AND every worker that joined it has left drain(): a straggler
        // must not meet the counters of the next region
        while (left_.load(std::memory_order_acquire) > 0 || active_.load(std::memory_order_acquire) > 0) spin_pause();
        open_.store(false, std::memory_order_release);
        while (active_.load(std::memory_order_acquire) > 0) spin_pause();
    }

private:
    static void spin_pause() {
#if TL_HAVE_SSE2
        _mm_pause();
#else
        std::this_thread::yield();
#endif
    }
    void drain() {
        for (;;) {
            const int i = next_.fetch_add(1);
            if (i >= n_tasks_) return;
            (*fn_)(i);
            left_.fetch_sub(1, std::memory_order_release);
        }
    }
    void work() {
        uint64_t seen = 0;
        for (;;) {
            // wait for a new region: spin first, then sleep
            int spins = 0;
            while (epoch_.load(std::memory_order_acquire) == seen) {
                if (++spins < 200000) {
                    spin_pause();
                    continue;
                }
                std::unique_lock<std::mutex> lock(mu_);
                sleepers_.fetch_add(1);
                cv_.wait(lock, [&] { return epoch_.load(std::memory_order_acquire) != seen; });
                sleepers_.fetch_sub(1);
            }
            seen = epoch_.load(std::memory_order_acquire);
            if (stop_.load()) return;
            // join the region only while it is open: register, then re-check (run() closes before it reuses the counters)
            active_.fetch_add(1, std::memory_order_acq_rel);
            if (open_.load(std::memory_order_acquire) && epoch_.load(std::memory_order_acquire) == seen) drain();
            active_.fetch_sub(1, std::memory_order_release);
        }
    }
    int n_;
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_;
    const std::function<void(int)>* fn_ = nullptr;
    int n_tasks_ = 0;
    std::atomic<int> next_{0}, left_{0}, active_{0}, sleepers_{0};
    std::atomic<uint64_t> epoch_{0};
    std::atomic<bool> stop_{false}, open_{false};
};

bool ensure_device(unsigned char** buf, size_t* cap, size_t need) {
    if (need <= *cap) return true;
    if (*buf) cudaFree(*buf);
    *buf = nullptr;
    *cap = 0;
    const size_t want = need + need / 4 + 4096;
    if (!cuda_ok(cudaMalloc(reinterpret_cast<void**>(buf), want), "cudaMalloc")) return false;
    *cap = want;
    return true;
}

bool ensure_pinned(unsigned char** buf, size_t* cap, size_t need) {
    if (need <= *cap) return true;
    if (*buf) cudaFreeHost(*buf);
    *buf = nullptr;
    *cap = 0;
    const size_t want = need + need / 4 + 4096;
    if (!cuda_ok(cudaMallocHost(reinterpret_cast<void**>(buf), want), "cudaMallocHost")) return false;
    *cap = want;
    return true;
}

size_t up256(size_t v) { return (v + 255) / 256 * 256; }

// One device -> host transfer of a step: `ticks` blocks of `bytes` bytes, dense on the device, `host_stride` bytes
// apart on the host; direct when the host buffer is pinned, else through the slot's staging block.
struct Down {
    unsigned char* host;
    size_t host_stride, bytes, dev_off, stage_off;
    int ticks;
    bool direct;
};

enum MapJob { kMapNone = 0, kMapU8, kMapBits };

struct HostSlot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    unsigned char *d_arena = nullptr, *d_out = nullptr, *h_stage = nullptr, *h_up = nullptr;
    size_t arena_cap = 0, out_cap = 0, stage_cap = 0, up_cap = 0;
    // the resident batch: sizes and device offsets of the arrays as uploaded by the previous steps
    bool resident = false;
    int32_t r_sizes[5] = {0, 0, 0, 0, 0};
    size_t r_off[TL_N_FIELDS] = {0};
    size_t r_bytes[TL_N_FIELDS] = {0};
    size_t r_total = 0;
    // the pending step
    bool pending = false;
    std::vector<Down> downs;
    MapJob map_job = kMapNone;
    bool has_flag = false;              // a narrow wire pass ran: its range flag comes back in the staging block
    unsigned char* map_host = nullptr;  // caller's float32 map
    size_t map_host_stride = 0;         // bytes between ticks
    size_t map_stage_off = 0, map_wire_row = 0, flag_stage_off = 0;
    int map_ticks = 0, map_cells = 0, map_wire_cells = 0;
    long long map_agents = 0;
    const float* path_lut = nullptr;  // full variant: the constant channels the host fills in
    int path_cells = 0;
};

}  // namespace

struct TlContext {
    int device;
    HostSlot slots[TL_HOST_SLOTS];
    Pool* pool;
    int n_threads;
    int64_t last_h2d, last_d2h;
    std::unordered_map<const void*, bool> pinned;  // host pointer -> page-locked?  (a stale entry costs speed, never correctness)
};

namespace {

bool is_pinned(TlContext* ctx, const void* p) {
    if (!p) return false;
    auto it = ctx->pinned.find(p);
    if (it != ctx->pinned.end()) return it->second;
    cudaPointerAttributes attr;
    bool pinned = false;
    if (cudaPointerGetAttributes(&attr, p) == cudaSuccess) pinned = attr.type == cudaMemoryTypeHost;
    else cudaGetLastError();  // not a CUDA pointer: clear the sticky-free error
    if (ctx->pinned.size() > 4096) ctx->pinned.clear();
    ctx->pinned[p] = pinned;
    return pinned;
}

int wire_cells_of(const TlObsParams* obs) {  // integer channels only: the path channels of `full` stay on the host
    const int w2 = obs->window * obs->window;
    return (obs->variant == TL_VARIANT_FULL ? 4 : obs->n_channels) * w2;
}

int64_t wire_row_bytes(const TlObsParams* obs, int wire) {
    const int64_t cells = wire_cells_of(obs);
    switch (wire) {
        case TL_WIRE_F32: return (int64_t)obs->n_channels * obs->window * obs->window * 4;
        case TL_WIRE_U8: return (cells + 3) / 4 * 4;
        case TL_WIRE_BITS: return (cells + 31) / 32 * 4;
        default: return -1;
    }
}

}  // namespace

extern "C" {

TlContext* tl_context_create(int device) {
    if (!cuda_ok(cudaSetDevice(device), "cudaSetDevice")) return nullptr;
    TlContext* ctx = new TlContext();
    ctx->device = device;
    ctx->pool = nullptr;
    const unsigned hw = std::thread::hardware_concurrency();
    ctx->n_threads = hw == 0 ? 1 : (hw > 8 ? 8 : (int)hw);
    ctx->last_h2d = ctx->last_d2h = 0;
    for (HostSlot& s : ctx->slots) {
        if (!cuda_ok(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), "cudaStreamCreate") ||
            !cuda_ok(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming), "cudaEventCreate")) {
            tl_context_destroy(ctx);
            return nullptr;
        }
    }
    return ctx;
}

void tl_context_destroy(TlContext* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (HostSlot& s : ctx->slots) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        if (s.d_arena) cudaFree(s.d_arena);
        if (s.d_out) cudaFree(s.d_out);
        if (s.h_stage) cudaFreeHost(s.h_stage);
        if (s.h_up) cudaFreeHost(s.h_up);
        if (s.done) cudaEventDestroy(s.done);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    delete ctx->pool;
    delete ctx;
}

int tl_context_set_threads(TlContext* ctx, int32_t n_threads) {
    if (!ctx || n_threads < 1 || n_threads > 256) return set_error("tl_context_set_threads: 1..256 threads"), TL_ERR_ARG;
    for (const HostSlot& s : ctx->slots)
        if (s.pending) return set_error("tl_context_set_threads: a step is pending"), TL_ERR_ARG;
    if (ctx->pool && ctx->pool->size() != n_threads) {
        delete ctx->pool;
        ctx->pool = nullptr;
    }
    ctx->n_threads = n_threads;
    return TL_OK;
}

int64_t tl_wire_row_bytes(const TlObsParams* obs, int32_t wire) {
    if (!obs) return set_error("tl_wire_row_bytes: obs is NULL"), TL_ERR_ARG;
    const int64_t row = wire_row_bytes(obs, wire);
    if (row < 0) return set_error("unknown wire format"), TL_ERR_ARG;
    return row;
}

int64_t tl_context_last_h2d_bytes(const TlContext* ctx) { return ctx ? ctx->last_h2d : 0; }
int64_t tl_context_last_d2h_bytes(const TlContext* ctx) { return ctx ? ctx->last_d2h : 0; }

int tl_host_submit(TlContext* ctx, int32_t slot_index, const TlHostStep* st) {
    if (!ctx || !st) return set_error("tl_host_submit: NULL pointer"), TL_ERR_ARG;
    if (slot_index < 0 || slot_index >= TL_HOST_SLOTS) return set_error("tl_host_submit: slot out of range"), TL_ERR_ARG;
    HostSlot& s = ctx->slots[slot_index];
    if (s.pending) return set_error("tl_host_submit: the slot has a step pending (tl_host_wait first)"), TL_ERR_ARG;
    const TlTownBatch* hb = st->batch;
    const TlObsParams* obs = st->obs;
    const uint32_t phases = st->phases;
    const int n_ticks = st->n_ticks;
    int rc = validate(hb, obs, st->rew, st->dyn, st->obs_out, st->rew_out, phases, n_ticks);
    if (rc != TL_OK) return rc;
    if (st->wire < TL_WIRE_F32 || st->wire > TL_WIRE_BITS) return set_error("unknown wire format"), TL_ERR_ARG;
    if (st->map_format != TL_MAP_F32 && st->map_format != TL_MAP_WIRE) return set_error("unknown map format"), TL_ERR_ARG;
    if (!cuda_ok(cudaSetDevice(ctx->device), "cudaSetDevice")) return TL_ERR_CUDA;
    ctx->last_h2d = ctx->last_d2h = 0;
    s.downs.clear();
    s.map_job = kMapNone;
    s.has_flag = false;
    if (hb->n_towns == 0 || n_ticks == 0) return TL_OK;  // nothing pending: tl_host_wait returns at once

    const size_t T = hb->n_towns, N = hb->n_agents, E = hb->n_entities, K = hb->max_edges, EW = hb->embed_words;
    const bool do_obs = phases & TL_PHASE_OBS;
    const bool do_state = phases & (TL_PHASE_DECAY | TL_PHASE_REWARD);
    const bool do_adv = phases & TL_PHASE_ADVANCE;
    const bool do_act = phases & TL_PHASE_ACTIONS;       // rewrites positions, agent entity words, last-action fields
    const bool do_led = phases & TL_PHASE_SOCIAL_DECAY;  // rewrites the relationship and rivalry ledgers
    const bool do_res = phases & TL_PHASE_RESERVATIONS;  // rewrites the reservation slots and TL_F_RESERVATION
    const bool do_decay = phases & TL_PHASE_DECAY, do_reward = phases & TL_PHASE_REWARD;

    // ---- the batch arrays: host pointer, bytes, whether the phases mutate them -------------------------------
    TlTownBatch db = *hb;
    db.arena_base = nullptr;
    db.arena_bytes = 0;
    struct Field {
        const void* host;
        void** dev;
        size_t bytes;
        bool mutated;
    } f[TL_N_FIELDS];
#define FIELD(idx, name, nbytes, mut) f[idx] = Field{hb->name, (void**)&db.name, hb->name ? (size_t)(nbytes) : 0, (mut)}
    FIELD(TL_FIELD_TOWN_AGENT_OFF, town_agent_off, (T + 1) * 4, false);
    FIELD(TL_FIELD_TOWN_ENT_OFF, town_ent_off, (T + 1) * 4, false);
    FIELD(TL_FIELD_TOWN_TICK, town_tick, T * 4, do_adv);
    FIELD(TL_FIELD_ENT, ent, E * 4, do_act || do_res);
    FIELD(TL_FIELD_POS, pos, N * 4, do_act);
    FIELD(TL_FIELD_NEEDS, needs, 3 * N * 8, do_decay);
    FIELD(TL_FIELD_WALLET, wallet, N * 8, false);
    FIELD(TL_FIELD_ATTENDANCE, attendance, N * 8, false);
    FIELD(TL_FIELD_WAGES_WITHHELD, wages_withheld, N * 8, false);
    FIELD(TL_FIELD_WAGES_PAID, wages_paid, N * 8, false);
    FIELD(TL_FIELD_PUNCTUALITY, punctuality, N * 8, false);
    FIELD(TL_FIELD_EPISODE_TOTAL, episode_total, N * 8, do_reward);
    FIELD(TL_FIELD_SOCIAL_IN, social_in, 3 * N * 8, false);
    FIELD(TL_FIELD_TERM_BLOCK_TICK, term_block_tick, N * 4, do_reward);
    FIELD(TL_FIELD_LATENESS, lateness, N * 4, false);
    FIELD(TL_FIELD_LAST_ACTION_DURATION, last_action_duration, N * 4, do_act);
    FIELD(TL_FIELD_EPISODE_TICK, episode_tick, N * 4, do_adv);
    FIELD(TL_FIELD_SLOT, slot, N * 4, false);
    FIELD(TL_FIELD_FLAGS, flags, N * 4, do_state || do_act || do_obs || do_res);  // the observation consumes TL_F_CTX_RESET
    FIELD(TL_FIELD_LAST_ACTION_HASH, last_action_hash, N * 4, do_act);
    FIELD(TL_FIELD_PERSONALITY, personality, 3 * N * 4, false);
    FIELD(TL_FIELD_TIE_COUNT, tie_count, N, do_led);
    FIELD(TL_FIELD_TIE_EMBED, tie_embed, N * K * EW * 8, do_led);
    FIELD(TL_FIELD_TIE_TRUST, tie_trust, N * K * 8, do_led);
    FIELD(TL_FIELD_TIE_FAM, tie_fam, N * K * 8, do_led);
    FIELD(TL_FIELD_TIE_RIV, tie_riv, N * K * 8, do_led);
    FIELD(TL_FIELD_RIV_COUNT, riv_count, N, do_led);
    FIELD(TL_FIELD_RIV_EMBED, riv_embed, N * K * EW * 8, do_led);
    FIELD(TL_FIELD_RIV_SCORE, riv_score, N * K * 8, do_led);
    FIELD(TL_FIELD_ENT_HOLDER, ent_holder, E * 4, false);
#undef FIELD

    // One copy for the whole batch only when the caller DECLARES that its arrays share one block
    // (TlTownBatch.arena_base / arena_bytes); otherwise one copy per array and nothing else is touched.
    uintptr_t lo = 0, hi = 0;
    bool arena = false;
    if (hb->arena_base && hb->arena_bytes > 0) {
        lo = (uintptr_t)hb->arena_base;
        hi = lo + (size_t)hb->arena_bytes;
        for (int i = 0; i < TL_N_FIELDS; ++i) {
            if (!f[i].bytes) continue;
            const uintptr_t p0 = (uintptr_t)f[i].host;
            if (p0 < lo || p0 + f[i].bytes > hi)
                return set_error("an array of the batch lies outside [arena_base, arena_base + arena_bytes)"), TL_ERR_ARG;
        }
        arena = true;
    }
    size_t off[TL_N_FIELDS];
    size_t cursor = arena ? up256(hi - lo) : 0;
    for (int i = 0; i < TL_N_FIELDS; ++i) {
        if (arena) off[i] = f[i].bytes ? (uintptr_t)f[i].host - lo : 0;
        else {
            off[i] = cursor;
            cursor += up256(f[i].bytes);
        }
    }
    // action words of the launch, [n_ticks][N] with the caller's tick stride, and the look-up tables in one block
    TlDynamicsParams ddyn;
    memset(&ddyn, 0, sizeof(ddyn));
    if (st->dyn) ddyn = *st->dyn;
    const size_t act_bytes = do_act ? ((size_t)(n_ticks - 1) * (size_t)st->dyn->actions_tick_stride + N) * 4 : 0;
    const size_t act_off = cursor;
    cursor += up256(act_bytes);
    TlObsParams dobs;
    memset(&dobs, 0, sizeof(dobs));
    if (obs) dobs = *obs;
    struct Lut {
        const float* host;
        const float** dev;
        size_t bytes;
    } luts[8];
    int n_luts = 0;
    size_t lut_bytes = 0;
    if (do_obs) {
        const size_t W2 = (size_t)obs->window * obs->window, R = obs->window / 2;
        auto lut = [&](const float* host, const float** dev, size_t bytes) {
            if (!host) return;
            luts[n_luts++] = Lut{host, dev, bytes};
            lut_bytes += up256(bytes);
        };
        lut(obs->time_lut, &dobs.time_lut, (size_t)obs->ticks_per_day * 2 * 4);
        lut(obs->progress_lut, &dobs.progress_lut, ((size_t)obs->ticks_per_day + 1) * 4);
        lut(obs->slot_lut, &dobs.slot_lut, (size_t)obs->max_slots * 4);
        lut(obs->agent_ratio_lut, &dobs.agent_ratio_lut, W2 * 4);
        lut(obs->tile_ratio_lut, &dobs.tile_ratio_lut, (W2 + 1) * 4);
        lut(obs->nearest_lut, &dobs.nearest_lut, (2 * R * R + 1) * 4);
        lut(obs->path_lut, &dobs.path_lut, 2 * W2 * 4);
        lut(obs->embed_lut, &dobs.embed_lut, 256 * 4);
    }
    const size_t lut_off = cursor;
    cursor += lut_bytes;

    // ---- residency: which arrays go up ----------------------------------------------------------------------
    const int32_t sizes[5] = {hb->n_towns, hb->n_agents, hb->n_entities, hb->max_edges, hb->embed_words};
    bool same = s.resident && s.d_arena && cursor <= s.arena_cap && s.r_total == cursor && !memcmp(sizes, s.r_sizes, sizeof(sizes));
    for (int i = 0; same && i < TL_N_FIELDS; ++i) same = s.r_off[i] == off[i] && s.r_bytes[i] == f[i].bytes;
    const uint64_t dirty = same ? st->dirty_fields : TL_DIRTY_ALL;
    if (!same) {
        // a reallocation drops the resident copy, so wait for anything still using the old block
        if (cursor > s.arena_cap && !cuda_ok(cudaStreamSynchronize(s.stream), "cudaStreamSynchronize")) return TL_ERR_CUDA;
        if (!ensure_device(&s.d_arena, &s.arena_cap, cursor)) return TL_ERR_CUDA;
    }
    if (!ensure_pinned(&s.h_up, &s.up_cap, lut_bytes ? lut_bytes : 256)) return TL_ERR_CUDA;

    bool all_dirty = true;
    for (int i = 0; i < TL_N_FIELDS; ++i) all_dirty = all_dirty && (!f[i].bytes || ((dirty >> i) & 1));
    if (arena && all_dirty) {
        if (!cuda_ok(cudaMemcpyAsync(s.d_arena, (const void*)lo, hi - lo, cudaMemcpyHostToDevice, s.stream), "H2D arena")) return TL_ERR_CUDA;
        ctx->last_h2d += (int64_t)(hi - lo);
    } else if (arena) {
        // dirty arrays that sit next to each other in the declared arena (only alignment padding between them) go up
        // as one copy per run
        int order[TL_N_FIELDS], n_dirty = 0;
        for (int i = 0; i < TL_N_FIELDS; ++i)
            if (f[i].bytes && ((dirty >> i) & 1)) order[n_dirty++] = i;
        std::sort(order, order + n_dirty, [&](int a, int b) { return off[a] < off[b]; });
        for (int k = 0; k < n_dirty;) {
            const size_t begin = off[order[k]];
            size_t end = begin + f[order[k]].bytes;
            int j = k + 1;
            while (j < n_dirty && off[order[j]] <= end + 256) end = off[order[j]] + f[order[j]].bytes, ++j;
            if (!cuda_ok(cudaMemcpyAsync(s.d_arena + begin, (const void*)(lo + begin), end - begin, cudaMemcpyHostToDevice, s.stream), "H2D arrays"))
                return TL_ERR_CUDA;
            ctx->last_h2d += (int64_t)(end - begin);
            k = j;
        }
    } else {
        for (int i = 0; i < TL_N_FIELDS; ++i) {
            if (!f[i].bytes || !((dirty >> i) & 1)) continue;
            if (!cuda_ok(cudaMemcpyAsync(s.d_arena + off[i], f[i].host, f[i].bytes, cudaMemcpyHostToDevice, s.stream), "H2D array"))
                return TL_ERR_CUDA;
            ctx->last_h2d += (int64_t)f[i].bytes;
        }
    }
    for (int i = 0; i < TL_N_FIELDS; ++i) *f[i].dev = f[i].bytes ? (void*)(s.d_arena + off[i]) : nullptr;
    if (do_act) {
        if (!cuda_ok(cudaMemcpyAsync(s.d_arena + act_off, st->dyn->actions, act_bytes, cudaMemcpyHostToDevice, s.stream), "H2D actions"))
            return TL_ERR_CUDA;
        ctx->last_h2d += (int64_t)act_bytes;
        ddyn.actions = (const uint32_t*)(s.d_arena + act_off);
    }
    if (n_luts) {  // gathered on the host (a few KB), one copy up
        size_t o = 0;
        for (int i = 0; i < n_luts; ++i) {
            memcpy(s.h_up + o, luts[i].host, luts[i].bytes);
            *luts[i].dev = (const float*)(s.d_arena + lut_off + o);
            o += up256(luts[i].bytes);
        }
        if (!cuda_ok(cudaMemcpyAsync(s.d_arena + lut_off, s.h_up, lut_bytes, cudaMemcpyHostToDevice, s.stream), "H2D tables")) return TL_ERR_CUDA;
        ctx->last_h2d += (int64_t)lut_bytes;
    }
    s.resident = true;
    memcpy(s.r_sizes, sizes, sizeof(sizes));
    for (int i = 0; i < TL_N_FIELDS; ++i) s.r_off[i] = off[i], s.r_bytes[i] = f[i].bytes;
    s.r_total = cursor;

    // ---- device outputs (dense per tick) and where each goes on the host -----------------------------------------
    TlObsOut doo;
    TlRewardOut dro;
    memset(&doo, 0, sizeof(doo));
    memset(&dro, 0, sizeof(dro));
    const TlObsOut* hoo = st->obs_out;
    const TlRewardOut* hro = st->rew_out;
    const size_t nt = (size_t)n_ticks;
    size_t ocur = 0, scur = 0;
    auto dev_block = [&](size_t bytes) {
        const size_t o = ocur;
        ocur += up256(bytes);
        return o;
    };
    auto down = [&](void* host, size_t host_stride, size_t bytes_per_tick, size_t dev_off) {
        Down d;
        d.host = (unsigned char*)host;
        d.host_stride = host_stride;
        d.bytes = bytes_per_tick;
        d.dev_off = dev_off;
        d.ticks = n_ticks;
        d.direct = is_pinned(ctx, host);
        d.stage_off = scur;
        if (!d.direct) scur += up256(bytes_per_tick * nt);
        s.downs.push_back(d);
    };
    size_t o_map = 0, o_feat = 0, o_sum = 0, o_wire = 0, o_flag = 0;
    const bool narrow = do_obs && st->wire != TL_WIRE_F32;
    if (do_obs) {
        const size_t me = (size_t)obs->n_channels * obs->window * obs->window;
        const size_t row = (size_t)wire_row_bytes(obs, st->wire);
        o_map = dev_block(nt * N * me * 4);
        o_feat = dev_block(nt * N * (size_t)obs->n_features * 4);
        if (hoo->summary) o_sum = dev_block(nt * N * TL_N_SUMMARY * 4);
        if (narrow) {
            o_wire = dev_block(nt * N * row);
            o_flag = dev_block(256);
        }
        doo.map_tick_stride = (int64_t)(N * me);
        doo.features_tick_stride = (int64_t)(N * obs->n_features);
        doo.summary_tick_stride = (int64_t)(N * TL_N_SUMMARY);
        if (!narrow) {  // TL_MAP_WIRE buffers state their tick stride in bytes
            down(hoo->map, (size_t)hoo->map_tick_stride * (st->map_format == TL_MAP_WIRE ? 1 : 4), N * me * 4, o_map);
        } else if (st->map_format == TL_MAP_WIRE) {
            down(hoo->map, (size_t)hoo->map_tick_stride, N * row, o_wire);
        } else {  // narrow wire, float32 delivery: always staged, expanded by the host threads in tl_host_wait
            s.map_job = st->wire == TL_WIRE_U8 ? kMapU8 : kMapBits;
            s.map_host = (unsigned char*)hoo->map;
            s.map_host_stride = (size_t)hoo->map_tick_stride * 4;
            s.map_stage_off = scur;
            scur += up256(nt * N * row);
            s.map_wire_row = row;
            s.map_ticks = n_ticks;
            s.map_agents = (long long)N;
            s.map_cells = (int)me;
            s.map_wire_cells = wire_cells_of(obs);
            s.path_lut = obs->variant == TL_VARIANT_FULL ? obs->path_lut : nullptr;
            s.path_cells = obs->variant == TL_VARIANT_FULL ? 2 * obs->window * obs->window : 0;
        }
        if (narrow) {
            s.has_flag = true;
            s.flag_stage_off = scur;
            scur += 256;
        }
        down(hoo->features, (size_t)hoo->features_tick_stride * 4, N * (size_t)obs->n_features * 4, o_feat);
        if (hoo->summary) down(hoo->summary, (size_t)hoo->summary_tick_stride * 4, N * TL_N_SUMMARY * 4, o_sum);
    }
    size_t o_rew = 0, o_comps = 0, o_term = 0, o_kpi = 0;
    const bool want_ro = do_state && hro;
    if (want_ro) {
        dro.reward_tick_stride = (int64_t)N;
        dro.comps_tick_stride = (int64_t)(N * TL_N_COMPS);
        dro.terminated_tick_stride = (int64_t)N;
        dro.kpi_tick_stride = (int64_t)(T * TL_N_KPI);
        if (hro->reward) o_rew = dev_block(nt * N * 4), down(hro->reward, (size_t)hro->reward_tick_stride * 4, N * 4, o_rew);
        if (hro->comps) o_comps = dev_block(nt * N * TL_N_COMPS * 8), down(hro->comps, (size_t)hro->comps_tick_stride * 8, N * TL_N_COMPS * 8, o_comps);
        if (hro->terminated) o_term = dev_block(nt * N), down(hro->terminated, (size_t)hro->terminated_tick_stride, N, o_term);
        if (hro->kpi) o_kpi = dev_block(nt * T * TL_N_KPI * 4), down(hro->kpi, (size_t)hro->kpi_tick_stride * 4, T * TL_N_KPI * 4, o_kpi);
    }
    if (!ensure_device(&s.d_out, &s.out_cap, ocur ? ocur : 256)) return TL_ERR_CUDA;
    if (do_obs) {
        doo.map = (float*)(s.d_out + o_map);
        doo.features = (float*)(s.d_out + o_feat);
        doo.summary = hoo->summary ? (int32_t*)(s.d_out + o_sum) : nullptr;
    }
    if (want_ro) {
        dro.reward = hro->reward ? (float*)(s.d_out + o_rew) : nullptr;
        dro.comps = hro->comps ? (double*)(s.d_out + o_comps) : nullptr;
        dro.terminated = hro->terminated ? (uint8_t*)(s.d_out + o_term) : nullptr;
        dro.kpi = hro->kpi ? (float*)(s.d_out + o_kpi) : nullptr;
    }

    // ---- the tick launch, then the wire pass over its map ----------------------------------------------------------
    LaunchPlan lp;
    rc = prepare(&db, obs ? &dobs : nullptr, st->rew, st->dyn ? &ddyn : nullptr, do_obs ? &doo : nullptr, want_ro ? &dro : nullptr, phases,
                 n_ticks, 0, lp);
    if (rc == TL_OK) rc = issue(lp, 0, s.stream);
    if (rc != TL_OK) return rc;
    if (narrow) {
        const size_t me = (size_t)obs->n_channels * obs->window * obs->window;
        const int cells = wire_cells_of(obs);
        const size_t row = (size_t)wire_row_bytes(obs, st->wire);
        const int words = (int)(row / 4);
        const long long rows = (long long)(nt * N);
        int* flag = (int*)(s.d_out + o_flag);
        if (!cuda_ok(cudaMemsetAsync(flag, 0, 4, s.stream), "cudaMemsetAsync")) return TL_ERR_CUDA;
        const long long items = st->wire == TL_WIRE_U8 ? rows * words : rows * words * 32;
        long long blocks = (items + 255) / 256;
        if (blocks > 148LL * 32) blocks = 148LL * 32;
        if (st->wire == TL_WIRE_U8)
            pack_map_u8_kernel<<<(unsigned)blocks, 256, 0, s.stream>>>((const float*)(s.d_out + o_map), (uint32_t*)(s.d_out + o_wire), rows,
                                                                      (int)me, cells, words, flag);
        else
            pack_map_bits_kernel<<<(unsigned)blocks, 256, 0, s.stream>>>((const float*)(s.d_out + o_map), (uint32_t*)(s.d_out + o_wire), rows,
                                                                        (int)me, cells, words, flag);
        if (!cuda_ok(cudaGetLastError(), "wire pass launch")) return TL_ERR_CUDA;
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }

    // ---- downloads ------------------------------------------------------------------------------------------------
    if (!ensure_pinned(&s.h_stage, &s.stage_cap, scur ? scur : 256)) return TL_ERR_CUDA;
    auto copy_down = [&](void* dst, const void* src, size_t bytes) {
        ctx->last_d2h += (int64_t)bytes;
        return cuda_ok(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s.stream), "D2H");
    };
    for (const Down& d : s.downs) {
        if (!d.direct) {
            if (!copy_down(s.h_stage + d.stage_off, s.d_out + d.dev_off, d.bytes * nt)) return TL_ERR_CUDA;
        } else if (d.host_stride == d.bytes || n_ticks == 1) {
            if (!copy_down(d.host, s.d_out + d.dev_off, d.bytes * nt)) return TL_ERR_CUDA;
        } else {
            for (int t = 0; t < n_ticks; ++t)
                if (!copy_down(d.host + (size_t)t * d.host_stride, s.d_out + d.dev_off + (size_t)t * d.bytes, d.bytes)) return TL_ERR_CUDA;
        }
    }
    if (s.map_job != kMapNone && !copy_down(s.h_stage + s.map_stage_off, s.d_out + o_wire, nt * N * s.map_wire_row)) return TL_ERR_CUDA;
    if (s.has_flag && !copy_down(s.h_stage + s.flag_stage_off, s.d_out + o_flag, 4)) return TL_ERR_CUDA;
    // Mutated state goes back (unless the caller keeps the batch resident).  With a declared, pinned arena whose
    // mutated arrays sit together (the layout of soa.py puts them first) one copy covers them: the bytes in between
    // belong to the arena and are equal on both sides.
    if (st->writeback) {
        uintptr_t wlo = UINTPTR_MAX, whi = 0;
        size_t wsum = 0;
        for (int i = 0; i < TL_N_FIELDS; ++i) {
            if (!f[i].bytes || !f[i].mutated) continue;
            const uintptr_t p0 = (uintptr_t)f[i].host;
            wlo = p0 < wlo ? p0 : wlo;
            whi = p0 + f[i].bytes > whi ? p0 + f[i].bytes : whi;
            wsum += f[i].bytes;
        }
        const bool one_copy = arena && wsum > 0 && (whi - wlo) <= wsum + wsum / 2 + 4096 && is_pinned(ctx, (const void*)lo);
        if (one_copy) {
            if (!copy_down((void*)wlo, s.d_arena + (wlo - lo), whi - wlo)) return TL_ERR_CUDA;
        } else {
            for (int i = 0; i < TL_N_FIELDS; ++i) {
                if (!f[i].bytes || !f[i].mutated) continue;
                // straight from the device arena; a pageable destination makes this copy synchronous (still correct)
                if (!copy_down(const_cast<void*>(f[i].host), s.d_arena + off[i], f[i].bytes)) return TL_ERR_CUDA;
            }
        }
    }
    if (!cuda_ok(cudaEventRecord(s.done, s.stream), "cudaEventRecord")) return TL_ERR_CUDA;
    s.pending = true;
    return TL_OK;
}

int tl_host_wait(TlContext* ctx, int32_t slot_index) {
    if (!ctx) return set_error("context is NULL"), TL_ERR_ARG;
    if (slot_index < 0 || slot_index >= TL_HOST_SLOTS) return set_error("tl_host_wait: slot out of range"), TL_ERR_ARG;
    HostSlot& s = ctx->slots[slot_index];
    if (!s.pending) return TL_OK;
    s.pending = false;
    if (!cuda_ok(cudaEventSynchronize(s.done), "cudaEventSynchronize")) return TL_ERR_CUDA;
    if (s.has_flag && *reinterpret_cast<const int*>(s.h_stage + s.flag_stage_off) != 0)
        return set_error("a map cell does not fit the wire format (a count above 1 with TL_WIRE_BITS, above 255 with TL_WIRE_U8)"),
               TL_ERR_UNSUPPORTED;
    bool staged = s.map_job != kMapNone;
    for (const Down& d : s.downs) staged = staged || !d.direct;
    if (!staged) return TL_OK;
    if (!ctx->pool) ctx->pool = new Pool(ctx->n_threads);
    const int parts = ctx->pool->size() * 4;
    // staged plain outputs: parallel memcpy into the caller's (pageable) buffers
    for (const Down& d : s.downs) {
        if (d.direct) continue;
        const size_t total = d.bytes * (size_t)d.ticks;
        if (total < (1u << 16)) {
            for (int t = 0; t < d.ticks; ++t) memcpy(d.host + (size_t)t * d.host_stride, s.h_stage + d.stage_off + (size_t)t * d.bytes, d.bytes);
            continue;
        }
        const Down dd = d;
        const unsigned char* stage = s.h_stage;
        ctx->pool->run(parts * d.ticks, [dd, stage, parts](int job) {
            const int t = job / parts, k = job % parts;
            const size_t b0 = dd.bytes * (size_t)k / parts, b1 = dd.bytes * (size_t)(k + 1) / parts;
            memcpy(dd.host + (size_t)t * dd.host_stride + b0, stage + dd.stage_off + (size_t)t * dd.bytes + b0, b1 - b0);
        });
    }
    if (s.map_job != kMapNone) {
        const HostSlot* sp = &s;
        const long long rows = s.map_agents * s.map_ticks;
        const int jobs = (int)(rows < parts ? rows : parts);
        const bool stream = g_opt.stream_stores.load() != 0;
        ctx->pool->run(jobs, [sp, rows, jobs, stream](int job) {
            const long long r0 = rows * job / jobs, r1 = rows * (job + 1) / jobs;
            for (long long r = r0; r < r1; ++r) {
                const long long t = r / sp->map_agents, a = r - t * sp->map_agents;
                const unsigned char* src = sp->h_stage + sp->map_stage_off + (size_t)r * sp->map_wire_row;
                float* dst = reinterpret_cast<float*>(sp->map_host + (size_t)t * sp->map_host_stride) + (size_t)a * sp->map_cells;
                if (sp->map_job == kMapU8) expand_u8_cells(src, dst, sp->map_wire_cells, stream);
                else expand_bit_cells(src, dst, sp->map_wire_cells, stream);
                if (sp->path_cells) memcpy(dst + sp->map_wire_cells, sp->path_lut, (size_t)sp->path_cells * 4);  // map.py:115-125
            }
#if TL_HAVE_SSE2
            _mm_sfence();  // streaming stores are visible before the region ends
#endif
        });
    }
    return TL_OK;
}

int tl_host_tick_world(TlContext* ctx, const TlTownBatch* hb, const TlObsParams* obs, const TlRewardParams* rew,
                       const TlDynamicsParams* dyn, const TlObsOut* hoo, const TlRewardOut* hro, uint32_t phases,
                       int32_t n_ticks) {
    if (!ctx) return set_error("context is NULL"), TL_ERR_ARG;
    TlHostStep st;
    memset(&st, 0, sizeof(st));
    st.batch = hb;
    st.obs = obs;
    st.rew = rew;
    st.dyn = dyn;
    st.obs_out = hoo;
    st.rew_out = hro;
    st.phases = phases;
    st.n_ticks = n_ticks;
    st.dirty_fields = TL_DIRTY_ALL;
    st.wire = TL_WIRE_F32;
    st.map_format = TL_MAP_F32;
    st.writeback = 1;
    const int rc = tl_host_submit(ctx, 0, &st);
    return rc != TL_OK ? rc : tl_host_wait(ctx, 0);
}

int tl_host_tick(TlContext* ctx, const TlTownBatch* hb, const TlObsParams* obs, const TlRewardParams* rew,
                 const TlObsOut* hoo, const TlRewardOut* hro, uint32_t phases, int32_t n_ticks) {
    if (phases & (TL_PHASE_ACTIONS | TL_PHASE_SOCIAL_DECAY))
        return set_error("the world-dynamics phases need TlDynamicsParams (tl_host_tick_world)"), TL_ERR_ARG;
    return tl_host_tick_world(ctx, hb, obs, rew, nullptr, hoo, hro, phases, n_ticks);
}

}  // extern "C"
