// townlet_b200.cu — sm_100a kernels and C ABI for Townlet's per-tick agent hot path.
//
// One kernel template, tick_kernel<VARIANT>, runs the phases of one or more
// ticks for a group of towns per CTA, in SimulationLoop.step order
// (reference src/townlet/core/sim_loop.py:632): need decay (world/grid.py:1439)
// -> faint predicate (lifecycle/manager.py:39-45) -> reward
// (rewards/engine.py:34-178) -> observation build
// (world/observations/service.py:201 and encoders/{map,features,social}.py).
//
// Shape of a CTA (256 threads, 8 warps) per tick:
//   stage   every town of the CTA gets a packed uint16 tile grid in shared
//           memory (agents:8 | objects:7 | reserved:1), built from the town's
//           entity list with shared-memory atomics — the device form of
//           build_local_cache (world/observations/cache.py:16-41);
//   phase A one THREAD per agent, 32 agents per chunk: coalesced SoA loads,
//           float64 decay / reward / guardrails, per-town KPI partials by
//           segmented warp shuffles, scalar features into a shared row;
//   phase B one WARP per agent: window gather from the staged grid into a
//           channel-major staging tile, local summary by REDUX, social
//           snippet by warp-cooperative stable ranking, then 16-byte
//           vectorised coalesced stores of the map and the feature row.
// The path is gather / elementwise and HBM-write bound; tensor cores are not
// used.  There is no CPU fallback in this library.

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_bf16.h>

#include <atomic>
#include <mutex>
#include <type_traits>
#include <unordered_map>

#include "townlet_b200.h"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kChunk = 32;          // agents per phase-A pass
constexpr int kMaxTownsPerCta = 8;  // towns sharing one CTA (grid form)
constexpr int kMaxTownsPerWarp = 16;  // towns sharing one CTA (entity form)
constexpr int kEntCache = 4;  // entity words per lane kept in registers (128 entities per town)

thread_local char g_error[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, const char* detail = "") { snprintf(g_error, sizeof(g_error), fmt, detail); }

struct KParams {
    TlTownBatch b;
    TlObsParams obs;
    TlRewardParams rew;
    TlObsOut oo;
    TlRewardOut ro;
    TlDynamicsParams dyn;
    uint32_t phases;
    int32_t n_ticks;
    int32_t towns_per_cta;  // most towns a CTA owns
    int32_t n_big_ctas;     // role form: the first n_big_ctas CTAs own towns_per_cta towns, the rest one fewer
    int32_t grid_tiles;    // grid_w * grid_h
    int32_t feat_stride;   // floats per staged feature row (== n_features)
    int32_t map_stride;    // floats per warp map staging tile (multiple of 4, >= map_elems + 4)
    // shared memory offsets in bytes
    int32_t off_grid;      // uint16 [towns_per_cta][grid_tiles]
    int32_t off_ctype;     // uint16 [towns_per_cta][grid_tiles] (4 x 4-bit counters) or -1
    int32_t off_feat;      // float [kChunk * feat_stride + 4]
    int32_t off_map;       // float [kWarps][map_stride]
    int32_t off_meta;      // ChunkMeta[kChunk] (grid form) / KpiAcc[towns_per_cta] (entity form)
    int32_t smem_bytes;
    int32_t chunk;         // entity form: agents per phase-A pass (<= 32)
    int32_t mode;          // 0 entity-list form with role-parallel warps, 1 dense shared-memory grid form
    int32_t off_social;    // role form: staged relation rows of a pass (4 x double + 2 x embed words per edge)
    int32_t kpi_out_warp;  // role form: warp that writes the per-town KPI rows of a tick (0 state warp, 2 a tile warp)
    int32_t off_state;     // role form: staged per-agent inputs of the state role (kStateStageBytes)
    int32_t off_kpi_raw;   // role form: per-agent KPI terms of a tick, two parities x 32 lanes (KpiRaw), summed by the KPI-output warp
    int32_t tmpl_mask;     // entity form: bit m set = a tile template for 16-byte misalignment m floats exists
    int32_t tmpl_off[4];   // entity form: float offset of template m inside the template area
};

// State the world-dynamics phases may rewrite inside a launch is read with plain loads when they run
// (MUT) and through the read-only path otherwise.
template <bool MUT, typename T>
__device__ __forceinline__ T ld_state(const T* ptr) {
    if constexpr (MUT) return *ptr;
    else return __ldg(ptr);
}

// 8-byte asynchronous global -> shared copy (LDGSTS) and the wait for all copies of this thread
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ int pos_x(int32_t p) { return (int)(int16_t)(p & 0xffff); }
__device__ __forceinline__ int pos_y(int32_t p) { return (int)(int16_t)((uint32_t)p >> 16); }

__device__ __forceinline__ double scale_bias(double value, double factor) {
    // RewardEngine._apply_reward_biases._scale, rewards/engine.py:226-231
    if (factor <= 0.0) return value;
    if (value >= 0.0) return value * factor;
    return value / factor;
}

// ---------------------------------------------------------------------------
// Phase A, reward part: RewardEngine.compute body for one agent
// (rewards/engine.py:71-172).  All float64, as in the reference.
// ---------------------------------------------------------------------------
struct RewardResult {
    double total;
    double comps[TL_N_COMPS];
    double cumulative;  // new RewardEngine._episode_totals entry
    int32_t block;      // new RewardEngine._termination_block entry
};

// Per-agent reward inputs, loaded up front so the global loads overlap.
struct RewardIn {
    double sb, sp, sa;                        // social bonus / penalty / avoidance (engine.py:54-63)
    double wages_paid, wages_withheld, punctuality;
    double episode_total;
    int32_t block;
};

__device__ __forceinline__ RewardIn load_reward_in(const KParams& p, int a) {
    RewardIn in;
    const size_t N = (size_t)p.b.n_agents;
    in.sb = in.sp = in.sa = 0.0;
    if (p.b.social_in) {
        in.sb = __ldg(p.b.social_in + a);
        in.sp = __ldg(p.b.social_in + N + a);
        in.sa = __ldg(p.b.social_in + 2 * N + a);
    }
    in.wages_paid = __ldg(p.b.wages_paid + a);
    in.wages_withheld = __ldg(p.b.wages_withheld + a);
    in.punctuality = __ldg(p.b.punctuality + a);
    in.episode_total = p.b.episode_total[a];
    in.block = p.b.term_block_tick[a];
    return in;
}

// RewardEngine._apply_reward_biases (rewards/engine.py:220-255): personality factors on the social, employment and
// survival components; `delta` accumulates adjusted - base in the reference's order.  Out of line: only worlds with the
// reward_multipliers flag and a profile run it, and its five float64 divisions are a fifth of the kernel's code.
struct RewardBias {
    double sb, sp, sa, wage, punct, survival, delta;
};
__device__ __noinline__ RewardBias apply_reward_biases(RewardBias v, double f0, double f1, double f2) {
    double delta = 0.0;
    if (f0 != 1.0) {
        double adj = scale_bias(v.sb, f0);
        delta += adj - v.sb, v.sb = adj;
        adj = scale_bias(v.sp, f0);
        delta += adj - v.sp, v.sp = adj;
        adj = scale_bias(v.sa, f0);
        delta += adj - v.sa, v.sa = adj;
    }
    if (f1 != 1.0) {
        double adj = scale_bias(v.wage, f1);
        delta += adj - v.wage, v.wage = adj;
        adj = scale_bias(v.punct, f1);
        delta += adj - v.punct, v.punct = adj;
    }
    if (f2 != 1.0) {
        const double adj = scale_bias(v.survival, f2);
        delta += adj - v.survival, v.survival = adj;
    }
    v.delta = delta;
    return v;
}

__device__ __forceinline__ void reward_agent(const TlRewardParams& r, const RewardIn& in, int tick,
                                             const double (&need)[3], uint32_t fl, bool terminated, unsigned reason,
                                             RewardResult& out) {
    double* c = out.comps;
#pragma unroll
    for (int i = 0; i < TL_N_COMPS; ++i) c[i] = 0.0;
    double total = r.survival_tick;
    c[TL_COMP_SURVIVAL] = r.survival_tick;
    double needs_penalty = 0.0;
#pragma unroll
    for (int k = 0; k < TL_N_NEEDS; ++k) {  // engine.py:76-80
        double deficit = fmax(0.0, 1.0 - need[k]);
        needs_penalty += r.need_weight[k] * (deficit * deficit);
    }
    total -= needs_penalty;
    c[TL_COMP_NEEDS_PENALTY] = -needs_penalty;
    const double sb = in.sb, sp = in.sp, sa = in.sa;
    total += sb + sp + sa;  // engine.py:87
    c[TL_COMP_SOCIAL_BONUS] = sb;
    c[TL_COMP_SOCIAL_PENALTY] = sp;
    c[TL_COMP_SOCIAL_AVOIDANCE] = sa;
    c[TL_COMP_SOCIAL] = sb + sp + sa;
    double wage = 0.0;  // engine.py:375-386
    if (r.wage_rate > 0.0) wage = in.wages_paid - in.wages_withheld;
    total += wage;
    c[TL_COMP_WAGE] = wage;
    double punct = 0.0;  // engine.py:388-398
    if (r.punctuality_bonus > 0.0) punct = r.punctuality_bonus * in.punctuality;
    total += punct;
    c[TL_COMP_PUNCTUALITY] = punct;
    unsigned profile = (fl & TL_F_PROFILE_MASK) >> TL_F_PROFILE_SHIFT;
    if (r.reward_personality && profile != 0) {  // engine.py:104-106, 220-255
        const RewardBias b = apply_reward_biases(RewardBias{c[TL_COMP_SOCIAL_BONUS], c[TL_COMP_SOCIAL_PENALTY], c[TL_COMP_SOCIAL_AVOIDANCE],
                                                            c[TL_COMP_WAGE], c[TL_COMP_PUNCTUALITY], c[TL_COMP_SURVIVAL], 0.0},
                                                 r.reward_bias[profile][0], r.reward_bias[profile][1], r.reward_bias[profile][2]);
        c[TL_COMP_SOCIAL_BONUS] = b.sb, c[TL_COMP_SOCIAL_PENALTY] = b.sp, c[TL_COMP_SOCIAL_AVOIDANCE] = b.sa;
        c[TL_COMP_SOCIAL] = b.sb + b.sp + b.sa;  // unchanged values add up to the same sum
        c[TL_COMP_WAGE] = b.wage, c[TL_COMP_PUNCTUALITY] = b.punct, c[TL_COMP_SURVIVAL] = b.survival;
        total += b.delta;
    }
    double penalty = 0.0;  // engine.py:400-413
    if (terminated) {
        if (reason == 1) penalty = r.faint_penalty;
        else if (reason == 2) penalty = r.eviction_penalty;
    }
    total += penalty;
    c[TL_COMP_TERMINAL_PENALTY] = penalty;
    const double pre_guard = total;
    // termination block: prune (engine.py:363-373), set (:113-114), test (:355-361)
    int32_t block = in.block;
    if (r.block_window < 0) block = TL_TERM_BLOCK_NONE;
    else if (block != TL_TERM_BLOCK_NONE && (int64_t)tick - (int64_t)block > (int64_t)r.block_window)
        block = TL_TERM_BLOCK_NONE;
    if (terminated) block = tick;
    const bool blocked = r.block_window >= 0 && block != TL_TERM_BLOCK_NONE &&
                         (int64_t)tick - (int64_t)block <= (int64_t)r.block_window;
    out.block = block;
    if ((terminated || blocked) && total > 0.0) total = 0.0;  // engine.py:116-122
    const double guard_adjust = total - pre_guard;
    const double post = fmax(fmin(total, r.clip_per_tick), -r.clip_per_tick);  // engine.py:126-128
    const double tick_clip_adjust = post - total;
    total = post;
    const double prev = terminated ? 0.0 : in.episode_total;  // engine.py:415-424
    double cumulative = prev + total;
    double episode_clip_adjust = 0.0;
    if (r.clip_per_episode > 0) {  // engine.py:130-139
        cumulative = fmax(fmin(cumulative, r.clip_per_episode), -r.clip_per_episode);
        const double new_total = cumulative - prev;
        episode_clip_adjust = new_total - total;
        total = new_total;
    }
    out.cumulative = cumulative;
    c[TL_COMP_CLIP_ADJUSTMENT] = guard_adjust + tick_clip_adjust + episode_clip_adjust;
    c[TL_COMP_TOTAL] = total;
    c[TL_COMP_HOMEOSTASIS] = c[TL_COMP_SURVIVAL] + c[TL_COMP_NEEDS_PENALTY];
    c[TL_COMP_WORK] = c[TL_COMP_WAGE] + c[TL_COMP_PUNCTUALITY];
    c[TL_COMP_GUARDRAIL_ACTIVE] = guard_adjust != 0.0 ? 1.0 : 0.0;
    c[TL_COMP_CLIPPED] = (tick_clip_adjust != 0.0 || episode_clip_adjust != 0.0) ? 1.0 : 0.0;
    out.total = total;
}

// Segmented (by town slot, contiguous lanes) suffix reduction: after the call
// the first lane of every segment holds the segment's reduction.
template <typename T, typename Op>
__device__ __forceinline__ T seg_reduce(T v, int seg, int lane, Op op) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        T other = __shfl_down_sync(0xffffffffu, v, off);
        int oseg = __shfl_down_sync(0xffffffffu, seg, off);
        if (lane + off < 32 && oseg == seg) v = op(v, other);
    }
    return v;
}

struct KpiAcc {
    double rsum, rsq, hmin, hsum, late, wallet;
    int starving, term, count;
};

// KPI terms of ONE agent and tick, handed from the state role to the warp that writes the KPI rows
struct KpiRaw {
    double reward, hunger, wallet;
    int late, flags;  // flags: bit 0 starving, bit 1 terminated
};

#include "agent_roles.cuh"
#include "tick_grid.cuh"
#include "tick_roles.cuh"

// Backward GAE scan, one thread per agent, coalesced over the agent axis of the time-major rollout
// (reference: compute_gae, src/townlet/policy/backends/pytorch/ppo_utils.py:55-67).  float32 arithmetic
// in the reference's operation order; built without FMA contraction (-fmad=false).
__global__ void __launch_bounds__(256) gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                                                  const float* __restrict__ dones, float* __restrict__ advantages,
                                                  float* __restrict__ returns, int n_steps, int n_agents, int bootstrap,
                                                  float gamma, float gamma_lambda) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_agents) return;
    float gae = 0.0f;
    float next_value = bootstrap ? values[(size_t)n_steps * n_agents + a] : 0.0f;
    for (int t = n_steps - 1; t >= 0; --t) {
        const size_t i = (size_t)t * n_agents + a;
        const float v = values[i];
        const float mask = 1.0f - dones[i];
        const float nv = bootstrap ? next_value : v;  // without a bootstrap row the reference reuses value[t]
        const float delta = rewards[i] + gamma * nv * mask - v;
        gae = delta + gamma_lambda * mask * gae;
        advantages[i] = gae;
        returns[i] = gae + v;
        next_value = v;
    }
}

// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3) and key (k0, k1) -> four uint32.  Counter-based, so a draw is a
// pure function of (seed, row, draw index): no generator state in device memory, replayable inside a CUDA graph.
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ float head_value(const float* p) { return __ldg(p); }
__device__ __forceinline__ float head_value(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// The sampling step of the batched policy (the reference takes log_softmax of the logits and the log
// probability of the chosen action per agent at batch size 1, policy/runner.py:478-488): one thread per agent row reads
// the policy logits and the value of the fused heads product, takes log_softmax in float32, samples the categorical by
// the Gumbel-max trick (argmax_j logp_j - log(-log u_j), u_j uniform in (0, 1) from Philox) and writes the action index,
// its log probability and the value.  Draw index of action j of row r: 4 * (*counter + offset) + j / 4, lane j % 4.
template <typename T>
__global__ void __launch_bounds__(256) sample_actions_kernel(const T* __restrict__ heads, long long n_rows, int row, int n_actions,
                                                             unsigned long long seed, const long long* __restrict__ counter,
                                                             long long offset, long long* __restrict__ actions,
                                                             float* __restrict__ log_probs, float* __restrict__ values) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const T* h = heads + r * row;
    float x[TL_MAX_ACTIONS];
    float top = -INFINITY;
#pragma unroll
    for (int j = 0; j < TL_MAX_ACTIONS; ++j) {
        x[j] = j < n_actions ? head_value(h + j) : -INFINITY;
        top = fmaxf(top, x[j]);
    }
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < TL_MAX_ACTIONS; ++j) sum += j < n_actions ? expf(x[j] - top) : 0.0f;
    const float lse = top + logf(sum);
    const unsigned long long base = 4ull * (unsigned long long)((counter ? *counter : 0ll) + offset);
    int best = 0;
    float best_score = -INFINITY, best_logp = 0.0f;
#pragma unroll
    for (int d = 0; d < TL_MAX_ACTIONS / 4; ++d) {
        if (4 * d >= n_actions) break;
        const unsigned long long c = base + d;
        const uint4 bits = philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)r, (uint32_t)((unsigned long long)r >> 32),
                                         (uint32_t)seed, (uint32_t)(seed >> 32));
        const uint32_t w[4] = {bits.x, bits.y, bits.z, bits.w};
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const int j = 4 * d + l;
            if (j >= n_actions) break;
            const float u = ((float)(w[l] >> 9) + 0.5f) * 1.1920928955078125e-07f;  // (0, 1), exact in float32
            const float logp = x[j] - lse;
            const float score = logp - logf(-logf(u));
            if (score > best_score) best_score = score, best = j, best_logp = logp;  // first maximum, as argmax
        }
    }
    actions[r] = best;
    log_probs[r] = best_logp;
    if (values) values[r] = head_value(h + n_actions);
}

// ---------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------
// float32 rows -> bfloat16 rows padded with zeros to a width the tensor-core GEMMs like (a multiple of 64): the cast the
// policy step does anyway (autocast), fused with the padding that makes the reduction dimension aligned.  One block walks
// rows; a thread converts one pair of neighbours (round to nearest even, as torch's float -> bfloat16 cast).
__global__ void __launch_bounds__(256) cast_pad_bf16_kernel(const float* __restrict__ src, __nv_bfloat162* __restrict__ dst,
                                                            long long n_rows, int width, int padded, int vec4,
                                                            long long dst_stride) {
    if (vec4) {  // width and padded are multiples of 4 and both bases 16 / 8-byte aligned: 16-byte loads, 8-byte stores
        const int quads = padded >> 2, in_quads = width >> 2;
        for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {
            const float4* in = reinterpret_cast<const float4*>(src + row * width);
            uint2* out = reinterpret_cast<uint2*>(dst + row * (dst_stride >> 1));
            for (int j = threadIdx.x; j < quads; j += blockDim.x) {
                const float4 v = j < in_quads ? __ldg(in + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
                out[j] = make_uint2(*reinterpret_cast<const unsigned*>(&lo), *reinterpret_cast<const unsigned*>(&hi));
            }
        }
        return;
    }
    const int pairs = padded >> 1;
    for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {
        const float* in = src + row * width;
        __nv_bfloat162* out = dst + row * (dst_stride >> 1);
        for (int j = threadIdx.x; j < pairs; j += blockDim.x) {
            const int e = 2 * j;
            const float a = e < width ? __ldg(in + e) : 0.0f;
            const float b = e + 1 < width ? __ldg(in + e + 1) : 0.0f;
            out[j] = __floats2bfloat162_rn(a, b);
        }
    }
}

int round_up(int v, int m) { return (v + m - 1) / m * m; }
bool cuda_ok(cudaError_t e, const char* what);

int cast_pad_rows(const float* src, void* dst, int64_t n_rows, int32_t width, int32_t padded_width, int64_t dst_stride,
                  cudaStream_t stream) {
    if (n_rows == 0) return TL_OK;
    const int vec4 = width % 4 == 0 && padded_width % 4 == 0 && dst_stride % 4 == 0 && ((uintptr_t)src & 15) == 0 &&
                     ((uintptr_t)dst & 7) == 0;
    const int items = vec4 ? padded_width / 4 : padded_width / 2;
    const int threads = items >= 256 ? 256 : (items + 31) / 32 * 32;
    const long long blocks = n_rows < 148LL * 16 ? n_rows : 148LL * 16;
    cast_pad_bf16_kernel<<<(unsigned)blocks, threads, 0, stream>>>(src, (__nv_bfloat162*)dst, n_rows, width, padded_width, vec4,
                                                                   dst_stride);
    if (!cuda_ok(cudaGetLastError(), "cast_pad_bf16 launch")) return TL_ERR_CUDA;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return TL_OK;
}

bool cuda_ok(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    snprintf(g_error, sizeof(g_error), "%s: %s", what, cudaGetErrorString(e));
    return false;
}

int validate(const TlTownBatch* b, const TlObsParams* obs, const TlRewardParams* rew, const TlDynamicsParams* dyn,
             const TlObsOut* oo, const TlRewardOut* ro, uint32_t phases, int n_ticks) {
    if (!b) return set_error("batch is NULL"), TL_ERR_ARG;
    if (n_ticks < 0) return set_error("n_ticks < 0"), TL_ERR_ARG;
    if (b->n_towns < 0 || b->n_agents < 0 || b->n_entities < 0) return set_error("negative batch size"), TL_ERR_ARG;
    if (b->max_edges < 1 || b->max_edges > TL_MAX_EDGES) return set_error("max_edges out of range"), TL_ERR_ARG;
    if (b->embed_words < 1 || b->embed_words > 4) return set_error("embed_words out of range"), TL_ERR_ARG;
    if (b->grid_w < 1 || b->grid_h < 1 || b->grid_w > TL_MAX_GRID || b->grid_h > TL_MAX_GRID)
        return set_error("grid size out of range"), TL_ERR_ARG;
    const uint32_t known = TL_PHASE_DECAY | TL_PHASE_REWARD | TL_PHASE_OBS | TL_PHASE_ADVANCE | TL_PHASE_ACTIONS |
                           TL_PHASE_SOCIAL_DECAY | TL_PHASE_RESERVATIONS;
    if (!(phases & known)) return set_error("empty phase mask"), TL_ERR_ARG;
    if (phases & ~known) return set_error("unknown phase bit"), TL_ERR_ARG;
    if (phases & (TL_PHASE_ACTIONS | TL_PHASE_SOCIAL_DECAY)) {
        if (!dyn) return set_error("dynamics params are NULL (use tl_tick_world)"), TL_ERR_ARG;
        if ((phases & TL_PHASE_ACTIONS) && !dyn->actions) return set_error("action words are NULL"), TL_ERR_ARG;
    }
    if ((phases & TL_PHASE_RESERVATIONS) && !b->ent_holder) return set_error("ent_holder is NULL (TL_PHASE_RESERVATIONS)"), TL_ERR_ARG;
    if (phases & (TL_PHASE_DECAY | TL_PHASE_REWARD)) {
        if (!rew) return set_error("reward params are NULL"), TL_ERR_ARG;
    }
    if (phases & (TL_PHASE_OBS | TL_PHASE_ADVANCE)) {
        if (!obs) return set_error("observation params are NULL"), TL_ERR_ARG;
    }
    if (phases & TL_PHASE_OBS) {
        // (map may be NULL when map_bf16 is given and the kernel writes the mirror itself: checked once the form is known)
        if (!oo || !oo->features || (!oo->map && !oo->map_bf16)) return set_error("observation outputs are NULL"), TL_ERR_ARG;
        if (obs->variant < 0 || obs->variant > 2) return set_error("unknown variant"), TL_ERR_ARG;
        if (obs->window < 3 || obs->window > TL_MAX_WINDOW || !(obs->window & 1))
            return set_error("window must be odd and in 3..21"), TL_ERR_ARG;
        if (obs->n_ctypes < 0 || obs->n_ctypes > TL_MAX_CTYPES) return set_error("too many object channels"), TL_ERR_ARG;
        const int expect_c = obs->variant == TL_VARIANT_HYBRID ? 4 : obs->variant == TL_VARIANT_FULL ? 6 : 5 + obs->n_ctypes;
        if (obs->n_channels != expect_c) return set_error("n_channels does not match the variant"), TL_ERR_ARG;
        if (obs->top_friends < 0 || obs->top_rivals < 0 || obs->top_friends + obs->top_rivals > TL_MAX_SOCIAL_SLOTS)
            return set_error("social slots out of range"), TL_ERR_ARG;
        if (obs->embed_dim < 1 || obs->embed_dim > 8 * b->embed_words)
            return set_error("embed_dim exceeds the stored embedding words"), TL_ERR_ARG;
        if (obs->n_features < TL_N_BASE_FEATURES) return set_error("n_features too small"), TL_ERR_ARG;
        if (!obs->time_lut || !obs->progress_lut || !obs->slot_lut || !obs->agent_ratio_lut || !obs->tile_ratio_lut ||
            !obs->nearest_lut || !obs->embed_lut || (obs->variant == TL_VARIANT_FULL && !obs->path_lut))
            return set_error("a look-up table pointer is NULL"), TL_ERR_ARG;
        if (obs->ticks_per_day < 1 || obs->max_slots < 1) return set_error("ticks_per_day / max_slots < 1"), TL_ERR_ARG;
        auto mirror_ok = [](const uint16_t* base, int row, int width, int64_t tick_stride, int data) {
            if (width == 0) width = row;
            return !((uintptr_t)base & 15) && !(row & 7) && !(width & 7) && !(tick_stride & 7) && tick_stride >= 0 && width >= data &&
                   row >= width;
        };
        if (oo->map_bf16 && !mirror_ok(oo->map_bf16, oo->map_bf16_row, oo->map_bf16_width, oo->map_bf16_tick_stride,
                                       obs->n_channels * obs->window * obs->window))
            return set_error("map_bf16: 16-byte aligned base, row >= width >= C*W*W, all multiples of 8 elements"), TL_ERR_ARG;
        if (oo->features_bf16 && !mirror_ok(oo->features_bf16, oo->features_bf16_row, oo->features_bf16_width,
                                            oo->features_bf16_tick_stride, obs->n_features))
            return set_error("features_bf16: 16-byte aligned base, row >= width >= F, all multiples of 8 elements"), TL_ERR_ARG;
    }
    (void)ro;
    return TL_OK;
}

// Kernel-selection switches (tl_set_option): process-wide, read when a launch is PLANNED, never per launch and never
// from the environment.  Defaults are the product configuration; the parity tests flip them to reach every kernel form.
struct Options {
    std::atomic<int> kernel_form{-1};   // -1 by entity count, 0 entity-list form with role warps, 1 dense shared-memory grid form
    std::atomic<int> pass_agents{0};    // entity form: agents per pass (0 = 32, shrunk to what a CTA owns)
    std::atomic<int> towns_per_cta{0};  // entity form: 0 = chosen from the batch shape
    std::atomic<int> generic_window{0}; // 1 = runtime-window kernels instead of the unrolled specialisations
    std::atomic<int> mirror_pass{0};    // 1 = bfloat16 mirrors by a cast pass after the launch instead of inside the kernel
    std::atomic<int> even_grid{1};      // round small grids up to a multiple of the SM count
    std::atomic<int> stream_stores{0};  // host expander: non-temporal stores into the caller's float32 map
    std::atomic<int> policy_pair{1};    // tl_policy_step: CTA pairs (tcgen05 cta_group::2, the fastest form); 0 = single CTAs
    std::atomic<int> policy_multicast{0};  // tl_policy_step: clusters of two single-CTA pipelines sharing the weight stream (TMA multicast)
};
Options g_opt;

struct DeviceInfo {
    int max_smem;  // opt-in dynamic shared memory per block, minus room for the kernels' static shared memory
    int sms;
    bool ok;
};

// attributes of a device, queried once
const DeviceInfo* device_info(int device) {
    static DeviceInfo info[64];
    static std::once_flag once[64];
    if (device < 0 || device >= 64) return nullptr;
    std::call_once(once[device], [device]() {
        DeviceInfo& d = info[device];
        d.ok = cudaDeviceGetAttribute(&d.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) == cudaSuccess &&
               cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess;
        d.max_smem -= 1024;  // the opt-in limit covers static + dynamic
    });
    return info[device].ok ? &info[device] : nullptr;
}

using TickKernel = void (*)(const KParams);

// Everything a launch needs, resolved once: the kernel instantiation, its grid and shared-memory carve-up, and the
// parameter block.  Issuing it is one kernel launch (plus the cast passes of the forms without an in-kernel mirror).
struct LaunchPlan {
    KParams kp;
    TickKernel kernel;
    int grid, threads;
    int device;
    int map_elems;
    bool cast_after;  // bfloat16 mirrors come from cast passes after the launch
    bool empty;       // nothing to do (empty batch or zero ticks)
};

int prepare(const TlTownBatch* b, const TlObsParams* obs, const TlRewardParams* rew, const TlDynamicsParams* dyn,
            const TlObsOut* oo, const TlRewardOut* ro, uint32_t phases, int n_ticks, int towns_per_cta_override,
            LaunchPlan& lp) {
    memset(&lp, 0, sizeof(lp));
    lp.empty = true;
    // an empty batch is a valid no-op (its zero-length arrays may well be NULL pointers)
    if (b && b->n_towns == 0 && b->n_agents == 0 && n_ticks >= 0) return TL_OK;
    int rc = validate(b, obs, rew, dyn, oo, ro, phases, n_ticks);
    if (rc != TL_OK) return rc;
    if (b->n_towns == 0 || n_ticks == 0) return TL_OK;
    lp.empty = false;

    KParams& kp = lp.kp;
    kp.b = *b;
    if (obs) kp.obs = *obs;
    if (rew) kp.rew = *rew;
    if (oo) {
        kp.oo = *oo;
        if (!kp.oo.map_bf16_width) kp.oo.map_bf16_width = kp.oo.map_bf16_row;
        if (!kp.oo.features_bf16_width) kp.oo.features_bf16_width = kp.oo.features_bf16_row;
    }
    if (ro) kp.ro = *ro;
    if (dyn) kp.dyn = *dyn;
    kp.phases = phases;
    kp.n_ticks = n_ticks;
    const bool do_obs = phases & TL_PHASE_OBS;
    if (!obs) {
        kp.obs.window = 3;
        kp.obs.ticks_per_day = 1;
        kp.obs.n_features = TL_N_BASE_FEATURES;
        kp.obs.social_base = -1;
        kp.obs.landmark_base = -1;
        kp.obs.personality_base = -1;
    }

    int device = 0;
    if (!cuda_ok(cudaGetDevice(&device), "cudaGetDevice")) return TL_ERR_CUDA;
    const DeviceInfo* di = device_info(device);
    if (!di) return set_error("cudaDeviceGetAttribute failed"), TL_ERR_CUDA;
    lp.device = device;
    const int max_smem = di->max_smem, sms = di->sms;

    const int opt_form = g_opt.kernel_form.load(), opt_pass = g_opt.pass_agents.load(), opt_tpc = g_opt.towns_per_cta.load();
    const bool opt_generic = g_opt.generic_window.load() != 0, opt_mirror_pass = g_opt.mirror_pass.load() != 0;
    const bool opt_even = g_opt.even_grid.load() != 0;

    const int avg_agents = b->n_towns > 0 ? (b->n_agents + b->n_towns - 1) / b->n_towns : 1;
    const int avg_entities = b->n_towns > 0 ? (b->n_entities + b->n_towns - 1) / b->n_towns : 1;
    const int map_elems = do_obs ? kp.obs.n_channels * kp.obs.window * kp.obs.window : 0;
    lp.map_elems = map_elems;
    kp.feat_stride = kp.obs.n_features;
    kp.map_stride = round_up(map_elems + 4, 4);
    kp.grid_tiles = b->grid_w * b->grid_h;
    // The entity-list form costs O(entities of the town) per agent, the dense grid form O(window);
    // the former wins until a town holds several hundred entities (DESIGN.md "Kernel forms").
    kp.mode = opt_form >= 0 ? (opt_form ? 1 : 0) : (avg_entities > 384 ? 1 : 0);

    int grid = 0, threads = kThreads;
    if (kp.mode != 1) {
        threads = 32 * kRoleWarps;
        int chunk = opt_pass > 0 ? opt_pass : 32;
        if (chunk > 32) chunk = 32;
        kp.chunk = chunk;
        const bool tpc_fixed = towns_per_cta_override > 0 || opt_tpc > 0;
        int tpc = towns_per_cta_override > 0 ? towns_per_cta_override : (opt_tpc > 0 ? opt_tpc : chunk / (avg_agents > 0 ? avg_agents : 1));
        if (tpc < 1) tpc = 1;
        if (tpc > kMaxTownsPerWarp) tpc = kMaxTownsPerWarp;
        if (!tpc_fixed) {
            // small batches: prefer more, emptier CTAs over idle SMs (measured on configs[1]:
            // 2 towns per 4-warp CTA beat 1, 3 and 4)
            const int want = 3 * sms;
            const int floor_tpc = 2 * avg_agents <= chunk ? 2 : 1;
            while (tpc > floor_tpc && (b->n_towns + tpc - 1) / tpc < want) --tpc;
        }
        {
            // staging only needs to hold the agents one warp owns per pass
            const int need = tpc * avg_agents;
            if (opt_pass <= 0 && need < chunk) kp.chunk = chunk = need < 8 ? 8 : (need + 7) / 8 * 8;
        }
        int off = 0;
        kp.off_grid = 0;
        kp.off_ctype = -1;
        kp.off_feat = off;
        // two row buffers of ((chunk * F + 4 + 3) & ~3) floats
        off += do_obs ? 2 * round_up((chunk * kp.feat_stride + 4) * 4, 16) : 16;
        kp.off_map = off;
        kp.tmpl_mask = 0;
        if (do_obs) {
            // which 16-byte misalignments (in floats) can a tile of the output tensor have?  (closed under tick
            // slots: a plan may be issued at any multiple of the tick stride)
            const int base = (int)(((uintptr_t)kp.oo.map >> 2) & 3);
            const int sa = map_elems & 3, st = (int)(kp.oo.map_tick_stride & 3);
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j) kp.tmpl_mask |= 1 << ((base + i * sa + j * st) & 3);
            const bool static_hybrid = kp.obs.variant == TL_VARIANT_HYBRID && kp.obs.window == 11 && !opt_generic;
            if (static_hybrid && kp.tmpl_mask == 1) kp.tmpl_mask = 0;  // aligned zero tiles come from registers
            int toff = 0;
            for (int m = 0; m < 4; ++m) {
                kp.tmpl_off[m] = toff;
                if ((kp.tmpl_mask >> m) & 1) toff += kp.map_stride;
            }
            off += toff * 4;
        }
        kp.off_meta = off;
        off += round_up(2 * (int)sizeof(KpiAcc) * tpc, 16);  // two tick parities
        kp.kpi_out_warp = (do_obs && kp.chunk <= 16) ? 2 : 0;  // measured: +5 % on configs[1], -3 % on 30-agent passes
        kp.off_kpi_raw = off;
        off += kp.kpi_out_warp ? 2 * 32 * (int)sizeof(KpiRaw) : 0;
        kp.off_state = off;
        off += kStateStageBytes;
        kp.off_social = off;
        if (do_obs || (phases & TL_PHASE_SOCIAL_DECAY)) off += chunk * b->max_edges * (4 * 8 + 2 * 8 * b->embed_words) + 256 * 4;
        kp.smem_bytes = off;
        if (off > max_smem) return set_error("observation staging does not fit in shared memory"), TL_ERR_UNSUPPORTED;
        kp.towns_per_cta = tpc;
        grid = (b->n_towns + tpc - 1) / tpc;
        kp.n_big_ctas = grid;  // every CTA owns tpc towns (the last one possibly fewer)
        if (!tpc_fixed && tpc > 1) {
            // A grid that is not a multiple of the SM count leaves some SMs one CTA more than others for the whole
            // launch (the CTAs loop over every tick); round it up and hand the towns out one fewer to the last CTAs.
            const int even = (grid + sms - 1) / sms * sms;
            if (even != grid && even <= b->n_towns && even / sms <= 8 && opt_even) {
                grid = even;
                kp.towns_per_cta = b->n_towns / grid + 1;
                kp.n_big_ctas = b->n_towns % grid;
                if (kp.n_big_ctas == 0) kp.towns_per_cta -= 1, kp.n_big_ctas = grid;
            }
        }
    } else {
        // towns per CTA: aim at one full phase-A warp (32 agents) per CTA
        int tpc = towns_per_cta_override > 0 ? towns_per_cta_override : kChunk / (avg_agents > 0 ? avg_agents : 1);
        if (tpc < 1) tpc = 1;
        if (tpc > kMaxTownsPerCta) tpc = kMaxTownsPerCta;
        for (;;) {
            int off = 0;
            kp.off_grid = off;
            off += do_obs ? round_up(tpc * kp.grid_tiles * 2, 16) : 0;
            kp.off_ctype = -1;
            if (do_obs && kp.obs.variant == TL_VARIANT_COMPACT && kp.obs.n_ctypes > 0) {
                kp.off_ctype = off;
                off += round_up(tpc * kp.grid_tiles * 2, 16);
            }
            kp.off_feat = off;
            off += do_obs ? round_up((kChunk * kp.feat_stride + 4) * 4, 16) : 16;
            kp.off_map = off;
            off += do_obs ? kWarps * kp.map_stride * 4 : 0;
            kp.off_meta = off;
            off += (int)sizeof(ChunkMeta) * kChunk;
            kp.smem_bytes = off;
            if (off <= max_smem || tpc == 1) break;
            --tpc;
        }
        if (kp.smem_bytes > max_smem) return set_error("town grid does not fit in shared memory"), TL_ERR_UNSUPPORTED;
        kp.towns_per_cta = tpc;
        grid = (b->n_towns + tpc - 1) / tpc;
        kp.n_big_ctas = grid;
    }
    lp.grid = grid;
    lp.threads = threads;

    auto pick_kernel = [&](TickKernel kernel) -> int {
        // opt in to the device's full dynamic shared memory once per kernel and device, so that launches
        // issue no attribute call (legal inside CUDA-graph stream capture).  Keyed by the function address.
        static std::mutex mu;
        static std::unordered_map<const void*, uint64_t> opted_in;
        const uint64_t bit = 1ull << (device & 63);
        std::lock_guard<std::mutex> lock(mu);
        uint64_t& mask = opted_in[reinterpret_cast<const void*>(kernel)];
        if (!(mask & bit)) {
            if (!cuda_ok(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem), "cudaFuncSetAttribute"))
                return TL_ERR_CUDA;
            mask |= bit;
        }
        lp.kernel = kernel;
        return TL_OK;
    };
    // The bfloat16 mirror of the map: written by the tiles role itself in the static hybrid form (the mirror_pass
    // option keeps it out, for A/B runs), by one cast pass per tick after the launch in every other form.
    const bool want_mirror = do_obs && (oo->map_bf16 != nullptr || oo->features_bf16 != nullptr);
    bool mirror_in_kernel = false;
    auto dispatch = [&]() -> int {
        if (kp.mode == 0) {
            const int win = do_obs ? kp.obs.window : 0;
            const bool world = phases & (TL_PHASE_ACTIONS | TL_PHASE_SOCIAL_DECAY | TL_PHASE_RESERVATIONS);
            auto pick = [&](auto variant, auto window) -> int {
                if (world) return pick_kernel(tick_kernel_roles<decltype(variant)::value, decltype(window)::value, true>);
                return pick_kernel(tick_kernel_roles<decltype(variant)::value, decltype(window)::value, false>);
            };
            using std::integral_constant;
            switch (do_obs ? kp.obs.variant : TL_VARIANT_HYBRID) {
                case TL_VARIANT_HYBRID:
                    if (win == 11 && !opt_generic && want_mirror && !opt_mirror_pass) {
                        mirror_in_kernel = true;
                        if (world) return pick_kernel(tick_kernel_roles<TL_VARIANT_HYBRID, 11, true, true>);
                        return pick_kernel(tick_kernel_roles<TL_VARIANT_HYBRID, 11, false, true>);
                    }
                    if (win == 11 && !opt_generic) return pick(integral_constant<int, TL_VARIANT_HYBRID>{}, integral_constant<int, 11>{});
                    return pick(integral_constant<int, TL_VARIANT_HYBRID>{}, integral_constant<int, 0>{});
                case TL_VARIANT_FULL:
                    if (win == 11 && !opt_generic) return pick(integral_constant<int, TL_VARIANT_FULL>{}, integral_constant<int, 11>{});
                    return pick(integral_constant<int, TL_VARIANT_FULL>{}, integral_constant<int, 0>{});
                case TL_VARIANT_COMPACT:
                    if (win == 7 && kp.obs.n_ctypes == 0 && !opt_generic)
                        return pick(integral_constant<int, TL_VARIANT_COMPACT>{}, integral_constant<int, 7>{});
                    return pick(integral_constant<int, TL_VARIANT_COMPACT>{}, integral_constant<int, 0>{});
                default: return set_error("unknown variant"), TL_ERR_ARG;
            }
        }
        switch (do_obs ? kp.obs.variant : TL_VARIANT_HYBRID) {
            case TL_VARIANT_HYBRID: return pick_kernel(tick_kernel<TL_VARIANT_HYBRID>);
            case TL_VARIANT_FULL: return pick_kernel(tick_kernel<TL_VARIANT_FULL>);
            case TL_VARIANT_COMPACT: return pick_kernel(tick_kernel<TL_VARIANT_COMPACT>);
            default: return set_error("unknown variant"), TL_ERR_ARG;
        }
    };
    rc = dispatch();
    if (rc != TL_OK) return rc;
    if (do_obs && !oo->map && !mirror_in_kernel)
        return set_error("map may be NULL only where the kernel writes map_bf16 itself (entity form, hybrid, window 11)"), TL_ERR_UNSUPPORTED;
    lp.cast_after = want_mirror && !mirror_in_kernel;
    return TL_OK;
}

// Issue a planned launch with every output moved `slot` tick strides on (slot 0 = the pointers the plan was made with).
int issue(const LaunchPlan& lp, int64_t slot, cudaStream_t stream) {
    if (lp.empty) return TL_OK;
    const KParams* kpp = &lp.kp;
    KParams moved;
    if (slot != 0) {
        moved = lp.kp;
        TlObsOut& o = moved.oo;
        TlRewardOut& r = moved.ro;
        if (o.map) o.map += slot * o.map_tick_stride;
        if (o.features) o.features += slot * o.features_tick_stride;
        if (o.summary) o.summary += slot * o.summary_tick_stride;
        if (o.map_bf16) o.map_bf16 += slot * o.map_bf16_tick_stride;
        if (o.features_bf16) o.features_bf16 += slot * o.features_bf16_tick_stride;
        if (r.reward) r.reward += slot * r.reward_tick_stride;
        if (r.comps) r.comps += slot * r.comps_tick_stride;
        if (r.terminated) r.terminated += slot * r.terminated_tick_stride;
        if (r.kpi) r.kpi += slot * r.kpi_tick_stride;
        kpp = &moved;
    }
    const KParams& kp = *kpp;
    lp.kernel<<<lp.grid, lp.threads, kp.smem_bytes, stream>>>(kp);
    if (!cuda_ok(cudaGetLastError(), "kernel launch")) return TL_ERR_CUDA;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!lp.cast_after) return TL_OK;
    int rc = TL_OK;
    for (int it = 0; it < kp.n_ticks; ++it) {
        if (kp.oo.map_bf16)
            rc = cast_pad_rows(kp.oo.map + (size_t)it * kp.oo.map_tick_stride, kp.oo.map_bf16 + (size_t)it * kp.oo.map_bf16_tick_stride,
                               kp.b.n_agents, lp.map_elems, kp.oo.map_bf16_width, kp.oo.map_bf16_row, stream);
        if (rc == TL_OK && kp.oo.features_bf16)
            rc = cast_pad_rows(kp.oo.features + (size_t)it * kp.oo.features_tick_stride,
                               kp.oo.features_bf16 + (size_t)it * kp.oo.features_bf16_tick_stride, kp.b.n_agents, kp.obs.n_features,
                               kp.oo.features_bf16_width, kp.oo.features_bf16_row, stream);
        if (rc != TL_OK) return rc;
    }
    return TL_OK;
}

int launch(const TlTownBatch* b, const TlObsParams* obs, const TlRewardParams* rew, const TlDynamicsParams* dyn,
           const TlObsOut* oo, const TlRewardOut* ro, uint32_t phases, int n_ticks, cudaStream_t stream,
           int towns_per_cta_override) {
    LaunchPlan lp;
    const int rc = prepare(b, obs, rew, dyn, oo, ro, phases, n_ticks, towns_per_cta_override, lp);
    return rc != TL_OK ? rc : issue(lp, 0, stream);
}

}  // namespace

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

const char* tl_last_error(void) { return g_error; }
int tl_abi_version(void) { return TL_ABI_VERSION; }

int tl_abi_struct_size(int idx) {
    switch (idx) {
        case 0: return (int)sizeof(TlTownBatch);
        case 1: return (int)sizeof(TlObsParams);
        case 2: return (int)sizeof(TlRewardParams);
        case 3: return (int)sizeof(TlObsOut);
        case 4: return (int)sizeof(TlRewardOut);
        case 5: return (int)sizeof(TlDynamicsParams);
        case 6: return (int)sizeof(TlHostStep);
        case 7: return (int)sizeof(TlPolicyWeights);
        case 8: return (int)sizeof(TlEmploymentParams);
        case 9: return (int)sizeof(TlEmploymentState);
        default: return -1;
    }
}

int tl_device_sm_count(void) {
    int device = 0, sms = 0;
    if (!cuda_ok(cudaGetDevice(&device), "cudaGetDevice")) return TL_ERR_CUDA;
    if (!cuda_ok(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device), "cudaDeviceGetAttribute"))
        return TL_ERR_CUDA;
    return sms;
}

int64_t tl_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int tl_gae(const float* rewards, const float* values, const float* dones, float* advantages, float* returns,
           int32_t n_steps, int32_t n_agents, int32_t bootstrap, double gamma, double gae_lambda, tl_stream_t stream) {
    if (!rewards || !values || !dones || !advantages || !returns) return set_error("tl_gae: NULL pointer"), TL_ERR_ARG;
    if (n_steps < 0 || n_agents < 0) return set_error("tl_gae: negative size"), TL_ERR_ARG;
    if (n_steps == 0 || n_agents == 0) return TL_OK;
    // the reference multiplies float32 tensors by Python floats: the scalars gamma and
    // gamma * gae_lambda (a float64 product) are rounded to float32 once
    const float g = (float)gamma, gl = (float)(gamma * gae_lambda);
    gae_kernel<<<(n_agents + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rewards, values, dones, advantages, returns, n_steps,
                                                                        n_agents, bootstrap ? 1 : 0, g, gl);
    if (!cuda_ok(cudaGetLastError(), "gae kernel launch")) return TL_ERR_CUDA;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return TL_OK;
}

int tl_sample_actions(const void* heads, int32_t heads_dtype, int64_t n_rows, int32_t row, int32_t n_actions, uint64_t seed,
                      const int64_t* counter, int64_t offset, int64_t* actions, float* log_probs, float* values,
                      tl_stream_t stream) {
    if (!heads || !actions || !log_probs) return set_error("tl_sample_actions: NULL pointer"), TL_ERR_ARG;
    if (heads_dtype != TL_DTYPE_F32 && heads_dtype != TL_DTYPE_BF16) return set_error("tl_sample_actions: unknown dtype"), TL_ERR_ARG;
    if (n_rows < 0 || n_actions < 1 || n_actions > TL_MAX_ACTIONS || row < n_actions + (values ? 1 : 0) || offset < 0)
        return set_error("tl_sample_actions: 1 <= n_actions <= 16, row >= n_actions (+ 1 with values), offset >= 0"), TL_ERR_ARG;
    if (n_rows == 0) return TL_OK;
    const unsigned blocks = (unsigned)((n_rows + 255) / 256);
    if (heads_dtype == TL_DTYPE_BF16)
        sample_actions_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)heads, n_rows, row, n_actions, seed,
                                                                      (const long long*)counter, offset, (long long*)actions,
                                                                      log_probs, values);
    else
        sample_actions_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)heads, n_rows, row, n_actions, seed,
                                                                      (const long long*)counter, offset, (long long*)actions,
                                                                      log_probs, values);
    if (!cuda_ok(cudaGetLastError(), "sample_actions launch")) return TL_ERR_CUDA;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return TL_OK;
}

int tl_cast_pad_bf16(const float* src, void* dst, int64_t n_rows, int32_t width, int32_t padded_width, tl_stream_t stream) {
    if (!src || !dst) return set_error("tl_cast_pad_bf16: NULL pointer"), TL_ERR_ARG;
    if (n_rows < 0 || width < 1 || padded_width < width || (padded_width & 1))
        return set_error("tl_cast_pad_bf16: padded_width must be even and >= width"), TL_ERR_ARG;
    return cast_pad_rows(src, dst, n_rows, width, padded_width, padded_width, (cudaStream_t)stream);
}

int tl_cast_pad_bf16_strided(const float* src, void* dst, int64_t n_rows, int32_t width, int32_t padded_width,
                             int64_t dst_stride, tl_stream_t stream) {
    if (!src || !dst) return set_error("tl_cast_pad_bf16: NULL pointer"), TL_ERR_ARG;
    if (n_rows < 0 || width < 1 || padded_width < width || (padded_width & 1) || dst_stride < padded_width || (dst_stride & 1))
        return set_error("tl_cast_pad_bf16: padded_width and dst_stride must be even, dst_stride >= padded_width >= width"), TL_ERR_ARG;
    return cast_pad_rows(src, dst, n_rows, width, padded_width, dst_stride, (cudaStream_t)stream);
}

int tl_obs_build(const TlTownBatch* batch, const TlObsParams* obs, const TlObsOut* out, tl_stream_t stream) {
    return launch(batch, obs, nullptr, nullptr, out, nullptr, TL_PHASE_OBS, 1, (cudaStream_t)stream, 0);
}

int tl_decay_reward(const TlTownBatch* batch, const TlRewardParams* rew, const TlRewardOut* out, uint32_t phases,
                    tl_stream_t stream) {
    if (phases & ~(TL_PHASE_DECAY | TL_PHASE_REWARD)) return set_error("tl_decay_reward accepts DECAY | REWARD only"), TL_ERR_ARG;
    return launch(batch, nullptr, rew, nullptr, nullptr, out, phases, 1, (cudaStream_t)stream, 0);
}

int tl_tick(const TlTownBatch* batch, const TlObsParams* obs, const TlRewardParams* rew, const TlObsOut* obs_out,
            const TlRewardOut* rew_out, uint32_t phases, int32_t n_ticks, tl_stream_t stream) {
    return launch(batch, obs, rew, nullptr, obs_out, rew_out, phases, n_ticks, (cudaStream_t)stream, 0);
}

int tl_tick_world(const TlTownBatch* batch, const TlObsParams* obs, const TlRewardParams* rew,
                  const TlDynamicsParams* dyn, const TlObsOut* obs_out, const TlRewardOut* rew_out, uint32_t phases,
                  int32_t n_ticks, tl_stream_t stream) {
    return launch(batch, obs, rew, dyn, obs_out, rew_out, phases, n_ticks, (cudaStream_t)stream, 0);
}

// ---- kernel-selection switches --------------------------------------------------
int tl_set_option(const char* name, int32_t value) {
    if (!name) return set_error("tl_set_option: NULL name"), TL_ERR_ARG;
    std::atomic<int>* slot = nullptr;
    if (!strcmp(name, "kernel_form")) slot = &g_opt.kernel_form;
    else if (!strcmp(name, "pass_agents")) slot = &g_opt.pass_agents;
    else if (!strcmp(name, "towns_per_cta")) slot = &g_opt.towns_per_cta;
    else if (!strcmp(name, "generic_window")) slot = &g_opt.generic_window;
    else if (!strcmp(name, "mirror_pass")) slot = &g_opt.mirror_pass;
    else if (!strcmp(name, "even_grid")) slot = &g_opt.even_grid;
    else if (!strcmp(name, "stream_stores")) slot = &g_opt.stream_stores;
    else if (!strcmp(name, "policy_pair")) slot = &g_opt.policy_pair;
    else if (!strcmp(name, "policy_multicast")) slot = &g_opt.policy_multicast;
    else if (!strcmp(name, "policy_multicast")) slot = &g_opt.policy_multicast;
    if (!slot) return set_error("tl_set_option: unknown option '%s'", name), TL_ERR_ARG;
    if (slot == &g_opt.kernel_form && (value < -1 || value > 1)) return set_error("kernel_form is -1, 0 or 1"), TL_ERR_ARG;
    if (slot == &g_opt.pass_agents && (value < 0 || value > 32)) return set_error("pass_agents is 0..32"), TL_ERR_ARG;
    if (slot == &g_opt.towns_per_cta && (value < 0 || value > kMaxTownsPerWarp)) return set_error("towns_per_cta is 0..16"), TL_ERR_ARG;
    slot->store(value);
    return TL_OK;
}

void tl_reset_options(void) {
    g_opt.kernel_form.store(-1);
    g_opt.pass_agents.store(0);
    g_opt.towns_per_cta.store(0);
    g_opt.generic_window.store(0);
    g_opt.mirror_pass.store(0);
    g_opt.even_grid.store(1);
    g_opt.stream_stores.store(0);
    g_opt.policy_pair.store(1);
    g_opt.policy_multicast.store(0);
}

// ---- planned launches -----------------------------------------------------------
struct TlPlan {
    LaunchPlan lp;
    cudaEvent_t* events;  // n_events timing events of the last tl_plan_launch_many(..., timed)
    int n_events, cap_events;
};

TlPlan* tl_plan_create(const TlTownBatch* batch, const TlObsParams* obs, const TlRewardParams* rew, const TlDynamicsParams* dyn,
                       const TlObsOut* obs_out, const TlRewardOut* rew_out, uint32_t phases, int32_t n_ticks) {
    TlPlan* plan = new TlPlan();
    plan->events = nullptr;
    plan->n_events = plan->cap_events = 0;
    if (prepare(batch, obs, rew, dyn, obs_out, rew_out, phases, n_ticks, 0, plan->lp) != TL_OK) {
        delete plan;
        return nullptr;
    }
    return plan;
}

void tl_plan_destroy(TlPlan* plan) {
    if (!plan) return;
    for (int i = 0; i < plan->cap_events; ++i) cudaEventDestroy(plan->events[i]);
    free(plan->events);
    delete plan;
}

int tl_plan_launch(const TlPlan* plan, int64_t tick_slot, tl_stream_t stream) {
    if (!plan) return set_error("plan is NULL"), TL_ERR_ARG;
    if (tick_slot < 0) return set_error("tick_slot < 0"), TL_ERR_ARG;
    return issue(plan->lp, tick_slot, (cudaStream_t)stream);
}

int tl_plan_launch_many(TlPlan* plan, int64_t slot0, int32_t n_launches, int64_t ring_slots, int32_t timed, tl_stream_t stream) {
    if (!plan) return set_error("plan is NULL"), TL_ERR_ARG;
    const int64_t per = plan->lp.empty ? 1 : plan->lp.kp.n_ticks;
    if (slot0 < 0 || n_launches < 0 || ring_slots < per) return set_error("tl_plan_launch_many: slot0 >= 0, ring_slots >= n_ticks"), TL_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const int want = timed ? n_launches + 1 : 0;
    if (want > plan->cap_events) {
        cudaEvent_t* grown = (cudaEvent_t*)realloc(plan->events, sizeof(cudaEvent_t) * want);
        if (!grown) return set_error("out of memory"), TL_ERR_ARG;
        plan->events = grown;
        for (int i = plan->cap_events; i < want; ++i) {
            if (!cuda_ok(cudaEventCreate(&plan->events[i]), "cudaEventCreate")) return TL_ERR_CUDA;
            plan->cap_events = i + 1;
        }
    }
    plan->n_events = want;
    const int64_t n_pos = ring_slots / per;  // launches that fit the ring before it wraps
    int64_t pos = (slot0 / per) % n_pos;
    if (timed && !cuda_ok(cudaEventRecord(plan->events[0], s), "cudaEventRecord")) return TL_ERR_CUDA;
    for (int i = 0; i < n_launches; ++i) {
        const int rc = issue(plan->lp, pos * per, s);
        if (rc != TL_OK) return rc;
        if (timed && !cuda_ok(cudaEventRecord(plan->events[i + 1], s), "cudaEventRecord")) return TL_ERR_CUDA;
        if (++pos == n_pos) pos = 0;
    }
    return TL_OK;
}

int tl_plan_launch_ms(const TlPlan* plan, float* ms, int32_t n) {
    if (!plan || !ms) return set_error("tl_plan_launch_ms: NULL pointer"), TL_ERR_ARG;
    if (n < 0 || n + 1 > plan->n_events) return set_error("tl_plan_launch_ms: more launches than the last timed call issued"), TL_ERR_ARG;
    for (int i = 0; i < n; ++i)
        if (!cuda_ok(cudaEventElapsedTime(&ms[i], plan->events[i], plan->events[i + 1]), "cudaEventElapsedTime")) return TL_ERR_CUDA;
    return TL_OK;
}

int tl_plan_grid(const TlPlan* plan, int32_t* grid, int32_t* threads, int32_t* smem_bytes) {
    if (!plan) return set_error("plan is NULL"), TL_ERR_ARG;
    if (grid) *grid = plan->lp.grid;
    if (threads) *threads = plan->lp.threads;
    if (smem_bytes) *smem_bytes = plan->lp.empty ? 0 : plan->lp.kp.smem_bytes;
    return TL_OK;
}

}  // extern "C"

#include "host_path.cuh"
#include "policy_mlp.cuh"
#include "employment.cuh"
