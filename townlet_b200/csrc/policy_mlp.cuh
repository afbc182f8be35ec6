// policy_mlp.cuh — the batched policy step of the rollout as ONE sm_100a kernel
// (included at the end of townlet_b200.cu; uses cuda_ok, set_error, g_launches, philox4x32_10).
//
// The reference runs ConflictAwarePolicyNetwork (src/townlet/policy/models.py:36-69) per agent at batch size 1
// (policy/runner.py:444-488): map encoder Linear(C*W*W, 256) + ReLU, feature encoder Linear(F, 256) + ReLU, shared
// Linear(512, 256) + ReLU, policy head Linear(256, A) and value head Linear(256, 1), then log_softmax and a categorical
// sample.  Here one CTA owns a tile of 128 agents and chains the four products on the 5th-generation tensor cores:
//
//   X[128 x 640] bf16 (the padded rows the tick kernel writes: 512 map columns | 128 feature columns)
//     --TMA-->  shared memory (128-byte swizzle)  --tcgen05.mma-->  TMEM  D1[128 x 512] fp32   (two encoders)
//     --tcgen05.ld, + bias, ReLU, bf16-->  shared memory H1[128 x 512] (written in the operand's swizzled layout)
//     --tcgen05.mma (weights by TMA)-->  TMEM  D2[128 x 256]  --+ bias, ReLU, bf16-->  H2[128 x 256]
//     --tcgen05.mma-->  TMEM  D3[128 x 16]  --+ bias-->  log_softmax, Gumbel-max sample (Philox), action / log-prob / value
//
// The activations never leave the SM (the three-GEMM library path moves ~130 MB of them through HBM per tick at 65 536
// agents) and the two encoders are two accumulations instead of one block-diagonal product with half its FLOPs wasted.
// Warp roles: warp 0 issues the TMA loads, warp 1 the MMAs (one elected lane), warps 2-9 are the epilogue (TMEM lane
// quarter = warp index mod 4, two warps per quarter splitting a layer's columns).  CTAs are persistent: each walks the
// tiles blockIdx.x, blockIdx.x + gridDim.x, ... so that tensor-memory allocation, barrier set-up and descriptor fetches
// are paid once, and the schedule is software-pipelined over the CTA's tiles (the tensor pipe works on tile i + 1 while the
// epilogue converts and samples tile i; accumulator columns ping-pong with the tile's parity).  Arithmetic: bfloat16
// operands, float32 accumulation, bias in float32, ReLU on the rounded bfloat16 pair (same result), hidden activations
// rounded to bfloat16 — the arithmetic of the bfloat16 autocast path it replaces.
// Three forms (tl_set_option): policy_mlp_kernel (default), policy_mlp_mc_kernel (clusters of two sharing the weight stream
// by TMA multicast), policy_mlp_pair_kernel (CTA pairs, tcgen05 cta_group::2).  DESIGN.md 5b has what each measures and why.

#include <cuda.h>
#include <cudaTypedefs.h>

namespace {

// -DTL_POLICY_TIMING: CTA 0 prints the cycles its first epilogue warp spends per tile waiting for each accumulator and
// converting it, and the MMA-issuing lane what it waits for (scripts/policy_phase_timing.py with an instrumented build; never
// part of the shipped library)
#ifdef TL_POLICY_TIMING
#define POL_TIMING(...) __VA_ARGS__
#else
#define POL_TIMING(...)
#endif

constexpr int kPolRows = 128;        // agents per CTA (UMMA M)
constexpr int kPolHidden = 256;      // hidden width (UMMA N of the encoder / shared products)
constexpr int kPolMapK = 512;        // padded map columns
constexpr int kPolFeatK = 128;       // padded feature columns
constexpr int kPolHeads = 16;        // policy logits + value, padded (UMMA N of the heads product)
constexpr int kPolKChunk = 64;       // K elements per pipeline step: one 128-byte swizzle row of bf16
constexpr int kPolStages = 3;
constexpr int kPolABytes = kPolRows * kPolKChunk * 2;      // 16 KB: an activation chunk [128 x 64]
constexpr int kPolBBytes = kPolHidden * kPolKChunk * 2;    // 32 KB: a weight chunk [256 x 64]
constexpr int kPolHBytes = kPolRows * 512 * 2;             // 128 KB: H1 [128 x 512] as 8 chunks (H2 reuses the first 4)
constexpr int kPolSmemB = 0;
constexpr int kPolSmemH = kPolStages * kPolBBytes;         // the A stages alias the start of H (X is dead before H1 is written)
constexpr int kPolSmemBar = kPolSmemH + kPolHBytes;
constexpr int kPolSmemBytes = kPolSmemBar + 256 + 1024;    // + barriers + slack for the 1024-byte alignment
constexpr int kPolEpiWarps = 8;                            // two per TMEM lane quarter, each taking half of a layer's columns
constexpr int kPolThreads = 64 + 32 * kPolEpiWarps;
// pipeline steps: 8 map chunks, 2 feature chunks, 8 shared-layer chunks, 4 heads chunks
constexpr int kStepsMap = kPolMapK / kPolKChunk, kStepsFeat = kPolFeatK / kPolKChunk;
constexpr int kStepsShared = 512 / kPolKChunk, kStepsHeads = kPolHidden / kPolKChunk;
// Operand hand-offs from the epilogue to the MMA issuer per tile: H1's map half, H1's feature half, H2.  (Finer hand-offs do
// not help: the shared layer accumulates into the map encoder's tensor-memory columns and cannot start before the epilogue
// has read all 256 of them — tried, and it read stale accumulators.)
constexpr int kHandoffs = 3;

struct PolicyParams {
    const float* b_enc;     // [512]
    const float* b_shared;  // [256]
    const float* b_heads;   // [16]
    long long n_rows;
    int n_actions;
    unsigned long long seed;
    const long long* counter;
    long long offset;
    long long* actions;
    float* log_probs;
    float* values;
    float* heads_out;  // optional [n_rows][16] float32 (tests)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}

// K-major operand tile in shared memory written with the 128-byte swizzle (rows of 64 bf16 = 128 bytes, groups of 8
// rows 1024 bytes apart): UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp: SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address, 16-byte units
    d |= (uint64_t)0 << 16;                        // leading byte offset: unused for a swizzled K-major tile one atom wide
    d |= (uint64_t)(1024u >> 4) << 32;             // stride byte offset: 8 rows x 128 bytes
    d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                        // layout type: SWIZZLE_128B
    return d;
}

// kind::f16 instruction descriptor: bf16 x bf16 -> f32, both operands K-major (InstrDescriptor in the same header)
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {  // arrives on the barrier when every MMA issued so far is done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// issue only: the registers are valid after tmem_ld_wait()
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, "
        "%28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 32 accumulator columns of one row -> + bias, ReLU, bf16 -> four 16-byte units of the row in the swizzled K-major
// operand layout (unit c of row r sits at unit position c ^ (r & 7) of the row's 128 bytes)
// (SMEM_BIAS: the bias vector is a shared-memory copy — the pair kernel's cluster-scope arrivals invalidate L1 all the time)
template <bool SMEM_BIAS = false>
__device__ __forceinline__ void hidden_store32(unsigned char* h, int row, int n0, const uint32_t (&acc)[32], const float* __restrict__ bias) {
    unsigned char* chunk = h + (n0 >> 6) * kPolABytes + row * 128;
    const int unit0 = (n0 & 63) >> 3;
    const float4* b4 = reinterpret_cast<const float4*>(bias + n0);  // every thread reads the same 32 floats: eight broadcast loads
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const float4 lo = SMEM_BIAS ? b4[2 * u] : __ldg(b4 + 2 * u), hi = SMEM_BIAS ? b4[2 * u + 1] : __ldg(b4 + 2 * u + 1);
        const float bv[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = 8 * u + 2 * k;
            // bias in float32, round the pair to bfloat16, ReLU on the packed pair: rounding is monotonic and keeps the sign,
            // so max(round(x), 0) == round(max(x, 0)) — three instructions per pair instead of five
            const __nv_bfloat162 pair = __floats2bfloat162_rn(__uint_as_float(acc[j]) + bv[2 * k], __uint_as_float(acc[j + 1]) + bv[2 * k + 1]);
            const __nv_bfloat162 relu = __hmax2(pair, __float2bfloat162_rn(0.0f));
            w[k] = *reinterpret_cast<const uint32_t*>(&relu);
        }
        *reinterpret_cast<uint4*>(chunk + (((unit0 + u) ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// Hidden units [n_begin, n_end) of this thread's row: accumulator columns tcol .. (tcol = the tensor-memory address of unit
// n_begin) -> hidden activations in shared memory, 32 at a time, the tensor-memory load of the next 32 in flight while the
// current 32 are converted
template <bool SMEM_BIAS = false>
__device__ __forceinline__ void hidden_epilogue(unsigned char* h, int row, uint32_t tcol, int n_begin, int n_end, const float* __restrict__ bias) {
    uint32_t a0[32], a1[32];
    tmem_ld32_issue(tcol, a0);
    tmem_ld_wait();
#pragma unroll 1
    for (int n0 = n_begin; n0 < n_end; n0 += 64) {
        tmem_ld32_issue(tcol + (n0 - n_begin) + 32, a1);  // (n_end - n_begin is a multiple of 64)
        hidden_store32<SMEM_BIAS>(h, row, n0, a0, bias);
        tmem_ld_wait();
        if (n0 + 64 < n_end) tmem_ld32_issue(tcol + (n0 - n_begin) + 64, a0);
        hidden_store32<SMEM_BIAS>(h, row, n0 + 32, a1, bias);
        tmem_ld_wait();
    }
}

// The heads accumulator of one row -> logits and value (+ bias), float32 log_softmax and the Gumbel-max sample of
// sample_actions_kernel (same Philox draws), written to the rollout rows.
__device__ __forceinline__ void policy_sample_row(const PolicyParams& p, long long r, const uint32_t (&hv)[16]) {
    float x[TL_MAX_ACTIONS];
    float top = -INFINITY;
#pragma unroll
    for (int j = 0; j < kPolHeads; ++j) {
        const float v = __uint_as_float(hv[j]) + __ldg(p.b_heads + j);
        if (p.heads_out) p.heads_out[r * kPolHeads + j] = v;
        x[j] = j < p.n_actions ? v : -INFINITY;
        top = fmaxf(top, x[j]);
        if (j == p.n_actions && p.values) p.values[r] = v;
    }
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < kPolHeads; ++j) sum += j < p.n_actions ? expf(x[j] - top) : 0.0f;
    const float lse = top + logf(sum);
    const unsigned long long base = 4ull * (unsigned long long)((p.counter ? *p.counter : 0ll) + p.offset);
    int best = 0;
    float best_score = -INFINITY, best_logp = 0.0f;
#pragma unroll
    for (int d = 0; d < kPolHeads / 4; ++d) {
        if (4 * d >= p.n_actions) break;
        const unsigned long long c = base + d;
        const uint4 bits = philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)r, (uint32_t)((unsigned long long)r >> 32),
                                         (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
        const uint32_t w[4] = {bits.x, bits.y, bits.z, bits.w};
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const int j = 4 * d + l;
            if (j >= p.n_actions) break;
            const float u = ((float)(w[l] >> 9) + 0.5f) * 1.1920928955078125e-07f;
            const float logp = x[j] - lse;
            const float score = logp - logf(-logf(u));
            if (score > best_score) best_score = score, best = j, best_logp = logp;
        }
    }
    p.actions[r] = best;
    p.log_probs[r] = best_logp;
}

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one box into the same shared-memory offset of both CTAs of a pair; each CTA's barrier (same offset) counts the bytes
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    const uint16_t mask = 3;
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar) {  // arrives on the barrier at this offset in BOTH CTAs of the pair
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}

// MC (policy_mlp_mc_kernel): the CTAs of a cluster of two run the same schedule on their own tiles and SHARE the weight stream:
// each loads half of every weight chunk and multicasts it into both CTAs' rings, so a weight byte leaves L2 once per pair
// (452 instead of 744 KB per tile and SM — the loads are what the single-CTA form waits for).  A stage is refilled when the
// MMAs of BOTH CTAs have read it (every commit on `empty` goes to both CTAs); everything else stays per CTA.
template <bool MC>
__device__ __forceinline__ void policy_mlp_body(const CUtensorMap& map_x, const CUtensorMap& map_wm, const CUtensorMap& map_wf,
                                                const CUtensorMap& map_ws, const CUtensorMap& map_wh, const PolicyParams& p) {
    extern __shared__ unsigned char pol_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(pol_smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* s_b = smem + kPolSmemB;
    unsigned char* s_h = smem + kPolSmemH;  // also the A stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kPolSmemBar);
    uint64_t* full = bars;                  // [stages] TMA landed
    uint64_t* empty = bars + kPolStages;    // [stages] MMAs that read the stage are done
    uint64_t* mma_done = bars + 2 * kPolStages;      // accumulator of a layer is complete (three times per tile)
    // [kHandoffs] the epilogue has written a part of the next A operand.  One barrier per hand-off of a tile (each completes
    // once per tile): with a single barrier a warp that ran one hand-off ahead would complete the phase of a slower warp.
    uint64_t* h_ready = bars + 2 * kPolStages + 1;
    uint64_t* d3_read = h_ready + kHandoffs;         // the epilogue has read the heads accumulator (once per tile)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d3_read + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (int)((p.n_rows + kPolRows - 1) / kPolRows);
    // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...  (MC: both CTAs of a pair walk the same number of tiles — the
    // shared weight ring keeps them in step — so the count is the pair's; a tile past the end loads zeros and writes nothing)
    const uint32_t mc_rank = MC ? cluster_rank() : 0u;
    const int n_local = MC ? ((int)blockIdx.x / 2 < (n_tiles + 1) / 2 ? ((n_tiles + 1) / 2 - 1 - (int)blockIdx.x / 2) / ((int)gridDim.x / 2) + 1 : 0)
                           : ((int)blockIdx.x < n_tiles ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kPolStages; ++s) mbar_init(full + s, 1), mbar_init(empty + s, MC ? 2 : 1);
        mbar_init(mma_done, 1);
        for (int k = 0; k < kHandoffs; ++k) mbar_init(h_ready + k, 32 * kPolEpiWarps);
        mbar_init(d3_read, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wm) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wf) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ws) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wh) : "memory");
    }
    if (warp == 1) {  // the whole tensor memory: the two encoders' accumulators need all 512 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (MC) cluster_sync_all();  // the peer multicasts into this CTA's ring and arrives on its barriers: initialised first
    else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    // The schedule both the TMA producer and the MMA issuer walk (one ring step each, in this order), software-pipelined over
    // the tiles of the CTA so that the tensor pipe works on tile i + 1 while the epilogue warps convert and sample tile i:
    //   ENC_A(0) ENC_B(0)   then per tile i:   SHARED(i)   ENC_A(i+1)   HEADS(i)   ENC_B(i+1)
    // ENC_A = map encoder (8 K chunks), ENC_B = feature encoder (2), SHARED = shared layer (8), HEADS = policy + value heads (4).
    // Accumulator columns ping-pong with the tile's parity: with P = 256 * (i & 1) and S = 256 - P,
    //   ENC_A(i) -> [P, P+256)   ENC_B(i) -> [S, S+256)   SHARED(i) -> [P, P+256) (ENC_A's columns, already converted)
    //   HEADS(i) -> [P, P+16) (SHARED's columns, already converted);   ENC_A(i+1) takes S as soon as ENC_B(i) is converted.
    // The A stages of the encoders alias chunks 4-6 of the hidden buffer: free once SHARED(i) has read H1.
    enum Seg { ENC_A, ENC_B, SHARED, HEADS };
    auto a_stage = [&](int s) { return s_h + (4 + s) * kPolABytes; };

    if (warp == 0) {
        // ================= TMA producer (one lane) =================
        if (lane == 0 && n_local > 0) {
            int g = 0;  // ring step
            auto segment = [&](Seg seg, int it) {
                const int row0 = ((int)blockIdx.x + it * (int)gridDim.x) * kPolRows;
                const int n_steps = seg == ENC_A ? kStepsMap : seg == ENC_B ? kStepsFeat : seg == SHARED ? kStepsShared : kStepsHeads;
                for (int j = 0; j < n_steps; ++j, ++g) {
                    const int s = g % kPolStages;
                    if (g >= kPolStages) mbar_wait(empty + s, ((g / kPolStages) - 1) & 1);
                    unsigned char* b_dst = s_b + s * kPolBBytes;
                    if (MC) {
                        // this CTA's half of the weight chunk (rows 128 r .. of 256; 8 r .. of the 16 head rows) into both rings;
                        // the full barrier of each CTA counts the whole chunk: its own half and the peer's
                        const bool heads = seg == HEADS, enc = seg == ENC_A || seg == ENC_B;
                        const int half_rows = heads ? kPolHeads / 2 : kPolHidden / 2;
                        mbar_expect_tx(full + s, (enc ? kPolABytes : 0) + 2 * half_rows * kPolKChunk * 2);
                        tma_load_2d_mc(b_dst + mc_rank * half_rows * kPolKChunk * 2,
                                       seg == ENC_A ? &map_wm : seg == ENC_B ? &map_wf : seg == SHARED ? &map_ws : &map_wh, full + s,
                                       j * kPolKChunk, (int)mc_rank * half_rows);
                        if (enc) {
                            if (seg == ENC_A && j == 0 && it > 0) mbar_wait(mma_done, (3 * (it - 1) + 1) & 1);
                            tma_load_2d(a_stage(s), &map_x, full + s, (seg == ENC_A ? 0 : kPolMapK) + j * kPolKChunk, row0);
                        }
                    } else if (seg == ENC_A || seg == ENC_B) {
                        mbar_expect_tx(full + s, kPolABytes + kPolBBytes);
                        tma_load_2d(b_dst, seg == ENC_A ? &map_wm : &map_wf, full + s, j * kPolKChunk, 0);
                        // the A stages alias the hidden buffer: the previous tile's shared layer must be done reading H1
                        if (seg == ENC_A && j == 0 && it > 0) mbar_wait(mma_done, (3 * (it - 1) + 1) & 1);
                        tma_load_2d(a_stage(s), &map_x, full + s, (seg == ENC_A ? 0 : kPolMapK) + j * kPolKChunk, row0);
                    } else if (seg == SHARED) {
                        mbar_expect_tx(full + s, kPolBBytes);
                        tma_load_2d(b_dst, &map_ws, full + s, j * kPolKChunk, 0);
                    } else {
                        mbar_expect_tx(full + s, kPolHeads * kPolKChunk * 2);
                        tma_load_2d(b_dst, &map_wh, full + s, j * kPolKChunk, 0);
                    }
                }
            };
            segment(ENC_A, 0);
            segment(ENC_B, 0);
            for (int it = 0; it < n_local; ++it) {
                segment(SHARED, it);
                if (it + 1 < n_local) segment(ENC_A, it + 1);
                segment(HEADS, it);
                if (it + 1 < n_local) segment(ENC_B, it + 1);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (one lane) =================
        if (lane == 0 && n_local > 0) {
            const uint32_t idesc_n256 = umma_idesc_bf16(kPolRows, kPolHidden), idesc_n16 = umma_idesc_bf16(kPolRows, kPolHeads);
            int g = 0;
            POL_TIMING(long long m_full = 0, m_h = 0, m_d3 = 0, m_t0 = clock64(), m_c;)
            auto segment = [&](Seg seg, int it) {
                const uint32_t P = 256u * (uint32_t)(it & 1), S = 256u - P;
                const int n_steps = seg == ENC_A ? kStepsMap : seg == ENC_B ? kStepsFeat : seg == SHARED ? kStepsShared : kStepsHeads;
                const uint32_t d_col = seg == ENC_B ? S : P;
                const uint32_t idesc = seg == HEADS ? idesc_n16 : idesc_n256;
                if (seg == ENC_B && it > 0) {  // these columns hold the previous tile's heads accumulator until it has been read
                    POL_TIMING(m_c = clock64();)
                    mbar_wait(d3_read, (it - 1) & 1);
                    POL_TIMING(m_d3 += clock64() - m_c;)
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                for (int j = 0; j < n_steps; ++j, ++g) {
                    const int s = g % kPolStages;
                    // operands written by the epilogue: H1's map half before SHARED step 0, its feature half before step 4 (the
                    // epilogue converts it under the first four steps), H2 before HEADS
                    if ((seg == SHARED && (j == 0 || j == kStepsShared / 2)) || (seg == HEADS && j == 0)) {
                        const int k = seg == HEADS ? 2 : (j == 0 ? 0 : 1);
                        POL_TIMING(m_c = clock64();)
                        mbar_wait(h_ready + k, it & 1);
                        POL_TIMING(m_h += clock64() - m_c;)
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    POL_TIMING(m_c = clock64();)
                    mbar_wait(full + s, (g / kPolStages) & 1);
                    POL_TIMING(m_full += clock64() - m_c;)
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = (seg == ENC_A || seg == ENC_B) ? smem_u32(a_stage(s)) : smem_u32(s_h + j * kPolABytes);
                    const uint32_t b_addr = smem_u32(s_b + s * kPolBBytes);
#pragma unroll
                    for (int k = 0; k < kPolKChunk / 16; ++k)  // UMMA K = 16 bf16 = 32 bytes along the swizzled row
                        umma_bf16(tmem + d_col, umma_desc_k128(a_addr + 32 * k), umma_desc_k128(b_addr + 32 * k), idesc, !(j == 0 && k == 0));
                    if (MC) umma_commit_mc(empty + s);  // the stage is free when both CTAs' MMAs have read it
                    else umma_commit(empty + s);
                }
                if (seg != ENC_A) umma_commit(mma_done);  // completions per tile, in this order: encoders, shared layer, heads
            };
            segment(ENC_A, 0);
            segment(ENC_B, 0);
            for (int it = 0; it < n_local; ++it) {
                segment(SHARED, it);
                if (it + 1 < n_local) segment(ENC_A, it + 1);
                segment(HEADS, it);
                if (it + 1 < n_local) segment(ENC_B, it + 1);
            }
            POL_TIMING(if (blockIdx.x == 0) printf("TL_POLICY_TIMING MMA issuer cycles per tile: total %lld, waiting for TMA %lld, for the epilogue's operands %lld, for the heads read %lld\n",
                                                   (clock64() - m_t0) / n_local, m_full / n_local, m_h / n_local, m_d3 / n_local);)
        }
    } else {
        // ================= epilogue: warps 2-9, one thread per row and column half =================
        const int quarter = warp & 3;          // the TMEM lanes this warp may read
        const int half = (warp - 2) >> 2;      // which half of a layer's columns
        const int row = 32 * quarter + lane;
        const uint32_t trow = tmem + ((uint32_t)(32 * quarter) << 16);
        uint32_t hv[16];           // the heads accumulator of the previous tile's row, sampled under the next tile's shared layer
        long long pending_r = -1;  // (-1: nothing pending)
        POL_TIMING(long long tw1 = 0, te1 = 0, tw2 = 0, te2 = 0, tw3 = 0, te3 = 0, tc = clock64(), tn;)
        for (int it = 0; it < n_local; ++it) {
            const uint32_t P = 256u * (uint32_t)(it & 1), S = 256u - P;
            const long long r = (long long)((int)blockIdx.x + it * (int)gridDim.x) * kPolRows + row;
            // encoders: map half [P, P+256) -> H1 chunks 0-3 first, then the feature half [S, S+256) -> chunks 4-7 (the shared
            // layer's first MMAs run under the second conversion)
            mbar_wait(mma_done, (3 * it) & 1);
            POL_TIMING(tn = clock64(); tw1 += tn - tc; tc = tn;)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int q = 0; q < 2; ++q) {  // the map half, then the feature half: 128 columns per column half of the warps
                const int n0 = 256 * q + 128 * half;
                hidden_epilogue(s_h, row, trow + (q == 0 ? P : S) + 128 * half, n0, n0 + 128, p.b_enc);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the MMA's async proxy
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(h_ready + q);
            }
            POL_TIMING(tn = clock64(); te1 += tn - tc; tc = tn;)
            if (pending_r >= 0) {  // the previous tile's sample: its arithmetic hides under this tile's shared layer
                if (pending_r < p.n_rows) policy_sample_row(p, pending_r, hv);
                pending_r = -1;
            }
            POL_TIMING(tn = clock64(); te3 += tn - tc; tc = tn;)
            // shared layer: [P, P+256) -> H2 (chunks 0-3 of the same buffer; the shared layer has finished reading H1)
            mbar_wait(mma_done, (3 * it + 1) & 1);
            POL_TIMING(tn = clock64(); tw2 += tn - tc; tc = tn;)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            hidden_epilogue(s_h, row, trow + P + 128 * half, 128 * half, 128 * half + 128, p.b_shared);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(h_ready + kHandoffs - 1);
            POL_TIMING(tn = clock64(); te2 += tn - tc; tc = tn;)
            // heads: [P, P+16) -> registers; the next tile's feature encoder may overwrite the columns at once
            // (every epilogue thread follows every phase of the barrier: a parity wait must never fall two phases behind)
            mbar_wait(mma_done, (3 * it + 2) & 1);
            POL_TIMING(tn = clock64(); tw3 += tn - tc; tc = tn;)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (half != 0) continue;  // one warp per quarter samples
            tmem_ld16(trow + P, hv);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(d3_read);
            pending_r = r;
        }
        if (pending_r >= 0 && pending_r < p.n_rows) policy_sample_row(p, pending_r, hv);
        POL_TIMING(if (blockIdx.x == 0 && warp == 2 && lane == 0 && n_local > 0)
                       printf("TL_POLICY_TIMING cycles per tile: wait-enc %lld epi-enc %lld wait-shared %lld epi-shared %lld wait-heads %lld sample %lld (%d tiles)\n",
                              tw1 / n_local, te1 / n_local, tw2 / n_local, te2 / n_local, tw3 / n_local, te3 / n_local, n_local);)
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (MC) cluster_sync_all();  // the peer's multicast loads and commits may still target this CTA
    else __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

__global__ void __launch_bounds__(kPolThreads, 1)
policy_mlp_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_wm, const __grid_constant__ CUtensorMap map_wf,
                  const __grid_constant__ CUtensorMap map_ws, const __grid_constant__ CUtensorMap map_wh, const PolicyParams p) {
    policy_mlp_body<false>(map_x, map_wm, map_wf, map_ws, map_wh, p);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPolThreads, 1)
policy_mlp_mc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_wm, const __grid_constant__ CUtensorMap map_wf,
                     const __grid_constant__ CUtensorMap map_ws, const __grid_constant__ CUtensorMap map_wh, const PolicyParams p) {
    policy_mlp_body<true>(map_x, map_wm, map_wf, map_ws, map_wh, p);
}

// ---------------------------------------------------------------------------------------------------------------------
// The same step for a PAIR of CTAs (a cluster of two, tcgen05 cta_group::2): one MMA covers 256 agents — each CTA stages its
// own 128 input rows and HALF of every weight chunk (128 of the 256 output rows), the tensor cores of both SMs read both
// halves — so the weight bytes that cross from L2 per agent are halved.  That is what bounds the single-CTA kernel: 744 KB
// per tile per SM at the ~43 bytes per cycle per SM the L2 slices deliver chip-wide is the 17 k cycles per tile it
// measures; a pair moves 452 KB per CTA and tile.  Same schedule, software-pipelined over the pair's tiles:
//   ENC_A(0) ENC_B(0)   then per tile i:   SHARED(i)   ENC_A(i+1)   HEADS(i)   ENC_B(i+1)
// Each CTA keeps its own accumulators (its 128 rows) in its own tensor memory and its own H1 / H2 in its own shared memory;
// the leader CTA (cluster rank 0) issues every MMA and commits to the barriers of both CTAs; both CTAs' TMA loads report to
// the leader's `full` barrier; the peer's epilogue threads arrive on the leader's barriers across the cluster.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kPairStages = 4;
constexpr int kPairBBytes = (kPolHidden / 2) * kPolKChunk * 2;  // 16 KB: this CTA's half of a weight chunk
constexpr int kPairSmemH = kPairStages * kPairBBytes;           // the A stages alias chunks 4-7 of H
constexpr int kPairSmemBar = kPairSmemH + kPolHBytes;
constexpr int kPairSmemBias = kPairSmemBar + 256;               // encoder (512) and shared-layer (256) biases, float32
constexpr int kPairEarlyA = 1;                                  // the map encoder's first input chunk gets a buffer of its own
constexpr int kPairSmemA = kPairSmemBias + (512 + kPolHidden) * 4 + 768;  // (1024-byte aligned)
constexpr int kPairSmemBytes = kPairSmemA + kPairEarlyA * kPolABytes + 1024;
static_assert(kPairSmemBytes <= 227 * 1024, "shared memory of one SM");
static_assert(kPairSmemA % 1024 == 0, "swizzled operand tiles are 1024-byte aligned");

#define PAIR_PROXY_FENCE() asm volatile("fence.proxy.async.shared::cta;" ::: "memory")
__device__ __forceinline__ uint32_t leader_addr(const void* local) {  // the same variable in the cluster's CTA 0
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(smem_u32(local)));
    return r;
}
// remote arrival with the default (CTA-scope) release: no cluster-scope membar (the data it publishes sits in shared memory)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    const uint32_t zero = 0;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate), "r"(zero)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {  // arrives on the barrier at this offset in BOTH CTAs
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPolThreads, 1)
policy_mlp_pair_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_wm,
                       const __grid_constant__ CUtensorMap map_wf, const __grid_constant__ CUtensorMap map_ws,
                       const __grid_constant__ CUtensorMap map_wh, const PolicyParams p) {
    extern __shared__ unsigned char pol_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(pol_smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* s_b = smem;
    unsigned char* s_h = smem + kPairSmemH;  // also the A stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kPairSmemBar);
    uint64_t* full = bars;                        // [stages] leader: both CTAs' loads of the stage have landed
    uint64_t* empty = bars + kPairStages;         // [stages] each CTA: the MMAs that read the stage are done
    uint64_t* mma_done = bars + 2 * kPairStages;  // each CTA: a layer's accumulator is complete (three times per tile)
    uint64_t* h_local = mma_done + 1;             // [kHandoffs] leader: its own epilogue has written a part of the next A operand
    uint64_t* d3_local = h_local + kHandoffs;     // leader: its own epilogue has read the heads accumulator
    uint64_t* peer_h = d3_local + 1;              // [kHandoffs] leader: the same from the peer's epilogue warps (remote arrivals)
    uint64_t* peer_d3 = peer_h + kHandoffs;       // leader: the peer's heads read
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(peer_d3 + 1);
    float* s_bias = reinterpret_cast<float*>(smem + kPairSmemBias);  // [0, 512) encoders, [512, 768) shared layer

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const bool leader = rank == 0;
    const int n_pair_tiles = (int)((p.n_rows + 2 * kPolRows - 1) / (2 * kPolRows));
    const int pair0 = blockIdx.x >> 1, pair_stride = gridDim.x >> 1;
    const int n_local = pair0 < n_pair_tiles ? (n_pair_tiles - 1 - pair0) / pair_stride + 1 : 0;  // tiles of this pair

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kPairStages; ++s) mbar_init(full + s, 1), mbar_init(empty + s, 1);
        mbar_init(mma_done, 1);
        for (int k = 0; k < kHandoffs; ++k) mbar_init(h_local + k, kPolEpiWarps), mbar_init(peer_h + k, kPolEpiWarps);  // one elected lane per epilogue warp
        mbar_init(d3_local, 4);
        mbar_init(peer_d3, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wm) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wf) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ws) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wh) : "memory");
    }
    for (int i = threadIdx.x; i < 512 + kPolHidden; i += kPolThreads) s_bias[i] = i < 512 ? __ldg(p.b_enc + i) : __ldg(p.b_shared + i - 512);
    if (warp == 1) {  // both CTAs, the same warp: the pair's tensor memory (all 512 columns of each SM)
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();  // barriers of both CTAs are initialised before anything arrives on them across the cluster
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    // the schedule and the accumulator columns of policy_mlp_kernel (P = 256 * (i & 1), S = 256 - P)
    enum Seg { ENC_A, ENC_B, SHARED, HEADS };
    // A stage of ring slot s: chunk 4 + s of the hidden buffer (free once the previous tile's shared layer has read H1) — except
    // the first chunks of the map encoder, which have buffers of their own so that their loads run under that shared layer
    // and the tensor pipe does not wait a load latency at every tile boundary
    unsigned char* s_a_early = smem + kPairSmemA;
    auto a_stage = [&](Seg seg, int j, int s) { return (seg == ENC_A && j < kPairEarlyA) ? s_a_early + j * kPolABytes : s_h + (4 + s) * kPolABytes; };

    if (warp == 0) {
        // ================= TMA producer (one lane, both CTAs) =================
        if (lane == 0 && n_local > 0) {
            int g = 0;  // ring step
            const int half_row = (int)rank * (kPolHidden / 2);
            auto segment = [&](Seg seg, int it) {
                const int row0 = (pair0 + it * pair_stride) * 2 * kPolRows + (int)rank * kPolRows;
                const int n_steps = seg == ENC_A ? kStepsMap : seg == ENC_B ? kStepsFeat : seg == SHARED ? kStepsShared : kStepsHeads;
                for (int j = 0; j < n_steps; ++j, ++g) {
                    const int s = g % kPairStages;
                    if (g >= kPairStages) mbar_wait(empty + s, ((g / kPairStages) - 1) & 1);
                    unsigned char* b_dst = s_b + s * kPairBBytes;
                    const uint32_t bar = leader_addr(full + s);
                    if (seg == ENC_A || seg == ENC_B) {
                        if (leader) mbar_expect_tx(full + s, 2 * (kPolABytes + kPairBBytes));
                        tma_load_2d_pair(b_dst, seg == ENC_A ? &map_wm : &map_wf, bar, j * kPolKChunk, half_row);
                        // the aliased A stages: the previous tile's shared layer must be done reading H1 (four stages keep this
                        // lane within one phase of that completion: its ring slot was freed by one of the shared layer's last steps)
                        if (seg == ENC_A && j == kPairEarlyA && it > 0) mbar_wait(mma_done, (3 * (it - 1) + 1) & 1);
                        tma_load_2d_pair(a_stage(seg, j, s), &map_x, bar, (seg == ENC_A ? 0 : kPolMapK) + j * kPolKChunk, row0);
                    } else if (seg == SHARED) {
                        if (leader) mbar_expect_tx(full + s, 2 * kPairBBytes);
                        tma_load_2d_pair(b_dst, &map_ws, bar, j * kPolKChunk, half_row);
                    } else {
                        if (leader) mbar_expect_tx(full + s, 2 * (kPolHeads / 2) * kPolKChunk * 2);
                        tma_load_2d_pair(b_dst, &map_wh, bar, j * kPolKChunk, (int)rank * (kPolHeads / 2));
                    }
                }
            };
            segment(ENC_A, 0);
            segment(ENC_B, 0);
            for (int it = 0; it < n_local; ++it) {
                segment(SHARED, it);
                if (it + 1 < n_local) segment(ENC_A, it + 1);
                segment(HEADS, it);
                if (it + 1 < n_local) segment(ENC_B, it + 1);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (one lane of the leader CTA) =================
        if (leader && lane == 0 && n_local > 0) {
            const uint32_t idesc_n256 = umma_idesc_bf16(2 * kPolRows, kPolHidden), idesc_n16 = umma_idesc_bf16(2 * kPolRows, kPolHeads);
            int g = 0;
            POL_TIMING(long long m_full = 0, m_h = 0, m_d3 = 0, m_t0 = clock64(), m_c;)
            auto segment = [&](Seg seg, int it) {
                const uint32_t P = 256u * (uint32_t)(it & 1), S = 256u - P;
                const int n_steps = seg == ENC_A ? kStepsMap : seg == ENC_B ? kStepsFeat : seg == SHARED ? kStepsShared : kStepsHeads;
                const uint32_t d_col = seg == ENC_B ? S : P;
                const uint32_t idesc = seg == HEADS ? idesc_n16 : idesc_n256;
                if (seg == ENC_B && it > 0) {  // these columns hold the previous tile's heads accumulator until both CTAs have read it
                    POL_TIMING(m_c = clock64();)
                    mbar_wait(d3_local, (it - 1) & 1);
                    mbar_wait(peer_d3, (it - 1) & 1);
                    POL_TIMING(m_d3 += clock64() - m_c;)
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                for (int j = 0; j < n_steps; ++j, ++g) {
                    const int s = g % kPairStages;
                    // operands written by the epilogue: H1's map half before SHARED step 0, its feature half before step 4 (the
                    // epilogue converts it under the first four steps), H2 before HEADS
                    if ((seg == SHARED && (j == 0 || j == kStepsShared / 2)) || (seg == HEADS && j == 0)) {
                        const int k = seg == HEADS ? 2 : (j == 0 ? 0 : 1);
                        POL_TIMING(m_c = clock64();)
                        mbar_wait(h_local + k, it & 1);
                        mbar_wait(peer_h + k, it & 1);
                        POL_TIMING(m_h += clock64() - m_c;)
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    POL_TIMING(m_c = clock64();)
                    mbar_wait(full + s, (g / kPairStages) & 1);
                    POL_TIMING(m_full += clock64() - m_c;)
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = (seg == ENC_A || seg == ENC_B) ? smem_u32(a_stage(seg, j, s)) : smem_u32(s_h + j * kPolABytes);
                    const uint32_t b_addr = smem_u32(s_b + s * kPairBBytes);
#pragma unroll
                    for (int k = 0; k < kPolKChunk / 16; ++k)
                        umma_bf16_pair(tmem + d_col, umma_desc_k128(a_addr + 32 * k), umma_desc_k128(b_addr + 32 * k), idesc, !(j == 0 && k == 0));
                    umma_commit_pair(empty + s);
                }
                if (seg != ENC_A) umma_commit_pair(mma_done);  // completions per tile, in this order: encoders, shared layer, heads
            };
            segment(ENC_A, 0);
            segment(ENC_B, 0);
            for (int it = 0; it < n_local; ++it) {
                segment(SHARED, it);
                if (it + 1 < n_local) segment(ENC_A, it + 1);
                segment(HEADS, it);
                if (it + 1 < n_local) segment(ENC_B, it + 1);
            }
            POL_TIMING(if (blockIdx.x == 0) printf("TL_POLICY_TIMING pair MMA issuer cycles per tile: total %lld, waiting for TMA %lld, for the epilogues' operands %lld, for the heads reads %lld\n",
                                                   (clock64() - m_t0) / n_local, m_full / n_local, m_h / n_local, m_d3 / n_local);)
        }
    } else {
        // ================= epilogue: warps 2-9 of both CTAs, one thread per row and column half =================
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = 32 * quarter + lane;
        const uint32_t trow = tmem + ((uint32_t)(32 * quarter) << 16);
        uint32_t hv[16];           // the heads accumulator of the previous tile's row, sampled under the next tile's shared layer
        long long pending_r = -1;  // (-1: nothing pending)
        // This warp's part of an operand is written: every lane fences its generic-proxy stores towards the async proxy (the
        // pair's MMA), the warp synchronises, ONE lane arrives on the leader's barrier — the peer's across the cluster, with the
        // default CTA-scope release (the 2-SM kernels of CUTLASS signal operand hand-offs the same way): a cluster-scope release
        // is a membar that also waits for the thread's outstanding global stores (the sampled actions) and cost the epilogue
        // warps ~3 k cycles per tile; the data published here sits in shared memory, written before the arrival is sent.
        auto arrive_warp = [&](uint64_t* bar) {
            PAIR_PROXY_FENCE();
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(bar);
                else mbar_arrive_remote(leader_addr(peer_h + (bar - h_local)));
            }
        };
        POL_TIMING(long long tw1 = 0, te1 = 0, tw2 = 0, te2 = 0, tw3 = 0, te3 = 0, tc = clock64(), tn;)
        for (int it = 0; it < n_local; ++it) {
            const uint32_t P = 256u * (uint32_t)(it & 1), S = 256u - P;
            const long long r = (long long)(pair0 + it * pair_stride) * 2 * kPolRows + (long long)rank * kPolRows + row;
            mbar_wait(mma_done, (3 * it) & 1);
            POL_TIMING(tn = clock64(); tw1 += tn - tc; tc = tn;)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int q = 0; q < 2; ++q) {  // the map half, then the feature half: 128 columns per column half of the warps
                const int n0 = 256 * q + 128 * half;
                hidden_epilogue<true>(s_h, row, trow + (q == 0 ? P : S) + 128 * half, n0, n0 + 128, s_bias);
                arrive_warp(h_local + q);
            }
            POL_TIMING(tn = clock64(); te1 += tn - tc; tc = tn;)
            if (pending_r >= 0) {  // the previous tile's sample: its arithmetic hides under this tile's shared layer
                if (pending_r < p.n_rows) policy_sample_row(p, pending_r, hv);
                pending_r = -1;
            }
            POL_TIMING(tn = clock64(); te3 += tn - tc; tc = tn;)
            mbar_wait(mma_done, (3 * it + 1) & 1);
            POL_TIMING(tn = clock64(); tw2 += tn - tc; tc = tn;)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            hidden_epilogue<true>(s_h, row, trow + P + 128 * half, 128 * half, 128 * half + 128, s_bias + 512);
            arrive_warp(h_local + kHandoffs - 1);
            POL_TIMING(tn = clock64(); te2 += tn - tc; tc = tn;)
            // (every epilogue thread follows every phase of the barrier: a parity wait must never fall two phases behind)
            mbar_wait(mma_done, (3 * it + 2) & 1);
            POL_TIMING(tn = clock64(); tw3 += tn - tc; tc = tn;)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (half != 0) continue;  // one warp per quarter samples
            tmem_ld16(trow + P, hv);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(d3_local);
                else mbar_arrive_remote(leader_addr(peer_d3));
            }
            pending_r = r;
        }
        if (pending_r >= 0 && pending_r < p.n_rows) policy_sample_row(p, pending_r, hv);
        POL_TIMING(if (blockIdx.x < 2 && warp == 2 && lane == 0 && n_local > 0)
                       printf("TL_POLICY_TIMING pair rank %u cycles per tile: wait-enc %lld epi-enc %lld wait-shared %lld epi-shared %lld wait-heads %lld sample %lld (%d tiles)\n",
                              rank, tw1 / n_local, te1 / n_local, tw2 / n_local, te2 / n_local, tw3 / n_local, te3 / n_local, n_local);)
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();  // neither CTA leaves (or frees tensor memory) while the pair's MMAs or remote arrivals may touch it
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = []() -> PFN_cuTensorMapEncodeTiled_v12000 {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }();
    return fn;
}

// bf16 row-major matrix [rows][cols] (row pitch in elements) as a 2-D tensor map with boxes of 64 columns x box_rows rows
bool make_bf16_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t box_rows) {
    auto encode = tensor_map_encoder();
    if (!encode) return set_error("cuTensorMapEncodeTiled is not available"), false;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {pitch * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kPolKChunk, box_rows};
    const cuuint32_t elem[2] = {1, 1};
    const CUresult rc = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, elem,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled failed"), false;
    return true;
}

}  // namespace

extern "C" {

int tl_policy_step(const void* inputs, int64_t n_rows, int64_t input_row, const TlPolicyWeights* w, uint64_t seed, const int64_t* counter,
                   int64_t offset, int64_t* actions, float* log_probs, float* values, float* heads_out, tl_stream_t stream) {
    if (!inputs || !w || !actions || !log_probs) return set_error("tl_policy_step: NULL pointer"), TL_ERR_ARG;
    if (!w->w_map || !w->w_feat || !w->w_shared || !w->w_heads || !w->b_enc || !w->b_shared || !w->b_heads)
        return set_error("tl_policy_step: NULL weight pointer"), TL_ERR_ARG;
    if (w->hidden != kPolHidden || w->map_in != kPolMapK || w->feat_in != kPolFeatK)
        return set_error("tl_policy_step: built for hidden 256, 512 map columns, 128 feature columns"), TL_ERR_UNSUPPORTED;
    if (w->n_actions < 1 || w->n_actions + (values ? 1 : 0) > kPolHeads) return set_error("tl_policy_step: 1 <= n_actions <= 15"), TL_ERR_ARG;
    if (input_row < kPolMapK + kPolFeatK || (input_row & 7) || ((uintptr_t)inputs & 15))
        return set_error("tl_policy_step: input rows are >= 640 bf16, a multiple of 8, 16-byte aligned"), TL_ERR_ARG;
    if (n_rows < 0 || offset < 0) return set_error("tl_policy_step: negative size"), TL_ERR_ARG;
    if (n_rows == 0) return TL_OK;
    int device = 0;
    if (!cuda_ok(cudaGetDevice(&device), "cudaGetDevice")) return TL_ERR_CUDA;
    const DeviceInfo* di = device_info(device);
    if (!di) return set_error("cudaDeviceGetAttribute failed"), TL_ERR_CUDA;
    const bool pair = g_opt.policy_pair.load() != 0;  // CTA pairs (cta_group::2): half the weight bytes per agent
    const bool mc = !pair && g_opt.policy_multicast.load() != 0;  // clusters of two sharing the weight stream by TMA multicast
    const uint32_t w_rows = (pair || mc) ? kPolHidden / 2 : kPolHidden, h_rows = (pair || mc) ? kPolHeads / 2 : kPolHeads;
    CUtensorMap mx, mwm, mwf, mws, mwh;
    if (!make_bf16_map(&mx, inputs, (uint64_t)n_rows, kPolMapK + kPolFeatK, (uint64_t)input_row, kPolRows) ||
        !make_bf16_map(&mwm, w->w_map, kPolHidden, kPolMapK, kPolMapK, w_rows) ||
        !make_bf16_map(&mwf, w->w_feat, kPolHidden, kPolFeatK, kPolFeatK, w_rows) ||
        !make_bf16_map(&mws, w->w_shared, kPolHidden, 512, 512, w_rows) ||
        !make_bf16_map(&mwh, w->w_heads, kPolHeads, kPolHidden, kPolHidden, h_rows))
        return TL_ERR_CUDA;
    static std::once_flag once;
    static cudaError_t attr_rc = cudaSuccess;
    std::call_once(once, [] {
        attr_rc = cudaFuncSetAttribute(policy_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPolSmemBytes);
        if (attr_rc == cudaSuccess) attr_rc = cudaFuncSetAttribute(policy_mlp_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes);
        if (attr_rc == cudaSuccess) attr_rc = cudaFuncSetAttribute(policy_mlp_mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPolSmemBytes);
    });
    if (!cuda_ok(attr_rc, "cudaFuncSetAttribute(policy_mlp_kernel)")) return TL_ERR_CUDA;
    PolicyParams p;
    p.b_enc = w->b_enc, p.b_shared = w->b_shared, p.b_heads = w->b_heads;
    p.n_rows = n_rows, p.n_actions = w->n_actions, p.seed = seed, p.counter = (const long long*)counter, p.offset = offset;
    p.actions = (long long*)actions, p.log_probs = log_probs, p.values = values, p.heads_out = heads_out;
    // persistent CTAs, one per SM (a CTA takes the whole shared memory and tensor memory of its SM)
    if (pair) {
        const long long n_pairs = (n_rows + 2 * kPolRows - 1) / (2 * kPolRows);
        const long long max_pairs = di->sms / 2;
        const unsigned grid = 2u * (unsigned)(n_pairs < max_pairs ? n_pairs : max_pairs);
        policy_mlp_pair_kernel<<<grid, kPolThreads, kPairSmemBytes, (cudaStream_t)stream>>>(mx, mwm, mwf, mws, mwh, p);
    } else if (mc) {
        const long long n_pairs = ((n_rows + kPolRows - 1) / kPolRows + 1) / 2;
        const long long max_pairs = di->sms / 2;
        const unsigned grid = 2u * (unsigned)(n_pairs < max_pairs ? n_pairs : max_pairs);
        policy_mlp_mc_kernel<<<grid, kPolThreads, kPolSmemBytes, (cudaStream_t)stream>>>(mx, mwm, mwf, mws, mwh, p);
    } else {
        const long long n_tiles = (n_rows + kPolRows - 1) / kPolRows;
        const unsigned grid = (unsigned)(n_tiles < di->sms ? n_tiles : di->sms);
        policy_mlp_kernel<<<grid, kPolThreads, kPolSmemBytes, (cudaStream_t)stream>>>(mx, mwm, mwf, mws, mwh, p);
    }
    if (!cuda_ok(cudaGetLastError(), "policy_mlp launch")) return TL_ERR_CUDA;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return TL_OK;
}

}  // extern "C"
