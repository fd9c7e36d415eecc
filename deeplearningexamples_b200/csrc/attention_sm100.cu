// Fused multi-head self-attention for sm_100a (head dim 64, S in {128,256,384,512}).
//
// Forward : one CTA per (batch, head, 128-query tile).  TMA stages Q, K, V of the head from the packed
//           [T, 3H] QKV matrix into 128B-swizzled smem; tcgen05.mma computes S = Q K^T (fp32 in TMEM)
//           128 keys at a time; 8 softmax warps run an online softmax straight out of TMEM, apply
//           dropout (Philox, regenerated in backward) and hand P (bf16, swizzled smem) back to the
//           tensor core for O += P V (V consumed MN-major, i.e. where the QKV GEMM wrote it).
//           The [B,A,S,S] score tensor the reference materialises 3x per layer never exists in HBM.
// Backward: one CTA per (batch, head); loops over (kv tile, q tile) pairs, recomputes P from the saved
//           log-sum-exp, and accumulates dV, dK (per kv tile) and dQ (all q tiles) in TMEM -- no atomics,
//           deterministic.  The P and dS tiles are written once to smem and read by the tensor core both
//           K-major (dQ = dS K) and MN-major (dV = P^T dO, dK = dS^T Q): same bytes, two descriptors.
//
// replaces BertSelfAttention.forward, PyTorch/LanguageModeling/BERT/modeling.py:349-376
// (transpose_for_scores, bmm, /sqrt(d), +mask, softmax, dropout, bmm, transpose+contiguous) and autograd.
#include "common.cuh"
#include "../../include/dle_b200.h"

namespace dle {

constexpr int HD = 64;                 // head dim
constexpr int TQ = 128;                // query tile / key chunk
constexpr int TILE_BYTES = TQ * HD * 2;    // 16 KB : one [128 x 64] bf16 tile (128 B rows)
constexpr int PT_BYTES = TQ * TQ * 2;      // 32 KB : one [128 x 128] bf16 tile = two 16 KB sub-tiles
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// byte offset of 8 consecutive columns (col % 8 == 0) of row r inside a [128 x 128] bf16 tile stored as two
// K-major SWIZZLE_128B sub-tiles (cols 0-63 | 64-127), 128 B per row, 16-byte chunks XOR-swizzled by row%8
__device__ __forceinline__ uint32_t pt_offset(int r, int col) {
    return (uint32_t)((col >> 6) * TILE_BYTES + r * 128 + ((((col & 63) >> 3) ^ (r & 7)) << 4));
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// =================================================================================================
// forward  (v3)
//   CTA   : one (batch, head, 128-query tile); sized for TWO co-resident CTAs per SM so one CTA's softmax overlaps the other's MMAs
//   smem  : Q 16 KB + K ring 2x16 KB + V ring 2x16 KB + P 32 KB + 512 B max exchange + barriers = 112.6 KB  -> 2 CTAs in 228 KB
//   TMEM  : S 128 cols + O 64 cols -> 256-column allocation                                             -> 2 CTAs in 512 cols
//   warps : 0..7 = softmax, 8 = TMA producer (+TMEM alloc), 9 = MMA issuer.  TWO threads per query row: the two warps of a TMEM lane
//           quarter each own 64 of the 128 key columns of a chunk (v2 had one thread per row = 2 softmax warps per scheduler with both
//           CTAs resident and was dependency-bound: ncu 44 % issue-active).  The halves agree on the chunk's row maximum through a
//           512 B shared exchange (bf16, rounded UP so that exp2(s - m) <= 1 still holds) and a 64-thread named barrier.
//   lazy rescale : the running maximum only moves when a chunk exceeds it by more than 2^8 (P stays <= 256, exact in the final
//           normalisation because l is accumulated against the same reference), so the O read-modify-write through TMEM is rare.
//   math  : packed FFMA2/FADD2 for scale-subtract and the row sums; dropout keeps arrive as bf16x2 AND-masks (common.cuh).
// =================================================================================================
constexpr int FWD_SOFTMAX_WARPS = 8;
constexpr int FWD_THREADS = (2 + FWD_SOFTMAX_WARPS) * 32;      // 320: warps 0..7 softmax, 8 TMA producer (+TMEM alloc), 9 MMA issuer (highest ids:
constexpr int FWD_WARP_TMA = FWD_SOFTMAX_WARPS, FWD_WARP_MMA = FWD_SOFTMAX_WARPS + 1;      // the arbiter favours them over the softmax warps)
constexpr int FWD_TMEM_COLS = 256;
constexpr int FWD_XCHG_BYTES = 2 * TQ * 2;                     // [half][row] bf16
constexpr int FWD_BAR_BYTES = 160;
constexpr int FWD_SMEM_BYTES = TILE_BYTES /*Q*/ + 4 * TILE_BYTES /*K,V rings*/ + PT_BYTES /*P*/ + FWD_XCHG_BYTES + FWD_BAR_BYTES;
constexpr float FWD_RESCALE_THRESH = 8.0f;                     // log2 domain
constexpr int HALF_BYTES = 64 * HD * 2;                        // 8 KB: 64 key rows of a K / V tile

struct AttnFwdParams {
    const float* mask;     // [B,S] additive or null
    bf16* ctx;             // [T, H]
    float* lse;            // [B,A,S]
    int B, S, A, H;
    int tok_stride_s, tok_stride_b;   // token row of (b, s) = b*tok_stride_b + s*tok_stride_s
    float scale_log2;      // (1/sqrt(d)) * log2(e)
    uint32_t drop_k2;      // attn_dropout_k2(thresh15); dropout off when drop_on == 0
    uint32_t drop_on; float drop_scale; uint32_t drop_stream; unsigned long long seed; const unsigned long long* seed_dev;
};

// S is produced in two 64-key halves (one N = 64 MMA group per softmax half) and each half's TMEM columns are handed back as soon as the
// owning warps hold them in registers for the exponentials (s_free): S_{j+1} is then computed by the tensor core WHILE chunk j's
// exponentials run, instead of after them (v3a: 20 % of all stall samples were the softmax warps waiting for S).
// staging tile of a drain: 128 rows x 128 B, 16-byte units XOR-swizzled by the row so that both the row-per-lane writes and the
// 8-lanes-per-row reads are conflict-free
__device__ __forceinline__ uint32_t drain_off(int row, int unit) { return (uint32_t)(row * 128 + ((unit ^ (row & 7)) << 4)); }

__global__ void __launch_bounds__(FWD_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnFwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int S = p.S, n_all = S / TQ;
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + TILE_BYTES;               // [2]
    uint8_t* sV = sK + 2 * TILE_BYTES;           // [2]
    uint8_t* sP = sV + 2 * TILE_BYTES;
    uint16_t* sX = reinterpret_cast<uint16_t*>(sP + PT_BYTES);          // [2][128] bf16 chunk maxima
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + PT_BYTES + FWD_XCHG_BYTES);
    uint64_t* q_full = bars;             // 1
    uint64_t* k_full = bars + 1;         // [2]
    uint64_t* k_empty = bars + 3;        // [2]
    uint64_t* v_full = bars + 5;         // [2]
    uint64_t* v_empty = bars + 7;        // [2]
    uint64_t* s_full = bars + 9;         // [2]  S half hf complete in TMEM
    uint64_t* s_free = bars + 11;        // [2]  the 4 warps of half hf hold it in registers (one arrival per warp)
    uint64_t* p_full = bars + 13;        // P_j complete in smem (one arrival per softmax warp)
    uint64_t* pv_done = bars + 14;       // PV_j retired
    uint64_t* nk_ready = bars + 15;      // the number of key chunks to process is known (one arrival per softmax warp)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 16);
    int* s_nk = reinterpret_cast<int*>(bars + 16) + 1;     // number of key chunks that hold at least one attendable key

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) __trap();          // SWIZZLE_128B tiles need a 1024-byte aligned base
        tma_prefetch_desc(&tmap_qkv);
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
            mbar_init(&s_full[i], 1); mbar_init(&s_free[i], FWD_SOFTMAX_WARPS / 2);
        }
        mbar_init(p_full, FWD_SOFTMAX_WARPS); mbar_init(pv_done, 1); mbar_init(nk_ready, FWD_SOFTMAX_WARPS);
        fence_barrier_init();
        *s_nk = 0;
    }
    if (warp == FWD_WARP_TMA) { tmem_alloc(tmem_ptr, FWD_TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + TQ;
    // Variable-length batches: trailing key chunks whose additive mask is <= -1000 for EVERY key contribute exp2(..) == 0 exactly in
    // fp32 (the running maximum comes from an attendable key), so they are skipped outright -- no load, no MMA, no softmax -- and the
    // result is bit-identical to processing them (padding-skipping of FasterTransformer-style inference; also
    // applies to training on LDDL's binned, partly padded batches).  The softmax warps scan the mask row while chunk 0 (always
    // needed) is already being loaded and multiplied; every role picks the count up through nk_ready before it goes past chunk 0.
    const bool scan = p.mask != nullptr && n_all > 1;
    auto chunks_to_do = [&]() -> int {
        if (!scan) return n_all;
        mbar_wait(nk_ready, 0);
        const int nk = *reinterpret_cast<volatile int*>(s_nk);
        return nk > 0 ? nk : n_all;
    };

    // Producer and MMA warps: warp-uniform loops (all lanes wait), tcgen05 / TMA instructions under elect_one() -- see gemm_sm100.cu.
    if (warp == FWD_WARP_TMA) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            mbar_expect_tx(q_full, TILE_BYTES);
            tma_load_3d(sQ, &tmap_qkv, q_full, h * HD, qt * TQ, b);
            mbar_expect_tx(&k_full[0], TILE_BYTES);
            tma_load_3d(sK, &tmap_qkv, &k_full[0], p.H + h * HD, 0, b);
            mbar_expect_tx(&v_full[0], TILE_BYTES);
            tma_load_3d(sV, &tmap_qkv, &v_full[0], 2 * p.H + h * HD, 0, b);
        }
        __syncwarp();
        const int n_chunks = chunks_to_do();
        for (int j = 1; j < n_chunks; ++j) {
            const int st = j & 1; const uint32_t ph = (j >> 1) & 1;
            mbar_wait(&k_empty[st], ph ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&k_full[st], TILE_BYTES);
                tma_load_3d(sK + st * TILE_BYTES, &tmap_qkv, &k_full[st], p.H + h * HD, j * TQ, b);
            }
            __syncwarp();
            mbar_wait(&v_empty[st], ph ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&v_full[st], TILE_BYTES);
                tma_load_3d(sV + st * TILE_BYTES, &tmap_qkv, &v_full[st], 2 * p.H + h * HD, j * TQ, b);
            }
            __syncwarp();
        }
    } else if (warp == FWD_WARP_MMA) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc_s = make_idesc_bf16(TQ, 64, false, false);
        constexpr uint32_t idesc_pv = make_idesc_bf16(TQ, HD, false, true);
        const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);
        auto issue_s = [&](int j_) {                       // S_j = Q K_j^T, one N = 64 group per key half
            const int st = j_ & 1;
            mbar_wait(&k_full[st], (j_ >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                if (j_ >= 1) { mbar_wait(&s_free[hf], (j_ - 1) & 1); tc_fence_after(); }     // half hf of S_{j-1} is in registers
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16_ss(tmem_S + hf * 64, make_smem_desc_sw128(aQ + kk * 32, 0, 1024),
                                     make_smem_desc_sw128(aK + st * TILE_BYTES + hf * HALF_BYTES + kk * 32, 0, 1024), idesc_s, kk > 0 ? 1u : 0u);
                    umma_commit(&s_full[hf]);
                    if (hf == 1) umma_commit(&k_empty[st]);
                }
                __syncwarp();
            }
        };
        mbar_wait(q_full, 0);
        issue_s(0);
        const int n_chunks = chunks_to_do();
        for (int j = 0; j < n_chunks; ++j) {
            const int st = j & 1; const uint32_t ph = (j >> 1) & 1;
            if (j + 1 < n_chunks) issue_s(j + 1);           // runs under chunk j's exponentials (its halves are released early)
            mbar_wait(p_full, j & 1);                        // P_j written
            tc_fence_after();
            mbar_wait(&v_full[st], ph);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)                // O += P_j V_j
                    umma_bf16_ss(tmem_O, make_smem_desc_sw128(aP + (kk >> 2) * TILE_BYTES + (kk & 3) * 32, 0, 1024),
                                 make_smem_desc_sw128(aV + st * TILE_BYTES + kk * 2048, TILE_BYTES, 1024), idesc_pv,
                                 (j > 0 || kk > 0) ? 1u : 0u);
                umma_commit(&v_empty[st]);
                umma_commit(pv_done);
            }
            __syncwarp();
        }
    } else {
        // ===================== softmax warps: two threads per query row, 64 key columns each =====================
        const int q4 = warp & 3;                      // TMEM lane quarter this warp may access (hardware rule: warp id % 4)
        const int hf = warp >> 2;                     // which 64-column half of the chunk
        const int r = q4 * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
        const uint32_t tS = tmem_S + lane_addr + hf * 64, tO = tmem_O + lane_addr + hf * 32;
        const uint32_t sp = smem_u32(sP);
        const unsigned long long seed = effective_seed(p.seed, p.seed_dev);
        float m_run = -INFINITY, l0 = 0.f;
        const unsigned long long drop_row = ((unsigned long long)(b * p.A + h) * S + (qt * TQ + r)) * (unsigned long long)S + hf * 64;
        const float* mrow = p.mask ? p.mask + (long long)b * S + hf * 64 : nullptr;
        // one pass over the mask row: which 128-key chunks carry a non-zero additive mask in MY half (warp-uniform bit set; typical
        // batches: none or the tail) and which is the last chunk with an attendable key anywhere (all 8 warps cover the row together)
        uint32_t chunk_masked = 0;
        if (mrow) {
            int last = 0;
            for (int j = 0; j < n_all; ++j) {
                const float2 m2 = __ldg(reinterpret_cast<const float2*>(mrow + j * TQ) + lane);
                if (__any_sync(0xffffffffu, m2.x != 0.f || m2.y != 0.f)) chunk_masked |= 1u << j;
                if (__any_sync(0xffffffffu, m2.x > -1000.0f || m2.y > -1000.0f)) last = j + 1;
            }
            if (scan) {
                if (lane == 0) { if (last > 0) atomicMax(s_nk, last); __threadfence_block(); mbar_arrive(nk_ready); }
            }
        }
        const int n_chunks = chunks_to_do();
        for (int j = 0; j < n_chunks; ++j) {
            const bool masked = (chunk_masked >> j) & 1u;
            const float* mk = mrow + j * TQ;
            mbar_wait(&s_full[hf], j & 1);
            tc_fence_after();
            uint32_t v[32];
            // ---- pass 1: maximum of my 64 columns
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int pc = 0; pc < 2; ++pc) {
                tmem_ld32(tS + pc * 32, v);
                tmem_ld_wait();
                if (masked) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 m4 = __ldg(reinterpret_cast<const float4*>(mk + pc * 32) + i);
                        mx0 = fmaxf(mx0, fmaf(__uint_as_float(v[4 * i]), p.scale_log2, m4.x * LOG2E));
                        mx1 = fmaxf(mx1, fmaf(__uint_as_float(v[4 * i + 1]), p.scale_log2, m4.y * LOG2E));
                        mx2 = fmaxf(mx2, fmaf(__uint_as_float(v[4 * i + 2]), p.scale_log2, m4.z * LOG2E));
                        mx3 = fmaxf(mx3, fmaf(__uint_as_float(v[4 * i + 3]), p.scale_log2, m4.w * LOG2E));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        mx0 = fmaxf(mx0, __uint_as_float(v[i])); mx1 = fmaxf(mx1, __uint_as_float(v[i + 1]));
                        mx2 = fmaxf(mx2, __uint_as_float(v[i + 2])); mx3 = fmaxf(mx3, __uint_as_float(v[i + 3]));
                    }
                }
            }
            float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            if (!masked) mx *= p.scale_log2;                          // scale > 0: max commutes with the scaling
            // ---- agree on the chunk maximum with the thread that owns the other 64 columns of this row
            const bf16 mine = __float2bfloat16_ru(mx);
            sX[hf * TQ + r] = __bfloat16_as_ushort(mine);
            named_bar_sync(1 + q4, 64);
            const float m_blk = fmaxf(__bfloat162float(mine), __bfloat162float(__ushort_as_bfloat16(sX[(hf ^ 1) * TQ + r])));
            const bool bump = (j == 0) || (m_blk > m_run + FWD_RESCALE_THRESH);
            const float m_new = bump ? fmaxf(m_run, m_blk) : m_run;
            const float alpha = bump ? ex2(m_run - m_new) : 1.0f;     // 0 at j == 0 (m_run = -inf)
            if (j >= 1) {                                              // PV_{j-1} retired: O is valid, the P buffer is free
                mbar_wait(pv_done, (j - 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.0f)) {          // rescale my 32 columns of this row of O
                    tmem_ld32(tO, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                    tmem_st32(tO, v);
                    tmem_st_wait();
                }
            }
            // ---- pass 2: p = exp2(s * scale - m), row sum (before dropout), dropout, bf16 -> swizzled smem
            const float neg_m = -m_new;
            float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
            for (int pc = 0; pc < 2; ++pc) {
                tmem_ld32(tS + pc * 32, v);
                tmem_ld_wait();
                if (pc == 1) {                                         // my half of S_j is in registers: its columns may take S_{j+1}
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_free[hf]);
                }
                float e[32];
                if (masked) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 m4 = __ldg(reinterpret_cast<const float4*>(mk + pc * 32) + i);
                        e[4 * i] = ex2(fmaf(__uint_as_float(v[4 * i]), p.scale_log2, m4.x * LOG2E) + neg_m);
                        e[4 * i + 1] = ex2(fmaf(__uint_as_float(v[4 * i + 1]), p.scale_log2, m4.y * LOG2E) + neg_m);
                        e[4 * i + 2] = ex2(fmaf(__uint_as_float(v[4 * i + 2]), p.scale_log2, m4.z * LOG2E) + neg_m);
                        e[4 * i + 3] = ex2(fmaf(__uint_as_float(v[4 * i + 3]), p.scale_log2, m4.w * LOG2E) + neg_m);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        float x0, x1;
                        ffma2(x0, x1, __uint_as_float(v[i]), __uint_as_float(v[i + 1]), p.scale_log2, p.scale_log2, neg_m, neg_m);
                        e[i] = ex2(x0); e[i + 1] = ex2(x1);
                    }
                }
#pragma unroll
                for (int i = 0; i < 32; i += 2) fadd2(rs0, rs1, rs0, rs1, e[i], e[i + 1]);
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] = pack_bf16(e[2 * i], e[2 * i + 1]);
                const int col = hf * 64 + pc * 32;
                if (p.drop_on != 0u) {                                 // the 1/(1-p) scale is applied once, in the epilogue
                    uint32_t km[16];
                    attn_dropout_masks16(seed, p.drop_stream, (drop_row + j * TQ + pc * 32) >> 5, p.drop_k2, km);
#pragma unroll
                    for (int i = 0; i < 16; ++i) pk[i] &= km[i];
                }
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    st_shared_v4(sp + pt_offset(r, col + g * 8), pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
            }
            const float rs = rs0 + rs1;
            l0 = l0 * alpha + rs;
            m_run = m_new;
            tc_fence_before();                                         // orders the O rescale (tcgen05.st) before PV_j
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);                       // one arrival per warp (count = 8)
        }
        // ---- epilogue: O * (1/(1-p)) / l -> ctx, lse
        mbar_wait(pv_done, (n_chunks - 1) & 1);
        tc_fence_after();
        float* xl = reinterpret_cast<float*>(sP);                      // the P buffer is free: exchange the two partial row sums
        xl[hf * TQ + r] = l0;
        named_bar_sync(1 + q4, 64);
        const float l_tot = l0 + xl[(hf ^ 1) * TQ + r];
        const float inv_l = p.drop_scale / l_tot;
        {
            // O rows go through the (idle) P~ buffer so that every global store instruction covers 4 rows x 128 contiguous bytes instead of
            // 32 lines 2 KB apart (see the backward drains): the stores of one CTA no longer hold the SM's LSU while the other CTA runs
            const uint32_t stage = smem_u32(sP) + 2048;                // past the 1 KB row-sum exchange above
            uint32_t v[32];
            tmem_ld32(tO, v);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 4; ++k)
                st_shared_v4(stage + drain_off(r, 4 * hf + k), pack_bf16(__uint_as_float(v[8 * k]) * inv_l, __uint_as_float(v[8 * k + 1]) * inv_l),
                             pack_bf16(__uint_as_float(v[8 * k + 2]) * inv_l, __uint_as_float(v[8 * k + 3]) * inv_l),
                             pack_bf16(__uint_as_float(v[8 * k + 4]) * inv_l, __uint_as_float(v[8 * k + 5]) * inv_l),
                             pack_bf16(__uint_as_float(v[8 * k + 6]) * inv_l, __uint_as_float(v[8 * k + 7]) * inv_l));
            named_bar_sync(5, FWD_SOFTMAX_WARPS * 32);
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int row = warp * 16 + it * 4 + (lane >> 3), ch = lane & 7;
                const uint4 w = lds_u4(stage + drain_off(row, ch));
                const long long tok = (long long)b * p.tok_stride_b + (long long)(qt * TQ + row) * p.tok_stride_s;
                st_global_v4(p.ctx + tok * p.H + h * HD + ch * 8, w.x, w.y, w.z, w.w);
            }
        }
        if (hf == 0) p.lse[((long long)b * p.A + h) * S + qt * TQ + r] = (m_run + log2f(l_tot)) * LN2;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == FWD_WARP_TMA) { tc_fence_after(); tmem_dealloc(tmem_base, FWD_TMEM_COLS); }
}

// =================================================================================================
// backward  (v4)
//   CTA   : one (batch, head); loops kv tiles j (outer) x query tiles i (inner), pair index t = j*n + i.
//   smem  : K_j, V_j 32 KB + Q_i, dO_i of ALL query tiles resident (<= 128 KB, loaded once per head) + P~ 32 KB + dS 32 KB + mask
//           = 226.3 KB at S = 512 (one CTA per SM).
//   TMEM  : dQ of all query tiles 4 x 64 | dK 64 | dV 64 | S half 64 | dP half 64 = 512 columns: every accumulator stays on chip
//           (no atomics, no scratch in HBM, deterministic), and S / dP are produced 64 KEY COLUMNS AT A TIME so that they fit beside them.
//   warps : 0..15 = two compute groups of 8 warps, 16 = TMA producer, 17 = MMA issuer: group g owns key half g of every pair
//           (thread = one query row x 32 keys).  The groups run half a pair apart, so while one computes P / dS from registers the
//           tensor core produces the other's S / dP: v3 measured no gain from overlapping whole-tile phases because its dQ went through
//           16 B/clk/SM of L2 atomics; v1 kept dQ on chip but had S and dP share columns, serialising every MMA behind a compute phase.
//   early release : a TMEM buffer is handed back (s_free / dp_free) as soon as its columns are in registers, not after they are used.
//   MMA order per pair : S(t,0) dP(t,0) S(t,1) dV(t-1) dP(t,1) dK(t-1) dQ(t-1) -- the accumulating MMAs trail by one pair, so the issuer
//           never blocks the next pair's S / dP behind operands (P~, dS) that the compute warps are still writing.
// =================================================================================================
// Two MMA-issuing warps.  The hand-off trace (tools/attn_trace.py) shows one in-order issuer spending ~3100 of
// the ~5500 clk pair period inside tcgen05.mma issue (~80 clk per 128x64x16 instruction) and the accumulating MMAs of pair t-1 reaching
// the tensor pipe ~4000 clk after their operands were ready, because they queue behind S / dP issues that wait on the compute warps.
// Warp 17 therefore issues only S / dP, warp 18 only dV / dK / dQ: each follows its own operands (measured 911 -> 857 us at B=128).
constexpr int BWD_THREADS = 608;       // warps 0..15: compute, 16: TMA + TMEM alloc, 17: S/dP issuer, 18: dV/dK/dQ issuer
constexpr int BWD_COMPUTE_WARPS = 16;
constexpr int BWD_WARP_TMA = 16, BWD_WARP_MMA = 17, BWD_WARP_MMA2 = 18;
constexpr int BWD_GROUP_WARPS = 8;

struct AttnBwdParams {
    const float* mask; const float* lse; const float* delta;
    float* dbias;          // [3H] fp32 or null: += column sums of dqkv
    bf16* dqkv;            // [T, 3H]
    int B, S, A, H;
    int tok_stride_s, tok_stride_b;
    float scale, scale_log2;
    uint32_t drop_k2; uint32_t drop_on; float drop_scale; uint32_t drop_stream; unsigned long long seed; const unsigned long long* seed_dev;
};

__host__ __device__ inline int bwd_smem_bytes(int S) {
    const int n = S / TQ;
    return 4 * TILE_BYTES /*K,V x 2 slots*/ + n * TILE_BYTES /*Q of every query tile*/ + 2 * TILE_BYTES /*dO ring*/ + 2 * PT_BYTES /*P,dS*/ + S * 4 + 256;
}

// delta[b,h,s] = sum_d dO[t, h*64+d] * O[t, h*64+d].  Eight lanes share one (token, head) row: consecutive lanes read consecutive 16-byte
// units, so a warp instruction covers 4 whole 128-byte lines (one thread per row touched 32 lines per instruction and ran at 3.7 TB/s).
__global__ void attn_delta_kernel(const bf16* __restrict__ dctx, const bf16* __restrict__ ctx, float* __restrict__ delta,
                                  int B, int S, int A, int seq_first) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long idx = gid >> 3;                                            // token * A + h
    const int unit = (int)(gid & 7);
    const long long total = (long long)B * S * A;
    float acc = 0.f;
    if (idx < total) {
        const uint4 ua = ld_global_nc_v4(dctx + idx * HD + unit * 8), uc = ld_global_nc_v4(ctx + idx * HD + unit * 8);
        float2 x, y;
        x = unpack_bf16(ua.x); y = unpack_bf16(uc.x); acc += x.x * y.x + x.y * y.y;
        x = unpack_bf16(ua.y); y = unpack_bf16(uc.y); acc += x.x * y.x + x.y * y.y;
        x = unpack_bf16(ua.z); y = unpack_bf16(uc.z); acc += x.x * y.x + x.y * y.y;
        x = unpack_bf16(ua.w); y = unpack_bf16(uc.w); acc += x.x * y.x + x.y * y.y;
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (idx < total && unit == 0) {
        const long long tok = idx / A; const int h = (int)(idx - tok * A);
        const int b = seq_first ? (int)(tok % B) : (int)(tok / S);
        const int s_ = seq_first ? (int)(tok / B) : (int)(tok - (long long)b * S);
        delta[((long long)b * A + h) * S + s_] = acc;
    }
}

// Measurement-only build switch (-DDLE_ATTN_TRACE, csrc/build.py --variant-trace): lane 0 of the MMA warp and of one compute warp per group
// of ONE CTA stamps clock64() at every barrier hand-off into a __device__ buffer, and every CTA records its SM and start / end
// %globaltimer.  tools/attn_trace.py turns that into the per-pair timeline DESIGN.md section 8 discusses.  Not compiled into the product.
#ifdef DLE_ATTN_TRACE
constexpr int TRACE_CAP = 4096, TRACE_CTA = 709, TRACE_MAX_CTAS = 8192;
__device__ unsigned long long g_attn_trace[4][TRACE_CAP];
__device__ int g_attn_trace_n[4];
__device__ unsigned long long g_attn_cta[TRACE_MAX_CTAS][3];
#define TR(code, t) do { if (tr_slot >= 0 && lane == 0 && tr_n < TRACE_CAP) g_attn_trace[tr_slot][tr_n++] = ((unsigned long long)clock64() << 16) | ((unsigned long long)(code) << 8) | (unsigned long long)((t) & 255); } while (0)
#define TR_END() do { if (tr_slot >= 0 && lane == 0) g_attn_trace_n[tr_slot] = tr_n; } while (0)
#else
#define TR(code, t) do {} while (0)
#define TR_END() do {} while (0)
#endif

// (608 threads leave 96 registers per thread, not 104: warps are placed 5 / 5 / 5 / 4 on the four 16 K-register sub-partitions, and a
// __maxnreg__(104) build fails to launch)
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do, const AttnBwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int S = p.S, n = S / TQ;
    uint8_t* sK = smem;                          // [2]  kv tile j lives in slot j & 1
    uint8_t* sV = sK + 2 * TILE_BYTES;           // [2]
    uint8_t* sQ = sV + 2 * TILE_BYTES;           // [n]  resident for the whole head
    uint8_t* sdO = sQ + n * TILE_BYTES;          // [2]  ring: the dO tile of pair t lives in slot t & 1
    uint8_t* sP = sdO + 2 * TILE_BYTES;
    uint8_t* sdS = sP + PT_BYTES;
    float* sMask = reinterpret_cast<float*>(sdS + PT_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sMask + S);
    uint64_t* kv_full = bars;         // [2]  K_j, V_j landed in slot j & 1   (completion j >> 1 of that slot)
    uint64_t* kv_empty = bars + 2;    // [2]  every MMA that reads the slot has retired
    uint64_t* q_full = bars + 4;      // [4]  Q_i landed (once per head)
    uint64_t* s_full = bars + 8;      // [2]  S half g in TMEM            (per pair)
    uint64_t* dp_full = bars + 10;    // [2]  dP half g in TMEM           (per pair)
    uint64_t* s_free = bars + 12;     // [2]  group g has S half g in registers   (8 warp arrivals per pair)
    uint64_t* dp_free = bars + 14;    // [2]
    uint64_t* p_full = bars + 16;     // P~ tile complete in smem         (16 warp arrivals per pair)
    uint64_t* ds_full = bars + 17;    // dS tile complete in smem         (16)
    uint64_t* dv_done = bars + 18;    // dV(t) retired: sP may be rewritten
    uint64_t* pair_done = bars + 19;  // dK(t), dQ(t) retired: sdS may be rewritten
    uint64_t* dkv_full = bars + 20;   // kv tile finished: dK, dV (and at the end dQ) complete
    uint64_t* dkv_read = bars + 21;   // accumulators drained (16)
    uint64_t* nk_ready = bars + 22;   // number of kv tiles to process is known (16 compute-warp arrivals)
    uint64_t* do_full = bars + 23;    // [2]  dO tile of pair t landed in slot t & 1   (completion t >> 1 of that slot)
    uint64_t* do_empty = bars + 25;   // [2]  2 arrivals: dP(t, 1) retired (S/dP issuer) and dV(t) retired (accumulating issuer)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 27);
    int* s_nk = reinterpret_cast<int*>(bars + 27) + 1;      // kv tiles that hold at least one attendable key
    volatile int* s_nk_loop = reinterpret_cast<volatile int*>(bars + 28);   // the compute warps' loop bound, re-read from here (see below)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.x, b = blockIdx.y;
#ifdef DLE_ATTN_TRACE
    const int cta_lin = blockIdx.y * gridDim.x + blockIdx.x;
    const int tr_slot = (cta_lin != TRACE_CTA) ? -1 : (warp == BWD_WARP_MMA ? 0 : warp == 0 ? 1 : warp == 8 ? 2 : warp == BWD_WARP_MMA2 ? 3 : -1);
    int tr_n = 0;
    if (threadIdx.x == 0 && cta_lin < TRACE_MAX_CTAS) {
        unsigned int smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_attn_cta[cta_lin][0] = smid; g_attn_cta[cta_lin][1] = global_timer_ns();
    }
    TR(1, 0);
#endif

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) __trap();          // SWIZZLE_128B tiles need a 1024-byte aligned base
        tma_prefetch_desc(&tmap_qkv); tma_prefetch_desc(&tmap_do);
        for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); mbar_init(&do_full[i], 1); mbar_init(&do_empty[i], 2); }
        for (int i = 0; i < 4; ++i) mbar_init(&q_full[i], 1);
        for (int g = 0; g < 2; ++g) {
            mbar_init(&s_full[g], 1); mbar_init(&dp_full[g], 1);
            mbar_init(&s_free[g], BWD_GROUP_WARPS); mbar_init(&dp_free[g], BWD_GROUP_WARPS);
        }
        mbar_init(p_full, BWD_COMPUTE_WARPS); mbar_init(ds_full, BWD_COMPUTE_WARPS);
        mbar_init(dv_done, 1); mbar_init(pair_done, 1); mbar_init(dkv_full, 1); mbar_init(dkv_read, BWD_COMPUTE_WARPS);
        mbar_init(nk_ready, BWD_COMPUTE_WARPS);
        fence_barrier_init();
        *s_nk = 0;
    }
    if (warp == BWD_WARP_TMA) { tmem_alloc(tmem_ptr, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // trailing kv tiles that are masked out for every key (additive mask <= -1000) have P == 0 exactly: dK = dV = 0 there and dQ gets
    // nothing from them, so they are skipped and their dK / dV rows are written as zeros (bit-identical to computing them).  The compute
    // warps find the count while they stage the mask row; the producer / MMA warps pick it up (nk_ready) before they go past kv tile 0.
    const bool scan = p.mask != nullptr && n > 1;
    auto tiles_to_do = [&]() -> int {
        if (!scan) return n;
        mbar_wait(nk_ready, 0);
        const int v_ = *reinterpret_cast<volatile int*>(s_nk);
        return v_ > 0 ? v_ : n;
    };
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_dQ = tmem_base, tmem_dK = tmem_base + 256, tmem_dV = tmem_base + 320, tmem_S = tmem_base + 384, tmem_dP = tmem_base + 448;

    // Producer and MMA warps: warp-uniform loops (all lanes wait), tcgen05 / TMA instructions under elect_one() -- see gemm_sm100.cu.
    if (warp == BWD_WARP_TMA) {
        // ===================== TMA producer =====================
        // K_j / V_j are double-buffered (tile j+1 is fetched while tile j is processed: the trace showed the single buffer costing ~6000 clk
        // per kv tile -- drain of the last MMAs, then 32 KB of TMA, then the first S -- a quarter of the kernel), and the room comes from
        // streaming dO through a two-slot ring instead of keeping it resident beside Q.
        auto load_kv = [&](int j_) {
            const int sl = j_ & 1;
            mbar_expect_tx(&kv_full[sl], 2 * TILE_BYTES);
            tma_load_3d(sK + sl * TILE_BYTES, &tmap_qkv, &kv_full[sl], p.H + h * HD, j_ * TQ, b);
            tma_load_3d(sV + sl * TILE_BYTES, &tmap_qkv, &kv_full[sl], 2 * p.H + h * HD, j_ * TQ, b);
        };
        auto load_do = [&](int u_) {                         // dO tile of pair u_ (query tile u_ % n)
            const int sl = u_ & 1;
            mbar_expect_tx(&do_full[sl], TILE_BYTES);
            tma_load_3d(sdO + sl * TILE_BYTES, &tmap_do, &do_full[sl], h * HD, (u_ % n) * TQ, b);
        };
        if (elect_one()) {
            load_kv(0);
            mbar_expect_tx(&q_full[0], TILE_BYTES);
            tma_load_3d(sQ, &tmap_qkv, &q_full[0], h * HD, 0, b);
            load_do(0);
            for (int i = 1; i < n; ++i) {
                mbar_expect_tx(&q_full[i], TILE_BYTES);
                tma_load_3d(sQ + i * TILE_BYTES, &tmap_qkv, &q_full[i], h * HD, i * TQ, b);
            }
            if (n >= 2) load_do(1);
        }
        __syncwarp();
        const int nk = tiles_to_do();
        const int total = nk * n;
        if (nk > 1) {
            if (elect_one()) load_kv(1);
            __syncwarp();
        }
        for (int u = (n >= 2) ? 2 : 1; u < total; ++u) {
            if (u >= 2) mbar_wait(&do_empty[u & 1], (uint32_t)((u >> 1) - 1) & 1u);     // pair u-2 no longer reads the slot
            if (elect_one()) load_do(u);
            __syncwarp();
            const int j = u / n;
            if (u - j * n == 0 && j >= 1 && j + 1 < nk) {    // tile j begins: tile j-1 is finishing, its slot takes tile j+1
                mbar_wait(&kv_empty[(j + 1) & 1], (uint32_t)((j - 1) >> 1) & 1u);
                if (elect_one()) load_kv(j + 1);
                __syncwarp();
            }
        }
    } else if (warp == BWD_WARP_MMA || warp == BWD_WARP_MMA2) {
        // ===================== MMA issuer(s) =====================
        constexpr uint32_t id_h = make_idesc_bf16(TQ, 64, false, false);      // S half, dP half: [128 q] x [64 keys], K = d
        constexpr uint32_t id_mm = make_idesc_bf16(TQ, HD, true, true);       // dV, dK
        constexpr uint32_t id_km = make_idesc_bf16(TQ, HD, false, true);      // dQ
        const uint32_t aK0 = smem_u32(sK), aV0 = smem_u32(sV), aQ0 = smem_u32(sQ), adO0 = smem_u32(sdO), aP = smem_u32(sP), adS = smem_u32(sdS);
        auto issue_s = [&](int i_, int j_, int g_) {        // S(t, g) = Q_i K_{j, half g}^T
            if (elect_one()) {
                const uint32_t aQ = aQ0 + i_ * TILE_BYTES, aKh = aK0 + (j_ & 1) * TILE_BYTES + g_ * HALF_BYTES;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16_ss(tmem_S, make_smem_desc_sw128(aQ + kk * 32, 0, 1024), make_smem_desc_sw128(aKh + kk * 32, 0, 1024), id_h, kk > 0 ? 1u : 0u);
                umma_commit(&s_full[g_]);
            }
            __syncwarp();
        };
        auto issue_dp = [&](int t_, int j_, int g_) {       // dP(t, g) = dO_i V_{j, half g}^T
            if (elect_one()) {
                const uint32_t adO = adO0 + (t_ & 1) * TILE_BYTES, aVh = aV0 + (j_ & 1) * TILE_BYTES + g_ * HALF_BYTES;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16_ss(tmem_dP, make_smem_desc_sw128(adO + kk * 32, 0, 1024), make_smem_desc_sw128(aVh + kk * 32, 0, 1024), id_h, kk > 0 ? 1u : 0u);
                umma_commit(&dp_full[g_]);
                if (g_ == 1) umma_commit(&do_empty[t_ & 1]);          // this issuer's last read of the dO slot
            }
            __syncwarp();
        };
        auto issue_dv = [&](int i_, int t_) {               // dV_j += P~^T dO_i
            if (elect_one()) {
                const uint32_t adO = adO0 + (t_ & 1) * TILE_BYTES;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                    umma_bf16_ss(tmem_dV, make_smem_desc_sw128(aP + kk * 2048, TILE_BYTES, 1024),
                                 make_smem_desc_sw128(adO + kk * 2048, TILE_BYTES, 1024), id_mm, (i_ > 0 || kk > 0) ? 1u : 0u);
                umma_commit(dv_done);
                umma_commit(&do_empty[t_ & 1]);
            }
            __syncwarp();
        };
        auto issue_dkdq = [&](int i_, int j_) {             // dK_j += dS^T Q_i ; dQ_i += dS K_j
            if (elect_one()) {
                const uint32_t aQ = aQ0 + i_ * TILE_BYTES, aK = aK0 + (j_ & 1) * TILE_BYTES;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                    umma_bf16_ss(tmem_dK, make_smem_desc_sw128(adS + kk * 2048, TILE_BYTES, 1024),
                                 make_smem_desc_sw128(aQ + kk * 2048, TILE_BYTES, 1024), id_mm, (i_ > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                    umma_bf16_ss(tmem_dQ + i_ * HD, make_smem_desc_sw128(adS + (kk >> 2) * TILE_BYTES + (kk & 3) * 32, 0, 1024),
                                 make_smem_desc_sw128(aK + kk * 2048, TILE_BYTES, 1024), id_km, (j_ > 0 || kk > 0) ? 1u : 0u);
                umma_commit(pair_done);
            }
            __syncwarp();
        };
        auto wait = [&](uint64_t* bar, int phase) { mbar_wait(bar, (uint32_t)phase & 1u); tc_fence_after(); };
        int nk = n;
        if (warp == BWD_WARP_MMA) {
            // ---- S / dP issuer: paced only by the compute warps taking the previous halves out of TMEM
            for (int j = 0; j < nk; ++j) {
                if (j == 1) nk = tiles_to_do();
                if (j >= nk) break;
                wait(&kv_full[j & 1], j >> 1); TR(10, j);
                for (int i = 0; i < n; ++i) {
                    const int t = j * n + i;
                    if (j == 0) { wait(&q_full[i], 0); TR(12, t); }
                    if (t >= 1) wait(&s_free[1], t - 1);
                    TR(13, t);
                    issue_s(i, j, 0); TR(14, t);
                    if (t >= 1) wait(&dp_free[1], t - 1);
                    wait(&do_full[t & 1], t >> 1);
                    TR(15, t);
                    issue_dp(t, j, 0); TR(16, t);
                    wait(&s_free[0], t); TR(17, t);
                    issue_s(i, j, 1); TR(18, t);
                    wait(&dp_free[0], t); TR(21, t);
                    issue_dp(t, j, 1); TR(22, t);
                }
            }
        } else {
            // ---- accumulating MMAs: dV(t) as soon as P~(t) is in shared memory, dK / dQ(t) as soon as dS(t) is.  kv_empty / dkv_full are
            // committed here: dS of the tile's last pair exists only after every S / dP MMA of the tile has been consumed, so when the
            // last dK / dQ retire nothing of either issuer still reads K_j / V_j.
            for (int j = 0; j < nk; ++j) {
                if (j == 1) nk = tiles_to_do();
                if (j >= nk) break;
                wait(&kv_full[j & 1], j >> 1); TR(10, j);
                if (j >= 1) { wait(dkv_read, j - 1); TR(11, j); }
                for (int i = 0; i < n; ++i) {
                    const int t = j * n + i;
                    if (j == 0) wait(&q_full[i], 0);
                    wait(p_full, t);
                    wait(&do_full[t & 1], t >> 1); TR(19, t);
                    issue_dv(i, t); TR(20, t);
                    wait(ds_full, t); TR(23, t);
                    issue_dkdq(i, j); TR(24, t);
                }
                if (elect_one()) {
                    umma_commit(&kv_empty[j & 1]);
                    umma_commit(dkv_full);
                }
                __syncwarp();
            }
        }
        TR_END();
    } else {
        // ===================== compute warps =====================
        const int q4 = warp & 3;                                // TMEM lane quarter (hardware rule: warp id % 4)
        const int wi = warp >> 2;                               // 0..3
        const int g = wi >> 1, c = wi & 1;                      // key half of the pair this warp's group owns; 32-column half of that
        const int r = q4 * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
        const int ct = threadIdx.x;
        {
            int last = 0;
            for (int i = ct; i < S; i += BWD_COMPUTE_WARPS * 32) {
                const float mv = p.mask ? p.mask[(long long)b * S + i] : 0.f;
                sMask[i] = mv * LOG2E;
                if (mv > -1000.0f) last = i / TQ + 1;
            }
            if (scan) {
                last = __reduce_max_sync(0xffffffffu, last);
                if (lane == 0) { if (last > 0) atomicMax(s_nk, last); __threadfence_block(); mbar_arrive(nk_ready); }
            }
        }
        named_bar_sync(1, BWD_COMPUTE_WARPS * 32);
        // ptxas keeps this loop bound on the stack (the register budget is spent on the 32-wide row buffers), and its reload at the end of
        // each kv tile queued behind the drain's global stores: every thread parks the value in shared memory and re-reads it from there
        // (a thread reads back what it wrote itself, so no barrier is needed).
        *s_nk_loop = tiles_to_do();
        TR(30, 0);
        const long long bh = (long long)b * p.A + h;
        // per-row softmax statistics of this thread's row: the pair loop fetches the NEXT pair's two values while it works on the
        // current one (the trace showed each pair opening with ~1000 clk of waiting for these loads when they were issued in place,
        // and ~580 clk when all tiles' values were kept per thread -- they did not fit in registers and came back from local memory)
        float lse_nxt = p.lse[bh * S + r], dl_nxt = p.delta[bh * S + r];
        float dbias_acc = 0.f;                                  // key / value bias gradient of my column, summed over the kv tiles
        const uint32_t aP = smem_u32(sP), adS = smem_u32(sdS);
        const unsigned long long seed = effective_seed(p.seed, p.seed_dev);
        const int kc = g * 64 + c * 32;                         // first key column (within the 128-key tile) of this thread
        const float c1 = p.drop_scale * p.scale;
        // drain of a finished kv tile; called from inside the first pair of the NEXT tile (after its P is computed, before it is stored) so
        // that the wait for the tile's last dK / dQ MMAs is covered by that pair's S -> P work, and after the loop for the last tile
        auto drain_kv = [&](int jd) {
            // ---- dV_j, dK_j complete: group 0 drains dK, group 1 drains dV (32 columns per warp).  The rows go through shared memory
            // (P~ / dS staging is idle here: every MMA of the tile has retired) so that each global store instruction covers 4 rows x 128
            // contiguous bytes: written straight from the row-per-lane registers, every STG.128 touched 32 different lines 6 KB apart
            // and the 2048 line writes of one drain held the LSU for ~2000 clk -- the loads that open the next pair queued behind them
            // (trace: ~2300-4300 idle clk at every kv-tile boundary).
            mbar_wait(dkv_full, jd & 1);
            TR(42, jd);
            tc_fence_after();
            {
                uint32_t v[32];
                tmem_ld32((g == 0 ? tmem_dK : tmem_dV) + lane_addr + c * 32, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(dkv_read);            // the accumulator columns are in registers: the next tile may overwrite them
                if (g == 1 && p.drop_on != 0u) {                 // dV accumulated keep-mask AND P: apply the 1/(1-p) factor here
#pragma unroll
                    for (int k = 0; k < 32; ++k) v[k] = __float_as_uint(__uint_as_float(v[k]) * p.drop_scale);
                }
                const uint32_t stage = (g == 0) ? adS : aP;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    st_shared_v4(stage + drain_off(r, 4 * c + k), pack_bf16(__uint_as_float(v[8 * k]), __uint_as_float(v[8 * k + 1])),
                                 pack_bf16(__uint_as_float(v[8 * k + 2]), __uint_as_float(v[8 * k + 3])), pack_bf16(__uint_as_float(v[8 * k + 4]), __uint_as_float(v[8 * k + 5])),
                                 pack_bf16(__uint_as_float(v[8 * k + 6]), __uint_as_float(v[8 * k + 7])));
                if (p.dbias != nullptr) {                        // key / value bias gradients: column sums of the stored bf16 values
                    float f[32];
#pragma unroll
                    for (int k = 0; k < 32; ++k) f[k] = __bfloat162float(__float2bfloat16_rn(__uint_as_float(v[k])));
                    dbias_acc += warp_column_sums32(f, lane);    // one atomic per head (below), not one per kv tile on the tile boundary
                }
                named_bar_sync(2 + g, BWD_GROUP_WARPS * 32);     // the group's 128 x 64 tile is staged
                const int gw = c * 4 + q4;                       // warp index within the group: 16 rows each
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int row = gw * 16 + it * 4 + (lane >> 3), ch = lane & 7;
                    const uint4 w = lds_u4(stage + drain_off(row, ch));
                    const long long tok = (long long)b * p.tok_stride_b + (long long)(jd * TQ + row) * p.tok_stride_s;
                    st_global_v4(p.dqkv + tok * (3LL * p.H) + (g + 1) * p.H + h * HD + ch * 8, w.x, w.y, w.z, w.w);
                }
            }
            named_bar_sync(1, BWD_COMPUTE_WARPS * 32);           // both staging tiles are read: P~ / dS of the next pair may be written
            TR(43, jd);
        };
        for (int j = 0; j < *s_nk_loop; ++j) {
            // my 32 key columns of this kv tile: additive mask (already x log2e) from shared memory, skipped entirely when it is all
            // zero (warp-uniform; unpadded batches)
            const uint32_t mk_addr = smem_u32(sMask + j * TQ + kc);
            const bool masked = __any_sync(0xffffffffu, sMask[j * TQ + kc + lane] != 0.f);
            for (int i = 0; i < n; ++i) {
                const int t = j * n + i;
                const uint32_t ph = (uint32_t)t & 1u;
                const float neg_lse2 = -lse_nxt * LOG2E;
                const float nd = -dl_nxt * p.scale;
                {
                    const int i2 = (i + 1 == n) ? 0 : i + 1;                     // query tile of the next pair (harmless reload after the last)
                    lse_nxt = p.lse[bh * S + i2 * TQ + r];
                    dl_nxt = p.delta[bh * S + i2 * TQ + r];
                }
                const unsigned long long drop_row = (unsigned long long)(bh * S + (i * TQ + r)) * (unsigned long long)S + j * TQ + kc;
                uint32_t pk[16];                 // undropped P, packed bf16x2 (32 values)
                uint32_t km[16];                 // keep-masks of my 32 columns (bf16x2 AND-masks)
                uint32_t v[32];
                // ---- P = exp2(S * scale + mask - lse)
                TR(31, t);
                mbar_wait(&s_full[g], ph);
                TR(32, t);
                tc_fence_after();
                tmem_ld32(tmem_S + lane_addr + c * 32, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_free[g]);         // the S columns may be overwritten (other half / next pair)
                TR(33, t);
                {
                    float e[32];
                    if (masked) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float m0, m1, m2, m3;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(m0), "=f"(m1), "=f"(m2), "=f"(m3) : "r"(mk_addr + q * 16));
                            e[4 * q + 0] = ex2(fmaf(__uint_as_float(v[4 * q + 0]), p.scale_log2, m0) + neg_lse2);
                            e[4 * q + 1] = ex2(fmaf(__uint_as_float(v[4 * q + 1]), p.scale_log2, m1) + neg_lse2);
                            e[4 * q + 2] = ex2(fmaf(__uint_as_float(v[4 * q + 2]), p.scale_log2, m2) + neg_lse2);
                            e[4 * q + 3] = ex2(fmaf(__uint_as_float(v[4 * q + 3]), p.scale_log2, m3) + neg_lse2);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 32; k += 2) {
                            float x0, x1;
                            ffma2(x0, x1, __uint_as_float(v[k]), __uint_as_float(v[k + 1]), p.scale_log2, p.scale_log2, neg_lse2, neg_lse2);
                            e[k] = ex2(x0); e[k + 1] = ex2(x1);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 16; ++k) pk[k] = pack_bf16(e[2 * k], e[2 * k + 1]);
                }
                if (p.drop_on != 0u) attn_dropout_masks16(seed, p.drop_stream, drop_row >> 5, p.drop_k2, km);
                TR(34, t);
                if (i == 0 && j > 0) drain_kv(j - 1);
                if (t >= 1) { mbar_wait(dv_done, ph ^ 1u); tc_fence_after(); }          // dV(t-1) retired: sP may be overwritten
                TR(35, t);
                // P~ = keep-mask AND P: the 1/(1-p) factor is folded into the dV drain and into the dS constants below
                if (p.drop_on != 0u) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        st_shared_v4(aP + pt_offset(r, kc + q * 8), pk[q * 4] & km[q * 4], pk[q * 4 + 1] & km[q * 4 + 1],
                                     pk[q * 4 + 2] & km[q * 4 + 2], pk[q * 4 + 3] & km[q * 4 + 3]);
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        st_shared_v4(aP + pt_offset(r, kc + q * 8), pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full);
                TR(36, t);
                // ---- dS = [ (mask & P) * dP / (1-p) - P * delta ] * scale
                mbar_wait(&dp_full[g], ph);
                TR(37, t);
                tc_fence_after();
                tmem_ld32(tmem_dP + lane_addr + c * 32, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&dp_free[g]);
                TR(38, t);
                uint32_t ds[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const uint32_t pm = (p.drop_on != 0u) ? (pk[k] & km[k]) : pk[k];
                    float t0, t1, u0, u1, d0, d1;
                    fmul2(t0, t1, __uint_as_float(pm << 16), __uint_as_float(pm & 0xFFFF0000u), __uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1]));
                    fmul2(u0, u1, __uint_as_float(pk[k] << 16), __uint_as_float(pk[k] & 0xFFFF0000u), nd, nd);
                    ffma2(d0, d1, t0, t1, c1, c1, u0, u1);
                    ds[k] = pack_bf16(d0, d1);
                }
                TR(39, t);
                if (t >= 1) { mbar_wait(pair_done, ph ^ 1u); tc_fence_after(); }        // dK / dQ(t-1) retired: sdS may be overwritten
                TR(40, t);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    st_shared_v4(adS + pt_offset(r, kc + q * 8), ds[q * 4], ds[q * 4 + 1], ds[q * 4 + 2], ds[q * 4 + 3]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(ds_full);
                TR(41, t);
            }
            if (j + 1 == *s_nk_loop) drain_kv(j);              // the last tile has no following pair to hide behind
        }
        if (p.dbias != nullptr) atomicAdd(p.dbias + (g + 1) * p.H + h * HD + c * 32 + lane, dbias_acc);
        // ---- skipped (fully masked) kv tiles: zero dK / dV rows
        for (int j = *s_nk_loop; j < n; ++j) {
            const long long tok = (long long)b * p.tok_stride_b + (long long)(j * TQ + r) * p.tok_stride_s;
            bf16* o = p.dqkv + tok * (3LL * p.H) + (g + 1) * p.H + h * HD + c * 32;
#pragma unroll
            for (int k = 0; k < 32; k += 8) st_global_v4(o + k, 0u, 0u, 0u, 0u);
        }
        // ---- all pairs done (dkv_full of the last kv tile implies every MMA retired): drain dQ, two 32-column chunks per warp, staged the
        // same way (query tile i in the i-th 16 KB quarter of the P~ / dS staging area)
        for (int ch = wi; ch < 2 * n; ch += 4) {
            const int i = ch >> 1, hf = ch & 1;
            uint32_t v[32];
            tmem_ld32(tmem_dQ + i * HD + lane_addr + hf * 32, v);
            tmem_ld_wait();
            const uint32_t stage = (i < 2 ? aP : adS) + (i & 1) * (TQ * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                st_shared_v4(stage + drain_off(r, 4 * hf + k), pack_bf16(__uint_as_float(v[8 * k]), __uint_as_float(v[8 * k + 1])),
                             pack_bf16(__uint_as_float(v[8 * k + 2]), __uint_as_float(v[8 * k + 3])), pack_bf16(__uint_as_float(v[8 * k + 4]), __uint_as_float(v[8 * k + 5])),
                             pack_bf16(__uint_as_float(v[8 * k + 6]), __uint_as_float(v[8 * k + 7])));
            if (p.dbias != nullptr) {                            // query bias gradient
                float f[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) f[k] = __bfloat162float(__float2bfloat16_rn(__uint_as_float(v[k])));
                const float cs = warp_column_sums32(f, lane);
                atomicAdd(p.dbias + h * HD + hf * 32 + lane, cs);
            }
        }
        named_bar_sync(1, BWD_COMPUTE_WARPS * 32);
        for (int it = 0; it < 8; ++it) {
            const int R = warp * 32 + it * 4 + (lane >> 3), ch = lane & 7;      // row of the head's [S, 64] dQ, 16-byte unit within it
            if (R < S) {
                const int i = R >> 7, row = R & (TQ - 1);
                const uint4 w = lds_u4((i < 2 ? aP : adS) + (i & 1) * (TQ * 128) + drain_off(row, ch));
                const long long tok = (long long)b * p.tok_stride_b + (long long)R * p.tok_stride_s;
                st_global_v4(p.dqkv + tok * (3LL * p.H) + h * HD + ch * 8, w.x, w.y, w.z, w.w);
            }
        }
        TR(44, 0);
        TR_END();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == BWD_WARP_TMA) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
#ifdef DLE_ATTN_TRACE
    if (threadIdx.x == 0 && cta_lin < TRACE_MAX_CTAS) g_attn_cta[cta_lin][2] = global_timer_ns();
#endif
}

}  // namespace dle

using namespace dle;

// 3-D view {cols, S, B} of a [tokens, cols] bf16 matrix whose token rows are ordered b*S+s (seq_first=0)
// or s*B+b (seq_first=1, the reference's [S,B,H] convention); box = {64 cols, 128 s, 1 b}
static int make_tmap_tokens_3d(CUtensorMap* out, const void* base, int B, int S, int cols, int seq_first) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (enc == nullptr) return DLE_ERR_CUDA;
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (cols * 2) % 16 != 0) return DLE_ERR_INVALID;
    const cuuint64_t row_bytes = (cuuint64_t)cols * 2;
    cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)S, (cuuint64_t)B};
    cuuint64_t gstride[2] = {seq_first ? row_bytes * B : row_bytes, seq_first ? row_bytes : row_bytes * S};
    cuuint32_t box[3] = {HD, TQ, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? DLE_OK : DLE_ERR_CUDA;
}

// dropout probability -> the kernels' parameters.  p is quantised to 1/32768 (common.cuh: attn_dropout_masks16); the rescale uses the
// quantised value so that E[P~] = P exactly.
static void attn_drop_params(float dropout_p, uint32_t* k2, uint32_t* on, float* scale) {
    const uint32_t t15 = dropout_p > 0.f ? dropout_thresh15(dropout_p) : 0u;
    *on = t15 != 0u ? 1u : 0u;
    *k2 = attn_dropout_k2(t15);
    *scale = t15 != 0u ? 1.0f / (1.0f - (float)t15 / 32768.0f) : 1.0f;
}
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: remember which devices have it (per kernel)
struct SmemAttrCache { int bytes[64] = {0}; };
template <typename K>
static int ensure_smem_attr(K kern, SmemAttrCache& c, int bytes, bool carveout) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return DLE_ERR_CUDA;
    if (bytes > c.bytes[dev]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return DLE_ERR_CUDA;
        if (carveout) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        c.bytes[dev] = bytes;
    }
    return DLE_OK;
}

static int attn_check(int B, int S, int A) {
    if (B <= 0 || A <= 0 || S <= 0 || S % TQ != 0 || S > 512) return DLE_ERR_INVALID;
    return DLE_OK;
}

extern "C" int dle_attn_fwd(const void* qkv, const float* mask, void* ctx, float* lse, int32_t B, int32_t S, int32_t A,
                            int32_t seq_first, float dropout_p, uint64_t seed, const uint64_t* seed_dev, uint32_t dropout_stream, void* stream) {
    DLE_CHECK_ARG(qkv && ctx && lse && attn_check(B, S, A) == DLE_OK && dropout_p >= 0.f && dropout_p < 1.f);
    const int H = A * HD;
    CUtensorMap tm;
    int rc = make_tmap_tokens_3d(&tm, qkv, B, S, 3 * H, seq_first);
    if (rc != DLE_OK) return rc;
    AttnFwdParams p;
    p.tok_stride_s = seq_first ? B : 1; p.tok_stride_b = seq_first ? 1 : S;
    p.mask = mask; p.ctx = reinterpret_cast<bf16*>(ctx); p.lse = lse; p.B = B; p.S = S; p.A = A; p.H = H;
    p.scale_log2 = 0.125f * LOG2E;
    attn_drop_params(dropout_p, &p.drop_k2, &p.drop_on, &p.drop_scale);
    p.drop_stream = dropout_stream; p.seed = seed; p.seed_dev = reinterpret_cast<const unsigned long long*>(seed_dev);
    static SmemAttrCache attr;
    rc = ensure_smem_attr(attn_fwd_kernel, attr, FWD_SMEM_BYTES, true);
    if (rc != DLE_OK) return rc;
    attn_fwd_kernel<<<dim3(S / TQ, A, B), FWD_THREADS, FWD_SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(tm, p);
    DLE_LAUNCH_CHECK();
    return DLE_OK;
}

extern "C" int dle_attn_bwd(const void* qkv, const float* mask, const void* ctx, const void* dctx, const float* lse, void* dqkv,
                            float* delta_ws, float* dbias_qkv, int32_t B, int32_t S, int32_t A, int32_t seq_first, float dropout_p, uint64_t seed,
                            const uint64_t* seed_dev, uint32_t dropout_stream, void* stream) {
    DLE_CHECK_ARG(qkv && ctx && dctx && lse && dqkv && delta_ws && attn_check(B, S, A) == DLE_OK && dropout_p >= 0.f && dropout_p < 1.f);
    const int H = A * HD;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CUtensorMap tq, td;
    int rc = make_tmap_tokens_3d(&tq, qkv, B, S, 3 * H, seq_first);
    if (rc != DLE_OK) return rc;
    rc = make_tmap_tokens_3d(&td, dctx, B, S, H, seq_first);
    if (rc != DLE_OK) return rc;
    const long long total = (long long)B * S * A;
    attn_delta_kernel<<<(unsigned)((total * 8 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const bf16*>(dctx), reinterpret_cast<const bf16*>(ctx), delta_ws, B, S, A, seq_first);
    DLE_LAUNCH_CHECK();
    AttnBwdParams p;
    p.mask = mask; p.lse = lse; p.delta = delta_ws; p.dbias = dbias_qkv; p.dqkv = reinterpret_cast<bf16*>(dqkv);
    p.B = B; p.S = S; p.A = A; p.H = H; p.tok_stride_s = seq_first ? B : 1; p.tok_stride_b = seq_first ? 1 : S; p.scale = 0.125f; p.scale_log2 = 0.125f * LOG2E;
    attn_drop_params(dropout_p, &p.drop_k2, &p.drop_on, &p.drop_scale);
    p.drop_stream = dropout_stream; p.seed = seed; p.seed_dev = reinterpret_cast<const unsigned long long*>(seed_dev);
    const int smem = bwd_smem_bytes(S);
    static SmemAttrCache attr;
    rc = ensure_smem_attr(attn_bwd_kernel, attr, smem, false);
    if (rc != DLE_OK) return rc;
    attn_bwd_kernel<<<dim3(A, B), BWD_THREADS, smem, st>>>(tq, td, p);
    DLE_LAUNCH_CHECK();
    return DLE_OK;
}

#ifdef DLE_ATTN_TRACE
// measurement-only (trace variant): copies the device trace buffers to host memory
extern "C" int dle_debug_attn_trace(unsigned long long* events /*[4][4096]*/, int* counts /*[4]*/, unsigned long long* ctas /*[8192][3]*/) {
    if (cudaDeviceSynchronize() != cudaSuccess) return DLE_ERR_CUDA;
    if (cudaMemcpyFromSymbol(events, dle::g_attn_trace, sizeof(unsigned long long) * 4 * dle::TRACE_CAP) != cudaSuccess) return DLE_ERR_CUDA;
    if (cudaMemcpyFromSymbol(counts, dle::g_attn_trace_n, sizeof(int) * 4) != cudaSuccess) return DLE_ERR_CUDA;
    if (cudaMemcpyFromSymbol(ctas, dle::g_attn_cta, sizeof(unsigned long long) * 3 * dle::TRACE_MAX_CTAS) != cudaSuccess) return DLE_ERR_CUDA;
    return DLE_OK;
}
#endif
