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
//   warps : 0 = TMA producer (+TMEM alloc), 1 = MMA issuer, 2..9 = softmax.  TWO threads per query row: the two warps of a TMEM lane
//           quarter each own 64 of the 128 key columns of a chunk (v2 had one thread per row = 2 softmax warps per scheduler with both
//           CTAs resident and was dependency-bound: ncu 44 % issue-active).  The halves agree on the chunk's row maximum through a
//           512 B shared exchange (bf16, rounded UP so that exp2(s - m) <= 1 still holds) and a 64-thread named barrier.
//   lazy rescale : the running maximum only moves when a chunk exceeds it by more than 2^8 (P stays <= 256, exact in the final
//           normalisation because l is accumulated against the same reference), so the O read-modify-write through TMEM is rare.
//   math  : packed FFMA2/FADD2 for scale-subtract and the row sums; dropout keeps arrive as bf16x2 AND-masks (common.cuh).
// =================================================================================================
constexpr int FWD_SOFTMAX_WARPS = 8;
constexpr int FWD_THREADS = (2 + FWD_SOFTMAX_WARPS) * 32;      // 320
constexpr int FWD_TMEM_COLS = 256;
constexpr int FWD_XCHG_BYTES = 2 * TQ * 2;                     // [half][row] bf16
constexpr int FWD_BAR_BYTES = 128;
constexpr int FWD_SMEM_BYTES = TILE_BYTES /*Q*/ + 4 * TILE_BYTES /*K,V rings*/ + PT_BYTES /*P*/ + FWD_XCHG_BYTES + FWD_BAR_BYTES;
constexpr float FWD_RESCALE_THRESH = 8.0f;                     // log2 domain

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

__global__ void __launch_bounds__(FWD_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnFwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int S = p.S, n_chunks = S / TQ;
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
    uint64_t* s_full = bars + 9;         // 1
    uint64_t* p_full = bars + 10;        // 1 (one arrival per softmax warp)
    uint64_t* pv_done = bars + 11;       // 1
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) __trap();          // SWIZZLE_128B tiles need a 1024-byte aligned base
        tma_prefetch_desc(&tmap_qkv);
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
        mbar_init(s_full, 1); mbar_init(p_full, FWD_SOFTMAX_WARPS); mbar_init(pv_done, 1);
        fence_barrier_init();
    }
    if (warp == 0) { tmem_alloc(tmem_ptr, FWD_TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + TQ;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_expect_tx(q_full, TILE_BYTES);
            tma_load_3d(sQ, &tmap_qkv, q_full, h * HD, qt * TQ, b);
            for (int j = 0; j < n_chunks; ++j) {
                const int st = j & 1; const uint32_t ph = (j >> 1) & 1;
                mbar_wait(&k_empty[st], ph ^ 1);
                mbar_expect_tx(&k_full[st], TILE_BYTES);
                tma_load_3d(sK + st * TILE_BYTES, &tmap_qkv, &k_full[st], p.H + h * HD, j * TQ, b);
                mbar_wait(&v_empty[st], ph ^ 1);
                mbar_expect_tx(&v_full[st], TILE_BYTES);
                tma_load_3d(sV + st * TILE_BYTES, &tmap_qkv, &v_full[st], 2 * p.H + h * HD, j * TQ, b);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16(TQ, TQ, false, false);
            constexpr uint32_t idesc_pv = make_idesc_bf16(TQ, HD, false, true);
            const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);
            auto issue_s = [&](int j_) {                       // S_j = Q K_j^T
                const int st = j_ & 1;
                mbar_wait(&k_full[st], (j_ >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16_ss(tmem_S, make_smem_desc_sw128(aQ + kk * 32, 0, 1024),
                                 make_smem_desc_sw128(aK + st * TILE_BYTES + kk * 32, 0, 1024), idesc_s, kk > 0 ? 1u : 0u);
                umma_commit(&k_empty[st]);
                umma_commit(s_full);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < n_chunks; ++j) {
                const int st = j & 1; const uint32_t ph = (j >> 1) & 1;
                // p_full(j): the softmax warps have read S_j completely and written P_j.  S_{j+1} goes FIRST (its columns are free), so the
                // next chunk's softmax starts while PV_j executes.
                mbar_wait(p_full, j & 1);
                tc_fence_after();
                if (j + 1 < n_chunks) issue_s(j + 1);
                mbar_wait(&v_full[st], ph);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)                // O += P_j V_j
                    umma_bf16_ss(tmem_O, make_smem_desc_sw128(aP + (kk >> 2) * TILE_BYTES + (kk & 3) * 32, 0, 1024),
                                 make_smem_desc_sw128(aV + st * TILE_BYTES + kk * 2048, TILE_BYTES, 1024), idesc_pv,
                                 (j > 0 || kk > 0) ? 1u : 0u);
                umma_commit(&v_empty[st]);
                umma_commit(pv_done);
            }
        }
    } else {
        // ===================== softmax warps: two threads per query row, 64 key columns each =====================
        const int q4 = warp & 3;                      // TMEM lane quarter this warp may access (hardware rule: warp id % 4)
        const int hf = (warp - 2) >> 2;               // which 64-column half of the chunk
        const int r = q4 * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
        const uint32_t tS = tmem_S + lane_addr + hf * 64, tO = tmem_O + lane_addr + hf * 32;
        const uint32_t sp = smem_u32(sP);
        const unsigned long long seed = effective_seed(p.seed, p.seed_dev);
        float m_run = -INFINITY, l0 = 0.f;
        const unsigned long long drop_row = ((unsigned long long)(b * p.A + h) * S + (qt * TQ + r)) * (unsigned long long)S + hf * 64;
        const float* mrow = p.mask ? p.mask + (long long)b * S + hf * 64 : nullptr;
        // which 128-key chunks carry a non-zero additive mask in MY half (warp-uniform bit set; typical batches: none or the tail)
        uint32_t chunk_masked = 0;
        if (mrow) {
            for (int j = 0; j < n_chunks; ++j) {
                const float2 m2 = __ldg(reinterpret_cast<const float2*>(mrow + j * TQ) + lane);
                if (__any_sync(0xffffffffu, m2.x != 0.f || m2.y != 0.f)) chunk_masked |= 1u << j;
            }
        }
        for (int j = 0; j < n_chunks; ++j) {
            const bool masked = (chunk_masked >> j) & 1u;
            const float* mk = mrow + j * TQ;
            mbar_wait(s_full, j & 1);
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
            tc_fence_before();
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
        const long long tok = (long long)b * p.tok_stride_b + (long long)(qt * TQ + r) * p.tok_stride_s;
        bf16* o = p.ctx + tok * p.H + h * HD + hf * 32;
        {
            uint32_t v[32];
            tmem_ld32(tO, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 8)
                st_global_v4(o + i, pack_bf16(__uint_as_float(v[i]) * inv_l, __uint_as_float(v[i + 1]) * inv_l),
                             pack_bf16(__uint_as_float(v[i + 2]) * inv_l, __uint_as_float(v[i + 3]) * inv_l),
                             pack_bf16(__uint_as_float(v[i + 4]) * inv_l, __uint_as_float(v[i + 5]) * inv_l),
                             pack_bf16(__uint_as_float(v[i + 6]) * inv_l, __uint_as_float(v[i + 7]) * inv_l));
        }
        if (hf == 0) p.lse[((long long)b * p.A + h) * S + qt * TQ + r] = (m_run + log2f(l_tot)) * LN2;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, FWD_TMEM_COLS); }
}

// =================================================================================================
// backward
// =================================================================================================
constexpr int BWD_THREADS = 576;       // warp 0: TMA + TMEM alloc, warp 1: MMA, warps 2..17: compute
constexpr int BWD_COMPUTE_THREADS = 512;   // 4 warps per TMEM lane quarter, each owning 32 of the 128 tile columns

struct AttnBwdParams {
    const float* mask; const float* lse; const float* delta;
    float* dq_acc;         // [B, A, S, 64] fp32 scratch: dQ accumulated over the kv tiles (unused when S == 128)
    float* dbias;          // [3H] fp32 or null: += column sums of dqkv
    bf16* dqkv;            // [T, 3H]
    int B, S, A, H;
    int tok_stride_s, tok_stride_b;
    float scale, scale_log2;
    uint32_t drop_k2; uint32_t drop_on; float drop_scale; uint32_t drop_stream; unsigned long long seed; const unsigned long long* seed_dev;
};

__host__ __device__ inline int bwd_smem_bytes(int S) {
    return 1024 + 2 * TILE_BYTES /*K,V*/ + 4 * TILE_BYTES /*Q,dO x2*/ + 2 * PT_BYTES /*P,dS*/ + S * 4 + 256;
}

// delta[b,h,s] = sum_d dO[t, h*64+d] * O[t, h*64+d]
__global__ void attn_delta_kernel(const bf16* __restrict__ dctx, const bf16* __restrict__ ctx, float* __restrict__ delta,
                                  int B, int S, int A, int seq_first) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // token * A + h
    const long long total = (long long)B * S * A;
    if (idx >= total) return;
    const long long tok = idx / A; const int h = (int)(idx - tok * A);
    const bf16* a = dctx + tok * (long long)(A * HD) + h * HD;
    const bf16* c = ctx + tok * (long long)(A * HD) + h * HD;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < HD; i += 8) {
        uint4 ua = ld_global_nc_v4(a + i), uc = ld_global_nc_v4(c + i);
        float2 x, y;
        x = unpack_bf16(ua.x); y = unpack_bf16(uc.x); acc += x.x * y.x + x.y * y.y;
        x = unpack_bf16(ua.y); y = unpack_bf16(uc.y); acc += x.x * y.x + x.y * y.y;
        x = unpack_bf16(ua.z); y = unpack_bf16(uc.z); acc += x.x * y.x + x.y * y.y;
        x = unpack_bf16(ua.w); y = unpack_bf16(uc.w); acc += x.x * y.x + x.y * y.y;
    }
    const int b = seq_first ? (int)(tok % B) : (int)(tok / S);
    const int s = seq_first ? (int)(tok / B) : (int)(tok - (long long)b * S);
    delta[((long long)b * A + h) * S + s] = acc;
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do, const AttnBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int S = p.S, n = S / TQ;
    uint8_t* sK = smem;
    uint8_t* sV = sK + TILE_BYTES;
    uint8_t* sQ = sV + TILE_BYTES;               // [2]
    uint8_t* sdO = sQ + 2 * TILE_BYTES;          // [2]
    uint8_t* sP = sdO + 2 * TILE_BYTES;
    uint8_t* sdS = sP + PT_BYTES;
    float* sMask = reinterpret_cast<float*>(sdS + PT_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sMask + S);
    uint64_t* qdo_full = bars;        // [2]
    uint64_t* qdo_empty = bars + 2;   // [2]
    uint64_t* kv_full = bars + 4;
    uint64_t* kv_empty = bars + 5;
    uint64_t* s_full = bars + 6;
    uint64_t* p_full = bars + 7;      // count 256
    uint64_t* dp_full = bars + 8;
    uint64_t* ds_full = bars + 9;     // count 256
    uint64_t* pair_done = bars + 10;
    uint64_t* dkv_full = bars + 11;
    uint64_t* dkv_read = bars + 12;   // one arrival per compute warp
    uint64_t* dv_done = bars + 13;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 14);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.x, b = blockIdx.y;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_qkv); tma_prefetch_desc(&tmap_do);
        for (int i = 0; i < 2; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
        mbar_init(kv_full, 1); mbar_init(kv_empty, 1); mbar_init(s_full, 1); mbar_init(p_full, BWD_COMPUTE_THREADS / 32);
        mbar_init(dp_full, 1); mbar_init(ds_full, BWD_COMPUTE_THREADS / 32); mbar_init(pair_done, 1); mbar_init(dkv_full, 1);
        mbar_init(dkv_read, BWD_COMPUTE_THREADS / 32); mbar_init(dv_done, 1);
        fence_barrier_init();
    }
    if (warp == 0) { tmem_alloc(tmem_ptr, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // S | dP | dV | dK | dQ(current pair only): S and dP no longer share columns, so S(t+1) is issued while dS(t) is being computed and
    // dP(t+1) while P(t+1) is -- every MMA hides behind a compute phase.  dQ of the pair is drained to an fp32 global scratch block by the
    // compute warps (same thread, same elements, fixed j order: deterministic) and converted to bf16 with the last kv tile.
    const uint32_t tmem_S = tmem_base, tmem_dP = tmem_base + 128, tmem_dV = tmem_base + 256, tmem_dK = tmem_base + 320, tmem_dQ = tmem_base + 384;

    if (warp == 0) {
        if (lane == 0) {
            for (int j = 0; j < n; ++j) {
                mbar_wait(kv_empty, (j & 1) ^ 1);
                mbar_expect_tx(kv_full, 2 * TILE_BYTES);
                tma_load_3d(sK, &tmap_qkv, kv_full, p.H + h * HD, j * TQ, b);
                tma_load_3d(sV, &tmap_qkv, kv_full, 2 * p.H + h * HD, j * TQ, b);
                for (int i = 0; i < n; ++i) {
                    const int t = j * n + i, st = t & 1;
                    mbar_wait(&qdo_empty[st], ((t >> 1) & 1) ^ 1);
                    mbar_expect_tx(&qdo_full[st], 2 * TILE_BYTES);
                    tma_load_3d(sQ + st * TILE_BYTES, &tmap_qkv, &qdo_full[st], h * HD, i * TQ, b);
                    tma_load_3d(sdO + st * TILE_BYTES, &tmap_do, &qdo_full[st], h * HD, i * TQ, b);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t id_kk = make_idesc_bf16(TQ, TQ, false, false);     // S, dP
            constexpr uint32_t id_mm = make_idesc_bf16(TQ, HD, true, true);       // dV, dK
            constexpr uint32_t id_km = make_idesc_bf16(TQ, HD, false, true);      // dQ
            const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP), adS = smem_u32(sdS);
            // Issue order per pair t = j*n + i (S and dP own separate TMEM columns):
            //   p_full(t):  S(t+1) -> dV(t)            (S columns are free once P(t) was computed from them)
            //   ds_full(t): dP(t+1) -> dK(t), dQ(t)    (dP columns are free once dS(t) was computed; dQ(t-1) was drained before ds_full(t))
            // so S(t+1) executes while the compute warps are in the dS(t) phase and dP(t+1) while they are in the P(t+1) phase.
            auto issue_s = [&](int t_) {
                const uint32_t aQ_ = smem_u32(sQ) + (t_ & 1) * TILE_BYTES;
                mbar_wait(&qdo_full[t_ & 1], (t_ >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16_ss(tmem_S, make_smem_desc_sw128(aQ_ + kk * 32, 0, 1024), make_smem_desc_sw128(aK + kk * 32, 0, 1024),
                                 id_kk, kk > 0 ? 1u : 0u);
                umma_commit(s_full);
            };
            auto issue_dp = [&](int t_) {                       // qdo_full[t_ & 1] was already observed by issue_s(t_)
                const uint32_t adO_ = smem_u32(sdO) + (t_ & 1) * TILE_BYTES;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16_ss(tmem_dP, make_smem_desc_sw128(adO_ + kk * 32, 0, 1024), make_smem_desc_sw128(aV + kk * 32, 0, 1024),
                                 id_kk, kk > 0 ? 1u : 0u);
                umma_commit(dp_full);
            };
            for (int j = 0; j < n; ++j) {
                mbar_wait(kv_full, j & 1);
                if (j >= 1) mbar_wait(dkv_read, (j - 1) & 1);       // dV/dK accumulators (and the last dQ of the previous tile) drained
                tc_fence_after();
                issue_s(j * n);                                      // first pair of this kv tile: both inputs of the compute warps up front
                issue_dp(j * n);
                for (int i = 0; i < n; ++i) {
                    const int t = j * n + i, st = t & 1;
                    const uint32_t aQ = smem_u32(sQ) + st * TILE_BYTES, adO = smem_u32(sdO) + st * TILE_BYTES;
                    mbar_wait(p_full, t & 1);                        // P~(t) in smem, S(t) consumed
                    tc_fence_after();
                    if (i + 1 < n) issue_s(t + 1);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)                  // dV_j += P~^T dO_i
                        umma_bf16_ss(tmem_dV, make_smem_desc_sw128(aP + kk * 2048, TILE_BYTES, 1024),
                                     make_smem_desc_sw128(adO + kk * 2048, TILE_BYTES, 1024), id_mm, (i > 0 || kk > 0) ? 1u : 0u);
                    umma_commit(dv_done);                            // sP may be overwritten
                    mbar_wait(ds_full, t & 1);                       // dS(t) in smem, dP(t) consumed, dQ(t-1) drained
                    tc_fence_after();
                    if (i + 1 < n) issue_dp(t + 1);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)                  // dK_j += dS^T Q_i
                        umma_bf16_ss(tmem_dK, make_smem_desc_sw128(adS + kk * 2048, TILE_BYTES, 1024),
                                     make_smem_desc_sw128(aQ + kk * 2048, TILE_BYTES, 1024), id_mm, (i > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)                  // dQ_ij = dS K_j (this pair only; accumulated over j in global memory)
                        umma_bf16_ss(tmem_dQ, make_smem_desc_sw128(adS + (kk >> 2) * TILE_BYTES + (kk & 3) * 32, 0, 1024),
                                     make_smem_desc_sw128(aK + kk * 2048, TILE_BYTES, 1024), id_km, kk > 0 ? 1u : 0u);
                    umma_commit(&qdo_empty[st]);
                    umma_commit(pair_done);
                }
                umma_commit(kv_empty);
                umma_commit(dkv_full);
            }
        }
    } else {
        // ===================== compute warps =====================
        const int q4 = warp & 3, qc = (warp - 2) >> 2;          // qc: which 32-column quarter of the tile
        const int r = q4 * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
        const int ct = threadIdx.x - 64;
        for (int i = ct; i < S; i += BWD_COMPUTE_THREADS) sMask[i] = p.mask ? p.mask[(long long)b * S + i] * LOG2E : 0.f;
        named_bar_sync(1, BWD_COMPUTE_THREADS);
        const long long bh = (long long)b * p.A + h;
        const uint32_t aP = smem_u32(sP), adS = smem_u32(sdS);
        const unsigned long long seed = effective_seed(p.seed, p.seed_dev);
        // dQ of pair (j, i): 128 x 64 fp32 in TMEM; this thread owns row r, columns [qc*16, +16) -- the same thread touches the same 16
        // addresses for every j, in program order, so the accumulation order is fixed.  j == 0 stores, 0 < j < n-1 adds (red), the last
        // kv tile reads the running sum back, adds its own part and writes bf16 (+ the query-bias column sums).
        auto drain_dq = [&](int j_, int i_) {
            uint32_t w[16];
            tmem_ld16(tmem_dQ + lane_addr + qc * 16, w);
            tmem_ld_wait();
            float* acc = p.dq_acc + ((bh * S + i_ * TQ + r) * HD + qc * 16);
            if (j_ + 1 < n) {
                if (j_ == 0) {
#pragma unroll
                    for (int k = 0; k < 16; k += 4)
                        *reinterpret_cast<float4*>(acc + k) = make_float4(__uint_as_float(w[k]), __uint_as_float(w[k + 1]), __uint_as_float(w[k + 2]), __uint_as_float(w[k + 3]));
                } else {
#pragma unroll
                    for (int k = 0; k < 16; k += 4)
                        red_add_v4_f32(acc + k, __uint_as_float(w[k]), __uint_as_float(w[k + 1]), __uint_as_float(w[k + 2]), __uint_as_float(w[k + 3]));
                }
                return;
            }
            float f[32];
#pragma unroll
            for (int k = 0; k < 16; ++k) f[k] = __uint_as_float(w[k]);
            if (n > 1) {
#pragma unroll
                for (int k = 0; k < 16; k += 4) {
                    const float4 prev = __ldcg(reinterpret_cast<const float4*>(acc + k));
                    f[k] += prev.x; f[k + 1] += prev.y; f[k + 2] += prev.z; f[k + 3] += prev.w;
                }
            }
            const long long tok_q = (long long)b * p.tok_stride_b + (long long)(i_ * TQ + r) * p.tok_stride_s;
            bf16* o = p.dqkv + tok_q * (3LL * p.H) + h * HD + qc * 16;
#pragma unroll
            for (int k = 0; k < 16; k += 8)
                st_global_v4(o + k, pack_bf16(f[k], f[k + 1]), pack_bf16(f[k + 2], f[k + 3]), pack_bf16(f[k + 4], f[k + 5]), pack_bf16(f[k + 6], f[k + 7]));
            if (p.dbias != nullptr) {                            // query bias gradient: column sums of the stored bf16 values
#pragma unroll
                for (int k = 0; k < 16; ++k) f[k] = __bfloat162float(__float2bfloat16_rn(f[k]));
#pragma unroll
                for (int k = 16; k < 32; ++k) f[k] = 0.f;
                const float cs = warp_column_sums32(f, lane);
                if (lane < 16) atomicAdd(p.dbias + h * HD + qc * 16 + lane, cs);
            }
        };
        for (int j = 0; j < n; ++j) {
            // my 32 key columns of this kv tile: additive mask (already x log2e) from shared memory, skipped entirely when it is all
            // zero (warp-uniform; unpadded batches).  Explicit ld.shared: the generic loads the compiler emitted for sMask[] went
            // through the global/local queue (ncu: 4 % of the kernel's samples on `lg` throttle at those eight loads).
            const uint32_t mk_addr = smem_u32(sMask + j * TQ + qc * 32);
            const bool masked = __any_sync(0xffffffffu, sMask[j * TQ + qc * 32 + lane] != 0.f);
            for (int i = 0; i < n; ++i) {
                const int t = j * n + i;
                const float lse2 = p.lse[bh * S + i * TQ + r] * LOG2E;
                const float dl = p.delta[bh * S + i * TQ + r];
                const unsigned long long drop_row = (unsigned long long)(bh * S + (i * TQ + r)) * (unsigned long long)S + j * TQ + qc * 32;
                uint32_t pk[16];                 // undropped P, packed bf16x2 (32 values)
                uint32_t km[16];                 // keep-masks of my 32 columns (bf16x2 AND-masks)
                mbar_wait(s_full, t & 1);
                tc_fence_after();
                if (t >= 1) { mbar_wait(dv_done, (t - 1) & 1); tc_fence_after(); }     // dV(t-1) retired: sP may be overwritten
                uint32_t v[32];
                {
                    tmem_ld32(tmem_S + lane_addr + qc * 32, v);
                    tmem_ld_wait();
                    float e[32];
                    const float neg_lse2 = -lse2;
                    if (masked) {
#pragma unroll
                        for (int g = 0; g < 8; ++g) {
                            float m0, m1, m2, m3;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(m0), "=f"(m1), "=f"(m2), "=f"(m3) : "r"(mk_addr + g * 16));
                            e[4 * g + 0] = ex2(fmaf(__uint_as_float(v[4 * g + 0]), p.scale_log2, m0) + neg_lse2);
                            e[4 * g + 1] = ex2(fmaf(__uint_as_float(v[4 * g + 1]), p.scale_log2, m1) + neg_lse2);
                            e[4 * g + 2] = ex2(fmaf(__uint_as_float(v[4 * g + 2]), p.scale_log2, m2) + neg_lse2);
                            e[4 * g + 3] = ex2(fmaf(__uint_as_float(v[4 * g + 3]), p.scale_log2, m3) + neg_lse2);
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
                    // P~ = keep-mask AND P: the 1/(1-p) factor is folded into the dV drain and into the dS constants below
                    if (p.drop_on != 0u) {
                        attn_dropout_masks16(seed, p.drop_stream, drop_row >> 5, p.drop_k2, km);
#pragma unroll
                        for (int g = 0; g < 4; ++g)
                            st_shared_v4(aP + pt_offset(r, qc * 32 + g * 8), pk[g * 4] & km[g * 4], pk[g * 4 + 1] & km[g * 4 + 1],
                                         pk[g * 4 + 2] & km[g * 4 + 2], pk[g * 4 + 3] & km[g * 4 + 3]);
                    } else {
#pragma unroll
                        for (int g = 0; g < 4; ++g)
                            st_shared_v4(aP + pt_offset(r, qc * 32 + g * 8), pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
                    }
                }
                tc_fence_before();
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full);
                // ---- dK/dQ(t-1) retired: sdS may be overwritten, and dQ(t-1) (previous query tile, same kv tile) is ready to be drained.
                //      This sits between the two phases so that it overlaps with dP(t) / S(t+1) on the tensor core.
                if (t >= 1) { mbar_wait(pair_done, (t - 1) & 1); tc_fence_after(); }
                if (i >= 1) drain_dq(j, i - 1);
                // ---- dS = [ (mask & P) * dP / (1-p) - P * delta ] * scale
                mbar_wait(dp_full, t & 1);
                tc_fence_after();
                {
                    tmem_ld32(tmem_dP + lane_addr + qc * 32, v);
                    tmem_ld_wait();
                    const float c1 = p.drop_scale * p.scale, nd = -dl * p.scale;
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
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        st_shared_v4(adS + pt_offset(r, qc * 32 + g * 8), ds[g * 4], ds[g * 4 + 1], ds[g * 4 + 2], ds[g * 4 + 3]);
                }
                tc_fence_before();
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(ds_full);
            }
            // ---- dV_j, dK_j complete (and with them the last pair's dQ): drain to global
            mbar_wait(dkv_full, j & 1);
            tc_fence_after();
            drain_dq(j, n - 1);
            const long long tok = (long long)b * p.tok_stride_b + (long long)(j * TQ + r) * p.tok_stride_s;
            uint32_t v[32];
            if (qc < 2) {                                        // 64-wide accumulators: two of the four column-warps drain them
            const int hf = qc;
#pragma unroll
            for (int which = 0; which < 2; ++which) {           // 0: dK (col block 1), 1: dV (col block 2)
                tmem_ld32((which == 0 ? tmem_dK : tmem_dV) + lane_addr + hf * 32, v);
                tmem_ld_wait();
                if (which == 1 && p.drop_on != 0u) {             // dV accumulated keep-mask AND P: apply the 1/(1-p) factor here
#pragma unroll
                    for (int k = 0; k < 32; ++k) v[k] = __float_as_uint(__uint_as_float(v[k]) * p.drop_scale);
                }
                bf16* o = p.dqkv + tok * (3LL * p.H) + (which + 1) * p.H + h * HD + hf * 32;
#pragma unroll
                for (int k = 0; k < 32; k += 8)
                    st_global_v4(o + k, pack_bf16(__uint_as_float(v[k]), __uint_as_float(v[k + 1])), pack_bf16(__uint_as_float(v[k + 2]), __uint_as_float(v[k + 3])),
                                 pack_bf16(__uint_as_float(v[k + 4]), __uint_as_float(v[k + 5])), pack_bf16(__uint_as_float(v[k + 6]), __uint_as_float(v[k + 7])));
                if (p.dbias != nullptr) {                        // key / value bias gradients: column sums of the stored bf16 values
                    float f[32];
#pragma unroll
                    for (int k = 0; k < 32; ++k) f[k] = __bfloat162float(__float2bfloat16_rn(__uint_as_float(v[k])));
                    const float cs = warp_column_sums32(f, lane);
                    atomicAdd(p.dbias + (which + 1) * p.H + h * HD + hf * 32 + lane, cs);
                }
            }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(dkv_read);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
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
    attn_delta_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(reinterpret_cast<const bf16*>(dctx), reinterpret_cast<const bf16*>(ctx), delta_ws, B, S, A, seq_first);
    DLE_LAUNCH_CHECK();
    AttnBwdParams p;
    p.mask = mask; p.lse = lse; p.delta = delta_ws; p.dq_acc = delta_ws + (long long)B * A * S; p.dbias = dbias_qkv; p.dqkv = reinterpret_cast<bf16*>(dqkv);
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
