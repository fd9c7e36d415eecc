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
// forward  (v2: sized for TWO co-resident CTAs per SM so one CTA's softmax overlaps the other's MMAs)
//   smem  : Q 16 KB + K ring 2x16 KB + V ring 2x16 KB + P 32 KB = 112 KB (+ barriers)      -> 2 CTAs in 228 KB
//   TMEM  : S 128 cols + O 64 cols -> 256-column allocation                                -> 2 CTAs in 512 cols
//   warps : 0 = TMA producer (+TMEM alloc), 1 = MMA issuer, 2..5 = softmax, ONE THREAD PER QUERY ROW
//           (row max / sum are thread-local: no cross-warp exchange, no named barrier)
// =================================================================================================
constexpr int FWD_THREADS = 192;
constexpr int FWD_SOFTMAX_THREADS = 128;
constexpr int FWD_TMEM_COLS = 256;
constexpr int FWD_SMEM_BYTES = TILE_BYTES /*Q*/ + 4 * TILE_BYTES /*K,V rings*/ + PT_BYTES /*P*/ + 256 /*barriers*/;

struct AttnFwdParams {
    const float* mask;     // [B,S] additive or null
    bf16* ctx;             // [T, H]
    float* lse;            // [B,A,S]
    int B, S, A, H;
    int tok_stride_s, tok_stride_b;   // token row of (b, s) = b*tok_stride_b + s*tok_stride_s
    float scale_log2;      // (1/sqrt(d)) * log2(e)
    uint32_t drop_thresh; float drop_scale; uint32_t drop_stream; unsigned long long seed;
};

__global__ void __launch_bounds__(FWD_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnFwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int S = p.S, n_chunks = S / TQ;
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + TILE_BYTES;               // [2]
    uint8_t* sV = sK + 2 * TILE_BYTES;           // [2]
    uint8_t* sP = sV + 2 * TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + PT_BYTES);
    uint64_t* q_full = bars;             // 1
    uint64_t* k_full = bars + 1;         // [2]
    uint64_t* k_empty = bars + 3;        // [2]
    uint64_t* v_full = bars + 5;         // [2]
    uint64_t* v_empty = bars + 7;        // [2]
    uint64_t* s_full = bars + 9;         // 1
    uint64_t* p_full = bars + 10;        // 1 (count 128)
    uint64_t* pv_done = bars + 11;       // 1
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) __trap();          // SWIZZLE_128B tiles need a 1024-byte aligned base
        tma_prefetch_desc(&tmap_qkv);
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
        mbar_init(s_full, 1); mbar_init(p_full, FWD_SOFTMAX_THREADS / 32); mbar_init(pv_done, 1);
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
            mbar_wait(q_full, 0);
            for (int j = 0; j < n_chunks; ++j) {
                const int st = j & 1; const uint32_t ph = (j >> 1) & 1;
                // S_j = Q K_j^T.  The S columns are free: the softmax warps finished reading S_{j-1} before p_full(j-1),
                // which this thread waited for before issuing PV_{j-1}.
                mbar_wait(&k_full[st], ph);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16_ss(tmem_S, make_smem_desc_sw128(aQ + kk * 32, 0, 1024),
                                 make_smem_desc_sw128(aK + st * TILE_BYTES + kk * 32, 0, 1024), idesc_s, kk > 0 ? 1u : 0u);
                umma_commit(&k_empty[st]);
                umma_commit(s_full);
                // O += P_j V_j
                mbar_wait(p_full, j & 1);
                tc_fence_after();
                mbar_wait(&v_full[st], ph);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                    umma_bf16_ss(tmem_O, make_smem_desc_sw128(aP + (kk >> 2) * TILE_BYTES + (kk & 3) * 32, 0, 1024),
                                 make_smem_desc_sw128(aV + st * TILE_BYTES + kk * 2048, TILE_BYTES, 1024), idesc_pv,
                                 (j > 0 || kk > 0) ? 1u : 0u);
                umma_commit(&v_empty[st]);
                umma_commit(pv_done);
            }
        }
    } else {
        // ===================== softmax warps: thread = query row =====================
        const int q4 = warp & 3;
        const int r = q4 * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
        const uint32_t sp = smem_u32(sP);
        float m_run = -INFINITY, l_run = 0.f;
        const unsigned long long drop_row = ((unsigned long long)(b * p.A + h) * S + (qt * TQ + r)) * (unsigned long long)S;
        const float* mrow = p.mask ? p.mask + (long long)b * S : nullptr;
        // which 128-key chunks carry a non-zero additive mask (warp-uniform bit set; typical batches: none or the tail)
        uint32_t chunk_masked = 0;
        if (mrow) {
            for (int j = 0; j < n_chunks; ++j) {
                const float4 m4 = __ldg(reinterpret_cast<const float4*>(mrow + j * TQ) + lane);
                if (__any_sync(0xffffffffu, m4.x != 0.f || m4.y != 0.f || m4.z != 0.f || m4.w != 0.f)) chunk_masked |= 1u << j;
            }
        }
        for (int j = 0; j < n_chunks; ++j) {
            const bool masked = (chunk_masked >> j) & 1u;
            const float* mk = mrow + j * TQ;
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            uint32_t v[2][32];
            // ---- pass 1: row max over the 128 keys of this chunk (next TMEM piece in flight while this one is reduced)
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
            tmem_ld32(tmem_S + lane_addr, v[0]);
#pragma unroll
            for (int pc = 0; pc < 4; ++pc) {
                tmem_ld_wait();
                if (pc < 3) tmem_ld32(tmem_S + lane_addr + (pc + 1) * 32, v[(pc + 1) & 1]);
                const uint32_t(&w)[32] = v[pc & 1];
                if (masked) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 m4 = __ldg(reinterpret_cast<const float4*>(mk + pc * 32) + i);
                        mx0 = fmaxf(mx0, fmaf(__uint_as_float(w[4 * i]), p.scale_log2, m4.x * LOG2E));
                        mx1 = fmaxf(mx1, fmaf(__uint_as_float(w[4 * i + 1]), p.scale_log2, m4.y * LOG2E));
                        mx2 = fmaxf(mx2, fmaf(__uint_as_float(w[4 * i + 2]), p.scale_log2, m4.z * LOG2E));
                        mx3 = fmaxf(mx3, fmaf(__uint_as_float(w[4 * i + 3]), p.scale_log2, m4.w * LOG2E));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        mx0 = fmaxf(mx0, __uint_as_float(w[i])); mx1 = fmaxf(mx1, __uint_as_float(w[i + 1]));
                        mx2 = fmaxf(mx2, __uint_as_float(w[i + 2])); mx3 = fmaxf(mx3, __uint_as_float(w[i + 3]));
                    }
                }
            }
            float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            if (!masked) mx *= p.scale_log2;                          // scale > 0: max commutes with the scaling
            const float m_new = fmaxf(m_run, mx);
            const float alpha = ex2(m_run - m_new);                    // 0 at j == 0
            tmem_ld32(tmem_S + lane_addr, v[0]);                       // first piece of pass 2 in flight during the wait below
            if (j >= 1) {                                              // PV_{j-1} retired: O is valid, the P buffer is free
                mbar_wait(pv_done, (j - 1) & 1);
                tc_fence_after();
            }
            // ---- pass 2: p = exp2(s - m), row sum (before dropout), dropout, bf16 -> swizzled smem
            float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
#pragma unroll
            for (int pc = 0; pc < 4; ++pc) {
                tmem_ld_wait();
                if (pc < 3) tmem_ld32(tmem_S + lane_addr + (pc + 1) * 32, v[(pc + 1) & 1]);
                const uint32_t(&w)[32] = v[pc & 1];
                float e[32];
                if (masked) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 m4 = __ldg(reinterpret_cast<const float4*>(mk + pc * 32) + i);
                        e[4 * i] = ex2(fmaf(__uint_as_float(w[4 * i]), p.scale_log2, m4.x * LOG2E) - m_new);
                        e[4 * i + 1] = ex2(fmaf(__uint_as_float(w[4 * i + 1]), p.scale_log2, m4.y * LOG2E) - m_new);
                        e[4 * i + 2] = ex2(fmaf(__uint_as_float(w[4 * i + 2]), p.scale_log2, m4.z * LOG2E) - m_new);
                        e[4 * i + 3] = ex2(fmaf(__uint_as_float(w[4 * i + 3]), p.scale_log2, m4.w * LOG2E) - m_new);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) e[i] = ex2(fmaf(__uint_as_float(w[i]), p.scale_log2, -m_new));
                }
#pragma unroll
                for (int i = 0; i < 32; i += 4) { rs0 += e[i]; rs1 += e[i + 1]; rs2 += e[i + 2]; rs3 += e[i + 3]; }
                const int col = pc * 32;
                if (p.drop_thresh != 0u) {                             // the 1/(1-p) scale is applied once, in the epilogue
                    const uint32_t keep = dropout_keep32(p.seed, p.drop_stream, (drop_row + j * TQ + col) >> 5, p.drop_thresh);
#pragma unroll
                    for (int i = 0; i < 32; ++i) e[i] = ((keep >> i) & 1u) ? e[i] : 0.f;
                }
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    st_shared_v4(sp + pt_offset(r, col + g * 8), pack_bf16(e[g * 8], e[g * 8 + 1]), pack_bf16(e[g * 8 + 2], e[g * 8 + 3]),
                                 pack_bf16(e[g * 8 + 4], e[g * 8 + 5]), pack_bf16(e[g * 8 + 6], e[g * 8 + 7]));
            }
            if (j >= 1 && __any_sync(0xffffffffu, alpha != 1.0f)) {    // rescale this row of O (64 columns)
                tmem_ld32(tmem_O + lane_addr, v[0]);
                tmem_ld32(tmem_O + lane_addr + 32, v[1]);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) { v[0][i] = __float_as_uint(__uint_as_float(v[0][i]) * alpha); v[1][i] = __float_as_uint(__uint_as_float(v[1][i]) * alpha); }
                tmem_st32(tmem_O + lane_addr, v[0]);
                tmem_st32(tmem_O + lane_addr + 32, v[1]);
                tmem_st_wait();
            }
            l_run = l_run * alpha + ((rs0 + rs1) + (rs2 + rs3));
            m_run = m_new;
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);                       // one arrival per warp (count = 4)
        }
        // ---- epilogue: O * (1/(1-p)) / l -> ctx, lse
        mbar_wait(pv_done, (n_chunks - 1) & 1);
        tc_fence_after();
        const float inv_l = p.drop_scale / l_run;
        const long long tok = (long long)b * p.tok_stride_b + (long long)(qt * TQ + r) * p.tok_stride_s;
        bf16* o = p.ctx + tok * p.H + h * HD;
        {
            uint32_t v[2][32];
            tmem_ld32(tmem_O + lane_addr, v[0]);
            tmem_ld32(tmem_O + lane_addr + 32, v[1]);
            tmem_ld_wait();
#pragma unroll
            for (int pc = 0; pc < 2; ++pc)
#pragma unroll
                for (int i = 0; i < 32; i += 8)
                    st_global_v4(o + pc * 32 + i, pack_bf16(__uint_as_float(v[pc][i]) * inv_l, __uint_as_float(v[pc][i + 1]) * inv_l),
                                 pack_bf16(__uint_as_float(v[pc][i + 2]) * inv_l, __uint_as_float(v[pc][i + 3]) * inv_l),
                                 pack_bf16(__uint_as_float(v[pc][i + 4]) * inv_l, __uint_as_float(v[pc][i + 5]) * inv_l),
                                 pack_bf16(__uint_as_float(v[pc][i + 6]) * inv_l, __uint_as_float(v[pc][i + 7]) * inv_l));
        }
        p.lse[((long long)b * p.A + h) * S + qt * TQ + r] = (m_run + log2f(l_run)) * LN2;
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
    float* dbias;          // [3H] fp32 or null: += column sums of dqkv
    bf16* dqkv;            // [T, 3H]
    int B, S, A, H;
    int tok_stride_s, tok_stride_b;
    float scale, scale_log2;
    uint32_t drop_thresh; float drop_scale; uint32_t drop_stream; unsigned long long seed;
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
    const uint32_t tmem_SP = tmem_base, tmem_dV = tmem_base + 128, tmem_dK = tmem_base + 192, tmem_dQ = tmem_base + 256;

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
            // Issue order (per pair t = j*n + i) is chosen so that the compute warps never wait behind MMAs they do not need:
            //   ... ds_full(t-1) -> [S(t)] -> dK(t-1), dQ(t-1) | p_full(t) -> dP(t) -> dV(t) | ds_full(t) -> [S(t+1)] -> dK(t), dQ(t) ...
            // S(t+1) goes first after ds_full(t) (its TMEM columns are free once dP(t) was consumed), so P(t+1) is computed
            // while dK(t)/dQ(t) execute; dP(t) goes before dV(t) so dS(t) is computed while dV(t) executes.
            auto issue_s = [&](int t_, uint32_t aQ_) {
                mbar_wait(&qdo_full[t_ & 1], (t_ >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16_ss(tmem_SP, make_smem_desc_sw128(aQ_ + kk * 32, 0, 1024), make_smem_desc_sw128(aK + kk * 32, 0, 1024),
                                 id_kk, kk > 0 ? 1u : 0u);
                umma_commit(s_full);
            };
            for (int j = 0; j < n; ++j) {
                mbar_wait(kv_full, j & 1);
                if (j >= 1) mbar_wait(dkv_read, (j - 1) & 1);       // dV/dK accumulators drained
                tc_fence_after();
                issue_s(j * n, smem_u32(sQ) + ((j * n) & 1) * TILE_BYTES);      // first pair of this kv tile
                for (int i = 0; i < n; ++i) {
                    const int t = j * n + i, st = t & 1;
                    const uint32_t aQ = smem_u32(sQ) + st * TILE_BYTES, adO = smem_u32(sdO) + st * TILE_BYTES;
                    // dP = dO_i V_j^T (overwrites S once the compute warps have consumed it), then dV_j += P~^T dO_i
                    mbar_wait(p_full, t & 1);
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16_ss(tmem_SP, make_smem_desc_sw128(adO + kk * 32, 0, 1024), make_smem_desc_sw128(aV + kk * 32, 0, 1024),
                                     id_kk, kk > 0 ? 1u : 0u);
                    umma_commit(dp_full);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        umma_bf16_ss(tmem_dV, make_smem_desc_sw128(aP + kk * 2048, TILE_BYTES, 1024),
                                     make_smem_desc_sw128(adO + kk * 2048, TILE_BYTES, 1024), id_mm, (i > 0 || kk > 0) ? 1u : 0u);
                    umma_commit(dv_done);                            // sP may be overwritten
                    // dS ready: next S first (same kv tile only: sK must stay), then dK_j += dS^T Q_i ; dQ_i += dS K_j
                    mbar_wait(ds_full, t & 1);
                    tc_fence_after();
                    if (i + 1 < n) issue_s(t + 1, smem_u32(sQ) + ((t + 1) & 1) * TILE_BYTES);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        umma_bf16_ss(tmem_dK, make_smem_desc_sw128(adS + kk * 2048, TILE_BYTES, 1024),
                                     make_smem_desc_sw128(aQ + kk * 2048, TILE_BYTES, 1024), id_mm, (i > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        umma_bf16_ss(tmem_dQ + i * HD, make_smem_desc_sw128(adS + (kk >> 2) * TILE_BYTES + (kk & 3) * 32, 0, 1024),
                                     make_smem_desc_sw128(aK + kk * 2048, TILE_BYTES, 1024), id_km, (j > 0 || kk > 0) ? 1u : 0u);
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
                uint32_t km = 0xffffffffu;       // keep-mask of my 32 columns
                mbar_wait(s_full, t & 1);
                tc_fence_after();
                if (t >= 1) { mbar_wait(dv_done, (t - 1) & 1); tc_fence_after(); }     // dV(t-1) retired: sP may be overwritten
                uint32_t v[32];
                {
                    tmem_ld32(tmem_SP + lane_addr + qc * 32, v);
                    tmem_ld_wait();
                    float e[32];
                    if (masked) {
#pragma unroll
                        for (int g = 0; g < 8; ++g) {
                            float m0, m1, m2, m3;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(m0), "=f"(m1), "=f"(m2), "=f"(m3) : "r"(mk_addr + g * 16));
                            e[4 * g + 0] = ex2(fmaf(__uint_as_float(v[4 * g + 0]), p.scale_log2, m0) - lse2);
                            e[4 * g + 1] = ex2(fmaf(__uint_as_float(v[4 * g + 1]), p.scale_log2, m1) - lse2);
                            e[4 * g + 2] = ex2(fmaf(__uint_as_float(v[4 * g + 2]), p.scale_log2, m2) - lse2);
                            e[4 * g + 3] = ex2(fmaf(__uint_as_float(v[4 * g + 3]), p.scale_log2, m3) - lse2);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 32; ++k) e[k] = ex2(fmaf(__uint_as_float(v[k]), p.scale_log2, -lse2));
                    }
#pragma unroll
                    for (int k = 0; k < 16; ++k) pk[k] = pack_bf16(e[2 * k], e[2 * k + 1]);
                    if (p.drop_thresh != 0u) {
                        km = dropout_keep32(p.seed, p.drop_stream, drop_row >> 5, p.drop_thresh);
#pragma unroll
                        for (int k = 0; k < 32; ++k) e[k] = ((km >> k) & 1u) ? e[k] * p.drop_scale : 0.f;
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        st_shared_v4(aP + pt_offset(r, qc * 32 + g * 8), pack_bf16(e[g * 8], e[g * 8 + 1]), pack_bf16(e[g * 8 + 2], e[g * 8 + 3]),
                                     pack_bf16(e[g * 8 + 4], e[g * 8 + 5]), pack_bf16(e[g * 8 + 6], e[g * 8 + 7]));
                }
                tc_fence_before();
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full);
                // ---- dS = P * (dP~ - delta) * scale
                mbar_wait(dp_full, t & 1);
                tc_fence_after();
                if (t >= 1) { mbar_wait(pair_done, (t - 1) & 1); tc_fence_after(); }   // dK/dQ(t-1) retired: sdS may be overwritten
                {
                    tmem_ld32(tmem_SP + lane_addr + qc * 32, v);
                    tmem_ld_wait();
                    float e[32];
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const float2 pp = unpack_bf16(pk[k]);
                        float d0 = __uint_as_float(v[2 * k]), d1 = __uint_as_float(v[2 * k + 1]);
                        d0 = ((km >> (2 * k)) & 1u) ? d0 * p.drop_scale : 0.f;
                        d1 = ((km >> (2 * k + 1)) & 1u) ? d1 * p.drop_scale : 0.f;
                        e[2 * k] = pp.x * (d0 - dl) * p.scale;
                        e[2 * k + 1] = pp.y * (d1 - dl) * p.scale;
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        st_shared_v4(adS + pt_offset(r, qc * 32 + g * 8), pack_bf16(e[g * 8], e[g * 8 + 1]), pack_bf16(e[g * 8 + 2], e[g * 8 + 3]),
                                     pack_bf16(e[g * 8 + 4], e[g * 8 + 5]), pack_bf16(e[g * 8 + 6], e[g * 8 + 7]));
                }
                tc_fence_before();
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(ds_full);
            }
            // ---- dV_j, dK_j complete: drain my 32 columns of each to global
            mbar_wait(dkv_full, j & 1);
            tc_fence_after();
            const long long tok = (long long)b * p.tok_stride_b + (long long)(j * TQ + r) * p.tok_stride_s;
            uint32_t v[32];
            if (qc < 2) {                                        // 64-wide accumulators: two of the four column-warps drain them
            const int hf = qc;
#pragma unroll
            for (int which = 0; which < 2; ++which) {           // 0: dK (col block 1), 1: dV (col block 2)
                tmem_ld32((which == 0 ? tmem_dK : tmem_dV) + lane_addr + hf * 32, v);
                tmem_ld_wait();
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
        // ---- all pairs done (dkv_full of the last j implies every MMA retired): drain dQ
        for (int i = 0; i < n && qc < 2; ++i) {
            const int hf = qc;
            uint32_t v[32];
            tmem_ld32(tmem_dQ + i * HD + lane_addr + hf * 32, v);
            tmem_ld_wait();
            const long long tok = (long long)b * p.tok_stride_b + (long long)(i * TQ + r) * p.tok_stride_s;
            bf16* o = p.dqkv + tok * (3LL * p.H) + h * HD + hf * 32;
#pragma unroll
            for (int k = 0; k < 32; k += 8)
                st_global_v4(o + k, pack_bf16(__uint_as_float(v[k]), __uint_as_float(v[k + 1])), pack_bf16(__uint_as_float(v[k + 2]), __uint_as_float(v[k + 3])),
                             pack_bf16(__uint_as_float(v[k + 4]), __uint_as_float(v[k + 5])), pack_bf16(__uint_as_float(v[k + 6]), __uint_as_float(v[k + 7])));
            if (p.dbias != nullptr) {                            // query bias gradient
                float f[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) f[k] = __bfloat162float(__float2bfloat16_rn(__uint_as_float(v[k])));
                const float cs = warp_column_sums32(f, lane);
                atomicAdd(p.dbias + h * HD + hf * 32 + lane, cs);
            }
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

static int attn_check(int B, int S, int A) {
    if (B <= 0 || A <= 0 || S <= 0 || S % TQ != 0 || S > 512) return DLE_ERR_INVALID;
    return DLE_OK;
}

extern "C" int dle_attn_fwd(const void* qkv, const float* mask, void* ctx, float* lse, int32_t B, int32_t S, int32_t A,
                            int32_t seq_first, float dropout_p, uint64_t seed, uint32_t dropout_stream, void* stream) {
    DLE_CHECK_ARG(qkv && ctx && lse && attn_check(B, S, A) == DLE_OK && dropout_p >= 0.f && dropout_p < 1.f);
    const int H = A * HD;
    CUtensorMap tm;
    int rc = make_tmap_tokens_3d(&tm, qkv, B, S, 3 * H, seq_first);
    if (rc != DLE_OK) return rc;
    AttnFwdParams p;
    p.tok_stride_s = seq_first ? B : 1; p.tok_stride_b = seq_first ? 1 : S;
    p.mask = mask; p.ctx = reinterpret_cast<bf16*>(ctx); p.lse = lse; p.B = B; p.S = S; p.A = A; p.H = H;
    p.scale_log2 = 0.125f * LOG2E;
    p.drop_thresh = dropout_p > 0.f ? dropout_thresh16(dropout_p) : 0u;
    p.drop_scale = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
    p.drop_stream = dropout_stream; p.seed = seed;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM_BYTES) != cudaSuccess) return DLE_ERR_CUDA;
        cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr_set = true;
    }
    attn_fwd_kernel<<<dim3(S / TQ, A, B), FWD_THREADS, FWD_SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(tm, p);
    DLE_LAUNCH_CHECK();
    return DLE_OK;
}

extern "C" int dle_attn_bwd(const void* qkv, const float* mask, const void* ctx, const void* dctx, const float* lse, void* dqkv,
                            float* delta_ws, float* dbias_qkv, int32_t B, int32_t S, int32_t A, int32_t seq_first, float dropout_p, uint64_t seed,
                            uint32_t dropout_stream, void* stream) {
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
    p.mask = mask; p.lse = lse; p.delta = delta_ws; p.dbias = dbias_qkv; p.dqkv = reinterpret_cast<bf16*>(dqkv);
    p.B = B; p.S = S; p.A = A; p.H = H; p.tok_stride_s = seq_first ? B : 1; p.tok_stride_b = seq_first ? 1 : S; p.scale = 0.125f; p.scale_log2 = 0.125f * LOG2E;
    p.drop_thresh = dropout_p > 0.f ? dropout_thresh16(dropout_p) : 0u;
    p.drop_scale = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
    p.drop_stream = dropout_stream; p.seed = seed;
    const int smem = bwd_smem_bytes(S);
    static int attr_smem = 0;
    if (smem > attr_smem) {
        if (cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return DLE_ERR_CUDA;
        attr_smem = smem;
    }
    attn_bwd_kernel<<<dim3(A, B), BWD_THREADS, smem, st>>>(tq, td, p);
    DLE_LAUNCH_CHECK();
    return DLE_OK;
}
