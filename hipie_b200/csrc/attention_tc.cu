// tcgen05 flash attention for the ViT-H global blocks (hd = 80, T % 128 == 0), sm_100a.
//
//   softmax((q*scale) k^T + rel_h[q, kh] + rel_w[q, kw]) v      (/root/reference/projects/HIPIE/hipie/backbone/vit.py:67-83,
//                                                                 backbone/utils.py:96-125)
//
// One CTA = 128 query rows of one (batch, head); 320 threads.  Both A operands of the two MMAs live in TENSOR MEMORY
// (TS-mode tcgen05.mma), so the tensor core only streams the small K / V^T tiles from shared memory:
//   warp 0      TMA producer: K tiles (64 keys x 80) and V^T tiles (80 x 64 keys) into a 3-stage, hardware-swizzled
//               shared-memory ring (head dim 80 = one 128B-swizzled 64-wide box + one 32B-swizzled 16-wide box)
//   warp 1      MMA issuer: S_j = Q K_j^T (A = Q in TMEM, M=128, N=64, K=80) into one of two TMEM score buffers and
//               O += P_j V_j (A = P_j in TMEM, M=128, N=80, K=64); QK runs two tiles ahead of PV
//   warps 2-9   softmax, two warps per TMEM lane quarter (each owns 32 of the tile's 64 key columns): load their Q row
//               from global once and park it in TMEM as packed bf16 pairs; per tile tcgen05.ld the scores, scale +
//               rel-pos bias (rel_w hoisted in registers, one prefetched rel_h scalar per tile since a 64-key tile is one
//               key row of the 64-wide grid), online softmax with lazy rescaling of the TMEM accumulator, and P written
//               back to TMEM with tcgen05.st (double buffered) as the A operand of the PV MMA
// Precision: PREC==3 evaluates Qh.Kh + Qh.Kl + Ql.Kh and Ph.Vh + Ph.Vl + Pl.Vh (bf16x3, fp32-class); PREC==1 plain bf16.
// History (profiles/): the first version kept Q and P in shared memory; ncu + a clock64 trace showed the tensor
// pipe starved re-reading the 4 KB A tile for every small-N instruction (≈1000 cycles per MMA batch), hence this layout.
#include "common.cuh"
#include "ptx.cuh"

namespace hipie {
using namespace ptx;

int make_tmap_bf16(CUtensorMap* out, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int batch,
                   int64_t bstride, int box_rows, int box_cols);

constexpr int FA_BM = 128, FA_BN = 64, FA_HD = 80, FA_STAGES = 3;
// tensor-memory map (columns)
constexpr int TM_S = 0;            // 2 x 64 fp32 score columns
constexpr int TM_O = 128;          // 80 fp32 output columns
constexpr int TM_Q = 208;          // Q hi: 40 cols of packed bf16 pairs, Q lo: next 40
constexpr int TM_P = 288;          // P: [buf][hi 32 cols | lo 32 cols]

struct FaMaps {
    CUtensorMap k64[2], k16[2], vt[2];   // [hi, lo]
};

struct FaParams {
    const __nv_bfloat16 *q_hi, *q_lo;
    int64_t q_bs, q_ts;
    const float *rel_h, *rel_w;   // (B,H,T,kh), (B,H,T,64) or null
    int kh;
    float* out_f32;
    __nv_bfloat16 *out_hi, *out_lo;
    int64_t o_bs, o_ts;
    int B, H, T;
    int q_col0, k_col0;           // column of head 0 inside the q / k rows
    float scale_log2e;
    long long* trace;             // debug: per-tile clock64 timestamps of CTA (0,0,0) or null
};

template <int PREC>
struct FaSmem {
    static constexpr int NPL = PREC == 3 ? 2 : 1;
    static constexpr int K64 = FA_BN * 128, K16 = FA_BN * 32;
    static constexpr int VT = FA_HD * 128;
    static constexpr int K_STAGE = NPL * (K64 + K16);
    static constexpr int V_STAGE = NPL * VT;
    static constexpr int OFF_V = FA_STAGES * K_STAGE;
    static constexpr int OFF_BAR = OFF_V + FA_STAGES * V_STAGE;
    static constexpr int TOTAL = OFF_BAR + 256 + 3 * 2 * FA_BM * 4 + 1024;
};

template <int PREC>
__global__ void __launch_bounds__(320, 1)
attn_tc_kernel(const __grid_constant__ FaMaps maps, const FaParams p) {
    using SM = FaSmem<PREC>;
    constexpr int NPL = SM::NPL;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::OFF_BAR);
    uint64_t* q_ready = bars;                      // [1]  Q parked in TMEM (8 warps)
    uint64_t* k_full = bars + 1;                   // [STAGES]
    uint64_t* k_empty = k_full + FA_STAGES;
    uint64_t* v_full = k_empty + FA_STAGES;
    uint64_t* v_empty = v_full + FA_STAGES;
    uint64_t* s_full = v_empty + FA_STAGES;        // [2]
    uint64_t* s_empty = s_full + 2;                // [2]
    uint64_t* p_full = s_empty + 2;                // [2]
    uint64_t* pv_done = p_full + 2;                // [2]  PV(j) commits to pv_done[j & 1]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 2);
    float* xchg = reinterpret_cast<float*>(smem + SM::OFF_BAR + 256);   // [3][2][128] max / row-sum exchange

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * FA_BM, h = blockIdx.y, b = blockIdx.z;
    const int ntiles = p.T / FA_BN;
    const bool trace_cta = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < NPL; ++i) {
            prefetch_tmap(&maps.k64[i]); prefetch_tmap(&maps.k16[i]); prefetch_tmap(&maps.vt[i]);
        }
    }
    if (warp == 1) {
        if (lane == 0) {
            mbar_init(q_ready, 8);
            for (int i = 0; i < FA_STAGES; ++i) {
                mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
                mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
            }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 8);
                mbar_init(&p_full[i], 8); mbar_init(&pv_done[i], 1);
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int j = 0; j < ntiles; ++j) {
                mbar_wait(&k_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&k_full[s], SM::K_STAGE);
                uint8_t* kb = smem + s * SM::K_STAGE;
                for (int pl = 0; pl < NPL; ++pl) {
                    tma_load_3d(kb + pl * (SM::K64 + SM::K16), &maps.k64[pl], &k_full[s], p.k_col0 + h * FA_HD, j * FA_BN, b);
                    tma_load_3d(kb + pl * (SM::K64 + SM::K16) + SM::K64, &maps.k16[pl], &k_full[s], p.k_col0 + h * FA_HD + 64, j * FA_BN, b);
                }
                mbar_wait(&v_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&v_full[s], SM::V_STAGE);
                uint8_t* vb = smem + SM::OFF_V + s * SM::V_STAGE;
                for (int pl = 0; pl < NPL; ++pl)
                    tma_load_3d(vb + pl * SM::VT, &maps.vt[pl], &v_full[s], b * p.T + j * FA_BN, h * FA_HD, 0);
                if (++s == FA_STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc_qk = make_idesc_bf16(FA_BM, FA_BN);
        constexpr uint32_t idesc_pv = make_idesc_bf16(FA_BM, FA_HD);
        const uint32_t tq_hi = tmem_base + TM_Q, tq_lo = tmem_base + TM_Q + 40;
        auto issue_qk = [&](int j) {
            const int st = j % FA_STAGES, sb = j & 1;
            const bool trq = trace_cta && lane == 0 && j < 32;
            if (trq) p.trace[j * 16 + 7] = clock64();
            mbar_wait(&k_full[st], (j / FA_STAGES) & 1);
            mbar_wait(&s_empty[sb], ((j >> 1) & 1) ^ 1);
            tc_fence_after();
            if (trq) p.trace[j * 16 + 8] = clock64();
            if (elect_one()) {
                const uint32_t kb = smem_u32(smem + st * SM::K_STAGE);
                const uint32_t d = tmem_base + TM_S + sb * FA_BN;
                const uint64_t k64h = make_kmajor_desc<128>(kb), k16h = make_kmajor_desc<32>(kb + SM::K64);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ts(d, tq_hi + 8 * k, k64h + 2 * k, idesc_qk, k > 0);
                umma_f16_ts(d, tq_hi + 32, k16h, idesc_qk, 1);
                if (PREC == 3) {
                    const uint64_t k64l = make_kmajor_desc<128>(kb + SM::K64 + SM::K16), k16l = make_kmajor_desc<32>(kb + 2 * SM::K64 + SM::K16);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16_ts(d, tq_hi + 8 * k, k64l + 2 * k, idesc_qk, 1);
                    umma_f16_ts(d, tq_hi + 32, k16l, idesc_qk, 1);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16_ts(d, tq_lo + 8 * k, k64h + 2 * k, idesc_qk, 1);
                    umma_f16_ts(d, tq_lo + 32, k16h, idesc_qk, 1);
                }
                umma_commit(&s_full[sb]);
                umma_commit(&k_empty[st]);
            }
            if (trq) p.trace[j * 16 + 9] = clock64();
            __syncwarp();
        };
        auto issue_pv = [&](int j) {
            const int st = j % FA_STAGES, sb = j & 1;
            const bool trm = trace_cta && lane == 0 && j < 32;
            if (trm) p.trace[j * 16 + 10] = clock64();
            mbar_wait(&p_full[sb], (j >> 1) & 1);
            if (trm) p.trace[j * 16 + 11] = clock64();
            mbar_wait(&v_full[st], (j / FA_STAGES) & 1);
            tc_fence_after();
            if (trm) p.trace[j * 16 + 12] = clock64();
            if (elect_one()) {
                const uint32_t vb = smem_u32(smem + SM::OFF_V + st * SM::V_STAGE);
                const uint32_t d = tmem_base + TM_O;
                const uint32_t tp_hi = tmem_base + TM_P + sb * 64, tp_lo = tp_hi + 32;
                const uint64_t v_hi = make_kmajor_desc<128>(vb);
                uint32_t accum = j > 0;
#pragma unroll
                for (int k = 0; k < FA_BN / 16; ++k) { umma_f16_ts(d, tp_hi + 8 * k, v_hi + 2 * k, idesc_pv, accum); accum = 1; }
                if (PREC == 3) {
                    const uint64_t v_lo = make_kmajor_desc<128>(vb + SM::VT);
#pragma unroll
                    for (int k = 0; k < FA_BN / 16; ++k) umma_f16_ts(d, tp_hi + 8 * k, v_lo + 2 * k, idesc_pv, 1);
#pragma unroll
                    for (int k = 0; k < FA_BN / 16; ++k) umma_f16_ts(d, tp_lo + 8 * k, v_hi + 2 * k, idesc_pv, 1);
                }
                umma_commit(&pv_done[sb]);
                umma_commit(&v_empty[st]);
            }
            if (trm) p.trace[j * 16 + 13] = clock64();
            __syncwarp();
        };
        // PV(j) (accumulator O) and QK(j+2) (accumulator S[j&1]) are independent chains: small-N MMAs into the same
        // accumulator are latency bound (~70 cycles each, measured), so the two chains are issued interleaved.
        auto issue_pv_qk = [&](int jp, int jq) {
            const int stp = jp % FA_STAGES, sbp = jp & 1;
            const int stq = jq % FA_STAGES, sbq = jq & 1;
            mbar_wait(&k_full[stq], (jq / FA_STAGES) & 1);
            mbar_wait(&s_empty[sbq], ((jq >> 1) & 1) ^ 1);
            mbar_wait(&v_full[stp], (jp / FA_STAGES) & 1);
            mbar_wait(&p_full[sbp], (jp >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t vb = smem_u32(smem + SM::OFF_V + stp * SM::V_STAGE);
                const uint32_t kb = smem_u32(smem + stq * SM::K_STAGE);
                const uint32_t dO = tmem_base + TM_O, dS = tmem_base + TM_S + sbq * FA_BN;
                const uint32_t tp_hi = tmem_base + TM_P + sbp * 64, tp_lo = tp_hi + 32;
                const uint64_t v_hi = make_kmajor_desc<128>(vb), v_lo = make_kmajor_desc<128>(vb + SM::VT);
                const uint64_t k64h = make_kmajor_desc<128>(kb), k16h = make_kmajor_desc<32>(kb + SM::K64);
                const uint64_t k64l = make_kmajor_desc<128>(kb + SM::K64 + SM::K16), k16l = make_kmajor_desc<32>(kb + 2 * SM::K64 + SM::K16);
                const uint32_t acc0 = jp > 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    umma_f16_ts(dO, tp_hi + 8 * k, v_hi + 2 * k, idesc_pv, k > 0 ? 1u : acc0);
                    umma_f16_ts(dS, tq_hi + 8 * k, k64h + 2 * k, idesc_qk, k > 0);
                }
                umma_f16_ts(dS, tq_hi + 32, k16h, idesc_qk, 1);
                if (PREC == 3) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        umma_f16_ts(dO, tp_hi + 8 * k, v_lo + 2 * k, idesc_pv, 1);
                        umma_f16_ts(dS, tq_hi + 8 * k, k64l + 2 * k, idesc_qk, 1);
                    }
                    umma_f16_ts(dS, tq_hi + 32, k16l, idesc_qk, 1);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        umma_f16_ts(dO, tp_lo + 8 * k, v_hi + 2 * k, idesc_pv, 1);
                        umma_f16_ts(dS, tq_lo + 8 * k, k64h + 2 * k, idesc_qk, 1);
                    }
                    umma_f16_ts(dS, tq_lo + 32, k16h, idesc_qk, 1);
                }
                umma_commit(&pv_done[sbp]);
                umma_commit(&v_empty[stp]);
                umma_commit(&s_full[sbq]);
                umma_commit(&k_empty[stq]);
            }
            __syncwarp();
        };
        mbar_wait(q_ready, 0);
        tc_fence_after();
        issue_qk(0);
        if (ntiles > 1) issue_qk(1);
        for (int j = 0; j < ntiles; ++j) {
            if (j + 2 < ntiles) issue_pv_qk(j, j + 2);
            else issue_pv(j);
        }
    } else {
        // ===================== softmax / epilogue (warps 2..9) =====================
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int r = quarter * 32 + lane;            // row of the Q tile == TMEM lane
        const int qrow = q0 + r;
        const uint32_t tm = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const bool has_rel = p.rel_h != nullptr;
        constexpr int HC = FA_BN / 2;                 // 32 key columns per thread
        constexpr float LOG2E = 1.4426950408889634f;
        // ---- park this thread's half of the Q row in TMEM (packed bf16 pairs = the TS-mode A operand layout) ----
        {
            const int64_t qoff = (int64_t)b * p.q_bs + (int64_t)qrow * p.q_ts + p.q_col0 + h * FA_HD + half * 40;
            for (int pl = 0; pl < NPL; ++pl) {
                const uint4* src = reinterpret_cast<const uint4*>((pl == 0 ? p.q_hi : p.q_lo) + qoff);
                uint32_t w[16];
                const uint4 a0 = src[0], a1 = src[1], a2 = src[2], a3 = src[3], a4 = src[4];
                w[0] = a0.x; w[1] = a0.y; w[2] = a0.z; w[3] = a0.w; w[4] = a1.x; w[5] = a1.y; w[6] = a1.z; w[7] = a1.w;
                w[8] = a2.x; w[9] = a2.y; w[10] = a2.z; w[11] = a2.w; w[12] = a3.x; w[13] = a3.y; w[14] = a3.z; w[15] = a3.w;
                tmem_st_32x32b_x16(tm + TM_Q + pl * 40 + half * 20, w);
                tmem_st_32x32b_x4(tm + TM_Q + pl * 40 + half * 20 + 16, a4.x, a4.y, a4.z, a4.w);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(q_ready);
        }
        float rw[HC];
        const float* relh_row = nullptr;
        if (has_rel) {
            const int64_t rowi = ((int64_t)b * p.H + h) * p.T + qrow;
            const float* rwp = p.rel_w + rowi * FA_BN + half * HC;
#pragma unroll
            for (int i = 0; i < HC; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(rwp + i);
                rw[i] = v.x * LOG2E; rw[i + 1] = v.y * LOG2E; rw[i + 2] = v.z * LOG2E; rw[i + 3] = v.w * LOG2E;
            }
            relh_row = p.rel_h + rowi * p.kh;
        }
        float m = -INFINITY, l = 0.f;          // m: running reference max (log2 domain, incl. rel-pos terms)
        const int o_c0 = half == 0 ? 0 : 48, o_c1 = half == 0 ? 48 : FA_HD;   // output columns owned (x16 granules)
        float rh_next = has_rel ? __ldg(relh_row) * LOG2E : 0.f;
        for (int j = 0; j < ntiles; ++j) {
            const int s = j & 1;
            const bool tr = trace_cta && warp == 2 && lane == 0 && j < 32;
            const float rh = rh_next;
            if (has_rel && j + 1 < ntiles) rh_next = __ldg(relh_row + j + 1) * LOG2E;   // prefetch for the next tile
            if (tr) p.trace[j * 16 + 0] = clock64();
            mbar_wait(&s_full[s], (j >> 1) & 1);
            tc_fence_after();
            if (tr) p.trace[j * 16 + 1] = clock64();
            uint32_t sv[HC];
            tmem_ld_32x32b_x32(tm + TM_S + s * FA_BN + half * HC, sv);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[s]);   // score buffer may be overwritten by QK(j+2)
            if (tr) p.trace[j * 16 + 2] = clock64();
            // t_i = s_i*scale*log2e + rel_w_i ; rel_h is uniform over the tile -> folded into the max / exponent offset
            float tmax = -INFINITY;
            float t[HC];
#pragma unroll
            for (int i = 0; i < HC; ++i) {
                t[i] = fmaf(__uint_as_float(sv[i]), p.scale_log2e, has_rel ? rw[i] : 0.f);
                tmax = fmaxf(tmax, t[i]);
            }
            tmax += rh;
            xchg[(s * 2 + half) * FA_BM + r] = tmax;    // exchange the tile max with the partner warp (other column half)
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
            tmax = fmaxf(tmax, xchg[(s * 2 + (half ^ 1)) * FA_BM + r]);
            if (tr) p.trace[j * 16 + 3] = clock64();
            // lazy rescale: keep the running reference max unless the tile exceeds it by more than 2^8
            float corr = 1.f;
            const bool need = tmax > m + 8.f;
            if (need) {
                corr = ex2_approx(m - tmax);   // m == -inf -> 0
                m = tmax;
                l *= corr;
            }
            const float off = m - rh;
            float rowsum = 0.f;
#pragma unroll
            for (int i = 0; i < HC; ++i) {
                t[i] = ex2_approx(t[i] - off);
                rowsum += t[i];
            }
            l += rowsum;
            if (tr) p.trace[j * 16 + 4] = clock64();
            if (j > 0 && __any_sync(0xffffffffu, need)) {
                // O must reflect PV(j-1) before it is rescaled (rare once the running max has settled)
                mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
                tc_fence_after();
                for (int c = o_c0; c < o_c1; c += 16) {
                    uint32_t o[16];
                    tmem_ld_32x32b_x16(tm + TM_O + c, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
                    tmem_st_32x32b_x16(tm + TM_O + c, o);
                }
                tmem_st_wait();
            }
            // P buffer s is free once PV(j-2) has retired
            mbar_wait(&pv_done[s], ((j >> 1) & 1) ^ 1);
            tc_fence_after();
            if (tr) p.trace[j * 16 + 5] = clock64();
            {
                uint32_t ph_[16], pl_[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (PREC == 3) split2(t[2 * i], t[2 * i + 1], ph_[i], pl_[i]);
                    else ph_[i] = pack_bf16x2(t[2 * i], t[2 * i + 1]);
                }
                tmem_st_32x32b_x16(tm + TM_P + s * 64 + half * 16, ph_);
                if (PREC == 3) tmem_st_32x32b_x16(tm + TM_P + s * 64 + 32 + half * 16, pl_);
                tmem_st_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[s]);
            if (tr) p.trace[j * 16 + 6] = clock64();
        }
        // ---- epilogue: O / l (row sum = both halves) ----
        xchg[(2 * 2 + half) * FA_BM + r] = l;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
        l += xchg[(2 * 2 + (half ^ 1)) * FA_BM + r];
        mbar_wait(&pv_done[(ntiles - 1) & 1], ((ntiles - 1) >> 1) & 1);
        tc_fence_after();
        const float inv = 1.f / l;
        const int64_t obase = (int64_t)b * p.o_bs + (int64_t)qrow * p.o_ts + (int64_t)h * FA_HD;
        for (int c = o_c0; c < o_c1; c += 16) {
            uint32_t o[16];
            tmem_ld_32x32b_x16(tm + TM_O + c, o);
            tmem_ld_wait();
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(o[i]) * inv;
            if (p.out_f32) {
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                    *reinterpret_cast<float4*>(p.out_f32 + obase + c + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
            }
            if (p.out_hi) {
                uint4 h0, h1, l0, l1;
                split2(f[0], f[1], h0.x, l0.x); split2(f[2], f[3], h0.y, l0.y);
                split2(f[4], f[5], h0.z, l0.z); split2(f[6], f[7], h0.w, l0.w);
                split2(f[8], f[9], h1.x, l1.x); split2(f[10], f[11], h1.y, l1.y);
                split2(f[12], f[13], h1.z, l1.z); split2(f[14], f[15], h1.w, l1.w);
                *reinterpret_cast<uint4*>(p.out_hi + obase + c) = h0;
                *reinterpret_cast<uint4*>(p.out_hi + obase + c + 8) = h1;
                if (p.out_lo) {
                    *reinterpret_cast<uint4*>(p.out_lo + obase + c) = l0;
                    *reinterpret_cast<uint4*>(p.out_lo + obase + c + 8) = l1;
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int PREC>
static int launch_fa(const FaMaps& maps, const FaParams& p, cudaStream_t st) {
    using SM = FaSmem<PREC>;
    static bool attr = false;
    if (!attr) {
        HIPIE_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel<PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
        attr = true;
    }
    dim3 grid(p.T / FA_BM, p.H, p.B);
    attn_tc_kernel<PREC><<<grid, 320, SM::TOTAL, st>>>(maps, p);
    HIPIE_CHECK_LAUNCH();
    return HIPIE_OK;
}

}  // namespace hipie

using namespace hipie;

extern "C" int hipie_attention_tc_traced(const void* q_hi, const void* q_lo, int64_t q_bs, int64_t q_ts, int q_col0, int q_width,
                                         const void* k_hi, const void* k_lo, int64_t k_bs, int64_t k_ts, int k_col0, int k_width,
                                         const void* vt_hi, const void* vt_lo, int64_t vt_ld, const float* rel_h, const float* rel_w,
                                         int kh, int kw, float* out_f32, void* out_hi, void* out_lo, int64_t o_bs, int64_t o_ts, int B,
                                         int H, int T, int hd, float scale, int prec, long long* trace, void* stream) {
    HIPIE_CHECK_ARG(q_hi && k_hi && vt_hi, "hipie_attention_tc: q/k/vt hi planes required");
    HIPIE_CHECK_ARG(prec == 1 || (prec == 3 && q_lo && k_lo && vt_lo), "hipie_attention_tc: prec/lo planes mismatch");
    HIPIE_CHECK_ARG(hd == FA_HD, "hipie_attention_tc: head dim must be 80 (got %d)", hd);
    HIPIE_CHECK_ARG(T > 0 && T % FA_BM == 0, "hipie_attention_tc: T (%d) must be a multiple of 128", T);
    HIPIE_CHECK_ARG((rel_h == nullptr) == (rel_w == nullptr), "hipie_attention_tc: rel_h and rel_w go together");
    HIPIE_CHECK_ARG(!rel_h || (kw == FA_BN && kh * kw == T), "hipie_attention_tc: rel-pos needs kw == 64 and kh*kw == T");
    HIPIE_CHECK_ARG(out_f32 || out_hi, "hipie_attention_tc: no output requested");
    HIPIE_CHECK_ARG(q_ts % 8 == 0 && q_bs % 8 == 0 && q_col0 % 8 == 0 && (reinterpret_cast<uintptr_t>(q_hi) & 15) == 0,
                    "hipie_attention_tc: q rows must be 16-byte aligned");
    (void)q_width;
    FaMaps maps;
    int rc;
    const void* kp[2] = {k_hi, k_lo};
    const void* vp[2] = {vt_hi, vt_lo};
    for (int pl = 0; pl < (prec == 3 ? 2 : 1); ++pl) {
        if ((rc = make_tmap_bf16(&maps.k64[pl], kp[pl], T, k_width, k_ts, B, k_bs, FA_BN, 64))) return rc;
        if ((rc = make_tmap_bf16(&maps.k16[pl], kp[pl], T, k_width, k_ts, B, k_bs, FA_BN, 16))) return rc;
        if ((rc = make_tmap_bf16(&maps.vt[pl], vp[pl], (int64_t)H * FA_HD, (int64_t)B * T, vt_ld, 1, 0, FA_HD, 64))) return rc;
    }
    if (prec == 1) { maps.k64[1] = maps.k64[0]; maps.k16[1] = maps.k16[0]; maps.vt[1] = maps.vt[0]; }
    FaParams p;
    p.q_hi = (const __nv_bfloat16*)q_hi; p.q_lo = (const __nv_bfloat16*)q_lo; p.q_bs = q_bs; p.q_ts = q_ts;
    p.rel_h = rel_h; p.rel_w = rel_w; p.kh = kh;
    p.out_f32 = out_f32; p.out_hi = (__nv_bfloat16*)out_hi; p.out_lo = (__nv_bfloat16*)out_lo;
    p.o_bs = o_bs; p.o_ts = o_ts; p.B = B; p.H = H; p.T = T; p.q_col0 = q_col0; p.k_col0 = k_col0;
    p.scale_log2e = scale * 1.4426950408889634f;
    p.trace = trace;
    cudaStream_t st = (cudaStream_t)stream;
    return prec == 3 ? launch_fa<3>(maps, p, st) : launch_fa<1>(maps, p, st);
}

// q / k: bf16 planes viewed as (B, T, row_width) with token stride q_ts / k_ts and batch stride q_bs / k_bs (elements);
// head h occupies columns [q_col0 + 80 h, +80).  vt: V transposed, (H*80 rows, B*T columns) planes with row stride vt_ld.
extern "C" int hipie_attention_tc(const void* q_hi, const void* q_lo, int64_t q_bs, int64_t q_ts, int q_col0, int q_width,
                                  const void* k_hi, const void* k_lo, int64_t k_bs, int64_t k_ts, int k_col0, int k_width,
                                  const void* vt_hi, const void* vt_lo, int64_t vt_ld, const float* rel_h, const float* rel_w,
                                  int kh, int kw, float* out_f32, void* out_hi, void* out_lo, int64_t o_bs, int64_t o_ts, int B,
                                  int H, int T, int hd, float scale, int prec, void* stream) {
    return hipie_attention_tc_traced(q_hi, q_lo, q_bs, q_ts, q_col0, q_width, k_hi, k_lo, k_bs, k_ts, k_col0, k_width, vt_hi, vt_lo, vt_ld, rel_h,
                                     rel_w, kh, kw, out_f32, out_hi, out_lo, o_bs, o_ts, B, H, T, hd, scale, prec, nullptr, stream);
}
