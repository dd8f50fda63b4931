// Fused semantic + panoptic post-processing (reference H/models/hipie_img.py:880-1023).
//
// The reference upsamples every query's 1/4-resolution mask logit to full resolution (bilinear, align_corners=False),
// takes the sigmoid, and then (a) contracts it with the per-query class probabilities into a (C, H, W) semantic map
// (einsum "qc,qhw->chw") and (b) takes an argmax over score*sigmoid for the kept queries plus three per-query pixel
// counts for the panoptic merge.  As separate ops that is Q*H*W fp32 written and re-read four times (~5 GB per image at
// Q=1200, 1024^2).  Here one kernel does it all from the low-resolution logits: each CTA owns a 32x8 pixel tile, stages
// the 10x4 low-resolution taps of 64 queries at a time in shared memory, evaluates sigmoid(bilinear()) straight into
// mma.sync A fragments (bf16 hi/lo split), and accumulates the (pixels x classes) tile on the tensor cores in bf16x3
// against the class-probability chunk; the argmax / area counters ride along on the same sigmoid values.
// HBM traffic drops to the low-resolution logits (+halo) and the outputs.
#include "common.cuh"

namespace hipie {

namespace {

constexpr int PP_TW = 32, PP_TH = 8;          // pixel tile per CTA (one warp per row, two m16 tiles per warp)
constexpr int PP_QC = 64;                     // queries per staged chunk
constexpr int PP_TR = 4, PP_TC = 10;          // low-resolution rows / cols a tile touches (stride 4)
constexpr int PP_TQ = PP_TR * PP_TC + 1;      // padded floats per query in the tap tile
constexpr int PP_PQ = PP_QC + 8;              // padded bf16 per class row of the staged P^T chunk
constexpr int PP_THREADS = 256;

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// masks (Q,h,w) f32 | pt_hi/pt_lo (NT*8, Qpad) bf16 = class probabilities transposed, split | scores (Qpad) f32, <0 = not kept
// sem (C,Hc,Wc) f32 | ids (Hc,Wc) i32 = -1 or 2*q + (sigmoid_q >= .5) | areas (3,Q) i32 = mask / original / intersection
template <int NT>
__global__ void __launch_bounds__(PP_THREADS)
seg_post_kernel(const float* __restrict__ masks, const __nv_bfloat16* __restrict__ pt_hi,
                const __nv_bfloat16* __restrict__ pt_lo, const float* __restrict__ scores, float* __restrict__ sem,
                int* __restrict__ ids, int* __restrict__ areas, int Q, int Qpad, int C, int h, int w, int Hc, int Wc) {
    extern __shared__ __align__(16) unsigned char pp_smem[];
    float* taps = reinterpret_cast<float*>(pp_smem);                                   // [PP_QC][PP_TQ]
    static_assert((PP_QC * PP_TQ * 4) % 16 == 0, "P^T chunk must stay 16-byte aligned");
    __nv_bfloat16* ph = reinterpret_cast<__nv_bfloat16*>(taps + PP_QC * PP_TQ);       // [NT*8][PP_PQ]
    __nv_bfloat16* pl = ph + NT * 8 * PP_PQ;
    float* sc = reinterpret_cast<float*>(pl + NT * 8 * PP_PQ);                         // [PP_QC]
    int* cnt = reinterpret_cast<int*>(sc + PP_QC);                                     // [3][Qpad]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int tx = blockIdx.x * PP_TW, ty = blockIdx.y * PP_TH;
    const int x_lo = tx / 4 - 1, y_lo = ty / 4 - 1;

    for (int i = tid; i < 3 * Qpad; i += PP_THREADS) cnt[i] = 0;

    // this thread's pixels: row y, columns tx + {g, g+8, g+16, g+24}  (slot ps = 2*mt + half)
    const int y = ty + warp;
    int r0, r1;
    float ly;
    {
        float s = fmaxf(0.25f * (y + 0.5f) - 0.5f, 0.f);
        int y0 = min((int)s, h - 1);
        ly = s - y0;
        r0 = y0 - y_lo;
        r1 = y0 + (y0 < h - 1) - y_lo;
    }
    const float ly0 = 1.f - ly;
    int o00[4], o01[4], o10[4], o11[4];
    float lx[4];
    bool pv[4];
#pragma unroll
    for (int ps = 0; ps < 4; ++ps) {
        const int x = tx + ps * 8 + g;
        float s = fmaxf(0.25f * (x + 0.5f) - 0.5f, 0.f);
        int x0 = min((int)s, w - 1);
        lx[ps] = s - x0;
        const int c0 = x0 - x_lo, c1 = x0 + (x0 < w - 1) - x_lo;
        o00[ps] = r0 * PP_TC + c0;
        o01[ps] = r0 * PP_TC + c1;
        o10[ps] = r1 * PP_TC + c0;
        o11[ps] = r1 * PP_TC + c1;
        pv[ps] = (x < Wc) && (y < Hc);
    }

    float acc[2][NT][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
    float best[4], bsig[4];
    int bidx[4];
#pragma unroll
    for (int ps = 0; ps < 4; ++ps) {
        best[ps] = -1.f;
        bsig[ps] = 0.f;
        bidx[ps] = -1;
    }
    const uint32_t tmask = 0x11111111u << t;

    for (int q0 = 0; q0 < Qpad; q0 += PP_QC) {
        __syncthreads();
        // ---- stage taps, P^T chunk, scores ---------------------------------------------------------------------
        for (int i = tid; i < PP_QC * PP_TR * PP_TC; i += PP_THREADS) {
            const int ql = i / (PP_TR * PP_TC), rc = i - ql * (PP_TR * PP_TC);
            const int r = rc / PP_TC, c = rc - r * PP_TC;
            const int q = q0 + ql;
            float v = -1e30f;                        // padded queries: sigmoid -> 0
            if (q < Q) {
                const int yy = min(max(y_lo + r, 0), h - 1), xx = min(max(x_lo + c, 0), w - 1);
                v = __ldg(masks + ((size_t)q * h + yy) * w + xx);
            }
            taps[ql * PP_TQ + rc] = v;
        }
        for (int i = tid; i < NT * 8 * (PP_QC / 8); i += PP_THREADS) {
            const int row = i / (PP_QC / 8), seg = i - row * (PP_QC / 8);
            const uint4 vh = *reinterpret_cast<const uint4*>(pt_hi + (size_t)row * Qpad + q0 + seg * 8);
            const uint4 vl = *reinterpret_cast<const uint4*>(pt_lo + (size_t)row * Qpad + q0 + seg * 8);
            *reinterpret_cast<uint4*>(ph + row * PP_PQ + seg * 8) = vh;
            *reinterpret_cast<uint4*>(pl + row * PP_PQ + seg * 8) = vl;
        }
        if (tid < PP_QC) sc[tid] = scores[q0 + tid];
        __syncthreads();

#pragma unroll 1
        for (int ks = 0; ks < PP_QC / 16; ++ks) {
            float sg[4][4];                           // [ps][j]  sigmoid(upsampled logit)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ql = ks * 16 + 2 * t + (j & 1) + (j >> 1) * 8;
                const float* T = taps + ql * PP_TQ;
                const float s_q = sc[ql];
                int n_orig = 0;
#pragma unroll
                for (int ps = 0; ps < 4; ++ps) {
                    const float lx1 = lx[ps], lx0 = 1.f - lx1;
                    const float v = ly0 * (lx0 * T[o00[ps]] + lx1 * T[o01[ps]]) + ly * (lx0 * T[o10[ps]] + lx1 * T[o11[ps]]);
                    const float s = rcp_approx(1.f + ex2_approx(-1.4426950408889634f * v));
                    sg[ps][j] = s;
                    const float pr = s * s_q;
                    if (s_q >= 0.f && pr > best[ps]) {
                        best[ps] = pr;
                        bsig[ps] = s;
                        bidx[ps] = q0 + ql;
                    }
                    const uint32_t b = __ballot_sync(0xffffffffu, pv[ps] && s >= 0.5f);
                    n_orig += __popc(b & tmask);
                }
                if (g == 0 && n_orig) atomicAdd(&cnt[Qpad + q0 + ql], n_orig);
            }
            uint32_t ah[2][4], al[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                split2(sg[2 * mt][0], sg[2 * mt][1], ah[mt][0], al[mt][0]);
                split2(sg[2 * mt + 1][0], sg[2 * mt + 1][1], ah[mt][1], al[mt][1]);
                split2(sg[2 * mt][2], sg[2 * mt][3], ah[mt][2], al[mt][2]);
                split2(sg[2 * mt + 1][2], sg[2 * mt + 1][3], ah[mt][3], al[mt][3]);
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int off = (nt * 8 + g) * PP_PQ + ks * 16 + 2 * t;
                const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(ph + off);
                const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(ph + off + 8);
                const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(pl + off);
                const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(pl + off + 8);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    mma16816(acc[mt][nt], ah[mt], bh0, bh1);
                    mma16816(acc[mt][nt], ah[mt], bl0, bl1);
                    mma16816(acc[mt][nt], al[mt], bh0, bh1);
                }
            }
        }
    }

    // ---- argmax across the quad (ties -> lowest query index, like torch.argmax over the kept list) -------------
#pragma unroll
    for (int ps = 0; ps < 4; ++ps) {
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best[ps], o);
            const float os = __shfl_xor_sync(0xffffffffu, bsig[ps], o);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx[ps], o);
            const bool take = (oi >= 0) && (bidx[ps] < 0 || ob > best[ps] || (ob == best[ps] && oi < bidx[ps]));
            if (take) {
                best[ps] = ob;
                bsig[ps] = os;
                bidx[ps] = oi;
            }
        }
        if (t == 0 && pv[ps]) {
            const int x = tx + ps * 8 + g;
            const int in = bsig[ps] >= 0.5f;
            ids[(size_t)y * Wc + x] = bidx[ps] < 0 ? -1 : 2 * bidx[ps] + in;
            if (bidx[ps] >= 0) {
                atomicAdd(&cnt[bidx[ps]], 1);
                if (in) atomicAdd(&cnt[2 * Qpad + bidx[ps]], 1);
            }
        }
    }
    // ---- semantic tile ------------------------------------------------------------------------------------------
    const size_t plane = (size_t)Hc * Wc;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int ps = 2 * mt + (e >> 1);
                const int c = nt * 8 + 2 * t + (e & 1);
                if (pv[ps] && c < C) sem[c * plane + (size_t)y * Wc + tx + ps * 8 + g] = acc[mt][nt][e];
            }
    __syncthreads();
    for (int i = tid; i < 3 * Qpad; i += PP_THREADS) {
        const int v = cnt[i];
        const int a = i / Qpad, q = i - a * Qpad;
        if (v && q < Q) atomicAdd(&areas[a * Q + q], v);
    }
}

template <int NT>
int launch_seg_post(const float* masks, const void* pt_hi, const void* pt_lo, const float* scores, float* sem, int* ids,
                    int* areas, int Q, int Qpad, int C, int h, int w, int Hc, int Wc, cudaStream_t st) {
    const int smem = PP_QC * PP_TQ * 4 + 2 * NT * 8 * PP_PQ * 2 + PP_QC * 4 + 3 * Qpad * 4;
    HIPIE_CHECK_ARG(smem <= 220 * 1024, "hipie_seg_postprocess: Q=%d needs %d B of shared memory", Q, smem);
    static int smem_set = 0;
    if (smem > smem_set) {
        HIPIE_CHECK_CUDA(cudaFuncSetAttribute(seg_post_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        smem_set = smem;
    }
    dim3 grid((Wc + PP_TW - 1) / PP_TW, (Hc + PP_TH - 1) / PP_TH);
    seg_post_kernel<NT><<<grid, PP_THREADS, smem, st>>>(masks, (const __nv_bfloat16*)pt_hi, (const __nv_bfloat16*)pt_lo, scores,
                                                       sem, ids, areas, Q, Qpad, C, h, w, Hc, Wc);
    HIPIE_CHECK_LAUNCH();
    return HIPIE_OK;
}

}  // namespace
}  // namespace hipie

using namespace hipie;

extern "C" int hipie_seg_postprocess(const float* masks, const void* pt_hi, const void* pt_lo, const float* scores,
                                     float* sem, int* ids, int* areas, int Q, int Qpad, int C, int h, int w, int stride,
                                     int Hc, int Wc, void* stream) {
    HIPIE_CHECK_ARG(masks && pt_hi && pt_lo && scores && sem && ids && areas, "hipie_seg_postprocess: null pointer");
    HIPIE_CHECK_ARG(stride == 4, "hipie_seg_postprocess: only mask stride 4 is implemented (got %d)", stride);
    HIPIE_CHECK_ARG(Q > 0 && Qpad >= Q && Qpad % PP_QC == 0 && C > 0 && C <= 136 && h > 1 && w > 1 && Hc > 0 && Wc > 0 &&
                        Hc <= 4 * h && Wc <= 4 * w,
                    "hipie_seg_postprocess: bad sizes Q=%d Qpad=%d C=%d h=%d w=%d Hc=%d Wc=%d", Q, Qpad, C, h, w, Hc, Wc);
    HIPIE_CHECK_CUDA(cudaMemsetAsync(areas, 0, sizeof(int) * 3 * Q, (cudaStream_t)stream));
    if (C <= 80)
        return launch_seg_post<10>(masks, pt_hi, pt_lo, scores, sem, ids, areas, Q, Qpad, C, h, w, Hc, Wc, (cudaStream_t)stream);
    return launch_seg_post<17>(masks, pt_hi, pt_lo, scores, sem, ids, areas, Q, Qpad, C, h, w, Hc, Wc, (cudaStream_t)stream);
}
