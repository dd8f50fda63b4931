// MaskCLIP re-scoring support kernels: everything around the ViT-L/14 forward that the reference does with
// torch ops on full-resolution tensors -- projects/HIPIE/hipie/open_vocab/clip.py:288-349 (image / mask resize to the CLIP input,
// max-pooled patch masks -> boolean attention mask), :351-361 + open_vocab/helper.py:79-109 (normalised dot products, logit scale,
// max over prompt synonyms) and hipie_img.py:592-609,735-747,840-866 (probability fusion with the model's own class scores).
// The transformer itself runs on hipie_gemm / hipie_layernorm / hipie_attention (key_mask).
#include "common.cuh"

namespace hipie {

// torch upsample_bilinear2d, align_corners = False: src = scale * (dst + 0.5) - 0.5 clamped at 0; second tap = +1 unless at the border
struct Tap { int i0, i1; float l0, l1; };
__device__ __forceinline__ Tap bilinear_tap(float scale, int dst, int in_size) {
    float s = scale * ((float)dst + 0.5f) - 0.5f;
    s = s < 0.f ? 0.f : s;
    Tap t;
    t.i0 = min((int)s, in_size - 1);
    t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
    t.l1 = s - (float)t.i0;
    t.l0 = 1.f - t.l1;
    return t;
}
__device__ __forceinline__ float bilinear_at(const float* img, int w, const Tap& ty, const Tap& tx) {
    const float* r0 = img + (int64_t)ty.i0 * w;
    const float* r1 = img + (int64_t)ty.i1 * w;
    return ty.l0 * (tx.l0 * __ldg(r0 + tx.i0) + tx.l1 * __ldg(r0 + tx.i1)) + ty.l1 * (tx.l0 * __ldg(r1 + tx.i0) + tx.l1 * __ldg(r1 + tx.i1));
}

// One warp per (query, patch): max over the P x P samples of the mask logits resized to S x S.  UP > 1: the source of that resize
// is itself the x UP bilinear upsample of the stored map cropped to (Hc, Wc) -- hipie_img.py:733-734 feeds MaskCLIP the upsampled,
// cropped masks -- evaluated on the fly (16 taps) instead of materialising (Q, Hc, Wc).
__global__ void __launch_bounds__(256)
patch_mask_kernel(const float* __restrict__ masks, int Q, int h, int w, int up, int Hc, int Wc, int S, int P, uint32_t* __restrict__ bits,
                  int row_words, int key_offset) {
    const int G = S / P;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (gw >= Q * G * G) return;
    const int q = gw / (G * G), pidx = gw - q * G * G, py = pidx / G, px = pidx - py * G;
    const float* img = masks + (int64_t)q * h * w;
    const float sy = (float)Hc / (float)S, sx = (float)Wc / (float)S;       // scale of the resize to the CLIP input
    const float inner = 1.f / (float)up;                                      // scale of the x UP upsample: in / out, exact for UP = 4
    float mx = -INFINITY;
    for (int s = lane; s < P * P; s += 32) {
        const int iy = s / P, ix = s - iy * P;
        const Tap ty = bilinear_tap(sy, py * P + iy, Hc), tx = bilinear_tap(sx, px * P + ix, Wc);
        float v;
        if (up == 1) {
            v = bilinear_at(img, w, ty, tx);
        } else {
            const Tap ya = bilinear_tap(inner, ty.i0, h), yb = bilinear_tap(inner, ty.i1, h);
            const Tap xa = bilinear_tap(inner, tx.i0, w), xb = bilinear_tap(inner, tx.i1, w);
            const float v00 = bilinear_at(img, w, ya, xa), v01 = bilinear_at(img, w, ya, xb);
            const float v10 = bilinear_at(img, w, yb, xa), v11 = bilinear_at(img, w, yb, xb);
            v = ty.l0 * (tx.l0 * v00 + tx.l1 * v01) + ty.l1 * (tx.l0 * v10 + tx.l1 * v11);
        }
        mx = fmaxf(mx, v);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) {
        // clip.py:302-309: patch_mask = max_pool2d(sigmoid(mask)); masked out where patch_mask < 0.5 (sigmoid is monotonic: pool the logits)
        const float sg = 1.f / (1.f + expf(-mx));
        if (sg < 0.5f) {
            const int key = key_offset + pidx;
            atomicOr(bits + (int64_t)q * row_words + (key >> 5), 1u << (key & 31));
        }
    }
}

// CLIP input patches: bilinear resize of the (3, H, W) image in 0..1 to S x S (clip.py:336-341), OpenAI normalisation (:291), and
// the im2col of the stride-P patch convolution: row = patch, column = c * P * P + iy * P + ix (Conv2d weight order), bf16 hi/lo.
__global__ void __launch_bounds__(256)
clip_patches_kernel(const float* __restrict__ image, int H, int W, int S, int P, float3 mean, float3 inv_std, __nv_bfloat16* __restrict__ hi,
                    __nv_bfloat16* __restrict__ lo, int ld) {
    const int G = S / P, K = 3 * P * P;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)G * G * K) return;
    const int row = (int)(idx / K), col = (int)(idx - (int64_t)row * K);
    const int c = col / (P * P), r = col - c * P * P, iy = r / P, ix = r - iy * P;
    const int py = row / G, px = row - py * G;
    const Tap ty = bilinear_tap((float)H / (float)S, py * P + iy, H), tx = bilinear_tap((float)W / (float)S, px * P + ix, W);
    float v = bilinear_at(image + (int64_t)c * H * W, W, ty, tx);
    const float m = c == 0 ? mean.x : (c == 1 ? mean.y : mean.z), is = c == 0 ? inv_std.x : (c == 1 ? inv_std.y : inv_std.z);
    v = (v - m) * is;
    const __nv_bfloat16 h16 = __float2bfloat16_rn(v);
    hi[(int64_t)row * ld + col] = h16;
    if (lo) lo[(int64_t)row * ld + col] = __float2bfloat16_rn(v - __bfloat162float(h16));
}

// One CTA per query row.  CLIP side: logits_n = raw_n / max(|mask_embed|, 1e-12) * logit_scale (text rows are unit vectors already),
// class logit = max over the class's prompts, probability = softmax over the classes (sigmoid for one class).  Model side: p =
// softmax(sigmoid(score) / temp) (temp > 0) or sigmoid(score).  Fusion (hipie_img.py:845-866): log(p^(1-w) * c^w) or
// log(p (1-w) + c w + 1e-9), w = alpha on classes seen in training and beta on the others.
//   mode 0: the fused log-probabilities            mode 2: softmax of them over the classes (hipie_img.py:747)
//   mode 1: sqrt(sigmoid(fused)^a * sigmoid(iou)^b) * [score(row 0) != -9999] (:592-609) plus its row max / first argmax (NMS keys)
__global__ void __launch_bounds__(128)
clip_fuse_kernel(const float* __restrict__ raw, int ld_raw, const float* __restrict__ mask_embed, int D, float logit_scale,
                 const int* __restrict__ seg, const float* __restrict__ scores, float temp, const int8_t* __restrict__ overlap, float alpha,
                 float beta, int agg_add, const float* __restrict__ iou, float fg_a, float fg_b, int mode, float* __restrict__ out,
                 float* __restrict__ row_max, int* __restrict__ row_arg, int C) {
    extern __shared__ float fs[];
    float* cl = fs;            // [C] clip logits -> clip probabilities
    float* pm = fs + C;        // [C] model probabilities
    __shared__ float red[4];
    __shared__ int redi[4];
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    auto block_sum = [&](float v) {
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if (lane == 0) red[wid] = v;
        __syncthreads();
        return red[0] + red[1] + red[2] + red[3];
    };
    auto block_max = [&](float v) {
#pragma unroll
        for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        __syncthreads();
        if (lane == 0) red[wid] = v;
        __syncthreads();
        return fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    };
    float ss = 0.f;
    for (int i = tid; i < D; i += 128) { const float v = mask_embed[(int64_t)r * D + i]; ss += v * v; }
    const float inv_norm = 1.f / fmaxf(sqrtf(block_sum(ss)), 1e-12f);
    const float* rr = raw + (int64_t)r * ld_raw;
    const float* sr = scores + (int64_t)r * C;
    float lmax = -INFINITY, smax = -INFINITY;
    for (int c = tid; c < C; c += 128) {
        float m = -INFINITY;
        for (int n = seg[c]; n < seg[c + 1]; ++n) m = fmaxf(m, rr[n] * inv_norm * logit_scale);
        cl[c] = m;
        lmax = fmaxf(lmax, m);
        const float sg = 1.f / (1.f + expf(-sr[c]));
        const float z = temp > 0.f ? sg / temp : sg;
        pm[c] = z;
        smax = fmaxf(smax, z);
    }
    lmax = block_max(lmax);
    smax = block_max(smax);
    float lsum = 0.f, ssum = 0.f;
    for (int c = tid; c < C; c += 128) {
        if (C > 1) { cl[c] = expf(cl[c] - lmax); lsum += cl[c]; }
        else cl[c] = 1.f / (1.f + expf(-cl[c]));
        if (temp > 0.f) { pm[c] = expf(pm[c] - smax); ssum += pm[c]; }
    }
    lsum = block_sum(lsum);
    ssum = block_sum(ssum);
    float fmax_ = -INFINITY;
    for (int c = tid; c < C; c += 128) {
        const float pc = C > 1 ? cl[c] / lsum : cl[c];
        const float p = temp > 0.f ? pm[c] / ssum : pm[c];
        const float wgt = overlap[c] ? alpha : beta;
        const float f = agg_add ? logf(p * (1.f - wgt) + pc * wgt + 1e-9f) : logf(powf(p, 1.f - wgt) * powf(pc, wgt));
        cl[c] = f;
        fmax_ = fmaxf(fmax_, f);
    }
    if (mode == 0) {
        __syncthreads();
        for (int c = tid; c < C; c += 128) out[(int64_t)r * C + c] = cl[c];
        return;
    }
    if (mode == 2) {
        fmax_ = block_max(fmax_);
        float fsum = 0.f;
        for (int c = tid; c < C; c += 128) { pm[c] = expf(cl[c] - fmax_); fsum += pm[c]; }
        fsum = block_sum(fsum);
        for (int c = tid; c < C; c += 128) out[(int64_t)r * C + c] = pm[c] / fsum;
        return;
    }
    // mode 1
    const float si = iou ? 1.f / (1.f + expf(-iou[r])) : 1.f;
    const float ib = iou ? powf(si, fg_b) : 1.f;
    float best = -INFINITY;
    int besti = 0x7fffffff;
    for (int c = tid; c < C; c += 128) {
        const float thing = scores[c] == -9999.0f ? 0.f : 1.f;            // row 0 of the score matrix (hipie_img.py:593)
        float pr = (1.f / (1.f + expf(-cl[c]))) * thing;
        if (iou) pr = sqrtf(powf(pr, fg_a) * ib);
        out[(int64_t)r * C + c] = pr;
        if (pr > best) { best = pr; besti = c; }
    }
    if (row_max) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
        __syncthreads();
        if (lane == 0) { red[wid] = best; redi[wid] = besti; }
        __syncthreads();
        if (tid == 0) {
            for (int k = 1; k < 4; ++k)
                if (red[k] > best || (red[k] == best && redi[k] < besti)) { best = red[k]; besti = redi[k]; }
            row_max[r] = best;
            row_arg[r] = besti == 0x7fffffff ? 0 : besti;
        }
    }
}

}  // namespace hipie

using namespace hipie;

extern "C" int hipie_maskclip_patch_mask(const float* masks, int Q, int h, int w, int up, int Hc, int Wc, int S, int P, uint32_t* bits,
                                         int row_words, int key_offset, void* stream) {
    HIPIE_CHECK_ARG(masks && bits, "hipie_maskclip_patch_mask: null pointer");
    HIPIE_CHECK_ARG(Q >= 0 && h > 0 && w > 0 && (up == 1 || up == 4) && S > 0 && P > 0 && S % P == 0, "hipie_maskclip_patch_mask: bad sizes");
    HIPIE_CHECK_ARG(up == 1 ? (Hc == h && Wc == w) : (Hc > 0 && Wc > 0 && Hc <= up * h && Wc <= up * w),
                    "hipie_maskclip_patch_mask: crop (%d, %d) outside the x%d map of (%d, %d)", Hc, Wc, up, h, w);
    const int G = S / P;
    HIPIE_CHECK_ARG(row_words * 32 >= key_offset + G * G, "hipie_maskclip_patch_mask: row_words too small");
    if (Q == 0) return HIPIE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(bits, 0, (size_t)Q * row_words * sizeof(uint32_t), st) != cudaSuccess) {
        set_error("hipie_maskclip_patch_mask: memset failed");
        return HIPIE_ECUDA;
    }
    const int64_t warps = (int64_t)Q * G * G;
    patch_mask_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(masks, Q, h, w, up, Hc, Wc, S, P, bits, row_words, key_offset);
    HIPIE_CHECK_LAUNCH();
    return HIPIE_OK;
}

extern "C" int hipie_clip_patches(const float* image, int H, int W, int S, int P, const float* mean3, const float* std3, void* hi, void* lo,
                                  int ld, void* stream) {
    HIPIE_CHECK_ARG(image && hi && mean3 && std3, "hipie_clip_patches: null pointer");
    HIPIE_CHECK_ARG(H > 0 && W > 0 && S > 0 && P > 0 && S % P == 0 && ld >= 3 * P * P, "hipie_clip_patches: bad sizes");
    const int G = S / P;
    const int64_t n = (int64_t)G * G * 3 * P * P;
    clip_patches_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        image, H, W, S, P, make_float3(mean3[0], mean3[1], mean3[2]), make_float3(1.f / std3[0], 1.f / std3[1], 1.f / std3[2]),
        (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, ld);
    HIPIE_CHECK_LAUNCH();
    return HIPIE_OK;
}

extern "C" int hipie_clip_fuse(const float* raw, int ld_raw, const float* mask_embed, int D, float logit_scale, const int* seg,
                               const float* scores, float temp, const int8_t* overlap, float alpha, float beta, int agg_add,
                               const float* iou, float fg_a, float fg_b, int mode, float* out, float* row_max, int* row_arg, int R, int C,
                               void* stream) {
    HIPIE_CHECK_ARG(raw && mask_embed && seg && scores && overlap && out, "hipie_clip_fuse: null pointer");
    HIPIE_CHECK_ARG(mode >= 0 && mode <= 2 && R >= 0 && C > 0 && C <= 8192 && D > 0, "hipie_clip_fuse: bad arguments");
    HIPIE_CHECK_ARG((row_max == nullptr) == (row_arg == nullptr), "hipie_clip_fuse: row_max and row_arg go together");
    if (R == 0) return HIPIE_OK;
    const int smem = 2 * C * (int)sizeof(float);
    HIPIE_ENSURE_SMEM(clip_fuse_kernel, smem);
    clip_fuse_kernel<<<R, 128, smem, (cudaStream_t)stream>>>(raw, ld_raw, mask_embed, D, logit_scale, seg, scores, temp, overlap, alpha, beta,
                                                            agg_add, iou, fg_a, fg_b, mode, out, row_max, row_arg, C);
    HIPIE_CHECK_LAUNCH();
    return HIPIE_OK;
}
