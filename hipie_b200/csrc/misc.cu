// Memory-bound helper kernels of the hot path: bf16 splitting, patch extraction, NHWC im2col,
// pixel shuffle for the 2x2 transposed convolutions, 2x2 max-pool, masked row softmax for the
// VL-fusion scores, and the fused CondInst dynamic mask head.
#include "common.cuh"

namespace hipie {

// ---- fp32 -> bf16 hi/lo, optional second addend ---------------------------------------------
__global__ void __launch_bounds__(256)
split_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ sum_f32,
             __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<const float4*>(a)[i];
        if (b) {
            const float4 w = reinterpret_cast<const float4*>(b)[i];
            v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
        if (sum_f32) reinterpret_cast<float4*>(sum_f32)[i] = v;
        if (hi) {
            uint2 h, l;
            split2(v.x, v.y, h.x, l.x);
            split2(v.z, v.w, h.y, l.y);
            reinterpret_cast<uint2*>(hi)[i] = h;
            if (lo) reinterpret_cast<uint2*>(lo)[i] = l;
        }
    }
}

// ---- patch extraction (PatchEmbed conv k=16,s=16 as a GEMM; utils.py:160-186) fused with the
//      pixel normalisation (hipie_img.py:880-898).  img (B,3,H,W) raw fp32; rows (B*hp*wp, 3*P*P)
//      in (c, py, px) column order == conv weight.view(out, -1).
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                int B, int H, int W, int P, float m0, float m1, float m2, float s0, float s1, float s2) {
    const int hp = H / P, wp = W / P;
    const int K = 3 * P * P;
    const int64_t total = (int64_t)B * hp * wp * K / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i * 4;
        const int col = (int)(e % K);
        const int64_t row = e / K;
        const int px = col % P, py = (col / P) % P, c = col / (P * P);
        const int tx = (int)(row % wp), ty = (int)((row / wp) % hp), b = (int)(row / ((int64_t)wp * hp));
        const float4 v = *reinterpret_cast<const float4*>(img + (((int64_t)b * 3 + c) * H + ty * P + py) * W + tx * P + px);
        const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
        const float istd = 1.f / (c == 0 ? s0 : (c == 1 ? s1 : s2));
        uint2 h, l;
        split2((v.x - mean) * istd, (v.y - mean) * istd, h.x, l.x);
        split2((v.z - mean) * istd, (v.w - mean) * istd, h.y, l.y);
        *reinterpret_cast<uint2*>(hi + e) = h;
        if (lo) *reinterpret_cast<uint2*>(lo + e) = l;
    }
}

// ---- NHWC im2col for k x k convs (pad, stride): rows (B*Ho*Wo, k*k*C), column order (ky,kx,c)
__global__ void __launch_bounds__(256)
im2col_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                   int B, int H, int W, int C, int ksz, int stride, int pad, int Ho, int Wo) {
    const int K = ksz * ksz * C;
    const int64_t total = (int64_t)B * Ho * Wo * K / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i * 4;
        const int col = (int)(e % K);
        const int64_t row = e / K;
        const int c = col % C, kx = (col / C) % ksz, ky = col / (C * ksz);
        const int ox = (int)(row % Wo), oy = (int)((row / Wo) % Ho), b = (int)(row / ((int64_t)Wo * Ho));
        const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (iy >= 0 && iy < H && ix >= 0 && ix < W)
            v = *reinterpret_cast<const float4*>(x + (((int64_t)b * H + iy) * W + ix) * C + c);
        uint2 h, l;
        split2(v.x, v.y, h.x, l.x);
        split2(v.z, v.w, h.y, l.y);
        *reinterpret_cast<uint2*>(hi + e) = h;
        if (lo) *reinterpret_cast<uint2*>(lo + e) = l;
    }
}

// ---- ConvTranspose2d(k=2,s=2) epilogue: GEMM output rows (B*H*W, 4*C) with column order
//      (dy, dx, c) -> NHWC (B, 2H, 2W, C); optional bf16 split output.
__global__ void __launch_bounds__(256)
pixel_shuffle2_kernel(const float* __restrict__ g, float* __restrict__ y, __nv_bfloat16* __restrict__ hi,
                      __nv_bfloat16* __restrict__ lo, int B, int H, int W, int C) {
    const int64_t total = (int64_t)B * H * W * 4 * C / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i * 4;
        const int c = (int)(e % C);
        const int q = (int)((e / C) % 4);
        const int64_t row = e / (4 * (int64_t)C);
        const int x = (int)(row % W), yy = (int)((row / W) % H), b = (int)(row / ((int64_t)W * H));
        const int dy = q >> 1, dx = q & 1;
        const float4 v = *reinterpret_cast<const float4*>(g + e);
        const int64_t o = ((((int64_t)b * 2 * H + 2 * yy + dy) * 2 * W) + 2 * x + dx) * C + c;
        if (y) *reinterpret_cast<float4*>(y + o) = v;
        if (hi) {
            uint2 h, l;
            split2(v.x, v.y, h.x, l.x);
            split2(v.z, v.w, h.y, l.y);
            *reinterpret_cast<uint2*>(hi + o) = h;
            if (lo) *reinterpret_cast<uint2*>(lo + o) = l;
        }
    }
}

// F.max_pool2d(kernel 3, stride 2, padding 1) on NHWC (the ResNet stem, detectron2 resnet.py:362-365): padding never wins the max
__global__ void __launch_bounds__(256)
maxpool3x3s2_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, __nv_bfloat16* __restrict__ hi,
                         __nv_bfloat16* __restrict__ lo, int B, int H, int W, int C) {
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const int64_t total = (int64_t)B * Ho * Wo * C / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i * 4;
        const int c = (int)(e % C);
        const int64_t row = e / C;
        const int ox = (int)(row % Wo), oy = (int)((row / Wo) % Ho), b = (int)(row / ((int64_t)Wo * Ho));
        float4 v = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int iy = 2 * oy - 1 + dy;
            if (iy < 0 || iy >= H) continue;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int ix = 2 * ox - 1 + dx;
                if (ix < 0 || ix >= W) continue;
                const float4 a = *reinterpret_cast<const float4*>(x + (((int64_t)b * H + iy) * W + ix) * C + c);
                v.x = fmaxf(v.x, a.x); v.y = fmaxf(v.y, a.y); v.z = fmaxf(v.z, a.z); v.w = fmaxf(v.w, a.w);
            }
        }
        if (y) *reinterpret_cast<float4*>(y + e) = v;
        if (hi) {
            uint2 h, l;
            split2(v.x, v.y, h.x, l.x);
            split2(v.z, v.w, h.y, l.y);
            *reinterpret_cast<uint2*>(hi + e) = h;
            if (lo) *reinterpret_cast<uint2*>(lo + e) = l;
        }
    }
}

__global__ void __launch_bounds__(256)
maxpool2_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, __nv_bfloat16* __restrict__ hi,
                     __nv_bfloat16* __restrict__ lo, int B, int H, int W, int C) {
    const int Ho = H / 2, Wo = W / 2;
    const int64_t total = (int64_t)B * Ho * Wo * C / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i * 4;
        const int c = (int)(e % C);
        const int64_t row = e / C;
        const int ox = (int)(row % Wo), oy = (int)((row / Wo) % Ho), b = (int)(row / ((int64_t)Wo * Ho));
        const float* p = x + (((int64_t)b * H + 2 * oy) * W + 2 * ox) * C + c;
        const float4 a = *reinterpret_cast<const float4*>(p);
        const float4 bb = *reinterpret_cast<const float4*>(p + C);
        const float4 cc = *reinterpret_cast<const float4*>(p + (int64_t)W * C);
        const float4 d = *reinterpret_cast<const float4*>(p + (int64_t)W * C + C);
        float4 v;
        v.x = fmaxf(fmaxf(a.x, bb.x), fmaxf(cc.x, d.x));
        v.y = fmaxf(fmaxf(a.y, bb.y), fmaxf(cc.y, d.y));
        v.z = fmaxf(fmaxf(a.z, bb.z), fmaxf(cc.z, d.z));
        v.w = fmaxf(fmaxf(a.w, bb.w), fmaxf(cc.w, d.w));
        if (y) *reinterpret_cast<float4*>(y + e) = v;
        if (hi) {
            uint2 h, l;
            split2(v.x, v.y, h.x, l.x);
            split2(v.z, v.w, h.y, l.y);
            *reinterpret_cast<uint2*>(hi + e) = h;
            if (lo) *reinterpret_cast<uint2*>(lo + e) = l;
        }
    }
}

// ---- masked row softmax (BiMultiHeadAttention, fuse_helper.py:79-109) --------------------------
// x (rows, n) fp32 -> p = softmax(clamp(x - (sub_rowmax ? rowmax : 0), +-clampv) + colbias[b, :])
// written as bf16 hi/lo (operand of the following P.V GEMM).  One block per row.
__global__ void __launch_bounds__(256)
row_softmax_kernel(const float* __restrict__ x, const float* __restrict__ colbias, int64_t rows_per_batch,
                   __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, float* __restrict__ p_f32,
                   int n, float clampv, int sub_rowmax, int fp16) {
    __shared__ float red[8];
    __shared__ float bc;
    const int64_t row = blockIdx.x;
    const float* xr = x + row * n;
    const float* cb = colbias ? colbias + (row / rows_per_batch) * n : nullptr;
    auto block_reduce = [&](float v, bool is_max) {
        v = is_max ? warp_max(v) : warp_sum(v);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            float r = red[0];
            for (int i = 1; i < 8; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
            bc = r;
        }
        __syncthreads();
        const float r = bc;
        __syncthreads();
        return r;
    };
    float pre = 0.f;
    if (sub_rowmax) {
        float m = -INFINITY;
        for (int c = threadIdx.x; c < n; c += 256) m = fmaxf(m, fminf(fmaxf(xr[c], -clampv), clampv));
        pre = block_reduce(m, true);
    }
    auto val = [&](int c) {
        float v = fminf(fmaxf(xr[c], -clampv), clampv);
        if (sub_rowmax) v = fminf(fmaxf(v - pre, -clampv), clampv);
        if (cb) v += cb[c];
        return v;
    };
    float m = -INFINITY;
    for (int c = threadIdx.x; c < n; c += 256) m = fmaxf(m, val(c));
    m = block_reduce(m, true);
    float s = 0.f;
    for (int c = threadIdx.x; c < n; c += 256) s += expf(val(c) - m);
    s = block_reduce(s, false);
    const float inv = 1.f / s;
    for (int c = threadIdx.x; c < n; c += 256) {
        const float pv = expf(val(c) - m) * inv;
        if (p_f32) p_f32[row * n + c] = pv;
        if (hi) {
            if (fp16) {
                reinterpret_cast<__half*>(hi)[row * n + c] = __float2half_rn(pv);
            } else {
                const __nv_bfloat16 h = __float2bfloat16_rn(pv);
                hi[row * n + c] = h;
                if (lo) lo[row * n + c] = __float2bfloat16_rn(pv - __bfloat162float(h));
            }
        }
    }
}

// Vectorised variants (n % 4 == 0, 16-byte aligned rows).  Same arithmetic as row_softmax_kernel.
__device__ __forceinline__ float rs_val(float x, float pre, float cb, float clampv, int sub_rowmax) {
    float v = fminf(fmaxf(x, -clampv), clampv);
    if (sub_rowmax) v = fminf(fmaxf(v - pre, -clampv), clampv);
    return v + cb;
}
__device__ __forceinline__ void rs_store4(float4 pv, int64_t off, __nv_bfloat16* hi, __nv_bfloat16* lo, float* p_f32, int fp16) {
    if (p_f32) *reinterpret_cast<float4*>(p_f32 + off) = pv;
    if (hi) {
        uint2 h, l;
        split2m(pv.x, pv.y, h.x, l.x, fp16);
        split2m(pv.z, pv.w, h.y, l.y, fp16);
        *reinterpret_cast<uint2*>(hi + off) = h;
        if (lo) *reinterpret_cast<uint2*>(lo + off) = l;
    }
}

// short rows (n <= 1024): one warp per row, the row lives in registers (one global read), shuffle reductions only
template <int NCH>   // float4 chunks per lane
__global__ void __launch_bounds__(256)
row_softmax_warp_kernel(const float* __restrict__ x, const float* __restrict__ colbias, int64_t rows, int64_t rows_per_batch,
                        __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, float* __restrict__ p_f32, int n,
                        float clampv, int sub_rowmax, int fp16) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* xr = x + row * n;
    const float* cb = colbias ? colbias + (row / rows_per_batch) * n : nullptr;
    float4 v[NCH];
    float pre = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int c = (i * 32 + lane) * 4;
        v[i] = c < n ? *reinterpret_cast<const float4*>(xr + c) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
    if (sub_rowmax) {
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < NCH; ++i)
            if ((i * 32 + lane) * 4 < n)     // clamp is monotone: max of clamped == clamp of max
                m = fmaxf(m, fminf(fmaxf(fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)), -clampv), clampv));
        pre = warp_max(m);
    }
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < n) {
            const float4 b = cb ? *reinterpret_cast<const float4*>(cb + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[i].x = rs_val(v[i].x, pre, b.x, clampv, sub_rowmax);
            v[i].y = rs_val(v[i].y, pre, b.y, clampv, sub_rowmax);
            v[i].z = rs_val(v[i].z, pre, b.z, clampv, sub_rowmax);
            v[i].w = rs_val(v[i].w, pre, b.w, clampv, sub_rowmax);
            m = fmaxf(m, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
        }
    }
    m = warp_max(m);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        if ((i * 32 + lane) * 4 < n) {
            v[i].x = expf(v[i].x - m); v[i].y = expf(v[i].y - m); v[i].z = expf(v[i].z - m); v[i].w = expf(v[i].w - m);
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    s = warp_sum(s);
    const float inv = 1.f / s;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < n) rs_store4(make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv), row * n + c, hi, lo, p_f32, fp16);
    }
}

// long rows: one block per row, the row is cached in shared memory (one global read)
__global__ void __launch_bounds__(512)
row_softmax_smem_kernel(const float* __restrict__ x, const float* __restrict__ colbias, int64_t rows_per_batch,
                        __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, float* __restrict__ p_f32, int n,
                        float clampv, int sub_rowmax, int fp16) {
    extern __shared__ __align__(16) float rs_row[];
    __shared__ float red[16];
    __shared__ float bc;
    const int64_t row = blockIdx.x;
    const float* xr = x + row * n;
    const float* cb = colbias ? colbias + (row / rows_per_batch) * n : nullptr;
    const int n4 = n >> 2;
    auto block_reduce = [&](float v, bool is_max) {
        v = is_max ? warp_max(v) : warp_sum(v);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x < 32) {
            float r = threadIdx.x < 16 ? red[threadIdx.x] : (is_max ? -INFINITY : 0.f);
            r = is_max ? warp_max(r) : warp_sum(r);
            if (threadIdx.x == 0) bc = r;
        }
        __syncthreads();
        const float r = bc;
        __syncthreads();
        return r;
    };
    float m0 = -INFINITY;
    for (int c = threadIdx.x; c < n4; c += 512) {
        const float4 a = *reinterpret_cast<const float4*>(xr + 4 * c);
        *reinterpret_cast<float4*>(rs_row + 4 * c) = a;
        m0 = fmaxf(m0, fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)));
    }
    float pre = 0.f;
    if (sub_rowmax) pre = fminf(fmaxf(block_reduce(m0, true), -clampv), clampv);
    else __syncthreads();
    float m = -INFINITY;
    for (int c = threadIdx.x; c < n4; c += 512) {
        float4 a = *reinterpret_cast<const float4*>(rs_row + 4 * c);
        const float4 b = cb ? *reinterpret_cast<const float4*>(cb + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
        a.x = rs_val(a.x, pre, b.x, clampv, sub_rowmax); a.y = rs_val(a.y, pre, b.y, clampv, sub_rowmax);
        a.z = rs_val(a.z, pre, b.z, clampv, sub_rowmax); a.w = rs_val(a.w, pre, b.w, clampv, sub_rowmax);
        *reinterpret_cast<float4*>(rs_row + 4 * c) = a;
        m = fmaxf(m, fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)));
    }
    m = block_reduce(m, true);
    float s = 0.f;
    for (int c = threadIdx.x; c < n4; c += 512) {
        float4 a = *reinterpret_cast<const float4*>(rs_row + 4 * c);
        a.x = expf(a.x - m); a.y = expf(a.y - m); a.z = expf(a.z - m); a.w = expf(a.w - m);
        *reinterpret_cast<float4*>(rs_row + 4 * c) = a;
        s += (a.x + a.y) + (a.z + a.w);
    }
    s = block_reduce(s, false);
    const float inv = 1.f / s;
    for (int c = threadIdx.x; c < n4; c += 512) {
        const float4 a = *reinterpret_cast<const float4*>(rs_row + 4 * c);
        rs_store4(make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv), row * n + 4 * c, hi, lo, p_f32, fp16);
    }
}

// ---- sine embedding of box reference points (deformable_transformer_dino.py:636-670, maskdino/utils/utils.py:74-100) ----
// pos (rows, 4) = (x, y, w, h) with row stride `ld`;  out (rows, 512) = [emb(y) | emb(x) | emb(w) | emb(h)],
// emb(v)[2k] = sin(v*2pi / 10000^(2k/128)), emb(v)[2k+1] = cos(same).  One thread per (row, k, component) pair of outputs.
__global__ void __launch_bounds__(256)
sine_embed_kernel(const float* __restrict__ pos, int64_t ld, int64_t rows, float* __restrict__ out,
                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // over rows * 256 output pairs
    if (i >= rows * 256) return;
    const int64_t r = i >> 8;
    const int pair = (int)(i & 255), g = pair >> 6, k = pair & 63;
    const int comp = g == 0 ? 1 : (g == 1 ? 0 : g);                         // [y, x, w, h]
    const float v = pos[r * ld + comp] * 6.283185307179586f;
    const float dim_t = powf(10000.f, (float)(2 * k) / 128.f);
    const float a = v / dim_t;
    const float s_ = sinf(a), c_ = cosf(a);
    const int64_t o = r * 512 + g * 128 + 2 * k;
    if (out) *reinterpret_cast<float2*>(out + o) = make_float2(s_, c_);
    if (hi) {
        uint32_t h, l;
        split2(s_, c_, h, l);
        *reinterpret_cast<uint32_t*>(hi + o) = h;
        if (lo) *reinterpret_cast<uint32_t*>(lo + o) = l;
    }
}

// ---- CondInst dynamic mask head, fused (ddetrs_dn.py:1390-1502, 1806-1870) --------------------
// feats (B, Hf*Wf, 8) NHWC fp32; params (B, Q, 169) = [w0 (8x10) | w1 (8x8) | w2 (1x8) | b0 8 | b1 8 | b2 1];
// ref_px (B, Q, 2) reference point in pixels.  Per (b, q): 3-layer 1x1 MLP over
// [ref - (grid*stride + stride/2), feats] at stride-8 resolution, then aligned_bilinear(x2):
//   out[2i+a, 2j+b] with a,b in {0,1}: replicate-pad / align_corners arithmetic reduces to
//   out[Y, X] = lerp of coarse[(Y-1)/2 ...] — evaluated exactly as the reference's pad+interpolate+pad.
// One block per (q-chunk, b): coarse logits staged in shared memory (Hf*Wf <= 160*160).
__global__ void __launch_bounds__(256)
condinst_kernel(const float* __restrict__ feats, const float* __restrict__ params, const float* __restrict__ ref_px,
                float* __restrict__ out, int B, int Q, int Hf, int Wf, int stride) {
    extern __shared__ float coarse[];  // Hf*Wf
    const int q = blockIdx.x, b = blockIdx.y;
    // the 169 dynamic parameters, re-laid out for 16-byte broadcast loads: w0 rows padded 10 -> 12 (the MLP is shared-memory
    // bound: one LDS per weight and 4 pixels was 169 wavefronts per iteration, now 47)
    __shared__ __align__(16) float prm[8 * 12 + 64 + 8 + 8 + 8 + 4];
    __shared__ float ref[2];
    float* w0 = prm;                // [8][12]
    float* w1 = prm + 96;           // [8][8]
    float* w2 = prm + 160;          // [8]
    float* b0 = prm + 168;
    float* b1 = prm + 176;
    if (threadIdx.x < 169) {
        const int i = threadIdx.x;
        const float v = params[((int64_t)b * Q + q) * 169 + i];
        if (i < 80) w0[(i / 10) * 12 + i % 10] = v;
        else if (i < 144) w1[i - 80] = v;
        else if (i < 152) w2[i - 144] = v;
        else if (i < 160) b0[i - 152] = v;
        else if (i < 168) b1[i - 160] = v;
        else prm[184] = v;
    }
    if (threadIdx.x >= 192 && threadIdx.x < 208) w0[((threadIdx.x - 192) >> 1) * 12 + 10 + (threadIdx.x & 1)] = 0.f;
    if (threadIdx.x < 2) ref[threadIdx.x] = ref_px[((int64_t)b * Q + q) * 2 + threadIdx.x];
    __syncthreads();
    const float b2 = prm[184];
    const int HW = Hf * Wf;
    const float* fb = feats + (int64_t)b * HW * 8;
    // 4 coarse pixels per thread and iteration: every (broadcast) weight read from shared memory feeds 4 FMAs, otherwise the
    // 152 LDS per pixel make the MLP shared-memory bound (measured 3.1 ms for 8 x 910 queries; this layout ~3x faster)
    constexpr int PX = 4;
    // the 8 feature channels of the NEXT 4 pixels are fetched while the current ones go through the 152-MAC MLP: with one CTA per SM
    // (the unrolled MLP keeps the weights in registers) nothing else would hide the L2 latency of these loads
    float4 nf[PX][2];
    auto fetch = [&](int i0n) {
#pragma unroll
        for (int u = 0; u < PX; ++u) {
            const int i = i0n + u * blockDim.x;
            const int ii = i < HW ? i : 0;
            nf[u][0] = ldg_nc_f4(fb + (int64_t)ii * 8);
            nf[u][1] = ldg_nc_f4(fb + (int64_t)ii * 8 + 4);
        }
    };
    fetch(threadIdx.x);
    for (int i0 = threadIdx.x; i0 < HW; i0 += blockDim.x * PX) {
        float in[PX][10];
        bool ok[PX];
#pragma unroll
        for (int u = 0; u < PX; ++u) {
            const int i = i0 + u * blockDim.x;
            ok[u] = i < HW;
            const int ii = ok[u] ? i : 0;
            const int y = ii / Wf, x = ii - y * Wf;
            in[u][0] = ref[0] - (float)(x * stride + stride / 2);
            in[u][1] = ref[1] - (float)(y * stride + stride / 2);
            const float4 f0 = nf[u][0], f1 = nf[u][1];
            in[u][2] = f0.x; in[u][3] = f0.y; in[u][4] = f0.z; in[u][5] = f0.w;
            in[u][6] = f1.x; in[u][7] = f1.y; in[u][8] = f1.z; in[u][9] = f1.w;
        }
        if (i0 + blockDim.x * PX < HW) fetch(i0 + blockDim.x * PX);
        float h0[PX][8], h1[PX][8];
        const float4 bb0a = *reinterpret_cast<const float4*>(b0), bb0b = *reinterpret_cast<const float4*>(b0 + 4);
        const float bb0[8] = {bb0a.x, bb0a.y, bb0a.z, bb0a.w, bb0b.x, bb0b.y, bb0b.z, bb0b.w};
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            float a[PX];
#pragma unroll
            for (int u = 0; u < PX; ++u) a[u] = bb0[o];
            const float4 wa = *reinterpret_cast<const float4*>(w0 + o * 12), wb = *reinterpret_cast<const float4*>(w0 + o * 12 + 4);
            const float2 wc = *reinterpret_cast<const float2*>(w0 + o * 12 + 8);
            const float wr[10] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x, wc.y};
#pragma unroll
            for (int k = 0; k < 10; ++k) {          // same accumulation order as before: bias, then k = 0..9
#pragma unroll
                for (int u = 0; u < PX; ++u) a[u] += wr[k] * in[u][k];
            }
#pragma unroll
            for (int u = 0; u < PX; ++u) h0[u][o] = fmaxf(a[u], 0.f);
        }
        const float4 bb1a = *reinterpret_cast<const float4*>(b1), bb1b = *reinterpret_cast<const float4*>(b1 + 4);
        const float bb1[8] = {bb1a.x, bb1a.y, bb1a.z, bb1a.w, bb1b.x, bb1b.y, bb1b.z, bb1b.w};
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            float a[PX];
#pragma unroll
            for (int u = 0; u < PX; ++u) a[u] = bb1[o];
            const float4 wa = *reinterpret_cast<const float4*>(w1 + o * 8), wb = *reinterpret_cast<const float4*>(w1 + o * 8 + 4);
            const float wr[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
#pragma unroll
                for (int u = 0; u < PX; ++u) a[u] += wr[k] * h0[u][k];
            }
#pragma unroll
            for (int u = 0; u < PX; ++u) h1[u][o] = fmaxf(a[u], 0.f);
        }
        float a[PX];
#pragma unroll
        for (int u = 0; u < PX; ++u) a[u] = b2;
        {
            const float4 wa = *reinterpret_cast<const float4*>(w2), wb = *reinterpret_cast<const float4*>(w2 + 4);
            const float wr[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
#pragma unroll
                for (int u = 0; u < PX; ++u) a[u] += wr[k] * h1[u][k];
            }
        }
#pragma unroll
        for (int u = 0; u < PX; ++u)
            if (ok[u]) coarse[i0 + u * blockDim.x] = a[u];
    }
    __syncthreads();
    // aligned_bilinear(factor 2): pad right/bottom by replicate -> (Hf+1, Wf+1); interpolate with
    // align_corners=True to (2Hf+1, 2Wf+1): sample position = Y/2 exactly; pad left/top by 1
    // (replicate) and crop to (2Hf, 2Wf): out[Y, X] = up[max(Y-1,0), max(X-1,0)].
    const int Ho = 2 * Hf, Wo = 2 * Wf;
    float* ob = out + ((int64_t)b * Q + q) * Ho * Wo;
    auto at = [&](int yy, int xx) { return coarse[min(yy, Hf - 1) * Wf + min(xx, Wf - 1)]; };
    auto px = [&](int Y, int X) {
        const int uy = max(Y - 1, 0), ux = max(X - 1, 0);
        const int y0 = uy >> 1, x0 = ux >> 1;
        const float fy = (uy & 1) ? 0.5f : 0.f, fx = (ux & 1) ? 0.5f : 0.f;
        const int y1 = y0 + 1, x1 = x0 + 1;
        const float v00 = at(y0, x0), v01 = at(y0, x1), v10 = at(y1, x0), v11 = at(y1, x1);
        const float top = v00 + (v01 - v00) * fx, bot = v10 + (v11 - v10) * fx;
        return top + (bot - top) * fy;
    };
    if ((Wo & 3) == 0) {          // 4 outputs per thread, 16-byte stores (rows are 16-byte aligned: Wo % 4 == 0)
        const int n4 = Ho * Wo / 4;
        for (int i = threadIdx.x; i < n4; i += blockDim.x) {
            const int Y = (i * 4) / Wo, X = i * 4 - Y * Wo;
            if (X == 0) {       // left border: the generic expression (clamped source columns)
                *reinterpret_cast<float4*>(ob + (int64_t)i * 4) = make_float4(px(Y, 0), px(Y, 1), px(Y, 2), px(Y, 3));
                continue;
            }
            // X = 4k > 0: source columns ux = X-1 .. X+2 touch coarse columns c, c+1, c+2 with c = X/2 - 1 (odd ux: midpoint of two
            // columns, even ux: one column) -- 6 shared-memory loads for 4 outputs instead of 16, the same arithmetic per output as px()
            const int uy = max(Y - 1, 0), y0 = uy >> 1, c = (X >> 1) - 1;
            const float fy = (uy & 1) ? 0.5f : 0.f;
            const float t0 = at(y0, c), t1 = at(y0, c + 1), t2 = at(y0, c + 2);
            const float u0 = at(y0 + 1, c), u1 = at(y0 + 1, c + 1), u2 = at(y0 + 1, c + 2);
            const float top0 = t0 + (t1 - t0) * 0.5f, bot0 = u0 + (u1 - u0) * 0.5f;
            const float top1 = t1, bot1 = u1;                            // even ux: v00 + (v01 - v00) * 0 == v00 for finite logits
            const float top2 = t1 + (t2 - t1) * 0.5f, bot2 = u1 + (u2 - u1) * 0.5f;
            const float top3 = t2, bot3 = u2;
            *reinterpret_cast<float4*>(ob + (int64_t)i * 4) = make_float4(top0 + (bot0 - top0) * fy, top1 + (bot1 - top1) * fy,
                                                                          top2 + (bot2 - top2) * fy, top3 + (bot3 - top3) * fy);
        }
    } else {
        for (int i = threadIdx.x; i < Ho * Wo; i += blockDim.x) {
            const int Y = i / Wo, X = i - Y * Wo;
            ob[i] = px(Y, X);
        }
    }
}

static inline int grid_for(int64_t n, int threads = 256) {
    int64_t b = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)num_sms() * 16;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace hipie

using namespace hipie;

// fp32 (rows, cols) -> fp16 plane + e4m3 planes of a prec-6 GEMM operand (gemm_tc.cu); thread = 4 consecutive columns.
// Activations: slot 0 = e4m3(h), slot 1 = e4m3(2^10 (x - h)); weights: slot 0 = e4m3(2^14 (w - h)), slot 1 = e4m3(2^4 h)
__global__ void __launch_bounds__(256)
split_f16_e4m3_kernel(const float* __restrict__ x, __half* __restrict__ h16, uint8_t* __restrict__ p8, int64_t rows, int cols, int weight) {
    const int c4 = cols >> 2;
    const int64_t total = rows * c4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / c4;
        const int c = (int)(i - r * c4) * 4;
        const float4 v = *reinterpret_cast<const float4*>(x + r * cols + c);
        const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
        uint2 h;
        h.x = *reinterpret_cast<const uint32_t*>(&h0);
        h.y = *reinterpret_cast<const uint32_t*>(&h1);
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        uint32_t first, second;
        if (weight) {      // [e4m3(2^14 (w - h)) | e4m3(2^4 h)]
            first = pack_e4m3x4((v.x - f0.x) * 16384.f, (v.y - f0.y) * 16384.f, (v.z - f1.x) * 16384.f, (v.w - f1.y) * 16384.f);
            second = pack_e4m3x4(f0.x * 16.f, f0.y * 16.f, f1.x * 16.f, f1.y * 16.f);
        } else {           // [e4m3(h) | e4m3(2^10 (a - h))]
            first = pack_e4m3x4(f0.x, f0.y, f1.x, f1.y);
            second = pack_e4m3x4((v.x - f0.x) * 1024.f, (v.y - f0.y) * 1024.f, (v.z - f1.x) * 1024.f, (v.w - f1.y) * 1024.f);
        }
        *reinterpret_cast<uint2*>(h16 + r * cols + c) = h;
        uint8_t* o8 = p8 + r * 2 * cols + e4m3_slot0(c);          // the two planes interleaved in 32-column groups (common.cuh)
        *reinterpret_cast<uint32_t*>(o8) = first;
        *reinterpret_cast<uint32_t*>(o8 + 32) = second;
    }
}

extern "C" int hipie_split_f16_e4m3(const float* x, void* h16, void* p8, int64_t rows, int cols, int weight, void* stream) {
    HIPIE_CHECK_ARG(x && h16 && p8, "hipie_split_f16_e4m3: null pointer");
    HIPIE_CHECK_ARG(rows >= 0 && cols > 0 && cols % 32 == 0, "hipie_split_f16_e4m3: cols must be a positive multiple of 32");
    if (rows == 0) return HIPIE_OK;
    const int64_t total = rows * (cols / 4);
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)num_sms() * 16);
    split_f16_e4m3_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, (__half*)h16, (uint8_t*)p8, rows, cols, weight);
    HIPIE_CHECK_LAUNCH();
    return HIPIE_OK;
}

extern "C" int hipie_split_bf16(const float* x, void* hi, void* lo, int64_t n, void* stream) {
    HIPIE_CHECK_ARG(x && hi, "hipie_split_bf16: null pointer");
    HIPIE_CHECK_ARG(n % 4 == 0, "hipie_split_bf16: n must be a multiple of 4");
    if (n == 0) return HIPIE_OK;
    split_kernel<<<grid_for(n / 4), 256, 0, (cudaStream_t)stream>>>(x, nullptr, nullptr, (__nv_bfloat16*)hi,
                                                                   (__nv_bfloat16*)lo, n / 4);
    HIPIE_CHECK_LAUNCH();
    return HIPIE_OK;
}

extern "C" int hipie_add_split(const float* a, const float* b, float* sum_f32, void* hi, void* lo, int64_t n,
                               void* stream) {
    HIPIE_CHECK_ARG(a && (sum_f32 || hi), "hipie_add_split: null pointer");
    HIPIE_CHECK_ARG(n % 4 == 0, "hipie_add_split: n must be a multiple of 4");
    if (n == 0) return HIPIE_OK;
    split_kernel<<<grid_for(n / 4), 256, 0, (cudaStream_t)stream>>>(a, b, sum_f32, (__nv_bfloat16*)hi,
                                                                   (__nv_bfloat16*)lo, n / 4);
    HIPIE_CHECK_LAUNCH();
    return HIPIE_OK;
}

extern "C" int hipie_patchify(const float* img, void* hi, void* lo, int B, int H, int W, int P, const float* mean3,
                              const float* std3, void* stream) {
    HIPIE_CHECK_ARG(img && hi && mean3 && std3, "hipie_patchify: null pointer");
    HIPIE_CHECK_ARG(P % 4 == 0 && H % P == 0 && W % P == 0, "hipie_patchify: H, W must be multiples of P, P of 4");
    const int64_t total = (int64_t)B * (H / P) * (W / P) * 3 * P * P / 4;
    patchify_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(img, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, B, H,
                                                                      W, P, mean3[0], mean3[1], mean3[2], std3[0],
                                                                      std3[1], std3[2]);
    HIPIE_CHECK_LAUNCH();
    return HIPIE_OK;
}

extern "C" int hipie_im2col_nhwc(const float* x, void* hi, void* lo, int B, int H, int W, int C, int ksz, int stride,
                                 int pad, void* stream) {
    HIPIE_CHECK_ARG(x && hi, "hipie_im2col_nhwc: null pointer");
    HIPIE_CHECK_ARG(C % 4 == 0, "hipie_im2col_nhwc: C must be a multiple of 4");
    const int Ho = (H + 2 * pad - ksz) / stride + 1, Wo = (W + 2 * pad - ksz) / stride + 1;
    const int64_t total = (int64_t)B * Ho * Wo * ksz * ksz * C / 4;
    im2col_nhwc_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, B,
                                                                         H, W, C, ksz, stride, pad, Ho, Wo);
    HIPIE_CHECK_LAUNCH();
    return HIPIE_OK;
}

extern "C" int hipie_pixel_shuffle2(const float* g, float* y, void* hi, void* lo, int B, int H, int W, int C,
                                    void* stream) {
    HIPIE_CHECK_ARG(g && (y || hi), "hipie_pixel_shuffle2: null pointer");
    HIPIE_CHECK_ARG(C % 4 == 0, "hipie_pixel_shuffle2: C must be a multiple of 4");
    const int64_t total = (int64_t)B * H * W * C;
    pixel_shuffle2_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(g, y, (__nv_bfloat16*)hi,
                                                                            (__nv_bfloat16*)lo, B, H, W, C);
    HIPIE_CHECK_LAUNCH();
    return HIPIE_OK;
}

extern "C" int hipie_maxpool2_nhwc(const float* x, float* y, void* hi, void* lo, int B, int H, int W, int C,
                                   void* stream) {
    HIPIE_CHECK_ARG(x && (y || hi), "hipie_maxpool2_nhwc: null pointer");
    HIPIE_CHECK_ARG(C % 4 == 0 && H % 2 == 0 && W % 2 == 0, "hipie_maxpool2_nhwc: bad sizes");
    const int64_t total = (int64_t)B * (H / 2) * (W / 2) * C / 4;
    maxpool2_nhwc_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, y, (__nv_bfloat16*)hi,
                                                                           (__nv_bfloat16*)lo, B, H, W, C);
    HIPIE_CHECK_LAUNCH();
    return HIPIE_OK;
}

extern "C" int hipie_maxpool3x3s2_nhwc(const float* x, float* y, void* hi, void* lo, int B, int H, int W, int C, void* stream) {
    HIPIE_CHECK_ARG(x && (y || hi), "hipie_maxpool3x3s2_nhwc: null pointer");
    HIPIE_CHECK_ARG(C % 4 == 0 && H > 0 && W > 0, "hipie_maxpool3x3s2_nhwc: bad sizes");
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const int64_t total = (int64_t)B * Ho * Wo * C / 4;
    maxpool3x3s2_nhwc_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, y, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, B, H, W, C);
    HIPIE_CHECK_LAUNCH();
    return HIPIE_OK;
}

extern "C" int hipie_row_softmax(const float* x, const float* colbias, int64_t rows, int64_t rows_per_batch, int n,
                                 float clampv, int sub_rowmax, void* hi, void* lo, float* p_f32, int out_fp16, void* stream) {
    const int fp16 = out_fp16 ? 1 : 0;
    HIPIE_CHECK_ARG(!fp16 || (hi && !lo), "hipie_row_softmax: out_fp16 writes one fp16 plane (hi set, lo NULL)");
    HIPIE_CHECK_ARG(x && (hi || p_f32), "hipie_row_softmax: null pointer");
    HIPIE_CHECK_ARG(rows >= 0 && n > 0 && rows_per_batch > 0, "hipie_row_softmax: bad sizes");
    if (rows == 0) return HIPIE_OK;
    const bool aligned = n % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(colbias) |
                                        reinterpret_cast<uintptr_t>(p_f32)) & 15) == 0 &&
                         ((reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 7) == 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (aligned && n <= 1024) {
        const unsigned blocks = (unsigned)((rows + 7) / 8);
        if (n <= 512)
            row_softmax_warp_kernel<4><<<blocks, 256, 0, st>>>(x, colbias, rows, rows_per_batch, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo,
                                                               p_f32, n, clampv, sub_rowmax, fp16);
        else
            row_softmax_warp_kernel<8><<<blocks, 256, 0, st>>>(x, colbias, rows, rows_per_batch, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo,
                                                               p_f32, n, clampv, sub_rowmax, fp16);
    } else if (aligned && (int64_t)n * 4 <= 200 * 1024) {
        const int smem = n * 4;
        HIPIE_ENSURE_SMEM(row_softmax_smem_kernel, smem);
        row_softmax_smem_kernel<<<(unsigned)rows, 512, smem, st>>>(x, colbias, rows_per_batch, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo,
                                                                   p_f32, n, clampv, sub_rowmax, fp16);
    } else {
        row_softmax_kernel<<<(unsigned)rows, 256, 0, st>>>(x, colbias, rows_per_batch, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, p_f32,
                                                           n, clampv, sub_rowmax, fp16);
    }
    HIPIE_CHECK_LAUNCH();
    return HIPIE_OK;
}

extern "C" int hipie_condinst_masks(const float* feats, const float* params, const float* ref_px, float* out, int B,
                                    int Q, int Hf, int Wf, int stride, void* stream) {
    HIPIE_CHECK_ARG(feats && params && ref_px && out, "hipie_condinst_masks: null pointer");
    HIPIE_CHECK_ARG(B > 0 && Q > 0 && Hf > 0 && Wf > 0 && (int64_t)Hf * Wf * 4 <= 200 * 1024,
                    "hipie_condinst_masks: bad sizes B=%d Q=%d Hf=%d Wf=%d", B, Q, Hf, Wf);
    const int smem = Hf * Wf * (int)sizeof(float);
    HIPIE_ENSURE_SMEM(condinst_kernel, smem);
    condinst_kernel<<<dim3(Q, B), 256, smem, (cudaStream_t)stream>>>(feats, params, ref_px, out, B, Q, Hf, Wf, stride);
    HIPIE_CHECK_LAUNCH();
    return HIPIE_OK;
}

extern "C" int hipie_sine_embed(const float* pos, int64_t ld, int64_t rows, float* out, void* hi, void* lo, void* stream) {
    if (rows == 0) return HIPIE_OK;
    HIPIE_CHECK_ARG(pos && (out || hi) && rows > 0 && ld >= 4, "hipie_sine_embed: bad arguments");
    const int64_t total = rows * 256;
    sine_embed_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pos, ld, rows, out, (__nv_bfloat16*)hi,
                                                                                        (__nv_bfloat16*)lo);
    HIPIE_CHECK_LAUNCH();
    return HIPIE_OK;
}
