// shade.cu — the screen-space passes: deferred lighting (forkergl.cpp:326-380) fused with the shadow filters
// (shadow.cpp:23-132) and the 8-bit quantise (buffer.cpp:113-126); SSAO (render.cpp:214-286); the in-place
// Gaussian blur as the recurrence it is (buffer.cpp:59-98); the SSAA box resolve (render.cpp:291-343); clears and
// layout conversions for the host accessors.
#include <cuda_pipeline.h>

#include <cmath>
#include <cstdlib>

#include "fgl_internal.h"

namespace
{
__device__ __forceinline__ V3 ld3s(const float* plane, size_t n, size_t idx) { return v3(plane[idx], plane[n + idx], plane[2 * n + idx]); }

__device__ __forceinline__ uint8_t quant8(float v) { return (uint8_t)f2i_x86(v * 254.99f); }  // buffer.cpp:121-122

// ---- clears ------------------------------------------------------------------------------------------------------
__global__ void k_fill(float* dst, size_t n, float value)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = value;
}

// ---- lighting, one thread per pixel (HardShadow or shadows off) ----------------------------------------------------
struct PixelIn
{
    V3    pos, nrm, lndc, albedo, emissive, param;
    float type;
};

__device__ __forceinline__ V3 shade_pixel(const LightPass& L, const PixelIn& in, float visibility, V3 lightDir, V3 viewDir)
{
    LightConsts lc;
    lc.shadowIntensity = L.shadowIntensity, lc.shadowOn = L.shadowOn;
    V3 rad = v3(L.lightColor[0], L.lightColor[1], L.lightColor[2]);
    if (in.type < 0.5f)  // deferred Blinn-Phong receives viewDir where the halfway vector is expected (forkergl.cpp:369)
        return blinn_phong_light(lc, lightDir, viewDir, in.nrm, visibility, in.albedo, in.emissive, in.param, rad);
    return pbr_light(lc, lightDir, viewDir, vnormalize(vadd(lightDir, viewDir)), in.nrm, visibility, in.albedo, in.emissive, in.param, rad);
}

template <int VEC>
__global__ void __launch_bounds__(128) k_lighting_hard(LightPass L)
{
    const size_t n = (size_t)L.W * L.H;
    const size_t begin = (size_t)L.row0 * L.W, end = (size_t)L.row1 * L.W;
    size_t       base = begin + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (base >= end) return;
    V3 eye = v3(L.eye[0], L.eye[1], L.eye[2]), lp = v3(L.lightPos[0], L.lightPos[1], L.lightPos[2]);

    float in[19][VEC];
    {
        const float* src[19];
        int          k = 0;
        for (int c = 0; c < 3; ++c) src[k++] = L.planes.p[FGL_PLANE_WORLDPOS] + c * n;
        for (int c = 0; c < 3; ++c) src[k++] = L.planes.p[FGL_PLANE_NORMAL] + c * n;
        for (int c = 0; c < 3; ++c) src[k++] = L.shadowOn ? L.planes.p[FGL_PLANE_LIGHTNDC] + c * n : nullptr;
        for (int c = 0; c < 3; ++c) src[k++] = L.planes.p[FGL_PLANE_ALBEDO] + c * n;
        for (int c = 0; c < 3; ++c) src[k++] = L.planes.p[FGL_PLANE_EMISSIVE] + c * n;
        for (int c = 0; c < 3; ++c) src[k++] = L.planes.p[FGL_PLANE_PARAM] + c * n;
        src[k++] = L.planes.p[FGL_PLANE_SHADINGTYPE];
#pragma unroll
        for (int j = 0; j < 19; ++j)
        {
            if (src[j] == nullptr)
            {
#pragma unroll
                for (int v = 0; v < VEC; ++v) in[j][v] = 0.f;
            }
            else if constexpr (VEC == 4)
            {
                float4 q = __ldcs(reinterpret_cast<const float4*>(src[j] + base));
                in[j][0] = q.x, in[j][1] = q.y, in[j][2] = q.z, in[j][3] = q.w;
            }
            else if constexpr (VEC == 2)
            {
                float2 q = __ldcs(reinterpret_cast<const float2*>(src[j] + base));
                in[j][0] = q.x, in[j][1] = q.y;
            }
            else in[j][0] = __ldcs(src[j] + base);
        }
    }
    float ao[VEC];
    if constexpr (VEC == 4)
    {
        float4 q = __ldcs(reinterpret_cast<const float4*>(L.planes.p[FGL_PLANE_AO] + base));
        ao[0] = q.x, ao[1] = q.y, ao[2] = q.z, ao[3] = q.w;
    }
    else if constexpr (VEC == 2)
    {
        float2 q = __ldcs(reinterpret_cast<const float2*>(L.planes.p[FGL_PLANE_AO] + base));
        ao[0] = q.x, ao[1] = q.y;
    }
    else ao[0] = __ldcs(L.planes.p[FGL_PLANE_AO] + base);

    float   outc[3][VEC];
    uint8_t q8[3 * VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v)
    {
        PixelIn p;
        p.pos = v3(in[0][v], in[1][v], in[2][v]);
        p.nrm = v3(in[3][v], in[4][v], in[5][v]);
        p.lndc = v3(in[6][v], in[7][v], in[8][v]);
        p.albedo = v3(in[9][v], in[10][v], in[11][v]);
        p.emissive = v3(in[12][v], in[13][v], in[14][v]);
        p.param = v3(in[15][v], in[16][v], in[17][v]);
        p.type = in[18][v];
        p.param.x *= ao[v];
        V3    lightDir = vnormalize(vsub(lp, p.pos)), viewDir = vnormalize(vsub(eye, p.pos));
        float visibility = 0.f;
        if (L.vis) visibility = L.vis[base + v];  // PCF / PCSS: filtered by the warp-per-pixel kernels of stream.cu
        else if (L.shadowOn)
        {   // shadow.cpp:109-132 with HardShadow (shadow.cpp:38-45)
            V3    sc = vadd(vscale(p.lndc, 0.5f), v3(0.5f, 0.5f, 0.5f));
            float bias = fmaxf(L.biasSlope * (1.f - vdot(p.nrm, lightDir)), L.biasMin);
            float sampled = shadow_lookup(L.sm, sc.x, sc.y);
            visibility = (sc.z <= sampled + bias) ? 1.f : 0.f;
        }
        V3 col = shade_pixel(L, p, visibility, lightDir, viewDir);
        outc[0][v] = col.x, outc[1][v] = col.y, outc[2][v] = col.z;
        q8[3 * v] = quant8(col.x), q8[3 * v + 1] = quant8(col.y), q8[3 * v + 2] = quant8(col.z);
    }
    if (L.writeF32)
    {
#pragma unroll
        for (int c = 0; c < 3; ++c)
        {
            float* dst = L.planes.p[FGL_PLANE_FRAME] + c * n + base;
            if constexpr (VEC == 4) __stcs(reinterpret_cast<float4*>(dst), make_float4(outc[c][0], outc[c][1], outc[c][2], outc[c][3]));
            else if constexpr (VEC == 2) __stcs(reinterpret_cast<float2*>(dst), make_float2(outc[c][0], outc[c][1]));
            else *dst = outc[c][0];
        }
    }
    if constexpr (VEC == 4)
    {
        uint32_t w[3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
            w[i] = (uint32_t)q8[4 * i] | ((uint32_t)q8[4 * i + 1] << 8) | ((uint32_t)q8[4 * i + 2] << 16) | ((uint32_t)q8[4 * i + 3] << 24);
        uint32_t* dst = reinterpret_cast<uint32_t*>(L.rgb8 + base * 3);
        dst[0] = w[0], dst[1] = w[1], dst[2] = w[2];
    }
    else if constexpr (VEC == 2)
    {
        uint16_t* dst = reinterpret_cast<uint16_t*>(L.rgb8 + base * 3);
        dst[0] = (uint16_t)(q8[0] | (q8[1] << 8)), dst[1] = (uint16_t)(q8[2] | (q8[3] << 8)), dst[2] = (uint16_t)(q8[4] | (q8[5] << 8));
    }
    else
    {
        L.rgb8[base * 3] = q8[0], L.rgb8[base * 3 + 1] = q8[1], L.rgb8[base * 3 + 2] = q8[2];
    }
}

// ---- quantise / SSAA ---------------------------------------------------------------------------------------------
__global__ void k_quantize(const float* frame, size_t n, uint8_t* rgb8)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    rgb8[i * 3] = quant8(frame[i]), rgb8[i * 3 + 1] = quant8(frame[n + i]), rgb8[i * 3 + 2] = quant8(frame[2 * n + i]);
}

// render.cpp:291-343: k x k integer box over the quantised image; int sum / float(k*k), truncated
__global__ void k_ssaa(const uint8_t* rgb8, int W, int H, int k, uint8_t* out, int orow0, int orow1)
{
    int ow = W / k;
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = orow0 + blockIdx.y;
    if (x >= ow || y >= orow1) return;
    int R = 0, G = 0, B = 0;
    for (int j = 0; j < k; ++j)
        for (int i = 0; i < k; ++i)
        {
            const uint8_t* p = rgb8 + ((size_t)(x * k + i) + (size_t)(y * k + j) * W) * 3;
            R += p[0], G += p[1], B += p[2];
        }
    float kk = (float)(k * k);
    R = f2i_x86((float)R / kk), G = f2i_x86((float)G / kk), B = f2i_x86((float)B / kk);
    uint8_t* o = out + ((size_t)x + (size_t)y * ow) * 3;
    o[0] = (uint8_t)R, o[1] = (uint8_t)G, o[2] = (uint8_t)B;
}

// ---- in-place two-pass Gaussian (buffer.cpp:59-98) -------------------------------------------------------------------
// Each pass is the 4th-order recurrence  y[w] = g0 x[w] + sum_i g_i (x[min(w+i, W-1)] + y'[w-i]),  y'[j<0] = a[0]
// (a[0] is still x[0] while w == 0 and y[0] afterwards), accumulated in the reference's order: centre, then for
// i = 1..4 the right tap, then the left tap.  Rows (H pass) / columns (V pass) are independent.
#define BLUR_G0 0.227027f
#define BLUR_G1 0.1945946f
#define BLUR_G2 0.1216216f
#define BLUR_G3 0.054054f
#define BLUR_G4 0.016216f

struct BlurState
{
    float h1, h2, h3, h4;  // y[w-1] .. y[w-4]
};
__device__ __forceinline__ float blur_step(BlurState& s, float x0, float x1, float x2, float x3, float x4, bool first)
{
    float r = x0 * BLUR_G0;
    r += x1 * BLUR_G1;
    r += s.h1 * BLUR_G1;
    r += x2 * BLUR_G2;
    r += s.h2 * BLUR_G2;
    r += x3 * BLUR_G3;
    r += s.h3 * BLUR_G3;
    r += x4 * BLUR_G4;
    r += s.h4 * BLUR_G4;
    if (first) s.h1 = s.h2 = s.h3 = s.h4 = r;  // from w == 1 on, every clamped left tap reads the new a[0]
    else s.h4 = s.h3, s.h3 = s.h2, s.h2 = s.h1, s.h1 = r;
    return r;
}

// Both passes are bound by the recurrence itself: from y[w-1] to y[w] there is one multiply and seven dependent adds
// (32 cycles; the summation order is part of the result), so a line of n samples costs >= 32 n cycles no matter how
// wide the machine is.  Parallelism therefore has to come from somewhere else:
//   * lines are independent: a CTA owns 32 lines, ONE warp walks them (lane = line) out of shared memory while the other
//     seven warps load tile i and store tile i - 2 (a three-stage ring, one __syncthreads per stage);
//   * a line is cut into CHUNKS of kBlurChunk samples that run concurrently (grid.y).  A chunk that does not begin at the
//     line's start walks kBlurWarm samples in front of its range first, from an arbitrary state: the recurrence forgets
//     (its feedback weights sum to 0.386), so after the warm-up its four history values have — so far always — become
//     bit-identical to the true ones.  That is CHECKED, not assumed: the chunk records the history it arrived with, and
//     k_blur_check compares it with what the chunk in front of it really wrote; a line with a mismatch is flagged and
//     recomputed serially by k_blur_fix.  If every boundary matches, every chunk continued from the exact state and the
//     result is the reference's, bit for bit.
// The passes are out of place (H: plane -> scratch, V: scratch -> plane), which is what lets a chunk read original
// samples in front of its range while its neighbour is writing there; the reference's in-place loop reads exactly the same
// mix (already blurred behind, still original ahead).
constexpr int kBlurTW = 96, kBlurLD = kBlurTW + 5;  // H: tile width, row pitch (101 = 5 mod 32: conflict-free)
constexpr int kBlurTR = 64;                         // V: tile rows
constexpr int kBlurChunk = 768, kBlurWarmUp = 64;   // samples per chunk (a multiple of both tile sizes), warm-up samples

struct BlurChunks
{
    int    begin, end;   // samples [begin, end) of every line are produced (begin > 0: a sort-first band's warm start, unchecked)
    int    nChunks;
    float* boundary;     // [line][chunk][4]: the history a chunk arrived with at its first sample (chunks >= 1)
};

__global__ void __launch_bounds__(256) k_blur_h(const float* src, float* dst, int W, int H, int rowBegin, BlurChunks C)
{
    __shared__ float tile[3][32][kBlurLD];
    const int row0 = rowBegin + blockIdx.x * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = blockIdx.y;
    const int outBegin = C.begin + chunk * kBlurChunk, outEnd = min(C.end, outBegin + kBlurChunk);
    const int walkBegin = chunk == 0 ? outBegin : max(0, outBegin - kBlurWarmUp);
    const int nT = (outEnd - walkBegin + kBlurTW - 1) / kBlurTW;
    BlurState st;
    st.h1 = st.h2 = st.h3 = st.h4 = 0.f;
    for (int i = 0; i < nT + 2; ++i)
    {
        if (warp == 0)
        {
            const int k = i - 1;
            if (k >= 0 && k < nT && row0 + lane < H)
            {
                float* t = tile[k % 3][lane];
                const int c0 = walkBegin + k * kBlurTW, tw = min(kBlurTW, outEnd - c0);
                if (k == 0) st.h1 = st.h2 = st.h3 = st.h4 = t[0];
                float x0 = t[0], x1 = t[1], x2 = t[2], x3 = t[3];
#pragma unroll 8
                for (int j = 0; j < tw; ++j)
                {
                    if (chunk > 0 && c0 + j == outBegin)
                    {
                        float4* bq = reinterpret_cast<float4*>(C.boundary + ((size_t)(row0 + lane) * C.nChunks + chunk) * 4);
                        *bq = make_float4(st.h1, st.h2, st.h3, st.h4);
                    }
                    float x4 = t[j + 4];
                    float r = blur_step(st, x0, x1, x2, x3, x4, c0 + j == 0);
                    t[j] = r;
                    x0 = x1, x1 = x2, x2 = x3, x3 = x4;
                }
            }
        }
        else
        {
            const int ks = i - 2;
            if (ks >= 0)
            {
                const int c0 = walkBegin + ks * kBlurTW, tw = min(kBlurTW, outEnd - c0);
                for (int r = warp - 1; r < 32; r += 7)
                {
                    int row = row0 + r;
                    if (row >= H) break;
                    for (int j = lane; j < tw; j += 32)
                        if (c0 + j >= outBegin) dst[(size_t)row * W + c0 + j] = tile[ks % 3][r][j];
                }
            }
            if (i < nT)
            {   // asynchronous copies (LDGSTS): all of a thread's loads are in flight at once, nothing is staged in registers
                const int c0 = walkBegin + i * kBlurTW, tw4 = min(kBlurTW, outEnd - c0) + 4;
                for (int e = threadIdx.x - 32; e < 32 * tw4; e += 224)
                {
                    int r = e / tw4, j = e - r * tw4, row = row0 + r;
                    if (row < H) __pipeline_memcpy_async(&tile[i % 3][r][j], &src[(size_t)row * W + min(c0 + j, W - 1)], 4);
                }
                __pipeline_commit();
            }
            __pipeline_wait_prior(0);
        }
        __syncthreads();
    }
}

// V pass: a CTA owns 32 columns, lane = column; chunks run along the rows.  C.begin > 0 (sort-first band): the first chunk
// starts from a row >= 64 above the band as if it were the plane's first row; its memory of the start decays as 0.61^n, far
// below fp32 resolution after 64 rows — the rows in front of the band belong to another GPU, so that one
// warm-up cannot be checked here (the group's frame hash is).  Look-ahead rows past the bottom edge re-read the last row.
__global__ void __launch_bounds__(256) k_blur_v(const float* src, float* dst, int W, int H, BlurChunks C)
{
    __shared__ float tile[3][kBlurTR + 4][32];
    const int  warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int  x = blockIdx.x * 32 + lane;
    const bool colOk = x < W;
    const int  chunk = blockIdx.y;
    const int  outBegin = C.begin + chunk * kBlurChunk, outEnd = min(C.end, outBegin + kBlurChunk);
    const int  walkBegin = chunk == 0 ? outBegin : max(0, outBegin - kBlurWarmUp);
    const int  nT = (outEnd - walkBegin + kBlurTR - 1) / kBlurTR;
    BlurState  st;
    st.h1 = st.h2 = st.h3 = st.h4 = 0.f;
    for (int i = 0; i < nT + 2; ++i)
    {
        if (warp == 0)
        {
            const int k = i - 1;
            if (k >= 0 && k < nT)
            {
                float(*t)[32] = tile[k % 3];
                const int r0 = walkBegin + k * kBlurTR, th = min(kBlurTR, outEnd - r0);
                if (k == 0) st.h1 = st.h2 = st.h3 = st.h4 = t[0][lane];
                float x0 = t[0][lane], x1 = t[1][lane], x2 = t[2][lane], x3 = t[3][lane];
#pragma unroll 8
                for (int j = 0; j < th; ++j)
                {
                    if (chunk > 0 && r0 + j == outBegin && colOk)
                    {
                        float4* bq = reinterpret_cast<float4*>(C.boundary + ((size_t)x * C.nChunks + chunk) * 4);
                        *bq = make_float4(st.h1, st.h2, st.h3, st.h4);
                    }
                    float x4 = t[j + 4][lane];
                    float r = blur_step(st, x0, x1, x2, x3, x4, chunk == 0 && r0 + j == C.begin);
                    t[j][lane] = r;
                    x0 = x1, x1 = x2, x2 = x3, x3 = x4;
                }
            }
        }
        else
        {
            const int ks = i - 2;
            if (ks >= 0 && colOk)
            {
                const int r0 = walkBegin + ks * kBlurTR, th = min(kBlurTR, outEnd - r0);
                for (int r = warp - 1; r < th; r += 7)
                    if (r0 + r >= outBegin) dst[(size_t)(r0 + r) * W + x] = tile[ks % 3][r][lane];
            }
            if (i < nT)
            {
                const int r0 = walkBegin + i * kBlurTR, th = min(kBlurTR, outEnd - r0);
                if (colOk)
                    for (int r = warp - 1; r < th + 4; r += 7) __pipeline_memcpy_async(&tile[i % 3][r][lane], &src[(size_t)min(r0 + r, H - 1) * W + x], 4);
                __pipeline_commit();
            }
            __pipeline_wait_prior(0);
        }
        __syncthreads();
    }
}

// boundary check: the four values in front of chunk k's first sample, as the chunk in front of it wrote them, against the
// history chunk k arrived with (bit for bit).  stride: distance between consecutive samples of a line, pitch: between lines.
__global__ void k_blur_check(const float* out, size_t stride, size_t pitch, int nLines, int lineBase, BlurChunks C, unsigned* lineFlags)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int line = i / C.nChunks, chunk = i - line * C.nChunks;
    if (line >= nLines || chunk == 0) return;
    const int     at = C.begin + chunk * kBlurChunk;
    const float*  q = out + (size_t)(lineBase + line) * pitch;
    const float4  h = *reinterpret_cast<const float4*>(C.boundary + ((size_t)(lineBase + line) * C.nChunks + chunk) * 4);
    const bool    same = __float_as_uint(q[(size_t)(at - 1) * stride]) == __float_as_uint(h.x) && __float_as_uint(q[(size_t)(at - 2) * stride]) == __float_as_uint(h.y) &&
                      __float_as_uint(q[(size_t)(at - 3) * stride]) == __float_as_uint(h.z) && __float_as_uint(q[(size_t)(at - 4) * stride]) == __float_as_uint(h.w);
    if (!same) atomicOr(lineFlags + (lineBase + line), 1u);
}
// flagged lines, recomputed from their first sample by one thread each (never taken so far; keeps the result exact if it ever is)
__global__ void k_blur_fix(const float* src, float* dst, size_t stride, size_t pitch, int nLines, int lineBase, int len, BlurChunks C, unsigned* lineFlags)
{
    int line = blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= nLines || !lineFlags[lineBase + line]) return;
    lineFlags[lineBase + line] = 0u;
    const float* x = src + (size_t)(lineBase + line) * pitch;
    float*       y = dst + (size_t)(lineBase + line) * pitch;
    BlurState    st;
    st.h1 = st.h2 = st.h3 = st.h4 = x[(size_t)C.begin * stride];
    for (int w = C.begin; w < C.end; ++w)
    {
        auto at = [&](int k) { return x[(size_t)min(w + k, len - 1) * stride]; };
        y[(size_t)w * stride] = blur_step(st, at(0), at(1), at(2), at(3), at(4), w == C.begin);
    }
}

// ---- in-place 3x3 box (buffer.cpp:35-57 / 140-162) --------------------------------------------------------------------
// Pixel (w, h) reads row h - 1 and (w - 1, h) already blurred and everything else still original (raster order, in
// place, clamped taps that read "whatever is there"), so the pixels of a wavefront t = w + 2 h are independent.  One CTA
// walks up to 1024 rows in lockstep: thread = row, row i lags row i - 1 by two columns, one __syncthreads per step; the
// blurred value of the row above travels through a shared-memory slot, the originals of the own row and of the row
// below are fetched eight steps at a time.  Not on the frame path (Render::Render never calls it): latency-bound by
// design, (W + 2 min(H, 1024)) steps per pass of 1024 rows.
__global__ void __launch_bounds__(1024) k_simple_blur(float* a, int W, int H, int rowBase, int nRows)
{
    __shared__ float exch[2][1024];
    const int   i = threadIdx.x, h = rowBase + i;
    const bool  live = i < nRows;
    const bool  topRow = h == 0, bottomRow = h == H - 1, fromGlobal = i == 0 && h > 0;
    const float s = 1 / 9.f;
    float       up0 = 0.f, up1 = 0.f, up2 = 0.f;  // blurred row h - 1 at w - 1, w, w + 1
    float       oL = 0.f, oM = 0.f, oR = 0.f;      // own row: blurred (w - 1), original w, original w + 1
    float       dn0 = 0.f, dn1 = 0.f, dn2 = 0.f;  // original row h + 1 at w - 1, w, w + 1
    const int   tEnd = W - 1 + 2 * (nRows - 1);   // last step (the first is -1: every row has a set-up step at w = -1)
    for (int T0 = -1; T0 <= tEnd; T0 += 8)
    {
        float bo[8], bd[8], bu[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
        {
            const int w = T0 + k - 2 * i;
            bo[k] = bd[k] = bu[k] = 0.f;
            if (live && w >= -1 && w <= W - 1)
            {
                const int c = min(w + 1, W - 1);
                bo[k] = a[(size_t)h * W + c];
                if (!bottomRow) bd[k] = a[(size_t)(h + 1) * W + c];
                if (fromGlobal) bu[k] = a[(size_t)(h - 1) * W + c];
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
        {
            const int t = T0 + k, w = t - 2 * i;
            if (live && w >= -1 && w <= W - 1)
            {
                float newU = up2;
                if (w + 1 <= W - 1 && !topRow) newU = fromGlobal ? bu[k] : exch[(t - 1) & 1][i - 1];
                up0 = up1, up1 = up2, up2 = newU;
                oM = oR, oR = bo[k];
                dn0 = dn1, dn1 = dn2, dn2 = bd[k];
                if (w >= 0)
                {
                    const float xl = w == 0 ? oM : oL, xm = oM, xr = oR;  // the own row as it stands: (w - 1) blurred, the rest original
                    const float al = topRow ? xl : (w == 0 ? up1 : up0), am = topRow ? xm : up1, ar = topRow ? xr : up2;
                    const float bl = bottomRow ? xl : (w == 0 ? dn1 : dn0), bm = bottomRow ? xm : dn1, br = bottomRow ? xr : dn2;
                    float       r = 0.f;
                    r += al * s, r += xl * s, r += bl * s;  // xOffset = -1: yOffset = -1, 0, +1
                    r += am * s, r += xm * s, r += bm * s;
                    r += ar * s, r += xr * s, r += br * s;
                    a[(size_t)h * W + w] = r;
                    exch[t & 1][i] = r;
                    oL = r;
                }
            }
            __syncthreads();
        }
    }
}

// ---- SSAO, one warp per pixel: lane = hemisphere sample (render.cpp:229-285) -----------------------------------------
// The kernel is bound by instruction issue (profiles/), so everything that cannot change the result is trimmed:
//  * a warp walks kSsaoPPW consecutive pixels: the matrices, the viewport and the loop-invariant address arithmetic stay in
//    registers instead of being re-fetched from the constant bank for every pixel; the next pixel's sample is loaded while
//    the current one is being projected.
//  * VIEWPORT_AFFINE: ForkerGL::SetViewportMatrix (forkergl.cpp:89-102) only fills [0][0], [0][3], [1][1], [1][3], [2][2],
//    [2][3]; with finite x, y, z the reference's row products ((0 + m0 x) + m1 y) + m2 z) + m3 w reduce to
//    (0 + m_k v_k) + m3 w bit for bit (adding 0 * finite = +-0 to a sum that is +0 or non-zero changes nothing), and
//    the w row is never used.  float -> int: cvttss2si only differs from a plain truncation outside +-2^31 (INT_MIN).
//    The same variant drops the "0 +" every Dot of the reference starts with (geometry.h:774-793): 0 + a differs from a only
//    for a = -0, and a sum whose LAST term is not -0 comes out the same whichever zero the partial sum carried (non-zero
//    term: t; +0: +0).  The last terms are viewProj[r][3] * 1 and viewport[r][3] * ndc.w with ndc.w > 0 or NaN; the host
//    selects this variant only when none of those seven matrix entries is -0.
//    Both shortcuts are taken by the whole warp or not at all (one vote per pixel instead of a branch per sample).
//  * background pixels (nothing drawn: depth = FLT_MAX, position = normal = 0) with the range check on: every sample sits
//    within the SSAO radius of the world origin; when the host has verified that the projection of that ball is finite
//    (S.backgroundIsOne: |w_clip| bounded away from 0, see fgl_run_ssao) the sample's depth is finite, so
//    "z >= cached + bias" fails against a cleared texel (FLT_MAX) and "|FLT_MAX - cached| < range" fails against a drawn
//    one: the occlusion is exactly 0 and the pixel's AO exactly 1 without looking at a single sample.
constexpr int kSsaoPPW = 4;

template <bool VIEWPORT_AFFINE>
__global__ void __launch_bounds__(256) k_ssao(SsaoPass S)
{
    const unsigned n = (unsigned)S.W * (unsigned)S.H;  // planes have < 2^31 pixels (plane_init)
    const int      lane = threadIdx.x & 31;
    const unsigned first = (unsigned)S.row0 * (unsigned)S.W + (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kSsaoPPW;
    const unsigned end = (unsigned)S.row1 * (unsigned)S.W;
    if (first >= end) return;
    const unsigned last = min(first + kSsaoPPW, end);
    const float    inf = __int_as_float(0x7f800000);
    // accepted unit-ball sample number 32 * pixel + lane of the replayed stream (geometry.h:957-966)
    // (one 16-byte record: the vector and its scale Lerp(0.1, 1, |v|^2), which k_compact computed with this kernel's arithmetic)
    const float4* b = S.ball + ((size_t)first * 32 + lane);
    float4        qNext = __ldcs(b);
    for (unsigned idx = first; idx < last; ++idx)
    {
        const float4 q = qNext;
        if (idx + 1 < last)
        {
            b += 32;
            qNext = __ldcs(b);
        }
        V3 v = v3(q.x, q.y, q.z);
        const float fragDepth = S.depth[idx];
        const float* wp = S.worldpos + idx;
        const V3     pos = v3(wp[0], wp[n], wp[2 * (size_t)n]);
        if (S.backgroundIsOne && fragDepth >= 3.402823466e+38f && pos.x == 0.f && pos.y == 0.f && pos.z == 0.f)
        {
            if (lane == 0) S.ao[idx] = 1.f;
            continue;
        }
        const float* np = S.normal + idx;
        const V3     nrm = v3(np[0], np[n], np[2 * (size_t)n]);
        // geometry.h:978-990.  Dot starts its sum at 0: "0 + a" differs from a only for a = -0, which can only change the SIGN of a
        // zero sum — invisible to the comparison
        if (!((v.x * nrm.x + v.y * nrm.y) + v.z * nrm.z > 0.f)) v = v3(-v.x, -v.y, -v.z);
        v = vscale(v, q.w);  // q.w = ssao_sample_scale(v): Lerp(0.1f, 1.0f, |v|^2), unchanged by the flip
        V3 sp = vadd(pos, vscale(v, S.radius));
        V4 sp4;
        sp4.x = sp.x, sp4.y = sp.y, sp4.z = sp.z, sp4.w = 1.f;
        V4 cs;
        if (VIEWPORT_AFFINE)
        {   // the rows' last terms (m[r][3] * 1) are not -0 (host check): the leading "0 +" of Dot cannot change the sums
            const float* m = S.viewProj;
            cs.x = ((m[0] * sp.x + m[1] * sp.y) + m[2] * sp.z) + m[3];
            cs.y = ((m[4] * sp.x + m[5] * sp.y) + m[6] * sp.z) + m[7];
            cs.z = ((m[8] * sp.x + m[9] * sp.y) + m[10] * sp.z) + m[11];
            cs.w = ((m[12] * sp.x + m[13] * sp.y) + m[14] * sp.z) + m[15];
        }
        else cs = mat4mul(S.viewProj, sp4);
        V4 ndc = vdivs4(cs, cs.w);
        // (ndc.w = cs.w * (1 / cs.w) is positive or NaN, so the second term has the sign of the viewport entry: not -0 either)
        float ssx = S.viewport[0] * ndc.x + S.viewport[3] * ndc.w;
        float ssy = S.viewport[5] * ndc.y + S.viewport[7] * ndc.w;
        float ssz = S.viewport[10] * ndc.z + S.viewport[11] * ndc.w;
        int   sx = (int)ssx, sy = (int)ssy;
        const bool plain = VIEWPORT_AFFINE && fabsf(ndc.x) < inf && fabsf(ndc.y) < inf && fabsf(ndc.z) < inf && fabsf(ssx) < 1.0e9f && fabsf(ssy) < 1.0e9f;
        if (!__all_sync(0xffffffffu, plain))
        {   // the general products and the x86 conversion, for every lane of the pixel
            V4 ss = mat4mul(S.viewport, ndc);
            ssx = ss.x, ssy = ss.y, ssz = ss.z;
            sx = f2i_x86(ssx), sy = f2i_x86(ssy);
        }
        long long li = (long long)sx + (long long)sy * S.W;  // unchecked linear index in the reference (buffer.h:37)
        bool      occ = false;
        if (li >= 0 && li < (long long)n)
        {
            float cached = __ldg(S.depth + li);
            if (ssz >= cached + S.bias) occ = S.rangeCheck ? (fabsf(fragDepth - cached) < S.rangeCheckRadius) : true;
        }
        int cnt = __popc(__ballot_sync(0xffffffffu, occ));
        if (lane == 0)
        {
            float o = 1.f - (float)cnt * (1.f / 32.f);  // 1/32 steps accumulate exactly
            S.ao[idx] = (o * o) * o;                    // pow(o, 3): k^3 / 32768 is exact in fp32
        }
    }
}

// ---- host accessor helpers -----------------------------------------------------------------------------------------
__global__ void k_ids(const unsigned long long* vis, size_t n, int* out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long k = vis[i];
    out[i] = k == FGL_VIS_EMPTY ? -1 : (int)(unsigned)(k & 0xffffffffull);
}
__global__ void k_soa_to_aos(const float* soa, float* aos, size_t n, int ch)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int c = 0; c < ch; ++c) aos[i * ch + c] = soa[(size_t)c * n + i];
}
__global__ void k_aos_to_soa(const float* aos, float* soa, size_t n, int ch)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int c = 0; c < ch; ++c) soa[(size_t)c * n + i] = aos[i * ch + c];
}

int check_launch(fgl_ctx* c, const char* what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fgl_fail(c, FGL_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return FGL_OK;
}
}  // namespace

int fgl_run_fill(fgl_ctx* c, float* dst, size_t n, float value)
{
    if (!n) return FGL_OK;
    LaunchScope ls(c, "fill", n * 4);
    unsigned    blocks = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)c->numSMs * 16);
    k_fill<<<blocks, 256, 0, c->stream>>>(dst, n, value);
    return check_launch(c, "fill");
}

int fgl_run_fill_rgb(fgl_ctx* c, float* dst, size_t nPixels, const float rgb[3])
{
    for (int ch = 0; ch < 3; ++ch)
        if (int rc = fgl_run_fill(c, dst + ch * nPixels, nPixels, rgb[ch])) return rc;
    return FGL_OK;
}

int fgl_run_lighting(fgl_ctx* c, const LightPass& L)
{
    size_t   nPix = (size_t)L.W * (L.row1 - L.row0);
    if (!nPix) return FGL_OK;  // an empty band
    uint64_t bytes = nPix * (80 + 3 + (L.writeF32 ? 12 : 0));
    {
        LaunchScope ls(c, "lighting", bytes + (L.vis ? nPix * 4 : 0));
        bool        vec = (((size_t)L.W * L.H) % 4 == 0) && (((size_t)L.row0 * L.W) % 4 == 0) && (nPix % 4 == 0);
        static int  vecWidth = getenv("FGL_LIGHT_VEC") ? atoi(getenv("FGL_LIGHT_VEC")) : 2;  // 2 pixels/thread: 64-bit loads, 40 % more resident warps than 4 (profiles/)
        if (vec && vecWidth == 4) k_lighting_hard<4><<<(unsigned)((nPix / 4 + 127) / 128), 128, 0, c->stream>>>(L);
        else if (vec && vecWidth == 2) k_lighting_hard<2><<<(unsigned)((nPix / 2 + 127) / 128), 128, 0, c->stream>>>(L);
        else k_lighting_hard<1><<<(unsigned)((nPix + 127) / 128), 128, 0, c->stream>>>(L);
    }
    return check_launch(c, "lighting");
}

int fgl_run_quantize(fgl_ctx* c, const float* frame, size_t nPixels, uint8_t* rgb8)
{
    LaunchScope ls(c, "quantize", nPixels * 15);
    k_quantize<<<(unsigned)((nPixels + 255) / 256), 256, 0, c->stream>>>(frame, nPixels, rgb8);
    return check_launch(c, "quantize");
}

int fgl_run_ssaa(fgl_ctx* c, const uint8_t* rgb8, int W, int H, int k, uint8_t* out, int row0, int row1)
{
    int ow = W / k, orow0 = row0 / k, orow1 = row1 / k;
    if (ow <= 0 || orow1 <= orow0) return FGL_OK;
    LaunchScope ls(c, "ssaa_resolve", (uint64_t)ow * (orow1 - orow0) * (3 * k * k + 3));
    dim3        grid((ow + 127) / 128, orow1 - orow0);
    k_ssaa<<<grid, 128, 0, c->stream>>>(rgb8, W, H, k, out, orow0, orow1);
    return check_launch(c, "ssaa");
}

int fgl_run_blur(fgl_ctx* c, float* plane, int W, int H, int channels, int kind, int hRow0, int hRow1, int vRow0, int vRow1)
{
    size_t n = (size_t)W * H;
    if (kind == FGL_BLUR_SIMPLE_3X3)
    {
        if (hRow0 != 0 || hRow1 != H || vRow0 != 0 || vRow1 != H)
            return fgl_fail(c, FGL_ERR_UNSUPPORTED, "fgl_blur: the in-place 3x3 box is a whole-plane recurrence (no row bands)");
        for (int ch = 0; ch < channels; ++ch)
            for (int row = 0; row < H; row += 1024)
            {
                LaunchScope ls(c, "simple_blur", (uint64_t)W * std::min(1024, H - row) * 8);
                k_simple_blur<<<1, 1024, 0, c->stream>>>(plane + ch * n, W, H, row, std::min(1024, H - row));
            }
        return check_launch(c, "simple blur");
    }
    if (kind != FGL_BLUR_TWO_PASS_GAUSSIAN) return fgl_fail(c, FGL_ERR_INVALID, "fgl_blur: unknown blur kind");
    if (hRow1 <= hRow0 || vRow1 <= vRow0) return FGL_OK;
    // scratch plane for the H pass + the chunks' boundary records and line flags
    BlurChunks CH, CV;
    CH.begin = 0, CH.end = W, CH.nChunks = (W + kBlurChunk - 1) / kBlurChunk;
    CV.begin = vRow0, CV.end = vRow1, CV.nChunks = (vRow1 - vRow0 + kBlurChunk - 1) / kBlurChunk;
    const size_t bndFloats = std::max((size_t)H * CH.nChunks, (size_t)W * CV.nChunks) * 4, nFlags = (size_t)std::max(W, H);
    const size_t n4 = (n + 3) & ~(size_t)3;  // the boundary records are read as float4
    if (int rc = fgl_reserve(c, c->blurTmp, n4 * 4 + bndFloats * 4 + nFlags * 4 + 64)) return rc;
    float*    tmp = (float*)c->blurTmp.p;
    float*    bnd = tmp + n4;
    unsigned* flags = (unsigned*)(bnd + bndFloats);
    CH.boundary = CV.boundary = bnd;
    cudaStream_t st = c->stream;
    for (int ch = 0; ch < channels; ++ch)
    {
        float* a = plane + ch * n;
        if (cudaMemsetAsync(flags, 0, nFlags * 4, st) != cudaSuccess) return check_launch(c, "blur");
        {
            LaunchScope ls(c, "blur_h", (size_t)(hRow1 - hRow0) * W * 8);
            k_blur_h<<<dim3((hRow1 - hRow0 + 31) / 32, CH.nChunks), 256, 0, st>>>(a, tmp, W, hRow1, hRow0, CH);
            if (CH.nChunks > 1)
            {
                const int lines = hRow1 - hRow0;
                k_blur_check<<<(lines * CH.nChunks + 255) / 256, 256, 0, st>>>(tmp, 1, (size_t)W, lines, hRow0, CH, flags);
                k_blur_fix<<<(lines + 127) / 128, 128, 0, st>>>(a, tmp, 1, (size_t)W, lines, hRow0, W, CH, flags);
            }
        }
        {
            LaunchScope ls(c, "blur_v", (size_t)(vRow1 - vRow0) * W * 8);
            k_blur_v<<<dim3((W + 31) / 32, CV.nChunks), 256, 0, st>>>(tmp, a, W, H, CV);
            if (CV.nChunks > 1)
            {
                k_blur_check<<<(W * CV.nChunks + 255) / 256, 256, 0, st>>>(a, (size_t)W, 1, W, 0, CV, flags);
                k_blur_fix<<<(W + 127) / 128, 128, 0, st>>>(tmp, a, (size_t)W, 1, W, 0, H, CV, flags);
            }
        }
    }
    return check_launch(c, "blur");
}

int fgl_run_ssao(fgl_ctx* c, const SsaoPass& S)
{
    size_t      nPix = (size_t)S.W * (S.row1 - S.row0);
    if (!nPix) return FGL_OK;
    LaunchScope ls(c, "ssao", nPix * (32 + 512));
    const float* m = S.viewport;  // ForkerGL::SetViewportMatrix structure (rows 0-2; the w row is not used by SSAO)
    auto         negZero = [](float f) { return f == 0.f && std::signbit(f); };
    const bool   affine = m[1] == 0.f && m[2] == 0.f && m[4] == 0.f && m[6] == 0.f && m[8] == 0.f && m[9] == 0.f && !negZero(m[3]) && !negZero(m[7]) &&
                        !negZero(m[11]) && !negZero(S.viewProj[3]) && !negZero(S.viewProj[7]) && !negZero(S.viewProj[11]) && !negZero(S.viewProj[15]);
    // Background shortcut (see k_ssao): a sample of a background pixel lies within R = 1.0001 * radius of the world origin
    // (|v| < 1, scale in [0.9, 1]).  Its clip-space w = (a, b, c) . P + d is at least |d| - R (|a| + |b| + |c|) in magnitude
    // (less a rounding margin); if that is comfortably positive and every row of the matrix stays far from overflow, the
    // projected point is finite — which is all the shortcut needs.
    SsaoPass S2 = S;
    {
        const float* vp = S.viewProj;
        const double R = 1.0001 * fabs((double)S.radius);
        const double wMin = fabs((double)vp[15]) - R * (fabs((double)vp[12]) + fabs((double)vp[13]) + fabs((double)vp[14]));
        double       rowMax = 0;
        for (int r = 0; r < 3; ++r) rowMax = std::max(rowMax, fabs((double)vp[4 * r + 3]) + R * (fabs((double)vp[4 * r]) + fabs((double)vp[4 * r + 1]) + fabs((double)vp[4 * r + 2])));
        double vpMax = 0;
        for (int i = 0; i < 12; ++i) vpMax = std::max(vpMax, fabs((double)S.viewport[i]));
        static const bool off = getenv("FGL_SSAO_NO_BG") != nullptr;
        const bool finiteInputs = std::isfinite(wMin) && std::isfinite(rowMax) && std::isfinite(vpMax) && std::isfinite((double)S.bias);
        S2.backgroundIsOne = !off && S.rangeCheck && S.rangeCheckRadius < 1.0e30f && finiteInputs && wMin > 1e-3 * (1.0 + fabs((double)vp[15])) &&
                             (rowMax / wMin) * (1.0 + vpMax) * 4.0 < 1.0e30;
    }
    const unsigned warps = (unsigned)((nPix + kSsaoPPW - 1) / kSsaoPPW), blocks = (warps + 7) / 8;
    if (affine) k_ssao<true><<<blocks, 256, 0, c->stream>>>(S2);
    else k_ssao<false><<<blocks, 256, 0, c->stream>>>(S2);
    return check_launch(c, "ssao");
}

int fgl_run_ids(fgl_ctx* c, const unsigned long long* vis, size_t n, int* out)
{
    LaunchScope ls(c, "ids", n * 12);
    k_ids<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(vis, n, out);
    return check_launch(c, "ids");
}

int fgl_run_aos(fgl_ctx* c, const float* src, float* dst, size_t nPixels, int channels, bool toAos)
{
    LaunchScope ls(c, toAos ? "soa_to_aos" : "aos_to_soa", nPixels * channels * 8);
    if (toAos) k_soa_to_aos<<<(unsigned)((nPixels + 255) / 256), 256, 0, c->stream>>>(src, dst, nPixels, channels);
    else k_aos_to_soa<<<(unsigned)((nPixels + 255) / 256), 256, 0, c->stream>>>(src, dst, nPixels, channels);
    return check_launch(c, "layout conversion");
}
