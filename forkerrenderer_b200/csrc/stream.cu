// stream.cu — device replay of the reference's global mt19937(5489) sample stream (utility.h:90-103) and the
// resolution of the PCSS consumption chain (shadow.cpp:92-106).
//
// The stream of a frame is a constant: Render::Render starts at draw 0, SSAO (if on) consumes rejection-sampled
// unit-ball vectors (geometry.h:957-966: 3 draws per candidate, accepted iff x^2+y^2+z^2 < 1), lighting then consumes
// rejection-sampled unit-disk vectors (geometry.h:968-976: 2 draws per candidate).  Because every candidate uses a
// fixed number of draws, candidate boundaries are known in advance and acceptance is a parallel predicate; the k-th
// accepted sample is found with a prefix sum.  The compacted samples are kept in HBM as a table (built once per
// (seed, pixel count, SSAO on/off) and charged to the consuming pass as read bytes, DESIGN.md "sample stream").
//
//   k_mt_checkpoints   single CTA: walks the generator block by block (624 words; three dependent phases of <= 227
//                      lanes) and stores its state every kCB blocks.
//   k_mt_generate      one CTA per checkpoint: regenerates kCB blocks of tempered 32-bit draws.
//   k_count/k_compact  accept flag per candidate, tile sums, scan, ordered compaction into the sample table.
//
// PCSS: pixel p (scan order) uses 32 samples and 64 more iff its blocker search found a blocker, so its offset is
// 32 p + 64 k(p), k(p) = number of earlier pixels with a blocker.  Pixels are first classified with a min/max filter
// of the shadow map over the search footprint (certainly no blocker / certainly a blocker / uncertain); the uncertain
// pixels are then resolved in scan order, in super-chunks from an exactly known state, by ONE persistent cooperative
// kernel (k_chain_fused): candidate offsets around a pilot prediction are evaluated in parallel, 32-row segments are
// turned into transfer tables, and CTA 0 composes the tables in order.  The chain state crosses GPUs through
// peer-memory mailboxes (k_peer_wait / k_peer_notify).
#include <cub/block/block_reduce.cuh>
#include <cub/block/block_scan.cuh>
#include <cub/device/device_scan.cuh>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "stream.h"

#ifndef FGL_CHAIN_NW
#define FGL_CHAIN_NW 2  // PCSS chain: 32-candidate words per row window (W = 32 * FGL_CHAIN_NW)
#endif

namespace
{
constexpr int    kMT = 624;
constexpr int    kCB = 256;                      // generator blocks per checkpoint
constexpr size_t kWindowCand = (size_t)64 << 20;  // candidates per build window
constexpr int    kChainNW = FGL_CHAIN_NW;                   // PCSS chain: candidate offsets evaluated per row = 32 * kChainNW, centred on the prediction

__device__ __forceinline__ uint32_t mt_mix(uint32_t a, uint32_t b)
{
    uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
// one generator step for the whole 624-word state: A -> B (blockDim.x threads, all must call)
__device__ __forceinline__ void mt_twist(const uint32_t* A, uint32_t* B)
{
    for (int i = threadIdx.x; i < 227; i += blockDim.x) B[i] = A[i + 397] ^ mt_mix(A[i], A[i + 1]);
    __syncthreads();
    for (int i = 227 + threadIdx.x; i < 454; i += blockDim.x) B[i] = B[i - 227] ^ mt_mix(A[i], A[i + 1]);
    __syncthreads();
    for (int i = 454 + threadIdx.x; i < kMT; i += blockDim.x) B[i] = B[i - 227] ^ mt_mix(A[i], i == kMT - 1 ? B[0] : A[i + 1]);
    __syncthreads();
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y)
{
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// ckpt[j] = generator state before block j * kCB is produced.  Extends [have, want).
__global__ void __launch_bounds__(256) k_mt_checkpoints(uint32_t* ckpt, long long have, long long want)
{
    __shared__ uint32_t s[2][kMT];
    for (int i = threadIdx.x; i < kMT; i += blockDim.x) s[0][i] = ckpt[(have - 1) * kMT + i];
    __syncthreads();
    int cur = 0;
    for (long long ck = have; ck < want; ++ck)
    {
        for (int b = 0; b < kCB; ++b)
        {
            mt_twist(s[cur], s[cur ^ 1]);
            cur ^= 1;
        }
        for (int i = threadIdx.x; i < kMT; i += blockDim.x) ckpt[ck * kMT + i] = s[cur][i];
    }
}

// out[(c * kCB + b) * 624 + i] = draw number ((firstCk + c) * kCB + b) * 624 + i of the stream
__global__ void __launch_bounds__(256) k_mt_generate(const uint32_t* ckpt, long long firstCk, long long nBlocks, uint32_t* out)
{
    __shared__ uint32_t s[2][kMT];
    long long           c = blockIdx.x;
    for (int i = threadIdx.x; i < kMT; i += blockDim.x) s[0][i] = ckpt[(firstCk + c) * kMT + i];
    __syncthreads();
    int cur = 0;
    for (int b = 0; b < kCB; ++b)
    {
        long long blk = c * kCB + b;
        if (blk >= nBlocks) break;
        mt_twist(s[cur], s[cur ^ 1]);
        cur ^= 1;
        for (int i = threadIdx.x; i < kMT; i += blockDim.x) out[blk * kMT + i] = mt_temper(s[cur][i]);
    }
}

// geometry.h:952-976 with g++'s right-to-left evaluation of constructor arguments:
// ball: draw0 -> z, draw1 -> y, draw2 -> x; disk: draw0 -> y, draw1 -> x.
template <int K>
__device__ __forceinline__ bool candidate(const uint32_t* raw, size_t cand, float& x, float& y, float& z)
{
    const uint32_t* r = raw + cand * K;
    if (K == 3)
    {
        z = random_m1p1(r[0]), y = random_m1p1(r[1]), x = random_m1p1(r[2]);
        return !(x * x + y * y + z * z >= 1.f);
    }
    y = random_m1p1(r[0]), x = random_m1p1(r[1]), z = 0.f;
    return !(x * x + y * y + 0.f * 0.f >= 1.f);
}

template <int K>
__global__ void __launch_bounds__(256) k_count(const uint32_t* raw, size_t nCand, int* tileCounts)
{
    typedef cub::BlockReduce<int, 256> Reduce;
    __shared__ typename Reduce::TempStorage tmp;
    size_t cand = (size_t)blockIdx.x * 256 + threadIdx.x;
    float  x, y, z;
    int    a = cand < nCand ? (int)candidate<K>(raw, cand, x, y, z) : 0;
    int    sum = Reduce(tmp).Sum(a);
    if (threadIdx.x == 0) tileCounts[blockIdx.x] = sum;
}

// Writes the accepted samples of this window at their global ordinal (accBase + rank inside the window) while the
// ordinal is below `need`; the thread that writes ordinal need - 1 records where the stream stands afterwards.
template <int K>
__global__ void __launch_bounds__(256) k_compact(const uint32_t* raw, size_t nCand, const int* tileOffsets, const unsigned long long* accBase,
                                                 unsigned long long need, float* out, unsigned long long rawBase, unsigned long long* rawEnd)
{
    typedef cub::BlockScan<int, 256> Scan;
    __shared__ typename Scan::TempStorage tmp;
    size_t cand = (size_t)blockIdx.x * 256 + threadIdx.x;
    float  x = 0.f, y = 0.f, z = 0.f;
    int    a = cand < nCand ? (int)candidate<K>(raw, cand, x, y, z) : 0;
    int    rank;
    Scan(tmp).ExclusiveSum(a, rank);
    if (!a) return;
    unsigned long long g = *accBase + (unsigned long long)tileOffsets[blockIdx.x] + (unsigned long long)rank;
    if (g >= need) return;
    if (K == 3) reinterpret_cast<float4*>(out)[g] = make_float4(x, y, z, ssao_sample_scale(v3(x, y, z)));  // one 16-byte record per ball sample
    else out[g * 2] = x, out[g * 2 + 1] = y;
    if (g == need - 1) *rawEnd = rawBase + (unsigned long long)K * (cand + 1);
}
__global__ void k_acc_add(unsigned long long* acc, const int* total) { *acc += (unsigned long long)*total; }

// ---- PCSS chain ----------------------------------------------------------------------------------------------------

// shadow.cpp:33-34 compares (double)d < 0.001; float(0.001) = 0.00100000004749745 is the smallest float above 0.001, so
// d < 0.001f is the same predicate for every float d
__device__ __forceinline__ float fix_depth(float d) { return (d < 0.001f) ? 1.f : d; }

// separable box min / max of the fixed-up shadow map over the window [x + lo, x + hi] x [y + lo, y + hi]
// (lo = -r, hi = r: centred box of the search footprint; lo = 0, hi = bw - 1: box anchored at its first texel).
// Each 1-D pass is a log-step filter in shared memory: m_k[i] = min(a[i .. i + 2^k - 1]) by doubling, and a window of
// w = hi - lo + 1 samples is the min of two overlapping m_K windows, K = floor(log2 w) — K + 1 operations per sample
// instead of w.  Samples outside the map are +inf / -inf, i.e. left out, as a clamped loop would.
constexpr int kMMW = 768;   // H pass: outputs per CTA (one row); window <= 257
constexpr int kMMR = 128;   // V pass: output rows per CTA (32 columns); window <= 129

__device__ __forceinline__ float2 mm2(float2 a, float2 b) { return make_float2(fminf(a.x, b.x), fmaxf(a.y, b.y)); }

// The box maps are only ever read at the texels the band's pixels map to (k_classify, k_pixel_masks, k_chunk_index), which
// k_shadow_coords bounds by a texel rectangle (rect = x0, y0, x1, y1; x1 < x0: nobody reads).  Output tiles that cannot
// intersect the rectangle grown by `margin` texels are skipped; rect == nullptr: the whole map.
struct MapRegion
{
    const int* rect;
    int        margin;
};
__device__ __forceinline__ bool region_skips(const MapRegion& g, int W, int H, int tx0, int tx1, int ty0, int ty1, int rowLo, int rowHi)
{
    if (!g.rect) return false;
    const int x0 = __ldg(g.rect), y0 = __ldg(g.rect + 1), x1 = __ldg(g.rect + 2), y1 = __ldg(g.rect + 3);
    if (x1 < x0 || y1 < y0) return true;
    // rowLo / rowHi: extra rows the H pass has to provide for the V pass's window
    return tx1 < x0 - g.margin || tx0 > x1 + g.margin || ty1 < y0 - g.margin + rowLo || ty0 > y1 + g.margin + rowHi;
}

template <bool FIRST>
__global__ void __launch_bounds__(256) k_minmax_h(const float* imin, const float* imax, int W, int H, int lo, int hi, float* omin, float* omax, MapRegion region)
{
    __shared__ float2 t[kMMW + 256];
    const int   y = blockIdx.y, x0 = blockIdx.x * kMMW, w = hi - lo + 1, nOut = min(kMMW, W - x0), n = nOut + w - 1;
    if (region_skips(region, W, H, x0, x0 + nOut - 1, y, y, lo, hi)) return;
    const float inf = __int_as_float(0x7f800000);
    for (int e = threadIdx.x; e < n; e += 256)
    {
        int    gx = x0 + lo + e;
        float2 v = make_float2(inf, -inf);
        if (gx >= 0 && gx < W)
        {
            if (FIRST)
            {
                float d = fix_depth(__ldg(imin + (size_t)y * W + gx));
                v = make_float2(d, d);
            }
            else v = make_float2(__ldg(imin + (size_t)y * W + gx), __ldg(imax + (size_t)y * W + gx));
        }
        t[e] = v;
    }
    __syncthreads();
    const int K = 31 - __clz(w);
    for (int k = 0; k < K; ++k)
    {
        const int step = 1 << k;
        float2    v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
            int e = threadIdx.x + 256 * q;
            if (e < n) v[q] = e + step < n ? mm2(t[e], t[e + step]) : t[e];
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
            int e = threadIdx.x + 256 * q;
            if (e < n) t[e] = v[q];
        }
        __syncthreads();
    }
    for (int o = threadIdx.x; o < nOut; o += 256)
    {
        float2 r = mm2(t[o], t[o + w - (1 << K)]);
        omin[(size_t)y * W + x0 + o] = r.x, omax[(size_t)y * W + x0 + o] = r.y;
    }
}

// V pass: 32 columns x kMMR output rows per CTA, ping-pong tiles in dynamic shared memory (2 x (kMMR + w - 1) x 32 float2 <= 128 KB)
__global__ void __launch_bounds__(512) k_minmax_v(const float* imin, const float* imax, int W, int H, int lo, int hi, float* omin, float* omax, MapRegion region)
{
    extern __shared__ __align__(16) unsigned char mmRaw[];
    const int   lane = threadIdx.x & 31, wid = threadIdx.x >> 5;  // 16 warps
    const int   x = blockIdx.x * 32 + lane, y0 = blockIdx.y * kMMR, w = hi - lo + 1, nOut = min(kMMR, H - y0), n = nOut + w - 1;
    if (region_skips(region, W, H, blockIdx.x * 32, blockIdx.x * 32 + 31, y0, y0 + nOut - 1, 0, 0)) return;
    // two tiles of (kMMR + w - 1) rows x 32 columns (the launch sizes the allocation to the window, so two CTAs share an SM)
    float2(*t0)[32] = reinterpret_cast<float2(*)[32]>(mmRaw);
    float2(*t1)[32] = t0 + (kMMR + w - 1);
    float2(*t[2])[32] = { t0, t1 };
    const float inf = __int_as_float(0x7f800000);
    for (int e = wid; e < n; e += 16)
    {
        int    gy = y0 + lo + e;
        float2 v = make_float2(inf, -inf);
        if (gy >= 0 && gy < H && x < W) v = make_float2(__ldg(imin + (size_t)gy * W + x), __ldg(imax + (size_t)gy * W + x));
        t[0][e][lane] = v;
    }
    __syncthreads();
    const int K = 31 - __clz(w);
    int       cur = 0;
    for (int k = 0; k < K; ++k)
    {
        const int step = 1 << k;
        for (int e = wid; e < n; e += 16) t[cur ^ 1][e][lane] = e + step < n ? mm2(t[cur][e][lane], t[cur][e + step][lane]) : t[cur][e][lane];
        __syncthreads();
        cur ^= 1;
    }
    if (x < W)
        for (int o = wid; o < nOut; o += 16)
        {
            float2 r = mm2(t[cur][o][lane], t[cur][o + w - (1 << K)][lane]);
            omin[(size_t)(y0 + o) * W + x] = r.x, omax[(size_t)(y0 + o) * W + x] = r.y;
        }
}

// plain loops, for windows wider than the tiles above are sized for
__global__ void k_minmax_h_wide(const float* sm, int W, int H, int lo, int hi, float* omin, float* omax)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    float mn = __int_as_float(0x7f800000), mx = -mn;
    for (int i = max(0, x + lo); i <= min(W - 1, x + hi); ++i)
    {
        float d = fix_depth(__ldg(sm + (size_t)y * W + i));
        mn = fminf(mn, d), mx = fmaxf(mx, d);
    }
    omin[(size_t)y * W + x] = mn, omax[(size_t)y * W + x] = mx;
}
__global__ void k_minmax_v_wide(const float* imin, const float* imax, int W, int H, int lo, int hi, float* omin, float* omax)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    float mn = __int_as_float(0x7f800000), mx = -mn;
    for (int j = max(0, y + lo); j <= min(H - 1, y + hi); ++j)
    {
        mn = fminf(mn, __ldg(imin + (size_t)j * W + x)), mx = fmaxf(mx, __ldg(imax + (size_t)j * W + x));
    }
    omin[(size_t)y * W + x] = mn, omax[(size_t)y * W + x] = mx;
}

// ---- chunk signatures ------------------------------------------------------------------------------------------------
// The unit square [-1,1)^2 of the disk samples is cut into 8 x 8 cells at the exact thresholds -1 + k/4.
// sig[c] has bit (8 ky + kx) set iff chunk c (32 consecutive accepted disk samples) has a sample in cell (kx, ky).
// A constant of the stream, like the sample table itself.
__device__ __forceinline__ int cell_of(float x)
{
    return (x >= -0.75f) + (x >= -0.5f) + (x >= -0.25f) + (x >= 0.f) + (x >= 0.25f) + (x >= 0.5f) + (x >= 0.75f);
}
__global__ void __launch_bounds__(256) k_signatures(const float2* disk, size_t nChunks, unsigned long long* sig)
{
    size_t chunk = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int    lane = threadIdx.x & 31;
    if (chunk >= nChunks) return;
    float2   d = __ldg(disk + chunk * 32 + lane);
    int      bit = cell_of(d.y) * 8 + cell_of(d.x);
    unsigned lo = __reduce_or_sync(0xffffffffu, bit < 32 ? 1u << bit : 0u);
    unsigned hi = __reduce_or_sync(0xffffffffu, bit >= 32 ? 1u << (bit - 32) : 0u);
    if (lane == 0) sig[chunk] = (unsigned long long)lo | ((unsigned long long)hi << 32);
}

struct ChainPass
{
    int          W, H;  // frame
    const float *worldpos, *normal, *lndc;
    float        lightPos[3], biasSlope, biasMin;
    ShadowMapD   sm;
    const float *smMin, *smMax;
    int          r;
    float        fsF;  // (float)pcss filter size: |tap offset| <= fsF
};

// per pixel: shadow coordinate + bias (shadow.cpp:109-118) from the G-buffer (deferred lighting; forward mode
// computes the same quantities per fragment in raster.cu)
// texel rectangle of the sites that can look at the box maps at all: the centre texel exactly as k_classify computes it
struct TexelRect
{
    int x0, y0, x1, y1;
};
__device__ __forceinline__ void rect_add(const ChainPass& P, float scx, float scy, TexelRect& r)
{
    float ulo = scx + (-P.fsF), uhi = scx + P.fsF, vlo = scy + (-P.fsF), vhi = scy + P.fsF;
    if (!(uhi < 0.f || ulo > 1.f || vhi < 0.f || vlo > 1.f) && ulo == ulo && vlo == vlo)
    {
        int cx = f2i_x86((float)P.sm.iw * clampf(scx, 0.f, 1.f)), cy = f2i_x86((float)P.sm.ih * clampf(scy, 0.f, 1.f));
        r.x0 = min(r.x0, cx), r.x1 = max(r.x1, cx), r.y0 = min(r.y0, cy), r.y1 = max(r.y1, cy);
    }
}
__global__ void k_rect_init(int* rect) { rect[0] = rect[1] = 0x7fffffff, rect[2] = rect[3] = -1; }
// texels of the rectangle grown by the widest margin the box maps are built with (byte model of the min / max passes)
__global__ void k_rect_area(const int* rect, int margin, int W, int H, unsigned* area)
{
    int x0 = max(0, rect[0] - margin), y0 = max(0, rect[1] - margin), x1 = min(W - 1, rect[2] + margin), y1 = min(H - 1, rect[3] + margin);
    *area = (rect[2] < rect[0] || x1 < x0 || y1 < y0) ? 0u : (unsigned)(x1 - x0 + 1) * (unsigned)(y1 - y0 + 1);
}

// grid-stride: every thread keeps its rectangle in registers, a warp merges once at the end (four atomics per warp of the
// whole launch — per-pixel atomics on four addresses would serialise in the L2)
__global__ void __launch_bounds__(256) k_shadow_coords(ChainPass P, size_t first, size_t last, float4* sc4, int* rect)
{
    const size_t n = (size_t)P.W * P.H, stride = (size_t)gridDim.x * blockDim.x;
    TexelRect    r;
    r.x0 = r.y0 = 0x7fffffff, r.x1 = r.y1 = -1;
    for (size_t idx = first + (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < last; idx += stride)
    {
        V3    pos = v3(P.worldpos[idx], P.worldpos[n + idx], P.worldpos[2 * n + idx]);
        V3    nrm = v3(P.normal[idx], P.normal[n + idx], P.normal[2 * n + idx]);
        V3    ln = v3(P.lndc[idx], P.lndc[n + idx], P.lndc[2 * n + idx]);
        V3    lightDir = vnormalize(vsub(v3(P.lightPos[0], P.lightPos[1], P.lightPos[2]), pos));
        V3    sc = vadd(vscale(ln, 0.5f), v3(0.5f, 0.5f, 0.5f));
        float bias = fmaxf(P.biasSlope * (1.f - vdot(nrm, lightDir)), P.biasMin);
        sc4[idx] = make_float4(sc.x, sc.y, sc.z, bias);
        if (rect) rect_add(P, sc.x, sc.y, r);
    }
    if (!rect) return;
    r.x0 = __reduce_min_sync(0xffffffffu, r.x0), r.y0 = __reduce_min_sync(0xffffffffu, r.y0);
    r.x1 = __reduce_max_sync(0xffffffffu, r.x1), r.y1 = __reduce_max_sync(0xffffffffu, r.y1);
    if ((threadIdx.x & 31) == 0 && r.x1 >= r.x0)
    {
        atomicMin(rect, r.x0), atomicMin(rect + 1, r.y0);
        atomicMax(rect + 2, r.x1), atomicMax(rect + 3, r.y1);
    }
}

// blocker-search class of every site (a site = one consumer of the lighting-phase stream, in consumption order:
// a pixel in deferred mode, a depth-test-passing fragment in forward mode)
__global__ void __launch_bounds__(256) k_classify(ChainPass P, size_t n, const float4* sc4, int* isU, int* isC1)
{
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    float4 s4 = sc4[idx];
    V3     sc = v3(s4.x, s4.y, s4.z);
    float  bias = s4.w;

    // every tap lands at u in [ulo, uhi] (the float add is monotone in the offset, |offset| <= fsF)
    float ulo = sc.x + (-P.fsF), uhi = sc.x + P.fsF, vlo = sc.y + (-P.fsF), vhi = sc.y + P.fsF;
    int   cls = 2;
    if (uhi < 0.f || ulo > 1.f || vhi < 0.f || vlo > 1.f) cls = 0;  // all taps outside the map: +inf, never a blocker
    else if (ulo == ulo && vlo == vlo && sc.z == sc.z)
    {
        bool  mayOut = ulo < 0.f || uhi > 1.f || vlo < 0.f || vhi > 1.f;
        float cu = clampf(sc.x, 0.f, 1.f), cv = clampf(sc.y, 0.f, 1.f);
        int   cx = f2i_x86((float)P.sm.iw * cu), cy = f2i_x86((float)P.sm.ih * cv);
        int   x0 = f2i_x86((float)P.sm.iw * fmaxf(ulo, 0.f)), x1 = f2i_x86((float)P.sm.iw * fminf(uhi, 1.f));
        int   y0 = f2i_x86((float)P.sm.ih * fmaxf(vlo, 0.f)), y1 = f2i_x86((float)P.sm.ih * fminf(vhi, 1.f));
        if (x0 >= cx - P.r && x1 <= cx + P.r && y0 >= cy - P.r && y1 <= cy + P.r && cx >= 0 && cx < P.sm.w && cy >= 0 && cy < P.sm.h)
        {
            float dmin = P.smMin[(size_t)cy * P.sm.w + cx], dmax = P.smMax[(size_t)cy * P.sm.w + cx];
            if (!mayOut && sc.z > dmax + bias) cls = 1;       // every tap blocks
            else if (!(sc.z > dmin + bias)) cls = 0;           // no tap can block
        }
    }
    isU[idx] = cls == 2, isC1[idx] = cls == 1;
}

// uncertain pixels, in scan order: pixel index, number of certain blockers before it, shadow coordinate + bias
__global__ void __launch_bounds__(256) k_gather_uncertain(size_t n, const int* isU, const int* posU, const int* c1pre, const float4* sc4, unsigned* Upix,
                                                          unsigned* Uc1, float4* Usc)
{
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n || !isU[idx]) return;
    int j = posU[idx];
    Upix[j] = (unsigned)idx, Uc1[j] = (unsigned)c1pre[idx], Usc[j] = sc4[idx];
}

// Per uncertain pixel: F = cells whose every tap certainly blocks, E = cells in which a tap can block at all.
// A chunk with a sample in an F cell has a blocker; a chunk with no sample in any E cell has none; only the rest
// needs its 32 taps evaluated.  Cell (kx, ky) covers sample coordinates [-1 + kx/4, -1 + (kx+1)/4]: the tap
// coordinate u = sc.x + float(x * fs) and its texel index are monotone in x, so the cell maps into the texel rectangle
// spanned by its two threshold columns / rows; box min / max maps anchored at the rectangle's first texel (a
// superset of the rectangle) give conservative answers.  One warp per pixel, two cells per lane.
struct MaskPass
{
    ShadowMapD   sm;
    const float *boxMin, *boxMax;
    int          bw;
    double       fs;
};
__global__ void __launch_bounds__(256) k_pixel_masks(MaskPass P, int nU, const float4* Usc, unsigned long long* UF, unsigned long long* UE)
{
    int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (j >= nU) return;
    float4   s = Usc[j];
    unsigned fbits[2], ebits[2];
#pragma unroll
    for (int h = 0; h < 2; ++h)
    {
        int   cell = lane + 32 * h, kx = cell & 7, ky = cell >> 3;
        // (the boundary offsets are recomputed per cell: a table of the nine values, in shared memory or in the parameter bank,
        // measured 60 % slower — the FP64 multiplies run beside the FP32 work)
        float tx0 = -1.f + 0.25f * (float)kx, tx1 = -1.f + 0.25f * (float)(kx + 1);
        float ty0 = -1.f + 0.25f * (float)ky, ty1 = -1.f + 0.25f * (float)(ky + 1);
        float u0 = s.x + (float)((double)tx0 * P.fs), u1 = s.x + (float)((double)tx1 * P.fs);
        float v0 = s.y + (float)((double)ty0 * P.fs), v1 = s.y + (float)((double)ty1 * P.fs);
        bool  none = u1 < 0.f || u0 > 1.f || v1 < 0.f || v0 > 1.f || !(u0 == u0) || !(v0 == v0);
        bool  full = u0 >= 0.f && u1 <= 1.f && v0 >= 0.f && v1 <= 1.f;
        bool  F = false, E = false;
        if (!none)
        {
            int x0 = f2i_x86((float)P.sm.iw * fmaxf(u0, 0.f)), x1 = f2i_x86((float)P.sm.iw * fminf(u1, 1.f));
            int y0 = f2i_x86((float)P.sm.ih * fmaxf(v0, 0.f)), y1 = f2i_x86((float)P.sm.ih * fminf(v1, 1.f));
            if (x1 - x0 + 1 > P.bw || y1 - y0 + 1 > P.bw || x0 < 0 || y0 < 0 || x0 >= P.sm.w || y0 >= P.sm.h) E = true;  // cannot bound: ambiguous
            else
            {
                float dmin = __ldg(P.boxMin + (size_t)y0 * P.sm.w + x0), dmax = __ldg(P.boxMax + (size_t)y0 * P.sm.w + x0);
                E = s.z > dmin + s.w;
                F = full && (s.z > dmax + s.w);
            }
        }
        fbits[h] = __ballot_sync(0xffffffffu, F), ebits[h] = __ballot_sync(0xffffffffu, E);
    }
    if (lane == 0)
    {
        UF[j] = (unsigned long long)fbits[0] | ((unsigned long long)fbits[1] << 32);
        UE[j] = (unsigned long long)ebits[0] | ((unsigned long long)ebits[1] << 32);
    }
}

// ---- the chain proper --------------------------------------------------------------------------------------------
// Row j of the uncertain list (pixel p_j, c1_j certain blockers before it) uses chunk p_j + 2 (c1_j + m_j), where
// m_j = number of uncertain rows before j that have a blocker.  m_j is only known once all earlier rows are, but
// its value can be predicted well: blocker flags evaluated at ANY nearby offset are draws from the same
// distribution, so the prefix sums of a "pilot" evaluation (every row at a crude offset, fully parallel) track the
// true m_j to within a random-walk drift.  Rows are then processed in super-chunks of up to kT rows from an exactly
// known state (j0, m0): each row is evaluated for the kWin candidates around its predicted offset (parallel), a
// single thread walks the rows whose flag depends on the candidate, and the first row whose true offset falls
// outside its window ends the super-chunk there (the next one restarts from that row with exact state), so the
// result never depends on the prediction — only the speed does.
struct ChainRows
{
    int                       nU;
    const unsigned *          Upix, *Uc1;
    const float4*             Usc;
    const unsigned long long *UF, *UE, *sig;
    ShadowMapD                sm;
    const float2*             disk;
    double                    fs;
    unsigned long long        base;  // chunk of site i with k earlier blockers = base + i + 2 k
    const unsigned long long* kDev;   // device-side hand-off: blockers found by the bands above (added to base as 2 k); NULL = none
    unsigned long long*       stats;  // diagnostics (FGL_CHAIN_STATS=1): pairs, decided-one, ambiguous, taps in E\F cells, warp tasks
};

// flags of up to 32 (row, chunk) pairs, one per lane: signature tests first, then the warp evaluates the ambiguous
// pairs tap by tap (four at a time: their sample loads and shadow-map gathers are independent).  Returns the ballot.
__device__ __forceinline__ uint32_t eval_pairs(const ChainRows& R, bool valid, int j, size_t chunk, int lane)
{
    unsigned long long sg = valid ? __ldg(R.sig + chunk) : 0ull;
    unsigned long long F = valid ? __ldg(R.UF + j) : 0ull, E = valid ? __ldg(R.UE + j) : 0ull;
    bool               one = (sg & F) != 0ull;
    uint32_t           w = __ballot_sync(0xffffffffu, valid && one);
    uint32_t           amb = __ballot_sync(0xffffffffu, valid && !one && (sg & E) != 0ull);
    if (R.stats)
    {
        uint32_t nv = __ballot_sync(0xffffffffu, valid);
        if (lane == 0)
        {
            atomicAdd(R.stats + 0, (unsigned long long)__popc(nv)), atomicAdd(R.stats + 1, (unsigned long long)__popc(w));
            atomicAdd(R.stats + 2, (unsigned long long)__popc(amb)), atomicAdd(R.stats + 4, 1ull);
        }
        for (uint32_t a = amb; a; a &= a - 1)
        {
            int                b = __ffs(a) - 1;
            size_t             cq = __shfl_sync(0xffffffffu, chunk, b);
            unsigned long long Fq = __shfl_sync(0xffffffffu, F, b), Eq = __shfl_sync(0xffffffffu, E, b);
            float2             d = __ldg(R.disk + cq * 32 + lane);
            int                bit = cell_of(d.y) * 8 + cell_of(d.x);
            uint32_t           hits = __ballot_sync(0xffffffffu, ((Eq & ~Fq) >> bit) & 1ull);
            if (lane == 0) atomicAdd(R.stats + 3, (unsigned long long)__popc(hits));
        }
    }
    while (amb)
    {
        int  b[4];
        bool v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
            v[q] = amb != 0;
            b[q] = v[q] ? __ffs(amb) - 1 : 0;
            amb &= amb - 1;
        }
        float4 s[4];
        float2 d[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
            int    jq = __shfl_sync(0xffffffffu, j, b[q]);
            size_t cq = __shfl_sync(0xffffffffu, chunk, b[q]);
            s[q] = v[q] ? __ldg(R.Usc + jq) : make_float4(0.f, 0.f, 0.f, 0.f);
            d[q] = v[q] ? __ldg(R.disk + cq * 32 + lane) : make_float2(0.f, 0.f);
        }
        float sd[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
            float ox = (float)((double)d[q].x * R.fs), oy = (float)((double)d[q].y * R.fs);
            sd[q] = v[q] ? shadow_lookup(R.sm, s[q].x + ox, s[q].y + oy) : __int_as_float(0x7f800000);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (__any_sync(0xffffffffu, s[q].z > sd[q] + s[q].w)) w |= 1u << b[q];
    }
    return w;
}

// pilot: every uncertain row at K crude offsets around "half of the uncertain rows before it have a blocker"; pilot[j] = how
// many of the K chunks had a blocker.  The prediction of the chain is the prefix sum of pilot / K: the error of a K-sample
// mean has the variance p (1 - p) / K per row, on top of the p (1 - p) of the row's own flag — the drift that ends a
// super-chunk grows (1 + 1 / K) / 2 times as fast as with a single sample.
__global__ void __launch_bounds__(256) k_chain_pilot(ChainRows R, int* pilot, int K)
{
    int  lane = threadIdx.x & 31;
    int  j = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = j < R.nU;
    size_t chunk = valid ? (size_t)R.base + R.Upix[j] + 2 * ((size_t)R.Uc1[j] + (size_t)(j >> 1)) : 0;
    int    cnt = 0;
    for (int q = 0; q < K; ++q)
    {
        uint32_t w = eval_pairs(R, valid, j, chunk + 2 * (size_t)q, lane);
        cnt += (int)((w >> lane) & 1u);
    }
    if (valid) pilot[j] = cnt;
}

enum { CH_J0 = 0, CH_M0 = 1, CH_DONE = 2, CH_ITERS = 3, CH_ARRIVE = 4, CH_EPOCH = 5, CH_ERROR = 6, CH_ARRIVE_ROWS = 7, CH_NBLOCKERS = 8, CH_NFILTERED = 9,
       CH_PACKED = 12 /* 64-bit, 8-byte aligned: the state the CTAs wait for, in one word (see chain_pack) */ };

// ---- the chain as ONE persistent cooperative kernel ----------------------------------------------------------------
// A super-chunk is cut into segments of kSeg = 32 rows.  Per iteration (one super-chunk from an exactly known state):
//   (1) every CTA evaluates whole segments: warp = row, lane = candidate offset of the row's window (signature test,
//       then the taps of the ambiguous candidates — only the taps that fall into a cell in which a texel can block);
//   (2) the same CTA turns the 32 x W window bits into the segment's transfer table: for every offset d the segment
//       can be entered with, the offset it is left with (G8) and the 32 blocker flags on the way (GM).  W threads
//       walk 32 rows each, out of shared memory;
//   (3) all CTAs arrive on a counter; CTA 0 composes the segment tables in order (32 warps compose groups of
//       segments for every entry, one thread chains the groups, the warps re-walk their group from its now known
//       entry), commits the longest prefix of segments whose windows contained the true offset, publishes the new
//       state (epoch flag; the other CTAs spin on it) and writes the committed rows' flags while the next iteration
//       is already being evaluated (tables are double-buffered by iteration parity).
// Nothing here depends on the prediction being right: a segment entered outside its window ends the super-chunk in
// front of it, and segment 0 is always valid (its rows' windows start at offset 0 and the offset is < 32).
constexpr int kSeg = 32;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// The state a super-chunk starts from travels from CTA 0 to the other CTAs as ONE 64-bit word, so that the load that sees the
// new epoch also carries the state (three more dependent L2 round trips per iteration otherwise):
//   bits 0..21 j0 / 32 (j0 is a multiple of the segment size until the chain is done), 22..48 m0, 49 done, 50..63 epoch mod 2^14.
// A waiter knows exactly which epoch comes next, so the truncated epoch is compared for equality.
constexpr int      kChainMaxRowsLog2 = 27;  // launch_chain refuses more uncertain rows than the fields hold
constexpr unsigned kEpochMask = 0x3fffu;
__device__ __forceinline__ unsigned long long chain_pack(unsigned j0, unsigned m0, bool done, unsigned epoch)
{
    return (unsigned long long)(j0 >> 5) | ((unsigned long long)m0 << 22) | ((unsigned long long)(done ? 1u : 0u) << 49) |
           ((unsigned long long)(epoch & kEpochMask) << 50);
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) { asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
// spins until the packed state carries `epoch`; gives up (returns false) after ~2 s
__device__ __forceinline__ bool spin_packed(const unsigned long long* p, unsigned epoch, unsigned long long& w)
{
    const long long t0 = clock64();
    while ((unsigned)((w = ld_acquire_u64(p)) >> 50) != (epoch & kEpochMask))
    {
        __nanosleep(20);
        if (clock64() - t0 > 4000000000LL) return false;
    }
    return true;
}

// spins until *p >= want; gives up (returns false) after ~2 s so a broken launch can never hang the device
__device__ __forceinline__ bool spin_until(const unsigned* p, unsigned want)
{
    const long long t0 = clock64();
    while (ld_acquire_u32(p) < want)
    {
        __nanosleep(20);
        if (clock64() - t0 > 4000000000LL) return false;
    }
    return true;
}

// position of the k-th (zero-based) set bit of m; m has more than k bits set
__device__ __forceinline__ int nth_set_bit(uint32_t m, int k)
{
    int pos = 0, c;
    c = __popc(m & 0xffffu);
    if (k >= c) k -= c, pos += 16, m >>= 16;
    c = __popc(m & 0xffu);
    if (k >= c) k -= c, pos += 8, m >>= 8;
    c = __popc(m & 0xfu);
    if (k >= c) k -= c, pos += 4, m >>= 4;
    c = __popc(m & 0x3u);
    if (k >= c) k -= c, pos += 2, m >>= 2;
    if (k >= (int)(m & 1u)) pos += 1;
    return pos;
}

constexpr int kMaxGrid = 160;              // CTAs (= SMs) the tables are sized for; a B200 has 148
constexpr int kMaxSpc = 6;                 // segments one CTA evaluates per iteration (FGL_CHAIN_SEGS caps the iteration below grid * kMaxSpc)
constexpr int kRowBatch = 3;               // rows a warp keeps in flight during the signature tests
constexpr int kRowsCta = kMaxSpc * kSeg;  // rows of those segments

// Shared memory of k_chain_fused (dynamic; CTA 0 also holds the tables of the whole iteration)
template <int NW>
struct ChainSmem
{
    static constexpr int W = 32 * NW;
    static constexpr int kMaxSegs = kMaxGrid * kMaxSpc;
    // evaluation (every CTA)
    float4             rowS[kRowsCta];    // shadow coordinate + bias of the row
    unsigned long long rowE[kRowsCta];    // cells in which a tap can block
    unsigned long long rowChunk0[kRowsCta];
    uint32_t           bits[kRowsCta][NW];  // window bits: bit l of word h = blocker flag of candidate offset lo + 32 h + l
    uint32_t           amb[kRowsCta * NW];  // candidates the signature test left undecided
    int                pre[kRowsCta * NW + 1];
    int                lo[kRowsCta];
    uint16_t           queue[kRowsCta * W];  // work list of the undecided candidates: (word << 5) | bit
    uint4              pairQ[32][8];      // per warp: the eight candidates of the current step (chunk lo / hi, work-list word, bit)
    // composition (CTA 0)
    uint8_t G[kMaxSegs * W];
    int     segLo[kMaxSegs], entry[kMaxSegs], after[kMaxSegs];
    short   GT[32 * W];
    int     groupD[33];
    int     ctl[8];
};

template <int NW>
__device__ __forceinline__ void chain_fused_body(const ChainRows& R, unsigned* state, const int* Ppre, uint8_t* G8all, uint32_t* GMall, int* segLoAll,
                                                 uint32_t* rowBits, int* rowLo, uint8_t* flagU, int segsPerIter, int pilotK)
{
    constexpr int W = 32 * NW;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    ChainSmem<NW>& S = *reinterpret_cast<ChainSmem<NW>*>(smemRaw);
    uint8_t* const sG = S.G;
    int* const     sSegLo = S.segLo;
    int* const     sEntry = S.entry;
    int* const     sAfter = S.after;
    short* const   sGT = S.GT;
    int* const     sGroupD = S.groupD;
    int* const     sCtl = S.ctl;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (unsigned it = 0;; ++it)
    {
        if (threadIdx.x == 0)
        {
            if (it == 0) sCtl[0] = 0, sCtl[1] = 0, sCtl[2] = 0;  // (the host zeroed the state)
            else if (blockIdx.x != 0)
            {   // CTA 0 left the next state in its own sCtl when it published it
                unsigned long long w = 0ull;
                const bool         ok = spin_packed(reinterpret_cast<const unsigned long long*>(state + CH_PACKED), it, w);
                sCtl[0] = (int)(((unsigned)w & 0x3fffffu) << 5), sCtl[1] = (int)((unsigned)(w >> 22) & 0x7ffffffu);
                sCtl[2] = ok ? (int)((w >> 49) & 1ull) : 1;
                if (!ok) atomicExch(state + CH_ERROR, 1u);
            }
        }
        __syncthreads();
        if (sCtl[2]) return;
        const int      j0 = sCtl[0];
        const unsigned m0 = (unsigned)sCtl[1];
        const int      nLive = min(segsPerIter * kSeg, R.nU - j0);
        const int      nSeg = (nLive + kSeg - 1) / kSeg;
        uint8_t*       G8 = G8all + (size_t)(it & 1) * segsPerIter * W;
        uint32_t*      GM = GMall + (size_t)(it & 1) * segsPerIter * W;
        int*           segLo = segLoAll + (size_t)(it & 1) * segsPerIter;
        const int      pre0 = __ldg(Ppre + j0);
        const size_t   kBefore = R.kDev ? (size_t)*R.kDev : 0;  // written by k_peer_wait before this kernel started
        // Evaluation is balanced over the grid by interleaving ROWS: CTA b takes rows b, b + G, b + 2 G, ... (hard rows
        // come in runs along shadow edges; whole segments per CTA left most SMs waiting for a few).  Its local row
        // rr = 32 i + warp is super-chunk row t = (32 i + warp) * G + b.  Tables are per SEGMENT: CTA b builds the tables
        // of segments b, b + G, ... from the rows' bits in global memory after a grid-wide barrier.
        const int mySegs = blockIdx.x < nSeg ? (nSeg - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
        constexpr int myWords = kRowsCta * NW;

        // (1a) warp per row (kMaxSpc rows per warp, kRowBatch of them in flight together): signature test of the W candidates
        static_assert(kMaxSpc % kRowBatch == 0, "rows are processed in whole batches");
        for (int ib = 0; ib < kMaxSpc; ib += kRowBatch)
        {
            unsigned           upix[kRowBatch], uc1[kRowBatch];
            unsigned long long F[kRowBatch], E[kRowBatch];
            float4             sc[kRowBatch];
            int                lo[kRowBatch], tt[kRowBatch];
            bool               rowValid[kRowBatch];
#pragma unroll
            for (int i = 0; i < kRowBatch; ++i)
            {
                tt[i] = (warp + 32 * (ib + i)) * (int)gridDim.x + (int)blockIdx.x;
                rowValid[i] = tt[i] < nLive;
                upix[i] = uc1[i] = 0u, F[i] = E[i] = 0ull, sc[i] = make_float4(0.f, 0.f, 0.f, 0.f), lo[i] = 0;
                if (rowValid[i])
                {
                    const int j = j0 + tt[i];
                    lo[i] = __ldg(Ppre + j);
                    upix[i] = __ldg(R.Upix + j), uc1[i] = __ldg(R.Uc1 + j), F[i] = __ldg(R.UF + j), E[i] = __ldg(R.UE + j), sc[i] = __ldg(R.Usc + j);
                }
            }
            unsigned long long sg[kRowBatch][NW];
            size_t             chunk0[kRowBatch];
#pragma unroll
            for (int i = 0; i < kRowBatch; ++i)
            {
                lo[i] = rowValid[i] ? max(0, (lo[i] - pre0 + (pilotK >> 1)) / pilotK - W / 2) : 0;  // Ppre counts K samples per row
                chunk0[i] = (size_t)R.base + 2 * kBefore + upix[i] + 2 * ((size_t)uc1[i] + m0 + (size_t)lo[i]);
#pragma unroll
                for (int h = 0; h < NW; ++h)
                {
                    int  d = lo[i] + 32 * h + lane;
                    bool valid = rowValid[i] && d <= tt[i];
                    sg[i][h] = valid ? __ldg(R.sig + chunk0[i] + 2 * (size_t)(32 * h + lane)) : 0ull;
                }
            }
#pragma unroll
            for (int i = 0; i < kRowBatch; ++i)
            {
                const int rr = warp + 32 * (ib + i);
#pragma unroll
                for (int h = 0; h < NW; ++h)
                {
                    bool     one = (sg[i][h] & F[i]) != 0ull;
                    uint32_t wOne = __ballot_sync(0xffffffffu, one);
                    uint32_t wAmb = __ballot_sync(0xffffffffu, !one && (sg[i][h] & E[i]) != 0ull);
                    if (lane == 0) S.bits[rr][h] = wOne, S.amb[rr * NW + h] = wAmb;
                }
                if (lane == 0) S.lo[rr] = lo[i], S.rowS[rr] = sc[i], S.rowE[rr] = E[i], S.rowChunk0[rr] = (unsigned long long)chunk0[i];
            }
        }
        __syncthreads();
        // (1b) work list of the undecided candidates of all rows of this CTA: exclusive prefix of their counts
        if (warp == 0)
        {
            constexpr int kPer = (kRowsCta * NW + 31) / 32;
            int           cnt[kPer], sum = 0;
#pragma unroll
            for (int k = 0; k < kPer; ++k)
            {
                int wi = lane * kPer + k;
                cnt[k] = wi < myWords ? __popc(S.amb[wi]) : 0;
                sum += cnt[k];
            }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            int run = incl - sum;
#pragma unroll
            for (int k = 0; k < kPer; ++k)
            {
                int wi = lane * kPer + k;
                if (wi < myWords) S.pre[wi] = run;
                run += cnt[k];
            }
            if (lane == 31) S.pre[myWords] = incl;
        }
        __syncthreads();
        // (1b') the work list itself: entry pre[w] + rank of bit b inside amb[w]  =  (w << 5) | b, written by the lane that owns
        // the bit (warp per word) — the candidates are then addressed directly instead of searched for in the prefix array
        for (int wi = warp; wi < myWords; wi += 32)
        {
            const uint32_t a = S.amb[wi];
            if ((a >> lane) & 1u) S.queue[S.pre[wi] + __popc(a & ((1u << lane) - 1u))] = (uint16_t)((wi << 5) | lane);
        }
        __syncthreads();
        // (1c) the undecided candidates, eight per warp step.  Lanes 0..7 take one candidate each off the work list (word,
        // bit inside the word, sample chunk).  Then, lane = tap: which of the candidate's 32 taps fall into a cell of E
        // (an undecided candidate has no sample in a cell of F; a tap outside E cannot block) — about 5 of 32.  Those taps
        // of all eight candidates are finally evaluated DENSELY, one tap per lane (the sample comes back from L1).
        {
            const int total = S.pre[myWords];
            for (int g0 = warp * 8; g0 < total; g0 += 32 * 8)
            {
                const int nq = min(8, total - g0);
                if (lane < nq)
                {
                    const unsigned e = S.queue[g0 + lane];
                    const int      loI = (int)(e >> 5);
                    const int                b = (int)(e & 31u);
                    const unsigned long long chunk = S.rowChunk0[loI / NW] + 2ull * (unsigned)(32 * (loI % NW) + b);
                    S.pairQ[warp][lane] = make_uint4((unsigned)chunk, (unsigned)(chunk >> 32), (unsigned)loI, (unsigned)b);
                }
                __syncwarp();
                // which taps fall into a cell of E: 8 lanes per candidate, 4 taps (two 16-byte loads) per lane, 4 candidates per
                // step.  Word 4 st + j, bit 8 a + t  <->  candidate 4 st + a, tap 4 t + j.
                uint32_t hm[8];
                int      cum[9];
                cum[0] = 0;
#pragma unroll
                for (int st = 0; st < 2; ++st)
                {
                    const int q = 4 * st + (lane >> 3), t = lane & 7;
                    bool      hit[4] = { false, false, false, false };
                    if (q < nq)
                    {
                        const uint4              pq = S.pairQ[warp][q];
                        const size_t             chunk = (size_t)pq.x | ((size_t)pq.y << 32);
                        const float4*            src = reinterpret_cast<const float4*>(R.disk + chunk * 32 + 4 * t);
                        const float4             s01 = __ldg(src), s23 = __ldg(src + 1);
                        const unsigned long long E = S.rowE[pq.z / NW];
                        // cell = cell_of(y) * 8 + cell_of(x): scaling by 4 is exact
                        hit[0] = (E >> ((__float2int_rd(s01.y * 4.f) + 4) * 8 + (__float2int_rd(s01.x * 4.f) + 4))) & 1ull;
                        hit[1] = (E >> ((__float2int_rd(s01.w * 4.f) + 4) * 8 + (__float2int_rd(s01.z * 4.f) + 4))) & 1ull;
                        hit[2] = (E >> ((__float2int_rd(s23.y * 4.f) + 4) * 8 + (__float2int_rd(s23.x * 4.f) + 4))) & 1ull;
                        hit[3] = (E >> ((__float2int_rd(s23.w * 4.f) + 4) * 8 + (__float2int_rd(s23.z * 4.f) + 4))) & 1ull;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                    {
                        hm[4 * st + j] = __ballot_sync(0xffffffffu, hit[j]);
                        cum[4 * st + j + 1] = cum[4 * st + j] + __popc(hm[4 * st + j]);
                    }
                }
                for (int base = 0; base < cum[8]; base += 32)
                {
                    const int hi = base + lane;
                    if (hi < cum[8])
                    {
                        int      wq = 0, c0 = 0;
                        uint32_t m = hm[0];
#pragma unroll
                        for (int jq = 1; jq < 8; ++jq)
                            if (cum[jq] <= hi) wq = jq, c0 = cum[jq], m = hm[jq];
                        const int    bit = nth_set_bit(m, hi - c0);
                        const int    q = 4 * (wq >> 2) + (bit >> 3), tap = 4 * (bit & 7) + (wq & 3);
                        const uint4  pq = S.pairQ[warp][q];
                        const size_t chunk = (size_t)pq.x | ((size_t)pq.y << 32);
                        const float2 d = __ldg(R.disk + chunk * 32 + tap);
                        const float4 sc = S.rowS[pq.z / NW];
                        const float  ox = (float)((double)d.x * R.fs), oy = (float)((double)d.y * R.fs);
                        if (sc.z > shadow_lookup(R.sm, sc.x + ox, sc.y + oy) + sc.w) atomicOr(&S.bits[pq.z / NW][pq.z % NW], 1u << pq.w);
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // window bits and origins of this CTA's rows -> global, then every CTA waits for all rows of the iteration
        for (int rr = threadIdx.x; rr < kRowsCta; rr += blockDim.x)
        {
            const int t = rr * (int)gridDim.x + (int)blockIdx.x;
            if (t < nLive)
            {
#pragma unroll
                for (int h = 0; h < NW; ++h) rowBits[(size_t)t * NW + h] = S.bits[rr][h];
                rowLo[t] = S.lo[rr];
            }
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0)
        {
            red_release_add_u32(state + CH_ARRIVE_ROWS, 1u);
            if (!spin_until(state + CH_ARRIVE_ROWS, gridDim.x * (it + 1))) atomicExch(state + CH_ERROR, 4u), sCtl[2] = 1;
        }
        __syncthreads();
        if (sCtl[2]) return;  // (a timed-out barrier: every CTA gives up by itself)
        // (2) transfer tables of segments blockIdx.x + si * gridDim.x: their rows come back from L2, W threads per segment walk them
        for (int e = threadIdx.x; e < mySegs * kSeg; e += blockDim.x)
        {
            const int si = e / kSeg, t = (blockIdx.x + si * gridDim.x) * kSeg + (e % kSeg);
            const bool live = t < nLive;
#pragma unroll
            for (int h = 0; h < NW; ++h) S.bits[e][h] = live ? __ldcg(rowBits + (size_t)t * NW + h) : 0u;
            S.lo[e] = live ? __ldcg(rowLo + t) : 0;
        }
        __syncthreads();
        if ((int)threadIdx.x < mySegs * W)
        {
            const int si = threadIdx.x / W, c = threadIdx.x % W;
            const int sgi = blockIdx.x + si * gridDim.x;
            const int rows = min(kSeg, nLive - sgi * kSeg), d0 = S.lo[si * kSeg] + c;
            int       d = d0;
            uint32_t  mask = 0u;
            bool      ok = true;
            for (int r = 0; r < rows; ++r)
            {
                int idx = d - S.lo[si * kSeg + r];
                if (idx < 0 || idx >= W || d > sgi * kSeg + r)
                {
                    ok = false;
                    break;
                }
                uint32_t bit = (S.bits[si * kSeg + r][idx >> 5] >> (idx & 31)) & 1u;
                mask |= bit << r, d += (int)bit;
            }
            G8[(size_t)sgi * W + c] = ok ? (uint8_t)(d - d0) : (uint8_t)255;
            GM[(size_t)sgi * W + c] = mask;
            if (c == 0) segLo[sgi] = S.lo[si * kSeg];
        }
        // (3) arrive
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) red_release_add_u32(state + CH_ARRIVE, 1u);
        if (blockIdx.x != 0) continue;

        if (threadIdx.x == 0)
        {
            bool ok = spin_until(state + CH_ARRIVE, gridDim.x * (it + 1));
            sCtl[3] = ok ? 0 : 1;
        }
        __syncthreads();
        if (sCtl[3])
        {
            if (threadIdx.x == 0)
            {
                atomicExch(state + CH_ERROR, 2u), atomicExch(state + CH_DONE, 1u);
                st_release_u64(reinterpret_cast<unsigned long long*>(state + CH_PACKED), chain_pack(0u, 0u, true, it + 1));
            }
            return;
        }
        for (int i = threadIdx.x; i < nSeg * W / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sG)[i] = __ldcg(reinterpret_cast<const uint32_t*>(G8) + i);
        for (int i = threadIdx.x; i < nSeg; i += blockDim.x) sSegLo[i] = __ldcg(segLo + i);
        __syncthreads();
        // compose: warp g owns segments [g * gs, (g + 1) * gs)
        const int gs = (nSeg + 31) / 32, sBeg = warp * gs, sEnd = min(nSeg, sBeg + gs);
        if (sBeg < nSeg)
        {
#pragma unroll
            for (int k = 0; k < NW; ++k)
            {
                const int c = lane + 32 * k, d0 = sSegLo[sBeg] + c;
                int       d = d0;
                bool      ok = true;
                for (int sgi = sBeg; sgi < sEnd; ++sgi)
                {
                    int cc = d - sSegLo[sgi];
                    if (cc < 0 || cc >= W)
                    {
                        ok = false;
                        break;
                    }
                    unsigned g = sG[sgi * W + cc];
                    if (g == 255u)
                    {
                        ok = false;
                        break;
                    }
                    d += (int)g;
                }
                sGT[warp * W + c] = ok ? (short)(d - d0) : (short)-1;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0)
        {   // chain the groups: entry offset of every group, -1 = not reached
            int  d = 0;
            bool alive = true;
            for (int g = 0; g < 32; ++g)
            {
                int first = g * gs;
                sGroupD[g] = (alive && first < nSeg) ? d : -1;
                if (!alive || first >= nSeg) continue;
                int cc = d - sSegLo[first];
                int v = (cc >= 0 && cc < W) ? (int)sGT[g * W + cc] : -1;
                if (v < 0) alive = false;  // the first invalid segment lies in this group: its warp finds it below
                else d += v;
            }
            sCtl[4] = nSeg;  // first invalid segment
        }
        __syncthreads();
        if (lane == 0 && sBeg < nSeg && sGroupD[warp] >= 0)
        {
            int d = sGroupD[warp];
            for (int sgi = sBeg; sgi < sEnd; ++sgi)
            {
                int      cc = d - sSegLo[sgi];
                unsigned g = (cc >= 0 && cc < W) ? (unsigned)sG[sgi * W + cc] : 255u;
                if (g == 255u)
                {
                    atomicMin(&sCtl[4], sgi);
                    break;
                }
                sEntry[sgi] = cc, d += (int)g, sAfter[sgi] = d;
            }
        }
        __syncthreads();
        const int sv = sCtl[4];  // segments [0, sv) are committed
        const int committed = min(sv * kSeg, nLive);
        if (threadIdx.x == 0)
        {
            unsigned nj0 = (unsigned)j0, nm0 = m0;
            bool     done = true;
            if (sv == 0) atomicExch(state + CH_ERROR, 3u), atomicExch(state + CH_DONE, 1u);  // cannot happen (segment 0 is always valid)
            else
            {
                nj0 = (unsigned)(j0 + committed), nm0 = m0 + (unsigned)sAfter[sv - 1], done = j0 + committed >= R.nU;
                // (read by the host and by k_peer_notify once the kernel has finished)
                state[CH_J0] = nj0, state[CH_M0] = nm0, state[CH_ITERS] = it + 1;
                if (done) state[CH_DONE] = 1u;
            }
            st_release_u64(reinterpret_cast<unsigned long long*>(state + CH_PACKED), chain_pack(nj0, nm0, done, it + 1));
            sCtl[0] = (int)nj0, sCtl[1] = (int)nm0, sCtl[2] = done ? 1 : 0;  // this CTA's own next iteration
        }
        for (int t = threadIdx.x; t < committed; t += blockDim.x)
        {
            int sgi = t >> 5;
            flagU[j0 + t] = (uint8_t)((__ldcg(GM + (size_t)sgi * W + sEntry[sgi]) >> (t & 31)) & 1u);
        }
        __syncthreads();
    }
}

// The kernel proper, in builds that differ only in their register budget.  One CTA of 1024 threads per SM: at 64 registers it
// owns the whole register file; the capped builds (48 / 40 / 32 registers) leave room for the blocks of the passes queued
// on the main stream behind it — SSAO, the blur — to become resident next to it and fill the issue slots this
// latency-bound kernel leaves empty (DESIGN.md §4 "Overlap").
#define FGL_CHAIN_KERNEL(NAME, ATTR)                                                                                                          \
    __global__ void ATTR NAME(ChainRows R, unsigned* state, const int* Ppre, uint8_t* G8all, uint32_t* GMall, int* segLoAll, uint32_t* rowBits, \
                              int* rowLo, uint8_t* flagU, int segsPerIter, int pilotK)                                                        \
    {                                                                                                                                         \
        chain_fused_body<kChainNW>(R, state, Ppre, G8all, GMall, segLoAll, rowBits, rowLo, flagU, segsPerIter, pilotK);                       \
    }
FGL_CHAIN_KERNEL(k_chain_fused_r64, __launch_bounds__(1024, 1))
FGL_CHAIN_KERNEL(k_chain_fused_r48, __maxnreg__(48))
FGL_CHAIN_KERNEL(k_chain_fused_r40, __maxnreg__(40))
FGL_CHAIN_KERNEL(k_chain_fused_r32, __launch_bounds__(1024, 2))
#undef FGL_CHAIN_KERNEL

// ---- device-side hand-off of the chain state between the bands of a sort-first group ----------------------------------
// Band r + 1 needs the number of blockers found in bands 0..r.  Instead of a host round trip per band (stream sync, NCCL
// send / recv, launch), the contexts exchange it through peer memory over NVLink: every context owns a mailbox of 16
// (epoch, value) slots; k_peer_notify of band r stores its running total into band r + 1's mailbox (value, system
// fence, epoch), k_peer_wait of band r + 1 — queued on its stream in front of the chain kernel — spins on the slot of
// the current frame.  A 30 s time-out turns a missing neighbour into an error instead of a hang.
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr unsigned long long kPeerFailed = ~0ull;  // travels in place of a blocker count when a band's wait or chain failed
__global__ void k_peer_wait(const unsigned long long* mailbox, unsigned long long epoch, unsigned long long* kDev, unsigned* err)
{
    const unsigned long long* slot = mailbox + (epoch & 15ull) * 2;
    const unsigned long long  t0 = global_timer_ns();
    while (ld_acquire_sys_u64(slot) != epoch)
    {
        __nanosleep(200);
        if (global_timer_ns() - t0 > 30000000000ull)
        {
            *err = 1u, *kDev = 0ull;
            return;
        }
    }
    const unsigned long long v = ld_acquire_sys_u64(slot + 1);
    if (v == kPeerFailed) *err = 2u, *kDev = 0ull;  // a band above gave up: pass the failure on instead of a count
    else *kDev = v;
}
__global__ void k_peer_notify(unsigned long long* nextMailbox, unsigned long long epoch, const unsigned long long* kDev, const unsigned* state,
                              unsigned long long hostBefore, unsigned nC1, int hasChain, unsigned long long* totalOut, const unsigned* err)
{
    unsigned long long total = hostBefore + (kDev ? *kDev : 0ull) + nC1 + (hasChain ? (unsigned long long)state[CH_M0] : 0ull);
    if (*err || (hasChain && (state[CH_ERROR] || !state[CH_DONE]))) total = kPeerFailed;
    *totalOut = total;
    if (nextMailbox)
    {
        unsigned long long* slot = nextMailbox + (epoch & 15ull) * 2;
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(slot + 1), "l"(total) : "memory");
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(slot), "l"(epoch) : "memory");
    }
}

__global__ void __launch_bounds__(256) k_pixel_flags(size_t n, const int* isU, const int* isC1, const int* posU, const uint8_t* flagU, int* hasBlocker)
{
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    hasBlocker[idx] = isU[idx] ? (int)flagU[posU[idx]] : isC1[idx];
}
// chunk index of every pixel, visibility 1 for the pixels without a blocker (shadow.cpp:96-99), list of the others.
//
// Deep-shadow shortcut (exact): a site whose 32 search taps all block (class "certain") has an average blocker depth of
// at least zbLo = (1 - 4e-6) min(search box) (ordered fp32 sum of 32 values >= that minimum), hence a penumbra of at most
// (z - zbLo) A / zbLo and a PCF filter radius of at most F = pcfFilter * penumbra (monotone in the blocker depth; 1e-5
// of slack for the fp32 roundings).  If all texels within ceil(iw F) + 2 of the site's texel fail the depth test —
// z > max(box) + bias, the box maximum assembled from k x k look-ups of the search-radius maximum map, k <= 4 — and
// no tap can leave the map, every one of the 64 PCF taps fails, and the visibility is exactly 0 / 64.
struct DeepShadow
{
    const int*   isC1;
    const float4* sc4;
    const float *smMin, *smMax;
    ShadowMapD   sm;
    int          r;  // radius of the smMin / smMax boxes
    double       pcfFilter;
    float        areaLight;
};
__device__ __forceinline__ bool chunk_index_pixel(size_t idx, size_t n, unsigned long long base, const int* kpre, const int* hasB, unsigned* chunkOf, float* vis,
                                                  unsigned* blockerList, unsigned* nBlockers, const DeepShadow& D, const unsigned long long* kDev)
{
    bool filtered = false;
    if (kDev) base += 2ull * *kDev;
    chunkOf[idx] = (unsigned)(base + idx + 2ull * (unsigned)kpre[idx]);
    float v = 1.f;
    if (hasB[idx])
    {
        unsigned entry = (unsigned)idx;
        if (D.isC1[idx])
        {
            const float4 s = D.sc4[idx];
            // centre texel as k_classify computed it (a certain site has all its search taps, hence its centre, inside the map)
            const int    cx = (int)((float)D.sm.iw * clampf(s.x, 0.f, 1.f)), cy = (int)((float)D.sm.ih * clampf(s.y, 0.f, 1.f));
            const float  dmin = D.smMin[(size_t)cy * D.sm.w + cx];
            if (dmin >= 0.002f && s.z == s.z)
            {
                const double zbLo = (double)dmin * (1.0 - 4e-6);
                const double F = D.pcfFilter * (((double)s.z - zbLo) * (double)D.areaLight / zbLo) * (1.0 + 1e-5);
                const bool   inside = F >= 0.0 && (double)s.x - F >= 1e-6 && (double)s.x + F <= 1.0 - 1e-6 && (double)s.y - F >= 1e-6 && (double)s.y + F <= 1.0 - 1e-6;
                const double rp = ceil((double)max(D.sm.iw, D.sm.ih) * F) + 2.0;
                if (inside && rp <= 4.0 * D.r)
                {
                    const int k = max(1, ((int)rp + D.r - 1) / D.r);
                    float     mx = -__int_as_float(0x7f800000);
                    for (int j = 0; j < k; ++j)
                        for (int i = 0; i < k; ++i)
                        {
                            int x = min(max(cx - (k - 1) * D.r + 2 * D.r * i, 0), D.sm.w - 1), y = min(max(cy - (k - 1) * D.r + 2 * D.r * j, 0), D.sm.h - 1);
                            mx = fmaxf(mx, __ldg(D.smMax + (size_t)y * D.sm.w + x));
                        }
                    if (s.z > mx + s.w) v = 0.f, entry = 0xffffffffu;
                }
            }
        }
        filtered = entry != 0xffffffffu;
    }
    vis[idx] = v;
    if (idx == n - 1) *nBlockers = (unsigned)(kpre[idx] + hasB[idx]);
    return filtered;
}

__global__ void __launch_bounds__(256) k_chunk_index(size_t n, unsigned long long base, const int* kpre, const int* hasB, unsigned* chunkOf, float* vis,
                                                     unsigned* blockerList, unsigned* nBlockers, DeepShadow D, const unsigned long long* kDev, unsigned* nFiltered)
{
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool   filtered = false;
    if (idx < n) filtered = chunk_index_pixel(idx, n, base, kpre, hasB, chunkOf, vis, blockerList, nBlockers, D, kDev);
    // the pixels that still need their 96 taps, as a DENSE list (a warp reserves its slots with one atomic; the order of the
    // list is arbitrary, its entries are independent of one another)
    const unsigned m = __ballot_sync(0xffffffffu, filtered);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    unsigned  slot = 0;
    if (lane == 0) slot = atomicAdd(nFiltered, (unsigned)__popc(m));
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (filtered) blockerList[slot + __popc(m & ((1u << lane) - 1u))] = (unsigned)idx;
}

// shadow.cpp:47-63 for one pixel by one warp: 64 taps, two per lane; the sum of 1/64 steps is exact
__device__ __forceinline__ float pcf_taps(const ShadowMapD& sm, float4 s, float filterSize, float2 d1, float2 d2)
{
    float a = shadow_lookup(sm, s.x + d1.x * filterSize, s.y + d1.y * filterSize);
    float b = shadow_lookup(sm, s.x + d2.x * filterSize, s.y + d2.y * filterSize);
    int   cnt = __popc(__ballot_sync(0xffffffffu, s.z <= a + s.w)) + __popc(__ballot_sync(0xffffffffu, s.z <= b + s.w));
    return (float)cnt * (1.f / 64.f);
}

// PCSS visibility of the pixels that have a blocker (shadow.cpp:92-106): persistent warps over the blocker list.
// The average blocker depth is an ORDERED fp32 sum over the blocking taps (shadow.cpp:78-84): the blocking lanes
// scatter their depth to shared memory at their rank, every lane then adds the n values in order (broadcast 128-bit
// reads) — 2-3 instructions per tap instead of a shuffle loop (most pixels with a blocker have all 32 taps blocked).
__global__ void __launch_bounds__(256) k_pcss_visibility(const unsigned* blockerList, const unsigned* nBlockers, const float4* sc4, const unsigned* chunkOf,
                                                        ShadowMapD sm, const float2* disk, double fs, double pcfFilter, float areaLight, float* vis)
{
    __shared__ __align__(16) float sDepth[8][32];
    const int      lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned nb = *nBlockers, nWarps = (gridDim.x * blockDim.x) >> 5;
    const unsigned ltMask = (1u << lane) - 1u;
    for (unsigned i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < nb; i += nWarps)
    {
        unsigned idx = blockerList[i];  // (pixels deep in shadow are not on the list: visibility 0 already written by k_chunk_index)
        float4   s = sc4[idx];
        size_t   first = (size_t)chunkOf[idx] * 32;
        float2   d0 = __ldg(disk + first + lane), d1 = __ldg(disk + first + 32 + lane), d2 = __ldg(disk + first + 64 + lane);
        float    ox = (float)((double)d0.x * fs), oy = (float)((double)d0.y * fs);
        float    sampleDepth = shadow_lookup(sm, s.x + ox, s.y + oy);
        bool     blocked = s.z > sampleDepth + s.w;
        unsigned mask = __ballot_sync(0xffffffffu, blocked);
        const int n = __popc(mask);
        if (blocked) sDepth[wid][__popc(mask & ltMask)] = sampleDepth;
        __syncwarp();
        float sum = 0.f;
        const float4* q = reinterpret_cast<const float4*>(sDepth[wid]);
#pragma unroll
        for (int k = 0; k < 8; ++k)
        {
            if (4 * k >= n) break;
            float4 v = q[k];
            sum += v.x;
            if (4 * k + 1 < n) sum += v.y;
            if (4 * k + 2 < n) sum += v.z;
            if (4 * k + 3 < n) sum += v.w;
        }
        __syncwarp();
        float dBlocker = mask ? sum / (float)n : 0.f;
        float v = 1.f;
        if (!(dBlocker < 0.001f))  // (double)dBlocker < 0.001 (shadow.cpp:101): float(0.001) is the smallest float above 0.001
        {
            float penumbra = (s.z - dBlocker) * areaLight / dBlocker;
            v = pcf_taps(sm, s, (float)(pcfFilter * (double)penumbra), d1, d2);
        }
        if (lane == 0) vis[idx] = v;
    }
}

// The same filter with the loads that do not depend on any arithmetic fetched ahead.  List entry -> coordinate + chunk index ->
// samples -> shadow-map gather -> (ordered sum) -> PCF gathers are five dependent round trips per pixel and a warp issues in
// order; the first three are interleaved with the two gather phases of the current pixel:
//   (A) list entry two pixels ahead, coordinate + chunk index one pixel ahead     — issued, not used
//   (B) current pixel: blocker search (one gather round trip), ordered sum, penumbra
//   (C) the next pixel's 96 samples (their address, the chunk index, arrived during B) — issued, not used
//   (D) current pixel: the 64 PCF taps (one gather round trip)
// 40 registers: six CTAs per SM instead of eight.  Selected by FGL_VIS_PIPELINE=1; NOT the default: on the dense list the plain
// kernel above is the faster one (C3 0.186 vs 0.202 ms, C5 0.155 vs 0.168 ms): with the deep-shadow entries gone every warp has
// real work queued behind it, the long-scoreboard stalls the fetch-ahead removes (25.7 -> 7.7 per issue, ncu) were already
// covered by the other 63 warps of the SM, and its extra instructions and lower occupancy cost more than they save.
__global__ void __launch_bounds__(256, 6) k_pcss_visibility_pipe(const unsigned* blockerList, const unsigned* nBlockers, const float4* sc4, const unsigned* chunkOf,
                                                                ShadowMapD sm, const float2* disk, double fs, double pcfFilter, float areaLight, float* vis)
{
    __shared__ __align__(16) float sDepth[8][32];
    const int      lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned nb = *nBlockers, nWarps = (gridDim.x * blockDim.x) >> 5;
    const unsigned ltMask = (1u << lane) - 1u;
    unsigned       i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= nb) return;
    // pipeline registers: pixel i (coordinate and samples), pixel i + nWarps (list entry)
    unsigned idx = blockerList[i];
    unsigned idxNext = i + nWarps < nb ? blockerList[i + nWarps] : 0u;
    float4   s = sc4[idx];
    float2   d0, d1, d2;
    {
        const float2* p = disk + (size_t)chunkOf[idx] * 32 + lane;
        d0 = p[0], d1 = p[32], d2 = p[64];
    }
    for (; i < nb; i += nWarps)
    {
        const bool more = i + nWarps < nb;
        // (A)
        const unsigned idxAfter = i + 2 * nWarps < nb ? blockerList[i + 2 * nWarps] : 0u;
        float4         sNext = make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned       chunkNext = 0u;
        if (more) sNext = sc4[idxNext], chunkNext = chunkOf[idxNext];
        asm volatile("" ::: "memory");
        // (B)
        float dBlocker;
        {
            float    ox = (float)((double)d0.x * fs), oy = (float)((double)d0.y * fs);
            float    sampleDepth = shadow_lookup(sm, s.x + ox, s.y + oy);
            bool     blocked = s.z > sampleDepth + s.w;
            unsigned mask = __ballot_sync(0xffffffffu, blocked);
            const int n = __popc(mask);
            if (blocked) sDepth[wid][__popc(mask & ltMask)] = sampleDepth;
            __syncwarp();
            float sum = 0.f;
            const float4* q = reinterpret_cast<const float4*>(sDepth[wid]);
#pragma unroll
            for (int k = 0; k < 8; ++k)
            {
                if (4 * k >= n) break;
                float4 v = q[k];
                sum += v.x;
                if (4 * k + 1 < n) sum += v.y;
                if (4 * k + 2 < n) sum += v.z;
                if (4 * k + 3 < n) sum += v.w;
            }
            __syncwarp();
            dBlocker = mask ? sum / (float)n : 0.f;
        }
        asm volatile("" ::: "memory");
        // (C)
        float2 e0 = make_float2(0.f, 0.f), e1 = e0, e2 = e0;
        if (more)
        {
            const float2* p = disk + (size_t)chunkNext * 32 + lane;
            e0 = p[0], e1 = p[32], e2 = p[64];
        }
        asm volatile("" ::: "memory");
        // (D)
        float v = 1.f;
        if (!(dBlocker < 0.001f))  // (double)dBlocker < 0.001 (shadow.cpp:101): float(0.001) is the smallest float above 0.001
        {
            float penumbra = (s.z - dBlocker) * areaLight / dBlocker;
            v = pcf_taps(sm, s, (float)(pcfFilter * (double)penumbra), d1, d2);
        }
        if (lane == 0) vis[idx] = v;
        idx = idxNext, idxNext = idxAfter, s = sNext, d0 = e0, d1 = e1, d2 = e2;
    }
}

// PCF visibility (shadow.cpp:47-63, 120-124): site i consumes accepted disk samples 64 i .. 64 i + 63; one warp per site
__global__ void __launch_bounds__(256) k_pcf_visibility(ShadowMapD sm, size_t first, size_t last, const float4* sc4, const float2* disk, float filterSize,
                                                       float* vis)
{
    int    lane = threadIdx.x & 31;
    size_t idx = first + (((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (idx >= last) return;
    float2 d1 = __ldg(disk + idx * 64 + lane), d2 = __ldg(disk + idx * 64 + 32 + lane);
    float  v = pcf_taps(sm, sc4[idx], filterSize, d1, d2);
    if (lane == 0) vis[idx] = v;
}
}  // namespace

// ====================================================================================================================
struct SampleStream
{
    DevBuf    ckpt;
    long long nCkpt = 0;  // checkpoints 0 .. nCkpt-1 exist
    DevBuf    window, tileCounts, tileOffsets, counters;
    // tables
    DevBuf             ball, disk;
    unsigned long long ballNeed = 0, ballRawEnd = 0;  // ball table holds ballNeed samples; the stream then stands at ballRawEnd
    unsigned long long diskNeed = 0, diskRawBegin = ~0ull;
    // frame state
    bool               ssaoThisFrame = false;
    unsigned long long ssaoSamples = 0;
    // chain scratch
    DevBuf smTmpMin, smTmpMax, smMin, smMax, boxMin, boxMax, sc4, isU, isC1, posU, c1pre, Upix, Uc1, Usc, UF, UE, bits, flagU, hasB, kpre, chunkOf, mState;
    DevBuf sig, vis, blockerList, pilot, Ppre, winLo, chainStats, rowBits, rect;
    bool   rectValid = false;  // rect holds the texel rectangle of this frame's sites (deferred lighting); else the box maps cover the whole map
    unsigned long long sigChunks = 0;
    // device-side hand-off (peer mailboxes)
    DevBuf              mailbox, peerLocal;  // 16 x (epoch, value); {kDev, total, err}
    unsigned long long* nextMailbox = nullptr;
    void*               nextMailboxIpc = nullptr;  // mapping to close
    bool                peerOn = false, peerWait = false, peerTotalOnDevice = false;
    unsigned long long  peerEpoch = 0;
    // FGL_VIS_PREPARE -> FGL_VIS_RESOLVE hand-over
    bool               prepValid = false, chainInFlight = false, chainWasShared = false;
    unsigned long long inflightBefore = 0;
    size_t             prepTotal = 0, prepLo = 0, prepHi = 0;
    int                prepNU = 0, prepNC1 = 0;
    unsigned long long chainTotal = 0;  // blockers found up to and including this context's band
    bool               chainCountValid = false;
};

static SampleStream* S_of(fgl_ctx* c)
{
    if (!c->stream_state) c->stream_state = new SampleStream();
    return c->stream_state;
}

void fgl_stream_destroy(fgl_ctx* c)
{
    SampleStream* s = c->stream_state;
    if (!s) return;
    if (s->nextMailboxIpc) cudaIpcCloseMemHandle(s->nextMailboxIpc);
    if (s->mailbox.p) cudaFree(s->mailbox.p);
    if (s->peerLocal.p) cudaFree(s->peerLocal.p);
    DevBuf* all[] = { &s->ckpt, &s->window, &s->tileCounts, &s->tileOffsets, &s->counters, &s->ball, &s->disk, &s->smTmpMin, &s->smTmpMax, &s->smMin,
                      &s->smMax, &s->boxMin, &s->boxMax, &s->UF, &s->UE, &s->sig, &s->vis, &s->blockerList, &s->pilot, &s->Ppre, &s->winLo, &s->chainStats, &s->rowBits, &s->rect, &s->sc4, &s->isU, &s->isC1, &s->posU, &s->c1pre, &s->Upix, &s->Uc1, &s->Usc, &s->bits, &s->flagU, &s->hasB, &s->kpre,
                      &s->chunkOf, &s->mState };
    for (DevBuf* b : all)
        if (b->p) cudaFree(b->p);
    delete s;
    c->stream_state = nullptr;
}

void fgl_stream_begin_frame(fgl_ctx* c)
{
    SampleStream* s = S_of(c);
    s->prepValid = false, s->chainInFlight = false, s->rectValid = false;
    if (s->peerOn) ++s->peerEpoch;
    s->ssaoThisFrame = false;
    s->ssaoSamples = 0;
}

static int ensure_checkpoints(fgl_ctx* c, SampleStream* s, long long wantCk)
{
    if (wantCk <= s->nCkpt) return FGL_OK;
    wantCk += wantCk / 8 + 64;  // amortise
    if (int rc = fgl_reserve(c, s->ckpt, (size_t)wantCk * kMT * 4)) return rc;
    if (s->nCkpt == 0)
    {   // std::mt19937 default seed (utility.h:92: a default-constructed engine)
        std::vector<uint32_t> st(kMT);
        st[0] = 5489u;
        for (int i = 1; i < kMT; ++i) st[i] = 1812433253u * (st[i - 1] ^ (st[i - 1] >> 30)) + (uint32_t)i;
        FGL_CUDA(c, cudaMemcpyAsync(s->ckpt.p, st.data(), kMT * 4, cudaMemcpyHostToDevice, c->stream));
        FGL_CUDA(c, cudaStreamSynchronize(c->stream));
        s->nCkpt = 1;
    }
    {
        LaunchScope ls(c, "mt_checkpoints", 0);
        k_mt_checkpoints<<<1, 256, 0, c->stream>>>((uint32_t*)s->ckpt.p, s->nCkpt, wantCk);
    }
    s->nCkpt = wantCk;
    return FGL_OK;
}

// Builds a table of `need` accepted samples (K = 3: unit ball, K = 2: unit disk) starting at draw rawBegin.
template <int K>
static int build_table(fgl_ctx* c, SampleStream* s, DevBuf& table, unsigned long long rawBegin, unsigned long long need, unsigned long long* rawEndOut)
{
    if (int rc = fgl_not_while_recording(c, "building a sample table (render the frame once without recording first)")) return rc;
    if (int rc = fgl_reserve(c, table, (size_t)need * (K == 3 ? 4 : K) * 4 + 16)) return rc;  // ball: x, y, z + the sample's SSAO scale
    if (int rc = fgl_reserve(c, s->counters, 64)) return rc;
    unsigned long long* dAcc = (unsigned long long*)s->counters.p;
    unsigned long long* dRawEnd = dAcc + 1;
    FGL_CUDA(c, cudaMemsetAsync(s->counters.p, 0, 64, c->stream));
    const double       rate = K == 3 ? 0.5235987 : 0.7853981;
    unsigned long long acc = 0, candDone = 0;
    while (acc < need)
    {
        unsigned long long remaining = need - acc;
        size_t             nCand = (size_t)std::min<double>((double)kWindowCand, remaining / rate * 1.002 + 65536.0);
        unsigned long long r0 = rawBegin + (unsigned long long)K * candDone, r1 = r0 + (unsigned long long)K * nCand;
        long long          firstCk = (long long)(r0 / kMT / kCB);
        long long          lastBlock = (long long)((r1 + kMT - 1) / kMT);  // exclusive
        long long          nCk = (lastBlock + kCB - 1) / kCB - firstCk;
        long long          nBlocks = lastBlock - firstCk * kCB;
        if (int rc = ensure_checkpoints(c, s, firstCk + nCk)) return rc;
        if (int rc = fgl_reserve(c, s->window, (size_t)nCk * kCB * kMT * 4)) return rc;
        {
            LaunchScope ls(c, "mt_generate", (uint64_t)nBlocks * kMT * 4);
            k_mt_generate<<<(unsigned)nCk, 256, 0, c->stream>>>((const uint32_t*)s->ckpt.p, firstCk, nBlocks, (uint32_t*)s->window.p);
        }
        const uint32_t* raw = (const uint32_t*)s->window.p + (r0 - (unsigned long long)firstCk * kCB * kMT);
        size_t          nTiles = (nCand + 255) / 256;
        if (int rc = fgl_reserve(c, s->tileCounts, (nTiles + 1) * 4)) return rc;
        if (int rc = fgl_reserve(c, s->tileOffsets, (nTiles + 1) * 4)) return rc;
        FGL_CUDA(c, cudaMemsetAsync((int*)s->tileCounts.p + nTiles, 0, 4, c->stream));
        {
            LaunchScope ls(c, "stream_count", (uint64_t)nCand * K * 4);
            k_count<K><<<(unsigned)nTiles, 256, 0, c->stream>>>(raw, nCand, (int*)s->tileCounts.p);
        }
        size_t tmpBytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, (int*)s->tileCounts.p, (int*)s->tileOffsets.p, (int)nTiles + 1, c->stream);
        if (int rc = fgl_reserve(c, c->scanTmp, tmpBytes)) return rc;
        {
            LaunchScope ls(c, "scan", nTiles * 8);
            cub::DeviceScan::ExclusiveSum(c->scanTmp.p, tmpBytes, (int*)s->tileCounts.p, (int*)s->tileOffsets.p, (int)nTiles + 1, c->stream);
        }
        {
            LaunchScope ls(c, "stream_compact", (uint64_t)nCand * K * 4);
            k_compact<K><<<(unsigned)nTiles, 256, 0, c->stream>>>(raw, nCand, (const int*)s->tileOffsets.p, dAcc, need, (float*)table.p, r0, dRawEnd);
        }
        ++c->launches;
        k_acc_add<<<1, 1, 0, c->stream>>>(dAcc, (const int*)s->tileOffsets.p + nTiles);
        unsigned long long host[2];
        FGL_CUDA(c, cudaMemcpyAsync(host, s->counters.p, 16, cudaMemcpyDeviceToHost, c->stream));
        FGL_CUDA(c, cudaStreamSynchronize(c->stream));
        acc = host[0];
        if (acc >= need && rawEndOut) *rawEndOut = host[1];
        candDone += nCand;
    }
    return FGL_OK;
}

int fgl_stream_prepare_ssao(fgl_ctx* c, SsaoPass& S)
{
    SampleStream*      s = S_of(c);
    unsigned long long need = (unsigned long long)S.W * S.H * 32;
    if (s->ballNeed != need)
    {
        s->ballNeed = 0;
        if (int rc = build_table<3>(c, s, s->ball, 0, need, &s->ballRawEnd)) return rc;
        s->ballNeed = need;
    }
    S.ball = (const float4*)s->ball.p;
    s->ssaoThisFrame = true, s->ssaoSamples = need;
    return FGL_OK;
}

// |(float)(d * fs)| <= fsF for every |d| < 1: the largest tap offset of the blocker search as a float
static float pcss_filter_bound(double fs)
{
    float fsF = (float)fs;
    if (!((double)fsF >= fs)) fsF = nextafterf(fsF, 1e30f);
    return fsF;
}

static int scan_ints(fgl_ctx* c, const int* in, int* out, size_t n)
{
    size_t tmpBytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, in, out, (int)n, c->stream);
    if (int rc = fgl_reserve(c, c->scanTmp, tmpBytes)) return rc;
    LaunchScope ls(c, "scan", n * 8);
    cub::DeviceScan::ExclusiveSum(c->scanTmp.p, tmpBytes, in, out, (int)n, c->stream);
    return FGL_OK;
}

// Several contexts may share a device (frames in flight: one context per frame, each driven by its own host thread).  Two
// persistent chain kernels must never be placed on the SMs at the same time — each needs a CTA on every SM, and two half-placed
// grids would wait for each other at their first barrier — so the chain launches of a device are ordered behind one another:
// every launch waits for the event recorded behind the previous one (whichever context issued it).
static std::mutex  g_chainMutex;
static cudaEvent_t g_chainDone[64] = { nullptr };

// The persistent chain kernel: one CTA of 1024 threads per SM, grid-wide barriers on global counters.
//   exclusive (shared = false): the 64-register build by a cooperative launch — co-residency guaranteed, the device is not
//     shared with other kernels meanwhile;
//   shared = true (the chain runs on its own stream while SSAO and the blur run on the main one): the 40-register build by a
//     plain launch, so that the other passes' blocks become resident next to it.  Co-residency then follows from one CTA per
//     SM being placed first (highest stream priority, issued ahead of SSAO); should a CTA ever be kept waiting beyond the
//     barriers' time-outs the kernel reports an error and the caller re-runs the chain exclusively (see the RESOLVE phase).
// samples per row of the pilot (FGL_CHAIN_PILOT_K; the pilot kernel and the chain kernel must agree)
static int chain_pilot_k()
{
    static const int k = getenv("FGL_CHAIN_PILOT_K") ? std::max(1, std::min(16, atoi(getenv("FGL_CHAIN_PILOT_K")))) : 4;
    return k;
}

static int launch_chain(fgl_ctx* c, SampleStream* s, ChainRows& R, int nU, size_t n, int nC1, cudaStream_t st, bool shared)
{
    typedef void (*ChainKernel)(ChainRows, unsigned*, const int*, uint8_t*, uint32_t*, int*, uint32_t*, int*, uint8_t*, int, int);
    static const int  segsEnv = getenv("FGL_CHAIN_SEGS") ? atoi(getenv("FGL_CHAIN_SEGS")) : 0;
    static const int  regsEnv = getenv("FGL_CHAIN_REGS") ? atoi(getenv("FGL_CHAIN_REGS")) : 0;
    static const int  coopEnv = getenv("FGL_CHAIN_COOP") ? atoi(getenv("FGL_CHAIN_COOP")) : -1;
    static bool       ready = false;
    const size_t      smemBytes = sizeof(ChainSmem<kChainNW>);
    const ChainKernel all[4] = { k_chain_fused_r64, k_chain_fused_r48, k_chain_fused_r40, k_chain_fused_r32 };
    if (!ready)
    {
        for (ChainKernel k : all)
        {
            int perSM = 0;
            FGL_CUDA(c, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
            FGL_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k, 1024, smemBytes));
            if (perSM < 1) return fgl_fail(c, FGL_ERR_CUDA, "pcss chain: the persistent kernel does not fit an SM");
        }
        ready = true;
    }
    if (nU >= (1 << kChainMaxRowsLog2)) return fgl_fail(c, FGL_ERR_UNSUPPORTED, "pcss chain: more than 2^27 uncertain pixels in one band");
    const int         regs = regsEnv ? regsEnv : (shared ? 40 : 64);
    const bool        coop = coopEnv >= 0 ? coopEnv != 0 : !shared;
    const ChainKernel kernel = regs <= 32 ? all[3] : regs <= 40 ? all[2] : regs <= 48 ? all[1] : all[0];
    FGL_CUDA(c, cudaMemsetAsync(s->mState.p, 0, 64, st));
    int grid = std::min(c->numSMs, kMaxGrid);  // one CTA per SM (kMaxSegs bounds the tables: at most kMaxGrid SMs)
    int segsPerIter = segsEnv > 0 ? segsEnv : kMaxSpc * grid;
    segsPerIter = std::max(1, std::min(segsPerIter, kMaxSpc * grid));
    const size_t tw = (size_t)segsPerIter * 32 * kChainNW;
    if (int rc = fgl_reserve(c, s->bits, 2 * tw * 5)) return rc;  // G8 (1 byte) + GM (4 bytes) per table entry, two parities
    if (int rc = fgl_reserve(c, s->winLo, 2 * (size_t)segsPerIter * 4)) return rc;
    uint8_t*     G8 = (uint8_t*)s->bits.p + 2 * tw * 4;
    uint32_t*    GM = (uint32_t*)s->bits.p;
    int*         segLo = (int*)s->winLo.p;
    unsigned*    state = (unsigned*)s->mState.p;
    const int*   Ppre = (const int*)s->Ppre.p;
    uint8_t*     flagU = (uint8_t*)s->flagU.p;
    const size_t nRows = (size_t)segsPerIter * 32;
    if (int rc = fgl_reserve(c, s->rowBits, nRows * kChainNW * 4 + nRows * 4)) return rc;
    uint32_t*    rowBits = (uint32_t*)s->rowBits.p;
    int*         rowLo = (int*)(rowBits + nRows * kChainNW);
    int          pilotK = chain_pilot_k();
    void*        args[] = { &R, &state, &Ppre, &G8, &GM, &segLo, &rowBits, &rowLo, &flagU, &segsPerIter, &pilotK };
    // algorithmic bytes: every uncertain row's record once (45 B) + every chunk signature of the band once (8 B)
    std::lock_guard<std::mutex> chainOrder(g_chainMutex);
    cudaEvent_t&                done = g_chainDone[c->device & 63];
    if (done) FGL_CUDA(c, cudaStreamWaitEvent(st, done, 0));
    else FGL_CUDA(c, cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    {
        LaunchScope ls(c, "pcss_chain", (uint64_t)nU * 45 + ((uint64_t)n + 2ull * ((uint64_t)nC1 + (uint64_t)nU)) * 8);
        if (coop) FGL_CUDA(c, cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3(1024), args, smemBytes, st));
        else
        {
            kernel<<<grid, 1024, smemBytes, st>>>(R, state, Ppre, G8, GM, segLo, rowBits, rowLo, flagU, segsPerIter, pilotK);
            FGL_CUDA(c, cudaGetLastError());
        }
    }
    FGL_CUDA(c, cudaEventRecord(done, st));
    s->chainWasShared = shared;
    return FGL_OK;
}

// Visibility of n stream consumers ("sites") in consumption order, given their shadow coordinate + bias.
// [siteLo, siteHi) = the sites whose visibility is needed (PCF); PCSS resolves the whole chain.
int fgl_stream_site_visibility(fgl_ctx* c, LightPass& L, size_t nTotal, const float4* sc4All, size_t siteLo, size_t siteHi, unsigned long long blockersBefore,
                               int phase)
{
    size_t n = nTotal;  // table sizes follow the whole site set; the kernels below work on [siteLo, siteHi)

    SampleStream*      s = S_of(c);
    cudaStream_t       st = c->stream;
    unsigned long long rawBegin = s->ssaoThisFrame ? s->ballRawEnd : 0ull;
    // PCF: 64 samples per pixel; PCSS: chunk index <= 3 n, plus the 96 samples of the last pixel
    bool               pcss = L.shadowMode == FGL_SHADOW_PCSS;
    unsigned long long nChunks = pcss ? (unsigned long long)n * 3 + 4 : (unsigned long long)n * 2;
    unsigned long long need = nChunks * 32;
    if (s->diskRawBegin != rawBegin || s->diskNeed < need)
    {
        s->diskNeed = 0, s->sigChunks = 0;
        if (int rc = build_table<2>(c, s, s->disk, rawBegin, need, nullptr)) return rc;
        s->diskNeed = need, s->diskRawBegin = rawBegin;
    }
    if (pcss && s->sigChunks < nChunks)
    {
        if (int rc = fgl_reserve(c, s->sig, (size_t)nChunks * 8)) return rc;
        LaunchScope ls(c, "stream_signatures", nChunks * 264);
        k_signatures<<<(unsigned)((nChunks * 32 + 255) / 256), 256, 0, st>>>((const float2*)s->disk.p, (size_t)nChunks, (unsigned long long*)s->sig.p);
        s->sigChunks = nChunks;
    }
    L.disk = (const float2*)s->disk.p;
    L.chunkOf = nullptr;
    if (int rc = fgl_reserve(c, s->vis, n * 4)) return rc;
    L.vis = (const float*)s->vis.p;

    ChainPass P;
    memset(&P, 0, sizeof P);
    P.sm = L.sm;
    if (pcss)
        if (int rc = fgl_not_while_recording(c, "a PCSS frame (its sample-stream chain reports counts to the host)")) return rc;
    if (!pcss)
    {
        if (phase == FGL_VIS_PREPARE || phase == FGL_VIS_LAUNCH) return FGL_OK;  // PCF offsets are closed-form: nothing depends on earlier bands
        size_t      nSites = siteHi - siteLo;
        LaunchScope ls(c, "pcf_visibility", nSites * (16 + 512 + 4));
        if (nSites) k_pcf_visibility<<<(unsigned)((nSites * 32 + 255) / 256), 256, 0, st>>>(L.sm, siteLo, siteHi, sc4All, L.disk, (float)L.pcfFilter, (float*)s->vis.p);
        return FGL_OK;
    }
    // from here on: the sites of this context's band, locally indexed; site i is global site siteLo + i
    n = siteHi - siteLo;
    const float4*            sc4In = sc4All + siteLo;
    float*                   visB = (float*)s->vis.p + siteLo;
    const unsigned long long chunkBase = (unsigned long long)siteLo + 2ull * blockersBefore;

    if (n == 0)
    {   // an empty band (more GPUs than rows): no pixel consumes samples, the running blocker count passes through unchanged
        s->prepValid = false, s->chainInFlight = false;
        if (phase == FGL_VIS_PREPARE) return FGL_OK;
        if (int rc = fgl_reserve(c, s->mState, 64)) return rc;
        if (s->peerOn)
        {
            unsigned long long* loc = (unsigned long long*)s->peerLocal.p;
            FGL_CUDA(c, cudaMemsetAsync(loc, 0, 24, st));
            if (s->peerWait)
            {
                LaunchScope ls(c, "pcss_peer_wait", 0);
                k_peer_wait<<<1, 1, 0, st>>>((const unsigned long long*)s->mailbox.p, s->peerEpoch, loc, (unsigned*)(loc + 2));
            }
            LaunchScope ls(c, "pcss_peer_notify", 0);
            k_peer_notify<<<1, 1, 0, st>>>(s->nextMailbox, s->peerEpoch, loc, (const unsigned*)s->mState.p, blockersBefore, 0u, 0, loc + 1, (const unsigned*)(loc + 2));
        }
        s->chainTotal = blockersBefore, s->chainCountValid = true, s->peerTotalOnDevice = s->peerOn;
        return FGL_OK;
    }
    // ---- PCSS chain --------------------------------------------------------------------------------------------
    size_t  smN = (size_t)L.sm.w * L.sm.h;
    DevBuf* f4[] = { &s->smTmpMin, &s->smTmpMax, &s->smMin, &s->smMax, &s->boxMin, &s->boxMax };
    for (DevBuf* b : f4)
        if (int rc = fgl_reserve(c, *b, smN * 4)) return rc;
    DevBuf* i4[] = { &s->isU, &s->isC1, &s->posU, &s->c1pre, &s->hasB, &s->kpre, &s->blockerList };
    for (DevBuf* b : i4)
        if (int rc = fgl_reserve(c, *b, (n + 1) * 4)) return rc;
    if (int rc = fgl_reserve(c, s->chunkOf, nTotal * 4)) return rc;
    if (int rc = fgl_reserve(c, s->mState, 64)) return rc;
    unsigned* chunkOfB = (unsigned*)s->chunkOf.p + siteLo;

    const float fsF = pcss_filter_bound(L.pcssFilter);
    int r = (int)ceil((double)L.sm.iw * (double)fsF) + 1;
    int bw = (int)ceil(0.25 * (double)L.sm.iw * (double)fsF) + 2;
    unsigned nb = (unsigned)((n + 255) / 256);
    int      nU = 0, nC1 = 0;  // uncertain sites; sites whose every tap blocks
    // Everything up to the pilot's prefix sums is independent of the blockers found in earlier bands: a sort-first
    // driver runs it (FGL_VIS_PREPARE) while it waits for the previous band's count, then resolves (FGL_VIS_RESOLVE).
    const bool prepared = (phase == FGL_VIS_RESOLVE || phase == FGL_VIS_LAUNCH) && s->prepValid && s->prepTotal == nTotal && s->prepLo == siteLo && s->prepHi == siteHi;
    // the chain may already be running (or finished) on its own stream: FGL_VIS_LAUNCH issued it, FGL_VIS_RESOLVE picks it up
    const bool inflight = phase == FGL_VIS_RESOLVE && prepared && s->chainInFlight && s->inflightBefore == blockersBefore;
    s->chainInFlight = false;
    if (prepared) nU = s->prepNU, nC1 = s->prepNC1;
    s->prepValid = false;
    ChainRows R;
    memset(&R, 0, sizeof R);
    R.sm = L.sm, R.disk = L.disk, R.fs = L.pcssFilter, R.sig = (const unsigned long long*)s->sig.p;
    R.base = phase == FGL_VIS_PREPARE ? (unsigned long long)siteLo : chunkBase;  // the pilot only predicts: any nearby offset will do
    if (!prepared)
    {
    {
        // byte model: 48 B per texel the passes really produce (two box maps, two passes each) — the whole map, or the texel
        // rectangle of the band's sites grown by the margin, counted on the device
        const unsigned* areaPtr = nullptr;
        if (s->rectValid)
        {
            unsigned* area = (unsigned*)s->rect.p + 4;
            k_rect_area<<<1, 1, 0, st>>>((const int*)s->rect.p, 3 * r + 2, L.sm.w, L.sm.h, area);
            areaPtr = area;
        }
        LaunchScope ls(c, "pcss_minmax", areaPtr ? 0 : smN * 48, areaPtr, 48);
        auto box = [&](int lo, int hi, float* omin, float* omax, int margin) -> int {
            const int w = hi - lo + 1;
            MapRegion region;
            region.rect = s->rectValid ? (const int*)s->rect.p : nullptr, region.margin = margin;
            float *   tmin = (float*)s->smTmpMin.p, *tmax = (float*)s->smTmpMax.p;
            if (w <= 129)
            {
                static bool attr = false;
                if (!attr)
                {
                    FGL_CUDA(c, cudaFuncSetAttribute(k_minmax_v, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 256 * 32 * 8));
                    attr = true;
                }
                k_minmax_h<true><<<dim3((L.sm.w + kMMW - 1) / kMMW, L.sm.h), 256, 0, st>>>(L.sm.d, nullptr, L.sm.w, L.sm.h, lo, hi, tmin, tmax, region);
                k_minmax_v<<<dim3((L.sm.w + 31) / 32, (L.sm.h + kMMR - 1) / kMMR), 512, (size_t)2 * (kMMR + w - 1) * 32 * 8, st>>>(tmin, tmax, L.sm.w, L.sm.h, lo, hi, omin,
                                                                                                                                  omax, region);
            }
            else
            {
                dim3 grid((L.sm.w + 127) / 128, L.sm.h);
                k_minmax_h_wide<<<grid, 128, 0, st>>>(L.sm.d, L.sm.w, L.sm.h, lo, hi, tmin, tmax);
                k_minmax_v_wide<<<grid, 128, 0, st>>>(tmin, tmax, L.sm.w, L.sm.h, lo, hi, omin, omax);
            }
            return FGL_OK;
        };
        // who reads where: k_classify / k_chunk_index at the centre texel and (deep-shadow shortcut) up to 3 r around it;
        // k_pixel_masks at a cell's first texel, at most r from the centre
        if (int rc = box(-r, r, (float*)s->smMin.p, (float*)s->smMax.p, 3 * r + 2)) return rc;
        if (int rc = box(0, bw - 1, (float*)s->boxMin.p, (float*)s->boxMax.p, r + 2)) return rc;
        c->launches += 3;
    }
    P.smMin = (const float*)s->smMin.p, P.smMax = (const float*)s->smMax.p, P.r = r, P.fsF = fsF;
    {
        LaunchScope ls(c, "pcss_classify", n * (16 + 8));
        k_classify<<<nb, 256, 0, st>>>(P, n, sc4In, (int*)s->isU.p, (int*)s->isC1.p);
    }
    FGL_CUDA(c, cudaMemsetAsync((int*)s->isU.p + n, 0, 4, st));
    FGL_CUDA(c, cudaMemsetAsync((int*)s->isC1.p + n, 0, 4, st));
    if (int rc = scan_ints(c, (const int*)s->isU.p, (int*)s->posU.p, n + 1)) return rc;
    if (int rc = scan_ints(c, (const int*)s->isC1.p, (int*)s->c1pre.p, n + 1)) return rc;
    FGL_CUDA(c, cudaMemcpyAsync(&nU, (int*)s->posU.p + n, 4, cudaMemcpyDeviceToHost, st));
    FGL_CUDA(c, cudaMemcpyAsync(&nC1, (int*)s->c1pre.p + n, 4, cudaMemcpyDeviceToHost, st));
    c->d2hBytes += 8;
    FGL_CUDA(c, cudaStreamSynchronize(st));
    if (int rc = fgl_reserve(c, s->Upix, (size_t)(nU + 1) * 4)) return rc;
    if (int rc = fgl_reserve(c, s->Uc1, (size_t)(nU + 1) * 4)) return rc;
    if (int rc = fgl_reserve(c, s->Usc, (size_t)(nU + 1) * 16)) return rc;
    if (int rc = fgl_reserve(c, s->UF, (size_t)(nU + 1) * 8)) return rc;
    if (int rc = fgl_reserve(c, s->UE, (size_t)(nU + 1) * 8)) return rc;
    if (int rc = fgl_reserve(c, s->flagU, (size_t)nU + 16)) return rc;
    c->lastUncertain = nU;
    if (nU > 0)
    {
        {
            LaunchScope ls(c, "pcss_gather", n * 12);
            k_gather_uncertain<<<nb, 256, 0, st>>>(n, (const int*)s->isU.p, (const int*)s->posU.p, (const int*)s->c1pre.p, sc4In,
                                                   (unsigned*)s->Upix.p, (unsigned*)s->Uc1.p, (float4*)s->Usc.p);
        }
        {
            MaskPass M;
            M.sm = L.sm, M.boxMin = (const float*)s->boxMin.p, M.boxMax = (const float*)s->boxMax.p, M.bw = bw, M.fs = L.pcssFilter;
            LaunchScope ls(c, "pcss_masks", (uint64_t)nU * (16 + 16 + 512));
            k_pixel_masks<<<(unsigned)(((size_t)nU * 32 + 255) / 256), 256, 0, st>>>(M, nU, (const float4*)s->Usc.p, (unsigned long long*)s->UF.p,
                                                                                    (unsigned long long*)s->UE.p);
        }
        R.nU = nU, R.Upix = (const unsigned*)s->Upix.p, R.Uc1 = (const unsigned*)s->Uc1.p, R.Usc = (const float4*)s->Usc.p;
        R.UF = (const unsigned long long*)s->UF.p, R.UE = (const unsigned long long*)s->UE.p;
        static const bool wantStats = getenv("FGL_CHAIN_STATS") != nullptr;
        R.stats = nullptr;
        if (wantStats)
        {
            if (int rc = fgl_reserve(c, s->chainStats, 128)) return rc;
            FGL_CUDA(c, cudaMemsetAsync(s->chainStats.p, 0, 128, st));
            R.stats = (unsigned long long*)s->chainStats.p;
        }
        if (int rc = fgl_reserve(c, s->pilot, (size_t)(nU + 1) * 4)) return rc;
        if (int rc = fgl_reserve(c, s->Ppre, (size_t)(nU + 1) * 4)) return rc;
        {
            LaunchScope ls(c, "pcss_chain_pilot", (uint64_t)nU * 48);
            k_chain_pilot<<<(nU + 255) / 256, 256, 0, st>>>(R, (int*)s->pilot.p, chain_pilot_k());
        }
        FGL_CUDA(c, cudaMemsetAsync((int*)s->pilot.p + nU, 0, 4, st));
        if (int rc = scan_ints(c, (const int*)s->pilot.p, (int*)s->Ppre.p, (size_t)nU + 1)) return rc;
    }
    }  // !prepared
    if (phase == FGL_VIS_PREPARE)
    {
        s->prepValid = true, s->prepTotal = nTotal, s->prepLo = siteLo, s->prepHi = siteHi, s->prepNU = nU, s->prepNC1 = nC1;
        return FGL_OK;
    }
    unsigned long long uncertainBlockers = 0;
    c->lastUncertain = nU;
    R.kDev = s->peerOn ? (const unsigned long long*)s->peerLocal.p : nullptr;
    if (!inflight)
    {
    if (s->peerOn)
    {   // blockers of the bands above arrive in this context's mailbox; rank 0 of the group (peerWait off) starts from zero
        unsigned long long* loc = (unsigned long long*)s->peerLocal.p;
        FGL_CUDA(c, cudaMemsetAsync(loc, 0, 24, st));  // count of the bands above, running total, error word: all per frame
        if (s->peerWait)
        {
            LaunchScope ls(c, "pcss_peer_wait", 0);
            k_peer_wait<<<1, 1, 0, st>>>((const unsigned long long*)s->mailbox.p, s->peerEpoch, loc, (unsigned*)(loc + 2));
        }
    }
    if (nU > 0)
    {
        R.nU = nU, R.Upix = (const unsigned*)s->Upix.p, R.Uc1 = (const unsigned*)s->Uc1.p, R.Usc = (const float4*)s->Usc.p;
        R.UF = (const unsigned long long*)s->UF.p, R.UE = (const unsigned long long*)s->UE.p;
        R.base = chunkBase;
        R.stats = s->chainStats.p && getenv("FGL_CHAIN_STATS") ? (unsigned long long*)s->chainStats.p : nullptr;
        if (int rc = launch_chain(c, s, R, nU, n, nC1, st, /*shared=*/phase == FGL_VIS_LAUNCH)) return rc;
    }
    if (s->peerOn)
    {   // the band below can start as soon as this band's chain has finished: signal it before anything else is queued
        unsigned long long* loc = (unsigned long long*)s->peerLocal.p;
        LaunchScope         ls(c, "pcss_peer_notify", 0);
        k_peer_notify<<<1, 1, 0, st>>>(s->nextMailbox, s->peerEpoch, loc, (const unsigned*)s->mState.p, blockersBefore, (unsigned)nC1, nU > 0 ? 1 : 0, loc + 1,
                                       (const unsigned*)(loc + 2));
    }
    }  // !inflight
    if (phase == FGL_VIS_LAUNCH)
    {
        s->chainInFlight = true, s->inflightBefore = blockersBefore;
        s->prepValid = true, s->prepTotal = nTotal, s->prepLo = siteLo, s->prepHi = siteHi, s->prepNU = nU, s->prepNC1 = nC1;
        return FGL_OK;
    }
    if (s->peerOn)
    {   // the wait for the band above may have timed out, or a band above may have failed: a wrong count must never be used silently
        unsigned long long pl[3] = { 0, 0, 0 };
        FGL_CUDA(c, cudaMemcpyAsync(pl, s->peerLocal.p, 24, cudaMemcpyDeviceToHost, st));
        FGL_CUDA(c, cudaStreamSynchronize(st));
        if (pl[2])
            return fgl_fail(c, FGL_ERR_STATE, pl[2] == 1 ? "PCSS chain hand-off: the band above never signalled (30 s)" : "PCSS chain hand-off: a band above failed");
    }
    if (nU > 0)
    {
        unsigned hs[8];
        FGL_CUDA(c, cudaMemcpyAsync(hs, s->mState.p, 32, cudaMemcpyDeviceToHost, st));
        c->d2hBytes += 32;
        FGL_CUDA(c, cudaStreamSynchronize(st));
        if ((hs[CH_ERROR] || !hs[CH_DONE]) && s->chainWasShared && !s->peerOn)
        {   // a CTA of the shared (plain-launch) build was kept off its SM beyond the barriers' time-outs: once more, exclusively
            R.nU = nU, R.Upix = (const unsigned*)s->Upix.p, R.Uc1 = (const unsigned*)s->Uc1.p, R.Usc = (const float4*)s->Usc.p;
            R.UF = (const unsigned long long*)s->UF.p, R.UE = (const unsigned long long*)s->UE.p, R.base = chunkBase, R.stats = nullptr;
            if (int rc = launch_chain(c, s, R, nU, n, nC1, st, false)) return rc;
            FGL_CUDA(c, cudaMemcpyAsync(hs, s->mState.p, 32, cudaMemcpyDeviceToHost, st));
            FGL_CUDA(c, cudaStreamSynchronize(st));
        }
        c->lastChainIters = (int)hs[CH_ITERS];
        uncertainBlockers = hs[CH_M0];
        if (hs[CH_ERROR] || !hs[CH_DONE]) return fgl_fail(c, FGL_ERR_STATE, "PCSS chain: persistent kernel failed (code " + std::to_string(hs[CH_ERROR]) + ")");
        if (s->chainStats.p && getenv("FGL_CHAIN_STATS"))
        {
            unsigned long long hq[16];
            FGL_CUDA(c, cudaMemcpyAsync(hq, s->chainStats.p, 128, cudaMemcpyDeviceToHost, st));
            FGL_CUDA(c, cudaStreamSynchronize(st));
            fprintf(stderr, "[chain stats] sites=%zu uncertain=%d iters=%d | pilot pairs=%llu one=%llu ambiguous=%llu taps_in_E=%llu\n", n, nU, c->lastChainIters,
                    hq[0], hq[1], hq[2], hq[3]);
        }
    }
    {
        LaunchScope ls(c, "pcss_flags", n * 16);
        k_pixel_flags<<<nb, 256, 0, st>>>(n, (const int*)s->isU.p, (const int*)s->isC1.p, (const int*)s->posU.p, (const uint8_t*)s->flagU.p, (int*)s->hasB.p);
    }
    if (int rc = scan_ints(c, (const int*)s->hasB.p, (int*)s->kpre.p, n)) return rc;
    {
        LaunchScope ls(c, "pcss_chunk_index", n * 20);
        DeepShadow D;
        D.isC1 = (const int*)s->isC1.p, D.sc4 = sc4In, D.smMin = (const float*)s->smMin.p, D.smMax = (const float*)s->smMax.p, D.sm = L.sm, D.r = r;
        D.pcfFilter = L.pcfFilter, D.areaLight = L.areaLight;
        FGL_CUDA(c, cudaMemsetAsync((unsigned*)s->mState.p + CH_NFILTERED, 0, 4, st));
        k_chunk_index<<<nb, 256, 0, st>>>(n, chunkBase, (const int*)s->kpre.p, (const int*)s->hasB.p, chunkOfB, visB, (unsigned*)s->blockerList.p,
                                          (unsigned*)s->mState.p + CH_NBLOCKERS, D, R.kDev, (unsigned*)s->mState.p + CH_NFILTERED);
    }
    {
        // bytes of the entries that are actually filtered (counted on the device by k_chunk_index): coordinate + chunk index + 96 samples + result
        LaunchScope ls(c, "pcss_visibility", 0, (const unsigned*)s->mState.p + CH_NFILTERED, 16 + 4 + 768 + 4);
        // (the fetch-ahead variant is measured, not the default: 0.202 vs 0.186 ms at C3, 0.168 vs 0.155 ms at C5 — profiles/r02c_*)
        static const bool pipelined = getenv("FGL_VIS_PIPELINE") && atoi(getenv("FGL_VIS_PIPELINE")) != 0;
        const unsigned*   nList = (const unsigned*)s->mState.p + CH_NFILTERED;  // the list is dense: its length is the number of filtered pixels
        if (pipelined)
            k_pcss_visibility_pipe<<<c->numSMs * 6, 256, 0, st>>>((const unsigned*)s->blockerList.p, nList, sc4In, chunkOfB, L.sm, L.disk, L.pcssFilter, L.pcfFilter,
                                                                 L.areaLight, visB);
        else
            k_pcss_visibility<<<c->numSMs * 8, 256, 0, st>>>((const unsigned*)s->blockerList.p, nList, sc4In, chunkOfB, L.sm, L.disk, L.pcssFilter, L.pcfFilter,
                                                            L.areaLight, visB);
    }
    L.chunkOf = (const unsigned*)s->chunkOf.p;
    // known on the host as soon as the chain kernel has finished — a sort-first driver can hand it to the next band while
    // this band's filter and lighting kernels are still running
    s->chainTotal = blockersBefore + (unsigned long long)nC1 + uncertainBlockers, s->chainCountValid = true;
    s->peerTotalOnDevice = s->peerOn;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fgl_fail(c, FGL_ERR_CUDA, std::string("pcss chain: ") + cudaGetErrorString(e));
    return FGL_OK;
}

// Deferred lighting: the sites are the pixels in scan order (forkergl.cpp:336-378 visits every pixel, background included).
int fgl_stream_prepare_lighting(fgl_ctx* c, LightPass& L, int phase)
{
    SampleStream* s = S_of(c);
    size_t        n = (size_t)L.W * L.H;
    if (int rc = fgl_reserve(c, s->sc4, n * 16)) return rc;
    ChainPass P;
    memset(&P, 0, sizeof P);
    P.W = L.W, P.H = L.H;
    P.worldpos = L.planes.p[FGL_PLANE_WORLDPOS], P.normal = L.planes.p[FGL_PLANE_NORMAL], P.lndc = L.planes.p[FGL_PLANE_LIGHTNDC];
    memcpy(P.lightPos, L.lightPos, 12);
    P.biasSlope = L.biasSlope, P.biasMin = L.biasMin, P.sm = L.sm;
    size_t lo = (size_t)L.row0 * L.W, hi = (size_t)L.row1 * L.W;
    if (!((phase == FGL_VIS_RESOLVE || phase == FGL_VIS_LAUNCH) && s->prepValid) && hi > lo)
    {
        const bool pcss = L.shadowMode == FGL_SHADOW_PCSS;
        static const bool noRect = getenv("FGL_NO_MAP_RECT") != nullptr;
        s->rectValid = pcss && !noRect;
        if (s->rectValid)
        {
            if (int rc = fgl_reserve(c, s->rect, 32)) return rc;
            P.fsF = pcss_filter_bound(L.pcssFilter);
            k_rect_init<<<1, 1, 0, c->stream>>>((int*)s->rect.p);
        }
        LaunchScope ls(c, "shadow_coords", (hi - lo) * (36 + 16));
        const unsigned blocks = (unsigned)std::min<size_t>((hi - lo + 255) / 256, (size_t)c->numSMs * 16);
        k_shadow_coords<<<blocks, 256, 0, c->stream>>>(P, lo, hi, (float4*)s->sc4.p, s->rectValid ? (int*)s->rect.p : nullptr);
    }
    // sort-first bands: the chain of this band starts from the number of blockers found in the bands before it
    return fgl_stream_site_visibility(c, L, n, (const float4*)s->sc4.p, lo, hi, c->chainBlockersBefore, phase);
}

// Blockers found up to and including this context's band (= input of the next band's chain).  Host hand-off: known without
// touching the stream; device-side hand-off: the running total lives on the device (blocks on the stream).
int fgl_stream_chain_total(fgl_ctx* c, unsigned long long* out)
{
    SampleStream* s = S_of(c);
    if (!s->chainCountValid) return fgl_fail(c, FGL_ERR_STATE, "no PCSS chain has run on this context");
    *out = s->chainTotal;
    if (s->peerTotalOnDevice)
    {
        unsigned long long v[3] = { 0, 0, 0 };
        FGL_CUDA(c, cudaMemcpyAsync(v, s->peerLocal.p, 24, cudaMemcpyDeviceToHost, c->stream));
        FGL_CUDA(c, cudaStreamSynchronize(c->stream));
        if (v[2]) return fgl_fail(c, FGL_ERR_STATE, "PCSS chain hand-off: the band above never signalled (30 s)");
        *out = v[1];
    }
    return FGL_OK;
}

bool fgl_stream_peer_on(fgl_ctx* c) { return c->stream_state && c->stream_state->peerOn; }

// ---- peer hand-off set-up ---------------------------------------------------------------------------------------------
static int peer_buffers(fgl_ctx* c, SampleStream* s)
{
    if (s->mailbox.p) return FGL_OK;
    if (int rc = fgl_reserve(c, s->mailbox, 16 * 2 * 8)) return rc;
    if (int rc = fgl_reserve(c, s->peerLocal, 64)) return rc;
    FGL_CUDA(c, cudaMemset(s->mailbox.p, 0, 16 * 2 * 8));
    FGL_CUDA(c, cudaMemset(s->peerLocal.p, 0, 64));
    return FGL_OK;
}
int fgl_stream_peer_mailbox(fgl_ctx* c, void** devPtr, void* ipcHandle64)
{
    SampleStream* s = S_of(c);
    if (int rc = peer_buffers(c, s)) return rc;
    if (devPtr) *devPtr = s->mailbox.p;
    if (ipcHandle64)
    {
        cudaIpcMemHandle_t h;
        FGL_CUDA(c, cudaIpcGetMemHandle(&h, s->mailbox.p));
        static_assert(sizeof h == 64, "CUDA IPC handles are 64 bytes");
        memcpy(ipcHandle64, &h, 64);
    }
    return FGL_OK;
}
int fgl_stream_peer_connect(fgl_ctx* c, void* nextDevPtr, const void* nextIpcHandle64, int waitPrev, int enable)
{
    SampleStream* s = S_of(c);
    if (s->nextMailboxIpc) cudaIpcCloseMemHandle(s->nextMailboxIpc), s->nextMailboxIpc = nullptr;
    s->nextMailbox = nullptr, s->peerOn = false, s->peerWait = false;
    if (!enable) return FGL_OK;
    if (int rc = peer_buffers(c, s)) return rc;
    // Epochs restart at 0 with every connect, so slots of an earlier session must not survive it.  The band above writes into
    // this mailbox: every context of the group has to return from this call before any of them renders a frame (a group
    // barrier; forkerrenderer_b200/multigpu.py and frh_group_connect have one).
    FGL_CUDA(c, cudaMemset(s->mailbox.p, 0, 16 * 2 * 8));
    FGL_CUDA(c, cudaMemset(s->peerLocal.p, 0, 64));
    if (nextIpcHandle64)
    {
        cudaIpcMemHandle_t h;
        memcpy(&h, nextIpcHandle64, 64);
        void* p = nullptr;
        FGL_CUDA(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        s->nextMailboxIpc = p, s->nextMailbox = (unsigned long long*)p;
    }
    else s->nextMailbox = (unsigned long long*)nextDevPtr;
    s->peerOn = true, s->peerWait = waitPrev != 0, s->peerEpoch = 0;
    return FGL_OK;
}
