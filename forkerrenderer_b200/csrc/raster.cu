// raster.cu — the raster passes (shadow / geometry / forward) of the frame as sm_100a kernels.
//
//   k_setup          one thread per triangle: the vertex program x3 (reference gshader.h:41-92 == phongshader.h:35-85
//                    == pbrshader.h:35-85, depthshader.h:21-28), viewport + integer snap + clamped bounding box
//                    (forkergl.cpp:241-255), classification for the two raster paths.
//   k_raster_small   one thread per small triangle; k_raster_blocks: one warp per 32x32 block of a large triangle's
//                    bounding box.  Coverage is the reference's double-precision barycentric test evaluated
//                    literally (geometry.cpp:20-56), pre-filtered by exact integer edge functions; the depth test
//                    (forkergl.cpp:180-200, strict-less, first submitted wins ties) is a 64-bit atomicMin on
//                    (orderable(depth) << 32 | primitive id), which makes the result independent of scheduling.
//   k_resolve_*      one thread per pixel: decode the winner, redo its barycentrics, run the fragment program
//                    (depthshader.h:30-36, gshader.h:95-201, phongshader.h:90-169, pbrshader.h:90-180) and write the
//                    SoA planes; pixels without a winner receive the planes' clear values (fused clear).
#include <cub/block/block_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>

#include "fgl_internal.h"

namespace
{
constexpr int kSmallArea = 32;  // clamped bbox area up to which a triangle takes the thread-per-triangle path
constexpr int kBlk = 32;        // block edge of the warp path
constexpr int kSaneCoord = 1 << 20;

__device__ __forceinline__ const DrawCmdD& find_draw(const DrawCmdD* draws, int nDraws, int prim)
{
    int lo = 0, hi = nDraws - 1;
    while (lo < hi)
    {
        int mid = (lo + hi + 1) >> 1;
        if (draws[mid].firstPrim <= prim) lo = mid;
        else hi = mid - 1;
    }
    return draws[lo];
}

__device__ __forceinline__ V3 ld3(const float* a, int i) { return v3(__ldg(a + 3 * (size_t)i), __ldg(a + 3 * (size_t)i + 1), __ldg(a + 3 * (size_t)i + 2)); }

struct SetupRegs
{
    int   X[3], Y[3];
    float d[3];
    int   xmin, xmax, ymin, ymax, flags;
};
__device__ __forceinline__ SetupRegs load_setup(const TriSetup* s, int prim)
{
    const int4* p = reinterpret_cast<const int4*>(s + prim);
    int4        a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    SetupRegs   r;
    r.X[0] = a.x, r.X[1] = a.y, r.X[2] = a.z, r.Y[0] = a.w;
    r.Y[1] = b.x, r.Y[2] = b.y, r.d[0] = __int_as_float(b.z), r.d[1] = __int_as_float(b.w);
    r.d[2] = __int_as_float(c.x);
    r.xmin = c.y & 0xffff, r.xmax = (c.y >> 16) & 0xffff;
    r.ymin = c.z & 0xffff, r.ymax = (c.z >> 16) & 0xffff;
    r.flags = c.w;
    return r;
}

// -------------------------------------------------------------------------------------------------------------
// The work items of k_raster_blocks (32 x 32 blocks of the large triangles' bounding boxes) are numbered by a two-level
// exclusive scan of the per-triangle block counts: inside k_setup every CTA (a tile of kSetupTile triangles) scans its own
// counts (blkScan[i] = blocks of the tile's triangles in front of triangle i) and stores its sum; k_tile_scan then scans the
// tile sums (one small CTA: a 10 M-triangle pass has 78 k tiles).  item -> (tile, triangle) is two short binary searches.
constexpr int kSetupTile = 128;

__global__ void __launch_bounds__(kSetupTile) k_setup(RasterPass P, int primBegin, int smallArea, int* tileBase)
{
    __shared__ int sWarpSum[kSetupTile / 32];
    const int  tile = blockIdx.x;
    const int  prim = primBegin + tile * kSetupTile + (int)threadIdx.x;
    const bool live = prim < P.nPrims;
    const DrawCmdD& d = find_draw(P.draws, P.nDraws, live ? prim : P.nPrims - 1);
    int             face = (live ? prim : P.nPrims - 1) - d.firstPrim;
    bool            shadowPass = P.passType == FGL_PASS_SHADOW;

    // ---- positions first: clip space, viewport, integer snap, bounding box (forkergl.cpp:241-255) — enough to decide whether
    // the triangle can touch this context's rows at all; the attributes follow only for the triangles that do
    V4    ndc[3], ws[3];
    float zn[3] = { 0.f, 0.f, 0.f }, oow[3] = { 0.f, 0.f, 0.f };
    int   pidx[3] = { 0, 0, 0 };
    if (d.preNdc)
    {   // fgl_draw_triangles: the caller ran the vertex program (Shader::ProcessVertex on the host)
        const float* q = d.preNdc + (size_t)face * 12;
#pragma unroll
        for (int k = 0; k < 3; ++k) ndc[k].x = __ldg(q + 4 * k), ndc[k].y = __ldg(q + 4 * k + 1), ndc[k].z = __ldg(q + 4 * k + 2), ndc[k].w = __ldg(q + 4 * k + 3);
        if (shadowPass)
        {
#pragma unroll
            for (int k = 0; k < 3; ++k) zn[k] = __ldg(d.preZ + (size_t)face * 3 + k);
        }
    }
    else
    {
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            pidx[k] = __ldg(d.pi + face * 3 + k);
            V3 p = ld3(d.pos, pidx[k]);
            V4 p4;
            p4.x = p.x, p4.y = p.y, p4.z = p.z, p4.w = 1.f;
            if (d.kind == FGL_SHADER_DEPTH)
            {
                V4 cs = mat4mul(d.lm, p4);
                ndc[k] = vdivs4(cs, cs.w);
                zn[k] = ndc[k].z;
                continue;
            }
            ws[k] = mat4mul(d.model, p4);
            V4 vs = mat4mul(d.view, ws[k]);
            V4 cs = mat4mul(d.proj, vs);
            oow[k] = 1.f / cs.w;
            ndc[k] = vdivs4(cs, cs.w);
        }
    }
    int   X[3], Y[3];
    float dz[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        V4 s = mat4mul(P.viewport, ndc[k]);
        X[k] = f2i_x86(s.x), Y[k] = f2i_x86(s.y);
        dz[k] = s.z;
    }
    int mnx = min(X[0], min(X[1], X[2])), mxx = max(X[0], max(X[1], X[2]));
    int mny = min(Y[0], min(Y[1], Y[2])), mxy = max(Y[0], max(Y[1], Y[2]));
    int xmin = clampi(mnx, 0, P.W - 1), xmax = clampi(mxx, 0, P.W - 1);
    int ymin = clampi(mny, 0, P.H - 1), ymax = clampi(mxy, 0, P.H - 1);
    // sort-first group: the pixel loop of the reference (forkergl.cpp:165-171) restricted to this context's rows
    ymin = max(ymin, P.cull0), ymax = min(ymax, P.cull1 - 1);

    TriCover tc = make_cover(X[0], Y[0], X[1], Y[1], X[2], Y[2]);
    int      flags = 0;
    bool     sane = max(max(abs(mnx), abs(mxx)), max(abs(mny), abs(mxy))) <= kSaneCoord && mnx != (int)0x80000000 && mny != (int)0x80000000;
    if (!tc.valid) flags = TRI_SKIP;
    // a triangle whose own bounding box misses the buffer only scans pixels outside itself (all rejected)
    if (sane && (mxx < 0 || mnx > P.W - 1 || mxy < 0 || mny > P.H - 1)) flags = TRI_SKIP;
    if (ymax < ymin) flags = TRI_SKIP;  // no row of this context's band
    int bw = xmax - xmin + 1, bh = ymax - ymin + 1, nb = 0;
    if (!(flags & TRI_SKIP))
    {
        if (bw * bh <= smallArea) flags |= TRI_SMALL;
        else
        {
            flags |= TRI_LARGE;
            nb = ((bw + kBlk - 1) / kBlk) * ((bh + kBlk - 1) / kBlk);
        }
        if (sane) flags |= 8;
    }
    if (!live) nb = 0;
    {   // blkScan[i] = blocks of the tile's triangles in front of triangle primBegin + i; tileBase[tile] = blocks of the tile
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        int       incl = nb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) sWarpSum[wid] = incl;
        __syncthreads();
        int excl = incl - nb, sum = 0;
#pragma unroll
        for (int w = 0; w < kSetupTile / 32; ++w)
        {
            if (w < wid) excl += sWarpSum[w];
            sum += sWarpSum[w];
        }
        if (live) P.blkScan[prim - primBegin] = excl;  // tile-local
        if (threadIdx.x == 0) tileBase[tile] = sum;
    }
    if (!live) return;
    P.nblk[prim] = nb;
    int4* so = reinterpret_cast<int4*>(P.setup + prim);
    if (flags & TRI_SKIP)
    {   // nothing else of a skipped triangle is ever read
        so[2] = make_int4(0, 0, 0, flags);
        return;
    }
    so[0] = make_int4(X[0], X[1], X[2], Y[0]);
    so[1] = make_int4(Y[1], Y[2], __float_as_int(dz[0]), __float_as_int(dz[1]));
    so[2] = make_int4(__float_as_int(dz[2]), xmin | (xmax << 16), ymin | (ymax << 16), flags);
    if (shadowPass)
    {
        P.zndc[prim] = make_float4(zn[0], zn[1], zn[2], 0.f);
        return;
    }
    float4* vo = reinterpret_cast<float4*>(P.vary + prim);
    if (d.preNdc)
    {
        const float4* q = reinterpret_cast<const float4*>(d.preVary + (size_t)face * 48);
#pragma unroll
        for (int i = 0; i < 12; ++i) vo[i] = __ldg(q + i);
        return;
    }
    // ---- attributes of the vertex program (gshader.h:41-92 == phongshader.h:35-85 == pbrshader.h:35-85), each times 1 / w_clip
    TriVary vy;
#pragma unroll
    for (int i = 0; i < 48; ++i) vy.f[i] = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        int   tidx = __ldg(d.ti + face * 3 + k);
        float tu = __ldg(d.uv + 2 * (size_t)tidx), tv = __ldg(d.uv + 2 * (size_t)tidx + 1);
        V3    nWS = mat3mul(d.normal, vnormalize(ld3(d.nrm, __ldg(d.ni + face * 3 + k))));  // mesh.cpp:46-50
        float w = oow[k];
        vy.f[42 + k] = w;
        vy.f[0 + 3 * k] = ws[k].x * w, vy.f[1 + 3 * k] = ws[k].y * w, vy.f[2 + 3 * k] = ws[k].z * w;
        vy.f[36 + k] = tu * w, vy.f[39 + k] = tv * w;
        vy.f[9 + 3 * k] = nWS.x * w, vy.f[10 + 3 * k] = nWS.y * w, vy.f[11 + 3 * k] = nWS.z * w;
        if (d.hasTangents)
        {
            V3 tWS = mat3mul(d.normal, vnormalize(ld3(d.tan, pidx[k])));
            vy.f[18 + 3 * k] = tWS.x * w, vy.f[19 + 3 * k] = tWS.y * w, vy.f[20 + 3 * k] = tWS.z * w;
        }
        if (P.shadowOn)
        {
            V4 ls = mat4mul(d.lightSpace, ws[k]);
            ls = vdivs4(ls, ls.w);
            vy.f[27 + 3 * k] = ls.x * w, vy.f[28 + 3 * k] = ls.y * w, vy.f[29 + 3 * k] = ls.z * w;
        }
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) vo[i] = make_float4(vy.f[4 * i], vy.f[4 * i + 1], vy.f[4 * i + 2], vy.f[4 * i + 3]);
}

// the tile-local half of the scan for counts that already exist (all triangles of a pass that was flushed in pieces)
__global__ void __launch_bounds__(kSetupTile) k_tile_local_scan(const int* nblk, int* blkScan, int* tileBase, int n)
{
    __shared__ int sWarpSum[kSetupTile / 32];
    const int i = blockIdx.x * kSetupTile + threadIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nb = i < n ? nblk[i] : 0;
    int       incl = nb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) sWarpSum[wid] = incl;
    __syncthreads();
    int excl = incl - nb, sum = 0;
#pragma unroll
    for (int w = 0; w < kSetupTile / 32; ++w)
    {
        if (w < wid) excl += sWarpSum[w];
        sum += sWarpSum[w];
    }
    if (i < n) blkScan[i] = excl;
    if (threadIdx.x == 0) tileBase[blockIdx.x] = sum;
}

// in-place exclusive scan of the nTiles tile sums, total behind them: one CTA walks the array in coalesced chunks of 4096
// (four consecutive values per thread), carrying the running sum
__global__ void __launch_bounds__(1024) k_tile_scan(int* tileBase, int nTiles)
{
    __shared__ int sWarp[32], sCarry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) sCarry = 0;
    __syncthreads();
    for (int base = 0; base < nTiles; base += 4096)
    {
        const int i0 = base + 4 * (int)threadIdx.x;
        int       v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = i0 + k < nTiles ? tileBase[i0 + k] : 0;
        const int sum = v[0] + v[1] + v[2] + v[3];
        int       incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) sWarp[wid] = incl;
        __syncthreads();
        const int carry = sCarry;
        int       before = 0, all = 0;
#pragma unroll
        for (int w = 0; w < 32; ++w)
        {
            const int ws = sWarp[w];
            if (w < wid) before += ws;
            all += ws;
        }
        int run = carry + before + incl - sum;
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            if (i0 + k < nTiles) tileBase[i0 + k] = run;
            run += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) sCarry = carry + all;
        __syncthreads();
    }
    if (threadIdx.x == 0) tileBase[nTiles] = sCarry;
}

// -------------------------------------------------------------------------------------------------------------
// Exact integer edge functions (valid while |coords| <= 2^20): a pixel can pass the reference's test only if
// all three are >= 0 after multiplying by the sign of the signed area (DESIGN.md "coverage pre-filter").
struct EdgeInt
{
    long long e0, e1, e2;        // values at the current pixel
    long long dx0, dx1, dx2;     // increments for px + 1
    long long dy0, dy1, dy2;     // increments for py + 1
};
__device__ __forceinline__ EdgeInt make_edges(const SetupRegs& s, int px, int py)
{
    long long ax = s.X[0], ay = s.Y[0], bx = s.X[1], by = s.Y[1], cx = s.X[2], cy = s.Y[2];
    long long rz = (bx - ax) * (cy - ay) - (cx - ax) * (by - ay);
    long long sg = rz < 0 ? -1 : 1;
    long long rx = (cx - ax) * (ay - py) - (ax - px) * (cy - ay);
    long long ry = (ax - px) * (by - ay) - (bx - ax) * (ay - py);
    EdgeInt   e;
    e.e1 = sg * rx, e.e2 = sg * ry, e.e0 = sg * (rz - rx - ry);
    e.dx1 = sg * (cy - ay), e.dy1 = -sg * (cx - ax);
    e.dx2 = -sg * (by - ay), e.dy2 = sg * (bx - ax);
    e.dx0 = -(e.dx1 + e.dx2), e.dy0 = -(e.dy1 + e.dy2);
    return e;
}

// MODE 0: depth test.  MODE 1 / 2 (forward mode with a stochastic shadow filter): count / record EVERY covered
// fragment per pixel, because each fragment that passes the depth test at submission time consumes samples of the
// stream (phongshader.h:139-149), not only the final winner.
enum { RM_DEPTH = 0, RM_COUNT = 1, RM_FILL = 2 };
template <int MODE, bool PRECHECK>
__device__ __forceinline__ void depth_test_pixel(const RasterPass& P, const TriCover& tc, const SetupRegs& s, int prim, int px, int py)
{
    V3 bary;
    if (!cover_test(tc, px, py, bary)) return;
    float z = vdot(bary, v3(s.d[0], s.d[1], s.d[2]));  // forkergl.cpp:180
    if (!(z < 3.402823466e+38f)) return;                // never below the FLT_MAX clear value (forkergl.cpp:189)
    if (MODE == RM_COUNT)
    {
        atomicAdd(P.fragCount + (size_t)px + (size_t)py * P.W, 1u);
        return;
    }
    if (MODE == RM_FILL)
    {
        size_t   pix = (size_t)px + (size_t)py * P.W;
        unsigned slot = P.fragOffset[pix] + atomicAdd(P.fragCount + pix, 1u);
        P.frags[slot] = ((unsigned long long)(unsigned)prim << 32) | __float_as_uint(z);
        return;
    }
    unsigned long long  key = ((unsigned long long)depth_to_ordered(z) << 32) | (unsigned)prim;
    unsigned long long* cell = P.vis + (size_t)px + (size_t)py * P.W;
    // the pre-read saves atomics under overdraw but serialises a load in front of each one
    if (!PRECHECK || key < *((volatile unsigned long long*)cell)) atomicMin(cell, key);
}

template <int MODE>
__global__ void __launch_bounds__(128) k_raster_small(RasterPass P, int primBegin)
{
    int prim = primBegin + blockIdx.x * blockDim.x + threadIdx.x;
    if (prim >= P.nPrims) return;
    int flags = __ldg(&reinterpret_cast<const int4*>(P.setup + prim)[2].w);
    if (!(flags & TRI_SMALL)) return;
    SetupRegs s = load_setup(P.setup, prim);
    TriCover  tc = make_cover(s.X[0], s.Y[0], s.X[1], s.Y[1], s.X[2], s.Y[2]);
    bool      sane = flags & 8;
    EdgeInt   e = make_edges(s, s.xmin, s.ymin);
    for (int px = s.xmin; px <= s.xmax; ++px)
    {
        long long c0 = e.e0, c1 = e.e1, c2 = e.e2;
        for (int py = s.ymin; py <= s.ymax; ++py)
        {
            if (!sane || (c0 | c1 | c2) >= 0) depth_test_pixel<MODE, true>(P, tc, s, prim, px, py);
            c0 += e.dy0, c1 += e.dy1, c2 += e.dy2;
        }
        e.e0 += e.dx0, e.e1 += e.dx1, e.e2 += e.dx2;
    }
}

// kLanesPerItem lanes share one work item (a 32 x 32 block of one triangle's bounding box).  32: a lane per column — the
// few huge triangles of the reference's scenes (the ground plane) fill their blocks.  8: four items per warp — most blocks of a
// high-triangle-count pass are a few pixels wide, and the per-item part (search, set-up record, the double-precision edge
// set-up) is shared by the lanes of an item; a lane then walks the columns l8, l8 + 8, ... of the block.
template <int MODE, int kLanesPerItem>
__global__ void __launch_bounds__(256) k_raster_blocks(RasterPass P, int primBegin, int nNew, const int* tileBase)
{
    const int l8 = threadIdx.x & (kLanesPerItem - 1);
    const int groupsPerGrid = (gridDim.x * blockDim.x) / kLanesPerItem;
    const int nTiles = (nNew + kSetupTile - 1) / kSetupTile;
    const int total = tileBase[nTiles];
    int       cachedTri = -1, cachedBegin = 0, cachedEnd = 0;
    SetupRegs s;
    TriCover  tc;
    for (int item = (blockIdx.x * blockDim.x + threadIdx.x) / kLanesPerItem; item < total; item += groupsPerGrid)
    {
        if (item < cachedBegin || item >= cachedEnd)
        {   // last tile whose first item is <= item, then the last triangle of that tile whose first item is <= item
            int lo = 0, hi = nTiles - 1;
            while (lo < hi)
            {
                int mid = (lo + hi + 1) >> 1;
                if (__ldg(tileBase + mid) <= item) lo = mid;
                else hi = mid - 1;
            }
            const int tile = lo, base = __ldg(tileBase + tile), t0 = tile * kSetupTile, rel = item - base;
            lo = 0, hi = min(kSetupTile, nNew - t0) - 1;
            while (lo < hi)
            {
                int mid = (lo + hi + 1) >> 1;
                if (__ldg(P.blkScan + t0 + mid) <= rel) lo = mid;
                else hi = mid - 1;
            }
            const int i = t0 + lo;
            cachedTri = primBegin + i;
            cachedBegin = base + __ldg(P.blkScan + i);
            cachedEnd = cachedBegin + __ldg(P.nblk + cachedTri);
            s = load_setup(P.setup, cachedTri);
            tc = make_cover(s.X[0], s.Y[0], s.X[1], s.Y[1], s.X[2], s.Y[2]);
        }
        int local = item - cachedBegin;
        int nbx = (s.xmax - s.xmin + kBlk) / kBlk;
        int bx = local % nbx, by = local / nbx;
        int x0 = s.xmin + bx * kBlk, x1 = min(x0 + kBlk - 1, s.xmax);
        int y0 = s.ymin + by * kBlk, y1 = min(y0 + kBlk - 1, s.ymax);
        if (x0 + l8 > x1) continue;
        bool    sane = s.flags & 8;
        EdgeInt col = make_edges(s, x0 + l8, y0);
        for (int px = x0 + l8; px <= x1; px += kLanesPerItem)  // (a single trip with a lane per column)
        {
            long long e0 = col.e0, e1 = col.e1, e2 = col.e2;
            for (int py = y0; py <= y1; ++py)
            {
                if (!sane || (e0 | e1 | e2) >= 0) depth_test_pixel<MODE, false>(P, tc, s, cachedTri, px, py);
                e0 += col.dy0, e1 += col.dy1, e2 += col.dy2;
            }
            if (kLanesPerItem == kBlk) break;
            col.e0 += kLanesPerItem * col.dx0, col.e1 += kLanesPerItem * col.dx1, col.e2 += kLanesPerItem * col.dx2;
        }
    }
}

// -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool decode_winner(const RasterPass& P, size_t idx, int px, int py, int& prim, V3& bary, float& depth)
{
    unsigned long long key = P.vis[idx];
    if (key == FGL_VIS_EMPTY) return false;
    prim = (int)(unsigned)(key & 0xffffffffull);
    SetupRegs s = load_setup(P.setup, prim);
    TriCover  tc = make_cover(s.X[0], s.Y[0], s.X[1], s.Y[1], s.X[2], s.Y[2]);
    cover_test(tc, px, py, bary);
    depth = vdot(bary, v3(s.d[0], s.d[1], s.d[2]));
    return true;
}

__global__ void __launch_bounds__(256) k_resolve_shadow(RasterPass P, float* shadowPlane, float* depthPlane)
{
    size_t n = (size_t)P.row1 * P.W;  // rows [row0, row1): everything, or this context's rows of the map in a sort-first group
    size_t idx = (size_t)P.row0 * P.W + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    int   px = (int)(idx % P.W), py = (int)(idx / P.W);
    int   prim;
    V3    bary;
    float depth, sh = 0.f;
    if (decode_winner(P, idx, px, py, prim, bary, depth))
    {
        float4 z = __ldg(P.zndc + prim);
        sh = interp(z.x, z.y, z.z, bary) * 0.5f + 0.5f;  // depthshader.h:30-36
    }
    else depth = 3.402823466e+38f;
    depthPlane[idx] = depth;
    if (!P.group) shadowPlane[idx] = sh;
    else
    {   // fused all-gather: the texel goes into every context's shadow map (peer stores over NVLink; [rank] is this context's own)
        for (int r = 0; r < P.peers.n; ++r) P.peers.p[r][idx] = sh;
    }
}

struct Surface
{
    V3    posWS, normal, lightNDC;
    float u, v;
};

// common head of the camera-space fragment programs: gshader.h:95-146 == phongshader.h:90-128 == pbrshader.h:90-128
__device__ __forceinline__ Surface interpolate_surface(const RasterPass& P, const DrawCmdD& d, const float* vy, V3 bary, int normalMapId)
{
    Surface s;
    V3      pos = v3(interp(vy[0], vy[3], vy[6], bary), interp(vy[1], vy[4], vy[7], bary), interp(vy[2], vy[5], vy[8], bary));
    float   tu = interp(vy[36], vy[37], vy[38], bary), tv = interp(vy[39], vy[40], vy[41], bary);
    V3      nrm = v3(interp(vy[9], vy[12], vy[15], bary), interp(vy[10], vy[13], vy[16], bary), interp(vy[11], vy[14], vy[17], bary));
    float   w = 1.f / vdot(v3(vy[42], vy[43], vy[44]), bary);
    pos = vscale(pos, w);
    tu *= w, tv *= w;
    nrm = vscale(nrm, w);
    V3 N = vnormalize(nrm);
    V3 normal = N;
    if (d.hasTangents && normalMapId >= 0)
    {
        V3 tg = v3(interp(vy[18], vy[21], vy[24], bary), interp(vy[19], vy[22], vy[25], bary), interp(vy[20], vy[23], vy[26], bary));
        tg = vscale(tg, w);
        V3 T = vnormalize(vadd(tg, v3(0.001f, 0.001f, 0.001f)));
        T = vnormalize(vsub(T, vscale(N, vdot(T, N))));
        V3 B = vnormalize(vcross(N, T));
        V3 sn = tex_sample(P.textures[normalMapId], tu, tv);
        sn = vnormalize(vsub(vscale(sn, 2.f), v3(1.f, 1.f, 1.f)));
        normal = vnormalize(v3(vdot(v3(T.x, B.x, N.x), sn), vdot(v3(T.y, B.y, N.y), sn), vdot(v3(T.z, B.z, N.z), sn)));
    }
    s.posWS = pos, s.normal = normal, s.u = tu, s.v = tv;
    s.lightNDC = v3(0.f, 0.f, 0.f);
    if (P.shadowOn)
    {
        V3 l = v3(interp(vy[27], vy[30], vy[33], bary), interp(vy[28], vy[31], vy[34], bary), interp(vy[29], vy[32], vy[35], bary));
        s.lightNDC = vscale(l, w);
    }
    return s;
}

__device__ __forceinline__ void st3(float* plane, size_t n, size_t idx, V3 v)
{
    plane[idx] = v.x, plane[n + idx] = v.y, plane[2 * n + idx] = v.z;
}

__device__ __forceinline__ void load_vary(const TriVary* vary, int prim, float* vy)
{
    const float4* p = reinterpret_cast<const float4*>(vary + prim);
#pragma unroll
    for (int i = 0; i < 12; ++i)
    {
        float4 q = __ldg(p + i);
        vy[4 * i] = q.x, vy[4 * i + 1] = q.y, vy[4 * i + 2] = q.z, vy[4 * i + 3] = q.w;
    }
}

// gshader.h:95-201 + the G-buffer writes of forkergl.cpp:211-223
__global__ void __launch_bounds__(128) k_resolve_geometry(RasterPass P, PlanesD out)
{
    int px = blockIdx.x * blockDim.x + threadIdx.x, py = (P.group ? P.row0 : 0) + blockIdx.y;
    if (px >= P.W) return;
    size_t n = (size_t)P.W * P.H, idx = (size_t)px + (size_t)py * P.W;
    int    prim;
    V3     bary;
    float  depth;
    V3     z3 = v3(0.f, 0.f, 0.f);
    bool   covered = decode_winner(P, idx, px, py, prim, bary, depth);
    if (py < P.row0 || py >= P.row1)
    {   // outside this GPU's band (+ halo): only the depth plane, which SSAO gathers from anywhere
        out.p[FGL_PLANE_DEPTH][idx] = covered ? depth : 3.402823466e+38f;
        return;
    }
    if (P.group)
    {   // sort-first group: the depth plane is exchanged — rows of this context's own band go into every context's plane
        // (halo rows are owned, and delivered, by the neighbouring band)
        const float dv = covered ? depth : 3.402823466e+38f;
        if (py >= P.own0 && py < P.own1)
            for (int r = 0; r < P.peers.n; ++r) P.peers.p[r][idx] = dv;
    }
    if (!covered)
    {
        if (!P.group) out.p[FGL_PLANE_DEPTH][idx] = 3.402823466e+38f;
        st3(out.p[FGL_PLANE_NORMAL], n, idx, z3);
        st3(out.p[FGL_PLANE_WORLDPOS], n, idx, z3);
        if (P.shadowOn) st3(out.p[FGL_PLANE_LIGHTNDC], n, idx, z3);
        st3(out.p[FGL_PLANE_ALBEDO], n, idx, z3);
        st3(out.p[FGL_PLANE_EMISSIVE], n, idx, z3);
        st3(out.p[FGL_PLANE_PARAM], n, idx, z3);
        out.p[FGL_PLANE_SHADINGTYPE][idx] = 0.f;
        out.p[FGL_PLANE_AO][idx] = 1.f;
        return;
    }
    const DrawCmdD&    d = find_draw(P.draws, P.nDraws, prim);
    const FglMaterial& m = d.mat;
    float              vy[48];
    load_vary(P.vary, prim, vy);
    Surface s = interpolate_surface(P, d, vy, bary, m.normal_map);
    V3      albedo, emissive, param;
    float   type;
    if (d.supportPBR)
    {
        albedo = m.base_color_map >= 0 ? tex_sample(P.textures[m.base_color_map], s.u, s.v) : v3(m.albedo[0], m.albedo[1], m.albedo[2]);
        emissive = m.pbr_emissive_map >= 0 ? tex_sample(P.textures[m.pbr_emissive_map], s.u, s.v) : v3(m.ke[0], m.ke[1], m.ke[2]);
        float roughness = m.roughness_map >= 0 ? tex_sample_float(P.textures[m.roughness_map], s.u, s.v) : m.roughness;
        float metalness = m.metalness_map >= 0 ? tex_sample_float(P.textures[m.metalness_map], s.u, s.v) : m.metalness;
        float ao = m.ao_map >= 0 ? tex_sample_float(P.textures[m.ao_map], s.u, s.v) : 1.f;
        param = v3(ao, metalness, roughness);
        type = 1.f;
    }
    else
    {
        emissive = m.emissive_map >= 0 ? tex_sample(P.textures[m.emissive_map], s.u, s.v) : v3(m.ke[0], m.ke[1], m.ke[2]);
        albedo = m.diffuse_map >= 0 ? tex_sample(P.textures[m.diffuse_map], s.u, s.v) : v3(m.kd[0], m.kd[1], m.kd[2]);
        float shininess = m.specular_map >= 0 ? tex_sample_float(P.textures[m.specular_map], s.u, s.v) + 5 : 1.f;
        param = v3(1.f, m.ks[0], shininess);
        type = 0.f;
    }
    if (!P.group) out.p[FGL_PLANE_DEPTH][idx] = depth;
    st3(out.p[FGL_PLANE_NORMAL], n, idx, s.normal);
    st3(out.p[FGL_PLANE_WORLDPOS], n, idx, s.posWS);
    if (P.shadowOn) st3(out.p[FGL_PLANE_LIGHTNDC], n, idx, s.lightNDC);
    st3(out.p[FGL_PLANE_ALBEDO], n, idx, albedo);
    st3(out.p[FGL_PLANE_EMISSIVE], n, idx, emissive);
    st3(out.p[FGL_PLANE_PARAM], n, idx, param);
    out.p[FGL_PLANE_SHADINGTYPE][idx] = type;
    out.p[FGL_PLANE_AO][idx] = 1.f;
}

// Forward mode (phongshader.h:90-169, pbrshader.h:90-180): only the depth-test winner's colour survives in the
// reference's FrameBuffer, so the programs run once per covered pixel.  Hard shadows only need the shadow map; the
// stochastic filters additionally need each winner's position in the sample stream (see stream.cu).
__global__ void __launch_bounds__(128) k_resolve_forward(RasterPass P, PlanesD out, LightPass L)
{
    int px = blockIdx.x * blockDim.x + threadIdx.x, py = P.row0 + blockIdx.y;
    if (px >= P.W || py >= P.row1) return;
    size_t n = (size_t)P.W * P.H, idx = (size_t)px + (size_t)py * P.W;
    int    prim;
    V3     bary;
    float  depth;
    if (!decode_winner(P, idx, px, py, prim, bary, depth))
    {
        out.p[FGL_PLANE_DEPTH][idx] = 3.402823466e+38f;
        return;  // FrameBuffer keeps its clear colour
    }
    out.p[FGL_PLANE_DEPTH][idx] = depth;
    const DrawCmdD&    d = find_draw(P.draws, P.nDraws, prim);
    const FglMaterial& m = d.mat;
    float              vy[48];
    load_vary(P.vary, prim, vy);
    bool    pbr = d.kind == FGL_SHADER_PBR;
    Surface s = interpolate_surface(P, d, vy, bary, pbr ? m.pbr_normal_map : m.normal_map);
    V3      lp = v3(d.lightPos[0], d.lightPos[1], d.lightPos[2]), ep = v3(d.eye[0], d.eye[1], d.eye[2]);
    V3      lightDir = vnormalize(vsub(lp, s.posWS)), viewDir = vnormalize(vsub(ep, s.posWS));
    V3      halfwayDir = vnormalize(vadd(lightDir, viewDir));
    float   visibility = 0.f;
    if (P.shadowOn && L.vis) visibility = L.vis[P.siteOfPixel[idx]];  // PCF / PCSS at this fragment's position in the stream
    else if (P.shadowOn)
    {   // shadow.cpp:109-132, HardShadow branch
        V3    sc = vadd(vscale(s.lightNDC, 0.5f), v3(0.5f, 0.5f, 0.5f));
        float bias = fmaxf(L.biasSlope * (1.f - vdot(s.normal, lightDir)), L.biasMin);
        float sampled = shadow_lookup(L.sm, sc.x, sc.y);
        visibility = (sc.z <= sampled + bias) ? 1.f : 0.f;
    }
    LightConsts lc;
    lc.shadowIntensity = L.shadowIntensity, lc.shadowOn = P.shadowOn;
    V3 lcol = v3(d.lightColor[0], d.lightColor[1], d.lightColor[2]);
    V3 color;
    if (pbr)
    {
        V3    albedo = m.base_color_map >= 0 ? tex_sample(P.textures[m.base_color_map], s.u, s.v) : v3(m.albedo[0], m.albedo[1], m.albedo[2]);
        V3    emissive = m.pbr_emissive_map >= 0 ? tex_sample(P.textures[m.pbr_emissive_map], s.u, s.v) : v3(m.pbr_ke[0], m.pbr_ke[1], m.pbr_ke[2]);
        float roughness = m.roughness_map >= 0 ? tex_sample_float(P.textures[m.roughness_map], s.u, s.v) : m.roughness;
        float metalness = m.metalness_map >= 0 ? tex_sample_float(P.textures[m.metalness_map], s.u, s.v) : m.metalness;
        float ao = m.ao_map >= 0 ? tex_sample_float(P.textures[m.ao_map], s.u, s.v) : 1.f;
        color = pbr_light(lc, lightDir, viewDir, halfwayDir, s.normal, visibility, albedo, emissive, v3(ao, metalness, roughness), lcol);
    }
    else
    {
        V3    diffuseColor = m.diffuse_map >= 0 ? tex_sample(P.textures[m.diffuse_map], s.u, s.v) : v3(m.kd[0], m.kd[1], m.kd[2]);
        V3    emissive = m.emissive_map >= 0 ? tex_sample(P.textures[m.emissive_map], s.u, s.v) : v3(m.ke[0], m.ke[1], m.ke[2]);
        float shininess = 1.f;
        if (m.specular_map >= 0) shininess = tex_sample_float(P.textures[m.specular_map], s.u, s.v) + 5;
        color = blinn_phong_light(lc, lightDir, halfwayDir, s.normal, visibility, diffuseColor, emissive, v3(m.ka[0], m.ks[0], shininess), lcol);
    }
    st3(out.p[FGL_PLANE_FRAME], n, idx, color);
}

// ---- forward mode + stochastic shadow filter: which fragments consumed samples, and in which order ------------------
// Per pixel: sort its fragments by primitive id (= submission order), the ones that pass the depth test at
// submission time are the running minima (strict <, forkergl.cpp:189).  Pass 1 counts them, pass 2 emits
// key = prim << 32 | px << 16 | py (the reference scans px outer / py inner inside a triangle, forkergl.cpp:169-171),
// value = pixel index, top bit set for the pixel's final winner.
template <bool EMIT>
__global__ void __launch_bounds__(128) k_frag_passing(RasterPass P, unsigned* nPass, const unsigned* passOff, unsigned long long* keys, unsigned* vals)
{
    size_t n = (size_t)P.W * P.H, pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= n) return;
    unsigned            cnt = P.fragCount[pix];
    unsigned long long* f = P.frags + P.fragOffset[pix];
    if (!EMIT)
    {   // insertion sort by (prim, depth bits): prim is unique per pixel
        for (unsigned i = 1; i < cnt; ++i)
        {
            unsigned long long k = f[i];
            unsigned           j = i;
            for (; j > 0 && f[j - 1] > k; --j) f[j] = f[j - 1];
            f[j] = k;
        }
    }
    float    runMin = 3.402823466e+38f;
    unsigned np = 0, out = EMIT ? passOff[pix] : 0;
    int      px = (int)(pix % P.W), py = (int)(pix / P.W);
    unsigned total = EMIT ? nPass[pix] : 0;
    for (unsigned i = 0; i < cnt; ++i)
    {
        float z = __uint_as_float((unsigned)(f[i] & 0xffffffffull));
        if (z < runMin)
        {
            runMin = z;
            if (EMIT)
            {
                keys[out + np] = (f[i] & 0xffffffff00000000ull) | ((unsigned long long)px << 16) | (unsigned long long)py;
                vals[out + np] = (unsigned)pix | (np + 1 == total ? 0x80000000u : 0u);
            }
            ++np;
        }
    }
    if (!EMIT) nPass[pix] = np;
}

__global__ void __launch_bounds__(256) k_site_of_pixel(size_t nSites, const unsigned* vals, int* siteOfPixel)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nSites) return;
    unsigned v = vals[i];
    if (v & 0x80000000u) siteOfPixel[v & 0x7fffffffu] = (int)i;
}

// shadow coordinate + bias of every stream consumer (phongshader.h:130-149 / pbrshader.h:131-150 up to the call of
// CalculateShadowVisibility): the fragment program's surface interpolation at that fragment
__global__ void __launch_bounds__(128) k_site_coords(RasterPass P, LightPass L, size_t nSites, const unsigned long long* keys, float4* sc4)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nSites) return;
    unsigned long long key = keys[i];
    int                prim = (int)(key >> 32), px = (int)((key >> 16) & 0xffff), py = (int)(key & 0xffff);
    SetupRegs          s = load_setup(P.setup, prim);
    TriCover           tc = make_cover(s.X[0], s.Y[0], s.X[1], s.Y[1], s.X[2], s.Y[2]);
    V3                 bary;
    cover_test(tc, px, py, bary);
    const DrawCmdD&    d = find_draw(P.draws, P.nDraws, prim);
    float              vy[48];
    load_vary(P.vary, prim, vy);
    Surface sf = interpolate_surface(P, d, vy, bary, d.kind == FGL_SHADER_PBR ? d.mat.pbr_normal_map : d.mat.normal_map);
    V3      lightDir = vnormalize(vsub(v3(d.lightPos[0], d.lightPos[1], d.lightPos[2]), sf.posWS));
    V3      sc = vadd(vscale(sf.lightNDC, 0.5f), v3(0.5f, 0.5f, 0.5f));
    float   bias = fmaxf(L.biasSlope * (1.f - vdot(sf.normal, lightDir)), L.biasMin);
    sc4[i] = make_float4(sc.x, sc.y, sc.z, bias);
}
}  // namespace

// -------------------------------------------------------------------------------------------------------------
// Forward mode with PCF / PCSS: enumerates the fragments that consumed samples, in consumption order (the "sites").
// Outputs (owned by the context): siteOfPixel (site of each pixel's final winner, -1 = no fragment) and the sites'
// shadow coordinates.  Re-runs the raster loops in counting / recording mode over all triangles of the pass.
int fgl_run_forward_sites(fgl_ctx* c, RasterPass& P, const LightPass& L, size_t* nSitesOut, const float4** sc4Out)
{
    if (int rc = fgl_not_while_recording(c, "a forward pass with a stochastic shadow filter (fragment counts are read back)")) return rc;
    cudaStream_t st = c->stream;
    size_t       nPix = (size_t)P.W * P.H;
    int          nPrims = P.nPrims;
    if (int rc = fgl_reserve(c, c->fragCount, (nPix + 1) * 4)) return rc;
    if (int rc = fgl_reserve(c, c->fragOffset, (nPix + 1) * 4)) return rc;
    if (int rc = fgl_reserve(c, c->nPass, (nPix + 1) * 4)) return rc;
    if (int rc = fgl_reserve(c, c->passOff, (nPix + 1) * 4)) return rc;
    if (int rc = fgl_reserve(c, c->siteOfPixel, nPix * 4)) return rc;
    P.fragCount = (unsigned*)c->fragCount.p, P.fragOffset = (unsigned*)c->fragOffset.p, P.siteOfPixel = (int*)c->siteOfPixel.p;
    // block work items over ALL triangles of the pass (the depth raster may have been flushed in pieces)
    if (int rc = fgl_reserve(c, c->blkScan, ((size_t)nPrims + 1) * 4)) return rc;
    P.blkScan = (int*)c->blkScan.p;
    const int nTiles = (nPrims + kSetupTile - 1) / kSetupTile;
    if (int rc = fgl_reserve(c, c->tileState, ((size_t)nTiles + 1) * 4)) return rc;
    int* tileBase = (int*)c->tileState.p;
    k_tile_local_scan<<<nTiles, kSetupTile, 0, st>>>(P.nblk, P.blkScan, tileBase, nPrims);
    k_tile_scan<<<1, 1024, 0, st>>>(tileBase, nTiles);
    size_t tmpBytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, (unsigned*)nullptr, (unsigned*)nullptr, (int)nPix + 1, st);
    if (int rc = fgl_reserve(c, c->scanTmp, tmpBytes)) return rc;
    auto raster = [&](int mode) {
        LaunchScope ls(c, mode == RM_COUNT ? "forward_frag_count" : "forward_frag_fill", 0);
        if (mode == RM_COUNT)
        {
            k_raster_small<RM_COUNT><<<(nPrims + 127) / 128, 128, 0, st>>>(P, 0);
            k_raster_blocks<RM_COUNT, 32><<<c->numSMs * 8, 256, 0, st>>>(P, 0, nPrims, tileBase);
        }
        else
        {
            k_raster_small<RM_FILL><<<(nPrims + 127) / 128, 128, 0, st>>>(P, 0);
            k_raster_blocks<RM_FILL, 32><<<c->numSMs * 8, 256, 0, st>>>(P, 0, nPrims, tileBase);
        }
        ++c->launches;
    };
    FGL_CUDA(c, cudaMemsetAsync(P.fragCount, 0, (nPix + 1) * 4, st));
    raster(RM_COUNT);
    cub::DeviceScan::ExclusiveSum(c->scanTmp.p, tmpBytes, P.fragCount, P.fragOffset, (int)nPix + 1, st);
    unsigned nFrags = 0;
    FGL_CUDA(c, cudaMemcpyAsync(&nFrags, P.fragOffset + nPix, 4, cudaMemcpyDeviceToHost, st));
    FGL_CUDA(c, cudaStreamSynchronize(st));
    if (int rc = fgl_reserve(c, c->frags, ((size_t)nFrags + 1) * 8)) return rc;
    P.frags = (unsigned long long*)c->frags.p;
    FGL_CUDA(c, cudaMemsetAsync(P.fragCount, 0, (nPix + 1) * 4, st));
    raster(RM_FILL);
    unsigned nb = (unsigned)((nPix + 127) / 128);
    {
        LaunchScope ls(c, "forward_frag_sort", (uint64_t)nFrags * 16);
        k_frag_passing<false><<<nb, 128, 0, st>>>(P, (unsigned*)c->nPass.p, nullptr, nullptr, nullptr);
    }
    FGL_CUDA(c, cudaMemsetAsync((unsigned*)c->nPass.p + nPix, 0, 4, st));
    cub::DeviceScan::ExclusiveSum(c->scanTmp.p, tmpBytes, (unsigned*)c->nPass.p, (unsigned*)c->passOff.p, (int)nPix + 1, st);
    unsigned nSites = 0;
    FGL_CUDA(c, cudaMemcpyAsync(&nSites, (unsigned*)c->passOff.p + nPix, 4, cudaMemcpyDeviceToHost, st));
    FGL_CUDA(c, cudaStreamSynchronize(st));
    if (int rc = fgl_reserve(c, c->siteKeys, ((size_t)nSites + 1) * 8 * 2)) return rc;
    if (int rc = fgl_reserve(c, c->siteVals, ((size_t)nSites + 1) * 4 * 2)) return rc;
    if (int rc = fgl_reserve(c, c->siteSc4, ((size_t)nSites + 1) * 16)) return rc;
    unsigned long long *keysIn = (unsigned long long*)c->siteKeys.p, *keysOut = keysIn + nSites + 1;
    unsigned *          valsIn = (unsigned*)c->siteVals.p, *valsOut = valsIn + nSites + 1;
    {
        LaunchScope ls(c, "forward_frag_emit", (uint64_t)nFrags * 8 + (uint64_t)nSites * 12);
        k_frag_passing<true><<<nb, 128, 0, st>>>(P, (unsigned*)c->nPass.p, (const unsigned*)c->passOff.p, keysIn, valsIn);
    }
    FGL_CUDA(c, cudaMemsetAsync(c->siteOfPixel.p, 0xFF, nPix * 4, st));
    if (nSites)
    {
        size_t sortBytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, sortBytes, keysIn, keysOut, valsIn, valsOut, (int)nSites, 0, 64, st);
        if (int rc = fgl_reserve(c, c->sortTmp, sortBytes)) return rc;
        {
            LaunchScope ls(c, "forward_site_sort", (uint64_t)nSites * 24);
            cub::DeviceRadixSort::SortPairs(c->sortTmp.p, sortBytes, keysIn, keysOut, valsIn, valsOut, (int)nSites, 0, 64, st);
        }
        {
            LaunchScope ls(c, "forward_site_coords", (uint64_t)nSites * (8 + 240 + 16));
            k_site_of_pixel<<<(nSites + 255) / 256, 256, 0, st>>>(nSites, valsOut, (int*)c->siteOfPixel.p);
            k_site_coords<<<(nSites + 127) / 128, 128, 0, st>>>(P, L, nSites, keysOut, (float4*)c->siteSc4.p);
            ++c->launches;
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fgl_fail(c, FGL_ERR_CUDA, std::string("forward sites: ") + cudaGetErrorString(e));
    *nSitesOut = nSites, *sc4Out = (const float4*)c->siteSc4.p;
    return FGL_OK;
}

int fgl_run_resolve_forward(fgl_ctx* c, const RasterPass& P, PlanesD planes, const LightPass& L)
{
    if (P.row1 <= P.row0) return FGL_OK;  // an empty band (more GPUs than rows)
    dim3        grid((P.W + 127) / 128, P.row1 - P.row0);
    LaunchScope ls(c, "resolve_forward", (uint64_t)P.W * (P.row1 - P.row0) * 24);
    k_resolve_forward<<<grid, 128, 0, c->stream>>>(P, planes, L);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fgl_fail(c, FGL_ERR_CUDA, std::string("resolve_forward: ") + cudaGetErrorString(e));
    return FGL_OK;
}

int fgl_run_raster(fgl_ctx* c, const RasterPass& P, PlanesD planes, uint8_t* rgb8, const LightPass* forwardLight)
{
    (void)rgb8;
    cudaStream_t st = c->stream;
    int          primBegin = c->flushedPrims;
    int          nNew = P.nPrims - primBegin;
    size_t       nPix = (size_t)P.W * P.H;
    int*         tileBase = nullptr;
    // Thread-per-triangle only pays off when there are enough small triangles to fill the machine; otherwise every
    // triangle takes the warp-per-block path (a small triangle is a single block).
    const int smallArea = nNew >= 65536 ? kSmallArea : 0;
    if (nNew > 0)
    {
        const int nTiles = (nNew + kSetupTile - 1) / kSetupTile;
        if (int rc = fgl_reserve(c, c->tileState, ((size_t)nTiles + 1) * 4)) return rc;
        tileBase = (int*)c->tileState.p;
        {
            LaunchScope ls(c, "setup", (uint64_t)nNew * 320);
            k_setup<<<nTiles, kSetupTile, 0, st>>>(P, primBegin, smallArea, tileBase);
        }
        {
            LaunchScope ls(c, "scan", (uint64_t)nTiles * 8);
            k_tile_scan<<<1, 1024, 0, st>>>(tileBase, nTiles);
        }
        if (smallArea > 0)
        {
            LaunchScope ls(c, "raster_small", 0);
            k_raster_small<RM_DEPTH><<<(nNew + 127) / 128, 128, 0, st>>>(P, primBegin);
        }
        {
            LaunchScope ls(c, "raster_blocks", 0);
            // passes with many triangles: mostly small blocks, eight lanes each; few triangles: the big ones dominate
            if (smallArea > 0) k_raster_blocks<RM_DEPTH, 8><<<c->numSMs * 8, 256, 0, st>>>(P, primBegin, nNew, tileBase);
            else k_raster_blocks<RM_DEPTH, 32><<<c->numSMs * 8, 256, 0, st>>>(P, primBegin, nNew, tileBase);
        }
    }
    // sort-first group: the resolves store into the other contexts' planes — not before those have begun this frame
    if (P.group)
        if (int rc = fgl_group_wait(c, FGL_GROUP_READY)) return rc;
    if (P.passType == FGL_PASS_SHADOW)
    {
        const size_t nBand = (size_t)P.W * (P.row1 - P.row0);
        LaunchScope ls(c, "resolve_shadow", (uint64_t)nBand * (12 + 4 * (P.group ? P.peers.n : 1)));
        if (nBand) k_resolve_shadow<<<(unsigned)((nBand + 255) / 256), 256, 0, st>>>(P, planes.p[FGL_PLANE_SHADOW], planes.p[FGL_PLANE_DEPTH]);
    }
    else
    {
        dim3 grid((P.W + 127) / 128, P.row1 - P.row0);
        if (P.passType == FGL_PASS_GEOMETRY)
        {
            const int rows = P.group ? P.row1 - P.row0 : P.H;  // stand-alone bands also resolve the depth of all other rows (SSAO)
            LaunchScope ls(c, "resolve_geometry", (uint64_t)P.W * (P.row1 - P.row0) * 88 + (uint64_t)P.W * (rows - (P.row1 - P.row0)) * 12 +
                                                      (P.group ? (uint64_t)P.W * (P.own1 - P.own0) * 4 * (P.peers.n - 1) : 0));
            if (rows > 0) k_resolve_geometry<<<dim3((P.W + 127) / 128, rows), 128, 0, st>>>(P, planes);
        }
        else if (P.passType == FGL_PASS_FORWARD && forwardLight && P.row1 > P.row0)
        {
            LaunchScope ls(c, "resolve_forward", (uint64_t)P.W * (P.row1 - P.row0) * 24);
            k_resolve_forward<<<grid, 128, 0, st>>>(P, planes, *forwardLight);
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fgl_fail(c, FGL_ERR_CUDA, std::string("raster launch: ") + cudaGetErrorString(e));
    return FGL_OK;
}
