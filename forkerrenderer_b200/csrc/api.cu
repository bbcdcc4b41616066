// api.cu — the C ABI of include/forkergl_b200.h over the CUDA kernels: context, resources, ForkerGL state, pass
// flushing, plane I/O and instrumentation.  This library has no CPU path: every entry either enqueues device work
// on the context's stream or fails with an error code.
#include <algorithm>
#include <cfloat>
#include <cstring>
#include <map>

#include "fgl_internal.h"
#include "stream.h"

static std::string g_createError;

int fgl_fail(fgl_ctx* c, int code, const std::string& msg)
{
    if (c) c->error = msg;
    else g_createError = msg;
    return code;
}

int fgl_not_while_recording(fgl_ctx* c, const char* what)
{
    if (!c || !c->recording) return FGL_OK;
    return fgl_fail(c, FGL_ERR_UNSUPPORTED, std::string(what) + " cannot be part of a recorded frame (fgl_frame_record_begin .. fgl_frame_record_end)");
}

int fgl_reserve(fgl_ctx* c, DevBuf& b, size_t bytes)
{
    if (bytes <= b.cap) return FGL_OK;
    if (c && c->recording)
        return fgl_fail(c, FGL_ERR_UNSUPPORTED, "a device buffer would have to grow while a frame is being recorded: render the same frame once without recording first");
    size_t want = std::max(bytes, b.cap + b.cap / 2);
    void*  p = nullptr;
    if (cudaMalloc(&p, want) != cudaSuccess)
    {
        cudaGetLastError();
        if (cudaMalloc(&p, bytes) != cudaSuccess) return fgl_fail(c, FGL_ERR_NOMEM, "cudaMalloc of " + std::to_string(bytes) + " bytes failed");
        want = bytes;
    }
    if (b.p && c)
    {   // recorded frames have the old address baked into their graphs
        for (auto& f : c->recorded)
            if (f.exec) cudaGraphExecDestroy(f.exec), f.exec = nullptr;
    }
    if (b.p)
    {   // growth keeps the contents (a pass can be flushed more than once)
        cudaMemcpyAsync(p, b.p, b.cap, cudaMemcpyDeviceToDevice, c->stream);
        cudaStreamSynchronize(c->stream);
        cudaFree(b.p);
    }
    b.p = p, b.cap = want;
    return FGL_OK;
}

static void release(DevBuf& b)
{
    if (b.p) cudaFree(b.p);
    b.p = nullptr, b.cap = 0;
}

// Per-launch timing records are bounded (kMaxTimingRecs; later launches are not recorded until fgl_reset_timings) and
// their events are recycled through c->eventPool, so a long run with timing on does not accumulate CUDA events.
static constexpr size_t kMaxTimingRecs = 1 << 16;
static cudaEvent_t take_event(fgl_ctx* c)
{
    cudaEvent_t e = nullptr;
    if (!c->eventPool.empty()) e = c->eventPool.back(), c->eventPool.pop_back();
    else cudaEventCreate(&e);
    return e;
}
bool fgl_time_begin(fgl_ctx* c, const char* name, uint64_t bytes, const unsigned* lateCount, uint64_t lateBytesEach)
{
    if (c->timings.size() >= kMaxTimingRecs) return false;
    TimingRec r;
    r.name = name, r.bytes = bytes, r.lateCount = lateCount, r.lateBytesEach = lateBytesEach;
    r.e0 = take_event(c), r.e1 = take_event(c);
    cudaEventRecord(r.e0, c->stream);
    c->timings.push_back(r);
    return true;
}
void fgl_time_end(fgl_ctx* c) { cudaEventRecord(c->timings.back().e1, c->stream); }

namespace
{
constexpr size_t kRecStagingBytes = (size_t)4 << 20;  // page-locked staging of one recorded frame's draw commands

#define ENTER(c)                                                      \
    if (!(c)) return FGL_ERR_INVALID;                                 \
    if (cudaSetDevice((c)->device) != cudaSuccess) return fgl_fail((c), FGL_ERR_CUDA, "cudaSetDevice failed")

void identity(float* m)
{
    memset(m, 0, 16 * sizeof(float));
    m[0] = m[5] = m[10] = m[15] = 1.f;
}

bool is1ch(int plane) { return plane == FGL_PLANE_DEPTH || plane == FGL_PLANE_SHADOW || plane == FGL_PLANE_SHADINGTYPE || plane == FGL_PLANE_AO; }

int plane_init(fgl_ctx* c, int plane, int w, int h, float value)
{
    if (w <= 0 || h <= 0 || w > 65535 || h > 65535 || (size_t)w * h > 0x7fffffffull) return fgl_fail(c, FGL_ERR_INVALID, "buffer size out of range");
    PlaneH& p = c->planes[plane];
    p.w = w, p.h = h, p.ch = is1ch(plane) ? 1 : 3;
    if (int rc = fgl_reserve(c, p.buf, (size_t)w * h * p.ch * 4)) return rc;
    p.fillPending = true, p.fillValue = value, p.fillIsRGB = false, p.fillPartial = false;
    return FGL_OK;
}

// executes a deferred clear (Buffer constructors of the reference, buffer.cpp:8-18,101-111)
int materialize(fgl_ctx* c, int plane)
{
    PlaneH& p = c->planes[plane];
    if (!p.fillPending || !p.buf.p) return FGL_OK;
    p.fillPending = false;
    size_t n = (size_t)p.w * p.h;
    if (p.fillPartial)
    {   // a band's passes wrote rows [validRow0, validRow1): clear the rest, channel by channel (SoA)
        p.fillPartial = false;
        for (int ch = 0; ch < p.ch; ++ch)
        {
            float* base = (float*)p.buf.p + (size_t)ch * n;
            float  v = p.fillIsRGB ? p.fillRGB[ch] : p.fillValue;
            if (int rc = fgl_run_fill(c, base, (size_t)p.validRow0 * p.w, v)) return rc;
            if (int rc = fgl_run_fill(c, base + (size_t)p.validRow1 * p.w, (size_t)(p.h - p.validRow1) * p.w, v)) return rc;
        }
        return FGL_OK;
    }
    if (p.fillIsRGB) return fgl_run_fill_rgb(c, (float*)p.buf.p, n, p.fillRGB);
    return fgl_run_fill(c, (float*)p.buf.p, n * p.ch, p.fillValue);
}

// The kernels of a band only touch rows [r0, r1): nothing to clear if the band's own passes wrote them.
int materialize_rows(fgl_ctx* c, int plane, int r0, int r1)
{
    PlaneH& p = c->planes[plane];
    if (p.fillPending && p.fillPartial && r0 >= p.validRow0 && r1 <= p.validRow1) return FGL_OK;
    return materialize(c, plane);
}

void band_of(fgl_ctx* c, int H, int& r0, int& r1)
{
    if (c->group.on) return fgl_group_band(c, H, r0, r1);
    r0 = std::max(0, std::min(c->row0, H));
    r1 = c->row1 < 0 ? H : std::max(r0, std::min(c->row1, H));
}
// Rows whose G-buffer / AO a band needs beyond itself: the V blur is warmed up over kBlurWarm rows above the band and
// looks 4 rows ahead; the H blur and SSAO must therefore cover [r0 - kBlurWarm, r1 + 4).
constexpr int kBlurWarm = 64, kBlurAhead = 4;
void halo_band_of(fgl_ctx* c, int H, int& r0, int& r1)
{
    band_of(c, H, r0, r1);
    r0 = std::max(0, r0 - kBlurWarm), r1 = std::min(H, r1 + kBlurAhead);
}

PlanesD planes_dev(fgl_ctx* c)
{
    PlanesD d;
    for (int i = 0; i <= FGL_PLANE_AO; ++i) d.p[i] = (float*)c->planes[i].buf.p;
    return d;
}

int upload_tex_table(fgl_ctx* c)
{
    if (!c->texTableDirty) return FGL_OK;
    if (int rc = fgl_not_while_recording(c, "the upload of a changed texture table")) return rc;
    std::vector<TexD> t(std::max<size_t>(1, c->textures.size()));
    for (size_t i = 0; i < c->textures.size(); ++i) t[i] = c->textures[i].desc;
    if (int rc = fgl_reserve(c, c->texTable, t.size() * sizeof(TexD))) return rc;
    FGL_CUDA(c, cudaMemcpyAsync(c->texTable.p, t.data(), t.size() * sizeof(TexD), cudaMemcpyHostToDevice, c->stream));
    c->h2dBytes += t.size() * sizeof(TexD);
    FGL_CUDA(c, cudaStreamSynchronize(c->stream));
    c->texTableDirty = false;
    return FGL_OK;
}

ShadowMapD shadow_map_dev(fgl_ctx* c)
{
    ShadowMapD   s;
    const PlaneH& p = c->planes[FGL_PLANE_SHADOW];
    s.d = (const float*)p.buf.p, s.w = p.w, s.h = p.h;
    s.iw = (int)((float)p.w - 0.001f), s.ih = (int)((float)p.h - 0.001f);  // shadow.cpp:27-28
    return s;
}

void fill_light_consts(fgl_ctx* c, LightPass& L)
{
    L.shadowOn = c->shadowOn, L.shadowMode = c->params.shadow_mode;
    L.biasSlope = c->params.shadow_bias_slope, L.biasMin = c->params.shadow_bias_min;
    L.shadowIntensity = c->params.shadow_intensity, L.areaLight = c->params.area_light_size;
    L.pcfFilter = c->params.pcf_filter_size, L.pcssFilter = c->params.pcss_blocker_filter_size;
    L.disk = nullptr, L.chunkOf = nullptr, L.vis = nullptr, L.useAO = 1;
    L.writeF32 = c->params.materialize_frame_f32;
}

// Rasterises the triangles submitted since the last flush and resolves the pass into its planes.
int flush(fgl_ctx* c)
{
    if (c->primCounter == c->flushedPrims) return FGL_OK;
    bool          shadowPass = c->pass == FGL_PASS_SHADOW;
    const PlaneH& target = c->planes[shadowPass ? FGL_PLANE_SHADOW : FGL_PLANE_DEPTH];
    const PlaneH& depth = c->planes[FGL_PLANE_DEPTH];
    if (!target.buf.p || !depth.buf.p || depth.w != target.w || depth.h != target.h)
    {
        c->flushedPrims = c->primCounter;
        return fgl_fail(c, FGL_ERR_STATE, "draw without matching Init*Buffer calls (the reference would index out of bounds)");
    }
    RasterPass P;
    memset(&P, 0, sizeof P);
    P.W = target.w, P.H = target.h;
    P.row0 = 0, P.row1 = P.H;
    if (!shadowPass)
    {
        if (c->pass == FGL_PASS_GEOMETRY) halo_band_of(c, P.H, P.row0, P.row1);
        else band_of(c, P.H, P.row0, P.row1);
    }
    P.passType = c->pass, P.shadowOn = c->shadowOn;
    memcpy(P.viewport, c->viewport, sizeof P.viewport);
    P.cull0 = 0, P.cull1 = P.H, P.own0 = 0, P.own1 = 0, P.group = 0;
    GroupState& grp = c->group;
    if (grp.on)
    {
        if (c->pass == FGL_PASS_FORWARD)
        {
            c->flushedPrims = c->primCounter;
            return fgl_fail(c, FGL_ERR_UNSUPPORTED, "a sort-first group renders deferred frames only");
        }
        if (P.W != grp.W || P.H != grp.H || c->planes[FGL_PLANE_SHADOW].buf.p != grp.expShadow || c->planes[FGL_PLANE_DEPTH].buf.p != grp.expDepth)
        {
            c->flushedPrims = c->primCounter;
            return fgl_fail(c, FGL_ERR_STATE, "the frame does not have the size the group was connected for (fgl_group_export)");
        }
        if (c->primCounter != c->flushedPrims && c->flushedPrims != 0)
        {
            c->flushedPrims = c->primCounter;
            return fgl_fail(c, FGL_ERR_UNSUPPORTED, "a sort-first group needs each raster pass submitted before it is flushed (one flush per pass)");
        }
        fgl_group_band(c, P.H, P.own0, P.own1);
        if (shadowPass) P.row0 = P.own0, P.row1 = P.own1;  // this context's rows of the shadow map
        P.cull0 = P.row0, P.cull1 = P.row1, P.group = 1;
        P.peers.n = grp.world;
        for (int r = 0; r < grp.world; ++r) P.peers.p[r] = shadowPass ? grp.peerShadow[r] : grp.peerDepth[r];
    }
    size_t  nPix = (size_t)P.W * P.H;
    DevBuf& vis = shadowPass ? c->visLight : c->visCamera;
    bool&   visClear = shadowPass ? c->visLightClear : c->visCamClear;
    int&    vw = shadowPass ? c->visLightW : c->visCamW;
    int&    vh = shadowPass ? c->visLightH : c->visCamH;
    // InitDepthBuffer (forkergl.cpp:60-63) re-creates the depth buffer the NEXT raster pass tests against
    if (c->depthInitPending) visClear = true, c->depthInitPending = c->depthInitBound = false;
    if (vw != P.W || vh != P.H) visClear = true;
    if (c->passRestarted && !visClear)
    {   // SetPassType restarted the primitive numbering, but the visibility buffer still holds (depth | id) keys of the pass
        // before it: their ids would be decoded against this pass's triangles.  The reference keeps testing against the old
        // depths here (no InitDepthBuffer in between); that sequence is not supported — say so instead of drawing garbage.
        c->flushedPrims = c->primCounter;
        return fgl_fail(c, FGL_ERR_UNSUPPORTED, "a raster pass was restarted (SetPassType) without InitDepthBuffer: depth carried over between passes is not supported");
    }
    c->passRestarted = false;
    if (int rc = fgl_reserve(c, vis, nPix * 8)) return rc;
    vw = P.W, vh = P.H;
    if (visClear && grp.on)
    {   // only the rows this context rasterises (the others are never read: the resolves stay inside [cull0, cull1))
        size_t rows = (size_t)(P.cull1 - P.cull0);
        LaunchScope ls(c, "vis_clear", rows * P.W * 8);
        if (rows) FGL_CUDA(c, cudaMemsetAsync((char*)vis.p + (size_t)P.cull0 * P.W * 8, 0xFF, rows * P.W * 8, c->stream));
        visClear = false;  // (rows outside the band keep old keys; fgl_group_disconnect schedules a full clear for stand-alone passes)
    }
    else if (visClear)
    {
        LaunchScope ls(c, "vis_clear", nPix * 8);
        FGL_CUDA(c, cudaMemsetAsync(vis.p, 0xFF, nPix * 8, c->stream));
        visClear = false;
    }
    int nPrims = c->primCounter, nNew = nPrims - c->flushedPrims;
    if (int rc = fgl_reserve(c, c->drawsDev, c->draws.size() * sizeof(DrawCmdD))) return rc;
    {
        const size_t drawBytes = c->draws.size() * sizeof(DrawCmdD);
        const void*  src = c->draws.data();
        if (c->recording)
        {   // a replay re-reads the host side of this copy: it has to live as long as the recorded frame does
            const size_t slot = (drawBytes + 255) & ~(size_t)255;
            if (c->recStagingUsed + slot > kRecStagingBytes)
            {
                c->flushedPrims = c->primCounter;
                return fgl_fail(c, FGL_ERR_UNSUPPORTED, "too many draw commands for a recorded frame");
            }
            memcpy((char*)c->recStaging + c->recStagingUsed, src, drawBytes);
            src = (char*)c->recStaging + c->recStagingUsed;
            c->recStagingUsed += slot;
        }
        FGL_CUDA(c, cudaMemcpyAsync(c->drawsDev.p, src, drawBytes, cudaMemcpyHostToDevice, c->stream));
        c->h2dBytes += drawBytes;
    }
    if (int rc = fgl_reserve(c, c->setup, (size_t)nPrims * sizeof(TriSetup))) return rc;
    if (shadowPass) { if (int rc = fgl_reserve(c, c->zndc, (size_t)nPrims * sizeof(float4))) return rc; }
    else if (int rc = fgl_reserve(c, c->vary, (size_t)nPrims * sizeof(TriVary))) return rc;
    if (int rc = fgl_reserve(c, c->nblk, ((size_t)nPrims + 1) * 4)) return rc;
    if (int rc = fgl_reserve(c, c->blkScan, ((size_t)nNew + 1) * 4)) return rc;
    FGL_CUDA(c, cudaMemsetAsync((int*)c->nblk.p + nPrims, 0, 4, c->stream));
    if (int rc = upload_tex_table(c)) return rc;
    P.draws = (const DrawCmdD*)c->drawsDev.p, P.nDraws = (int)c->draws.size(), P.nPrims = nPrims;
    P.setup = (TriSetup*)c->setup.p, P.vary = (TriVary*)c->vary.p, P.zndc = (float4*)c->zndc.p;
    P.nblk = (int*)c->nblk.p, P.blkScan = (int*)c->blkScan.p;
    P.vis = (unsigned long long*)vis.p;
    P.textures = (const TexD*)c->texTable.p;

    bool      fullBand = P.row0 == 0 && P.row1 == P.H;
    LightPass L;
    memset(&L, 0, sizeof L);
    if (shadowPass)
    {
        c->planes[FGL_PLANE_SHADOW].fillPending = false;
        c->planes[FGL_PLANE_DEPTH].fillPending = false;
    }
    else if (c->pass == FGL_PASS_GEOMETRY)
    {
        static const int written[] = { FGL_PLANE_DEPTH, FGL_PLANE_NORMAL, FGL_PLANE_WORLDPOS, FGL_PLANE_LIGHTNDC, FGL_PLANE_ALBEDO,
                                       FGL_PLANE_EMISSIVE, FGL_PLANE_PARAM, FGL_PLANE_SHADINGTYPE, FGL_PLANE_AO };
        for (int pl : written)
        {
            if (pl == FGL_PLANE_LIGHTNDC && !c->shadowOn) continue;
            const PlaneH& q = c->planes[pl];
            if (!q.buf.p || q.w != P.W || q.h != P.H)
            {
                c->flushedPrims = c->primCounter;
                return fgl_fail(c, FGL_ERR_STATE, "geometry pass without InitGeometryBuffers of the depth buffer's size");
            }
            PlaneH& w = c->planes[pl];
            if (fullBand || pl == FGL_PLANE_DEPTH) w.fillPending = false, w.fillPartial = false;  // the resolve writes every depth texel, band or not
            else if (w.fillPending && !w.fillPartial) w.fillPartial = true, w.validRow0 = P.row0, w.validRow1 = P.row1;  // clear of the other rows deferred
            else if (w.fillPending && (w.validRow0 != P.row0 || w.validRow1 != P.row1))
            {
                if (int rc = materialize(c, pl)) return rc;
            }
        }
    }
    else if (c->pass == FGL_PASS_FORWARD)
    {
        const PlaneH& f = c->planes[FGL_PLANE_FRAME];
        if (!f.buf.p || f.w != P.W || f.h != P.H)
        {
            c->flushedPrims = c->primCounter;
            return fgl_fail(c, FGL_ERR_STATE, "forward pass without InitFrameBuffer of the depth buffer's size");
        }
        if (int rc = materialize(c, FGL_PLANE_FRAME)) return rc;
        if (fullBand) c->planes[FGL_PLANE_DEPTH].fillPending = false;
        else if (int rc = materialize(c, FGL_PLANE_DEPTH)) return rc;
        fill_light_consts(c, L);
        if (c->shadowOn)
        {
            if (!c->planes[FGL_PLANE_SHADOW].buf.p)
            {
                c->flushedPrims = c->primCounter;
                return fgl_fail(c, FGL_ERR_STATE, "shadows are on but no shadow pass has run");
            }
            if (int rc = materialize(c, FGL_PLANE_SHADOW)) return rc;
            L.sm = shadow_map_dev(c);
        }
    }
    else
    {   // LightingPass: the reference's DrawTriangle would skip the depth test and run no program; nothing to draw
        c->flushedPrims = c->primCounter;
        return FGL_OK;
    }
    bool stochasticForward = c->pass == FGL_PASS_FORWARD && c->shadowOn && c->params.shadow_mode != FGL_SHADOW_HARD;
    int  rc = fgl_run_raster(c, P, planes_dev(c), nullptr, stochasticForward ? nullptr : &L);
    c->flushedPrims = c->primCounter;
    c->frameRgb8Valid = c->bandRgb8Valid = false;
    if (!rc && grp.on) rc = fgl_group_signal(c, shadowPass ? FGL_GROUP_SHADOW : FGL_GROUP_DEPTH);
    if (rc || !stochasticForward) return rc;
    // Forward + PCF / PCSS: every fragment that passed the depth test when it was submitted consumed samples, so the
    // winners' stream positions depend on all of them.  (Exact for a pass flushed once, which is how Render::DoForwardPass
    // submits it; a pass flushed in pieces restarts the sample stream at each flush.)
    size_t        nSites = 0;
    const float4* sc4 = nullptr;
    if ((rc = fgl_run_forward_sites(c, P, L, &nSites, &sc4))) return rc;
    if (nSites)
        if ((rc = fgl_stream_site_visibility(c, L, nSites, sc4, 0, nSites, 0))) return rc;
    return fgl_run_resolve_forward(c, P, planes_dev(c), L);
}

// A raster pass that ends without a single triangle still consumed the InitDepthBuffer issued for it (the reference's
// depth buffer was re-created): its winner-id plane is empty, and the next pass must not inherit the pending re-init.
void retire_empty_pass(fgl_ctx* c)
{
    if (!c->depthInitBound) return;
    c->depthInitBound = false;
    if (c->primCounter != 0 || !c->depthInitPending) return;
    if (c->pass == FGL_PASS_LIGHTING) return;
    (c->pass == FGL_PASS_SHADOW ? c->visLightClear : c->visCamClear) = true;
    c->depthInitPending = false;
}

void save_host_state(const fgl_ctx* c, RecordedFrame::HostState& h)
{
    static_assert(FGL_PLANE_AO + 1 <= 16, "HostState::planes");
    for (int i = 0; i <= FGL_PLANE_AO; ++i)
    {
        const PlaneH& p = c->planes[i];
        auto&         q = h.planes[i];
        q.w = p.w, q.h = p.h, q.ch = p.ch, q.fillPending = p.fillPending, q.fillIsRGB = p.fillIsRGB, q.fillPartial = p.fillPartial, q.fillValue = p.fillValue;
        memcpy(q.fillRGB, p.fillRGB, sizeof q.fillRGB);
        q.validRow0 = p.validRow0, q.validRow1 = p.validRow1;
    }
    h.frameRgb8Valid = c->frameRgb8Valid, h.bandRgb8Valid = c->bandRgb8Valid, h.visCamClear = c->visCamClear, h.visLightClear = c->visLightClear;
    h.depthInitPending = c->depthInitPending, h.depthInitBound = c->depthInitBound, h.passRestarted = c->passRestarted;
    h.ssaaW = c->ssaaW, h.ssaaH = c->ssaaH, h.visCamW = c->visCamW, h.visCamH = c->visCamH, h.visLightW = c->visLightW, h.visLightH = c->visLightH;
    h.pass = c->pass, h.primCounter = c->primCounter, h.flushedPrims = c->flushedPrims;
}
void restore_host_state(fgl_ctx* c, const RecordedFrame::HostState& h)
{
    for (int i = 0; i <= FGL_PLANE_AO; ++i)
    {
        PlaneH&     p = c->planes[i];
        const auto& q = h.planes[i];
        p.w = q.w, p.h = q.h, p.ch = q.ch, p.fillPending = q.fillPending, p.fillIsRGB = q.fillIsRGB, p.fillPartial = q.fillPartial, p.fillValue = q.fillValue;
        memcpy(p.fillRGB, q.fillRGB, sizeof p.fillRGB);
        p.validRow0 = q.validRow0, p.validRow1 = q.validRow1;
    }
    c->frameRgb8Valid = h.frameRgb8Valid, c->bandRgb8Valid = h.bandRgb8Valid, c->visCamClear = h.visCamClear, c->visLightClear = h.visLightClear;
    c->depthInitPending = h.depthInitPending, c->depthInitBound = h.depthInitBound, c->passRestarted = h.passRestarted;
    c->ssaaW = h.ssaaW, c->ssaaH = h.ssaaH, c->visCamW = h.visCamW, c->visCamH = h.visCamH, c->visLightW = h.visLightW, c->visLightH = h.visLightH;
    c->pass = h.pass, c->primCounter = h.primCounter, c->flushedPrims = h.flushedPrims;
}

int ensure_rgb8(fgl_ctx* c)
{
    if (c->frameRgb8Valid) return FGL_OK;
    PlaneH& f = c->planes[FGL_PLANE_FRAME];
    if (!f.buf.p) return fgl_fail(c, FGL_ERR_STATE, "no frame buffer");
    if (int rc = materialize(c, FGL_PLANE_FRAME)) return rc;
    size_t n = (size_t)f.w * f.h;
    if (int rc = fgl_reserve(c, c->frameRgb8, n * 3 + 16)) return rc;
    if (int rc = fgl_run_quantize(c, (const float*)f.buf.p, n, (uint8_t*)c->frameRgb8.p)) return rc;
    c->frameRgb8Valid = true;
    return FGL_OK;
}
}  // namespace

// ================================================================================================================
extern "C" {

void fgl_default_params(FglParams* p)
{
    p->shadow_mode = FGL_SHADOW_PCSS;
    p->pcf_filter_size = 0.007;
    p->pcss_blocker_filter_size = 0.005;
    p->area_light_size = 2.5f;
    p->shadow_bias_slope = 0.009f;
    p->shadow_bias_min = 0.007f;
    p->shadow_intensity = 0.6f;
    p->ssao_radius = 0.075f;
    p->ssao_range_check_radius = 0.01f;
    p->ssao_bias = 0.0005f;
    p->ssao_range_check = 1;
    p->materialize_frame_f32 = 1;
}

int fgl_create(int device, fgl_ctx** out)
{
    if (!out) return fgl_fail(nullptr, FGL_ERR_INVALID, "out_ctx is NULL");
    int         n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fgl_fail(nullptr, FGL_ERR_CUDA, std::string("no usable CUDA device (this library has no CPU path): ") + cudaGetErrorString(e));
    if (device < 0 || device >= n) return fgl_fail(nullptr, FGL_ERR_INVALID, "cuda_device out of range");
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fgl_fail(nullptr, FGL_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major < 10)
        return fgl_fail(nullptr, FGL_ERR_UNSUPPORTED, std::string("device ") + prop.name + " is not sm_100a (kernels are built for Blackwell only)");
    fgl_ctx* c = new fgl_ctx();
    c->device = device;
    c->numSMs = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
    if ((e = cudaStreamCreateWithFlags(&c->ownStream, cudaStreamNonBlocking)) != cudaSuccess)
    {
        delete c;
        return fgl_fail(nullptr, FGL_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
    }
    c->stream = c->ownStream;
    fgl_default_params(&c->params);
    identity(c->viewport), identity(c->viewProj), identity(c->lightSpace);
    *out = c;
    return FGL_OK;
}

void fgl_destroy(fgl_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->chainStream) cudaStreamSynchronize(c->chainStream);
    cudaStreamSynchronize(c->stream);
    if (c->chainStream) cudaStreamDestroy(c->chainStream), cudaEventDestroy(c->evChainGo), cudaEventDestroy(c->evChainDone);
    fgl_group_release(c);
    release(c->group.flags);
    fgl_stream_destroy(c);
    for (auto& p : c->planes) release(p.buf);
    release(c->frameRgb8), release(c->ssaaRgb8), release(c->blurTmp), release(c->visCamera), release(c->visLight), release(c->texTable);
    release(c->drawsDev), release(c->setup), release(c->vary), release(c->zndc), release(c->nblk), release(c->blkScan), release(c->scanTmp), release(c->tileState);
    for (auto& b : c->preChunks) release(b);
    for (auto& t : c->textures) release(t.data);
    for (auto& v : c->vertices) release(v.pos), release(v.uv), release(v.nrm), release(v.tan);
    for (auto& m : c->meshes) release(m.pi), release(m.ti), release(m.ni);
    for (auto& t : c->timings) cudaEventDestroy(t.e0), cudaEventDestroy(t.e1);
    for (auto& e : c->eventPool) cudaEventDestroy(e);
    if (c->recording)
    {   // an unfinished recording: close the capture so that the stream can be destroyed
        cudaGraph_t g = nullptr;
        cudaStreamEndCapture(c->stream, &g);
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
    }
    if (c->recStaging) cudaFreeHost(c->recStaging);
    for (auto& f : c->recorded)
    {
        if (f.exec) cudaGraphExecDestroy(f.exec);
        if (f.staging) cudaFreeHost(f.staging);
    }
    cudaStreamDestroy(c->ownStream);
    delete c;
}

const char* fgl_last_error(fgl_ctx* c) { return c ? c->error.c_str() : g_createError.c_str(); }
const char* fgl_backend_name(void) { return "cuda-sm_100a"; }

int fgl_set_stream(fgl_ctx* c, void* s)
{
    ENTER(c);
    if (int rc = fgl_not_while_recording(c, "fgl_set_stream")) return rc;
    if (int rc = flush(c)) return rc;
    FGL_CUDA(c, cudaStreamSynchronize(c->stream));
    c->stream = s ? (cudaStream_t)s : c->ownStream;
    return FGL_OK;
}
int fgl_set_params(fgl_ctx* c, const FglParams* p)
{
    ENTER(c);
    if (!p) return fgl_fail(c, FGL_ERR_INVALID, "params is NULL");
    if (p->shadow_mode < FGL_SHADOW_HARD || p->shadow_mode > FGL_SHADOW_PCSS) return fgl_fail(c, FGL_ERR_INVALID, "bad shadow_mode");
    if (int rc = flush(c)) return rc;
    c->params = *p;
    return FGL_OK;
}
int fgl_sync(fgl_ctx* c)
{
    ENTER(c);
    if (int rc = fgl_not_while_recording(c, "fgl_sync")) return rc;
    if (int rc = flush(c)) return rc;
    if (c->chainStream) FGL_CUDA(c, cudaStreamSynchronize(c->chainStream));
    FGL_CUDA(c, cudaStreamSynchronize(c->stream));
    return FGL_OK;
}

// ---- resources ---------------------------------------------------------------------------------------------------
static int upload(fgl_ctx* c, DevBuf& b, const void* src, size_t bytes)
{
    if (int rc = fgl_reserve(c, b, std::max<size_t>(bytes, 16))) return rc;
    if (bytes) FGL_CUDA(c, cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice));
    c->h2dBytes += bytes;
    return FGL_OK;
}

int fgl_upload_texture(fgl_ctx* c, const uint8_t* texels, int w, int h, int bpp, int wrap, int filter, int* id)
{
    ENTER(c);
    if (int rc = fgl_not_while_recording(c, "fgl_upload_texture")) return rc;
    if (!texels || w <= 0 || h <= 0 || (bpp != 1 && bpp != 3 && bpp != 4) || !id || wrap < 0 || wrap > 3 || filter < 0 || filter > 1)
        return fgl_fail(c, FGL_ERR_INVALID, "fgl_upload_texture: bad arguments");
    TextureH t;
    if (int rc = upload(c, t.data, texels, (size_t)w * h * bpp)) return rc;
    t.desc.data = (const uint8_t*)t.data.p, t.desc.w = w, t.desc.h = h, t.desc.bpp = bpp, t.desc.wrap = wrap, t.desc.filter = filter;
    c->textures.push_back(t);
    c->texTableDirty = true;
    *id = (int)c->textures.size() - 1;
    return FGL_OK;
}

int fgl_upload_vertices(fgl_ctx* c, const float* pos, int np, const float* uv, int nt, const float* nrm, int nn, const float* tan, int ntan, int* id)
{
    ENTER(c);
    if (int rc = fgl_not_while_recording(c, "fgl_upload_vertices")) return rc;
    if (!id || np < 0 || nt < 0 || nn < 0 || ntan < 0) return fgl_fail(c, FGL_ERR_INVALID, "fgl_upload_vertices: bad arguments");
    VerticesH v;
    v.nPos = pos ? np : 0, v.nUv = uv ? nt : 0, v.nNrm = nrm ? nn : 0, v.nTan = tan ? ntan : 0;
    if (int rc = upload(c, v.pos, pos, (size_t)v.nPos * 12)) return rc;
    if (int rc = upload(c, v.uv, uv, (size_t)v.nUv * 8)) return rc;
    if (int rc = upload(c, v.nrm, nrm, (size_t)v.nNrm * 12)) return rc;
    if (int rc = upload(c, v.tan, tan, (size_t)v.nTan * 12)) return rc;
    c->vertices.push_back(v);
    *id = (int)c->vertices.size() - 1;
    return FGL_OK;
}

int fgl_upload_mesh(fgl_ctx* c, int vid, int nFaces, const int* pi, const int* ti, const int* ni, const FglMaterial* mat, int hasTangents,
                    int supportPBR, int* id)
{
    ENTER(c);
    if (int rc = fgl_not_while_recording(c, "fgl_upload_mesh")) return rc;
    if (!id || !mat || vid < 0 || vid >= (int)c->vertices.size() || nFaces < 0 || (nFaces && (!pi || !ti || !ni)))
        return fgl_fail(c, FGL_ERR_INVALID, "fgl_upload_mesh: bad arguments");
    const VerticesH& vb = c->vertices[vid];
    if (hasTangents && vb.nTan == 0) return fgl_fail(c, FGL_ERR_INVALID, "fgl_upload_mesh: has_tangents without tangents");
    for (int i = 0; i < nFaces * 3; ++i)
        if (pi[i] < 0 || pi[i] >= vb.nPos || ti[i] < 0 || ti[i] >= vb.nUv || ni[i] < 0 || ni[i] >= vb.nNrm || (hasTangents && pi[i] >= vb.nTan))
            return fgl_fail(c, FGL_ERR_INVALID, "fgl_upload_mesh: index out of range");
    const int maps[] = { mat->diffuse_map, mat->specular_map, mat->normal_map, mat->emissive_map, mat->base_color_map, mat->roughness_map,
                         mat->metalness_map, mat->ao_map, mat->pbr_normal_map, mat->pbr_emissive_map };
    for (int t : maps)
        if (t >= (int)c->textures.size()) return fgl_fail(c, FGL_ERR_INVALID, "fgl_upload_mesh: texture id out of range");
    MeshH m;
    m.vertices = vid, m.nFaces = nFaces, m.mat = *mat, m.hasTangents = hasTangents ? 1 : 0, m.supportPBR = supportPBR ? 1 : 0;
    if (int rc = upload(c, m.pi, pi, (size_t)nFaces * 12)) return rc;
    if (int rc = upload(c, m.ti, ti, (size_t)nFaces * 12)) return rc;
    if (int rc = upload(c, m.ni, ni, (size_t)nFaces * 12)) return rc;
    c->meshes.push_back(m);
    *id = (int)c->meshes.size() - 1;
    return FGL_OK;
}

// ---- ForkerGL state --------------------------------------------------------------------------------------------------
int fgl_init_frame_buffer(fgl_ctx* c, int w, int h)
{
    ENTER(c);
    if (int rc = flush(c)) return rc;
    c->frameRgb8Valid = c->bandRgb8Valid = false;
    return plane_init(c, FGL_PLANE_FRAME, w, h, 0.f);
}
int fgl_init_depth_buffer(fgl_ctx* c, int w, int h)
{
    ENTER(c);
    if (int rc = flush(c)) return rc;
    retire_empty_pass(c);
    c->depthInitPending = true;
    return plane_init(c, FGL_PLANE_DEPTH, w, h, FLT_MAX);
}
int fgl_init_shadow_buffer(fgl_ctx* c, int w, int h)
{
    ENTER(c);
    if (int rc = flush(c)) return rc;
    return plane_init(c, FGL_PLANE_SHADOW, w, h, 0.f);
}
int fgl_init_geometry_buffers(fgl_ctx* c, int w, int h)
{
    ENTER(c);
    if (int rc = flush(c)) return rc;
    static const int zero[] = { FGL_PLANE_NORMAL, FGL_PLANE_WORLDPOS, FGL_PLANE_ALBEDO, FGL_PLANE_EMISSIVE, FGL_PLANE_PARAM, FGL_PLANE_SHADINGTYPE };
    for (int pl : zero)
        if (int rc = plane_init(c, pl, w, h, 0.f)) return rc;
    if (c->shadowOn)
        if (int rc = plane_init(c, FGL_PLANE_LIGHTNDC, w, h, 0.f)) return rc;
    return plane_init(c, FGL_PLANE_AO, w, h, 1.f);
}
int fgl_clear_color(fgl_ctx* c, const float rgb[3])
{
    ENTER(c);
    if (!rgb) return fgl_fail(c, FGL_ERR_INVALID, "rgb is NULL");
    if (int rc = flush(c)) return rc;
    PlaneH& f = c->planes[FGL_PLANE_FRAME];
    if (!f.buf.p) return fgl_fail(c, FGL_ERR_STATE, "ClearColor before InitFrameBuffer");
    f.fillPending = true, f.fillIsRGB = true;
    memcpy(f.fillRGB, rgb, 12);
    c->frameRgb8Valid = c->bandRgb8Valid = false;
    return FGL_OK;
}
int fgl_set_viewport(fgl_ctx* c, int x, int y, int w, int h)  // forkergl.cpp:89-102
{
    ENTER(c);
    if (int rc = flush(c)) return rc;
    identity(c->viewport);
    c->viewport[0] = w / 2.f;
    c->viewport[5] = h / 2.f;
    c->viewport[3] = x + w / 2.f;
    c->viewport[7] = y + h / 2.f;
    c->viewport[10] = 1 / 2.f;
    c->viewport[11] = 1 / 2.f;
    return FGL_OK;
}
int fgl_get_viewport_matrix(fgl_ctx* c, float o[16]) { ENTER(c); memcpy(o, c->viewport, 64); return FGL_OK; }
int fgl_set_view_projection_matrix(fgl_ctx* c, const float m[16]) { ENTER(c); memcpy(c->viewProj, m, 64); return FGL_OK; }
int fgl_get_view_projection_matrix(fgl_ctx* c, float o[16]) { ENTER(c); memcpy(o, c->viewProj, 64); return FGL_OK; }
int fgl_set_light_space_matrix(fgl_ctx* c, const float m[16]) { ENTER(c); memcpy(c->lightSpace, m, 64); return FGL_OK; }
int fgl_get_light_space_matrix(fgl_ctx* c, float o[16]) { ENTER(c); memcpy(o, c->lightSpace, 64); return FGL_OK; }
int fgl_set_render_mode(fgl_ctx* c, int mode)
{
    ENTER(c);
    if (mode != FGL_MODE_FORWARD && mode != FGL_MODE_DEFERRED) return fgl_fail(c, FGL_ERR_INVALID, "bad render mode");
    c->mode = mode;
    return FGL_OK;
}
int fgl_get_render_mode(fgl_ctx* c, int* mode) { ENTER(c); *mode = c->mode; return FGL_OK; }
int fgl_set_pass_type(fgl_ctx* c, int pass)
{
    ENTER(c);
    if (pass < FGL_PASS_FORWARD || pass > FGL_PASS_SHADOW) return fgl_fail(c, FGL_ERR_INVALID, "bad pass type");
    if (int rc = flush(c)) return rc;
    retire_empty_pass(c);
    if (!c->preChunks.empty())
    {   // arrays of fgl_draw_triangles: consumed by the flush above
        FGL_CUDA(c, cudaStreamSynchronize(c->stream));
        for (size_t i = 0; i + 1 < c->preChunks.size(); ++i) release(c->preChunks[i]);
        DevBuf last = c->preChunks.back();
        c->preChunks.clear();
        c->preChunks.push_back(last);
        c->preUsed = 0;
    }
    c->pass = pass;
    c->primCounter = c->flushedPrims = 0;
    c->draws.clear();
    c->passRestarted = true;
    c->depthInitBound = c->depthInitPending;  // an InitDepthBuffer issued before SetPassType belongs to this pass
    return FGL_OK;
}
int fgl_set_shadow_status(fgl_ctx* c, int on)
{
    ENTER(c);
    if (int rc = flush(c)) return rc;
    c->shadowOn = on ? 1 : 0;
    return FGL_OK;
}
int fgl_begin_frame(fgl_ctx* c)
{
    ENTER(c);
    if (int rc = flush(c)) return rc;
    if (c->chainEventPending)
    {   // a chain that nobody picked up: the new frame's passes must not overtake it
        FGL_CUDA(c, cudaStreamWaitEvent(c->stream, c->evChainDone, 0));
        c->chainEventPending = false;
    }
    c->chainBlockersBefore = 0;  // a band's input is set after fgl_begin_frame, every frame (fgl_set_chain_blockers_before)
    fgl_stream_begin_frame(c);
    if (c->group.on)
    {   // the previous frame has been consumed here: the peers may store the next frame's rows into this context's planes
        GroupState& g = c->group;
        ++g.epoch;
        g.readyWaited = g.shadowWaited = g.depthWaited = g.bandWaited = false;
        // clears of the exchanged planes are never executed in a group: every row is written by the band that owns it
        c->planes[FGL_PLANE_SHADOW].fillPending = c->planes[FGL_PLANE_DEPTH].fillPending = false;
        return fgl_group_signal(c, FGL_GROUP_READY);
    }
    return FGL_OK;
}
int fgl_set_chain_blockers_before(fgl_ctx* c, uint64_t blockers)
{
    ENTER(c);
    c->chainBlockersBefore = blockers;
    return FGL_OK;
}
int fgl_get_chain_blockers(fgl_ctx* c, uint64_t* out)
{
    ENTER(c);
    if (int rc = fgl_not_while_recording(c, "fgl_get_chain_blockers")) return rc;
    if (!out) return fgl_fail(c, FGL_ERR_INVALID, "out is NULL");
    unsigned long long v = 0;
    if (int rc = fgl_stream_chain_total(c, &v)) return rc;
    *out = v;
    return FGL_OK;
}
int fgl_chain_peer_mailbox(fgl_ctx* c, void** device_ptr, void* ipc_handle, size_t ipc_handle_bytes)
{
    ENTER(c);
    if (ipc_handle && ipc_handle_bytes < 64) return fgl_fail(c, FGL_ERR_INVALID, "fgl_chain_peer_mailbox: a CUDA IPC handle needs 64 bytes");
    return fgl_stream_peer_mailbox(c, device_ptr, ipc_handle);
}
int fgl_chain_peer_connect(fgl_ctx* c, void* next_device_ptr, const void* next_ipc_handle, int wait_for_previous, int enable)
{
    ENTER(c);
    if (int rc = flush(c)) return rc;
    return fgl_stream_peer_connect(c, next_device_ptr, next_ipc_handle, wait_for_previous, enable);
}
int fgl_set_row_band(fgl_ctx* c, int row0, int row1)
{
    ENTER(c);
    if (row0 < 0 || (row1 >= 0 && row1 < row0)) return fgl_fail(c, FGL_ERR_INVALID, "bad row band");
    if (c->group.on) return fgl_fail(c, FGL_ERR_STATE, "the row band of a context in a sort-first group follows from its rank (fgl_group_connect)");
    if (int rc = flush(c)) return rc;
    c->row0 = row0, c->row1 = row1;
    return FGL_OK;
}

// ---- sort-first group --------------------------------------------------------------------------------------------------
int fgl_group_export(fgl_ctx* c, int w, int h, FglGroupMember* out)
{
    ENTER(c);
    if (int rc = fgl_not_while_recording(c, "fgl_group_export")) return rc;
    if (!out || w <= 0 || h <= 0) return fgl_fail(c, FGL_ERR_INVALID, "fgl_group_export: bad arguments");
    if (int rc = flush(c)) return rc;
    if (c->group.on) return fgl_fail(c, FGL_ERR_STATE, "fgl_group_export: disconnect the current group first");
    memset(out, 0, sizeof *out);
    // the exchanged planes get their final size now, so that their device addresses stay put
    if (int rc = plane_init(c, FGL_PLANE_SHADOW, w, h, 0.f)) return rc;
    if (int rc = plane_init(c, FGL_PLANE_DEPTH, w, h, FLT_MAX)) return rc;
    if (int rc = fgl_reserve(c, c->frameRgb8, (size_t)w * h * 3 + 16)) return rc;
    if (int rc = fgl_reserve(c, c->group.flags, sizeof(GroupFlagsD))) return rc;
    FGL_CUDA(c, cudaMemset(c->group.flags.p, 0, sizeof(GroupFlagsD)));
    void* chain = nullptr;
    if (int rc = fgl_stream_peer_mailbox(c, &chain, out->chain_ipc)) return rc;
    GroupState& g = c->group;
    g.expShadow = c->planes[FGL_PLANE_SHADOW].buf.p, g.expDepth = c->planes[FGL_PLANE_DEPTH].buf.p, g.expRgb8 = c->frameRgb8.p;
    g.W = w, g.H = h;
    out->shadow_ptr = g.expShadow, out->depth_ptr = g.expDepth, out->frame_ptr = g.expRgb8, out->flags_ptr = g.flags.p, out->chain_ptr = chain;
    out->width = w, out->height = h, out->device = c->device;
    cudaIpcMemHandle_t hnd;
    static_assert(sizeof hnd == 64, "CUDA IPC handles are 64 bytes");
    FGL_CUDA(c, cudaIpcGetMemHandle(&hnd, g.expShadow));
    memcpy(out->shadow_ipc, &hnd, 64);
    FGL_CUDA(c, cudaIpcGetMemHandle(&hnd, g.expDepth));
    memcpy(out->depth_ipc, &hnd, 64);
    FGL_CUDA(c, cudaIpcGetMemHandle(&hnd, g.expRgb8));
    memcpy(out->frame_ipc, &hnd, 64);
    FGL_CUDA(c, cudaIpcGetMemHandle(&hnd, g.flags.p));
    memcpy(out->flags_ipc, &hnd, 64);
    return FGL_OK;
}

int fgl_group_connect(fgl_ctx* c, int rank, int world, const FglGroupMember* m, int sameProcess)
{
    ENTER(c);
    if (int rc = fgl_not_while_recording(c, "fgl_group_connect")) return rc;
    if (!m || world < 1 || world > kMaxGroup || rank < 0 || rank >= world) return fgl_fail(c, FGL_ERR_INVALID, "fgl_group_connect: bad arguments");
    if (int rc = flush(c)) return rc;
    GroupState& g = c->group;
    if (g.on) return fgl_fail(c, FGL_ERR_STATE, "fgl_group_connect: already connected");
    if (!g.expShadow || m[rank].shadow_ptr != g.expShadow || m[rank].flags_ptr != g.flags.p)
        return fgl_fail(c, FGL_ERR_STATE, "fgl_group_connect: members[rank] is not this context's fgl_group_export record");
    for (int r = 0; r < world; ++r)
        if (m[r].width != g.W || m[r].height != g.H) return fgl_fail(c, FGL_ERR_INVALID, "fgl_group_connect: the members were exported for different frame sizes");
    auto map = [&](const unsigned char* ipc, void* ptr, int peerDevice, void** out) -> int {
        if (sameProcess)
        {
            if (peerDevice != c->device)
            {
                cudaError_t e = cudaDeviceEnablePeerAccess(peerDevice, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fgl_fail(c, FGL_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
                cudaGetLastError();
            }
            *out = ptr;
            return FGL_OK;
        }
        cudaIpcMemHandle_t hnd;
        memcpy(&hnd, ipc, 64);
        FGL_CUDA(c, cudaIpcOpenMemHandle(out, hnd, cudaIpcMemLazyEnablePeerAccess));
        g.ipcMapped.push_back(*out);
        return FGL_OK;
    };
    for (int r = 0; r < world; ++r)
    {
        if (r == rank)
        {
            g.peerShadow[r] = (float*)g.expShadow, g.peerDepth[r] = (float*)g.expDepth, g.peerFlags[r] = (GroupFlagsD*)g.flags.p;
            continue;
        }
        void *a = nullptr, *b = nullptr, *f = nullptr;
        if (int rc = map(m[r].shadow_ipc, m[r].shadow_ptr, m[r].device, &a)) return rc;
        if (int rc = map(m[r].depth_ipc, m[r].depth_ptr, m[r].device, &b)) return rc;
        if (int rc = map(m[r].flags_ipc, m[r].flags_ptr, m[r].device, &f)) return rc;
        g.peerShadow[r] = (float*)a, g.peerDepth[r] = (float*)b, g.peerFlags[r] = (GroupFlagsD*)f;
    }
    if (rank == 0) g.rootRgb8 = (uint8_t*)g.expRgb8;
    else
    {
        void* p = nullptr;
        if (int rc = map(m[0].frame_ipc, m[0].frame_ptr, m[0].device, &p)) return rc;
        g.rootRgb8 = (uint8_t*)p;
    }
    // the PCSS chain's blocker count travels band to band through the mailboxes of stream.cu
    const bool last = rank == world - 1;
    if (int rc = fgl_stream_peer_connect(c, last || !sameProcess ? nullptr : m[rank + 1].chain_ptr, last || sameProcess ? nullptr : m[rank + 1].chain_ipc, rank > 0, world > 1))
        return rc;
    g.rank = rank, g.world = world, g.epoch = 0, g.on = true;
    c->row0 = 0, c->row1 = -1;
    return FGL_OK;
}

int fgl_group_disconnect(fgl_ctx* c)
{
    ENTER(c);
    if (int rc = flush(c)) return rc;
    if (c->chainStream) FGL_CUDA(c, cudaStreamSynchronize(c->chainStream));
    FGL_CUDA(c, cudaStreamSynchronize(c->stream));
    fgl_stream_peer_connect(c, nullptr, nullptr, 0, 0);
    fgl_group_release(c);
    c->visCamClear = c->visLightClear = true;
    return FGL_OK;
}

int fgl_group_read_frame(fgl_ctx* c, void* dst, size_t bytes)
{
    ENTER(c);
    GroupState& g = c->group;
    if (!g.on || g.rank != 0) return fgl_fail(c, FGL_ERR_STATE, "fgl_group_read_frame: only rank 0 of a connected group holds the frame");
    if (!dst || bytes != (size_t)g.W * g.H * 3) return fgl_fail(c, FGL_ERR_INVALID, "fgl_group_read_frame: size mismatch");
    if (int rc = flush(c)) return rc;
    if (int rc = fgl_group_wait(c, FGL_GROUP_BAND)) return rc;
    FGL_CUDA(c, cudaMemcpyAsync(dst, g.expRgb8, bytes, cudaMemcpyDeviceToHost, c->stream));
    unsigned long long err = 0;
    FGL_CUDA(c, cudaMemcpyAsync(&err, &((GroupFlagsD*)g.flags.p)->error, 8, cudaMemcpyDeviceToHost, c->stream));
    FGL_CUDA(c, cudaStreamSynchronize(c->stream));
    c->d2hBytes += bytes;
    if (err) return fgl_fail(c, FGL_ERR_STATE, "sort-first group: a device-side wait for context " + std::to_string(err - 1) + " timed out (30 s)");
    return FGL_OK;
}

// ---- draws -----------------------------------------------------------------------------------------------------------
int fgl_draw_mesh(fgl_ctx* c, int meshId, int kind, const FglUniforms* un)
{
    ENTER(c);
    if (!un || meshId < 0 || meshId >= (int)c->meshes.size()) return fgl_fail(c, FGL_ERR_INVALID, "fgl_draw_mesh: bad mesh");
    if (kind < FGL_SHADER_DEPTH || kind > FGL_SHADER_PBR) return fgl_fail(c, FGL_ERR_INVALID, "fgl_draw_mesh: bad shader kind");
    if (c->pass == FGL_PASS_GEOMETRY && kind != FGL_SHADER_G)
        return fgl_fail(c, FGL_ERR_STATE, "geometry pass requires GShader (reference forkergl.cpp:214 dynamic_cast)");
    if (c->pass == FGL_PASS_SHADOW && kind != FGL_SHADER_DEPTH) return fgl_fail(c, FGL_ERR_STATE, "shadow pass requires DepthShader");
    if (c->pass == FGL_PASS_FORWARD && kind != FGL_SHADER_BLINN_PHONG && kind != FGL_SHADER_PBR)
        return fgl_fail(c, FGL_ERR_STATE, "forward pass requires BlinnPhongShader or PBRShader");
    const MeshH& m = c->meshes[meshId];
    if (m.nFaces == 0) return FGL_OK;
    if ((long long)c->primCounter + m.nFaces > 0x7fffffffLL) return fgl_fail(c, FGL_ERR_INVALID, "too many triangles in one pass");
    const VerticesH& vb = c->vertices[m.vertices];
    DrawCmdD         d;
    memset(&d, 0, sizeof d);
    d.pos = (const float*)vb.pos.p, d.uv = (const float*)vb.uv.p, d.nrm = (const float*)vb.nrm.p, d.tan = (const float*)vb.tan.p;
    d.pi = (const int*)m.pi.p, d.ti = (const int*)m.ti.p, d.ni = (const int*)m.ni.p;
    d.mat = m.mat, d.hasTangents = m.hasTangents, d.supportPBR = m.supportPBR, d.kind = kind;
    d.firstPrim = c->primCounter, d.nFaces = m.nFaces;
    memcpy(d.model, un->model, 64), memcpy(d.view, un->view, 64), memcpy(d.proj, un->projection, 64);
    memcpy(d.normal, un->normal, 36), memcpy(d.lightSpace, un->light_space, 64);
    memcpy(d.lightPos, un->light_position, 12), memcpy(d.lightColor, un->light_color, 12), memcpy(d.eye, un->eye_position, 12);
    for (int i = 0; i < 4; ++i)  // uLightSpaceMatrix * uModelMatrix (depthshader.h:23-24): Dot(row, column), geometry.h:784-793
        for (int j = 0; j < 4; ++j)
        {
            float r = 0.f;
            for (int k = 0; k < 4; ++k) r += un->light_space[i * 4 + k] * un->model[k * 4 + j];
            d.lm[i * 4 + j] = r;
        }
    c->draws.push_back(d);
    c->primCounter += m.nFaces;
    return FGL_OK;
}

// Device copy of a host array that has to stay alive until the pass ends (draws are deferred): chunks are only ever added
// while a pass is open, so pointers already recorded in draw commands stay valid.
static int stash(fgl_ctx* c, const float* src, size_t srcBytes, const float** out)
{
    const size_t bytes = (srcBytes + 255) & ~(size_t)255;  // slots are 256-byte aligned; only srcBytes are read from the caller
    if (c->preChunks.empty() || c->preUsed + bytes > c->preChunks.back().cap)
    {
        DevBuf b;
        size_t want = std::max<size_t>(bytes, (size_t)8 << 20);
        if (cudaMalloc(&b.p, want) != cudaSuccess) return fgl_fail(c, FGL_ERR_NOMEM, "cudaMalloc of " + std::to_string(want) + " bytes failed");
        b.cap = want;
        c->preChunks.push_back(b);
        c->preUsed = 0;
    }
    char* dst = (char*)c->preChunks.back().p + c->preUsed;
    FGL_CUDA(c, cudaMemcpyAsync(dst, src, srcBytes, cudaMemcpyHostToDevice, c->stream));
    FGL_CUDA(c, cudaStreamSynchronize(c->stream));  // the caller's arrays may be gone when this call returns
    c->h2dBytes += srcBytes;
    c->preUsed += bytes;
    *out = (const float*)dst;
    return FGL_OK;
}

int fgl_draw_triangles(fgl_ctx* c, int meshId, int kind, const FglUniforms* un, int n, const float* ndc, const float* vary, const float* lightZ)
{
    ENTER(c);
    if (int rc = fgl_not_while_recording(c, "fgl_draw_triangles (its arrays are copied synchronously)")) return rc;
    if (!un || meshId < 0 || meshId >= (int)c->meshes.size() || n < 0 || (n && !ndc)) return fgl_fail(c, FGL_ERR_INVALID, "fgl_draw_triangles: bad arguments");
    if (kind < FGL_SHADER_DEPTH || kind > FGL_SHADER_PBR) return fgl_fail(c, FGL_ERR_INVALID, "fgl_draw_triangles: bad shader kind");
    if (c->pass == FGL_PASS_GEOMETRY && kind != FGL_SHADER_G)
        return fgl_fail(c, FGL_ERR_STATE, "geometry pass requires GShader (reference forkergl.cpp:214 dynamic_cast)");
    if (c->pass == FGL_PASS_SHADOW && kind != FGL_SHADER_DEPTH) return fgl_fail(c, FGL_ERR_STATE, "shadow pass requires DepthShader");
    if (c->pass == FGL_PASS_FORWARD && kind != FGL_SHADER_BLINN_PHONG && kind != FGL_SHADER_PBR)
        return fgl_fail(c, FGL_ERR_STATE, "forward pass requires BlinnPhongShader or PBRShader");
    if (n && (kind == FGL_SHADER_DEPTH ? !lightZ : !vary)) return fgl_fail(c, FGL_ERR_INVALID, "fgl_draw_triangles: the shader kind's per-triangle array is NULL");
    if (n == 0) return FGL_OK;
    if ((long long)c->primCounter + n > 0x7fffffffLL) return fgl_fail(c, FGL_ERR_INVALID, "too many triangles in one pass");
    const MeshH& m = c->meshes[meshId];
    DrawCmdD     d;
    memset(&d, 0, sizeof d);
    if (int rc = stash(c, ndc, (size_t)n * 48, &d.preNdc)) return rc;
    if (kind == FGL_SHADER_DEPTH) { if (int rc = stash(c, lightZ, (size_t)n * 12, &d.preZ)) return rc; }
    else if (int rc = stash(c, vary, (size_t)n * 192, &d.preVary)) return rc;
    d.mat = m.mat, d.hasTangents = m.hasTangents, d.supportPBR = m.supportPBR, d.kind = kind;
    d.firstPrim = c->primCounter, d.nFaces = n;
    memcpy(d.lightPos, un->light_position, 12), memcpy(d.lightColor, un->light_color, 12), memcpy(d.eye, un->eye_position, 12);
    c->draws.push_back(d);
    c->primCounter += n;
    return FGL_OK;
}

// LightPass of the lighting loop (forkergl.cpp:326-380) for the current state; shared by the two entry points below
static int build_light_pass(fgl_ctx* c, const float eye[3], const float lpos[3], const float lcol[3], LightPass& L, bool& fullBand)
{
    if (!eye || !lpos || !lcol) return fgl_fail(c, FGL_ERR_INVALID, "NULL argument");
    if (int rc = flush(c)) return rc;
    PlaneH& frame = c->planes[FGL_PLANE_FRAME];
    if (!frame.buf.p) return fgl_fail(c, FGL_ERR_STATE, "DrawScreenSpacePixels before InitFrameBuffer");
    static const int inputs[] = { FGL_PLANE_NORMAL, FGL_PLANE_WORLDPOS, FGL_PLANE_ALBEDO, FGL_PLANE_EMISSIVE, FGL_PLANE_PARAM, FGL_PLANE_SHADINGTYPE, FGL_PLANE_AO };
    for (int pl : inputs)
    {
        const PlaneH& q = c->planes[pl];
        if (!q.buf.p || q.w != frame.w || q.h != frame.h) return fgl_fail(c, FGL_ERR_STATE, "DrawScreenSpacePixels: G-buffers missing or of another size");
    }
    memset(&L, 0, sizeof L);
    L.W = frame.w, L.H = frame.h;
    band_of(c, L.H, L.row0, L.row1);
    for (int pl : inputs)
        if (int rc = materialize_rows(c, pl, L.row0, L.row1)) return rc;
    fill_light_consts(c, L);
    if (c->shadowOn)
    {
        const PlaneH& q = c->planes[FGL_PLANE_LIGHTNDC];
        if (!q.buf.p || q.w != frame.w || q.h != frame.h || !c->planes[FGL_PLANE_SHADOW].buf.p)
            return fgl_fail(c, FGL_ERR_STATE, "DrawScreenSpacePixels: shadows on without LightSpaceNDCPosGBuffer / ShadowBuffer");
        if (int rc = materialize_rows(c, FGL_PLANE_LIGHTNDC, L.row0, L.row1)) return rc;
        if (c->group.on)
        {   // the other bands' rows of the shadow map arrive by peer stores
            c->planes[FGL_PLANE_SHADOW].fillPending = false;
            if (int rc = fgl_group_wait(c, FGL_GROUP_SHADOW)) return rc;
        }
        else if (int rc = materialize(c, FGL_PLANE_SHADOW)) return rc;
        L.sm = shadow_map_dev(c);
    }
    memcpy(L.eye, eye, 12), memcpy(L.lightPos, lpos, 12), memcpy(L.lightColor, lcol, 12);
    fullBand = L.row0 == 0 && L.row1 == L.H;
    L.planes = planes_dev(c);
    return FGL_OK;
}

// Optional first half of fgl_draw_screen_space_pixels: everything of a PCSS frame that does not depend on the bands above
// this one (shadow coordinates, min/max maps, classification, cell masks, the pilot).  A sort-first driver calls it
// before it waits for the previous band's blocker count; single-GPU callers never need it.
int fgl_prepare_screen_space_pixels(fgl_ctx* c, const float eye[3], const float lpos[3], const float lcol[3], int ssao_follows)
{
    ENTER(c);
    LightPass L;
    bool      fullBand = false;
    if (int rc = build_light_pass(c, eye, lpos, lcol, L, fullBand)) return rc;
    if (!(c->shadowOn && c->params.shadow_mode == FGL_SHADOW_PCSS)) return FGL_OK;
    if (ssao_follows)
    {   // SSAO consumes the sample stream first (render.cpp:204-209): the lighting phase starts where its 32 samples per pixel end
        SsaoPass S;
        memset(&S, 0, sizeof S);
        S.W = L.W, S.H = L.H;
        if (int rc = fgl_stream_prepare_ssao(c, S)) return rc;
    }
    if (int rc = fgl_stream_prepare_lighting(c, L, FGL_VIS_PREPARE)) return rc;
    // The chain itself can be issued right away when its input is known: a whole-frame context (nothing above it), or a
    // band whose input arrives on the device (fgl_chain_peer_connect).  It goes to its own stream, behind an event, so
    // that whatever the caller queues next on the main stream — SSAO, the blur — runs concurrently with it.
    static const bool noOverlap = getenv("FGL_NO_CHAIN_OVERLAP") != nullptr;
    // (per-kernel event timing is only meaningful without concurrency: the instrumented frames of bench.py run serially)
    if (noOverlap || c->timing || !(fullBand || fgl_stream_peer_on(c))) return FGL_OK;
    if (!c->chainStream)
    {
        int prLo = 0, prHi = 0;  // the chain is the critical path of the frame: its CTAs are placed before anything else
        cudaDeviceGetStreamPriorityRange(&prLo, &prHi);
        FGL_CUDA(c, cudaStreamCreateWithPriority(&c->chainStream, cudaStreamNonBlocking, prHi));
        FGL_CUDA(c, cudaEventCreateWithFlags(&c->evChainGo, cudaEventDisableTiming));
        FGL_CUDA(c, cudaEventCreateWithFlags(&c->evChainDone, cudaEventDisableTiming));
    }
    FGL_CUDA(c, cudaEventRecord(c->evChainGo, c->stream));
    FGL_CUDA(c, cudaStreamWaitEvent(c->chainStream, c->evChainGo, 0));
    cudaStream_t mainStream = c->stream;
    c->stream = c->chainStream;
    int rc = fgl_stream_prepare_lighting(c, L, FGL_VIS_LAUNCH);
    c->stream = mainStream;
    if (rc) return rc;
    FGL_CUDA(c, cudaEventRecord(c->evChainDone, c->chainStream));
    c->chainEventPending = true;
    return FGL_OK;
}

int fgl_draw_screen_space_pixels(fgl_ctx* c, const float eye[3], const float lpos[3], const float lcol[3])
{
    ENTER(c);
    LightPass L;
    bool      fullBand = false;
    if (int rc = build_light_pass(c, eye, lpos, lcol, L, fullBand)) return rc;
    PlaneH& frame = c->planes[FGL_PLANE_FRAME];
    if (L.writeF32)
    {
        if (fullBand) frame.fillPending = false;
        else if (int rc = materialize(c, FGL_PLANE_FRAME)) return rc;
    }
    size_t n = (size_t)L.W * L.H;
    if (int rc = fgl_reserve(c, c->frameRgb8, n * 3 + 16)) return rc;
    L.rgb8 = (uint8_t*)c->frameRgb8.p;
    if (c->group.on)
    {
        if (c->frameRgb8.p != c->group.expRgb8 || L.W != c->group.W || L.H != c->group.H)
            return fgl_fail(c, FGL_ERR_STATE, "the frame does not have the size the group was connected for (fgl_group_export)");
        L.rgb8 = c->group.rootRgb8;  // the band's 8-bit rows go straight into rank 0's frame
        if (int rc = fgl_group_wait(c, FGL_GROUP_READY)) return rc;
    }
    if (c->chainEventPending)
    {   // a chain issued by fgl_prepare_screen_space_pixels: everything from here on is ordered behind it
        FGL_CUDA(c, cudaStreamWaitEvent(c->stream, c->evChainDone, 0));
        c->chainEventPending = false;
    }
    if (c->shadowOn && c->params.shadow_mode != FGL_SHADOW_HARD)
        if (int rc = fgl_stream_prepare_lighting(c, L, FGL_VIS_RESOLVE)) return rc;  // continues a prepared chain, else does it all
    if (int rc = fgl_run_lighting(c, L)) return rc;
    c->frameRgb8Valid = fullBand;  // with a partial band only the band rows are current; readers take rows of the band
    c->bandRgb8Valid = !c->group.on;
    if (c->group.on) return fgl_group_signal(c, FGL_GROUP_BAND, /*rootOnly=*/true);
    return FGL_OK;
}

int fgl_ssao(fgl_ctx* c)
{
    ENTER(c);
    if (int rc = flush(c)) return rc;
    const PlaneH& frame = c->planes[FGL_PLANE_FRAME];
    const PlaneH& dp = c->planes[FGL_PLANE_DEPTH];
    if (!frame.buf.p || !dp.buf.p) return fgl_fail(c, FGL_ERR_STATE, "SSAO before InitFrameBuffer / InitDepthBuffer");
    static const int inputs[] = { FGL_PLANE_NORMAL, FGL_PLANE_WORLDPOS, FGL_PLANE_DEPTH, FGL_PLANE_AO };
    for (int pl : inputs)
    {
        const PlaneH& q = c->planes[pl];
        if (!q.buf.p || q.w != frame.w || q.h != frame.h) return fgl_fail(c, FGL_ERR_STATE, "SSAO: G-buffers missing or of another size");
    }
    SsaoPass S;
    memset(&S, 0, sizeof S);
    S.W = frame.w, S.H = frame.h;
    halo_band_of(c, S.H, S.row0, S.row1);
    for (int pl : inputs)
        if (int rc = pl == FGL_PLANE_DEPTH ? materialize(c, pl) : materialize_rows(c, pl, S.row0, S.row1)) return rc;  // depth is gathered from anywhere
    S.worldpos = (const float*)c->planes[FGL_PLANE_WORLDPOS].buf.p, S.normal = (const float*)c->planes[FGL_PLANE_NORMAL].buf.p;
    S.depth = (const float*)dp.buf.p, S.ao = (float*)c->planes[FGL_PLANE_AO].buf.p;
    memcpy(S.viewProj, c->viewProj, 64), memcpy(S.viewport, c->viewport, 64);
    S.radius = c->params.ssao_radius, S.rangeCheckRadius = c->params.ssao_range_check_radius, S.bias = c->params.ssao_bias;
    S.rangeCheck = c->params.ssao_range_check;
    if (int rc = fgl_stream_prepare_ssao(c, S)) return rc;
    if (int rc = fgl_group_wait(c, FGL_GROUP_DEPTH)) return rc;  // group: the other bands' depth rows arrive by peer stores
    return fgl_run_ssao(c, S);
}

int fgl_blur(fgl_ctx* c, int plane, int kind)
{
    ENTER(c);
    if (plane < 0 || plane > FGL_PLANE_AO) return fgl_fail(c, FGL_ERR_INVALID, "fgl_blur: not an fp32 plane");
    if (int rc = flush(c)) return rc;
    PlaneH& p = c->planes[plane];
    if (!p.buf.p) return fgl_fail(c, FGL_ERR_STATE, "fgl_blur: plane not initialised");
    if (plane == FGL_PLANE_FRAME) c->frameRgb8Valid = c->bandRgb8Valid = false;
    int h0, h1, v0, v1;
    halo_band_of(c, p.h, h0, h1);
    if (int rc = kind == FGL_BLUR_TWO_PASS_GAUSSIAN ? materialize_rows(c, plane, h0, h1) : materialize(c, plane)) return rc;
    band_of(c, p.h, v0, v1);
    v0 = std::max(0, v0 - kBlurWarm);
    return fgl_run_blur(c, (float*)p.buf.p, p.w, p.h, p.ch, kind, h0, h1, v0, v1);
}

int fgl_ssaa_resolve(fgl_ctx* c, int k)
{
    ENTER(c);
    if (k < 1) return fgl_fail(c, FGL_ERR_INVALID, "fgl_ssaa_resolve: kernel size < 1");
    if (c->group.on) return fgl_fail(c, FGL_ERR_UNSUPPORTED, "SSAA is not available in a sort-first group");
    if (int rc = flush(c)) return rc;
    const PlaneH& f = c->planes[FGL_PLANE_FRAME];
    if (!f.buf.p) return fgl_fail(c, FGL_ERR_STATE, "SSAA before InitFrameBuffer");
    int r0, r1;
    band_of(c, f.h, r0, r1);
    bool fullBand = r0 == 0 && r1 == f.h;
    if (!fullBand && (r0 % k != 0 || (r1 % k != 0 && r1 != f.h)))
        return fgl_fail(c, FGL_ERR_INVALID, "fgl_ssaa_resolve: row band boundaries must be multiples of the SSAA kernel size");
    if (!(c->frameRgb8Valid || (!fullBand && c->bandRgb8Valid)))
        if (int rc = ensure_rgb8(c)) return rc;
    int ow = f.w / k, oh = f.h / k;
    c->ssaaW = ow, c->ssaaH = oh;
    if (int rc = fgl_reserve(c, c->ssaaRgb8, (size_t)ow * oh * 3 + 16)) return rc;
    return fgl_run_ssaa(c, (const uint8_t*)c->frameRgb8.p, f.w, f.h, k, (uint8_t*)c->ssaaRgb8.p, r0, std::min(r1, oh * k));
}

// ---- buffers ---------------------------------------------------------------------------------------------------------
int fgl_plane_info(fgl_ctx* c, int plane, int* w, int* h, int* ch, int* bpc)
{
    ENTER(c);
    if (plane < 0 || plane >= FGL_PLANE_COUNT || !w || !h || !ch || !bpc) return fgl_fail(c, FGL_ERR_INVALID, "fgl_plane_info: bad arguments");
    if (int rc = flush(c)) return rc;  // pending draws define the size of the winner-id planes
    if (plane <= FGL_PLANE_AO) *w = c->planes[plane].w, *h = c->planes[plane].h, *ch = is1ch(plane) ? 1 : 3, *bpc = 4;
    else if (plane == FGL_PLANE_FRAME_RGB8) *w = c->planes[FGL_PLANE_FRAME].w, *h = c->planes[FGL_PLANE_FRAME].h, *ch = 3, *bpc = 1;
    else if (plane == FGL_PLANE_SSAA_RGB8) *w = c->ssaaW, *h = c->ssaaH, *ch = 3, *bpc = 1;
    else if (plane == FGL_PLANE_PRIMID_CAMERA) *w = c->visCamW, *h = c->visCamH, *ch = 1, *bpc = 4;
    else *w = c->visLightW, *h = c->visLightH, *ch = 1, *bpc = 4;
    return FGL_OK;
}

// device pointer + byte size of a plane in the reference's layout (AoS); 3-channel fp32 planes go through scanTmp
static int plane_as_aos(fgl_ctx* c, int plane, const void** src, size_t* bytes)
{
    if (plane <= FGL_PLANE_AO)
    {
        PlaneH& p = c->planes[plane];
        size_t  n = (size_t)p.w * p.h;
        *bytes = n * p.ch * 4;
        if (!p.buf.p) { *src = nullptr; *bytes = 0; return FGL_OK; }
        if (int rc = materialize(c, plane)) return rc;
        if (p.ch == 1) { *src = p.buf.p; return FGL_OK; }
        if (int rc = fgl_reserve(c, c->scanTmp, *bytes)) return rc;
        if (int rc = fgl_run_aos(c, (const float*)p.buf.p, (float*)c->scanTmp.p, n, p.ch, true)) return rc;
        *src = c->scanTmp.p;
        return FGL_OK;
    }
    if (plane == FGL_PLANE_FRAME_RGB8 && c->group.on)
    {   // the 8-bit frame of a sort-first group lives on rank 0, its rows are stored there by every band's lighting kernel
        if (c->group.rank != 0) return fgl_fail(c, FGL_ERR_STATE, "the 8-bit frame of a sort-first group is gathered on rank 0");
        if (int rc = fgl_group_wait(c, FGL_GROUP_BAND)) return rc;
        *src = c->group.expRgb8, *bytes = (size_t)c->group.W * c->group.H * 3;
        return FGL_OK;
    }
    if (plane == FGL_PLANE_FRAME_RGB8)
    {
        if (int rc = ensure_rgb8(c)) return rc;
        *src = c->frameRgb8.p, *bytes = (size_t)c->planes[FGL_PLANE_FRAME].w * c->planes[FGL_PLANE_FRAME].h * 3;
        return FGL_OK;
    }
    if (plane == FGL_PLANE_SSAA_RGB8)
    {
        *src = c->ssaaRgb8.p, *bytes = (size_t)c->ssaaW * c->ssaaH * 3;
        return FGL_OK;
    }
    bool    cam = plane == FGL_PLANE_PRIMID_CAMERA;
    DevBuf& vis = cam ? c->visCamera : c->visLight;
    size_t  n = (size_t)(cam ? c->visCamW : c->visLightW) * (cam ? c->visCamH : c->visLightH);
    *bytes = n * 4;
    if (!n) { *src = nullptr; return FGL_OK; }
    bool& pendingClear = cam ? c->visCamClear : c->visLightClear;
    if (pendingClear && vis.p)
    {
        FGL_CUDA(c, cudaMemsetAsync(vis.p, 0xFF, n * 8, c->stream));
        pendingClear = false;
    }
    if (int rc = fgl_reserve(c, c->scanTmp, n * 4)) return rc;
    if (int rc = fgl_run_ids(c, (const unsigned long long*)vis.p, n, (int*)c->scanTmp.p)) return rc;
    *src = c->scanTmp.p;
    return FGL_OK;
}

int fgl_read_plane(fgl_ctx* c, int plane, void* dst, size_t bytes)
{
    ENTER(c);
    if (int rc = fgl_not_while_recording(c, "fgl_read_plane")) return rc;
    if (plane < 0 || plane >= FGL_PLANE_COUNT || !dst) return fgl_fail(c, FGL_ERR_INVALID, "fgl_read_plane: bad arguments");
    if (int rc = flush(c)) return rc;
    const void* src = nullptr;
    size_t      n = 0;
    if (int rc = plane_as_aos(c, plane, &src, &n)) return rc;
    if (bytes != n) return fgl_fail(c, FGL_ERR_INVALID, "fgl_read_plane: size mismatch (have " + std::to_string(n) + ")");
    if (n) FGL_CUDA(c, cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, c->stream));
    FGL_CUDA(c, cudaStreamSynchronize(c->stream));
    c->d2hBytes += n;
    return FGL_OK;
}

int fgl_host_alloc(fgl_ctx* c, size_t bytes, void** out)
{
    ENTER(c);
    if (!out) return fgl_fail(c, FGL_ERR_INVALID, "fgl_host_alloc: out is NULL");
    *out = nullptr;
    FGL_CUDA(c, cudaHostAlloc(out, std::max<size_t>(bytes, 16), cudaHostAllocDefault));
    return FGL_OK;
}
int fgl_host_free(fgl_ctx* c, void* p)
{
    ENTER(c);
    if (p) FGL_CUDA(c, cudaFreeHost(p));
    return FGL_OK;
}

int fgl_write_plane(fgl_ctx* c, int plane, const void* src, size_t bytes)
{
    ENTER(c);
    if (int rc = fgl_not_while_recording(c, "fgl_write_plane")) return rc;
    if (plane < 0 || plane > FGL_PLANE_AO || !src) return fgl_fail(c, FGL_ERR_INVALID, "fgl_write_plane: bad arguments");
    if (int rc = flush(c)) return rc;
    PlaneH& p = c->planes[plane];
    size_t  n = (size_t)p.w * p.h;
    if (!p.buf.p || bytes != n * p.ch * 4) return fgl_fail(c, FGL_ERR_INVALID, "fgl_write_plane: size mismatch");
    p.fillPending = false;
    c->h2dBytes += bytes;
    if (plane == FGL_PLANE_FRAME) c->frameRgb8Valid = c->bandRgb8Valid = false;
    if (p.ch == 1) FGL_CUDA(c, cudaMemcpyAsync(p.buf.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
    else
    {
        if (int rc = fgl_reserve(c, c->scanTmp, bytes)) return rc;
        FGL_CUDA(c, cudaMemcpyAsync(c->scanTmp.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
        if (int rc = fgl_run_aos(c, (const float*)c->scanTmp.p, (float*)p.buf.p, n, p.ch, false)) return rc;
    }
    FGL_CUDA(c, cudaStreamSynchronize(c->stream));
    return FGL_OK;
}

int fgl_copy_plane_rows_to_device(fgl_ctx* c, int plane, int row0, int row1, void* dst, size_t bytes)
{
    ENTER(c);
    if (int rc = fgl_not_while_recording(c, "fgl_copy_plane_rows_to_device")) return rc;
    if (plane < 0 || plane >= FGL_PLANE_COUNT || !dst || row0 < 0 || row1 < row0) return fgl_fail(c, FGL_ERR_INVALID, "fgl_copy_plane_rows_to_device: bad arguments");
    if (int rc = flush(c)) return rc;
    int w, h, ch, bpc;
    fgl_plane_info(c, plane, &w, &h, &ch, &bpc);
    if (row1 > h) return fgl_fail(c, FGL_ERR_INVALID, "fgl_copy_plane_rows_to_device: rows out of range");
    size_t rowBytes = (size_t)w * ch * bpc, want = rowBytes * (row1 - row0);
    if (bytes != want) return fgl_fail(c, FGL_ERR_INVALID, "fgl_copy_plane_rows_to_device: size mismatch (need " + std::to_string(want) + ")");
    const void* src = nullptr;
    size_t      n = 0;
    if (plane == FGL_PLANE_FRAME_RGB8 && c->bandRgb8Valid && !c->frameRgb8Valid)
    {   // sort-first: this GPU lit only its band; its rows are current
        int b0, b1;
        band_of(c, h, b0, b1);
        if (row0 < b0 || row1 > b1) return fgl_fail(c, FGL_ERR_STATE, "rows outside the band this context rendered");
        src = c->frameRgb8.p;
    }
    else if (int rc = plane_as_aos(c, plane, &src, &n)) return rc;
    if (want) FGL_CUDA(c, cudaMemcpyAsync(dst, (const uint8_t*)src + rowBytes * row0, want, cudaMemcpyDeviceToDevice, c->stream));
    return FGL_OK;
}

// ---- recorded frames (SURVEY.md §8 f, N4: multi-frame use) ----------------------------------------------------------------
// A frame without host read-backs is a fixed sequence of copies, clears and kernels on one stream: captured once into a CUDA
// graph, it is replayed with a single launch.  Nothing is executed while recording; every entry point that would have to wait
// for the device or allocate refuses (fgl_not_while_recording), and the recording then ends with an error instead of a graph.
int fgl_frame_record_begin(fgl_ctx* c)
{
    ENTER(c);
    if (c->recording) return fgl_fail(c, FGL_ERR_STATE, "fgl_frame_record_begin: already recording");
    if (c->group.on) return fgl_fail(c, FGL_ERR_UNSUPPORTED, "a context of a sort-first group cannot record frames (its kernels wait for other GPUs)");
    if (c->timing) return fgl_fail(c, FGL_ERR_UNSUPPORTED, "switch per-kernel timing off before recording a frame");
    if (int rc = flush(c)) return rc;
    if (c->chainEventPending)
    {   // (an event of another stream recorded outside the capture cannot be waited for inside it)
        FGL_CUDA(c, cudaStreamWaitEvent(c->stream, c->evChainDone, 0));
        c->chainEventPending = false;
    }
    if (!c->recStaging) FGL_CUDA(c, cudaHostAlloc(&c->recStaging, kRecStagingBytes, cudaHostAllocDefault));
    c->recStagingUsed = 0;
    c->recLaunches0 = c->launches, c->recH2d0 = c->h2dBytes;
    // relaxed: other host threads (frames in flight on other contexts) keep calling CUDA while this thread records
    FGL_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
    c->recording = true;
    return FGL_OK;
}

int fgl_frame_record_end(fgl_ctx* c, int* frameId)
{
    ENTER(c);
    if (!c->recording) return fgl_fail(c, FGL_ERR_STATE, "fgl_frame_record_end without fgl_frame_record_begin");
    int         rcFlush = flush(c);  // triangles still pending belong to the frame
    std::string why = rcFlush ? c->error : std::string();
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    c->recording = false;
    RecordedFrame::HostState hs;
    save_host_state(c, hs);
    const uint64_t launches = c->launches - c->recLaunches0, h2d = c->h2dBytes - c->recH2d0;
    c->launches = c->recLaunches0, c->h2dBytes = c->recH2d0;  // nothing has run yet: replays are counted when they are issued
    // host-side bookkeeping now describes a frame that has not been rendered: make the next eager pass start clean
    c->visCamClear = c->visLightClear = true;
    c->frameRgb8Valid = c->bandRgb8Valid = false;
    if (e != cudaSuccess || !graph || rcFlush)
    {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return fgl_fail(c, FGL_ERR_UNSUPPORTED, "the frame could not be recorded: " + (why.empty() ? std::string(cudaGetErrorString(e)) : why));
    }
    if (!frameId)
    {
        cudaGraphDestroy(graph);
        return fgl_fail(c, FGL_ERR_INVALID, "fgl_frame_record_end: frame_id is NULL");
    }
    size_t nodes = 0;
    cudaGraphGetNodes(graph, nullptr, &nodes);
    RecordedFrame* f = nullptr;
    if (*frameId >= 0)
    {
        if (*frameId >= (int)c->recorded.size() || !c->recorded[*frameId].staging)
        {
            cudaGraphDestroy(graph);
            return fgl_fail(c, FGL_ERR_INVALID, "fgl_frame_record_end: no such recorded frame");
        }
        f = &c->recorded[*frameId];
        // a replay of the old recording may still be reading its staging
        cudaStreamSynchronize(c->stream);
        cudaGraphExecUpdateResultInfo info;
        memset(&info, 0, sizeof info);
        if (f->exec && cudaGraphExecUpdate(f->exec, graph, &info) != cudaSuccess)
        {   // another shape (a pass more or less): instantiate afresh
            cudaGetLastError();
            cudaGraphExecDestroy(f->exec);
            f->exec = nullptr;
        }
        if (f->staging) cudaFreeHost(f->staging);
        f->staging = nullptr;
    }
    else
    {
        c->recorded.emplace_back();
        f = &c->recorded.back();
        *frameId = (int)c->recorded.size() - 1;
    }
    if (!f->exec)
    {
        e = cudaGraphInstantiate(&f->exec, graph, 0);
        if (e != cudaSuccess)
        {
            cudaGraphDestroy(graph);
            f->exec = nullptr;
            return fgl_fail(c, FGL_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
        }
    }
    cudaGraphDestroy(graph);
    f->staging = c->recStaging, c->recStaging = nullptr;  // the graph's copy nodes read from it at every replay
    f->launches = launches, f->h2dBytes = h2d, f->nodes = nodes, f->host = hs;
    return FGL_OK;
}

int fgl_frame_record_abort(fgl_ctx* c)
{
    ENTER(c);
    if (!c->recording) return FGL_OK;
    cudaGraph_t graph = nullptr;
    cudaStreamEndCapture(c->stream, &graph);
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    c->recording = false;
    c->launches = c->recLaunches0, c->h2dBytes = c->recH2d0;
    // whatever the aborted sequence left pending or marked valid never happened
    c->draws.clear();
    c->primCounter = c->flushedPrims = 0;
    c->visCamClear = c->visLightClear = true;
    c->depthInitPending = c->depthInitBound = false;
    c->frameRgb8Valid = c->bandRgb8Valid = false;
    return FGL_OK;
}

int fgl_frame_replay(fgl_ctx* c, int frameId)
{
    ENTER(c);
    if (int rc = fgl_not_while_recording(c, "fgl_frame_replay")) return rc;
    if (frameId < 0 || frameId >= (int)c->recorded.size() || !c->recorded[frameId].staging) return fgl_fail(c, FGL_ERR_INVALID, "fgl_frame_replay: no such recorded frame");
    if (!c->recorded[frameId].exec)
        return fgl_fail(c, FGL_ERR_STATE, "fgl_frame_replay: a device buffer was re-allocated after this frame was recorded (another frame size?): record it again");
    if (c->group.on) return fgl_fail(c, FGL_ERR_UNSUPPORTED, "fgl_frame_replay: not inside a sort-first group");
    if (int rc = flush(c)) return rc;
    const RecordedFrame& f = c->recorded[frameId];
    FGL_CUDA(c, cudaGraphLaunch(f.exec, c->stream));
    c->launches += f.launches, c->h2dBytes += f.h2dBytes;
    restore_host_state(c, f.host);  // the planes now hold the recorded frame
    c->draws.clear();
    return FGL_OK;
}

int fgl_frame_release(fgl_ctx* c, int frameId)
{
    ENTER(c);
    if (frameId < 0 || frameId >= (int)c->recorded.size()) return fgl_fail(c, FGL_ERR_INVALID, "fgl_frame_release: no such recorded frame");
    RecordedFrame& f = c->recorded[frameId];
    cudaStreamSynchronize(c->stream);
    if (f.exec) cudaGraphExecDestroy(f.exec);
    if (f.staging) cudaFreeHost(f.staging);
    f = RecordedFrame();
    return FGL_OK;
}

int fgl_frame_info(fgl_ctx* c, int frameId, int* nodes, int* launches)
{
    ENTER(c);
    if (frameId < 0 || frameId >= (int)c->recorded.size() || !c->recorded[frameId].staging) return fgl_fail(c, FGL_ERR_INVALID, "fgl_frame_info: no such recorded frame");
    if (nodes) *nodes = (int)c->recorded[frameId].nodes;
    if (launches) *launches = (int)c->recorded[frameId].launches;
    return FGL_OK;
}

// ---- instrumentation ---------------------------------------------------------------------------------------------------
int fgl_enable_timing(fgl_ctx* c, int on)
{
    ENTER(c);
    if (on)
        if (int rc = fgl_not_while_recording(c, "per-kernel event timing")) return rc;
    c->timing = on != 0;
    return FGL_OK;
}
int fgl_reset_timings(fgl_ctx* c)
{
    ENTER(c);
    if (int rc = fgl_not_while_recording(c, "fgl_reset_timings")) return rc;
    cudaStreamSynchronize(c->stream);
    for (auto& t : c->timings) c->eventPool.push_back(t.e0), c->eventPool.push_back(t.e1);
    c->timings.clear();
    return FGL_OK;
}
int fgl_get_timings(fgl_ctx* c, FglTiming* out, int max, int* count)
{
    ENTER(c);
    if (int rc = fgl_not_while_recording(c, "fgl_get_timings")) return rc;
    FGL_CUDA(c, cudaStreamSynchronize(c->stream));
    std::vector<std::string> order;
    std::map<std::string, FglTiming> agg;
    for (auto& t : c->timings)
    {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t.e0, t.e1);
        auto it = agg.find(t.name);
        if (it == agg.end())
        {
            FglTiming f;
            memset(&f, 0, sizeof f);
            strncpy(f.name, t.name.c_str(), sizeof f.name - 1);
            it = agg.emplace(t.name, f).first;
            order.push_back(t.name);
        }
        uint64_t bytes = t.bytes;
        if (t.lateCount)
        {   // byte figure that depends on a count the device produced (e.g. the PCSS entries that were actually filtered)
            unsigned n = 0;
            if (cudaMemcpy(&n, t.lateCount, 4, cudaMemcpyDeviceToHost) == cudaSuccess) bytes += (uint64_t)n * t.lateBytesEach;
        }
        it->second.ms_total += ms, it->second.launches += 1, it->second.algorithmic_bytes += bytes;
    }
    int n = 0;
    for (auto& name : order)
        if (out && n < max) out[n++] = agg[name];
    if (count) *count = n;
    return FGL_OK;
}
int fgl_launch_count(fgl_ctx* c, uint64_t* o) { ENTER(c); if (o) *o = c->launches; return FGL_OK; }
int fgl_transfer_bytes(fgl_ctx* c, uint64_t* h2d, uint64_t* d2h)
{
    ENTER(c);
    if (h2d) *h2d = c->h2dBytes;
    if (d2h) *d2h = c->d2hBytes;
    return FGL_OK;
}

}  // extern "C"
