// fgl_internal.h — context layout and kernel entry points shared by the translation units of libforkergl_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "device_math.cuh"

// ---- device-side records -----------------------------------------------------------------------------------
// One draw call (= one Mesh::Draw of the reference, mesh.cpp:10-25) as the kernels see it.
struct DrawCmdD
{
    const float *pos, *uv, *nrm, *tan;
    const int *  pi, *ti, *ni;
    FglMaterial  mat;
    int          hasTangents, supportPBR, kind, firstPrim, nFaces;
    float        model[16], view[16], proj[16], normal[9], lightSpace[16];
    float        lm[16];  // DepthShader: uLightSpaceMatrix * uModelMatrix (depthshader.h:23-24)
    float        lightPos[3], lightColor[3], eye[3];
    // fgl_draw_triangles: the vertex programs ran on the host; per triangle 12 floats of NDC, 48 of varyings, 3 of light-space z
    const float *preNdc, *preVary, *preZ;
};

// Per-triangle setup record, 48 bytes (three 16-byte words):
//   w0 = X0 X1 X2 Y0 ; w1 = Y1 Y2 d0 d1 ; w2 = d2, bbox x (lo16 = xmin, hi16 = xmax), bbox y, flags
struct TriSetup
{
    int4 w0, w1, w2;
};
enum { TRI_SKIP = 1, TRI_SMALL = 2, TRI_LARGE = 4 };

// Per-triangle varyings of the camera-space programs (gshader.h:71-92), every attribute already times 1/w_clip.
// 48 floats = 192 bytes: posWS[3], nrmWS[3], tanWS[3], lightNDC[3] as xyz triples, then u[3], v[3], oow[3], pad.
struct TriVary
{
    float f[48];
};

// grow-only device allocation
struct DevBuf
{
    void*  p = nullptr;
    size_t cap = 0;
};

// ---- sort-first group (one context per GPU of an NVLink box; include/forkergl_b200.h "fgl_group_*") ---------------------
constexpr int kMaxGroup = 16;
// where a kernel stores a band's rows: the same plane in every context of the group (this context included)
struct PeerPlanes
{
    float* p[kMaxGroup];
    int    n;
};
// one per context, in its own HBM; every peer stores the current frame number into ITS slot of the arrays (peer stores over
// NVLink, system-scope release), the owner's wait kernels spin on them
struct GroupFlagsD
{
    unsigned long long ready[kMaxGroup];   // peer r has begun frame e: this context's planes may receive frame e's rows from... (see group.cu)
    unsigned long long shadow[kMaxGroup];  // peer r's rows of the shadow map of frame e have arrived
    unsigned long long depth[kMaxGroup];   // peer r's rows of the camera depth plane of frame e have arrived
    unsigned long long band[kMaxGroup];    // (rank 0 only) peer r's rows of the finished 8-bit frame e have arrived
    unsigned long long error;              // a wait timed out
};
struct GroupState
{
    bool               on = false;
    int                rank = 0, world = 1, W = 0, H = 0;
    unsigned long long epoch = 0;
    DevBuf             flags;  // this context's GroupFlagsD
    GroupFlagsD*       peerFlags[kMaxGroup] = { nullptr };
    float*             peerShadow[kMaxGroup] = { nullptr };
    float*             peerDepth[kMaxGroup] = { nullptr };
    uint8_t*           rootRgb8 = nullptr;  // rank 0's 8-bit frame
    void *             expShadow = nullptr, *expDepth = nullptr, *expRgb8 = nullptr;  // what fgl_group_export handed out
    std::vector<void*> ipcMapped;
    bool               readyWaited = false, shadowWaited = false, depthWaited = false, bandWaited = false;  // per frame
};

struct RasterPass
{
    // target
    int W, H, row0, row1;  // buffer size and the row band camera passes are restricted to
    int passType, shadowOn;
    // sort-first group: only triangles that touch rows [cull0, cull1) are set up and rasterised (everything: 0, H);
    // rows [own0, own1) of the pass's exchanged plane (shadow map / camera depth) are stored into every context of the group
    int        cull0, cull1, own0, own1, group;
    PeerPlanes peers;
    float viewport[16];
    // geometry
    const DrawCmdD* draws;
    int             nDraws, nPrims;
    TriSetup*       setup;
    TriVary*        vary;   // camera passes
    float4*         zndc;   // shadow pass: DepthShader's vPositionNDC z row
    int*            nblk;   // per triangle: number of 32x32 blocks of the large path (0 otherwise)
    int*            blkScan;  // exclusive scan of nblk, nPrims + 1 entries
    unsigned long long* vis;
    const TexD*     textures;
    // forward mode with a stochastic shadow filter: per-pixel fragment lists and the winner's stream position
    unsigned *          fragCount, *fragOffset;
    unsigned long long* frags;
    const int*          siteOfPixel;
};

struct PlanesD
{
    float* p[FGL_PLANE_AO + 1];  // SoA: channel c of pixel i at p[plane][c * N + i]
};

struct LightPass
{
    int     W, H, row0, row1;
    PlanesD planes;
    ShadowMapD sm;
    int     shadowOn, shadowMode, useAO, writeF32;
    float   eye[3], lightPos[3], lightColor[3];
    float   biasSlope, biasMin, shadowIntensity, areaLight;
    double  pcfFilter, pcssFilter;
    uint8_t* rgb8;
    const float2* disk;        // accepted unit-disk samples of the replayed stream (lighting phase)
    const unsigned* chunkOf;   // PCSS: per pixel index of its first 32-sample chunk; NULL = PCF (2 * pixel)
    const float* vis;          // PCF / PCSS: visibility per pixel, computed by the warp-per-pixel filter kernels (stream.cu)
};

// ---- host-side objects -----------------------------------------------------------------------------------

struct PlaneH
{
    int    w = 0, h = 0, ch = 0;
    DevBuf buf;
    bool   fillPending = false;
    float  fillValue = 0.f;
    float  fillRGB[3] = { 0, 0, 0 };
    bool   fillIsRGB = false;
    // sort-first bands: rows [validRow0, validRow1) were written by this band's passes; only the others still need the clear
    // value, and only if somebody reads them (a host read of the whole plane) — fillPending stays set until then
    bool   fillPartial = false;
    int    validRow0 = 0, validRow1 = 0;
};

struct TextureH
{
    DevBuf data;
    TexD   desc;
};
struct VerticesH
{
    DevBuf pos, uv, nrm, tan;
    int    nPos = 0, nUv = 0, nNrm = 0, nTan = 0;
};
struct MeshH
{
    int         vertices = -1, nFaces = 0;
    DevBuf      pi, ti, ni;
    FglMaterial mat;
    int         hasTangents = 0, supportPBR = 0;
};

struct TimingRec
{
    std::string name;
    cudaEvent_t e0, e1;
    uint64_t    bytes;
    const unsigned* lateCount = nullptr;  // device counter read when the timings are fetched: bytes += *lateCount * lateBytesEach
    uint64_t        lateBytesEach = 0;
};

// A frame recorded as a CUDA graph (include/forkergl_b200.h "fgl_frame_*"): the instantiated graph, the page-locked staging
// its host-to-device copies read from at every replay, and what one replay amounts to for the counters.
struct RecordedFrame
{
    cudaGraphExec_t exec = nullptr;
    void*           staging = nullptr;
    uint64_t        launches = 0, h2dBytes = 0;
    size_t          nodes = 0;
    // the context's host-side bookkeeping as the recorded call sequence left it (restored by every replay, so that the context
    // looks exactly as it does after rendering the frame eagerly): plane sizes and pending clears, validity of the 8-bit
    // images, the visibility buffers' sizes, the pass state
    struct HostState
    {
        struct Plane { int w, h, ch; bool fillPending, fillIsRGB, fillPartial; float fillValue, fillRGB[3]; int validRow0, validRow1; } planes[16];
        bool frameRgb8Valid, bandRgb8Valid, visCamClear, visLightClear, depthInitPending, depthInitBound, passRestarted;
        int  ssaaW, ssaaH, visCamW, visCamH, visLightW, visLightH, pass, primCounter, flushedPrims;
    } host;
};

struct fgl_ctx
{
    int          device = 0, numSMs = 148;  // SM count of the device (persistent grids are sized in multiples of it)
    std::string  error;
    cudaStream_t ownStream = nullptr, stream = nullptr;
    // the PCSS chain kernel runs on its own stream so that SSAO and the blur (main stream) overlap it
    cudaStream_t chainStream = nullptr;
    cudaEvent_t  evChainGo = nullptr, evChainDone = nullptr;
    bool         chainEventPending = false;
    FglParams    params;
    float        viewport[16], viewProj[16], lightSpace[16];
    int          mode = FGL_MODE_FORWARD, pass = FGL_PASS_FORWARD, shadowOn = 1;
    int          row0 = 0, row1 = -1;  // band; row1 < 0 = whole buffer
    unsigned long long chainBlockersBefore = 0;  // sort-first: blockers found by the bands above this one (PCSS chain input)

    PlaneH planes[FGL_PLANE_AO + 1];
    DevBuf frameRgb8, ssaaRgb8, blurTmp;
    bool   frameRgb8Valid = false, bandRgb8Valid = false;
    int    ssaaW = 0, ssaaH = 0;

    // visibility buffers (depth|primitive id keys)
    DevBuf visCamera, visLight;
    int    visCamW = 0, visCamH = 0, visLightW = 0, visLightH = 0;
    bool   visCamClear = false, visLightClear = false, depthInitPending = false, depthInitBound = false, passRestarted = false;

    // resources
    std::vector<TextureH>  textures;
    std::vector<VerticesH> vertices;
    std::vector<MeshH>     meshes;
    DevBuf                 texTable;
    bool                   texTableDirty = true;

    // pending draws of the current pass
    std::vector<DrawCmdD> draws;
    int                   primCounter = 0;  // submission index of the next triangle in this pass
    int                   flushedPrims = 0;
    DevBuf                drawsDev, setup, vary, zndc, nblk, blkScan, scanTmp, tileState;
    DevBuf                fragCount, fragOffset, frags, nPass, passOff, siteOfPixel, siteKeys, siteVals, siteSc4, sortTmp;
    std::vector<DevBuf>   preChunks;  // device copies of fgl_draw_triangles arrays, alive until the pass ends
    size_t                preUsed = 0;
    void*                 pinned = nullptr;
    size_t                pinnedCap = 0;

    // sample stream (rng.cu)
    struct SampleStream* stream_state = nullptr;

    // instrumentation
    bool                   timing = false;
    std::vector<TimingRec> timings;
    std::vector<cudaEvent_t> eventPool;
    uint64_t               launches = 0, h2dBytes = 0, d2hBytes = 0;
    GroupState             group;
    int                    lastUncertain = 0, lastChainIters = 0;  // PCSS: uncertain pixels / super-chunks of the last chain (diagnostics)

    // frame recording (CUDA graph capture of the context's stream)
    bool                       recording = false;
    void*                      recStaging = nullptr;  // page-locked arena the recorded host-to-device copies read from
    size_t                     recStagingUsed = 0;
    uint64_t                   recLaunches0 = 0, recH2d0 = 0;
    std::vector<RecordedFrame> recorded;
};

// error helpers ------------------------------------------------------------------------------------------------
int fgl_fail(fgl_ctx* c, int code, const std::string& msg);
#define FGL_CUDA(c, call)                                                                                   \
    do                                                                                                      \
    {                                                                                                       \
        cudaError_t e__ = (call);                                                                           \
        if (e__ != cudaSuccess)                                                                             \
            return fgl_fail((c), FGL_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));        \
    } while (0)

int  fgl_reserve(fgl_ctx* c, DevBuf& b, size_t bytes);  // grow-only device allocation
// Steps that synchronise with the host or allocate cannot be part of a recorded frame:
//   if (int rc = fgl_not_while_recording(c, "...")) return rc;
int  fgl_not_while_recording(fgl_ctx* c, const char* what);
bool fgl_time_begin(fgl_ctx* c, const char* name, uint64_t bytes, const unsigned* lateCount, uint64_t lateBytesEach);
void fgl_time_end(fgl_ctx* c);

// RAII bracket used around every kernel launch: counts the launch and, when enabled, records CUDA events.
struct LaunchScope
{
    fgl_ctx* c;
    bool     recording = false;
    LaunchScope(fgl_ctx* ctx, const char* name, uint64_t bytes, const unsigned* lateCount = nullptr, uint64_t lateBytesEach = 0) : c(ctx)
    {
        ++c->launches;
        if (c->timing) recording = fgl_time_begin(c, name, bytes, lateCount, lateBytesEach);
    }
    ~LaunchScope()
    {
        if (recording) fgl_time_end(c);
    }
};

// sort-first group (group.cu)
enum { FGL_GROUP_READY = 0, FGL_GROUP_SHADOW = 1, FGL_GROUP_DEPTH = 2, FGL_GROUP_BAND = 3 };
int  fgl_group_signal(fgl_ctx* c, int what, bool rootOnly = false);  // store this frame's number into the peers' flag slots
int  fgl_group_wait(fgl_ctx* c, int what);                           // device-side wait for every peer's slot (once per frame and kind)
void fgl_group_band(const fgl_ctx* c, int H, int& r0, int& r1);
void fgl_group_release(fgl_ctx* c);

// kernels (raster.cu) ------------------------------------------------------------------------------------------
int fgl_run_raster(fgl_ctx* c, const RasterPass& P, PlanesD planes, uint8_t* rgb8, const LightPass* forwardLight);
int fgl_run_forward_sites(fgl_ctx* c, RasterPass& P, const LightPass& L, size_t* nSites, const float4** sc4);
int fgl_run_resolve_forward(fgl_ctx* c, const RasterPass& P, PlanesD planes, const LightPass& L);
// kernels (shade.cu)
int fgl_run_fill(fgl_ctx* c, float* dst, size_t n, float value);
int fgl_run_fill_rgb(fgl_ctx* c, float* dst, size_t nPixels, const float rgb[3]);
int fgl_run_lighting(fgl_ctx* c, const LightPass& L);
int fgl_run_quantize(fgl_ctx* c, const float* frame, size_t nPixels, uint8_t* rgb8);
int fgl_run_ssaa(fgl_ctx* c, const uint8_t* rgb8, int W, int H, int k, uint8_t* out, int row0, int row1);
int fgl_run_blur(fgl_ctx* c, float* plane, int W, int H, int channels, int kind, int hRow0, int hRow1, int vRow0, int vRow1);
int fgl_run_ids(fgl_ctx* c, const unsigned long long* vis, size_t n, int* out);
int fgl_run_aos(fgl_ctx* c, const float* soa, float* aos, size_t nPixels, int channels, bool toAos);
struct SsaoPass
{
    int          W, H, row0, row1;
    const float *worldpos, *normal, *depth;
    float*       ao;
    float        viewProj[16], viewport[16];
    float        radius, rangeCheckRadius, bias;
    int          rangeCheck;
    int          backgroundIsOne;  // set by fgl_run_ssao: background pixels are exactly AO = 1 (see k_ssao)
    const float4* ball;  // accepted unit-ball samples, 32 per pixel: x, y, z and the sample's scale (ssao_sample_scale)
};
int fgl_run_ssao(fgl_ctx* c, const SsaoPass& S);
