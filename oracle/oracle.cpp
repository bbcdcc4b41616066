// oracle/oracle.cpp — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's per-frame path.
//
// *** This file is the parity checker, not the product.  Only tests/, __graft_entry__.smoke() and bench.py's
// *** cpu_baseline / --impl reference legs may load it.  The product library (forkerrenderer_b200/csrc) never
// *** links, imports or calls anything in oracle/.
//
// It implements the C ABI of include/forkergl_b200.h on the CPU, single-threaded, in the reference's own
// order of evaluation: immediate-mode draws (per face: vertex program x3, bounding-box scan px-outer/py-inner,
// double-precision barycentric, strict-less depth test, fragment program, buffer writes), then the
// sequential SSAO / in-place blur / lighting / SSAA loops, all consuming ONE mt19937(5489) stream.
// Every function cites the reference file:line it restates.  Arithmetic is scalar fp32 with separate
// multiply/add roundings (compile with -ffp-contract=off, no -march, no -ffast-math), doubles where the
// reference uses them.
//
// Parity pin: checked against the UNMODIFIED reference built by oracle/Makefile (oracle/_ref/ref_driver) on
// scenes C1..C4 + PBR — see tests/golden/ (hashes of the reference's raw buffers) and tests/test_oracle_*.py.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../include/forkergl_b200.h"

namespace
{
// ---------------------------------------------------------------------------------------------------------
// small vector helpers with the reference's operation order (reference src/geometry.h)
struct V3
{
    float x, y, z;
};
struct V4
{
    float x, y, z, w;
};
inline V3 v3(float x, float y, float z) { return V3{ x, y, z }; }
inline V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 mul(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 scale(V3 a, float f) { return v3(a.x * f, a.y * f, a.z * f); }
// geometry.h:881-889 — Dot accumulates from 0.f, left to right
inline float dot(V3 a, V3 b)
{
    float r = 0.f;
    r += a.x * b.x;
    r += a.y * b.y;
    r += a.z * b.z;
    return r;
}
inline float dot4(const float* row, V4 v)
{
    float r = 0.f;
    r += row[0] * v.x;
    r += row[1] * v.y;
    r += row[2] * v.z;
    r += row[3] * v.w;
    return r;
}
inline float dot3(const float* row, V3 v)
{
    float r = 0.f;
    r += row[0] * v.x;
    r += row[1] * v.y;
    r += row[2] * v.z;
    return r;
}
// geometry.h:774-782 — matrix x vector = per-row Dot
inline V4 mat4(const float* m, V4 v) { return V4{ dot4(m, v), dot4(m + 4, v), dot4(m + 8, v), dot4(m + 12, v) }; }
inline V3 mat3(const float* m, V3 v) { return v3(dot3(m, v), dot3(m + 3, v), dot3(m + 6, v)); }
// geometry.h:335-341 — vector / scalar = reciprocal, then multiply
inline V3 divs(V3 a, float f)
{
    float inv = 1.f / f;
    return v3(a.x * inv, a.y * inv, a.z * inv);
}
inline V4 divs4(V4 a, float f)
{
    float inv = 1.f / f;
    return V4{ a.x * inv, a.y * inv, a.z * inv, a.w * inv };
}
inline float length(V3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }  // geometry.h:370-371
inline V3    normalize(V3 a) { return divs(scale(a, 1.f), length(a)); }             // geometry.h:906-910
inline V3    cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float clampf(float v, float lo, float hi) { return std::min(hi, std::max(v, lo)); }  // utility.h:32-35
inline int   clampi(int v, int lo, int hi) { return std::min(hi, std::max(v, lo)); }
// geometry.h:912-916 — Lerp(t, a, b) = (1 - t) * a + t * b (scalar-times-vector multiplies component * scalar)
inline V3 lerp3(float t, V3 a, V3 b) { return add(scale(a, 1 - t), scale(b, t)); }

const float kGamma = 2.2;            // constant.h:32
const float kInvGamma = 1.f / 2.2f;  // constant.h:33
const float kInvPi = 0.31830988618379067154;

// ---------------------------------------------------------------------------------------------------------
// mt19937 + Random01 (reference src/utility.h:90-103; libstdc++ generate_canonical<float,24> over one draw)
struct Mt19937
{
    uint32_t s[624];
    int      idx;
    void seed(uint32_t v)
    {
        s[0] = v;
        for (int i = 1; i < 624; ++i) s[i] = 1812433253u * (s[i - 1] ^ (s[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    uint32_t next()
    {
        if (idx >= 624)
        {
            for (int i = 0; i < 624; ++i)
            {
                uint32_t y = (s[i] & 0x80000000u) | (s[(i + 1) % 624] & 0x7fffffffu);
                s[i] = s[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = s[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
};

struct Texture
{
    int                  w = 0, h = 0, bpp = 0, wrap = 0, filter = 0;
    std::vector<uint8_t> data;
};
struct Vertices
{
    std::vector<float> pos, uv, nrm, tan;
};
struct MeshD
{
    int              vertices = -1, nFaces = 0;
    std::vector<int> pi, ti, ni;
    FglMaterial      mat;
    int              hasTangents = 0, supportPBR = 0;
};

struct Plane
{
    int                w = 0, h = 0, ch = 0;
    std::vector<float> v;
    void init(int W, int H, int C, float val)
    {
        w = W, h = H, ch = C;
        v.assign((size_t)W * H * C, val);
    }
};
}  // namespace

struct fgl_ctx
{
    std::string error;
    FglParams   params;
    float       viewport[16], viewProj[16], lightSpace[16];
    int         mode = FGL_MODE_FORWARD, pass = FGL_PASS_FORWARD, shadowOn = 1;
    Plane       planes[FGL_PLANE_AO + 1];
    std::vector<int>     idsCamera, idsLight;
    std::vector<uint8_t> frameRgb8, ssaaRgb8;
    int                  ssaaW = 0, ssaaH = 0;
    int                  primCounter = 0;
    Mt19937              rng;
    uint64_t             draws = 0;
    std::vector<Texture>  textures;
    std::vector<Vertices> vertices;
    std::vector<MeshD>    meshes;
};

static std::string g_createError;

namespace
{
int fail(fgl_ctx* c, int code, const std::string& msg)
{
    if (c) c->error = msg;
    else g_createError = msg;
    return code;
}

void identity(float* m)
{
    memset(m, 0, 16 * sizeof(float));
    m[0] = m[5] = m[10] = m[15] = 1.f;
}

// utility.h:90-98 + libstdc++ generate_canonical: float(u32) * 2^-32, clamped below 1
inline float Random01(fgl_ctx* c)
{
    uint32_t u = c->rng.next();
    ++c->draws;
    float r = (float)u / 4294967296.f;
    if (r >= 1.f) r = std::nextafter(1.f, 0.f);
    return r;
}
inline float Random(fgl_ctx* c, float a, float b) { return a + (b - a) * Random01(c); }  // utility.h:100-103

// geometry.h:952-966.  g++ evaluates constructor arguments right to left: draw0 -> z, draw1 -> y, draw2 -> x
// (measured).
V3 RandomVectorInUnitSphere(fgl_ctx* c)
{
    for (;;)
    {
        float z = Random(c, -1.f, 1.f);
        float y = Random(c, -1.f, 1.f);
        float x = Random(c, -1.f, 1.f);
        if (x * x + y * y + z * z >= 1.f) continue;
        return v3(x, y, z);
    }
}
// geometry.h:968-976: Vector3f(Random(-1,1), Random(-1,1), 0): draw0 -> y, draw1 -> x
V3 RandomVectorInUnitDisk(fgl_ctx* c)
{
    for (;;)
    {
        float y = Random(c, -1, 1);
        float x = Random(c, -1, 1);
        if (x * x + y * y + 0.f * 0.f >= 1.f) continue;
        return v3(x, y, 0.f);
    }
}
// geometry.h:978-990
V3 RandomVectorInHemisphere(fgl_ctx* c, V3 n)
{
    V3 v = RandomVectorInUnitSphere(c);
    return dot(v, n) > 0.f ? v : v3(-v.x, -v.y, -v.z);
}

// ---------------------------------------------------------------------------------------------------------
// Texture sampling (reference src/texture.h:41-145, tgaimage.cpp:304-311)
V3 texel(const Texture& t, int x, int y)  // getColorFromImage, [0,255]; OOB = black; grey value lives in b
{
    if (t.data.empty() || x < 0 || y < 0 || x >= t.w || y >= t.h) return v3(0, 0, 0);
    const uint8_t* p = &t.data[((size_t)x + (size_t)y * t.w) * t.bpp];
    uint8_t        bgra[4] = { 0, 0, 0, 0 };
    for (int i = 0; i < t.bpp; ++i) bgra[i] = p[i];
    return v3((float)bgra[2], (float)bgra[1], (float)bgra[0]);
}

void wrapCoord(const Texture& t, float& u, float& v)  // texture.h:62-83
{
    if (t.wrap == FGL_WRAP_REPEAT)
    {
        u = u - std::floor(u);
        v = v - std::floor(v);
    }
    else if (t.wrap == FGL_WRAP_MIRRORED_REPEAT)
    {
        int   xi = std::floor(u), yi = std::floor(v);
        float rx = u - xi, ry = v - yi;
        u = xi % 2 == 0 ? rx : 1.f - rx;
        v = yi % 2 == 0 ? ry : 1.f - ry;
    }
    else if (t.wrap == FGL_WRAP_CLAMP_TO_EDGE)
    {
        u = clampf(u, 0.f, 1.f);
        v = clampf(v, 0.f, 1.f);
    }
}

V3 colorFromFiltering(const Texture& t, float u, float v)  // texture.h:86-132
{
    float w = t.w - 0.001, h = t.h - 0.001;
    if (t.filter == FGL_FILTER_LINEAR)
    {
        float px = u * w, py = v * h;
        float tlx = std::floor(px - 0.5f), tly = std::floor(py - 0.5f);
        float tx = px - (tlx + 0.5f), ty = py - (tly + 0.5f);
        int   x0 = tlx, y0 = tly, x1 = tlx + 1.f, y1 = tly + 1.f;
        int   sx[4] = { x0, x1, x0, x1 }, sy[4] = { y0, y0, y1, y1 };
        if (t.wrap != FGL_WRAP_NOWRAP)
            for (int i = 0; i < 4; ++i) sx[i] = clampi(sx[i], 0, t.w - 1), sy[i] = clampi(sy[i], 0, t.h - 1);
        V3 c0 = texel(t, sx[0], sy[0]), c1 = texel(t, sx[1], sy[1]), c2 = texel(t, sx[2], sy[2]),
           c3 = texel(t, sx[3], sy[3]);
        V3 cx1 = lerp3(tx, c0, c1), cx2 = lerp3(tx, c2, c3);
        return lerp3(ty, cx1, cx2);
    }
    int ix = std::floor(u * w), iy = std::floor(v * h);
    return texel(t, ix, iy);
}

V3 Sample(const Texture& t, float u, float v)  // texture.h:41-45 (vector / 255.f = reciprocal-multiply)
{
    wrapCoord(t, u, v);
    return divs(colorFromFiltering(t, u, v), 255.f);
}
float SampleFloat(const Texture& t, float u, float v)  // texture.h:47-51 (true division of the b channel)
{
    wrapCoord(t, u, v);
    return colorFromFiltering(t, u, v).z / 255.f;
}

// ---------------------------------------------------------------------------------------------------------
// Shadow filters (reference src/shaders/shadow.cpp:23-132)
float SampleShadowMap(fgl_ctx* c, float u, float v)  // shadow.cpp:23-35
{
    const Plane& sm = c->planes[FGL_PLANE_SHADOW];
    if (u < 0.f || u > 1.f || v < 0.f || v > 1.f) return std::numeric_limits<float>::infinity();
    int   w = sm.w - 0.001f;
    int   h = sm.h - 0.001f;
    int   iu = (int)((float)w * u);
    int   iv = (int)((float)h * v);
    float depth = sm.v[(size_t)iu + (size_t)iv * sm.w];
    return depth < 0.001 ? 1.f : depth;
}

float PCF(fgl_ctx* c, V3 sc, float bias, float filterSize)  // shadow.cpp:47-63
{
    float visibility = 0.f;
    float inv = 1.f / (float)64;
    for (int i = 0; i < 64; ++i)
    {
        V3    d = RandomVectorInUnitDisk(c);
        float u = sc.x + d.x * filterSize, v = sc.y + d.y * filterSize;
        float sampleDepth = SampleShadowMap(c, u, v);
        if (sc.z <= sampleDepth + bias) visibility += inv;
    }
    return visibility;
}

float FindAverageBlockDepth(fgl_ctx* c, V3 sc, float bias)  // shadow.cpp:65-90
{
    float  blockerDepth = 0.f, numBlockers = 0.f;
    double fs = c->params.pcss_blocker_filter_size;  // a double literal in the reference: vector * double
    for (int i = 0; i < 32; ++i)
    {
        V3    d = RandomVectorInUnitDisk(c);
        float ox = (float)(d.x * fs), oy = (float)(d.y * fs);  // Vector2f * double: product in double, stored as float
        float sampleDepth = SampleShadowMap(c, sc.x + ox, sc.y + oy);
        if (sc.z > sampleDepth + bias)
        {
            blockerDepth += sampleDepth;
            numBlockers += 1.f;
        }
    }
    if (numBlockers < 1.f) return 0.f;
    return blockerDepth / numBlockers;
}

float PCSS(fgl_ctx* c, V3 sc, float bias)  // shadow.cpp:92-106
{
    float dReceiver = sc.z;
    float dBlocker = FindAverageBlockDepth(c, sc, bias);
    if (dBlocker < 0.001) return 1.f;
    float penumbra = (dReceiver - dBlocker) * c->params.area_light_size / dBlocker;
    float filterSize = c->params.pcf_filter_size * penumbra;  // double * float -> float
    return PCF(c, sc, bias, filterSize);
}

float CalculateShadowVisibility(fgl_ctx* c, V3 ndc, V3 n, V3 l)  // shadow.cpp:109-132
{
    V3    sc = add(scale(ndc, 0.5f), v3(0.5f, 0.5f, 0.5f));
    float bias = std::max(c->params.shadow_bias_slope * (1.f - dot(n, l)), c->params.shadow_bias_min);
    if (c->params.shadow_mode == FGL_SHADOW_PCF) return PCF(c, sc, bias, (float)c->params.pcf_filter_size);
    if (c->params.shadow_mode == FGL_SHADOW_PCSS) return PCSS(c, sc, bias);
    float sampled = SampleShadowMap(c, sc.x, sc.y);  // HardShadow, shadow.cpp:38-45
    return (sc.z <= sampled + bias) ? 1.f : 0.f;
}

// ---------------------------------------------------------------------------------------------------------
// Lighting models
V3 pow3(V3 v, float p) { return v3(std::pow(v.x, p), std::pow(v.y, p), std::pow(v.z, p)); }
V3 clamp01(V3 v) { return v3(clampf(v.x, 0.f, 1.f), clampf(v.y, 0.f, 1.f), clampf(v.z, 0.f, 1.f)); }

// reference src/shaders/phongshader.h:171-215
V3 BlinnPhongLight(fgl_ctx* c, V3 lightDir, V3 halfwayDir, V3 normal, float visibility, V3 diffuseColor,
                   V3 emissive, V3 param, V3 lightColor)
{
    V3    dl = pow3(diffuseColor, kGamma), el = pow3(emissive, kGamma);
    float ao = param.x, ks = param.y, shininess = param.z;
    float diff = std::max(0.f, dot(lightDir, normal));
    float spec = std::pow(std::max(0.f, dot(halfwayDir, normal)), shininess);
    V3    ambient = scale(mul(v3(0.3f, 0.3f, 0.3f), dl), ao);
    V3    diffuse = scale(scale(dl, diff), ao);
    V3    specular = scale(v3(ks, ks, ks), spec);
    if (c->shadowOn)
    {
        float shadow = (1 - visibility) * c->params.shadow_intensity;
        visibility = 1 - shadow;
        diffuse = scale(diffuse, visibility);
        specular = scale(specular, visibility);
    }
    V3 color = add(ambient, mul(add(add(diffuse, specular), el), lightColor));
    V3 den = add(color, v3(1.f, 1.f, 1.f));
    color = v3(color.x / den.x, color.y / den.y, color.z / den.z);  // Vector / Vector: true division
    color = pow3(color, kInvGamma);
    return clamp01(color);
}

// reference src/shaders/pbrshader.h:182-288
V3 PBRLight(fgl_ctx* c, V3 lightDir, V3 viewDir, V3 halfwayDir, V3 normal, float visibility, V3 albedo,
            V3 emissive, V3 param, V3 lightRadiance)
{
    V3    al = pow3(albedo, kGamma), el = pow3(emissive, kGamma);
    float ao = param.x, metalness = param.y, roughness = param.z;
    float NdotV = std::max(dot(normal, viewDir), 0.f);
    float NdotL = std::max(dot(normal, lightDir), 0.f);
    float NdotH = std::max(dot(normal, halfwayDir), 0.f);
    float HdotV = std::max(dot(halfwayDir, viewDir), 0.f);
    V3    F0 = lerp3(metalness, v3(0.04f, 0.04f, 0.04f), al);
    // distributionGGX pbrshader.h:256-266
    float a = roughness * roughness, a2 = a * a, NdotH2 = NdotH * NdotH;
    float den = (NdotH2 * (a2 - 1.f) + 1.f);
    float NDF = a2 * kInvPi / (den * den);
    // geometrySmith pbrshader.h:268-282
    float ka = roughness + 1.f, k = ka * ka / 8.f;
    float ggx1 = NdotV / (NdotV * (1 - k) + k);
    float ggx2 = NdotL / (NdotL * (1 - k) + k);
    float G = ggx1 * ggx2;
    // fresnelSchlick pbrshader.h:284-288
    float om = std::max(1.f - HdotV, 0.f);
    float p5 = std::pow(om, 5.f);
    V3    F = add(F0, scale(sub(v3(1.f, 1.f, 1.f), F0), p5));
    V3    DGF = scale(F, NDF * G);  // "NDF * G * F" = (NDF*G) * F -> F * scalar
    float denominator = 4 * NdotV * NdotL + 0.001f;
    V3    specular = divs(DGF, denominator);
    V3    kd = sub(v3(1.f, 1.f, 1.f), F);
    kd = scale(kd, 1.f - metalness);
    V3 brdf = add(scale(mul(kd, al), kInvPi), specular);
    V3 Lo = scale(mul(brdf, lightRadiance), NdotL);
    if (c->shadowOn)
    {
        float shadow = (1 - visibility) * c->params.shadow_intensity;
        visibility = 1 - shadow;
        Lo = scale(Lo, visibility);
    }
    V3 color = Lo;
    color = add(color, scale(mul(v3(0.3f, 0.3f, 0.3f), al), ao));  // Color3(0.3): double literal -> float 0.3f
    color = add(color, el);
    V3 d = add(color, v3(1.f, 1.f, 1.f));
    color = v3(color.x / d.x, color.y / d.y, color.z / d.z);
    color = pow3(color, kInvGamma);
    return clamp01(color);
}

// ---------------------------------------------------------------------------------------------------------
// The four programs.  Varyings of one triangle, as the reference keeps them in the Shader object.
struct Varyings
{
    V4    ndc[3];
    V3    posWS[3], nrmWS[3], tanWS[3], lightNDC[3];  // all already multiplied by 1/w_clip (PCI)
    float u[3], v[3], oow[3];
    V3    depthNdc[3];  // DepthShader: vPositionNDC columns
};

struct Draw
{
    const MeshD*    mesh;
    const Vertices* vb;
    FglUniforms     un;
    int             kind;
    float           lm[16];  // DepthShader: uLightSpaceMatrix * uModelMatrix
};

V3 fetch3(const std::vector<float>& a, int i) { return v3(a[(size_t)i * 3], a[(size_t)i * 3 + 1], a[(size_t)i * 3 + 2]); }

// depthshader.h:21-28 / gshader.h:41-92 (== phongshader.h:35-85 == pbrshader.h:35-85)
void ProcessVertex(fgl_ctx* c, const Draw& d, int face, int k, Varyings& o)
{
    const MeshD& m = *d.mesh;
    V3           p = fetch3(d.vb->pos, m.pi[face * 3 + k]);
    if (d.kind == FGL_SHADER_DEPTH)
    {
        V4 cs = mat4(d.lm, V4{ p.x, p.y, p.z, 1.f });
        V4 ndc = divs4(cs, cs.w);
        o.depthNdc[k] = v3(ndc.x, ndc.y, ndc.z);
        o.ndc[k] = ndc;
        return;
    }
    V4    ws = mat4(d.un.model, V4{ p.x, p.y, p.z, 1.f });
    V4    vs = mat4(d.un.view, ws);
    V4    cs = mat4(d.un.projection, vs);
    int   ti = m.ti[face * 3 + k];
    float tu = d.vb->uv[(size_t)ti * 2], tv = d.vb->uv[(size_t)ti * 2 + 1];
    V3    nWS = mat3(d.un.normal, normalize(fetch3(d.vb->nrm, m.ni[face * 3 + k])));  // mesh.cpp:46-50
    V3    tWS = v3(0, 0, 0);
    if (m.hasTangents) tWS = mat3(d.un.normal, normalize(fetch3(d.vb->tan, m.pi[face * 3 + k])));
    V4 ls = V4{ 0, 0, 0, 0 };
    if (c->shadowOn)
    {
        ls = mat4(d.un.light_space, ws);
        ls = divs4(ls, ls.w);
    }
    float oow = 1.f / cs.w;
    o.oow[k] = oow;
    o.posWS[k] = v3(ws.x * oow, ws.y * oow, ws.z * oow);
    o.u[k] = tu * oow, o.v[k] = tv * oow;
    o.nrmWS[k] = scale(nWS, oow);
    if (m.hasTangents) o.tanWS[k] = scale(tWS, oow);
    if (c->shadowOn) o.lightNDC[k] = v3(ls.x * oow, ls.y * oow, ls.z * oow);
    o.ndc[k] = divs4(cs, cs.w);
}

inline float interp(float a0, float a1, float a2, V3 b)  // Matrix row . bary (geometry.h:774-782)
{
    float r = 0.f;
    r += a0 * b.x;
    r += a1 * b.y;
    r += a2 * b.z;
    return r;
}
inline V3 interp3(const V3* a, V3 b)
{
    return v3(interp(a[0].x, a[1].x, a[2].x, b), interp(a[0].y, a[1].y, a[2].y, b), interp(a[0].z, a[1].z, a[2].z, b));
}

struct Surface
{
    V3    posWS, normal, lightNDC;
    float u, v;
};

// common head of the three camera-space fragment programs: gshader.h:95-146 == phongshader.h:90-128
Surface InterpolateSurface(fgl_ctx* c, const Draw& d, const Varyings& vy, V3 bary, bool hasNormalMap, int normalMapId)
{
    Surface s;
    V3      pos = interp3(vy.posWS, bary);
    float   tu = interp(vy.u[0], vy.u[1], vy.u[2], bary), tv = interp(vy.v[0], vy.v[1], vy.v[2], bary);
    V3      nrm = interp3(vy.nrmWS, bary);
    float   w = 1.f / dot(v3(vy.oow[0], vy.oow[1], vy.oow[2]), bary);
    pos = scale(pos, w);
    tu *= w, tv *= w;
    nrm = scale(nrm, w);
    V3 N = normalize(nrm);
    V3 normal = N;
    if (d.mesh->hasTangents && hasNormalMap)
    {
        V3 tg = scale(interp3(vy.tanWS, bary), w);
        V3 T = normalize(add(tg, v3(0.001f, 0.001f, 0.001f)));
        T = normalize(sub(T, scale(N, dot(T, N))));
        V3 B = normalize(cross(N, T));
        V3 sn = Sample(c->textures[normalMapId], tu, tv);
        sn = normalize(sub(scale(sn, 2.f), v3(1.f, 1.f, 1.f)));
        // TbnMatrix columns T, B, N; rows dotted with sn
        normal = normalize(v3(dot(v3(T.x, B.x, N.x), sn), dot(v3(T.y, B.y, N.y), sn), dot(v3(T.z, B.z, N.z), sn)));
    }
    s.posWS = pos, s.normal = normal, s.u = tu, s.v = tv;
    s.lightNDC = v3(0, 0, 0);
    if (c->shadowOn) s.lightNDC = scale(interp3(vy.lightNDC, bary), w);
    return s;
}

struct GOut
{
    V3    normal, pos, lightNDC, albedo, emissive, param;
    float type;
};

// gshader.h:95-201
GOut GFragment(fgl_ctx* c, const Draw& d, const Varyings& vy, V3 bary)
{
    const FglMaterial& m = d.mesh->mat;
    Surface            s = InterpolateSurface(c, d, vy, bary, m.normal_map >= 0, m.normal_map);
    GOut               o;
    o.normal = s.normal, o.pos = s.posWS, o.lightNDC = s.lightNDC;
    auto tex = [&](int id) -> const Texture& { return c->textures[id]; };
    if (d.mesh->supportPBR)
    {
        o.albedo = m.base_color_map >= 0 ? Sample(tex(m.base_color_map), s.u, s.v) : v3(m.albedo[0], m.albedo[1], m.albedo[2]);
        o.emissive = m.pbr_emissive_map >= 0 ? Sample(tex(m.pbr_emissive_map), s.u, s.v) : v3(m.ke[0], m.ke[1], m.ke[2]);
        float roughness = m.roughness_map >= 0 ? SampleFloat(tex(m.roughness_map), s.u, s.v) : m.roughness;
        float metalness = m.metalness_map >= 0 ? SampleFloat(tex(m.metalness_map), s.u, s.v) : m.metalness;
        float ao = m.ao_map >= 0 ? SampleFloat(tex(m.ao_map), s.u, s.v) : 1.f;
        o.param = v3(ao, metalness, roughness);
        o.type = 1.f;
    }
    else
    {
        o.emissive = m.emissive_map >= 0 ? Sample(tex(m.emissive_map), s.u, s.v) : v3(m.ke[0], m.ke[1], m.ke[2]);
        o.albedo = m.diffuse_map >= 0 ? Sample(tex(m.diffuse_map), s.u, s.v) : v3(m.kd[0], m.kd[1], m.kd[2]);
        float shininess = m.specular_map >= 0 ? SampleFloat(tex(m.specular_map), s.u, s.v) + 5 : 1.f;
        o.param = v3(1.f, m.ks[0], shininess);
        o.type = 0.f;
    }
    return o;
}

// phongshader.h:90-169
V3 BlinnPhongFragment(fgl_ctx* c, const Draw& d, const Varyings& vy, V3 bary)
{
    const FglMaterial& m = d.mesh->mat;
    Surface            s = InterpolateSurface(c, d, vy, bary, m.normal_map >= 0, m.normal_map);
    V3 lp = v3(d.un.light_position[0], d.un.light_position[1], d.un.light_position[2]);
    V3 ep = v3(d.un.eye_position[0], d.un.eye_position[1], d.un.eye_position[2]);
    V3 lightDir = normalize(sub(lp, s.posWS));
    V3 viewDir = normalize(sub(ep, s.posWS));
    V3 halfwayDir = normalize(add(lightDir, viewDir));
    float visibility = 0.f;
    if (c->shadowOn) visibility = CalculateShadowVisibility(c, s.lightNDC, s.normal, lightDir);
    auto tex = [&](int id) -> const Texture& { return c->textures[id]; };
    V3    diffuseColor = m.diffuse_map >= 0 ? Sample(tex(m.diffuse_map), s.u, s.v) : v3(m.kd[0], m.kd[1], m.kd[2]);
    V3    emissive = m.emissive_map >= 0 ? Sample(tex(m.emissive_map), s.u, s.v) : v3(m.ke[0], m.ke[1], m.ke[2]);
    float shininess = 1.f;
    if (m.specular_map >= 0) shininess = SampleFloat(tex(m.specular_map), s.u, s.v) + 5;
    V3 param = v3(m.ka[0], m.ks[0], shininess);
    return BlinnPhongLight(c, lightDir, halfwayDir, s.normal, visibility, diffuseColor, emissive, param,
                           v3(d.un.light_color[0], d.un.light_color[1], d.un.light_color[2]));
}

// pbrshader.h:90-180
V3 PBRFragment(fgl_ctx* c, const Draw& d, const Varyings& vy, V3 bary)
{
    const FglMaterial& m = d.mesh->mat;
    Surface            s = InterpolateSurface(c, d, vy, bary, m.pbr_normal_map >= 0, m.pbr_normal_map);
    V3 lp = v3(d.un.light_position[0], d.un.light_position[1], d.un.light_position[2]);
    V3 ep = v3(d.un.eye_position[0], d.un.eye_position[1], d.un.eye_position[2]);
    V3 lightDir = normalize(sub(lp, s.posWS));
    V3 viewDir = normalize(sub(ep, s.posWS));
    V3 halfwayDir = normalize(add(lightDir, viewDir));
    float visibility = 0.f;
    if (c->shadowOn) visibility = CalculateShadowVisibility(c, s.lightNDC, s.normal, lightDir);
    auto tex = [&](int id) -> const Texture& { return c->textures[id]; };
    V3 albedo = m.base_color_map >= 0 ? Sample(tex(m.base_color_map), s.u, s.v) : v3(m.albedo[0], m.albedo[1], m.albedo[2]);
    V3 emissive = m.pbr_emissive_map >= 0 ? Sample(tex(m.pbr_emissive_map), s.u, s.v) : v3(m.pbr_ke[0], m.pbr_ke[1], m.pbr_ke[2]);
    float roughness = m.roughness_map >= 0 ? SampleFloat(tex(m.roughness_map), s.u, s.v) : m.roughness;
    float metalness = m.metalness_map >= 0 ? SampleFloat(tex(m.metalness_map), s.u, s.v) : m.metalness;
    float ao = m.ao_map >= 0 ? SampleFloat(tex(m.ao_map), s.u, s.v) : 1.f;
    return PBRLight(c, lightDir, viewDir, halfwayDir, s.normal, visibility, albedo, emissive, v3(ao, metalness, roughness),
                    v3(d.un.light_color[0], d.un.light_color[1], d.un.light_color[2]));
}

// geometry.cpp:20-56 — double-precision barycentric on integer-snapped vertices; returns false if outside
bool Barycentric(const int X[3], const int Y[3], int px, int py, V3& out)
{
    float  ax = X[0], ay = Y[0], bx = X[1], by = Y[1], cx = X[2], cy = Y[2], fx = px, fy = py;
    double s0x = bx - ax, s0y = cx - ax, s0z = ax - fx;  // float subtraction, then widened
    double s1x = by - ay, s1y = cy - ay, s1z = ay - fy;
    double rx = s0y * s1z - s0z * s1y, ry = s0z * s1x - s0x * s1z, rz = s0x * s1y - s0y * s1x;
    if (std::abs(rz) > 1e-2)
    {
        double inv = 1.f / rz;
        rx *= inv;
        ry *= inv;
        float r0 = (float)(1.f - (rx + ry)), r1 = (float)rx, r2 = (float)ry;
        if (r0 < 0.f || r1 < 0.f || r2 < 0.f) return false;
        out = v3(r0, r1, r2);
        return true;
    }
    return false;
}

void set3(Plane& p, int x, int y, V3 v)
{
    float* q = &p.v[((size_t)x + (size_t)y * p.w) * 3];
    q[0] = v.x, q[1] = v.y, q[2] = v.z;
}
V3 get3(const Plane& p, int x, int y)
{
    const float* q = &p.v[((size_t)x + (size_t)y * p.w) * 3];
    return v3(q[0], q[1], q[2]);
}

// forkergl.cpp:239-324 (DrawTriangle) + :165-234 (DrawTriangleSubTask)
void DrawTriangle(fgl_ctx* c, const Draw& d, const Varyings& vy, int primId)
{
    int   X[3], Y[3];
    float depths[3];
    for (int i = 0; i < 3; ++i)
    {
        V4 s = mat4(c->viewport, vy.ndc[i]);
        X[i] = (int)s.x, Y[i] = (int)s.y;  // C truncation (x86 cvttss2si)
        depths[i] = s.z;
    }
    bool   shadowPass = c->pass == FGL_PASS_SHADOW;
    Plane& depth = c->planes[FGL_PLANE_DEPTH];
    int    w = shadowPass ? c->planes[FGL_PLANE_SHADOW].w : depth.w;
    int    h = shadowPass ? c->planes[FGL_PLANE_SHADOW].h : depth.h;
    int    xMin = clampi(std::min(X[0], std::min(X[1], X[2])), 0, w - 1), yMin = clampi(std::min(Y[0], std::min(Y[1], Y[2])), 0, h - 1);
    int    xMax = clampi(std::max(X[0], std::max(X[1], X[2])), 0, w - 1), yMax = clampi(std::max(Y[0], std::max(Y[1], Y[2])), 0, h - 1);
    std::vector<int>& ids = shadowPass ? c->idsLight : c->idsCamera;
    for (int px = xMin; px <= xMax; ++px)
        for (int py = yMin; py <= yMax; ++py)
        {
            V3 bary;
            if (!Barycentric(X, Y, px, py, bary)) continue;
            float  z = dot(bary, v3(depths[0], depths[1], depths[2]));
            size_t idx = (size_t)px + (size_t)py * w;
            if (c->pass != FGL_PASS_LIGHTING)
            {
                if (z >= depth.v[idx]) continue;
                depth.v[idx] = z;
                ids[idx] = primId;
            }
            if (c->pass == FGL_PASS_SHADOW)
            {
                float nz = interp(vy.depthNdc[0].z, vy.depthNdc[1].z, vy.depthNdc[2].z, bary);  // depthshader.h:30-36
                c->planes[FGL_PLANE_SHADOW].v[idx] = nz * 0.5f + 0.5f;
            }
            else if (c->pass == FGL_PASS_GEOMETRY)
            {
                GOut o = GFragment(c, d, vy, bary);
                set3(c->planes[FGL_PLANE_NORMAL], px, py, o.normal);
                set3(c->planes[FGL_PLANE_WORLDPOS], px, py, o.pos);
                if (c->shadowOn) set3(c->planes[FGL_PLANE_LIGHTNDC], px, py, o.lightNDC);
                set3(c->planes[FGL_PLANE_ALBEDO], px, py, o.albedo);
                set3(c->planes[FGL_PLANE_EMISSIVE], px, py, o.emissive);
                set3(c->planes[FGL_PLANE_PARAM], px, py, o.param);
                c->planes[FGL_PLANE_SHADINGTYPE].v[idx] = o.type;
            }
            else if (c->pass == FGL_PASS_FORWARD)
            {
                V3 col = d.kind == FGL_SHADER_PBR ? PBRFragment(c, d, vy, bary) : BlinnPhongFragment(c, d, vy, bary);
                set3(c->planes[FGL_PLANE_FRAME], px, py, col);
            }
        }
}

bool okPlane(int p) { return p >= 0 && p < FGL_PLANE_COUNT; }
}  // namespace

// =========================================================================================================
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

int fgl_create(int, fgl_ctx** out)
{
    if (!out) return fail(nullptr, FGL_ERR_INVALID, "out_ctx is NULL");
    fgl_ctx* c = new fgl_ctx();
    fgl_default_params(&c->params);
    identity(c->viewport), identity(c->viewProj), identity(c->lightSpace);
    c->rng.seed(5489u);
    *out = c;
    return FGL_OK;
}
void        fgl_destroy(fgl_ctx* c) { delete c; }
const char* fgl_last_error(fgl_ctx* c) { return c ? c->error.c_str() : g_createError.c_str(); }
const char* fgl_backend_name(void) { return "oracle-cpu"; }
int         fgl_set_stream(fgl_ctx*, void*) { return FGL_OK; }
int         fgl_set_params(fgl_ctx* c, const FglParams* p)
{
    if (!c || !p) return FGL_ERR_INVALID;
    c->params = *p;
    return FGL_OK;
}
int fgl_sync(fgl_ctx*) { return FGL_OK; }

int fgl_upload_texture(fgl_ctx* c, const uint8_t* texels, int w, int h, int bpp, int wrap, int filter, int* id)
{
    if (!c || !texels || w <= 0 || h <= 0 || (bpp != 1 && bpp != 3 && bpp != 4) || !id)
        return fail(c, FGL_ERR_INVALID, "fgl_upload_texture: bad arguments");
    Texture t;
    t.w = w, t.h = h, t.bpp = bpp, t.wrap = wrap, t.filter = filter;
    t.data.assign(texels, texels + (size_t)w * h * bpp);
    c->textures.push_back(std::move(t));
    *id = (int)c->textures.size() - 1;
    return FGL_OK;
}

int fgl_upload_vertices(fgl_ctx* c, const float* pos, int np, const float* uv, int nt, const float* nrm, int nn,
                        const float* tan, int ntan, int* id)
{
    if (!c || !id || np < 0 || nt < 0 || nn < 0) return fail(c, FGL_ERR_INVALID, "fgl_upload_vertices: bad arguments");
    Vertices v;
    if (pos) v.pos.assign(pos, pos + (size_t)np * 3);
    if (uv) v.uv.assign(uv, uv + (size_t)nt * 2);
    if (nrm) v.nrm.assign(nrm, nrm + (size_t)nn * 3);
    if (tan) v.tan.assign(tan, tan + (size_t)ntan * 3);
    c->vertices.push_back(std::move(v));
    *id = (int)c->vertices.size() - 1;
    return FGL_OK;
}

int fgl_upload_mesh(fgl_ctx* c, int vid, int nFaces, const int* pi, const int* ti, const int* ni, const FglMaterial* mat,
                    int hasTangents, int supportPBR, int* id)
{
    if (!c || !id || !mat || vid < 0 || vid >= (int)c->vertices.size() || nFaces < 0)
        return fail(c, FGL_ERR_INVALID, "fgl_upload_mesh: bad arguments");
    const Vertices& vb = c->vertices[vid];
    if (hasTangents && vb.tan.empty()) return fail(c, FGL_ERR_INVALID, "fgl_upload_mesh: has_tangents without tangents");
    MeshD m;
    m.vertices = vid, m.nFaces = nFaces;
    m.pi.assign(pi, pi + (size_t)nFaces * 3);
    m.ti.assign(ti, ti + (size_t)nFaces * 3);
    m.ni.assign(ni, ni + (size_t)nFaces * 3);
    for (int i = 0; i < nFaces * 3; ++i)
        if (m.pi[i] < 0 || (size_t)m.pi[i] * 3 >= vb.pos.size() || m.ti[i] < 0 || (size_t)m.ti[i] * 2 >= vb.uv.size() ||
            m.ni[i] < 0 || (size_t)m.ni[i] * 3 >= vb.nrm.size())
            return fail(c, FGL_ERR_INVALID, "fgl_upload_mesh: index out of range");
    m.mat = *mat;
    m.hasTangents = hasTangents, m.supportPBR = supportPBR;
    c->meshes.push_back(std::move(m));
    *id = (int)c->meshes.size() - 1;
    return FGL_OK;
}

// ---- state ------------------------------------------------------------------------------------------------
int fgl_init_frame_buffer(fgl_ctx* c, int w, int h)  // forkergl.cpp:55-58
{
    c->planes[FGL_PLANE_FRAME].init(w, h, 3, 0.f);
    return FGL_OK;
}
int fgl_init_depth_buffer(fgl_ctx* c, int w, int h)  // forkergl.cpp:60-63
{
    c->planes[FGL_PLANE_DEPTH].init(w, h, 1, std::numeric_limits<float>::max());
    return FGL_OK;
}
int fgl_init_shadow_buffer(fgl_ctx* c, int w, int h)  // forkergl.cpp:65-68
{
    c->planes[FGL_PLANE_SHADOW].init(w, h, 1, 0.f);
    return FGL_OK;
}
int fgl_init_geometry_buffers(fgl_ctx* c, int w, int h)  // forkergl.cpp:70-81
{
    c->planes[FGL_PLANE_NORMAL].init(w, h, 3, 0.f);
    c->planes[FGL_PLANE_WORLDPOS].init(w, h, 3, 0.f);
    if (c->shadowOn) c->planes[FGL_PLANE_LIGHTNDC].init(w, h, 3, 0.f);
    c->planes[FGL_PLANE_ALBEDO].init(w, h, 3, 0.f);
    c->planes[FGL_PLANE_EMISSIVE].init(w, h, 3, 0.f);
    c->planes[FGL_PLANE_PARAM].init(w, h, 3, 0.f);
    c->planes[FGL_PLANE_SHADINGTYPE].init(w, h, 1, 0.f);
    c->planes[FGL_PLANE_AO].init(w, h, 1, 1.f);
    return FGL_OK;
}
int fgl_clear_color(fgl_ctx* c, const float rgb[3])  // forkergl.cpp:84-87, buffer.cpp:128-137
{
    Plane& f = c->planes[FGL_PLANE_FRAME];
    for (size_t i = 0; i < (size_t)f.w * f.h; ++i) f.v[i * 3] = rgb[0], f.v[i * 3 + 1] = rgb[1], f.v[i * 3 + 2] = rgb[2];
    return FGL_OK;
}
int fgl_set_viewport(fgl_ctx* c, int x, int y, int w, int h)  // forkergl.cpp:89-102
{
    identity(c->viewport);
    c->viewport[0] = w / 2.f;
    c->viewport[5] = h / 2.f;
    c->viewport[3] = x + w / 2.f;
    c->viewport[7] = y + h / 2.f;
    c->viewport[10] = 1 / 2.f;
    c->viewport[11] = 1 / 2.f;
    return FGL_OK;
}
int fgl_get_viewport_matrix(fgl_ctx* c, float o[16]) { memcpy(o, c->viewport, 64); return FGL_OK; }
int fgl_set_view_projection_matrix(fgl_ctx* c, const float m[16]) { memcpy(c->viewProj, m, 64); return FGL_OK; }
int fgl_get_view_projection_matrix(fgl_ctx* c, float o[16]) { memcpy(o, c->viewProj, 64); return FGL_OK; }
int fgl_set_light_space_matrix(fgl_ctx* c, const float m[16]) { memcpy(c->lightSpace, m, 64); return FGL_OK; }
int fgl_get_light_space_matrix(fgl_ctx* c, float o[16]) { memcpy(o, c->lightSpace, 64); return FGL_OK; }
int fgl_set_render_mode(fgl_ctx* c, int mode) { c->mode = mode; return FGL_OK; }
int fgl_get_render_mode(fgl_ctx* c, int* mode) { *mode = c->mode; return FGL_OK; }
int fgl_set_pass_type(fgl_ctx* c, int pass)
{
    c->pass = pass;
    c->primCounter = 0;
    const Plane& d = c->planes[pass == FGL_PASS_SHADOW ? FGL_PLANE_SHADOW : FGL_PLANE_DEPTH];
    if (pass == FGL_PASS_SHADOW) c->idsLight.assign((size_t)d.w * d.h, -1);
    else if (pass != FGL_PASS_LIGHTING) c->idsCamera.assign((size_t)d.w * d.h, -1);
    return FGL_OK;
}
int fgl_set_shadow_status(fgl_ctx* c, int on) { c->shadowOn = on ? 1 : 0; return FGL_OK; }
int fgl_begin_frame(fgl_ctx* c)
{
    c->rng.seed(5489u);
    c->draws = 0;
    return FGL_OK;
}
int fgl_set_row_band(fgl_ctx*, int, int) { return FGL_OK; }  // the oracle always computes the whole frame
int fgl_set_chain_blockers_before(fgl_ctx*, uint64_t) { return FGL_OK; }
int fgl_get_chain_blockers(fgl_ctx* c, uint64_t*) { return fail(c, FGL_ERR_UNSUPPORTED, "the CPU oracle renders whole frames"); }

// ---- draws --------------------------------------------------------------------------------------------------
int fgl_draw_mesh(fgl_ctx* c, int meshId, int kind, const FglUniforms* un)  // mesh.cpp:10-25
{
    if (!c || !un || meshId < 0 || meshId >= (int)c->meshes.size()) return fail(c, FGL_ERR_INVALID, "fgl_draw_mesh: bad mesh");
    if (c->pass == FGL_PASS_GEOMETRY && kind != FGL_SHADER_G)
        return fail(c, FGL_ERR_STATE, "geometry pass requires GShader (reference forkergl.cpp:214 dynamic_cast)");
    Draw d;
    d.mesh = &c->meshes[meshId];
    d.vb = &c->vertices[d.mesh->vertices];
    d.un = *un;
    d.kind = kind;
    for (int i = 0; i < 4; ++i)  // uLightSpaceMatrix * uModelMatrix (depthshader.h:23-24), Dot(row, col)
        for (int j = 0; j < 4; ++j)
        {
            float r = 0.f;
            for (int k = 0; k < 4; ++k) r += un->light_space[i * 4 + k] * un->model[k * 4 + j];
            d.lm[i * 4 + j] = r;
        }
    for (int f = 0; f < d.mesh->nFaces; ++f)
    {
        Varyings vy;
        memset(&vy, 0, sizeof vy);
        for (int k = 0; k < 3; ++k) ProcessVertex(c, d, f, k, vy);
        DrawTriangle(c, d, vy, c->primCounter++);
    }
    return FGL_OK;
}

// ForkerGL::DrawTriangle (forkergl.cpp:239-324) for triangles whose vertex programs the caller ran: same rasteriser and
// fragment programs as above, the varyings come from the caller's arrays (layout: include/forkergl_b200.h)
int fgl_draw_triangles(fgl_ctx* c, int meshId, int kind, const FglUniforms* un, int n, const float* ndc, const float* vary, const float* lightZ)
{
    if (!c || !un || meshId < 0 || meshId >= (int)c->meshes.size() || n < 0 || (n && !ndc)) return fail(c, FGL_ERR_INVALID, "fgl_draw_triangles: bad arguments");
    if (c->pass == FGL_PASS_GEOMETRY && kind != FGL_SHADER_G)
        return fail(c, FGL_ERR_STATE, "geometry pass requires GShader (reference forkergl.cpp:214 dynamic_cast)");
    if (n && (kind == FGL_SHADER_DEPTH ? !lightZ : !vary)) return fail(c, FGL_ERR_INVALID, "fgl_draw_triangles: the shader kind's per-triangle array is NULL");
    Draw d;
    d.mesh = &c->meshes[meshId];
    d.vb = &c->vertices[d.mesh->vertices];
    d.un = *un;
    d.kind = kind;
    memset(d.lm, 0, sizeof d.lm);
    for (int t = 0; t < n; ++t)
    {
        Varyings vy;
        memset(&vy, 0, sizeof vy);
        for (int k = 0; k < 3; ++k)
        {
            const float* q = ndc + (size_t)t * 12 + 4 * k;
            vy.ndc[k] = V4{ q[0], q[1], q[2], q[3] };
            if (kind == FGL_SHADER_DEPTH)
            {
                vy.depthNdc[k] = v3(q[0], q[1], lightZ[(size_t)t * 3 + k]);
                continue;
            }
            const float* f = vary + (size_t)t * 48;
            vy.posWS[k] = v3(f[3 * k], f[3 * k + 1], f[3 * k + 2]);
            vy.nrmWS[k] = v3(f[9 + 3 * k], f[10 + 3 * k], f[11 + 3 * k]);
            vy.tanWS[k] = v3(f[18 + 3 * k], f[19 + 3 * k], f[20 + 3 * k]);
            vy.lightNDC[k] = v3(f[27 + 3 * k], f[28 + 3 * k], f[29 + 3 * k]);
            vy.u[k] = f[36 + k], vy.v[k] = f[39 + k], vy.oow[k] = f[42 + k];
        }
        DrawTriangle(c, d, vy, c->primCounter++);
    }
    return FGL_OK;
}

// forkergl.cpp:326-380
int fgl_draw_screen_space_pixels(fgl_ctx* c, const float eye[3], const float lpos[3], const float lcol[3])
{
    Plane& frame = c->planes[FGL_PLANE_FRAME];
    int    W = frame.w, H = frame.h;
    V3     eyePos = v3(eye[0], eye[1], eye[2]), lightPos = v3(lpos[0], lpos[1], lpos[2]), rad = v3(lcol[0], lcol[1], lcol[2]);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
        {
            size_t idx = (size_t)x + (size_t)y * W;
            V3     pos = get3(c->planes[FGL_PLANE_WORLDPOS], x, y), nrm = get3(c->planes[FGL_PLANE_NORMAL], x, y);
            V3     lndc = c->shadowOn ? get3(c->planes[FGL_PLANE_LIGHTNDC], x, y) : v3(0, 0, 0);
            V3     albedo = get3(c->planes[FGL_PLANE_ALBEDO], x, y), emissive = get3(c->planes[FGL_PLANE_EMISSIVE], x, y);
            V3     param = get3(c->planes[FGL_PLANE_PARAM], x, y);
            float  type = c->planes[FGL_PLANE_SHADINGTYPE].v[idx];
            param.x *= c->planes[FGL_PLANE_AO].v[idx];
            V3    lightDir = normalize(sub(lightPos, pos)), viewDir = normalize(sub(eyePos, pos));
            float visibility = 0.f;
            if (c->shadowOn) visibility = CalculateShadowVisibility(c, lndc, nrm, lightDir);
            V3 color;
            if (type < 0.5f)  // deferred Blinn-Phong receives viewDir as "halfway" (forkergl.cpp:369)
                color = BlinnPhongLight(c, lightDir, viewDir, nrm, visibility, albedo, emissive, param, rad);
            else
                color = PBRLight(c, lightDir, viewDir, normalize(add(lightDir, viewDir)), nrm, visibility, albedo, emissive, param, rad);
            set3(frame, x, y, color);
        }
    return FGL_OK;
}

// render.cpp:214-286
int fgl_ssao(fgl_ctx* c)
{
    Plane&      ao = c->planes[FGL_PLANE_AO];
    const Plane &wp = c->planes[FGL_PLANE_WORLDPOS], &nm = c->planes[FGL_PLANE_NORMAL], &dp = c->planes[FGL_PLANE_DEPTH];
    int         W = c->planes[FGL_PLANE_FRAME].w, H = c->planes[FGL_PLANE_FRAME].h;
    const float kernelScale = 1.f / 32;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
        {
            V3    pos = get3(wp, x, y), nrm = get3(nm, x, y);
            float fragDepth = dp.v[(size_t)x + (size_t)y * W];
            float occlusion = 0.f;
            for (int s = 0; s < 32; ++s)
            {
                V3    dir = RandomVectorInHemisphere(c, nrm);
                float sc = length(dir);
                sc = (1 - 0.1f) * 1.0f + 0.1f * (sc * sc);  // Lerp(0.1f, 1.0f, sc*sc), utility.h:26-29
                dir = scale(dir, sc);
                V3 sp = add(pos, scale(dir, c->params.ssao_radius));
                V4 cs = mat4(c->viewProj, V4{ sp.x, sp.y, sp.z, 1.f });
                V4 ndc = divs4(cs, cs.w);
                V4 ss = mat4(c->viewport, ndc);
                int       sx = (int)ss.x, sy = (int)ss.y;
                long long li = (long long)sx + (long long)sy * dp.w;  // unchecked index in the reference (buffer.h:37)
                if (li < 0 || li >= (long long)dp.w * dp.h) continue;  // UB there; defined as "no occlusion" 
                float cached = dp.v[(size_t)li];
                if (ss.z >= cached + c->params.ssao_bias)
                {
                    if (c->params.ssao_range_check)
                    {
                        float rc = std::abs(fragDepth - cached) < c->params.ssao_range_check_radius ? 1.f : 0.f;
                        occlusion += kernelScale * rc;
                    }
                    else
                        occlusion += kernelScale;
                }
            }
            occlusion = 1.f - occlusion;
            occlusion = std::pow(occlusion, (float)3);
            ao.v[(size_t)x + (size_t)y * W] = occlusion;
        }
    return FGL_OK;
}

// buffer.cpp:35-98 / 140-203 — in place, raster order
int fgl_blur(fgl_ctx* c, int plane, int kind)
{
    if (!okPlane(plane) || plane > FGL_PLANE_AO) return fail(c, FGL_ERR_INVALID, "fgl_blur: not an fp32 plane");
    Plane& p = c->planes[plane];
    int    W = p.w, H = p.h, C = p.ch;
    auto   at = [&](int x, int y, int ch) -> float& { return p.v[((size_t)x + (size_t)y * W) * C + ch]; };
    if (kind == FGL_BLUR_SIMPLE_3X3)
    {
        const float s = 1 / 9.f;
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w)
                for (int ch = 0; ch < C; ++ch)
                {
                    float r = 0.f;
                    for (int xo = -1; xo <= 1; ++xo)
                        for (int yo = -1; yo <= 1; ++yo) r += at(clampi(w + xo, 0, W - 1), clampi(h + yo, 0, H - 1), ch) * s;
                    at(w, h, ch) = r;
                }
        return FGL_OK;
    }
    const float g[5] = { 0.227027, 0.1945946, 0.1216216, 0.054054, 0.016216 };
    for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w)
            for (int ch = 0; ch < C; ++ch)
            {
                float r = at(w, h, ch) * g[0];
                for (int i = 1; i < 5; ++i)
                {
                    r += at(clampi(w + i, 0, W - 1), h, ch) * g[i];
                    r += at(clampi(w - i, 0, W - 1), h, ch) * g[i];
                }
                at(w, h, ch) = r;
            }
    for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w)
            for (int ch = 0; ch < C; ++ch)
            {
                float r = at(w, h, ch) * g[0];
                for (int i = 1; i < 5; ++i)
                {
                    r += at(w, clampi(h + i, 0, H - 1), ch) * g[i];
                    r += at(w, clampi(h - i, 0, H - 1), ch) * g[i];
                }
                at(w, h, ch) = r;
            }
    return FGL_OK;
}

static void quantizeFrame(fgl_ctx* c)  // buffer.cpp:113-126
{
    const Plane& f = c->planes[FGL_PLANE_FRAME];
    c->frameRgb8.resize((size_t)f.w * f.h * 3);
    for (size_t i = 0; i < (size_t)f.w * f.h * 3; ++i) c->frameRgb8[i] = (uint8_t)(f.v[i] * 254.99f);
}

// render.cpp:291-343
int fgl_ssaa_resolve(fgl_ctx* c, int k)
{
    if (k < 1) return fail(c, FGL_ERR_INVALID, "fgl_ssaa_resolve: kernel size < 1");
    quantizeFrame(c);
    const Plane& f = c->planes[FGL_PLANE_FRAME];
    int          ow = f.w / k, oh = f.h / k;
    c->ssaaW = ow, c->ssaaH = oh;
    c->ssaaRgb8.assign((size_t)ow * oh * 3, 0);
    for (int x = 0; x < ow; ++x)
        for (int y = 0; y < oh; ++y)
        {
            int R = 0, G = 0, B = 0;
            for (int i = 0; i < k; ++i)
                for (int j = 0; j < k; ++j)
                {
                    const uint8_t* p = &c->frameRgb8[((size_t)(x * k + i) + (size_t)(y * k + j) * f.w) * 3];
                    R += p[0], G += p[1], B += p[2];
                }
            R /= (float)(k * k);
            G /= (float)(k * k);
            B /= (float)(k * k);
            uint8_t* o = &c->ssaaRgb8[((size_t)x + (size_t)y * ow) * 3];
            o[0] = (uint8_t)R, o[1] = (uint8_t)G, o[2] = (uint8_t)B;
        }
    return FGL_OK;
}

// ---- buffers --------------------------------------------------------------------------------------------------
int fgl_plane_info(fgl_ctx* c, int plane, int* w, int* h, int* ch, int* bpc)
{
    if (!c || !okPlane(plane)) return fail(c, FGL_ERR_INVALID, "fgl_plane_info: bad plane");
    if (plane <= FGL_PLANE_AO) *w = c->planes[plane].w, *h = c->planes[plane].h, *ch = c->planes[plane].ch ? c->planes[plane].ch : ((plane == FGL_PLANE_DEPTH || plane == FGL_PLANE_SHADOW || plane == FGL_PLANE_SHADINGTYPE || plane == FGL_PLANE_AO) ? 1 : 3), *bpc = 4;
    else if (plane == FGL_PLANE_FRAME_RGB8) *w = c->planes[FGL_PLANE_FRAME].w, *h = c->planes[FGL_PLANE_FRAME].h, *ch = 3, *bpc = 1;
    else if (plane == FGL_PLANE_SSAA_RGB8) *w = c->ssaaW, *h = c->ssaaH, *ch = 3, *bpc = 1;
    else if (plane == FGL_PLANE_PRIMID_CAMERA) *w = c->planes[FGL_PLANE_DEPTH].w, *h = c->planes[FGL_PLANE_DEPTH].h, *ch = 1, *bpc = 4;
    else *w = c->planes[FGL_PLANE_SHADOW].w, *h = c->planes[FGL_PLANE_SHADOW].h, *ch = 1, *bpc = 4;
    return FGL_OK;
}

int fgl_read_plane(fgl_ctx* c, int plane, void* dst, size_t bytes)
{
    if (!c || !okPlane(plane) || !dst) return fail(c, FGL_ERR_INVALID, "fgl_read_plane: bad arguments");
    const void* src = nullptr;
    size_t      n = 0;
    if (plane <= FGL_PLANE_AO) src = c->planes[plane].v.data(), n = c->planes[plane].v.size() * 4;
    else if (plane == FGL_PLANE_FRAME_RGB8)
    {
        quantizeFrame(c);
        src = c->frameRgb8.data(), n = c->frameRgb8.size();
    }
    else if (plane == FGL_PLANE_SSAA_RGB8) src = c->ssaaRgb8.data(), n = c->ssaaRgb8.size();
    else if (plane == FGL_PLANE_PRIMID_CAMERA) src = c->idsCamera.data(), n = c->idsCamera.size() * 4;
    else src = c->idsLight.data(), n = c->idsLight.size() * 4;
    if (bytes != n) return fail(c, FGL_ERR_INVALID, "fgl_read_plane: size mismatch (have " + std::to_string(n) + ")");
    if (n) memcpy(dst, src, n);
    return FGL_OK;
}

int fgl_write_plane(fgl_ctx* c, int plane, const void* src, size_t bytes)
{
    if (!c || plane < 0 || plane > FGL_PLANE_AO || !src) return fail(c, FGL_ERR_INVALID, "fgl_write_plane: bad arguments");
    if (bytes != c->planes[plane].v.size() * 4) return fail(c, FGL_ERR_INVALID, "fgl_write_plane: size mismatch");
    memcpy(c->planes[plane].v.data(), src, bytes);
    return FGL_OK;
}

int fgl_copy_plane_rows_to_device(fgl_ctx* c, int, int, int, void*, size_t)
{
    return fail(c, FGL_ERR_UNSUPPORTED, "the CPU oracle has no device memory");
}

int fgl_enable_timing(fgl_ctx*, int) { return FGL_OK; }
int fgl_reset_timings(fgl_ctx*) { return FGL_OK; }
int fgl_get_timings(fgl_ctx*, FglTiming*, int, int* n) { if (n) *n = 0; return FGL_OK; }
int fgl_launch_count(fgl_ctx*, uint64_t* o) { if (o) *o = 0; return FGL_OK; }
int fgl_group_export(fgl_ctx* c, int, int, FglGroupMember*) { return fail(c, FGL_ERR_UNSUPPORTED, "the CPU oracle renders whole frames"); }
int fgl_group_connect(fgl_ctx* c, int, int, const FglGroupMember*, int) { return fail(c, FGL_ERR_UNSUPPORTED, "the CPU oracle renders whole frames"); }
int fgl_group_read_frame(fgl_ctx* c, void*, size_t) { return fail(c, FGL_ERR_UNSUPPORTED, "the CPU oracle renders whole frames"); }
int fgl_group_disconnect(fgl_ctx*) { return FGL_OK; }
int fgl_transfer_bytes(fgl_ctx*, uint64_t* a, uint64_t* b) { if (a) *a = 0; if (b) *b = 0; return FGL_OK; }
int fgl_prepare_screen_space_pixels(fgl_ctx*, const float*, const float*, const float*, int) { return FGL_OK; }  // nothing to split on one CPU
int fgl_chain_peer_mailbox(fgl_ctx* c, void**, void*, size_t) { return fail(c, FGL_ERR_UNSUPPORTED, "the CPU oracle has no peer memory"); }
int fgl_chain_peer_connect(fgl_ctx* c, void*, const void*, int, int enable) { return enable ? fail(c, FGL_ERR_UNSUPPORTED, "the CPU oracle has no peer memory") : FGL_OK; }
int fgl_host_alloc(fgl_ctx*, size_t bytes, void** out) { if (!out) return FGL_ERR_INVALID; *out = malloc(bytes ? bytes : 16); return *out ? FGL_OK : FGL_ERR_INVALID; }
int fgl_host_free(fgl_ctx*, void* p) { free(p); return FGL_OK; }
// recorded frames are CUDA graphs: the oracle executes every call as it arrives and has nothing to replay
int fgl_frame_record_begin(fgl_ctx* c) { return fail(c, FGL_ERR_UNSUPPORTED, "the CPU oracle does not record frames"); }
int fgl_frame_record_end(fgl_ctx* c, int*) { return fail(c, FGL_ERR_UNSUPPORTED, "the CPU oracle does not record frames"); }
int fgl_frame_record_abort(fgl_ctx*) { return FGL_OK; }
int fgl_frame_replay(fgl_ctx* c, int) { return fail(c, FGL_ERR_UNSUPPORTED, "the CPU oracle does not record frames"); }
int fgl_frame_info(fgl_ctx* c, int, int*, int*) { return fail(c, FGL_ERR_UNSUPPORTED, "the CPU oracle does not record frames"); }
int fgl_frame_release(fgl_ctx*, int) { return FGL_OK; }

// oracle-only: number of mt19937 draws consumed since fgl_begin_frame (used by the stream-accounting tests)
uint64_t orc_rng_draws(fgl_ctx* c) { return c->draws; }
// oracle-only: raw access to the sample stream for known-answer tests
uint32_t orc_mt19937_nth(uint32_t seed, uint64_t n)
{
    Mt19937 g;
    g.seed(seed);
    uint32_t v = 0;
    for (uint64_t i = 0; i <= n; ++i) v = g.next();
    return v;
}
float orc_random01_from_u32(uint32_t u)
{
    float r = (float)u / 4294967296.f;
    if (r >= 1.f) r = std::nextafter(1.f, 0.f);
    return r;
}
}  // extern "C"
