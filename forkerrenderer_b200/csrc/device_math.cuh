// device_math.cuh — fp32 vector helpers, texture sampling, shadow lookups and the two lighting models, written
// for sm_100a with the reference's order of evaluation so that results are bit-comparable with the CPU renderer.
//
// The whole library is compiled with -fmad=false (no FMA contraction: the reference is x86-64 SSE2 scalar code
// without FMA), IEEE division and square root.  Every helper cites the reference file:line
// whose arithmetic it restates.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/forkergl_b200.h"

#define FGL_HD __host__ __device__ __forceinline__
#define FGL_D __device__ __forceinline__

struct V3
{
    float x, y, z;
};
struct V4
{
    float x, y, z, w;
};

FGL_HD V3 v3(float x, float y, float z)
{
    V3 r;
    r.x = x, r.y = y, r.z = z;
    return r;
}
FGL_HD V3 vadd(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
FGL_HD V3 vsub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
FGL_HD V3 vmul(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
FGL_HD V3 vscale(V3 a, float f) { return v3(a.x * f, a.y * f, a.z * f); }
// reference geometry.h:881-889 — Dot accumulates from 0.f, left to right
FGL_HD float vdot(V3 a, V3 b)
{
    float r = 0.f;
    r += a.x * b.x;
    r += a.y * b.y;
    r += a.z * b.z;
    return r;
}
FGL_HD float dot4(const float* row, V4 v)
{
    float r = 0.f;
    r += row[0] * v.x;
    r += row[1] * v.y;
    r += row[2] * v.z;
    r += row[3] * v.w;
    return r;
}
FGL_HD float dot3(const float* row, V3 v)
{
    float r = 0.f;
    r += row[0] * v.x;
    r += row[1] * v.y;
    r += row[2] * v.z;
    return r;
}
// reference geometry.h:774-782 — matrix x vector = one Dot per row
FGL_HD V4 mat4mul(const float* m, V4 v)
{
    V4 r;
    r.x = dot4(m, v), r.y = dot4(m + 4, v), r.z = dot4(m + 8, v), r.w = dot4(m + 12, v);
    return r;
}
FGL_HD V3 mat3mul(const float* m, V3 v) { return v3(dot3(m, v), dot3(m + 3, v), dot3(m + 6, v)); }
// reference geometry.h:335-341 — vector / scalar is "reciprocal, then multiply"
FGL_HD V3 vdivs(V3 a, float f)
{
    float inv = 1.f / f;
    return v3(a.x * inv, a.y * inv, a.z * inv);
}
FGL_HD V4 vdivs4(V4 a, float f)
{
    float inv = 1.f / f;
    V4    r;
    r.x = a.x * inv, r.y = a.y * inv, r.z = a.z * inv, r.w = a.w * inv;
    return r;
}
FGL_HD float vlength(V3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }  // geometry.h:370-371
FGL_HD V3    vnormalize(V3 a) { return vdivs(a, vlength(a)); }                   // geometry.h:906-910 (v*1 is exact)
// SSAO (render.cpp:255-259): the scale of a hemisphere sample, Lerp(0.1f, 1.0f, len * len) with the reference's argument order.
// A function of the sample alone (and the same for v and -v), so the sample table stores it next to the vector.
FGL_HD float ssao_sample_scale(V3 v)
{
    float sc = vlength(v);
    return (1 - 0.1f) * 1.0f + 0.1f * (sc * sc);
}
FGL_HD V3    vcross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
FGL_HD float clampf(float v, float lo, float hi) { return fminf(hi, fmaxf(v, lo)); }  // utility.h:32-35
FGL_HD int   clampi(int v, int lo, int hi) { return min(hi, max(v, lo)); }
// reference geometry.h:912-916 — Lerp(t, a, b) = (1 - t) * a + t * b
FGL_HD V3 vlerp(float t, V3 a, V3 b) { return vadd(vscale(a, 1 - t), vscale(b, t)); }

// C float->int conversion as x86-64 cvttss2si performs it: NaN / out of range -> INT_MIN.
FGL_HD int f2i_x86(float f)
{
    if (!(f > -2147483904.f && f < 2147483648.f)) return (int)0x80000000;
    return (int)f;
}

// interpolation of one varying row with the barycentric vector (Matrix row . bary, geometry.h:774-782)
FGL_HD float interp(float a0, float a1, float a2, V3 b)
{
    float r = 0.f;
    r += a0 * b.x;
    r += a1 * b.y;
    r += a2 * b.z;
    return r;
}

// ---------------------------------------------------------------------------------------------------------
// depth keys: 64-bit (orderable(depth) << 32) | primitive id, minimised with atomicMin
FGL_HD uint32_t depth_to_ordered(float d)
{
    if (d == 0.f) d = 0.f;  // -0 and +0 tie in the reference's `>=` test
    uint32_t b;
#ifdef __CUDA_ARCH__
    b = __float_as_uint(d);
#else
    union { float f; uint32_t u; } cv;
    cv.f = d;
    b = cv.u;
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
FGL_HD float ordered_to_depth(uint32_t o)
{
    uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    union { float f; uint32_t u; } cv;
    cv.u = b;
    return cv.f;
#endif
}
#define FGL_VIS_EMPTY 0xFFFFFFFFFFFFFFFFull

// ---------------------------------------------------------------------------------------------------------
// Coverage: reference geometry.cpp:20-56 evaluated literally in double on the integer-snapped vertices.
struct TriCover
{
    float  ax, ay;
    double s0x, s0y, s1x, s1y;  // (bx-ax), (cx-ax), (by-ay), (cy-ay): fp32 subtractions, widened
    double rz, inv;             // signed 2*area and 1.f / rz (double division)
    bool   valid;               // |rz| > 1e-2
};
FGL_HD TriCover make_cover(int X0, int Y0, int X1, int Y1, int X2, int Y2)
{
    TriCover t;
    float    ax = (float)X0, ay = (float)Y0, bx = (float)X1, by = (float)Y1, cx = (float)X2, cy = (float)Y2;
    t.ax = ax, t.ay = ay;
    t.s0x = (double)(bx - ax), t.s0y = (double)(cx - ax);
    t.s1x = (double)(by - ay), t.s1y = (double)(cy - ay);
    t.rz = t.s0x * t.s1y - t.s0y * t.s1x;
    t.valid = fabs(t.rz) > 1e-2;
    t.inv = 1.0 / t.rz;
    return t;
}
FGL_HD bool cover_test(const TriCover& t, int px, int py, V3& bary)
{
    float  fx = (float)px, fy = (float)py;
    double s0z = (double)(t.ax - fx), s1z = (double)(t.ay - fy);
    double rx = t.s0y * s1z - s0z * t.s1y;
    double ry = s0z * t.s1x - t.s0x * s1z;
    rx = rx * t.inv;
    ry = ry * t.inv;
    float r0 = (float)(1.0 - (rx + ry)), r1 = (float)rx, r2 = (float)ry;
    if (r0 < 0.f || r1 < 0.f || r2 < 0.f) return false;
    bary = v3(r0, r1, r2);
    return true;
}

// ---------------------------------------------------------------------------------------------------------
// Textures (reference src/texture.h:41-145, tgaimage.cpp:304-311)
struct TexD
{
    const uint8_t* data;
    int            w, h, bpp, wrap, filter;
};

FGL_D V3 tex_texel(const TexD& t, int x, int y)  // [0,255]; out of image = black; a grey image keeps its value in b
{
    if (x < 0 || y < 0 || x >= t.w || y >= t.h) return v3(0.f, 0.f, 0.f);
    const uint8_t* p = t.data + ((size_t)x + (size_t)y * t.w) * t.bpp;
    if (t.bpp == 1) return v3(0.f, 0.f, (float)__ldg(p));
    return v3((float)__ldg(p + 2), (float)__ldg(p + 1), (float)__ldg(p));
}

FGL_D void tex_wrap(const TexD& t, float& u, float& v)  // texture.h:62-83
{
    if (t.wrap == FGL_WRAP_REPEAT)
    {
        u = u - floorf(u);
        v = v - floorf(v);
    }
    else if (t.wrap == FGL_WRAP_MIRRORED_REPEAT)
    {
        int   xi = f2i_x86(floorf(u)), yi = f2i_x86(floorf(v));
        float rx = u - (float)xi, ry = v - (float)yi;
        u = xi % 2 == 0 ? rx : 1.f - rx;
        v = yi % 2 == 0 ? ry : 1.f - ry;
    }
    else if (t.wrap == FGL_WRAP_CLAMP_TO_EDGE)
    {
        u = clampf(u, 0.f, 1.f);
        v = clampf(v, 0.f, 1.f);
    }
}

FGL_D V3 tex_filter(const TexD& t, float u, float v)  // texture.h:86-132
{
    float w = (float)((double)t.w - 0.001), h = (float)((double)t.h - 0.001);
    if (t.filter == FGL_FILTER_LINEAR)
    {
        float px = u * w, py = v * h;
        float tlx = floorf(px - 0.5f), tly = floorf(py - 0.5f);
        float tx = px - (tlx + 0.5f), ty = py - (tly + 0.5f);
        int   x0 = f2i_x86(tlx), y0 = f2i_x86(tly), x1 = f2i_x86(tlx + 1.f), y1 = f2i_x86(tly + 1.f);
        if (t.wrap != FGL_WRAP_NOWRAP)
        {
            x0 = clampi(x0, 0, t.w - 1), x1 = clampi(x1, 0, t.w - 1);
            y0 = clampi(y0, 0, t.h - 1), y1 = clampi(y1, 0, t.h - 1);
        }
        V3 c0 = tex_texel(t, x0, y0), c1 = tex_texel(t, x1, y0), c2 = tex_texel(t, x0, y1), c3 = tex_texel(t, x1, y1);
        V3 cx1 = vlerp(tx, c0, c1), cx2 = vlerp(tx, c2, c3);
        return vlerp(ty, cx1, cx2);
    }
    int ix = f2i_x86(floorf(u * w)), iy = f2i_x86(floorf(v * h));
    return tex_texel(t, ix, iy);
}
FGL_D V3 tex_sample(const TexD& t, float u, float v)  // texture.h:41-45 (vector / 255.f: reciprocal-multiply)
{
    tex_wrap(t, u, v);
    return vdivs(tex_filter(t, u, v), 255.f);
}
FGL_D float tex_sample_float(const TexD& t, float u, float v)  // texture.h:47-51 (true division of channel b)
{
    tex_wrap(t, u, v);
    return tex_filter(t, u, v).z / 255.f;
}

// ---------------------------------------------------------------------------------------------------------
// Shadow map point lookup (reference src/shaders/shadow.cpp:23-35)
struct ShadowMapD
{
    const float* d;
    int          w, h;
    int          iw, ih;  // (int)(W - 0.001f), (int)(H - 0.001f)
};
FGL_D float shadow_lookup(const ShadowMapD& sm, float u, float v)
{
    if (!(u >= 0.f && u <= 1.f && v >= 0.f && v <= 1.f)) return __int_as_float(0x7f800000);  // NaN coordinates would index out of bounds in the reference
    // u, v are in [0, 1] here, so the products are in [0, iw] x [0, ih]: the plain truncating conversion is the x86 one
    // (no INT_MIN case), and the texel index fits 32 bits (shadow maps have < 2^31 texels: fgl_init_shadow_buffer)
    int   iu = (int)((float)sm.iw * u);
    int   iv = (int)((float)sm.ih * v);
    float depth = __ldg(sm.d + (iu + iv * sm.w));
    return (depth < 0.001f) ? 1.f : depth;  // == ((double)depth < 0.001), shadow.cpp:33-34: float(0.001) is the smallest float above 0.001
}

// ---------------------------------------------------------------------------------------------------------
// powf: the reference calls glibc powf (correctly rounded in all but rare cases).  CUDA's powf is within a few
// ulp of it, which only moves 8-bit colours that sit on a quantisation boundary (DESIGN.md "tolerances").
FGL_D float fgl_pow(float a, float b) { return powf(a, b); }
// The gamma conversions (exponents 2.2 and 1 / 2.2: positive, not odd integers): pow(+-0, y) is +0 by definition, in glibc and in
// CUDA alike, so a zero base needs no evaluation — emissive colours are zero almost everywhere and background pixels are zero
// in every channel, which makes this three to nine of a pixel's ten powf calls.
FGL_D float fgl_pow_gamma(float a, float p) { return a == 0.f ? 0.f : powf(a, p); }
// The Blinn-Phong highlight pow(max(0, N.H), shininess): pow(x, +-0) is 1 for every x and pow(+0, y > 0) is +0, again by
// definition in both libraries — background pixels (shininess 0) and surfaces turned away from the highlight need no evaluation.
// (+0 only: pow(-0, odd integer) is -0)
FGL_D float fgl_pow_spec(float a, float p) { return p == 0.f ? 1.f : (__float_as_uint(a) == 0u && p > 0.f ? 0.f : powf(a, p)); }
FGL_D V3    vpow(V3 v, float p) { return v3(fgl_pow_gamma(v.x, p), fgl_pow_gamma(v.y, p), fgl_pow_gamma(v.z, p)); }
FGL_D V3    vclamp01(V3 v) { return v3(clampf(v.x, 0.f, 1.f), clampf(v.y, 0.f, 1.f), clampf(v.z, 0.f, 1.f)); }

struct LightConsts
{
    float shadowIntensity;
    int   shadowOn;
};

// reference src/shaders/phongshader.h:171-215
FGL_D V3 blinn_phong_light(const LightConsts& lc, V3 lightDir, V3 halfwayDir, V3 normal, float visibility, V3 diffuseColor,
                           V3 emissive, V3 param, V3 lightColor)
{
    const float kGamma = 2.2f, kInvGamma = 1.f / 2.2f;
    V3          dl = vpow(diffuseColor, kGamma), el = vpow(emissive, kGamma);
    float       ao = param.x, ks = param.y, shininess = param.z;
    float       diff = fmaxf(0.f, vdot(lightDir, normal));
    float       spec = fgl_pow_spec(fmaxf(0.f, vdot(halfwayDir, normal)), shininess);
    V3          ambient = vscale(vmul(v3(0.3f, 0.3f, 0.3f), dl), ao);
    V3          diffuse = vscale(vscale(dl, diff), ao);
    V3          specular = vscale(v3(ks, ks, ks), spec);
    if (lc.shadowOn)
    {
        float shadow = (1 - visibility) * lc.shadowIntensity;
        visibility = 1 - shadow;
        diffuse = vscale(diffuse, visibility);
        specular = vscale(specular, visibility);
    }
    V3 color = vadd(ambient, vmul(vadd(vadd(diffuse, specular), el), lightColor));
    V3 den = vadd(color, v3(1.f, 1.f, 1.f));
    color = v3(color.x / den.x, color.y / den.y, color.z / den.z);
    color = vpow(color, kInvGamma);
    return vclamp01(color);
}

// reference src/shaders/pbrshader.h:182-288
FGL_D V3 pbr_light(const LightConsts& lc, V3 lightDir, V3 viewDir, V3 halfwayDir, V3 normal, float visibility, V3 albedo,
                   V3 emissive, V3 param, V3 lightRadiance)
{
    const float kGamma = 2.2f, kInvGamma = 1.f / 2.2f, kInvPi = 0.31830988618379067154f;
    V3          al = vpow(albedo, kGamma), el = vpow(emissive, kGamma);
    float       ao = param.x, metalness = param.y, roughness = param.z;
    float       NdotV = fmaxf(vdot(normal, viewDir), 0.f);
    float       NdotL = fmaxf(vdot(normal, lightDir), 0.f);
    float       NdotH = fmaxf(vdot(normal, halfwayDir), 0.f);
    float       HdotV = fmaxf(vdot(halfwayDir, viewDir), 0.f);
    V3          F0 = vlerp(metalness, v3(0.04f, 0.04f, 0.04f), al);
    float       a = roughness * roughness, a2 = a * a, NdotH2 = NdotH * NdotH;  // pbrshader.h:256-266
    float       den = (NdotH2 * (a2 - 1.f) + 1.f);
    float       NDF = a2 * kInvPi / (den * den);
    float       ka = roughness + 1.f, k = ka * ka / 8.f;  // pbrshader.h:268-282
    float       ggx1 = NdotV / (NdotV * (1 - k) + k);
    float       ggx2 = NdotL / (NdotL * (1 - k) + k);
    float       G = ggx1 * ggx2;
    float       om = fmaxf(1.f - HdotV, 0.f);  // pbrshader.h:284-288
    float       p5 = fgl_pow(om, 5.f);
    V3          F = vadd(F0, vscale(vsub(v3(1.f, 1.f, 1.f), F0), p5));
    V3          DGF = vscale(F, NDF * G);
    float       denominator = 4 * NdotV * NdotL + 0.001f;
    V3          specular = vdivs(DGF, denominator);
    V3          kd = vsub(v3(1.f, 1.f, 1.f), F);
    kd = vscale(kd, 1.f - metalness);
    V3 brdf = vadd(vscale(vmul(kd, al), kInvPi), specular);
    V3 Lo = vscale(vmul(brdf, lightRadiance), NdotL);
    if (lc.shadowOn)
    {
        float shadow = (1 - visibility) * lc.shadowIntensity;
        visibility = 1 - shadow;
        Lo = vscale(Lo, visibility);
    }
    V3 color = Lo;
    color = vadd(color, vscale(vmul(v3(0.3f, 0.3f, 0.3f), al), ao));
    color = vadd(color, el);
    V3 d = vadd(color, v3(1.f, 1.f, 1.f));
    color = v3(color.x / d.x, color.y / d.y, color.z / d.z);
    color = vpow(color, kInvGamma);
    return vclamp01(color);
}

// Random01 of the reference (utility.h:90-98; libstdc++ generate_canonical<float,24> over one 32-bit draw)
FGL_HD float random01_from_u32(uint32_t u)
{
    float r = (float)u * 2.3283064365386963e-10f;  // float(u) (round to nearest) * 2^-32 (exact scaling)
    return r >= 1.f ? 0.99999994f : r;
}
// Random(a, b) = a + (b - a) * Random01()  (utility.h:100-103), for (a, b) = (-1, 1)
FGL_HD float random_m1p1(uint32_t u) { return -1.f + 2.f * random01_from_u32(u); }
