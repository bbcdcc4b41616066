// programs.cpp — the four shader programs on the HOST: Shader::ProcessVertex / ProcessFragment of DepthShader, GShader,
// BlinnPhongShader and PBRShader (reference src/shaders/depthshader.h:21-36, gshader.h:41-201, phongshader.h:35-215,
// pbrshader.h:35-288), written against the facade's own types.
//
// ProcessVertex is on the per-triangle submission path (the reference's Mesh::Draw loop, mesh.cpp:10-25, followed by
// ForkerGL::DrawTriangle): it evaluates the vertex program with the expressions and the order of the device program
// (csrc/raster.cu k_setup) and stores the triangle's varyings in the packed layout fgl_draw_triangles takes.
// ProcessFragment is for callers that drive the Shader interface by hand; the device passes run their own fragment programs.
#include <cmath>
#include <limits>
#include <random>

#include "forkergl.h"
#include "model.h"
#include "shader.h"
#include "shadow.h"

namespace
{
// ---- vertex stage shared by the three camera programs (gshader.h:41-92 == phongshader.h:35-85 == pbrshader.h:35-85) --------
Point4f CameraVertex(Shader& s, int face, int vert, const Matrix4x4f& model, const Matrix4x4f& view, const Matrix4x4f& proj,
                     const Matrix3x3f& normalMatrix, const Matrix4x4f& lightSpace)
{
    const Mesh& mesh = *s.mesh;
    const bool  tangents = mesh.GetModel().HasTangents(), shadows = Shadow::GetShadowStatus();
    Point4f     ws = model * Point4f(mesh.Vert(face, vert), 1.f);
    Point4f     cs = proj * (view * ws);
    Vector2f    uv = mesh.TexCoord(face, vert);
    Vector3f    n = normalMatrix * mesh.Normal(face, vert);
    Vector3f    t = tangents ? normalMatrix * mesh.Tangent(face, vert) : Vector3f(0.f);
    Point4f     ls;
    if (shadows)
    {
        ls = lightSpace * ws;
        ls = ls / ls.w;
    }
    // perspective-correct interpolation: every attribute travels multiplied by 1 / w_clip
    const Float oow = 1.f / cs.w;
    float*      f = s.varyings;
    if (vert == 0)
        for (int i = 0; i < 48; ++i) f[i] = 0.f;
    f[42 + vert] = oow;
    f[3 * vert] = ws.x * oow, f[3 * vert + 1] = ws.y * oow, f[3 * vert + 2] = ws.z * oow;
    f[36 + vert] = uv.x * oow, f[39 + vert] = uv.y * oow;
    f[9 + 3 * vert] = n.x * oow, f[10 + 3 * vert] = n.y * oow, f[11 + 3 * vert] = n.z * oow;
    if (tangents) f[18 + 3 * vert] = t.x * oow, f[19 + 3 * vert] = t.y * oow, f[20 + 3 * vert] = t.z * oow;
    if (shadows) f[27 + 3 * vert] = ls.x * oow, f[28 + 3 * vert] = ls.y * oow, f[29 + 3 * vert] = ls.z * oow;
    return cs / cs.w;
}

// one varying row against the barycentric vector (matrix x vector = Dot per row, accumulated from 0)
inline Float Row(Float a0, Float a1, Float a2, const Vector3f& b)
{
    Float r = 0.f;
    r += a0 * b.x;
    r += a1 * b.y;
    r += a2 * b.z;
    return r;
}
inline Vector3f Attr(const float* f, int base, const Vector3f& b)  // vertex-major triple
{
    return Vector3f(Row(f[base], f[base + 3], f[base + 6], b), Row(f[base + 1], f[base + 4], f[base + 7], b), Row(f[base + 2], f[base + 5], f[base + 8], b));
}

struct Surface
{
    Point3f  position, lightNDC;
    Vector3f normal;
    Vector2f uv;
};
// head of the camera-space fragment programs (gshader.h:95-146 == phongshader.h:90-128 == pbrshader.h:90-128)
Surface Interpolate(const Shader& s, const Vector3f& bary, const std::shared_ptr<Texture>& normalMap)
{
    const float* f = s.varyings;
    Surface      o;
    Point3f      pos = Attr(f, 0, bary);
    Vector2f     uv(Row(f[36], f[37], f[38], bary), Row(f[39], f[40], f[41], bary));
    Vector3f     nrm = Attr(f, 9, bary);
    const Float  w = 1.f / Dot(Vector3f(f[42], f[43], f[44]), bary);
    pos *= w, uv *= w, nrm *= w;
    const Vector3f N = Normalize(nrm);
    o.normal = N;
    if (s.mesh->GetModel().HasTangents() && normalMap)
    {
        Vector3f tg = Attr(f, 18, bary);
        tg *= w;
        Vector3f T = Normalize(tg + Vector3f(0.001f));
        T = Normalize(T - Dot(T, N) * N);
        const Vector3f B = Normalize(Cross(N, T));
        Vector3f       sn = normalMap->Sample(uv);
        sn = Normalize(sn * 2.f - Vector3f(1.f));
        o.normal = Normalize(Vector3f(Dot(Vector3f(T.x, B.x, N.x), sn), Dot(Vector3f(T.y, B.y, N.y), sn), Dot(Vector3f(T.z, B.z, N.z), sn)));
    }
    o.position = pos, o.uv = uv;
    if (Shadow::GetShadowStatus())
    {
        o.lightNDC = Attr(f, 27, bary);
        o.lightNDC *= w;
    }
    return o;
}

inline Color3 PowV(const Color3& c, Float p) { return Color3(std::pow(c.x, p), std::pow(c.y, p), std::pow(c.z, p)); }
inline Color3 Clamp01(const Color3& c) { return Color3(Clamp(c.x, 0.f, 1.f), Clamp(c.y, 0.f, 1.f), Clamp(c.z, 0.f, 1.f)); }
inline Color3 Tonemap(Color3 c)  // Reinhard, gamma, clamp
{
    const Color3 d = c + Color3(1.f);
    c = Color3(c.x / d.x, c.y / d.y, c.z / d.z);
    return Clamp01(PowV(c, InvGamma));
}
}  // namespace

// ---- DepthShader -------------------------------------------------------------------------------------------------------
Point4f DepthShader::ProcessVertex(int faceIdx, int vertIdx)
{
    Point4f cs = (uLightSpaceMatrix * uModelMatrix) * Point4f(mesh->Vert(faceIdx, vertIdx), 1.f);
    Point4f ndc = cs / cs.w;
    lightZ[vertIdx] = ndc.z;
    return ndc;
}
bool DepthShader::ProcessFragment(const Vector3f& bary, Color3& gl_Color)
{
    gl_Color.z = Row(lightZ[0], lightZ[1], lightZ[2], bary) * 0.5f + 0.5f;
    return false;
}

// ---- GShader -----------------------------------------------------------------------------------------------------------
Point4f GShader::ProcessVertex(int faceIdx, int vertIdx)
{
    return CameraVertex(*this, faceIdx, vertIdx, uModelMatrix, uViewMatrix, uProjectionMatrix, uNormalMatrix, uLightSpaceMatrix);
}
bool GShader::ProcessFragment(const Vector3f& bary, Color3&)
{
    std::shared_ptr<const Material>    m = mesh->GetMaterial();
    std::shared_ptr<const PBRMaterial> p = mesh->GetPBRMaterial();
    const Surface                      s = Interpolate(*this, bary, m->normalMap);
    outNormalWS = s.normal, outPositionWS = s.position;
    if (Shadow::GetShadowStatus()) outLightSpaceNDC = s.lightNDC;
    if (mesh->GetModel().SupportPBR())
    {
        outAlbedo = p->HasBaseColorMap() ? p->baseColorMap->Sample(s.uv) : p->albedo;
        outEmissive = p->HasEmssiveMap() ? p->emissiveMap->Sample(s.uv) : m->ke;  // (the Material's ke, as in gshader.h:166)
        const Float roughness = p->HasRoughnessMap() ? p->roughnessMap->SampleFloat(s.uv) : p->roughness;
        const Float metalness = p->HasMetalnessMap() ? p->metalnessMap->SampleFloat(s.uv) : p->metalness;
        const Float ao = p->HasAmbientOcclusionMap() ? p->ambientOcclusionMap->SampleFloat(s.uv) : 1.f;
        outParam = Vector3f(ao, metalness, roughness);
        outShadingType = 1.f;
    }
    else
    {
        outEmissive = m->HasEmissiveMap() ? m->emissiveMap->Sample(s.uv) : m->ke;
        outAlbedo = m->HasDiffuseMap() ? m->diffuseMap->Sample(s.uv) : m->kd;
        outParam = Vector3f(1.f, m->ks.r, m->HasSpecularMap() ? m->specularMap->SampleFloat(s.uv) + 5 : 1.f);
        outShadingType = 0.f;
    }
    return false;
}

// ---- BlinnPhongShader ----------------------------------------------------------------------------------------------------
Point4f BlinnPhongShader::ProcessVertex(int faceIdx, int vertIdx)
{
    return CameraVertex(*this, faceIdx, vertIdx, uModelMatrix, uViewMatrix, uProjectionMatrix, uNormalMatrix, uLightSpaceMatrix);
}
bool BlinnPhongShader::ProcessFragment(const Vector3f& bary, Color3& gl_Color)
{
    std::shared_ptr<const Material> m = mesh->GetMaterial();
    const Surface                   s = Interpolate(*this, bary, m->normalMap);
    const Vector3f lightDir = Normalize(uPointLight.position - s.position), viewDir = Normalize(uEyePos - s.position);
    const Vector3f halfwayDir = Normalize(lightDir + viewDir);
    Float          visibility = 0.f;
    if (Shadow::GetShadowStatus()) visibility = Shadow::CalculateShadowVisibility(ForkerGL::ShadowBuffer, s.lightNDC, s.normal, lightDir);
    const Color3 diffuseColor = m->HasDiffuseMap() ? m->diffuseMap->Sample(s.uv) : m->kd;
    const Color3 emissive = m->HasEmissiveMap() ? m->emissiveMap->Sample(s.uv) : m->ke;
    const Float  shininess = m->HasSpecularMap() ? m->specularMap->SampleFloat(s.uv) + 5 : 1.f;
    gl_Color = CalculateLight(lightDir, halfwayDir, s.normal, visibility, diffuseColor, emissive, Vector3f(m->ka.x, m->ks.x, shininess), uPointLight.color);
    return false;
}
Color3 BlinnPhongShader::CalculateLight(const Vector3f& lightDir, const Vector3f& halfwayDir, const Vector3f& normal, Float visibility,
                                        const Color3& diffuseColor, const Color3& emissive, const Vector3f& param, const Color3& lightColor)
{
    const Color3 dl = PowV(diffuseColor, Gamma), el = PowV(emissive, Gamma);
    const Float  ao = param.x, ks = param.y, shininess = param.z;
    const Float  diff = std::max(0.f, Dot(lightDir, normal));
    const Float  spec = std::pow(std::max(0.f, Dot(halfwayDir, normal)), shininess);
    const Color3 ambient = Color3(0.3f) * dl * ao;
    Color3       diffuse = dl * diff * ao, specular = Color3(ks) * spec;
    if (Shadow::GetShadowStatus())
    {
        const Float shadow = (1 - visibility) * 0.6f;
        visibility = 1 - shadow;
        diffuse *= visibility, specular *= visibility;
    }
    return Tonemap(ambient + (diffuse + specular + el) * lightColor);
}

// ---- PBRShader -----------------------------------------------------------------------------------------------------------
Point4f PBRShader::ProcessVertex(int faceIdx, int vertIdx)
{
    return CameraVertex(*this, faceIdx, vertIdx, uModelMatrix, uViewMatrix, uProjectionMatrix, uNormalMatrix, uLightSpaceMatrix);
}
bool PBRShader::ProcessFragment(const Vector3f& bary, Color3& gl_Color)
{
    std::shared_ptr<const PBRMaterial> p = mesh->GetPBRMaterial();
    const Surface                      s = Interpolate(*this, bary, p->normalMap);
    const Vector3f lightDir = Normalize(uPointLight.position - s.position), viewDir = Normalize(uEyePos - s.position);
    const Vector3f halfwayDir = Normalize(lightDir + viewDir);
    Float          visibility = 0.f;
    if (Shadow::GetShadowStatus()) visibility = Shadow::CalculateShadowVisibility(ForkerGL::ShadowBuffer, s.lightNDC, s.normal, lightDir);
    const Color3 albedo = p->HasBaseColorMap() ? p->baseColorMap->Sample(s.uv) : p->albedo;
    const Color3 emissive = p->HasEmssiveMap() ? p->emissiveMap->Sample(s.uv) : p->ke;
    const Float  roughness = p->HasRoughnessMap() ? p->roughnessMap->SampleFloat(s.uv) : p->roughness;
    const Float  metalness = p->HasMetalnessMap() ? p->metalnessMap->SampleFloat(s.uv) : p->metalness;
    const Float  ao = p->HasAmbientOcclusionMap() ? p->ambientOcclusionMap->SampleFloat(s.uv) : 1.f;
    gl_Color = CalculateLight(lightDir, viewDir, halfwayDir, s.normal, visibility, albedo, emissive, Vector3f(ao, metalness, roughness), uPointLight.color);
    return false;
}
Color3 PBRShader::CalculateLight(const Vector3f& lightDir, const Vector3f& viewDir, const Vector3f& halfwayDir, const Vector3f& normal,
                                 Float visibility, const Color3& albedo, const Color3& emissive, const Vector3f& param, const Color3& lightRadiance)
{
    const Color3 al = PowV(albedo, Gamma), el = PowV(emissive, Gamma);
    const Float  ao = param.x, metalness = param.y, roughness = param.z;
    const Float  NdotV = std::max(Dot(normal, viewDir), 0.f), NdotL = std::max(Dot(normal, lightDir), 0.f);
    const Float  NdotH = std::max(Dot(normal, halfwayDir), 0.f), HdotV = std::max(Dot(halfwayDir, viewDir), 0.f);
    const Color3 F0 = Color3(0.04f) * (1 - metalness) + al * metalness;  // Lerp(metalness, 0.04, albedo)
    // GGX normal distribution, Schlick-GGX geometry (Smith), Fresnel-Schlick
    const Float a = roughness * roughness, a2 = a * a, den = (NdotH * NdotH) * (a2 - 1.f) + 1.f;
    const Float NDF = a2 * InvPi / (den * den);
    const Float kr = roughness + 1.f, k = kr * kr / 8.f;
    const Float G = (NdotV / (NdotV * (1 - k) + k)) * (NdotL / (NdotL * (1 - k) + k));
    const Float p5 = std::pow(std::max(1.f - HdotV, 0.f), 5.f);
    const Color3 F = F0 + (Color3(1.f) - F0) * p5;
    const Color3 specular = (F * (NDF * G)) / (4 * NdotV * NdotL + 0.001f);
    Color3       kd = Color3(1.f) - F;
    kd *= 1.f - metalness;
    Color3 Lo = ((kd * al) * InvPi + specular) * lightRadiance * NdotL;
    if (Shadow::GetShadowStatus())
    {
        const Float shadow = (1 - visibility) * 0.6f;
        visibility = 1 - shadow;
        Lo *= visibility;
    }
    return Tonemap(Lo + Color3(0.3f) * al * ao + el);
}

// ---- Shadow: the filters on the host (reference src/shaders/shadow.cpp:23-132) ----------------------------------------------
namespace Shadow
{
namespace
{
// the reference's global sample stream (utility.h:90-103): one default-seeded mt19937 behind generate_canonical<float, 24>
std::mt19937 g_Generator;
Float        Random01()
{
    static std::uniform_real_distribution<Float> distribution(0.f, 1.f);
    return distribution(g_Generator);
}
Float RandomM1P1() { return -1.f + (1.f - -1.f) * Random01(); }
void  DiskSample(Float& x, Float& y)  // geometry.h:968-976; g++ evaluates the constructor arguments right to left: y is drawn first
{
    do
    {
        y = RandomM1P1();
        x = RandomM1P1();
    } while (x * x + y * y + 0.f * 0.f >= 1.f);
}
Float Lookup(const Buffer1f& map, Float u, Float v)
{
    if (u < 0.f || u > 1.f || v < 0.f || v > 1.f) return std::numeric_limits<Float>::infinity();
    const int   w = (int)((Float)map.GetWidth() - 0.001f), h = (int)((Float)map.GetHeight() - 0.001f);
    const Float depth = map.GetValue((int)((Float)w * u), (int)((Float)h * v));
    return (double)depth < 0.001 ? 1.f : depth;
}
Float Filter(const Buffer1f& map, const Vector3f& sc, Float bias, Float size)  // PCF, 64 taps
{
    Float visibility = 0.f;
    for (int i = 0; i < 64; ++i)
    {
        Float dx, dy;
        DiskSample(dx, dy);
        if (sc.z <= Lookup(map, sc.x + dx * size, sc.y + dy * size) + bias) visibility += 1.f / 64.f;
    }
    return visibility;
}
}  // namespace

void ResetHostSampleStream() { g_Generator = std::mt19937(); }

Float CalculateShadowVisibility(const Buffer1f& shadowMap, const Vector3f& positionLightSpaceNDC, const Vector3f& normal, const Vector3f& lightDir)
{
    const Vector3f sc = positionLightSpaceNDC * 0.5f + Vector3f(0.5f);
    const Float    bias = std::max(0.009f * (1.f - Dot(normal, lightDir)), 0.007f);
    const double   kPcfFilter = 0.007, kSearchFilter = 0.005;
    if (GetShadowMode() == Hard) return sc.z <= Lookup(shadowMap, sc.x, sc.y) + bias ? 1.f : 0.f;
    if (GetShadowMode() == PCF) return Filter(shadowMap, sc, bias, (Float)kPcfFilter);
    // PCSS: 32-tap blocker search, penumbra from the average blocker depth, then the 64-tap filter
    Float sum = 0.f, blockers = 0.f;
    for (int i = 0; i < 32; ++i)
    {
        Float dx, dy;
        DiskSample(dx, dy);
        const Float d = Lookup(shadowMap, sc.x + (Float)((double)dx * kSearchFilter), sc.y + (Float)((double)dy * kSearchFilter));
        if (sc.z > d + bias) sum += d, blockers += 1.f;
    }
    const Float dBlocker = blockers < 1.f ? 0.f : sum / blockers;
    if ((double)dBlocker < 0.001) return 1.f;
    const Float penumbra = (sc.z - dBlocker) * 2.5f / dBlocker;
    return Filter(shadowMap, sc, bias, (Float)(kPcfFilter * (double)penumbra));
}
}  // namespace Shadow
