// shader.h — Shader interface of the drop-in facade (reference src/shaders/shader.h:18-31) plus the four
// programs (depthshader.h, gshader.h, phongshader.h, pbrshader.h) with their public uniform fields preserved.
//
// Two ways to the device:
//   * Mesh::Draw -> ForkerGL::DrawMesh: one indexed draw, the vertex AND fragment programs run as CUDA code
//     (csrc/raster.cu k_setup / k_resolve_*) selected by Kind().  This is the fast path.
//   * the reference's own per-face loop (mesh.cpp:10-25): shader.Use, 3 x ProcessVertex, ForkerGL::DrawTriangle.
//     ProcessVertex below runs the vertex program on the HOST (same expressions, same order as the device program, so the
//     values are the same bits) and leaves the triangle's varyings in `varyings` / `lightZ`; DrawTriangle hands them to the
//     device in batches (fgl_draw_triangles), where the rasteriser and the fragment programs run.
// ProcessFragment is the host-side fragment program (host/programs.cpp): it is what a caller that drives the interface by hand
// gets; no device pass calls it.  User-defined Shader subclasses can be written against this interface, but only the four
// programs below have device counterparts: Kind() of an unknown subclass is -1 and drawing with it is an error (stated
// limitation).
#pragma once

#include <memory>

#include "geometry.h"
#include "light.h"
#include "mesh.h"

struct FglUniforms;

struct Shader
{
    std::shared_ptr<const Mesh> mesh;

    Shader() : mesh(nullptr) {}
    virtual ~Shader() {}

    void Use(std::shared_ptr<const Mesh> m) { mesh = m; }

    virtual Point4f ProcessVertex(int faceIdx, int vertIdx) = 0;                        // reference shader.h:28
    virtual bool    ProcessFragment(const Vector3f& baryCoord, Color3& gl_Color) = 0;   // reference shader.h:30

    // FGL_SHADER_* of the device program that implements this class, -1 if none.
    virtual int Kind() const { return -1; }
    // Packs the public uniform fields for the device program.
    virtual void FillUniforms(FglUniforms& u) const = 0;

    // The current triangle's varyings as ProcessVertex left them, in the layout of fgl_draw_triangles
    // (include/forkergl_b200.h): camera programs fill `varyings`, DepthShader fills `lightZ`.
    float varyings[48] = { 0 };
    float lightZ[3] = { 0, 0, 0 };
};

struct DepthShader : public Shader
{
    Matrix4x4f uModelMatrix;
    Matrix4x4f uLightSpaceMatrix;
    Point4f ProcessVertex(int faceIdx, int vertIdx) override;
    bool    ProcessFragment(const Vector3f& baryCoord, Color3& gl_Color) override;
    int     Kind() const override;
    void    FillUniforms(FglUniforms& u) const override;
};

struct GShader : public Shader
{
    Matrix4x4f uModelMatrix, uViewMatrix, uProjectionMatrix;
    Matrix3x3f uNormalMatrix;
    Matrix4x4f uLightSpaceMatrix;
    // outputs of ProcessFragment (reference gshader.h:33-40; ForkerGL::DrawTriangle writes them to the G-buffers)
    Vector3f outNormalWS, outPositionWS, outLightSpaceNDC;
    Color3   outAlbedo, outEmissive;
    Vector3f outParam;
    Float    outShadingType = 0.f;
    Point4f ProcessVertex(int faceIdx, int vertIdx) override;
    bool    ProcessFragment(const Vector3f& baryCoord, Color3& gl_Color) override;
    int     Kind() const override;
    void    FillUniforms(FglUniforms& u) const override;
};

struct BlinnPhongShader : public Shader
{
    Matrix4x4f uModelMatrix, uViewMatrix, uProjectionMatrix;
    Matrix3x3f uNormalMatrix;
    PointLight uPointLight;
    Point3f    uEyePos;
    Matrix4x4f uLightSpaceMatrix;
    Point4f ProcessVertex(int faceIdx, int vertIdx) override;
    bool    ProcessFragment(const Vector3f& baryCoord, Color3& gl_Color) override;
    int     Kind() const override;
    void    FillUniforms(FglUniforms& u) const override;
    // reference phongshader.h:171-215
    static Color3 CalculateLight(const Vector3f& lightDir, const Vector3f& halfwayDir, const Vector3f& normal, Float visibility,
                                 const Color3& diffuseColor, const Color3& emissive, const Vector3f& param, const Color3& lightColor);
};

struct PBRShader : public Shader
{
    Matrix4x4f uModelMatrix, uViewMatrix, uProjectionMatrix;
    Matrix3x3f uNormalMatrix;
    PointLight uPointLight;
    Point3f    uEyePos;
    Matrix4x4f uLightSpaceMatrix;
    Point4f ProcessVertex(int faceIdx, int vertIdx) override;
    bool    ProcessFragment(const Vector3f& baryCoord, Color3& gl_Color) override;
    int     Kind() const override;
    void    FillUniforms(FglUniforms& u) const override;
    // reference pbrshader.h:182-288
    static Color3 CalculateLight(const Vector3f& lightDir, const Vector3f& viewDir, const Vector3f& halfwayDir, const Vector3f& normal,
                                 Float visibility, const Color3& albedo, const Color3& emissive, const Vector3f& param,
                                 const Color3& lightRadiance);
};
