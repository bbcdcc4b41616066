// shader.h — Shader interface of the drop-in facade (reference src/shaders/shader.h:18-31) plus the four
// programs (depthshader.h, gshader.h, phongshader.h, pbrshader.h) with their public uniform fields preserved.
//
// Virtual ProcessVertex/ProcessFragment callbacks cannot run on the device: the four known programs execute as
// CUDA code (csrc/programs.cuh) selected by Kind().  User-defined Shader subclasses are outside the drop-in
// surface (stated limitation): Kind() of an unknown subclass is -1 and drawing with it
// is an error.
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

    // FGL_SHADER_* of the device program that implements this class, -1 if none.
    virtual int Kind() const { return -1; }
    // Packs the public uniform fields for the device program.
    virtual void FillUniforms(FglUniforms& u) const = 0;
};

struct DepthShader : public Shader
{
    Matrix4x4f uModelMatrix;
    Matrix4x4f uLightSpaceMatrix;
    int  Kind() const override;
    void FillUniforms(FglUniforms& u) const override;
};

struct GShader : public Shader
{
    Matrix4x4f uModelMatrix, uViewMatrix, uProjectionMatrix;
    Matrix3x3f uNormalMatrix;
    Matrix4x4f uLightSpaceMatrix;
    int  Kind() const override;
    void FillUniforms(FglUniforms& u) const override;
};

struct BlinnPhongShader : public Shader
{
    Matrix4x4f uModelMatrix, uViewMatrix, uProjectionMatrix;
    Matrix3x3f uNormalMatrix;
    PointLight uPointLight;
    Point3f    uEyePos;
    Matrix4x4f uLightSpaceMatrix;
    int  Kind() const override;
    void FillUniforms(FglUniforms& u) const override;
};

struct PBRShader : public Shader
{
    Matrix4x4f uModelMatrix, uViewMatrix, uProjectionMatrix;
    Matrix3x3f uNormalMatrix;
    PointLight uPointLight;
    Point3f    uEyePos;
    Matrix4x4f uLightSpaceMatrix;
    int  Kind() const override;
    void FillUniforms(FglUniforms& u) const override;
};
