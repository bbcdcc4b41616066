// shader.cpp — packs the Shader subclasses' public uniform fields into the C-ABI struct.
#include "shader.h"

#include <cstring>

#include "forkergl_b200.h"

namespace
{
void Put(float* dst, const Matrix4x4f& m)
{
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) dst[r * 4 + c] = m[r][c];
}
void Put(float* dst, const Matrix3x3f& m)
{
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) dst[r * 3 + c] = m[r][c];
}
void Put(float* dst, const Vector3f& v) { dst[0] = v.x, dst[1] = v.y, dst[2] = v.z; }
}  // namespace

int DepthShader::Kind() const { return FGL_SHADER_DEPTH; }
void DepthShader::FillUniforms(FglUniforms& u) const
{
    memset(&u, 0, sizeof u);
    Put(u.model, uModelMatrix);
    Put(u.light_space, uLightSpaceMatrix);
}

int GShader::Kind() const { return FGL_SHADER_G; }
void GShader::FillUniforms(FglUniforms& u) const
{
    memset(&u, 0, sizeof u);
    Put(u.model, uModelMatrix);
    Put(u.view, uViewMatrix);
    Put(u.projection, uProjectionMatrix);
    Put(u.normal, uNormalMatrix);
    Put(u.light_space, uLightSpaceMatrix);
}

int BlinnPhongShader::Kind() const { return FGL_SHADER_BLINN_PHONG; }
void BlinnPhongShader::FillUniforms(FglUniforms& u) const
{
    memset(&u, 0, sizeof u);
    Put(u.model, uModelMatrix);
    Put(u.view, uViewMatrix);
    Put(u.projection, uProjectionMatrix);
    Put(u.normal, uNormalMatrix);
    Put(u.light_space, uLightSpaceMatrix);
    Put(u.light_position, uPointLight.position);
    Put(u.light_color, uPointLight.color);
    Put(u.eye_position, uEyePos);
}

int PBRShader::Kind() const { return FGL_SHADER_PBR; }
void PBRShader::FillUniforms(FglUniforms& u) const
{
    memset(&u, 0, sizeof u);
    Put(u.model, uModelMatrix);
    Put(u.view, uViewMatrix);
    Put(u.projection, uProjectionMatrix);
    Put(u.normal, uNormalMatrix);
    Put(u.light_space, uLightSpaceMatrix);
    Put(u.light_position, uPointLight.position);
    Put(u.light_color, uPointLight.color);
    Put(u.eye_position, uEyePos);
}
