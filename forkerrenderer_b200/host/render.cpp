// render.cpp — frame orchestration: the pass sequence, matrices and per-model shader set-up of reference
// src/render.cpp:33-343, issuing device work through the ForkerGL facade.  Constants: render.cpp:9-14.
#include "render.h"

#include "forkergl.h"
#include "forkergl_b200.h"
#include "shader.h"
#include "shadow.h"

static const Float s_ShadowViewSize = 3.0f;
static const Float s_ShadowNearPlane = 0.1f;
static const Float s_ShadowFarPlane = 20.f;
static const Float s_CameraNearPlane = 0.01f;
static const Float s_CameraFarPlane = 20.f;

namespace Render
{
static int BufferWidth(const Scene& s) { return s.IsSSAAOn() ? s.GetWidth() * s.GetSSAAKernelSize() : s.GetWidth(); }
static int BufferHeight(const Scene& s) { return s.IsSSAAOn() ? s.GetHeight() * s.GetSSAAKernelSize() : s.GetHeight(); }

static Matrix4x4f CameraProjection(const Scene& scene)
{
    Float ratio = scene.GetRatio();
    return (scene.GetProjectionType() == Camera::Orthographic)
               ? scene.GetCamera().GetOrthographicMatrix(-1.f * ratio, 1.f * ratio, -1.f, 1.f, s_CameraNearPlane,
                                                         s_CameraFarPlane)
               : scene.GetCamera().GetPerspectiveMatrix(45.f, ratio, s_CameraNearPlane, s_CameraFarPlane);
}

// reference render.cpp:33-38
void Preconfigure(const Scene& scene)
{
    ForkerGL::SetViewportMatrix(0, 0, BufferWidth(scene), BufferHeight(scene));
    ForkerGL::TextureWrapMode(Texture::NoWrap);
    ForkerGL::TextureFilterMode(Texture::Nearest);
}

// reference render.cpp:40-58, split where a sort-first multi-GPU driver has to exchange the PCSS chain state
// (fgl_set_chain_blockers_before) between the bands: everything up to the lighting loop, then the rest.
void RenderGeometryStage(const Scene& scene)
{
    fgl_ctx* ctx = ForkerGL::Context();
    FglParams params = ForkerGL::Params();
    params.shadow_mode = (int)Shadow::GetShadowMode();
    ForkerGL::Check(fgl_set_params(ctx, &params), "params");
    ForkerGL::Check(fgl_set_shadow_status(ctx, Shadow::GetShadowStatus() ? 1 : 0), "shadow status");
    ForkerGL::Check(fgl_begin_frame(ctx), "begin frame");

    DoShadowPass(scene);
    if (ForkerGL::GetRenderMode() == ForkerGL::Forward)
        DoForwardPass(scene);
    else
    {
        DoGeometryPass(scene);
        // first half of DoLightingPass (reference render.cpp:195-209)
        ForkerGL::InitFrameBuffer(BufferWidth(scene), BufferHeight(scene));
        ForkerGL::SetPassType(ForkerGL::LightingPass);
        ForkerGL::ClearColor(Color3(0.12f, 0.12f, 0.12f));  // overwritten by the lighting loop, as in the reference
        // nothing but a head start for the lighting loop: a PCSS frame's sample-stream chain starts here, on its own stream,
        // and runs while SSAO and the blur do (include/forkergl_b200.h, fgl_prepare_screen_space_pixels)
        ForkerGL::PrepareScreenSpacePixels(scene, scene.IsSSAOOn());
        if (scene.IsSSAOOn()) DoSSAO(scene);
    }
    ForkerGL::FlushTriangles();  // per-triangle submission: nothing stays behind in the host-side batch
}

void RenderLightingStage(const Scene& scene)
{
    if (ForkerGL::GetRenderMode() != ForkerGL::Forward) ForkerGL::DrawScreenSpacePixels(scene);  // render.cpp:210
    DoSSAA(scene);
    ForkerGL::FlushTriangles();
}

void Render(const Scene& scene)
{
    RenderGeometryStage(scene);
    RenderLightingStage(scene);
}

// reference render.cpp:60-97
void DoShadowPass(const Scene& scene)
{
    if (!Shadow::GetShadowStatus()) return;
    ForkerGL::InitShadowBuffer(BufferWidth(scene), BufferHeight(scene));
    ForkerGL::InitDepthBuffer(BufferWidth(scene), BufferHeight(scene));
    ForkerGL::SetPassType(ForkerGL::ShadowPass);

    Matrix4x4f viewSM = MakeLookAtMatrix(scene.GetPointLight().position, Vector3f(0.f));
    Matrix4x4f projSM = MakeOrthographicMatrix(-s_ShadowViewSize * scene.GetRatio(), s_ShadowViewSize * scene.GetRatio(),
                                               -s_ShadowViewSize, s_ShadowViewSize, s_ShadowNearPlane, s_ShadowFarPlane);
    ForkerGL::SetLightSpaceMatrix(projSM * viewSM);

    for (int i = 0; i < (int)scene.GetModelCount(); ++i)
    {
        DepthShader depthShader;
        depthShader.uModelMatrix = scene.GetModelMatrix(i);
        depthShader.uLightSpaceMatrix = ForkerGL::GetLightSpaceMatrix();
        scene.GetModel(i).Render(depthShader);
    }
}

// reference render.cpp:99-157
void DoForwardPass(const Scene& scene)
{
    ForkerGL::InitFrameBuffer(BufferWidth(scene), BufferHeight(scene));
    ForkerGL::InitDepthBuffer(BufferWidth(scene), BufferHeight(scene));
    ForkerGL::ClearColor(Color3(0.12f, 0.12f, 0.12f));
    ForkerGL::SetPassType(ForkerGL::ForwardPass);

    Matrix4x4f viewMatrix = scene.GetCamera().GetViewMatrix();
    Matrix4x4f projectionMatrix = CameraProjection(scene);
    ForkerGL::SetViewProjectionMatrix(projectionMatrix * viewMatrix);

    for (int i = 0; i < (int)scene.GetModelCount(); ++i)
    {
        const Model& model = scene.GetModel(i);
        if (!model.SupportPBR())
        {
            BlinnPhongShader s;
            s.uModelMatrix = scene.GetModelMatrix(i);
            s.uViewMatrix = viewMatrix;
            s.uProjectionMatrix = projectionMatrix;
            s.uNormalMatrix = MakeNormalMatrix(s.uModelMatrix);
            s.uPointLight = scene.GetPointLight();
            s.uEyePos = scene.GetCamera().GetPosition();
            if (Shadow::GetShadowStatus()) s.uLightSpaceMatrix = ForkerGL::GetLightSpaceMatrix();
            model.Render(s);
        }
        else
        {
            PBRShader s;
            s.uModelMatrix = scene.GetModelMatrix(i);
            s.uViewMatrix = viewMatrix;
            s.uProjectionMatrix = projectionMatrix;
            s.uNormalMatrix = MakeNormalMatrix(s.uModelMatrix);
            s.uPointLight = scene.GetPointLight();
            s.uEyePos = scene.GetCamera().GetPosition();
            if (Shadow::GetShadowStatus()) s.uLightSpaceMatrix = ForkerGL::GetLightSpaceMatrix();
            model.Render(s);
        }
    }
}

// reference render.cpp:159-193
void DoGeometryPass(const Scene& scene)
{
    ForkerGL::InitGeometryBuffers(BufferWidth(scene), BufferHeight(scene));
    ForkerGL::InitDepthBuffer(BufferWidth(scene), BufferHeight(scene));
    ForkerGL::SetPassType(ForkerGL::GeometryPass);

    Matrix4x4f viewMatrix = scene.GetCamera().GetViewMatrix();
    Matrix4x4f projectionMatrix = CameraProjection(scene);
    ForkerGL::SetViewProjectionMatrix(projectionMatrix * viewMatrix);

    for (int i = 0; i < (int)scene.GetModelCount(); ++i)
    {
        GShader g;
        g.uModelMatrix = scene.GetModelMatrix(i);
        g.uViewMatrix = viewMatrix;
        g.uProjectionMatrix = projectionMatrix;
        g.uNormalMatrix = MakeNormalMatrix(g.uModelMatrix);
        g.uLightSpaceMatrix = ForkerGL::GetLightSpaceMatrix();
        scene.GetModel(i).Render(g);
    }
}

// reference render.cpp:195-212
void DoLightingPass(const Scene& scene)
{
    ForkerGL::InitFrameBuffer(BufferWidth(scene), BufferHeight(scene));
    ForkerGL::SetPassType(ForkerGL::LightingPass);
    ForkerGL::ClearColor(Color3(0.12f, 0.12f, 0.12f));  // overwritten by the lighting loop, as in the reference
    if (scene.IsSSAOOn()) DoSSAO(scene);
    ForkerGL::DrawScreenSpacePixels(scene);
}

// reference render.cpp:214-289
void DoSSAO(const Scene&)
{
    ForkerGL::InvalidateHostMirrors();
    ForkerGL::Check(fgl_ssao(ForkerGL::Context()), "SSAO");
    ForkerGL::AmbientOcclusionGBuffer.TwoPassGaussianBlurDenoised();
}

// reference render.cpp:291-343
void DoSSAA(const Scene& scene)
{
    if (!scene.IsSSAAOn()) return;
    ForkerGL::InvalidateHostMirrors();
    ForkerGL::Check(fgl_ssaa_resolve(ForkerGL::Context(), scene.GetSSAAKernelSize()), "SSAA");
}
}  // namespace Render
