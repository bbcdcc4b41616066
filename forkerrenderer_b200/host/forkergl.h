// forkergl.h — struct ForkerGL of the drop-in facade: the same static state / draw-call surface as reference
// src/forkergl.h:15-81, implemented as thin calls into the C ABI (include/forkergl_b200.h).  Nothing else in
// the host code touches CUDA.
#pragma once

#include "buffer.h"
#include "geometry.h"
#include "texture.h"
#include "tgaimage.h"

class Scene;
class Mesh;
struct Shader;
struct fgl_ctx;
struct FglParams;

struct ForkerGL
{
    ForkerGL() = delete;

    enum RenderMode { Forward, Deferred };
    enum PassType { ForwardPass, GeometryPass, LightingPass, ShadowPass };

    static Texture::WrapMode   TextureWrapping;
    static Texture::FilterMode TextureFiltering;
    static void TextureWrapMode(Texture::WrapMode wrapMode);
    static void TextureFilterMode(Texture::FilterMode filterMode);

    // The static buffers of the reference (forkergl.h:34-47), here handles onto device planes (buffer.h).  One list feeds the
    // declarations, the definitions and the host-mirror bookkeeping in forkergl.cpp:  X(type, name, C-ABI plane).
#define FGL_FACADE_BUFFERS(X)                                      \
    X(Buffer3f, FrameBuffer, FGL_PLANE_FRAME)                      \
    X(Buffer1f, DepthBuffer, FGL_PLANE_DEPTH)                      \
    X(Buffer1f, ShadowBuffer, FGL_PLANE_SHADOW)                    \
    X(Buffer3f, NormalGBuffer, FGL_PLANE_NORMAL)                   \
    X(Buffer3f, WorldPosGBuffer, FGL_PLANE_WORLDPOS)               \
    X(Buffer3f, LightSpaceNDCPosGBuffer, FGL_PLANE_LIGHTNDC)       \
    X(Buffer3f, AlbedoGBuffer, FGL_PLANE_ALBEDO)                   \
    X(Buffer3f, EmissiveGBuffer, FGL_PLANE_EMISSIVE)               \
    X(Buffer3f, ParamGBuffer, FGL_PLANE_PARAM)                     \
    X(Buffer1f, ShadingTypeGBuffer, FGL_PLANE_SHADINGTYPE)         \
    X(Buffer1f, AmbientOcclusionGBuffer, FGL_PLANE_AO)
#define FGL_FACADE_DECLARE(TYPE, NAME, PLANE) static TYPE NAME;
    FGL_FACADE_BUFFERS(FGL_FACADE_DECLARE)
#undef FGL_FACADE_DECLARE
    static TGAImage AntiAliasedImage;  // the SSAA result (render.cpp:291-343), fetched from the device on demand

    static void InitFrameBuffer(int width, int height);
    static void InitDepthBuffer(int width, int height);
    static void InitShadowBuffer(int width, int height);
    static void InitGeometryBuffers(int width, int height);

    static void       ClearColor(const Color3& color);
    static void       SetViewportMatrix(int x, int y, int w, int h);
    static Matrix4x4f GetViewportMatrix();
    static void       SetViewProjectionMatrix(const Matrix4x4f& matrix);
    static Matrix4x4f GetViewProjectionMatrix();
    static void       SetLightSpaceMatrix(const Matrix4x4f& matrix);
    static Matrix4x4f GetLightSpaceMatrix();
    static void       SetRenderMode(enum RenderMode mode);
    static RenderMode GetRenderMode();
    static void       SetPassType(enum PassType type);

    // Rasterization.  DrawTriangle is the reference's entry (forkergl.h:74): one triangle whose three NDC positions the caller
    // obtained from shader.ProcessVertex; triangles are batched and reach the device at the next state change, draw of another
    // kind or read-back (fgl_draw_triangles), primitive ids in call order.  DrawMesh is what Mesh::Draw calls by default: one
    // indexed draw per mesh, vertex programs on the device.
    static void DrawTriangle(const Point4f ndcVerts[3], Shader& shader);
    static void DrawMesh(const Mesh& mesh, Shader& shader);
    // Mesh::Draw submits face by face through ProcessVertex + DrawTriangle (the reference's loop, mesh.cpp:10-25) instead of
    // one DrawMesh per mesh.  Same image either way; off by default.
    static void SetPerTriangleSubmission(bool on);
    static bool GetPerTriangleSubmission();
    static void FlushTriangles();  // hands a pending DrawTriangle batch to the device (implied by every other entry point)
    static void DrawScreenSpacePixels(const Scene& scene);
    static void PrepareScreenSpacePixels(const Scene& scene, bool ssaoFollows);  // optional, multi-GPU: see fgl_prepare_screen_space_pixels

    // --- additions over the reference surface (device plumbing) ---
    static fgl_ctx* Context();            // lazily created on device $FGL_DEVICE (default 0); aborts loudly on failure
    static void     Shutdown();
    static void     Check(int status, const char* what);  // throws std::runtime_error with fgl_last_error
    static void     InvalidateHostMirrors();              // called after every device pass
    static void     FetchAntiAliasedImage();              // device SSAA image -> AntiAliasedImage
    static struct FglParams& Params();                    // constants pushed to the device at Render::Render
};
