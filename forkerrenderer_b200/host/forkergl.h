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

    // Buffers (handles onto device planes; see buffer.h)
    static Buffer3f FrameBuffer;
    static Buffer1f DepthBuffer;
    static Buffer1f ShadowBuffer;
    static Buffer3f NormalGBuffer;
    static Buffer3f WorldPosGBuffer;
    static Buffer3f LightSpaceNDCPosGBuffer;
    static Buffer3f AlbedoGBuffer;
    static Buffer3f EmissiveGBuffer;
    static Buffer3f ParamGBuffer;
    static Buffer1f ShadingTypeGBuffer;
    static Buffer1f AmbientOcclusionGBuffer;
    static TGAImage AntiAliasedImage;

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

    // Rasterization.  DrawMesh is what Mesh::Draw calls (one indexed draw per mesh).
    static void DrawMesh(const Mesh& mesh, Shader& shader);
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
