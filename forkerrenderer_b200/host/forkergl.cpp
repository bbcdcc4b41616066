// forkergl.cpp — ForkerGL facade over the C ABI.  Mirrors reference src/forkergl.cpp:44-142 (state setters),
// :55-81 (Init*), :326-380 (DrawScreenSpacePixels entry); the work itself is CUDA (csrc/).
#include "forkergl.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "forkergl_b200.h"
#include "light.h"
#include "scene.h"
#include "shader.h"
#include "shadow.h"

Texture::WrapMode   ForkerGL::TextureWrapping = Texture::NoWrap;
Texture::FilterMode ForkerGL::TextureFiltering = Texture::Nearest;

#define FGL_FACADE_DEFINE(TYPE, NAME, PLANE) TYPE ForkerGL::NAME;
FGL_FACADE_BUFFERS(FGL_FACADE_DEFINE)
#undef FGL_FACADE_DEFINE
TGAImage ForkerGL::AntiAliasedImage;

static fgl_ctx*             s_Ctx = nullptr;
static ForkerGL::RenderMode s_RenderMode = ForkerGL::Forward;

// ---- DrawTriangle batching -------------------------------------------------------------------------------------------
// Consecutive DrawTriangle calls with the same mesh, program and uniforms form one batch; anything else the caller does
// goes through Context(), which flushes the batch first, so the device sees the triangles in call order.
namespace
{
struct TriangleBatch
{
    int                meshId = -1, kind = -1;
    FglUniforms        uniforms;
    std::vector<float> ndc, varyings, lightZ;
    bool               open = false;
} s_Batch;
bool s_PerTriangle = false, s_Flushing = false;
}  // namespace

void ForkerGL::SetPerTriangleSubmission(bool on) { s_PerTriangle = on; }
bool ForkerGL::GetPerTriangleSubmission() { return s_PerTriangle; }

void ForkerGL::FlushTriangles()
{
    if (!s_Batch.open || s_Flushing) return;
    s_Flushing = true;
    s_Batch.open = false;
    const int n = (int)(s_Batch.ndc.size() / 12);
    int       rc = fgl_draw_triangles(s_Ctx, s_Batch.meshId, s_Batch.kind, &s_Batch.uniforms, n, s_Batch.ndc.data(),
                                      s_Batch.varyings.empty() ? nullptr : s_Batch.varyings.data(), s_Batch.lightZ.empty() ? nullptr : s_Batch.lightZ.data());
    s_Batch.ndc.clear(), s_Batch.varyings.clear(), s_Batch.lightZ.clear();
    s_Flushing = false;
    Check(rc, "DrawTriangle");
}

void ForkerGL::DrawTriangle(const Point4f ndcVerts[3], Shader& shader)
{
    if (shader.Kind() < 0)
        throw std::runtime_error("ForkerGL::DrawTriangle: only DepthShader, GShader, BlinnPhongShader and PBRShader have device "
                                 "fragment programs (user-defined Shader subclasses are unsupported)");
    if (!shader.mesh) throw std::runtime_error("ForkerGL::DrawTriangle: shader.Use(mesh) was not called");
    FglUniforms u;
    shader.FillUniforms(u);
    const int meshId = shader.mesh->DeviceId(), kind = shader.Kind();  // (may upload the model: goes through Context())
    if (s_Batch.open && (s_Batch.meshId != meshId || s_Batch.kind != kind || memcmp(&s_Batch.uniforms, &u, sizeof u) != 0)) FlushTriangles();
    if (!s_Batch.open)
    {
        Context();
        s_Batch.open = true, s_Batch.meshId = meshId, s_Batch.kind = kind, s_Batch.uniforms = u;
        InvalidateHostMirrors();
    }
    for (int v = 0; v < 3; ++v)
    {
        const float q[4] = { ndcVerts[v].x, ndcVerts[v].y, ndcVerts[v].z, ndcVerts[v].w };
        s_Batch.ndc.insert(s_Batch.ndc.end(), q, q + 4);
    }
    if (kind == FGL_SHADER_DEPTH) s_Batch.lightZ.insert(s_Batch.lightZ.end(), shader.lightZ, shader.lightZ + 3);
    else s_Batch.varyings.insert(s_Batch.varyings.end(), shader.varyings, shader.varyings + 48);
}

fgl_ctx* ForkerGL::Context()
{
    if (s_Ctx && s_Batch.open && !s_Flushing) FlushTriangles();
    if (!s_Ctx)
    {
        const char* dev = getenv("FGL_DEVICE");
        int         rc = fgl_create(dev ? atoi(dev) : 0, &s_Ctx);
        if (rc != FGL_OK)
        {
            // No CPU fallback: without the CUDA backend the product path stops here, loudly.
            std::string msg = std::string("ForkerGL: cannot create the CUDA context: ") + fgl_last_error(nullptr);
            fprintf(stderr, "%s\n", msg.c_str());
            throw std::runtime_error(msg);
        }
    }
    return s_Ctx;
}

FglParams& ForkerGL::Params()
{
    static FglParams p;
    static bool      init = false;
    if (!init)
    {
        fgl_default_params(&p);
        init = true;
    }
    return p;
}

void ForkerGL::Shutdown()
{
    s_Batch = TriangleBatch();
    if (s_Ctx) fgl_destroy(s_Ctx);
    s_Ctx = nullptr;
}

void ForkerGL::Check(int status, const char* what)
{
    if (status != FGL_OK)
        throw std::runtime_error(std::string(what) + ": " + fgl_last_error(s_Ctx));
}

void ForkerGL::InvalidateHostMirrors()
{
#define FGL_FACADE_ADDRESS(TYPE, NAME, PLANE) &NAME,
    Buffer* all[] = { FGL_FACADE_BUFFERS(FGL_FACADE_ADDRESS) };
#undef FGL_FACADE_ADDRESS
    for (Buffer* b : all)
    {
        b->push();  // host edits made through SetValue reach the device before it runs
        b->m_HostValid = false;
    }
}

void ForkerGL::TextureWrapMode(Texture::WrapMode wrapMode) { TextureWrapping = wrapMode; }
void ForkerGL::TextureFilterMode(Texture::FilterMode filterMode) { TextureFiltering = filterMode; }

void ForkerGL::InitFrameBuffer(int width, int height)
{
    Check(fgl_init_frame_buffer(Context(), width, height), "InitFrameBuffer");
    FrameBuffer = Buffer3f(width, height, FGL_PLANE_FRAME);
}

void ForkerGL::InitDepthBuffer(int width, int height)
{
    Check(fgl_init_depth_buffer(Context(), width, height), "InitDepthBuffer");
    DepthBuffer = Buffer1f(width, height, FGL_PLANE_DEPTH);
}

void ForkerGL::InitShadowBuffer(int width, int height)
{
    Check(fgl_init_shadow_buffer(Context(), width, height), "InitShadowBuffer");
    ShadowBuffer = Buffer1f(width, height, FGL_PLANE_SHADOW);
}

void ForkerGL::InitGeometryBuffers(int width, int height)
{
    Check(fgl_init_geometry_buffers(Context(), width, height), "InitGeometryBuffers");
    NormalGBuffer = Buffer3f(width, height, FGL_PLANE_NORMAL);
    WorldPosGBuffer = Buffer3f(width, height, FGL_PLANE_WORLDPOS);
    if (Shadow::GetShadowStatus()) LightSpaceNDCPosGBuffer = Buffer3f(width, height, FGL_PLANE_LIGHTNDC);
    AlbedoGBuffer = Buffer3f(width, height, FGL_PLANE_ALBEDO);
    EmissiveGBuffer = Buffer3f(width, height, FGL_PLANE_EMISSIVE);
    ParamGBuffer = Buffer3f(width, height, FGL_PLANE_PARAM);
    ShadingTypeGBuffer = Buffer1f(width, height, FGL_PLANE_SHADINGTYPE);
    AmbientOcclusionGBuffer = Buffer1f(width, height, FGL_PLANE_AO);
}

void ForkerGL::ClearColor(const Color3& color)
{
    float rgb[3] = { color.x, color.y, color.z };
    Check(fgl_clear_color(Context(), rgb), "ClearColor");
    FrameBuffer.m_HostValid = false;
}

void ForkerGL::SetViewportMatrix(int x, int y, int w, int h)
{
    Check(fgl_set_viewport(Context(), x, y, w, h), "SetViewportMatrix");
}

static Matrix4x4f FromFlat(const float* f)
{
    Matrix4x4f m;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) m[r][c] = f[r * 4 + c];
    return m;
}
static void ToFlat(const Matrix4x4f& m, float* f)
{
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) f[r * 4 + c] = m[r][c];
}

Matrix4x4f ForkerGL::GetViewportMatrix()
{
    float f[16];
    Check(fgl_get_viewport_matrix(Context(), f), "GetViewportMatrix");
    return FromFlat(f);
}

void ForkerGL::SetViewProjectionMatrix(const Matrix4x4f& matrix)
{
    float f[16];
    ToFlat(matrix, f);
    Check(fgl_set_view_projection_matrix(Context(), f), "SetViewProjectionMatrix");
}

Matrix4x4f ForkerGL::GetViewProjectionMatrix()
{
    float f[16];
    Check(fgl_get_view_projection_matrix(Context(), f), "GetViewProjectionMatrix");
    return FromFlat(f);
}

void ForkerGL::SetLightSpaceMatrix(const Matrix4x4f& matrix)
{
    float f[16];
    ToFlat(matrix, f);
    Check(fgl_set_light_space_matrix(Context(), f), "SetLightSpaceMatrix");
}

Matrix4x4f ForkerGL::GetLightSpaceMatrix()
{
    float f[16];
    Check(fgl_get_light_space_matrix(Context(), f), "GetLightSpaceMatrix");
    return FromFlat(f);
}

void ForkerGL::SetRenderMode(enum RenderMode mode)
{
    s_RenderMode = mode;
    if (s_Ctx) fgl_set_render_mode(s_Ctx, (int)mode);
}

ForkerGL::RenderMode ForkerGL::GetRenderMode() { return s_RenderMode; }

void ForkerGL::SetPassType(enum PassType type)
{
    Check(fgl_set_render_mode(Context(), (int)s_RenderMode), "SetRenderMode");
    Check(fgl_set_pass_type(Context(), (int)type), "SetPassType");
}

void ForkerGL::DrawMesh(const Mesh& mesh, Shader& shader)
{
    if (shader.Kind() < 0)
        throw std::runtime_error("ForkerGL::DrawMesh: only DepthShader, GShader, BlinnPhongShader and PBRShader "
                                 "have device programs (user-defined Shader subclasses are unsupported)");
    FglUniforms u;
    shader.FillUniforms(u);
    Check(fgl_draw_mesh(Context(), mesh.DeviceId(), shader.Kind(), &u), "DrawMesh");
    InvalidateHostMirrors();
}

void ForkerGL::DrawScreenSpacePixels(const Scene& scene)
{
    Point3f eye = scene.GetCamera().GetPosition();
    Point3f lp = scene.GetPointLight().position;
    Color3  lc = scene.GetPointLight().color;
    float   e[3] = { eye.x, eye.y, eye.z }, p[3] = { lp.x, lp.y, lp.z }, c[3] = { lc.x, lc.y, lc.z };
    InvalidateHostMirrors();
    Check(fgl_draw_screen_space_pixels(Context(), e, p, c), "DrawScreenSpacePixels");
}

// sort-first drivers: the band-independent first half of DrawScreenSpacePixels (include/forkergl_b200.h)
void ForkerGL::PrepareScreenSpacePixels(const Scene& scene, bool ssaoFollows)
{
    Point3f eye = scene.GetCamera().GetPosition();
    Point3f lp = scene.GetPointLight().position;
    Color3  lc = scene.GetPointLight().color;
    float   e[3] = { eye.x, eye.y, eye.z }, p[3] = { lp.x, lp.y, lp.z }, c[3] = { lc.x, lc.y, lc.z };
    InvalidateHostMirrors();
    Check(fgl_prepare_screen_space_pixels(Context(), e, p, c, ssaoFollows ? 1 : 0), "PrepareScreenSpacePixels");
}

void ForkerGL::FetchAntiAliasedImage()
{
    int w = 0, h = 0, ch = 0, bpc = 0;
    Check(fgl_plane_info(Context(), FGL_PLANE_SSAA_RGB8, &w, &h, &ch, &bpc), "plane_info");
    if (w == 0 || h == 0) return;
    std::vector<std::uint8_t> rgb((size_t)w * h * 3);
    Check(fgl_read_plane(Context(), FGL_PLANE_SSAA_RGB8, rgb.data(), rgb.size()), "read SSAA image");
    TGAImage img(w, h, TGAImage::RGB);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
        {
            const std::uint8_t* s = &rgb[((size_t)x + (size_t)y * w) * 3];
            img.Set(x, y, TGAColor(s[0], s[1], s[2]));
        }
    AntiAliasedImage = img;
}
