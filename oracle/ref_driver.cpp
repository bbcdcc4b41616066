// oracle/ref_driver.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A replacement for the reference's src/main.cpp (reference main.cpp:18-55) that links against the
// UNMODIFIED reference objects and, instead of writing 8-bit TGAs, dumps every ForkerGL buffer as raw
// little-endian fp32 (plus the 8-bit images) so that parity can be checked bit-for-bit.  It calls only the
// reference's public API: Scene, Render::Preconfigure, Render::Do*Pass, ForkerGL::*, Model::Render, Shader.
//
// Usage: ref_driver --assets DIR --scene FILE --out DIR [--shadow hard|pcf|pcss] [--wrap 0..3]
//                   [--filter 0|1] [--ids] [--tga] [--frames N] [--quiet]
//        ref_driver --buffer-test OUTDIR | --matrices <12 floats>
//   --assets   directory that contains obj/ (the reference opens model paths relative to the CWD)
//   --wrap/--filter are applied BEFORE the Scene is constructed, because textures capture the modes at
//   load time (reference model.cpp:425).
#include <spdlog/spdlog.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>

#include "gshader.h"
#include "output.h"
#include "pbrshader.h"
#include "phongshader.h"
#include "render.h"
#include "utility.h"

extern int g_fglRefShadowMode;

namespace
{
double Now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void Dump1f(const std::string& path, const Buffer1f& b)
{
    int w = b.GetWidth(), h = b.GetHeight();
    if (w == 0 || h == 0) return;
    std::vector<float> v((size_t)w * h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) v[(size_t)x + (size_t)y * w] = b.GetValue(x, y);
    FILE* f = fopen(path.c_str(), "wb");
    fwrite(v.data(), sizeof(float), v.size(), f);
    fclose(f);
}

void Dump3f(const std::string& path, const Buffer3f& b)
{
    int w = b.GetWidth(), h = b.GetHeight();
    if (w == 0 || h == 0) return;
    std::vector<float> v((size_t)w * h * 3);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
        {
            Vector3f c = b.GetValue(x, y);
            size_t   i = ((size_t)x + (size_t)y * w) * 3;
            v[i] = c.x, v[i + 1] = c.y, v[i + 2] = c.z;
        }
    FILE* f = fopen(path.c_str(), "wb");
    fwrite(v.data(), sizeof(float), v.size(), f);
    fclose(f);
}

void DumpRGB(const std::string& path, const TGAImage& img)
{
    int w = img.GetWidth(), h = img.GetHeight();
    if (w == 0 || h == 0) return;
    std::vector<unsigned char> v((size_t)w * h * 3);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
        {
            TGAColor c = img.Get(x, y);
            size_t   i = ((size_t)x + (size_t)y * w) * 3;
            v[i] = c.r, v[i + 1] = c.g, v[i + 2] = c.b;
        }
    FILE* f = fopen(path.c_str(), "wb");
    fwrite(v.data(), 1, v.size(), f);
    fclose(f);
}

// Winner-ID program: same clip-space position as GShader / DepthShader, fragment colour = primitive
// ordinal + 1.  Shader is an open interface in the reference (shader.h:18-31), so this uses only its
// public surface.  ProcessVertex(face, 0) is the first call DrawTriangle's caller makes per face
// (mesh.cpp:16-23), which is where the ordinal advances.
struct IdShader : public Shader
{
    Matrix4x4f uModelMatrix, uViewMatrix, uProjectionMatrix, uLightSpaceMatrix;
    bool       lightSpace = false;
    int        counter = 0;
    int        current = 0;

    Point4f ProcessVertex(int faceIdx, int vertIdx) override
    {
        if (vertIdx == 0) current = ++counter;
        if (lightSpace)
        {
            Point4f cs = uLightSpaceMatrix * uModelMatrix * Point4f(mesh->Vert(faceIdx, vertIdx), 1.f);
            return cs / cs.w;
        }
        Point4f ws = uModelMatrix * Point4f(mesh->Vert(faceIdx, vertIdx), 1.f);
        Point4f vs = uViewMatrix * ws;
        Point4f cs = uProjectionMatrix * vs;
        return cs / cs.w;
    }
    bool ProcessFragment(const Vector3f&, Color3& c) override
    {
        c = Color3((Float)current, 0.f, 0.f);
        return false;
    }
};

void DumpIds(const std::string& path, const Scene& scene, bool lightSpace, int w, int h)
{
    ForkerGL::InitFrameBuffer(w, h);
    ForkerGL::InitDepthBuffer(w, h);
    ForkerGL::SetPassType(ForkerGL::ForwardPass);
    Float      ratio = scene.GetRatio();
    Matrix4x4f view = scene.GetCamera().GetViewMatrix();
    Matrix4x4f proj = (scene.GetProjectionType() == Camera::Orthographic)
                          ? scene.GetCamera().GetOrthographicMatrix(-1.f * ratio, 1.f * ratio, -1.f, 1.f,
                                                                    0.01f, 20.f)
                          : scene.GetCamera().GetPerspectiveMatrix(45.f, ratio, 0.01f, 20.f);
    IdShader s;
    s.lightSpace = lightSpace;
    s.uViewMatrix = view;
    s.uProjectionMatrix = proj;
    s.uLightSpaceMatrix = ForkerGL::GetLightSpaceMatrix();
    for (int i = 0; i < (int)scene.GetModelCount(); ++i)
    {
        s.uModelMatrix = scene.GetModelMatrix(i);
        scene.GetModel(i).Render(s);
    }
    std::vector<int> ids((size_t)w * h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            ids[(size_t)x + (size_t)y * w] = (int)ForkerGL::FrameBuffer.GetValue(x, y).x - 1;
    FILE* f = fopen(path.c_str(), "wb");
    fwrite(ids.data(), sizeof(int), ids.size(), f);
    fclose(f);
}
}  // namespace

// --buffer-test OUTDIR: the reference's own Buffer1f / Buffer3f post-processing (buffer.cpp:35-98, 140-203) on
// deterministic inputs (an LCG, restated in tests/parity.py), raw results dumped for tests/golden/make_golden.py.
static float Lcg01(uint32_t& s)
{
    s = s * 1664525u + 1013904223u;
    return (float)((s >> 8) & 0xffffffu) / 16777216.f;
}
static int BufferTest(const std::string& out)
{
    const int shapes[][2] = { { 5, 4 }, { 33, 7 }, { 1, 9 }, { 9, 1 }, { 64, 48 } };
    for (auto& sh : shapes)
    {
        const int W = sh[0], H = sh[1];
        for (int kind = 0; kind < 2; ++kind)
        {
            uint32_t seed = 12345u + (uint32_t)(W * 131 + H);
            Buffer1f b1(W, H, Buffer::Zero);
            Buffer3f b3(W, H, Buffer::Zero);
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) b1.SetValue(x, y, Lcg01(seed));
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x)
                {
                    float r = Lcg01(seed), g = Lcg01(seed), bb = Lcg01(seed);
                    b3.SetValue(x, y, Vector3f(r, g, bb));
                }
            if (kind == 0) b1.SimpleBlurDenoised(), b3.SimpleBlurDenoised();
            else b1.TwoPassGaussianBlurDenoised(), b3.TwoPassGaussianBlurDenoised();
            char name[256];
            snprintf(name, sizeof name, "%s/buffer_%s_%dx%d.raw", out.c_str(), kind == 0 ? "simple" : "gauss", W, H);
            FILE* f = fopen(name, "wb");
            if (!f) return 3;
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x)
                {
                    float v = b1.GetValue(x, y);
                    fwrite(&v, 4, 1, f);
                }
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x)
                {
                    Vector3f v = b3.GetValue(x, y);
                    float    q[3] = { v.x, v.y, v.z };
                    fwrite(q, 4, 3, f);
                }
            fclose(f);
        }
    }
    return 0;
}

// --matrices tx ty tz rotY scale ex ey ez cx cy cz ratio: the reference's host uniform builders (geometry.cpp:60-68,92-179,
// geometry.h:784-793) on one parameter set; prints the 105 floats as hex words in the order of the facade's
// frh_test_matrices: model(16) normal(9) lookat(16) persp(16) ortho(16) ortho*lookat(16) persp*lookat(16).
static int MatrixTest(int argc, const char* argv[])
{
    if (argc != 14) return 2;
    float a[12];
    for (int i = 0; i < 12; ++i) a[i] = strtof(argv[2 + i], nullptr);
    Matrix4x4f M = MakeModelMatrix(Vector3f(a[0], a[1], a[2]), a[3], a[4]);
    Matrix3x3f N = MakeNormalMatrix(M);
    Matrix4x4f L = MakeLookAtMatrix(Vector3f(a[5], a[6], a[7]), Vector3f(a[8], a[9], a[10]));
    Matrix4x4f P = MakePerspectiveMatrix(45.f, a[11], 0.01f, 20.f);
    Matrix4x4f O = MakeOrthographicMatrix(-3 * a[11], 3 * a[11], -3, 3, 0.1f, 20.f);
    Matrix4x4f OL = O * L, PL = P * L;
    auto put = [](float f) {
        uint32_t u;
        memcpy(&u, &f, 4);
        printf("%08x ", u);
    };
    auto put4 = [&](const Matrix4x4f& m) { for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) put(m[r][c]); };
    put4(M);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) put(N[r][c]);
    put4(L), put4(P), put4(O), put4(OL), put4(PL);
    printf("\n");
    return 0;
}

// --fragments: known-answer vectors of the reference's own vertex + fragment programs.  After the shadow pass, every mesh of
// every model goes through GShader and through the forward program of its model (BlinnPhongShader / PBRShader, uniforms as
// in render.cpp:121-154, 176-191); on faces 0, n/3 and 2n/3 ProcessFragment is evaluated at three barycentric points and
// the outputs are printed as hex words: 19 of GShader (normal, position, light-space NDC, albedo, emissive, param, type)
// and the 3 of gl_Color.  A probe Shader drives the programs from inside Model::Render (the meshes are not public) and
// returns a degenerate triangle, so nothing is rasterised.
struct FragmentProbe : public Shader
{
    Shader*  inner = nullptr;
    GShader* g = nullptr;
    Point4f  ProcessVertex(int faceIdx, int vertIdx) override
    {
        inner->Use(mesh);
        inner->ProcessVertex(faceIdx, vertIdx);
        const int n = mesh->NumFaces();
        if (vertIdx == 2 && (faceIdx == 0 || faceIdx == n / 3 || faceIdx == 2 * n / 3))
        {
            const Vector3f pts[3] = { Vector3f(1.f / 3, 1.f / 3, 1.f / 3), Vector3f(0.6f, 0.3f, 0.1f), Vector3f(0.05f, 0.15f, 0.8f) };
            for (const Vector3f& b : pts)
            {
                Color3 c(0.f);
                inner->ProcessFragment(b, c);
                auto put = [](float f) {
                    uint32_t u;
                    memcpy(&u, &f, 4);
                    printf("%08x ", u);
                };
                auto put3 = [&](const Vector3f& v) { put(v.x), put(v.y), put(v.z); };
                if (g) put3(g->outNormalWS), put3(g->outPositionWS), put3(g->outLightSpaceNDC), put3(g->outAlbedo), put3(g->outEmissive), put3(g->outParam), put(g->outShadingType);
                else put3(c);
            }
        }
        return Point4f(0.f, 0.f, 0.f, 1.f);
    }
    bool ProcessFragment(const Vector3f&, Color3&) override { return false; }
};

static void FragmentTest(const Scene& scene)
{
    Render::Preconfigure(scene);
    Render::DoShadowPass(scene);
    ForkerGL::InitFrameBuffer(scene.GetWidth(), scene.GetHeight());
    ForkerGL::InitDepthBuffer(scene.GetWidth(), scene.GetHeight());
    ForkerGL::SetPassType(ForkerGL::ForwardPass);
    Float      ratio = scene.GetRatio();
    Matrix4x4f view = scene.GetCamera().GetViewMatrix();
    Matrix4x4f proj = (scene.GetProjectionType() == Camera::Orthographic)
                          ? scene.GetCamera().GetOrthographicMatrix(-1.f * ratio, 1.f * ratio, -1.f, 1.f, 0.01f, 20.f)
                          : scene.GetCamera().GetPerspectiveMatrix(45.f, ratio, 0.01f, 20.f);
    for (int i = 0; i < (int)scene.GetModelCount(); ++i)
    {
        const Model&  model = scene.GetModel(i);
        FragmentProbe probe;
        GShader       gs;
        gs.uModelMatrix = scene.GetModelMatrix(i), gs.uViewMatrix = view, gs.uProjectionMatrix = proj;
        gs.uNormalMatrix = MakeNormalMatrix(gs.uModelMatrix), gs.uLightSpaceMatrix = ForkerGL::GetLightSpaceMatrix();
        probe.inner = &gs, probe.g = &gs;
        model.Render(probe);
        probe.g = nullptr;
        BlinnPhongShader bp;
        PBRShader        pb;
        bp.uModelMatrix = pb.uModelMatrix = scene.GetModelMatrix(i), bp.uViewMatrix = pb.uViewMatrix = view;
        bp.uProjectionMatrix = pb.uProjectionMatrix = proj, bp.uNormalMatrix = pb.uNormalMatrix = MakeNormalMatrix(bp.uModelMatrix);
        bp.uPointLight = pb.uPointLight = scene.GetPointLight(), bp.uEyePos = pb.uEyePos = scene.GetCamera().GetPosition();
        bp.uLightSpaceMatrix = pb.uLightSpaceMatrix = ForkerGL::GetLightSpaceMatrix();
        probe.inner = model.SupportPBR() ? (Shader*)&pb : (Shader*)&bp;
        model.Render(probe);
    }
    printf("\n");
}

int main(int argc, const char* argv[])
{
    if (argc == 3 && std::string(argv[1]) == "--buffer-test") return BufferTest(argv[2]);
    if (argc >= 2 && std::string(argv[1]) == "--matrices") return MatrixTest(argc, argv);
    std::string assets = ".", sceneFile, out = ".", shadow = "pcss";
    int         wrap = 0, filter = 0, frames = 1;
    bool        ids = false, tga = false, quiet = false, fragments = false;
    for (int i = 1; i < argc; ++i)
    {
        std::string a = argv[i];
        auto        next = [&]() { return std::string(i + 1 < argc ? argv[++i] : ""); };
        if (a == "--assets") assets = next();
        else if (a == "--scene") sceneFile = next();
        else if (a == "--out") out = next();
        else if (a == "--shadow") shadow = next();
        else if (a == "--wrap") wrap = atoi(next().c_str());
        else if (a == "--filter") filter = atoi(next().c_str());
        else if (a == "--frames") frames = atoi(next().c_str());
        else if (a == "--ids") ids = true;
        else if (a == "--tga") tga = true;
        else if (a == "--fragments") fragments = true;
        else if (a == "--quiet") quiet = true;
        else
        {
            fprintf(stderr, "unknown argument %s\n", a.c_str());
            return 2;
        }
    }
    if (sceneFile.empty())
    {
        fprintf(stderr, "usage: ref_driver --assets DIR --scene FILE --out DIR [...]\n");
        return 2;
    }
    g_fglRefShadowMode = (shadow == "hard") ? 0 : (shadow == "pcf") ? 1 : 2;

    spdlog::set_pattern("[%^%l%$] %v");
    spdlog::set_level(quiet ? spdlog::level::warn : spdlog::level::debug);

    // absolute paths before chdir
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd)) return 3;
    auto abs = [&](const std::string& p) { return (p.size() && p[0] == '/') ? p : std::string(cwd) + "/" + p; };
    sceneFile = abs(sceneFile);
    out = abs(out);
    if (chdir(assets.c_str()) != 0)
    {
        fprintf(stderr, "cannot chdir to %s\n", assets.c_str());
        return 3;
    }

    ForkerGL::TextureWrapMode((Texture::WrapMode)wrap);
    ForkerGL::TextureFilterMode((Texture::FilterMode)filter);

    double t0 = Now();
    Scene  scene(sceneFile);
    double tLoad = Now() - t0;
    if (fragments)
    {
        FragmentTest(scene);
        return 0;
    }

    // Render::Preconfigure (render.cpp:33-38) also resets the texture-mode statics to NoWrap/Nearest; that
    // has no effect on already-loaded textures (fact 9) and is kept as is.
    Render::Preconfigure(scene);

    double tShadow = 0, tRaster = 0, tLight = 0, tAA = 0, tFrame = 0;
    for (int f = 0; f < frames; ++f)
    {
        // Same sequence as Render::Render (render.cpp:40-58), timed per pass.
        double a = Now();
        Render::DoShadowPass(scene);
        double b = Now();
        double c, d;
        if (ForkerGL::GetRenderMode() == ForkerGL::Forward)
        {
            Render::DoForwardPass(scene);
            c = d = Now();
        }
        else
        {
            Render::DoGeometryPass(scene);
            c = Now();
            Render::DoLightingPass(scene);
            d = Now();
        }
        Render::DoSSAA(scene);
        double e = Now();
        if (f == 0)
        {
            tShadow = b - a, tRaster = c - b, tLight = d - c, tAA = e - d, tFrame = e - a;
            int W = ForkerGL::FrameBuffer.GetWidth(), H = ForkerGL::FrameBuffer.GetHeight();
            Dump1f(out + "/depth.f32", ForkerGL::DepthBuffer);
            Dump1f(out + "/shadow.f32", ForkerGL::ShadowBuffer);
            Dump3f(out + "/frame.f32", ForkerGL::FrameBuffer);
            if (ForkerGL::GetRenderMode() == ForkerGL::Deferred)
            {
                Dump3f(out + "/normal.f32", ForkerGL::NormalGBuffer);
                Dump3f(out + "/worldpos.f32", ForkerGL::WorldPosGBuffer);
                if (Shadow::GetShadowStatus()) Dump3f(out + "/lightndc.f32", ForkerGL::LightSpaceNDCPosGBuffer);
                Dump3f(out + "/albedo.f32", ForkerGL::AlbedoGBuffer);
                Dump3f(out + "/emissive.f32", ForkerGL::EmissiveGBuffer);
                Dump3f(out + "/param.f32", ForkerGL::ParamGBuffer);
                Dump1f(out + "/shadingtype.f32", ForkerGL::ShadingTypeGBuffer);
                Dump1f(out + "/ao.f32", ForkerGL::AmbientOcclusionGBuffer);
            }
            DumpRGB(out + "/frame.u8", ForkerGL::FrameBuffer.GenerateImage());
            if (scene.IsSSAAOn()) DumpRGB(out + "/ssaa.u8", ForkerGL::AntiAliasedImage);
            if (tga)
            {
                Output::OutputFrameBuffer();
                Output::OutputSSAAImage();
                Output::OutputShadowBuffer();
                Output::OutputZBuffer();
                Output::OutputNormalGBuffer();
                Output::OutputWorldPosGBuffer();
                Output::OutputAlbedoGBuffer();
                Output::OutputParamGBuffer();
                Output::OutputShadingTypeGBuffer();
                Output::OutputAmbientOcclusionGBuffer();
            }
            FILE* m = fopen((out + "/meta.json").c_str(), "w");
            fprintf(m,
                    "{\"width\": %d, \"height\": %d, \"out_width\": %d, \"out_height\": %d, \"ssaa\": %d, "
                    "\"ssaa_k\": %d, \"ssao\": %d, \"shadow\": %d, \"shadow_mode\": \"%s\", \"deferred\": %d, "
                    "\"wrap\": %d, \"filter\": %d, \"t_load\": %.6f, \"t_shadow\": %.6f, \"t_raster\": %.6f, "
                    "\"t_lighting\": %.6f, \"t_aa\": %.6f, \"t_frame\": %.6f}\n",
                    W, H, scene.GetWidth(), scene.GetHeight(), (int)scene.IsSSAAOn(), scene.GetSSAAKernelSize(),
                    (int)scene.IsSSAOOn(), (int)Shadow::GetShadowStatus(), shadow.c_str(),
                    (int)(ForkerGL::GetRenderMode() == ForkerGL::Deferred), wrap, filter, tLoad, tShadow,
                    tRaster, tLight, tAA, tFrame);
            fclose(m);
            if (ids)
            {
                DumpIds(out + "/ids_camera.i32", scene, false, W, H);
                if (Shadow::GetShadowStatus()) DumpIds(out + "/ids_light.i32", scene, true, W, H);
            }
        }
        printf("{\"frame\": %d, \"t_shadow\": %.6f, \"t_raster\": %.6f, \"t_lighting\": %.6f, \"t_aa\": %.6f, "
               "\"t_frame\": %.6f}\n",
               f, b - a, c - b, d - c, e - d, e - a);
        fflush(stdout);
    }
    return 0;
}
