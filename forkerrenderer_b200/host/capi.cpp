// capi.cpp — C entry points over the C++ facade, for ctypes (tests, bench.py) and other FFI callers.
// These are conveniences above the real boundary (include/forkergl_b200.h): load a .scene with the facade's
// loaders, run Render::Preconfigure + Render::Render, hand out the fgl context for plane reads.
#include <unistd.h>

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "forkergl.h"
#include "forkergl_b200.h"
#include "output.h"
#include "render.h"
#include "shader.h"
#include "shadow.h"

static std::string s_Error;

template <typename F>
static int Guard(F&& f)
{
    try
    {
        f();
        return 0;
    }
    catch (const std::exception& e)
    {
        s_Error = e.what();
        return 1;
    }
}

namespace
{
struct ReplayState
{
    enum { Cold, Warm, Recorded, Unsupported } state = Cold;
    int                        frameId = -1;
    std::vector<unsigned char> key;
    std::string                why;
};
ReplayState s_Replay;

std::vector<unsigned char> ReplayKey(const Scene& s, int shadow_mode, int materialize)
{
    std::vector<unsigned char> k;
    auto put = [&](const void* p, size_t n) { k.insert(k.end(), (const unsigned char*)p, (const unsigned char*)p + n); };
    const Scene* sp = &s;
    put(&sp, sizeof sp), put(&shadow_mode, sizeof shadow_mode), put(&materialize, sizeof materialize);
    const Vector3f eye = s.GetCamera().GetPosition();
    const Matrix4x4f view = s.GetCamera().GetViewMatrix();
    const PointLight& l = s.GetPointLight();
    put(&eye, sizeof eye), put(&view, sizeof view), put(&l.position, sizeof l.position), put(&l.color, sizeof l.color);
    // the run-time constants pushed by Render::Render (fgl_set_params) are baked into a recording as kernel parameters
    const FglParams& prm = ForkerGL::Params();
    const double     pd[2] = { prm.pcf_filter_size, prm.pcss_blocker_filter_size };
    const float      pf[7] = { prm.area_light_size, prm.shadow_bias_slope, prm.shadow_bias_min, prm.shadow_intensity, prm.ssao_radius,
                               prm.ssao_range_check_radius, prm.ssao_bias };
    put(pd, sizeof pd), put(pf, sizeof pf), put(&prm.ssao_range_check, sizeof prm.ssao_range_check);
    const int misc[8] = { s.GetWidth(), s.GetHeight(), s.IsSSAAOn(), s.GetSSAAKernelSize(), s.IsSSAOOn(), (int)ForkerGL::GetRenderMode(),
                          Shadow::GetShadowStatus(), (int)s.GetProjectionType() };
    put(misc, sizeof misc);
    for (unsigned i = 0; i < s.GetModelCount(); ++i)
    {
        const Matrix4x4f m = s.GetModelMatrix(i);
        put(&m, sizeof m);
    }
    return k;
}
}  // namespace

extern "C" {

const char* frh_last_error() { return s_Error.c_str(); }

fgl_ctx* frh_context()
{
    fgl_ctx* c = nullptr;
    Guard([&] { c = ForkerGL::Context(); });
    return c;
}

void frh_shutdown() { ForkerGL::Shutdown(); }

// Loads a .scene.  Model paths inside it are relative to assets_dir (the reference resolves them against the
// CWD; we chdir for the duration of the load).  wrap/filter are the ForkerGL texture modes in force while the
// textures load (reference model.cpp:425).
int frh_scene_load(const char* assets_dir, const char* scene_file, int wrap, int filter, void** out_scene)
{
    return Guard([&] {
        char cwd[4096];
        if (!getcwd(cwd, sizeof cwd)) throw std::runtime_error("getcwd failed");
        std::string scenePath = scene_file;
        if (scenePath.empty() || scenePath[0] != '/') scenePath = std::string(cwd) + "/" + scenePath;
        if (assets_dir && *assets_dir && chdir(assets_dir) != 0)
            throw std::runtime_error(std::string("cannot chdir to ") + assets_dir);
        ForkerGL::TextureWrapMode((Texture::WrapMode)wrap);
        ForkerGL::TextureFilterMode((Texture::FilterMode)filter);
        Scene* s = nullptr;
        try
        {
            s = new Scene(scenePath);
        }
        catch (...)
        {
            if (chdir(cwd) != 0) {}
            throw;
        }
        if (chdir(cwd) != 0) {}
        if (!s->IsValid())
        {
            delete s;
            throw std::runtime_error(std::string("failed to load scene ") + scene_file);
        }
        for (unsigned i = 0; i < s->GetModelCount(); ++i) s->GetModel(i).UploadToDevice();
        *out_scene = s;
    });
}

void frh_scene_free(void* scene)
{
    // a recording refers to this scene's device meshes and is keyed by its address: it goes with the scene
    if (s_Replay.frameId >= 0) Guard([&] { fgl_frame_release(ForkerGL::Context(), s_Replay.frameId); });
    s_Replay = ReplayState();
    delete (Scene*)scene;
}

// info[0..7] = width, height, ssaa on, ssaa k, ssao on, deferred, shadow on, triangle count
int frh_scene_info(void* scene, int* info)
{
    return Guard([&] {
        Scene* s = (Scene*)scene;
        int    tris = 0;
        for (unsigned i = 0; i < s->GetModelCount(); ++i) tris += s->GetModel(i).GetNumFaces();
        info[0] = s->GetWidth(), info[1] = s->GetHeight(), info[2] = s->IsSSAAOn(), info[3] = s->GetSSAAKernelSize();
        info[4] = s->IsSSAOOn(), info[5] = ForkerGL::GetRenderMode() == ForkerGL::Deferred;
        info[6] = Shadow::GetShadowStatus(), info[7] = tris;
    });
}

int frh_set_camera(void* scene, const float eye[3], const float look_at[3])
{
    return Guard([&] {
        Scene* s = (Scene*)scene;
        s->GetCamera().SetPosition(eye[0], eye[1], eye[2]);
        s->GetCamera().SetLookAtPos(look_at[0], look_at[1], look_at[2]);
    });
}

int frh_set_point_light(void* scene, const float position[3], const float color[3])
{
    return Guard([&] {
        PointLight& l = ((Scene*)scene)->GetPointLight();
        l.position = Vector3f(position[0], position[1], position[2]);
        l.color = Vector3f(color[0], color[1], color[2]);
    });
}

// Mesh::Draw face by face through Shader::ProcessVertex + ForkerGL::DrawTriangle (reference mesh.cpp:10-25) instead of one
// indexed draw per mesh
void frh_set_per_triangle_submission(int on) { ForkerGL::SetPerTriangleSubmission(on != 0); }

// One frame: Render::Preconfigure + Render::Render (reference main.cpp:37-40).
int frh_render(void* scene, int shadow_mode, int materialize_frame_f32)
{
    return Guard([&] {
        Scene* s = (Scene*)scene;
        Shadow::SetShadowMode((Shadow::Mode)shadow_mode);
        ForkerGL::Params().materialize_frame_f32 = materialize_frame_f32;
        Render::Preconfigure(*s);
        Render::Render(*s);
    });
}

// Multi-frame use (SURVEY.md §8 f, N4; the reference's main renders one frame and exits, main.cpp:18-55): the frame as a
// recorded CUDA graph (include/forkergl_b200.h "fgl_frame_*").  The first call renders eagerly (buffers and sample tables get
// their sizes), the second records Render::Preconfigure + Render::Render and replays the recording, every further call is ONE
// graph launch; after frh_set_camera / frh_set_point_light (or another shadow mode) the frame is recorded again and the graph
// patched in place.  Frames that cannot be recorded (PCSS, forward mode with a stochastic filter, per-triangle submission,
// a sort-first group) are rendered eagerly, every time.  *out_replayed: 1 if this call was a graph launch.

int frh_render_replay(void* scene, int shadow_mode, int materialize_frame_f32, int* out_replayed)
{
    return Guard([&] {
        Scene* s = (Scene*)scene;
        if (out_replayed) *out_replayed = 0;
        auto eager = [&] {
            Shadow::SetShadowMode((Shadow::Mode)shadow_mode);
            ForkerGL::Params().materialize_frame_f32 = materialize_frame_f32;
            Render::Preconfigure(*s);
            Render::Render(*s);
        };
        fgl_ctx* ctx = ForkerGL::Context();
        std::vector<unsigned char> key = ReplayKey(*s, shadow_mode, materialize_frame_f32);
        ReplayState& R = s_Replay;
        if (R.state == ReplayState::Unsupported && key == R.key) return eager();
        if (R.state == ReplayState::Cold || (R.state == ReplayState::Unsupported && key != R.key))
        {   // sizes first; (an unsupported frame gets another chance when its parameters change, e.g. PCSS -> hard)
            eager();
            if (R.state == ReplayState::Cold) R.state = ReplayState::Warm;
            else R.state = ReplayState::Warm, R.key.clear();
            return;
        }
        if (!(R.state == ReplayState::Recorded && key == R.key))
        {
            ForkerGL::FlushTriangles();
            bool ok = fgl_frame_record_begin(ctx) == FGL_OK;  // refused inside a sort-first group, with timing on, or by a back end without graphs
            if (!ok) R.why = fgl_last_error(ctx);
            if (ok)
            {
                try
                {
                    eager();
                }
                catch (const std::exception& e)
                {
                    ok = false, R.why = e.what();
                }
            }
            if (ok && fgl_frame_record_end(ctx, &R.frameId) != FGL_OK) ok = false, R.why = fgl_last_error(ctx);
            if (!ok)
            {
                fgl_frame_record_abort(ctx);
                R.state = ReplayState::Unsupported, R.key = key;
                return eager();
            }
            R.state = ReplayState::Recorded, R.key = key;
        }
        ForkerGL::Check(fgl_frame_replay(ctx, R.frameId), "replay");
        ForkerGL::InvalidateHostMirrors();
        if (out_replayed) *out_replayed = 1;
    });
}
// why the last frh_render_replay fell back to eager rendering ("" if it did not)
const char* frh_replay_fallback_reason() { return s_Replay.state == ReplayState::Unsupported ? s_Replay.why.c_str() : ""; }

// The same frame in two calls (sort-first multi-GPU: the PCSS chain state is exchanged between them).
int frh_render_begin(void* scene, int shadow_mode, int materialize_frame_f32)
{
    return Guard([&] {
        Scene* s = (Scene*)scene;
        Shadow::SetShadowMode((Shadow::Mode)shadow_mode);
        ForkerGL::Params().materialize_frame_f32 = materialize_frame_f32;
        Render::Preconfigure(*s);
        Render::RenderGeometryStage(*s);
    });
}
int frh_render_finish(void* scene)
{
    return Guard([&] { Render::RenderLightingStage(*(Scene*)scene); });
}

// Known-answer vectors of the HOST vertex + fragment programs (Shader::ProcessVertex / ProcessFragment, host/programs.cpp) in
// the enumeration of `ref_driver --fragments`: after the shadow pass, every mesh of every model through GShader and its
// model's forward program; faces 0, n/3, 2n/3 at three barycentric points; 19 words of GShader outputs / 3 of gl_Color each.
int frh_test_fragments(void* scenePtr, int shadow_mode, float* out, int max, int* count)
{
    return Guard([&] {
        Scene& scene = *(Scene*)scenePtr;
        Shadow::SetShadowMode((Shadow::Mode)shadow_mode);
        Shadow::ResetHostSampleStream();
        Render::Preconfigure(scene);
        fgl_ctx* ctx = ForkerGL::Context();
        FglParams params = ForkerGL::Params();
        params.shadow_mode = shadow_mode;
        ForkerGL::Check(fgl_set_params(ctx, &params), "params");
        ForkerGL::Check(fgl_set_shadow_status(ctx, Shadow::GetShadowStatus() ? 1 : 0), "shadow status");
        ForkerGL::Check(fgl_begin_frame(ctx), "begin frame");
        Render::DoShadowPass(scene);
        int  n = 0;
        auto put = [&](float f) { if (n < max) out[n] = f; ++n; };
        auto put3 = [&](const Vector3f& v) { put(v.x), put(v.y), put(v.z); };
        Float      ratio = scene.GetRatio();
        Matrix4x4f view = scene.GetCamera().GetViewMatrix();
        Matrix4x4f proj = scene.GetProjectionType() == Camera::Orthographic
                              ? scene.GetCamera().GetOrthographicMatrix(-1.f * ratio, 1.f * ratio, -1.f, 1.f, 0.01f, 20.f)
                              : scene.GetCamera().GetPerspectiveMatrix(45.f, ratio, 0.01f, 20.f);
        const Vector3f pts[3] = { Vector3f(1.f / 3, 1.f / 3, 1.f / 3), Vector3f(0.6f, 0.3f, 0.1f), Vector3f(0.05f, 0.15f, 0.8f) };
        for (int i = 0; i < (int)scene.GetModelCount(); ++i)
        {
            const Model& model = scene.GetModel(i);
            GShader      gs;
            gs.uModelMatrix = scene.GetModelMatrix(i), gs.uViewMatrix = view, gs.uProjectionMatrix = proj;
            gs.uNormalMatrix = MakeNormalMatrix(gs.uModelMatrix), gs.uLightSpaceMatrix = ForkerGL::GetLightSpaceMatrix();
            BlinnPhongShader bp;
            PBRShader        pb;
            bp.uModelMatrix = pb.uModelMatrix = scene.GetModelMatrix(i), bp.uViewMatrix = pb.uViewMatrix = view;
            bp.uProjectionMatrix = pb.uProjectionMatrix = proj, bp.uNormalMatrix = pb.uNormalMatrix = MakeNormalMatrix(bp.uModelMatrix);
            bp.uPointLight = pb.uPointLight = scene.GetPointLight(), bp.uEyePos = pb.uEyePos = scene.GetCamera().GetPosition();
            bp.uLightSpaceMatrix = pb.uLightSpaceMatrix = ForkerGL::GetLightSpaceMatrix();
            Shader* programs[2] = { &gs, model.SupportPBR() ? (Shader*)&pb : (Shader*)&bp };
            for (int pass = 0; pass < 2; ++pass)
                for (auto& kv : model.Meshes())
                {
                    Shader&   sh = *programs[pass];
                    const int nf = kv.second->NumFaces();
                    for (int f = 0; f < nf; ++f)
                    {
                        if (!(f == 0 || f == nf / 3 || f == 2 * nf / 3)) continue;
                        sh.Use(kv.second);
                        for (int v = 0; v < 3; ++v) sh.ProcessVertex(f, v);
                        for (const Vector3f& b : pts)
                        {
                            Color3 c(0.f);
                            sh.ProcessFragment(b, c);
                            if (pass == 0) put3(gs.outNormalWS), put3(gs.outPositionWS), put3(gs.outLightSpaceNDC), put3(gs.outAlbedo), put3(gs.outEmissive), put3(gs.outParam), put(gs.outShadingType);
                            else put3(c);
                        }
                    }
                }
        }
        *count = n;
    });
}

// ---- sort-first group: one frame on several GPUs (include/forkergl_b200.h "fgl_group_*", csrc/group.cu) ---------------------
// One process per GPU.  The host side is these four calls; the processes only have to hand each other their 368-byte member
// records once (any transport: bench.py uses torch.distributed's all_gather for that and for nothing else).
//   frh_group_export(scene, member)            this process's record for the scene's frame size
//   frh_group_connect(rank, world, members)    all records in rank order; then a barrier across the processes
//   frh_render(scene, ...)                      every frame: this rank's band, everything else happens on the devices
//   frh_group_read_frame(dst, bytes)            rank 0: the finished 8-bit frame
int frh_group_member_bytes() { return (int)sizeof(FglGroupMember); }
int frh_group_export(void* scene, void* member)
{
    return Guard([&] {
        Scene* s = (Scene*)scene;
        int    k = s->IsSSAAOn() ? s->GetSSAAKernelSize() : 1;
        ForkerGL::Check(fgl_group_export(ForkerGL::Context(), s->GetWidth() * k, s->GetHeight() * k, (FglGroupMember*)member), "group export");
    });
}
int frh_group_connect(int rank, int world, const void* members, int same_process)
{
    return Guard([&] { ForkerGL::Check(fgl_group_connect(ForkerGL::Context(), rank, world, (const FglGroupMember*)members, same_process), "group connect"); });
}
int frh_group_read_frame(void* dst, size_t bytes)
{
    return Guard([&] { ForkerGL::Check(fgl_group_read_frame(ForkerGL::Context(), dst, bytes), "group read frame"); });
}
int frh_group_disconnect()
{
    return Guard([&] { ForkerGL::Check(fgl_group_disconnect(ForkerGL::Context()), "group disconnect"); });
}

// Output::* of the reference's main (main.cpp:43-52) into `dir`.
int frh_output_tga(const char* dir)
{
    return Guard([&] {
        Output::SetDirectory(dir);
        Output::OutputFrameBuffer();
        Output::OutputSSAAImage();
        Output::OutputShadowBuffer();
        Output::OutputZBuffer();
        if (ForkerGL::GetRenderMode() == ForkerGL::Deferred)
        {
            Output::OutputNormalGBuffer();
            Output::OutputWorldPosGBuffer();
            Output::OutputAlbedoGBuffer();
            Output::OutputParamGBuffer();
            Output::OutputShadingTypeGBuffer();
            Output::OutputAmbientOcclusionGBuffer();
        }
    });
}

// Host matrix builders, for the bit-exactness test against the reference's (tests/test_host_math.py).
// out = model(16) normal(9) lookat(16) persp(16) ortho(16) ortho*lookat(16) persp*lookat(16) = 105 floats
void frh_test_matrices(const float t[3], float rot, float scale, const float eye[3], const float center[3],
                       float ratio, float* out)
{
    Matrix4x4f M = MakeModelMatrix(Vector3f(t[0], t[1], t[2]), rot, scale);
    Matrix3x3f N = MakeNormalMatrix(M);
    Matrix4x4f L = MakeLookAtMatrix(Vector3f(eye[0], eye[1], eye[2]), Vector3f(center[0], center[1], center[2]));
    Matrix4x4f P = MakePerspectiveMatrix(45.f, ratio, 0.01f, 20.f);
    Matrix4x4f O = MakeOrthographicMatrix(-3 * ratio, 3 * ratio, -3, 3, 0.1f, 20.f);
    Matrix4x4f OL = O * L, PL = P * L;
    int        k = 0;
    auto put4 = [&](const Matrix4x4f& m) { for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out[k++] = m[r][c]; };
    put4(M);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) out[k++] = N[r][c];
    put4(L), put4(P), put4(O), put4(OL), put4(PL);
}
}  // extern "C"
