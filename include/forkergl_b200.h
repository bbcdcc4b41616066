/* forkergl_b200.h — the C-ABI boundary of the B200-native ForkerRenderer hot path.
 *
 * The reference (forkercat/ForkerRenderer) has no FFI layer: its hot path is reached through the C++ surface of
 * `struct ForkerGL` (reference src/forkergl.h:15-81), `Shader` (src/shaders/shader.h:18-31), `Buffer1f/3f`
 * (src/buffer.h:32-93) and `Texture` (src/texture.h:11-145).  This header is what a host-side facade with those
 * same C++ signatures binds to (ours lives in forkerrenderer_b200/host/, the binding a reference maintainer
 * would add is shown in INTEGRATION.md).  Everything here is plain C: pointers, sizes, ints.  No torch types,
 * no C++ types, no CUDA types (a stream is passed as void*).
 *
 * Conventions
 *   - every call returns FGL_OK (0) or an FGL_ERR_* code; fgl_last_error(ctx) gives the text.  The library
 *     never aborts and never falls back to the CPU: with no usable CUDA device fgl_create fails.
 *   - matrices are 16 (or 9) floats, ROW-major, i.e. the memory image of the reference's Matrix4x4f/Matrix3x3f
 *     (rows of Vectors, src/geometry.h:744-746).
 *   - one ctx per GPU; calls on one ctx must be serialised by the caller; different ctxs are independent.
 *   - all work is enqueued on the ctx's stream; only fgl_sync and the read calls block.
 *   - pixel (x, y) has linear index x + y*W with y pointing up (src/buffer.h:37), exactly as in the reference.
 */
#ifndef FORKERGL_B200_H
#define FORKERGL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fgl_ctx fgl_ctx;

enum
{
    FGL_OK = 0,
    FGL_ERR_INVALID = 1,     /* bad argument / handle */
    FGL_ERR_CUDA = 2,        /* a CUDA runtime call failed (text has the CUDA error) */
    FGL_ERR_STATE = 3,       /* call sequence the reference would have asserted / crashed on */
    FGL_ERR_NOMEM = 4,
    FGL_ERR_UNSUPPORTED = 5
};

/* ForkerGL::RenderMode, ForkerGL::PassType — src/forkergl.h:19-31 (same numeric values) */
enum { FGL_MODE_FORWARD = 0, FGL_MODE_DEFERRED = 1 };
enum { FGL_PASS_FORWARD = 0, FGL_PASS_GEOMETRY = 1, FGL_PASS_LIGHTING = 2, FGL_PASS_SHADOW = 3 };

/* Texture::WrapMode, Texture::FilterMode — src/texture.h:14-26 (same numeric values) */
enum { FGL_WRAP_NOWRAP = 0, FGL_WRAP_REPEAT = 1, FGL_WRAP_MIRRORED_REPEAT = 2, FGL_WRAP_CLAMP_TO_EDGE = 3 };
enum { FGL_FILTER_NEAREST = 0, FGL_FILTER_LINEAR = 1 };

/* The four programs of src/shaders/{depthshader,gshader,phongshader,pbrshader}.h.  Virtual Shader callbacks
 * cannot run on the device; the facade maps each Shader subclass to its kind. */
enum { FGL_SHADER_DEPTH = 0, FGL_SHADER_G = 1, FGL_SHADER_BLINN_PHONG = 2, FGL_SHADER_PBR = 3 };

/* Shadow filter: compile-time macros in the reference (src/shaders/shadow.h:15-16), run-time here. */
enum { FGL_SHADOW_HARD = 0, FGL_SHADOW_PCF = 1, FGL_SHADOW_PCSS = 2 };

/* Buffer::InitType — src/buffer.h:14-20 */
enum { FGL_INIT_ZERO = 0, FGL_INIT_ONE = 1, FGL_INIT_MAX_POSITIVE = 2, FGL_INIT_MIN_NEGATIVE = 3 };

/* Planes = the public static buffers of ForkerGL (src/forkergl.h:41-54).  3-channel planes are stored on the
 * device as three SoA fp32 planes; fgl_read_plane returns them in the reference's AoS order (x,y,z per pixel). */
enum
{
    FGL_PLANE_FRAME = 0,        /* FrameBuffer              3 x f32 */
    FGL_PLANE_DEPTH = 1,        /* DepthBuffer              1 x f32 */
    FGL_PLANE_SHADOW = 2,       /* ShadowBuffer             1 x f32 */
    FGL_PLANE_NORMAL = 3,       /* NormalGBuffer            3 x f32 */
    FGL_PLANE_WORLDPOS = 4,     /* WorldPosGBuffer          3 x f32 */
    FGL_PLANE_LIGHTNDC = 5,     /* LightSpaceNDCPosGBuffer  3 x f32 */
    FGL_PLANE_ALBEDO = 6,       /* AlbedoGBuffer            3 x f32 */
    FGL_PLANE_EMISSIVE = 7,     /* EmissiveGBuffer          3 x f32 */
    FGL_PLANE_PARAM = 8,        /* ParamGBuffer             3 x f32 */
    FGL_PLANE_SHADINGTYPE = 9,  /* ShadingTypeGBuffer       1 x f32 */
    FGL_PLANE_AO = 10,          /* AmbientOcclusionGBuffer  1 x f32 */
    FGL_PLANE_FRAME_RGB8 = 11,  /* FrameBuffer.GenerateImage() (src/buffer.cpp:113-126), 3 x u8, R,G,B */
    FGL_PLANE_SSAA_RGB8 = 12,   /* AntiAliasedImage (src/render.cpp:291-343),           3 x u8, R,G,B */
    FGL_PLANE_PRIMID_CAMERA = 13, /* depth-test winner of the last camera raster pass, int32 submission index, -1 = none */
    FGL_PLANE_PRIMID_LIGHT = 14,  /* same for the last shadow pass */
    FGL_PLANE_COUNT = 15
};

/* Texture descriptor = what Texture's constructor captures (src/texture.h:28-36): the TGAImage bytes in memory
 * order (row 0 first, `bpp` bytes per texel in TGA channel order B,G,R[,A] or one grey byte; src/tgaimage.cpp:304-311)
 * plus the wrap/filter modes that were current at load time (src/model.cpp:425). */
int fgl_upload_texture(fgl_ctx* ctx, const uint8_t* texels, int width, int height, int bytes_per_texel,
                       int wrap_mode, int filter_mode, int* out_texture_id);

/* Material + PBRMaterial of one mesh (src/materials/material.h:13-41, pbrmaterial.h:13-46).
 * Texture ids are handles from fgl_upload_texture, or -1 for "no map". */
typedef struct FglMaterial
{
    float ka[3], kd[3], ks[3], ke[3];     /* Material */
    float pbr_ke[3], albedo[3];           /* PBRMaterial */
    float roughness, metalness;
    int   diffuse_map, specular_map, normal_map, emissive_map;                             /* Material */
    int   base_color_map, roughness_map, metalness_map, ao_map, pbr_normal_map, pbr_emissive_map; /* PBRMaterial */
} FglMaterial;

/* Model-level vertex arrays (src/model.h:48-51).  Normals/tangents are the RAW stored vectors; the per-fetch
 * Normalize of Mesh::Normal/Tangent (src/mesh.cpp:46-56) is applied on the device.  tangents may be NULL. */
int fgl_upload_vertices(fgl_ctx* ctx, const float* positions_xyz, int n_positions, const float* texcoords_uv,
                        int n_texcoords, const float* normals_xyz, int n_normals, const float* tangents_xyz,
                        int n_tangents, int* out_vertices_id);

/* One Mesh (src/mesh.h:53-61): 3 indices per face into the arrays above (0-based), tangent index == position
 * index (src/model.cpp:440-470).  has_tangents / support_pbr are the owning Model's flags (src/model.h:40-41). */
int fgl_upload_mesh(fgl_ctx* ctx, int vertices_id, int n_faces, const int* position_idx, const int* texcoord_idx,
                    const int* normal_idx, const FglMaterial* material, int has_tangents, int support_pbr,
                    int* out_mesh_id);

/* The uniform fields of the four Shader subclasses (e.g. src/shaders/gshader.h:23-30, phongshader.h:23-31). */
typedef struct FglUniforms
{
    float model[16];
    float view[16];
    float projection[16];
    float normal[9];
    float light_space[16];
    float light_position[3];
    float light_color[3];
    float eye_position[3];
} FglUniforms;

/* Constants the reference hard-codes as macros / file statics (SURVEY.md §5 "Config / flags"); defaults from
 * fgl_default_params are the reference's values. */
typedef struct FglParams
{
    int    shadow_mode;                /* FGL_SHADOW_*; reference ships PCSS (shadow.h:16) */
    double pcf_filter_size;            /* PCF_FILTER_SIZE 0.007 (a double literal in the reference) */
    double pcss_blocker_filter_size;   /* PCSS_BLOCKER_SEARCH_FILTER_SIZE 0.005 */
    float  area_light_size;            /* AREA_LIGHT_SIZE 2.5f */
    float  shadow_bias_slope;          /* 0.009f  shadow.cpp:116 */
    float  shadow_bias_min;            /* 0.007f  shadow.cpp:116 */
    float  shadow_intensity;           /* 0.6f    phongshader.h:197, pbrshader.h:229 */
    float  ssao_radius;                /* 0.075f  render.cpp:222 */
    float  ssao_range_check_radius;    /* 0.01f   render.cpp:224 */
    float  ssao_bias;                  /* 0.0005f render.cpp:258 */
    int    ssao_range_check;           /* true    render.cpp:223 */
    int    materialize_frame_f32;      /* 1: also keep the fp32 FrameBuffer planes (reference behaviour);
                                          0: lighting writes only the 8-bit image (bench fast path)   */
} FglParams;
/* PCF_NUM_SAMPLES (64), PCSS_BLOCKER_SEARCH_NUM_SAMPLES (32) and the SSAO sample count (32) are fixed: the
 * replayed sample stream is organised in chunks of 32 (DESIGN.md "sample stream"). */

void fgl_default_params(FglParams* out);

/* ---- context ---------------------------------------------------------------------------------------- */
int         fgl_create(int cuda_device, fgl_ctx** out_ctx);
void        fgl_destroy(fgl_ctx* ctx);
const char* fgl_last_error(fgl_ctx* ctx);      /* ctx may be NULL: error of the last failed fgl_create */
const char* fgl_backend_name(void);            /* "cuda-sm_100a" for the product library */
int         fgl_set_stream(fgl_ctx* ctx, void* cuda_stream); /* cudaStream_t; NULL = ctx-owned stream */
int         fgl_set_params(fgl_ctx* ctx, const FglParams* params);
int         fgl_sync(fgl_ctx* ctx);

/* ---- ForkerGL state (src/forkergl.h:37-67) ------------------------------------------------------------ */
int fgl_init_frame_buffer(fgl_ctx* ctx, int width, int height);      /* ForkerGL::InitFrameBuffer   forkergl.cpp:55 */
int fgl_init_depth_buffer(fgl_ctx* ctx, int width, int height);      /* ForkerGL::InitDepthBuffer   forkergl.cpp:60 */
int fgl_init_shadow_buffer(fgl_ctx* ctx, int width, int height);     /* ForkerGL::InitShadowBuffer  forkergl.cpp:65 */
int fgl_init_geometry_buffers(fgl_ctx* ctx, int width, int height);  /* ForkerGL::InitGeometryBuffers forkergl.cpp:70 */
int fgl_clear_color(fgl_ctx* ctx, const float rgb[3]);               /* ForkerGL::ClearColor        forkergl.cpp:84 */
int fgl_set_viewport(fgl_ctx* ctx, int x, int y, int w, int h);      /* ForkerGL::SetViewportMatrix forkergl.cpp:89 */
int fgl_get_viewport_matrix(fgl_ctx* ctx, float out16[16]);
int fgl_set_view_projection_matrix(fgl_ctx* ctx, const float m16[16]); /* forkergl.cpp:109 */
int fgl_get_view_projection_matrix(fgl_ctx* ctx, float out16[16]);
int fgl_set_light_space_matrix(fgl_ctx* ctx, const float m16[16]);     /* forkergl.cpp:119 */
int fgl_get_light_space_matrix(fgl_ctx* ctx, float out16[16]);
int fgl_set_render_mode(fgl_ctx* ctx, int mode);                       /* forkergl.cpp:129 */
int fgl_get_render_mode(fgl_ctx* ctx, int* out_mode);
int fgl_set_pass_type(fgl_ctx* ctx, int pass_type);                    /* forkergl.cpp:139 */
int fgl_set_shadow_status(fgl_ctx* ctx, int on);                       /* Shadow::SetShadowStatus shadow.cpp:13 */

/* Start of Render::Render (src/render.cpp:40): rewinds the replayed mt19937 sample stream to position 0 — the
 * reference process renders exactly one frame, so every frame sees the stream from its seed. */
int fgl_begin_frame(fgl_ctx* ctx);

/* Sort-first multi-GPU: restrict the camera-space shading passes (G-buffer resolve, SSAO, blur, lighting, SSAA) to
 * buffer rows [row_begin, row_end) (plus the few halo rows SSAO and the blur recurrence need).  The shadow pass and
 * the camera depth plane always cover the whole buffer (SSAO gathers depth anywhere).  (0, -1) = everything. */
int fgl_set_row_band(fgl_ctx* ctx, int row_begin, int row_end);
/* Device-side hand-off of the PCSS chain state inside a sort-first group on one NVLink / NVSwitch box (one context per
 * GPU, usually one process per GPU).  Every context owns a mailbox in device memory; the context of band r + 1 waits
 * ON THE DEVICE (a one-thread kernel queued in front of its chain kernel) for the running blocker count that the
 * context of band r stores into it by a peer write when its own chain kernel has finished — no host round trip, no
 * fgl_set_chain_blockers_before / fgl_get_chain_blockers per band.
 *   fgl_chain_peer_mailbox: this context's mailbox as a device pointer (same process) and / or as a 64-byte CUDA IPC handle.
 *   fgl_chain_peer_connect: the NEXT band's mailbox (pointer, or IPC handle from another process; both NULL for the last
 *   band), whether this band waits for a previous one (0 for the first band), enable = 0 switches the mechanism off.
 * Frames are counted from the connect call; all contexts of the group must render the same sequence of frames. */
int fgl_chain_peer_mailbox(fgl_ctx* ctx, void** out_device_ptr, void* out_ipc_handle, size_t ipc_handle_bytes);
int fgl_chain_peer_connect(fgl_ctx* ctx, void* next_device_ptr, const void* next_ipc_handle, int wait_for_previous, int enable);
/* Sort-first PCSS: the sample-stream position of a band depends on how many pixels of the bands before it found a
 * blocker (shadow.cpp:96-105).  Set that count before the lighting pass of a band; read the running total (count
 * before + this band's) after it and hand it to the next band.  Both default to / start from 0 for a whole frame. */
int fgl_set_chain_blockers_before(fgl_ctx* ctx, uint64_t blockers);
int fgl_get_chain_blockers(fgl_ctx* ctx, uint64_t* out_blockers);

/* ---- sort-first group: one frame on several GPUs of one NVLink / NVSwitch box ------------------------------------------
 * One context per GPU (usually one process per GPU).  Context r renders row band r of the screen and of the shadow map and
 * its kernels store what the other bands need straight into the other contexts' planes over NVLink: the shadow map rows
 * (ShadowBuffer of every context), the camera depth rows (DepthBuffer of every context; SSAO gathers depth anywhere), the
 * finished 8-bit rows (rank 0's frame) and the PCSS chain's running blocker count (the next band's mailbox).  Ordering is
 * kept on the device by epoch flags; the host only exchanges the FglGroupMember records once (csrc/group.cu).
 *   fgl_group_export   sizes this context's exchanged planes for a width x height frame and describes them: CUDA IPC
 *                      handles (other processes) and raw device pointers (contexts of the same process)
 *   fgl_group_connect  members[0..world) = the records of all contexts in rank order (this context's own at [rank]).
 *                      EVERY context must have returned from fgl_group_connect before any of them begins a frame.
 *                      From then on Render::Render (or the raw pass calls) on this context renders its band; all contexts
 *                      must render the same sequence of frames.  Deferred mode only; SSAA is not available in a group; every
 *                      raster pass of a group frame is submitted whole (one flush) and has at least one triangle (a pass
 *                      without any would never signal its rows to the other bands: they report a time-out after 30 s).
 *   fgl_group_read_frame  rank 0: waits (on the device) for every band of the current frame, then copies the 8-bit frame
 *                      (FGL_PLANE_FRAME_RGB8 layout) to host memory; other ranks: FGL_ERR_STATE
 *   fgl_group_disconnect  back to a stand-alone context */
typedef struct FglGroupMember
{
    unsigned char shadow_ipc[64], depth_ipc[64], frame_ipc[64], flags_ipc[64], chain_ipc[64];
    void *        shadow_ptr, *depth_ptr, *frame_ptr, *flags_ptr, *chain_ptr;
    int           width, height, device, reserved;
} FglGroupMember;
int fgl_group_export(fgl_ctx* ctx, int width, int height, FglGroupMember* out_member);
int fgl_group_connect(fgl_ctx* ctx, int rank, int world, const FglGroupMember* members, int same_process);
int fgl_group_read_frame(fgl_ctx* ctx, void* dst_host, size_t dst_bytes);
int fgl_group_disconnect(fgl_ctx* ctx);

/* ---- draw submission ------------------------------------------------------------------------------- */
/* Mesh::Draw (src/mesh.cpp:10-25) for every face of the mesh: vertex program x3 + ForkerGL::DrawTriangle
 * (src/forkergl.cpp:239-324).  Primitive ids follow submission order.  The triangles are rasterised when the
 * pass is flushed: at the next fgl_set_pass_type / fgl_init_* / fgl_draw_screen_space_pixels / read / sync. */
int fgl_draw_mesh(fgl_ctx* ctx, int mesh_id, int shader_kind, const FglUniforms* uniforms);

/* ForkerGL::DrawTriangle (src/forkergl.h:74, src/forkergl.cpp:239-324) for a batch of triangles whose vertex programs the CALLER
 * has already run (Shader::ProcessVertex, src/shaders/shader.h:28): what arrives is what the reference's DrawTriangle sees —
 * the three NDC positions ProcessVertex returned — plus the varyings the shader object holds when ProcessFragment is called.
 *   ndc        12 floats per triangle: x,y,z,w of vertex 0, 1, 2
 *   varyings   camera programs (G / Blinn-Phong / PBR), 48 floats per triangle, every attribute already times 1/w_clip
 *              (gshader.h:71-92): [0..8] world position (vertex-major: x,y,z of vertex 0, then 1, 2), [9..17] world normal,
 *              [18..26] world tangent, [27..35] light-space NDC, [36..38] u, [39..41] v, [42..44] 1/w_clip, [45..47] unused.
 *              NULL for DepthShader.
 *   light_z    DepthShader only: 3 floats per triangle, the z row of vPositionNDC (depthshader.h:21-36); NULL otherwise
 * mesh_id supplies the material and the owning model's flags; uniforms supplies light / eye for the forward programs.
 * Primitive ids continue the pass's submission order, triangle by triangle; the arrays are copied during the call. */
int fgl_draw_triangles(fgl_ctx* ctx, int mesh_id, int shader_kind, const FglUniforms* uniforms, int n_triangles, const float* ndc,
                       const float* varyings, const float* light_z);

/* ForkerGL::DrawScreenSpacePixels (src/forkergl.cpp:326-380), the deferred lighting loop. */
int fgl_draw_screen_space_pixels(fgl_ctx* ctx, const float eye_position[3], const float light_position[3],
                                 const float light_color[3]);
/* Optional head start for the lighting pass, callable once the geometry pass has been submitted (before fgl_ssao).
 * PCSS frames: everything that does not depend on the row bands above this context's band (shadow coordinates,
 * classification, cell masks, pilot), and — when the chain's input is known (a whole-frame context) or arrives on
 * the device (fgl_chain_peer_connect) — the sample-stream chain kernel itself, issued on a stream of its own so that
 * SSAO and the blur overlap it.  fgl_draw_screen_space_pixels picks the result up; without this call it does all of
 * it itself.  A sort-first driver using the HOST hand-off calls it before waiting for fgl_get_chain_blockers of the
 * band above.  ssao_follows: non-zero iff fgl_ssao will run between this call and fgl_draw_screen_space_pixels (SSAO
 * consumes the reference's sample stream before the lighting loop does, render.cpp:204-209). */
int fgl_prepare_screen_space_pixels(fgl_ctx* ctx, const float eye_position[3], const float light_position[3],
                                    const float light_color[3], int ssao_follows);

/* Render::DoSSAO without its trailing blur (src/render.cpp:214-286) */
int fgl_ssao(fgl_ctx* ctx);
/* Buffer1f/3f::TwoPassGaussianBlurDenoised (src/buffer.cpp:59-98,164-203) and SimpleBlurDenoised
 * (src/buffer.cpp:35-57,140-162), in-place semantics reproduced. */
enum { FGL_BLUR_SIMPLE_3X3 = 0, FGL_BLUR_TWO_PASS_GAUSSIAN = 1 };
int fgl_blur(fgl_ctx* ctx, int plane, int blur_kind);
/* Render::DoSSAA (src/render.cpp:291-343): quantise the frame buffer, k x k integer box. */
int fgl_ssaa_resolve(fgl_ctx* ctx, int kernel_size);

/* ---- buffers --------------------------------------------------------------------------------------- */
int fgl_plane_info(fgl_ctx* ctx, int plane, int* out_width, int* out_height, int* out_channels,
                   int* out_bytes_per_channel);
/* Blocking copy to / from HOST memory in the reference's layout (AoS for 3-channel planes). */
int fgl_read_plane(fgl_ctx* ctx, int plane, void* dst_host, size_t dst_bytes);
int fgl_write_plane(fgl_ctx* ctx, int plane, const void* src_host, size_t src_bytes);
/* Page-locked host memory for fgl_read_plane / fgl_write_plane: a read into it is one DMA at PCIe speed instead of a
 * staged copy into pageable memory (the reference's Buffer::GetValue loops have no counterpart; this is the boundary
 * a caller reads finished frames through).  Plain malloc/free semantics; the ctx only selects the device. */
int fgl_host_alloc(fgl_ctx* ctx, size_t bytes, void** out_ptr);
int fgl_host_free(fgl_ctx* ctx, void* ptr);
/* Same, but dst is DEVICE memory on the ctx's GPU (e.g. a torch tensor's data_ptr) and the copy is enqueued
 * on the ctx's stream.  Rows [row_begin,row_end) only; used by the multi-GPU gather. */
int fgl_copy_plane_rows_to_device(fgl_ctx* ctx, int plane, int row_begin, int row_end, void* dst_device,
                                  size_t dst_bytes);

/* ---- recorded frames: multi-frame use (SURVEY.md §8 f, N4; the reference's main is one-shot, src/main.cpp:18-55) ----------
 * A frame whose passes need no host read-back is a fixed sequence of copies, clears and kernels on the ctx's stream: hard-shadow
 * or PCF lighting, with or without SSAO, the blur and SSAA.  It can be recorded ONCE as a CUDA graph and replayed with one
 * launch per frame.  Not recordable (FGL_ERR_UNSUPPORTED, the context stays usable): PCSS (its sample-stream chain reports
 * counts to the host), forward mode with PCF / PCSS, fgl_draw_triangles, a context of a sort-first group.
 *   fgl_frame_record_begin  every call on this context up to fgl_frame_record_end is recorded instead of executed — the usual
 *                           frame sequence from fgl_begin_frame to fgl_draw_screen_space_pixels / fgl_ssaa_resolve.  The same
 *                           frame (same buffer sizes, same scene) must have been rendered once before without recording:
 *                           buffers and sample tables are sized then, nothing may allocate or wait for the device while
 *                           recording.  Reads, uploads and fgl_sync are refused until the recording ends.
 *   fgl_frame_record_end    *frame_id < 0: creates a replayable frame and stores its id; *frame_id >= 0: UPDATES that frame in
 *                           place to the newly recorded parameters (a moved camera or light: same passes, other uniforms) —
 *                           the graph is patched, not rebuilt, when its shape is unchanged.  The recorded frame has NOT been
 *                           rendered yet: replay it before reading planes.
 *   fgl_frame_record_abort  ends a recording without keeping anything (after an error inside the sequence)
 *   fgl_frame_replay        enqueues the whole frame on the ctx's stream; planes are then read as after an eager frame
 *   fgl_frame_info          number of graph nodes (kernels + copies + clears) and of kernel launches of one replay
 *   fgl_frame_release       frees a recorded frame
 * A recorded frame becomes invalid (fgl_frame_replay: FGL_ERR_STATE) when a device buffer it uses is re-allocated afterwards,
 * e.g. by rendering a larger frame on the same context.  Everything the recorded calls passed by value — uniforms, FglParams,
 * buffer sizes — is part of the recording: change it by recording again into the same id. */
int fgl_frame_record_begin(fgl_ctx* ctx);
int fgl_frame_record_end(fgl_ctx* ctx, int* frame_id);
int fgl_frame_record_abort(fgl_ctx* ctx);
int fgl_frame_replay(fgl_ctx* ctx, int frame_id);
int fgl_frame_info(fgl_ctx* ctx, int frame_id, int* out_nodes, int* out_kernel_launches);
int fgl_frame_release(fgl_ctx* ctx, int frame_id);

/* ---- instrumentation --------------------------------------------------------------------------------- */
/* Per-kernel CUDA-event timing.  When enabled every kernel launch is bracketed by events on the ctx stream;
 * fgl_get_timings returns, for the launches since the last fgl_reset_timings, up to `max` records. */
typedef struct FglTiming
{
    char     name[48];
    float    ms_total;
    int      launches;
    uint64_t algorithmic_bytes;   /* per SURVEY.md §8(d), summed over the launches */
} FglTiming;
int fgl_enable_timing(fgl_ctx* ctx, int on);
int fgl_reset_timings(fgl_ctx* ctx);
int fgl_get_timings(fgl_ctx* ctx, FglTiming* out, int max, int* out_count);
/* number of kernel launches issued by this library since ctx creation */
int fgl_launch_count(fgl_ctx* ctx, uint64_t* out);
/* bytes this library copied host -> device and device -> host since ctx creation (draw commands, texture tables, plane
 * reads / writes, the few words of chain state): what a caller's frame really moves over PCIe */
int fgl_transfer_bytes(fgl_ctx* ctx, uint64_t* out_host_to_device, uint64_t* out_device_to_host);

#ifdef __cplusplus
}
#endif
#endif /* FORKERGL_B200_H */
