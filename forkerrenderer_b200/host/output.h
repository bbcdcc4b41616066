// output.h — namespace Output of the drop-in facade: one TGA dump per ForkerGL buffer, under the reference's entry-point
// names (src/output.h:7-20) and file names (src/output.cpp:12-86).  The list below is the single source of both the
// declarations here and the definitions in output.cpp.
#pragma once
#include <string>

// X(entry point suffix, ForkerGL buffer, file name below the output directory)
#define FGL_OUTPUT_GBUFFER_DUMPS(X)                                              \
    X(NormalGBuffer, NormalGBuffer, "/gbuffer_normal.tga")                       \
    X(WorldPosGBuffer, WorldPosGBuffer, "/gbuffer_worldpos.tga")                 \
    X(AlbedoGBuffer, AlbedoGBuffer, "/gbuffer_albedo.tga")                       \
    X(ParamGBuffer, ParamGBuffer, "/gbuffer_param.tga")                          \
    X(ShadingTypeGBuffer, ShadingTypeGBuffer, "/gbuffer_shading_type.tga")       \
    X(AmbientOcclusionGBuffer, AmbientOcclusionGBuffer, "/gbuffer_ambient_occlusion.tga")

namespace Output
{
void SetDirectory(const std::string& dir);  // default "output" (reference output.cpp:14)
// frame, depth ("zbuffer.tga"), shadow map and the SSAA image have their own rules (empty buffers are skipped, the SSAA
// image is fetched from the device first)
void OutputFrameBuffer(), OutputZBuffer(), OutputShadowBuffer(), OutputSSAAImage();
#define FGL_OUTPUT_DECLARE(NAME, BUFFER, FILE) void Output##NAME();
FGL_OUTPUT_GBUFFER_DUMPS(FGL_OUTPUT_DECLARE)
#undef FGL_OUTPUT_DECLARE
}  // namespace Output
