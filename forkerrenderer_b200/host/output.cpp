// output.cpp — TGA dumps (reference src/output.cpp:12-86), same file names, written with the facade's RLE
// encoder so that files are byte-comparable with the oracle's.
#include "output.h"

#include "forkergl.h"

static std::string s_Dir = "output";

namespace Output
{
void SetDirectory(const std::string& dir) { s_Dir = dir; }

void OutputFrameBuffer() { ForkerGL::FrameBuffer.GenerateImage().WriteTgaFile(s_Dir + "/framebuffer.tga"); }
void OutputZBuffer() { ForkerGL::DepthBuffer.GenerateImage().WriteTgaFile(s_Dir + "/zbuffer.tga"); }
void OutputShadowBuffer()
{
    if (ForkerGL::ShadowBuffer.GetWidth() != 0)
        ForkerGL::ShadowBuffer.GenerateImage().WriteTgaFile(s_Dir + "/shadowmap.tga");
}
void OutputSSAAImage()
{
    ForkerGL::FetchAntiAliasedImage();
    if (ForkerGL::AntiAliasedImage.GetWidth() != 0)
        ForkerGL::AntiAliasedImage.WriteTgaFile(s_Dir + "/framebuffer_SSAA.tga");
}
#define FGL_OUT3(FN, BUF, FILE) \
    void FN() { if (ForkerGL::BUF.GetWidth() != 0) ForkerGL::BUF.GenerateImage().WriteTgaFile(s_Dir + FILE); }
FGL_OUT3(OutputNormalGBuffer, NormalGBuffer, "/gbuffer_normal.tga")
FGL_OUT3(OutputWorldPosGBuffer, WorldPosGBuffer, "/gbuffer_worldpos.tga")
FGL_OUT3(OutputAlbedoGBuffer, AlbedoGBuffer, "/gbuffer_albedo.tga")
FGL_OUT3(OutputParamGBuffer, ParamGBuffer, "/gbuffer_param.tga")
FGL_OUT3(OutputShadingTypeGBuffer, ShadingTypeGBuffer, "/gbuffer_shading_type.tga")
FGL_OUT3(OutputAmbientOcclusionGBuffer, AmbientOcclusionGBuffer, "/gbuffer_ambient_occlusion.tga")
}  // namespace Output
