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
#define FGL_OUTPUT_DEFINE(NAME, BUFFER, FILE) \
    void Output##NAME() { if (ForkerGL::BUFFER.GetWidth() != 0) ForkerGL::BUFFER.GenerateImage().WriteTgaFile(s_Dir + FILE); }
FGL_OUTPUT_GBUFFER_DUMPS(FGL_OUTPUT_DEFINE)
#undef FGL_OUTPUT_DEFINE
}  // namespace Output
