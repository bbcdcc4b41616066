// main.cpp — CLI with the reference's command line (reference src/main.cpp:18-55):
//   forkerrenderer <scene file> [--shadow hard|pcf|pcss] [--wrap N] [--filter N]
// Scene -> Preconfigure -> Render -> TGA dumps into ./output/.
#include <cstdio>
#include <cstring>
#include <exception>
#include <string>

#include "forkergl.h"
#include "output.h"
#include "render.h"
#include "shadow.h"

int main(int argc, const char* argv[])
{
    if (argc < 2)
    {
        fprintf(stderr, "Usage: %s <scene file> [--shadow hard|pcf|pcss] [--wrap 0..3] [--filter 0|1]\n", argv[0]);
        return 1;
    }
    try
    {
        for (int i = 2; i + 1 < argc; i += 2)
        {
            std::string k = argv[i], v = argv[i + 1];
            if (k == "--shadow") Shadow::SetShadowMode(v == "hard" ? Shadow::Hard : v == "pcf" ? Shadow::PCF : Shadow::PCSS);
            else if (k == "--wrap") ForkerGL::TextureWrapMode((Texture::WrapMode)atoi(v.c_str()));
            else if (k == "--filter") ForkerGL::TextureFilterMode((Texture::FilterMode)atoi(v.c_str()));
        }
        Scene scene(argv[1]);
        if (!scene.IsValid()) return 2;
        Render::Preconfigure(scene);
        Render::Render(scene);
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
        ForkerGL::Shutdown();
    }
    catch (const std::exception& e)
    {
        fprintf(stderr, "forkerrenderer: %s\n", e.what());
        return 3;
    }
    return 0;
}
