// output.h — namespace Output of the drop-in facade (reference src/output.h:7-20): TGA dumps of every buffer.
#pragma once
#include <string>

namespace Output
{
void SetDirectory(const std::string& dir);  // default "output" (reference output.cpp:14)
void OutputFrameBuffer();
void OutputZBuffer();
void OutputShadowBuffer();
void OutputSSAAImage();
void OutputNormalGBuffer();
void OutputWorldPosGBuffer();
void OutputAlbedoGBuffer();
void OutputParamGBuffer();
void OutputShadingTypeGBuffer();
void OutputAmbientOcclusionGBuffer();
}  // namespace Output
