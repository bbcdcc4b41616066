// render.h — namespace Render of the drop-in facade (reference src/render.h:19-38): pass sequencing.
#pragma once

#include "scene.h"

namespace Render
{
void Preconfigure(const Scene& scene);
void Render(const Scene& scene);
void RenderGeometryStage(const Scene& scene);  // shadow + raster (+ SSAO): everything before the lighting loop
void RenderLightingStage(const Scene& scene);  // lighting loop + SSAA
void DoShadowPass(const Scene& scene);
void DoForwardPass(const Scene& scene);
void DoGeometryPass(const Scene& scene);
void DoLightingPass(const Scene& scene);
void DoSSAO(const Scene& scene);
void DoSSAA(const Scene& scene);
}  // namespace Render
