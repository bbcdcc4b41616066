// render.h — namespace Render of the drop-in facade: the pass sequence of the reference (src/render.h:19-38) under its
// own names, plus the two halves a sort-first multi-GPU driver runs on either side of the PCSS chain hand-off.
#pragma once

#include "scene.h"

namespace Render
{
// reference surface: Preconfigure, then Render = shadow pass, forward | geometry + lighting (+ SSAO), SSAA
void Preconfigure(const Scene& scene);
void Render(const Scene& scene);
void DoShadowPass(const Scene& scene), DoForwardPass(const Scene& scene), DoGeometryPass(const Scene& scene), DoLightingPass(const Scene& scene);
void DoSSAO(const Scene& scene), DoSSAA(const Scene& scene);

// additions: Render(scene) == RenderGeometryStage(scene) followed by RenderLightingStage(scene)
void RenderGeometryStage(const Scene& scene);  // shadow + raster (+ SSAO): everything before the lighting loop
void RenderLightingStage(const Scene& scene);  // lighting loop + SSAA
}  // namespace Render
