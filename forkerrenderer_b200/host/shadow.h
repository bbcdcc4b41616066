// shadow.h — namespace Shadow of the drop-in facade (reference src/shaders/shadow.h:25-45).
// The on/off status is forwarded to the device context; the filter (hard / PCF / PCSS), a compile-time macro in
// the reference (shadow.h:15-16), is a run-time mode here.  The device passes filter in csrc/stream.cu / shade.cu;
// CalculateShadowVisibility is the same arithmetic on the host (host/programs.cpp) for the host-side fragment programs.
#pragma once

#include "buffer.h"
#include "geometry.h"

namespace Shadow
{
enum Mode { Hard = 0, PCF = 1, PCSS = 2 };
void SetShadowStatus(bool status);
bool GetShadowStatus();
void SetShadowMode(Mode mode);  // default PCSS, the reference's shipped configuration
Mode GetShadowMode();
// reference shadow.cpp:109-132 (Hard / PCF / PCSS by GetShadowMode(); consumes the host's own mt19937 stream)
Float CalculateShadowVisibility(const Buffer1f& shadowMap, const Vector3f& positionLightSpaceNDC, const Vector3f& normal, const Vector3f& lightDir);
void  ResetHostSampleStream();  // the host-side filters restart their mt19937 at its default seed (a fresh reference process)
}  // namespace Shadow
