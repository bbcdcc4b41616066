// shadow.h — namespace Shadow of the drop-in facade (reference src/shaders/shadow.h:25-45).
// The on/off status is forwarded to the device context; the filter (hard / PCF / PCSS), a compile-time macro in
// the reference (shadow.h:15-16), is a run-time mode here.  The filter arithmetic itself lives in
// csrc/shadow.cuh.
#pragma once

namespace Shadow
{
enum Mode { Hard = 0, PCF = 1, PCSS = 2 };
void SetShadowStatus(bool status);
bool GetShadowStatus();
void SetShadowMode(Mode mode);  // default PCSS, the reference's shipped configuration
Mode GetShadowMode();
}  // namespace Shadow
