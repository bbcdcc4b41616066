// shadow.cpp — Shadow status / mode forwarding (reference src/shaders/shadow.cpp:9-21).
#include "shadow.h"

#include "forkergl.h"
#include "forkergl_b200.h"

static bool         s_IsShadowOn = true;
static Shadow::Mode s_Mode = Shadow::PCSS;

namespace Shadow
{
void SetShadowStatus(bool status) { s_IsShadowOn = status; }
bool GetShadowStatus() { return s_IsShadowOn; }
void SetShadowMode(Mode mode) { s_Mode = mode; }
Mode GetShadowMode() { return s_Mode; }
}  // namespace Shadow
