// oracle/ref_shadow_dispatch.cpp — TEST INFRASTRUCTURE ONLY.
//
// Defines the reference's `namespace Shadow` entry points (declared in reference src/shaders/shadow.h:25-45)
// as run-time dispatchers over the three compile-time variants built by ref_shadow_variant.cpp.
// The global mt19937 lives in the inline function Random01 (reference src/utility.h:90-98) and is therefore
// shared by all variants, exactly as in a single-variant build.
#include "shadow.h"
#include "buffer.h"

#define FGL_DECL(NS)                                                                              \
    namespace NS                                                                                  \
    {                                                                                             \
    Float SampleShadowMap(const Buffer1f&, const Vector2f&);                                      \
    Float HardShadow(const Buffer1f&, const Vector3f&, Float);                                    \
    Float PCF(const Buffer1f&, const Vector3f&, Float, Float);                                    \
    Float FindAverageBlockDepth(const Buffer1f&, const Vector3f&, Float);                         \
    Float PCSS(const Buffer1f&, const Vector3f&, Float);                                          \
    Float CalculateShadowVisibility(const Buffer1f&, const Vector3f&, const Vector3f&,            \
                                    const Vector3f&);                                             \
    }
FGL_DECL(ShadowHard)
FGL_DECL(ShadowPCF)
FGL_DECL(ShadowPCSS)

int g_fglRefShadowMode = 2;  // 0 hard, 1 PCF, 2 PCSS (the reference's shipped default)

static bool s_status = true;  // reference default, src/shaders/shadow.cpp:9

namespace Shadow
{
void SetShadowStatus(bool status) { s_status = status; }
bool GetShadowStatus() { return s_status; }

Float SampleShadowMap(const Buffer1f& m, const Vector2f& uv) { return ShadowPCSS::SampleShadowMap(m, uv); }
Float HardShadow(const Buffer1f& m, const Vector3f& c, Float b) { return ShadowPCSS::HardShadow(m, c, b); }
Float PCF(const Buffer1f& m, const Vector3f& c, Float b, Float f) { return ShadowPCSS::PCF(m, c, b, f); }
Float FindAverageBlockDepth(const Buffer1f& m, const Vector3f& c, Float b)
{
    return ShadowPCSS::FindAverageBlockDepth(m, c, b);
}
Float PCSS(const Buffer1f& m, const Vector3f& c, Float b) { return ShadowPCSS::PCSS(m, c, b); }

Float CalculateShadowVisibility(const Buffer1f& shadowMap, const Vector3f& ndc, const Vector3f& n,
                                const Vector3f& l)
{
    switch (g_fglRefShadowMode)
    {
        case 0: return ShadowHard::CalculateShadowVisibility(shadowMap, ndc, n, l);
        case 1: return ShadowPCF::CalculateShadowVisibility(shadowMap, ndc, n, l);
        default: return ShadowPCSS::CalculateShadowVisibility(shadowMap, ndc, n, l);
    }
}
}  // namespace Shadow
