// oracle/ref_shadow_variant.cpp — TEST INFRASTRUCTURE ONLY.
//
// The reference selects its shadow filter with compile-time macros (reference src/shaders/shadow.h:15-16,
// consumed only at src/shaders/shadow.cpp:120-129).  To get all three filters into ONE reference binary
// without editing or copying the reference sources, this wrapper compiles the reference's own shadow.cpp
// three times, each time into a differently named namespace (the token `Shadow` is renamed by the
// preprocessor) and with the mode macros re-defined after the reference header has been seen.
// `#pragma once` in shadow.h makes the second inclusion (from inside shadow.cpp) a no-op, so the
// re-definitions below are the ones shadow.cpp sees.
//
//   -DFGL_VARIANT=0 -DFGL_VARIANT_NS=ShadowHard    hard shadow   (both macros off)
//   -DFGL_VARIANT=1 -DFGL_VARIANT_NS=ShadowPCF     PCF           (SOFT_SHADOW_PCF)
//   -DFGL_VARIANT=2 -DFGL_VARIANT_NS=ShadowPCSS    PCSS          (as shipped)
#define Shadow FGL_VARIANT_NS
#include "shadow.h"
#undef SOFT_SHADOW_PCF
#undef SOFT_SHADOW_PCSS
#if FGL_VARIANT == 1
#define SOFT_SHADOW_PCF
#elif FGL_VARIANT == 2
#define SOFT_SHADOW_PCSS
#endif
#include "shadow.cpp"
