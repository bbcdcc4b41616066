// stream.h — replay of the reference's single global mt19937 sample stream (utility.h:90-103) on the device.
#pragma once
#include "fgl_internal.h"

void fgl_stream_destroy(fgl_ctx* c);
// Render::Render starts a frame at stream position 0 (the reference process renders exactly one frame).
void fgl_stream_begin_frame(fgl_ctx* c);
// SSAO consumes the stream first (render.cpp:204-209): pixel p takes accepted unit-ball samples 32p .. 32p+31.
int fgl_stream_prepare_ssao(fgl_ctx* c, SsaoPass& S);
// Deferred lighting continues where SSAO stopped: PCF takes 64 accepted unit-disk samples per pixel, PCSS 32 plus
// 64 more iff the pixel's blocker search found a blocker (shadow.cpp:92-106) — a frame-long dependency chain that
// is resolved here into a per-pixel chunk index.
// phase: the part of the PCSS resolution that does not depend on earlier bands (PREPARE), the rest (RESOLVE), or both.
// LAUNCH: PREPARE (if still to do) plus the issue of the chain kernel itself, on whatever stream c->stream is at the time
// (fgl_prepare_screen_space_pixels points it at the context's chain stream so that SSAO and the blur overlap the chain).
enum { FGL_VIS_ALL = 0, FGL_VIS_PREPARE = 1, FGL_VIS_RESOLVE = 2, FGL_VIS_LAUNCH = 3 };
int fgl_stream_prepare_lighting(fgl_ctx* c, LightPass& L, int phase = FGL_VIS_ALL);
// Generic form: n consumers in consumption order with their shadow coordinate + bias (device array); leaves L.vis set.
int fgl_stream_site_visibility(fgl_ctx* c, LightPass& L, size_t n, const float4* sc4, size_t siteLo, size_t siteHi, unsigned long long blockersBefore,
                               int phase = FGL_VIS_ALL);
int fgl_stream_chain_total(fgl_ctx* c, unsigned long long* out);
// Device-side hand-off of the chain state through peer memory (include/forkergl_b200.h: fgl_chain_peer_*)
int fgl_stream_peer_mailbox(fgl_ctx* c, void** devPtr, void* ipcHandle64);
int fgl_stream_peer_connect(fgl_ctx* c, void* nextDevPtr, const void* nextIpcHandle64, int waitPrev, int enable);
bool fgl_stream_peer_on(fgl_ctx* c);
