// group.cu — the sort-first group: several contexts, one per GPU of an NVLink / NVSwitch box, render one frame together.
//
// Every context owns a contiguous band of rows — of the screen AND of the shadow map (same size, render.cpp:62-66).  What a
// band's passes produce and another band needs is STORED INTO THE OTHER CONTEXTS' PLANES BY THE PRODUCING KERNEL ITSELF, over
// NVLink (peer pointers obtained through CUDA IPC, or plain pointers inside one process):
//     k_resolve_shadow      rows of the shadow map          -> every context's ShadowBuffer        (fused all-gather, 4 B / texel)
//     k_resolve_geometry    rows of the camera depth plane   -> every context's DepthBuffer         (SSAO gathers depth anywhere)
//     k_lighting_hard       rows of the finished 8-bit frame -> rank 0's frame                       (fused gather, 3 B / pixel)
//     k_peer_notify         the PCSS chain's blocker count   -> the next band's mailbox              (stream.cu)
// and ordering is kept by epoch flags (GroupFlagsD): a one-block kernel stores the frame number into its slot in the peers'
// flag arrays behind the producing kernel (system-scope release), a one-warp kernel in front of the first consumer spins on
// the slots of all peers (acquire).  No host round trip, no NCCL call, nothing replicated but the (culled) per-triangle set-up:
//     begin_frame      signal READY        "frame e may be written into my planes" (the previous frame has been consumed)
//     shadow pass      set-up culls to the band's shadow rows, raster, wait READY, resolve -> peers, signal SHADOW
//     geometry pass    set-up culls to the band's rows + halo, raster, resolve (G-buffer local, depth -> peers), signal DEPTH
//     SSAO             wait DEPTH;   lighting preparation / lighting: wait SHADOW
//     lighting         RGB8 rows -> rank 0, signal BAND (rank 0's slot array only);  rank 0 reads the frame after wait BAND
// A wait gives up after 30 s (error word, reported by the next host synchronisation) instead of hanging the device.
#include <cstring>

#include "fgl_internal.h"

namespace
{
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct FlagTargets
{
    unsigned long long* slot[kMaxGroup];  // &peerFlags[r]->kind[myRank], nullptr = skip
    int                 n;
};
// lane r stores the epoch into peer r's slot; the kernel boundary before it + the system fence order it behind this context's
// peer stores of the producing kernel
__global__ void k_group_signal(FlagTargets T, unsigned long long epoch)
{
    int r = threadIdx.x;
    if (r >= T.n || !T.slot[r]) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(T.slot[r]), "l"(epoch) : "memory");
}
// lane r waits for slot r (written by peer r) to reach the epoch
__global__ void k_group_wait(const unsigned long long* slots, int n, unsigned long long epoch, unsigned long long* error)
{
    int r = threadIdx.x;
    if (r < n)
    {
        const unsigned long long t0 = timer_ns();
        while (ld_acquire_sys(slots + r) < epoch)
        {
            __nanosleep(100);
            if (timer_ns() - t0 > 30000000000ull)
            {
                *error = 1ull + (unsigned long long)r;
                break;
            }
        }
    }
    __syncwarp();
    __threadfence_system();
}
}  // namespace

void fgl_group_band(const fgl_ctx* c, int H, int& r0, int& r1)
{
    const GroupState& g = c->group;
    int               per = (H + g.world - 1) / g.world;
    r0 = std::min(H, g.rank * per), r1 = std::min(H, r0 + per);
}

static unsigned long long* slot_of(GroupFlagsD* f, int what, int rank)
{
    switch (what)
    {
    case FGL_GROUP_READY: return f->ready + rank;
    case FGL_GROUP_SHADOW: return f->shadow + rank;
    case FGL_GROUP_DEPTH: return f->depth + rank;
    default: return f->band + rank;
    }
}

int fgl_group_signal(fgl_ctx* c, int what, bool rootOnly)
{
    GroupState& g = c->group;
    if (!g.on) return FGL_OK;
    FlagTargets T;
    memset(&T, 0, sizeof T);
    T.n = g.world;
    for (int r = 0; r < g.world; ++r)
        if (!rootOnly || r == 0) T.slot[r] = slot_of(g.peerFlags[r], what, g.rank);
    LaunchScope ls(c, "group_signal", 0);
    k_group_signal<<<1, 32, 0, c->stream>>>(T, g.epoch);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fgl_fail(c, FGL_ERR_CUDA, std::string("group signal: ") + cudaGetErrorString(e));
    return FGL_OK;
}

int fgl_group_wait(fgl_ctx* c, int what)
{
    GroupState& g = c->group;
    if (!g.on) return FGL_OK;
    bool& done = what == FGL_GROUP_READY ? g.readyWaited : what == FGL_GROUP_SHADOW ? g.shadowWaited : what == FGL_GROUP_DEPTH ? g.depthWaited : g.bandWaited;
    if (done) return FGL_OK;
    done = true;
    GroupFlagsD* mine = (GroupFlagsD*)g.flags.p;
    LaunchScope  ls(c, "group_wait", 0);
    k_group_wait<<<1, 32, 0, c->stream>>>(slot_of(mine, what, 0), g.world, g.epoch, &mine->error);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fgl_fail(c, FGL_ERR_CUDA, std::string("group wait: ") + cudaGetErrorString(e));
    return FGL_OK;
}

void fgl_group_release(fgl_ctx* c)
{
    GroupState& g = c->group;
    for (void* p : g.ipcMapped) cudaIpcCloseMemHandle(p);
    g.ipcMapped.clear();
    g.on = false;
}
