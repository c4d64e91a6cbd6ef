// peer.cu -- peer-mapped slab buffers and the DEVICE-INITIATED halo exchange of the multi-GPU 2D solver.
//
// The reference has no parallelism; BASELINE.json asks for the 8192^2 grid cut into row slabs over
// the GPUs of one box with the halos travelling over NVLink.  One process per GPU: every rank allocates its two psi
// buffers and a small flag block with nlsb_peer_alloc (plain cudaMalloc, so the allocation has an IPC handle), the
// handles are exchanged once through torch.distributed, and every rank maps its two neighbours' buffers
// (cudaIpcOpenMemHandle; NVLink peer access).  From then on a halo exchange is ONE kernel per rank and needs no
// host work, so the m-step cycle (m RK4 launches + exchange) is captured in a CUDA graph:
//
//   halo_exchange_kernel (a few CTAs)
//     1. publish READY(e) in both neighbours' flag blocks: "my kernels that read my halo rows are done, you may
//        overwrite them" (stream order guarantees it: the exchange kernel follows the step kernels);
//     2. wait for the neighbours' READY(e), then copy my boundary rows straight into their halo rows with 16-byte
//        stores on the peer-mapped pointers (NVLink), fence at system scope;
//     3. the CTA that finishes last publishes DATA(e) in both neighbours' flag blocks and waits for their DATA(e):
//        when the kernel exits, my halo rows hold the neighbours' rows of epoch e.
//   e is a counter kept in device memory and advanced by the kernel itself, so graph replays need no new arguments.
//   Every wait gives up after about two seconds and records the failure in the state block (a lost neighbour must
//   not hang the GPU); nlsb_dev_halo_status reads it back.
//
// Flags are 64-bit epochs written with st.release.sys and polled with ld.acquire.sys; a rank's own flag block lives
// in its own memory (the poll never crosses NVLink).

#include "kernels.h"
#include "peer_flags.cuh"

#include <cstdint>

namespace nlsb {

namespace {

struct HaloArgs {
    const double2 *src_up;      // my first `halo` owned rows -> lower halo of the rank above (null: no such rank)
    double2 *dst_up;            // peer-mapped
    const double2 *src_down;    // my last `halo` owned rows -> upper halo of the rank below
    double2 *dst_down;
    unsigned long long count;   // double2 elements per direction
    unsigned long long *state;  // local: [0] epoch, [1] ticket, [2] timeouts seen
    unsigned long long *mine;   // local flag block: [0] READY from up, [8] READY from down, [16] DATA from up, [24] DATA from down
    unsigned long long *up;     // the flag block of the rank above (peer-mapped), or null
    unsigned long long *down;
    long long timeout_cycles;
};

__global__ void __launch_bounds__(256) halo_exchange_kernel(const HaloArgs a)
{
    __shared__ int s_last;
    const unsigned long long epoch = a.state[0] + 1;      // the same for every CTA: state[0] changes after the last ticket
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) {
            // 1. my halo rows may be overwritten (what the rank ABOVE reads as "ready from down" and vice versa)
            if (a.up) st_release_sys(a.up + kReadyFromDown, epoch);
            if (a.down) st_release_sys(a.down + kReadyFromUp, epoch);
        }
        bool ok = true;
        if (a.up) ok = wait_epoch(a.mine + kReadyFromUp, epoch, a.timeout_cycles) && ok;
        if (a.down) ok = wait_epoch(a.mine + kReadyFromDown, epoch, a.timeout_cycles) && ok;
        if (!ok) atomicAdd(a.state + 2, 1ull);
    }
    __syncthreads();

    // 2. boundary rows -> the neighbours' halo rows (peer stores)
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.count; i += stride) {
        if (a.up) a.dst_up[i] = a.src_up[i];
        if (a.down) a.dst_down[i] = a.src_down[i];
    }
    __threadfence_system();
    __syncthreads();

    // 3. last CTA: publish DATA(e), wait for the neighbours' DATA(e), advance the epoch
    if (threadIdx.x == 0) {
        const unsigned long long ticket = atomicAdd(a.state + 1, 1ull);
        s_last = ticket == gridDim.x - 1;
        if (s_last) {
            __threadfence_system();
            if (a.up) st_release_sys(a.up + kDataFromDown, epoch);
            if (a.down) st_release_sys(a.down + kDataFromUp, epoch);
            bool ok = true;
            if (a.up) ok = wait_epoch(a.mine + kDataFromUp, epoch, a.timeout_cycles) && ok;
            if (a.down) ok = wait_epoch(a.mine + kDataFromDown, epoch, a.timeout_cycles) && ok;
            if (!ok) atomicAdd(a.state + 2, 1ull);
            a.state[1] = 0;
            __threadfence();
            a.state[0] = epoch;
        }
    }
}

}  // namespace

int launch_halo_exchange(const double2 *src_up, double2 *dst_up, const double2 *src_down, double2 *dst_down,
                         size_t count, unsigned long long *state, unsigned long long *mine, unsigned long long *up,
                         unsigned long long *down, double timeout_seconds, cudaStream_t stream)
{
    HaloArgs a{};
    a.src_up = up ? src_up : nullptr; a.dst_up = up ? dst_up : nullptr;
    a.src_down = down ? src_down : nullptr; a.dst_down = down ? dst_down : nullptr;
    a.count = count; a.state = state; a.mine = mine; a.up = up; a.down = down;
    a.timeout_cycles = (long long)(timeout_seconds * 1.9e9);
    // enough CTAs to keep NVLink busy (a few MiB per exchange), few enough to start at once beside other work
    size_t ctas = (count + 256 * 16 - 1) / (256 * 16);
    if (ctas < 1) ctas = 1;
    if (ctas > 32) ctas = 32;
    halo_exchange_kernel<<<(unsigned)ctas, 256, 0, stream>>>(a);
    count_launches(1);
    return (int)cudaGetLastError();
}

}  // namespace nlsb

using namespace nlsb;

extern "C" {

int nlsb_peer_alloc(size_t bytes, void **ptr)
{
    if (!ptr || bytes == 0) return fail(NLSB_EINVAL, "peer_alloc: bad arguments");
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail((int)e, "peer_alloc: cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    e = cudaMemset(p, 0, bytes);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail((int)e, "peer_alloc: cudaMemset: %s", cudaGetErrorString(e));
    }
    *ptr = p;
    return 0;
}

int nlsb_peer_free(void *ptr)
{
    cudaError_t e = cudaFree(ptr);
    return e == cudaSuccess ? 0 : fail((int)e, "peer_free: %s", cudaGetErrorString(e));
}

int nlsb_peer_export(const void *ptr, unsigned char *handle64)
{
    if (!ptr || !handle64) return fail(NLSB_EINVAL, "peer_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void *>(ptr));
    if (e != cudaSuccess) return fail((int)e, "peer_export: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    for (int i = 0; i < 64; ++i) handle64[i] = reinterpret_cast<const unsigned char *>(&h)[i];
    return 0;
}

int nlsb_peer_open(const unsigned char *handle64, void **ptr)
{
    if (!handle64 || !ptr) return fail(NLSB_EINVAL, "peer_open: null argument");
    cudaIpcMemHandle_t h;
    for (int i = 0; i < 64; ++i) reinterpret_cast<unsigned char *>(&h)[i] = handle64[i];
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail((int)e, "peer_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
    *ptr = p;
    return 0;
}

int nlsb_peer_close(void *ptr)
{
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    return e == cudaSuccess ? 0 : fail((int)e, "peer_close: %s", cudaGetErrorString(e));
}

int nlsb_peer_enable_access(int peer_device)
{
    int dev = 0, can = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail((int)e, "peer_enable_access: %s", cudaGetErrorString(e));
    if (peer_device == dev) return 0;
    e = cudaDeviceCanAccessPeer(&can, dev, peer_device);
    if (e != cudaSuccess) return fail((int)e, "peer_enable_access: %s", cudaGetErrorString(e));
    if (!can) return fail(NLSB_EINVAL, "device %d cannot access device %d", dev, peer_device);
    e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        return 0;
    }
    return e == cudaSuccess ? 0 : fail((int)e, "peer_enable_access: %s", cudaGetErrorString(e));
}

int nlsb_dev_halo_exchange(const double *src_up, double *dst_up, const double *src_down, double *dst_down,
                           size_t complex_count, void *state, void *flags_mine, void *flags_up, void *flags_down,
                           double timeout_seconds, nlsb_stream_t stream)
{
    if (!state || !flags_mine) return fail(NLSB_EINVAL, "dev_halo_exchange: null state / flag block");
    if ((flags_up && (!src_up || !dst_up)) || (flags_down && (!src_down || !dst_down)))
        return fail(NLSB_EINVAL, "dev_halo_exchange: a neighbour without row pointers");
    if (!(timeout_seconds > 0.0)) timeout_seconds = 2.0;
    int rc = launch_halo_exchange(reinterpret_cast<const double2 *>(src_up), reinterpret_cast<double2 *>(dst_up),
                                  reinterpret_cast<const double2 *>(src_down), reinterpret_cast<double2 *>(dst_down),
                                  complex_count, static_cast<unsigned long long *>(state),
                                  static_cast<unsigned long long *>(flags_mine), static_cast<unsigned long long *>(flags_up),
                                  static_cast<unsigned long long *>(flags_down), timeout_seconds,
                                  static_cast<cudaStream_t>(stream));
    if (rc) return fail(rc, "dev_halo_exchange: %s", cudaGetErrorString((cudaError_t)rc));
    return 0;
}

int nlsb_dev_halo_status(const void *state, unsigned long long *epoch, unsigned long long *timeouts)
{
    if (!state) return fail(NLSB_EINVAL, "dev_halo_status: null state");
    unsigned long long host[3] = {0, 0, 0};
    cudaError_t e = cudaMemcpy(host, state, sizeof(host), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail((int)e, "dev_halo_status: %s", cudaGetErrorString(e));
    if (epoch) *epoch = host[0];
    if (timeouts) *timeouts = host[2];
    return 0;
}

void nlsb_add_kernel_launches(unsigned long long n) { count_launches(n); }

}  // extern "C"
