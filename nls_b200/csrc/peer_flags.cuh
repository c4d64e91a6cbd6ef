// peer_flags.cuh -- system-scope flag helpers and the flag-block layout shared by the halo-exchange kernel (peer.cu) and
// by the step kernel that performs the exchange in its own epilogue (stream_2d.cu).
#pragma once

#include <cuda_runtime.h>

namespace nlsb {

// A rank's flag block (256 zero-initialised bytes from nlsb_peer_alloc, 64-byte spacing): 64-bit epochs written by
// the NEIGHBOURS with st.release.sys and polled by the owner with ld.acquire.sys.
//   READY(e) from a neighbour: "my kernels that read my halo rows before exchange e are done -- overwrite them"
//   DATA(e)  from a neighbour: "your halo rows now hold my boundary rows of exchange e"
constexpr int kReadyFromUp = 0, kReadyFromDown = 8, kDataFromUp = 16, kDataFromDown = 24;
// A rank's private state block (64 zero-initialised bytes): [0] exchanges completed (epoch), [1] ticket counter of
// the running exchange, [2] waits that timed out.

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// spin until *flag >= epoch; false when the wait timed out (a lost neighbour must not hang the device)
__device__ __forceinline__ bool wait_epoch(const unsigned long long *flag, unsigned long long epoch, long long timeout_cycles)
{
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < epoch) {
        if (clock64() - t0 > timeout_cycles) return false;
        __nanosleep(64);
    }
    return true;
}

}  // namespace nlsb
