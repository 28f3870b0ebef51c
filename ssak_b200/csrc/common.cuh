// common.cuh -- device helpers shared by the sm_100a CTC lattice kernels.
//
// PTX wrappers for the pieces of the Blackwell execution model the kernels use:
//   * mbarrier (init / arrive.expect_tx / try_wait.parity) as the completion mechanism of
//   * cp.async.bulk (1-D TMA bulk copy, SASS UBLKCP) global -> shared, used to prefetch
//     emission rows for the next frames into a shared-memory ring,
//   * MUFU ex2 / lg2 for the log2-domain log-sum-exp,
// plus the emission-ring bookkeeping (aligned super-range copies so that rows of any
// element alignment can be moved by the 16-byte-granular bulk copy engine).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ssak {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
// Finite stand-in for log(0): keeps (-inf) - (-inf) NaNs out of the log-sum-exp without
// branches.  ex2(kNeg - x) == 0 for every finite x that can occur, and kNeg absorbs the
// O(T) additions it sees, so an unreachable state stays "minus infinity".
constexpr float kNeg = -1.0e30f;
constexpr float kNegTest = -1.0e29f;  // value < kNegTest  <=>  log(0)

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// log2(2^a + 2^b): one MUFU.EX2 + one MUFU.LG2.
__device__ __forceinline__ float lse2(float a, float b) {
    const float m = fmaxf(a, b);
    const float d = fminf(a, b) - m;
    return m + lg2_approx(1.0f + ex2_approx(d));
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
                 : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// non-blocking probe of the phase with the given parity
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// named barrier among a subset of the CTA's warps (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ------------------------------------------------------------------- bulk copy (1-D TMA)
// global -> shared::cta, completion counted in bytes on `bar`.  dst, src 16-byte aligned,
// bytes a positive multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// --------------------------------------------------------------------------- row ring
// A ring of `stages` chunks of `chunk` frame rows each.  A row of V floats that starts at an
// arbitrary 4-byte-aligned global address is fetched as the enclosing 16-byte-aligned byte
// range; its first element then sits (addr & 15) bytes into the slot.
struct RowRing {
    unsigned char *slots;  // stages * chunk * slot_bytes, 16-byte aligned
    uint64_t *full;        // one mbarrier per stage
    int chunk;             // frames per stage
    int stages;
    int slot_bytes;        // >= round16(4V) + 32
    int row_bytes;         // 4V
};

__host__ __device__ __forceinline__ int ring_slot_bytes(int V) {
    return ((4 * V + 15) & ~15) + 32;
}

// Producer-side cursor: the next chunk to fetch (kept by the one thread that issues copies).
struct RingProducer {
    const float *src;    // global address of the first frame row of the next chunk
    int64_t step_elems;  // elements between consecutive frames (negative: time runs backwards)
    int stage;           // stage the next chunk goes to
    int remaining;       // frames not issued yet
};

// Issue the copies of the next chunk (executed by ONE thread).
__device__ __forceinline__ void ring_issue_next(const RowRing &r, RingProducer &pr) {
    const int n = pr.remaining < r.chunk ? pr.remaining : r.chunk;
    if (n <= 0) return;
    uint32_t total = 0;
    const float *s = pr.src;
    for (int f = 0; f < n; ++f, s += pr.step_elems) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(s);
        total += (uint32_t)(((a + r.row_bytes + 15) & ~(uintptr_t)15) - (a & ~(uintptr_t)15));
    }
    mbar_arrive_expect_tx(&r.full[pr.stage], total);
    unsigned char *dst = r.slots + (size_t)pr.stage * r.chunk * r.slot_bytes;
    s = pr.src;
    for (int f = 0; f < n; ++f, s += pr.step_elems, dst += r.slot_bytes) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(s);
        const uintptr_t a0 = a & ~(uintptr_t)15;
        bulk_g2s(dst, reinterpret_cast<const void *>(a0),
                 (uint32_t)(((a + r.row_bytes + 15) & ~(uintptr_t)15) - a0), &r.full[pr.stage]);
    }
    pr.src = s;
    pr.remaining -= n;
    if (++pr.stage == r.stages) pr.stage = 0;
}

// Consumer-side cursor over the frames of the ring: no divisions in the time loop.
struct RingPos {
    int f, stage, phase, slot;
    uint32_t a15, a15_step;
    __device__ __forceinline__ void init(const float *first_row, int64_t step_elems) {
        f = stage = phase = slot = 0;
        a15 = (uint32_t)(reinterpret_cast<uintptr_t>(first_row) & 15);
        a15_step = (uint32_t)((step_elems * 4) & 15);
    }
    // wait for the chunk when entering it, return the V floats of the current frame
    __device__ __forceinline__ const float *row(const RowRing &r) const { return row(r, r.full); }
    // same, waiting on another per-stage barrier array (e.g. "rows post-processed")
    __device__ __forceinline__ const float *row(const RowRing &r, uint64_t *bars) const {
        if (f == 0) mbar_wait(&bars[stage], (uint32_t)phase);
        return reinterpret_cast<const float *>(r.slots + (size_t)slot * r.slot_bytes + a15);
    }
    __device__ __forceinline__ bool last_of_chunk(const RowRing &r) const { return f == r.chunk - 1; }
    __device__ __forceinline__ void advance(const RowRing &r) {
        a15 = (a15 + a15_step) & 15u;
        ++slot;
        if (++f == r.chunk) {
            f = 0;
            if (++stage == r.stages) {
                stage = 0;
                phase ^= 1;
                slot = 0;
            }
        }
    }
};

// Order-preserving float <-> int map, so that a warp-wide float max is one integer REDUX.
__device__ __forceinline__ int float_order_key(float x) {
    const int b = __float_as_int(x);
    return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float float_from_order_key(int k) {
    return __int_as_float(k ^ ((k >> 31) & 0x7fffffff));
}
__device__ __forceinline__ float warp_max(float x) {
    return float_from_order_key(__reduce_max_sync(0xffffffffu, float_order_key(x)));
}

}  // namespace ssak

// --------------------------------------------------------------------- host-side helpers
#include "../../include/ssak_b200.h"

namespace ssak {
void set_last_cuda_error(cudaError_t e);
inline int check_launch() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_cuda_error(e);
        return SSAK_ERR_CUDA;
    }
    return SSAK_OK;
}
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Number of SMs of the current device (cached per device ordinal; 148 on a B200).  The launch-shape heuristics
// derive the "one CTA per SM" thresholds from it instead of hard-coding the B200's count.
inline int device_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    const int slot = dev & 63;
    int n = cached[slot];
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[slot] = n;   // (a benign race: every thread writes the same value)
    }
    return n;
}

constexpr int kMaxDynSmem = 227 * 1024;
// Opt a kernel instantiation into the full 227 KB of dynamic shared memory ONCE per device.  The attribute is
// per-function, per-device state: setting it to the exact size of every launch let two host threads with different
// shapes interleave "set" and "launch" (one launch then failed with invalid-value).
template <auto Kern>
inline cudaError_t ensure_max_smem() {
    static unsigned long long done = 0;   // bit per device ordinal
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (__atomic_load_n(&done, __ATOMIC_ACQUIRE) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
    if (e == cudaSuccess) __atomic_fetch_or(&done, bit, __ATOMIC_RELEASE);
    return e;
}
}  // namespace ssak
