// common.cuh -- shared helpers for the sm_100a kernels of the Marius hot path.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <string>

#include "../../include/marius_b200.h"

namespace mb {

constexpr int kWarp = 32;

// --- error plumbing -------------------------------------------------------------------------------
void set_error(const std::string& msg);
extern std::atomic<uint64_t> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define MB_CUDA_TRY(expr)                                                                                   \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess) {                                                                            \
            ::mb::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                            std::to_string(__LINE__) + ")");                                                \
            return MB_ERR_CUDA;                                                                             \
        }                                                                                                   \
    } while (0)

#define MB_TRY(expr)                    \
    do {                                \
        mb_status _s = (expr);          \
        if (_s != MB_OK) return _s;     \
    } while (0)

#define MB_REQUIRE(cond, msg)                                          \
    do {                                                               \
        if (!(cond)) {                                                 \
            ::mb::set_error(std::string("invalid argument: ") + msg);  \
            return MB_ERR_INVALID;                                     \
        }                                                              \
    } while (0)

// launch + count + check
#define MB_LAUNCH_CHECK()                                 \
    do {                                                  \
        ::mb::count_launch();                             \
        MB_CUDA_TRY(cudaGetLastError());                  \
    } while (0)

int sm_count();  // SMs of the current device (148 on B200)

// --- device helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// streaming 128-bit global accesses (rows are touched once per batch: keep them out of L1)
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ld_f4(const float4* p) { return *p; }
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Adagrad rule of Batch::accumulateGradients (data/batch.cpp:67-69), one element, op-for-op in fp32 with IEEE
// round-to-nearest sqrt/div and NO fma contraction, so the result only depends on (g, s, lr):
//   ds = g*g ; s' = s + ds ; de = (-lr) * (g / (sqrt(s') + 1e-10f))
__device__ __forceinline__ void adagrad_rule(float g, float s, float neg_lr, float& de, float& ds, float& s_new) {
    ds = __fmul_rn(g, g);
    s_new = __fadd_rn(s, ds);
    float den = __fadd_rn(__fsqrt_rn(s_new), 1e-10f);
    de = __fmul_rn(neg_lr, __fdiv_rn(g, den));
}

}  // namespace mb
