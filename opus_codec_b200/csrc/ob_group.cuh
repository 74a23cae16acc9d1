// ob_group.cuh -- the cooperating set of lanes a device function runs on: a warp (band reconstruction),
// a thread block (synthesis), or -- when this code is compiled by g++ for tests/host_emul -- one lane.
// All cooperative device code is written against this interface: strided loops `for (j = g.lane; j < n;
// j += g.n)`, g.sync() where lanes exchange data through shared memory, g.sum()/g.sum_u32() all-reduces.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
struct ObWarp {
    int lane;
    static constexpr int n = 32;
    __device__ __forceinline__ ObWarp() : lane((int)(threadIdx.x & 31)) {}
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ float sum(float v) const
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ __forceinline__ uint32_t sum_u32(uint32_t v) const
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
};

// A whole thread block; red points at >= 32 floats of shared memory reserved for reductions.
struct ObBlock {
    int lane, n;
    float *red;
    __device__ __forceinline__ ObBlock(float *r) : lane((int)threadIdx.x), n((int)blockDim.x), red(r) {}
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ float sum(float v) const
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((lane & 31) == 0) red[lane >> 5] = v;
        __syncthreads();
        float t = 0.f;
        for (int w = 0; w < (n + 31) >> 5; w++) t += red[w];
        return t;
    }
    // Scan of affine maps m -> a*m + b, lane order = application order.  (ea, eb): the composition of all lanes BEFORE this one (identity for
    // lane 0); (ta, tb): of all lanes.  Shuffles inside a warp, the warps' totals through `red`: two barriers instead of two per doubling step.
    __device__ __forceinline__ void affine_scan(float a, float b, float &ea, float &eb, float &ta, float &tb) const
    {
        const int wl = lane & 31, w = lane >> 5, nw = (n + 31) >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float pa = __shfl_up_sync(0xffffffffu, a, o), pb = __shfl_up_sync(0xffffffffu, b, o);
            if (wl >= o) { b = a * pb + b; a = a * pa; }
        }
        float ia = __shfl_up_sync(0xffffffffu, a, 1), ib = __shfl_up_sync(0xffffffffu, b, 1);
        if (wl == 0) { ia = 1.f; ib = 0.f; }
        __syncthreads();
        if (wl == 31) { red[2 * w] = a; red[2 * w + 1] = b; }
        __syncthreads();
        float ca = 1.f, cb = 0.f;
        ea = ia; eb = ib;
        for (int v = 0; v < nw; v++) {
            if (v == w) { ea = ia * ca; eb = ia * cb + ib; }
            cb = red[2 * v] * cb + red[2 * v + 1]; ca = red[2 * v] * ca;
        }
        ta = ca; tb = cb;
    }
};
#endif

// A single lane: the whole "group" is one thread (thread-per-stream device code, and the host emulation in tests/).
#ifdef __CUDACC__
#define OB_SOLO_FN __host__ __device__ __forceinline__
#else
#define OB_SOLO_FN inline
#endif
struct ObSolo {
    static constexpr int lane = 0;
    static constexpr int n = 1;
    OB_SOLO_FN void sync() const {}
    OB_SOLO_FN void pace(int) const {}
    OB_SOLO_FN void set_base(int) const {}
    OB_SOLO_FN float sum(float v) const { return v; }
    OB_SOLO_FN uint32_t sum_u32(uint32_t v) const { return v; }
    OB_SOLO_FN void affine_scan(float a, float b, float &ea, float &eb, float &ta, float &tb) const { ea = 1.f; eb = 0.f; ta = a; tb = b; }
};
