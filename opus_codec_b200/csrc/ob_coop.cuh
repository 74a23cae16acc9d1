// ob_coop.cuh -- cross-lane primitives of the warp-per-stream ENCODER: reductions, arg-max, first- and second-order
// linear-recurrence scans, integer suffix sums.  Every cooperative stage of enc_*.cuh is written against these, so that
// the same source runs as
//   ObWarp   32 lanes of a CUDA warp (the product);
//   ObSolo   one host lane, sums and recurrences in the REFERENCE's order: the stage is then bit-identical to the reference's
//            pure-C build (tests/test_host_emul.py proves the arithmetic this way);
//   ObSoloW  one host thread that evaluates every reduction / scan in the WARP's order (32 strided partial sums + xor
//            butterfly, chunk-per-lane scans combined by a Hillis-Steele scan): the packets the GPU must produce, so a GPU
//            mismatch against this emulation is a synchronisation bug and nothing else.
// The element-wise loops (`for (j = g.lane; j < n; j += g.n)`) need no emulation: with one lane they visit every j.
#pragma once
#include "dec_symbols.cuh"   // OB_DEV, ob_imin
#include "ob_group.cuh"

// warp-order host emulation group (see above); element-wise code sees one lane
struct ObSoloW {
    static constexpr int lane = 0;
    static constexpr int n = 1;
    static constexpr int order = 32;
    OB_SOLO_FN void sync() const {}
    OB_SOLO_FN void pace(int) const {}
    OB_SOLO_FN void set_base(int) const {}
};

#ifdef __CUDACC__
// A warp of the encoder kernel that PACES itself against the other warps of its thread block.  Every warp codes its own stream, but
// the encoder is ~45 k instructions of warp-uniform code: when the warps of an SM wander through it independently, every one of them
// misses the instruction cache on its own (measured: 25 of 33 stall cycles per issued instruction were `no_instruction`).  pace(stage)
// publishes how far this warp has come (frame-major, stage-minor) and waits until every other warp of the block is at least as far, so
// that all of them fetch the same few thousand instructions at the same time.  It is a performance hint only: no data depends on it,
// a warp that skips stages simply publishes a later stage, and the warp that is furthest behind never waits.
struct ObWarpPaced : ObWarp {
    volatile int *all;          // [nw] progress of the block's warps, in shared memory
    int nw, w;                  // warps in the block, this warp
    int level, slack;           // pace points of a finer level than `level` are skipped (1: frame stages, 2: + every band, 3: + every leaf); a warp may be `slack` ids ahead
    mutable int base;           // frame-major base of the stage ids (set_base before every frame)
    __device__ __forceinline__ ObWarpPaced(volatile int *a, int n_warps, int lvl, int slk) : ObWarp(), all(a), nw(n_warps), w((int)(threadIdx.x >> 5)), level(lvl), slack(slk), base(0) {}
    __device__ __forceinline__ void set_base(int b) const { base = b; }
    __device__ __forceinline__ void publish(int id) const { __syncwarp(); if (lane == 0) all[w] = id; __syncwarp(); }
    // stage ids: 1..63 frame stages; 64 + 64*band: start of a band; 64 + 64*band + k: k-th leaf of the band
    __device__ __forceinline__ void pace(int stage) const
    {
        const int lvl = stage < 64 ? 1 : ((stage & 63) == 0 ? 2 : 3);
        if (nw <= 1 || lvl > level) return;
        const int id = base + stage;
        publish(id);
        unsigned ns = 200;
        for (;;) {
            int v = lane < nw ? all[lane] : 0x7fffffff;
            v = __reduce_min_sync(0xffffffffu, v);
            if (v >= id - slack) break;
            __nanosleep(ns);                                        // back off: a waiting warp must not eat the issue slots of the ones it waits for
            if (ns < 3200) ns *= 2;
        }
    }
};
#endif

template <class G> struct ObOrder { static constexpr int value = 32; };
template <> struct ObOrder<ObSolo> { static constexpr int value = 1; };

#ifdef __CUDACC__
#define OB_COOP __device__ __forceinline__
// lane-strided loops of the warp-per-stream encoder run 1-6 iterations (a band is <= 176 coefficients): unrolled four times they only cost
// instruction-cache footprint, which is what bounds that kernel
#define OB_ROLLED _Pragma("unroll 1")
#define OB_ROLLED_G _Pragma("unroll 1")
#else
#define OB_COOP static inline
#define OB_ROLLED
#define OB_ROLLED_G
#endif

#ifdef __CUDACC__
// ---------------------------------------------------------------- device: ObWarp ------------------------------------------------
template <class F> OB_COOP float ob_psum(const ObWarp &g, int n, float init, F f)
{
    float s = 0.f;
    OB_ROLLED
    for (int j = g.lane; j < n; j += 32) s = s + f(j);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = s + __shfl_xor_sync(0xffffffffu, s, o);
    return init + s;
}
template <class F> OB_COOP void ob_psum2(const ObWarp &g, int n, float &a, float &b, F f)       // f(j, a, b) accumulates into a and b
{
    float sa = 0.f, sb = 0.f;
    OB_ROLLED
    for (int j = g.lane; j < n; j += 32) f(j, sa, sb);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sa = sa + __shfl_xor_sync(0xffffffffu, sa, o); sb = sb + __shfl_xor_sync(0xffffffffu, sb, o); }
    a = a + sa; b = b + sb;
}
template <class F> OB_COOP float ob_pmax(const ObWarp &g, int n, float init, F f)
{
    float s = init;
    OB_ROLLED
    for (int j = g.lane; j < n; j += 32) { const float v = f(j); s = v > s ? v : s; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const float v = __shfl_xor_sync(0xffffffffu, s, o); s = v > s ? v : s; }
    return s;
}
template <class F> OB_COOP uint32_t ob_psum_u32(const ObWarp &g, int n, F f)
{
    uint32_t s = 0;
    OB_ROLLED
    for (int j = g.lane; j < n; j += 32) s += f(j);
    return __reduce_add_sync(0xffffffffu, s);
}
template <class F> OB_COOP uint32_t ob_por_u32(const ObWarp &g, int n, F f)
{
    uint32_t s = 0;
    OB_ROLLED
    for (int j = g.lane; j < n; j += 32) s |= f(j);
    return __reduce_or_sync(0xffffffffu, s);
}
#endif

// ---------------------------------------------------------------- host: ObSolo (reference order) -------------------------------
template <class F> OB_COOP float ob_psum(const ObSolo &, int n, float init, F f) { float s = init; for (int j = 0; j < n; j++) s = s + f(j); return s; }
template <class F> OB_COOP void ob_psum2(const ObSolo &, int n, float &a, float &b, F f) { for (int j = 0; j < n; j++) f(j, a, b); }
template <class F> OB_COOP float ob_pmax(const ObSolo &, int n, float init, F f) { float s = init; for (int j = 0; j < n; j++) { const float v = f(j); s = v > s ? v : s; } return s; }
template <class F> OB_COOP uint32_t ob_psum_u32(const ObSolo &, int n, F f) { uint32_t s = 0; for (int j = 0; j < n; j++) s += f(j); return s; }
template <class F> OB_COOP uint32_t ob_por_u32(const ObSolo &, int n, F f) { uint32_t s = 0; for (int j = 0; j < n; j++) s |= f(j); return s; }

#ifndef __CUDACC__
// ---------------------------------------------------------------- host: ObSoloW (warp order) -----------------------------------
static inline float ob_tree32(float *p)
{
    for (int o = 16; o > 0; o >>= 1) { float q[32]; for (int l = 0; l < 32; l++) q[l] = p[l] + p[l ^ o]; for (int l = 0; l < 32; l++) p[l] = q[l]; }
    return p[0];
}
template <class F> OB_COOP float ob_psum(const ObSoloW &, int n, float init, F f)
{
    float p[32];
    for (int l = 0; l < 32; l++) p[l] = 0.f;
    for (int j = 0; j < n; j++) p[j & 31] = p[j & 31] + f(j);
    return init + ob_tree32(p);
}
template <class F> OB_COOP void ob_psum2(const ObSoloW &, int n, float &a, float &b, F f)
{
    float pa[32], pb[32];
    for (int l = 0; l < 32; l++) pa[l] = pb[l] = 0.f;
    for (int j = 0; j < n; j++) f(j, pa[j & 31], pb[j & 31]);
    a = a + ob_tree32(pa); b = b + ob_tree32(pb);
}
template <class F> OB_COOP float ob_pmax(const ObSoloW &, int n, float init, F f) { return ob_pmax(ObSolo(), n, init, f); }
template <class F> OB_COOP uint32_t ob_psum_u32(const ObSoloW &, int n, F f) { return ob_psum_u32(ObSolo(), n, f); }
template <class F> OB_COOP uint32_t ob_por_u32(const ObSoloW &, int n, F f) { return ob_por_u32(ObSolo(), n, f); }
#endif

// ---- first-order recurrence  y_i = u(i) + a * y_{i-1}  (i = 0 .. n-1, y_{-1} = y0), put(i, y_i); returns y_{n-1} ----------------
// rev: run over i = n-1 .. 0 instead (the recurrence index still counts from the first element visited).
// Warp form: lane l owns the contiguous chunk [l*L, (l+1)*L); pass 1 gives each chunk's zero-state response and a^len, a Hillis-Steele
// scan of the affine maps gives every chunk's entry state, pass 2 re-runs the chunk from it.  u(i) is evaluated twice and must not
// read what put() of ANOTHER chunk writes.
template <class U, class P> OB_COOP float ob_scan1(const ObSolo &, int n, float a, float y0, bool rev, U u, P put)
{
    float y = y0;
    for (int t = 0; t < n; t++) { const int i = rev ? n - 1 - t : t; y = u(i) + a * y; put(i, y); }
    return y;
}
#ifdef __CUDACC__
template <class U, class P> OB_COOP float ob_scan1(const ObWarp &g, int n, float a, float y0, bool rev, U u, P put)
{
    const int L = (n + 31) >> 5, lo = ob_imin(n, g.lane * L), hi = ob_imin(n, lo + L);
    float e = 0.f, A = 1.f;
    for (int t = lo; t < hi; t++) { const int i = rev ? n - 1 - t : t; e = u(i) + a * e; A = A * a; }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float Ap = __shfl_up_sync(0xffffffffu, A, o), ep = __shfl_up_sync(0xffffffffu, e, o);
        if (g.lane >= o) { e = A * ep + e; A = A * Ap; }
    }
    const float yend = A * y0 + e;                                   // state after this lane's chunk
    float y = __shfl_up_sync(0xffffffffu, yend, 1);
    if (g.lane == 0) y = y0;
    for (int t = lo; t < hi; t++) { const int i = rev ? n - 1 - t : t; y = u(i) + a * y; put(i, y); }
    __syncwarp();
    return __shfl_sync(0xffffffffu, yend, 31);
}
#else
template <class U, class P> OB_COOP float ob_scan1(const ObSoloW &, int n, float a, float y0, bool rev, U u, P put)
{
    const int L = (n + 31) >> 5;
    float e[32], A[32];
    for (int l = 0; l < 32; l++) {
        const int lo = ob_imin(n, l * L), hi = ob_imin(n, lo + L);
        e[l] = 0.f; A[l] = 1.f;
        for (int t = lo; t < hi; t++) { const int i = rev ? n - 1 - t : t; e[l] = u(i) + a * e[l]; A[l] = A[l] * a; }
    }
    for (int o = 1; o < 32; o <<= 1) {
        float e2[32], A2[32];
        for (int l = 0; l < 32; l++) { if (l >= o) { e2[l] = A[l] * e[l - o] + e[l]; A2[l] = A[l] * A[l - o]; } else { e2[l] = e[l]; A2[l] = A[l]; } }
        for (int l = 0; l < 32; l++) { e[l] = e2[l]; A[l] = A2[l]; }
    }
    for (int l = 0; l < 32; l++) {
        const int lo = ob_imin(n, l * L), hi = ob_imin(n, lo + L);
        float y = l == 0 ? y0 : A[l - 1] * y0 + e[l - 1];
        for (int t = lo; t < hi; t++) { const int i = rev ? n - 1 - t : t; y = u(i) + a * y; put(i, y); }
    }
    return A[31] * y0 + e[31];
}
#endif

// ---- second-order recurrence: state s = (s0, s1), step(i, s0, s1) advances the state by one sample (and emits its output); the
// homogeneous part of step is s -> M s with the constant matrix M = [[m00, m01], [m10, m11]].  Same chunk-per-lane scheme. -----------
struct ObState2 { float s0, s1; };
template <class S> OB_COOP ObState2 ob_scan2(const ObSolo &, int n, float, float, float, float, ObState2 st, S step)
{
    for (int i = 0; i < n; i++) step(i, st.s0, st.s1, true);
    return st;
}
#ifdef __CUDACC__
template <class S> OB_COOP ObState2 ob_scan2(const ObWarp &g, int n, float m00, float m01, float m10, float m11, ObState2 st, S step)
{
    const int L = (n + 31) >> 5, lo = ob_imin(n, g.lane * L), hi = ob_imin(n, lo + L);
    float e0 = 0.f, e1 = 0.f, a00 = 1.f, a01 = 0.f, a10 = 0.f, a11 = 1.f;         // zero-state response and M^len of this chunk
    for (int i = lo; i < hi; i++) {
        step(i, e0, e1, false);
        const float b00 = m00 * a00 + m01 * a10, b01 = m00 * a01 + m01 * a11, b10 = m10 * a00 + m11 * a10, b11 = m10 * a01 + m11 * a11;
        a00 = b00; a01 = b01; a10 = b10; a11 = b11;
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float p00 = __shfl_up_sync(0xffffffffu, a00, o), p01 = __shfl_up_sync(0xffffffffu, a01, o), p10 = __shfl_up_sync(0xffffffffu, a10, o),
                    p11 = __shfl_up_sync(0xffffffffu, a11, o), q0 = __shfl_up_sync(0xffffffffu, e0, o), q1 = __shfl_up_sync(0xffffffffu, e1, o);
        if (g.lane >= o) {
            const float n0 = a00 * q0 + a01 * q1 + e0, n1 = a10 * q0 + a11 * q1 + e1;
            const float b00 = a00 * p00 + a01 * p10, b01 = a00 * p01 + a01 * p11, b10 = a10 * p00 + a11 * p10, b11 = a10 * p01 + a11 * p11;
            e0 = n0; e1 = n1; a00 = b00; a01 = b01; a10 = b10; a11 = b11;
        }
    }
    const float y0 = a00 * st.s0 + a01 * st.s1 + e0, y1 = a10 * st.s0 + a11 * st.s1 + e1;     // state after this lane's chunk
    float s0 = __shfl_up_sync(0xffffffffu, y0, 1), s1 = __shfl_up_sync(0xffffffffu, y1, 1);
    if (g.lane == 0) { s0 = st.s0; s1 = st.s1; }
    for (int i = lo; i < hi; i++) step(i, s0, s1, true);
    __syncwarp();
    ObState2 r;
    r.s0 = __shfl_sync(0xffffffffu, y0, 31); r.s1 = __shfl_sync(0xffffffffu, y1, 31);
    return r;
}
#else
template <class S> OB_COOP ObState2 ob_scan2(const ObSoloW &, int n, float m00, float m01, float m10, float m11, ObState2 st, S step)
{
    const int L = (n + 31) >> 5;
    float e0[32], e1[32], a00[32], a01[32], a10[32], a11[32];
    for (int l = 0; l < 32; l++) {
        const int lo = ob_imin(n, l * L), hi = ob_imin(n, lo + L);
        e0[l] = e1[l] = 0.f; a00[l] = a11[l] = 1.f; a01[l] = a10[l] = 0.f;
        for (int i = lo; i < hi; i++) {
            step(i, e0[l], e1[l], false);
            const float b00 = m00 * a00[l] + m01 * a10[l], b01 = m00 * a01[l] + m01 * a11[l], b10 = m10 * a00[l] + m11 * a10[l], b11 = m10 * a01[l] + m11 * a11[l];
            a00[l] = b00; a01[l] = b01; a10[l] = b10; a11[l] = b11;
        }
    }
    for (int o = 1; o < 32; o <<= 1) {
        float E0[32], E1[32], B00[32], B01[32], B10[32], B11[32];
        for (int l = 0; l < 32; l++) {
            if (l >= o) {
                const int p = l - o;
                E0[l] = a00[l] * e0[p] + a01[l] * e1[p] + e0[l]; E1[l] = a10[l] * e0[p] + a11[l] * e1[p] + e1[l];
                B00[l] = a00[l] * a00[p] + a01[l] * a10[p]; B01[l] = a00[l] * a01[p] + a01[l] * a11[p];
                B10[l] = a10[l] * a00[p] + a11[l] * a10[p]; B11[l] = a10[l] * a01[p] + a11[l] * a11[p];
            } else { E0[l] = e0[l]; E1[l] = e1[l]; B00[l] = a00[l]; B01[l] = a01[l]; B10[l] = a10[l]; B11[l] = a11[l]; }
        }
        for (int l = 0; l < 32; l++) { e0[l] = E0[l]; e1[l] = E1[l]; a00[l] = B00[l]; a01[l] = B01[l]; a10[l] = B10[l]; a11[l] = B11[l]; }
    }
    for (int l = 0; l < 32; l++) {
        const int lo = ob_imin(n, l * L), hi = ob_imin(n, lo + L);
        float s0, s1;
        if (l == 0) { s0 = st.s0; s1 = st.s1; }
        else { s0 = a00[l - 1] * st.s0 + a01[l - 1] * st.s1 + e0[l - 1]; s1 = a10[l - 1] * st.s0 + a11[l - 1] * st.s1 + e1[l - 1]; }
        for (int i = lo; i < hi; i++) step(i, s0, s1, true);
    }
    ObState2 r;
    r.s0 = a00[31] * st.s0 + a01[31] * st.s1 + e0[31]; r.s1 = a10[31] * st.s0 + a11[31] * st.s1 + e1[31];
    return r;
}
#endif

// ---- first-order recurrence given as an exact per-sample step: step(i, y, emit) advances the state y over sample i with the reference's own
// arithmetic (and emits what it has to when emit is set); the homogeneous part of a step is y -> a * y.  Same chunk-per-lane scheme as ob_scan1. ----
template <class S> OB_COOP float ob_scan1s(const ObSolo &, int n, float, float y0, S step)
{
    float y = y0;
    for (int i = 0; i < n; i++) step(i, y, true);
    return y;
}
#ifdef __CUDACC__
template <class S> OB_COOP float ob_scan1s(const ObWarp &g, int n, float a, float y0, S step)
{
    const int L = (n + 31) >> 5, lo = ob_imin(n, g.lane * L), hi = ob_imin(n, lo + L);
    float e = 0.f, A = 1.f;
    for (int i = lo; i < hi; i++) { step(i, e, false); A = A * a; }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float Ap = __shfl_up_sync(0xffffffffu, A, o), ep = __shfl_up_sync(0xffffffffu, e, o);
        if (g.lane >= o) { e = A * ep + e; A = A * Ap; }
    }
    const float yend = A * y0 + e;
    float y = __shfl_up_sync(0xffffffffu, yend, 1);
    if (g.lane == 0) y = y0;
    for (int i = lo; i < hi; i++) step(i, y, true);
    __syncwarp();
    return __shfl_sync(0xffffffffu, yend, 31);
}
#else
template <class S> OB_COOP float ob_scan1s(const ObSoloW &, int n, float a, float y0, S step)
{
    const int L = (n + 31) >> 5;
    float e[32], A[32];
    for (int l = 0; l < 32; l++) {
        const int lo = ob_imin(n, l * L), hi = ob_imin(n, lo + L);
        e[l] = 0.f; A[l] = 1.f;
        for (int i = lo; i < hi; i++) { step(i, e[l], false); A[l] = A[l] * a; }
    }
    for (int o = 1; o < 32; o <<= 1) {
        float e2[32], A2[32];
        for (int l = 0; l < 32; l++) { if (l >= o) { e2[l] = A[l] * e[l - o] + e[l]; A2[l] = A[l] * A[l - o]; } else { e2[l] = e[l]; A2[l] = A[l]; } }
        for (int l = 0; l < 32; l++) { e[l] = e2[l]; A[l] = A2[l]; }
    }
    for (int l = 0; l < 32; l++) {
        const int lo = ob_imin(n, l * L), hi = ob_imin(n, lo + L);
        float y = l == 0 ? y0 : A[l - 1] * y0 + e[l - 1];
        for (int i = lo; i < hi; i++) step(i, y, true);
    }
    return A[31] * y0 + e[31];
}
#endif

// ---- plain inclusive prefix sum of f(i) (float), out(i, sum_{t<=i} f(t)); warp form: chunk per lane + shuffle scan of the totals ----
template <class F, class P> OB_COOP void ob_prefix_sum(const ObSolo &, int n, float y0, F f, P out)
{
    float y = y0;
    for (int i = 0; i < n; i++) { y = y + f(i); out(i, y); }
}
#ifdef __CUDACC__
template <class F, class P> OB_COOP void ob_prefix_sum(const ObWarp &g, int n, float y0, F f, P out)
{
    const int L = (n + 31) >> 5, lo = ob_imin(n, g.lane * L), hi = ob_imin(n, lo + L);
    float e = 0.f;
    for (int i = lo; i < hi; i++) e = e + f(i);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float p = __shfl_up_sync(0xffffffffu, e, o); if (g.lane >= o) e = p + e; }
    float y = __shfl_up_sync(0xffffffffu, e, 1);
    y = g.lane == 0 ? y0 : y0 + y;
    for (int i = lo; i < hi; i++) { y = y + f(i); out(i, y); }
    __syncwarp();
}
#else
template <class F, class P> OB_COOP void ob_prefix_sum(const ObSoloW &, int n, float y0, F f, P out)
{
    const int L = (n + 31) >> 5;
    float e[32];
    for (int l = 0; l < 32; l++) { const int lo = ob_imin(n, l * L), hi = ob_imin(n, lo + L); e[l] = 0.f; for (int i = lo; i < hi; i++) e[l] = e[l] + f(i); }
    for (int o = 1; o < 32; o <<= 1) { float q[32]; for (int l = 0; l < 32; l++) q[l] = l >= o ? e[l - o] + e[l] : e[l]; for (int l = 0; l < 32; l++) e[l] = q[l]; }
    for (int l = 0; l < 32; l++) {
        const int lo = ob_imin(n, l * L), hi = ob_imin(n, lo + L);
        float y = l == 0 ? y0 : y0 + e[l - 1];
        for (int i = lo; i < hi; i++) { y = y + f(i); out(i, y); }
    }
}
#endif

// ---- warp-uniform broadcast of a value one lane holds (host: identity) ----
#ifdef __CUDACC__
OB_COOP int ob_bcast(const ObWarp &, int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
OB_COOP float ob_bcastf(const ObWarp &, float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
#endif
OB_COOP int ob_bcast(const ObSolo &, int v, int) { return v; }
OB_COOP float ob_bcastf(const ObSolo &, float v, int) { return v; }
#ifndef __CUDACC__
OB_COOP int ob_bcast(const ObSoloW &, int v, int) { return v; }
OB_COOP float ob_bcastf(const ObSoloW &, float v, int) { return v; }
#endif
