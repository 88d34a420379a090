// noise.cuh — "dllm_noise v1": counter-based standard-normal generator shared by the seeded sampling loop.
//
// The reference draws its noise from an unseeded thread_rng (diffuse-llm-rs/src/lib.rs:875-878, :1107-1109, :1201), so its
// stream cannot be reproduced; SURVEY.md §8(d) asks for a counter-based generator that the CPU oracle reproduces bit for
// bit.  Element i of stream s under `seed` depends on (seed, s, i) only:
//     key = splitmix64(splitmix64(seed) + s);   r = splitmix64(key + (i >> 1))          (one draw per PAIR of elements)
//     a = r >> 41 (23 bits), b = (r >> 17) & 0xFFFFFF (24 bits)
//     u = (2a+1) * 2^-24 in (0,1);   radius = sqrt(-2 ln u);   theta = (2b+1) * pi / 2^24
//     element 2p = radius * cos(theta), element 2p+1 = radius * sin(theta)               (Box-Muller)
// Every step is either exact (integer work, int -> float of < 2^24, powers of two) or ONE correctly rounded IEEE f32
// operation in a fixed order (mul, add, fma, div, sqrt): ln and sin / cos are polynomials evaluated with fmaf, the angle is
// reduced in the integer domain.  No libm / SFU function is involved, so gcc's and nvcc's results are identical.
#pragma once
#include <stdint.h>

__host__ __device__ __forceinline__ uint64_t dn_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t dn_key(uint64_t seed, uint64_t stream) {
    return dn_splitmix64(dn_splitmix64(seed) + stream);
}

#ifdef __CUDACC__
// polynomial sin / cos on (0, pi/4]
__device__ __forceinline__ void dn_sincos_poly(float phi, float *s, float *c) {
    const float p2 = __fmul_rn(phi, phi);
    float ps = __fmaf_rn(p2, 2.75573192e-6f, -1.98412698e-4f);      //  1/9!, -1/7!
    ps = __fmaf_rn(p2, ps, 8.33333333e-3f);                        //  1/5!
    ps = __fmaf_rn(p2, ps, -1.66666667e-1f);                       // -1/3!
    ps = __fmaf_rn(p2, ps, 1.0f);
    *s = __fmul_rn(phi, ps);
    float pc = __fmaf_rn(p2, -2.75573192e-7f, 2.48015873e-5f);      // -1/10!, 1/8!
    pc = __fmaf_rn(p2, pc, -1.38888889e-3f);                       // -1/6!
    pc = __fmaf_rn(p2, pc, 4.16666667e-2f);                        //  1/4!
    pc = __fmaf_rn(p2, pc, -0.5f);
    *c = __fmaf_rn(p2, pc, 1.0f);
}

// the two normals of pair p
__device__ __forceinline__ void dn_normal_pair(uint64_t key, uint64_t p, float *z0, float *z1) {
    const uint64_t r = dn_splitmix64(key + p);
    const uint32_t a = (uint32_t)(r >> 41), b = (uint32_t)(r >> 17) & 0xFFFFFFu;
    // ---- radius = sqrt(-2 ln u), u = v * 2^-24, v = 2a+1 ----
    const uint32_t v = 2u * a + 1u;
    int e = 31 - __clz((int)v);                                   // msb of v: 0..23
    float m = __fmul_rn((float)v, __uint_as_float((uint32_t)(127 - e) << 23));      // v * 2^-e in [1, 2): exact
    if (m > 1.41421356f) { m = __fmul_rn(m, 0.5f); e += 1; }      // m in (0.707, 1.4143]
    const float t = __fdiv_rn(__fsub_rn(m, 1.0f), __fadd_rn(m, 1.0f));
    const float t2 = __fmul_rn(t, t);
    float pl = __fmaf_rn(t2, 1.11111111e-1f, 1.42857143e-1f);       // 1/9, 1/7
    pl = __fmaf_rn(t2, pl, 0.2f);
    pl = __fmaf_rn(t2, pl, 3.33333343e-1f);
    pl = __fmaf_rn(t2, pl, 1.0f);
    const float ln_m = __fmul_rn(__fmul_rn(2.0f, t), pl);         // ln m = 2 atanh(t)
    const float ln_u = __fmaf_rn((float)(e - 24), 6.93147181e-1f, ln_m);
    const float radius = __fsqrt_rn(__fmul_rn(-2.0f, ln_u));
    // ---- angle theta = (2b+1) * pi / 2^24: quadrant = b >> 22, reflected about pi/4 in the integer domain ----
    const uint32_t quad = b >> 22, f = b & 0x3FFFFFu;
    const bool refl = f >= 0x200000u;
    const uint32_t g = refl ? 0x3FFFFFu - f : f;                  // < 2^21
    const float phi = __fmul_rn((float)(2u * g + 1u), 1.87253514e-7f);   // (2g+1) * (pi/2) / 2^23, in (0, pi/4)
    float sp, cp;
    dn_sincos_poly(phi, &sp, &cp);
    const float sq = refl ? cp : sp, cq = refl ? sp : cp;         // sin / cos of the angle inside the quadrant
    float cs, sn;
    switch (quad) {
        case 0: cs = cq; sn = sq; break;
        case 1: cs = -sq; sn = cq; break;
        case 2: cs = -cq; sn = -sq; break;
        default: cs = sq; sn = -cq; break;
    }
    *z0 = __fmul_rn(radius, cs);
    *z1 = __fmul_rn(radius, sn);
}
__device__ __forceinline__ float dn_normal(uint64_t key, uint64_t i) {
    float z0, z1;
    dn_normal_pair(key, i >> 1, &z0, &z1);
    return (i & 1) ? z1 : z0;
}
#endif
