// refmath_check.cpp -- compares pbrt_v2_b200/csrc/device/refmath.h (compiled as plain C++) with the host libm
// (glibc) over the argument ranges pbrt's hot path uses.  Prints "name total mismatches".
//   unary functions: EXHAUSTIVE over every float in the range;  binary: N random pairs.
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <thread>
#include <vector>
#include <atomic>
#include <random>
#include "../../pbrt_v2_b200/csrc/device/refmath.h"

static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

template <typename F, typename G>
static void exhaustive(const char *name, F ours, G libm, float lo, float hi, int nthreads, uint32_t step) {
    // iterates bit patterns of [lo, hi]; lo, hi >= 0 or both <= 0
    uint32_t a = f2u(lo), b = f2u(hi);
    if (a > b) std::swap(a, b);
    std::atomic<uint64_t> mism(0), total(0);
    std::atomic<uint32_t> firstBad(0);
    std::vector<std::thread> th;
    for (int k = 0; k < nthreads; ++k)
        th.emplace_back([&, k] {
            uint64_t m = 0, n = 0;
            for (uint64_t u = (uint64_t)a + (uint64_t)k * step; u <= b; u += (uint64_t)nthreads * step) {
                float x = u2f((uint32_t)u);
                if (f2u(ours(x)) != f2u(libm(x))) { if (!m) firstBad = (uint32_t)u; ++m; }
                ++n;
            }
            mism += m; total += n;
        });
    for (auto &t : th) t.join();
    printf("%s %llu %llu", name, (unsigned long long)total.load(), (unsigned long long)mism.load());
    if (mism.load()) { float x = u2f(firstBad.load()); printf("  first x=%a ours=%a libm=%a", x, ours(x), libm(x)); }
    printf("\n");
}

int main(int argc, char **argv) {
    int nthreads = argc > 1 ? atoi(argv[1]) : 8;
    uint32_t step = argc > 2 ? (uint32_t)atoi(argv[2]) : 1;   // 1 = exhaustive
    uint64_t npairs = argc > 3 ? strtoull(argv[3], nullptr, 10) : 200000000ull;
    exhaustive("sinf[0,7]", refmath::ref_sinf, sinf, 0.f, 7.f, nthreads, step);
    exhaustive("cosf[0,7]", refmath::ref_cosf, cosf, 0.f, 7.f, nthreads, step);
    exhaustive("sinf[-7,0]", refmath::ref_sinf, sinf, -0.f, -7.f, nthreads, step);
    exhaustive("expf[-210,0]", refmath::ref_expf, expf, -0.f, -210.f, nthreads, step);
    exhaustive("expf[0,90]", refmath::ref_expf, expf, 0.f, 90.f, nthreads, step);
    exhaustive("acosf[0,1]", refmath::ref_acosf, acosf, 0.f, 1.f, nthreads, step);
    exhaustive("acosf[-1,0]", refmath::ref_acosf, acosf, -0.f, -1.f, nthreads, step);
    exhaustive("atanf[0,1e30]", refmath::ref_atanf, atanf, 0.f, 1e30f, nthreads, step);
    {
        std::atomic<uint64_t> m1(0), m2(0), m3(0);
        std::vector<std::thread> th;
        for (int k = 0; k < nthreads; ++k)
            th.emplace_back([&, k] {
                std::mt19937_64 g(1234 + k);
                std::uniform_real_distribution<float> U(-1, 1), U01(0, 1), E(0, 10001);
                uint64_t a = 0, b = 0, c = 0;
                for (uint64_t i = 0; i < npairs / nthreads; ++i) {
                    float y = U(g), x = U(g);
                    if (i & 1) { y *= 1e-3f; }
                    if (f2u(atan2f(y, x)) != f2u(refmath::ref_atan2f(y, x))) ++a;
                    float bs = U01(g), e = E(g);
                    if (i % 3 == 0) e = 1.f / (e + 1.f);
                    if (f2u(powf(bs, e)) != f2u(refmath::ref_powf(bs, e))) ++b;
                    float bs2 = 1.f - bs * bs * 1e-2f;   // near 1 (Blinn cos theta_h close to 1)
                    if (f2u(powf(bs2, e)) != f2u(refmath::ref_powf(bs2, e))) ++c;
                }
                m1 += a; m2 += b; m3 += c;
            });
        for (auto &t : th) t.join();
        printf("atan2f %llu %llu\n", (unsigned long long)npairs, (unsigned long long)m1.load());
        printf("powf %llu %llu\n", (unsigned long long)npairs, (unsigned long long)m2.load());
        printf("powf_near1 %llu %llu\n", (unsigned long long)npairs, (unsigned long long)m3.load());
    }
    // spot values
    float sx[] = {0.f, -0.f, 1.f, 0.5f, 1e-40f};
    for (float x : sx)
        for (float y : {0.5f, 2.f, 1e4f, 1e-4f})
            if (f2u(powf(x, y)) != f2u(refmath::ref_powf(x, y))) printf("powf spot mismatch %a %a\n", x, y);
    return 0;
}
