#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../../include/pbrtgpu.h"
extern "C" int doh_bvh_build(const float *bounds, uint32_t n, uint32_t maxPrimsInNode, PbrtGpuBVHNode *nodes, uint32_t *nNodes,
                             uint32_t *order, uint32_t *nLevels, unsigned maxBlocks, int reverse);
int main() {
    srand(1);
    for (int it = 0; it < 600; ++it) {
        uint32_t n = 1 + rand() % (it < 500 ? 300 : 40000);
        int kind = rand() % 4;
        std::vector<float> b(6 * (size_t)n);
        for (uint32_t i = 0; i < n; ++i) {
            float c[3], e[3];
            for (int k = 0; k < 3; ++k) {
                c[k] = kind == 0 ? rand() / (float)RAND_MAX : kind == 1 ? (float)(rand() % 5) : kind == 2 ? (float)(rand() % 2) : (k == 0 ? (float)i : 0.f);
                e[k] = kind == 3 ? .25f : (rand() % 3) * .5f;
            }
            for (int k = 0; k < 3; ++k) { b[6 * i + k] = c[k] - e[k]; b[6 * i + 3 + k] = c[k] + e[k]; }
        }
        std::vector<PbrtGpuBVHNode> nodes(2 * (size_t)n);
        std::vector<uint32_t> order(n);
        uint32_t nn = 0, nl = 0;
        const uint32_t maxp[] = {1, 2, 4, 255};
        if (doh_bvh_build(b.data(), n, maxp[rand() % 4], nodes.data(), &nn, order.data(), &nl, rand() % 3, rand() % 2)) return 1;
        if (nn == 0 || nn > 2 * n - 1) { printf("bad node count\n"); return 2; }
    }
    printf("ok\n");
    return 0;
}
