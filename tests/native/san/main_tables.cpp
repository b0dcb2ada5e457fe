#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include "../../../include/pbrtgpu.h"
extern "C" int toh_ld_samples(const PbrtGpuRenderDesc *rd, int taskNum, int x0, int y0, int x1, int y1, int general, int rounds, float *out);
int main() {
    const int spps[] = {1, 2, 16, 256, 2048};
    for (int spp : spps)
        for (int general = 0; general < 2; ++general) {
            PbrtGpuRenderDesc rd;
            memset(&rd, 0, sizeof(rd));
            rd.xResolution = rd.yResolution = 32; rd.samplesPerPixel = spp; rd.nTasks = 4; rd.taskEnd = 4;
            rd.cropWindow[1] = rd.cropWindow[3] = 1.f; rd.shutterClose = 1.f;
            rd.n1D = 14; rd.n2D = 9;
            for (int i = 0; i < 14; ++i) rd.size1D[i] = 1;
            for (int i = 0; i < 9; ++i) rd.size2D[i] = 1;
            if (general && spp <= 256) { rd.integrator = 1; rd.n1D = 3; rd.n2D = 2; rd.size1D[1] = 7; rd.size2D[0] = 5; rd.size2D[1] = 64; }
            int nf = 5;
            for (int j = 0; j < rd.n1D; ++j) nf += rd.size1D[j];
            for (int j = 0; j < rd.n2D; ++j) nf += 2 * rd.size2D[j];
            std::vector<float> out((size_t)spp * nf);
            if (toh_ld_samples(&rd, 1, 0, 8, 16, 16, general, spp <= 16 ? 3 : 1, out.data())) { printf("rc\n"); return 1; }
        }
    printf("ok\n");
    return 0;
}
