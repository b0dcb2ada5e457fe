// refdump.cpp -- TEST INFRASTRUCTURE.  OUR program, linked against the UNMODIFIED
// reference library (oracle/_ref/lib/libpbrt.a, built by oracle/Makefile from
// /root/reference/src), that prints known-answer vectors by calling the
// reference's own public classes.  tests/golden/make_golden.py turns its output
// into the committed fixtures.
//
//   refdump ld   <x0> <x1> <y0> <y1> <spp> <xres> <yres> <nTasks> <taskNum> <shutterOpen> <shutterClose>
//        first pixel of sub-sampler `taskNum`: spp rows of 37 floats
//        (imageX imageY lensU lensV time, 14 one-D, 9 two-D pairs) + the RNG draw that follows
//   refdump rng  <seed> <n>        n RandomUInt() values of RNG(seed)
#include "pbrt.h"
#include "rng.h"
#include "sampler.h"
#include "samplers/lowdiscrepancy.h"
#include "integrators/path.h"
#include "integrators/emission.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

int main(int argc, char **argv) {
    if (argc >= 4 && !strcmp(argv[1], "rng")) {
        RNG rng((uint32_t)strtoul(argv[2], NULL, 10));
        int n = atoi(argv[3]);
        for (int i = 0; i < n; ++i) printf("%u\n", rng.RandomUInt());
        return 0;
    }
    if (argc >= 13 && !strcmp(argv[1], "ld")) {
        int x0 = atoi(argv[2]), x1 = atoi(argv[3]), y0 = atoi(argv[4]), y1 = atoi(argv[5]);
        int spp = atoi(argv[6]);
        int nTasks = atoi(argv[9]), taskNum = atoi(argv[10]);
        float so = (float)atof(argv[11]), sc = (float)atof(argv[12]);
        LDSampler mainSampler(x0, x1, y0, y1, spp, so, sc);
        PathIntegrator surf(5);
        EmissionIntegrator vol(1.f);
        Sample *origSample = new Sample(&mainSampler, &surf, &vol, NULL);
        Sampler *sub = mainSampler.GetSubSampler(taskNum, nTasks);
        if (!sub) { printf("EMPTY\n"); return 0; }
        RNG rng(taskNum);
        int maxSamples = sub->MaximumSampleCount();
        Sample *samples = origSample->Duplicate(maxSamples);
        int n = sub->GetMoreSamples(samples, rng);
        printf("%d %d %d\n", n, (int)origSample->n1D.size(), (int)origSample->n2D.size());
        for (int i = 0; i < n; ++i) {
            printf("%.9g %.9g %.9g %.9g %.9g", samples[i].imageX, samples[i].imageY, samples[i].lensU,
                   samples[i].lensV, samples[i].time);
            for (size_t j = 0; j < samples[i].n1D.size(); ++j) printf(" %.9g", samples[i].oneD[j][0]);
            for (size_t j = 0; j < samples[i].n2D.size(); ++j)
                printf(" %.9g %.9g", samples[i].twoD[j][0], samples[i].twoD[j][1]);
            printf("\n");
        }
        printf("%u\n", rng.RandomUInt());
        return 0;
    }
    fprintf(stderr, "usage: refdump ld ... | refdump rng seed n\n");
    return 1;
}
