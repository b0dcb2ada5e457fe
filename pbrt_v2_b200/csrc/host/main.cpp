// main.cpp -- `pbrt_b200 [--ncores n] [--outfile f] [--quiet] scene.pbrt`:
// command-line twin of the reference's src/main/pbrt.cpp:41-82 for scenes whose
// Accelerator / SurfaceIntegrator / Renderer select the GPU hot path.
//
// Several GPUs: start one process per GPU, e.g.
//     for r in 0 1 2 3; do pbrt_b200 --rank $r --world 4 --rendezvous /tmp/pbrt.id scene.pbrt & done; wait
// (or under any launcher that sets RANK / WORLD_SIZE / LOCAL_RANK).  Rank r renders on device LOCAL_RANK (default r), the
// ranks find each other through the rendezvous FILE (rank 0 writes the 128-byte NCCL id, the others wait for it), the
// films are summed onto rank 0 with one ncclReduce inside the library, and rank 0 writes the image.
#include "scene.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

static bool ExchangeId(const std::string &path, int rank, unsigned char id[PBRTGPU_COMM_ID_BYTES]) {
    if (rank == 0) {
        if (pbrtgpu_comm_unique_id(id) != PBRTGPU_OK) return false;
        const std::string tmp = path + ".tmp";
        FILE *fp = fopen(tmp.c_str(), "wb");
        if (!fp) return false;
        fwrite(id, 1, PBRTGPU_COMM_ID_BYTES, fp);
        fclose(fp);
        return rename(tmp.c_str(), path.c_str()) == 0;   // atomic: a reader never sees a partial id
    }
    for (int tries = 0; tries < 6000; ++tries) {   // up to 10 minutes
        FILE *fp = fopen(path.c_str(), "rb");
        if (fp) {
            const size_t n = fread(id, 1, PBRTGPU_COMM_ID_BYTES, fp);
            fclose(fp);
            if (n == PBRTGPU_COMM_ID_BYTES) return true;
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(100));
    }
    return false;
}

int main(int argc, char *argv[]) {
    pb::Options opt;
    std::vector<std::string> files;
    auto envInt = [](const char *k, int d) { const char *e = getenv(k); return e ? atoi(e) : d; };
    int rank = envInt("RANK", 0), world = envInt("WORLD_SIZE", 1), localRank = envInt("LOCAL_RANK", -1);
    std::string rendezvous = getenv("PBRTGPU_RENDEZVOUS") ? getenv("PBRTGPU_RENDEZVOUS") : "";
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--ncores")) opt.nCores = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--rank") && i + 1 < argc) rank = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--world") && i + 1 < argc) world = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) localRank = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--rendezvous") && i + 1 < argc) rendezvous = argv[++i];
        else if (!strcmp(argv[i], "--outfile")) opt.imageFile = argv[++i];
        else if (!strcmp(argv[i], "--quick")) opt.quickRender = true;
        else if (!strcmp(argv[i], "--quiet")) opt.quiet = true;
        else if (!strcmp(argv[i], "--verbose")) {}
        else if (!strcmp(argv[i], "--help") || !strcmp(argv[i], "-h")) {
            printf("usage: pbrt_b200 [--ncores n] [--outfile filename] [--quick] [--quiet] [--rank r --world n --rendezvous file [--device d]] <filename.pbrt> ...\n");
            return 0;
        } else files.push_back(argv[i]);
    }
    if (files.empty()) files.push_back("-");   // "Parse scene from standard input" (main/pbrt.cpp:71-74)
    PbrtGpuComm *comm = nullptr;
    if (world > 1) {
        if (rank < 0 || rank >= world || rendezvous.empty()) { fprintf(stderr, "pbrt_b200: --world needs --rank and --rendezvous\n"); return 1; }
        if (pbrtgpu_set_device(localRank >= 0 ? localRank : rank) != PBRTGPU_OK) { fprintf(stderr, "Error: %s\n", pbrtgpu_last_error()); return 1; }
        unsigned char id[PBRTGPU_COMM_ID_BYTES];
        if (!ExchangeId(rendezvous, rank, id)) { fprintf(stderr, "Error: rendezvous through \"%s\" failed: %s\n", rendezvous.c_str(), pbrtgpu_last_error()); return 1; }
        if (pbrtgpu_comm_create(id, rank, world, &comm) != PBRTGPU_OK) { fprintf(stderr, "Error: %s\n", pbrtgpu_last_error()); return 1; }
    } else if (localRank >= 0 && pbrtgpu_set_device(localRank) != PBRTGPU_OK) { fprintf(stderr, "Error: %s\n", pbrtgpu_last_error()); return 1; }
    for (const std::string &f : files) {
        pb::LoadedScene ls;
        if (!pb::LoadSceneFile(f, opt, &ls)) {
            fprintf(stderr, "Error: %s\n", ls.error.c_str());
            return 1;
        }
        ls.renderer->comm = comm;
        ls.renderer->commRank = rank;
        if (!ls.renderer->Render(ls.scene.get())) return 1;
        if (!opt.quiet && rank == 0) {
            const PbrtGpuStats &s = ls.renderer->lastStats;
            printf("Rendering: %.3f s on device (%.2f Msamples/s), %llu camera samples, %llu+%llu rays\n",
                   s.renderMs * 1e-3, s.cameraSamples / (s.renderMs * 1e3), (unsigned long long)s.cameraSamples,
                   (unsigned long long)s.closestRays, (unsigned long long)s.shadowRays);
        }
    }
    if (comm) {
        pbrtgpu_comm_destroy(comm);
        if (rank == 0) remove(rendezvous.c_str());
    }
    return 0;
}
