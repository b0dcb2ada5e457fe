// temporary stub -- replaced by pbrtgpu.cu
#include "../../../include/pbrtgpu.h"
extern "C" {
int pbrtgpu_device_count(void) { int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0; }
int pbrtgpu_set_device(int) { return PBRTGPU_ENODEVICE; }
const char *pbrtgpu_last_error(void) { return "stub"; }
int pbrtgpu_scene_create(const PbrtGpuSceneDesc *, PbrtGpuScene **) { return PBRTGPU_ENODEVICE; }
void pbrtgpu_scene_destroy(PbrtGpuScene *) {}
int pbrtgpu_render(PbrtGpuScene *, const PbrtGpuRenderDesc *, PbrtGpuPixel *, PbrtGpuStats *) { return PBRTGPU_ENODEVICE; }
int pbrtgpu_render_device(PbrtGpuScene *, const PbrtGpuRenderDesc *, void *, void *, PbrtGpuStats *) { return PBRTGPU_ENODEVICE; }
int pbrtgpu_intersect(PbrtGpuScene *, const PbrtGpuRay *, int, PbrtGpuHit *) { return PBRTGPU_ENODEVICE; }
int pbrtgpu_intersect_p(PbrtGpuScene *, const PbrtGpuRay *, int, uint8_t *) { return PBRTGPU_ENODEVICE; }
int pbrtgpu_ld_samples(const PbrtGpuRenderDesc *, int, float *) { return PBRTGPU_ENODEVICE; }
}
