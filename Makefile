# Top-level build: the CUDA hot-path library, the host-side drop-in library and
# CLI, and (test infrastructure) the oracle.  `python -c "import __graft_entry__ as g; g.build()"`
# runs exactly this.
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
PKG       := pbrt_v2_b200
HOSTDIR   := $(PKG)/csrc/host
DEVDIR    := $(PKG)/csrc/device
# -fmad=false: the reference is compiled for x86-64 without FMA; contraction would
# change hit/miss decisions and with them the RNG stream.
NVFLAGS   := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
             -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
             -Xcompiler -fPIC -Xptxas -v
# host code must not contract to FMA either (no -march=native)
HOSTFLAGS := -O2 -std=c++17 -ffp-contract=off -fPIC -Wall -Wno-unused-function

HOSTSRC   := transform parser shapes bvh measured envmap exr tga renderer render_call capi
HOSTOBJ   := $(patsubst %,build/host/%.o,$(HOSTSRC))
DEVSRC    := $(wildcard $(DEVDIR)/*.cu)
DEVHDR    := $(wildcard $(DEVDIR)/*.cuh) $(wildcard $(DEVDIR)/*.inc) $(wildcard $(DEVDIR)/*.h) include/pbrtgpu.h
# extra -D switches for kernel experiments, e.g.  make NVDEFS=-DPBRT_SPECULAR=1  (see device/dshade.cuh) or
# -DK_TRACE_MIN_BLOCKS=8; the shipped library is built with none
NVDEFS    ?=

# kernel experiments, built next to the product library so that a GPU visit can A/B them without compiling on the box
# (scripts/ab_variants.py; tests/test_zz_gpu_variants.py runs the parity suite over them).  Nothing loads them by default.
#   make variants VARIANT_NAMES="pairs foo"   with   VARIANT_DEFS_foo := -DSOMETHING=1
VARIANT_NAMES ?=
# Kernel experiments of the current round (the earlier ones and their verdicts: profiles/README.md).  Traversal variants
# that lost on the B200 and can still be built:  pairs (-DPBRT_PAIR_NODES=1), ww (-DPBRT_TRACE_LOOP=1),
# sstack (-DPBRT_SMEM_STACK=32), top9_sstack (-DPBRT_TOP_LEVELS=9 -DPBRT_SMEM_STACK=32)
VARIANT_DEFS_pairs := -DPBRT_PAIR_NODES=1
VARIANT_DEFS_ww := -DPBRT_TRACE_LOOP=1
VARIANT_DEFS_sstack := -DPBRT_SMEM_STACK=32
VARIANT_DEFS_top9_sstack := -DPBRT_TOP_LEVELS=9 -DPBRT_SMEM_STACK=32
# register bounds of the vertex wavefront's kernels
VARIANT_DEFS_ws5 := -DK_WSHADE_MIN_BLOCKS=5
VARIANT_DEFS_ws6 := -DK_WSHADE_MIN_BLOCKS=6
VARIANT_DEFS_we10 := -DK_WEVAL_MIN_BLOCKS=10
# MetalMaterial on the device (host-verified; first GPU run: PBRTGPU_LIB=ab/libpbrtgpu_metal.so pytest tests/test_gpu_specular.py -m gpu)
VARIANT_DEFS_metal := -DPBRT_METAL=1
# ... plus InfiniteAreaLight and pixel filters through the weight table: the whole of SURVEY 8(f) rank 4 -- and the
# zero-initialised pivot arrays of MatInverse (danim.cuh PBRT_SAFE_INVERSE).
# PBRTGPU_LIB=ab/libpbrtgpu_f4.so pytest tests/test_gpu_refusals.py tests/test_gpu_specular.py -m gpu
VARIANT_DEFS_f4 := -DPBRT_METAL=1 -DPBRT_INFINITE=1 -DPBRT_FILTER=1 -DPBRT_SAFE_INVERSE=1
# leaf tests in batches of >= n lanes (PBRT_TRACE_LOOP 2), on top of ws6
VARIANT_DEFS_ws6_lb8 := -DK_WSHADE_MIN_BLOCKS=6 -DPBRT_TRACE_LOOP=2 -DPBRT_LEAF_BATCH=8
VARIANT_DEFS_ws6_lb12 := -DK_WSHADE_MIN_BLOCKS=6 -DPBRT_TRACE_LOOP=2 -DPBRT_LEAF_BATCH=12
VARIANT_DEFS_ws6_lb20 := -DK_WSHADE_MIN_BLOCKS=6 -DPBRT_TRACE_LOOP=2 -DPBRT_LEAF_BATCH=20
VARIANTS  := $(patsubst %,$(PKG)/ab/libpbrtgpu_%.so,$(VARIANT_NAMES))

all: $(PKG)/libpbrtgpu.so $(PKG)/libpbrthost.so $(PKG)/pbrt_b200 oracle variants

variants: $(VARIANTS)

$(PKG)/ab/libpbrtgpu_%.so: $(DEVSRC) $(DEVHDR)
	@mkdir -p $(PKG)/ab build
	$(NVCC) $(NVFLAGS) $(VARIANT_DEFS_$*) -shared $(DEVSRC) -o $@ 2> build/ptxas_$*.log || (cat build/ptxas_$*.log; false)

$(PKG)/libpbrtgpu.so: $(DEVSRC) $(DEVHDR)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) $(NVDEFS) -shared $(DEVSRC) -o $@ 2> build/ptxas.log || (cat build/ptxas.log; false)
	@grep -E "registers|spill|error" build/ptxas.log | sort | uniq -c | sort -rn | head -20 || true

build/host/%.o: $(HOSTDIR)/%.cpp $(HOSTDIR)/scene.h $(HOSTDIR)/geom.h include/pbrtgpu.h
	@mkdir -p build/host
	$(CXX) $(HOSTFLAGS) -c $< -o $@

$(PKG)/libpbrthost.so: $(HOSTOBJ) $(PKG)/libpbrtgpu.so
	$(CXX) -shared $(HOSTOBJ) -o $@ -L$(PKG) -lpbrtgpu -Wl,-rpath,'$$ORIGIN' -lpthread -ldl -lz

$(PKG)/pbrt_b200: build/host/main.o $(PKG)/libpbrthost.so
	$(CXX) build/host/main.o -o $@ -L$(PKG) -lpbrthost -lpbrtgpu -Wl,-rpath,'$$ORIGIN' -lpthread -ldl

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf build $(PKG)/*.so $(PKG)/ab $(PKG)/pbrt_b200
	$(MAKE) -C oracle clean

.PHONY: all oracle variants clean
