# Build of the B200-native MyTinyGL back end, its CPU oracle and the reference builds.
#
#   make product   mytinygl_b200/lib/libMyTinyGL_b200.so  (gl* front end + sm_100a back end) and
#                  tests/scenes/_build/libscenes_b200.so  (scene recipes linked against it)
#   make oracle    oracle/_build/libfront_oracle.so       (front end + CPU restatement; tests only)
#   make ref       oracle/_ref/libref_{strict,shipped,shipped_v3}.so from the UNMODIFIED reference
#                  sources where they lie (needs $(REF); skipped when the tree is absent)
#
# Everything that produces pixels on the host is compiled with IEEE semantics.
REF      ?= /root/reference
CUDA     ?= /usr/local/cuda
NVCC     ?= $(CUDA)/bin/nvcc
CXX      ?= g++
CC       ?= gcc

IEEE     := -fno-fast-math -ffp-contract=off
CXXFLAGS := -O2 -g -fPIC -std=c++17 -Wall -Wextra $(IEEE) -Iinclude
CFLAGS   := -O2 -g -fPIC -std=gnu99 -Wall $(IEEE) -Iinclude
# -fmad=false: the reference's strict build never contracts a*b+c
NVFLAGS  := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true \
            -prec-sqrt=true -ftz=false -Xcompiler -fPIC -Iinclude

FRONT_SRC := mytinygl_b200/csrc/front/gl_front.cpp mytinygl_b200/csrc/front/gl_api_state.cpp
FRONT_HDR := $(wildcard mytinygl_b200/csrc/front/*.h) include/mtgl_dev.h include/mtgl_context.h include/GL/gl.h
DEV_SRC   := $(wildcard mytinygl_b200/csrc/dev/*.cu)
DEV_HDR   := $(wildcard mytinygl_b200/csrc/dev/*.cuh) include/mtgl_dev.h
SCENES    := tests/scenes/scenes.c
REF_SRC   := $(REF)/src/gl_api.c $(REF)/src/raster.c $(REF)/src/textures.c $(REF)/src/vbo.c $(REF)/src/lists.c

PRODUCT   := mytinygl_b200/lib/libMyTinyGL_b200.so
SCENELIB  := tests/scenes/_build/libscenes_b200.so
ORACLELIB := oracle/_build/libfront_oracle.so
REFLIBS   := oracle/_ref/libref_strict.so oracle/_ref/libref_shipped.so oracle/_ref/libref_shipped_v3.so

.PHONY: all product oracle ref clean
all: product oracle ref

product: $(PRODUCT) $(SCENELIB)
oracle: $(ORACLELIB)
ifneq ($(wildcard $(REF)/src/raster.c),)
ref: $(REFLIBS)
else
ref:
	@echo "reference tree $(REF) not present: keeping prebuilt oracle/_ref"
endif

build/front/%.o: mytinygl_b200/csrc/front/%.cpp $(FRONT_HDR)
	@mkdir -p $(dir $@)
	$(CXX) $(CXXFLAGS) -c $< -o $@

build/dev/%.o: mytinygl_b200/csrc/dev/%.cu $(DEV_HDR)
	@mkdir -p $(dir $@)
	$(NVCC) $(NVFLAGS) -c $< -o $@

FRONT_OBJ := $(patsubst mytinygl_b200/csrc/front/%.cpp,build/front/%.o,$(FRONT_SRC))
DEV_OBJ   := $(patsubst mytinygl_b200/csrc/dev/%.cu,build/dev/%.o,$(DEV_SRC))

$(PRODUCT): $(FRONT_OBJ) $(DEV_OBJ)
	@mkdir -p $(dir $@)
	$(NVCC) -shared -o $@ $^ -Xlinker -Bsymbolic -cudart shared -L$(CUDA)/lib64

$(SCENELIB): $(SCENES) tests/scenes/harness_b200.c $(PRODUCT)
	@mkdir -p $(dir $@)
	$(CC) $(CFLAGS) -shared -o $@ $(SCENES) tests/scenes/harness_b200.c \
	    -Lmytinygl_b200/lib -lMyTinyGL_b200 -Wl,-rpath,'$$ORIGIN/../../../mytinygl_b200/lib' -lm

build/oracle/mtgl_oracle.o: oracle/mtgl_oracle.c include/mtgl_dev.h
	@mkdir -p $(dir $@)
	$(CC) $(CFLAGS) -c $< -o $@

$(ORACLELIB): $(FRONT_OBJ) build/oracle/mtgl_oracle.o $(SCENES) tests/scenes/harness_b200.c
	@mkdir -p $(dir $@)
	$(CC) $(CFLAGS) -DMTGL_HARNESS_ORACLE -c tests/scenes/harness_b200.c -o build/oracle/harness.o
	$(CC) $(CFLAGS) -c $(SCENES) -o build/oracle/scenes.o
	$(CXX) -shared -o $@ $(FRONT_OBJ) build/oracle/mtgl_oracle.o build/oracle/harness.o build/oracle/scenes.o \
	    -Wl,-Bsymbolic -lm

# The reference, unmodified, compiled in place.  strict = canonical parity oracle (bit-identical to -O0),
# shipped = the reference's own Makefile:3 flags (timing baseline), shipped_v3 = the same with a portable
# -march for GPU-box hosts that lack this container's ISA extensions.
REF_COMMON := -std=gnu99 -fPIC -shared -w -I$(REF)/include -I$(REF) -Wl,-Bsymbolic
oracle/_ref/libref_strict.so: $(REF_SRC) oracle/ref_shim.c $(SCENES)
	@mkdir -p $(dir $@)
	$(CC) -O2 $(IEEE) $(REF_COMMON) -o $@ $(REF_SRC) oracle/ref_shim.c $(SCENES) -lm
oracle/_ref/libref_shipped.so: $(REF_SRC) oracle/ref_shim.c $(SCENES)
	@mkdir -p $(dir $@)
	$(CC) -O3 -march=native -ffast-math $(REF_COMMON) -o $@ $(REF_SRC) oracle/ref_shim.c $(SCENES) -lm
oracle/_ref/libref_shipped_v3.so: $(REF_SRC) oracle/ref_shim.c $(SCENES)
	@mkdir -p $(dir $@)
	$(CC) -O3 -march=x86-64-v3 -ffast-math $(REF_COMMON) -o $@ $(REF_SRC) oracle/ref_shim.c $(SCENES) -lm

clean:
	rm -rf build mytinygl_b200/lib tests/scenes/_build oracle/_build
