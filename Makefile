# Builds libkmbart_sm100.so (the C-ABI CUDA library) and the standalone CUDA test binaries.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr
PKG := km-bart_b200
SRCS := $(wildcard $(PKG)/csrc/*.cu)
OBJS := $(patsubst $(PKG)/csrc/%.cu,build/%.o,$(SRCS))
LIB := $(PKG)/libkmbart_sm100.so

all: $(LIB) tests

build/%.o: $(PKG)/csrc/%.cu $(wildcard $(PKG)/csrc/*.cuh) include/kmbart.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart

tests: build/gemm_test
build/gemm_test: tests/cuda/gemm_test.cu $(LIB)
	$(NVCC) $(ARCH) -O2 -std=c++17 -o $@ $< -L$(PKG) -lkmbart_sm100 -Xlinker -rpath -Xlinker '$$ORIGIN/../$(PKG)'

clean:
	rm -rf build $(LIB)
.PHONY: all tests clean
