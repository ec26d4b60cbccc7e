# Build libgpar_b200.so (sm_100a only) and the oracle's C restatement.
NVCC      ?= nvcc
PKG       := gpar-at-scale_b200
CSRC      := $(PKG)/csrc
LIBDIR    := $(PKG)/lib
NVFLAGS   := -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v
SRCS      := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,build/%.o,$(SRCS))

all: $(LIBDIR)/libgpar_b200.so oracle

build/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/gpar_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIBDIR)/libgpar_b200.so: $(OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) -shared -o $@ $(OBJS)

oracle: oracle/_build/liboracle_c.so
oracle/_build/liboracle_c.so: $(wildcard oracle/c/*.c)
	@mkdir -p oracle/_build
	gcc -O3 -march=x86-64-v3 -fPIC -shared -fopenmp -o $@ $^ -lm

clean:
	rm -rf build $(LIBDIR) oracle/_build

.PHONY: all oracle clean
