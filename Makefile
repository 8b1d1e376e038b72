# Builds the C-ABI shared library of sm_100a kernels (no GPU needed: nvcc cross-compiles).
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall
SRC := $(wildcard graphtrans_b200/csrc/*.cu)
OBJ := $(patsubst graphtrans_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := graphtrans_b200/lib/libgraphtrans_b200.so

all: $(LIB)

build/%.o: graphtrans_b200/csrc/%.cu graphtrans_b200/csrc/common.cuh include/graphtrans_b200.h $(wildcard graphtrans_b200/csrc/*.cuh)
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	@mkdir -p graphtrans_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -cudart shared

clean:
	rm -rf build $(LIB)
.PHONY: all clean
