# Builds libvqgnn.so (the C-ABI CUDA library, sm_100a only) in-tree so it travels with the gpurun snapshot.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall -Xptxas -v
SRC := $(wildcard vq_gnn_b200/csrc/*.cu)
OBJ := $(SRC:.cu=.o)
LIB := vq_gnn_b200/libvqgnn.so

all: $(LIB)

%.o: %.cu vq_gnn_b200/csrc/common.cuh vq_gnn_b200/csrc/mp_common.cuh include/vqgnn.h
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ)

clean:
	rm -f $(OBJ) $(LIB)
.PHONY: all clean
