# Builds libvqgnn.so (the C-ABI CUDA library, sm_100a only) in-tree so it travels with the gpurun snapshot.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall -Xptxas -v
SRC := $(wildcard vq_gnn_b200/csrc/*.cu)
OBJ := $(SRC:.cu=.o)
LIB := vq_gnn_b200/libvqgnn.so

all: $(LIB)

%.o: %.cu $(wildcard vq_gnn_b200/csrc/*.cuh) include/vqgnn.h
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ)

# The UNMODIFIED reference's two Python packages, copied (never committed: oracle/_ref/ is git-ignored) next to the
# oracle so that the GPU box -- which has no /root/reference -- can run the real reference as the CPU arm of bench.py
# (`--impl reference`, kind "reference") and in tests/test_oracle_vs_reference.py.  Test/bench infrastructure only.
REFERENCE ?= /root/reference
oracle_ref:
	@if [ -d $(REFERENCE)/vq_gnn_v2 ]; then \
	  rm -rf oracle/_ref && mkdir -p oracle/_ref && \
	  (cd $(REFERENCE) && find vq_gnn_v1 vq_gnn_v2 -name '*.py' | tar -cf - -T -) | tar -xf - -C oracle/_ref && \
	  echo "oracle/_ref: copied `find oracle/_ref -name '*.py' | wc -l` files from $(REFERENCE)"; \
	else echo "oracle/_ref: $(REFERENCE) not present, keeping what is there"; fi

clean:
	rm -f $(OBJ) $(LIB)
.PHONY: all clean oracle_ref
