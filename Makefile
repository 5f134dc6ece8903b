# Builds libiak3d.so (sm_100a only) and the host-compiled copy of the scalar math used by the CPU tests.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden
CSRC      := iak3d_b200/csrc
OUT       := iak3d_b200/lib
OBJS      := $(OUT)/api.o $(OUT)/api_fit.o $(OUT)/tables.o $(OUT)/covbuild.o $(OUT)/chol.o $(OUT)/gemm.o $(OUT)/grad.o $(OUT)/mat.o $(OUT)/design.o $(OUT)/voronoi.o $(OUT)/peer.o $(OUT)/tail.o
HDRS      := $(CSRC)/common.cuh $(CSRC)/chol_pair.cuh $(CSRC)/iak3d_math.h include/iak3d.h

all: $(OUT)/libiak3d.so tests/_hostmath.so

$(OUT)/libiak3d.so: $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart_static -lpthread -ldl -lrt

# closed forms round like the reference's operation order: no FMA contraction in the table / assembly kernels
$(OUT):
	mkdir -p $(OUT)
$(OBJS): | $(OUT)

$(OUT)/tables.o: $(CSRC)/tables.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -fmad=false -c $< -o $@
$(OUT)/covbuild.o: $(CSRC)/covbuild.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -fmad=false -c $< -o $@
$(OUT)/voronoi.o: $(CSRC)/voronoi.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -fmad=false -c $< -o $@
$(OUT)/design.o: $(CSRC)/design.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -fmad=false -c $< -o $@
$(OUT)/%.o: $(CSRC)/%.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -c $< -o $@

# test-only: the same iak3d_math.h compiled for the host (never loaded by the product)
tests/_hostmath.so: tests/hostmath.cpp $(CSRC)/iak3d_math.h
	g++ -O2 -ffp-contract=off -fPIC -shared -o $@ tests/hostmath.cpp -lm

clean:
	rm -f $(OUT)/*.o $(OUT)/libiak3d.so tests/_hostmath.so
.PHONY: all clean
