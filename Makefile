# Builds the product library (CUDA, sm_100a), the peaks microbenchmark and the test oracle.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v
CSRC      := bisip_b200/csrc
LIB       := $(CSRC)/libbisip_b200.so
HDRS      := $(wildcard $(CSRC)/*.cuh) include/bisip_b200.h

all: $(LIB) tools/peaks oracle

$(LIB): $(CSRC)/api.cu $(HDRS)
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(CSRC)/api.cu 2> $(CSRC)/ptxas.log || (cat $(CSRC)/ptxas.log; false)

# developer build with per-phase cycle counters (not shipped, not loaded by the package)
dbg: $(CSRC)/api.cu $(HDRS)
	$(NVCC) $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DBISIP_PHASE_TIMING -o $(CSRC)/libbisip_b200_dbg.so $(CSRC)/api.cu

tools/peaks: tools/peaks.cu
	$(NVCC) $(ARCH) -O3 -lineinfo -o $@ $<

oracle:
	python -c "from oracle import oracle; oracle.build()"
	./oracle/build_ref.sh

clean:
	rm -f $(LIB) tools/peaks $(CSRC)/ptxas.log
	rm -rf oracle/_build

.PHONY: all oracle clean dbg
