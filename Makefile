# Builds libgf_b200.so (sm_100a only) in-tree.  `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
CSRC := ground_fusion_b200/csrc
OUT  := ground_fusion_b200/libgf_b200.so
# -fmad=false: the front end is bit-exact with OpenCV's separately-rounded float ops (FMA only where written)
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -fmad=false
SRCS := $(wildcard $(CSRC)/*.cu)
HDRS := $(wildcard $(CSRC)/*.cuh) include/gf_b200.h

all: $(OUT) oracle

$(OUT): $(SRCS) $(HDRS)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(SRCS) -Xptxas -v 2> build_ptxas.log || (cat build_ptxas.log; false)

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(OUT) build_ptxas.log; $(MAKE) -C oracle clean
.PHONY: all oracle clean
