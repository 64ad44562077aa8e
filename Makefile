# Top-level build: the sm_100a shared library (the product), the C++ backend shim and the CPU oracles.
#   make            -> paris_b200/libparis_b200.so + oracle/ (checkers)
#   make lib        -> product library only
NVCC     ?= /usr/local/cuda/bin/nvcc
HOSTCXX  ?= $(firstword $(wildcard /usr/bin/g++) g++)
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -O3 -std=c++17 -lineinfo -ccbin $(HOSTCXX) -Xcompiler -fPIC -Xcompiler -Wall \
            -Xcompiler -ffp-contract=off $(EXTRA_NVFLAGS)
CSRC     := paris_b200/csrc
SRCS     := $(CSRC)/api.cu $(CSRC)/weight.cu $(CSRC)/filter.cu $(CSRC)/backproject.cu $(CSRC)/backproject_tma.cu $(CSRC)/phantom.cu
OBJS     := $(SRCS:.cu=.o)
HDRS     := $(wildcard $(CSRC)/*.cuh) include/paris_b200.h
LIB      := paris_b200/libparis_b200.so

all: lib oracle

lib: $(LIB)

$(CSRC)/%.o: $(CSRC)/%.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -ccbin $(HOSTCXX) -o $@ $(OBJS) -lcudart

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(OBJS) $(LIB)
	$(MAKE) -C oracle clean

.PHONY: all lib oracle clean
