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
DROPIN   := paris_b200/libparis_b200_dropin.so
CPPSRC   := paris_b200/cpp/b200/backend.cpp paris_b200/cpp/pipeline.cpp paris_b200/cpp/dropin_api.cpp
CPPHDR   := paris_b200/cpp/b200/backend.h paris_b200/cpp/pipeline.h paris_b200/cpp/paris_types.h

all: lib oracle

lib: $(LIB) $(DROPIN)

$(CSRC)/%.o: $(CSRC)/%.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -ccbin $(HOSTCXX) -o $@ $(OBJS) -lcudart

# the C++ host layer (namespace paris::b200 + stage wrappers + the reference-shaped loop), over the C ABI
$(DROPIN): $(CPPSRC) $(CPPHDR) include/paris_b200.h $(LIB)
	$(HOSTCXX) -std=c++14 -O2 -fPIC -Wall -Wextra -shared -o $@ $(CPPSRC) -Lparis_b200 -lparis_b200 -Wl,-rpath,'$$ORIGIN'

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(OBJS) $(LIB) $(DROPIN)
	$(MAKE) -C oracle clean

.PHONY: all lib oracle clean
