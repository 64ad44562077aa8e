# Top-level build: the sm_100a shared library (the product), the C++ backend shim and the CPU oracles.
#   make            -> paris_b200/libparis_b200.so + oracle/ (checkers)
#   make lib        -> product library only
NVCC     ?= /usr/local/cuda/bin/nvcc
HOSTCXX  ?= $(firstword $(wildcard /usr/bin/g++) g++)
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -O3 -std=c++17 -lineinfo -ccbin $(HOSTCXX) -Xcompiler -fPIC -Xcompiler -Wall \
            -Xcompiler -ffp-contract=off $(EXTRA_NVFLAGS)
CSRC     := paris_b200/csrc
SRCS     := $(CSRC)/api.cu $(CSRC)/weight.cu $(CSRC)/filter.cu $(CSRC)/backproject.cu $(CSRC)/backproject_tma.cu $(CSRC)/phantom.cu $(CSRC)/group.cu
OBJS     := $(SRCS:.cu=.o)
HDRS     := $(wildcard $(CSRC)/*.cuh) include/paris_b200.h
LIB      := paris_b200/libparis_b200.so
DROPIN   := paris_b200/libparis_b200_dropin.so
CPPDIR   := paris_b200/cpp
CPPSRC   := $(CPPDIR)/b200/backend.cpp $(CPPDIR)/pipeline.cpp $(CPPDIR)/dropin_api.cpp $(CPPDIR)/io_api.cpp \
            $(CPPDIR)/io/his.cpp $(CPPDIR)/io/ddbvf.cpp $(CPPDIR)/io/filesystem.cpp $(CPPDIR)/io/source.cpp \
            $(CPPDIR)/io/sink.cpp $(CPPDIR)/program_options.cpp $(CPPDIR)/task.cpp
CPPHDR   := $(wildcard $(CPPDIR)/*.h $(CPPDIR)/b200/*.h $(CPPDIR)/io/*.h)
CLI      := paris_b200/bin/paris_b200

all: lib oracle

lib: $(LIB) $(DROPIN) $(CLI)

$(CSRC)/%.o: $(CSRC)/%.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -ccbin $(HOSTCXX) -o $@ $(OBJS) -lcudart

# the C++ host layer (namespace paris::b200 + stage wrappers + the reference-shaped loop), over the C ABI
$(DROPIN): $(CPPSRC) $(CPPHDR) include/paris_b200.h $(LIB)
	$(HOSTCXX) -std=c++17 -O2 -fPIC -Wall -Wextra -shared -o $@ $(CPPSRC) -Lparis_b200 -lparis_b200 -Wl,-rpath,'$$ORIGIN'

# the command-line driver (src/main.cpp of the reference on the B200 backend)
$(CLI): $(CPPDIR)/main.cpp $(DROPIN)
	mkdir -p paris_b200/bin
	$(HOSTCXX) -std=c++17 -O2 -Wall -Wextra -pthread -o $@ $(CPPDIR)/main.cpp -Lparis_b200 -lparis_b200_dropin -lparis_b200 -Wl,-rpath,'$$ORIGIN/..'

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(OBJS) $(LIB) $(DROPIN) $(CLI)
	$(MAKE) -C oracle clean

.PHONY: all lib oracle clean

# Checked build of the library (compute-sanitizer is closed on the GPU pool): every shared-memory sample load of the
# backprojection is tested against the CTA's ring of staged boxes (csrc/backproject_tma.cu, PB_BOUNDS_CHECK); select it
# with PARIS_B200_LIB=paris_b200/libparis_b200_check.so, read the counters with scripts/sanitize_case.py
check-lib: paris_b200/libparis_b200_check.so
paris_b200/libparis_b200_check.so: $(SRCS) $(HDRS)
	$(NVCC) $(NVFLAGS) -DPB_BOUNDS_CHECK -c $(CSRC)/backproject_tma.cu -o /tmp/pb_backproject_tma_check.o
	$(NVCC) $(ARCH) -shared -ccbin $(HOSTCXX) -o $@ $(filter-out $(CSRC)/backproject_tma.o,$(OBJS)) /tmp/pb_backproject_tma_check.o -lcudart

