# Builds the product in-tree for sm_100a only:
#   gpupsat_b200/libgpsat.so   C-ABI library (include/gpsat.h)
#   gpupsat_b200/gpupsat       command-line front end with the reference's surface
# and the test infrastructure (oracle/, tests/emu).  `make` = everything that can be built on this machine.
NVCC ?= /usr/local/cuda/bin/nvcc
CXX  ?= g++
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -std=c++17 -O3 $(ARCH) -lineinfo -Xcompiler -fPIC -Xptxas -v $(EXTRA_NVFLAGS)
CSRC := gpupsat_b200/csrc
DEPS := $(wildcard $(CSRC)/*.h $(CSRC)/*.inl include/*.h)

all: lib cli oracle emu
lib: gpupsat_b200/libgpsat.so
cli: gpupsat_b200/gpupsat

build/%.o: $(CSRC)/%.cu $(DEPS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)
build/host_formula.o: $(CSRC)/host_formula.cpp $(DEPS)
	@mkdir -p build
	$(CXX) -std=c++17 -O2 -Wall -fPIC -c $< -o $@

gpupsat_b200/libgpsat.so: build/kernels.o build/gpsat_api.o build/gpsat_multi.o build/host_formula.o
	$(NVCC) -shared $(ARCH) -o $@ $^ -cudart shared -ldl

gpupsat_b200/gpupsat: $(CSRC)/gpupsat_main.cpp gpupsat_b200/libgpsat.so include/gpsat.h
	$(CXX) -std=c++17 -O2 -Wall -o $@ $< -Lgpupsat_b200 -lgpsat -Wl,-rpath,'$$ORIGIN'

oracle:
	$(MAKE) -C oracle all
emu:
	$(CXX) -std=c++17 -O2 -Wall -Wno-unused-variable -fPIC -shared -o tests/emu/libgpsat_emu.so tests/emu/emu.cpp $(CSRC)/host_formula.cpp

clean:
	rm -rf build gpupsat_b200/libgpsat.so gpupsat_b200/gpupsat tests/emu/libgpsat_emu.so
.PHONY: all lib cli oracle emu clean
