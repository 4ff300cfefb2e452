# Build of the product library (nvcc, sm_100a only) and of the test-only kernel-logic emulator.
#   make            -> edsparser_b200/libedsparser_b200.so   (the product; needs a B200 at run time)
#   make emu        -> tests/emu/libedsparser_emu.so         (TEST INFRASTRUCTURE: same sources under g++
#                      with tests/emu/cuda_emu.h; used by the CPU test tier only, never by the product)
#   make oracle     -> oracle/ (CPU restatement + reference build, test infrastructure)
NVCC     ?= /usr/local/cuda/bin/nvcc
CXX      := $(firstword $(wildcard /usr/bin/g++) g++)
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := -std=c++17 -O3 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v
CSRC     := edsparser_b200/csrc
SRCS     := $(CSRC)/msa.cu $(CSRC)/leds.cu $(CSRC)/vcf.cu $(CSRC)/capi.cu $(CSRC)/shard.cu
HDRS     := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh) include/edsparser_b200.h
OBJS     := $(patsubst $(CSRC)/%.cu,build/%.o,$(SRCS))
EMUOBJS  := $(patsubst $(CSRC)/%.cu,build/emu_%.o,$(SRCS)) build/emu_runtime.o

.PHONY: all lib emu oracle host clean
all: lib host
lib: edsparser_b200/libedsparser_b200.so

build:
	mkdir -p build

build/%.o: $(CSRC)/%.cu $(HDRS) | build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)

edsparser_b200/libedsparser_b200.so: $(OBJS)
	$(NVCC) -shared $(ARCH) -o $@ $^ -ldl

emu: tests/emu/libedsparser_emu.so

build/emu_%.o: $(CSRC)/%.cu $(HDRS) tests/emu/cuda_emu.h | build
	$(CXX) -std=c++17 -O1 -g -fPIC -Wall -Wno-unused-function -Wno-unknown-pragmas -DEDSB_EMU -Itests/emu -x c++ -c $< -o $@

build/emu_runtime.o: tests/emu/cuda_emu.cpp tests/emu/cuda_emu.h | build
	$(CXX) -std=c++17 -O1 -g -fPIC -Wall -Itests/emu -c $< -o $@

tests/emu/libedsparser_emu.so: $(EMUOBJS)
	$(CXX) -shared -pthread -o $@ $^ -ldl

# host layer: the reference's transforms API (C++17) + the msa2eds / eds2leds tools, over the C ABI
HOST     := edsparser_b200/host
HOSTINC  := -I$(HOST)/include -I$(HOST)/tools
host: edsparser_b200/bin/msa2eds edsparser_b200/bin/eds2leds edsparser_b200/bin/vcf2eds

edsparser_b200/libedsparser_host.a: $(HOST)/src/host.cpp $(wildcard $(HOST)/include/edsparser/*.hpp $(HOST)/include/edsparser/*/*.hpp) include/edsparser_b200.h | build
	$(CXX) -std=c++17 -O2 -fPIC -Wall $(HOSTINC) -c $(HOST)/src/host.cpp -o build/host.o
	ar rcs $@ build/host.o

edsparser_b200/bin/%: $(HOST)/tools/%.cpp $(HOST)/tools/cli_common.hpp edsparser_b200/libedsparser_host.a edsparser_b200/libedsparser_b200.so
	mkdir -p edsparser_b200/bin
	$(CXX) -std=c++17 -O2 -Wall $(HOSTINC) $< edsparser_b200/libedsparser_host.a -Ledsparser_b200 -ledsparser_b200 -Wl,-rpath,'$$ORIGIN/..' -o $@

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf build edsparser_b200/libedsparser_b200.so edsparser_b200/libedsparser_host.a edsparser_b200/bin tests/emu/libedsparser_emu.so
