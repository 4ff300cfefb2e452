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

# The reference's own C++ tests, UNMODIFIED (sources stay under /root/reference; only binaries are produced, into the
# git-ignored tests/_refbin/ which travels to the GPU box): compiled with -Dmain=reference_main next to
# tests/refwrap/run_each.cpp (runs the test functions one by one), linked against this repo's host layer + the product
# library (*_gpu), against the emulator build (*_emu, CPU tier) and against the unmodified reference library (*_ref:
# which of the reference's asserts hold against the reference itself -> tests/golden/ref_cpp_tests.json).
REF      ?= /root/reference
REFTESTS := test_msa test_merge test_sources test_stats
REFFLAGS := -std=c++17 -O1 -rdynamic -Wno-unused-variable -Wno-unused-result
REFMAIN  := -Dmain=reference_main -c
REFOURS  := $(HOSTINC) -I$(HOST)/include/edsparser
reftests: $(patsubst %,tests/_refbin/%_gpu,$(REFTESTS)) $(patsubst %,tests/_refbin/%_emu,$(REFTESTS)) $(patsubst %,tests/_refbin/%_ref,$(REFTESTS))
tests/_refbin/%_gpu: $(REF)/tests/cpp/%.cpp tests/refwrap/run_each.cpp edsparser_b200/libedsparser_host.a edsparser_b200/libedsparser_b200.so
	mkdir -p tests/_refbin
	$(CXX) $(REFFLAGS) $(REFOURS) $(REFMAIN) $< -o $@.o
	$(CXX) $(REFFLAGS) $@.o tests/refwrap/run_each.cpp edsparser_b200/libedsparser_host.a -Ledsparser_b200 -ledsparser_b200 -ldl -Wl,-rpath,'$$ORIGIN/../../edsparser_b200' -o $@ && rm -f $@.o
tests/_refbin/%_emu: $(REF)/tests/cpp/%.cpp tests/refwrap/run_each.cpp edsparser_b200/libedsparser_host.a tests/emu/libedsparser_emu.so
	mkdir -p tests/_refbin
	$(CXX) $(REFFLAGS) $(REFOURS) $(REFMAIN) $< -o $@.o
	$(CXX) $(REFFLAGS) $@.o tests/refwrap/run_each.cpp edsparser_b200/libedsparser_host.a -Ltests/emu -ledsparser_emu -ldl -Wl,-rpath,'$$ORIGIN/../emu' -o $@ && rm -f $@.o
tests/_refbin/%_ref: $(REF)/tests/cpp/%.cpp tests/refwrap/run_each.cpp oracle/_ref/libedsparser_ref.a
	mkdir -p tests/_refbin
	$(CXX) $(REFFLAGS) -Ioracle/sdsl_standin -I$(REF)/src/cpp/lib $(REFMAIN) $< -o $@.o
	$(CXX) $(REFFLAGS) -fopenmp $@.o tests/refwrap/run_each.cpp oracle/_ref/libedsparser_ref.a -ldl -o $@ && rm -f $@.o
