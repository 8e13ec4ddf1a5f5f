# Builds the product library (CUDA, sm_100a only) and the test-side oracle.
#   make            -> cutrace_b200/lib/libcutrace_b200.so  (C-ABI of include/cutrace.h)
#   make oracle     -> oracle/libcutrace_oracle.so (+ oracle/_ref/*.so when the reference tree exists)
#   make cli        -> bin/cutrace (C++ host: JSON/STL loader, JPEG writer, main)
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -std=c++17 -O3 $(ARCH) -lineinfo -Xcompiler -fPIC -Xptxas -v
SRC := cutrace_b200/csrc
OBJ := build/obj
LIB := cutrace_b200/lib/libcutrace_b200.so
CU := $(SRC)/api.cu $(SRC)/bvh_build.cu $(SRC)/render.cu $(SRC)/output.cu
OBJS := $(patsubst $(SRC)/%.cu,$(OBJ)/%.o,$(CU))
HDRS := $(wildcard $(SRC)/*.cuh) include/cutrace.h

all: $(LIB) host cli
lib: $(LIB)

$(OBJ)/%.o: $(SRC)/%.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJ)/$*.ptxas.log || (cat $(OBJ)/$*.ptxas.log; exit 1)

$(LIB): $(OBJS)
	@mkdir -p cutrace_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -cudart shared

HOST := cutrace_b200/host
HOSTLIB := cutrace_b200/lib/libcutrace_host.so
HOSTCXX := $(if $(wildcard /usr/bin/g++),/usr/bin/g++,g++)
HOSTFLAGS := -std=c++17 -O2 -ffp-contract=off -fPIC -Wall -pthread

host: $(HOSTLIB)
$(HOSTLIB): $(HOST)/scene_loader.cpp $(HOST)/jpeg.cpp $(HOST)/host_api.cpp $(wildcard $(HOST)/*.hpp) include/cutrace.h include/cutrace_host.h
	@mkdir -p cutrace_b200/lib
	$(HOSTCXX) $(HOSTFLAGS) -shared -o $@ $(HOST)/scene_loader.cpp $(HOST)/jpeg.cpp $(HOST)/host_api.cpp

cli: bin/cutrace
bin/cutrace: $(HOST)/main.cpp $(HOST)/scene_loader.cpp $(HOST)/jpeg.cpp $(wildcard $(HOST)/*.hpp) $(LIB)
	@mkdir -p bin
	$(HOSTCXX) $(HOSTFLAGS) -o $@ $(HOST)/main.cpp $(HOST)/scene_loader.cpp $(HOST)/jpeg.cpp -Lcutrace_b200/lib -lcutrace_b200 \
	  -Wl,-rpath,'$$ORIGIN/../cutrace_b200/lib' -L/usr/local/cuda/lib64 -Wl,-rpath,/usr/local/cuda/lib64

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf build bin $(LIB) $(HOSTLIB)
.PHONY: all lib oracle clean host cli
