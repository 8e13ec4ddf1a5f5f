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

all: $(LIB)

$(OBJ)/%.o: $(SRC)/%.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJ)/$*.ptxas.log || (cat $(OBJ)/$*.ptxas.log; exit 1)

$(LIB): $(OBJS)
	@mkdir -p cutrace_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -cudart shared

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf build $(LIB)
.PHONY: all oracle clean
