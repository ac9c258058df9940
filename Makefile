# Build libmvgeo.so (sm_100a only) in-tree. `python -c "import __graft_entry__ as g; g.build()"` runs `make -j`.
NVCC ?= nvcc
PKG  := 2025_icra_multi_view_robot_pose_estimation_b200
SRCS := decode dlt fk encode pnp geom api
OBJ  := $(addprefix build/,$(addsuffix .o,$(SRCS)))
HDR  := include/mvgeo.h $(PKG)/csrc/common.cuh $(PKG)/csrc/dlt_device.cuh $(PKG)/csrc/fk_device.cuh
OUT  := $(PKG)/libmvgeo.so
NVFLAGS := -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Iinclude \
           -Xptxas -v --expt-relaxed-constexpr

$(OUT): $(OBJ)
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -shared -o $@ $(OBJ)
	cat $(OBJ:.o=.ptxas.log) > build_ptxas.log

build/%.o: $(PKG)/csrc/%.cu $(HDR)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c -o $@ $< 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

clean:
	rm -f $(OUT) build_ptxas.log build/*.o build/*.ptxas.log
