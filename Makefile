# Build libmvgeo.so (sm_100a only) in-tree. `python -c "import __graft_entry__ as g; g.build()"` runs this.
NVCC ?= nvcc
PKG  := 2025_icra_multi_view_robot_pose_estimation_b200
SRC  := $(PKG)/csrc/decode.cu $(PKG)/csrc/dlt.cu $(PKG)/csrc/fk.cu $(PKG)/csrc/encode.cu $(PKG)/csrc/pnp.cu $(PKG)/csrc/geom.cu $(PKG)/csrc/api.cu
HDR  := include/mvgeo.h $(PKG)/csrc/common.cuh $(PKG)/csrc/dlt_device.cuh $(PKG)/csrc/fk_device.cuh
OUT  := $(PKG)/libmvgeo.so
NVFLAGS := -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Iinclude \
           -Xptxas -v --expt-relaxed-constexpr

$(OUT): $(SRC) $(HDR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(SRC) 2> build_ptxas.log || (cat build_ptxas.log; exit 1)

clean:
	rm -f $(OUT) build_ptxas.log
