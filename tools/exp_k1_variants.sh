#!/bin/bash
# Builds K1 variants (cape_cell_fit.cu with -D switches) into tools/_variants/<name>.so (run here, nvcc cross-compiles),
# and, on the GPU box (`tools/exp_k1_variants.sh run`), times each with tools/k1_time.py and checks CAPE parity.
# usage: tools/exp_k1_variants.sh build "name:-DFOO=1 -DBAR=2" ...   |   tools/exp_k1_variants.sh run [pytest]
set -e
cd "$(dirname "$0")/.."
P=rgb-d-slam_b200
if [ "$1" = build ]; then
    shift
    python $P/build.py > /dev/null
    for spec in "$@"; do
        name=${spec%%:*}; flags=${spec#*:}
        nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
            -fmad=false $flags -c $P/csrc/cape_cell_fit.cu -o tools/_variants/$name.o
        objs=$(ls $P/build/*.o | grep -v cape_cell_fit.o)
        nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/_variants/$name.so tools/_variants/$name.o $objs -cudart static
        rm tools/_variants/$name.o
        echo built $name
    done
else
    cp $P/librgbdslam_b200.so /tmp/orig.so
    for so in tools/_variants/*.so; do
        cp $so $P/librgbdslam_b200.so
        echo "== $(basename $so .so): $(python tools/k1_time.py 2>&1 | tail -1)"
        if [ "$2" = pytest ]; then python -m pytest tests/test_cape_gpu.py -x -q 2>&1 | tail -2; fi
    done
    cp /tmp/orig.so $P/librgbdslam_b200.so
fi
