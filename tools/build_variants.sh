#!/bin/bash
# A/B build variants of libdiqt_b200 (see DIQT_PDL_LATE_TRIGGER / DIQT_LOAD_NC in csrc/common.cuh) -> build/variants/<name>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
v() { name=$1; shift; DIQT_OUT=$PWD/build/variants/$name.so DIQT_BUILD_DIR=$PWD/build/variants/obj_$name DIQT_EXTRA_FLAGS="$*" bash diffusioniqt_b200/csrc/build.sh 2>&1 | tail -1; }
v cg_early -DDIQT_PDL_LATE_TRIGGER=0 &
v nc_early -DDIQT_LOAD_NC=1 -DDIQT_PDL_LATE_TRIGGER=0 &
v nc_late -DDIQT_LOAD_NC=1 -DDIQT_PDL_LATE_TRIGGER=1 &
wait
