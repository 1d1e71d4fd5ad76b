#!/bin/bash
# Probe build of the library (-DTPLS_PROBE): the streaming kernels take experiment switches from TPLS_DBG
# (rowpass.cu / stream_common.cuh).  Output: cmtf_pls_b200/libtpls_b200_probe.so, loaded with TPLS_B200_LIB=...
set -e
cd "$(dirname "$0")/../cmtf_pls_b200/csrc"
O=/tmp/tpls_probe_obj; mkdir -p $O
ARCH="-gencode arch=compute_100a,code=sm_100a"
rm -f $O/*.o
pids=""
for f in passes rowpass covpass multiproj rank1 small xchg driver transform ops; do
  nvcc -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH ${PROBE_FLAG--DTPLS_PROBE} ${PROBE_DEFS:-} -c $f.cu -o $O/$f.o &
  pids="$pids $!"
done
for p in $pids; do wait $p; done
nvcc -shared $ARCH -o ../libtpls_b200_probe${PROBE_TAG:-}.so $O/*.o -ldl -lcudart
ls -la ../libtpls_b200_probe${PROBE_TAG:-}.so
