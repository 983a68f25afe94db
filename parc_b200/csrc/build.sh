#!/usr/bin/env bash
# Builds parc_b200/libparc_b200.so for sm_100a (B200).  Cross-compiles without a GPU.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${here}/../libparc_b200.so"
NVCC="${NVCC:-nvcc}"
COMMON="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
mkdir -p "${here}/build"
# FMA contraction stays enabled: every rounding that decides an index or a branch uses explicit *_rn
# intrinsics (parc_common.cuh), which never contract.
for f in motion_query fk heightfield api dataset_sweep tracker_step table_build motion_opt peer_gather; do
  "${NVCC}" ${COMMON} ${EXTRA_NVCC_FLAGS:-} -c "${here}/${f}.cu" -o "${here}/build/${f}.o" &
done
"${NVCC}" ${COMMON} ${EXTRA_NVCC_FLAGS:-} -c "${here}/body_loss.cu" -o "${here}/build/body_loss.o" &
wait
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -shared -o "${out}" "${here}"/build/{motion_query,fk,heightfield,api,body_loss,dataset_sweep,tracker_step,table_build,motion_opt,peer_gather}.o
echo "built ${out}"
