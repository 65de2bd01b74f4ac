#!/bin/bash
# A/B of differently built libraries (build_variants/*.so, made on the build host) on one GPU box.
LIB=lunar_module_ascent_trajectory_optimiser_b200/liblmato_b200.so
cp $LIB /tmp/lib_keep.so
for v in build_variants/*.so; do
  echo "== $v"
  cp $v $LIB
  timeout 300 python ${PROBE:-tools/gpu_variant_probe.py} 2>&1 | tail -6
done
cp /tmp/lib_keep.so $LIB
