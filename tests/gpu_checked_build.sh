#!/bin/bash
# The GPU suite against the CHECKED build of the library (python -m dge_b200.build --variant checks -DDGE_CHECKS=1:
# bounds assertions on the scattered stores of sort / partition / expand and on the gathered ids of the blend
# staging, common.cuh:DGE_CHECK). A failed assertion prints its site and traps, which fails the test it happens in.
mkdir -p gpurun_out
LIB=dge_b200/_build/var_checks/libdge_b200.so
[ -f $LIB ] || { echo "build the variant first: python -m dge_b200.build --variant checks -DDGE_CHECKS=1"; exit 1; }
DGE_B200_LIB=$PWD/$LIB timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/checked_suite.log
DGE_B200_LIB=$PWD/$LIB timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a gpurun_out/checked_suite.log
grep -c "DGE_CHECK failed" gpurun_out/checked_suite.log
