#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
  timeout 60 python tools/tc_probe.py selftest $v > gpurun_out/r2c_selftest_$v.json 2> gpurun_out/r2c_selftest_$v.err; echo "selftest $v rc=$?"; cat gpurun_out/r2c_selftest_$v.json | cut -c1-2500; tail -2 gpurun_out/r2c_selftest_$v.err
done
