#!/bin/bash
mkdir -p gpurun_out
DSIM_INPUTS_READY=1 DSIM_LIB=$PWD/mujoco_drone_b200/variants/tl.so timeout 300 python tools/timeline_graph.py c4 > gpurun_out/r2d_timeline_c4_ready.log 2>&1
head -34 gpurun_out/r2d_timeline_c4_ready.log
DSIM_INPUTS_READY=1 DSIM_LIB=$PWD/mujoco_drone_b200/variants/tl.so timeout 300 python tools/timeline_graph.py c4 524288 > gpurun_out/r2d_timeline_c4x4_ready.log 2>&1
head -22 gpurun_out/r2d_timeline_c4x4_ready.log
