#!/bin/bash
mkdir -p gpurun_out
{
PROBE_SCHEDS=pixel timeout 300 python tools/r02_probe.py triangle spheres1080 mirror1080 bunny4k
PROBE_SCHEDS=pixel,launches,frame timeout 300 python tools/r02_probe.py synthetic10m
} > gpurun_out/r02_probe12.log 2>&1
cat gpurun_out/r02_probe12.log
