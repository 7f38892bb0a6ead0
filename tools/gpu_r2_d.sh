#!/bin/bash
# round 2, call D: stage-A grid cap experiment (two streams on disjoint SMs) + split-B
for cap in 0 74 96 112; do echo "== stage A grid cap $cap"; AIRPOSE_STAGEA_GRID=$cap timeout 300 python tools/gpu_probe.py trunk 2>&1 | grep "n=128"; done
echo "== split B"; AIRPOSE_TRUNK_SPLIT_B=1 timeout 300 python tools/gpu_probe.py trunk 2>&1 | grep "n=128"
echo "== split B cap 74"; AIRPOSE_TRUNK_SPLIT_B=1 AIRPOSE_STAGEA_GRID=74 timeout 300 python tools/gpu_probe.py trunk 2>&1 | grep "n=128"
echo "== chunk 32, cap 74"; AIRPOSE_TRUNK_CHUNK=32 AIRPOSE_STAGEA_GRID=74 timeout 300 python tools/gpu_probe.py trunk 2>&1 | grep "n=128"
