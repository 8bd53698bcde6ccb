#!/bin/bash
# What the chunked (multi-GPU) walk of the evaluation kernel costs on one GPU: grid stride against
# chunks of two sizes, without and with copies of the chunks' gradient ranges to local stand-in
# buffers (scripts/kbench.cu, build with scripts/kbench_build.sh cur "").
cd /root/repo
./build/kbench/kb_cur 13682 4456117 28987644 1 1 plain | grep KBENCH
for c in 224 992; do KBENCH_CHUNK=$c ./build/kbench/kb_cur 13682 4456117 28987644 1 1 chunk$c | grep "KBENCH"; for pe in 1 7; do KBENCH_PEERS=$pe KBENCH_CHUNK=$c ./build/kbench/kb_cur 13682 4456117 28987644 1 1 chunk${c}_peers$pe | grep "KBENCH"; done; done
