#!/bin/bash
# usage: scripts/ptxas_regs.sh [extra nvcc flags]  -> registers / spills of every polling kernel
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xptxas -v "$@" -c ground-plane-polling_b200/csrc/gpp_launch.cu -o /tmp/ptxas_regs.o 2>&1 | \
 awk '/Compiling entry function/{name=$0; sub(/.*function ./,"",name); sub(/. for.*/,"",name)} /spill/{sp=$0} /Used/{print name " | " $0 " | " sp}' | grep -E "poll" | sed 's/ptxas info    : //; s/_ZN3gpp//; s/EvNS_.*E |/ |/'
