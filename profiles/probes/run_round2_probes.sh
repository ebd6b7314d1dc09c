#!/bin/bash
# First GPU call of the next round (about one minute of box time):
#   /usr/local/graft/bin/gpurun --timeout 300 -- 'bash profiles/probes/run_round2_probes.sh'
# Builds the two probes written at the end of round 1 on the box and leaves their output under gpurun_out/.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out scratch
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc -O3 -std=c++17 $ARCH -o scratch/epilogue_store_probe profiles/probes/epilogue_store_probe.cu -lcuda || exit 1
nvcc -O3 -std=c++17 $ARCH -o scratch/pdl_chain_probe profiles/probes/pdl_chain_probe.cu || exit 1
timeout 120 scratch/epilogue_store_probe | tee gpurun_out/r2_epilogue_store_probe.txt
timeout 60 scratch/pdl_chain_probe | tee gpurun_out/r2_pdl_chain_probe.txt
