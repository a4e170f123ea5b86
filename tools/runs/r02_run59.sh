#!/bin/bash
# round-2 GPU call 59: randomised comparison of the host pipeline with the device run (tools/fuzz_host.py)
mkdir -p gpurun_out
timeout 200 python tools/fuzz_host.py 45 1 > gpurun_out/r02_fuzz_host_seed1.json 2> gpurun_out/r02_fuzz_host.err
tail -3 gpurun_out/r02_fuzz_host.err
cat gpurun_out/r02_fuzz_host_seed1.json
