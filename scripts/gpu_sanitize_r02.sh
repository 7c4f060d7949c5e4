#!/bin/bash
# compute-sanitizer (memcheck, racecheck) over the kernels added or changed late in round 2: DQN kernels, the fused kernel with 8-column data-gradient items
O=gpurun_out; T=r02p
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_dqn.py tests/test_gpu_brain.py -m gpu -q -k "select_actions or replay_write or (test_fused_kernel_matches_layered_kernels and 20-2-7-False)" > $O/memcheck_$T.log 2>&1; echo "memcheck rc=$?"; tail -4 $O/memcheck_$T.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_dqn.py tests/test_gpu_brain.py -m gpu -q -k "replay_write or (test_fused_kernel_matches_layered_kernels and 20-2-7-False)" > $O/racecheck_$T.log 2>&1; echo "racecheck rc=$?"; tail -4 $O/racecheck_$T.log
