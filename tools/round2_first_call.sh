#!/usr/bin/env bash
# First GPU call of round 2: everything that was written after round 1's GPU minutes ran out, in the order that
# keeps a failure from hiding the rest.  Every step writes under gpurun_out/ and none stops the script.
#
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
#
set -u
mkdir -p gpurun_out
run() {  # name, timeout seconds, command...
    local name=$1 limit=$2; shift 2
    echo "=== $name" | tee -a gpurun_out/round2_first_call.log
    timeout "$limit" "$@" > "gpurun_out/$name.out" 2> "gpurun_out/$name.err"
    echo "exit $? ($name)" | tee -a gpurun_out/round2_first_call.log
}
run gpu_tests            900 python -m pytest tests -m gpu -x -q
run aligned_rows_probe   300 python tools/probe/aligned_rows_probe.py
run aligned_rows_test    120 env SHRIMPY_TEST_UNMEASURED=1 python -m pytest tests/test_deskew_gpu.py -q -k whole_sector
run paged_stack_one_gpu  180 env SHRIMPY_TEST_UNMEASURED=1 python -m pytest tests/test_paged_stack_gpu.py -q -x
run host_call_probe      300 python tools/probe/host_call_probe.py
run bench                600 python bench.py
tail -n 3 gpurun_out/*.out | cut -c1-400
# then, in a 2-GPU call (gpurun --gpus 2), the scan split with no exchange step (paged_stack.py) beside the measured one:
#   for t in peer vmm; do python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
#       --master-port 29611 tools/scan_split_bench.py --transport $t > gpurun_out/scan_split_$t.json 2> gpurun_out/scan_split_$t.err; done
