#!/bin/bash
# debug helper: run the 2-rank worker with each rank under compute-sanitizer memcheck
mode=${1:-gpu}; n=${2:-30000}; f=${3:-96}; leaves=${4:-31}
export MASTER_ADDR=127.0.0.1 MASTER_PORT=29533 WORLD_SIZE=2 NO_GRAPH=1
for r in 0 1; do
  RANK=$r LOCAL_RANK=$r timeout 500 compute-sanitizer --tool memcheck --print-limit 8 python tests/multi_worker.py $mode $n $f $leaves > gpurun_out/san_$r.log 2>&1 &
done
wait
for r in 0 1; do echo "=== rank $r"; grep -v "^JSON" gpurun_out/san_$r.log | grep -E "Invalid|at |by thread|Address|ERROR SUMMARY|Error|error" | head -30; done
