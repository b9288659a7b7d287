#!/bin/bash
# Run every GPU parity test in its own process (a faulting kernel must not take the other tests down),
# with a per-test timeout, collecting logs under gpurun_out/.
mkdir -p gpurun_out/tests
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
TESTS=${TESTS:-$(python - <<'PY'
import re
src = open("tests/test_gpu_parity.py").read()
print(" ".join(re.findall(r"^def (test_\w+)", src, flags=re.M)))
PY
)}
: > gpurun_out/summary.txt
for t in $TESTS; do
  timeout ${PER_TEST_TIMEOUT:-420} python -m pytest "tests/test_gpu_parity.py" -k "$t" -m gpu -q -s -x -p no:cacheprovider \
      > gpurun_out/tests/$t.log 2>&1
  rc=$?
  echo "$t rc=$rc $(grep -E 'passed|failed|error' gpurun_out/tests/$t.log | tail -1)" >> gpurun_out/summary.txt
done
grep -h "\[parity\]" gpurun_out/tests/*.log > gpurun_out/parity.txt 2>/dev/null
cat gpurun_out/summary.txt
