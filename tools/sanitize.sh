#!/bin/bash
# compute-sanitizer over every kernel of the hot path: one small proof of each circuit
# family + the commitment / Merkle edge cases.  ONE tool per gpurun call (B200_PROFILING.md):
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh memcheck'      then, in another call,
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh racecheck'
# Writes gpurun_out/sanitize_<tool>.log; the summary line of each run is kept under profiles/.
set -o pipefail
tool=${1:-memcheck}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# the same test selection first WITHOUT the sanitizer: it must be green before a tool is attached
SEL='tiny_circuit or aes_gcm_tag_small or feistel or public_inputs or merkle_fewer or from_values_small or device_witness or hash_and_merkle'
python -m pytest tests/test_gpu_prove.py tests/test_gpu_commit.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitize_plain.log 2>&1 || { tail -5 gpurun_out/sanitize_plain.log; echo "plain run failed: not sanitizing"; exit 1; }
timeout 1300 compute-sanitizer --tool "$tool" --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_prove.py tests/test_gpu_commit.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitize_$tool.log 2>&1
rc=$?
tail -6 gpurun_out/sanitize_$tool.log
echo "compute-sanitizer --tool $tool exit code $rc"
exit $rc
