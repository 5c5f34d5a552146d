mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/c15_tests.log 2>&1
tail -3 gpurun_out/c15_tests.log
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_smoke.py > gpurun_out/c15_san_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize smoke ok|Error|error" gpurun_out/c15_san_$tool.log | head -8
done
