mkdir -p gpurun_out
timeout 300 python tests/sanitize_smoke.py 2>&1 | tail -2 | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck python tests/sanitize_smoke.py > gpurun_out/memcheck_r02f.log 2>&1; grep -E "ERROR SUMMARY|sanitize smoke" gpurun_out/memcheck_r02f.log | cut -c1-160
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python tests/sanitize_smoke.py > gpurun_out/racecheck_r02f.log 2>&1; grep -E "RACECHECK SUMMARY|sanitize smoke" gpurun_out/racecheck_r02f.log | cut -c1-160
timeout 600 compute-sanitizer --tool synccheck python tests/sanitize_smoke.py > gpurun_out/synccheck_r02f.log 2>&1; grep -E "ERROR SUMMARY|sanitize smoke" gpurun_out/synccheck_r02f.log | cut -c1-160
