# round 2, job 14 (1 GPU): parity after the capacity-miss fix (full log kept), host-side profile of one step
timeout 900 python -X faulthandler -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference_cuda.py -m gpu -q -x --tb=short -v > gpurun_out/r02_pytest_parity.log 2>&1
grep -n "PASSED\|FAILED\|ERROR" gpurun_out/r02_pytest_parity.log | tail -4; grep -n "passed\|failed" gpurun_out/r02_pytest_parity.log | tail -2; grep -n "Fatal\|illegal\|File \"" gpurun_out/r02_pytest_parity.log | head -12
timeout 300 python tools/host_profile.py --steps 300 > gpurun_out/r02_host_profile.txt 2>&1; head -45 gpurun_out/r02_host_profile.txt | cut -c1-160
