# Round-2 GPU validation of prefs.devices.b200.libm = 'glibc' (recorded command of the last gpurun call)
mkdir -p gpurun_out
( time timeout 80 python -m pytest tests/test_parity_gpu.py -q -x -k "every_libm_call or cuba_1000-True" ) > gpurun_out/r2v_tests.log 2>&1
tail -8 gpurun_out/r2v_tests.log
