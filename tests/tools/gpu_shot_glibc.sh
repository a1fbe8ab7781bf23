# Round-2 GPU validation of prefs.devices.b200.libm = 'glibc' (recorded command of the last gpurun call)
mkdir -p gpurun_out
( time timeout 150 python -m pytest tests/test_parity_gpu.py -q -x -k "glibc_math or cuba_1000-True or cobahh_1000-False or test_device_math_identical" ) > gpurun_out/r2w_tests.log 2>&1
tail -8 gpurun_out/r2w_tests.log
( time timeout 60 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2w_smoke.log 2>&1
tail -2 gpurun_out/r2w_smoke.log
