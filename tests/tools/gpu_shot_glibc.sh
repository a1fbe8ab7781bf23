# Round-2 GPU validation of prefs.devices.b200.libm = 'glibc' (recorded command of the gpurun calls)
mkdir -p gpurun_out
( time timeout 60 python bench.py --workload cobahh_256k --libm glibc --phases --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2z_phases_glibc.json 2> gpurun_out/r2z_phases_glibc.err
grep PHASE gpurun_out/r2z_phases_glibc.err | head -12
( time timeout 240 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2z_bench_default.json 2> gpurun_out/r2z_bench_default.err
tail -3 gpurun_out/r2z_bench_default.err
python - <<'P'
import json
line = [l for l in open("gpurun_out/r2z_bench_default.json") if l.startswith("{")][-1]
d = json.loads(line)
print(json.dumps({k: d.get(k) for k in ("value", "us_per_timestep", "parity_check", "cpu_baseline")})[:3000])
print([(c["config"]["workload"][:30], c.get("value"), c.get("error")) for c in d.get("configs", [])])
P
