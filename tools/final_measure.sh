# round-end measurement batch on one B200 (every profiled command runs once without the profiler first)
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo bench rc=$?
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_bench_ref.json 2> gpurun_out/r02f_bench_ref.err; echo ref rc=$?
timeout 300 python bench.py --steps 2 --warmup 1 --no-knn25m > gpurun_out/r02f_plain.json 2>/dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-knn25m > gpurun_out/r02f_ncu_launch.log 2>&1; echo launches rc=$?
python tools/e2e_once.py > gpurun_out/r02f_e2e_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_e2e_launches.csv python tools/e2e_once.py > gpurun_out/r02f_e2e_ncu.log 2>&1; echo e2e launches rc=$?
python tools/knn_once.py 2>&1 | tail -3
python tools/timeline.py 2>&1 | grep -E "^step|per-CTA" | tail -4
