set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo bench rc=$?
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_bench_ref.json 2> gpurun_out/r02f_bench_ref.err; echo ref rc=$?
timeout 300 python bench.py --steps 2 --warmup 1 --no-knn25m > gpurun_out/r02f_plain.json 2>/dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-knn25m > gpurun_out/r02f_ncu_launch.log 2>&1; echo launches rc=$?
MRS_NO_ITEM_AVG=1 python tools/prof_pass.py --passes 3 > gpurun_out/r02f_pp.log 2>&1 && MRS_NO_ITEM_AVG=1 ncu --set full --clock-control none --import-source on -k regex:"item_tiled_kernel|predict_mae_tiled|user_sum_kernel" -s 6 -c 3 -o gpurun_out/r02f_pass python tools/prof_pass.py --passes 3 > gpurun_out/r02f_ncu_pass.log 2>&1; echo ncu pass rc=$?
NO_TIMING=1 ncu --set full --clock-control none --import-source on -k regex:"similarity_wide|sort_rank|pers_mae|dev_pre" -s 4 -c 4 -o gpurun_out/r02f_knn python tools/knn_once.py > gpurun_out/r02f_ncu_knn.log 2>&1; echo ncu knn rc=$?
python tools/timeline.py 2>&1 | grep -E "^step|per-CTA" | tail -5
