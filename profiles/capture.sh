# Evidence capture run on the GPU box: plain runs first, then the ncu launch list and one --set full capture.
mkdir -p gpurun_out
set -x
python bench.py > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err || exit 1
python bench.py --steps 2 --warmup 3 > gpurun_out/plain_r01d.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_r01d_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_describe|k_scores|k_phase_b|k_pa_|k_integral|k_fill|k_twothird|k_half' --launch-skip 60 -c 14 -o gpurun_out/prof_r01d -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_r01d_b.log 2>&1
ls -la gpurun_out
