# ncu --set full of the descriptor kernel alone (profiles/capture.sh stops before it); run after a plain run exited 0
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 > gpurun_out/plain_r01e.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'k_describe|k_text' --launch-skip 6 -c 4 -o gpurun_out/prof_r01d_describe -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_r01d_c.log 2>&1
ls -la gpurun_out
