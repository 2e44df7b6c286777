# Evidence capture of round 2 on the GPU box (one GPU): plain runs first (they must exit 0), then the ncu launch list and ONE
# --set full capture of every kernel of the path (profiles/prof_step.py: one launch sequence of 256 pairs + bag of words +
# text + MIP).  Usage: bash profiles/capture_r02.sh r02z
TAG=${1:-r02}
mkdir -p gpurun_out
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || exit 1
python bench.py --steps 700 --warmup 5 > gpurun_out/bench_${TAG}_sustained.json 2> gpurun_out/bench_${TAG}_sustained.err || exit 1
python bench.py --config ucf --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_ucf.json 2> gpurun_out/bench_${TAG}_ucf.err || exit 1
python bench.py --config kth --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_kth.json 2> gpurun_out/bench_${TAG}_kth.err || exit 1
python profiles/prof_step.py 3 > gpurun_out/plain_$TAG.log 2>&1 || exit 1
N=$(grep -o "launches per iteration [0-9]*" gpurun_out/plain_$TAG.log | tail -1 | grep -o "[0-9]*$")
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python profiles/prof_step.py 3 > gpurun_out/ncu_${TAG}_a.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip $((2 * N)) -c $N -o gpurun_out/prof_$TAG -f python profiles/prof_step.py 3 > gpurun_out/ncu_${TAG}_b.log 2>&1
ls -la gpurun_out | tail -5
