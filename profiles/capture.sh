# Evidence capture run on the GPU box: plain runs first (they must exit 0), then the ncu launch list and one --set full
# capture of a whole launch sequence (every kernel of one batch of 256 pairs).  Usage: bash profiles/capture.sh r01e
TAG=${1:-r01f}
mkdir -p gpurun_out
set -x
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || exit 1
python bench.py --steps 2 --warmup 3 > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_${TAG}_a.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 84 -c 28 -o gpurun_out/prof_$TAG -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_${TAG}_b.log 2>&1
ls -la gpurun_out
