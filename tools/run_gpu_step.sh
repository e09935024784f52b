set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/perf_probe.py > gpurun_out/r2_probe_t256.txt 2>&1; head -3 gpurun_out/r2_probe_t256.txt
ACIDS_B200_LIB=$PWD/acids_transforms_b200/variants/libacids_t128.so python tools/perf_probe.py > gpurun_out/r2_probe_t128.txt 2>&1; head -3 gpurun_out/r2_probe_t128.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -c 3000 gpurun_out/r2_bench_a.json; tail -5 gpurun_out/r2_bench_a.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_a.json 2> gpurun_out/r2_bench_ref_a.err; tail -c 1500 gpurun_out/r2_bench_ref_a.json
