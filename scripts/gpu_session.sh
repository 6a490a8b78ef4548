#!/usr/bin/env bash
# gpu_session.sh — one GPU call (1 GPU) that produces everything the 1-GPU evidence under profiles/ comes from:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_session.sh'
# Stages (each bounded by its own timeout, outputs under gpurun_out/; a failing stage does not stop the next):
#   tests     python -m pytest tests -m gpu
#   knobs     tools/ozaki_knobs.py: parity + timing of the tcgen05 kernel and of its diagnostics flags (cost of loads / C update / no wave sync)
#   peaks     bin/umma_rate: burst table with random operands, sustained int8 peak (zero and random operands)
#   ncu       launch list of a short bench run + one --set full capture of the tcgen05 GEMM kernel at the bench launch shape
#   sanitize  compute-sanitizer memcheck / racecheck / synccheck of the tcgen05 parity tests on small shapes
#   bench     python bench.py (N=32768, value + band-pipelined e2e + roofline + cpu_baseline)
# Choose stages with STAGES="tests knobs ..." (default: all).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
STAGES="${STAGES:-tests knobs peaks ncu sanitize bench}"
run() { local name=$1 limit=$2; shift 2; echo "=== $name" >&2; timeout "$limit" "$@"; echo "=== $name exit $?" >&2; }  # markers on stderr: stdout may be a data file
for s in $STAGES; do
  case $s in
    tests)    run tests 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log ;;
    knobs)    run knobs 900 python tools/ozaki_knobs.py --out gpurun_out/ozaki_knobs.jsonl --time 8192 16384 32768 --tstamp-n 16384 ;;
    peaks)    run umma_random 120 env UMMA_RANDOM=1 bin/umma_rate > gpurun_out/umma_rate_random.jsonl 2> gpurun_out/umma_rate.err
              run umma_sustain 120 env UMMA_SUSTAIN=1 bin/umma_rate > gpurun_out/umma_rate_sustained.jsonl 2>> gpurun_out/umma_rate.err
              tail -2 gpurun_out/umma_rate_sustained.jsonl ;;
    ncu)      run ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
                  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-refcuda --no-secondary > gpurun_out/ncu_bench.log 2>&1
              run ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:ozaki_gemm -s 1 -c 1 -o gpurun_out/prof_ozaki_bench_shape -f \
                  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-refcuda --no-secondary > gpurun_out/ncu_full.log 2>&1
              # back in the build container: python tools/ncu_summary.py gpurun_out/prof_ozaki_bench_shape.ncu-rep \
              #     --traffic-json profiles/ozaki_traffic_rNN.json --shape 32768,8192,32768 > profiles/ncu_ozaki_bench_shape_rNN.txt
              ;;
    sanitize) for tool in memcheck racecheck synccheck; do
                run sanitize_$tool 600 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
                    -k "ozaki_gemm_seeded_vs_oracle or ozaki_gemm_index_fill or special_values or different_streams" > gpurun_out/sanitizer_$tool.log 2>&1
                tail -4 gpurun_out/sanitizer_$tool.log
              done ;;
    bench)    run bench 1500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench_n1.json ;;
  esac
done
