#!/bin/bash
# Evidence of one build on a 1-GPU box (tools/evidence.sh outdir): GPU tests, bench lines, launch list, ncu --set full
# of the step's kernels (dim 1280) and of the dim-2560 kernels, cuFFT strawman, parity sweep.
out=$1; mkdir -p $out
python -m pytest tests -m gpu -q > $out/tests.log 2>&1; echo "tests rc=$?"; tail -1 $out/tests.log
python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err; echo "reference rc=$?"
python bench.py --config 5 --steps 3 --warmup 2 > $out/bench_config5.json 2> $out/bench_config5.err; echo "config5 rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu --no-configs --row-kernel 1 > $out/bench_warp_kernel.json 2>/dev/null; echo "warp-kernel rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
    python bench.py --draws 256 --steps 1 --warmup 1 --no-cpu --no-configs --no-fp64-leg > $out/ncu_launches.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"group_rows|hot_cols|fit_kernel|fft_convolve|psd_quad|tiled_pass|ao_zone|resample" \
    -s 18 -c 9 -o $out/prof python bench.py --draws 256 --steps 1 --warmup 1 --no-cpu --no-configs --no-fp64-leg > $out/ncu_full.log 2>&1; echo "ncu full rc=$?"
PSFR_BENCH_BATCH_ONLY=1 ncu --set full --clock-control none --import-source on -k regex:"group_rows|hot_cols|pass_kernel" \
    -s 19 -c 5 -o $out/prof_2560 python bench.py --config 5 --steps 1 --warmup 1 --no-cpu --draws5 16 > $out/ncu_2560.log 2>&1; echo "ncu 2560 rc=$?"
python tools/cufft_check.py --out $out/cufft_strawman.json > $out/cufft.log 2>&1; echo "cufft rc=$?"
python tools/parity_sweep.py --out $out/parity_sweep.json > $out/sweep.log 2>&1; echo "sweep rc=$?"
