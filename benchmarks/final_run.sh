set -u
O=gpurun_out/final
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/gpu_tests.log 2>&1; echo "tests rc=$?" >> $O/gpu_tests.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/launches.log 2>&1; echo "launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"concat_fwd_vec4|soft_argmin_fwd|patch_loss_fold_v3" -s 9 -c 3 -o $O/step_kernels -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-variant --no-stock-variant > $O/ncu_step.log 2>&1; echo "ncu step rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gwc_bwd_systolic -s 3 -c 1 -o $O/gwc_bwd -f python benchmarks/variant_bench.py --only gwc > $O/ncu_gwc.log 2>&1; echo "ncu gwc rc=$?"
timeout 900 python benchmarks/kernel_sweep.py --out $O/kernel_sweep.json > $O/kernel_sweep.log 2>&1; echo "sweep rc=$?"
timeout 600 python benchmarks/variant_bench.py --out $O/variants_all.json > $O/variants_all.log 2>&1; echo "variants rc=$?"
tail -3 $O/gpu_tests.log
