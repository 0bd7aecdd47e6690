#!/bin/bash
# One gpurun call: smoke, the bench command, its ncu launch list, one --set full capture of the headline kernel,
# compute-sanitizer on small configurations.  Outputs under gpurun_out/ (copied into profiles/ by hand).
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/r2_box.txt
(timeout 600 python __graft_entry__.py smoke > $O/r2_smoke.log 2>&1; echo "smoke exit $?")
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also"
(timeout 600 $CMD > $O/r2_prof_bench.json 2> $O/r2_prof_bench.err; echo "bench exit $?")
(timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv \
   --log-file $O/r2_launches_bench_steps2.csv $CMD > $O/r2_ncu_list.log 2>&1; echo "ncu list exit $?")
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep_tc16 -s 40 -c 1 -f -o $O/r2_fused16 \
   $CMD > $O/r2_ncu_full.log 2>&1; echo "ncu full exit $?")
# compute-sanitizer: one attempt, kept for the record (this pool answers that the tool is closed; see profiles/README.md)
(timeout 300 compute-sanitizer --tool memcheck --log-file $O/r2_sanitizer_memcheck.log python scripts/sanitize_small.py all > $O/r2_sanitize_all.out 2>&1; echo "memcheck exit $?")
# the same small configurations without the tool: every kernel family runs and the library's own checks pass
(timeout 300 python scripts/sanitize_small.py all > $O/r2_small_all.out 2>&1; echo "small configurations exit $?")
tail -3 $O/r2_smoke.log; tail -c 300 $O/r2_prof_bench.json; tail -3 $O/r2_sanitize_all.out; tail -8 $O/r2_small_all.out
