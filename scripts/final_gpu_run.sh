#!/bin/bash
# One GPU box, one GPU: the evidence set of a round (tests, smoke, every config's bench line, the reference arm,
# the ncu launch list and the full capture of the dominant kernels).  Outputs under gpurun_out/final/.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/final_gpu_run.sh'
set -u
O=gpurun_out/final
mkdir -p $O
python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_c5_reference.json 2> $O/bench_c5_reference.err; echo "ref rc=$?"
python bench.py > $O/bench_c5_n1.json 2> $O/bench_c5_n1.err; echo "c5 rc=$?"
for c in c1 c2; do python bench.py --config $c > $O/bench_$c.json 2> $O/bench_$c.err; echo "$c rc=$?"; done
for c in c3 c4; do python bench.py --config $c --no-cpu-baseline > $O/bench_$c.json 2> $O/bench_$c.err; echo "$c rc=$?"; done
VQGNN_CUPROF=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $O/launches_c5_step.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graphs --sync-vq > $O/ncu_launches.log 2>&1; echo "launches rc=$?"
VQGNN_CUPROF=1 ncu --profile-from-start off --set full --clock-control none --import-source on \
  --kernel-name regex:"mp_fwd_rows_kernel|vq_assign_tc_kernel|segsum_kernel|tail_materialize_kernel" --launch-count 14 \
  -o $O/prof_c5_step -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graphs --sync-vq > $O/ncu_full.log 2>&1; echo "ncu rc=$?"
for f in $O/bench_*.json; do python scripts/show_bench.py $f 2>/dev/null | head -3; done
