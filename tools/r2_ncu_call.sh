#!/bin/bash
# One gpurun call: an ncu --set full capture of every kernel of the path in every BASELINE shape (tools/ncu_targets.py), each
# summarised ON THE BOX (the reports together exceed what gpurun copies back; only those named in NCU_KEEP travel), then the
# launch list of a short bench run. Outputs under gpurun_out/.
mkdir -p gpurun_out
for s in ${NCU_SCENARIOS:-c5_64m c5_8m c5_1m c5_64m_slot c5_init c3_16m c3_1m c2 c4 churn churn_slot interop}; do
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_ncu_$s \
      python tools/ncu_targets.py $s > gpurun_out/r2_ncu_$s.log 2>&1
  echo "$s rc=$?" >> gpurun_out/r2_ncu_rc.txt
  python tools/ncu_summary.py gpurun_out/r2_ncu_$s.ncu-rep gpurun_out/r2_ncu_${s}_summary.txt > /dev/null 2>&1
  case " ${NCU_KEEP:-c3_16m c5_8m} " in *" $s "*) ;; *) rm -f gpurun_out/r2_ncu_$s.ncu-rep ;; esac
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 5 --warmup 3 > gpurun_out/r2_launches_bench.log 2>&1
cat gpurun_out/r2_ncu_rc.txt; du -sh gpurun_out
