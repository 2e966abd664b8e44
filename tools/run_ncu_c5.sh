#!/bin/bash
# Config 5: ncu --set full per kernel at the two ends of the point-count sweep (K = 64; SIMT default and BF16 tensor-core path),
# plus the FPS phase table.  Reports are summarised on the box and deleted (gpurun_out/ <= 64 MiB).
set -u
O=gpurun_out
export SN2_NCU_ONE=1
python tools/prof_fps_plain.py > $O/r2_fps_phases.txt 2>&1
summarise() {
  python tools/summarize_ncu.py $O/$1.ncu-rep > $O/$1.md 2> $O/$1.err
  rm -f $O/$1.ncu-rep
}
KR='regex:fps_bucket_kernel|fps_kernel|sa_fused_kernel|sa1_tc_kernel|knn3_grid|fp1_head|fp2_kernel|sa_pre'
for cfg in "4096 64 64 0" "65536 8 64 0" "16384 64 64 3" "65536 8 2000 3"; do
  tag=$(echo $cfg | tr ' ' '_')
  A="python tools/ncu_target.py $cfg"
  $A > $O/ncu_plain_c5.log 2>&1 && ncu --set full --clock-control none -k "$KR" -c 16 -f -o $O/r2_ncu_config5_$tag $A > $O/ncu_c5_$tag.log 2>&1
  summarise r2_ncu_config5_$tag
done
ls -la $O | tail -12; du -sh $O
