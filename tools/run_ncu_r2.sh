#!/bin/bash
# Round-2 ncu evidence (one gpurun call; each ncu command runs only after the same command exited 0 without ncu).
# Reports are summarised ON the box (tools/summarize_ncu.py + the raw CSV page, gzipped) and the .ncu-rep files deleted:
# gpurun_out/ may not exceed 64 MiB.
set -u
O=gpurun_out
export SN2_NCU_ONE=1
A="python tools/ncu_target.py 16384 64 2000 0"      # config 2, default (SIMT) path, one forward
B2="python tools/ncu_target.py 16384 64 2000 2"     # same, SA1 second layer on tcgen05 (TF32)
T="python tools/ncu_target_train.py 32 10000"       # one config-3 step (eager)
P="python tools/ncu_target_parcel.py 200"           # parcel tiling on a 200 m x 200 m cloud (1.3 M points, 256 plots)
summarise() {  # $1 = report stem, $2 = optional traffic json
  python tools/summarize_ncu.py $O/$1.ncu-rep ${2:-} > $O/$1.md 2> $O/$1.err
  ncu -i $O/$1.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > $O/$1.raw.csv.gz
  rm -f $O/$1.ncu-rep
}
$A > $O/ncu_plain_a.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_cfg2.csv $A > /dev/null 2>&1
$A > $O/ncu_plain_a.log 2>&1 && ncu --set full --clock-control none --import-source on \
   -k regex:"fps_bucket_kernel|fps_kernel|sa_fused_kernel|knn3_grid|fp1_head|fp2_kernel|fp3_kernel|global_sa|sa_pre|project_|grid_build|ingest" \
   -c 30 -f -o $O/r2_full_cfg2 $A > $O/ncu_full_a.log 2>&1
summarise r2_full_cfg2 $O/r2_ncu_traffic_cfg2.json
$B2 > $O/ncu_plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sa1_tc_kernel -c 2 -f -o $O/r2_full_sa1tc $B2 > $O/ncu_full_b.log 2>&1
summarise r2_full_sa1tc
$T > $O/ncu_plain_t.log 2>&1 && ncu --set full --clock-control none --import-source on \
   -k regex:"lrb_bwd_kernel|lrb_fwd_kernel|lrb_bwd_reduce|lrb_small|edge_msg_fwd|head_|segment_max_fwd|adam_|pointwise|kde_lut|bn_finalize|sa1t_|sa2t_|sa_t_" \
   -c 60 -f -o $O/r2_full_train $T > $O/ncu_full_t.log 2>&1
summarise r2_full_train
$P > $O/ncu_plain_p.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"extract_plots|pgrid|finalize_bands|hardveg" -c 7 -f -o $O/r2_full_parcel $P > $O/ncu_full_p.log 2>&1
summarise r2_full_parcel
ls -la $O/ | tail -30
for f in a b t p; do tail -n 2 $O/ncu_full_$f.log; done
du -sh $O
