#!/bin/bash
# Round-2 ncu evidence (one gpurun call; each ncu command runs only after the same command exited 0 without ncu).
set -u
O=gpurun_out
A="python tools/ncu_target.py 16384 64 2000 0"      # config 2, default (SIMT) path
B2="python tools/ncu_target.py 16384 64 2000 2"     # same, SA1 second layer on tcgen05 (TF32)
T="python tools/ncu_target_train.py 32 10000"       # config 3 step
P="python tools/ncu_target_parcel.py 200"           # parcel tiling on a 200 m x 200 m cloud (1.3 M points, 256 plots)
$A > $O/ncu_plain_a.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_cfg2.csv $A > /dev/null 2>&1
$A > $O/ncu_plain_a.log 2>&1 && ncu --set full --clock-control none --import-source on -s 27 -c 27 -f -o $O/r2_full_cfg2 $A > $O/ncu_full_a.log 2>&1
$B2 > $O/ncu_plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sa1_tc_kernel -s 1 -c 1 -f -o $O/r2_full_sa1tc $B2 > $O/ncu_full_b.log 2>&1
$T > $O/ncu_plain_t.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"lrb_|head_|pointwise|adam_|bn_|segment_max|edge_msg|interp|kde_lut|project_plotwise|ball_|knn3" -s 62 -c 70 -f -o $O/r2_full_train $T > $O/ncu_full_t.log 2>&1
$P > $O/ncu_plain_p.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"extract_plots|pgrid|finalize_bands|hardveg" -s 7 -c 7 -f -o $O/r2_full_parcel $P > $O/ncu_full_p.log 2>&1
ls -la $O/*.ncu-rep; tail -2 $O/ncu_full_a.log $O/ncu_full_b.log $O/ncu_full_t.log $O/ncu_full_p.log
