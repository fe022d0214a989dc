source scripts/sweep.sh
run group_cull HM_X=1
run group_cull_again HM_X=1
