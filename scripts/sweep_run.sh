source scripts/sweep.sh
run greedy_collapse HM_X=1
run dp_collapse HM_BVH_COLLAPSE=dp
run dp_collapse_again HM_BVH_COLLAPSE=dp
