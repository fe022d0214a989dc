source scripts/sweep.sh
run stage_split HM_X=1
