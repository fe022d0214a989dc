source scripts/sweep.sh
run outlined HM_X=1
run outlined_again HM_X=1
