source scripts/sweep.sh
run nearest_first HM_X=1
run nearest_first_again HM_X=1
