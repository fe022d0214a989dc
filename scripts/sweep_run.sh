source scripts/sweep.sh
run rep2 HM_X=1
run rep1 HM_LIB=$V/libhairmsnn_rep1.so
run rep3 HM_LIB=$V/libhairmsnn_rep3.so
run rep2n16 HM_LIB=$V/libhairmsnn_rep2n16.so
run rep4n20 HM_LIB=$V/libhairmsnn_rep4n20.so
