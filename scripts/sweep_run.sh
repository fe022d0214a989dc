source scripts/sweep.sh
run refill8 HM_X=1
run refill4 HM_LIB=$V/libhairmsnn_refill4.so
run refill2 HM_LIB=$V/libhairmsnn_refill2.so
run refill12 HM_LIB=$V/libhairmsnn_refill12.so
run refill16 HM_LIB=$V/libhairmsnn_refill16.so
